/* qacompute_oracle.c -- CPU restatement of the reference's qaCompute for the surface metaSNV uses:
 * `qaCompute [-q INT] -c INT -d -i <in.bam> <out>` (metaSNV.py:63-65; src/qaTools/qaCompute.cpp).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md). Pinned: tests/test_oracle_cpu.py compares its two
 * output files byte for byte with the UNMODIFIED reference built into oracle/_ref/qaCompute_ref.
 * It keeps the reference's serial shape (difference array per contig, prefix sum, clamped histogram);
 * BAM decoding goes through the oracle's own reader (obam.c) because htslib is absent.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "obam.h"

static FILE *g_out, *g_detail;
static int g_max_cov = 30, g_silent;

/* compute_print_cov, qaCompute.cpp:125-221 */
static void compute_print_cov(int *data, const char *name, uint32_t chr_size, uint64_t *hist_all)
{
    int32_t cov = 0; uint64_t sum = 0;
    uint64_t *hist = (uint64_t *)calloc(g_max_cov + 1, sizeof(uint64_t));
    for (uint32_t i = 0; i < chr_size; ++i) {                     /* :142-165 */
        cov += data[i];
        sum += (uint64_t)(int64_t)cov;
        int b = cov > g_max_cov ? g_max_cov : cov;
        if (b >= 0) { ++hist_all[b]; ++hist[b]; }                 /* negative coverage is undefined behaviour in the reference */
    }
    if (g_detail) {                                               /* :192-205 */
        fprintf(g_detail, "%s\t%d\t", name, chr_size);
        for (int i = 1; i <= g_max_cov; ++i) {
            uint64_t c = 0;
            for (int x = i; x <= g_max_cov; ++x) c += hist[x];
            fprintf(g_detail, "%d\t", (int)c);
        }
        fprintf(g_detail, "\n");
    }
    if (!g_silent) {
        fprintf(stdout, "Coverage sum %lu ! \n", (unsigned long)sum);
        fprintf(stdout, "Average coverage over %s : %3.2f\n", name, (double)sum / chr_size);
    }
    fprintf(g_out, "%s\t%d\t%3.5f\n", name, chr_size, (double)sum / chr_size);     /* :217 */
    free(hist);
}

/* printSkipped, qaCompute.cpp:226-263 */
static void print_skipped(const obam_hdr *h, int start, int end)
{
    for (int i = start; i < end; ++i) {
        if (!g_silent) {
            printf("Computing %s of size %u... \n", h->target_name[i], h->target_len[i]);
            printf("Coverage sum %d ! \n", 0);
            printf("Average coverage over %s : %3.5f\n", h->target_name[i], 0.0);
        }
        fprintf(g_out, "%s\t%d\t%3.5f\n", h->target_name[i], h->target_len[i], 0.0);
        if (g_detail) {
            fprintf(g_detail, "%s\t%d\t", h->target_name[i], h->target_len[i]);
            for (int k = 1; k <= g_max_cov; ++k) fprintf(g_detail, "%d\t", 0);
            fprintf(g_detail, "\n");
        }
    }
}

int main(int argc, char **argv)
{
    int min_qual = 1, do_detail = 0, arg;
    while ((arg = getopt(argc, argv, "mdip:s:q:c:h:x:a:")) >= 0) {        /* :312-354 */
        switch (arg) {
            case 'd': do_detail = 1; break;
            case 'i': g_silent = 1; break;
            case 'q': min_qual = atoi(optarg); break;
            case 'c': g_max_cov = atoi(optarg); break;
            default: fprintf(stderr, "qacompute_oracle: only -q -c -d -i are restated\n"); return -1;
        }
    }
    if (argc - optind != 2) return 1;
    obam_file *fp = obam_open(argv[optind]);
    if (!fp) { fprintf(stderr, "qaCompute: Failed to open file %s\n", argv[optind]); return 1; }
    obam_hdr *head = obam_hdr_read(fp);
    if (!head) return 1;
    if (!(g_out = fopen(argv[optind + 1], "wt"))) return 1;
    if (do_detail) {
        char *fn = (char *)malloc(strlen(argv[optind + 1]) + 16);
        sprintf(fn, "%s.detail", argv[optind + 1]);
        g_detail = fopen(fn, "wt");
        fprintf(stdout, "Printing details in %s!\n", fn);
        free(fn);
    }
    uint64_t total_len = 0;
    for (int i = 0; i < head->n_targets; ++i) total_len += head->target_len[i];
    uint32_t unmapped = 0, zero_q = 0, total = 0, proper = 0, dups = 0, used = 0, chr_size = 0;
    int *chr = NULL; int32_t cur = -1;
    uint64_t *hist_all = (uint64_t *)calloc(g_max_cov + 1, sizeof(uint64_t));
    fprintf(g_out, "Chromosome\tSeq_lem\tAvg_Cov\n");                    /* :439 */
    obam_rec b; memset(&b, 0, sizeof b);
    while (obam_read1(fp, &b) >= 0) {                                   /* :441-593 */
        if (b.flag & 4) ++unmapped;
        else {
            if (b.tid != cur) {
                if (b.tid == -1) { ++unmapped; ++total; continue; }
                if (cur != -1) {
                    if (!g_silent) fprintf(stdout, "Basing coverage on %u reads\n", used);
                    used = 0;
                    compute_print_cov(chr, head->target_name[cur], chr_size, hist_all);
                }
                chr_size = head->target_len[b.tid];
                chr = (int *)realloc(chr, ((size_t)chr_size + 1) * sizeof(int));
                memset(chr, 0, ((size_t)chr_size + 1) * sizeof(int));
                if (cur + 1 != b.tid && cur != -1) print_skipped(head, cur + 1, b.tid);
                if (cur == -1) { cur = b.tid; print_skipped(head, 0, cur); } else cur = b.tid;
                if (!g_silent) printf("Computing %s of size %u... \n", head->target_name[b.tid], chr_size);
            }
            if (b.mapq >= min_qual) {
                if (b.flag & 2) ++proper;
                if (b.flag & 1024) ++dups;
                else {                                                  /* :530-552 */
                    const uint32_t *cig = obam_cigar(&b);
                    uint32_t pp = (uint32_t)b.pos + 1;
                    int i = 0;
                    if (b.n_cigar > 0 && ((cig[0] & 0xf) == 4 || (cig[0] & 0xf) == 5)) { ++cig; ++i; }
                    while (i < b.n_cigar) {
                        ++i;
                        if ((*cig & 0xf) != 0) pp += *cig >> 4;
                        else {
                            if (pp <= chr_size) ++chr[pp];              /* beyond the array the reference writes out of bounds */
                            pp += *cig >> 4;
                            if (pp >= chr_size) --chr[chr_size - 1]; else --chr[pp];
                        }
                        ++cig;
                    }
                    ++used;
                }
            } else ++zero_q;
        }
        ++total;
    }
    if (cur != -1) compute_print_cov(chr, head->target_name[cur], chr_size, hist_all);       /* :596 */
    if (cur != head->n_targets) print_skipped(head, cur + 1, head->n_targets);               /* :600-602 */
    fprintf(g_out, "\nCov*X\tPercentage\tNr. of bases\n");                                   /* :623-640 */
    for (int i = 1; i <= g_max_cov; ++i) {
        uint64_t c = 0;
        for (int x = i; x <= g_max_cov; ++x) c += hist_all[x];
        fprintf(g_out, "%d\t%3.5f\t%lu\n", i, (double)c / total_len * 100, (unsigned long)c);
    }
    fprintf(g_out, "\nOther\n");                                                             /* :642-654 */
    fprintf(g_out, "Total number of reads: %u\n", total);
    fprintf(g_out, "Total number of duplicates found and ignored: %u\n", dups);
    fprintf(g_out, "Percentage of unmapped reads: %3.5f\n", 100 * ((double)unmapped / total));
    fprintf(g_out, "Percentage of sub-par quality mappings: %3.5f\n", 100 * ((double)zero_q / total));
    int32_t pairs = total / 2;
    fprintf(g_out, "Number of proper paired reads: %u\n", proper);
    fprintf(g_out, "Percentage of proper pairs: %3.5f\n", (double)(100 * (double)proper / 2) / pairs);
    fclose(g_out);
    if (g_detail) fclose(g_detail);
    return 0;
}
