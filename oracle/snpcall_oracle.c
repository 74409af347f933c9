/* snpcall_oracle.c -- CPU restatement of the reference's snpCall (src/snpCaller/call_vC.cpp + gene.h).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): used by tests/, smoke() and bench.py's CPU legs.
 * Pinned: tests/test_oracle_cpu.py checks it byte for byte against the UNMODIFIED reference
 * compiled into oracle/_ref/snpCall_ref (and against the committed vectors in tests/golden/).
 *
 * Usage is the reference's: snpcall_oracle [-f ref.fa] [-g ann] [-i indiv] [-c INT] [-t INT] [-p FLOAT] < mpileup.txt
 * Each block cites the reference lines it follows.
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define TOK_MAX 10000                     /* call_vC.cpp:482-483 */

/* ---- toksplit, call_vC.cpp:92-111 */
static const char *toksplit(const char *src, char sep, char *tok, size_t lgh)
{
    if (src) {
        while (*src == ' ') src++;
        while (*src && *src != sep) { if (lgh) { *tok++ = *src; --lgh; } src++; }
        if (*src && *src == sep) src++;
    }
    *tok = 0;
    return src;
}

/* ---- annotation: gene blocks per sequence id (call_vC.cpp:116-160) and genomes (:162-195) */
typedef struct { long start, end; char *name; char strand; } gene_t;
typedef struct { char *id; gene_t *genes; int n; } block_t;
typedef struct { char *name; char *seq; long len; } genome_t;
static block_t *g_blocks; static int g_nblocks;
static genome_t *g_genomes; static int g_ngenomes;

static block_t *find_block(const char *id)
{
    block_t *hit = NULL;
    for (int i = 0; i < g_nblocks; ++i) if (!strcmp(g_blocks[i].id, id)) hit = &g_blocks[i];   /* a later block replaces an earlier one (:145,160) */
    return hit;
}
static genome_t *find_genome(const char *name)
{
    genome_t *hit = NULL;
    for (int i = 0; i < g_ngenomes; ++i) if (!strcmp(g_genomes[i].name, name)) hit = &g_genomes[i];
    return hit;
}
/* gene.h:29-37,67: letters outside A,T,C,G,N are stored as 'A' */
static char pack_letter(char c) { return (c == 'A' || c == 'T' || c == 'C' || c == 'G' || c == 'N') ? c : 'A'; }

static void load_annotation(FILE *genomes, FILE *genes)
{
    static char line[10000], tok[TOK_MAX + 1];
    char cur[TOK_MAX + 1] = "";
    int have = 0;
    block_t b; memset(&b, 0, sizeof b);
    if (fgets(line, 10000, genes)) {                                   /* header (:129) */
        while (fgets(line, 10000, genes)) {
            char col[9][TOK_MAX + 1];
            const char *rest = line;
            for (int k = 0; k < 9; ++k) { rest = toksplit(rest, '\t', tok, TOK_MAX); strcpy(col[k], tok); }
            /* the id (column 2) is looked at only when more text follows it (:135-153) */
            const char *q = line; for (int k = 0; k < 3; ++k) q = toksplit(q, '\t', tok, TOK_MAX);
            if (*q) {
                if (!have) { strcpy(cur, col[2]); have = 1; }
                else if (strcmp(cur, col[2])) {
                    b.id = strdup(cur);
                    g_blocks = (block_t *)realloc(g_blocks, (g_nblocks + 1) * sizeof(block_t)); g_blocks[g_nblocks++] = b;
                    memset(&b, 0, sizeof b); strcpy(cur, col[2]);
                }
            }
            /* loadGenome (:243-271): name col 1, start col 6, end col 7 (1-based), strand col 8 */
            b.genes = (gene_t *)realloc(b.genes, (b.n + 1) * sizeof(gene_t));
            b.genes[b.n].name = strdup(col[1]);
            b.genes[b.n].start = atol(col[6]) - 1; b.genes[b.n].end = atol(col[7]) - 1;
            b.genes[b.n].strand = col[8][0];
            ++b.n;
        }
    }
    b.id = strdup(cur);                                               /* "Add the last one!" (:157-160) */
    g_blocks = (block_t *)realloc(g_blocks, (g_nblocks + 1) * sizeof(block_t)); g_blocks[g_nblocks++] = b;

    char *name = strdup(""); char *seq = NULL; long len = 0, cap = 0; int skip = 0;
    while (fgets(line, 10000, genomes)) {
        size_t l = strlen(line);
        if (l) line[l - 1] = 0;                                        /* :170 drops the last character */
        if (line[0] == '>') {
            if (len > 0 && !skip) {                                    /* :173-177 */
                g_genomes = (genome_t *)realloc(g_genomes, (g_ngenomes + 1) * sizeof(genome_t));
                g_genomes[g_ngenomes].name = name; g_genomes[g_ngenomes].seq = seq; g_genomes[g_ngenomes].len = len; ++g_ngenomes;
                seq = NULL; len = cap = 0;
            } else free(name);
            name = strdup(line + 1);
            skip = find_block(name) == NULL;                           /* :180-184 */
        } else if (!skip) {
            l = strlen(line);
            if (len + (long)l + 1 > cap) { cap = (len + l + 1) * 2; seq = (char *)realloc(seq, cap); }
            for (size_t i = 0; i < l; ++i) seq[len++] = pack_letter(line[i]);
        }
    }
    g_genomes = (genome_t *)realloc(g_genomes, (g_ngenomes + 1) * sizeof(genome_t));     /* :193 */
    g_genomes[g_ngenomes].name = name; g_genomes[g_ngenomes].seq = seq; g_genomes[g_ngenomes].len = len; ++g_ngenomes;
    fprintf(stderr, "Genomes loaded!\n");
}

/* Genome::getSequence, gene.h:76-90 */
static void get_sequence(const genome_t *g, long start, long end, char *out)
{
    int n = 0;
    if (g && !(end < start) && !(end > g->len))
        for (long i = start; i <= end; ++i) out[n++] = i < g->len ? g->seq[i] : 'A';
    out[n] = 0;
}
/* revComplement, call_vC.cpp:299-314 */
static void rev_complement(char *s)
{
    char t[8]; int n = 0;
    for (int i = (int)strlen(s) - 1; i >= 0; --i) {
        if (s[i] == 'A') t[n++] = 'T'; else if (s[i] == 'T') t[n++] = 'A';
        else if (s[i] == 'C') t[n++] = 'G'; else if (s[i] == 'G') t[n++] = 'C';
    }
    t[n] = 0; strcpy(s, t);
}
/* codon table, gene.h:3-25; 0 for anything that is not one of the 64 codons (std::map default) */
static char codon_aa(const char *c)
{
    static const char *tab[] = {
        "TAA","X","TGA","X","TAG","X","GCT","A","GCC","A","GCA","A","GCG","A","CGT","R","CGC","R","CGA","R","CGG","R","AGA","R","AGG","R",
        "AAT","N","AAC","N","GAT","D","GAC","D","TGT","C","TGC","C","CAA","Q","CAG","Q","GAA","E","GAG","E","GGT","G","GGC","G","GGA","G","GGG","G",
        "CAT","H","CAC","H","ATT","I","ATC","I","ATA","I","TTA","L","TTG","L","CTT","L","CTC","L","CTA","L","CTG","L","AAA","K","AAG","K","ATG","M",
        "TTT","F","TTC","F","CCT","P","CCC","P","CCA","P","CCG","P","TCT","S","TCC","S","TCA","S","TCG","S","AGT","S","AGC","S",
        "ACT","T","ACC","T","ACA","T","ACG","T","TGG","W","TAT","Y","TAC","Y","GTA","V","GTG","V","GTT","V","GTC","V", NULL};
    for (int i = 0; tab[i]; i += 2) if (!strcmp(tab[i], c)) return tab[i + 1][0];
    return 0;
}

int main(int argc, char **argv)
{
    FILE *genomes = NULL, *genes = NULL, *indivf = NULL;
    int min_cov = 4, thr = 4; double frac = 0.01;                       /* call_vC.cpp:26-36 */
    int c;
    opterr = 0;
    while ((c = getopt(argc, argv, "hdab:f:g:i:c:p:t:")) != -1) switch (c) {   /* :346-410 */
        case 'h': return -1;
        case 'a': case 'd': case 'b': break;
        case 'f': if (!(genomes = fopen(optarg, "r"))) { fprintf(stderr, "Cannot open %s\n", optarg); return -1; } break;
        case 'g': if (!(genes = fopen(optarg, "r"))) { fprintf(stderr, "Cannot open %s\n", optarg); return -1; } break;
        case 'i': if (!(indivf = fopen(optarg, "w"))) { fprintf(stderr, "Cannot open %s\n", optarg); return -1; } break;
        case 'c': min_cov = (int)atol(optarg); break;
        case 'p': frac = atof(optarg); break;
        case 't': thr = (int)atol(optarg); break;
        default: abort();
    }
    if (optind < argc) { printf("Non-option argument %s\n", argv[optind]); return 0; }

    char *line = NULL; size_t cap = 0; ssize_t len;
    /* first line: sample count, then dropped (:423-434) */
    len = getline(&line, &cap, stdin);
    unsigned tabs = 0;
    for (ssize_t i = 0; i < len; ++i) if (line[i] == '\t') ++tabs;
    int S = len > 0 ? (int)(tabs + 1 - 3) / 3 : 0;
    fprintf(stderr, "Identified %d samples\n", S);
    if (genomes && genes) {
        fprintf(stderr, "Found reference genomes and annotation file.\nLoading Genomes...\n");
        load_annotation(genomes, genes);
        fclose(genomes);
    }
    if (S < 0) S = 0;
    /* counts[symbol][sample], symbol order . , a c g t A C G T ; index 0 = all samples (:435-444) */
    static const char SYM[] = ".,acgtACGT";
    long *cnt[10];
    for (int k = 0; k < 10; ++k) cnt[k] = (long *)calloc(S + 1, sizeof(long));
    char *tok = (char *)malloc(TOK_MAX + 1);
    char name[TOK_MAX + 1] = "";
    int genome_loaded = 0, has_genes = 0;
    const block_t *blk = NULL;
    int warned = 0;

    while ((len = getline(&line, &cap, stdin)) > 0) {                   /* :466 */
        line[--len] = 0;                                               /* :475 */
        for (int k = 0; k < 10; ++k) memset(cnt[k], 0, (S + 1) * sizeof(long));
        int pos = 0; long lP = 0; char base = 0;
        const char *rest = toksplit(line, '\t', tok, TOK_MAX);
        while (*rest) {                                                /* :490-541 */
            if (pos == 0) { if (strcmp(name, tok)) { strcpy(name, tok); genome_loaded = 0; } }
            else if (pos == 1) lP = atol(tok) - 1;
            else if (pos == 2) base = tok[0];
            else if (pos > 3 && pos % 3 == 1) {
                int i = 0, l = (int)strlen(tok), smp = pos / 3;
                while (i < l) {
                    switch (tok[i]) {
                        case '^': ++i; break;
                        case '+': case '-': {
                            char num[32]; int nn = 0;
                            while (isdigit((unsigned char)tok[++i])) if (nn < 30) num[nn++] = tok[i];
                            num[nn] = 0;
                            i += atoi(num) - 1;
                            break;
                        }
                        case '*': case '$': case 'N': case 'n': break;
                        default: {
                            const char *p = strchr(SYM, tok[i]);
                            if (p && tok[i] && smp <= S) { ++cnt[p - SYM][0]; ++cnt[p - SYM][smp]; }
                            /* any other symbol crashes the reference (SURVEY.md Annex E #16) */
                            break;
                        }
                    }
                    ++i;
                }
            }
            ++pos;
            rest = toksplit(rest, '\t', tok, TOK_MAX);
        }
        long cov = 0, nonref = 0;
        for (int k = 0; k < 10; ++k) cov += cnt[k][0];                  /* :545 */
        for (int k = 2; k < 10; ++k) nonref += cnt[k][0];
        if ((int)cov < min_cov) continue;                               /* :547-552 */
        if ((int)nonref < thr) continue;
        if (!genome_loaded) {                                          /* loadGenome, :205-214 */
            blk = find_block(name);
            has_genes = blk != NULL;
            genome_loaded = 1;
        }
        const gene_t *gene = NULL;                                     /* first gene in file order containing lP (:273-279,567-574) */
        if (blk) for (int i = 0; i < blk->n; ++i)
            if (blk->genes[i].start <= blk->genes[i].end && blk->genes[i].start <= lP && lP <= blk->genes[i].end) { gene = &blk->genes[i]; break; }

        static const int order[4] = {2, 3, 5, 4};                      /* "actg" (:561) as indices into SYM */
        char *out[2]; size_t on[2] = {0, 0}, om[2] = {0, 0}; out[0] = out[1] = NULL;
        int write_pop = 0;
        for (int oi = 0; oi < 4; ++oi) {
            const int lo = order[oi], up = lo + 4;
            if (SYM[lo] == base) continue;                             /* :580 */
            long n = cnt[lo][0] + cnt[up][0];                          /* :583-584 */
            int which = -1;
            if (n >= thr && (double)n >= (int)cov * frac) { which = 0; write_pop = 1; }      /* :588 */
            else for (int s = 1; s <= S; ++s) if ((int)(cnt[lo][s] + cnt[up][s]) >= thr) { which = 1; break; }   /* :593-600 */
            if (which < 0) continue;
            char ann[64] = ".";
            if (has_genes && gene) {                                   /* :604-633 */
                if (gene->start < gene->end) {
                    long cp = (lP - gene->start) % 3, cs = lP - cp;
                    char oldc[8], newc[8];
                    get_sequence(find_genome(name), cs, cs + 2, oldc);
                    strcpy(newc, oldc);
                    if ((size_t)cp < strlen(newc)) newc[cp] = (char)toupper(SYM[lo]);
                    if (gene->strand == '-') { rev_complement(oldc); rev_complement(newc); }
                    snprintf(ann, sizeof ann, "%s[%s-%s]", codon_aa(newc) == codon_aa(oldc) ? "S" : "N", oldc, newc);
                } else { fprintf(stderr, "Will not handle circular genes\n"); continue; }
            }
            size_t need = 64 + 24 * (size_t)(S + 1);
            if (on[which] + need > om[which]) { om[which] = (on[which] + need) * 2; out[which] = (char *)realloc(out[which], om[which]); }
            char *o = out[which] + on[which];
            o += sprintf(o, ",%ld|%c|%s", n, SYM[up], ann);             /* :625-636 */
            for (int s = 1; s <= S; ++s) o += sprintf(o, "|%d", (int)(cnt[lo][s] + cnt[up][s]));
            on[which] = (size_t)(o - out[which]);
        }
        for (int which = 0; which < 2; ++which) {
            if (which == 0 ? !write_pop : on[1] == 0) continue;        /* :641,:653 */
            FILE *f = which == 0 ? stdout : indivf;
            if (!f) {                                                  /* :655-657 */
                if (!warned) fprintf(stderr, "Individual SNPs detected, but no individual output file specified (-i option).\n");
                warned = 1;
                continue;
            }
            fprintf(f, "%s\t%s\t%ld\t%c\t", name, gene ? gene->name : "-", lP + 1, base);      /* :645-651 */
            for (int s = 1; s <= S; ++s) {                             /* getCoverageString, :316-325 */
                long cv = 0; for (int k = 0; k < 10; ++k) cv += cnt[k][s];
                fprintf(f, s == 1 ? "%d" : "|%d", (int)cv);
            }
            fputc('\t', f);
            if (on[which]) fwrite(out[which] + 1, 1, on[which] - 1, f);                        /* drop the leading comma */
            fputc('\n', f);
        }
        free(out[0]); free(out[1]);
    }
    if (indivf) fclose(indivf);
    return 0;
}
