// snpCall -- drop-in for the reference's src/snpCaller/snpCall (call_vC.cpp), GPU backed.
//
// Command line, exit codes and output formats are those of call_vC.cpp:346-416 (getopt string
// "hdab:f:g:i:c:p:t:"), so the unchanged metaSNV.py:166-176 drives it. Two input modes on stdin:
//   * a one-line job descriptor "#MSNV1\t<ref.fa>\t<bed or ->\t<bam list>" written by this
//     repository's `samtools mpileup` stand-in: BAMs are decoded here and the pileup itself runs on
//     the GPU (no text is ever rendered);
//   * classic `samtools mpileup` text (any real samtools): parsed on the host into count tiles, then
//     the same GPU call / compaction kernels run.
// There is no CPU calling path: without a CUDA device the program fails with a non-zero status.
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/msnv.h"
#include "../host/pileup_input.hpp"
#include "../host/snp_output.hpp"
#include "../host/text_pileup.hpp"

using namespace msnv;

static int print_usage()
{
    fprintf(stderr, "\n");
    fprintf(stderr, "metaSNV --- metagenomic SNV caller (B200 build)\n\n");
    fprintf(stderr, "Usage:   snpCall [options] <stdin.mpileup> \n");
    fprintf(stderr, "Options: \n");
    fprintf(stderr, "     -f,     faidx indexed reference metagenome \n ");
    fprintf(stderr, "    -g,     gene annotation file [NULL].\n");
    fprintf(stderr, "     -i,     individual SNPs output file [NULL].\n\n");
    fprintf(stderr, "SNP definition: \n");
    fprintf(stderr, "     -c,     minimum coverage (mapped reads) per position [4]\n ");
    fprintf(stderr, "    -p,     minimum non-reference nucleotide allele frequency per position [0.01].\n");
    fprintf(stderr, "     -t,     minimum number of non-reference nucleotides per position [4].\n\n");
    fprintf(stderr, "Note: Expecting samtools mpileup as standard input\n\n");
    return 1;
}

static bool verbose() { static int v = getenv("MSNV_VERBOSE") ? 1 : 0; return v != 0; }
static double g_t0 = 0;
static void stage(const char* what);
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void stage(const char* what) { if (verbose()) fprintf(stderr, "[msnv %8.3f s] %s\n", now_s() - g_t0, what); }

static int pick_device(const std::string& indiv_path)
{
    const int n = msnv_device_count();
    if (n <= 0) return -1;
    if (const char* e = getenv("MSNV_DEVICE")) return atoi(e) % n;
    // metaSNV.py names the per-split outputs "...best_split_<k>" (metaSNV.py:199-207): split k -> GPU k mod n
    size_t p = indiv_path.rfind("best_split_");
    if (p != std::string::npos) return atoi(indiv_path.c_str() + p + 11) % n;
    return 0;
}

struct Job { std::string ref, bed, list; };

static int run_direct(const Job& job, const msnv_call_params& prm, const std::string& fasta_opt, const std::string& genes_opt,
                      FILE* indiv, const std::string& indiv_path)
{
    const double t_start = now_s();
    std::vector<std::string> bams;
    {
        std::ifstream in(job.list);
        if (!in) { fprintf(stderr, "snpCall: cannot open %s\n", job.list.c_str()); return 1; }
        std::string l;
        while (std::getline(in, l)) {
            while (!l.empty() && (l.back() == '\r' || l.back() == ' ')) l.pop_back();
            if (!l.empty()) bams.push_back(l);
        }
    }
    const uint32_t S = (uint32_t)bams.size();
    fprintf(stderr, "Identified %d samples\n", (int)S);
    if (S == 0) return 0;

    std::string err;
    BamHeader hdr;
    { BamReader r; if (!r.open(bams[0])) { fprintf(stderr, "snpCall: %s\n", r.error().c_str()); return 1; } hdr = r.header(); }
    Bed bed; bool has_bed = job.bed != "-";
    if (has_bed && !bed.load(job.bed, err)) { fprintf(stderr, "snpCall: %s\n", err.c_str()); return 1; }
    ShardLayout layout;
    if (!layout.build(hdr, has_bed ? &bed : nullptr, err)) { fprintf(stderr, "snpCall: %s\n", err.c_str()); return 1; }
    Fasta fa;
    if (!fa.load(job.ref, err)) { fprintf(stderr, "snpCall: %s\n", err.c_str()); return 1; }
    std::vector<int64_t> ref_len(hdr.names.size(), -1);
    for (size_t t = 0; t < hdr.names.size(); ++t) { int fi = fa.find(hdr.names[t]); if (fi >= 0) ref_len[t] = (int64_t)fa.seqs[fi].size(); }
    if (layout.n_positions == 0) return 0;
    std::vector<uint8_t> ref = shard_reference(layout, hdr, fa);

    Annotation ann;
    if (!fasta_opt.empty() && !genes_opt.empty()) {
        fprintf(stderr, "Found reference genomes and annotation file.\nLoading Genomes...\n");
        if (!ann.load(genes_opt, fasta_opt, err)) { fprintf(stderr, "%s\n", err.c_str()); return 255; }
        fprintf(stderr, "Genomes loaded!\n");
    }

    stage("reference and annotation loaded");
    const int dev = pick_device(indiv_path);
    msnv_ctx* ctx = nullptr;
    if (dev < 0 || msnv_create(dev, &ctx) != MSNV_OK) {
        fprintf(stderr, "snpCall: no usable CUDA device (%s); this build has no CPU calling path\n", msnv_last_error(ctx));
        msnv_destroy(ctx);
        return 1;
    }
    stage("CUDA context created");
    if (msnv_shard_begin(ctx, S, layout.n_positions, ref.data()) != MSNV_OK) {
        fprintf(stderr, "snpCall: %s\n", msnv_last_error(ctx)); msnv_destroy(ctx); return 1;
    }

    // ---- decode all BAMs (one thread per file at a time), upload each as soon as it is ready
    int n_threads = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("MSNV_THREADS")) n_threads = atoi(e);
    if (n_threads < 1) n_threads = 1;
    int inflate_threads = 1;
    if ((int)S < n_threads) { inflate_threads = n_threads / (int)S; n_threads = (int)S; }
    std::vector<SampleReads> reads(S);
    std::vector<DecodeStats> stats(S);
    std::vector<char> done(S, 0);
    std::atomic<uint32_t> next(0);
    std::atomic<bool> failed(false);
    std::mutex mu; std::string first_err;
    const double t_dec0 = now_s();
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&]() {
            for (;;) {
                uint32_t s = next.fetch_add(1);
                if (s >= S || failed) break;
                std::string e;
                if (!decode_sample_for_pileup(bams[s], layout, ref_len, inflate_threads, reads[s], stats[s], e)) {
                    std::lock_guard<std::mutex> lk(mu);
                    if (first_err.empty()) first_err = e;
                    failed = true;
                }
                std::lock_guard<std::mutex> lk(mu);
                done[s] = 1;
            }
        });
    // the context is single-threaded: uploads happen here, in sample order
    double t_h2d = 0; uint64_t h2d_bytes = 0;
    int rc = 0;
    for (uint32_t s = 0; s < S && !failed; ++s) {
        for (;;) { { std::lock_guard<std::mutex> lk(mu); if (done[s]) break; } std::this_thread::sleep_for(std::chrono::microseconds(200)); }
        if (failed) break;
        const double a = now_s();
        msnv_sample_reads v = reads[s].view();
        if (msnv_shard_add_sample(ctx, s, &v) != MSNV_OK || msnv_shard_sync(ctx) != MSNV_OK) {
            fprintf(stderr, "snpCall: %s\n", msnv_last_error(ctx)); rc = 1; failed = true; break;
        }
        h2d_bytes += reads[s].bytes();
        reads[s] = SampleReads();                          // release the host copy
        t_h2d += now_s() - a;
    }
    for (auto& th : pool) th.join();
    const double t_dec1 = now_s();
    stage("BAMs decoded and uploaded");
    if (failed) {
        if (!first_err.empty()) fprintf(stderr, "snpCall: %s\n", first_err.c_str());
        msnv_destroy(ctx);
        return rc ? rc : 1;
    }
    // the reference consumes the first pileup line without calling it (call_vC.cpp:423-434)
    int64_t first_col = -1;
    DecodeStats tot;
    for (uint32_t s = 0; s < S; ++s) {
        if (stats[s].first_column >= 0 && (first_col < 0 || stats[s].first_column < first_col)) first_col = stats[s].first_column;
        tot.records += stats[s].records; tot.accepted += stats[s].accepted; tot.dropped_by_cap += stats[s].dropped_by_cap;
        tot.aligned_bases += stats[s].aligned_bases; tot.pairs += stats[s].pairs; tot.compressed_bytes += stats[s].compressed_bytes;
        tot.seconds += stats[s].seconds; tot.inflate_seconds += stats[s].inflate_seconds;
    }
    if (first_col >= 0 && msnv_shard_mask_position(ctx, (uint32_t)first_col) != MSNV_OK) {
        fprintf(stderr, "snpCall: %s\n", msnv_last_error(ctx)); msnv_destroy(ctx); return 1;
    }

    msnv_hits hits;
    const double t_run0 = now_s();
    if (msnv_shard_run(ctx, &prm, &hits) != MSNV_OK) { fprintf(stderr, "snpCall: %s\n", msnv_last_error(ctx)); msnv_destroy(ctx); return 1; }
    const double t_run1 = now_s();
    stage("kernels done");

    // test hook: per-sample, per-position A,C,G,T,N counts of the whole shard (msnv_shard_counts)
    if (const char* dump = getenv("MSNV_DUMP_COUNTS")) {
        FILE* f = fopen(dump, "wb");
        FILE* g = fopen((std::string(dump) + ".layout").c_str(), "w");
        if (f && g) {
            std::vector<uint16_t> buf((size_t)layout.n_positions * 5);
            for (uint32_t s = 0; s < S; ++s) {
                if (msnv_shard_counts(ctx, s, 0, layout.n_positions, buf.data()) != MSNV_OK) { fprintf(stderr, "snpCall: %s\n", msnv_last_error(ctx)); break; }
                fwrite(buf.data(), 2, buf.size(), f);
            }
            fprintf(g, "%u\t%u\t%lld\n", S, layout.n_positions, (long long)first_col);
            for (const auto& c : layout.ctgs) fprintf(g, "%s\t%u\t%u\n", hdr.names[c.tid].c_str(), c.offset, c.len);
        }
        if (f) fclose(f);
        if (g) fclose(g);
    }

    HitWriter w;
    w.pop_out = stdout; w.indiv_out = indiv; w.ann = ann.active() ? &ann : nullptr;
    std::vector<HitWriter::Contig> ctgs;
    for (const auto& c : layout.ctgs) ctgs.push_back(HitWriter::Contig{hdr.names[c.tid], c.offset, c.len});
    w.write(hits, HitWriter::shard_locator(ctgs, ref.data()));
    fflush(stdout);
    const double t_end = now_s();

    msnv_timings tm; msnv_get_timings(ctx, &tm);
    if (const char* pj = getenv("MSNV_PERF_JSON")) {
        FILE* f = fopen(pj, "a");
        if (f) {
            fprintf(f,
                    "{\"tool\": \"snpCall\", \"device\": %d, \"samples\": %u, \"positions\": %u, \"records\": %llu, \"reads\": %llu, "
                    "\"aligned_bases\": %llu, \"pairs\": %llu, \"dropped_by_cap\": %llu, \"bam_bytes\": %llu, \"h2d_bytes\": %llu, \"hits\": %u, "
                    "\"decode_threads\": %d, \"decode_wall_s\": %.6f, \"decode_cpu_s\": %.6f, \"inflate_cpu_s\": %.6f, \"h2d_s\": %.6f, "
                    "\"gpu_run_wall_s\": %.6f, \"format_s\": %.6f, \"total_s\": %.6f, "
                    "\"ms_index\": %.4f, \"ms_pileup\": %.4f, \"ms_call\": %.4f, \"ms_compact\": %.4f, \"ms_gather\": %.4f, "
                    "\"items\": %llu, \"launches\": %u}\n",
                    dev, S, layout.n_positions, (unsigned long long)tot.records, (unsigned long long)tot.accepted,
                    (unsigned long long)tot.aligned_bases, (unsigned long long)tot.pairs, (unsigned long long)tot.dropped_by_cap,
                    (unsigned long long)tot.compressed_bytes, (unsigned long long)h2d_bytes, hits.n_hits, n_threads * inflate_threads,
                    t_dec1 - t_dec0, tot.seconds, tot.inflate_seconds, t_h2d, t_run1 - t_run0, t_end - t_run1, t_end - t_start, tm.ms_index,
                    tm.ms_pileup, tm.ms_call, tm.ms_compact, tm.ms_gather, (unsigned long long)tm.n_items, tm.kernel_launches);
            fclose(f);
        }
    }
    stage("output written");
    // the process is about to exit: the driver reclaims the device; freeing every block one by one
    // (each cudaFree is a device-wide synchronisation) would only add seconds
    if (getenv("MSNV_CLEAN_EXIT")) msnv_destroy(ctx);
    return 0;
}

int main(int argc, char** argv)
{
    g_t0 = now_s();
    FILE* individualFile = NULL;
    std::string fasta_opt, genes_opt, indiv_path;
    msnv_call_params prm; prm.min_coverage = 4; prm.calling_threshold = 4; prm.min_fraction = 0.01;
    int c;
    opterr = 0;
    while ((c = getopt(argc, argv, "hdab:f:g:i:c:p:t:")) != -1) switch (c) {
        case 'h': print_usage(); return -1;
        case 'a': break;
        case 'd': break;
        case 'b': break;
        case 'f': { FILE* f = fopen(optarg, "r"); if (!f) { fprintf(stderr, "Cannot open %s\n", optarg); return -1; } fclose(f); fasta_opt = optarg; break; }
        case 'g': { FILE* f = fopen(optarg, "r"); if (!f) { fprintf(stderr, "Cannot open %s\n", optarg); return -1; } fclose(f); genes_opt = optarg; break; }
        case 'i':
            individualFile = fopen(optarg, "w");
            if (!individualFile) { fprintf(stderr, "Cannot open %s\n", optarg); return -1; }
            indiv_path = optarg;
            break;
        case 'c': prm.min_coverage = (int32_t)atol(optarg); break;
        case 'p': prm.min_fraction = atof(optarg); break;
        case 't': prm.calling_threshold = (int32_t)atol(optarg); break;
        case '?':
            if (optopt == 'f') fprintf(stderr, "Option -%c requires a reference file.\n", optopt);
            else if (optopt == 'g') fprintf(stderr, "Option -%c requires an annotation file.\n", optopt);
            else if (optopt == 'i') fprintf(stderr, "Option -%c requires an output filename.\n", optopt);
            else if (isprint(optopt)) fprintf(stderr, "Unknown option `-%c'.\n", optopt);
            else { fprintf(stderr, "Unknown option character `\\x%x'.\n", optopt); return 1; }
            abort();                                   // the reference falls through to abort() (call_vC.cpp:391-409)
        default: abort();
    }
    for (int index = optind; index < argc; index++) { printf("Non-option argument %s\n", argv[index]); return 0; }

    // first line of stdin decides the mode
    std::string first;
    {
        int ch;
        while ((ch = fgetc(stdin)) != EOF) { first.push_back((char)ch); if (ch == '\n') break; }
    }
    int rc;
    if (first.compare(0, 7, "#MSNV1\t") == 0) {
        while (!first.empty() && (first.back() == '\n' || first.back() == '\r')) first.pop_back();
        std::vector<std::string> f; size_t p = 7;
        for (;;) { size_t q = first.find('\t', p); f.push_back(first.substr(p, q == std::string::npos ? q : q - p)); if (q == std::string::npos) break; p = q + 1; }
        if (f.size() != 3) { fprintf(stderr, "snpCall: malformed job descriptor on stdin\n"); return 1; }
        Job job{f[0], f[1], f[2]};
        rc = run_direct(job, prm, fasta_opt, genes_opt, individualFile, indiv_path);
    } else {
        rc = run_text_mode(first, stdin, prm, fasta_opt, genes_opt, individualFile, pick_device(indiv_path));
    }
    if (individualFile) fclose(individualFile);
    fflush(stdout); fflush(stderr);
    stage("exit");
    if (!getenv("MSNV_CLEAN_EXIT")) _exit(rc & 0xff);      // skip static destructors / CUDA runtime teardown
    return rc;
}
