// synth.cc -- see synth.hpp.
#include "synth.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

#include "bam.hpp"

namespace msnv {

using namespace synth;

static Model default_model(uint64_t seed, int n_samples)
{
    Model m;
    m.seed = seed; m.n_samples = n_samples; m.read_len = 100;
    m.depth_x100 = 1000; m.presence_ppm = 1000000; m.paired_pct = 50;
    m.site_ppm = 5000; m.err_ppm = 2000; m.nbase_ppm = 500; m.refn_ppm = 100;
    m.indel_pct_x10 = 10; m.clip_pct_x10 = 20; m.mapq0_pct_x10 = 10;
    return m;
}

static uint32_t scaled(double len, double scale, uint32_t floor_len = 2000)
{
    double v = std::floor(len * scale);
    return (uint32_t)(v < floor_len ? floor_len : v);
}

bool synth_preset(const std::string& name, double scale, int n_samples, uint64_t seed, SynthConfig& cfg, std::string& err)
{
    cfg = SynthConfig();
    if (scale <= 0) scale = 1.0;
    if (name == "c1") {                     // tutorial shape: 3 genomes, 1/2/3 subspecies, 160 samples
        int S = n_samples > 0 ? n_samples : 160;
        cfg.model = default_model(seed ? seed : 20211124, S);
        const double lens[3] = {200000, 250000, 300000};
        for (int g = 0; g < 3; ++g) {
            SynthGenome G; G.taxid = 100001 + g; G.n_sub = g + 1;
            G.contig_lens.push_back(scaled(lens[g], scale));
            cfg.genomes.push_back(G);
        }
    } else if (name == "c2") {              // one 5 Mb genome, 1000 samples, 10x
        int S = n_samples > 0 ? n_samples : 1000;
        cfg.model = default_model(seed ? seed : 20211125, S);
        SynthGenome G; G.taxid = 200001; G.n_sub = 2;
        G.contig_lens.push_back(scaled(5e6, scale));
        cfg.genomes.push_back(G);
    } else if (name == "c3") {              // ProGenomes2 scale: 1000 genomes x 4 Mb in 50 contigs, 500 samples
        int S = n_samples > 0 ? n_samples : 500;
        cfg.model = default_model(seed ? seed : 20211126, S);
        cfg.model.depth_x100 = 500; cfg.model.presence_ppm = 100000;
        // scale >= 0.01: that fraction of the 1000 genomes at full size (0.125 = one of 8 createOptimumSplit shards);
        // smaller scales (tests): a handful of short genomes
        const bool full = scale >= 0.01;
        int NG = full ? (int)std::max(2.0, std::floor(1000 * scale + 0.5)) : (int)std::max(8.0, std::floor(1000 * std::min(1.0, scale * 50)));
        for (int g = 0; g < NG; ++g) {
            SynthGenome G; G.taxid = 300001 + g; G.n_sub = 1 + g % 3;
            int nc = full ? 50 : 5;
            for (int c = 0; c < nc; ++c) G.contig_lens.push_back(full ? 80000u : scaled(4e6 / nc, std::min(1.0, scale * 20), 1500));
            cfg.genomes.push_back(G);
        }
    } else if (name == "c4") {              // deep coverage: one 3 Mb genome, 20 samples at 2000x + spikes
        int S = n_samples > 0 ? n_samples : 20;
        cfg.model = default_model(seed ? seed : 20211127, S);
        cfg.model.depth_x100 = 200000;
        SynthGenome G; G.taxid = 400001; G.n_sub = 2;
        uint32_t len = scaled(3e6, scale, 4000);
        G.contig_lens.push_back(len);
        cfg.genomes.push_back(G);
        for (int k = 0; k < 10; ++k) {
            SynthSpike sp; sp.tid = 0; sp.len = 200; sp.depth = 12000;
            sp.start = (uint32_t)((uint64_t)len * (2 * k + 1) / 20);
            if (sp.start + 400 < len && sp.start > 200) cfg.spikes.push_back(sp);
        }
    } else if (name == "c5") {              // 50 genomes x 3 Mb, 200 samples, annotation
        int S = n_samples > 0 ? n_samples : 200;
        cfg.model = default_model(seed ? seed : 20211128, S);
        for (int g = 0; g < 50; ++g) {
            SynthGenome G; G.taxid = 500001 + g; G.n_sub = 1 + g % 3;
            G.contig_lens.push_back(scaled(3e6, scale, 3000));
            cfg.genomes.push_back(G);
        }
        cfg.annotation = true;
    } else {
        err = "unknown preset " + name + " (expected c1..c5)";
        return false;
    }
    return true;
}

std::vector<SynthContig> synth_contigs(const SynthConfig& cfg)
{
    std::vector<SynthContig> v;
    for (size_t g = 0; g < cfg.genomes.size(); ++g)
        for (size_t c = 0; c < cfg.genomes[g].contig_lens.size(); ++c) {
            SynthContig k;
            k.name = std::to_string(cfg.genomes[g].taxid) + "." + cfg.project + ".c" + std::to_string(c);
            k.len = cfg.genomes[g].contig_lens[c];
            k.genome = (int)g;
            v.push_back(k);
        }
    return v;
}

namespace {

struct Rec { uint32_t pos; uint8_t kind; uint32_t f; };   // kind 0 = R1/single, 1 = R2, 2 = spike

struct Writer {
    const SynthConfig& cfg;
    const std::vector<SynthContig>& ctgs;
    int sample;
    BamWriter bw;
    SynthStats st;
    std::string seq; std::vector<uint8_t> qual; std::vector<uint32_t> cig;

    Writer(const SynthConfig& c, const std::vector<SynthContig>& k, int s) : cfg(c), ctgs(k), sample(s) {}

    // Fill seq/qual for a read following `ops` from reference position pos.
    void make_bases(uint32_t tid, uint64_t read_id, uint32_t pos, const uint32_t* ops, int n_ops)
    {
        const Model& m = cfg.model;
        const SynthGenome& G = cfg.genomes[ctgs[tid].genome];
        seq.clear(); qual.clear();
        int64_t rp = pos; int j = 0;
        for (int k = 0; k < n_ops; ++k) {
            uint32_t op = ops[k] & 0xf, len = ops[k] >> 4;
            if (op == CIG_M || op == CIG_EQ || op == CIG_X) {
                for (uint32_t i = 0; i < len; ++i, ++j, ++rp) {
                    seq.push_back(read_base(m, sample, ctgs[tid].genome, G.n_sub, tid, read_id, j, rp));
                    qual.push_back(read_qual(m, sample, tid, read_id, j));
                }
            } else if (op == CIG_I || op == CIG_S) {
                for (uint32_t i = 0; i < len; ++i, ++j) {
                    seq.push_back(read_base(m, sample, ctgs[tid].genome, G.n_sub, tid, read_id, j, -1));
                    qual.push_back(read_qual(m, sample, tid, read_id, j));
                }
            } else if (op == CIG_D || op == CIG_N) {
                rp += len;
            }
        }
    }

    void junk_after(uint32_t tid, uint32_t pos, uint64_t read_id)
    {
        const Model& m = cfg.model;
        uint64_t h = h3(m.seed ^ ST_JUNK, ((uint64_t)sample << 32) | tid, read_id, 0);
        if (!chance(h, cfg.junk_pct_x10, 1000)) return;
        uint32_t k = urand(mix64(h), 8);
        uint16_t flag;
        if (k < 2) flag = FLAG_DUP;
        else if (k == 2) flag = FLAG_SECONDARY;
        else if (k == 3) flag = FLAG_QCFAIL;
        else flag = FLAG_PAIRED | FLAG_MUNMAP | FLAG_READ1;      // orphan: paired but not proper
        if (mix64(h ^ 3) & 1) flag |= FLAG_REVERSE;
        uint32_t L = (uint32_t)m.read_len;
        if (pos + L > ctgs[tid].len) return;
        uint32_t op = L << 4 | CIG_M;
        make_bases(tid, read_id | (1ull << 62), pos, &op, 1);
        cig.assign(1, op);
        char nm[64]; snprintf(nm, sizeof nm, "j%d_%u_%llu", sample, tid, (unsigned long long)read_id);
        bw.write((int32_t)tid, (int32_t)pos, 30, flag, nm, cig, seq, qual, -1, -1, 0);
        ++st.junk;
    }

    void write_contig(uint32_t tid)
    {
        const Model& m = cfg.model;
        const SynthContig& C = ctgs[tid];
        std::vector<Rec> recs;
        bool present = sample_has_genome(m, sample, C.genome);
        bool paired = sample_paired(m, sample);
        int32_t D = paired ? sample_mate_offset(m, sample) : 0;
        uint32_t span = frag_span(m, paired, D);
        uint32_t nf = 0;
        if (present && C.len > span) {
            nf = n_fragments(m, C.len, paired);
            for (uint32_t f = 0; f < nf; ++f) {
                uint32_t x = frag_start(m, sample, tid, C.len, span, nf, f);
                recs.push_back(Rec{x, 0, f});
                if (paired) recs.push_back(Rec{x + (uint32_t)D, 1, f});
            }
        }
        uint32_t spike_base = 0;
        for (size_t k = 0; k < cfg.spikes.size(); ++k) {
            const SynthSpike& sp = cfg.spikes[k];
            if (sp.tid != (int)tid || !present) continue;
            uint32_t L = (uint32_t)m.read_len;
            uint32_t lo = sp.start - (L - 1), w = sp.len + L - 1;
            uint32_t n = (uint32_t)((uint64_t)sp.depth * w / L);
            for (uint32_t i = 0; i < n; ++i) {
                uint64_t h = h3(m.seed ^ ST_DEPTH, ((uint64_t)sample << 32) | tid, k, i);
                recs.push_back(Rec{lo + urand(h, w), 2, spike_base + i});
            }
            spike_base += n;
        }
        std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) {
            if (a.pos != b.pos) return a.pos < b.pos;
            if (a.kind != b.kind) return a.kind < b.kind;
            return a.f < b.f;
        });
        char nm[64];
        for (const Rec& r : recs) {
            if (r.kind == 2) {
                uint32_t L = (uint32_t)m.read_len, op = L << 4 | CIG_M;
                uint64_t rid = (1ull << 61) | r.f;
                make_bases(tid, rid, r.pos, &op, 1);
                cig.assign(1, op);
                snprintf(nm, sizeof nm, "k%d_%u_%u", sample, tid, r.f);
                bw.write((int32_t)tid, (int32_t)r.pos, 40, (r.f & 1) ? FLAG_REVERSE : 0, nm, cig, seq, qual, -1, -1, 0);
                ++st.reads; st.aligned_bases += L;
                continue;
            }
            int mate = r.kind;
            ReadShape sh = read_shape(m, sample, tid, r.f, mate, paired);
            uint64_t rid = (uint64_t)r.f * 2 + (uint64_t)mate;
            make_bases(tid, rid, r.pos, sh.ops, sh.n_ops);
            cig.assign(sh.ops, sh.ops + sh.n_ops);
            uint16_t flag = 0; int32_t mtid = -1, mpos = -1, tlen = 0;
            if (paired) {
                ReadShape other = read_shape(m, sample, tid, r.f, 1 - mate, paired);
                uint32_t x = mate == 0 ? r.pos : r.pos - (uint32_t)D;
                int32_t frag_end = std::max((int32_t)x + (mate == 0 ? sh.rlen : other.rlen),
                                            (int32_t)x + D + (mate == 0 ? other.rlen : sh.rlen));
                flag = FLAG_PAIRED | FLAG_PROPER_PAIR | (mate == 0 ? (FLAG_READ1 | FLAG_MREVERSE) : (FLAG_READ2 | FLAG_REVERSE));
                mtid = (int32_t)tid;
                mpos = mate == 0 ? (int32_t)(x + D) : (int32_t)x;
                tlen = mate == 0 ? frag_end - (int32_t)x : -(frag_end - (int32_t)x);
            } else if (sh.reverse) flag |= FLAG_REVERSE;
            snprintf(nm, sizeof nm, "r%d_%u_%u", sample, tid, r.f);
            bw.write((int32_t)tid, (int32_t)r.pos, sh.mapq, flag, nm, cig, seq, qual, mtid, mpos, tlen);
            ++st.reads;
            for (int k = 0; k < sh.n_ops; ++k) if ((sh.ops[k] & 0xf) == CIG_M) st.aligned_bases += sh.ops[k] >> 4;
            junk_after(tid, r.pos, rid);
        }
    }

    void write_unmapped()
    {
        const Model& m = cfg.model;
        uint64_t n = st.reads * cfg.unmapped_pct_x10 / 1000;
        uint32_t L = (uint32_t)m.read_len;
        for (uint64_t i = 0; i < n; ++i) {
            seq.assign(L, 'A'); qual.assign(L, 20);
            for (uint32_t j = 0; j < L; ++j) seq[j] = "ACGT"[mix64(m.seed ^ (i * 131 + j) ^ ((uint64_t)sample << 40)) & 3];
            cig.clear();
            char nm[64]; snprintf(nm, sizeof nm, "u%d_%llu", sample, (unsigned long long)i);
            bw.write(-1, -1, 0, FLAG_UNMAP, nm, cig, seq, qual, -1, -1, 0);
            ++st.unmapped;
        }
    }
};

bool mkdir_p(const std::string& d)
{
    std::string cur;
    for (size_t i = 0; i <= d.size(); ++i) {
        if (i == d.size() || d[i] == '/') {
            if (!cur.empty() && cur != "/") { if (mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) return false; }
        }
        if (i < d.size()) cur.push_back(d[i]);
    }
    return true;
}

void write_annotation(const SynthConfig& cfg, const std::vector<SynthContig>& ctgs, const std::string& path)
{
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return;
    fprintf(f, "gene_id\texternal_id\tsequence_id\ttype\tgene_info\tlength\tstart\tend\tstrand\tstart_codon\tstop_codon\tgc\n");
    long gid = 0;
    for (size_t t = 0; t < ctgs.size(); ++t) {
        int k = 0;
        for (uint32_t s = 1; s + 1000 < ctgs[t].len; s += 1000) {
            uint64_t h = h3(cfg.model.seed, 77, t, s);
            uint32_t start = s + urand(h, 50) + 1, end = start + 899;
            char strand = (mix64(h) & 1) ? '+' : '-';
            ++gid; ++k;
            fprintf(f, "%ld\t%s.%d\t%s\tCDS\t<annotation synthetic>\t%u\t%u\t%u\t%c\t\t\t\n", gid, ctgs[t].name.c_str(), k,
                    ctgs[t].name.c_str(), end - start + 1, start, end, strand);
            if (chance(mix64(h ^ 9), 10, 100)) {               // 10% overlapping gene on either strand
                uint32_t s2 = start + 600, e2 = s2 + 599;
                if (e2 < ctgs[t].len) {
                    ++gid; ++k;
                    fprintf(f, "%ld\t%s.%d\t%s\tCDS\t<annotation synthetic>\t%u\t%u\t%u\t%c\t\t\t\n", gid, ctgs[t].name.c_str(), k,
                            ctgs[t].name.c_str(), e2 - s2 + 1, s2, e2, (mix64(h ^ 11) & 1) ? '+' : '-');
                }
            }
        }
    }
    fclose(f);
}

}  // namespace

bool synth_write(const SynthConfig& cfg, const std::string& dir, int threads, SynthStats& stats, std::string& err)
{
    std::vector<SynthContig> ctgs = synth_contigs(cfg);
    if (!mkdir_p(dir + "/bam")) { err = "cannot create " + dir; return false; }
    {   // reference FASTA, 60 columns
        FILE* f = fopen((dir + "/ref.fa").c_str(), "w");
        if (!f) { err = "cannot write ref.fa"; return false; }
        std::string line;
        for (size_t t = 0; t < ctgs.size(); ++t) {
            fprintf(f, ">%s\n", ctgs[t].name.c_str());
            for (uint32_t p = 0; p < ctgs[t].len; p += 60) {
                line.clear();
                for (uint32_t q = p; q < p + 60 && q < ctgs[t].len; ++q) line.push_back(ref_base(cfg.model, (uint32_t)t, q));
                line.push_back('\n');
                fwrite(line.data(), 1, line.size(), f);
            }
        }
        fclose(f);
    }
    if (cfg.annotation) write_annotation(cfg, ctgs, dir + "/annotation.txt");

    BamHeader hdr;
    for (auto& c : ctgs) { hdr.names.push_back(c.name); hdr.lens.push_back(c.len); }
    hdr.text = make_sam_header_text(hdr.names, hdr.lens);

    int S = cfg.model.n_samples;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > S) nt = S;
    std::atomic<int> next(0);
    std::atomic<bool> failed(false);
    std::vector<SynthStats> per(S);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&]() {
            for (;;) {
                int s = next.fetch_add(1);
                if (s >= S) break;
                Writer w(cfg, ctgs, s);
                char nm[64]; snprintf(nm, sizeof nm, "/bam/s%04d.bam", s);
                if (!w.bw.open(dir + nm, hdr, 1)) { failed = true; continue; }
                for (uint32_t tid = 0; tid < ctgs.size(); ++tid) w.write_contig(tid);
                w.write_unmapped();
                if (!w.bw.close()) failed = true;
                per[s] = w.st;
            }
        });
    for (auto& x : th) x.join();
    if (failed) { err = "failed writing BAM files under " + dir; return false; }
    std::ofstream list(dir + "/all_samples");
    for (int s = 0; s < S; ++s) {
        char nm[64]; snprintf(nm, sizeof nm, "/bam/s%04d.bam", s);
        list << dir << nm << "\n";
        stats.reads += per[s].reads; stats.aligned_bases += per[s].aligned_bases;
        stats.junk += per[s].junk; stats.unmapped += per[s].unmapped;
    }
    return true;
}

bool sam_to_bam(const std::string& sam_path, const std::string& bam_path, std::string& err)
{
    std::ifstream in(sam_path);
    if (!in) { err = "cannot open " + sam_path; return false; }
    BamHeader hdr;
    std::string line;
    std::vector<std::string> body;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        if (line[0] == '@') {
            hdr.text += line + "\n";
            if (line.compare(0, 3, "@SQ") == 0) {
                std::string sn; uint32_t ln = 0;
                std::stringstream ss(line); std::string tok;
                while (std::getline(ss, tok, '\t')) {
                    if (tok.compare(0, 3, "SN:") == 0) sn = tok.substr(3);
                    if (tok.compare(0, 3, "LN:") == 0) ln = (uint32_t)strtoul(tok.c_str() + 3, nullptr, 10);
                }
                hdr.names.push_back(sn); hdr.lens.push_back(ln);
            }
        } else body.push_back(line);
    }
    BamWriter bw;
    if (!bw.open(bam_path, hdr, 1)) { err = "cannot write " + bam_path; return false; }
    for (const std::string& l : body) {
        std::vector<std::string> c;
        std::stringstream ss(l); std::string tok;
        while (std::getline(ss, tok, '\t')) c.push_back(tok);
        if (c.size() < 11) { err = "SAM line with fewer than 11 columns: " + l; return false; }
        int tid = c[2] == "*" ? -1 : hdr.find(c[2]);
        int mtid = c[6] == "*" ? -1 : (c[6] == "=" ? tid : hdr.find(c[6]));
        std::vector<uint32_t> cig;
        if (c[5] != "*") {
            const char* p = c[5].c_str();
            while (*p) {
                char* e; unsigned long n = strtoul(p, &e, 10);
                const char* ops = "MIDNSHP=X"; const char* q = strchr(ops, *e);
                if (!q || !*e) { err = "bad CIGAR " + c[5]; return false; }
                cig.push_back((uint32_t)n << 4 | (uint32_t)(q - ops));
                p = e + 1;
            }
        }
        std::string seq = c[9] == "*" ? std::string() : c[9];
        std::vector<uint8_t> qual;
        if (c[10] != "*") for (char ch : c[10]) qual.push_back((uint8_t)(ch - 33));
        bw.write(tid, atoi(c[3].c_str()) - 1, (uint8_t)atoi(c[4].c_str()), (uint16_t)atoi(c[1].c_str()), c[0], cig, seq, qual,
                 mtid, atoi(c[7].c_str()) - 1, atoi(c[8].c_str()));
    }
    return bw.close();
}

}  // namespace msnv
