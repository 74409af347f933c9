// synth_model.h -- stateless (counter-based) model of a metagenomic read set.
//
// Every property of every read is a pure function of (seed, sample, contig, fragment, mate), so
// the same data can be produced as BAM files on the host (synth.cc: parity and end-to-end runs)
// and, at the full BASELINE.json shapes, directly as structure-of-arrays read batches on the
// device (csrc/gpu: roofline runs) without materialising hundreds of GB of BAM.
// Shapes follow SURVEY.md section 8(d): 100 bp reads, phred {2:5%,12:5%,20:10%,30:30%,37:50%},
// 0.2% substitution errors, 1% reads with a 1-3 bp indel, 2% with a 5-20 bp soft clip, 0.05% N.
//
// Plain C structs and inline functions only (no STL) so that nvcc can use it in device code.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MSNV_HD __host__ __device__ __forceinline__
#else
#define MSNV_HD inline
#endif

namespace msnv {
namespace synth {

MSNV_HD uint64_t mix64(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
MSNV_HD uint64_t h3(uint64_t seed, uint64_t a, uint64_t b, uint64_t c)
{
    return mix64(mix64(mix64(seed ^ (a * 0xd6e8feb86659fd93ull)) ^ (b * 0xa0761d6478bd642full)) ^ (c * 0xe7037ed1a0b428dbull));
}
// uniform integer in [0, n)
MSNV_HD uint32_t urand(uint64_t h, uint32_t n) { return (uint32_t)(((h >> 32) * (uint64_t)n) >> 32); }
// true with probability num/den (den <= 2^20)
MSNV_HD bool chance(uint64_t h, uint32_t num, uint32_t den) { return urand(h, den) < num; }

enum Stream : uint64_t {
    ST_REF = 1, ST_SITE = 2, ST_CLUSTER = 3, ST_PRESENT = 4, ST_PAIRED = 5, ST_FRAG = 6, ST_READ = 7,
    ST_BASE = 8, ST_QUAL = 9, ST_JUNK = 10, ST_DEPTH = 11
};

struct Model {
    uint64_t seed;
    int32_t  n_samples;
    int32_t  read_len;        // L
    uint32_t depth_x100;      // mean depth per present (sample, genome), times 100
    uint32_t presence_ppm;    // probability (ppm) that a sample carries a genome
    uint32_t paired_pct;      // percent of samples that are paired-end
    uint32_t site_ppm;        // SNV sites per million reference positions
    uint32_t err_ppm;         // substitution error rate per base
    uint32_t nbase_ppm;       // rate of N in reads
    uint32_t refn_ppm;        // rate of N in the reference
    uint32_t indel_pct_x10;   // reads (per mille) carrying one 1-3 bp insertion or deletion
    uint32_t clip_pct_x10;    // reads (per mille) carrying one 5-20 bp soft clip
    uint32_t mapq0_pct_x10;   // accepted reads (per mille) with mapq 0
};

MSNV_HD char ref_base(const Model& m, uint32_t ctg, uint32_t p)
{
    uint64_t h = h3(m.seed, ST_REF, ctg, p);
    if (chance(h, m.refn_ppm, 1000000)) return 'N';
    return "ACGT"[(h >> 3) & 3];
}

// SNV site description. kind: 0 none, 1 shared polymorphism (alt at ~30% in every sample),
// 2 subspecies marker (alt in the samples of one cluster), 3 private (alt in one sample only).
struct Site { int kind; int alt; int cluster; int sample; };

MSNV_HD Site site_at(const Model& m, uint32_t ctg, uint32_t p, int refcode /*0..3, or <0 for N*/, int n_sub)
{
    Site s; s.kind = 0; s.alt = 0; s.cluster = 0; s.sample = 0;
    uint64_t h = h3(m.seed, ST_SITE, ctg, p);
    if (refcode < 0 || !chance(h, m.site_ppm, 1000000)) return s;
    uint64_t g = mix64(h);
    uint32_t k = urand(g, 10);
    s.kind = k < 2 ? 1 : (k < 8 ? 2 : 3);
    s.alt = (refcode + 1 + (int)urand(mix64(g), 3)) & 3;
    s.cluster = (int)urand(mix64(g ^ 0x51), (uint32_t)n_sub);
    s.sample = (int)urand(mix64(g ^ 0x77), (uint32_t)m.n_samples);
    return s;
}

MSNV_HD int base_code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

MSNV_HD int sample_cluster(const Model& m, int sample, int genome, int n_sub)
{
    return (int)urand(h3(m.seed, ST_CLUSTER, (uint64_t)sample, (uint64_t)genome), (uint32_t)n_sub);
}
MSNV_HD bool sample_has_genome(const Model& m, int sample, int genome)
{
    return chance(h3(m.seed, ST_PRESENT, (uint64_t)sample, (uint64_t)genome), m.presence_ppm, 1000000);
}
MSNV_HD bool sample_paired(const Model& m, int sample)
{
    return chance(h3(m.seed, ST_PAIRED, (uint64_t)sample, 0), m.paired_pct, 100);
}
// Distance between the two mates' leftmost coordinates; four classes so that some samples have
// overlapping mates (d < L) and others do not.
MSNV_HD int32_t sample_mate_offset(const Model& m, int sample)
{
    uint32_t k = urand(h3(m.seed, ST_PAIRED, (uint64_t)sample, 1), 4);
    int num = k == 0 ? 3 : k == 1 ? 6 : k == 2 ? 15 : 20;
    return (int32_t)(num * m.read_len / 10);
}

// Number of fragments of (sample, contig); one read per fragment for single-end samples, two for
// paired ones.
MSNV_HD uint32_t n_fragments(const Model& m, uint32_t ctg_len, bool paired)
{
    uint64_t bases = (uint64_t)ctg_len * m.depth_x100 / 100;
    uint64_t per_frag = (uint64_t)m.read_len * (paired ? 2 : 1);
    return (uint32_t)(bases / per_frag);
}
// Largest reference span a fragment can have (mate offset + read + longest deletion).
MSNV_HD uint32_t frag_span(const Model& m, bool paired, int32_t mate_off)
{
    return (uint32_t)(m.read_len + 3 + (paired ? mate_off : 0));
}
// Leftmost coordinate of fragment f: bin f of n equal bins over [0, len-span], plus jitter inside
// the bin, hence non-decreasing in f.
MSNV_HD uint32_t frag_start(const Model& m, int sample, uint32_t ctg, uint32_t ctg_len, uint32_t span,
                            uint32_t n_frag, uint32_t f)
{
    uint64_t room = ctg_len > span ? ctg_len - span : 0;
    uint64_t b0 = room * f / n_frag, b1 = room * (f + 1) / n_frag;
    uint32_t w = (uint32_t)(b1 - b0);
    uint64_t h = h3(m.seed ^ ST_FRAG, (uint64_t)sample, ctg, f);
    return (uint32_t)b0 + (w ? urand(h, w) : 0);
}

// CIGAR shape of a read: up to 3 operations.
struct ReadShape {
    int n_ops;
    uint32_t ops[3];      // BAM encoding len<<4 | op
    int32_t rlen;         // reference span
    int32_t lead_clip;    // query bases before the first aligned base
    uint8_t mapq;
    bool reverse;
};

MSNV_HD ReadShape read_shape(const Model& m, int sample, uint32_t ctg, uint32_t f, int mate, bool paired)
{
    ReadShape r;
    uint64_t h = h3(m.seed ^ ST_READ, (uint64_t)sample, ctg, (uint64_t)f * 2 + (uint64_t)mate);
    const int L = m.read_len;
    uint32_t cls = urand(h, 1000);
    uint64_t g = mix64(h);
    r.lead_clip = 0;
    if (cls < m.indel_pct_x10) {
        int len = 1 + (int)urand(g, 3);
        int at = L * 3 / 10 + (int)urand(mix64(g), (uint32_t)(L * 4 / 10));
        bool ins = (mix64(g ^ 5) & 1) != 0;
        r.n_ops = 3;
        if (ins) {
            r.ops[0] = (uint32_t)at << 4 | 0; r.ops[1] = (uint32_t)len << 4 | 1; r.ops[2] = (uint32_t)(L - at - len) << 4 | 0;
            r.rlen = L - len;
        } else {
            r.ops[0] = (uint32_t)at << 4 | 0; r.ops[1] = (uint32_t)len << 4 | 2; r.ops[2] = (uint32_t)(L - at) << 4 | 0;
            r.rlen = L + len;
        }
    } else if (cls < m.indel_pct_x10 + m.clip_pct_x10) {
        int len = 5 + (int)urand(g, 16);
        bool lead = (mix64(g ^ 9) & 1) != 0;
        r.n_ops = 2;
        if (lead) { r.ops[0] = (uint32_t)len << 4 | 4; r.ops[1] = (uint32_t)(L - len) << 4 | 0; r.lead_clip = len; }
        else      { r.ops[0] = (uint32_t)(L - len) << 4 | 0; r.ops[1] = (uint32_t)len << 4 | 4; }
        r.rlen = L - len;
    } else {
        r.n_ops = 1; r.ops[0] = (uint32_t)L << 4 | 0; r.rlen = L;
    }
    r.mapq = chance(mix64(h ^ 0x33), m.mapq0_pct_x10, 1000) ? 0 : (uint8_t)(20 + urand(mix64(h ^ 0x44), 41));
    r.reverse = paired ? (mate == 1) : ((mix64(h ^ 0x55) & 1) != 0);
    return r;
}

MSNV_HD uint8_t read_qual(const Model& m, int sample, uint32_t ctg, uint64_t read_id, int j)
{
    uint32_t u = urand(h3(m.seed ^ ST_QUAL, ((uint64_t)sample << 32) | ctg, read_id, (uint64_t)j), 100);
    return u < 5 ? 2 : u < 10 ? 12 : u < 20 ? 20 : u < 50 ? 30 : 37;
}

// Base of a read at reference position p (for aligned bases) or a random base (for inserted and
// clipped bases, refp < 0). Returns an ASCII letter in "ACGTN".
MSNV_HD char read_base(const Model& m, int sample, int genome, int n_sub, uint32_t ctg, uint64_t read_id, int j,
                       int64_t refp)
{
    uint64_t h = h3(m.seed ^ ST_BASE, ((uint64_t)sample << 32) | ctg, read_id, (uint64_t)j);
    if (chance(h, m.nbase_ppm, 1000000)) return 'N';
    uint64_t g = mix64(h);
    if (refp < 0) return "ACGT"[g & 3];
    char rb = ref_base(m, ctg, (uint32_t)refp);
    int rc = base_code(rb);
    int truth = rc < 0 ? (int)(g & 3) : rc;
    if (rc >= 0) {
        Site s = site_at(m, ctg, (uint32_t)refp, rc, n_sub);
        if (s.kind == 1) { if (chance(mix64(g ^ 1), 30, 100)) truth = s.alt; }
        else if (s.kind == 2) { if (sample_cluster(m, sample, genome, n_sub) == s.cluster && chance(mix64(g ^ 2), 98, 100)) truth = s.alt; }
        else if (s.kind == 3) { if (sample == s.sample && chance(mix64(g ^ 3), 60, 100)) truth = s.alt; }
    }
    if (chance(mix64(g ^ 4), m.err_ppm, 1000000)) truth = (truth + 1 + (int)urand(mix64(g ^ 6), 3)) & 3;
    return "ACGT"[truth];
}

}  // namespace synth
}  // namespace msnv
