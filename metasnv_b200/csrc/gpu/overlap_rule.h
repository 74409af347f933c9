// overlap_rule.h -- mpileup's mate-overlap quality rule (htslib tweak_overlap_quality, SURVEY.md
// Annex A.2) on the staged quality bytes of include/msnv.h, one position at a time and four
// positions at a time (byte-lane arithmetic). Host + device: tests/test_host_cpu.py compiles this
// header with gcc and checks the four-lane form against the scalar one exhaustively.
//
// A staged quality byte is min(phred,127) with bit 7 set when the base is not A/C/G/T. For one
// reference position that both mates align to, with a = the mate that comes first in the file:
//   bases equal    : qa' = min(qa + qb, 127) (htslib caps at 200; only "q >= 13" is ever used), qb' = 0
//   bases differ   : the better one keeps 0.8 * q (truncated), the other drops to 0; a wins ties
// "equal" for flagged bases means both are flagged (htslib compares the 4-bit codes; a read base
// that is N equals only another N -- IUPAC codes other than N are documented as unsupported).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MSNV_RULE_HD __host__ __device__ __forceinline__
#else
#define MSNV_RULE_HD static inline
#endif

// (int)(0.8 * q) for every q < 256, in integers
MSNV_RULE_HD uint32_t msnv_q08(uint32_t q) { return (q * 205u) >> 8; }

MSNV_RULE_HD void msnv_overlap_rule(uint32_t va, uint32_t vb, bool same, uint32_t& na, uint32_t& nb)
{
    const uint32_t fa = va & 0x80u, fb = vb & 0x80u, qa = va & 0x7fu, qb = vb & 0x7fu;
    if (same) { uint32_t q = qa + qb; if (q > 127u) q = 127u; na = fa | q; nb = fb; }
    else if (qa >= qb) { na = fa | msnv_q08(qa); nb = fb; }
    else { na = fa; nb = fb | msnv_q08(qb); }
}

// 2-bit base codes of one seq2 byte, one per byte lane
MSNV_RULE_HD uint32_t msnv_spread_bases(uint32_t s)
{
    s = (s * 4097u) & 0x000f000fu;        // two 2-bit pairs per half word
    return (s * 65u) & 0x03030303u;
}

// msnv_q08 in every byte lane (q <= 255 per lane)
MSNV_RULE_HD uint32_t msnv_q08x4(uint32_t q)
{
    const uint32_t e = q & 0x00ff00ffu, o = (q >> 8) & 0x00ff00ffu;
    return (((e * 205u) >> 8) & 0x00ff00ffu) | ((o * 205u) & 0xff00ff00u);
}

// Four positions at once. va/vb: quality words of the two mates, xa/xb: their bases (msnv_spread_bases),
// m: 0xff in the byte lanes the rule applies to (the other lanes pass through unchanged).
MSNV_RULE_HD void msnv_overlap_rule4(uint32_t va, uint32_t vb, uint32_t xa, uint32_t xb, uint32_t m, uint32_t& oa, uint32_t& ob)
{
    const uint32_t fa = va & 0x80808080u, fb = vb & 0x80808080u, qa = va & 0x7f7f7f7fu, qb = vb & 0x7f7f7f7fu;
    const uint32_t d = xa ^ xb;
    const uint32_t diff7 = ((d | (d << 1)) << 6) & 0x80808080u;                 // bit 7: the 2-bit codes differ
    const uint32_t any_n = fa | fb;
    const uint32_t same7 = (fa & fb) | (~any_n & ~diff7 & 0x80808080u);         // bit 7: "bases equal"
    const uint32_t sum = qa + qb;                                               // <= 254 per lane: no carry between lanes
    const uint32_t qs = (sum | (((sum & 0x80808080u) >> 7) * 127u)) & 0x7f7f7f7fu;   // min(sum, 127)
    const uint32_t ge7 = ((qa | 0x80808080u) - qb) & 0x80808080u;               // bit 7: qa >= qb (no borrow between lanes)
    const uint32_t ms = (same7 >> 7) * 255u, mg = (ge7 >> 7) * 255u;            // byte masks
    const uint32_t na = fa | (qs & ms) | (msnv_q08x4(qa) & ~ms & mg);
    const uint32_t nb = fb | (msnv_q08x4(qb) & ~ms & ~mg);
    oa = (va & ~m) | (na & m);
    ob = (vb & ~m) | (nb & m);
}

// The same rule reduced to what the pileup consumes: a corrected quality is only ever compared with the threshold 13
// (mpileup -Q default; nothing else reads it), so it is enough to know per base whether it still passes:
//   bases equal  : a passes iff qa + qb >= 13, b never
//   bases differ : the better one (a on ties) passes iff its quality is >= 17 (0.8 * 17 = 13.6 -> 13, 0.8 * 16 -> 12), the other never
// d: XOR of the two mates' bases (msnv_spread_bases of the XOR of their seq2 bytes). In the lanes of m the outputs carry
// the input's flag bit and quality 16 (passes) or 0 (fails); other lanes pass through unchanged.
MSNV_RULE_HD void msnv_overlap_pass4(uint32_t va, uint32_t vb, uint32_t d, uint32_t m, uint32_t& oa, uint32_t& ob)
{
    const uint32_t H = 0x80808080u;
    const uint32_t fa = va & H, fb = vb & H, qa = va & ~H, qb = vb & ~H;
    const uint32_t diff7 = (d | (d << 1)) << 6;                                  // bit 7: the 2-bit codes differ
    const uint32_t same7 = (fa & fb) | ~(fa | fb | diff7);                       // bit 7: "bases equal"
    const uint32_t sum = qa + qb;                                                // <= 254 per lane: no carry between lanes
    const uint32_t s13 = ((sum & ~H) + 0x73737373u) | sum;                       // bit 7: qa + qb >= 13
    const uint32_t ge7 = (qa | H) - qb;                                          // bit 7: qa >= qb (no borrow between lanes)
    const uint32_t a17 = qa + 0x6f6f6f6fu, b17 = qb + 0x6f6f6f6fu;               // bit 7: quality >= 17
    const uint32_t pa = ((same7 & s13) | (~same7 & ge7 & a17)) & H;
    const uint32_t pb = ~same7 & ~ge7 & b17 & H;
    oa = (va & ~m) | ((fa | (pa >> 3)) & m);
    ob = (vb & ~m) | ((fb | (pb >> 3)) & m);
}

// The verdicts alone, for all four lanes: bit 7 of a lane of pa7 / pb7 = the base of mate a / b still passes Q13 after the
// rule (what msnv_overlap_pass4 encodes as quality 16). The caller masks the lanes the rule does not apply to.
MSNV_RULE_HD void msnv_overlap_verdict4(uint32_t va, uint32_t vb, uint32_t d, uint32_t& pa7, uint32_t& pb7)
{
    const uint32_t H = 0x80808080u;
    const uint32_t fa = va & H, fb = vb & H, qa = va & ~H, qb = vb & ~H;
    const uint32_t diff7 = (d | (d << 1)) << 6;                                  // bit 7: the 2-bit codes differ
    const uint32_t same7 = (fa & fb) | ~(fa | fb | diff7);                       // bit 7: "bases equal"
    const uint32_t sum = qa + qb;                                                // <= 254 per lane: no carry between lanes
    const uint32_t s13 = ((sum & ~H) + 0x73737373u) | sum;                       // bit 7: qa + qb >= 13
    const uint32_t ge7 = (qa | H) - qb;                                          // bit 7: qa >= qb (no borrow between lanes)
    const uint32_t a17 = qa + 0x6f6f6f6fu, b17 = qb + 0x6f6f6f6fu;               // bit 7: quality >= 17
    pa7 = ((same7 & s13) | (~same7 & ge7 & a17)) & H;
    pb7 = ~same7 & ~ge7 & b17 & H;
}

// byte-lane mask of the positions p0 .. p0+3 that lie in [lo, hi); the quad must intersect the range
MSNV_RULE_HD uint32_t msnv_quad_mask(int32_t p0, int32_t lo, int32_t hi)
{
    const int32_t k0 = lo - p0 > 0 ? lo - p0 : 0, k1 = hi - p0 < 4 ? hi - p0 : 4;
    return (0xffffffffu << (8 * k0)) & (0xffffffffu >> (8 * (4 - k1)));
}
