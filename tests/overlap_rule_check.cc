#include <cstdio>
#include <cstdlib>
#include "overlap_rule.h"
int main() {
    long bad = 0;
    // exhaustive per lane: every (va, vb, base a, base b), in every lane position, with the other lanes random
    uint64_t rng = 88172645463325252ull;
    auto rnd = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (uint32_t)(rng >> 16); };
    for (int lane = 0; lane < 4; ++lane)
        for (uint32_t a = 0; a < 256; ++a)
            for (uint32_t b = 0; b < 256; ++b)
                for (uint32_t ba = 0; ba < 4; ++ba)
                    for (uint32_t bb = 0; bb < 4; ++bb) {
                        uint32_t va = rnd(), vb = rnd(), sa = rnd() & 0xff, sb = rnd() & 0xff, m = 0;
                        for (int k = 0; k < 4; ++k) if (rnd() & 1) m |= 0xffu << (8 * k);
                        m |= 0xffu << (8 * lane);
                        va = (va & ~(0xffu << (8 * lane))) | a << (8 * lane);
                        vb = (vb & ~(0xffu << (8 * lane))) | b << (8 * lane);
                        sa = (sa & ~(3u << (2 * lane))) | ba << (2 * lane);
                        sb = (sb & ~(3u << (2 * lane))) | bb << (2 * lane);
                        uint32_t oa, ob;
                        msnv_overlap_rule4(va, vb, msnv_spread_bases(sa), msnv_spread_bases(sb), m, oa, ob);
                        for (int k = 0; k < 4; ++k) {
                            uint32_t xa = (va >> (8 * k)) & 0xff, xb = (vb >> (8 * k)) & 0xff, ea = xa, eb = xb;
                            if ((m >> (8 * k)) & 0xff) {
                                uint32_t ca = (sa >> (2 * k)) & 3, cb = (sb >> (2 * k)) & 3;
                                bool same = ((xa | xb) & 0x80u) ? ((xa & xb & 0x80u) != 0) : (ca == cb);
                                msnv_overlap_rule(xa, xb, same, ea, eb);
                            }
                            if (((oa >> (8 * k)) & 0xff) != ea || ((ob >> (8 * k)) & 0xff) != eb) ++bad;
                        }
                        // the threshold form the kernel uses: same flag bits, and "passes Q13" exactly where the full rule does
                        uint32_t pa, pb;
                        msnv_overlap_pass4(va, vb, msnv_spread_bases(sa ^ sb), m, pa, pb);
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t fa = (oa >> (8 * k)) & 0xff, fb = (ob >> (8 * k)) & 0xff, ga = (pa >> (8 * k)) & 0xff, gb = (pb >> (8 * k)) & 0xff;
                            if ((m >> (8 * k)) & 0xff) {
                                if ((fa & 0x80) != (ga & 0x80) || (fb & 0x80) != (gb & 0x80)) ++bad;
                                if (((fa & 0x7f) >= 13) != ((ga & 0x7f) >= 13) || ((fb & 0x7f) >= 13) != ((gb & 0x7f) >= 13)) ++bad;
                            } else if (fa != ga || fb != gb) ++bad;
                        }
                        // the verdicts alone (what mate_kernel stores): bit 7 of a lane = "passes Q13" of the threshold form
                        uint32_t qa7, qb7;
                        msnv_overlap_verdict4(va, vb, msnv_spread_bases(sa ^ sb), qa7, qb7);
                        for (int k = 0; k < 4; ++k)
                            if ((m >> (8 * k)) & 0xff) {
                                if (((qa7 >> (8 * k + 7)) & 1u) != ((((pa >> (8 * k)) & 0x7f) >= 13) ? 1u : 0u)) ++bad;
                                if (((qb7 >> (8 * k + 7)) & 1u) != ((((pb >> (8 * k)) & 0x7f) >= 13) ? 1u : 0u)) ++bad;
                            }
                        if ((qa7 | qb7) & 0x7f7f7f7fu) ++bad;
                    }
    for (uint32_t q = 0; q < 256; ++q) if (msnv_q08(q) != (uint32_t)(0.8 * (double)q)) ++bad;
    for (int p0 = 0; p0 < 16; p0 += 4) for (int lo = 0; lo < 20; ++lo) for (int hi = lo + 1; hi < 24; ++hi) {
        if (hi <= p0 || lo >= p0 + 4) continue;
        uint32_t m = msnv_quad_mask(p0, lo, hi), e = 0;
        for (int k = 0; k < 4; ++k) if (p0 + k >= lo && p0 + k < hi) e |= 0xffu << (8 * k);
        if (m != e) ++bad;
    }
    printf("mismatches: %ld\n", bad);
    return bad != 0;
}
