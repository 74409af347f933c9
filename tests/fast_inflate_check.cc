// fast_inflate_check -- csrc/host/fast_inflate.cc against zlib: every block type (stored, fixed, dynamic), sizes from 0
// to 64 KiB, data from incompressible to highly repetitive, every compression level and strategy; truncated and
// bit-flipped streams must be refused or, if they still decode, decode to what zlib says (never write out of bounds).
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fast_inflate.hpp"

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 16); }

static std::vector<uint8_t> make_data(int kind, size_t n)
{
    std::vector<uint8_t> d(n);
    switch (kind) {
        case 0: for (auto& x : d) x = (uint8_t)rnd(); break;                                  // incompressible
        case 1: for (auto& x : d) x = "ACGT"[rnd() & 3]; break;                               // 2 bits of entropy
        case 2: for (size_t i = 0; i < n; ++i) d[i] = (uint8_t)("IIIIIIIIFFFF<<<7"[rnd() & 15]); break;   // quality-like
        case 3: for (size_t i = 0; i < n; ++i) d[i] = (uint8_t)(i % 7 == 0 ? rnd() : 'x'); break;          // long matches
        case 4: memset(d.data(), 0, n); break;                                                // one symbol
        default: {                                                                            // BAM-like records
            size_t i = 0;
            while (i < n) {
                const size_t l = 40 + rnd() % 200;
                for (size_t k = 0; k < l && i < n; ++k, ++i) d[i] = (uint8_t)(k < 36 ? (k * 37 + (rnd() & 1)) : (k & 1 ? "ACGT"[rnd() & 3] : 'I' - (rnd() % 5 == 0)));
            }
        }
    }
    return d;
}

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t>& d, int level, int strategy)
{
    z_stream zs; memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
    std::vector<uint8_t> out(deflateBound(&zs, (uLong)d.size()) + 64);
    zs.next_in = (Bytef*)d.data(); zs.avail_in = (uInt)d.size();
    zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}

static bool zlib_inflate(const uint8_t* in, size_t n, std::vector<uint8_t>& out, size_t expect)
{
    z_stream zs; memset(&zs, 0, sizeof zs);
    inflateInit2(&zs, -15);
    out.assign(expect + 1, 0);
    zs.next_in = (Bytef*)in; zs.avail_in = (uInt)n;
    zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    const int r = inflate(&zs, Z_FINISH);
    const bool ok = r == Z_STREAM_END && zs.total_out == expect;
    inflateEnd(&zs);
    out.resize(expect);
    return ok;
}

int main()
{
    long cases = 0, bad = 0, corrupt_accepted = 0;
    const size_t sizes[] = {0, 1, 2, 3, 7, 8, 9, 63, 64, 257, 258, 259, 1000, 4096, 30000, 65279, 65280, 65535, 65536};
    const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE, Z_FILTERED};
    for (int kind = 0; kind < 6; ++kind)
        for (size_t n : sizes)
            for (int level : {0, 1, 4, 6, 9})
                for (int st : strategies) {
                    const std::vector<uint8_t> d = make_data(kind, n), c = deflate_raw(d, level, st);
                    std::vector<uint8_t> guard(n + 64, 0xA5);
                    ++cases;
                    if (!msnv::fast_inflate(c.data(), c.size(), guard.data() + 32, n) || memcmp(guard.data() + 32, d.data(), n)) {
                        ++bad; fprintf(stderr, "MISMATCH kind %d n %zu level %d strategy %d\n", kind, n, level, st);
                    }
                    for (int k = 0; k < 32; ++k) if (guard[k] != 0xA5 || guard[n + 32 + k] != 0xA5) { ++bad; fprintf(stderr, "WROTE OUT OF BOUNDS kind %d n %zu\n", kind, n); break; }
                    // wrong expected size must be refused
                    if (n && msnv::fast_inflate(c.data(), c.size(), guard.data() + 32, n - 1)) { ++bad; fprintf(stderr, "ACCEPTED SHORT OUTPUT kind %d n %zu\n", kind, n); }
                    if (msnv::fast_inflate(c.data(), c.size(), guard.data() + 32, n + 1)) { ++bad; fprintf(stderr, "ACCEPTED LONG OUTPUT kind %d n %zu\n", kind, n); }
                    // truncation and bit flips: refuse, or agree with zlib
                    if (n >= 64 && level && (cases % 3) == 0) {
                        for (int t = 0; t < 6; ++t) {
                            std::vector<uint8_t> cc = c;
                            if (t < 2) cc.resize(cc.size() - 1 - rnd() % (cc.size() / 2));
                            else cc[rnd() % cc.size()] ^= (uint8_t)(1u << (rnd() & 7));
                            std::vector<uint8_t> g2(n + 64, 0x5A), z;
                            const bool mine = msnv::fast_inflate(cc.data(), cc.size(), g2.data() + 32, n);
                            const bool theirs = zlib_inflate(cc.data(), cc.size(), z, n);
                            for (int k = 0; k < 32; ++k) if (g2[k] != 0x5A || g2[n + 32 + k] != 0x5A) { ++bad; fprintf(stderr, "CORRUPT INPUT WROTE OUT OF BOUNDS\n"); break; }
                            if (mine && (!theirs || memcmp(g2.data() + 32, z.data(), n))) { ++corrupt_accepted; ++bad; fprintf(stderr, "ACCEPTED CORRUPT STREAM kind %d n %zu t %d\n", kind, n, t); }
                        }
                    }
                }
    // multi-block streams (Z_FULL_FLUSH between pieces: stored empty blocks in between, non-final blocks of every type)
    for (int rep = 0; rep < 50; ++rep) {
        std::vector<uint8_t> all, comp(200000);
        z_stream zs; memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, 1 + rep % 9, Z_DEFLATED, -15, 8, rep % 5 == 0 ? Z_FIXED : Z_DEFAULT_STRATEGY);
        zs.next_out = comp.data(); zs.avail_out = (uInt)comp.size();
        const int pieces = 1 + rnd() % 6;
        for (int p = 0; p < pieces; ++p) {
            const std::vector<uint8_t> d = make_data((int)(rnd() % 6), rnd() % 9000);
            all.insert(all.end(), d.begin(), d.end());
            zs.next_in = (Bytef*)d.data(); zs.avail_in = (uInt)d.size();
            deflate(&zs, p + 1 == pieces ? Z_FINISH : (rnd() & 1 ? Z_FULL_FLUSH : Z_SYNC_FLUSH));
        }
        comp.resize(zs.total_out);
        deflateEnd(&zs);
        std::vector<uint8_t> o(all.size() + 1);
        ++cases;
        if (!msnv::fast_inflate(comp.data(), comp.size(), o.data(), all.size()) || memcmp(o.data(), all.data(), all.size())) { ++bad; fprintf(stderr, "MISMATCH multi-block %d\n", rep); }
    }
    printf("cases: %ld mismatches: %ld\n", cases, bad);
    return bad != 0;
}
