// fast_inflate.cc -- see fast_inflate.hpp. Written from RFC 1951; no code of zlib / libdeflate / htslib.
#include "fast_inflate.hpp"

#include <cstring>

namespace msnv {
namespace {

// ---- table entries -------------------------------------------------------------------------------
// bits 0..4   code bits to consume (for a link: the primary table's width)
// bits 5..8   extra bits that follow the code (lengths: 0..5, distances: 0..13; for a link: subtable index bits)
// bits 9..10  kind
// bit  11     (literals) the entry carries TWO literals: value = first | second << 8, the bit count covers both codes
// bits 16..31 literal byte | length or distance base | 0 = end of block, 1 = invalid symbol | subtable offset
// Bit patterns no code maps to hold INVALID (a "special" entry), so the hot loop tests the kind only.
enum : uint32_t { K_LITERAL = 0, K_BASE = 1, K_SPECIAL = 2, K_LINK = 3 };
constexpr uint32_t entry(uint32_t nbits, uint32_t extra, uint32_t kind, uint32_t value) { return nbits | extra << 5 | kind << 9 | value << 16; }
constexpr uint32_t INVALID = entry(1, 0, K_SPECIAL, 1);
constexpr uint32_t DOUBLE = 1u << 11;
inline uint32_t e_bits(uint32_t e) { return e & 31u; }
inline uint32_t e_extra(uint32_t e) { return (e >> 5) & 15u; }
inline uint32_t e_kind(uint32_t e) { return (e >> 9) & 3u; }
inline uint32_t e_value(uint32_t e) { return e >> 16; }

constexpr int LL_PRIMARY = 11, D_PRIMARY = 8, CL_PRIMARY = 7;
constexpr int LL_SIZE = 2048 + 1024, D_SIZE = 1024;          // primary + subtables; build_table refuses to overflow them

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum Alphabet { LITLEN, DIST, CODELEN };

inline uint32_t symbol_entry(Alphabet a, uint32_t sym, uint32_t nbits)
{
    switch (a) {
        case LITLEN:
            if (sym < 256) return entry(nbits, 0, K_LITERAL, sym);
            if (sym == 256) return entry(nbits, 0, K_SPECIAL, 0);
            if (sym < 286) return entry(nbits, kLenExtra[sym - 257], K_BASE, kLenBase[sym - 257]);
            return entry(nbits, 0, K_SPECIAL, 1);
        case DIST:
            if (sym < 30) return entry(nbits, kDistExtra[sym], K_BASE, kDistBase[sym]);
            return entry(nbits, 0, K_SPECIAL, 1);
        default:
            return entry(nbits, 0, K_LITERAL, sym);
    }
}

struct Reverse8 { uint8_t t[256]; constexpr Reverse8() : t() { for (int i = 0; i < 256; ++i) { int r = 0; for (int k = 0; k < 8; ++k) r |= ((i >> k) & 1) << (7 - k); t[i] = (uint8_t)r; } } };
constexpr Reverse8 kReverse8;
// the low n (<= 15) bits of v in reverse order
inline uint32_t reverse_bits(uint32_t v, int n) { return ((uint32_t)kReverse8.t[v & 0xff] << 8 | kReverse8.t[(v >> 8) & 0xff]) >> (16 - n); }

// Canonical Huffman code of `lens[0..n)` (0 = unused symbol) -> look-up table indexed by the next
// `primary` input bits (LSB first). Over-subscribed codes are refused; incomplete ones are accepted
// (RFC 1951 allows a single distance code; unused patterns stay invalid).
bool build_table(Alphabet a, const uint8_t* lens, int n, int primary, uint32_t* table, int capacity)
{
    int count[16] = {0};
    for (int i = 0; i < n; ++i) ++count[lens[i]];
    count[0] = 0;
    int left = 1;
    for (int l = 1; l <= 15; ++l) { left = left * 2 - count[l]; if (left < 0) return false; }
    uint32_t next[16]; uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }

    const int psize = 1 << primary;
    for (int i = 0; i < psize; ++i) table[i] = INVALID;
    // widest code behind every primary prefix that needs a subtable
    static_assert(LL_PRIMARY >= D_PRIMARY && LL_PRIMARY >= CL_PRIMARY, "scratch arrays are sized for the widest primary table");
    uint8_t sub_bits[1 << LL_PRIMARY];
    memset(sub_bits, 0, (size_t)psize);
    uint32_t codes[288];
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        codes[s] = reverse_bits(next[l]++, l);
        if (l > primary) {
            uint8_t& w = sub_bits[codes[s] & (uint32_t)(psize - 1)];
            if (l - primary > w) w = (uint8_t)(l - primary);
        }
    }
    int used = psize;
    for (int p = 0; p < psize; ++p)
        if (sub_bits[p]) {
            const int size = 1 << sub_bits[p];
            if (used + size > capacity) return false;
            table[p] = entry((uint32_t)primary, sub_bits[p], K_LINK, (uint32_t)used);
            for (int i = 0; i < size; ++i) table[used + i] = INVALID;
            used += size;
        }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        if (l <= primary) {
            const uint32_t e = symbol_entry(a, (uint32_t)s, (uint32_t)l);
            for (uint32_t i = codes[s]; i < (uint32_t)psize; i += 1u << l) table[i] = e;
        } else {
            const uint32_t link = table[codes[s] & (uint32_t)(psize - 1)];
            const uint32_t e = symbol_entry(a, (uint32_t)s, (uint32_t)(l - primary));
            const uint32_t size = 1u << e_extra(link);
            for (uint32_t i = codes[s] >> primary; i < size; i += 1u << (l - primary)) table[e_value(link) + i] = e;
        }
    }
    if (a == LITLEN) {
        // two short literal codes that fit the primary index together become one entry: BAM payload is literal-heavy
        // (qualities, packed bases) with codes of 2-6 bits, and a table look-up per symbol is what bounds the decoder
        uint32_t single[1 << LL_PRIMARY];
        memcpy(single, table, sizeof(uint32_t) * (size_t)psize);
        for (int i = 0; i < psize; ++i) {
            const uint32_t e1 = single[i];
            if (e_kind(e1) != K_LITERAL) continue;
            const uint32_t l1 = e_bits(e1);
            if ((int)l1 >= primary) continue;
            const uint32_t e2 = single[(uint32_t)i >> l1];             // the index bits behind the first code, zero-extended
            if (e_kind(e2) != K_LITERAL || l1 + e_bits(e2) > (uint32_t)primary) continue;
            table[i] = entry(l1 + e_bits(e2), 0, K_LITERAL, e_value(e1) | e_value(e2) << 8) | DOUBLE;
        }
    }
    return true;
}

struct FixedTables {
    uint32_t ll[LL_SIZE], d[D_SIZE];
    bool ok;
    FixedTables()
    {
        uint8_t lens[288];
        for (int i = 0; i < 144; ++i) lens[i] = 8;
        for (int i = 144; i < 256; ++i) lens[i] = 9;
        for (int i = 256; i < 280; ++i) lens[i] = 7;
        for (int i = 280; i < 288; ++i) lens[i] = 8;
        ok = build_table(LITLEN, lens, 288, LL_PRIMARY, ll, LL_SIZE);
        for (int i = 0; i < 32; ++i) lens[i] = 5;
        ok = build_table(DIST, lens, 32, D_PRIMARY, d, D_SIZE) && ok;
    }
};

// ---- bit reader: LSB-first, at most 56..63 valid bits in `buf` ------------------------------------
struct Bits {
    const uint8_t* in; const uint8_t* end;
    uint64_t buf = 0; int cnt = 0;          // bits in `buf`, including `padded` zero bits invented past `end`
    int padded = 0;
    inline void refill()
    {
        if (end - in >= 8) {
            uint64_t w; memcpy(&w, in, 8);                  // little-endian hosts only (x86-64, aarch64)
            buf |= w << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56) {
                if (in < end) buf |= (uint64_t)*in++ << cnt; else padded += 8;
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void consume(int n) { buf >>= n; cnt -= n; }
    inline uint32_t take(int n) { const uint32_t v = peek(n); consume(n); return v; }
    // invented bits were consumed <=> fewer real bits remain than zero bits were appended
    inline bool ran_past_end() const { return cnt < padded; }
};

}  // namespace

bool fast_inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len)
{
    static const FixedTables fixed;
    if (!fixed.ok) return false;
    uint32_t ll_dyn[LL_SIZE], d_dyn[D_SIZE];
    Bits b; b.in = in; b.end = in + in_len;
    uint8_t* const out0 = out; uint8_t* const out_end = out + out_len;

    for (bool last = false; !last;) {
        b.refill();
        last = b.take(1) != 0;
        const uint32_t type = b.take(2);
        if (b.ran_past_end()) return false;
        const uint32_t* ll; const uint32_t* dt;
        if (type == 0) {                                    // stored: back to a byte boundary, LEN, ~LEN, bytes
            b.consume(b.cnt & 7);
            b.refill();
            if (b.ran_past_end() || b.cnt - b.padded < 32) return false;
            const uint32_t len = b.take(16), nlen = b.take(16);
            if ((len ^ nlen) != 0xffffu) return false;
            // hand the whole bytes still in the bit buffer back to the input
            const uint8_t* src = b.in - (b.cnt - b.padded) / 8;
            if ((size_t)(b.end - src) < len || (size_t)(out_end - out) < len) return false;
            memcpy(out, src, len);
            out += len;
            b.in = src + len; b.buf = 0; b.cnt = 0; b.padded = 0;
            continue;
        } else if (type == 1) {
            ll = fixed.ll; dt = fixed.d;
        } else if (type == 2) {
            const uint32_t hlit = b.take(5) + 257, hdist = b.take(5) + 1, hclen = b.take(4) + 4;
            if (hlit > 286 || hdist > 30) return false;
            uint8_t cl[19] = {0};
            b.refill();
            for (uint32_t i = 0; i < hclen; ++i) { if (b.cnt < 3) b.refill(); cl[kClOrder[i]] = (uint8_t)b.take(3); }
            uint32_t clt[1 << CL_PRIMARY];
            if (!build_table(CODELEN, cl, 19, CL_PRIMARY, clt, 1 << CL_PRIMARY)) return false;
            uint8_t lens[286 + 30 + 138];
            uint32_t n = 0;
            while (n < hlit + hdist) {
                b.refill();
                const uint32_t e = clt[b.peek(CL_PRIMARY)];
                if (e_kind(e) != K_LITERAL) return false;
                b.consume((int)e_bits(e));
                const uint32_t sym = e_value(e);
                if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                uint32_t rep, val = 0;
                if (sym == 16) { if (!n) return false; val = lens[n - 1]; rep = 3 + b.take(2); }
                else if (sym == 17) rep = 3 + b.take(3);
                else rep = 11 + b.take(7);
                if (n + rep > hlit + hdist) return false;
                memset(lens + n, (int)val, rep);
                n += rep;
            }
            if (b.ran_past_end() || lens[256] == 0) return false;
            if (!build_table(LITLEN, lens, (int)hlit, LL_PRIMARY, ll_dyn, LL_SIZE)) return false;
            if (!build_table(DIST, lens + hlit, (int)hdist, D_PRIMARY, d_dyn, D_SIZE)) return false;
            ll = ll_dyn; dt = d_dyn;
        } else return false;

        // ---- symbols of one Huffman block
        // Fast loop while there is slack on both sides (no bounds tests inside): up to three literal entries (six
        // literals) per refill of the bit buffer, then at most one match (258 bytes + 8 of copy slack).
        bool block_done = false;
        while (out_end - out >= 6 + 258 + 8 && b.end - b.in >= 8) {
            b.refill();
            uint32_t e = ll[b.peek(LL_PRIMARY)];
            // an entry of the primary table holds one literal or two: both bytes are stored, the pointer moves by 1 or 2
#define MSNV_PUT_LITERALS(e) do { b.consume((int)e_bits(e)); out[0] = (uint8_t)e_value(e); out[1] = (uint8_t)(e_value(e) >> 8); out += 1 + (((e) >> 11) & 1u); } while (0)
            if (e_kind(e) == K_LITERAL) {
                MSNV_PUT_LITERALS(e);
                e = ll[b.peek(LL_PRIMARY)];
                if (e_kind(e) == K_LITERAL) {
                    MSNV_PUT_LITERALS(e);
                    e = ll[b.peek(LL_PRIMARY)];
                    if (e_kind(e) == K_LITERAL) { MSNV_PUT_LITERALS(e); continue; }
                }
                b.refill();                                 // keeps the low bits: `e` still describes the next symbol
            }
#undef MSNV_PUT_LITERALS
            if (e_kind(e) == K_LINK) { b.consume(LL_PRIMARY); e = ll[e_value(e) + b.peek((int)e_extra(e))]; }
            b.consume((int)e_bits(e));
            if (e_kind(e) == K_LITERAL) { *out++ = (uint8_t)e_value(e); continue; }           // (subtable entries are single)
            if (e_kind(e) == K_SPECIAL) {
                if (e_value(e) != 0) return false;
                block_done = true;
                break;
            }
            const uint32_t len = e_value(e) + b.take((int)e_extra(e));
            uint32_t f = dt[b.peek(D_PRIMARY)];
            if (e_kind(f) == K_LINK) { b.consume(D_PRIMARY); f = dt[e_value(f) + b.peek((int)e_extra(f))]; }
            if (e_kind(f) != K_BASE) return false;
            b.consume((int)e_bits(f));
            const uint32_t dist = e_value(f) + b.take((int)e_extra(f));
            if (dist > (size_t)(out - out0)) return false;
            const uint8_t* src = out - dist;
            if (dist >= 8) {
                for (uint32_t i = 0; i < len; i += 8) { uint64_t w; memcpy(&w, src + i, 8); memcpy(out + i, &w, 8); }
            } else if (dist == 1) {
                memset(out, *src, len);
            } else {
                for (uint32_t i = 0; i < len; ++i) out[i] = src[i];
            }
            out += len;
        }
        // Careful loop: the tail of the block, every access checked.
        while (!block_done) {
            b.refill();                                     // >= 56 bits: code 15 + extra 5 + code 15 + extra 13 = 48
            uint32_t e = ll[b.peek(LL_PRIMARY)];
            if (e_kind(e) == K_LINK) { b.consume(LL_PRIMARY); e = ll[e_value(e) + b.peek((int)e_extra(e))]; }
            b.consume((int)e_bits(e));
            if (e_kind(e) == K_LITERAL) {
                if (out >= out_end) return false;
                *out++ = (uint8_t)e_value(e);
                if (e & DOUBLE) {
                    if (out >= out_end) return false;
                    *out++ = (uint8_t)(e_value(e) >> 8);
                }
                continue;
            }
            if (e_kind(e) == K_SPECIAL) {
                if (e_value(e) != 0) return false;
                break;                                      // end of block
            }
            const uint32_t len = e_value(e) + b.take((int)e_extra(e));
            uint32_t f = dt[b.peek(D_PRIMARY)];
            if (e_kind(f) == K_LINK) { b.consume(D_PRIMARY); f = dt[e_value(f) + b.peek((int)e_extra(f))]; }
            if (e_kind(f) != K_BASE) return false;
            b.consume((int)e_bits(f));
            const uint32_t dist = e_value(f) + b.take((int)e_extra(f));
            if (dist > (size_t)(out - out0) || len > (size_t)(out_end - out)) return false;
            const uint8_t* src = out - dist;
            for (uint32_t i = 0; i < len; ++i) out[i] = src[i];
            out += len;
        }
        if (b.ran_past_end()) return false;
    }
    return out == out_end && !b.ran_past_end();
}

}  // namespace msnv
