// crc32_fold.cc -- see crc32_fold.hpp
#include "crc32_fold.hpp"

#include <zlib.h>

#if defined(__x86_64__)
#include <immintrin.h>

namespace {

// ---- fold constants, derived at start-up from the polynomial itself
// x^n mod P over GF(2) for the CRC-32 polynomial P = x^32 + 0x04C11DB7 (most significant bit = highest power)
constexpr uint32_t kPolyLow = 0x04C11DB7u;
uint32_t x_pow_mod_p(unsigned n)
{
    uint32_t r = 1;                         // x^0
    for (unsigned i = 0; i < n; ++i) {
        const bool carry = r & 0x80000000u;
        r <<= 1;
        if (carry) r ^= kPolyLow;
    }
    return r;
}
// gzip's CRC is bit-reflected; a carry-less product of two reflected 64-bit operands comes out one bit short of its
// own reflection, which the constant absorbs: constant = reflect32(x^n mod P) << 1
uint64_t fold_constant(unsigned n)
{
    const uint32_t v = x_pow_mod_p(n);
    uint32_t r = 0;
    for (int b = 0; b < 32; ++b) if (v & (1u << b)) r |= 0x80000000u >> b;
    return (uint64_t)r << 1;
}

struct FoldKeys { __m128i by512, by128; };

__attribute__((target("pclmul,sse4.1")))
const FoldKeys& fold_keys()
{
    // moving a 128-bit lane D bits towards the end of the message multiplies its low half by x^(D+32) and its high
    // half by x^(D-32) (all mod P); D = 512 while four lanes run side by side, D = 128 when they are merged
    static const FoldKeys k = {
        _mm_set_epi64x((long long)fold_constant(512 - 32), (long long)fold_constant(512 + 32)),
        _mm_set_epi64x((long long)fold_constant(128 - 32), (long long)fold_constant(128 + 32)),
    };
    return k;
}

// lane * x^D mod P (as a 128-bit value congruent to it), xor the data that sits D bits further on
__attribute__((target("pclmul,sse4.1")))
inline __m128i fold_onto(__m128i lane, __m128i key, __m128i next)
{
    const __m128i lo = _mm_clmulepi64_si128(lane, key, 0x00), hi = _mm_clmulepi64_si128(lane, key, 0x11);
    return _mm_xor_si128(_mm_xor_si128(lo, hi), next);
}

// CRC register after `len` bytes (len >= 64, a multiple of 16), starting from register value `reg`.
// Invariant of the loop: (lanes || unread bytes) has the same remainder as (reg-adjusted message read so far ||
// unread bytes). At the end one 16-byte lane is left; its remainder is taken by zlib as if it were message bytes
// under a zero register, so there is no reduction code here.
__attribute__((target("pclmul,sse4.1")))
uint32_t crc32_fold_bulk(const uint8_t* p, size_t len, uint32_t reg)
{
    const FoldKeys& k = fold_keys();
    __m128i lane[4];
    for (int i = 0; i < 4; ++i) lane[i] = _mm_loadu_si128((const __m128i*)(p + 16 * i));
    lane[0] = _mm_xor_si128(lane[0], _mm_cvtsi32_si128((int)reg));       // the register lines up with the first four bytes
    p += 64; len -= 64;
    for (; len >= 64; p += 64, len -= 64)
        for (int i = 0; i < 4; ++i) lane[i] = fold_onto(lane[i], k.by512, _mm_loadu_si128((const __m128i*)(p + 16 * i)));
    __m128i acc = lane[0];
    for (int i = 1; i < 4; ++i) acc = fold_onto(acc, k.by128, lane[i]);
    for (; len >= 16; p += 16, len -= 16) acc = fold_onto(acc, k.by128, _mm_loadu_si128((const __m128i*)p));
    alignas(16) uint8_t rest[16];
    _mm_store_si128((__m128i*)rest, acc);
    // zlib's crc32(c, buf) is ~remainder(register = ~c, buf): a zero register is c = ~0, and the result is inverted back
    return ~(uint32_t)crc32(0xffffffffu, rest, 16);
}

bool have_clmul() { static const bool ok = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1"); return ok; }

}  // namespace
#endif

namespace msnv {

uint32_t crc32_member(const uint8_t* data, size_t n)
{
#if defined(__x86_64__)
    if (n >= 64 && have_clmul()) {
        const size_t bulk = n & ~(size_t)15;
        const uint32_t reg = crc32_fold_bulk(data, bulk, 0xffffffffu);    // zlib's crc32(0, ...) starts from register ~0
        // hand the register to zlib for the tail: crc32(c, ...) works on ~c
        return (uint32_t)crc32((uLong)(~reg), data + bulk, (uInt)(n - bulk));
    }
#endif
    return (uint32_t)crc32(crc32(0L, nullptr, 0), data, (uInt)n);
}

}  // namespace msnv
