/* htslib/khash.h STAND-IN (qaCompute.cpp:27,31-32,456). TEST INFRASTRUCTURE ONLY.
 * qaCompute only declares an unused string-set type and, in the `-a` subsampling mode
 * (never used by metaSNV.py:63-65), two hash helpers. The two helpers are the public
 * X31 string hash and Wang's 32-bit integer mix. */
#ifndef ORACLE_HTSLIB_KHASH_STANDIN_H
#define ORACLE_HTSLIB_KHASH_STANDIN_H
#include <stdint.h>
typedef uint32_t khint_t;
#define KHASH_SET_INIT_STR(name) typedef struct kh_##name##_s { int unused; } kh_##name##_t;
#define khash_t(name) kh_##name##_t
static inline khint_t __ac_X31_hash_string(const char *s)
{
    khint_t h = (khint_t)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (khint_t)*s;
    return h;
}
static inline khint_t __ac_Wang_hash(khint_t key)
{
    key += ~(key << 15); key ^= (key >> 10); key += (key << 3);
    key ^= (key >> 6);   key += ~(key << 11); key ^= (key >> 16);
    return key;
}
#endif
