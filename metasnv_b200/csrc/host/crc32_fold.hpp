// crc32_fold.hpp -- CRC-32 (IEEE 802.3, as in gzip/BGZF) of one BGZF member.
// On x86-64 with PCLMULQDQ the bulk is folded 64 bytes at a time with carry-less multiplies (the method of Gopal et
// al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ Instruction", Intel 2009). This file's own
// statement of it: the fold constants x^(D+32), x^(D-32) mod P (D = 512, 128) are computed at start-up from
// P = 0x104C11DB7, the lanes live in an array, and the final 16-byte lane, the tail and every other CPU go through
// zlib's crc32() - there is no hand-written reduction step. After inflation the CRC is ~15 % of the BGZF reader's time
// with zlib's table-driven routine (2 GB/s); folded it is noise.
#pragma once
#include <cstddef>
#include <cstdint>

namespace msnv {
uint32_t crc32_member(const uint8_t* data, size_t n);
}
