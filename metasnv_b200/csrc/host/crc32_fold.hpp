// crc32_fold.hpp -- CRC-32 (IEEE 802.3, as in gzip/BGZF) of one BGZF member.
// On x86-64 with PCLMULQDQ the bulk is folded 64 bytes at a time with carry-less multiplies (the method of Gopal et
// al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ Instruction", Intel 2009: fold constants
// x^(512+32), x^(512-32), x^(128+32), x^(128-32), x^64 mod P and the Barrett pair for P = 0x104C11DB7, bit-reflected);
// the tail and every other CPU go through zlib's crc32(). After inflation the CRC is ~15 % of the BGZF reader's time
// with zlib's table-driven routine (2 GB/s); folded it is noise.
#pragma once
#include <cstddef>
#include <cstdint>

namespace msnv {
uint32_t crc32_member(const uint8_t* data, size_t n);
}
