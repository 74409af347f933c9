// fast_inflate.hpp -- raw DEFLATE (RFC 1951) decoder for BGZF members: input and output sizes are
// known up front (BSIZE / ISIZE, SAMv1 4.1), a member is at most 64 KiB, and nothing is streamed.
//
// BGZF inflation is what bounds the path from BAM files (60 % of the decoder's CPU
// time is zlib's inflate). This decoder is written for that one job: a 64-bit bit buffer refilled eight
// bytes at a time, one table look-up per symbol (10-bit primary table for literals/lengths, 8-bit for
// distances, subtables for longer codes), word-wide match copies. It decodes exactly `out_len` bytes
// or reports failure; the caller (bgzf.cc) then falls back to zlib, and the member's CRC-32 is checked
// either way. tests/fast_inflate_check.cc compares it with zlib on every block type and on corrupt input.
#pragma once
#include <cstddef>
#include <cstdint>

namespace msnv {

// true iff [in, in + in_len) is a complete raw DEFLATE stream that inflates to exactly out_len bytes
bool fast_inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);

}  // namespace msnv
