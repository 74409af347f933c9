// bam.hpp -- BAM header / record reader and writer on top of bgzf.hpp (SAMv1 spec section 4.2).
// Replaces, for this path, htslib's sam_open / sam_hdr_read / sam_read1 used by the reference
// (qaCompute.cpp:365-368,441) and the BAM decoding inside `samtools mpileup` / `samtools view -H`
// (metaSNV.py:83,160-165).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "bgzf.hpp"

namespace msnv {

enum : uint16_t {
    FLAG_PAIRED = 0x1, FLAG_PROPER_PAIR = 0x2, FLAG_UNMAP = 0x4, FLAG_MUNMAP = 0x8,
    FLAG_REVERSE = 0x10, FLAG_MREVERSE = 0x20, FLAG_READ1 = 0x40, FLAG_READ2 = 0x80,
    FLAG_SECONDARY = 0x100, FLAG_QCFAIL = 0x200, FLAG_DUP = 0x400, FLAG_SUPPLEMENTARY = 0x800
};
enum : uint32_t { CIG_M = 0, CIG_I = 1, CIG_D = 2, CIG_N = 3, CIG_S = 4, CIG_H = 5, CIG_P = 6, CIG_EQ = 7, CIG_X = 8 };

struct BamHeader {
    std::string text;                       // SAM header text (what `samtools view -H` prints)
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
    int find(const std::string& name) const;
};

// Fixed 32-byte core of a BAM alignment, as laid out on disk after block_size (little endian).
struct BamCore {
    int32_t  tid, pos;
    uint8_t  l_read_name, mapq;
    uint16_t bin, n_cigar, flag;
    int32_t  l_seq, mtid, mpos, tlen;
};
static_assert(sizeof(BamCore) == 32, "BamCore must match the on-disk layout");

// View of one record inside the reader's buffer; valid until the next call to next().
struct BamRecord {
    BamCore core;                            // copied out (the buffer gives no alignment guarantee)
    const char*    qname = nullptr;          // NUL terminated, l_read_name bytes
    const uint8_t* cigar = nullptr;          // n_cigar little-endian u32 (may be unaligned)
    const uint8_t* seq = nullptr;            // 4-bit packed, (l_seq+1)/2 bytes
    const uint8_t* qual = nullptr;           // l_seq bytes
    uint32_t cigar_at(int i) const { uint32_t v; __builtin_memcpy(&v, cigar + 4 * i, 4); return v; }
};

class BamReader {
public:
    bool open(const std::string& path, int threads = 1);
    const BamHeader& header() const { return hdr_; }
    // 1 = record delivered, 0 = clean EOF, -1 = error (see error()).
    int next(BamRecord& rec);
    const std::string& error() const { return err_; }
    double inflate_seconds() const { return bg_.inflate_seconds(); }
    uint64_t compressed_size() const { return bg_.compressed_size(); }
    uint64_t compressed_bytes_read() const { return bg_.compressed_bytes_read(); }
    // virtual offset of the record the next call to next() returns / jump to a record boundary
    uint64_t tell() const { return bg_.tell(); }
    bool seek(uint64_t virtual_offset) { if (!bg_.seek(virtual_offset)) { err_ = bg_.error(); return false; } return true; }
private:
    BgzfReader bg_;
    BamHeader hdr_;
    std::string err_;
};

class BamWriter {
public:
    bool open(const std::string& path, const BamHeader& hdr, int level = 1);
    // seq: one ASCII base per element ("=ACMGRSVTWYHKDBN" alphabet, others -> N); qual: raw phred.
    void write(int32_t tid, int32_t pos, uint8_t mapq, uint16_t flag, const std::string& qname,
               const std::vector<uint32_t>& cigar, const std::string& seq, const std::vector<uint8_t>& qual,
               int32_t mtid, int32_t mpos, int32_t tlen);
    bool close();
private:
    BgzfWriter bg_;
    std::vector<uint8_t> rec_;
};

// Where the records of every reference sequence start in a coordinate-sorted BAM: lets a reader that needs a few
// contigs (one genome bin of createOptimumSplit.py:43-60) seek to them instead of inflating the whole file, which
// is what `samtools mpileup -l` - and round 1 of this build - did once per split (metaSNV.py:157-165).
// Sources, in order: "<bam>.tidx" / a caller-given sidecar written by this repository's qaCompute during the coverage
// pass (one scan of every BAM that metaSNV.py runs anyway, metaSNV.py:58-69), else the linear index of a standard
// "<bam>.bai" (SAMv1 5.2). A sidecar records the size of the BAM it was made for and is ignored when that differs.
struct TidIndex {
    static constexpr uint64_t NONE = ~0ull;
    std::vector<uint64_t> first;          // [n_ref] virtual offset of the first record of the tid, NONE = no record
    bool valid() const { return !first.empty(); }
    bool save(const std::string& path, uint64_t bam_bytes) const;
    bool load_sidecar(const std::string& path, uint64_t bam_bytes, size_t n_ref);
    bool load_bai(const std::string& path, size_t n_ref);
    // "<bam>.tidx", `extra` (may be empty), "<bam>.bai", "<bam without .bam>.bai"
    bool find_for(const std::string& bam_path, const std::string& extra, uint64_t bam_bytes, size_t n_ref);
};

// SAM header text for a list of contigs (@HD + @SQ lines), the form `samtools view -H` prints.
std::string make_sam_header_text(const std::vector<std::string>& names, const std::vector<uint32_t>& lens);

// UCSC binning scheme (SAMv1 5.3); only needed so that written BAMs are spec conformant.
int reg2bin(int64_t beg, int64_t end);

}  // namespace msnv
