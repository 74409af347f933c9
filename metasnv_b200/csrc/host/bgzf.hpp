// bgzf.hpp -- BGZF block reader / writer on zlib (htslib is not assumed; SURVEY.md Annex F).
//
// The reference gets BGZF through htslib (qaCompute.cpp:276,441) and through the external
// `samtools` binary (metaSNV.py:83,160). Neither is available to this build, so the container
// format is implemented here from the SAMv1 specification, section 4.1.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace msnv {

// Sequential BGZF writer. One instance per output file; not thread safe.
class BgzfWriter {
public:
    BgzfWriter() = default;
    ~BgzfWriter();
    bool open(const std::string& path, int level = 1);
    void write(const void* data, size_t n);
    // Ends the current block if it could not take `n` more bytes (keeps BAM records whole when
    // the caller wants block-aligned records; not required by the format).
    void reserve(size_t n);
    bool close();            // flushes and appends the 28-byte EOF marker
    uint64_t bytes_in() const { return bytes_in_; }
private:
    void flush_block();
    FILE* fp_ = nullptr;
    int level_ = 1;
    std::vector<uint8_t> ubuf_;
    std::vector<uint8_t> cbuf_;
    uint64_t bytes_in_ = 0;
    bool ok_ = true;
};

// Streaming BGZF reader. The whole compressed file is mapped; members are located by their BSIZE
// field and inflated in batches, optionally by several threads (members are independent).
class BgzfReader {
public:
    BgzfReader() = default;
    ~BgzfReader();
    BgzfReader(const BgzfReader&) = delete;
    BgzfReader& operator=(const BgzfReader&) = delete;
    // threads: worker threads used to inflate one batch of members (1 = inline).
    bool open(const std::string& path, int threads = 1);
    void close();
    // Reads exactly n bytes unless EOF; returns the number of bytes delivered, or -1 on a
    // corrupt stream.
    long read(void* dst, size_t n);
    // Zero-copy access: pointer to at least n contiguous inflated bytes (assembling across member
    // boundaries into a side buffer when needed). nullptr at EOF/short stream.
    const uint8_t* fetch(size_t n);
    const std::string& error() const { return err_; }
    uint64_t compressed_size() const { return size_; }
    uint64_t compressed_bytes_read() const { return cread_; }     // members actually inflated (less than the file after a seek)
    double inflate_seconds() const { return inflate_s_; }
    // BGZF virtual file offset (SAMv1 4.1.1: compressed offset of the member << 16 | offset inside its data) of
    // the next byte read() would deliver, and repositioning to one.
    uint64_t tell() const;
    bool seek(uint64_t virtual_offset);
private:
    bool fill();             // inflate the next batch; false at EOF or error
    struct BatchMember { uint64_t cstart; uint64_t ooff; };
    std::vector<BatchMember> batch_;  // members of the current batch: where each starts in the file and in out_
    uint64_t cread_ = 0;
    int fd_ = -1;
    const uint8_t* map_ = nullptr;
    uint64_t size_ = 0, cpos_ = 0;
    int threads_ = 1;
    std::vector<uint8_t> out_;        // inflated bytes of the current batch
    size_t opos_ = 0, olen_ = 0;
    std::vector<uint8_t> side_;       // for fetch() across batch boundaries
    std::string err_;
    double inflate_s_ = 0;
};

}  // namespace msnv
