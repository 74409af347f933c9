// bgzf.cc -- see bgzf.hpp.
#include "bgzf.hpp"
#include "crc32_fold.hpp"
#include "fast_inflate.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace msnv {

static const size_t kBlockPayload = 0xff00;   // uncompressed bytes per member (htslib's choice too)
static const uint8_t kEofMarker[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                        0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// ------------------------------------------------------------------ writer
BgzfWriter::~BgzfWriter() { if (fp_) close(); }

bool BgzfWriter::open(const std::string& path, int level)
{
    fp_ = fopen(path.c_str(), "wb");
    level_ = level;
    ubuf_.clear(); ubuf_.reserve(kBlockPayload);
    cbuf_.resize(1 << 16);
    ok_ = fp_ != nullptr;
    return ok_;
}

void BgzfWriter::flush_block()
{
    if (ubuf_.empty() || !ok_) return;
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level_, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { ok_ = false; return; }
    zs.next_in = ubuf_.data(); zs.avail_in = (uInt)ubuf_.size();
    zs.next_out = cbuf_.data() + 18; zs.avail_out = (uInt)(cbuf_.size() - 18 - 8);
    int zr = deflate(&zs, Z_FINISH);
    deflateEnd(&zs);
    if (zr != Z_STREAM_END) { ok_ = false; return; }
    size_t clen = zs.total_out, total = 18 + clen + 8;
    uint8_t* h = cbuf_.data();
    const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(h, hdr, 16);
    h[16] = (uint8_t)((total - 1) & 0xff); h[17] = (uint8_t)((total - 1) >> 8);
    uint32_t crc = (uint32_t)crc32(crc32(0L, nullptr, 0), ubuf_.data(), (uInt)ubuf_.size());
    uint32_t isz = (uint32_t)ubuf_.size();
    memcpy(h + 18 + clen, &crc, 4);
    memcpy(h + 18 + clen + 4, &isz, 4);
    if (fwrite(h, 1, total, fp_) != total) ok_ = false;
    ubuf_.clear();
}

void BgzfWriter::write(const void* data, size_t n)
{
    const uint8_t* p = (const uint8_t*)data;
    bytes_in_ += n;
    while (n) {
        size_t k = kBlockPayload - ubuf_.size();
        if (k > n) k = n;
        ubuf_.insert(ubuf_.end(), p, p + k);
        p += k; n -= k;
        if (ubuf_.size() == kBlockPayload) flush_block();
    }
}

void BgzfWriter::reserve(size_t n)
{
    if (n <= kBlockPayload && ubuf_.size() + n > kBlockPayload) flush_block();
}

bool BgzfWriter::close()
{
    if (!fp_) return false;
    flush_block();
    if (fwrite(kEofMarker, 1, 28, fp_) != 28) ok_ = false;
    if (fclose(fp_) != 0) ok_ = false;
    fp_ = nullptr;
    return ok_;
}

// ------------------------------------------------------------------ reader
BgzfReader::~BgzfReader() { close(); }

bool BgzfReader::open(const std::string& path, int threads)
{
    close();
    threads_ = threads < 1 ? 1 : threads;
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) { err_ = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd_, &st) != 0) { err_ = "cannot stat " + path; return false; }
    size_ = (uint64_t)st.st_size;
    if (size_) {
        void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m == MAP_FAILED) { err_ = "cannot map " + path; return false; }
        map_ = (const uint8_t*)m;
        madvise(m, size_, MADV_SEQUENTIAL);
    }
    cpos_ = 0; opos_ = olen_ = 0;
    return true;
}

void BgzfReader::close()
{
    if (map_) munmap((void*)map_, size_);
    if (fd_ >= 0) ::close(fd_);
    map_ = nullptr; fd_ = -1; size_ = cpos_ = 0; opos_ = olen_ = 0;
}

namespace {
struct Member { uint64_t coff; uint32_t clen; uint32_t isize; uint32_t crc; uint64_t ooff; };

bool inflate_member(const uint8_t* src, const Member& m, uint8_t* dst)
{
    if (m.isize == 0) return true;
    // the path's own decoder first (fast_inflate.hpp); zlib only for what it refuses. The CRC is checked either way.
    static const bool use_zlib_only = getenv("MSNV_ZLIB_INFLATE") != nullptr;      // A/B switch for tests and timing
    if (!use_zlib_only && fast_inflate(src + m.coff, m.clen, dst + m.ooff, m.isize))
        return crc32_member(dst + m.ooff, m.isize) == m.crc;
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = (Bytef*)(src + m.coff); zs.avail_in = m.clen;
    zs.next_out = dst + m.ooff; zs.avail_out = m.isize;
    int zr = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (zr != Z_STREAM_END || zs.total_out != m.isize) return false;
    return crc32_member(dst + m.ooff, m.isize) == m.crc;
}
}  // namespace

bool BgzfReader::fill()
{
    // many readers are open at once (one per sample), so the single-threaded batch stays small
    const size_t kBatchBytes = threads_ > 1 ? ((size_t)8 << 20) : ((size_t)512 << 10);
    std::vector<Member> ms;
    uint64_t total = 0;
    batch_.clear();
    while (cpos_ < size_ && total < kBatchBytes) {
        if (size_ - cpos_ < 18) { err_ = "truncated BGZF member header"; return false; }
        const uint8_t* h = map_ + cpos_;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { err_ = "not a BGZF member"; return false; }
        uint32_t xlen = h[10] | h[11] << 8;
        if (size_ - cpos_ < 12 + (uint64_t)xlen) { err_ = "truncated BGZF extra field"; return false; }
        int bsize = -1;
        for (uint32_t off = 0; off + 4 <= xlen;) {
            const uint8_t* x = h + 12 + off;
            uint32_t slen = x[2] | x[3] << 8;
            if (off + 4 + slen > xlen) break;                     // sub-field runs past the extra field: malformed
            if (x[0] == 'B' && x[1] == 'C' && slen == 2) bsize = x[4] | x[5] << 8;
            off += 4 + slen;
        }
        if (bsize < 0) { err_ = "BGZF member without BC subfield"; return false; }
        uint64_t msize = (uint64_t)bsize + 1;
        if (msize < 12 + xlen + 8 || cpos_ + msize > size_) { err_ = "truncated BGZF member"; return false; }
        Member m;
        m.coff = cpos_ + 12 + xlen;
        m.clen = (uint32_t)(msize - 12 - xlen - 8);
        memcpy(&m.crc, map_ + cpos_ + msize - 8, 4);
        memcpy(&m.isize, map_ + cpos_ + msize - 4, 4);
        if (m.isize > (1u << 16)) { err_ = "BGZF member larger than 64 KiB"; return false; }
        m.ooff = total;
        batch_.push_back(BatchMember{cpos_, total});
        total += m.isize;
        ms.push_back(m);
        cpos_ += msize;
        cread_ += msize;
    }
    if (ms.empty()) return false;
    auto t0 = std::chrono::steady_clock::now();
    out_.resize(total);
    bool ok = true;
    int nt = threads_;
    if ((size_t)nt > ms.size()) nt = (int)ms.size();
    if (nt <= 1) {
        for (const Member& m : ms) ok = inflate_member(map_, m, out_.data()) && ok;
    } else {
        std::vector<std::thread> th;
        std::vector<char> oks(nt, 1);
        for (int t = 0; t < nt; ++t)
            th.emplace_back([&, t]() {
                for (size_t i = t; i < ms.size(); i += nt)
                    if (!inflate_member(map_, ms[i], out_.data())) oks[t] = 0;
            });
        for (auto& x : th) x.join();
        for (char c : oks) ok = ok && c;
    }
    inflate_s_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!ok) { err_ = "BGZF inflate/CRC failure"; return false; }
    opos_ = 0; olen_ = total;
    if (total == 0) return cpos_ < size_ ? fill() : false;   // only empty members (EOF marker)
    return true;
}

uint64_t BgzfReader::tell() const
{
    if (opos_ >= olen_) return cpos_ << 16;                       // the next byte is the first of the next member
    size_t lo = 0, hi = batch_.size();                            // last member that starts at or before opos_
    while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (batch_[mid].ooff <= opos_) lo = mid; else hi = mid; }
    return batch_[lo].cstart << 16 | (uint64_t)(opos_ - batch_[lo].ooff);
}

bool BgzfReader::seek(uint64_t voff)
{
    const uint64_t c = voff >> 16, u = voff & 0xffff;
    if (c > size_) { err_ = "seek beyond the end of the file"; return false; }
    cpos_ = c; opos_ = olen_ = 0; batch_.clear(); err_.clear();
    if (u == 0) return true;
    if (!fill() || u > olen_) { if (err_.empty()) err_ = "seek into a truncated BGZF member"; return false; }
    opos_ = u;
    return true;
}

long BgzfReader::read(void* dst, size_t n)
{
    uint8_t* d = (uint8_t*)dst;
    size_t got = 0;
    while (got < n) {
        if (opos_ == olen_) {
            if (!fill()) return err_.empty() ? (long)got : -1;
        }
        size_t k = olen_ - opos_;
        if (k > n - got) k = n - got;
        memcpy(d + got, out_.data() + opos_, k);
        opos_ += k; got += k;
    }
    return (long)got;
}

const uint8_t* BgzfReader::fetch(size_t n)
{
    if (olen_ - opos_ >= n) {
        const uint8_t* p = out_.data() + opos_;
        opos_ += n;
        return p;
    }
    // straddles a batch boundary: assemble in the side buffer
    std::vector<uint8_t> tmp(n);
    long r = read(tmp.data(), n);
    if (r != (long)n) return nullptr;
    side_.swap(tmp);
    return side_.data();
}

}  // namespace msnv
