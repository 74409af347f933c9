// bam.cc -- see bam.hpp.
#include "bam.hpp"

#include <unistd.h>

#include <cstdio>
#include <cstring>

namespace msnv {

int BamHeader::find(const std::string& name) const
{
    for (size_t i = 0; i < names.size(); ++i) if (names[i] == name) return (int)i;
    return -1;
}

bool BamReader::open(const std::string& path, int threads)
{
    if (!bg_.open(path, threads)) { err_ = bg_.error(); return false; }
    uint8_t b[8];
    if (bg_.read(b, 8) != 8 || memcmp(b, "BAM\1", 4) != 0) { err_ = path + ": not a BAM file"; return false; }
    uint32_t l_text; memcpy(&l_text, b + 4, 4);
    hdr_.text.resize(l_text);
    if (l_text && bg_.read(&hdr_.text[0], l_text) != (long)l_text) { err_ = path + ": truncated header"; return false; }
    while (!hdr_.text.empty() && hdr_.text.back() == '\0') hdr_.text.pop_back();
    int32_t n_ref;
    if (bg_.read(&n_ref, 4) != 4 || n_ref < 0) { err_ = path + ": truncated header"; return false; }
    hdr_.names.resize(n_ref); hdr_.lens.resize(n_ref);
    for (int i = 0; i < n_ref; ++i) {
        int32_t l_name;
        if (bg_.read(&l_name, 4) != 4 || l_name < 1) { err_ = path + ": truncated header"; return false; }
        std::string nm(l_name, '\0');
        if (bg_.read(&nm[0], l_name) != l_name) { err_ = path + ": truncated header"; return false; }
        nm.resize(strlen(nm.c_str()));
        hdr_.names[i] = nm;
        if (bg_.read(&hdr_.lens[i], 4) != 4) { err_ = path + ": truncated header"; return false; }
    }
    return true;
}

int BamReader::next(BamRecord& rec)
{
    int32_t block_size;
    long r = bg_.read(&block_size, 4);
    if (r == 0) return 0;
    if (r != 4 || block_size < 32) { err_ = "corrupt BAM record"; return -1; }
    const uint8_t* p = bg_.fetch((size_t)block_size);
    if (!p) { err_ = "truncated BAM record"; return -1; }
    memcpy(&rec.core, p, 32);
    const BamCore* c = &rec.core;
    size_t need = 32 + (size_t)c->l_read_name + 4 * (size_t)c->n_cigar + (size_t)((c->l_seq + 1) / 2) + (size_t)c->l_seq;
    if (c->l_seq < 0 || need > (size_t)block_size) { err_ = "corrupt BAM record"; return -1; }
    rec.qname = (const char*)(p + 32);
    rec.cigar = p + 32 + c->l_read_name;
    rec.seq = rec.cigar + 4 * (size_t)c->n_cigar;
    rec.qual = rec.seq + (c->l_seq + 1) / 2;
    return 1;
}

// ---- tid index
namespace {
const char kTidxMagic[8] = {'M', 'S', 'N', 'V', 'T', 'I', 'X', '1'};
}

bool TidIndex::save(const std::string& path, uint64_t bam_bytes) const
{
    const std::string tmp = path + ".tmp" + std::to_string((unsigned long long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    const uint64_t n = first.size();
    bool ok = fwrite(kTidxMagic, 1, 8, f) == 8 && fwrite(&bam_bytes, 8, 1, f) == 1 && fwrite(&n, 8, 1, f) == 1 &&
              (n == 0 || fwrite(first.data(), 8, n, f) == n);
    ok = fclose(f) == 0 && ok;
    if (ok) ok = rename(tmp.c_str(), path.c_str()) == 0;          // concurrent writers: last complete file wins
    if (!ok) remove(tmp.c_str());
    return ok;
}

bool TidIndex::load_sidecar(const std::string& path, uint64_t bam_bytes, size_t n_ref)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char magic[8]; uint64_t size = 0, n = 0;
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, kTidxMagic, 8) == 0 && fread(&size, 8, 1, f) == 1 && fread(&n, 8, 1, f) == 1 &&
              size == bam_bytes && n == n_ref;
    std::vector<uint64_t> v;
    if (ok) { v.resize(n); ok = n == 0 || fread(v.data(), 8, n, f) == n; }
    fclose(f);
    if (ok) first.swap(v);
    return ok;
}

bool TidIndex::load_bai(const std::string& path, size_t n_ref)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    auto rd = [&](void* p, size_t n) { return fread(p, 1, n, f) == n; };
    char magic[4]; int32_t n = 0;
    bool ok = rd(magic, 4) && memcmp(magic, "BAI\1", 4) == 0 && rd(&n, 4) && n >= 0 && (size_t)n == n_ref;
    std::vector<uint64_t> v(ok ? n_ref : 0, NONE);
    for (int32_t r = 0; ok && r < n; ++r) {
        int32_t n_bin = 0;
        ok = rd(&n_bin, 4) && n_bin >= 0;
        uint64_t lo = NONE;
        for (int32_t b = 0; ok && b < n_bin; ++b) {
            uint32_t bin = 0; int32_t n_chunk = 0;
            ok = rd(&bin, 4) && rd(&n_chunk, 4) && n_chunk >= 0;
            for (int32_t c = 0; ok && c < n_chunk; ++c) {
                uint64_t cb = 0, ce = 0;
                ok = rd(&cb, 8) && rd(&ce, 8);
                if (ok && bin != 37450 && cb < lo) lo = cb;           // 37450: the pseudo-bin with the read counts
            }
        }
        int32_t n_intv = 0;
        ok = ok && rd(&n_intv, 4) && n_intv >= 0;
        for (int32_t i = 0; ok && i < n_intv; ++i) { uint64_t io = 0; ok = rd(&io, 8); }
        if (ok) v[r] = lo;
    }
    fclose(f);
    if (ok) first.swap(v);
    return ok;
}

bool TidIndex::find_for(const std::string& bam_path, const std::string& extra, uint64_t bam_bytes, size_t n_ref)
{
    if (load_sidecar(bam_path + ".tidx", bam_bytes, n_ref)) return true;
    if (!extra.empty() && load_sidecar(extra, bam_bytes, n_ref)) return true;
    if (load_bai(bam_path + ".bai", n_ref)) return true;
    if (bam_path.size() > 4 && bam_path.compare(bam_path.size() - 4, 4, ".bam") == 0 && load_bai(bam_path.substr(0, bam_path.size() - 4) + ".bai", n_ref)) return true;
    return false;
}

int reg2bin(int64_t beg, int64_t end)
{
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

std::string make_sam_header_text(const std::vector<std::string>& names, const std::vector<uint32_t>& lens)
{
    std::string t = "@HD\tVN:1.6\tSO:coordinate\n";
    for (size_t i = 0; i < names.size(); ++i)
        t += "@SQ\tSN:" + names[i] + "\tLN:" + std::to_string(lens[i]) + "\n";
    return t;
}

bool BamWriter::open(const std::string& path, const BamHeader& hdr, int level)
{
    if (!bg_.open(path, level)) return false;
    bg_.write("BAM\1", 4);
    uint32_t l_text = (uint32_t)hdr.text.size();
    bg_.write(&l_text, 4);
    bg_.write(hdr.text.data(), l_text);
    int32_t n_ref = (int32_t)hdr.names.size();
    bg_.write(&n_ref, 4);
    for (int i = 0; i < n_ref; ++i) {
        int32_t l_name = (int32_t)hdr.names[i].size() + 1;
        bg_.write(&l_name, 4);
        bg_.write(hdr.names[i].c_str(), l_name);
        bg_.write(&hdr.lens[i], 4);
    }
    return true;
}

static uint8_t base_code(char c)
{
    switch (c) {
        case '=': return 0;  case 'A': case 'a': return 1;  case 'C': case 'c': return 2;
        case 'M': case 'm': return 3;  case 'G': case 'g': return 4;  case 'R': case 'r': return 5;
        case 'S': case 's': return 6;  case 'V': case 'v': return 7;  case 'T': case 't': return 8;
        case 'W': case 'w': return 9;  case 'Y': case 'y': return 10; case 'H': case 'h': return 11;
        case 'K': case 'k': return 12; case 'D': case 'd': return 13; case 'B': case 'b': return 14;
        default: return 15;
    }
}

void BamWriter::write(int32_t tid, int32_t pos, uint8_t mapq, uint16_t flag, const std::string& qname,
                      const std::vector<uint32_t>& cigar, const std::string& seq, const std::vector<uint8_t>& qual,
                      int32_t mtid, int32_t mpos, int32_t tlen)
{
    int64_t rlen = 0;
    for (uint32_t c : cigar) {
        uint32_t op = c & 0xf;
        if (op == CIG_M || op == CIG_D || op == CIG_N || op == CIG_EQ || op == CIG_X) rlen += c >> 4;
    }
    BamCore core;
    core.tid = tid; core.pos = pos;
    core.l_read_name = (uint8_t)(qname.size() + 1); core.mapq = mapq;
    core.bin = (uint16_t)reg2bin(pos, pos + (rlen ? rlen : 1));
    core.n_cigar = (uint16_t)cigar.size(); core.flag = flag;
    core.l_seq = (int32_t)seq.size(); core.mtid = mtid; core.mpos = mpos; core.tlen = tlen;
    size_t l_seq = seq.size();
    int32_t block_size = (int32_t)(32 + core.l_read_name + 4 * cigar.size() + (l_seq + 1) / 2 + l_seq);
    rec_.resize(4 + (size_t)block_size);
    uint8_t* p = rec_.data();
    memcpy(p, &block_size, 4); p += 4;
    memcpy(p, &core, 32); p += 32;
    memcpy(p, qname.c_str(), core.l_read_name); p += core.l_read_name;
    if (!cigar.empty()) memcpy(p, cigar.data(), 4 * cigar.size());
    p += 4 * cigar.size();
    memset(p, 0, (l_seq + 1) / 2);
    for (size_t i = 0; i < l_seq; ++i) p[i >> 1] |= (uint8_t)(base_code(seq[i]) << ((~i & 1) << 2));
    p += (l_seq + 1) / 2;
    for (size_t i = 0; i < l_seq; ++i) p[i] = i < qual.size() ? qual[i] : 0xff;
    bg_.write(rec_.data(), rec_.size());
}

bool BamWriter::close() { return bg_.close(); }

}  // namespace msnv
