// snp_output.cc -- see snp_output.hpp.
#include "snp_output.hpp"

#include <algorithm>
#include <cstdlib>
#include <memory>
#include <cstring>

namespace msnv {

namespace {

// call_vC.cpp:92-111: skip leading blanks, copy up to the separator, step over it.
const char* next_token(const char* src, std::string& tok)
{
    tok.clear();
    while (*src == ' ') ++src;
    while (*src && *src != '\t') tok.push_back(*src++);
    if (*src == '\t') ++src;
    return src;
}

// Standard genetic code keyed by codon, as listed in gene.h:3-25 ('X' = stop).
char amino_acid(const std::string& codon)
{
    if (codon.size() != 3) return '\0';
    int v = 0;
    for (char c : codon) {
        int b = c == 'T' ? 0 : c == 'C' ? 1 : c == 'A' ? 2 : c == 'G' ? 3 : -1;
        if (b < 0) return '\0';                       // codons with N are not in the reference's table
        v = v * 4 + b;
    }
    // order T,C,A,G on each of the three positions
    static const char table[65] = "FFLLSSSSYYXXCCXWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    return table[v];
}

// call_vC.cpp:299-314 -- letters other than A/C/G/T vanish
std::string rev_complement(const std::string& s)
{
    std::string r;
    for (size_t i = s.size(); i-- > 0;) {
        char c = s[i];
        if (c == 'A') r.push_back('T'); else if (c == 'T') r.push_back('A');
        else if (c == 'C') r.push_back('G'); else if (c == 'G') r.push_back('C');
    }
    return r;
}

inline void put_uint(std::string& b, uint32_t v)
{
    char tmp[12]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) b.push_back(tmp[--n]);
}

}  // namespace

bool Annotation::load(const std::string& genes_path, const std::string& fasta_path, std::string& err)
{
    FILE* g = fopen(genes_path.c_str(), "r");
    if (!g) { err = "Cannot open " + genes_path; return false; }
    std::vector<char> line(10000);
    std::string tok, cur_name;
    bool have = false;
    Block cur;
    auto flush = [&]() { if (have) blocks_[cur_name] = cur; cur = Block(); };
    if (fgets(line.data(), 10000, g)) {               // header line
        while (fgets(line.data(), 10000, g)) {
            // columns (0-based): 1 gene name, 2 sequence id, 6 start, 7 end (1-based), 8 strand
            std::vector<std::string> col;
            const char* rest = line.data();
            size_t ll = strlen(rest);
            std::string ln(rest, ll);
            const char* p = ln.c_str();
            for (int k = 0; k < 9; ++k) { p = next_token(p, tok); col.push_back(tok); }
            bool more_after_id = false;                // the reference looks at the id only when text follows it
            {
                const char* q = ln.c_str(); std::string t;
                for (int k = 0; k < 3; ++k) q = next_token(q, t);
                more_after_id = *q != 0;
            }
            if (more_after_id) {
                if (!have) { cur_name = col[2]; have = true; }
                else if (cur_name != col[2]) { flush(); cur_name = col[2]; }
            }
            GeneRec r;
            r.name = col[1];
            r.start = atol(col[6].c_str()) - 1;
            r.end = atol(col[7].c_str()) - 1;
            std::string st = col[8];
            r.strand = st.empty() ? '\0' : st[0];
            cur.genes.push_back(r);
        }
    }
    have = true;                                       // the reference stores the trailing block unconditionally
    flush();
    fclose(g);

    FILE* f = fopen(fasta_path.c_str(), "r");
    if (!f) { err = "Cannot open " + fasta_path; return false; }
    std::string name, genome;
    bool skip = false;
    auto norm = [](char c) { return (c == 'A' || c == 'T' || c == 'C' || c == 'G' || c == 'N') ? c : 'A'; };
    while (fgets(line.data(), 10000, f)) {
        size_t l = strlen(line.data());
        if (l) line[l - 1] = '\0';                     // the reference drops the last character, whatever it is
        if (line[0] == '>') {
            if (!genome.empty() && !skip) { genomes_[name] = genome; genome.clear(); }
            name = line.data() + 1;
            skip = blocks_.find(name) == blocks_.end();
        } else if (!skip) {
            for (const char* c = line.data(); *c; ++c) genome.push_back(norm(*c));
        }
    }
    genomes_[name] = genome;
    fclose(f);
    active_ = true;
    return true;
}

bool Annotation::select(const std::string& contig)
{
    cur_ = nullptr; by_start_.clear(); heap_.clear(); next_ = 0; last_pos_ = -1;
    auto it = blocks_.find(contig);
    if (it == blocks_.end()) return false;
    cur_ = &it->second;
    for (uint32_t i = 0; i < cur_->genes.size(); ++i)
        if (cur_->genes[i].start <= cur_->genes[i].end) by_start_.push_back(i);      // "goes around" genes are ignored
    std::stable_sort(by_start_.begin(), by_start_.end(),
                     [&](uint32_t a, uint32_t b) { return cur_->genes[a].start < cur_->genes[b].start; });
    return true;
}

const GeneRec* Annotation::gene_at(long pos0)
{
    if (!cur_) return nullptr;
    if (pos0 < last_pos_) { heap_.clear(); next_ = 0; }           // out-of-order query (hand-written text input): restart the sweep
    last_pos_ = pos0;
    auto cmp = [](uint32_t a, uint32_t b) { return a > b; };       // min-heap on file order
    while (next_ < by_start_.size() && cur_->genes[by_start_[next_]].start <= pos0) {
        heap_.push_back(by_start_[next_++]);
        std::push_heap(heap_.begin(), heap_.end(), cmp);
    }
    // drop genes that ended before pos0; a gene deeper in the heap that has ended is dropped when it
    // surfaces, which is enough because only the top is ever reported
    while (!heap_.empty() && cur_->genes[heap_.front()].end < pos0) {
        std::pop_heap(heap_.begin(), heap_.end(), cmp);
        heap_.pop_back();
    }
    return heap_.empty() ? nullptr : &cur_->genes[heap_.front()];
}

std::string Annotation::codon(const std::string& contig, long start, long end) const
{
    auto it = genomes_.find(contig);
    if (it == genomes_.end()) return "";
    const std::string& g = it->second;
    const long len = (long)g.size();
    if (end < start || end > len) return "";          // gene.h:78-82 (note: end == length is let through)
    std::string r;
    for (long i = start; i <= end; ++i) r.push_back(i < len ? g[(size_t)i] : 'A');
    return r;
}

HitWriter::Locator HitWriter::shard_locator(const std::vector<Contig>& contigs, const uint8_t* shard_ref)
{
    // hits arrive in ascending order, so the contig cursor only moves forward
    auto ci = std::make_shared<size_t>(0);
    return [&contigs, shard_ref, ci](uint32_t p, const std::string*& name, long& pos0, char& refc) {
        size_t& c = *ci;
        if (p < contigs[c].offset) c = 0;
        while (c + 1 < contigs.size() && p >= contigs[c + 1].offset) ++c;
        name = &contigs[c].name; pos0 = (long)(p - contigs[c].offset); refc = (char)shard_ref[p];
    };
}

void HitWriter::write(const msnv_hits& hits, const Locator& locate)
{
    const uint32_t S = hits.n_samples;
    static const int order[4] = {0, 1, 3, 2};          // the reference walks "actg" (call_vC.cpp:561)
    static const char letter[4] = {'A', 'C', 'G', 'T'};
    for (uint32_t h = 0; h < hits.n_hits; ++h) {
        const uint32_t p = hits.pos[h];
        const std::string* cname = nullptr; long pos0 = 0; char refc = 'N';
        locate(p, cname, pos0, refc);
        struct { const std::string& name; } C{*cname};
        // gene lookup state, as the reference keeps it across lines
        const GeneRec* gene = nullptr;
        if (ann && ann->active()) {
            if (loaded_contig != C.name) { has_genes = ann->select(C.name); loaded_contig = C.name; }
            gene = ann->gene_at(pos0);
        }
        const uint16_t* cov = hits.cov + (size_t)h * S;
        std::string* dst[2] = {&buf_pop_, &buf_ind_};
        const uint8_t mask[2] = {hits.pop_mask[h], hits.ind_mask[h]};
        for (int kind = 0; kind < 2; ++kind) {
            if (!mask[kind]) continue;
            std::string& b = *dst[kind];
            const size_t line_start = b.size();
            b += C.name; b.push_back('\t');
            b += gene ? gene->name : std::string("-"); b.push_back('\t');
            put_uint(b, (uint32_t)(pos0 + 1)); b.push_back('\t');
            b.push_back(refc); b.push_back('\t');
            for (uint32_t s = 0; s < S; ++s) { if (s) b.push_back('|'); put_uint(b, cov[s]); }
            b.push_back('\t');
            bool first = true; bool any = false;
            for (int oi = 0; oi < 4; ++oi) {
                const int a = order[oi];
                if (!(mask[kind] >> a & 1)) continue;
                std::string ann_txt = ".";
                if (has_genes && gene) {
                    if (gene->start < gene->end) {
                        const long cp = (pos0 - gene->start) % 3, cs = pos0 - cp;
                        std::string oldc = ann->codon(C.name, cs, cs + 2);
                        std::string newc = oldc;
                        if ((size_t)cp < newc.size()) newc[(size_t)cp] = letter[a];
                        if (gene->strand == '-') { oldc = rev_complement(oldc); newc = rev_complement(newc); }
                        ann_txt = (amino_acid(newc) == amino_acid(oldc)) ? "S" : "N";
                        ann_txt += "[" + oldc + "-" + newc + "]";
                    } else {
                        continue;                      // single-base gene: the reference drops the allele (call_vC.cpp:614-617)
                    }
                }
                if (!first) b.push_back(',');
                first = false; any = true;
                put_uint(b, hits.total[(size_t)h * 5 + 1 + a]); b.push_back('|');
                b.push_back(letter[a]); b.push_back('|');
                b += ann_txt;
                const uint16_t* al = hits.allele + ((size_t)h * 4 + a) * S;
                for (uint32_t s = 0; s < S; ++s) { b.push_back('|'); put_uint(b, al[s]); }
            }
            b.push_back('\n');
            if (kind == 1 && !any) { b.resize(line_start); continue; }      // no individual text -> no line (call_vC.cpp:653)
            if (kind == 0) ++pop_lines; else ++indiv_lines;
        }
        if (buf_pop_.size() > (1u << 20)) { fwrite(buf_pop_.data(), 1, buf_pop_.size(), pop_out); buf_pop_.clear(); }
        if (buf_ind_.size() > (1u << 20)) {
            if (indiv_out) fwrite(buf_ind_.data(), 1, buf_ind_.size(), indiv_out);
            buf_ind_.clear();
        }
    }
    if (!buf_pop_.empty()) { fwrite(buf_pop_.data(), 1, buf_pop_.size(), pop_out); buf_pop_.clear(); }
    if (!buf_ind_.empty()) {
        if (indiv_out) fwrite(buf_ind_.data(), 1, buf_ind_.size(), indiv_out);
        else if (!warned_no_indiv) {
            fprintf(stderr, "Individual SNPs detected, but no individual output file specified (-i option).\n");
            warned_no_indiv = true;
        }
        buf_ind_.clear();
    }
}

}  // namespace msnv
