// text_pileup.cc -- see text_pileup.hpp.
#include "text_pileup.hpp"

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "snp_output.hpp"

namespace msnv {

namespace {

const size_t kTokLimit = 10000;        // call_vC.cpp:482-483: toksplit(..., 10000)

struct LineMeta { uint32_t name_id; long pos0; char refc; };

// Count one bases column (call_vC.cpp:507-534) into letter counts a,c,g,t and matches.
inline void count_column(const char* tok, size_t len, uint32_t cnt[5])
{
    size_t i = 0;
    while (i < len) {
        switch (tok[i]) {
            case '^': ++i; break;
            case '+': case '-': {
                long skip = 0; bool any = false;
                while (i + 1 < len + 1 && isdigit((unsigned char)tok[i + 1])) { ++i; skip = skip * 10 + (tok[i] - '0'); any = true; if (skip > 100000000) skip = 100000000; }
                ++i;                     // first character after the digits
                (void)any;
                i += (size_t)(skip > 0 ? skip : 0);
                continue;                // the reference's `i += skip-1; ++i` lands on the same index
            }
            case '*': case '$': case 'N': case 'n': break;
            case '.': case ',': ++cnt[4]; break;
            case 'a': case 'A': ++cnt[0]; break;
            case 'c': case 'C': ++cnt[1]; break;
            case 'g': case 'G': ++cnt[2]; break;
            case 't': case 'T': ++cnt[3]; break;
            default: break;              // the reference crashes on other symbols (SURVEY.md Annex E #16)
        }
        ++i;
    }
}

}  // namespace

int run_text_mode(const std::string& first_line, FILE* in, const msnv_call_params& prm, const std::string& fasta_opt,
                  const std::string& genes_opt, FILE* indiv, int device)
{
    unsigned tabs = 0;
    for (char ch : first_line) if (ch == '\t') ++tabs;
    const int nr = first_line.empty() ? 0 : (int)(tabs + 1 - 3) / 3;         // call_vC.cpp:424-431
    fprintf(stderr, "Identified %d samples\n", nr);

    Annotation ann;
    std::string err;
    if (!fasta_opt.empty() && !genes_opt.empty()) {
        fprintf(stderr, "Found reference genomes and annotation file.\nLoading Genomes...\n");
        if (!ann.load(genes_opt, fasta_opt, err)) { fprintf(stderr, "%s\n", err.c_str()); return 255; }
        fprintf(stderr, "Genomes loaded!\n");
    }
    if (nr <= 0) {                       // nothing can be counted; drain the input like the reference would
        char buf[1 << 16];
        while (fread(buf, 1, sizeof buf, in) > 0) { }
        return 0;
    }
    const uint32_t S = (uint32_t)nr;

    msnv_ctx* ctx = nullptr;
    // lines per batch: a multiple of the tile, about 256 MB of count tiles
    uint64_t lines = ((uint64_t)256 << 20) / (10ull * S) / MSNV_TILE * MSNV_TILE;
    if (lines < MSNV_TILE) lines = MSNV_TILE;
    if (lines > 64ull * MSNV_TILE) lines = 64ull * MSNV_TILE;
    const uint32_t B = (uint32_t)lines, n_tiles = B / MSNV_TILE;
    std::vector<uint64_t> acgt((size_t)B * S);
    std::vector<uint16_t> match((size_t)B * S);
    std::vector<uint8_t> ref(B);
    std::vector<LineMeta> meta(B);
    std::vector<std::string> names;
    (void)n_tiles;

    HitWriter w;
    w.pop_out = stdout; w.indiv_out = indiv; w.ann = ann.active() ? &ann : nullptr;

    auto flush = [&](uint32_t n_lines) -> int {
        if (n_lines == 0) return 0;
        if (!ctx) {
            if (device < 0 || msnv_create(device, &ctx) != MSNV_OK) {
                fprintf(stderr, "snpCall: no usable CUDA device (%s); this build has no CPU calling path\n", msnv_last_error(ctx));
                return 1;
            }
        }
        const uint32_t P = (n_lines + MSNV_TILE - 1) / MSNV_TILE * MSNV_TILE;
        for (uint32_t i = n_lines; i < P; ++i) ref[i] = 0;
        msnv_hits hits;
        if (msnv_call_counts(ctx, S, P, ref.data(), acgt.data(), match.data(), &prm, &hits) != MSNV_OK) {
            fprintf(stderr, "snpCall: %s\n", msnv_last_error(ctx));
            return 1;
        }
        w.write(hits, [&](uint32_t p, const std::string*& name, long& pos0, char& refc) {
            name = &names[meta[p].name_id]; pos0 = meta[p].pos0; refc = meta[p].refc;
        });
        return 0;
    };

    char* line = nullptr; size_t cap = 0; ssize_t len;
    std::string tok;
    uint32_t n = 0;
    int rc = 0;
    while ((len = getline(&line, &cap, in)) > 0) {
        line[--len] = '\0';                                  // call_vC.cpp:475 drops the last character
        if (n == 0) {
            std::fill(acgt.begin(), acgt.end(), 0);
            std::fill(match.begin(), match.end(), 0);
        }
        const uint32_t tile = n / MSNV_TILE, off = n % MSNV_TILE;
        LineMeta lm{0, -1, '\0'};
        // toksplit loop (call_vC.cpp:483-541): a token is examined only while text follows it
        const char* p = line;
        int pos = 0;
        auto next = [&]() {
            tok.clear();
            while (*p == ' ') ++p;
            while (*p && *p != '\t') { if (tok.size() < kTokLimit) tok.push_back(*p); ++p; }
            if (*p == '\t') ++p;
        };
        next();
        std::string cname;
        while (*p) {
            if (pos == 0) cname = tok;
            else if (pos == 1) lm.pos0 = atol(tok.c_str()) - 1;
            else if (pos == 2) lm.refc = tok.empty() ? '\0' : tok[0];
            else if (pos > 3 && pos % 3 == 1) {
                const uint32_t s = (uint32_t)(pos / 3) - 1;
                if (s < S) {
                    uint32_t c[5] = {0, 0, 0, 0, 0};
                    count_column(tok.c_str(), tok.size(), c);
                    // 16-bit count lanes; a column is at most kTokLimit characters wide (call_vC.cpp:92-111 truncates
                    // there), so this only guards the packing below against a change of that limit
                    static_assert(kTokLimit <= 65535, "column counts are packed into 16-bit lanes");
                    const size_t o = ((size_t)tile * S + s) * MSNV_TILE + off;
                    acgt[o] = (uint64_t)c[0] | (uint64_t)c[1] << 16 | (uint64_t)c[2] << 32 | (uint64_t)c[3] << 48;
                    match[o] = (uint16_t)c[4];
                }
            }
            ++pos;
            next();
        }
        if (names.empty() || names.back() != cname) names.push_back(cname);
        lm.name_id = (uint32_t)names.size() - 1;
        meta[n] = lm;
        // a reference character of 0 would mean "do not call"; an empty column cannot be called anyway
        ref[n] = lm.refc ? (uint8_t)lm.refc : (uint8_t)' ';
        if (++n == B) {
            if ((rc = flush(n)) != 0) break;
            n = 0;
            std::string keep = names.back();
            names.assign(1, keep);
        }
    }
    if (rc == 0) rc = flush(n);
    free(line);
    fflush(stdout);
    if (ctx) msnv_destroy(ctx);
    return rc;
}

}  // namespace msnv
