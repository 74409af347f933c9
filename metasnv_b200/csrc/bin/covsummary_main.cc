// metaSNV_covSummary -- SURVEY.md 8(f) rank 2: the step behind the coverage pass, host only.
//
//   metaSNV_covSummary <project_dir>
//
// does what metaSNV.py:compute_summary (metaSNV.py:96-121) starts S + 1 Python processes for:
//   * per `cov/<bam>.cov` + `cov/<bam>.cov.detail` (qaCompute's two files): `cov/<bam>.cov.summary`, one line per taxon
//     (contig name up to the first '.'): length-weighted average coverage, per cent of the bases covered at >= 1x and >= 2x
//     (src/computeGenomeCoverage.py:7-52);
//   * over all `cov/*.summary`: `<project>.all_cov.tab` and `<project>.all_perc.tab`, taxa x samples
//     (src/collapse_coverages.py:9-39).
// The text is byte-identical to the scripts' (tests/test_covsummary_cpu.py): same IEEE double arithmetic in the same order,
// "%f", taxa in order of first appearance in the summary and in byte order in the matrices, samples in the order of the sorted
// paths. Inputs the scripts die on (a .cov shorter than its .detail, a row with too few columns, a taxon missing from a sample)
// end this program with exit code 1 and a message instead of a Python traceback.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

bool read_line(FILE* f, std::string& out)             // like file.readline(): keeps the '\n'; false at EOF
{
    out.clear();
    int c;
    while ((c = fgetc(f)) != EOF) { out.push_back((char)c); if (c == '\n') break; }
    return !out.empty();
}

std::vector<std::string> split_tab(const std::string& s)       // str.split('\t')
{
    std::vector<std::string> v;
    size_t p = 0;
    for (;;) { const size_t q = s.find('\t', p); v.push_back(s.substr(p, q == std::string::npos ? q : q - p)); if (q == std::string::npos) break; p = q + 1; }
    return v;
}

std::vector<std::string> split_ws(const std::string& s)        // str.split()
{
    std::vector<std::string> v;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && isspace((unsigned char)s[i])) ++i;
        size_t j = i;
        while (j < s.size() && !isspace((unsigned char)s[j])) ++j;
        if (j > i) v.push_back(s.substr(i, j - i));
        i = j;
    }
    return v;
}

// int(x) / float(x) of a token that may carry white space around it
bool to_int(const std::string& t, long long& v)
{
    const char* b = t.c_str(); char* e = nullptr;
    while (isspace((unsigned char)*b)) ++b;
    if (!*b) return false;
    v = strtoll(b, &e, 10);
    if (e == b) return false;
    while (isspace((unsigned char)*e)) ++e;
    return *e == 0;
}
bool to_double(const std::string& t, double& v)
{
    const char* b = t.c_str(); char* e = nullptr;
    while (isspace((unsigned char)*b)) ++b;
    if (!*b) return false;
    v = strtod(b, &e);
    if (e == b) return false;
    while (isspace((unsigned char)*e)) ++e;
    return *e == 0;
}

struct Taxon { std::string id; double len = 0, cov_len = 0, x1 = 0, x2 = 0; };

// computeGenomeCoverage.py: <cov> <cov.detail> -> <cov.summary>
bool summarize(const std::string& cov_path, std::string& err)
{
    FILE* cov = fopen(cov_path.c_str(), "r");
    FILE* det = fopen((cov_path + ".detail").c_str(), "r");
    if (!cov || !det) { err = "cannot open " + cov_path + (cov ? ".detail" : ""); if (cov) fclose(cov); if (det) fclose(det); return false; }
    std::vector<Taxon> taxa;                                  // in order of first appearance (dict order)
    std::unordered_map<std::string, size_t> where;
    std::string cl, xl;
    read_line(cov, cl);                                       // the .cov header
    bool ok = true;
    for (;;) {
        const bool have_cov = read_line(cov, cl);
        if (!read_line(det, xl)) break;
        const std::vector<std::string> c = split_tab(have_cov ? cl : std::string()), x = split_tab(xl);
        if (c[0] != x[0]) printf("Mismatch in names %s != %s\n", c[0].c_str(), x[0].c_str());
        long long len, b1, b2; double avg;
        if (c.size() < 3 || x.size() < 4 || !to_int(c[1], len) || !to_double(c[2], avg) || !to_int(x[2], b1) || !to_int(x[3], b2)) {
            err = cov_path + ": a row of the coverage files has too few or unreadable columns";
            ok = false;
            break;
        }
        const std::string id = c[0].substr(0, c[0].find('.'));
        auto it = where.find(id);
        if (it == where.end()) { it = where.emplace(id, taxa.size()).first; taxa.push_back(Taxon{id}); }
        Taxon& t = taxa[it->second];
        t.len += (double)len;
        t.cov_len += avg * (double)len;
        t.x1 += (double)b1;
        t.x2 += (double)b2;
    }
    fclose(cov); fclose(det);
    if (!ok) return false;
    FILE* out = fopen((cov_path + ".summary").c_str(), "w");
    if (!out) { err = "cannot write " + cov_path + ".summary"; return false; }
    fputs("TaxId\tAverage_cov\tPercentage_1x\tPercentage_2x\n", out);
    for (const Taxon& t : taxa) fprintf(out, "%s\t%f\t%f\t%f\n", t.id.c_str(), t.cov_len / t.len, t.x1 / t.len * 100, t.x2 / t.len * 100);
    fclose(out);
    return true;
}

std::vector<std::string> list_dir(const std::string& dir, const std::string& suffix)      // sorted(glob(dir + '/*' + suffix))
{
    std::vector<std::string> v;
    if (DIR* d = opendir(dir.c_str())) {
        while (dirent* e = readdir(d)) {
            const std::string n = e->d_name;
            if (n.empty() || n[0] == '.') continue;                                        // glob's '*' skips dot files
            if (n.size() >= suffix.size() && n.compare(n.size() - suffix.size(), suffix.size(), suffix) == 0) v.push_back(dir + "/" + n);
        }
        closedir(d);
    }
    std::sort(v.begin(), v.end());
    return v;
}

std::string base_name(std::string p)                  // os.path.basename
{
    const size_t s = p.rfind('/');
    return s == std::string::npos ? p : p.substr(s + 1);
}

}  // namespace

int main(int argc, char** argv)
{
    if (argc != 2) { fprintf(stderr, "usage: metaSNV_covSummary <project_dir>\n"); return 1; }
    std::string project = argv[1];
    const std::string cov_dir = project + "/cov";
    const std::vector<std::string> covs = list_dir(cov_dir, ".cov");
    if (covs.empty()) { fprintf(stderr, "Coverage files not found.\n"); return 1; }
    std::string err;
    for (const std::string& c : covs)
        if (!summarize(c, err)) { fprintf(stderr, "metaSNV_covSummary: %s\n", err.c_str()); return 1; }

    // ---- collapse_coverages.py
    std::string name = project;                           // os.path.basename(project_dir): empty when the path ends with '/'
    name = base_name(name);
    std::vector<std::string> bamfiles;
    std::map<std::string, std::unordered_map<std::string, std::string>> avg, per;        // taxon -> sample -> text (taxa in byte order)
    for (const std::string& f : list_dir(cov_dir, ".summary")) {
        std::string b = base_name(f);
        b = b.size() >= strlen(".cov.summary") ? b.substr(0, b.size() - strlen(".cov.summary")) : std::string();
        FILE* in = fopen(f.c_str(), "r");
        if (!in) { fprintf(stderr, "metaSNV_covSummary: cannot open %s\n", f.c_str()); return 1; }
        std::string line;
        for (int i = 0; read_line(in, line); ++i) {
            if (i == 0) continue;
            const std::vector<std::string> t = split_ws(line);
            if (t.size() < 3) { fprintf(stderr, "metaSNV_covSummary: %s: a summary line with fewer than three columns\n", f.c_str()); fclose(in); return 1; }
            avg[t[0]][b] = t[1];
            per[t[0]][b] = t[2];
        }
        fclose(in);
        bamfiles.push_back(b);
    }
    auto write_matrix = [&](const std::map<std::string, std::unordered_map<std::string, std::string>>& m, const char* header, const std::string& path) {
        FILE* out = fopen(path.c_str(), "w");
        if (!out) { fprintf(stderr, "metaSNV_covSummary: cannot write %s\n", path.c_str()); return false; }
        fputc('\t', out);
        for (size_t i = 0; i < bamfiles.size(); ++i) { if (i) fputc('\t', out); fputs(bamfiles[i].c_str(), out); }
        fputs("\nTaxId\t", out);
        for (size_t i = 0; i < bamfiles.size(); ++i) { if (i) fputc('\t', out); fputs(header, out); }
        fputc('\n', out);
        for (const auto& row : avg) {                                                     // (the script walks avg_cov's taxa for both matrices)
            const auto it = m.find(row.first);
            fprintf(out, "%s\t", row.first.c_str());
            for (size_t i = 0; i < bamfiles.size(); ++i) {
                const std::string* cell = nullptr;
                if (it != m.end()) { const auto c = it->second.find(bamfiles[i]); if (c != it->second.end()) cell = &c->second; }
                if (!cell) {
                    fprintf(stderr, "metaSNV_covSummary: taxon %s is missing from %s\n", row.first.c_str(), bamfiles[i].c_str());
                    fclose(out);
                    return false;
                }
                if (i) fputc('\t', out);
                fputs(cell->c_str(), out);
            }
            fputc('\n', out);
        }
        fclose(out);
        return true;
    };
    if (!write_matrix(avg, "Average_cov", project + "/" + name + ".all_cov.tab")) return 1;
    if (!write_matrix(per, "Percentage_1x", project + "/" + name + ".all_perc.tab")) return 1;
    return 0;
}
