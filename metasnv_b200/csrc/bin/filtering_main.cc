// filtering_main.cc -- `metaSNV_Filtering`: the consumer right behind the hot path.
//
// Drop-in for the reference's metaSNV_Filtering.py (same command line, same inputs in the project
// directory, same `filtered/pop/<taxon>.filtered.freq` and `filtered/ind/...` files): host-only C++,
// one thread per taxon; a line is only tokenised by the taxon it belongs to (the reference splits every
// line S-wide for every taxon).
//   filter I  (metaSNV_Filtering.py:108-148): samples of interest per taxon from <proj>.all_cov.tab /
//             <proj>.all_perc.tab: average depth >= -d and breadth >= -b, at least -m such samples
//   filter II (metaSNV_Filtering.py:156-242): per called position of the taxon, the samples of interest
//             with site coverage >= -c; keep the position when their share is >= -p; per alternative
//             allele one row  CHROM:GENE:POS:REF>ALT:ANNOT <tab> allele_count/site_coverage ... (-1
//             where the sample is not informative)
// Numbers are written the way Python's str(float) writes them (shortest round-trip digits, fixed
// notation for 1e-4 <= |x| < 1e16 with a trailing ".0" for integers, else d.ddde-XX), so the files are
// byte-identical to the reference's; tests/test_filtering_cpu.py holds the comparison (the tolerance
// north_star allows for allele frequencies, 1e-6 relative, is not needed).
#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Args {
    std::string projdir;
    double b = 40.0, d = 5.0, c = 5.0, p = 0.50;
    int m = 2, n_threads = 1;
    bool ind = false;
};

[[noreturn]] void die(const std::string& msg) { fprintf(stderr, "%s\n", msg.c_str()); exit(1); }

int usage()
{
    fprintf(stderr, "usage: metaSNV_filtering.py [-h] [-b FLOAT] [-d FLOAT] [-m INT] [-c FLOAT] [-p FLOAT] [--ind]\n"
                    "                            [--n_threads : Number of Processes] Proj\n");
    return 2;
}

std::vector<std::string> split_ws(const std::string& s)        // Python's str.split()
{
    std::vector<std::string> out;
    size_t i = 0, n = s.size();
    while (i < n) {
        while (i < n && isspace((unsigned char)s[i])) ++i;
        size_t j = i;
        while (j < n && !isspace((unsigned char)s[j])) ++j;
        if (j > i) out.emplace_back(s, i, j - i);
        i = j;
    }
    return out;
}

// Python 3 str(float): repr() = shortest digits that round-trip; exponent notation below 1e-4 and from 1e16
void append_py_float(std::string& out, double x)
{
    if (x != x) { out += "nan"; return; }
    if (x == 1.0 / 0.0) { out += "inf"; return; }
    if (x == -1.0 / 0.0) { out += "-inf"; return; }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest
    std::string s(buf, r.ptr);
    size_t k = 0;
    if (s[0] == '-') { out += '-'; k = 1; }
    const size_t e = s.find('e');
    std::string digits;
    for (size_t i = k; i < e; ++i) if (s[i] != '.') digits += s[i];
    const int exp10 = atoi(s.c_str() + e + 1);          // value = d.ddd * 10^exp10
    if (exp10 < -4 || exp10 >= 16) {
        out += digits[0];
        if (digits.size() > 1) { out += '.'; out.append(digits, 1, std::string::npos); }
        char eb[16];
        snprintf(eb, sizeof eb, "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
        out += eb;
    } else if (exp10 < 0) {
        out += "0.";
        out.append((size_t)(-exp10 - 1), '0');
        out += digits;
    } else {
        const size_t int_len = (size_t)exp10 + 1;
        if (digits.size() <= int_len) {
            out += digits;
            out.append(int_len - digits.size(), '0');
            out += ".0";
        } else {
            out.append(digits, 0, int_len);
            out += '.';
            out.append(digits, int_len, std::string::npos);
        }
    }
}

double py_float(const std::string& s, const std::string& what)
{
    char* end = nullptr;
    double v = strtod(s.c_str(), &end);
    if (end == s.c_str() || *end) die("ERROR: could not convert string to float: '" + s + "' (" + what + ")");
    return v;
}

long py_int(const std::string& s, const std::string& what)
{
    char* end = nullptr;
    long v = strtol(s.c_str(), &end, 10);
    if (end == s.c_str() || *end) die("ERROR: invalid literal for int(): '" + s + "' (" + what + ")");
    return v;
}

bool is_file(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
bool is_dir(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }

void rm_tree(const std::string& p)
{
    DIR* d = opendir(p.c_str());
    if (d) {
        while (dirent* e = readdir(d)) {
            if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
            const std::string q = p + "/" + e->d_name;
            struct stat st;
            if (lstat(q.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) rm_tree(q); else unlink(q.c_str());
        }
        closedir(d);
    }
    rmdir(p.c_str());
}

// files of `dir` whose name starts with `prefix`, in directory order (what glob.glob returns)
std::vector<std::string> glob_prefix(const std::string& dir, const std::string& prefix)
{
    std::vector<std::string> out;
    DIR* d = opendir(dir.c_str());
    if (!d) return out;
    while (dirent* e = readdir(d))
        if (!strncmp(e->d_name, prefix.c_str(), prefix.size())) out.push_back(dir + "/" + e->d_name);
    closedir(d);
    return out;
}

struct Taxon {
    std::string id;
    std::vector<std::string> samples;       // samples of interest, in header order
    std::vector<size_t> index;              // their columns in the called_SNPs files
};

// filter I
std::vector<Taxon> relevant_taxa(const Args& a, const std::string& cov_file, const std::string& perc_file)
{
    std::ifstream cov(cov_file), per(perc_file);
    std::string lc, lp;
    std::getline(cov, lc); std::getline(per, lp);
    const std::vector<std::string> header_cov = split_ws(lc), header_per = split_ws(lp);
    std::getline(cov, lc); std::getline(per, lp);               // second row: column descriptions
    if (header_cov != header_per) die("ERROR: Coverage file headers do not match!");
    std::vector<Taxon> out;
    while (std::getline(cov, lc) && std::getline(per, lp)) {
        std::vector<std::string> c = split_ws(lc), p = split_ws(lp);
        if (c.empty() || p.empty()) die("ERROR: empty line in the coverage tables");
        if (c[0] != p[0]) die("ERROR: TaxIDs in the coverage files are not in the same order!");
        Taxon t; t.id = c[0];
        const size_t n = std::min(c.size(), p.size()) - 1;
        bool complete = false;
        for (size_t i = 0; i < n; ++i) {
            if (i >= header_cov.size()) die("ERROR: more columns than samples in the coverage tables");
            if (py_float(c[i + 1], cov_file) >= a.d && py_float(p[i + 1], perc_file) >= a.b) t.samples.push_back(header_cov[i]);
            if (i + 1 == header_cov.size()) complete = true;
        }
        if (complete && (long)t.samples.size() >= a.m) {
            auto it = std::find_if(out.begin(), out.end(), [&](const Taxon& o) { return o.id == t.id; });
            if (it != out.end()) *it = t; else out.push_back(t);       // a dict: the last row of a taxon wins
        }
    }
    return out;
}

// fields of s[b, e) separated by c, as (begin, end) offsets: Python's str.split(c) without the copies
void split_span(const std::string& s, size_t b, size_t e, char c, std::vector<std::pair<size_t, size_t>>& out)
{
    out.clear();
    size_t i = b;
    for (;;) {
        const void* hit = i < e ? memchr(s.data() + i, c, e - i) : nullptr;
        if (!hit) { out.emplace_back(i, e); break; }
        const size_t j = (size_t)((const char*)hit - s.data());
        out.emplace_back(i, j);
        i = j + 1;
    }
}

long span_int(const std::string& s, std::pair<size_t, size_t> f, const std::string& what)
{
    long v = 0;
    auto r = std::from_chars(s.data() + f.first, s.data() + f.second, v);
    if (r.ec != std::errc() || r.ptr != s.data() + f.second) return py_int(s.substr(f.first, f.second - f.first), what);   // "+3", " 3": Python's rules
    return v;
}

double span_float(const std::string& s, std::pair<size_t, size_t> f, const std::string& what)
{
    double v = 0;
    auto r = std::from_chars(s.data() + f.first, s.data() + f.second, v);
    if (r.ec != std::errc() || r.ptr != s.data() + f.second) return py_float(s.substr(f.first, f.second - f.first), what);
    return v;
}

// filter II for one taxon
void filter_two(const Taxon& t, const Args& a, const std::vector<std::string>& snp_files, const std::string& outdir)
{
    FILE* out = nullptr;
    const std::string out_path = outdir + "/" + t.id + ".filtered.freq";
    std::string row, line;
    std::vector<long> site_cov;
    std::vector<std::pair<size_t, size_t>> tok, cov_f, alleles, xs;
    for (const std::string& f : snp_files) {
        std::ifstream in(f);
        while (std::getline(in, line)) {
            // taxon = contig name up to its first '.'
            size_t b = 0;
            while (b < line.size() && isspace((unsigned char)line[b])) ++b;
            if (b >= line.size()) die("ERROR: empty line in " + f);
            size_t e = b;
            while (e < line.size() && !isspace((unsigned char)line[e]) && line[e] != '.') ++e;
            if (e - b != t.id.size() || line.compare(b, e - b, t.id) != 0) continue;
            // whitespace-separated columns (str.split())
            tok.clear();
            for (size_t i = b, n = line.size(); i < n;) {
                while (i < n && isspace((unsigned char)line[i])) ++i;
                size_t j = i;
                while (j < n && !isspace((unsigned char)line[j])) ++j;
                if (j > i) tok.emplace_back(i, j);
                i = j;
            }
            if (tok.size() < 6) die("ERROR: SNP FILE " + f + " is corrupted");
            split_span(line, tok[4].first, tok[4].second, '|', cov_f);
            site_cov.resize(cov_f.size());
            for (size_t i = 0; i < cov_f.size(); ++i) site_cov[i] = span_int(line, cov_f[i], f);
            size_t nr_good = 0;
            for (size_t idx : t.index) {
                if (idx >= site_cov.size()) die("ERROR: SNP FILE " + f + " is corrupted");
                if (!((double)site_cov[idx] < a.c || site_cov[idx] == 0)) ++nr_good;
            }
            if ((double)nr_good / (double)t.index.size() < a.p) continue;
            if (!out) {
                if (is_file(out_path)) die("ERROR: " + out_path + " exists already");
                out = fopen(out_path.c_str(), "w");
                if (!out) die("ERROR: cannot write " + out_path);
                printf("Generating: %s\n", out_path.c_str());
                row = "\t";
                for (size_t i = 0; i < t.samples.size(); ++i) { if (i) row += '\t'; row += t.samples[i]; }
                row += '\n';
                fwrite(row.data(), 1, row.size(), out);
            }
            split_span(line, tok[5].first, tok[5].second, ',', alleles);
            for (const auto& al : alleles) {
                split_span(line, al.first, al.second, '|', xs);
                if (xs.size() < 3) die("ERROR: SNP FILE " + f + " is corrupted");
                if (xs.size() - 3 != site_cov.size()) {
                    printf("ERROR: SNP FILE %s is corrupted\n", f.c_str());
                    die("ERROR: Site coverage and SNP coverage string have uneven length!");
                }
                row.clear();
                for (int k = 0; k < 4; ++k) { if (k) row += ':'; row.append(line, tok[k].first, tok[k].second - tok[k].first); }
                row += '>'; row.append(line, xs[1].first, xs[1].second - xs[1].first);
                row += ':'; row.append(line, xs[2].first, xs[2].second - xs[2].first);
                for (size_t idx : t.index) {
                    row += '\t';
                    if ((double)site_cov[idx] >= a.c && site_cov[idx] != 0) append_py_float(row, span_float(line, xs[3 + idx], f) / (double)site_cov[idx]);
                    else row += "-1";
                }
                row += '\n';
                fwrite(row.data(), 1, row.size(), out);
            }
        }
    }
    if (out) { printf("closing: %s\n", t.id.c_str()); fclose(out); }
}

void run_pass(const std::vector<Taxon>& taxa, const Args& a, const std::vector<std::string>& files, const std::string& outdir)
{
    std::atomic<size_t> next(0);
    std::vector<std::thread> pool;
    const int n = std::max(1, std::min<int>(a.n_threads, (int)taxa.size()));
    for (int i = 0; i < n; ++i)
        pool.emplace_back([&]() { for (size_t k; (k = next.fetch_add(1)) < taxa.size();) filter_two(taxa[k], a, files, outdir); });
    for (auto& th : pool) th.join();
}

}  // namespace

int main(int argc, char** argv)
{
    Args a;
    for (int i = 1; i < argc; ++i) {
        const std::string o = argv[i];
        auto val = [&](const char* name) -> std::string {
            if (i + 1 >= argc) { fprintf(stderr, "metaSNV_filtering.py: error: argument %s: expected one argument\n", name); exit(2); }
            return argv[++i];
        };
        if (o == "-h" || o == "--help") { usage(); return 0; }
        else if (o == "--version") { printf("metaSNV_filtering.py 2.0\n"); return 0; }
        else if (o == "--debug") {}
        else if (o == "-b") a.b = py_float(val("-b"), "-b");
        else if (o == "-d") a.d = py_float(val("-d"), "-d");
        else if (o == "-c") a.c = py_float(val("-c"), "-c");
        else if (o == "-p") a.p = py_float(val("-p"), "-p");
        else if (o == "-m") a.m = (int)py_int(val("-m"), "-m");
        else if (o == "--n_threads") a.n_threads = (int)py_int(val("--n_threads"), "--n_threads");
        else if (o == "--ind") a.ind = true;
        else if (!o.empty() && o[0] == '-' && o.size() > 1) { fprintf(stderr, "metaSNV_filtering.py: error: unrecognized arguments: %s\n", o.c_str()); return usage(); }
        else if (a.projdir.empty()) a.projdir = o;
        else { fprintf(stderr, "metaSNV_filtering.py: error: unrecognized arguments: %s\n", o.c_str()); return usage(); }
    }
    if (a.projdir.empty()) { fprintf(stderr, "metaSNV_filtering.py: error: the following arguments are required: Proj\n"); return usage(); }
    while (a.projdir.size() > 1 && a.projdir.back() == '/') a.projdir.pop_back();
    const std::string name = a.projdir.substr(a.projdir.rfind('/') == std::string::npos ? 0 : a.projdir.rfind('/') + 1);
    const std::string cov_file = a.projdir + "/" + name + ".all_cov.tab", perc_file = a.projdir + "/" + name + ".all_perc.tab";
    const std::string all_samples = a.projdir + "/all_samples";

    // file_check (metaSNV_Filtering.py:57-76)
    printf("Checking for necessary input files...\n");
    if (is_file(cov_file) && is_file(perc_file)) printf("found: '%s' \nfound:'%s'\n", cov_file.c_str(), perc_file.c_str());
    else die("\nERROR: No such file '" + cov_file + "',\nERROR: No such file '" + perc_file + "'");
    if (is_file(all_samples)) printf("found: '%s'\n\n", all_samples.c_str());
    else die("\nERROR: No such file '" + all_samples + "'");

    std::vector<Taxon> taxa = relevant_taxa(a, cov_file, perc_file);

    std::vector<std::string> snp_header;
    {
        std::ifstream in(all_samples);
        std::string l;
        while (std::getline(in, l)) {
            if (!l.empty() && l.back() == '\r') l.pop_back();
            const size_t s = l.rfind('/');
            snp_header.push_back(s == std::string::npos ? l : l.substr(s + 1));
        }
    }
    for (Taxon& t : taxa)
        for (const std::string& n : t.samples) {
            auto it = std::find(snp_header.begin(), snp_header.end(), n);
            if (it == snp_header.end()) die("ERROR: '" + n + "' is not in list (all_samples)");
            t.index.push_back((size_t)(it - snp_header.begin()));
        }

    const std::string filt = a.projdir + "/filtered/";
    if (is_dir(a.projdir + "/filtered")) rm_tree(a.projdir + "/filtered");
    if (mkdir((a.projdir + "/filtered").c_str(), 0777) != 0 || mkdir((filt + "/pop").c_str(), 0777) != 0) die("ERROR: cannot create " + filt);
    run_pass(taxa, a, glob_prefix(a.projdir + "/snpCaller", "called"), filt + "/pop");
    if (a.ind) {
        if (!is_dir(filt + "/ind")) mkdir((filt + "/ind").c_str(), 0777);
        run_pass(taxa, a, glob_prefix(a.projdir + "/snpCaller", "indiv"), filt + "/ind");
    }
    return 0;
}
