// msnv_synth -- command line front end of the synthetic data writer (host/synth.hpp).
//   msnv_synth --preset c1 [--scale F] [--samples N] [--seed S] [--threads T] [--depth X] --out DIR
//   msnv_synth --sam in.sam --bam out.bam
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../host/synth.hpp"

int main(int argc, char** argv)
{
    std::string preset, out, sam, bam;
    double scale = 1.0, depth = -1;
    int samples = 0, threads = 0;
    unsigned long long seed = 0;
    bool annotate = false, describe = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&](const char* nm) { if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", nm); exit(2); } return argv[++i]; };
        if (a == "--preset") preset = need("--preset");
        else if (a == "--scale") scale = atof(need("--scale"));
        else if (a == "--samples") samples = atoi(need("--samples"));
        else if (a == "--seed") seed = strtoull(need("--seed"), nullptr, 10);
        else if (a == "--threads") threads = atoi(need("--threads"));
        else if (a == "--depth") depth = atof(need("--depth"));
        else if (a == "--out") out = need("--out");
        else if (a == "--sam") sam = need("--sam");
        else if (a == "--bam") bam = need("--bam");
        else if (a == "--annotation") annotate = true;
        else if (a == "--describe") describe = true;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    std::string err;
    if (!sam.empty()) {
        if (bam.empty()) { fprintf(stderr, "--sam needs --bam\n"); return 2; }
        if (!msnv::sam_to_bam(sam, bam, err)) { fprintf(stderr, "msnv_synth: %s\n", err.c_str()); return 1; }
        return 0;
    }
    if (describe && !preset.empty()) {          // print the configuration as JSON (consumed by bench.py / tests for msnv_shard_synth)
        msnv::SynthConfig cfg;
        if (!msnv::synth_preset(preset, scale, samples, seed, cfg, err)) { fprintf(stderr, "msnv_synth: %s\n", err.c_str()); return 1; }
        if (depth > 0) cfg.model.depth_x100 = (uint32_t)(depth * 100);
        const msnv::synth::Model& m = cfg.model;
        printf("{\"seed\": %llu, \"n_samples\": %d, \"read_len\": %d, \"depth_x100\": %u, \"presence_ppm\": %u, \"paired_pct\": %u, "
               "\"site_ppm\": %u, \"err_ppm\": %u, \"nbase_ppm\": %u, \"refn_ppm\": %u, \"indel_pct_x10\": %u, \"clip_pct_x10\": %u, "
               "\"mapq0_pct_x10\": %u, \"spikes\": %zu, \"genome_n_sub\": [",
               (unsigned long long)m.seed, m.n_samples, m.read_len, m.depth_x100, m.presence_ppm, m.paired_pct, m.site_ppm, m.err_ppm,
               m.nbase_ppm, m.refn_ppm, m.indel_pct_x10, m.clip_pct_x10, m.mapq0_pct_x10, cfg.spikes.size());
        for (size_t g = 0; g < cfg.genomes.size(); ++g) printf("%s%d", g ? ", " : "", cfg.genomes[g].n_sub);
        printf("], \"contig_len\": [");
        auto ctgs = msnv::synth_contigs(cfg);
        for (size_t k = 0; k < ctgs.size(); ++k) printf("%s%u", k ? ", " : "", ctgs[k].len);
        printf("], \"contig_genome\": [");
        for (size_t k = 0; k < ctgs.size(); ++k) printf("%s%d", k ? ", " : "", ctgs[k].genome);
        printf("], \"contig_name\": [");
        for (size_t k = 0; k < ctgs.size(); ++k) printf("%s\"%s\"", k ? ", " : "", ctgs[k].name.c_str());
        printf("]}\n");
        return 0;
    }
    if (preset.empty() || out.empty()) {
        fprintf(stderr, "usage: msnv_synth --preset c1..c5 [--scale F] [--samples N] [--seed S] [--threads T] [--depth X] [--annotation] --out DIR\n"
                        "       msnv_synth --sam in.sam --bam out.bam\n");
        return 2;
    }
    msnv::SynthConfig cfg;
    if (!msnv::synth_preset(preset, scale, samples, seed, cfg, err)) { fprintf(stderr, "msnv_synth: %s\n", err.c_str()); return 1; }
    if (depth > 0) cfg.model.depth_x100 = (uint32_t)(depth * 100);
    if (annotate) cfg.annotation = true;
    msnv::SynthStats st;
    if (!msnv::synth_write(cfg, out, threads, st, err)) { fprintf(stderr, "msnv_synth: %s\n", err.c_str()); return 1; }
    printf("{\"preset\": \"%s\", \"samples\": %d, \"contigs\": %zu, \"reads\": %llu, \"aligned_bases\": %llu, \"junk\": %llu, \"unmapped\": %llu}\n",
           preset.c_str(), cfg.model.n_samples, msnv::synth_contigs(cfg).size(), (unsigned long long)st.reads,
           (unsigned long long)st.aligned_bases, (unsigned long long)st.junk, (unsigned long long)st.unmapped);
    return 0;
}
