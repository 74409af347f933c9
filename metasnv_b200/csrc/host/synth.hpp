// synth.hpp -- synthetic metagenome writer (reference FASTA, annotation, sorted BAMs, sample list)
// for the shapes in BASELINE.json:configs / SURVEY.md section 8(d). The reference ships no test
// data, so parity and benchmark inputs are generated here.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "synth_model.h"

namespace msnv {

struct SynthGenome {
    int taxid = 0;
    int n_sub = 1;                       // number of subspecies clusters
    std::vector<uint32_t> contig_lens;
};
struct SynthSpike { int tid; uint32_t start, len, depth; };

struct SynthConfig {
    synth::Model model;
    std::vector<SynthGenome> genomes;
    std::vector<SynthSpike> spikes;      // extra single-end reads piled on short regions (cap test)
    uint32_t junk_pct_x10 = 40;          // per mille of reads followed by a filtered record
    uint32_t unmapped_pct_x10 = 10;      // per mille of reads mirrored by an unmapped record at EOF
    bool annotation = false;
    std::string project = "synth";
};

// Presets: "c1".."c5" (BASELINE.json configs 0..4) scaled by `scale` in genome length and, where
// n_samples > 0, with the sample count overridden.
bool synth_preset(const std::string& name, double scale, int n_samples, uint64_t seed, SynthConfig& cfg,
                  std::string& err);

struct SynthContig { std::string name; uint32_t len; int genome; };
std::vector<SynthContig> synth_contigs(const SynthConfig& cfg);

struct SynthStats {
    uint64_t reads = 0, aligned_bases = 0, junk = 0, unmapped = 0;
};

// Writes <dir>/ref.fa, <dir>/annotation.txt (if cfg.annotation), <dir>/bam/s<k>.bam and
// <dir>/all_samples. threads <= 0 means all cores.
bool synth_write(const SynthConfig& cfg, const std::string& dir, int threads, SynthStats& stats, std::string& err);

// Converts a (restricted) SAM text file into BAM: @SQ lines and the 11 mandatory columns.
bool sam_to_bam(const std::string& sam_path, const std::string& bam_path, std::string& err);

}  // namespace msnv
