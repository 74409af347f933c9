// text_pileup.hpp -- classic snpCall input: `samtools mpileup` text on stdin.
// The host only tokenises and counts characters exactly as call_vC.cpp:466-541 does (tab split with
// the 10000-character token limit of toksplit, '^' / '+n' / '-n' skipping, the ten counted symbols);
// thresholds, compaction and per-hit gathering run on the GPU through msnv_call_counts().
#pragma once
#include <cstdio>
#include <string>

#include "../../../include/msnv.h"

namespace msnv {

// first_line: the line already consumed from `in` (with its newline, may be empty at EOF).
// Returns the process exit status.
int run_text_mode(const std::string& first_line, FILE* in, const msnv_call_params& prm, const std::string& fasta_opt,
                  const std::string& genes_opt, FILE* indiv, int device);

}  // namespace msnv
