// snp_output.hpp -- called_SNPs / indiv_called line writer and the optional gene / codon annotation
// of snpCall (-g). Format and rules follow call_vC.cpp:561-667 (line layout :645-651,:659-665,
// allele entries :625-636, coverage string :316-325) and gene.h (genome packing :29-37,:67,
// codon table :3-25); SURVEY.md Annex B and D list the quirks that are reproduced on purpose.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../../include/msnv.h"

namespace msnv {

struct GeneRec { long start, end; std::string name; char strand; };   // 0-based closed interval

class Annotation {
public:
    // genes: annotation table (-g); fasta: reference (-f). Mirrors indexGenomeAndGenes
    // (call_vC.cpp:116-199): genes are grouped in contiguous blocks per sequence_id (a reappearing id
    // replaces the earlier block); only sequences that have genes are kept from the FASTA, keyed by
    // the whole header line.
    bool load(const std::string& genes_path, const std::string& fasta_path, std::string& err);
    bool active() const { return active_; }
    // Start serving queries for one contig (loadGenome, call_vC.cpp:205-284). Returns whether the
    // contig has a gene block.
    bool select(const std::string& contig);
    // First gene (file order) whose closed interval contains pos0; nullptr if none. Positions must
    // be queried in ascending order after select().
    const GeneRec* gene_at(long pos0);
    // Reference codon text as the reference's Genome::getSequence returns it ("" when out of range).
    std::string codon(const std::string& contig, long start, long end) const;
    bool has_sequence(const std::string& contig) const { return genomes_.count(contig) != 0; }
private:
    struct Block { std::vector<GeneRec> genes; };
    bool active_ = false;
    std::map<std::string, Block> blocks_;
    std::map<std::string, std::string> genomes_;      // letters normalised the way gene.h packs them
    // sweep state of the selected contig
    const Block* cur_ = nullptr;
    std::vector<uint32_t> by_start_;                   // gene indices ordered by start
    size_t next_ = 0;
    long last_pos_ = -1;
    std::vector<uint32_t> heap_;                       // active gene indices, min-heap on file order
};

// Writes the lines of one shard's hits. `contig_of` / names / lengths map shard coordinates back
// to contigs.
struct HitWriter {
    FILE* pop_out = nullptr;          // stdout of snpCall
    FILE* indiv_out = nullptr;        // -i file (may be null)
    bool warned_no_indiv = false;
    Annotation* ann = nullptr;
    // state that the reference keeps across lines (call_vC.cpp:460-465,554-559)
    std::string loaded_contig;
    bool has_genes = false;
    uint64_t pop_lines = 0, indiv_lines = 0;

    struct Contig { std::string name; uint32_t offset, len; };
    // Maps a hit's coordinate to (contig name, 0-based position, reference character as printed).
    typedef std::function<void(uint32_t p, const std::string*& name, long& pos0, char& refc)> Locator;
    void write(const msnv_hits& hits, const Locator& locate);
    // Locator for shard coordinates (direct-from-BAM mode).
    static Locator shard_locator(const std::vector<Contig>& contigs, const uint8_t* shard_ref);
private:
    std::string buf_pop_, buf_ind_;
};

}  // namespace msnv
