// samtools -- stand-in for the two samtools sub-commands metaSNV.py runs (metaSNV.py:83,160-165),
// to be placed first on PATH so that the unchanged orchestrator drives the GPU path:
//   samtools view -H <bam>                     prints the SAM header text of the BAM
//   samtools mpileup -f REF [-l BED] -B -b LIST  prints a one-line job descriptor that this
//                                              repository's snpCall turns into a GPU pileup
// Anything else is rejected: this is not a general samtools.
#include <cstdio>
#include <cstring>
#include <string>

#include "../host/bam.hpp"

int main(int argc, char** argv)
{
    if (argc >= 4 && !strcmp(argv[1], "view") && !strcmp(argv[2], "-H")) {
        msnv::BamReader r;
        if (!r.open(argv[3])) { fprintf(stderr, "samtools view: %s\n", r.error().c_str()); return 1; }
        const std::string& t = r.header().text;
        fwrite(t.data(), 1, t.size(), stdout);
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "mpileup")) {
        std::string ref, bed = "-", list;
        for (int i = 2; i < argc; ++i) {
            if (!strcmp(argv[i], "-f") && i + 1 < argc) ref = argv[++i];
            else if (!strcmp(argv[i], "-l") && i + 1 < argc) bed = argv[++i];
            else if (!strcmp(argv[i], "-b") && i + 1 < argc) list = argv[++i];
            else if (!strcmp(argv[i], "-B")) { }
            else { fprintf(stderr, "samtools mpileup (metasnv_b200 stand-in): unsupported argument %s\n", argv[i]); return 1; }
        }
        if (ref.empty() || list.empty()) { fprintf(stderr, "samtools mpileup (metasnv_b200 stand-in): -f REF and -b LIST are required\n"); return 1; }
        printf("#MSNV1\t%s\t%s\t%s\n", ref.c_str(), bed.c_str(), list.c_str());
        return 0;
    }
    if (argc >= 2 && (!strcmp(argv[1], "--version") || !strcmp(argv[1], "version"))) {
        printf("samtools (metasnv_b200 stand-in: view -H, mpileup -f/-l/-B/-b only)\n");
        return 0;
    }
    fprintf(stderr, "samtools (metasnv_b200 stand-in): only `view -H BAM` and `mpileup -f REF [-l BED] -B -b LIST` are supported\n");
    return 1;
}
