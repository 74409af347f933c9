/* htslib/sam.h STAND-IN for building the UNMODIFIED reference src/qaTools/qaCompute.cpp
 * (it includes <htslib/sam.h> at qaCompute.cpp:26; htslib is not in this image and is
 * not vendored by the reference). TEST INFRASTRUCTURE ONLY. Exposes exactly the names
 * qaCompute.cpp uses, implemented on the oracle's own reader (oracle/obam.c).
 * Constant values are those of the SAMv1 specification (FLAG bits, CIGAR op codes). */
#ifndef ORACLE_HTSLIB_SAM_STANDIN_H
#define ORACLE_HTSLIB_SAM_STANDIN_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../obam.h"

#define BAM_FPAIRED        1
#define BAM_FPROPER_PAIR   2
#define BAM_FUNMAP         4
#define BAM_FMUNMAP        8
#define BAM_FREVERSE      16
#define BAM_FMREVERSE     32
#define BAM_FREAD1        64
#define BAM_FREAD2       128
#define BAM_FSECONDARY   256
#define BAM_FQCFAIL      512
#define BAM_FDUP        1024
#define BAM_FSUPPLEMENTARY 2048

#define BAM_CMATCH      0
#define BAM_CINS        1
#define BAM_CDEL        2
#define BAM_CREF_SKIP   3
#define BAM_CSOFT_CLIP  4
#define BAM_CHARD_CLIP  5
#define BAM_CPAD        6
#define BAM_CEQUAL      7
#define BAM_CDIFF       8
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK  0xf

typedef struct { obam_file *f; } htsFile;
typedef obam_hdr bam_hdr_t;

typedef struct {
    int32_t  tid, pos;
    uint16_t bin;
    uint8_t  qual;
    uint16_t l_qname;
    uint16_t flag;
    uint32_t n_cigar;
    int32_t  l_qseq, mtid, mpos, isize;
} bam1_core_t;

typedef struct {
    bam1_core_t core;
    uint8_t *data;
    obam_rec rec;
} bam1_t;

static inline htsFile *sam_open(const char *path, const char *mode)
{
    (void)mode;
    obam_file *f = obam_open(path);
    if (!f) return NULL;
    htsFile *h = (htsFile *)malloc(sizeof(htsFile));
    h->f = f;
    return h;
}
static inline int hts_close(htsFile *h) { if (h) { obam_close(h->f); free(h); } return 0; }
static inline bam_hdr_t *sam_hdr_read(htsFile *h) { return obam_hdr_read(h->f); }
static inline bam1_t *bam_init1(void) { return (bam1_t *)calloc(1, sizeof(bam1_t)); }
static inline void bam_destroy1(bam1_t *b) { if (b) { obam_rec_free(&b->rec); free(b); } }
static inline int sam_read1(htsFile *h, bam_hdr_t *hdr, bam1_t *b)
{
    (void)hdr;
    int r = obam_read1(h->f, &b->rec);
    if (r < 0) return r;
    b->core.tid = b->rec.tid;     b->core.pos = b->rec.pos;     b->core.bin = b->rec.bin;
    b->core.qual = b->rec.mapq;   b->core.l_qname = b->rec.l_qname;
    b->core.flag = b->rec.flag;   b->core.n_cigar = b->rec.n_cigar;
    b->core.l_qseq = b->rec.l_qseq; b->core.mtid = b->rec.mtid;
    b->core.mpos = b->rec.mpos;   b->core.isize = b->rec.isize;
    b->data = b->rec.data;
    return r;
}
#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))

#endif
