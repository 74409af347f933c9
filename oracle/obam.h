/* obam.h -- tiny single-threaded BGZF/BAM reader for the ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product path;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build or run it. It is deliberately independent of the product's own
 * BGZF/BAM code (metasnv_b200/csrc/host) so that a decoder bug cannot hide in both.
 *
 * htslib is not vendored in /root/reference (qaCompute.cpp:26-27 includes it from the
 * system) and is absent from this image, so this follows the published SAMv1
 * specification (section 4, "The BAM format", and 4.1 "The BGZF compression format").
 */
#ifndef OBAM_H
#define OBAM_H
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct obam_file obam_file;

typedef struct {
    int32_t   n_targets;
    char    **target_name;
    uint32_t *target_len;
    char     *text;       /* SAM header text (what `samtools view -H` prints) */
    uint32_t  l_text;
} obam_hdr;

typedef struct {
    int32_t  tid, pos;
    uint16_t l_qname;    /* stored length incl. NUL padding to a multiple of 4 */
    uint8_t  mapq;
    uint16_t bin, n_cigar, flag;
    int32_t  l_qseq, mtid, mpos, isize;
    /* variable part, all pointing into data */
    uint8_t *data;        /* qname\0 | cigar u32[n_cigar] | seq 4-bit | qual | tags */
    int32_t  l_data, m_data;
} obam_rec;

obam_file *obam_open(const char *path);
void       obam_close(obam_file *f);
obam_hdr  *obam_hdr_read(obam_file *f);
void       obam_hdr_free(obam_hdr *h);
/* returns >=0 on success, -1 at EOF, < -1 on error */
int        obam_read1(obam_file *f, obam_rec *r);
void       obam_rec_free(obam_rec *r);   /* frees r->data only */

static inline char     *obam_qname(const obam_rec *r) { return (char *)r->data; }
static inline uint32_t *obam_cigar(const obam_rec *r) { return (uint32_t *)(r->data + r->l_qname); }
static inline uint8_t  *obam_seq(const obam_rec *r)   { return r->data + r->l_qname + 4 * (int)r->n_cigar; }
static inline uint8_t  *obam_qual(const obam_rec *r)  { return obam_seq(r) + ((r->l_qseq + 1) >> 1); }
static inline int       obam_seqi(const uint8_t *s, int i) { return (s[i >> 1] >> ((~i & 1) << 2)) & 0xf; }

#ifdef __cplusplus
}
#endif
#endif
