/* obam.c -- see obam.h. TEST INFRASTRUCTURE ONLY (oracle). SAMv1 spec section 4. */
#include "obam.h"
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

struct obam_file {
    FILE    *fp;
    uint8_t *cbuf;          /* one compressed BGZF block  (<= 64 KiB) */
    uint8_t *ubuf;          /* one inflated BGZF block    (<= 64 KiB) */
    int      ulen, upos;
    int      eof, err;
};

static uint32_t le32(const uint8_t *p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24; }
static uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }

obam_file *obam_open(const char *path)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return NULL;
    obam_file *f = (obam_file *)calloc(1, sizeof(*f));
    f->fp = fp;
    f->cbuf = (uint8_t *)malloc(1 << 16);
    f->ubuf = (uint8_t *)malloc(1 << 16);
    return f;
}

void obam_close(obam_file *f)
{
    if (!f) return;
    fclose(f->fp); free(f->cbuf); free(f->ubuf); free(f);
}

/* Load and inflate the next BGZF member. Returns 0 ok, -1 EOF, -2 error. */
static int next_block(obam_file *f)
{
    uint8_t hdr[12];
    for (;;) {
        size_t n = fread(hdr, 1, 12, f->fp);
        if (n == 0) { f->eof = 1; return -1; }
        if (n != 12 || hdr[0] != 0x1f || hdr[1] != 0x8b || hdr[2] != 8 || !(hdr[3] & 4)) return f->err = -2;
        int xlen = le16(hdr + 10);
        if (fread(f->cbuf, 1, xlen, f->fp) != (size_t)xlen) return f->err = -2;
        int bsize = -1;
        for (int off = 0; off + 4 <= xlen; ) {           /* find the 'BC' subfield */
            int slen = le16(f->cbuf + off + 2);
            if (f->cbuf[off] == 'B' && f->cbuf[off + 1] == 'C' && slen == 2) bsize = le16(f->cbuf + off + 4);
            off += 4 + slen;
        }
        if (bsize < 0) return f->err = -2;
        int clen = bsize + 1 - 12 - xlen;                 /* deflate payload + crc32 + isize */
        if (clen < 8 || fread(f->cbuf, 1, clen, f->fp) != (size_t)clen) return f->err = -2;
        uint32_t isize = le32(f->cbuf + clen - 4), crc = le32(f->cbuf + clen - 8);
        if (isize > (1u << 16)) return f->err = -2;
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) return f->err = -2;
        zs.next_in = f->cbuf; zs.avail_in = clen - 8;
        zs.next_out = f->ubuf; zs.avail_out = 1 << 16;
        int zr = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (zr != Z_STREAM_END || zs.total_out != isize) return f->err = -2;
        if ((uint32_t)crc32(crc32(0L, NULL, 0), f->ubuf, isize) != crc) return f->err = -2;
        f->ulen = (int)isize; f->upos = 0;
        if (isize) return 0;                              /* empty blocks (EOF marker) are skipped */
    }
}

/* Read exactly n bytes of the uncompressed stream. Returns n, 0 at clean EOF, <0 error. */
static int uread(obam_file *f, void *dst, int n)
{
    uint8_t *d = (uint8_t *)dst; int got = 0;
    while (got < n) {
        if (f->upos == f->ulen) {
            int r = next_block(f);
            if (r == -1) return got == 0 ? 0 : -2;
            if (r < 0) return r;
        }
        int k = f->ulen - f->upos; if (k > n - got) k = n - got;
        memcpy(d + got, f->ubuf + f->upos, k);
        f->upos += k; got += k;
    }
    return got;
}

obam_hdr *obam_hdr_read(obam_file *f)
{
    uint8_t b[8];
    if (uread(f, b, 8) != 8 || memcmp(b, "BAM\1", 4)) return NULL;
    obam_hdr *h = (obam_hdr *)calloc(1, sizeof(*h));
    h->l_text = le32(b + 4);
    h->text = (char *)malloc(h->l_text + 1);
    if (h->l_text && uread(f, h->text, h->l_text) != (int)h->l_text) { obam_hdr_free(h); return NULL; }
    h->text[h->l_text] = 0;
    if (uread(f, b, 4) != 4) { obam_hdr_free(h); return NULL; }
    h->n_targets = (int32_t)le32(b);
    h->target_name = (char **)calloc(h->n_targets ? h->n_targets : 1, sizeof(char *));
    h->target_len = (uint32_t *)calloc(h->n_targets ? h->n_targets : 1, sizeof(uint32_t));
    for (int i = 0; i < h->n_targets; ++i) {
        if (uread(f, b, 4) != 4) { obam_hdr_free(h); return NULL; }
        int l = (int)le32(b);
        h->target_name[i] = (char *)malloc(l + 1);
        if (uread(f, h->target_name[i], l) != l) { obam_hdr_free(h); return NULL; }
        h->target_name[i][l] = 0;
        if (uread(f, b, 4) != 4) { obam_hdr_free(h); return NULL; }
        h->target_len[i] = le32(b);
    }
    return h;
}

void obam_hdr_free(obam_hdr *h)
{
    if (!h) return;
    if (h->target_name) for (int i = 0; i < h->n_targets; ++i) free(h->target_name[i]);
    free(h->target_name); free(h->target_len); free(h->text); free(h);
}

int obam_read1(obam_file *f, obam_rec *r)
{
    uint8_t b[36];
    int n = uread(f, b, 4);
    if (n == 0) return -1;
    if (n != 4) return -2;
    int32_t block_size = (int32_t)le32(b);
    if (block_size < 32) return -3;
    if (uread(f, b + 4, 32) != 32) return -2;
    r->tid = (int32_t)le32(b + 4);  r->pos = (int32_t)le32(b + 8);
    int l_read_name = b[12];        r->mapq = b[13];
    r->bin = le16(b + 14);          r->n_cigar = le16(b + 16);
    r->flag = le16(b + 18);         r->l_qseq = (int32_t)le32(b + 20);
    r->mtid = (int32_t)le32(b + 24); r->mpos = (int32_t)le32(b + 28);
    r->isize = (int32_t)le32(b + 32);
    int rest = block_size - 32;
    int pad = (4 - (l_read_name & 3)) & 3;               /* keep the CIGAR words 4-byte aligned */
    int need = rest + pad;
    if (need > r->m_data) {
        r->m_data = need + 64;
        r->data = (uint8_t *)realloc(r->data, r->m_data);
    }
    if (uread(f, r->data, l_read_name) != l_read_name) return -2;
    memset(r->data + l_read_name, 0, pad);
    if (uread(f, r->data + l_read_name + pad, rest - l_read_name) != rest - l_read_name) return -2;
    r->l_qname = (uint16_t)(l_read_name + pad);
    r->l_data = need;
    return need;
}

void obam_rec_free(obam_rec *r) { free(r->data); r->data = NULL; r->m_data = r->l_data = 0; }
