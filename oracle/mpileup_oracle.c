/* mpileup_oracle.c -- CPU restatement of `samtools mpileup -f REF [-l BED] -B -b LIST`
 * as invoked by the reference at metaSNV.py:160-165, plus `samtools view -H` (metaSNV.py:83).
 *
 * TEST INFRASTRUCTURE ONLY: may be built/run by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py. The product never calls it.
 *
 * PARITY UNPINNED: samtools/htslib is an un-vendored, un-pinned external dependency of the
 * reference (README.md:19, DEVELOPER.md:46, .github/workflows/main.yml:33-34) and is absent from
 * this image, and the reference holds no test vectors for it. This
 * file restates the published algorithm of samtools 1.9 / htslib 1.9 (released 2018-07-18; the
 * first release with `-d 8000` per file, Annex A; the pileup iterator and the overlap rule are
 * the same in 1.10). Whether later releases changed the overlap rule's handling of two mates that
 * disagree with EQUAL qualities (here, as in 1.9: the earlier read keeps 0.8 x its quality, the
 * later one drops to 0) could not be checked in this image: treat that case as version dependent.
 * The restated functions (bam_plcmd.c: mplp_func,
 * mpileup, pileup_seq; htslib sam.c: bam_plp_push / bam_plp_next / bam_plp_auto / bam_mplp_auto,
 * resolve_cigar2, overlap_push / tweak_overlap_quality) as summarised in SURVEY.md Annex A.
 * It is written as htslib writes it -- a per-file buffered pileup iterator that renders text --
 * which is deliberately a different decomposition from the product (host filter + GPU gather),
 * so the two only agree if both implement the same semantics.
 *
 * Defaults that apply for the reference's command line (SURVEY.md Annex A): exclude flags
 * UNMAP|SECONDARY|QCFAIL|DUP, orphans skipped, min mapq 0, min base quality 13, max depth 8000
 * per file, mate-overlap detection on, BAQ off (-B).
 */
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "obam.h"

#define F_PAIRED 1
#define F_PROPER 2
#define F_UNMAP 4
#define F_MUNMAP 8
#define F_REVERSE 16
#define F_SECONDARY 256
#define F_QCFAIL 512
#define F_DUP 1024

#define MIN_BASEQ 13
#define MAX_DEPTH 8000

/* ---------------------------------------------------------------- nt16 tables (SAMv1 4.2.3) */
static const char NT16_STR[] = "=ACMGRSVTWYHKDBN";
static unsigned char NT16_TABLE[256];
static void init_nt16(void)
{
    memset(NT16_TABLE, 15, 256);
    for (int i = 0; i < 16; ++i) {
        NT16_TABLE[(unsigned char)NT16_STR[i]] = (unsigned char)i;
        NT16_TABLE[(unsigned char)tolower(NT16_STR[i])] = (unsigned char)i;
    }
    NT16_TABLE['0'] = 1; NT16_TABLE['1'] = 2; NT16_TABLE['2'] = 4; NT16_TABLE['3'] = 8;
}

/* ---------------------------------------------------------------- reference FASTA */
typedef struct { char *name; char *seq; int64_t len; } fa_seq;
static fa_seq *g_fa; static int g_nfa;

static void load_fasta(const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f) { fprintf(stderr, "[mpileup_oracle] cannot open reference %s\n", path); exit(1); }
    char *line = NULL; size_t cap = 0; ssize_t n;
    int64_t m = 0; fa_seq *cur = NULL;
    while ((n = getline(&line, &cap, f)) > 0) {
        while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
        if (line[0] == '>') {
            g_fa = (fa_seq *)realloc(g_fa, (g_nfa + 1) * sizeof(fa_seq));
            cur = &g_fa[g_nfa++];
            char *e = line + 1; while (*e && !isspace((unsigned char)*e)) ++e;   /* faidx: name = first word */
            *e = 0;
            cur->name = strdup(line + 1); cur->seq = NULL; cur->len = 0; m = 0;
        } else if (cur) {
            if (cur->len + n + 1 > m) { m = (cur->len + n + 1) * 2; cur->seq = (char *)realloc(cur->seq, m); }
            memcpy(cur->seq + cur->len, line, n); cur->len += n;
        }
    }
    free(line); fclose(f);
}
/* tid -> FASTA record (NULL when the contig is not in the FASTA), built once from the header */
static const fa_seq **g_fa_tid;
static int cmp_fa_name(const void *a, const void *b)
{
    const fa_seq *x = *(const fa_seq *const *)a, *y = *(const fa_seq *const *)b;
    int c = strcmp(x->name, y->name);
    return c ? c : (x < y ? -1 : x > y);          /* equal names: file order, the first one wins */
}
static void index_fasta(int n_targets, char **target_name)
{
    const fa_seq **sorted = (const fa_seq **)malloc((g_nfa ? g_nfa : 1) * sizeof(*sorted));
    for (int i = 0; i < g_nfa; ++i) sorted[i] = &g_fa[i];
    qsort(sorted, g_nfa, sizeof(*sorted), cmp_fa_name);
    g_fa_tid = (const fa_seq **)calloc(n_targets ? n_targets : 1, sizeof(*g_fa_tid));
    for (int t = 0; t < n_targets; ++t) {
        int lo = 0, hi = g_nfa;                    /* lower bound of target_name[t] */
        while (lo < hi) {
            int mid = (lo + hi) / 2;
            if (strcmp(sorted[mid]->name, target_name[t]) < 0) lo = mid + 1; else hi = mid;
        }
        if (lo < g_nfa && !strcmp(sorted[lo]->name, target_name[t])) g_fa_tid[t] = sorted[lo];
    }
    free(sorted);
}

/* ---------------------------------------------------------------- BED (-l) */
typedef struct { char *name; int64_t beg, end; } bed_iv;
static bed_iv *g_bed; static int g_nbed; static int g_has_bed;

static void load_bed(const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f) { fprintf(stderr, "[mpileup_oracle] cannot open %s\n", path); exit(1); }
    char name[4096]; char rest[4096]; char line[8192];
    while (fgets(line, sizeof line, f)) {
        long long a = 0, b = 0;
        int k = sscanf(line, "%4095s %lld %lld", name, &a, &b);
        (void)rest;
        if (k < 2 || name[0] == '#') continue;
        g_bed = (bed_iv *)realloc(g_bed, (g_nbed + 1) * sizeof(bed_iv));
        g_bed[g_nbed].name = strdup(name);
        if (k == 2) { g_bed[g_nbed].beg = a - 1; g_bed[g_nbed].end = a; }     /* 2 columns: 1-based position */
        else        { g_bed[g_nbed].beg = a;     g_bed[g_nbed].end = b; }     /* 3 columns: BED, 0-based half open */
        ++g_nbed;
    }
    fclose(f);
    g_has_bed = 1;
}
/* per-tid interval lists, built once */
typedef struct { int n; int64_t *beg, *end; } bed_list;
static bed_list *g_bed_tid;
static void index_bed(const obam_hdr *h)
{
    g_bed_tid = (bed_list *)calloc(h->n_targets, sizeof(bed_list));
    for (int t = 0; t < h->n_targets; ++t)
        for (int i = 0; i < g_nbed; ++i)
            if (!strcmp(g_bed[i].name, h->target_name[t])) {
                bed_list *L = &g_bed_tid[t];
                L->beg = (int64_t *)realloc(L->beg, (L->n + 1) * sizeof(int64_t));
                L->end = (int64_t *)realloc(L->end, (L->n + 1) * sizeof(int64_t));
                L->beg[L->n] = g_bed[i].beg; L->end[L->n] = g_bed[i].end; ++L->n;
            }
}
static int bed_overlap(int tid, int64_t beg, int64_t end)
{
    const bed_list *L = &g_bed_tid[tid];
    for (int i = 0; i < L->n; ++i) if (L->beg[i] < end && beg < L->end[i]) return 1;
    return 0;
}

/* ---------------------------------------------------------------- per-file pileup iterator */
typedef struct node {
    obam_rec b;              /* deep copy */
    int32_t beg, end;        /* [beg, end) on the reference */
    /* cigar resolver state (resolve_cigar2): k = op index, x = ref coord of op start, y = query coord */
    int k, x, y, s_end;
    struct node *next;
} node;

typedef struct {
    int is_del, is_refskip, is_head, is_tail, indel, qpos;
    node *nd;
} plp1;

typedef struct { char *key; node *val; int used; } oslot;   /* open-addressing qname -> node */

typedef struct {
    obam_file *fp;
    obam_rec rec;            /* scratch for reading */
    node *head, *tail;       /* tail is an empty sentinel, as in htslib */
    int n_nodes;             /* allocated list nodes incl. the sentinel (htslib: mp->cnt) */
    int tid, pos, max_tid, max_pos, is_eof;
    plp1 *plp; int max_plp;
    oslot *ov; int ov_cap, ov_n;
    /* current column handed to the merger */
    int cur_tid, cur_pos, cur_n, has_cur;
} piter;

static const obam_hdr *g_hdr;

static int cigar_rlen(const obam_rec *b)
{
    const uint32_t *c = obam_cigar(b); int l = 0;
    for (int i = 0; i < b->n_cigar; ++i) {
        int op = c[i] & 0xf;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) l += c[i] >> 4;
    }
    return l;
}

/* ---- overlap hash */
static uint32_t str_hash(const char *s) { uint32_t h = 2166136261u; while (*s) { h ^= (unsigned char)*s++; h *= 16777619u; } return h; }
static void ov_grow(piter *it);
static oslot *ov_find(piter *it, const char *key)
{
    if (!it->ov_cap) return NULL;
    uint32_t i = str_hash(key) & (it->ov_cap - 1);
    while (it->ov[i].used) {
        if (it->ov[i].used == 1 && !strcmp(it->ov[i].key, key)) return &it->ov[i];
        i = (i + 1) & (it->ov_cap - 1);
    }
    return NULL;
}
static void ov_put(piter *it, const char *key, node *val)
{
    if ((it->ov_n + 1) * 2 > it->ov_cap) ov_grow(it);
    uint32_t i = str_hash(key) & (it->ov_cap - 1);
    while (it->ov[i].used == 1) i = (i + 1) & (it->ov_cap - 1);
    if (it->ov[i].used == 0) ++it->ov_n;             /* tombstones (2) are reused without growing */
    it->ov[i].key = strdup(key); it->ov[i].val = val; it->ov[i].used = 1;
}
static void ov_del(oslot *s) { free(s->key); s->key = NULL; s->used = 2; }
static void ov_grow(piter *it)
{
    int ocap = it->ov_cap; oslot *o = it->ov;
    it->ov_cap = ocap ? ocap * 2 : 1024;
    it->ov = (oslot *)calloc(it->ov_cap, sizeof(oslot)); it->ov_n = 0;
    for (int i = 0; i < ocap; ++i) if (o[i].used == 1) { ov_put(it, o[i].key, o[i].val); free(o[i].key); }
    free(o);
}
static void overlap_remove(piter *it, const obam_rec *b)
{
    if (b) { oslot *s = ov_find(it, obam_qname(b)); if (s) ov_del(s); }
    else for (int i = 0; i < it->ov_cap; ++i) if (it->ov[i].used == 1) ov_del(&it->ov[i]);
}

/* Walk to the first/next aligned (M,=,X) base at or after reference offset *iref (relative to
 * the read start); returns 0 and sets *iseq, or -1 when the alignment is exhausted.
 * Restates the effect of htslib's cigar_iref2iseq_set/next for M/I/D/N/S/H/P/=/X cigars. */
typedef struct { const uint32_t *c; int n, k; int x, y; } cwalk;   /* x,y = ref/query offset at start of op k */
static void cwalk_init(cwalk *w, const obam_rec *b) { w->c = obam_cigar(b); w->n = b->n_cigar; w->k = 0; w->x = 0; w->y = 0; }
/* smallest aligned reference offset >= want; returns 1 and fills *roff,*qoff or 0 if none */
static int cwalk_seek(cwalk *w, int want, int *roff, int *qoff)
{
    while (w->k < w->n) {
        int op = w->c[w->k] & 0xf, len = w->c[w->k] >> 4;
        if (op == 0 || op == 7 || op == 8) {
            if (want < w->x + len) {
                int r = want > w->x ? want : w->x;
                *roff = r; *qoff = w->y + (r - w->x);
                return 1;
            }
            w->x += len; w->y += len;
        } else if (op == 2 || op == 3) w->x += len;
        else if (op == 1 || op == 4) w->y += len;
        ++w->k;
    }
    return 0;
}

/* tweak_overlap_quality (htslib sam.c): a = earlier mate, b = later mate. */
static void tweak_overlap_quality(obam_rec *a, obam_rec *b)
{
    cwalk wa, wb; cwalk_init(&wa, a); cwalk_init(&wb, b);
    uint8_t *aq = obam_qual(a), *bq = obam_qual(b);
    const uint8_t *as = obam_seq(a), *bs = obam_seq(b);
    int64_t ref = b->pos;                     /* absolute reference coordinate to look at next */
    for (;;) {
        int ra, qa, rb, qb;
        if (!cwalk_seek(&wa, (int)(ref - a->pos), &ra, &qa)) break;
        if (ra + (int64_t)a->pos > ref) ref = ra + (int64_t)a->pos;
        if (!cwalk_seek(&wb, (int)(ref - b->pos), &rb, &qb)) break;
        if (rb + (int64_t)b->pos > ref) { ref = rb + (int64_t)b->pos; continue; }   /* b skips ahead: re-seek a */
        /* both reads have an aligned base at `ref` */
        if (obam_seqi(as, qa) == obam_seqi(bs, qb)) {
            int q = aq[qa] + bq[qb];
            aq[qa] = (uint8_t)(q > 200 ? 200 : q);
            bq[qb] = 0;
        } else if (aq[qa] >= bq[qb]) {
            aq[qa] = (uint8_t)(0.8 * aq[qa]);
            bq[qb] = 0;
        } else {
            bq[qb] = (uint8_t)(0.8 * bq[qb]);
            aq[qa] = 0;
        }
        ++ref;
    }
}

static void overlap_push(piter *it, node *nd)
{
    obam_rec *b = &nd->b;
    if ((b->flag & F_MUNMAP) || !(b->flag & F_PROPER)) return;
    if ((b->mtid >= 0 && b->tid != b->mtid) ||
        (llabs((long long)b->isize) >= 2LL * b->l_qseq && b->mpos >= nd->end)) return;
    oslot *s = ov_find(it, obam_qname(b));
    if (!s) {
        if (b->mpos >= b->pos) ov_put(it, obam_qname(b), nd);   /* only reads whose mate is still to come */
    } else {
        tweak_overlap_quality(&s->val->b, b);
        ov_del(s);
    }
}

static node *node_alloc(piter *it) { ++it->n_nodes; return (node *)calloc(1, sizeof(node)); }
static void node_free(piter *it, node *p) { --it->n_nodes; free(p->b.data); free(p); }

static void rec_copy(obam_rec *dst, const obam_rec *src)
{
    uint8_t *d = (uint8_t *)malloc(src->l_data > 0 ? src->l_data : 1);
    memcpy(d, src->data, src->l_data);
    *dst = *src; dst->data = d; dst->m_data = src->l_data;
}

/* bam_plp_push */
static void plp_push(piter *it, const obam_rec *b)
{
    if (!b) { it->is_eof = 1; return; }
    if (b->tid < 0 || (b->flag & F_UNMAP)) { overlap_remove(it, b); return; }
    if (it->tid == b->tid && it->pos == b->pos && it->n_nodes > MAX_DEPTH) { overlap_remove(it, b); return; }
    node *t = it->tail;
    rec_copy(&t->b, b);
    t->beg = b->pos; t->end = b->pos + cigar_rlen(b);
    t->k = -1; t->x = t->y = 0; t->s_end = t->end - 1;
    if (b->tid < it->max_tid || (b->tid == it->max_tid && t->beg < it->max_pos)) {
        fprintf(stderr, "[mpileup_oracle] the input is not sorted\n"); exit(1);
    }
    it->max_tid = b->tid; it->max_pos = t->beg;
    if (t->end > it->pos || t->b.tid > it->tid) {
        overlap_push(it, t);
        t->next = node_alloc(it);
        it->tail = t->next;
    } else { free(t->b.data); t->b.data = NULL; }
}

/* resolve_cigar2: fill p for node nd at reference position pos (beg <= pos < end). */
static int resolve_cigar(plp1 *p, node *nd, int pos)
{
    obam_rec *b = &nd->b; const uint32_t *cigar = obam_cigar(b);
    int k;
    if (nd->k == -1) {                    /* first time: locate the first reference-consuming op */
        if (b->n_cigar == 1) {
            int op = cigar[0] & 0xf;
            if (op == 0 || op == 7 || op == 8) { nd->k = 0; nd->x = b->pos; nd->y = 0; }
        } else {
            nd->x = b->pos; nd->y = 0;
            for (k = 0; k < b->n_cigar; ++k) {
                int op = cigar[k] & 0xf, l = cigar[k] >> 4;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) break;
                else if (op == 1 || op == 4) nd->y += l;
            }
            nd->k = k;
        }
        if (nd->k == -1) return 0;
    } else {                              /* advance to the op containing pos */
        int l = cigar[nd->k] >> 4;
        if (pos - nd->x >= l) {
            int op = cigar[nd->k] & 0xf;
            /* leave the current op */
            if (op == 0 || op == 7 || op == 8) { nd->x += l; nd->y += l; }
            else if (op == 2 || op == 3) nd->x += l;
            for (k = nd->k + 1; k < b->n_cigar; ++k) {
                op = cigar[k] & 0xf; l = cigar[k] >> 4;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) break;
                else if (op == 1 || op == 4) nd->y += l;
            }
            nd->k = k;
        }
    }
    if (nd->k >= b->n_cigar) return 0;
    {
        int op = cigar[nd->k] & 0xf, l = cigar[nd->k] >> 4;
        p->is_del = p->indel = p->is_refskip = 0;
        if (nd->x + l - 1 == pos && nd->k + 1 < b->n_cigar) {     /* peek at the next op */
            int op2 = cigar[nd->k + 1] & 0xf, l2 = cigar[nd->k + 1] >> 4;
            if (op2 == 2) p->indel = -l2;
            else if (op2 == 1) p->indel = l2;
            else if (op2 == 6 && nd->k + 2 < b->n_cigar) {        /* padding then maybe an insertion */
                int l3 = 0, kk;
                for (kk = nd->k + 2; kk < b->n_cigar; ++kk) {
                    int op3 = cigar[kk] & 0xf;
                    if (op3 == 1) l3 += cigar[kk] >> 4;
                    else if (op3 == 2 || op3 == 0 || op3 == 3 || op3 == 7 || op3 == 8) break;
                }
                if (l3 > 0) p->indel = l3;
            }
        }
        if (op == 0 || op == 7 || op == 8) p->qpos = nd->y + (pos - nd->x);
        else if (op == 2 || op == 3) { p->is_del = 1; p->qpos = nd->y; p->is_refskip = (op == 3); }
        p->is_head = (pos == b->pos); p->is_tail = (pos == nd->s_end);
    }
    p->nd = nd;
    return 1;
}

/* bam_plp_next: returns 1 with a column in it->cur_*, or 0 if more reads are needed / finished */
static int plp_next(piter *it)
{
    if (it->is_eof && it->head == it->tail) return 0;
    while (it->is_eof || it->max_tid > it->tid || (it->max_tid == it->tid && it->max_pos > it->pos)) {
        int n_plp = 0;
        node **pp = &it->head;
        while (*pp != it->tail) {
            node *p = *pp;
            if (p->b.tid < it->tid || (p->b.tid == it->tid && p->end <= it->pos)) {
                overlap_remove(it, &p->b);
                *pp = p->next; node_free(it, p);
            } else {
                if (p->b.tid == it->tid && p->beg <= it->pos) {
                    if (n_plp == it->max_plp) {
                        it->max_plp = it->max_plp ? it->max_plp << 1 : 256;
                        it->plp = (plp1 *)realloc(it->plp, sizeof(plp1) * it->max_plp);
                    }
                    if (resolve_cigar(&it->plp[n_plp], p, it->pos)) ++n_plp;
                }
                pp = &(*pp)->next;
            }
        }
        it->cur_tid = it->tid; it->cur_pos = it->pos; it->cur_n = n_plp;
        if (it->head != it->tail && it->tid < it->head->b.tid) { it->tid = it->head->b.tid; it->pos = it->head->beg; }
        else if (it->head != it->tail && it->pos < it->head->beg) it->pos = it->head->beg;
        else ++it->pos;
        if (n_plp) return 1;
        if (it->is_eof && it->head == it->tail) break;
    }
    return 0;
}

/* mplp_func: next read of this file that passes mpileup's default filters; 0 at EOF */
static int next_read(piter *it)
{
    for (;;) {
        int r = obam_read1(it->fp, &it->rec);
        if (r < 0) {
            if (r < -1) { fprintf(stderr, "[mpileup_oracle] truncated or corrupt BAM\n"); exit(1); }
            return 0;
        }
        obam_rec *b = &it->rec;
        if (b->tid < 0 || (b->flag & F_UNMAP)) continue;
        if (b->flag & (F_UNMAP | F_SECONDARY | F_QCFAIL | F_DUP)) continue;
        if (g_has_bed) {
            int rl = cigar_rlen(b);
            if (!bed_overlap(b->tid, b->pos, b->pos + (rl ? rl : 1))) continue;   /* bam_endpos: rlen 0 counts as 1 */
        }
        {
            const fa_seq *fs = b->tid < g_hdr->n_targets ? g_fa_tid[b->tid] : NULL;
            if (fs && fs->len <= b->pos) {
                fprintf(stderr, "[mpileup_oracle] Skipping because %d is outside of %ld [ref:%d]\n", b->pos, (long)fs->len, b->tid);
                continue;
            }
        }
        /* min mapq 0: nothing to do. Orphans: */
        if ((b->flag & F_PAIRED) && !(b->flag & F_PROPER)) continue;
        return 1;
    }
}

/* bam_plp_auto */
static int plp_auto(piter *it)
{
    if (plp_next(it)) return 1;
    if (it->is_eof) return 0;
    while (next_read(it)) {
        plp_push(it, &it->rec);
        if (plp_next(it)) return 1;
    }
    plp_push(it, NULL);
    return plp_next(it);
}

/* ---------------------------------------------------------------- text rendering (pileup_seq) */
static char *g_out; static size_t g_on, g_om;
static inline void oputc(int c) { if (g_on == g_om) { g_om = g_om ? g_om * 2 : 1 << 20; g_out = (char *)realloc(g_out, g_om); } g_out[g_on++] = (char)c; }
static void oputs(const char *s) { while (*s) oputc(*s++); }
static void oputi(long v) { char b[32]; snprintf(b, sizeof b, "%ld", v); oputs(b); }

static void pileup_seq(const plp1 *p, int pos, int64_t ref_len, const char *ref)
{
    const obam_rec *b = &p->nd->b;
    int rev = (b->flag & F_REVERSE) != 0, j;
    if (p->is_head) { oputc('^'); oputc(b->mapq > 93 ? 126 : b->mapq + 33); }
    if (!p->is_del) {
        int c = p->qpos < b->l_qseq ? NT16_STR[obam_seqi(obam_seq(b), p->qpos)] : 'N';
        if (ref) {
            int rb = pos < ref_len ? ref[pos] : 'N';
            if (c == '=' || NT16_TABLE[c] == NT16_TABLE[rb]) c = rev ? ',' : '.';
            else c = rev ? tolower(c) : toupper(c);
        } else {
            if (c == '=') c = rev ? ',' : '.';
            else c = rev ? tolower(c) : toupper(c);
        }
        oputc(c);
    } else oputc(p->is_refskip ? (rev ? '<' : '>') : '*');
    if (p->indel > 0) {
        oputc('+'); oputi(p->indel);
        for (j = 1; j <= p->indel; ++j) {
            int c = NT16_STR[obam_seqi(obam_seq(b), p->qpos + j)];
            oputc(rev ? tolower(c) : toupper(c));
        }
    } else if (p->indel < 0) {
        oputc('-'); oputi(-p->indel);
        for (j = 1; j <= -p->indel; ++j) {
            int c = (ref && (int64_t)pos + j < ref_len) ? ref[pos + j] : 'N';
            oputc(rev ? tolower(c) : toupper(c));
        }
    }
    if (p->is_tail) oputc('$');
}

/* ---------------------------------------------------------------- main */
static int view_header(const char *path)
{
    obam_file *f = obam_open(path);
    if (!f) { fprintf(stderr, "[mpileup_oracle] cannot open %s\n", path); return 1; }
    obam_hdr *h = obam_hdr_read(f);
    if (!h) { fprintf(stderr, "[mpileup_oracle] cannot read header of %s\n", path); return 1; }
    fwrite(h->text, 1, strlen(h->text), stdout);
    obam_hdr_free(h); obam_close(f);
    return 0;
}

int main(int argc, char **argv)
{
    const char *ref_path = NULL, *bed_path = NULL, *list_path = NULL;
    if (argc >= 4 && !strcmp(argv[1], "view") && !strcmp(argv[2], "-H")) return view_header(argv[3]);
    if (argc < 2 || strcmp(argv[1], "mpileup")) {
        fprintf(stderr, "usage: mpileup_oracle mpileup -f REF [-l BED] -B -b LIST | mpileup_oracle view -H BAM\n");
        return 1;
    }
    for (int i = 2; i < argc; ++i) {
        if (!strcmp(argv[i], "-f") && i + 1 < argc) ref_path = argv[++i];
        else if (!strcmp(argv[i], "-l") && i + 1 < argc) bed_path = argv[++i];
        else if (!strcmp(argv[i], "-b") && i + 1 < argc) list_path = argv[++i];
        else if (!strcmp(argv[i], "-B")) { }
        else { fprintf(stderr, "[mpileup_oracle] unsupported argument %s\n", argv[i]); return 1; }
    }
    if (!list_path) { fprintf(stderr, "[mpileup_oracle] -b LIST is required\n"); return 1; }
    init_nt16();
    if (ref_path) load_fasta(ref_path);
    if (bed_path) load_bed(bed_path);

    /* open all files */
    int n = 0; piter *its = NULL;
    {
        FILE *lf = fopen(list_path, "r");
        if (!lf) { fprintf(stderr, "[mpileup_oracle] cannot open %s\n", list_path); return 1; }
        char line[8192];
        while (fgets(line, sizeof line, lf)) {
            size_t l = strlen(line);
            while (l && (line[l - 1] == '\n' || line[l - 1] == '\r' || line[l - 1] == ' ')) line[--l] = 0;
            if (!l) continue;
            its = (piter *)realloc(its, (n + 1) * sizeof(piter));
            memset(&its[n], 0, sizeof(piter));
            its[n].fp = obam_open(line);
            if (!its[n].fp) { fprintf(stderr, "[mpileup_oracle] cannot open %s\n", line); return 1; }
            obam_hdr *h = obam_hdr_read(its[n].fp);
            if (!h) { fprintf(stderr, "[mpileup_oracle] cannot read header of %s\n", line); return 1; }
            if (n == 0) g_hdr = h; else obam_hdr_free(h);       /* header of the first file is used throughout */
            its[n].max_tid = its[n].max_pos = -1;
            its[n].head = its[n].tail = node_alloc(&its[n]);
            ++n;
        }
        fclose(lf);
    }
    if (n == 0) return 0;
    if (g_has_bed) index_bed(g_hdr);
    index_fasta(g_hdr->n_targets, g_hdr->target_name);

    /* bam_mplp_auto: merge the per-file columns by (tid,pos) */
    uint64_t *ipos = (uint64_t *)malloc(n * sizeof(uint64_t));
    uint64_t min = (uint64_t)-1;
    for (int i = 0; i < n; ++i) ipos[i] = (uint64_t)-1;
    int cached_tid = -1; const fa_seq *cached_fa = NULL;
    setvbuf(stdout, NULL, _IOFBF, 1 << 20);
    for (;;) {
        uint64_t new_min = (uint64_t)-1;
        for (int i = 0; i < n; ++i) {
            if (ipos[i] == min) {
                its[i].has_cur = plp_auto(&its[i]);
                ipos[i] = its[i].has_cur ? ((uint64_t)its[i].cur_tid << 32 | (uint32_t)its[i].cur_pos) : 0;
            }
            if (its[i].has_cur && ipos[i] < new_min) new_min = ipos[i];
        }
        min = new_min;
        if (new_min == (uint64_t)-1) break;
        int tid = (int)(new_min >> 32), pos = (int)(uint32_t)new_min;
        if (g_has_bed && !bed_overlap(tid, pos, pos + 1)) continue;
        if (tid != cached_tid) { cached_tid = tid; cached_fa = tid < g_hdr->n_targets ? g_fa_tid[tid] : NULL; }
        const char *ref = cached_fa ? cached_fa->seq : NULL;
        int64_t ref_len = cached_fa ? cached_fa->len : 0;
        g_on = 0;
        oputs(g_hdr->target_name[tid]); oputc('\t'); oputi(pos + 1); oputc('\t');
        oputc((ref && pos < ref_len) ? ref[pos] : 'N');
        for (int i = 0; i < n; ++i) {
            int n_plp = (its[i].has_cur && ipos[i] == min) ? its[i].cur_n : 0;
            int cnt = 0;
            for (int j = 0; j < n_plp; ++j) {
                const plp1 *p = &its[i].plp[j];
                int c = p->qpos < p->nd->b.l_qseq ? obam_qual(&p->nd->b)[p->qpos] : 0;
                if (c >= MIN_BASEQ) ++cnt;
            }
            oputc('\t'); oputi(cnt); oputc('\t');
            if (n_plp == 0) { oputs("*\t*"); continue; }
            int shown = 0;
            for (int j = 0; j < n_plp; ++j) {
                const plp1 *p = &its[i].plp[j];
                int c = p->qpos < p->nd->b.l_qseq ? obam_qual(&p->nd->b)[p->qpos] : 0;
                if (c >= MIN_BASEQ) { ++shown; pileup_seq(p, pos, ref_len, ref); }
            }
            if (!shown) oputc('*');
            oputc('\t');
            shown = 0;
            for (int j = 0; j < n_plp; ++j) {
                const plp1 *p = &its[i].plp[j];
                int c = p->qpos < p->nd->b.l_qseq ? obam_qual(&p->nd->b)[p->qpos] : 0;
                if (c >= MIN_BASEQ) { c += 33; if (c > 126) c = 126; oputc(c); ++shown; }
            }
            if (!shown) oputc('*');
        }
        oputc('\n');
        fwrite(g_out, 1, g_on, stdout);
    }
    fflush(stdout);
    return 0;
}
