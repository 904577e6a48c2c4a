/*
 * th_host.c -- host C above the GPU C ABI: derives the reference's floating-point fields from the
 * integer results, applies the adapter logic and record filters, and formats the output text.
 * Compiled with -ffp-contract=off so that every expression rounds like the reference's plain SSE2
 * doubles ("Host floating point").
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>
#include <pthread.h>
#include <time.h>
#include <unistd.h>
#include "th_host.h"

#define TH_SLOTS 4096 /* CHUNK_READ_N, src/tidehunter.h:10: tandem_seq_t slots are reused every 4096 reads */
#define TH_MAX_LANES 64 /* e.g. 8 GPUs x 8 contexts */

typedef struct { char *s; size_t l, m; } str_t;

struct th_host {
    th_host_para p;
    th_gpu_ctx *gpu;                        /* = lane[0] */
    th_gpu_ctx *lane[TH_MAX_LANES]; int n_lanes;
    long long n_failed;       /* consensus tasks the GPU path reported as failed since th_host_create (their records are missing) */
    char *five_rc, *three_rc; int five_len, three_len;
    str_t out;
    str_t qual[TH_SLOTS];     /* persistent quality buffers: qual.l is never reset in the reference */
    int64_t read_counter;
    th_gpu_stats stats;
};

static char g_err[1024];
const char *th_host_last_error(void) { return g_err; }
static void set_err(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }

void th_host_default_para(th_host_para *p) {
    memset(p, 0, sizeof(*p));
    th_gpu_default_params(&p->gpu);
    p->out_fmt = 1; p->min_len = 30; p->ada_match_rat = 0.8f; p->chunk_reads = 4096;
}

static void str_reserve(str_t *s, size_t extra) {
    if (s->l + extra + 1 > s->m) { s->m = (s->l + extra + 1) * 2; s->s = (char *)realloc(s->s, s->m); }
}
static void str_write(str_t *s, const char *d, size_t n) { str_reserve(s, n); memcpy(s->s + s->l, d, n); s->l += n; s->s[s->l] = 0; }
static void str_printf(str_t *s, const char *fmt, ...) {
    va_list ap; int n;
    str_reserve(s, 128);
    va_start(ap, fmt); n = vsnprintf(s->s + s->l, s->m - s->l, fmt, ap); va_end(ap);
    if ((size_t)n >= s->m - s->l) { str_reserve(s, (size_t)n + 1); va_start(ap, fmt); n = vsnprintf(s->s + s->l, s->m - s->l, fmt, ap); va_end(ap); }
    s->l += n;
}

static char *revcomp(const char *s, int l) { /* src/seq.c:89-95 */
    char *r = (char *)malloc(l + 1); int i;
    for (i = 0; i < l; ++i) {
        char c = s[i], o = 'N';
        switch (c) { case 'A': case 'a': o = 'T'; break; case 'C': case 'c': o = 'G'; break; case 'G': case 'g': o = 'C'; break; case 'T': case 't': o = 'A'; break; default: o = 'N'; }
        r[l - i - 1] = o;
    }
    r[l] = 0;
    return r;
}

th_host *th_host_create(const th_host_para *p, int device) { return th_host_create_multi(p, 1, &device); }

/* One process, several GPUs: `lanes` contexts on each device, chunk c of a th_host_run goes to lane c % n_lanes and
 * consecutive lanes sit on different devices.  Reads are independent, so this is the whole multi-GPU story of the
 * command line front end: no exchange between devices, results are formatted in input order by the caller's thread. */
th_host *th_host_create_multi(const th_host_para *p, int n_devices, const int *devices) {
    th_host *h = (th_host *)calloc(1, sizeof(th_host));
    if (n_devices < 1 || !devices) { set_err("th_host_create_multi: no device given"); free(h); return NULL; }
    h->p = *p;
    h->p.gpu.need_cov = (p->out_fmt == 3 || p->out_fmt == 4 || p->min_cov > 0 || p->min_frac > 0.0) ? 1 : 0;
    if (h->p.chunk_reads <= 0) h->p.chunk_reads = 4096;
    if (h->p.lanes <= 0) { const char *e = getenv("TH_HOST_LANES"); h->p.lanes = e ? atoi(e) : 4; } /* 4 contexts x 4096 reads: the best end-to-end setting measured on B200 (profiles/) */
    if (h->p.lanes < 1) h->p.lanes = 1;
    h->p.lanes *= n_devices;
    if (h->p.lanes > TH_MAX_LANES) h->p.lanes = TH_MAX_LANES;
    for (h->n_lanes = 0; h->n_lanes < h->p.lanes; ++h->n_lanes) {
        { const int dev = devices[h->n_lanes % n_devices];
          h->lane[h->n_lanes] = th_gpu_create(&h->p.gpu, dev < 0 ? 0 : dev); }
        if (!h->lane[h->n_lanes]) { int i; set_err("%s", th_gpu_last_error()); for (i = 0; i < h->n_lanes; ++i) th_gpu_destroy(h->lane[i]); free(h); return NULL; }
    }
    h->gpu = h->lane[0];
    if (p->five_seq && p->three_seq) {
        h->five_len = (int)strlen(p->five_seq); h->three_len = (int)strlen(p->three_seq);
        h->p.five_seq = strdup(p->five_seq); h->p.three_seq = strdup(p->three_seq);
        h->five_rc = revcomp(p->five_seq, h->five_len); h->three_rc = revcomp(p->three_seq, h->three_len);
    } else { h->p.five_seq = h->p.three_seq = NULL; }
    return h;
}
void th_host_destroy(th_host *h) {
    int i;
    if (!h) return;
    for (i = 0; i < h->n_lanes; ++i) th_gpu_destroy(h->lane[i]);
    free(h->five_rc); free(h->three_rc); free((void *)h->p.five_seq); free((void *)h->p.three_seq);
    for (i = 0; i < TH_SLOTS; ++i) free(h->qual[i].s);
    free(h->out.s); free(h);
}
void th_host_stats(const th_host *h, th_gpu_stats *s) { *s = h->stats; }
long long th_host_failed_tasks(const th_host *h) { return h->n_failed; }
void th_host_set_read_index(th_host *h, long long first) { h->read_counter = first; }
th_gpu_ctx *th_host_gpu(th_host *h) { return h->gpu; }

/* Infix edit distance with threshold (edlib_align_HW, src/edlib_align.c:73-85): edit distance,
 * first end location, and for it the smallest start reaching that distance (edlib/src/edlib.cpp:141-236).
 * Case-insensitive equality.  infix_ed_plain is the column-by-column definition; infix_ed evaluates the same two
 * passes with Myers' bit-vector recurrences (Myers 1999, block form of Hyyro 2003: vertical deltas Pv/Mv per 64 rows,
 * horizontal delta carried from block to block), 64 DP cells per word operation. */
/* character classes of the adapter search: edlib is given the five equalities a/A c/C g/G t/T n/N (src/edlib_align.c:21-27)
 * and compares every other byte verbatim, so 'R' and 'r' (or '@' and '`') are different symbols */
static unsigned char ed_class[256];
static unsigned char nt4_class[256]; /* the read's nt4 code (src/seq.c:15-32): what abpoa_gen_cons compares */
__attribute__((constructor)) static void init_classes(void) { /* filled when the library is loaded */
    int c;
    for (c = 0; c < 256; ++c) { ed_class[c] = (unsigned char)c; nt4_class[c] = 4; }
    ed_class['a'] = 'A'; ed_class['c'] = 'C'; ed_class['g'] = 'G'; ed_class['t'] = 'T'; ed_class['n'] = 'N';
    nt4_class[0] = 0; nt4_class[1] = 1; nt4_class[2] = 2; nt4_class[3] = 3; nt4_class['-'] = 5;
    nt4_class['A'] = nt4_class['a'] = 0; nt4_class['C'] = nt4_class['c'] = 1; nt4_class['G'] = nt4_class['g'] = 2; nt4_class['T'] = nt4_class['t'] = 3;
}
#define ED_EQ(a, b) (ed_class[(unsigned char)(a)] == ed_class[(unsigned char)(b)])

static int infix_ed_plain(const char *q, int ql, const char *t, int tl, int *start, int *end, int k) {
    int i, j, best = -1, best_end = -1, best_start = -1;
    int *col;
    if (ql <= 0 || tl <= 0) return -1;
    col = (int *)malloc(sizeof(int) * (ql + 1));
    for (i = 0; i <= ql; ++i) col[i] = i;
    for (j = 0; j < tl; ++j) {
        int diag = col[0];
        col[0] = 0;
        for (i = 1; i <= ql; ++i) {
            int up = col[i - 1] + 1, left = col[i] + 1, d = diag + (ED_EQ(q[i - 1], t[j]) ? 0 : 1);
            int v = d < up ? d : up;
            if (left < v) v = left;
            diag = col[i]; col[i] = v;
        }
        if (best < 0 || col[ql] < best) { best = col[ql]; best_end = j; }
    }
    if (best > ql) best = ql;
    if (k >= 0 && best > k) { free(col); return -1; }
    for (i = 0; i <= ql; ++i) col[i] = i;
    for (j = 0; j <= best_end; ++j) {
        int diag = col[0];
        col[0] = j + 1;
        for (i = 1; i <= ql; ++i) {
            int up = col[i - 1] + 1, left = col[i] + 1, d = diag + (ED_EQ(q[ql - i], t[best_end - j]) ? 0 : 1);
            int v = d < up ? d : up;
            if (left < v) v = left;
            diag = col[i]; col[i] = v;
        }
        if (col[ql] == best) best_start = best_end - j;
    }
    free(col);
    *start = best_start; *end = best_end;
    return best;
}

#define ED_MAXW 4 /* adapters up to 256 bases take the bit-vector path */
/* one text column through the W blocks; hin0 = horizontal delta along the top row (0: a match may start anywhere,
 * +1: the top row counts text characters); returns the delta of the bottom row (pattern row ql) */
static inline int ed_column(int W, uint64_t last_bit, const uint64_t *eq, uint64_t *Pv, uint64_t *Mv, int hin0) {
    int hin = hin0, b;
    for (b = 0; b < W; ++b) {
        const uint64_t top = b == W - 1 ? last_bit : (uint64_t)1 << 63;
        uint64_t Eq = eq[b], pv = Pv[b], mv = Mv[b], Xv, Xh, Ph, Mh;
        int hout = 0;
        Xv = Eq | mv;
        if (hin < 0) Eq |= 1;
        Xh = (((Eq & pv) + pv) ^ pv) | Eq;
        Ph = mv | ~(Xh | pv);
        Mh = pv & Xh;
        if (Ph & top) hout = 1; else if (Mh & top) hout = -1;
        Ph <<= 1; Mh <<= 1;
        if (hin < 0) Mh |= 1; else if (hin > 0) Ph |= 1;
        Pv[b] = Mh | ~(Xv | Ph);
        Mv[b] = Ph & Xv;
        hin = hout;
    }
    return hin;
}
static int infix_ed(const char *q, int ql, const char *t, int tl, int *start, int *end, int k) {
    uint64_t peq[256][ED_MAXW], Pv[ED_MAXW], Mv[ED_MAXW];
    int W, i, j, b, score, best = -1, best_end = -1, best_start = -1;
    uint64_t last_bit;
    if (ql <= 0 || tl <= 0) return -1;
    if (ql > 64 * ED_MAXW) return infix_ed_plain(q, ql, t, tl, start, end, k);
    W = (ql + 63) / 64; last_bit = (uint64_t)1 << ((ql - 1) & 63);
    memset(peq, 0, sizeof(peq));
    for (i = 0; i < ql; ++i) peq[ed_class[(unsigned char)q[i]]][i >> 6] |= (uint64_t)1 << (i & 63);
    for (b = 0; b < W; ++b) { Pv[b] = ~(uint64_t)0; Mv[b] = 0; }
    score = ql;
    for (j = 0; j < tl; ++j) {
        score += ed_column(W, last_bit, peq[ed_class[(unsigned char)t[j]]], Pv, Mv, 0);
        if (best < 0 || score < best) { best = score; best_end = j; }
    }
    if (best > ql) best = ql;
    if (k >= 0 && best > k) return -1;
    /* backwards from the end location: reversed adapter against the reversed text prefix, top row counting text */
    memset(peq, 0, sizeof(peq));
    for (i = 0; i < ql; ++i) peq[ed_class[(unsigned char)q[ql - 1 - i]]][i >> 6] |= (uint64_t)1 << (i & 63);
    for (b = 0; b < W; ++b) { Pv[b] = ~(uint64_t)0; Mv[b] = 0; }
    score = ql;
    for (j = 0; j <= best_end; ++j) {
        score += ed_column(W, last_bit, peq[ed_class[(unsigned char)t[best_end - j]]], Pv, Mv, 1);
        if (score == best) best_start = best_end - j;
    }
    *start = best_start; *end = best_end;
    return best;
}

/* self-test hook for tests/: random adapters and texts through both implementations; returns the number of disagreements */
int th_host_selftest_infix(int trials, unsigned seed) {
    int bad = 0, it;
    unsigned long long x = seed * 2654435761ull + 88172645463325252ull;
#define RND() (x ^= x << 13, x ^= x >> 7, x ^= x << 17, (unsigned)(x >> 11))
    for (it = 0; it < trials; ++it) {
        const int ql = 1 + (int)(RND() % (it % 7 == 0 ? 300 : 130));
        const int tl = 1 + (int)(RND() % (it % 5 == 0 ? 3000 : 400));
        const int sigma = 2 + (int)(RND() % 4);
        char *q = (char *)malloc((size_t)ql + 1), *t = (char *)malloc((size_t)tl + 1);
        int i, s1 = -7, e1 = -7, s2 = -7, e2 = -7, r1, r2, k;
        static const char al[] = "ACGTNacgtn";
        for (i = 0; i < ql; ++i) { const unsigned c = RND() % sigma, lower = RND() % 4 == 0; q[i] = al[c + (lower ? 5 : 0)]; }
        for (i = 0; i < tl; ++i) { const unsigned c = RND() % sigma, lower = RND() % 4 == 0; t[i] = al[c + (lower ? 5 : 0)]; }
        if (it % 3 == 0 && tl > ql + 10) { /* plant a noisy copy of the adapter */
            const int off = (int)(RND() % (unsigned)(tl - ql));
            for (i = 0; i < ql; ++i) if (RND() % 8) t[off + i] = q[i];
        }
        k = it % 4 == 0 ? -1 : (int)(ql * (RND() % 100) / 100.0);
        r1 = infix_ed_plain(q, ql, t, tl, &s1, &e1, k); r2 = infix_ed(q, ql, t, tl, &s2, &e2, k);
        if (r1 != r2 || (r1 >= 0 && (s1 != s2 || e1 != e2))) ++bad;
        free(q); free(t);
    }
#undef RND
    return bad;
}

/* -s: collect_ed_res (src/gen_cons.c:89-110) -- the best infix hit of an adapter in the read and one more on each side */
typedef struct { int ed, start, end; } ed_res_t;
static int collect_ed_res(float ada_match_rat, const char *q, int qlen, const char *seq, int seq_len, ed_res_t *res) {
    int n = 0, ed, start = 0, end = 0, k = (int)(qlen * (1 - ada_match_rat));
    ed = infix_ed(q, qlen, seq, seq_len, &start, &end, k);
    if (ed != -1) {
        res[0].ed = ed; res[0].start = start; res[0].end = end; n++;
        if (res[0].start >= qlen) {
            ed = infix_ed(q, qlen, seq, res[0].start, &start, &end, k);
            if (ed != -1) { res[n].ed = ed; res[n].start = start; res[n].end = end; n++; }
        }
        if (res[0].end <= seq_len - qlen) {
            ed = infix_ed(q, qlen, seq + res[0].end, seq_len - res[0].end, &start, &end, k);
            if (ed != -1) { res[n].ed = ed; res[n].start = res[0].end + start; res[n].end = res[0].end + end; n++; }
        }
    }
    return n;
}
/* get_full_len_seq (src/gen_cons.c:112-126): adapter pair with the smallest total distance and >= min_len between them */
static int full_len_pair(int min_len, int left_n, const ed_res_t *left, int right_n, const ed_res_t *right, int *tar_start, int *tar_end) {
    int tot_ed = INT32_MAX, i, j;
    for (i = 0; i < left_n; ++i)
        for (j = 0; j < right_n; ++j)
            if (right[j].start - left[i].end - 1 >= min_len && tot_ed > left[i].ed + right[j].ed) {
                tot_ed = left[i].ed + right[j].ed;
                *tar_start = left[i].end + 1; *tar_end = right[j].start - 1;
            }
    return tot_ed;
}

/* Adapter trimming of a consensus (src/gen_cons.c:224-291).  The consensus of a tandem repeat is circular -- a unit may
 * start anywhere in it -- so both adapters are searched in two copies laid end to end, and the insert is what lies between
 * the upstream adapter's end and the downstream adapter's start.  Two orientations are tried, in this order:
 *   1: 5' adapter ... reverse complement of the 3' adapter      2: 3' adapter ... reverse complement of the 5' adapter
 * The second one is only looked at when the first left a non-zero total distance, and only wins with a strictly smaller
 * one (an orientation-1 hit pair that does not delimit an insert sets no distance to beat).  Returns the orientation
 * taken (0: none; the consensus is left alone) and cuts cons_seq / cons_qual / *cons_len down to the insert. */
typedef struct { int ed, start, end; } ada_hit_t;
static int ada_find(const char *ada, int ada_len, float match_rat, const char *text, int text_len, ada_hit_t *hit) {
    hit->start = hit->end = -1;
    hit->ed = infix_ed(ada, ada_len, text, text_len, &hit->start, &hit->end, (int)(ada_len * (1 - match_rat)));
    return hit->ed != -1;
}
static int trim_to_adapters(const th_host *h, char *cons_seq, uint8_t *cons_qual, int *cons_len_io) {
    const th_host_para *p = &h->p;
    const int cons_len = *cons_len_io, twice = cons_len << 1;
    const struct { const char *up; int up_len; const char *down; int down_len; } ori[2] = {
        {p->five_seq, h->five_len, h->three_rc, h->three_len},
        {p->three_seq, h->three_len, h->five_rc, h->five_len}};
    char *ring = (char *)malloc((size_t)twice + 1);
    int o, taken = 0, to_beat = INT32_MAX, ins_start = -1, ins_end = -1;
    memcpy(ring, cons_seq, cons_len); memcpy(ring + cons_len, cons_seq, cons_len); ring[twice] = 0;
    for (o = 0; o < 2 && to_beat != 0; ++o) {
        ada_hit_t up, down;
        int s, e;
        if (!ada_find(ori[o].up, ori[o].up_len, p->ada_match_rat, ring, twice, &up)) continue;
        if (!ada_find(ori[o].down, ori[o].down_len, p->ada_match_rat, ring, twice, &down)) continue;
        if (up.ed + down.ed >= to_beat) continue;
        s = up.end + 1;
        if (down.start > up.end) e = down.start - 1;                       /* downstream adapter in the same turn */
        else if (down.end + cons_len < twice && down.start + cons_len > up.end) e = down.start + cons_len - 1; /* one turn later */
        else continue;
        ins_start = s; ins_end = e; taken = o + 1;
        if (o == 0) to_beat = up.ed + down.ed;
    }
    if (ins_start > 0 && ins_end > ins_start) {
        const int n = ins_end - ins_start + 1;
        int k;
        memcpy(cons_seq, ring + ins_start, n); cons_seq[n] = 0;
        if (cons_qual) { /* same cut of the doubled quality string; cons_qual has room for two turns */
            memcpy(cons_qual + cons_len, cons_qual, cons_len);
            for (k = 0; k < n; ++k) cons_qual[k] = cons_qual[ins_start + k];
        }
        *cons_len_io = n;
    }
    free(ring);
    return taken;
}

typedef struct { /* one record of tandem_seq_t */
    int cons_start, cons_end, cons_len, full_length, pos_n;
    double copy_num, ave_match;
    const int32_t *sub_pos;
    size_t seq_off;          /* into the per-read consensus text */
} rec_t;

/* one read: turn its tasks into records (seqs_msa tail + write_tandem_cons_seq) and print them */
static void emit_read(th_host *h, str_t *out, long long *n_failed, const th_gpu_result *R, int r, const char *name, const char *seq, int len, int64_t global_index) {
    const th_host_para *p = &h->p;
    const int with_qual = (p->out_fmt == 3 || p->out_fmt == 4);
    int t, i, n_rec = 0, m_rec = 0;
    int32_t sc_pos[2] = {0, 0};
    rec_t *rec = NULL;
    str_t cons_txt = {0, 0, 0};
    str_t *qs = &h->qual[global_index % TH_SLOTS];
    size_t qual_base = qs->l; (void)qual_base;
    if (R->read_status && R->read_status[r] != 0 && R->read_task_off[r + 1] == R->read_task_off[r]) {
        /* the read itself failed on the GPU (chain ranking or partition limits) before any task existed: say so, count it */
        if ((*n_failed)++ < 20) fprintf(stderr, "[th_host] read %s: failed on the GPU before the consensus stage (code %d); its records are missing\n", name, R->read_status[r]);
    }
    for (t = R->read_task_off[r]; t < R->read_task_off[r + 1]; ++t) {
        const int p0 = R->task_pos_off[t], pos_n = R->task_pos_off[t + 1] - p0;
        const int32_t *pos = R->pos + p0;
        if (p->gpu.only_unit) { /* write_tandem_unit */
            if (n_rec == m_rec) { m_rec = m_rec ? m_rec * 2 : 4; rec = (rec_t *)realloc(rec, sizeof(rec_t) * m_rec); }
            memset(&rec[n_rec], 0, sizeof(rec_t)); rec[n_rec].pos_n = pos_n; rec[n_rec].sub_pos = pos; ++n_rec;
            continue;
        }
        if (R->task_status[t] != 0) { /* e.g. TH_ERR_LEN: a unit beyond the int16 score range (the reference switches to int32 there) */
            if ((*n_failed)++ < 20) fprintf(stderr, "[th_host] read %s: consensus task failed on the GPU (code %d); record dropped\n", name, R->task_status[t]);
            continue;
        }
        {
            const int c0 = R->task_cons_off[t]; int cons_len = R->task_cons_off[t + 1] - c0;
            const int n_seqs = R->task_n_seqs[t];
            const uint8_t *cb = R->cons_base + c0; const int32_t *cov = R->cons_cov + c0;
            char *cons_seq; uint8_t *cons_qual = NULL;
            double ave_match = 0, copy_num; int cons_start, cons_end, full_length = 0, skip = 0, min_cov = 0;
            if (cons_len <= 0) continue; /* the reference would spin here (src/gen_cons.c:206) */
            /* min-cov filter of abpoa_gen_cons (src/abpoa_cons.c:52-98) */
            if (p->min_frac > 0.0) min_cov = (int)(n_seqs * p->min_frac); else if (p->min_cov > 0) min_cov = p->min_cov;
            if (min_cov > 0) {
                if (n_seqs <= 2) {
                    int _min_cov = 2, l0 = pos[1] - pos[0], l1 = pos[2] - pos[1];
                    if (l0 != l1) _min_cov = 1; else for (i = 0; i < l0; ++i) if (nt4_class[(unsigned char)seq[pos[0] + 1 + i]] != nt4_class[(unsigned char)seq[pos[1] + 1 + i]]) { _min_cov = 1; break; }
                    if (_min_cov < min_cov) skip = 1;
                } else for (i = 0; i < cons_len; ++i) if (cov[i] < min_cov) { skip = 1; break; }
            }
            if (skip) continue; /* reference: cons_len = 0 -> spins; we drop the record */
            cons_seq = (char *)malloc((size_t)cons_len * 2 + 2);
            for (i = 0; i < cons_len; ++i) cons_seq[i] = "ACGTN"[cb[i] > 4 ? 4 : cb[i]];
            cons_seq[cons_len] = 0;
            if (with_qual) { /* phred from coverage, src/abpoa_cons.c:100-107; n_seqs <= 2 -> '!' */
                cons_qual = (uint8_t *)malloc((size_t)cons_len * 2 + 2);
                for (i = 0; i < cons_len; ++i) {
                    if (n_seqs <= 2) cons_qual[i] = 33;
                    else {
                        double x = 13.8 * (1.25 * cov[i] / n_seqs - 0.25);
                        double pr = 1 - 1.0 / (1.0 + pow(2.718281828459045, -1 * x));
                        cons_qual[i] = (uint8_t)(33 + (int)(-10 * log10(pr) + 0.499));
                    }
                }
            }
            for (i = 0; i < pos_n - 1; ++i) { /* src/gen_cons.c:208-214 */
                int ulen = pos[i + 1] - pos[i], iden_n = R->iden_n[p0 + i];
                ave_match += (iden_n * 100 / (ulen + 0.0));
            }
            copy_num = n_seqs;
            cons_start = pos[0] - R->ext[4 * t + 1];
            copy_num += (R->ext[4 * t + 0] + 1.0) / cons_len;
            cons_end = pos[pos_n - 1] + R->ext[4 * t + 3] + 1;
            copy_num += (R->ext[4 * t + 2] + 1.0) / cons_len;
            if (p->five_seq && p->three_seq && cons_len > h->five_len + h->three_len) /* src/gen_cons.c:224-291 */
                full_length = trim_to_adapters(h, cons_seq, cons_qual, &cons_len);
            if (!p->only_full_length || full_length > 0) { /* write_tandem_cons_seq, src/gen_cons.c:10-62 */
                int keep = !(cons_len < p->min_len || cons_len > p->gpu.max_p);
                if (keep && p->only_longest && n_rec == 1) {
                    if (cons_end - cons_start > rec[0].cons_end - rec[0].cons_start) { n_rec = 0; cons_txt.l = 0; }
                    else keep = 0;
                }
                if (keep) {
                    if (n_rec == m_rec) { m_rec = m_rec ? m_rec * 2 : 4; rec = (rec_t *)realloc(rec, sizeof(rec_t) * m_rec); }
                    rec[n_rec].cons_start = cons_start; rec[n_rec].cons_end = cons_end; rec[n_rec].cons_len = cons_len;
                    rec[n_rec].full_length = full_length; rec[n_rec].pos_n = pos_n; rec[n_rec].sub_pos = pos;
                    rec[n_rec].copy_num = copy_num; rec[n_rec].ave_match = ave_match / (pos_n - 1);
                    rec[n_rec].seq_off = cons_txt.l;
                    str_write(&cons_txt, cons_seq, cons_len);
                    if (cons_qual) str_write(qs, (const char *)cons_qual, cons_len); /* appended at qual.l, never rewound */
                    ++n_rec;
                }
            }
            free(cons_seq); free(cons_qual);
        }
    }
    /* single_copy_full_len_seq, src/gen_cons.c:128-171 (tidehunter_core runs it after the chains, src/tidehunter.c:49-51) */
    if (p->single_copy == 1 && p->only_full_length && p->five_seq && p->three_seq && len >= p->gpu.k) {
        ed_res_t e5[3], e3[3]; int n5, n3, tar_start = -1, tar_end = -1, tot_ed, full_length = 0, cons_len = 0;
        n5 = collect_ed_res(p->ada_match_rat, p->five_seq, h->five_len, seq, len, e5);
        n3 = collect_ed_res(p->ada_match_rat, h->three_rc, h->three_len, seq, len, e3);
        tot_ed = full_len_pair(p->min_len, n5, e5, n3, e3, &tar_start, &tar_end);
        if (tot_ed != INT32_MAX) { sc_pos[0] = tar_start; sc_pos[1] = tar_end; cons_len = tar_end - tar_start + 1; full_length = 1; }
        if (tot_ed > 0) { /* reverse strand */
            n5 = collect_ed_res(p->ada_match_rat, h->five_rc, h->five_len, seq, len, e5);
            n3 = collect_ed_res(p->ada_match_rat, p->three_seq, h->three_len, seq, len, e3);
            if (full_len_pair(p->min_len, n3, e3, n5, e5, &tar_start, &tar_end) < tot_ed) {
                sc_pos[0] = tar_start; sc_pos[1] = tar_end; cons_len = tar_end - tar_start + 1; full_length = 2;
            }
        }
        if (full_length > 0) {
            int keep = 1;
            if (!p->gpu.only_unit) {
                keep = !(cons_len < p->min_len || cons_len > p->gpu.max_p);
                if (keep && p->only_longest && n_rec == 1) {
                    if (sc_pos[1] - sc_pos[0] > rec[0].cons_end - rec[0].cons_start) { n_rec = 0; cons_txt.l = 0; }
                    else keep = 0;
                }
            }
            if (keep) {
                if (n_rec == m_rec) { m_rec = m_rec ? m_rec * 2 : 4; rec = (rec_t *)realloc(rec, sizeof(rec_t) * m_rec); }
                memset(&rec[n_rec], 0, sizeof(rec_t));
                rec[n_rec].pos_n = 2; rec[n_rec].sub_pos = sc_pos;
                if (!p->gpu.only_unit) {
                    rec[n_rec].cons_start = sc_pos[0]; rec[n_rec].cons_end = sc_pos[1]; rec[n_rec].cons_len = cons_len;
                    rec[n_rec].full_length = full_length; rec[n_rec].copy_num = 1.0; rec[n_rec].ave_match = 100.0;
                    rec[n_rec].seq_off = cons_txt.l;
                    str_write(&cons_txt, seq + sc_pos[0], cons_len);   /* the read's own characters, case kept */
                    if (with_qual) { str_reserve(qs, cons_len); memset(qs->s + qs->l, 33, cons_len); qs->l += cons_len; qs->s[qs->l] = 0; }
                }
                ++n_rec;
            }
        }
    }
    /* mini_tandem_output, src/main.c:214-271 */
    {
        str_t *o = out; int ci, j; size_t qoff = 0;
        for (ci = 0; ci < n_rec; ++ci) {
            const rec_t *c = rec + ci;
            if (p->gpu.only_unit) {
                if (p->out_fmt == 1) {
                    for (i = 0; i < c->pos_n - 1; ++i) {
                        str_printf(o, ">%s_rep%d_sub%d\n", name, ci, i);
                        if (c->sub_pos[i + 1] > c->sub_pos[i]) str_write(o, seq + c->sub_pos[i] + 1, c->sub_pos[i + 1] - c->sub_pos[i]);
                        str_write(o, "\n", 1);
                    }
                } else if (p->out_fmt == 2) { /* the tabular unit output drops the last base (`<` vs `<=`) */
                    for (i = 0; i < c->pos_n - 1; ++i) {
                        str_printf(o, "%s\trep%d\tsub%d\t", name, ci, i);
                        if (c->sub_pos[i + 1] - 1 > c->sub_pos[i]) str_write(o, seq + c->sub_pos[i] + 1, c->sub_pos[i + 1] - 1 - c->sub_pos[i]);
                        str_write(o, "\n", 1);
                    }
                }
                continue;
            }
            if (p->out_fmt == 1 || p->out_fmt == 3)
                str_printf(o, "%c%s_rep%d_%.1f %d_%d_%d_%d_%.1f_%d_", p->out_fmt == 1 ? '>' : '@', name, ci, c->copy_num, len, c->cons_start + 1, c->cons_end + 1, c->cons_len, c->ave_match, c->full_length);
            else
                str_printf(o, "%s\trep%d\t%.1f\t%d\t%d\t%d\t%d\t%.1f\t%d\t", name, ci, c->copy_num, len, c->cons_start + 1, c->cons_end + 1, c->cons_len, c->ave_match, c->full_length);
            str_printf(o, "%d", c->sub_pos[0] + 2);
            for (j = 1; j < c->pos_n - 1; ++j) str_printf(o, ",%d", c->sub_pos[j] + 2);
            str_printf(o, ",%d%c", c->sub_pos[j] + 1, (p->out_fmt == 1 || p->out_fmt == 3) ? '\n' : '\t');
            str_write(o, cons_txt.s + c->seq_off, c->cons_len);
            if (p->out_fmt == 3) str_write(o, "\n+\n", 3); else if (p->out_fmt == 4) str_write(o, "\t", 1);
            if (with_qual) { /* printed from offset 0 of the slot's buffer, whatever it holds (main.c:258-262) */
                str_write(o, qs->s + qoff, c->cons_len);
                qoff += c->cons_len;
            }
            str_write(o, "\n", 1);
        }
    }
    free(rec); free(cons_txt.s);
}

static void add_stats(th_gpu_stats *a, const th_gpu_stats *b) {
    a->ms_h2d += b->ms_h2d; a->ms_pack += b->ms_pack; a->ms_seed += b->ms_seed; a->ms_chain += b->ms_chain; a->ms_select += b->ms_select;
    a->ms_partition += b->ms_partition; a->ms_poa += b->ms_poa; a->ms_ksw += b->ms_ksw; a->ms_d2h += b->ms_d2h; a->ms_total += b->ms_total;
    a->n_bases += b->n_bases; a->n_hits += b->n_hits; a->n_chain_evals += b->n_chain_evals; a->n_poa_cells += b->n_poa_cells; a->n_poa_rows += b->n_poa_rows;
    a->n_ksw_cells += b->n_ksw_cells; a->n_tasks += b->n_tasks; a->n_launches += b->n_launches; a->h2d_bytes += b->h2d_bytes; a->d2h_bytes += b->d2h_bytes; a->n_ksw_cells_full += b->n_ksw_cells_full;
}

/* One th_host_run in flight.  Chunks are cut by work, not only by count: a chunk ends after chunk_reads reads or once it
 * holds TH_CHUNK_BASES_PER_READ x chunk_reads bases (the reference balances reads over its threads dynamically,
 * src/main.c:273-291; with mixed read lengths equal read counts are unequal work).  Lanes TAKE chunks -- a free lane takes
 * the first chunk nobody has yet -- so a lane that drew long reads does not hold up the others.  The caller's thread formats
 * the chunks strictly in input order (the FASTQ slot quirk and the output order are sequential).  A lane copies its chunk's
 * result out of the context's buffers (copy_result) and takes its next chunk at once; it only waits when the formatter is
 * more than max_ahead chunks behind.  Every chunk before a taken one is taken too, so the formatter never waits on an
 * untaken chunk. */
#define TH_CHUNK_BASES_PER_READ 12288
static int *cut_chunks(const th_host *h, int n, const int32_t *lens, int *n_chunks) {
    const long long cap_b = (long long)h->p.chunk_reads * TH_CHUNK_BASES_PER_READ;
    int *start = (int *)malloc(sizeof(int) * ((size_t)n + 2)), nc = 0, r = 0;
    while (r < n) {
        long long b = 0; int m = 0;
        start[nc++] = r;
        while (r < n && m < h->p.chunk_reads && (m == 0 || b + lens[r] <= cap_b)) { b += lens[r]; ++r; ++m; }
    }
    start[nc] = n;
    *n_chunks = nc;
    return start;
}
enum { CH_PENDING = 0, CH_READY = 1, CH_EMITTED = 2, CH_FAILED = 3 };
typedef struct {
    th_host *h; int n, n_chunks; const char *const *seqs; const int32_t *lens;
    const int *start;      /* chunk c = reads [start[c], start[c + 1]) */
    int next;              /* first chunk no lane has taken yet */
    int n_emitted;         /* chunks formatted so far (they are formatted in order) */
    int max_ahead;         /* a lane takes chunk c only while c - n_emitted < max_ahead: bounds the copies waiting to be formatted */
    pthread_mutex_t mu; pthread_cond_t cv;
    int *state; th_gpu_result *res; void **own; int abort; char err[512];
} run_job;
typedef struct { run_job *job; int lane; double t_gpu, t_wait; } lane_arg;
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* A chunk's result lives in buffers of the GPU context and is overwritten by the context's next chunk.  The lane copies
 * it out (a few MB) so that it can start its next chunk at once instead of waiting until the formatter, which works in
 * input order, has got to this one.  One allocation holds all arrays; *own is what the formatter frees. */
static int copy_result(const th_gpu_result *s, th_gpu_result *d, void **own) {
    const size_t n = (size_t)s->n_reads, nt = (size_t)s->n_tasks;
    const size_t n_pos = nt && s->task_pos_off ? (size_t)s->task_pos_off[nt] : 0, n_cons = nt && s->task_cons_off ? (size_t)s->task_cons_off[nt] : 0;
    const int cov_is_alias = (const void *)s->cons_cov == (const void *)s->iden_n;
    const size_t sz[11] = { 4 * (n + 1), 4 * (nt + 1), 4 * n_pos, 4 * nt, 4 * (nt + 1), n_cons, cov_is_alias ? 0 : 4 * n_cons, 4 * n_pos, 16 * nt, 4 * nt, 4 * n };
    const void *src[11] = { s->read_task_off, s->task_pos_off, s->pos, s->task_n_seqs, s->task_cons_off, s->cons_base, s->cons_cov, s->iden_n, s->ext, s->task_status, s->read_status };
    void *dst[11]; size_t tot = 256, off = 0; int i; char *blk;
    for (i = 0; i < 11; ++i) tot += (sz[i] + 15) & ~(size_t)15;
    blk = (char *)malloc(tot);
    if (!blk) return -1;
    for (i = 0; i < 11; ++i) {
        dst[i] = NULL;
        if (src[i] && sz[i]) { dst[i] = blk + off; memcpy(dst[i], src[i], sz[i]); off += (sz[i] + 15) & ~(size_t)15; }
        else if (src[i] && i != 6) { dst[i] = blk + off; off += 16; }   /* present but empty: keep a valid pointer */
    }
    *d = *s;
    d->read_task_off = (const int32_t *)dst[0]; d->task_pos_off = (const int32_t *)dst[1]; d->pos = (const int32_t *)dst[2];
    d->task_n_seqs = (const int32_t *)dst[3]; d->task_cons_off = (const int32_t *)dst[4]; d->cons_base = (const uint8_t *)dst[5];
    d->iden_n = (const int32_t *)dst[7]; d->cons_cov = cov_is_alias ? d->iden_n : (const int32_t *)dst[6];
    d->ext = (const int32_t *)dst[8]; d->task_status = (const int32_t *)dst[9]; d->read_status = (const int32_t *)dst[10];
    *own = blk;
    return 0;
}

static void *lane_main(void *arg_) {
    lane_arg *a = (lane_arg *)arg_; run_job *J = a->job; th_host *h = J->h;
    for (;;) {
        int c, c0, m, rc, stop;
        th_gpu_result R;
        { const double t0 = now_s();
          pthread_mutex_lock(&J->mu);
          while (!J->abort && J->next < J->n_chunks && J->next - J->n_emitted >= J->max_ahead) pthread_cond_wait(&J->cv, &J->mu);
          stop = J->abort; c = J->next; if (!stop && c < J->n_chunks) J->next = c + 1;
          pthread_mutex_unlock(&J->mu);
          a->t_wait += now_s() - t0; }
        if (stop || c >= J->n_chunks) break;
        c0 = J->start[c]; m = J->start[c + 1] - c0;
        { const double t0 = now_s();
          rc = th_gpu_process_chunk(h->lane[a->lane], m, J->seqs + c0, J->lens + c0, &R);
          if (rc) { pthread_mutex_lock(&J->mu); if (!J->err[0]) snprintf(J->err, sizeof(J->err), "%s", th_gpu_last_error()); pthread_mutex_unlock(&J->mu); }
          else if (copy_result(&R, &J->res[c], &J->own[c])) { rc = -1; pthread_mutex_lock(&J->mu); if (!J->err[0]) snprintf(J->err, sizeof(J->err), "out of memory copying a chunk's result"); pthread_mutex_unlock(&J->mu); }
          a->t_gpu += now_s() - t0; }
        pthread_mutex_lock(&J->mu);
        if (rc) { J->state[c] = CH_FAILED; J->abort = 1; } else J->state[c] = CH_READY;
        pthread_cond_broadcast(&J->cv);
        stop = J->abort;
        pthread_mutex_unlock(&J->mu);
        if (stop) break;
    }
    return NULL;
}

/* Formats the m reads of one finished chunk into h->out, in input order, on several threads: a read's records depend
 * only on that read (and on its own quality slot, src/main.c:266-267 -- distinct for reads less than TH_SLOTS apart), so
 * blocks of at most TH_SLOTS reads are split into contiguous segments with a buffer each.  The adapter searches of the
 * full-length / single-copy modes (src/gen_cons.c:85-171) are the heavy part; the reference runs them on its thread pool. */
typedef struct { th_host *h; const th_gpu_result *R; const char *const *names, *const *seqs; const int32_t *lens; int r0, r1; int64_t g0; str_t out; long long failed; } fmt_arg;
static void *fmt_main(void *a_) {
    fmt_arg *a = (fmt_arg *)a_; int r;
    for (r = a->r0; r < a->r1; ++r) emit_read(a->h, &a->out, &a->failed, a->R, r, a->names[r], a->seqs[r], a->lens[r], a->g0 + r);
    return NULL;
}
static int fmt_threads(void) {
    static int n = 0;
    if (n == 0) {
        const char *e = getenv("TH_HOST_FMT_THREADS"), *lw = getenv("LOCAL_WORLD_SIZE"); /* set by torchrun: processes sharing this host's cores */
        long c = sysconf(_SC_NPROCESSORS_ONLN);
        if (lw && atoi(lw) > 1) c /= atoi(lw);           /* one process per GPU: each takes its share of the cores, not all of them */
        n = e ? atoi(e) : (int)(c >= 16 ? 8 : c >= 4 ? c / 2 : c >= 2 ? 2 : 1);
        if (lw && atoi(lw) > 1 && !e && c >= 2 && c < 16) n = (int)c; /* lane threads sleep in blocking waits: formatting may use the whole share */
        if (n < 1) n = 1; if (n > 32) n = 32;
    }
    return n;
}
static void emit_chunk(th_host *h, const th_gpu_result *R, int m, const char *const *names, const char *const *seqs, const int32_t *lens) {
    int b0;
    for (b0 = 0; b0 < m; b0 += TH_SLOTS) {
        const int b1 = b0 + TH_SLOTS < m ? b0 + TH_SLOTS : m, nb = b1 - b0;
        int T = fmt_threads(), t;
        { const char *e = getenv("TH_HOST_FMT_GRAIN"); const int grain = e && atoi(e) > 0 ? atoi(e) : 64; /* reads per thread worth a thread */
          if (nb < grain * T) T = nb / grain > 0 ? nb / grain : 1; }
        if (T <= 1) { int r; for (r = b0; r < b1; ++r) emit_read(h, &h->out, &h->n_failed, R, r, names[r], seqs[r], lens[r], h->read_counter + r); continue; }
        {
            fmt_arg *fa = (fmt_arg *)calloc((size_t)T, sizeof(fmt_arg)); pthread_t *th = (pthread_t *)calloc((size_t)T, sizeof(pthread_t));
            for (t = 0; t < T; ++t) {
                fa[t].h = h; fa[t].R = R; fa[t].names = names; fa[t].seqs = seqs; fa[t].lens = lens; fa[t].g0 = h->read_counter;
                fa[t].r0 = b0 + (int)((long long)nb * t / T); fa[t].r1 = b0 + (int)((long long)nb * (t + 1) / T);
                pthread_create(&th[t], NULL, fmt_main, &fa[t]);
            }
            for (t = 0; t < T; ++t) {
                pthread_join(th[t], NULL);
                if (fa[t].out.l) str_write(&h->out, fa[t].out.s, fa[t].out.l);
                h->n_failed += fa[t].failed; free(fa[t].out.s);
            }
            free(fa); free(th);
        }
    }
}

const char *th_host_run(th_host *h, int n, const char *const *names, const char *const *seqs, const int32_t *lens, size_t *out_len) {
    int c, n_chunks = 0;
    int *start = cut_chunks(h, n, lens, &n_chunks);
    h->out.l = 0; str_reserve(&h->out, 16); h->out.s[0] = 0;
    memset(&h->stats, 0, sizeof(h->stats));
    if (n_chunks <= 1 || h->n_lanes == 1) {
        for (c = 0; c < n_chunks; ++c) {
            const int c0 = start[c], m = start[c + 1] - c0;
            th_gpu_result R;
            if (th_gpu_process_chunk(h->gpu, m, seqs + c0, lens + c0, &R)) { set_err("%s", th_gpu_last_error()); free(start); *out_len = 0; return NULL; }
            emit_chunk(h, &R, m, names + c0, seqs + c0, lens + c0);
            h->read_counter += m;
            add_stats(&h->stats, &R.stats);
        }
    } else {
        run_job J; pthread_t th[TH_MAX_LANES]; lane_arg la[TH_MAX_LANES]; int n_thr = h->n_lanes < n_chunks ? h->n_lanes : n_chunks, failed = 0;
        memset(&J, 0, sizeof(J));
        J.h = h; J.n = n; J.n_chunks = n_chunks; J.seqs = seqs; J.lens = lens; J.start = start; J.next = 0;
        J.state = (int *)calloc(n_chunks, sizeof(int)); J.res = (th_gpu_result *)calloc(n_chunks, sizeof(th_gpu_result)); J.own = (void **)calloc(n_chunks, sizeof(void *));
        J.n_emitted = 0; J.max_ahead = 2 * n_thr + 2;
        pthread_mutex_init(&J.mu, NULL); pthread_cond_init(&J.cv, NULL);
        const double t_run0 = now_s(); double t_emit = 0, t_mwait = 0;
        for (c = 0; c < n_thr; ++c) { la[c].job = &J; la[c].lane = c; la[c].t_gpu = la[c].t_wait = 0; pthread_create(&th[c], NULL, lane_main, &la[c]); }
        for (c = 0; c < n_chunks && !failed; ++c) {
            const int c0 = start[c], m = start[c + 1] - c0;
            double t0 = now_s();
            pthread_mutex_lock(&J.mu);
            while (J.state[c] == CH_PENDING && !J.abort) pthread_cond_wait(&J.cv, &J.mu);
            failed = J.state[c] != CH_READY;
            pthread_mutex_unlock(&J.mu);
            t_mwait += now_s() - t0;
            if (failed) break;
            t0 = now_s();
            emit_chunk(h, &J.res[c], m, names + c0, seqs + c0, lens + c0);
            t_emit += now_s() - t0;
            h->read_counter += m;
            add_stats(&h->stats, &J.res[c].stats);
            free(J.own[c]); J.own[c] = NULL;
            pthread_mutex_lock(&J.mu); J.state[c] = CH_EMITTED; J.n_emitted = c + 1; pthread_cond_broadcast(&J.cv); pthread_mutex_unlock(&J.mu);
        }
        if (failed) { pthread_mutex_lock(&J.mu); J.abort = 1; pthread_cond_broadcast(&J.cv); pthread_mutex_unlock(&J.mu); }
        for (c = 0; c < n_thr; ++c) pthread_join(th[c], NULL);
        if (getenv("TH_HOST_TIMING")) {
            fprintf(stderr, "[th_host_run] %d reads, %d chunks, %d lanes: %.1f ms; formatting %.1f ms, waiting for chunks %.1f ms;", n, n_chunks, n_thr, 1e3 * (now_s() - t_run0), 1e3 * t_emit, 1e3 * t_mwait);
            for (c = 0; c < n_thr; ++c) fprintf(stderr, " lane%d gpu %.1f wait %.1f;", c, 1e3 * la[c].t_gpu, 1e3 * la[c].t_wait);
            fprintf(stderr, "\n");
        }
        pthread_mutex_destroy(&J.mu); pthread_cond_destroy(&J.cv);
        for (c = 0; c < n_chunks; ++c) free(J.own[c]);
        free(J.own);
        if (failed) { set_err("%s", J.err[0] ? J.err : "a GPU lane failed"); free(J.state); free(J.res); free(start); *out_len = 0; return NULL; }
        free(J.state); free(J.res);
    }
    free(start);
    *out_len = h->out.l;
    return h->out.s;
}
