/*
 * th_oracle.c -- CPU restatement of TideHunter v1.5.5's per-read pipeline (seeding, chaining,
 * partition, consensus driver, result formatting).  TEST INFRASTRUCTURE ONLY (see th_oracle.h).
 * ksw2 lives in th_ksw.c, abPOA in th_poa.c.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include "th_oracle.h"

#define MIN2(a, b) ((a) < (b) ? (a) : (b))
#define MAX2(a, b) ((a) > (b) ? (a) : (b))

/* src/main.c:325-362, src/tidehunter.h:9-41 */
void tho_default_para(tho_para_t *p) {
    memset(p, 0, sizeof(*p));
    p->k = 8; p->w = 1; p->hpc = 0;
    p->min_copy = 2; p->max_div = 0.25; p->min_p = 30; p->max_p = 10000;
    p->min_cov = 0; p->min_frac = 0.0;
    p->match = 2; p->mismatch = 4; p->gap_open1 = 4; p->gap_open2 = 24; p->gap_ext1 = 2; p->gap_ext2 = 1;
    p->out_fmt = 1; p->min_len = 30;
    p->ada_match_rat = 0.8f;
    p->pn16 = 16;
}

/* ------------------------------------------------------------------ nt4 (src/seq.c:15-32, 77-87) */
static uint8_t nt4(unsigned char c) {
    switch (c) {
    case 0: case 'A': case 'a': return 0;
    case 1: case 'C': case 'c': return 1;
    case 2: case 'G': case 'g': return 2;
    case 3: case 'T': case 't': return 3;
    case '-': return 5;
    default: return 4;
    }
}
void tho_get_bseq(const char *seq, int len, uint8_t *bseq) {
    int i;
    for (i = 0; i < len; ++i) bseq[i] = nt4((unsigned char)seq[i]);
}

/* ------------------------------------------------------------------ seeding (src/tandem_hit.c) */
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* tandem_hit.c:37-56 : every k-mer, keyed by its 2-bit code, positioned at its last base */
static int seeds_direct(const uint8_t *bseq, int len, int k, int hpc, uint64_t *h) {
    uint32_t key = 0, mask = (uint32_t)(((uint64_t)1 << 2 * k) - 1);
    int l = 0, n = 0, pos;
    for (pos = 0; pos < len; ++pos) {
        int c = bseq[pos];
        if (c >= 4) { key = 0; l = 0; continue; }
        if (hpc) while (pos + 1 < len && bseq[pos + 1] == c) ++pos;
        key = key << 2 | (uint32_t)c;
        if (++l >= k) { key &= mask; h[n++] = (uint64_t)key << 32 | (uint32_t)pos; }
    }
    return n;
}

/* tandem_hit.c:97-157 : (w,k) window minimizers ordered by the raw key */
typedef struct { uint32_t x, y; } mm_t;
static int seeds_minimizer(const uint8_t *bseq, int len, int k, int w, int hpc, uint64_t *h) {
    int i, j, l = 0, n = 0, span = 0, bp = 0, minp = 0;
    uint32_t key = 0, mask = (uint32_t)((1ULL << 2 * k) - 1);
    mm_t buf[256], mn = {UINT32_MAX, UINT32_MAX};
    int tq[32], tq_front = 0, tq_count = 0;
#define EMIT(e) (h[n++] = (uint64_t)(e).x << 32 | (e).y)
    for (i = 0; i < len; ++i) {
        int c = bseq[i];
        mm_t info = {UINT32_MAX, UINT32_MAX};
        if (c < 4) {
            if (hpc) {
                int skip = 1;
                if (i + 1 < len && bseq[i + 1] == c) {
                    for (skip = 2; i + skip < len; ++skip) if (bseq[i + skip] != c) break;
                    i += skip - 1;
                }
                tq[(tq_count++ + tq_front) & 0x1f] = skip;
                span += skip;
                if (tq_count > k) { span -= tq[tq_front++]; tq_front &= 0x1f; --tq_count; }
            } else span = l + 1 < k ? l + 1 : k;
            key = (key << 2 | (uint32_t)c) & mask;
            ++l;
            if (l >= k && span < 256) { info.x = key; info.y = (uint32_t)i; }
        } else { l = 0; tq_count = tq_front = 0; span = 0; key = 0; }
        buf[bp] = info;
        if (l == w + k - 1 && mn.x != UINT32_MAX) {
            for (j = bp + 1; j < w; ++j) if (mn.x == buf[j].x && buf[j].y != mn.y) EMIT(buf[j]);
            for (j = 0; j < bp; ++j) if (mn.x == buf[j].x && buf[j].y != mn.y) EMIT(buf[j]);
        }
        if (info.x <= mn.x) {
            if (l >= w + k && mn.x != UINT32_MAX) EMIT(mn);
            mn = info; minp = bp;
        } else if (bp == minp) {
            if (l >= w + k - 1 && mn.x != UINT32_MAX) EMIT(mn);
            for (j = bp + 1, mn.x = UINT32_MAX; j < w; ++j) if (mn.x >= buf[j].x) { mn = buf[j]; minp = j; }
            for (j = 0; j <= bp; ++j) if (mn.x >= buf[j].x) { mn = buf[j]; minp = j; }
            if (l >= w + k - 1 && mn.x != UINT32_MAX) {
                for (j = bp + 1; j < w; ++j) if (mn.x == buf[j].x && mn.y != buf[j].y) EMIT(buf[j]);
                for (j = 0; j <= bp; ++j) if (mn.x == buf[j].x && mn.y != buf[j].y) EMIT(buf[j]);
            }
        }
        if (++bp == w) bp = 0;
    }
    if (mn.x != UINT32_MAX) EMIT(mn);
#undef EMIT
    return n;
}

/* tandem_hit.c:171-237 : sort seeds; for every occurrence keep its distance to the nearest earlier
 * occurrence of the same key that is >= min_p away, if <= max_p; sort hits by (end, period). */
int tho_collect_hits(const uint8_t *bseq, int len, const tho_para_t *p, uint64_t **hits_) {
    *hits_ = NULL;
    if (len - p->w <= 0) return 0;
    uint64_t *h = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(len + 1));
    int hn = p->w > 1 ? seeds_minimizer(bseq, len, p->k, p->w, p->hpc, h) : seeds_direct(bseq, len, p->k, p->hpc, h);
    if (hn == 0) { free(h); return 0; }
    qsort(h, hn, sizeof(uint64_t), cmp_u64);
    uint64_t *hits = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)hn);
    uint32_t min_p = (uint32_t)p->min_p, max_p = (uint32_t)p->max_p;
    int i, j, kk, n = 0, s = 0;
    for (i = 1; i <= hn; ++i) {
        if (i == hn || (h[i] >> 32) != (h[i - 1] >> 32)) {
            for (j = s + 1; j < i; ++j) {
                uint32_t d = 0;
                for (kk = j - 1; kk >= s; --kk) { d = (uint32_t)(h[j] - h[kk]); if (d >= min_p) break; }
                if (d >= min_p && d <= max_p) hits[n++] = (h[j] & 0xffffffffULL) << 32 | d;
            }
            s = i;
        }
    }
    free(h);
    qsort(hits, n, sizeof(uint64_t), cmp_u64);
    *hits_ = hits;
    return n;
}

/* ------------------------------------------------------------------ chaining (src/tandem_chain.c) */
static inline int ilog2_32(uint32_t v) { /* :14-19; ilog2(0) = -1 */
    int r = -1;
    while (v) { ++r; v >>= 1; }
    return r;
}
enum { NO_CON = 0, REG_CON = 1, SAME_CON = 2, OVL_CON = 3 };
/* :128-166 */
static inline int con_score(int cs, int ce, int ps, int pe, int k, int *score) {
    int cp = ce - cs, pp = pe - ps;
    if (cs <= ps || cp >= pp * 1.8 || pp >= cp * 1.8) return NO_CON;
    int de = abs(ce - pe), ds = abs(cs - ps), dpd = abs(cp - pp);
    int matched = MIN2(de, k) + MIN2(ds, k);
    int gap = dpd * dpd / 2 + ilog2_32((uint32_t)(de + ds)) / 2;
    *score = matched - gap;
    if (dpd == 0) return matched < 2 * k ? OVL_CON : SAME_CON;
    return REG_CON;
}

typedef struct { int *cell; int len, score; } wchain_t; /* cell = flat ids */
typedef struct { int id, score; } srank_t;

static void merge_sort_rank(srank_t *a, srank_t *tmp, int n) { /* stable, score descending (:21-43 + glibc merge sort) */
    if (n < 2) return;
    int h = n / 2, i = 0, j = h, k = 0;
    merge_sort_rank(a, tmp, h); merge_sort_rank(a + h, tmp, n - h);
    while (i < h && j < n) tmp[k++] = (a[j].score > a[i].score) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(srank_t) * n);
}

int tho_tandem_chain(const uint64_t *hits, int hit_n, const tho_para_t *p, tho_chain_t *out) {
    memset(out, 0, sizeof(*out));
    if (hit_n < 2) return 0;
    int k = p->k, i, j, tot_n = 0;
    int *st = (int *)malloc(sizeof(int) * hit_n), *en = (int *)malloc(sizeof(int) * hit_n);
    int *score = (int *)malloc(sizeof(int) * hit_n), *from = (int *)malloc(sizeof(int) * hit_n);
    int *row = (int *)malloc(sizeof(int) * hit_n);
    int *row_beg = (int *)malloc(sizeof(int) * (hit_n + 1));
    int8_t *tracked = (int8_t *)calloc(hit_n, 1);
    for (i = 0; i < hit_n; ++i) { /* rows = distinct ends; init_dp :113-126 */
        int end = (int)(hits[i] >> 32), period = (int)(uint32_t)hits[i];
        if (i == 0 || end != en[i - 1]) row_beg[tot_n++] = i;
        row[i] = tot_n - 1; en[i] = end; st[i] = end - period;
        score[i] = k + MIN2(k, period); from[i] = -1;
    }
    row_beg[tot_n] = hit_n;
    /* main DP :325-356 */
    int64_t n_evals = 0;
    int cur_i, cur, pre_i, pre;
    for (cur_i = 1; cur_i < tot_n; ++cur_i) {
        for (cur = row_beg[cur_i]; cur < row_beg[cur_i + 1]; ++cur) {
            int max_score = score[cur], max_pre = -1, max_h = en[cur] - st[cur], iter_n = 0, stop = 0;
            for (pre_i = cur_i - 1; pre_i >= 0 && !stop; --pre_i) {
                int gt = 0;
                if (en[row_beg[pre_i]] < st[cur]) break;
                for (pre = row_beg[pre_i]; pre < row_beg[pre_i + 1]; ++pre) {
                    int cs, res = con_score(st[cur], en[cur], st[pre], en[pre], k, &cs), s;
                    ++n_evals;
                    if (res == NO_CON) continue;
                    s = score[pre] + cs;
                    if (s > max_score) {
                        max_score = s; max_pre = pre;
                        if (res == SAME_CON || res == OVL_CON) { stop = 1; break; }
                        gt = 1;
                    } else if (res == OVL_CON) { stop = 1; break; }
                }
                if (stop) break;
                if (gt) iter_n = 0;
                else if (++iter_n >= max_h) break;
            }
            if (max_score > score[cur]) { score[cur] = max_score; from[cur] = max_pre; }
        }
    }
    /* rank cells :21-43 : i from tot_n-1 down, j up, score > 0, stable sort by score descending */
    srank_t *rank = (srank_t *)malloc(sizeof(srank_t) * hit_n), *tmp = (srank_t *)malloc(sizeof(srank_t) * hit_n);
    int score_n = 0;
    for (i = tot_n - 1; i >= 0; --i)
        for (j = row_beg[i]; j < row_beg[i + 1]; ++j)
            if (score[j] > 0) { rank[score_n].id = j; rank[score_n++].score = score[j]; }
    merge_sort_rank(rank, tmp, score_n);
    free(tmp);
    /* greedy chain extraction :358-372 */
    int top_N = 1000, ch_n = 0;
    wchain_t *chain = (wchain_t *)calloc(top_N + 1, sizeof(wchain_t));
    int *chain_idx = (int *)malloc(sizeof(int) * top_N);
    for (i = 0; i < top_N; ++i) chain_idx[i] = i;
    int *cellbuf = (int *)malloc(sizeof(int) * tot_n);
    for (i = 0; i < score_n && ch_n < top_N; ++i) {
        int c = rank[i].id, _i, in_chain = 0;
        { /* is_in_chain :170-185 ; NB cell_start comes from the first cell of the row */
            int cell_start = st[row_beg[row[c]]], cell_end = en[c];
            for (_i = 0; _i < ch_n; ++_i) {
                wchain_t *ch = chain + chain_idx[_i];
                if (ch->len <= 0) continue;
                int chain_start = st[ch->cell[0]], chain_end = en[ch->cell[ch->len - 1]];
                if (chain_end < cell_start) break;
                else if (chain_start > cell_end) continue;
                else if (cell_end - chain_start >= (chain_end - chain_start) / 2) { in_chain = 1; break; }
            }
        }
        if (in_chain) continue;
        /* backtrack_dp :86-111 */
        int accepted = 0;
        if (!tracked[c]) {
            int sc = score[c], cur_c = c, len = 0;
            while (1) {
                tracked[cur_c] = 1; cellbuf[len++] = cur_c;
                int pr = from[cur_c];
                if (pr == -1) break;
                if (tracked[pr]) { sc -= score[pr]; break; }
                cur_c = pr;
            }
            wchain_t *ch = chain + ch_n;
            ch->cell = (int *)realloc(ch->cell, sizeof(int) * len);
            for (j = 0; j < len; ++j) ch->cell[j] = cellbuf[len - 1 - j];
            ch->len = len; ch->score = sc;
            if (len > 1) { /* is_overlap_chain :54-83 */
                int ovl = 0;
                if (ch_n > 0 && ch->len > 0) {
                    int start = st[ch->cell[ch->len - 1]];
                    for (j = ch_n - 1; j >= 0; --j) {
                        wchain_t *o = chain + j;
                        if (o->len <= 0) continue;
                        if (en[o->cell[o->len - 1]] <= start) break;
                        int s1 = st[o->cell[0]], e1 = st[o->cell[o->len - 1]];
                        int s2 = st[ch->cell[0]], e2 = st[ch->cell[ch->len - 1]];
                        int mn = MIN2(e1 - s1, e2 - s2), ovlp_len = MIN2(e1, e2) - MAX2(s1, s2);
                        if (ovlp_len / (mn + 0.0) >= 0.5) {
                            if (o->score > ch->score) ovl = 1;
                            else o->len = 0;
                            break;
                        }
                    }
                }
                if (!ovl) accepted = 1;
            }
        }
        if (accepted) ++ch_n;
        /* sort_chain :188-207 (selection-style, stale `i` kept as in the reference) */
        if (ch_n >= 2) {
            int _j;
            for (_i = 0; _i < ch_n - 1; ++_i) {
                int ii = chain_idx[_i];
                if (chain[ii].len <= 0) continue;
                int ch_end1 = en[chain[ii].cell[chain[ii].len - 1]];
                for (_j = _i + 1; _j < ch_n; ++_j) {
                    int jj = chain_idx[_j];
                    if (chain[jj].len <= 0) continue;
                    int ch_end2 = en[chain[jj].cell[chain[jj].len - 1]];
                    if (ch_end1 < ch_end2) { chain_idx[_i] = jj; chain_idx[_j] = ii; ch_end1 = ch_end2; }
                }
            }
        }
    }
    /* post-process :392-399 : ascending end, chains with >= 3 cells */
    out->n_cells = hit_n; out->start = st; out->end = en; out->score = score; out->from = from; out->row = row;
    out->n_evals = n_evals;
    out->chain_off = (int *)malloc(sizeof(int) * (ch_n + 2));
    out->est_start = (int *)malloc(sizeof(int) * (ch_n + 1)); out->est_period = (int *)malloc(sizeof(int) * (ch_n + 1));
    int tot = 0;
    for (i = 0; i < ch_n; ++i) if (chain[chain_idx[i]].len >= 3) tot += chain[chain_idx[i]].len;
    out->cells = (int *)malloc(sizeof(int) * (tot + 1));
    out->chain_off[0] = 0;
    for (i = ch_n - 1; i >= 0; --i) {
        wchain_t *ch = chain + chain_idx[i];
        if (ch->len - 1 < 2) continue; /* copy_chain :210 */
        memcpy(out->cells + out->chain_off[out->n_chain], ch->cell, sizeof(int) * ch->len);
        out->est_start[out->n_chain] = st[ch->cell[0]];
        out->est_period[out->n_chain] = en[ch->cell[0]] - st[ch->cell[0]];
        out->chain_off[out->n_chain + 1] = out->chain_off[out->n_chain] + ch->len;
        out->n_chain++;
    }
    for (i = 0; i <= top_N; ++i) free(chain[i].cell);
    free(chain); free(chain_idx); free(cellbuf); free(rank); free(tracked); free(row_beg);
    return out->n_chain;
}
void tho_chain_free(tho_chain_t *c) {
    free(c->start); free(c->end); free(c->score); free(c->from); free(c->row);
    free(c->chain_off); free(c->cells); free(c->est_start); free(c->est_period);
    memset(c, 0, sizeof(*c));
}

/* ------------------------------------------------------------------ partition (src/partition.c:171-276) */
int tho_partition(const uint8_t *bseq, int len, const tho_chain_t *c, int chain_i, const tho_para_t *p, int **par_pos_) {
    const int *cell = c->cells + c->chain_off[chain_i];
    int ch_len = c->chain_off[chain_i + 1] - c->chain_off[chain_i];
    int est_start = c->est_start[chain_i], est_period = c->est_period[chain_i];
    int last_start = c->start[cell[ch_len - 1]];
    int *par_pos = (int *)malloc(sizeof(int) * (size_t)(len + 4)), par_n = 0;
    int i, k = p->k, ch_i, s, e, s1, e1, s2, e2;
    /* left pass (:186-228) never iterates: est_ch_i is always 0 (tandem_chain.c:251-255) */
    par_pos[par_n++] = est_start;
    par_pos[par_n++] = est_start + est_period;
    ch_i = 0; s = est_start; e = est_start + est_period;
    while (ch_i < ch_len - 1 && e <= last_start) {
        s1 = s; e1 = e; s2 = e2 = -1;
        for (i = ch_i + 1; i < ch_len; ++i) {
            s2 = c->start[cell[i]]; e2 = c->end[cell[i]];
            if (s2 == e) {
                par_pos[par_n++] = e2;
                ch_i = i; s = s2; e = e2;
                break;
            } else if (s2 > e) {
                int n_cigar; uint32_t *cigar;
                int ql = s2 - s1 + k, tl = e2 - e1 + k;
                int iden_n = tho_ksw2_global(bseq + s1 - k + 1, ql, bseq + e1 - k + 1, tl, &n_cigar, &cigar);
                if (iden_n >= MIN2(ql, tl) * (1 - p->max_div)) {
                    s = e; e = e2 - tho_ksw2_backtrack_left_end(n_cigar, cigar, ql, tl, s2 - e);
                    if (e == s) { ch_i = ch_len; free(cigar); break; }
                    par_pos[par_n++] = e;
                    ch_i = i - 1;
                } else {
                    par_pos[par_n++] = -1; par_pos[par_n++] = s2; par_pos[par_n++] = e2;
                    ch_i = i; s = s2; e = e2;
                }
                free(cigar);
                break;
            } else { s1 = s2; e1 = e2; }
        }
        if (i == ch_len) break; /* the reference would spin here; unreachable because e <= last_start */
    }
    *par_pos_ = par_pos;
    return par_n;
}

/* ------------------------------------------------------------------ result records (src/gen_cons.c:10-83) */
typedef struct { /* the persistent part of tandem_seq_t that survives between chunks (qual.l quirk) */
    char *seq_s; size_t seq_l, seq_m;
    char *qual_s; size_t qual_l, qual_m;
} slot_t;

static tho_cons_t *push_cons(tho_read_t *r) {
    if (r->n_cons == r->m_cons) {
        r->m_cons = r->m_cons ? r->m_cons << 1 : 1;
        r->cons = (tho_cons_t *)realloc(r->cons, sizeof(tho_cons_t) * r->m_cons);
    }
    tho_cons_t *c = r->cons + r->n_cons++;
    memset(c, 0, sizeof(*c));
    return c;
}
static void free_cons(tho_cons_t *c) { free(c->sub_pos); free(c->cons_seq); free(c->cons_qual); free(c->iden_n); }

/* write_tandem_cons_seq :10-62 */
static void write_cons(tho_read_t *r, const tho_para_t *p, const char *cons_seq, const uint8_t *cons_qual, int cons_len,
                       int start, int end, double copy_num, double ave_match, int full_length, const int *par_pos, int pos_n,
                       const tho_cons_t *raw) {
    if (cons_len < p->min_len || cons_len > p->max_p) return;
    if (p->only_longest && r->n_cons == 1) {
        if (end - start > r->cons[0].cons_end - r->cons[0].cons_start) {
            if (r->cons[0].cons_qual) { /* the replaced record's quality bytes stay in the buffer (qual.l is not rewound) */
                r->dropped_qual = (char *)realloc(r->dropped_qual, r->dropped_l + r->cons[0].cons_len + 1);
                memcpy(r->dropped_qual + r->dropped_l, r->cons[0].cons_qual, r->cons[0].cons_len); r->dropped_l += r->cons[0].cons_len;
            }
            free_cons(r->cons); r->n_cons = 0;
        }
        else return;
    }
    tho_cons_t *c = push_cons(r);
    if (raw) { c->n_seqs = raw->n_seqs; c->raw_cons_len = raw->raw_cons_len; c->lext_q = raw->lext_q; c->lext_t = raw->lext_t; c->rext_q = raw->rext_q; c->rext_t = raw->rext_t;
        if (raw->iden_n) { c->iden_n = (int *)malloc(sizeof(int) * pos_n); memcpy(c->iden_n, raw->iden_n, sizeof(int) * (pos_n - 1)); } }
    c->cons_seq = (char *)malloc(cons_len + 1); memcpy(c->cons_seq, cons_seq, cons_len); c->cons_seq[cons_len] = 0;
    if (cons_qual) { c->cons_qual = (char *)malloc(cons_len + 1); memcpy(c->cons_qual, cons_qual, cons_len); c->cons_qual[cons_len] = 0; }
    c->cons_start = start; c->cons_end = end; c->copy_num = copy_num; c->full_length = full_length;
    c->cons_len = cons_len; c->ave_match = ave_match; c->pos_n = pos_n;
    c->sub_pos = (int *)malloc(sizeof(int) * pos_n); memcpy(c->sub_pos, par_pos, sizeof(int) * pos_n);
}
/* write_tandem_unit :64-83 */
static void write_unit(tho_read_t *r, const int *par_pos, int pos_n) {
    tho_cons_t *c = push_cons(r);
    c->pos_n = pos_n;
    c->sub_pos = (int *)malloc(sizeof(int) * pos_n); memcpy(c->sub_pos, par_pos, sizeof(int) * pos_n);
}

/* ------------------------------------------------------------------ adapters: infix edit distance
 * edlib_align_HW (src/edlib_align.c:73-85) = edlibAlign(mode HW, task LOC, threshold k), of which only
 * editDistance, endLocations[0] and startLocations[0] are used.  edlib (edlib/src/edlib.cpp:141-236)
 * reports: editDistance = min over target end positions of the infix edit distance (or -1 if > k);
 * endLocations = all end positions reaching it, ascending; startLocations[i] = endLocation minus the
 * LAST position of a prefix (SHW) search of reverse(query) in reverse(target[0..end]) bounded by
 * editDistance, i.e. the SMALLEST start with ed(query, target[start..end]) == editDistance.
 * Plain O(nm) dynamic programming restatement; equalities = case-insensitive A/C/G/T/N. */
static int edlib_hw(const char *q, int ql, const char *t, int tl, int *start, int *end, int k) {
    int i, j, best = -1, best_end = -1;
    if (ql <= 0 || tl <= 0) return -1;
    int *col = (int *)malloc(sizeof(int) * (ql + 1));
    for (i = 0; i <= ql; ++i) col[i] = i;
    for (j = 0; j < tl; ++j) {
        int diag = col[0];
        col[0] = 0; /* infix: free start anywhere in the target */
        for (i = 1; i <= ql; ++i) {
            int up = col[i - 1] + 1, left = col[i] + 1;
            int eq = (q[i - 1] | 0x20) == (t[j] | 0x20);
            int d = diag + (eq ? 0 : 1), v = d < up ? d : up;
            if (left < v) v = left;
            diag = col[i]; col[i] = v;
        }
        if (best < 0 || col[ql] < best) { best = col[ql]; best_end = j; }
    }
    if (best > ql) best = ql; /* never exceeds deleting the whole query (end location -1 case, unreachable for k < ql) */
    if (k >= 0 && best > k) { free(col); return -1; }
    {
        int best_start = -1;
        for (i = 0; i <= ql; ++i) col[i] = i;
        for (j = 0; j <= best_end; ++j) { /* prefix search on the reversed sequences */
            int diag = col[0];
            col[0] = j + 1;
            for (i = 1; i <= ql; ++i) {
                int up = col[i - 1] + 1, left = col[i] + 1;
                int eq = (q[ql - i] | 0x20) == (t[best_end - j] | 0x20);
                int d = diag + (eq ? 0 : 1), v = d < up ? d : up;
                if (left < v) v = left;
                diag = col[i]; col[i] = v;
            }
            if (col[ql] == best) best_start = best_end - j; /* keep the last one */
        }
        *start = best_start; *end = best_end;
    }
    free(col);
    return best;
}

static char *rc_seq(const char *s, int l) { /* src/seq.c:89-95 */
    char *r = (char *)malloc(l + 1); int i;
    for (i = 0; i < l; ++i) { uint8_t c = nt4((unsigned char)s[i]); r[l - i - 1] = "TGCAN"[c > 4 ? 4 : c]; }
    r[l] = 0;
    return r;
}

/* ------------------------------------------------------------------ consensus driver */
/* src/abpoa_cons.c:30-120 */
static int gen_cons(const tho_para_t *p, const uint8_t *bseq, int seq_len, const int *pos, int pos_n,
                    uint8_t *cons_bseq, uint8_t *cons_qual, int *n_seqs_, int64_t *poa_cells) {
    int i, n_seqs = 0, cons_len = 0;
    int *seq_lens = (int *)malloc(sizeof(int) * pos_n);
    const uint8_t **seqs = (const uint8_t **)malloc(sizeof(uint8_t *) * pos_n);
    for (i = 0; i < pos_n - 1; ++i) {
        int start = pos[i], end = pos[i + 1];
        if (start < 0 || end < 0 || start >= seq_len - 1 || end + 1 > seq_len) continue;
        seq_lens[n_seqs] = end - start; seqs[n_seqs] = bseq + start + 1; ++n_seqs;
    }
    *n_seqs_ = n_seqs;
    int min_cov = 0;
    if (p->min_frac > 0.0) min_cov = (int)(n_seqs * p->min_frac);
    else if (p->min_cov > 0) min_cov = p->min_cov;
    if (n_seqs <= 2) {
        if (n_seqs <= 1) { fprintf(stderr, "[tho] Not enough sequences to perform msa.\n"); exit(1); }
        int skip = 0;
        cons_len = seq_lens[0];
        if (min_cov > 0) {
            int _min_cov = 2;
            if (seq_lens[0] != seq_lens[1]) _min_cov = 1;
            else for (i = 0; i < cons_len; ++i) if (seqs[0][i] != seqs[1][i]) { _min_cov = 1; break; }
            if (_min_cov < min_cov) skip = 1;
        }
        if (!skip) for (i = 0; i < cons_len; ++i) { cons_bseq[i] = seqs[0][i]; if (cons_qual) cons_qual[i] = 33; }
        else cons_len = 0;
    } else {
        int *cov = (int *)malloc(sizeof(int) * (size_t)(seq_len + 2));
        int skip = 0;
        cons_len = tho_abpoa_cons(p, n_seqs, seqs, seq_lens, cons_bseq, cov, poa_cells);
        if (min_cov > 0) for (i = 0; i < cons_len; ++i) if (cov[i] < min_cov) { skip = 1; break; }
        if (cons_qual) {
            for (i = 0; i < cons_len; ++i) { /* :100-107 */
                double x = 13.8 * (1.25 * cov[i] / n_seqs - 0.25);
                double pr = 1 - 1.0 / (1.0 + pow(2.718281828459045, -1 * x));
                cons_qual[i] = (uint8_t)(33 + (int)(-10 * log10(pr) + 0.499));
            }
        }
        if (skip) cons_len = 0;
        free(cov);
    }
    free(seq_lens); free(seqs);
    return cons_len;
}

/* src/gen_cons.c:173-301 */
static void seqs_msa(int seq_len, const uint8_t *bseq, int par_n, const int *par_pos, tho_read_t *r, const tho_para_t *p,
                     const char *five_rc, const char *three_rc) {
    char *cons_seq = (char *)malloc(seq_len + 1);
    uint8_t *cons_bseq = (uint8_t *)malloc(seq_len + 2), *cons_qual = NULL;
    if (p->out_fmt == 3 || p->out_fmt == 4) cons_qual = (uint8_t *)malloc(seq_len + 2);
    int i = 0, j, k, s;
    int five_len = p->five_seq ? (int)strlen(p->five_seq) : 0, three_len = p->three_seq ? (int)strlen(p->three_seq) : 0;
    while (i < par_n - p->min_copy) {
        if (par_pos[i] < 0) { i++; continue; }
        for (j = i + 1; j < par_n; ++j) if (par_pos[j] < 0) break;
        if (j - i > p->min_copy) {
            if (p->only_unit) write_unit(r, par_pos + i, j - i);
            else {
                int n_seqs, cons_len;
                tho_cons_t raw; memset(&raw, 0, sizeof(raw));
                cons_len = gen_cons(p, bseq, seq_len, par_pos + i, j - i, cons_bseq, cons_qual, &n_seqs, &r->n_poa_cells);
                if (cons_len == 0) { fprintf(stderr, "[tho] cons_len == 0: the reference would spin here (gen_cons.c:206)\n"); exit(1); }
                double ave_match = 0;
                raw.n_seqs = n_seqs; raw.raw_cons_len = cons_len; raw.iden_n = (int *)malloc(sizeof(int) * (j - i));
                for (k = i; k < j - 1; ++k) {
                    int start = par_pos[k], end = par_pos[k + 1], len = end - start;
                    int iden_n = tho_ksw2_global(bseq + start + 1, len, cons_bseq, cons_len, NULL, NULL);
                    r->n_ksw_cells += (int64_t)len * cons_len;
                    raw.iden_n[k - i] = iden_n;
                    ave_match += (iden_n * 100 / (len + 0.0));
                }
                for (s = 0; s < cons_len; ++s) cons_seq[s] = "ACGTN"[cons_bseq[s]];
                cons_seq[cons_len] = 0;
                int max_q, max_t, cons_start, cons_end; double copy_num = n_seqs;
                tho_ksw2_left_ext(cons_bseq, cons_len, bseq, par_pos[i] + 1, &max_q, &max_t); cons_start = par_pos[i] - max_t;
                r->n_ksw_cells += (int64_t)cons_len * (par_pos[i] + 1);
                raw.lext_q = max_q; raw.lext_t = max_t;
                copy_num += (max_q + 1.0) / cons_len;
                tho_ksw2_ext(cons_bseq, cons_len, bseq + par_pos[j - 1] + 1, seq_len - par_pos[j - 1] - 1, &max_q, &max_t); cons_end = par_pos[j - 1] + max_t + 1;
                r->n_ksw_cells += (int64_t)cons_len * (seq_len - par_pos[j - 1] - 1);
                raw.rext_q = max_q; raw.rext_t = max_t;
                copy_num += (max_q + 1.0) / cons_len;
                int full_length = 0;
                if (p->five_seq && p->three_seq && cons_len > five_len + three_len) { /* :224-291 */
                    char *cons2 = (char *)malloc((cons_len << 1) + 1); uint8_t *qual2 = NULL;
                    memcpy(cons2, cons_seq, cons_len); memcpy(cons2 + cons_len, cons_seq, cons_len); cons2[cons_len << 1] = 0;
                    if (cons_qual) { qual2 = (uint8_t *)malloc(cons_len << 1); memcpy(qual2, cons_qual, cons_len); memcpy(qual2 + cons_len, cons_qual, cons_len); }
                    int tar_start = -1, tar_end = -1, tot_ed = INT32_MAX, _5_ed, _3_ed, _5_start = -1, _5_end = -1, _3_start = -1, _3_end = -1;
                    int k5 = (int)(five_len * (1 - p->ada_match_rat)), k3 = (int)(three_len * (1 - p->ada_match_rat));
                    _5_ed = edlib_hw(p->five_seq, five_len, cons2, cons_len << 1, &_5_start, &_5_end, k5);
                    if (_5_ed == -1) goto REV;
                    _3_ed = edlib_hw(three_rc, three_len, cons2, cons_len << 1, &_3_start, &_3_end, k3);
                    if (_3_ed == -1) goto REV;
                    if (_3_start <= _5_end) {
                        if (_3_end + cons_len < cons_len << 1 && _3_start + cons_len > _5_end) {
                            tar_start = _5_end + 1; tar_end = _3_start + cons_len - 1; full_length = 1; tot_ed = _5_ed + _3_ed;
                        }
                    } else { tar_start = _5_end + 1; tar_end = _3_start - 1; tot_ed = _5_ed + _3_ed; full_length = 1; }
                    if (tot_ed == 0) goto WRITE_CONS;
REV:
                    _5_ed = edlib_hw(five_rc, five_len, cons2, cons_len << 1, &_5_start, &_5_end, k5);
                    if (_5_ed == -1) goto WRITE_CONS;
                    _3_ed = edlib_hw(p->three_seq, three_len, cons2, cons_len << 1, &_3_start, &_3_end, k3);
                    if (_3_ed == -1) goto WRITE_CONS;
                    if (_5_ed + _3_ed < tot_ed) {
                        if (_5_start <= _3_end) {
                            if (_5_end + cons_len < cons_len << 1 && _5_start + cons_len > _3_end) { tar_start = _3_end + 1; tar_end = _5_start + cons_len - 1; full_length = 2; }
                        } else { tar_start = _3_end + 1; tar_end = _5_start - 1; full_length = 2; }
                    }
WRITE_CONS:
                    if (tar_start > 0 && tar_end > tar_start) {
                        memcpy(cons_seq, cons2 + tar_start, tar_end - tar_start + 1);
                        cons_seq[tar_end - tar_start + 1] = 0;
                        if (cons_qual) for (k = tar_start; k <= tar_end; ++k) cons_qual[k - tar_start] = qual2[k];
                        cons_len = tar_end - tar_start + 1;
                    }
                    free(cons2); free(qual2);
                }
                if (!p->only_full_length || full_length > 0)
                    write_cons(r, p, cons_seq, cons_qual, cons_len, cons_start, cons_end, copy_num, ave_match / (j - i - 1), full_length, par_pos + i, j - i, &raw);
                free(raw.iden_n);
            }
        }
        i = j + 1;
    }
    free(cons_seq); free(cons_bseq); free(cons_qual);
}

/* ------------------------------------------------------------------ -s: single-copy full-length reads
 * collect_ed_res src/gen_cons.c:89-110: best infix hit of the adapter, then one more on each side of it */
typedef struct { int ed, start, end; } ed_res_t;
static int collect_ed_res(const tho_para_t *p, const char *q, int qlen, const char *seq, int seq_len, ed_res_t *res) {
    int n = 0, ed, start = 0, end = 0, k = (int)(qlen * (1 - p->ada_match_rat));
    ed = edlib_hw(q, qlen, seq, seq_len, &start, &end, k);
    if (ed != -1) {
        res[0].ed = ed; res[0].start = start; res[0].end = end; n++;
        if (res[0].start >= qlen) {
            ed = edlib_hw(q, qlen, seq, res[0].start, &start, &end, k);
            if (ed != -1) { res[n].ed = ed; res[n].start = start; res[n].end = end; n++; }
        }
        if (res[0].end <= seq_len - qlen) {
            ed = edlib_hw(q, qlen, seq + res[0].end, seq_len - res[0].end, &start, &end, k);
            if (ed != -1) { res[n].ed = ed; res[n].start = res[0].end + start; res[n].end = res[0].end + end; n++; }
        }
    }
    return n;
}
/* get_full_len_seq src/gen_cons.c:112-126 */
static int get_full_len_seq(const tho_para_t *p, int left_n, const ed_res_t *left, int right_n, const ed_res_t *right, int *tar_start, int *tar_end) {
    int tot_ed = INT32_MAX, i, j;
    for (i = 0; i < left_n; ++i)
        for (j = 0; j < right_n; ++j)
            if (right[j].start - left[i].end - 1 >= p->min_len && tot_ed > left[i].ed + right[j].ed) {
                tot_ed = left[i].ed + right[j].ed;
                *tar_start = left[i].end + 1; *tar_end = right[j].start - 1;
            }
    return tot_ed;
}
/* single_copy_full_len_seq src/gen_cons.c:128-171 */
static void single_copy_full_len(int seq_len, const char *seq, tho_read_t *r, const tho_para_t *p, const char *five_rc, const char *three_rc) {
    int cons_len = 0, full_length = 0, tar_start = -1, tar_end = -1, tot_ed, _5_n, _3_n, par_pos[2];
    int five_len = (int)strlen(p->five_seq), three_len = (int)strlen(p->three_seq);
    ed_res_t _5[3], _3[3];
    _5_n = collect_ed_res(p, p->five_seq, five_len, seq, seq_len, _5);
    _3_n = collect_ed_res(p, three_rc, three_len, seq, seq_len, _3);
    tot_ed = get_full_len_seq(p, _5_n, _5, _3_n, _3, &tar_start, &tar_end);
    if (tot_ed != INT32_MAX) { par_pos[0] = tar_start; par_pos[1] = tar_end; cons_len = tar_end - tar_start + 1; full_length = 1; }
    if (tot_ed > 0) {
        _5_n = collect_ed_res(p, five_rc, five_len, seq, seq_len, _5);
        _3_n = collect_ed_res(p, p->three_seq, three_len, seq, seq_len, _3);
        if (get_full_len_seq(p, _3_n, _3, _5_n, _5, &tar_start, &tar_end) < tot_ed) {
            par_pos[0] = tar_start; par_pos[1] = tar_end; cons_len = tar_end - tar_start + 1; full_length = 2;
        }
    }
    if (full_length > 0) {
        if (p->only_unit) write_unit(r, par_pos, 2);
        else {
            uint8_t *cons_qual = NULL; int i;
            if (p->out_fmt == 3 || p->out_fmt == 4) { cons_qual = (uint8_t *)malloc(cons_len > 0 ? cons_len : 1); for (i = 0; i < cons_len; ++i) cons_qual[i] = 33; }
            write_cons(r, p, seq + par_pos[0], cons_qual, cons_len, par_pos[0], par_pos[1], 1.0, 100.0, full_length, par_pos, 2, NULL);
            free(cons_qual);
        }
    }
}

/* src/tidehunter.c:23-60 */
void tho_process_read(const char *seq, int len, const tho_para_t *p, tho_read_t *r) {
    if (len < p->k) return;
    uint8_t *bseq = (uint8_t *)malloc(len + 1);
    tho_get_bseq(seq, len, bseq);
    uint64_t *hits = NULL;
    int hit_n = tho_collect_hits(bseq, len, p, &hits);
    tho_chain_t ch;
    int ch_n = tho_tandem_chain(hits, hit_n, p, &ch), ci;
    r->n_hits += hit_n; r->n_chain_evals += ch.n_evals;
    free(hits);
    char *five_rc = p->five_seq ? rc_seq(p->five_seq, (int)strlen(p->five_seq)) : NULL;
    char *three_rc = p->three_seq ? rc_seq(p->three_seq, (int)strlen(p->three_seq)) : NULL;
    for (ci = 0; ci < ch_n; ++ci) {
        int *par_pos, par_n = tho_partition(bseq, len, &ch, ci, p, &par_pos);
        if (par_n >= p->min_copy + 1) seqs_msa(len, bseq, par_n, par_pos, r, p, five_rc, three_rc);
        free(par_pos);
    }
    if (p->single_copy == 1 && p->only_full_length && p->five_seq && p->three_seq)
        single_copy_full_len(len, seq, r, p, five_rc, three_rc);
    free(five_rc); free(three_rc);
    if (hit_n >= 2) tho_chain_free(&ch);
    free(bseq);
}
void tho_read_free(tho_read_t *r) {
    int i;
    for (i = 0; i < r->n_cons; ++i) free_cons(r->cons + i);
    free(r->cons); free(r->dropped_qual);
    memset(r, 0, sizeof(*r));
}

/* ------------------------------------------------------------------ output (src/main.c:214-271) */
static void bprintf(char **buf, size_t *l, size_t *m, const char *fmt, ...) __attribute__((format(printf, 4, 5)));
#include <stdarg.h>
static void bprintf(char **buf, size_t *l, size_t *m, const char *fmt, ...) {
    va_list ap; int n;
    while (1) {
        va_start(ap, fmt);
        n = vsnprintf(*buf ? *buf + *l : NULL, *buf ? *m - *l : 0, fmt, ap);
        va_end(ap);
        if (*buf && (size_t)n < *m - *l) break;
        *m = (*m + n + 1) * 2; *buf = (char *)realloc(*buf, *m);
    }
    *l += n;
}
static void bwrite(char **buf, size_t *l, size_t *m, const char *s, size_t n) {
    if (*l + n + 1 > *m) { *m = (*l + n + 1) * 2; *buf = (char *)realloc(*buf, *m); }
    memcpy(*buf + *l, s, n); *l += n; (*buf)[*l] = 0;
}

/* `quals` lets the caller emulate the reference's never-reset qual.l (main.c:266-267): it points at
 * the text the reference would print for each record's quality line, or NULL to use the record's own */
static void format_read(const char *name, const char *seq, int len, const tho_read_t *r, const tho_para_t *p,
                        char **buf, size_t *bl, size_t *bm, const char *qual_override) {
    int ci, i, j; size_t qoff = 0;
    for (ci = 0; ci < r->n_cons; ++ci) {
        const tho_cons_t *c = r->cons + ci;
        if (p->only_unit) {
            if (p->out_fmt == 1) {
                for (i = 0; i < c->pos_n - 1; ++i) {
                    bprintf(buf, bl, bm, ">%s_rep%d_sub%d\n", name, ci, i);
                    for (j = c->sub_pos[i] + 1; j <= c->sub_pos[i + 1]; ++j) bwrite(buf, bl, bm, seq + j, 1);
                    bwrite(buf, bl, bm, "\n", 1);
                }
            } else if (p->out_fmt == 2) {
                for (i = 0; i < c->pos_n - 1; ++i) {
                    bprintf(buf, bl, bm, "%s\trep%d\tsub%d\t", name, ci, i);
                    for (j = c->sub_pos[i] + 1; j < c->sub_pos[i + 1]; ++j) bwrite(buf, bl, bm, seq + j, 1);
                    bwrite(buf, bl, bm, "\n", 1);
                }
            }
        } else {
            if (p->out_fmt == 1 || p->out_fmt == 3) {
                bprintf(buf, bl, bm, "%c%s_rep%d_%.1f %d_%d_%d_%d_%.1f_%d_", p->out_fmt == 1 ? '>' : '@', name, ci, c->copy_num, len,
                        c->cons_start + 1, c->cons_end + 1, c->cons_len, c->ave_match, c->full_length);
            } else {
                bprintf(buf, bl, bm, "%s\trep%d\t%.1f\t%d\t%d\t%d\t%d\t%.1f\t%d\t", name, ci, c->copy_num, len,
                        c->cons_start + 1, c->cons_end + 1, c->cons_len, c->ave_match, c->full_length);
            }
            bprintf(buf, bl, bm, "%d", c->sub_pos[0] + 2);
            for (i = 1; i < c->pos_n - 1; ++i) bprintf(buf, bl, bm, ",%d", c->sub_pos[i] + 2);
            bprintf(buf, bl, bm, ",%d%c", c->sub_pos[i] + 1, (p->out_fmt == 1 || p->out_fmt == 3) ? '\n' : '\t');
            bwrite(buf, bl, bm, c->cons_seq, c->cons_len);
            if (p->out_fmt == 3) bwrite(buf, bl, bm, "\n+\n", 3);
            else if (p->out_fmt == 4) bwrite(buf, bl, bm, "\t", 1);
            if (p->out_fmt == 3 || p->out_fmt == 4) {
                if (qual_override) bwrite(buf, bl, bm, qual_override + qoff, c->cons_len);
                else bwrite(buf, bl, bm, c->cons_qual, c->cons_len);
                qoff += c->cons_len;
            }
            bwrite(buf, bl, bm, "\n", 1);
        }
    }
}
void tho_format_read(const char *name, const char *seq, int len, const tho_read_t *r, const tho_para_t *p, char **buf, size_t *bl, size_t *bm) {
    format_read(name, seq, len, r, p, buf, bl, bm, NULL);
}

/* ------------------------------------------------------------------ batch runner */
typedef struct { int n; const char *const *seqs; const int *lens; const tho_para_t *p; tho_read_t *res; volatile int *next; } job_t;
static void *worker(void *a) {
    job_t *jb = (job_t *)a;
    while (1) {
        int i = __sync_fetch_and_add(jb->next, 1);
        if (i >= jb->n) break;
        tho_process_read(jb->seqs[i], jb->lens[i], jb->p, jb->res + i);
    }
    return NULL;
}

#define CHUNK_READ_N 4096 /* src/tidehunter.h:10 */
char *tho_run_batch(int n, const char *const *names, const char *const *seqs, const int *lens, const tho_para_t *p,
                    int n_threads, size_t *out_len, int64_t counters[4]) {
    tho_read_t *res = (tho_read_t *)calloc(n > 0 ? n : 1, sizeof(tho_read_t));
    volatile int next = 0; int i;
    job_t jb = {n, seqs, lens, p, res, &next};
    if (n_threads <= 1) worker(&jb);
    else {
        pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
        for (i = 0; i < n_threads; ++i) pthread_create(tid + i, NULL, worker, &jb);
        for (i = 0; i < n_threads; ++i) pthread_join(tid[i], NULL);
        free(tid);
    }
    char *buf = NULL; size_t bl = 0, bm = 0;
    bwrite(&buf, &bl, &bm, "", 0);
    /* qual.l is never reset in the reference (main.c:266-267 vs gen_cons.c:23-30): each of the 4096
     * chunk slots keeps appending qualities while printing always starts at offset 0 */
    slot_t *slots = (slot_t *)calloc(CHUNK_READ_N, sizeof(slot_t));
    int with_qual = (p->out_fmt == 3 || p->out_fmt == 4) && !p->only_unit;
    if (counters) counters[0] = counters[1] = counters[2] = counters[3] = 0;
    for (i = 0; i < n; ++i) {
        const char *qov = NULL;
        if (with_qual) {
            slot_t *s = slots + i % CHUNK_READ_N; int ci; size_t need = res[i].dropped_l;
            for (ci = 0; ci < res[i].n_cons; ++ci) need += res[i].cons[ci].cons_len;
            if (s->qual_l + need + 1 > s->qual_m) {
                size_t m2 = (s->qual_l + need + 1) * 2;
                s->qual_s = (char *)realloc(s->qual_s, m2);
                memset(s->qual_s + s->qual_m, '?', m2 - s->qual_m); /* the reference would print uninitialised heap here */
                s->qual_m = m2;
            }
            if (res[i].dropped_l) { memcpy(s->qual_s + s->qual_l, res[i].dropped_qual, res[i].dropped_l); s->qual_l += res[i].dropped_l; }
            for (ci = 0; ci < res[i].n_cons; ++ci) {
                memcpy(s->qual_s + s->qual_l, res[i].cons[ci].cons_qual, res[i].cons[ci].cons_len);
                s->qual_l += res[i].cons[ci].cons_len;
            }
            qov = s->qual_s;
        }
        format_read(names[i], seqs[i], lens[i], res + i, p, &buf, &bl, &bm, qov);
        if (counters) { counters[0] += res[i].n_hits; counters[1] += res[i].n_chain_evals; counters[2] += res[i].n_poa_cells; counters[3] += res[i].n_ksw_cells; }
        tho_read_free(res + i);
    }
    for (i = 0; i < CHUNK_READ_N; ++i) free(slots[i].qual_s);
    free(slots); free(res);
    *out_len = bl;
    return buf;
}
void tho_free(void *p) { free(p); }
