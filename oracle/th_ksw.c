/*
 * th_ksw.c -- plain-integer restatement of the ksw2 calls on TideHunter's hot path.
 * TEST INFRASTRUCTURE ONLY (see th_oracle.h).
 *
 * Reference: ksw2/ksw2_extz2_sse.c:23-304 (Suzuki-Kasahara differential SSE kernel),
 *            ksw2/ksw2.h:119-176 (backtrack state machine, exact-max bookkeeping),
 *            src/ksw2_align.c:11-17 (scoring 1/-2, gap 2+1*g), :62-173 (wrappers).
 * The reference stores 8-bit differences; mathematically it is the affine recurrence below on
 * 32-bit scores, with the tie-breaking rules of the traceback flags and of the per-anti-diagonal
 * arg-max.  Only those rules and the boundary conditions are restated, not the SIMD layout.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "th_oracle.h"

#define KQ 2   /* gap_open, src/ksw2_align.c:11 */
#define KE 1   /* gap_ext */
#define NEG (-0x40000000)

/* ksw2_extz2_sse.c:57-59 with mat from ksw2_align.c:12-17: match 1, mismatch -2,
 * wildcard (code m-1 = 4) scores -e because mat[24] == 0. */
static inline int sc(uint8_t a, uint8_t b) {
    if (a == 4 || b == 4) return -KE;
    return a == b ? 1 : -2;
}

static uint32_t *push_cigar(int *n, int *m, uint32_t *c, uint32_t op, int len) { /* ksw2.h:103-114 */
    if (*n == 0 || op != (c[*n - 1] & 0xf)) {
        if (*n == *m) { *m = *m ? *m << 1 : 4; c = (uint32_t *)realloc(c, (size_t)*m * 4); }
        c[(*n)++] = (uint32_t)len << 4 | op;
    } else c[*n - 1] += (uint32_t)len << 4;
    return c;
}

/* Global alignment, w = -1, zdrop = -1, flag = 0 (ksw2_align.c:117-151).
 * Recurrence (target index i = rows, query index j = columns):
 *   H(i,j) = max(H(i-1,j-1)+s, E(i,j), F(i,j));  E(i+1,j) = max(H(i,j)-q, E(i,j)) - e;  F likewise.
 * Flags per cell (ksw2_extz2_sse.c:171-196): d = 1 iff E > diag (strict), then 2 iff F > max (strict);
 * E-continuation iff E(i,j) > H(i,j)-q, F-continuation likewise (strict).
 * Returns the number of M columns with equal codes (ksw2_get_xid, ksw2_align.c:62-86; N==N counts). */
int tho_ksw2_global(const uint8_t *q, int ql, const uint8_t *t, int tl, int *n_cigar_, uint32_t **cigar_) {
    if (n_cigar_) *n_cigar_ = 0;
    if (cigar_) *cigar_ = NULL;
    if (ql <= 0 || tl <= 0) return 0;
    uint8_t *p = (uint8_t *)malloc((size_t)ql * tl);
    int *Hrow = (int *)malloc(sizeof(int) * (ql + 1)); /* Hrow[j+1] = H(i-1, j); Hrow[0] = H(i-1,-1) */
    int *E = (int *)malloc(sizeof(int) * ql);          /* E(i, j) arriving at the current row */
    int i, j;
    Hrow[0] = 0;
    for (j = 0; j < ql; ++j) { Hrow[j + 1] = -(KQ + KE * (j + 1)); E[j] = Hrow[j + 1] - KQ - KE; }
    for (i = 0; i < tl; ++i) {
        int hleft = -(KQ + KE * (i + 1)); /* H(i,-1) */
        int hdiag = Hrow[0];              /* H(i-1,-1) */
        int F = hleft - KQ - KE;          /* F(i,0) */
        Hrow[0] = hleft;
        for (j = 0; j < ql; ++j) {
            int z = hdiag + sc(t[i], q[j]), d = 0, e = E[j];
            if (e > z) { d = 1; z = e; }
            if (F > z) { d = 2; z = F; }
            if (e > z - KQ) d |= 0x08;
            if (F > z - KQ) d |= 0x10;
            p[(size_t)i * ql + j] = (uint8_t)d;
            E[j] = (e > z - KQ ? e : z - KQ) - KE;
            F = (F > z - KQ ? F : z - KQ) - KE;
            hdiag = Hrow[j + 1];
            Hrow[j + 1] = z;
        }
    }
    /* ksw_backtrack, ksw2.h:119-151 (is_rot band bookkeeping never forces a state for a full matrix) */
    int n = 0, m = 0, state = 0; uint32_t *cig = NULL;
    i = tl - 1; j = ql - 1;
    while (i >= 0 && j >= 0) {
        int tmp = p[(size_t)i * ql + j];
        if (state == 0) state = tmp & 7;
        else if (!(tmp >> (state + 2) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        if (state == 0) { cig = push_cigar(&n, &m, cig, 0, 1); --i; --j; }
        else if (state == 1) { cig = push_cigar(&n, &m, cig, 2, 1); --i; }
        else { cig = push_cigar(&n, &m, cig, 1, 1); --j; }
    }
    if (i >= 0) cig = push_cigar(&n, &m, cig, 2, i + 1);
    if (j >= 0) cig = push_cigar(&n, &m, cig, 1, j + 1);
    for (i = 0; i < n >> 1; ++i) { uint32_t x = cig[i]; cig[i] = cig[n - 1 - i]; cig[n - 1 - i] = x; }
    /* ksw2_get_xid */
    int iden = 0, qi = 0, ti = 0;
    for (i = 0; i < n; ++i) {
        int op = cig[i] & 0xf, len = cig[i] >> 4;
        if (op == 0) { for (j = 0; j < len; ++j) if (q[qi + j] == t[ti + j]) ++iden; qi += len; ti += len; }
        else if (op == 1) qi += len;
        else ti += len;
    }
    free(p); free(Hrow); free(E);
    if (n_cigar_) *n_cigar_ = n;
    if (cigar_) *cigar_ = cig; else free(cig);
    return iden;
}

/* src/ksw2_align.c:88-115 */
int tho_ksw2_backtrack_left_end(int n_cigar, const uint32_t *cigar, int qlen, int tlen, int q_left_ext) {
    int t_left_ext = 0, i, q_remain = q_left_ext;
    (void)qlen; (void)tlen;
    for (i = n_cigar - 1; i >= 0; --i) {
        int op = cigar[i] & 0xf, len = cigar[i] >> 4;
        if (op == 0) {
            if (len >= q_remain) return t_left_ext + q_remain;
            t_left_ext += len; q_remain -= len;
        } else if (op == 1) {
            if (len >= q_remain) return t_left_ext;
            q_remain -= len;
        } else if (op == 2) t_left_ext += len;
    }
    if (q_remain > 0) { fprintf(stderr, "[tho] Error: unmatched cigar and q_left_ext.\n"); exit(1); }
    return t_left_ext;
}

/* Extension, flag = EXTZ_ONLY|SCORE_ONLY, zdrop = -1 (ksw2_align.c:153-159).
 * Result = first cell reaching the global maximum of H (must be > 0, ksw2.h:153-176 with
 * ez->max starting at 0) in the reference's visiting order: anti-diagonal r = i+j ascending, and
 * inside a diagonal [st0,en0]: en0, then 4 interleaved lanes over [st0,en1), then [en1,en0)
 * (ksw2_extz2_sse.c:224-261).  The order is encoded as a rank so a row-major sweep can apply it. */
static inline int64_t diag_rank(int t, int r, int ql, int tl) {
    int st0 = r - ql + 1 > 0 ? r - ql + 1 : 0, en0 = r < tl - 1 ? r : tl - 1;
    if (t == en0) return 0;
    int en1 = st0 + (en0 - st0) / 4 * 4;
    if (t < en1) return 1 + (int64_t)((t - st0) & 3) * 0x40000000LL + (t - st0) / 4;
    return 1 + 4 * 0x40000000LL + (t - en1);
}

void tho_ksw2_ext(const uint8_t *q, int ql, const uint8_t *t, int tl, int *max_q, int *max_t) {
    *max_q = *max_t = -1;
    if (ql <= 0 || tl <= 0) return;
    int *Hrow = (int *)malloc(sizeof(int) * (ql + 1));
    int *E = (int *)malloc(sizeof(int) * ql);
    int i, j, best = 0, best_r = 0; int64_t best_rank = 0;
    Hrow[0] = 0;
    for (j = 0; j < ql; ++j) { Hrow[j + 1] = -(KQ + KE * (j + 1)); E[j] = Hrow[j + 1] - KQ - KE; }
    for (i = 0; i < tl; ++i) {
        int hleft = -(KQ + KE * (i + 1)), hdiag = Hrow[0], F = hleft - KQ - KE;
        Hrow[0] = hleft;
        for (j = 0; j < ql; ++j) {
            int z = hdiag + sc(t[i], q[j]), e = E[j];
            if (e > z) z = e;
            if (F > z) z = F;
            E[j] = (e > z - KQ ? e : z - KQ) - KE;
            F = (F > z - KQ ? F : z - KQ) - KE;
            hdiag = Hrow[j + 1];
            Hrow[j + 1] = z;
            if (z > 0 && z >= best) {
                int r = i + j;
                if (z > best || r < best_r || (r == best_r && diag_rank(i, r, ql, tl) < best_rank)) {
                    best = z; best_r = r; best_rank = diag_rank(i, r, ql, tl);
                    *max_t = i; *max_q = j;
                }
            }
        }
    }
    free(Hrow); free(E);
}

/* src/ksw2_align.c:161-173 */
void tho_ksw2_left_ext(const uint8_t *q, int ql, const uint8_t *t, int tl, int *max_q, int *max_t) {
    uint8_t *rq = (uint8_t *)malloc(ql > 0 ? ql : 1), *rt = (uint8_t *)malloc(tl > 0 ? tl : 1);
    int i;
    for (i = 0; i < ql; ++i) rq[i] = q[ql - i - 1];
    for (i = 0; i < tl; ++i) rt[i] = t[tl - i - 1];
    tho_ksw2_ext(rq, ql, rt, tl, max_q, max_t);
    free(rq); free(rt);
}
