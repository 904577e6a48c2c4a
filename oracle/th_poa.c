/*
 * th_poa.c -- scalar restatement of the abPOA (v1.3.0) calls on TideHunter's hot path:
 * global, convex-gap, adaptive-banded sequence-to-graph alignment + heaviest-column consensus.
 * TEST INFRASTRUCTURE ONLY (see th_oracle.h).
 *
 * The reference computes rows in SIMD vectors of `pn` lanes; band edges, the in-vector F
 * propagation and the row arg-max are all vector-granular, so `pn` is a parameter here (16 = AVX2
 * int16, the reference default).  Score arithmetic wraps to 16 bit exactly like _mm256_add_epi16 /
 * _mm256_sub_epi16 when the reference would pick its int16 path.
 *
 * Reference: abPOA/src/simd_abpoa_align.c (convex path: :248-377 backtrack, :382-479 variables,
 *   :538-610 first row, :613-647 SIMD_SET_F, :835-958 row DP, :976-1015 max/adaptive band,
 *   :1079-1107 core, :1583-1650 entry), abPOA/src/abpoa_graph.c (:150-277 sorts, :279-359 msa rank,
 *   :604-648 heaviest column, :1020-1124 node/edge, :1218-1288 add alignment),
 *   abPOA/src/abpoa_align.c:293-411 (driver), abPOA/src/abpoa_align.h:34-35 (band macros).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "th_oracle.h"

#define MAX2(a, b) ((a) > (b) ? (a) : (b))
#define MIN2(a, b) ((a) < (b) ? (a) : (b))
#define MAX3(a, b, c) MAX2(MAX2(a, b), c)
#define MIN3(a, b, c) MIN2(MIN2(a, b), c)

typedef struct {
    int in_n, in_m, *in_id;
    int out_n, out_m, *out_id, *out_w;
    int aln_n, aln_m, *aln_id;
    uint64_t *read_ids; int read_ids_n;
    int max_out_id;
    uint8_t base;
} pnode_t;

typedef struct {
    pnode_t *node; int node_n, node_m;
    int *index_to_node_id, *node_id_to_index, *max_pos_left, *max_pos_right, *max_remain, *msa_rank;
    int aux_m;
} pgraph_t;

enum { OP_M = 0, OP_I = 1, OP_D = 2 };
typedef struct { int op, node, q, len; } pcig_t; /* M: node,q ; D: node ; I: q (last query pos), len */

#define PUSH(arr, n, m, type, val) do { if ((n) == (m)) { (m) = (m) ? (m) << 1 : 4; (arr) = (type *)realloc((arr), sizeof(type) * (size_t)(m)); } (arr)[(n)++] = (val); } while (0)

static pgraph_t *g_init(void) {
    pgraph_t *g = (pgraph_t *)calloc(1, sizeof(pgraph_t));
    g->node_m = 1024; g->node = (pnode_t *)calloc(g->node_m, sizeof(pnode_t));
    g->node_n = 2; /* 0 = source, 1 = sink (abpoa.h:25-26) */
    return g;
}
static void g_free(pgraph_t *g) {
    int i;
    for (i = 0; i < g->node_m; ++i) {
        free(g->node[i].in_id); free(g->node[i].out_id); free(g->node[i].out_w);
        free(g->node[i].aln_id); free(g->node[i].read_ids);
    }
    free(g->node); free(g->index_to_node_id); free(g->node_id_to_index); free(g->max_pos_left);
    free(g->max_pos_right); free(g->max_remain); free(g->msa_rank); free(g);
}
static int g_add_node(pgraph_t *g, uint8_t base) { /* abpoa_graph.c:1054-1061 */
    if (g->node_n == g->node_m) {
        g->node = (pnode_t *)realloc(g->node, sizeof(pnode_t) * (size_t)g->node_m * 2);
        memset(g->node + g->node_m, 0, sizeof(pnode_t) * (size_t)g->node_m);
        g->node_m <<= 1;
    }
    g->node[g->node_n].base = base;
    return g->node_n++;
}
/* abpoa_graph.c:1063-1106 */
static void g_add_edge(pgraph_t *g, int from, int to, int check_edge, int w, int add_read_id, int read_id, int read_ids_n) {
    pnode_t *f = g->node + from, *t = g->node + to;
    int i;
    if (check_edge) {
        for (i = 0; i < f->out_n; ++i)
            if (f->out_id[i] == to) { f->out_w[i] += w; goto ADD_READ_ID; }
    }
    PUSH(t->in_id, t->in_n, t->in_m, int, from);
    { int m2 = f->out_m, n2 = f->out_n; PUSH(f->out_w, n2, m2, int, w); }
    PUSH(f->out_id, f->out_n, f->out_m, int, to);
ADD_READ_ID:
    if (add_read_id) {
        if (f->read_ids_n < read_ids_n) {
            f->read_ids = (uint64_t *)realloc(f->read_ids, sizeof(uint64_t) * read_ids_n);
            for (i = f->read_ids_n; i < read_ids_n; ++i) f->read_ids[i] = 0;
            f->read_ids_n = read_ids_n;
        }
        f->read_ids[read_id / 64] |= 1ULL << (read_id & 0x3f);
    }
}
/* abpoa_graph.c:1020-1044 */
static int g_get_aligned_id(pgraph_t *g, int node_id, uint8_t base) {
    int i;
    for (i = 0; i < g->node[node_id].aln_n; ++i) {
        int a = g->node[node_id].aln_id[i];
        if (g->node[a].base == base) return a;
    }
    return -1;
}
static void g_add_aligned1(pnode_t *n, int id) { PUSH(n->aln_id, n->aln_n, n->aln_m, int, id); }
static void g_add_aligned(pgraph_t *g, int node_id, int aligned_id) {
    int i, n0 = g->node[node_id].aln_n;
    for (i = 0; i < n0; ++i) {
        int a = g->node[node_id].aln_id[i];
        g_add_aligned1(g->node + a, aligned_id);
        g_add_aligned1(g->node + aligned_id, a);
    }
    g_add_aligned1(g->node + node_id, aligned_id);
    g_add_aligned1(g->node + aligned_id, node_id);
}
/* abpoa_graph.c:1108-1124 (first sequence) */
static void g_add_sequence(pgraph_t *g, const uint8_t *seq, int seq_l, int read_id, int read_ids_n) {
    int i, last = 0, cur;
    for (i = 0; i < seq_l; ++i) {
        cur = g_add_node(g, seq[i]);
        g_add_edge(g, last, cur, 0, 1, 1, read_id, read_ids_n);
        last = cur;
    }
    g_add_edge(g, last, 1, 0, 1, 1, read_id, read_ids_n);
}
/* abpoa_graph.c:1218-1284 with beg = source, end = sink, inc_both_ends = 1, use_read_ids = 1 */
static void g_add_alignment(pgraph_t *g, const uint8_t *seq, int seq_l, const pcig_t *cig, int n_cig, int read_id, int tot_read_n) {
    int read_ids_n = 1 + ((tot_read_n - 1) >> 6);
    if (g->node_n == 2) { g_add_sequence(g, seq, seq_l, read_id, read_ids_n); return; }
    if (n_cig == 0) return;
    int i, j, query_id = -1, last_new = 0, last_id = 0, new_id, aligned_id;
    for (i = 0; i < n_cig; ++i) {
        if (cig[i].op == OP_M) {
            int node_id = cig[i].node;
            query_id++;
            if (g->node[node_id].base != seq[query_id]) {
                if ((aligned_id = g_get_aligned_id(g, node_id, seq[query_id])) != -1) {
                    g_add_edge(g, last_id, aligned_id, 1 - last_new, 1, 1, read_id, read_ids_n);
                    last_id = aligned_id; last_new = 0;
                } else {
                    new_id = g_add_node(g, seq[query_id]);
                    g_add_edge(g, last_id, new_id, 0, 1, 1, read_id, read_ids_n);
                    last_id = new_id; last_new = 1;
                    g_add_aligned(g, node_id, new_id);
                }
            } else {
                g_add_edge(g, last_id, node_id, 1 - last_new, 1, 1, read_id, read_ids_n);
                last_id = node_id; last_new = 0;
            }
        } else if (cig[i].op == OP_I) {
            int len = cig[i].len;
            query_id += len;
            for (j = len - 1; j >= 0; --j) {
                new_id = g_add_node(g, seq[query_id - j]);
                g_add_edge(g, last_id, new_id, 0, 1, 1, read_id, read_ids_n);
                last_id = new_id; last_new = 1;
            }
        }
    }
    g_add_edge(g, last_id, 1, 1 - last_new, 1, 1, read_id, read_ids_n);
}

static void g_aux_reserve(pgraph_t *g) {
    if (g->node_n > g->aux_m) {
        g->aux_m = g->node_n * 2;
        g->index_to_node_id = (int *)realloc(g->index_to_node_id, sizeof(int) * g->aux_m);
        g->node_id_to_index = (int *)realloc(g->node_id_to_index, sizeof(int) * g->aux_m);
        g->max_pos_left = (int *)realloc(g->max_pos_left, sizeof(int) * g->aux_m);
        g->max_pos_right = (int *)realloc(g->max_pos_right, sizeof(int) * g->aux_m);
        g->max_remain = (int *)realloc(g->max_remain, sizeof(int) * g->aux_m);
        g->msa_rank = (int *)realloc(g->msa_rank, sizeof(int) * g->aux_m);
    }
}

/* abpoa_graph.c:150-277 */
static void g_topological_sort(pgraph_t *g) {
    int n = g->node_n, i, j;
    g_aux_reserve(g);
    int *deg = (int *)malloc(sizeof(int) * n), *que = (int *)malloc(sizeof(int) * (n + 1));
    int qh = 0, qt = 0, index = 0;
    for (i = 0; i < n; ++i) deg[i] = g->node[i].in_n;
    que[qt++] = 0;
    while (qh < qt) { /* BFS_set_node_index */
        int cur = que[qh++];
        g->index_to_node_id[index] = cur; g->node_id_to_index[cur] = index++;
        if (cur == 1) break;
        for (i = 0; i < g->node[cur].out_n; ++i) {
            int out = g->node[cur].out_id[i], ok = 1;
            if (--deg[out] == 0) {
                for (j = 0; j < g->node[out].aln_n; ++j)
                    if (deg[g->node[out].aln_id[j]] != 0) { ok = 0; break; }
                if (!ok) continue;
                que[qt++] = out;
                for (j = 0; j < g->node[out].aln_n; ++j) que[qt++] = g->node[out].aln_id[j];
            }
        }
    }
    if (index != n) { fprintf(stderr, "[tho] Failed to set node index (%d/%d).\n", index, n); exit(1); }
    for (i = 0; i < n; ++i) { g->max_pos_right[i] = 0; g->max_pos_left[i] = n; }
    /* BFS_set_node_remain */
    for (i = 0; i < n; ++i) { deg[i] = g->node[i].out_n; g->max_remain[i] = 0; }
    qh = qt = 0; que[qt++] = 1; g->max_remain[1] = -1;
    while (qh < qt) {
        int cur = que[qh++];
        if (cur != 1) {
            int max_w = -1, max_id = -1;
            for (i = 0; i < g->node[cur].out_n; ++i)
                if (g->node[cur].out_w[i] > max_w) { max_w = g->node[cur].out_w[i]; max_id = g->node[cur].out_id[i]; }
            g->max_remain[cur] = g->max_remain[max_id] + 1;
        }
        if (cur == 0) break;
        for (i = 0; i < g->node[cur].in_n; ++i) {
            int in = g->node[cur].in_id[i];
            if (--deg[in] == 0) que[qt++] = in;
        }
    }
    free(deg); free(que);
}

/* ---------------- banded convex-gap DP ---------------- */
typedef struct {
    int bits, pn, log_n, inf_min;
    int qlen, dp_sn, gn;
    int *dp_beg, *dp_end, *dp_beg_sn, *dp_end_sn;
    size_t *row_off;  /* offset (in values) of vector dp_beg_sn of each row inside each state array */
    int32_t *H, *E1, *E2, *F1, *F2; size_t cap, used;
} pdp_t;

static inline int32_t Wv(const pdp_t *d, int64_t x) { return d->bits == 16 ? (int32_t)(int16_t)x : (int32_t)x; }

static void dp_alloc_row(pdp_t *d, int i, int beg_sn, int end_sn) {
    size_t need = (size_t)(end_sn - beg_sn + 2) * d->pn, k;
    if (d->used + need > d->cap) {
        d->cap = (d->used + need) * 2;
        d->H = (int32_t *)realloc(d->H, d->cap * 4); d->E1 = (int32_t *)realloc(d->E1, d->cap * 4);
        d->E2 = (int32_t *)realloc(d->E2, d->cap * 4); d->F1 = (int32_t *)realloc(d->F1, d->cap * 4);
        d->F2 = (int32_t *)realloc(d->F2, d->cap * 4);
    }
    d->row_off[i] = d->used;
    for (k = 0; k < need; ++k) /* poison: nothing below may depend on unwritten cells */
        d->H[d->used + k] = d->E1[d->used + k] = d->E2[d->used + k] = d->F1[d->used + k] = d->F2[d->used + k] = 0x3f3f3f3f;
    d->used += need;
}
/* address of column c of row i in state array A (c must lie in the stored vectors) */
#define AT(A, i, c) ((A)[d->row_off[i] + (size_t)((c) - d->dp_beg[i])])
static inline int stored(const pdp_t *d, int i, int sn) { return sn >= d->dp_beg_sn[i] && sn <= d->dp_end_sn[i] + 1; }

/* simd_abpoa_align.c:613-647 : in-vector max-plus prefix propagation of F, lane by lane */
static void set_F(const pdp_t *d, int32_t *F, int set_num, int e) {
    int pn = d->pn, k, l, cov_bit = set_num; int32_t tmp[64];
    for (k = 0; k < d->log_n; ++k) {
        int sh = 1 << k;
        if (set_num == pn) {
            for (l = 0; l < pn; ++l) tmp[l] = l >= sh ? Wv(d, (int64_t)F[l - sh] - (int64_t)e * sh) : d->inf_min;
        } else {
            if (k > 0) cov_bit += sh;
            int c = cov_bit > pn ? pn : cov_bit; /* PRE_MASK[pn] == PRE_MASK[pn-1], SUF_MIN[pn] == SUF_MIN[pn-1] */
            for (l = 0; l < pn; ++l) tmp[l] = (l >= sh && l <= c) ? Wv(d, (int64_t)F[l - sh] - (int64_t)e * sh) : d->inf_min;
        }
        for (l = 0; l < pn; ++l) if (tmp[l] > F[l]) F[l] = tmp[l];
    }
}

static int poa_align(pgraph_t *g, const tho_para_t *p, const uint8_t *query, int qlen, pcig_t **cig_, int *n_cig_, int64_t *cells) {
    int m = 5, mat[25], i, j, k, l;
    { /* gen_simple_mat, abpoa_align.c:10-23 */
        int match = p->match < 0 ? -p->match : p->match, mis = p->mismatch > 0 ? -p->mismatch : p->mismatch;
        for (i = 0; i < m - 1; ++i) { for (j = 0; j < m - 1; ++j) mat[i * m + j] = i == j ? match : mis; mat[i * m + m - 1] = 0; }
        for (j = 0; j < m; ++j) mat[(m - 1) * m + j] = 0;
    }
    int max_mat = p->match < 0 ? -p->match : p->match, min_mis = p->mismatch > 0 ? p->mismatch : -p->mismatch;
    int o1 = p->gap_open1, e1 = p->gap_ext1, o2 = p->gap_open2, e2 = p->gap_ext2, oe1 = o1 + e1, oe2 = o2 + e2;
    const int oe2_raw = oe2, e2_raw = e2; /* inf_min and the int16 choice use the options as given in every gap mode (:1613-1614) */
    /* affine mode (gap_open2 == 0, abpoa_align.c:85-88): simd_abpoa_ag_dp (:739-833) is NOT the convex recurrence with
     * one gap function removed: an insertion opens from M only (F from the row's diagonal values, before E is folded
     * in, :807-815), and the E handed to the next row is inf_min wherever F strictly won the cell (SIMDSetIfEqual,
     * :822-827) -- insertions and deletions cannot be adjacent.  The second pair (E2, F2) is held at inf_min, so the
     * convex backtrack (:248-377) degenerates to simd_abpoa_ag_backtrack (:160-246).  Pinned on reference outputs
     * (tests/golden/gapmode_golden.json, tools/option_fuzz.py). */
    const int affine = p->gap_open1 > 0 && p->gap_open2 == 0;
    /* linear mode (gap_open1 == 0, abpoa_align.c:86): simd_abpoa_lg_first_dp / lg_dp / lg_backtrack (:557-570, :649-736,
     * :108-158) -- a single matrix H = max(M, max_pre H[pre][j] - e, H[j-1] - e); the oracle only (the GPU path rejects
     * the option), pinned by tools/option_fuzz.py against the reference */
    const int linear = p->gap_open1 == 0;
    int beg_index = g->node_id_to_index[0], end_index = g->node_id_to_index[1], gn = end_index - beg_index + 1;
    pdp_t D, *d = &D; memset(d, 0, sizeof(D));
    { /* simd_abpoa_align.c:1610-1621 */
        int len = qlen > gn ? qlen : gn, max_score = MAX2(qlen * max_mat, len * e1 + o1);
        if (max_score <= INT16_MAX - min_mis - oe1 - oe2_raw) {
            d->bits = 16; d->pn = p->pn16;
            d->inf_min = MAX3(INT16_MIN + min_mis, INT16_MIN + oe1, INT16_MIN + oe2_raw) + 31 * MAX2(e1, e2_raw);
        } else {
            d->bits = 32; d->pn = p->pn16 / 2;
            if (getenv("THO_DEBUG")) fprintf(stderr, "[tho] int32 alignment: qlen %d, graph rows %d\n", qlen, gn);
            d->inf_min = MAX3(INT32_MIN + min_mis, INT32_MIN + oe1, INT32_MIN + oe2_raw) + 31 * MAX2(e1, e2_raw);
        }
    }
    int pn = d->pn, inf_min = d->inf_min;
    for (d->log_n = 0; (1 << d->log_n) < pn; ++d->log_n) ;
    int dp_sn = (qlen + pn) / pn; /* (matrix_col_n + pn - 1) / pn, :396 */
    int w = 10 + (int)(0.01f * qlen); /* wb + (int)(wf*qlen), float arithmetic, :393 */
    d->qlen = qlen; d->dp_sn = dp_sn; d->gn = gn;
    d->dp_beg = (int *)malloc(sizeof(int) * gn); d->dp_end = (int *)malloc(sizeof(int) * gn);
    d->dp_beg_sn = (int *)malloc(sizeof(int) * gn); d->dp_end_sn = (int *)malloc(sizeof(int) * gn);
    d->row_off = (size_t *)malloc(sizeof(size_t) * gn);
    /* query profile (:438-446) and query index (:447-451) */
    int32_t *qp = (int32_t *)malloc(sizeof(int32_t) * (size_t)m * dp_sn * pn), *qi = (int32_t *)malloc(sizeof(int32_t) * (size_t)(qlen / pn + 1) * pn);
    for (k = 0; k < m; ++k) {
        int32_t *_qp = qp + (size_t)k * dp_sn * pn; _qp[0] = 0;
        for (j = 0; j < qlen; ++j) _qp[j + 1] = mat[k * m + query[j]];
        for (j = qlen + 1; j < dp_sn * pn; ++j) _qp[j] = 0;
    }
    for (i = 0; i <= qlen; ++i) qi[i] = i;
    for (i = qlen + 1; i < (qlen / pn + 1) * pn; ++i) qi[i] = -1;
    /* predecessor lists in in_id order (:461-471); every node is reachable from the source here */
    int **pre_index = (int **)calloc(gn, sizeof(int *)), *pre_n = (int *)calloc(gn, sizeof(int));
    for (i = 1; i < gn; ++i) {
        int node_id = g->index_to_node_id[beg_index + i];
        pre_n[i] = g->node[node_id].in_n;
        pre_index[i] = (int *)malloc(sizeof(int) * (pre_n[i] ? pre_n[i] : 1));
        for (j = 0; j < pre_n[i]; ++j) pre_index[i][j] = g->node_id_to_index[g->node[node_id].in_id[j]] - beg_index;
    }
#define REMAIN(id) (g->max_remain[id] - g->max_remain[1] - 1)
#define AD_BEGIN(id) MAX2(0, MIN2(g->max_pos_left[id], qlen - REMAIN(id)) - w)
#define AD_END(id) MIN2(qlen, MAX2(g->max_pos_right[id], qlen - REMAIN(id)) + w)
    /* first row (:538-555, :591-610) */
    g->max_pos_left[0] = g->max_pos_right[0] = 0;
    for (i = 0; i < g->node[0].out_n; ++i) { int o = g->node[0].out_id[i]; g->max_pos_left[o] = g->max_pos_right[o] = 1; }
    d->dp_beg_sn[0] = 0; d->dp_end_sn[0] = AD_END(0) / pn;
    d->dp_beg[0] = 0; d->dp_end[0] = (d->dp_end_sn[0] + 1) * pn - 1;
    dp_alloc_row(d, 0, 0, d->dp_end_sn[0]);
    {
        int _end_sn = MIN2(d->dp_end_sn[0] + 1, dp_sn - 1);
        for (j = 0; j < (_end_sn + 1) * pn; ++j) AT(d->H, 0, j) = AT(d->E1, 0, j) = AT(d->E2, 0, j) = inf_min;
        AT(d->H, 0, 0) = 0; AT(d->E1, 0, 0) = -oe1; AT(d->E2, 0, 0) = affine ? inf_min : -oe2; AT(d->F1, 0, 0) = AT(d->F2, 0, 0) = inf_min;
        if (linear) for (j = 1; j <= d->dp_end[0]; ++j) { AT(d->H, 0, j) = Wv(d, -e1 * j); AT(d->F1, 0, j) = AT(d->F2, 0, j) = inf_min; }
        else for (j = 1; j <= d->dp_end[0]; ++j) {
            AT(d->F1, 0, j) = Wv(d, -o1 - e1 * j); AT(d->F2, 0, j) = affine ? inf_min : Wv(d, -o2 - e2 * j);
            AT(d->H, 0, j) = MAX2(AT(d->F1, 0, j), AT(d->F2, 0, j));
        }
        if (cells) *cells += (int64_t)(d->dp_end_sn[0] + 1) * pn;
    }
    /* rows in topological order (:1083-1099) */
    int dp_i;
    for (dp_i = 1; dp_i < gn - 1; ++dp_i) {
        int node_id = g->index_to_node_id[beg_index + dp_i];
        const int32_t *q = qp + (size_t)g->node[node_id].base * dp_sn * pn;
        int beg = AD_BEGIN(node_id), end = AD_END(node_id);
        int beg_sn = beg / pn, end_sn, min_pre_beg_sn = INT_MAX, max_pre_end_sn = -1, sn;
        for (i = 0; i < pre_n[dp_i]; ++i) {
            int pre = pre_index[dp_i][i];
            if (min_pre_beg_sn > d->dp_beg_sn[pre]) min_pre_beg_sn = d->dp_beg_sn[pre];
            if (max_pre_end_sn < d->dp_end_sn[pre]) max_pre_end_sn = d->dp_end_sn[pre];
        }
        if (beg_sn < min_pre_beg_sn) beg_sn = min_pre_beg_sn;
        d->dp_beg_sn[dp_i] = beg_sn; d->dp_beg[dp_i] = beg_sn * pn;
        end_sn = d->dp_end_sn[dp_i] = end / pn; d->dp_end[dp_i] = (end_sn + 1) * pn - 1;
        if (beg_sn > end_sn) { fprintf(stderr, "[tho] empty band at row %d\n", dp_i); exit(1); }
        dp_alloc_row(d, dp_i, beg_sn, end_sn);
        if (cells) *cells += (int64_t)(end_sn - beg_sn + 1) * pn;
        int32_t *rH = &AT(d->H, dp_i, beg_sn * pn) - (size_t)beg_sn * pn; /* column-indexed views */
        int32_t *rE1 = &AT(d->E1, dp_i, beg_sn * pn) - (size_t)beg_sn * pn, *rE2 = &AT(d->E2, dp_i, beg_sn * pn) - (size_t)beg_sn * pn;
        int32_t *rF1 = &AT(d->F1, dp_i, beg_sn * pn) - (size_t)beg_sn * pn, *rF2 = &AT(d->F2, dp_i, beg_sn * pn) - (size_t)beg_sn * pn;
        if (linear) { /* simd_abpoa_lg_dp (:649-736) */
            for (i = 0; i < pre_n[dp_i]; ++i) {
                int pre = pre_index[dp_i][i], _beg_sn, _end_sn; int32_t first;
                int pre_beg_sn = d->dp_beg_sn[pre], pre_end = d->dp_end[pre];
                if (pre_beg_sn < beg_sn) { _beg_sn = beg_sn; first = stored(d, pre, beg_sn - 1) ? AT(d->H, pre, (beg_sn - 1) * pn + pn - 1) : inf_min; }
                else { _beg_sn = pre_beg_sn; first = inf_min; }
                _end_sn = MIN3((pre_end + 1) / pn, end_sn, dp_sn - 1);
                if (i == 0) {
                    for (sn = beg_sn; sn < _beg_sn && sn <= end_sn + 1; ++sn) for (l = 0; l < pn; ++l) rH[sn * pn + l] = inf_min;
                    for (sn = MAX2(_end_sn + 1, beg_sn); sn <= MIN2(end_sn + 1, dp_sn - 1); ++sn) for (l = 0; l < pn; ++l) rH[sn * pn + l] = inf_min;
                }
                for (sn = _beg_sn; sn <= _end_sn; ++sn) {
                    for (l = 0; l < pn; ++l) {
                        const int32_t mm = Wv(d, (int64_t)(l ? AT(d->H, pre, sn * pn + l - 1) : first) + q[sn * pn + l]);
                        const int32_t ee = Wv(d, (int64_t)AT(d->H, pre, sn * pn + l) - e1);
                        const int32_t v = MAX2(mm, ee);
                        if (i == 0 || v > rH[sn * pn + l]) rH[sn * pn + l] = v;
                    }
                    first = AT(d->H, pre, sn * pn + pn - 1);
                }
            }
            {
                int32_t first = rH[beg_sn * pn]; /* lane 0 of the first vector, the other lanes hold inf_min (:719) */
                for (sn = beg_sn; sn <= end_sn; ++sn) {
                    int set_num; int32_t *h = rH + sn * pn;
                    if (sn < min_pre_beg_sn) { fprintf(stderr, "[tho] sn_i < min_pre_beg_sn\n"); exit(1); }
                    else if (sn > max_pre_end_sn) set_num = sn == max_pre_end_sn + 1 ? 1 : 0;
                    else set_num = pn;
                    if (first > h[0]) h[0] = first;
                    set_F(d, h, set_num, e1);
                    first = Wv(d, (int64_t)h[pn - 1] - e1);
                }
                for (j = beg_sn * pn; j < (end_sn + 1) * pn; ++j) rE1[j] = rE2[j] = rF1[j] = rF2[j] = inf_min;
            }
        } else {
        for (i = 0; i < pre_n[dp_i]; ++i) { /* M and E from every predecessor (:862-915) */
            int pre = pre_index[dp_i][i], _beg_sn, _end_sn; int32_t first;
            int pre_beg_sn = d->dp_beg_sn[pre], pre_end_sn = d->dp_end_sn[pre], pre_end = d->dp_end[pre];
            if (pre_beg_sn < beg_sn) { _beg_sn = beg_sn; first = stored(d, pre, beg_sn - 1) ? AT(d->H, pre, (beg_sn - 1) * pn + pn - 1) : inf_min; }
            else { _beg_sn = pre_beg_sn; first = inf_min; }
            _end_sn = MIN3((pre_end + 1) / pn, end_sn, dp_sn - 1);
            if (i == 0) {
                for (sn = beg_sn; sn < _beg_sn && sn <= end_sn + 1; ++sn) for (l = 0; l < pn; ++l) rH[sn * pn + l] = inf_min;
                for (sn = MAX2(_end_sn + 1, beg_sn); sn <= MIN2(end_sn + 1, dp_sn - 1); ++sn) for (l = 0; l < pn; ++l) rH[sn * pn + l] = inf_min;
            }
            for (sn = _beg_sn; sn <= _end_sn; ++sn) {
                for (l = 0; l < pn; ++l) {
                    int32_t v = l ? AT(d->H, pre, sn * pn + l - 1) : first;
                    if (i == 0 || v > rH[sn * pn + l]) rH[sn * pn + l] = v;
                }
                first = AT(d->H, pre, sn * pn + pn - 1);
            }
            _end_sn = MIN2(pre_end_sn, end_sn);
            if (i == 0) {
                for (sn = beg_sn; sn < _beg_sn && sn <= end_sn; ++sn) for (l = 0; l < pn; ++l) rE1[sn * pn + l] = rE2[sn * pn + l] = inf_min;
                for (sn = MAX2(_end_sn + 1, beg_sn); sn <= end_sn; ++sn) for (l = 0; l < pn; ++l) rE1[sn * pn + l] = rE2[sn * pn + l] = inf_min;
            }
            for (sn = _beg_sn; sn <= _end_sn; ++sn)
                for (l = 0; l < pn; ++l) {
                    int32_t a = AT(d->E1, pre, sn * pn + l), b = AT(d->E2, pre, sn * pn + l);
                    if (i == 0 || a > rE1[sn * pn + l]) rE1[sn * pn + l] = a;
                    if (i == 0 || b > rE2[sn * pn + l]) rE2[sn * pn + l] = b;
                }
        }
        for (j = beg_sn * pn; j < (end_sn + 1) * pn; ++j) rH[j] = Wv(d, (int64_t)rH[j] + q[j]);
        int32_t first = rH[beg_sn * pn], first2 = first; /* the row's first cell feeds its own F (:924) */
        for (sn = beg_sn; sn <= end_sn; ++sn) {
            int set_num;
            int32_t *h = rH + sn * pn, *x1 = rE1 + sn * pn, *x2 = rE2 + sn * pn, *f1 = rF1 + sn * pn, *f2 = rF2 + sn * pn;
            if (sn < min_pre_beg_sn) { fprintf(stderr, "[tho] sn_i < min_pre_beg_sn\n"); exit(1); }
            else if (sn > max_pre_end_sn) set_num = sn == max_pre_end_sn + 1 ? 2 : 1;
            else set_num = pn;
            if (affine) { /* simd_abpoa_ag_dp: F from M, E masked where F won */
                for (l = 0; l < pn; ++l) f1[l] = Wv(d, (int64_t)(l ? h[l - 1] : first) - oe1);
                set_F(d, f1, set_num, e1);
                first = MAX2(h[pn - 1], Wv(d, (int64_t)f1[pn - 1] + o1));
                for (l = 0; l < pn; ++l) {
                    const int32_t tmp = MAX2(h[l], x1[l]);
                    h[l] = MAX2(tmp, f1[l]);
                    x1[l] = h[l] == tmp ? MAX2(Wv(d, (int64_t)x1[l] - e1), Wv(d, (int64_t)h[l] - oe1)) : inf_min;
                    x2[l] = inf_min; f2[l] = inf_min;
                }
                continue;
            }
            for (l = 0; l < pn; ++l) h[l] = MAX3(h[l], x1[l], x2[l]);
            for (l = 0; l < pn; ++l) {
                f1[l] = Wv(d, (int64_t)(l ? h[l - 1] : first) - oe1);
                f2[l] = Wv(d, (int64_t)(l ? h[l - 1] : first2) - oe2);
            }
            set_F(d, f1, set_num, e1); set_F(d, f2, set_num, e2);
            first = MAX2(h[pn - 1], Wv(d, (int64_t)f1[pn - 1] + o1));
            first2 = MAX2(h[pn - 1], Wv(d, (int64_t)f2[pn - 1] + o2));
            for (l = 0; l < pn; ++l) {
                h[l] = MAX3(h[l], f1[l], f2[l]);
                x1[l] = MAX2(Wv(d, (int64_t)x1[l] - e1), Wv(d, (int64_t)h[l] - oe1));
                x2[l] = MAX2(Wv(d, (int64_t)x2[l] - e2), Wv(d, (int64_t)h[l] - oe2));
            }
        }
        } /* !linear */
        /* row arg-max with the lane-ordered tie-break (:991-1005), then widen out-neighbours (:1007-1015) */
        {
            int32_t a[64], b[64]; int max = inf_min, max_i = -1;
            for (l = 0; l < pn; ++l) { a[l] = rH[end_sn * pn + l]; b[l] = qi[end_sn * pn + l]; }
            if (end_sn == qlen / pn) for (l = 0; l < pn; ++l) if (0 > b[l]) a[l] = inf_min;
            for (sn = beg_sn; sn < end_sn; ++sn)
                for (l = 0; l < pn; ++l) if (rH[sn * pn + l] > a[l]) { a[l] = rH[sn * pn + l]; b[l] = qi[sn * pn + l]; }
            for (l = 0; l < pn; ++l) if (a[l] > max) { max = a[l]; max_i = b[l]; }
            int out_i = max_i + 1;
            for (i = 0; i < g->node[node_id].out_n; ++i) {
                int o = g->node[node_id].out_id[i];
                if (out_i > g->max_pos_right[o]) g->max_pos_right[o] = out_i;
                if (out_i < g->max_pos_left[o]) g->max_pos_left[o] = out_i;
            }
        }
    }
    /* best end cell (:976-989) */
    int best_score = inf_min, best_i = 0, best_j = 0;
    for (i = 0; i < g->node[1].in_n; ++i) {
        int in_dp_i = g->node_id_to_index[g->node[1].in_id[i]] - beg_index;
        int end = qlen > d->dp_end[in_dp_i] ? d->dp_end[in_dp_i] : qlen;
        int32_t s = AT(d->H, in_dp_i, end);
        if (s > best_score) { best_score = s; best_i = in_dp_i; best_j = end; }
    }
    /* backtrack by value comparison (:248-377) */
    enum { M_OP = 1, E1_OP = 2, E2_OP = 4, E_OP = 6, F1_OP = 8, F2_OP = 16, F_OP = 24, ALL_OP = 31 };
    pcig_t *cig = NULL; int n_c = 0, m_c = 0, cur_op = ALL_OP, hit, id;
    i = best_i; j = best_j; id = g->index_to_node_id[i + beg_index];
    if (best_j < qlen) { pcig_t c = {OP_I, -1, qlen - 1, qlen - j}; PUSH(cig, n_c, m_c, pcig_t, c); }
#define INBAND(r, c) ((c) >= d->dp_beg[r] && (c) <= d->dp_end[r])
#define PUSH_I1(qpos) do { if (n_c && cig[n_c - 1].op == OP_I) cig[n_c - 1].len += 1; else { pcig_t c_ = {OP_I, -1, (qpos), 1}; PUSH(cig, n_c, m_c, pcig_t, c_); } } while (0)
    while (linear && i > 0 && j > 0) { /* simd_abpoa_lg_backtrack (:108-158): match, then deletion, else insertion */
        int s = mat[m * g->node[id].base + query[j - 1]];
        int32_t hij = AT(d->H, i, j);
        hit = 0;
        for (k = 0; k < pre_n[i]; ++k) {
            int pre = pre_index[i][k];
            if (!INBAND(pre, j - 1)) continue;
            if ((int64_t)AT(d->H, pre, j - 1) + s == (int64_t)hij) {
                pcig_t c = {OP_M, id, j - 1, 1}; PUSH(cig, n_c, m_c, pcig_t, c);
                hit = 1; i = pre; --j; id = g->index_to_node_id[i + beg_index];
                break;
            }
        }
        if (hit == 0) for (k = 0; k < pre_n[i]; ++k) {
            int pre = pre_index[i][k];
            if (!INBAND(pre, j)) continue;
            if ((int64_t)AT(d->H, pre, j) - e1 == (int64_t)hij) {
                pcig_t c = {OP_D, id, j - 1, 1}; PUSH(cig, n_c, m_c, pcig_t, c);
                hit = 1; i = pre; id = g->index_to_node_id[i + beg_index];
                break;
            }
        }
        if (hit == 0) { PUSH_I1(j - 1); --j; }
    }
    while (i > 0 && j > 0) {
        int s = mat[m * g->node[id].base + query[j - 1]];
        int32_t hij = AT(d->H, i, j);
        hit = 0;
        if (cur_op & M_OP) {
            for (k = 0; k < pre_n[i]; ++k) {
                int pre = pre_index[i][k];
                if (!INBAND(pre, j - 1)) continue;
                /* the reference compares in promoted int arithmetic (score_t + int): no 16-bit wrap here */
                if ((int64_t)AT(d->H, pre, j - 1) + s == (int64_t)hij) {
                    pcig_t c = {OP_M, id, j - 1, 1}; PUSH(cig, n_c, m_c, pcig_t, c);
                    cur_op = ALL_OP; hit = 1; i = pre; --j; id = g->index_to_node_id[i + beg_index];
                    break;
                }
            }
        }
        if (hit == 0 && (cur_op & E_OP)) {
            for (k = 0; k < pre_n[i]; ++k) {
                int pre = pre_index[i][k];
                if (!INBAND(pre, j)) continue;
                if (cur_op & E1_OP) {
                    int32_t pe1 = AT(d->E1, pre, j);
                    int ok = (cur_op & M_OP) ? (hij == pe1) : ((int64_t)AT(d->E1, i, j) == (int64_t)pe1 - e1);
                    if (ok) {
                        if ((int64_t)AT(d->H, pre, j) - oe1 == (int64_t)pe1) cur_op = M_OP | F_OP; else cur_op = E1_OP;
                        pcig_t c = {OP_D, id, j - 1, 1}; PUSH(cig, n_c, m_c, pcig_t, c);
                        hit = 1; i = pre; id = g->index_to_node_id[i + beg_index];
                        break;
                    }
                }
                if (cur_op & E2_OP) {
                    int32_t pe2 = AT(d->E2, pre, j);
                    int ok = (cur_op & M_OP) ? (hij == pe2) : ((int64_t)AT(d->E2, i, j) == (int64_t)pe2 - e2);
                    if (ok) {
                        if ((int64_t)AT(d->H, pre, j) - oe2 == (int64_t)pe2) cur_op = M_OP | F_OP; else cur_op = E2_OP;
                        pcig_t c = {OP_D, id, j - 1, 1}; PUSH(cig, n_c, m_c, pcig_t, c);
                        hit = 1; i = pre; id = g->index_to_node_id[i + beg_index];
                        break;
                    }
                }
            }
        }
        if (hit == 0 && (cur_op & F_OP)) {
            if (cur_op & F1_OP) {
                int32_t f1 = AT(d->F1, i, j);
                if (!(cur_op & M_OP) || hij == f1) {
                    if ((int64_t)AT(d->H, i, j - 1) - oe1 == (int64_t)f1) { cur_op = M_OP | E_OP; hit = 1; }
                    else if ((int64_t)AT(d->F1, i, j - 1) - e1 == (int64_t)f1) { cur_op = F1_OP; hit = 1; }
                    else { fprintf(stderr, "[tho] Error in cg_backtrack (F1)\n"); exit(1); }
                }
            }
            if (hit == 0 && (cur_op & F2_OP)) {
                int32_t f2 = AT(d->F2, i, j);
                if (!(cur_op & M_OP) || hij == f2) {
                    if ((int64_t)AT(d->H, i, j - 1) - oe2 == (int64_t)f2) { cur_op = M_OP | E_OP; hit = 1; }
                    else if ((int64_t)AT(d->F2, i, j - 1) - e2 == (int64_t)f2) { cur_op = F2_OP; hit = 1; }
                    else { fprintf(stderr, "[tho] Error in cg_backtrack (F2)\n"); exit(1); }
                }
            }
            PUSH_I1(j - 1); --j;
            hit = 1;
        }
        if (hit == 0) { fprintf(stderr, "[tho] Error in cg_backtrack (5) i=%d j=%d op=%d\n", i, j, cur_op); exit(1); }
    }
    if (j > 0) { /* leading insertion merges with a preceding (later-in-query) I run, abpoa_align.h:54-73 */
        if (n_c && cig[n_c - 1].op == OP_I) cig[n_c - 1].len += j;
        else { pcig_t c = {OP_I, -1, j - 1, j}; PUSH(cig, n_c, m_c, pcig_t, c); }
    }
    for (i = 0; i < n_c >> 1; ++i) { pcig_t t = cig[i]; cig[i] = cig[n_c - 1 - i]; cig[n_c - 1 - i] = t; }
    *cig_ = cig; *n_cig_ = n_c;
    for (i = 0; i < gn; ++i) free(pre_index[i]);
    free(pre_index); free(pre_n); free(qp); free(qi);
    free(d->dp_beg); free(d->dp_end); free(d->dp_beg_sn); free(d->dp_end_sn); free(d->row_off);
    free(d->H); free(d->E1); free(d->E2); free(d->F1); free(d->F2);
    return best_score;
}

/* heaviest-column consensus: abpoa_graph.c:279-359 (DFS msa rank), :604-648 */
static int g_consensus(pgraph_t *g, int n_seq, uint8_t *cons, int *cov) {
    int n = g->node_n, i, j, k;
    g_aux_reserve(g);
    int *deg = (int *)malloc(sizeof(int) * n), *stk = (int *)malloc(sizeof(int) * (n + 1)), sp = 0, msa_rank = 0;
    for (i = 0; i < n; ++i) deg[i] = g->node[i].in_n;
    stk[sp++] = 0; g->msa_rank[0] = -1;
    while (sp > 0) {
        int cur = stk[--sp];
        if (g->msa_rank[cur] < 0) {
            g->msa_rank[cur] = msa_rank;
            for (i = 0; i < g->node[cur].aln_n; ++i) g->msa_rank[g->node[cur].aln_id[i]] = msa_rank;
            msa_rank++;
        }
        if (cur == 1) break;
        for (i = 0; i < g->node[cur].out_n; ++i) {
            int out = g->node[cur].out_id[i], ok = 1;
            if (--deg[out] == 0) {
                for (j = 0; j < g->node[out].aln_n; ++j) if (deg[g->node[out].aln_id[j]] != 0) { ok = 0; break; }
                if (!ok) continue;
                stk[sp++] = out; g->msa_rank[out] = -1;
                for (j = 0; j < g->node[out].aln_n; ++j) { int a = g->node[out].aln_id[j]; stk[sp++] = a; g->msa_rank[a] = -1; }
            }
        }
    }
    int msa_l = g->msa_rank[1] - 1;
    int *rc_weight = (int *)calloc((size_t)(msa_l > 0 ? msa_l : 1) * 5, sizeof(int));
    int *msa_node = (int *)calloc((size_t)(msa_l > 0 ? msa_l : 1) * 5, sizeof(int));
    for (i = 0; i < msa_l; ++i) rc_weight[i * 5 + 4] = n_seq;
    for (i = 2; i < n; ++i) { /* abpoa_set_row_column_weight */
        int rank = g->msa_rank[i];
        for (k = 0; k < g->node[i].aln_n; ++k) rank = MAX2(rank, g->msa_rank[g->node[i].aln_id[k]]);
        for (k = 0; k < g->node[i].read_ids_n; ++k) rc_weight[(rank - 1) * 5 + g->node[i].base] += __builtin_popcountll(g->node[i].read_ids[k]);
        rc_weight[(rank - 1) * 5 + 4] -= rc_weight[(rank - 1) * 5 + g->node[i].base];
        msa_node[(rank - 1) * 5 + g->node[i].base] = i;
    }
    int last_id = 0, cons_i = 0;
    for (i = 0; i < msa_l; ++i) { /* abpoa_heaviest_column_consensus */
        int max_w = 0, max_base = 5, gap_w = n_seq, w;
        for (j = 0; j < 4; ++j) { w = rc_weight[i * 5 + j]; if (w > max_w) { max_base = j; max_w = w; } gap_w -= w; }
        if (max_w >= gap_w) {
            int cur = msa_node[i * 5 + max_base];
            g->node[last_id].max_out_id = cur; last_id = cur;
            cov[cons_i++] = max_w;
        }
    }
    g->node[last_id].max_out_id = 1;
    int id = g->node[0].max_out_id, l = 0; /* abpoa_store_consensus, :467-478 */
    while (id != 1) { cons[l++] = g->node[id].base; id = g->node[id].max_out_id; }
    free(deg); free(stk); free(rc_weight); free(msa_node);
    return l;
}

/* abpoa_msa -> abpoa_poa (abpoa_align.c:293-328, 357-411) with TideHunter's parameters
 * (src/abpoa_cons.c:12-28): plain in-order POA, no seeding, no strand ambiguity. */
int tho_abpoa_cons(const tho_para_t *p, int n_seqs, const uint8_t *const *seqs, const int *lens, uint8_t *cons, int *cov, int64_t *poa_cells) {
    pgraph_t *g = g_init();
    int i, cons_l;
    for (i = 0; i < n_seqs; ++i) {
        pcig_t *cig = NULL; int n_cig = 0;
        if (g->node_n > 2) {
            g_topological_sort(g);
            poa_align(g, p, seqs[i], lens[i], &cig, &n_cig, poa_cells);
        }
        g_add_alignment(g, seqs[i], lens[i], cig, n_cig, i, n_seqs);
        free(cig);
    }
    if (g->node_n <= 2) { g_free(g); return 0; }
    cons_l = g_consensus(g, n_seqs, cons, cov);
    g_free(g);
    return cons_l;
}
