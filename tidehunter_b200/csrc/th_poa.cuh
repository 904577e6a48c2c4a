// th_poa.cuh -- adaptive-banded partial-order alignment + heaviest-column consensus, one warp per task.
//
// Replaces (reference, /root/reference): src/abpoa_cons.c:30-120 (abpoa_gen_cons) and the abPOA calls
// below it: abPOA/src/abpoa_align.c:293-411 (abpoa_msa / abpoa_poa), abPOA/src/simd_abpoa_align.c
// (convex-gap banded DP :835-958, row arg-max :991-1015, best cell :976-989, backtrack :248-377),
// abPOA/src/abpoa_graph.c (:197-238 max_remain, :1020-1124 node/edge/aligned, :1218-1288 add alignment,
// :279-359 + :604-648 heaviest-column consensus).
//
// What is kept bit-exact and why it is enough:
//  * DP values are 16-bit wrapping integers exactly as the reference's AVX2 int16 path; two adjacent
//    columns are packed in one register (s16x2) and updated with DPX/video SIMD ops.
//  * Band edges are rounded to whole emulated SIMD vectors of `pn` lanes (pn = 16), and the row
//    arg-max that steers the adaptive band uses the reference's lane-ordered tie-break.
//  * Rows are processed in a topological order that keeps aligned-node groups contiguous; it is NOT
//    the reference's BFS order.  The alignment does not depend on which topological order is used:
//    a row's band and values depend only on its predecessors' rows, the backtrack walks in_id order,
//    and max_remain is a function of graph structure only.  The order is maintained incrementally
//    (new nodes are merged in right before the next existing node of the alignment path), which
//    replaces the per-sequence BFS (abpoa_graph.c:150-195) by a parallel merge.
//  * read-id bitsets are not stored: popcount(read_ids) of a node equals the sum of its out-edge
//    weights, because every sequence adds weight 1 to exactly one out-edge of each node it visits.
//  * The consensus DFS (msa rank) and column vote are literal.
//
// HBM layout (per warp "slab"): graph arrays (SoA, int32), edge pool with per-node in/out linked lists
// in insertion order, two order buffers, per-row band metadata, query profile (5 x int16 rows), cigar,
// and the DP arena holding, per row, H|E1|E2|F1|F2 as int16 over the row's band only.
#pragma once
#include "th_common.cuh"

struct PoaTask {
    int64_t seq_off;     // offset of the read in the nt4 buffer
    int32_t read, unit_off, n_seqs;
    int32_t ncap;        // node capacity = sum(unit len) + 2
    int32_t qmax;        // longest unit
    int32_t cons_off;    // output offset (capacity = ncap - 2)
};

#define POA_WARPS 4
#define POA_MAXPRE 32
#define POA_NEGP 0x80008000u

struct PoaWs {
    int32_t *out_head, *out_tail, *in_head, *in_tail, *aln_n, *aln, *n2i, *remain, *mpl, *mpr, *hs;
    int32_t *e_to, *e_from, *e_w, *e_no, *e_ni;
    int32_t *ord, *ord2, *ev_anchor, *ev_node, *hi_idx;
    uint32_t *row_off; int32_t *row_bsn, *row_esn;
    int16_t *qp; uint32_t *cigar; uint8_t *base;
    int16_t *arena; uint32_t arena_cap;
    int32_t qp_stride;
};

__host__ __device__ inline size_t poa_fixed_bytes(int ncap, int qmax, int nseq) {
    size_t ecap = (size_t)ncap + nseq + 2;
    size_t b = 0;
    b += (size_t)ncap * 4 * (11 + 4 - 1 + 1);        // 11 node arrays (aln counts as 4) = 15 x int32... see carve
    b += ecap * 4 * 5;
    b += (size_t)ncap * 4 * 3;                         // ord, ord2, hi_idx
    b += (size_t)(qmax + 2) * 4 * 2;                   // events
    b += (size_t)ncap * 4 * 3;                         // row meta
    b += (size_t)5 * (qmax + 1 + 128) * 2;             // profile
    b += (size_t)(qmax + ncap + 8) * 4;                // cigar
    b += (size_t)ncap + 64;                            // base
    return (b + 4095) & ~(size_t)4095;
}

__device__ inline void poa_carve(PoaWs &w, uint8_t *slab, size_t slab_bytes, int ncap, int qmax, int nseq) {
    size_t ecap = (size_t)ncap + nseq + 2;
    int32_t *p = reinterpret_cast<int32_t *>(slab);
    w.out_head = p; p += ncap; w.out_tail = p; p += ncap; w.in_head = p; p += ncap; w.in_tail = p; p += ncap;
    w.aln_n = p; p += ncap; w.aln = p; p += 4 * (size_t)ncap; w.n2i = p; p += ncap; w.remain = p; p += ncap;
    w.mpl = p; p += ncap; w.mpr = p; p += ncap; w.hs = p; p += ncap;
    w.e_to = p; p += ecap; w.e_from = p; p += ecap; w.e_w = p; p += ecap; w.e_no = p; p += ecap; w.e_ni = p; p += ecap;
    w.ord = p; p += ncap; w.ord2 = p; p += ncap; w.hi_idx = p; p += ncap;
    w.ev_anchor = p; p += qmax + 2; w.ev_node = p; p += qmax + 2;
    w.row_off = reinterpret_cast<uint32_t *>(p); p += ncap; w.row_bsn = p; p += ncap; w.row_esn = p; p += ncap;
    w.qp_stride = (qmax + 1 + 128) & ~1;
    w.qp = reinterpret_cast<int16_t *>(p); p += ((size_t)5 * w.qp_stride * 2 + 3) / 4;
    w.cigar = reinterpret_cast<uint32_t *>(p); p += qmax + ncap + 8;
    w.base = reinterpret_cast<uint8_t *>(p);
    size_t fixed = poa_fixed_bytes(ncap, qmax, nseq);
    w.arena = reinterpret_cast<int16_t *>(slab + fixed);
    size_t ab = slab_bytes > fixed ? slab_bytes - fixed : 0;
    size_t ne = ab / 2; if (ne > 0xfffffff0ull) ne = 0xfffffff0ull;
    w.arena_cap = (uint32_t)ne;
}

__device__ __forceinline__ uint32_t ld32(const int16_t *p) { return *reinterpret_cast<const uint32_t *>(p); }
__device__ __forceinline__ void st32(int16_t *p, uint32_t v) { *reinterpret_cast<uint32_t *>(p) = v; }

// graph edits (lane 0 only) ------------------------------------------------------------------
__device__ inline void g_add_edge(PoaWs &w, int &edge_n, int from, int to, bool check) {
    if (check) {
        for (int e = w.out_head[from]; e >= 0; e = w.e_no[e])
            if (w.e_to[e] == to) { w.e_w[e] += 1; return; }
    }
    int e = edge_n++;
    w.e_to[e] = to; w.e_from[e] = from; w.e_w[e] = 1; w.e_no[e] = -1; w.e_ni[e] = -1;
    if (w.out_tail[from] < 0) w.out_head[from] = e; else w.e_no[w.out_tail[from]] = e;
    w.out_tail[from] = e;
    if (w.in_tail[to] < 0) w.in_head[to] = e; else w.e_ni[w.in_tail[to]] = e;
    w.in_tail[to] = e;
}
__device__ inline int g_new_node(PoaWs &w, int &node_n, uint8_t b) {
    int v = node_n++;
    w.base[v] = b; w.out_head[v] = w.out_tail[v] = w.in_head[v] = w.in_tail[v] = -1; w.aln_n[v] = 0;
    return v;
}

// one warp aligns sequence `query` to the graph and merges it in.  Returns an error code.
__device__ int poa_add_sequence(PoaWs &w, const DevParams &P, const uint8_t *query, int qlen, int &node_n, int &edge_n,
                                int *s_pre /* POA_MAXPRE*4 ints of shared scratch for this warp */,
                                unsigned long long &cells, unsigned long long &rows) {
    const int lane = lane_id();
    const int n = node_n, pn = P.pn;
    const int o1 = P.o1, e1 = P.e1, o2 = P.o2, e2 = P.e2, oe1 = o1 + e1, oe2 = o2 + e2;
    const int mis = P.mismatch > 0 ? P.mismatch : -P.mismatch, mat = P.match < 0 ? -P.match : P.match;
    { // int16 path only (simd_abpoa_align.c:1610-1621)
        int len = qlen > n ? qlen : n;
        int max_score = max(qlen * mat, len * e1 + o1);
        if (max_score > 32767 - mis - oe1 - oe2) return TH_ERR_LEN;
    }
    const int inf_min = max(max(-32768 + mis, -32768 + oe1), -32768 + oe2) + 31 * max(e1, e2);
    const uint32_t INFP = pk(inf_min, inf_min);
    const int wband = 10 + (int)(0.01f * (float)qlen); // wb + (int)(wf*qlen), float (simd_abpoa_align.c:393)
    // ---- order index, heaviest successor, max_remain --------------------------------------
    for (int i = lane; i < n; i += 32) w.n2i[w.ord[i]] = i;
    __syncwarp();
    for (int v = lane; v < n; v += 32) { // first out-edge with maximum weight (abpoa_graph.c:216-226)
        int mw = -1, mt = -1;
        for (int e = w.out_head[v]; e >= 0; e = w.e_no[e]) if (w.e_w[e] > mw) { mw = w.e_w[e]; mt = w.e_to[e]; }
        w.hs[v] = mt;
        w.mpl[v] = n; w.mpr[v] = 0; // reset of abpoa_topological_sort (abpoa_graph.c:267-272)
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) { int h = w.hs[w.ord[i]]; w.hi_idx[i] = h >= 0 ? w.n2i[h] : 0x7fffffff; }
    __syncwarp();
    // remain by index, 32 indices at a time from the sink backwards; chains inside a chunk are
    // collapsed by pointer jumping on shuffles (remain[v] = remain[heaviest successor] + 1)
    int32_t *ri = w.row_bsn; // temporary: remain by index (row_bsn is rewritten by the DP below)
    for (int cb = ((n - 1) / 32) * 32; cb >= 0; cb -= 32) {
        const int idx = cb + lane;
        int ptr = 0x7fffffff, dist = 0;
        if (idx < n) { ptr = w.hi_idx[idx]; dist = 1; if (idx == n - 1) { ptr = 0x7fffffff; dist = -1; } }
#pragma unroll
        for (int rnd = 0; rnd < 5; ++rnd) {
            const bool inside = ptr < cb + 32;
            const int tl = inside ? ptr - cb : 0;
            const int pd = __shfl_sync(TH_FULL, dist, tl), pp = __shfl_sync(TH_FULL, ptr, tl);
            if (inside) { dist += pd; ptr = pp; }
        }
        if (idx < n) {
            const int val = (ptr == 0x7fffffff ? 0 : ri[ptr]) + dist;
            ri[idx] = val;
            w.remain[w.ord[idx]] = val;
        }
        __syncwarp();
    }
    // ---- query profile (simd_abpoa_align.c:438-446); N row/column score 0 -----------------
    const int prof_w = ((qlen / pn + 1) * pn + 64 + 1) & ~1;
    for (int j = lane; j < prof_w; j += 32) {
        const int qc = (j >= 1 && j <= qlen) ? min((int)query[j - 1], 4) : -1;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            int s = 0;
            if (qc >= 0 && qc < 4 && b < 4) s = (qc == b) ? mat : -mis;
            w.qp[b * w.qp_stride + j] = (int16_t)s;
        }
    }
    // ---- first row (simd_abpoa_align.c:538-555, 591-610) ------------------------------------
    if (lane == 0) {
        w.mpl[0] = w.mpr[0] = 0;
        for (int e = w.out_head[0]; e >= 0; e = w.e_no[e]) { w.mpl[w.e_to[e]] = 1; w.mpr[w.e_to[e]] = 1; }
    }
    __syncwarp();
    uint32_t used = 0;
    {
        const int r = w.remain[0];
        const int end = min(qlen, max(w.mpr[0], qlen - r) + wband);
        const int esn = end / pn, width = (esn + 1) * pn;
        if ((uint64_t)used + 5ull * width > w.arena_cap) return TH_ERR_ARENA;
        if (lane == 0) { w.row_off[0] = used; w.row_bsn[0] = 0; w.row_esn[0] = esn; }
        int16_t *H = w.arena + used, *E1 = H + width, *E2 = E1 + width, *F1 = E2 + width, *F2 = F1 + width;
        for (int j = lane; j < width; j += 32) {
            int f1 = -o1 - e1 * j, f2 = -o2 - e2 * j;
            H[j] = (int16_t)(j == 0 ? 0 : max((int)(int16_t)f1, (int)(int16_t)f2));
            E1[j] = (int16_t)(j == 0 ? -oe1 : inf_min); E2[j] = (int16_t)(j == 0 ? -oe2 : inf_min);
            F1[j] = (int16_t)(j == 0 ? inf_min : f1); F2[j] = (int16_t)(j == 0 ? inf_min : f2);
        }
        used += 5u * width; cells += width; rows += 1;
    }
    __syncwarp();
    // ---- rows in topological order ----------------------------------------------------------
    const uint32_t OE1P = pk(oe1, oe1), OE2P = pk(oe2, oe2), E1P = pk(e1, e1), E2P = pk(e2, e2), E12P = pk(e1, e2);
    const int lam_bits = pn - 1;
    for (int i = 1; i < n - 1; ++i) {
        const int v = w.ord[i];
        const int r = w.remain[v];
        const int beg0 = max(0, min(w.mpl[v], qlen - r) - wband), end0 = min(qlen, max(w.mpr[v], qlen - r) + wband);
        // predecessors (in_id order); every lane walks the same list
        int np = 0, min_pre_bsn = 0x7fffffff;
        for (int e = w.in_head[v]; e >= 0; e = w.e_ni[e]) {
            const int pi = w.n2i[w.e_from[e]];
            const int pb = w.row_bsn[pi];
            min_pre_bsn = min(min_pre_bsn, pb);
            if (np < POA_MAXPRE && lane == 0) { s_pre[np * 4] = (int)w.row_off[pi]; s_pre[np * 4 + 1] = pb * pn; s_pre[np * 4 + 2] = (w.row_esn[pi] + 1) * pn - 1; s_pre[np * 4 + 3] = pi; }
            ++np;
        }
        if (np > POA_MAXPRE) return TH_ERR_CAP;
        __syncwarp();
        const int bsn = max(beg0 / pn, min_pre_bsn), esn = end0 / pn;
        if (bsn > esn) return TH_ERR_BAND;
        const int beg = bsn * pn, dend = (esn + 1) * pn - 1, width = dend - beg + 1;
        if ((uint64_t)used + 5ull * width > w.arena_cap) return TH_ERR_ARENA;
        if (lane == 0) { w.row_off[i] = used; w.row_bsn[i] = bsn; w.row_esn[i] = esn; }
        int16_t *H = w.arena + used, *E1 = H + width, *E2 = E1 + width, *F1 = E2 + width, *F2 = F1 + width;
        used += 5u * width; cells += width; rows += 1;
        const int16_t *qrow = w.qp + (int)w.base[v] * w.qp_stride;
        const bool mask_tail = esn == qlen / pn;
        uint32_t best = 0, carryH = 0; int carryF1 = 0, carryF2 = 0;
        const int nchunk = (width + 63) >> 6;
        for (int ch = 0; ch < nchunk; ++ch) {
            const int j = beg + (ch << 6) + 2 * lane;
            const bool in_row = j <= dend;
            uint32_t Mx = INFP, E1x = INFP, E2x = INFP;
            for (int p = 0; p < np; ++p) {
                const int poff = s_pre[p * 4], pb = s_pre[p * 4 + 1], pe = s_pre[p * 4 + 2], pw = pe - pb + 1;
                const int16_t *Hp = w.arena + (uint32_t)poff;
                const bool inb = j >= pb && j <= pe;
                const uint32_t Xh = inb ? ld32(Hp + (j - pb)) : INFP;
                uint32_t prev = __shfl_up_sync(TH_FULL, Xh, 1);
                if (lane == 0) { const int jm = j - 1; int pv = (jm >= pb && jm <= pe) ? (int)Hp[jm - pb] : inf_min; prev = (uint32_t)(uint16_t)pv << 16; }
                Mx = __vmaxs2(Mx, __funnelshift_r(prev, Xh, 16));
                if (inb) { E1x = __vmaxs2(E1x, ld32(Hp + pw + (j - pb))); E2x = __vmaxs2(E2x, ld32(Hp + 2 * pw + (j - pb))); }
            }
            const uint32_t S = ld32(qrow + j);
            const uint32_t Ms = __vadd2(Mx, S);
            const uint32_t Hme = __vmaxs2(__vmaxs2(Ms, E1x), E2x);
            uint32_t hp = __shfl_up_sync(TH_FULL, Hme, 1);
            if (lane == 0) hp = ch == 0 ? (Ms << 16) : carryH;
            const uint32_t Hsh = __funnelshift_r(hp, Hme, 16);
            uint32_t A1 = __vsubss2(Hsh, OE1P), A2 = __vsubss2(Hsh, OE2P);
            if (lane == 0 && ch > 0) {
                A1 = pk(max(lo16(A1), carryF1 - e1), hi16(A1));
                A2 = pk(max(lo16(A2), carryF2 - e2), hi16(A2));
            }
            A1 = __vmaxs2(A1, (__vsubss2(A1, E1P) << 16) | 0x8000u);
            A2 = __vmaxs2(A2, (__vsubss2(A2, E2P) << 16) | 0x8000u);
            uint32_t TT = __byte_perm(A1, A2, 0x7632); // lo = F1 at this lane's odd column, hi = F2
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(TH_FULL, TT, d);
                o = __vsubss2(o, pk(2 * d * e1, 2 * d * e2));
                if (lane >= d) TT = __vmaxs2(TT, o);
            }
            uint32_t Pv = __shfl_up_sync(TH_FULL, TT, 1);
            if (lane == 0) Pv = POA_NEGP;
            Pv = __vsubss2(Pv, E12P); // (F1[j-1]-e1, F2[j-1]-e2)
            uint32_t Fa = __vmaxs2(A1, (Pv & 0xffffu) | 0x80000000u);
            uint32_t Fb = __vmaxs2(A2, (Pv >> 16) | 0x80000000u);
            Fa = __vmaxs2(Fa, (__vsubss2(Fa, E1P) << 16) | 0x8000u);
            Fb = __vmaxs2(Fb, (__vsubss2(Fb, E2P) << 16) | 0x8000u);
            const uint32_t Hn = __vmaxs2(Hme, __vmaxs2(Fa, Fb));
            const uint32_t E1o = __vmaxs2(__vsub2(E1x, E1P), __vsub2(Hn, OE1P));
            const uint32_t E2o = __vmaxs2(__vsub2(E2x, E2P), __vsub2(Hn, OE2P));
            carryH = __shfl_sync(TH_FULL, Hme, 31) & 0xffff0000u;
            carryF1 = hi16(__shfl_sync(TH_FULL, Fa, 31)); carryF2 = hi16(__shfl_sync(TH_FULL, Fb, 31));
            if (in_row) {
                const int c = j - beg;
                st32(H + c, Hn); st32(E1 + c, E1o); st32(E2 + c, E2o); st32(F1 + c, Fa); st32(F2 + c, Fb);
                // row arg-max key: value, then lane (j mod pn) ascending, then vector order with end_sn first
                const int vsn = j / pn, vr = vsn == esn ? 0 : vsn - bsn + 1;
                const uint32_t tail = (uint32_t)(0xfff - vr);
                if (!(mask_tail && j > qlen)) {
                    uint32_t key = ((uint32_t)(lo16(Hn) + 32768) << 16) | ((uint32_t)(lam_bits - (j & lam_bits)) << 12) | tail;
                    best = max(best, key);
                }
                if (!(mask_tail && j + 1 > qlen)) {
                    uint32_t key = ((uint32_t)(hi16(Hn) + 32768) << 16) | ((uint32_t)(lam_bits - ((j + 1) & lam_bits)) << 12) | tail;
                    best = max(best, key);
                }
            }
        }
        best = __reduce_max_sync(TH_FULL, best);
        if (lane == 0) { // simd_abpoa_max_in_row + simd_abpoa_ada_max_i
            const int val = (int)(best >> 16) - 32768;
            int max_i = -1;
            if (best != 0 && val > inf_min) {
                const int lam = lam_bits - (int)((best >> 12) & 0xf), vr = 0xfff - (int)(best & 0xfff);
                const int vsn = vr == 0 ? esn : bsn + vr - 1;
                max_i = vsn * pn + lam;
            }
            const int out_i = max_i + 1;
            for (int e = w.out_head[v]; e >= 0; e = w.e_no[e]) {
                const int o = w.e_to[e];
                if (out_i > w.mpr[o]) w.mpr[o] = out_i;
                if (out_i < w.mpl[o]) w.mpl[o] = out_i;
            }
        }
        __syncwarp();
    }
    // ---- best end cell, backtrack by value comparison (lane 0) -------------------------------
    int n_cig = 0, err = TH_OK;
    if (lane == 0) {
#define ROWP(idx) (w.arena + w.row_off[idx])
#define RBEG(idx) (w.row_bsn[idx] * pn)
#define REND(idx) ((w.row_esn[idx] + 1) * pn - 1)
#define RW(idx) ((w.row_esn[idx] - w.row_bsn[idx] + 1) * pn)
        int best_score = inf_min, bi = 0, bj = 0;
        for (int e = w.in_head[1]; e >= 0; e = w.e_ni[e]) {
            const int pi = w.n2i[w.e_from[e]];
            const int end = qlen > REND(pi) ? REND(pi) : qlen;
            const int s = ROWP(pi)[end - RBEG(pi)];
            if (s > best_score) { best_score = s; bi = pi; bj = end; }
        }
        enum { M_OP = 1, E1_OP = 2, E2_OP = 4, E_OP = 6, F1_OP = 8, F2_OP = 16, F_OP = 24, ALL_OP = 31 };
        int i = bi, j = bj, cur_op = ALL_OP;
        uint32_t *cg = w.cigar;
#define PUSH_I(len_) do { if (n_cig && (cg[n_cig - 1] & 3) == 1) cg[n_cig - 1] += (uint32_t)(len_) << 2; else cg[n_cig++] = ((uint32_t)(len_) << 2) | 1; } while (0)
        if (bj < qlen) PUSH_I(qlen - bj);
        while (i > 0 && j > 0 && err == TH_OK) {
            const int v = w.ord[i];
            const int16_t *Hi = ROWP(i); const int ib = RBEG(i), iw = RW(i);
            const int s = (query[j - 1] < 4 && w.base[v] < 4) ? (query[j - 1] == w.base[v] ? mat : -mis) : 0;
            const int hij = Hi[j - ib];
            bool hit = false;
            if (cur_op & M_OP) {
                for (int e = w.in_head[v]; e >= 0; e = w.e_ni[e]) {
                    const int pi = w.n2i[w.e_from[e]];
                    if (j - 1 < RBEG(pi) || j - 1 > REND(pi)) continue;
                    if ((int)ROWP(pi)[j - 1 - RBEG(pi)] + s == hij) {
                        cg[n_cig++] = ((uint32_t)v << 2) | 0;
                        cur_op = ALL_OP; hit = true; i = pi; --j;
                        break;
                    }
                }
            }
            if (!hit && (cur_op & E_OP)) {
                for (int e = w.in_head[v]; e >= 0 && !hit; e = w.e_ni[e]) {
                    const int pi = w.n2i[w.e_from[e]];
                    if (j < RBEG(pi) || j > REND(pi)) continue;
                    const int16_t *Hp = ROWP(pi); const int pc = j - RBEG(pi), pw = RW(pi);
                    if (cur_op & E1_OP) {
                        const int pe1 = Hp[pw + pc];
                        const bool ok = (cur_op & M_OP) ? (hij == pe1) : ((int)Hi[iw + j - ib] == pe1 - e1);
                        if (ok) {
                            cur_op = ((int)Hp[pc] - oe1 == pe1) ? (M_OP | F_OP) : E1_OP;
                            cg[n_cig++] = ((uint32_t)v << 2) | 2; hit = true; i = pi;
                            break;
                        }
                    }
                    if (cur_op & E2_OP) {
                        const int pe2 = Hp[2 * pw + pc];
                        const bool ok = (cur_op & M_OP) ? (hij == pe2) : ((int)Hi[2 * iw + j - ib] == pe2 - e2);
                        if (ok) {
                            cur_op = ((int)Hp[pc] - oe2 == pe2) ? (M_OP | F_OP) : E2_OP;
                            cg[n_cig++] = ((uint32_t)v << 2) | 2; hit = true; i = pi;
                            break;
                        }
                    }
                }
            }
            if (!hit && (cur_op & F_OP)) {
                if (j - 1 < ib) { err = TH_ERR_BACKTRACK; break; }
                if (cur_op & F1_OP) {
                    const int f1 = Hi[3 * iw + j - ib];
                    if (!(cur_op & M_OP) || hij == f1) {
                        if ((int)Hi[j - 1 - ib] - oe1 == f1) { cur_op = M_OP | E_OP; hit = true; }
                        else if ((int)Hi[3 * iw + j - 1 - ib] - e1 == f1) { cur_op = F1_OP; hit = true; }
                        else { err = TH_ERR_BACKTRACK; break; }
                    }
                }
                if (!hit && (cur_op & F2_OP)) {
                    const int f2 = Hi[4 * iw + j - ib];
                    if (!(cur_op & M_OP) || hij == f2) {
                        if ((int)Hi[j - 1 - ib] - oe2 == f2) { cur_op = M_OP | E_OP; hit = true; }
                        else if ((int)Hi[4 * iw + j - 1 - ib] - e2 == f2) { cur_op = F2_OP; hit = true; }
                        else { err = TH_ERR_BACKTRACK; break; }
                    }
                }
                PUSH_I(1); --j; hit = true;
            }
            if (!hit) { err = TH_ERR_BACKTRACK; break; }
        }
        if (err == TH_OK && j > 0) PUSH_I(j);
#undef PUSH_I
    }
    err = __shfl_sync(TH_FULL, err, 0);
    if (err != TH_OK) return err;
    n_cig = __shfl_sync(TH_FULL, n_cig, 0);
    // ---- merge the alignment into the graph (abpoa_graph.c:1218-1284), lane 0 ---------------
    int n_ev = 0;
    if (lane == 0) {
        int query_id = -1, last_id = 0; bool last_new = false; int pend = 0; // events [pend, n_ev) wait for their anchor
        for (int c = n_cig - 1; c >= 0; --c) {
            const uint32_t cv = w.cigar[c]; const int op = cv & 3;
            if (op == 0) {
                const int node = (int)(cv >> 2);
                ++query_id;
                const uint8_t qb = query[query_id];
                int bf = w.n2i[node], bl = bf;               // block of the aligned group in the old order
                for (int a = 0; a < w.aln_n[node]; ++a) { const int x = w.n2i[w.aln[node * 4 + a]]; bf = min(bf, x); bl = max(bl, x); }
                for (; pend < n_ev; ++pend) w.ev_anchor[pend] = bf;
                if (w.base[node] != qb) {
                    int aid = -1;
                    for (int a = 0; a < w.aln_n[node]; ++a) { const int x = w.aln[node * 4 + a]; if (w.base[x] == qb) { aid = x; break; } }
                    if (aid != -1) { g_add_edge(w, edge_n, last_id, aid, !last_new); last_id = aid; last_new = false; }
                    else {
                        const int x = g_new_node(w, node_n, qb);
                        g_add_edge(w, edge_n, last_id, x, false); last_id = x; last_new = true;
                        const int n0 = w.aln_n[node]; // abpoa_add_graph_aligned_node (:1036-1044)
                        for (int a = 0; a < n0; ++a) { const int y = w.aln[node * 4 + a]; w.aln[y * 4 + w.aln_n[y]++] = x; w.aln[x * 4 + w.aln_n[x]++] = y; }
                        w.aln[node * 4 + w.aln_n[node]++] = x; w.aln[x * 4 + w.aln_n[x]++] = node;
                        w.ev_node[n_ev] = x; w.ev_anchor[n_ev] = bl + 1; ++n_ev; pend = n_ev;
                    }
                } else { g_add_edge(w, edge_n, last_id, node, !last_new); last_id = node; last_new = false; }
            } else if (op == 1) {
                const int len = (int)(cv >> 2);
                query_id += len;
                for (int jj = len - 1; jj >= 0; --jj) {
                    const int x = g_new_node(w, node_n, query[query_id - jj]);
                    g_add_edge(w, edge_n, last_id, x, false); last_id = x; last_new = true;
                    w.ev_node[n_ev++] = x;
                }
            }
        }
        g_add_edge(w, edge_n, last_id, 1, !last_new);
        for (; pend < n_ev; ++pend) w.ev_anchor[pend] = n - 1; // before the sink
    }
    n_ev = __shfl_sync(TH_FULL, n_ev, 0);
    node_n = __shfl_sync(TH_FULL, node_n, 0); edge_n = __shfl_sync(TH_FULL, edge_n, 0);
    __syncwarp();
    // new order: old node at index i moves to i + #(events with anchor <= i); event e lands at anchor_e + e
    for (int i = lane; i < n; i += 32) {
        int lo = 0, hi = n_ev;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (w.ev_anchor[mid] <= i) lo = mid + 1; else hi = mid; }
        w.ord2[i + lo] = w.ord[i];
    }
    for (int e = lane; e < n_ev; e += 32) w.ord2[w.ev_anchor[e] + e] = w.ev_node[e];
    __syncwarp();
    { int32_t *t = w.ord; w.ord = w.ord2; w.ord2 = t; }
    return TH_OK;
}

// heaviest-column consensus (abpoa_graph.c:279-359, 604-648, 467-478); lane 0.  Returns cons_len.
__device__ int poa_consensus(PoaWs &w, int node_n, int n_seq, uint8_t *cons, int32_t *cov) {
    int32_t *deg = w.n2i, *stk = w.ord2, *rank = w.remain;
    int32_t *rcw = reinterpret_cast<int32_t *>(w.arena); // 5 x msa_l weights, then 5 x msa_l node ids
    for (int i = 0; i < node_n; ++i) { int d = 0; for (int e = w.in_head[i]; e >= 0; e = w.e_ni[e]) ++d; deg[i] = d; }
    int sp = 0, msa_rank = 0;
    stk[sp++] = 0; rank[0] = -1;
    while (sp > 0) {
        const int cur = stk[--sp];
        if (rank[cur] < 0) {
            rank[cur] = msa_rank;
            for (int a = 0; a < w.aln_n[cur]; ++a) rank[w.aln[cur * 4 + a]] = msa_rank;
            ++msa_rank;
        }
        if (cur == 1) break;
        for (int e = w.out_head[cur]; e >= 0; e = w.e_no[e]) {
            const int o = w.e_to[e];
            if (--deg[o] == 0) {
                bool ok = true;
                for (int a = 0; a < w.aln_n[o]; ++a) if (deg[w.aln[o * 4 + a]] != 0) { ok = false; break; }
                if (!ok) continue;
                stk[sp++] = o; rank[o] = -1;
                for (int a = 0; a < w.aln_n[o]; ++a) { const int x = w.aln[o * 4 + a]; stk[sp++] = x; rank[x] = -1; }
            }
        }
    }
    const int msa_l = rank[1] - 1;
    if (msa_l <= 0) return 0;
    if ((uint64_t)msa_l * 10 * 2 > w.arena_cap) return -1;
    int32_t *nodeid = rcw + 5 * (size_t)msa_l;
    for (int i = 0; i < 5 * msa_l; ++i) { rcw[i] = 0; nodeid[i] = 0; }
    for (int i = 2; i < node_n; ++i) { // abpoa_set_row_column_weight; popcount(read_ids) == sum of out weights
        int rk = rank[i];
        for (int a = 0; a < w.aln_n[i]; ++a) rk = max(rk, rank[w.aln[i * 4 + a]]);
        int wsum = 0;
        for (int e = w.out_head[i]; e >= 0; e = w.e_no[e]) wsum += w.e_w[e];
        const int b = w.base[i] > 4 ? 4 : w.base[i];
        rcw[(rk - 1) * 5 + b] += wsum;
        nodeid[(rk - 1) * 5 + b] = i;
    }
    int last_id = 0, cons_i = 0;
    int32_t *max_out = w.hs;
    for (int i = 0; i < msa_l; ++i) {
        int max_w = 0, max_base = 5, gap_w = n_seq;
        for (int b = 0; b < 4; ++b) { const int x = rcw[i * 5 + b]; if (x > max_w) { max_base = b; max_w = x; } gap_w -= x; }
        if (max_w >= gap_w) {
            const int cur = nodeid[i * 5 + max_base];
            max_out[last_id] = cur; last_id = cur;
            cov[cons_i++] = max_w;
        }
    }
    max_out[last_id] = 1;
    int id = max_out[0], l = 0;
    while (id != 1 && l <= node_n) { cons[l++] = w.base[id]; id = max_out[id]; }
    return l;
}

// persistent warps pull tasks from an atomic counter
__global__ void __launch_bounds__(POA_WARPS * 32)
poa_kernel(DevParams P, int n_tasks, const PoaTask *__restrict__ tasks, const int32_t *__restrict__ task_order,
           const int32_t *__restrict__ u_start, const int32_t *__restrict__ u_len, const uint8_t *__restrict__ bseq,
           uint8_t *slabs, size_t slab_bytes, int *task_counter,
           uint8_t *__restrict__ cons_base, int32_t *__restrict__ cons_cov, int32_t *__restrict__ cons_len,
           int32_t *__restrict__ task_status, unsigned long long *__restrict__ stat_cells, unsigned long long *__restrict__ stat_rows) {
    __shared__ int s_pre[POA_WARPS][POA_MAXPRE * 4];
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * POA_WARPS + wib;
    uint8_t *slab = slabs + (size_t)gw * slab_bytes;
    unsigned long long cells = 0, rows = 0;
    while (true) {
        int ti = 0;
        if (lane == 0) ti = atomicAdd(task_counter, 1);
        ti = __shfl_sync(TH_FULL, ti, 0);
        if (ti >= n_tasks) break;
        const int t = task_order ? task_order[ti] : ti;
        const PoaTask T = tasks[t];
        const uint8_t *rseq = bseq + T.seq_off;
        uint8_t *cons = cons_base + T.cons_off; int32_t *cov = cons_cov + T.cons_off;
        if (T.n_seqs < 2) { if (lane == 0) { cons_len[t] = 0; task_status[t] = TH_ERR_CAP; } continue; } // the reference aborts here (abpoa_cons.c:58)
        if (T.n_seqs <= 2) { // src/abpoa_cons.c:57-80: the first unit verbatim
            const int l0 = u_len[T.unit_off]; const uint8_t *s0 = rseq + u_start[T.unit_off];
            for (int i = lane; i < l0; i += 32) { cons[i] = s0[i]; cov[i] = 0; }
            if (lane == 0) { cons_len[t] = l0; task_status[t] = TH_OK; }
            continue;
        }
        if (poa_fixed_bytes(T.ncap, T.qmax, T.n_seqs) + 4096 > slab_bytes) { if (lane == 0) { cons_len[t] = 0; task_status[t] = TH_ERR_ARENA; } continue; }
        PoaWs w;
        poa_carve(w, slab, slab_bytes, T.ncap, T.qmax, T.n_seqs);
        // first sequence: a chain of new nodes (abpoa_graph.c:1108-1124)
        const int l0 = u_len[T.unit_off]; const uint8_t *s0 = rseq + u_start[T.unit_off];
        for (int i = lane; i < l0 + 2; i += 32) { w.out_head[i] = w.out_tail[i] = w.in_head[i] = w.in_tail[i] = -1; w.aln_n[i] = 0; }
        __syncwarp();
        for (int i = lane; i <= l0; i += 32) {
            const int from = i == 0 ? 0 : 1 + i, to = i == l0 ? 1 : 2 + i;
            w.e_to[i] = to; w.e_from[i] = from; w.e_w[i] = 1; w.e_no[i] = -1; w.e_ni[i] = -1;
            w.out_head[from] = w.out_tail[from] = i; w.in_head[to] = w.in_tail[to] = i;
            if (i < l0) { w.base[2 + i] = s0[i]; w.ord[1 + i] = 2 + i; }
        }
        if (lane == 0) { w.ord[0] = 0; w.ord[l0 + 1] = 1; w.base[0] = w.base[1] = 4; }
        __syncwarp();
        int node_n = l0 + 2, edge_n = l0 + 1, err = TH_OK;
        for (int s = 1; s < T.n_seqs && err == TH_OK; ++s)
            err = poa_add_sequence(w, P, rseq + u_start[T.unit_off + s], u_len[T.unit_off + s], node_n, edge_n, s_pre[wib], cells, rows);
        int cl = 0;
        if (err == TH_OK) {
            if (lane == 0) { cl = poa_consensus(w, node_n, T.n_seqs, cons, cov); }
            cl = __shfl_sync(TH_FULL, cl, 0);
            if (cl < 0) { err = TH_ERR_ARENA; cl = 0; }
        }
        if (lane == 0) { cons_len[t] = cl; task_status[t] = err; }
        __syncwarp();
    }
    if (lane == 0 && cells) { atomicAdd(stat_cells, cells); atomicAdd(stat_rows, rows); }
}
