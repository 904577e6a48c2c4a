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
// in insertion order, two order buffers, per-row band metadata, room for the query bit-planes of long queries
// (short ones keep them in shared memory), cigar,
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
#define POA_RING 64       // rows of metadata kept in shared memory per warp (power of two)

// Row descriptor (static per alignment, by row index):  x = first predecessor row (-1: none),
//   y = np | base << 10 | node << 13,  z = qlen - max_remain term of the band centre,
//   w = second predecessor row (np == 2) or offset into plist (np > 2).
// Row metadata (written when the row is computed):  x = arena offset, y = first column, z = last column,
//   w = max_i + 1 of the row (what the reference scatters into max_pos_left/right of the successors).
struct PoaWs {
    int4 *rdesc, *rmeta;
    int32_t *out_head, *out_tail, *in_head, *in_tail, *aln_n, *aln, *n2i, *ri, *hs;
    int32_t *e_to, *e_from, *e_w, *e_no, *e_ni, *plist;
    int32_t *ord, *ord2, *ev_anchor, *ev_node, *hi_idx;
    int16_t *qp; uint32_t *cigar; int32_t *cigq; uint8_t *base;
    int16_t *arena; uint32_t arena_cap;
    int32_t qp_stride, ncap;
};

__host__ __device__ inline size_t poa_fixed_bytes(int ncap, int qmax, int nseq) {
    size_t ecap = (size_t)ncap + nseq + 2;
    size_t b = 0;
    b += (size_t)ncap * 16 * 2;                        // rdesc, rmeta
    b += (size_t)ncap * 4 * (9 + 3);                   // 9 node arrays (aln counts as 4) = 12 x int32
    b += ecap * 4 * 6;                                 // 5 edge arrays + plist
    b += (size_t)ncap * 4 * 3;                         // ord, ord2, hi_idx
    b += (size_t)(qmax + 2) * 4 * 2;                   // events
    b += (size_t)5 * (qmax + 1 + 128) * 2 + 8;         // profile
    b += (size_t)(qmax + ncap + 8) * 4 * 2;            // cigar, cigq
    b += (size_t)ncap + 64;                            // base
    return (b + 4095) & ~(size_t)4095;
}

__device__ inline void poa_carve(PoaWs &w, uint8_t *slab, size_t slab_bytes, int ncap, int qmax, int nseq) {
    size_t ecap = (size_t)ncap + nseq + 2;
    w.ncap = ncap;
    w.rdesc = reinterpret_cast<int4 *>(slab); w.rmeta = w.rdesc + ncap;
    int32_t *p = reinterpret_cast<int32_t *>(w.rmeta + ncap);
    w.out_head = p; p += ncap; w.out_tail = p; p += ncap; w.in_head = p; p += ncap; w.in_tail = p; p += ncap;
    w.aln_n = p; p += ncap; w.aln = p; p += 4 * (size_t)ncap; w.n2i = p; p += ncap; w.ri = p; p += ncap; w.hs = p; p += ncap;
    w.e_to = p; p += ecap; w.e_from = p; p += ecap; w.e_w = p; p += ecap; w.e_no = p; p += ecap; w.e_ni = p; p += ecap; w.plist = p; p += ecap;
    w.ord = p; p += ncap; w.ord2 = p; p += ncap; w.hi_idx = p; p += ncap;
    w.ev_anchor = p; p += qmax + 2; w.ev_node = p; p += qmax + 2;
    w.qp_stride = (qmax + 1 + 128) & ~1;
    w.qp = reinterpret_cast<int16_t *>(p); p += ((size_t)5 * w.qp_stride * 2 + 3) / 4;
    w.cigar = reinterpret_cast<uint32_t *>(p); p += qmax + ncap + 8;
    w.cigq = p; p += qmax + ncap + 8;
    w.base = reinterpret_cast<uint8_t *>(p);
    size_t fixed = poa_fixed_bytes(ncap, qmax, nseq);
    w.arena = reinterpret_cast<int16_t *>(slab + fixed);
    size_t ab = slab_bytes > fixed ? slab_bytes - fixed : 0;
    size_t ne = ab / 2; if (ne > 0xfffffff0ull) ne = 0xfffffff0ull;
    w.arena_cap = (uint32_t)ne;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// graph edit used for the final edge into the sink (lane 0 only)
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

// DP arena layout.  A row of W columns (W a multiple of pn) that starts at int16 offset `off` owns 5 W int16:
//   W/2 records of 16 bytes, record q = columns (beg + 2q, beg + 2q + 1) as four s16x2 words {H, E1, E2, F1},
//   followed by W/2 words of F2 pairs.  One 16-byte access moves everything the next rows need from a column
//   pair, and a backtrack step touches one or two sectors per row instead of five.
__device__ __forceinline__ int s16_at(uint32_t word, int odd) { return odd ? hi16(word) : lo16(word); }

#ifndef POA_SETUP_U
#define POA_SETUP_U 2      // batches of 32 nodes whose edge-list walks are interleaved when the row descriptors are built
#endif
#define POA_PEQ_W 40     // words per query bit-plane kept in shared memory (queries up to ~1200 columns; longer ones use the slab)
struct PoaSmem { int4 desc[POA_RING]; int4 meta[POA_RING]; int4 pre[POA_MAXPRE]; uint4 last[32]; uint32_t peq[5 * POA_PEQ_W]; };

// contributions of one predecessor row to columns (j, j+1) of the current row: M from H[p][j-1], H[p][j];
// E1, E2 from the same columns.  pm = the predecessor's row metadata, recs = its records ({H, E1, E2, F1} per column
// pair).  Cells outside the predecessor's band count as inf_min (simd_abpoa_align.c:860-905).
// The row this warp computed last (when it fits one 64-column chunk) is still in shared memory in the same record
// layout, so `recs` is a generic pointer to either place and one code path serves both.
__device__ __forceinline__ const uint4 *poa_row_recs(const uint32_t *A32, const uint4 *last, uint32_t off, uint32_t last_off) {
    return off == last_off ? last : reinterpret_cast<const uint4 *>(A32) + (off >> 3); // row offsets are multiples of 40 int16
}
__device__ __forceinline__ void poa_pred(const uint4 *recs, const int4 pm, const int j, const uint32_t INFP,
                                         uint32_t &Mx, uint32_t &E1x, uint32_t &E2x) {
    const int pb = pm.y, l2 = (j - pb) >> 1; // band starts are even, so is j
    uint32_t Xh = INFP, prev = INFP;
    if (j >= pb && j <= pm.z) { const uint4 r = recs[l2]; Xh = r.x; E1x = __vmaxs2(E1x, r.y); E2x = __vmaxs2(E2x, r.z); }
    if (j > pb && j - 1 <= pm.z) prev = recs[l2 - 1].x; // H of the pair to the left; only its upper half (column j-1) is used
    Mx = __vmaxs2(Mx, __funnelshift_r(prev, Xh, 16));
}

// one warp aligns sequence `query` to the graph and merges it in.  Returns an error code.
// AFFINE: abPOA's affine gap mode (gap_open2 == 0, simd_abpoa_ag_dp, simd_abpoa_align.c:739-833).  It is not the convex
// recurrence minus one gap function: an insertion opens from M only (F is built from the row's diagonal values before E
// is folded in), and the E handed to the next row is inf_min wherever F strictly won the cell (SIMDSetIfEqual), so
// insertions and deletions are never adjacent.  The second pair (E2, F2) is held at inf_min, which makes the convex
// backtrack below behave exactly as simd_abpoa_ag_backtrack (:160-246).
template <int LP, bool AFFINE> // LP = log2 of the emulated vector width pn (4: AVX2 int16, the reference build; 3: SSE)
__device__ int poa_add_sequence(PoaWs &w, const DevParams &P, const uint8_t *query, int qlen, int &node_n, int &edge_n,
                                PoaSmem &sm, unsigned long long &cells, unsigned long long &rows, long long *ph) {
#ifdef POA_PROFILE   // per-phase warp-cycle counters (tools/profile_step.py); cost ~16 registers, off in the product build
    long long t_ph = clock64();
#define PH(k) do { long long t_ = clock64(); ph[k] += t_ - t_ph; t_ph = t_; } while (0)
#else
#define PH(k) do { } while (0)
#endif
    const int lane = lane_id();
    const int n = node_n; constexpr int pn = 1 << LP, lp = LP;
    const int o1 = P.o1, e1 = P.e1, o2 = P.o2, e2 = P.e2, oe1 = o1 + e1, oe2 = o2 + e2;
    const int mis = P.mismatch > 0 ? P.mismatch : -P.mismatch, mat = P.match < 0 ? -P.match : P.match;
    { // int16 path only (simd_abpoa_align.c:1610-1621)
        int len = qlen > n ? qlen : n;
        int max_score = max(qlen * mat, len * e1 + o1);
        if (max_score > 32767 - mis - oe1 - max(oe2, P.o2_raw + P.e2_raw) - 64 * max(max(e1, e2), P.e2_raw)) return TH_ERR_LEN;
    }
    // simd_abpoa_align.c:1613-1614 uses the option values as given, whatever the gap mode
    const int inf_min = max(max(-32768 + mis, -32768 + oe1), -32768 + P.o2_raw + P.e2_raw) + 31 * max(e1, P.e2_raw);
    const uint32_t INFP = pk(inf_min, inf_min);
    const int wband = 10 + (int)(0.01f * (float)qlen); // wb + (int)(wf*qlen), float (simd_abpoa_align.c:393)
    // ---- row descriptors: order index, predecessors by row, heaviest successor ----------------
    for (int i = lane; i < n; i += 32) w.n2i[w.ord[i]] = i;
    __syncwarp();
    { // four batches of 32 nodes walk their edge lists in lock-step: the loads of a step are independent across the
      // batches, so four of these dependent (DRAM/L2-latency) pointer chases are in flight per lane instead of one
        constexpr int U = POA_SETUP_U;
        int pl_base = 0, bad = 0;
        for (int i0 = 0; i0 < n; i0 += 32 * U) {
            int np[U], p0[U], p1[U], hi[U], v[U], b[U], ei[U], eo[U], mw[U], mt[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + 32 * u + lane;
                np[u] = 0; p0[u] = -1; p1[u] = -1; hi[u] = 0x7fffffff; b[u] = 0; mw[u] = -1; mt[u] = -1;
                v[u] = i < n ? w.ord[i] : -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                ei[u] = -1; eo[u] = -1;
                if (v[u] >= 0) { b[u] = w.base[v[u]]; ei[u] = w.in_head[v[u]]; eo[u] = w.out_head[v[u]]; }
            }
            while (true) { // predecessors by row, in in_id order
                bool any = false;
#pragma unroll
                for (int u = 0; u < U; ++u) any |= ei[u] >= 0;
                if (!any) break;
                int fr[U], nx[U];
#pragma unroll
                for (int u = 0; u < U; ++u) if (ei[u] >= 0) { fr[u] = w.e_from[ei[u]]; nx[u] = w.e_ni[ei[u]]; }
#pragma unroll
                for (int u = 0; u < U; ++u) if (ei[u] >= 0) {
                    const int pi = w.n2i[fr[u]];
                    if (np[u] == 0) p0[u] = pi; else if (np[u] == 1) p1[u] = pi;
                    ++np[u]; ei[u] = nx[u];
                }
            }
            while (true) { // first out-edge with maximum weight (abpoa_graph.c:216-226)
                bool any = false;
#pragma unroll
                for (int u = 0; u < U; ++u) any |= eo[u] >= 0;
                if (!any) break;
                int ww[U], to[U], nx[U];
#pragma unroll
                for (int u = 0; u < U; ++u) if (eo[u] >= 0) { ww[u] = w.e_w[eo[u]]; to[u] = w.e_to[eo[u]]; nx[u] = w.e_no[eo[u]]; }
#pragma unroll
                for (int u = 0; u < U; ++u) if (eo[u] >= 0) { if (ww[u] > mw[u]) { mw[u] = ww[u]; mt[u] = to[u]; } eo[u] = nx[u]; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) if (mt[u] >= 0) hi[u] = w.n2i[mt[u]];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + 32 * u + lane;
                const int cnt = np[u] > 2 ? np[u] : 0;
                int inc = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(TH_FULL, inc, d); if (lane >= d) inc += o; }
                const int off = pl_base + inc - cnt;
                pl_base += __shfl_sync(TH_FULL, inc, 31);
                if (i < n) {
                    if (np[u] > 2) { int k = 0; for (int e = w.in_head[v[u]]; e >= 0; e = w.e_ni[e]) w.plist[off + k++] = w.n2i[w.e_from[e]]; }
                    w.hi_idx[i] = hi[u];
                    bad |= np[u] > 1023 || (i > 0 && i < n - 1 && (unsigned)(np[u] - 1) >= POA_MAXPRE); // the row loop relies on 1 <= np <= POA_MAXPRE
                    w.rdesc[i] = make_int4(p0[u], min(np[u], 1023) | (min(b[u], 7) << 10) | (v[u] << 13), 0, np[u] > 2 ? off : p1[u]);
                }
            }
        }
        if (__any_sync(TH_FULL, bad)) return TH_ERR_CAP;
    }
    __syncwarp();
    // max_remain by index, 32 indices at a time from the sink backwards; chains inside a chunk are
    // collapsed by pointer jumping on shuffles (remain[v] = remain[heaviest successor] + 1)
    int32_t *ri = w.ri;
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
            reinterpret_cast<int32_t *>(w.rdesc + idx)[2] = qlen - val;
        }
        __syncwarp();
    }
    // ---- query profile (simd_abpoa_align.c:438-446) as bit-planes: bit j of plane b < 4 = (query[j-1] == b), bit j of
    // plane 4 = (query[j-1] is A/C/G/T and 1 <= j <= qlen).  The score pair of columns (j, j+1) against node base b is
    // then two word loads, shifts and IMADs: mat where the match bit is set, -mis where only the valid bit is set, 0
    // elsewhere (N in the query, column 0, columns past the query end) or when the node itself is N -- exactly the
    // reference's 5 x qlen table, without a trip to a table in the slab at the head of every row's dependency chain.
    const int prof_w = ((qlen / pn + 1) * pn + 64 + 1) & ~1;
    const int peq_w = (prof_w >> 5) + 1;
    uint32_t *const peq = peq_w <= POA_PEQ_W ? sm.peq : reinterpret_cast<uint32_t *>(w.qp);
    for (int j0 = 0; j0 < peq_w * 32; j0 += 32) {
        const int j = j0 + lane;
        const int qc = (j >= 1 && j <= qlen) ? (int)query[j - 1] : 7;
        const uint32_t b0 = __ballot_sync(TH_FULL, qc == 0), b1 = __ballot_sync(TH_FULL, qc == 1), b2 = __ballot_sync(TH_FULL, qc == 2),
                       b3 = __ballot_sync(TH_FULL, qc == 3), bv = __ballot_sync(TH_FULL, qc < 4);
        if (lane < 5) peq[lane * peq_w + (j0 >> 5)] = lane == 0 ? b0 : lane == 1 ? b1 : lane == 2 ? b2 : lane == 3 ? b3 : bv;
    }
    // ---- first row (simd_abpoa_align.c:538-555, 591-610) ------------------------------------
    uint32_t used = 0;
    {
        const int end = min(qlen, max(0, qlen - ri[0]) + wband);
        const int esn = end >> lp, width = (esn + 1) << lp;
        if ((uint64_t)used + 5ull * width > w.arena_cap) return TH_ERR_ARENA;
        if (lane == 0) { const int4 m = make_int4(0, 0, width - 1, 1); w.rmeta[0] = m; sm.meta[0] = m; } // the source hands 1 to its successors (:549-552)
        uint4 *R = reinterpret_cast<uint4 *>(w.arena); uint32_t *F2w = reinterpret_cast<uint32_t *>(w.arena) + 2 * width;
        for (int q = lane; q < width / 2; q += 32) {
            uint32_t hh = 0, ee1 = 0, ee2 = 0, ff1 = 0, ff2 = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * q + h;
                const int f1 = -o1 - e1 * j, f2 = -o2 - e2 * j;
                const int vh = j == 0 ? 0 : (AFFINE ? (int)(int16_t)f1 : max((int)(int16_t)f1, (int)(int16_t)f2));
                const int v1 = j == 0 ? -oe1 : inf_min, v2 = (j == 0 && !AFFINE) ? -oe2 : inf_min, vf1 = j == 0 ? inf_min : f1, vf2 = (j == 0 || AFFINE) ? inf_min : f2;
                hh |= (uint32_t)(uint16_t)vh << (16 * h); ee1 |= (uint32_t)(uint16_t)v1 << (16 * h); ee2 |= (uint32_t)(uint16_t)v2 << (16 * h);
                ff1 |= (uint32_t)(uint16_t)vf1 << (16 * h); ff2 |= (uint32_t)(uint16_t)vf2 << (16 * h);
            }
            R[q] = make_uint4(hh, ee1, ee2, ff1); F2w[q] = ff2;
        }
        used += 5u * width; cells += width; rows += 1;
    }
    __syncwarp();
    PH(0);
    // ---- rows in topological order ----------------------------------------------------------
    // Value bounds that make plain 16-bit adds exact here: every H is >= inf_min - mis (the M term), so
    // H - oe, E - e and the final F never reach -32768: the reference's saturating subtractions
    // (_mm256_subs_epi16 in SIMD_SET_F) only ever clip intermediate candidates that lose the max anyway.
    // F1[j] = max_k<=j (A1[k] - e1 (j-k)) is evaluated as a prefix MAX of G[k] = A1[k] + e1 (k - j0) over the
    // chunk (no subtraction inside the scan, hence no underflow), F = G - e1 (j - j0).  The int16 range check
    // above leaves 64 max(e1,e2) of headroom for G.
    const uint32_t NOE1P = pk(-oe1, -oe1), NOE2P = pk(-oe2, -oe2), NE1P = pk(-e1, -e1), NE2P = pk(-e2, -e2);
    const uint32_t C1 = pk(e1 * 2 * lane - oe1, e1 * (2 * lane + 1) - oe1), C2 = pk(e2 * 2 * lane - oe2, e2 * (2 * lane + 1) - oe2);
    const uint32_t NJ1 = pk(-e1 * 2 * lane, -e1 * (2 * lane + 1)), NJ2 = pk(-e2 * 2 * lane, -e2 * (2 * lane + 1));
    const int lam_bits = pn - 1;
    const uint32_t lamk_lo = (uint32_t)(lam_bits - ((2 * lane) & lam_bits)) << 12, lamk_hi = (uint32_t)(lam_bits - ((2 * lane + 1) & lam_bits)) << 12;
    const int lane_vec = (2 * lane) >> lp;
    const int qsn = qlen >> lp;
    const uint32_t *A32 = reinterpret_cast<const uint32_t *>(w.arena);
    uint32_t *A32w = reinterpret_cast<uint32_t *>(w.arena);
    uint32_t last_off = 0xffffffffu; // arena offset of the row whose values sit in sm.last
    const uint32_t NEGMIS2 = pk(-mis, -mis), XMM = (uint32_t)(uint16_t)mat ^ (uint32_t)(uint16_t)(-mis); // -mis ^ XMM == mat per half
    const uint32_t KLO = lamk_lo | 0xfffu, KHI = lamk_hi | 0xfffu;
    const uint32_t used_rows0 = used;
    int4 *const rmeta_g = w.rmeta; const int4 *const rdesc_g = w.rdesc; const int32_t *const plist_g = w.plist;
    const uint32_t arena_cap = w.arena_cap;
    const uint32_t *const vrow = peq + 4 * peq_w; // valid plane
    for (int i0 = 0; i0 < n - 1; i0 += 32) {
      { const int idx = i0 + lane; // descriptors of the next 32 rows
        if (idx < n) sm.desc[idx & (POA_RING - 1)] = rdesc_g[idx];
        __syncwarp(); }
      const int i_end = min(i0 + 32, n - 1);
      for (int i = max(i0, 1); i < i_end; ++i) {
        const int4 d = sm.desc[i & (POA_RING - 1)];
        const int np = d.y & 1023; // 1..POA_MAXPRE, checked when the descriptors were built
        // band: what the predecessors' row maxima and max_remain say (abpoa_align.h:34-35, simd_abpoa_align.c:846-854)
        const int4 pm0 = (i - d.x < POA_RING) ? sm.meta[d.x & (POA_RING - 1)] : rmeta_g[d.x];
        int mpl = min(n, pm0.w), mpr = max(0, pm0.w), min_pre_beg = pm0.y;
        if (np > 1) {
            for (int p = 1; p < np; ++p) {
                const int pi = np == 2 ? d.w : plist_g[d.w + p];
                const int4 m = (i - pi < POA_RING) ? sm.meta[pi & (POA_RING - 1)] : rmeta_g[pi];
                mpl = min(mpl, m.w); mpr = max(mpr, m.w); min_pre_beg = min(min_pre_beg, m.y);
                if (lane == 0) sm.pre[p] = m;
            }
            __syncwarp();
        }
        const int beg0 = max(0, min(mpl, d.z) - wband), end0 = min(qlen, max(mpr, d.z) + wband);
        const int beg = max((beg0 >> lp) << lp, min_pre_beg), esn = end0 >> lp, dend = ((esn + 1) << lp) - 1;
        const int bsn = beg >> lp, width = dend - beg + 1;
        const uint32_t w5 = 5u * (uint32_t)width;
        if ((beg > dend) | (w5 > arena_cap - used)) return beg > dend ? TH_ERR_BAND : TH_ERR_ARENA;
        const uint32_t row_off = used;
        used += w5; // cells and rows are derived from `used` after the loop
        const int vb = (d.y >> 10) & 7;
        const uint32_t *const prow = peq + (vb & 3) * peq_w;
        const uint32_t s_neg = vb < 4 ? NEGMIS2 : 0u, s_xm = vb < 4 ? XMM : 0u;  // an N node scores 0 against everything
        const uint32_t rec0 = (row_off >> 3) + lane;                                  // this lane's record in chunk 0 (16-byte units; row_off is a multiple of 40)
        const uint32_t f20 = (row_off >> 1) + 2u * (uint32_t)width + lane;              // ... and its F2 pair (words)
        const int jmax = esn == qsn ? qlen : dend;       // columns past the query end do not compete for the row maximum
        const int vlast = esn - bsn;                      // the row's last vector is visited first by the reference's arg-max
        int best = INT_MIN;
        if (width <= 64) {
            // ---- the usual row: one 64-column chunk, no carries between chunks --------------------------------------
            const int j = beg + 2 * lane;
            uint32_t Mx = INFP, E1x = INFP, E2x = INFP;
            poa_pred(poa_row_recs(A32, sm.last, (uint32_t)pm0.x, last_off), pm0, j, INFP, Mx, E1x, E2x);
            for (int p = 1; p < np; ++p) {
                const int4 pm = sm.pre[p];
                poa_pred(poa_row_recs(A32, sm.last, (uint32_t)pm.x, last_off), pm, j, INFP, Mx, E1x, E2x);
            }
            const uint32_t sh = j & 31, t = (prow[j >> 5] >> sh) & 3u, v = (vrow[j >> 5] >> sh) & 3u; // bits of columns j, j + 1 (j is even)
            const uint32_t S = (s_neg ^ (((t | (t << 15)) & 0x10001u) * s_xm)) & (((v | (v << 15)) & 0x10001u) * 0xffffu);
            const uint32_t Ms = __vadd2(Mx, S);
            const uint32_t Hme = __vimax3_s16x2(Ms, E1x, E2x);
            const uint32_t Hf = AFFINE ? Ms : Hme;        // what an insertion may open from
            uint32_t hp = __shfl_up_sync(TH_FULL, Hf, 1);
            if (lane == 0) hp = Ms << 16;
            const uint32_t Hsh = __funnelshift_r(hp, Hf, 16);
            uint32_t G1 = __vadd2(Hsh, C1), G2 = __vadd2(Hsh, C2);   // G = (Hme[j-1] - oe) + e (j - j0)
            G1 = __vmaxs2(G1, (G1 << 16) | 0x8000u); G2 = __vmaxs2(G2, (G2 << 16) | 0x8000u); // odd column also sees the even one
            uint32_t TT = __byte_perm(G1, G2, 0x7632); // lo = G1 at this lane's odd column, hi = G2
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) TT = __vmaxs2(TT, __shfl_up_sync(TH_FULL, TT, dd)); // lanes < dd get their own value back
            uint32_t Pv = __shfl_up_sync(TH_FULL, TT, 1);
            if (lane == 0) Pv = POA_NEGP;
            G1 = __vmaxs2(G1, __byte_perm(Pv, Pv, 0x1010)); G2 = __vmaxs2(G2, __byte_perm(Pv, Pv, 0x3232));
            const uint32_t Fa = __vadd2(G1, NJ1), Fb = AFFINE ? INFP : __vadd2(G2, NJ2);
            const uint32_t Hn = __vimax3_s16x2(Hme, Fa, Fb);
            uint32_t E1o = __viaddmax_s16x2(E1x, NE1P, __vadd2(Hn, NOE1P));
            if (AFFINE) { const uint32_t keep = __vcmpeq2(Hn, Hme); E1o = (E1o & keep) | (INFP & ~keep); } // F won the cell: no deletion from it
            const uint32_t E2o = AFFINE ? INFP : __viaddmax_s16x2(E2x, NE2P, __vadd2(Hn, NOE2P));
            sm.last[lane] = make_uint4(Hn, E1o, E2o, 0); // every reader of the old contents is past the scan's shuffles
            if (j <= dend) { reinterpret_cast<uint4 *>(A32w)[rec0] = make_uint4(Hn, E1o, E2o, Fa); A32w[f20] = Fb; }
            const uint32_t sub = lane_vec == vlast ? 0u : (uint32_t)(lane_vec + 1);
            const int klo = (int)__byte_perm(Hn, KLO - sub, 0x1054), khi = (int)__byte_perm(Hn, KHI - sub, 0x3254);
            best = max(j <= jmax ? klo : INT_MIN, j < jmax ? khi : INT_MIN); // jmax <= dend
            last_off = row_off;
        } else {
        uint32_t carryH = 0, carryF = POA_NEGP; // carryF: (F1 - e1, F2 - e2) of the previous chunk's last column
        const int nchunk = (width + 63) >> 6;
        for (int ch = 0; ch < nchunk; ++ch) {
            const int j = beg + (ch << 6) + 2 * lane;
            uint32_t Mx = INFP, E1x = INFP, E2x = INFP;
            poa_pred(poa_row_recs(A32, sm.last, (uint32_t)pm0.x, last_off), pm0, j, INFP, Mx, E1x, E2x);
            for (int p = 1; p < np; ++p) {
                const int4 pm = sm.pre[p];
                poa_pred(poa_row_recs(A32, sm.last, (uint32_t)pm.x, last_off), pm, j, INFP, Mx, E1x, E2x);
            }
            const uint32_t sh = j & 31, t = (prow[j >> 5] >> sh) & 3u, v = (vrow[j >> 5] >> sh) & 3u;
            const uint32_t S = (s_neg ^ (((t | (t << 15)) & 0x10001u) * s_xm)) & (((v | (v << 15)) & 0x10001u) * 0xffffu);
            const uint32_t Ms = __vadd2(Mx, S);
            const uint32_t Hme = __vimax3_s16x2(Ms, E1x, E2x);
            const uint32_t Hf = AFFINE ? Ms : Hme;
            uint32_t hp = __shfl_up_sync(TH_FULL, Hf, 1);
            if (lane == 0) hp = ch == 0 ? (Ms << 16) : carryH;
            const uint32_t Hsh = __funnelshift_r(hp, Hf, 16);
            uint32_t G1 = __vadd2(Hsh, C1), G2 = __vadd2(Hsh, C2);   // G = (Hme[j-1] - oe) + e (j - j0)
            if (lane == 0) { G1 = __vmaxs2(G1, (carryF & 0xffffu) | 0x80000000u); G2 = __vmaxs2(G2, (carryF >> 16) | 0x80000000u); }
            G1 = __vmaxs2(G1, (G1 << 16) | 0x8000u); G2 = __vmaxs2(G2, (G2 << 16) | 0x8000u); // odd column also sees the even one
            uint32_t TT = __byte_perm(G1, G2, 0x7632); // lo = G1 at this lane's odd column, hi = G2
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) TT = __vmaxs2(TT, __shfl_up_sync(TH_FULL, TT, dd)); // lanes < dd get their own value back
            uint32_t Pv = __shfl_up_sync(TH_FULL, TT, 1);
            if (lane == 0) Pv = POA_NEGP;
            G1 = __vmaxs2(G1, __byte_perm(Pv, Pv, 0x1010)); G2 = __vmaxs2(G2, __byte_perm(Pv, Pv, 0x3232));
            const uint32_t Fa = __vadd2(G1, NJ1), Fb = AFFINE ? INFP : __vadd2(G2, NJ2);
            const uint32_t Hn = __vimax3_s16x2(Hme, Fa, Fb);
            uint32_t E1o = __viaddmax_s16x2(E1x, NE1P, __vadd2(Hn, NOE1P));
            if (AFFINE) { const uint32_t keep = __vcmpeq2(Hn, Hme); E1o = (E1o & keep) | (INFP & ~keep); } // F won the cell: no deletion from it
            const uint32_t E2o = AFFINE ? INFP : __viaddmax_s16x2(E2x, NE2P, __vadd2(Hn, NOE2P));
            if (ch + 1 < nchunk) {
                carryH = __shfl_sync(TH_FULL, Hf, 31) & 0xffff0000u;
                const uint32_t fa = __shfl_sync(TH_FULL, Fa, 31), fb = __shfl_sync(TH_FULL, Fb, 31);
                carryF = __vadd2(__byte_perm(fa, fb, 0x7632), pk(-e1, -e2)); // G of column j0 - 1 in the next chunk's frame
            }
            if (j <= dend) { reinterpret_cast<uint4 *>(A32w)[rec0 + (ch << 5)] = make_uint4(Hn, E1o, E2o, Fa); A32w[f20 + (ch << 5)] = Fb; }
            { // row arg-max key (signed compare): value, then lane (j mod pn) ascending, then vector order with end_sn first
                const int rel = lane_vec + (ch << (6 - lp));
                const uint32_t sub = rel == vlast ? 0u : (uint32_t)(rel + 1);
                const int klo = (int)__byte_perm(Hn, KLO - sub, 0x1054), khi = (int)__byte_perm(Hn, KHI - sub, 0x3254);
                best = max(best, max(j <= jmax ? klo : INT_MIN, j < jmax ? khi : INT_MIN)); // jmax <= dend
            }
        }
        last_off = 0xffffffffu; // sm.last keeps an older row; it is matched by offset, and that offset is forgotten here
        }
        best = __reduce_max_sync(TH_FULL, best);
        { // simd_abpoa_max_in_row + simd_abpoa_ada_max_i: successors pull max_i + 1 from this row's metadata
            const int val = best >> 16;
            const int lam = lam_bits - (int)((best >> 12) & 0xf), vr = 0xfff - (int)(best & 0xfff);
            const int vsn = vr == 0 ? esn : bsn + vr - 1;
            const int max_i = (best != INT_MIN && val > inf_min) ? vsn * pn + lam : -1;
            if (lane == 0) { const int4 m = make_int4((int)row_off, beg, dend, max_i + 1); sm.meta[i & (POA_RING - 1)] = m; rmeta_g[i] = m; }
        }
        __syncwarp();
      }
    }
    cells += (used - used_rows0) / 5u; rows += (unsigned long long)max(n - 2, 0);
    PH(1);
    // ---- best end cell (simd_abpoa_align.c:976-989): sink's in-neighbours in in_id order, strict > ----
    int bi = 0, bj = 0;
    {
        const int4 ds = w.rdesc[n - 1];
        const int nps = ds.y & 1023;
        int best_score = inf_min;
        for (int p0 = 0; p0 < nps; p0 += 32) {
            const int p = p0 + lane;
            int s = -0x7fffffff, pi = 0, end = 0;
            if (p < nps) {
                pi = p == 0 ? ds.x : (nps == 2 ? ds.w : w.plist[ds.w + p]);
                const int4 m = w.rmeta[pi];
                end = qlen > m.z ? m.z : qlen;
                s = s16_at(A32[(m.x >> 1) + 4 * ((end - m.y) >> 1)], (end - m.y) & 1);
            }
            const int mx = __reduce_max_sync(TH_FULL, s);
            if (mx > best_score) {
                const unsigned bm = __ballot_sync(TH_FULL, s == mx);
                const int f = __ffs(bm) - 1;
                best_score = mx; bi = __shfl_sync(TH_FULL, pi, f); bj = __shfl_sync(TH_FULL, end, f);
            }
        }
    }
    // ---- backtrack by value comparison (simd_abpoa_align.c:248-377).  The walk is sequential, but every
    // step's loads (own row, all predecessors) are issued together by the warp, row metadata comes from a
    // shared-memory window filled 32 rows at a time, and the DP values of the rows ahead are prefetched into L2.
    int n_cig = 0, err = TH_OK;
    {
        enum { M_OP = 1, E1_OP = 2, E2_OP = 4, E_OP = 6, F1_OP = 8, F2_OP = 16, F_OP = 24, ALL_OP = 31 };
        int i = bi, j = bj, cur_op = ALL_OP;
        uint32_t *cg = w.cigar; int32_t *cq = w.cigq;
        for (int t = lane; t < qlen - bj; t += 32) { cg[t] = 1; cq[t] = qlen - 1 - t; } // unaligned query tail
        if (bj < qlen) n_cig = qlen - bj;
        int wlo = n, whi = -1; // rows [wlo, whi] are in the window
        while (i > 0 && j > 0) {
            if (i < wlo || (i < wlo + 32 && wlo > 0)) {
                const int top = i < wlo ? i + 1 : wlo; // load rows [top-32, top)
                if (i < wlo) whi = i;
                const int r = top - 32 + lane;
                if (r >= 0) {
                    const int4 m = w.rmeta[r];
                    sm.desc[r & (POA_RING - 1)] = w.rdesc[r]; sm.meta[r & (POA_RING - 1)] = m;
                    const int rw = m.z - m.y + 1;
                    // the path crosses a row close to that row's maximum (the column that steered the band); the graph holds
                    // several nodes per query column, so extrapolating j along the row index would drift off within a few rows.
                    // (Prefetching whole rows with cp.async.bulk.prefetch.L2 was measured: +25 GB of DRAM reads per 8192 reads
                    // and no shorter backtrack, so only the sectors around the predicted crossing are requested.)
                    int jp = m.w > 0 ? m.w - 1 : j - (i - r); jp = min(max(jp, m.y), m.z);
                    const int c0 = max(jp - 7, m.y) - m.y, c1 = max(jp - 2, m.y) - m.y, c2 = min(jp + 4, m.z) - m.y;
                    const uint32_t *Rr = A32 + (m.x >> 1);
                    prefetch_l2(Rr + 4 * (c0 >> 1)); prefetch_l2(Rr + 4 * (c1 >> 1)); prefetch_l2(Rr + 4 * (c2 >> 1)); prefetch_l2(Rr + 2 * rw + ((jp - m.y) >> 1));
                }
                wlo = max(0, top - 32); whi = min(whi, wlo + POA_RING - 1);
                __syncwarp();
            }
            const int4 d = sm.desc[i & (POA_RING - 1)], mi = sm.meta[i & (POA_RING - 1)];
            const int np = d.y & 1023, vb = (d.y >> 10) & 7, v = d.y >> 13;
            if (np > 32) { err = TH_ERR_CAP; break; }
            const int qb = query[j - 1];
            const int s = (qb < 4 && vb < 4) ? (qb == vb ? mat : -mis) : 0;
            const int ib = mi.y, iw = mi.z - mi.y + 1;
            const int c = j - ib, q = c >> 1, odd = c & 1;
            const uint32_t rb = (uint32_t)(mi.x >> 1);
            // lane p holds predecessor p.  Most steps are matches: test those first, with the two loads they need.
            int pi = 0; bool in1 = false, in0 = false, podd = false; uint32_t pbw = 0;
            if (lane < np) {
                pi = lane == 0 ? d.x : (np == 2 ? d.w : w.plist[d.w + lane]);
                const int4 pm = pi >= wlo ? sm.meta[pi & (POA_RING - 1)] : w.rmeta[pi];
                const int cp = j - pm.y;
                podd = cp & 1; pbw = (uint32_t)(pm.x >> 1) + 4 * (cp >> 1);
                in1 = cp >= 1 && j - 1 <= pm.z; in0 = cp >= 0 && j <= pm.z;
            }
            const int hij = s16_at(A32[rb + 4 * q], odd);
            if (cur_op & M_OP) {
                uint32_t hpw = 0; // the H word that holds column j-1 of the predecessor
                if (in1) hpw = A32[podd ? pbw : pbw - 4];
                const int a = podd ? lo16(hpw) : hi16(hpw);
                const unsigned mm = __ballot_sync(TH_FULL, in1 && a + s == hij);
                if (mm) {
                    const int f = __ffs(mm) - 1;
                    if (lane == 0) { cg[n_cig] = ((uint32_t)v << 2) | 0; cq[n_cig] = j - 1; }
                    ++n_cig; cur_op = ALL_OP; i = __shfl_sync(TH_FULL, pi, f); --j;
                    continue;
                }
            }
            // not a match: the other states of this cell, of the cell to its left, and of the predecessors at column j
            const uint4 r0 = *reinterpret_cast<const uint4 *>(A32 + rb + 4 * q);
            const uint32_t g0 = A32[rb + 2 * iw + q];
            uint4 r1 = r0; uint32_t g1 = g0;
            if (!odd && c >= 2) { r1 = *reinterpret_cast<const uint4 *>(A32 + rb + 4 * (q - 1)); g1 = A32[rb + 2 * iw + q - 1]; }
            const int e1ij = s16_at(r0.y, odd), e2ij = s16_at(r0.z, odd), f1 = s16_at(r0.w, odd), f2 = s16_at(g0, odd);
            const int hm1 = s16_at(r1.x, !odd), f1m1 = s16_at(r1.w, !odd), f2m1 = s16_at(g1, !odd); // column j-1 (used only when c >= 1)
            int b = 0, x1 = 0, x2 = 0;
            if (in0) { const uint4 rp = *reinterpret_cast<const uint4 *>(A32 + pbw); b = s16_at(rp.x, podd); x1 = s16_at(rp.y, podd); x2 = s16_at(rp.z, podd); }
            if (cur_op & E_OP) {
                const bool ok1 = (cur_op & E1_OP) && in0 && ((cur_op & M_OP) ? (hij == x1) : (e1ij == x1 - e1));
                const bool ok2 = (cur_op & E2_OP) && in0 && ((cur_op & M_OP) ? (hij == x2) : (e2ij == x2 - e2));
                const unsigned em = __ballot_sync(TH_FULL, ok1 || ok2);
                if (em) {
                    const int f = __ffs(em) - 1;
                    int nop;
                    if (ok1) nop = (b - oe1 == x1) ? (M_OP | F_OP) : E1_OP;
                    else nop = (b - oe2 == x2) ? (M_OP | F_OP) : E2_OP;
                    cur_op = __shfl_sync(TH_FULL, nop, f);
                    if (lane == 0) { cg[n_cig] = ((uint32_t)v << 2) | 2; cq[n_cig] = j - 1; }
                    ++n_cig; i = __shfl_sync(TH_FULL, pi, f);
                    continue;
                }
            }
            if (cur_op & F_OP) {
                if (c < 1) { err = TH_ERR_BACKTRACK; break; }
                bool hit = false;
                if (cur_op & F1_OP) {
                    if (!(cur_op & M_OP) || hij == f1) {
                        if (hm1 - oe1 == f1) { cur_op = M_OP | E_OP; hit = true; }
                        else if (f1m1 - e1 == f1) { cur_op = F1_OP; hit = true; }
                        else { err = TH_ERR_BACKTRACK; break; }
                    }
                }
                if (!hit && (cur_op & F2_OP)) {
                    if (!(cur_op & M_OP) || hij == f2) {
                        if (hm1 - oe2 == f2) { cur_op = M_OP | E_OP; hit = true; }
                        else if (f2m1 - e2 == f2) { cur_op = F2_OP; hit = true; }
                        else { err = TH_ERR_BACKTRACK; break; }
                    }
                }
                if (lane == 0) { cg[n_cig] = 1; cq[n_cig] = j - 1; }
                ++n_cig; --j;
                continue;
            }
            err = TH_ERR_BACKTRACK; break;
        }
        if (err != TH_OK) return err;
        for (int t = lane; t < j; t += 32) { cg[n_cig + t] = 1; cq[n_cig + t] = j - 1 - t; } // unaligned query head
        if (j > 0) n_cig += j;
        __syncwarp();
    }
    PH(2);
    // ---- merge the alignment into the graph (abpoa_graph.c:1218-1284), 32 path steps at a time ----------
    // A path visits every node and every aligned group at most once, so all steps touch distinct adjacency lists
    // and groups; ids of new nodes/edges are creation-ordered prefix sums, exactly what the sequential walk yields.
    int n_ev = 0;
    {
        const int node_n0 = node_n;
        int last_id = 0, last_new = 0, pend_lo = 0; // events [pend_lo, n_ev) wait for the next match column
        for (int k0 = 0; k0 < n_cig; k0 += 32) {
            const int k = k0 + lane;
            int op = 2, v = 0, q = 0;
            if (k < n_cig) { const uint32_t cv = w.cigar[n_cig - 1 - k]; op = cv & 3; v = (int)(cv >> 2); q = w.cigq[n_cig - 1 - k]; }
            const bool prod = op != 2;
            int tgt = -1, isnew = 0, bf = 0, bl = 0, an = 0, al[4] = {0, 0, 0, 0};
            uint8_t qb = 0;
            if (prod) qb = query[q];
            if (op == 0) {
                bf = w.n2i[v]; bl = bf; an = w.aln_n[v];
                int aid = -1;
                for (int a = 0; a < an; ++a) {
                    al[a] = w.aln[v * 4 + a];
                    const int x = w.n2i[al[a]]; bf = min(bf, x); bl = max(bl, x);
                    if (aid < 0 && w.base[al[a]] == qb) aid = al[a];
                }
                if (w.base[v] == qb) tgt = v; else if (aid >= 0) tgt = aid; else isnew = 1;
            } else if (op == 1) isnew = 1;
            const unsigned newm = __ballot_sync(TH_FULL, isnew), prodm = __ballot_sync(TH_FULL, prod), mm = __ballot_sync(TH_FULL, op == 0);
            const unsigned below = (1u << lane) - 1;
            const int ev = n_ev + __popc(newm & below); // event id == rank of the new node
            if (isnew) tgt = node_n0 + ev;
            // previous producing step
            int from = last_id, from_new = last_new;
            { const unsigned lower = prodm & below; const int src = lower ? 31 - __clz(lower) : 0;
              const int pt = __shfl_sync(TH_FULL, tgt, src), pnw = __shfl_sync(TH_FULL, isnew, src);
              if (lower) { from = pt; from_new = pnw; } }
            // new nodes first (their adjacency heads must exist before edges are linked)
            if (isnew) {
                if (tgt >= w.ncap) err = TH_ERR_CAP;
                else { w.base[tgt] = qb; w.out_head[tgt] = w.out_tail[tgt] = w.in_head[tgt] = w.in_tail[tgt] = -1; w.aln_n[tgt] = 0; w.ev_node[ev] = tgt; }
            }
            if (__any_sync(TH_FULL, err != TH_OK)) return TH_ERR_CAP;
            if (isnew && op == 0) { // abpoa_add_graph_aligned_node (:1036-1044): all-pairs with the old group
                if (an >= 4) err = TH_ERR_CAP; // a column holds at most 5 distinct codes (ACGT + N): 4 aligned nodes per node
                else {
                    for (int a = 0; a < an; ++a) { const int y = al[a]; w.aln[y * 4 + w.aln_n[y]] = tgt; w.aln_n[y] += 1; w.aln[tgt * 4 + a] = y; }
                    w.aln[v * 4 + an] = tgt; w.aln_n[v] = an + 1; w.aln[tgt * 4 + an] = v; w.aln_n[tgt] = an + 1;
                    w.ev_anchor[ev] = bl + 1;
                }
            }
            if (__any_sync(TH_FULL, err != TH_OK)) return TH_ERR_CAP;
            // insertion events take the first index of the next match column's group
            if (mm) {
                const int fm = __ffs(mm) - 1, bff = __shfl_sync(TH_FULL, bf, fm);
                const int hi_ev = n_ev + __popc(newm & ((1u << fm) - 1)); // events created before that step
                for (int e = pend_lo + lane; e < hi_ev; e += 32) w.ev_anchor[e] = bff;
            }
            {
                const unsigned higher = mm & ~(below | (1u << lane));
                const int src = higher ? __ffs(higher) - 1 : 0;
                const int bfn = __shfl_sync(TH_FULL, bf, src);
                if (op == 1 && higher) w.ev_anchor[ev] = bfn;
            }
            { // events after the last match column of this batch stay pending
                const int lastm = mm ? 31 - __clz(mm) : -1;
                if (mm) pend_lo = n_ev + __popc(newm & ((2u << lastm) - 1));
            }
            __syncwarp();
            // edges (abpoa_add_graph_edge :1063-1106): weight + 1 on an existing edge, else append to both lists
            int found = -1;
            if (prod && !from_new && !isnew)
                for (int e = w.out_head[from]; e >= 0; e = w.e_no[e]) if (w.e_to[e] == tgt) { found = e; break; }
            const bool mk = prod && found < 0;
            const unsigned mkm = __ballot_sync(TH_FULL, mk);
            if (found >= 0) w.e_w[found] += 1;
            if (mk) {
                const int e = edge_n + __popc(mkm & below);
                w.e_to[e] = tgt; w.e_from[e] = from; w.e_w[e] = 1; w.e_no[e] = -1; w.e_ni[e] = -1;
                const int ot = w.out_tail[from];
                if (ot < 0) w.out_head[from] = e; else w.e_no[ot] = e;
                w.out_tail[from] = e;
                const int it = w.in_tail[tgt];
                if (it < 0) w.in_head[tgt] = e; else w.e_ni[it] = e;
                w.in_tail[tgt] = e;
            }
            edge_n += __popc(mkm);
            n_ev += __popc(newm);
            if (prodm) { const int src = 31 - __clz(prodm); last_id = __shfl_sync(TH_FULL, tgt, src); last_new = __shfl_sync(TH_FULL, isnew, src); }
            __syncwarp();
        }
        node_n = node_n0 + n_ev;
        if (lane == 0) g_add_edge(w, edge_n, last_id, 1, !last_new);
        edge_n = __shfl_sync(TH_FULL, edge_n, 0);
        for (int e = pend_lo + lane; e < n_ev; e += 32) w.ev_anchor[e] = n - 1; // before the sink
    }
    __syncwarp();
    PH(3);
    // new order: old node at index i moves to i + #(events with anchor <= i); event e lands at anchor_e + e
    for (int i = lane; i < n; i += 32) {
        int lo = 0, hi = n_ev;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (w.ev_anchor[mid] <= i) lo = mid + 1; else hi = mid; }
        w.ord2[i + lo] = w.ord[i];
    }
    for (int e = lane; e < n_ev; e += 32) w.ord2[w.ev_anchor[e] + e] = w.ev_node[e];
    __syncwarp();
    if (lane == 0) { int32_t *t = w.ord; w.ord = w.ord2; w.ord2 = t; }
    __syncwarp();
    PH(4);
#undef PH
    return TH_OK;
}

// heaviest-column consensus (abpoa_graph.c:279-359, 604-648, 467-478).  All lanes call; returns cons_len.
// The DFS that assigns MSA ranks is order dependent (stack discipline, abpoa_graph.c:279-339) and stays on lane 0;
// in-degrees, column weights and the column vote are data parallel.
__device__ int poa_consensus(PoaWs &w, int node_n, int n_seq, uint8_t *cons, int32_t *cov) {
    const int lane = lane_id();
    int32_t *deg = w.n2i, *stk = w.ord2, *rank = w.ri;
    int32_t *rcw = reinterpret_cast<int32_t *>(w.arena); // 5 x msa_l weights, then 5 x msa_l node ids
    for (int i = lane; i < node_n; i += 32) { int d = 0; for (int e = w.in_head[i]; e >= 0; e = w.e_ni[e]) ++d; deg[i] = d; }
    __syncwarp();
    int msa_l = 0;
    if (lane == 0) {
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
                    const int an = w.aln_n[o];
                    for (int a = 0; a < an; ++a) if (deg[w.aln[o * 4 + a]] != 0) { ok = false; break; }
                    if (!ok) continue;
                    stk[sp++] = o; rank[o] = -1;
                    for (int a = 0; a < an; ++a) { const int x = w.aln[o * 4 + a]; stk[sp++] = x; rank[x] = -1; }
                }
            }
        }
        msa_l = rank[1] - 1;
    }
    msa_l = __shfl_sync(TH_FULL, msa_l, 0);
    __syncwarp();
    if (msa_l <= 0) return 0;
    if ((uint64_t)msa_l * 10 * 2 > w.arena_cap) return -1;
    int32_t *nodeid = rcw + 5 * (size_t)msa_l;
    for (int i = lane; i < 5 * msa_l; i += 32) { rcw[i] = 0; nodeid[i] = 0; }
    __syncwarp();
    // abpoa_set_row_column_weight; popcount(read_ids) == sum of out weights.  A column is one aligned group and a
    // group holds one node per base, so every (column, base) slot has a single writer.
    for (int i = 2 + lane; i < node_n; i += 32) {
        int rk = rank[i];
        for (int a = 0; a < w.aln_n[i]; ++a) rk = max(rk, rank[w.aln[i * 4 + a]]);
        int wsum = 0;
        for (int e = w.out_head[i]; e >= 0; e = w.e_no[e]) wsum += w.e_w[e];
        const int b = w.base[i] > 4 ? 4 : w.base[i];
        rcw[(rk - 1) * 5 + b] += wsum;
        nodeid[(rk - 1) * 5 + b] = i;
    }
    __syncwarp();
    int cons_i = 0;
    for (int c0 = 0; c0 < msa_l; c0 += 32) { // abpoa_heaviest_column_consensus (:604-629): first heaviest base, kept iff max_w >= gap weight
        const int c = c0 + lane;
        int max_w = 0, max_base = 5, gap_w = n_seq; bool sel = false;
        if (c < msa_l) {
            for (int b = 0; b < 4; ++b) { const int x = rcw[c * 5 + b]; if (x > max_w) { max_base = b; max_w = x; } gap_w -= x; }
            sel = max_w >= gap_w && max_base < 5;
        }
        const unsigned m = __ballot_sync(TH_FULL, sel);
        if (sel) { const int pos = cons_i + __popc(m & ((1u << lane) - 1)); cons[pos] = w.base[nodeid[c * 5 + max_base]]; cov[pos] = max_w; }
        cons_i += __popc(m);
    }
    return cons_i;
}

// persistent warps pull tasks from an atomic counter
#ifndef POA_MIN_BLOCKS
#define POA_MIN_BLOCKS 8   // 64 registers: 32 warps per SM.  Every phase of the kernel is a dependent chain, so resident warps are what hides latency
#endif
__global__ void __launch_bounds__(POA_WARPS * 32, POA_MIN_BLOCKS)
poa_kernel(DevParams P, int n_tasks, const PoaTask *__restrict__ tasks, const int32_t *__restrict__ task_order,
           const int32_t *__restrict__ u_start, const int32_t *__restrict__ u_len, const uint8_t *__restrict__ bseq,
           uint8_t *slabs, size_t slab_bytes, int *task_counter,
           uint8_t *__restrict__ cons_base, int32_t *__restrict__ cons_cov, int32_t *__restrict__ cons_len,
           int32_t *__restrict__ task_status, unsigned long long *__restrict__ stat_cells, unsigned long long *__restrict__ stat_rows,
           unsigned long long *__restrict__ stat_phase) {
    __shared__ PoaSmem s_mem[POA_WARPS];
    __shared__ PoaWs s_ws[POA_WARPS]; // workspace pointers live in shared memory: one LDS instead of re-deriving them under register pressure
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * POA_WARPS + wib;
    uint8_t *slab = slabs + (size_t)gw * slab_bytes;
    unsigned long long cells = 0, rows = 0;
#ifdef POA_PROFILE
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#else
    long long *ph = nullptr;
#endif
    while (true) {
#ifdef POA_PROFILE
        long long t_task = clock64();
#endif
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
        PoaWs &w = s_ws[wib];
        __syncwarp();
        if (lane == 0) poa_carve(w, slab, slab_bytes, T.ncap, T.qmax, T.n_seqs);
        __syncwarp();
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
        {
            const uint8_t *q = rseq + u_start[T.unit_off + s]; const int ql = u_len[T.unit_off + s];
            if (!P.affine) err = P.pn == 16 ? poa_add_sequence<4, false>(w, P, q, ql, node_n, edge_n, s_mem[wib], cells, rows, ph)
                                            : poa_add_sequence<3, false>(w, P, q, ql, node_n, edge_n, s_mem[wib], cells, rows, ph);
            else err = P.pn == 16 ? poa_add_sequence<4, true>(w, P, q, ql, node_n, edge_n, s_mem[wib], cells, rows, ph)
                                  : poa_add_sequence<3, true>(w, P, q, ql, node_n, edge_n, s_mem[wib], cells, rows, ph);
        }
        int cl = 0;
#ifdef POA_PROFILE
        long long t_c0 = clock64();
#endif
        if (err == TH_OK) {
            cl = poa_consensus(w, node_n, T.n_seqs, cons, cov);
            if (cl < 0) { err = TH_ERR_ARENA; cl = 0; }
        }
        if (lane == 0) { cons_len[t] = cl; task_status[t] = err; }
        __syncwarp();
#ifdef POA_PROFILE
        { long long t_ = clock64(); ph[5] += t_ - t_c0; ph[6] += t_ - t_task; }
#endif
    }
    if (lane == 0 && cells) {
        atomicAdd(stat_cells, cells); atomicAdd(stat_rows, rows);
#ifdef POA_PROFILE
        for (int k = 0; k < 7; ++k) atomicAdd(stat_phase + k, (unsigned long long)ph[k]);
#endif
    }
}
