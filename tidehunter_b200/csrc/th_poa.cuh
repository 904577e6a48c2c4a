// th_poa.cuh -- adaptive-banded partial-order alignment + heaviest-column consensus on lane groups.
//
// Replaces (reference, /root/reference): src/abpoa_cons.c:30-120 (abpoa_gen_cons) and the abPOA calls
// below it: abPOA/src/abpoa_align.c:293-411 (abpoa_msa / abpoa_poa), abPOA/src/simd_abpoa_align.c
// (convex-gap banded DP :835-958, affine :739-833, row arg-max :991-1015, best cell :976-989, backtrack :248-377),
// abPOA/src/abpoa_graph.c (:197-238 max_remain, :1020-1124 node/edge/aligned, :1218-1288 add alignment,
// :279-359 + :604-648 heaviest-column consensus).
//
// Mapping.  A task (one consensus) is worked on by a GROUP of LPT lanes: LPT = 16 puts two tasks in one warp, LPT = 32
// one.  A lane owns four adjacent columns of a row chunk (two s16x2 registers), so a chunk is 4 LPT columns: the usual
// band of a 1 kb unit (48-80 columns) is one chunk of a 16-lane group.  The kernel is bound by instructions issued per
// row (band bookkeeping, address arithmetic, scan rounds), not by cell throughput: with two groups in a warp every
// one of those instructions serves two rows.  Both groups of a warp run the same instruction stream ("lockstep"):
// every loop that contains a warp-synchronous operation runs while ANY group needs it, a group that has nothing to do
// executes it with its loads and stores predicated off, and the cross-lane operations are full-warp shuffles of
// width LPT.  Groups fetch their tasks independently, so a finished group starts its next task while its neighbour
// is still in the middle of one; only the phases of one alignment (setup, rows, backtrack, merge) are shared.
//
// What is kept bit-exact and why it is enough:
//  * DP values are 16-bit wrapping integers exactly as the reference's AVX2 int16 path (DPX / video SIMD ops).
//  * Band edges are rounded to whole emulated SIMD vectors of `pn` lanes, and the row arg-max that steers the
//    adaptive band uses the reference's lane-ordered tie-break.
//  * Rows are processed in a topological order that keeps aligned-node groups contiguous; it is NOT the reference's
//    BFS order.  The alignment does not depend on which topological order is used: a row's band and values depend
//    only on its predecessors' rows, the backtrack walks in_id order, and max_remain is a function of graph
//    structure only.  The order is maintained incrementally (new nodes are merged in right before the next existing
//    node of the alignment path), which replaces the per-sequence BFS (abpoa_graph.c:150-195) by a parallel merge.
//  * read-id bitsets are not stored: popcount(read_ids) of a node equals the sum of its out-edge weights.
//  * The consensus DFS (msa rank) and column vote are literal.
//
// DP storage: 7 bytes per banded cell.  A row of W columns owns three int16 planes H | E1 | E2 (what successor rows and
// the backtrack's deletion tests read) and one byte per cell that says whether the cell's value is a match / mismatch
// step and from which predecessor (bit 0: H == max_p H[p][j-1] + s; bits 1..4: the first predecessor reaching that
// maximum, which is the first one the reference's value comparison accepts).  F1 / F2 are NOT stored: the backtrack only
// needs them at the few cells where an insertion is taken, and recomputes that row chunk from the predecessors' planes
// (same code as the forward pass).  The last POA_NRING rows (when they fit one chunk) also stay in shared memory, where
// the next rows read them.  The backtrack stages, for a 64-row window, each row's predecessors and the codes of 32
// columns around its arg-max (where the path crosses the row) in shared memory: a match step (85 % of the path) is two
// shared-memory loads and no cross-lane operation; everything else takes the general value-comparison step.
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
#define POA_MAXPRE 32     // most predecessors a row may have in the 32-lane kernel (16 in the 16-lane kernel): more -> TH_ERR_CAP
#define POA_NEGP 0x80008000u
#define POA_NRING 4       // rows kept in shared memory for their successors (power of two)
#define POA_WIN 64        // rows of metadata kept in shared memory (power of two)

// Row descriptor (static per alignment, by row index):  x = first predecessor row (-1: none),
//   y = np | base << 10 | node << 13,  z = qlen - max_remain term of the band centre,
//   w = second predecessor row.  rdesc2 holds predecessor rows 2..5; rows with more keep all of them in plist at rdesc2.x
//   (then rdesc2.y = predecessor 2 ... is not used: np > 6 reads plist[rdesc2.x + p]).
// Row metadata (written when the row is computed):  x = arena offset (int16 units), y = first column, z = last column,
//   w = max_i + 1 of the row (what the reference scatters into max_pos_left/right of the successors).
// Node records (kept up to date by the merge, so that the per-alignment setup walks no adjacency list):
//   nrec[v] = in-degree, then the first three in-neighbours in in_id order;
//   brec[v] = heaviest out-edge (the first one of maximum weight, abpoa_graph.c:216-226): target node, weight, edge id.
struct PoaWs {
    int4 *rdesc, *rdesc2, *rmeta, *nrec, *brec;
    int32_t *out_head, *out_tail, *in_head, *in_tail, *aln_n, *aln, *n2i, *ri, *hs;
    int32_t *e_to, *e_from, *e_w, *e_no, *e_ni, *plist;
    int32_t *ord, *ord2, *ev_anchor, *ev_node, *hi_idx;
    uint32_t *q8g; uint32_t *cigar; int32_t *cigq; uint8_t *base;
    int16_t *arena; uint32_t arena_cap;
    int32_t ncap;
};

__host__ __device__ inline size_t poa_fixed_bytes(int ncap, int qmax, int nseq) {
    size_t ecap = (size_t)ncap + nseq + 2;
    size_t b = 0;
    b += (size_t)ncap * 16 * 5;                        // rdesc, rdesc2, rmeta, nrec, brec
    b += (size_t)ncap * 4 * (9 + 3);                   // 9 node arrays (aln counts as 4) = 12 x int32
    b += ecap * 4 * 6;                                 // 5 edge arrays + plist
    b += (size_t)ncap * 4 * 3;                         // ord, ord2, hi_idx
    b += (size_t)(qmax + 2) * 4 * 2;                   // events
    b += (size_t)(qmax + 512) + 8;                     // query bytes (when they do not fit shared memory)
    b += (size_t)(qmax + ncap + 8) * 4 * 2;            // cigar, cigq
    b += (size_t)ncap + 64;                            // base
    return (b + 4095) & ~(size_t)4095;
}

// slab sizes (bytes) of one task: graph arrays + the DP arena of ONE alignment (recycled for every unit): rows <= nodes, three
// int16 planes and a code byte per banded cell.  Typical: the graph holds <= ~2.5 units worth of nodes; the adaptive band is
// 2w+1 columns around the predecessors' row maxima, rounded to whole vectors (measured mean ~61 columns on 1 kb units).
__host__ __device__ inline unsigned long long poa_slab_need(int ncap, int qmax, int n_seqs, bool full) {
    const unsigned long long fixed = poa_fixed_bytes(ncap, qmax, n_seqs);
    const unsigned long long fullb = (unsigned long long)ncap * (((unsigned long long)qmax + 64) * 7 + 16);
    if (full) { // the wide pass: five int32 planes per cell; rows are banded (2 w + 1 columns around the predecessors' maxima, plus the
        // spread of those maxima), so full width is only needed for short queries
        const unsigned long long wd = (unsigned long long)qmax + 64, band = 4ull * (10 + qmax / 100) + 512;
        return fixed + (unsigned long long)ncap * (wd < band ? wd : band) * 20 + 4096;
    }
    const int wband = 10 + qmax / 100;
    unsigned long long rows_typ = (unsigned long long)qmax * 5 / 2 + 64; if (rows_typ > (unsigned long long)ncap) rows_typ = ncap;
    unsigned long long width_typ = 2ull * wband + 64; if (width_typ > (unsigned long long)qmax + 64) width_typ = (unsigned long long)qmax + 64;
    unsigned long long typ = rows_typ * (width_typ * 7 + 16); if (typ < (1ull << 20)) typ = 1ull << 20;
    return fixed + (typ < fullb ? typ : fullb) + 4096;
}

__device__ inline void poa_carve(PoaWs &w, uint8_t *slab, size_t slab_bytes, int ncap, int qmax, int nseq) {
    size_t ecap = (size_t)ncap + nseq + 2;
    w.ncap = ncap;
    w.rdesc = reinterpret_cast<int4 *>(slab); w.rdesc2 = w.rdesc + ncap; w.rmeta = w.rdesc2 + ncap; w.nrec = w.rmeta + ncap; w.brec = w.nrec + ncap;
    int32_t *p = reinterpret_cast<int32_t *>(w.brec + ncap);
    w.out_head = p; p += ncap; w.out_tail = p; p += ncap; w.in_head = p; p += ncap; w.in_tail = p; p += ncap;
    w.aln_n = p; p += ncap; w.aln = p; p += 4 * (size_t)ncap; w.n2i = p; p += ncap; w.ri = p; p += ncap; w.hs = p; p += ncap;
    w.e_to = p; p += ecap; w.e_from = p; p += ecap; w.e_w = p; p += ecap; w.e_no = p; p += ecap; w.e_ni = p; p += ecap; w.plist = p; p += ecap;
    w.ord = p; p += ncap; w.ord2 = p; p += ncap; w.hi_idx = p; p += ncap;
    w.ev_anchor = p; p += qmax + 2; w.ev_node = p; p += qmax + 2;
    w.q8g = reinterpret_cast<uint32_t *>(p); p += (qmax + 512 + 3) / 4 + 1;
    w.cigar = reinterpret_cast<uint32_t *>(p); p += qmax + ncap + 8;
    w.cigq = p; p += qmax + ncap + 8;
    w.base = reinterpret_cast<uint8_t *>(p);
    size_t fixed = poa_fixed_bytes(ncap, qmax, nseq);
    w.arena = reinterpret_cast<int16_t *>(slab + fixed);
    size_t ab = slab_bytes > fixed ? slab_bytes - fixed : 0;
    size_t ne = ab / 2; if (ne > 0x7ffffff0ull) ne = 0x7ffffff0ull; // row offsets are kept in an int; the sign bit marks "in the shared-memory ring"
    w.arena_cap = (uint32_t)ne;
}

// int16 units a row of `width` columns takes in the arena: three planes and the byte codes, rounded to 16 bytes
__host__ __device__ __forceinline__ uint32_t poa_row_size(int width) { return ((uint32_t)(3 * width + (width >> 1)) + 7u) & ~7u; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
// asynchronous 16-byte copy global -> shared (LDGSTS): in flight while the warp goes on
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ int s16_at(uint32_t word, int odd) { return odd ? hi16(word) : lo16(word); }

// a new edge `from -> to` with id e enters the node records (one lane per `to` and per `from` at a time)
__device__ __forceinline__ void g_note_in(PoaWs &w, int to, int from) {
    int32_t *r = reinterpret_cast<int32_t *>(w.nrec + to);
    const int k = r[0];
    if (k < 3) r[1 + k] = from;
    r[0] = k + 1;
}
// edge e out of `from` now has weight wt: it becomes the heaviest one if it beats the current one, or ties with it
// and comes earlier in the out list (edge ids grow along a node's out list)
__device__ __forceinline__ void g_note_out(PoaWs &w, int from, int to, int e, int wt) {
    int32_t *r = reinterpret_cast<int32_t *>(w.brec + from);
    const int bw = r[1], be = r[2];
    if (wt > bw || (wt == bw && e <= be)) { r[0] = to; r[1] = wt; r[2] = e; }
}
// row of predecessor p of a row with descriptors d, d2 (np > 6: all of them are in plist)
__device__ __forceinline__ int poa_pred_row(const int4 d, const int4 d2, const int32_t *plist, const int np, const int p) {
    if (np > 6) return plist[d2.x + p];
    return p == 0 ? d.x : p == 1 ? d.w : p == 2 ? d2.x : p == 3 ? d2.y : p == 4 ? d2.z : d2.w;
}
// graph edit used for the final edge into the sink (one lane of the group)
__device__ inline void g_add_edge(PoaWs &w, int &edge_n, int from, int to, bool check) {
    if (check) {
        for (int e = w.out_head[from]; e >= 0; e = w.e_no[e])
            if (w.e_to[e] == to) { const int wt = w.e_w[e] + 1; w.e_w[e] = wt; g_note_out(w, from, to, e, wt); return; }
    }
    int e = edge_n++;
    g_note_in(w, to, from); g_note_out(w, from, to, e, 1);
    w.e_to[e] = to; w.e_from[e] = from; w.e_w[e] = 1; w.e_no[e] = -1; w.e_ni[e] = -1;
    if (w.out_tail[from] < 0) w.out_head[from] = e; else w.e_no[w.out_tail[from]] = e;
    w.out_tail[from] = e;
    if (w.in_tail[to] < 0) w.in_head[to] = e; else w.e_ni[w.in_tail[to]] = e;
    w.in_tail[to] = e;
}

// ---- lane groups --------------------------------------------------------------------------------------------------
// Every member function is a warp-synchronous operation over ALL 32 lanes (full mask); the group only sees its own part
// of the result.  They must be called from code all lanes of the warp execute together.
template <int LPT> struct PoaG {
    int gl, gofs;
    __device__ __forceinline__ PoaG() { const int l = threadIdx.x & 31; gl = l & (LPT - 1); gofs = l & ~(LPT - 1); }
    __device__ __forceinline__ unsigned ballot(bool p) const {
        const unsigned b = __ballot_sync(TH_FULL, p);
        if (LPT == 32) return b;
        return (b >> gofs) & 0xffffu;
    }
    __device__ __forceinline__ bool any(bool p) const { return ballot(p) != 0; }
    template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(TH_FULL, v, src, LPT); }
    template <class T> __device__ __forceinline__ T shfl_up(T v, int d) const { return __shfl_up_sync(TH_FULL, v, d, LPT); }
    __device__ __forceinline__ int rmax(int v) const {
        if (LPT == 32) return __reduce_max_sync(TH_FULL, v);
        const int a = __reduce_max_sync(TH_FULL, gofs ? INT_MIN : v), b = __reduce_max_sync(TH_FULL, gofs ? v : INT_MIN);
        return gofs ? b : a;
    }
};

// One row of the backtrack window: hdr = {first predecessor row, second predecessor row or offset into plist,
// np | base << 10 | node << 13 (the row descriptor's y), first staged column}, code = the M-codes of 32 columns from there.
struct PoaBt { int4 hdr; uint4 code[2]; };
template <int LPT> struct PoaSmem {
    static constexpr int CW = LPT * 4;               // columns per chunk
    static constexpr int RINGC = CW + 16;            // widest row a ring slot holds (columns)
    static constexpr int RINGW = 3 * RINGC / 2;      // words per ring slot (three planes)
    static constexpr int Q4N = LPT == 16 ? 320 : 1536; // 4-column groups of query codes kept in shared memory (16 bits each)
    union {
        struct { int4 meta[POA_WIN]; uint32_t ring[POA_NRING * RINGW]; } f; // forward pass: row metadata, recent rows
        PoaBt bt[POA_WIN];                                                   // backtrack: window of rows
    } u;
    int4 pre[LPT];                                    // metadata of predecessors 1.. of the row being computed (np <= LPT)
    uint16_t q4[Q4N];
    int plist_n;
    PoaWs ws;
};

// Per-lane constants of the row chunk: for the lane's registers r = 0, 1 (columns c = 4 gl + 2 r, c + 1 of the chunk)
//   C1 = (e1 c - oe1, e1 (c + 1) - oe1), C2 likewise: what turns H[j - 1] into the scan's G;  NJ1 / NJ2 = (-e c, -e (c + 1)): G back to F;
//   KLO / KHI: the tie-break bits of the row arg-max key for the even / odd column.
// They depend on the lane and the options only; a table in shared memory is filled once per block and read where needed
// (a volatile load: held in registers across the row loop they would push its other state out).
struct PoaLaneK { uint32_t C1[2], C2[2], NJ1[2], NJ2[2], KLO[2], KHI[2]; };
__device__ __forceinline__ uint4 lds128v(uint32_t saddr) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr)); return v;
}
__device__ inline void poa_fill_lane_k(PoaLaneK *tab, int lpt, const DevParams &P) {
    const int lam_bits = P.pn - 1;
    for (int gl = threadIdx.x; gl < lpt; gl += blockDim.x) {
        PoaLaneK k;
        for (int r = 0; r < 2; ++r) {
            const int c = 4 * gl + 2 * r;
            k.C1[r] = pk(P.e1 * c - P.oe1, P.e1 * (c + 1) - P.oe1); k.C2[r] = pk(P.e2 * c - P.oe2, P.e2 * (c + 1) - P.oe2);
            k.NJ1[r] = pk(-P.e1 * c, -P.e1 * (c + 1)); k.NJ2[r] = pk(-P.e2 * c, -P.e2 * (c + 1));
            k.KLO[r] = ((uint32_t)(lam_bits - (c & lam_bits)) << 12) | 0xfffu; k.KHI[r] = ((uint32_t)(lam_bits - ((c + 1) & lam_bits)) << 12) | 0xfffu;
        }
        tab[gl] = k;
    }
}

// One predecessor row's contribution to the four columns (j0 .. j0 + 3) of a lane: m = H[p][j - 1] (the diagonal), x1 / x2 =
// E1 / E2[p][j].  pm.x = source of the row's planes (sign bit: slot of the shared-memory ring, else arena offset in int16
// units), pm.y / pm.z = its first / last column.  Cells outside the predecessor's band count as inf_min
// (simd_abpoa_align.c:860-905).  Band edges are multiples of 8 columns, j0 of 4: a lane's columns are all inside or all outside.
template <int LPT, bool AFFINE>
__device__ __forceinline__ void poa_pred(const uint32_t *A32, const uint32_t *ring, const int4 pm, const bool act, const int j0, const uint32_t INFP,
                                         uint32_t (&m)[2], uint32_t (&x1)[2], uint32_t (&x2)[2]) {
    const int pb = pm.y, ps = (pm.z - pb + 1) >> 1, wi = (j0 - pb) >> 1;
    // a plane holds ps words; wi and ps are even.  Columns j0 .. j0 + 3 inside the band <=> 0 <= wi < ps;  column j0 - 1 inside <=> 2 <= wi <= ps
    const bool in0 = act && (unsigned)wi < (unsigned)ps, in1 = act && (unsigned)(wi - 2) < (unsigned)ps;
    uint2 h = make_uint2(INFP, INFP), a = h, b = h; uint32_t left = INFP;
    if (pm.x < 0) { // one of the last POA_NRING rows: shared memory
        const uint32_t *Pp = ring + (pm.x & (POA_NRING - 1)) * PoaSmem<LPT>::RINGW + wi;
        if (in0) { h = *reinterpret_cast<const uint2 *>(Pp); a = *reinterpret_cast<const uint2 *>(Pp + ps); if (!AFFINE) b = *reinterpret_cast<const uint2 *>(Pp + 2 * ps); }
        if (in1) left = Pp[-1]; // upper half = H[p][j0 - 1]
    } else {
        const uint32_t *Pp = A32 + ((uint32_t)pm.x >> 1) + wi;
        if (in0) { h = *reinterpret_cast<const uint2 *>(Pp); a = *reinterpret_cast<const uint2 *>(Pp + ps); if (!AFFINE) b = *reinterpret_cast<const uint2 *>(Pp + 2 * ps); }
        if (in1) left = Pp[-1];
    }
    m[0] = __funnelshift_r(left, h.x, 16); m[1] = __funnelshift_r(h.x, h.y, 16);
    x1[0] = a.x; x1[1] = a.y; x2[0] = b.x; x2[1] = b.y;
}

// One chunk (4 LPT columns) of one row: M / E from the predecessors, scores, the F scans, H and the E handed on.
// `act` is false for a group that only keeps the warp company: it loads nothing, and what it computes is discarded.
// pm0 / pre[1..np-1]: predecessor metadata with x = source of the row's planes (sign bit: slot of the shared-memory ring,
// else arena offset in int16 units), y / z = its first / last column.  Cells outside a predecessor's band count as
// inf_min (simd_abpoa_align.c:860-905).  j0 = first column of this lane.  ch0: first chunk of the row.
// AFFINE: abPOA's affine gap mode (gap_open2 == 0, simd_abpoa_ag_dp, simd_abpoa_align.c:739-833).  It is not the convex
// recurrence minus one gap function: an insertion opens from M only (F is built from the row's diagonal values before E
// is folded in), and the E handed to the next row is inf_min wherever F strictly won the cell (SIMDSetIfEqual), so
// insertions and deletions are never adjacent.  The second pair (E2, F2) is held at inf_min.
template <int LPT, bool AFFINE>
__device__ __forceinline__ void poa_chunk(const PoaG<LPT> &g, const DevParams &P, const uint32_t lk, const uint32_t *A32, const uint32_t *ring, const int4 *pre,
                                          const int4 pm0, const int np, const bool act, const int j0, const bool ch0,
                                          const uint32_t carryH, const uint32_t carryF, const uint32_t basew, const uint16_t *q4,
                                          uint32_t (&Hn)[2], uint32_t (&E1o)[2], uint32_t (&E2o)[2], uint32_t (&Fa)[2], uint32_t (&Fb)[2], uint32_t (&Hf)[2], uint32_t &mcode) {
    const uint32_t INFP = P.INFP;
    uint32_t Mx[2], E1x[2], E2x[2];
    poa_pred<LPT, AFFINE>(A32, ring, pm0, act, j0, INFP, Mx, E1x, E2x);
    uint32_t midx[2] = {0u, 0u};   // per column: the first predecessor whose H[p][j-1] is the maximum
#pragma unroll 1
    for (int p = 1; p < np; ++p) { // no warp-synchronous operation in here: groups may run different trip counts
        uint32_t m[2], x1[2], x2[2];
        poa_pred<LPT, AFFINE>(A32, ring, pre[p], act, j0, INFP, m, x1, x2);
        const uint32_t pp = (uint32_t)p * 0x00020002u; // the index, already in its place in the code
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t gt = __vcmpgts2(m[r], Mx[r]);
            midx[r] = (midx[r] & ~gt) | (pp & gt);
            Mx[r] = __vmaxs2(Mx[r], m[r]); E1x[r] = __vmaxs2(E1x[r], x1[r]); if (!AFFINE) E2x[r] = __vmaxs2(E2x[r], x2[r]);
        }
    }
    // scores of the lane's four columns against the node base: q4 holds four 4-bit query codes per 16-bit entry, code of
    // column j = query[j-1] (0..3), or 8 where the score is 0 (column 0, N in the query, columns past the query end);
    // basew = the node base in every byte, 4 for an N node (scores 0 against everything).  x = q ^ base: 0 = match,
    // 1..3 = mismatch, >= 4 = no score.
    uint32_t S[2];
    {
        const uint32_t w4 = act ? (uint32_t)q4[j0 >> 2] : 0x8888u;
        const uint32_t x = prmt(w4 & 0x0f0fu, (w4 >> 4) & 0x0f0fu, 0x5140) ^ basew; // one code per byte
        const uint32_t hb = 0x80808080u - x, vb = x + 0x7c7c7c7cu; // sign bit of each byte: match / no score (bytes are <= 15: no borrows or carries)
        const uint32_t m0 = prmt(hb, 0, 0x9988), m1 = prmt(hb, 0, 0xbbaa), v0 = prmt(vb, 0, 0x9988), v1 = prmt(vb, 0, 0xbbaa);
        S[0] = (P.NEGMIS2 ^ (m0 & P.XMM)) & ~v0; S[1] = (P.NEGMIS2 ^ (m1 & P.XMM)) & ~v1;
    }
    uint32_t Ms[2], Hme[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        Ms[r] = __vadd2(Mx[r], S[r]);
        Hme[r] = AFFINE ? __vmaxs2(Ms[r], E1x[r]) : __vimax3_s16x2(Ms[r], E1x[r], E2x[r]);
        Hf[r] = AFFINE ? Ms[r] : Hme[r];               // what an insertion may open from
    }
    // F1[j] = max_k<=j (A1[k] - e1 (j-k)) with A1[k] = Hf[k-1] - oe1 is evaluated as a prefix MAX of G[k] = A1[k] + e1 (k - j0)
    // over the chunk (no subtraction inside the scan, hence no underflow), F = G - e1 (j - j0).  The int16 range check of
    // the alignment leaves 4 LPT max(e1, e2) of headroom for G.  The row's first cell feeds its own F (:924).
    uint32_t hp = g.shfl_up(Hf[1], 1);
    if (g.gl == 0) hp = ch0 ? (Ms[0] << 16) : carryH;
    const uint32_t Hsh0 = __funnelshift_r(hp, Hf[0], 16), Hsh1 = __funnelshift_r(Hf[0], Hf[1], 16);
    uint32_t G1[2], G2[2];
    { const uint4 c = lds128v(lk);                        // C1[0], C1[1], C2[0], C2[1]
      G1[0] = __vadd2(Hsh0, c.x); G1[1] = __vadd2(Hsh1, c.y); G2[0] = __vadd2(Hsh0, c.z); G2[1] = __vadd2(Hsh1, c.w); }
    if (!ch0 && g.gl == 0) { G1[0] = __vmaxs2(G1[0], (carryF & 0xffffu) | 0x80000000u); G2[0] = __vmaxs2(G2[0], (carryF >> 16) | 0x80000000u); }
#pragma unroll
    for (int r = 0; r < 2; ++r) { // the odd column also sees the even one
        G1[r] = __vmaxs2(G1[r], (G1[r] << 16) | 0x8000u);
        if (!AFFINE) G2[r] = __vmaxs2(G2[r], (G2[r] << 16) | 0x8000u);
    }
    G1[1] = __vmaxs2(G1[1], prmt(G1[0], 0, 0x3232));
    if (!AFFINE) G2[1] = __vmaxs2(G2[1], prmt(G2[0], 0, 0x3232));
    uint32_t TT = prmt(G1[1], G2[1], 0x7632); // lo = G1 at this lane's last column, hi = G2
#pragma unroll
    for (int dd = 1; dd < LPT; dd <<= 1) TT = __vmaxs2(TT, g.shfl_up(TT, dd)); // lanes < dd get their own value back
    uint32_t Pv = g.shfl_up(TT, 1);
    if (g.gl == 0) Pv = POA_NEGP;
    const uint32_t P1 = prmt(Pv, 0, 0x1010), P2 = prmt(Pv, 0, 0x3232);
    const uint4 nj = lds128v(lk + 16);                    // NJ1[0], NJ1[1], NJ2[0], NJ2[1]
    const uint32_t NJ1[2] = {nj.x, nj.y}, NJ2[2] = {nj.z, nj.w};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        G1[r] = __vmaxs2(G1[r], P1);
        Fa[r] = __vadd2(G1[r], NJ1[r]);
        if (AFFINE) { Fb[r] = INFP; Hn[r] = __vmaxs2(Hme[r], Fa[r]); }
        else { G2[r] = __vmaxs2(G2[r], P2); Fb[r] = __vadd2(G2[r], NJ2[r]); Hn[r] = __vimax3_s16x2(Hme[r], Fa[r], Fb[r]); }
        // Value bounds that make plain 16-bit adds exact here: every H is >= inf_min - mis (the M term), so H - oe and
        // E - e never reach -32768: the reference's saturating subtractions only ever clip candidates that lose the max.
        E1o[r] = __viaddmax_s16x2(E1x[r], P.NE1P, __vadd2(Hn[r], P.NOE1P));
        if (AFFINE) { const uint32_t keep = __vcmpeq2(Hn[r], Hme[r]); E1o[r] = (E1o[r] & keep) | (INFP & ~keep); E2o[r] = INFP; } // F won the cell: no deletion from it
        else E2o[r] = __viaddmax_s16x2(E2x[r], P.NE2P, __vadd2(Hn[r], P.NOE2P));
        midx[r] |= __vcmpeq2(Ms[r], Hn[r]) & 0x00010001u;  // the cell is a match / mismatch step from that predecessor
    }
    mcode = prmt(midx[0], midx[1], 0x6420);                // one byte per column
}

// carries from this chunk to the next one of the same row (all lanes call)
template <int LPT>
__device__ __forceinline__ void poa_chunk_carry(const PoaG<LPT> &g, const DevParams &P, const uint32_t (&Hf)[2], const uint32_t (&Fa)[2], const uint32_t (&Fb)[2],
                                                uint32_t &carryH, uint32_t &carryF) {
    carryH = g.shfl(Hf[1], LPT - 1) & 0xffff0000u;
    const uint32_t fa = g.shfl(Fa[1], LPT - 1), fb = g.shfl(Fb[1], LPT - 1);
    carryF = __vadd2(prmt(fa, fb, 0x7632), P.PE12); // G of column j0 - 1 in the next chunk's frame
}

// ---- the wide path ------------------------------------------------------------------------------------------------
// One alignment in plain 32-bit arithmetic, one task per warp (32-lane kernel only).  It serves what the packed 16-bit
// path hands over: alignments the reference itself runs with 32-bit scores (simd_abpoa_align.c:1610-1621: more than about
// 16.3 k graph rows or query bases; vectors of pn / 2 lanes, its own inf_min), alignments whose scores leave no headroom
// for the packed scan, tasks that overflowed their arena, and rows with more than 16 predecessors.  For an alignment the
// reference runs in 16 bits the band arithmetic keeps pn and the 16-bit inf_min: no value that survives a max ever wraps
// or saturates there (every H is >= inf_min - mismatch, and the reference's saturating F shifts only clip losers), so
// exact integers reproduce it bit for bit.  Simple by design: every row keeps all five planes H | E1 | E2 | F1 | F2 as
// int32 in the arena (20 bytes per cell) and the backtrack is the reference's value comparison throughout.
template <bool AFFINE>
__device__ void poa_dp_wide(PoaSmem<32> &sm, const DevParams &P, bool &ok, int &err, const int n, const int qlen, const uint16_t *q4,
                            int &n_cig, unsigned long long &cells, unsigned long long &rows) {
    const int lane = lane_id();
    PoaWs &w = sm.ws;
    constexpr int CW = 128;
    const int o1 = P.o1, e1 = P.e1, o2 = P.o2, e2 = P.e2, oe1 = P.oe1, oe2 = P.oe2, mis = P.mis_abs, mat = P.mat_abs;
    int lp = P.lp, inf_min = P.inf_min;
    { // score width as the reference picks it for THIS alignment
        const int len = qlen > n ? qlen : n;
        const long long max_score = max((long long)qlen * mat, (long long)len * e1 + o1);
        if (max_score > 32767 - mis - oe1 - (P.o2_raw + P.e2_raw)) {
            lp = P.lp - 1;
            inf_min = max(max(INT_MIN + mis, INT_MIN + oe1), INT_MIN + P.o2_raw + P.e2_raw) + 31 * max(e1, P.e2_raw);
        }
    }
    const int pn = 1 << lp, lam_bits = pn - 1;
    const int wband = 10 + (int)(0.01f * (float)qlen); // wb + (int)(wf*qlen), float (simd_abpoa_align.c:393)
    int4 *const rmeta_g = w.rmeta; const int4 *const rdesc_g = w.rdesc, *const rdesc2_g = w.rdesc2; const int32_t *const plist_g = w.plist;
    int32_t *const A = reinterpret_cast<int32_t *>(w.arena);
    const uint32_t cap = w.arena_cap / 2; // int32 units
    uint32_t used = 0;
    n_cig = 0;
    // ---- first row (simd_abpoa_align.c:538-555, 591-610)
    if (ok) {
        const int end = min(qlen, max(0, qlen - w.ri[0]) + wband);
        const int esn = end >> lp, width = (esn + 1) << lp;
        if (5ull * width > cap) { err = TH_ERR_ARENA; ok = false; }
        else {
            if (lane == 0) rmeta_g[0] = make_int4(0, 0, width - 1, 1); // the source hands 1 to its successors (:549-552)
            for (int j = lane; j < width; j += 32) {
                const int f1 = -o1 - e1 * j, f2 = -o2 - e2 * j;
                A[j] = j == 0 ? 0 : (AFFINE ? f1 : max(f1, f2));
                A[width + j] = j == 0 ? -oe1 : inf_min; A[2 * width + j] = (j == 0 && !AFFINE) ? -oe2 : inf_min;
                A[3 * width + j] = j == 0 ? inf_min : f1; A[4 * width + j] = (j == 0 || AFFINE) ? inf_min : f2;
            }
            used = 5u * width; cells += width; rows += 1;
        }
    }
    __syncwarp();
    // ---- rows in topological order
    const int qsn = qlen >> lp;
    for (int i = 1; ok && i < n - 1; ++i) {
        const int4 d = rdesc_g[i], d2 = rdesc2_g[i];
        const int np = d.y & 1023; // 1..32, checked when the descriptors were built
        int mpl = n, mpr = 0, min_pre_beg = INT_MAX;
        for (int p = 0; p < np; ++p) { // band: what the predecessors' row maxima and max_remain say (abpoa_align.h:34-35, simd_abpoa_align.c:846-854)
            const int4 m = rmeta_g[poa_pred_row(d, d2, plist_g, np, p)];
            mpl = min(mpl, m.w); mpr = max(mpr, m.w); min_pre_beg = min(min_pre_beg, m.y);
            if (lane == 0) sm.pre[p] = m;
        }
        __syncwarp();
        const int beg0 = max(0, min(mpl, d.z) - wband), end0 = min(qlen, max(mpr, d.z) + wband);
        const int beg = max((beg0 >> lp) << lp, min_pre_beg), esn = end0 >> lp, dend = ((esn + 1) << lp) - 1, bsn = beg >> lp;
        const int width = dend - beg + 1;
        if (beg > dend) { err = TH_ERR_BAND; ok = false; break; }
        if (5ull * (uint32_t)width > (unsigned long long)(cap - used)) { err = TH_ERR_ARENA; ok = false; break; }
        const uint32_t row_off = used; used += 5u * (uint32_t)width; cells += width;
        const int vb = (d.y >> 10) & 7;
        const int jmax = esn == qsn ? qlen : dend, vlast = esn - bsn;
        int32_t *R = A + row_off;
        long long best = LLONG_MIN;
        int carryH = 0, carryF1 = INT_MIN / 2, carryF2 = INT_MIN / 2; // carryF: F of the previous chunk's last column, one extension further
        const int nchunk = (width + CW - 1) / CW;
        for (int ch = 0; ch < nchunk; ++ch) {
            const int j0 = beg + ch * CW + 4 * lane;
            int Mx[4], E1x[4], E2x[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { Mx[c] = inf_min; E1x[c] = inf_min; E2x[c] = inf_min; }
            for (int p = 0; p < np; ++p) { // cells outside the predecessor's band count as inf_min (simd_abpoa_align.c:860-905)
                const int4 pm = sm.pre[p];
                const int pw = pm.z - pm.y + 1; const int32_t *Pp = A + (uint32_t)pm.x;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = j0 + c;
                    if (j - 1 >= pm.y && j - 1 <= pm.z) Mx[c] = max(Mx[c], Pp[j - 1 - pm.y]);
                    if (j >= pm.y && j <= pm.z) { E1x[c] = max(E1x[c], Pp[pw + j - pm.y]); if (!AFFINE) E2x[c] = max(E2x[c], Pp[2 * pw + j - pm.y]); }
                }
            }
            const uint32_t w4 = q4[j0 >> 2];
            int Ms[4], Hme[4], Hf[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int qc = (w4 >> (4 * c)) & 0xf;
                const int sc = (qc < 4 && vb < 4) ? (qc == vb ? mat : -mis) : 0;
                Ms[c] = Mx[c] + sc;
                Hme[c] = AFFINE ? max(Ms[c], E1x[c]) : max(Ms[c], max(E1x[c], E2x[c]));
                Hf[c] = AFFINE ? Ms[c] : Hme[c]; // what an insertion may open from
            }
            // F1[j] = max_k<=j (A1[k] - e1 (j-k)), A1[k] = Hf[k-1] - oe1, as a prefix max of G[k] = A1[k] + e1 (k - chunk start); the
            // row's first cell feeds its own F (:924)
            int hp = __shfl_up_sync(TH_FULL, Hf[3], 1);
            if (lane == 0) hp = ch == 0 ? Ms[0] : carryH;
            int G1[4], G2[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int prev = c == 0 ? hp : Hf[c - 1], k = 4 * lane + c;
                G1[c] = prev - oe1 + e1 * k; G2[c] = prev - oe2 + e2 * k;
            }
            if (ch > 0 && lane == 0) { G1[0] = max(G1[0], carryF1); G2[0] = max(G2[0], carryF2); }
#pragma unroll
            for (int c = 1; c < 4; ++c) { G1[c] = max(G1[c], G1[c - 1]); G2[c] = max(G2[c], G2[c - 1]); }
            int T1 = G1[3], T2 = G2[3];
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const int a = __shfl_up_sync(TH_FULL, T1, dd), b = __shfl_up_sync(TH_FULL, T2, dd);
                if (lane >= dd) { T1 = max(T1, a); T2 = max(T2, b); }
            }
            int P1 = __shfl_up_sync(TH_FULL, T1, 1), P2 = __shfl_up_sync(TH_FULL, T2, 1);
            if (lane == 0) { P1 = INT_MIN / 2; P2 = INT_MIN / 2; }
            int Hn[4], E1o[4], E2o[4], Fa[4], Fb[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = 4 * lane + c;
                Fa[c] = max(G1[c], P1) - e1 * k; Fb[c] = AFFINE ? inf_min : max(G2[c], P2) - e2 * k;
                Hn[c] = AFFINE ? max(Hme[c], Fa[c]) : max(Hme[c], max(Fa[c], Fb[c]));
                E1o[c] = max(E1x[c] - e1, Hn[c] - oe1);
                if (AFFINE) { if (Hn[c] != Hme[c]) E1o[c] = inf_min; E2o[c] = inf_min; } // F won the cell: no deletion from it
                else E2o[c] = max(E2x[c] - e2, Hn[c] - oe2);
            }
            if (ch + 1 < nchunk) { // (F - e) of the chunk's last column, in the next chunk's frame (k = 0)
                carryH = __shfl_sync(TH_FULL, Hf[3], 31);
                carryF1 = __shfl_sync(TH_FULL, Fa[3], 31) - e1; carryF2 = __shfl_sync(TH_FULL, Fb[3], 31) - e2;
            }
            const int rel = ((4 * lane) >> lp) + ch * (CW >> lp);
            const int sub = rel == vlast ? 0 : rel + 1;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = j0 + c;
                if (j <= dend) {
                    const int o = j - beg;
                    R[o] = Hn[c]; R[width + o] = E1o[c]; R[2 * width + o] = E2o[c]; R[3 * width + o] = Fa[c]; R[4 * width + o] = Fb[c];
                }
                if (j <= jmax) { // row arg-max key: value, then lane (j mod pn) ascending, then vector order with end_sn first
                    const long long key = ((long long)Hn[c] << 32) | (long long)(((unsigned)(lam_bits - (j & lam_bits)) << 12) | (unsigned)(0xfff - sub));
                    best = max(best, key);
                }
            }
        }
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) { const long long o = __shfl_xor_sync(TH_FULL, best, dd); best = max(best, o); }
        { // simd_abpoa_max_in_row + simd_abpoa_ada_max_i: successors pull max_i + 1 from this row's metadata
            const int val = (int)(best >> 32);
            const int lam = lam_bits - (int)((best >> 12) & 0xf), vr = 0xfff - (int)(best & 0xfff);
            const int vsn = vr == 0 ? esn : bsn + vr - 1;
            const int max_i = (best != LLONG_MIN && val > inf_min) ? vsn * pn + lam : -1;
            if (lane == 0) rmeta_g[i] = make_int4((int)row_off, beg, dend, max_i + 1);
        }
        __syncwarp();
    }
    if (ok) rows += (unsigned long long)max(n - 2, 0);
    // ---- best end cell (simd_abpoa_align.c:976-989): sink's in-neighbours in in_id order, strict >
    int bi = 0, bj = 0;
    if (ok) {
        const int4 ds = rdesc_g[n - 1], ds2 = rdesc2_g[n - 1];
        const int nps = ds.y & 1023;
        int best_score = inf_min;
        for (int p0 = 0; p0 < nps; p0 += 32) {
            const int p = p0 + lane;
            long long s = LLONG_MIN; int pi = 0, end = 0;
            if (p < nps) {
                pi = poa_pred_row(ds, ds2, plist_g, nps, p);
                const int4 m = rmeta_g[pi];
                end = qlen > m.z ? m.z : qlen;
                s = A[(uint32_t)m.x + (uint32_t)(end - m.y)];
            }
            long long mx = s;
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) mx = max(mx, __shfl_xor_sync(TH_FULL, mx, dd));
            if (mx > (long long)best_score) {
                const unsigned bm = __ballot_sync(TH_FULL, s == mx);
                const int f = __ffs(bm) - 1;
                best_score = (int)mx; bi = __shfl_sync(TH_FULL, pi, f); bj = __shfl_sync(TH_FULL, end, f);
            }
        }
    }
    // ---- backtrack by value comparison (simd_abpoa_align.c:248-377): lane p holds predecessor p
    if (ok) {
        enum { M_OP = 1, E1_OP = 2, E2_OP = 4, E_OP = 6, F1_OP = 8, F2_OP = 16, F_OP = 24, ALL_OP = 31 };
        int i = bi, j = bj, cur_op = ALL_OP;
        uint32_t *cg = w.cigar; int32_t *cq = w.cigq;
        for (int t = lane; t < qlen - bj; t += 32) { cg[t] = 1; cq[t] = qlen - 1 - t; } // unaligned query tail
        if (bj < qlen) n_cig = qlen - bj;
        while (i > 0 && j > 0) {
            const int4 d = rdesc_g[i], d2 = rdesc2_g[i], mi = rmeta_g[i];
            const int np = d.y & 1023, vb = (d.y >> 10) & 7, v = d.y >> 13;
            const int qc = (q4[j >> 2] >> (4 * (j & 3))) & 0xf;
            const int s = (qc < 4 && vb < 4) ? (qc == vb ? mat : -mis) : 0;
            const int ib = mi.y, iw = mi.z - mi.y + 1;
            const int32_t *Ri = A + (uint32_t)mi.x;
            const int hij = Ri[j - ib];
            int pi = 0; bool in1 = false, in0 = false; int4 pm = make_int4(0, 0, -1, 0);
            if (lane < np) {
                pi = poa_pred_row(d, d2, plist_g, np, lane);
                pm = rmeta_g[pi];
                in1 = j - 1 >= pm.y && j - 1 <= pm.z; in0 = j >= pm.y && j <= pm.z;
            }
            const int pw = pm.z - pm.y + 1; const int32_t *Rp = A + (uint32_t)pm.x;
            if (cur_op & M_OP) {
                const long long a = in1 ? (long long)Rp[j - 1 - pm.y] : 0; // promoted arithmetic, as the reference compares
                const unsigned mm = __ballot_sync(TH_FULL, in1 && a + s == (long long)hij);
                if (mm) {
                    const int f = __ffs(mm) - 1;
                    if (lane == 0) { cg[n_cig] = ((uint32_t)v << 2) | 0; cq[n_cig] = j - 1; }
                    ++n_cig; cur_op = ALL_OP; i = __shfl_sync(TH_FULL, pi, f); --j;
                    continue;
                }
            }
            if (cur_op & E_OP) {
                long long b = 0, x1 = 0, x2 = inf_min, e1ij = Ri[iw + j - ib], e2ij = AFFINE ? inf_min : Ri[2 * iw + j - ib];
                if (in0) { b = Rp[j - pm.y]; x1 = Rp[pw + j - pm.y]; if (!AFFINE) x2 = Rp[2 * pw + j - pm.y]; }
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
            if ((cur_op & F_OP) && j > ib) {
                const long long f1 = Ri[3 * iw + j - ib], f2 = Ri[4 * iw + j - ib], hm1 = Ri[j - 1 - ib];
                const long long f1m1 = Ri[3 * iw + j - 1 - ib], f2m1 = Ri[4 * iw + j - 1 - ib];
                bool hit = false, bad = false;
                if (cur_op & F1_OP) {
                    if (!(cur_op & M_OP) || hij == f1) {
                        if (hm1 - oe1 == f1) { cur_op = M_OP | E_OP; hit = true; }
                        else if (f1m1 - e1 == f1) { cur_op = F1_OP; hit = true; }
                        else bad = true;
                    }
                }
                if (!hit && !bad && (cur_op & F2_OP)) {
                    if (!(cur_op & M_OP) || hij == f2) {
                        if (hm1 - oe2 == f2) { cur_op = M_OP | E_OP; hit = true; }
                        else if (f2m1 - e2 == f2) { cur_op = F2_OP; hit = true; }
                        else bad = true;
                    }
                }
                if (bad) { err = TH_ERR_BACKTRACK; ok = false; break; }
                if (lane == 0) { cg[n_cig] = 1; cq[n_cig] = j - 1; } // the insertion is taken whether or not a test above fired (:357-361)
                ++n_cig; --j;
                continue;
            }
            err = TH_ERR_BACKTRACK; ok = false; break;
        }
        if (ok) {
            for (int t = lane; t < j; t += 32) { cg[n_cig + t] = 1; cq[n_cig + t] = j - 1 - t; } // unaligned query head
            if (j > 0) n_cig += j;
        }
    }
    if (!ok) n_cig = 0;
    __syncwarp();
}

// abPOA's linear gap mode (gap_open1 == 0, abpoa_align.c:86) on the wide path: simd_abpoa_lg_first_dp / lg_dp / lg_backtrack
// (simd_abpoa_align.c:557-570, 649-736, 108-158).  One matrix, H = max(M, max_pre H[pre][j] - e, H[j-1] - e), but the
// reference's vector code is not that textbook recurrence at the band edges and it is reproduced as it is:
//  * a row stores one vector more than its band (all inf_min) and successors read it;
//  * predecessor p contributes to the vectors max(beg_sn, pre_beg_sn) .. min(pre_end_sn + 1, end_sn, dp_sn - 1) only; the
//    first predecessor also writes inf_min to the row's other vectors;
//  * the horizontal pass runs vector by vector with a carried `first`, and in vectors past the predecessors' last one the
//    in-vector propagation is masked (set_num 1 or 0: SIMDShiftOneN with PRE_MASK / SUF_MIN, :613-647), which leaves
//    lanes the textbook recurrence would fill untouched.
// Values wrap to 16 bits where the reference runs the alignment with int16 vectors.  Lanes 0..pn-1 hold one vector in the
// horizontal pass; everything else is column parallel.  The backtrack is match, then deletion, else insertion.
__device__ void poa_dp_linear(PoaSmem<32> &sm, const DevParams &P, bool &ok, int &err, const int n, const int qlen, const uint16_t *q4,
                              int &n_cig, unsigned long long &cells, unsigned long long &rows) {
    const int lane = lane_id();
    PoaWs &w = sm.ws;
    const int e1 = P.e1, oe1 = P.oe1, mis = P.mis_abs, mat = P.mat_abs;
    int lp = P.lp, inf_min = P.inf_min;
    bool b16 = true;
    { // score width as the reference picks it for THIS alignment (simd_abpoa_align.c:1610-1621)
        const int len = qlen > n ? qlen : n;
        const long long max_score = max((long long)qlen * mat, (long long)len * e1 + P.o1);
        if (max_score > 32767 - mis - oe1 - (P.o2_raw + P.e2_raw)) {
            lp = P.lp - 1; b16 = false;
            inf_min = max(max(INT_MIN + mis, INT_MIN + oe1), INT_MIN + P.o2_raw + P.e2_raw) + 31 * max(e1, P.e2_raw);
        }
    }
    auto Wv = [&](int x) { return b16 ? (int)(int16_t)x : x; };
    const int pn = 1 << lp, lam_bits = pn - 1;
    const int dp_sn = (qlen + pn) >> lp;
    const int wband = 10 + (int)(0.01f * (float)qlen);
    int4 *const rmeta_g = w.rmeta; const int4 *const rdesc_g = w.rdesc, *const rdesc2_g = w.rdesc2; const int32_t *const plist_g = w.plist;
    int32_t *const A = reinterpret_cast<int32_t *>(w.arena);
    const uint32_t cap = w.arena_cap / 2;
    uint32_t used = 0;
    n_cig = 0;
    auto score = [&](int j, int vb) { const int qc = (q4[j >> 2] >> (4 * (j & 3))) & 0xf; return (qc < 4 && vb < 4) ? (qc == vb ? mat : -mis) : 0; };
    // ---- first row: H[0][j] = -e j inside the band, the extra vector inf_min
    if (ok) {
        const int end = min(qlen, max(0, qlen - w.ri[0]) + wband);
        const int esn = end >> lp, width = (esn + 1) << lp;
        if ((unsigned long long)width + pn > cap) { err = TH_ERR_ARENA; ok = false; }
        else {
            if (lane == 0) rmeta_g[0] = make_int4(0, 0, width - 1, 1);
            for (int j = lane; j < width + pn; j += 32) A[j] = j < width ? Wv(-e1 * j) : inf_min;
            used = (uint32_t)(width + pn); cells += width; rows += 1;
        }
    }
    __syncwarp();
    const int qsn = qlen >> lp;
    for (int i = 1; ok && i < n - 1; ++i) {
        const int4 d = rdesc_g[i], d2 = rdesc2_g[i];
        const int np = d.y & 1023, vb = (d.y >> 10) & 7;
        int mpl = n, mpr = 0, min_pre_beg = INT_MAX, max_pre_end = -1;
        for (int p = 0; p < np; ++p) {
            const int4 m = rmeta_g[poa_pred_row(d, d2, plist_g, np, p)];
            mpl = min(mpl, m.w); mpr = max(mpr, m.w); min_pre_beg = min(min_pre_beg, m.y); max_pre_end = max(max_pre_end, m.z);
            if (lane == 0) sm.pre[p] = m;
        }
        __syncwarp();
        const int beg0 = max(0, min(mpl, d.z) - wband), end0 = min(qlen, max(mpr, d.z) + wband);
        const int bsn = max(beg0 >> lp, min_pre_beg >> lp), esn = end0 >> lp, max_pre_end_sn = max_pre_end >> lp;
        const int beg = bsn << lp, dend = ((esn + 1) << lp) - 1, width = dend - beg + 1;
        if (bsn > esn) { err = TH_ERR_BAND; ok = false; break; }
        if ((unsigned long long)width + pn > (unsigned long long)(cap - used)) { err = TH_ERR_ARENA; ok = false; break; }
        const uint32_t row_off = used; used += (uint32_t)(width + pn); cells += width;
        int32_t *R = A + row_off;
        // ---- match and deletion candidates, one predecessor after the other (:662-711)
        for (int p = 0; p < np; ++p) {
            const int4 pm = sm.pre[p];
            const int32_t *Hp = A + (uint32_t)pm.x;
            const int pbsn = pm.y >> lp, pesn = pm.z >> lp;
            int fb_sn, first;
            if (pbsn < bsn) { fb_sn = bsn; first = (bsn - 1 <= pesn + 1) ? Hp[(bsn << lp) - 1 - pm.y] : inf_min; }
            else { fb_sn = pbsn; first = inf_min; }
            const int fe_sn = min(min(pesn + 1, esn), dp_sn - 1);
            for (int j = beg + lane; j <= dend + pn; j += 32) {
                const int sn = j >> lp;
                if (sn >= fb_sn && sn <= fe_sn) {
                    const int left = j == (fb_sn << lp) ? first : Hp[j - 1 - pm.y];
                    const int v = max(Wv(left + score(j, vb)), Wv(Hp[j - pm.y] - e1));
                    if (p == 0 || v > R[j - beg]) R[j - beg] = v;
                } else if (p == 0) R[j - beg] = inf_min;
            }
            __syncwarp();
        }
        // ---- horizontal pass, vector by vector (:713-733)
        {
            int first = R[0];
            for (int sn = bsn; sn <= esn; ++sn) {
                const int set_num = sn > max_pre_end_sn ? (sn == max_pre_end_sn + 1 ? 1 : 0) : pn;
                int h = lane < pn ? R[((sn - bsn) << lp) + lane] : inf_min;
                if (lane == 0) h = max(h, first);
                int cov = set_num;
                for (int k = 0; (1 << k) < pn; ++k) {
                    const int sh = 1 << k;
                    int t = __shfl_up_sync(TH_FULL, h, sh);
                    bool take = lane >= sh && lane < pn;
                    if (set_num != pn) { if (k > 0) cov += sh; take = take && lane <= min(cov, pn); }
                    t = take ? Wv(t - e1 * sh) : inf_min;
                    h = max(h, t);
                }
                if (lane < pn) R[((sn - bsn) << lp) + lane] = h;
                first = Wv(__shfl_sync(TH_FULL, h, pn - 1) - e1);
            }
        }
        __syncwarp();
        // ---- row arg-max (:991-1005) as on the convex path
        const int jmax = esn == qsn ? qlen : dend, vlast = esn - bsn;
        long long best = LLONG_MIN;
        for (int j = beg + lane; j <= jmax; j += 32) {
            const int rel = (j - beg) >> lp, sub = rel == vlast ? 0 : rel + 1;
            const long long key = ((long long)R[j - beg] << 32) | (long long)(((unsigned)(lam_bits - (j & lam_bits)) << 12) | (unsigned)(0xfff - sub));
            best = max(best, key);
        }
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) { const long long o = __shfl_xor_sync(TH_FULL, best, dd); best = max(best, o); }
        {
            const int val = (int)(best >> 32);
            const int lam = lam_bits - (int)((best >> 12) & 0xf), vr = 0xfff - (int)(best & 0xfff);
            const int vsn = vr == 0 ? esn : bsn + vr - 1;
            const int max_i = (best != LLONG_MIN && val > inf_min) ? vsn * pn + lam : -1;
            if (lane == 0) rmeta_g[i] = make_int4((int)row_off, beg, dend, max_i + 1);
        }
        __syncwarp();
    }
    if (ok) rows += (unsigned long long)max(n - 2, 0);
    // ---- best end cell (:976-989)
    int bi = 0, bj = 0;
    if (ok) {
        const int4 ds = rdesc_g[n - 1], ds2 = rdesc2_g[n - 1];
        const int nps = ds.y & 1023;
        int best_score = inf_min;
        for (int p = 0; p < nps; ++p) { // every lane walks the list: the sink has few in-neighbours
            const int pi = poa_pred_row(ds, ds2, plist_g, nps, p);
            const int4 m = rmeta_g[pi];
            const int end = qlen > m.z ? m.z : qlen;
            const int s = A[(uint32_t)m.x + (uint32_t)(end - m.y)];
            if (s > best_score) { best_score = s; bi = pi; bj = end; }
        }
    }
    // ---- backtrack (:108-158): lane p holds predecessor p
    if (ok) {
        int i = bi, j = bj;
        uint32_t *cg = w.cigar; int32_t *cq = w.cigq;
        for (int t = lane; t < qlen - bj; t += 32) { cg[t] = 1; cq[t] = qlen - 1 - t; }
        if (bj < qlen) n_cig = qlen - bj;
        while (i > 0 && j > 0) {
            const int4 d = rdesc_g[i], d2 = rdesc2_g[i], mi = rmeta_g[i];
            const int np = d.y & 1023, vb = (d.y >> 10) & 7, v = d.y >> 13;
            const long long hij = A[(uint32_t)mi.x + (uint32_t)(j - mi.y)];
            const int s = score(j, vb);
            int pi = 0; long long hm = 0, hd = 0; bool in1 = false, in0 = false;
            if (lane < np) {
                pi = poa_pred_row(d, d2, plist_g, np, lane);
                const int4 pm = rmeta_g[pi];
                in1 = j - 1 >= pm.y && j - 1 <= pm.z; in0 = j >= pm.y && j <= pm.z;
                if (in1) hm = A[(uint32_t)pm.x + (uint32_t)(j - 1 - pm.y)];
                if (in0) hd = A[(uint32_t)pm.x + (uint32_t)(j - pm.y)];
            }
            const unsigned mm = __ballot_sync(TH_FULL, in1 && hm + s == hij);
            if (mm) {
                if (lane == 0) { cg[n_cig] = ((uint32_t)v << 2) | 0; cq[n_cig] = j - 1; }
                ++n_cig; i = __shfl_sync(TH_FULL, pi, __ffs(mm) - 1); --j;
                continue;
            }
            const unsigned dm = __ballot_sync(TH_FULL, in0 && hd - e1 == hij);
            if (dm) {
                if (lane == 0) { cg[n_cig] = ((uint32_t)v << 2) | 2; cq[n_cig] = j - 1; }
                ++n_cig; i = __shfl_sync(TH_FULL, pi, __ffs(dm) - 1);
                continue;
            }
            if (lane == 0) { cg[n_cig] = 1; cq[n_cig] = j - 1; }
            ++n_cig; --j;
        }
        for (int t = lane; t < j; t += 32) { cg[n_cig + t] = 1; cq[n_cig + t] = j - 1 - t; }
        if (j > 0) n_cig += j;
    }
    if (!ok) n_cig = 0;
    __syncwarp();
}

// Aligns sequence `query` of every active group to its graph and merges it in.  ALL lanes of the warp call; `act` says
// whether this lane's group takes part.  err is the group's status (TH_OK going in).
template <int LPT, bool AFFINE>
__device__ void poa_add_sequence(PoaSmem<LPT> &sm, const DevParams &P, const uint32_t lk, const bool act, const uint8_t *query, const int qlen_in,
                                 int &node_n, int &edge_n, int &err, unsigned long long &cells, unsigned long long &rows, long long *ph) {
#ifdef POA_PROFILE   // per-phase warp-cycle counters (tools/profile_step.py); off in the product build
    long long t_ph = clock64();
#define PH(k) do { long long t_ = clock64(); ph[k] += t_ - t_ph; t_ph = t_; } while (0)
#else
#define PH(k) do { } while (0)
#endif
    constexpr int CW = PoaSmem<LPT>::CW, RINGW = PoaSmem<LPT>::RINGW, RINGC = PoaSmem<LPT>::RINGC;
    const PoaG<LPT> g;
    const int gl = g.gl;
    PoaWs &w = sm.ws;
    bool ok = act;                                     // this group is (still) working on the alignment
    const int n = act ? node_n : 0, qlen = act ? qlen_in : 0;
    const int lp = P.lp, pn = P.pn;                   // log2 / width of the emulated vector (16: AVX2 int16, the reference build; 8: SSE)
    const int o1 = P.o1, e1 = P.e1, o2 = P.o2, e2 = P.e2, oe1 = P.oe1, oe2 = P.oe2;
    const int mis = P.mis_abs, mat = P.mat_abs;
    if (LPT != 32 && ok) { // packed 16-bit scores only where the reference uses them (simd_abpoa_align.c:1610-1621) and with CW max(e) of
        // headroom for the scan frame; the rest is the wide path's (second pass)
        const int len = qlen > n ? qlen : n;
        const int max_score = max(qlen * mat, len * e1 + o1);
        if (P.linear || max_score > 32767 - mis - oe1 - max(oe2, P.o2_raw + P.e2_raw) - CW * max(max(e1, e2), P.e2_raw)) { err = TH_ERR_LEN; ok = false; }
    }
    const int inf_min = P.inf_min;
    const int wband = 10 + (int)(0.01f * (float)qlen); // wb + (int)(wf*qlen), float (simd_abpoa_align.c:393)
    int4 *const rmeta_g = w.rmeta; int4 *const rdesc_g = w.rdesc; int32_t *const plist_g = w.plist;
    // ---- row descriptors: order index, predecessors by row, heaviest successor ----------------
    { const int nn = ok ? n : 0;
      for (int i = gl; i < nn; i += LPT) w.n2i[w.ord[i]] = i; }
    if (gl == 0) sm.plist_n = 0;
    __syncwarp();
    { // the node records hold each node's first three in-neighbours and its heaviest out-edge: a descriptor costs three
      // dependent loads (node, records, their indices) and no list walk; four nodes per lane are in flight
        constexpr int U = 4;
        int bad = 0;
        const int nn = ok ? n : 0;
        for (int i0 = 0; i0 < nn; i0 += LPT * U) {
            int v[U], b[U]; int4 nr[U], br[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const int i = i0 + LPT * u + gl; v[u] = i < nn ? w.ord[i] : -1; }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                nr[u] = make_int4(0, -1, -1, -1); br[u] = make_int4(-1, 0, 0, 0); b[u] = 0;
                if (v[u] >= 0) { nr[u] = w.nrec[v[u]]; br[u] = w.brec[v[u]]; b[u] = w.base[v[u]]; }
            }
            int p0[U], p1[U], p2[U], hi[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                p0[u] = nr[u].x > 0 ? w.n2i[nr[u].y] : -1; p1[u] = nr[u].x > 1 ? w.n2i[nr[u].z] : -1; p2[u] = nr[u].x > 2 ? w.n2i[nr[u].w] : -1;
                hi[u] = br[u].x >= 0 ? w.n2i[br[u].x] : 0x7fffffff;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + LPT * u + gl;
                if (i < nn) {
                    const int np = nr[u].x;
                    int4 d2 = make_int4(p2[u], -1, -1, -1);
                    if (np > 3) { // rare: the record holds three in-neighbours, the rest come from the list
                        int off = 0, k = 0;
                        if (np > 6) { off = atomicAdd(&sm.plist_n, np); d2.x = off; } // the lists' order among rows does not matter
                        for (int e = w.in_head[v[u]]; e >= 0; e = w.e_ni[e], ++k) {
                            const int pi = w.n2i[w.e_from[e]];
                            if (np > 6) plist_g[off + k] = pi; else if (k == 3) d2.y = pi; else if (k == 4) d2.z = pi; else if (k == 5) d2.w = pi;
                        }
                    }
                    w.rdesc2[i] = d2;
                    w.hi_idx[i] = hi[u];
                    bad |= np > 1023 || (i > 0 && i < nn - 1 && (unsigned)(np - 1) >= (unsigned)min(LPT, POA_MAXPRE)); // the row loop relies on 1 <= np <= LPT
                    rdesc_g[i] = make_int4(p0[u], min(np, 1023) | (min(b[u], 7) << 10) | (v[u] << 13), 0, p1[u]);
                }
            }
        }
        if (g.any(bad)) { if (ok) err = TH_ERR_CAP; ok = false; }
    }
    __syncwarp();
    // max_remain by index, LPT indices at a time from the sink backwards; chains inside a chunk are
    // collapsed by pointer jumping on shuffles (remain[v] = remain[heaviest successor] + 1)
    int32_t *ri = w.ri;
    {
        int cb = ok ? ((n - 1) / LPT) * LPT : -1;
        while (__any_sync(TH_FULL, cb >= 0)) {
            const bool a = cb >= 0;
            const int idx = cb + gl;
            int ptr = 0x7fffffff, dist = 0;
            if (a && idx < n) { ptr = w.hi_idx[idx]; dist = 1; if (idx == n - 1) { ptr = 0x7fffffff; dist = -1; } }
#pragma unroll
            for (int rnd = 0; (1 << rnd) < LPT; ++rnd) {
                const bool inside = a && ptr < cb + LPT;
                const int tl = inside ? ptr - cb : 0;
                const int pd = g.shfl(dist, tl), pp = g.shfl(ptr, tl);
                if (inside) { dist += pd; ptr = pp; }
            }
            if (a && idx < n) {
                const int val = (ptr == 0x7fffffff ? 0 : ri[ptr]) + dist;
                ri[idx] = val;
                reinterpret_cast<int32_t *>(rdesc_g + idx)[2] = qlen - val;
            }
            __syncwarp();
            if (a) cb -= LPT;
        }
    }
    // ---- query codes (the reference's 5 x qlen profile, simd_abpoa_align.c:438-446, reduced to what it encodes): 4 bits per
    // column, code of column j = query[j-1] for an A/C/G/T base, 8 (score 0 against every node) for column 0, N and
    // columns past the end
    const int q4n = (qlen + pn + CW + 8) >> 2;          // entries: every column a lane of any chunk may look at
    uint16_t *const q4 = q4n <= PoaSmem<LPT>::Q4N ? sm.q4 : reinterpret_cast<uint16_t *>(w.q8g);
    if (ok) for (int wi = gl; wi < q4n; wi += LPT) {
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = 4 * wi + b;
            uint32_t qc = 8;
            if (j >= 1 && j <= qlen) { qc = query[j - 1]; if (qc > 3) qc = 8; }
            word |= qc << (4 * b);
        }
        q4[wi] = (uint16_t)word;
    }
    int n_cig = 0;
    if constexpr (LPT == 32) { // the 32-lane kernel is the wide path: 32-bit arithmetic, see poa_dp_wide
        if (P.linear) poa_dp_linear(sm, P, ok, err, n, qlen, q4, n_cig, cells, rows);
        else poa_dp_wide<AFFINE>(sm, P, ok, err, n, qlen, q4, n_cig, cells, rows);
        PH(1);
    } else {
    // ---- first row (simd_abpoa_align.c:538-555, 591-610) ------------------------------------
    uint32_t used = 0;
    uint32_t *const A32w = reinterpret_cast<uint32_t *>(w.arena);
    const uint32_t *const A32 = A32w;
    const uint32_t arena_cap = w.arena_cap;
    uint32_t *const ring = sm.u.f.ring;
    if (ok) {
        const int end = min(qlen, max(0, qlen - ri[0]) + wband);
        const int esn = end >> lp, width = (esn + 1) << lp;
        if ((unsigned long long)poa_row_size(width) + 64 > arena_cap) { err = TH_ERR_ARENA; ok = false; }
        else {
            if (gl == 0) { const int4 m = make_int4(0, 0, width - 1, 1); rmeta_g[0] = m; sm.u.f.meta[0] = m; } // the source hands 1 to its successors (:549-552)
            const int ps = width >> 1;
            for (int q = gl; q < ps; q += LPT) {
                uint32_t hh = 0, ee1 = 0, ee2 = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = 2 * q + h;
                    const int f1 = -o1 - e1 * j, f2 = -o2 - e2 * j;
                    const int vh = j == 0 ? 0 : (AFFINE ? (int)(int16_t)f1 : max((int)(int16_t)f1, (int)(int16_t)f2));
                    const int v1 = j == 0 ? -oe1 : inf_min, v2 = (j == 0 && !AFFINE) ? -oe2 : inf_min;
                    hh |= (uint32_t)(uint16_t)vh << (16 * h); ee1 |= (uint32_t)(uint16_t)v1 << (16 * h); ee2 |= (uint32_t)(uint16_t)v2 << (16 * h);
                }
                A32w[q] = hh; A32w[ps + q] = ee1; if (!AFFINE) A32w[2 * ps + q] = ee2;
                if (width <= RINGC) { ring[q] = hh; ring[ps + q] = ee1; if (!AFFINE) ring[2 * ps + q] = ee2; }
            }
            used = poa_row_size(width); cells += width; rows += 1;
        }
    }
    __syncwarp();
    PH(0);
    // ---- rows in topological order ----------------------------------------------------------
    const int lam_bits = pn - 1;
    const int lane_vec = (4 * gl) >> lp;
    const int qsn = qlen >> lp;
    uint32_t ncell = 0;
    {
        int i = 1;
        bool ract = ok && n > 2;
        int4 dn = ract ? rdesc_g[1] : make_int4(0, 0, 0, 0), dn2 = ract ? w.rdesc2[1] : make_int4(0, 0, 0, 0);
        const int4 *const rdesc2_g = w.rdesc2;
        while (__any_sync(TH_FULL, ract)) {
            // ---- band and predecessors of row i (no warp-synchronous operation before the chunk)
            const int4 d = dn, d2 = dn2;
            int4 pm0 = make_int4(0, 0, -1, 0);
            int np = 0, beg = 0, dend = -1, esn = 0, width = 0;
            uint32_t row_off = 0;
            bool go = false;
            if (ract) {
                if (i + 1 < n - 1) { dn = rdesc_g[i + 1]; dn2 = rdesc2_g[i + 1]; } // next row's descriptors: in flight while this row is computed
                np = d.y & 1023;                           // 1..LPT, checked when the descriptors were built
                // band: what the predecessors' row maxima and max_remain say (abpoa_align.h:34-35, simd_abpoa_align.c:846-854)
                int4 m = (i - d.x < POA_WIN) ? sm.u.f.meta[d.x & (POA_WIN - 1)] : rmeta_g[d.x];
                int mpl = min(n, m.w), mpr = max(0, m.w), min_pre_beg = m.y;
                if (i - d.x < POA_NRING && m.z - m.y < RINGC) m.x = (int)(0x80000000u | (uint32_t)(d.x & (POA_NRING - 1)));
                pm0 = m;
                if (np > 1) {
#pragma unroll 1
                    for (int p = 1; p < np; ++p) {
                        const int pi = poa_pred_row(d, d2, plist_g, np, p);
                        int4 q = (i - pi < POA_WIN) ? sm.u.f.meta[pi & (POA_WIN - 1)] : rmeta_g[pi];
                        mpl = min(mpl, q.w); mpr = max(mpr, q.w); min_pre_beg = min(min_pre_beg, q.y);
                        if (i - pi < POA_NRING && q.z - q.y < RINGC) q.x = (int)(0x80000000u | (uint32_t)(pi & (POA_NRING - 1)));
                        sm.pre[p] = q;                     // every lane of the group writes the same value and reads back its own write
                    }
                }
                const int beg0 = max(0, min(mpl, d.z) - wband), end0 = min(qlen, max(mpr, d.z) + wband);
                beg = max((beg0 >> lp) << lp, min_pre_beg); esn = end0 >> lp; dend = ((esn + 1) << lp) - 1;
                width = dend - beg + 1;
                if (beg > dend) err = TH_ERR_BAND;
                else if ((unsigned long long)poa_row_size(width) + 64 > (unsigned long long)(arena_cap - used)) err = TH_ERR_ARENA;
                else { go = true; row_off = used; used += poa_row_size(width); ncell += (uint32_t)width; }
                if (!go) { ract = false; ok = false; width = 0; }
            }
            const int ps = width >> 1, nchunk = (width + CW - 1) / CW;
            const bool cache_row = width <= RINGC;
            const int vb = (d.y >> 10) & 7;
            const uint32_t basew = (uint32_t)min(vb, 4) * 0x01010101u;
            const int jmax = esn == qsn ? qlen : dend;   // columns past the query end do not compete for the row maximum
            const int vlast = esn - (beg >> lp);          // the row's last vector is visited first by the reference's arg-max
            int best = INT_MIN;
            uint32_t carryH = 0, carryF = POA_NEGP;
            const bool multi = __any_sync(TH_FULL, go && nchunk > 1);
            auto chunk = [&](const int ch, const bool more) {
                const bool a = go && ch < nchunk;
                const int j0 = beg + ch * CW + 4 * gl;
                uint32_t Hn[2], E1o[2], E2o[2], Fa[2], Fb[2], Hf[2], mcode;
                poa_chunk<LPT, AFFINE>(g, P, lk, A32, ring, sm.pre, pm0, a ? np : 0, a, j0, ch == 0, carryH, carryF, basew, q4, Hn, E1o, E2o, Fa, Fb, Hf, mcode);
                if (a && j0 <= dend) {
                    const uint32_t wo = (row_off >> 1) + (uint32_t)((j0 - beg) >> 1);
                    A32w[(row_off >> 1) + 3u * (uint32_t)ps + (uint32_t)((j0 - beg) >> 2)] = mcode;
                    *reinterpret_cast<uint2 *>(A32w + wo) = make_uint2(Hn[0], Hn[1]);
                    *reinterpret_cast<uint2 *>(A32w + wo + ps) = make_uint2(E1o[0], E1o[1]);
                    if (!AFFINE) *reinterpret_cast<uint2 *>(A32w + wo + 2 * ps) = make_uint2(E2o[0], E2o[1]);
                    if (cache_row) { // the slot held row i - POA_NRING, which no row reads from the ring any more (a two-chunk row would
                                     // otherwise overwrite it between its own chunks)
                        uint32_t *rs = ring + (i & (POA_NRING - 1)) * RINGW + ((j0 - beg) >> 1);
                        *reinterpret_cast<uint2 *>(rs) = make_uint2(Hn[0], Hn[1]);
                        *reinterpret_cast<uint2 *>(rs + ps) = make_uint2(E1o[0], E1o[1]);
                        if (!AFFINE) *reinterpret_cast<uint2 *>(rs + 2 * ps) = make_uint2(E2o[0], E2o[1]);
                    }
                }
                if (a) { // row arg-max key (signed compare): value, then lane (j mod pn) ascending, then vector order with end_sn first
                    const int rel = lane_vec + ch * (CW >> lp);
                    const uint32_t sub = rel == vlast ? 0u : (uint32_t)(rel + 1);
                    const uint4 kk = lds128v(lk + 32);    // KLO[0], KLO[1], KHI[0], KHI[1]
                    const uint32_t KLO[2] = {kk.x, kk.y}, KHI[2] = {kk.z, kk.w};
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int j = j0 + 2 * r;
                        const int klo = (int)prmt(Hn[r], KLO[r] - sub, 0x1054), khi = (int)prmt(Hn[r], KHI[r] - sub, 0x3254);
                        best = max(best, max(j <= jmax ? klo : INT_MIN, j < jmax ? khi : INT_MIN)); // jmax <= dend
                    }
                }
                if (more) poa_chunk_carry<LPT>(g, P, Hf, Fa, Fb, carryH, carryF);
            };
            { int ch = 0; // rows wider than one chunk: the other group keeps the warp company
              do { chunk(ch, multi); ++ch; } while (multi && __any_sync(TH_FULL, go && ch < nchunk)); }
            const int rbest = g.rmax(best);
            if (go) { // simd_abpoa_max_in_row + simd_abpoa_ada_max_i: successors pull max_i + 1 from this row's metadata
                const int val = rbest >> 16;
                const int lam = lam_bits - (int)((rbest >> 12) & 0xf), vr = 0xfff - (int)(rbest & 0xfff);
                const int vsn = vr == 0 ? esn : (beg >> lp) + vr - 1;
                const int max_i = (rbest != INT_MIN && val > inf_min) ? vsn * pn + lam : -1;
                if (gl == 0) { const int4 m = make_int4((int)row_off, beg, dend, max_i + 1); sm.u.f.meta[i & (POA_WIN - 1)] = m; rmeta_g[i] = m; }
                ++i; ract = i < n - 1;
            }
            __syncwarp();
        }
    }
    if (ok) { cells += ncell; rows += (unsigned long long)max(n - 2, 0); }
    PH(1);
    // ---- best end cell (simd_abpoa_align.c:976-989): sink's in-neighbours in in_id order, strict > ----
    const int16_t *const A16 = w.arena;
    int bi = 0, bj = 0;
    {
        const int4 ds = ok ? rdesc_g[n - 1] : make_int4(0, 0, 0, 0), ds2 = ok ? w.rdesc2[n - 1] : make_int4(0, 0, 0, 0);
        const int nps = ok ? (ds.y & 1023) : 0;
        int best_score = inf_min;
        for (int p0 = 0; __any_sync(TH_FULL, p0 < nps); p0 += LPT) {
            const int p = p0 + gl;
            int s = -0x7fffffff, pi = 0, end = 0;
            if (p < nps) {
                pi = poa_pred_row(ds, ds2, plist_g, nps, p);
                const int4 m = rmeta_g[pi];
                end = qlen > m.z ? m.z : qlen;
                s = A16[(uint32_t)m.x + (uint32_t)(end - m.y)];
            }
            const int mx = g.rmax(s);
            const unsigned bm = g.ballot(s == mx);
            const int f = bm ? __ffs(bm) - 1 : 0;
            const int fpi = g.shfl(pi, f), fend = g.shfl(end, f);
            if (p0 < nps && mx > best_score) { best_score = mx; bi = fpi; bj = fend; }
        }
    }
    // ---- backtrack (simd_abpoa_align.c:248-377).  The walk is sequential.  A window of 64 rows is staged in shared memory
    // 32 rows at a time: each row's predecessors and the M-codes of 32 columns around its arg-max.  Where the code says
    // "match / mismatch step from predecessor k" the step is taken from it (the code is exactly the outcome of the
    // reference's first value test); every other step -- deletions, insertions, columns outside the staged ones -- is the
    // reference's value comparison, lane p of the group holding predecessor p and reading the arena.
    {
        enum { M_OP = 1, E1_OP = 2, E2_OP = 4, E_OP = 6, F1_OP = 8, F2_OP = 16, F_OP = 24, ALL_OP = 31 };
        int i = ok ? bi : 0, j = ok ? bj : 0, cur_op = ALL_OP;
        uint32_t *cg = w.cigar; int32_t *cq = w.cigq;
        if (ok) {
            for (int t = gl; t < qlen - bj; t += LPT) { cg[t] = 1; cq[t] = qlen - 1 - t; } // unaligned query tail
            if (bj < qlen) n_cig = qlen - bj;
        }
        int wlo = n, whi = -1; // rows [wlo, whi] are in the window
        const uint8_t *const A8 = reinterpret_cast<const uint8_t *>(w.arena);
        bool bact;
        while (__any_sync(TH_FULL, bact = (ok && i > 0 && j > 0))) {
            const bool need = bact && (i < wlo || (i < wlo + 32 && wlo > 0));
            if (__any_sync(TH_FULL, need)) {
                // The walk is entering the 32 rows staged last time: their codes, copied asynchronously while the 32 rows before
                // them were walked, have to be there now.  Then the next 32 rows are put in flight.
                cp_async_wait_all();
                __syncwarp();
                bool jumped = false;
                if (need) {
                    const int top = i < wlo ? i + 1 : wlo; // stage rows [top-32, top)
                    jumped = i < wlo;
                    if (jumped) whi = i;
                    for (int r = top - 32 + gl; r < top; r += LPT) if (r >= 0) {
                        const int4 m = rmeta_g[r], d = rdesc_g[r];
                        // the path crosses a row close to that row's maximum (the column that steered the band): stage 32 columns
                        // around it, from a multiple of 16 columns into the row
                        const int jp = m.w > 0 ? m.w - 1 : j - (i - r);
                        const int wd = m.z - m.y + 1;
                        int k16 = max(0, jp - m.y - 8) & ~15; k16 = min(k16, max(0, wd - 32) & ~15);
                        const uint4 *src = reinterpret_cast<const uint4 *>(A8 + 2ull * ((uint32_t)m.x + 3u * (uint32_t)wd) + (uint32_t)k16);
                        PoaBt &R = sm.u.bt[r & (POA_WIN - 1)];
                        cp_async16(&R.code[0], src); cp_async16(&R.code[1], src + 1); // may run past a narrow row's codes: the arena keeps 128 bytes of slack
                        R.hdr = make_int4(d.x, d.w, d.y, m.y + k16);
                    }
                    wlo = max(0, top - 32); whi = min(whi, wlo + POA_WIN - 1);
                }
                cp_async_commit();
                if (__any_sync(TH_FULL, jumped)) cp_async_wait_all(); // a window started from scratch is walked at once
                __syncwarp();
            }
            const PoaBt &R = sm.u.bt[i & (POA_WIN - 1)];
            const int4 h = R.hdr;
            const int np = h.z & 1023, vb = (h.z >> 10) & 7, v = h.z >> 13;
            const int cc = j - h.w;
            const bool coded = bact && (cur_op & M_OP) && (unsigned)cc < 32u;
            const int code = coded ? (int)reinterpret_cast<const uint8_t *>(R.code)[cc] : 0;
            bool stepped = false;
            if (coded && (code & 1)) { // a match / mismatch step from predecessor code >> 1
                const int k = code >> 1;
                int pi = k == 0 ? h.x : h.y;
                if (k > 1) pi = poa_pred_row(make_int4(h.x, 0, 0, h.y), w.rdesc2[i], plist_g, np, k);
                if (gl == 0) { cg[n_cig] = ((uint32_t)v << 2) | 0; cq[n_cig] = j - 1; }
                ++n_cig; cur_op = ALL_OP; i = pi; --j; stepped = true;
            }
            if (__any_sync(TH_FULL, bact && !stepped)) { // the reference's step by value comparison (for some group)
                const bool todo = bact && !stepped;
                const int qb = todo ? (q4[j >> 2] >> (4 * (j & 3))) & 0xf : 8;
                const int s = (qb < 4 && vb < 4) ? (qb == vb ? mat : -mis) : 0;
                int4 mi = make_int4(0, 0, -1, 0);
                if (todo) mi = rmeta_g[i];
                const int ib = mi.y, iw = mi.z - mi.y + 1;
                const int hij = todo ? (int)A16[(uint32_t)mi.x + (uint32_t)(j - ib)] : 0;
                // lane p holds predecessor p
                int pi = 0; bool in1 = false, in0 = false; int4 pm = make_int4(0, 0, -1, 0);
                if (todo && gl < np) {
                    pi = gl < 2 ? (gl == 0 ? h.x : h.y) : poa_pred_row(make_int4(h.x, 0, 0, h.y), w.rdesc2[i], plist_g, np, gl);
                    pm = rmeta_g[pi];
                    in1 = j - 1 >= pm.y && j - 1 <= pm.z; in0 = j >= pm.y && j <= pm.z;
                }
                if (__any_sync(TH_FULL, todo && (cur_op & M_OP) && !coded)) { // match test where no code was staged
                    const bool tm = todo && (cur_op & M_OP) && !coded;
                    int a = 0;
                    if (tm && in1) a = A16[(uint32_t)pm.x + (uint32_t)(j - 1 - pm.y)];
                    const unsigned mm = g.ballot(tm && in1 && a + s == hij);
                    const int f = mm ? __ffs(mm) - 1 : 0;
                    const int fpi = g.shfl(pi, f);
                    if (mm) {
                        if (gl == 0) { cg[n_cig] = ((uint32_t)v << 2) | 0; cq[n_cig] = j - 1; }
                        ++n_cig; cur_op = ALL_OP; i = fpi; --j; stepped = true;
                    }
                }
                if (__any_sync(TH_FULL, todo && !stepped && (cur_op & E_OP))) { // deletions: predecessors at column j
                    const bool te = todo && !stepped && (cur_op & E_OP);
                    int b = 0, x1 = 0, x2 = inf_min, e1ij = 0, e2ij = inf_min;
                    if (te && in0) {
                        const uint32_t pw = (uint32_t)(pm.z - pm.y + 1), po = (uint32_t)pm.x + (uint32_t)(j - pm.y);
                        b = A16[po]; x1 = A16[po + pw]; if (!AFFINE) x2 = A16[po + 2 * pw];
                    }
                    if (te && !(cur_op & M_OP)) { const uint32_t io = (uint32_t)mi.x + (uint32_t)(j - ib); e1ij = A16[io + iw]; if (!AFFINE) e2ij = A16[io + 2 * iw]; }
                    const bool ok1 = te && (cur_op & E1_OP) && in0 && ((cur_op & M_OP) ? (hij == x1) : (e1ij == x1 - e1));
                    const bool ok2 = te && (cur_op & E2_OP) && in0 && ((cur_op & M_OP) ? (hij == x2) : (e2ij == x2 - e2));
                    const unsigned em = g.ballot(ok1 || ok2);
                    const int f = em ? __ffs(em) - 1 : 0;
                    int nop;
                    if (ok1) nop = (b - oe1 == x1) ? (M_OP | F_OP) : E1_OP;
                    else nop = (b - oe2 == x2) ? (M_OP | F_OP) : E2_OP;
                    const int fop = g.shfl(nop, f), fpi = g.shfl(pi, f);
                    if (em) {
                        cur_op = fop;
                        if (gl == 0) { cg[n_cig] = ((uint32_t)v << 2) | 2; cq[n_cig] = j - 1; }
                        ++n_cig; i = fpi; stepped = true;
                    }
                }
                if (__any_sync(TH_FULL, todo && !stepped)) { // insertions: F1 / F2 of this row at columns j and j - 1, recomputed
                    const bool tf = todo && !stepped && (cur_op & F_OP) && j > ib;
                    if (todo && !stepped && !tf) { err = TH_ERR_BACKTRACK; ok = false; }
                    // predecessors' planes from the arena (the ring belongs to the forward pass)
                    int4 p0m = make_int4(0, 0, -1, 0);
                    if (tf) {
                        p0m = rmeta_g[h.x];
                        if (gl >= 1 && gl < np) sm.pre[gl] = rmeta_g[pi]; // lane p holds predecessor p (np <= LPT)
                    }
                    __syncwarp();
                    const uint32_t basew = (uint32_t)min(vb, 4) * 0x01010101u;
                    const int cj = tf ? (j - ib) / CW : -1;
                    int f1 = 0, f2 = 0, f1m1 = 0, f2m1 = 0;
                    uint32_t carryH = 0, carryF = POA_NEGP;
                    for (int ch = 0; __any_sync(TH_FULL, ch <= cj); ++ch) {
                        const bool a = ch <= cj;
                        const int j0 = ib + ch * CW + 4 * gl;
                        uint32_t Hn[2], E1o[2], E2o[2], Fa[2], Fb[2], Hf[2], mcode;
                        poa_chunk<LPT, AFFINE>(g, P, lk, A32, ring, sm.pre, p0m, a ? np : 0, a, j0, ch == 0, carryH, carryF, basew, q4, Hn, E1o, E2o, Fa, Fb, Hf, mcode);
#pragma unroll
                        for (int x = 0; x < 2; ++x) { // x = 0: column j, x = 1: column j - 1
                            const int rel = j - x - ib - ch * CW;
                            const bool inch = a && rel >= 0 && rel < CW;
                            const int own = (rel >> 2) & (LPT - 1), r = (rel >> 1) & 1, hf = rel & 1;
                            const int va = s16_at(r ? Fa[1] : Fa[0], hf), vbb = s16_at(r ? Fb[1] : Fb[0], hf);
                            const int ta = g.shfl(va, own), tb = g.shfl(vbb, own);
                            if (inch) { if (x == 0) { f1 = ta; f2 = tb; } else { f1m1 = ta; f2m1 = tb; } }
                        }
                        poa_chunk_carry<LPT>(g, P, Hf, Fa, Fb, carryH, carryF);
                    }
                    __syncwarp();                       // the staged predecessor metadata has been read: a later step may overwrite it
                    if (tf) {
                        const int hm1 = A16[(uint32_t)mi.x + (uint32_t)(j - 1 - ib)];
                        bool hit = false, bad = false;
                        if (cur_op & F1_OP) {
                            if (!(cur_op & M_OP) || hij == f1) {
                                if (hm1 - oe1 == f1) { cur_op = M_OP | E_OP; hit = true; }
                                else if (f1m1 - e1 == f1) { cur_op = F1_OP; hit = true; }
                                else bad = true;
                            }
                        }
                        if (!hit && !bad && (cur_op & F2_OP)) {
                            if (!(cur_op & M_OP) || hij == f2) {
                                if (hm1 - oe2 == f2) { cur_op = M_OP | E_OP; hit = true; }
                                else if (f2m1 - e2 == f2) { cur_op = F2_OP; hit = true; }
                                else bad = true;
                            }
                        }
                        if (bad) { err = TH_ERR_BACKTRACK; ok = false; }
                        else { // simd_abpoa_align.c:357-361: the insertion is taken whether or not a test above fired
                            if (gl == 0) { cg[n_cig] = 1; cq[n_cig] = j - 1; }
                            ++n_cig; --j;
                        }
                    }
                }
            }
        }
        if (ok) {
            for (int t = gl; t < j; t += LPT) { cg[n_cig + t] = 1; cq[n_cig + t] = j - 1 - t; } // unaligned query head
            if (j > 0) n_cig += j;
        }
        // the walk usually ends with the next 32 rows' codes still in flight: they land in the window, which shares its
        // shared memory with the next alignment's row metadata and ring (found by compute-sanitizer racecheck)
        cp_async_wait_all();
        __syncwarp();
    }
    }
    PH(2);
    // ---- merge the alignment into the graph (abpoa_graph.c:1218-1284), LPT path steps at a time ----------
    // A path visits every node and every aligned group at most once, so all steps touch distinct adjacency lists
    // and groups; ids of new nodes/edges are creation-ordered prefix sums, exactly what the sequential walk yields.
    int n_ev = 0;
    {
        const int node_n0 = node_n;
        int last_id = 0, last_new = 0, pend_lo = 0; // events [pend_lo, n_ev) wait for the next match column
        if (!ok) n_cig = 0;
        const unsigned below = (1u << gl) - 1;
        for (int k0 = 0; __any_sync(TH_FULL, ok && k0 < n_cig); k0 += LPT) {
            const bool ma = ok && k0 < n_cig;
            const int k = k0 + gl;
            int op = 2, v = 0, q = 0;
            if (ma && k < n_cig) { const uint32_t cv = w.cigar[n_cig - 1 - k]; op = cv & 3; v = (int)(cv >> 2); q = w.cigq[n_cig - 1 - k]; }
            const bool prod = op != 2;
            int tgt = -1, isnew = 0, bf = 0, bl = 0, an = 0, al[4] = {0, 0, 0, 0};
            uint8_t qb = 0;
            if (prod) qb = query[q];
            if (op == 0) {
                bf = w.n2i[v]; bl = bf; an = min(w.aln_n[v], 4);
                int aid = -1;
                for (int a = 0; a < an; ++a) {
                    al[a] = w.aln[v * 4 + a];
                    const int x = w.n2i[al[a]]; bf = min(bf, x); bl = max(bl, x);
                    if (aid < 0 && w.base[al[a]] == qb) aid = al[a];
                }
                if (w.base[v] == qb) tgt = v; else if (aid >= 0) tgt = aid; else isnew = 1;
            } else if (op == 1) isnew = 1;
            const unsigned newm = g.ballot(isnew), prodm = g.ballot(prod), mm = g.ballot(op == 0);
            const int ev = n_ev + __popc(newm & below); // event id == rank of the new node
            if (isnew) tgt = node_n0 + ev;
            // previous producing step
            int from = last_id, from_new = last_new;
            { const unsigned lower = prodm & below; const int src = lower ? 31 - __clz(lower) : 0;
              const int pt = g.shfl(tgt, src), pnw = g.shfl(isnew, src);
              if (lower) { from = pt; from_new = pnw; } }
            // new nodes first (their adjacency heads must exist before edges are linked)
            int lerr = 0;
            if (isnew) {
                if (tgt >= w.ncap) lerr = 1;
                else { w.base[tgt] = qb; w.out_head[tgt] = w.out_tail[tgt] = w.in_head[tgt] = w.in_tail[tgt] = -1; w.aln_n[tgt] = 0; w.hs[tgt] = 0; w.ev_node[ev] = tgt;
                       w.nrec[tgt] = make_int4(0, -1, -1, -1); w.brec[tgt] = make_int4(-1, 0, 0, 0); }
            }
            if (isnew && op == 0 && an >= 4) lerr = 1; // a column holds at most 5 distinct codes (ACGT + N): 4 aligned nodes per node
            if (g.any(lerr)) { err = TH_ERR_CAP; ok = false; }
            const bool go = ok && ma;
            if (go && prod) w.hs[tgt] += 1;              // reads through the node (a path visits a node once: no two lanes share a target)
            if (go && isnew && op == 0) { // abpoa_add_graph_aligned_node (:1036-1044): all-pairs with the old group
                for (int a = 0; a < an; ++a) { const int y = al[a]; w.aln[y * 4 + w.aln_n[y]] = tgt; w.aln_n[y] += 1; w.aln[tgt * 4 + a] = y; }
                w.aln[v * 4 + an] = tgt; w.aln_n[v] = an + 1; w.aln[tgt * 4 + an] = v; w.aln_n[tgt] = an + 1;
                w.ev_anchor[ev] = bl + 1;
            }
            // insertion events take the first index of the next match column's group
            {
                const int fm = mm ? __ffs(mm) - 1 : 0, bff = g.shfl(bf, fm);
                if (go && mm) {
                    const int hi_ev = n_ev + __popc(newm & ((1u << fm) - 1)); // events created before that step
                    for (int e = pend_lo + gl; e < hi_ev; e += LPT) w.ev_anchor[e] = bff;
                }
            }
            {
                const unsigned higher = mm & ~(below | (1u << gl));
                const int src = higher ? __ffs(higher) - 1 : 0;
                const int bfn = g.shfl(bf, src);
                if (go && op == 1 && higher) w.ev_anchor[ev] = bfn;
            }
            { // events after the last match column of this batch stay pending
                const int lastm = mm ? 31 - __clz(mm) : -1;
                if (mm) pend_lo = n_ev + __popc(newm & ((2u << lastm) - 1));
            }
            __syncwarp();
            // edges (abpoa_add_graph_edge :1063-1106): weight + 1 on an existing edge, else append to both lists
            int found = -1;
            if (go && prod && !from_new && !isnew) { // most steps follow the heaviest edge: look there before walking the list
                const int4 br = w.brec[from];
                if (br.x == tgt) found = br.z;
                else for (int e = w.out_head[from]; e >= 0; e = w.e_no[e]) if (w.e_to[e] == tgt) { found = e; break; }
            }
            const bool mk = go && prod && found < 0;
            const unsigned mkm = g.ballot(mk);
            if (found >= 0) { const int wt = w.e_w[found] + 1; w.e_w[found] = wt; g_note_out(w, from, tgt, found, wt); }
            if (mk) {
                const int e = edge_n + __popc(mkm & below);
                g_note_in(w, tgt, from); g_note_out(w, from, tgt, e, 1);
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
            { const int src = prodm ? 31 - __clz(prodm) : 0;
              const int lt = g.shfl(tgt, src), ln = g.shfl(isnew, src);
              if (prodm) { last_id = lt; last_new = ln; } }
            __syncwarp();
        }
        if (ok) {
            node_n = node_n0 + n_ev;
            if (gl == 0) g_add_edge(w, edge_n, last_id, 1, !last_new);
        }
        edge_n = g.shfl(edge_n, 0);
        if (ok) for (int e = pend_lo + gl; e < n_ev; e += LPT) w.ev_anchor[e] = n - 1; // before the sink
    }
    __syncwarp();
    PH(3);
    // new order: old node at index i moves to i + #(events with anchor <= i); event e lands at anchor_e + e
    if (ok) {
        for (int i = gl; i < n; i += LPT) {
            int lo = 0, hi = n_ev;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (w.ev_anchor[mid] <= i) lo = mid + 1; else hi = mid; }
            w.ord2[i + lo] = w.ord[i];
        }
        for (int e = gl; e < n_ev; e += LPT) w.ord2[w.ev_anchor[e] + e] = w.ev_node[e];
    }
    __syncwarp();
    if (ok && gl == 0) { int32_t *t = w.ord; w.ord = w.ord2; w.ord2 = t; }
    __syncwarp();
    PH(4);
#undef PH
}

// Heaviest-column consensus without the rank DFS.  The reference orders the MSA columns (aligned groups) by a stack-driven
// DFS (abpoa_graph.c:279-339), which is one topological order of the column DAG; `ord` is another one, with every group
// contiguous.  Only columns that EMIT a base matter (max_w >= n_seq - sum_w, :604-629), and two emitted columns that no
// path orders would have disjoint read sets, hence max_w = sum_w = n_seq / 2 in both: unless at least two such columns
// exist (counted in `ambiguous`), every topological order yields the same consensus, and this one is data parallel.
// The weight of a node (popcount of its read ids) is the number of reads through it, kept in hs[] by the merge.
template <int LPT>
__device__ int poa_consensus_fast(PoaSmem<LPT> &sm, const bool act, const int node_n, const int n_seq, uint8_t *cons, int32_t *cov, int &ambiguous) {
    const PoaG<LPT> g;
    const int gl = g.gl;
    PoaWs &w = sm.ws;
    const int n = act ? node_n : 0;
    for (int i = gl; i < n; i += LPT) w.n2i[w.ord[i]] = i;
    __syncwarp();
    int cons_i = 0; ambiguous = 0;
    for (int i0 = 1; __any_sync(TH_FULL, i0 < n - 1); i0 += LPT) {
        const int i = i0 + gl;
        bool sel = false, amb = false; int max_w = 0, node = 0;
        if (i < n - 1) {
            const int v = w.ord[i], an = min(w.aln_n[v], 4);
            bool first = true;
            int sum = 0, key = 0;
            { const int b = w.base[v], hw = w.hs[v]; if (b < 4) { sum += hw; key = (hw << 3) | (7 - b); node = v; } }
            for (int a = 0; a < an; ++a) {
                const int x = w.aln[v * 4 + a];
                if (w.n2i[x] < i) first = false;
                const int b = w.base[x], hw = w.hs[x];
                if (b < 4) { sum += hw; const int k = (hw << 3) | (7 - b); if (k > key) { key = k; node = x; } } // first heaviest base in A, C, G, T order
            }
            max_w = key >> 3;
            const int gap_w = n_seq - sum;
            sel = first && max_w > 0 && max_w >= gap_w;
            amb = sel && max_w == gap_w && sum == max_w;
        }
        const unsigned m = g.ballot(sel), am = g.ballot(amb);
        if (sel) { const int pos = cons_i + __popc(m & ((1u << gl) - 1)); cons[pos] = w.base[node]; cov[pos] = max_w; }
        cons_i += __popc(m); ambiguous += __popc(am);
    }
    return cons_i;
}

// heaviest-column consensus (abpoa_graph.c:279-359, 604-648, 467-478).  All lanes of the warp call; returns cons_len of the
// lane's group (-1: the arena is too small).  The DFS that assigns MSA ranks is order dependent (stack discipline,
// abpoa_graph.c:279-339) and runs on one lane per group; in-degrees, column weights and the column vote are data parallel.
template <int LPT>
__device__ int poa_consensus(PoaSmem<LPT> &sm, const bool act, int node_n, int n_seq, uint8_t *cons, int32_t *cov) {
    const PoaG<LPT> g;
    const int gl = g.gl;
    PoaWs &w = sm.ws;
    int32_t *deg = w.n2i, *stk = w.ord2, *rank = w.ri;
    int32_t *rcw = reinterpret_cast<int32_t *>(w.arena); // 5 x msa_l weights, then 5 x msa_l node ids
    if (!act) node_n = 0;
    for (int i = gl; i < node_n; i += LPT) { int d = 0; for (int e = w.in_head[i]; e >= 0; e = w.e_ni[e]) ++d; deg[i] = d; }
    __syncwarp();
    int msa_l = 0;
    if (act && gl == 0) {
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
                    bool okk = true;
                    const int an = w.aln_n[o];
                    for (int a = 0; a < an; ++a) if (deg[w.aln[o * 4 + a]] != 0) { okk = false; break; }
                    if (!okk) continue;
                    stk[sp++] = o; rank[o] = -1;
                    for (int a = 0; a < an; ++a) { const int x = w.aln[o * 4 + a]; stk[sp++] = x; rank[x] = -1; }
                }
            }
        }
        msa_l = rank[1] - 1;
    }
    msa_l = g.shfl(msa_l, 0);
    __syncwarp();
    int ret = 0;
    if (msa_l <= 0) msa_l = 0;
    else if ((uint64_t)msa_l * 10 * 2 > w.arena_cap) { ret = -1; msa_l = 0; }
    int32_t *nodeid = rcw + 5 * (size_t)msa_l;
    for (int i = gl; i < 5 * msa_l; i += LPT) { rcw[i] = 0; nodeid[i] = 0; }
    __syncwarp();
    // abpoa_set_row_column_weight; popcount(read_ids) == sum of out weights.  A column is one aligned group and a
    // group holds one node per base, so every (column, base) slot has a single writer.
    if (msa_l > 0) for (int i = 2 + gl; i < node_n; i += LPT) {
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
    for (int c0 = 0; __any_sync(TH_FULL, c0 < msa_l); c0 += LPT) { // abpoa_heaviest_column_consensus (:604-629): first heaviest base, kept iff max_w >= gap weight
        const int c = c0 + gl;
        int max_w = 0, max_base = 5, gap_w = n_seq; bool sel = false;
        if (c < msa_l) {
            for (int b = 0; b < 4; ++b) { const int x = rcw[c * 5 + b]; if (x > max_w) { max_base = b; max_w = x; } gap_w -= x; }
            sel = max_w >= gap_w && max_base < 5;
        }
        const unsigned m = g.ballot(sel);
        if (sel) { const int pos = cons_i + __popc(m & ((1u << gl) - 1)); cons[pos] = w.base[nodeid[c * 5 + max_base]]; cov[pos] = max_w; }
        cons_i += __popc(m);
    }
    return ret < 0 ? -1 : cons_i;
}

// persistent warps; every group pulls its tasks from an atomic counter
#ifndef POA_MIN_BLOCKS16
#define POA_MIN_BLOCKS16 5   // 16-lane groups: 20 warps = 40 tasks per SM, 102 registers (6 blocks: +2 %, 7: spills; measured)
#endif
#ifndef POA_MIN_BLOCKS32
#define POA_MIN_BLOCKS32 4
#endif
template <int LPT>
__global__ void __launch_bounds__(POA_WARPS * 32, LPT == 16 ? POA_MIN_BLOCKS16 : POA_MIN_BLOCKS32)
poa_kernel(DevParams P, int n_tasks, const int *__restrict__ n_tasks_dev, const PoaTask *__restrict__ tasks, const int32_t *__restrict__ task_order,
           const int32_t *__restrict__ u_start, const int32_t *__restrict__ u_len, const uint8_t *__restrict__ bseq,
           uint8_t *slabs, size_t slab_bytes, const unsigned long long *__restrict__ slab_bytes_dev, size_t slab_cap, int *task_counter,
           uint8_t *__restrict__ cons_base, int32_t *__restrict__ cons_cov, int32_t *__restrict__ cons_len,
           int32_t *__restrict__ task_status, unsigned long long *__restrict__ stat_cells, unsigned long long *__restrict__ stat_rows,
           unsigned long long *__restrict__ stat_phase, int max_groups, int32_t *__restrict__ retry_list, TaskTotals *tot) {
    constexpr int GPW = 32 / LPT;                      // groups per warp
    extern __shared__ __align__(16) unsigned char poa_smem_raw[];
    PoaSmem<LPT> *s_all = reinterpret_cast<PoaSmem<LPT> *>(poa_smem_raw);
    PoaLaneK *lk_tab = reinterpret_cast<PoaLaneK *>(s_all + POA_WARPS * GPW);
    poa_fill_lane_k(lk_tab, LPT, P);
    __syncthreads();
    const PoaG<LPT> g;
    const int gl = g.gl, wib = threadIdx.x >> 5, gib = wib * GPW + (g.gofs ? 1 : 0);
    PoaSmem<LPT> &sm = s_all[gib];
    const uint32_t lk = (uint32_t)__cvta_generic_to_shared(lk_tab + gl);
    const int gg = blockIdx.x * (POA_WARPS * GPW) + gib;
    if (n_tasks_dev) { // second pass: the first one left the task count and the slab size these tasks need in device memory
        n_tasks = *n_tasks_dev;
        slab_bytes = ((size_t)*slab_bytes_dev + 255) & ~(size_t)255;
        max_groups = (int)min((unsigned long long)max_groups, slab_bytes ? (unsigned long long)(slab_cap / slab_bytes) : 0ull); // no room for a single slab: the tasks keep their status
    }
    uint8_t *slab = slabs + (size_t)gg * slab_bytes;
    unsigned long long cells = 0, rows = 0;
#ifdef POA_PROFILE
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_all = clock64();
#else
    long long *ph = nullptr;
#endif
    // group state
    bool has = false, done = gg >= max_groups; // groups beyond max_groups have no slab
    int t = 0, s = 0, node_n = 0, edge_n = 0, err = TH_OK;
    PoaTask T; T.n_seqs = 0; T.unit_off = 0; T.seq_off = 0; T.cons_off = 0; T.ncap = 0; T.qmax = 0; T.read = 0;
    while (true) {
        // ---- groups without a task fetch one -----------------------------------------------------------------------
        bool fresh = false;
        {
            int ti = 0;
            if (!has && !done && gl == 0) ti = atomicAdd(task_counter, 1);
            ti = g.shfl(ti, 0);
            if (!has && !done) {
                if (ti >= n_tasks) done = true;
                else { t = task_order ? task_order[ti] : ti; T = tasks[t]; has = true; fresh = true; s = 1; err = TH_OK; }
            }
        }
        if (!__any_sync(TH_FULL, has)) break;
        const uint8_t *rseq = bseq + T.seq_off;
        uint8_t *cons = cons_base + T.cons_off; int32_t *cov = cons_cov + T.cons_off;
        if (fresh) { // tasks that need no graph
            if (T.n_seqs < 2) { if (gl == 0) { cons_len[t] = 0; task_status[t] = TH_ERR_CAP; } has = false; } // the reference aborts here (abpoa_cons.c:58)
            else if (T.n_seqs <= 2) { // src/abpoa_cons.c:57-80: the first unit verbatim
                const int l0 = u_len[T.unit_off]; const uint8_t *s0 = rseq + u_start[T.unit_off];
                for (int i = gl; i < l0; i += LPT) { cons[i] = s0[i]; cov[i] = 0; }
                if (gl == 0) { cons_len[t] = l0; task_status[t] = TH_OK; }
                has = false;
            } else if (poa_fixed_bytes(T.ncap, T.qmax, T.n_seqs) + 4096 > slab_bytes) {
                if (gl == 0) {
                    cons_len[t] = 0; task_status[t] = TH_ERR_ARENA;
                    if (retry_list) { retry_list[atomicAdd(&tot->retry_n, 1)] = t; atomicMax(&tot->slab_full, poa_slab_need(T.ncap, T.qmax, T.n_seqs, true)); }
                }
                has = false;
            }
            if (!has) fresh = false;
        }
        __syncwarp();
        if (fresh && gl == 0) poa_carve(sm.ws, slab, slab_bytes, T.ncap, T.qmax, T.n_seqs);
        __syncwarp();
        if (fresh) { // first sequence: a chain of new nodes (abpoa_graph.c:1108-1124)
            PoaWs &w = sm.ws;
            const int l0 = u_len[T.unit_off];
            for (int i = gl; i < l0 + 2; i += LPT) { w.out_head[i] = w.out_tail[i] = w.in_head[i] = w.in_tail[i] = -1; w.aln_n[i] = 0; }
        }
        __syncwarp();
        if (fresh) {
            PoaWs &w = sm.ws;
            const int l0 = u_len[T.unit_off]; const uint8_t *s0 = rseq + u_start[T.unit_off];
            for (int i = gl; i <= l0; i += LPT) {
                const int from = i == 0 ? 0 : 1 + i, to = i == l0 ? 1 : 2 + i;
                w.e_to[i] = to; w.e_from[i] = from; w.e_w[i] = 1; w.e_no[i] = -1; w.e_ni[i] = -1;
                w.out_head[from] = w.out_tail[from] = i; w.in_head[to] = w.in_tail[to] = i;
                w.nrec[to] = make_int4(1, from, -1, -1); w.brec[from] = make_int4(to, 1, i, 0);
                if (i < l0) { w.base[2 + i] = s0[i]; w.ord[1 + i] = 2 + i; w.hs[2 + i] = 1; }
            }
            if (gl == 0) { w.ord[0] = 0; w.ord[l0 + 1] = 1; w.base[0] = w.base[1] = 4; w.nrec[0] = make_int4(0, -1, -1, -1); w.brec[1] = make_int4(-1, 0, 0, 0); }
            node_n = l0 + 2; edge_n = l0 + 1;
        }
        __syncwarp();
        // ---- one more sequence for every group that has a task -----------------------------------------------------
        if (__any_sync(TH_FULL, has)) {
            const bool act = has && err == TH_OK && s < T.n_seqs;
            const uint8_t *q = rseq; int ql = 0;
            if (act) { q = rseq + u_start[T.unit_off + s]; ql = u_len[T.unit_off + s]; }
            if (!P.affine) poa_add_sequence<LPT, false>(sm, P, lk, act, q, ql, node_n, edge_n, err, cells, rows, ph);
            else poa_add_sequence<LPT, true>(sm, P, lk, act, q, ql, node_n, edge_n, err, cells, rows, ph);
            if (act) ++s;
        }
        // ---- finished tasks: consensus -----------------------------------------------------------------------------
        const bool fin = has && (err != TH_OK || s >= T.n_seqs);
        if (__any_sync(TH_FULL, fin)) {
#ifdef POA_PROFILE
            long long t_c0 = clock64();
#endif
            int amb = 0;
            int cl = poa_consensus_fast<LPT>(sm, fin && err == TH_OK, node_n, T.n_seqs, cons, cov, amb);
            if (__any_sync(TH_FULL, fin && err == TH_OK && amb >= 2)) { // column order not settled by the graph: the reference's DFS decides
                const bool dfs = fin && err == TH_OK && amb >= 2;
                const int cl2 = poa_consensus<LPT>(sm, dfs, node_n, T.n_seqs, cons, cov);
                if (dfs) cl = cl2;
            }
            if (fin) {
                if (err == TH_OK && cl < 0) err = TH_ERR_ARENA;
                if (err != TH_OK) cl = 0;
                if (gl == 0) {
                    cons_len[t] = cl; task_status[t] = err;
                    if (retry_list && (err == TH_ERR_ARENA || err == TH_ERR_CAP || err == TH_ERR_LEN)) { // a full-width slab, 32-lane groups and 32-bit scores take it in the second pass
                        retry_list[atomicAdd(&tot->retry_n, 1)] = t;
                        atomicMax(&tot->slab_full, poa_slab_need(T.ncap, T.qmax, T.n_seqs, true));
                    }
                }
                has = false;
            }
#ifdef POA_PROFILE
            ph[5] += clock64() - t_c0;
#endif
        }
        __syncwarp();
    }
    if (gl == 0 && cells) { atomicAdd(stat_cells, cells); atomicAdd(stat_rows, rows); }
#ifdef POA_PROFILE
    if ((threadIdx.x & 31) == 0) {
        ph[6] = clock64() - t_all;
        for (int k = 0; k < 7; ++k) atomicAdd(stat_phase + k, (unsigned long long)ph[k]);
    }
#endif
}
