// th_ksw.cuh -- warp-systolic affine-gap alignment with ksw2's exact tie-breaking, no traceback matrix.
//
// Replaces ksw_extz2_sse (ksw2/ksw2_extz2_sse.c:23-304) + ksw_backtrack (ksw2/ksw2.h:119-151) as used
// through src/ksw2_align.c:62-173: scoring +1/-2, N = -1, gap 2+1*g, full matrix (w = -1, zdrop = -1).
//
// One warp per alignment.  Lane l owns C consecutive query columns of a 32*C-column block and walks
// down the target rows one step behind lane l-1 (systolic wavefront); H/F of the block's last column
// are handed to the next lane by shuffle, and to the next column block through a per-warp boundary
// array in global memory (16 B per target row), so memory is O(tlen) for any size.
//
// Instead of storing ksw2's 5-bit direction matrix and walking it backwards, every DP state carries
// the value its traceback would produce ("payload"): the traceback parent of each (cell, state) is a
// pure function of the cell's own scores (diag unless E > diag, then F unless F > max; E/F continue
// iff strictly better than re-opening), so quantities that are sums along the traced path can be
// pushed forward with the scores.  Payloads:
//   iden  : number of M columns with equal codes (ksw2_get_xid, src/ksw2_align.c:62-86)
//   istop : target bases consumed when the path has consumed X = qlen - q_left_ext query bases, i.e.
//           tlen - ksw2_backtrack_left_end(...) (src/ksw2_align.c:88-115)
// Extension mode keeps no payload but reproduces the reference's arg-max visiting order
// (anti-diagonal by anti-diagonal, last cell first, four interleaved lanes, tail; ksw2_extz2_sse.c:224-261).
#pragma once
#include "th_common.cuh"

#define KSW_Q 2
#define KSW_E 1
enum { KSW_GLOBAL = 0, KSW_GLOBAL_STOP = 1, KSW_EXT = 2 };

// rank of target index t inside anti-diagonal r (lower = visited earlier by the reference)
__device__ __forceinline__ long long ksw_diag_rank(int t, int r, int ql, int tl) {
    int st0 = r - ql + 1 > 0 ? r - ql + 1 : 0, en0 = r < tl - 1 ? r : tl - 1;
    if (t == en0) return 0;
    int en1 = st0 + (en0 - st0) / 4 * 4;
    if (t < en1) return 1 + (long long)((t - st0) & 3) * 0x40000000ll + (t - st0) / 4;
    return 1 + 4 * 0x40000000ll + (t - en1);
}
__device__ __forceinline__ bool ksw_ext_better(int z1, int i1, int j1, int z2, int i2, int j2, int ql, int tl) {
    if (z1 != z2) return z1 > z2;
    if (i2 < 0) return true;
    int r1 = i1 + j1, r2 = i2 + j2;
    if (r1 != r2) return r1 < r2;
    return ksw_diag_rank(i1, r1, ql, tl) < ksw_diag_rank(i2, r2, ql, tl);
}

// bnd: 2*tl int4 entries of per-warp scratch.  Results: GLOBAL -> out0 = iden; GLOBAL_STOP -> out0 =
// iden, out1 = t_left_ext; EXT -> out0 = max_q, out1 = max_t.  All lanes must call; all get the result.
template <int MODE, int C>
__device__ void ksw_warp(const uint8_t *q, int ql, const uint8_t *t, int tl, int X,
                         int4 *bnd, int &out0, int &out1) {
    const int lane = lane_id();
    out0 = MODE == KSW_EXT ? -1 : 0; out1 = MODE == KSW_EXT ? -1 : 0;
    if (ql <= 0 || tl <= 0) return;
    // Extension: only cells with a positive score can become the maximum, and a cell (i, j) with i > j scores at most
    // (j + 1) matches minus a gap of i - j bases = 2 j - 1 - i (likewise with i and j swapped), so rows from 2 ql on and
    // columns from 2 tl on are never positive: they are not computed.  The visiting order still refers to the full matrix.
    const int qlr = ql, tlr = tl;
    if (MODE == KSW_EXT) { ql = min(qlr, 2 * tlr); tl = min(tlr, 2 * qlr); }
    const int BW = 32 * C;
    const int nblk = (ql + BW - 1) / BW;
    int bestz = 0, besti = -1, bestj = -1;
    int resz = 0, resp = 0;
    for (int b = 0; b < nblk; ++b) {
        const int jb = b * BW;
        const int bw = min(ql - jb, BW), nl = (bw + C - 1) / C;
        const int j0 = jb + lane * C;
        const int4 *bin = bnd + (size_t)(b & 1) * tl;
        int4 *bout = bnd + (size_t)((b + 1) & 1) * tl;
        int Hp[C], Ea[C], pH[C], pE[C], qb[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = j0 + c;
            Hp[c] = -(KSW_Q + KSW_E * (j + 1));
            Ea[c] = Hp[c] - KSW_Q - KSW_E;
            pH[c] = 0; pE[c] = 0;
            qb[c] = j < ql ? q[j] : 6; // sentinel: columns past the query end only produce dead values
        }
        int hdiag = j0 == 0 ? 0 : -(KSW_Q + KSW_E * j0), phdiag = 0;
        int oH = 0, oF = 0, oPH = 0, oPF = 0;
        const int nstep = tl + nl - 1;
        for (int s = 0; s < nstep; ++s) {
            const int i = s - lane;
            int iH = __shfl_up_sync(TH_FULL, oH, 1), iF = __shfl_up_sync(TH_FULL, oF, 1);
            int iPH = 0, iPF = 0;
            if (MODE != KSW_EXT) { iPH = __shfl_up_sync(TH_FULL, oPH, 1); iPF = __shfl_up_sync(TH_FULL, oPF, 1); }
            if (lane == 0) {
                if (b == 0) { iH = -(KSW_Q + KSW_E * (s + 1)); iF = iH - KSW_Q - KSW_E; iPH = 0; iPF = 0; }
                else if (s < tl) { int4 v = bin[s]; iH = v.x; iF = v.y; iPH = v.z; iPF = v.w; }
            }
            if (i >= 0 && i < tl && lane < nl) {
                const int tb = t[i];
                int hd = hdiag, phd = phdiag, F = iF, pF = iPF;
                hdiag = iH; phdiag = iPH;
                int rowk = INT_MIN;
                const int vcols = ql - j0; // columns past the query end hold dead values
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int eq = tb == qb[c];
                    const int sc = ((tb | qb[c]) & 4) ? -KSW_E : (eq ? 1 : -2);
                    int z = hd + sc, pz = 0;
                    const int e = Ea[c];
                    if (MODE != KSW_EXT) {
                        pz = phd + eq;
                        if (MODE == KSW_GLOBAL_STOP && j0 + c == X) pz = (pz & 0xffff) | (i << 16);
                        if (e > z) pz = pE[c];
                    }
                    z = max(z, e);
                    if (MODE != KSW_EXT) { if (F > z) pz = pF; }
                    z = max(z, F);
                    const int t1 = z - KSW_Q;
                    if (MODE != KSW_EXT) {
                        pE[c] = e > t1 ? pE[c] : pz;
                        pF = F > t1 ? pF : pz;
                        if (MODE == KSW_GLOBAL_STOP && j0 + c + 1 == X) pF = (pF & 0xffff) | ((i + 1) << 16);
                    }
                    Ea[c] = max(e, t1) - KSW_E;
                    F = max(F, t1) - KSW_E;
                    hd = Hp[c]; phd = pH[c];
                    Hp[c] = z; pH[c] = pz;
                    if (MODE == KSW_EXT) { // row maximum of this lane's columns, first column on ties (smaller j = earlier anti-diagonal)
                        if (c < vcols) rowk = max(rowk, z * 32 + (31 - c));
                    }
                }
                if (MODE == KSW_EXT) {
                    const int zr = rowk >> 5; // arithmetic shift: floor, exact because 0 <= 31 - c < 32
                    if (zr > 0 && zr >= bestz) {
                        const int jr = j0 + 31 - (rowk & 31);
                        if (ksw_ext_better(zr, i, jr, bestz, besti, bestj, qlr, tlr)) { bestz = zr; besti = i; bestj = jr; }
                    }
                }
                oH = Hp[C - 1]; oF = F; oPH = pH[C - 1]; oPF = pF;
                if (lane == nl - 1 && b + 1 < nblk) bout[i] = make_int4(oH, oF, oPH, oPF);
            }
        }
        if (b == nblk - 1 && MODE != KSW_EXT) { // cell (tl-1, ql-1) lives in lane nl-1, column (ql-1-jb) % C
            const int cc = (ql - 1 - jb) - (nl - 1) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) if (c == cc) { resz = Hp[c]; resp = pH[c]; }
            resz = __shfl_sync(TH_FULL, resz, nl - 1); resp = __shfl_sync(TH_FULL, resp, nl - 1);
        }
        __syncwarp();
    }
    if (MODE == KSW_EXT) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            int oz = __shfl_xor_sync(TH_FULL, bestz, d), oi = __shfl_xor_sync(TH_FULL, besti, d), oj = __shfl_xor_sync(TH_FULL, bestj, d);
            if (oi >= 0 && (besti < 0 || ksw_ext_better(oz, oi, oj, bestz, besti, bestj, qlr, tlr))) { bestz = oz; besti = oi; bestj = oj; }
        }
        out0 = bestj; out1 = besti;
    } else {
        (void)resz;
        out0 = MODE == KSW_GLOBAL_STOP ? (resp & 0xffff) : resp;
        if (MODE == KSW_GLOBAL_STOP) out1 = tl - (int)((unsigned)resp >> 16);
    }
}

// ---------------------------------------------------------------------------------------------
// Two global alignments in one warp (usually against the same target): alignment A in the low, B in the high 16 bits of
// every DP word (scores and identity payloads fit int16 for sequences up to KSW2_MAXLEN), so one VIMNMX.S16x2
// (with its per-half "which operand won" predicates steering the payload selects) updates two cells.  Same
// recurrences, boundary values and tie-breaking as ksw_warp<KSW_GLOBAL>; sequences must not contain N (the
// caller checks and falls back), because the score is computed as 1 - 3 * min(q ^ t, 1).
// ---------------------------------------------------------------------------------------------
#define KSW2_MAXLEN 8000
#define KSW2_BIAS 16400   // > 2 * KSW2_MAXLEN + gap costs, and BIAS + KSW2_MAXLEN < 32768
__device__ __forceinline__ uint32_t ksw_sel2(uint32_t a, uint32_t b, bool ph, bool pl) { // (ph ? a : b).hi, (pl ? a : b).lo
    uint32_t r = b;
    if (pl) r = __byte_perm(r, a, 0x3254);
    if (ph) r = __byte_perm(r, a, 0x7610);
    return r;
}
__device__ __forceinline__ uint32_t pk2(int v) { return ((uint32_t)(uint16_t)v) * 0x10001u; }

// Banded mode (Bu >= 0): column block b (columns jb .. jb + 32 C - 1) only computes the target rows jb - Bu .. jb + 32 C + Bl - 1,
// a staircase of rectangles around the diagonals -Bl .. Bu (d = column - row).  The cells left out are replaced by values
// that are lower bounds of the true ones (the score of "gap along the first row, then gap down the column",
// -(i + j + 6) in ksw2's scoring), so every banded value is at most the full matrix's and at least the best path that stays
// inside the band.  A path that touches a diagonal d > max(0, n - m) has at least d inserted query bases, as many target
// bases deleted again minus (n - m), and two gap openings: it scores at most 2 n - m - 4 - 3 d (n = query, m = target
// length); likewise 2 m - n - 4 - 3 |d| below the band.  CERTIFICATE: if the banded score is strictly above both bounds
// for d = Bu + 1 and d = -(Bl + 1), no optimal or co-optimal path leaves the band, so the score, every traceback decision
// on the optimal path (a competing candidate that reached the cell through a left-out cell would complete to a path
// above the bound) and with them the identity payload equal the full matrix's.  The caller checks it (ksw_band_certified)
// and widens the band or runs the full matrix when it fails.  scA / scB = score of the final cell (INT_MIN when the
// final cell was not computed).  Checked on the CPU against the oracle by tools/ksw_band_check.py (tools/sim/ksw_band_sim.c
// is the scalar model of this routine).  Requires, for both alignments, Bu >= max(0, ql - tl) and Bl >= max(0, tl - ql).
template <int C>
__device__ void ksw_warp_global2(const uint8_t *qa, int qla, const uint8_t *ta, int tla, const uint8_t *qb_, int qlb, const uint8_t *tb_, int tlb,
                                 int4 *bnd, const int Bu, const int Bl, int &idenA, int &idenB, int &scA, int &scB) {
    const int lane = lane_id();
    idenA = 0; idenB = 0; scA = INT_MIN; scB = INT_MIN;
    if (qla <= 0 || tla <= 0) { qla = 0; tla = 0; }
    if (qlb <= 0 || tlb <= 0) { qlb = 0; tlb = 0; }
    const int ql = max(qla, qlb), tl = max(tla, tlb);
    if (ql <= 0 || tl <= 0) return;
    const bool full = Bu < 0;
    uint32_t capA = 0, capB = 0;   // payload of cell (tl_x - 1, ql_x - 1), captured when its row is computed
    uint32_t capHA = 0, capHB = 0; // its (biased) score; 0 = not computed (biased scores are >= 1)
    const int BW = 32 * C;
    const int nblk = (ql + BW - 1) / BW;
    const int blkA = qla > 0 ? (qla - 1) / BW : -1, blkB = qlb > 0 ? (qlb - 1) / BW : -1;
    // Scores are kept biased by +KSW2_BIAS per half, so every half stays in [1, 32767]: adding small per-half
    // deltas with ordinary 32-bit integer arithmetic can then neither borrow nor carry across the halves, and the
    // compiler is free to put those adds on the FMA pipe (IMAD) while the 16x2 max ops take the ALU pipe.
    // (Banded mode keeps the range: every computed cell's E is >= -(i + j + 8), by induction down its column.)
    const uint32_t ONE2 = 0x00010001u, Q2 = (uint32_t)KSW_Q * 0x10001u, E2 = (uint32_t)KSW_E * 0x10001u;
    int p_hi = 0;                  // rows [.., p_hi) of the previous column block were computed
    for (int b = 0; b < nblk; ++b) {
        const int jb = b * BW;
        const int bw = min(ql - jb, BW), nl = (bw + C - 1) / C;
        const int j0 = jb + lane * C;
        const int r_lo = full ? 0 : max(0, jb - Bu), r_hi = full ? tl : min(tl, jb + BW + Bl);
        if (r_lo >= r_hi) break;   // the caller's band always reaches the last block; kept as a guard (no certificate then)
        const int4 *bin = bnd + (size_t)(b & 1) * tl;
        int4 *bout = bnd + (size_t)((b + 1) & 1) * tl;
        uint32_t Hp[C], Ea[C], pH[C], pE[C], qq[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = j0 + c;
            const int h0 = KSW2_BIAS - (r_lo == 0 ? KSW_Q + KSW_E * (j + 1) : r_lo + j + 5); // row r_lo - 1: the matrix's first row, or the lower bound
            Hp[c] = pk2(h0); Ea[c] = pk2(h0 - KSW_Q - KSW_E);
            pH[c] = 0; pE[c] = 0;
            const uint32_t a = j < qla ? qa[j] : 8u, bb = j < qlb ? qb_[j] : 8u; // 8 never equals a target code
            qq[c] = a | bb << 16;
        }
        uint32_t hdiag, phdiag = 0;    // cell (r_lo - 1, j0 - 1)
        if (r_lo == 0) hdiag = j0 == 0 ? pk2(KSW2_BIAS) : pk2(KSW2_BIAS - (KSW_Q + KSW_E * j0));
        else if (lane == 0) { const int4 v = bin[r_lo - 1]; hdiag = (uint32_t)v.x; phdiag = (uint32_t)v.z; } // computed by the previous block
        else hdiag = pk2(KSW2_BIAS - (r_lo + j0 + 4));
        uint32_t oH = 0, oF = 0, oPH = 0, oPF = 0;
        const int nstep = (r_hi - r_lo) + nl - 1;
        for (int s = 0; s < nstep; ++s) {
            const int i = r_lo + s - lane;
            uint32_t iH = __shfl_up_sync(TH_FULL, oH, 1), iF = __shfl_up_sync(TH_FULL, oF, 1);
            uint32_t iPH = __shfl_up_sync(TH_FULL, oPH, 1), iPF = __shfl_up_sync(TH_FULL, oPF, 1);
            if (lane == 0) {
                if (b == 0) { const int h0 = KSW2_BIAS - (KSW_Q + KSW_E * (i + 1)); iH = pk2(h0); iF = pk2(h0 - KSW_Q - KSW_E); iPH = 0; iPF = 0; }
                else if (i < p_hi) { const int4 v = bin[i]; iH = (uint32_t)v.x; iF = (uint32_t)v.y; iPH = (uint32_t)v.z; iPF = (uint32_t)v.w; }
                else if (i < r_hi) { const int h0 = KSW2_BIAS - (i + jb + 5); iH = pk2(h0); iF = pk2(h0 - KSW_Q - KSW_E); iPH = 0; iPF = 0; } // below the previous block's rows
            }
            if (i >= r_lo && i < r_hi && lane < nl) {
                const uint32_t tb2 = (i < tla ? (uint32_t)ta[i] : 9u) | (i < tlb ? (uint32_t)tb_[i] : 9u) << 16; // 9: past the end, never equal
                uint32_t hd = hdiag, phd = phdiag, F = iF, pF = iPF;
                hdiag = iH; phdiag = iPH;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const uint32_t mm = __vminu2(qq[c] ^ tb2, ONE2);           // 0 = equal, 1 = different, per half
                    uint32_t z = hd + ONE2 - 3u * mm;                          // + (1 - 3 mm) per half
                    uint32_t pz = phd + ONE2 - mm;                             // identity count + (1 - mm)
                    const uint32_t e = Ea[c];
                    bool gh, gl;
                    z = __vibmax_s16x2(z, e, &gh, &gl); pz = ksw_sel2(pz, pE[c], gh, gl);   // E wins only if strictly greater
                    z = __vibmax_s16x2(z, F, &gh, &gl); pz = ksw_sel2(pz, pF, gh, gl);      // then F, again strictly
                    const uint32_t t1 = z - Q2;
                    uint32_t m = __vibmax_s16x2(t1, e, &gh, &gl); pE[c] = ksw_sel2(pz, pE[c], gh, gl); Ea[c] = m - E2;
                    m = __vibmax_s16x2(t1, F, &gh, &gl); pF = ksw_sel2(pz, pF, gh, gl); F = m - E2;
                    hd = Hp[c]; phd = pH[c];
                    Hp[c] = z; pH[c] = pz;
                }
                oH = Hp[C - 1]; oF = F; oPH = pH[C - 1]; oPF = pF;
                if (lane == nl - 1 && b + 1 < nblk) bout[i] = make_int4((int)oH, (int)oF, (int)oPH, (int)oPF);
                if (i == tla - 1 && b == blkA) {
                    const int cc = (qla - 1 - jb) - lane * C;
#pragma unroll
                    for (int c = 0; c < C; ++c) if (c == cc) { capA = pH[c]; capHA = Hp[c]; }
                }
                if (i == tlb - 1 && b == blkB) {
                    const int cc = (qlb - 1 - jb) - lane * C;
#pragma unroll
                    for (int c = 0; c < C; ++c) if (c == cc) { capB = pH[c]; capHB = Hp[c]; }
                }
            }
        }
        if (b == blkA) {
            idenA = (int)(__shfl_sync(TH_FULL, capA, (qla - 1 - jb) / C) & 0xffffu);
            const int h = (int)(__shfl_sync(TH_FULL, capHA, (qla - 1 - jb) / C) & 0xffffu);
            if (h) scA = h - KSW2_BIAS;
        }
        if (b == blkB) {
            idenB = (int)(__shfl_sync(TH_FULL, capB, (qlb - 1 - jb) / C) >> 16);
            const int h = (int)(__shfl_sync(TH_FULL, capHB, (qlb - 1 - jb) / C) >> 16);
            if (h) scB = h - KSW2_BIAS;
        }
        p_hi = r_hi;
        __syncwarp();
    }
}

// The band's certificate for one alignment (query length n, target length m, banded score sc): see ksw_warp_global2.
__device__ __forceinline__ bool ksw_band_certified(int n, int m, int sc, int Bu, int Bl) {
    if (n <= 0 || m <= 0) return true;
    if (sc == INT_MIN) return false;
    return sc > max(2 * n - m - 4 - 3 * (Bu + 1), 2 * m - n - 4 - 3 * (Bl + 1));
}
// Smallest half-widths (above max(0, n - m), below max(0, m - n)) whose certificate a score >= sc passes.
__device__ __forceinline__ void ksw_band_needed(int n, int m, int sc, int &bu, int &bl) {
    bu = 0; bl = 0;
    if (n <= 0 || m <= 0) return;
    const int nu = 2 * n - m - 4 - sc, nl = 2 * m - n - 4 - sc;   // 3 (B + 1) must exceed these
    bu = max(0, nu >= 0 ? nu / 3 : 0); bl = max(0, nl >= 0 ? nl / 3 : 0);
}
// target rows the banded routine walks, summed over the column blocks (full matrix: blocks x tl)
__device__ __forceinline__ int ksw_band_rows(int ql, int tl, int BW, int Bu, int Bl) {
    int rows = 0;
    for (int jb = 0; jb < ql; jb += BW) rows += max(0, min(tl, jb + BW + Bl) - max(0, jb - Bu));
    return rows;
}
// cells of an n x m matrix inside those row ranges
__device__ __forceinline__ unsigned long long ksw_band_cells(int n, int m, int BW, int Bu, int Bl) {
    unsigned long long cells = 0;
    for (int jb = 0; jb < n; jb += BW) cells += (unsigned long long)max(0, min(m, jb + BW + Bl) - min(m, max(0, jb - Bu))) * (unsigned)min(BW, n - jb);
    return cells;
}

#ifndef KSW_BANDED
#define KSW_BANDED 1          // identity alignments try a certified band before the full matrix (0: always the full matrix)
#endif
#ifndef KSW_BAND_MARGIN
#define KSW_BAND_MARGIN 0.04f // added to what the last certified pair needed (a failed band costs a second pass, a wide one a few rows)
#endif
#ifndef KSW_BAND_ALPHA0
#define KSW_BAND_ALPHA0 0.19f // first band half-width / length: what 15 % divergence needs (tools/ksw_band_check.py)
#endif
// Identity counts of two global alignments (unit vs consensus, src/gen_cons.c:208-216): banded first, as wide as the warp's
// recent alignments needed plus a margin (`alpha` = half-width / length, carried from pair to pair by the caller); a failed
// certificate tells the width that is enough, because the banded score is a lower bound of the true one; the full matrix is
// the last resort, and the first choice where the band would leave out less than an eighth of the rows.  The results never
// depend on `alpha`.  `path` (optional): 1 = certified at the first width, 2 = at the second, 0 = full matrix only, 3 / 4 =
// full matrix after one / two failed bands.  ncell counts the cells computed, failed attempts included.
template <int C>
__device__ void ksw_pair_identity(const uint8_t *qa, int qla, const uint8_t *ta, int tla, const uint8_t *qb, int qlb, const uint8_t *tb, int tlb,
                                  int4 *bnd, float &alpha, int &idenA, int &idenB, unsigned long long &ncell, int *path = nullptr) {
    int s0 = 0, s1 = 0, tried = 0;
    const int BW = 32 * C, ql = max(qla, qlb), tl = max(tla, tlb), rows_full = ((ql + BW - 1) / BW) * tl;
    const int dq = max(0, max(qla - tla, qlb - tlb)), dt = max(0, max(tla - qla, tlb - qlb));
    int Bu = dq + 16 + (int)(alpha * (float)max(ql, tl)), Bl = dt + 16 + (int)(alpha * (float)max(ql, tl));
    bool banded = KSW_BANDED && qla > 0 && qlb > 0 && tla > 0 && tlb > 0, done = false;
    while (true) { // one call site: the routine is inlined once
        banded = banded && tried < 2 && ksw_band_rows(ql, tl, BW, Bu, Bl) * 8 <= rows_full * 7; // else not worth it: the full matrix
        ksw_warp_global2<C>(qa, qla, ta, tla, qb, qlb, tb, tlb, bnd, banded ? Bu : -1, banded ? Bl : -1, idenA, idenB, s0, s1);
        if (!banded) { ncell += (unsigned long long)max(qla, 0) * max(tla, 0) + (unsigned long long)max(qlb, 0) * max(tlb, 0); break; }
        ncell += ksw_band_cells(qla, tla, BW, Bu, Bl) + ksw_band_cells(qlb, tlb, BW, Bu, Bl);
        ++tried;
        if (s0 == INT_MIN || s1 == INT_MIN) { banded = false; continue; }
        done = ksw_band_certified(qla, tla, s0, Bu, Bl) && ksw_band_certified(qlb, tlb, s1, Bu, Bl);
        int ua, la, ub, lb;
        ksw_band_needed(qla, tla, s0, ua, la); ksw_band_needed(qlb, tlb, s1, ub, lb);
        const int nu = max(ua, ub), nl = max(la, lb);
        if (done) { // what this pair needed, relative to its length, steers the next one
            const float need = (float)max(max(nu - dq, nl - dt), 0) / (float)max(ql, tl);
            alpha = 0.5f * alpha + 0.5f * fminf(fmaxf(need + KSW_BAND_MARGIN, 0.05f), 0.5f);
            break;
        }
        Bu = max(Bu, nu) + 8; Bl = max(Bl, nl) + 8; alpha = fminf(alpha + 0.03f, 0.5f);
    }
    if (path) *path = done ? tried : (tried ? 2 + tried : 0);
}

// ---------------------------------------------------------------------------------------------
// Two score-only extensions (boundary extensions of consensus sequences, src/gen_cons.c:217-223; any two, the caller pairs
// extensions of similar target length) in one warp.  Extension A lives in the
// low, B in the high 16 bits of every DP word, biased like ksw_warp_global2; the row maximum of a lane's columns is one
// packed max per column, and candidates are ranked exactly as ksw_warp<KSW_EXT> ranks them (the reference's visiting order).  No N allowed
// (the caller falls back to the single-alignment routine), lengths <= KSW2_MAXLEN.
// ---------------------------------------------------------------------------------------------
template <int C>
__device__ void ksw_warp_ext2(const uint8_t *qa, int qla, const uint8_t *ta, int tla, const uint8_t *qb_, int qlb, const uint8_t *tb_, int tlb,
                              int4 *bnd, int &mqA, int &mtA, int &mqB, int &mtB) {
    const int lane = lane_id();
    mqA = mtA = mqB = mtB = -1;
    if (qla <= 0 || tla <= 0) { qla = 0; tla = 0; }
    if (qlb <= 0 || tlb <= 0) { qlb = 0; tlb = 0; }
    // rows from 2 ql on and columns from 2 tl on cannot hold a positive score (see ksw_warp): not computed
    const int qlra = qla, tlra = tla, qlrb = qlb, tlrb = tlb;
    qla = min(qlra, 2 * tlra); tla = min(tlra, 2 * qlra); qlb = min(qlrb, 2 * tlrb); tlb = min(tlrb, 2 * qlrb);
    const int ql = max(qla, qlb), tl = max(tla, tlb);
    if (ql <= 0 || tl <= 0) return;
    const int BW = 32 * C;
    const int nblk = (ql + BW - 1) / BW;
    int bzA = 0, biA = -1, bjA = -1, bzB = 0, biB = -1, bjB = -1;
    const uint32_t ONE2 = 0x00010001u, Q2 = (uint32_t)KSW_Q * 0x10001u, E2 = (uint32_t)KSW_E * 0x10001u;
    int2 *bnd2 = reinterpret_cast<int2 *>(bnd);
    for (int b = 0; b < nblk; ++b) {
        const int jb = b * BW;
        const int bw = min(ql - jb, BW), nl = (bw + C - 1) / C;
        const int j0 = jb + lane * C;
        const int2 *bin = bnd2 + (size_t)(b & 1) * tl;
        int2 *bout = bnd2 + (size_t)((b + 1) & 1) * tl;
        uint32_t Hp[C], Ea[C], qq[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = j0 + c;
            const int h0 = KSW2_BIAS - (KSW_Q + KSW_E * (j + 1));
            Hp[c] = pk2(h0); Ea[c] = pk2(h0 - KSW_Q - KSW_E);
            const uint32_t a = j < qla ? qa[j] : 8u, bb = j < qlb ? qb_[j] : 8u; // 8 never equals a target code
            qq[c] = a | bb << 16;
        }
        uint32_t hdiag = j0 == 0 ? pk2(KSW2_BIAS) : pk2(KSW2_BIAS - (KSW_Q + KSW_E * j0));
        uint32_t oH = 0, oF = 0;
        const int nstep = tl + nl - 1;
        for (int s = 0; s < nstep; ++s) {
            const int i = s - lane;
            uint32_t iH = __shfl_up_sync(TH_FULL, oH, 1), iF = __shfl_up_sync(TH_FULL, oF, 1);
            if (lane == 0) {
                if (b == 0) { const int h0 = KSW2_BIAS - (KSW_Q + KSW_E * (s + 1)); iH = pk2(h0); iF = pk2(h0 - KSW_Q - KSW_E); }
                else if (s < tl) { const int2 v = bin[s]; iH = (uint32_t)v.x; iF = (uint32_t)v.y; }
            }
            if (i >= 0 && i < tl && lane < nl) {
                const uint32_t tb2 = (i < tla ? (uint32_t)ta[i] : 9u) | (i < tlb ? (uint32_t)tb_[i] : 9u) << 16; // 9: past the end, never equal
                uint32_t hd = hdiag, F = iF;
                hdiag = iH;
                uint32_t rowm = 0; // biased scores are >= 1: 0 is below all of them
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const uint32_t mm = __vminu2(qq[c] ^ tb2, ONE2);           // 0 = equal, 1 = different, per half
                    uint32_t z = hd + ONE2 - 3u * mm;                          // + (1 - 3 mm) per half
                    const uint32_t e = Ea[c];
                    z = __vmaxs2(__vmaxs2(z, e), F);
                    const uint32_t t1 = z - Q2;
                    Ea[c] = __vmaxs2(e, t1) - E2;
                    F = __vmaxs2(F, t1) - E2;
                    hd = Hp[c];
                    Hp[c] = z;
                    rowm = __vmaxs2(rowm, z);
                }
                // Row maximum of this lane's columns.  Columns past a query's end hold dead values, but a dead value is at least 2
                // below a live one this lane has already seen (the cell diagonally above-left of the first dead column, or a
                // live cell further left in the same row), so it never passes the test against the lane's best below; the column
                // -- the first one on ties (smaller j = earlier anti-diagonal), hence a live one -- is only looked up then.
                const int zrA = (int)(rowm & 0xffffu) - KSW2_BIAS, zrB = (int)(rowm >> 16) - KSW2_BIAS;
                const bool ha = i < tla && zrA > 0 && zrA >= bzA, hb = i < tlb && zrB > 0 && zrB >= bzB;
                if (ha || hb) {
                    int ca = 0, cb = 0;
#pragma unroll
                    for (int c = C - 1; c >= 0; --c) { const uint32_t x = Hp[c] ^ rowm; if ((x & 0xffffu) == 0) ca = c; if ((x >> 16) == 0) cb = c; }
                    if (ha && ksw_ext_better(zrA, i, j0 + ca, bzA, biA, bjA, qlra, tlra)) { bzA = zrA; biA = i; bjA = j0 + ca; }
                    if (hb && ksw_ext_better(zrB, i, j0 + cb, bzB, biB, bjB, qlrb, tlrb)) { bzB = zrB; biB = i; bjB = j0 + cb; }
                }
                oH = Hp[C - 1]; oF = F;
                if (lane == nl - 1 && b + 1 < nblk) bout[i] = make_int2((int)oH, (int)oF);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        int oz = __shfl_xor_sync(TH_FULL, bzA, d), oi = __shfl_xor_sync(TH_FULL, biA, d), oj = __shfl_xor_sync(TH_FULL, bjA, d);
        if (oi >= 0 && (biA < 0 || ksw_ext_better(oz, oi, oj, bzA, biA, bjA, qlra, tlra))) { bzA = oz; biA = oi; bjA = oj; }
        oz = __shfl_xor_sync(TH_FULL, bzB, d); oi = __shfl_xor_sync(TH_FULL, biB, d); oj = __shfl_xor_sync(TH_FULL, bjB, d);
        if (oi >= 0 && (biB < 0 || ksw_ext_better(oz, oi, oj, bzB, biB, bjB, qlrb, tlrb))) { bzB = oz; biB = oi; bjB = oj; }
    }
    mqA = bjA; mtA = biA; mqB = bjB; mtB = biB;
}
