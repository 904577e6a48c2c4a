// th_partition.cuh -- unit partitioning along a chain, and the post-consensus alignments.
//
// partition_kernel replaces get_partition_pos_with_narrow_global_alignment (src/partition.c:171-276):
// one warp per read walks each of the read's chains; inter-anchor windows are aligned with the
// warp-systolic ksw routine in GLOBAL_STOP mode, which returns iden_n and the projected boundary
// (ksw2_global_with_cigar + ksw2_backtrack_left_end) without a traceback matrix.
// The left pass of the reference (:186-228) never iterates because est_ch_i is always 0
// (src/tandem_chain.c:251-255), so par_pos = [est_start, est_start+est_period, right pass...].
//
// ksw_items_kernel runs the post-consensus alignments of seqs_msa (src/gen_cons.c:208-223): identity
// of every unit against the consensus (ksw2_global) and the two boundary extensions
// (ksw2_left_ext / ksw2_right_ext), one warp per item, persistent warps on an atomic work counter.
#pragma once
#include "th_common.cuh"
#include "th_ksw.cuh"

#define PART_WARPS 4
#ifndef PART_MIN_BLOCKS
#define PART_MIN_BLOCKS 4   // resident blocks per SM the grid is sized for (tuning knob)
#endif
__global__ void __launch_bounds__(PART_WARPS * 32, PART_MIN_BLOCKS)
partition_kernel(DevParams P, int n_reads, const int64_t *__restrict__ roff, const int32_t *__restrict__ rlen,
                 const uint8_t *__restrict__ bseq, const int32_t *__restrict__ hend, const int32_t *__restrict__ hper,
                 const int32_t *__restrict__ cells, const int32_t *__restrict__ pch_n, const int32_t *__restrict__ pch_off,
                 const int32_t *__restrict__ pch_len, int32_t *__restrict__ par, int32_t *__restrict__ par_off, int32_t *__restrict__ par_n,
                 int4 *bnd_all, int64_t bnd_stride, int *read_counter, int32_t *__restrict__ read_status,
                 unsigned long long *__restrict__ stat_cells, const int32_t *__restrict__ order) {
    const int lane = lane_id();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int4 *bnd = bnd_all + (int64_t)gw * bnd_stride;
    unsigned long long ncell = 0;
    const int k = P.k;
    while (true) {
        int r = 0;
        if (lane == 0) r = atomicAdd(read_counter, 1);
        r = __shfl_sync(TH_FULL, r, 0);
        if (r >= n_reads) break;
        r = order[r];                                  // reads with the most hits first (the chaining DP's queue order)
        const int nch = pch_n[r];
        if (nch == 0) continue;
        const int64_t off = roff[r], hoff = off / 2; const int L = rlen[r];
        const int32_t *en = hend + off, *pr = hper + off, *cl = cells + off;
        const uint8_t *bs = bseq + off;
        int32_t *out = par + 2 * off; const int cap = 2 * L;
        int used = 0;
        for (int c = 0; c < nch; ++c) {
            const int32_t *cc = cl + pch_off[hoff + c]; const int len = pch_len[hoff + c];
            const int est_start = en[cc[0]] - pr[cc[0]], est_period = pr[cc[0]];
            const int last_start = en[cc[len - 1]] - pr[cc[len - 1]];
            const int p0 = used; int pn_ = 0; bool overflow = false;
#define PUSH(v) do { if (used < cap) { if (lane == 0) out[used] = (v); ++used; ++pn_; } else overflow = true; } while (0)
            PUSH(est_start); PUSH(est_start + est_period);
            int ch_i = 0, s = est_start, e = est_start + est_period, guard = 0;
            while (ch_i < len - 1 && e <= last_start && !overflow) {
                if (++guard > 4 * L + 64) { overflow = true; break; } // the reference would spin; never seen, but never hang the GPU
                int s1 = s, e1 = e, i; bool brk = false;
                for (i = ch_i + 1; i < len; ++i) {
                    const int e2 = en[cc[i]], s2 = e2 - pr[cc[i]];
                    if (s2 == e) { PUSH(e2); ch_i = i; s = s2; e = e2; brk = true; break; }
                    else if (s2 > e) {
                        const int ql = s2 - s1 + k, tl = e2 - e1 + k;
                        int iden = 0, tle = 0;
                        if (ql >= 65536 || tl >= 65536) { if (lane == 0) read_status[r] = TH_ERR_LEN; overflow = true; brk = true; break; }
                        ksw_warp<KSW_GLOBAL_STOP, 4>(bs + s1 - k + 1, ql, bs + e1 - k + 1, tl, ql - (s2 - e), bnd, iden, tle);
                        ncell += (unsigned long long)ql * tl;
                        if ((double)iden >= (double)min(ql, tl) * (1 - P.max_div)) {
                            s = e; e = e2 - tle;
                            if (e == s) { ch_i = len; brk = true; break; }
                            PUSH(e); ch_i = i - 1;
                        } else { PUSH(-1); PUSH(s2); PUSH(e2); ch_i = i; s = s2; e = e2; }
                        brk = true; break;
                    } else { s1 = s2; e1 = e2; }
                }
                if (!brk) break;
            }
#undef PUSH
            if (overflow && lane == 0 && read_status[r] == TH_OK) read_status[r] = TH_ERR_CAP;
            if (lane == 0) { par_off[hoff + c] = p0; par_n[hoff + c] = overflow ? 0 : pn_; }
        }
    }
    if (lane == 0 && ncell) { atomicAdd(stat_cells, ncell); atomicAdd(stat_cells + 1, ncell); }
}

// item kinds: 0: global identity of unit (offset a, length b) vs consensus -> out_iden[out];
// 3: two units (a, b) and (a2, b2) against the same consensus, packed 16x2 -> out_iden[out], out_iden[out + 1];
// 4: two units of different tasks/reads, unit 2 = (seq_off2 + a2, b2) vs the consensus of task2 -> out_iden[out], out_iden[out2];
// 1: left extension (target = reversed read prefix of length a) -> out_ext[out .. out+1];
// 2: right extension (target from a, length b) -> out_ext[out .. out+1]
struct KswItem { int32_t kind, task, a, b, a2, b2, out, task2, out2, pad; int64_t seq_off, seq_off2; };

__device__ __forceinline__ bool warp_has_n(const uint8_t *s, int l) {
    const int lane = lane_id();
    bool f = false;
    for (int i = lane; i < l; i += 32) f |= s[i] >= 4;
    return __any_sync(TH_FULL, f);
}

#define KSW_WARPS 4
#ifndef KSW_PAIR_MIN_BLOCKS
#define KSW_PAIR_MIN_BLOCKS 4
#endif
#ifndef KSW2_C
#define KSW2_C 8     // columns per lane of the packed identity alignments: 256-column blocks follow the band closely (16: 42.5 ms, 8: 38.3 ms per 8,192 R2C2 reads for the ksw stage)
#endif
#define KSW_MIN_BLOCKS 4
#ifndef KSW_EXT_MIN_BLOCKS
#define KSW_EXT_MIN_BLOCKS 4
#endif

// Post-consensus alignments of seqs_msa (src/gen_cons.c:208-223) run as three kernels with their own register
// budgets, persistent warps on atomic work counters:
//   ksw_pair_kernel   kinds 3/4: two unit-vs-consensus global alignments per warp (packed 16x2); pairs that do not
//                     qualify (N bases, > KSW2_MAXLEN) are appended to a redo list
//   ksw_single_kernel kind 0 and the redo list: one global alignment per warp, 32-bit
//   ksw_ext_kernel    kinds 1/2: boundary extensions (score-only, arg-max in the reference's visiting order)
__global__ void __launch_bounds__(KSW_WARPS * 32, KSW_PAIR_MIN_BLOCKS)
ksw_pair_kernel(int n_items, const KswItem *__restrict__ items, const uint8_t *__restrict__ bseq,
                const uint8_t *__restrict__ cons_base, const int32_t *__restrict__ cons_off, const int32_t *__restrict__ cons_len,
                int4 *bnd_all, int64_t bnd_stride, int *counter, int32_t *__restrict__ redo_list, int *redo_count,
                int32_t *__restrict__ out_iden, unsigned long long *__restrict__ stat_cells, float *band_alpha) {
    const int lane = lane_id();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int4 *bnd = bnd_all + (int64_t)gw * bnd_stride;
    unsigned long long ncell = 0, nfull = 0; // cells computed / cells of the full matrices (what the reference computes)
    // Band half-width as a fraction of the alignment's length: every warp adapts its own from pair to pair
    // (ksw_pair_identity); it starts from a running mean of what the warps of the context's earlier launches ended with (plain
    // loads and stores of one float in device memory: lost updates do not matter, and a stale or odd value costs a few rows
    // or one retry, never a result), so a launch with few pairs per warp already knows what the data needs.
    float alpha = 0.f;
    if (lane == 0) alpha = *(volatile float *)band_alpha;
    alpha = fminf(fmaxf(__shfl_sync(TH_FULL, alpha, 0), 0.02f), 0.5f);
    bool any_pair = false;
    while (true) {
        int it = 0;
        if (lane == 0) it = atomicAdd(counter, 1);
        it = __shfl_sync(TH_FULL, it, 0);
        if (it >= n_items) break;
        const KswItem I = items[it];
        const int t2 = I.kind == 4 ? I.task2 : I.task;
        const int cl = cons_len[I.task], cl2 = cons_len[t2];
        const uint8_t *cons = cons_base + cons_off[I.task], *cons2 = cons_base + cons_off[t2];
        const uint8_t *qa = bseq + I.seq_off + I.a, *qb = bseq + (I.kind == 4 ? I.seq_off2 : I.seq_off) + I.a2;
        const int out2 = I.kind == 4 ? I.out2 : I.out + 1;
        if (cl <= 0 && cl2 <= 0) { if (lane == 0) { out_iden[I.out] = 0; out_iden[out2] = 0; } continue; }
        bool ok = cl > 0 && cl2 > 0 && max(max(I.b, I.b2), max(cl, cl2)) <= KSW2_MAXLEN;
        ok = ok && !warp_has_n(qa, I.b) && !warp_has_n(qb, I.b2) && !warp_has_n(cons, cl) && (cons2 == cons || !warp_has_n(cons2, cl2));
        if (!ok) { if (lane == 0) redo_list[atomicAdd(redo_count, 1)] = it; continue; }
        int o0 = 0, o1 = 0;
        ksw_pair_identity<KSW2_C>(qa, I.b, cons, cl, qb, I.b2, cons2, cl2, bnd, alpha, o0, o1, ncell);
        any_pair = true;
        nfull += (unsigned long long)I.b * cl + (unsigned long long)I.b2 * cl2;
        if (lane == 0) { out_iden[I.out] = o0; out_iden[out2] = o1; }
    }
    if (lane == 0 && ncell) { atomicAdd(stat_cells, ncell); atomicAdd(stat_cells + 1, nfull); }
    if (lane == 0 && any_pair) { volatile float *g = band_alpha; *g = 0.75f * *g + 0.25f * alpha; } // damped: one odd pair does not steer the next launch
}

__global__ void __launch_bounds__(KSW_WARPS * 32, KSW_MIN_BLOCKS)
ksw_single_kernel(int n_single, const KswItem *__restrict__ singles, const KswItem *__restrict__ pairs,
                  const int32_t *__restrict__ redo_list, const int *__restrict__ redo_count, const uint8_t *__restrict__ bseq,
                  const uint8_t *__restrict__ cons_base, const int32_t *__restrict__ cons_off, const int32_t *__restrict__ cons_len,
                  int4 *bnd_all, int64_t bnd_stride, int *counter, int32_t *__restrict__ out_iden, unsigned long long *__restrict__ stat_cells) {
    const int lane = lane_id();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int4 *bnd = bnd_all + (int64_t)gw * bnd_stride;
    unsigned long long ncell = 0;
    const int n_work = n_single + 2 * *redo_count; // each redo pair is two alignments
    while (true) {
        int it = 0;
        if (lane == 0) it = atomicAdd(counter, 1);
        it = __shfl_sync(TH_FULL, it, 0);
        if (it >= n_work) break;
        int task, a, b, out; int64_t so;
        if (it < n_single) { const KswItem I = singles[it]; task = I.task; a = I.a; b = I.b; out = I.out; so = I.seq_off; }
        else {
            const KswItem I = pairs[redo_list[(it - n_single) >> 1]];
            if (((it - n_single) & 1) == 0) { task = I.task; a = I.a; b = I.b; out = I.out; so = I.seq_off; }
            else if (I.kind == 4) { task = I.task2; a = I.a2; b = I.b2; out = I.out2; so = I.seq_off2; }
            else { task = I.task; a = I.a2; b = I.b2; out = I.out + 1; so = I.seq_off; }
        }
        const int cl = cons_len[task];
        int o0 = 0, o1 = 0;
        if (cl > 0) { // ksw2_global(query = unit, target = consensus), src/gen_cons.c:211
            ksw_warp<KSW_GLOBAL, 16>(bseq + so + a, b, cons_base + cons_off[task], cl, 0, bnd, o0, o1);
            ncell += (unsigned long long)b * cl;
        }
        if (lane == 0) out_iden[out] = o0;
    }
    if (lane == 0 && ncell) { atomicAdd(stat_cells, ncell); atomicAdd(stat_cells + 1, ncell); }
}

// Boundary extensions, two per warp (packed 16x2): `order` lists the items by target length, so entries 2k and 2k + 1 are
// of similar size.  A pair with N in a sequence or one longer than KSW2_MAXLEN runs one after the other, 32-bit.
__device__ __forceinline__ void ksw_ext_prepare(const KswItem &I, const uint8_t *bseq, const uint8_t *cons, int cl, uint8_t *rev,
                                                const uint8_t *&q, const uint8_t *&t, int &tl) {
    const int lane = lane_id();
    if (I.kind == 1) { // ksw2_left_ext: both sequences reversed (src/ksw2_align.c:161-173)
        tl = I.a; const uint8_t *rs = bseq + I.seq_off;
        const int cle = min(cl, 2 * max(tl, 0)), tle = min(tl, 2 * cl); // the part of each the extension looks at (ksw_warp)
        uint8_t *rq = rev, *rt = rev + ((cle + 15) & ~15);
        for (int i = lane; i < cle; i += 32) rq[i] = cons[cl - 1 - i];
        for (int i = lane; i < tle; i += 32) rt[i] = rs[tl - 1 - i];
        q = rq; t = rt;
    } else { q = cons; t = bseq + I.seq_off + I.a; tl = I.b; } // ksw2_right_ext (src/ksw2_align.c:153-159)
}
__global__ void __launch_bounds__(KSW_WARPS * 32, KSW_EXT_MIN_BLOCKS)
ksw_ext_kernel(int n_items, const KswItem *__restrict__ items, const int32_t *__restrict__ order, const uint8_t *__restrict__ bseq,
               const uint8_t *__restrict__ cons_base, const int32_t *__restrict__ cons_off, const int32_t *__restrict__ cons_len,
               uint8_t *rev_all, int64_t rev_stride, int4 *bnd_all, int64_t bnd_stride, int *counter,
               int32_t *__restrict__ out_ext, unsigned long long *__restrict__ stat_cells) {
    const int lane = lane_id();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int4 *bnd = bnd_all + (int64_t)gw * bnd_stride;
    uint8_t *rev = rev_all + (int64_t)gw * rev_stride; // reversed copies for left extensions: one half of it per item of the pair
    unsigned long long ncell = 0, nfull = 0;
    const int n_pairs = (n_items + 1) >> 1;
    while (true) {
        int it = 0;
        if (lane == 0) it = atomicAdd(counter, 1);
        it = __shfl_sync(TH_FULL, it, 0);
        if (it >= n_pairs) break;
        const bool hasb = 2 * it + 1 < n_items;
        const KswItem A = items[order[2 * it]], B = items[order[hasb ? 2 * it + 1 : 2 * it]];
        const int cla = cons_len[A.task], clb = hasb ? cons_len[B.task] : 0;
        const uint8_t *qa = nullptr, *ta = nullptr, *qb = nullptr, *tb = nullptr; int tla = 0, tlb = 0;
        if (cla > 0) ksw_ext_prepare(A, bseq, cons_base + cons_off[A.task], cla, rev, qa, ta, tla);
        if (clb > 0) ksw_ext_prepare(B, bseq, cons_base + cons_off[B.task], clb, rev + rev_stride / 2, qb, tb, tlb);
        __syncwarp();
        int aq = -1, at = -1, bq = -1, bt = -1;
        // effective sizes: rows from 2 cl on and columns from 2 tl on are never computed (ksw_warp)
        const int cae = min(cla, 2 * max(tla, 0)), tae = min(max(tla, 0), 2 * cla), cbe = min(clb, 2 * max(tlb, 0)), tbe = min(max(tlb, 0), 2 * clb);
        bool two = cae > 0 && cbe > 0 && tae > 0 && tbe > 0 && max(max(cae, cbe), max(tae, tbe)) <= KSW2_MAXLEN;
        if (two) two = !warp_has_n(qa, cae) && !warp_has_n(ta, tae) && !warp_has_n(qb, cbe) && !warp_has_n(tb, tbe);
        if (two) ksw_warp_ext2<16>(qa, cla, ta, tla, qb, clb, tb, tlb, bnd, aq, at, bq, bt);
        else {
            if (cla > 0) ksw_warp<KSW_EXT, 16>(qa, cla, ta, tla, 0, bnd, aq, at);
            if (clb > 0) ksw_warp<KSW_EXT, 16>(qb, clb, tb, tlb, 0, bnd, bq, bt);
        }
        ncell += (unsigned long long)cae * tae + (unsigned long long)cbe * tbe; // cells computed
        nfull += (unsigned long long)max(cla, 0) * max(tla, 0) + (unsigned long long)max(clb, 0) * max(tlb, 0);
        __syncwarp();
        if (lane == 0) { out_ext[A.out] = aq; out_ext[A.out + 1] = at; if (hasb) { out_ext[B.out] = bq; out_ext[B.out + 1] = bt; } }
    }
    if (lane == 0 && ncell) { atomicAdd(stat_cells, ncell); atomicAdd(stat_cells + 1, nfull); }
}
