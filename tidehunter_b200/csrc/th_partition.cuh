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
__global__ void __launch_bounds__(PART_WARPS * 32)
partition_kernel(DevParams P, int n_reads, const int64_t *__restrict__ roff, const int32_t *__restrict__ rlen,
                 const uint8_t *__restrict__ bseq, const int32_t *__restrict__ hend, const int32_t *__restrict__ hper,
                 const int32_t *__restrict__ cells, const int32_t *__restrict__ pch_n, const int32_t *__restrict__ pch_off,
                 const int32_t *__restrict__ pch_len, int32_t *__restrict__ par, int32_t *__restrict__ par_off, int32_t *__restrict__ par_n,
                 int4 *bnd_all, int64_t bnd_stride, int *read_counter, int32_t *__restrict__ read_status,
                 unsigned long long *__restrict__ stat_cells) {
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
    if (lane == 0 && ncell) atomicAdd(stat_cells, ncell);
}

// item = {kind, task, a, b}: kind 0: global identity of unit (qoff=a,len=b) vs consensus;
// kind 1: left extension (target = reversed read prefix of length a); kind 2: right extension (target from a, length b)
struct KswItem { int32_t kind, task, a, b; int64_t seq_off; };

#define KSW_WARPS 4
__global__ void __launch_bounds__(KSW_WARPS * 32)
ksw_items_kernel(int n_items, const KswItem *__restrict__ items, const uint8_t *__restrict__ bseq,
                 const uint8_t *__restrict__ cons_base, const int32_t *__restrict__ cons_off, const int32_t *__restrict__ cons_len,
                 uint8_t *rev_all, int64_t rev_stride, int4 *bnd_all, int64_t bnd_stride, int *counter,
                 int32_t *__restrict__ out_iden /* per item */, int32_t *__restrict__ out_ext /* 2 per item */,
                 unsigned long long *__restrict__ stat_cells) {
    const int lane = lane_id();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int4 *bnd = bnd_all + (int64_t)gw * bnd_stride;
    uint8_t *rev = rev_all + (int64_t)gw * rev_stride; // reversed copies for the left extension
    unsigned long long ncell = 0;
    while (true) {
        int it = 0;
        if (lane == 0) it = atomicAdd(counter, 1);
        it = __shfl_sync(TH_FULL, it, 0);
        if (it >= n_items) break;
        const KswItem I = items[it];
        const int cl = cons_len[I.task];
        const uint8_t *cons = cons_base + cons_off[I.task];
        int o0 = 0, o1 = 0;
        if (cl <= 0) { if (lane == 0) { out_iden[it] = 0; out_ext[2 * it] = -1; out_ext[2 * it + 1] = -1; } continue; }
        if (I.kind == 0) { // ksw2_global(query = unit, target = consensus), src/gen_cons.c:211
            ksw_warp<KSW_GLOBAL, 16>(bseq + I.seq_off + I.a, I.b, cons, cl, 0, bnd, o0, o1);
            ncell += (unsigned long long)I.b * cl;
            if (lane == 0) out_iden[it] = o0;
        } else if (I.kind == 1) { // ksw2_left_ext: both sequences reversed (src/ksw2_align.c:161-173)
            const int tl = I.a; const uint8_t *rs = bseq + I.seq_off;
            uint8_t *rq = rev, *rt = rev + ((cl + 15) & ~15);
            for (int i = lane; i < cl; i += 32) rq[i] = cons[cl - 1 - i];
            for (int i = lane; i < tl; i += 32) rt[i] = rs[tl - 1 - i];
            __syncwarp();
            ksw_warp<KSW_EXT, 16>(rq, cl, rt, tl, 0, bnd, o0, o1);
            ncell += (unsigned long long)cl * tl;
            if (lane == 0) { out_ext[2 * it] = o0; out_ext[2 * it + 1] = o1; }
            __syncwarp();
        } else {
            ksw_warp<KSW_EXT, 16>(cons, cl, bseq + I.seq_off + I.a, I.b, 0, bnd, o0, o1);
            ncell += (unsigned long long)cl * I.b;
            if (lane == 0) { out_ext[2 * it] = o0; out_ext[2 * it + 1] = o1; }
        }
    }
    if (lane == 0 && ncell) atomicAdd(stat_cells, ncell);
}
