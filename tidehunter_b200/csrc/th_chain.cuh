// th_chain.cuh -- period-consistent chaining DP and chain extraction.
//
// Replaces src/tandem_chain.c:113-166 (init_dp, get_con_score), :290-356 (main DP) and
// :21-111, :170-223, :251-255, :358-403 (ranking, greedy extraction, ordering) of the reference.
//
// chain_dp_kernel: one warp per read.  Cells are the hits in (end, period) order; with one hit per
// end (always true for -w 1) the reference's scan "rows downwards, stop rules in order" is a linear
// scan over earlier hits, evaluated 32 predecessors at a time: a shuffle prefix-max reproduces the
// running max_score each predecessor would have seen, and ballots pick the first lane that triggers
// one of the ordered stop rules.  Reads with repeated ends (possible only with minimizer seeds) go to
// chain_dp_generic_kernel, a literal one-thread restatement.
#pragma once
#include "th_common.cuh"

enum { CON_NO = 0, CON_REG = 1, CON_SAME = 2, CON_OVL = 3 };

// get_con_score (tandem_chain.c:151-166).  The reference's double test `cur_p >= pre_p * 1.8` equals
// 5*cur_p >= 9*pre_p for all int periods: when 9*pre_p/5 is an integer N the correctly rounded product
// fl(pre_p * 1.8) is exactly N (1.8's representation error is 2.5e-17 relative, below half an ulp),
// otherwise the exact product is at least 0.2 away from any integer.
// SMALL: all periods < 2^27 (max_p bounds them), so the products fit 32 bits.
template <bool SMALL = false>
__device__ __forceinline__ int con_score(int cs, int ce, int ps, int pe, int k, int &score) {
    int cp = ce - cs, pp = pe - ps;
    if (SMALL) { if (cs <= ps || 5 * cp >= 9 * pp || 5 * pp >= 9 * cp) return CON_NO; }
    else if (cs <= ps || 5ll * cp >= 9ll * pp || 5ll * pp >= 9ll * cp) return CON_NO;
    int de = abs(ce - pe), ds = abs(cs - ps), dpd = abs(cp - pp);
    int matched = min(de, k) + min(ds, k);
    uint32_t v = (uint32_t)(de + ds);
    int lg = v ? 31 - __clz(v) : -1;
    score = matched - (dpd * dpd / 2 + lg / 2);
    if (dpd == 0) return matched < 2 * k ? CON_OVL : CON_SAME;
    return CON_REG;
}

#define CHAIN_WARPS 4
#ifndef CHAIN_BLOCKS_PER_SM
#define CHAIN_BLOCKS_PER_SM 8   // persistent grid: blocks per SM (tuning knob)
#endif
#ifndef CHAIN_BUCKETS
#define CHAIN_BUCKETS 1         // 0: always the plain window scan
#endif
#define CHAIN_BK_S 512          // period buckets [S j, S j + 2 S), stride S: every hit is in two of them
#define CHAIN_BK_MAX 64         // most buckets kept in shared memory (max_p <= 31,744); more: plain scan

// Period buckets (round 2).  The reference scans, for every cell, ALL earlier cells that end within one period -- ~300 of
// them on 10 kb reads, most of them chance matches or harmonics with other periods.  A predecessor q can only raise the
// running maximum when dpd^2 < 2 (gmax + 2k), dpd = |period(cur) - period(q)|, gmax = the largest score of the read so far:
// otherwise con_score's penalty floor(dpd^2 / 2) alone exceeds every score, and the stop rules that need no improvement
// (overlap, dpd == 0) cannot fire either.  Skipping such a predecessor changes nothing the reference computes.  So the hits
// are kept a second time sorted into period buckets (stable: index order inside a bucket), with the scores written through,
// and a cell scans only the one bucket that contains [period - D, period + D] (D = the bound above, <= S / 2), newest entry
// first, with the same ordered stop rules as the plain scan.  The number of predecessors the reference evaluates (the work
// counter) follows from indices: one hit per end position, so rows = index differences.  The plain scan remains for cells
// whose window is as long as their period (the reference's "max_h rows without improvement" rule could fire; it cannot
// otherwise), for D > S / 2 (scores beyond 32 k), more than CHAIN_BK_MAX buckets, or reads with more hits than half their
// length (the bucket arrays reuse chunk-wide buffers that are free at this point).  Model and check against the oracle:
// tools/sim/chain_bucket_sim.py.
template <bool SMALL>
__global__ void __launch_bounds__(CHAIN_WARPS * 32)
chain_dp_kernel(DevParams P, int n_reads, const int64_t *__restrict__ roff, const int32_t *__restrict__ nhits,
                const int32_t *__restrict__ hend, const int32_t *__restrict__ hper,
                int32_t *score, int32_t *from, int32_t *__restrict__ generic_flag,
                unsigned long long *__restrict__ eval_count, const int32_t *__restrict__ order, int *__restrict__ next_read,
                int32_t *bk_idx, int32_t *bk_en, int32_t *bk_pr, int32_t *bk_sc) {
    __shared__ int s_off_all[CHAIN_WARPS][CHAIN_BK_MAX + 1], s_fill_all[CHAIN_WARPS][CHAIN_BK_MAX];
    const int lane = lane_id();
    int *const s_off = s_off_all[threadIdx.x >> 5], *const s_fill = s_fill_all[threadIdx.x >> 5];
    const int nbk = (int)min((unsigned)(P.max_p / CHAIN_BK_S) + 2u, 1u << 20);
    const bool can_bucket = CHAIN_BUCKETS && bk_idx != nullptr && nbk <= CHAIN_BK_MAX;
    const unsigned below = (1u << lane) - 1u;
    unsigned long long evals = 0;
    // a read is one warp from start to end, and its time grows with hits x hits per period: warps take the reads from a
    // queue ordered by hit count, most hits first, so that the longest reads do not start last
    while (true) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(next_read, 1);
        qi = __shfl_sync(TH_FULL, qi, 0);
        if (qi >= n_reads) break;
        const int r = order[qi];
        const int n = nhits[r];
        const int64_t off = roff[r];
        const int32_t *en = hend + off, *pr = hper + off;
        int32_t *sc = score + off, *fr = from + off;
        if (n < 2) { if (lane == 0) generic_flag[r] = 0; continue; }
        // repeated ends? (rows with more than one cell) -> generic kernel
        int dup = 0;
        for (int i = lane + 1; i < n; i += 32) dup |= (en[i] == en[i - 1]);
        dup = __any_sync(TH_FULL, dup);
        if (lane == 0) generic_flag[r] = dup;
        if (dup) continue;
        for (int i = lane; i < n; i += 32) { sc[i] = P.k + min(P.k, pr[i]); fr[i] = -1; } // init_dp
        const bool bucketed = can_bucket && 2ll * n <= roff[r + 1] - off;
        int32_t *const bi = bk_idx + off, *const be = bk_en + off, *const bp = bk_pr + off, *const bs = bk_sc + off;
        if (bucketed) { // stable counting sort of the entries e = 2 i + t (hit i, bucket period / S - 1 + t) by bucket
            for (int b = lane; b < nbk; b += 32) s_fill[b] = 0;
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                const int b = min(pr[i] / CHAIN_BK_S, nbk - 1);
                atomicAdd(&s_fill[b], 1);
                if (b >= 1) atomicAdd(&s_fill[b - 1], 1);
            }
            __syncwarp();
            if (lane == 0) { int acc = 0; for (int b = 0; b < nbk; ++b) { const int c = s_fill[b]; s_off[b] = acc; s_fill[b] = acc; acc += c; } s_off[nbk] = acc; }
            __syncwarp();
            for (int base = 0; base < 2 * n; base += 32) {
                const int e = base + lane, i = e >> 1;
                int key = -1 - lane, pv = 0, ev = 0;   // entries without a bucket get keys of their own
                if (e < 2 * n) { pv = pr[i]; ev = en[i]; const int b = min(pv / CHAIN_BK_S, nbk - 1) - 1 + (e & 1); if (b >= 0) key = b; }
                const unsigned m = __match_any_sync(TH_FULL, key);
                int pos = 0;
                if (key >= 0) pos = s_fill[key] + __popc(m & below);
                __syncwarp();
                if (key >= 0 && (m & below) == 0) s_fill[key] += __popc(m);
                if (key >= 0) { bi[pos] = i; be[pos] = ev; bp[pos] = pv; bs[pos] = P.k + min(P.k, pv); }
                __syncwarp();
            }
            for (int b = lane; b < nbk; b += 32) s_fill[b] = 0;   // from here on: entries of the bucket with index < cur
            __syncwarp();
            if (lane == 0) { const int b = min(pr[0] / CHAIN_BK_S, nbk - 1); s_fill[b] = 1; if (b >= 1) s_fill[b - 1] = 1; }
        }
        __syncwarp();
        int gmax = 2 * P.k;   // >= every score of the read so far
        for (int cur = 1; cur < n; ++cur) {
            const int ce = en[cur], cp = pr[cur], cs = ce - cp;
            const int init = P.k + min(P.k, cp);
            int max_score = init, best_pre = -1, iter_in = 0;
            const int max_h = cp;
            bool fast = false;
            int wlo = 0;
            const int T2 = 2 * (min(gmax, 1 << 27) + 2 * P.k);   // (scores that large send the cell to the plain scan through Dc)
            const int Dc = (int)sqrtf((float)T2) + 1;     // >= the largest dpd with dpd^2 < T2
            if (bucketed && Dc <= CHAIN_BK_S / 2) {
                // first index whose end is >= cs (the window's far end): 32-ary search over the ends before cur
                int lo = 0, hi = cur;
                while (hi - lo > 32) {
                    const int step = (hi - lo + 31) >> 5, probe = lo + lane * step;
                    const int v = probe < hi ? en[probe] : INT_MAX;
                    const unsigned m = __ballot_sync(TH_FULL, v >= cs);
                    if (m == 0) lo = lo + 31 * step + 1;    // every probed end is before cs
                    else { const int f = __ffs(m) - 1; hi = min(hi, lo + f * step); if (f) lo = lo + (f - 1) * step + 1; }
                }
                { const int probe = lo + lane; const int v = probe < hi ? en[probe] : INT_MAX;
                  const unsigned m = __ballot_sync(TH_FULL, v >= cs);
                  wlo = m ? min(hi, lo + __ffs(m) - 1) : hi; }
                fast = cur - wlo < max_h;
            }
            if (fast) {
                const int j = max(cp - Dc, 0) / CHAIN_BK_S;
                const int seg_lo = s_off[j];
                int stop_idx = -1;
                for (int base = seg_lo + s_fill[j] - 1; base >= seg_lo; base -= 32) {
                    const int p = base - lane;
                    const bool valid = p >= seg_lo;
                    int pe = 0, pp = 1, psc = 0, pq = 0;
                    if (valid) { pe = be[p]; pp = bp[p]; psc = bs[p]; pq = bi[p]; }
                    const bool cstop = !valid || pe < cs;           // stop BEFORE this predecessor
                    const int dpd = abs(cp - pp);                   // <= 2 S inside a bucket: the square fits
                    int con = 0, cls = CON_NO;
                    if (!cstop && dpd * dpd < T2) cls = con_score<SMALL>(cs, ce, pe - pp, pe, P.k, con);
                    const int s = cls != CON_NO ? psc + con : INT_MIN;
                    const unsigned cm = __ballot_sync(TH_FULL, cstop);
                    const int first_c = cm ? __ffs(cm) - 1 : 32;
                    int first_a;
                    if (__reduce_max_sync(TH_FULL, s) <= max_score) {
                        const unsigned am = __ballot_sync(TH_FULL, !cstop && cls == CON_OVL);
                        first_a = am ? __ffs(am) - 1 : 32;
                    } else {
                        int inc = s;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) inc = max(inc, __shfl_up_sync(TH_FULL, inc, d)); // lanes < d get their own value
                        int excl = __shfl_up_sync(TH_FULL, inc, 1);
                        excl = lane == 0 ? max_score : max(max_score, excl);
                        const bool imp = cls != CON_NO && s > excl;
                        const bool stop_after = (imp && (cls == CON_SAME || cls == CON_OVL)) || (!imp && cls == CON_OVL);
                        const unsigned am = __ballot_sync(TH_FULL, stop_after && !cstop);
                        first_a = am ? __ffs(am) - 1 : 32;
                        const bool processed = lane < first_c && lane <= first_a;
                        const int bm = __reduce_max_sync(TH_FULL, processed ? s : INT_MIN);
                        if (bm > max_score) {
                            const unsigned wm = __ballot_sync(TH_FULL, processed && imp && s == bm);
                            max_score = bm; best_pre = __shfl_sync(TH_FULL, pq, __ffs(wm) - 1);
                        }
                    }
                    if (first_a < first_c) stop_idx = __shfl_sync(TH_FULL, pq, first_a);   // a stop rule fired at that predecessor
                    if (first_c < 32 || first_a < 32) break;
                }
                evals += (unsigned long long)(cur - max(wlo, stop_idx));   // rows the reference walks: down to the stop or the window's end
            } else {
            // the next batch of predecessors is loaded while the current one is evaluated (its scores are final: they lie at
            // least 32 cells behind the current one); a batch is ~100 instructions, about one L2 round trip
            int n_pe = 0, n_pp = 1, n_psc = 0;
            { const int pre = cur - 1 - lane; if (pre >= 0) { n_pe = en[pre]; n_pp = pr[pre]; n_psc = sc[pre]; } }
            for (int base = cur - 1; base >= 0; base -= 32) {
                const int pre = base - lane;
                const bool valid = pre >= 0;
                const int pe = n_pe, pp = n_pp, psc = n_psc;
                { const int nx = pre - 32; n_pe = 0; n_pp = 1; n_psc = 0; if (nx >= 0) { n_pe = en[nx]; n_pp = pr[nx]; n_psc = sc[nx]; } }
                const bool cstop = !valid || pe < cs;           // stop BEFORE this predecessor
                int con = 0, cls = CON_NO;
                if (!cstop) cls = con_score<SMALL>(cs, ce, pe - pp, pe, P.k, con);
                const int s = cls != CON_NO ? psc + con : INT_MIN;
                const unsigned cm = __ballot_sync(TH_FULL, cstop);
                const int first_c = cm ? __ffs(cm) - 1 : 32;
                int first_a, cnt;
                if (__reduce_max_sync(TH_FULL, s) <= max_score) {
                    // no predecessor of this batch can improve (the usual case once the best one, a near one, is
                    // found): only the non-improving stop rules apply, no running maximum is needed
                    cnt = iter_in + lane + 1;
                    const unsigned am = __ballot_sync(TH_FULL, !cstop && (cls == CON_OVL || cnt >= max_h));
                    first_a = am ? __ffs(am) - 1 : 32;
                    evals += min(first_c, first_a + 1);
                } else {
                    // running max each lane would have seen = max(max_score, s of earlier lanes)
                    int inc = s;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) inc = max(inc, __shfl_up_sync(TH_FULL, inc, d)); // lanes < d get their own value
                    int excl = __shfl_up_sync(TH_FULL, inc, 1);
                    excl = lane == 0 ? max_score : max(max_score, excl);
                    const bool imp = cls != CON_NO && s > excl;
                    const unsigned impm = __ballot_sync(TH_FULL, imp);
                    // rows without improvement so far (iter_n), as of this lane
                    const unsigned blw = impm & (0xffffffffu >> (31 - lane));
                    cnt = blw ? lane - (31 - __clz(blw)) : iter_in + lane + 1;
                    const bool stop_after = (imp && (cls == CON_SAME || cls == CON_OVL)) || (!imp && cls == CON_OVL) || (!imp && cnt >= max_h);
                    const unsigned am = __ballot_sync(TH_FULL, stop_after && !cstop);
                    first_a = am ? __ffs(am) - 1 : 32;
                    const bool processed = lane < first_c && lane <= first_a;
                    evals += min(first_c, first_a + 1);
                    const int sp = processed ? s : INT_MIN;
                    const int bm = __reduce_max_sync(TH_FULL, sp);
                    if (bm > max_score) {
                        const unsigned wm = __ballot_sync(TH_FULL, processed && imp && s == bm);
                        max_score = bm; best_pre = base - (__ffs(wm) - 1);
                    }
                }
                if (first_c < 32 || first_a < 32) break;
                iter_in = __shfl_sync(TH_FULL, cnt, 31);
            }
            }
            __syncwarp();   // every lane has read the bucket fill counts (an empty bucket scan has no vote in it) before lane 0 moves them
            if (lane == 0 && max_score > init) { sc[cur] = max_score; fr[cur] = best_pre; }
            if (bucketed) { // the cell becomes a predecessor: its entries are the next ones of its two buckets; the score is written through
                gmax = max(gmax, max_score);
                if (lane == 0) {
                    const int b = min(cp / CHAIN_BK_S, nbk - 1);
                    if (max_score > init) { bs[s_off[b] + s_fill[b]] = max_score; if (b >= 1) bs[s_off[b - 1] + s_fill[b - 1]] = max_score; }
                    s_fill[b] += 1; if (b >= 1) s_fill[b - 1] += 1;
                }
            }
            __syncwarp();
        }
    }
    evals = __shfl_sync(TH_FULL, evals, 0);
    if (lane == 0 && evals) atomicAdd(eval_count, evals);
}

// Literal restatement for reads whose hits repeat an end position (ragged DP rows); thread per read.
__global__ void chain_dp_generic_kernel(DevParams P, int n_reads, const int64_t *__restrict__ roff, const int32_t *__restrict__ nhits,
                                        const int32_t *__restrict__ hend, const int32_t *__restrict__ hper,
                                        int32_t *score, int32_t *from, const int32_t *__restrict__ generic_flag,
                                        int32_t *rowbeg_scratch) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads || !generic_flag[r]) return;
    const int n = nhits[r]; const int64_t off = roff[r];
    const int32_t *en = hend + off, *pr = hper + off; int32_t *sc = score + off, *fr = from + off;
    int32_t *row_beg = rowbeg_scratch + off; // n+1 entries fit: capacity L >= n+1
    int tot = 0;
    for (int i = 0; i < n; ++i) {
        if (i == 0 || en[i] != en[i - 1]) row_beg[tot++] = i;
        sc[i] = P.k + min(P.k, pr[i]); fr[i] = -1;
    }
    row_beg[tot] = n;
    for (int ci = 1; ci < tot; ++ci)
        for (int cur = row_beg[ci]; cur < row_beg[ci + 1]; ++cur) {
            int max_score = sc[cur], max_pre = -1, max_h = pr[cur], iter_n = 0; bool stop = false;
            int ce = en[cur], cs = ce - pr[cur];
            for (int pi = ci - 1; pi >= 0 && !stop; --pi) {
                bool gt = false;
                if (en[row_beg[pi]] < cs) break;
                for (int pre = row_beg[pi]; pre < row_beg[pi + 1]; ++pre) {
                    int con, cls = con_score(cs, ce, en[pre] - pr[pre], en[pre], P.k, con);
                    if (cls == CON_NO) continue;
                    int s = sc[pre] + con;
                    if (s > max_score) { max_score = s; max_pre = pre; if (cls == CON_SAME || cls == CON_OVL) { stop = true; break; } gt = true; }
                    else if (cls == CON_OVL) { stop = true; break; }
                }
                if (stop) break;
                if (gt) iter_n = 0; else if (++iter_n >= max_h) break;
            }
            if (max_score > sc[cur]) { sc[cur] = max_score; fr[cur] = max_pre; }
        }
}

// ---------------------------------------------------------------------------------------------
// Ranking: cells with score > 0 ordered by score descending, ties by the reference's listing order
// (row descending, then period ascending; tandem_chain.c:32-43 + glibc's stable merge-sort qsort).
// Block per read.  key = score<<40 | last_cell_of_row<<20 | (0xfffff - index_in_row), sorted
// descending; requires score < 2^24 and hit_n < 2^20 (flagged otherwise).
// ---------------------------------------------------------------------------------------------
#define RANK_THREADS 256
#define RANK_SMEM_CAP 4096
__global__ void __launch_bounds__(RANK_THREADS)
rank_kernel(int n_reads, const int64_t *__restrict__ roff, const int32_t *__restrict__ nhits,
            const int32_t *__restrict__ hend, const int32_t *__restrict__ score,
            uint64_t *__restrict__ gscratch, int64_t gcap, int32_t *__restrict__ rank, int32_t *__restrict__ nrank,
            int32_t *__restrict__ read_status) {
    __shared__ uint64_t sbuf[RANK_SMEM_CAP];
    __shared__ int s_cnt, s_bad;
    for (int r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const int n = nhits[r]; const int64_t off = roff[r];
        if (n < 2) { if (threadIdx.x == 0) nrank[r] = 0; continue; }
        const int npow = next_pow2(n);
        uint64_t *buf = npow <= RANK_SMEM_CAP ? sbuf : gscratch + (int64_t)blockIdx.x * gcap;
        if (threadIdx.x == 0) { s_cnt = 0; s_bad = n >= (1 << 20); }
        __syncthreads();
        for (int i = threadIdx.x; i < npow; i += blockDim.x) {
            uint64_t v = 0;
            if (i < n && score[off + i] > 0) {
                int lo = i, hi = i; const int e = hend[off + i];
                while (lo > 0 && hend[off + lo - 1] == e) --lo;
                while (hi + 1 < n && hend[off + hi + 1] == e) ++hi;
                const int sc = score[off + i];
                if (sc >= (1 << 24)) s_bad = 1;
                v = ((uint64_t)(uint32_t)sc << 40) | ((uint64_t)(uint32_t)hi << 20) | (uint64_t)(0xfffff - (i - lo));
                atomicAdd(&s_cnt, 1);
            }
            buf[i] = v;
        }
        __syncthreads();
        const int cnt = s_cnt;
        block_bitonic_sort<true>(buf, npow);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const uint64_t v = buf[i];
            int hi = (int)((v >> 20) & 0xfffff), lo = hi; const int e = hend[off + hi];
            while (lo > 0 && hend[off + lo - 1] == e) --lo;
            rank[off + i] = lo + (0xfffff - (int)(v & 0xfffff));
        }
        if (threadIdx.x == 0) { nrank[r] = s_bad ? 0 : cnt; if (s_bad) read_status[r] = TH_ERR_CAP; }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Greedy chain extraction; one warp per read.  Restatement of tandem_chain.c:358-403 on flat cell ids.
// The walk over ranked candidates is sequential in the reference, but a candidate only changes state when it
// is NOT inside an existing chain (is_in_chain, :170-185), and most are: 32 candidates are screened at a time
// against the current chain set; the first survivor is processed exactly as the reference does (lane 0), and
// the candidates behind it are screened again because the chain set may have changed.
// Output: post-chains (>= 3 cells, ascending end).  Per-read scratch (all at the read's base offset, capacity L):
//   tracked[L] u8, ch_off/ch_len/ch_score/ch_idx [L/2+1 via half offsets], cells[L]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int row_first(const int32_t *en, int c) { while (c > 0 && en[c - 1] == en[c]) --c; return c; }

#define SELECT_WARPS 4
__global__ void __launch_bounds__(SELECT_WARPS * 32)
chain_select_kernel(int n_reads, const int64_t *__restrict__ roff, const int32_t *__restrict__ nhits,
                    const int32_t *__restrict__ hend, const int32_t *__restrict__ hper,
                    const int32_t *__restrict__ score, const int32_t *__restrict__ from,
                    const int32_t *__restrict__ rank, const int32_t *__restrict__ nrank,
                    uint8_t *tracked, int32_t *ch_off, int32_t *ch_len, int32_t *ch_score, int32_t *ch_idx,
                    int32_t *cells, int32_t *__restrict__ pch_n, int32_t *pch_off, int32_t *pch_len, const int ragged_ok) {
    const int lane = lane_id();
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_reads) return;
    const int n = nhits[r]; const int64_t off = roff[r], hoff = off / 2;
    if (n < 2) { if (lane == 0) pch_n[r] = 0; return; }
    const int32_t *en = hend + off, *pr = hper + off, *sc = score + off, *fr = from + off, *rk = rank + off;
    uint8_t *trk = tracked + off; int32_t *cl = cells + off;
    int32_t *coff = ch_off + hoff, *clen = ch_len + hoff, *cscore = ch_score + hoff, *cidx = ch_idx + hoff;
    const int cap = n / 2 + 1, top_N = 1000;
    for (int i = lane; i < n; i += 32) trk[i] = 0;
    for (int i = lane; i < cap && i < top_N; i += 32) cidx[i] = i;
    __syncwarp();
    const int score_n = nrank[r];
    int ch_n = 0, pool = 0; // pool: next free slot in cl[]; the slot of chain ch_n is rewritten until accepted
    bool need_sort_check = false;
#define ST(c) (en[c] - pr[c])
    for (int b0 = 0; b0 < score_n && ch_n < top_N && ch_n < cap; b0 += 32) {
        const int ci = b0 + lane;
        const bool valid = ci < score_n;
        int c = 0, cell_start = 0, cell_end = 0;
        if (valid) { c = rk[ci]; cell_start = ST(row_first(en, c)); cell_end = en[c]; }
        unsigned remaining = __ballot_sync(TH_FULL, valid), surv = 0;
        bool rescreen = true;
        while (remaining && ch_n < top_N && ch_n < cap) {
            bool in_chain = false;
            if (rescreen && ((remaining >> lane) & 1)) { // is_in_chain (:170-185); cell_start is taken from the first cell of the row
                for (int _i = 0; _i < ch_n; ++_i) {
                    const int k = cidx[_i];
                    const int kl = clen[k];
                    if (kl <= 0) continue;
                    const int ko = coff[k];
                    const int c1 = cl[ko], c2 = cl[ko + kl - 1];
                    const int chain_start = ST(c1), chain_end = en[c2];
                    if (chain_end < cell_start) break;
                    else if (chain_start > cell_end) continue;
                    else if (cell_end - chain_start >= (chain_end - chain_start) / 2) { in_chain = true; break; }
                }
            }
            if (rescreen) surv = __ballot_sync(TH_FULL, ((remaining >> lane) & 1) && !in_chain);
            surv &= remaining;
            if (!surv) break;
            const int f = __ffs(surv) - 1;
            const int cc = __shfl_sync(TH_FULL, c, f);
            int resc = 0;
            if (lane == 0) {
                bool accepted = false, changed = false;
                if (!trk[cc]) { // backtrack_dp (:86-111)
                    int s = sc[cc], cur = cc, len = 0;
                    while (true) {
                        trk[cur] = 1; cl[pool + len++] = cur;
                        const int p = fr[cur];
                        if (p == -1) break;
                        if (trk[p]) { s -= sc[p]; break; }
                        cur = p;
                    }
                    for (int a = 0, b = len - 1; a < b; ++a, --b) { int t = cl[pool + a]; cl[pool + a] = cl[pool + b]; cl[pool + b] = t; }
                    coff[ch_n] = pool; clen[ch_n] = len; cscore[ch_n] = s;
                    if (len > 1) { // is_overlap_chain (:54-83)
                        bool ovl = false;
                        if (ch_n > 0) {
                            const int start = ST(cl[pool + len - 1]);
                            for (int j = ch_n - 1; j >= 0; --j) {
                                if (clen[j] <= 0) continue;
                                if (en[cl[coff[j] + clen[j] - 1]] <= start) break;
                                const int s1 = ST(cl[coff[j]]), e1 = ST(cl[coff[j] + clen[j] - 1]);
                                const int s2 = ST(cl[pool]), e2 = ST(cl[pool + len - 1]);
                                const int mn = min(e1 - s1, e2 - s2), ovlp = min(e1, e2) - max(s1, s2);
                                if (ovlp / (mn + 0.0) >= 0.5) {
                                    if (cscore[j] > s) ovl = true; else { clen[j] = 0; changed = true; }
                                    break;
                                }
                            }
                        }
                        accepted = !ovl;
                    }
                }
                if (accepted) { pool += clen[ch_n]; ++ch_n; changed = true; }
                if (changed) need_sort_check = true;
                // sort_chain (:188-207) runs after every candidate in the reference; it performs no swap when the live
                // chain ends are already non-increasing, so that O(ch_n) test replaces the O(ch_n^2) pass, and the test
                // itself is skipped while the chain set is unchanged since the last time it found them ordered.
                if (ch_n >= 2 && need_sort_check) {
                    bool sorted = true; int last_end = INT_MAX;
                    for (int _i = 0; _i < ch_n; ++_i) {
                        const int k = cidx[_i];
                        if (clen[k] <= 0) continue;
                        const int e = en[cl[coff[k] + clen[k] - 1]];
                        if (e > last_end) { sorted = false; break; }
                        last_end = e;
                    }
                    if (sorted) need_sort_check = false;
                    else { changed = true; // the literal pass (incl. its stale `i`); may leave the list unsorted, so check again next time
                        for (int _i = 0; _i < ch_n - 1; ++_i) {
                            const int ii = cidx[_i];
                            if (clen[ii] <= 0) continue;
                            int end1 = en[cl[coff[ii] + clen[ii] - 1]];
                            for (int _j = _i + 1; _j < ch_n; ++_j) {
                                const int jj = cidx[_j];
                                if (clen[jj] <= 0) continue;
                                const int end2 = en[cl[coff[jj] + clen[jj] - 1]];
                                if (end1 < end2) { cidx[_i] = jj; cidx[_j] = ii; end1 = end2; }
                            }
                        }
                    }
                }
                resc = changed;
            }
            ch_n = __shfl_sync(TH_FULL, ch_n, 0);
            rescreen = __shfl_sync(TH_FULL, resc, 0) != 0; // an unchanged chain set keeps the screening of the lanes behind f valid
            __syncwarp();
            remaining &= ~((2u << f) - 1u);
        }
    }
#undef ST
    // post-process (:392-399): ascending end, chains with >= 3 cells
    if (lane == 0) {
        int pn = 0;
        for (int i = ch_n - 1; i >= 0; --i) {
            const int k = cidx[i];
            if (clen[k] - 1 < 2) continue;
            pch_off[hoff + pn] = coff[k]; pch_len[hoff + pn] = clen[k]; ++pn;
        }
        pch_n[r] = pn;
    }
    (void)ragged_ok;
}
