// th_tasks.cuh -- consensus tasks and post-consensus alignment items, built on the device from the partition result.
//
// Replaces the bookkeeping half of seqs_msa (src/gen_cons.c:173-223): par_pos runs between -1 separators with more
// than min_copy entries become consensus tasks (src/gen_cons.c:191-200, src/tidehunter.c:42); the units that enter the
// POA are those inside the read (src/abpoa_cons.c:40-50); every unit of a run is aligned to the consensus afterwards
// (ksw2_global, :208-216) and the run's two ends are extended (:217-223).  Three passes over the reads: count, prefix
// sums, fill -- so that the only thing the host has to wait for between the partition and the consensus kernels is six
// totals to size its buffers with.
#pragma once
#include "th_common.cuh"
#include "th_partition.cuh"
#include "th_poa.cuh"

// visits the tasks of read r in the order the reference emits their records: f(par, i, j) for the run par[i..j)
template <class F>
__device__ __forceinline__ void th_for_each_task(int r, int min_copy, const int64_t *roff, const int32_t *pch_n, const int32_t *par,
                                                 const int32_t *par_off, const int32_t *par_n, F f) {
    const int64_t off = roff[r], hoff = off / 2;
    const int32_t *p = par + 2 * off;
    const int nch = pch_n[r];
    for (int c = 0; c < nch; ++c) {
        const int n = par_n[hoff + c]; const int32_t *pp = p + par_off[hoff + c];
        if (n < min_copy + 1) continue; // src/tidehunter.c:42
        int i = 0;
        while (i < n - min_copy) {
            if (pp[i] < 0) { ++i; continue; }
            int j;
            for (j = i + 1; j < n; ++j) if (pp[j] < 0) break;
            if (j - i > min_copy) f(pp, i, j);
            i = j + 1;
        }
    }
}

// the units of a run that enter the consensus (src/abpoa_cons.c:40-50)
__device__ __forceinline__ void th_run_units(const int32_t *pp, int i, int j, int L, int &nseq, int &sum, int &qmax) {
    nseq = 0; sum = 0; qmax = 0;
    for (int q = i; q < j - 1; ++q) {
        const int start = pp[q], end = pp[q + 1];
        if (start < 0 || end < 0 || start >= L - 1 || end + 1 > L) continue;
        ++nseq; sum += end - start; qmax = max(qmax, end - start);
    }
}

__global__ void task_count_kernel(int n_reads, int min_copy, const int64_t *__restrict__ roff, const int32_t *__restrict__ rlen,
                                  const int32_t *__restrict__ pch_n, const int32_t *__restrict__ par, const int32_t *__restrict__ par_off,
                                  const int32_t *__restrict__ par_n, const int32_t *__restrict__ nhits, int32_t *__restrict__ counts,
                                  TaskTotals *tot) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    int c[TC_N] = {0, 0, 0, 0, 0, 0};
    unsigned long long slab = 0, wide = 0; long long dense = 0; int max_key = 0;
    const int L = rlen[r];
    th_for_each_task(r, min_copy, roff, pch_n, par, par_off, par_n, [&](const int32_t *pp, int i, int j) {
        int nseq, sum, qmax; th_run_units(pp, i, j, L, nseq, sum, qmax);
        c[TC_TASKS] += 1; c[TC_UNITS] += nseq; c[TC_POS] += j - i; c[TC_PAIR3] += (j - 1 - i) >> 1; c[TC_LEFT] += (j - 1 - i) & 1; c[TC_CONS] += sum + 4;
        if (nseq > 2) {
            const unsigned long long s = poa_slab_need(sum + 2, qmax, nseq, false), sw = poa_slab_need(sum + 2, qmax, nseq, true);
            if (s > slab) slab = s;
            if (sw > wide) wide = sw;
        }
        dense += qmax + 64;
        const long long key = (long long)(sum + 2) * nseq; max_key = max(max_key, (int)min(key, 0x7fffffffll));
    });
#pragma unroll
    for (int k = 0; k < TC_N; ++k) counts[(size_t)k * n_reads + r] = c[k];
    if (slab) atomicMax(&tot->slab_typ, slab);
    if (wide) atomicMax(&tot->slab_wide, wide);
    if (dense) atomicAdd((unsigned long long *)&tot->dense_bound, (unsigned long long)dense);
    if (max_key) atomicMax(&tot->max_key, max_key);
    atomicAdd((unsigned long long *)&tot->n_hits, (unsigned long long)nhits[r]);
}

// exclusive prefix sums of TC_N arrays of n ints each, in place; totals to tot->n.  One block.
__global__ void __launch_bounds__(1024) task_scan_kernel(int n, int32_t *__restrict__ counts, TaskTotals *tot) {
    __shared__ int s_part[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + 1023) / 1024;
    for (int k = 0; k < TC_N; ++k) {
        int32_t *a = counts + (size_t)k * n;
        const int lo = min(n, tid * per), hi = min(n, lo + per);
        int sum = 0;
        for (int i = lo; i < hi; ++i) sum += a[i];
        int inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(TH_FULL, inc, d); if (lane >= d) inc += o; }
        if (lane == 31) s_part[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int v = s_part[lane], w = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(TH_FULL, w, d); if (lane >= d) w += o; }
            s_part[lane] = w - v; // exclusive
            if (lane == 31) tot->n[k] = w;
        }
        __syncthreads();
        int run = s_part[wid] + inc - sum;
        for (int i = lo; i < hi; ++i) { const int v = a[i]; a[i] = run; run += v; }
        __syncthreads();
    }
}

// left-over unit of a run with an odd number of units: two of them (of different tasks) share a warp later
struct LeftUnit { int32_t task, a, b, out; int64_t seq_off; };

__global__ void task_fill_kernel(int n_reads, int min_copy, int only_unit, const int64_t *__restrict__ roff, const int32_t *__restrict__ rlen,
                                 const int32_t *__restrict__ pch_n, const int32_t *__restrict__ par, const int32_t *__restrict__ par_off,
                                 const int32_t *__restrict__ par_n, const int32_t *__restrict__ offs, TaskTotals *tot,
                                 PoaTask *__restrict__ tasks, int32_t *__restrict__ ustart, int32_t *__restrict__ ulen, int32_t *__restrict__ pos,
                                 int32_t *__restrict__ read_task_off, int32_t *__restrict__ task_pos_off, int32_t *__restrict__ task_n_seqs,
                                 int32_t *__restrict__ task_key, int32_t *__restrict__ ext_key, KswItem *__restrict__ items, LeftUnit *__restrict__ left) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    int t = offs[(size_t)TC_TASKS * n_reads + r], u = offs[(size_t)TC_UNITS * n_reads + r], ps = offs[(size_t)TC_POS * n_reads + r];
    int p3 = offs[(size_t)TC_PAIR3 * n_reads + r], lf = offs[(size_t)TC_LEFT * n_reads + r], co = offs[(size_t)TC_CONS * n_reads + r];
    const int nt = tot->n[TC_TASKS];
    const int n_pairs = tot->n[TC_PAIR3] + (tot->n[TC_LEFT] >> 1), n_single = tot->n[TC_LEFT] & 1;
    KswItem *exts = items + n_pairs + n_single;
    read_task_off[r] = t;
    if (r == n_reads - 1) read_task_off[n_reads] = nt;
    if (r == 0) task_pos_off[0] = 0;
    const int L = rlen[r]; const int64_t so = roff[r];
    int max_ext = 0;
    th_for_each_task(r, min_copy, roff, pch_n, par, par_off, par_n, [&](const int32_t *pp, int i, int j) {
        PoaTask T; T.seq_off = so; T.read = r; T.unit_off = u;
        int nseq = 0, sum = 0, qmax = 0;
        for (int q = i; q < j - 1; ++q) { // src/abpoa_cons.c:40-50
            const int start = pp[q], end = pp[q + 1];
            if (start < 0 || end < 0 || start >= L - 1 || end + 1 > L) continue;
            ustart[u] = start + 1; ulen[u] = end - start; ++u; ++nseq; sum += end - start; qmax = max(qmax, end - start);
        }
        T.n_seqs = nseq; T.ncap = sum + 2; T.qmax = qmax; T.cons_off = co;
        co += sum + 4;
        tasks[t] = T;
        task_n_seqs[t] = nseq;
        const long long key = (long long)(sum + 2) * nseq; task_key[t] = (int)min(key, 0x7fffffffll);
        const int p0 = ps;
        for (int q = i; q < j; ++q) pos[ps++] = pp[q];
        task_pos_off[t + 1] = ps;
        if (!only_unit) { // post-consensus alignments of seqs_msa (src/gen_cons.c:208-223); units go two per warp
            for (int q = i; q < j - 1; q += 2) {
                if (q + 1 < j - 1) {
                    KswItem it; it.kind = 3; it.task = t; it.seq_off = so; it.out = p0 + (q - i);
                    it.a = pp[q] + 1; it.b = pp[q + 1] - pp[q]; it.a2 = pp[q + 1] + 1; it.b2 = pp[q + 2] - pp[q + 1];
                    it.task2 = 0; it.out2 = 0; it.pad = 0; it.seq_off2 = 0;
                    items[p3++] = it;
                } else { LeftUnit lu; lu.task = t; lu.a = pp[q] + 1; lu.b = pp[q + 1] - pp[q]; lu.out = p0 + (q - i); lu.seq_off = so; left[lf++] = lu; }
            }
            KswItem le; le.kind = 1; le.task = t; le.a = pp[i] + 1; le.b = 0; le.a2 = le.b2 = 0; le.seq_off = so; le.out = 4 * t; le.task2 = le.out2 = le.pad = 0; le.seq_off2 = 0;
            KswItem re = le; re.kind = 2; re.a = pp[j - 1] + 1; re.b = L - pp[j - 1] - 1; re.out = 4 * t + 2;
            exts[2 * t] = le; exts[2 * t + 1] = re;
            { const int ka = min(max(le.a, 0), 2 * qmax), kb = min(max(re.b, 0), 2 * qmax); // rows the extension computes: at most twice the consensus length
              ext_key[2 * t] = ka; ext_key[2 * t + 1] = kb; max_ext = max(max_ext, max(ka, kb)); }
        }
        ++t;
    });
    if (max_ext > 0) atomicMax(&tot->max_ext, max_ext);
}

// left-over units, two per item (kind 4); an odd last one is the single item
__global__ void task_pair_left_kernel(const TaskTotals *__restrict__ tot, const LeftUnit *__restrict__ left, KswItem *__restrict__ items) {
    const int n_left = tot->n[TC_LEFT], n3 = tot->n[TC_PAIR3];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * k >= n_left) return;
    const LeftUnit a = left[2 * k];
    KswItem it; it.task = a.task; it.a = a.a; it.b = a.b; it.out = a.out; it.seq_off = a.seq_off; it.pad = 0;
    it.a2 = it.b2 = it.task2 = it.out2 = 0; it.seq_off2 = 0;
    if (2 * k + 1 < n_left) {
        const LeftUnit b = left[2 * k + 1];
        it.kind = 4; it.task2 = b.task; it.a2 = b.a; it.b2 = b.b; it.out2 = b.out; it.seq_off2 = b.seq_off;
        items[n3 + k] = it;
    } else { it.kind = 0; items[n3 + (n_left >> 1)] = it; }
}

// Order by decreasing key in 1024 size classes (a counting sort; the order inside a class is arbitrary).  Used for the task
// order of the persistent POA groups (largest first: the order only balances the load, every task's result is independent
// of it) and to pair boundary extensions of similar target length.  n = mult x tot->n[TC_TASKS].  One block.
// n_direct >= 0: n and the largest key are given by the caller instead (the chaining DP's read order by hit count).
__global__ void __launch_bounds__(1024) bucket_order_kernel(const TaskTotals *__restrict__ tot, int mult, const int32_t *__restrict__ max_key,
                                                            const int32_t *__restrict__ key, int32_t *__restrict__ order, int n_direct = -1, int max_key_direct = 1) {
    __shared__ int s_cnt[1024], s_part[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nt = n_direct >= 0 ? n_direct : mult * tot->n[TC_TASKS];
    const unsigned long long mk = (unsigned long long)max(n_direct >= 0 ? max_key_direct : *max_key, 1);
    s_cnt[tid] = 0;
    __syncthreads();
    for (int t = tid; t < nt; t += 1024) atomicAdd(&s_cnt[1023 - (int)min((unsigned long long)key[t] * 1023ull / mk, 1023ull)], 1);
    __syncthreads();
    { // exclusive scan of the 1024 counts
        const int v = s_cnt[tid]; int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(TH_FULL, inc, d); if (lane >= d) inc += o; }
        if (lane == 31) s_part[wid] = inc;
        __syncthreads();
        if (wid == 0) { int w = s_part[lane], x = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(TH_FULL, x, d); if (lane >= d) x += o; }
            s_part[lane] = x - w; }
        __syncthreads();
        s_cnt[tid] = s_part[wid] + inc - v;
    }
    __syncthreads();
    for (int t = tid; t < nt; t += 1024) order[atomicAdd(&s_cnt[1023 - (int)min((unsigned long long)key[t] * 1023ull / mk, 1023ull)], 1)] = t;
}

// dense consensus offsets: exclusive prefix sum of cons_len over the tasks (one block), total to tot->cons_total
__global__ void __launch_bounds__(1024) cons_scan_kernel(TaskTotals *tot, const int32_t *__restrict__ cons_len, int32_t *__restrict__ task_cons_off) {
    __shared__ long long s_part[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = tot->n[TC_TASKS];
    const int per = (n + 1023) / 1024;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    long long sum = 0;
    for (int i = lo; i < hi; ++i) sum += cons_len[i];
    long long inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const long long o = __shfl_up_sync(TH_FULL, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) s_part[wid] = inc;
    __syncthreads();
    if (wid == 0) { long long v = s_part[lane], w = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const long long o = __shfl_up_sync(TH_FULL, w, d); if (lane >= d) w += o; }
        s_part[lane] = w - v; if (lane == 31) tot->cons_total = w; }
    __syncthreads();
    long long run = s_part[wid] + inc - sum;
    for (int i = lo; i < hi; ++i) { task_cons_off[i] = (int32_t)run; run += cons_len[i]; }
    if (tid == 0) task_cons_off[n] = (int32_t)tot->cons_total;
}

// dense copy of the consensus sequences (and coverages): one warp per task
__global__ void cons_gather_kernel(const TaskTotals *__restrict__ tot, const PoaTask *__restrict__ tasks, const int32_t *__restrict__ cons_len,
                                   const int32_t *__restrict__ task_cons_off, const uint8_t *__restrict__ cons_b, const int32_t *__restrict__ cons_c,
                                   uint8_t *__restrict__ dense_b, int32_t *__restrict__ dense_c, long long dense_cap) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= tot->n[TC_TASKS]) return;
    const int l = cons_len[w]; const long long d = task_cons_off[w]; const int64_t s = tasks[w].cons_off;
    if (d + l > dense_cap) return; // the host fetches what did not fit with a second, exact copy
    for (int i = lane; i < l; i += 32) { dense_b[d + i] = cons_b[s + i]; if (dense_c) dense_c[d + i] = cons_c[s + i]; }
}

// consensus buffer offset of every task, as an array (what the ksw kernels index by task)
__global__ void task_cons_off_kernel(int nt, const PoaTask *__restrict__ tasks, int32_t *__restrict__ cons_off) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nt) cons_off[t] = tasks[t].cons_off;
}
