// th_seed.cuh -- read packing, k-mer seeding and hit-distance detection.
//
// Replaces (reference, /root/reference): src/seq.c:77-87 (get_bseq), src/tandem_hit.c:37-56
// (direct_hash), :97-157 (minimizer_hash), :171-225 (collect_hash_hit) and src/ksort.h:101-151.
//
// Data layout in HBM: all reads of a chunk are concatenated, each starting at a 64-base aligned
// offset; the pack kernel turns the ASCII bytes into (a) nt4 byte codes for the DP kernels,
// (b) 2-bit packed words (32 bases per uint64, base i at bits 2*(i%32)) and (c) an N-mask (1 bit per
// base) that the seeding kernel reads.  Seeds are (key<<32|pos) words sorted per read inside one
// thread block (shared memory up to SEED_SMEM_CAP seeds, global scratch beyond); a hit is
// (end<<32|period), sorted the same way.  Algorithmic HBM bytes per read: L/4 + L/8 in, 8*hit_n out.
#pragma once
#include "th_common.cuh"

#define SEED_SMEM_CAP 16384           // 128 KB of 64-bit seeds
#define SEED_THREADS 512
#define SEED_WARPS (SEED_THREADS / 32)
#define SEED_HIST_BYTES (SEED_WARPS * 256 * 4)                 // digit counters of the radix passes
#define SEED_SMEM_BYTES (SEED_SMEM_CAP * 8 + SEED_HIST_BYTES)  // largest launch: seed / hit buffers of a 16 k read + the counters
#define SEED_INVALID 0xffffffffffffffffull

// ASCII -> nt4 (src/seq.c:15-32): ACGT/acgt and raw 0..3 -> 0..3, '-' -> 5, everything else 4
__device__ __forceinline__ uint8_t nt4_code(uint8_t c) {
    uint8_t u = c & 0xDF;
    uint8_t r = 4;
    if (u == 'A') r = 0; else if (u == 'C') r = 1; else if (u == 'G') r = 2; else if (u == 'T') r = 3;
    if (c < 4) r = c;
    if (c == '-') r = 5;
    // '-' & 0xDF == 0x0D, never a letter; letters with bit 5 cleared collide only with their own case
    return r;
}

// one thread = 32 bases: 32 B in, 32 B codes + 8 B packed + 4 B mask out (coalesced 16-byte accesses)
__global__ void pack_kernel(const uint8_t *__restrict__ ascii, uint8_t *__restrict__ bseq,
                            uint64_t *__restrict__ pack2, uint32_t *__restrict__ nmask, int64_t n_words) {
    int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= n_words) return;
    const uint4 *in = reinterpret_cast<const uint4 *>(ascii + wi * 32);
    uint4 *out = reinterpret_cast<uint4 *>(bseq + wi * 32);
    uint64_t pw = 0; uint32_t nm = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint4 v = in[h];
        uint32_t w4[4] = {v.x, v.y, v.z, v.w}, o4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t o = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                uint8_t c = nt4_code((uint8_t)(w4[q] >> (8 * b)));
                int i = h * 16 + q * 4 + b;
                o |= (uint32_t)c << (8 * b);
                pw |= (uint64_t)(c & 3) << (2 * i);
                nm |= (uint32_t)(c >= 4) << i;
            }
            o4[q] = o;
        }
        out[h] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
    }
    pack2[wi] = pw;
    nmask[wi] = nm;
}

// ---- seeders for the non-default options (-H, -w > 1), all threads of the block ------------------------------------
// The reference walks the read once (src/tandem_hit.c:37-56, :97-157) carrying a small state: the rolling k-mer, the
// number of bases since the last N, with -H the lengths of the last k homopolymer runs, and with -w the last w k-mers and
// their rightmost minimum.  That state is a function of a bounded stretch of the read behind the current position -- the
// last w + k steps (a step is a base, with -H a homopolymer run; an N is a step of its own) -- so every thread takes a tile
// of 32 positions, replays the w + k steps before its tile from a blank state without emitting anything, and then emits
// what the steps that START inside its tile emit.  Why the replay reproduces the reference's state:
//  * k-mer, run lengths, span: functions of the last k steps;
//  * the base counter l only enters comparisons against k, w + k - 1 and w + k: after w + k replayed steps without an N
//    both counters are >= w + k + 1, and an N inside the replay resets both at the same place;
//  * the window holds the k-mers of the last w steps, and its `minimum` is always the rightmost smallest entry of the
//    window (new entries win ties, the rescan after the minimum left the window keeps the last of equals), hence
//    determined by the window's content;
//  * the window slot index only serves as a circular order, so its phase is free.
// Seeds go to out[] through a shared counter: their order is irrelevant, the list is sorted next.
struct SeedMin { uint32_t x, y; };
__device__ __forceinline__ void seed_emit(uint64_t *out, int cap, int *counter, uint32_t x, uint32_t y) {
    const int slot = atomicAdd(counter, 1);
    if (slot < cap) out[slot] = (uint64_t)x << 32 | y;
}
// first position >= a where a step starts (with -H: not in the middle of a homopolymer run)
__device__ __forceinline__ int seed_step_start(const uint8_t *b, int len, int a, int hpc) {
    if (hpc) while (a > 0 && a < len && b[a] < 4 && b[a - 1] == b[a]) ++a;
    return a;
}
// start of the step `steps` steps before the step starting at a
__device__ __forceinline__ int seed_steps_back(const uint8_t *b, int a, int steps, int hpc) {
    while (a > 0 && steps > 0) {
        --a;
        if (hpc) while (a > 0 && b[a] < 4 && b[a - 1] == b[a]) --a;
        --steps;
    }
    return a;
}
__device__ void seeds_parallel(const uint8_t *bseq, int len, int k, int w, int hpc, uint64_t *out, int cap, int *counter) {
    const uint32_t mask = (uint32_t)((1ull << 2 * k) - 1), NONE = 0xffffffffu;
    for (int t0 = threadIdx.x * 32; t0 < len; t0 += blockDim.x * 32) {
        const int own_lo = seed_step_start(bseq, len, t0, hpc);
        int own_hi = min(t0 + 32, len);                  // steps starting in [own_lo, own_hi) are this thread's
        const bool last_tile = t0 + 32 >= len;
        if (own_lo >= own_hi && !last_tile) continue;
        int i = seed_steps_back(bseq, own_lo, w + k, hpc);
        int l = 0, span = 0, bp = 0, minp = 0;
        uint32_t key = 0;
        int rq[32], rq_front = 0, rq_count = 0;           // lengths of the last k runs (-H)
        SeedMin win[256], mn = {NONE, NONE};             // w <= 255 (th_gpu_create)
        if (w > 1) for (int j = 0; j < w; ++j) win[j] = mn;
        for (; i < own_hi; ++i) {
            const bool emit = i >= own_lo;
            const int c = bseq[i];
            SeedMin info = {NONE, NONE};
            if (c < 4) {
                if (hpc) {
                    int run = 1;
                    while (i + run < len && bseq[i + run] == c) ++run;
                    i += run - 1;
                    rq[(rq_count++ + rq_front) & 31] = run; span += run;
                    if (rq_count > k) { span -= rq[rq_front]; rq_front = (rq_front + 1) & 31; --rq_count; }
                } else span = min(l + 1, k);
                key = (key << 2 | (uint32_t)c) & mask;
                ++l;
                if (w <= 1) { if (l >= k && emit) seed_emit(out, cap, counter, key, (uint32_t)i); continue; } // direct_hash with -H
                if (l >= k && span < 256) { info.x = key; info.y = (uint32_t)i; }
            } else { l = 0; rq_count = rq_front = 0; span = 0; key = 0; if (w <= 1) continue; }
            win[bp] = info;
            if (l == w + k - 1 && mn.x != NONE && emit) // the first full window: its other copies of the minimum
                for (int j = 0; j < w; ++j) if (j != bp && win[j].x == mn.x && win[j].y != mn.y) seed_emit(out, cap, counter, win[j].x, win[j].y);
            if (info.x <= mn.x) { // a new minimum enters
                if (l >= w + k && mn.x != NONE && emit) seed_emit(out, cap, counter, mn.x, mn.y);
                mn = info; minp = bp;
            } else if (bp == minp) { // the minimum left the window: report it, find the next one (the last of equals, oldest slot first)
                if (l >= w + k - 1 && mn.x != NONE && emit) seed_emit(out, cap, counter, mn.x, mn.y);
                mn.x = NONE;
                for (int q = 1; q <= w; ++q) { const int j = bp + q < w ? bp + q : bp + q - w; if (mn.x >= win[j].x) { mn = win[j]; minp = j; } }
                if (l >= w + k - 1 && mn.x != NONE && emit)
                    for (int j = 0; j < w; ++j) if (win[j].x == mn.x && win[j].y != mn.y) seed_emit(out, cap, counter, win[j].x, win[j].y);
            }
            if (++bp == w) bp = 0;
        }
        if (last_tile && w > 1 && mn.x != NONE) seed_emit(out, cap, counter, mn.x, mn.y); // the minimum left when the read ends
    }
}

// ---- block-wide helpers of the default path ----------------------------------------------------------------------------
// exclusive prefix sum of one int per thread (SEED_THREADS threads); returns the thread's offset, total in `total`
__device__ __forceinline__ int seed_block_scan(int v, int *s_part, int &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(TH_FULL, x, d); if (lane >= d) x += y; }
    __syncthreads();                                   // s_part may still be read from the previous call
    if (lane == 31) s_part[wid] = x;
    __syncthreads();
    int p = lane < SEED_WARPS ? s_part[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(TH_FULL, p, d); if (lane >= d) p += y; }
    total = __shfl_sync(TH_FULL, p, 31);
    const int wbase = __shfl_sync(TH_FULL, p, max(wid - 1, 0));
    return x - v + (wid > 0 ? wbase : 0);
}
// One stable counting pass of an LSD radix sort over an 8-bit digit: src[0..n) -> dst, hist = SEED_WARPS x 256 ints (8 per
// thread in the scan: SEED_WARPS x 256 = 8 SEED_THREADS).  Warp w owns the w-th contiguous part of the input, so "stable" is (warp, position inside the warp's part); inside a chunk of 32 the
// rank among equal digits comes from __match_any_sync.
__device__ void seed_radix_pass(const uint32_t *src, uint32_t *dst, int n, int shift, uint32_t dmask, int *hist, int *s_part) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int per = ((n + SEED_WARPS - 1) / SEED_WARPS + 31) & ~31; // elements per warp, a multiple of 32
    const int lo = min(wid * per, n), hi = min(lo + per, n);
    for (int j = threadIdx.x; j < SEED_WARPS * 256; j += blockDim.x) hist[j] = 0;
    __syncthreads();
    int *my = hist + wid * 256;
    for (int j0 = lo; j0 < hi; j0 += 32) {
        const int j = j0 + lane;
        const bool in = j < hi;
        const uint32_t d = in ? (src[j] >> shift) & dmask : 0x1ffu; // lanes past the end form their own group
        const unsigned peers = __match_any_sync(TH_FULL, d);
        if (in && (peers & ((1u << lane) - 1)) == 0) my[d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    { // exclusive scan in (digit, warp) order: thread t takes the 8 consecutive entries 8t .. 8t+7 of that order
        int v[8], sum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int e = threadIdx.x * 8 + q; v[q] = hist[(e % SEED_WARPS) * 256 + e / SEED_WARPS]; sum += v[q]; }
        int total; int base = seed_block_scan(sum, s_part, total);
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int e = threadIdx.x * 8 + q; hist[(e % SEED_WARPS) * 256 + e / SEED_WARPS] = base; base += v[q]; }
    }
    __syncthreads();
    for (int j0 = lo; j0 < hi; j0 += 32) {
        const int j = j0 + lane;
        const bool in = j < hi;
        const uint32_t v = in ? src[j] : 0, d = in ? (v >> shift) & dmask : 0x1ffu;
        const unsigned peers = __match_any_sync(TH_FULL, d);
        const unsigned lower = peers & ((1u << lane) - 1);
        int base = 0;
        if (in) base = my[d];
        __syncwarp();
        if (in) { dst[base + __popc(lower)] = v; if (lower == 0) my[d] = base + __popc(peers); }
        __syncwarp();
    }
    __syncthreads();
}

// One block per read (grid-stride).  gscratch: 2 * gcap uint64 per block (sort buffer for long reads +
// hit staging).  Output: hend/hper at the read's base offset, nhits[r].
// Fast path (default options, reads up to 32 k bases): seeds are 32-bit words (key << bits(L) | pos), hits are
// (end << bits(L) | period); both sorts and the hit staging stay in shared memory.  Same total order as the
// reference's 64-bit words (key major, position minor / end major, period minor), so any correct sort is exact.
__global__ void __launch_bounds__(SEED_THREADS, 2)
seed_kernel(DevParams P, int n_reads, int smem_bytes, const int64_t *__restrict__ roff, const int32_t *__restrict__ rlen,
            const uint8_t *__restrict__ bseq, const uint64_t *__restrict__ pack2, const uint32_t *__restrict__ nmask,
            uint64_t *__restrict__ gscratch, int64_t gcap,
            int32_t *__restrict__ hend, int32_t *__restrict__ hper, int32_t *__restrict__ nhits) {
    extern __shared__ uint64_t sbuf[];
    __shared__ int s_cnt;
    __shared__ int s_part[32];
    uint64_t *gbuf = gscratch + (int64_t)blockIdx.x * 2 * gcap, *gtmp = gbuf + gcap;
    const uint32_t kmask = (uint32_t)((1ull << 2 * P.k) - 1);
    for (int r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const int L = rlen[r];
        const int64_t off = roff[r];
        if (L < P.k || L - P.w <= 0) { if (threadIdx.x == 0) nhits[r] = 0; continue; }
        const int npow = next_pow2(L);
        const int bitsL = 32 - __clz(npow - 1 > 0 ? npow - 1 : 1);
        const uint64_t *pw = pack2 + off / 32; const uint32_t *nm = nmask + off / 32;
        const int cap = (L + 31) & ~31;                  // buffer entries of the default path: a tile of 32 positions per thread
        if (P.w <= 1 && !P.hpc && 2 * P.k + bitsL < 32 && P.min_p >= 1 && cap <= SEED_THREADS * 32 && 8 * cap + SEED_HIST_BYTES <= smem_bytes) {
            uint32_t *buf = reinterpret_cast<uint32_t *>(sbuf), *hbuf = buf + cap;  // 2 x cap x 4 B
            int *hist = reinterpret_cast<int *>(hbuf + cap);                          // SEED_WARPS x 256 digit counters
            const uint32_t posmask = (1u << bitsL) - 1;
            // rolling 2-bit k-mer, one tile of 32 positions per thread (cap <= 16384: at most 512 tiles), k-1 bases of warm-up
            // (tandem_hit.c:37-56); the tile's seeds stay in registers until their place in the compacted list is known
            const int base0 = threadIdx.x * 32;
            uint32_t valid = 0;
            if (base0 < L) {
                uint32_t key = 0; int l = 0;
                int start = base0 - (P.k - 1); if (start < 0) start = 0;
                const int stop = base0 + 32 < L ? base0 + 32 : L;
                const uint64_t w0 = pw[start >> 5], w1 = pw[base0 >> 5];
                const uint32_t m0 = nm[start >> 5], m1 = nm[base0 >> 5];
                for (int p = start; p < base0; ++p) {
                    if ((m0 >> (p & 31)) & 1) { key = 0; l = 0; }
                    else { key = ((key << 2) | (uint32_t)((w0 >> (2 * (p & 31))) & 3)) & kmask; ++l; }
                }
                for (int q = 0; base0 + q < stop; ++q) {
                    if ((m1 >> q) & 1) { key = 0; l = 0; }
                    else {
                        key = ((key << 2) | (uint32_t)((w1 >> (2 * q)) & 3)) & kmask;
                        if (++l >= P.k) { hbuf[base0 + q] = key << bitsL | (uint32_t)(base0 + q); valid |= 1u << q; }
                    }
                }
            }
            int n_seed;
            {
                int o = seed_block_scan(__popc(valid), s_part, n_seed);
                for (uint32_t m = valid; m; m &= m - 1) buf[o++] = hbuf[base0 + __ffs(m) - 1]; // the thread's own writes
            }
            __syncthreads();
            if (n_seed == 0) { if (threadIdx.x == 0) nhits[r] = 0; continue; }
            // sort by key, positions ascending inside a key: the list is in position order, so a STABLE sort on the 2k key bits is
            // the reference's order (key major, position minor; ksort.h:101-151 sorts the 64-bit words)
            uint32_t *sorted = buf, *other = hbuf;
            for (int sh = 0; sh < 2 * P.k; sh += 8) {
                const int bits = min(8, 2 * P.k - sh);
                seed_radix_pass(sorted, other, n_seed, bitsL + sh, (1u << bits) - 1, hist, s_part);
                uint32_t *t = sorted; sorted = other; other = t;
            }
            // nearest earlier occurrence of the same key at distance >= min_p (tandem_hit.c:186-214).  A position yields at most
            // one hit, so the (end, period) order of the hit list is the position order: periods are scattered by position and
            // compacted in order -- no second sort
            for (int j = threadIdx.x; j < cap; j += blockDim.x) other[j] = 0;
            __syncthreads();
            for (int j = threadIdx.x; j < n_seed; j += blockDim.x) {
                const uint32_t cur = sorted[j], key = cur >> bitsL, pos = cur & posmask; uint32_t d = 0; bool found = false;
                for (int kk = j - 1; kk >= 0; --kk) {
                    const uint32_t o = sorted[kk];
                    if ((o >> bitsL) != key) break;
                    d = pos - (o & posmask);
                    if (d >= P.min_p) { found = true; break; }
                }
                if (found && d <= P.max_p) other[pos] = d;   // d >= min_p >= 1: 0 means no hit
            }
            __syncthreads();
            {
                uint32_t has = 0;
                for (int q = 0; q < 32 && base0 + q < L; ++q) if (other[base0 + q]) has |= 1u << q;
                int n_hit; int o = seed_block_scan(__popc(has), s_part, n_hit);
                for (uint32_t m = has; m; m &= m - 1) { const int q = __ffs(m) - 1; hend[off + o] = base0 + q; hper[off + o] = (int32_t)other[base0 + q]; ++o; }
                if (threadIdx.x == 0) nhits[r] = n_hit;
            }
            __syncthreads();
            continue;
        }
        // ---- general path: 64-bit words, any option ----
        const int smem_words = smem_bytes / 8;
        uint64_t *buf = npow <= smem_words ? sbuf : gbuf;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if (P.w <= 1 && !P.hpc) {
            for (int t = threadIdx.x; t * 32 < npow; t += blockDim.x) {
                int base0 = t * 32;
                if (base0 >= L) { for (int p = base0; p < base0 + 32 && p < npow; ++p) buf[p] = SEED_INVALID; continue; }
                uint32_t key = 0; int l = 0, cnt = 0;
                int start = base0 - (P.k - 1); if (start < 0) start = 0;
                int stop = base0 + 32 < L ? base0 + 32 : L;
                uint64_t w0 = pw[start >> 5], w1 = pw[base0 >> 5];
                uint32_t m0 = nm[start >> 5], m1 = nm[base0 >> 5];
                for (int p = start; p < stop; ++p) {
                    bool cur = p >= base0;
                    uint64_t wd = cur ? w1 : w0; uint32_t md = cur ? m1 : m0;
                    uint64_t v = SEED_INVALID;
                    if ((md >> (p & 31)) & 1) { key = 0; l = 0; }
                    else {
                        key = ((key << 2) | (uint32_t)((wd >> (2 * (p & 31))) & 3)) & kmask;
                        if (++l >= P.k) { v = (uint64_t)key << 32 | (uint32_t)p; if (cur) ++cnt; }
                    }
                    if (cur) buf[p] = v;
                }
                for (int p = stop; p < base0 + 32 && p < npow; ++p) buf[p] = SEED_INVALID;
                if (cnt) atomicAdd(&s_cnt, cnt);
            }
        } else {
            seeds_parallel(bseq + off, L, P.k, P.w, P.hpc, buf, npow, &s_cnt);
            __syncthreads();
            if (threadIdx.x == 0 && s_cnt > npow) s_cnt = npow; // more seeds than bases cannot happen (one minimum leaves per step); a guard, not a path
            __syncthreads();
            for (int p = s_cnt + threadIdx.x; p < npow; p += blockDim.x) buf[p] = SEED_INVALID;
        }
        __syncthreads();
        const int n_seed = s_cnt;
        __syncthreads();
        if (n_seed == 0) { if (threadIdx.x == 0) nhits[r] = 0; continue; }
        block_bitonic_sort<false>(buf, npow);
        // nearest earlier occurrence of the same key at distance >= min_p (tandem_hit.c:186-214); hits are compacted
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        for (int j0 = 0; j0 < n_seed; j0 += blockDim.x) {
            const int j = j0 + threadIdx.x;
            uint64_t v = SEED_INVALID;
            if (j < n_seed) {
                uint64_t cur = buf[j]; uint32_t key = (uint32_t)(cur >> 32), pos = (uint32_t)cur, d = 0; bool found = false;
                for (int kk = j - 1; kk >= 0; --kk) {
                    uint64_t o = buf[kk];
                    if ((uint32_t)(o >> 32) != key) break;
                    d = pos - (uint32_t)o;
                    if (d >= P.min_p) { found = true; break; }
                }
                if (found && d <= P.max_p) v = (uint64_t)pos << 32 | d;
            }
            const unsigned m = __ballot_sync(TH_FULL, v != SEED_INVALID);
            int wbase = 0;
            if ((threadIdx.x & 31) == 0 && m) wbase = atomicAdd(&s_cnt, __popc(m));
            wbase = __shfl_sync(TH_FULL, wbase, 0);
            if (v != SEED_INVALID) gtmp[wbase + __popc(m & ((1u << (threadIdx.x & 31)) - 1))] = v;
        }
        __syncthreads();
        const int n_hit = s_cnt;
        const int hpow = next_pow2(n_hit);
        uint64_t *hb = hpow <= smem_words ? sbuf : gtmp; // the seeds are no longer needed
        if (hb != gtmp) for (int j = threadIdx.x; j < n_hit; j += blockDim.x) hb[j] = gtmp[j];
        for (int j = n_hit + threadIdx.x; j < hpow; j += blockDim.x) hb[j] = SEED_INVALID;
        __syncthreads();
        if (n_hit > 1) block_bitonic_sort<false>(hb, hpow);
        for (int j = threadIdx.x; j < n_hit; j += blockDim.x) {
            uint64_t v = hb[j];
            hend[off + j] = (int32_t)(v >> 32);
            hper[off + j] = (int32_t)(uint32_t)v;
        }
        if (threadIdx.x == 0) nhits[r] = n_hit;
        __syncthreads();
    }
}
