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
#define SEED_THREADS 1024
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

// literal sequential seeders for the non-default options (-H and -w > 1); one thread per read.
__device__ int seeds_direct_hpc(const uint8_t *bseq, int len, int k, uint64_t *h) {
    uint32_t key = 0, mask = (uint32_t)((1ull << 2 * k) - 1);
    int l = 0, n = 0;
    for (int pos = 0; pos < len; ++pos) {
        int c = bseq[pos];
        if (c >= 4) { key = 0; l = 0; continue; }
        while (pos + 1 < len && bseq[pos + 1] == c) ++pos;
        key = key << 2 | (uint32_t)c;
        if (++l >= k) { key &= mask; h[n++] = (uint64_t)key << 32 | (uint32_t)pos; }
    }
    return n;
}
__device__ int seeds_minimizer(const uint8_t *bseq, int len, int k, int w, int hpc, uint64_t *h) {
    struct mm { uint32_t x, y; };
    int l = 0, n = 0, span = 0, bp = 0, minp = 0;
    uint32_t key = 0, mask = (uint32_t)((1ull << 2 * k) - 1);
    mm buf[256]; mm mn = {0xffffffffu, 0xffffffffu};
    int tq[32], tq_front = 0, tq_count = 0;
    for (int j = 0; j < 256; ++j) buf[j] = mn;
#define TH_EMIT(e) (h[n++] = (uint64_t)(e).x << 32 | (e).y)
    for (int i = 0; i < len; ++i) {
        int c = bseq[i];
        mm info = {0xffffffffu, 0xffffffffu};
        if (c < 4) {
            if (hpc) {
                int skip = 1;
                if (i + 1 < len && bseq[i + 1] == c) {
                    for (skip = 2; i + skip < len; ++skip) if (bseq[i + skip] != c) break;
                    i += skip - 1;
                }
                tq[(tq_count++ + tq_front) & 0x1f] = skip;
                span += skip;
                if (tq_count > k) { span -= tq[tq_front++]; tq_front &= 0x1f; --tq_count; }
            } else span = l + 1 < k ? l + 1 : k;
            key = (key << 2 | (uint32_t)c) & mask;
            ++l;
            if (l >= k && span < 256) { info.x = key; info.y = (uint32_t)i; }
        } else { l = 0; tq_count = tq_front = 0; span = 0; key = 0; }
        buf[bp] = info;
        if (l == w + k - 1 && mn.x != 0xffffffffu) {
            for (int j = bp + 1; j < w; ++j) if (mn.x == buf[j].x && buf[j].y != mn.y) TH_EMIT(buf[j]);
            for (int j = 0; j < bp; ++j) if (mn.x == buf[j].x && buf[j].y != mn.y) TH_EMIT(buf[j]);
        }
        if (info.x <= mn.x) {
            if (l >= w + k && mn.x != 0xffffffffu) TH_EMIT(mn);
            mn = info; minp = bp;
        } else if (bp == minp) {
            if (l >= w + k - 1 && mn.x != 0xffffffffu) TH_EMIT(mn);
            mn.x = 0xffffffffu;
            for (int j = bp + 1; j < w; ++j) if (mn.x >= buf[j].x) { mn = buf[j]; minp = j; }
            for (int j = 0; j <= bp; ++j) if (mn.x >= buf[j].x) { mn = buf[j]; minp = j; }
            if (l >= w + k - 1 && mn.x != 0xffffffffu) {
                for (int j = bp + 1; j < w; ++j) if (mn.x == buf[j].x && mn.y != buf[j].y) TH_EMIT(buf[j]);
                for (int j = 0; j <= bp; ++j) if (mn.x == buf[j].x && mn.y != buf[j].y) TH_EMIT(buf[j]);
            }
        }
        if (++bp == w) bp = 0;
    }
    if (mn.x != 0xffffffffu) TH_EMIT(mn);
#undef TH_EMIT
    return n;
}

// One block per read (grid-stride).  gscratch: 2 * gcap uint64 per block (sort buffer for long reads +
// hit staging).  Output: hend/hper at the read's base offset, nhits[r].
// Fast path (default options, reads up to 32 k bases): seeds are 32-bit words (key << bits(L) | pos), hits are
// (end << bits(L) | period); both sorts and the hit staging stay in shared memory.  Same total order as the
// reference's 64-bit words (key major, position minor / end major, period minor), so any correct sort is exact.
__global__ void __launch_bounds__(SEED_THREADS, 1)
seed_kernel(DevParams P, int n_reads, const int64_t *__restrict__ roff, const int32_t *__restrict__ rlen,
            const uint8_t *__restrict__ bseq, const uint64_t *__restrict__ pack2, const uint32_t *__restrict__ nmask,
            uint64_t *__restrict__ gscratch, int64_t gcap,
            int32_t *__restrict__ hend, int32_t *__restrict__ hper, int32_t *__restrict__ nhits) {
    extern __shared__ uint64_t sbuf[];
    __shared__ int s_cnt;
    uint64_t *gbuf = gscratch + (int64_t)blockIdx.x * 2 * gcap, *gtmp = gbuf + gcap;
    const uint32_t kmask = (uint32_t)((1ull << 2 * P.k) - 1);
    for (int r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const int L = rlen[r];
        const int64_t off = roff[r];
        if (L < P.k || L - P.w <= 0) { if (threadIdx.x == 0) nhits[r] = 0; continue; }
        const int npow = next_pow2(L);
        const int bitsL = 32 - __clz(npow - 1 > 0 ? npow - 1 : 1);
        const uint64_t *pw = pack2 + off / 32; const uint32_t *nm = nmask + off / 32;
        if (P.w <= 1 && !P.hpc && 2 * P.k + bitsL < 32 && 2 * bitsL < 32 && npow <= SEED_SMEM_CAP) {
            uint32_t *buf = reinterpret_cast<uint32_t *>(sbuf), *hbuf = buf + npow; // 2 x npow x 4 B <= 128 KB
            const uint32_t posmask = (1u << bitsL) - 1;
            if (threadIdx.x == 0) s_cnt = 0;
            __syncthreads();
            // rolling 2-bit k-mer, 32 positions per thread, k-1 bases of warm-up (tandem_hit.c:37-56)
            for (int t = threadIdx.x; t * 32 < npow; t += blockDim.x) {
                const int base0 = t * 32;
                if (base0 >= L) { for (int p = base0; p < base0 + 32 && p < npow; ++p) buf[p] = 0xffffffffu; continue; }
                uint32_t key = 0; int l = 0, cnt = 0;
                int start = base0 - (P.k - 1); if (start < 0) start = 0;
                const int stop = base0 + 32 < L ? base0 + 32 : L;
                const uint64_t w0 = pw[start >> 5], w1 = pw[base0 >> 5];
                const uint32_t m0 = nm[start >> 5], m1 = nm[base0 >> 5];
                for (int p = start; p < stop; ++p) {
                    const bool cur = p >= base0;
                    const uint64_t wd = cur ? w1 : w0; const uint32_t md = cur ? m1 : m0;
                    uint32_t v = 0xffffffffu;
                    if ((md >> (p & 31)) & 1) { key = 0; l = 0; }
                    else {
                        key = ((key << 2) | (uint32_t)((wd >> (2 * (p & 31))) & 3)) & kmask;
                        if (++l >= P.k) { v = key << bitsL | (uint32_t)p; if (cur) ++cnt; }
                    }
                    if (cur) buf[p] = v;
                }
                for (int p = stop; p < base0 + 32 && p < npow; ++p) buf[p] = 0xffffffffu;
                if (cnt) atomicAdd(&s_cnt, cnt);
            }
            __syncthreads();
            const int n_seed = s_cnt;
            __syncthreads();
            if (n_seed == 0) { if (threadIdx.x == 0) nhits[r] = 0; continue; }
            block_bitonic_sort<false, uint32_t>(buf, npow);
            // nearest earlier occurrence of the same key at distance >= min_p (tandem_hit.c:186-214); hits are compacted
            if (threadIdx.x == 0) s_cnt = 0;
            __syncthreads();
            for (int j0 = 0; j0 < n_seed; j0 += blockDim.x) {
                const int j = j0 + threadIdx.x;
                uint32_t v = 0xffffffffu;
                if (j < n_seed) {
                    const uint32_t cur = buf[j], key = cur >> bitsL, pos = cur & posmask; uint32_t d = 0; bool found = false;
                    for (int kk = j - 1; kk >= 0; --kk) {
                        const uint32_t o = buf[kk];
                        if ((o >> bitsL) != key) break;
                        d = pos - (o & posmask);
                        if (d >= P.min_p) { found = true; break; }
                    }
                    if (found && d <= P.max_p) v = pos << bitsL | d;
                }
                const unsigned m = __ballot_sync(TH_FULL, v != 0xffffffffu);
                int wbase = 0;
                if ((threadIdx.x & 31) == 0 && m) wbase = atomicAdd(&s_cnt, __popc(m));
                wbase = __shfl_sync(TH_FULL, wbase, 0);
                if (v != 0xffffffffu) hbuf[wbase + __popc(m & ((1u << (threadIdx.x & 31)) - 1))] = v;
            }
            __syncthreads();
            const int n_hit = s_cnt;
            const int hpow = next_pow2(n_hit);
            for (int j = n_hit + threadIdx.x; j < hpow; j += blockDim.x) hbuf[j] = 0xffffffffu;
            __syncthreads();
            if (n_hit > 1) block_bitonic_sort<false, uint32_t>(hbuf, hpow);
            for (int j = threadIdx.x; j < n_hit; j += blockDim.x) {
                const uint32_t v = hbuf[j];
                hend[off + j] = (int32_t)(v >> bitsL);
                hper[off + j] = (int32_t)(v & posmask);
            }
            if (threadIdx.x == 0) nhits[r] = n_hit;
            __syncthreads();
            continue;
        }
        // ---- general path: 64-bit words, any option ----
        uint64_t *buf = npow <= SEED_SMEM_CAP ? sbuf : gbuf;
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
            if (threadIdx.x == 0) {
                int n = P.w > 1 ? seeds_minimizer(bseq + off, L, P.k, P.w, P.hpc, buf) : seeds_direct_hpc(bseq + off, L, P.k, buf);
                s_cnt = n;
            }
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
        uint64_t *hb = hpow <= SEED_SMEM_CAP ? sbuf : gtmp; // the seeds are no longer needed
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
