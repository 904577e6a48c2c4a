// th_common.cuh -- shared device helpers for the TideHunter hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define TH_WARP 32
#define TH_FULL 0xffffffffu

// error codes reported per read / per task
enum { TH_OK = 0, TH_ERR_ARENA = 1, TH_ERR_BAND = 2, TH_ERR_BACKTRACK = 3, TH_ERR_CAP = 4, TH_ERR_LEN = 5 };

struct DevParams {
    int k, w, hpc, min_copy;
    uint32_t min_p, max_p;
    double max_div;
    int match, mismatch, o1, e1, o2, e2; // o2/e2 as used in the recurrences (affine mode: o1 + 1 / e1, see th_gpu_create)
    int affine;                          // abPOA's affine gap mode (gap_open2 == 0): own recurrences, see th_poa.cuh
    int o2_raw, e2_raw;                  // the caller's values: abPOA derives inf_min and the int16 range check from them in every gap mode
    int pn;          // int16 lanes of the emulated abPOA vector (16)
    int only_unit;
    int linear;                          // abPOA's linear gap mode (gap_open1 == 0): every consensus task runs on the wide path (poa_dp_linear)
    // derived once on the host (dev_params_from): the POA kernel reads them straight from the constant bank
    int lp;                              // log2(pn)
    int mat_abs, mis_abs, oe1, oe2, inf_min; // |match|, |mismatch|, o + e, abPOA's int16 "minus infinity" (simd_abpoa_align.c:1613-1614)
    uint32_t INFP, NOE1P, NOE2P, NE1P, NE2P, PE12, NEGMIS2, XMM; // the same as s16x2 pairs: (inf, inf), (-oe1, -oe1), ..., (-e1, -e2), (-mis, -mis), mat ^ -mis
};

// per-read counters, one array each (index = TC_* x n_reads + read)
enum { TC_TASKS = 0, TC_UNITS, TC_POS, TC_PAIR3, TC_LEFT, TC_CONS, TC_N };

struct TaskTotals {
    int32_t n[TC_N];            // sums of the per-read counters
    int32_t max_key;            // largest task size key (ncap x n_seqs, saturated): scales the size buckets of the task order
    int32_t retry_n;            // tasks the first POA pass handed to the second one
    int32_t max_ext, pad0;      // longest extension target (scales the length classes that pair extensions of similar size)
    unsigned long long slab_typ, slab_full; // largest typical slab of all tasks; largest full-width slab of the retried ones
    long long n_hits, dense_bound;          // hits of all reads; bound used for the dense consensus copy
    long long cons_total;                   // dense consensus length (after the POA)
    unsigned long long slab_wide;           // largest full-width slab any task of the chunk could ask of the second pass
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---------------------------------------------------------------------------------------------
// Block-wide bitonic sort of n (power of two) 64-bit keys; `a` may live in shared or global memory.
// Every key is distinct in all our uses (position / index in the low bits), so the result is the
// unique sorted order and any correct sort is bit-exact with the reference's radix sort
// (src/ksort.h:101-151) and its stable qsort (src/tandem_chain.c:21-43).
// ---------------------------------------------------------------------------------------------
template <bool DESC, class K = uint64_t>
__device__ void block_bitonic_sort(K *a, int n) {
    const int half = n >> 1;
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int p = threadIdx.x; p < half; p += blockDim.x) {
                int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                int x = i | j;
                bool up = ((i & k) == 0) != DESC;
                K ai = a[i], ax = a[x];
                if ((ai > ax) == up) { a[i] = ax; a[x] = ai; }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---------------------------------------------------------------------------------------------
// packed signed 16x2 helpers (DPX / video SIMD; all map to single SASS ops on sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pk(int lo, int hi) { return ((uint32_t)(uint16_t)lo) | ((uint32_t)(uint16_t)hi << 16); }
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(int16_t)(v & 0xffff); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)(int16_t)(v >> 16); }
