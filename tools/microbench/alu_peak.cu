// alu_peak.cu -- integer-pipe microbenchmark for the roofline denominators of the DP kernels ("lanes/clk and the
// s16x2 DPX rate must be microbenchmarked on the box").  One persistent grid (148 x 8 blocks x 256 threads), every thread runs
// ILP independent dependency chains of one operation; rate = warp-instructions per clock per SM at the measured SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_peak tools/microbench/alu_peak.cu && ./alu_peak
// Prints one JSON object.  Not part of the product; written in round 1, to be run first thing in round 2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ILP 8
#define ITERS 4096

enum Op { OP_VADD2, OP_VMAXS2, OP_VIMAX3, OP_VIBMAX, OP_VIADDMAX, OP_IADD, OP_IMAD, OP_LOP3, OP_PRMT, OP_SHFL, OP_REDUX, OP_COUNT };
static const char *op_name[OP_COUNT] = {"vadd2", "vmaxs2", "vimax3_s16x2", "vibmax_s16x2+2pred", "viaddmax_s16x2", "iadd", "imad", "lop3", "prmt", "shfl_up", "redux_max"};

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t a[ILP], b = seed | 0x00010001u, c = seed * 3u + 7u;
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = seed + threadIdx.x * 17u + i * 0x01010101u;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == OP_VADD2) a[i] = __vadd2(a[i], b);
            else if (OP == OP_VMAXS2) a[i] = __vmaxs2(a[i], b ^ a[(i + 1) % ILP]);
            else if (OP == OP_VIMAX3) a[i] = __vimax3_s16x2(a[i], b, c ^ it);
            else if (OP == OP_VIBMAX) { bool ph, pl; a[i] = __vibmax_s16x2(a[i], b + it, &ph, &pl); if (ph) b ^= 1u; if (pl) c ^= 2u; }
            else if (OP == OP_VIADDMAX) a[i] = __viaddmax_s16x2(a[i], b, c);
            else if (OP == OP_IADD) a[i] = a[i] + b + it;
            else if (OP == OP_IMAD) a[i] = a[i] * b + c;
            else if (OP == OP_LOP3) a[i] = (a[i] & b) ^ (c | it);
            else if (OP == OP_PRMT) a[i] = __byte_perm(a[i], b, 0x3254 + (it & 1));
            else if (OP == OP_SHFL) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1) + 1u;
            else if (OP == OP_REDUX) a[i] = __reduce_max_sync(0xffffffffu, a[i] + it);
        }
    }
    const long long t1 = clock64();
    uint32_t s = b ^ c;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(int n_sm, int blocks_per_sm, uint32_t *d_out, long long *d_cyc, double sm_mhz) {
    const int grid = n_sm * blocks_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<grid, 256>>>(d_out, 12345u, d_cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<grid, 256>>>(d_out, 54321u, d_cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inst = (double)grid * 8 /*warps*/ * ILP * ITERS;
    const double per_clk_sm = warp_inst / (ms * 1e-3 * sm_mhz * 1e6) / n_sm;
    printf("  {\"op\": \"%s\", \"ms\": %.3f, \"warp_inst_per_clk_per_sm\": %.3f, \"lane_ops_per_s\": %.4g, \"blocks_per_sm\": %d},\n",
           op_name[OP], ms, per_clk_sm, warp_inst * 32 / (ms * 1e-3), blocks_per_sm);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
    cudaDeviceProp p; if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double sm_mhz = clk_khz / 1e3;
    const int n_sm = p.multiProcessorCount, bps = 8;
    uint32_t *d_out; long long *d_cyc;
    cudaMalloc(&d_out, sizeof(uint32_t) * n_sm * bps * 256); cudaMalloc(&d_cyc, sizeof(long long) * n_sm * bps);
    printf("{\"device\": \"%s\", \"n_sm\": %d, \"sm_mhz_max\": %.0f, \"note\": \"rates assume the maximum SM clock; compare with nvidia-smi under load\", \"ops\": [\n", p.name, n_sm, sm_mhz);
    run<OP_VADD2>(n_sm, bps, d_out, d_cyc, sm_mhz); run<OP_VMAXS2>(n_sm, bps, d_out, d_cyc, sm_mhz); run<OP_VIMAX3>(n_sm, bps, d_out, d_cyc, sm_mhz);
    run<OP_VIBMAX>(n_sm, bps, d_out, d_cyc, sm_mhz); run<OP_VIADDMAX>(n_sm, bps, d_out, d_cyc, sm_mhz); run<OP_IADD>(n_sm, bps, d_out, d_cyc, sm_mhz);
    run<OP_IMAD>(n_sm, bps, d_out, d_cyc, sm_mhz); run<OP_LOP3>(n_sm, bps, d_out, d_cyc, sm_mhz); run<OP_PRMT>(n_sm, bps, d_out, d_cyc, sm_mhz);
    run<OP_SHFL>(n_sm, bps, d_out, d_cyc, sm_mhz); run<OP_REDUX>(n_sm, bps, d_out, d_cyc, sm_mhz);
    printf("  {\"op\": \"end\"}\n]}\n");
    cudaFree(d_out); cudaFree(d_cyc);
    return 0;
}
