// th_api.cu -- C ABI (include/th_gpu.h) and host-side batch orchestration for the TideHunter hot path.
//
// One context per GPU.  A chunk of reads is staged in pinned memory (reads at 64-base aligned
// offsets, gaps filled with 'N'), copied to the device, and pushed through
//   pack -> seed -> chain DP -> rank -> chain select -> partition -> [host: split runs into tasks]
//        -> POA consensus -> ksw identity/extension items -> gather -> D2H
// on the context's stream.  The only host work between kernels is bookkeeping (prefix sums, splitting
// par_pos runs at -1 exactly as seqs_msa does, src/gen_cons.c:191-200, and sizing per-warp slabs).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include <limits.h>
#include <string>
#include "../../include/th_gpu.h"
#include "th_common.cuh"
#include "th_seed.cuh"
#include "th_chain.cuh"
#include "th_ksw.cuh"
#include "th_partition.cuh"
#include "th_poa.cuh"
#include "th_tasks.cuh"

static thread_local std::string g_err;
static void set_err(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
#include <stdarg.h>
static void set_err(const char *fmt, ...) {
    char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_err = buf;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return -1; } } while (0)
#define CKP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return nullptr; } } while (0)

struct DBuf { // growable device buffer
    void *p = nullptr; size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { set_err("cudaMalloc(%zu): %s", want, cudaGetErrorString(e)); return -1; }
        cap = want; return 0;
    }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct HBuf { // growable pinned host buffer
    void *p = nullptr; size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) { set_err("cudaMallocHost(%zu): %s", want, cudaGetErrorString(e)); return -1; }
        cap = want; return 0;
    }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct th_gpu_ctx {
    int device = 0, n_sm = 148;
    double share = 1.0; // fraction of the SM slots the persistent grids of this context claim (TH_GPU_SHARE; several contexts can then co-run)
    th_gpu_params params;
    DevParams dp;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[16];
    cudaEvent_t mark[4]; // th_gpu_mark: step brackets for callers that time several contexts together
    cudaEvent_t ev_block; // host waits go through a blocking-sync event: a waiting lane thread sleeps instead of spinning on a core
    // resident chunk
    int n_reads = 0; int64_t bpad = 0; int max_len = 0;
    std::vector<int64_t> h_roff; std::vector<int32_t> h_rlen;
    HBuf h_ascii;
    DBuf d_ascii, d_bseq, d_pack, d_nmask, d_roff, d_rlen;
    DBuf d_hend, d_hper, d_nhits, d_score, d_from, d_gflag, d_rank, d_nrank, d_tracked;
    DBuf d_choff, d_chlen, d_chscore, d_chidx, d_cells, d_pchn, d_pchoff, d_pchlen;
    DBuf d_par, d_paroff, d_parn, d_rstatus, d_scratch, d_scratch2, d_bnd, d_rev, d_counters;
    DBuf d_parstream, d_parused, d_pardoff;
    DBuf d_kswalpha; // ksw_pair_kernel's band-width estimate (one float), kept across chunks
    DBuf d_tasks, d_torder, d_ustart, d_ulen, d_slabs, d_consb, d_consc, d_consl, d_tstatus, d_items, d_iden, d_ext;
    DBuf d_gsrc, d_gdst, d_glen, d_dense_b, d_dense_c, d_redo;
    DBuf d_tcounts, d_totals, d_tkey, d_ekey, d_eorder, d_left, d_pos, d_rtoff, d_tposoff, d_tnseqs, d_tconsoff, d_retry, d_rorder;
    // host result storage (pinned: the final copies are asynchronous and the host waits once)
    HBuf h_totals, h_rtoff, h_tposoff, h_pos, h_tnseqs, h_tconsoff, h_tstatus, h_iden, h_ext, h_rstatus, h_counters, h_consb, h_consc;
    th_gpu_stats stats;
};

extern "C" void th_gpu_default_params(th_gpu_params *p) {
    memset(p, 0, sizeof(*p));
    p->k = 8; p->w = 1; p->hpc = 0; p->min_copy = 2; p->max_div = 0.25; p->min_p = 30; p->max_p = 10000;
    p->match = 2; p->mismatch = 4; p->gap_open1 = 4; p->gap_open2 = 24; p->gap_ext1 = 2; p->gap_ext2 = 1;
    p->only_unit = 0; p->need_cov = 0; p->simd_lanes16 = 16;
}
extern "C" const char *th_gpu_last_error(void) { return g_err.c_str(); }
extern "C" int th_gpu_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }

// the kernels' view of the options (pure host code: no device is touched, so the mapping is testable without a GPU)
static DevParams dev_params_from(const th_gpu_params *p) {
    DevParams d; memset(&d, 0, sizeof(d));
    d.k = p->k; d.w = p->w; d.hpc = p->hpc; d.min_copy = p->min_copy; d.min_p = (uint32_t)p->min_p; d.max_p = (uint32_t)p->max_p;
    d.max_div = p->max_div; d.match = p->match; d.mismatch = p->mismatch; d.o1 = p->gap_open1; d.e1 = p->gap_ext1; d.o2 = p->gap_open2; d.e2 = p->gap_ext2;
    d.o2_raw = p->gap_open2; d.e2_raw = p->gap_ext2;
    // abPOA's affine mode (gap_open2 == 0, abpoa_align.c:85-88) has its own recurrences (template parameter AFFINE of
    // poa_add_sequence); the unused second gap function only has to stay inside the int16 headroom checks
    d.linear = p->gap_open1 == 0;        // abpoa_set_gap_mode (abpoa_align.c:85-89): linear wins over affine
    d.affine = !d.linear && p->gap_open2 == 0;
    if (d.affine) { d.o2 = d.o1 + 1; d.e2 = d.e1; }
    d.pn = p->simd_lanes16; d.only_unit = p->only_unit;
    d.lp = d.pn == 16 ? 4 : 3;
    d.mat_abs = d.match < 0 ? -d.match : d.match; d.mis_abs = d.mismatch < 0 ? -d.mismatch : d.mismatch;
    d.oe1 = d.o1 + d.e1; d.oe2 = d.o2 + d.e2;
    { // simd_abpoa_align.c:1613-1614 uses the option values as given, whatever the gap mode
        const int a = -32768 + d.mis_abs, b = -32768 + d.oe1, c2 = -32768 + d.o2_raw + d.e2_raw;
        d.inf_min = std::max(std::max(a, b), c2) + 31 * std::max(d.e1, d.e2_raw);
    }
    auto pk2 = [](int lo, int hi) { return ((uint32_t)(uint16_t)lo) | ((uint32_t)(uint16_t)hi << 16); };
    d.INFP = pk2(d.inf_min, d.inf_min); d.NOE1P = pk2(-d.oe1, -d.oe1); d.NOE2P = pk2(-d.oe2, -d.oe2);
    d.NE1P = pk2(-d.e1, -d.e1); d.NE2P = pk2(-d.e2, -d.e2); d.PE12 = pk2(-d.e1, -d.e2);
    d.NEGMIS2 = pk2(-d.mis_abs, -d.mis_abs); d.XMM = ((uint32_t)(uint16_t)d.mat_abs ^ (uint32_t)(uint16_t)(-d.mis_abs)) * 0x10001u;
    return d;
}
// test hook: the fields of DevParams as int32 in declaration order (max_div as round(max_div * 1e6)); returns the count
extern "C" int th_gpu_debug_dev_params(const th_gpu_params *p, int32_t cap, int32_t *out) {
    const DevParams d = dev_params_from(p);
    const int32_t v[] = {d.k, d.w, d.hpc, d.min_copy, (int32_t)d.min_p, (int32_t)d.max_p, (int32_t)(d.max_div * 1e6 + 0.5), d.match, d.mismatch,
                         d.o1, d.e1, d.o2, d.e2, d.affine, d.o2_raw, d.e2_raw, d.pn, d.only_unit, d.linear};
    const int n = (int)(sizeof(v) / sizeof(v[0]));
    for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
    return n;
}

extern "C" th_gpu_ctx *th_gpu_create(const th_gpu_params *p, int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { set_err("no CUDA device available (there is no CPU fallback)"); return nullptr; }
    if (device < 0 || device >= n) { set_err("device %d out of range (0..%d)", device, n - 1); return nullptr; }
    if (p->k < 2 || p->k > 16) { set_err("k must be in 2..16"); return nullptr; }
    if (p->w < 1 || p->w > 255) { set_err("w must be in 1..255"); return nullptr; }
    if (p->simd_lanes16 != 16 && p->simd_lanes16 != 8) { set_err("simd_lanes16 must be 8 or 16"); return nullptr; }
    if (p->gap_open1 < 0 || p->gap_open2 < 0 || p->gap_ext1 <= 0 || p->gap_ext2 < 0) { set_err("gap penalties must not be negative (and the first extension penalty positive)"); return nullptr; }
    CKP(cudaSetDevice(device));
    th_gpu_ctx *c = new th_gpu_ctx();
    c->device = device; c->params = *p;
    cudaDeviceProp prop; CKP(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    if (const char *e = getenv("TH_GPU_SHARE")) { const double v = atof(e); if (v > 0.05 && v <= 1.0) c->share = v; }
    c->dp = dev_params_from(p);
    CKP(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 16; ++i) CKP(cudaEventCreate(&c->ev[i]));
    for (int i = 0; i < 4; ++i) CKP(cudaEventCreate(&c->mark[i]));
    CKP(cudaEventCreateWithFlags(&c->ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
    CKP(cudaFuncSetAttribute(seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEED_SMEM_BYTES));
    CKP(cudaFuncSetAttribute(poa_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(PoaSmem<16>) * POA_WARPS * 2 + sizeof(PoaLaneK) * 16)));
    CKP(cudaFuncSetAttribute(poa_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(PoaSmem<32>) * POA_WARPS + sizeof(PoaLaneK) * 32)));
    memset(&c->stats, 0, sizeof(c->stats));
    { const float a0 = KSW_BAND_ALPHA0;
      if (c->d_kswalpha.ensure(sizeof(float)) || cudaMemcpy(c->d_kswalpha.p, &a0, sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { set_err("cannot allocate device memory"); th_gpu_destroy(c); return nullptr; } }
    return c;
}

extern "C" void th_gpu_destroy(th_gpu_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DBuf *ds[] = {&c->d_ascii, &c->d_bseq, &c->d_pack, &c->d_nmask, &c->d_roff, &c->d_rlen, &c->d_hend, &c->d_hper, &c->d_nhits, &c->d_score, &c->d_from,
                  &c->d_gflag, &c->d_rank, &c->d_nrank, &c->d_tracked, &c->d_choff, &c->d_chlen, &c->d_chscore, &c->d_chidx, &c->d_cells, &c->d_pchn, &c->d_pchoff,
                  &c->d_pchlen, &c->d_par, &c->d_paroff, &c->d_parn, &c->d_rstatus, &c->d_scratch, &c->d_scratch2, &c->d_bnd, &c->d_rev, &c->d_counters,
                  &c->d_parstream, &c->d_parused, &c->d_pardoff, &c->d_tasks, &c->d_torder, &c->d_ustart, &c->d_ulen, &c->d_slabs, &c->d_consb, &c->d_consc,
                  &c->d_consl, &c->d_tstatus, &c->d_items, &c->d_iden, &c->d_ext, &c->d_gsrc, &c->d_gdst, &c->d_glen, &c->d_dense_b, &c->d_dense_c, &c->d_redo,
                  &c->d_tcounts, &c->d_totals, &c->d_tkey, &c->d_ekey, &c->d_eorder, &c->d_left, &c->d_pos, &c->d_rtoff, &c->d_tposoff, &c->d_tnseqs, &c->d_tconsoff, &c->d_retry, &c->d_rorder, &c->d_kswalpha};
    for (DBuf *b : ds) b->release();
    HBuf *hs[] = {&c->h_ascii, &c->h_totals, &c->h_rtoff, &c->h_tposoff, &c->h_pos, &c->h_tnseqs, &c->h_tconsoff, &c->h_tstatus, &c->h_iden, &c->h_ext, &c->h_rstatus,
                  &c->h_counters, &c->h_consb, &c->h_consc};
    for (HBuf *b : hs) b->release();
    for (int i = 0; i < 16; ++i) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(c->mark[i]);
    cudaEventDestroy(c->ev_block);
    cudaStreamDestroy(c->stream);
    delete c;
}

// ---------------------------------------------------------------------------------------------
// Wait for the context's stream without burning a host core: with several contexts per GPU and one process per GPU,
// spinning waits (the runtime's default on a lightly loaded host) starve the threads that format records.
static cudaError_t th_wait(th_gpu_ctx *c) {
    cudaError_t e = cudaEventRecord(c->ev_block, c->stream);
    return e != cudaSuccess ? e : cudaEventSynchronize(c->ev_block);
}

extern "C" int th_gpu_upload(th_gpu_ctx *c, int32_t n_reads, const char *const *seq, const int32_t *seq_len) {
    CK(cudaSetDevice(c->device));
    c->n_reads = n_reads;
    c->h_roff.assign(n_reads + 1, 0); c->h_rlen.assign(seq_len, seq_len + n_reads);
    int64_t off = 0; int max_len = 0;
    for (int i = 0; i < n_reads; ++i) {
        c->h_roff[i] = off;
        off += (((int64_t)seq_len[i] + 63) & ~63ll) + 64; // 64-aligned start, >= 64 bases of slack
        max_len = std::max(max_len, seq_len[i]);
    }
    c->h_roff[n_reads] = off; c->bpad = off; c->max_len = max_len;
    if (c->h_ascii.ensure((size_t)off + 64)) return -1;
    char *h = c->h_ascii.as<char>();
    for (int i = 0; i < n_reads; ++i) {
        memcpy(h + c->h_roff[i], seq[i], seq_len[i]);
        memset(h + c->h_roff[i] + seq_len[i], 'N', (size_t)(c->h_roff[i + 1] - c->h_roff[i] - seq_len[i]));
    }
    if (c->d_ascii.ensure((size_t)off + 64) || c->d_roff.ensure(sizeof(int64_t) * (n_reads + 1)) || c->d_rlen.ensure(sizeof(int32_t) * (n_reads + 1))) return -1;
    CK(cudaEventRecord(c->ev[0], c->stream));
    CK(cudaMemcpyAsync(c->d_ascii.p, h, (size_t)off, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_roff.p, c->h_roff.data(), sizeof(int64_t) * (n_reads + 1), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_rlen.p, c->h_rlen.data(), sizeof(int32_t) * n_reads, cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(c->ev[1], c->stream));
    CK(th_wait(c));
    float ms = 0; cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->stats.ms_h2d = ms;
    c->stats.h2d_bytes = off + (int64_t)(sizeof(int64_t) + sizeof(int32_t)) * n_reads;
    return 0;
}

__global__ void ksw_test_kernel(int n, int mode, const uint8_t *__restrict__ buf, const int64_t *__restrict__ qoff, const int32_t *__restrict__ ql,
                                const int64_t *__restrict__ toff, const int32_t *__restrict__ tl, const int32_t *__restrict__ arg,
                                int4 *bnd_all, int64_t bnd_stride, int32_t *__restrict__ out2) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    int4 *bnd = bnd_all + (int64_t)w * bnd_stride;
    int o0 = 0, o1 = 0;
    if (mode == 0) ksw_warp<KSW_GLOBAL, 16>(buf + qoff[w], ql[w], buf + toff[w], tl[w], 0, bnd, o0, o1);
    else if (mode == 1) ksw_warp<KSW_GLOBAL_STOP, 4>(buf + qoff[w], ql[w], buf + toff[w], tl[w], ql[w] - arg[w], bnd, o0, o1);
    else if (mode == 2) ksw_warp<KSW_EXT, 16>(buf + qoff[w], ql[w], buf + toff[w], tl[w], 0, bnd, o0, o1);
    else if (mode == 4) { // entries 2k and 2k + 1 as the two halves of one packed identity alignment, exactly as ksw_pair_kernel runs it;
        // arg = first band half-width in thousandths of the length (0: the kernel's default); out = (identity, path taken)
        if (w & 1) return;
        const int v = w + 1 < n ? w + 1 : w;
        float alpha = arg[w] > 0 ? 0.001f * (float)arg[w] : KSW_BAND_ALPHA0;
        int a = 0, b = 0, path = 0; unsigned long long nc = 0;
        ksw_pair_identity<KSW2_C>(buf + qoff[w], ql[w], buf + toff[w], tl[w], buf + qoff[v], ql[v], buf + toff[v], tl[v], bnd, alpha, a, b, nc, &path);
        if (lane == 0) { out2[2 * w] = a; out2[2 * w + 1] = path; if (v != w) { out2[2 * v] = b; out2[2 * v + 1] = path; } }
        return;
    }
    else { // mode 3: entries 2k and 2k + 1 (queries of equal length) as the two halves of one packed extension
        if (w & 1) return;
        int a0 = -1, a1 = -1, b0 = -1, b1 = -1;
        if (w + 1 < n) ksw_warp_ext2<16>(buf + qoff[w], ql[w], buf + toff[w], tl[w], buf + qoff[w + 1], ql[w + 1], buf + toff[w + 1], tl[w + 1], bnd, a0, a1, b0, b1);
        else ksw_warp<KSW_EXT, 16>(buf + qoff[w], ql[w], buf + toff[w], tl[w], 0, bnd, a0, a1);
        if (lane == 0) { out2[2 * w] = a0; out2[2 * w + 1] = a1; if (w + 1 < n) { out2[2 * w + 2] = b0; out2[2 * w + 3] = b1; } }
        return;
    }
    if (lane == 0) { out2[2 * w] = o0; out2[2 * w + 1] = o1; }
}

static float ev_ms(th_gpu_ctx *c, int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]); return ms; }

extern "C" int th_gpu_process_resident(th_gpu_ctx *c, th_gpu_result *out) {
    CK(cudaSetDevice(c->device));
    const int n = c->n_reads; const int64_t B = c->bpad; cudaStream_t st = c->stream;
    const DevParams P = c->dp;
    th_gpu_stats &S = c->stats;
    float keep_h2d = S.ms_h2d; int64_t keep_h2db = S.h2d_bytes;
    memset(&S, 0, sizeof(S)); S.ms_h2d = keep_h2d; S.h2d_bytes = keep_h2db;
    memset(out, 0, sizeof(*out));
    static const int32_t zero2[2] = {0, 0};
    if (n == 0) { out->read_task_off = zero2; out->task_pos_off = zero2; out->task_cons_off = zero2; return 0; }
    for (int i = 0; i < n; ++i) S.n_bases += c->h_rlen[i];
    const size_t B4 = (size_t)B * 4;
    if (c->d_bseq.ensure(B + 64) || c->d_pack.ensure(B / 4 + 64) || c->d_nmask.ensure(B / 8 + 64) ||
        c->d_hend.ensure(B4) || c->d_hper.ensure(B4) || c->d_nhits.ensure(4 * (size_t)n) || c->d_score.ensure(B4) || c->d_from.ensure(B4) ||
        c->d_gflag.ensure(4 * (size_t)n) || c->d_rank.ensure(B4) || c->d_nrank.ensure(4 * (size_t)n) || c->d_tracked.ensure(B) ||
        c->d_choff.ensure(B4 / 2 + 256) || c->d_chlen.ensure(B4 / 2 + 256) || c->d_chscore.ensure(B4 / 2 + 256) || c->d_chidx.ensure(B4 / 2 + 256) ||
        c->d_cells.ensure(B4) || c->d_pchn.ensure(4 * (size_t)n) || c->d_pchoff.ensure(B4 / 2 + 256) || c->d_pchlen.ensure(B4 / 2 + 256) ||
        c->d_par.ensure(2 * B4) || c->d_paroff.ensure(B4 / 2 + 256) || c->d_parn.ensure(B4 / 2 + 256) || c->d_rstatus.ensure(4 * (size_t)n) ||
        c->d_counters.ensure(256) || c->d_tcounts.ensure(4 * (size_t)TC_N * n + 64) || c->d_totals.ensure(sizeof(TaskTotals)) ||
        c->d_rtoff.ensure(4 * (size_t)(n + 1)) || c->h_totals.ensure(sizeof(TaskTotals)) || c->h_rtoff.ensure(4 * (size_t)(n + 1)) ||
        c->h_rstatus.ensure(4 * (size_t)n) || c->h_counters.ensure(256))
        return -1;
    // counters: [0] chain evals (u64), [1] poa cells, [2] poa rows, [3] ksw cells, [8..] int work counters
    CK(cudaMemsetAsync(c->d_counters.p, 0, 256, st));
    CK(cudaMemsetAsync(c->d_rstatus.p, 0, 4 * (size_t)n, st));
    CK(cudaMemsetAsync(c->d_totals.p, 0, sizeof(TaskTotals), st));
    unsigned long long *cnt64 = c->d_counters.as<unsigned long long>();
    int *cnt32 = c->d_counters.as<int>() + 16;
    TaskTotals *d_tot = c->d_totals.as<TaskTotals>();
    const int64_t *roff = c->d_roff.as<int64_t>(); const int32_t *rlen = c->d_rlen.as<int32_t>();
    int ei = 2; // event index
    // ---- pack ----
    CK(cudaEventRecord(c->ev[ei++], st)); // 2
    { const int64_t nw = B / 32; pack_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(c->d_ascii.as<uint8_t>(), c->d_bseq.as<uint8_t>(), c->d_pack.as<uint64_t>(), c->d_nmask.as<uint32_t>(), nw); S.n_launches++; }
    CK(cudaEventRecord(c->ev[ei++], st)); // 3
    // ---- seed ----
    {
        int gcap = 1; while (gcap < c->max_len) gcap <<= 1;
        // shared memory per block: the default path needs two 32-bit buffers of the longest read plus the radix counters;
        // up to ~113 KB two blocks share an SM (reads up to ~12 k bases), beyond that one block with the full 144 KB
        int smem = SEED_SMEM_BYTES, per_sm = 1;
        if (P.w <= 1 && !P.hpc) {
            const int need = 8 * ((c->max_len + 31) & ~31) + SEED_HIST_BYTES;
            if (need <= 113 * 1024) { smem = need; per_sm = 2; }
        }
        const int grid = std::min(n, c->n_sm * per_sm);
        if (c->d_scratch.ensure((size_t)grid * 2 * gcap * 8)) return -1;
        seed_kernel<<<grid, SEED_THREADS, smem, st>>>(P, n, smem, roff, rlen, c->d_bseq.as<uint8_t>(), c->d_pack.as<uint64_t>(), c->d_nmask.as<uint32_t>(),
                                                      c->d_scratch.as<uint64_t>(), gcap, c->d_hend.as<int32_t>(), c->d_hper.as<int32_t>(), c->d_nhits.as<int32_t>());
        S.n_launches++;
    }
    CK(cudaEventRecord(c->ev[ei++], st)); // 4
    // ---- chain DP ----
    {
        const int grid = std::min((n + CHAIN_WARPS - 1) / CHAIN_WARPS, std::max(1, (int)(c->n_sm * CHAIN_BLOCKS_PER_SM * c->share)));
        if (c->d_rorder.ensure(4 * (size_t)n + 64)) return -1;
        bucket_order_kernel<<<1, 1024, 0, st>>>(nullptr, 0, nullptr, c->d_nhits.as<int32_t>(), c->d_rorder.as<int32_t>(), n, c->max_len); // reads by hit count, most first
        // period buckets of the chaining DP: four chunk-wide int arrays that are only written later (chain cells, par_pos x 2, ranks)
        if (P.max_p < (1u << 27)) // hit periods are <= max_p (src/tandem_hit.c:204): the 1.8x gate fits 32-bit products
            chain_dp_kernel<true><<<grid, CHAIN_WARPS * 32, 0, st>>>(P, n, roff, c->d_nhits.as<int32_t>(), c->d_hend.as<int32_t>(), c->d_hper.as<int32_t>(),
                                                                  c->d_score.as<int32_t>(), c->d_from.as<int32_t>(), c->d_gflag.as<int32_t>(), cnt64 + 0,
                                                                  c->d_rorder.as<int32_t>(), cnt32 + 7,
                                                                  c->d_cells.as<int32_t>(), c->d_par.as<int32_t>(), c->d_par.as<int32_t>() + B, c->d_rank.as<int32_t>());
        else
            chain_dp_kernel<false><<<grid, CHAIN_WARPS * 32, 0, st>>>(P, n, roff, c->d_nhits.as<int32_t>(), c->d_hend.as<int32_t>(), c->d_hper.as<int32_t>(),
                                                                   c->d_score.as<int32_t>(), c->d_from.as<int32_t>(), c->d_gflag.as<int32_t>(), cnt64 + 0,
                                                                   c->d_rorder.as<int32_t>(), cnt32 + 7,
                                                                   c->d_cells.as<int32_t>(), c->d_par.as<int32_t>(), c->d_par.as<int32_t>() + B, c->d_rank.as<int32_t>());
        S.n_launches += 2;
        if (P.w > 1) { // repeated ends can only come from minimizer seeds
            chain_dp_generic_kernel<<<(n + 63) / 64, 64, 0, st>>>(P, n, roff, c->d_nhits.as<int32_t>(), c->d_hend.as<int32_t>(), c->d_hper.as<int32_t>(),
                                                                c->d_score.as<int32_t>(), c->d_from.as<int32_t>(), c->d_gflag.as<int32_t>(), c->d_rank.as<int32_t>());
            S.n_launches++;
        }
    }
    CK(cudaEventRecord(c->ev[ei++], st)); // 5
    // ---- rank + select ----
    {
        int gcap = 1; while (gcap < c->max_len) gcap <<= 1;
        const int grid = std::min(n, c->n_sm * 4);
        if (c->d_scratch2.ensure((size_t)grid * gcap * 8)) return -1;
        rank_kernel<<<grid, RANK_THREADS, 0, st>>>(n, roff, c->d_nhits.as<int32_t>(), c->d_hend.as<int32_t>(), c->d_score.as<int32_t>(), c->d_scratch2.as<uint64_t>(), gcap,
                                                  c->d_rank.as<int32_t>(), c->d_nrank.as<int32_t>(), c->d_rstatus.as<int32_t>());
        chain_select_kernel<<<(n + SELECT_WARPS - 1) / SELECT_WARPS, SELECT_WARPS * 32, 0, st>>>(n, roff, c->d_nhits.as<int32_t>(), c->d_hend.as<int32_t>(), c->d_hper.as<int32_t>(), c->d_score.as<int32_t>(),
                                                         c->d_from.as<int32_t>(), c->d_rank.as<int32_t>(), c->d_nrank.as<int32_t>(), c->d_tracked.as<uint8_t>(),
                                                         c->d_choff.as<int32_t>(), c->d_chlen.as<int32_t>(), c->d_chscore.as<int32_t>(), c->d_chidx.as<int32_t>(),
                                                         c->d_cells.as<int32_t>(), c->d_pchn.as<int32_t>(), c->d_pchoff.as<int32_t>(), c->d_pchlen.as<int32_t>(), 1);
        S.n_launches += 2;
    }
    CK(cudaEventRecord(c->ev[ei++], st)); // 6
    // ---- partition, then the tasks it implies are counted on the device ----
    const int n_pwarps = std::min(n, std::max(4, (int)(c->n_sm * PART_MIN_BLOCKS * PART_WARPS * c->share)));
    const int64_t bnd_stride = 2 * (int64_t)(c->max_len + 64);
    {
        const int grid = (n_pwarps + PART_WARPS - 1) / PART_WARPS;
        if (c->d_bnd.ensure((size_t)grid * PART_WARPS * bnd_stride * sizeof(int4))) return -1;
        partition_kernel<<<grid, PART_WARPS * 32, 0, st>>>(P, n, roff, rlen, c->d_bseq.as<uint8_t>(), c->d_hend.as<int32_t>(), c->d_hper.as<int32_t>(), c->d_cells.as<int32_t>(),
                                                          c->d_pchn.as<int32_t>(), c->d_pchoff.as<int32_t>(), c->d_pchlen.as<int32_t>(), c->d_par.as<int32_t>(),
                                                          c->d_paroff.as<int32_t>(), c->d_parn.as<int32_t>(), c->d_bnd.as<int4>(), bnd_stride, cnt32 + 0,
                                                          c->d_rstatus.as<int32_t>(), cnt64 + 3, c->d_rorder.as<int32_t>());
        task_count_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, c->params.min_copy, roff, rlen, c->d_pchn.as<int32_t>(), c->d_par.as<int32_t>(), c->d_paroff.as<int32_t>(),
                                                          c->d_parn.as<int32_t>(), c->d_nhits.as<int32_t>(), c->d_tcounts.as<int32_t>(), d_tot);
        task_scan_kernel<<<1, 1024, 0, st>>>(n, c->d_tcounts.as<int32_t>(), d_tot);
        S.n_launches += 3;
    }
    CK(cudaEventRecord(c->ev[ei++], st)); // 7
    // ---- first (and only mid-chunk) wait: the totals that size everything downstream ----
    TaskTotals *ht = c->h_totals.as<TaskTotals>();
    CK(cudaMemcpyAsync(ht, d_tot, sizeof(TaskTotals), cudaMemcpyDeviceToHost, st));
    CK(th_wait(c));
    const int nt = ht->n[TC_TASKS], n_units = ht->n[TC_UNITS], n_pos = ht->n[TC_POS];
    const int n_pairs = ht->n[TC_PAIR3] + (ht->n[TC_LEFT] >> 1), n_singles = ht->n[TC_LEFT] & 1, n_exts = P.only_unit ? 0 : 2 * nt;
    const int64_t cons_total = ht->n[TC_CONS];
    S.n_hits = ht->n_hits; S.n_tasks = nt;
    S.d2h_bytes += sizeof(TaskTotals);
    if (c->d_tasks.ensure(sizeof(PoaTask) * (size_t)nt + 64) || c->d_torder.ensure(4 * (size_t)nt + 64) || c->d_tkey.ensure(4 * (size_t)nt + 64) || c->d_ekey.ensure(8 * (size_t)nt + 64) || c->d_eorder.ensure(8 * (size_t)nt + 64) ||
        c->d_ustart.ensure(4 * (size_t)n_units + 64) || c->d_ulen.ensure(4 * (size_t)n_units + 64) || c->d_pos.ensure(4 * (size_t)n_pos + 64) ||
        c->d_tposoff.ensure(4 * (size_t)(nt + 1)) || c->d_tnseqs.ensure(4 * (size_t)nt + 64) || c->d_tconsoff.ensure(4 * (size_t)(nt + 1)) ||
        c->d_items.ensure(sizeof(KswItem) * ((size_t)n_pairs + n_singles + 2 * (size_t)nt) + 64) || c->d_left.ensure(sizeof(LeftUnit) * (size_t)ht->n[TC_LEFT] + 64) ||
        c->d_consb.ensure((size_t)cons_total + 64) || c->d_consc.ensure(4 * (size_t)cons_total + 64) || c->d_consl.ensure(4 * (size_t)nt + 64) ||
        c->d_tstatus.ensure(4 * (size_t)nt + 64) || c->d_iden.ensure(4 * (size_t)n_pos + 64) || c->d_ext.ensure(16 * (size_t)nt + 64) || c->d_retry.ensure(4 * (size_t)nt + 64) ||
        c->h_tposoff.ensure(4 * (size_t)(nt + 1)) || c->h_pos.ensure(4 * (size_t)n_pos + 64) || c->h_tnseqs.ensure(4 * (size_t)nt + 64) || c->h_tconsoff.ensure(4 * (size_t)(nt + 1)) ||
        c->h_tstatus.ensure(4 * (size_t)nt + 64) || c->h_iden.ensure(4 * (size_t)n_pos + 64) || c->h_ext.ensure(16 * (size_t)nt + 64)) return -1;
    task_fill_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, c->params.min_copy, P.only_unit, roff, rlen, c->d_pchn.as<int32_t>(), c->d_par.as<int32_t>(), c->d_paroff.as<int32_t>(),
                                                     c->d_parn.as<int32_t>(), c->d_tcounts.as<int32_t>(), d_tot, c->d_tasks.as<PoaTask>(), c->d_ustart.as<int32_t>(), c->d_ulen.as<int32_t>(),
                                                     c->d_pos.as<int32_t>(), c->d_rtoff.as<int32_t>(), c->d_tposoff.as<int32_t>(), c->d_tnseqs.as<int32_t>(), c->d_tkey.as<int32_t>(),
                                                     c->d_ekey.as<int32_t>(), c->d_items.as<KswItem>(), c->d_left.as<LeftUnit>());
    S.n_launches++;
    CK(cudaMemsetAsync(c->d_consl.p, 0, 4 * (size_t)nt + 64, st));
    CK(cudaMemsetAsync(c->d_tstatus.p, 0, 4 * (size_t)nt + 64, st));
    CK(cudaMemsetAsync(c->d_iden.p, 0, 4 * (size_t)n_pos + 64, st));
    CK(cudaMemsetAsync(c->d_ext.p, 0xff, 16 * (size_t)nt + 64, st));
    CK(cudaEventRecord(c->ev[ei++], st)); // 8
    long long dense_cap = 0;
    if (nt > 0 && !P.only_unit) {
        if (ht->n[TC_LEFT] > 0) { task_pair_left_kernel<<<(ht->n[TC_LEFT] / 2 + 128) / 128, 128, 0, st>>>(d_tot, c->d_left.as<LeftUnit>(), c->d_items.as<KswItem>()); S.n_launches++; }
        bucket_order_kernel<<<1, 1024, 0, st>>>(d_tot, 1, &d_tot->max_key, c->d_tkey.as<int32_t>(), c->d_torder.as<int32_t>());
        bucket_order_kernel<<<1, 1024, 0, st>>>(d_tot, 2, &d_tot->max_ext, c->d_ekey.as<int32_t>(), c->d_eorder.as<int32_t>());
        S.n_launches += 2;
        // ---- POA: 16-lane groups, two tasks per warp ----
        size_t slab_typ = ((size_t)ht->slab_typ + 255) & ~(size_t)255; // slabs hold 16-byte accesses
        if (slab_typ == 0) slab_typ = 1 << 20;
        constexpr int GPB16 = POA_WARPS * 2, GPB32 = POA_WARPS;
        double poa_blocks = POA_MIN_BLOCKS16;            // resident POA blocks per SM this context asks for (tuning: TH_POA_BLOCKS)
        if (const char *e = getenv("TH_POA_BLOCKS")) { const double v = atof(e); if (v >= 0.5 && v <= POA_MIN_BLOCKS16) poa_blocks = v; }
        const int want_groups = std::min((int)(c->n_sm * poa_blocks * GPB16 * c->share), std::max(nt, 1));
        // The memory budget is only consulted when the slabs have to grow: cudaMemGetInfo is a driver call that takes from a
        // fraction of a millisecond to 70 ms (measured: it added 13 ms per chunk on average, on the host, between the task
        // kernels and the POA launch), and chunks of one run are alike, so the slabs of the previous chunk nearly always do.
        size_t free_b = 0, total_b = 0, budget = c->d_slabs.cap;
        const bool fits = (size_t)((want_groups + GPB16 - 1) / GPB16) * GPB16 * slab_typ <= c->d_slabs.cap && (((size_t)ht->slab_wide + 511) & ~(size_t)255) <= c->d_slabs.cap;
        if (!fits) {
            CK(cudaMemGetInfo(&free_b, &total_b));
            budget = (size_t)((double)(free_b + c->d_slabs.cap) * 0.6);
            budget = std::min(budget, total_b / 5); // several contexts share the device (the host layer runs four): none may take it all
        }
        int ngroups = (int)std::min<size_t>((size_t)want_groups, std::max<size_t>(1, budget / slab_typ));
        int grid = (ngroups + GPB16 - 1) / GPB16;
        // several contexts of one process (the host layer's lanes) size their slabs from the same free-memory reading:
        // if the allocation loses that race, run with fewer resident groups instead of failing the chunk
        // the second pass reuses these slabs: room for at least one full-width slab of the chunk's largest task, so that no
        // task is left without a pass that can hold it (when even that exceeds the budget the task keeps its error status)
        const size_t wide_one = std::min<size_t>(((size_t)ht->slab_wide + 511) & ~(size_t)255, budget);
        while (c->d_slabs.ensure(std::max((size_t)grid * GPB16 * slab_typ, wide_one))) {
            cudaGetLastError();
            if (grid <= 4) return -1;
            grid = (grid + 1) / 2;
        }
        if (getenv("TH_GPU_DEBUG")) fprintf(stderr, "[th_gpu] POA: %d tasks on %d blocks (%d groups), slabs of %zu bytes, budget %.1f GB%s\n", nt, grid, grid * GPB16, slab_typ, budget / 1e9, fits ? " (slabs of the previous chunk reused, memory not queried)" : "");
        CK(cudaEventRecord(c->ev[12], st));   // 12, 13: the packed POA kernel alone (TH_GPU_DEBUG prints it beside the stage's bracket)
        poa_kernel<16><<<grid, POA_WARPS * 32, sizeof(PoaSmem<16>) * GPB16 + sizeof(PoaLaneK) * 16, st>>>(
            P, nt, nullptr, c->d_tasks.as<PoaTask>(), c->d_torder.as<int32_t>(), c->d_ustart.as<int32_t>(), c->d_ulen.as<int32_t>(), c->d_bseq.as<uint8_t>(),
            c->d_slabs.as<uint8_t>(), slab_typ, nullptr, c->d_slabs.cap, cnt32 + 1, c->d_consb.as<uint8_t>(), c->d_consc.as<int32_t>(), c->d_consl.as<int32_t>(),
            c->d_tstatus.as<int32_t>(), cnt64 + 1, cnt64 + 2, cnt64 + 16, grid * GPB16, c->d_retry.as<int32_t>(), d_tot);
        CK(cudaEventRecord(c->ev[13], st));
        // second pass, one task per warp with full-width slabs, driven from the device: tasks whose DP arena overflowed the
        // typical slab and rows with more than 16 predecessors were put on a list by the first pass, together with the
        // slab size they need.  It reuses the first pass's slabs, with as many resident warps as fit; when not even one
        // slab fits, the tasks keep their error status (the host layer reports and drops those records).  With an empty
        // list (the usual case) the launch ends at once.
        {
            const int rgrid = std::min(c->n_sm * POA_MIN_BLOCKS32, std::max(1, (nt + GPB32 - 1) / GPB32));
            poa_kernel<32><<<rgrid, POA_WARPS * 32, sizeof(PoaSmem<32>) * GPB32 + sizeof(PoaLaneK) * 32, st>>>(
                P, 0, &d_tot->retry_n, c->d_tasks.as<PoaTask>(), c->d_retry.as<int32_t>(), c->d_ustart.as<int32_t>(), c->d_ulen.as<int32_t>(), c->d_bseq.as<uint8_t>(),
                c->d_slabs.as<uint8_t>(), 0, &d_tot->slab_full, c->d_slabs.cap, cnt32 + 6, c->d_consb.as<uint8_t>(), c->d_consc.as<int32_t>(), c->d_consl.as<int32_t>(),
                c->d_tstatus.as<int32_t>(), cnt64 + 1, cnt64 + 2, cnt64 + 16, rgrid * GPB32, nullptr, d_tot);
        }
        S.n_launches += 2;
        CK(cudaEventRecord(c->ev[ei++], st)); // 9
        // ---- post-consensus ksw items: pairs, singles, extensions, one kernel each (own register budgets) ----
        {
            const int64_t rev_stride = 4 * (int64_t)(c->max_len + 64); // two (consensus + flank) pairs of reversed copies per warp
            const KswItem *d_pairs = c->d_items.as<KswItem>(), *d_singles = d_pairs + n_pairs, *d_exts = d_singles + n_singles;
            auto grid_for = [&](int n_work, int min_blocks) { const int kw = std::min(std::max(n_work, 1), std::max(4, (int)(c->n_sm * min_blocks * KSW_WARPS * c->share))); return (kw + KSW_WARPS - 1) / KSW_WARPS; };
            const int g_pair = grid_for(n_pairs, KSW_PAIR_MIN_BLOCKS), g_single = grid_for(n_singles + 64, KSW_MIN_BLOCKS), g_ext = grid_for((n_exts + 1) / 2, KSW_EXT_MIN_BLOCKS);
            const int g_max = std::max(g_pair, std::max(g_single, g_ext));
            if (c->d_bnd.ensure((size_t)g_max * KSW_WARPS * bnd_stride * sizeof(int4)) || c->d_rev.ensure((size_t)g_ext * KSW_WARPS * rev_stride) ||
                c->d_redo.ensure(4 * (size_t)n_pairs + 64) || c->d_glen.ensure(4 * (size_t)nt + 64)) return -1;
            task_cons_off_kernel<<<(nt + 255) / 256, 256, 0, st>>>(nt, c->d_tasks.as<PoaTask>(), c->d_glen.as<int32_t>());
            S.n_launches++;
            if (n_pairs > 0) {
                ksw_pair_kernel<<<g_pair, KSW_WARPS * 32, 0, st>>>(n_pairs, d_pairs, c->d_bseq.as<uint8_t>(), c->d_consb.as<uint8_t>(), c->d_glen.as<int32_t>(),
                                                                 c->d_consl.as<int32_t>(), c->d_bnd.as<int4>(), bnd_stride, cnt32 + 2, c->d_redo.as<int32_t>(), cnt32 + 5,
                                                                 c->d_iden.as<int32_t>(), cnt64 + 3, c->d_kswalpha.as<float>());
                S.n_launches++;
            }
            ksw_single_kernel<<<g_single, KSW_WARPS * 32, 0, st>>>(n_singles, d_singles, d_pairs, c->d_redo.as<int32_t>(), cnt32 + 5, c->d_bseq.as<uint8_t>(),
                                                                 c->d_consb.as<uint8_t>(), c->d_glen.as<int32_t>(), c->d_consl.as<int32_t>(), c->d_bnd.as<int4>(), bnd_stride,
                                                                 cnt32 + 3, c->d_iden.as<int32_t>(), cnt64 + 3);
            S.n_launches++;
            if (n_exts > 0) {
                ksw_ext_kernel<<<g_ext, KSW_WARPS * 32, 0, st>>>(n_exts, d_exts, c->d_eorder.as<int32_t>(), c->d_bseq.as<uint8_t>(), c->d_consb.as<uint8_t>(), c->d_glen.as<int32_t>(),
                                                               c->d_consl.as<int32_t>(), c->d_rev.as<uint8_t>(), rev_stride, c->d_bnd.as<int4>(), bnd_stride, cnt32 + 4,
                                                               c->d_ext.as<int32_t>(), cnt64 + 3);
                S.n_launches++;
            }
        }
        CK(cudaEventRecord(c->ev[ei++], st)); // 10
        // ---- dense consensus: offsets by a scan over the tasks, one gather ----
        dense_cap = std::min<long long>(ht->dense_bound, cons_total) + 64;
        if (c->d_dense_b.ensure((size_t)dense_cap + 64) || c->h_consb.ensure((size_t)dense_cap + 64) ||
            (c->params.need_cov && (c->d_dense_c.ensure(4 * (size_t)dense_cap + 64) || c->h_consc.ensure(4 * (size_t)dense_cap + 64)))) return -1;
        cons_scan_kernel<<<1, 1024, 0, st>>>(d_tot, c->d_consl.as<int32_t>(), c->d_tconsoff.as<int32_t>());
        cons_gather_kernel<<<(nt * 32 + 255) / 256, 256, 0, st>>>(d_tot, c->d_tasks.as<PoaTask>(), c->d_consl.as<int32_t>(), c->d_tconsoff.as<int32_t>(), c->d_consb.as<uint8_t>(),
                                                                c->d_consc.as<int32_t>(), c->d_dense_b.as<uint8_t>(), c->params.need_cov ? c->d_dense_c.as<int32_t>() : nullptr, dense_cap);
        S.n_launches += 2;
        CK(cudaMemcpyAsync(c->h_consb.p, c->d_dense_b.p, (size_t)dense_cap, cudaMemcpyDeviceToHost, st));
        if (c->params.need_cov) CK(cudaMemcpyAsync(c->h_consc.p, c->d_dense_c.p, 4 * (size_t)dense_cap, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(c->h_tconsoff.p, c->d_tconsoff.p, 4 * (size_t)(nt + 1), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(c->h_iden.p, c->d_iden.p, 4 * (size_t)n_pos, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(c->h_ext.p, c->d_ext.p, 16 * (size_t)nt, cudaMemcpyDeviceToHost, st));
    } else {
        CK(cudaEventRecord(c->ev[ei++], st)); CK(cudaEventRecord(c->ev[ei++], st)); // 9, 10
        CK(cudaMemsetAsync(c->d_tconsoff.p, 0, 4 * (size_t)(nt + 1), st));
        CK(cudaMemcpyAsync(c->h_tconsoff.p, c->d_tconsoff.p, 4 * (size_t)(nt + 1), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(c->h_iden.p, c->d_iden.p, 4 * (size_t)n_pos, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(c->h_ext.p, c->d_ext.p, 16 * (size_t)nt, cudaMemcpyDeviceToHost, st));
    }
    // ---- everything else the host layer reads, then the second and last wait ----
    CK(cudaMemcpyAsync(c->h_rtoff.p, c->d_rtoff.p, 4 * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(c->h_tposoff.p, c->d_tposoff.p, 4 * (size_t)(nt + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(c->h_pos.p, c->d_pos.p, 4 * (size_t)n_pos, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(c->h_tnseqs.p, c->d_tnseqs.p, 4 * (size_t)nt, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(c->h_tstatus.p, c->d_tstatus.p, 4 * (size_t)nt, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(c->h_rstatus.p, c->d_rstatus.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(c->h_counters.p, c->d_counters.p, 256, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ht, d_tot, sizeof(TaskTotals), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(c->ev[ei++], st)); // 11
    CK(th_wait(c));
    int32_t *h_rtoff = c->h_rtoff.as<int32_t>(), *h_tstatus = c->h_tstatus.as<int32_t>();
    if (nt > 0 && !P.only_unit && ht->cons_total + 64 > dense_cap) { // consensus longer than the units it came from: fetch the rest with an exact copy (never seen)
        dense_cap = ht->cons_total + 64;
        if (c->d_dense_b.ensure((size_t)dense_cap + 64) || c->h_consb.ensure((size_t)dense_cap + 64) ||
            (c->params.need_cov && (c->d_dense_c.ensure(4 * (size_t)dense_cap + 64) || c->h_consc.ensure(4 * (size_t)dense_cap + 64)))) return -1;
        cons_gather_kernel<<<(nt * 32 + 255) / 256, 256, 0, st>>>(d_tot, c->d_tasks.as<PoaTask>(), c->d_consl.as<int32_t>(), c->d_tconsoff.as<int32_t>(), c->d_consb.as<uint8_t>(),
                                                                c->d_consc.as<int32_t>(), c->d_dense_b.as<uint8_t>(), c->params.need_cov ? c->d_dense_c.as<int32_t>() : nullptr, dense_cap);
        CK(cudaMemcpyAsync(c->h_consb.p, c->d_dense_b.p, (size_t)dense_cap, cudaMemcpyDeviceToHost, st));
        if (c->params.need_cov) CK(cudaMemcpyAsync(c->h_consc.p, c->d_dense_c.p, 4 * (size_t)dense_cap, cudaMemcpyDeviceToHost, st));
        CK(th_wait(c));
    }
    if (ht->retry_n > 0 && getenv("TH_GPU_DEBUG")) fprintf(stderr, "[th_gpu] %d of %d POA tasks went to the second pass (%llu-byte slabs)\n", ht->retry_n, nt, (unsigned long long)ht->slab_full);
    // a read-level failure (chain ranking or partition limits) fails the read's tasks; th_gpu_result.read_status lets the
    // caller see it even when the read ended up with no task at all
    { const int32_t *rs = c->h_rstatus.as<int32_t>();
      for (int r = 0; r < n; ++r) if (rs[r]) for (int t = h_rtoff[r]; t < h_rtoff[r + 1]; ++t) if (!h_tstatus[t]) h_tstatus[t] = rs[r]; }
    S.d2h_bytes += 4ll * (n + 1) + 4ll * (nt + 1) * 2 + 8ll * n_pos + 4ll * nt * 2 + 16ll * nt + 4ll * n + 256 + sizeof(TaskTotals) +
                   (nt > 0 && !P.only_unit ? dense_cap * (c->params.need_cov ? 5 : 1) : 0);
    const unsigned long long *hc = c->h_counters.as<unsigned long long>();
    S.n_chain_evals = (int64_t)hc[0]; S.n_poa_cells = (int64_t)hc[1]; S.n_poa_rows = (int64_t)hc[2]; S.n_ksw_cells = (int64_t)hc[3]; S.n_ksw_cells_full = (int64_t)hc[4];
    S.ms_pack = ev_ms(c, 2, 3); S.ms_seed = ev_ms(c, 3, 4); S.ms_chain = ev_ms(c, 4, 5); S.ms_select = ev_ms(c, 5, 6); S.ms_partition = ev_ms(c, 6, 7);
    S.ms_poa = ev_ms(c, 8, 9); S.ms_ksw = ev_ms(c, 9, 10); S.ms_d2h = ev_ms(c, 10, 11);
    S.ms_total = ev_ms(c, 2, 11);
    if (nt > 0 && !P.only_unit && getenv("TH_GPU_DEBUG"))
        fprintf(stderr, "[th_gpu] POA stage %.2f ms: %.2f before the packed kernel (host + ordering kernels), %.2f packed kernel, %.2f wide pass\n", S.ms_poa, ev_ms(c, 8, 12), ev_ms(c, 12, 13), ev_ms(c, 13, 9));
    out->n_reads = n; out->n_tasks = nt;
    out->read_task_off = h_rtoff; out->task_pos_off = c->h_tposoff.as<int32_t>(); out->pos = c->h_pos.as<int32_t>();
    out->task_n_seqs = c->h_tnseqs.as<int32_t>(); out->task_cons_off = c->h_tconsoff.as<int32_t>(); out->cons_base = c->h_consb.as<uint8_t>();
    out->cons_cov = c->params.need_cov ? c->h_consc.as<int32_t>() : c->h_iden.as<int32_t>(); // only read when coverage was asked for
    out->iden_n = c->h_iden.as<int32_t>(); out->ext = c->h_ext.as<int32_t>(); out->task_status = h_tstatus;
    out->read_status = c->h_rstatus.as<int32_t>();
    out->stats = S;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_err("kernel failure: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

extern "C" int th_gpu_process_chunk(th_gpu_ctx *c, int32_t n_reads, const char *const *seq, const int32_t *seq_len, th_gpu_result *out) {
    if (th_gpu_upload(c, n_reads, seq, seq_len)) return -1;
    return th_gpu_process_resident(c, out);
}

// ---- device-side step brackets ------------------------------------------------------------------
extern "C" int th_gpu_mark(th_gpu_ctx *c, int32_t slot) {
    if (slot < 0 || slot >= 4) { set_err("mark slot out of range"); return -1; }
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->mark[slot], c->stream));
    return 0;
}
extern "C" int th_gpu_mark_elapsed(th_gpu_ctx *a, int32_t slot_a, th_gpu_ctx *b, int32_t slot_b, float *ms) {
    if (slot_a < 0 || slot_a >= 4 || slot_b < 0 || slot_b >= 4 || a->device != b->device) { set_err("bad marks"); return -1; }
    CK(cudaSetDevice(a->device));
    CK(cudaEventSynchronize(a->mark[slot_a])); CK(cudaEventSynchronize(b->mark[slot_b]));
    CK(cudaEventElapsedTime(ms, a->mark[slot_a], b->mark[slot_b]));
    return 0;
}

// ---- stage probes -----------------------------------------------------------------------------
extern "C" int th_gpu_debug_counters(th_gpu_ctx *c, int32_t cap, int64_t *out) {
    CK(cudaSetDevice(c->device));
    if (!c->d_counters.p) { set_err("no chunk processed yet"); return -1; }
    int64_t h[32];
    CK(cudaMemcpy(h, c->d_counters.p, sizeof(h), cudaMemcpyDeviceToHost));
    const int m = std::min(cap, 32);
    for (int i = 0; i < m; ++i) out[i] = h[i];
    return m;
}
extern "C" int th_gpu_debug_hits(th_gpu_ctx *c, int32_t read, int32_t cap, int32_t *end, int32_t *period) {
    CK(cudaSetDevice(c->device));
    if (read < 0 || read >= c->n_reads) { set_err("read out of range"); return -1; }
    int32_t n = 0; CK(cudaMemcpy(&n, c->d_nhits.as<int32_t>() + read, 4, cudaMemcpyDeviceToHost));
    const int m = std::min(n, cap);
    CK(cudaMemcpy(end, c->d_hend.as<int32_t>() + c->h_roff[read], 4 * (size_t)m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(period, c->d_hper.as<int32_t>() + c->h_roff[read], 4 * (size_t)m, cudaMemcpyDeviceToHost));
    return n;
}
extern "C" int th_gpu_debug_chain_dp(th_gpu_ctx *c, int32_t read, int32_t cap, int32_t *score, int32_t *from) {
    CK(cudaSetDevice(c->device));
    if (read < 0 || read >= c->n_reads) { set_err("read out of range"); return -1; }
    int32_t n = 0; CK(cudaMemcpy(&n, c->d_nhits.as<int32_t>() + read, 4, cudaMemcpyDeviceToHost));
    const int m = std::min(n, cap);
    CK(cudaMemcpy(score, c->d_score.as<int32_t>() + c->h_roff[read], 4 * (size_t)m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(from, c->d_from.as<int32_t>() + c->h_roff[read], 4 * (size_t)m, cudaMemcpyDeviceToHost));
    return n;
}
extern "C" int th_gpu_debug_chains(th_gpu_ctx *c, int32_t read, int32_t cap, int32_t *n_chain, int32_t *chain_len, int32_t *cells) {
    CK(cudaSetDevice(c->device));
    if (read < 0 || read >= c->n_reads) { set_err("read out of range"); return -1; }
    int32_t nch = 0; CK(cudaMemcpy(&nch, c->d_pchn.as<int32_t>() + read, 4, cudaMemcpyDeviceToHost));
    *n_chain = nch;
    const int64_t off = c->h_roff[read], hoff = off / 2;
    std::vector<int32_t> po(nch), pl(nch);
    if (nch) { CK(cudaMemcpy(po.data(), c->d_pchoff.as<int32_t>() + hoff, 4 * (size_t)nch, cudaMemcpyDeviceToHost));
               CK(cudaMemcpy(pl.data(), c->d_pchlen.as<int32_t>() + hoff, 4 * (size_t)nch, cudaMemcpyDeviceToHost)); }
    int tot = 0;
    for (int i = 0; i < nch; ++i) {
        chain_len[i] = pl[i];
        if (tot + pl[i] <= cap) CK(cudaMemcpy(cells + tot, c->d_cells.as<int32_t>() + off + po[i], 4 * (size_t)pl[i], cudaMemcpyDeviceToHost));
        tot += pl[i];
    }
    return tot;
}
extern "C" int th_gpu_debug_par_pos(th_gpu_ctx *c, int32_t read, int32_t chain, int32_t cap, int32_t *par_pos) {
    CK(cudaSetDevice(c->device));
    if (read < 0 || read >= c->n_reads || !c->d_par.p) { set_err("read out of range"); return -1; }
    int32_t nch = 0; CK(cudaMemcpy(&nch, c->d_pchn.as<int32_t>() + read, 4, cudaMemcpyDeviceToHost));
    if (chain < 0 || chain >= nch) return -1;
    const int64_t off = c->h_roff[read], hoff = off / 2;
    int32_t n = 0, po = 0;
    CK(cudaMemcpy(&n, c->d_parn.as<int32_t>() + hoff + chain, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&po, c->d_paroff.as<int32_t>() + hoff + chain, 4, cudaMemcpyDeviceToHost));
    const int m = std::min(n, cap);
    if (m > 0) CK(cudaMemcpy(par_pos, c->d_par.as<int32_t>() + 2 * off + po, 4 * (size_t)m, cudaMemcpyDeviceToHost));
    return n;
}

extern "C" int th_gpu_ksw_batch(th_gpu_ctx *c, int32_t n, int32_t mode, const uint8_t *const *q, const int32_t *ql,
                                const uint8_t *const *t, const int32_t *tl, const int32_t *arg, int32_t *out2) {
    CK(cudaSetDevice(c->device));
    if (n <= 0) return 0;
    std::vector<int64_t> qoff(n), toff(n); int64_t tot = 0; int maxt = 0;
    for (int i = 0; i < n; ++i) { qoff[i] = tot; tot += ql[i]; toff[i] = tot; tot += tl[i]; maxt = std::max(maxt, tl[i]); }
    std::vector<uint8_t> buf(tot + 16);
    for (int i = 0; i < n; ++i) { memcpy(buf.data() + qoff[i], q[i], ql[i]); memcpy(buf.data() + toff[i], t[i], tl[i]); }
    std::vector<int32_t> zero(n, 0);
    DBuf dbuf, dq, dt, dql, dtl, darg, dout, dbnd;
    const int64_t bnd_stride = 2 * (int64_t)(maxt + 16);
    int rc = -1;
    do {
        if (dbuf.ensure(tot + 16) || dq.ensure(8 * (size_t)n) || dt.ensure(8 * (size_t)n) || dql.ensure(4 * (size_t)n) || dtl.ensure(4 * (size_t)n) ||
            darg.ensure(4 * (size_t)n) || dout.ensure(8 * (size_t)n) || dbnd.ensure((size_t)n * bnd_stride * sizeof(int4))) break;
        if (cudaMemcpy(dbuf.p, buf.data(), tot, cudaMemcpyHostToDevice) != cudaSuccess) break;
        cudaMemcpy(dq.p, qoff.data(), 8 * (size_t)n, cudaMemcpyHostToDevice); cudaMemcpy(dt.p, toff.data(), 8 * (size_t)n, cudaMemcpyHostToDevice);
        cudaMemcpy(dql.p, ql, 4 * (size_t)n, cudaMemcpyHostToDevice); cudaMemcpy(dtl.p, tl, 4 * (size_t)n, cudaMemcpyHostToDevice);
        cudaMemcpy(darg.p, arg ? arg : zero.data(), 4 * (size_t)n, cudaMemcpyHostToDevice);
        ksw_test_kernel<<<(n * 32 + 127) / 128, 128, 0, c->stream>>>(n, mode, dbuf.as<uint8_t>(), dq.as<int64_t>(), dql.as<int32_t>(), dt.as<int64_t>(), dtl.as<int32_t>(),
                                                                   darg.as<int32_t>(), dbnd.as<int4>(), bnd_stride, dout.as<int32_t>());
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { set_err("ksw_test_kernel: %s", cudaGetErrorString(e)); break; }
        if (cudaMemcpy(out2, dout.p, 8 * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        rc = 0;
    } while (0);
    dbuf.release(); dq.release(); dt.release(); dql.release(); dtl.release(); darg.release(); dout.release(); dbnd.release();
    return rc;
}
