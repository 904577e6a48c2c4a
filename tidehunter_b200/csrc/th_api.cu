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
    DBuf d_tasks, d_torder, d_ustart, d_ulen, d_slabs, d_consb, d_consc, d_consl, d_tstatus, d_items, d_iden, d_ext;
    DBuf d_gsrc, d_gdst, d_glen, d_dense_b, d_dense_c, d_redo;
    HBuf h_tmp, h_tmp2, h_consb, h_consc;
    // host result storage
    std::vector<int32_t> r_read_task_off, r_task_pos_off, r_pos, r_task_n_seqs, r_task_cons_off, r_cons_cov, r_iden, r_ext, r_task_status;
    std::vector<uint8_t> r_cons_base;
    // debug copies of the last chunk
    std::vector<int32_t> dbg_par_stream; std::vector<int64_t> dbg_par_doff;
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
    d.affine = p->gap_open2 == 0;
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
                         d.o1, d.e1, d.o2, d.e2, d.affine, d.o2_raw, d.e2_raw, d.pn, d.only_unit};
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
    if (p->gap_open1 <= 0 || p->gap_open2 < 0 || p->gap_ext1 <= 0 || p->gap_ext2 < 0) { set_err("abPOA's linear gap mode (O1 = 0) is not implemented on the GPU path; convex (O1, O2 > 0) and affine (O2 = 0) are"); return nullptr; }
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
    CKP(cudaFuncSetAttribute(seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEED_SMEM_CAP * 8));
    CKP(cudaFuncSetAttribute(poa_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(PoaSmem<16>) * POA_WARPS * 2 + sizeof(PoaLaneK) * 16)));
    CKP(cudaFuncSetAttribute(poa_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(PoaSmem<32>) * POA_WARPS + sizeof(PoaLaneK) * 32)));
    memset(&c->stats, 0, sizeof(c->stats));
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
                  &c->d_consl, &c->d_tstatus, &c->d_items, &c->d_iden, &c->d_ext, &c->d_gsrc, &c->d_gdst, &c->d_glen, &c->d_dense_b, &c->d_dense_c, &c->d_redo};
    for (DBuf *b : ds) b->release();
    HBuf *hs[] = {&c->h_ascii, &c->h_tmp, &c->h_tmp2, &c->h_consb, &c->h_consc};
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

// generic range gather: one warp per range
template <class T>
__global__ void gather_kernel(int n, const int64_t *__restrict__ src_off, const int64_t *__restrict__ dst_off, const int32_t *__restrict__ len,
                              const T *__restrict__ src, T *__restrict__ dst) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const T *s = src + src_off[w]; T *d = dst + dst_off[w]; const int l = len[w];
    for (int i = lane; i < l; i += 32) d[i] = s[i];
}

// serialises every read's partition result as [n_chains, (par_n, values...)*] at the start of its par region
__global__ void par_stream_kernel(int n_reads, const int64_t *__restrict__ roff, const int32_t *__restrict__ pch_n,
                                  const int32_t *__restrict__ par, const int32_t *__restrict__ par_off, const int32_t *__restrict__ par_n,
                                  int32_t *__restrict__ stream, int32_t *__restrict__ used) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const int64_t off = roff[r], hoff = off / 2;
    int32_t *o = stream + 2 * off; const int32_t *p = par + 2 * off;
    const int nch = pch_n[r];
    int u = 0; o[u++] = nch;
    for (int c = 0; c < nch; ++c) {
        const int n = par_n[hoff + c], po = par_off[hoff + c];
        o[u++] = n;
        for (int i = 0; i < n; ++i) o[u++] = p[po + i];
    }
    used[r] = u;
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
    else ksw_warp<KSW_EXT, 16>(buf + qoff[w], ql[w], buf + toff[w], tl[w], 0, bnd, o0, o1);
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
    c->r_read_task_off.assign(n + 1, 0);
    c->r_task_pos_off.assign(1, 0); c->r_pos.clear(); c->r_task_n_seqs.clear(); c->r_task_cons_off.assign(1, 0);
    c->r_cons_base.clear(); c->r_cons_cov.clear(); c->r_iden.clear(); c->r_ext.clear(); c->r_task_status.clear();
    if (n == 0) { out->read_task_off = c->r_read_task_off.data(); out->task_pos_off = c->r_task_pos_off.data(); out->task_cons_off = c->r_task_cons_off.data(); return 0; }
    for (int i = 0; i < n; ++i) S.n_bases += c->h_rlen[i];
    const size_t B4 = (size_t)B * 4;
    if (c->d_bseq.ensure(B + 64) || c->d_pack.ensure(B / 4 + 64) || c->d_nmask.ensure(B / 8 + 64) ||
        c->d_hend.ensure(B4) || c->d_hper.ensure(B4) || c->d_nhits.ensure(4 * (size_t)n) || c->d_score.ensure(B4) || c->d_from.ensure(B4) ||
        c->d_gflag.ensure(4 * (size_t)n) || c->d_rank.ensure(B4) || c->d_nrank.ensure(4 * (size_t)n) || c->d_tracked.ensure(B) ||
        c->d_choff.ensure(B4 / 2 + 256) || c->d_chlen.ensure(B4 / 2 + 256) || c->d_chscore.ensure(B4 / 2 + 256) || c->d_chidx.ensure(B4 / 2 + 256) ||
        c->d_cells.ensure(B4) || c->d_pchn.ensure(4 * (size_t)n) || c->d_pchoff.ensure(B4 / 2 + 256) || c->d_pchlen.ensure(B4 / 2 + 256) ||
        c->d_par.ensure(2 * B4) || c->d_paroff.ensure(B4 / 2 + 256) || c->d_parn.ensure(B4 / 2 + 256) || c->d_rstatus.ensure(4 * (size_t)n) ||
        c->d_parstream.ensure(2 * B4) || c->d_parused.ensure(4 * (size_t)n) || c->d_pardoff.ensure(8 * (size_t)n) || c->d_counters.ensure(256))
        return -1;
    // counters: [0] chain evals (u64), [1] poa cells, [2] poa rows, [3] ksw cells, [8..] int work counters
    CK(cudaMemsetAsync(c->d_counters.p, 0, 256, st));
    CK(cudaMemsetAsync(c->d_rstatus.p, 0, 4 * (size_t)n, st));
    unsigned long long *cnt64 = c->d_counters.as<unsigned long long>();
    int *cnt32 = c->d_counters.as<int>() + 16;
    const int64_t *roff = c->d_roff.as<int64_t>(); const int32_t *rlen = c->d_rlen.as<int32_t>();
    int ei = 2; // event index
    // ---- pack ----
    CK(cudaEventRecord(c->ev[ei++], st)); // 2
    { const int64_t nw = B / 32; pack_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(c->d_ascii.as<uint8_t>(), c->d_bseq.as<uint8_t>(), c->d_pack.as<uint64_t>(), c->d_nmask.as<uint32_t>(), nw); S.n_launches++; }
    CK(cudaEventRecord(c->ev[ei++], st)); // 3
    // ---- seed ----
    {
        int gcap = 1; while (gcap < c->max_len) gcap <<= 1;
        const int grid = std::min(n, c->n_sm);
        if (c->d_scratch.ensure((size_t)grid * 2 * gcap * 8)) return -1;
        seed_kernel<<<grid, SEED_THREADS, SEED_SMEM_CAP * 8, st>>>(P, n, roff, rlen, c->d_bseq.as<uint8_t>(), c->d_pack.as<uint64_t>(), c->d_nmask.as<uint32_t>(),
                                                                  c->d_scratch.as<uint64_t>(), gcap, c->d_hend.as<int32_t>(), c->d_hper.as<int32_t>(), c->d_nhits.as<int32_t>());
        S.n_launches++;
    }
    CK(cudaEventRecord(c->ev[ei++], st)); // 4
    // ---- chain DP ----
    {
        const int grid = std::min((n + CHAIN_WARPS - 1) / CHAIN_WARPS, std::max(1, (int)(c->n_sm * CHAIN_BLOCKS_PER_SM * c->share)));
        if (P.max_p < (1u << 27)) // hit periods are <= max_p (src/tandem_hit.c:204): the 1.8x gate fits 32-bit products
            chain_dp_kernel<true><<<grid, CHAIN_WARPS * 32, 0, st>>>(P, n, roff, c->d_nhits.as<int32_t>(), c->d_hend.as<int32_t>(), c->d_hper.as<int32_t>(),
                                                                  c->d_score.as<int32_t>(), c->d_from.as<int32_t>(), c->d_gflag.as<int32_t>(), cnt64 + 0);
        else
            chain_dp_kernel<false><<<grid, CHAIN_WARPS * 32, 0, st>>>(P, n, roff, c->d_nhits.as<int32_t>(), c->d_hend.as<int32_t>(), c->d_hper.as<int32_t>(),
                                                                   c->d_score.as<int32_t>(), c->d_from.as<int32_t>(), c->d_gflag.as<int32_t>(), cnt64 + 0);
        S.n_launches++;
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
    // ---- partition ----
    const int n_pwarps = std::min(n, std::max(4, (int)(c->n_sm * PART_MIN_BLOCKS * PART_WARPS * c->share)));
    const int64_t bnd_stride = 2 * (int64_t)(c->max_len + 64);
    {
        if (c->d_bnd.ensure((size_t)n_pwarps * bnd_stride * sizeof(int4))) return -1;
        const int grid = (n_pwarps + PART_WARPS - 1) / PART_WARPS;
        if (c->d_bnd.ensure((size_t)grid * PART_WARPS * bnd_stride * sizeof(int4))) return -1;
        partition_kernel<<<grid, PART_WARPS * 32, 0, st>>>(P, n, roff, rlen, c->d_bseq.as<uint8_t>(), c->d_hend.as<int32_t>(), c->d_hper.as<int32_t>(), c->d_cells.as<int32_t>(),
                                                          c->d_pchn.as<int32_t>(), c->d_pchoff.as<int32_t>(), c->d_pchlen.as<int32_t>(), c->d_par.as<int32_t>(),
                                                          c->d_paroff.as<int32_t>(), c->d_parn.as<int32_t>(), c->d_bnd.as<int4>(), bnd_stride, cnt32 + 0,
                                                          c->d_rstatus.as<int32_t>(), cnt64 + 3);
        par_stream_kernel<<<(n + 63) / 64, 64, 0, st>>>(n, roff, c->d_pchn.as<int32_t>(), c->d_par.as<int32_t>(), c->d_paroff.as<int32_t>(), c->d_parn.as<int32_t>(),
                                                       c->d_parstream.as<int32_t>(), c->d_parused.as<int32_t>());
        S.n_launches += 2;
    }
    CK(cudaEventRecord(c->ev[ei++], st)); // 7
    // ---- bring the partition streams to the host (dense) ----
    std::vector<int32_t> used(n), nhits_h(n);
    CK(cudaMemcpyAsync(used.data(), c->d_parused.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nhits_h.data(), c->d_nhits.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(th_wait(c));
    std::vector<int64_t> doff(n + 1, 0), soff(n);
    for (int i = 0; i < n; ++i) { doff[i + 1] = doff[i] + used[i]; soff[i] = 2 * c->h_roff[i]; S.n_hits += nhits_h[i]; }
    const int64_t tot_stream = doff[n];
    if (c->d_gsrc.ensure(8 * (size_t)n) || c->d_gdst.ensure(8 * (size_t)n) || c->d_dense_c.ensure(4 * (size_t)tot_stream + 64)) return -1;
    CK(cudaMemcpyAsync(c->d_gsrc.p, soff.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_gdst.p, doff.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, st));
    gather_kernel<int32_t><<<(n * 32 + 255) / 256, 256, 0, st>>>(n, c->d_gsrc.as<int64_t>(), c->d_gdst.as<int64_t>(), c->d_parused.as<int32_t>(), c->d_parstream.as<int32_t>(), c->d_dense_c.as<int32_t>());
    S.n_launches++;
    c->dbg_par_stream.resize(tot_stream); c->dbg_par_doff = doff;
    CK(cudaMemcpyAsync(c->dbg_par_stream.data(), c->d_dense_c.p, 4 * (size_t)tot_stream, cudaMemcpyDeviceToHost, st));
    CK(th_wait(c));
    S.d2h_bytes += 4 * tot_stream + 8 * (int64_t)n;
    // ---- host: split runs into tasks exactly as seqs_msa (src/gen_cons.c:191-200) ----
    std::vector<PoaTask> tasks; std::vector<int32_t> ustart, ulen; std::vector<KswItem> items;
    int pending_single = -1; // index of an unpaired single-unit item
    int64_t cons_total = 0;
    const int min_copy = c->params.min_copy;
    for (int r = 0; r < n; ++r) {
        const int32_t *sp = c->dbg_par_stream.data() + doff[r];
        const int L = c->h_rlen[r];
        int u = 0; const int nch = sp[u++];
        for (int ch = 0; ch < nch; ++ch) {
            const int par_n = sp[u++]; const int32_t *par = sp + u; u += par_n;
            if (par_n < min_copy + 1) continue; // src/tidehunter.c:42
            int i = 0;
            while (i < par_n - min_copy) {
                if (par[i] < 0) { ++i; continue; }
                int j;
                for (j = i + 1; j < par_n; ++j) if (par[j] < 0) break;
                if (j - i > min_copy) {
                    PoaTask T; memset(&T, 0, sizeof(T));
                    T.read = r; T.seq_off = c->h_roff[r]; T.unit_off = (int32_t)ustart.size();
                    int nseq = 0, sum = 0, qmax = 0;
                    for (int q = i; q < j - 1; ++q) { // src/abpoa_cons.c:40-50
                        const int start = par[q], end = par[q + 1];
                        if (start < 0 || end < 0 || start >= L - 1 || end + 1 > L) continue;
                        ustart.push_back(start + 1); ulen.push_back(end - start); ++nseq; sum += end - start; qmax = std::max(qmax, end - start);
                    }
                    T.n_seqs = nseq; T.ncap = sum + 2; T.qmax = qmax; T.cons_off = (int32_t)cons_total;
                    cons_total += sum + 4;
                    const int t = (int)tasks.size();
                    tasks.push_back(T);
                    for (int q = i; q < j; ++q) c->r_pos.push_back(par[q]);
                    c->r_task_pos_off.push_back((int32_t)c->r_pos.size());
                    c->r_task_n_seqs.push_back(nseq);
                    if (!P.only_unit) { // post-consensus alignments of seqs_msa (src/gen_cons.c:208-223); units go two per warp
                        const int p0 = (int)c->r_pos.size() - (j - i);
                        for (int q = i; q < j - 1; q += 2) {
                            KswItem it; memset(&it, 0, sizeof(it));
                            it.task = t; it.seq_off = c->h_roff[r]; it.out = p0 + (q - i);
                            it.a = par[q] + 1; it.b = par[q + 1] - par[q];
                            if (q + 1 < j - 1) { it.kind = 3; it.a2 = par[q + 1] + 1; it.b2 = par[q + 2] - par[q + 1]; items.push_back(it); }
                            else if (pending_single >= 0) { // pair the left-over unit with the previous task's left-over
                                KswItem &o = items[pending_single];
                                o.kind = 4; o.task2 = t; o.a2 = it.a; o.b2 = it.b; o.out2 = it.out; o.seq_off2 = it.seq_off;
                                pending_single = -1;
                            } else { it.kind = 0; pending_single = (int)items.size(); items.push_back(it); }
                        }
                        KswItem le; memset(&le, 0, sizeof(le)); le.kind = 1; le.task = t; le.a = par[i] + 1; le.seq_off = c->h_roff[r]; le.out = 4 * t; items.push_back(le);
                        KswItem re; memset(&re, 0, sizeof(re)); re.kind = 2; re.task = t; re.a = par[j - 1] + 1; re.b = L - par[j - 1] - 1; re.seq_off = c->h_roff[r]; re.out = 4 * t + 2; items.push_back(re);
                    }
                }
                i = j + 1;
            }
        }
        c->r_read_task_off[r + 1] = (int32_t)tasks.size();
    }
    const int nt = (int)tasks.size();
    S.n_tasks = nt;
    c->r_task_status.assign(nt, 0); c->r_task_cons_off.assign(nt + 1, 0);
    c->r_iden.assign(c->r_pos.size(), 0); c->r_ext.assign((size_t)nt * 4, -1);
    CK(cudaEventRecord(c->ev[ei++], st)); // 8
    if (nt > 0 && !P.only_unit) {
        // ---- POA ----
        std::vector<int32_t> order(nt);
        for (int i = 0; i < nt; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return (int64_t)tasks[a].ncap * tasks[a].n_seqs > (int64_t)tasks[b].ncap * tasks[b].n_seqs; });
        // Slabs: graph arrays + the DP arena of ONE alignment (recycled for every unit): rows <= nodes, three int16 planes
        // (H, E1, E2) and a code byte per banded cell.  Typical: the graph holds <= ~2.5 units worth of nodes; the adaptive band is 2w+1
        // columns around the predecessors' row maxima, rounded to whole vectors (measured mean ~61 columns on 1 kb units)
        auto slab_need = [](const PoaTask &T, bool full) -> size_t {
            const size_t fixed = poa_fixed_bytes(T.ncap, T.qmax, T.n_seqs);
            const size_t fullb = (size_t)T.ncap * (((size_t)T.qmax + 64) * 7 + 16);
            if (full) return fixed + fullb + 4096;
            const int wband = 10 + T.qmax / 100;
            const size_t rows_typ = std::min<size_t>((size_t)T.ncap, (size_t)T.qmax * 5 / 2 + 64);
            const size_t width_typ = std::min<size_t>((size_t)T.qmax + 64, (size_t)2 * wband + 64);
            const size_t typ = std::max<size_t>(rows_typ * (width_typ * 7 + 16), (size_t)1 << 20);
            return fixed + std::min(typ, fullb) + 4096;
        };
        size_t slab_typ = 0;
        for (const PoaTask &T : tasks) if (T.n_seqs > 2) slab_typ = std::max(slab_typ, slab_need(T, false));
        slab_typ = (slab_typ + 255) & ~(size_t)255; // slabs hold 16-byte accesses
        if (c->d_tasks.ensure(sizeof(PoaTask) * (size_t)nt) || c->d_torder.ensure(4 * (size_t)nt) || c->d_ustart.ensure(4 * ustart.size() + 64) ||
            c->d_ulen.ensure(4 * ulen.size() + 64) || c->d_consb.ensure((size_t)cons_total + 64) || c->d_consc.ensure(4 * (size_t)cons_total + 64) ||
            c->d_consl.ensure(4 * (size_t)nt) || c->d_tstatus.ensure(4 * (size_t)nt)) return -1;
        CK(cudaMemcpyAsync(c->d_tasks.p, tasks.data(), sizeof(PoaTask) * (size_t)nt, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(c->d_torder.p, order.data(), 4 * (size_t)nt, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(c->d_ustart.p, ustart.data(), 4 * ustart.size(), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(c->d_ulen.p, ulen.data(), 4 * ulen.size(), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(c->d_consl.p, 0, 4 * (size_t)nt, st));
        CK(cudaMemsetAsync(c->d_tstatus.p, 0, 4 * (size_t)nt, st));
        size_t free_b = 0, total_b = 0; CK(cudaMemGetInfo(&free_b, &total_b));
        if (slab_typ == 0) slab_typ = 1 << 20;
        size_t budget = (size_t)((double)(free_b + c->d_slabs.cap) * 0.6);
        budget = std::min(budget, total_b / 5); // several contexts share the device (the host layer runs four): none may take it all
        // first pass: 16-lane groups, two tasks per warp
        constexpr int GPB16 = POA_WARPS * 2, GPB32 = POA_WARPS;
        int ngroups = (int)std::min<size_t>((size_t)(c->n_sm * POA_MIN_BLOCKS16 * GPB16 * c->share), std::max<size_t>(1, budget / slab_typ));
        ngroups = std::min(ngroups, std::max(nt, 1));
        int grid = (ngroups + GPB16 - 1) / GPB16;
        // several contexts of one process (the host layer's lanes) size their slabs from the same free-memory reading:
        // if the allocation loses that race, run with fewer resident groups instead of failing the chunk
        while (c->d_slabs.ensure((size_t)grid * GPB16 * slab_typ)) {
            cudaGetLastError();
            if (grid <= 4) return -1;
            grid = (grid + 1) / 2;
        }
        poa_kernel<16><<<grid, POA_WARPS * 32, sizeof(PoaSmem<16>) * GPB16 + sizeof(PoaLaneK) * 16, st>>>(P, nt, c->d_tasks.as<PoaTask>(), c->d_torder.as<int32_t>(), c->d_ustart.as<int32_t>(), c->d_ulen.as<int32_t>(),
                                                 c->d_bseq.as<uint8_t>(), c->d_slabs.as<uint8_t>(), slab_typ, cnt32 + 1, c->d_consb.as<uint8_t>(), c->d_consc.as<int32_t>(),
                                                 c->d_consl.as<int32_t>(), c->d_tstatus.as<int32_t>(), cnt64 + 1, cnt64 + 2, cnt64 + 16, grid * GPB16);
        S.n_launches++;
        CK(cudaMemcpyAsync(c->r_task_status.data(), c->d_tstatus.p, 4 * (size_t)nt, cudaMemcpyDeviceToHost, st));
        CK(th_wait(c));
        // second pass, one task per warp with full-width slabs: tasks whose DP arena overflowed the typical slab, and rows
        // with more than 16 predecessors.  The slab is sized for the retried tasks only; when memory is short the pass
        // runs with fewer resident warps, and if even one slab cannot be had the tasks keep their error status (the host
        // layer reports and drops those records) -- the chunk goes on.
        std::vector<int32_t> retry;
        for (int t = 0; t < nt; ++t) if (c->r_task_status[t] == TH_ERR_ARENA || c->r_task_status[t] == TH_ERR_CAP) retry.push_back(t);
        if (!retry.empty()) {
            size_t slab_full = 0;
            for (int t : retry) slab_full = std::max(slab_full, slab_need(tasks[t], true));
            slab_full = (slab_full + 255) & ~(size_t)255;
            if (getenv("TH_GPU_DEBUG")) fprintf(stderr, "[th_gpu] %d of %d POA tasks are retried with %zu-byte slabs (first pass: %zu)\n", (int)retry.size(), nt, slab_full, slab_typ);
            CK(cudaMemGetInfo(&free_b, &total_b));
            budget = (size_t)((double)(free_b + c->d_slabs.cap) * 0.8);
            int rg = (int)std::min<size_t>(retry.size(), std::max<size_t>(1, budget / slab_full));
            rg = std::min(rg, (int)(c->n_sm * POA_MIN_BLOCKS32 * GPB32));
            bool have = false;
            while (true) {
                const int rgrid = (rg + GPB32 - 1) / GPB32;
                // fewer groups than a block holds: the extra groups of the last block find no task and never touch their slab
                if (!c->d_slabs.ensure((size_t)std::min(rg, rgrid * GPB32) * slab_full)) { have = true; break; }
                cudaGetLastError();
                if (rg <= 1) break;
                rg = (rg + 1) / 2;
            }
            if (have) {
                const int rgrid = (rg + GPB32 - 1) / GPB32;
                CK(cudaMemcpyAsync(c->d_torder.p, retry.data(), 4 * retry.size(), cudaMemcpyHostToDevice, st));
                CK(cudaMemsetAsync(cnt32 + 1, 0, 4, st));
                poa_kernel<32><<<rgrid, POA_WARPS * 32, sizeof(PoaSmem<32>) * GPB32 + sizeof(PoaLaneK) * 32, st>>>(P, std::min((int)retry.size(), INT_MAX), c->d_tasks.as<PoaTask>(), c->d_torder.as<int32_t>(), c->d_ustart.as<int32_t>(), c->d_ulen.as<int32_t>(),
                                                          c->d_bseq.as<uint8_t>(), c->d_slabs.as<uint8_t>(), slab_full, cnt32 + 1, c->d_consb.as<uint8_t>(), c->d_consc.as<int32_t>(),
                                                          c->d_consl.as<int32_t>(), c->d_tstatus.as<int32_t>(), cnt64 + 1, cnt64 + 2, cnt64 + 16, rg);
                S.n_launches++;
            } else set_err("POA retry: no memory for one %zu-byte slab; %d tasks keep their error status", slab_full, (int)retry.size());
        }
        CK(cudaEventRecord(c->ev[ei++], st)); // 9
        // ---- post-consensus ksw items ----
        const int ni = (int)items.size();
        {
            std::vector<int32_t> coff(nt);
            for (int t = 0; t < nt; ++t) coff[t] = tasks[t].cons_off;
            // pairs, singles, extensions: one kernel each (own register budgets)
            std::stable_sort(items.begin(), items.end(), [](const KswItem &x, const KswItem &y) {
                auto grp = [](int k) { return k >= 3 ? 0 : (k == 0 ? 1 : 2); };
                return grp(x.kind) < grp(y.kind); });
            int n_pairs = 0, n_singles = 0, n_exts = 0;
            for (const KswItem &it : items) { if (it.kind >= 3) ++n_pairs; else if (it.kind == 0) ++n_singles; else ++n_exts; }
            if (c->d_items.ensure(sizeof(KswItem) * (size_t)ni + 64) || c->d_iden.ensure(4 * c->r_pos.size() + 64) || c->d_ext.ensure(16 * (size_t)nt + 64) || c->d_glen.ensure(4 * (size_t)nt)) return -1;
            CK(cudaMemcpyAsync(c->d_items.p, items.data(), sizeof(KswItem) * (size_t)ni, cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(c->d_iden.p, 0, 4 * c->r_pos.size() + 64, st));
            CK(cudaMemcpyAsync(c->d_glen.p, coff.data(), 4 * (size_t)nt, cudaMemcpyHostToDevice, st));
            const int64_t rev_stride = 2 * (int64_t)(c->max_len + 64);
            const KswItem *d_pairs = c->d_items.as<KswItem>(), *d_singles = d_pairs + n_pairs, *d_exts = d_singles + n_singles;
            auto grid_for = [&](int n_work, int min_blocks) { const int kw = std::min(std::max(n_work, 1), std::max(4, (int)(c->n_sm * min_blocks * KSW_WARPS * c->share))); return (kw + KSW_WARPS - 1) / KSW_WARPS; };
            const int g_pair = grid_for(n_pairs, KSW_PAIR_MIN_BLOCKS), g_single = grid_for(n_singles + 64, KSW_MIN_BLOCKS), g_ext = grid_for(n_exts, KSW_EXT_MIN_BLOCKS);
            const int g_max = std::max(g_pair, std::max(g_single, g_ext));
            if (c->d_bnd.ensure((size_t)g_max * KSW_WARPS * bnd_stride * sizeof(int4)) || c->d_rev.ensure((size_t)g_ext * KSW_WARPS * rev_stride) ||
                c->d_redo.ensure(4 * (size_t)n_pairs + 64)) return -1;
            if (n_pairs > 0) {
                ksw_pair_kernel<<<g_pair, KSW_WARPS * 32, 0, st>>>(n_pairs, d_pairs, c->d_bseq.as<uint8_t>(), c->d_consb.as<uint8_t>(), c->d_glen.as<int32_t>(),
                                                                 c->d_consl.as<int32_t>(), c->d_bnd.as<int4>(), bnd_stride, cnt32 + 2, c->d_redo.as<int32_t>(), cnt32 + 5,
                                                                 c->d_iden.as<int32_t>(), cnt64 + 3);
                S.n_launches++;
            }
            ksw_single_kernel<<<g_single, KSW_WARPS * 32, 0, st>>>(n_singles, d_singles, d_pairs, c->d_redo.as<int32_t>(), cnt32 + 5, c->d_bseq.as<uint8_t>(),
                                                                 c->d_consb.as<uint8_t>(), c->d_glen.as<int32_t>(), c->d_consl.as<int32_t>(), c->d_bnd.as<int4>(), bnd_stride,
                                                                 cnt32 + 3, c->d_iden.as<int32_t>(), cnt64 + 3);
            S.n_launches++;
            if (n_exts > 0) {
                ksw_ext_kernel<<<g_ext, KSW_WARPS * 32, 0, st>>>(n_exts, d_exts, c->d_bseq.as<uint8_t>(), c->d_consb.as<uint8_t>(), c->d_glen.as<int32_t>(),
                                                               c->d_consl.as<int32_t>(), c->d_rev.as<uint8_t>(), rev_stride, c->d_bnd.as<int4>(), bnd_stride, cnt32 + 4,
                                                               c->d_ext.as<int32_t>(), cnt64 + 3);
                S.n_launches++;
            }
        }
        CK(cudaEventRecord(c->ev[ei++], st)); // 10
        // ---- results to the host ----
        std::vector<int32_t> cl(nt);
        CK(cudaMemcpyAsync(cl.data(), c->d_consl.p, 4 * (size_t)nt, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(c->r_task_status.data(), c->d_tstatus.p, 4 * (size_t)nt, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(c->r_iden.data(), c->d_iden.p, 4 * c->r_pos.size(), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(c->r_ext.data(), c->d_ext.p, 16 * (size_t)nt, cudaMemcpyDeviceToHost, st));
        CK(th_wait(c));
        std::vector<int64_t> gs(nt), gd(nt);
        int64_t tot = 0;
        for (int t = 0; t < nt; ++t) { gs[t] = tasks[t].cons_off; gd[t] = tot; c->r_task_cons_off[t] = (int32_t)tot; tot += cl[t]; }
        c->r_task_cons_off[nt] = (int32_t)tot;
        c->r_cons_base.resize(tot); c->r_cons_cov.assign(tot, 0);
        if (tot > 0) {
            if (c->d_gsrc.ensure(8 * (size_t)nt) || c->d_gdst.ensure(8 * (size_t)nt) || c->d_dense_b.ensure(tot + 64) || c->d_dense_c.ensure(4 * (size_t)tot + 64)) return -1;
            CK(cudaMemcpyAsync(c->d_gsrc.p, gs.data(), 8 * (size_t)nt, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(c->d_gdst.p, gd.data(), 8 * (size_t)nt, cudaMemcpyHostToDevice, st));
            gather_kernel<uint8_t><<<(nt * 32 + 255) / 256, 256, 0, st>>>(nt, c->d_gsrc.as<int64_t>(), c->d_gdst.as<int64_t>(), c->d_consl.as<int32_t>(), c->d_consb.as<uint8_t>(), c->d_dense_b.as<uint8_t>());
            S.n_launches++;
            CK(cudaMemcpyAsync(c->r_cons_base.data(), c->d_dense_b.p, (size_t)tot, cudaMemcpyDeviceToHost, st));
            if (c->params.need_cov) {
                gather_kernel<int32_t><<<(nt * 32 + 255) / 256, 256, 0, st>>>(nt, c->d_gsrc.as<int64_t>(), c->d_gdst.as<int64_t>(), c->d_consl.as<int32_t>(), c->d_consc.as<int32_t>(), c->d_dense_c.as<int32_t>());
                S.n_launches++;
                CK(cudaMemcpyAsync(c->r_cons_cov.data(), c->d_dense_c.p, 4 * (size_t)tot, cudaMemcpyDeviceToHost, st));
            }
        }
        CK(cudaEventRecord(c->ev[ei++], st)); // 11
        CK(th_wait(c));
        S.d2h_bytes += 4ll * nt * 2 + 4ll * (int64_t)c->r_pos.size() + 16ll * nt + tot * (c->params.need_cov ? 5 : 1);
        S.ms_poa = ev_ms(c, 8, 9); S.ms_ksw = ev_ms(c, 9, 10); S.ms_d2h = ev_ms(c, 10, 11);
    } else {
        CK(cudaEventRecord(c->ev[ei++], st)); CK(cudaEventRecord(c->ev[ei++], st)); CK(cudaEventRecord(c->ev[ei++], st));
        CK(th_wait(c));
    }
    // per-read status -> tasks of that read
    { std::vector<int32_t> rs(n);
      CK(cudaMemcpyAsync(rs.data(), c->d_rstatus.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st)); CK(th_wait(c));
      for (int r = 0; r < n; ++r) if (rs[r]) for (int t = c->r_read_task_off[r]; t < c->r_read_task_off[r + 1]; ++t) if (!c->r_task_status[t]) c->r_task_status[t] = rs[r]; }
    unsigned long long hc[4];
    CK(cudaMemcpy(hc, c->d_counters.p, sizeof(hc), cudaMemcpyDeviceToHost));
    S.n_chain_evals = (int64_t)hc[0]; S.n_poa_cells = (int64_t)hc[1]; S.n_poa_rows = (int64_t)hc[2]; S.n_ksw_cells = (int64_t)hc[3];
    S.ms_pack = ev_ms(c, 2, 3); S.ms_seed = ev_ms(c, 3, 4); S.ms_chain = ev_ms(c, 4, 5); S.ms_select = ev_ms(c, 5, 6); S.ms_partition = ev_ms(c, 6, 7);
    S.ms_total = ev_ms(c, 2, ei - 1);
    out->n_reads = n; out->n_tasks = nt;
    out->read_task_off = c->r_read_task_off.data(); out->task_pos_off = c->r_task_pos_off.data(); out->pos = c->r_pos.data();
    out->task_n_seqs = c->r_task_n_seqs.data(); out->task_cons_off = c->r_task_cons_off.data(); out->cons_base = c->r_cons_base.data();
    out->cons_cov = c->r_cons_cov.data(); out->iden_n = c->r_iden.data(); out->ext = c->r_ext.data(); out->task_status = c->r_task_status.data();
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
    if (read < 0 || read >= c->n_reads || c->dbg_par_doff.empty()) { set_err("read out of range"); return -1; }
    const int32_t *sp = c->dbg_par_stream.data() + c->dbg_par_doff[read];
    int u = 0; const int nch = sp[u++];
    if (chain < 0 || chain >= nch) return -1;
    for (int ch = 0; ch < nch; ++ch) {
        const int n = sp[u++];
        if (ch == chain) { for (int i = 0; i < n && i < cap; ++i) par_pos[i] = sp[u + i]; return n; }
        u += n;
    }
    return -1;
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
