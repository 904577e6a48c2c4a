/*
 * th_gpu.h -- C ABI of the B200-native replacement for TideHunter's per-read hot path.
 *
 * The seam it replaces (reference, paths relative to /root/reference):
 *   int tidehunter_core(kseq_t *read_seq, tandem_seq_t *tseq, mini_tandem_para *mtp,
 *                       abpoa_t *ab, abpoa_para_t *abpt);            src/tidehunter.h:85, src/tidehunter.c:23-60
 * called once per read by mini_tandem_thread_main (src/main.c:273-291) for every read of a chunk
 * (src/main.c:402-425).  Here the whole chunk crosses the boundary in one call: reads in, one flat
 * structure-of-arrays of integer results out, in input order.  Everything integer is computed on the
 * GPU (pack, seeding, chaining, partition, ksw2-style alignments, abPOA-style consensus); the few
 * floating-point scalars the reference derives from those integers (ave_match, copy_num, phred) and
 * the record formatting stay in host C (host/th_host.c) exactly as in src/gen_cons.c / src/main.c.
 *
 * Plain C types only; no CUDA or torch types in any signature.  There is no CPU fallback: every
 * entry point fails (non-zero / NULL + th_gpu_last_error()) when no CUDA device is usable.
 */
#ifndef TH_GPU_H
#define TH_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TH_GPU_ABI_VERSION 2

/* Numeric fields of mini_tandem_para (src/tidehunter.h:47-61) that reach the hot path. */
typedef struct {
    int32_t k, w, hpc;                 /* -k -w -H   (src/tandem_hit.c:160-167) */
    int32_t min_copy;                  /* -c         (src/gen_cons.c:193,200; src/tidehunter.c:42) */
    double  max_div;                   /* -e         (src/partition.c:254) */
    int64_t min_p, max_p;              /* -p -P      (src/tandem_hit.c:227-237) */
    int32_t match, mismatch;           /* -M -X      (src/abpoa_cons.c:12-28) */
    int32_t gap_open1, gap_open2;      /* -O */
    int32_t gap_ext1, gap_ext2;        /* -E */
    int32_t only_unit;                 /* -u: stop after partition (src/gen_cons.c:201-202) */
    int32_t need_cov;                  /* consensus coverage wanted (-f 3/4 or -r; src/abpoa_cons.c:86) */
    int32_t simd_lanes16;              /* abPOA int16 lanes per emulated SIMD vector; 16 = AVX2 build of the reference */
} th_gpu_params;

/* defaults of mini_tandem_init_para (src/main.c:325-362) */
void th_gpu_default_params(th_gpu_params *p);

/* Device-side time per stage for the last chunk (CUDA events on the library's own stream), plus
 * algorithmic work counters (the roofline numerators of SURVEY.md section 8d). */
typedef struct {
    float ms_h2d, ms_pack, ms_seed, ms_chain, ms_select, ms_partition, ms_poa, ms_ksw, ms_d2h, ms_total;
    int64_t n_bases, n_hits, n_chain_evals, n_poa_cells, n_poa_rows, n_ksw_cells, n_tasks;
    int64_t n_launches;                /* kernels launched for the chunk */
    int64_t h2d_bytes, d2h_bytes;
    int64_t n_ksw_cells_full;          /* n_ksw_cells counts the cells computed (certified bands, cut extension matrices);
                                          this is the full matrices' count, i.e. what the reference computes */
} th_gpu_stats;

/* Result of one chunk: structure of arrays, owned by the context, valid until the next call on it.
 * A "task" is one run of unit boundaries handed to seqs_msa's inner block (src/gen_cons.c:191-298):
 * par_pos[i..j) between -1 separators with more than min_copy entries.  Tasks of a read appear in the
 * order the reference would emit their records. */
typedef struct {
    int32_t n_reads, n_tasks;
    const int32_t *read_task_off;      /* n_reads+1 */
    const int32_t *task_pos_off;       /* n_tasks+1, into pos[] and iden_n[] */
    const int32_t *pos;                /* unit boundaries, 0-based like the reference's par_pos */
    const int32_t *task_n_seqs;        /* units that entered the consensus (src/abpoa_cons.c:40-50) */
    const int32_t *task_cons_off;      /* n_tasks+1, into cons_base[] / cons_cov[] */
    const uint8_t *cons_base;          /* consensus, nt4 codes 0..4 */
    const int32_t *cons_cov;           /* per-base coverage (heaviest-column weight); 0 when n_seqs <= 2 */
    const int32_t *iden_n;             /* ksw2_global identity count of unit u vs consensus: iden_n[task_pos_off[t]+u] */
    const int32_t *ext;                /* 4 per task: left max_q, left max_t, right max_q, right max_t (src/gen_cons.c:217-223) */
    const int32_t *task_status;        /* 0 ok; non-zero = th_gpu error code for that task */
    const int32_t *read_status;        /* n_reads; non-zero = the read itself failed before any task existed (chain ranking / partition limits) */
    th_gpu_stats stats;
} th_gpu_result;

typedef struct th_gpu_ctx th_gpu_ctx;

/* Create a context on CUDA device `device` (one context per GPU / per process rank). NULL on failure. */
th_gpu_ctx *th_gpu_create(const th_gpu_params *p, int device);
void th_gpu_destroy(th_gpu_ctx *ctx);

/* Run the hot path over `n_reads` reads held in HOST memory (ASCII, not NUL terminated).
 * Blocking; includes host->device and device->host copies.  Returns 0 on success. */
int th_gpu_process_chunk(th_gpu_ctx *ctx, int32_t n_reads, const char *const *seq, const int32_t *seq_len,
                         th_gpu_result *out);

/* Same work with the reads already resident in device memory from the previous th_gpu_process_chunk /
 * th_gpu_upload call on this context (used to time the device-only leg of the benchmark). */
int th_gpu_upload(th_gpu_ctx *ctx, int32_t n_reads, const char *const *seq, const int32_t *seq_len);
int th_gpu_process_resident(th_gpu_ctx *ctx, th_gpu_result *out);

/* Device-side step brackets: th_gpu_mark records CUDA event `slot` (0..3) on the context's stream;
 * th_gpu_mark_elapsed returns the device time from mark (a, slot_a) to mark (b, slot_b) of two contexts
 * on the same GPU (negative when b's mark was reached first).  bench.py times a step that spans
 * several contexts as max(end marks) - min(start marks). */
int th_gpu_mark(th_gpu_ctx *ctx, int32_t slot);
int th_gpu_mark_elapsed(th_gpu_ctx *a, int32_t slot_a, th_gpu_ctx *b, int32_t slot_b, float *ms);

/* Stage probes used by the parity tests (results are written to caller-allocated host arrays). */
int th_gpu_debug_hits(th_gpu_ctx *ctx, int32_t read, int32_t cap, int32_t *end, int32_t *period);      /* returns hit_n */
int th_gpu_debug_chain_dp(th_gpu_ctx *ctx, int32_t read, int32_t cap, int32_t *score, int32_t *from);   /* returns hit_n */
int th_gpu_debug_chains(th_gpu_ctx *ctx, int32_t read, int32_t cap, int32_t *n_chain, int32_t *chain_len, int32_t *cells); /* returns total cells */
int th_gpu_debug_par_pos(th_gpu_ctx *ctx, int32_t read, int32_t chain, int32_t cap, int32_t *par_pos);  /* returns par_n */

/* Raw device work / cycle counters of the last chunk (profiling aid): [0] chain evals, [1] POA cells, [2] POA rows,
 * [3] ksw cells, [16..22] POA warp-cycles per phase (setup, rows, backtrack, merge, reorder, consensus, total). */
/* test hook without a device: how the options reach the kernels (k, w, hpc, min_copy, min_p, max_p, max_div * 1e6, match,
 * mismatch, o1, e1, o2, e2, affine, o2 and e2 as given, vector width, only_unit, linear); returns the number of fields */
int th_gpu_debug_dev_params(const th_gpu_params *p, int32_t cap, int32_t *out);
int th_gpu_debug_counters(th_gpu_ctx *ctx, int32_t cap, int64_t *out);                                   /* returns entries written */

/* Stand-alone ksw2-style alignments on nt4-coded HOST sequences (tests of the alignment kernels).
 * mode 0: global -> out[0] = iden_n;  mode 1: global + left-end projection with q_left_ext = arg ->
 * out[0] = iden_n, out[1] = t_left_ext;  mode 2: extension -> out[0] = max_q, out[1] = max_t;  mode 3: the same for
 * entries 2k and 2k+1 together (no N), through the packed two-extensions-per-warp routine. */
int th_gpu_ksw_batch(th_gpu_ctx *ctx, int32_t n, int32_t mode, const uint8_t *const *q, const int32_t *ql,
                     const uint8_t *const *t, const int32_t *tl, const int32_t *arg, int32_t *out2);

const char *th_gpu_last_error(void);
int th_gpu_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
