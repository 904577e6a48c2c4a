/*
 * th_host.h -- host-side C layer above the GPU C ABI (include/th_gpu.h).
 *
 * Mirrors what the reference keeps on the CPU around tidehunter_core: the option structure
 * (mini_tandem_para, src/tidehunter.h:47-61), the floating-point scalars derived from the integer
 * alignment results (src/gen_cons.c:204-223, src/abpoa_cons.c:100-107), the adapter / full-length
 * logic (src/gen_cons.c:224-291), record filtering (src/gen_cons.c:10-16) and the output formats of
 * mini_tandem_output (src/main.c:214-271), including its quirks.
 */
#ifndef TH_HOST_H
#define TH_HOST_H
#include <stddef.h>
#include <stdint.h>
#include "../include/th_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    th_gpu_params gpu;
    int out_fmt;            /* -f 1 FASTA, 2 tab, 3 FASTQ, 4 tab+qual */
    int min_len;            /* -m */
    int min_cov;            /* -r (integer form) */
    double min_frac;        /* -r (fraction form) */
    int only_longest;       /* -l */
    int only_full_length;   /* -F */
    int single_copy;        /* -s */
    float ada_match_rat;    /* -a */
    const char *five_seq, *three_seq; /* -5 / -3 adapter sequences (already read from their files) or NULL */
    int chunk_reads;        /* reads per GPU chunk (the reference uses 4096, src/tidehunter.h:10) */
    int lanes;              /* GPU contexts (own stream + buffers) the chunks of one th_host_run rotate over; chunk c+1 is
                               uploaded and processed while chunk c is formatted, and their kernels fill each other's
                               tails.  <= 0: default (TH_HOST_LANES or 4); 1 = strictly serial.  Per device with th_host_create_multi */
} th_host_para;

void th_host_default_para(th_host_para *p);

typedef struct th_host th_host;

/* device < 0: use CUDA device 0.  NULL on failure (see th_host_last_error). */
th_host *th_host_create(const th_host_para *p, int device);
/* `lanes` contexts on each of n_devices CUDA devices (one process drives several GPUs; reads are independent) */
th_host *th_host_create_multi(const th_host_para *p, int n_devices, const int *devices);
void th_host_destroy(th_host *h);

/* Process reads [0,n) and append the text the reference would print for them to an internal buffer;
 * returns a pointer to it (valid until the next call) and its length.  Reads are numbered across
 * calls so that the FASTQ quality slot-reuse quirk of the reference (src/main.c:266-267) is kept. */
const char *th_host_run(th_host *h, int n, const char *const *names, const char *const *seqs, const int32_t *lens, size_t *out_len);

/* Index, in the whole input, of the first read of the NEXT th_host_run, for a process that handles part of an input (one
 * rank of a sharded run): quality slots are then numbered as in the whole input.  That alone does not make a part's FASTQ
 * output (-f 3/4) equal to the reference's beyond 4,096 reads: the reference never rewinds a slot's quality buffer
 * (src/main.c:262-266), so from read 4,096 on it prints the qualities of the FIRST record ever written to the slot, which may
 * belong to another process's part.  Sharded FASTA / tabular output is exact; sharded FASTQ is exact up to 4,096 reads. */
void th_host_set_read_index(th_host *h, long long first);
/* stats of the last th_host_run (summed over its chunks) */
void th_host_stats(const th_host *h, th_gpu_stats *s);
/* consensus tasks that failed on the GPU since th_host_create (status != 0, e.g. a unit beyond the int16 score range);
 * their records are missing from the output and a message went to stderr */
long long th_host_failed_tasks(const th_host *h);
th_gpu_ctx *th_host_gpu(th_host *h);
const char *th_host_last_error(void);
/* test hook: the bit-vector adapter search against its column-by-column definition on random inputs; returns the
 * number of disagreements (0 expected) */
int th_host_selftest_infix(int trials, unsigned seed);

#ifdef __cplusplus
}
#endif
#endif
