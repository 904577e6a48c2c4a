/*
 * th_oracle.h -- CPU restatement ("port") of TideHunter v1.5.5's per-read hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (tidehunter_b200/, host/) may include, link or
 * execute this code; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_pin.py) against
 *   (a) the reference's own known answers: README.md:222 (test_50x4) and the golden stdout md5s of
 *       SURVEY.md section 8(c) for test.fq (-f 1..4), test_1000x10.fa and test_50x4.fa, and
 *   (b) outputs of the unmodified reference compiled here by oracle/Makefile into oracle/_ref/
 *       (fixtures under tests/golden/ made by tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef TH_ORACLE_H
#define TH_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* numeric fields of mini_tandem_para (src/tidehunter.h:47-61) + the SIMD-width emulation knob */
typedef struct {
    int k, w, hpc;
    int min_copy, min_cov;
    double max_div, min_frac;
    int64_t min_p, max_p;
    int match, mismatch, gap_open1, gap_open2, gap_ext1, gap_ext2;
    int out_fmt, min_len, only_unit, only_longest, only_full_length, single_copy;
    float ada_match_rat;
    const char *five_seq, *three_seq; /* NUL-terminated adapter sequences or NULL */
    int pn16;                         /* int16 lanes per emulated abPOA vector: 16 = AVX2 (reference default) */
} tho_para_t;

void tho_default_para(tho_para_t *p); /* src/main.c:325-362 */

/* ---- stage functions (each usable on its own from ctypes) ---- */

/* src/seq.c:77-87 */
void tho_get_bseq(const char *seq, int len, uint8_t *bseq);

/* src/tandem_hit.c:227-237.  Returns hit_n; *hits is malloc'ed (end<<32|period), caller frees. */
int tho_collect_hits(const uint8_t *bseq, int len, const tho_para_t *p, uint64_t **hits);

/* chaining result, flattened.  cell c of chain i = cells[chain_off[i] + c] = index into the flat dp
 * arrays (dp rows are flattened in row-major order: row = distinct end, ascending period). */
typedef struct {
    int n_cells;                 /* == hit_n */
    int *start, *end, *score, *from; /* per flat cell; from = flat index or -1 */
    int *row;                    /* dp row (distinct-end index) of each flat cell */
    int n_chain;
    int *chain_off;              /* n_chain+1 */
    int *cells;                  /* flat cell ids */
    int *est_start, *est_period; /* per chain */
    int64_t n_evals;             /* get_con_score evaluations (the roofline unit for chaining) */
} tho_chain_t;
/* src/tandem_chain.c:290-403 */
int tho_tandem_chain(const uint64_t *hits, int hit_n, const tho_para_t *p, tho_chain_t *out);
void tho_chain_free(tho_chain_t *c);

/* src/partition.c:171-276; returns par_n, *par_pos malloc'ed */
int tho_partition(const uint8_t *bseq, int len, const tho_chain_t *c, int chain_i, const tho_para_t *p, int **par_pos);

/* ksw2 restatement (ksw2/ksw2_extz2_sse.c:23-304, ksw2/ksw2.h:119-176, src/ksw2_align.c) */
int  tho_ksw2_global(const uint8_t *q, int ql, const uint8_t *t, int tl, int *n_cigar, uint32_t **cigar); /* returns iden_n; cigar optional */
void tho_ksw2_ext(const uint8_t *q, int ql, const uint8_t *t, int tl, int *max_q, int *max_t);        /* ksw2_right_ext */
void tho_ksw2_left_ext(const uint8_t *q, int ql, const uint8_t *t, int tl, int *max_q, int *max_t);   /* ksw2_left_ext */
int  tho_ksw2_backtrack_left_end(int n_cigar, const uint32_t *cigar, int qlen, int tlen, int q_left_ext);

/* abPOA restatement: consensus of n_seqs sequences (global, convex gap, adaptive band, HC consensus).
 * Returns cons_len (0 if none); cons/cov caller-allocated with >= sum(lens)+2 entries.
 * poa_cells (optional) accumulates the banded DP cell count. */
int tho_abpoa_cons(const tho_para_t *p, int n_seqs, const uint8_t *const *seqs, const int *lens,
                   uint8_t *cons, int *cov, int64_t *poa_cells);

/* ---- whole path ---- */
typedef struct {
    int cons_start, cons_end, cons_len;
    double copy_num, ave_match;
    int full_length;
    int pos_n; int *sub_pos;
    char *cons_seq;            /* cons_len chars (not NUL terminated) or NULL (-u) */
    char *cons_qual;           /* cons_len chars or NULL */
    /* raw integer results behind the FP fields (what the GPU path returns) */
    int n_seqs, raw_cons_len;
    int *iden_n;               /* pos_n-1 */
    int lext_q, lext_t, rext_q, rext_t;
} tho_cons_t;

typedef struct {
    int n_cons, m_cons;
    tho_cons_t *cons;
    /* work counters for roofline numerators */
    int64_t n_hits, n_chain_evals, n_poa_cells, n_ksw_cells;
    /* qualities of records that -l replaced by a longer one: write_tandem_cons_seq rewinds seq.l but not qual.l
     * (src/gen_cons.c:12-16 vs :23-30), so they stay in the slot's quality buffer in front of the kept record's */
    char *dropped_qual; size_t dropped_l;
} tho_read_t;

/* src/tidehunter.c:23-60 */
void tho_process_read(const char *seq, int len, const tho_para_t *p, tho_read_t *out);
void tho_read_free(tho_read_t *r);

/* src/main.c:214-271 -- formats one read's records exactly like mini_tandem_output; appends to *buf */
void tho_format_read(const char *name, const char *seq, int len, const tho_read_t *r, const tho_para_t *p,
                     char **buf, size_t *buf_l, size_t *buf_m);

/* convenience for ctypes: run n reads (threads >= 1) and return the concatenated output text */
char *tho_run_batch(int n, const char *const *names, const char *const *seqs, const int *lens,
                    const tho_para_t *p, int n_threads, size_t *out_len, int64_t counters[4]);
void tho_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
