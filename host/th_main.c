/*
 * th_main.c -- command line front end with TideHunter's options (src/main.c:16-147, 438-535), running
 * the per-read hot path on the GPU through host/th_host.c -> include/th_gpu.h.
 *
 *   tidehunter-b200 [options] in.fa/fq[.gz] > cons.fa
 *
 * Same flags and output formats as the reference; `-t` is accepted and ignored (the GPU replaces the
 * pthread pool), `--device N` selects the CUDA device, `--chunk N` the reads per GPU chunk,
 * `--lanes N` the GPU contexts the chunks rotate over (pipelining).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <getopt.h>
#include <zlib.h>
#include <time.h>
#include "th_host.h"

#define PROG "tidehunter-b200"

static const struct option long_opt[] = {
    {"kmer-length", 1, NULL, 'k'}, {"window-size", 1, NULL, 'w'}, {"HPC-kmer", 0, NULL, 'H'},
    {"min-copy", 1, NULL, 'c'}, {"max-diverg", 1, NULL, 'e'}, {"min-period", 1, NULL, 'p'}, {"max-period", 1, NULL, 'P'},
    {"match", 1, NULL, 'M'}, {"mismatch", 1, NULL, 'X'}, {"gap_open", 1, NULL, 'O'}, {"gap_ext", 1, NULL, 'E'},
    {"five-prime", 1, NULL, '5'}, {"three-prime", 1, NULL, '3'}, {"ada-match-rat", 1, NULL, 'a'},
    {"output", 1, NULL, 'o'}, {"min-len", 1, NULL, 'm'}, {"min-cov", 1, NULL, 'r'}, {"unit-seq", 0, NULL, 'u'},
    {"longest", 0, NULL, 'l'}, {"full-len", 0, NULL, 'F'}, {"single-copy", 0, NULL, 's'}, {"out-fmt", 1, NULL, 'f'},
    {"thread", 1, NULL, 't'}, {"help", 0, NULL, 'h'}, {"version", 0, NULL, 'v'},
    {"device", 1, NULL, 1001}, {"chunk", 1, NULL, 1002}, {"lanes", 1, NULL, 1003}, {"devices", 1, NULL, 1004},
    {0, 0, 0, 0}};

static long long parse_num(const char *str) { /* th_parse_num, src/main.c:54-64 */
    char *p; double x = strtod(str, &p);
    if (*p == 'G' || *p == 'g') x *= 1e9; else if (*p == 'M' || *p == 'm') x *= 1e6; else if (*p == 'K' || *p == 'k') x *= 1e3;
    return (long long)(x + .499);
}

static int usage(void) {
    fprintf(stderr, "\n%s: tandem repeat detection and consensus calling from noisy long reads (B200 GPU path)\n\n", PROG);
    fprintf(stderr, "Usage:   %s [options] in.fa/fq > cons.fa\n\n", PROG);
    fprintf(stderr, "Options (identical to TideHunter v1.5.5):\n"
                    "  -k INT  k-mer length (<=16) [8]        -w INT  minimizer window [1]       -H  HPC k-mers\n"
                    "  -c INT  min copy number (>=2) [2]      -e FLT  max divergence [0.25]\n"
                    "  -p INT  min period (>=2) [30]          -P INT  max period [10K]\n"
                    "  -M/-X INT match/mismatch [2/4]         -O INT(,INT) gap open [4,24]       -E INT(,INT) gap ext [2,1]\n"
                    "  -5/-3 STR adapter FASTA files          -a FLT  adapter match ratio [0.80]\n"
                    "  -o STR  output file [stdout]           -m INT  min consensus length [30]  -r FLT|INT min coverage\n"
                    "  -u unit sequences only   -l longest only   -F full-length only   -s single-copy full-length (with -F -5 -3)\n"
                    "  -f INT  1 FASTA, 2 tabular, 3 FASTQ, 4 tabular+quality [1]\n"
                    "  -t INT  accepted, ignored              --device INT CUDA device [0]       --chunk INT reads per GPU chunk [4096]\n          --lanes INT GPU contexts per device the chunks rotate over [4]\n          --devices LIST  comma-separated CUDA devices driven by this one process, e.g. 0,1,2,3 [--device]\n\n");
    return 1;
}

/* ---- input: host/th_reader.h parses batches on a reader thread while the GPU lanes work on the previous batch ---- */
#include <pthread.h>
#include "th_reader.h"

static char *read_first_seq(const char *fn) { /* get_seq_from_fx, src/main.c:157-171: first record with a non-empty sequence */
    th_reader *r = thr_open(fn); th_batch b; int stop = 0; char *res = NULL;
    if (!r) { fprintf(stderr, "[%s] fail to open %s\n", PROG, fn); exit(1); }
    memset(&b, 0, sizeof(b));
    if (thr_read_batch(r, &b, 1, &stop) > 0 && b.lens[0] > 0) res = strdup(b.seqs[0]);
    thr_batch_free(&b); thr_close(r);
    if (!res) { fprintf(stderr, "[%s] No sequence found in %s.\n", PROG, fn); exit(1); }
    return res;
}

typedef struct {
    th_reader *r; int batch_reads, first_batch;
    th_batch slot[2]; int state[2];       /* 0 = free, 1 = filled */
    int done;                             /* reader reached the end of the input (or the reference's stop condition) */
    pthread_mutex_t mu; pthread_cond_t cv;
} prefetch_t;

static void *reader_main(void *arg) {
    prefetch_t *q = (prefetch_t *)arg; int k = 0, stop = 0, first = 1;
    for (;;) {
        pthread_mutex_lock(&q->mu);
        while (q->state[k] != 0) pthread_cond_wait(&q->cv, &q->mu);
        pthread_mutex_unlock(&q->mu);
        { const int n = stop ? 0 : thr_read_batch(q->r, &q->slot[k], first ? q->first_batch : q->batch_reads, &stop);
          first = 0;
          pthread_mutex_lock(&q->mu);
          if (n > 0) q->state[k] = 1; else q->done = 1;
          pthread_cond_broadcast(&q->cv);
          pthread_mutex_unlock(&q->mu);
          if (n <= 0) break; }
        k ^= 1;
    }
    return NULL;
}

int main(int argc, char *argv[]) {
    long long n_failed = 0;
    th_host_para p; int c, device = 0, n_dev = 0, devs[16]; const char *out_fn = NULL, *five_fn = NULL, *three_fn = NULL; char *s;
    th_host_default_para(&p);
    if (argc < 2) return usage();
    while ((c = getopt_long(argc, argv, "k:w:m:Hhvc:e:p:P:M:X:E:O:5:3:a:o:ur:qslFf:t:", long_opt, NULL)) >= 0) {
        switch (c) {
        case 'k': p.gpu.k = atoi(optarg); break;
        case 'w': p.gpu.w = atoi(optarg); break;
        case 'H': p.gpu.hpc = 1; break;
        case 'c': p.gpu.min_copy = atoi(optarg); break;
        case 'e': p.gpu.max_div = atof(optarg); break;
        case 'p': p.gpu.min_p = parse_num(optarg); break;
        case 'P': p.gpu.max_p = parse_num(optarg); break;
        case 'M': p.gpu.match = atoi(optarg); break;
        case 'X': p.gpu.mismatch = atoi(optarg); break;
        case 'O': p.gpu.gap_open1 = (int)strtol(optarg, &s, 10); if (*s == ',') p.gpu.gap_open2 = (int)strtol(s + 1, &s, 10); break;
        case 'E': p.gpu.gap_ext1 = (int)strtol(optarg, &s, 10); if (*s == ',') p.gpu.gap_ext2 = (int)strtol(s + 1, &s, 10); break;
        case '5': five_fn = optarg; break;
        case '3': three_fn = optarg; break;
        case 'a': p.ada_match_rat = (float)atof(optarg); break;
        case 'o': out_fn = optarg; break;
        case 'm': p.min_len = atoi(optarg); break;
        case 'r': { double v = strtod(optarg, &s); /* src/main.c:492-495 */
                    if (v < 1.0) { p.min_frac = v; p.min_cov = 0; } else { p.min_cov = (int)(v + .499); p.min_frac = 0.0; }
                    break; }
        case 'u': p.gpu.only_unit = 1; break;
        case 'l': p.only_longest = 1; break;
        case 'F': p.only_full_length = 1; break;
        case 's': p.single_copy = 1; break;
        case 'f': p.out_fmt = atoi(optarg); break;
        case 't': break;
        case 'q': break;
        case 1001: device = atoi(optarg); break;
        case 1002: p.chunk_reads = atoi(optarg); if (p.chunk_reads <= 0) { fprintf(stderr, "[main] --chunk must be a positive number of reads\n"); return 1; } break;
        case 1003: p.lanes = atoi(optarg); if (p.lanes <= 0) { fprintf(stderr, "[main] --lanes must be a positive number\n"); return 1; } break;
        case 1004: { char *q = optarg; n_dev = 0; while (*q && n_dev < 16) { devs[n_dev++] = (int)strtol(q, &q, 10); if (*q == ',') ++q; else break; } break; }
        case 'v': printf("%s (TideHunter v1.5.5 compatible)\n", PROG); return 0;
        case 'h': default: return usage();
        }
    }
    /* option validation of the reference (src/main.c:505-523) */
    if (p.gpu.k > 16) { fprintf(stderr, "[main] k-mer length must be no larger than 16\n"); return 1; }
    if (p.gpu.min_copy < 2) { fprintf(stderr, "[main] min copy number must be >= 2\n"); return 1; }
    if (p.gpu.min_p < 2) { fprintf(stderr, "[main] min period must be >= 2\n"); return 1; }
    if (p.gpu.max_p > 4294967295LL) { fprintf(stderr, "[main] max period must be <= 4294967295\n"); return 1; } /* MAX_PERIOD, src/tidehunter.h:23 */
    if (p.gpu.max_p >= 65536) fprintf(stderr, "[main] note: with -P >= 65536 a read whose partition window reaches 65,536 bases is reported as failed on the GPU path (message, non-zero exit) instead of being processed\n");
    if (p.out_fmt < 1 || p.out_fmt > 4) { fprintf(stderr, "[main] unknown output format %d\n", p.out_fmt); return 1; }
    if (p.gpu.only_unit && p.out_fmt > 2) { fprintf(stderr, "[main] -u only works with -f 1/2\n"); return 1; }
    if (p.only_full_length && !(five_fn && three_fn)) { fprintf(stderr, "[main] -F needs -5 and -3\n"); return 1; }
    if (optind + 1 > argc) return usage();
    if (five_fn && three_fn) { p.five_seq = read_first_seq(five_fn); p.three_seq = read_first_seq(three_fn); }
    {
        struct timespec t0, t1; FILE *out = out_fn ? fopen(out_fn, "w") : stdout;
        th_host *h; prefetch_t q; pthread_t rt; int k = 0; long long tot_reads = 0; double t_wait = 0, t_run = 0, t_write = 0, t_create;
        struct timespec ta, tb;
#define TSPAN(a, b) ((b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec))
        clock_gettime(CLOCK_MONOTONIC, &t0);
        if (!out) { fprintf(stderr, "[main] cannot open %s\n", out_fn); return 1; }
        memset(&q, 0, sizeof(q));
        /* one th_host_run covers several chunks so that its GPU lanes overlap (host/th_host.h): four chunks per lane,
         * so the lanes drain only once per 16 chunks; the first batch is one chunk per lane to get the GPUs going early */
        q.first_batch = p.chunk_reads * (p.lanes > 0 ? p.lanes : 4) * (n_dev > 0 ? n_dev : 1);
        q.batch_reads = q.first_batch * 4;
        q.r = thr_open(argv[optind]);
        if (!q.r) { fprintf(stderr, "[main] fail to open %s\n", argv[optind]); return 1; }
        pthread_mutex_init(&q.mu, NULL); pthread_cond_init(&q.cv, NULL);
        pthread_create(&rt, NULL, reader_main, &q);   /* the first batch is parsed while the CUDA contexts come up */
        if (n_dev == 0) { devs[0] = device; n_dev = 1; }
        clock_gettime(CLOCK_MONOTONIC, &ta);
        h = th_host_create_multi(&p, n_dev, devs);
        if (!h) { fprintf(stderr, "[main] %s\n", th_host_last_error()); return 1; }
        clock_gettime(CLOCK_MONOTONIC, &tb); t_create = TSPAN(ta, tb);
        for (;;) {
            th_batch *b = &q.slot[k]; size_t ol; const char *txt; int have;
            clock_gettime(CLOCK_MONOTONIC, &ta);
            pthread_mutex_lock(&q.mu);
            while (q.state[k] == 0 && !q.done) pthread_cond_wait(&q.cv, &q.mu);
            have = q.state[k] == 1;
            pthread_mutex_unlock(&q.mu);
            clock_gettime(CLOCK_MONOTONIC, &tb); t_wait += TSPAN(ta, tb);
            if (!have) break;
            txt = th_host_run(h, b->n, (const char *const *)b->names, (const char *const *)b->seqs, b->lens, &ol);
            if (!txt) { fprintf(stderr, "[main] %s\n", th_host_last_error()); return 1; }
            clock_gettime(CLOCK_MONOTONIC, &ta); t_run += TSPAN(tb, ta);
            if (fwrite(txt, 1, ol, out) != ol) { fprintf(stderr, "[main] short write to the output (disk full or pipe closed)\n"); return 1; }
            clock_gettime(CLOCK_MONOTONIC, &tb); t_write += TSPAN(ta, tb);
            tot_reads += b->n;
            pthread_mutex_lock(&q.mu); q.state[k] = 0; pthread_cond_broadcast(&q.cv); pthread_mutex_unlock(&q.mu);
            k ^= 1;
        }
        pthread_join(rt, NULL);
        thr_batch_free(&q.slot[0]); thr_batch_free(&q.slot[1]); thr_close(q.r);
        pthread_mutex_destroy(&q.mu); pthread_cond_destroy(&q.cv);
        n_failed = th_host_failed_tasks(h);
        if (n_failed > 0) fprintf(stderr, "[main] ERROR: %lld read(s) / consensus task(s) could not run on the GPU path; their records are missing from the output\n", n_failed);
        th_host_destroy(h);
        if (fflush(out) != 0 || (out != stdout && fclose(out) != 0)) { fprintf(stderr, "[main] failed to flush the output\n"); return 1; }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "[main] Real time: %.3f sec; reads: %lld\n", (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec), tot_reads);
        if (getenv("TH_HOST_TIMING")) fprintf(stderr, "[main] contexts %.2f s, waiting for the reader %.2f s, th_host_run %.2f s, writing %.2f s\n", t_create, t_wait, t_run, t_write);
    }
    return n_failed > 0 ? 2 : 0; /* message + non-zero exit, as the reference does for what it cannot process */
}
