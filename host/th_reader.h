/* th_reader.h -- FASTA/FASTQ(.gz) batch reader for the command line front end.
 *
 * Record semantics are those of the reference's reader (kseq_read, /root/reference/src/kseq.h:175-217, driven by
 * mini_tandem_read_seq, src/main.c:173-182), restated over whole buffer spans instead of per-character calls:
 *   - a record starts at the next '>' or '@' (anything before the first one is skipped);
 *   - name = bytes up to the first isspace() character; the rest of that line is the comment (dropped);
 *   - sequence = the following lines, copied verbatim (newlines removed, empty lines skipped) until a line that
 *     starts with '>', '@' or '+'; after each line, a trailing '\r' of the ACCUMULATED sequence is dropped when
 *     the sequence is longer than one byte (kseq.h:133);
 *   - '+' starts a FASTQ quality block: the rest of that line is skipped, then whole lines are consumed until at
 *     least seq.l quality bytes were read; a record whose quality is missing or of a different length is an error
 *     (kseq returns -2): the reference stops filling the current chunk there and drops the record.
 * Reads are returned in batches; all bytes of a batch live in two arenas (no per-read malloc), and a batch can be
 * parsed by a reader thread while the GPU works on the previous one (host/th_main.c).
 */
#ifndef TH_READER_H
#define TH_READER_H
#include <ctype.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#ifndef THR_BUFSZ
#define THR_BUFSZ (1 << 22)
#endif

typedef struct {
    gzFile fp;
    unsigned char *buf;
    int beg, end, eof;
    int last_char;      /* header character already consumed ('>' / '@'), 0 if none */
    int chunk_fill;     /* reads in the reference's current 4096-read chunk (CHUNK_READ_N, src/tidehunter.h:10) */
} th_reader;

typedef struct {
    int n, cap;
    char *sbuf; size_t sl, sm;       /* sequences, each NUL-terminated */
    char *nbuf; size_t nl, nm;       /* names, each NUL-terminated */
    size_t *soff, *noff;             /* offsets into the arenas */
    int32_t *lens;
    char **seqs, **names;            /* pointer views, valid after thr_read_batch returns */
} th_batch;

static inline int thr_fill(th_reader *r) { /* 1 if bytes are available */
    if (r->beg < r->end) return 1;
    if (r->eof) return 0;
    r->beg = 0; r->end = gzread(r->fp, r->buf, THR_BUFSZ);
    if (r->end <= 0) { r->end = 0; r->eof = 1; return 0; }
    return 1;
}
static inline int thr_getc(th_reader *r) { return thr_fill(r) ? r->buf[r->beg++] : -1; }

static inline th_reader *thr_open(const char *fn) {
    th_reader *r = (th_reader *)calloc(1, sizeof(th_reader));
    r->fp = strcmp(fn, "-") ? gzopen(fn, "r") : gzdopen(0, "r");
    if (!r->fp) { free(r); return NULL; }
    gzbuffer(r->fp, 1 << 20);
    r->buf = (unsigned char *)malloc(THR_BUFSZ);
    return r;
}
static inline void thr_close(th_reader *r) { if (r) { gzclose(r->fp); free(r->buf); free(r); } }

static inline void thr_reserve(char **s, size_t *m, size_t need) {
    if (need > *m) { size_t nm = *m ? *m : 1 << 16; while (nm < need) nm <<= 1; *s = (char *)realloc(*s, nm); *m = nm; }
}

/* appends the rest of the current line (without the newline) to *s at *l; returns the delimiter ('\n') or -1 at EOF;
 * *got = 1 if any buffer span (even an empty one) was consumed, as ks_getuntil2's `gotany` */
static inline int thr_rest_of_line(th_reader *r, char **s, size_t *l, size_t *m, int *got) {
    *got = 0;
    while (thr_fill(r)) {
        const unsigned char *p = r->buf + r->beg;
        const unsigned char *nl = (const unsigned char *)memchr(p, '\n', (size_t)(r->end - r->beg));
        const size_t k = nl ? (size_t)(nl - p) : (size_t)(r->end - r->beg);
        *got = 1;
        if (s) { thr_reserve(s, m, *l + k + 2); memcpy(*s + *l, p, k); *l += k; }
        r->beg += (int)k + (nl ? 1 : 0);
        if (nl) return '\n';
    }
    return -1;
}

/* one record into the arenas of b (not yet committed).  Returns seq length >= 0, -1 at EOF, -2 on a bad quality block */
static inline int thr_read_record(th_reader *r, th_batch *b, size_t *name_len) {
    int c, got;
    size_t l0 = b->sl, n0 = b->nl;
    if (r->last_char == 0) {
        for (;;) { /* jump to the next header line */
            if (!thr_fill(r)) return -1;
            while (r->beg < r->end && r->buf[r->beg] != '>' && r->buf[r->beg] != '@') ++r->beg;
            if (r->beg < r->end) { r->last_char = r->buf[r->beg++]; break; }
        }
    }
    /* name: up to the first white space */
    got = 0; c = -1;
    while (thr_fill(r)) {
        int i = r->beg;
        while (i < r->end && !isspace(r->buf[i])) ++i;
        thr_reserve(&b->nbuf, &b->nm, b->nl + (size_t)(i - r->beg) + 2);
        memcpy(b->nbuf + b->nl, r->buf + r->beg, (size_t)(i - r->beg)); b->nl += (size_t)(i - r->beg);
        got = 1;
        if (i < r->end) { c = r->buf[i]; r->beg = i + 1; break; }
        r->beg = i;
    }
    if (!got) { b->nl = n0; return -1; }
    *name_len = b->nl - n0;
    thr_reserve(&b->nbuf, &b->nm, b->nl + 2); b->nbuf[b->nl] = 0;
    if (c != '\n' && c != -1) thr_rest_of_line(r, NULL, NULL, NULL, &got); /* comment */
    /* sequence lines */
    while ((c = thr_getc(r)) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        thr_reserve(&b->sbuf, &b->sm, b->sl + 2);
        b->sbuf[b->sl++] = (char)c;
        thr_rest_of_line(r, &b->sbuf, &b->sl, &b->sm, &got);
        if (got && b->sl - l0 > 1 && b->sbuf[b->sl - 1] == '\r') --b->sl; /* no strip when the stream ended right after c (kseq.h:130) */
    }
    if (c == '>' || c == '@') r->last_char = c;
    thr_reserve(&b->sbuf, &b->sm, b->sl + 2); b->sbuf[b->sl] = 0;
    if (c != '+') return (int)(b->sl - l0); /* FASTA (at EOF last_char keeps the header character: the next call finds no name and ends) */
    /* FASTQ: skip the '+' line, then whole lines until enough quality bytes were seen */
    while ((c = thr_getc(r)) != -1 && c != '\n') ;
    if (c == -1) return -2;
    {
        size_t ql = 0; char *q = NULL; size_t qm = 0; const size_t sl = b->sl - l0;
        for (;;) {
            c = thr_rest_of_line(r, &q, &ql, &qm, &got);
            if (!got && r->eof && r->beg >= r->end) break;              /* ks_getuntil2 returned -1 */
            if (ql > 1 && q && q[ql - 1] == '\r') --ql;
            if (ql >= sl) break;
        }
        free(q);
        r->last_char = 0;
        if (ql != sl) return -2;
    }
    return (int)(b->sl - l0);
}

/* Fills b with up to max_reads records.  Returns the number of reads (0: end of input).  `*stop` is set when the
 * reference would stop reading for good: a bad quality block as the first record of one of its 4096-read chunks
 * (mini_tandem_read_seq returns 0 there and the main loop ends, src/main.c:402). */
static inline int thr_read_batch(th_reader *r, th_batch *b, int max_reads, int *stop) {
    int i;
    b->n = 0; b->sl = 0; b->nl = 0; *stop = 0;
    while (b->n < max_reads) {
        size_t s0 = b->sl, n0 = b->nl, name_len = 0;
        const int l = thr_read_record(r, b, &name_len);
        if (l == -1) { b->sl = s0; b->nl = n0; break; }
        if (l == -2) { /* record dropped; the reference's current chunk ends here */
            b->sl = s0; b->nl = n0;
            if (r->chunk_fill == 0) *stop = 1;
            r->chunk_fill = 0;
            if (*stop) break;
            continue; /* the next reference chunk starts right after the dropped record */
        }
        if (b->n == b->cap) {
            b->cap = b->cap ? b->cap * 2 : 4096;
            b->soff = (size_t *)realloc(b->soff, sizeof(size_t) * b->cap); b->noff = (size_t *)realloc(b->noff, sizeof(size_t) * b->cap);
            b->lens = (int32_t *)realloc(b->lens, sizeof(int32_t) * b->cap);
            b->seqs = (char **)realloc(b->seqs, sizeof(char *) * b->cap); b->names = (char **)realloc(b->names, sizeof(char *) * b->cap);
        }
        b->soff[b->n] = s0; b->noff[b->n] = n0; b->lens[b->n] = l;
        b->sl += 1; b->nl += 1; /* keep the terminators */
        ++b->n;
        if (++r->chunk_fill == 4096) r->chunk_fill = 0;
    }
    for (i = 0; i < b->n; ++i) { b->seqs[i] = b->sbuf + b->soff[i]; b->names[i] = b->nbuf + b->noff[i]; }
    return b->n;
}
static inline void thr_batch_free(th_batch *b) { free(b->sbuf); free(b->nbuf); free(b->soff); free(b->noff); free(b->lens); free(b->seqs); free(b->names); memset(b, 0, sizeof(*b)); }
#endif
