/* th_reader_dump.c -- prints what host/th_reader.h reads from a FASTA/FASTQ(.gz) file, one line per read:
 * "<name>\t<len>\t<sequence with bytes outside 33..126 as \xHH>".  Built by tests/test_reader.py (gcc, no GPU) and
 * compared with the same dump made through the reference's own reader (tests/golden/reader_golden.json).
 * usage: th_reader_dump file [batch_reads] */
#include <stdio.h>
#include "th_reader.h"

static void put_escaped(const char *s, size_t l) {
    size_t i;
    for (i = 0; i < l; ++i) { unsigned char c = (unsigned char)s[i]; if (c > 32 && c < 127 && c != '\\') putchar(c); else printf("\\x%02x", c); }
}

int main(int argc, char **argv) {
    th_reader *r; th_batch b; int i, stop = 0, batch = argc > 2 ? atoi(argv[2]) : 1000;
    if (argc < 2) return 1;
    r = thr_open(argv[1]); if (!r) return 1;
    memset(&b, 0, sizeof(b));
    while (!stop && thr_read_batch(r, &b, batch, &stop) > 0)
        for (i = 0; i < b.n; ++i) { put_escaped(b.names[i], strlen(b.names[i])); printf("\t%d\t", b.lens[i]); put_escaped(b.seqs[i], (size_t)b.lens[i]); putchar('\n'); }
    thr_batch_free(&b); thr_close(r);
    return 0;
}
