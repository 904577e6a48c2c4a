/* kseq_dump.c -- TEST INFRASTRUCTURE: prints what the reference's own reader sees in a FASTA/FASTQ(.gz) file.
 * Compiled by oracle/Makefile against the reference's kseq.h WHERE IT LIES (/root/reference/src/kseq.h, nothing is
 * copied) and driven like the reference drives it: chunks of CHUNK_READ_N = 4096 slots sharing one stream, reading
 * stops for good when a chunk comes back empty (mini_tandem_read_seq + the main loop, src/main.c:173-182, 402).
 * Output: one line per read, "<name>\t<len>\t<sequence with bytes outside 33..126 as \xHH>".  Used once, in this
 * container, to make tests/golden/reader_golden.json for host/th_reader.h. */
#include <stdio.h>
#include <stdlib.h>
#include <zlib.h>
#include "kseq.h"
KSEQ_INIT(gzFile, gzread)

static void put_escaped(const char *s, size_t l) {
    size_t i;
    for (i = 0; i < l; ++i) { unsigned char c = (unsigned char)s[i]; if (c > 32 && c < 127 && c != '\\') putchar(c); else printf("\\x%02x", c); }
}

int main(int argc, char **argv) {
    enum { CHUNK = 4096 };
    gzFile fp; kstream_t *fs; kseq_t *rs; int i, n;
    if (argc < 2) return 1;
    fp = gzopen(argv[1], "r"); if (!fp) return 1;
    fs = ks_init(fp);
    rs = (kseq_t *)calloc(CHUNK, sizeof(kseq_t));
    for (i = 0; i < CHUNK; ++i) rs[i].f = fs;
    for (;;) {
        n = 0;
        while (kseq_read(rs + n) >= 0) { ++n; if (n >= CHUNK) break; }
        if (n == 0) break;
        for (i = 0; i < n; ++i) { put_escaped(rs[i].name.s, rs[i].name.l); printf("\t%d\t", (int)rs[i].seq.l); put_escaped(rs[i].seq.s, rs[i].seq.l); putchar('\n'); }
    }
    gzclose(fp);
    return 0;
}
