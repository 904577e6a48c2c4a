/* CPU model of the banded "rectangle" variant of the unit-vs-consensus identity alignment (ksw_warp_global2 in
 * tidehunter_b200/csrc/th_ksw.cuh): a scalar restatement of the kernel's recurrence with the same block geometry, fake
 * boundary values and certificate, used by tools/ksw_band_check.py to check "certificate passes => identity equals the
 * full alignment's" against the oracle and to measure pass rates.  Development aid, not part of the product. */
#include <stdint.h>
#include <stdlib.h>
#define Q 2
#define E 1
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
/* returns 1 when the certificate holds (then *iden is the full alignment's identity count); band < 0: full matrix */
int ksw_band_sim(const uint8_t *q, int ql, const uint8_t *t, int tl, int BW, int Bu, int Bl, int full, int *score, int *iden, long long *cells) {
    int *Hp = malloc(sizeof(int) * ql), *Ea = malloc(sizeof(int) * ql), *pH = malloc(sizeof(int) * ql), *pE = malloc(sizeof(int) * ql);
    int *bnd = malloc(sizeof(int) * 8 * (tl + 1));
    int nblk = (ql + BW - 1) / BW, got = 0, p_lo = 0, p_hi = 0;
    *score = 0; *iden = 0; *cells = 0;
    for (int b = 0; b < nblk; ++b) {
        const int jb = b * BW, bw = imin(ql - jb, BW);
        const int r_lo = full ? 0 : imax(0, jb - Bu), r_hi = full ? tl : imin(tl, jb + BW + Bl);
        int *bin = bnd + 4 * (tl + 1) * (b & 1), *bout = bnd + 4 * (tl + 1) * ((b + 1) & 1);
        if (r_lo >= r_hi) break;
        *cells += (long long)(r_hi - r_lo) * bw;
        for (int j = jb; j < jb + bw; ++j) {
            Hp[j] = r_lo == 0 ? -(Q + E * (j + 1)) : -(r_lo + j + 5);
            Ea[j] = Hp[j] - Q - E; pH[j] = 0; pE[j] = 0;
        }
        int hdiag, phdiag = 0;
        if (r_lo == 0) hdiag = jb == 0 ? 0 : -(Q + E * jb);
        else { hdiag = bin[4 * (r_lo - 1)]; phdiag = bin[4 * (r_lo - 1) + 2]; }
        for (int i = r_lo; i < r_hi; ++i) {
            int iH, iF, iPH = 0, iPF = 0;
            if (b == 0) { iH = -(Q + E * (i + 1)); iF = iH - Q - E; }
            else if (i < p_hi) { iH = bin[4 * i]; iF = bin[4 * i + 1]; iPH = bin[4 * i + 2]; iPF = bin[4 * i + 3]; }
            else { iH = -(i + jb + 5); iF = iH - Q - E; }
            int hd = hdiag, phd = phdiag, F = iF, pF = iPF;
            hdiag = iH; phdiag = iPH;
            for (int j = jb; j < jb + bw; ++j) {
                const int eq = q[j] == t[i];
                int z = hd + (eq ? 1 : -2), pz = phd + eq;
                const int e = Ea[j];
                if (e > z) pz = pE[j];
                z = imax(z, e);
                if (F > z) pz = pF;
                z = imax(z, F);
                const int t1 = z - Q;
                pE[j] = e > t1 ? pE[j] : pz;
                pF = F > t1 ? pF : pz;
                Ea[j] = imax(e, t1) - E;
                F = imax(F, t1) - E;
                hd = Hp[j]; phd = pH[j];
                Hp[j] = z; pH[j] = pz;
            }
            bout[4 * i] = Hp[jb + bw - 1]; bout[4 * i + 1] = F; bout[4 * i + 2] = pH[jb + bw - 1]; bout[4 * i + 3] = pF;
            if (i == tl - 1 && b == nblk - 1) { *score = Hp[ql - 1]; *iden = pH[ql - 1]; got = 1; }
        }
        p_lo = r_lo; p_hi = r_hi; (void)p_lo;
    }
    free(Hp); free(Ea); free(pH); free(pE); free(bnd);
    if (full) return got;
    if (!got) return 0;
    const int up = 2 * ql - tl - 4 - 3 * (Bu + 1), lo = 2 * tl - ql - 4 - 3 * (Bl + 1);
    return *score > imax(up, lo);
}
