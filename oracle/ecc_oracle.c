/* ecc_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the Reed-Solomon stage that follows frame assembly in rs41mod:
 *   - GF(2^8), f = 0x11D, alpha = 2 and the RS(255,231) code with b = 0, t = 12 (demod/mod/bch_ecc_mod.c:55-59, :98, :742-763)
 *   - the decoder rs_decode() = rs_decode_ErrEra() without erasures (:877-966): syndromes at alpha^0..alpha^23 (:638-651),
 *     Euclidean key-equation solver stopped at deg r < t (polyGF_lfsr :547-582), its three failure exits
 *     (deg Omega >= deg Lambda -> -3, Lambda(0) == 0 -> -2, fewer roots than deg Lambda -> -1), root search over x = 1..255,
 *     Forney values with the b == 0 factor (:596-612)
 *   - rs41_ecc() (demod/mod/rs41mod.c:1703-1976) for option.ecc 1 and 2: two interleaved codewords over the zero-padded
 *     frame (parity at 8 / 32, message bytes from 56 alternating), the second pass that rewrites the fixed packet ids and
 *     the zero block before retrying (:1738-1762; `--json`, what auto_rx runs, selects ecc = 2, :2703-2706), write-back and
 *     the return code; the erasure / known-byte passes of --ecc3/--ecc4 are NOT restated
 *   - crc16() / check_CRC() (:284-319).
 * Only decisions that are visible in the output are mirrored (which words are declared uncorrectable, what a decode beyond
 * t errors does); the polynomial bookkeeping is written independently with bounded arrays.
 * Pinned by tests/test_ecc.py against `oracle/_ref/rs41mod -r [--ecc2] --softin -i` (frame hex + [OK]/[NO] + error count).
 */
#include "oracle.h"
#include <string.h>

#define RS_N 255
#define RS_R 24
#define RS_K 231
#define RS_T 12
#define PLEN 32                       /* enough for degree <= 24 polynomials of the key equation */

static unsigned char EXP[256], LOG[256], GEN[RS_R + 1];
static int tab_ready = 0;

static void gen_tables(void)
{
    if (tab_ready) return;
    unsigned b = 1;
    for (int i = 0; i < 256; i++) {                          /* exp_a[i] = alpha^i, exp_a[255] = 1 (GF_genTab :151-154) */
        EXP[i] = (unsigned char)b;
        b <<= 1; if (b & 0x100) b ^= 0x11D;
    }
    LOG[0] = 0;
    for (int i = 0; i < 255; i++) LOG[EXP[i]] = (unsigned char)i;
    /* g(X) = prod_{i<24} (X - alpha^i)   (:753-760) */
    memset(GEN, 0, sizeof GEN); GEN[0] = 1;
    for (int i = 0; i < RS_R; i++) {
        unsigned char a = EXP[i % 255];
        for (int j = i + 1; j >= 0; j--) {
            unsigned char lo = j > 0 ? GEN[j - 1] : 0;        /* X * g */
            unsigned char hi = GEN[j] ? EXP[(LOG[GEN[j]] + LOG[a]) % 255] : 0;
            GEN[j] = lo ^ hi;
        }
    }
    tab_ready = 1;
}

static unsigned char gmul(unsigned char p, unsigned char q) { return (p && q) ? EXP[(LOG[p] + LOG[q]) % 255] : 0; }   /* :360-365 */
static unsigned char ginv(unsigned char p) { return p ? EXP[255 - LOG[p]] : 0; }                                        /* :367-370 */

static int pdeg(const unsigned char *p, int n) { int d = n - 1; while (d >= 0 && p[d] == 0) d--; return d; }           /* deg 0 = -1 (:403-408) */

static unsigned char peval(const unsigned char *p, int n, unsigned char x)                                              /* :378-390 */
{
    unsigned char y = 0;
    for (int i = n - 1; i >= 0; i--) y = gmul(y, x) ^ p[i];
    return y;
}

int ecco_rs255_encode(unsigned char cw[RS_N])                                                                           /* rs_encode :860-874 */
{
    gen_tables();
    unsigned char r[RS_N];
    memset(r, 0, RS_R); memcpy(r + RS_R, cw + RS_R, RS_K);
    for (int d = RS_N - 1; d >= RS_R; d--) {                 /* remainder of m(X) X^24 by the monic g */
        unsigned char c = r[d];
        if (!c) continue;
        for (int i = 0; i <= RS_R; i++) r[d - i] ^= gmul(GEN[RS_R - i], c);
    }
    memcpy(cw, r, RS_R);
    return 0;
}

/* returns the number of corrected symbols, or -1 / -2 / -3 with cw untouched (rs_decode_ErrEra :877-958, nera = 0) */
int ecco_rs255_decode(unsigned char cw[RS_N], unsigned char *err_pos, unsigned char *err_val)
{
    gen_tables();
    unsigned char S[PLEN]; int any = 0;
    if (err_pos) memset(err_pos, 0, RS_R);
    if (err_val) memset(err_val, 0, RS_R);
    memset(S, 0, sizeof S);
    for (int i = 0; i < RS_R; i++) { S[i] = peval(cw, RS_N, EXP[i]); if (S[i]) any = 1; }
    if (!any) return 0;
    /* key equation: remainder sequence of (X^24, S), cofactor sequence (0, 1); stop at deg r < t.  (The reference starts one step
       earlier with (S, X^24) and cofactors (1, 0); its first division has quotient 0 and only swaps the pairs, :560-572.) */
    unsigned char r0[PLEN], r1[PLEN], s0[PLEN], s1[PLEN], q[PLEN], t2[PLEN];
    memset(r0, 0, PLEN); r0[RS_R] = 1; memcpy(r1, S, PLEN);
    memset(s0, 0, PLEN); memset(s1, 0, PLEN); s1[0] = 1;
    while (pdeg(r1, PLEN) >= RS_T) {
        int dq = pdeg(r1, PLEN), dp = pdeg(r0, PLEN);
        memset(q, 0, PLEN);
        unsigned char lead = ginv(r1[dq]);
        while (dp >= dq) {                                   /* r0 <- r0 mod r1, q <- r0 div r1 */
            unsigned char c = gmul(r0[dp], lead);
            q[dp - dq] = c;
            for (int i = 0; i <= dq; i++) r0[dp - i] ^= gmul(r1[dq - i], c);
            dp = pdeg(r0, PLEN);
        }
        memcpy(t2, r0, PLEN); memcpy(r0, r1, PLEN); memcpy(r1, t2, PLEN);      /* (r0, r1) <- (r1, remainder) */
        memcpy(t2, s0, PLEN);                                                   /* s2 = q*s1 + s0 */
        int ds = pdeg(s1, PLEN), dqq = pdeg(q, PLEN);
        for (int i = 0; i <= dqq; i++) for (int j = 0; j <= ds; j++) if (i + j < PLEN) t2[i + j] ^= gmul(q[i], s1[j]);
        memcpy(s0, s1, PLEN); memcpy(s1, t2, PLEN);
    }
    unsigned char *Om = r1, *La = s1;
    int dL = pdeg(La, PLEN), dO = pdeg(Om, PLEN);
    if (dO >= dL) return -3;
    unsigned char gamma = La[0];
    if (!gamma) return -2;
    unsigned char gi = ginv(gamma);
    for (int i = 0; i <= dL; i++) La[i] = gmul(La[i], gi);
    for (int i = 0; i <= dO; i++) Om[i] = gmul(Om[i], gi);
    unsigned char dLa[PLEN]; memset(dLa, 0, PLEN);
    for (int i = 1; i <= dL; i++) if (i & 1) dLa[i - 1] = La[i];                /* formal derivative in characteristic 2 (:584-594) */
    unsigned char pos[RS_R], val[RS_R]; int nerr = 0;
    for (int x = 1; x < 256 && nerr < dL; x++) {                                /* :933-946 */
        if (peval(La, PLEN, (unsigned char)x)) continue;
        unsigned char x1 = ginv((unsigned char)x);
        pos[nerr] = LOG[x1] % 255;
        unsigned char w = peval(Om, PLEN, (unsigned char)x), z = peval(dLa, PLEN, (unsigned char)x);
        val[nerr] = z ? gmul(x1, gmul(w, ginv(z))) : 0;                         /* b == 0: Y = x^-1 * Omega(x)/Lambda'(x) (:605-606) */
        nerr++;
    }
    if (nerr < dL) return -1;
    for (int i = 0; i < nerr; i++) {
        cw[pos[i]] ^= val[i];
        if (err_pos) err_pos[i] = pos[i];
        if (err_val) err_val[i] = val[i];
    }
    return nerr;
}

int ecco_crc16(const unsigned char *data, int len)                                                                       /* rs41mod.c:284-305 */
{
    int rem = 0xFFFF;
    for (int i = 0; i < len; i++) {
        rem ^= data[i] << 8;
        for (int j = 0; j < 8; j++) rem = (rem & 0x8000) ? ((rem << 1) ^ 0x1021) & 0xFFFF : (rem << 1) & 0xFFFF;
    }
    return rem;
}

/* 0 = CRC ok, 1 = CRC wrong, -1 = not that packet / does not fit (check_CRC :307-319) */
int ecco_rs41_check_crc(const unsigned char *frame, int pos, int pck)
{
    if (((pck >> 8) & 0xFF) != frame[pos]) return -1;
    int crclen = frame[pos + 1];
    if (pos + crclen + 4 > 518) return -1;
    int crcdat = frame[pos + 2 + crclen] | (frame[pos + 3 + crclen] << 8);
    return crcdat != ecco_crc16(frame + pos + 2, crclen);
}

#define FRAME_LEN 518
#define NDATA_LEN 320
#define MSGPOS 56
#define PARPOS 8
/* rs41_ecc (rs41mod.c:1703-1976) for option.ecc = level (1 or 2).  frame is modified in place like gpx->frame. */
int ecco_rs41_ecc(unsigned char frame[FRAME_LEN], int frmlen, int level)
{
    unsigned char cw1[RS_N], cw2[RS_N];
    memset(cw1, 0, RS_N); memset(cw2, 0, RS_N);
    if (frmlen > FRAME_LEN) frmlen = FRAME_LEN;
    for (int i = frmlen; i < FRAME_LEN; i++) frame[i] = 0;
    for (int i = 0; i < RS_R; i++) { cw1[i] = frame[PARPOS + i]; cw2[i] = frame[PARPOS + RS_R + i]; }
    for (int i = 0; i < RS_K; i++) { cw1[RS_R + i] = frame[MSGPOS + 2 * i]; cw2[RS_R + i] = frame[MSGPOS + 2 * i + 1]; }
    int e1 = ecco_rs255_decode(cw1, 0, 0), e2 = ecco_rs255_decode(cw2, 0, 0);
    if (level >= 2 && (e1 < 0 || e2 < 0)) {                                     /* 2nd pass: fixed packet ids, zero block (:1738-1762) */
        static const int idpos[5] = {0x039, 0x065, 0x093, 0x0B5, 0x112};
        static const int idpck[5] = {0x7928, 0x7A2A, 0x7C1E, 0x7D59, 0x7B15};
        for (int k = 0; k < 5; k++) { frame[idpos[k]] = (idpck[k] >> 8) & 0xFF; frame[idpos[k] + 1] = idpck[k] & 0xFF; }
        int ft = 0; unsigned char b = frame[0x039 - 1];
        for (int i = 0; i < 4; i++) ft += ((b >> i) & 1) - ((b >> (i + 4)) & 1);
        if (ft < -2) { for (int i = NDATA_LEN + 7; i < FRAME_LEN - 2; i++) frame[i] = 0; }
        else {
            for (int i = NDATA_LEN; i < FRAME_LEN; i++) frame[i] = 0;
            frame[0x12B] = 0x76; frame[0x12B + 1] = 0x11;
            for (int i = 0x12B + 2; i < NDATA_LEN - 2; i++) frame[i] = 0;
            frame[NDATA_LEN - 2] = 0xEC; frame[NDATA_LEN - 1] = 0xC7;
        }
        for (int i = 0; i < RS_K; i++) { cw1[RS_R + i] = frame[MSGPOS + 2 * i]; cw2[RS_R + i] = frame[MSGPOS + 2 * i + 1]; }   /* parity part stays as pass 1 left it */
        e1 = ecco_rs255_decode(cw1, 0, 0); e2 = ecco_rs255_decode(cw2, 0, 0);
    }
    for (int i = 0; i < RS_R; i++) { frame[PARPOS + i] = cw1[i]; frame[PARPOS + RS_R + i] = cw2[i]; }
    for (int i = 0; i < RS_K; i++) { frame[MSGPOS + 2 * i] = cw1[RS_R + i]; frame[MSGPOS + 1 + 2 * i] = cw2[RS_R + i]; }
    if (e1 < 0 || e2 < 0) return -((e1 < 0 ? 1 : 0) | (e2 < 0 ? 2 : 0));
    return e1 + e2;
}
