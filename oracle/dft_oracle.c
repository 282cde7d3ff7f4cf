/* dft_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference sonde-type detector scan/dft_detect.c in its production mode
 * (`dft_detect [-c] [-v] [-t s] --iq --bw <kHz> --dc - <sr> <8|16|32>`, auto_rx/autorx/scan.py:541-547):
 * IQ DC removal (dft_detect.c:556-591), IF low-pass (:665-716, :1131-1151), FM discriminator (:780-811),
 * matched-filter bank (:1175-1266), FFT correlator getCorrDFT (:357-443), header bit check (:821-901),
 * M10/M20 and iMet disambiguation and the main decision loop (:1492-1666).
 * Structure is ours (state struct instead of file-scope statics, in-memory stream instead of stdin, results
 * as records instead of printf); the arithmetic follows the reference.  The reference is built -Ofast, so
 * even CPU<->CPU float sums are not reproducible: this file is pinned against oracle/_ref/dft_detect on
 * detected type, sign, sample position (exact) and score (1e-3 abs) by tests/test_dft_oracle_vs_ref.py.
 * The legacy FM-audio/WAV input and the --IQ baseband decimator are not restated (not on the auto_rx IQ path).
 */
#include "oracle.h"
#include <math.h>
#include <complex.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define NRS 18
#define IDX_IMETAFSK 15
#define IDX_RS 16
#define IDX_I4 17
#define N_BWIQ 4
#define FM_GAIN 0.8

typedef struct { int sps; const char *header; float BT; float thres; int herrs; const char *type; int tn; int lpFM; int lpIQ; } hdrdef_t;

/* the header bank, scan/dft_detect.c:43-126 (strings) and :177-194 (table) */
static const hdrdef_t HDR[NRS] = {
    {2500, "10011010100110010101101001010101", 1.0f, 0.65f, 2, "DFM9", 2, 0, 1},
    {4800, "0000100001101101010100111000100001000100011010010100100000011111", 0.5f, 0.70f, 2, "RS41", 3, 0, 1},
    {4800, "10100110011001101001" "10100110011001101001" "1010011001100110100110101010100110101001", 0.5f, 0.70f, 3, "RS92", 4, 0, 1},
    {4800, "0101011000001000" "0001110010010111" "0001101010100111" "0011110100111110", 1.0f, 0.60f, 8, "LMS6", 8, 0, 1},
    {4800, "0000000001" "0101010101" "0001001001" "0001001001", 0.5f, 0.80f, 2, "IMET5", 24, 0, 1},
    {9616, "0010100111" "0010100111" "0001001001" "0010010101", 1.0f, 0.70f, 2, "MK2LMS", 18, 1, 2},
    {9608, "1001100110010100110010011001" "1010", 1.0f, 0.76f, 2, "M10", 5, 1, 2},
    {2400, "110011001101001101001101010100101010110010101010", 1.0f, 0.70f, 2, "MEISEI", 9, 0, 2},
    {4800, "10100110010110101001" "10010101011010010101" "10101001010101010101" "10011001010110101001", 1.0f, 0.70f, 2, "RD94RD41", 10, 0, 1},
    {2400, "1001100110011001" "1001101010101010", 1.5f, 0.80f, 2, "MRZ", 12, 0, 1},
    {1200, "10101010" "10101010" "10110100" "00101011", 1.0f, 0.65f, 2, "MTS01", 13, 0, 0},
    {5800, "01010101010101010101010101010101", 1.5f, 0.80f, 2, "C34C50", 15, 0, 2},
    {4800, "10101010" "10101010" "10101010" "00101101" "11010100", 1.0f, 0.65f, 2, "WXR301", 16, 0, 3},
    {5000, "10101010" "10101010" "10101010" "11000001" "10010100", 1.0f, 0.65f, 2, "WXRPN9", 17, 0, 3},
    {9600, "0000" "11110000111100001111000011110000" "1111" "0000" "10101100110010101100101010101100" "1111", 1.0f, 0.80f, 2, "IMET1AB", 29, 1, 3},
    {9600, "11110000111100001111000011110000" "11110000111100001111000011110000", 0.5f, 0.80f, 4, "IMETafsk", 25, 1, 1},
    {9600, "0000" "1111" "0000" "1111" "0000" "1111" "0000" "1111", 0.5f, 0.80f, 2, "IMET1RS", 28, 0, 3},
    {9600, "0000" "1111" "0000" "1111" "0000" "1111" "0000" "1111", 0.5f, 0.80f, 2, "IMET4", 26, 1, 1},
};
#define SKIP_A 11   /* C34C50  (-DNOC34C50, scan/Makefile:1)  */
#define SKIP_B 14   /* IMET1AB (-DNOIMET1AB, scan/Makefile:1) */

typedef struct {
    int sr, bits, opt_dc, opt_cont, singleLp;
    double set_lpIQ;
    int N, K, L, M, LOG2N, delay, lpFMtaps, lpIQtaps, hLenMax;
    float spb[NRS]; int hLen[NRS], Lj[NRS]; const char *type[NRS]; int tn[NRS]; float dc[NRS], df[NRS];
    float *ws_lpFM[2], *ws_lpIQ[3];
    float complex *Fm[IDX_RS], *WS[2], *ew, *X, *Z, *cx, *lpIQ_buf;
    float *xn, *db, *buf_fm[N_BWIQ];
    char *rawbits;
    unsigned sample_in, sample_out;
    /* IQ dc */
    double sumx, sumy; float avgx, avgy; unsigned cnt, maxcnt;
    float complex z0_fm0, z0_fm1, z0_fm2, z0;
    /* read_bufbit statics */
    unsigned rcount; float rbitgrenze;
    /* input */
    const unsigned char *in; long nsamp, rd;
    /* --IQ <fq> baseband translate + decimate (dft_detect.c:751-761, 1021-1111) */
    int wide, sr_base, decM, dectaps, lut_len; unsigned sample_decX, sample_decM; double xlt_fq;
    float *ws_dec; float complex *ex, *decX, *decMbuf;
    float complex *if_out; long if_cap, if_n;                      /* optional copy of the decimated IF stream */
} dfto_t;

static void fft_raw(dfto_t *s, float complex *Z)                       /* dft_detect.c:295-329 */
{
    int N = s->N, i, j = 1, k, st, l, l2;
    for (i = 1; i < N; i++) {
        if (i < j) { float complex T = Z[j - 1]; Z[j - 1] = Z[i - 1]; Z[i - 1] = T; }
        k = N / 2;
        while (k < j) { j -= k; k /= 2; }
        j += k;
    }
    for (st = 0; st < s->LOG2N; st++) {
        l2 = 1 << st; l = l2 << 1;
        float complex w1 = 1.0f, w2 = s->ew[st];
        for (j = 1; j <= l2; j++) {
            for (i = j; i <= N; i += l) {
                k = i + l2;
                float complex T = Z[k - 1] * w1;
                Z[k - 1] = Z[i - 1] - T;
                Z[i - 1] = Z[i - 1] + T;
            }
            w1 = w1 * w2;
        }
    }
}
static void fft_real(dfto_t *s, const float *x, float complex *Z) { for (int i = 0; i < s->N; i++) Z[i] = x[i]; fft_raw(s, Z); }
static void nidft(dfto_t *s, const float complex *Zi, float complex *z) { for (int i = 0; i < s->N; i++) z[i] = conjf(Zi[i]); fft_raw(s, z); }

static double sinc_(double x) { return x == 0 ? 1 : sin(M_PI * x) / (M_PI * x); }
static int lowpass_init(float f, int taps, float **pws)                 /* dft_detect.c:665-697 */
{
    if (taps % 2 == 0) taps++;
    if (taps < 1) taps = 1;
    float *ws = (float *)calloc(2 * taps + 1, sizeof(float));
    double norm = 0;
    for (int n = 0; n < taps; n++) {
        double w = 7938 / 18608.0 - 9240 / 18608.0 * cos(2 * M_PI * n / (taps - 1)) + 1430 / 18608.0 * cos(4 * M_PI * n / (taps - 1));
        double h = 2 * f * sinc_(2 * f * (n - (taps - 1) / 2));
        ws[n] = w * h;
        norm += ws[n];
    }
    for (int n = 0; n < taps; n++) ws[n] /= norm;
    for (int n = 0; n < taps; n++) ws[taps + n] = ws[n];
    *pws = ws;
    return taps;
}
static float complex lowpass(const float complex *buffer, unsigned sample, unsigned taps, const float *ws)   /* :709-716 */
{
    float complex w = 0;
    int S = taps - (sample % taps);
    for (unsigned n = 0; n < taps; n++) w += buffer[n] * ws[S + n];
    return w;
}
static double Qf(double x) { return 0.5 - 0.5 * erf(x / 1.4142135624); }   /* :977-982 */
static double pulse(double t, double sigma) { return Qf((t - 0.5) / sigma) - Qf((t + 0.5) / sigma); }

static int init(dfto_t *s)                                              /* dft_detect.c:995-1285, --iq branch */
{
    int taps, j, i;
    float lpIQ_bw[3] = {6e3f, 12e3f, 22e3f}, lpFM_bw[2] = {4e3f, 10e3f};
    s->decM = 1;
    if (s->wide) {                                                      /* :1021-1068 */
        int IF_sr = 48000, decM = 1, sr_base = s->sr;
        if (s->set_lpIQ > IF_sr) IF_sr = s->set_lpIQ;
        int wideIF = IF_sr > 60e3;
        if (IF_sr > sr_base) IF_sr = sr_base;
        if (IF_sr < sr_base) { while (sr_base % IF_sr) IF_sr += 1; decM = sr_base / IF_sr; }
        float f_lp = (IF_sr + 20e3) / (4.0 * sr_base), t_bw = (IF_sr - 20e3);
        if (wideIF) { f_lp = (IF_sr + 60e3) / (4.0 * sr_base); t_bw = (IF_sr - 60e3); }
        if (t_bw < 0) t_bw = 10e3;
        t_bw /= sr_base;
        taps = 4.0 / t_bw; if (taps % 2 == 0) taps++;
        s->dectaps = lowpass_init(f_lp, taps, &s->ws_dec);
        s->sr_base = sr_base; s->sr = IF_sr; s->decM = decM;
        /* rotation look-up table (:1070-1101) */
        int W = 2 * 8, d, k, freq = (int)(s->xlt_fq * (double)sr_base + 0.5), freq0 = freq;
        for (d = W; d > 0; d--) if (sr_base % d == 0) break;
        if (d == 0) d = 1;
        for (k = 0; k < W / 2; k++) {
            if ((freq + k) % d == 0) { freq0 = freq + k; break; }
            if ((freq - k) % d == 0) { freq0 = freq - k; break; }
        }
        s->lut_len = sr_base / d;
        double f0 = freq0 / (double)sr_base;
        s->ex = (float complex *)calloc(s->lut_len + 1, sizeof(float complex));
        for (int n = 0; n < s->lut_len; n++) { double t = f0 * (double)n; s->ex[n] = cexp(t * 2 * M_PI * I); }
        s->decX = (float complex *)calloc(s->dectaps + 1, sizeof(float complex));
        s->decMbuf = (float complex *)calloc(s->decM + 1, sizeof(float complex));
    }
    int sr = s->sr;
    taps = 4 * sr / 2e3; if (taps % 2 == 0) taps++;
    taps = lowpass_init(lpFM_bw[0] / (float)sr, taps, &s->ws_lpFM[0]);
    taps = lowpass_init(lpFM_bw[1] / (float)sr, taps, &s->ws_lpFM[1]);
    s->lpFMtaps = taps;
    taps = 4 * sr / 4e3; if (taps % 2 == 0) taps++;
    if (s->set_lpIQ > 100.0) { lpIQ_bw[0] = lpIQ_bw[1] = lpIQ_bw[2] = s->set_lpIQ; s->singleLp = 1; }
    for (j = 0; j < 3; j++) taps = lowpass_init(lpIQ_bw[j] / (float)sr / 2.0, taps, &s->ws_lpIQ[j]);
    s->lpIQtaps = taps;
    s->lpIQ_buf = (float complex *)calloc(taps + 3, sizeof(float complex));
    s->maxcnt = sr / 32;
    if (s->decM > 1) s->maxcnt *= s->decM;                              /* :1156 */
    int hLen = 0, Lmax = 0;
    for (j = 0; j < NRS; j++) {
        s->type[j] = HDR[j].type; s->tn[j] = HDR[j].tn;
        s->spb[j] = sr / (float)HDR[j].sps;
        s->hLen[j] = strlen(HDR[j].header);
        s->Lj[j] = s->hLen[j] * s->spb[j] + 0.5;
        if (j != SKIP_A && j != SKIP_B) { if (s->hLen[j] > hLen) hLen = s->hLen[j]; if (s->Lj[j] > Lmax) Lmax = s->Lj[j]; }
    }
    s->hLenMax = hLen;
    int L = 2 * Lmax, M = 3 * L, p2 = 1;
    while (p2 < M) p2 <<= 1;
    while (p2 < 0x2000) p2 <<= 1;
    s->N = p2; s->K = p2 - L; s->L = L;
    s->LOG2N = log(p2) / log(2) + 0.1;
    s->delay = L / 16; s->M = p2 + s->delay + 8;
    s->rawbits = (char *)calloc(hLen + 1, 1);
    for (j = 0; j < N_BWIQ; j++) s->buf_fm[j] = (float *)calloc(s->M + 1, sizeof(float));
    int N = s->N;
    s->xn = (float *)calloc(N + 1, sizeof(float)); s->db = (float *)calloc(N + 1, sizeof(float));
    s->ew = (float complex *)calloc(s->LOG2N + 1, sizeof(float complex));
    s->X = (float complex *)calloc(N + 1, sizeof(float complex)); s->Z = (float complex *)calloc(N + 1, sizeof(float complex));
    s->cx = (float complex *)calloc(N + 1, sizeof(float complex));
    for (int n = 0; n < s->LOG2N; n++) s->ew[n] = cexp(-I * M_PI / (float)(1 << n));
    float *match = (float *)calloc(L + 1, sizeof(float)), *m = (float *)calloc(N + 1, sizeof(float));
    for (j = 0; j < IDX_RS; j++) {                                        /* :1223-1266 */
        s->Fm[j] = (float complex *)calloc(N + 1, sizeof(float complex));
        const char *bits = HDR[j].header; float spb = s->spb[j];
        double sigma = sqrt(log(2)) / (2 * M_PI * HDR[j].BT);
        for (i = 0; i < s->Lj[j]; i++) {
            int pos = i / spb;
            double t = (i - pos * spb) / spb - 0.5;
            double b = ((bits[pos] & 0x1) - 0.5) * 2.0 * pulse(t, sigma);
            if (pos > 0) b += ((bits[pos - 1] & 0x1) - 0.5) * 2.0 * pulse(t + 1, sigma);
            if (pos < hLen - 1) b += ((bits[pos + 1] & 0x1) - 0.5) * 2.0 * pulse(t - 1, sigma);   /* global hLen: the quirk of :1249 */
            match[i] = b;
        }
        double n2 = 0; for (i = 0; i < s->Lj[j]; i++) n2 += (double)match[i] * match[i];
        float nm = sqrt(n2);
        for (i = 0; i < s->Lj[j]; i++) match[i] /= nm;
        for (i = 0; i < s->Lj[j]; i++) m[s->Lj[j] - 1 - i] = match[i];
        while (i < N) m[i++] = 0.0;
        fft_real(s, m, s->Fm[j]);
    }
    for (j = 0; j < 2; j++) {
        s->WS[j] = (float complex *)calloc(N + 1, sizeof(float complex));
        for (i = 0; i < s->lpFMtaps; i++) m[i] = s->ws_lpFM[j][i];
        while (i < N) m[i++] = 0.0;
        fft_real(s, m, s->WS[j]);
    }
    free(match); free(m);
    return s->K;
}

static int read_csample(dfto_t *s, float complex *z)                    /* :556-591 */
{
    if (s->rd >= s->nsamp) return -1;
    float x, y;
    if (s->bits == 32) { const float *f = (const float *)s->in + 2 * s->rd; x = f[0]; y = f[1]; }
    else if (s->bits == 16) { const short *b = (const short *)s->in + 2 * s->rd; x = b[0] / 32768.0; y = b[1] / 32768.0; }
    else { const unsigned char *u = s->in + 2 * s->rd; x = (u[0] - 128) / 128.0; y = (u[1] - 128) / 128.0; }
    s->rd++;
    *z = (x - s->avgx) + I * (y - s->avgy);
    s->sumx += x; s->sumy += y; s->cnt += 1;
    if (s->cnt == s->maxcnt) { s->avgx = s->sumx / (float)s->maxcnt; s->avgy = s->sumy / (float)s->maxcnt; s->sumx = 0; s->sumy = 0; s->cnt = 0; }
    return 0;
}

static int read_cblock(dfto_t *s)                                      /* f32read_cblock, :593-657 */
{
    int n;
    for (n = 0; n < s->decM; n++) { if (read_csample(s, &s->decMbuf[n])) break; }
    return n;
}

static int buf_sample(dfto_t *s)                                        /* f32buf_sample, :737-819 (--iq) */
{
    float complex z, z_fm0, z_fm1, z_fm2, w; float sv[N_BWIQ]; double gain = FM_GAIN;
    if (s->wide) {                                                      /* :751-761 */
        if (read_cblock(s) < s->decM) return -1;
        for (int j = 0; j < s->decM; j++) {
            s->decX[s->sample_decX] = s->decMbuf[j] * s->ex[s->sample_decM];
            s->sample_decM += 1; if (s->sample_decM >= (unsigned)s->lut_len) s->sample_decM = 0;
            s->sample_decX += 1; if (s->sample_decX >= (unsigned)s->dectaps) s->sample_decX = 0;
        }
        z = lowpass(s->decX, s->sample_decX, s->dectaps, s->ws_dec);
        if (s->if_out && s->if_n < s->if_cap) s->if_out[s->if_n] = z;
        s->if_n++;
    }
    else if (read_csample(s, &z)) return -1;
    s->lpIQ_buf[s->sample_in % s->lpIQtaps] = z;
    z_fm0 = lowpass(s->lpIQ_buf, s->sample_in + 1, s->lpIQtaps, s->ws_lpIQ[0]);
    if (s->singleLp) { z_fm1 = z_fm0; z_fm2 = z_fm0; }
    else { z_fm1 = lowpass(s->lpIQ_buf, s->sample_in + 1, s->lpIQtaps, s->ws_lpIQ[1]); z_fm2 = lowpass(s->lpIQ_buf, s->sample_in + 1, s->lpIQtaps, s->ws_lpIQ[2]); }
    w = z_fm0 * conjf(s->z0_fm0); sv[0] = gain * carg(w) / M_PI; s->z0_fm0 = z_fm0;
    if (s->singleLp) { sv[1] = sv[0]; sv[2] = sv[0]; s->z0_fm1 = z_fm1; s->z0_fm2 = z_fm2; }
    else {
        w = z_fm1 * conjf(s->z0_fm1); sv[1] = gain * carg(w) / M_PI; s->z0_fm1 = z_fm1;
        w = z_fm2 * conjf(s->z0_fm2); sv[2] = gain * carg(w) / M_PI; s->z0_fm2 = z_fm2;
    }
    w = z * conjf(s->z0); sv[3] = gain * carg(w) / M_PI; s->z0 = z;
    for (int i = 0; i < N_BWIQ; i++) s->buf_fm[i][s->sample_in % s->M] = sv[i];
    s->sample_out = s->sample_in - s->delay;
    s->sample_in += 1;
    return 0;
}

static int get_corr(dfto_t *s, int j, float *maxv, unsigned *maxvpos)    /* getCorrDFT, :357-443 */
{
    int K = s->K, N = s->N, Lj = s->Lj[j], M = s->M, i, mp = -1;
    float mx = 0, mx2 = 0, dc = 0;
    s->dc[j] = 0;
    if (K + Lj > N) return -1;
    unsigned pos = s->sample_out;
    const float *bufs = s->buf_fm[HDR[j].lpIQ];
    for (i = 0; i < K + Lj; i++) s->xn[i] = bufs[(pos + M - (K + Lj - 1) + i) % M];
    while (i < N) s->xn[i++] = 0.0;
    fft_real(s, s->xn, s->X);
    if (s->opt_dc) {
        for (i = K - Lj; i < K + Lj; i++) dc += s->xn[i];
        dc /= 2.0 * (float)Lj;
        s->X[0] -= N * dc * 0.98;
    }
    s->dc[j] = dc;
    for (i = 0; i < N; i++) s->X[i] *= s->WS[HDR[j].lpFM][i];
    nidft(s, s->X, s->cx);
    for (i = 0; i < N; i++) s->xn[i] = crealf(s->cx[i]) / (float)N;
    for (i = 0; i < N; i++) s->Z[i] = s->X[i] * s->Fm[j][i];
    nidft(s, s->Z, s->cx);
    for (i = Lj - 1; i < K + Lj; i++) {
        float re = crealf(s->cx[i]);
        if (re * re > mx2) { mx = re; mx2 = mx * mx; mp = i; }
    }
    if (mp == Lj - 1 || mp == K + Lj - 1) return -4;
    unsigned mpos = pos - (K + Lj - 1) + mp;
    double xnorm = 0;
    for (i = 0; i < Lj; i++) xnorm += s->xn[mp - i] * s->xn[mp - i];
    xnorm = sqrt(xnorm);
    mx /= xnorm * N;
    mpos -= s->lpFMtaps / 2;
    *maxv = mx; *maxvpos = mpos;
    if (s->opt_dc) s->df[j] = s->dc[j] / (2.0 * FM_GAIN * s->decM);     /* :439 */
    return mp;
}

static void read_bufbit(dfto_t *s, int symlen, char *bits, unsigned mvp, int reset, float dc, int j)   /* :821-864 */
{
    double sum = 0.0; const float *bufs = s->buf_fm[HDR[j].lpIQ]; int M = s->M;
    if (reset) { s->rcount = 0; s->rbitgrenze = 0; }
    s->rbitgrenze += s->spb[j];
    do { sum += bufs[(s->rcount + mvp + M) % M] - dc; s->rcount++; } while (s->rcount < s->rbitgrenze);
    if (symlen == 2) {
        s->rbitgrenze += s->spb[j];
        do { sum -= bufs[(s->rcount + mvp + M) % M] - dc; s->rcount++; } while (s->rcount < s->rbitgrenze);
    }
    if (symlen != 2) *bits = sum >= 0 ? '1' : '0';
    else { bits[0] = sum >= 0 ? '1' : '0'; bits[1] = sum >= 0 ? '0' : '1'; }
}
static int headcmp(dfto_t *s, unsigned mvp, int inv, int j)               /* :866-901, symlen 1 */
{
    int errs = 0, len = s->hLen[j], pos; char sign = inv ? 1 : 0;
    float dc = s->opt_dc ? s->dc[j] : 0.0f;
    for (pos = 0; pos < len; pos++) read_bufbit(s, 1, s->rawbits + pos, mvp + 1 - (int)(s->hLen[j] * s->spb[j]), pos == 0, dc, j);
    s->rawbits[pos] = 0;
    while (len > 0) { if ((s->rawbits[len - 1] ^ sign) != HDR[j].header[len - 1]) errs++; len--; }
    return errs;
}
static unsigned frm_M10(dfto_t *s, unsigned mvp, int inv, int j)          /* :932-965 */
{
    float dc = s->opt_dc ? s->dc[j] : 0.0f; char bit0, mb[2], frmbit[17]; int pos2;
    int ofs = ((int)strlen(HDR[j].header) - 28) / 2;
    if (ofs < 0 || ofs > 8) ofs = 0;
    bit0 = 0x30 + (inv > 0);
    for (pos2 = 0; pos2 < 16; pos2++) {
        if (pos2 < ofs) mb[0] = HDR[j].header[28 + 2 * pos2] ^ (inv > 0);
        else read_bufbit(s, 2, mb, mvp, pos2 == ofs, dc, j);
        frmbit[pos2] = 0x31 ^ (bit0 ^ mb[0]);
        bit0 = mb[0];
    }
    unsigned b0 = 0, b1 = 0;
    for (int i = 0; i < 8; i++) { b0 = (b0 << 1) | (frmbit[i] == '1'); b1 = (b1 << 1) | (frmbit[8 + i] == '1'); }
    return (b0 << 8) | b1;
}
static int hw(unsigned b) { int d = 0; for (int i = 0; i < 8; i++) { d += b & 1; b >>= 1; } return d; }

/* Whole run of the detector over an in-memory IQ stream.  Returns the process exit code the reference would
 * return (header_found * tn, :1660-1666); detections (the lines the reference prints) go to dets[]. */
static int run_core(dfto_t *s, const void *iq, long nsamp, int bits, int sample_rate, double bw_khz, int opt_dc, int opt_cont, double tl,
                    dfto_det_t *dets, int maxdets, int *ndets, dfto_blk_t *blks, int maxblks, int *nblks);

int dfto_run(const void *iq, long nsamp, int bits, int sample_rate, double bw_khz, int opt_dc, int opt_cont, double tl,
             dfto_det_t *dets, int maxdets, int *ndets, dfto_blk_t *blks, int maxblks, int *nblks)
{
    dfto_t S; memset(&S, 0, sizeof S);
    return run_core(&S, iq, nsamp, bits, sample_rate, bw_khz, opt_dc, opt_cont, tl, dets, maxdets, ndets, blks, maxblks, nblks);
}

/* `dft_detect [-c] [-t tl] --IQ fq --bw bw_khz [--dc] - sr_base bits`: translate by -fq*sr_base, decimate to the IF rate,
 * then the same detector (dft_detect.c:1383-1392).  if_out (may be NULL) receives the decimated IF samples (re,im pairs). */
int dfto_run_wideband(const void *iq, long nsamp, int bits, int sr_base, double fq, double bw_khz, int opt_dc, int opt_cont, double tl,
                      dfto_det_t *dets, int maxdets, int *ndets, float *if_out, long if_cap, long *if_n, int *if_sr, int *decM, int *dectaps)
{
    dfto_t S; memset(&S, 0, sizeof S);
    if (fq < -0.5) fq = -0.5;
    if (fq > 0.5) fq = 0.5;
    S.wide = 1; S.xlt_fq = -fq; S.if_out = (float complex *)if_out; S.if_cap = if_cap;
    int rc = run_core(&S, iq, nsamp, bits, sr_base, bw_khz, opt_dc, opt_cont, tl, dets, maxdets, ndets, NULL, 0, NULL);
    if (if_n) *if_n = S.if_n;
    if (if_sr) *if_sr = S.sr;
    if (decM) *decM = S.decM;
    if (dectaps) *dectaps = S.dectaps;
    return rc;
}

static int run_core(dfto_t *s, const void *iq, long nsamp, int bits, int sample_rate, double bw_khz, int opt_dc, int opt_cont, double tl,
                    dfto_det_t *dets, int maxdets, int *ndets, dfto_blk_t *blks, int maxblks, int *nblks)
{
    s->sr = sample_rate; s->bits = bits; s->opt_dc = opt_dc; s->opt_cont = opt_cont;
    if (bw_khz < 1.0) bw_khz = 0.0;
    s->set_lpIQ = bw_khz * 1e3; s->in = (const unsigned char *)iq; s->nsamp = nsamp;
    int K = init(s), j, k = 0, nd = 0, nb = 0;
    float mv[NRS]; unsigned mv_pos[NRS], mv0_pos[NRS]; int mp[NRS];
    for (j = 0; j < NRS; j++) { mv[j] = 0; mv_pos[j] = 0; mp[j] = 0; }
    int j_max = 0, header_found = 0; float mv_max = 0; unsigned frm2 = 0;
    while (buf_sample(s) == 0) {
        if (tl > 0 && s->sample_in > (tl + 1) * s->sr) break;
        k++;
        if (k < K - 4) continue;
        for (j = 0; j <= IDX_IMETAFSK; j++) {
            if (j == SKIP_A || j == SKIP_B) continue;
            mv0_pos[j] = mv_pos[j];
            mp[j] = get_corr(s, j, mv + j, mv_pos + j);
            if (blks && nb < maxblks) { blks[nb].hdr = j; blks[nb].mp = mp[j]; blks[nb].score = mv[j]; blks[nb].pos = mv_pos[j]; blks[nb].dc = s->dc[j]; blks[nb].sample_in = s->sample_in; nb++; }
        }
        k = 0;
        header_found = 0;
        for (j = 0; j <= IDX_IMETAFSK; j++) {
            if (!(mp[j] > 0 && (mv[j] > HDR[j].thres || mv[j] < -HDR[j].thres))) continue;
            if (!(mv_pos[j] > mv0_pos[j])) continue;
            int herrs = headcmp(s, mv_pos[j], mv[j] < 0, j);
            if (herrs >= HDR[j].herrs) continue;
            int jj = j;
            if (!strncmp(s->type[j], "M10", 3) || !strncmp(s->type[j], "M20", 3)) {                 /* :1526-1540 */
                unsigned bytes = frm_M10(s, mv_pos[j], mv[j] < 0, j);
                int h = hw(bytes & 0x0F);
                if (h < 2 || (h == 2 && (bytes & 0xF0) == 0x20)) { s->type[j] = "M20"; s->tn[j] = 6; }
                else { s->type[j] = "M10"; s->tn[j] = 5; }
                frm2 = bytes;
            }
            if (!strncmp(s->type[j], "IMETafsk", 8)) {                                                /* :1542-1607 */
                int n, m, N = s->N, D = N / 2 - 3; float pow2200 = 0, pow2400 = 0;
                for (n = 0; n < N; n++) { s->xn[n] = 0; s->db[n] = 0; }
                n = 0;
                while (n < s->sr) {
                    if (buf_sample(s)) break;
                    s->xn[n % D] = s->buf_fm[HDR[j].lpIQ][s->sample_out % s->M];
                    n++;
                    if (n % D == 0) { fft_real(s, s->xn, s->X); for (m = 0; m < N; m++) s->db[m] += cabsf(s->X[m]); }
                }
                float dfb = (1 / (float)N) * s->sr;
                m = 50.0 / dfb; if (m < 1) m = 1;
                if (2500 * N / (float)s->sr > N / 2) goto ende;
                int b22 = 2200 * N / (float)s->sr, b24 = 2400 * N / (float)s->sr;
                for (n = 0; n < m; n++) pow2200 += s->db[b22 - m / 4 + n];
                for (n = 0; n < m; n++) pow2400 += s->db[b24 - m / 4 + n];
                mv[j] = fabsf(mv[j]);
                if (pow2200 > pow2400) {
                    int b8 = 800 * N / (float)s->sr; float pow800 = 0;
                    for (n = 0; n < m; n++) pow800 += s->db[b8 - m / 4 + n];
                    if (pow2200 > pow800) {
                        int j0 = j;
                        jj = (s->set_lpIQ > 50e3) ? IDX_RS : IDX_I4;
                        mv[jj] = mv[j0]; mv_pos[jj] = mv_pos[j0]; s->dc[jj] = s->dc[j0]; s->df[jj] = s->df[j0]; mv[j0] = 0;
                        header_found = 1;
                    } else mv[j] = 0;
                } else mv[j] = 0;
            } else header_found = 1;
            if (header_found) {
                if (mv[jj] > HDR[jj].thres || mv[jj] < -HDR[jj].thres) {
                    if (dets && nd < maxdets) {
                        dfto_det_t *d = &dets[nd];
                        d->hdr = jj; d->tn = s->tn[jj]; d->score = mv[jj]; d->pos = mv_pos[jj]; d->df_hz = s->df[jj] * (s->wide ? s->sr_base : s->sr) /* df*sr_base, :1628 */; d->m10 = frm2 & 0xFFFF;
                        strncpy(d->type, s->type[jj], sizeof d->type - 1); d->type[sizeof d->type - 1] = 0;
                    }
                    nd++;
                    if (!strncmp(s->type[jj], "M10", 3) || !strncmp(s->type[jj], "M20", 3)) frm2 = 0;
                }
                if (fabsf(mv_max) < fabsf(mv[jj])) { mv_max = mv[jj]; j_max = jj; }
            }
            /* the reference's loop variable j itself is reassigned on an iMet relabel (:1585) and the scan resumes from there */
            j = jj;
        }
        if (header_found && !s->opt_cont) break;
        header_found = 0;
        for (j = 0; j < NRS; j++) mv[j] = 0.0;
    }
ende:
    if (ndets) *ndets = nd;
    if (nblks) *nblks = nb;
    if (mv_max) header_found = (mv_max < 0 && j_max < 3) ? -1 : 1; else header_found = 0;
    int rc = header_found * s->tn[j_max];
    for (j = 0; j < 2; j++) { free(s->ws_lpFM[j]); free(s->WS[j]); }
    for (j = 0; j < 3; j++) free(s->ws_lpIQ[j]);
    for (j = 0; j < IDX_RS; j++) free(s->Fm[j]);
    for (j = 0; j < N_BWIQ; j++) free(s->buf_fm[j]);
    free(s->ws_dec); free(s->ex); free(s->decX); free(s->decMbuf);
    free(s->lpIQ_buf); free(s->rawbits); free(s->xn); free(s->db); free(s->ew); free(s->X); free(s->Z); free(s->cx);
    return rc;
}
