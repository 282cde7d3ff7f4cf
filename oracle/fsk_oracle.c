/* fsk_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * A from-scratch plain-C restatement of the reference non-coherent 2/4-FSK demodulator
 * (radiosonde_auto_rx utils/fsk.c fsk_demod_core/fsk_demod_freq_est + utils/kiss_fft.c), written so
 * that every IEEE-754 single operation happens with the same operands and in the same order as in
 * the reference built with its own flags (-O3, no FMA contraction).  Structure is ours (iterative
 * FFT over an explicit digit-reversal permutation, staged frame pipeline, per-frame side records);
 * the arithmetic is the reference's.  Each block cites the reference lines it follows.
 *
 * Build: gcc -O2 -ffp-contract=off  (NEVER -march=native / -ffast-math).
 * Parity pin: tests/test_oracle_vs_ref.py diffs this file bit-for-bit against oracle/_ref/libfsk_ref.so.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>

typedef struct { float re, im; } cf;

struct fsko {
    /* configuration (fsk.h:48-66) */
    int Fs, Rs, Ts, M, P, Nsym, N, Nmem, Ndft, Nbits, f1_tx, fs_tx;
    int est_min, est_max, est_space, est_type;
    float tc;
    /* tables */
    float *hann;            /* fsk.c:91-98  */
    cf    *tw;              /* kiss_fft.c:357-363 */
    int   *perm;            /* leaf position -> input index (kiss_fft.c:282-296 recursion unrolled) */
    int    nstage, radix[32], mlen[32], fstride[32];
    cf    *phi_ft;          /* fine-timing oscillator sequence, input independent (fsk.c:682-698) */
    /* persistent demod state (fsk.h:69-90) */
    float *Sf;
    cf     phi_c[FSKO_MMAX];
    cf    *f_dc;            /* M*Nmem */
    float  norm_rx_timing, ppm, EbNodB, snr_est;
    float  f_est[FSKO_MMAX], f2_est[FSKO_MMAX];
    int    nin;
    /* last-frame scratch kept for inspection */
    cf    *f_int;           /* M*(Nsym+1)*P */
    cf    *fftin, *fftout;
    float *scratch;
    fsko_frame_t last;
    float  eye[FSKO_EYE_TR][FSKO_EYE_IND];
    int    neyetr, neyesamp;
};

/* ---- complex helpers: operand order of utils/comp_prim.h:57-85 ------------------------------- */
static inline cf c_mul(cf a, cf b) { cf r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }
static inline cf c_add(cf a, cf b) { cf r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }
static inline cf c_sub(cf a, cf b) { cf r; r.re = a.re - b.re; r.im = a.im - b.im; return r; }
static inline cf c_scale(float s, cf b) { cf r; r.re = s * b.re; r.im = s * b.im; return r; }
static inline cf c_conj(cf a) { cf r; r.re = a.re; r.im = -a.im; return r; }
static inline cf c_expj(float phi) { cf r; r.re = cosf(phi); r.im = sinf(phi); return r; }   /* comp_prim.h:95 */

/* ---- FFT plan: kiss_fft's factorisation (kiss_fft.c:309-331) and recursion (kiss_fft.c:237-301),
 *      flattened into (1) an input permutation and (2) a bottom-up list of butterfly stages. ------- */
static void plan_fft(struct fsko *s)
{
    int n = s->Ndft, ns = 0, p = 4;
    double floor_sqrt = floorf(sqrtf((double)n));
    int rem = n;
    do {
        while (rem % p) {
            switch (p) { case 4: p = 2; break; case 2: p = 3; break; default: p += 2; break; }
            if (p > floor_sqrt) p = rem;
        }
        rem /= p;
        s->radix[ns] = p; s->mlen[ns] = rem; ns++;
    } while (rem > 1);
    s->nstage = ns;
    int fs = 1;
    for (int i = 0; i < ns; i++) { assert(s->radix[i] == 2 || s->radix[i] == 4); s->fstride[i] = fs; fs *= s->radix[i]; }
    /* leaf position sum(k_i*m_i)  <-  input index sum(k_i*fstride_i) */
    for (int pos = 0; pos < n; pos++) {
        int r = pos, idx = 0;
        for (int i = 0; i < ns; i++) { int k = r / s->mlen[i]; r -= k * s->mlen[i]; idx += k * s->fstride[i]; }
        s->perm[pos] = idx;
    }
    for (int i = 0; i < n; i++) {
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        double phase = -2 * pi * i / n;                 /* kiss_fft.c:358-359 */
        s->tw[i].re = (float)cosf(phase);               /* _kiss_fft_guts.h:136-146 */
        s->tw[i].im = (float)sinf(phase);
    }
}

void fsko_fft(const fsko_t *s, const float *in_, float *out_)
{
    const cf *in = (const cf *)in_; cf *F = (cf *)out_;
    int n = s->Ndft;
    for (int pos = 0; pos < n; pos++) F[pos] = in[s->perm[pos]];
    for (int st = s->nstage - 1; st >= 0; st--) {
        int p = s->radix[st], m = s->mlen[st], fstr = s->fstride[st];
        for (int base = 0; base < n; base += p * m) {
            for (int u = 0; u < m; u++) {
                cf *X = F + base + u;
                if (p == 2) {                                   /* kiss_fft.c:21-42 */
                    cf t = c_mul(X[m], s->tw[fstr * u]);
                    X[m] = c_sub(X[0], t);
                    X[0] = c_add(X[0], t);
                } else {                                        /* kiss_fft.c:44-90, forward branch :82-87 */
                    cf s0 = c_mul(X[m],     s->tw[fstr * u]);
                    cf s1 = c_mul(X[2 * m], s->tw[fstr * u * 2]);
                    cf s2 = c_mul(X[3 * m], s->tw[fstr * u * 3]);
                    cf s5 = c_sub(X[0], s1);
                    cf a  = c_add(X[0], s1);
                    cf s3 = c_add(s0, s2);
                    cf s4 = c_sub(s0, s2);
                    X[2 * m] = c_sub(a, s3);
                    X[0]     = c_add(a, s3);
                    X[m].re     = s5.re + s4.im;  X[m].im     = s5.im - s4.re;
                    X[3 * m].re = s5.re - s4.im;  X[3 * m].im = s5.im + s4.re;
                }
            }
        }
    }
}

/* ---- create / config (fsk.c:114-201) ---------------------------------------------------------- */
fsko_t *fsko_create(int Fs, int Rs, int M, int P, int Nsym, int f1_tx, int fs_tx)
{
    assert(Fs > 0 && Rs > 0 && f1_tx > 0 && fs_tx > 0 && P > 0 && Nsym > 0);
    assert((Fs % Rs) == 0); assert(((Fs / Rs) % P) == 0); assert(M == 2 || M == 4);
    struct fsko *s = (struct fsko *)calloc(1, sizeof *s);
    float bin_width_Hz = 0.1 * Rs;                                   /* fsk.c:135 */
    float Ndft = (float)Fs / bin_width_Hz;                           /* fsk.c:136 */
    Ndft = pow(2.0, ceil(log2(Ndft)));                               /* fsk.c:137 */
    s->Fs = Fs; s->Rs = Rs; s->Ts = Fs / Rs; s->M = M; s->P = P; s->Nsym = Nsym;
    s->N = s->Ts * Nsym; s->Ndft = (int)Ndft; s->tc = 0.95 * Ndft / Fs;   /* fsk.c:146-148 */
    s->Nmem = s->N + 2 * s->Ts; s->f1_tx = f1_tx; s->fs_tx = fs_tx; s->nin = s->N;
    s->Nbits = (M == 2) ? Nsym : 2 * Nsym;
    s->est_min = 0; s->est_max = Fs; s->est_space = 0.75 * Rs;       /* fsk.c:156-158 */
    s->est_type = 0;   /* the reference leaves this uninitialised until fsk_set_freq_est_alg (fsk_demod.c:230) */
    int n = s->Ndft, nfi = (Nsym + 1) * P;
    s->hann = (float *)malloc(sizeof(float) * n); s->tw = (cf *)malloc(sizeof(cf) * n);
    s->perm = (int *)malloc(sizeof(int) * n); s->Sf = (float *)calloc(n, sizeof(float));
    s->scratch = (float *)calloc(n, sizeof(float));
    s->fftin = (cf *)malloc(sizeof(cf) * n); s->fftout = (cf *)malloc(sizeof(cf) * n);
    s->f_dc = (cf *)calloc((size_t)M * s->Nmem, sizeof(cf)); s->f_int = (cf *)calloc((size_t)M * nfi, sizeof(cf));
    s->phi_ft = (cf *)malloc(sizeof(cf) * nfi);
    for (int i = 0; i < n; i++)                                      /* fsk.c:96 */
        s->hann[i] = 0.5 - 0.5 * cosf(2.0 * M_PI * (float)i / (float)(n - 1));
    plan_fft(s);
    for (int m = 0; m < M; m++) s->phi_c[m] = c_expj(0);
    /* fsk.c:682-698: the timing oscillator restarts at (1,0) every frame and never sees the input */
    cf dphift = c_expj(2 * M_PI * ((float)(Rs) / (float)(P * Rs))), ph = {1, 0};
    for (int i = 0; i < nfi; i++) { s->phi_ft[i] = ph; ph = c_mul(ph, dphift); }
    /* stats_init (fsk.c:939-971) */
    int dec = ceil(((float)P * 2) / FSKO_EYE_IND);
    s->neyesamp = (P * 2) / dec; s->neyetr = M * (FSKO_EYE_TR / M);
    return s;
}

void fsko_destroy(fsko_t *s)
{
    if (!s) return;
    free(s->hann); free(s->tw); free(s->perm); free(s->Sf); free(s->scratch); free(s->fftin); free(s->fftout);
    free(s->f_dc); free(s->f_int); free(s->phi_ft); free(s);
}

void fsko_set_freq_est_limits(fsko_t *s, int lo, int hi)
{
    assert(lo >= -s->Fs / 2); assert(hi <= s->Fs / 2); assert(hi > lo);
    s->est_min = lo; s->est_max = hi;
}
void fsko_set_freq_est_alg(fsko_t *s, int t) { s->est_type = t; }
int  fsko_nin(const fsko_t *s) { return s->nin; }
int  fsko_nbits(const fsko_t *s) { return s->Nbits; }
int  fsko_ndft(const fsko_t *s) { return s->Ndft; }
int  fsko_param(const fsko_t *s, const char *k)
{
    if (!strcmp(k, "N")) return s->N;       if (!strcmp(k, "Ts")) return s->Ts;
    if (!strcmp(k, "Nmem")) return s->Nmem; if (!strcmp(k, "Ndft")) return s->Ndft;
    if (!strcmp(k, "P")) return s->P;       if (!strcmp(k, "Nsym")) return s->Nsym;
    if (!strcmp(k, "Nbits")) return s->Nbits; if (!strcmp(k, "M")) return s->M;
    return -1;
}
const float *fsko_Sf(const fsko_t *s) { return s->Sf; }
const float *fsko_hann(const fsko_t *s) { return s->hann; }
const float *fsko_twiddles(const fsko_t *s) { return (const float *)s->tw; }
const float *fsko_f_int(const fsko_t *s) { return (const float *)s->f_int; }
const float *fsko_f_dc(const fsko_t *s) { return (const float *)s->f_dc; }
const float *fsko_phi_c(const fsko_t *s) { return (const float *)s->phi_c; }
void fsko_last_frame(const fsko_t *s, fsko_frame_t *o) { *o = s->last; }
int  fsko_eye(const fsko_t *s, float eye[FSKO_EYE_TR][FSKO_EYE_IND], int *ns)
{ memcpy(eye, s->eye, sizeof s->eye); *ns = s->neyesamp; return s->neyetr; }

/* ---- tone frequency estimation (fsk.c:438-590) ------------------------------------------------ */
static void freq_est(struct fsko *s, const cf *x)
{
    int n = s->Ndft, Fs = s->Fs, nin = s->nin, M = s->M;
    int st = (s->est_min * n) / Fs + n / 2; if (st < 0) st = 0;          /* fsk.c:459 */
    int en = (s->est_max * n) / Fs + n / 2; if (en > n) en = n;          /* fsk.c:460 */
    int f_zero = (s->est_space * n) / Fs;                                /* fsk.c:463 */
    int numffts = floor((float)nin / (n / 2)) - 1;                       /* fsk.c:465 */
    float tc = s->tc;
    for (int j = 0; j < numffts; j++) {
        int a = j * n / 2;
        for (int i = 0; i < n; i++) {                                    /* fsk.c:470-478 */
            s->fftin[i].re = s->hann[i] * x[i + a].re;
            s->fftin[i].im = s->hann[i] * x[i + a].im;
        }
        fsko_fft(s, (const float *)s->fftin, (float *)s->fftout);
        for (int i = 0; i < n; i++) {                                    /* fsk.c:484-502: shift, |.|, IIR */
            cf v = s->fftout[(i + n / 2) % n];
            float mag2 = (v.re * v.re) + (v.im * v.im);
            s->Sf[i] = (s->Sf[i] * (1 - tc)) + (sqrtf(mag2) * tc);
        }
    }
    /* peak picker on a scratch copy of Sf (fsk.c:507-547).  (With numffts<=0 the reference reads
       uninitialised stack; we define the scratch as a copy of Sf.) */
    int freqi[FSKO_MMAX];
    memcpy(s->scratch, s->Sf, sizeof(float) * n);
    for (int i = 0; i < M; i++) {
        int imax = 0; float max = 0;
        for (int j = st; j < en; j++) if (s->scratch[j] > max) { max = s->scratch[j]; imax = j; }
        int f_min = imax - f_zero; f_min = f_min < 0 ? 0 : f_min;
        int f_max = imax + f_zero; f_max = f_max > n ? n : f_max;
        for (int j = f_min; j < f_max; j++) s->scratch[j] = 0;
        freqi[i] = imax - n / 2;
    }
    for (int i = 1; i < M; ) {                                           /* gnome sort, fsk.c:533-542 */
        if (freqi[i] >= freqi[i - 1]) i++;
        else { int t = freqi[i]; freqi[i] = freqi[i - 1]; freqi[i - 1] = t; if (i > 1) i--; }
    }
    for (int i = 0; i < M; i++) s->f_est[i] = (float)(freqi[i]) * ((float)Fs / (float)n);   /* fsk.c:546 */

    /* mask estimator (fsk.c:552-581) */
    float *mask = s->scratch;
    for (int i = 0; i < n; i++) mask[i] = 0.0;
    for (int i = 0; i < 3; i++) mask[i] = 1.0;
    int bin = 0;
    for (int m = 1; m <= M - 1; m++) {
        bin = round((float)m * s->fs_tx * n / Fs) - 1;                   /* fsk.c:557 */
        /* the reference iterates a size_t from bin: a negative bin means zero iterations (fsk.c:558) */
        if (bin >= 0) for (int i = bin; i <= bin + 2; i++) mask[i] = 1.0;
    }
    int len_mask = bin + 2 + 1;
    int b_max = st; float corr_max = 0.0;
    for (int b = st; b < en - len_mask; b++) {
        float corr = 0.0;
        for (int i = 0; i < len_mask; i++) corr += mask[i] * s->Sf[b + i];
        if (corr > corr_max) { corr_max = corr; b_max = b; }
    }
    float foff = (b_max - n / 2) * Fs / n;                               /* integer maths, fsk.c:578 */
    for (int m = 0; m < M; m++) s->f2_est[m] = foff + m * s->fs_tx;
}

/* ---- one modem frame (fsk.c:593-915) ---------------------------------------------------------- */
void fsko_frame(fsko_t *s, const float *iq, uint8_t *rx_bits, float *rx_sd)
{
    const cf *x = (const cf *)iq;
    int N = s->N, Ts = s->Ts, Rs = s->Rs, Fs = s->Fs, nsym = s->Nsym, nin = s->nin, P = s->P, Nmem = s->Nmem, M = s->M;
    int nfi = (nsym + 1) * P, nold = Nmem - nin;
    cf *f_dc = s->f_dc, *f_int = s->f_int;
    (void)Rs;
    s->last.nin_used = nin;

    freq_est(s, x);
    const float *f_est = s->est_type ? s->f2_est : s->f_est;             /* fsk.c:630-634 */

    for (int m = 0; m < M; m++)                                          /* fsk.c:637-640 */
        for (int i = 0, j = Nmem - nold; i < nold; i++, j++) f_dc[m * Nmem + i] = f_dc[m * Nmem + j];

    for (int m = 0; m < M; m++) {                                        /* fsk.c:644-651 */
        cf dphi = c_expj(2 * M_PI * ((f_est[m]) / (float)(Fs)));
        cf ph = s->phi_c[m];
        for (int i = nold, j = 0; i < Nmem; i++, j++) {
            ph = c_mul(ph, dphi);
            f_dc[m * Nmem + i] = c_mul(x[j], c_conj(ph));
        }
        float av = sqrtf((ph.re * ph.re) + (ph.im * ph.im));             /* comp_prim.h:133-139 */
        ph.re = ph.re / av; ph.im = ph.im / av;
        s->phi_c[m] = ph;
    }

    for (int i = 0; i < nfi; i++) {                                      /* fsk.c:660-668 */
        int st = i * Ts / P, en = st + Ts - 1;
        for (int m = 0; m < M; m++) {
            cf acc = {0, 0};
            for (int j = st; j <= en; j++) acc = c_add(acc, f_dc[m * Nmem + j]);
            f_int[m * nfi + i] = acc;
        }
    }

    cf t_c = {0, 0};                                                     /* fsk.c:682-698 */
    for (int i = 0; i < nfi; i++) {
        float ft1 = 0;
        for (int m = 0; m < M; m++) {
            cf v = f_int[m * nfi + i];
            ft1 += (v.re * v.re) + (v.im * v.im);
        }
        t_c = c_add(t_c, c_scale(ft1, s->phi_ft[i]));
    }
    if (isnan(t_c.re) || isnan(t_c.im)) return;                          /* fsk.c:702-704: outputs left stale */

    float norm_rx_timing = atan2f(t_c.im, t_c.re) / (2 * M_PI);          /* fsk.c:707 */
    float rx_timing = norm_rx_timing * (float)P;
    float old = s->norm_rx_timing;
    s->norm_rx_timing = norm_rx_timing;
    float d = norm_rx_timing - old;
    if (fabsf(d) < .2) {                                                 /* fsk.c:717-720 */
        float appm = 1e6 * d / (float)nsym;
        s->ppm = .9 * s->ppm + .1 * appm;
    }
    if (norm_rx_timing > 0.25) s->nin = N + Ts / 2;                      /* fsk.c:725-730 (no burst/lock) */
    else if (norm_rx_timing < -0.25) s->nin = N - Ts / 2;
    else s->nin = N;

    int low = (int)floorf(rx_timing);                                    /* fsk.c:737-739 */
    float fract = rx_timing - (float)low;
    int high = (int)ceilf(rx_timing);

    float meanebno = 0, stdebno = 0;
    for (int i = 0; i < nsym; i++) {                                     /* fsk.c:751-817 */
        int st = (i + 1) * P;
        cf t[FSKO_MMAX]; float tmax[FSKO_MMAX];
        for (int m = 0; m < M; m++) {
            t[m] = c_scale(1 - fract, f_int[m * nfi + st + low]);
            t[m] = c_add(t[m], c_scale(fract, f_int[m * nfi + st + high]));
            tmax[m] = (t[m].re * t[m].re) + (t[m].im * t[m].im);
        }
        float max = tmax[0], min = tmax[0]; int sym = 0;
        for (int m = 0; m < M; m++) {
            if (tmax[m] > max) { max = tmax[m]; sym = m; }
            if (tmax[m] < min) min = tmax[m];
        }
        if (rx_bits) {
            if (M == 2) rx_bits[i] = sym == 1;
            else { rx_bits[i * 2 + 1] = sym & 1; rx_bits[i * 2] = (sym & 2) >> 1; }
        }
        if (rx_sd) {
            for (int m = 0; m < M; m++) tmax[m] = sqrtf(tmax[m]);
            if (M == 2) rx_sd[i] = tmax[0] - tmax[1];
            else {
                float a = -tmax[0], b = -tmax[0];       /* a = rx_sd[2i+1], b = rx_sd[2i]  (fsk.c:796-803) */
                a += tmax[1];  b += -tmax[1];
                a += -tmax[2]; b += tmax[2];
                a += tmax[3];  b += tmax[3];
                rx_sd[i * 2 + 1] = a; rx_sd[i * 2] = b;
            }
        }
        stdebno += max;                                                  /* fsk.c:810-814 */
        meanebno += sqrtf(max);
    }
    s->last.mean_sum = meanebno; s->last.std_sum = stdebno;
    meanebno = meanebno / (float)nsym;                                   /* fsk.c:822-833 */
    stdebno = (stdebno / (float)nsym) - (meanebno * meanebno);
    if (stdebno > 0.0) stdebno = sqrt(stdebno); else stdebno = 0.0;
    s->EbNodB = -6 + (20 * log10f((1e-6 + meanebno) / (1e-6 + stdebno)));
    s->snr_est = .5 * s->snr_est + .5 * s->EbNodB;                       /* fsk.c:845 */

    float fc_avg = 0.0, fc_tx = 0.0;                                     /* fsk.c:851-856 */
    for (int m = 0; m < M; m++) { fc_avg += f_est[m] / M; fc_tx += (s->f1_tx + m * s->fs_tx) / M; }

    /* eye traces (fsk.c:864-903) */
    int dec = ceil(((float)P * 2) / FSKO_EYE_IND);
    int neyesamp = (P * 2) / dec, neyeoffset = high + 1, eye_traces = FSKO_EYE_TR / M;
    s->neyesamp = neyesamp; s->neyetr = M * eye_traces;
    for (int i = 0; i < eye_traces; i++)
        for (int m = 0; m < M; m++)
            for (int j = 0; j < neyesamp; j++) {
                int ind = 2 * P * i + neyeoffset + j * dec;
                /* ind == -1 happens (high == -2): the reference then reads the VLA out of bounds --
                   the previous tone's last integrator for m>=1, stack garbage for m==0.  We read the
                   flat array the same way and define m==0 as zero. */
                cf v = {0, 0};
                int flat = m * nfi + ind;
                if (flat >= 0 && ind < nfi) v = f_int[flat];
                s->eye[i * M + m][j] = sqrtf((v.re * v.re) + (v.im * v.im));
            }
    float eye_max = 0;
    for (int i = 0; i < M * eye_traces; i++)
        for (int j = 0; j < neyesamp; j++) if (fabsf(s->eye[i][j]) > eye_max) eye_max = fabsf(s->eye[i][j]);
    for (int i = 0; i < M * eye_traces; i++)
        for (int j = 0; j < neyesamp; j++) s->eye[i][j] = s->eye[i][j] / eye_max;

    s->last.nin_next = s->nin; s->last.norm_rx_timing = norm_rx_timing; s->last.rx_timing = rx_timing;
    s->last.ppm = s->ppm; s->last.EbNodB = s->EbNodB; s->last.snr_est = s->snr_est;
    s->last.foff = fc_tx - fc_avg;
    for (int m = 0; m < M; m++) s->last.f_est[m] = f_est[m];
}

/* ---- the CLI frame pump over memory (fsk_demod.c:281-436) ------------------------------------- */
long fsko_run_stream(fsko_t *s, const void *in, long nsamp, int fmt, int soft, int softinv,
                     void *out, long *consumed, fsko_frame_t *recs, long max_frames)
{
    long pos = 0, nf = 0, nb = s->Nbits;
    int cap = s->N + 2 * s->Ts;
    float *mod = (float *)malloc(sizeof(float) * 2 * cap);
    float *sd = (float *)calloc(nb, sizeof(float));       /* persistent like the CLI's sdbuf/bitbuf */
    uint8_t *bits = (uint8_t *)calloc(nb, 1);
    while (pos + s->nin <= nsamp) {
        int nin = s->nin;
        if (fmt == 0) {                                                  /* fsk_demod.c:303-309 */
            const int16_t *r = (const int16_t *)in + 2 * pos;
            for (int i = 0; i < nin; i++) { mod[2 * i] = ((float)r[2 * i]) / 1000; mod[2 * i + 1] = ((float)r[2 * i + 1] / 1000); }
        } else if (fmt == 1) {
            memcpy(mod, (const float *)in + 2 * pos, sizeof(float) * 2 * nin);
        } else if (fmt == 2) {                                           /* fsk_demod.c:294-299 */
            const uint8_t *r = (const uint8_t *)in + 2 * pos;
            for (int i = 0; i < nin; i++) { mod[2 * i] = ((float)r[2 * i] - 127.0) / 128.0; mod[2 * i + 1] = ((float)r[2 * i + 1] - 127.0) / 128.0; }
        } else {                                                         /* fsk_demod.c:285-290 */
            const int16_t *r = (const int16_t *)in + pos;
            for (int i = 0; i < nin; i++) { mod[2 * i] = ((float)r[i]) / 1000; mod[2 * i + 1] = 0.0; }
        }
        if (soft) fsko_frame(s, mod, NULL, sd); else fsko_frame(s, mod, bits, NULL);
        if (soft) {
            if (softinv) for (long j = 0; j < nb; j++) sd[j] = sd[j] * -1.0;      /* fsk_demod.c:424-428 */
            memcpy((float *)out + nf * nb, sd, sizeof(float) * nb);
        } else memcpy((uint8_t *)out + nf * nb, bits, nb);
        if (recs && nf < max_frames) recs[nf] = s->last;
        pos += nin; nf++;
    }
    if (consumed) *consumed = pos;
    free(mod); free(sd); free(bits);
    return nf;
}
