/* fsk_compat.cu -- the reference's single-channel modem API (utils/fsk.h:105-205) exported unchanged from libb200fsk.so
 * (include/fsk.h).  Thin host code over the batched C-ABI (include/b200fsk.h) with one channel; the demodulation itself is the CUDA
 * frame kernel, bit-identical to utils/fsk.c.  The modulators (fsk.c:266-411) are host arithmetic for test-signal generation. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/b200fsk.h"
#include "../../include/fsk.h"

namespace {
struct Impl {
    b200fsk_t *h = nullptr;
    int device = 0;
    long frames = 0;
};

[[noreturn]] void die(const char *what)
{
    /* the reference asserts (fsk.c:120-130) or would crash; this library has no CPU path to fall back to */
    fprintf(stderr, "libb200fsk (fsk.h API): %s\n", what);
    exit(70);
}

Impl *impl(struct FSK *fsk) { if (!fsk || !fsk->b200) die("NULL struct FSK"); return (Impl *)fsk->b200; }

void ensure_handle(struct FSK *fsk)
{
    Impl *im = impl(fsk);
    if (im->h) return;
    b200fsk_cfg_t cfg;
    cfg.Fs = fsk->Fs; cfg.Rs = fsk->Rs; cfg.M = fsk->mode; cfg.P = fsk->P; cfg.Nsym = fsk->Nsym;
    cfg.f1_tx = fsk->f1_tx; cfg.fs_tx = fsk->fs_tx;
    cfg.est_min = fsk->est_min < -fsk->Fs / 2 ? -fsk->Fs / 2 : fsk->est_min;    /* fsk_create leaves [0, Fs]; the estimator only spans +-Fs/2 (fsk.c:459-462) */
    cfg.est_max = fsk->est_max > fsk->Fs / 2 ? fsk->Fs / 2 : fsk->est_max;
    cfg.est_type = fsk->freq_est_type ? 1 : 0;
    if (const char *d = getenv("B200FSK_DEVICE")) im->device = atoi(d);
    if (b200fsk_create(&im->h, &cfg, 1, im->device) != B200FSK_OK) {
        char msg[600]; snprintf(msg, sizeof msg, "cannot create the GPU modem: %s", b200fsk_last_error()); die(msg);
    }
    b200fsk_enable_stats(im->h, 1);
    if (fsk->burst_mode || fsk->lock_nin) b200fsk_set_lock_nin(im->h, 1);
}

void refresh(struct FSK *fsk)
{
    Impl *im = impl(fsk);
    b200fsk_stats_t st;
    if (b200fsk_get_stats(im->h, 0, &st) != B200FSK_OK) die(b200fsk_last_error());
    fsk->nin = st.nin; fsk->ppm = st.ppm; fsk->norm_rx_timing = st.norm_rx_timing; fsk->EbNodB = st.EbNodB;
    for (int m = 0; m < fsk->mode; ++m) { fsk->f_est[m] = st.f_est[m]; fsk->f2_est[m] = st.f_est[m]; }
    b200fsk_get_Sf(im->h, 0, fsk->Sf);
    struct MODEM_STATS *ms = fsk->stats;
    ms->snr_est = st.snr_est; ms->clock_offset = st.ppm; ms->rx_timing = st.rx_timing; ms->foff = st.foff;   /* fsk.c:845-862 */
    ms->neyetr = st.neyetr; ms->neyesamp = st.neyesamp;
    if (fsk->normalise_eye) memcpy(ms->rx_eye, st.rx_eye, sizeof ms->rx_eye);
    else {
        static_assert(sizeof(ms->rx_eye) == sizeof(float) * B200FSK_EYE_TR * B200FSK_EYE_IND, "eye layout");
        b200fsk_get_eye_raw(im->h, 0, &ms->rx_eye[0][0]);
    }
    for (int m = 0; m < fsk->mode; ++m) ms->f_est[m] = st.f_est[m];
}

void demod_frame(struct FSK *fsk, void *out, COMP *in, int soft)
{
    ensure_handle(fsk);
    Impl *im = impl(fsk);
    int frames = 0;
    const long nin = fsk->nin;
    if (b200fsk_push_host(im->h, in, B200FSK_FMT_F32, nin, nin, 1, out, fsk->Nbits, soft, 0, &frames) != B200FSK_OK) die(b200fsk_last_error());
    if (frames != 1) die("fsk_demod: the caller must pass exactly fsk_nin() samples");
    im->frames++;
    refresh(fsk);
}

COMP cexpj(float phi) { COMP r; r.real = cosf(phi); r.imag = sinf(phi); return r; }                       /* comp_prim.h:113-118 */
COMP cmul(COMP a, COMP b) { COMP r; r.real = a.real * b.real - a.imag * b.imag; r.imag = a.real * b.imag + a.imag * b.real; return r; }
COMP cnorm(COMP a) { COMP r; float av = sqrtf(a.real * a.real + a.imag * a.imag); r.real = a.real / av; r.imag = a.imag / av; return r; }
int next_symbol(const struct FSK *fsk, const uint8_t *bits, size_t &bit_i)                              /* MSB first, fsk.c:287-293 */
{
    int sym = 0;
    for (int m = fsk->mode; m >>= 1;) { sym = (sym << 1) | (bits[bit_i] == 1 ? 1 : 0); ++bit_i; }
    return sym;
}
}  // namespace

extern "C" struct FSK *fsk_create_hbr(int Fs, int Rs, int M, int P, int Nsym, int tx_f1, int tx_fs)
{
    /* the reference asserts on these (fsk.c:120-130) */
    if (!(Fs > 0 && Rs > 0 && tx_f1 > 0 && tx_fs > 0 && P > 0 && Nsym > 0)) die("fsk_create: non-positive parameter");
    if (Fs % Rs) die("fsk_create: Fs/Rs must be an integer");
    if ((Fs / Rs) % P) die("fsk_create: Ts/P must be an integer");
    if (M != 2 && M != 4) die("fsk_create: M must be 2 or 4");
    struct FSK *fsk = (struct FSK *)calloc(1, sizeof(struct FSK));
    if (!fsk) return NULL;
    fsk->Fs = Fs; fsk->Rs = Rs; fsk->Ts = Fs / Rs; fsk->P = P; fsk->Nsym = Nsym; fsk->mode = M;
    fsk->N = fsk->Ts * Nsym; fsk->Nmem = fsk->N + 2 * fsk->Ts; fsk->Nbits = M == 2 ? Nsym : 2 * Nsym;
    fsk->f1_tx = tx_f1; fsk->fs_tx = tx_fs;
    float bin_width_Hz = 0.1 * Rs;                                    /* fsk.c:135-137 */
    float Ndft = (float)Fs / bin_width_Hz;
    Ndft = pow(2.0, ceil(log2(Ndft)));
    fsk->Ndft = (int)Ndft;
    fsk->tc = 0.95 * Ndft / Fs;                                       /* fsk.c:148 */
    fsk->est_min = 0; fsk->est_max = Fs; fsk->est_space = 0.75 * Rs;  /* fsk.c:156-158 */
    fsk->nin = fsk->N; fsk->normalise_eye = 1;
    fsk->tx_phase_c = cexpj(0); for (int m = 0; m < MODE_M_MAX; ++m) fsk->phi_c[m] = cexpj(0);
    fsk->hann_table = (float *)malloc(sizeof(float) * fsk->Ndft);
    fsk->Sf = (float *)calloc(fsk->Ndft, sizeof(float));
    fsk->stats = (struct MODEM_STATS *)calloc(1, sizeof(struct MODEM_STATS));
    if (!fsk->hann_table || !fsk->Sf || !fsk->stats) die("out of memory");
    for (int i = 0; i < fsk->Ndft; ++i) fsk->hann_table[i] = 0.5 - 0.5 * cosf(2.0 * M_PI * (float)i / (float)(fsk->Ndft - 1));   /* fsk.c:96 */
    fsk->stats->neyesamp = 0; fsk->stats->neyetr = 0;
    fsk->b200 = new Impl();
    return fsk;
}

extern "C" struct FSK *fsk_create(int Fs, int Rs, int M, int tx_f1, int tx_fs)
{
    return fsk_create_hbr(Fs, Rs, M, FSK_DEFAULT_P, FSK_DEFAULT_NSYM, tx_f1, tx_fs);
}

extern "C" void fsk_destroy(struct FSK *fsk)
{
    if (!fsk) return;
    Impl *im = (Impl *)fsk->b200;
    if (im) { if (im->h) b200fsk_destroy(im->h); delete im; }
    free(fsk->hann_table); free(fsk->Sf); free(fsk->stats); free(fsk);
}

extern "C" void fsk_set_freq_est_limits(struct FSK *fsk, int est_min, int est_max)
{
    Impl *im = impl(fsk);
    if (!(est_min >= -fsk->Fs / 2 && est_max <= fsk->Fs / 2 && est_max > est_min)) die("fsk_set_freq_est_limits: limits outside +-Fs/2 or empty (fsk.c:1019-1021)");
    if (im->h) {
        if (im->frames) die("fsk_set_freq_est_limits after the first fsk_demod call is not supported");
        b200fsk_destroy(im->h); im->h = nullptr;
    }
    fsk->est_min = est_min; fsk->est_max = est_max;
}

extern "C" void fsk_set_freq_est_alg(struct FSK *fsk, int est_type)
{
    Impl *im = impl(fsk);
    if (im->h) {
        if (im->frames) die("fsk_set_freq_est_alg after the first fsk_demod call is not supported");
        b200fsk_destroy(im->h); im->h = nullptr;
    }
    fsk->freq_est_type = est_type;
}

extern "C" void fsk_clear_estimators(struct FSK *fsk)
{
    ensure_handle(fsk);
    if (b200fsk_clear_estimators(impl(fsk)->h) != B200FSK_OK) die(b200fsk_last_error());
    memset(fsk->Sf, 0, sizeof(float) * fsk->Ndft);                    /* fsk.c:984-988 */
    fsk->nin = fsk->N;
}

extern "C" void fsk_get_demod_stats(struct FSK *fsk, struct MODEM_STATS *stats)
{
    impl(fsk);
    const struct MODEM_STATS *s = fsk->stats;                         /* same field list as fsk.c:997-1011 */
    stats->clock_offset = s->clock_offset; stats->snr_est = s->snr_est; stats->rx_timing = s->rx_timing; stats->foff = s->foff;
    stats->neyesamp = s->neyesamp; stats->neyetr = s->neyetr;
    memcpy(stats->rx_eye, s->rx_eye, sizeof stats->rx_eye);
    memcpy(stats->f_est, s->f_est, fsk->mode * sizeof(float));
    stats->sync = 0; stats->nr = s->nr; stats->Nc = s->Nc;
}

extern "C" void fsk_stats_normalise_eye(struct FSK *fsk, int normalise_enable) { impl(fsk); fsk->normalise_eye = normalise_enable; }

extern "C" void fsk_enable_burst_mode(struct FSK *fsk)
{
    Impl *im = impl(fsk);
    fsk->burst_mode = 1;
    if (im->h) b200fsk_set_lock_nin(im->h, 1);
}

extern "C" uint32_t fsk_nin(struct FSK *fsk) { impl(fsk); return (uint32_t)fsk->nin; }

extern "C" void fsk_demod(struct FSK *fsk, uint8_t rx_bits[], COMP fsk_in[]) { demod_frame(fsk, rx_bits, fsk_in, 0); }
extern "C" void fsk_demod_sd(struct FSK *fsk, float rx_sd[], COMP fsk_in[]) { demod_frame(fsk, rx_sd, fsk_in, 1); }

/* ---- modulators: one frame of Nsym symbols, Ts samples each; the oscillator advances by the symbol's phase step per sample ---- */
extern "C" void fsk_mod_c(struct FSK *fsk, COMP fsk_out[], uint8_t tx_bits[])
{
    impl(fsk);
    COMP step[MODE_M_MAX], ph = fsk->tx_phase_c;
    for (int m = 0; m < fsk->mode; ++m) step[m] = cexpj(2 * M_PI * ((float)(fsk->f1_tx + (fsk->fs_tx * m)) / (float)(fsk->Fs)));
    size_t bit_i = 0;
    for (int i = 0; i < fsk->Nsym; ++i) {
        const COMP d = step[next_symbol(fsk, tx_bits, bit_i)];
        for (int j = 0; j < fsk->Ts; ++j) {
            ph = cmul(ph, d);
            fsk_out[i * fsk->Ts + j].real = 2 * ph.real; fsk_out[i * fsk->Ts + j].imag = 2 * ph.imag;
        }
    }
    fsk->tx_phase_c = cnorm(ph);
}

extern "C" void fsk_mod(struct FSK *fsk, float fsk_out[], uint8_t tx_bits[])
{
    impl(fsk);
    COMP step[MODE_M_MAX], ph = fsk->tx_phase_c;
    for (int m = 0; m < fsk->mode; ++m) step[m] = cexpj(2 * M_PI * ((float)(fsk->f1_tx + (fsk->fs_tx * m)) / (float)(fsk->Fs)));
    size_t bit_i = 0;
    for (int i = 0; i < fsk->Nsym; ++i) {
        const COMP d = step[next_symbol(fsk, tx_bits, bit_i)];
        for (int j = 0; j < fsk->Ts; ++j) { ph = cmul(ph, d); fsk_out[i * fsk->Ts + j] = 2 * ph.real; }
    }
    fsk->tx_phase_c = cnorm(ph);
}

extern "C" void fsk_mod_ext_vco(struct FSK *fsk, float vco_out[], uint8_t tx_bits[])
{
    impl(fsk);
    size_t bit_i = 0;
    for (int i = 0; i < fsk->Nsym; ++i) {
        const int sym = next_symbol(fsk, tx_bits, bit_i);
        for (int j = 0; j < fsk->Ts; ++j) vco_out[i * fsk->Ts + j] = fsk->f1_tx + fsk->fs_tx * (float)sym;
    }
}
