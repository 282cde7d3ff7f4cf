/* ref_fsk_stream.c -- TEST INFRASTRUCTURE ONLY.  OUR driver around the UNMODIFIED reference modem:
 * it includes the reference's public header (utils/fsk.h) at build time from /root/reference and calls
 * only its public API (fsk_create_hbr, fsk_set_freq_est_limits, fsk_set_freq_est_alg, fsk_nin,
 * fsk_demod, fsk_demod_sd, fsk_get_demod_stats).  It reproduces the CLI frame pump of
 * utils/fsk_demod.c:281-436 over an in-memory stream so that tests/bench can run many channels from
 * threads without spawning processes.  Linked into oracle/_ref/libfsk_ref.so (git-ignored). */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "fsk.h"
#include "codec2_fdmdv.h"

/* rec layout per frame (floats): nin_used, nin_next, norm_rx_timing, ppm, EbNodB, snr_est, f_est[4] = 10 floats */
#define REF_REC 10

long ref_fsk_run_stream_passes(int Fs, int Rs, int M, int P, int nsym, int fs_tx, int lower, int upper, int mask,
                               const void *in, long nsamp, int fmt, int soft, int softinv,
                               void *out, long *consumed, float *recs, long max_frames, float *Sf_out, int passes);

long ref_fsk_run_stream(int Fs, int Rs, int M, int P, int nsym, int fs_tx, int lower, int upper, int mask,
                        const void *in, long nsamp, int fmt, int soft, int softinv,
                        void *out, long *consumed, float *recs, long max_frames, float *Sf_out)
{
    return ref_fsk_run_stream_passes(Fs, Rs, M, P, nsym, fs_tx, lower, upper, mask, in, nsamp, fmt, soft, softinv, out, consumed, recs, max_frames, Sf_out, 1);
}

/* `passes` > 1: the SAME modem object walks the same buffer `passes` times (each pass restarts at sample 0 and drops the
 * sub-frame remainder, the modem state carries over) -- what a handle of the batched library does when bench.py launches it
 * repeatedly on one resident buffer.  out / consumed / recs describe the LAST pass. */
long ref_fsk_run_stream_passes(int Fs, int Rs, int M, int P, int nsym, int fs_tx, int lower, int upper, int mask,
                               const void *in, long nsamp, int fmt, int soft, int softinv,
                               void *out, long *consumed, float *recs, long max_frames, float *Sf_out, int passes)
{
    struct FSK *fsk = fsk_create_hbr(Fs, Rs, M, P, nsym, 1000, fs_tx);
    struct MODEM_STATS stats;
    fsk_set_freq_est_limits(fsk, lower, upper);
    fsk_set_freq_est_alg(fsk, mask);
    long pos = 0, nf = 0, nb = fsk->Nbits;
    int pass;
    COMP *mod = (COMP *)malloc(sizeof(COMP) * (fsk->N + 2 * fsk->Ts));
    float *sd = (float *)calloc(nb, sizeof(float));
    uint8_t *bits = (uint8_t *)calloc(nb, 1);
    for (pass = 0; pass < passes; pass++) {
    pos = 0; nf = 0;
    while (pos + (long)fsk_nin(fsk) <= nsamp) {
        int nin = fsk_nin(fsk), i;
        if (fmt == 0) {
            const int16_t *r = (const int16_t *)in + 2 * pos;
            for (i = 0; i < nin; i++) { mod[i].real = ((float)r[2 * i]) / FDMDV_SCALE; mod[i].imag = ((float)r[2 * i + 1] / FDMDV_SCALE); }
        } else if (fmt == 1) {
            memcpy(mod, (const float *)in + 2 * pos, sizeof(COMP) * nin);
        } else if (fmt == 2) {
            const uint8_t *r = (const uint8_t *)in + 2 * pos;
            for (i = 0; i < nin; i++) { mod[i].real = ((float)r[2 * i] - 127.0) / 128.0; mod[i].imag = ((float)r[2 * i + 1] - 127.0) / 128.0; }
        } else {
            const int16_t *r = (const int16_t *)in + pos;
            for (i = 0; i < nin; i++) { mod[i].real = ((float)r[i]) / FDMDV_SCALE; mod[i].imag = 0.0; }
        }
        if (soft) fsk_demod_sd(fsk, sd, mod); else fsk_demod(fsk, bits, mod);
        if (out) {
            if (soft) {
                if (softinv) { long j; for (j = 0; j < nb; j++) sd[j] = sd[j] * -1.0; }
                memcpy((float *)out + nf * nb, sd, sizeof(float) * nb);
            } else memcpy((uint8_t *)out + nf * nb, bits, nb);
        }
        if (recs && nf < max_frames) {
            float *r = recs + REF_REC * nf; int m;
            float *fe = fsk->freq_est_type ? fsk->f2_est : fsk->f_est;
            fsk_get_demod_stats(fsk, &stats);
            r[0] = nin; r[1] = fsk_nin(fsk); r[2] = fsk->norm_rx_timing; r[3] = fsk->ppm;
            r[4] = fsk->EbNodB; r[5] = stats.snr_est;
            for (m = 0; m < 4; m++) r[6 + m] = m < M ? fe[m] : 0;
        }
        pos += nin; nf++;
    }
    }
    if (consumed) *consumed = pos;
    if (Sf_out) memcpy(Sf_out, fsk->Sf, sizeof(float) * fsk->Ndft);
    free(mod); free(sd); free(bits);
    fsk_destroy(fsk);
    return nf;
}
