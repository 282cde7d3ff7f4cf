/* fsk.h -- the reference's single-channel FSK modem API (utils/fsk.h:105-205) exported UNCHANGED by libb200fsk.so.
 *
 * A program written against radiosonde_auto_rx's utils/fsk.h -- the reference's own utils/fsk_demod.c included -- compiles against
 * this directory and links with -lb200fsk without edits (tests/test_fsk_capi.py does exactly that with the unmodified file).
 * Every function keeps its name, argument list and meaning; the demodulation runs on the GPU through the batched library
 * (include/b200fsk.h) with one channel, so each fsk_demod() call costs a host<->device round trip: this surface is for drop-in
 * compatibility, the batched entry points are the fast path.  There is no CPU demodulator behind it: without a CUDA device the
 * first call that needs the device prints a message and exits.
 *
 * struct FSK: the members the reference documents as readable by the driving code are kept up to date after every fsk_demod*()
 * call (Sf is a host mirror of the device array; f_est/f2_est both hold the tones of the estimator in use).  Members that only the
 * reference's own implementation touches (f_dc, fft_cfg) are absent. */
#ifndef B200_COMPAT_FSK_H
#define B200_COMPAT_FSK_H
#include <stdint.h>
#include "comp.h"
#include "modem_stats.h"
#ifdef __cplusplus
extern "C" {
#endif

#define MODE_2FSK 2
#define MODE_4FSK 4
#define MODE_M_MAX 4
#define FSK_SCALE 16383
#define FSK_DEFAULT_P 8           /* fsk.h:45 */
#define FSK_DEFAULT_NSYM 50       /* fsk.h:46 */

struct FSK {
    /* static parameters, fsk_create_core (fsk.c:114-201) */
    int Ndft, Fs, N, Rs, Ts, Nmem, P, Nsym, Nbits;
    int f1_tx, fs_tx;
    int mode;                     /* 2 or 4 */
    float tc;
    int est_min, est_max, est_space;
    float *hann_table;            /* [Ndft], host copy */
    /* demodulator state visible to the caller */
    float *Sf;                    /* [Ndft] smoothed magnitude spectrum after the last frame (host mirror) */
    COMP phi_c[MODE_M_MAX];       /* not mirrored (lives on the device) */
    float norm_rx_timing;
    COMP tx_phase_c;              /* modulator phase (fsk_mod*) */
    float EbNodB;
    float f_est[MODE_M_MAX];
    float f2_est[MODE_M_MAX];
    int   freq_est_type;          /* 0: peak picker, 1: mask correlator (fsk_set_freq_est_alg) */
    float ppm;
    int nin;                      /* fsk_nin() */
    int burst_mode, lock_nin;
    struct MODEM_STATS *stats;
    int normalise_eye;
    void *b200;                   /* implementation handle (batched GPU library), not for callers */
};

struct FSK *fsk_create(int Fs, int Rs, int M, int tx_f1, int tx_fs);                          /* fsk.c:215 */
struct FSK *fsk_create_hbr(int Fs, int Rs, int M, int P, int Nsym, int tx_f1, int tx_fs);     /* fsk.c:235 */
void fsk_destroy(struct FSK *fsk);                                                            /* fsk.c:249 */
void fsk_set_freq_est_limits(struct FSK *fsk, int fmin, int fmax);                            /* fsk.c:1017 (before the first demod call) */
void fsk_set_freq_est_alg(struct FSK *fsk, int est_type);                                     /* fsk.c:1031 (before the first demod call) */
void fsk_clear_estimators(struct FSK *fsk);                                                   /* fsk.c:981 */
void fsk_get_demod_stats(struct FSK *fsk, struct MODEM_STATS *stats);                         /* fsk.c:991 */
void fsk_stats_normalise_eye(struct FSK *fsk, int normalise_enable);                          /* fsk.c:1026 */
void fsk_enable_burst_mode(struct FSK *fsk);                                                  /* fsk.c:976: nin stays fixed */
uint32_t fsk_nin(struct FSK *fsk);                                                            /* fsk.c:424 */
void fsk_demod(struct FSK *fsk, uint8_t rx_bits[], COMP fsk_in[]);                            /* fsk.c:930: fsk_nin() samples -> Nbits bits */
void fsk_demod_sd(struct FSK *fsk, float rx_bits[], COMP fsk_in[]);                           /* fsk.c:934: ... -> Nbits soft decisions */
/* modulators (host arithmetic, test-signal generation): one frame of Nsym symbols -> N samples */
void fsk_mod(struct FSK *fsk, float fsk_out[], uint8_t tx_bits[]);                            /* fsk.c:266 */
void fsk_mod_c(struct FSK *fsk, COMP fsk_out[], uint8_t tx_bits[]);                           /* fsk.c:321 */
void fsk_mod_ext_vco(struct FSK *fsk, float vco_out[], uint8_t tx_bits[]);                    /* fsk.c:376 */

#ifdef __cplusplus
}
#endif
#endif
