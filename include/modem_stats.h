/* modem_stats.h -- struct MODEM_STATS as fsk_get_demod_stats() fills it (source-compatible stand-in for utils/modem_stats.h:37-78).
 * Field names and array bounds are the reference's (they are what callers index); the FFT/waterfall state the reference keeps in
 * this struct (kiss_fft_cfg) is not part of the FSK path and is absent. */
#ifndef B200_COMPAT_MODEM_STATS_H
#define B200_COMPAT_MODEM_STATS_H
#include "comp.h"
#ifdef __cplusplus
extern "C" {
#endif

#define MODEM_STATS_NC_MAX      50
#define MODEM_STATS_NR_MAX      8
#define MODEM_STATS_ET_MAX      8      /* eye traces              (== B200FSK_EYE_TR)  */
#define MODEM_STATS_EYE_IND_MAX 160    /* samples per eye trace   (== B200FSK_EYE_IND) */
#define MODEM_STATS_NSPEC       512
#define MODEM_STATS_MAX_F_HZ    4000
#define MODEM_STATS_MAX_F_EST   4

struct MODEM_STATS {
    int   Nc;
    float snr_est;                                   /* fsk: the smoothed Eb/N0 estimate in dB (fsk.c:845) */
    COMP  rx_symbols[MODEM_STATS_NR_MAX][MODEM_STATS_NC_MAX + 1];   /* unused by the FSK modem */
    int   nr;
    int   sync;
    float foff;                                      /* Hz, fsk.c:851-856 */
    float rx_timing;                                 /* samples */
    float clock_offset;                              /* ppm */
    float sync_metric;
    float rx_eye[MODEM_STATS_ET_MAX][MODEM_STATS_EYE_IND_MAX];
    int   neyetr, neyesamp;
    float f_est[MODEM_STATS_MAX_F_EST];              /* tone estimates of the estimator in use */
    float fft_buf[2 * MODEM_STATS_NSPEC];            /* unused here */
};

#ifdef __cplusplus
}
#endif
#endif
