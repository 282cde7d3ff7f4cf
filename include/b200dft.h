/* b200dft.h -- C-ABI of the B200-native batched sonde-type detector (dft_detect).
 *
 * Drop-in boundary for scan/dft_detect.c in its production IQ mode
 *     dft_detect [-v] [-c] [-t sec] --iq --bw <kHz> [--dc] - <sample_rate> <8|16|32>
 * (auto_rx/autorx/scan.py:541-547, scan_async.py:143-148).  The reference is a stdin->stdout executable with
 * all state in file-scope statics; there is no in-process API, so the entry points below are what a binding
 * replacing "one dft_detect process per candidate frequency" (scan.py:621, scan_async.py:209-222) would call:
 *
 *   reference (file:line)                                        here
 *   ----------------------------------------------------------   -------------------------------------
 *   argv parsing + init_buffers          dft_detect.c:1368-1476   b200dft_create (sample_rate, bw, dc)
 *   while(f32buf_sample) ... getCorrDFT  dft_detect.c:1492-1512   b200dft_run_host / b200dft_run_device
 *   threshold/headcmp/M10/iMet + printf  dft_detect.c:1514-1652   result records (type, score, pos, df)
 *   exit code  header_found * tn         dft_detect.c:1660-1666   b200dft_result_t.exit_code
 *
 * One handle = nchan independent IQ streams (one per candidate frequency), processed concurrently.
 * Detected type, sign and sample position are identical to the reference's; scores agree to ~1e-4 (the
 * reference itself is built -Ofast, so its float sums are not reproducible bit for bit).
 * Plain pointers and sizes only.  No CPU fallback: without a CUDA device b200dft_create fails.
 */
#ifndef B200DFT_H
#define B200DFT_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B200DFT_OK       0
#define B200DFT_EINVAL  -1
#define B200DFT_ECUDA   -2
#define B200DFT_ERANGE  -4

#define B200DFT_FMT_CS16 0    /* int16 I,Q  /32768   (dft_detect.c:567-571) */
#define B200DFT_FMT_F32  1    /* float I,Q           (dft_detect.c:561-565) */
#define B200DFT_FMT_CU8  2    /* uint8 I,Q  (u-128)/128 (dft_detect.c:573-577) */

#define B200DFT_NHDR 18       /* rows of rs_hdr[] (dft_detect.c:177-194) */
#define B200DFT_MAXDET 64     /* detections kept per channel */

typedef struct b200dft b200dft_t;

typedef struct {              /* one line the reference prints (dft_detect.c:1620-1634) */
    int      hdr;             /* row of the header bank                                 */
    int      tn;              /* type number (dft_detect.c:152-170)                     */
    float    score;           /* signed normalised correlation, "%.4f"                  */
    uint32_t pos;             /* "sample: N" (-v)                                       */
    float    df_hz;           /* " , %+.1fHz" (--dc --iq)                               */
    uint32_t m10;             /* M10/M20 frame bytes (-v "[%04X]")                      */
    char     type[12];        /* "RS41", "DFM9", "M20", ...                             */
} b200dft_det_t;

typedef struct {
    int exit_code;            /* what the process would return: +-tn of the best score, 0 = nothing */
    int ndet;                 /* lines printed (capped at B200DFT_MAXDET in dets[])                  */
    b200dft_det_t dets[B200DFT_MAXDET];
} b200dft_result_t;

typedef struct {
    int sample_rate, N_DFT, K, L, M, delay, lpIQtaps, lpFMtaps, nhdr_active, nchan;
} b200dft_info_t;

/* bw_khz is the --bw argument (required: the CUDA path implements the single-IF-filter mode that --bw selects,
 * dft_detect.c:1132-1137, which is what auto_rx always passes); opt_dc is --dc. */
int  b200dft_create(b200dft_t **out, int sample_rate, double bw_khz, int opt_dc, int nchan, int device);
void b200dft_destroy(b200dft_t *h);
/* `--ths x` (dft_detect.c:1426-1431): replace the per-header detection thresholds by one value (before the first run) */
int  b200dft_set_threshold(b200dft_t *h, double ths);
int  b200dft_info(const b200dft_t *h, b200dft_info_t *info);
const char *b200dft_last_error(void);

/* Whole detector run over nchan in-memory streams of nsamp samples each (channel stride in_stride samples).
 * opt_cont = -c (keep going after a detection); tl = -t seconds (<=0: none).  res[nchan] receives what each
 * reference process would have printed and returned. Host buffers: H2D inside the call. */
int  b200dft_run_host(b200dft_t *h, const void *h_iq, int fmt, long in_stride, long nsamp, int opt_cont, double tl,
                      b200dft_result_t *res);
/* Same with the IQ already on the device. */
int  b200dft_run_device(b200dft_t *h, const void *d_iq, int fmt, long in_stride, long nsamp, int opt_cont, double tl,
                        b200dft_result_t *res);

/* ---- streaming form of the --iq mode: what ONE long-running dft_detect process does between reads, for nchan streams at once ----
 * begin(fmt, -c, -t) resets the stream; every push appends nsamp samples per channel and advances the detector as far as the data
 * allows.  res[nchan] must be the same array for the whole stream: new detections are APPENDED to res[c].dets (the caller may print
 * them and reset res[c].ndet to 0 between pushes), res[c].exit_code is final once done[c] != 0 (first detection without -c, the -t
 * limit, or flush != 0 = end of input).  Results are identical to one whole-buffer call over the concatenation of all pushes;
 * the handle keeps only the K+L+... samples later blocks can still reach, so work and memory per push do not grow with the stream. */
int  b200dft_stream_begin(b200dft_t *h, int fmt, int opt_cont, double tl);
int  b200dft_stream_push_host(b200dft_t *h, const void *h_iq, long in_stride, long nsamp, int flush, b200dft_result_t *res, int *done);

/* ---- wideband mode: `dft_detect --IQ <fq> --bw <kHz> [--dc] - <sr_base> <bits>` --------------------------------
 * The reference translates ONE candidate frequency per process out of a wide capture (rotation LUT + decimating low-pass,
 * dft_detect.c:751-761, tables :1021-1111; the same front as demod/mod/iq_dec.c:514-545) and then runs the detector at the
 * IF rate.  Here one handle takes the capture once and nchan values of fq (the --IQ argument: fraction of sr_base, clipped to
 * +-0.5 like :1388-1389); each result is what the process started with that fq would have printed and returned. */
typedef struct { int sr_base, if_sr, decM, dectaps, nchan; } b200dft_wbinfo_t;
int  b200dft_create_wideband(b200dft_t **out, int sr_base, double bw_khz, int opt_dc, int nchan, int device);
int  b200dft_wideband_info(const b200dft_t *h, b200dft_wbinfo_t *info);
int  b200dft_run_wideband_host(b200dft_t *h, const void *h_wide, int fmt, long nwide, const double *fq, int opt_cont, double tl,
                               b200dft_result_t *res);
int  b200dft_run_wideband_device(b200dft_t *h, const void *d_wide, int fmt, long nwide, const double *fq, int opt_cont, double tl,
                                 b200dft_result_t *res);
/* channeliser alone: d_if_out[c*out_stride + m] (float I,Q pairs, device memory) = IF sample m of channel c; *nout = nwide/decM.
 * This is the stream `iq_dec --iq <fq> - <sr_base> <bits>` would emit before its output conversion (iq_dec.c:514-545). */
int  b200dft_channelise_device(b200dft_t *h, const void *d_wide, int fmt, long nwide, const double *fq, void *d_if_out, long out_stride,
                               long *nout);
int  b200dft_last_channelise_ms(b200dft_t *h, float *ms);

/* ---- scanner peak stage: where to point the detector -------------------------------------------------------------
 * auto_rx: rtl_power spectrum -> median noise floor -> detect_peaks -> sorted / quantised / clipped candidate list
 * (auto_rx/autorx/scan.py:974-1063, auto_rx/autorx/utils.py:435-586).  Here the spectrum is computed from the capture that is
 * already in HBM for the channeliser, so candidates and detections come from the same samples. */
/* averaged Hann periodogram: power_db[k] / freq_hz[k], k < nfft (power of two, 64..16384), ascending frequency around centre_hz;
 * *step_hz = sr_base/nfft.  (rtl_power is an external program: no reference restatement exists for this function.) */
int  b200dft_scan_psd_device(const void *d_wide, int fmt, long nwide, int sr_base, double centre_hz, int nfft, double *freq_hz,
                             double *power_db, double *step_hz, int device, float *kernel_ms);
/* detect_peaks(x, mph, mpd) with the other arguments at the values sonde_search uses (utils.py:435-586); returns the peak count,
 * indices ascending in idx_out (at most max_out written).  use_mph = 0: mph=None. */
int  b200dft_scan_detect_peaks(const double *x, int n, int use_mph, double mph, double mpd, int *idx_out, int max_out);
/* scan.py:1005-1063: noise floor (median), peaks above floor + snr_threshold at least min_distance apart, strongest first, rounded to
 * quantization, duplicates dropped, outside [min_freq, max_freq] +- quantization/2 dropped, first max_peaks kept. */
int  b200dft_scan_candidates(const double *freq_hz, const double *power_db, int n, double step_hz, double snr_threshold_db, double min_distance_hz,
                             double quantization_hz, double min_freq_hz, double max_freq_hz, int max_peaks, double *out_hz, int max_out,
                             double *noise_floor_db);
const char *b200dft_scan_last_error(void);

long  b200dft_launch_count(const b200dft_t *h);
/* device time (ms) of the front-end kernels and of the correlator kernel in the most recent run */
int   b200dft_last_kernel_ms(b200dft_t *h, float *front_ms, float *corr_ms);

#ifdef __cplusplus
}
#endif
#endif
