/* b200fsk.h -- C-ABI of the B200-native batched non-coherent 2/4-FSK demodulator.
 *
 * Drop-in boundary for the IQ->bits hot path of radiosonde_auto_rx (reference: utils/fsk.c,
 * utils/kiss_fft.c, utils/fsk_demod.c).  The reference has no in-process plugin ABI: its contract is
 * the small C API of utils/fsk.h:105-205 plus the stdin/stdout byte formats of the fsk_demod CLI
 * (utils/fsk_demod.c:281-436).  This header is what a binding for that path would bind:
 *
 *   reference interface (file:line)                          entry point here
 *   ------------------------------------------------------   ------------------------------------
 *   fsk_create_hbr           utils/fsk.h:125, fsk.c:235      b200fsk_create          (x nchan)
 *   fsk_set_freq_est_limits  utils/fsk.h:139, fsk.c:1017     b200fsk_cfg_t.est_min/est_max
 *   fsk_set_freq_est_alg     utils/fsk.h:205, fsk.c:1031     b200fsk_cfg_t.est_type
 *   fsk_destroy              utils/fsk.h:144, fsk.c:249      b200fsk_destroy
 *   fsk_nin                  utils/fsk.h:171, fsk.c:424      b200fsk_nin
 *   fsk_demod / fsk_demod_sd utils/fsk.h:179-189, fsk.c:930  b200fsk_demod_device / b200fsk_push_host
 *   fsk_get_demod_stats      utils/fsk.h:193, fsk.c:991      b200fsk_get_stats
 *   fsk_clear_estimators     utils/fsk.h:191, fsk.c:981      b200fsk_clear_estimators
 *   CLI frame pump + sample conversion fsk_demod.c:281-436   b200fsk_push_host (fmt, soft, softinv)
 *
 * Plain pointers and sizes only; no torch / C++ types.  All functions return 0 on success or a
 * negative B200FSK_E* code (the reference asserts instead).  There is NO CPU
 * fallback: without a CUDA device b200fsk_create fails with B200FSK_ECUDA.
 *
 * A handle is one homogeneous batch: nchan independent channels sharing one modem configuration,
 * each with its own persistent demodulator state on the GPU (Sf, oscillator phases, integrator
 * memory, timing, nin).  Frames of one channel are strictly sequential; channels run concurrently.
 * Results are bit-identical to the reference per channel (hard bits, soft decisions, nin sequence).
 */
#ifndef B200FSK_H
#define B200FSK_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FSK_OK       0
#define B200FSK_EINVAL  -1   /* configuration the reference would assert on (fsk.c:120-130)  */
#define B200FSK_ECUDA   -2   /* CUDA runtime error / no device / extension unusable           */
#define B200FSK_ENOMEM  -3   /* frame working set does not fit one SM's shared memory         */
#define B200FSK_ERANGE  -4   /* argument out of range                                         */

/* input sample formats (utils/fsk_demod.c:284-310) */
#define B200FSK_FMT_CS16 0   /* interleaved int16 I,Q ; value/1000 (FDMDV_SCALE)              */
#define B200FSK_FMT_F32  1   /* interleaved float I,Q (struct COMP, utils/comp.h:33-36)       */
#define B200FSK_FMT_CU8  2   /* interleaved uint8 I,Q ; (u-127)/128                           */
#define B200FSK_FMT_S16  3   /* real int16 ; value/1000, imag = 0                             */

#define B200FSK_MMAX     4
#define B200FSK_EYE_TR   8   /* MODEM_STATS_ET_MAX      utils/modem_stats.h:40 */
#define B200FSK_EYE_IND  160 /* MODEM_STATS_EYE_IND_MAX utils/modem_stats.h:41 */

typedef struct b200fsk b200fsk_t;

/* Arguments of fsk_create_hbr + fsk_set_freq_est_limits + fsk_set_freq_est_alg. */
typedef struct {
    int Fs, Rs, M, P, Nsym;     /* fsk.c:114                                                   */
    int f1_tx, fs_tx;           /* modulator tone 0 / spacing; fs_tx drives the mask estimator  */
    int est_min, est_max;       /* Hz, within [-Fs/2, Fs/2]                  (fsk.c:1017-1024)  */
    int est_type;               /* 0 = peak picker, 1 = mask correlator      (fsk.c:631-634)    */
} b200fsk_cfg_t;

/* Derived sizes (struct FSK fields the reference CLI reads: N, Ts, Nbits, Ndft ...). */
typedef struct {
    int Ts, N, Nmem, Ndft, Nbits, nin_max, nin_min, nchan;
    size_t smem_bytes;          /* dynamic shared memory per channel CTA                        */
} b200fsk_info_t;

/* What fsk_get_demod_stats + struct FSK expose after the most recent frame of a channel. */
typedef struct {
    float snr_est;              /* "EbNodB" in the CLI JSON (fsk_demod.c:365)                   */
    float EbNodB;               /* instantaneous (fsk.c:833)                                    */
    float ppm;                  /* fsk->ppm == stats.clock_offset                               */
    float rx_timing, norm_rx_timing, foff;
    float f_est[B200FSK_MMAX];  /* estimator in use                                             */
    int   nin;                  /* fsk_nin() for the next frame                                 */
    int   neyetr, neyesamp;
    float rx_eye[B200FSK_EYE_TR][B200FSK_EYE_IND];
    long long frames;           /* frames demodulated so far on this channel                    */
    long long samples;          /* input samples consumed so far                                */
} b200fsk_stats_t;

int  b200fsk_create(b200fsk_t **out, const b200fsk_cfg_t *cfg, int nchan, int device);
void b200fsk_destroy(b200fsk_t *h);
int  b200fsk_info(const b200fsk_t *h, b200fsk_info_t *info);
const char *b200fsk_last_error(void);

/* fsk_nin() for every channel (host array of nchan ints). Synchronises the handle's stream. */
int  b200fsk_nin(b200fsk_t *h, int *nin_out);
/* fsk_clear_estimators (fsk.c:981-989) on every channel. */
int  b200fsk_clear_estimators(b200fsk_t *h);
/* Upper bound of frames a stream of nsamp samples can hold. */
long b200fsk_max_frames(const b200fsk_t *h, long nsamp);

/* ---- device-resident path (the kernel boundary) -------------------------------------------------
 * d_in        device pointer; channel c's stream starts at element c*in_stride (+ d_in_off[c] if given),
 *             counted in SAMPLES of `fmt`.
 * nsamp       samples available per channel (stream end, measured from c*in_stride).
 * max_frames  frame cap per channel for this call (<=0: no cap).
 * d_out       device pointer; channel c writes frame k at  c*out_stride + k*Nbits  (elements:
 *             float if soft else uint8).  softinv negates soft output (fsk_demod.c:424-428).
 * Each channel demodulates whole frames while cursor+nin <= nsamp (the CLI's fread loop, :281);
 * per-channel frame/consumed counters are kept on the device (b200fsk_counts fetches them).
 * `stream` is a cudaStream_t (void* here to keep the header CUDA-free). Asynchronous. */
int  b200fsk_demod_device(b200fsk_t *h, const void *d_in, int fmt, long in_stride, const int *d_in_off,
                          long nsamp, int max_frames, void *d_out, long out_stride,
                          int soft, int softinv, void *stream);
/* frames[c], consumed[c] of the most recent demod call (host arrays; either may be NULL). Synchronises. */
int  b200fsk_counts(b200fsk_t *h, int *frames, long *consumed);

/* ---- host-buffer path (what the CLI / a ctypes or cgo caller uses) ------------------------------
 * Append `nsamp` new samples per channel (h_in: [nchan][nsamp] samples of fmt, channel stride
 * in_stride samples) to each channel's stream, demodulate every whole frame now available (at most
 * max_frames per channel if >0), and copy results to h_out ([nchan] rows of out_stride elements).
 * Unconsumed samples (< nin) are carried inside the handle to the next push, exactly like the bytes
 * a blocking fread would still be waiting for.  frames[c] receives the frames written for channel c.
 * Synchronous: H2D + kernel + D2H all inside the call. */
int  b200fsk_push_host(b200fsk_t *h, const void *h_in, int fmt, long in_stride, long nsamp, int max_frames,
                       void *h_out, long out_stride, int soft, int softinv, int *frames);

/* Frames one push of nsamp new samples can emit per channel (size h_out rows with this * Nbits). */
long b200fsk_push_frames_bound(const b200fsk_t *h, long nsamp, int max_frames);

/* Demod statistics of channel c after its latest frame (fsk_get_demod_stats + fsk->ppm/f_est/nin).
 * EbNodB/snr_est use the host libm (log10f) on per-frame sums produced by the kernel, so they are
 * bit-identical to the reference's.  Requires b200fsk_enable_stats(h,1) before the frames of interest. */
int  b200fsk_enable_stats(b200fsk_t *h, int on);
int  b200fsk_get_stats(b200fsk_t *h, int chan, b200fsk_stats_t *st);
/* fsk->Sf of channel c (Ndft floats; the CLI prints the first Ndft/2 as "samp_fft", fsk_demod.c:397-402). */
int  b200fsk_get_Sf(b200fsk_t *h, int chan, float *Sf_out);

/* ---- next row of the path: soft-bit header sync + RS41 frame assembly, soft bits stay in HBM -----------
 * Reference: find_softbinhead / corr_softhdb (demod/mod/demod_mod.c:1692-1761) + the --softin frame loop of rs41mod
 * (demod/mod/rs41mod.c:2875-2960), i.e. what `rs41mod --softin [-i] -r` does with fsk_demod's stdout before its ECC stage.
 * d_soft: [nchan] rows of `stride` floats as written by b200fsk_demod_device (soft=1); per-channel bit counts in d_nbits
 * (device, may be NULL -> nbits for all).  softinv = --softinv, inv = -i.  Frames are returned grouped by channel, in order. */
typedef struct {
    int   chan, len, nbytes_read;      /* len: 320 or 518 from the frame-type byte; nbytes_read < 518: stream ended inside the frame */
    long long first_bit;               /* index of the first payload soft bit                                                   */
    float score;                       /* header correlation                                                                    */
    unsigned char bytes[518];          /* de-whitened frame (bytes 0..7: header)                                                */
    float bytescore[518];              /* soft bit of smallest magnitude per byte (input of the reference's ecc3/4 erasure logic) */
    int   ec;                          /* ecc > 0: rs41_ecc() result: corrected symbols, or -1 / -2 / -3 = word 1 / word 2 / both failed    */
    int   crc_mask;                    /* ecc > 0: bit k = check_CRC() == 0 for block k of 7928 FRAME, 7A2A PTU, 7C1E GPS1, 7D59 GPS2,   */
                                       /*          7B15 GPS3, 76xx ZERO at their standard positions (rs41mod.c:307-319, 338-392)       */
} b200fsk_rs41_frame_t;
int  b200fsk_rs41_framesync_device(const float *d_soft, long stride, const int *d_nbits, int nbits, int nchan, int softinv, int inv,
                                   b200fsk_rs41_frame_t *h_frames, int max_frames, int *nframes, int device, void *stream);
/* the same with rs41mod's Reed-Solomon stage applied to every frame before it is returned (rs41_ecc,
 * demod/mod/rs41mod.c:1703-1976; decoder demod/mod/bch_ecc_mod.c:877-966).  ecc = 0: none (identical to the call above),
 * 1: `--ecc` (rs41mod's default), 2: `--ecc2`, which `--json` selects (rs41mod.c:2703-2706) -- the level auto_rx runs at
 * (auto_rx/autorx/decode.py:909).  bytes[] then holds the corrected frame, exactly what `rs41mod -r` prints. */
int  b200fsk_rs41_frames_device(const float *d_soft, long stride, const int *d_nbits, int nbits, int nchan, int softinv, int inv, int ecc,
                                b200fsk_rs41_frame_t *h_frames, int max_frames, int *nframes, int device, void *stream);
/* Streaming form: one handle = nchan running streams (what one `rs41mod --softin` process per channel keeps between reads: header
 * ring, frame buffer, position inside a frame).  Each push continues where the previous one stopped, a frame that straddles two pushes
 * is returned by the push that completes it; flush != 0 marks end of stream (a frame cut short is then returned the way the reference
 * prints it at EOF, rs41mod.c:2481-2484).  The whole-buffer calls above are create + push(flush) + destroy. */
typedef struct b200fsk_rs41sync b200fsk_rs41sync_t;
int  b200fsk_rs41sync_create(b200fsk_rs41sync_t **out, int nchan, int softinv, int inv, int ecc, int device);
int  b200fsk_rs41sync_push_device(b200fsk_rs41sync_t *h, const float *d_soft, long stride, const int *d_nbits, int nbits, int flush,
                                  b200fsk_rs41_frame_t *h_frames, int max_frames, int *nframes, void *stream);
void b200fsk_rs41sync_destroy(b200fsk_rs41sync_t *h);
const char *b200fsk_framesync_last_error(void);
/* device time (ms, CUDA events) of the frame-sync kernel in this thread's most recent call; < 0 if none */
float b200fsk_framesync_last_kernel_ms(void);

/* fsk_enable_burst_mode / lock_nin of the reference (fsk.c:724: the next nin is not adapted, it stays at its current value) for
 * every channel of the handle; takes effect with the next launch. */
int  b200fsk_set_lock_nin(b200fsk_t *h, int on);
/* the un-normalised eye traces of channel c's last frame (what the reference keeps when fsk_stats_normalise_eye(fsk, 0) was
 * called, fsk.c:869-903); eye must hold B200FSK_EYE_TR * B200FSK_EYE_IND floats, rows of B200FSK_EYE_IND */
int  b200fsk_get_eye_raw(b200fsk_t *h, int c, float *eye);

/* kernel launches issued by this handle so far (bench.py's gpu_launches) */
long b200fsk_launch_count(const b200fsk_t *h);
/* device time (ms) of the most recent demod kernel measured with CUDA events on its own stream; <0 if none */
float b200fsk_last_kernel_ms(b200fsk_t *h);
/* which frame kernel the handle's most recent launch used: 3 = fsk_duo_kernel (one CTA per SM owning one or two channels: the default up to
 * 4 channels per SM), 2 = fsk_wpc_kernel (up to 15 channels per CTA sharing one lane-per-chain oscillator warp: larger batches),
 * 1 = fsk_demod_pipe_kernel (one warp-specialised CTA per channel) and 0 = fsk_demod_kernel (one stage at a time): the modem configurations the
 * first two are not eligible for; -1 before the first launch.  All four produce identical bits.
 * Environment (experiments and tests only): B200FSK_KERNEL=serial|pipe|duo|wpc forces a kernel for handles created afterwards,
 * B200FSK_DUO=0 disables the duo kernel's automatic selection, B200FSK_DUO_G=1|2 its channels per CTA, B200FSK_WPC_THRESH=<n> moves the
 * duo/wpc cross-over, B200FSK_L2PERSIST=0 drops the persisting L2 window of the wpc kernel. */
int b200fsk_last_kernel_kind(const b200fsk_t *h);

#ifdef __cplusplus
}
#endif
#endif
