/* fsk_kernel.cuh -- the batched fsk_demod frame kernel (sm_100a).
 *
 * One CTA owns one channel for the whole launch and walks its stream frame by frame, exactly like
 * the reference's `while(fread(..., fsk_nin()))` loop (utils/fsk_demod.c:281), with the complete
 * per-frame pipeline of fsk_demod_core (utils/fsk.c:593-915) fused in shared memory:
 *
 *   HBM --(coalesced 4/8-byte loads, int16->float /1000 fused)--> x[nin]            (fsk_demod.c:302-309)
 *   A  overlapping Hann-windowed Ndft-point FFTs, one warp per FFT, kiss_fft op order   (fsk.c:465-503)
 *      |X| -> Sf IIR (serial over FFTs, parallel over bins)
 *   B  tone estimate: warp-shuffle argmax (peak) or mask correlation argmax            (fsk.c:507-581)
 *      -> host-built (cos,sin) LUT indexed by bin: no device transcendental
 *   C  per-tone oscillator recurrence (one lane per tone, serial) + down-conversion     (fsk.c:637-656)
 *   D  integrate-and-dump at (Nsym+1)*P offsets, sequential adds per output            (fsk.c:659-668)
 *   E  fine timing: energy x table oscillator, sequential accumulate, exact atan2f     (fsk.c:682-731)
 *   F  re-sample, decide, soft/hard output straight to HBM, EbNo sums                  (fsk.c:737-817)
 *
 * Arithmetic is IEEE-754 single with explicit round-to-nearest intrinsics (never contracted to FMA)
 * and the reference's summation order, so outputs are bit-identical to the reference.
 * Algorithmic HBM traffic: 8 B (f32) or 4 B (cs16) read per sample + Nbits*4/nin B written.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "exact_math.h"

#define FSK_THREADS 256
#define FSK_WARPS   (FSK_THREADS / 32)
#define FSK_MAXSTAGE 8

#ifdef FSK_PROFILE   /* phase timing for profiles/: thread 0 accumulates clock64() deltas per phase */
#define PROF_DECL long long prof_t = clock64(), prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PROF(k) do { if (threadIdx.x == 0) { const long long t_ = clock64(); prof_acc[k] += t_ - prof_t; prof_t = t_; } } while (0)
#define PROF_FLUSH do { if (threadIdx.x == 0 && p.prof) for (int k_ = 0; k_ < 12; ++k_) p.prof[blockIdx.x * 16 + k_] += prof_acc[k_]; } while (0)
#else
#define PROF_DECL
#define PROF(k)
#define PROF_FLUSH
#endif

struct FskFrameRec {            /* per-frame side record (host turns the sums into EbNodB with libm) */
    int   nin_used, nin_next;
    float norm_rx_timing, ppm;
    float mean_sum, std_sum;    /* fsk.c:811,814 */
    int   tabidx[4];            /* LUT index of each tone estimate */
    int   nanflag;              /* fsk.c:702-704 early return taken */
    int   high_sample;
};

struct FskKParams {
    int Ts, M, P, Nsym, N, Nmem, Ndft, Nbits, nfi, ncap, tailcap, xcap;
    int st, en, f_zero, len_mask, est_type;
    int nstage, radix[FSK_MAXSTAGE], log2m[FSK_MAXSTAGE], fstride[FSK_MAXSTAGE];
    float tc, omtc;
    int neyesamp, eye_dec, eye_traces;
    /* shared-memory byte offsets */
    int o_x, o_fdc, o_fint, o_fft, o_mag, o_Sf, o_hann, o_tw, o_perm, o_scr, o_mask, o_smax;
    /* shared-memory byte offsets of the pipelined kernel (fsk_pipe_kernel.cuh) */
    int q_ring, q_fint, q_fft, q_mag, q_Sf, q_hann, q_tw, q_perm, q_scr, q_mask, q_smax, q_tbuf, q_stage_bytes;
    /* tables (global, read-only) */
    const float *hann; const float2 *tw; const unsigned short *perm; const float2 *phi_ft;
    const float *mask; const float2 *dphi_tab; const unsigned short *invperm;
    /* persistent per-channel state (global) */
    float *Sf; float2 *phi_c; float2 *fdc_tail; float *nrt; float *ppm; int *nin;
    unsigned int *last_out; long long *tot_frames; long long *tot_samples;
    /* I/O of this launch */
    const void *in; int fmt; long long in_stride; int *in_off; long long nsamp; int max_frames;
    void *carry_dst; int carry_cap; int *in_off_next;
    void *out; long long out_stride; int soft, softinv; int *frame_base;
    FskFrameRec *recs; long long rec_stride; int write_recs;
    int *nframes; long long *consumed; float *eye_raw;
    float2 *g_ring;             /* pipelined kernel: per-channel oscillator ring in global memory [nchan][M][PIPE_R] */
    int osc_pipelined;          /* pipelined kernel: software-pipelined oscillator stores (wins with one CTA per SM, loses with two) */
    int lock_nin;               /* fsk.c:724: burst_mode / lock_nin -- the next nin is not adapted */
    long long *prof;            /* FSK_PROFILE builds only: per-phase cycle counters [nchan][16] */
    /* throughput kernel (fsk_wpc_kernel.cuh): channels per CTA, ring geometry, shared-memory byte offsets, global f_int */
    int nchan, w_G, w_ring, w_blk;
    int w_hann, w_tw, w_perm, w_mask, w_chan0, w_chan_stride, w_ring_off, w_fft_off, w_Sf_off, w_scr_off;
    /* two-channels-per-CTA kernel (fsk_duo_kernel.cuh): shared-memory byte offsets */
    int d_hann, d_tw, d_mask, d_chan0, d_chan_stride, d_ring_off, d_fint_off, d_fft_off, d_mag_off, d_Sf_off, d_scr_off, d_term_off, d_smax_off, d_tcap, d_phift;
    int   d_G;                      /* duo kernel: channels per CTA in this launch (1 or 2) */
    float2 *w_fint; int *w_abort; int w_fint_slots, w_tcap, w_fp;
};

__device__ __forceinline__ float2 cmul_rn(float2 a, float2 b)       /* comp_prim.h:57-65 */
{
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                       __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cadd_rn(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 csub_rn(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
__device__ __forceinline__ float mag2_rn(float2 v) { return __fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)); }

/* One sample of the input stream -> COMP, the conversions of fsk_demod.c:284-310. */
__device__ __forceinline__ float2 load_sample(const void *in, int fmt, long long idx)
{
    if (fmt == 1) return __ldg((const float2 *)in + idx);
    if (fmt == 0) {
        const short2 s = __ldg((const short2 *)in + idx);
        return make_float2(__fdiv_rn((float)s.x, 1000.0f), __fdiv_rn((float)s.y, 1000.0f));
    }
    if (fmt == 2) {
        const uchar2 u = __ldg((const uchar2 *)in + idx);   /* ((float)u - 127.0)/128.0 is exact in float */
        return make_float2(__fdiv_rn(__fsub_rn((float)u.x, 127.0f), 128.0f), __fdiv_rn(__fsub_rn((float)u.y, 127.0f), 128.0f));
    }
    const short s = __ldg((const short *)in + idx);
    return make_float2(__fdiv_rn((float)s, 1000.0f), 0.0f);
}

template <int FMT> __device__ __forceinline__ float2 load_sample_t(const void *in, long long idx)
{
    if (FMT == 1) return __ldg((const float2 *)in + idx);
    if (FMT == 0) {
        const short2 s = __ldg((const short2 *)in + idx);
        return make_float2(__fdiv_rn((float)s.x, 1000.0f), __fdiv_rn((float)s.y, 1000.0f));
    }
    if (FMT == 2) {
        const uchar2 u = __ldg((const uchar2 *)in + idx);
        return make_float2(__fdiv_rn(__fsub_rn((float)u.x, 127.0f), 128.0f), __fdiv_rn(__fsub_rn((float)u.y, 127.0f), 128.0f));
    }
    const short s = __ldg((const short *)in + idx);
    return make_float2(__fdiv_rn((float)s, 1000.0f), 0.0f);
}

/* The LAST use of a sample (the mixer's read, one frame after the estimator's): evict-first, so that it leaves L2 to the samples that are
 * still waiting for their second read. */
template <int FMT> __device__ __forceinline__ float2 load_sample_last(const void *in, long long idx)
{
    if (FMT == 1) return __ldcs((const float2 *)in + idx);
    if (FMT == 0) {
        const short2 s = __ldcs((const short2 *)in + idx);
        return make_float2(__fdiv_rn((float)s.x, 1000.0f), __fdiv_rn((float)s.y, 1000.0f));
    }
    if (FMT == 2) {
        const uchar2 u = __ldcs((const uchar2 *)in + idx);
        return make_float2(__fdiv_rn(__fsub_rn((float)u.x, 127.0f), 128.0f), __fdiv_rn(__fsub_rn((float)u.y, 127.0f), 128.0f));
    }
    const short s = __ldcs((const short *)in + idx);
    return make_float2(__fdiv_rn((float)s, 1000.0f), 0.0f);
}
/* In-place mixed-radix DIT FFT of one warp's buffer; F already holds the permuted, windowed input.
 * Same butterflies, twiddle indices and operand order as kiss_fft's kf_bfly2 / kf_bfly4 forward
 * (kiss_fft.c:21-42, 44-90), executed stage by stage bottom-up.  PAD inserts one float2 of padding per 8
 * elements (index i -> i + i/8) so that the stride-8/stride-32 butterfly stages do not bank-conflict. */
template <bool PAD> __device__ __forceinline__ int fpad(int i) { return PAD ? i + (i >> 3) : i; }

template <bool PAD>
__device__ __forceinline__ void warp_fft(float2 *F, const float2 *tw, const FskKParams &p, int lane)
{
    const int lgn = 31 - __clz(p.Ndft);
#pragma unroll 1
    for (int st = p.nstage - 1; st >= 0; --st) {
        const int radix = p.radix[st], lm = p.log2m[st], m = 1 << lm, fs = p.fstride[st];
        const int lr = (radix == 2) ? 1 : 2;
        const int nb = 1 << (lgn - lr);
#pragma unroll 1
        for (int b = lane; b < nb; b += 32) {
            const int blk = b >> lm, u = b & (m - 1);
            const int i0 = (blk << (lm + lr)) + u;
            if (radix == 2) {
                float2 *X0 = F + fpad<PAD>(i0), *X1 = F + fpad<PAD>(i0 + m);
                const float2 t = cmul_rn(*X1, tw[fs * u]);
                const float2 x0 = *X0;
                *X1 = csub_rn(x0, t);
                *X0 = cadd_rn(x0, t);
            } else {
                float2 *X0 = F + fpad<PAD>(i0), *X1 = F + fpad<PAD>(i0 + m), *X2 = F + fpad<PAD>(i0 + 2 * m), *X3 = F + fpad<PAD>(i0 + 3 * m);
                const float2 s0 = cmul_rn(*X1, tw[fs * u]);
                const float2 s1 = cmul_rn(*X2, tw[fs * u * 2]);
                const float2 s2 = cmul_rn(*X3, tw[fs * u * 3]);
                const float2 x0 = *X0;
                const float2 s5 = csub_rn(x0, s1);
                const float2 a  = cadd_rn(x0, s1);
                const float2 s3 = cadd_rn(s0, s2);
                const float2 s4 = csub_rn(s0, s2);
                *X2 = csub_rn(a, s3);
                *X0 = cadd_rn(a, s3);
                *X1 = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));
                *X3 = make_float2(__fsub_rn(s5.x, s4.y), __fadd_rn(s5.y, s4.x));
            }
        }
        __syncwarp();
    }
}

/* The same FFT on NB independent buffers at once: the butterflies of different buffers are independent, so their
 * shared-memory loads and FP32 chains overlap (one warp per FFT is latency-bound: ~10 cycles per instruction). */
template <bool PAD, int NB>
__device__ __forceinline__ void warp_fft_multi(float2 *const (&F)[NB], const float2 *tw, const FskKParams &p, int lane)
{
    const int lgn = 31 - __clz(p.Ndft);
#pragma unroll 1
    for (int st = p.nstage - 1; st >= 0; --st) {
        const int radix = p.radix[st], lm = p.log2m[st], m = 1 << lm, fs = p.fstride[st];
        const int lr = (radix == 2) ? 1 : 2;
        const int nb = 1 << (lgn - lr);
#pragma unroll 1
        for (int b = lane; b < nb; b += 32) {
            const int blk = b >> lm, u = b & (m - 1);
            const int i0 = (blk << (lm + lr)) + u;
            if (radix == 2) {
                const int a0 = fpad<PAD>(i0), a1 = fpad<PAD>(i0 + m);
                const float2 w = tw[fs * u];
                float2 x0[NB], x1[NB];
#pragma unroll
                for (int q = 0; q < NB; ++q) { x0[q] = F[q][a0]; x1[q] = F[q][a1]; }
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const float2 t = cmul_rn(x1[q], w);
                    F[q][a1] = csub_rn(x0[q], t);
                    F[q][a0] = cadd_rn(x0[q], t);
                }
            } else {
                const int a0 = fpad<PAD>(i0), a1 = fpad<PAD>(i0 + m), a2 = fpad<PAD>(i0 + 2 * m), a3 = fpad<PAD>(i0 + 3 * m);
                const float2 w1 = tw[fs * u], w2 = tw[fs * u * 2], w3 = tw[fs * u * 3];
                float2 x0[NB], x1[NB], x2[NB], x3[NB];
#pragma unroll
                for (int q = 0; q < NB; ++q) { x0[q] = F[q][a0]; x1[q] = F[q][a1]; x2[q] = F[q][a2]; x3[q] = F[q][a3]; }
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const float2 s0 = cmul_rn(x1[q], w1);
                    const float2 s1 = cmul_rn(x2[q], w2);
                    const float2 s2 = cmul_rn(x3[q], w3);
                    const float2 s5 = csub_rn(x0[q], s1);
                    const float2 a  = cadd_rn(x0[q], s1);
                    const float2 s3 = cadd_rn(s0, s2);
                    const float2 s4 = csub_rn(s0, s2);
                    F[q][a2] = csub_rn(a, s3);
                    F[q][a0] = cadd_rn(a, s3);
                    F[q][a1] = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));
                    F[q][a3] = make_float2(__fsub_rn(s5.x, s4.y), __fadd_rn(s5.y, s4.x));
                }
            }
        }
        __syncwarp();
    }
}

/* Tone estimate of one frame by ONE warp, from the final Sf of that frame: the peak picker (fsk.c:507-547) or
 * the mask correlator (fsk.c:566-581), whichever fsk_set_freq_est_alg selected.  Returns, for each tone, the
 * index into the host-built (cos,sin) oscillator-step LUT.  Ties resolve to the lowest bin, as the
 * reference's strict `>` scan does. */
__device__ __forceinline__ void estimate_tones(const FskKParams &p, const float *s_Sf, float *scr, const float *s_mask,
                                               int lane, int tabidx[4])
{
    const int M = p.M, Ndft = p.Ndft;
    tabidx[0] = tabidx[1] = tabidx[2] = tabidx[3] = 0;
    if (p.est_type == 0) {
        for (int i = lane; i < Ndft; i += 32) scr[i] = s_Sf[i];
        __syncwarp();
        int freqi[4] = {0, 0, 0, 0};
        for (int k = 0; k < M; ++k) {
            float mx = 0.0f; int im = 0;
            for (int j = p.st + lane; j < p.en; j += 32) { const float v = scr[j]; if (v > mx) { mx = v; im = j; } }
            for (int off = 16; off; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, mx, off);
                const int   oi = __shfl_xor_sync(0xffffffffu, im, off);
                if (ov > mx || (ov == mx && oi < im)) { mx = ov; im = oi; }
            }
            const int f_min = max(im - p.f_zero, 0), f_max = min(im + p.f_zero, Ndft);
            for (int j = f_min + lane; j < f_max; j += 32) scr[j] = 0.0f;
            __syncwarp();
            freqi[k] = im - Ndft / 2;
        }
        for (int i = 1; i < M;) {                               /* gnome sort, fsk.c:533-542 */
            if (freqi[i] >= freqi[i - 1]) ++i;
            else { const int t = freqi[i]; freqi[i] = freqi[i - 1]; freqi[i - 1] = t; if (i > 1) --i; }
        }
        for (int m = 0; m < M; ++m) tabidx[m] = freqi[m] + Ndft / 2;
    } else {
        float cm = 0.0f; int bm = p.st;
        for (int b = p.st + lane; b < p.en - p.len_mask; b += 32) {
            float corr = 0.0f;
            for (int i = 0; i < p.len_mask; ++i) corr = __fadd_rn(corr, __fmul_rn(s_mask[i], s_Sf[b + i]));
            if (corr > cm) { cm = corr; bm = b; }
        }
        for (int off = 16; off; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, cm, off);
            const int   ob = __shfl_xor_sync(0xffffffffu, bm, off);
            if (ov > cm || (ov == cm && ob < bm)) { cm = ov; bm = ob; }
        }
        for (int m = 0; m < M; ++m) tabidx[m] = bm * M + m;
    }
}

__global__ void __launch_bounds__(FSK_THREADS) fsk_demod_kernel(const FskKParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ float2 s_phi[4];
    __shared__ int    s_tabidx[4];
    __shared__ float  s_tc[2];
    __shared__ int    s_nin, s_low, s_high, s_nan;
    __shared__ float  s_nrt, s_ppm, s_fract;

    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = p.M, Ndft = p.Ndft, Nmem = p.Nmem, nfi = p.nfi, P = p.P, Ts = p.Ts, nsym = p.Nsym;

    float2 *x      = (float2 *)(smem + p.o_x);        /* raw frame, later the timing terms */
    float2 *fdc    = (float2 *)(smem + p.o_fdc);      /* [M][Nmem] */
    float2 *fint   = (float2 *)(smem + p.o_fint);     /* [M][nfi]  */
    float2 *fftbuf = (float2 *)(smem + p.o_fft);      /* [warps][Ndft] */
    float  *magbuf = (float *)(smem + p.o_mag);       /* [warps][Ndft] */
    float  *s_Sf   = (float *)(smem + p.o_Sf);
    float  *s_hann = (float *)(smem + p.o_hann);
    float2 *s_tw   = (float2 *)(smem + p.o_tw);
    unsigned short *s_perm = (unsigned short *)(smem + p.o_perm);
    float  *scr    = (float *)(smem + p.o_scr);
    float  *s_mask = (float *)(smem + p.o_mask);
    float2 *s_maxs = (float2 *)(smem + p.o_smax);     /* per symbol (max, sqrt(max)) */

    /* ---- launch prologue: tables + persistent state -> shared memory ---- */
    for (int i = tid; i < Ndft; i += FSK_THREADS) {
        s_hann[i] = p.hann[i]; s_tw[i] = p.tw[i]; s_perm[i] = p.perm[i]; s_mask[i] = p.mask[i];
        s_Sf[i] = p.Sf[(long long)c * Ndft + i];
    }
    for (int e = tid; e < M * p.tailcap; e += FSK_THREADS) {
        const int m = e / p.tailcap, i = e - m * p.tailcap;
        fdc[m * Nmem + Nmem - p.tailcap + i] = p.fdc_tail[((long long)c * M + m) * p.tailcap + i];
    }
    if (tid < M) s_phi[tid] = p.phi_c[c * 4 + tid];
    if (tid == 0) { s_nin = p.nin[c]; s_nrt = p.nrt[c]; s_ppm = p.ppm[c]; s_nan = 0; s_high = 0; }
    __syncthreads();

    const long long base = (long long)c * p.in_stride;
    const long long start = p.in_off ? (long long)p.in_off[c] : 0;
    long long cursor = start;
    const int fbase = p.frame_base ? p.frame_base[c] : 0;
    int frames = 0;
    unsigned char *outb = (unsigned char *)p.out + (long long)c * p.out_stride * (p.soft ? 4 : 1);

    PROF_DECL
    while (true) {
        const int nin = s_nin;
        if ((p.max_frames > 0 && frames >= p.max_frames) || cursor + nin > p.nsamp) break;
        const int nold = Nmem - nin;

        /* ---- load + convert the frame (coalesced) ---- */
        for (int i = tid; i < nin; i += FSK_THREADS) x[i] = load_sample(p.in, p.fmt, base + cursor + i);
        __syncthreads();
        PROF(0);

        /* ---- A: overlapping windowed FFTs -> Sf (fsk.c:465-503) ---- */
        const int numffts = (int)floorf(__fdiv_rn((float)nin, (float)(Ndft / 2))) - 1;
        for (int r0 = 0; r0 < numffts; r0 += FSK_WARPS) {
            const int j = r0 + warp;
            if (j < numffts) {
                float2 *F = fftbuf + warp * Ndft;
                const int a = j * Ndft / 2;
                for (int pos = lane; pos < Ndft; pos += 32) {
                    const int idx = s_perm[pos];
                    const float h = s_hann[idx];
                    const float2 v = x[a + idx];
                    F[pos] = make_float2(__fmul_rn(h, v.x), __fmul_rn(h, v.y));
                }
                __syncwarp();
                warp_fft<false>(F, s_tw, p, lane);
                for (int i = lane; i < Ndft; i += 32) magbuf[warp * Ndft + i] = __fsqrt_rn(mag2_rn(F[i]));
            }
            __syncthreads();
            const int cnt = min(FSK_WARPS, numffts - r0);
            for (int i = tid; i < Ndft; i += FSK_THREADS) {
                float s = s_Sf[i];
                const int src = (i + Ndft / 2) & (Ndft - 1);           /* fftshift, fsk.c:484-489 */
                for (int jj = 0; jj < cnt; ++jj)
                    s = __fadd_rn(__fmul_rn(s, p.omtc), __fmul_rn(magbuf[jj * Ndft + src], p.tc));   /* fsk.c:500 */
                s_Sf[i] = s;
            }
            __syncthreads();
        }

        PROF(1);
        /* ---- B + C(serial part): estimator, integrator memory shift, oscillators: warp 0 ---- */
        if (warp == 0) {
            int tabidx[4];
            estimate_tones(p, s_Sf, scr, s_mask, lane, tabidx);
            PROF(2);
            /* integrator memory shift (fsk.c:637-640): source and destination never overlap (Nsym >= 4) */
            for (int e = lane; e < M * nold; e += 32) {
                const int m = e / nold, i = e - m * nold;
                fdc[m * Nmem + i] = fdc[m * Nmem + nin + i];
            }
            /* per-tone local oscillator (fsk.c:644-651): a serial recurrence, one lane per tone */
            if (lane < M) {
                const int ti = tabidx[lane];
                s_tabidx[lane] = ti;
                const float2 d = __ldg(p.dphi_tab + ti);
                float2 ph = s_phi[lane];
                float2 *dst = fdc + lane * Nmem + nold;
#pragma unroll 4
                for (int j = 0; j < nin; ++j) { ph = cmul_rn(ph, d); dst[j] = ph; }
                const float av = __fsqrt_rn(mag2_rn(ph));               /* comp_normalize, comp_prim.h:133-139 */
                s_phi[lane] = make_float2(__fdiv_rn(ph.x, av), __fdiv_rn(ph.y, av));
            }
        }
        __syncthreads();
        PROF(3);

        /* ---- C(parallel part): down-convert, f_dc = x * conj(phi)  (fsk.c:648) ---- */
        for (int i = tid; i < nin; i += FSK_THREADS) {
            const float2 v = x[i];
            for (int m = 0; m < M; ++m) {
                const float2 ph = fdc[m * Nmem + nold + i];
                fdc[m * Nmem + nold + i] = cmul_rn(v, make_float2(ph.x, -ph.y));
            }
        }
        __syncthreads();
        PROF(4);

        /* ---- D: integrate and dump (fsk.c:659-668) ---- */
        for (int e = tid; e < M * nfi; e += FSK_THREADS) {
            const int m = e / nfi, i = e - m * nfi;
            const float2 *src = fdc + m * Nmem + (i * Ts) / P;
            float2 acc = make_float2(0.0f, 0.0f);
            for (int j = 0; j < Ts; ++j) acc = cadd_rn(acc, src[j]);
            fint[e] = acc;
        }
        __syncthreads();
        PROF(5);

        /* ---- E: fine timing (fsk.c:682-731) ---- */
        for (int i = tid; i < nfi; i += FSK_THREADS) {
            float ft1 = 0.0f;
            for (int m = 0; m < M; ++m) ft1 = __fadd_rn(ft1, mag2_rn(fint[m * nfi + i]));
            const float2 ph = __ldg(p.phi_ft + i);
            x[i] = make_float2(__fmul_rn(ft1, ph.x), __fmul_rn(ft1, ph.y));
        }
        __syncthreads();
        PROF(6);
        if (warp == 0 && lane < 2) {                                    /* sequential accumulate: re / im lanes */
            const float *t = (const float *)x + lane;
            float acc = 0.0f;
#pragma unroll 8
            for (int i = 0; i < nfi; ++i) acc = __fadd_rn(acc, t[2 * i]);
            s_tc[lane] = acc;
        }
        __syncthreads();
        PROF(7);
        if (tid == 0) {
            const float tr = s_tc[0], ti = s_tc[1];
            int nan = 0, nin_next = nin, high = s_high;
            if (isnan(tr) || isnan(ti)) nan = 1;                        /* fsk.c:702-704 */
            else {
                const float nrt = (float)__ddiv_rn((double)em_atan2f(ti, tr), 6.283185307179586476925286766559);
                const float rx_timing = __fmul_rn(nrt, (float)P);
                const float d = __fsub_rn(nrt, s_nrt);
                s_nrt = nrt;
                if (fabsf(d) < 0.2f) {                                  /* (double)|d| < .2  <=>  |d| < 0.2f */
                    const float appm = (float)__ddiv_rn(__dmul_rn(1e6, (double)d), (double)(float)nsym);
                    s_ppm = (float)__dadd_rn(__dmul_rn(.9, (double)s_ppm), __dmul_rn(.1, (double)appm));
                }
                if (p.lock_nin) nin_next = nin;
                else if (nrt > 0.25f) nin_next = p.N + Ts / 2;
                else if (nrt < -0.25f) nin_next = p.N - Ts / 2;
                else nin_next = p.N;
                const int low = (int)floorf(rx_timing);
                s_low = low; s_fract = __fsub_rn(rx_timing, (float)low);
                high = (int)ceilf(rx_timing); s_high = high;
            }
            s_nan = nan; s_nin = nin_next;
        }
        __syncthreads();
        PROF(8);

        /* ---- F: re-sample, decide, write (fsk.c:751-817) ---- */
        const long long ofs = (long long)(fbase + frames) * p.Nbits;
        if (!s_nan) {
            const int low = s_low, high = s_high; const float fract = s_fract, omf = __fsub_rn(1.0f, fract);
            for (int i = tid; i < nsym; i += FSK_THREADS) {
                const int st = (i + 1) * P;
                float tmax[4];
                for (int m = 0; m < M; ++m) {
                    const float2 a = fint[m * nfi + st + low], b = fint[m * nfi + st + high];
                    const float2 t = make_float2(__fadd_rn(__fmul_rn(omf, a.x), __fmul_rn(fract, b.x)),
                                                 __fadd_rn(__fmul_rn(omf, a.y), __fmul_rn(fract, b.y)));
                    tmax[m] = mag2_rn(t);
                }
                float mx = tmax[0]; int sym = 0;
                for (int m = 1; m < M; ++m) if (tmax[m] > mx) { mx = tmax[m]; sym = m; }
                s_maxs[i] = make_float2(mx, __fsqrt_rn(mx));
                if (!p.soft) {
                    if (M == 2) outb[ofs + i] = (unsigned char)(sym == 1);
                    else { outb[ofs + 2 * i + 1] = (unsigned char)(sym & 1); outb[ofs + 2 * i] = (unsigned char)((sym & 2) >> 1); }
                } else {
                    float *o = (float *)outb + ofs;
                    for (int m = 0; m < M; ++m) tmax[m] = __fsqrt_rn(tmax[m]);
                    if (M == 2) { const float v = __fsub_rn(tmax[0], tmax[1]); o[i] = p.softinv ? -v : v; }
                    else {                                              /* fsk.c:796-803 */
                        float a = -tmax[0], b = -tmax[0];
                        a = __fadd_rn(a, tmax[1]);  b = __fadd_rn(b, -tmax[1]);
                        a = __fadd_rn(a, -tmax[2]); b = __fadd_rn(b, tmax[2]);
                        a = __fadd_rn(a, tmax[3]);  b = __fadd_rn(b, tmax[3]);
                        o[2 * i + 1] = p.softinv ? -a : a; o[2 * i] = p.softinv ? -b : b;
                    }
                }
            }
        } else {
            /* stale frame: the reference re-emits its (already inverted) previous buffer (fsk_demod.c:424-433) */
            for (int i = tid; i < p.Nbits; i += FSK_THREADS) {
                if (p.soft) {
                    float *o = (float *)outb;
                    const float prev = (fbase + frames) > 0 ? o[ofs - p.Nbits + i] : __uint_as_float(p.last_out[(long long)c * p.Nbits + i]);
                    o[ofs + i] = p.softinv ? -prev : prev;
                } else {
                    outb[ofs + i] = (fbase + frames) > 0 ? outb[ofs - p.Nbits + i] : (unsigned char)p.last_out[(long long)c * p.Nbits + i];
                }
            }
        }
        __syncthreads();
        PROF(9);

        /* ---- EbNo sums (fsk.c:808-815) + frame record: warp 1 lanes 0/1, sequential ---- */
        if (warp == 1) {
            float acc = 0.0f;
            if (lane < 2 && !s_nan) {
                const float *t = (const float *)s_maxs + lane;
#pragma unroll 4
                for (int i = 0; i < nsym; ++i) acc = __fadd_rn(acc, t[2 * i]);
            }
            const float mean_sum = __shfl_sync(0xffffffffu, acc, 1);
            if (lane == 0 && p.write_recs) {
                FskFrameRec r;
                r.nin_used = nin; r.nin_next = s_nin; r.norm_rx_timing = s_nrt; r.ppm = s_ppm;
                r.mean_sum = mean_sum; r.std_sum = acc; r.nanflag = s_nan; r.high_sample = s_high;
                for (int m = 0; m < 4; ++m) r.tabidx[m] = m < M ? s_tabidx[m] : 0;
                p.recs[(long long)c * p.rec_stride + fbase + frames] = r;
            }
        }
        cursor += nin; ++frames;
        __syncthreads();
        PROF(10);
    }
    PROF_FLUSH;

    /* ---- launch epilogue: state back to HBM ---- */
    for (int i = tid; i < Ndft; i += FSK_THREADS) p.Sf[(long long)c * Ndft + i] = s_Sf[i];
    for (int e = tid; e < M * p.tailcap; e += FSK_THREADS) {
        const int m = e / p.tailcap, i = e - m * p.tailcap;
        p.fdc_tail[((long long)c * M + m) * p.tailcap + i] = fdc[m * Nmem + Nmem - p.tailcap + i];
    }
    if (tid < M) p.phi_c[c * 4 + tid] = s_phi[tid];
    if (tid == 0) {
        p.nin[c] = s_nin; p.nrt[c] = s_nrt; p.ppm[c] = s_ppm;
        p.nframes[c] = fbase + frames; p.consumed[c] = cursor - start;
        p.tot_frames[c] += frames; p.tot_samples[c] += cursor - start;
        if (p.frame_base) p.frame_base[c] = fbase + frames;
    }
    if (frames > 0) {
        const long long ofs = (long long)(fbase + frames - 1) * p.Nbits;
        for (int i = tid; i < p.Nbits; i += FSK_THREADS)
            p.last_out[(long long)c * p.Nbits + i] = p.soft ? __float_as_uint(((const float *)outb)[ofs + i]) : (unsigned int)outb[ofs + i];
        if (p.write_recs && !s_nan) {                                   /* raw eye samples of the last frame (fsk.c:869-890) */
            const int neyeoffset = s_high + 1, tot = p.eye_traces * M * p.neyesamp;
            for (int e = tid; e < tot; e += FSK_THREADS) {
                const int row = e / p.neyesamp, j = e - row * p.neyesamp, i = row / M, m = row - i * M;
                const int ind = 2 * P * i + neyeoffset + j * p.eye_dec, flat = m * nfi + ind;
                float2 v = make_float2(0.0f, 0.0f);
                if (flat >= 0 && ind < nfi) v = fint[flat];
                p.eye_raw[(long long)c * (8 * 160) + row * 160 + j] = __fsqrt_rn(mag2_rn(v));
            }
        }
    }
    /* ---- carry the unconsumed tail to the end of the next staging buffer's carry area (push mode).
     *      The destination lies left of the source, so a chunked copy staged through shared memory is
     *      safe even when both live in the same buffer. ---- */
    if (p.carry_cap > 0) {
        const long long tail = p.nsamp - cursor;
        const int es = p.fmt == 1 ? 8 : (p.fmt == 0 ? 4 : 2);
        const long long nhalf = tail * es / 2;                          /* copy in 2-byte units */
        const int chunk = p.xcap * 4;                                   /* xcap float2 = xcap*4 halfwords */
        unsigned short *stage = (unsigned short *)(smem + p.o_x);
        const unsigned short *src = (const unsigned short *)((const unsigned char *)p.in + (base + cursor) * es);
        unsigned short *dst = (unsigned short *)((unsigned char *)p.carry_dst + (base + p.carry_cap - tail) * es);
        for (long long k0 = 0; k0 < nhalf; k0 += chunk) {
            const int n = (int)min((long long)chunk, nhalf - k0);
            __syncthreads();
            for (int i = tid; i < n; i += FSK_THREADS) stage[i] = src[k0 + i];
            __syncthreads();
            for (int i = tid; i < n; i += FSK_THREADS) dst[k0 + i] = stage[i];
        }
        if (tid == 0) p.in_off_next[c] = (int)(p.carry_cap - tail);
    }
}
