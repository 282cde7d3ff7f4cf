/* dft_detect.cu -- libb200dft.so: batched sonde-type detector on B200 (C-ABI of include/b200dft.h).
 *
 * Reference: scan/dft_detect.c (production IQ mode, `--iq --bw B --dc`).  The reference runs, per candidate
 * frequency, one single-threaded process that (a) removes the IQ DC, low-pass filters the IF, FM-demodulates
 * (dft_detect.c:556-591, 709-716, 780-811) and (b) every K-4 samples cross-correlates the last K+L_j FM
 * samples with 14 matched header templates through three 8192-point FFTs per header (getCorrDFT, :357-443,
 * 93 % of its CPU time), then (c) applies thresholds / header bit checks (:1514-1652).
 *
 * Here the batch of channels is processed by three kernels:
 *   dft_dcavg_kernel   block means of the raw IQ (the DC estimate the reference refreshes every sr/32 samples)
 *   dft_front_kernel   fused  int16->float, DC subtraction, 49-tap IF FIR, FM discriminator x2  -> two FM streams
 *                      in HBM (filtered: rings 0-2 of the reference; unfiltered: ring 3)
 *   dft_corr_kernel    one CTA per (channel, block, header): window -> DC estimate -> 97-tap FM low-pass ->
 *                      DIRECT time-domain matched-filter correlation with register-tiled sliding dot products
 *                      (the FFTs of the reference are replaced outright: same circular-convolution definition,
 *                      evaluated exactly, FP32 FMA) -> peak (first maximum of corr^2), normalisation, and, for
 *                      threshold crossings, the header bit re-slice (headcmp) and the M10/M20 frame bytes.
 * The per-channel decision loop of main() (mv_pos ordering, -c, exit code) runs on the host over the tiny
 * per-block result records.  No CPU fallback exists for the DSP.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>
#include "../../include/b200dft.h"

#define NRS 18
#define IDX_IMETAFSK 15
#define IDX_RS 16
#define IDX_I4 17
#define SKIP_A 11                /* C34C50:  -DNOC34C50  (scan/Makefile:1)  */
#define SKIP_B 14                /* IMET1AB: -DNOIMET1AB (scan/Makefile:1)  */
#define FM_GAIN 0.8
#define CORR_THREADS 384
#define CORR_R 16                /* outputs per thread (register tile of the sliding dot product) */
#define FRONT_TILE 512

static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...)
{
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf; return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(B200DFT_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)
extern "C" const char *b200dft_last_error(void) { return g_err.c_str(); }

struct HdrDef { int sps; const char *header; float BT, thres; int herrs; const char *type; int tn, lpFM, lpIQ; };
/* the header bank: strings scan/dft_detect.c:43-126, table :177-194 */
static const HdrDef HDR[NRS] = {
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

struct BlkRes {                  /* getCorrDFT outputs of one (channel, block, header) + candidate check */
    float score; unsigned pos; float dc; int mp; int herrs; unsigned m10;
};

struct HdrDev {                  /* per active header, device side */
    int j, L, Lpad, hLen, use_raw, lpFM, match_off;     /* Lpad = L rounded up to 8; match_off into the match table */
    float spb, thres;
    unsigned char bits[96];     /* header bits 0/1 */
};

struct DftParams {
    int nchan, nact, N, K, L, lpIQtaps, lpFMtaps, tpad, delay, opt_dc, maxcnt, nblk, navg;
    long long nsamp, in_stride;
    int fmt;
    const void *iq;
    const float *ws_lpIQ;        /* lpIQtaps */
    const float *ws_lpFM;        /* [2][lpFMtaps] */
    const float *match;          /* concatenated, each padded to a multiple of 8 with zeros */
    const HdrDev *hdr;           /* nact */
    float2 *avg;                 /* [nchan][navg] DC estimate valid for DC-block i */
    float *fm_filt, *fm_raw;     /* [nchan][nsamp] */
    BlkRes *res;                 /* [nz][nblk][nact] */
    const int *chan_list;        /* [nz] channel of grid.z entry (NULL: identity)                                   */
    const long long *sched_off;  /* [nz] samples consumed outside the block counter before this schedule (iMet branch) */
    long long pos_base;          /* streaming: absolute index of sample 0 of the buffers (added to reported positions) */
};

/* raw sample -> float pair, the conversions of f32read_csample (dft_detect.c:561-578) */
__device__ __forceinline__ float2 dft_load(const void *iq, int fmt, long long idx)
{
    if (fmt == 1) return __ldg((const float2 *)iq + idx);
    /* x/32768.0 and (u-128)/128.0 rounded to float (:568-576) are exact power-of-two scalings of small integers: a float multiply is bit-identical */
    if (fmt == 0) { const short2 s = __ldg((const short2 *)iq + idx); return make_float2((float)s.x * (1.0f / 32768.0f), (float)s.y * (1.0f / 32768.0f)); }
    const uchar2 u = __ldg((const uchar2 *)iq + idx);
    return make_float2((float)((int)u.x - 128) * (1.0f / 128.0f), (float)((int)u.y - 128) * (1.0f / 128.0f));
}

/* ---- K0: mean of every block of maxcnt raw samples (IQdc, dft_detect.c:583-588); avg[i+1] applies to block i+1 ---- */
__global__ void dft_dcavg_kernel(const DftParams p)
{
    const int c = blockIdx.y, blk = blockIdx.x;
    const long long s0 = (long long)blk * p.maxcnt;
    double sx = 0.0, sy = 0.0;
    for (int i = threadIdx.x; i < p.maxcnt; i += blockDim.x) {
        const long long n = s0 + i;
        if (n < p.nsamp) { const float2 v = dft_load(p.iq, p.fmt, (long long)c * p.in_stride + n); sx += v.x; sy += v.y; }
    }
    __shared__ double shx[4], shy[4];
    for (int o = 16; o; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
    if ((threadIdx.x & 31) == 0) { shx[threadIdx.x >> 5] = sx; shy[threadIdx.x >> 5] = sy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        sx = shx[0] + shx[1] + shx[2] + shx[3]; sy = shy[0] + shy[1] + shy[2] + shy[3];
        float2 a = make_float2((float)(sx / (double)(float)p.maxcnt), (float)(sy / (double)(float)p.maxcnt));
        if (s0 + p.maxcnt > p.nsamp) a = make_float2(0.f, 0.f);            /* incomplete block: never becomes an estimate */
        p.avg[(long long)c * p.navg + blk + 1] = a;
        if (blk == 0) p.avg[(long long)c * p.navg] = make_float2(0.f, 0.f);
    }
}


/* ---- KW: wideband front (dft_detect --IQ <fq>): DC block on the shared capture, rotate by the per-channel LUT phase, decimating FIR
 *          (dft_detect.c:751-761 with the tables of :1021-1111).  One CTA = one channel x CH_TM outputs; the rotated window is
 *          staged once in shared memory, each warp then forms whole outputs (lanes stride the taps, conflict-free float2 reads). */
#define CH_THREADS 256
#define CH_OW 4                                  /* outputs a warp accumulates together (shared coefficient load, 8 independent FMA chains) */
struct ChanParams {
    const void *wide; int fmt; long long nwide;
    const float2 *wavg; int wmaxcnt;          /* IQdc means of the wide stream (maxcnt = IF_sr/32*decM, :1155-1156) */
    int decM, taps, lut_len, TM;
    const float *ws;                          /* decimation low-pass (:1043-1058) */
    const double *f0;                         /* [nchan] freq0/sr_base of the rotation LUT (:1075-1096) */
    float2 *out; long long nout, out_stride;
};

/* (cos,sin)(t*2*pi) for the LUT entry ex[n], t = f0*n, formed like the reference's cexp(t*2*M_PI*I) (:1100): the double argument is
 * reduced with a three-term 2*pi (FMA keeps each product exact), then the fast-path sincos. */
__device__ __forceinline__ double2 lut_cs(double f0, int n)
{
    const double t = f0 * (double)n;
    const double arg = (t * 2) * M_PI;
    const double k = rint(arg * 0.15915494309189535);
    double r = fma(-k, 6.283185307179586232e+00, arg);
    r = fma(-k, 2.449293598294706414e-16, r);
    r = fma(-k, -5.989539619436679332e-33, r);
    double sn, cs;
    sincos(r, &sn, &cs);
    return make_double2(cs, sn);
}

__global__ void __launch_bounds__(CH_THREADS) dft_chan_kernel(const ChanParams q)
{
    extern __shared__ float2 xs[];
    const int c = blockIdx.y, decM = q.decM, taps = q.taps;
    const long long m0 = (long long)blockIdx.x * q.TM;
    const int tm = (int)min((long long)q.TM, q.nout - m0);
    const int W = (tm - 1) * decM + taps;
    float *wsm = (float *)(xs + q.TM * decM + taps);               /* (window padded by one decM row) */
    const long long g0 = m0 * decM + decM - taps;                /* wide index of xs[0]; output m ends at wide sample m*decM+decM-1 */
    const double f0 = q.f0[c];
    for (int i = threadIdx.x; i < taps; i += CH_THREADS) wsm[i] = q.ws[i];
    /* Rotation: thread t owns window samples ib+t, ib+t+256, ...  Its LUT index advances by 256 (mod lut_len) per step, so after one full
       evaluation the entry is carried by a double-precision angle addition (error ~1e-16 per step, < 60 steps; the value is then rounded to
       float like the table).  On a LUT wrap it is re-evaluated, which also reproduces the table's phase jump when freq0 is not a multiple
       of the divisor d (:1082-1091 may leave it unsnapped).  64-bit divisions happen once per thread, not per sample. */
    const long long gb = g0 < 0 ? 0 : g0;
    const int ib = (int)(gb - g0);
    for (int i = threadIdx.x; i < ib; i += CH_THREADS) xs[i] = make_float2(0.f, 0.f);       /* before the stream: the calloc'ed ring (:1104) */
    for (int i = W + threadIdx.x; i < q.TM * decM + taps; i += CH_THREADS) xs[i] = make_float2(0.f, 0.f);
    const bool carry = q.lut_len > 2 * CH_THREADS;
    const double2 st = lut_cs(f0, CH_THREADS);                   /* e^{j 2 pi f0 256} */
    long long gq = (gb + threadIdx.x) / q.wmaxcnt;               /* DC-block index of this thread's sample ... */
    int gr = (int)((gb + threadIdx.x) % q.wmaxcnt);              /* ... and position inside that block */
    int n = (int)((gb + threadIdx.x) % q.lut_len);
    double2 e = lut_cs(f0, n);
    int i = ib + threadIdx.x;
    float2 rn = i < W ? dft_load(q.wide, q.fmt, g0 + i) : make_float2(0.f, 0.f);
    for (; i < W; i += CH_THREADS) {
        const float2 r = rn;
        if (i + CH_THREADS < W) rn = dft_load(q.wide, q.fmt, g0 + i + CH_THREADS);     /* next sample in flight while this one is rotated */
        const float2 a = q.wavg[gq];
        const float x = r.x - a.x, y = r.y - a.y;                                      /* :627-628 */
        const float ex = (float)e.x, ey = (float)e.y;
        xs[i] = make_float2(x * ex - y * ey, x * ey + y * ex);                         /* decMbuf[j]*ex[] (:755) */
        gr += CH_THREADS; while (gr >= q.wmaxcnt) { gr -= q.wmaxcnt; ++gq; }
        n += CH_THREADS;
        if (carry && n < q.lut_len) e = make_double2(e.x * st.x - e.y * st.y, e.x * st.y + e.y * st.x);
        else { n %= q.lut_len; e = lut_cs(f0, n); }
    }
    __syncthreads();
    float2 *out = q.out + (long long)c * q.out_stride + m0;
    /* FIR: a warp forms CH_OW consecutive outputs at a time; lanes stride the taps (consecutive float2 words: conflict-free), one
       coefficient load feeds 2*CH_OW independent FMA chains.  Outputs past tm read the zero-filled pad and are not stored. */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = warp * CH_OW; o < tm; o += (CH_THREADS / 32) * CH_OW) {
        const float2 *xp = xs + o * decM;
        const bool full = o + CH_OW <= tm;
        float ax[CH_OW], ay[CH_OW];
#pragma unroll
        for (int u = 0; u < CH_OW; ++u) { ax[u] = 0.f; ay[u] = 0.f; }
        if (full) {
            for (int k = lane; k < taps; k += 32) {
                const float w = wsm[k];
#pragma unroll
                for (int u = 0; u < CH_OW; ++u) { const float2 v = xp[u * decM + k]; ax[u] = fmaf(w, v.x, ax[u]); ay[u] = fmaf(w, v.y, ay[u]); }
            }
        } else {
            for (int k = lane; k < taps; k += 32) {
                const float w = wsm[k];
#pragma unroll
                for (int u = 0; u < CH_OW; ++u) if (o + u < tm) { const float2 v = xp[u * decM + k]; ax[u] = fmaf(w, v.x, ax[u]); ay[u] = fmaf(w, v.y, ay[u]); }
            }
        }
#pragma unroll
        for (int u = 0; u < CH_OW; ++u) {
            for (int d = 16; d; d >>= 1) { ax[u] += __shfl_xor_sync(0xffffffffu, ax[u], d); ay[u] += __shfl_xor_sync(0xffffffffu, ay[u], d); }
            if (lane == 0 && o + u < tm) out[o + u] = make_float2(ax[u], ay[u]);
        }
    }
}

/* ---- K1: fused front end: DC subtraction -> IF FIR -> FM discriminator (filtered + unfiltered) ------------------ */
__global__ void __launch_bounds__(FRONT_TILE) dft_front_kernel(const DftParams p)
{
    extern __shared__ float2 sm[];
    const int taps = p.lpIQtaps;
    float2 *z = sm;                                  /* [FRONT_TILE + taps]  : samples t0-taps .. t0+TILE-1 */
    float2 *y = sm + FRONT_TILE + taps;              /* [FRONT_TILE + 1]     : filtered t0-1 .. t0+TILE-1    */
    __shared__ float ws[256];
    const int c = blockIdx.y;
    const long long t0 = (long long)blockIdx.x * FRONT_TILE;
    for (int i = threadIdx.x; i < taps; i += FRONT_TILE) ws[i] = p.ws_lpIQ[i];
    for (int i = threadIdx.x; i < FRONT_TILE + taps; i += FRONT_TILE) {
        const long long n = t0 - taps + i;
        float2 v = make_float2(0.f, 0.f);
        if (n >= 0 && n < p.nsamp) {
            const float2 r = dft_load(p.iq, p.fmt, (long long)c * p.in_stride + n);
            const float2 a = p.avg[(long long)c * p.navg + (int)(n / p.maxcnt)];
            v = make_float2(r.x - a.x, r.y - a.y);                           /* dft_detect.c:581 */
        }
        z[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < FRONT_TILE + 1; i += FRONT_TILE) {          /* y[n] = sum_k ws[k] z[n-k]  (:709-716) */
        float2 acc = make_float2(0.f, 0.f);
        const float2 *zp = z + taps - 1 + i;                                 /* sample n = t0-1+i */
#pragma unroll 7
        for (int k = 0; k < taps; ++k) { const float w = ws[k]; const float2 v = zp[-k]; acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); }
        y[i] = acc;
    }
    __syncthreads();
    const long long n = t0 + threadIdx.x;
    if (n < p.nsamp) {
        const float g = (float)(FM_GAIN / M_PI);
        const float2 a = y[threadIdx.x + 1], b = y[threadIdx.x];             /* w = z_fm * conj(z_fm_prev)  (:780-781) */
        const float wr = a.x * b.x + a.y * b.y, wi = a.y * b.x - a.x * b.y;
        p.fm_filt[(long long)c * p.nsamp + n] = g * atan2f(wi, wr);
        const float2 a2 = z[taps + threadIdx.x], b2 = z[taps + threadIdx.x - 1];
        const float vr = a2.x * b2.x + a2.y * b2.y, vi = a2.y * b2.x - a2.x * b2.y;
        p.fm_raw[(long long)c * p.nsamp + n] = g * atan2f(vi, vr);            /* ring 3 (:798-800) */
    }
}

/* ---- K2: matched-filter correlator, one CTA per (block, header, channel) ---------------------------------- */
__device__ __forceinline__ int padr(int i) { return i + (i >> 4); }   /* one float of padding per 16: threads 16 apart hit distinct banks */

__device__ __forceinline__ float fm_at(const float *fm, long long n) { return n >= 0 ? __ldg(fm + n) : 0.0f; }

/* acc[r] = sum_t x[o0 + r + t] * coef[t], r < CORR_R, t < ncoef (multiple of CORR_R); x in the padded layout, o0 a multiple of
 * CORR_R.  The CORR_R running samples live in registers and slide: one shared-memory load (+ one broadcast coefficient)
 * per CORR_R FMAs. */
__device__ __forceinline__ void sliding_dot(const float *xp, int o0, const float *coef, int ncoef, float (&acc)[CORR_R])
{
    float xs[CORR_R];
#pragma unroll
    for (int r = 0; r < CORR_R; ++r) { acc[r] = 0.0f; xs[r] = xp[padr(o0) + r]; }
    for (int t0 = 0; t0 < ncoef; t0 += CORR_R) {
        const int nb = padr(o0 + t0 + CORR_R);
#pragma unroll
        for (int u = 0; u < CORR_R; ++u) {
            const float m = coef[t0 + u];
#pragma unroll
            for (int r = 0; r < CORR_R; ++r) acc[r] = fmaf(xs[(u + r) & (CORR_R - 1)], m, acc[r]);
            xs[u] = xp[nb + u];
        }
    }
}

/* read_bufbit (dft_detect.c:821-864): integrate-and-dump bit slicer on the un-low-passed FM stream */
__device__ void slice_bits(const float *fm, long long mvp, float spb, float dc, int symlen, int nsym, unsigned &rcount, float &grenze,
                           unsigned char *out)
{
    for (int b = 0; b < nsym; ++b) {
        double sum = 0.0;
        grenze += spb;
        do { sum += (double)(fm_at(fm, mvp + rcount) - dc); rcount++; } while ((float)rcount < grenze);
        if (symlen == 2) {
            grenze += spb;
            do { sum -= (double)(fm_at(fm, mvp + rcount) - dc); rcount++; } while ((float)rcount < grenze);
        }
        out[b] = sum >= 0 ? 1 : 0;
    }
}

__global__ void __launch_bounds__(CORR_THREADS, 2) dft_corr_kernel(const DftParams p)
{
    extern __shared__ float smf[];
    const int b = blockIdx.x, a = blockIdx.y, z = blockIdx.z, tid = threadIdx.x;
    const int c = p.chan_list ? p.chan_list[z] : z;
    const HdrDev &h = p.hdr[a];
    const int K = p.K, L = h.L, W = K + L, taps = p.lpFMtaps, tpad = p.tpad, H = tpad - 1;   /* H: history points kept before the window */
    const int nx = H + W + 2 * CORR_R;                /* raw points incl. tile overrun */
    float *xw = smf;                                  /* padded: xw[padr(i)] = x'[i - H]  */
    float *xf = xw + padr(nx) + 16;                   /* padded: filtered window + overrun */
    float *mt = xf + padr(W + h.Lpad + 2 * CORR_R) + 16;   /* [Lpad] matched filter      */
    float *wr = mt + h.Lpad;                          /* [tpad] reversed FM low-pass taps, front-padded with zeros */
    __shared__ float red[CORR_THREADS / 32], redv[CORR_THREADS / 32]; __shared__ int redi[CORR_THREADS / 32];
    __shared__ float s_dc, s_best; __shared__ int s_mp;

    const float *fm = (h.use_raw ? p.fm_raw : p.fm_filt) + (long long)c * p.nsamp;
    /* the reference triggers after n = (b+1)(K-4) samples: pos = sample_out = n-1-delay (dft_detect.c:813-816,1498) */
    const long long n_trig = (p.sched_off ? p.sched_off[z] : 0) + (long long)(b + 1) * (K - 4), pos = n_trig - 1 - p.delay, w0 = pos - (W - 1);
    if (n_trig > p.nsamp) {                                                  /* this schedule has fewer blocks than the launch */
        if (tid == 0) { BlkRes r; r.score = 0.0f; r.pos = 0; r.dc = 0.0f; r.mp = -1; r.herrs = 99; r.m10 = 0; p.res[((long long)z * p.nblk + b) * p.nact + a] = r; }
        return;
    }
    for (int i = tid; i < nx; i += CORR_THREADS) xw[padr(i)] = (i >= H && i < H + W) ? fm_at(fm, w0 + i - H) : 0.0f;   /* :376 */
    for (int i = tid; i < h.Lpad; i += CORR_THREADS) mt[i] = p.match[h.match_off + i];
    for (int i = tid; i < tpad; i += CORR_THREADS) { const int k = tpad - 1 - i; wr[i] = k < taps ? p.ws_lpFM[h.lpFM * taps + k] : 0.0f; }
    __syncthreads();
    /* DC over the last 2L samples of the window (:388-391); X[0] -= N*dc*0.98 == subtract 0.98*dc from all N points */
    float dc = 0.0f;
    if (p.opt_dc) {
        float s = 0.0f;
        for (int i = K - L + tid; i < K + L; i += CORR_THREADS) s += xw[padr(H + i)];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((tid & 31) == 0) red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) { float t = 0; for (int w = 0; w < CORR_THREADS / 32; ++w) t += red[w]; s_dc = (float)((double)t / (2.0 * (double)(float)L)); }
        __syncthreads();
        dc = s_dc;
        const float sub = (float)((double)dc * 0.98);
        /* the points before the window are the zero padding of the N-point circular buffer, shifted by -0.98dc like everything else */
        for (int i = tid; i < H + W; i += CORR_THREADS) xw[padr(i)] -= sub;
        __syncthreads();
    }
    /* FM low-pass == circular convolution with the taps (X *= WS, :396): xf[n] = sum_k wl[k] x'[n-k] = sum_t wr[t] xw[n+t] */
    for (int n0 = tid * CORR_R; n0 < W + h.Lpad + CORR_R; n0 += CORR_THREADS * CORR_R) {
        float acc[CORR_R];
        if (n0 < W) sliding_dot(xw, n0, wr, tpad, acc);
#pragma unroll
        for (int r = 0; r < CORR_R; ++r) xf[padr(n0) + r] = (n0 + r < W) ? acc[r] : 0.0f;   /* zeros beyond the window feed tile overruns */
    }
    __syncthreads();
    /* correlation: corr[lag] = sum_t xf[lag+t]*match[t], lag = i-(L-1) in [0, K]  (Z = X*Fm, :403-404) */
    const int nlag = K + 1;
    float best = -1.0f, bestv = 0.0f; int besti = -1;
    for (int lag0 = tid * CORR_R; lag0 < nlag; lag0 += CORR_THREADS * CORR_R) {
        float acc[CORR_R];
        sliding_dot(xf, lag0, mt, h.Lpad, acc);
#pragma unroll
        for (int r = 0; r < CORR_R; ++r) {
            const int lag = lag0 + r;
            if (lag < nlag) { const float v2 = acc[r] * acc[r]; if (v2 > best) { best = v2; bestv = acc[r]; besti = lag; } }
        }
    }
    /* first maximum of corr^2 (strict >, ascending lag, :413-421): larger value wins, ties -> smaller lag.
       The reference starts from mx2 = 0, so an all-zero window yields mp = -1. */
    if (!(best > 0.0f)) { best = 0.0f; besti = 0x7fffffff; }
    for (int o = 16; o; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o), ov = __shfl_xor_sync(0xffffffffu, bestv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; bestv = ov; besti = oi; }
    }
    if ((tid & 31) == 0) { red[tid >> 5] = best; redv[tid >> 5] = bestv; redi[tid >> 5] = besti; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < CORR_THREADS / 32; ++w)
            if (red[w] > best || (red[w] == best && redi[w] < besti)) { best = red[w]; bestv = redv[w]; besti = redi[w]; }
        s_mp = (best > 0.0f) ? besti + (L - 1) : -1; s_best = bestv;
    }
    __syncthreads();
    const int mp = s_mp;
    BlkRes r; r.score = 0.0f; r.pos = 0; r.dc = dc; r.mp = mp; r.herrs = 99; r.m10 = 0;
    if (mp == L - 1 || mp == K + L - 1) r.mp = -4;                            /* Randwert, :422 */
    float nrm = 0.0f;
    if (r.mp > 0) {                                                          /* xnorm over the L filtered samples under the template (:427-429) */
        for (int i = tid; i < L; i += CORR_THREADS) { const float v = xf[padr(mp - i)]; nrm = fmaf(v, v, nrm); }
        for (int o = 16; o; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    }
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = nrm;
    __syncthreads();
    if (tid == 0) {
        if (r.mp > 0) {
            double t = 0; for (int w = 0; w < CORR_THREADS / 32; ++w) t += red[w];
            r.score = (float)((double)s_best / sqrt(t));
            const long long mpos = w0 + mp - taps / 2;                        /* :425,433 */
            r.pos = (unsigned)(mpos + p.pos_base);
            if (r.score > h.thres || r.score < -h.thres) {                    /* candidate: headcmp (:866-901) */
                unsigned char bits[96]; unsigned rc = 0; float gr = 0.0f;
                const float hdc = p.opt_dc ? dc : 0.0f;
                slice_bits(fm, mpos + 1 - (int)(h.hLen * h.spb), h.spb, hdc, 1, h.hLen, rc, gr, bits);
                const int inv = r.score < 0;
                int errs = 0;
                for (int i = 0; i < h.hLen; ++i) errs += ((bits[i] ^ inv) != h.bits[i]);
                r.herrs = errs;
                if (h.j == 6) {                                              /* frm_M10 (:932-965): Manchester bits after the header */
                    const int ofs = (h.hLen - 28) / 2;
                    unsigned char mb; unsigned char bit0 = inv ? 1 : 0; unsigned bytes = 0; unsigned rc2 = 0; float gr2 = 0.0f;
                    for (int q = 0; q < 16; ++q) {
                        if (q < ofs) mb = h.bits[28 + 2 * q] ^ inv;
                        else slice_bits(fm, mpos, h.spb, hdc, 2, 1, rc2, gr2, &mb);
                        const unsigned fb = 1u ^ (bit0 ^ mb);
                        bytes = (bytes << 1) | fb;
                        bit0 = mb;
                    }
                    r.m10 = bytes;
                }
            }
        }
        p.res[((long long)z * p.nblk + b) * p.nact + a] = r;
    }
}

/* ---- K3: iMet AFSK power test (dft_detect.c:1542-1607): |DFT| of successive D-sample segments of the filtered FM stream, summed
 *          over m bins around 2200 / 2400 / 800 Hz.  One CTA per pending event; one warp per bin at a time. ---------------------- */
struct ImetEv { int c; int n; long long src; };
__global__ void dft_imet_kernel(const DftParams p, const ImetEv *ev, int m, int b22, int b24, int b8, float *out)
{
    const ImetEv e = ev[blockIdx.x];
    const float *fm = p.fm_filt + (long long)e.c * p.nsamp + e.src;
    const int N = p.N, D = N / 2 - 3, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __shared__ float pw[3];
    if (threadIdx.x < 3) pw[threadIdx.x] = 0.0f;
    __syncthreads();
    const int bins0[3] = {b22 - m / 4, b24 - m / 4, b8 - m / 4};
    for (int s0 = 0; s0 + D <= e.n; s0 += D) {
        for (int t = warp; t < 3 * m; t += nw) {
            const int g = t / m, bin = bins0[g] + (t - g * m);
            float re = 0.0f, im = 0.0f;
            for (int i = lane; i < D; i += 32) {
                const int k = (int)(((long long)bin * i) & (N - 1));             /* N is a power of two */
                float sn, cs; sincospif(-2.0f * (float)k / (float)N, &sn, &cs);
                const float x = __ldg(fm + s0 + i);
                re = fmaf(x, cs, re); im = fmaf(x, sn, im);
            }
            for (int o = 16; o; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
            if (lane == 0) atomicAdd(&pw[g], sqrtf(re * re + im * im));
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) out[blockIdx.x * 3 + threadIdx.x] = pw[threadIdx.x];
}

/* Per-channel state of the decision loop of main() (dft_detect.c:1498-1652), resumable because the iMet AFSK branch
 * (:1542-1607) swallows one extra second of samples OUTSIDE the block counter and so shifts every later block. */
struct ChanRun {
    float mv[NRS]; unsigned mv_pos[NRS], mv0_pos[NRS]; int mp[NRS]; float dcv[NRS], dfv[NRS];
    const char *type[NRS]; int tn[NRS];
    int j_max; float mv_max; unsigned frm2;
    long long off;            /* samples consumed outside the block counter so far */
    int b;                    /* next block of the current schedule */
    bool done, need;
    bool starved;             /* streaming: the next step needs samples that have not been pushed yet */
    bool wait_imet; long long imet_ntrig; int imet_n, imet_hf;   /* an iMet candidate waits for its power test */
    b200dft_result_t *out;
};

/* =============================== host side =================================================================== */
struct b200dft {
    int sr, opt_dc, nchan, device; double bw_khz;
    int N, K, L, M, delay, lpIQtaps, lpFMtaps, nact, maxcnt;
    std::vector<HdrDev> hdr; std::vector<int> act;      /* active header rows */
    float spb[NRS]; int hLen[NRS], Lj[NRS];
    float thres[NRS];              /* detection thresholds: rs_hdr[].thres, or the --ths override for every header (:1426-1431) */
    float *d_ws_lpIQ = 0, *d_ws_lpFM = 0, *d_match = 0; HdrDev *d_hdr = 0;
    float2 *d_avg = 0; float *d_fm_filt = 0, *d_fm_raw = 0; BlkRes *d_res = 0; void *d_iq = 0;
    long long cap_samp = 0, cap_samp_run = 0, cap_iq_bytes = 0; int cap_blk = 0, cap_avg = 0;
    int *d_list = 0; long long *d_soff = 0;
    size_t corr_smem_max = 0;
    cudaStream_t stream = 0; cudaEvent_t ev[4] = {0, 0, 0, 0};
    long launches = 0; bool timed = false;
    /* wideband (--IQ) mode */
    bool wide = false; int sr_base = 0, decM = 1, dectaps = 0, wmaxcnt = 0, ch_tm = 0; size_t ch_smem = 0;
    float *d_ws_dec = 0; double *d_f0 = 0; float2 *d_wavg = 0; float2 *d_if = 0; void *d_wide = 0;
    long long cap_wavg = 0, cap_if = 0, cap_wide_bytes = 0;
    cudaEvent_t evc[2] = {0, 0}; bool ch_timed = false;
    /* streaming (b200dft_stream_*): the tail of the raw input that later blocks still need, and the decision state of every channel */
    bool st_active = false; int st_fmt = 0, st_cont = 0; double st_tl = -1.0;
    long long st_abs0 = 0, st_have = 0, st_total = 0, st_cap = 0;
    void *d_sbuf[2] = {0, 0}; int st_cur = 0;
    std::vector<ChanRun> st_run; b200dft_result_t *st_res = 0;
};

static double sinc_(double x) { return x == 0 ? 1 : sin(M_PI * x) / (M_PI * x); }
static int lowpass_init(float f, int taps, std::vector<float> &ws)     /* dft_detect.c:665-697 */
{
    if (taps % 2 == 0) taps++;
    if (taps < 1) taps = 1;
    ws.assign(taps, 0.0f);
    double norm = 0;
    for (int n = 0; n < taps; n++) {
        double w = 7938 / 18608.0 - 9240 / 18608.0 * cos(2 * M_PI * n / (taps - 1)) + 1430 / 18608.0 * cos(4 * M_PI * n / (taps - 1));
        double hh = 2 * f * sinc_(2 * f * (n - (taps - 1) / 2));
        ws[n] = w * hh;
        norm += ws[n];
    }
    for (int n = 0; n < taps; n++) ws[n] /= norm;
    return taps;
}
static double Qf(double x) { return 0.5 - 0.5 * erf(x / 1.4142135624); }
static double pulse(double t, double sigma) { return Qf((t - 0.5) / sigma) - Qf((t + 0.5) / sigma); }

extern "C" int b200dft_create(b200dft_t **out, int sample_rate, double bw_khz, int opt_dc, int nchan, int device)
{
    if (!out || sample_rate < 1 || nchan <= 0) return fail(B200DFT_EINVAL, "bad arguments");
    if (!(bw_khz * 1e3 > 100.0)) return fail(B200DFT_EINVAL, "--bw <kHz> is required (single IF filter mode, dft_detect.c:1132-1137)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(B200DFT_ECUDA, "no CUDA device: libb200dft has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(B200DFT_ERANGE, "device");
    CK(cudaSetDevice(device));
    b200dft *h = new b200dft();
    for (int j = 0; j < NRS; j++) h->thres[j] = HDR[j].thres;
    h->sr = sample_rate; h->opt_dc = opt_dc ? 1 : 0; h->nchan = nchan; h->device = device; h->bw_khz = bw_khz;
    const int sr = sample_rate;
    /* init_buffers(), --iq branch (dft_detect.c:1113-1156) */
    std::vector<float> wfm0, wfm1, wiq;
    int taps = 4 * sr / 2e3; if (taps % 2 == 0) taps++;
    taps = lowpass_init(4e3f / (float)sr, taps, wfm0);
    taps = lowpass_init(10e3f / (float)sr, taps, wfm1);
    h->lpFMtaps = taps;
    taps = 4 * sr / 4e3; if (taps % 2 == 0) taps++;
    const float set_lpIQ = bw_khz * 1e3;
    taps = lowpass_init(set_lpIQ / (float)sr / 2.0, taps, wiq);
    h->lpIQtaps = taps;
    if (taps > 256) { delete h; return fail(B200DFT_EINVAL, "IF filter too long"); }
    h->maxcnt = sr / 32;
    int hLenMax = 0, Lmax = 0;
    for (int j = 0; j < NRS; j++) {                                          /* :1175-1184 */
        h->spb[j] = sr / (float)HDR[j].sps;
        h->hLen[j] = strlen(HDR[j].header);
        h->Lj[j] = h->hLen[j] * h->spb[j] + 0.5;
        if (j != SKIP_A && j != SKIP_B) { if (h->hLen[j] > hLenMax) hLenMax = h->hLen[j]; if (h->Lj[j] > Lmax) Lmax = h->Lj[j]; }
    }
    int L = 2 * Lmax, M = 3 * L, p2 = 1;
    while (p2 < M) p2 <<= 1;
    while (p2 < 0x2000) p2 <<= 1;
    h->N = p2; h->K = p2 - L; h->L = L; h->delay = L / 16; h->M = p2 + h->delay + 8;
    /* matched filters (:1223-1259), time domain only: the reference's m[L-1-i] = match[i] reversal + FFT product
       is the correlation sum_t x[lag+t]*match[t] evaluated directly by the kernel */
    std::vector<float> match_all; std::vector<float> match(L + 1);
    for (int j = 0; j <= IDX_IMETAFSK; j++) {
        if (j == SKIP_A || j == SKIP_B) continue;
        const char *bits = HDR[j].header; const float spb = h->spb[j]; const int Lj = h->Lj[j];
        if (h->K + Lj > h->N) continue;
        const double sigma = sqrt(log(2)) / (2 * M_PI * HDR[j].BT);
        for (int i = 0; i < Lj; i++) {
            int pos = i / spb;
            double t = (i - pos * spb) / spb - 0.5;
            double bb = ((bits[pos] & 0x1) - 0.5) * 2.0 * pulse(t, sigma);
            if (pos > 0) bb += ((bits[pos - 1] & 0x1) - 0.5) * 2.0 * pulse(t + 1, sigma);
            if (pos < hLenMax - 1) bb += ((bits[pos + 1] & 0x1) - 0.5) * 2.0 * pulse(t - 1, sigma);   /* global hLen: :1249 */
            match[i] = bb;
        }
        double n2 = 0; for (int i = 0; i < Lj; i++) n2 += (double)match[i] * match[i];
        const float nm = sqrt(n2);
        for (int i = 0; i < Lj; i++) match[i] /= nm;
        HdrDev hd; memset(&hd, 0, sizeof hd);
        hd.j = j; hd.L = Lj; hd.Lpad = (Lj + CORR_R - 1) / CORR_R * CORR_R; hd.hLen = h->hLen[j]; hd.use_raw = HDR[j].lpIQ == 3;
        hd.lpFM = HDR[j].lpFM; hd.match_off = (int)match_all.size(); hd.spb = spb; hd.thres = HDR[j].thres; h->thres[j] = HDR[j].thres;
        for (int i = 0; i < hd.hLen; i++) hd.bits[i] = bits[i] & 1;
        for (int i = 0; i < hd.Lpad; i++) match_all.push_back(i < Lj ? match[i] : 0.0f);
        h->hdr.push_back(hd); h->act.push_back(j);
        const int tpad = (h->lpFMtaps + CORR_R - 1) / CORR_R * CORR_R, W = h->K + Lj, nx = tpad - 1 + W + 2 * CORR_R, nf = W + hd.Lpad + 2 * CORR_R;
        size_t sm = ((size_t)(nx + (nx >> 4) + 16) + (size_t)(nf + (nf >> 4) + 16) + hd.Lpad + tpad + 64) * 4;
        if (sm > h->corr_smem_max) h->corr_smem_max = sm;
    }
    h->nact = (int)h->hdr.size();
    int maxsmem = 0;
    CK(cudaDeviceGetAttribute(&maxsmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (h->corr_smem_max + 1024 > (size_t)maxsmem) { delete h; return fail(B200DFT_EINVAL, "correlator window does not fit shared memory"); }
    {   /* per-function attribute: allow the device maximum (minus the kernel's static shared memory) */
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, (const void *)dft_corr_kernel));
        CK(cudaFuncSetAttribute(dft_corr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsmem - (int)fa.sharedSizeBytes));
    }
    std::vector<float> wfm(2 * h->lpFMtaps);
    for (int i = 0; i < h->lpFMtaps; i++) { wfm[i] = wfm0[i]; wfm[h->lpFMtaps + i] = wfm1[i]; }
    CK(cudaMalloc(&h->d_ws_lpIQ, wiq.size() * 4)); CK(cudaMemcpy(h->d_ws_lpIQ, wiq.data(), wiq.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_ws_lpFM, wfm.size() * 4)); CK(cudaMemcpy(h->d_ws_lpFM, wfm.data(), wfm.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_match, match_all.size() * 4)); CK(cudaMemcpy(h->d_match, match_all.data(), match_all.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_hdr, h->hdr.size() * sizeof(HdrDev))); CK(cudaMemcpy(h->d_hdr, h->hdr.data(), h->hdr.size() * sizeof(HdrDev), cudaMemcpyHostToDevice));
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++) CK(cudaEventCreate(&h->ev[i]));
    *out = h;
    return B200DFT_OK;
}

extern "C" int b200dft_set_threshold(b200dft_t *h, double ths)
{
    if (!h) return fail(B200DFT_EINVAL, "null");
    CK(cudaSetDevice(h->device));
    for (int j = 0; j < NRS; j++) h->thres[j] = (float)ths;               /* --ths x: one threshold for every header (:1426-1431) */
    for (size_t a = 0; a < h->hdr.size(); a++) h->hdr[a].thres = (float)ths;
    CK(cudaMemcpy(h->d_hdr, h->hdr.data(), h->hdr.size() * sizeof(HdrDev), cudaMemcpyHostToDevice));
    return B200DFT_OK;
}

extern "C" void b200dft_destroy(b200dft_t *h)
{
    if (!h) return;
    cudaSetDevice(h->device); cudaDeviceSynchronize();
    void *ptrs[] = {h->d_ws_lpIQ, h->d_ws_lpFM, h->d_match, h->d_hdr, h->d_avg, h->d_fm_filt, h->d_fm_raw, h->d_res, h->d_iq, h->d_list, h->d_soff, h->d_sbuf[0], h->d_sbuf[1],
                    h->d_ws_dec, h->d_f0, h->d_wavg, h->d_if, h->d_wide};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < 4; i++) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 2; i++) if (h->evc[i]) cudaEventDestroy(h->evc[i]);
    cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int b200dft_info(const b200dft_t *h, b200dft_info_t *info)
{
    if (!h || !info) return fail(B200DFT_EINVAL, "null");
    info->sample_rate = h->sr; info->N_DFT = h->N; info->K = h->K; info->L = h->L; info->M = h->M; info->delay = h->delay;
    info->lpIQtaps = h->lpIQtaps; info->lpFMtaps = h->lpFMtaps; info->nhdr_active = h->nact; info->nchan = h->nchan;
    return B200DFT_OK;
}

static void chan_init(ChanRun &s, b200dft_result_t *out)
{
    memset(&s, 0, sizeof s);
    for (int j = 0; j < NRS; j++) { s.type[j] = HDR[j].type; s.tn[j] = HDR[j].tn; }
    s.out = out; memset(out, 0, sizeof *out);
}
static void chan_finish(ChanRun &s)
{
    int hf; if (s.mv_max) hf = (s.mv_max < 0 && s.j_max < 3) ? -1 : 1; else hf = 0;
    s.out->exit_code = hf * s.tn[s.j_max]; s.done = true; s.need = false;
}
static void chan_report(const b200dft *h, ChanRun &s, int j)                 /* the printf block, :1611-1642 */
{
    if (s.mv[j] > h->thres[j] || s.mv[j] < -h->thres[j]) {
        if (s.out->ndet < B200DFT_MAXDET) {
            b200dft_det_t *d = &s.out->dets[s.out->ndet];
            d->hdr = j; d->tn = s.tn[j]; d->score = s.mv[j]; d->pos = s.mv_pos[j]; d->df_hz = s.dfv[j] * (h->wide ? h->sr_base : h->sr); d->m10 = s.frm2 & 0xFFFF;
            memset(d->type, 0, sizeof d->type); strncpy(d->type, s.type[j], sizeof d->type - 1);
        }
        s.out->ndet++;
        if (j == 6) s.frm2 = 0;
    }
    if (fabsf(s.mv_max) < fabsf(s.mv[j])) { s.mv_max = s.mv[j]; s.j_max = j; }
}

/* Closing part of one block iteration (dft_detect.c:1649-1651) */
static void chan_end_block(b200dft *h, ChanRun &s, int header_found, bool imet_event, int opt_cont)
{
    if (header_found && !opt_cont) { chan_finish(s); return; }
    for (int j = 0; j < NRS; j++) s.mv[j] = 0.0f;
    if (imet_event) { s.off += (long long)s.b * (h->K - 4); s.b = 0; s.need = true; }   /* later blocks follow a shifted schedule */
}

/* The iMet branch once the power sums are known (dft_detect.c:1570-1607). verdict: 1 relabel to IMET4/IMET1RS, 0 drop, -1 goto ende */
static void chan_resume_imet(b200dft *h, ChanRun &s, int verdict, int opt_cont)
{
    const int j0 = IDX_IMETAFSK;
    int header_found = s.imet_hf;
    s.wait_imet = false;
    s.off += s.imet_n;
    if (verdict < 0) { chan_finish(s); return; }
    s.mv[j0] = fabsf(s.mv[j0]);
    if (verdict == 1) {
        const int j = (h->bw_khz * 1e3 > 50e3) ? IDX_RS : IDX_I4;
        s.mv[j] = s.mv[j0]; s.mv_pos[j] = s.mv_pos[j0]; s.dcv[j] = s.dcv[j0]; s.dfv[j] = s.dfv[j0]; s.mv[j0] = 0.0f;
        header_found = 1;
        chan_report(h, s, j);
    } else s.mv[j0] = 0.0f;
    chan_end_block(h, s, header_found, true, opt_cont);
}

/* Advance one channel over the results of its current schedule; stops at the end of the stream, at the first detection
 * (no -c), when the schedule's results run out, or at an iMet candidate (s.wait_imet: the caller runs the power test). */
static void chan_advance(b200dft *h, ChanRun &s, const BlkRes *res, int nblk_launch, long long neff, int opt_cont, bool final = true)
{
    const int nact = h->nact, Kb = h->K - 4;
    s.need = false;
    while (!s.done && !s.need && !s.wait_imet) {
        const long long n_trig = s.off + (long long)(s.b + 1) * Kb;
        if (n_trig > neff) { if (final) chan_finish(s); else s.starved = true; return; }
        if (s.b >= nblk_launch) { s.need = true; return; }
        const BlkRes *rb = res + (size_t)s.b * nact;
        if (!final && neff - n_trig < h->sr) {
            /* streaming: an iMet AFSK candidate in this block would read one more second of samples (:1542-1607) that has not arrived
               yet -- look before touching any state, and wait for the next push if so */
            for (int a = 0; a < nact; a++) {
                const int j = h->act[a];
                if (j != IDX_IMETAFSK) continue;
                const float mvj = rb[a].mp > 0 ? rb[a].score : s.mv[j];
                const unsigned pj = rb[a].mp > 0 ? rb[a].pos : s.mv_pos[j];
                if (rb[a].mp > 0 && (mvj > h->thres[j] || mvj < -h->thres[j]) && pj > s.mv_pos[j] && rb[a].herrs < HDR[j].herrs) { s.starved = true; return; }
            }
        }
        for (int a = 0; a < nact; a++) {
            const int j = h->act[a];
            s.mv0_pos[j] = s.mv_pos[j];
            s.mp[j] = rb[a].mp; s.dcv[j] = rb[a].dc;
            if (rb[a].mp > 0) { s.mv[j] = rb[a].score; s.mv_pos[j] = rb[a].pos; }   /* getCorrDFT leaves them untouched on failure */
            if (h->opt_dc) s.dfv[j] = s.dcv[j] / (2.0 * FM_GAIN * h->decM);            /* :439 */
        }
        s.b++;
        int header_found = 0;
        for (int a = 0; a < nact; a++) {
            const int j = h->act[a];
            if (!(s.mp[j] > 0 && (s.mv[j] > h->thres[j] || s.mv[j] < -h->thres[j]))) continue;
            if (!(s.mv_pos[j] > s.mv0_pos[j])) continue;
            if (!(rb[a].herrs < HDR[j].herrs)) continue;
            if (j == 6) {                                                       /* M10 / M20 (:1526-1540) */
                const unsigned bytes = rb[a].m10; int hwt = 0; for (int q = 0; q < 4; q++) hwt += (bytes >> q) & 1;
                if (hwt < 2 || (hwt == 2 && (bytes & 0xF0) == 0x20)) { s.type[j] = "M20"; s.tn[j] = 6; } else { s.type[j] = "M10"; s.tn[j] = 5; }
                s.frm2 = bytes;
            }
            if (j == IDX_IMETAFSK) {                                            /* always the last active header: park the channel */
                const long long avail = neff - n_trig;
                s.imet_n = (int)(avail < h->sr ? (avail > 0 ? avail : 0) : h->sr);   /* one more second is read outside the block counter */
                s.imet_ntrig = n_trig; s.imet_hf = header_found; s.wait_imet = true;
                return;
            }
            header_found = 1;
            chan_report(h, s, j);
        }
        chan_end_block(h, s, header_found, false, opt_cont);
    }
}

static int run_common(b200dft *h, const void *d_iq, int fmt, long in_stride, long nsamp, int opt_cont, double tl, b200dft_result_t *res,
                      bool skip_iqdc = false)
{
    /* -t: blocks only trigger while sample_in <= (tl+1)*sr (:1494) */
    long neff = nsamp;
    if (tl > 0) { const long lim = (long)floor((tl + 1) * h->sr); if (neff > lim) neff = lim; }
    const int nblk = (int)(neff / (h->K - 4));
    const int navg = (int)(neff / h->maxcnt) + 2;
    if (neff > h->cap_samp) {
        if (h->d_fm_filt) cudaFree(h->d_fm_filt); if (h->d_fm_raw) cudaFree(h->d_fm_raw);
        CK(cudaMalloc(&h->d_fm_filt, (size_t)h->nchan * neff * 4)); CK(cudaMalloc(&h->d_fm_raw, (size_t)h->nchan * neff * 4));
        h->cap_samp = neff;
    }
    h->cap_samp_run = neff;
    if (navg > h->cap_avg) { if (h->d_avg) cudaFree(h->d_avg); CK(cudaMalloc(&h->d_avg, (size_t)h->nchan * navg * sizeof(float2))); h->cap_avg = navg; }
    if (nblk > h->cap_blk) {
        if (h->d_res) cudaFree(h->d_res); CK(cudaMalloc(&h->d_res, (size_t)h->nchan * nblk * h->nact * sizeof(BlkRes))); h->cap_blk = nblk;
    }
    if (!h->d_list) { CK(cudaMalloc(&h->d_list, sizeof(int) * h->nchan)); CK(cudaMalloc(&h->d_soff, sizeof(long long) * h->nchan)); }
    DftParams p; memset(&p, 0, sizeof p);
    p.nchan = h->nchan; p.nact = h->nact; p.N = h->N; p.K = h->K; p.L = h->L; p.lpIQtaps = h->lpIQtaps; p.lpFMtaps = h->lpFMtaps; p.tpad = (h->lpFMtaps + CORR_R - 1) / CORR_R * CORR_R;
    p.delay = h->delay; p.opt_dc = h->opt_dc; p.maxcnt = h->maxcnt; p.nblk = nblk; p.navg = navg; p.nsamp = neff; p.in_stride = in_stride;
    p.fmt = fmt; p.iq = d_iq; p.ws_lpIQ = h->d_ws_lpIQ; p.ws_lpFM = h->d_ws_lpFM; p.match = h->d_match; p.hdr = h->d_hdr;
    p.avg = h->d_avg; p.fm_filt = h->d_fm_filt; p.fm_raw = h->d_fm_raw; p.res = h->d_res;
    cudaStream_t s = h->stream;
    CK(cudaEventRecord(h->ev[0], s));
    if (neff > 0) {
        dim3 g0((unsigned)(navg - 1), h->nchan);
        if (skip_iqdc) CK(cudaMemsetAsync(h->d_avg, 0, (size_t)h->nchan * navg * sizeof(float2), s));   /* --IQ: the DC block ran on the wide stream */
        else dft_dcavg_kernel<<<g0, 128, 0, s>>>(p);
        dim3 g1((unsigned)((neff + FRONT_TILE - 1) / FRONT_TILE), h->nchan);
        const size_t sm1 = (size_t)(2 * FRONT_TILE + h->lpIQtaps + 1) * sizeof(float2);
        dft_front_kernel<<<g1, FRONT_TILE, sm1, s>>>(p);
        h->launches += 2;
    }
    CK(cudaEventRecord(h->ev[1], s));
    if (nblk > 0) {
        dim3 g2(nblk, h->nact, h->nchan);
        dft_corr_kernel<<<g2, CORR_THREADS, h->corr_smem_max, s>>>(p);
        h->launches += 1;
    }
    CK(cudaEventRecord(h->ev[2], s));
    CK(cudaGetLastError());
    std::vector<BlkRes> hres((size_t)h->nchan * nblk * h->nact);
    if (nblk > 0) CK(cudaMemcpyAsync(hres.data(), h->d_res, hres.size() * sizeof(BlkRes), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    h->timed = true;
    std::vector<ChanRun> st(h->nchan);
    for (int c = 0; c < h->nchan; c++) chan_init(st[c], &res[c]);
    std::vector<int> todo(h->nchan);                 /* channels that still have a schedule to walk */
    for (int c = 0; c < h->nchan; c++) todo[c] = c;
    std::vector<BlkRes> r2;
    const BlkRes *rbase = hres.data(); int nb_cur = nblk; bool identity = true;
    const float dfb = (1 / (float)h->N) * h->sr;
    int m_imet = 50.0 / dfb; if (m_imet < 1) m_imet = 1;
    const bool imet_ok = !(2500 * h->N / (float)h->sr > h->N / 2);
    while (!todo.empty()) {
        const int nz = (int)todo.size();
        std::vector<int> pend;
        for (int z = 0; z < nz; z++) {
            const int c = todo[z];
            chan_advance(h, st[c], rbase + (size_t)(identity ? c : z) * nb_cur * h->nact, nb_cur, neff, opt_cont);
            if (st[c].wait_imet) pend.push_back(c);
        }
        if (!pend.empty()) {                                                   /* iMet candidates: one small kernel decides them all */
            const int ne = (int)pend.size();
            std::vector<ImetEv> ev(ne); std::vector<float> pw(3 * ne);
            for (int i = 0; i < ne; i++) { ev[i].c = pend[i]; ev[i].n = st[pend[i]].imet_n; ev[i].src = st[pend[i]].imet_ntrig - h->delay; }
            ImetEv *d_ev = 0; float *d_pw = 0;
            CK(cudaMalloc(&d_ev, sizeof(ImetEv) * ne)); CK(cudaMalloc(&d_pw, sizeof(float) * 3 * ne));
            CK(cudaMemcpy(d_ev, ev.data(), sizeof(ImetEv) * ne, cudaMemcpyHostToDevice));
            dft_imet_kernel<<<ne, 256, 0, s>>>(p, d_ev, m_imet, (int)(2200 * h->N / (float)h->sr), (int)(2400 * h->N / (float)h->sr), (int)(800 * h->N / (float)h->sr), d_pw);
            h->launches += 1;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(pw.data(), d_pw, sizeof(float) * 3 * ne, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            cudaFree(d_ev); cudaFree(d_pw);
            for (int i = 0; i < ne; i++) {
                const int verdict = !imet_ok ? -1 : ((pw[3 * i] > pw[3 * i + 1] && pw[3 * i] > pw[3 * i + 2]) ? 1 : 0);
                chan_resume_imet(h, st[pend[i]], verdict, opt_cont);
            }
        }
        /* channels whose block schedule was shifted by an iMet event: correlate again from the new offset (rare, small) */
        std::vector<int> next;
        for (int z = 0; z < nz; z++) if (!st[todo[z]].done && st[todo[z]].need) next.push_back(todo[z]);
        todo.swap(next);
        if (todo.empty()) break;
        const int nz2 = (int)todo.size();
        std::vector<long long> off(nz2); int nb2 = 0;
        for (int z = 0; z < nz2; z++) { off[z] = st[todo[z]].off; const int nb = (int)((neff - off[z]) / (h->K - 4)); if (nb > nb2) nb2 = nb; }
        if (nb2 <= 0) { for (int z = 0; z < nz2; z++) chan_finish(st[todo[z]]); break; }
        if ((long long)nz2 * nb2 > (long long)h->nchan * h->cap_blk) return fail(B200DFT_ERANGE, "relaunch buffer");
        CK(cudaMemcpy(h->d_list, todo.data(), sizeof(int) * nz2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_soff, off.data(), sizeof(long long) * nz2, cudaMemcpyHostToDevice));
        DftParams q = p; q.nblk = nb2; q.chan_list = h->d_list; q.sched_off = h->d_soff;
        dim3 g2(nb2, h->nact, nz2);
        dft_corr_kernel<<<g2, CORR_THREADS, h->corr_smem_max, s>>>(q);
        h->launches += 1;
        CK(cudaGetLastError());
        r2.resize((size_t)nz2 * nb2 * h->nact);
        CK(cudaMemcpyAsync(r2.data(), h->d_res, r2.size() * sizeof(BlkRes), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        rbase = r2.data(); nb_cur = nb2; identity = false;
    }
    return B200DFT_OK;
}

/* ---- streaming form (ADVICE r01 / VERDICT r01 #4): the handle keeps what the reference process keeps between reads ---------------------
 * The reference is a sample loop with all state in statics: DC-block sums, FIR ring, previous IF sample, FM rings, block counter,
 * mv/mv_pos per header.  Every one of them depends on a BOUNDED past: the FM value of sample n needs the raw samples of the previous
 * DC block and lpIQtaps more, a correlation block needs K+L+delay+lpFMtaps FM samples.  So the handle keeps the raw tail that the next
 * undecided block of any channel can still reach (rounded down to a DC-block boundary minus one block, so the kernels recompute exactly
 * the reference's values for everything they are asked about) plus the per-channel decision state (ChanRun), and every push runs the
 * same kernels as the whole-buffer call over [tail + new samples] with absolute positions.  Work per push is O(new samples + K+L). */
extern "C" int b200dft_stream_begin(b200dft_t *h, int fmt, int opt_cont, double tl)
{
    if (!h) return fail(B200DFT_EINVAL, "null");
    if (h->wide) return fail(B200DFT_EINVAL, "streaming is implemented for the --iq mode");
    if (fmt < 0 || fmt > 2) return fail(B200DFT_ERANGE, "fmt");
    h->st_active = true; h->st_fmt = fmt; h->st_cont = opt_cont ? 1 : 0; h->st_tl = tl;
    h->st_abs0 = h->st_have = h->st_total = 0; h->st_res = 0;
    h->st_run.assign(h->nchan, ChanRun());
    return B200DFT_OK;
}

static int stream_step(b200dft *h, bool final)
{
    const int Kb = h->K - 4, nact = h->nact;
    const long long neff = h->st_have, neff_abs = h->st_abs0 + neff;
    std::vector<int> todo;
    for (int c = 0; c < h->nchan; c++) {
        ChanRun &s = h->st_run[c];
        if (s.done) continue;
        s.off += (long long)s.b * Kb; s.b = 0; s.starved = false; s.need = false;      /* the schedule restarts at the next undecided block */
        todo.push_back(c);
    }
    if (todo.empty()) return B200DFT_OK;
    int nbmax = 0;
    for (int c : todo) { const long long nb = (neff_abs - h->st_run[c].off) / Kb; if (nb > nbmax) nbmax = (int)nb; }
    if (nbmax <= 0) { if (final) for (int c : todo) chan_finish(h->st_run[c]); return B200DFT_OK; }
    const int navg = (int)(neff / h->maxcnt) + 2;
    if (neff > h->cap_samp) {
        if (h->d_fm_filt) cudaFree(h->d_fm_filt); if (h->d_fm_raw) cudaFree(h->d_fm_raw);
        const long long cap = neff + neff / 2;
        CK(cudaMalloc(&h->d_fm_filt, (size_t)h->nchan * cap * 4)); CK(cudaMalloc(&h->d_fm_raw, (size_t)h->nchan * cap * 4));
        h->cap_samp = cap;
    }
    h->cap_samp_run = neff;
    if (navg > h->cap_avg) { if (h->d_avg) cudaFree(h->d_avg); CK(cudaMalloc(&h->d_avg, (size_t)h->nchan * navg * sizeof(float2))); h->cap_avg = navg; }
    if (nbmax > h->cap_blk) { if (h->d_res) cudaFree(h->d_res); CK(cudaMalloc(&h->d_res, (size_t)h->nchan * nbmax * nact * sizeof(BlkRes))); h->cap_blk = nbmax; }
    if (!h->d_list) { CK(cudaMalloc(&h->d_list, sizeof(int) * h->nchan)); CK(cudaMalloc(&h->d_soff, sizeof(long long) * h->nchan)); }
    DftParams p; memset(&p, 0, sizeof p);
    p.nchan = h->nchan; p.nact = nact; p.N = h->N; p.K = h->K; p.L = h->L; p.lpIQtaps = h->lpIQtaps; p.lpFMtaps = h->lpFMtaps; p.tpad = (h->lpFMtaps + CORR_R - 1) / CORR_R * CORR_R;
    p.delay = h->delay; p.opt_dc = h->opt_dc; p.maxcnt = h->maxcnt; p.nblk = nbmax; p.navg = navg; p.nsamp = neff; p.in_stride = h->st_cap;
    p.fmt = h->st_fmt; p.iq = h->d_sbuf[h->st_cur]; p.ws_lpIQ = h->d_ws_lpIQ; p.ws_lpFM = h->d_ws_lpFM; p.match = h->d_match; p.hdr = h->d_hdr;
    p.avg = h->d_avg; p.fm_filt = h->d_fm_filt; p.fm_raw = h->d_fm_raw; p.res = h->d_res; p.pos_base = h->st_abs0;
    cudaStream_t s = h->stream;
    CK(cudaEventRecord(h->ev[0], s));
    {
        dim3 g0((unsigned)(navg - 1), h->nchan);
        dft_dcavg_kernel<<<g0, 128, 0, s>>>(p);
        dim3 g1((unsigned)((neff + FRONT_TILE - 1) / FRONT_TILE), h->nchan);
        dft_front_kernel<<<g1, FRONT_TILE, (size_t)(2 * FRONT_TILE + h->lpIQtaps + 1) * sizeof(float2), s>>>(p);
        h->launches += 2;
    }
    CK(cudaEventRecord(h->ev[1], s));
    const float dfb = (1 / (float)h->N) * h->sr;
    int m_imet = 50.0 / dfb; if (m_imet < 1) m_imet = 1;
    const bool imet_ok = !(2500 * h->N / (float)h->sr > h->N / 2);
    std::vector<BlkRes> r2;
    while (!todo.empty()) {
        const int nz = (int)todo.size();
        std::vector<long long> off(nz); int nb2 = 0;
        for (int z = 0; z < nz; z++) {
            const ChanRun &st = h->st_run[todo[z]];
            off[z] = st.off - h->st_abs0;
            const int nb = (int)((neff_abs - st.off) / Kb); if (nb > nb2) nb2 = nb;
        }
        if (nb2 > 0) {
            if ((long long)nz * nb2 > (long long)h->nchan * h->cap_blk) return fail(B200DFT_ERANGE, "relaunch buffer");
            CK(cudaMemcpyAsync(h->d_list, todo.data(), sizeof(int) * nz, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(h->d_soff, off.data(), sizeof(long long) * nz, cudaMemcpyHostToDevice, s));
            DftParams q = p; q.nblk = nb2; q.chan_list = h->d_list; q.sched_off = h->d_soff;
            dim3 g2(nb2, nact, nz);
            dft_corr_kernel<<<g2, CORR_THREADS, h->corr_smem_max, s>>>(q);
            h->launches += 1;
            CK(cudaGetLastError());
            r2.resize((size_t)nz * nb2 * nact);
            CK(cudaMemcpyAsync(r2.data(), h->d_res, r2.size() * sizeof(BlkRes), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
        }
        std::vector<int> pend;
        for (int z = 0; z < nz; z++) {
            ChanRun &st = h->st_run[todo[z]];
            chan_advance(h, st, r2.data() + (size_t)z * nb2 * nact, nb2, neff_abs, h->st_cont, final);
            if (st.wait_imet) pend.push_back(todo[z]);
        }
        if (!pend.empty()) {
            const int ne = (int)pend.size();
            std::vector<ImetEv> ev(ne); std::vector<float> pw(3 * ne);
            for (int i = 0; i < ne; i++) { const ChanRun &st = h->st_run[pend[i]]; ev[i].c = pend[i]; ev[i].n = st.imet_n; ev[i].src = st.imet_ntrig - h->delay - h->st_abs0; }
            ImetEv *d_ev = 0; float *d_pw = 0;
            CK(cudaMalloc(&d_ev, sizeof(ImetEv) * ne)); CK(cudaMalloc(&d_pw, sizeof(float) * 3 * ne));
            CK(cudaMemcpy(d_ev, ev.data(), sizeof(ImetEv) * ne, cudaMemcpyHostToDevice));
            dft_imet_kernel<<<ne, 256, 0, s>>>(p, d_ev, m_imet, (int)(2200 * h->N / (float)h->sr), (int)(2400 * h->N / (float)h->sr), (int)(800 * h->N / (float)h->sr), d_pw);
            h->launches += 1;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(pw.data(), d_pw, sizeof(float) * 3 * ne, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            cudaFree(d_ev); cudaFree(d_pw);
            for (int i = 0; i < ne; i++) {
                const int verdict = !imet_ok ? -1 : ((pw[3 * i] > pw[3 * i + 1] && pw[3 * i] > pw[3 * i + 2]) ? 1 : 0);
                chan_resume_imet(h, h->st_run[pend[i]], verdict, h->st_cont);
            }
        }
        std::vector<int> next;
        for (int z = 0; z < nz; z++) { const ChanRun &st = h->st_run[todo[z]]; if (!st.done && st.need && !st.starved) next.push_back(todo[z]); }
        todo.swap(next);
    }
    CK(cudaEventRecord(h->ev[2], s));
    h->timed = true;
    return B200DFT_OK;
}

extern "C" int b200dft_stream_push_host(b200dft_t *h, const void *h_iq, long in_stride, long nsamp, int flush, b200dft_result_t *res, int *done)
{
    if (!h || !res || (nsamp > 0 && !h_iq)) return fail(B200DFT_EINVAL, "null");
    if (!h->st_active) return fail(B200DFT_EINVAL, "b200dft_stream_begin first");
    if (nsamp < 0) return fail(B200DFT_ERANGE, "nsamp");
    CK(cudaSetDevice(h->device));
    if (!h->st_res) { h->st_res = res; for (int c = 0; c < h->nchan; c++) chan_init(h->st_run[c], &res[c]); }
    else if (h->st_res != res) return fail(B200DFT_EINVAL, "the result array must stay the same for the whole stream");
    const int es = h->st_fmt == 1 ? 8 : (h->st_fmt == 0 ? 4 : 2);
    long long lim = -1;                                  /* -t: the reference stops reading at (tl+1)*sr samples (:1494) */
    if (h->st_tl > 0) lim = (long long)floor((h->st_tl + 1) * h->sr);
    long long acc = nsamp;
    if (lim >= 0 && h->st_total + acc > lim) acc = lim > h->st_total ? lim - h->st_total : 0;
    const bool final = flush || (lim >= 0 && h->st_total + acc >= lim);
    if (h->st_have + acc > h->st_cap) {                  /* grow both buffers, keep the tail */
        const long long ncap = (h->st_have + acc) * 2 + 4096;
        void *nb[2] = {0, 0};
        CK(cudaMalloc(&nb[0], (size_t)h->nchan * ncap * es)); CK(cudaMalloc(&nb[1], (size_t)h->nchan * ncap * es));
        CK(cudaStreamSynchronize(h->stream));
        if (h->st_have > 0)
            CK(cudaMemcpy2D(nb[0], (size_t)ncap * es, h->d_sbuf[h->st_cur], (size_t)h->st_cap * es, (size_t)h->st_have * es, h->nchan, cudaMemcpyDeviceToDevice));
        for (int i = 0; i < 2; i++) { if (h->d_sbuf[i]) cudaFree(h->d_sbuf[i]); h->d_sbuf[i] = nb[i]; }
        h->st_cap = ncap; h->st_cur = 0;
    }
    if (acc > 0)
        CK(cudaMemcpy2DAsync((char *)h->d_sbuf[h->st_cur] + (size_t)h->st_have * es, (size_t)h->st_cap * es, h_iq, (size_t)in_stride * es, (size_t)acc * es,
                             h->nchan, cudaMemcpyHostToDevice, h->stream));
    h->st_have += acc; h->st_total += acc;
    const int rc = stream_step(h, final);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    /* drop what no later block can reach */
    long long keep = -1;
    for (int c = 0; c < h->nchan; c++) {
        const ChanRun &st = h->st_run[c];
        if (st.done) continue;
        const long long n_trig = st.off + (long long)(st.b + 1) * (h->K - 4);
        const long long w0 = n_trig - 1 - h->delay - (h->K + h->L - 1) - (h->lpFMtaps + 2 * CORR_R + 16);
        const long long from = w0 - h->lpIQtaps - h->maxcnt - 8;
        if (keep < 0 || from < keep) keep = from;
    }
    if (keep > h->st_abs0) {
        const long long nabs = keep / h->maxcnt * h->maxcnt;
        const long long shift = nabs - h->st_abs0;
        if (shift > 0 && shift <= h->st_have) {
            const long long rest = h->st_have - shift;
            if (rest > 0)
                CK(cudaMemcpy2DAsync(h->d_sbuf[h->st_cur ^ 1], (size_t)h->st_cap * es, (char *)h->d_sbuf[h->st_cur] + (size_t)shift * es, (size_t)h->st_cap * es,
                                     (size_t)rest * es, h->nchan, cudaMemcpyDeviceToDevice, h->stream));
            h->st_cur ^= 1; h->st_have = rest; h->st_abs0 = nabs;
        }
    }
    if (done) for (int c = 0; c < h->nchan; c++) done[c] = h->st_run[c].done ? 1 : 0;
    return B200DFT_OK;
}

extern "C" int b200dft_run_device(b200dft_t *h, const void *d_iq, int fmt, long in_stride, long nsamp, int opt_cont, double tl,
                                  b200dft_result_t *res)
{
    if (!h || !d_iq || !res) return fail(B200DFT_EINVAL, "null");
    if (fmt < 0 || fmt > 2 || nsamp < 0) return fail(B200DFT_ERANGE, "fmt/nsamp");
    CK(cudaSetDevice(h->device));
    return run_common(h, d_iq, fmt, in_stride, nsamp, opt_cont, tl, res);
}

extern "C" int b200dft_run_host(b200dft_t *h, const void *h_iq, int fmt, long in_stride, long nsamp, int opt_cont, double tl,
                                b200dft_result_t *res)
{
    if (!h || !h_iq || !res) return fail(B200DFT_EINVAL, "null");
    if (fmt < 0 || fmt > 2 || nsamp < 0) return fail(B200DFT_ERANGE, "fmt/nsamp");
    CK(cudaSetDevice(h->device));
    const int es = fmt == 1 ? 8 : (fmt == 0 ? 4 : 2);
    const size_t need = (size_t)h->nchan * nsamp * es;
    if ((long long)need > h->cap_iq_bytes) { if (h->d_iq) cudaFree(h->d_iq); CK(cudaMalloc(&h->d_iq, need + 16)); h->cap_iq_bytes = need; }
    if (nsamp > 0)
        CK(cudaMemcpy2DAsync(h->d_iq, (size_t)nsamp * es, h_iq, (size_t)in_stride * es, (size_t)nsamp * es, h->nchan, cudaMemcpyHostToDevice, h->stream));
    return run_common(h, h->d_iq, fmt, nsamp, nsamp, opt_cont, tl, res);
}


/* ---- wideband mode: `dft_detect --IQ <fq> --bw <kHz> [--dc] - <sr_base> <bits>` for nchan values of fq over ONE capture ------------- */
extern "C" int b200dft_create_wideband(b200dft_t **out, int sr_base, double bw_khz, int opt_dc, int nchan, int device)
{
    if (!out || sr_base < 1 || nchan <= 0) return fail(B200DFT_EINVAL, "bad arguments");
    /* IF rate, decimation and low-pass exactly as init_buffers() derives them (dft_detect.c:1021-1068) */
    int IF_sr = 48000, decM = 1;
    const float set_lpIQ = bw_khz * 1e3;
    if (set_lpIQ > IF_sr) IF_sr = set_lpIQ;
    const int wideIF = IF_sr > 60e3;
    if (IF_sr > sr_base) IF_sr = sr_base;
    if (IF_sr < sr_base) { while (sr_base % IF_sr) IF_sr += 1; decM = sr_base / IF_sr; }
    float f_lp = (IF_sr + 20e3) / (4.0 * sr_base), t_bw = (IF_sr - 20e3);
    if (wideIF) { f_lp = (IF_sr + 60e3) / (4.0 * sr_base); t_bw = (IF_sr - 60e3); }
    if (t_bw < 0) t_bw = 10e3;
    t_bw /= sr_base;
    int taps = 4.0 / t_bw; if (taps % 2 == 0) taps++;
    std::vector<float> ws;
    taps = lowpass_init(f_lp, taps, ws);
    const int rc = b200dft_create(out, IF_sr, bw_khz, opt_dc, nchan, device);
    if (rc) return rc;
    b200dft *h = *out;
    h->wide = true; h->sr_base = sr_base; h->decM = decM; h->dectaps = taps;
    h->wmaxcnt = IF_sr / 32; if (decM > 1) h->wmaxcnt *= decM;                     /* :1155-1156 */
    int maxsmem = 0;
    CK(cudaDeviceGetAttribute(&maxsmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    int tm = 64, budget_kb = 100;                                                  /* outputs per CTA: window <= budget (occupancy vs. re-rotated overlap) */
    if (const char *e = getenv("B200DFT_CH_SMEM_KB")) { const int v = atoi(e); if (v >= 16 && v <= 200) budget_kb = v; }   /* tuning knob for profiles/wideband_bench.py */
    while (tm > 1 && ((size_t)tm * decM + taps) * sizeof(float2) > (size_t)budget_kb * 1024) tm >>= 1;
    h->ch_tm = tm; h->ch_smem = ((size_t)tm * decM + taps) * sizeof(float2) + (size_t)taps * sizeof(float) + 16;
    if (h->ch_smem + 1024 > (size_t)maxsmem) { b200dft_destroy(h); *out = 0; return fail(B200DFT_EINVAL, "decimation filter does not fit shared memory"); }
    {   cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, (const void *)dft_chan_kernel));
        CK(cudaFuncSetAttribute(dft_chan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsmem - (int)fa.sharedSizeBytes)); }
    CK(cudaMalloc(&h->d_ws_dec, ws.size() * 4)); CK(cudaMemcpy(h->d_ws_dec, ws.data(), (size_t)taps * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_f0, sizeof(double) * nchan));
    for (int i = 0; i < 2; i++) CK(cudaEventCreate(&h->evc[i]));
    return B200DFT_OK;
}

extern "C" int b200dft_wideband_info(const b200dft_t *h, b200dft_wbinfo_t *info)
{
    if (!h || !info || !h->wide) return fail(B200DFT_EINVAL, "not a wideband handle");
    info->sr_base = h->sr_base; info->if_sr = h->sr; info->decM = h->decM; info->dectaps = h->dectaps; info->nchan = h->nchan;
    return B200DFT_OK;
}

/* rotation frequency of the LUT: freq0/sr_base with freq0 snapped like :1072-1096 */
static double wb_f0(int sr_base, double fq, int *lut_len)
{
    if (fq < -0.5) fq = -0.5;
    if (fq > 0.5) fq = 0.5;
    const double xlt = -fq;                                                        /* :1390 */
    const int W = 2 * 8;
    int d, k, freq = (int)(xlt * (double)sr_base + 0.5), freq0 = freq;
    for (d = W; d > 0; d--) if (sr_base % d == 0) break;
    if (d == 0) d = 1;
    for (k = 0; k < W / 2; k++) {
        if ((freq + k) % d == 0) { freq0 = freq + k; break; }
        if ((freq - k) % d == 0) { freq0 = freq - k; break; }
    }
    *lut_len = sr_base / d;
    return freq0 / (double)sr_base;
}

static int channelise_common(b200dft *h, const void *d_wide, int fmt, long nwide, const double *fq, float2 *d_out, long out_stride, long nout)
{
    std::vector<double> f0(h->nchan); int lut_len = 1;
    for (int c = 0; c < h->nchan; c++) f0[c] = wb_f0(h->sr_base, fq[c], &lut_len);
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(h->d_f0, f0.data(), sizeof(double) * h->nchan, cudaMemcpyHostToDevice, s));
    const long long nwavg = nwide / h->wmaxcnt + 2;
    if (nwavg > h->cap_wavg) { if (h->d_wavg) cudaFree(h->d_wavg); CK(cudaMalloc(&h->d_wavg, nwavg * sizeof(float2))); h->cap_wavg = nwavg; }
    CK(cudaEventRecord(h->evc[0], s));
    if (nout > 0) {
        DftParams p; memset(&p, 0, sizeof p);
        p.nchan = 1; p.maxcnt = h->wmaxcnt; p.nsamp = nwide; p.in_stride = 0; p.fmt = fmt; p.iq = d_wide; p.avg = h->d_wavg; p.navg = (int)nwavg;
        dft_dcavg_kernel<<<dim3((unsigned)(nwavg - 1), 1), 128, 0, s>>>(p);
        ChanParams q; memset(&q, 0, sizeof q);
        q.wide = d_wide; q.fmt = fmt; q.nwide = nwide; q.wavg = h->d_wavg; q.wmaxcnt = h->wmaxcnt; q.decM = h->decM; q.taps = h->dectaps;
        q.lut_len = lut_len; q.TM = h->ch_tm; q.ws = h->d_ws_dec; q.f0 = h->d_f0; q.out = d_out; q.nout = nout; q.out_stride = out_stride;
        dft_chan_kernel<<<dim3((unsigned)((nout + h->ch_tm - 1) / h->ch_tm), h->nchan), CH_THREADS, h->ch_smem, s>>>(q);
        h->launches += 2;
    }
    CK(cudaEventRecord(h->evc[1], s));
    CK(cudaGetLastError());
    h->ch_timed = true;
    return B200DFT_OK;
}

extern "C" int b200dft_channelise_device(b200dft_t *h, const void *d_wide, int fmt, long nwide, const double *fq, void *d_if_out, long out_stride,
                                         long *nout)
{
    if (!h || !h->wide || !d_wide || !fq || !d_if_out) return fail(B200DFT_EINVAL, "null / not a wideband handle");
    if (fmt < 0 || fmt > 2 || nwide < 0) return fail(B200DFT_ERANGE, "fmt/nwide");
    CK(cudaSetDevice(h->device));
    const long n = nwide / h->decM;                                                /* f32read_cblock: only whole blocks (:754) */
    if (out_stride < n) return fail(B200DFT_ERANGE, "out_stride");
    const int rc = channelise_common(h, d_wide, fmt, nwide, fq, (float2 *)d_if_out, out_stride, n);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    if (nout) *nout = n;
    return B200DFT_OK;
}

extern "C" int b200dft_run_wideband_device(b200dft_t *h, const void *d_wide, int fmt, long nwide, const double *fq, int opt_cont, double tl,
                                           b200dft_result_t *res)
{
    if (!h || !h->wide || !d_wide || !fq || !res) return fail(B200DFT_EINVAL, "null / not a wideband handle");
    if (fmt < 0 || fmt > 2 || nwide < 0) return fail(B200DFT_ERANGE, "fmt/nwide");
    CK(cudaSetDevice(h->device));
    long n = nwide / h->decM;
    if (tl > 0) { const long lim = (long)floor((tl + 1) * h->sr); if (n > lim) n = lim; }     /* nothing later can trigger (:1494) */
    if (n > h->cap_if) { if (h->d_if) cudaFree(h->d_if); CK(cudaMalloc(&h->d_if, (size_t)h->nchan * (n + 1) * sizeof(float2))); h->cap_if = n; }
    const int rc = channelise_common(h, d_wide, fmt, nwide, fq, h->d_if, n, n);
    if (rc) return rc;
    return run_common(h, h->d_if, B200DFT_FMT_F32, n, n, opt_cont, tl, res, true);
}

extern "C" int b200dft_run_wideband_host(b200dft_t *h, const void *h_wide, int fmt, long nwide, const double *fq, int opt_cont, double tl,
                                         b200dft_result_t *res)
{
    if (!h || !h->wide || !h_wide || !fq || !res) return fail(B200DFT_EINVAL, "null / not a wideband handle");
    if (fmt < 0 || fmt > 2 || nwide < 0) return fail(B200DFT_ERANGE, "fmt/nwide");
    CK(cudaSetDevice(h->device));
    const int es = fmt == 1 ? 8 : (fmt == 0 ? 4 : 2);
    const size_t need = (size_t)nwide * es;
    if ((long long)need > h->cap_wide_bytes) { if (h->d_wide) cudaFree(h->d_wide); CK(cudaMalloc(&h->d_wide, need + 16)); h->cap_wide_bytes = need; }
    if (nwide > 0) CK(cudaMemcpyAsync(h->d_wide, h_wide, need, cudaMemcpyHostToDevice, h->stream));
    return b200dft_run_wideband_device(h, h->d_wide, fmt, nwide, fq, opt_cont, tl, res);
}

extern "C" int b200dft_last_channelise_ms(b200dft_t *h, float *ms)
{
    if (!h || !h->ch_timed || !ms) return fail(B200DFT_EINVAL, "no channeliser run yet");
    CK(cudaEventSynchronize(h->evc[1]));
    CK(cudaEventElapsedTime(ms, h->evc[0], h->evc[1]));
    return B200DFT_OK;
}

extern "C" long b200dft_launch_count(const b200dft_t *h) { return h ? h->launches : 0; }
extern "C" int b200dft_last_kernel_ms(b200dft_t *h, float *front_ms, float *corr_ms)
{
    if (!h || !h->timed) return fail(B200DFT_EINVAL, "no run yet");
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, h->ev[0], h->ev[1])); CK(cudaEventElapsedTime(&b, h->ev[1], h->ev[2]));
    if (front_ms) *front_ms = a; if (corr_ms) *corr_ms = b;
    return B200DFT_OK;
}
