/* fsk_fft128.cuh -- four 128-point kiss-order FFTs per warp in two register passes (sm_100a).
 *
 * The generic warp FFT (fsk_kernel.cuh: warp_fft_multi) walks kiss_fft's 4*4*4*2 factorisation (kiss_fft.c:309-331) one stage at a
 * time through shared memory: 4 round trips per window, stride-2 / stride-8 accesses that bank-conflict two ways even with the
 * i + i/8 padding, and a scatter pass in front.  The ncu capture of the throughput kernel (profiles/r02_wpc_a_lines.txt) shows that
 * path at ~30 % of all shared-memory wavefronts (half of them excess) and ~35 % of the instructions.
 * Here the SAME butterflies, in the same operand order (kf_bfly2 kiss_fft.c:21-42, kf_bfly4 :44-90 forward branch), are regrouped:
 *
 *   pass 1  a lane owns one aligned block of 8 leaf slots: it loads the 8 input samples those slots stand for straight from global
 *           memory (the DIT input permutation is a fixed stride pattern: idx0 + {0,64,16,80,32,96,48,112}), multiplies by its 8
 *           Hann weights (kept in registers for the whole kernel), runs the radix-2 (m=1) and radix-4 (m=2) stages in registers
 *           and stores the block once (stride 72 B between lanes: conflict-free);
 *   pass 2  a lane owns the 16 slots {u + 8j}: radix-4 (m=8) and radix-4 (m=32) in registers, then |X| of its 16 bins goes back to
 *           shared memory as floats for the bin-parallel Sf recursion (fsk.c:484-500), which needs the windows in order.
 *
 * One shared-memory round trip instead of five, every access conflict-free, no scatter.  Bit-exactness is unaffected: butterflies
 * within a stage are independent, only their placement changed (tests/test_fsk_gpu.py runs all kernels against the oracle).
 */
#pragma once
#include "fsk_kernel.cuh"

#define F128_MSTRIDE 136         /* floats per |X| row when the magnitudes go to a per-frame buffer (MAGOUT) */
#define F128_STRIDE 152          /* float2 per FFT buffer: 144 (i + i/8 padding) + 8, so that buffers q and q+1 sit 16 banks apart */

__device__ __forceinline__ void kf_bfly2_rn(float2 &X0, float2 &X1, const float2 w)
{
    const float2 t = cmul_rn(X1, w);
    X1 = csub_rn(X0, t);
    X0 = cadd_rn(X0, t);
}
__device__ __forceinline__ void kf_bfly4_rn(float2 &X0, float2 &X1, float2 &X2, float2 &X3, const float2 w1, const float2 w2, const float2 w3)
{
    const float2 s0 = cmul_rn(X1, w1);
    const float2 s1 = cmul_rn(X2, w2);
    const float2 s2 = cmul_rn(X3, w3);
    const float2 s5 = csub_rn(X0, s1);
    const float2 a  = cadd_rn(X0, s1);
    const float2 s3 = cadd_rn(s0, s2);
    const float2 s4 = csub_rn(s0, s2);
    X2 = csub_rn(a, s3);
    X0 = cadd_rn(a, s3);
    X1 = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));
    X3 = make_float2(__fsub_rn(s5.x, s4.y), __fadd_rn(s5.y, s4.x));
}

/* input index of leaf slot 8*B + t (plan_fft in fsk_batch.cu: slot = 32 k0 + 8 k1 + 2 k2 + k3  ->  index k0 + 4 k1 + 16 k2 + 64 k3) */
__device__ __forceinline__ int f128_idx0(int B) { return (B >> 2) + 4 * (B & 3); }
__device__ __forceinline__ constexpr int f128_off(int t) { return 16 * (t >> 1) + 64 * (t & 1); }

struct F128Lane {                /* per-lane constants of the two passes */
    float h[8];                  /* Hann weights of this lane's pass-1 block */
    int   idx0;                  /* first input index of the block */
};

__device__ __forceinline__ void f128_init(F128Lane &L, const float *s_hann, int lane)
{
    L.idx0 = f128_idx0(lane & 15);
#pragma unroll
    for (int t = 0; t < 8; ++t) L.h[t] = s_hann[L.idx0 + f128_off(t)];
}

/* Windows j0 .. j0+3 of the frame whose first sample is in[gpos]: Hann, FFT, |X|, then Sf_out[i] = IIR over those windows of
 * Sf_src[i] (fsk.c:466-503).  F: 4 buffers of F128_STRIDE float2.  Windows >= numffts are skipped. */
/* MAGOUT = false: as described.  MAGOUT = true (the per-channel pipeline, where several warps share a frame's windows): |X| of
 * window j goes to magout[j * mstride + bin] (mstride = 136: rows 8 banks apart) and the caller runs the ordered recursion. */
template <int FMT, bool MAGOUT = false>
__device__ __forceinline__ void est_round128(const FskKParams &p, long long gpos, int j0, int numffts, float2 *F, const float2 *s_tw,
                                             const F128Lane &L, const float *Sf_src, float *Sf_out, int lane, float *magout = nullptr)
{
    /* ---- pass 1: blocks (q, B) with B = lane & 15, q = lane >> 4 and (lane >> 4) + 2 ---- */
    const int B = lane & 15, qa = lane >> 4;
    float2 x[2][8];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int j = j0 + qa + 2 * r;
#pragma unroll
        for (int t = 0; t < 8; ++t)
            x[r][t] = (j < numffts) ? load_sample_t<FMT>(p.in, gpos + (long long)j * 64 + L.idx0 + f128_off(t)) : make_float2(0.0f, 0.0f);
    }
    const float2 w0 = s_tw[0], w16 = s_tw[16], w32 = s_tw[32], w48 = s_tw[48];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int t = 0; t < 8; ++t) x[r][t] = make_float2(__fmul_rn(L.h[t], x[r][t].x), __fmul_rn(L.h[t], x[r][t].y));   /* fsk.c:470-478 */
        kf_bfly2_rn(x[r][0], x[r][1], w0); kf_bfly2_rn(x[r][2], x[r][3], w0);                 /* radix 2, m = 1 */
        kf_bfly2_rn(x[r][4], x[r][5], w0); kf_bfly2_rn(x[r][6], x[r][7], w0);
        kf_bfly4_rn(x[r][0], x[r][2], x[r][4], x[r][6], w0, w0, w0);                          /* radix 4, m = 2, u = 0 */
        kf_bfly4_rn(x[r][1], x[r][3], x[r][5], x[r][7], w16, w32, w48);                       /* u = 1: tw[16 u], tw[32 u], tw[48 u] */
        float2 *dst = F + (qa + 2 * r) * F128_STRIDE + 9 * B;
#pragma unroll
        for (int t = 0; t < 8; ++t) dst[t] = x[r][t];
    }
    __syncwarp();
    /* ---- pass 2: slots {u + 8 j} of FFT q, u = lane & 7, q = lane >> 3 ---- */
    const int u = lane & 7, q = lane >> 3;
    float2 y[16];
    {
        const float2 *src = F + q * F128_STRIDE + u;
#pragma unroll
        for (int j = 0; j < 16; ++j) y[j] = src[9 * j];
    }
    __syncwarp();                                                    /* every slot is in registers: the buffers may be overwritten */
    {
        const float2 a1 = s_tw[4 * u], a2 = s_tw[8 * u], a3 = s_tw[12 * u];                    /* radix 4, m = 8: tw[4 u k] */
#pragma unroll
        for (int b = 0; b < 4; ++b) kf_bfly4_rn(y[4 * b], y[4 * b + 1], y[4 * b + 2], y[4 * b + 3], a1, a2, a3);
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {                                 /* radix 4, m = 32: butterfly v = u + 8 jj, tw[v k] */
        const int v = u + 8 * jj;
        kf_bfly4_rn(y[jj], y[jj + 4], y[jj + 8], y[jj + 12], s_tw[v], s_tw[2 * v], s_tw[3 * v]);
    }
    if (MAGOUT) {
        if (j0 + q < numffts) {
            float *mag = magout + (j0 + q) * F128_MSTRIDE + u;
#pragma unroll
            for (int j = 0; j < 16; ++j) mag[8 * j] = __fsqrt_rn(mag2_rn(y[j]));
        }
        __syncwarp();
        return;
    }
    {
        float *mag = (float *)(F + q * F128_STRIDE) + 8 * q + u;      /* +8 q: the four FFTs' rows land in four different bank octets */
#pragma unroll
        for (int j = 0; j < 16; ++j) mag[8 * j] = __fsqrt_rn(mag2_rn(y[j]));
    }
    __syncwarp();
    /* ---- Sf IIR in window order, bin-parallel (fsk.c:484-500: fftshift, then Sf = Sf (1-tc) + |X| tc) ---- */
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = lane + 32 * r, src = (i + 64) & 127;
        float s = Sf_src[i];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
            if (j0 + qq < numffts) s = __fadd_rn(__fmul_rn(s, p.omtc), __fmul_rn(((const float *)(F + qq * F128_STRIDE))[8 * qq + src], p.tc));
        Sf_out[i] = s;
    }
    __syncwarp();
}

/* Compact form of est_round128<FMT, true> for kernels whose warps run DIFFERENT roles on one SM (fsk_duo_kernel): there the 32 KB
 * instruction cache level, not the issue slots, is the scarce resource (profiles/r02_notes.md: 28 % of the warp samples of the first
 * version were "no instruction", hit rate 62 %), so the body is kept near 10 KB: pass 1 is a rolled loop over the two blocks of a lane
 * with the next block's samples prefetched, and the square roots run as a rolled in-place pass over the rows.  Same butterflies,
 * same operand order, same results. */
template <int FMT>
__device__ __forceinline__ void est_round128c(const FskKParams &p, long long gpos, int j0, int numffts, float2 *F, const float2 *s_tw,
                                           const F128Lane &L, int lane, float *magout)
{
    const int B = lane & 15, qa = lane >> 4;
    const float2 w0 = s_tw[0], w16 = s_tw[16], w32 = s_tw[32], w48 = s_tw[48];
    float2 xn[8];
#pragma unroll
    for (int t = 0; t < 8; ++t)
        xn[t] = (j0 + qa < numffts) ? load_sample_t<FMT>(p.in, gpos + (long long)(j0 + qa) * 64 + L.idx0 + f128_off(t)) : make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        float2 x[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) x[t] = xn[t];
        if (r == 0) {
            const int j = j0 + qa + 2;
#pragma unroll
            for (int t = 0; t < 8; ++t)
                xn[t] = (j < numffts) ? load_sample_t<FMT>(p.in, gpos + (long long)j * 64 + L.idx0 + f128_off(t)) : make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) x[t] = make_float2(__fmul_rn(L.h[t], x[t].x), __fmul_rn(L.h[t], x[t].y));   /* fsk.c:470-478 */
        kf_bfly2_rn(x[0], x[1], w0); kf_bfly2_rn(x[2], x[3], w0);
        kf_bfly2_rn(x[4], x[5], w0); kf_bfly2_rn(x[6], x[7], w0);
        kf_bfly4_rn(x[0], x[2], x[4], x[6], w0, w0, w0);
        kf_bfly4_rn(x[1], x[3], x[5], x[7], w16, w32, w48);
        float2 *dst = F + (qa + 2 * r) * F128_STRIDE + 9 * B;
#pragma unroll
        for (int t = 0; t < 8; ++t) dst[t] = x[t];
    }
    __syncwarp();
    /* pass 2 in two rolled sub-passes over this lane's own 16 slots (same lane writes and reads: no warp sync in between) */
    const int u = lane & 7, q = lane >> 3;
    float2 *base = F + q * F128_STRIDE + u;
    {
        const float2 a1 = s_tw[4 * u], a2 = s_tw[8 * u], a3 = s_tw[12 * u];                    /* radix 4, m = 8 */
#pragma unroll 2
        for (int b = 0; b < 4; ++b) {
            float2 *sl = base + 36 * b;
            float2 y0 = sl[0], y1 = sl[9], y2 = sl[18], y3 = sl[27];
            kf_bfly4_rn(y0, y1, y2, y3, a1, a2, a3);
            sl[0] = y0; sl[9] = y1; sl[18] = y2; sl[27] = y3;
        }
    }
    {
        const bool live = j0 + q < numffts;
        float *mag = magout + (j0 + q) * F128_MSTRIDE + u;
#pragma unroll 2
        for (int jj = 0; jj < 4; ++jj) {                             /* radix 4, m = 32, then |X|^2 */
            const int v = u + 8 * jj;
            const float2 *sl = base + 9 * jj;
            float2 y0 = sl[0], y1 = sl[36], y2 = sl[72], y3 = sl[108];
            kf_bfly4_rn(y0, y1, y2, y3, s_tw[v], s_tw[2 * v], s_tw[3 * v]);
            if (live) { mag[8 * jj] = mag2_rn(y0); mag[8 * (jj + 4)] = mag2_rn(y1); mag[8 * (jj + 8)] = mag2_rn(y2); mag[8 * (jj + 12)] = mag2_rn(y3); }
        }
    }
    __syncwarp();
    const int nw = min(4, numffts - j0);
#pragma unroll 1
    for (int w = 0; w < nw; ++w) {                                   /* |X| = sqrt(|X|^2), in place, rolled */
        float *row = magout + (j0 + w) * F128_MSTRIDE + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) row[32 * k] = __fsqrt_rn(row[32 * k]);
    }
    __syncwarp();
}
