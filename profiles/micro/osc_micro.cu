// Micro-benchmark: cycles per step of the per-tone oscillator recurrence phi *= dphi (fsk.c:646-647)
// in several instruction forms.  One CTA; timing by clock64 in lane 0 of the measured warp.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#define STEPS 1500
#define REPS 40
__device__ __forceinline__ float2 cmul_rn(float2 a, float2 b) {
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ unsigned long long pk(float lo, float hi) { return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo); }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__device__ __forceinline__ void sts128(float2 *dst, float2 a, float2 b)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
}
// the software-pipelined loop of fsk_pipe_kernel.cuh::osc_pairs (two register sets, stores of one set between the steps of the other)
__device__ __forceinline__ float2 osc_pipelined(float2 ph, const float2 d, float2 *dst, int n)
{
    int j = 0;
    float2 a[8], b[8];
    a[0] = cmul_rn(ph, d);
#pragma unroll
    for (int s = 1; s < 8; ++s) a[s] = cmul_rn(a[s - 1], d);
    for (; j + 32 <= n; j += 16) {
        b[0] = cmul_rn(a[7], d);
#pragma unroll
        for (int s = 1; s < 8; ++s) { b[s] = cmul_rn(b[s - 1], d); if (s & 1) sts128(dst + j + s - 1, a[s - 1], a[s]); }
        a[0] = cmul_rn(b[7], d);
#pragma unroll
        for (int s = 1; s < 8; ++s) { a[s] = cmul_rn(a[s - 1], d); if (s & 1) sts128(dst + j + 8 + s - 1, b[s - 1], b[s]); }
    }
#pragma unroll
    for (int s = 1; s < 8; s += 2) sts128(dst + j + s - 1, a[s - 1], a[s]);
    ph = a[7]; j += 8;
    for (; j < n; j += 2) { float2 x = cmul_rn(ph, d); ph = cmul_rn(x, d); *(float4 *)(dst + j) = make_float4(x.x, x.y, ph.x, ph.y); }
    return ph;
}

template <int MODE> __global__ void k(float2 *out, const float2 *dtab, long long *cyc, int nactive_warps)
{
    extern __shared__ float2 ring[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= nactive_warps) return;
    float2 d = dtab[lane & 1], ph = make_float2(1.f, 0.f);
    float2 *dst = ring + (warp * 2 + (lane & 1)) * 1536;
    long long t0 = clock64();
    if (lane < 2) {
        for (int r = 0; r < REPS; ++r) {
            if (MODE == 0) {           // scalar, unroll 4, STS.64 per step (current kernel)
#pragma unroll 4
                for (int j = 0; j < STEPS; ++j) { ph = cmul_rn(ph, d); dst[j] = ph; }
            } else if (MODE == 1) {    // scalar, unroll 8, STS.128 per two steps
#pragma unroll 4
                for (int j = 0; j < STEPS; j += 2) { float2 a = cmul_rn(ph, d); ph = cmul_rn(a, d); *(float4 *)(dst + j) = make_float4(a.x, a.y, ph.x, ph.y); }
            } else if (MODE == 2) {    // packed f32x2: (pr,pi)*(dr,dr) + (pi,pr)*(-di,di)
                unsigned long long drr = pk(d.x, d.x), dii = pk(-d.y, d.y), p = pk(ph.x, ph.y);
                const float onef = dtab[2].x; const unsigned long long one2 = pk(onef, onef);
#pragma unroll 4
                for (int j = 0; j < STEPS; ++j) {
                    unsigned long long ps = (p >> 32) | (p << 32);
                    // t + u as fma(t, 1, u) with an opaque 1.0 so ptxas cannot contract a product into the add
                    p = fma2(mul2(p, drr), one2, mul2(ps, dii));
                    *(unsigned long long *)(dst + j) = p;
                }
                ph = make_float2(__uint_as_float((unsigned)p), __uint_as_float((unsigned)(p >> 32)));
            } else if (MODE == 4) {    // packed products (FMUL2 x2) + scalar adds, STS.128 per 2 steps
                const unsigned long long drr = pk(d.x, d.x), dii = pk(-d.y, d.y);
#pragma unroll 4
                for (int j = 0; j < STEPS; j += 2) {
                    unsigned long long p = pk(ph.x, ph.y), ps = pk(ph.y, ph.x);
                    unsigned long long t = mul2(p, drr), u = mul2(ps, dii);
                    float2 a = make_float2(__fadd_rn(__uint_as_float((unsigned)t), __uint_as_float((unsigned)u)),
                                           __fadd_rn(__uint_as_float((unsigned)(t >> 32)), __uint_as_float((unsigned)(u >> 32))));
                    p = pk(a.x, a.y); ps = pk(a.y, a.x);
                    t = mul2(p, drr); u = mul2(ps, dii);
                    ph = make_float2(__fadd_rn(__uint_as_float((unsigned)t), __uint_as_float((unsigned)u)),
                                     __fadd_rn(__uint_as_float((unsigned)(t >> 32)), __uint_as_float((unsigned)(u >> 32))));
                    *(float4 *)(dst + j) = make_float4(a.x, a.y, ph.x, ph.y);
                }
            } else if (MODE == 5) {    // software-pipelined, STS.128 of the previous 8 steps between the current ones
                ph = osc_pipelined(ph, d, dst, STEPS - (STEPS & 1));
            } else if (MODE == 3) {    // scalar without any store (pure chain)
#pragma unroll 8
                for (int j = 0; j < STEPS; ++j) { ph = cmul_rn(ph, d); }
            }
            if (r % 7 != 3) {   // keep some un-normalised reps so that rounding differences between variants would show in the result
                const float av = __fsqrt_rn(__fadd_rn(__fmul_rn(ph.x, ph.x), __fmul_rn(ph.y, ph.y)));
                ph = make_float2(__fdiv_rn(ph.x, av), __fdiv_rn(ph.y, av));
            }
        }
    }
    long long t1 = clock64();
    if (lane < 2) out[warp * 2 + lane] = ph;
    if (lane == 0) cyc[warp] = t1 - t0;
}
template <int MODE> void run(const char *name, int warps, float2 *out, float2 *dtab, long long *cyc)
{
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<MODE><<<1, 256, 8 * 2 * 1536 * 8>>>(out, dtab, cyc, warps);
    cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[8]; float2 o[16];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost); cudaMemcpy(o, out, sizeof o, cudaMemcpyDeviceToHost);
    printf("%-44s warps=%d  cycles/step: w0=%.2f w%d=%.2f   phi0=(%a,%a)\n", name, warps, h[0] / double(STEPS * REPS),
           warps - 1, h[warps - 1] / double(STEPS * REPS), o[0].x, o[0].y);
}
int main()
{
    float2 *out, *dtab; long long *cyc;
    cudaMalloc(&out, 256); cudaMalloc(&dtab, 32); cudaMalloc(&cyc, 64);
    float2 hd[3] = {{0.9510565f, -0.3090171f}, {0.9510564f, 0.3090169f}, {1.0f, 1.0f}};
    cudaMemcpy(dtab, hd, sizeof hd, cudaMemcpyHostToDevice);
    for (int w : {1, 2, 4, 5, 8}) {
        run<0>("scalar unroll4 STS.64 (current)", w, out, dtab, cyc);
    }
    run<1>("scalar, STS.128 per 2 steps", 1, out, dtab, cyc);
    run<2>("packed f32x2 (mul2,mul2,add2 + swap)", 1, out, dtab, cyc);
    run<2>("packed f32x2 (mul2,mul2,add2 + swap)", 4, out, dtab, cyc);
    run<3>("scalar, no store (pure chain)", 1, out, dtab, cyc);
    run<4>("FMUL2 products + scalar FADD, STS.128/2", 1, out, dtab, cyc);
    run<4>("FMUL2 products + scalar FADD, STS.128/2", 5, out, dtab, cyc);
    for (int w : {1, 2, 5, 8}) run<5>("software-pipelined 2x8 steps, STS.128 interleaved", w, out, dtab, cyc);
    for (int w : {2, 5, 8}) run<1>("scalar, STS.128 per 2 steps", w, out, dtab, cyc);
    return 0;
}
