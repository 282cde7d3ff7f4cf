// Round-2 micro-benchmarks that decide the shape of the split oscillator / consumer kernel:
//   A  mbarrier.try_wait on a barrier nobody completes: iterations per microsecond for several suspend-time hints
//   B  __nanosleep(ns): cycles actually slept
//   C  lane-per-chain oscillator warp (32 independent phi *= dphi chains, utils/fsk.c:646-647): cycles per step with
//      no store / STG.64 [step][lane] / STG.128 "value+flag" words / STS.64, for 1, 4 (one per SM sub-partition) and 8 warps
//   D  cost of __threadfence() and st.release.gpu after a burst of global stores
//   E  round trip of a relaxed / acquire global load of an L2-resident word
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o r02_micro r02_micro.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 cmul_rn(float2 a, float2 b)
{
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// ---------------- A ----------------
__global__ void k_trywait(long long *out, unsigned hint, int use_hint)
{
    __shared__ unsigned long long bar;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t0 = gtime(); const long long c0 = clock64();
        long long it = 0;
        while (gtime() - t0 < 200000ull) {
            unsigned ok;
            if (use_hint)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(0u), "r"(hint) : "memory");
            else
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(0u) : "memory");
            ++it;
            if (ok) break;
        }
        out[0] = it; out[1] = clock64() - c0; out[2] = (long long)(gtime() - t0);
    }
}
// ---------------- B ----------------
__global__ void k_nanosleep(long long *out, unsigned ns)
{
    if (threadIdx.x == 0) {
        const long long c0 = clock64();
        for (int i = 0; i < 64; ++i) __nanosleep(ns);
        out[0] = (clock64() - c0) / 64;
    }
}
// ---------------- C ----------------
#define RING 4096
__device__ int g_active_lanes = 32;
template <int MODE>
__global__ void k_osc(float2 *ring, const float2 *dtab, long long *cyc, float2 *res, int steps, int nwarps)
{
    extern __shared__ float2 sring[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= nwarps) return;
    const int gw = blockIdx.x * nwarps + warp;
    const float2 d = dtab[lane & 7];
    float2 ph = make_float2(1.0f, 0.0f);
    float2 *g = ring + (size_t)gw * RING * (MODE == 2 ? 64 : 32);
    float2 *s = sring + (warp & 3) * 32 * 66;
    const long long t0 = clock64();
    if (MODE == 6) {                // software-pipelined: the pairs of the previous 8 steps are stored between the steps of the current 8
        float2 a[8], b[8];
        a[0] = cmul_rn(ph, d);
#pragma unroll
        for (int q = 1; q < 8; ++q) a[q] = cmul_rn(a[q - 1], d);
        for (int pos = 0; pos + 16 <= steps; pos += 16) {
            float2 *sa = s + lane * 66 + (pos & 63), *sb = s + lane * 66 + ((pos + 8) & 63);
            b[0] = cmul_rn(a[7], d);
#pragma unroll
            for (int q = 1; q < 8; ++q) { b[q] = cmul_rn(b[q - 1], d); if (q & 1) *reinterpret_cast<float4 *>(sa + q - 1) = make_float4(a[q - 1].x, a[q - 1].y, a[q].x, a[q].y); }
            a[0] = cmul_rn(b[7], d);
#pragma unroll
            for (int q = 1; q < 8; ++q) { a[q] = cmul_rn(a[q - 1], d); if (q & 1) *reinterpret_cast<float4 *>(sb + q - 1) = make_float4(b[q - 1].x, b[q - 1].y, b[q].x, b[q].y); }
        }
        ph = a[7];
    } else
    for (int pos = 0; pos < steps; pos += 8) {
        float2 a[8];
        a[0] = cmul_rn(ph, d);
#pragma unroll
        for (int q = 1; q < 8; ++q) a[q] = cmul_rn(a[q - 1], d);
        ph = a[7];
        if (MODE == 1) {            // STG.64, layout [slot][lane]
#pragma unroll
            for (int q = 0; q < 8; ++q) g[(size_t)((pos + q) & (RING - 1)) * 32 + lane] = a[q];
        } else if (MODE == 2) {     // STG.128 {re, flag, im, flag}
            const float fl = __int_as_float(pos / RING + 1);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4 *>(g + ((size_t)((pos + q) & (RING - 1)) * 32 + lane) * 2) = make_float4(a[q].x, fl, a[q].y, fl);
        } else if (MODE == 3) {     // STS.64 chain-major [lane][66], 64-step tile
#pragma unroll
            for (int q = 0; q < 8; ++q) s[lane * 66 + ((pos + q) & 63)] = a[q];
        } else if (MODE == 5) {     // STS.128 chain-major [lane][66], two steps per store
#pragma unroll
            for (int q = 0; q < 8; q += 2)
                *reinterpret_cast<float4 *>(s + lane * 66 + ((pos + q) & 63)) = make_float4(a[q].x, a[q].y, a[q + 1].x, a[q + 1].y);
        } else if (MODE == 7) {     // as 5, but only the first g_active_lanes lanes store (the others keep computing)
            if (lane < g_active_lanes) {
#pragma unroll
            for (int q = 0; q < 8; q += 2)
                *reinterpret_cast<float4 *>(s + lane * 66 + ((pos + q) & 63)) = make_float4(a[q].x, a[q].y, a[q + 1].x, a[q + 1].y);
            }
        } else if (MODE == 4) {     // STG.128, two steps of one lane per store: layout [slot/2][lane] float4
#pragma unroll
            for (int q = 0; q < 8; q += 2)
                *reinterpret_cast<float4 *>(g + ((size_t)(((pos + q) & (RING - 1)) >> 1) * 32 + lane) * 2) = make_float4(a[q].x, a[q].y, a[q + 1].x, a[q + 1].y);
        }
    }
    const long long t1 = clock64();
    res[gw * 32 + lane] = ph;
    if (lane == 0) cyc[gw] = t1 - t0;
}
// ---------------- D ----------------
__global__ void k_fence(float2 *ring, long long *out, int *flag)
{
    const int lane = threadIdx.x;
    long long acc_f = 0, acc_r = 0, acc_n = 0;
    for (int r = 0; r < 64; ++r) {
        for (int q = 0; q < 16; ++q) ring[(r * 16 + q) * 32 + lane] = make_float2((float)r, (float)q);
        long long c0 = clock64();
        __threadfence();
        acc_f += clock64() - c0;
        for (int q = 0; q < 16; ++q) ring[(r * 16 + q) * 32 + lane + 65536] = make_float2((float)r, (float)q);
        c0 = clock64();
        if (lane == 0) asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(flag), "r"(r) : "memory");
        __syncwarp();
        acc_r += clock64() - c0;
        for (int q = 0; q < 16; ++q) ring[(r * 16 + q) * 32 + lane + 131072] = make_float2((float)r, (float)q);
        c0 = clock64();
        if (lane == 0) asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(flag + 32), "r"(r) : "memory");
        __syncwarp();
        acc_n += clock64() - c0;
    }
    if (lane == 0) { out[0] = acc_f / 64; out[1] = acc_r / 64; out[2] = acc_n / 64; }
}
// ---------------- E ----------------
__global__ void k_poll(int *flag, long long *out)
{
    if (threadIdx.x == 0) {
        int v = 0; long long c0 = clock64();
        for (int i = 0; i < 256; ++i) { int t; asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(t) : "l"(flag + (v & 1)) : "memory"); v += t; }
        out[0] = (clock64() - c0) / 256;
        c0 = clock64();
        for (int i = 0; i < 256; ++i) { int t; asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(t) : "l"(flag + (v & 1)) : "memory"); v += t; }
        out[1] = (clock64() - c0) / 256;
        c0 = clock64();
        for (int i = 0; i < 256; ++i) { int t; asm volatile("ld.volatile.global.b32 %0, [%1];" : "=r"(t) : "l"(flag + (v & 1)) : "memory"); v += t; }
        out[2] = (clock64() - c0) / 256;
        out[3] = v;
    }
}

// ---------------- F: one oscillator warp + one helper warp that copies every finished 64-step tile to a chain-major global ring ----------------
__global__ void k_oschelper(float2 *ring, const float2 *dtab, long long *out, int *flag, int steps, int do_fence)
{
    __shared__ __align__(16) float2 tile[2][32 * 66];
    __shared__ int s_done, s_copied;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_done = 0; s_copied = 0; }
    __syncthreads();
    if (warp == 0) {
        const float2 d = dtab[lane & 7];
        float2 ph = make_float2(1.0f, 0.0f);
        const long long t0 = clock64();
        for (int t = 0; t * 64 < steps; ++t) {
            while (t - *(volatile int *)&s_copied >= 2) __nanosleep(100);
            float2 *s = tile[t % 2] + lane * 66;
            for (int pos = 0; pos < 64; pos += 8) {
                float2 a[8];
                a[0] = cmul_rn(ph, d);
#pragma unroll
                for (int q = 1; q < 8; ++q) a[q] = cmul_rn(a[q - 1], d);
                ph = a[7];
#pragma unroll
                for (int q = 0; q < 8; q += 2) *reinterpret_cast<float4 *>(s + pos + q) = make_float4(a[q].x, a[q].y, a[q + 1].x, a[q + 1].y);
            }
            __syncwarp();
            if (lane == 0) asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_done)), "r"(t + 1) : "memory");
        }
        if (lane == 0) { out[0] = clock64() - t0; out[4] = __float_as_int(ph.x); }
    } else if (warp == 1) {
        long long acc_copy = 0, acc_fence = 0, acc_wait = 0;
        for (int t = 0; t * 64 < steps; ++t) {
            long long c0 = clock64();
            if (lane == 0) { int v; do { asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&s_done)) : "memory"); if (v <= t) __nanosleep(100); } while (v <= t); }
            __syncwarp();
            acc_wait += clock64() - c0; c0 = clock64();
            const float2 *s = tile[t % 2];
            for (int ch = 0; ch < 32; ++ch) {
                const float4 v = *reinterpret_cast<const float4 *>(s + ch * 66 + 2 * lane);
                *reinterpret_cast<float4 *>(ring + (size_t)ch * RING + ((t * 64) & (RING - 1)) + 2 * lane) = v;
            }
            acc_copy += clock64() - c0; c0 = clock64();
            if (do_fence) __threadfence();
            if (lane == 0) { asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(flag), "r"((t + 1) * 64) : "memory"); *(volatile int *)&s_copied = t + 1; }
            __syncwarp();
            acc_fence += clock64() - c0;
        }
        if (lane == 0) { out[1] = acc_wait; out[2] = acc_copy; out[3] = acc_fence; }
    }
}

// ---------------- G: oscillator warp(s) + one helper warp whose lanes ship finished tiles with cp.async.bulk (TMA) and publish a per-tile flag ----------------
#define GT 64
#define GROW 66
#define GNBUF 3
__global__ void k_osctma(float2 *ring, const float2 *dtab, long long *out, int *flags, int steps, int nosc, int lanes_per_warp)
{
    extern __shared__ __align__(128) unsigned char smem_g[];
    float2 *tiles = (float2 *)smem_g;                  // [nosc][GNBUF][lanes_per_warp][GROW]
    __shared__ int s_done[4], s_free[4][GNBUF];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_elems = lanes_per_warp * GROW;
    if (threadIdx.x < 4) { s_done[threadIdx.x] = 0; for (int b = 0; b < GNBUF; ++b) s_free[threadIdx.x][b] = 0; }
    __syncthreads();
    if (warp < nosc) {
        const float2 d = dtab[lane & 7];
        float2 ph = make_float2(1.0f, 0.0f);
        const long long t0 = clock64();
        for (int t = 0; t * GT < steps; ++t) {
            const int b = t % GNBUF;
            if (t >= GNBUF) { while (*(volatile int *)&s_free[warp][b] < t - GNBUF + 1) __nanosleep(50); }
            float2 *s = tiles + ((size_t)(warp * GNBUF + b)) * tile_elems + lane * GROW;
            for (int pos = 0; pos < GT; pos += 8) {
                float2 a[8];
                a[0] = cmul_rn(ph, d);
#pragma unroll
                for (int q = 1; q < 8; ++q) a[q] = cmul_rn(a[q - 1], d);
                ph = a[7];
                if (lane < lanes_per_warp) {
#pragma unroll
                for (int q = 0; q < 8; q += 2) *reinterpret_cast<float4 *>(s + pos + q) = make_float4(a[q].x, a[q].y, a[q + 1].x, a[q + 1].y);
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_done[warp])), "r"(t + 1) : "memory");
        }
        if (lane == 0) { out[warp] = clock64() - t0; out[8 + warp] = __float_as_int(ph.x); }
    } else if (warp == nosc) {
        // helper warp: lane = ow * GNBUF + b owns buffer b of oscillator warp ow
        const int ow = lane / GNBUF, b = lane % GNBUF;
        long long lat = 0; int ntile = 0;
        if (ow < nosc) {
            const unsigned bytes = (unsigned)tile_elems * 8u;
            for (int t = b; t * GT < steps; t += GNBUF) {
                int v; do { asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&s_done[ow])) : "memory"); if (v <= t) __nanosleep(50); } while (v <= t);
                const long long c0 = clock64();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const float2 *src = tiles + ((size_t)(ow * GNBUF + b)) * tile_elems;
                float2 *dst = ring + ((size_t)(blockIdx.x * 4 + ow) * 64 + (t & 63)) * tile_elems;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"((unsigned)__cvta_generic_to_shared(src)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                *(volatile int *)&s_free[ow][b] = t + 1;                 // smem buffer reusable
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                __threadfence();
                asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(flags + (blockIdx.x * 4 + ow) * 64 + (t & 63)), "r"(t + 1) : "memory");
                lat += clock64() - c0; ++ntile;
            }
            out[16 + lane] = ntile ? lat / ntile : 0;
        }
    }
}

template <int MODE> void run_osc(const char *name, int ctas, int nwarps, float2 *ring, float2 *dtab, long long *cyc, float2 *res)
{
    const int steps = 1 << 16;
    cudaFuncSetAttribute(k_osc<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_osc<MODE><<<ctas, 32 * nwarps, 200 * 1024>>>(ring, dtab, cyc, res, steps, nwarps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[64]; cudaMemcpy(h, cyc, sizeof(long long) * ctas * nwarps, cudaMemcpyDeviceToHost);
    double mx = 0, mn = 1e30; for (int i = 0; i < ctas * nwarps; ++i) { double v = h[i] / (double)steps; mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
    printf("C %-46s ctas=%d warps/cta=%d  cycles/step min %.2f max %.2f\n", name, ctas, nwarps, mn, mx);
}

int main()
{
    long long *out; cudaMalloc(&out, 4096); long long h[8];
    for (int use = 0; use < 2; ++use)
        for (unsigned hint : {0u, 1000u, 20000u, 1000000u}) {
            if (!use && hint) continue;
            k_trywait<<<1, 32>>>(out, hint, use); cudaDeviceSynchronize();
            cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
            printf("A try_wait %s hint=%u ns: %lld iterations in %lld ns (%lld cycles) => %.1f ns / call\n", use ? "with" : "no", hint, h[0], h[2], h[1], (double)h[2] / h[0]);
        }
    for (unsigned ns : {0u, 20u, 100u, 200u, 500u, 1000u, 4000u}) {
        k_nanosleep<<<1, 32>>>(out, ns); cudaDeviceSynchronize(); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("B nanosleep(%u): %lld cycles\n", ns, h[0]);
    }
    float2 *ring, *dtab, *res; long long *cyc; int *flag;
    cudaMalloc(&ring, (size_t)64 * RING * 64 * 8); cudaMalloc(&dtab, 64); cudaMalloc(&res, 64 * 32 * 8); cudaMalloc(&cyc, 64 * 8); cudaMalloc(&flag, 65536); cudaMemset(flag, 0, 65536);
    float2 hd[8] = {{0.9510565f, -0.3090171f}, {0.9510564f, 0.3090169f}, {0.8090170f, 0.5877852f}, {0.3090170f, 0.9510565f},
                    {0.9510565f, 0.3090171f}, {0.9510564f, -0.3090169f}, {0.8090170f, -0.5877852f}, {0.3090170f, -0.9510565f}};
    cudaMemcpy(dtab, hd, sizeof hd, cudaMemcpyHostToDevice);
    for (int w : {1, 2, 4}) {
        run_osc<0>("no store", 1, w, ring, dtab, cyc, res);
        run_osc<1>("STG.64 [slot][lane]", 1, w, ring, dtab, cyc, res);
        run_osc<2>("STG.128 value+flag words", 1, w, ring, dtab, cyc, res);
        run_osc<4>("STG.128 two steps per store", 1, w, ring, dtab, cyc, res);
        run_osc<3>("STS.64 chain-major [lane][66]", 1, w, ring, dtab, cyc, res);
        run_osc<5>("STS.128 chain-major, 2 steps/store", 1, w, ring, dtab, cyc, res);
        run_osc<6>("STS.128 chain-major, software-pipelined", 1, w, ring, dtab, cyc, res);
    }
    run_osc<1>("STG.64 [slot][lane], 4 CTAs", 4, 4, ring, dtab, cyc, res);
    run_osc<2>("STG.128 value+flag, 4 CTAs", 4, 4, ring, dtab, cyc, res);
    for (int L : {8, 16, 32}) {
        cudaMemcpyToSymbol(g_active_lanes, &L, 4);
        char nm[64]; snprintf(nm, sizeof nm, "STS.128 chain-major, %d storing lanes", L);
        run_osc<7>(nm, 1, 1, ring, dtab, cyc, res); run_osc<7>(nm, 1, 4, ring, dtab, cyc, res);
    }
    {
        cudaFuncSetAttribute(k_osctma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        const int steps = 1 << 16;
        const int cfgs[5][3] = {{1, 1, 32}, {1, 2, 32}, {1, 4, 8}, {1, 4, 16}, {8, 4, 8}};   // ctas, osc warps, lanes per warp
        for (auto &c : cfgs) {
            cudaMemset(out, 0, 512);
            k_osctma<<<c[0], 32 * (c[1] + 1), (size_t)c[1] * GNBUF * c[2] * GROW * 8>>>(ring, dtab, out, flag, steps, c[1], c[2]);
            cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("G: %s\n", cudaGetErrorString(e)); break; }
            long long g[64]; cudaMemcpy(g, out, sizeof g, cudaMemcpyDeviceToHost);
            printf("G osc+TMA ctas=%d oscwarps=%d lanes=%d: osc cycles/step w0 %.2f w%d %.2f; tile ship+publish latency %lld cycles (buffer 1: %lld)\n", c[0], c[1], c[2],
                   g[0] / (double)steps, c[1] - 1, g[c[1] - 1] / (double)steps, g[16], g[17]);
        }
    }
    for (int f = 0; f < 2; ++f) {
        const int steps = 1 << 16;
        k_oschelper<<<1, 64>>>(ring, dtab, out, flag, steps, f); cudaDeviceSynchronize(); cudaMemcpy(h, out, 40, cudaMemcpyDeviceToHost);
        printf("F osc+helper fence=%d: osc %.2f cycles/step; helper per tile: wait %.0f copy %.0f fence+publish %.0f cycles\n", f, h[0] / (double)steps,
               h[1] / (steps / 64.0), h[2] / (steps / 64.0), h[3] / (steps / 64.0));
    }
    k_fence<<<1, 32>>>(ring, out, flag); cudaDeviceSynchronize(); cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    printf("D after 16 STG.64: __threadfence %lld cycles, st.release.gpu %lld cycles, st.relaxed %lld cycles\n", h[0], h[1], h[2]);
    k_poll<<<1, 32>>>(flag, out); cudaDeviceSynchronize(); cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    printf("E dependent global load round trip: relaxed %lld, acquire %lld, volatile %lld cycles\n", h[0], h[1], h[2]);
    return 0;
}
