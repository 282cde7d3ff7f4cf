/* exact_math.h -- transcendental-free, rounding-explicit single-precision helpers shared by host and device.
 *
 * The reference's only per-frame libm call on the demod path is atan2f (utils/fsk.c:707).  To keep the
 * whole frame loop on the GPU *and* bit-identical to the reference, the device evaluates atan2f with the
 * same algorithm the reference's libm uses (glibc 2.39: Sun fdlibm single-precision atanf/atan2f --
 * breakpoint argument reduction + odd/even polynomial), written with explicit round-to-nearest
 * mul/add/div so nvcc can never contract an FMA into it.  tests/test_exact_math.py checks the host build of
 * this header against the system atan2f bit for bit; the GPU test checks device == host.
 */
#pragma once
#include <stdint.h>
#include <string.h>
#if defined(__CUDA_ARCH__)
#define EM_MUL(a, b) __fmul_rn((a), (b))
#define EM_ADD(a, b) __fadd_rn((a), (b))
#define EM_SUB(a, b) __fsub_rn((a), (b))
#define EM_DIV(a, b) __fdiv_rn((a), (b))
#define EM_SQRT(a)   __fsqrt_rn((a))
#else
#include <math.h>
#define EM_MUL(a, b) ((a) * (b))
#define EM_ADD(a, b) ((a) + (b))
#define EM_SUB(a, b) ((a) - (b))
#define EM_DIV(a, b) ((a) / (b))
#define EM_SQRT(a)   sqrtf((a))
#endif
#ifdef __CUDACC__
#define EM_HD __host__ __device__ __forceinline__
#else
#define EM_HD static inline
#endif

EM_HD uint32_t em_f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
EM_HD float em_u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

EM_HD float em_atanf(float x)
{
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                          9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                          4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
    int32_t hx = (int32_t)em_f2u(x), ix = hx & 0x7fffffff, id;
    if (ix >= 0x4c000000) {                       /* |x| >= 2^25 */
        if (ix > 0x7f800000) return EM_ADD(x, x); /* NaN */
        if (hx > 0) return EM_ADD(atanhi[3], atanlo[3]);
        return EM_SUB(-atanhi[3], atanlo[3]);
    }
    if (ix < 0x3ee00000) {                        /* |x| < 0.4375 */
        if (ix < 0x31000000) return x;            /* |x| < 2^-29 */
        id = -1;
    } else {
        x = em_u2f((uint32_t)ix);
        if (ix < 0x3f980000) {                    /* |x| < 1.1875 */
            if (ix < 0x3f300000) { id = 0; x = EM_DIV(EM_SUB(EM_MUL(2.0f, x), 1.0f), EM_ADD(2.0f, x)); }
            else                 { id = 1; x = EM_DIV(EM_SUB(x, 1.0f), EM_ADD(x, 1.0f)); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = EM_DIV(EM_SUB(x, 1.5f), EM_ADD(1.0f, EM_MUL(1.5f, x))); }
            else                 { id = 3; x = EM_DIV(-1.0f, x); }
        }
    }
    float z = EM_MUL(x, x), w = EM_MUL(z, z);
    float s1 = EM_MUL(z, EM_ADD(aT[0], EM_MUL(w, EM_ADD(aT[2], EM_MUL(w, EM_ADD(aT[4], EM_MUL(w, EM_ADD(aT[6],
               EM_MUL(w, EM_ADD(aT[8], EM_MUL(w, aT[10])))))))))));
    float s2 = EM_MUL(w, EM_ADD(aT[1], EM_MUL(w, EM_ADD(aT[3], EM_MUL(w, EM_ADD(aT[5], EM_MUL(w, EM_ADD(aT[7],
               EM_MUL(w, aT[9])))))))));
    if (id < 0) return EM_SUB(x, EM_MUL(x, EM_ADD(s1, s2)));
    z = EM_SUB(atanhi[id], EM_SUB(EM_SUB(EM_MUL(x, EM_ADD(s1, s2)), atanlo[id]), x));
    return (hx < 0) ? -z : z;
}

EM_HD float em_atan2f(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
                pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    int32_t hx = (int32_t)em_f2u(x), ix = hx & 0x7fffffff;
    int32_t hy = (int32_t)em_f2u(y), iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return EM_ADD(x, y);
    if (hx == 0x3f800000) return em_atanf(y);
    int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) { case 0: case 1: return y; case 2: return EM_ADD(pi, tiny); default: return EM_SUB(-pi, tiny); }
    }
    if (ix == 0) return (hy < 0) ? EM_SUB(-pi_o_2, tiny) : EM_ADD(pi_o_2, tiny);
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) { case 0: return EM_ADD(pi_o_4, tiny); case 1: return EM_SUB(-pi_o_4, tiny);
                         case 2: return EM_ADD(EM_MUL(3.0f, pi_o_4), tiny); default: return EM_SUB(EM_MUL(-3.0f, pi_o_4), tiny); }
        } else {
            switch (m) { case 0: return 0.0f; case 1: return -0.0f; case 2: return EM_ADD(pi, tiny); default: return EM_SUB(-pi, tiny); }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? EM_SUB(-pi_o_2, tiny) : EM_ADD(pi_o_2, tiny);
    int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = EM_ADD(pi_o_2, EM_MUL(0.5f, pi_lo));
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = em_atanf(em_u2f(em_f2u(EM_DIV(y, x)) & 0x7fffffffu));
    switch (m) {
        case 0: return z;
        case 1: return em_u2f(em_f2u(z) ^ 0x80000000u);
        case 2: return EM_SUB(pi, EM_SUB(z, pi_lo));
        default: return EM_SUB(EM_SUB(z, pi_lo), pi);
    }
}
