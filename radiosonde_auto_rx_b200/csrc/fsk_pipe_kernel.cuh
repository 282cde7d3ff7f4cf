/* fsk_pipe_kernel.cuh -- warp-specialised, software-pipelined fsk_demod frame kernel (sm_100a).
 *
 * Why: the per-tone local oscillator phi_c *= dphi (utils/fsk.c:646-647) is a strictly serial complex
 * recurrence -- 4 FMUL + 2 FADD per sample that must stay un-fused and in order to be bit-exact -- and one
 * warp issues one FP32 instruction per ~2 cycles (profiles/r01_notes.md), so ONE channel can never exceed
 * ~1 sample per 11-14 SM cycles.  Everything else in fsk_demod_core is parallel or much shorter.  This kernel
 * therefore makes the oscillator warp the only thing on the critical path and keeps it running without ever
 * waiting, with four warp roles inside the channel's CTA communicating through shared-memory rings and
 * release/acquire counters:
 *
 *   EST  (5 warps)  frame k+1:  windowed FFTs straight from HBM/L2 -> |X| -> Sf IIR -> tone estimate -> dphi LUT
 *                               (needs only c_{k+1} = c_k + nin_k, which is known when frame k starts)  fsk.c:438-590
 *   OSC  (1 warp)   frame k:    phi recurrence, one lane per tone, into a ring indexed by absolute sample
 *                               number; never waits in steady state                                 fsk.c:644-651
 *   CONS (4 warps)  frame k:    stream behind OSC in 128-sample blocks: x*conj(phi) in place in the ring,
 *                               then every integrate-and-dump window that became complete            fsk.c:648,659-668
 *   TIME (1 warp)   frame k:    streams behind CONS: timing terms, the sequential timing accumulate, exact
 *                               atan2f, nin/ppm, publishes frame k+1                                 fsk.c:682-731
 *   DEC  (1 warp)   frame k:    re-sample + decide + outputs to HBM, sequential EbNo sums, frame record fsk.c:737-817
 *
 * The integrator memory of the reference (last Nmem-nin down-converted samples, fsk.c:637-640) falls out of
 * indexing f_dc by absolute sample number: a sample is mixed exactly once, with the oscillator of the frame
 * that first saw it, and frame k's windows start at c_{k+1}-Nmem.
 * Used when numffts = floor(nin/(Ndft/2))-1 is the same for the three possible nin (true for every auto_rx
 * production line); otherwise the host falls back to the serial kernel in fsk_kernel.cuh.
 * Arithmetic, operand order and summation order are identical to fsk_kernel.cuh, hence to the reference.
 */
#pragma once
#include "fsk_kernel.cuh"
#include "fsk_fft128.cuh"

#ifdef FSK_PROFILE
#define PTIC long long pt_ = clock64();
#define PTOC(a) do { pacc[a] += clock64() - pt_; } while (0)
#define PDECL long long pacc[4] = {0, 0, 0, 0}; const long long prole0_ = clock64();
#define PFLUSH(base, n) do { if (lane == 0 && p.prof) { for (int q_ = 0; q_ < (n); ++q_) p.prof[blockIdx.x * 16 + (base) + q_] += pacc[q_]; p.prof[blockIdx.x * 16 + (base) + (n)] += clock64() - prole0_; } } while (0)
#else
#define PTIC
#define PTOC(a)
#define PDECL
#define PFLUSH(base, n)
#endif

#ifndef PIPE_GRING
#define PIPE_GRING 0            /* experiment, not the default (measured: 19.8 vs 20.8 Gsps at 256 channels, 13.7 vs 12.9 at 128): 1: the oscillator writes its phases to a per-channel ring in GLOBAL memory (L2) instead of shared memory.
                                   Its stores then queue in the SM's global/local instruction queue, not in the shared-memory queue that the
                                   FFT / mixer / integrator traffic of both resident CTAs keeps busy; in the ncu capture 40 % of the oscillator
                                   loop's samples were the first FMUL waiting on registers that queued STS.128 had not read yet
                                   (profiles/r02_notes.md).  The mixer reads the phases back from L2. */
#endif
#define PIPE_THREADS 384        /* 12 warps: OSC 1, TIME 1, DEC 1, CONS 4, EST 4, 1 idle (keeps the oscillator's sub-partition quiet) */
#define PIPE_R    1024          /* ring capacity per tone, samples (power of two, multiple of PIPE_BLK) */
#define PIPE_RS   (PIPE_R + 2)  /* row stride: the tones' rows sit 16 B apart in bank space, so the oscillator's two 16-byte stores (one lane per
                                   tone) no longer collide -- they were 2-way conflicts, 22 % of the kernel's shared-memory wavefronts (profiles/r02) */
#define PIPE_BLK  (32 * PIPE_NCONS)   /* consumer block = one sample per CONS thread */
#ifndef PIPE_VARIANT
#define PIPE_VARIANT 0
#endif
#if PIPE_VARIANT == 1          /* experiment: 5 EST warps, no idle warp */
#define PIPE_NEST 5
#define PIPE_NCONS 4
#elif PIPE_VARIANT == 2        /* experiment: 3 CONS (96-sample blocks), 5 EST, idle kept */
#define PIPE_NEST 5
#define PIPE_NCONS 3
#else
#define PIPE_NEST 4             /* EST warps  */
#define PIPE_NCONS 4            /* CONS warps */
#endif
#define PIPE_TB   64            /* timing terms per batch */
#define PIPE_OSC_CHUNK 256       /* oscillator steps between flow-control checks / position publications */
#define PIPE_NFB  3             /* FFTs a warp advances together */

enum { ROLE_OSC = 0, ROLE_TIME, ROLE_DEC, ROLE_CONS, ROLE_EST, ROLE_IDLE };

struct PipeDesc {               /* per-frame descriptor, 4 slots */
    int    c;                   /* frame start, samples relative to this call's stream start */
    int    nin;
    float2 dphi[4];
    int    tabidx[4];
    float2 phi_end[4];
    /* timing result of the frame (TIME -> DEC) */
    int    nan, low, high, nin_next, exists_next;
    float  fract, nrt, ppm;
};

__device__ __forceinline__ int ld_acq(const int *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel(int *p, int v)
{
    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
/* Counter + doorbell.  Correctness rests on the monotonic shared-memory counters (release/acquire); the doorbell
 * is an mbarrier (arrival count 1) the producer rings after every counter update so that a waiter can SUSPEND in
 * hardware (mbarrier.try_wait with a time hint) instead of spinning: polling loops were 62 % of all issued
 * instructions before (profiles/r01_notes.md).  A waiter that falls two rings behind merely sleeps one hint. */
enum { DB_PUB = 0, DB_EST, DB_OSC, DB_FREE, DB_FINT, DB_TIME, DB_DEC, DB_COUNT };

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_ring(unsigned long long *bar)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity, unsigned hint_ns)
{
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity), "r"(hint_ns) : "memory");
    return ok != 0;
}
/* Warp-uniform wait for `*flag >= want` (lane 0 waits, result broadcast); gives up when `*stop` is raised. */
__device__ __forceinline__ bool warp_wait_ge(const int *flag, int want, const int *stop, int lane,
                                             unsigned long long *bell, unsigned &par, unsigned hint_ns = 1000)
{
    int ok = 1;
    if (lane == 0) {
        while (ld_acq(flag) < want) {
            if (ld_acq(stop)) { ok = (ld_acq(flag) >= want); break; }
            if (mbar_try_wait(bell, par, hint_ns)) par ^= 1u;
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    return ok != 0;
}
__device__ __forceinline__ void group_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

/* 16-byte shared store that the scheduler must keep where it is written (see osc_pairs) */
__device__ __forceinline__ void sts128(float2 *dst, float2 a, float2 b)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
}

__device__ __forceinline__ void sts128a(unsigned addr, float2 a, float2 b)      /* the same on a shared-window byte address */
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
}

/* n oscillator steps (n even, dst 16-byte aligned) written to consecutive ring slots; paired 16-byte stores.
 * The recurrence is one dependent FMUL->FADD chain per lane; what it must not do is wait for its own stores.  A shared store reads its
 * source registers only when the memory pipe dequeues it, and until then the next write to those registers stalls (write-after-read on
 * the short scoreboard: 18 % of this warp's samples in the round-1 profile, more when a second CTA loads the same SM's memory pipe).
 * With PIPE_OSC_PIPELINED the loop is software-pipelined over two register sets of 8 steps: while set B is being computed the pairs of
 * set A are stored one by one between the steps, and vice versa -- every stored register is rewritten 8 steps (~100 cycles) after its
 * store was issued.  Measured (profiles/r01_notes.md): 17.8 -> 15.1 cycles/step with one channel per SM, but no gain (22-24 -> 26) with
 * two channels per SM, where the stores queue behind the other warps' shared-memory traffic wherever they are placed; the headline
 * configuration (256 channels on 148 SMs) is in the second regime, so the default stays the plain loop. */
#ifndef PIPE_OSC_PIPELINED
#define PIPE_OSC_PIPELINED 0
#endif
__device__ __forceinline__ float2 osc_pairs(float2 ph, const float2 d, float2 *dst, int n, bool pipelined = false)
{
    int j = 0;
#if !PIPE_GRING
    if (PIPE_OSC_PIPELINED || pipelined) {
        /* 16 steps per iteration in 16 registers, every pair stored ONE pair after it was computed; the last two pairs of an iteration
         * are stored at the start of the next one (after the loop for the last iteration).  What this arrangement buys:
         *  - a stored register is rewritten >= 13 steps (~140 cycles) after its store was issued (no write-after-read wait on the data),
         *  - no store in the last two steps of the iteration, so the pointer increment at the loop end does not wait for a store that
         *    still holds the address register (that wait was 10 % of the step time in the ncu capture: profiles/r02_notes.md),
         *  - no prologue/epilogue code: with the control code around it the oscillator loop stays inside the 6 KB L0 instruction
         *    cache of its sub-partition. */
        float2 v[16];
        unsigned sa = (unsigned)__cvta_generic_to_shared(dst);       /* byte address of dst + j */
        bool pend = false;
#pragma unroll 1
        for (; j + 16 <= n; j += 16) {
            v[0] = cmul_rn(ph, d); v[1] = cmul_rn(v[0], d);
            if (pend) sts128a(sa - 32, v[12], v[13]);
            v[2] = cmul_rn(v[1], d); v[3] = cmul_rn(v[2], d);
            if (pend) sts128a(sa - 16, v[14], v[15]);
            sts128a(sa, v[0], v[1]);
#pragma unroll
            for (int q = 2; q < 8; ++q) {
                v[2 * q] = cmul_rn(v[2 * q - 1], d); v[2 * q + 1] = cmul_rn(v[2 * q], d);
                if (q < 7) sts128a(sa + 16 * (q - 1), v[2 * q - 2], v[2 * q - 1]);
            }
            ph = v[15]; pend = true; sa += 128;
        }
        if (pend) { sts128a(sa - 32, v[12], v[13]); sts128a(sa - 16, v[14], v[15]); }
#pragma unroll 1
        for (; j < n; j += 2) {
            const float2 t = cmul_rn(ph, d);
            ph = cmul_rn(t, d);
            sts128(dst + j, t, ph);
        }
        return ph;
    }
#endif
#pragma unroll 4
    for (; j < n; j += 2) {
        const float2 a = cmul_rn(ph, d);
        ph = cmul_rn(a, d);
        *reinterpret_cast<float4 *>(dst + j) = make_float4(a.x, a.y, ph.x, ph.y);
    }
    return ph;
}

template <int M, int FMT, bool OSCP>      /* OSCP: software-pipelined oscillator stores (one CTA per SM) */
__global__ void __launch_bounds__(PIPE_THREADS, 2) fsk_demod_pipe_kernel(const FskKParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ PipeDesc desc[4];
    __shared__ int s_published, s_est_done, s_osc_pos, s_free_pos, s_fint_total, s_time_done, s_decide_done, s_stop;
    __shared__ int s_frames, s_nin_state, s_high_last, s_nan_last;
    __shared__ float s_nrt_state, s_ppm_state;
    __shared__ float2 s_phi0[4];
    __shared__ unsigned long long s_bell[DB_COUNT];
    __shared__ int s_ok_cons, s_ok_est;      /* verdict of a group's polling warp, published through the group's named barrier */

    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Ndft = p.Ndft, Nmem = p.Nmem, nfi = p.nfi, P = p.P, Ts = p.Ts, nsym = p.Nsym;
    const int step = Ts / P, nin_min = p.N - Ts / 2;
    constexpr int lgM = (M == 2) ? 1 : 2;

    float2 *ring   = (float2 *)(smem + p.q_ring);     /* [M][PIPE_R]: phi, then f_dc in place */
    float2 *fintb  = (float2 *)(smem + p.q_fint);     /* [2][M][nfi]: frame k uses buffer k&1 */
    float2 *fftbuf = (float2 *)(smem + p.q_fft);      /* [PIPE_NEST][Ndft] */
    float  *magbuf = (float *)(smem + p.q_mag);       /* [numffts][Ndft]: |X| of every window of one frame */
    float  *Sfbuf  = (float *)(smem + p.q_Sf);        /* [2][Ndft]: Sf after frame k lives in buffer k&1 */
    float  *s_hann = (float *)(smem + p.q_hann);
    float2 *s_tw   = (float2 *)(smem + p.q_tw);
    unsigned short *s_iperm = (unsigned short *)(smem + p.q_perm);
    float  *scr    = (float *)(smem + p.q_scr);
    float  *s_mask = (float *)(smem + p.q_mask);
    float2 *s_maxs = (float2 *)(smem + p.q_smax);
    float  *tre    = (float *)(smem + p.q_tbuf), *tim = tre + PIPE_TB;     /* [PIPE_TB] timing terms of one batch, real / imaginary */

    /* ---- prologue ---- */
#pragma unroll 1
    for (int i = tid; i < Ndft; i += PIPE_THREADS) {
        s_hann[i] = p.hann[i]; s_tw[i] = p.tw[i]; s_iperm[i] = p.invperm[i]; s_mask[i] = p.mask[i];
        Sfbuf[Ndft + i] = p.Sf[(long long)c * Ndft + i];                 /* "after frame -1" -> buffer 1 */
    }
    for (int e = tid; e < M * p.tailcap; e += PIPE_THREADS) {           /* integrator memory -> ring slots [-tailcap,0) */
        const int m = e / p.tailcap, i = e - m * p.tailcap;
        ring[m * PIPE_RS + ((i - p.tailcap) & (PIPE_R - 1))] = p.fdc_tail[((long long)c * M + m) * p.tailcap + i];
    }
    const long long start = p.in_off ? (long long)p.in_off[c] : 0;
    const long long gbase = (long long)c * p.in_stride + start;
    const int nsamp_rel = (int)(p.nsamp - start);
    const int fbase = p.frame_base ? p.frame_base[c] : 0;
    if (tid < M) s_phi0[tid] = p.phi_c[c * 4 + tid];
    if (tid == 0) {
        const int nin0 = p.nin[c];
        s_nin_state = nin0; s_nrt_state = p.nrt[c]; s_ppm_state = p.ppm[c];
        s_frames = 0; s_high_last = 0; s_nan_last = 0;
        s_est_done = 0; s_osc_pos = 0; s_free_pos = -p.tailcap; s_fint_total = 0; s_time_done = 0; s_decide_done = 0;
        const bool exists0 = nin0 <= nsamp_rel;
        desc[0].c = 0; desc[0].nin = nin0; desc[1].c = nin0;
        s_published = exists0 ? 1 : 0; s_stop = exists0 ? 0 : 1;
        for (int i = 0; i < DB_COUNT; ++i) mbar_init(&s_bell[i], 1);
    }
    __syncthreads();

    /* Warp roles.  SM sub-partition = warp id % 4: the oscillator shares its sub-partition only with the two
       light roles (TIME, DEC).  CTAs that share an SM with another CTA (bid >= #SMs) rotate by one warp so the
       two oscillators sit on different sub-partitions. */
    const int rw = (warp + 12 - (blockIdx.x >= 148 ? 1 : 0)) % 12;
    int role, ridx;                                                     /* ridx: index inside the role's group */
    switch (rw) {      /* sub-partition 0: OSC, DEC (light), one idle warp | 1: TIME, CONS, CONS | 2: CONS, EST, EST | 3: CONS, EST, EST */
        case 0: role = ROLE_OSC; ridx = 0; break;
        case 4: role = ROLE_DEC; ridx = 0; break;
#if PIPE_VARIANT == 1
        case 8: role = ROLE_EST; ridx = 4; break;
#else
        case 8: role = ROLE_IDLE; ridx = 0; break;
#endif
        case 1: role = ROLE_TIME; ridx = 0; break;
        case 5: role = ROLE_CONS; ridx = 0; break;
        case 9: role = ROLE_CONS; ridx = 1; break;
        case 2: role = ROLE_CONS; ridx = 2; break;
#if PIPE_VARIANT == 2
        case 3: role = ROLE_EST; ridx = 4; break;
#else
        case 3: role = ROLE_CONS; ridx = 3; break;
#endif
        case 6: role = ROLE_EST; ridx = 0; break;
        case 10: role = ROLE_EST; ridx = 1; break;
        case 7: role = ROLE_EST; ridx = 2; break;
        default: role = ROLE_EST; ridx = 3; break;     /* rw == 11 */
    }
    unsigned char *outb = (unsigned char *)p.out + (long long)c * p.out_stride * (p.soft ? 4 : 1);

    if (role == ROLE_OSC) {
        /* =========================== OSC ======================================================== */
        float2 ph = lane < M ? s_phi0[lane] : make_float2(1.0f, 0.0f);
#if PIPE_GRING
        float2 *myring = p.g_ring + ((size_t)c * M + (lane < M ? lane : 0)) * PIPE_R;
#else
        float2 *myring = ring + (lane < M ? lane : 0) * PIPE_RS;
#endif
        int pos = 0;
        unsigned par_est = 0, par_free = 0, par_pub = 0;
        PDECL
        for (int k = 0;; ++k) {
            { PTIC const bool ok_ = warp_wait_ge(&s_est_done, k + 1, &s_stop, lane, &s_bell[DB_EST], par_est); PTOC(0); if (!ok_) break; }
            const float2 d = desc[k & 3].dphi[lane < M ? lane : 0];
            int done = 0, target = nin_min, nin = -1;
            bool stopped = false;
            while (true) {
                while (done < target) {
                    const int n = min(PIPE_OSC_CHUNK - (pos & (PIPE_OSC_CHUNK - 1)), target - done);
                    int ok = 1;
                    PTIC
                    if (lane == 0) {                                     /* ring space: slot of pos+n-1 must be free */
                        while (pos + n - ld_acq(&s_free_pos) > PIPE_R) { if (ld_acq(&s_stop)) { ok = 0; break; } if (mbar_try_wait(&s_bell[DB_FREE], par_free, 1000)) par_free ^= 1u; }
                    }
                    ok = __shfl_sync(0xffffffffu, ok, 0);
                    PTOC(1);
                    if (!ok) { stopped = true; break; }
                    { PTIC
                    if (lane < M) {
                        float2 *dst = myring + (pos & (PIPE_R - 1));
                        int j = 0;
                        if (pos & 1) { ph = cmul_rn(ph, d); dst[0] = ph; j = 1; }          /* align to a 16-byte pair */
                        const int npair = (n - j) & ~1;
                        ph = osc_pairs(ph, d, dst + j, npair, OSCP);
                        j += npair;
                        if (j < n) { ph = cmul_rn(ph, d); dst[j] = ph; }
                    }
                    __syncwarp();
                    PTOC(3); }
                    pos += n; done += n;
                    if (lane == 0) { st_rel(&s_osc_pos, pos); mbar_ring(&s_bell[DB_OSC]); }
                }
                if (stopped || nin >= 0) break;
                { PTIC const bool ok_ = warp_wait_ge(&s_published, k + 1, &s_stop, lane, &s_bell[DB_PUB], par_pub); PTOC(2); if (!ok_) { stopped = true; break; } }
                nin = desc[k & 3].nin; target = nin;
            }
            if (stopped) break;
            if (lane < M) {                                             /* comp_normalize, comp_prim.h:133-139 */
                const float av = __fsqrt_rn(mag2_rn(ph));
                ph = make_float2(__fdiv_rn(ph.x, av), __fdiv_rn(ph.y, av));
                desc[k & 3].phi_end[lane] = ph;
            }
            __syncwarp();
        }
#ifdef FSK_PROFILE
        pacc[2] = pacc[3];   /* report pure recurrence cycles in the 'wait publish' slot */
#endif
        PFLUSH(0, 3);
    } else if (role == ROLE_TIME) {
        /* =========================== TIME: fine timing, publish next frame ========================== */
        float nrt_state = s_nrt_state, ppm = s_ppm_state;
        int k = 0;
        bool exists = ld_acq(&s_published) >= 1;
        unsigned par_fint = 0;
        PDECL
        while (exists) {
            const int ck = desc[k & 3].c, nin = desc[k & 3].nin;
            const float2 *fint = fintb + (size_t)(k & 1) * M * nfi;
            float acc = 0.0f;
            float2 pf0 = __ldg(p.phi_ft + min(lane, nfi - 1)), pf1 = __ldg(p.phi_ft + min(lane + 32, nfi - 1));
            for (int i0 = 0; i0 < nfi; i0 += PIPE_TB) {                  /* fsk.c:686-698, streamed */
                const int cnt = min(PIPE_TB, nfi - i0);
                { PTIC warp_wait_ge(&s_fint_total, k * nfi + i0 + cnt, &s_stop, lane, &s_bell[DB_FINT], par_fint); PTOC(0); }
                const float2 q0 = pf0, q1 = pf1;
                if (i0 + PIPE_TB < nfi) {                                /* prefetch the next batch's table entries */
                    pf0 = __ldg(p.phi_ft + min(i0 + PIPE_TB + lane, nfi - 1));
                    pf1 = __ldg(p.phi_ft + min(i0 + PIPE_TB + lane + 32, nfi - 1));
                }
                if (lane < cnt) {
                    float ft1 = 0.0f;
                    for (int m = 0; m < M; ++m) ft1 = __fadd_rn(ft1, mag2_rn(fint[m * nfi + i0 + lane]));
                    tre[lane] = __fmul_rn(ft1, q0.x); tim[lane] = __fmul_rn(ft1, q0.y);
                }
                if (lane + 32 < cnt) {
                    float ft1 = 0.0f;
                    for (int m = 0; m < M; ++m) ft1 = __fadd_rn(ft1, mag2_rn(fint[m * nfi + i0 + lane + 32]));
                    tre[lane + 32] = __fmul_rn(ft1, q1.x); tim[lane + 32] = __fmul_rn(ft1, q1.y);
                }
                __syncwarp();
                if (lane < 2) {                                          /* lane 0: real parts, lane 1: imaginary parts, in order */
                    const float *t = lane ? tim : tre;
                    int j = 0;
#pragma unroll 4
                    for (; j + 4 <= cnt; j += 4) {
                        const float4 v = *reinterpret_cast<const float4 *>(t + j);
                        acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
                    }
                    for (; j < cnt; ++j) acc = __fadd_rn(acc, t[j]);
                }
                __syncwarp();
            }
            const float tr = __shfl_sync(0xffffffffu, acc, 0), ti = __shfl_sync(0xffffffffu, acc, 1);
            int nan = 0, nin_next = nin, low = 0, high = s_high_last; float fract = 0.0f;
            if (isnan(tr) || isnan(ti)) nan = 1;                         /* fsk.c:702-704 */
            else {
                const float nrt = (float)__ddiv_rn((double)em_atan2f(ti, tr), 6.283185307179586476925286766559);
                const float rx_timing = __fmul_rn(nrt, (float)P);
                const float d = __fsub_rn(nrt, nrt_state);
                nrt_state = nrt;
                if (fabsf(d) < 0.2f) {
                    const float appm = (float)__ddiv_rn(__dmul_rn(1e6, (double)d), (double)(float)nsym);
                    ppm = (float)__dadd_rn(__dmul_rn(.9, (double)ppm), __dmul_rn(.1, (double)appm));
                }
                if (p.lock_nin) nin_next = nin;
                else if (nrt > 0.25f) nin_next = p.N + Ts / 2;
                else if (nrt < -0.25f) nin_next = p.N - Ts / 2;
                else nin_next = p.N;
                low = (int)floorf(rx_timing); fract = __fsub_rn(rx_timing, (float)low); high = (int)ceilf(rx_timing);
            }
            /* publish frame k+1 as early as possible: EST(k+2) and the end of OSC(k+1) hang on it */
            const int c_next = ck + nin;
            const bool exists_next = (p.max_frames <= 0 || k + 1 < p.max_frames) && (c_next + nin_next <= nsamp_rel);
            if (lane == 0) {
                s_nin_state = nin_next; s_nrt_state = nrt_state; s_ppm_state = ppm; s_high_last = high; s_nan_last = nan;
                if (exists_next) { desc[(k + 1) & 3].nin = nin_next; desc[(k + 2) & 3].c = c_next + nin_next; st_rel(&s_published, k + 2); mbar_ring(&s_bell[DB_PUB]); }
                PipeDesc &dk = desc[k & 3];
                dk.nan = nan; dk.low = low; dk.high = high; dk.nin_next = nin_next; dk.exists_next = exists_next ? 1 : 0;
                dk.fract = fract; dk.nrt = nrt_state; dk.ppm = ppm;
                st_rel(&s_time_done, k + 1); mbar_ring(&s_bell[DB_TIME]);
            }
            __syncwarp();
            exists = exists_next; ++k;
        }
        PFLUSH(7, 1);
    } else if (role == ROLE_DEC) {
        /* =========================== DEC: decisions, outputs, EbNo sums, frame record ================ */
        unsigned par_time = 0;
        for (int k = 0;; ++k) {
            if (!warp_wait_ge(&s_time_done, k + 1, &s_stop, lane, &s_bell[DB_TIME], par_time, 2000)) break;
            const PipeDesc &dk = desc[k & 3];
            const int nan = dk.nan, low = dk.low, high = dk.high, nin = dk.nin, exists_next = dk.exists_next;
            const float fract = dk.fract;
            const float2 *fint = fintb + (size_t)(k & 1) * M * nfi;
            const long long ofs = (long long)(fbase + k) * p.Nbits;
            if (!nan) {                                                  /* fsk.c:751-817 */
                const float omf = __fsub_rn(1.0f, fract);
                for (int i = lane; i < nsym; i += 32) {
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
                        else {
                            float a = -tmax[0], b = -tmax[0];
                            a = __fadd_rn(a, tmax[1]);  b = __fadd_rn(b, -tmax[1]);
                            a = __fadd_rn(a, -tmax[2]); b = __fadd_rn(b, tmax[2]);
                            a = __fadd_rn(a, tmax[3]);  b = __fadd_rn(b, tmax[3]);
                            o[2 * i + 1] = p.softinv ? -a : a; o[2 * i] = p.softinv ? -b : b;
                        }
                    }
                }
            } else {                                                     /* stale frame, fsk_demod.c:424-433 */
                for (int i = lane; i < p.Nbits; i += 32) {
                    if (p.soft) {
                        float *o = (float *)outb;
                        const float prev = (fbase + k) > 0 ? o[ofs - p.Nbits + i] : __uint_as_float(p.last_out[(long long)c * p.Nbits + i]);
                        o[ofs + i] = p.softinv ? -prev : prev;
                    } else {
                        outb[ofs + i] = (fbase + k) > 0 ? outb[ofs - p.Nbits + i] : (unsigned char)p.last_out[(long long)c * p.Nbits + i];
                    }
                }
            }
            __syncwarp();
            float es = 0.0f;                                             /* EbNo sums (fsk.c:808-815), sequential */
            if (lane < 2 && !nan) {
                const float *t = (const float *)s_maxs + lane;
#pragma unroll 8
                for (int i = 0; i < nsym; ++i) es = __fadd_rn(es, t[2 * i]);
            }
            const float mean_sum = __shfl_sync(0xffffffffu, es, 1);
            if (lane == 0) {
                if (p.write_recs) {
                    FskFrameRec r;
                    r.nin_used = nin; r.nin_next = dk.nin_next; r.norm_rx_timing = dk.nrt; r.ppm = dk.ppm;
                    r.mean_sum = mean_sum; r.std_sum = es; r.nanflag = nan; r.high_sample = high;
                    for (int m = 0; m < 4; ++m) r.tabidx[m] = m < M ? dk.tabidx[m] : 0;
                    p.recs[(long long)c * p.rec_stride + fbase + k] = r;
                }
                s_frames = k + 1;
                __threadfence_block();
                st_rel(&s_decide_done, k + 1); mbar_ring(&s_bell[DB_DEC]);
                if (!exists_next) { st_rel(&s_stop, 1); for (int i = 0; i < DB_COUNT; ++i) mbar_ring(&s_bell[i]); }   /* wake every sleeper */
            }
            __syncwarp();
            if (!exists_next) break;
        }
    } else if (role == ROLE_CONS) {
        /* =========================== CONS: down-convert + integrate, streaming ===================== */
        const int ct = ridx * 32 + lane, NCT = PIPE_NCONS * 32;
        unsigned par_pub = 0, par_dec = 0, par_osc = 0;
        PDECL
        for (int k = 0;; ++k) {
            { PTIC
              /* only the group's first warp polls; the others block in the named barrier (no issue slots spent) */
              if (ridx == 0) {
                  const bool ok_ = warp_wait_ge(&s_published, k + 1, &s_stop, lane, &s_bell[DB_PUB], par_pub, 2000) && warp_wait_ge(&s_decide_done, k - 1, &s_stop, lane, &s_bell[DB_DEC], par_dec, 2000);
                  if (lane == 0) *(volatile int *)&s_ok_cons = ok_ ? 1 : 0;
              }
              group_bar(2, NCT);
              const bool ok_ = *(volatile int *)&s_ok_cons != 0;
              PTOC(0); if (!ok_) break; }
            const int ck = desc[k & 3].c, nin = desc[k & 3].nin, base = ck + nin - Nmem, fend = ck + nin;
            float2 *fint = fintb + (size_t)(k & 1) * M * nfi;
            int i_done = 0;
            float2 xn = make_float2(0.0f, 0.0f);
            if (ct < min(PIPE_BLK, fend - ck)) xn = load_sample_t<FMT>(p.in, gbase + ck + ct);
#if PIPE_GRING
            float2 phn[M]; bool phv = false;                         /* this thread's phases of the next block, fetched early when the oscillator is ahead */
#endif
            for (int s0 = ck; s0 < fend; s0 += PIPE_BLK) {
                const int cnt = min(PIPE_BLK, fend - s0);
                const float2 xv = xn;
                if (s0 + PIPE_BLK + ct < fend) xn = load_sample_t<FMT>(p.in, gbase + s0 + PIPE_BLK + ct);   /* prefetch next block */
                { PTIC
                  if (ridx == 0) warp_wait_ge(&s_osc_pos, s0 + cnt, &s_stop, lane, &s_bell[DB_OSC], par_osc);   /* a published frame is always completed by OSC */
                  group_bar(2, NCT);
                  PTOC(1); }
                if (ct < cnt) {
                    const int slot = (s0 + ct) & (PIPE_R - 1);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
#if PIPE_GRING
                        const float2 ph = phv ? phn[m] : __ldcg(p.g_ring + ((size_t)c * M + m) * PIPE_R + slot);
#else
                        const float2 ph = ring[m * PIPE_RS + slot];
#endif
                        ring[m * PIPE_RS + slot] = cmul_rn(xv, make_float2(ph.x, -ph.y));       /* fsk.c:648 */
                    }
                }
#if PIPE_GRING
                {   /* the next block's phases, if the oscillator has already published them (it usually runs ahead): hides the L2 round trip */
                    const int s1 = s0 + PIPE_BLK;
                    phv = false;
                    if (s1 + ct < fend && ld_acq(&s_osc_pos) >= min(s1 + PIPE_BLK, fend)) {
                        const int slot1 = (s1 + ct) & (PIPE_R - 1);
#pragma unroll
                        for (int m = 0; m < M; ++m) phn[m] = __ldcg(p.g_ring + ((size_t)c * M + m) * PIPE_R + slot1);
                        phv = true;
                    }
                }
#endif
                group_bar(2, NCT);
                const int mixed_end = s0 + cnt;
                int i_hi = 0;
                if (mixed_end - base >= Ts) i_hi = min(nfi, (mixed_end - base - Ts) / step + 1);
                const int ni = i_hi - i_done;
                /* fsk.c:660-668.  Threads are grouped by tone (a warp reads one ring row at consecutive window starts: no bank
                   conflicts); with windows every 2nd sample the Ts sequential adds take their operands as 16-byte pairs. */
                const bool pair_path = (step == 2) && !(Ts & 1);
                const int par = base & 1;
                for (int e = ct; e < (ni << lgM); e += NCT) {
                    int m, ii;
                    if (M == 2) { m = e >= ni ? 1 : 0; ii = e - m * ni; } else { m = e / ni; ii = e - m * ni; }
                    const int i = i_done + ii, a0 = base + i * step;
                    const float2 *rg = ring + m * PIPE_RS;
                    float2 acc = make_float2(0.0f, 0.0f);
                    if (pair_path && !par) {
#pragma unroll 5
                        for (int j = 0; j < Ts; j += 2) {
                            const float4 v = *reinterpret_cast<const float4 *>(rg + ((a0 + j) & (PIPE_R - 1)));
                            acc = cadd_rn(acc, make_float2(v.x, v.y)); acc = cadd_rn(acc, make_float2(v.z, v.w));
                        }
                    } else if (pair_path) {
                        { const float4 v = *reinterpret_cast<const float4 *>(rg + ((a0 - 1) & (PIPE_R - 1))); acc = cadd_rn(acc, make_float2(v.z, v.w)); }
#pragma unroll 4
                        for (int j = 1; j < Ts - 1; j += 2) {
                            const float4 v = *reinterpret_cast<const float4 *>(rg + ((a0 + j) & (PIPE_R - 1)));
                            acc = cadd_rn(acc, make_float2(v.x, v.y)); acc = cadd_rn(acc, make_float2(v.z, v.w));
                        }
                        { const float4 v = *reinterpret_cast<const float4 *>(rg + ((a0 + Ts - 1) & (PIPE_R - 1))); acc = cadd_rn(acc, make_float2(v.x, v.y)); }
                    } else {
#pragma unroll 5
                        for (int j = 0; j < Ts; ++j) acc = cadd_rn(acc, rg[(a0 + j) & (PIPE_R - 1)]);
                    }
                    fint[m * nfi + i] = acc;
                }
                group_bar(2, NCT);
                if (ct == 0) {
                    st_rel(&s_fint_total, k * nfi + i_hi); mbar_ring(&s_bell[DB_FINT]);
                    st_rel(&s_free_pos, min(base + i_hi * step, fend - p.tailcap)); mbar_ring(&s_bell[DB_FREE]);
                }
                i_done = i_hi;
            }
        }
        if (ridx == 0) PFLUSH(4, 2);
    } else if (role == ROLE_EST) {
        /* =========================== EST: FFTs -> Sf -> tone estimate, one frame ahead =============== */
        const int ew = ridx;
        const int et = ew * 32 + lane, NET = PIPE_NEST * 32;
        const int numffts = (int)floorf(__fdiv_rn((float)nin_min, (float)(Ndft / 2))) - 1;   /* same for all 3 nin (host-checked) */
        const bool fused128 = (Ndft == 128);
        F128Lane L128;
        if (fused128) f128_init(L128, s_hann, lane);
        const int mstride = fused128 ? F128_MSTRIDE : Ndft;
        float2 *F = fftbuf + ew * (fused128 ? 4 * F128_STRIDE : PIPE_NFB * (Ndft + (Ndft >> 3)));       /* this warp's FFT buffers */
        unsigned par_pub = 0;
        PDECL
        for (int k = 0;; ++k) {
            { PTIC
              if (ew == 0) { const bool ok_ = warp_wait_ge(&s_published, k, &s_stop, lane, &s_bell[DB_PUB], par_pub, 2000); if (lane == 0) *(volatile int *)&s_ok_est = ok_ ? 1 : 0; }
              group_bar(1, NET);
              const bool ok_ = *(volatile int *)&s_ok_est != 0;
              PTOC(0); if (!ok_) break; }
            const int ck = desc[k & 3].c;
            if (!(ck + nin_min <= nsamp_rel && (p.max_frames <= 0 || k < p.max_frames))) break;   /* frame k cannot exist */
            const float *Sf_in = Sfbuf + ((k + 1) & 1) * Ndft;
            float *Sf_out = Sfbuf + (k & 1) * Ndft;
            if (fused128) {
                /* Ndft = 128: four windows per round in two register passes (fsk_fft128.cuh).  A super-round = one round per EST warp
                   (16 windows); their |X| rows go to magbuf, then the ordered Sf recursion runs over them (fsk.c:500) */
                for (int r0 = 0; 4 * r0 < numffts; r0 += PIPE_NEST) {
                    { PTIC
                    if (4 * (r0 + ew) < numffts)
                        est_round128<FMT, true>(p, gbase + ck, 4 * (r0 + ew), numffts, F, s_tw, L128, nullptr, nullptr, lane, magbuf - 4 * r0 * F128_MSTRIDE);
                    PTOC(2); }
                    { PTIC
                    group_bar(1, NET);
                    const int jend = min(numffts, 4 * (r0 + PIPE_NEST));
                    const float *Sf_src = r0 ? Sf_out : Sf_in;
                    for (int i = et; i < Ndft; i += NET) {
                        float s = Sf_src[i];
                        const int src = (i + Ndft / 2) & (Ndft - 1);
                        for (int jj = 4 * r0; jj < jend; ++jj)
                            s = __fadd_rn(__fmul_rn(s, p.omtc), __fmul_rn(magbuf[(jj - 4 * r0) * F128_MSTRIDE + src], p.tc));
                        Sf_out[i] = s;
                    }
                    group_bar(1, NET);
                    PTOC(3); }
                }
            } else
            /* this warp's windows: j = ew, ew+NEST, ... taken PIPE_NFB at a time so that the loads, butterflies and
               magnitudes of independent FFTs overlap (fsk.c:466-503) */
            for (int j0 = ew; j0 < numffts; j0 += PIPE_NEST * PIPE_NFB) {
                float2 *Fq[PIPE_NFB];
                { PTIC
#pragma unroll
                for (int g = 0; g < PIPE_NFB; ++g) Fq[g] = F + g * (Ndft + (Ndft >> 3));
#pragma unroll 1
                for (int ib = 0; ib < Ndft; ib += 128) {                /* 128 samples of each window per pass: 4 loads per lane per FFT in flight */
                    float2 v[PIPE_NFB][4];
#pragma unroll
                    for (int g = 0; g < PIPE_NFB; ++g) {
                        const int j = j0 + g * PIPE_NEST;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int i = ib + q * 32 + lane;
                            v[g][q] = (j < numffts && i < Ndft) ? load_sample_t<FMT>(p.in, gbase + ck + j * (Ndft / 2) + i) : make_float2(0.0f, 0.0f);
                        }
                    }
#pragma unroll
                    for (int g = 0; g < PIPE_NFB; ++g) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {                    /* fsk.c:470-478: window, scatter to DIT order */
                            const int i = ib + q * 32 + lane;
                            if (i < Ndft) { const float h = s_hann[i]; Fq[g][fpad<true>(s_iperm[i])] = make_float2(__fmul_rn(h, v[g][q].x), __fmul_rn(h, v[g][q].y)); }
                        }
                    }
                }
                __syncwarp();
                PTOC(1); }
                { PTIC
                warp_fft_multi<true, PIPE_NFB>(Fq, s_tw, p, lane);
#pragma unroll
                for (int g = 0; g < PIPE_NFB; ++g) {
                    const int j = j0 + g * PIPE_NEST;
                    if (j < numffts)
                        for (int i = lane; i < Ndft; i += 32) magbuf[j * Ndft + i] = __fsqrt_rn(mag2_rn(Fq[g][fpad<true>(i)]));
                }
                __syncwarp();
                PTOC(2); }
            }
            if (!fused128) { PTIC
            group_bar(1, NET);
            for (int i = et; i < Ndft; i += NET) {                       /* Sf IIR over the frame's FFTs, in order (fsk.c:500) */
                float s = Sf_in[i];
                const int src = (i + Ndft / 2) & (Ndft - 1);
#pragma unroll 4
                for (int jj = 0; jj < numffts; ++jj)
                    s = __fadd_rn(__fmul_rn(s, p.omtc), __fmul_rn(magbuf[jj * mstride + src], p.tc));
                Sf_out[i] = s;
            }
            group_bar(1, NET);
            PTOC(3); }
            if (ew == 0) {
                int tabidx[4];
                estimate_tones(p, Sf_out, scr, s_mask, lane, tabidx);
                if (lane < M) { const int ti = tabidx[lane]; desc[k & 3].tabidx[lane] = ti; desc[k & 3].dphi[lane] = __ldg(p.dphi_tab + ti); }
                __syncwarp();
                if (lane == 0) { st_rel(&s_est_done, k + 1); mbar_ring(&s_bell[DB_EST]); }
            }
        }
        if (ew == 0) PFLUSH(11, 4);
    }
    __syncthreads();

    /* ---- epilogue: state back to HBM ---- */
    const int frames = s_frames;
    const int c_end = frames > 0 ? desc[frames & 3].c : 0;             /* desc[k+1].c is written when frame k is published */
    const float *Sf_fin = Sfbuf + ((frames + 1) & 1) * Ndft;            /* buffer (frames-1)&1 */
    const float2 *fint = fintb + (size_t)((frames + 1) & 1) * M * nfi;  /* f_int of the last frame */
    for (int i = tid; i < Ndft; i += PIPE_THREADS) p.Sf[(long long)c * Ndft + i] = Sf_fin[i];
    for (int e = tid; e < M * p.tailcap; e += PIPE_THREADS) {
        const int m = e / p.tailcap, i = e - m * p.tailcap;
        p.fdc_tail[((long long)c * M + m) * p.tailcap + i] = ring[m * PIPE_RS + ((c_end - p.tailcap + i) & (PIPE_R - 1))];
    }
    if (tid < M) p.phi_c[c * 4 + tid] = frames > 0 ? desc[(frames - 1) & 3].phi_end[tid] : s_phi0[tid];
    if (tid == 0) {
        p.nin[c] = s_nin_state; p.nrt[c] = s_nrt_state; p.ppm[c] = s_ppm_state;
        p.nframes[c] = fbase + frames; p.consumed[c] = c_end;
        p.tot_frames[c] += frames; p.tot_samples[c] += c_end;
        if (p.frame_base) p.frame_base[c] = fbase + frames;
    }
    if (frames > 0) {
        const long long ofs = (long long)(fbase + frames - 1) * p.Nbits;
        for (int i = tid; i < p.Nbits; i += PIPE_THREADS)
            p.last_out[(long long)c * p.Nbits + i] = p.soft ? __float_as_uint(((const float *)outb)[ofs + i]) : (unsigned int)outb[ofs + i];
        if (p.write_recs && !s_nan_last) {
            const int neyeoffset = s_high_last + 1, tot = p.eye_traces * M * p.neyesamp;
            for (int e = tid; e < tot; e += PIPE_THREADS) {
                const int row = e / p.neyesamp, j = e - row * p.neyesamp, i = row / M, m = row - i * M;
                const int ind = 2 * P * i + neyeoffset + j * p.eye_dec, flat = m * nfi + ind;
                float2 v = make_float2(0.0f, 0.0f);
                if (flat >= 0 && ind < nfi) v = fint[flat];
                p.eye_raw[(long long)c * (8 * 160) + row * 160 + j] = __fsqrt_rn(mag2_rn(v));
            }
        }
    }
    if (p.carry_cap > 0) {                                             /* same tail carry as the serial kernel */
        const long long cursor = start + c_end, tail = p.nsamp - cursor;
        constexpr int es = FMT == 1 ? 8 : (FMT == 0 ? 4 : 2);
        const long long nhalf = tail * es / 2;
        const int chunk = (int)(p.q_stage_bytes / 2);
        __syncthreads();
        unsigned short *stage = (unsigned short *)(smem + p.q_fint);    /* f_int is dead now (eye already written) */
        const long long base = (long long)c * p.in_stride;
        const unsigned short *src = (const unsigned short *)((const unsigned char *)p.in + (base + cursor) * es);
        unsigned short *dst = (unsigned short *)((unsigned char *)p.carry_dst + (base + p.carry_cap - tail) * es);
        for (long long k0 = 0; k0 < nhalf; k0 += chunk) {
            const int n = (int)min((long long)chunk, nhalf - k0);
            __syncthreads();
            for (int i = tid; i < n; i += PIPE_THREADS) stage[i] = src[k0 + i];
            __syncthreads();
            for (int i = tid; i < n; i += PIPE_THREADS) dst[k0 + i] = stage[i];
        }
        if (tid == 0) p.in_off_next[c] = (int)(p.carry_cap - tail);
    }
}
