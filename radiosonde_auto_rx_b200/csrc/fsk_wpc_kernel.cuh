/* fsk_wpc_kernel.cuh -- the THROUGHPUT form of the batched fsk_demod kernel (sm_100a): one CTA owns up to 16 channels.
 *
 * Why a second kernel.  fsk_pipe_kernel.cuh spends a 12-warp CTA on ONE channel so that the channel's oscillator recurrence
 * (utils/fsk.c:646-647) is the only thing on its critical path: right for <= 2 channels per SM, but every SM then hosts two
 * oscillator warps that use 2 of 32 lanes (46 % of all useful warp-instructions) and ten warps that poll each other.  With many
 * channels per GPU the per-channel latency is irrelevant and the instruction count per sample is everything.  Here:
 *
 *   warp 0        OSC   lane-per-chain oscillator: lane = channel*M + tone, all chains of the CTA advance in lockstep over the
 *                       absolute sample position; 6 FP32 instructions + one 16-byte shared store per two steps serve up to 32
 *                       chains.  Frame boundaries (normalise, next dphi, next nin: fsk.c:644-651, 724-731) are per-lane events.
 *   warp 1+g      one warp per channel runs the whole rest of fsk_demod_core for its channel SEQUENTIALLY -- no role groups,
 *                       no named barriers, no polling between roles:
 *                         EST(k+1)  windowed FFTs -> |X| -> Sf IIR -> tone estimate -> dphi LUT          fsk.c:438-590
 *                         CONS(k)   x*conj(phi) in place in the channel's ring, integrate-and-dump         fsk.c:648,659-668
 *                                   + timing terms and their sequential accumulation                      fsk.c:682-698
 *                         TIME(k)   exact atan2f, ppm, next nin                                           fsk.c:700-731
 *                         DEC(k)    re-sample, decide, soft/hard output, EbNo sums, frame record          fsk.c:737-817
 *                       The warp consumes oscillator output in blocks as soon as it exists and fills the gaps with EST rounds of
 *                       the next frame (PIPE-style one-frame look-ahead: c_{k+1} = c_k + nin_k is known when frame k starts).
 *
 * f_int ((Nsym+1)*P*M complex per frame, 24 KB at the RS41 production line) cannot stay in shared memory for 16 channels: it
 * lives in a per-channel global buffer that the same warp re-reads from L2 for the decisions.
 * A CTA depends on nothing outside itself, so a batch larger than the resident set simply runs in waves.
 * Arithmetic, operand order and summation order are those of fsk_kernel.cuh / fsk_pipe_kernel.cuh, hence of the reference;
 * the same bit-exact GPU tests run against all three kernels (B200FSK_KERNEL=serial|pipe|wpc).
 */
#pragma once
#include "fsk_pipe_kernel.cuh"
#include "fsk_fft128.cuh"

#ifndef WPC_MAXG
#define WPC_MAXG   15           /* channel warps per CTA: 16 warps = 512 threads leave 128 registers per thread */
#endif
#ifndef WPC_NFB
#define WPC_NFB    4            /* FFTs a channel warp advances together */
#endif
#ifndef WPC_CTAS
#define WPC_CTAS   1            /* CTAs per SM the kernel is sized for (registers, shared memory) */
#endif
#define WPC_CHUNK  64           /* oscillator steps between position publications */
#define WPC_SPIN_LIMIT (1LL << 33)   /* watchdog: cycles a wait may last before the CTA gives up (~4 s) */

struct WpcCtl {                 /* static shared control block */
    int   oscpos;               /* oscillator output exists for positions < oscpos (all chains) */
    int   abort;
    int   free_[WPC_MAXG];      /* per channel: oldest ring position its warp still needs */
    int   seqE[WPC_MAXG];       /* est descriptor k is valid when seqE >= k+1 */
    int   seqT[WPC_MAXG];       /* nin descriptor k is valid when seqT >= k+1 */
    int   ninv[WPC_MAXG][2];    /* nin of frame k (slot k&1), or -1: the frame does not exist */
    int   nsrel[WPC_MAXG];      /* samples available to the channel in this launch */
    float2 dphi[WPC_MAXG][2][4];
};

__device__ __forceinline__ int ld_vol(const int *p) { return *(const volatile int *)p; }

/* spin until *flag >= want (acquire); false when the CTA aborted or the watchdog fired */
__device__ __forceinline__ bool wpc_wait_ge(const int *flag, int want, WpcCtl *ctl, unsigned ns)
{
    if (ld_acq(flag) >= want) return true;
    const long long t0 = clock64();
    for (unsigned it = 1;; ++it) {
        __nanosleep(ns);
        if (ld_acq(flag) >= want) return true;
        if ((it & 31u) == 0) {                                       /* watchdog and abort only now and then: the loop must stay cheap */
            if (ld_vol(&ctl->abort)) return false;
            if (clock64() - t0 > WPC_SPIN_LIMIT) { *(volatile int *)&ctl->abort = 1; return false; }
        }
    }
}

/* the same with exponential back-off (ns0, 2 ns0, ... up to nsmax) for waits that are known to be long and off the critical path:
 * the ncu capture of the duo kernel counted more than half of all issued warp instructions inside fixed-interval poll loops */
__device__ __forceinline__ bool wpc_wait_ge_bo(const int *flag, int want, WpcCtl *ctl, unsigned ns0, unsigned nsmax)
{
    if (ld_acq(flag) >= want) return true;
    const long long t0 = clock64();
    unsigned ns = ns0;
    for (unsigned it = 1;; ++it) {
        __nanosleep(ns);
        if (ld_acq(flag) >= want) return true;
#if !defined(DUO_BACKOFF) || DUO_BACKOFF
        ns = min(ns * 2u, nsmax);
#endif
        if ((it & 7u) == 0) {
            if (ld_vol(&ctl->abort)) return false;
            if (clock64() - t0 > WPC_SPIN_LIMIT) { *(volatile int *)&ctl->abort = 1; return false; }
        }
    }
}

template <int M, int FMT>
__global__ void __launch_bounds__((WPC_MAXG + 1) * 32, WPC_CTAS) fsk_wpc_kernel(const FskKParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ WpcCtl ctl;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = p.w_G;                                             /* channels per CTA */
    const int Ndft = p.Ndft, Nmem = p.Nmem, nfi = p.nfi, P = p.P, Ts = p.Ts, nsym = p.Nsym;
    const int step = Ts / P, nin_min = p.N - Ts / 2;
    const int RN = p.w_ring, RS = RN + 2, RMASK = RN - 1, BLK = p.w_blk;
    constexpr int lgM = (M == 2) ? 1 : 2;

    float  *s_hann = (float *)(smem + p.w_hann);
    float2 *s_tw   = (float2 *)(smem + p.w_tw);
    unsigned short *s_iperm = (unsigned short *)(smem + p.w_perm);
    float  *s_mask = (float *)(smem + p.w_mask);

    /* ---- CTA prologue ---- */
    for (int i = tid; i < Ndft; i += blockDim.x) { s_hann[i] = p.hann[i]; s_tw[i] = p.tw[i]; s_iperm[i] = p.invperm[i]; s_mask[i] = p.mask[i]; }
    if (tid == 0) { ctl.oscpos = 0; ctl.abort = 0; }
    if (tid < WPC_MAXG) { ctl.free_[tid] = -p.tailcap; ctl.seqE[tid] = 0; ctl.seqT[tid] = 0; ctl.nsrel[tid] = -1; }
    __syncthreads();

    if (warp == 0) {
        /* =========================== OSC: lane-per-chain oscillator ============================== */
        const int g = lane >> lgM, m = lane & (M - 1);
        const int c = blockIdx.x * G + g;
        const bool valid = g < G && c < p.nchan;
        /* the channel warps publish nsrel in their prologue */
        if (valid) wpc_wait_ge(&ctl.nsrel[g], 0, &ctl, 100);
        const int nsrel = valid ? ld_vol(&ctl.nsrel[g]) : 0;
        float2 *myring = (float2 *)(smem + p.w_chan0 + (size_t)(valid ? g : 0) * p.w_chan_stride + p.w_ring_off) + (size_t)m * RS;
        float2 ph = valid ? p.phi_c[c * 4 + m] : make_float2(1.0f, 0.0f);
        float2 d = make_float2(1.0f, 0.0f);
        int k = 0, cstart = 0, e_lim = 0;
        bool have_nin = false;
        bool act = valid && nin_min <= nsrel;                        /* frame 0 can exist (max_frames <= 0 or >= 1) */
        if (act) {
            if (wpc_wait_ge(&ctl.seqE[g], 1, &ctl, 200)) { d = ctl.dphi[g][0][m]; e_lim = nin_min; } else act = false;
        }
        int pos = 0;
        long long spin_t0 = 0;
        while (true) {
            /* the frame's real length, as soon as TIME(k-1) has published it (-1: the frame does not exist) */
            if (act && !have_nin && ld_acq(&ctl.seqT[g]) >= k + 1) {
                const int nin = ld_vol(&ctl.ninv[g][k & 1]);
                if (nin < 0) act = false; else { have_nin = true; e_lim = cstart + nin; }
            }
            if (!__any_sync(0xffffffffu, act)) break;
            const int lim = __reduce_min_sync(0xffffffffu, act ? e_lim - pos : (1 << 30));
            if (lim <= 0) {                                          /* per-lane frame events at this position */
                if (act && pos == e_lim) {
                    if (!have_nin) {                                 /* nin_min steps done and the length still unknown: wait for it */
                        if (!wpc_wait_ge(&ctl.seqT[g], k + 1, &ctl, 100)) act = false;
                    } else {                                         /* frame end: comp_normalize (comp_prim.h:133-139), next frame */
                        const float av = __fsqrt_rn(mag2_rn(ph));
                        ph = make_float2(__fdiv_rn(ph.x, av), __fdiv_rn(ph.y, av));
                        p.phi_c[c * 4 + m] = ph;                     /* state after the last frame that really ended */
                        ++k; cstart = e_lim; have_nin = false;
                        if (!(cstart + nin_min <= nsrel && (p.max_frames <= 0 || k < p.max_frames))) act = false;
                        else if (wpc_wait_ge(&ctl.seqE[g], k + 1, &ctl, 100)) { d = ctl.dphi[g][k & 1][m]; e_lim = cstart + nin_min; }
                        else act = false;
                    }
                }
                __syncwarp();
                continue;
            }
            int n = min(lim, WPC_CHUNK - (pos & (WPC_CHUNK - 1)));
            /* ring space: the slot of pos+n-1 must be free in every live channel */
            const int minfree = __reduce_min_sync(0xffffffffu, act ? ld_vol(&ctl.free_[g]) : (1 << 30));
            if (pos + n - minfree > RN) {
                const int fit = (RN - (pos - minfree)) & ~1;         /* take what fits rather than wait for the whole chunk */
                if (fit <= 0) {
                    if (!spin_t0) spin_t0 = clock64();
                    if (ld_vol(&ctl.abort)) break;
                    if (clock64() - spin_t0 > WPC_SPIN_LIMIT) { *(volatile int *)&ctl.abort = 1; break; }
                    __nanosleep(400);
                    continue;
                }
                n = min(n, fit);
            }
            spin_t0 = 0;
            if (act) {                                              /* n steps, fsk.c:646-647; paired 16-byte stores.  Only ACTIVE chains: a channel
                                                                       that has ended is not covered by the ring-space check any more, and its
                                                                       warp may still be consuming the last blocks of its last frame */
                float2 *dst = myring + (pos & RMASK);
                int j = 0;
                if (pos & 1) { ph = cmul_rn(ph, d); dst[0] = ph; j = 1; }
                const int npair = (n - j) & ~1;
                ph = osc_pairs(ph, d, dst + j, npair);
                j += npair;
                if (j < n) { ph = cmul_rn(ph, d); dst[j] = ph; }
            }
            __syncwarp();
            pos += n;
            if (lane == 0) st_rel(&ctl.oscpos, pos);
        }
        return;
    }

    /* =============================== channel warp ================================================== */
    const int g = warp - 1, c = blockIdx.x * G + g;
    if (g >= G || c >= p.nchan) return;
    unsigned char *cb = smem + p.w_chan0 + (size_t)g * p.w_chan_stride;
    float2 *ring  = (float2 *)(cb + p.w_ring_off);                   /* [M][RS]: phi from OSC, then f_dc in place */
    float2 *fftb  = (float2 *)(cb + p.w_fft_off);                    /* [WPC_NFB][Ndft + Ndft/8]; time-shared with tbuf / smax */
    float  *Sfbuf = (float *)(cb + p.w_Sf_off);                      /* [2][Ndft]: Sf after frame k lives in buffer k&1 */
    float  *scr   = (float *)fftb;                                   /* estimator scratch: runs when no FFT is in flight */
    float  *tre   = (float *)fftb, *tim = tre + p.w_tcap;   /* timing terms of one block, real / imaginary */
    float2 *s_maxs = fftb;                                           /* per symbol (max, sqrt(max)) */
    unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
    /* this channel's f_int in global memory (L2): one slot per (SM, channel warp), so the hot set is the resident set */
    float2 *fintg = p.w_fint + ((size_t)(WPC_CTAS == 1 ? smid % (unsigned)p.w_fint_slots : blockIdx.x) * WPC_MAXG + g) * M * nfi;

    for (int i = lane; i < Ndft; i += 32) Sfbuf[Ndft + i] = p.Sf[(long long)c * Ndft + i];      /* "after frame -1" -> buffer 1 */
    for (int e = lane; e < M * p.tailcap; e += 32) {                 /* integrator memory -> ring slots [-tailcap, 0) */
        const int m = e / p.tailcap, i = e - m * p.tailcap;
        ring[m * RS + ((i - p.tailcap) & RMASK)] = p.fdc_tail[((long long)c * M + m) * p.tailcap + i];
    }
    const long long start = p.in_off ? (long long)p.in_off[c] : 0;
    const long long gbase = (long long)c * p.in_stride + start;
    const int nsamp_rel = (int)(p.nsamp - start);
    const int fbase = p.frame_base ? p.frame_base[c] : 0;
    unsigned char *outb = (unsigned char *)p.out + (long long)c * p.out_stride * (p.soft ? 4 : 1);
    const int numffts = (int)floorf(__fdiv_rn((float)nin_min, (float)(Ndft / 2))) - 1;   /* same for all 3 nin (host-checked) */
    const int nrounds = (numffts + WPC_NFB - 1) / WPC_NFB;
    const int FP = p.w_fp;                                           /* float2 per FFT buffer */
    __syncwarp();

    int nin = p.nin[c];
    float nrt_state = p.nrt[c], ppm = p.ppm[c];
    int high_last = 0, nan_last = 0;
    bool exists = nin <= nsamp_rel;
    if (lane == 0) {
        ctl.ninv[g][0] = exists ? nin : -1;
        st_rel(&ctl.seqT[g], 1);
        st_rel(&ctl.nsrel[g], nsamp_rel < 0 ? 0 : nsamp_rel);
    }

    /* one EST round: windows [r*NFB, r*NFB+NFB) of the frame that starts at fc; Sf_in -> Sf_out (fsk.c:466-503) */
    F128Lane L128;
    const bool fused128 = (WPC_NFB == 4) && (Ndft == 128);          /* fsk_fft128.cuh: two register passes, conflict-free */
    if (fused128) f128_init(L128, s_hann, lane);
    auto est_round = [&](int fc, int r, const float *Sf_in, float *Sf_out) {
        const int j0 = r * WPC_NFB;
        if (fused128) { est_round128<FMT>(p, gbase + fc, j0, numffts, fftb, s_tw, L128, (r == 0) ? Sf_in : Sf_out, Sf_out, lane); return; }
        float2 *Fq[WPC_NFB];
#pragma unroll
        for (int q = 0; q < WPC_NFB; ++q) Fq[q] = fftb + q * FP;
#pragma unroll 1
        for (int ib = 0; ib < Ndft; ib += 128) {
            float2 v[WPC_NFB][4];
#pragma unroll
            for (int q = 0; q < WPC_NFB; ++q) {
                const int j = j0 + q;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = ib + u * 32 + lane;
                    v[q][u] = (j < numffts && i < Ndft) ? load_sample_t<FMT>(p.in, gbase + fc + j * (Ndft / 2) + i) : make_float2(0.0f, 0.0f);
                }
            }
#pragma unroll
            for (int q = 0; q < WPC_NFB; ++q) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {                        /* fsk.c:470-478: window, scatter to DIT order */
                    const int i = ib + u * 32 + lane;
                    if (i < Ndft) { const float h = s_hann[i]; Fq[q][fpad<true>(s_iperm[i])] = make_float2(__fmul_rn(h, v[q][u].x), __fmul_rn(h, v[q][u].y)); }
                }
            }
        }
        __syncwarp();
        warp_fft_multi<true, WPC_NFB>(Fq, s_tw, p, lane);
        const float *Sf_src = (r == 0) ? Sf_in : Sf_out;
        for (int i = lane; i < Ndft; i += 32) {                      /* |X| and the Sf IIR in window order (fsk.c:484-500) */
            float s = Sf_src[i];
            const int src = fpad<true>((i + Ndft / 2) & (Ndft - 1));
#pragma unroll
            for (int q = 0; q < WPC_NFB; ++q)
                if (j0 + q < numffts) s = __fadd_rn(__fmul_rn(s, p.omtc), __fmul_rn(__fsqrt_rn(mag2_rn(Fq[q][src])), p.tc));
            Sf_out[i] = s;
        }
        __syncwarp();
    };
    /* tone estimate of frame kk from its final Sf, descriptor to the oscillator */
    int tab_cur[4] = {0, 0, 0, 0}, tab_next[4] = {0, 0, 0, 0};
    auto est_finish = [&](int kk, const float *Sf_fin, int tab[4]) {
        estimate_tones(p, Sf_fin, scr, s_mask, lane, tab);
        if (lane < M) ctl.dphi[g][kk & 1][lane] = __ldg(p.dphi_tab + tab[lane]);
        __syncwarp();
        if (lane == 0) st_rel(&ctl.seqE[g], kk + 1);
    };

    /* frame 0's estimate cannot overlap anything */
    if (nin_min <= nsamp_rel) {
        for (int r = 0; r < nrounds; ++r) est_round(0, r, Sfbuf + Ndft, Sfbuf);
        est_finish(0, Sfbuf, tab_cur);
    }

    int k = 0, ck = 0, frames = 0;
    bool dead = false;
    while (exists && !dead) {
        const int ck1 = ck + nin, base = ck1 - Nmem, fend = ck1;
        const bool est_next = (ck1 + nin_min <= nsamp_rel) && (p.max_frames <= 0 || k + 1 < p.max_frames);
        const float *Sf_in = Sfbuf + (k & 1) * Ndft;                 /* after frame k */
        float *Sf_out = Sfbuf + ((k + 1) & 1) * Ndft;
        int round = est_next ? 0 : nrounds;
        int i_done = 0;
        float acc = 0.0f;                                            /* lanes 0/1: re/im of the timing sum (fsk.c:686-698) */
        float2 xn[4];                                                /* the next block's samples, fetched one block ahead */
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int t = u * 32 + lane; xn[u] = (t < BLK && ck + t < fend) ? load_sample_last<FMT>(p.in, gbase + ck + t) : make_float2(0.0f, 0.0f); }
        for (int s0 = ck; s0 < fend; s0 += BLK) {
            const int cnt = min(BLK, fend - s0);
            /* oscillator output not there yet: spend the time on the next frame's FFTs */
            while (ld_acq(&ctl.oscpos) < s0 + cnt) {
                if (round < nrounds) {
                    est_round(ck1, round, Sf_in, Sf_out);
                    if (++round == nrounds) est_finish(k + 1, Sf_out, tab_next);
                } else if (!wpc_wait_ge(&ctl.oscpos, s0 + cnt, &ctl, 400)) { dead = true; break; }
            }
            if (dead) break;
            /* down-convert in place: f_dc = x * conj(phi)  (fsk.c:648); BLK <= 128: at most 4 samples per lane */
            {
                float2 xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    xv[u] = xn[u];
                    const int t = BLK + u * 32 + lane;
                    xn[u] = (u * 32 + lane < BLK && s0 + t < fend) ? load_sample_last<FMT>(p.in, gbase + s0 + t) : make_float2(0.0f, 0.0f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = u * 32 + lane;
                    if (t < cnt) {
                        const int slot = (s0 + t) & RMASK;
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            const float2 ph = ring[m * RS + slot];
                            ring[m * RS + slot] = cmul_rn(xv[u], make_float2(ph.x, -ph.y));
                        }
                    }
                }
            }
            __syncwarp();
            /* every integrate-and-dump window that is complete now (fsk.c:659-668), its timing term (fsk.c:686-696) */
            const int mixed_end = s0 + cnt;
            int i_hi = 0;
            if (mixed_end - base >= Ts) i_hi = min(nfi, (mixed_end - base - Ts) / step + 1);
            const int ni = i_hi - i_done;
            const bool pair_path = (step == 2) && !(Ts & 1);         /* windows start every 2nd sample: 16-byte loads, no bank conflicts */
            const int par = base & 1;
            for (int e = lane; e < ni; e += 32) {
                const int i = i_done + e, a0 = base + i * step;
                float2 a[M];
#pragma unroll
                for (int m = 0; m < M; ++m) a[m] = make_float2(0.0f, 0.0f);
                if (pair_path && !par) {                             /* a0 even: Ts/2 aligned pairs */
#pragma unroll 5
                    for (int j = 0; j < Ts; j += 2) {
                        const int slot = (a0 + j) & RMASK;
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            const float4 v = *reinterpret_cast<const float4 *>(ring + m * RS + slot);
                            a[m] = cadd_rn(a[m], make_float2(v.x, v.y)); a[m] = cadd_rn(a[m], make_float2(v.z, v.w));
                        }
                    }
                } else if (pair_path) {                              /* a0 odd: the aligned pairs straddle the window's ends */
                    {
                        const int slot = (a0 - 1) & RMASK;
#pragma unroll
                        for (int m = 0; m < M; ++m) { const float4 v = *reinterpret_cast<const float4 *>(ring + m * RS + slot); a[m] = cadd_rn(a[m], make_float2(v.z, v.w)); }
                    }
#pragma unroll 4
                    for (int j = 1; j < Ts - 1; j += 2) {
                        const int slot = (a0 + j) & RMASK;
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            const float4 v = *reinterpret_cast<const float4 *>(ring + m * RS + slot);
                            a[m] = cadd_rn(a[m], make_float2(v.x, v.y)); a[m] = cadd_rn(a[m], make_float2(v.z, v.w));
                        }
                    }
                    {
                        const int slot = (a0 + Ts - 1) & RMASK;
#pragma unroll
                        for (int m = 0; m < M; ++m) { const float4 v = *reinterpret_cast<const float4 *>(ring + m * RS + slot); a[m] = cadd_rn(a[m], make_float2(v.x, v.y)); }
                    }
                } else {
#pragma unroll 5
                    for (int j = 0; j < Ts; ++j) {
                        const int slot = (a0 + j) & RMASK;
#pragma unroll
                        for (int m = 0; m < M; ++m) a[m] = cadd_rn(a[m], ring[m * RS + slot]);
                    }
                }
                float ft1 = 0.0f;
#pragma unroll
                for (int m = 0; m < M; ++m) { fintg[(size_t)m * nfi + i] = a[m]; ft1 = __fadd_rn(ft1, mag2_rn(a[m])); }
                const float2 pf = __ldg(p.phi_ft + i);
                tre[e] = __fmul_rn(ft1, pf.x); tim[e] = __fmul_rn(ft1, pf.y);
            }
            __syncwarp();
            if (lane < 2) {                                          /* lane 0: real parts, lane 1: imaginary parts, in order */
                const float *t = lane ? tim : tre;
                int j = 0;
                for (; j + 4 <= ni; j += 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(t + j);
                    acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
                }
                for (; j < ni; ++j) acc = __fadd_rn(acc, t[j]);
            }
            __syncwarp();
            if (lane == 0) *(volatile int *)&ctl.free_[g] = min(base + i_hi * step, fend - p.tailcap);
            i_done = i_hi;
            /* keep the look-ahead moving even when the oscillator is far ahead: one EST round per block */
            if (round < nrounds) {
                est_round(ck1, round, Sf_in, Sf_out);
                if (++round == nrounds) est_finish(k + 1, Sf_out, tab_next);
            }
        }
        if (dead) break;
        while (round < nrounds) {
            est_round(ck1, round, Sf_in, Sf_out);
            if (++round == nrounds) est_finish(k + 1, Sf_out, tab_next);
        }

        /* ---- TIME: fsk.c:700-731 (all lanes compute the same scalars) ---- */
        const float tr = __shfl_sync(0xffffffffu, acc, 0), ti = __shfl_sync(0xffffffffu, acc, 1);
        int nan = 0, nin_next = nin, low = 0, high = high_last; float fract = 0.0f;
        if (isnan(tr) || isnan(ti)) nan = 1;                         /* fsk.c:702-704 */
        else {
            const float nrt = (float)__ddiv_rn((double)em_atan2f(ti, tr), 6.283185307179586476925286766559);
            const float rx_timing = __fmul_rn(nrt, (float)P);
            const float dd = __fsub_rn(nrt, nrt_state);
            nrt_state = nrt;
            if (fabsf(dd) < 0.2f) {
                const float appm = (float)__ddiv_rn(__dmul_rn(1e6, (double)dd), (double)(float)nsym);
                ppm = (float)__dadd_rn(__dmul_rn(.9, (double)ppm), __dmul_rn(.1, (double)appm));
            }
            if (p.lock_nin) nin_next = nin;
            else if (nrt > 0.25f) nin_next = p.N + Ts / 2;
            else if (nrt < -0.25f) nin_next = p.N - Ts / 2;
            else nin_next = p.N;
            low = (int)floorf(rx_timing); fract = __fsub_rn(rx_timing, (float)low); high = (int)ceilf(rx_timing);
        }
        const bool exists_next = (p.max_frames <= 0 || k + 1 < p.max_frames) && (ck1 + nin_next <= nsamp_rel);
        if (lane == 0) { ctl.ninv[g][(k + 1) & 1] = exists_next ? nin_next : -1; st_rel(&ctl.seqT[g], k + 2); }
        high_last = high; nan_last = nan;

        /* ---- DEC: fsk.c:737-817; f_int comes back from L2 ---- */
        const long long ofs = (long long)(fbase + k) * p.Nbits;
        __syncwarp();
        if (!nan) {
            const float omf = __fsub_rn(1.0f, fract);
            constexpr int DB = (M == 2) ? 4 : 2;                      /* symbols per lane whose L2 reads are in flight together */
            for (int i0 = lane; i0 < nsym; i0 += 32 * DB) {
                float2 fa[DB][M], fb[DB][M];
#pragma unroll
                for (int u = 0; u < DB; ++u) {
                    const int i = i0 + 32 * u, st = (i + 1) * P;
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        fa[u][m] = i < nsym ? __ldcg(fintg + (size_t)m * nfi + st + low) : make_float2(0.0f, 0.0f);
                        fb[u][m] = i < nsym ? __ldcg(fintg + (size_t)m * nfi + st + high) : make_float2(0.0f, 0.0f);
                    }
                }
#pragma unroll
                for (int u = 0; u < DB; ++u) {
                const int i = i0 + 32 * u;
                if (i >= nsym) break;
                float tmax[4];
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float2 a = fa[u][m], b = fb[u][m];
                    const float2 t = make_float2(__fadd_rn(__fmul_rn(omf, a.x), __fmul_rn(fract, b.x)),
                                                 __fadd_rn(__fmul_rn(omf, a.y), __fmul_rn(fract, b.y)));
                    tmax[m] = mag2_rn(t);
                }
                float mx = tmax[0]; int sym = 0;
#pragma unroll
                for (int m = 1; m < M; ++m) if (tmax[m] > mx) { mx = tmax[m]; sym = m; }
                s_maxs[i] = make_float2(mx, __fsqrt_rn(mx));
                if (!p.soft) {
                    if (M == 2) outb[ofs + i] = (unsigned char)(sym == 1);
                    else { outb[ofs + 2 * i + 1] = (unsigned char)(sym & 1); outb[ofs + 2 * i] = (unsigned char)((sym & 2) >> 1); }
                } else {
                    float *o = (float *)outb + ofs;
#pragma unroll
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
        if (lane == 0 && p.write_recs) {
            FskFrameRec r;
            r.nin_used = nin; r.nin_next = nin_next; r.norm_rx_timing = nrt_state; r.ppm = ppm;
            r.mean_sum = mean_sum; r.std_sum = es; r.nanflag = nan; r.high_sample = high;
            for (int m = 0; m < 4; ++m) r.tabidx[m] = m < M ? tab_cur[m] : 0;
            p.recs[(long long)c * p.rec_stride + fbase + k] = r;
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 4; ++m) tab_cur[m] = tab_next[m];
        ++frames; ++k; ck = ck1; nin = nin_next; exists = exists_next;
    }
    if (dead && lane == 0 && p.w_abort) atomicExch(p.w_abort, 1);

    /* ---- epilogue: this channel's state back to HBM ---- */
    const int c_end = ck;
    const float *Sf_fin = Sfbuf + ((frames + 1) & 1) * Ndft;         /* buffer (frames-1)&1 */
    for (int i = lane; i < Ndft; i += 32) p.Sf[(long long)c * Ndft + i] = Sf_fin[i];
    for (int e = lane; e < M * p.tailcap; e += 32) {
        const int m = e / p.tailcap, i = e - m * p.tailcap;
        p.fdc_tail[((long long)c * M + m) * p.tailcap + i] = ring[m * RS + ((c_end - p.tailcap + i) & RMASK)];
    }
    if (lane == 0) {
        p.nin[c] = nin; p.nrt[c] = nrt_state; p.ppm[c] = ppm;
        p.nframes[c] = fbase + frames; p.consumed[c] = c_end;
        p.tot_frames[c] += frames; p.tot_samples[c] += c_end;
        if (p.frame_base) p.frame_base[c] = fbase + frames;
    }
    if (frames > 0) {
        const long long ofs = (long long)(fbase + frames - 1) * p.Nbits;
        for (int i = lane; i < p.Nbits; i += 32)
            p.last_out[(long long)c * p.Nbits + i] = p.soft ? __float_as_uint(((const float *)outb)[ofs + i]) : (unsigned int)outb[ofs + i];
        if (p.write_recs && !nan_last) {                             /* raw eye samples of the last frame (fsk.c:869-890) */
            const int neyeoffset = high_last + 1, tot = p.eye_traces * M * p.neyesamp;
            for (int e = lane; e < tot; e += 32) {
                const int row = e / p.neyesamp, j = e - row * p.neyesamp, i = row / M, m = row - i * M;
                const int ind = 2 * P * i + neyeoffset + j * p.eye_dec, flat = m * nfi + ind;
                float2 v = make_float2(0.0f, 0.0f);
                if (flat >= 0 && ind < nfi) v = __ldcg(fintg + flat);
                p.eye_raw[(long long)c * (8 * 160) + row * 160 + j] = __fsqrt_rn(mag2_rn(v));
            }
        }
    }
    if (p.carry_cap > 0) {                                           /* carry the unconsumed tail (push mode); dst lies left of src */
        const long long cursor = start + c_end, tail = p.nsamp - cursor;
        constexpr int es = FMT == 1 ? 8 : (FMT == 0 ? 4 : 2);
        const long long nhalf = tail * es / 2;
        const long long cbase = (long long)c * p.in_stride;
        const unsigned short *src = (const unsigned short *)((const unsigned char *)p.in + (cbase + cursor) * es);
        unsigned short *dst = (unsigned short *)((unsigned char *)p.carry_dst + (cbase + p.carry_cap - tail) * es);
        for (long long k0 = 0; k0 < nhalf; k0 += 32 * 8) {
            unsigned short v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const long long i = k0 + u * 32 + lane; v[u] = i < nhalf ? src[i] : (unsigned short)0; }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 8; ++u) { const long long i = k0 + u * 32 + lane; if (i < nhalf) dst[i] = v[u]; }
            __syncwarp();
        }
        if (lane == 0) p.in_off_next[c] = (int)(p.carry_cap - tail);
    }
}
