/* fsk_duo_kernel.cuh -- one CTA per SM that owns one or two channels: the batched fsk_demod kernel for up to 4 channels per SM (sm_100a).
 *
 * Why.  With two one-channel CTAs of fsk_demod_pipe_kernel on an SM the oscillator recurrence (utils/fsk.c:646-647) runs at 20.8 cycles/step
 * against 15-17 with one CTA and 10.65 alone (profiles/r02_notes.md): two oscillator warps at 2 of 32 lanes, twenty latency-bound role
 * warps behind named barriers, an instruction footprint that no longer stays cached.  fsk_wpc_kernel showed that ONE sequential warp with
 * enough registers does a channel's whole consumer side at ~0.22 IPC.  This kernel combines the two: per channel a handful of sequential
 * warps with flag hand-offs (no groups, no named barriers), one oscillator warp for both channels that owns its SM sub-partition.
 *
 *   slot 0       OSC   lane = channel*M + tone (fsk_wpc_kernel's protocol; only ACTIVE chains store); slots 4, 8, 12 exit at once:
 *                      sub-partition 0 stays quiet                                                                   fsk.c:644-651
 *   per channel  E0    EST rounds r = 0 mod 3 (4 windows each, est_round128c), the ordered Sf recursion over its own and the helpers'
 *                      |X| rows, the tone estimate -> dphi for the oscillator; runs one frame ahead                   fsk.c:438-590
 *                E1,E2 EST rounds r = 1, 2 mod 3 -> |X| rows for E0 (two row sets each, eh_done / iir_h flags)
 *                C0,C1 one warp per half of the tones: blocks of 128 samples behind the oscillator: x*conj(phi) in place,
 *                      integrate-and-dump into the frame's f_int (shared memory, double-buffered); samples two blocks ahead in
 *                      registers; on whole blocks two outputs per lane side by side                                   fsk.c:648,659-668
 *                D     timing terms from f_int, their sequential accumulation, atan2f, ppm, next nin -> publishes frame k+1;
 *                      decisions, outputs, EbNo sums, frame record, state write-back                                  fsk.c:686-817
 *   role map (DUO_ROLEMAP 2): warps that run the same code share a sub-partition and its 6 KB L0 instruction cache:
 *                      1: C0 C1 of both channels | 2: D E2 | 3: E0 E1.
 *
 * The design constraint that shaped the code: the instruction cache.  14 warps in 6 roles share 32 KB behind the L0s; the first version
 * executed ~70 KB of hot code and spent 28 % of its warp samples without an instruction (profiles/r02_notes.md section 7).  Hence one call
 * site for the FFT body, rolled loops wherever a role is latency-bound anyway, an oscillator loop without prologue/epilogue
 * (profiles/code_size.py prints the static size per source region).
 *
 * p.d_G = channels per CTA of this launch: 2, or 1 while the batch has at most one channel per SM (then the second channel's warps exit).
 * Eligibility (fsk_batch.cu): Ndft = 128 (the fused FFT), the pipeline's window-count condition, a frame of at least two blocks, shared
 * memory.  Arithmetic and order are those of the other kernels; the same bit-exact tests run with B200FSK_KERNEL=duo and by default.
 * Compile-time switches kept for experiments (profiles/r02_notes.md): DUO_NE (estimator warps, 2 or 3), DUO_ROLEMAP (0..2),
 * DUO_PREFETCH2, DUO_BACKOFF, DUO_CILP; FSK_PROFILE adds per-role cycle counters (profiles/duo_timing.py).
 */
#pragma once
#include "fsk_wpc_kernel.cuh"

#ifdef FSK_PROFILE
#define DW(expr) ([&]() { const long long t_ = clock64(); const bool r_ = (expr); dwait += clock64() - t_; return r_; }())
#define DPROF_BEGIN long long dwait = 0, dt0 = 0, dt1 = 0, dt2 = 0, dtm = 0; const long long drole0 = clock64(); (void)dt0; (void)dt1; (void)dt2; (void)dtm;
#define DT_MARK dtm = clock64();
#define DT_ADD(v) do { const long long n_ = clock64(); v += n_ - dtm; dtm = n_; } while (0)
#define DPROF_RAW(q, v) do { if (lane == 0 && p.prof) p.prof[(long long)(blockIdx.x * p.d_G + (g < 0 ? 0 : g)) * 16 + (q)] += (v); } while (0)
#define DPROF_END(slot) do { if (lane == 0 && p.prof) { p.prof[(long long)(blockIdx.x * p.d_G + (g < 0 ? 0 : g)) * 16 + 2 * (slot)] += dwait; p.prof[(long long)(blockIdx.x * p.d_G + (g < 0 ? 0 : g)) * 16 + 2 * (slot) + 1] += clock64() - drole0; } } while (0)
#else
#define DW(expr) (expr)
#define DPROF_BEGIN
#define DPROF_END(slot)
#define DT_MARK
#define DT_ADD(v)
#define DPROF_RAW(q, v)
#endif
#define DUO_G       2
#define DUO_THREADS 512          /* 16 warp slots; slot % 4 == 0 is sub-partition 0: slot 0 = OSC, slots 4, 8, 12 exit at once */
#define DUO_RN      1024         /* ring slots per tone (f_dc; the oscillator writes its phases here first) */
#define DUO_RS      (DUO_RN + 2)
#define DUO_BLK     128
#ifndef DUO_NE
#define DUO_NE      3            /* EST warps per channel: E0 (leader: its rounds + the ordered Sf recursion + the estimate), E1, E2 (helpers) */
#endif
#ifndef DUO_PREFETCH2
#define DUO_PREFETCH2 1
#endif
#ifndef DUO_BACKOFF
#define DUO_BACKOFF 1
#endif
#ifndef DUO_CILP
#define DUO_CILP 0               /* integrate two outputs per lane at a time */
#endif
#define DUO_MAGSETS (1 + 2 * (DUO_NE - 1))   /* |X| row sets of 4 windows: E0 uses set 0, every helper alternates between two of its own */

struct DuoChan {                 /* per-channel hand-off words (static shared memory) */
    int eh_done[DUO_NE - 1];     /* helper e: rounds finished since launch (monotonic) */
    int iir_h[DUO_NE - 1];       /* E0: rounds of helper e consumed by the Sf recursion (monotonic): the helper may reuse a row set two rounds later */
    int cons_blocks[2];          /* C0 / C1: blocks whose f_int entries are complete (monotonic) */
    int free2[2];                /* C0 / C1: oldest ring position still needed */
    int ring_ready[2];           /* C0 / C1: integrator memory preloaded, the oscillator may write */
    int dec_frames;              /* D: frames fully decided: C may overwrite f_int buffer k&1 once dec_frames >= k-1 */
    int tab[2][4];               /* LUT indices of frame k's tones (slot k&1), for the frame record */
    int cstart[4];               /* frame start c_k (slot k&3), written by D together with ninv */
};

template <int M, int FMT>
__global__ void __launch_bounds__(DUO_THREADS, 1) fsk_duo_kernel(const FskKParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ WpcCtl ctl;
    __shared__ DuoChan dc[DUO_G];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Nmem = p.Nmem, nfi = p.nfi, P = p.P, Ts = p.Ts, nsym = p.Nsym;
    constexpr int Ndft = 128;
    const int step = Ts / P, nin_min = p.N - Ts / 2;
    constexpr int lgM = (M == 2) ? 1 : 2;
    constexpr int RN = DUO_RN, RS = DUO_RS, RMASK = DUO_RN - 1, BLK = DUO_BLK;

    float  *s_hann = (float *)(smem + p.d_hann);
    float2 *s_tw   = (float2 *)(smem + p.d_tw);
    float  *s_mask = (float *)(smem + p.d_mask);
    float2 *s_phift = (float2 *)(smem + p.d_phift);                  /* fine-timing oscillator table (fsk.c:682-698), nfi entries */

    for (int i = tid; i < Ndft; i += DUO_THREADS) { s_hann[i] = p.hann[i]; s_tw[i] = p.tw[i]; s_mask[i] = p.mask[i]; }
    for (int i = tid; i < nfi; i += DUO_THREADS) s_phift[i] = p.phi_ft[i];
    if (tid == 0) { ctl.oscpos = 0; ctl.abort = 0; }
    if (tid < WPC_MAXG) { ctl.free_[tid] = -p.tailcap; ctl.seqE[tid] = 0; ctl.seqT[tid] = 0; ctl.nsrel[tid] = -1; }
    if (tid < DUO_G) { for (int e = 0; e < DUO_NE - 1; ++e) { dc[tid].eh_done[e] = 0; dc[tid].iir_h[e] = 0; } dc[tid].cons_blocks[0] = dc[tid].cons_blocks[1] = 0; dc[tid].free2[0] = dc[tid].free2[1] = -p.tailcap; dc[tid].ring_ready[0] = dc[tid].ring_ready[1] = 0; dc[tid].dec_frames = 0; }
    __syncthreads();

    /* ---------------------------------------------------------------- role table ---------------------------------------------------- */
    int g = -1, role = -1;                       /* role: 0 C0, 1 E0, 2 E1, 3 D, 4 C1, 5 E2; sub-partition = slot % 4 (0 is the oscillator's) */
#ifndef DUO_ROLEMAP
#define DUO_ROLEMAP 2
#endif
    {
        /* slot -> role: sub-partition 1 = slots 1,5,9,13; 2 = 2,6,10,14; 3 = 3,7,11,15 */
#if DUO_ROLEMAP == 0
        constexpr signed char R[16] = { -1, 0, 1, 2,  -1, 3, 4, 5,  -1, 0, 1, 2,  -1, 3, 4, 5 };   /* SMSP1: C0 D | SMSP2: E0 C1 | SMSP3: E1 E2 */
#elif DUO_ROLEMAP == 1
        constexpr signed char R[16] = { -1, 0, 1, 2,  -1, 5, 4, 3,  -1, 0, 1, 2,  -1, 5, 4, 3 };   /* SMSP1: C0 E2 | SMSP2: E0 C1 | SMSP3: E1 D */
#else
        constexpr signed char R[16] = { -1, 0, 3, 1,  -1, 4, 5, 2,  -1, 0, 3, 1,  -1, 4, 5, 2 };   /* SMSP1: C0 C1 | SMSP2: D E2 | SMSP3: E0 E1 (same code per sub-partition) */
#endif
        role = R[warp];
        if (role >= 0) g = warp >= 8 ? 1 : 0;
    }

    if (warp == 0) {
        /* =========================== OSC (same protocol as fsk_wpc_kernel) ============================== */
        const int og = lane >> lgM, m = lane & (M - 1);
        const int c = blockIdx.x * p.d_G + og;
        const bool valid = og < p.d_G && c < p.nchan;
        if (valid) { wpc_wait_ge(&ctl.nsrel[og], 0, &ctl, 100); wpc_wait_ge(&dc[og].ring_ready[0], 1, &ctl, 100); wpc_wait_ge(&dc[og].ring_ready[1], 1, &ctl, 100); }
        const int nsrel = valid ? ld_vol(&ctl.nsrel[og]) : 0;
        float2 *myring = (float2 *)(smem + p.d_chan0 + (size_t)(valid ? og : 0) * p.d_chan_stride + p.d_ring_off) + (size_t)m * RS;
        float2 ph = valid ? p.phi_c[c * 4 + m] : make_float2(1.0f, 0.0f);
        float2 d = make_float2(1.0f, 0.0f);
        int k = 0, cstart = 0, e_lim = 0;
        bool have_nin = false;
        bool act = valid && nin_min <= nsrel;
        if (act) { if (wpc_wait_ge(&ctl.seqE[og], 1, &ctl, 200)) { d = ctl.dphi[og][0][m]; e_lim = nin_min; } else act = false; }
        int pos = 0;
        long long spin_t0 = 0;
#ifdef FSK_PROFILE
        long long o_step = 0, o_ev = 0; const long long o_t0 = clock64();
#endif
        while (true) {
            if (act && !have_nin && ld_acq(&ctl.seqT[og]) >= k + 1) {
                const int nin = ld_vol(&ctl.ninv[og][k & 1]);
                if (nin < 0) act = false; else { have_nin = true; e_lim = cstart + nin; }
            }
            if (!__any_sync(0xffffffffu, act)) break;
            const int lim = __reduce_min_sync(0xffffffffu, act ? e_lim - pos : (1 << 30));
            if (lim <= 0) {
#ifdef FSK_PROFILE
                const long long te_ = clock64();
#endif
                if (act && pos == e_lim) {
                    if (!have_nin) { if (!wpc_wait_ge(&ctl.seqT[og], k + 1, &ctl, 100)) act = false; }
                    else {
                        const float av = __fsqrt_rn(mag2_rn(ph));
                        ph = make_float2(__fdiv_rn(ph.x, av), __fdiv_rn(ph.y, av));
                        p.phi_c[c * 4 + m] = ph;
                        ++k; cstart = e_lim; have_nin = false;
                        if (!(cstart + nin_min <= nsrel && (p.max_frames <= 0 || k < p.max_frames))) act = false;
                        else if (wpc_wait_ge(&ctl.seqE[og], k + 1, &ctl, 100)) { d = ctl.dphi[og][k & 1][m]; e_lim = cstart + nin_min; }
                        else act = false;
                    }
                }
                __syncwarp();
#ifdef FSK_PROFILE
                o_ev += clock64() - te_;
#endif
                continue;
            }
            int n = min(lim, 256 - (pos & 255));
            const int minfree = __reduce_min_sync(0xffffffffu, act ? min(ld_vol(&dc[og].free2[0]), ld_vol(&dc[og].free2[1])) : (1 << 30));
            if (pos + n - minfree > RN) {
                const int fit = (RN - (pos - minfree)) & ~1;
                if (fit <= 0) {
                    if (!spin_t0) spin_t0 = clock64();
                    if (ld_vol(&ctl.abort)) break;
                    if (clock64() - spin_t0 > WPC_SPIN_LIMIT) { *(volatile int *)&ctl.abort = 1; break; }
                    __nanosleep(200);
                    continue;
                }
                n = min(n, fit);
            }
            spin_t0 = 0;
#ifdef FSK_PROFILE
            const long long ts_ = clock64();
#endif
            if (act) {                                               /* only active chains: see fsk_wpc_kernel.cuh */
                float2 *dst = myring + (pos & RMASK);
                int j = 0;
                if (pos & 1) { ph = cmul_rn(ph, d); dst[0] = ph; j = 1; }
                const int npair = (n - j) & ~1;
                ph = osc_pairs(ph, d, dst + j, npair, true);
                j += npair;
                if (j < n) { ph = cmul_rn(ph, d); dst[j] = ph; }
            }
            __syncwarp();
#ifdef FSK_PROFILE
            o_step += clock64() - ts_;
#endif
            pos += n;
            if (lane == 0) st_rel(&ctl.oscpos, pos);
        }
#ifdef FSK_PROFILE
        if (lane == 0 && p.prof) { long long *q = p.prof + (long long)(blockIdx.x * p.d_G) * 16; q[12] += o_step; q[13] += o_ev; q[14] += clock64() - o_t0; q[15] += pos; }
#endif
        return;
    }
    if (g < 0) return;
    const int c = blockIdx.x * p.d_G + g;
    if (g >= p.d_G || c >= p.nchan) return;                         /* d_G = 1: one channel per CTA (at most one channel per SM in the batch) */

    /* ---------------------------------------------------------------- per-channel memory ----------------------------------------------- */
    unsigned char *cb = smem + p.d_chan0 + (size_t)g * p.d_chan_stride;
    float2 *ring  = (float2 *)(cb + p.d_ring_off);                   /* [M][RS] */
    float2 *fintb = (float2 *)(cb + p.d_fint_off);                   /* [2][M][nfi] */
    float2 *fftE  = (float2 *)(cb + p.d_fft_off);                    /* [DUO_NE warps][4][F128_STRIDE] */
    float  *magb  = (float *)(cb + p.d_mag_off);                     /* [DUO_MAGSETS][4][F128_MSTRIDE] */
    float  *Sfbuf = (float *)(cb + p.d_Sf_off);                      /* [2][Ndft] */
    float  *scr   = (float *)(cb + p.d_scr_off);
    float  *terms = (float *)(cb + p.d_term_off);                    /* [2 slots][2][tcap] */
    float2 *s_maxs = (float2 *)(cb + p.d_smax_off);
    DuoChan &D = dc[g];
    const long long start = p.in_off ? (long long)p.in_off[c] : 0;
    const long long gbase = (long long)c * p.in_stride + start;
    const int nsamp_rel = (int)(p.nsamp - start);
    const int fbase = p.frame_base ? p.frame_base[c] : 0;
    const int numffts = (int)floorf(__fdiv_rn((float)nin_min, (float)(Ndft / 2))) - 1;
    const int nrounds = (numffts + 3) / 4;
    const int tcap = p.d_tcap;
    const int sshift = (step & (step - 1)) == 0 ? 31 - __clz(step) : -1;     /* step is a power of two in every configuration the
                                                                                 reference's CLIs produce: shift instead of a ~40-instruction division per block */
#define DUO_DIVSTEP(x) (sshift >= 0 ? ((x) >> sshift) : (x) / step)

    if (role == 5 && DUO_NE < 3) return;
    if (role == 1 || role == 2 || role == 5) {
        /* =========================== E0 / E1 / E2: estimator, one frame ahead ============================== */
        const int me = role == 1 ? 0 : (role == 2 ? 1 : 2);          /* 0: leader (its rounds + Sf recursion + estimate); 1, 2: helpers */
        DPROF_BEGIN
        F128Lane L128; f128_init(L128, s_hann, lane);
        float2 *F = fftE + me * 4 * F128_STRIDE;
        if (me == 0) for (int i = lane; i < Ndft; i += 32) Sfbuf[Ndft + i] = p.Sf[(long long)c * Ndft + i];   /* "after frame -1" */
        __syncwarp();
        int hcnt[DUO_NE - 1];                                        /* leader: rounds of helper e consumed; helper e: hcnt[e-1] = rounds produced */
#pragma unroll
        for (int e = 0; e < DUO_NE - 1; ++e) hcnt[e] = 0;
        int kk = 0, ck = 0, mycnt = 0;                               /* helper: rounds it produced */
        for (;; ++kk) {
            if (kk > 0) {                                            /* c_kk = c_{kk-1} + nin_{kk-1}: known once frame kk-1 is published */
                if (!DW(wpc_wait_ge_bo(&ctl.seqT[g], kk, &ctl, 200, 1600))) break;   /* a frame ahead: long wait, latency irrelevant */
                const int ninp = ld_vol(&ctl.ninv[g][(kk - 1) & 1]);
                if (ninp < 0) break;                                 /* frame kk-1 does not exist */
                ck = ld_vol(&D.cstart[(kk - 1) & 3]) + ninp;
            }
            if (!(ck + nin_min <= nsamp_rel && (p.max_frames <= 0 || kk < p.max_frames))) break;
            const float *Sf_in = Sfbuf + ((kk + 1) & 1) * Ndft;
            float *Sf_out = Sfbuf + (kk & 1) * Ndft;
            bool dead = false;
            for (int r = me; r < nrounds; r += DUO_NE) {
                /* one call site for the FFT body (instruction-cache footprint): helpers write one of their two row sets, the leader set 0 */
                float *rows = magb;
                if (me > 0) {
                    if (mycnt >= 2 && !DW(wpc_wait_ge_bo(&D.iir_h[me - 1], mycnt - 1, &ctl, 100, 800))) { dead = true; break; }   /* the set is free again */
                    rows = magb + (1 + 2 * (me - 1) + (mycnt & 1)) * 4 * F128_MSTRIDE;
                }
                est_round128c<FMT>(p, gbase + ck, 4 * r, numffts, F, s_tw, L128, lane, rows - 4 * r * F128_MSTRIDE);
                if (me > 0) {
                    ++mycnt;
                    if (lane == 0) st_rel(&D.eh_done[me - 1], mycnt);
                    continue;
                }
                /* leader: rounds r .. r+NE-1 in window order through the Sf recursion (fsk.c:484-500) */
                const float *hrow[DUO_NE - 1];
#pragma unroll
                for (int e = 0; e < DUO_NE - 1; ++e) {
                    hrow[e] = nullptr;
                    if (r + 1 + e < nrounds) {
                        if (!DW(wpc_wait_ge_bo(&D.eh_done[e], hcnt[e] + 1, &ctl, 100, 400))) { dead = true; break; }
                        hrow[e] = magb + (1 + 2 * e + (hcnt[e] & 1)) * 4 * F128_MSTRIDE;
                    }
                }
                if (dead) break;
                const float *Sf_src = r ? Sf_out : Sf_in;
#pragma unroll 2
                for (int q = 0; q < 4; ++q) {                        /* two bins in flight: enough for a latency-bound chain, half the code */
                    const int i = lane + 32 * q, src = (i + 64) & 127;
                    float s = Sf_src[i];
#pragma unroll
                    for (int w = 0; w < 4; ++w) if (4 * r + w < numffts) s = __fadd_rn(__fmul_rn(s, p.omtc), __fmul_rn(rows[w * F128_MSTRIDE + src], p.tc));
#pragma unroll
                    for (int e = 0; e < DUO_NE - 1; ++e) {
                        if (hrow[e]) {
#pragma unroll
                            for (int w = 0; w < 4; ++w) if (4 * (r + 1 + e) + w < numffts) s = __fadd_rn(__fmul_rn(s, p.omtc), __fmul_rn(hrow[e][w * F128_MSTRIDE + src], p.tc));
                        }
                    }
                    Sf_out[i] = s;
                }
                __syncwarp();
#pragma unroll
                for (int e = 0; e < DUO_NE - 1; ++e) if (hrow[e]) { ++hcnt[e]; if (lane == 0) st_rel(&D.iir_h[e], hcnt[e]); }
            }
            if (dead) break;
            if (me == 0) {
                int tab[4];
                estimate_tones(p, Sf_out, scr, s_mask, lane, tab);
                if (lane < M) { ctl.dphi[g][kk & 1][lane] = __ldg(p.dphi_tab + tab[lane]); D.tab[kk & 1][lane] = tab[lane]; }
                __syncwarp();
                if (lane == 0) st_rel(&ctl.seqE[g], kk + 1);
            }
        }
#ifndef DUO_PROF_DETAIL
        DPROF_END(me == 0 ? 1 : (me == 1 ? 2 : 5));
#endif
        return;                                                      /* Sf is written back by D (it knows how many frames really existed) */
    }

    if (role == 0 || role == 4) {
        /* =========================== C0 / C1: mix + integrate for one half of the tones ============================== */
        DPROF_BEGIN
        constexpr int MH = M / 2;                                    /* tones per C warp */
        const int tg = role == 0 ? 0 : 1, m0 = tg * MH;
        for (int e = lane; e < MH * p.tailcap; e += 32) {            /* integrator memory -> ring slots [-tailcap, 0) */
            const int m = m0 + e / p.tailcap, i = e % p.tailcap;
            ring[m * RS + ((i - p.tailcap) & RMASK)] = p.fdc_tail[((long long)c * M + m) * p.tailcap + i];
        }
        __syncwarp();
        if (lane == 0) { if (tg == 0) st_rel(&ctl.nsrel[g], nsamp_rel < 0 ? 0 : nsamp_rel); st_rel(&D.ring_ready[tg], 1); }
        int k = 0, nblk_pub = 0, c_end = 0;
        float2 xa[4], xb[4];                                         /* samples of the next block and of the one after it */
        bool primed = false;
        for (;; ++k) {
            if (!DW(wpc_wait_ge_bo(&ctl.seqT[g], k + 1, &ctl, 100, 400))) break;
            const int nin = ld_vol(&ctl.ninv[g][k & 1]);
            if (nin < 0) break;
            const int ck = ld_vol(&D.cstart[k & 3]);
            const int base = ck + nin - Nmem, fend = ck + nin;
            if (k >= 2 && !DW(wpc_wait_ge(&D.dec_frames, k - 1, &ctl, 200))) break;      /* f_int buffer k&1 is free */
            float2 *fint = fintb + (size_t)(k & 1) * M * nfi;
            int i_done = 0;
            bool dead = false;
            if (!primed) {                                           /* the first two blocks of this launch */
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = u * 32 + lane;
                    xa[u] = (ck + t < nsamp_rel) ? load_sample_t<FMT>(p.in, gbase + ck + t) : make_float2(0.0f, 0.0f);
                    xb[u] = (ck + BLK + t < nsamp_rel) ? load_sample_t<FMT>(p.in, gbase + ck + BLK + t) : make_float2(0.0f, 0.0f);
                }
                primed = true;
            }
            for (int s0 = ck; s0 < fend; s0 += BLK) {
                const int cnt = min(BLK, fend - s0);
                if (!DW(wpc_wait_ge(&ctl.oscpos, s0 + cnt, &ctl, 100))) { dead = true; break; }
                DT_MARK
                /* samples: two blocks in flight (one was not enough to cover an L2/HBM round trip: 12 % of this warp's samples waited
                 * here).  Block starts continue across the frame boundary (the next frame starts at fend; it is at least two blocks long) */
                const int n1 = (s0 + BLK < fend) ? s0 + BLK : fend;
#if DUO_PREFETCH2
                const int n2 = (n1 >= fend) ? n1 + BLK : ((n1 + BLK < fend) ? n1 + BLK : fend);
#else
                const int n2 = n1;                                   /* debugging aid: one block in flight (xa is reloaded below) */
#endif
                float2 xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
#if DUO_PREFETCH2
                    xv[u] = xa[u]; xa[u] = xb[u];
#else
                    xv[u] = xa[u];
#endif
                    const int t = n2 + u * 32 + lane;
                    xb[u] = (t < nsamp_rel) ? load_sample_t<FMT>(p.in, gbase + t) : make_float2(0.0f, 0.0f);
#if !DUO_PREFETCH2
                    xa[u] = xb[u];
#endif
                }
                if (cnt == BLK) {                                    /* whole block (all but the last of a frame): no per-sample guards */
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int slot = (s0 + u * 32 + lane) & RMASK;
#pragma unroll
                        for (int mm = 0; mm < MH; ++mm) { float2 *rp = ring + (m0 + mm) * RS + slot; const float2 ph = *rp; *rp = cmul_rn(xv[u], make_float2(ph.x, -ph.y)); }
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int t = u * 32 + lane;
                        if (t < cnt) {
                            const int slot = (s0 + t) & RMASK;
#pragma unroll
                            for (int mm = 0; mm < MH; ++mm) { const int m = m0 + mm; const float2 ph = ring[m * RS + slot]; ring[m * RS + slot] = cmul_rn(xv[u], make_float2(ph.x, -ph.y)); }
                        }
                    }
                }
                __syncwarp();
                DT_ADD(dt0);
                const int mixed_end = s0 + cnt;
                int i_hi = 0;
                if (mixed_end - base >= Ts) i_hi = min(nfi, DUO_DIVSTEP(mixed_end - base - Ts) + 1);
                const int ni = i_hi - i_done;
                const bool pair_path = (step == 2) && !(Ts & 1);
                const int par = base & 1;
                auto integ_one = [&](const int i, float2 (&a)[MH]) {
                    const int a0 = base + i * step;
#pragma unroll
                    for (int mm = 0; mm < MH; ++mm) a[mm] = make_float2(0.0f, 0.0f);
                    if (pair_path && !par) {
#pragma unroll 5
                        for (int j = 0; j < Ts; j += 2) {
                            const int slot = (a0 + j) & RMASK;
#pragma unroll
                            for (int mm = 0; mm < MH; ++mm) {
                                const float4 v = *reinterpret_cast<const float4 *>(ring + (m0 + mm) * RS + slot);
                                a[mm] = cadd_rn(a[mm], make_float2(v.x, v.y)); a[mm] = cadd_rn(a[mm], make_float2(v.z, v.w));
                            }
                        }
                    } else if (pair_path) {
                        { const int slot = (a0 - 1) & RMASK;
#pragma unroll
                          for (int mm = 0; mm < MH; ++mm) { const float4 v = *reinterpret_cast<const float4 *>(ring + (m0 + mm) * RS + slot); a[mm] = cadd_rn(a[mm], make_float2(v.z, v.w)); } }
#pragma unroll 4
                        for (int j = 1; j < Ts - 1; j += 2) {
                            const int slot = (a0 + j) & RMASK;
#pragma unroll
                            for (int mm = 0; mm < MH; ++mm) {
                                const float4 v = *reinterpret_cast<const float4 *>(ring + (m0 + mm) * RS + slot);
                                a[mm] = cadd_rn(a[mm], make_float2(v.x, v.y)); a[mm] = cadd_rn(a[mm], make_float2(v.z, v.w));
                            }
                        }
                        { const int slot = (a0 + Ts - 1) & RMASK;
#pragma unroll
                          for (int mm = 0; mm < MH; ++mm) { const float4 v = *reinterpret_cast<const float4 *>(ring + (m0 + mm) * RS + slot); a[mm] = cadd_rn(a[mm], make_float2(v.x, v.y)); } }
                    } else {
#pragma unroll 5
                        for (int j = 0; j < Ts; ++j) {
                            const int slot = (a0 + j) & RMASK;
#pragma unroll
                            for (int mm = 0; mm < MH; ++mm) a[mm] = cadd_rn(a[mm], ring[(m0 + mm) * RS + slot]);
                        }
                    }
                };
#if !DUO_CILP
                /* fast path (every block but the first and last of a frame): 64 outputs = two per lane, integrated side by side (two
                 * independent add chains per tone), 16-byte loads at immediate offsets from one pointer per tone (no ring wrap inside) */
                const int sl0 = (base + i_done * 2 - par) & RMASK;
                if (pair_path && ni == 64 && sl0 + 2 * 63 + Ts + 2 <= RN) {
                    const int ia = i_done + lane, h = Ts >> 1;
                    float2 accA[MH], accB[MH];
#pragma unroll
                    for (int mm = 0; mm < MH; ++mm) {
                        const float4 *q4 = reinterpret_cast<const float4 *>(ring + (m0 + mm) * RS + sl0 + 2 * lane);   /* output B: 64 slots = 32 float4 on */
                        float2 a = make_float2(0.0f, 0.0f), b = make_float2(0.0f, 0.0f);
                        if (!par) {
#pragma unroll 5
                            for (int q = 0; q < h; ++q) {
                                const float4 va = q4[q], vb = q4[q + 32];
                                a = cadd_rn(a, make_float2(va.x, va.y)); b = cadd_rn(b, make_float2(vb.x, vb.y));
                                a = cadd_rn(a, make_float2(va.z, va.w)); b = cadd_rn(b, make_float2(vb.z, vb.w));
                            }
                        } else {
                            { const float4 va = q4[0], vb = q4[32]; a = cadd_rn(a, make_float2(va.z, va.w)); b = cadd_rn(b, make_float2(vb.z, vb.w)); }
#pragma unroll 4
                            for (int q = 1; q < h; ++q) {
                                const float4 va = q4[q], vb = q4[q + 32];
                                a = cadd_rn(a, make_float2(va.x, va.y)); b = cadd_rn(b, make_float2(vb.x, vb.y));
                                a = cadd_rn(a, make_float2(va.z, va.w)); b = cadd_rn(b, make_float2(vb.z, vb.w));
                            }
                            { const float4 va = q4[h], vb = q4[h + 32]; a = cadd_rn(a, make_float2(va.x, va.y)); b = cadd_rn(b, make_float2(vb.x, vb.y)); }
                        }
                        accA[mm] = a; accB[mm] = b;
                    }
#pragma unroll
                    for (int mm = 0; mm < MH; ++mm) { fint[(size_t)(m0 + mm) * nfi + ia] = accA[mm]; fint[(size_t)(m0 + mm) * nfi + ia + 32] = accB[mm]; }
                } else
                for (int e = lane; e < ni; e += 32) {
                    const int ia = i_done + e;
                    float2 a0v[MH];
                    integ_one(ia, a0v);
#pragma unroll
                    for (int mm = 0; mm < MH; ++mm) fint[(size_t)(m0 + mm) * nfi + ia] = a0v[mm];
                }
#else
                for (int e = lane; e < ni; e += 64) {                /* two outputs per lane in flight (independent add chains) */
                    const bool two = e + 32 < ni;
                    const int ia = i_done + e, ib = two ? ia + 32 : ia;
                    float2 a0v[MH], a1v[MH];
                    integ_one(ia, a0v); integ_one(ib, a1v);
#pragma unroll
                    for (int mm = 0; mm < MH; ++mm) { fint[(size_t)(m0 + mm) * nfi + ia] = a0v[mm]; if (two) fint[(size_t)(m0 + mm) * nfi + ib] = a1v[mm]; }
                }
#endif
                __syncwarp();
                DT_ADD(dt1);
                ++nblk_pub;
                if (lane == 0) {
                    st_rel(&D.cons_blocks[tg], nblk_pub);
                    *(volatile int *)&D.free2[tg] = min(base + i_hi * step, fend - p.tailcap);
                }
                i_done = i_hi;
                DT_ADD(dt2);
            }
            if (dead) break;
            c_end = fend;
        }
        DPROF_END(tg == 0 ? 0 : 4);
#ifdef DUO_PROF_DETAIL
        if (tg == 1) { DPROF_RAW(4, dt0); DPROF_RAW(5, dt1); DPROF_RAW(10, dt2); }
#endif
        /* the integrator memory of the next launch: the ring rows of these tones are this warp's */
        for (int e = lane; e < MH * p.tailcap; e += 32) {
            const int m = m0 + e / p.tailcap, i = e % p.tailcap;
            p.fdc_tail[((long long)c * M + m) * p.tailcap + i] = ring[m * RS + ((c_end - p.tailcap + i) & RMASK)];
        }
        if (tg == 0 && p.carry_cap > 0) {                            /* carry the unconsumed tail (push mode); dst lies left of src */
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
        return;
    }

    /* =============================== D: timing accumulate, TIME, DEC, frame records, state ======================================= */
    unsigned char *outb = (unsigned char *)p.out + (long long)c * p.out_stride * (p.soft ? 4 : 1);
    DPROF_BEGIN
    int nin = p.nin[c];
    float nrt_state = p.nrt[c], ppm = p.ppm[c];
    int high_last = 0, nan_last = 0;
    bool exists = nin <= nsamp_rel;
    if (lane == 0) { D.cstart[0] = 0; ctl.ninv[g][0] = exists ? nin : -1; st_rel(&ctl.seqT[g], 1); }
    int k = 0, ck = 0, frames = 0, blk_seen = 0;
    bool dead = false;
    while (exists && !dead) {
        const int ck1 = ck + nin;
        const int nb = (nin + BLK - 1) / BLK;
        float acc = 0.0f;
        {
        const int base = ck1 - Nmem;
        const float2 *fintk = fintb + (size_t)(k & 1) * M * nfi;
        float *tre = terms, *tim = terms + tcap;
        int i_done = 0;
        for (int b = 0; b < nb; ++b) {                               /* fsk.c:686-698, block by block, in order */
            if (!DW(wpc_wait_ge(&D.cons_blocks[0], blk_seen + 1, &ctl, 100)) || !DW(wpc_wait_ge(&D.cons_blocks[1], blk_seen + 1, &ctl, 100))) { dead = true; break; }
            const int mixed_end = min(ck + (b + 1) * BLK, ck1);
            int i_hi = 0;
            if (mixed_end - base >= Ts) i_hi = min(nfi, DUO_DIVSTEP(mixed_end - base - Ts) + 1);
            const int ni = i_hi - i_done;
            DT_MARK
            for (int e = lane; e < ni; e += 64) {                    /* the timing terms of the outputs this block completed, two per lane in flight */
                const int e1 = e + 32 < ni ? e + 32 : e;
                const int ia = i_done + e, ib = i_done + e1;
                float fa = 0.0f, fb = 0.0f;
#pragma unroll
                for (int m = 0; m < M; ++m) { fa = __fadd_rn(fa, mag2_rn(fintk[(size_t)m * nfi + ia])); fb = __fadd_rn(fb, mag2_rn(fintk[(size_t)m * nfi + ib])); }
                const float2 pa = s_phift[ia], pb = s_phift[ib];
                tre[e] = __fmul_rn(fa, pa.x); tim[e] = __fmul_rn(fa, pa.y);
                tre[e1] = __fmul_rn(fb, pb.x); tim[e1] = __fmul_rn(fb, pb.y);
            }
            __syncwarp();
            DT_ADD(dt0);
            if (lane < 2) {
                const float *t = lane ? tim : tre;
                int j = 0;
#pragma unroll 8
                for (; j + 4 <= ni; j += 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(t + j);
                    acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
                }
                for (; j < ni; ++j) acc = __fadd_rn(acc, t[j]);
            }
            __syncwarp();
            DT_ADD(dt1);
            ++blk_seen;
            i_done = i_hi;
        }
        }
        if (dead) break;
        DT_MARK
        const float tr = __shfl_sync(0xffffffffu, acc, 0), ti = __shfl_sync(0xffffffffu, acc, 1);
        int nan = 0, nin_next = nin, low = 0, high = high_last; float fract = 0.0f;
        if (isnan(tr) || isnan(ti)) nan = 1;
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
        if (lane == 0) { D.cstart[(k + 1) & 3] = ck1; ctl.ninv[g][(k + 1) & 1] = exists_next ? nin_next : -1; st_rel(&ctl.seqT[g], k + 2); }
        high_last = high; nan_last = nan;

        const float2 *fint = fintb + (size_t)(k & 1) * M * nfi;
        const long long ofs = (long long)(fbase + k) * p.Nbits;
        __syncwarp();
        if (!nan) {
            const float omf = __fsub_rn(1.0f, fract);
            for (int i = lane; i < nsym; i += 32) {
                const int st = (i + 1) * P;
                float tmax[4];
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float2 a = fint[(size_t)m * nfi + st + low], b = fint[(size_t)m * nfi + st + high];
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
        } else {
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
        float es = 0.0f;
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
            for (int m = 0; m < 4; ++m) r.tabidx[m] = m < M ? ld_vol(&D.tab[k & 1][m]) : 0;
            p.recs[(long long)c * p.rec_stride + fbase + k] = r;
        }
        if (p.write_recs && !nan && !exists_next) {                  /* raw eye samples of the last frame (fsk.c:869-890), before the buffer is released */
            const int neyeoffset = high + 1, tot = p.eye_traces * M * p.neyesamp;
            for (int e = lane; e < tot; e += 32) {
                const int row = e / p.neyesamp, j = e - row * p.neyesamp, i = row / M, m = row - i * M;
                const int ind = 2 * P * i + neyeoffset + j * p.eye_dec, flat = m * nfi + ind;
                float2 v = make_float2(0.0f, 0.0f);
                if (flat >= 0 && ind < nfi) v = fint[flat];
                p.eye_raw[(long long)c * (8 * 160) + row * 160 + j] = __fsqrt_rn(mag2_rn(v));
            }
        }
        __syncwarp();
        ++frames; ++k; ck = ck1; nin = nin_next; exists = exists_next;
        if (lane == 0) st_rel(&D.dec_frames, k);
        DT_ADD(dt2);
    }
    DPROF_END(3);
#ifdef DUO_PROF_DETAIL
    DPROF_RAW(2, dt0); DPROF_RAW(3, dt1); DPROF_RAW(11, dt2);
#endif
    if (dead && lane == 0 && p.w_abort) atomicExch(p.w_abort, 1);
    /* ---- this channel's state back to HBM ---- */
    {
        /* Sf after the last frame that existed: E0 may have estimated one frame more; wait until it has stopped touching the buffers */
        if (frames > 0) {                                            /* no frame in this launch: Sf in HBM is already right (E0 may not even have loaded it yet) */
            wpc_wait_ge(&ctl.seqE[g], frames, &ctl, 200);
            const float *Sf_fin = Sfbuf + ((frames + 1) & 1) * Ndft;
            for (int i = lane; i < Ndft; i += 32) p.Sf[(long long)c * Ndft + i] = Sf_fin[i];
        }
    }
    if (lane == 0) {
        p.nin[c] = nin; p.nrt[c] = nrt_state; p.ppm[c] = ppm;
        p.nframes[c] = fbase + frames; p.consumed[c] = ck;
        p.tot_frames[c] += frames; p.tot_samples[c] += ck;
        if (p.frame_base) p.frame_base[c] = fbase + frames;
    }
    if (frames > 0) {
        const long long ofs = (long long)(fbase + frames - 1) * p.Nbits;
        for (int i = lane; i < p.Nbits; i += 32)
            p.last_out[(long long)c * p.Nbits + i] = p.soft ? __float_as_uint(((const float *)outb)[ofs + i]) : (unsigned int)outb[ofs + i];
    }
    (void)nan_last;
}
