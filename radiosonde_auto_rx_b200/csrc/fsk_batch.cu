/* fsk_batch.cu -- host side of libb200fsk.so: the C-ABI of include/b200fsk.h over fsk_kernel.cuh.
 *
 * Host responsibilities (all plain C-style code, no torch types):
 *   * derive the modem geometry exactly as fsk_create_core does (utils/fsk.c:114-201);
 *   * build every table that involves a transcendental with the HOST libm, the same one the reference
 *     links: Hann window (fsk.c:96), kiss_fft twiddles (kiss_fft.c:357-363), the fine-timing
 *     oscillator sequence (fsk.c:682-698) and the (cos,sin) oscillator-step LUT for every tone
 *     estimate the quantised estimators can return (fsk.c:546, 578-581, 645);
 *   * own the per-channel persistent state in HBM, the staging/carry buffers of the push path, and
 *     turn the kernel's per-frame sums into EbNodB / snr_est with libm's log10f (fsk.c:822-845).
 * There is no CPU demodulation path here: without a CUDA device every entry point fails.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>
#include "../../include/b200fsk.h"
#include "fsk_pipe_kernel.cuh"
#include "fsk_wpc_kernel.cuh"
#include "fsk_duo_kernel.cuh"

static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...)
{
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf; return code;
}
/* inside b200fsk_create, once the handle exists: a failure releases everything allocated so far */
#define CKH(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { b200fsk_destroy(h); return fail(B200FSK_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(B200FSK_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

struct b200fsk {
    b200fsk_cfg_t cfg; int nchan, device;
    int Ts, N, Nmem, Ndft, Nbits, nfi, ncap, tailcap, xcap;
    FskKParams kp;                     /* template of kernel params (tables/state pointers filled) */
    size_t smem;                       /* serial kernel */
    size_t smem_pipe = 0; bool use_pipe = false;   /* pipelined kernel */
    size_t smem_wpc = 0; bool can_wpc = false, force_wpc = false; int wpc_G = 0, nsm = 0;   /* throughput kernel (many channels per CTA) */
    float2 *d_fint = 0; int *d_abort = 0; int wpc_thresh = 1 << 30;
    float2 *d_gring = 0;
    bool can_duo = false, force_duo = false, duo_auto = false; size_t smem_duo = 0;   /* two channels per CTA, one CTA per SM */
    size_t fint_bytes = 0; int l2_persist = 0; size_t l2_window = 0; float l2_ratio = 0.0f;   /* L2 persistence window over the f_int slots */
    unsigned short *d_invperm = 0;
    std::vector<float> h_fest;         /* LUT: tone estimate in Hz per table index */
    /* device tables */
    float *d_hann = 0; float2 *d_tw = 0; unsigned short *d_perm = 0; float2 *d_phi_ft = 0; float *d_mask = 0; float2 *d_dphi = 0;
    /* device state */
    float *d_Sf = 0; float2 *d_phi = 0; float2 *d_tail = 0; float *d_nrt = 0, *d_ppm = 0; int *d_nin = 0;
    unsigned int *d_last = 0; long long *d_totf = 0, *d_tots = 0;
    int *d_nframes = 0; long long *d_consumed = 0; float *d_eye = 0;
    FskFrameRec *d_recs = 0; long long rec_cap = 0;      /* per channel */
    /* push path */
    void *d_stage[2] = {0, 0}; int cur = 0;  /* double-buffered staging: kernel t reads [cur], carries its tail into [cur^1] */
    long long stage_cap = 0; int stage_es = 0; int carry_cap = 0; int *d_off = 0; int *d_fbase = 0;
    cudaStream_t copy_stream = 0, out_stream = 0; cudaEvent_t ev_h2d[2] = {0, 0}, ev_k[2] = {0, 0};
    void *d_out = 0; long long out_cap = 0; int out_es = 0;
    std::vector<int> h_tail;           /* per-channel carried samples (host mirror) */
    /* stats */
    int stats_on = 0; bool recs_pending = false; int pending_base0 = 0;
    std::vector<float> snr_est, EbNodB; std::vector<FskFrameRec> last_rec; std::vector<char> have_rec;
    std::vector<FskFrameRec> h_recs; std::vector<int> h_nframes;
    cudaStream_t stream = 0; cudaEvent_t ev0 = 0, ev1 = 0; bool timed = false;
    long launches = 0; int last_kind = -1;
};

const char *b200fsk_last_error(void) { return g_err.c_str(); }

static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

/* kiss_fft factorisation: 4s first, then 2 (kiss_fft.c:309-331), flattened to bottom-up stages. */
static int plan_fft(int n, FskKParams &kp, std::vector<unsigned short> &perm)
{
    int ns = 0, rem = n, p = 4, mlen[FSK_MAXSTAGE];
    double floor_sqrt = floorf(sqrtf((double)n));
    do {
        while (rem % p) {
            switch (p) { case 4: p = 2; break; case 2: p = 3; break; default: p += 2; break; }
            if (p > floor_sqrt) p = rem;
        }
        rem /= p;
        if (ns >= FSK_MAXSTAGE || (p != 2 && p != 4)) return -1;
        kp.radix[ns] = p; mlen[ns] = rem; kp.log2m[ns] = ilog2(rem); ++ns;
    } while (rem > 1);
    kp.nstage = ns;
    int fs = 1;
    for (int i = 0; i < ns; ++i) { kp.fstride[i] = fs; fs *= kp.radix[i]; }
    perm.resize(n);
    for (int pos = 0; pos < n; ++pos) {          /* leaf slot -> input index of the DIT recursion (kiss_fft.c:282-296) */
        int r = pos, idx = 0;
        for (int i = 0; i < ns; ++i) { int k = r / mlen[i]; r -= k * mlen[i]; idx += k * kp.fstride[i]; }
        perm[pos] = (unsigned short)idx;
    }
    return 0;
}

/* the pipelined kernel is specialised on tone count and input sample format */
typedef void (*pipe_fn_t)(const FskKParams);
static const void *pipe_kernel_fn(int M, int fmt, int oscp = 0)
{
    static const pipe_fn_t tab[2][2][4] = {
        {{fsk_demod_pipe_kernel<2, 0, false>, fsk_demod_pipe_kernel<2, 1, false>, fsk_demod_pipe_kernel<2, 2, false>, fsk_demod_pipe_kernel<2, 3, false>},
         {fsk_demod_pipe_kernel<4, 0, false>, fsk_demod_pipe_kernel<4, 1, false>, fsk_demod_pipe_kernel<4, 2, false>, fsk_demod_pipe_kernel<4, 3, false>}},
        {{fsk_demod_pipe_kernel<2, 0, true>, fsk_demod_pipe_kernel<2, 1, true>, fsk_demod_pipe_kernel<2, 2, true>, fsk_demod_pipe_kernel<2, 3, true>},
         {fsk_demod_pipe_kernel<4, 0, true>, fsk_demod_pipe_kernel<4, 1, true>, fsk_demod_pipe_kernel<4, 2, true>, fsk_demod_pipe_kernel<4, 3, true>}}};
    return (const void *)tab[oscp ? 1 : 0][M == 4 ? 1 : 0][fmt & 3];
}

static const void *wpc_kernel_fn(int M, int fmt)
{
    static const pipe_fn_t tab[2][4] = {
        {fsk_wpc_kernel<2, 0>, fsk_wpc_kernel<2, 1>, fsk_wpc_kernel<2, 2>, fsk_wpc_kernel<2, 3>},
        {fsk_wpc_kernel<4, 0>, fsk_wpc_kernel<4, 1>, fsk_wpc_kernel<4, 2>, fsk_wpc_kernel<4, 3>}};
    return (const void *)tab[M == 4 ? 1 : 0][fmt & 3];
}

static const void *duo_kernel_fn(int M, int fmt)
{
    static const pipe_fn_t tab[2][4] = {
        {fsk_duo_kernel<2, 0>, fsk_duo_kernel<2, 1>, fsk_duo_kernel<2, 2>, fsk_duo_kernel<2, 3>},
        {fsk_duo_kernel<4, 0>, fsk_duo_kernel<4, 1>, fsk_duo_kernel<4, 2>, fsk_duo_kernel<4, 3>}};
    return (const void *)tab[M == 4 ? 1 : 0][fmt & 3];
}

/* Max dynamic shared memory is a per-FUNCTION attribute: always raise it to what the device allows for that function
   (opt-in maximum minus its static shared memory) so that handles with different working sets can coexist. */
static cudaError_t allow_max_smem(const void *fn, int maxsmem_optin)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, fn);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsmem_optin - (int)fa.sharedSizeBytes);
}

template <class T> static cudaError_t upload(T **dptr, const std::vector<T> &v)
{
    cudaError_t e = cudaMalloc((void **)dptr, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}
template <class T> static cudaError_t zalloc(T **dptr, size_t n)
{
    cudaError_t e = cudaMalloc((void **)dptr, n * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemset(*dptr, 0, n * sizeof(T));
}

extern "C" void b200fsk_destroy(b200fsk_t *h);

extern "C" int b200fsk_create(b200fsk_t **out, const b200fsk_cfg_t *cfg, int nchan, int device)
{
    if (!out || !cfg || nchan <= 0) return fail(B200FSK_EINVAL, "bad arguments");
    const int Fs = cfg->Fs, Rs = cfg->Rs, M = cfg->M, P = cfg->P, Nsym = cfg->Nsym;
    /* the reference's asserts (fsk.c:120-130, 1019-1021) become error codes */
    if (Fs <= 0 || Rs <= 0 || cfg->f1_tx <= 0 || cfg->fs_tx <= 0 || P <= 0 || Nsym <= 0) return fail(B200FSK_EINVAL, "non-positive parameter");
    if (Fs % Rs) return fail(B200FSK_EINVAL, "Fs/Rs must be an integer");
    if ((Fs / Rs) % P) return fail(B200FSK_EINVAL, "Ts/P must be an integer");
    if (M != 2 && M != 4) return fail(B200FSK_EINVAL, "M must be 2 or 4");
    if (cfg->est_min < -Fs / 2 || cfg->est_max > Fs / 2 || cfg->est_max <= cfg->est_min) return fail(B200FSK_EINVAL, "estimator limits");
    if (Nsym < 4) return fail(B200FSK_EINVAL, "Nsym >= 4 required");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(B200FSK_ECUDA, "no CUDA device: libb200fsk has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(B200FSK_ERANGE, "device %d out of range", device);
    CK(cudaSetDevice(device));

    b200fsk *h = new b200fsk();
    h->cfg = *cfg; h->nchan = nchan; h->device = device;
    float bin_width_Hz = 0.1 * Rs;                                   /* fsk.c:135-137 */
    float Ndft_f = (float)Fs / bin_width_Hz;
    Ndft_f = pow(2.0, ceil(log2(Ndft_f)));
    const int Ndft = (int)Ndft_f, Ts = Fs / Rs, N = Ts * Nsym, Nmem = N + 2 * Ts, nfi = (Nsym + 1) * P;
    if (Ndft < 4 || Ndft > 4096) { b200fsk_destroy(h); return fail(B200FSK_EINVAL, "Ndft=%d unsupported", Ndft); }
    h->Ts = Ts; h->N = N; h->Nmem = Nmem; h->Ndft = Ndft; h->nfi = nfi; h->Nbits = (M == 2) ? Nsym : 2 * Nsym;
    h->ncap = N + Ts / 2; h->tailcap = 2 * Ts + Ts / 2; h->xcap = h->ncap > nfi ? h->ncap : nfi;

    FskKParams &kp = h->kp; memset(&kp, 0, sizeof kp);
    kp.Ts = Ts; kp.M = M; kp.P = P; kp.Nsym = Nsym; kp.N = N; kp.Nmem = Nmem; kp.Ndft = Ndft; kp.Nbits = h->Nbits;
    kp.nfi = nfi; kp.ncap = h->ncap; kp.tailcap = h->tailcap; kp.xcap = h->xcap;
    kp.tc = 0.95 * Ndft_f / Fs;                                      /* fsk.c:148 */
    kp.omtc = 1 - kp.tc;                                             /* the (1-tc) of fsk.c:500, a float op */
    int st = (cfg->est_min * Ndft) / Fs + Ndft / 2; if (st < 0) st = 0;       /* fsk.c:459-460 */
    int en = (cfg->est_max * Ndft) / Fs + Ndft / 2; if (en > Ndft) en = Ndft;
    const int est_space = 0.75 * Rs;                                 /* fsk.c:158 */
    kp.st = st; kp.en = en; kp.f_zero = (est_space * Ndft) / Fs; kp.est_type = cfg->est_type ? 1 : 0;
    kp.eye_dec = ceil(((float)P * 2) / B200FSK_EYE_IND);             /* fsk.c:864-871 */
    kp.neyesamp = (P * 2) / kp.eye_dec; kp.eye_traces = B200FSK_EYE_TR / M;

    std::vector<unsigned short> perm;
    if (plan_fft(Ndft, kp, perm)) { b200fsk_destroy(h); return fail(B200FSK_EINVAL, "FFT plan"); }
    std::vector<float> hann(Ndft), mask(Ndft, 0.0f);
    std::vector<float2> tw(Ndft), phi_ft(nfi);
    for (int i = 0; i < Ndft; ++i) hann[i] = 0.5 - 0.5 * cosf(2.0 * M_PI * (float)i / (float)(Ndft - 1));   /* fsk.c:96 */
    for (int i = 0; i < Ndft; ++i) {                                 /* kiss_fft.c:357-363 */
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        double phase = -2 * pi * i / Ndft;
        tw[i].x = (float)cosf(phase); tw[i].y = (float)sinf(phase);
    }
    {   /* fine-timing oscillator sequence (fsk.c:682-698): restarts at (1,0) each frame, input independent */
        const float ph = 2 * M_PI * ((float)(Rs) / (float)(P * Rs));
        const float dr = cosf(ph), di = sinf(ph);
        float pr = 1, pi_ = 0;
        for (int i = 0; i < nfi; ++i) {
            phi_ft[i].x = pr; phi_ft[i].y = pi_;
            const float nr = pr * dr - pi_ * di, ni = pr * di + pi_ * dr;
            pr = nr; pi_ = ni;
        }
    }
    {   /* mask of the second estimator (fsk.c:552-560) */
        for (int i = 0; i < 3; ++i) mask[i] = 1.0f;
        int bin = 0;
        for (int m = 1; m <= M - 1; ++m) {
            bin = round((float)m * cfg->fs_tx * Ndft / Fs) - 1;
            if (bin >= 0) for (int i = bin; i <= bin + 2 && i < Ndft; ++i) mask[i] = 1.0f;   /* size_t loop: bin<0 => no iterations */
        }
        kp.len_mask = bin + 2 + 1;
    }
    /* oscillator-step LUT: every value the quantised estimators can produce */
    std::vector<float2> dphi; std::vector<float> &fest = h->h_fest;
    if (!kp.est_type) {
        dphi.resize(Ndft); fest.resize(Ndft);
        for (int b = 0; b < Ndft; ++b) {
            const float f = (float)(b - Ndft / 2) * ((float)Fs / (float)Ndft);       /* fsk.c:546 */
            const float phi = 2 * M_PI * ((f) / (float)(Fs));                         /* fsk.c:645 */
            fest[b] = f; dphi[b].x = cosf(phi); dphi[b].y = sinf(phi);
        }
    } else {
        dphi.resize((size_t)Ndft * M); fest.resize((size_t)Ndft * M);
        for (int b = 0; b < Ndft; ++b) {
            const float foff = (b - Ndft / 2) * Fs / Ndft;                            /* integer maths, fsk.c:578 */
            for (int m = 0; m < M; ++m) {
                const float f = foff + m * cfg->fs_tx;                                /* fsk.c:581 */
                const float phi = 2 * M_PI * ((f) / (float)(Fs));
                fest[(size_t)b * M + m] = f; dphi[(size_t)b * M + m].x = cosf(phi); dphi[(size_t)b * M + m].y = sinf(phi);
            }
        }
    }

    /* shared-memory layout */
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return (int)r; };
    kp.o_x = take((size_t)h->xcap * 8); kp.o_fdc = take((size_t)M * Nmem * 8); kp.o_fint = take((size_t)M * nfi * 8);
    kp.o_fft = take((size_t)FSK_WARPS * Ndft * 8); kp.o_mag = take((size_t)FSK_WARPS * Ndft * 4);
    kp.o_Sf = take(Ndft * 4); kp.o_hann = take(Ndft * 4); kp.o_tw = take(Ndft * 8); kp.o_perm = take(Ndft * 2);
    kp.o_scr = take(Ndft * 4); kp.o_mask = take(Ndft * 4); kp.o_smax = take((size_t)Nsym * 8);
    h->smem = o;
    /* pipelined kernel: usable when numffts does not depend on which of the three nin values a frame has */
    {
        const int nmin = N - Ts / 2, nmax = N + Ts / 2;
        const int f0 = (int)floor((float)nmin / (Ndft / 2)), f1 = (int)floor((float)nmax / (Ndft / 2));
        size_t q = 0;
        auto takeq = [&](size_t bytes) { size_t r = q; q += (bytes + 15) & ~(size_t)15; return (int)r; };
        kp.q_ring = takeq((size_t)M * PIPE_RS * 8); kp.q_fint = takeq((size_t)2 * M * nfi * 8); kp.q_stage_bytes = 2 * M * nfi * 8;
        kp.q_fft = takeq((size_t)PIPE_NEST * (Ndft == 128 ? 4 * F128_STRIDE : PIPE_NFB * (Ndft + Ndft / 8)) * 8);
        kp.q_mag = takeq(Ndft == 128 ? (size_t)4 * PIPE_NEST * F128_MSTRIDE * 4 : (size_t)(f0 > 1 ? f0 - 1 : 1) * Ndft * 4);
        kp.q_Sf = takeq((size_t)2 * Ndft * 4); kp.q_hann = takeq(Ndft * 4); kp.q_tw = takeq(Ndft * 8); kp.q_perm = takeq(Ndft * 2);
        kp.q_scr = takeq(Ndft * 4); kp.q_mask = takeq(Ndft * 4); kp.q_smax = takeq((size_t)Nsym * 8); kp.q_tbuf = takeq(PIPE_TB * 8);
        h->smem_pipe = q;
        const char *force = getenv("B200FSK_KERNEL");
        h->use_pipe = (f0 == f1) && (f0 - 1 >= 1) && (3 * Ts + 2 * PIPE_BLK + 16 <= PIPE_R) && (nfi >= 2) && (Ndft <= 256) && (Ts % 2 == 0 || true) &&
                      (h->smem_pipe + 1024 <= (size_t)227 * 1024);
        if (force && !strcmp(force, "serial")) h->use_pipe = false;
    }
    int maxsmem = 0;
    CKH(cudaDeviceGetAttribute(&maxsmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CKH(cudaDeviceGetAttribute(&h->nsm, cudaDevAttrMultiProcessorCount, device));
    kp.nchan = nchan;
    {   /* throughput kernel: G channels per CTA (one warp each) + one lane-per-chain oscillator warp; same eligibility as the pipeline */
        const int nmin = N - Ts / 2, nmax = N + Ts / 2;
        const int f0 = (int)floor((float)nmin / (Ndft / 2)), f1 = (int)floor((float)nmax / (Ndft / 2));
        const int Gmax = 32 / M;
        size_t q = 0;
        auto takew = [&](size_t bytes) { size_t r = q; q += (bytes + 15) & ~(size_t)15; return (int)r; };
        kp.w_hann = takew(Ndft * 4); kp.w_tw = takew(Ndft * 8); kp.w_perm = takew(Ndft * 2); kp.w_mask = takew(Ndft * 4);
        q = (q + 127) & ~(size_t)127; kp.w_chan0 = (int)q;
        auto layout = [&](int ring, int blk, size_t &stride) {
            size_t o2 = 0;
            auto t2 = [&](size_t bytes) { size_t r = o2; o2 += (bytes + 15) & ~(size_t)15; return (int)r; };
            kp.w_ring_off = t2((size_t)M * (ring + 2) * 8);
            kp.w_fp = (WPC_NFB == 4 && Ndft == 128) ? F128_STRIDE : Ndft + Ndft / 8;
            size_t work = (size_t)WPC_NFB * kp.w_fp * 8;                /* FFT buffers, time-shared with the timing terms / per-symbol maxima */
            /* outputs one block can complete: blk/step new ones + those that were waiting on the integrator memory (Nmem - nin <= 2.5 Ts) */
            kp.w_tcap = (blk + 3 * Ts + 8) & ~3;
            if (work < (size_t)kp.w_tcap * 8) work = (size_t)kp.w_tcap * 8;
            if (work < (size_t)Nsym * 8) work = (size_t)Nsym * 8;
            if (work < (size_t)Ndft * 4) work = (size_t)Ndft * 4;                    /* ... and the estimator scratch */
            kp.w_fft_off = t2(work); kp.w_Sf_off = t2((size_t)2 * Ndft * 4); kp.w_scr_off = kp.w_fft_off;
            /* channel stride = 32 B (M=2) / 64 B (M=4) modulo 128: the oscillator's 16-byte stores of a quarter warp hit 8 distinct bank groups */
            const size_t want = (M == 2) ? 32 : 64;
            o2 = (o2 + 127) & ~(size_t)127; o2 += want;
            stride = o2;
        };
        int ring = 512, blk = 128; size_t stride = 0;
        layout(ring, blk, stride);
        const size_t budget = (size_t)maxsmem / WPC_CTAS - 2048 - (WPC_CTAS > 1 ? 1024 : 0);
        int G = (int)((budget - kp.w_chan0) / stride);
        if (G < Gmax && 3 * Ts + 2 * 64 + 16 <= 256) {                            /* smaller ring: more channels per CTA */
            size_t s2 = 0; layout(256, 64, s2);
            const int G2 = (int)((budget - kp.w_chan0) / s2);
            if (G2 > G) { G = G2; ring = 256; blk = 64; stride = s2; } else layout(ring, blk, stride);
        }
        if (G > Gmax) G = Gmax;
        if (G > WPC_MAXG) G = WPC_MAXG;
        kp.w_ring = ring; kp.w_blk = blk; kp.w_chan_stride = (int)stride;
        h->wpc_G = G;
        h->wpc_thresh = 4 * h->nsm;                               /* measured cross-over (profiles/r02_notes.md): beyond 4 channels per SM the throughput kernel wins */
        if (const char *t = getenv("B200FSK_WPC_THRESH")) h->wpc_thresh = atoi(t);
        h->can_wpc = (f0 == f1) && (f0 - 1 >= 1) && (3 * Ts + 2 * blk + 16 <= ring) && (nfi >= 2) && (Ndft <= 256) && G >= 1;
        const char *force = getenv("B200FSK_KERNEL");
        h->force_wpc = h->can_wpc && force && !strcmp(force, "wpc");
        if (force && (!strcmp(force, "serial") || !strcmp(force, "pipe"))) h->can_wpc = false;
        if (force && !strcmp(force, "wpc") && !h->can_wpc) fprintf(stderr, "b200fsk: B200FSK_KERNEL=wpc but this modem is not eligible\n");
    }
    {   /* two-channels-per-CTA kernel: needs the fused 128-point FFT and the pipeline's window-count condition */
        const int nmin = N - Ts / 2, nmax = N + Ts / 2;
        const int f0 = (int)floor((float)nmin / (Ndft / 2)), f1 = (int)floor((float)nmax / (Ndft / 2));
        size_t q = 0;
        auto take = [&](size_t bytes) { size_t r = q; q += (bytes + 15) & ~(size_t)15; return (int)r; };
        kp.d_hann = take(Ndft * 4); kp.d_tw = take(Ndft * 8); kp.d_mask = take(Ndft * 4); kp.d_phift = take((size_t)nfi * 8);
        q = (q + 127) & ~(size_t)127; kp.d_chan0 = (int)q;
        size_t o2 = 0;
        auto t2 = [&](size_t bytes) { size_t r = o2; o2 += (bytes + 15) & ~(size_t)15; return (int)r; };
        kp.d_tcap = (DUO_BLK + 3 * Ts + 8) & ~3;
        kp.d_ring_off = t2((size_t)M * DUO_RS * 8); kp.d_fint_off = t2((size_t)2 * M * nfi * 8); kp.d_fft_off = t2((size_t)DUO_NE * 4 * F128_STRIDE * 8);
        kp.d_mag_off = t2((size_t)DUO_MAGSETS * 4 * F128_MSTRIDE * 4); kp.d_Sf_off = t2((size_t)2 * Ndft * 4); kp.d_scr_off = t2(Ndft * 4);
        kp.d_term_off = t2((size_t)2 * kp.d_tcap * 4); kp.d_smax_off = t2((size_t)Nsym * 8);
        o2 = (o2 + 127) & ~(size_t)127; o2 += (M == 2) ? 32 : 64;          /* bank offset between the two channels' rings (oscillator stores) */
        kp.d_chan_stride = (int)o2;
        h->smem_duo = (size_t)kp.d_chan0 + (size_t)DUO_G * o2;
        h->can_duo = (Ndft == 128) && (f0 == f1) && (f0 - 1 >= 1) && (3 * Ts + 2 * DUO_BLK + 16 <= DUO_RN) && (N - Ts / 2 > 2 * DUO_BLK) && (nfi >= 2) && (32 / M >= DUO_G) &&
                     (h->smem_duo + 4096 <= (size_t)maxsmem);
        const char *force = getenv("B200FSK_KERNEL");
        h->force_duo = h->can_duo && force && !strcmp(force, "duo");
        h->duo_auto = true;                                       /* measured (profiles/r02_notes.md): wins for nsm < nchan <= 4 nsm (one or two full waves of
                                                                     2 channels per SM); at most one channel per SM the one-CTA-per-channel pipeline is faster */
        if (const char *t = getenv("B200FSK_DUO")) h->duo_auto = atoi(t) != 0;
        if (force && strcmp(force, "duo")) h->can_duo = false;
    }
    const bool serial_fits = h->smem + 256 <= (size_t)maxsmem;
    if (h->use_pipe && h->smem_pipe + 1024 > (size_t)maxsmem) h->use_pipe = false;
    if (!h->use_pipe && !serial_fits) { b200fsk_destroy(h); return fail(B200FSK_ENOMEM, "frame working set %zu B exceeds shared memory %d B", h->smem, maxsmem); }
    /* the attribute is per function, not per handle: always allow the device maximum so handles of different sizes coexist */
    CKH(allow_max_smem((const void *)fsk_demod_kernel, maxsmem));
    if (h->use_pipe) {
        for (int f = 0; f < 4; ++f) {
            CKH(allow_max_smem(pipe_kernel_fn(M, f, 0), maxsmem));
            CKH(allow_max_smem(pipe_kernel_fn(M, f, 1), maxsmem));
        }
    }
    if (h->use_pipe) { CKH(zalloc(&h->d_gring, (size_t)nchan * M * PIPE_R)); kp.g_ring = h->d_gring; }
    if (h->can_duo) { for (int f = 0; f < 4; ++f) CKH(allow_max_smem(duo_kernel_fn(M, f), maxsmem)); if (!h->d_abort) { CKH(zalloc(&h->d_abort, 1)); kp.w_abort = h->d_abort; } }
    if (h->can_wpc) {
        for (int f = 0; f < 4; ++f) CKH(allow_max_smem(wpc_kernel_fn(M, f), maxsmem));
        {   /* f_int spill: one slot per (SM, channel warp); with WPC_CTAS > 1 two CTAs may share an SM -> fall back to per-channel slots */
            const int slots = (WPC_CTAS == 1) ? h->nsm : nchan;       /* nchan >= number of CTAs */
            kp.w_fint_slots = slots;
            CKH(zalloc(&h->d_fint, (size_t)slots * WPC_MAXG * M * nfi));
            h->fint_bytes = (size_t)slots * WPC_MAXG * M * nfi * sizeof(float2);
            /* The f_int slots are rewritten once per frame by the CTA that owns them and read back by the same warp: a hot set of
               #SMs x 15 x 24 KB (53 MB at the RS41 line) that the input stream would otherwise push out of L2 between rewrites
               (ncu: 8 B/sample of extra DRAM writes).  Pin it with a persisting access-policy window on the launching stream. */
            const char *e = getenv("B200FSK_L2PERSIST");
            int maxp = 0, maxw = 0;
            cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, device);
            cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, device);
            if ((!e || atoi(e)) && maxp > 0 && maxw > 0 && WPC_CTAS == 1) {
                const size_t want = h->fint_bytes < (size_t)maxp ? h->fint_bytes : (size_t)maxp;
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
                    h->l2_persist = 1;
                    h->l2_window = h->fint_bytes < (size_t)maxw ? h->fint_bytes : (size_t)maxw;
                    h->l2_ratio = (float)((double)want / (double)h->l2_window); if (h->l2_ratio > 1.0f) h->l2_ratio = 1.0f;
                } else cudaGetLastError();
            }
        }
        if (!h->d_abort) CKH(zalloc(&h->d_abort, 1));
        kp.w_fint = h->d_fint; kp.w_abort = h->d_abort;
    }

    std::vector<unsigned short> invperm(Ndft);
    for (int pos = 0; pos < Ndft; ++pos) invperm[perm[pos]] = (unsigned short)pos;
    CKH(upload(&h->d_invperm, invperm));
    CKH(upload(&h->d_hann, hann)); CKH(upload(&h->d_tw, tw)); CKH(upload(&h->d_perm, perm));
    CKH(upload(&h->d_phi_ft, phi_ft)); CKH(upload(&h->d_mask, mask)); CKH(upload(&h->d_dphi, dphi));
    CKH(zalloc(&h->d_Sf, (size_t)nchan * Ndft)); CKH(zalloc(&h->d_tail, (size_t)nchan * M * h->tailcap));
    CKH(zalloc(&h->d_nrt, nchan)); CKH(zalloc(&h->d_ppm, nchan)); CKH(zalloc(&h->d_last, (size_t)nchan * h->Nbits));
    CKH(zalloc(&h->d_totf, nchan)); CKH(zalloc(&h->d_tots, nchan)); CKH(zalloc(&h->d_nframes, nchan)); CKH(zalloc(&h->d_consumed, nchan));
    CKH(zalloc(&h->d_eye, (size_t)nchan * 8 * 160)); CKH(zalloc(&h->d_off, (size_t)2 * nchan));
    {
        std::vector<float2> phi((size_t)nchan * 4, make_float2(cosf(0), sinf(0)));   /* comp_exp_j(0), fsk.c:164 */
        std::vector<int> nin(nchan, N);
        CKH(upload(&h->d_phi, phi)); CKH(upload(&h->d_nin, nin));
    }
    kp.hann = h->d_hann; kp.tw = h->d_tw; kp.perm = h->d_perm; kp.phi_ft = h->d_phi_ft; kp.mask = h->d_mask; kp.dphi_tab = h->d_dphi; kp.invperm = h->d_invperm;
    kp.Sf = h->d_Sf; kp.phi_c = h->d_phi; kp.fdc_tail = h->d_tail; kp.nrt = h->d_nrt; kp.ppm = h->d_ppm; kp.nin = h->d_nin;
    kp.last_out = h->d_last; kp.tot_frames = h->d_totf; kp.tot_samples = h->d_tots;
    kp.nframes = h->d_nframes; kp.consumed = h->d_consumed; kp.eye_raw = h->d_eye;
    h->snr_est.assign(nchan, 0.0f); h->EbNodB.assign(nchan, 0.0f); h->last_rec.resize(nchan); h->have_rec.assign(nchan, 0);
    h->h_tail.assign(nchan, 0); h->h_nframes.assign(nchan, 0);
    CKH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CKH(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CKH(cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
    CKH(cudaEventCreate(&h->ev0)); CKH(cudaEventCreate(&h->ev1));
    for (int i = 0; i < 2; ++i) { CKH(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming)); CKH(cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming)); }
    CKH(zalloc(&h->d_fbase, nchan));
    *out = h;
    return B200FSK_OK;
}

extern "C" void b200fsk_destroy(b200fsk_t *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void *ptrs[] = {h->d_invperm, h->d_hann, h->d_tw, h->d_perm, h->d_phi_ft, h->d_mask, h->d_dphi, h->d_Sf, h->d_phi, h->d_tail, h->d_nrt,
                    h->d_ppm, h->d_nin, h->d_last, h->d_totf, h->d_tots, h->d_nframes, h->d_consumed, h->d_eye, h->d_recs,
                    h->d_stage[0], h->d_stage[1], h->d_off, h->d_out, h->d_fbase, h->d_fint, h->d_abort, h->d_gring};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->out_stream) cudaStreamDestroy(h->out_stream);
    for (int i = 0; i < 2; ++i) { if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]); if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]); }
    cudaGetLastError();
    delete h;
}

extern "C" int b200fsk_info(const b200fsk_t *h, b200fsk_info_t *info)
{
    if (!h || !info) return fail(B200FSK_EINVAL, "null");
    info->Ts = h->Ts; info->N = h->N; info->Nmem = h->Nmem; info->Ndft = h->Ndft; info->Nbits = h->Nbits;
    info->nin_max = h->N + h->Ts / 2; info->nin_min = h->N - h->Ts / 2; info->nchan = h->nchan; info->smem_bytes = h->use_pipe ? h->smem_pipe : h->smem;
    return B200FSK_OK;
}

extern "C" long b200fsk_max_frames(const b200fsk_t *h, long nsamp)
{
    const int nmin = h->N - h->Ts / 2;
    return nsamp / (nmin > 0 ? nmin : 1) + 1;
}

/* frames one push of nsamp new samples can emit per channel (carried samples included) */
extern "C" long b200fsk_push_frames_bound(const b200fsk_t *h, long nsamp, int max_frames)
{
    if (max_frames > 0) return max_frames;
    int max_tail = 0; for (int c = 0; c < h->nchan; ++c) max_tail = h->h_tail[c] > max_tail ? h->h_tail[c] : max_tail;
    return b200fsk_max_frames(h, (long)max_tail + nsamp);
}

static int ensure_recs(b200fsk *h, long long frames_per_chan)
{
    if (frames_per_chan <= h->rec_cap) return 0;
    if (h->d_recs) cudaFree(h->d_recs);
    h->rec_cap = frames_per_chan + 64;
    CK(cudaMalloc((void **)&h->d_recs, (size_t)h->nchan * h->rec_cap * sizeof(FskFrameRec)));
    return 0;
}

/* EbNodB / snr_est chain on the host with the reference's libm (fsk.c:822-845). */
static void digest_records(b200fsk *h, const FskFrameRec *recs, long long stride, const int *nframes, int base0)
{
    const int nsym = h->cfg.Nsym;
    for (int c = 0; c < h->nchan; ++c) {
        for (int k = base0; k < nframes[c]; ++k) {
            const FskFrameRec &r = recs[(long long)c * stride + k];
            if (!r.nanflag) {
                float meanebno = r.mean_sum / (float)nsym;
                float stdebno = (r.std_sum / (float)nsym) - (meanebno * meanebno);
                if (stdebno > 0.0) stdebno = sqrt(stdebno); else stdebno = 0.0;
                h->EbNodB[c] = -6 + (20 * log10f((1e-6 + meanebno) / (1e-6 + stdebno)));
                h->snr_est[c] = .5 * h->snr_est[c] + .5 * h->EbNodB[c];
            }
            h->last_rec[c] = r; h->have_rec[c] = 1;
        }
    }
}

static int flush_pending(b200fsk *h)
{
    if (!h->recs_pending) return 0;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(h->h_nframes.data(), h->d_nframes, sizeof(int) * h->nchan, cudaMemcpyDeviceToHost));
    int mx = 0; for (int c = 0; c < h->nchan; ++c) mx = h->h_nframes[c] > mx ? h->h_nframes[c] : mx;
    if (mx > 0) {
        h->h_recs.resize((size_t)h->nchan * mx);
        CK(cudaMemcpy2D(h->h_recs.data(), (size_t)mx * sizeof(FskFrameRec), h->d_recs, (size_t)h->rec_cap * sizeof(FskFrameRec),
                        (size_t)mx * sizeof(FskFrameRec), h->nchan, cudaMemcpyDeviceToHost));
        digest_records(h, h->h_recs.data(), mx, h->h_nframes.data(), h->pending_base0);
    }
    h->recs_pending = false;
    return 0;
}

static int launch(b200fsk *h, FskKParams &kp, cudaStream_t s, bool time_it)
{
    if (time_it) CK(cudaEventRecord(h->ev0, s));
    /* kernel choice by channels per SM (all measured, profiles/r02_notes.md): up to 4: the duo kernel (one oscillator warp on its own
       sub-partition + specialised warps per channel; one channel per CTA while the batch has at most one channel per SM, two otherwise);
       beyond: 15 channels per CTA (wpc), where the per-sample instruction count decides.  The one-CTA-per-channel pipeline (pipe) and the
       serial kernel serve the modem configurations the duo kernel is not eligible for (N_dft != 128, nin-dependent window count). */
    const bool wpc = h->can_wpc && !h->force_duo && (h->force_wpc || h->nchan > h->wpc_thresh);
    const bool duo = !wpc && h->can_duo && (h->force_duo || (h->duo_auto && h->nchan <= 4 * h->nsm));
    if (duo) {
        void *args[] = {(void *)&kp};
        kp.d_G = (!h->force_duo && h->nchan <= h->nsm) ? 1 : DUO_G;   /* at most one channel per SM: spread them, one channel per CTA (19.9 vs 19.0 Gsps at 148) */
        if (const char *t = getenv("B200FSK_DUO_G")) { const int v = atoi(t); if (v >= 1 && v <= DUO_G) kp.d_G = v; }   /* experiments: 1 = one channel per CTA */
        CK(cudaLaunchKernel(duo_kernel_fn(kp.M, kp.fmt), dim3((h->nchan + kp.d_G - 1) / kp.d_G), dim3(DUO_THREADS), args, h->smem_duo, s));
        h->last_kind = 3;
    } else if (wpc) {
        /* channels per CTA: fill whole waves of one CTA per SM */
        int G = h->wpc_G;
        const int slots = h->nsm * WPC_CTAS, per_wave = G * slots, waves = (h->nchan + per_wave - 1) / per_wave;
        const int Gfit = (h->nchan + waves * slots - 1) / (waves * slots);
        if (Gfit < G) G = Gfit;
        if (G < 1) G = 1;
        kp.w_G = G;
        const size_t smem = (size_t)kp.w_chan0 + (size_t)G * kp.w_chan_stride;
        if (h->l2_persist) {
            cudaStreamAttrValue av; memset(&av, 0, sizeof av);
            av.accessPolicyWindow.base_ptr = h->d_fint; av.accessPolicyWindow.num_bytes = h->l2_window; av.accessPolicyWindow.hitRatio = h->l2_ratio;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) { cudaGetLastError(); h->l2_persist = 0; }
        }
        void *args[] = {(void *)&kp};
        CK(cudaLaunchKernel(wpc_kernel_fn(kp.M, kp.fmt), dim3((h->nchan + G - 1) / G), dim3((G + 1) * 32), args, smem, s));
        h->last_kind = 2;
    } else if (h->use_pipe) {
        h->last_kind = 1;
        /* one CTA per SM: the oscillator's stores are worth software-pipelining (15.4 vs 12.9 Gsps at 128 channels); with two CTAs per
           SM the second CTA's shared-memory traffic makes it a loss (18.0 vs 20.8 at 256): profiles/r02_notes.md */
        kp.osc_pipelined = h->nchan <= h->nsm ? 1 : 0;
        if (const char *e = getenv("B200FSK_OSC_PIPELINED")) kp.osc_pipelined = atoi(e);
        void *args[] = {(void *)&kp};
        CK(cudaLaunchKernel(pipe_kernel_fn(kp.M, kp.fmt, kp.osc_pipelined), dim3(h->nchan), dim3(PIPE_THREADS), args, h->smem_pipe, s));
    } else { fsk_demod_kernel<<<h->nchan, FSK_THREADS, h->smem, s>>>(kp); h->last_kind = 0; }
    CK(cudaGetLastError());
    if (time_it) { CK(cudaEventRecord(h->ev1, s)); h->timed = true; }
    h->launches++;
    return 0;
}

/* the throughput kernel's watchdog: a CTA that waited ~4 s for its own oscillator gives up and raises this flag */
static int check_abort(b200fsk *h)
{
    if (!h->d_abort) return 0;
    int a = 0;
    CK(cudaMemcpy(&a, h->d_abort, sizeof a, cudaMemcpyDeviceToHost));
    if (a) { cudaMemset(h->d_abort, 0, sizeof a); return fail(B200FSK_ECUDA, "fsk_wpc_kernel watchdog: a channel group stalled"); }
    return 0;
}

extern "C" int b200fsk_demod_device(b200fsk_t *h, const void *d_in, int fmt, long in_stride, const int *d_in_off,
                                    long nsamp, int max_frames, void *d_out, long out_stride,
                                    int soft, int softinv, void *stream)
{
    if (!h || !d_in || !d_out) return fail(B200FSK_EINVAL, "null pointer");
    if (fmt < 0 || fmt > 3 || nsamp < 0) return fail(B200FSK_ERANGE, "fmt/nsamp");
    if (nsamp > 0x7fff0000L) return fail(B200FSK_ERANGE, "nsamp %ld per channel exceeds the kernels' 31-bit sample index: split the stream", nsamp);
    CK(cudaSetDevice(h->device));
    int rc = flush_pending(h); if (rc) return rc;
    const long long maxf = max_frames > 0 ? max_frames : b200fsk_max_frames(h, nsamp);
    if (out_stride < maxf * h->Nbits) return fail(B200FSK_ERANGE, "out_stride %ld < %lld (use b200fsk_max_frames)", out_stride, maxf * h->Nbits);
    FskKParams kp = h->kp;
    if (h->stats_on) { rc = ensure_recs(h, maxf); if (rc) return rc; }
    kp.in = d_in; kp.fmt = fmt; kp.in_stride = in_stride; kp.in_off = (int *)d_in_off; kp.nsamp = nsamp; kp.max_frames = max_frames;
    kp.carry_dst = 0; kp.carry_cap = 0; kp.in_off_next = 0;
    kp.out = d_out; kp.out_stride = out_stride; kp.soft = soft ? 1 : 0; kp.softinv = softinv ? 1 : 0; kp.frame_base = 0;
    kp.recs = h->d_recs; kp.rec_stride = h->rec_cap; kp.write_recs = h->stats_on;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    rc = launch(h, kp, s, true); if (rc) return rc;
    if (h->stats_on) {
        if (s != h->stream) CK(cudaStreamSynchronize(s));
        h->recs_pending = true; h->pending_base0 = 0;
    }
    return B200FSK_OK;
}

extern "C" int b200fsk_counts(b200fsk_t *h, int *frames, long *consumed)
{
    if (!h) return fail(B200FSK_EINVAL, "null");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    { const int rc = check_abort(h); if (rc) return rc; }
    if (frames) CK(cudaMemcpy(frames, h->d_nframes, sizeof(int) * h->nchan, cudaMemcpyDeviceToHost));
    if (consumed) {
        std::vector<long long> t(h->nchan);
        CK(cudaMemcpy(t.data(), h->d_consumed, sizeof(long long) * h->nchan, cudaMemcpyDeviceToHost));
        for (int c = 0; c < h->nchan; ++c) consumed[c] = (long)t[c];
    }
    return B200FSK_OK;
}

extern "C" int b200fsk_nin(b200fsk_t *h, int *nin_out)
{
    if (!h || !nin_out) return fail(B200FSK_EINVAL, "null");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(nin_out, h->d_nin, sizeof(int) * h->nchan, cudaMemcpyDeviceToHost));
    return B200FSK_OK;
}

extern "C" int b200fsk_clear_estimators(b200fsk_t *h)
{
    if (!h) return fail(B200FSK_EINVAL, "null");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(h->d_Sf, 0, sizeof(float) * (size_t)h->nchan * h->Ndft));          /* fsk.c:984-986 */
    std::vector<int> nin(h->nchan, h->N);                                              /* fsk.c:988 */
    CK(cudaMemcpy(h->d_nin, nin.data(), sizeof(int) * h->nchan, cudaMemcpyHostToDevice));
    return B200FSK_OK;
}

static int esize_of(int fmt) { return fmt == 1 ? 8 : (fmt == 0 ? 4 : 2); }

extern "C" int b200fsk_push_host(b200fsk_t *h, const void *h_in, int fmt, long in_stride, long nsamp, int max_frames,
                                 void *h_out, long out_stride, int soft, int softinv, int *frames)
{
    if (!h || (!h_in && nsamp > 0) || !h_out) return fail(B200FSK_EINVAL, "null pointer");
    if (fmt < 0 || fmt > 3 || nsamp < 0) return fail(B200FSK_ERANGE, "fmt/nsamp");
    CK(cudaSetDevice(h->device));
    int rc = flush_pending(h); if (rc) return rc;
    const int es = esize_of(fmt), nchan = h->nchan, oes = soft ? 4 : 1;
    int max_tail = 0; for (int c = 0; c < nchan; ++c) max_tail = h->h_tail[c] > max_tail ? h->h_tail[c] : max_tail;
    if (h->d_stage[0] && h->stage_es != es && max_tail > 0) return fail(B200FSK_EINVAL, "input format changed with samples still carried");
    /* Long uncapped pushes are cut into time chunks so that the H2D copy of chunk t+1 overlaps the kernel of chunk t
       (the kernel is latency-bound per channel, so chunking along time costs nothing). */
    int nchunk = 1;
    if (max_frames <= 0 && nsamp >= 16L * h->ncap) { nchunk = 16; while (nchunk > 1 && nsamp / nchunk < 8L * h->ncap) nchunk >>= 1; }
    const long chunk = (nsamp + nchunk - 1) / nchunk;
    const long long need_carry = (max_frames > 0 ? (long long)max_tail + nsamp : 0) + h->ncap + 8;
    if (!h->d_stage[0] || h->stage_es != es || h->carry_cap < need_carry || h->stage_cap < (long long)h->carry_cap + chunk) {
        const int new_carry = (int)(need_carry > h->carry_cap ? need_carry : h->carry_cap);
        const long long new_cap = (long long)new_carry + chunk;
        void *ns[2] = {0, 0};
        CK(cudaMalloc(&ns[0], (size_t)nchan * new_cap * es)); CK(cudaMalloc(&ns[1], (size_t)nchan * new_cap * es));
        CK(cudaDeviceSynchronize());
        if (h->d_stage[0] && max_tail > 0)      /* move the carried tails (right-aligned before carry_cap of the current buffer) */
            for (int c = 0; c < nchan; ++c)
                if (h->h_tail[c])
                    CK(cudaMemcpy((char *)ns[h->cur] + ((size_t)c * new_cap + new_carry - h->h_tail[c]) * es,
                                  (char *)h->d_stage[h->cur] + ((size_t)c * h->stage_cap + h->carry_cap - h->h_tail[c]) * es,
                                  (size_t)h->h_tail[c] * es, cudaMemcpyDeviceToDevice));
        for (int i = 0; i < 2; ++i) { if (h->d_stage[i]) cudaFree(h->d_stage[i]); h->d_stage[i] = ns[i]; }
        h->stage_cap = new_cap; h->stage_es = es; h->carry_cap = new_carry;
        std::vector<int> off(2 * nchan);
        for (int c = 0; c < nchan; ++c) off[h->cur * nchan + c] = off[(h->cur ^ 1) * nchan + c] = new_carry - h->h_tail[c];
        CK(cudaMemcpy(h->d_off, off.data(), sizeof(int) * 2 * nchan, cudaMemcpyHostToDevice));
    }
    const long long maxf = b200fsk_push_frames_bound(h, nsamp, max_frames);
    if (out_stride < maxf * h->Nbits) return fail(B200FSK_ERANGE, "out_stride %ld < %lld", out_stride, maxf * h->Nbits);
    if (!h->d_out || h->out_es != oes || h->out_cap < maxf * h->Nbits) {
        CK(cudaDeviceSynchronize());
        if (h->d_out) cudaFree(h->d_out);
        h->out_cap = maxf * h->Nbits + 1024; h->out_es = oes;
        CK(cudaMalloc(&h->d_out, (size_t)nchan * h->out_cap * oes));
    }
    if (h->stats_on) { rc = ensure_recs(h, maxf); if (rc) return rc; }
    cudaStream_t s = h->stream, cs = h->copy_stream;
    CK(cudaMemsetAsync(h->d_fbase, 0, sizeof(int) * nchan, s));
    CK(cudaEventRecord(h->ev0, s));
    long long rows_sent = 0;
    for (int t = 0; t < nchunk; ++t) {
        const long s0 = (long)t * chunk, len = (s0 + chunk <= nsamp ? chunk : (nsamp > s0 ? nsamp - s0 : 0));
        const int b = h->cur;
        if (len > 0) {
            CK(cudaStreamWaitEvent(cs, h->ev_k[b], 0));            /* the kernel that last read this buffer has finished */
            CK(cudaMemcpy2DAsync((char *)h->d_stage[b] + (size_t)h->carry_cap * es, (size_t)h->stage_cap * es,
                                 (const char *)h_in + (size_t)s0 * es, (size_t)in_stride * es, (size_t)len * es, nchan, cudaMemcpyHostToDevice, cs));
            CK(cudaEventRecord(h->ev_h2d[b], cs));
            CK(cudaStreamWaitEvent(s, h->ev_h2d[b], 0));
        }
        FskKParams kp = h->kp;
        kp.in = h->d_stage[b]; kp.fmt = fmt; kp.in_stride = h->stage_cap; kp.in_off = h->d_off + b * nchan; kp.nsamp = (long long)h->carry_cap + len;
        kp.max_frames = max_frames; kp.carry_dst = h->d_stage[b ^ 1]; kp.carry_cap = h->carry_cap; kp.in_off_next = h->d_off + (b ^ 1) * nchan;
        kp.out = h->d_out; kp.out_stride = h->out_cap; kp.soft = soft ? 1 : 0; kp.softinv = softinv ? 1 : 0; kp.frame_base = h->d_fbase;
        kp.recs = h->d_recs; kp.rec_stride = h->rec_cap; kp.write_recs = h->stats_on;
        rc = launch(h, kp, s, false); if (rc) return rc;
        CK(cudaEventRecord(h->ev_k[b], s));
        h->cur ^= 1;
        /* every channel has, after this chunk, at least floor((samples so far - nin_max)/nin_max) whole frames: ship those rows
           of the output now, overlapped with the following chunks (rows beyond are sent after the last kernel) */
        if (nchunk > 1 && !h->stats_on) {
            const long long fed = (long long)s0 + len;
            const long long lb = fed > 2LL * h->ncap ? (fed - h->ncap) / h->ncap - 1 : 0;
            if (lb > rows_sent) {
                CK(cudaStreamWaitEvent(h->out_stream, h->ev_k[b], 0));
                CK(cudaMemcpy2DAsync((char *)h_out + (size_t)rows_sent * h->Nbits * oes, (size_t)out_stride * oes,
                                     (char *)h->d_out + (size_t)rows_sent * h->Nbits * oes, (size_t)h->out_cap * oes,
                                     (size_t)(lb - rows_sent) * h->Nbits * oes, nchan, cudaMemcpyDeviceToHost, h->out_stream));
                rows_sent = lb;
            }
        }
    }
    CK(cudaEventRecord(h->ev1, s)); h->timed = true;
    CK(cudaMemcpyAsync(h->h_nframes.data(), h->d_nframes, sizeof(int) * nchan, cudaMemcpyDeviceToHost, s));
    std::vector<int> off(nchan);
    CK(cudaMemcpyAsync(off.data(), h->d_off + h->cur * nchan, sizeof(int) * nchan, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    int mx = 0;
    for (int c = 0; c < nchan; ++c) { h->h_tail[c] = h->carry_cap - off[c]; mx = h->h_nframes[c] > mx ? h->h_nframes[c] : mx; if (frames) frames[c] = h->h_nframes[c]; }
    if (mx > rows_sent) {
        CK(cudaMemcpy2DAsync((char *)h_out + (size_t)rows_sent * h->Nbits * oes, (size_t)out_stride * oes,
                             (char *)h->d_out + (size_t)rows_sent * h->Nbits * oes, (size_t)h->out_cap * oes,
                             (size_t)(mx - rows_sent) * h->Nbits * oes, nchan, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(h->out_stream));
    { const int rc2 = check_abort(h); if (rc2) return rc2; }
    if (mx > 0) {
        if (h->stats_on) {
            h->h_recs.resize((size_t)nchan * mx);
            CK(cudaMemcpy2DAsync(h->h_recs.data(), (size_t)mx * sizeof(FskFrameRec), h->d_recs, (size_t)h->rec_cap * sizeof(FskFrameRec),
                                 (size_t)mx * sizeof(FskFrameRec), nchan, cudaMemcpyDeviceToHost, s));
        }
        CK(cudaStreamSynchronize(s));
        if (h->stats_on) digest_records(h, h->h_recs.data(), mx, h->h_nframes.data(), 0);
    }
    return B200FSK_OK;
}

extern "C" int b200fsk_enable_stats(b200fsk_t *h, int on)
{
    if (!h) return fail(B200FSK_EINVAL, "null");
    int rc = flush_pending(h); if (rc) return rc;
    h->stats_on = on ? 1 : 0;
    return B200FSK_OK;
}

extern "C" int b200fsk_get_stats(b200fsk_t *h, int c, b200fsk_stats_t *st)
{
    if (!h || !st) return fail(B200FSK_EINVAL, "null");
    if (c < 0 || c >= h->nchan) return fail(B200FSK_ERANGE, "channel");
    CK(cudaSetDevice(h->device));
    int rc = flush_pending(h); if (rc) return rc;
    CK(cudaDeviceSynchronize());
    memset(st, 0, sizeof *st);
    const int M = h->cfg.M, P = h->cfg.P;
    long long tf = 0, ts = 0; float ppm = 0, nrt = 0; int nin = 0;
    CK(cudaMemcpy(&tf, h->d_totf + c, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&ts, h->d_tots + c, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&ppm, h->d_ppm + c, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&nrt, h->d_nrt + c, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&nin, h->d_nin + c, 4, cudaMemcpyDeviceToHost));
    st->frames = tf; st->samples = ts; st->ppm = ppm; st->norm_rx_timing = nrt; st->rx_timing = nrt * (float)P; st->nin = nin;
    st->snr_est = h->snr_est[c]; st->EbNodB = h->EbNodB[c];
    st->neyesamp = h->kp.neyesamp; st->neyetr = M * h->kp.eye_traces;       /* stats_init, fsk.c:939-971 */
    if (h->have_rec[c]) {
        const FskFrameRec &r = h->last_rec[c];
        float fc_avg = 0.0, fc_tx = 0.0;                                      /* fsk.c:851-856 */
        for (int m = 0; m < M; ++m) {
            st->f_est[m] = h->h_fest[r.tabidx[m]];
            fc_avg += st->f_est[m] / M; fc_tx += (h->cfg.f1_tx + m * h->cfg.fs_tx) / M;
        }
        st->foff = fc_tx - fc_avg;
        std::vector<float> eye(8 * 160);
        CK(cudaMemcpy(eye.data(), h->d_eye + (size_t)c * 8 * 160, sizeof(float) * 8 * 160, cudaMemcpyDeviceToHost));
        float eye_max = 0;                                                    /* normalise_eye, fsk.c:892-903 */
        for (int i = 0; i < st->neyetr; ++i) for (int j = 0; j < st->neyesamp; ++j) if (fabsf(eye[i * 160 + j]) > eye_max) eye_max = fabsf(eye[i * 160 + j]);
        for (int i = 0; i < st->neyetr; ++i) for (int j = 0; j < st->neyesamp; ++j) st->rx_eye[i][j] = eye[i * 160 + j] / eye_max;
    }
    return B200FSK_OK;
}

extern "C" int b200fsk_get_Sf(b200fsk_t *h, int c, float *Sf_out)
{
    if (!h || !Sf_out) return fail(B200FSK_EINVAL, "null");
    if (c < 0 || c >= h->nchan) return fail(B200FSK_ERANGE, "channel");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(Sf_out, h->d_Sf + (size_t)c * h->Ndft, sizeof(float) * h->Ndft, cudaMemcpyDeviceToHost));
    return B200FSK_OK;
}

#ifdef FSK_PROFILE
/* profiling build only: attach a zeroed [nchan][16] cycle-counter array and read it back */
extern "C" int b200fsk_prof_read(b200fsk_t *h, long long *out)
{
    if (!h->kp.prof) { long long *d = 0; CK(zalloc(&d, (size_t)h->nchan * 16)); h->kp.prof = d; return 0; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, h->kp.prof, sizeof(long long) * h->nchan * 16, cudaMemcpyDeviceToHost));
    CK(cudaMemset(h->kp.prof, 0, sizeof(long long) * h->nchan * 16));
    return 0;
}
#endif

extern "C" int b200fsk_set_lock_nin(b200fsk_t *h, int on)
{
    if (!h) return fail(B200FSK_EINVAL, "null");
    h->kp.lock_nin = on ? 1 : 0;
    return B200FSK_OK;
}

extern "C" int b200fsk_get_eye_raw(b200fsk_t *h, int c, float *eye)
{
    if (!h || !eye) return fail(B200FSK_EINVAL, "null");
    if (c < 0 || c >= h->nchan) return fail(B200FSK_ERANGE, "channel");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(eye, h->d_eye + (size_t)c * 8 * 160, sizeof(float) * 8 * 160, cudaMemcpyDeviceToHost));
    return B200FSK_OK;
}

extern "C" long b200fsk_launch_count(const b200fsk_t *h) { return h ? h->launches : 0; }
extern "C" int b200fsk_last_kernel_kind(const b200fsk_t *h) { return h ? h->last_kind : -1; }

extern "C" float b200fsk_last_kernel_ms(b200fsk_t *h)
{
    if (!h || !h->timed) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0f;
    return ms;
}
