/* softsync.cu -- soft-bit header sync + RS41 frame assembly on the GPU, part of libb200fsk.so.
 *
 * Reference: find_softbinhead / corr_softhdb (demod/mod/demod_mod.c:1692-1761) and the --softin frame loop of rs41mod
 * (demod/mod/rs41mod.c:2875-2960).  In the reference chain `fsk_demod -s | rs41mod --softin -i` every soft bit crosses a
 * pipe; here the soft decisions that fsk_demod_*_kernel left in HBM are scanned in place and only frames come back:
 * one warp per channel, 32 candidate header positions per iteration (64-tap normalised correlation each, double
 * accumulation like the reference), the first hit wins, the following 4080 soft bits are packed into 510 bytes
 * (little-endian bits, XOR de-whitening mask, per-byte minimum-magnitude soft score), then the scan resumes with the
 * ring exactly as the reference leaves it (frame payload bits never enter the header ring).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "../../include/b200fsk.h"

#define SS_HEADLEN 64
#define SS_FRAMESTART 8
#define SS_FRAME_LEN 518
#define SS_NDATA_LEN 320
#define SS_POS_FRAME 0x039
#define SS_POS_GPS1 0x093
#define SS_WARPS 4

__constant__ unsigned char c_hdr[64];      /* header bits (rs41mod.c:170)  */
__constant__ unsigned char c_mask[64];     /* de-whitening (rs41mod.c:175) */
static const char RS41_HDR[] = "0000100001101101010100111000100001000100011010010100100000011111";
static const unsigned char RS41_HDR_BYTES[8] = {0x86, 0x35, 0xf4, 0x40, 0x93, 0xdf, 0x1a, 0x60};
static const unsigned char RS41_MASK[64] = {
    0x96, 0x83, 0x3E, 0x51, 0xB1, 0x49, 0x08, 0x98, 0x32, 0x05, 0x59, 0x0E, 0xF9, 0x44, 0xC6, 0x26,
    0x21, 0x60, 0xC2, 0xEA, 0x79, 0x5D, 0x6D, 0xA1, 0x54, 0x69, 0x47, 0x0C, 0xDC, 0xE8, 0x5C, 0xF1,
    0xF7, 0x76, 0x82, 0x7F, 0x07, 0x99, 0xA2, 0x2C, 0x93, 0x7C, 0x30, 0x63, 0xF5, 0x10, 0x2E, 0x61,
    0xD0, 0xBC, 0xB4, 0xB6, 0x06, 0xAA, 0xF4, 0x23, 0x78, 0x6E, 0x3B, 0xAE, 0xBF, 0x7B, 0x4C, 0xC1};

struct SsParams {
    const float *soft; long long stride; const int *nbits_c; int nbits; int nchan, softinv, inv, maxpc;
    b200fsk_rs41_frame_t *frames; int *counts;
};

__global__ void __launch_bounds__(SS_WARPS * 32) rs41_framesync_kernel(const SsParams p)
{
    __shared__ float s_w[SS_WARPS][96];
    __shared__ unsigned char s_frame[SS_WARPS][SS_FRAME_LEN + 2];
    __shared__ float s_score[SS_WARPS][SS_FRAME_LEN];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, c = blockIdx.x * SS_WARPS + wp;
    if (c >= p.nchan) return;
    float *w = s_w[wp]; unsigned char *frame = s_frame[wp]; float *score = s_score[wp];
    const float *soft = p.soft + (long long)c * p.stride;
    const int nbits = p.nbits_c ? p.nbits_c[c] : p.nbits;
    for (int i = lane; i < 96; i += 32) w[i] = 0.0f;                          /* hdb.sbuf = calloc (rs41mod.c:2865) */
    for (int i = lane; i < SS_FRAME_LEN; i += 32) { frame[i] = 0; score[i] = 0.0f; }
    __syncwarp();
    if (lane < 8) { const unsigned char hb[8] = {0x86, 0x35, 0xf4, 0x40, 0x93, 0xdf, 0x1a, 0x60}; frame[lane] = hb[lane]; }   /* rs41mod.c:2766 */
    __syncwarp();
    int pos = 0, nf = 0;
    while (pos < nbits) {
        const int ncand = min(32, nbits - pos);
        if (lane < ncand) { float s = __ldg(soft + pos + lane); w[63 + lane] = p.softinv ? -s : s; }
        __syncwarp();
        float mv = 0.0f; int hit = 0;
        if (lane < ncand) {                                                   /* corr_softhdb (demod_mod.c:1692-1716) */
            double sum = 0.0, normx = 0.0;
#pragma unroll 8
            for (int i = 0; i < SS_HEADLEN; ++i) {
                const float x = w[lane + i], y = c_hdr[i] ? 1.0f : -1.0f;
                sum += (double)(y * x); normx += (double)(x * x);
            }
            mv = (float)(sum / sqrt(normx * 64.0));
            hit = (fabsf(mv) > 0.7f) && !((double)mv * (0.5 - p.inv) < 0);    /* ths (rs41mod.c:2864) + polarity rule (:2889-2892) */
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (!m) {                                                             /* slide: the ring keeps the last 63 scanned bits */
            float t0 = w[ncand + lane], t1 = (lane + 32 < 63) ? w[ncand + lane + 32] : 0.0f;
            __syncwarp();
            w[lane] = t0; if (lane + 32 < 63) w[lane + 32] = t1;
            __syncwarp();
            pos += ncand;
            continue;
        }
        const int f = __ffs(m) - 1;
        const float mvf = __shfl_sync(0xffffffffu, mv, f);
        const int first = pos + f + 1;
        const int nread = min(8 * (SS_FRAME_LEN - SS_FRAMESTART), nbits - first), nby = nread / 8;
        for (int k = lane; k < nby; k += 32) {                                /* rs41mod.c:2896-2956 */
            const int bc = SS_FRAMESTART + k; const float *sp = soft + first + 8 * k;
            int byte = 0; float mn = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float s = __ldg(sp + q); if (p.softinv) s = -s;
                int bit = s >= 0.0f; float sb = s;
                if (p.inv) { bit ^= 1; sb = -sb; }
                if (q == 0 || fabsf(sb) < fabsf(mn)) mn = sb;
                byte |= bit << q;                                             /* bits2byte: little endian (rs41mod.c:224-234) */
            }
            frame[bc] = (unsigned char)(byte ^ c_mask[bc & 63]); score[bc] = mn;
        }
        __syncwarp();
        const int byte_count = SS_FRAMESTART + nby;
        if (byte_count < SS_POS_GPS1) { for (int i = byte_count + lane; i < SS_FRAME_LEN; i += 32) frame[i] = 0; __syncwarp(); }   /* print_frame :2481-2484 */
        if (nf < p.maxpc) {
            b200fsk_rs41_frame_t *o = p.frames + (long long)c * p.maxpc + nf;
            for (int i = lane; i < SS_FRAME_LEN; i += 32) { o->bytes[i] = frame[i]; o->bytescore[i] = score[i]; }
            if (lane == 0) {
                const unsigned char b = frame[SS_POS_FRAME - 1]; int ft = 0;
                for (int i = 0; i < 4; ++i) ft += ((b >> i) & 1) - ((b >> (i + 4)) & 1);   /* frametype (rs41mod.c:407-415) */
                o->chan = c; o->len = ft >= 0 ? SS_NDATA_LEN : SS_FRAME_LEN; o->nbytes_read = byte_count; o->first_bit = first; o->score = mvf;
            }
        }
        ++nf;
        {   /* the header ring is untouched by the payload: it still ends at the hit bit */
            float t0 = w[f + 1 + lane], t1 = (lane + 32 < 63) ? w[f + 1 + lane + 32] : 0.0f;
            __syncwarp();
            w[lane] = t0; if (lane + 32 < 63) w[lane + 32] = t1;
            __syncwarp();
        }
        pos = first + nread;
        if (pos >= nbits && byte_count < SS_FRAME_LEN) break;                 /* EOF inside the frame */
    }
    if (lane == 0) p.counts[c] = nf;
}

static thread_local char g_sserr[256];
extern "C" const char *b200fsk_framesync_last_error(void) { return g_sserr; }
#define SCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_sserr, sizeof g_sserr, "%s: %s", #call, cudaGetErrorString(e_)); return B200FSK_ECUDA; } } while (0)

extern "C" int b200fsk_rs41_framesync_device(const float *d_soft, long stride, const int *d_nbits, int nbits, int nchan, int softinv, int inv,
                                             b200fsk_rs41_frame_t *h_frames, int max_frames, int *nframes, int device, void *stream)
{
    if (!d_soft || !h_frames || !nframes || nchan <= 0 || (nbits < 0)) return B200FSK_EINVAL;
    SCK(cudaSetDevice(device));
    static bool init = false;
    if (!init) {
        unsigned char hb[64]; for (int i = 0; i < 64; ++i) hb[i] = RS41_HDR[i] & 1;
        SCK(cudaMemcpyToSymbol(c_hdr, hb, 64)); SCK(cudaMemcpyToSymbol(c_mask, RS41_MASK, 64));
        init = true;
    }
    (void)RS41_HDR_BYTES;
    const int maxpc = nbits / (8 * (SS_FRAME_LEN - SS_FRAMESTART) + SS_HEADLEN) + 2;
    b200fsk_rs41_frame_t *d_fr = 0; int *d_cnt = 0;
    SCK(cudaMalloc(&d_fr, (size_t)nchan * maxpc * sizeof(b200fsk_rs41_frame_t))); SCK(cudaMalloc(&d_cnt, sizeof(int) * nchan));
    SsParams p; p.soft = d_soft; p.stride = stride; p.nbits_c = d_nbits; p.nbits = nbits; p.nchan = nchan; p.softinv = softinv ? 1 : 0; p.inv = inv ? 1 : 0;
    p.maxpc = maxpc; p.frames = d_fr; p.counts = d_cnt;
    cudaStream_t s = (cudaStream_t)stream;
    rs41_framesync_kernel<<<(nchan + SS_WARPS - 1) / SS_WARPS, SS_WARPS * 32, 0, s>>>(p);
    SCK(cudaGetLastError());
    std::vector<int> cnt(nchan);
    SCK(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int) * nchan, cudaMemcpyDeviceToHost, s));
    SCK(cudaStreamSynchronize(s));
    int tot = 0;
    for (int c = 0; c < nchan; ++c) {
        const int n = cnt[c] < maxpc ? cnt[c] : maxpc;
        const int take = (tot + n <= max_frames) ? n : (max_frames - tot > 0 ? max_frames - tot : 0);
        if (take > 0) SCK(cudaMemcpy(h_frames + tot, d_fr + (size_t)c * maxpc, (size_t)take * sizeof(b200fsk_rs41_frame_t), cudaMemcpyDeviceToHost));
        tot += take;
    }
    *nframes = tot;
    cudaFree(d_fr); cudaFree(d_cnt);
    return B200FSK_OK;
}
