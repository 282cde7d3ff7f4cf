/* softsync.cu -- soft-bit header sync + RS41 frame assembly on the GPU, part of libb200fsk.so.
 *
 * Reference: find_softbinhead / corr_softhdb (demod/mod/demod_mod.c:1692-1761) and the --softin frame loop of rs41mod
 * (demod/mod/rs41mod.c:2875-2960).  In the reference chain `fsk_demod -s | rs41mod --softin -i` every soft bit crosses a
 * pipe; here the soft decisions that fsk_demod_*_kernel left in HBM are scanned in place and only frames come back:
 * one warp per channel, 32 candidate header positions per iteration (64-tap normalised correlation each, double
 * accumulation like the reference), the first hit wins, the following 4080 soft bits are packed into 510 bytes
 * (little-endian bits, XOR de-whitening mask, per-byte minimum-magnitude soft score), then the scan resumes with the
 * ring exactly as the reference leaves it (frame payload bits never enter the header ring).
 *
 * with ecc = 1|2 the warp also runs rs41mod's Reed-Solomon stage on the frame while it is still in shared
 * memory (rs41_ecc, demod/mod/rs41mod.c:1703-1976 for --ecc / --ecc2; decoder demod/mod/bch_ecc_mod.c:877-966): two
 * interleaved RS(255,231) words over GF(2^8)/0x11D.  Syndromes (24 lanes, Horner), root search and Forney values (all lanes)
 * are warp-parallel; the short Euclidean key-equation recursion runs on one lane.  The correction is applied in the running
 * frame buffer like gpx->frame, then the six standard blocks' CRC-16 are checked (check_CRC, rs41mod.c:284-319).
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
    const float *soft; long long stride; const int *nbits_c; int nbits; int nchan, softinv, inv, maxpc, ecc, flush;
    b200fsk_rs41_frame_t *frames; int *counts; struct SsState *state;
};

/* ---- GF(2^8) f = 0x11D, alpha = 2 (bch_ecc_mod.c:55-59): exp table doubled so products need no modulo -------------------- */
#define RS_N 255
#define RS_R 24
#define RS_K 231
#define RS_T 12
#define RS_PLEN 32
struct GfTab { unsigned char ex[512]; unsigned char lg[256]; };
__device__ __forceinline__ unsigned gmul(const GfTab &g, unsigned a, unsigned b) { return (a && b) ? g.ex[g.lg[a] + g.lg[b]] : 0u; }
__device__ __forceinline__ unsigned ginv(const GfTab &g, unsigned a) { return a ? g.ex[255 - g.lg[a]] : 0u; }
__device__ __forceinline__ int pdeg(const unsigned char *p) { int d = RS_PLEN - 1; while (d >= 0 && p[d] == 0) --d; return d; }
__device__ __forceinline__ unsigned peval(const GfTab &g, const unsigned char *p, int n, unsigned x)
{
    unsigned y = 0;
    for (int i = n - 1; i >= 0; --i) y = gmul(g, y, x) ^ p[i];
    return y;
}

struct RsScratch { unsigned char S[RS_PLEN], r0[RS_PLEN], r1[RS_PLEN], s0[RS_PLEN], s1[RS_PLEN], q[RS_PLEN], t2[RS_PLEN], dLa[RS_PLEN]; int dL, dO, status; };

/* rs_decode (bch_ecc_mod.c:877-966, no erasures) by one warp on a codeword in shared memory.  Returns (uniformly) the number of
 * corrected symbols or -1/-2/-3 with cw untouched. */
__device__ int rs255_decode_warp(const GfTab &g, unsigned char *cw, RsScratch &w, int lane)
{
    unsigned sy = 0;
    if (lane < RS_R) {                                                       /* S_i = cw(alpha^i), i = 0..23 (:638-651) */
        const unsigned x = g.ex[lane];
        for (int n = RS_N - 1; n >= 0; --n) sy = gmul(g, sy, x) ^ cw[n];
        w.S[lane] = (unsigned char)sy;
    } else w.S[lane] = 0;
    if (!__ballot_sync(0xffffffffu, sy != 0)) return 0;
    __syncwarp();
    if (lane == 0) {
        /* key equation S*Lambda = Omega mod X^24: remainder sequence of (X^24, S) with cofactors (0, 1), stopped at deg r < t
           (polyGF_lfsr :547-582; its first step only swaps the start pair) */
        for (int i = 0; i < RS_PLEN; ++i) { w.r0[i] = 0; w.r1[i] = w.S[i]; w.s0[i] = 0; w.s1[i] = 0; }
        w.r0[RS_R] = 1; w.s1[0] = 1;
        while (pdeg(w.r1) >= RS_T) {
            const int dq = pdeg(w.r1); int dp = pdeg(w.r0);
            for (int i = 0; i < RS_PLEN; ++i) w.q[i] = 0;
            const unsigned lead = ginv(g, w.r1[dq]);
            while (dp >= dq) {
                const unsigned c = gmul(g, w.r0[dp], lead);
                w.q[dp - dq] = (unsigned char)c;
                for (int i = 0; i <= dq; ++i) w.r0[dp - i] ^= (unsigned char)gmul(g, w.r1[dq - i], c);
                dp = pdeg(w.r0);
            }
            for (int i = 0; i < RS_PLEN; ++i) { const unsigned char t = w.r0[i]; w.r0[i] = w.r1[i]; w.r1[i] = t; w.t2[i] = w.s0[i]; }
            const int ds = pdeg(w.s1), dqq = pdeg(w.q);
            for (int i = 0; i <= dqq; ++i) for (int j = 0; j <= ds; ++j) if (i + j < RS_PLEN) w.t2[i + j] ^= (unsigned char)gmul(g, w.q[i], w.s1[j]);
            for (int i = 0; i < RS_PLEN; ++i) { w.s0[i] = w.s1[i]; w.s1[i] = w.t2[i]; }
        }
        const int dL = pdeg(w.s1), dO = pdeg(w.r1);                          /* Lambda = s1, Omega = r1 */
        int st = 1;
        if (dO >= dL) st = -3;                                               /* :911-914 */
        else if (w.s1[0] == 0) st = -2;                                      /* :915-925 */
        else {
            const unsigned gi = ginv(g, w.s1[0]);
            for (int i = 0; i <= dL; ++i) w.s1[i] = (unsigned char)gmul(g, w.s1[i], gi);
            for (int i = 0; i <= dO; ++i) w.r1[i] = (unsigned char)gmul(g, w.r1[i], gi);
            for (int i = 0; i < RS_PLEN; ++i) w.dLa[i] = 0;
            for (int i = 1; i <= dL; ++i) if (i & 1) w.dLa[i - 1] = w.s1[i];  /* formal derivative (:584-594) */
        }
        w.dL = dL; w.dO = dO; w.status = st;
    }
    __syncwarp();
    if (w.status < 0) return w.status;
    const int dL = w.dL;
    unsigned roots = 0;                                                       /* lane tests x = lane+1, lane+33, ... (:933-946) */
    for (int k = 0; k < 8; ++k) { const int x = lane + 1 + 32 * k; if (x < 256 && peval(g, w.s1, dL + 1, x) == 0) roots |= 1u << k; }
    int nerr = __popc(roots);
    for (int o = 16; o; o >>= 1) nerr += __shfl_xor_sync(0xffffffffu, nerr, o);
    if (nerr < dL) return -1;                                                 /* :949 */
    for (int k = 0; k < 8; ++k) if (roots >> k & 1) {
        const unsigned x = lane + 1 + 32 * k, x1 = ginv(g, x);
        const unsigned wv = peval(g, w.r1, w.dO + 1, x), z = peval(g, w.dLa, dL + 1, x);
        const unsigned val = z ? gmul(g, x1, gmul(g, wv, ginv(g, z))) : 0u;   /* b == 0: Y = x^-1 Omega(x)/Lambda'(x) (:596-612) */
        cw[g.lg[x1] % 255] ^= (unsigned char)val;
    }
    __syncwarp();
    return nerr;
}

/* rs41_ecc (rs41mod.c:1703-1976), option.ecc = level (1|2), on the frame in shared memory */
__device__ int rs41_ecc_warp(const GfTab &g, unsigned char *frame, int frmlen, int level, unsigned char *cw1, unsigned char *cw2, RsScratch &w, int lane)
{
    for (int i = frmlen + lane; i < SS_FRAME_LEN; i += 32) frame[i] = 0;     /* 0-padding (:1726) */
    __syncwarp();
    for (int i = lane; i < RS_N; i += 32) {
        cw1[i] = i < RS_R ? frame[8 + i] : frame[56 + 2 * (i - RS_R)];
        cw2[i] = i < RS_R ? frame[8 + RS_R + i] : frame[56 + 2 * (i - RS_R) + 1];
    }
    __syncwarp();
    int e1 = rs255_decode_warp(g, cw1, w, lane); __syncwarp();
    int e2 = rs255_decode_warp(g, cw2, w, lane); __syncwarp();
    if (level >= 2 && (e1 < 0 || e2 < 0)) {                                   /* 2nd pass (:1738-1762) */
        if (lane == 0) {
            const int idpos[5] = {0x039, 0x065, 0x093, 0x0B5, 0x112}, idpck[5] = {0x7928, 0x7A2A, 0x7C1E, 0x7D59, 0x7B15};
            for (int k = 0; k < 5; ++k) { frame[idpos[k]] = (idpck[k] >> 8) & 0xFF; frame[idpos[k] + 1] = idpck[k] & 0xFF; }
        }
        __syncwarp();
        const unsigned char b = frame[SS_POS_FRAME - 1]; int ft = 0;
        for (int i = 0; i < 4; ++i) ft += ((b >> i) & 1) - ((b >> (i + 4)) & 1);
        if (ft < -2) { for (int i = SS_NDATA_LEN + 7 + lane; i < SS_FRAME_LEN - 2; i += 32) frame[i] = 0; }
        else {
            for (int i = SS_NDATA_LEN + lane; i < SS_FRAME_LEN; i += 32) frame[i] = 0;
            for (int i = 0x12B + 2 + lane; i < SS_NDATA_LEN - 2; i += 32) frame[i] = 0;
            if (lane == 0) { frame[0x12B] = 0x76; frame[0x12B + 1] = 0x11; frame[SS_NDATA_LEN - 2] = 0xEC; frame[SS_NDATA_LEN - 1] = 0xC7; }
        }
        __syncwarp();
        for (int i = RS_R + lane; i < RS_N; i += 32) { cw1[i] = frame[56 + 2 * (i - RS_R)]; cw2[i] = frame[56 + 2 * (i - RS_R) + 1]; }   /* parity part as pass 1 left it */
        __syncwarp();
        e1 = rs255_decode_warp(g, cw1, w, lane); __syncwarp();
        e2 = rs255_decode_warp(g, cw2, w, lane); __syncwarp();
    }
    for (int i = lane; i < RS_N; i += 32) {                                   /* write-back (:1952-1959) */
        if (i < RS_R) { frame[8 + i] = cw1[i]; frame[8 + RS_R + i] = cw2[i]; }
        else { frame[56 + 2 * (i - RS_R)] = cw1[i]; frame[56 + 2 * (i - RS_R) + 1] = cw2[i]; }
    }
    __syncwarp();
    if (e1 < 0 || e2 < 0) return -((e1 < 0 ? 1 : 0) | (e2 < 0 ? 2 : 0));
    return e1 + e2;
}

/* check_CRC (rs41mod.c:307-319) of block k of the standard frame layout; lanes 0..5 -> bit k of the result = CRC ok */
__device__ unsigned rs41_crc_mask_warp(const unsigned char *frame, int lane)
{
    const int bpos[6] = {0x039, 0x065, 0x093, 0x0B5, 0x112, 0x12B}, btyp[6] = {0x79, 0x7A, 0x7C, 0x7D, 0x7B, 0x76};
    int ok = 0;
    if (lane < 6) {
        const int pos = bpos[lane];
        if (frame[pos] == btyp[lane]) {
            const int len = frame[pos + 1];
            if (pos + len + 4 <= SS_FRAME_LEN) {
                int rem = 0xFFFF;
                for (int i = 0; i < len; ++i) {
                    rem ^= frame[pos + 2 + i] << 8;
                    for (int j = 0; j < 8; ++j) rem = (rem & 0x8000) ? ((rem << 1) ^ 0x1021) & 0xFFFF : (rem << 1) & 0xFFFF;
                }
                ok = rem == (frame[pos + 2 + len] | (frame[pos + 3 + len] << 8));
            }
        }
    }
    return __ballot_sync(0xffffffffu, ok) & 0x3Fu;
}

/* Per-channel state that survives between pushes of one stream (all of it is what the reference keeps in its process: the header ring
 * hdb.sbuf, gpx->frame / frm_bytescore, the position inside a frame, plus the <8 soft bits of an unfinished byte). */
struct SsState {
    int mode;                      /* 0: searching for a header, 1: inside a frame */
    int byte_count, ncarry, nframes_total;
    long long consumed;            /* soft bits handed in by earlier pushes */
    long long first_bit; float mv; /* of the frame in progress */
    float ring[64];                /* last 63 scanned bits (search) / the ring as the hit left it (in frame) */
    float carry[8];                /* bits of an incomplete byte, logically in front of the next push */
    unsigned char frame[SS_FRAME_LEN + 2];
    float score[SS_FRAME_LEN];
};

__global__ void __launch_bounds__(SS_WARPS * 32) rs41_framesync_kernel(const SsParams p)
{
    __shared__ float s_w[SS_WARPS][96];
    __shared__ unsigned char s_frame[SS_WARPS][SS_FRAME_LEN + 2];
    __shared__ float s_score[SS_WARPS][SS_FRAME_LEN];
    __shared__ float s_carry[SS_WARPS][8];
    __shared__ GfTab s_gf;
    __shared__ unsigned char s_cw[SS_WARPS][2][256];
    __shared__ RsScratch s_rs[SS_WARPS];
    if (p.ecc > 0) {                                                          /* exp / log tables (GF_genTab, bch_ecc_mod.c:136-170) */
        if (threadIdx.x == 0) {
            unsigned b = 1;
            for (int i = 0; i < 255; ++i) { s_gf.ex[i] = s_gf.ex[i + 255] = (unsigned char)b; s_gf.lg[b] = (unsigned char)i; b <<= 1; if (b & 0x100) b ^= 0x11D; }
            s_gf.ex[510] = s_gf.ex[0]; s_gf.ex[511] = s_gf.ex[1]; s_gf.lg[0] = 0;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, c = blockIdx.x * SS_WARPS + wp;
    if (c >= p.nchan) return;
    float *w = s_w[wp]; unsigned char *frame = s_frame[wp]; float *score = s_score[wp]; float *carry = s_carry[wp];
    SsState &st = p.state[c];
    const float *soft = p.soft + (long long)c * p.stride;
    /* this push = the carried bits of an unfinished byte followed by the new soft bits */
    const int ncarry = st.ncarry;
    const int nb = ncarry + (p.nbits_c ? p.nbits_c[c] : p.nbits);
    const long long base = st.consumed - ncarry;                              /* stream index of local bit 0 */
    for (int i = lane; i < 96; i += 32) w[i] = i < 63 ? st.ring[i] : 0.0f;   /* first push: zeros = calloc'ed hdb.sbuf (rs41mod.c:2865) */
    for (int i = lane; i < SS_FRAME_LEN; i += 32) { frame[i] = st.frame[i]; score[i] = st.score[i]; }
    if (lane < 8) carry[lane] = st.carry[lane];
    __syncwarp();
    auto rd = [&](int i) -> float { const float v = i < ncarry ? carry[i] : __ldg(soft + (i - ncarry)); return p.softinv ? -v : v; };
    int pos = 0, nf = 0, mode = st.mode, byte_count = st.byte_count;
    long long first_g = st.first_bit; float mvf = st.mv;
    bool ended = false;
    while (true) {
        if (mode == 0) {                                                      /* ---- find_softbinhead (demod_mod.c:1741-1761) ---- */
            if (pos >= nb) break;
            const int ncand = min(32, nb - pos);
            if (lane < ncand) w[63 + lane] = rd(pos + lane);
            __syncwarp();
            float mv = 0.0f; int hit = 0;
            if (lane < ncand) {                                               /* corr_softhdb (demod_mod.c:1692-1716) */
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
            const int adv = m ? (__ffs(m) - 1) + 1 : ncand;                    /* bits that entered the ring */
            {   /* slide: the ring keeps the last 63 scanned bits; on a hit it ends at the hit bit (payload never enters it) */
                const float t0 = w[adv + lane], t1 = (lane + 32 < 63) ? w[adv + lane + 32] : 0.0f;
                __syncwarp();
                w[lane] = t0; if (lane + 32 < 63) w[lane + 32] = t1;
                __syncwarp();
            }
            pos += adv;
            if (!m) continue;
            mvf = __shfl_sync(0xffffffffu, mv, adv - 1);
            first_g = base + pos; byte_count = SS_FRAMESTART; mode = 1;
        }
        /* ---- frame assembly (rs41mod.c:2896-2956): whole bytes only; <8 left-over bits wait for the next push ---- */
        const int nby = min(SS_FRAME_LEN - byte_count, (nb - pos) / 8);
        for (int k = lane; k < nby; k += 32) {
            const int bc = byte_count + k;
            int byte = 0; float mn = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float s = rd(pos + 8 * k + q);
                int bit = s >= 0.0f; float sb = s;
                if (p.inv) { bit ^= 1; sb = -sb; }
                if (q == 0 || fabsf(sb) < fabsf(mn)) mn = sb;
                byte |= bit << q;                                             /* bits2byte: little endian (rs41mod.c:224-234) */
            }
            frame[bc] = (unsigned char)(byte ^ c_mask[bc & 63]); score[bc] = mn;
        }
        __syncwarp();
        byte_count += nby; pos += 8 * nby;
        if (byte_count < SS_FRAME_LEN) {                                      /* out of bits inside the frame */
            if (!p.flush) break;                                              /* stream continues: resume here on the next push */
            ended = true;                                                     /* EOF: an incomplete byte is dropped like the reference's bit loop does */
            if (byte_count < SS_POS_GPS1) { for (int i = byte_count + lane; i < SS_FRAME_LEN; i += 32) frame[i] = 0; __syncwarp(); }   /* print_frame :2481-2484 */
        }
        int flen, ec = 0; unsigned crc = 0;
        {
            const unsigned char b = frame[SS_POS_FRAME - 1]; int ft = 0;
            for (int i = 0; i < 4; ++i) ft += ((b >> i) & 1) - ((b >> (i + 4)) & 1);       /* frametype (rs41mod.c:407-415) */
            flen = ft >= 0 ? SS_NDATA_LEN : SS_FRAME_LEN;
        }
        if (p.ecc > 0) {                                                      /* print_frame: ec = rs41_ecc(gpx, len) (rs41mod.c:2525-2527) */
            ec = rs41_ecc_warp(s_gf, frame, flen, p.ecc, s_cw[wp][0], s_cw[wp][1], s_rs[wp], lane);
            crc = rs41_crc_mask_warp(frame, lane);
        }
        if (nf < p.maxpc) {
            b200fsk_rs41_frame_t *o = p.frames + (long long)c * p.maxpc + nf;
            for (int i = lane; i < SS_FRAME_LEN; i += 32) { o->bytes[i] = frame[i]; o->bytescore[i] = score[i]; }
            if (lane == 0) { o->chan = c; o->len = flen; o->nbytes_read = byte_count; o->first_bit = first_g; o->score = mvf; o->ec = ec; o->crc_mask = (int)crc; }
        }
        ++nf; mode = 0;
        if (ended) { pos = nb; break; }                                        /* EOF inside the frame ends the run (rs41mod.c main loop) */
    }
    /* save the stream state */
    const int left = (mode == 1 && !ended) ? nb - pos : 0;                     /* < 8 bits of an unfinished byte */
    float cv = 0.0f;
    if (lane < left) { const float v = pos + lane < ncarry ? carry[pos + lane] : __ldg(soft + (pos + lane - ncarry)); cv = v; }   /* raw (un-negated) value */
    for (int i = lane; i < 63; i += 32) st.ring[i] = w[i];
    for (int i = lane; i < SS_FRAME_LEN; i += 32) { st.frame[i] = frame[i]; st.score[i] = score[i]; }
    __syncwarp();
    if (lane < 8) st.carry[lane] = cv;
    if (lane == 0) {
        st.mode = mode; st.byte_count = byte_count; st.ncarry = left; st.consumed = base + nb; st.first_bit = first_g; st.mv = mvf;
        st.nframes_total += nf; p.counts[c] = nf;
    }
}

static thread_local char g_sserr[256];
static thread_local float g_ss_ms = -1.0f;
extern "C" float b200fsk_framesync_last_kernel_ms(void) { return g_ss_ms; }
extern "C" const char *b200fsk_framesync_last_error(void) { return g_sserr; }
#define SCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_sserr, sizeof g_sserr, "%s: %s", #call, cudaGetErrorString(e_)); return B200FSK_ECUDA; } } while (0)

struct b200fsk_rs41sync { int nchan, softinv, inv, ecc, device; SsState *d_state; b200fsk_rs41_frame_t *d_fr; int *d_cnt; int cap_pc; };

/* __constant__ symbols exist once PER DEVICE: upload the 128 bytes on every call for the device that is current (cudaSetDevice by the
   caller).  A once-per-process flag left every GPU but the first with all-zero header bits and de-whitening mask, and raced between
   threads. */
static int ss_init_tables(void)
{
    unsigned char hb[64]; for (int i = 0; i < 64; ++i) hb[i] = RS41_HDR[i] & 1;
    SCK(cudaMemcpyToSymbol(c_hdr, hb, 64)); SCK(cudaMemcpyToSymbol(c_mask, RS41_MASK, 64));
    return B200FSK_OK;
}

extern "C" int b200fsk_rs41sync_create(b200fsk_rs41sync_t **out, int nchan, int softinv, int inv, int ecc, int device)
{
    if (!out || nchan <= 0 || ecc < 0 || ecc > 2) return B200FSK_EINVAL;
    SCK(cudaSetDevice(device));
    if (int rc = ss_init_tables()) return rc;
    b200fsk_rs41sync *h = new b200fsk_rs41sync();
    h->nchan = nchan; h->softinv = softinv ? 1 : 0; h->inv = inv ? 1 : 0; h->ecc = ecc; h->device = device; h->d_fr = 0; h->d_cnt = 0; h->cap_pc = 0;
    std::vector<SsState> st(nchan);
    memset(st.data(), 0, sizeof(SsState) * nchan);
    for (int c = 0; c < nchan; ++c) memcpy(st[c].frame, RS41_HDR_BYTES, 8);                       /* rs41mod.c:2766 */
    SCK(cudaMalloc(&h->d_state, sizeof(SsState) * nchan));
    SCK(cudaMemcpy(h->d_state, st.data(), sizeof(SsState) * nchan, cudaMemcpyHostToDevice));
    SCK(cudaMalloc(&h->d_cnt, sizeof(int) * nchan));
    *out = h;
    return B200FSK_OK;
}

extern "C" void b200fsk_rs41sync_destroy(b200fsk_rs41sync_t *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->d_state) cudaFree(h->d_state); if (h->d_fr) cudaFree(h->d_fr); if (h->d_cnt) cudaFree(h->d_cnt);
    delete h;
}

extern "C" int b200fsk_rs41sync_push_device(b200fsk_rs41sync_t *h, const float *d_soft, long stride, const int *d_nbits, int nbits, int flush,
                                            b200fsk_rs41_frame_t *h_frames, int max_frames, int *nframes, void *stream)
{
    if (!h || (!d_soft && nbits > 0) || !h_frames || !nframes || nbits < 0) return B200FSK_EINVAL;
    SCK(cudaSetDevice(h->device));
    const int nchan = h->nchan;
    const int maxpc = nbits / (8 * (SS_FRAME_LEN - SS_FRAMESTART) + SS_HEADLEN) + 3;      /* + the frame in progress + an EOF fragment */
    if (maxpc > h->cap_pc) { if (h->d_fr) cudaFree(h->d_fr); SCK(cudaMalloc(&h->d_fr, (size_t)nchan * maxpc * sizeof(b200fsk_rs41_frame_t))); h->cap_pc = maxpc; }
    b200fsk_rs41_frame_t *d_fr = h->d_fr; int *d_cnt = h->d_cnt;
    SsParams p; p.soft = d_soft; p.stride = stride; p.nbits_c = d_nbits; p.nbits = nbits; p.nchan = nchan; p.softinv = h->softinv; p.inv = h->inv;
    p.maxpc = maxpc; p.ecc = h->ecc; p.flush = flush ? 1 : 0; p.frames = d_fr; p.counts = d_cnt; p.state = h->d_state;
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t e0, e1; SCK(cudaEventCreate(&e0)); SCK(cudaEventCreate(&e1));
    SCK(cudaEventRecord(e0, s));
    rs41_framesync_kernel<<<(nchan + SS_WARPS - 1) / SS_WARPS, SS_WARPS * 32, 0, s>>>(p);
    SCK(cudaEventRecord(e1, s));
    SCK(cudaGetLastError());
    std::vector<int> cnt(nchan);
    SCK(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int) * nchan, cudaMemcpyDeviceToHost, s));
    SCK(cudaStreamSynchronize(s));
    SCK(cudaEventElapsedTime(&g_ss_ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    /* frame records: one copy of the used prefix of every channel's slot row when that is most of the buffer, else per channel */
    int tot = 0, maxn = 0; long long sum = 0;
    for (int c = 0; c < nchan; ++c) { const int n = cnt[c] < maxpc ? cnt[c] : maxpc; sum += n; if (n > maxn) maxn = n; }
    if (sum > 0 && nchan > 8) {
        std::vector<b200fsk_rs41_frame_t> all((size_t)nchan * maxn);
        SCK(cudaMemcpy2D(all.data(), (size_t)maxn * sizeof(b200fsk_rs41_frame_t), d_fr, (size_t)maxpc * sizeof(b200fsk_rs41_frame_t),
                         (size_t)maxn * sizeof(b200fsk_rs41_frame_t), nchan, cudaMemcpyDeviceToHost));
        for (int c = 0; c < nchan; ++c) {
            const int n = cnt[c] < maxpc ? cnt[c] : maxpc;
            const int take = (tot + n <= max_frames) ? n : (max_frames - tot > 0 ? max_frames - tot : 0);
            if (take > 0) memcpy(h_frames + tot, all.data() + (size_t)c * maxn, (size_t)take * sizeof(b200fsk_rs41_frame_t));
            tot += take;
        }
    } else {
        for (int c = 0; c < nchan; ++c) {
            const int n = cnt[c] < maxpc ? cnt[c] : maxpc;
            const int take = (tot + n <= max_frames) ? n : (max_frames - tot > 0 ? max_frames - tot : 0);
            if (take > 0) SCK(cudaMemcpy(h_frames + tot, d_fr + (size_t)c * maxpc, (size_t)take * sizeof(b200fsk_rs41_frame_t), cudaMemcpyDeviceToHost));
            tot += take;
        }
    }
    *nframes = tot;
    return B200FSK_OK;
}

/* whole-buffer calls: one stream per call = create + push(flush) + destroy */
extern "C" int b200fsk_rs41_frames_device(const float *d_soft, long stride, const int *d_nbits, int nbits, int nchan, int softinv, int inv, int ecc,
                                          b200fsk_rs41_frame_t *h_frames, int max_frames, int *nframes, int device, void *stream)
{
    if (!d_soft || !h_frames || !nframes || nchan <= 0 || (nbits < 0) || ecc < 0 || ecc > 2) return B200FSK_EINVAL;
    b200fsk_rs41sync_t *h = 0;
    int rc = b200fsk_rs41sync_create(&h, nchan, softinv, inv, ecc, device);
    if (rc) return rc;
    rc = b200fsk_rs41sync_push_device(h, d_soft, stride, d_nbits, nbits, 1, h_frames, max_frames, nframes, stream);
    b200fsk_rs41sync_destroy(h);
    return rc;
}

extern "C" int b200fsk_rs41_framesync_device(const float *d_soft, long stride, const int *d_nbits, int nbits, int nchan, int softinv, int inv,
                                             b200fsk_rs41_frame_t *h_frames, int max_frames, int *nframes, int device, void *stream)
{
    return b200fsk_rs41_frames_device(d_soft, stride, d_nbits, nbits, nchan, softinv, inv, 0, h_frames, max_frames, nframes, device, stream);
}
