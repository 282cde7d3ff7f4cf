/* softsync_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the soft-bit consumer that follows fsk_demod in the auto_rx RS41 chain
 * (`fsk_demod -s ... | rs41mod --softin -i`, auto_rx/autorx/decode.py:901-909): the 64-bit header search on float soft
 * bits (find_softbinhead / corr_softhdb, demod/mod/demod_mod.c:1692-1761) and the frame assembly loop of rs41mod
 * (demod/mod/rs41mod.c:2875-2960: polarity rule, bit inversion, little-endian bytes, XOR de-whitening mask, per-byte
 * minimum soft score, frame length from the frame-type byte :407-415, :2483-2491).  SURVEY.md §8f-1.
 * Pinned by tests/test_softsync.py against `oracle/_ref/rs41mod -r --softin -i` (raw frame hex of frames the
 * Reed-Solomon stage leaves untouched); with ecc > 0 the Reed-Solomon stage of oracle/ecc_oracle.c runs in the same
 * frame buffer (tests/test_ecc.py).  Telemetry parsing is NOT restated.
 */
#include "oracle.h"
#include <math.h>
#include <string.h>

static const char RS41_HDR[] = "0000100001101101010100111000100001000100011010010100100000011111";   /* rs41mod.c:170 */
static const unsigned char RS41_HDR_BYTES[8] = {0x86, 0x35, 0xf4, 0x40, 0x93, 0xdf, 0x1a, 0x60};         /* rs41mod.c:172 */
static const unsigned char RS41_MASK[64] = {                                                                /* rs41mod.c:175-182 */
    0x96, 0x83, 0x3E, 0x51, 0xB1, 0x49, 0x08, 0x98, 0x32, 0x05, 0x59, 0x0E, 0xF9, 0x44, 0xC6, 0x26,
    0x21, 0x60, 0xC2, 0xEA, 0x79, 0x5D, 0x6D, 0xA1, 0x54, 0x69, 0x47, 0x0C, 0xDC, 0xE8, 0x5C, 0xF1,
    0xF7, 0x76, 0x82, 0x7F, 0x07, 0x99, 0xA2, 0x2C, 0x93, 0x7C, 0x30, 0x63, 0xF5, 0x10, 0x2E, 0x61,
    0xD0, 0xBC, 0xB4, 0xB6, 0x06, 0xAA, 0xF4, 0x23, 0x78, 0x6E, 0x3B, 0xAE, 0xBF, 0x7B, 0x4C, 0xC1};
#define HEADLEN 64
#define FRAMESTART 8
#define FRAME_LEN 518
#define NDATA_LEN 320
#define POS_FRAME 0x039
#define POS_GPS1 0x093

/* soft: nbits float soft decisions of one channel; softinv = --softinv (negate on read), inv = -i (gpx.option.inv).
 * Returns the number of frames; frames[k] receives length, position of the first payload bit, header score, bytes, scores. */
int sso_rs41(const float *soft, long nbits, int softinv, int inv, int ecc, sso_frame_t *frames, int maxframes)
{
    float sbuf[HEADLEN]; int bufpos = -1, nf = 0; long pos = 0;
    unsigned char frame[FRAME_LEN]; float bytescore[FRAME_LEN];
    memset(sbuf, 0, sizeof sbuf); memset(frame, 0, sizeof frame); memset(bytescore, 0, sizeof bytescore);
    memcpy(frame, RS41_HDR_BYTES, 8);                                       /* rs41mod.c:2766 */
    while (1) {
        /* find_softbinhead (demod_mod.c:1741-1761) */
        int found = 0; float mv = 0;
        while (pos < nbits) {
            float s = soft[pos++]; if (softinv) s = -s;
            bufpos = (bufpos + 1) % HEADLEN; sbuf[bufpos] = s;
            double sum = 0, normx = 0, normy = 0; int j = bufpos + 1;       /* corr_softhdb (:1692-1716) */
            for (int i = 0; i < HEADLEN; i++) {
                if (j >= HEADLEN) j = 0;
                float x = sbuf[j], y = 2.0 * (RS41_HDR[i] & 1) - 1.0;
                sum += y * sbuf[j]; normx += x * x; normy += y * y; j++;
            }
            sum /= sqrt(normx * normy);
            mv = sum;
            if (fabs(mv) > 0.7f) { found = 1; break; }                      /* hdb.ths = 0.7 (rs41mod.c:2864) */
        }
        if (!found) break;
        if (mv * (0.5 - inv) < 0) continue;                                 /* polarity rule, option.aut == 0 (rs41mod.c:2889-2892) */
        int byte_count = FRAMESTART, b8 = 0; char bitbuf[8]; float sb8[8];
        const long first = pos;
        while (byte_count < FRAME_LEN && pos < nbits) {                     /* rs41mod.c:2896-2956 */
            float s = soft[pos++]; if (softinv) s = -s;
            int bit = (s >= 0.0); float sb = s;
            if (inv) { bit ^= 1; sb = -sb; }
            sb8[b8] = sb; bitbuf[b8] = bit; b8++;
            if (b8 == 8) {
                float m = sb8[0];
                for (int q = 1; q < 8; q++) if (fabs(sb8[q]) < fabs(m)) m = sb8[q];
                int byte = 0; for (int q = 0; q < 8; q++) byte |= bitbuf[q] << q;   /* bits2byte: little endian (:224-234) */
                bytescore[byte_count] = m;
                frame[byte_count] = byte ^ RS41_MASK[byte_count % 64];
                byte_count++; b8 = 0;
            }
        }
        /* print_frame (rs41mod.c:2472-2491): short frames are zero-filled, otherwise bytes of the previous frame remain */
        if (byte_count < POS_GPS1) for (int i = byte_count; i < FRAME_LEN; i++) frame[i] = 0;
        int ft = 0; unsigned char b = frame[POS_FRAME - 1];
        for (int i = 0; i < 4; i++) ft += ((b >> i) & 1) - ((b >> (i + 4)) & 1);
        int ec = 0, crcm = 0;
        if (ecc > 0) {
            static const int bpos[6] = {0x039, 0x065, 0x093, 0x0B5, 0x112, 0x12B}, bpck[6] = {0x7928, 0x7A2A, 0x7C1E, 0x7D59, 0x7B15, 0x7600};
            ec = ecco_rs41_ecc(frame, ft >= 0 ? NDATA_LEN : FRAME_LEN, ecc);     /* print_frame: ec = rs41_ecc(gpx, len) (rs41mod.c:2525-2527) */
            for (int k = 0; k < 6; k++) if (ecco_rs41_check_crc(frame, bpos[k], bpck[k]) == 0) crcm |= 1 << k;
        }
        if (nf < maxframes) {
            sso_frame_t *f = &frames[nf];
            f->len = ft >= 0 ? NDATA_LEN : FRAME_LEN; f->first_bit = first; f->score = mv; f->nbytes_read = byte_count; f->ec = ec; f->crc_mask = crcm;
            memcpy(f->bytes, frame, FRAME_LEN); memcpy(f->bytescore, bytescore, sizeof bytescore);
        }
        nf++;
        if (pos >= nbits && byte_count < FRAME_LEN) break;                  /* EOF inside the frame ends the run */
    }
    return nf;
}
