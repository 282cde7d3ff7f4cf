/* dft_detect.c -- drop-in replacement for the reference detector CLI scan/dft_detect.c in its IQ mode
 *     dft_detect [-v] [-c] [-s] [-t sec] --iq --bw <kHz> [--dc] [--ths x] - <sr> <8|16|32>   (stdin: raw IQ)
 * Same stdout lines ("<TYPE>: <score>[ , <+f>Hz]", "sample: N" with -v), same exit code (+-type number).
 * All DSP runs on the GPU through include/b200dft.h.  The stream is analysed incrementally (about half a second
 * at a time) so that, like the reference, the process answers as soon as a header is found when -c is not given.
 *     dft_detect ... --IQ <fq> --bw <kHz> [--dc] - <sr_base> <bits>   baseband capture: translate by -fq*sr_base, decimate to the IF
 *                                                                     rate (dft_detect.c:1383-1392), then the same detector
 * Not supported here (the auto_rx IQ scan never uses them): WAV / FM-audio input, running without --bw, --min, -d2, --ch2. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <errno.h>
#include "../../include/b200dft.h"

int main(int argc, char **argv)
{
    int verbose = 0, cont = 0, silent = 0, iq = 0, dc = 0, sr = 0, bits = 0, pcmraw = 0, wide = 0;
    double tl = -1.0, bw = 0.0, fq = 0.0, ths = 0.0; int have_ths = 0;
    for (char **a = argv + 1; *a; ++a) {
        if (!strcmp(*a, "-h") || !strcmp(*a, "--help")) {
            fprintf(stderr, "%s [options] - <sr> <bits>   (raw IQ on stdin)\n  options:\n       -v          (verbose)\n       -c          (continuous)\n"
                            "       --iq        (IF iq-data)\n       --bw <kHz>  (set IQ filter bw/kHz)\n       --dc\n       -t <sec>\n", argv[0]);
            return 0;
        } else if (!strcmp(*a, "-v") || !strcmp(*a, "--verbose")) verbose = 1;
        else if (!strcmp(*a, "--iq")) iq = 1;
        else if (!strcmp(*a, "--bw")) { if (!*++a) return -1; bw = atof(*a); if (bw < 1.0) bw = 0.0; }
        else if (!strcmp(*a, "--dc")) dc = 1;
        else if (!strcmp(*a, "-c") || !strcmp(*a, "--cnt")) cont = 1;
        else if (!strcmp(*a, "-s") || !strcmp(*a, "--silent")) silent = 1;
        else if (!strcmp(*a, "-t") || !strcmp(*a, "--time")) { if (!*++a) return -50; tl = atof(*a); }
        else if (!strcmp(*a, "-")) {
            if (!*++a) return -1; sr = atoi(*a);
            if (!*++a) return -1; bits = atoi(*a);
            if (sr < 1 || (bits != 8 && bits != 16 && bits != 32)) { fprintf(stderr, "- <sr> <bs>\n"); return -1; }
            pcmraw = 1;
        } else if (!strcmp(*a, "--IQ")) { if (!*++a) return -1; fq = atof(*a); iq = 1; wide = 1; }
        else if (!strcmp(*a, "--ths")) { if (!*++a) return -50; ths = atof(*a); have_ths = 1; }
        else if (!strcmp(*a, "--min") || !strcmp(*a, "-L") || !strcmp(*a, "-d2") || !strcmp(*a, "--ch2")) { fprintf(stderr, "dft_detect(b200): option %s not supported\n", *a); return -50; }
        else { fprintf(stderr, "dft_detect(b200): file/WAV input not supported, use: --iq --bw <kHz> - <sr> <bits>\n"); return -50; }
    }
    setbuf(stdout, NULL);
    if (!pcmraw || !iq || bw <= 0.0) { fprintf(stderr, "dft_detect(b200): needs --iq --bw <kHz> - <sr> <bits>\n"); return -50; }
    b200dft_t *h = NULL;
    const int crc = wide ? b200dft_create_wideband(&h, sr, bw, dc, 1, 0) : b200dft_create(&h, sr, bw, dc, 1, 0);
    if (crc != B200DFT_OK) { fprintf(stderr, "error: init buffers (%s)\n", b200dft_last_error()); return -50; }
    if (have_ths && b200dft_set_threshold(h, ths) != B200DFT_OK) { fprintf(stderr, "dft_detect(b200): %s\n", b200dft_last_error()); return -50; }
    if (wide) {                                         /* -t counts IF samples (:1494) = (tl+1)*sr input samples, sr = IF*dec */
        b200dft_wbinfo_t wb; b200dft_wideband_info(h, &wb);
        fprintf(stderr, "IF: %d\n", wb.if_sr); fprintf(stderr, "dec: %d\n", wb.decM);      /* :1067-1068 */
    }
    const int fmt = bits == 16 ? B200DFT_FMT_CS16 : (bits == 32 ? B200DFT_FMT_F32 : B200DFT_FMT_CU8), es = bits / 4;
    b200dft_result_t res; memset(&res, 0, sizeof res);
    if (!wide) {
        /* --iq: streaming handle -- every read is pushed once, the library keeps the K+L samples later blocks still need (O(n) work,
           constant memory, every detection printed: the reference's sample loop, dft_detect.c:1492-1652) */
        const long step = sr / 4 > 4096 ? sr / 4 : 4096;
        unsigned char *buf = malloc((size_t)step * es);
        int done = 0, eof = 0;
        if (b200dft_stream_begin(h, fmt, cont, tl) != B200DFT_OK) { fprintf(stderr, "dft_detect(b200): %s\n", b200dft_last_error()); return -50; }
        while (!done) {
            long n = 0;
            for (;;) {                                     /* one read: a live pipe must not wait for a full buffer */
                ssize_t r = read(0, buf, (size_t)step * es);
                if (r < 0 && errno == EINTR) continue;
                if (r <= 0) { eof = 1; break; }
                long bytes = r;
                while (bytes % es) { ssize_t r2 = read(0, buf + bytes, es - bytes % es); if (r2 <= 0) { eof = 1; bytes -= bytes % es; break; } bytes += r2; }
                n = bytes / es;
                break;
            }
            if (b200dft_stream_push_host(h, buf, n, n, eof, &res, &done) != B200DFT_OK) { fprintf(stderr, "dft_detect(b200): %s\n", b200dft_last_error()); return -50; }
            for (int i = 0; i < res.ndet && i < B200DFT_MAXDET; i++) {                /* dft_detect.c:1620-1634 */
                const b200dft_det_t *d = &res.dets[i];
                if (silent) continue;
                if (verbose) fprintf(stdout, "sample: %d\n", (int)d->pos);
                fprintf(stdout, "%s: %.4f", d->type, d->score);
                if ((!strncmp(d->type, "M10", 3) || !strncmp(d->type, "M20", 3)) && verbose) fprintf(stdout, " [%04X]", d->m10 & 0xFFFF);
                if (dc) {
                    fprintf(stdout, " , %+.1fHz", d->df_hz);
                    if (verbose) { fprintf(stdout, "   [ fq-ofs: %+.6f", d->df_hz / sr); fprintf(stdout, " = %+.1fHz ]", d->df_hz); }
                }
                fprintf(stdout, "\n");
            }
            if (res.ndet > B200DFT_MAXDET) fprintf(stderr, "dft_detect(b200): %d detections in one read exceed the %d kept\n", res.ndet, B200DFT_MAXDET);
            res.ndet = 0;
            if (eof && !done) break;
        }
        free(buf);
        b200dft_destroy(h);
        return res.exit_code;
    }
    long cap = (tl > 0 ? (long)((tl + 1) * sr) + 16 : (long)60 * sr), n = 0, step = sr / 2, next = step;
    unsigned char *buf = malloc((size_t)cap * es);
    int eof = 0, printed = 0;
    while (!eof) {
        while (n < next && n < cap) {
            ssize_t r = read(0, buf + n * es, (size_t)((next < cap ? next : cap) - n) * es);
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) { eof = 1; break; }
            long bytes = r;
            while (bytes % es) { ssize_t r2 = read(0, buf + n * es + bytes, es - bytes % es); if (r2 <= 0) { eof = 1; bytes -= bytes % es; break; } bytes += r2; }
            n += bytes / es;
        }
        if (n >= cap && tl <= 0) { cap *= 2; buf = realloc(buf, (size_t)cap * es); }
        const int rrc = wide ? b200dft_run_wideband_host(h, buf, fmt, n, &fq, cont, tl, &res) : b200dft_run_host(h, buf, fmt, n, n, cont, tl, &res);
        if (rrc != B200DFT_OK) { fprintf(stderr, "dft_detect(b200): %s\n", b200dft_last_error()); return -50; }
        for (int i = printed; i < res.ndet && i < B200DFT_MAXDET; i++) {            /* dft_detect.c:1620-1634 */
            const b200dft_det_t *d = &res.dets[i];
            if (silent) continue;
            if (verbose) fprintf(stdout, "sample: %d\n", (int)d->pos);
            fprintf(stdout, "%s: %.4f", d->type, d->score);
            if ((!strncmp(d->type, "M10", 3) || !strncmp(d->type, "M20", 3)) && verbose) fprintf(stdout, " [%04X]", d->m10 & 0xFFFF);
            if (dc) {
                fprintf(stdout, " , %+.1fHz", d->df_hz);
                if (verbose) { fprintf(stdout, "   [ fq-ofs: %+.6f", d->df_hz / sr); fprintf(stdout, " = %+.1fHz ]", d->df_hz); }
            }
            fprintf(stdout, "\n");
        }
        printed = res.ndet;
        if (res.ndet > 0 && !cont) break;
        if (tl > 0 && n >= (long)((tl + 1) * sr)) break;
        next = n + step;
        if (next > cap && tl > 0) next = cap;
        if (tl > 0 && n >= cap) break;
    }
    b200dft_destroy(h);
    return res.exit_code;
}
