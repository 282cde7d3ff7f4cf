/* fsk_demod.c -- drop-in replacement for the reference CLI utils/fsk_demod.c: same argv, same stdin sample formats,
 * same stdout bytes (uint8 bits or float32 soft decisions, flushed as frames complete), same stderr JSON statistics,
 * exit codes and SIGTERM behaviour -- but every sample is demodulated on the GPU through the C-ABI of
 * include/b200fsk.h (one channel; the batched entry points are what a scanner uses for hundreds).
 * Host code only: argument parsing, the read loop, test-frame BER bookkeeping and text formatting. */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <getopt.h>
#include <signal.h>
#include <unistd.h>
#include <fcntl.h>
#include <errno.h>
#include "../../include/b200fsk.h"

#define TEST_FRAME_SIZE 100          /* utils/fsk_demod.c:30 */
#define DEFAULT_P 10                 /* utils/fsk_demod.c:71 (the help text says 8, the code uses 10) */
#define DEFAULT_NSYM 50              /* utils/fsk.h:46 */

static void on_term(int s) { if (s == SIGTERM) _exit(0); }     /* utils/fsk_demod.c:46-51 */

static void usage(const char *a0)
{
    fprintf(stderr, "usage: %s [options] (2|4) SampleRate SymbolRate InputModemRawFile OutputFile\n", a0);
    fprintf(stderr, " -c --cs16          The raw input file will be in complex signed 16 bit format.\n");
    fprintf(stderr, " -d --cu8           The raw input file will be in complex unsigned 8 bit format.\n");
    fprintf(stderr, "                    If neither -c nor -d are used, the input should be in signed 16 bit format.\n");
    fprintf(stderr, " -f --testframes    Testframe mode, prints stats to stderr when a testframe is detected, if -t (JSON) \n");
    fprintf(stderr, "                    is enabled stats will be in JSON format\n");
    fprintf(stderr, " -t[r] --stats=[r]  Print out modem statistics to stderr in JSON.\n");
    fprintf(stderr, "                    r, if provided, sets the number of modem frames between statistic printouts.\n");
    fprintf(stderr, " -s --soft-dec      The output file will be in a soft-decision format, with one 32-bit float per bit.\n");
    fprintf(stderr, "                    If -s is not used, the output will be in a 1 byte-per-bit format.\n");
    fprintf(stderr, " -i --softinv       Invert soft decision polarity.\n");
    fprintf(stderr, " -p P               The demod internals operate at a rate of Fs/P, default %d\n", DEFAULT_P);
    fprintf(stderr, " --fsk_lower freq   lower limit of freq estimator (default 0 for real input, -Fs/2  for complex input)\n");
    fprintf(stderr, " --fsk_upper freq   upper limit of freq estimator (default Fs/2)\n");
    fprintf(stderr, " --nsym Nsym        number of symbols used for estimators. Default %d\n", DEFAULT_NSYM);
    fprintf(stderr, " --mask TxFreqSpace Use \"mask\" freq estimator (default is \"peak\" estimator)\n");
    exit(1);
}

int main(int argc, char **argv)
{
    int P = DEFAULT_P, nsym = DEFAULT_NSYM, M, Fs, Rs, soft = 0, softinv = 0, testframes = 0, stats = 0, stats_rate = 8;
    int complex_in = 1, bytes_per = 2, lower = 0, upper = 0, user_lower = 0, user_upper = 0, mask = 0, tone_sep = 100;
    static struct option lo[] = {{"help", 0, 0, 'h'}, {"softinv", 0, 0, 'i'}, {"conv", 1, 0, 'p'}, {"cs16", 0, 0, 'c'}, {"cu8", 0, 0, 'd'},
        {"fsk_lower", 1, 0, 'b'}, {"fsk_upper", 1, 0, 'u'}, {"stats", 2, 0, 't'}, {"soft-dec", 0, 0, 's'}, {"testframes", 0, 0, 'f'},
        {"nsym", 1, 0, 'n'}, {"mask", 1, 0, 'm'}, {0, 0, 0, 0}};
    int o, idx = 0;
    while ((o = getopt_long(argc, argv, "fhilp:cdt::sb:u:m", lo, &idx)) != -1) {
        switch (o) {
        case 'c': complex_in = 2; bytes_per = 2; break;
        case 'd': complex_in = 2; bytes_per = 1; break;
        case 'f': testframes = 1; break;
        case 'i': softinv = 1; break;
        case 't': stats = 1; if (optarg) { stats_rate = atoi(optarg); if (!stats_rate) stats_rate = 8; } break;
        case 's': soft = 1; break;
        case 'p': P = atoi(optarg); break;
        case 'b': if (optarg) { lower = atoi(optarg); user_lower = 1; } break;
        case 'u': if (optarg) { upper = atoi(optarg); user_upper = 1; } break;
        case 'n': if (optarg) nsym = atoi(optarg); break;
        case 'm': mask = 1; if (optarg) tone_sep = atoi(optarg); break;      /* the reference dereferences a NULL optarg for short -m */
        case 'l': break;
        default: usage(argv[0]);
        }
    }
    int dx = optind;
    if (argc - dx < 5) { fprintf(stderr, "Too few arguments\n"); usage(argv[0]); }
    if (argc - dx > 5) { fprintf(stderr, "Too many arguments\n"); usage(argv[0]); }
    M = atoi(argv[dx]); Fs = atoi(argv[dx + 1]); Rs = atoi(argv[dx + 2]);
    if (M != 2 && M != 4) { fprintf(stderr, "Mode %d is not valid. Mode must be 2 or 4.\n", M); usage(argv[0]); }
    int fdin = strcmp(argv[dx + 3], "-") ? open(argv[dx + 3], O_RDONLY) : 0;
    FILE *fout = strcmp(argv[dx + 4], "-") ? fopen(argv[dx + 4], "w") : stdout;
    if (!user_lower) lower = (complex_in == 1) ? 0 : -Fs / 2;
    if (!user_upper) upper = Fs / 2;
    fprintf(stderr, "Setting estimator limits to %d to %d Hz.\n", lower, upper);
    if (fdin < 0 || !fout) { fprintf(stderr, "Couldn't open files\n"); exit(1); }

    b200fsk_cfg_t cfg = {Fs, Rs, M, P, nsym, 1000, tone_sep, lower, upper, mask};
    b200fsk_t *h = NULL;
    if (b200fsk_create(&h, &cfg, 1, 0) != B200FSK_OK) {            /* the reference asserts on bad configurations */
        fprintf(stderr, "fsk_demod: %s\n", b200fsk_last_error());
        abort();
    }
    b200fsk_info_t info; b200fsk_info(h, &info);
    if (stats) b200fsk_enable_stats(h, 1);
    const int fmt = complex_in == 1 ? B200FSK_FMT_S16 : (bytes_per == 1 ? B200FSK_FMT_CU8 : B200FSK_FMT_CS16);
    const int es = (fmt == B200FSK_FMT_CS16) ? 4 : 2;

    uint8_t tx[TEST_FRAME_SIZE], rx[TEST_FRAME_SIZE]; int tfcnt = 0, bitcnt = 0, biterr = 0;
    if (testframes) { srand(158324); for (int i = 0; i < TEST_FRAME_SIZE; i++) { tx[i] = rand() & 1; rx[i] = 0; } }   /* :247-251 */
    int stats_loop = 0, stats_ctr = 0;
    if (stats) { float loop_time = ((float)info.N) / ((float)Fs); stats_loop = (int)(1 / (stats_rate * loop_time)); }
    if (signal(SIGTERM, on_term) == SIG_ERR) printf("\ncan't catch SIGTERM\n");

    const long CHUNK = 1 << 18;                                    /* samples per read when data is plentiful */
    unsigned char *inbuf = malloc((size_t)CHUNK * es);
    const long maxf_cap = b200fsk_push_frames_bound(h, CHUNK + info.nin_max, 0) + 4;
    void *outbuf = malloc((size_t)maxf_cap * info.Nbits * (soft ? 4 : 1));
    long sample_count = 0; int eof = 0, nin = info.N;
    long have = 0;                                                 /* samples pushed but not yet consumed (carried inside the handle) */
    while (!eof || have >= nin) {
        /* block until one more frame's worth is available, then take whatever else is already there */
        long got = 0;
        /* frames allowed in this call: stop on the frame that prints statistics so that they describe that frame */
        int maxf = 0;
        if (testframes) maxf = 1; else if (stats) maxf = stats_ctr < 0 ? 1 : stats_ctr + 2;
        /* Read only what this call can consume: with a frame cap, more than maxf frames' worth would pile up inside the
           handle (O(file) memory, quadratic tail copies on file input); and never block on read() while a whole frame is already there. */
        long want = CHUNK, need = nin > have ? nin - have : 0;
        if (maxf > 0) { want = (long)maxf * info.nin_max - have; if (want > CHUNK) want = CHUNK; if (want < need) want = need; }
        if (!eof && want > 0 && need > 0) {
            while (got < want) {
                ssize_t r = read(fdin, inbuf + got * es, (size_t)(want - got) * es);
                if (r < 0 && errno == EINTR) continue;
                if (r <= 0) { eof = 1; break; }
                /* keep partial samples aligned */
                long bytes = r;
                while (bytes % es) { ssize_t r2 = read(fdin, inbuf + got * es + bytes, es - bytes % es); if (r2 <= 0) { eof = 1; bytes -= bytes % es; break; } bytes += r2; }
                got += bytes / es;
                if (got >= need) break;
            }
        }
        int frames = 0;
        if (b200fsk_push_host(h, inbuf, fmt, got, got, maxf, outbuf, maxf_cap * info.Nbits, soft, softinv, &frames) != B200FSK_OK) {
            fprintf(stderr, "fsk_demod: %s\n", b200fsk_last_error()); exit(1);
        }
        b200fsk_stats_t st; b200fsk_get_stats(h, 0, &st);
        have = have + got - (st.samples - sample_count);
        sample_count = st.samples; nin = st.nin;
        for (int f = 0; f < frames; f++) {
            int tf_detected = 0;
            if (testframes) {                                       /* :319-357 */
                for (int j = 0; j < info.Nbits; j++) {
                    memmove(rx, rx + 1, TEST_FRAME_SIZE - 1);
                    if (soft) { float v = ((float *)outbuf)[f * info.Nbits + j]; if (softinv) v = -v; rx[TEST_FRAME_SIZE - 1] = v < 0.0; }
                    else rx[TEST_FRAME_SIZE - 1] = ((uint8_t *)outbuf)[f * info.Nbits + j];
                    int errs = 0;
                    for (int i = 0; i < TEST_FRAME_SIZE; i++) errs += rx[i] != tx[i];
                    if (errs < 0.1 * TEST_FRAME_SIZE) {
                        tf_detected = 1; tfcnt++; bitcnt += TEST_FRAME_SIZE; biterr += errs;
                        if (!stats) fprintf(stderr, "errs: %d FSK BER %f, bits tested %d, bit errors %d\n", errs, ((float)biterr / (float)bitcnt), bitcnt, biterr);
                    }
                }
            }
            if (stats) {
                if (stats_ctr < 0 || tf_detected) {                 /* :360-413; this is always the last frame of the call */
                    fprintf(stderr, "{");
                    fprintf(stderr, "\"samples\": %ld, \"EbNodB\": %5.1f, \"ppm\": %4d,", sample_count, st.snr_est, (int)st.ppm);
                    fprintf(stderr, " \"f1_est\":%.1f, \"f2_est\":%.1f", st.f_est[0], st.f_est[1]);
                    if (M == 4) fprintf(stderr, ", \"f3_est\":%.1f, \"f4_est\":%.1f", st.f_est[2], st.f_est[3]);
                    if (!testframes) {
                        fprintf(stderr, ",\t\"eye_diagram\":[");
                        for (int i = 0; i < st.neyetr; i++) {
                            fprintf(stderr, "[");
                            for (int j = 0; j < st.neyesamp; j++) { fprintf(stderr, "%f ", st.rx_eye[i][j]); if (j < st.neyesamp - 1) fprintf(stderr, ","); }
                            fprintf(stderr, "]");
                            if (i < st.neyetr - 1) fprintf(stderr, ",");
                        }
                        fprintf(stderr, "],");
                        fprintf(stderr, "\"samp_fft\":[");
                        float *Sf = malloc(sizeof(float) * info.Ndft);
                        b200fsk_get_Sf(h, 0, Sf);
                        for (int i = 0; i < info.Ndft / 2; i++) { fprintf(stderr, "%f ", Sf[i]); if (i < info.Ndft / 2 - 1) fprintf(stderr, ","); }
                        free(Sf);
                        fprintf(stderr, "]");
                    }
                    if (testframes) fprintf(stderr, ", \"frames\":%d, \"bits\":%d, \"errs\":%d", tfcnt, bitcnt, biterr);
                    fprintf(stderr, "}\n");
                    if (stats_ctr < 0) stats_ctr = stats_loop;
                }
                if (!testframes) stats_ctr--;
            }
        }
        if (frames) {
            fwrite(outbuf, soft ? 4 : 1, (size_t)frames * info.Nbits, fout);
            if (fout == stdout) fflush(stdout);
        }
        if (eof && got == 0 && frames == 0) break;
    }
    fclose(fout);
    b200fsk_destroy(h);
    return 0;
}
