/*
 * orc_cli.c — command-line driver for the CPU ORACLE (test infrastructure).
 * Mirrors `swiftover -t bed -c CHAIN -i IN -o OUT -u UNMATCHED` (main.d:41-52,
 * bed.d:72) for the BED path so the oracle can be diffed against the CUDA
 * product and timed as the CPU baseline.   Usage:
 *   orc_cli bed CHAIN IN OUT UNMATCHED [threads]
 */
#include "swo_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static char *slurp(const char *path, size_t *len) {
    FILE *f = strcmp(path, "-") ? fopen(path, "rb") : stdin;
    if (!f) return NULL;
    size_t cap = 1 << 20, n = 0;
    char *p = malloc(cap);
    for (;;) {
        size_t r = fread(p + n, 1, cap - n, f);
        n += r;
        if (r == 0) break;
        if (n == cap) p = realloc(p, cap *= 2);
    }
    if (f != stdin) fclose(f);
    *len = n;
    return p;
}
static double now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv) {
    if (argc < 6 || strcmp(argv[1], "bed")) {
        fprintf(stderr, "usage: %s bed CHAIN IN OUT UNMATCHED [threads]\n", argv[0]);
        return 1;
    }
    char err[256] = "";
    double t0 = now();
    orc_chainfile *cf = orc_chain_load(argv[2], err, sizeof err);
    if (!cf) {
        fprintf(stderr, "%s\n", err);
        return 1;
    }
    double t1 = now();
    size_t n;
    char *in = slurp(argv[3], &n);
    if (!in) {
        fprintf(stderr, "Input file does not exist\n");
        return 1;
    }
    int threads = argc > 6 ? atoi(argv[6]) : 1;
    orc_buf lifted = {0, 0, 0}, unm = {0, 0, 0};
    orc_bed_stats st;
    double t2 = now();
    int rc = orc_lift_bed_mt(cf, in, n, threads, &lifted, &unm, &st, err, sizeof err);
    double t3 = now();
    if (rc) {
        fprintf(stderr, "error %d: %s\n", rc, err);
        return 1;
    }
    FILE *fo = strcmp(argv[4], "-") ? fopen(argv[4], "wb") : stdout;
    FILE *fu = fopen(argv[5], "wb");
    if (!fo || !fu) {
        fprintf(stderr, "cannot open outputs\n");
        return 1;
    }
    fwrite(lifted.data, 1, lifted.len, fo);
    fwrite(unm.data, 1, unm.len, fu);
    if (fo != stdout) fclose(fo);
    fclose(fu);
    fprintf(stderr, "oracle: build %.1f ms, lift %.1f ms, %llu records, %llu rows, %llu unmatched, %d threads\n",
            (t1 - t0) * 1e3, (t3 - t2) * 1e3, (unsigned long long)st.n_records, (unsigned long long)st.n_rows,
            (unsigned long long)st.n_unmatched, threads);
    orc_buf_free(&lifted);
    orc_buf_free(&unm);
    orc_chain_free(cf);
    free(in);
    return 0;
}
