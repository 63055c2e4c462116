/* Latency of small commitments through the C ABI, without Python (what the Rust shim would see for the m = 2 demo shapes):
 *     gcc -O2 -I include tools/small_commit_latency.c -o /tmp/scl -L plonky2_demo_b200 -lpcs -Wl,-rpath,$PWD/plonky2_demo_b200
 * Prints one JSON line: median microseconds per call. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "pcs.h"

/* oracle/oracle.c (the CPU port of the reference; test infrastructure) */
int ref_commit_from_values(uint64_t* values, size_t w, unsigned lg_d, unsigned rate_bits, unsigned cap_height, const uint64_t* salts,
                           size_t salt_w, uint64_t* leaves, uint64_t* digests, uint64_t* cap);

static double now_us(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e6 + t.tv_nsec * 1e-3;
}
static int cmp(const void* a, const void* b) { return (*(const double*)a > *(const double*)b) - (*(const double*)a < *(const double*)b); }

static double commit_us(size_t w, unsigned lg_d, int from_values, int reps) {
    const size_t d = (size_t)1 << lg_d;
    uint64_t** polys = malloc(w * sizeof(*polys));
    uint64_t** outs = malloc(w * sizeof(*outs));
    for (size_t j = 0; j < w; j++) {
        polys[j] = malloc(d * 8);
        outs[j] = malloc(d * 8);
        for (size_t i = 0; i < d; i++) polys[j][i] = (j * 1315423911u + i * 2654435761u) % 0xFFFFFFFF00000001ULL;
    }
    uint64_t cap[16 * 4];
    double* ts = malloc(reps * sizeof(double));
    for (int r = 0; r < reps + 3; r++) {
        pcs_batch* b = NULL;
        double t0 = now_us();
        int rc = from_values ? pcs_commit_from_values((const uint64_t* const*)polys, w, lg_d, 3, 4, NULL, 0, 0, outs, cap, &b)
                             : pcs_commit_from_coeffs((const uint64_t* const*)polys, w, lg_d, 3, 4, NULL, 0, 0, cap, &b);
        double t1 = now_us();
        if (rc) { fprintf(stderr, "commit failed: %s\n", pcs_last_error()); exit(2); }
        pcs_batch_free(b);
        if (r >= 3) ts[r - 3] = t1 - t0;
    }
    qsort(ts, reps, sizeof(double), cmp);
    return ts[reps / 2];
}

static double cpu_us(size_t w, unsigned lg_d, int reps) {
    const size_t d = (size_t)1 << lg_d, n = d << 3;
    uint64_t* v = malloc(w * d * 8);
    uint64_t* leaves = malloc(n * w * 8);
    uint64_t* dig = malloc(2 * n * 32 + 64);
    uint64_t cap[16 * 4];
    double* ts = malloc(reps * sizeof(double));
    for (int r = 0; r < reps + 1; r++) {
        for (size_t i = 0; i < w * d; i++) v[i] = (i * 2654435761u + r) % 0xFFFFFFFF00000001ULL;
        double t0 = now_us();
        ref_commit_from_values(v, w, lg_d, 3, 4, NULL, 0, leaves, dig, cap);
        if (r >= 1) ts[r - 1] = now_us() - t0;
    }
    qsort(ts, reps, sizeof(double), cmp);
    double m = ts[reps / 2];
    free(v); free(leaves); free(dig); free(ts);
    return m;
}

int main(void) {
    if (pcs_init(0, NULL)) { fprintf(stderr, "%s\n", pcs_last_error()); return 2; }
    uint64_t st[12] = {0};
    double tp[50];
    for (int r = 0; r < 50; r++) { double t0 = now_us(); pcs_poseidon_permute(st, 1); tp[r] = now_us() - t0; }
    qsort(tp, 50, sizeof(double), cmp);
    printf("{\"permute_1_state_us\": %.1f", tp[25]);
    const size_t ws[4] = {84, 135, 20, 16};
    const int fv[4] = {1, 1, 1, 0};
    const char* names[4] = {"constants_sigmas", "wires", "zs_partial_products", "quotient_chunks"};
    for (unsigned lg_d = 3; lg_d <= 15; lg_d += 12) {
        double total = 0;
        for (int k = 0; k < 4; k++) {
            double us = commit_us(ws[k], lg_d, fv[k], 21);
            total += us;
            printf(", \"%s_2^%u_us\": %.1f", names[k], lg_d, us);
        }
        printf(", \"total_2^%u_us\": %.1f", lg_d, total);
    }
    /* where does the device path start to win?  wires-shaped from_values commits (135 polynomials, coefficients returned) */
    printf(", \"crossover_135_polys\": [");
    for (unsigned lg_d = 3; lg_d <= 13; lg_d++)
        printf("%s{\"lg_d\": %u, \"leaves\": %u, \"gpu_us\": %.1f, \"cpu_port_us\": %.1f}", lg_d > 3 ? ", " : "", lg_d, 8u << lg_d,
               commit_us(135, lg_d, 1, 11), cpu_us(135, lg_d, lg_d > 10 ? 3 : 7));
    printf("]}\n");
    pcs_shutdown();
    return 0;
}
