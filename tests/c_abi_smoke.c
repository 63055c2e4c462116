/* C-ABI smoke test: what a host language without Python sees of libpcs.so.
 *
 *     gcc -O2 -I include tests/c_abi_smoke.c -o /tmp/c_abi_smoke -L plonky2_demo_b200 -lpcs -L oracle -loracle \
 *         -Wl,-rpath,$PWD/plonky2_demo_b200 -Wl,-rpath,$PWD/oracle
 *     /tmp/c_abi_smoke [n_devices]
 *
 * It does what the Rust shim of INTEGRATION.md does for PolynomialBatch::from_coeffs / from_values
 * (plonky2/src/fri/oracle.rs:43-98) and for the query phase's tree.get / tree.prove (fri/prover.rs:183-216):
 * commit from separately allocated host vectors (Vec<Vec<F>>), read the cap, rows and Merkle paths back, and -- with
 * n_devices > 1 -- the same commitment through pcs_multi_* over several GPUs from this one process.
 * The checker is the CPU oracle (liboracle.so, test infrastructure): cap, digests, rows and paths must be bit-identical.
 * Exit code 0 = all equal; prints one JSON line.                                                                      */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "pcs.h"

/* oracle/oracle.c (test infrastructure) */
int ref_commit_from_coeffs(const uint64_t* coeffs, size_t w, unsigned lg_d, unsigned rate_bits, unsigned cap_height,
                           const uint64_t* salts, size_t salt_w, uint64_t* leaves, uint64_t* digests, uint64_t* cap);
int ref_merkle_prove(const uint64_t* digests, size_t n, unsigned cap_height, size_t leaf_index, uint64_t* siblings);
int ref_fft_batch(uint64_t* polys, size_t w, unsigned lg_n, int inverse, unsigned zero_factor);

#define P 0xFFFFFFFF00000001ULL

static uint64_t sm_state;
static uint64_t splitmix64(void) {
    uint64_t z = (sm_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

#define CHECK(call)                                                                   \
    do {                                                                              \
        int _rc = (call);                                                             \
        if (_rc != 0) {                                                               \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, _rc, pcs_last_error());    \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

static double now_ms(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

int main(int argc, char** argv) {
    const int n_devices = argc > 1 ? atoi(argv[1]) : 1;
    const size_t w = 135;
    const unsigned lg_d = argc > 2 ? (unsigned)atoi(argv[2]) : 12, rate_bits = 3, cap_height = 4;
    const size_t d = (size_t)1 << lg_d, n = d << rate_bits, n_cap = (size_t)1 << cap_height, n_dig = 2 * (n - n_cap);
    const unsigned n_sib = lg_d + rate_bits - cap_height;

    /* Vec<PolynomialCoeffs<F>>: w separately allocated vectors (pageable memory, like a Rust Vec) */
    uint64_t** polys = malloc(w * sizeof(*polys));
    uint64_t* flat = malloc(w * d * 8);
    for (size_t j = 0; j < w; j++) {
        polys[j] = malloc(d * 8);
        sm_state = 0x5EED0000ULL + j;
        for (size_t i = 0; i < d; i++) {
            uint64_t v;
            do v = splitmix64(); while (v >= P);
            polys[j][i] = flat[j * d + i] = v;
        }
    }
    polys[0][0] = P + 2;   /* non-canonical inputs are legal (goldilocks_field.rs:33-37) */
    flat[0] = 2;

    /* the checker */
    uint64_t* ref_leaves = malloc(n * w * 8);
    uint64_t* ref_dig = malloc(n_dig * 32 + 32);
    uint64_t ref_cap[16 * 4];
    double t0 = now_ms();
    if (ref_commit_from_coeffs(flat, w, lg_d, rate_bits, cap_height, NULL, 0, ref_leaves, ref_dig, ref_cap)) return 3;
    double cpu_ms = now_ms() - t0;

    /* ---- single GPU: PolynomialBatch::from_coeffs ---- */
    CHECK(pcs_init(0, NULL));
    uint64_t cap[16 * 4];
    pcs_batch* b = NULL;
    CHECK(pcs_commit_from_coeffs((const uint64_t* const*)polys, w, lg_d, rate_bits, cap_height, NULL, 0, PCS_KEEP_COEFFS, cap, &b));
    t0 = now_ms();
    pcs_batch* b2 = NULL;
    CHECK(pcs_commit_from_coeffs((const uint64_t* const*)polys, w, lg_d, rate_bits, cap_height, NULL, 0, 0, cap, &b2));
    double gpu_ms = now_ms() - t0;
    pcs_batch_free(b2);
    int ok_cap = memcmp(cap, ref_cap, sizeof(ref_cap)) == 0;

    size_t nl, ll, nd;
    unsigned ch;
    CHECK(pcs_batch_shape(b, &nl, &ll, &nd, &ch));
    int ok_shape = nl == n && ll == w && nd == n_dig && ch == cap_height;

    uint64_t* dig = malloc(n_dig * 32 + 32);
    CHECK(pcs_batch_digests(b, dig));
    int ok_dig = memcmp(dig, ref_dig, n_dig * 32) == 0;

    /* query phase: 28 rows + paths (fri/prover.rs:183-216) */
    enum { Q = 28 };
    uint64_t idx[Q];
    sm_state = 99;
    for (int k = 0; k < Q; k++) idx[k] = splitmix64() % n;
    idx[0] = 0;
    idx[1] = n - 1;
    uint64_t* rows = malloc(Q * w * 8);
    uint64_t* sib = malloc((size_t)Q * n_sib * 32);
    uint64_t* ref_sib = malloc((size_t)n_sib * 32);
    CHECK(pcs_batch_get_rows(b, idx, Q, rows));
    CHECK(pcs_batch_prove_many(b, idx, Q, sib));
    int ok_rows = 1, ok_paths = 1;
    for (int k = 0; k < Q; k++) {
        ok_rows &= memcmp(rows + (size_t)k * w, ref_leaves + idx[k] * w, w * 8) == 0;
        if (ref_merkle_prove(ref_dig, n, cap_height, idx[k], ref_sib) != (int)n_sib) return 3;
        ok_paths &= memcmp(sib + (size_t)k * n_sib * 4, ref_sib, (size_t)n_sib * 32) == 0;
    }
    /* polynomials[i] are kept (OpeningSet::new needs them, plonk/proof.rs:316-321) */
    uint64_t* c3 = malloc(d * 8);
    CHECK(pcs_batch_coeffs(b, 3, c3));
    int ok_coeffs = memcmp(c3, flat + 3 * d, d * 8) == 0;
    pcs_batch_free(b);

    /* ---- PolynomialBatch::from_values: the same commitment from point values ---- */
    uint64_t* vals = malloc(w * d * 8);
    memcpy(vals, flat, w * d * 8);
    if (ref_fft_batch(vals, w, lg_d, 0, 0)) return 3;
    uint64_t** vptr = malloc(w * sizeof(*vptr));
    uint64_t** cptr = malloc(w * sizeof(*cptr));
    uint64_t* cback = malloc(w * d * 8);
    for (size_t j = 0; j < w; j++) {
        vptr[j] = vals + j * d;
        cptr[j] = cback + j * d;
    }
    uint64_t capv[16 * 4];
    pcs_batch* bv = NULL;
    CHECK(pcs_commit_from_values((const uint64_t* const*)vptr, w, lg_d, rate_bits, cap_height, NULL, 0, 0, cptr, capv, &bv));
    int ok_values = memcmp(capv, ref_cap, sizeof(ref_cap)) == 0 && memcmp(cback, flat, w * d * 8) == 0;
    pcs_batch_free(bv);

    /* ---- several GPUs, one process ---- */
    int ok_multi = 1, multi_devs = 0;
    double multi_ms = 0;
    if (n_devices >= 1) {
        CHECK(pcs_multi_init(NULL, n_devices));
        multi_devs = pcs_multi_devices(NULL);
        uint64_t capm[16 * 4];
        pcs_multi_batch* mb = NULL;
        CHECK(pcs_multi_commit_from_coeffs((const uint64_t* const*)polys, w, lg_d, rate_bits, cap_height, NULL, 0, 0, capm, &mb));
        pcs_multi_batch_free(mb);
        t0 = now_ms();
        CHECK(pcs_multi_commit_from_coeffs((const uint64_t* const*)polys, w, lg_d, rate_bits, cap_height, NULL, 0, 0, capm, &mb));
        multi_ms = now_ms() - t0;
        ok_multi &= memcmp(capm, ref_cap, sizeof(ref_cap)) == 0;
        CHECK(pcs_multi_batch_get_rows(mb, idx, Q, rows));
        for (int k = 0; k < Q; k++) {
            ok_multi &= memcmp(rows + (size_t)k * w, ref_leaves + idx[k] * w, w * 8) == 0;
            CHECK(pcs_multi_batch_prove(mb, idx[k], sib));
            if (ref_merkle_prove(ref_dig, n, cap_height, idx[k], ref_sib) != (int)n_sib) return 3;
            ok_multi &= memcmp(sib, ref_sib, (size_t)n_sib * 32) == 0;
        }
        pcs_multi_batch_free(mb);
    }
    pcs_shutdown();

    const int ok = ok_cap && ok_shape && ok_dig && ok_rows && ok_paths && ok_coeffs && ok_values && ok_multi;
    printf("{\"c_abi_smoke\": %s, \"w\": %zu, \"lg_d\": %u, \"cap\": %d, \"shape\": %d, \"digests\": %d, \"rows\": %d, \"paths\": %d, "
           "\"coeffs\": %d, \"from_values\": %d, \"multi\": %d, \"multi_devices\": %d, \"gpu_commit_ms\": %.3f, "
           "\"multi_commit_ms\": %.3f, \"cpu_oracle_ms\": %.1f}\n",
           ok ? "true" : "false", w, lg_d, ok_cap, ok_shape, ok_dig, ok_rows, ok_paths, ok_coeffs, ok_values, ok_multi, multi_devs,
           gpu_ms, multi_ms, cpu_ms);
    return ok ? 0 : 1;
}
