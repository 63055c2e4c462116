/*
 * ORACLE -- test infrastructure, NOT the product.
 *
 * CPU restatement (plain C + OpenMP) of the reference's PolynomialBatch commit path:
 *   batched Goldilocks (I)FFT / coset LDE  ->  transpose + bit-reverse  ->  Poseidon Merkle tree.
 * The reference is Rust (nightly-2023-06-30); no Rust toolchain exists in the build image, so
 * the reference itself cannot be compiled here ("port", not "reference").  Parity is PINNED by
 * every golden number the reference's own tests hold for this path (tests/golden/reference_kats.json:
 * 4 Poseidon-12 KATs, the 256-entry bit-reversal table, the field constants) plus the reference's
 * property tests restated in tests/ (FFT == naive evaluation, zero_factor equivalence, coset FFT ==
 * naive evaluation, every Merkle proof verifies against the cap).  The reference holds NO golden
 * LDE vectors, Merkle roots or proofs, so caps/digests are pinned by construction only.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference) may load
 * this library; the product (libpcs.so) never does.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "gl64.h"
#include "poseidon_constants.h"

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * util/src/lib.rs:35 log2_strict ; plonky2/src/util/mod.rs:30 reverse_bits
 * ---------------------------------------------------------------------------------------- */
static int log2_strict(size_t n) {
    int r = __builtin_ctzll(n);
    return ((n >> r) == 1) ? r : -1;
}

static inline uint64_t bitreverse64(uint64_t x) {
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x);
}

static inline size_t reverse_bits(size_t x, unsigned bits) {
    if (bits == 0) return 0;
    return (size_t)(bitreverse64((uint64_t)x) >> (64 - bits));
}

/* util/src/lib.rs:62,188: out[brev(i)] = in[i] (semantics only; the reference's cache-oblivious
 * variant computes the same permutation). */
API void ref_reverse_index_bits(uint64_t* a, size_t n) {
    int lg = log2_strict(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = reverse_bits(i, (unsigned)lg);
        if (i < j) {
            uint64_t t = a[i];
            a[i] = a[j];
            a[j] = t;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * field/src/fft.rs:14-33 fft_root_table: row k (k = lg_m - 1) holds the first max(2^k, 2) powers
 * of the primitive 2^(k+1)-th root.  Flat layout: rows concatenated, row_off[k] gives the start.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int lg_n;
    gl_t* data;
    size_t row_off[GL_TWO_ADICITY + 1];
} root_table_t;

static void root_table_build(root_table_t* t, int lg_n) {
    t->lg_n = lg_n;
    size_t total = 0;
    for (int lg_m = 1; lg_m <= lg_n; lg_m++) {
        size_t half = (size_t)1 << (lg_m - 1);
        t->row_off[lg_m - 1] = total;
        total += half < 2 ? 2 : half;
    }
    t->data = (gl_t*)malloc((total ? total : 1) * sizeof(gl_t));
    gl_t bases[GL_TWO_ADICITY + 1];
    if (lg_n > 0) {
        gl_t base = gl_primitive_root_of_unity((unsigned)lg_n);
        bases[0] = base;
        for (int i = 1; i < lg_n; i++) {
            base = gl_sqr(base);
            bases[i] = base;
        }
    }
    for (int lg_m = 1; lg_m <= lg_n; lg_m++) {
        size_t half = (size_t)1 << (lg_m - 1);
        size_t len = half < 2 ? 2 : half;
        gl_t base = bases[lg_n - lg_m];
        gl_t* row = t->data + t->row_off[lg_m - 1];
        gl_t cur = 1; /* types.rs:429 powers() */
        for (size_t j = 0; j < len; j++) {
            row[j] = cur;
            cur = gl_mul(cur, base);
        }
    }
}

static void root_table_free(root_table_t* t) { free(t->data); }

/* ------------------------------------------------------------------------------------------
 * field/src/fft.rs:169-206 fft_classic (+ the scalar body of fft_classic_simd :142-160):
 * bit-reverse the input, collapse the first r stages into "replicate", then radix-2 DIT stages.
 * ---------------------------------------------------------------------------------------- */
static void fft_classic(gl_t* v, int r, int lg_n, const root_table_t* t) {
    size_t n = (size_t)1 << lg_n;
    ref_reverse_index_bits(v, n);
    if (r > 0) {
        size_t mask = ~(((size_t)1 << r) - 1);
        for (size_t i = 0; i < n; i++) v[i] = v[i & mask];
    }
    for (int lg_half_m = r; lg_half_m < lg_n; lg_half_m++) {
        size_t half_m = (size_t)1 << lg_half_m, m = half_m << 1;
        const gl_t* omega = t->data + t->row_off[lg_half_m];
        for (size_t k = 0; k < n; k += m) {
            for (size_t j = 0; j < half_m; j++) {
                gl_t tt = gl_mul(omega[j], v[k + half_m + j]);
                gl_t u = v[k + j];
                v[k + j] = gl_add(u, tt);
                v[k + half_m + j] = gl_sub(u, tt);
            }
        }
    }
}

/* field/src/fft.rs:72-95 ifft_with_options */
static void ifft_inplace(gl_t* v, int lg_n, const root_table_t* t) {
    size_t n = (size_t)1 << lg_n;
    gl_t n_inv = gl_inverse_2exp((unsigned)lg_n);
    fft_classic(v, 0, lg_n, t);
    v[0] = gl_mul(v[0], n_inv);
    if (n > 1) v[n / 2] = gl_mul(v[n / 2], n_inv);
    for (size_t i = 1; i < n / 2; i++) {
        size_t j = n - i;
        gl_t ci = gl_mul(v[j], n_inv), cj = gl_mul(v[i], n_inv);
        v[i] = ci;
        v[j] = cj;
    }
}

/* fft_with_options / ifft_with_options on a batch [w][n] (natural order in and out). */
API int ref_fft_batch(uint64_t* polys, size_t w, unsigned lg_n, int inverse, unsigned zero_factor) {
    root_table_t t;
    root_table_build(&t, (int)lg_n);
    size_t n = (size_t)1 << lg_n;
#pragma omp parallel for schedule(dynamic)
    for (size_t j = 0; j < w; j++) {
        if (inverse)
            ifft_inplace(polys + j * n, (int)lg_n, &t);
        else
            fft_classic(polys + j * n, (int)zero_factor, (int)lg_n, &t);
    }
    root_table_free(&t);
    for (size_t i = 0; i < w * n; i++) polys[i] = gl_canon(polys[i]);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * plonky2/src/fri/oracle.rs:100-125 lde_values (without salts):
 *   p.lde(rate_bits)                      polynomial/mod.rs:201-203 (zero pad to N)
 *    .coset_fft_with_options(shift, Some(rate_bits), table)   polynomial/mod.rs:282-295
 * out is poly-major [w][N], natural order: out[j][i] = P_j(shift * w_N^i).
 * ---------------------------------------------------------------------------------------- */
static void lde_one(const gl_t* coeffs, size_t d, int lg_n, unsigned rate_bits, gl_t shift,
                    const root_table_t* t, gl_t* out) {
    size_t n = (size_t)1 << lg_n;
    gl_t pw = 1;
    for (size_t i = 0; i < d; i++) { /* shift.powers().zip(coeffs) -- zeros beyond d stay zero */
        out[i] = gl_mul(pw, coeffs[i]);
        pw = gl_mul(pw, shift);
    }
    memset(out + d, 0, (n - d) * sizeof(gl_t));
    fft_classic(out, (int)rate_bits, lg_n, t);
}

API int ref_coset_lde_batch(const uint64_t* coeffs /*[w][d]*/, size_t w, unsigned lg_d,
                            unsigned rate_bits, uint64_t shift, uint64_t* out /*[w][N]*/) {
    int lg_n = (int)(lg_d + rate_bits);
    if (lg_n > GL_TWO_ADICITY) return -1;
    size_t d = (size_t)1 << lg_d, n = (size_t)1 << lg_n;
    root_table_t t;
    root_table_build(&t, lg_n);
#pragma omp parallel for schedule(dynamic)
    for (size_t j = 0; j < w; j++) lde_one(coeffs + j * d, d, lg_n, rate_bits, shift, &t, out + j * n);
    root_table_free(&t);
#pragma omp parallel for
    for (size_t i = 0; i < w * n; i++) out[i] = gl_canon(out[i]);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Poseidon-12.  plonky2/src/hash/poseidon.rs.
 * ---------------------------------------------------------------------------------------- */
#define PW 12
#define N_PARTIAL 22
#define HALF_FULL 4

/* poseidon.rs:522-528 sbox_monomial x -> x^7 */
static inline gl_t sbox(gl_t x) {
    gl_t x2 = gl_sqr(x), x4 = gl_sqr(x2), x3 = gl_mul(x, x2);
    return gl_mul(x3, x4);
}

/* poseidon.rs:484-495 constant_layer */
static inline void constant_layer(gl_t* s, int round) {
    for (int i = 0; i < PW; i++) s[i] = gl_add(s[i], ALL_ROUND_CONSTANTS[i + PW * round]);
}

/* One 32-bit half of the MDS product: y = circ-correlation(s) (exact in int64, inputs < 2^32, results
 * < 2^43).  The 12-point circulant is evaluated through x^12 - 1 = prod over zeta in {1,-1,i,-i} of
 * (x^3 - zeta) with x^4 = zeta: three 4-point integer DFTs, three 3x3 twisted products whose constants
 * are [16,32,16], [-1,-8,2] and (2+i),(-4-i),(16-i), three inverse DFTs -- the decomposition the
 * reference's x86-64 path uses (poseidon_goldilocks.rs:307-441 mds_multiply_freq), written here as the
 * add/shift network tests/golden/make_golden.py:mds_network_check derives and checks against the plain
 * circulant sums of poseidon.rs:178-198 mds_row_shf. */
static inline void mds_half(const int64_t* s, int64_t* y) {
    int64_t A[3], B[3], P[3], Q[3];
    for (int j = 0; j < 3; j++) {
        int64_t u = s[j] + s[j + 6], v = s[j + 3] + s[j + 9];
        A[j] = u + v;
        B[j] = u - v;
        P[j] = s[j] - s[j + 6];
        Q[j] = s[j + 3] - s[j + 9];
    }
    int64_t t = A[0] + A[1] + A[2];
    int64_t Ya[3] = {(t + A[2]) << 4, (t + A[0]) << 4, (t + A[1]) << 4};
    int64_t Yb[3] = {(B[2] << 3) - (B[1] << 1) - B[0], -(B[0] << 3) - B[1] - (B[2] << 1),
                     (B[0] << 1) - (B[1] << 3) - B[2]};
    int64_t re[3], im[3];
    re[0] = (P[0] << 1) - Q[0] + P[1] - (Q[1] << 4) + P[2] + (Q[2] << 2);
    im[0] = P[0] + (Q[0] << 1) + (P[1] << 4) + Q[1] - (P[2] << 2) + Q[2];
    re[1] = -(P[0] << 2) + Q[0] + (P[1] << 1) - Q[1] + P[2] - (Q[2] << 4);
    im[1] = -P[0] - (Q[0] << 2) + P[1] + (Q[1] << 1) + (P[2] << 4) + Q[2];
    re[2] = (P[0] << 4) + Q[0] - (P[1] << 2) + Q[1] + (P[2] << 1) - Q[2];
    im[2] = -P[0] + (Q[0] << 4) - P[1] - (Q[1] << 2) + P[2] + (Q[2] << 1);
    for (int j = 0; j < 3; j++) {
        int64_t e1 = Ya[j] + Yb[j], e2 = Ya[j] - Yb[j];
        y[j] = e1 + re[j];
        y[j + 3] = e2 + im[j];
        y[j + 6] = e1 - re[j];
        y[j + 9] = e2 - im[j];
    }
}

/* poseidon.rs:242-262 mds_layer in the form of the x86-64 override (poseidon_goldilocks.rs:217-248):
 * every lane is split in 32-bit halves so that both half-products fit in 64 bits, the halves are
 * recombined to 96 bits and folded once; MDS_MATRIX_DIAG is 8 on lane 0 and 0 elsewhere. */
static inline void mds_layer(gl_t* s) {
    int64_t lo[PW], hi[PW], yl[PW], yh[PW];
    for (int i = 0; i < PW; i++) {
        lo[i] = (int64_t)(uint32_t)s[i];
        hi[i] = (int64_t)(s[i] >> 32);
    }
    mds_half(lo, yl);
    mds_half(hi, yh);
    yl[0] += lo[0] * (int64_t)MDS_MATRIX_DIAG[0];
    yh[0] += hi[0] * (int64_t)MDS_MATRIX_DIAG[0];
    for (int r = 0; r < PW; r++) {
        u128 sum = (u128)(uint64_t)yl[r] + ((u128)(uint64_t)yh[r] << 32);
        s[r] = gl_reduce96((uint64_t)sum, (uint32_t)(sum >> 64));
    }
}

/* The plain circulant sums of poseidon.rs:178-198 mds_row_shf (what mds_layer computes on targets
 * without the override); kept for ref_poseidon_permute(naive = 2), which tests/ compare with the
 * network above on random states. */
static inline void mds_layer_plain(gl_t* s) {
    uint64_t lo[2 * PW], hi[2 * PW], out[PW];
    for (int i = 0; i < PW; i++) {
        lo[i] = lo[i + PW] = (uint32_t)s[i];
        hi[i] = hi[i + PW] = s[i] >> 32;
    }
    for (int r = 0; r < PW; r++) {
        uint64_t al = 0, ah = 0;
        for (int i = 0; i < PW; i++) {
            al += lo[i + r] * MDS_MATRIX_CIRC[i];
            ah += hi[i + r] * MDS_MATRIX_CIRC[i];
        }
        al += lo[r] * MDS_MATRIX_DIAG[r];
        ah += hi[r] * MDS_MATRIX_DIAG[r];
        u128 sum = (u128)al + ((u128)ah << 32);
        out[r] = gl_reduce96((uint64_t)sum, (uint32_t)(sum >> 64));
    }
    memcpy(s, out, sizeof(out));
}

/* poseidon.rs:574-582 full_rounds */
static inline void full_rounds(gl_t* s, int* round) {
    for (int k = 0; k < HALF_FULL; k++) {
        constant_layer(s, *round);
        for (int i = 0; i < PW; i++) s[i] = sbox(s[i]);
        mds_layer(s);
        (*round)++;
    }
}

/* poseidon.rs:584-596 partial_rounds (fast form): partial_first_constant_layer :312,
 * mds_partial_layer_init :340, mds_partial_layer_fast :401 */
static inline void partial_rounds_fast(gl_t* s, int* round) {
    for (int i = 0; i < PW; i++) s[i] = gl_add(s[i], FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]);
    gl_t res[PW];
    res[0] = s[0];
    for (int c = 1; c < PW; c++) res[c] = 0;
    for (int r = 1; r < PW; r++)
        for (int c = 1; c < PW; c++)
            res[c] = gl_add(res[c], gl_mul(s[r], FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + (c - 1)]));
    memcpy(s, res, sizeof(res));
    const gl_t m00 = MDS_MATRIX_CIRC[0] + MDS_MATRIX_DIAG[0];
    for (int k = 0; k < N_PARTIAL; k++) {
        s[0] = gl_add(sbox(s[0]), FAST_PARTIAL_ROUND_CONSTANTS[k]);
        /* d = [M00 | w_hat] . state as an exact 160-bit sum (poseidon.rs:30-44 add_u160_u128).  The
         * low and high words of the 12 products are summed separately so that no carry has to be
         * tested: a carry out of 128 bits happens on about every other term of random data, and a
         * branch on it is mispredicted as often. */
        u128 first = (u128)s[0] * m00;
        u128 sum_lo = (uint64_t)first, sum_hi = (uint64_t)(first >> 64);
        for (int i = 1; i < PW; i++) {
            u128 term = (u128)s[i] * FAST_PARTIAL_ROUND_W_HATS[k * 11 + i - 1];
            sum_lo += (uint64_t)term;
            sum_hi += (uint64_t)(term >> 64);
        }
        sum_hi += (uint64_t)(sum_lo >> 64); /* < 2^68 */
        /* poseidon.rs:46-53 reduce_u160 */
        uint64_t lo_hi = (uint64_t)sum_hi, lo_lo = (uint64_t)sum_lo;
        uint32_t acc_hi = (uint32_t)(sum_hi >> 64);
        gl_t red_hi = gl_reduce96(lo_hi, acc_hi);
        gl_t d = gl_reduce128(((u128)red_hi << 64) + lo_lo);
        for (int i = 1; i < PW; i++) s[i] = gl_add(s[i], gl_mul(s[0], FAST_PARTIAL_ROUND_VS[k * 11 + i - 1]));
        s[0] = d;
    }
    *round += N_PARTIAL;
}

/* poseidon.rs:613-621 partial_rounds_naive */
static inline void partial_rounds_naive(gl_t* s, int* round) {
    for (int k = 0; k < N_PARTIAL; k++) {
        constant_layer(s, *round);
        s[0] = sbox(s[0]);
        mds_layer(s);
        (*round)++;
    }
}

/* poseidon.rs:599-609 poseidon */
static inline void poseidon(gl_t* s) {
    int round = 0;
    full_rounds(s, &round);
    partial_rounds_fast(s, &round);
    full_rounds(s, &round);
}

/* poseidon.rs:623-633 poseidon_naive */
static inline void poseidon_naive(gl_t* s) {
    int round = 0;
    full_rounds(s, &round);
    partial_rounds_naive(s, &round);
    full_rounds(s, &round);
}

/* poseidon_naive with the plain circulant MDS: no optimised form anywhere (the definition) */
static void poseidon_plain(gl_t* s) {
    for (int round = 0; round < 2 * HALF_FULL + N_PARTIAL; round++) {
        constant_layer(s, round);
        if (round < HALF_FULL || round >= HALF_FULL + N_PARTIAL)
            for (int i = 0; i < PW; i++) s[i] = sbox(s[i]);
        else
            s[0] = sbox(s[0]);
        mds_layer_plain(s);
    }
}

/* naive: 0 = poseidon (fast partial rounds), 1 = poseidon_naive, 2 = the plain definition */
API int ref_poseidon_permute(uint64_t* states /*[n][12]*/, size_t n, int naive) {
#pragma omp parallel for
    for (size_t i = 0; i < n; i++) {
        gl_t* s = states + i * PW;
        if (naive == 2)
            poseidon_plain(s);
        else if (naive)
            poseidon_naive(s);
        else
            poseidon(s);
        for (int k = 0; k < PW; k++) s[k] = gl_canon(s[k]);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * plonky2/src/hash/hashing.rs:119-146 hash_n_to_m_no_pad (overwrite-mode sponge, rate 8, 4 outputs)
 * plonky2/src/plonk/config.rs:55-66   hash_or_noop
 * plonky2/src/hash/hashing.rs:98-115  compress (two_to_one)
 * ---------------------------------------------------------------------------------------- */
static void hash_no_pad(const gl_t* in, size_t len, gl_t out[4]) {
    gl_t st[PW];
    memset(st, 0, sizeof(st));
    for (size_t off = 0; off < len; off += 8) {
        size_t c = len - off < 8 ? len - off : 8;
        memcpy(st, in + off, c * sizeof(gl_t)); /* set_from_slice: overwrite first c lanes */
        poseidon(st);
    }
    /* len == 0: no permutation is applied and the squeeze returns zeros (hashing.rs:133-141) */
    for (int i = 0; i < 4; i++) out[i] = gl_canon(st[i]);
}

static void hash_or_noop(const gl_t* in, size_t len, gl_t out[4]) {
    if (len * 8 <= 32) {
        for (int i = 0; i < 4; i++) out[i] = (size_t)i < len ? gl_canon(in[i]) : 0;
    } else {
        hash_no_pad(in, len, out);
    }
}

static void two_to_one(const gl_t l[4], const gl_t r[4], gl_t out[4]) {
    gl_t st[PW];
    memcpy(st, l, 4 * sizeof(gl_t));
    memcpy(st + 4, r, 4 * sizeof(gl_t));
    memset(st + 8, 0, 4 * sizeof(gl_t));
    poseidon(st);
    for (int i = 0; i < 4; i++) out[i] = gl_canon(st[i]);
}

API int ref_hash_or_noop(const uint64_t* in /*[n][len]*/, size_t n, size_t len, uint64_t* out /*[n][4]*/) {
#pragma omp parallel for
    for (size_t i = 0; i < n; i++) hash_or_noop(in + i * len, len, out + i * 4);
    return 0;
}

API int ref_two_to_one(const uint64_t* l, const uint64_t* r, size_t n, uint64_t* out) {
#pragma omp parallel for
    for (size_t i = 0; i < n; i++) two_to_one(l + 4 * i, r + 4 * i, out + 4 * i);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * plonky2/src/hash/merkle_tree.rs:69-96 fill_subtree (recursive; rayon::join -> omp task),
 * :98-132 fill_digests_buf, :135-166 MerkleTree::new.
 * leaves: row-major [n][len]; digests: [2(n - 2^cap)][4] in the reference's interleaved layout;
 * cap: [2^cap][4].
 * ---------------------------------------------------------------------------------------- */
static void fill_subtree(gl_t* digests_buf, size_t digests_len /*in hashes*/, const gl_t* leaves,
                         size_t n_leaves, size_t leaf_len, gl_t out[4]) {
    if (digests_len == 0) {
        hash_or_noop(leaves, leaf_len, out);
        return;
    }
    size_t half = digests_len / 2;
    gl_t* left_buf = digests_buf;                    /* half-1 hashes */
    gl_t* left_digest = digests_buf + (half - 1) * 4;
    gl_t* right_digest = digests_buf + half * 4;
    gl_t* right_buf = digests_buf + (half + 1) * 4;  /* half-1 hashes */
    size_t hl = n_leaves / 2;
    gl_t ld[4], rd[4];
    if (n_leaves >= 1024) {
#pragma omp task shared(ld)
        fill_subtree(left_buf, half - 1, leaves, hl, leaf_len, ld);
#pragma omp task shared(rd)
        fill_subtree(right_buf, half - 1, leaves + hl * leaf_len, hl, leaf_len, rd);
#pragma omp taskwait
    } else {
        fill_subtree(left_buf, half - 1, leaves, hl, leaf_len, ld);
        fill_subtree(right_buf, half - 1, leaves + hl * leaf_len, hl, leaf_len, rd);
    }
    memcpy(left_digest, ld, sizeof(ld));
    memcpy(right_digest, rd, sizeof(rd));
    two_to_one(ld, rd, out);
}

API int ref_merkle_build(const uint64_t* leaves, size_t n, size_t leaf_len, unsigned cap_height,
                         uint64_t* digests, uint64_t* cap) {
    int lg = log2_strict(n);
    if (lg < 0) return -2;                 /* log2_strict panic */
    if ((int)cap_height > lg) return -3;   /* merkle_tree.rs:136-142 assert */
    size_t n_cap = (size_t)1 << cap_height;
    size_t num_digests = 2 * (n - n_cap);
    if (num_digests == 0) { /* merkle_tree.rs:107-116 */
#pragma omp parallel for
        for (size_t i = 0; i < n; i++) hash_or_noop(leaves + i * leaf_len, leaf_len, cap + 4 * i);
        return 0;
    }
    size_t sub_digests = num_digests >> cap_height, sub_leaves = n >> cap_height;
#pragma omp parallel
#pragma omp single
    for (size_t s = 0; s < n_cap; s++) {
#pragma omp task
        fill_subtree(digests + s * sub_digests * 4, sub_digests, leaves + s * sub_leaves * leaf_len,
                     sub_leaves, leaf_len, cap + 4 * s);
    }
    return 0;
}

/* merkle_tree.rs:173-207 prove: siblings [num_layers][4] */
API int ref_merkle_prove(const uint64_t* digests, size_t n, unsigned cap_height, size_t leaf_index,
                         uint64_t* siblings) {
    int lg = log2_strict(n);
    int num_layers = lg - (int)cap_height;
    size_t num_digests = 2 * (n - ((size_t)1 << cap_height));
    size_t tree_index = leaf_index >> num_layers;
    size_t tree_len = num_digests >> cap_height;
    const uint64_t* tree = digests + tree_len * tree_index * 4;
    size_t pair_index = leaf_index & (((size_t)1 << num_layers) - 1);
    for (int i = 0; i < num_layers; i++) {
        size_t parity = pair_index & 1;
        pair_index >>= 1;
        size_t siblings_index = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
        size_t sibling_index = 2 * siblings_index + (1 - parity);
        memcpy(siblings + 4 * i, tree + 4 * sibling_index, 32);
    }
    return num_layers;
}

/* plonky2/src/hash/merkle_proofs.rs:54-77 verify_merkle_proof_to_cap; returns 1 if it verifies */
API int ref_merkle_verify(const uint64_t* leaf, size_t leaf_len, size_t leaf_index,
                          const uint64_t* cap, const uint64_t* siblings, unsigned n_siblings) {
    gl_t cur[4], nxt[4];
    hash_or_noop(leaf, leaf_len, cur);
    size_t index = leaf_index;
    for (unsigned i = 0; i < n_siblings; i++) {
        if (index & 1)
            two_to_one(siblings + 4 * i, cur, nxt);
        else
            two_to_one(cur, siblings + 4 * i, nxt);
        memcpy(cur, nxt, sizeof(cur));
        index >>= 1;
    }
    return memcmp(cur, cap + 4 * index, 32) == 0;
}

/* ------------------------------------------------------------------------------------------
 * plonky2/src/util/mod.rs:22-28 transpose + util/src/lib.rs:188 reverse_index_bits_in_place(leaves):
 * leaves[brev(i)][j] = lde[j][i].  salts: NULL or [salt_w][N] extra columns (oracle.rs:119-123,
 * the reference draws them from OsRng; the caller supplies them so runs are reproducible).
 * ---------------------------------------------------------------------------------------- */
API int ref_transpose_bitrev(const uint64_t* lde /*[w][n]*/, size_t w, size_t n, uint64_t* leaves /*[n][w]*/) {
    int lg = log2_strict(n);
    if (lg < 0) return -2;
#pragma omp parallel for
    for (size_t i = 0; i < n; i++) {
        size_t bi = reverse_bits(i, (unsigned)lg);
        for (size_t j = 0; j < w; j++) leaves[bi * w + j] = lde[j * n + i];
    }
    return 0;
}

/* plonky2/src/fri/oracle.rs:68-98 from_coeffs.  Outputs: leaves [N][w+salt_w], digests, cap. */
API int ref_commit_from_coeffs(const uint64_t* coeffs /*[w][d]*/, size_t w, unsigned lg_d, unsigned rate_bits,
                               unsigned cap_height, const uint64_t* salts, size_t salt_w,
                               uint64_t* leaves, uint64_t* digests, uint64_t* cap) {
    int lg_n = (int)(lg_d + rate_bits);
    size_t n = (size_t)1 << lg_n;
    size_t wt = w + salt_w;
    uint64_t* lde = (uint64_t*)malloc(wt * n * sizeof(uint64_t));
    if (!lde) return -4;
    int rc = ref_coset_lde_batch(coeffs, w, lg_d, rate_bits, GL_COSET_SHIFT, lde);
    if (rc) { free(lde); return rc; }
    if (salt_w) memcpy(lde + w * n, salts, salt_w * n * sizeof(uint64_t));
    ref_transpose_bitrev(lde, wt, n, leaves);
    free(lde);
    return ref_merkle_build(leaves, n, wt, cap_height, digests, cap);
}

/* plonky2/src/fri/oracle.rs:43-65 from_values: IFFT every column (values is overwritten with the
 * coefficients, which the reference keeps as `polynomials`), then from_coeffs. */
API int ref_commit_from_values(uint64_t* values /*[w][d] in: values, out: coeffs*/, size_t w, unsigned lg_d,
                               unsigned rate_bits, unsigned cap_height, const uint64_t* salts, size_t salt_w,
                               uint64_t* leaves, uint64_t* digests, uint64_t* cap) {
    ref_fft_batch(values, w, lg_d, 1, 0);
    return ref_commit_from_coeffs(values, w, lg_d, rate_bits, cap_height, salts, salt_w, leaves, digests, cap);
}

/* Scalar field ops exposed for the reference's field-op grid test (prime_field_testing.rs:78-125). */
API uint64_t ref_gl_add(uint64_t a, uint64_t b) { return gl_canon(gl_add(a, b)); }
API uint64_t ref_gl_sub(uint64_t a, uint64_t b) { return gl_canon(gl_sub(a, b)); }
API uint64_t ref_gl_mul(uint64_t a, uint64_t b) { return gl_canon(gl_mul(a, b)); }
API uint64_t ref_gl_pow(uint64_t a, uint64_t e) { return gl_canon(gl_pow(a, e)); }
API uint64_t ref_gl_inverse_2exp(unsigned e) { return gl_inverse_2exp(e); }
API uint64_t ref_gl_primitive_root_of_unity(unsigned lg) { return gl_canon(gl_primitive_root_of_unity(lg)); }

/* Naive evaluation P(x) by Horner (polynomial/mod.rs eval) -- used by the FFT==naive tests. */
API uint64_t ref_poly_eval(const uint64_t* coeffs, size_t n, uint64_t x) {
    gl_t acc = 0;
    for (size_t i = n; i-- > 0;) acc = gl_add(gl_mul(acc, x), coeffs[i]);
    return gl_canon(acc);
}

API int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Thread count of the OpenMP loops (the reference uses the global rayon pool = all cores, maybe_rayon/src/lib.rs).
 * torchrun exports OMP_NUM_THREADS=1 to its children: bench.py's reference arm overrides that here. */
API void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* FRI opening proof arithmetic */
#include "fri.c"
