/*
 * ORACLE (test infrastructure, NOT the product): CPU restatement of the arithmetic of the reference's FRI opening
 * proof -- the consumers of a committed batch's coefficients.  #included by oracle.c (one
 * translation unit, so the field / FFT restatement above is shared).  The protocol glue around these (Challenger,
 * query rounds, verifier) is restated in oracle/fri_ref.py.
 *
 * Follows
 *   field/src/extension/quadratic.rs:184-192, goldilocks_extensions.rs:14-28   F[X]/(X^2 - 7) multiplication
 *   field/src/polynomial/mod.rs:157-163 eval (Horner), :297-310 to_extension / mul_extension
 *   plonky2/src/util/reducing.rs:84-96 reduce_polys_base, :104-107 shift_poly
 *   field/src/polynomial/division.rs:75-88 divide_by_linear
 *   field/src/polynomial/mod.rs:201-203 lde, :277-295 coset_fft
 *   plonky2/src/plonk/plonk_common.rs:116-128 reduce_with_powers (fri/prover.rs:93-101)
 * Extension elements are [a, b] = a + b X, Vec<F::Extension> = flat u64[2n].
 *
 * Parity: pinned by the reference's field-extension test identities restated in tests/test_fri_oracle.py
 * (quadratic.rs tests via extension-field algebra against Python big-int arithmetic; division.rs tests:
 * divide_by_linear == long division; polynomial eval == FFT values) and by the restated FRI verifier accepting the
 * restated prover's proofs (fri/verifier.rs).  The reference holds no golden FRI proofs.
 */

typedef struct {
    gl_t a, b;
} ext_t;

static inline ext_t ext_mul(ext_t x, ext_t y) { /* quadratic.rs:184-192, W = 7 */
    ext_t r;
    r.a = gl_add(gl_mul(x.a, y.a), gl_mul(7, gl_mul(x.b, y.b)));
    r.b = gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a));
    return r;
}
static inline ext_t ext_add(ext_t x, ext_t y) {
    ext_t r = {gl_add(x.a, y.a), gl_add(x.b, y.b)};
    return r;
}
static inline ext_t ext_mul_base(ext_t x, gl_t s) { /* scalar_mul */
    ext_t r = {gl_mul(x.a, s), gl_mul(x.b, s)};
    return r;
}
static inline ext_t ext_pow(ext_t b, uint64_t e) { /* types.rs exp_u64 */
    ext_t acc = {1, 0};
    while (e) {
        if (e & 1) acc = ext_mul(acc, b);
        b = ext_mul(b, b);
        e >>= 1;
    }
    return acc;
}
static inline void ext_store(uint64_t* out, ext_t v) {
    out[0] = gl_canon(v.a);
    out[1] = gl_canon(v.b);
}

API void ref_ext_mul(const uint64_t x[2], const uint64_t y[2], uint64_t out[2]) {
    ext_t a = {x[0], x[1]}, b = {y[0], y[1]};
    ext_store(out, ext_mul(a, b));
}

API void ref_ext_pow(const uint64_t x[2], uint64_t e, uint64_t out[2]) {
    ext_t a = {x[0], x[1]};
    ext_store(out, ext_pow(a, e));
}

/* proof.rs:316-322 eval_commitment: p.to_extension().eval(z) for every row of coeffs [w][d]; out [w][2].
 * polynomial/mod.rs:157-163: coeffs.iter().rev().fold(ZERO, |acc, &c| acc * x + c) */
API void ref_eval_base_polys_ext(const uint64_t* coeffs, size_t w, size_t d, const uint64_t z[2], uint64_t* out) {
    ext_t zz = {z[0], z[1]};
#pragma omp parallel for schedule(dynamic)
    for (size_t j = 0; j < w; j++) {
        ext_t acc = {0, 0};
        for (size_t i = d; i-- > 0;) {
            acc = ext_mul(acc, zz);
            acc.a = gl_add(acc.a, coeffs[j * d + i]); /* to_extension(): (c, 0) */
        }
        ext_store(out + 2 * j, acc);
    }
}

/* the same Horner evaluation of an extension polynomial [n][2] (verifier.rs:238 final_poly.eval) */
API void ref_ext_poly_eval(const uint64_t* coeffs, size_t n, const uint64_t z[2], uint64_t out[2]) {
    ext_t zz = {z[0], z[1]}, acc = {0, 0};
    for (size_t i = n; i-- > 0;) {
        ext_t c = {coeffs[2 * i], coeffs[2 * i + 1]};
        acc = ext_add(ext_mul(acc, zz), c);
    }
    ext_store(out, acc);
}

/* reducing.rs:84-96 reduce_polys_base: sum_j alpha^j * polys[j] over rows of [k][d]; out [d][2].
 * (poly.mul_extension(base_power): every coefficient times the extension scalar) */
API void ref_reduce_polys_base(const uint64_t* polys, size_t k, size_t d, const uint64_t alpha[2], uint64_t* out) {
    ext_t a = {alpha[0], alpha[1]};
    ext_t* pw = (ext_t*)malloc((k ? k : 1) * sizeof(ext_t));
    ext_t cur = {1, 0};
    for (size_t j = 0; j < k; j++) { /* base.powers() */
        pw[j] = cur;
        cur = ext_mul(cur, a);
    }
#pragma omp parallel for
    for (size_t i = 0; i < d; i++) {
        ext_t acc = {0, 0};
        for (size_t j = 0; j < k; j++) acc = ext_add(acc, ext_mul_base(pw[j], polys[j * d + i]));
        ext_store(out + 2 * i, acc);
    }
    free(pw);
}

/* division.rs:75-88 divide_by_linear, then `quotient.coeffs.push(ZERO)` (oracle.rs:197): out [n][2] */
API void ref_ext_divide_by_linear(const uint64_t* coeffs, size_t n, const uint64_t z[2], uint64_t* out) {
    ext_t zz = {z[0], z[1]}, acc = {0, 0};
    /* bs (in reverse order) = scan(acc = acc * z + c) from the top coefficient; drop the last (the remainder) */
    for (size_t i = n; i-- > 0;) {
        ext_t c = {coeffs[2 * i], coeffs[2 * i + 1]};
        acc = ext_add(ext_mul(acc, zz), c); /* = bs entry for coefficient i: quotient coefficient i - 1 */
        if (i > 0) ext_store(out + 2 * (i - 1), acc);
    }
    if (n > 0) {
        out[2 * (n - 1)] = 0;
        out[2 * (n - 1) + 1] = 0;
    }
}

/* reducing.rs:104-107 shift_poly (acc *= alpha^count) followed by oracle.rs:199 final_poly += quotient */
API void ref_ext_scale_add(uint64_t* acc, const uint64_t s[2], const uint64_t* q, size_t n) {
    ext_t ss = {s[0], s[1]};
    for (size_t i = 0; i < n; i++) {
        ext_t a = {acc[2 * i], acc[2 * i + 1]}, qq = {q[2 * i], q[2 * i + 1]};
        ext_store(acc + 2 * i, ext_add(ext_mul(a, ss), qq));
    }
}

/* p.lde(rate_bits).coset_fft(shift.into()) for an extension polynomial [d][2] -> values [N][2], natural order
 * (oracle.rs:202-207).  coset_fft (polynomial/mod.rs:277-295): c_i *= shift^i in the extension, then the FFT over
 * the extension with the root table of F::Extension::primitive_root_of_unity(lg N).  For lg N <= 32 that root is
 * the base field's (quadratic.rs:70-74: EXT_POWER_OF_TWO_GENERATOR^2 = [POWER_OF_TWO_GENERATOR, 0], checked in
 * tests/test_fri_oracle.py), and (a, b) * (r, 0) = (a r, b r), so every butterfly acts on the two components
 * separately: the extension FFT IS fft_classic on each component. */
API int ref_ext_coset_lde(const uint64_t* coeffs, unsigned lg_d, unsigned rate_bits, uint64_t shift, uint64_t* out) {
    int lg_n = (int)(lg_d + rate_bits);
    if (lg_n > GL_TWO_ADICITY) return -1;
    size_t d = (size_t)1 << lg_d, n = (size_t)1 << lg_n;
    gl_t* comp = (gl_t*)calloc(2 * n, sizeof(gl_t)); /* lde(): zero padding */
    if (!comp) return -4;
    ext_t pw = {1, 0}, sh = {shift, 0};
    for (size_t i = 0; i < d; i++) {
        ext_t c = {coeffs[2 * i], coeffs[2 * i + 1]};
        ext_t m = ext_mul(pw, c);
        comp[i] = m.a;
        comp[n + i] = m.b;
        pw = ext_mul(pw, sh);
    }
    root_table_t t;
    root_table_build(&t, lg_n);
    fft_classic(comp, 0, lg_n, &t);
    fft_classic(comp + n, 0, lg_n, &t);
    root_table_free(&t);
    for (size_t i = 0; i < n; i++) {
        out[2 * i] = gl_canon(comp[i]);
        out[2 * i + 1] = gl_canon(comp[n + i]);
    }
    free(comp);
    return 0;
}

/* prover.rs:93-101: coeffs.par_chunks_exact(arity).map(|chunk| reduce_with_powers(chunk, beta));
 * plonk_common.rs:116-128: sum = sum * alpha + term over the chunk in reverse.  in [n][2] -> out [n / arity][2] */
API void ref_fri_fold(const uint64_t* coeffs, size_t n, size_t arity, const uint64_t beta[2], uint64_t* out) {
    ext_t b = {beta[0], beta[1]};
    size_t n_out = n / arity; /* chunks_exact drops a short tail */
    for (size_t i = 0; i < n_out; i++) {
        ext_t sum = {0, 0};
        for (size_t j = arity; j-- > 0;) {
            ext_t c = {coeffs[2 * (i * arity + j)], coeffs[2 * (i * arity + j) + 1]};
            sum = ext_add(ext_mul(sum, b), c);
        }
        ext_store(out + 2 * i, sum);
    }
}
