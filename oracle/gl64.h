/*
 * ORACLE (test infrastructure, NOT the product): CPU restatement of the reference's
 * Goldilocks field arithmetic.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use anything under oracle/.
 *
 * Follows /root/reference/field/src/goldilocks_field.rs:
 *   representation  : any u64 (non-canonical allowed)            :23-25, 33-37
 *   canonicalise    : one conditional subtract                   :171-178
 *   add / sub       : wrapping op + EPSILON fix-ups              :199-258
 *   mul             : u128 product + reduce128                   :267-274, 356-369
 *   reduce96        :                                            :347-352
 * and field/src/types.rs: exp (square-and-multiply), inverse_2exp :227-262,
 *   primitive_root_of_unity :268-272.
 */
#pragma once
#include <stdint.h>

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL /* 2^64 mod p */
#define GL_TWO_ADICITY 32
#define GL_POWER_OF_TWO_GENERATOR 1753635133440165772ULL
#define GL_COSET_SHIFT 7ULL

typedef uint64_t gl_t;
typedef unsigned __int128 u128;

static inline gl_t gl_canon(gl_t x) { return x >= GL_P ? x - GL_P : x; }

static inline gl_t gl_add(gl_t a, gl_t b) {
    uint64_t s = a + b;
    uint64_t over = (uint64_t)0 - (uint64_t)(s < a); /* all-ones on wrap */
    uint64_t s2 = s + (over & GL_EPS);
    if (__builtin_expect(s2 < s, 0)) s2 += GL_EPS; /* double overflow: only if both inputs > p (rare) */
    return s2;
}

static inline gl_t gl_sub(gl_t a, gl_t b) {
    uint64_t d = a - b;
    uint64_t under = (uint64_t)0 - (uint64_t)(a < b);
    uint64_t d2 = d - (under & GL_EPS);
    if (__builtin_expect(d2 > d, 0)) d2 -= GL_EPS; /* double underflow (rare) */
    return d2;
}

/* x + y mod 2^64 with the wrap folded back in; exact when x + y < 2^64 + p. */
/* Branch-free like the reference's x86 `add; sbb` trick (goldilocks_field.rs:309-334): the wrap
 * happens ~50% of the time on random data, so a branch here would be mispredicted every other call. */
static inline uint64_t gl_add_no_canon(uint64_t x, uint64_t y) {
    uint64_t r, adj;
#if defined(__x86_64__)
    __asm__("add %2, %0\n\tsbb %k1, %k1" : "=&r"(r), "=&r"(adj) : "r"(y), "0"(x) : "cc");
#else
    r = x + y;
    adj = (uint64_t)0 - (uint64_t)(r < x);
    adj &= GL_EPS;
#endif
    return r + adj;
}

static inline gl_t gl_reduce96(uint64_t lo, uint32_t hi) {
    return gl_add_no_canon(lo, (uint64_t)hi * GL_EPS);
}

static inline gl_t gl_reduce128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t t0 = lo - hi_hi;
    if (__builtin_expect(lo < hi_hi, 0)) { t0 -= GL_EPS; __asm__("" : "+r"(t0)); } /* "exceedingly rare" (goldilocks_field.rs:362) */
    uint64_t t1 = hi_lo * GL_EPS;
    return gl_add_no_canon(t0, t1);
}

static inline gl_t gl_mul(gl_t a, gl_t b) { return gl_reduce128((u128)a * (u128)b); }
static inline gl_t gl_sqr(gl_t a) { return gl_mul(a, a); }

static inline gl_t gl_pow(gl_t base, uint64_t e) {
    gl_t acc = 1;
    while (e) {
        if (e & 1) acc = gl_mul(acc, base);
        base = gl_sqr(base);
        e >>= 1;
    }
    return acc;
}

/* types.rs:227-262 inverse_2exp, for exp <= TWO_ADICITY: 2^-e = p - ((p-1) >> e). */
static inline gl_t gl_inverse_2exp(unsigned e) { return GL_P - ((GL_P - 1) >> e); }

/* types.rs:268-272 */
static inline gl_t gl_primitive_root_of_unity(unsigned n_log) {
    gl_t b = GL_POWER_OF_TWO_GENERATOR;
    for (unsigned i = n_log; i < GL_TWO_ADICITY; i++) b = gl_sqr(b);
    return b;
}
