// Quadratic extension of Goldilocks, F[X]/(X^2 - 7), for host and device.
//
// GPU counterpart of the reference's QuadraticExtension<GoldilocksField>
// (field/src/extension/quadratic.rs:184-192 mul, field/src/goldilocks_extensions.rs:14-28: W = 7).
// An element is the pair (a, b) = a + b*X, stored like the reference's [F; 2] (to_basefield_array).
// Device values are loose u64 like everywhere in the engine; canon() where they leave it.
#pragma once
#include <stdint.h>

#include "gl64.cuh"

namespace gl {

struct ext2 {
    uint64_t a, b;
};

// ---- 160-bit lazy accumulator: sum of 64x64 products, reduced once ---------------------------------
// The reference reduces such sums the same way in its extension-field kernels (goldilocks_extensions.rs:120-180
// accumulate u128 products with a carry count and reduce once); the result is the same field element.
struct acc160 {
    uint64_t lo, hi;
    uint32_t top;
};

__device__ __forceinline__ void acc_zero(acc160& A) {
    A.lo = 0;
    A.hi = 0;
    A.top = 0;
}

// A += x * y  (any u64 x, y)
__device__ __forceinline__ void acc_mac(acc160& A, uint64_t x, uint64_t y) {
    unsigned __int128 q = (unsigned __int128)x * y;
    uint64_t ql = (uint64_t)q, qh = (uint64_t)(q >> 64);
    asm("{\n\t"
        "add.cc.u64  %0, %0, %3;\n\t"
        "addc.cc.u64 %1, %1, %4;\n\t"
        "addc.u32    %2, %2, 0;\n\t"
        "}"
        : "+l"(A.lo), "+l"(A.hi), "+r"(A.top)
        : "l"(ql), "l"(qh));
}

// lo + hi*2^64 + top*2^128 -> loose u64, using 2^128 = 2^96 * 2^32 = -2^32 (mod p); top*2^32 <= p - 1 is canonical
__device__ __forceinline__ uint64_t acc_reduce(const acc160& A) {
    return sub_lc(reduce128(A.lo, A.hi), (uint64_t)A.top << 32);
}

// loose x loose -> loose.  c0 = a0 b0 + 7 a1 b1, c1 = a0 b1 + a1 b0, each as ONE lazy reduction
__device__ __forceinline__ ext2 ext_mul(ext2 x, ext2 y) {
    acc160 c0, c1;
    acc_zero(c0);
    acc_zero(c1);
    uint64_t t = canon(mul(x.b, y.b));
    acc_mac(c0, x.a, y.a);
    acc_mac(c0, t, 7);
    acc_mac(c1, x.a, y.b);
    acc_mac(c1, x.b, y.a);
    ext2 r;
    r.a = acc_reduce(c0);
    r.b = acc_reduce(c1);
    return r;
}

// loose + loose -> loose (canonicalises one operand)
__device__ __forceinline__ ext2 ext_add(ext2 x, ext2 y) {
    ext2 r;
    r.a = add_lc(x.a, canon(y.a));
    r.b = add_lc(x.b, canon(y.b));
    return r;
}

__device__ __forceinline__ ext2 ext_canon(ext2 x) {
    ext2 r;
    r.a = canon(x.a);
    r.b = canon(x.b);
    return r;
}

__device__ __forceinline__ ext2 ext_pow(ext2 base, uint64_t e) {
    ext2 acc = {1, 0};
    while (e) {
        if (e & 1) acc = ext_mul(acc, base);
        base = ext_mul(base, base);
        e >>= 1;
    }
    return ext_canon(acc);
}

}  // namespace gl

// ---- host-side scalar arithmetic (setup only: points, shifts, powers of a few challenges) -------------
namespace glh {
constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
struct ext2 {
    uint64_t a, b;
};
static inline uint64_t mulmod(uint64_t a, uint64_t b) { return (uint64_t)((unsigned __int128)a * b % P); }
static inline uint64_t addmod(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a + b) % P); }
static inline uint64_t powmod(uint64_t a, uint64_t e) {
    uint64_t acc = 1;
    a %= P;
    for (; e; e >>= 1) {
        if (e & 1) acc = mulmod(acc, a);
        a = mulmod(a, a);
    }
    return acc;
}
static inline ext2 ext_mul(ext2 x, ext2 y) {
    ext2 r;
    r.a = addmod(mulmod(x.a, y.a), mulmod(7, mulmod(x.b, y.b)));
    r.b = addmod(mulmod(x.a, y.b), mulmod(x.b, y.a));
    return r;
}
static inline ext2 ext_pow(ext2 x, uint64_t e) {
    ext2 acc = {1, 0};
    x.a %= P;
    x.b %= P;
    for (; e; e >>= 1) {
        if (e & 1) acc = ext_mul(acc, x);
        x = ext_mul(x, x);
    }
    return acc;
}
}  // namespace glh
