// Goldilocks field (p = 2^64 - 2^32 + 1) device arithmetic for sm_100a.
//
// GPU counterpart of the reference's field/src/goldilocks_field.rs (add/sub :199-258,
// mul = u128 product + reduce128 :267-274,356-369).  The reference keeps NON-canonical u64s and
// canonicalises on compare/serialise (:33-37,171-178); this engine accepts any u64 at its
// boundary, canonicalises on load, and emits canonical values, so parity = plain equality.
//
// B200 has no 64-bit integer multiplier: a 64x64->128 product is four IMAD.WIDE.U32
// (32x32+64->64) on the FMA-heavy pipe, the reduction is IADD3/LOP3 work on the ALU pipe.
// Identities used (2^64 = 2^32 - 1, 2^96 = -1 mod p):
//   r3*2^96 + r2*2^64 + r1*2^32 + r0  ==  (r0 - r2 - r3) + (r1 + r2)*2^32   (mod p)
#pragma once
#include <stdint.h>

namespace gl {

constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
constexpr uint64_t EPS = 0xFFFFFFFFULL;  // 2^64 mod p

__device__ __forceinline__ uint64_t canon(uint64_t x) { return x >= P ? x - P : x; }

// a, b canonical -> canonical.
__device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) {
    uint64_t nb = P - b;  // in (0, p]
    uint64_t r = a - nb;
    return a < nb ? r + P : r;
}

__device__ __forceinline__ uint64_t sub(uint64_t a, uint64_t b) {
    uint64_t r = a - b;
    return a < b ? r + P : r;
}

__device__ __forceinline__ uint64_t neg(uint64_t a) { return a ? P - a : 0; }

// 64x64 -> 128 as four 32x32+64 multiply-adds; (hi, lo) returned through references.
__device__ __forceinline__ void mul_wide(uint64_t a, uint64_t b, uint64_t& lo, uint64_t& hi) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
    uint32_t b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    uint64_t t0 = (uint64_t)a0 * b0;
    uint64_t t1 = (uint64_t)a0 * b1 + (t0 >> 32);            // cannot overflow
    uint64_t t2 = (uint64_t)a1 * b0 + (uint32_t)t1;          // cannot overflow
    uint64_t t3 = (uint64_t)a1 * b1 + (t1 >> 32) + (t2 >> 32);  // cannot overflow
    lo = (t2 << 32) | (uint32_t)t0;
    hi = t3;
}

// 64-bit square: three products.
__device__ __forceinline__ void sqr_wide(uint64_t a, uint64_t& lo, uint64_t& hi) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
    uint64_t t0 = (uint64_t)a0 * a0;
    uint64_t m = (uint64_t)a0 * a1;
    uint64_t t1 = m + (t0 >> 32);           // < 2^64
    uint64_t t2 = m + (uint32_t)t1;         // < 2^64
    uint64_t t3 = (uint64_t)a1 * a1 + (t1 >> 32) + (t2 >> 32);
    lo = (t2 << 32) | (uint32_t)t0;
    hi = t3;
}

// Any 128-bit value -> canonical.  Mirrors reduce128 (goldilocks_field.rs:356-369).
__device__ __forceinline__ uint64_t reduce128(uint64_t lo, uint64_t hi) {
    uint32_t hi_hi = (uint32_t)(hi >> 32), hi_lo = (uint32_t)hi;
    // t0 = lo - hi_hi (mod p), lo arbitrary u64
    uint64_t t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= EPS;  // borrow: -2^64 == -EPS; cannot underflow
    // t1 = hi_lo * (2^32 - 1) = (hi_lo << 32) - hi_lo  < 2^64
    uint64_t t1 = ((uint64_t)hi_lo << 32) - hi_lo;
    uint64_t r = t0 + t1;
    if (r < t1) r += EPS;  // wrapped: +2^64 == +EPS; cannot wrap again
    return canon(r);
}

__device__ __forceinline__ uint64_t mul(uint64_t a, uint64_t b) {
    uint64_t lo, hi;
    mul_wide(a, b, lo, hi);
    return reduce128(lo, hi);
}

__device__ __forceinline__ uint64_t sqr(uint64_t a) {
    uint64_t lo, hi;
    sqr_wide(a, lo, hi);
    return reduce128(lo, hi);
}

__device__ __forceinline__ uint64_t pow(uint64_t base, uint64_t e) {
    uint64_t acc = 1;
    while (e) {
        if (e & 1) acc = mul(acc, base);
        base = sqr(base);
        e >>= 1;
    }
    return acc;
}

// ---- unreduced accumulation: acc (192-bit: a0,a1,a2) += x * c --------------------------------
// Used where many products with compile-time constants are summed before ONE reduction
// (the reference does the same with its u160 accumulator, poseidon.rs:38-53,401-431).
struct Acc192 {
    uint64_t lo, mid;  // 128-bit
    uint32_t hi;       // overflow counter (sum of < 2^32 products fits)
};

__device__ __forceinline__ void acc_init(Acc192& a) { a.lo = 0; a.mid = 0; a.hi = 0; }

__device__ __forceinline__ void acc_mac(Acc192& a, uint64_t x, uint64_t c) {
    uint64_t lo, hi;
    mul_wide(x, c, lo, hi);
    uint64_t nlo = a.lo + lo;
    uint64_t carry = nlo < lo;
    uint64_t nmid = a.mid + hi;
    uint32_t c2 = nmid < hi;
    uint64_t nmid2 = nmid + carry;
    c2 += nmid2 < carry;
    a.lo = nlo;
    a.mid = nmid2;
    a.hi += c2;
}

__device__ __forceinline__ void acc_add64(Acc192& a, uint64_t x) {
    uint64_t nlo = a.lo + x;
    uint64_t carry = nlo < x;
    uint64_t nmid = a.mid + carry;
    a.hi += nmid < carry;
    a.lo = nlo;
    a.mid = nmid;
}

// value = lo + mid*2^64 + hi*2^128 ; 2^128 = 2^64*2^64 == (2^32-1)^2 = 2^64 - 2^33 + 1 == -2^32 (mod p)
__device__ __forceinline__ uint64_t acc_reduce(const Acc192& a) {
    // fold hi*2^128 == -hi*2^32 into the 128-bit part first: reduce128(lo, mid) - hi*2^32
    uint64_t r = reduce128(a.lo, a.mid);
    uint64_t h = (uint64_t)a.hi << 32;  // hi < 2^32 -> h < 2^64, may be >= p
    return sub(r, canon(h));
}

}  // namespace gl
