// Poseidon-12 permutation over Goldilocks, one permutation per thread, state in registers.
//
// GPU counterpart of Poseidon::poseidon (plonky2/src/hash/poseidon.rs:599-609):
//   4 full rounds -> 22 partial rounds -> 4 full rounds, S-box x^7, MDS = circ(17,15,41,16,2,28,
//   13,13,39,18,34,20) + diag(8,0,..) (poseidon_goldilocks.rs:24-25).
// Output is bit-identical (after canonicalisation) to the reference for every input; the
// reference's own `consistency` test (poseidon.rs:777-790) licenses any algebraically equal
// evaluation order, and tests/ pin this kernel to the reference KATs (poseidon_goldilocks.rs:461-482).
//
// Design for the B200 integer pipes (no tensor cores: 64-bit modular arithmetic):
//  * full-round MDS: every lane is split in 32-bit halves; each output row is two chains of
//    IMAD.WIDE.U32 (32x6-bit MAC into a 64-bit accumulator, < 2^42), one 96-bit fold per row.
//    The NEXT round's constant is pre-loaded into the accumulator, so constant_layer is free.
//  * partial rounds: "lazy" form (tests/golden/make_golden.py:derive_lazy_tables).  Lanes 1..11
//    are never materialised during the 22 rounds; each round's lane-0 value is a dot product of
//    compile-time constants with the 11 post-init lanes and the earlier S-box outputs,
//    accumulated unreduced (one reduction per round instead of twelve).
#pragma once
#include "gl64.cuh"
#include "poseidon_constants.cuh"

namespace pcs {

constexpr int SPONGE_WIDTH = 12;
constexpr int SPONGE_RATE = 8;

__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
    uint64_t x2 = gl::sqr(x);
    uint64_t x4 = gl::sqr(x2);
    uint64_t x3 = gl::mul(x, x2);
    return gl::mul(x3, x4);
}

// value = lo + hi*2^64 (hi < 2^32) -> canonical
__device__ __forceinline__ uint64_t reduce96(uint64_t lo, uint32_t hi) {
    uint64_t t1 = ((uint64_t)hi << 32) - hi;  // hi * EPS
    uint64_t r = lo + t1;
    if (r < t1) r += gl::EPS;
    return gl::canon(r);
}

// s <- MDS * s  (+ optional constant vector, added for free inside the accumulators)
template <bool ADD_RC>
__device__ __forceinline__ void mds_full(uint64_t (&s)[12], const uint64_t* __restrict__ rc) {
    constexpr uint32_t CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        lo[i] = (uint32_t)s[i];
        hi[i] = (uint32_t)(s[i] >> 32);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        uint64_t al, ah;
        if (ADD_RC) {
            uint64_t c = rc[r];
            al = (uint32_t)c;
            ah = c >> 32;
        } else {
            al = 0;
            ah = 0;
        }
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al += (uint64_t)lo[(i + r) % 12] * CIRC[i];
            ah += (uint64_t)hi[(i + r) % 12] * CIRC[i];
        }
        if (r == 0) {
            al += (uint64_t)lo[0] * 8u;
            ah += (uint64_t)hi[0] * 8u;
        }
        // value = al + ah*2^32, al, ah < 2^43
        uint64_t hs = ah << 32;
        uint64_t l = al + hs;
        uint32_t h = (uint32_t)(ah >> 32) + (l < hs);
        s[r] = reduce96(l, h);
    }
}

// The permutation.  Input lanes must be canonical; output lanes are canonical.
__device__ __forceinline__ void poseidon12(uint64_t (&s)[12]) {
    using namespace pconst;
    // ---- first 4 full rounds; RC of round 0 added explicitly, the rest fused into the MDS ----
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl::add(s[i], RC[i]);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        // after round 3 the next additive constant is FIRST_RC (partial_first_constant_layer)
        mds_full<true>(s, r < 3 ? &RC[12 * (r + 1)] : FIRST_RC);
    }

    // ---- mds_partial_layer_init (poseidon.rs:340-366): t_c = sum_r s_r * INIT[r][c], c,r in 1..11
    uint64_t t[12];
    t[0] = s[0];
#pragma unroll
    for (int c = 1; c < 12; c++) {
        gl::Acc192 acc;
        gl::acc_init(acc);
#pragma unroll
        for (int r = 1; r < 12; r++) gl::acc_mac(acc, s[r], INIT_T[(c - 1) * 11 + (r - 1)]);
        t[c] = gl::acc_reduce(acc);
    }

    // ---- 22 lazy partial rounds ----
    uint64_t x[22];
    uint64_t s0 = t[0];
#pragma unroll
    for (int k = 0; k < 22; k++) {
        uint64_t xk = gl::add(sbox7(s0), PARTIAL_RC[k]);
        x[k] = xk;
        gl::Acc192 acc;
        gl::acc_init(acc);
        gl::acc_mac(acc, xk, 25);  // M00 = circ[0] + diag[0]
#pragma unroll
        for (int i = 1; i < 12; i++) gl::acc_mac(acc, t[i], W_HATS[k * 11 + i - 1]);
#pragma unroll
        for (int q = 0; q < k; q++) gl::acc_mac(acc, x[q], LAZY_C[k * (k - 1) / 2 + q]);
        s0 = gl::acc_reduce(acc);
    }
    s[0] = s0;
#pragma unroll
    for (int i = 1; i < 12; i++) {
        gl::Acc192 acc;
        gl::acc_init(acc);
        gl::acc_add64(acc, t[i]);
#pragma unroll
        for (int q = 0; q < 22; q++) gl::acc_mac(acc, x[q], VS[q * 11 + i - 1]);
        s[i] = gl::acc_reduce(acc);
    }

    // ---- last 4 full rounds ----
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl::add(s[i], RC[12 * 26 + i]);
#pragma unroll 1
    for (int r = 26; r < 30; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        if (r < 29)
            mds_full<true>(s, &RC[12 * (r + 1)]);
        else
            mds_full<false>(s, nullptr);
    }
}

}  // namespace pcs
