// Poseidon-12 with ONE permutation spread over a half-warp (16 lanes, 12 of them holding one state element each).
//
// The throughput kernels (poseidon.cuh) run one permutation per thread: ~17 k dependent-ish instructions, 18 us on an
// otherwise idle SM (tools/poseidon_bench.cu with one warp, profiles/r02_small_commits.md).  That is the right shape for 10^8
// permutations, but a Merkle tree is a CHAIN: a 135-element leaf is 17 permutations one after the other and every tree level
// waits for the one below, so small commitments (the m = 2 demo: 64 leaves) and the top levels of every tree are pure
// latency.  Here the 12 S-boxes of a full round run on 12 lanes at once and the MDS layer is 11 shuffles + 12 small-constant
// multiply-adds per lane: ~135 instructions per lane and round instead of ~1250 per thread, ~4x lower latency per permutation
// at ~3x the total instruction count -- used only where the GPU is not full anyway (<= 8192 permutations in flight).
//
// Form: the reference's poseidon_naive (plonky2/src/hash/poseidon.rs:613-633): 30 rounds of { + ALL_ROUND_CONSTANTS[12 r + i],
// x^7 on every lane (full rounds r < 4, r >= 26) or on lane 0 only, MDS = circ(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0,..)
// (poseidon_goldilocks.rs:24-25; mds_row_shf, poseidon.rs:223-240) }.  The reference's `consistency` test (poseidon.rs:777-790)
// asserts naive == fast; tests/ pin this kernel to the reference KATs through pcs_poseidon_permute (n <= 8192).
#pragma once
#include "poseidon.cuh"

namespace pcs {

constexpr int COOP_LANES = 16;                 // lanes per permutation (a half-warp)
constexpr size_t COOP_MAX_PERMS = 8192;        // below this many independent permutations per launch the GPU is latency bound

__device__ __forceinline__ uint64_t coop_shfl(uint64_t v, unsigned src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, (int)src, COOP_LANES);
    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), (int)src, COOP_LANES);
    return gl::pack(lo, hi);
}

// y_i = sum_k CIRC[k] * s[(i + k) % 12] + DIAG[i] * s_i, on 32-bit halves (sums < 2^41), folded back to a loose u64
__device__ __forceinline__ uint64_t coop_mds(uint64_t s, unsigned i) {
    constexpr uint32_t CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    const uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
    uint64_t al0 = (uint64_t)lo * CIRC[0], ah0 = (uint64_t)hi * CIRC[0], al1 = 0, ah1 = 0;
#pragma unroll
    for (int k = 1; k < 12; k++) {
        unsigned src = i + k;
        src -= src >= 12 ? 12 : 0;
        const uint32_t l = __shfl_sync(0xffffffffu, lo, (int)src, COOP_LANES);
        const uint32_t h = __shfl_sync(0xffffffffu, hi, (int)src, COOP_LANES);
        if (k & 1) {
            al1 += (uint64_t)l * CIRC[k];
            ah1 += (uint64_t)h * CIRC[k];
        } else {
            al0 += (uint64_t)l * CIRC[k];
            ah0 += (uint64_t)h * CIRC[k];
        }
    }
    if (i == 0) {   // MDS_MATRIX_DIAG[0] = 8
        al0 += (uint64_t)lo << 3;
        ah0 += (uint64_t)hi << 3;
    }
    return gl::fold_halves(al0 + al1, ah0 + ah1);
}

// s = state element of lane i (lanes 12..15: anything, ignored); rc = ALL_ROUND_CONSTANTS (shared memory copy).
// Every lane of the warp must call this (full-mask shuffles).  Returns the permuted element, loose.
__device__ __forceinline__ uint64_t poseidon12_coop(uint64_t s, unsigned i, const uint64_t* __restrict__ rc) {
    const unsigned ii = i < 12 ? i : 0;
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        s = gl::add_lc(s, rc[12 * r + ii]);
        const bool full = r < 4 || r >= 26;
        if (full || i == 0) s = sbox7(s);
        s = coop_mds(s, i);
    }
    return s;
}

__device__ __forceinline__ void coop_load_rc(uint64_t* rc_sh) {
    for (unsigned e = threadIdx.x; e < 360; e += blockDim.x) rc_sh[e] = pconst::RC[e];
    __syncthreads();
}

}  // namespace pcs
