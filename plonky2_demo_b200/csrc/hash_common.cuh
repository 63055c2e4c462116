// Shared pieces of the hashing kernels (digest layout addressing).
#pragma once
#include "common.cuh"
#include "poseidon.cuh"

namespace pcs {

constexpr int HASH_THREADS = 128;

// Slot of node k (global index within its level) in the reference digest layout
// (merkle_tree.rs:43-51; closed form of prove()'s index math :197-201).
__device__ __forceinline__ uint64_t* digest_slot(uint64_t* digests, uint64_t* cap, unsigned lg_sub,
                                                 unsigned level, size_t node) {
    unsigned lg_per = lg_sub - level;  // log2(nodes per subtree at this level)
    size_t s = node >> lg_per;
    if (lg_per == 0) return cap + 4 * s;  // subtree root lives only in the cap
    size_t k = node & (((size_t)1 << lg_per) - 1);
    size_t sub_len = 2 * (((size_t)1 << lg_sub) - 1);
    size_t idx = 2 * (((k >> 1) << (level + 1)) + ((size_t)1 << level) - 1) + (k & 1);
    return digests + (s * sub_len + idx) * 4;
}

static inline unsigned grid_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace pcs
