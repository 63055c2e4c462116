// Batched Poseidon-12 permutation (pcs_poseidon_permute): one thread per state.
// Reference: Poseidon::poseidon, plonky2/src/hash/poseidon.rs:599-609.
#include "hash_common.cuh"

namespace pcs {

__global__ void __launch_bounds__(HASH_THREADS) k_permute(uint64_t* __restrict__ states, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s[12];
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(states + i * 12);
#pragma unroll
    for (int k = 0; k < 6; k++) {
        ulonglong2 v = p[k];
        s[2 * k] = gl::canon(v.x);
        s[2 * k + 1] = gl::canon(v.y);
    }
    poseidon12(s);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(states + i * 12);
#pragma unroll
    for (int k = 0; k < 6; k++) o[k] = make_ulonglong2(s[2 * k], s[2 * k + 1]);
}

// FRI proof-of-work search (plonky2/src/fri/prover.rs:115-160): candidate c goes into lane `pos` of the duplex
// intermediate state, the state is permuted and lane 7 (the last rate element, `squeeze().last()`) must have at least
// `min_lz` leading zeros.  Every thread tests one candidate of [base, base + n); the smallest hit wins (what the
// reference finds with one rayon thread).
__global__ void __launch_bounds__(HASH_THREADS) k_pow_search(const uint64_t* __restrict__ state, unsigned pos,
                                                            unsigned min_lz, uint64_t base, uint64_t n,
                                                            unsigned long long* __restrict__ best) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = gl::canon(state[k]);
    const uint64_t cand = base + i;
#pragma unroll
    for (int k = 0; k < 12; k++)
        if ((unsigned)k == pos) s[k] = cand;
    poseidon12(s);
    if ((unsigned)__clzll((long long)s[7]) >= min_lz) atomicMin(best, (unsigned long long)cand);
}

cudaError_t launch_pow_search(const uint64_t* state_dev, unsigned pos, unsigned min_lz, uint64_t base, uint64_t n,
                              unsigned long long* best_dev, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_pow_search<<<grid_for(n, HASH_THREADS), HASH_THREADS, 0, st>>>(state_dev, pos, min_lz, base, n, best_dev);
    return cudaGetLastError();
}

cudaError_t launch_permute(uint64_t* states, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_permute<<<grid_for(n, HASH_THREADS), HASH_THREADS, 0, st>>>(states, n);
    return cudaGetLastError();
}

}  // namespace pcs
