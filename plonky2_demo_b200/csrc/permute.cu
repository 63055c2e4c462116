// Batched Poseidon-12 permutation (pcs_poseidon_permute): one thread per state.
// Reference: Poseidon::poseidon, plonky2/src/hash/poseidon.rs:599-609.
#include "hash_common.cuh"
#include "poseidon_coop.cuh"

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

// Latency form (<= COOP_MAX_PERMS states, e.g. the Challenger's single duplexing): one state per half-warp.
__global__ void __launch_bounds__(128) k_permute_coop(uint64_t* __restrict__ states, size_t n) {
    __shared__ uint64_t rc[360];
    coop_load_rc(rc);
    const size_t k0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / COOP_LANES;
    const unsigned i = threadIdx.x & (COOP_LANES - 1);
    const bool active = k0 < n;
    uint64_t* p = states + (active ? k0 : 0) * 12;
    uint64_t s = i < 12 ? p[i] : 0;
    s = gl::canon(poseidon12_coop(s, i, rc));
    if (active && i < 12) p[i] = s;
}

cudaError_t launch_permute(uint64_t* states, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (n <= COOP_MAX_PERMS)
        k_permute_coop<<<grid_for(n * COOP_LANES, 128), 128, 0, st>>>(states, n);
    else
        k_permute<<<grid_for(n, HASH_THREADS), HASH_THREADS, 0, st>>>(states, n);
    return cudaGetLastError();
}

}  // namespace pcs
