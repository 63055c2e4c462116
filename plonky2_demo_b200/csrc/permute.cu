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

cudaError_t launch_permute(uint64_t* states, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_permute<<<grid_for(n, HASH_THREADS), HASH_THREADS, 0, st>>>(states, n);
    return cudaGetLastError();
}

}  // namespace pcs
