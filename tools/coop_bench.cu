// Latency of ONE chain of cooperative (half-warp) Poseidon permutations against one chain of per-thread permutations.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I plonky2_demo_b200/csrc -o tools/coop_bench tools/coop_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "poseidon_coop.cuh"
__global__ void k_coop(uint64_t* io, int reps) {
    __shared__ uint64_t rc[360];
    pcs::coop_load_rc(rc);
    const unsigned i = threadIdx.x & 15;
    uint64_t s = io[threadIdx.x];
    for (int r = 0; r < reps; r++) s = pcs::poseidon12_coop(s, i, rc);
    io[threadIdx.x] = s;
}
__global__ void k_single(uint64_t* io, int reps) {
    uint64_t s[12];
    for (int k = 0; k < 12; k++) s[k] = io[k * 32 + threadIdx.x];
    for (int r = 0; r < reps; r++) pcs::poseidon12<false>(s);
    for (int k = 0; k < 4; k++) io[k * 32 + threadIdx.x] = s[k];
}
int main() {
    uint64_t* d; cudaMalloc(&d, 4096); cudaMemset(d, 1, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int reps = 170;
    for (int it = 0; it < 2; it++) { cudaEventRecord(e0); k_coop<<<1, 32>>>(d, reps); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"coop_us_per_permutation\": %.2f", ms * 1e3 / reps);
    for (int it = 0; it < 2; it++) { cudaEventRecord(e0); k_single<<<1, 32>>>(d, reps); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf(", \"single_thread_us_per_permutation\": %.2f, \"status\": \"%s\"}\n", ms * 1e3 / reps, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
