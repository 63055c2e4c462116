// FP64-pipe rate of the MDS network in isolation: mds_net_d on (lo, hi) halves, outputs rescaled by 2^-9 (exact) so that
// the loop can run forever.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I plonky2_demo_b200/csrc -I include -o tools/mdsnet_bench tools/mdsnet_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "poseidon.cuh"

#ifndef MINBLOCKS
#define MINBLOCKS 5
#endif
__global__ void __launch_bounds__(128, MINBLOCKS) k(double* io, size_t n, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dl[12], dh[12];
#pragma unroll
    for (int j = 0; j < 12; j++) { dl[j] = io[(size_t)j * n + i]; dh[j] = io[(size_t)(12 + j) * n + i]; }
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
        double yl[12], yh[12];
        pcs::mds_net_d<false>(dl, yl);
        pcs::mds_net_d<false>(dh, yh);
#pragma unroll
        for (int j = 0; j < 12; j++) { dl[j] = yl[j] * 0.001953125; dh[j] = yh[j] * 0.001953125; }
    }
#pragma unroll
    for (int j = 0; j < 12; j++) { io[(size_t)j * n + i] = dl[j]; io[(size_t)(12 + j) * n + i] = dh[j]; }
}
int main() {
    size_t n = (size_t)148 * 128 * MINBLOCKS * 4;
    int reps = 2000;
    double* d; cudaMalloc(&d, n * 24 * 8); cudaMemset(d, 0, n * 24 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<(unsigned)(n / 128), 128>>>(d, n, reps); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<(unsigned)(n / 128), 128>>>(d, n, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    double ops = (double)n * reps * (2 * 78 + 24);
    printf("{\"regs\": %d, \"ms\": %.3f, \"fp64_instr_per_clk_per_sm\": %.1f, \"status\": \"%s\"}\n", fa.numRegs, ms,
           ops / (ms * 1e-3 * 1.965e9 * 148), cudaGetErrorString(cudaGetLastError()));
    return 0;
}
