// Issue rate of the S-box layer alone (12 independent x^7 chains per thread, like a full round without the MDS layer).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "poseidon.cuh"
#ifndef MINBLOCKS
#define MINBLOCKS 5
#endif
#ifndef NL
#define NL 12
#endif
__global__ void __launch_bounds__(128, MINBLOCKS) k(uint64_t* io, size_t n, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x[NL];
#pragma unroll
    for (int j = 0; j < NL; j++) x[j] = io[(size_t)j * n + i];
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int j = 0; j < NL; j++) x[j] = pcs::sbox7(x[j]);
    }
    uint64_t acc = 0;
#pragma unroll
    for (int j = 0; j < NL; j++) acc ^= x[j];
    io[i] = acc;
}
int main() {
    size_t n = (size_t)148 * 128 * MINBLOCKS * 4;
    int reps = 1000;
    uint64_t* d; cudaMalloc(&d, n * NL * 8); cudaMemset(d, 1, n * NL * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<(unsigned)(n / 128), 128>>>(d, n, reps); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<(unsigned)(n / 128), 128>>>(d, n, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    double sboxes = (double)n * reps * NL;
    double clk = ms * 1e-3 * 1.965e9 * 148;      // SM-cycles
    printf("{\"lanes\": %d, \"regs\": %d, \"blocks\": %d, \"ms\": %.3f, \"sm_clk_per_sbox_per_128thr\": %.3f, \"cycles_per_sbox_per_scheduler_warp\": %.1f}\n", NL, fa.numRegs, MINBLOCKS,
           ms, clk * 128 / sboxes / 128, clk * 4 / (sboxes / 32));
    return 0;
}
