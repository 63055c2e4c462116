// How much does STATIC interleaving of independent FP64 (MDS network) and integer (S-box) work inside one thread buy over
// phase-by-phase execution?  DEP=1: the network input depends on the S-box output and vice versa (like one permutation);
// DEP=0: two independent streams in the same loop body (like two permutations skewed by half a round).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "poseidon.cuh"
#ifndef DEP
#define DEP 1
#endif
#ifndef NSBOX
#define NSBOX 2
#endif
#ifndef MINBLOCKS
#define MINBLOCKS 5
#endif
__global__ void __launch_bounds__(128, MINBLOCKS) k(uint64_t* io, size_t n, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dl[12];
    uint64_t x[NSBOX];
#pragma unroll
    for (int j = 0; j < 12; j++) dl[j] = (double)(io[(size_t)j * n + i] & 0xFFFFFFFFu);
#pragma unroll
    for (int j = 0; j < NSBOX; j++) x[j] = io[(size_t)(12 + j) * n + i];
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int j = 0; j < NSBOX; j++) x[j] = pcs::sbox7(x[j]);
#if DEP
        dl[0] = pcs::u32_as_double((uint32_t)x[0]);
#endif
        double y[12];
        pcs::mds_net_d<false>(dl, y);
#pragma unroll
        for (int j = 0; j < 12; j++) dl[j] = y[j] * 0.00390625;
#if DEP
        x[0] += (uint64_t)__double2loint(dl[0] + pcs::TWO52);
#endif
    }
    uint64_t acc = 0;
#pragma unroll
    for (int j = 0; j < NSBOX; j++) acc ^= x[j];
    io[i] = acc + (uint64_t)dl[3];
}
int main() {
    size_t n = (size_t)148 * 128 * MINBLOCKS * 4;
    int reps = 2000;
    uint64_t* d; cudaMalloc(&d, n * 24 * 8); cudaMemset(d, 1, n * 24 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<(unsigned)(n / 128), 128>>>(d, n, reps); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<(unsigned)(n / 128), 128>>>(d, n, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    double fp = 78 + 12, in = 72.0 * NSBOX;
    double iters = (double)n * reps;
    printf("{\"dep\": %d, \"nsbox\": %d, \"regs\": %d, \"ms\": %.3f, \"sm_clk_per_iter_per_128thr\": %.1f, \"issue_slots_used_pct\": %.1f}\n", DEP, NSBOX,
           fa.numRegs, ms, ms * 1e-3 * 1.965e9 * 148 * 128 / iters / 128, 100.0 * (fp + in + 6) * iters / (ms * 1e-3 * 1.965e9 * 148 * 128));
    return 0;
}
