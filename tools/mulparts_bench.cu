// Where do the 121 cycles of an S-box go?  Separate rates of (a) the 64x64->128 product, (b) reduce128, (c) both.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "gl64.cuh"
#ifndef MODE
#define MODE 0
#endif
constexpr int NL = 8;
__global__ void __launch_bounds__(128, 5) k(uint64_t* io, size_t n, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x[NL], y[NL];
#pragma unroll
    for (int j = 0; j < NL; j++) { x[j] = io[(size_t)j * n + i]; y[j] = io[(size_t)(NL + j) * n + i]; }
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int j = 0; j < NL; j++) {
#if MODE == 0   // product only: (hi, lo) of x*y, folded back with one xor each
            unsigned __int128 q = (unsigned __int128)x[j] * y[j];
            x[j] = (uint64_t)q ^ 0x9E3779B97F4A7C15ULL;
            y[j] = (uint64_t)(q >> 64) | 1;
#elif MODE == 1  // reduction only
            x[j] = gl::reduce128(x[j], y[j]);
            y[j] += 0x9E3779B97F4A7C15ULL;
#else            // full modmul
            x[j] = gl::mul(x[j], y[j]);
#endif
        }
    }
    uint64_t acc = 0;
#pragma unroll
    for (int j = 0; j < NL; j++) acc ^= x[j] ^ y[j];
    io[i] = acc;
}
int main() {
    size_t n = (size_t)148 * 128 * 5 * 4;
    int reps = 2000;
    uint64_t* d; cudaMalloc(&d, n * 2 * NL * 8); cudaMemset(d, 3, n * 2 * NL * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<(unsigned)(n / 128), 128>>>(d, n, reps); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<(unsigned)(n / 128), 128>>>(d, n, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)n * reps * NL;
    printf("{\"mode\": %d, \"ms\": %.3f, \"cycles_per_op_per_scheduler_warp\": %.2f}\n", MODE, ms, ms * 1e-3 * 1.965e9 * 148 * 4 / (ops / 32));
    return 0;
}
