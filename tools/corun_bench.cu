// Do integer S-box warps and FP64 network warps co-run on one SM?  MODE 0: all warps S-box; 1: all warps network;
// 2: even warps S-box, odd warps network (same per-warp work as in 0 / 1).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "poseidon.cuh"
#ifndef MODE
#define MODE 2
#endif
__global__ void __launch_bounds__(128, 5) k(uint64_t* io, size_t n, int reps_s, int reps_n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool sbox_warp = MODE == 0 || (MODE == 2 && ((threadIdx.x >> 5) & 1) == 0);
    if (sbox_warp) {
        uint64_t x[12];
#pragma unroll
        for (int j = 0; j < 12; j++) x[j] = io[(size_t)j * n + i];
#pragma unroll 1
        for (int r = 0; r < reps_s; r++) {
#pragma unroll
            for (int j = 0; j < 12; j++) x[j] = pcs::sbox7(x[j]);
        }
        uint64_t acc = 0;
#pragma unroll
        for (int j = 0; j < 12; j++) acc ^= x[j];
        io[i] = acc;
    } else {
        double dl[12], dh[12];
#pragma unroll
        for (int j = 0; j < 12; j++) { dl[j] = (double)(io[(size_t)j * n + i] & 0xFFFFFFFFu); dh[j] = dl[j] + 1.0; }
#pragma unroll 1
        for (int r = 0; r < reps_n; r++) {
            double yl[12], yh[12];
            pcs::mds_net_d<false>(dl, yl);
            pcs::mds_net_d<false>(dh, yh);
#pragma unroll
            for (int j = 0; j < 12; j++) { dl[j] = yl[j] * 0.00390625; dh[j] = yh[j] * 0.00390625; }
        }
        io[i] = (uint64_t)(dl[3] + dh[5]);
    }
}
int main() {
    size_t n = (size_t)148 * 128 * 5 * 4;
    // per-warp work chosen so that both kinds of warp take about the same time alone: 12 S-boxes = 1452 cycles, 2 networks + rescale = 180 x 2
    int reps_s = 200, reps_n = 800;
    uint64_t* d; cudaMalloc(&d, n * 12 * 8); cudaMemset(d, 1, n * 12 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<(unsigned)(n / 128), 128>>>(d, n, reps_s, reps_n); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<(unsigned)(n / 128), 128>>>(d, n, reps_s, reps_n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"mode\": %d, \"ms\": %.3f, \"status\": \"%s\"}\n", MODE, ms, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
