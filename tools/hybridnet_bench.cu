// Can the MDS network of a narrow (21-bit) piece run on the ALU pipe (IADD3 / LEA / SHF, 32-bit, no carries) UNDER the FP64
// network of the wide (43-bit) piece?  MODE 0: FP64 network only; 1: int32 network only; 2: both in the same loop body.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "poseidon.cuh"
#ifndef MODE
#define MODE 2
#endif
#ifndef FORCE_ALU
#define FORCE_ALU 1
#endif
// ALU-pipe-only building blocks: add.cc / sub.cc always become IADD3 (a plain add may become IMAD.IADD on the FMA pipe,
// which DFMA shares), shf.l becomes SHF (a plain shl may become IMAD.SHL)
struct A32 {
    uint32_t v;
    __device__ __forceinline__ A32() {}
    __device__ __forceinline__ A32(uint32_t x) : v(x) {}
};
__device__ __forceinline__ A32 operator+(A32 a, A32 b) { A32 r; asm("add.cc.u32 %0, %1, %2;" : "=r"(r.v) : "r"(a.v), "r"(b.v)); return r; }
__device__ __forceinline__ A32 operator-(A32 a, A32 b) { A32 r; asm("sub.cc.u32 %0, %1, %2;" : "=r"(r.v) : "r"(a.v), "r"(b.v)); return r; }
__device__ __forceinline__ A32 operator<<(A32 a, int k) { A32 r; asm("shf.l.wrap.b32 %0, 0, %1, %2;" : "=r"(r.v) : "r"(a.v), "r"(k)); return r; }
#if FORCE_ALU
typedef A32 I32;
__device__ __forceinline__ I32 ZERO32() { return A32(0u); }
#else
typedef uint32_t I32;
__device__ __forceinline__ I32 ZERO32() { return 0u; }
#endif
#if FORCE_ALU
__device__ __forceinline__ uint32_t RAW(A32 a) { return a.v; }
#else
__device__ __forceinline__ uint32_t RAW(uint32_t a) { return a; }
#endif
__device__ __forceinline__ void mds_net_i32(const I32 (&s)[12], I32 (&y)[12]) {
    I32 A[3], B[3], P[3], Q[3];
    const I32 Z = ZERO32();
#pragma unroll
    for (int j = 0; j < 3; j++) {
        I32 u = s[j] + s[j + 6], v = s[j + 3] + s[j + 9];
        A[j] = u + v; B[j] = u - v; P[j] = s[j] - s[j + 6]; Q[j] = s[j + 3] - s[j + 9];
    }
    I32 t = A[0] + A[1] + A[2];
    I32 Ya[3] = {t + A[2], t + A[0], t + A[1]};
    I32 Yb[3] = {(B[2] << 3) - (B[1] << 1) - B[0], Z - (B[0] << 3) - B[1] - (B[2] << 1), (B[0] << 1) - (B[1] << 3) - B[2]};
    I32 re[3], im[3];
    re[0] = (P[0] << 1) - Q[0] + P[1] - (Q[1] << 4) + P[2] + (Q[2] << 2);
    im[0] = P[0] + (Q[0] << 1) + (P[1] << 4) + Q[1] - (P[2] << 2) + Q[2];
    re[1] = Z - (P[0] << 2) + Q[0] + (P[1] << 1) - Q[1] + P[2] - (Q[2] << 4);
    im[1] = Z - P[0] - (Q[0] << 2) + P[1] + (Q[1] << 1) + (P[2] << 4) + Q[2];
    re[2] = (P[0] << 4) + Q[0] - (P[1] << 2) + Q[1] + (P[2] << 1) - Q[2];
    im[2] = Z - P[0] + (Q[0] << 4) - P[1] - (Q[1] << 2) + P[2] + (Q[2] << 1);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        I32 e1 = (Ya[j] << 4) + Yb[j], e2 = (Ya[j] << 4) - Yb[j];
        y[j] = e1 + re[j]; y[j + 3] = e2 + im[j]; y[j + 6] = e1 - re[j]; y[j + 9] = e2 - im[j];
    }
    y[0] = y[0] + (s[0] << 3);
}
__global__ void __launch_bounds__(128, 5) k(uint64_t* io, size_t n, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d[12];
    I32 b[12];
#pragma unroll
    for (int j = 0; j < 12; j++) { d[j] = (double)(io[(size_t)j * n + i] & 0xFFFFFFFFu); b[j] = I32((uint32_t)io[(size_t)j * n + i] & 0x1FFFFF); }
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
        if (MODE != 1) {
            double y[12];
            pcs::mds_net_d<false>(d, y);
#pragma unroll
            for (int j = 0; j < 12; j++) d[j] = y[j] * 0.00390625;
        }
        if (MODE != 0) {
            I32 y[12];
            mds_net_i32(b, y);
#pragma unroll
            for (int j = 0; j < 12; j++) b[j] = I32(RAW(y[j]) & 0x1FFFFF);
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 12; j++) acc ^= RAW(b[j]);
    io[i] = acc + (uint64_t)d[3];
}
int main() {
    size_t n = (size_t)148 * 128 * 5 * 4;
    int reps = 4000;
    uint64_t* d; cudaMalloc(&d, n * 12 * 8); cudaMemset(d, 1, n * 12 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<(unsigned)(n / 128), 128>>>(d, n, reps); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<(unsigned)(n / 128), 128>>>(d, n, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"mode\": %d, \"ms\": %.3f, \"cycles_per_iter_per_scheduler_warp\": %.1f}\n", MODE, ms, ms * 1e-3 * 1.965e9 * 148 * 4 / ((double)n * reps / 32));
    return 0;
}
