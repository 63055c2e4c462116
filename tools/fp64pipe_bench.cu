// FP64 pipe microbenchmark for B200: DFMA/DADD issue rate alone and co-issued with integer work.
// The Poseidon MDS layer is an add/shift network on exact integers < 2^44, which doubles represent
// exactly -- so it can run on the otherwise idle FP64 pipe.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, double seed, uint32_t iseed) {
    double a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 + 7, a3 = a0 + 99;
    double b0 = a0 * 7, b1 = a1 * 5, b2 = a2 * 11, b3 = a3 * 13;
    uint32_t i0 = threadIdx.x + iseed, i1 = i0 * 3, i2 = i0 ^ 0x55, i3 = i0 + 9, j0 = i0 * 7, j1 = i1 * 5, j2 = i2 * 11, j3 = i3 * 13;
    uint64_t w0 = i0, w1 = i1, w2 = i2, w3 = i3;
    const double c = 0.99999, d = 1e-9;
    for (int i = 0; i < ITERS; i++) {
        if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4 || MODE == 6 || MODE == 8) {  // 8 independent DFMA
            a0 = fma(a0, c, d); a1 = fma(a1, c, d); a2 = fma(a2, c, d); a3 = fma(a3, c, d);
            b0 = fma(b0, c, d); b1 = fma(b1, c, d); b2 = fma(b2, c, d); b3 = fma(b3, c, d);
        }
        if (MODE == 1) {  // 8 DADD
            a0 += d; a1 += d; a2 += d; a3 += d; b0 += d; b1 += d; b2 += d; b3 += d;
        }
        if (MODE == 2) {  // + 8 LOP3 (ALU pipe)
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i0) : "r"(iseed), "r"(j0));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i1) : "r"(iseed), "r"(j1));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i2) : "r"(iseed), "r"(j2));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i3) : "r"(iseed), "r"(j3));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(j0) : "r"(iseed), "r"(i0));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(j1) : "r"(iseed), "r"(i1));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(j2) : "r"(iseed), "r"(i2));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(j3) : "r"(iseed), "r"(i3));
        }
        if (MODE == 3) {  // + 8 LOP3 + 8 IMAD  (all three pipes)
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i0) : "r"(iseed), "r"(j0));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i1) : "r"(iseed), "r"(j1));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i2) : "r"(iseed), "r"(j2));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i3) : "r"(iseed), "r"(j3));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(j0) : "r"(iseed), "r"(i0));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(j1) : "r"(iseed), "r"(i1));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(j2) : "r"(iseed), "r"(i2));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(j3) : "r"(iseed), "r"(i3));
        }
        if (MODE == 4 || MODE == 5) {  // 4 IMAD.WIDE.U32 (64-bit accumulate), with (4) or without (5) the DFMAs
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w0) : "r"(i0), "r"(iseed));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w1) : "r"(i1), "r"(iseed));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2) : "r"(i2), "r"(iseed));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w3) : "r"(i3), "r"(iseed));
        }
        if (MODE == 6 || MODE == 7) {  // 8 IADD3 (32-bit adds, no carry), with (6) or without (7) the DFMAs
            asm volatile("add.u32 %0, %0, %1;" : "+r"(i0) : "r"(j0));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(i1) : "r"(j1));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(i2) : "r"(j2));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(i3) : "r"(j3));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(j0) : "r"(i1));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(j1) : "r"(i2));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(j2) : "r"(i3));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(j3) : "r"(i0));
        }
        if (MODE == 8 || MODE == 9) {  // 8 x (add.cc + addc) = 16 carry-chained adds, with (8) or without (9) the DFMAs
            asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(i0), "+r"(j0) : "r"(i1), "r"(j1));
            asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(i1), "+r"(j1) : "r"(i2), "r"(j2));
            asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(i2), "+r"(j2) : "r"(i3), "r"(j3));
            asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(i3), "+r"(j3) : "r"(i0), "r"(j0));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (double)(w0 ^ w1 ^ w2 ^ w3) + a0 + a1 + a2 + a3 + b0 + b1 + b2 + b3 + (double)(i0 ^ i1 ^ i2 ^ i3 ^ j0 ^ j1 ^ j2 ^ j3);
}

template <int MODE>
void run(const char* name, double fp64_ops, double int_ops, double* d_out, int sms) {
    int blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d_out, 1.0, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d_out, 2.0, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)blocks * 256 * ITERS;
    printf("{\"mode\": \"%s\", \"ms\": %.4f, \"fp64_per_clk_per_sm\": %.2f, \"int_per_clk_per_sm\": %.2f, \"total_per_clk_per_sm\": %.2f}\n", name, ms,
           n * fp64_ops / (ms * 1e-3) / sms / 1.965e9, n * int_ops / (ms * 1e-3) / sms / 1.965e9, n * (fp64_ops + int_ops) / (ms * 1e-3) / sms / 1.965e9);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* d_out;
    cudaMalloc(&d_out, (size_t)sms * 8 * 256 * 8);
    run<0>("dfma_x8", 8, 0, d_out, sms);
    run<1>("dadd_x8", 8, 0, d_out, sms);
    run<2>("dfma_x8+lop3_x8", 8, 8, d_out, sms);
    run<3>("dfma_x8+lop3_x4+imad_x4", 8, 8, d_out, sms);
    run<5>("imadwide_x4", 0, 4, d_out, sms);
    run<4>("dfma_x8+imadwide_x4", 8, 4, d_out, sms);
    run<7>("iadd_x8", 0, 8, d_out, sms);
    run<6>("dfma_x8+iadd_x8", 8, 8, d_out, sms);
    run<9>("addcc_addc_x4 (8 instr)", 0, 8, d_out, sms);
    run<8>("dfma_x8+addcc_addc_x4", 8, 8, d_out, sms);
    printf("{\"status\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
