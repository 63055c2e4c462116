// Integer-pipe microbenchmark for B200: measured issue rates of the instructions the Goldilocks
// kernels are made of (IMAD.WIDE.U32, IMAD, IADD3, LOP3, SHF), alone and mixed.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o intpipe_bench intpipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 16

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 ^ 0x1234567, a3 = a0 + 99;
    uint32_t b0 = a0 * 7, b1 = a1 * 5, b2 = a2 * 11, b3 = a3 * 13;
    uint64_t w0 = a0, w1 = a1, w2 = a2, w3 = a3, w4 = b0, w5 = b1, w6 = b2, w7 = b3;
    uint32_t m = seed | 1;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int u = 0; u < UNROLL / 8; u++) {
            if (MODE == 0) {  // IMAD.WIDE.U32 x8 independent chains
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w0) : "r"(a0), "r"(m));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w1) : "r"(a1), "r"(m));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2) : "r"(a2), "r"(m));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w3) : "r"(a3), "r"(m));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w4) : "r"(b0), "r"(m));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w5) : "r"(b1), "r"(m));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w6) : "r"(b2), "r"(m));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w7) : "r"(b3), "r"(m));
            } else if (MODE == 1) {  // IMAD (32-bit) x8
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(m), "r"(b0));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(m), "r"(b1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a2) : "r"(m), "r"(b2));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(m), "r"(b3));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b0) : "r"(m), "r"(a0));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b1) : "r"(m), "r"(a1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b2) : "r"(m), "r"(a2));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b3) : "r"(m), "r"(a3));
            } else if (MODE == 2) {  // IADD3 x8 (add.cc / addc pairs = 64-bit adds)
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a0), "+r"(a1) : "r"(b0), "r"(b1));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a2), "+r"(a3) : "r"(b2), "r"(b3));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(b0), "+r"(b1) : "r"(a2), "r"(a3));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(b2), "+r"(b3) : "r"(a0), "r"(a1));
            } else if (MODE == 3) {  // LOP3 x8
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(m), "r"(b0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a1) : "r"(m), "r"(b1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(m), "r"(b2));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a3) : "r"(m), "r"(b3));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b0) : "r"(m), "r"(a0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b1) : "r"(m), "r"(a1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b2) : "r"(m), "r"(a2));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b3) : "r"(m), "r"(a3));
            } else if (MODE == 4) {  // 4 IMAD.WIDE + 4 IADD3-class (mixed, independent)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w0) : "r"(a0), "r"(m));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(b0), "+r"(b1) : "r"(a2), "r"(a3));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w1) : "r"(a1), "r"(m));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2) : "r"(a2), "r"(m));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(b2), "+r"(b3) : "r"(a0), "r"(a1));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w3) : "r"(a3), "r"(m));
            } else if (MODE == 5) {  // 2 IMAD.WIDE + 6 ALU
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w0) : "r"(a0), "r"(m));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(b0), "+r"(b1) : "r"(a2), "r"(a3));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(m), "r"(b0));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w1) : "r"(a1), "r"(m));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(b2), "+r"(b3) : "r"(a0), "r"(a1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a1) : "r"(m), "r"(b1));
            } else if (MODE == 6) {  // SHF (funnel shift) x8
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a0) : "r"(b0));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a1) : "r"(b1));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a2) : "r"(b2));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a3) : "r"(b3));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(b0) : "r"(a0));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(b1) : "r"(a1));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(b2) : "r"(a2));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(b3) : "r"(a3));
            } else if (MODE == 7) {  // IMAD.WIDE with carry-out (mad.lo.cc + madc.hi) style 64-bit MAC chain
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a0), "+r"(a1) : "r"(b0), "r"(m));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a2), "+r"(a3) : "r"(b1), "r"(m));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(b2), "+r"(b3) : "r"(b0), "r"(m));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(b0), "+r"(b1) : "r"(a1), "r"(m));
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3 ^ (uint32_t)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) ^
                                                 (uint32_t)((w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) >> 32);
}

template <int MODE>
void run(const char* name, double ops_per_iter, uint32_t* d_out, int sms) {
    int blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d_out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d_out, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 256 * ITERS * (UNROLL / 8) * ops_per_iter;
    printf("{\"mode\": \"%s\", \"ms\": %.4f, \"thread_ops_per_s\": %.4e, \"ops_per_clk_per_sm_at_1965MHz\": %.2f}\n", name, ms,
           ops / (ms * 1e-3), ops / (ms * 1e-3) / sms / 1.965e9);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    uint32_t* d_out;
    cudaMalloc(&d_out, (size_t)sms * 8 * 256 * 4);
    run<0>("imad_wide_u32", 8, d_out, sms);
    run<1>("imad_lo_u32", 8, d_out, sms);
    run<2>("iadd3_pairs(8 instr)", 8, d_out, sms);
    run<3>("lop3", 8, d_out, sms);
    run<6>("shf", 8, d_out, sms);
    run<4>("mix_4wide_4alu(8 instr)", 8, d_out, sms);
    run<5>("mix_2wide_6alu(8 instr)", 8, d_out, sms);
    run<7>("madlo_cc_madc_hi(8 ptx instr)", 8, d_out, sms);
    cudaError_t e = cudaDeviceSynchronize();
    printf("{\"status\": \"%s\"}\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
