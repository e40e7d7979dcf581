// pipe_bench.cu — issue-rate micro-benchmark of the integer instructions the ring kernels are made of.
// Reports thread-instructions per clock per SM for IMAD (lo), IMAD.HI, IMAD.WIDE, IADD3, LOP3 and mixes,
// i.e. the denominators behind DESIGN.md section 3 ("which pipe binds").
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, uint32_t m0, uint32_t m1) {
    uint32_t a[8];
    uint64_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 7 + i + m0; w[i] = a[i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = a[i] * m0 + m1;                                   // IMAD
                if (MODE == 1) a[i] = __umulhi(a[i], m0) + m1;                          // IMAD.HI
                if (MODE == 2) w[i] += (uint64_t)(uint32_t)w[i] * m0;                   // IMAD.WIDE
                if (MODE == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m0));   // IADD (compiler's choice of pipe)
                if (MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m0), "r"(m1));
                if (MODE == 5) { a[i] = a[i] * m0 + m1; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[(i + 4) & 7]) : "r"(m0), "r"(m1)); }  // 1 IMAD : 1 LOP3
                if (MODE == 6) { a[i] = a[i] * m0 + m1; a[i] = __umulhi(a[i], m1) ; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[(i + 4) & 7]) : "r"(m0), "r"(m1)); }  // 2 FMA : 1 ALU
                if (MODE == 7) { float f = __uint_as_float(a[i]); f = f * 1.0001f + 0.5f; a[i] = __float_as_uint(f); }   // FFMA
                if (MODE == 8) { float f = __uint_as_float(a[i]); f = f * 1.0001f + 0.5f; a[i] = __float_as_uint(f); a[(i + 4) & 7] = a[(i + 4) & 7] * m0 + m1; } // FFMA + IMAD
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <int MODE>
void run(const char* name, int per_iter, uint32_t* d) {
    int iters = 4096, blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(d, 16, 3, 5);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<blocks, 256>>>(d, iters, 3, 5);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double instr = (double)blocks * 256 * iters * per_iter;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-34s %8.3f ms  %7.2f T instr/s  %6.1f thread-instr/clk/SM (at %d MHz)\n", name, ms, instr / ms / 1e9,
           instr / (ms * 1e-3) / (clk * 1e3) / 148, clk / 1000);
}

int main() {
    uint32_t* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("IMAD (lo)", 32, d);
    run<1>("IMAD.HI", 32, d);
    run<2>("IMAD.WIDE (64-bit acc)", 32, d);
    run<3>("IADD", 32, d);
    run<4>("LOP3", 32, d);
    run<5>("IMAD + LOP3 (1:1)", 64, d);
    run<6>("IMAD + IMAD.HI + LOP3 (2:1)", 96, d);
    run<7>("FFMA", 32, d);
    run<8>("FFMA + IMAD (1:1)", 64, d);
    return 0;
}
