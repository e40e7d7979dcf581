// halfwarp_bench.cu - does the ALU pipe (16 lanes wide: one LOP3 warp-instruction per two cycles) skip an EMPTY half of a
// warp?  VERDICT round 1 (item 2) asked for this test: if a warp whose upper 16 lanes are inactive occupied the pipe for
// one cycle only, two half-full warps per scheduler would issue at 1 instruction per cycle and a Keccak kernel short of
// streams (k_agg_coefs_il at 8,192 streams per GPU, a lone warp per scheduler) could double its rate by spreading its
// lanes over twice the warps.  Measures LOP3 warp-instructions per cycle and scheduler for 1, 2 and 4 warps per scheduler
// with all 32 lanes, the lower 16 lanes, or the even lanes active.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, uint32_t m0, uint32_t m1, int pattern) {
    const int lane = threadIdx.x & 31;
    const bool on = pattern == 0 ? true : (pattern == 1 ? lane < 16 : (lane & 1) == 0);
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 7 + i + m0;
    if (on) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m0), "r"(m1));
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

__global__ void __launch_bounds__(512) kmix(uint32_t* out, int iters, uint32_t m0, uint32_t m1) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 7 + i + m0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m0), "r"(m1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i + 1]) : "r"(m0), "r"(m1));
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

__global__ void __launch_bounds__(512) kshfl(uint32_t* out, int iters, uint32_t m0, uint32_t m1) {
    uint32_t a[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) a[i] = threadIdx.x * 7 + i + m0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            // 5 LOP3 on a rotating register, then one exchange with the partner lane on another register
#pragma unroll
            for (int j = 0; j < 5; ++j)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[(i + 3 * j + 1) % 18]) : "r"(m0), "r"(m1));
            a[i] = __shfl_xor_sync(0xFFFFFFFFu, a[i], 1);
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(m0), "r"(m1));
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 18; ++i) x ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 148 * 512 * 4);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const char* names[3] = {"32 lanes active", "lanes 0-15 active", "even lanes active"};
    for (int pattern = 0; pattern < 3; ++pattern) {
        for (int threads = 128; threads <= 512; threads *= 2) {
            const int iters = 1 << 16;
            k<<<148, threads>>>(d, 64, 3, 5, pattern);
            cudaDeviceSynchronize();
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a);
            k<<<148, threads>>>(d, iters, 3, 5, pattern);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            const double warp_instr_per_sched = (double)(threads / 32 / 4) * iters * 32;
            const double cycles = ms * 1e-3 * clk * 1e3;
            printf("%-18s %d warp(s) per scheduler: %8.3f ms, %.3f LOP3 warp-instructions per cycle and scheduler (at %d MHz nominal)\n",
                   names[pattern], threads / 128, ms, warp_instr_per_sched / cycles, clk / 1000);
        }
    }
    // a lone warp per scheduler with instructions of TWO pipes interleaved (LOP3 on the ALU pipe, IMAD on the FMA pipe):
    // if a warp could issue every cycle this would reach 1.0; the "one instruction every two cycles per warp" rule says 0.5
    for (int threads = 128; threads <= 512; threads *= 2) {
        const int iters = 1 << 16;
        kmix<<<148, threads>>>(d, 64, 3, 5);
        cudaDeviceSynchronize();
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a);
        kmix<<<148, threads>>>(d, iters, 3, 5);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double warp_instr_per_sched = (double)(threads / 32 / 4) * iters * 32;
        const double cycles = ms * 1e-3 * clk * 1e3;
        printf("LOP3 + IMAD 1:1     %d warp(s) per scheduler: %8.3f ms, %.3f warp-instructions per cycle and scheduler\n",
               threads / 128, ms, warp_instr_per_sched / cycles);
    }
    // the instruction mix of the two-lane Keccak round (k_agg_coefs_il): 90 ALU instructions to 17 SHFL, all independent
    for (int threads = 128; threads <= 512; threads *= 2) {
        const int iters = 1 << 15;
        kshfl<<<148, threads>>>(d, 64, 3, 5);
        cudaDeviceSynchronize();
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a);
        kshfl<<<148, threads>>>(d, iters, 3, 5);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double groups_per_sched = (double)(threads / 32 / 4) * iters;
        const double cycles = ms * 1e-3 * clk * 1e3;
        printf("90 LOP3 + 17 SHFL   %d warp(s) per scheduler: %8.3f ms, %.1f cycles per (90 + 17)-instruction group per scheduler\n",
               threads / 128, ms, cycles / groups_per_sched);
    }
    return 0;
}
