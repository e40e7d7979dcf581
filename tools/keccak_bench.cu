// keccak_bench.cu — micro-benchmark of Keccak-f[1600] variants (one state per thread, registers only).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../lattice_cryptography_b200/csrc keccak_bench.cu -o keccak_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "lcb_device.cuh"
using namespace lcb;

__constant__ uint64_t c_rc[24] = LCB_KECCAK_RC_INIT;

template <int UNROLL>
__device__ __forceinline__ void keccak64(uint64_t (&s)[25]) {
#pragma unroll UNROLL
    for (int round = 0; round < 24; ++round) {
        uint64_t c[5];
#pragma unroll
        for (int x = 0; x < 5; ++x) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
#pragma unroll
        for (int x = 0; x < 5; ++x) {
            uint64_t dd = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma unroll
            for (int y = 0; y < 5; ++y) s[x + 5 * y] ^= dd;
        }
        uint64_t b[25];
#pragma unroll
        for (int i = 0; i < 25; ++i) b[keccak_pi(i)] = KECCAK_RHO[i] ? rotl64(s[i], KECCAK_RHO[i]) : s[i];
#pragma unroll
        for (int y = 0; y < 5; ++y)
#pragma unroll
            for (int x = 0; x < 5; ++x)
                s[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        s[0] ^= c_rc[round];
    }
}

// explicit 32-bit halves: LOP3 (xor3 / chi) + funnel shifts only
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t chi(uint32_t a, uint32_t b, uint32_t c) {   // a ^ (~b & c)
    uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
template <int N>
__device__ __forceinline__ void rot(uint32_t lo, uint32_t hi, uint32_t& olo, uint32_t& ohi) {
    if (N == 0) { olo = lo; ohi = hi; }
    else if (N < 32) { ohi = __funnelshift_l(lo, hi, N); olo = __funnelshift_l(hi, lo, N); }
    else if (N == 32) { olo = hi; ohi = lo; }
    else { ohi = __funnelshift_l(hi, lo, N - 32); olo = __funnelshift_l(lo, hi, N - 32); }
}
template <int I> struct RhoPi {
    __device__ static __forceinline__ void run(const uint32_t (&lo)[25], const uint32_t (&hi)[25], const uint32_t (&dlo)[5],
                                               const uint32_t (&dhi)[5], uint32_t (&blo)[25], uint32_t (&bhi)[25]) {
        constexpr int R[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
        uint32_t tl = lo[I] ^ dlo[I % 5], th = hi[I] ^ dhi[I % 5];
        rot<R[I]>(tl, th, blo[keccak_pi(I)], bhi[keccak_pi(I)]);
        RhoPi<I + 1>::run(lo, hi, dlo, dhi, blo, bhi);
    }
};
template <> struct RhoPi<25> {
    __device__ static __forceinline__ void run(const uint32_t (&)[25], const uint32_t (&)[25], const uint32_t (&)[5],
                                               const uint32_t (&)[5], uint32_t (&)[25], uint32_t (&)[25]) {}
};
template <int UNROLL>
__device__ __forceinline__ void keccak32(uint32_t (&lo)[25], uint32_t (&hi)[25]) {
#pragma unroll UNROLL
    for (int round = 0; round < 24; ++round) {
        uint32_t clo[5], chi_[5], dlo[5], dhi[5];
#pragma unroll
        for (int x = 0; x < 5; ++x) {
            clo[x] = xor3(xor3(lo[x], lo[x + 5], lo[x + 10]), lo[x + 15], lo[x + 20]);
            chi_[x] = xor3(xor3(hi[x], hi[x + 5], hi[x + 10]), hi[x + 15], hi[x + 20]);
        }
#pragma unroll
        for (int x = 0; x < 5; ++x) {
            uint32_t rl, rh;
            rot<1>(clo[(x + 1) % 5], chi_[(x + 1) % 5], rl, rh);
            dlo[x] = clo[(x + 4) % 5] ^ rl;
            dhi[x] = chi_[(x + 4) % 5] ^ rh;
        }
        uint32_t blo[25], bhi[25];
        RhoPi<0>::run(lo, hi, dlo, dhi, blo, bhi);
#pragma unroll
        for (int y = 0; y < 5; ++y)
#pragma unroll
            for (int x = 0; x < 5; ++x) {
                lo[x + 5 * y] = chi(blo[x + 5 * y], blo[(x + 1) % 5 + 5 * y], blo[(x + 2) % 5 + 5 * y]);
                hi[x + 5 * y] = chi(bhi[x + 5 * y], bhi[(x + 1) % 5 + 5 * y], bhi[(x + 2) % 5 + 5 * y]);
            }
        lo[0] ^= (uint32_t)c_rc[round];
        hi[0] ^= (uint32_t)(c_rc[round] >> 32);
    }
}

template <int V, int UNROLL>
__global__ void __launch_bounds__(128) k_bench(uint64_t* out, int iters) {
    uint64_t s[25];
    for (int i = 0; i < 25; ++i) s[i] = (uint64_t)(threadIdx.x + blockIdx.x * blockDim.x) * 0x9E3779B97F4A7C15ull + i;
    if (V == 0) {
        for (int it = 0; it < iters; ++it) keccak64<UNROLL>(s);
    } else {
        uint32_t lo[25], hi[25];
        for (int i = 0; i < 25; ++i) { lo[i] = (uint32_t)s[i]; hi[i] = (uint32_t)(s[i] >> 32); }
        for (int it = 0; it < iters; ++it) keccak32<UNROLL>(lo, hi);
        for (int i = 0; i < 25; ++i) s[i] = ((uint64_t)hi[i] << 32) | lo[i];
    }
    uint64_t x = 0;
    for (int i = 0; i < 25; ++i) x ^= s[i];
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
}

template <int V, int UNROLL>
void run(const char* name, uint64_t* d_out, uint64_t* ref) {
    const int blocks = 148 * 8, iters = 200;
    k_bench<V, UNROLL><<<blocks, 128>>>(d_out, 2);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_bench<V, UNROLL><<<blocks, 128>>>(d_out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    uint64_t h; cudaMemcpy(&h, d_out + 5, 8, cudaMemcpyDeviceToHost);
    if (*ref == 0) *ref = h;
    printf("%-22s %8.3f ms  %7.3f Gperm/s  %s (%s)\n", name, ms, (double)blocks * 128 * iters / ms / 1e6,
           h == *ref ? "match" : "MISMATCH", cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint64_t* d_out; cudaMalloc(&d_out, 148 * 8 * 128 * 8);
    uint64_t ref = 0;
    run<0, 2>("u64 unroll 2", d_out, &ref);
    run<0, 4>("u64 unroll 4", d_out, &ref);
    run<0, 8>("u64 unroll 8", d_out, &ref);
    run<0, 24>("u64 unroll 24", d_out, &ref);
    run<1, 2>("u32 unroll 2", d_out, &ref);
    run<1, 4>("u32 unroll 4", d_out, &ref);
    run<1, 6>("u32 unroll 6", d_out, &ref);
    run<1, 8>("u32 unroll 8", d_out, &ref);
    run<1, 12>("u32 unroll 12", d_out, &ref);
    run<1, 24>("u32 unroll 24", d_out, &ref);
    return 0;
}
