// keccak_bench2.cu — can Keccak-f[1600] go faster than the ALU-pipe limit of the LOP3/SHF form by
// moving some of the 64-bit rotations onto the (otherwise idle) FMA-heavy pipe?
//   rotl64((lo,hi), n), 0<n<32:   P = lo * 2^n (IMAD.WIDE.U32);  ohi = hi * 2^n + P.hi (IMAD);
//                                 olo = hi32(hi * 2^n) + P.lo (IMAD.HI.U32)
// (the sums are carry-free: the addends occupy disjoint bits).  Also folds theta's D into the
// application (a ^ C[x-1] ^ rot(C[x+1]) is one LOP3), which saves 10 LOP3 per round.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 keccak_bench2.cu -o keccak_bench2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__constant__ uint64_t c_rc[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
__constant__ uint32_t c_pw[32];      // 2^n, filled at run time so that ptxas cannot turn the multiplies into shifts

__host__ __device__ constexpr int keccak_pi(int i) { return (i / 5) + 5 * ((2 * (i % 5) + 3 * (i / 5)) % 5); }

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t chi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
template <int N>
__device__ __forceinline__ void rot_shf(uint32_t lo, uint32_t hi, uint32_t& olo, uint32_t& ohi) {
    if (N == 0) { olo = lo; ohi = hi; }
    else if (N < 32) { ohi = __funnelshift_l(lo, hi, N); olo = __funnelshift_l(hi, lo, N); }
    else if (N == 32) { olo = hi; ohi = lo; }
    else { ohi = __funnelshift_l(hi, lo, N - 32); olo = __funnelshift_l(lo, hi, N - 32); }
}
template <int N>
__device__ __forceinline__ void rot_imad(uint32_t lo, uint32_t hi, uint32_t& olo, uint32_t& ohi) {
    static_assert(N % 32 != 0, "");
    const uint32_t pw = c_pw[N & 31];
    uint32_t a = N < 32 ? lo : hi, b = N < 32 ? hi : lo, plo, phi;
    asm("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(plo), "=r"(phi) : "r"(a), "r"(pw));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ohi) : "r"(b), "r"(pw), "r"(phi));
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(olo) : "r"(b), "r"(pw), "r"(plo));
}

__device__ constexpr int RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

// MASK bit i set: lane i's rho rotation on the FMA pipe.  TH: how many of the five theta rotations.
template <int I, uint32_t MASK>
struct RhoPi {
    __device__ static __forceinline__ void run(const uint32_t (&lo)[25], const uint32_t (&hi)[25], const uint32_t (&clo)[5],
                                               const uint32_t (&chi_)[5], const uint32_t (&rlo)[5], const uint32_t (&rhi)[5],
                                               uint32_t (&blo)[25], uint32_t (&bhi)[25]) {
        constexpr int x = I % 5;
        uint32_t tl = xor3(lo[I], clo[(x + 4) % 5], rlo[(x + 1) % 5]);
        uint32_t th = xor3(hi[I], chi_[(x + 4) % 5], rhi[(x + 1) % 5]);
        if constexpr (((MASK >> I) & 1) && RHO[I] % 32 != 0) rot_imad<RHO[I]>(tl, th, blo[keccak_pi(I)], bhi[keccak_pi(I)]);
        else rot_shf<RHO[I]>(tl, th, blo[keccak_pi(I)], bhi[keccak_pi(I)]);
        RhoPi<I + 1, MASK>::run(lo, hi, clo, chi_, rlo, rhi, blo, bhi);
    }
};
template <uint32_t MASK>
struct RhoPi<25, MASK> {
    __device__ static __forceinline__ void run(const uint32_t (&)[25], const uint32_t (&)[25], const uint32_t (&)[5],
                                               const uint32_t (&)[5], const uint32_t (&)[5], const uint32_t (&)[5],
                                               uint32_t (&)[25], uint32_t (&)[25]) {}
};

template <uint32_t MASK, int TH, int UNROLL>
__device__ __forceinline__ void keccak32(uint32_t (&lo)[25], uint32_t (&hi)[25]) {
#pragma unroll UNROLL
    for (int round = 0; round < 24; ++round) {
        uint32_t clo[5], chi_[5], rlo[5], rhi[5];
#pragma unroll
        for (int x = 0; x < 5; ++x) {
            clo[x] = xor3(xor3(lo[x], lo[x + 5], lo[x + 10]), lo[x + 15], lo[x + 20]);
            chi_[x] = xor3(xor3(hi[x], hi[x + 5], hi[x + 10]), hi[x + 15], hi[x + 20]);
        }
#pragma unroll
        for (int x = 0; x < 5; ++x) {
            if (x < TH) rot_imad<1>(clo[x], chi_[x], rlo[x], rhi[x]);
            else rot_shf<1>(clo[x], chi_[x], rlo[x], rhi[x]);
        }
        uint32_t blo[25], bhi[25];
        RhoPi<0, MASK>::run(lo, hi, clo, chi_, rlo, rhi, blo, bhi);
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

template <uint32_t MASK, int TH, int UNROLL>
__global__ void __launch_bounds__(128) k_bench(uint64_t* out, int iters) {
    uint32_t lo[25], hi[25];
    for (int i = 0; i < 25; ++i) {
        uint64_t s = (uint64_t)(threadIdx.x + blockIdx.x * blockDim.x) * 0x9E3779B97F4A7C15ull + i;
        lo[i] = (uint32_t)s; hi[i] = (uint32_t)(s >> 32);
    }
    for (int it = 0; it < iters; ++it) keccak32<MASK, TH, UNROLL>(lo, hi);
    uint64_t x = 0;
    for (int i = 0; i < 25; ++i) x ^= ((uint64_t)hi[i] << 32) | lo[i];
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
}

template <uint32_t MASK, int TH, int UNROLL>
void run(const char* name, uint64_t* d_out, uint64_t* ref) {
    const int blocks = 148 * 8, iters = 200;
    k_bench<MASK, TH, UNROLL><<<blocks, 128>>>(d_out, 2);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_bench<MASK, TH, UNROLL><<<blocks, 128>>>(d_out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    uint64_t h; cudaMemcpy(&h, d_out + 5, 8, cudaMemcpyDeviceToHost);
    if (*ref == 0) *ref = h;
    printf("%-34s %8.3f ms  %7.3f Gperm/s  %s (%s)\n", name, ms, (double)blocks * 128 * iters / ms / 1e6,
           h == *ref ? "match" : "MISMATCH", cudaGetErrorString(cudaGetLastError()));
}

__global__ void __launch_bounds__(128) k_ref(uint64_t* out, int iters) {
    uint64_t s[25];
    for (int i = 0; i < 25; ++i) s[i] = (uint64_t)(threadIdx.x + blockIdx.x * blockDim.x) * 0x9E3779B97F4A7C15ull + i;
    for (int it = 0; it < iters; ++it)
        for (int round = 0; round < 24; ++round) {
            uint64_t c[5], b[25];
            for (int x = 0; x < 5; ++x) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
            for (int x = 0; x < 5; ++x) {
                uint64_t r = c[(x + 1) % 5], dd = c[(x + 4) % 5] ^ ((r << 1) | (r >> 63));
                for (int y = 0; y < 5; ++y) s[x + 5 * y] ^= dd;
            }
            _Pragma("unroll") for (int i = 0; i < 25; ++i) b[keccak_pi(i)] = RHO[i] ? ((s[i] << RHO[i]) | (s[i] >> (64 - RHO[i]))) : s[i];
            for (int y = 0; y < 5; ++y)
                for (int x = 0; x < 5; ++x) s[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
            s[0] ^= c_rc[round];
        }
    uint64_t x = 0;
    for (int i = 0; i < 25; ++i) x ^= s[i];
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
}

// first K lanes with a non-trivial rotation
constexpr uint32_t first_k(int k) {
    uint32_t m = 0;
    for (int i = 1; i < 25 && k > 0; ++i) { m |= 1u << i; --k; }
    return m;
}

int main() {
    uint32_t pw[32];
    for (int i = 0; i < 32; ++i) pw[i] = 1u << i;
    cudaMemcpyToSymbol(c_pw, pw, sizeof(pw));
    uint64_t* d_out; cudaMalloc(&d_out, 148 * 8 * 128 * 8);
    uint64_t ref = 0;
    k_ref<<<148 * 8, 128>>>(d_out, 200);
    cudaMemcpy(&ref, d_out + 5, 8, cudaMemcpyDeviceToHost);
    run<0, 0, 2>("fold-D, all SHF, unroll 2", d_out, &ref);
    run<0, 0, 4>("fold-D, all SHF, unroll 4", d_out, &ref);
    run<first_k(4), 0, 2>("4 rho on FMA, u2", d_out, &ref);
    run<first_k(8), 0, 2>("8 rho on FMA, u2", d_out, &ref);
    run<first_k(12), 0, 2>("12 rho on FMA, u2", d_out, &ref);
    run<first_k(16), 0, 2>("16 rho on FMA, u2", d_out, &ref);
    run<first_k(20), 0, 2>("20 rho on FMA, u2", d_out, &ref);
    run<first_k(24), 0, 2>("24 rho on FMA, u2", d_out, &ref);
    run<first_k(24), 5, 2>("24 rho + 5 theta on FMA, u2", d_out, &ref);
    run<first_k(12), 0, 4>("12 rho on FMA, u4", d_out, &ref);
    run<first_k(16), 0, 4>("16 rho on FMA, u4", d_out, &ref);
    run<first_k(20), 0, 4>("20 rho on FMA, u4", d_out, &ref);
    run<first_k(16), 2, 2>("16 rho + 2 theta on FMA, u2", d_out, &ref);
    run<first_k(20), 3, 2>("20 rho + 3 theta on FMA, u2", d_out, &ref);
    return 0;
}
