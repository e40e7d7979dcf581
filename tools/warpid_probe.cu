// Which hardware warp slot (%warpid) and hence which scheduler (slot mod 4) do the warps of co-resident 128-thread
// blocks get?  Prints the (block, warp) -> %warpid map of a few SMs for a 592-block persistent-style launch.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 4) probe(unsigned* out, int spin) {
    extern __shared__ unsigned sm[];
    unsigned smid, wid, nwid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    asm volatile("mov.u32 %0, %%nwarpid;" : "=r"(nwid));
    if ((threadIdx.x & 31) == 0) {
        unsigned* o = out + (blockIdx.x * 4 + (threadIdx.x >> 5)) * 4;
        o[0] = smid; o[1] = wid; o[2] = nwid; o[3] = blockIdx.x;
    }
    // keep the block resident until everything is launched
    long long t0 = clock64();
    while (clock64() - t0 < spin) sm[threadIdx.x] += 1;
}
int main() {
    const int grid = 592;
    unsigned* d; cudaMalloc(&d, grid * 16 * sizeof(unsigned));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
    probe<<<grid, 128, 52 * 1024>>>(d, 2000000);
    cudaDeviceSynchronize();
    static unsigned h[592 * 16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (unsigned s = 0; s < 3; ++s) {
        printf("SM %u:", s);
        for (int i = 0; i < grid * 4; ++i) if (h[4 * i] == s) printf(" b%u.w%d->slot%u", h[4 * i + 3], i & 3, h[4 * i + 1]);
        printf("  (nwarpid %u)\n", h[2]);
    }
    int hist[4] = {0, 0, 0, 0}, same = 0;
    for (int b = 0; b < grid; ++b) { bool ok = true; for (int w = 0; w < 4; ++w) { hist[h[(4 * b + w) * 4 + 1] & 3]++; if ((h[(4 * b + w) * 4 + 1] & 3) != (unsigned)w) ok = false; } same += ok; }
    printf("slot&3 histogram %d %d %d %d; blocks whose warp w sits on slot&3 == w: %d of %d\n", hist[0], hist[1], hist[2], hist[3], same, grid);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
