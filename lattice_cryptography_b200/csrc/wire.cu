// wire.cu — packed wire format for keys and signatures (sm_100a).
//
// The reference has no serialisation at all: keys print as memory addresses and signatures are Python
// objects (one_time_keys.py:197-237).  This is the engine's own, opt-in format: every polynomial is its
// 256 values v_i = (x_i + bias) mod 2^16, each stored in `bits` bits, value i at bit offset i*bits,
// least significant bit first (byte = offset >> 3, bit = offset & 7) - 32*bits bytes per polynomial.
//   signatures : x = centred coefficient, bias = vf_bd, bits = ceil(log2(2*vf_bd + 1))   (11 / 13)
//   NTT-form keys : x = slot value in [0, q), bias = 0, bits = ceil(log2 q)               (14 / 16)
// Both kernels are pure HBM streams: one warp per polynomial, global traffic as 16-byte / 4-byte
// coalesced accesses, the bit shuffling in a per-warp shared-memory row.
#include "engine.h"

namespace lcb {

namespace {

constexpr int WBS = 256;                       // threads per block
constexpr int WARPS = WBS / 32;
constexpr int ROW_WORDS = 8 * 16 + 4;          // packed polynomial at 16 bits + slack read by the funnel

__global__ void __launch_bounds__(WBS) k_pack(const uint16_t* __restrict__ in, int64_t npoly, int bits, uint32_t bias,
                                              uint32_t* __restrict__ out, uint8_t* __restrict__ in_range) {
    __shared__ uint32_t rows[WARPS][ROW_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* row = rows[warp];
    uint8_t* rowb = reinterpret_cast<uint8_t*>(row);
    const uint32_t limit = 1u << bits;
    const int words = 8 * bits;
    for (int64_t p = (int64_t)blockIdx.x * WARPS + warp; p < npoly; p += (int64_t)gridDim.x * WARPS) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(in + p * D) + lane);     // 8 values of this lane
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        // 8 values x bits = `bits` whole bytes per lane, at byte lane*bits of the row
        uint64_t acc = 0;
        int fill = 0, nb = 0;
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t v = (((w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu) + bias) & 0xFFFFu;
            ok = ok && v < limit;
            acc |= (uint64_t)(v & (limit - 1)) << fill;
            fill += bits;
            while (fill >= 8) {
                rowb[lane * bits + nb++] = (uint8_t)acc;
                acc >>= 8;
                fill -= 8;
            }
        }
        ok = __all_sync(0xFFFFFFFFu, ok);
        __syncwarp();
        uint32_t* o = out + p * words;
        for (int k = lane; k < words; k += 32) o[k] = row[k];
        if (in_range && lane == 0) in_range[p] = ok ? 1 : 0;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(WBS) k_unpack(const uint32_t* __restrict__ in, int64_t npoly, int bits, uint32_t bias,
                                                uint16_t* __restrict__ out) {
    __shared__ uint32_t rows[WARPS][ROW_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* row = rows[warp];
    const uint32_t mask = (1u << bits) - 1;
    const int words = 8 * bits;
    for (int k = lane; k < ROW_WORDS; k += 32) row[k] = 0;
    for (int64_t p = (int64_t)blockIdx.x * WARPS + warp; p < npoly; p += (int64_t)gridDim.x * WARPS) {
        __syncwarp();
        const uint32_t* src = in + p * words;
        for (int k = lane; k < words; k += 32) row[k] = __ldg(src + k);
        __syncwarp();
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int pos = (lane * 8 + i) * bits;
            const uint32_t v = __funnelshift_r(row[pos >> 5], row[(pos >> 5) + 1], pos & 31) & mask;
            const uint32_t x = (v - bias) & 0xFFFFu;
            if (i & 1) w[i >> 1] |= x << 16;
            else w[i >> 1] = x;
        }
        reinterpret_cast<uint4*>(out + p * D)[lane] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

int wire_grid(int64_t npoly, int num_sms) {
    const int64_t blocks = (npoly + WARPS - 1) / WARPS;
    const int64_t cap = (int64_t)num_sms * 8;          // 8 resident 256-thread blocks per SM, grid-stride beyond
    return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

cudaError_t launch_pack(const RingCtx& c, const void* in, int64_t npoly, int bits, int bias, uint8_t* out,
                        uint8_t* in_range, cudaStream_t st) {
    if (npoly <= 0) return cudaSuccess;
    if (bits < 1 || bits > 16) return cudaErrorInvalidValue;
    k_pack<<<wire_grid(npoly, c.num_sms), WBS, 0, st>>>(static_cast<const uint16_t*>(in), npoly, bits,
                                                         (uint32_t)bias & 0xFFFFu, reinterpret_cast<uint32_t*>(out),
                                                         in_range);
    return cudaGetLastError();
}

cudaError_t launch_unpack(const RingCtx& c, const uint8_t* in, int64_t npoly, int bits, int bias, void* out,
                          cudaStream_t st) {
    if (npoly <= 0) return cudaSuccess;
    if (bits < 1 || bits > 16) return cudaErrorInvalidValue;
    k_unpack<<<wire_grid(npoly, c.num_sms), WBS, 0, st>>>(reinterpret_cast<const uint32_t*>(in), npoly, bits,
                                                           (uint32_t)bias & 0xFFFFu, static_cast<uint16_t*>(out));
    return cudaGetLastError();
}

}  // namespace lcb
