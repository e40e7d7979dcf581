// sampler.cu — SHAKE256 and the fused squeeze->decode samplers (sm_100a).
//
// Replaces lattice_algebra's binary_digest / decode2indices / decode2coef / decode2polycoefs /
// hash2polynomial / hash2polynomialvector (restated in oracle/lattice_algebra/__init__.py; call
// sites lm_one_time_sigs.py:70-91,142-160, adaptor_sigs.py:86-96, bklm_one_time_agg_sigs.py:81).
//
// One SHAKE256 stream per THREAD: the 1600-bit state lives in 50 registers and the permutation is
// pure LOP3/SHF work at the ALU-pipe issue limit.  The digest is never written to HBM: each
// 136-byte rate block is spilled to a per-thread column of shared memory ([word][thread],
// conflict-free because all threads of a warp consume the stream in lock-step) and read back by a
// big-endian bit cursor that feeds the index / coefficient decoder directly.
#include "engine.h"
#include "sampler_device.cuh"

namespace lcb {

namespace {

constexpr int SBS = 128;          // threads per block = SHAKE streams per block

// Absorb an input into a fresh state; leaves the state PERMUTED, i.e. its first 136 bytes are the
// first squeeze block.  `rate` is this block's [34][SBS] staging area in shared memory.
__device__ __forceinline__ void absorb(KeccakState& s, uint32_t* rate, int tid, const InputView& in) {
#pragma unroll
    for (int i = 0; i < 25; ++i) { s.lo[i] = 0; s.hi[i] = 0; }
    const int64_t tot = in.total();
    const int64_t nblocks = tot / 136 + 1;          // the pad byte always needs room
    const int64_t last = nblocks * 136 - 1;
    for (int64_t blk = 0; blk < nblocks; ++blk) {
        for (int w = 0; w < RATE_WORDS; ++w) rate[w * SBS + tid] = in.word_at(blk * RATE_WORDS + w, tot, last);
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            s.lo[i] ^= rate[(2 * i) * SBS + tid];
            s.hi[i] ^= rate[(2 * i + 1) * SBS + tid];
        }
        keccak_f1600(s, c_keccak_rc);
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SBS) k_shake256(const uint8_t* __restrict__ in, const int64_t* __restrict__ off,
                                                  int64_t n, uint8_t* __restrict__ out, int64_t out_len) {
    __shared__ uint32_t rate[RATE_WORDS * SBS];
    const int tid = threadIdx.x;
    const int64_t inst = (int64_t)blockIdx.x * SBS + tid;
    if (inst >= n) return;   // no block-level sync in this kernel
    InputView iv{nullptr, 0, in + off[inst], off[inst + 1] - off[inst]};
    KeccakState s;
    absorb(s, rate, tid, iv);
    const uint8_t* rb = reinterpret_cast<const uint8_t*>(rate);
    uint8_t* o = out + inst * out_len;
    int64_t produced = 0;
    while (produced < out_len) {
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            rate[(2 * i) * SBS + tid] = s.lo[i];
            rate[(2 * i + 1) * SBS + tid] = s.hi[i];
        }
        for (int k = 0; k < 136 && produced < out_len; ++k, ++produced)
            o[produced] = rb[((k >> 2) * SBS + tid) * 4 + (k & 3)];
        if (produced < out_len) keccak_f1600(s, c_keccak_rc);
    }
}

// ------------------------------------------------------------------------------------------------
// Fused SHAKE256 squeeze + decode2polycoefs (sampler_device.cuh), one stream per thread.
// Shared memory per block: ring [54][P] u32, bmap [8][P] u32, two modulus tables, the piece-weight table.
template <int P>   // streams (threads) per block; compile-time so that column addressing is shifts, not multiplies
__global__ void __launch_bounds__(P, 640 / P) k_sampler(SamplerArgs a) {   // 5 x 128 streams resident per SM
    extern __shared__ uint32_t smem[];
    uint32_t* ring = smem;                       // [RING_WORDS][P]
    uint32_t* bmap = ring + RING_WORDS * P;      // [8][P]   (directly after the ring, see StreamCols)
    uint32_t* mutab = bmap + 8 * P;
    uint32_t* r16tab = mutab + 260;
    uint8_t* wtab = reinterpret_cast<uint8_t*>(r16tab + 260);     // [257][pieces]
    const int pieces = weight_pieces(a.idx_bits, a.mag_bits);

    const int tid = threadIdx.x;
    const int64_t inst_raw = (int64_t)blockIdx.x * P + tid;
    const bool live = inst_raw < a.n;
    const int64_t inst = live ? inst_raw : a.n - 1;
    // the offsets of this stream's input are requested first, so that their latency hides behind the table fill
    const int64_t item = a.paired ? inst >> 1 : inst;
    const bool second = a.paired && (inst & 1);
    const int64_t msg_begin = __ldg(a.off + item), msg_end = __ldg(a.off + item + 1);
    fill_mod_tables(mutab, r16tab, a.wt);
    fill_weight_table(wtab, a.wt, a.bd, pieces);
    __syncthreads();

    const InputView iv{reinterpret_cast<const uint32_t*>(second ? a.salt2 : a.salt), second ? a.salt2_len : a.salt_len,
                       a.msgs + msg_begin, msg_end - msg_begin};
    const DecodeParams dp{a.bd, a.wt, a.vec_len, a.idx_bits, a.mag_bits, a.pad_bits};
    const StreamCols sc{ring + tid, bmap + tid, P, mutab, r16tab, wtab, pieces, a.idx_scratch + inst_raw, a.idx_stride};
    int16_t* dense = a.out_dense ? a.out_dense + inst * a.dense_stride : nullptr;
    uint32_t* pairs = a.out_pairs ? reinterpret_cast<uint32_t*>(a.out_pairs) + inst * a.vec_len * (int64_t)a.wt : nullptr;
    const int wt = a.wt;
    sample_stream(dp, iv, sc, [&](int poly, int e, int idx, int coef) {
        if (!live) return;
        if (dense) dense[(int64_t)poly * D + idx] = (int16_t)coef;
        if (pairs) pairs[(int64_t)poly * wt + e] = (uint32_t)idx | ((uint32_t)(uint16_t)(int16_t)coef << 16);
    });
}

// ------------------------------------------------------------------------------------------------
// BKLM aggregation coefficients (bklm_one_time_agg_sigs.py:60-81), ag_wt = ag_bd = 1: coefficient i is
// +-X^k with k = first digest byte and sign = top bit of the second byte of
// SHAKE256(ag_salt || str(i) || agmsg).  Every stream absorbs the SAME O(N)-byte message behind a
// salt of 8..27 bytes, so the message cannot be pre-absorbed; each thread walks it with aligned
// 32-bit loads (warp-uniform addresses wherever the decimal index has the same number of digits)
// and funnel shifts.  N streams x ~N*124/136 permutations: the O(N^2) Keccak work of the reference.
constexpr int ABS = 64;   // threads per block (finer blocks balance the ~1 s streams over 148 SMs)

__global__ void __launch_bounds__(ABS) k_agg_coefs(SamplerArgs a) {
    __shared__ uint32_t edge[RATE_WORDS * ABS];
    const int64_t inst = (int64_t)blockIdx.x * ABS + threadIdx.x;
    if (inst >= a.n) return;
    // salt' = ag_salt || decimal(index), at most 27 bytes, kept in eight registers as words
    const uint64_t index = (uint64_t)(a.index_first + inst);
    uint32_t sw[8];
    const uint32_t* salt_w = reinterpret_cast<const uint32_t*>(a.salt);
#pragma unroll
    for (int i = 0; i < 8; ++i) sw[i] = i < (SALT_BYTES / 4) ? salt_w[i] : 0;
    int ndig = 1;
    for (uint64_t v = index; v >= 10; v /= 10) ++ndig;
    const int slen = a.salt_len + ndig;
    {
        uint64_t v = index;
        for (int dpos = slen - 1; dpos >= a.salt_len; --dpos, v /= 10) {
            const uint32_t ch = (uint32_t)('0' + (v % 10));
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if ((dpos >> 2) == i) sw[i] |= ch << (8 * (dpos & 3));
        }
    }
    const int64_t tot = (int64_t)slen + a.shared_len;
    const int64_t nblocks = tot / 136 + 1;
    const int64_t last = nblocks * 136 - 1;
    const uint8_t* msg = a.msgs;
    auto salt_byte = [&](int p) -> uint32_t {
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if ((p >> 2) == i) w = sw[i];
        return (w >> (8 * (p & 3))) & 0xFFu;
    };
    auto word_at = [&](int64_t k) -> uint32_t {
        const int64_t p = 4 * k;
        if (p >= slen && p + 4 <= tot) return load_u32_unaligned(msg + (p - slen));
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t q = p + b;
            uint32_t byte = q < slen ? salt_byte((int)q) : (q < tot ? (uint32_t)__ldg(msg + (q - slen)) : (q == tot ? 0x1Fu : 0u));
            if (q == last) byte ^= 0x80u;
            v |= byte << (8 * b);
        }
        return v;
    };
    KeccakState s;
#pragma unroll
    for (int i = 0; i < 25; ++i) { s.lo[i] = 0; s.hi[i] = 0; }
    for (int64_t blk = 0; blk < nblocks; ++blk) {
        const int64_t k0 = blk * RATE_WORDS;
        // interior blocks: every word is a plain unaligned message word
        const bool interior = 4 * k0 >= slen && 4 * (k0 + RATE_WORDS) <= tot;
        if (interior) {
            const uint8_t* p = msg + (4 * k0 - slen);
            const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
            const uint32_t* w = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
            const unsigned sh = (unsigned)(addr & 3) * 8;
            uint32_t prev = __ldg(w);
#pragma unroll
            for (int i = 0; i < 17; ++i) {
                const uint32_t m1 = __ldg(w + 2 * i + 1);
                const uint32_t m2 = (i < 16 || sh) ? __ldg(w + 2 * i + 2) : 0u;
                s.lo[i] ^= __funnelshift_r(prev, m1, sh);
                s.hi[i] ^= __funnelshift_r(m1, m2, sh);
                prev = m2;
            }
        } else {
            // first / last blocks (salt, tail, padding): assemble word by word through shared memory
            for (int w2 = 0; w2 < RATE_WORDS; ++w2) edge[w2 * ABS + threadIdx.x] = word_at(k0 + w2);
#pragma unroll
            for (int i = 0; i < 17; ++i) {
                s.lo[i] ^= edge[(2 * i) * ABS + threadIdx.x];
                s.hi[i] ^= edge[(2 * i + 1) * ABS + threadIdx.x];
            }
        }
        keccak_f1600(s, c_keccak_rc);
    }
    const uint32_t k = s.lo[0] & 0xFFu;                 // first digest byte: the 8 index bits
    const int sgn = (s.lo[0] >> 15) & 1u ? 1 : -1;      // top bit of the second byte: the sign bit
    reinterpret_cast<uint32_t*>(a.out_pairs)[inst] = k | ((uint32_t)(uint16_t)(int16_t)sgn << 16);
}

}  // namespace

cudaError_t launch_shake256(const uint8_t* in, const int64_t* off, int64_t n, uint8_t* out, int64_t out_len,
                            cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + SBS - 1) / SBS;
    k_shake256<<<(unsigned)blocks, SBS, 0, st>>>(in, off, n, out, out_len);
    return cudaGetLastError();
}

cudaError_t launch_sampler(const SamplerArgs& a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // Grids under two full waves (5 resident 128-stream blocks per SM, register-limited) run as 64-stream blocks instead:
    // the ALU-bound streams then spread evenly (e.g. 2^16 streams: 6.92 blocks per SM, at most 7) instead
    // of leaving SMs with 3 blocks waiting for SMs with 4.
    const bool narrow = (a.n + SBS - 1) / SBS < (int64_t)num_sms * 10;
    const int threads = narrow ? SBS / 2 : SBS;
    if (!a.idx_scratch || a.idx_stride < (a.n + threads - 1) / threads * threads) return cudaErrorInvalidValue;
    const int pieces = weight_pieces(a.idx_bits, a.mag_bits);
    if (pieces > WT_K) return cudaErrorInvalidValue;
    size_t smem = (size_t)(RING_WORDS + 8) * threads * 4 + 2 * 260 * 4 + (size_t)(257 * pieces + 15) / 16 * 16;
    auto kern = narrow ? k_sampler<SBS / 2> : k_sampler<SBS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t blocks = (a.n + threads - 1) / threads;
    kern<<<(unsigned)blocks, threads, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_agg_coefs(const SamplerArgs& a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    int64_t blocks = (a.n + ABS - 1) / ABS;
    k_agg_coefs<<<(unsigned)blocks, ABS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace lcb
