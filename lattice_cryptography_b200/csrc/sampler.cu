// sampler.cu — SHAKE256 and the fused squeeze->decode samplers (sm_100a).
//
// Replaces lattice_algebra's binary_digest / decode2indices / decode2coef / decode2polycoefs /
// hash2polynomial / hash2polynomialvector (restated in oracle/lattice_algebra/__init__.py; call
// sites lm_one_time_sigs.py:70-91,142-160, adaptor_sigs.py:86-96, bklm_one_time_agg_sigs.py:81).
//
// One SHAKE256 stream per THREAD: the 1600-bit state lives in 50 registers, the permutation is
// LOP3/SHF work on the ALU pipe, and the digest is never written to HBM: each 136-byte rate block
// is spilled to a per-thread column of shared memory ([word][thread], conflict-free because all
// threads of a warp consume the stream in lock-step) and read back by a big-endian bit cursor that
// feeds the index / coefficient decoder directly.
#include "engine.h"

namespace lcb {

namespace {

constexpr int SBS = 128;          // threads per block = SHAKE streams per block
constexpr int RATE_WORDS = 34;    // 136 bytes

__constant__ uint64_t c_rc[24] = LCB_KECCAK_RC_INIT;

__device__ __forceinline__ int decimal_digits(uint64_t v) {
    int n = 1;
    while (v >= 10) { v /= 10; ++n; }
    return n;
}

// Absorb salt || [decimal index] || msg into a fresh state; leaves the state PERMUTED, i.e. its
// first 136 bytes are the first squeeze block.
struct InputView {
    const uint8_t* salt; int salt_len;
    uint64_t index; int ndig;        // ndig == 0: no index suffix
    const uint8_t* msg; int64_t msg_len;
    __device__ __forceinline__ int64_t total() const { return (int64_t)salt_len + ndig + msg_len; }
    __device__ __forceinline__ uint8_t at(int64_t p) const {
        if (p < salt_len) return salt[p];
        p -= salt_len;
        if (p < ndig) {
            uint64_t v = index;
            for (int i = ndig - 1 - (int)p; i > 0; --i) v /= 10;
            return (uint8_t)('0' + (v % 10));
        }
        return __ldg(msg + (p - ndig));
    }
};

__device__ __forceinline__ void absorb(uint64_t (&s)[25], uint32_t* rate, int tid, const InputView& in) {
#pragma unroll
    for (int i = 0; i < 25; ++i) s[i] = 0;
    uint8_t* rb = reinterpret_cast<uint8_t*>(rate);
    const int64_t total = in.total();
    int64_t pos = 0;
    bool done = false;
    while (!done) {
#pragma unroll
        for (int w = 0; w < RATE_WORDS; ++w) rate[w * SBS + tid] = 0;
        int k = 0;
        for (; k < 136 && pos < total; ++k, ++pos) rb[((k >> 2) * SBS + tid) * 4 + (k & 3)] = in.at(pos);
        if (k < 136) {
            rb[((k >> 2) * SBS + tid) * 4 + (k & 3)] ^= 0x1F;    // SHAKE domain bits + first pad bit
            rb[((135 >> 2) * SBS + tid) * 4 + (135 & 3)] ^= 0x80;  // last pad bit
            done = true;
        }
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            uint32_t lo = rate[(2 * i) * SBS + tid], hi = rate[(2 * i + 1) * SBS + tid];
            s[i] ^= ((uint64_t)hi << 32) | lo;
        }
        keccak_f1600(s, c_rc);
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SBS) k_shake256(const uint8_t* __restrict__ in, const int64_t* __restrict__ off,
                                                  int64_t n, uint8_t* __restrict__ out, int64_t out_len) {
    __shared__ uint32_t rate[RATE_WORDS * SBS];
    const int tid = threadIdx.x;
    const int64_t inst = (int64_t)blockIdx.x * SBS + tid;
    if (inst >= n) return;   // no block-level sync in this kernel
    InputView iv{nullptr, 0, 0, 0, in + off[inst], off[inst + 1] - off[inst]};
    uint64_t s[25];
    absorb(s, rate, tid, iv);
    const uint8_t* rb = reinterpret_cast<const uint8_t*>(rate);
    uint8_t* o = out + inst * out_len;
    int64_t produced = 0;
    while (produced < out_len) {
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            rate[(2 * i) * SBS + tid] = (uint32_t)s[i];
            rate[(2 * i + 1) * SBS + tid] = (uint32_t)(s[i] >> 32);
        }
        for (int k = 0; k < 136 && produced < out_len; ++k, ++produced)
            o[produced] = rb[((k >> 2) * SBS + tid) * 4 + (k & 3)];
        if (produced < out_len) keccak_f1600(s, c_rc);
    }
}

// ------------------------------------------------------------------------------------------------
// Fused SHAKE256 squeeze + decode2polycoefs.  Shared memory per block:
//   rate  [34][SBS] u32   big-endian stream words of the current rate block
//   bmap  [8][SBS]  u32   bitmap of still-unused positions (d = 256)
//   idxb  [wt][SBS] u8    indices in draw order (coefficients are drawn after ALL indices)
__global__ void __launch_bounds__(SBS) k_sampler(SamplerArgs a) {
    extern __shared__ uint32_t smem[];
    uint32_t* rate = smem;
    uint32_t* bmap = rate + RATE_WORDS * SBS;
    uint8_t* idxb = reinterpret_cast<uint8_t*>(bmap + 8 * SBS);

    const int tid = threadIdx.x;
    const int64_t inst_raw = (int64_t)blockIdx.x * SBS + tid;
    const bool live = inst_raw < a.n;
    const int64_t inst = live ? inst_raw : a.n - 1;

    InputView iv;
    iv.salt = a.salt;
    iv.salt_len = a.salt_len;
    if (a.shared_msg) {
        iv.index = (uint64_t)(a.index_first + inst);
        iv.ndig = decimal_digits(iv.index);
        iv.msg = a.msgs;
        iv.msg_len = a.shared_len;
    } else {
        iv.index = 0;
        iv.ndig = 0;
        iv.msg = a.msgs + a.off[inst];
        iv.msg_len = a.off[inst + 1] - a.off[inst];
    }
    uint64_t s[25];
    absorb(s, rate, tid, iv);

    // big-endian bit cursor over the squeeze stream
    uint64_t buf = 0;
    int nbits = 0;
    int wpos = RATE_WORDS;      // forces a dump of the current (already permuted) state first
    bool first_block = true;
    auto get = [&](int n) -> uint32_t {
        if (nbits < n) {
            if (wpos == RATE_WORDS) {
                if (!first_block) keccak_f1600(s, c_rc);
                first_block = false;
#pragma unroll
                for (int i = 0; i < 17; ++i) {
                    rate[(2 * i) * SBS + tid] = __byte_perm((uint32_t)s[i], 0, 0x0123);
                    rate[(2 * i + 1) * SBS + tid] = __byte_perm((uint32_t)(s[i] >> 32), 0, 0x0123);
                }
                wpos = 0;
            }
            buf = (buf << 32) | rate[wpos * SBS + tid];
            ++wpos;
            nbits += 32;
        }
        nbits -= n;
        return (uint32_t)(buf >> nbits) & ((1u << n) - 1u);   // n <= 24
    };

    const int wt = a.wt;
    for (int poly = 0; poly < a.vec_len; ++poly) {
#pragma unroll
        for (int w = 0; w < 8; ++w) bmap[w * SBS + tid] = 0xFFFFFFFFu;
        int16_t* dense = a.out_dense ? a.out_dense + inst * a.dense_stride + (int64_t)poly * D : nullptr;
        uint32_t* pairs = a.out_pairs
                              ? reinterpret_cast<uint32_t*>(a.out_pairs) + (inst * a.vec_len + poly) * (int64_t)wt
                              : nullptr;
        for (int f = 0; f <= 2 * wt; ++f) {
            // ---- field description (warp-uniform)
            int width;
            uint32_t mod;
            const bool is_coef = f >= wt && f < 2 * wt;
            if (f == 0) { width = LOGD; mod = 0; }
            else if (f < wt) { width = a.idx_bits; mod = (uint32_t)(D - f); }
            else if (is_coef) { width = 1 + a.mag_bits; mod = (uint32_t)a.bd; }
            else { width = a.pad_bits; mod = 0; }
            const int limb = (mod != 0 && mod <= 256) ? 24 : 16;
            const uint32_t mu = mod ? 0xFFFFFFFFu / mod : 0;
            // ---- consume the field, most significant limb first (Horner, reduced mod `mod`)
            uint32_t r = 0, sign = 0;
            int rem = width;
            bool want_sign = is_coef;
            while (rem > 0) {
                int take = want_sign ? 1 : ((rem % limb) ? (rem % limb) : limb);
                uint32_t v = get(take);
                rem -= take;
                if (want_sign) { sign = v; want_sign = false; }
                else if (mod) {
                    uint32_t x = (r << take) | v;
                    uint32_t t = x - __umulhi(x, mu) * mod;
                    t = t >= mod ? t - mod : t;
                    r = t >= mod ? t - mod : t;
                } else r = v;
            }
            // ---- act on it
            if (f < wt) {
                uint32_t selw, word, pos;
                if (f == 0) {
                    selw = r >> 5;
                    pos = r & 31;
                    word = bmap[selw * SBS + tid];
                } else {
                    // r-th (0-based) still-unused position in ascending order
                    uint32_t k = r;
                    bool found = false;
                    selw = 0; word = 0;
#pragma unroll
                    for (uint32_t w = 0; w < 8; ++w) {
                        uint32_t cand = bmap[w * SBS + tid];
                        uint32_t c = __popc(cand);
                        if (!found) {
                            if (k < c) { found = true; selw = w; word = cand; }
                            else k -= c;
                        }
                    }
                    uint32_t wd = word, c;
                    pos = 0;
                    c = __popc(wd & 0xFFFFu); if (k >= c) { k -= c; pos += 16; wd >>= 16; }
                    c = __popc(wd & 0xFFu);   if (k >= c) { k -= c; pos += 8;  wd >>= 8; }
                    c = __popc(wd & 0xFu);    if (k >= c) { k -= c; pos += 4;  wd >>= 4; }
                    c = __popc(wd & 0x3u);    if (k >= c) { k -= c; pos += 2;  wd >>= 2; }
                    c = wd & 1u;              if (k >= c) { pos += 1; }
                }
                bmap[selw * SBS + tid] = word & ~(1u << pos);
                idxb[f * SBS + tid] = (uint8_t)(selw * 32 + pos);
            } else if (is_coef) {
                const int e = f - wt;
                const int idx = idxb[e * SBS + tid];
                const int mag = 1 + (int)r;
                const int coef = sign ? mag : -mag;
                if (live) {
                    if (dense) dense[idx] = (int16_t)coef;
                    if (pairs) pairs[e] = (uint32_t)idx | ((uint32_t)(uint16_t)(int16_t)coef << 16);
                }
            }
        }
    }
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
    size_t smem = (size_t)(RATE_WORDS + 8) * SBS * 4 + (size_t)a.wt * SBS;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_sampler, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    int64_t blocks = (a.n + SBS - 1) / SBS;
    k_sampler<<<(unsigned)blocks, SBS, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_agg_coefs(const SamplerArgs& a, cudaStream_t st) { return launch_sampler(a, st); }

}  // namespace lcb
