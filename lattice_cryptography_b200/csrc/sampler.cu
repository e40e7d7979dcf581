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

namespace lcb {

namespace {

constexpr int SBS = 128;          // threads per block = SHAKE streams per block
constexpr int RATE_WORDS = 34;    // 136 bytes

__constant__ uint64_t c_rc[24] = LCB_KECCAK_RC_INIT;

// ---- message access ------------------------------------------------------------------------------
// 32-bit little-endian word of a byte string at an arbitrary byte offset, from aligned loads only.
// The word must lie inside the string except for its alignment slack (never crosses the aligned
// word that holds the last valid byte).
__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8;
    const uint32_t w0 = __ldg(w);
    const uint32_t w1 = sh ? __ldg(w + 1) : 0u;
    return __funnelshift_r(w0, w1, sh);
}

// The hash input of one stream: salt (warp-uniform, from the kernel parameter) || msg (ragged).
struct InputView {
    const uint32_t* salt_w;     // salt bytes, zero padded, as words
    int salt_len;
    const uint8_t* msg;
    int64_t msg_len;
    __device__ __forceinline__ int64_t total() const { return (int64_t)salt_len + msg_len; }
    __device__ __forceinline__ uint32_t byte_at(int64_t p) const {
        if (p < salt_len) return (salt_w[p >> 2] >> (8 * (p & 3))) & 0xFFu;
        return __ldg(msg + (p - salt_len));
    }
    // stream word k (bytes 4k .. 4k+3) with SHAKE padding applied: pad_pos = total(), last = index of
    // the final byte of the final block
    __device__ __forceinline__ uint32_t word_at(int64_t k, int64_t tot, int64_t last) const {
        const int64_t p = 4 * k;
        if (p + 4 <= salt_len) return salt_w[k];
        if (p >= salt_len && p + 4 <= tot) return load_u32_unaligned(msg + (p - salt_len));
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t q = p + b;
            uint32_t byte = q < tot ? byte_at(q) : (q == tot ? 0x1Fu : 0u);
            if (q == last) byte ^= 0x80u;
            v |= byte << (8 * b);
        }
        return v;
    }
};

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
        if (produced < out_len) keccak_f1600(s, c_rc);
    }
}

// ------------------------------------------------------------------------------------------------
// Fused SHAKE256 squeeze + decode2polycoefs.  Shared memory per block:
//   rate  [34][SBS] u32   big-endian stream words of the current rate block
//   bmap  [8][SBS]  u32   bitmap of still-unused positions (d = 256)
//   idxb  [wt][SBS] u8    indices in draw order (coefficients are drawn after ALL indices)
//
// A field of L bits is reduced modulo m (the number of unused positions, or bd) without big
// integers: for m <= 256 the field is cut into 16-bit pieces h_j and sum_j h_j * (2^(16j) mod m)
// < 2^28 is reduced once; for larger m a 16-bit Horner recurrence with one Barrett step per piece.
__global__ void __launch_bounds__(SBS) k_sampler(SamplerArgs a) {
    extern __shared__ uint32_t smem[];
    uint32_t* rate = smem;
    uint32_t* bmap = rate + RATE_WORDS * SBS;
    uint32_t* mutab = bmap + 8 * SBS;                  // [257] floor((2^32-1)/m)
    uint32_t* r16tab = mutab + 260;                    // [257] 2^16 mod m
    uint8_t* idxb = reinterpret_cast<uint8_t*>(r16tab + 260);

    const int tid = threadIdx.x;
    const int64_t inst_raw = (int64_t)blockIdx.x * SBS + tid;
    const bool live = inst_raw < a.n;
    const int64_t inst = live ? inst_raw : a.n - 1;
    for (int mth = 1 + tid; mth <= 256; mth += SBS) {
        mutab[mth] = 0xFFFFFFFFu / (uint32_t)mth;
        r16tab[mth] = 65536u % (uint32_t)mth;
    }
    __syncthreads();

    const InputView iv{reinterpret_cast<const uint32_t*>(a.salt), a.salt_len, a.msgs + a.off[inst],
                       a.off[inst + 1] - a.off[inst]};
    const int64_t in_total = iv.total();
    const int64_t in_blocks = in_total / 136 + 1;      // the pad byte always needs room
    const int64_t in_last = in_blocks * 136 - 1;
    int64_t in_blk = 0;
    KeccakState s;
#pragma unroll
    for (int i = 0; i < 25; ++i) { s.lo[i] = 0; s.hi[i] = 0; }

    // Big-endian bit cursor over the squeeze stream.  The ONLY call site of the permutation is the
    // refill below: the first refill absorbs every input block (xor + permute), later ones squeeze.
    uint64_t buf = 0;
    int nbits = 0;
    int wpos = RATE_WORDS;
    auto get = [&](int n) -> uint32_t {       // next n bits (1 <= n <= 32), most significant first
        if (nbits < n) {
            if (wpos == RATE_WORDS) {
                do {
                    if (in_blk < in_blocks) {
                        for (int w = 0; w < RATE_WORDS; ++w)
                            rate[w * SBS + tid] = iv.word_at(in_blk * RATE_WORDS + w, in_total, in_last);
#pragma unroll
                        for (int i = 0; i < 17; ++i) {
                            s.lo[i] ^= rate[(2 * i) * SBS + tid];
                            s.hi[i] ^= rate[(2 * i + 1) * SBS + tid];
                        }
                        ++in_blk;
                    }
                    keccak_f1600(s, c_rc);
                } while (in_blk < in_blocks);
#pragma unroll
                for (int i = 0; i < 17; ++i) {
                    rate[(2 * i) * SBS + tid] = __byte_perm(s.lo[i], 0, 0x0123);
                    rate[(2 * i + 1) * SBS + tid] = __byte_perm(s.hi[i], 0, 0x0123);
                }
                wpos = 0;
            }
            buf = (buf << 32) | rate[wpos * SBS + tid];
            ++wpos;
            nbits += 32;
        }
        nbits -= n;
        return (uint32_t)((buf >> nbits) & ((1ull << n) - 1ull));
    };
    const uint32_t bd_mu = 0xFFFFFFFFu / (uint32_t)a.bd, bd_r16 = 65536u % (uint32_t)a.bd;

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
            int width;                 // value bits after the optional sign bit
            uint32_t mod;              // 0: raw value (<= 32 bits);  1: value not needed (skip)
            const bool is_coef = f >= wt && f < 2 * wt;
            if (f == 0) { width = LOGD; mod = 0; }
            else if (f < wt) { width = a.idx_bits; mod = (uint32_t)(D - f); }
            else if (is_coef) { width = a.mag_bits; mod = (uint32_t)a.bd; }
            else { width = a.pad_bits; mod = 1; }
            const bool small = mod >= 2 && mod <= 256, big = mod > 256;
            const uint32_t mu = is_coef ? bd_mu : (small ? mutab[mod] : 0u);
            const uint32_t r16 = is_coef ? bd_r16 : (small ? r16tab[mod] : 0u);   // 2^16 mod m
            // ---- consume it, most significant bits first, through ONE call site of the bit cursor.
            // small m: 32-bit pieces c, acc <- ((acc*r16 + c>>16)*r16 + (c&0xFFFF)) mod m  (< 2^25 before
            // the reduction); large m: 16-bit Horner pieces; the leading piece absorbs the odd bits.
            uint32_t sign = 0, r = 0;
            bool want_sign = is_coef;
            int rem = width + (is_coef ? 1 : 0);
            while (rem > 0) {
                const int mask = big ? 15 : 31;
                const int take = want_sign ? 1 : ((rem & mask) ? (rem & mask) : mask + 1);
                const uint32_t c = get(take);
                rem -= take;
                if (want_sign) { sign = c; want_sign = false; }
                else if (mod == 0) r = c;
                else if (mod != 1) {
                    const uint32_t x = small ? (r * r16 + (c >> 16)) * r16 + (c & 0xFFFFu) : ((r << take) | c);
                    uint32_t t = x - __umulhi(x, mu) * mod;
                    t = t >= mod ? t - mod : t;
                    r = t >= mod ? t - mod : t;
                }
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
        keccak_f1600(s, c_rc);
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
    size_t smem = (size_t)(RATE_WORDS + 8) * SBS * 4 + 2 * 260 * 4 + (size_t)a.wt * SBS;
    cudaError_t e = cudaFuncSetAttribute(k_sampler, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t blocks = (a.n + SBS - 1) / SBS;
    k_sampler<<<(unsigned)blocks, SBS, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_agg_coefs(const SamplerArgs& a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    int64_t blocks = (a.n + ABS - 1) / ABS;
    k_agg_coefs<<<(unsigned)blocks, ABS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace lcb
