// sampler_device.cuh — one SHAKE256 stream per thread, squeeze fused with lattice_algebra's
// decode2polycoefs (index set + signed bounded coefficients).  Shared by k_sampler (sampler.cu) and
// by k_verify (ring.cu), whose half-warps hash the challenges of their own upcoming signatures.
//
// Per-thread scratch lives in shared-memory COLUMNS: element w of a thread's column is base[w*pitch]
// (pitch = threads sharing the region), so a warp walking its streams in lock-step never conflicts.
#pragma once
#include "lcb_device.cuh"

namespace lcb {

constexpr int RATE_WORDS = 34;    // 136-byte SHAKE256 rate

static __constant__ uint64_t c_keccak_rc[24] = LCB_KECCAK_RC_INIT;

// 32-bit little-endian word of a byte string at an arbitrary byte offset, from aligned loads only
// (never touches an aligned word that holds no valid byte).
__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8;
    const uint32_t w0 = __ldg(w);
    const uint32_t w1 = sh ? __ldg(w + 1) : 0u;
    return __funnelshift_r(w0, w1, sh);
}

// The hash input of one stream: salt (warp-uniform words) || msg (ragged bytes).
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
    // stream word k (bytes 4k .. 4k+3) with SHAKE padding applied: 0x1F at byte `tot`, 0x80 at byte `last`
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

struct DecodeParams {
    int bd, wt, vec_len;
    int idx_bits;              // LOGD + secpar
    int mag_bits;              // btd - 1
    int pad_bits;              // 8*nb - bti - wt*btd
};

struct StreamCols {
    uint32_t* rate;            // [34] column: big-endian stream words of the current rate block
    uint32_t* bmap;            // [8]  column: bitmap of still-unused positions (d = 256)
    uint8_t* idxb;             // [wt] byte column: indices in draw order (coefficients come after ALL indices)
    int pitch;
    const uint32_t* mutab;     // [257] floor((2^32-1)/m)   (block-shared)
    const uint32_t* r16tab;    // [257] 2^16 mod m
};

__device__ __forceinline__ void fill_mod_tables(uint32_t* mutab, uint32_t* r16tab) {
    for (int m = 1 + threadIdx.x; m <= 256; m += blockDim.x) {
        mutab[m] = 0xFFFFFFFFu / (uint32_t)m;
        r16tab[m] = 65536u % (uint32_t)m;
    }
}

// SHAKE256(salt || msg) -> vec_len polynomials; emit(poly, e, index, coefficient) in draw order.
//
// A field of L bits is reduced modulo m (the number of unused positions, or bd) without big integers:
// for m <= 256 as 32-bit pieces c, acc <- ((acc*r16 + c>>16)*r16 + (c&0xFFFF)) mod m with r16 = 2^16 mod m
// (one Barrett step per 32 bits); for larger m a 16-bit Horner recurrence.  Fields whose value is not
// needed (magnitudes when bd = 1, pad bits) are skipped 32 bits at a time.  The permutation has ONE
// call site: the first refill absorbs every input block (xor + permute), later refills squeeze.
template <typename Emit>
__device__ __forceinline__ void sample_stream(const DecodeParams& dp, const InputView& iv, const StreamCols& sc,
                                              Emit&& emit) {
    const int P = sc.pitch;
    const int64_t in_total = iv.total();
    const int64_t in_blocks = in_total / 136 + 1;      // the pad byte always needs room
    const int64_t in_last = in_blocks * 136 - 1;
    int64_t in_blk = 0;
    KeccakState s;
#pragma unroll
    for (int i = 0; i < 25; ++i) { s.lo[i] = 0; s.hi[i] = 0; }
    // Big-endian bit cursor: `cur` is the stream word being consumed (`off` bits of it already used),
    // `nxt` the word after it; a field of n <= 32 bits is one funnel shift.  fetch() is the only place
    // that touches the sponge: the first call absorbs every input block, later ones squeeze.
    int wpos = RATE_WORDS;
    auto fetch = [&]() -> uint32_t {
        if (wpos == RATE_WORDS) {
            do {
                if (in_blk < in_blocks) {
                    for (int w = 0; w < RATE_WORDS; ++w)
                        sc.rate[w * P] = iv.word_at(in_blk * RATE_WORDS + w, in_total, in_last);
#pragma unroll
                    for (int i = 0; i < 17; ++i) {
                        s.lo[i] ^= sc.rate[(2 * i) * P];
                        s.hi[i] ^= sc.rate[(2 * i + 1) * P];
                    }
                    ++in_blk;
                }
                keccak_f1600(s, c_keccak_rc);
            } while (in_blk < in_blocks);
#pragma unroll
            for (int i = 0; i < 17; ++i) {
                sc.rate[(2 * i) * P] = __byte_perm(s.lo[i], 0, 0x0123);
                sc.rate[(2 * i + 1) * P] = __byte_perm(s.hi[i], 0, 0x0123);
            }
            wpos = 0;
        }
        return sc.rate[(wpos++) * P];
    };
    // `nxt` runs one word ahead of `cur`; the refill loop is the single call site of fetch() (it also
    // performs the two initial fetches: off starts at 64).
    uint32_t cur = 0, nxt = 0;
    unsigned off = 64;
    auto get = [&](int n) -> uint32_t {       // next n bits (1 <= n <= 32), most significant first
        while (off >= 32) {
            off -= 32;
            cur = nxt;
            nxt = fetch();
        }
        const uint32_t v = __funnelshift_l(nxt, cur, off) >> (32 - n);
        off += n;
        return v;
    };
    const uint32_t bd_mu = 0xFFFFFFFFu / (uint32_t)dp.bd, bd_r16 = 65536u % (uint32_t)dp.bd;
    const int wt = dp.wt;
    for (int poly = 0; poly < dp.vec_len; ++poly) {
#pragma unroll
        for (int w = 0; w < 8; ++w) sc.bmap[w * P] = 0xFFFFFFFFu;
        for (int f = 0; f <= 2 * wt; ++f) {
            // ---- field description (warp-uniform)
            int width;                 // value bits after the optional sign bit
            uint32_t mod;              // 0: raw value (<= 32 bits);  1: value not needed (skip)
            const bool is_coef = f >= wt && f < 2 * wt;
            if (f == 0) { width = LOGD; mod = 0; }
            else if (f < wt) { width = dp.idx_bits; mod = (uint32_t)(D - f); }
            else if (is_coef) { width = dp.mag_bits; mod = (uint32_t)dp.bd; }
            else { width = dp.pad_bits; mod = 1; }
            const bool small = mod >= 2 && mod <= 256, big = mod > 256;
            const uint32_t mu = is_coef ? bd_mu : (small ? sc.mutab[mod] : 0u);
            const uint32_t r16 = is_coef ? bd_r16 : (small ? sc.r16tab[mod] : 0u);
            // ---- consume it, most significant bits first, through the single call site of the cursor
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
                    word = sc.bmap[selw * P];
                } else {
                    // r-th (0-based) still-unused position in ascending order
                    uint32_t k = r;
                    bool found = false;
                    selw = 0; word = 0;
#pragma unroll
                    for (uint32_t w = 0; w < 8; ++w) {
                        uint32_t cand = sc.bmap[w * P];
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
                sc.bmap[selw * P] = word & ~(1u << pos);
                sc.idxb[f * P] = (uint8_t)(selw * 32 + pos);
            } else if (is_coef) {
                const int e = f - wt;
                const int mag = 1 + (int)r;
                emit(poly, e, (int)sc.idxb[e * P], sign ? mag : -mag);
            }
        }
    }
}

}  // namespace lcb
