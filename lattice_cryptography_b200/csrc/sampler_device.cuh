// sampler_device.cuh — one SHAKE256 stream per thread, squeeze fused with lattice_algebra's
// decode2polycoefs (index set + signed bounded coefficients).  Device-side building block of k_sampler
// (sampler.cu); kept in a header so that other kernels can embed a stream (an experiment that hashed the
// challenges inside k_verify was measured 15 % slower than two kernels and dropped).
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
    // One whole 136-byte rate block (stream words 34*blk .. 34*blk+33) into the shared-memory column `col`
    // (element w at col[w * P]).  Because 136 is a multiple of 4, every stream word of every block sits at
    // the same byte phase of the message: 35 aligned, mutually independent loads (batched by the compiler,
    // one round trip instead of 34 dependent ones) and 34 funnel shifts produce the interior words; only the
    // words that touch the salt, the end of the message or the padding are patched through word_at().
    __device__ __forceinline__ void load_block(int64_t blk, int64_t tot, int64_t last, uint32_t* col, int P) const {
        const int64_t k0 = blk * RATE_WORDS;
        const uintptr_t a = reinterpret_cast<uintptr_t>(msg) + (uintptr_t)(4 * k0 - salt_len);   // stream byte 4*k0
        const uint32_t* aw = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const unsigned sh = (unsigned)(a & 3) * 8;
        // an aligned word may be read iff it overlaps [msg, msg + msg_len): word indices [j_lo, j_hi) of aw[]
        const int64_t first = (int64_t)((reinterpret_cast<uintptr_t>(msg) & ~(uintptr_t)3) - reinterpret_cast<uintptr_t>(aw)) >> 2;
        const int64_t end = (int64_t)(reinterpret_cast<uintptr_t>(msg) + (uintptr_t)msg_len + 3 - reinterpret_cast<uintptr_t>(aw)) >> 2;
        const int j_lo = (int)(first < 0 ? 0 : (first > RATE_WORDS + 1 ? RATE_WORDS + 1 : first));
        const int j_hi = msg_len > 0 ? (int)(end < 0 ? 0 : (end > RATE_WORDS + 1 ? RATE_WORDS + 1 : end)) : 0;
#ifdef LCB_CHECKED
        // every aligned word this block may read lies inside [msg & ~3, (msg + msg_len + 3) & ~3)
        if (j_hi > j_lo) {
            const uintptr_t lo_ok = reinterpret_cast<uintptr_t>(msg) & ~(uintptr_t)3;
            const uintptr_t hi_ok = (reinterpret_cast<uintptr_t>(msg) + (uintptr_t)msg_len + 3) & ~(uintptr_t)3;
            LCB_CHECK(reinterpret_cast<uintptr_t>(aw + j_lo) >= lo_ok && reinterpret_cast<uintptr_t>(aw + j_hi) <= hi_ok);
        }
#endif
        uint32_t prev = (0 >= j_lo && 0 < j_hi) ? __ldg(aw) : 0u;
#pragma unroll
        for (int w = 0; w < RATE_WORDS; ++w) {
            const uint32_t next = (w + 1 >= j_lo && w + 1 < j_hi) ? __ldg(aw + w + 1) : 0u;
            col[w * P] = __funnelshift_r(prev, next, sh);
            prev = next;
        }
        const int64_t ks = ((int64_t)salt_len + 3) >> 2;          // stream words [0, ks) touch the salt
        for (int64_t k = k0; k < ks && k < k0 + RATE_WORDS; ++k) col[(k - k0) * P] = word_at(k, tot, last);
        const int64_t kt = tot >> 2;                               // the word that holds the 0x1F marker
        if (kt >= k0 && kt < k0 + RATE_WORDS) col[(kt - k0) * P] = word_at(kt, tot, last);
        for (int64_t k = (kt + 1 > k0 ? kt + 1 : k0); k < k0 + RATE_WORDS; ++k)
            col[(k - k0) * P] = (k == (last >> 2)) ? 0x80000000u : 0u;
    }
};

// Ring geometry seen by the decoder: the shipped degree as compile-time constants (every shift and loop bound
// folds), or any power-of-two degree 32..1024 at run time.
struct Geo256 {
    static constexpr bool kFixed = true;
    __device__ __forceinline__ int d() const { return D; }
    __device__ __forceinline__ int logd() const { return LOGD; }
    __device__ __forceinline__ int words() const { return D / 32; }
};
struct GeoAny {
    static constexpr bool kFixed = false;
    int d_, logd_;
    __device__ __forceinline__ int d() const { return d_; }
    __device__ __forceinline__ int logd() const { return logd_; }
    __device__ __forceinline__ int words() const { return d_ >> 5; }
};

struct DecodeParams {
    int bd, wt, vec_len;
    int idx_bits;              // LOGD + secpar
    int mag_bits;              // btd - 1
    int pad_bits;              // 8*nb - bti - wt*btd
};

constexpr int RING_EXTRA = 20;                       // words a field may need beyond the current block
constexpr int RING_WORDS = RATE_WORDS + RING_EXTRA;  // per-stream window over the squeeze stream
constexpr int MAX_FIELD_BITS = 32 * (RING_EXTRA - 1);

template <typename IdxT>
struct StreamColsT {
    uint32_t* ring;            // [RING_WORDS] column: big-endian stream words, window over the digest
    uint32_t* bmap;            // [d / 32] column: bitmap of still-unused positions; MUST follow `ring`
    int pitch;
    const uint32_t* mutab;     // [257] floor((2^32-1)/m)   (block-shared)
    const uint32_t* r16tab;    // [257] 2^16 mod m
    const uint8_t* wtab;       // [257][wstride] 2^(32k) mod m, rows copied by load_mod_tables
    int wstride;               // pieces per row
    IdxT* idxs;                // global scratch, index e of this stream at idxs[e * idx_stride]
    int64_t idx_stride;
};
using StreamCols = StreamColsT<uint8_t>;

// Only the rows a sampler call can touch: the index moduli D-wt+1 .. D-1 (bd keeps its constants in registers).
__device__ __forceinline__ void fill_mod_tables(uint32_t* mutab, uint32_t* r16tab, int wt, int d = D) {
    for (int m = max(1, d - wt + 1) + (int)threadIdx.x; m <= d; m += blockDim.x) {
        mutab[m] = 0xFFFFFFFFu / (uint32_t)m;
        if (r16tab) r16tab[m] = 65536u % (uint32_t)m;
    }
}

// Weights of the 32-bit pieces of a big-endian field modulo m <= 256: wtab[m][k] = 2^(32k) mod m.
// Only the rows a sampler call can touch are copied into shared memory: the index moduli D-wt+1 .. D-1 and bd.  A row
// holds `pieces` = max(2, ceil(widest field / 32)) <= WT_K bytes there (WT_K in the per-context global table).
constexpr int WT_K = 20;          // MAX_FIELD_BITS / 32 + 1
__host__ __device__ __forceinline__ int weight_pieces(int idx_bits, int mag_bits) {
    const int w = idx_bits > mag_bits ? idx_bits : mag_bits;
    const int p = (w + 31) >> 5;
    return p < 2 ? 2 : p;
}
// The three tables of the fast decoder, copied from the per-context global table (engine.h: mod_tab; made on the host
// by fill_sampler_mod_table, sampler.cu): no division or modulo in the kernel.  (Computed in the kernel, the dependent
// modulo chains of the 21 rows kept warp 0 of every block busy for ~20 us behind 19 ALU-bound warps while its three
// sibling warps sat at the barrier: 12.8 % of the challenge sampler's warp samples for 0.3 % of its instructions.
// Removing them changed the kernel's time by nothing - 1.741 vs 1.743 ms per 2^20 challenges, the other warps kept the
// ALU pipe fed - but a prologue without division routines is the simpler one.)
__device__ __forceinline__ void load_mod_tables(const uint32_t* __restrict__ tab, uint32_t* mutab, uint32_t* r16tab,
                                                uint8_t* wtab, int wt, int bd, int pieces) {
    const int m_first = max(1, D - wt + 1);
    for (int m = m_first + (int)threadIdx.x; m <= D; m += blockDim.x) {
        mutab[m] = __ldg(tab + m);
        r16tab[m] = __ldg(tab + 260 + m);
    }
    const uint8_t* gw = reinterpret_cast<const uint8_t*>(tab + 520);
    const int m_lo = max(2, D - wt + 1), m_hi = D - 1;
    for (int i = m_lo + (int)threadIdx.x; i <= m_hi + 1; i += blockDim.x) {
        const int m = i <= m_hi ? i : bd;
        if (m < 2 || m > 256) continue;
        for (int k = 0; k < pieces; ++k) wtab[m * pieces + k] = __ldg(gw + m * WT_K + k);
    }
}

__device__ __forceinline__ uint32_t barrett_small(uint32_t x, uint32_t mu, uint32_t mod) {
    uint32_t t = x - __umulhi(x, mu) * mod;          // in [0, 3 mod)
    t = min(t, t - mod);                             // unsigned wrap-around: subtracts only when t >= mod
    return min(t, t - mod);
}

// SHAKE256(salt || msg) -> vec_len polynomials; emit(poly, e, index, coefficient) in draw order.
//
// The digest is consumed through a per-stream WINDOW of RING_WORDS stream words in shared memory.
// ensure(bits) - called once at the start of every field and the ONLY call site of the permutation -
// tops the window up by one 136-byte block whenever the field would not fit (the first call absorbs
// every input block).  Inside a field the cursor is a pure funnel shift over two window words, so the
// per-field code is straight-line and specialised:
//   index field   : (8 + secpar) bits reduced mod (#unused positions) as a weighted sum of 32-bit pieces
//                   (weights 2^(32k) mod m from a block-shared table, 64-bit accumulator, one fold), then
//                   the acc-th unused position is taken from a 256-bit bitmap (popcount search);
//   coefficient   : sign bit, then magnitude 1 + (field mod bd) the same way (bd <= 256), or a 16-bit
//                   Horner recurrence (larger bd, only key_ch), or nothing at all when bd == 1;
//   pad bits      : skipped.
// Indices are parked in a global scratch column until their coefficients arrive (the coefficients of a
// polynomial are drawn after ALL its indices).
// Where the digest words come from: the stream's own sponge (the normal case: squeeze fused with the decoder, the
// digest never leaves the SM), or a buffer that ANOTHER warp fills (k_sampler_coop: one producer runs the sponge, one
// consumer warp per polynomial decodes behind it - the low-latency form for a handful of streams).
struct SpongeFeed {
    static constexpr bool kBuffered = false;
};
struct BufferFeed {
    static constexpr bool kBuffered = true;
    const uint32_t* words;               // digest as (even-bit half, odd-bit half) pairs per 64-bit sponge word - the
                                         // producer's registers as they are; this polynomial's first pair at [0]
    const volatile unsigned* produced;   // 32-bit stream words the producer has published so far (shared memory)
    unsigned base;                       // index of words[0] in the producer's stream
    unsigned next;                       // words consumed so far
    static __device__ __forceinline__ uint32_t spread16(uint32_t x) {      // bit i of the low half -> bit 2 i
        x &= 0xFFFFu;
        x = (x | (x << 8)) & 0x00FF00FFu;
        x = (x | (x << 4)) & 0x0F0F0F0Fu;
        x = (x | (x << 2)) & 0x33333333u;
        return (x | (x << 1)) & 0x55555555u;
    }
    __device__ __forceinline__ void refill(uint32_t* col, int P) {
        const unsigned need = base + next + RATE_WORDS;
        while (*produced < need) __nanosleep(100);
        __threadfence_block();
#pragma unroll
        for (int i = 0; i < RATE_WORDS / 2; ++i) {
            // the interleaving the producer skipped: 64-bit word = even bits from one half, odd bits from the other
            const uint2 h = __ldcg(reinterpret_cast<const uint2*>(words + next) + i);
            col[(2 * i) * P] = __byte_perm(spread16(h.x) | (spread16(h.y) << 1), 0, 0x0123);
            col[(2 * i + 1) * P] = __byte_perm(spread16(h.x >> 16) | (spread16(h.y >> 16) << 1), 0, 0x0123);
        }
        next += RATE_WORDS;
    }
};

template <typename Geo, typename IdxT, typename Feed, typename Emit>
__device__ __forceinline__ void sample_stream_t(const Geo& geo, const DecodeParams& dp, const InputView& iv,
                                                const StreamColsT<IdxT>& sc, Feed& feed, Emit&& emit) {
    const int P = sc.pitch;
    const int64_t in_total = iv.total();
    // the pad byte always needs room; 32-bit division whenever the input is shorter than 4 GiB
    const int64_t in_blocks = (in_total >> 32) == 0 ? (int64_t)((uint32_t)in_total / 136u) + 1 : in_total / 136 + 1;
    const int64_t in_last = in_blocks * 136 - 1;
    int64_t in_blk = 0;
    KeccakState s;
#pragma unroll
    for (int i = 0; i < 25; ++i) { s.lo[i] = 0; s.hi[i] = 0; }
    int rp = 0;                 // read cursor, in bits from the start of the window
    int nw = 0;                 // valid words in the window
    auto ensure = [&](int bits) {
        if (rp + bits > 32 * nw) {
            const int drop = rp >> 5, keep = nw - drop;          // keep <= RING_EXTRA
            for (int i = 0; i < keep; ++i) sc.ring[i * P] = sc.ring[(drop + i) * P];
            rp -= 32 * drop;
            if constexpr (Feed::kBuffered) {
                feed.refill(sc.ring + keep * P, P);
            } else {
                do {
                    if (in_blk < in_blocks) {
                        // the not-yet-valid part of the window doubles as staging for the input block
                        iv.load_block(in_blk, in_total, in_last, sc.ring + keep * P, P);
#pragma unroll
                        for (int i = 0; i < 17; ++i) {
                            s.lo[i] ^= sc.ring[(keep + 2 * i) * P];
                            s.hi[i] ^= sc.ring[(keep + 2 * i + 1) * P];
                        }
                        ++in_blk;
                    }
                    keccak_f1600(s, c_keccak_rc);
                } while (in_blk < in_blocks);
#pragma unroll
                for (int i = 0; i < 17; ++i) {
                    sc.ring[(keep + 2 * i) * P] = __byte_perm(s.lo[i], 0, 0x0123);
                    sc.ring[(keep + 2 * i + 1) * P] = __byte_perm(s.hi[i], 0, 0x0123);
                }
            }
            nw = keep + RATE_WORDS;
        }
    };
    auto take = [&](int n) -> uint32_t {      // next n bits (1 <= n <= 32), most significant first; no refill
        const int wi = rp >> 5;
        const uint32_t v = __funnelshift_l(sc.ring[(wi + 1) * P], sc.ring[wi * P], rp & 31) >> (32 - n);
        rp += n;
        return v;
    };
    // Value of the next `width` bits modulo m (2 <= m <= 256).  The field is cut into 32-bit pieces
    // aligned to its END (the most significant piece holds the `head` odd bits); their weighted sum
    // sum_k piece_k * (2^(32k) mod m) < 2^45 is accumulated in 64 bits - independent multiply-adds
    // on the FMA pipe instead of a serial reduce-per-piece chain on the ALU pipe - and folded once.
    auto field_small = [&](int width, uint32_t m, uint32_t mu, uint32_t r16, const uint8_t* wrow) -> uint32_t {
        const int np = (width + 31) >> 5;
        const int head = width - 32 * (np - 1);
        const uint32_t* rw = sc.ring + (rp >> 5) * P;
        uint32_t c = __funnelshift_l(rw[P], rw[0], rp & 31) >> (32 - head);
        uint64_t acc = (uint64_t)c * wrow[np - 1];
        rp += head;
        rw = sc.ring + (rp >> 5) * P;
        const int sh = rp & 31;
        uint32_t prev = rw[0];
        for (int k = np - 2; k >= 0; --k) {
            rw += P;
            const uint32_t nxt = rw[0];
            c = __funnelshift_l(nxt, prev, sh);
            prev = nxt;
            acc += (uint64_t)c * wrow[k];
        }
        rp += 32 * (np - 1);
        const uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);     // hi < 2^13
        return barrett_small(hi * wrow[1] + (lo >> 16) * r16 + (lo & 0xFFFFu), mu, m);
    };
    // Value of the next `width` bits modulo m < 2^16 as a 16-bit Horner recurrence (any width; generic degrees and
    // the bd > 256 sampler of key_ch), and the same for moduli up to 2^31 with 64-bit remainders (wide contexts).
    auto field_horner = [&](int width, uint32_t m, uint32_t mu) -> uint32_t {
        uint32_t r = 0;
        int rem = width;
        while (rem > 0) {
            const int n = (rem & 15) ? (rem & 15) : 16;
            r = barrett_small((r << n) | take(n), mu, m);
            rem -= n;
        }
        return r;
    };
    auto field_horner64 = [&](int width, uint32_t m) -> uint32_t {
        uint64_t r = 0;
        int rem = width;
        while (rem > 0) {
            const int n = (rem & 31) ? (rem & 31) : 32;
            r = ((r << n) | take(n)) % m;
            rem -= n;
        }
        return (uint32_t)r;
    };
    const uint32_t bd = (uint32_t)dp.bd;
    const uint32_t bd_mu = 0xFFFFFFFFu / bd, bd_r16 = 65536u % bd;
    const int wt = dp.wt;
    for (int poly = 0; poly < dp.vec_len; ++poly) {
        if constexpr (Geo::kFixed) {
#pragma unroll
            for (int w = 0; w < 8; ++w) sc.bmap[w * P] = 0xFFFFFFFFu;
        } else {
            for (int w = 0; w < geo.words(); ++w) sc.bmap[w * P] = 0xFFFFFFFFu;
        }
        for (int f = 0; f <= 2 * wt; ++f) {
            const bool is_idx = f < wt, is_coef = f >= wt && f < 2 * wt;
            ensure(f == 0 ? geo.logd() : (is_idx ? dp.idx_bits : (is_coef ? 1 + dp.mag_bits : dp.pad_bits)));   // the one call site
            if (is_idx) {
                // ---- one position of the index set
                uint32_t selw, word, pos;
                if (f == 0) {
                    const uint32_t r = take(geo.logd());
                    selw = r >> 5;
                    pos = r & 31;
                    word = sc.bmap[selw * P];
                } else {
                    const uint32_t m = (uint32_t)(geo.d() - f);
                    uint32_t k = 0;
                    if (m == 1) rp += dp.idx_bits;
                    else if constexpr (Geo::kFixed) k = field_small(dp.idx_bits, m, sc.mutab[m], sc.r16tab[m], sc.wtab + m * sc.wstride);
                    else k = field_horner(dp.idx_bits, m, sc.mutab[m]);
                    // k-th (0-based) still-unused position in ascending order
                    bool found = false;
                    selw = 0; word = 0;
                    if constexpr (Geo::kFixed) {
#pragma unroll
                        for (uint32_t w = 0; w < 8; ++w) {
                            const uint32_t cand = sc.bmap[w * P];
                            const uint32_t cnt = __popc(cand);
                            if (!found) {
                                if (k < cnt) { found = true; selw = w; word = cand; }
                                else k -= cnt;
                            }
                        }
                    } else {
                        for (uint32_t w = 0; w < (uint32_t)geo.words() && !found; ++w) {
                            const uint32_t cand = sc.bmap[w * P];
                            const uint32_t cnt = __popc(cand);
                            if (k < cnt) { found = true; selw = w; word = cand; }
                            else k -= cnt;
                        }
                    }
                    uint32_t wd = word, cnt;
                    pos = 0;
                    cnt = __popc(wd & 0xFFFFu); if (k >= cnt) { k -= cnt; pos += 16; wd >>= 16; }
                    cnt = __popc(wd & 0xFFu);   if (k >= cnt) { k -= cnt; pos += 8;  wd >>= 8; }
                    cnt = __popc(wd & 0xFu);    if (k >= cnt) { k -= cnt; pos += 4;  wd >>= 4; }
                    cnt = __popc(wd & 0x3u);    if (k >= cnt) { k -= cnt; pos += 2;  wd >>= 2; }
                    cnt = wd & 1u;              if (k >= cnt) { pos += 1; }
                }
                sc.bmap[selw * P] = word & ~(1u << pos);
                LCB_CHECK(selw * 32 + pos < (uint32_t)geo.d() && ((word >> pos) & 1u));     // an unused position of the ring
                sc.idxs[f * sc.idx_stride] = (IdxT)(selw * 32 + pos);
            } else if (is_coef) {
                // ---- one coefficient
                const int e = f - wt;
                const uint32_t sign = take(1);
                uint32_t r = 0;
                if (bd == 1) {
                    rp += dp.mag_bits;
                } else if (Geo::kFixed && bd <= 256) {
                    r = field_small(dp.mag_bits, bd, bd_mu, bd_r16, sc.wtab + bd * sc.wstride);
                } else if (bd < 65536u) {
                    r = field_horner(dp.mag_bits, bd, bd_mu);
                } else {
                    r = field_horner64(dp.mag_bits, bd);
                }
                const int mag = 1 + (int)r;
                emit(poly, e, (int)sc.idxs[e * sc.idx_stride], sign ? mag : -mag);
            } else {
                rp += dp.pad_bits;
            }
        }
    }
}

template <typename Emit>
__device__ __forceinline__ void sample_stream(const DecodeParams& dp, const InputView& iv, const StreamCols& sc, Emit&& emit) {
    SpongeFeed feed;
    sample_stream_t<Geo256, uint8_t>(Geo256{}, dp, iv, sc, feed, static_cast<Emit&&>(emit));
}

}  // namespace lcb
