/* lcb200 — C ABI of the B200-native batched engine for the lattice-cryptography hot path.
 *
 * The reference (b-g-goodell/lattice-cryptography) is pure Python and has no FFI; its seam is
 * the ten names it imports from `lattice_algebra` (one_time_keys.py:4-5, lm_one_time_sigs.py:3,
 * bklm_one_time_agg_sigs.py:1, adaptor_sigs.py:1) plus its scheme functions.  Each entry point
 * below names the reference code it replaces.  See INTEGRATION.md for the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 (LCB_OK) or a negative lcb_status; nothing throws across the ABI;
 *   - the caller owns every buffer; a data pointer may be HOST or DEVICE memory (detected per
 *     call with cudaPointerGetAttributes; host buffers are staged through engine-owned device
 *     scratch inside the call); scalar/offset arrays follow the same rule;
 *   - one ctx = one GPU + one ordered stream; a ctx is not re-entrant; distinct ctxs are independent;
 *   - calls are asynchronous only when EVERY buffer is device memory; otherwise they return after
 *     the results are in the caller's host buffer.  lcb_synchronize() waits for the stream.
 *
 * Data formats (d = ring degree, a power of two in 32..1024; l = vector length; q prime, q = 1 mod 2d)
 *   coefficient form : int16_t[d], centred residues in [-(q-1)/2, (q-1)/2], natural order
 *   NTT form         : uint16_t[d], residues in [0, q); slot p holds a(psi^(2*bitrev_logd(p)+1)),
 *                      psi = least primitive 2d-th root of unity mod q (the reference's `rou`)
 *   pairs            : int16_t[wt][2] = (index, coefficient) in sampler DRAW ORDER
 *   ragged bytes     : uint8_t blob + int64_t off[n+1]; item i is blob[off[i] .. off[i+1])
 * WIDE contexts (q >= 2^16, up to 2^31): every int16_t / uint16_t array of this header then holds int32_t /
 * uint32_t elements of the same shape, and the int32_t BKLM partial sums become int64_t; pass the wider arrays
 * through the same pointers.  The shipped parameter sets (d = 256, q = 11777 / 39937) run on register-resident
 * half-warp kernels; every other (d, q) runs on one-polynomial-per-warp shared-memory kernels (ring_generic.cu).
 * Caller-owned DEVICE buffers of 16-bit data must be 16-byte aligned (LCB_ERR_CUDA, "misaligned address", otherwise).
 */
#ifndef LCB200_H
#define LCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lcb_ctx lcb_ctx;

typedef enum lcb_status {
    LCB_OK = 0,
    LCB_ERR_INVALID = -1,      /* bad argument / unsupported parameter set            */
    LCB_ERR_CUDA = -2,         /* a CUDA runtime call failed (see lcb_last_error)     */
    LCB_ERR_NO_DEVICE = -3,    /* no sm_100 device; there is NO CPU fallback          */
    LCB_ERR_NO_KEY_CH = -4,    /* an operation needs key_ch and lcb_set_key_ch was not called */
    LCB_ERR_OOM = -5
} lcb_status;

#define LCB_SALT_MAX 32

/* Scheme constants, mirroring the reference's `pp` dict (lm_one_time_sigs.py:36-55,
 * bklm_one_time_agg_sigs.py:27-44, adaptor_sigs.py:37-71).  Salts are NUL-terminated ASCII. */
typedef struct lcb_scheme {
    int32_t sk_bd, sk_wt;
    int32_t ch_bd, ch_wt;
    int32_t ag_bd, ag_wt;
    int32_t wit_bd, wit_wt;
    char sk_salt[LCB_SALT_MAX];   /* 'SK_SALT' (the engine appends LEFT / RIGHT)           */
    char ch_salt[LCB_SALT_MAX];   /* 'CH_SALT'                                             */
    char ag_salt[LCB_SALT_MAX];   /* 'AG_SALT' (the engine appends the decimal index)      */
    char wit_salt[LCB_SALT_MAX];  /* 'WIT_SALT'                                            */
} lcb_scheme;

const char* lcb_strerror(int status);
const char* lcb_last_error(const lcb_ctx* ctx);     /* detail string of the last failure    */
int lcb_version(void);
/* 1 for the checked build (make CHECKED=1: device-side asserts on every hand-rolled bound of a global access - the
 * stand-in for compute-sanitizer's memcheck, which is not available on the GPU pool), 0 for the production build.
 * lcb_checked_selftest launches a kernel whose check fails on purpose: LCB_OK on the production build, LCB_ERR_CUDA
 * ("device-side assert triggered", the context is unusable afterwards) on the checked one - proof that checks are live. */
int lcb_build_is_checked(void);
int lcb_checked_selftest(lcb_ctx* ctx);

/* LatticeParameters(modulus=q, degree=d, length=l) + secpar (lattice_algebra; constructed at
 * lm_one_time_sigs.py:19-21; the reference's container tests use (d, q) = (32, 193), tests/test_one_time_keys.py:12-33).
 * Supported: d a power of two in 32..1024, 3 <= q < 2^31 prime with q % (2d) == 1, 1 <= l <= 64,
 * 1 <= secpar <= 512.  device = CUDA ordinal. */
int lcb_ctx_create(lcb_ctx** out, int device, int secpar, int q, int d, int l);
int lcb_ctx_destroy(lcb_ctx* ctx);
int lcb_ctx_set_stream(lcb_ctx* ctx, void* cuda_stream);   /* run on a caller-owned cudaStream_t  */
int lcb_synchronize(lcb_ctx* ctx);
int lcb_ctx_root_of_unity(const lcb_ctx* ctx);             /* == LatticeParameters.rou            */

/* SchemeParameters.key_ch (one_time_keys.py:259-299): upload the public row, coefficient form
 * int16[l][d]; kept NTT-resident on the device until replaced. */
int lcb_set_key_ch(lcb_ctx* ctx, const int16_t* key_ch_coef);

/* hashlib.shake_256(item).digest(out_len) per item (lattice_algebra binary_digest without salt). */
int lcb_shake256_batch(lcb_ctx* ctx, const uint8_t* in, const int64_t* in_off, int64_t n,
                       uint8_t* out, int64_t out_len);

/* The unseeded keygen path for batches (make_random_seed, lm_one_time_sigs.py:58-61, draws secpar random bits per
 * key from the OS): ONE 32-byte secret from the host's entropy source is expanded on the device into n seed
 * bitstrings, seed i = the first secpar bits (most significant bit of each byte first) of
 * SHAKE256(secret32 || le64(first + i)) as ASCII '0'/'1' - exactly the strings keygen hashes, so seed i reproduces
 * key i.  seeds: uint8[n][secpar] (host or device); feed it to lcb_lm_keygen_batch with seed_off[i] = i * secpar. */
int lcb_expand_seeds(lcb_ctx* ctx, const uint8_t* secret32, int64_t first, int64_t n, uint8_t* seeds);

/* hash2polynomialvector / hash2polynomial (lm_one_time_sigs.py:70-91,142-160;
 * adaptor_sigs.py:86-96): SHAKE256(salt || item_i) -> vec_len polynomials with `wt` distinct
 * non-zero positions and coefficients in +-[1..bd].  Either output may be NULL.
 *   out_dense : int16[n][vec_len][d]    out_pairs : int16[n][vec_len][wt][2] */
int lcb_hash2polyvec_batch(lcb_ctx* ctx, const char* salt, const uint8_t* msgs, const int64_t* msg_off,
                           int64_t n, int bd, int wt, int vec_len, int16_t* out_dense, int16_t* out_pairs);

/* Polynomial.__init__ / get_coef_rep / __mul__ (lattice_algebra): transforms and the negacyclic
 * product, npoly independent polynomials. */
int lcb_ntt_fwd_batch(lcb_ctx* ctx, const int16_t* coef, int64_t npoly, uint16_t* ntt);
int lcb_ntt_inv_batch(lcb_ctx* ctx, const uint16_t* ntt, int64_t npoly, int16_t* coef);
int lcb_poly_mul_batch(lcb_ctx* ctx, const int16_t* a, const int16_t* b, int64_t npoly, int16_t* out);
/* Polynomial.ntt_representation exactly as lattice_algebra stores it (parity level L3): the 2d-point
 * cyclic transform of the zero-padded coefficients, natural order, centred: rep[k] = a(rou^k), int16[npoly][2d]. */
int lcb_ntt_reference_repr_batch(lcb_ctx* ctx, const int16_t* coef, int64_t npoly, int16_t* rep);

/* make_one_key / keygen_core (lm_one_time_sigs.py:64-97,126-138; adaptor_sigs.py:104-137).
 * seeds: ragged ASCII bitstrings.  Any output may be NULL.
 *   sk_coef int16[n][2][l][d], sk_ntt uint16[n][2][l][d], vk_ntt uint16[n][2][d], vk_coef int16[n][2][d] */
int lcb_lm_keygen_batch(lcb_ctx* ctx, const lcb_scheme* sch, const uint8_t* seeds, const int64_t* seed_off,
                        int64_t n, int16_t* sk_coef, uint16_t* sk_ntt, uint16_t* vk_ntt, int16_t* vk_coef);

/* make_signature_challenge (lm_one_time_sigs.py:141-160; adaptor_sigs.py:168-188): chmsg item i is
 * the full hash input after the salt, i.e. str(otvk)+', '+msg.  out_pairs int16[n][ch_wt][2]. */
int lcb_challenge_batch(lcb_ctx* ctx, const lcb_scheme* sch, const uint8_t* chmsg, const int64_t* chmsg_off,
                        int64_t n, int16_t* out_pairs);

/* sign (lm_one_time_sigs.py:163-170) and adaptor presign (adaptor_sigs.py:191-195):
 * sig = sk_left ** c + sk_right, sig int16[n][l][d] coefficient form. */
int lcb_lm_sign_batch(lcb_ctx* ctx, const lcb_scheme* sch, const uint16_t* sk_ntt, const uint8_t* chmsg,
                      const int64_t* chmsg_off, int64_t n, int16_t* sig);

/* verify (lm_one_time_sigs.py:173-191), adaptor preverify / verify (adaptor_sigs.py:198-217,247-266):
 * verdict[i] = max|sig_i| <= bd && max weight <= wt && key_ch*sig == vk_left*c + vk_right (+ st).
 * st_ntt may be NULL (LM verify, preverify) or uint16[n][d] (adaptor verify).
 * The bound is tested on the int16 values AS GIVEN: coefficient-form inputs MUST be centred residues (what every
 * engine entry point emits and what the reference's get_coef_rep() returns).  A non-canonical representative
 * such as x + q is not re-centred first; it fails the bound (verdict 0) whenever |x + q| > bd.
 * At most 2^30 items per call (LCB_ERR_INVALID beyond; the same for the packed and the witness-verify entry points). */
int lcb_lm_verify_batch(lcb_ctx* ctx, const lcb_scheme* sch, const uint16_t* vk_ntt, const uint8_t* chmsg,
                        const int64_t* chmsg_off, const int16_t* sig, const uint16_t* st_ntt, int64_t n,
                        int bd, int wt, uint8_t* verdict);

/* ---- Packed wire format.  The reference has no serialisation: keys print as memory
 * addresses and signatures are live Python objects (one_time_keys.py:197-237), so this format is the
 * engine's own and opt-in.  A polynomial is its 256 values v_i = (x_i + bias) mod 2^16, `bits` bits each,
 * value i at bit offset i*bits, least significant bit first: 32*bits bytes per polynomial, polynomials
 * back to back.  Signatures: x = centred coefficient, bias = vf_bd, bits = ceil(log2(2*vf_bd+1)) (11 at
 * secpar 128, 13 at 256).  NTT-form keys: x = slot value, bias = 0, bits = ceil(log2 q) (14 / 16).
 * Defined for d == 256, q < 2^16 contexts (LCB_ERR_INVALID otherwise).
 * `values` is int16 or uint16 [npoly][d]; packed buffers must be 4-byte aligned (LCB_ERR_INVALID otherwise);
 * lcb_lm_verify_packed_batch reads the shipped widths directly only from 16-byte aligned device buffers and
 * unpacks into scratch first when a caller-owned device buffer is merely 4-byte aligned.
 * in_range (nullable) uint8[npoly]: 1 when every value of the polynomial was representable. */
int lcb_pack_batch(lcb_ctx* ctx, const void* values, int64_t npoly, int bits, int bias, uint8_t* packed,
                   uint8_t* in_range);
int lcb_unpack_batch(lcb_ctx* ctx, const uint8_t* packed, int64_t npoly, int bits, int bias, void* values);

/* lcb_lm_verify_batch (verify, lm_one_time_sigs.py:173-191) on packed inputs: vk_packed uint8[n][2][32*vk_bits]
 * (bias 0), sig_packed uint8[n][l][32*sig_bits] (bias sig_bias).  Same verdicts as unpacking first.  The shipped
 * widths (sig/vk = 11/14 and 13/16 bits) are read by the verify kernel directly; other widths are unpacked
 * into engine scratch first.  Host-resident signatures (here and in lcb_lm_verify_batch) cross PCIe in
 * chunks on a second stream while the kernels of the previous chunk run. */
int lcb_lm_verify_packed_batch(lcb_ctx* ctx, const lcb_scheme* sch, const uint8_t* vk_packed, int vk_bits,
                               const uint8_t* chmsg, const int64_t* chmsg_off, const uint8_t* sig_packed,
                               int sig_bits, int sig_bias, int64_t n, int bd, int wt, uint8_t* verdict);

/* make_agg_coefs (bklm_one_time_agg_sigs.py:78-81): coefficient i = H2P(ag_salt+str(first+i) || agmsg),
 * out_pairs int16[count][ag_wt][2].  ag_wt = ag_bd = 1 (the shipped tables: signed monomials) takes the dedicated
 * shared-message kernels; any other 1 <= ag_wt <= d, ag_bd >= 1 runs the full sampler per coefficient. */
int lcb_bklm_agg_coefs(lcb_ctx* ctx, const lcb_scheme* sch, const uint8_t* agmsg, int64_t agmsg_len,
                       int64_t first, int64_t count, int16_t* out_pairs);

/* aggregate (bklm_one_time_agg_sigs.py:92-96), one shard: partial[l][d] (int32, residues mod q,
 * not centred; ag_pairs int16[count][ag_wt][2], general polynomials when ag_wt > 1 or ag_bd > 1)
 * = sum_i sig_sorted[i] ** ag_i over this shard's `count` signatures whose global
 * sorted positions start at `first`.  ag_pairs from lcb_bklm_agg_coefs (same first/count) or NULL to
 * derive them here from agmsg.  Shards are summed by the caller (NCCL reduce) and finished below. */
int lcb_bklm_aggregate_partial(lcb_ctx* ctx, const lcb_scheme* sch, const int16_t* sig_sorted,
                               const int16_t* ag_pairs, const uint8_t* agmsg, int64_t agmsg_len,
                               int64_t first, int64_t count, int32_t* partial);
int lcb_bklm_aggregate_finish(lcb_ctx* ctx, const int32_t* partial_sum, int16_t* ag_sig);

/* aggregate_verify (bklm_one_time_agg_sigs.py:99-116), one shard of the right-hand side:
 * partial[d] (int32 mod q, NTT form) = sum_i (vk_left_i*c_i + vk_right_i) * ag_i. */
int lcb_bklm_aggverify_partial(lcb_ctx* ctx, const lcb_scheme* sch, const uint16_t* vk_ntt_sorted,
                               const uint8_t* chmsg_sorted, const int64_t* chmsg_off, const int16_t* ag_pairs,
                               const uint8_t* agmsg, int64_t agmsg_len, int64_t first, int64_t count,
                               int32_t* partial);
/* bounds 1 <= n <= avf_bd, 1 <= w <= avf_wt, 1 <= total <= ag_cap, then key_ch*ag_sig == sum. */
int lcb_bklm_aggverify_finish(lcb_ctx* ctx, const int32_t* partial_sum, const int16_t* ag_sig, int64_t total,
                              int ag_cap, int avf_bd, int avf_wt, uint8_t* verdict);

/* adaptor make_one_wit / witgen (adaptor_sigs.py:80-101,140-151):
 *   wit_coef int16[n][l][d], st_ntt uint16[n][d], st_coef int16[n][d]; any may be NULL. */
int lcb_adaptor_witgen_batch(lcb_ctx* ctx, const lcb_scheme* sch, const uint8_t* seeds, const int64_t* seed_off,
                             int64_t n, int16_t* wit_coef, uint16_t* st_ntt, int16_t* st_coef);
/* adapt = presig + wit (adaptor_sigs.py:220-221); extract = sig - presig (:224-226); int16[n][l][d]. */
int lcb_vec_add_batch(lcb_ctx* ctx, const int16_t* a, const int16_t* b, int64_t npoly, int16_t* out);
int lcb_vec_sub_batch(lcb_ctx* ctx, const int16_t* a, const int16_t* b, int64_t npoly, int16_t* out);
/* witness_verify (adaptor_sigs.py:229-237): bounds + key_ch*wit == st. */
int lcb_adaptor_witness_verify_batch(lcb_ctx* ctx, const int16_t* wit_coef, const uint16_t* st_ntt, int64_t n,
                                     int bd, int wt, uint8_t* verdict);

/* ---- Multi-device context: ONE call shards a HOST-resident batch over several GPUs of the box.
 * devices[] lists CUDA ordinals (an ordinal may repeat: two contexts on one GPU).  Independent units (keygen, sign,
 * verify) are split into contiguous ranges by the reference's distribute_tasks rule (lm_one_time_sigs.py:194-215: the
 * first n % ndev shards are one longer), one host thread per device, no data-path communication.  BKLM aggregate /
 * aggregate_verify (bklm_one_time_agg_sigs.py:92-116) shard the SORTED list; the per-device int32 partial sums are
 * gathered on device 0 by peer copies (NVLink where peer access exists), added by a kernel and finished there - the
 * exchange step that lattice_cryptography_b200/distributed.py performs with one NCCL reduce when there is one
 * process per GPU.  Every data pointer of these calls must be HOST memory; device-resident work goes through the
 * per-device contexts (lcb_mctx_ctx).  Arguments are those of the single-device calls. */
typedef struct lcb_mctx lcb_mctx;
int lcb_mctx_create(lcb_mctx** out, const int* devices, int ndev, int secpar, int q, int d, int l);
int lcb_mctx_destroy(lcb_mctx* m);
int lcb_mctx_ndev(const lcb_mctx* m);
lcb_ctx* lcb_mctx_ctx(lcb_mctx* m, int i);
const char* lcb_mctx_last_error(const lcb_mctx* m);
int lcb_mctx_set_key_ch(lcb_mctx* m, const int16_t* key_ch_coef);
int lcb_mctx_lm_keygen_batch(lcb_mctx* m, const lcb_scheme* sch, const uint8_t* seeds, const int64_t* seed_off, int64_t n,
                             int16_t* sk_coef, uint16_t* sk_ntt, uint16_t* vk_ntt, int16_t* vk_coef);
int lcb_mctx_lm_sign_batch(lcb_mctx* m, const lcb_scheme* sch, const uint16_t* sk_ntt, const uint8_t* chmsg,
                           const int64_t* chmsg_off, int64_t n, int16_t* sig);
int lcb_mctx_lm_verify_batch(lcb_mctx* m, const lcb_scheme* sch, const uint16_t* vk_ntt, const uint8_t* chmsg,
                             const int64_t* chmsg_off, const int16_t* sig, const uint16_t* st_ntt, int64_t n, int bd,
                             int wt, uint8_t* verdict);
int lcb_mctx_lm_verify_packed_batch(lcb_mctx* m, const lcb_scheme* sch, const uint8_t* vk_packed, int vk_bits,
                                    const uint8_t* chmsg, const int64_t* chmsg_off, const uint8_t* sig_packed,
                                    int sig_bits, int sig_bias, int64_t n, int bd, int wt, uint8_t* verdict);
/* whole aggregate / aggregate_verify in one call: coefficient derivation, sharded sums, device-side reduce, finish */
int lcb_mctx_bklm_aggregate(lcb_mctx* m, const lcb_scheme* sch, const int16_t* sig_sorted, const uint8_t* agmsg,
                            int64_t agmsg_len, int64_t n, int16_t* ag_sig);
int lcb_mctx_bklm_aggregate_verify(lcb_mctx* m, const lcb_scheme* sch, const uint16_t* vk_ntt_sorted,
                                   const uint8_t* chmsg_sorted, const int64_t* chmsg_off, const uint8_t* agmsg,
                                   int64_t agmsg_len, int64_t n, const int16_t* ag_sig, int ag_cap, int avf_bd,
                                   int avf_wt, uint8_t* verdict);

/* Instrumentation: kernels launched by this ctx since creation (bench.py's gpu_launches). */
int64_t lcb_launch_count(const lcb_ctx* ctx);
/* Optional per-kernel timing with CUDA events on the ctx stream (what bench.py's roofline reads).
 * kernel names: sampler shake256 ntt_fwd ntt_inv poly_mul matvec sign verify vec_addsub agg_coefs
 * agg_partial agg_finish aggv_partial aggv_finish.  lcb_profile_read synchronises the stream. */
int lcb_profile_enable(lcb_ctx* ctx, int on);
int lcb_profile_reset(lcb_ctx* ctx);
int lcb_profile_read(lcb_ctx* ctx, const char* kernel, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* LCB200_H */
