/* cocg -- C ABI of the B200 (sm_100a) kernels behind collaborative-circom's MPC proving path.
 *
 * The reference (pure Rust) has no FFI; its operator seam is the trait set in
 * /root/reference/mpc-core/src/traits.rs consumed by CoGroth16<T,P> / CoPlonk<T,P>.  Every entry point below
 * is what a Rust `impl {MSMProvider, FFTProvider, PrimeFieldMpcProtocol} for {Rep3Protocol, ShamirProtocol,
 * PlainDriver}` would bind through `extern "C"` (see INTEGRATION.md for the binding), and names the trait
 * method / impl (file:line under /root/reference) it replaces.
 *
 * Conventions
 *  - All functions return 0 on success, nonzero on failure; cocg_last_error() gives the message.  Nothing
 *    throws or aborts across the boundary (the reference's compute ops are infallible, traits.rs:535-568).
 *  - One cocg_ctx per driver / thread (the reference's drivers are `&mut self`, traits.rs:43); contexts are
 *    independent and may be used concurrently from different threads.  Work is enqueued on the context's CUDA
 *    stream; entry points taking HOST pointers synchronise before returning, entry points taking DEVICE
 *    pointers are asynchronous until cocg_sync().
 *  - Field elements: 32 bytes (Fr of either curve), little-endian limbs, MONTGOMERY form (R = 2^256) -- the
 *    in-memory layout of arkworks' Fp<MontBackend<_,4>> a Rust caller holds.  Share vectors are plain arrays
 *    of such elements, one array per share component (the reference's SoA Rep3PrimeFieldShareVec{a,b},
 *    mpc-core/src/protocols/rep3/fieldshare.rs:232-236).
 *  - Affine points: x | y (G2: x.c0 | x.c1 | y.c0 | y.c1), each coordinate 32 B (BN254) / 48 B (BLS12-381)
 *    Montgomery limbs; all-zero = point at infinity (zkey layout, circom-types/src/traits.rs:107-155).
 *  - Jacobian points (outputs): X | Y | Z in the same coordinate encoding; Z == 0 is infinity.  Equality is
 *    group equality (X/Z^2, Y/Z^3) as for arkworks' Projective.
 */
#ifndef COCG_H
#define COCG_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define COCG_API __attribute__((visibility("default")))
#else
#define COCG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cocg_ctx cocg_ctx;

enum { COCG_BN254 = 0, COCG_BLS12_381 = 1 };
enum { COCG_G1 = 1, COCG_G2 = 2 };
/* element-wise ops (cocg_vec_op) */
enum { COCG_OP_MUL = 0, COCG_OP_ADD = 1, COCG_OP_SUB = 2, COCG_OP_NEG = 3, COCG_OP_TO_MONT = 4, COCG_OP_FROM_MONT = 5 };

/* ---- lifecycle ------------------------------------------------------------------------------------- */
COCG_API int cocg_create(cocg_ctx** out, int device, int curve);
COCG_API void cocg_destroy(cocg_ctx* ctx);
COCG_API const char* cocg_last_error(cocg_ctx* ctx); /* ctx may be NULL: error of the last failed cocg_create */
COCG_API int cocg_version(void);
/* Use an externally owned cudaStream_t (e.g. the torch current stream); NULL restores the context's own. */
COCG_API int cocg_set_stream(cocg_ctx* ctx, void* cuda_stream);
/* high != 0: the context's own stream gets the device's highest scheduling priority (its pending thread blocks go ahead of those of
 * other streams' running kernels); 0: default priority.  Call while the context is idle. */
COCG_API int cocg_set_stream_priority(cocg_ctx* ctx, int high);
COCG_API int cocg_sync(cocg_ctx* ctx);
/* Number of kernel launches issued by this context since creation (bench.py's `gpu_launches`). */
COCG_API uint64_t cocg_launch_count(cocg_ctx* ctx);

/* Device-side timing per kernel class (CUDA events on the launching stream; bench.py's roofline numbers).  read: waits
 * for the stream, then returns the summed milliseconds and the number of timed scopes of `cls` since the last reset. */
enum { COCG_PROF_MSM_SORT = 0, COCG_PROF_MSM_ACCUMULATE = 1, COCG_PROF_MSM_REDUCE = 2, COCG_PROF_NTT = 3, COCG_PROF_VEC = 4,
       COCG_PROF_SPMV = 5, COCG_PROF_CLASSES = 6 };
COCG_API int cocg_profile_enable(cocg_ctx* ctx, int on);
COCG_API int cocg_profile_read(cocg_ctx* ctx, int cls, double* total_ms, uint64_t* scopes);
COCG_API int cocg_profile_reset(cocg_ctx* ctx);
/* Measures, on this device and now, the throughput of a pure dependent chain of the library's own Montgomery product on every SM
 * (base_field = 0: Fr, 1: Fq of the context's curve), in 10^9 products / s: the multiplier-issue ceiling bench.py reports the MSM
 * and NTT kernels against (a ~5 ms kernel). */
COCG_API int cocg_fp_mul_ceiling(cocg_ctx* ctx, int base_field, double* gmul_per_s);

/* ---- device memory (so that share vectors can stay resident between MPC network rounds) ------------- */
COCG_API int cocg_malloc(cocg_ctx* ctx, size_t bytes, void** dptr);
COCG_API int cocg_free(cocg_ctx* ctx, void* dptr);
COCG_API int cocg_h2d(cocg_ctx* ctx, void* dptr, const void* hptr, size_t bytes); /* synchronous */
COCG_API int cocg_d2h(cocg_ctx* ctx, void* hptr, const void* dptr, size_t bytes); /* synchronous */
COCG_API int cocg_memset0(cocg_ctx* ctx, void* dptr, size_t bytes);
COCG_API int cocg_d2d(cocg_ctx* ctx, void* dst, const void* src, size_t bytes); /* asynchronous; clone_from_slice rep3.rs:710-725 */
/* Page-locked host buffers: the staging memory for share vectors crossing PCIe (MPC network rounds, witness upload). */
COCG_API int cocg_host_alloc(cocg_ctx* ctx, size_t bytes, void** hptr);
COCG_API int cocg_host_free(cocg_ctx* ctx, void* hptr);

/* ---- a7: element-wise share-vector arithmetic (DEVICE pointers, n elements) --------------------------
 * Replaces PrimeFieldMpcProtocol::{add_vec, sub_assign_vec, neg_vec_in_place, mul (plain/Shamir local)}
 * traits.rs:67,146-149,161; rep3.rs:581-593,634-648,672-679; plain.rs:111-285.  `out` may alias `a`/`b`. */
COCG_API int cocg_vec_op(cocg_ctx* ctx, int op, const void* a, const void* b, void* out, size_t n);
/* out[i] = a * x[i] + y[i] (y may be NULL: out = a * x); a: HOST pointer to one Montgomery Fr, x / y / out DEVICE, out may alias.
 * The linear combinations of the Shamir driver: Lagrange interpolation at the king and re-sharing (shamir.rs:302-384), Vandermonde
 * extraction of the double-random pairs (shamir.rs:904-1010), ShamirCore::share (shamir/shamir_core.rs:8-33). */
COCG_API int cocg_vec_axpy(cocg_ctx* ctx, const void* a, const void* x, const void* y, void* out, size_t n);
/* a5: x[i] <- x[i] * c * g^i.  Replaces distribute_powers_and_mul_by_const traits.rs:177, rep3.rs:681-688.
 * g, c: HOST pointers to one Montgomery Fr each. */
COCG_API int cocg_vec_scale_powers(cocg_ctx* ctx, void* x, size_t n, const void* g, const void* c);
/* a6: REP3 mul_vec local step out[i] = aa*ba + aa*bb + ab*ba + mask[i] (mask may be NULL).
 * Replaces rep3.rs:656-660; the send_next/recv_prev round (rep3.rs:661-662) stays with the caller. */
COCG_API int cocg_rep3_mul_local(cocg_ctx* ctx, const void* aa, const void* ab, const void* ba, const void* bb,
                        const void* mask, void* out, size_t n);

/* Same with the zero-mask produced inside the kernel: mask[i] = F(seed_own, ctr, i) - F(seed_prev, ctr, i), F = the
 * counter-addressed ChaCha12 field PRF of csrc/prf.cuh.  Replaces rep3.rs:656-660 together with
 * Rep3Rand::masking_field_element (rep3/rngs.rs:37-46).  seed_*: HOST pointers to 32 bytes; ctr: a per-call counter the
 * three parties advance in lock-step. */
COCG_API int cocg_rep3_mul_local_prf(cocg_ctx* ctx, const void* aa, const void* ab, const void* ba, const void* bb,
                            const void* seed_own, const void* seed_prev, uint32_t ctr, void* out, size_t n);
/* out[i] = F(seed, ctr, i) (DEVICE, n elements), and the same function for one element on the HOST (no context needed):
 * Rep3Rand::random_fes (rngs.rs:42-46) for rand() / masking of single elements. */
COCG_API int cocg_prf_fill(cocg_ctx* ctx, const void* seed, uint32_t ctr, void* out, size_t n);
COCG_API int cocg_prf_field_host(int curve, const void* seed, uint32_t ctr, uint64_t idx, void* out);

/* ---- a4 (+a5): in-order radix-2 NTT over Fr, in place (DEVICE pointers) -------------------------------
 * Replaces FFTProvider::{fft_in_place, ifft_in_place} traits.rs:535-555; rep3.rs:880-921, shamir.rs:826-863,
 * plain.rs:369-400.  vecs: k device pointers (HOST array), each 2^log_n elements (k = 2 for a REP3 share
 * vector, 1 for Shamir / plain).  root: HOST pointer to the domain generator (Montgomery) -- the caller's
 * `domain.group_gen`, i.e. the snarkjs root for Groth16 (groth16.rs:62-71).  inverse != 0 computes the
 * inverse transform (uses root^-1, scales by n^-1).  coset_g: NULL, or HOST pointer to g; fuses
 * distribute_powers_and_mul_by_const(c = 1): forward -> input is pre-scaled by g^i, inverse -> output is
 * post-scaled by g^i (the ifft; distribute_powers pair of groth16.rs:175-186). */
COCG_API int cocg_ntt(cocg_ctx* ctx, void* const* vecs, int k, unsigned log_n, const void* root, int inverse,
             const void* coset_g);

/* ---- a1: variable-base MSM over resident public bases -------------------------------------------------
 * Replaces MSMProvider::msm_public_points traits.rs:561-568; rep3.rs:934-947, shamir.rs:1027-1039,
 * plain.rs:408-416 (-> ark-ec msm_unchecked).  Bases are uploaded once per zkey query (the reference
 * borrows &ZKey for the whole prove, groth16.rs:113-117) and addressed by handle.
 * pts: HOST pointer, n points, `stride` bytes apart (2*coordinate size for packed zkey bytes; 72/136/104/200
 * for arkworks Affine structs -- the trailing infinity flag is ignored, infinity is x = y = 0);
 * mont = 1 if coordinates are Montgomery limbs, 0 if canonical. */
COCG_API int cocg_bases_upload(cocg_ctx* ctx, int group, const void* pts, size_t n, size_t stride, int mont, uint64_t* handle);
COCG_API int cocg_bases_free(cocg_ctx* ctx, uint64_t handle);
/* Validates the resident points of `handle`: every point must satisfy the curve equation and, with check_subgroup != 0, lie in the
 * prime-order subgroup ([r]P = O; (0, 0) = infinity passes).  Replaces the per-point `is_on_curve` / `is_in_correct_subgroup_assuming_
 * on_curve` of the reference's zkey parser (circom-types/src/traits.rs:107-155, rayon loop :555-570).  *n_bad = number of offending
 * points, *first_bad (may be NULL) = smallest offending index.  Synchronises. */
COCG_API int cocg_bases_check(cocg_ctx* ctx, uint64_t handle, int check_subgroup, size_t* n_bad, size_t* first_bad);
/* The window plan of the resident table built for a query of n points: signed digits of window_bits bits, `windows` of them per scalar
 * (= table rows = additions per scalar).  No GPU needed; bench.py reports the additions per MSM from it. */
COCG_API int cocg_msm_plan(int curve, size_t n, int* window_bits, int* windows);
/* Synthetic bases generated in HBM: P0 + i*Q for two points derived from `seed` (HOST pointer, 32 bytes) -- valid, distinct
 * curve points for the 2^20..2^22 benchmark configurations, for which no zkey ships (csrc/gen.cu). */
COCG_API int cocg_bases_generate(cocg_ctx* ctx, int group, size_t n, const void* seed, uint64_t* handle);
/* The slice [first, first + n) of the same sequence: a rank of a multi-GPU run generates only the shard it accumulates. */
COCG_API int cocg_bases_generate_range(cocg_ctx* ctx, int group, size_t first, size_t n, const void* seed, uint64_t* handle);
/* Copy n packed affine points starting at `off` back to the HOST. */
COCG_API int cocg_bases_download(cocg_ctx* ctx, uint64_t handle, size_t off, size_t n, void* out);
/* Non-owning alias of another context's bases on the same device: the three REP3 drivers of one process borrow the
 * same &ZKey (tests/tests/circom/e2e_tests/mod.rs:55-70).  The owner must outlive the alias. */
COCG_API int cocg_bases_share(cocg_ctx* ctx, cocg_ctx* owner, uint64_t owner_handle, uint64_t* handle);
/* out[j] = sum_i scalars[j][i] * bases[off + i], i < n, for each of the k share components.
 * scalars: HOST array of k DEVICE pointers; scalars_mont = 1 if Montgomery form.
 * out_jacobian: HOST pointer, k Jacobian points. */
COCG_API int cocg_msm(cocg_ctx* ctx, uint64_t bases, size_t off, size_t n, const void* const* scalars, int k,
             int scalars_mont, void* out_jacobian);
/* Several queries times the SAME scalars: out[q][j] = sum_i scalars[j][i] * bases[q][offs[q] + i].  One digit sort per component
 * is shared by all queries whose tables have the same window width -- the a_query / b_g1_query / b_g2_query / l_query MSMs of
 * create_proof_with_assignment all take aux_assignment (groth16.rs:221-225, 251-255).  out_jacobian: HOST array of nq HOST
 * pointers, k Jacobian points each. */
COCG_API int cocg_msm_multi(cocg_ctx* ctx, const uint64_t* bases, const size_t* offs, int nq, size_t n, const void* const* scalars, int k,
                   int scalars_mont, void* const* out_jacobian);
/* Same with HOST scalar pointers (staged through pinned memory); the drop-in call for a host-resident caller. */
COCG_API int cocg_msm_host(cocg_ctx* ctx, uint64_t bases, size_t off, size_t n, const void* const* scalars, int k,
                  int scalars_mont, void* out_jacobian);

/* ---- a8: evaluate_constraint as CSR sparse mat-vec ----------------------------------------------------
 * Replaces the loop groth16.rs:159-166 over evaluate_constraint traits.rs:180-185 (rep3.rs:690-708,
 * plain.rs:243-251).  Matrix: rowptr[rows+1], col[nnz] (u32), coeff[nnz] (Montgomery Fr); uploaded once. */
COCG_API int cocg_csr_upload(cocg_ctx* ctx, const uint32_t* rowptr, const uint32_t* col, const void* coeff, size_t rows,
                    size_t nnz, uint64_t* handle);
/* Same with the coefficients in another stored form: COCG_FORM_R2 = value * R^2, as section 4 of a snarkjs zkey holds them
 * (circom-types/src/groth16/zkey.rs:181-194, traits.rs:65-67); COCG_FORM_CANONICAL = plain integers.  Converted on the device. */
enum { COCG_FORM_MONT = 0, COCG_FORM_R2 = 1, COCG_FORM_CANONICAL = 2 };
COCG_API int cocg_csr_upload_form(cocg_ctx* ctx, const uint32_t* rowptr, const uint32_t* col, const void* coeff, size_t rows, size_t nnz,
                         int coeff_form, uint64_t* handle);
/* Copy a resident matrix back to the HOST (any of rowptr / col / coeff may be NULL; *nnz is set when nnz != NULL). */
COCG_API int cocg_csr_download(cocg_ctx* ctx, uint64_t handle, uint32_t* rowptr, uint32_t* col, void* coeff, size_t* nnz);
COCG_API int cocg_csr_free(cocg_ctx* ctx, uint64_t handle);
COCG_API int cocg_csr_share(cocg_ctx* ctx, cocg_ctx* owner, uint64_t owner_handle, uint64_t* handle);
/* out[r] = sum_k coeff[k] * z[col[k]]; z = public inputs followed by the witness share, given as two DEVICE
 * arrays (z_pub: npub elements, may be NULL meaning zeros -- parties that do not add the public part,
 * rep3.rs:600-608 -- and z_wit).  out: DEVICE pointer, rows elements. */
COCG_API int cocg_spmv(cocg_ctx* ctx, uint64_t csr, const void* z_pub, size_t npub, const void* z_wit, void* out);

/* ---- CoPlonk vector primitives (csrc/poly.cu) -- DEVICE pointers, n Fr elements ---------------------------
 * out[i] = src[idx[i]] (idx: DEVICE u32 array; an index >= src_n selects zero).  The wire buffers of round 1: witness values picked
 * through the zkey's A / B / C maps (co-plonk/src/round1.rs:121-166, plonk_utils::get_witness lib.rs:113-137). */
COCG_API int cocg_vec_gather(cocg_ctx* ctx, const void* src, size_t src_n, const uint32_t* idx, void* out, size_t n);
/* Inclusive prefix scan, op = COCG_OP_MUL (array_prod_mul's `open[i] *= open[i-1]`, round2.rs:33-35) or COCG_OP_ADD (the closed form of
 * div_by_zerofier, round5.rs:97-115).  out may alias x. */
COCG_API int cocg_vec_scan(cocg_ctx* ctx, int op, const void* x, void* out, size_t n);
/* out[i] = 1 / x[i] by Montgomery's trick (inv_many's `y.inverse()` loop, rep3.rs:544-558).  Zero inputs give zero outputs and are
 * counted in *zeros (may be NULL: no synchronisation then) so the caller can raise "cannot compute inverse of zero".  out may alias x. */
COCG_API int cocg_vec_inv(cocg_ctx* ctx, const void* x, void* out, size_t n, size_t* zeros);
/* out (HOST, one Fr) = sum_i coeffs[i] * point^i: evaluate_poly_public (rep3.rs:923-928, round4.rs:136-142) per share component.
 * point: HOST pointer to one Montgomery Fr.  Synchronises. */
COCG_API int cocg_poly_eval(cocg_ctx* ctx, const void* coeffs, size_t n, const void* point, void* out);
/* out[i] = sum_k factors[k] * vecs[k][i] over the vectors with lens[k] > i (nv <= 8; factors: HOST array of nv Montgomery Fr):
 * the mul_with_public / add loops of compute_r and compute_wxi (round5.rs:140-330).  out must not alias an input. */
COCG_API int cocg_vec_lincomb(cocg_ctx* ctx, int nv, const void* const* vecs, const size_t* lens, const void* factors, void* out, size_t n);

/* out[i] = value (HOST pointer to one Fr): a shared scalar broadcast to a vector (`vec![r_inv[0].clone(); len]`, round2.rs:25). */
COCG_API int cocg_vec_fill(cocg_ctx* ctx, void* out, size_t n, const void* value);

/* ---- CoPlonk fused round kernels (csrc/plonk.cu) ----------------------------------------------------------------------------
 * Round 2, compute_z (round2.rs:146-206): numerator / denominator factors for ONE share component.
 * out = n1 n2 n3 d1 d2 d3; add_public = 1 for the component that receives public addends (add_with_public, rep3.rs:600-608). */
typedef struct cocg_plonk_z_args {
  const void *a, *b, *c;                 /* DEVICE: wire buffers, n elements */
  const void *sigma1, *sigma2, *sigma3;  /* DEVICE: 4n evaluations each (element 4 i is read) */
  const void *beta, *gamma, *k1, *k2;    /* HOST: one Montgomery Fr each */
  const void* omega;                     /* HOST: generator of the n domain */
  void* out[6];                          /* DEVICE */
} cocg_plonk_z_args;
COCG_API int cocg_plonk_z_factors(cocg_ctx* ctx, const cocg_plonk_z_args* args, size_t n, int add_public);
/* Round 3, compute_t (round3.rs:237-471) in two levels (see csrc/plonk.cu for the algebra): level 1 writes the party's additive, masked
 * shares of 6 product vectors (out: 6 x n4), level 2 those of t and tz (out: 2 x n4).  The caller re-shares after each level. */
typedef struct cocg_plonk_quotient_args {
  int components;                        /* 1 plain, 2 REP3 (a | b) */
  int pub_comp;                          /* component that receives public addends: 0 plain / party 0, 1 party 1, -1 party 2 */
  size_t n4;                             /* extended domain size 4n */
  size_t n_public, n_lagrange;           /* public inputs; Lagrange polynomials resident (>= max(1, n_public)) */
  const void *eval_a[2], *eval_b[2], *eval_c[2], *eval_z[2];   /* DEVICE: n4 evaluations per component */
  const void* buffer_a[2];               /* DEVICE: wire buffer a (its first n_public elements feed PI) */
  const void *sigma1, *sigma2, *sigma3, *qm, *ql, *qr, *qo, *qc;  /* DEVICE: n4 public evaluations each */
  const void* lagrange;                  /* DEVICE: n_lagrange x n4 evaluations, contiguous */
  const void* level1[2];                 /* DEVICE (level 2 only): the re-shared level-1 vectors, 6 x n4 per component */
  const void *beta, *gamma, *alpha, *k1, *k2, *omega_n, *omega_4n;  /* HOST Fr */
  const void* blinders;                  /* HOST: b0..b8 as 9 x 2 Fr (component a | b; b ignored when components = 1) */
  const void* scalar_products;           /* HOST (level 2): shares of b1b3 b0b3 b1b2 b0b2 b5b8 b5b7 b5b6 b4b8 b4b7 b4b6, 10 x 2 Fr */
  const void *seed_own, *seed_prev;      /* HOST: 32-byte PRF seeds (REP3 zero-masks; csrc/prf.cuh) */
  uint32_t ctr;                          /* first PRF vector counter; level 1 consumes 6, level 2 consumes 2 */
  void* out;                             /* DEVICE */
} cocg_plonk_quotient_args;
COCG_API int cocg_plonk_quotient_l1(cocg_ctx* ctx, const cocg_plonk_quotient_args* args);
COCG_API int cocg_plonk_quotient_l2(cocg_ctx* ctx, const cocg_plonk_quotient_args* args);
/* Coefficients of T from ifft(t), ifft(tz) (n4 = 4n each): negate / divide by Z_H / add / split into t1 (n + 1) | t2 (n + 1) | t3 (n + 6)
 * (round3.rs:433-468); elements t1[n], t2[n] are left for the caller's b9 / b10 patches. */
COCG_API int cocg_plonk_t_finish(cocg_ctx* ctx, const void* ct, const void* ctz, size_t n, void* t1, void* t2, void* t3);

/* ---- K7: O(1) group operations on HOST Jacobian points (proof assembly, groth16.rs:257-312) -----------
 * op: 0 add(a,b)  1 scalar-mul(a, b = canonical 32-byte scalar)  2 to_affine(a) -> packed affine
 *     3 from_affine(a)  4 neg(a)  5 double(a)  6 generator() (a, b ignored).  Replaces EcMpcProtocol::{add_points, scalar_mul_public_point,
 *     ...} traits.rs:472-522 for single points. */
COCG_API int cocg_ec_op(cocg_ctx* ctx, int group, int op, const void* a, const void* b, void* out);

#ifdef __cplusplus
}
#endif
#endif /* COCG_H */
