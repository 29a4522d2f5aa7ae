/* cocg CPU oracle -- TEST INFRASTRUCTURE ONLY (never linked into or called from the product path).
 *
 * Plain-C restatement (64-bit limbs, unsigned __int128) of the reference's CPU hot path so the
 * CUDA kernels can be checked bit-for-bit and so bench.py has a CPU arm ("port"):
 *   - batched Fr arithmetic / REP3 mul_vec local step   mpc-core/src/protocols/rep3.rs:581-688
 *   - evaluate_constraint (sparse mat-vec)               rep3.rs:690-708, co-groth16/src/groth16.rs:159-166
 *   - in-order radix-2 NTT/iNTT, snarkjs roots           rep3.rs:880-921, groth16.rs:57-77
 *   - distribute_powers_and_mul_by_const                 rep3.rs:681-688
 *   - variable-base MSM (signed-digit Pippenger)         rep3.rs:934-947 -> ark-ec 0.4.2 msm_unchecked
 * (paths relative to /root/reference).  The arithmetic itself lives in arkworks 0.4.x, which is
 * not vendored and cannot be built here (no cargo): parity is pinned through the Python big-int
 * oracle, which in turn is pinned by the reference's fixtures (tests/test_oracle_pinned.py).
 *
 * Element layout everywhere: little-endian u64 limbs in MONTGOMERY form (what a Rust caller holding
 * ark_ff::Fp would hand over); affine points packed (x,y) with (0,0) = infinity. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "params64_gen.h"

#define FP bn_fr
#define NL 4
#define FP_MOD BN254_FR_MOD
#define FP_R1 BN254_FR_R1
#define FP_R2 BN254_FR_R2
#define FP_INV BN254_FR_INV
#include "fp_tmpl.h"

#define FP bn_fq
#define NL 4
#define FP_MOD BN254_FQ_MOD
#define FP_R1 BN254_FQ_R1
#define FP_R2 BN254_FQ_R2
#define FP_INV BN254_FQ_INV
#include "fp_tmpl.h"

#define FP bls_fr
#define NL 4
#define FP_MOD BLS381_FR_MOD
#define FP_R1 BLS381_FR_R1
#define FP_R2 BLS381_FR_R2
#define FP_INV BLS381_FR_INV
#include "fp_tmpl.h"

#define FP bls_fq
#define NL 6
#define FP_MOD BLS381_FQ_MOD
#define FP_R1 BLS381_FQ_R1
#define FP_R2 BLS381_FQ_R2
#define FP_INV BLS381_FQ_INV
#include "fp_tmpl.h"

#define F2 bn_fq2
#define FQ bn_fq
#define FQ_R1_INIT BN254_FQ_R1
#include "fp2_tmpl.h"

#define F2 bls_fq2
#define FQ bls_fq
#define FQ_R1_INIT BLS381_FQ_R1
#include "fp2_tmpl.h"

#define EC bn_g1
#define EF bn_fq
#include "ec_tmpl.h"
#define EC bn_g2
#define EF bn_fq2
#include "ec_tmpl.h"
#define EC bls_g1
#define EF bls_fq
#include "ec_tmpl.h"
#define EC bls_g2
#define EF bls_fq2
#include "ec_tmpl.h"

#define EXPORT __attribute__((visibility("default")))

EXPORT int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
EXPORT void orc_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------ Fr vectors */
#define FR_DISPATCH(curve, CALL_BN, CALL_BLS) do { if ((curve) == 0) { CALL_BN; } else { CALL_BLS; } } while (0)

/* op: 0 mul, 1 add, 2 sub, 3 neg(a), 4 to_mont(a), 5 from_mont(a) */
#define DEF_VEC_OP(P)                                                                              \
  static void P##_vec_op(int op, const P##_t *a, const P##_t *b, P##_t *o, size_t n) {             \
    _Pragma("omp parallel for schedule(static)") for (size_t i = 0; i < n; i++) {                  \
      switch (op) {                                                                                \
        case 0: P##_mul(&o[i], &a[i], &b[i]); break;                                               \
        case 1: P##_add(&o[i], &a[i], &b[i]); break;                                               \
        case 2: P##_sub(&o[i], &a[i], &b[i]); break;                                               \
        case 3: P##_neg(&o[i], &a[i]); break;                                                      \
        case 4: P##_to_mont(&o[i], &a[i]); break;                                                  \
        default: P##_from_mont(&o[i], &a[i]); break;                                               \
      }                                                                                            \
    }                                                                                              \
  }                                                                                                \
  /* mul_vec local step, rep3.rs:656-660: aa*ba + aa*bb + ab*ba + mask */                          \
  static void P##_rep3_mul_local(const P##_t *aa, const P##_t *ab, const P##_t *ba,                \
                                 const P##_t *bb, const P##_t *mask, P##_t *o, size_t n) {         \
    _Pragma("omp parallel for schedule(static)") for (size_t i = 0; i < n; i++) {                  \
      P##_t t0, t1, t2;                                                                            \
      P##_mul(&t0, &aa[i], &ba[i]);                                                                \
      P##_mul(&t1, &aa[i], &bb[i]);                                                                \
      P##_mul(&t2, &ab[i], &ba[i]);                                                                \
      P##_add(&t0, &t0, &t1);                                                                      \
      P##_add(&t0, &t0, &t2);                                                                      \
      if (mask) P##_add(&t0, &t0, &mask[i]);                                                       \
      o[i] = t0;                                                                                   \
    }                                                                                              \
  }                                                                                                \
  /* x_i <- x_i * c * g^i  (rep3.rs:681-688), chunked so the dependent chain is per chunk */       \
  static void P##_distribute_powers(P##_t *x, size_t n, const P##_t *g, const P##_t *c) {          \
    size_t chunk = 4096;                                                                           \
    size_t nch = (n + chunk - 1) / chunk;                                                          \
    _Pragma("omp parallel for schedule(static)") for (size_t ch = 0; ch < nch; ch++) {             \
      uint64_t e[1] = {(uint64_t)(ch * chunk)};                                                    \
      P##_t pw;                                                                                    \
      P##_pow(&pw, g, e, 1);                                                                       \
      P##_mul(&pw, &pw, c);                                                                        \
      size_t hi = (ch + 1) * chunk < n ? (ch + 1) * chunk : n;                                     \
      for (size_t i = ch * chunk; i < hi; i++) {                                                   \
        P##_mul(&x[i], &x[i], &pw);                                                                \
        P##_mul(&pw, &pw, g);                                                                      \
      }                                                                                            \
    }                                                                                              \
  }                                                                                                \
  /* CSR sparse mat-vec: out[r] = sum_k coeff[k] * z[col[k]]  (evaluate_constraint) */             \
  static void P##_spmv(const uint32_t *rowptr, const uint32_t *col, const P##_t *coeff,            \
                       const P##_t *z, P##_t *o, size_t rows) {                                    \
    _Pragma("omp parallel for schedule(static)") for (size_t r = 0; r < rows; r++) {               \
      P##_t acc, t;                                                                                \
      memset(&acc, 0, sizeof(acc));                                                                \
      for (uint32_t k = rowptr[r]; k < rowptr[r + 1]; k++) {                                       \
        P##_mul(&t, &coeff[k], &z[col[k]]);                                                        \
        P##_add(&acc, &acc, &t);                                                                   \
      }                                                                                            \
      o[r] = acc;                                                                                  \
    }                                                                                              \
  }                                                                                                \
  /* in-order radix-2 DFT, in place; inverse: caller passes omega^-1, we scale by n^-1 */          \
  static void P##_ntt(P##_t *a, unsigned logn, const P##_t *omega, int inverse) {                  \
    size_t n = (size_t)1 << logn;                                                                  \
    if (logn == 0) return;                                                                         \
    for (size_t i = 0; i < n; i++) {                                                               \
      size_t j = 0;                                                                                \
      for (unsigned b = 0; b < logn; b++) j |= ((i >> b) & 1) << (logn - 1 - b);                   \
      if (j > i) { P##_t t = a[i]; a[i] = a[j]; a[j] = t; }                                        \
    }                                                                                              \
    P##_t *tw = (P##_t *)malloc((n / 2) * sizeof(P##_t));                                          \
    size_t chunk = 1024;                                                                           \
    size_t nch = (n / 2 + chunk - 1) / chunk;                                                      \
    _Pragma("omp parallel for schedule(static)") for (size_t ch = 0; ch < nch; ch++) {             \
      uint64_t e[1] = {(uint64_t)(ch * chunk)};                                                    \
      P##_t pw;                                                                                    \
      P##_pow(&pw, omega, e, 1);                                                                   \
      size_t hi = (ch + 1) * chunk < n / 2 ? (ch + 1) * chunk : n / 2;                             \
      for (size_t i = ch * chunk; i < hi; i++) { tw[i] = pw; P##_mul(&pw, &pw, omega); }           \
    }                                                                                              \
    for (unsigned s = 0; s < logn; s++) {                                                          \
      size_t m = (size_t)1 << s;                                                                   \
      size_t stride = n / (2 * m);                                                                 \
      _Pragma("omp parallel for schedule(static)") for (size_t idx = 0; idx < n / 2; idx++) {      \
        size_t k = (idx / m) * 2 * m, j = idx % m;                                                 \
        P##_t t, u = a[k + j];                                                                     \
        P##_mul(&t, &a[k + j + m], &tw[j * stride]);                                               \
        P##_add(&a[k + j], &u, &t);                                                                \
        P##_sub(&a[k + j + m], &u, &t);                                                            \
      }                                                                                            \
    }                                                                                              \
    free(tw);                                                                                      \
    if (inverse) {                                                                                 \
      P##_t nn, ninv;                                                                              \
      memset(&nn, 0, sizeof(nn));                                                                  \
      nn.l[0] = n;                                                                                 \
      P##_to_mont(&nn, &nn);                                                                       \
      P##_inv(&ninv, &nn);                                                                         \
      _Pragma("omp parallel for schedule(static)") for (size_t i = 0; i < n; i++)                  \
          P##_mul(&a[i], &a[i], &ninv);                                                            \
    }                                                                                              \
  }

DEF_VEC_OP(bn_fr)
DEF_VEC_OP(bls_fr)

EXPORT void orc_fr_vec_op(int curve, int op, const void *a, const void *b, void *o, size_t n) {
  FR_DISPATCH(curve, bn_fr_vec_op(op, a, b, o, n), bls_fr_vec_op(op, a, b, o, n));
}
EXPORT void orc_rep3_mul_local(int curve, const void *aa, const void *ab, const void *ba, const void *bb,
                               const void *mask, void *o, size_t n) {
  FR_DISPATCH(curve, bn_fr_rep3_mul_local(aa, ab, ba, bb, mask, o, n), bls_fr_rep3_mul_local(aa, ab, ba, bb, mask, o, n));
}
EXPORT void orc_distribute_powers(int curve, void *x, size_t n, const void *g, const void *c) {
  FR_DISPATCH(curve, bn_fr_distribute_powers(x, n, g, c), bls_fr_distribute_powers(x, n, g, c));
}
EXPORT void orc_spmv(int curve, const uint32_t *rowptr, const uint32_t *col, const void *coeff, const void *z,
                     void *o, size_t rows) {
  FR_DISPATCH(curve, bn_fr_spmv(rowptr, col, coeff, z, o, rows), bls_fr_spmv(rowptr, col, coeff, z, o, rows));
}
EXPORT void orc_ntt(int curve, void *a, unsigned logn, const void *omega, int inverse) {
  FR_DISPATCH(curve, bn_fr_ntt(a, logn, omega, inverse), bls_fr_ntt(a, logn, omega, inverse));
}

/* ------------------------------------------------------------------ MSM */
static int msm_window(size_t n) {
  if (n < 32) return 3;
  int lg = 63 - __builtin_clzll((unsigned long long)n);
  return lg * 69 / 100 + 2;
}
/* signed c-bit digits of plain 256-bit integers: s = sum_w dig[w] * 2^(c*w), |dig| <= 2^(c-1) */
static void signed_digits(const uint64_t *sc, size_t n, int c, int nwin, int32_t *dig) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    const uint64_t *s = sc + 4 * i;
    int carry = 0;
    for (int w = 0; w < nwin; w++) {
      int bit = w * c;
      uint64_t raw = 0;
      if (bit < 256) {
        raw = s[bit / 64] >> (bit % 64);
        if (bit % 64 + c > 64 && bit / 64 + 1 < 4) raw |= s[bit / 64 + 1] << (64 - bit % 64);
        raw &= ((uint64_t)1 << c) - 1;
      }
      int64_t d = (int64_t)raw + carry;
      carry = 0;
      if (d > ((int64_t)1 << (c - 1))) { d -= (int64_t)1 << c; carry = 1; }
      dig[i * (size_t)nwin + w] = (int32_t)d;
    }
  }
}

#define DEF_MSM(G, FR, NBITS)                                                                      \
  static void G##_msm(const G##_aff *pts, const FR##_t *sc_mont, size_t n, G##_jac *out) {         \
    if (n == 0) { G##_set_inf(out); return; }                                                      \
    int c = msm_window(n);                                                                         \
    int nwin = NBITS / c + 1;                                                                      \
    uint64_t *plain = (uint64_t *)malloc(n * 32);                                                  \
    FR##_vec_op(5, sc_mont, NULL, (FR##_t *)plain, n); /* into_bigint */                           \
    int32_t *dig = (int32_t *)malloc(n * (size_t)nwin * sizeof(int32_t));                          \
    signed_digits(plain, n, c, nwin, dig);                                                         \
    free(plain);                                                                                   \
    G##_jac *win = (G##_jac *)malloc(nwin * sizeof(G##_jac));                                      \
    _Pragma("omp parallel for schedule(dynamic, 1)") for (int w = 0; w < nwin; w++)                \
        G##_msm_one_window(pts, dig, n, c, w, nwin, &win[w]);                                      \
    G##_msm_fold(win, nwin, c, out);                                                               \
    free(win);                                                                                     \
    free(dig);                                                                                     \
  }

DEF_MSM(bn_g1, bn_fr, 254)
DEF_MSM(bn_g2, bn_fr, 254)
DEF_MSM(bls_g1, bls_fr, 255)
DEF_MSM(bls_g2, bls_fr, 255)

/* group: 1 = G1, 2 = G2.  pts: n packed affine points; scalars: n Fr (Montgomery); out: Jacobian (X,Y,Z). */
EXPORT void orc_msm(int curve, int group, const void *pts, const void *scalars, size_t n, void *out) {
  if (curve == 0) {
    if (group == 1) bn_g1_msm(pts, scalars, n, out); else bn_g2_msm(pts, scalars, n, out);
  } else {
    if (group == 1) bls_g1_msm(pts, scalars, n, out); else bls_g2_msm(pts, scalars, n, out);
  }
}

/* ------------------------------------------------------------------ small EC helpers
 * op: 0 add(jac a, jac b) -> jac ; 1 mul(jac a, plain 4-limb scalar b) -> jac ; 2 to_affine(jac a) -> aff ;
 *     3 from_affine(aff a) -> jac ; 4 neg(jac a) -> jac ; 5 double */
#define DEF_EC_OP(G)                                                                               \
  static void G##_op(int op, const void *a, const void *b, void *o) {                              \
    switch (op) {                                                                                  \
      case 0: { G##_jac r; G##_add(&r, (const G##_jac *)a, (const G##_jac *)b); *(G##_jac *)o = r; break; } \
      case 1: { G##_jac r; G##_mul(&r, (const G##_jac *)a, (const uint64_t *)b); *(G##_jac *)o = r; break; } \
      case 2: { G##_aff r; G##_to_aff(&r, (const G##_jac *)a); *(G##_aff *)o = r; break; }          \
      case 3: { G##_jac r; G##_from_aff(&r, (const G##_aff *)a); *(G##_jac *)o = r; break; }        \
      case 4: { G##_jac r; G##_neg(&r, (const G##_jac *)a); *(G##_jac *)o = r; break; }             \
      default: { G##_jac r; G##_dbl(&r, (const G##_jac *)a); *(G##_jac *)o = r; break; }            \
    }                                                                                              \
  }
DEF_EC_OP(bn_g1)
DEF_EC_OP(bn_g2)
DEF_EC_OP(bls_g1)
DEF_EC_OP(bls_g2)

EXPORT void orc_ec_op(int curve, int group, int op, const void *a, const void *b, void *o) {
  if (curve == 0) {
    if (group == 1) bn_g1_op(op, a, b, o); else bn_g2_op(op, a, b, o);
  } else {
    if (group == 1) bls_g1_op(op, a, b, o); else bls_g2_op(op, a, b, o);
  }
}

/* Deterministic synthetic bases: P_0 = k0*G-like start point `p0` (affine), P_{i+1} = P_i + Q, all
 * normalised to affine.  Used by tests/bench to make 2^20 valid curve points quickly on the CPU.
 * Each chunk of 1024 restarts from p0 + (chunk*1024)*Q computed by scalar mul. */
#define DEF_GEN(G, FQP)                                                                            \
  static void G##_gen_chain(const G##_aff *p0, const G##_aff *q, size_t n, G##_aff *out) {         \
    size_t chunk = 1024;                                                                           \
    size_t nch = (n + chunk - 1) / chunk;                                                          \
    _Pragma("omp parallel for schedule(dynamic, 1)") for (size_t ch = 0; ch < nch; ch++) {         \
      G##_jac acc, qj, t;                                                                          \
      uint64_t k[4] = {(uint64_t)(ch * chunk), 0, 0, 0};                                           \
      G##_from_aff(&qj, q);                                                                        \
      G##_mul(&t, &qj, k);                                                                         \
      G##_from_aff(&acc, p0);                                                                      \
      G##_add(&acc, &acc, &t);                                                                     \
      size_t hi = (ch + 1) * chunk < n ? (ch + 1) * chunk : n;                                     \
      for (size_t i = ch * chunk; i < hi; i++) {                                                   \
        G##_to_aff(&out[i], &acc);                                                                 \
        G##_madd(&acc, &acc, q);                                                                   \
      }                                                                                            \
    }                                                                                              \
  }
DEF_GEN(bn_g1, bn_fq)
DEF_GEN(bn_g2, bn_fq2)
DEF_GEN(bls_g1, bls_fq)
DEF_GEN(bls_g2, bls_fq2)

EXPORT void orc_gen_chain(int curve, int group, const void *p0, const void *q, size_t n, void *out) {
  if (curve == 0) {
    if (group == 1) bn_g1_gen_chain(p0, q, n, out); else bn_g2_gen_chain(p0, q, n, out);
  } else {
    if (group == 1) bls_g1_gen_chain(p0, q, n, out); else bls_g2_gen_chain(p0, q, n, out);
  }
}
