// Groth16 verification on the host: the acceptance check the reference runs after every proof (`Groth16::verify`,
// /root/reference/co-circom/co-groth16/src/verifier.rs:23-43 -> ark_groth16::verify_proof; `co-circom verify`,
// co-circom/src/bin/co-circom.rs:640-720).  Three pairings per proof are O(1) latency-bound work, so -- like the O(1) group operations
// of proof assembly (csrc/api.cu) -- it runs on the calling host thread with the host branch of fp.cuh / ec.cuh; no GPU is needed.
//
//   e(A, B) == e(alpha, beta) * e(vk_x, gamma) * e(C, delta),   vk_x = IC[0] + sum_i public_i * IC[i + 1]
// is checked as  prod of four Miller loops, one final exponentiation, == 1.
//
// Optimal-ate Miller loop with affine steps on the twist (one Fq2 inversion per step), Fq12 as the flat extension
// Fq[w] / (w^12 - 2a w^6 + a^2 + 1) with u = w^6 - a, xi = a + u (BN254: a = 9, D-type twist, loop 6x + 2 plus the two Frobenius steps;
// BLS12-381: a = 1, M-type twist, loop |x|), final exponentiation by plain square-and-multiply with (q^12 - 1) / r
// (pairing_params_gen.h).  Any bilinear non-degenerate pairing validates the Groth16 equation, so lines are scaled by subfield
// elements freely and the sign of the BLS parameter is ignored.  ~0.05 s (BN254) / ~0.2 s (BLS12-381) per proof on one core.
#pragma once
#include <string>
#include <vector>

#include "driver.hpp"
#include "pairing_params_gen.h"

namespace cohost {

template <class P, int A, bool TWIST_D>
struct PairingEngine {
  using Fq = cocg::Fp<P>;
  using Fq2 = cocg::Fp2<P>;
  using G1 = cocg::Affine<Fq>;
  using G2 = cocg::Affine<Fq2>;
  struct F12 {
    Fq c[12];
  };

  static Fq small(int v) {
    Fq r = Fq::zero(), one = Fq::one();
    for (int i = 0; i < v; i++) r = cocg::fp_add(r, one);
    return r;
  }
  static F12 f12_one() {
    F12 r;
    for (auto& x : r.c) x = Fq::zero();
    r.c[0] = Fq::one();
    return r;
  }
  static bool f12_is_one(const F12& x) {
    if (!(x.c[0] == Fq::one())) return false;
    for (int i = 1; i < 12; i++)
      if (!x.c[i].is_zero()) return false;
    return true;
  }
  // schoolbook product, zero coefficients skipped (line values have three or four), then w^12 = 2a w^6 - (a^2 + 1)
  static F12 f12_mul(const F12& x, const F12& y) {
    static const Fq c6 = small(2 * A), c0 = cocg::fp_neg(small(A * A + 1));
    Fq t[23];
    for (auto& v : t) v = Fq::zero();
    bool ynz[12];
    for (int j = 0; j < 12; j++) ynz[j] = !y.c[j].is_zero();
    for (int i = 0; i < 12; i++) {
      if (x.c[i].is_zero()) continue;
      for (int j = 0; j < 12; j++)
        if (ynz[j]) t[i + j] = cocg::fp_add(t[i + j], cocg::fp_mul(x.c[i], y.c[j]));
    }
    for (int k = 22; k >= 12; k--) {
      if (t[k].is_zero()) continue;
      t[k - 6] = cocg::fp_add(t[k - 6], cocg::fp_mul(t[k], c6));
      t[k - 12] = cocg::fp_add(t[k - 12], cocg::fp_mul(t[k], c0));
    }
    F12 r;
    for (int i = 0; i < 12; i++) r.c[i] = t[i];
    return r;
  }
  static F12 f12_pow(const F12& x, const uint32_t* e, int limbs) {
    F12 r = f12_one();
    bool started = false;
    for (int i = 32 * limbs - 1; i >= 0; i--) {
      if (started) r = f12_mul(r, r);
      if ((e[i >> 5] >> (i & 31)) & 1) {
        r = started ? f12_mul(r, x) : x;
        started = true;
      }
    }
    return r;
  }
  // z = z0 + z1 u (u = w^6 - a) times w^shift, added into out
  static void add_embedded(F12& out, const Fq2& z, int shift) {
    static const Fq a = small(A);
    out.c[shift] = cocg::fp_add(out.c[shift], cocg::fp_sub(z.c0, cocg::fp_mul(a, z.c1)));
    out.c[shift + 6] = cocg::fp_add(out.c[shift + 6], z.c1);
  }
  static Fq2 f2_scale(const Fq2& z, const Fq& s) { return Fq2{cocg::fp_mul(z.c0, s), cocg::fp_mul(z.c1, s)}; }
  static Fq2 f2_conj(const Fq2& z) { return Fq2{z.c0, cocg::fp_neg(z.c1)}; }

  // Line through R and Q (tangent at R when Q == nullptr) on the twist, evaluated at P; R <- R + Q (or 2R).
  // Returns false when the line is vertical (R = -Q): cannot happen for points of prime order r inside the loop.
  static bool line_and_step(G2& R, const G2* Q, const G1& Pt, F12& line) {
    using namespace cocg;
    Fq2 num, den;
    if (!Q || (R.x == Q->x && R.y == Q->y)) {
      Fq2 x2 = f_sqr(R.x);
      num = f_add(f_dbl(x2), x2);
      den = f_dbl(R.y);
    } else {
      num = f_sub(Q->y, R.y);
      den = f_sub(Q->x, R.x);
    }
    if (den.is_zero()) return false;
    const Fq2 lam = f_mul(num, f_inv(den));
    const Fq2 k0 = f_sub(R.y, f_mul(lam, R.x));  // yR - lam xR
    const Fq2 k1 = f2_scale(lam, Pt.x);          // lam xP
    for (auto& v : line.c) v = Fq::zero();
    if (TWIST_D) {  // l = -yP + lam xP w + (yR - lam xR) w^3
      add_embedded(line, k1, 1);
      add_embedded(line, k0, 3);
      line.c[0] = fp_sub(line.c[0], Pt.y);
    } else {        // l w^3 = -yP w^3 + lam xP w^2 + (yR - lam xR)
      add_embedded(line, k1, 2);
      add_embedded(line, k0, 0);
      line.c[3] = fp_sub(line.c[3], Pt.y);
    }
    const Fq2 xq = Q ? Q->x : R.x;
    const Fq2 x3 = f_sub(f_sub(f_sqr(lam), R.x), xq);
    const Fq2 y3 = f_sub(f_mul(lam, f_sub(R.x, x3)), R.y);
    R = G2{x3, y3};
    return true;
  }

  static Fq2 load_f2(const uint32_t* c0, const uint32_t* c1) {
    Fq2 r;
    memcpy(r.c0.l, c0, sizeof(r.c0.l));
    memcpy(r.c1.l, c1, sizeof(r.c1.l));
    return r;
  }

  // f_{loop, Q}(P); P, Q finite
  static F12 miller(const G1& Pt, const G2& Q, const uint32_t* loop, int loop_bits, const Fq2* gamma2, const Fq2* gamma3) {
    F12 acc = f12_one(), l;
    G2 R = Q;
    for (int i = loop_bits - 2; i >= 0; i--) {
      if (!line_and_step(R, nullptr, Pt, l)) throw Error("pairing: degenerate doubling step (point of small order)");
      acc = f12_mul(f12_mul(acc, acc), l);
      if ((loop[i >> 5] >> (i & 31)) & 1) {
        if (!line_and_step(R, &Q, Pt, l)) throw Error("pairing: degenerate addition step (point of small order)");
        acc = f12_mul(acc, l);
      }
    }
    if (TWIST_D) {  // BN: Q1 = pi(Q), Q2 = pi^2(Q); lines through (R, Q1) and (R + Q1, -Q2)
      using namespace cocg;
      G2 q1{f_mul(f2_conj(Q.x), *gamma2), f_mul(f2_conj(Q.y), *gamma3)};
      G2 q2{f_mul(f2_conj(q1.x), *gamma2), f_mul(f2_conj(q1.y), *gamma3)};
      G2 nq2{q2.x, f_neg(q2.y)};
      if (!line_and_step(R, &q1, Pt, l)) throw Error("pairing: degenerate Frobenius step");
      acc = f12_mul(acc, l);
      if (!line_and_step(R, &nq2, Pt, l)) throw Error("pairing: degenerate Frobenius step");
      acc = f12_mul(acc, l);
    }
    return acc;
  }
};

// ---- per-curve glue ------------------------------------------------------------------------------------------------------------
struct Bn254Pairing {
  using E = PairingEngine<cocg::Bn254FqP, 9, true>;
  static E::F12 miller(const E::G1& p, const E::G2& q) {
    static const uint32_t loop[3] = PAIRING_BN254_LOOP;
    static const uint32_t g2c0[8] = PAIRING_BN254_GAMMA2_C0, g2c1[8] = PAIRING_BN254_GAMMA2_C1, g3c0[8] = PAIRING_BN254_GAMMA3_C0,
                          g3c1[8] = PAIRING_BN254_GAMMA3_C1;
    const E::Fq2 gamma2 = E::load_f2(g2c0, g2c1), gamma3 = E::load_f2(g3c0, g3c1);
    return E::miller(p, q, loop, PAIRING_BN254_LOOP_BITS, &gamma2, &gamma3);
  }
  static bool final_is_one(const E::F12& f) {
    static const uint32_t e[PAIRING_BN254_FINAL_EXP_LIMBS] = PAIRING_BN254_FINAL_EXP;
    return E::f12_is_one(E::f12_pow(f, e, PAIRING_BN254_FINAL_EXP_LIMBS));
  }
  static const uint32_t* order() {
    static const uint32_t r[8] = PAIRING_BN254_SUBGROUP_ORDER;
    return r;
  }
  static E::Fq b1() { static const uint32_t v[8] = BN254_G1_B; E::Fq r; memcpy(r.l, v, sizeof(v)); return r; }
  static E::Fq2 b2() { static const uint32_t c0[8] = BN254_G2_B_C0, c1[8] = BN254_G2_B_C1; return E::load_f2(c0, c1); }
  static E::G1 g1_generator() { static const uint32_t g[2][8] = BN254_G1_GEN; E::G1 r; memcpy(&r, g, sizeof(r)); return r; }
  static E::G2 g2_generator() { static const uint32_t g[4][8] = BN254_G2_GEN; E::G2 r; memcpy(&r, g, sizeof(r)); return r; }
};
struct Bls381Pairing {
  using E = PairingEngine<cocg::Bls381FqP, 1, false>;
  static E::F12 miller(const E::G1& p, const E::G2& q) {
    static const uint32_t loop[3] = PAIRING_BLS381_LOOP;
    return E::miller(p, q, loop, PAIRING_BLS381_LOOP_BITS, nullptr, nullptr);
  }
  static bool final_is_one(const E::F12& f) {
    static const uint32_t e[PAIRING_BLS381_FINAL_EXP_LIMBS] = PAIRING_BLS381_FINAL_EXP;
    return E::f12_is_one(E::f12_pow(f, e, PAIRING_BLS381_FINAL_EXP_LIMBS));
  }
  static const uint32_t* order() {
    static const uint32_t r[8] = PAIRING_BLS381_SUBGROUP_ORDER;
    return r;
  }
  static E::Fq b1() { static const uint32_t v[12] = BLS381_G1_B; E::Fq r; memcpy(r.l, v, sizeof(v)); return r; }
  static E::Fq2 b2() { static const uint32_t c0[12] = BLS381_G2_B_C0, c1[12] = BLS381_G2_B_C1; return E::load_f2(c0, c1); }
  static E::G1 g1_generator() { static const uint32_t g[2][12] = BLS381_G1_GEN; E::G1 r; memcpy(&r, g, sizeof(r)); return r; }
  static E::G2 g2_generator() { static const uint32_t g[4][12] = BLS381_G2_GEN; E::G2 r; memcpy(&r, g, sizeof(r)); return r; }
};

// k * p for a canonical little-endian scalar of 8 x 32 bits (double-and-add, XYZZ accumulators)
template <class F>
cocg::XYZZ<F> scalar_mul_affine(const cocg::Affine<F>& p, const uint32_t* k) {
  cocg::XYZZ<F> acc = cocg::xyzz_inf<F>();
  for (int i = 255; i >= 0; i--) {
    acc = cocg::xyzz_dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) cocg::xyzz_madd(acc, p);
  }
  return acc;
}
template <class F>
cocg::Affine<F> xyzz_to_affine(const cocg::XYZZ<F>& p) {
  if (p.is_inf()) return cocg::Affine<F>{F::zero(), F::zero()};
  F zi = cocg::f_inv(p.zzz);             // 1 / ZZZ
  F zzi = cocg::f_mul(zi, zi);           // 1 / ZZ^3 ...
  zzi = cocg::f_mul(zzi, cocg::f_sqr(p.zz));  // ZZ^2 / ZZZ^2 = 1 / ZZ   (ZZ^3 = ZZZ^2)
  return cocg::Affine<F>{cocg::f_mul(p.x, zzi), cocg::f_mul(p.y, zi)};
}
template <class F>
bool on_curve(const cocg::Affine<F>& p, const F& b) {  // y^2 = x^3 + b; infinity counts as on the curve
  if (p.is_inf()) return true;
  return cocg::f_sqr(p.y) == cocg::f_add(cocg::f_mul(cocg::f_sqr(p.x), p.x), b);
}

struct VerifyInput {  // packed affine Montgomery, as everywhere in the C ABI
  const uint64_t *alpha_g1, *beta_g2, *gamma_g2, *delta_g2;
  const uint64_t* ic;  // n_ic G1 points
  size_t n_ic;
  const uint64_t* proof;   // A | B | C
  const uint64_t* pub;     // n_ic - 1 Montgomery Fr (without the leading 1)
};

// Returns true iff the proof is accepted.  Throws Error for malformed input (counts, points off the curve / outside the subgroup),
// mirroring where the reference fails at deserialisation (circom-types/src/traits.rs:160-184) rather than at the pairing check.
template <class CP, class FrP>
bool groth16_verify_impl(const VerifyInput& in) {
  using E = typename CP::E;
  using Fq = typename E::Fq;
  using Fq2 = typename E::Fq2;
  using G1 = typename E::G1;
  using G2 = typename E::G2;
  constexpr size_t lq = sizeof(Fq) / 8;
  auto g1 = [&](const uint64_t* p) { G1 r; memcpy(&r, p, sizeof(r)); return r; };
  auto g2 = [&](const uint64_t* p) { G2 r; memcpy(&r, p, sizeof(r)); return r; };
  if (in.n_ic == 0) throw Error("verify: the verification key has no IC points");
  const G1 A = g1(in.proof), C = g1(in.proof + 6 * lq), alpha = g1(in.alpha_g1);
  const G2 B = g2(in.proof + 2 * lq), beta = g2(in.beta_g2), gamma = g2(in.gamma_g2), delta = g2(in.delta_g2);
  const Fq b1 = CP::b1();
  const Fq2 b2 = CP::b2();
  for (const G1* p : {&A, &C, &alpha})
    if (!on_curve(*p, b1)) throw Error("verify: G1 point is not on the curve");
  for (const G2* p : {&B, &beta, &gamma, &delta}) {
    if (!on_curve(*p, b2)) throw Error("verify: G2 point is not on the curve");
    if (!scalar_mul_affine(*p, CP::order()).is_inf()) throw Error("verify: G2 point is not in the prime-order subgroup");
  }
  // the reference checks EVERY proof and vk point at parse time (circom-types/src/traits.rs:160-232): alpha and the IC points as well
  for (const G1* p : {&A, &C, &alpha})
    if (!scalar_mul_affine(*p, CP::order()).is_inf()) throw Error("verify: G1 point is not in the prime-order subgroup");
  // vk_x
  if (!on_curve(g1(in.ic), b1) || !scalar_mul_affine(g1(in.ic), CP::order()).is_inf()) throw Error("verify: IC point is not on the curve / in the subgroup");
  cocg::XYZZ<Fq> acc = cocg::xyzz_from_affine(g1(in.ic));
  for (size_t i = 1; i < in.n_ic; i++) {
    G1 ic = g1(in.ic + i * 2 * lq);
    if (!on_curve(ic, b1)) throw Error("verify: IC point is not on the curve");
    if (!scalar_mul_affine(ic, CP::order()).is_inf()) throw Error("verify: IC point is not in the prime-order subgroup");
    cocg::Fp<FrP> s;
    memcpy(s.l, in.pub + (i - 1) * 4, 32);
    s = cocg::fp_from_mont(s);
    cocg::xyzz_add(acc, scalar_mul_affine(ic, s.l));
  }
  const G1 vkx = xyzz_to_affine(acc);
  auto neg = [](const G1& p) { return G1{p.x, cocg::fp_neg(p.y)}; };
  typename E::F12 f = E::f12_one();
  const std::pair<G1, G2> pairs[4] = {{A, B}, {neg(alpha), beta}, {neg(vkx), gamma}, {neg(C), delta}};
  for (const auto& pr : pairs) {
    if (pr.first.is_inf() || pr.second.is_inf()) continue;  // e(O, Q) = e(P, O) = 1
    f = E::f12_mul(f, CP::miller(pr.first, pr.second));
  }
  return CP::final_is_one(f);
}

inline bool groth16_verify(int curve, const VerifyInput& in) {
  if (curve == COCG_BN254) return groth16_verify_impl<Bn254Pairing, cocg::Bn254FrP>(in);
  if (curve == COCG_BLS12_381) return groth16_verify_impl<Bls381Pairing, cocg::Bls381FrP>(in);
  throw Error("verify: unknown curve");
}

}  // namespace cohost
