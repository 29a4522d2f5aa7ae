// Plonk verification on the host -- `Plonk::verify` of the reference (/root/reference/co-circom/co-plonk/src/plonk.rs:123-283) and the
// verifier half of `co-circom verify plonk` (co-circom/src/bin/co-circom.rs:696-712): Keccak-256 transcript (types.rs:125-176), the six
// Fiat-Shamir challenges (plonk.rs:44-121), Lagrange evaluations and PI (lib.rs:160-199), r0 / D / E / F and the final pairing check
//     e(Wxi + u Wxiw, X_2) == e(xi Wxi + u xi w Wxiw - E + F, G2)
// on the pairing of pairing.hpp.  The transcript and challenge code is what the prover rounds 2-5 will share.
// O(1) work per proof, no GPU.
#pragma once
#include <array>

#include "groth16.hpp"
#include "verify_json.hpp"

namespace cohost {

// ---- Keccak-256 (original padding 0x01, as sha3::Keccak256 in the reference; not NIST SHA3-256) ------------------------------------
inline void keccak_f1600(uint64_t a[25]) {
  static const uint64_t rc[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
                                  0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
                                  0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
                                  0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
                                  0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int rot[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};  // index x + 5 y
  auto rol = [](uint64_t v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; };
  for (int round = 0; round < 24; round++) {
    uint64_t c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], rot[x + 5 * y]);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= rc[round];
  }
}
inline std::array<uint8_t, 32> keccak256(const uint8_t* data, size_t n) {
  constexpr size_t rate = 136;
  std::vector<uint8_t> msg(data, data + n);
  msg.push_back(0x01);
  while (msg.size() % rate) msg.push_back(0);
  msg.back() |= 0x80;
  uint64_t st[25] = {};
  for (size_t off = 0; off < msg.size(); off += rate) {
    for (size_t i = 0; i < rate / 8; i++) {
      uint64_t w = 0;
      for (int k = 7; k >= 0; k--) w = (w << 8) | msg[off + 8 * i + k];
      st[i] ^= w;
    }
    keccak_f1600(st);
  }
  std::array<uint8_t, 32> out;
  for (int i = 0; i < 32; i++) out[i] = (uint8_t)(st[i / 8] >> (8 * (i % 8)));
  return out;
}

template <class CP, class FrP>
struct PlonkVerifier {
  using E = typename CP::E;
  using Fq = typename E::Fq;
  using Fr = cocg::Fp<FrP>;
  using G1 = typename E::G1;
  using G2 = typename E::G2;
  using X1 = cocg::XYZZ<Fq>;
  static constexpr size_t lq = sizeof(Fq) / 8;

  // Keccak256Transcript (types.rs:125-176): canonical big-endian bytes; the point at infinity is 2 x byte_len zero bytes
  struct Transcript {
    std::vector<uint8_t> buf;
    template <class F>
    void add_field(const F& mont, size_t bytes) {
      F c = cocg::fp_from_mont(mont);
      for (size_t i = 0; i < bytes; i++) {
        size_t k = bytes - 1 - i;  // byte k of the little-endian value
        buf.push_back((uint8_t)(c.l[k / 4] >> (8 * (k % 4))));
      }
    }
    void add_scalar(const Fr& s) { add_field(s, 32); }
    void add_point(const G1& p) {
      const size_t qlen = (Fq::Params::BITS + 7) / 8;
      if (p.is_inf()) { buf.insert(buf.end(), 2 * qlen, 0); return; }
      add_field(p.x, qlen);
      add_field(p.y, qlen);
    }
    Fr get_challenge() const {  // from_be_bytes_mod_order of the 32-byte digest
      auto h = keccak256(buf.data(), buf.size());
      Fr v;
      for (int i = 0; i < 8; i++) v.l[i] = ((uint32_t)h[31 - 4 * i]) | ((uint32_t)h[30 - 4 * i] << 8) | ((uint32_t)h[29 - 4 * i] << 16) | ((uint32_t)h[28 - 4 * i] << 24);
      while (true) {  // v < 2^256 < 6 r: subtract the modulus until canonical
        bool ge = true;
        for (int i = 7; i >= 0; i--) {
          if (v.l[i] > FrP::mod(i)) break;
          if (v.l[i] < FrP::mod(i)) { ge = false; break; }
        }
        if (!ge) break;
        uint64_t br = 0;
        for (int i = 0; i < 8; i++) {
          uint64_t dlt = (uint64_t)v.l[i] - FrP::mod(i) - br;
          v.l[i] = (uint32_t)dlt;
          br = (dlt >> 63) & 1;
        }
      }
      return cocg::fp_to_mont(v);
    }
  };

  struct Proof {
    G1 a, b, c, z, t1, t2, t3, wxi, wxiw;
    Fr eval_a, eval_b, eval_c, eval_s1, eval_s2, eval_zw;
  };
  struct Vk {
    size_t n_public = 0, power = 0;
    Fr k1, k2;
    G1 qm, ql, qr, qo, qc, s1, s2, s3;
    G2 x2;
  };
  struct Challenges {
    Fr alpha, beta, gamma, xi, v[5], u;
  };

  static Challenges challenges(const Vk& vk, const Proof& p, const std::vector<Fr>& pub) {  // plonk.rs:44-121
    Challenges ch;
    Transcript t;
    for (const G1* q : {&vk.qm, &vk.ql, &vk.qr, &vk.qo, &vk.qc, &vk.s1, &vk.s2, &vk.s3}) t.add_point(*q);
    for (const Fr& s : pub) t.add_scalar(s);
    t.add_point(p.a); t.add_point(p.b); t.add_point(p.c);
    ch.beta = t.get_challenge();
    t = Transcript();
    t.add_scalar(ch.beta);
    ch.gamma = t.get_challenge();
    t = Transcript();
    t.add_scalar(ch.beta); t.add_scalar(ch.gamma); t.add_point(p.z);
    ch.alpha = t.get_challenge();
    t = Transcript();
    t.add_scalar(ch.alpha); t.add_point(p.t1); t.add_point(p.t2); t.add_point(p.t3);
    ch.xi = t.get_challenge();
    t = Transcript();
    t.add_scalar(ch.xi);
    for (const Fr* s : {&p.eval_a, &p.eval_b, &p.eval_c, &p.eval_s1, &p.eval_s2, &p.eval_zw}) t.add_scalar(*s);
    ch.v[0] = t.get_challenge();
    for (int i = 1; i < 5; i++) ch.v[i] = cocg::fp_mul(ch.v[i - 1], ch.v[0]);
    t = Transcript();
    t.add_point(p.wxi); t.add_point(p.wxiw);
    ch.u = t.get_challenge();
    return ch;
  }

  static X1 mul(const G1& p, const Fr& k_mont) {
    Fr k = cocg::fp_from_mont(k_mont);
    return scalar_mul_affine(p, k.l);
  }
  static X1 neg(X1 p) { p.y = cocg::fp_neg(p.y); return p; }

  static bool verify(int curve, const Vk& vk, const Proof& p, const std::vector<Fr>& pub) {
    using namespace cocg;
    if (vk.n_public != pub.size()) throw Error("Invalid number of public inputs");
    const Fq b1c = CP::b1();
    for (const G1* q : {&p.a, &p.b, &p.c, &p.z, &p.t1, &p.t2, &p.t3, &p.wxi, &p.wxiw, &vk.qm, &vk.ql, &vk.qr, &vk.qo, &vk.qc, &vk.s1, &vk.s2, &vk.s3})
    {
      if (!on_curve(*q, b1c)) throw Error("verify: G1 point is not on the curve");
      // is_in_correct_subgroup_assuming_on_curve at deserialisation (circom-types/src/traits.rs:160-232); G1 of BLS12-381 has a cofactor
      if (!scalar_mul_affine(*q, CP::order()).is_inf()) throw Error("verify: G1 point is not in the prime-order subgroup");
    }
    if (!on_curve(vk.x2, CP::b2())) throw Error("verify: G2 point is not on the curve");
    if (!scalar_mul_affine(vk.x2, CP::order()).is_inf()) throw Error("verify: G2 point is not in the prime-order subgroup");
    const Challenges ch = challenges(vk, p, pub);
    // Domains::new: root_of_unity_pow = roots_of_unity[power] (types.rs:59-99)
    Fr omega;
    {
      cohost::Fr w = root_of_unity_for_groth16(curve, vk.power).omega;
      memcpy(omega.l, w.l, 32);
    }
    // calculate_lagrange_evaluations + calculate_pi (lib.rs:160-199)
    Fr xin = ch.xi, one = Fr::one(), nfr = one;
    for (size_t i = 0; i < vk.power; i++) { xin = fp_sqr(xin); nfr = fp_add(nfr, nfr); }
    const Fr zh = fp_sub(xin, one);
    const size_t l_len = vk.n_public > 1 ? vk.n_public : 1;
    std::vector<Fr> l(l_len);
    Fr w = one;
    for (size_t i = 0; i < l_len; i++) {
      l[i] = fp_mul(fp_mul(w, zh), fp_inv(fp_mul(nfr, fp_sub(ch.xi, w))));
      w = fp_mul(w, omega);
    }
    Fr pi = Fr::zero();
    for (size_t i = 0; i < pub.size(); i++) pi = fp_sub(pi, fp_mul(l[i], pub[i]));
    // calculate_r0_d (plonk.rs:170-226)
    const Fr e2 = fp_mul(fp_sqr(ch.alpha), l[0]);
    const Fr e3a = fp_add(fp_add(p.eval_a, fp_mul(p.eval_s1, ch.beta)), ch.gamma);
    const Fr e3b = fp_add(fp_add(p.eval_b, fp_mul(p.eval_s2, ch.beta)), ch.gamma);
    const Fr e3c = fp_add(p.eval_c, ch.gamma);
    const Fr e3 = fp_mul(fp_mul(fp_mul(fp_mul(e3a, e3b), e3c), p.eval_zw), ch.alpha);
    const Fr r0 = fp_sub(fp_sub(pi, e2), e3);
    X1 d = mul(vk.qm, fp_mul(p.eval_a, p.eval_b));
    xyzz_add(d, mul(vk.ql, p.eval_a));
    xyzz_add(d, mul(vk.qr, p.eval_b));
    xyzz_add(d, mul(vk.qo, p.eval_c));
    xyzz_madd(d, vk.qc);
    const Fr betaxi = fp_mul(ch.beta, ch.xi);
    const Fr d2a1 = fp_add(fp_add(p.eval_a, betaxi), ch.gamma);
    const Fr d2a2 = fp_add(fp_add(p.eval_b, fp_mul(betaxi, vk.k1)), ch.gamma);
    const Fr d2a3 = fp_add(fp_add(p.eval_c, fp_mul(betaxi, vk.k2)), ch.gamma);
    const Fr d2a = fp_mul(fp_mul(fp_mul(d2a1, d2a2), d2a3), ch.alpha);
    xyzz_add(d, mul(p.z, fp_add(fp_add(d2a, e2), ch.u)));
    const Fr d3c = fp_mul(fp_mul(ch.alpha, ch.beta), p.eval_zw);
    xyzz_add(d, neg(mul(vk.s3, fp_mul(fp_mul(e3a, e3b), d3c))));
    X1 d4 = xyzz_from_affine(p.t1);
    xyzz_add(d4, mul(p.t2, xin));
    xyzz_add(d4, mul(p.t3, fp_sqr(xin)));
    xyzz_add(d, neg(mul(xyzz_to_affine(d4), zh)));
    // calculate_e, calculate_f (plonk.rs:228-257)
    Fr e = fp_mul(ch.v[0], p.eval_a);
    e = fp_add(e, fp_mul(ch.v[1], p.eval_b));
    e = fp_add(e, fp_mul(ch.v[2], p.eval_c));
    e = fp_add(e, fp_mul(ch.v[3], p.eval_s1));
    e = fp_add(e, fp_mul(ch.v[4], p.eval_s2));
    e = fp_add(e, fp_mul(ch.u, p.eval_zw));
    e = fp_sub(e, r0);
    X1 f = d;
    xyzz_add(f, mul(p.a, ch.v[0]));
    xyzz_add(f, mul(p.b, ch.v[1]));
    xyzz_add(f, mul(p.c, ch.v[2]));
    xyzz_add(f, mul(vk.s1, ch.v[3]));
    xyzz_add(f, mul(vk.s2, ch.v[4]));
    // valid_pairing (plonk.rs:259-282)
    const Fr s = fp_mul(fp_mul(ch.u, ch.xi), omega);
    X1 a1 = xyzz_from_affine(p.wxi);
    xyzz_add(a1, mul(p.wxiw, ch.u));
    X1 bb = mul(p.wxi, ch.xi);
    xyzz_add(bb, mul(p.wxiw, s));
    xyzz_add(bb, neg(mul(CP::g1_generator(), e)));
    xyzz_add(bb, f);
    const G1 a1a = xyzz_to_affine(a1), nb = xyzz_to_affine(neg(bb));
    typename E::F12 acc = E::f12_one();
    if (!a1a.is_inf() && !vk.x2.is_inf()) acc = E::f12_mul(acc, CP::miller(a1a, vk.x2));
    if (!nb.is_inf()) acc = E::f12_mul(acc, CP::miller(nb, CP::g2_generator()));
    return CP::final_is_one(acc);
  }

  // ---- JSON front end (circom-types/src/plonk/{verification_key,proof}.rs) -------------------------------------------------------
  using JV = JsonVerifier<CP, FrP>;
  static G1 g1(const JsonValue& v) {
    uint64_t buf[2 * lq];
    JV::g1(v, buf);
    G1 r;
    memcpy(&r, buf, sizeof(r));
    return r;
  }
  static Fr fr(const JsonValue& v) { return field_from_decimal<Fr>(v.text()); }
  static size_t count(const JsonValue& v) {
    const std::string& s = v.text();
    if (s.empty() || s.size() > 9) throw Error("json: expected a small non-negative integer");
    size_t n = 0;
    for (char c : s) {
      if (c < '0' || c > '9') throw Error("json: expected a small non-negative integer");
      n = n * 10 + (size_t)(c - '0');
    }
    return n;
  }
  static void parse(const JsonValue& vkj, const JsonValue& pj, const JsonValue& pubj, Vk& vk, Proof& p, std::vector<Fr>& pub) {
    vk.n_public = count(vkj.at("nPublic"));
    vk.power = count(vkj.at("power"));
    if (vk.power > 28) throw Error("verify: domain too large");
    vk.k1 = fr(vkj.at("k1")); vk.k2 = fr(vkj.at("k2"));
    vk.qm = g1(vkj.at("Qm")); vk.ql = g1(vkj.at("Ql")); vk.qr = g1(vkj.at("Qr")); vk.qo = g1(vkj.at("Qo")); vk.qc = g1(vkj.at("Qc"));
    vk.s1 = g1(vkj.at("S1")); vk.s2 = g1(vkj.at("S2")); vk.s3 = g1(vkj.at("S3"));
    {
      uint64_t buf[4 * lq];
      JV::g2(vkj.at("X_2"), buf);
      memcpy(&vk.x2, buf, sizeof(vk.x2));
    }
    p.a = g1(pj.at("A")); p.b = g1(pj.at("B")); p.c = g1(pj.at("C")); p.z = g1(pj.at("Z"));
    p.t1 = g1(pj.at("T1")); p.t2 = g1(pj.at("T2")); p.t3 = g1(pj.at("T3")); p.wxi = g1(pj.at("Wxi")); p.wxiw = g1(pj.at("Wxiw"));
    p.eval_a = fr(pj.at("eval_a")); p.eval_b = fr(pj.at("eval_b")); p.eval_c = fr(pj.at("eval_c"));
    p.eval_s1 = fr(pj.at("eval_s1")); p.eval_s2 = fr(pj.at("eval_s2")); p.eval_zw = fr(pj.at("eval_zw"));
    if (pubj.kind != JsonValue::ARRAY) throw Error("verify: the public inputs must be an array");
    for (const auto& v : pubj.arr) pub.push_back(fr(v));
  }
  // challenges_out: NULL or 6 x 4 limbs -- alpha, beta, gamma, xi, v[0], u (Montgomery), the values of the reference's challenge KAT
  static bool verify_json(int curve, const JsonValue& vkj, const JsonValue& pj, const JsonValue& pubj, uint64_t* challenges_out) {
    Vk vk;
    Proof p;
    std::vector<Fr> pub;
    parse(vkj, pj, pubj, vk, p, pub);
    if (challenges_out) {
      const Challenges ch = challenges(vk, p, pub);
      const Fr* v[6] = {&ch.alpha, &ch.beta, &ch.gamma, &ch.xi, &ch.v[0], &ch.u};
      for (int i = 0; i < 6; i++) memcpy(challenges_out + 4 * i, v[i]->l, 32);
    }
    return verify(curve, vk, p, pub);
  }
};

inline bool plonk_verify_json(const char* vk, size_t vk_len, const char* proof, size_t proof_len, const char* pub, size_t pub_len,
                              uint64_t* challenges_out /* NULL or 6 x 4 limbs */) {
  JsonValue v = JsonReader(vk, vk_len).parse(), p = JsonReader(proof, proof_len).parse(), u = JsonReader(pub, pub_len).parse();
  const int curve = curve_from_name(v.at("curve").text());
  if (curve_from_name(p.at("curve").text()) != curve) throw Error("verify: proof and verification key are over different curves");
  if (p.at("protocol").text() != "plonk" || v.at("protocol").text() != "plonk") throw Error("verify: not a plonk proof / key");
  if (curve == COCG_BN254) return PlonkVerifier<Bn254Pairing, cocg::Bn254FrP>::verify_json(curve, v, p, u, challenges_out);
  return PlonkVerifier<Bls381Pairing, cocg::Bls381FrP>::verify_json(curve, v, p, u, challenges_out);
}

}  // namespace cohost
