// `co-circom verify` inputs (co-circom/src/bin/co-circom.rs:640-720 under /root/reference/co-circom): snarkjs verification_key.json
// (circom-types/src/groth16/verification_key.rs: vk_alpha_1, vk_beta_2, vk_gamma_2, vk_delta_2, IC, curve), proof.json
// (groth16/proof.rs:7-29) and public.json (array of decimal strings) -> the packed Montgomery blocks groth16_verify takes.
// A deliberately small JSON reader: objects, arrays, strings and bare numbers are all the three files contain.
#pragma once
#include <map>
#include <memory>

#include "pairing.hpp"

namespace cohost {

struct JsonValue {
  enum Kind { STRING, ARRAY, OBJECT } kind = STRING;
  std::string str;  // STRING (also bare numbers / literals, kept as text)
  std::vector<JsonValue> arr;
  std::vector<std::pair<std::string, JsonValue>> obj;
  const JsonValue& at(const std::string& key) const {
    if (kind != OBJECT) throw Error("json: expected an object");
    for (const auto& kv : obj)
      if (kv.first == key) return kv.second;
    throw Error("json: missing key \"" + key + "\"");
  }
  const JsonValue& at(size_t i) const {
    if (kind != ARRAY || i >= arr.size()) throw Error("json: expected an array of at least " + std::to_string(i + 1) + " elements");
    return arr[i];
  }
  const std::string& text() const {
    if (kind != STRING) throw Error("json: expected a string");
    return str;
  }
};

class JsonReader {
 public:
  JsonReader(const char* p, size_t n) : p_(p), end_(p + n) {}
  JsonValue parse() {
    JsonValue v = value(0);
    ws();
    if (p_ != end_) throw Error("json: trailing characters");
    return v;
  }

 private:
  const char *p_, *end_;
  void ws() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) p_++; }
  char peek() { ws(); if (p_ >= end_) throw Error("json: unexpected end"); return *p_; }
  std::string string() {
    if (peek() != '"') throw Error("json: expected a string");
    p_++;
    std::string s;
    while (p_ < end_ && *p_ != '"') {
      if (*p_ == '\\') {
        if (++p_ >= end_) break;
        switch (*p_) {
          case 'n': s.push_back('\n'); break;
          case 't': s.push_back('\t'); break;
          case 'r': s.push_back('\r'); break;
          case 'b': s.push_back('\b'); break;
          case 'f': s.push_back('\f'); break;
          case '"': case '\\': case '/': s.push_back(*p_); break;
          case 'u': throw Error("json: \\u escapes are not supported");
          default: throw Error("json: invalid escape");  // serde_json rejects anything else
        }
        p_++;
      } else {
        s.push_back(*p_++);
      }
    }
    if (p_ >= end_) throw Error("json: unterminated string");
    p_++;
    return s;
  }
  JsonValue value(int depth) {
    if (depth > 16) throw Error("json: nesting too deep");
    JsonValue v;
    char c = peek();
    if (c == '"') {
      v.str = string();
    } else if (c == '[') {
      v.kind = JsonValue::ARRAY;
      p_++;
      if (peek() == ']') { p_++; return v; }
      while (true) {
        v.arr.push_back(value(depth + 1));
        char d = peek();
        p_++;
        if (d == ']') break;
        if (d != ',') throw Error("json: expected ',' or ']'");
      }
    } else if (c == '{') {
      v.kind = JsonValue::OBJECT;
      p_++;
      if (peek() == '}') { p_++; return v; }
      while (true) {
        std::string k = string();
        if (peek() != ':') throw Error("json: expected ':'");
        p_++;
        for (const auto& kv : v.obj)
          if (kv.first == k) throw Error("json: duplicate key \"" + k + "\"");  // serde's derived Deserialize: `duplicate field`
        v.obj.emplace_back(std::move(k), value(depth + 1));
        char d = peek();
        p_++;
        if (d == '}') break;
        if (d != ',') throw Error("json: expected ',' or '}'");
      }
    } else {  // number / true / false / null: kept as text
      const char* s = p_;
      while (p_ < end_ && *p_ != ',' && *p_ != ']' && *p_ != '}' && *p_ != ' ' && *p_ != '\n' && *p_ != '\r' && *p_ != '\t') p_++;
      if (p_ == s) throw Error("json: unexpected character");
      v.str.assign(s, p_);
    }
    return v;
  }
};

// decimal string -> Montgomery limbs of field F; rejects anything that is not a canonical residue (ark's FromStr does the same)
template <class F>
F field_from_decimal(const std::string& s) {
  if (s.empty()) throw Error("field element: empty string");
  F v = F::zero();
  for (char ch : s) {
    if (ch < '0' || ch > '9') throw Error("field element: not a decimal number: " + s);
    uint64_t carry = (uint64_t)(ch - '0');
    for (int i = 0; i < F::N; i++) {
      uint64_t t = (uint64_t)v.l[i] * 10 + carry;
      v.l[i] = (uint32_t)t;
      carry = t >> 32;
    }
    if (carry) throw Error("field element: larger than the modulus");
  }
  for (int i = F::N - 1; i >= 0; i--) {
    uint32_t m = F::Params::mod(i);
    if (v.l[i] < m) break;
    if (v.l[i] > m || i == 0) throw Error("field element: larger than the modulus");
  }
  return cocg::fp_to_mont(v);
}

template <class CP, class FrP>
struct JsonVerifier {
  using E = typename CP::E;
  using Fq = typename E::Fq;
  using Fq2 = typename E::Fq2;
  static constexpr size_t lq = sizeof(Fq) / 8;

  // [x, y, z] Jacobian strings -> packed affine (traits.rs:160-184: Projective::new(x, y, z).into_affine())
  static void g1(const JsonValue& v, uint64_t* out) {
    Fq x = field_from_decimal<Fq>(v.at(0).text()), y = field_from_decimal<Fq>(v.at(1).text()), z = field_from_decimal<Fq>(v.at(2).text());
    memset(out, 0, 2 * lq * 8);
    if (z.is_zero()) return;
    // packed (0, 0) means infinity below this layer; with z != 0 it is the off-curve point (0, 0), which the reference rejects
    if (x.is_zero() && y.is_zero()) throw Error("verify: G1 point is not on the curve");
    if (!(z == Fq::one())) {
      Fq zi = cocg::fp_inv(z), zi2 = cocg::fp_sqr(zi);
      x = cocg::fp_mul(x, zi2);
      y = cocg::fp_mul(y, cocg::fp_mul(zi2, zi));
    }
    memcpy(out, x.l, lq * 8);
    memcpy(out + lq, y.l, lq * 8);
  }
  static Fq2 f2(const JsonValue& v) { return Fq2{field_from_decimal<Fq>(v.at(0).text()), field_from_decimal<Fq>(v.at(1).text())}; }
  static void g2(const JsonValue& v, uint64_t* out) {
    Fq2 x = f2(v.at(0)), y = f2(v.at(1)), z = f2(v.at(2));
    memset(out, 0, 4 * lq * 8);
    if (z.is_zero()) return;
    if (x.is_zero() && y.is_zero()) throw Error("verify: G2 point is not on the curve");
    if (!(z == Fq2::one())) {
      Fq2 zi = cocg::f_inv(z), zi2 = cocg::f_sqr(zi);
      x = cocg::f_mul(x, zi2);
      y = cocg::f_mul(y, cocg::f_mul(zi2, zi));
    }
    memcpy(out, x.c0.l, lq * 8);
    memcpy(out + lq, x.c1.l, lq * 8);
    memcpy(out + 2 * lq, y.c0.l, lq * 8);
    memcpy(out + 3 * lq, y.c1.l, lq * 8);
  }

  static bool verify(int curve, const JsonValue& vk, const JsonValue& proof, const JsonValue& pub) {
    std::vector<uint64_t> alpha(2 * lq), beta(4 * lq), gamma(4 * lq), delta(4 * lq), pr(8 * lq);
    g1(vk.at("vk_alpha_1"), alpha.data());
    g2(vk.at("vk_beta_2"), beta.data());
    g2(vk.at("vk_gamma_2"), gamma.data());
    g2(vk.at("vk_delta_2"), delta.data());
    const JsonValue& icv = vk.at("IC");
    if (icv.kind != JsonValue::ARRAY || pub.kind != JsonValue::ARRAY) throw Error("verify: IC and the public inputs must be arrays");
    // ark_groth16::prepare_inputs: public inputs + 1 must equal the number of IC points (SynthesisError::MalformedVerifyingKey)
    if (pub.arr.size() + 1 != icv.arr.size()) throw Error("verify: number of public inputs does not match the verification key");
    std::vector<uint64_t> ic(icv.arr.size() * 2 * lq), pubs(pub.arr.size() * 4 + 4);
    for (size_t i = 0; i < icv.arr.size(); i++) g1(icv.arr[i], ic.data() + i * 2 * lq);
    for (size_t i = 0; i < pub.arr.size(); i++) {
      cocg::Fp<FrP> s = field_from_decimal<cocg::Fp<FrP>>(pub.arr[i].text());
      memcpy(pubs.data() + 4 * i, s.l, 32);
    }
    g1(proof.at("pi_a"), pr.data());
    g2(proof.at("pi_b"), pr.data() + 2 * lq);
    g1(proof.at("pi_c"), pr.data() + 6 * lq);
    VerifyInput in{alpha.data(), beta.data(), gamma.data(), delta.data(), ic.data(), icv.arr.size(), pr.data(), pubs.data()};
    return groth16_verify(curve, in);
  }
};

// curve names as the reference writes them (circom-types/src/traits.rs:18-31)
inline int curve_from_name(const std::string& s) {
  if (s == "bn128") return COCG_BN254;
  if (s == "bls12381") return COCG_BLS12_381;
  throw Error("verify: unknown curve \"" + s + "\"");
}

inline bool groth16_verify_json(const char* vk, size_t vk_len, const char* proof, size_t proof_len, const char* pub, size_t pub_len) {
  JsonValue v = JsonReader(vk, vk_len).parse(), p = JsonReader(proof, proof_len).parse(), u = JsonReader(pub, pub_len).parse();
  const int curve = curve_from_name(v.at("curve").text());
  if (curve_from_name(p.at("curve").text()) != curve) throw Error("verify: proof and verification key are over different curves");
  if (p.at("protocol").text() != "groth16" || v.at("protocol").text() != "groth16") throw Error("verify: not a groth16 proof / key");
  if (curve == COCG_BN254) return JsonVerifier<Bn254Pairing, cocg::Bn254FrP>::verify(curve, v, p, u);
  return JsonVerifier<Bls381Pairing, cocg::Bls381FrP>::verify(curve, v, p, u);
}

}  // namespace cohost
