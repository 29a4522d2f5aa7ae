// Output-side data formats of the proving path: what `co-circom generate-proof` / `split-witness` write and read
// around CoGroth16::prove.  Host-only code (no GPU needed), the values cross the C ABI in the same Montgomery limbs as everywhere.
//   proof JSON            Groth16Proof<P> via serde_json (circom-types/src/groth16/proof.rs:7-29); points as snarkjs does:
//                         G1 [x, y, "1"] (infinity ["0","1","0"], traits.rs:186-193), G2 [[x0,x1],[y0,y1],["1","0"]] (traits.rs:225-232),
//                         "protocol":"groth16", "curve": "bn128" | "bls12381" (traits.rs:18-31); written compact (serde_json::to_writer,
//                         co-circom/src/bin/co-circom.rs:540)
//   public inputs JSON    array of decimal strings without the leading constant 1 (co-circom.rs:611-629)
//   SharedWitness file    bincode 1.x of { public_inputs: bytes, witness: bytes } (co-circom-snarks/src/lib.rs:24-41), each `bytes` =
//                         u64 length + the ark-serialize 0.4 compressed encoding (serde_compat.rs:5-24): Vec<F> = u64 count + 32-byte
//                         little-endian canonical elements; Rep3PrimeFieldShareVec = a then b (rep3/fieldshare.rs:232-236),
//                         ShamirPrimeFieldShareVec = a (shamir/fieldshare.rs:152-155).  The reference ships no .shared fixture, so
//                         this layout is restated from the two crates' published formats (bincode default options: fixed-width
//                         little-endian integers), not pinned by a golden file.
// (paths under /root/reference/co-circom unless noted)
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "driver.hpp"
#include "formats.hpp"

namespace cohost {

// decimal digits of a little-endian integer of n 32-bit limbs
inline std::string limbs_to_decimal(const uint32_t* limbs, int n) {
  std::vector<uint32_t> v(limbs, limbs + n);
  std::string out;
  while (true) {
    bool zero = true;
    for (uint32_t x : v) zero = zero && x == 0;
    if (zero) break;
    uint64_t rem = 0;
    for (int i = n - 1; i >= 0; i--) {  // divide by 10^9
      uint64_t cur = (rem << 32) | v[i];
      v[i] = (uint32_t)(cur / 1000000000u);
      rem = cur % 1000000000u;
    }
    bool last = true;
    for (uint32_t x : v) last = last && x == 0;
    for (int k = 0; k < 9 && (!last || rem); k++) {
      out.push_back((char)('0' + rem % 10));
      rem /= 10;
    }
  }
  if (out.empty()) out = "0";
  std::reverse(out.begin(), out.end());
  return out;
}

// Montgomery limbs -> decimal string, for the scalar (is_fr) or base field of `curve`
inline std::string field_to_decimal(int curve, bool is_fr, const uint64_t* mont) {
  auto conv = [&](auto tag) {
    using F = decltype(tag);
    F x;
    memcpy(x.l, mont, sizeof(x.l));
    F c = cocg::fp_from_mont(x);
    return limbs_to_decimal(c.l, F::N);
  };
  if (curve == COCG_BN254) return is_fr ? conv(cocg::Bn254Fr()) : conv(cocg::Bn254Fq());
  return is_fr ? conv(cocg::Bls381Fr()) : conv(cocg::Bls381Fq());
}

inline bool limbs_zero(const uint64_t* p, size_t n) {
  for (size_t i = 0; i < n; i++)
    if (p[i]) return false;
  return true;
}

// proof: A | B | C packed affine, Montgomery (what cohost_*_prove writes)
inline std::string proof_to_json(int curve, const uint64_t* proof) {
  const size_t lq = curve == COCG_BN254 ? 4 : 6;
  auto q = [&](const uint64_t* p) { return "\"" + field_to_decimal(curve, false, p) + "\""; };
  auto g1 = [&](const uint64_t* p) {
    if (limbs_zero(p, 2 * lq)) return std::string("[\"0\",\"1\",\"0\"]");
    return "[" + q(p) + "," + q(p + lq) + ",\"1\"]";
  };
  const uint64_t *a = proof, *b = proof + 2 * lq, *c = proof + 6 * lq;
  if (limbs_zero(b, 4 * lq)) throw Error("proof_to_json: pi_b is the point at infinity");  // serialize_g2 unwraps xy()
  std::string s = "{\"pi_a\":" + g1(a);
  s += ",\"pi_b\":[[" + q(b) + "," + q(b + lq) + "],[" + q(b + 2 * lq) + "," + q(b + 3 * lq) + "],[\"1\",\"0\"]]";
  s += ",\"pi_c\":" + g1(c);
  s += std::string(",\"protocol\":\"groth16\",\"curve\":\"") + (curve == COCG_BN254 ? "bn128" : "bls12381") + "\"}";
  return s;
}

// PlonkProof via serde_json (circom-types/src/plonk/proof.rs:7-87): A B C Z T1 T2 T3 Wxi Wxiw as G1 [x, y, "1"], then eval_a eval_b eval_c
// eval_s1 eval_s2 eval_zw as decimal strings, "protocol":"plonk", "curve".  proof: 9 packed affine points | 6 Fr (what cohost_plonk_prove writes)
inline std::string plonk_proof_to_json(int curve, const uint64_t* proof) {
  const size_t lq = curve == COCG_BN254 ? 4 : 6;
  auto q = [&](const uint64_t* p) { return "\"" + field_to_decimal(curve, false, p) + "\""; };
  auto g1 = [&](const uint64_t* p) {
    if (limbs_zero(p, 2 * lq)) return std::string("[\"0\",\"1\",\"0\"]");
    return "[" + q(p) + "," + q(p + lq) + ",\"1\"]";
  };
  static const char* pn[9] = {"A", "B", "C", "Z", "T1", "T2", "T3", "Wxi", "Wxiw"};
  static const char* en[6] = {"eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw"};
  std::string s = "{";
  for (int i = 0; i < 9; i++) s += std::string(i ? "," : "") + "\"" + pn[i] + "\":" + g1(proof + (size_t)i * 2 * lq);
  for (int i = 0; i < 6; i++) s += std::string(",\"") + en[i] + "\":\"" + field_to_decimal(curve, true, proof + 18 * lq + (size_t)i * 4) + "\"";
  s += std::string(",\"protocol\":\"plonk\",\"curve\":\"") + (curve == COCG_BN254 ? "bn128" : "bls12381") + "\"}";
  return s;
}

// pub: n_public + 1 Montgomery Fr with the constant 1 in front (SharedWitness::public_inputs); the 1 is skipped
inline std::string public_inputs_to_json(int curve, const uint64_t* pub, size_t count) {
  std::string s = "[";
  for (size_t i = 1; i < count; i++) {
    if (i > 1) s += ",";
    s += "\"" + field_to_decimal(curve, true, pub + 4 * i) + "\"";
  }
  return s + "]";
}

// ---------------------------------------------------------------- SharedWitness files
namespace detail {
inline void put_u64(std::vector<uint8_t>& o, uint64_t v) {
  for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i)));
}
inline void put_vec(std::vector<uint8_t>& o, const FrOps& fr, const uint64_t* mont, size_t n) {  // ark Vec<F>, compressed
  put_u64(o, n);
  size_t at = o.size();
  o.resize(at + n * 32);
  for (size_t i = 0; i < n; i++) {
    Fr x;
    memcpy(x.l, mont + 4 * i, 32);
    Fr c = fr.from_mont(x);
    memcpy(o.data() + at + i * 32, c.l, 32);  // little-endian host
  }
}
struct Reader {
  const uint8_t* p;
  size_t left;
  uint64_t u64() {
    if (left < 8) throw Error("shared witness: truncated");
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
    p += 8;
    left -= 8;
    return v;
  }
  // ark Vec<F> -> Montgomery limbs; rejects non-canonical elements like Validate::Yes
  std::vector<Fr> vec(const FrOps& fr, const uint8_t* modulus) {
    uint64_t n = u64();
    if (n > left / 32) throw Error("shared witness: truncated");
    std::vector<Fr> out(n);
    Fr r2 = fr.un(fr.zero(), [](auto x) { return decltype(x)::r2(); });
    for (uint64_t i = 0; i < n; i++) {
      for (int k = 31; k >= 0; k--) {
        if (p[k] < modulus[k]) break;
        if (p[k] > modulus[k] || k == 0) throw Error("shared witness: field element is not canonical");
      }
      Fr c;
      memcpy(c.l, p, 32);
      out[i] = fr.mul(c, r2);
      p += 32;
      left -= 32;
    }
    return out;
  }
};
}  // namespace detail

// comps: 2 share components (REP3: a, b) or 1 (Shamir), n Montgomery Fr each; pub: n_pub Montgomery Fr (leading 1 included)
inline std::vector<uint8_t> shared_witness_encode(int curve, const uint64_t* pub, size_t n_pub, const uint64_t* const* comps, int k, size_t n) {
  FrOps fr{curve};
  std::vector<uint8_t> o;
  o.reserve(32 + 32 * (n_pub + k * n));
  detail::put_u64(o, 8 + 32 * n_pub);
  detail::put_vec(o, fr, pub, n_pub);
  detail::put_u64(o, (uint64_t)k * (8 + 32 * n));
  for (int j = 0; j < k; j++) detail::put_vec(o, fr, comps[j], n);
  return o;
}
struct SharedWitnessData {
  std::vector<Fr> public_inputs;
  std::vector<Fr> comps[2];
};
inline SharedWitnessData shared_witness_decode(int curve, const uint8_t* data, size_t len, int k) {
  FrOps fr{curve};
  const uint8_t* modulus = curve == COCG_BN254 ? kBn254R : kBls381R;
  detail::Reader rd{data, len};
  SharedWitnessData w;
  uint64_t l1 = rd.u64();
  if (l1 > rd.left) throw Error("shared witness: truncated");
  {
    detail::Reader sub{rd.p, (size_t)l1};
    w.public_inputs = sub.vec(fr, modulus);
    if (sub.left) throw Error("shared witness: trailing bytes after the public inputs");
  }
  rd.p += l1;
  rd.left -= l1;
  uint64_t l2 = rd.u64();
  if (l2 > rd.left) throw Error("shared witness: truncated");
  detail::Reader sub{rd.p, (size_t)l2};
  for (int j = 0; j < k; j++) w.comps[j] = sub.vec(fr, modulus);
  if (sub.left) throw Error("shared witness: trailing bytes after the witness share (wrong protocol?)");
  if (k == 2 && w.comps[0].size() != w.comps[1].size()) throw Error("shared witness: share components differ in length");
  return w;
}

}  // namespace cohost
