// Round 1 of the collaborative Plonk prover on the cocg kernels, generic over the MPC driver -- the first of the five rounds of
// CoPlonk::prove (/root/reference/co-circom/co-plonk/src/lib.rs:80-99), i.e. /root/reference/co-circom/co-plonk/src/round1.rs:
//   Round1Challenges::{random, deterministic}   :86-108   (11 blinders; b_i = i in the reference's KAT configuration)
//   calculate_additions                         :213-242  + plonk_utils::get_witness (lib.rs:113-137)
//   compute_wire_polynomials                    :121-209  (3 iFFT of n, 3 FFT of 4n, blind_coefficients lib.rs:140-158)
//   round1                                      :263-300  (3 MSMs over p_tau, open_point_many)
// plus the Plonk zkey reader for what this round consumes (circom-types/src/plonk/zkey.rs:160-330: header, additions, wire maps,
// p_tau).  Rounds 2-5 are not built.  The wire buffers are gathered on the host from the party's witness share
// (index rules of get_witness), everything from the iFFT on runs on the GPU.
#pragma once
#include "formats.hpp"
#include "groth16.hpp"

namespace cohost {

struct PlonkZKeyFile {
  int curve = 0;
  size_t n_vars = 0, n_public = 0, domain_size = 0, pow = 0, n_additions = 0, n_constraints = 0, n8q = 0;
  const uint8_t *additions = nullptr, *map_a = nullptr, *map_b = nullptr, *map_c = nullptr, *p_tau = nullptr;
  const uint8_t *k1 = nullptr, *k2 = nullptr, *vk_g1 = nullptr, *x_2 = nullptr;  // header tail; nullptr when the header stops early
  // selector / sigma / Lagrange sections (n coefficients + 4n extended evaluations per polynomial, Montgomery Fr), when present
  const uint8_t *sel[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *sigma = nullptr, *lagrange = nullptr;
  PlonkZKeyFile(const uint8_t* data, size_t len) {
    BinFile bf(data, len, "zkey");
    auto s1 = bf.take(1);
    if (s1.second < 4 || BinFile::u32(s1.first) != 2) throw Error("zkey: not a plonk key (protocol id != 2)");
    auto h = bf.take(2);
    const uint8_t* p = h.first;
    const uint8_t* end = h.first + h.second;
    auto need = [&](size_t k) { if ((size_t)(end - p) < k) throw Error("zkey: truncated header"); };
    need(4);
    n8q = BinFile::u32(p); p += 4;
    if (n8q != 32 && n8q != 48) throw Error("zkey: unexpected base field byte size");
    curve = n8q == 32 ? COCG_BN254 : COCG_BLS12_381;
    need(n8q);
    if (memcmp(p, curve == COCG_BN254 ? kBn254Q : kBls381Q, n8q) != 0) throw Error("zkey: invalid prime in header");
    p += n8q;
    need(4 + 32);
    if (BinFile::u32(p) != 32 || memcmp(p + 4, curve == COCG_BN254 ? kBn254R : kBls381R, 32) != 0) throw Error("zkey: invalid prime in header");
    p += 36;
    need(20);
    n_vars = BinFile::u32(p); n_public = BinFile::u32(p + 4); domain_size = BinFile::u32(p + 8);
    n_additions = BinFile::u32(p + 12); n_constraints = BinFile::u32(p + 16);
    p += 20;
    // VerifyingKey::new (plonk/zkey.rs:329-355): k1, k2 (Montgomery Fr), Qm Ql Qr Qo Qc S1 S2 S3 (G1, Montgomery), X_2 (G2) -- rounds 2-5
    if ((size_t)(end - p) >= 64 + 8 * 2 * n8q + 4 * n8q) {
      k1 = p; k2 = p + 32;
      vk_g1 = p + 64;
      x_2 = vk_g1 + 8 * 2 * n8q;
    }
    if (domain_size == 0 || (domain_size & (domain_size - 1))) throw Error("zkey: invalid domain size");  // PlonkProofError::InvalidDomainSize
    while (((size_t)1 << pow) < domain_size) pow++;
    if (n_constraints > domain_size || n_vars < n_additions + n_public + 1) throw Error("zkey: inconsistent header");
    auto sec = [&](uint32_t id, size_t bytes, const char* what) {
      auto s = bf.take(id);
      if (s.second < bytes) throw Error(std::string("zkey: section too short: ") + what);
      return s.first;
    };
    additions = sec(3, n_additions * (8 + 64), "additions");
    map_a = sec(4, n_constraints * 4, "A map");
    map_b = sec(5, n_constraints * 4, "B map");
    map_c = sec(6, n_constraints * 4, "C map");
    p_tau = sec(14, (domain_size + 6) * 2 * n8q, "p_tau");
    const size_t poly_bytes = 5 * domain_size * 32;
    auto opt = [&](uint32_t id, size_t bytes) -> const uint8_t* {
      auto it = bf.sections.find(id);
      return it != bf.sections.end() && it->second.second >= bytes ? it->second.first : nullptr;
    };
    for (int k = 0; k < 5; k++) sel[k] = opt(7 + k, poly_bytes);               // Qm, Ql, Qr, Qo, Qc  (zkey.rs:236-249)
    sigma = opt(12, 3 * poly_bytes);
    lagrange = opt(13, (n_public > 1 ? n_public : 1) * poly_bytes);
  }
};

struct PlonkZKey {  // what round 1 needs, p_tau resident in HBM
  int curve = 0, device = 0;
  cocg_ctx* owner = nullptr;
  size_t n_vars = 0, n_public = 0, domain_size = 0, pow = 0, n_additions = 0, n_constraints = 0;
  struct Addition { uint32_t s1, s2; Fr f1, f2; };  // factors in Montgomery form (montgomery_bigint_from_reader)
  std::vector<Addition> additions;
  std::vector<uint32_t> map_a, map_b, map_c;
  uint64_t p_tau = 0;
};

struct Round1Proof {
  Point commit_a, commit_b, commit_c;  // packed affine
};

// host-side share arithmetic the round needs, per driver kind
template <class T> struct ShareOps;
template <> struct ShareOps<PlainDriver> {
  static FieldShare promote(PlainDriver& d, const Fr& v) { return FieldShare{v, d.fr.zero()}; }
};
template <> struct ShareOps<Rep3Protocol> {
  static FieldShare promote(Rep3Protocol& d, const Fr& v) {  // fieldshare.rs:55-79: (v, 0) / (0, v) / (0, 0)
    return FieldShare{d.id() == 0 ? v : d.fr.zero(), d.id() == 1 ? v : d.fr.zero()};
  }
};

template <class T>
class CoPlonkRound1 {
 public:
  explicit CoPlonkRound1(T& driver) : driver(driver) {}
  T& driver;

  // public_inputs: n_public + 1 values (the leading one is overwritten with zero, PlonkWitness::new types.rs:105-108);
  // wit_a / wit_b: the party's share components of values[n_public + 1 ..] on the HOST (wit_b unused by the plain driver)
  Round1Proof round1(const PlonkZKey& zk, uint64_t p_tau_handle, const Fr* public_inputs, const Fr* wit_a, const Fr* wit_b, bool deterministic) {
    const FrOps& fr = driver.fr;
    const size_t n = zk.domain_size, base = zk.n_vars - zk.n_additions;
    // ---- witness with additions (calculate_additions + get_witness)
    std::vector<FieldShare> add_w;
    add_w.reserve(zk.n_additions);
    auto get = [&](size_t i) -> FieldShare {
      if (i <= zk.n_public) return ShareOps<T>::promote(driver, i == 0 ? fr.zero() : public_inputs[i]);
      if (i < base) return FieldShare{wit_a[i - zk.n_public - 1], wit_b ? wit_b[i - zk.n_public - 1] : fr.zero()};
      if (i < zk.n_vars) return add_w[i - base];
      throw Error("CorruptedWitness(" + std::to_string(i) + ")");
    };
    for (const auto& ad : zk.additions) {
      FieldShare w1 = get(ad.s1), w2 = get(ad.s2);
      add_w.push_back(FieldShare{fr.add(fr.mul(ad.f1, w1.a), fr.mul(ad.f2, w2.a)), fr.add(fr.mul(ad.f1, w1.b), fr.mul(ad.f2, w2.b))});
    }
    // ---- blinders
    FieldShare b[11];
    for (int i = 0; i < 11; i++) b[i] = deterministic ? ShareOps<T>::promote(driver, small(i)) : driver.rand();
    Domain dom, ext;
    dom.log_n = (unsigned)zk.pow;
    ext.log_n = (unsigned)zk.pow + 2;
    dom.group_gen = root_of_unity_for_groth16(zk.curve, zk.pow).omega;       // roots_of_unity[pow]   (co-plonk/src/types.rs:83-89)
    ext.group_gen = root_of_unity_for_groth16(zk.curve, zk.pow + 2).omega;   // roots_of_unity[pow + 2]
    PointShare commits[3];
    const std::vector<uint32_t>* maps[3] = {&zk.map_a, &zk.map_b, &zk.map_c};
    for (int k = 0; k < 3; k++) {
      std::vector<Fr> ha(n, fr.zero()), hb(n, fr.zero());
      for (size_t i = 0; i < zk.n_constraints; i++) {
        FieldShare w = get((*maps[k])[i]);
        ha[i] = w.a;
        hb[i] = w.b;
      }
      FieldShareVec poly = driver.share_vec_from_host(ha.data(), hb.data(), n);
      driver.ifft_in_place(poly, dom);                      // coefficients of the wire polynomial
      FieldShareVec eval = extend(poly, 4 * n, n);
      driver.fft_in_place(eval, ext);                       // extended evaluations (used by rounds 2-3)
      driver.release(eval);
      FieldShareVec blinded = extend(poly, n + 2, n);       // blind_coefficients with coeff_rev = b[2k .. 2k+2]
      driver.release(poly);
      const FieldShare &b_lo = b[2 * k], &b_hi = b[2 * k + 1];
      patch(blinded.a, n, b_lo.a, b_hi.a);
      if (blinded.b.p) patch(blinded.b, n, b_lo.b, b_hi.b);
      commits[k] = driver.msm_public_points(1, p_tau_handle, 0, n + 2, blinded);
      driver.release(blinded);
    }
    Round1Proof pr;  // open_point_many (rep3.rs:855-861)
    pr.commit_a = driver.to_affine(1, driver.open_point(1, commits[0]));
    pr.commit_b = driver.to_affine(1, driver.open_point(1, commits[1]));
    pr.commit_c = driver.to_affine(1, driver.open_point(1, commits[2]));
    return pr;
  }

 private:
  Fr small(uint64_t v) const {
    Fr acc = driver.fr.zero(), basev = driver.fr.one();
    for (; v; v >>= 1) { if (v & 1) acc = driver.fr.add(acc, basev); basev = driver.fr.add(basev, basev); }
    return acc;
  }
  // a zero-extended copy of the first `keep` elements
  FieldShareVec extend(const FieldShareVec& v, size_t len, size_t keep) {
    FieldShareVec o;
    o.a = ext1(v.a, len, keep);
    if (v.b.p) o.b = ext1(v.b, len, keep);
    return o;
  }
  DevVec ext1(const DevVec& v, size_t len, size_t keep) {
    DevVec o = driver.alloc(len);
    check(driver.ctx, cocg_d2d(driver.ctx, o.p, v.p, keep * 32), "cocg_d2d");
    check(driver.ctx, cocg_memset0(driver.ctx, o.at(keep), (len - keep) * 32), "cocg_memset0");
    return o;
  }
  // res[0] -= b_hi; res[1] -= b_lo; res[n] = b_hi; res[n+1] = b_lo
  void patch(DevVec& v, size_t n, const Fr& b_lo, const Fr& b_hi) {
    Fr head[2];
    check(driver.ctx, cocg_d2h(driver.ctx, head, v.p, 64), "cocg_d2h");
    head[0] = driver.fr.sub(head[0], b_hi);
    head[1] = driver.fr.sub(head[1], b_lo);
    check(driver.ctx, cocg_h2d(driver.ctx, v.p, head, 64), "cocg_h2d");
    Fr tail[2] = {b_hi, b_lo};
    check(driver.ctx, cocg_h2d(driver.ctx, v.at(n), tail, 64), "cocg_h2d");
  }
};

}  // namespace cohost
