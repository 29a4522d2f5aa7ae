// CoPlonk<T>: the collaborative Plonk prover on the cocg kernels, generic over the MPC driver -- the five rounds of CoPlonk::prove
// (/root/reference/co-circom/co-plonk/src/lib.rs:80-99):
//   round 1  round1.rs:86-300   blinders, calculate_additions, wire buffers -> iFFT(n), FFT(4n), blind, 3 commitments
//   round 2  round2.rs:146-300  beta / gamma, z(X): factor products, array_prod_mul, inv_many, rotate, iFFT, FFT(4n), blind, [z]
//   round 3  round3.rs:237-520  alpha, quotient t(X) on the 4n domain, iFFT(4n) x 2, division by Z_H, split, [t1] [t2] [t3]
//   round 4  round4.rs:114-165  xi, evaluations of a, b, c at xi and z at xi*omega (opened), sigma1 / sigma2 at xi (public)
//   round 5  round5.rs:96-380   v, linearisation r(X), W_xi, W_xi*omega (division by X - point), 2 commitments
// plus the Plonk zkey reader (circom-types/src/plonk/zkey.rs:160-420).  Same round structure, same names; every vector step is a
// kernel on the driver's context, the share vectors stay in HBM from the witness upload to the last commitment.  Differences from
// the reference, none visible in the proof: the 52 sequential mul_vec rounds of compute_t are two exchanges (csrc/plonk.cu),
// independent products of round 2 share a network round (mul_vec_many), the dependent chains (prefix products, division by X - xi,
// Horner) are scans.  With the reference's deterministic blinders (b_i = i) the proof equals its round KATs bit for bit.
#pragma once
#include <map>

#include "formats.hpp"
#include "groth16.hpp"
#include "plonk_verify.hpp"
#include "shamir.hpp"

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
    // signal indices feed a device gather: refuse what the reference's get_witness would (CorruptedWitness, lib.rs:113-137)
    for (const uint8_t* m : {map_a, map_b, map_c})
      for (size_t i = 0; i < n_constraints; i++)
        if (BinFile::u32(m + 4 * i) >= n_vars) throw Error("zkey: wire map index out of range (CorruptedWitness)");
  }
};

// A Plonk proving key resident in HBM (buffers owned by `owner`; drivers on the same device read them directly)
struct PlonkZKey {
  int curve = 0, device = 0;
  cocg_ctx* owner = nullptr;
  size_t n_vars = 0, n_public = 0, domain_size = 0, pow = 0, n_additions = 0, n_constraints = 0;
  struct Addition { uint32_t s1, s2; Fr f1, f2; };  // factors in Montgomery form (montgomery_bigint_from_reader)
  std::vector<Addition> additions;
  std::vector<uint32_t> map_a, map_b, map_c;
  uint64_t p_tau = 0;
  // rounds 2-5
  bool full = false;
  Fr k1, k2;
  Point vk_g1[8];                       // Qm Ql Qr Qo Qc S1 S2 S3, packed affine Montgomery (transcript of round 2)
  void* sel_coef[5] = {};               // n coefficients each: Qm Ql Qr Qo Qc
  void* sel_eval[5] = {};               // 4n evaluations each
  void* sig_coef[3] = {};
  void* sig_eval[3] = {};
  void* lagrange = nullptr;             // n_lagrange x 4n evaluations, contiguous
  size_t n_lagrange = 0;
  uint32_t* d_map[3] = {};              // wire maps on the device, padded to n with 0xffffffff (= zero)
  std::vector<void*> owned;             // everything cocg_malloc'ed on `owner`
};

struct Round1Proof {
  Point commit_a, commit_b, commit_c;  // packed affine
};
struct PlonkProof {  // circom-types/src/plonk/proof.rs
  Point a, b, c, z, t1, t2, t3, wxi, wxiw;  // packed affine Montgomery
  Fr eval_a, eval_b, eval_c, eval_s1, eval_s2, eval_zw;
};

// Keccak256Transcript (co-plonk/src/types.rs:125-176) with the curve chosen at run time; the byte layout lives in plonk_verify.hpp
class PlonkTranscript {
 public:
  explicit PlonkTranscript(int curve) : curve_(curve) {}
  void add_scalar(const Fr& s) {
    if (curve_ == COCG_BN254) { cocg::Bn254Fr v; memcpy(v.l, s.l, 32); bn_.add_scalar(v); }
    else { cocg::Bls381Fr v; memcpy(v.l, s.l, 32); bls_.add_scalar(v); }
  }
  void add_point(const Point& affine) {
    if (curve_ == COCG_BN254) { typename Bn::G1 g; memcpy(&g, affine.l, sizeof(g)); bn_.add_point(g); }
    else { typename Bls::G1 g; memcpy(&g, affine.l, sizeof(g)); bls_.add_point(g); }
  }
  Fr get_challenge() const {
    Fr r;
    if (curve_ == COCG_BN254) { auto c = bn_.get_challenge(); memcpy(r.l, c.l, 32); }
    else { auto c = bls_.get_challenge(); memcpy(r.l, c.l, 32); }
    return r;
  }

 private:
  using Bn = PlonkVerifier<Bn254Pairing, cocg::Bn254FrP>;
  using Bls = PlonkVerifier<Bls381Pairing, cocg::Bls381FrP>;
  int curve_;
  Bn::Transcript bn_;
  Bls::Transcript bls_;
};

// host-side share arithmetic the rounds need, per driver kind
template <class T> struct ShareOps;
template <> struct ShareOps<PlainDriver> {
  static FieldShare promote(PlainDriver& d, const Fr& v) { return FieldShare{v, d.fr.zero()}; }
};
template <> struct ShareOps<Rep3Protocol> {
  static FieldShare promote(Rep3Protocol& d, const Fr& v) {  // fieldshare.rs:55-79: (v, 0) / (0, v) / (0, 0)
    return FieldShare{d.id() == 0 ? v : d.fr.zero(), d.id() == 1 ? v : d.fr.zero()};
  }
};

template <> struct ShareOps<ShamirProtocol> {
  static FieldShare promote(ShamirProtocol& d, const Fr& v) { return FieldShare{v, d.fr.zero()}; }  // shamir.rs:625-627: every party holds the value
};

template <class T>
class CoPlonk {
 public:
  explicit CoPlonk(T& driver) : driver(driver), fr(driver.fr) {}
  T& driver;
  const FrOps& fr;
  // test hook: component a of named intermediate vectors (the sum over the three parties is the plain value)
  std::map<std::string, std::vector<Fr>>* trace = nullptr;
  double round_s[5] = {0, 0, 0, 0, 0};  // host wall-clock per round of the last prove

  static constexpr int K = T::kComponents;

  // public_inputs: n_public + 1 values (the leading one is overwritten with zero, PlonkWitness::new types.rs:105-108);
  // wit_a / wit_b: the party's share components of values[n_public + 1 ..] (HOST, or DEVICE with wit_on_device; wit_b unused by
  // the plain driver).  deterministic: the reference's KAT blinders b_i = i (round1.rs:101-108).  stop_after_round1 keeps the
  // round-1 entry point of the C ABI.
  PlonkProof prove(const PlonkZKey& zk, uint64_t p_tau_handle, const Fr* public_inputs, const void* wit_a, const void* wit_b, bool deterministic,
                   bool wit_on_device = false, bool stop_after_round1 = false) {
    if (!stop_after_round1 && !zk.full) throw Error("plonk: the zkey holds no selector / sigma / Lagrange sections (rounds 2-5 need them)");
    zk_ = &zk;
    p_tau_ = p_tau_handle;
    n_ = zk.domain_size;
    dom_.log_n = (unsigned)zk.pow;
    ext_.log_n = (unsigned)zk.pow + 2;
    dom_.group_gen = root_of_unity_for_groth16(zk.curve, zk.pow).omega;       // roots_of_unity[pow]   (co-plonk/src/types.rs:83-89)
    ext_.group_gen = root_of_unity_for_groth16(zk.curve, zk.pow + 2).omega;   // roots_of_unity[pow + 2]
    PlonkProof pr;
    double t0 = now();
    round1(public_inputs, wit_a, wit_b, deterministic, wit_on_device, pr);
    round_s[0] = now() - t0;
    if (!stop_after_round1) {
      t0 = now(); round2(public_inputs, pr); round_s[1] = now() - t0;
      t0 = now(); round3(pr); round_s[2] = now() - t0;
      t0 = now(); round4(pr); round_s[3] = now() - t0;
      t0 = now(); round5(public_inputs, pr); round_s[4] = now() - t0;
    }
    cleanup();
    return pr;
  }

 private:
  const PlonkZKey* zk_ = nullptr;
  uint64_t p_tau_ = 0;
  size_t n_ = 0;
  Domain dom_, ext_;
  FieldShare b_[11];
  FieldShareVec buf_[3], poly_[3], eval_[3];  // wire buffers (n), blinded coefficients (n + 2), extended evaluations (4n)
  FieldShareVec poly_z_, eval_z_;             // n + 3, 4n
  FieldShareVec t_[3];                        // n + 1, n + 1, n + 6
  Fr beta_, gamma_, alpha_, xi_, v_[5];
  static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

  void cleanup() {
    for (int k = 0; k < 3; k++) { driver.release(buf_[k]); driver.release(poly_[k]); driver.release(eval_[k]); driver.release(t_[k]); }
    driver.release(poly_z_);
    driver.release(eval_z_);
  }
  void record(const char* name, const DevVec& v, size_t n) {
    if (!trace) return;
    std::vector<Fr> h(n);
    check(driver.ctx, cocg_d2h(driver.ctx, h.data(), v.p, n * 32), "cocg_d2h");
    (*trace)[name] = std::move(h);
  }
  const DevVec& comp(const FieldShareVec& v, int k) const { return k == 0 ? v.a : v.b; }
  DevVec& comp(FieldShareVec& v, int k) { return k == 0 ? v.a : v.b; }
  const Fr& comp(const FieldShare& v, int k) const { return k == 0 ? v.a : v.b; }

  // a zero-extended copy of the first `keep` elements
  FieldShareVec extend(const FieldShareVec& v, size_t len, size_t keep) {
    FieldShareVec o;
    for (int k = 0; k < K; k++) {
      DevVec d = driver.alloc(len);
      check(driver.ctx, cocg_d2d(driver.ctx, d.p, comp(v, k).p, keep * 32), "cocg_d2d");
      check(driver.ctx, cocg_memset0(driver.ctx, d.at(keep), (len - keep) * 32), "cocg_memset0");
      comp(o, k) = d;
    }
    return o;
  }
  // plonk_utils::blind_coefficients (lib.rs:140-158): v has room for n + |rev| coefficients; rev = coeff_rev reversed
  void blind(FieldShareVec& v, size_t n, const std::vector<FieldShare>& rev) {
    const size_t m = rev.size();
    for (int k = 0; k < K; k++) {
      std::vector<Fr> head(m), tail(m);
      check(driver.ctx, cocg_d2h(driver.ctx, head.data(), comp(v, k).p, m * 32), "cocg_d2h");
      for (size_t i = 0; i < m; i++) { head[i] = fr.sub(head[i], comp(rev[i], k)); tail[i] = comp(rev[i], k); }
      check(driver.ctx, cocg_h2d(driver.ctx, comp(v, k).p, head.data(), m * 32), "cocg_h2d");
      check(driver.ctx, cocg_h2d(driver.ctx, comp(v, k).at(n), tail.data(), m * 32), "cocg_h2d");
    }
  }
  // element 0 of a share vector += a public value (add_with_public: the component that carries public addends)
  void add_public_at0(FieldShareVec& v, const Fr& c) {
    const int pc = driver.pub_comp();
    if (pc < 0 || pc >= K) return;
    Fr h;
    check(driver.ctx, cocg_d2h(driver.ctx, h.l, comp(v, pc).p, 32), "cocg_d2h");
    h = fr.add(h, c);
    check(driver.ctx, cocg_h2d(driver.ctx, comp(v, pc).p, h.l, 32), "cocg_h2d");
  }
  void sub_share_at0(FieldShareVec& v, const FieldShare& s) {
    for (int k = 0; k < K; k++) {
      Fr h;
      check(driver.ctx, cocg_d2h(driver.ctx, h.l, comp(v, k).p, 32), "cocg_d2h");
      h = fr.sub(h, comp(s, k));
      check(driver.ctx, cocg_h2d(driver.ctx, comp(v, k).p, h.l, 32), "cocg_h2d");
    }
  }
  void set_share_at(FieldShareVec& v, size_t i, const FieldShare& s) {
    for (int k = 0; k < K; k++) check(driver.ctx, cocg_h2d(driver.ctx, comp(v, k).at(i), comp(s, k).l, 32), "cocg_h2d");
  }
  Point commit(const FieldShareVec& poly, size_t len) {
    PointShare c = driver.msm_public_points(1, p_tau_, 0, len, poly);
    return driver.to_affine(1, driver.open_point(1, c));
  }
  // The commitments of one round (round1.rs:276-292, round3.rs:498-517, round5.rs:351-358): the polynomials are multiplied against
  // the same p_tau, so ONE MSM call takes them as extra scalar vectors (their bucket sets are reduced together, one synchronisation)
  // and open_point_many opens them with one message.  `len` covers the longest; every vector must hold zeros from its own end to len.
  std::vector<Point> commit_many(const std::vector<const FieldShareVec*>& polys, size_t len) {
    std::vector<PointShare> c = driver.msm_public_points_many(1, p_tau_, len, polys);
    std::vector<Point> opened = driver.open_point_many(1, c);
    for (auto& p : opened) p = driver.to_affine(1, p);
    return opened;
  }
  void zero_tail(FieldShareVec& v, size_t from, size_t to) {
    if (to <= from) return;
    for (int k = 0; k < K; k++) check(driver.ctx, cocg_memset0(driver.ctx, comp(v, k).at(from), (to - from) * 32), "cocg_memset0");
  }

  // ------------------------------------------------------------------------------------------------ round 1
  void round1(const Fr* public_inputs, const void* wit_a, const void* wit_b, bool deterministic, bool wit_on_device, PlonkProof& pr) {
    const PlonkZKey& zk = *zk_;
    const size_t n = n_, base = zk.n_vars - zk.n_additions, n_wit = base - zk.n_public - 1;
    for (int i = 0; i < 11; i++) b_[i] = deterministic ? ShareOps<T>::promote(driver, fr.from_u64(i)) : driver.rand();
    // ---- the signal vector: public inputs (trivial shares) | private witness | additions
    FieldShareVec sig = driver.alloc_share(zk.n_vars);
    {
      std::vector<Fr> ha(zk.n_public + 1), hb(zk.n_public + 1);
      for (size_t i = 0; i <= zk.n_public; i++) {
        FieldShare s = ShareOps<T>::promote(driver, i == 0 ? fr.zero() : public_inputs[i]);
        ha[i] = s.a;
        hb[i] = s.b;
      }
      check(driver.ctx, cocg_h2d(driver.ctx, sig.a.p, ha.data(), ha.size() * 32), "cocg_h2d");
      if (K == 2) check(driver.ctx, cocg_h2d(driver.ctx, sig.b.p, hb.data(), hb.size() * 32), "cocg_h2d");
      const void* w[2] = {wit_a, wit_b};
      for (int k = 0; k < K; k++) {
        if (!n_wit) break;
        if (wit_on_device) check(driver.ctx, cocg_d2d(driver.ctx, comp(sig, k).at(zk.n_public + 1), w[k], n_wit * 32), "cocg_d2d");
        else check(driver.ctx, cocg_h2d(driver.ctx, comp(sig, k).at(zk.n_public + 1), w[k], n_wit * 32), "cocg_h2d");
      }
    }
    if (zk.n_additions) {  // calculate_additions (round1.rs:213-242): each may read earlier additions -- a dependent chain, host side
      std::vector<Fr> wa(n_wit), wb(K == 2 ? n_wit : 0);
      if (wit_on_device) {
        check(driver.ctx, cocg_d2h(driver.ctx, wa.data(), wit_a, n_wit * 32), "cocg_d2h");
        if (K == 2) check(driver.ctx, cocg_d2h(driver.ctx, wb.data(), wit_b, n_wit * 32), "cocg_d2h");
      } else {
        memcpy(wa.data(), wit_a, n_wit * 32);
        if (K == 2) memcpy(wb.data(), wit_b, n_wit * 32);
      }
      std::vector<FieldShare> add_w;
      add_w.reserve(zk.n_additions);
      auto get = [&](size_t i) -> FieldShare {
        if (i <= zk.n_public) return ShareOps<T>::promote(driver, i == 0 ? fr.zero() : public_inputs[i]);
        if (i < base) return FieldShare{wa[i - zk.n_public - 1], K == 2 ? wb[i - zk.n_public - 1] : fr.zero()};
        if (i < base + add_w.size()) return add_w[i - base];
        throw Error("CorruptedWitness(" + std::to_string(i) + ")");
      };
      for (const auto& ad : zk.additions) {
        FieldShare w1 = get(ad.s1), w2 = get(ad.s2);
        add_w.push_back(FieldShare{fr.add(fr.mul(ad.f1, w1.a), fr.mul(ad.f2, w2.a)), fr.add(fr.mul(ad.f1, w1.b), fr.mul(ad.f2, w2.b))});
      }
      std::vector<Fr> ha(zk.n_additions), hb(zk.n_additions);
      for (size_t i = 0; i < zk.n_additions; i++) { ha[i] = add_w[i].a; hb[i] = add_w[i].b; }
      check(driver.ctx, cocg_h2d(driver.ctx, sig.a.at(base), ha.data(), ha.size() * 32), "cocg_h2d");
      if (K == 2) check(driver.ctx, cocg_h2d(driver.ctx, sig.b.at(base), hb.data(), hb.size() * 32), "cocg_h2d");
    }
    // ---- compute_wire_polynomials (round1.rs:121-209)
    // the three wires go through each transform as one launch sequence (a 2^18 transform alone is less than one wave of blocks)
    FieldShareVec wire_poly[3];
    for (int w = 0; w < 3; w++) {
      buf_[w] = driver.alloc_share(n);
      for (int k = 0; k < K; k++)
        check(driver.ctx, cocg_vec_gather(driver.ctx, comp(sig, k).p, zk.n_vars, zk.d_map[w], comp(buf_[w], k).p, n), "cocg_vec_gather");
      wire_poly[w] = extend(buf_[w], n, n);
    }
    driver.ifft_many({&wire_poly[0], &wire_poly[1], &wire_poly[2]}, dom_);
    for (int w = 0; w < 3; w++) eval_[w] = extend(wire_poly[w], 4 * n, n);
    driver.fft_many({&eval_[0], &eval_[1], &eval_[2]}, ext_);
    for (int w = 0; w < 3; w++) {
      poly_[w] = extend(wire_poly[w], n + 2, n);
      driver.release(wire_poly[w]);
      blind(poly_[w], n, {b_[2 * w + 1], b_[2 * w]});  // coeff_rev = b[2w .. 2w + 2]
    }
    driver.release(sig);
    std::vector<Point> commits = commit_many({&poly_[0], &poly_[1], &poly_[2]}, n + 2);
    pr.a = commits[0];
    pr.b = commits[1];
    pr.c = commits[2];
    if (trace) { record("buffer_a", buf_[0].a, n); record("poly_a", poly_[0].a, n + 2); record("eval_a", eval_[0].a, 4 * n); }
  }

  // ------------------------------------------------------------------------------------------------ round 2
  void round2(const Fr* public_inputs, PlonkProof& pr) {
    const PlonkZKey& zk = *zk_;
    const size_t n = n_;
    PlonkTranscript tr(zk.curve);
    for (int i = 0; i < 8; i++) tr.add_point(zk.vk_g1[i]);
    for (size_t i = 1; i <= zk.n_public; i++) tr.add_scalar(public_inputs[i]);  // the leading zero is dropped (round1.rs:42-46)
    tr.add_point(pr.a);
    tr.add_point(pr.b);
    tr.add_point(pr.c);
    beta_ = tr.get_challenge();
    PlonkTranscript tr2(zk.curve);
    tr2.add_scalar(beta_);
    gamma_ = tr2.get_challenge();
    // ---- compute_z (round2.rs:146-241)
    FieldShareVec fac[6];
    for (int j = 0; j < 6; j++) fac[j] = driver.alloc_share(n);
    for (int k = 0; k < K; k++) {
      cocg_plonk_z_args za;
      za.a = comp(buf_[0], k).p; za.b = comp(buf_[1], k).p; za.c = comp(buf_[2], k).p;
      za.sigma1 = zk.sig_eval[0]; za.sigma2 = zk.sig_eval[1]; za.sigma3 = zk.sig_eval[2];
      za.beta = beta_.l; za.gamma = gamma_.l; za.k1 = zk.k1.l; za.k2 = zk.k2.l; za.omega = dom_.group_gen.l;
      for (int j = 0; j < 6; j++) za.out[j] = comp(fac[j], k).p;
      check(driver.ctx, cocg_plonk_z_factors(driver.ctx, &za, n, k == driver.pub_comp() ? 1 : 0), "cocg_plonk_z_factors");
    }
    std::vector<FieldShareVec> m1 = driver.mul_vec_many({{&fac[0], &fac[1]}, {&fac[3], &fac[4]}});
    std::vector<FieldShareVec> m2 = driver.mul_vec_many({{&m1[0], &fac[2]}, {&m1[1], &fac[5]}});
    driver.release_many(m1);
    for (int j = 0; j < 6; j++) driver.release(fac[j]);
    FieldShareVec num = driver.array_prod_mul(m2[0]);
    FieldShareVec den = driver.array_prod_mul(m2[1]);
    driver.release_many(m2);
    FieldShareVec den_inv = driver.inv_many(den);
    driver.release(den);
    FieldShareVec bz = driver.mul_vec(num, den_inv);
    driver.release(num);
    driver.release(den_inv);
    FieldShareVec poly = driver.alloc_share(n);  // buffer_z.rotate_right(1)
    for (int k = 0; k < K; k++) {
      check(driver.ctx, cocg_d2d(driver.ctx, comp(poly, k).at(1), comp(bz, k).p, (n - 1) * 32), "cocg_d2d");
      check(driver.ctx, cocg_d2d(driver.ctx, comp(poly, k).p, comp(bz, k).at(n - 1), 32), "cocg_d2d");
    }
    driver.release(bz);
    if (trace) record("buffer_z", poly.a, n);
    driver.ifft_in_place(poly, dom_);
    eval_z_ = extend(poly, 4 * n, n);
    driver.fft_in_place(eval_z_, ext_);
    poly_z_ = extend(poly, n + 3, n);
    driver.release(poly);
    blind(poly_z_, n, {b_[8], b_[7], b_[6]});  // coeff_rev = b[6..9]
    if (trace) record("poly_z", poly_z_.a, n + 3);
    pr.z = commit(poly_z_, n + 3);
  }

  // ------------------------------------------------------------------------------------------------ round 3
  void round3(PlonkProof& pr) {
    const PlonkZKey& zk = *zk_;
    const size_t n = n_, n4 = 4 * n;
    PlonkTranscript tr(zk.curve);
    tr.add_scalar(beta_);
    tr.add_scalar(gamma_);
    tr.add_point(pr.z);
    alpha_ = tr.get_challenge();
    // products of the blinders that the blinding-polynomial products ap*bp, cp*zp, cp*zwp need (one small network round)
    static const int pa[10] = {1, 0, 1, 0, 5, 5, 5, 4, 4, 4}, pb[10] = {3, 3, 2, 2, 8, 7, 6, 8, 7, 6};
    std::vector<FieldShare> xa, xb;
    for (int j = 0; j < 10; j++) { xa.push_back(b_[pa[j]]); xb.push_back(b_[pb[j]]); }
    std::vector<FieldShare> sp = driver.mul_many(xa, xb);
    Fr blinders[9][2], sprod[10][2];
    for (int j = 0; j < 9; j++) { blinders[j][0] = b_[j].a; blinders[j][1] = b_[j].b; }
    for (int j = 0; j < 10; j++) { sprod[j][0] = sp[j].a; sprod[j][1] = sp[j].b; }
    cocg_plonk_quotient_args qa;
    memset(&qa, 0, sizeof(qa));
    qa.components = K;
    qa.pub_comp = driver.pub_comp();
    qa.n4 = n4;
    qa.n_public = zk.n_public;
    qa.n_lagrange = zk.n_lagrange;
    for (int k = 0; k < K; k++) {
      qa.eval_a[k] = comp(eval_[0], k).p; qa.eval_b[k] = comp(eval_[1], k).p; qa.eval_c[k] = comp(eval_[2], k).p;
      qa.eval_z[k] = comp(eval_z_, k).p;
      qa.buffer_a[k] = comp(buf_[0], k).p;
    }
    qa.sigma1 = zk.sig_eval[0]; qa.sigma2 = zk.sig_eval[1]; qa.sigma3 = zk.sig_eval[2];
    qa.qm = zk.sel_eval[0]; qa.ql = zk.sel_eval[1]; qa.qr = zk.sel_eval[2]; qa.qo = zk.sel_eval[3]; qa.qc = zk.sel_eval[4];
    qa.lagrange = zk.lagrange;
    qa.beta = beta_.l; qa.gamma = gamma_.l; qa.alpha = alpha_.l; qa.k1 = zk.k1.l; qa.k2 = zk.k2.l;
    qa.omega_n = dom_.group_gen.l; qa.omega_4n = ext_.group_gen.l;
    qa.blinders = blinders;
    qa.scalar_products = sprod;
    qa.seed_own = driver.seed_own();
    qa.seed_prev = driver.seed_prev();
    // level 1: six product vectors, one exchange
    DevVec l1 = driver.alloc(6 * n4);
    qa.ctr = driver.take_ctr(6);
    qa.out = l1.p;
    check(driver.ctx, cocg_plonk_quotient_l1(driver.ctx, &qa), "cocg_plonk_quotient_l1");
    FieldShareVec l1s = driver.reshare(l1);
    // level 2: t and tz, one exchange
    for (int k = 0; k < K; k++) qa.level1[k] = comp(l1s, k).p;
    DevVec l2 = driver.alloc(2 * n4);
    qa.ctr = driver.take_ctr(2);
    qa.out = l2.p;
    check(driver.ctx, cocg_plonk_quotient_l2(driver.ctx, &qa), "cocg_plonk_quotient_l2");
    FieldShareVec tt = driver.reshare(l2);
    driver.release(l1s);
    FieldShareVec t = driver.slice(tt, 0, n4), tz = driver.slice(tt, n4, n4);
    if (trace) { record("t_evals", t.a, n4); record("tz_evals", tz.a, n4); }
    driver.ifft_many({&t, &tz}, ext_);
    t_[0] = driver.alloc_share(n + 6);  // t1, t2 hold n + 1 coefficients; zero up to n + 6 so that one MSM call commits all three
    t_[1] = driver.alloc_share(n + 6);
    t_[2] = driver.alloc_share(n + 6);
    zero_tail(t_[0], n, n + 6);
    zero_tail(t_[1], n, n + 6);
    for (int k = 0; k < K; k++)
      check(driver.ctx, cocg_plonk_t_finish(driver.ctx, comp(t, k).p, comp(tz, k).p, n, comp(t_[0], k).p, comp(t_[1], k).p, comp(t_[2], k).p), "cocg_plonk_t_finish");
    driver.release(tt);
    set_share_at(t_[0], n, b_[9]);   // t1.push(b[9])
    sub_share_at0(t_[1], b_[9]);     // t2[0] -= b[9]
    set_share_at(t_[1], n, b_[10]);  // t2.push(b[10])
    sub_share_at0(t_[2], b_[10]);    // t3[0] -= b[10]
    if (trace) { record("t1", t_[0].a, n + 1); record("t2", t_[1].a, n + 1); record("t3", t_[2].a, n + 6); }
    std::vector<Point> ct = commit_many({&t_[0], &t_[1], &t_[2]}, n + 6);
    pr.t1 = ct[0];
    pr.t2 = ct[1];
    pr.t3 = ct[2];
  }

  // ------------------------------------------------------------------------------------------------ round 4
  void round4(PlonkProof& pr) {
    const PlonkZKey& zk = *zk_;
    const size_t n = n_;
    PlonkTranscript tr(zk.curve);
    tr.add_scalar(alpha_);
    tr.add_point(pr.t1);
    tr.add_point(pr.t2);
    tr.add_point(pr.t3);
    xi_ = tr.get_challenge();
    const Fr xiw = fr.mul(xi_, dom_.group_gen);
    std::vector<FieldShare> ev = {driver.evaluate_poly_public(poly_[0], n + 2, xi_), driver.evaluate_poly_public(poly_[1], n + 2, xi_),
                                  driver.evaluate_poly_public(poly_[2], n + 2, xi_), driver.evaluate_poly_public(poly_z_, n + 3, xiw)};
    std::vector<Fr> opened = driver.open_many(ev);
    pr.eval_a = opened[0];
    pr.eval_b = opened[1];
    pr.eval_c = opened[2];
    pr.eval_zw = opened[3];
    pr.eval_s1 = driver.eval_public(DevVec{zk.sig_coef[0], n}, n, xi_);
    pr.eval_s2 = driver.eval_public(DevVec{zk.sig_coef[1], n}, n, xi_);
  }

  // ------------------------------------------------------------------------------------------------ round 5
  // div_by_zerofier(inout, 1, beta) (round5.rs:97-115): y_i = (y_{i-1} - x_i) / beta  ==  y_i = -beta^-(i+1) * sum_{j<=i} beta^j x_j
  void div_by_linear(FieldShareVec& v, size_t len, const Fr& beta) {
    const Fr one = fr.one(), inv = fr.inv(beta), ninv = fr.neg(inv);
    for (int k = 0; k < K; k++) {
      void* p = comp(v, k).p;
      check(driver.ctx, cocg_vec_scale_powers(driver.ctx, p, len, beta.l, one.l), "cocg_vec_scale_powers");
      check(driver.ctx, cocg_vec_scan(driver.ctx, COCG_OP_ADD, p, p, len), "cocg_vec_scan");
      check(driver.ctx, cocg_vec_scale_powers(driver.ctx, p, len, inv.l, ninv.l), "cocg_vec_scale_powers");
    }
  }
  void lincomb(DevVec& out, size_t len, const std::vector<std::pair<const void*, size_t>>& vecs, const std::vector<Fr>& f) {
    std::vector<const void*> ptr;
    std::vector<size_t> lens;
    for (auto& v : vecs) { ptr.push_back(v.first); lens.push_back(v.second); }
    check(driver.ctx, cocg_vec_lincomb(driver.ctx, (int)ptr.size(), ptr.data(), lens.data(), f.data(), out.p, len), "cocg_vec_lincomb");
  }
  void round5(const Fr* public_inputs, PlonkProof& pr) {
    const PlonkZKey& zk = *zk_;
    const size_t n = n_, len = n + 6;
    PlonkTranscript tr(zk.curve);
    tr.add_scalar(xi_);
    for (const Fr* s : {&pr.eval_a, &pr.eval_b, &pr.eval_c, &pr.eval_s1, &pr.eval_s2, &pr.eval_zw}) tr.add_scalar(*s);
    v_[0] = tr.get_challenge();
    for (int i = 1; i < 5; i++) v_[i] = fr.mul(v_[i - 1], v_[0]);
    // ---- compute_r (round5.rs:140-250)
    Fr xin = xi_;
    for (size_t i = 0; i < zk.pow; i++) xin = fr.sqr(xin);
    const Fr one = fr.one(), zh = fr.sub(xin, one), nfr = fr.from_u64(n);
    const size_t l_len = zk.n_public > 1 ? zk.n_public : 1;  // calculate_lagrange_evaluations (lib.rs:160-185)
    std::vector<Fr> l(l_len);
    Fr w = one;
    for (size_t i = 0; i < l_len; i++) {
      l[i] = fr.mul(fr.mul(w, zh), fr.inv(fr.mul(nfr, fr.sub(xi_, w))));
      w = fr.mul(w, dom_.group_gen);
    }
    Fr eval_pi = fr.zero();
    for (size_t i = 0; i < zk.n_public; i++) eval_pi = fr.sub(eval_pi, fr.mul(l[i], public_inputs[i + 1]));
    const Fr coef_ab = fr.mul(pr.eval_a, pr.eval_b), betaxi = fr.mul(beta_, xi_);
    const Fr e2a = fr.add(fr.add(pr.eval_a, betaxi), gamma_), e2b = fr.add(fr.add(pr.eval_b, fr.mul(betaxi, zk.k1)), gamma_),
             e2c = fr.add(fr.add(pr.eval_c, fr.mul(betaxi, zk.k2)), gamma_);
    const Fr e2 = fr.mul(fr.mul(fr.mul(e2a, e2b), e2c), alpha_);
    const Fr e3a = fr.add(fr.add(pr.eval_a, fr.mul(beta_, pr.eval_s1)), gamma_), e3b = fr.add(fr.add(pr.eval_b, fr.mul(beta_, pr.eval_s2)), gamma_);
    const Fr e3 = fr.mul(fr.mul(fr.mul(e3a, e3b), pr.eval_zw), alpha_);
    const Fr e4 = fr.mul(fr.sqr(alpha_), l[0]), e24 = fr.add(e2, e4);
    const Fr xin2 = fr.sqr(xin);
    const int pc = driver.pub_comp();
    FieldShareVec r = driver.alloc_share(len);
    for (int k = 0; k < K; k++) {
      std::vector<std::pair<const void*, size_t>> vecs = {{comp(poly_z_, k).p, n + 3}, {comp(t_[2], k).p, n + 6}, {comp(t_[1], k).p, n + 1}, {comp(t_[0], k).p, n + 1}};
      std::vector<Fr> f = {e24, fr.neg(fr.mul(zh, xin2)), fr.neg(fr.mul(zh, xin)), fr.neg(zh)};
      DevVec pub;
      if (k == pc) {  // the public part of r: qm*ab + ql*a + qr*b + qo*c + qc - s3 * e3 * beta
        pub = driver.alloc(n);
        lincomb(pub, n, {{zk.sel_coef[0], n}, {zk.sel_coef[1], n}, {zk.sel_coef[2], n}, {zk.sel_coef[3], n}, {zk.sel_coef[4], n}, {zk.sig_coef[2], n}},
                {coef_ab, pr.eval_a, pr.eval_b, pr.eval_c, one, fr.neg(fr.mul(e3, beta_))});
        vecs.push_back({pub.p, n});
        f.push_back(one);
      }
      lincomb(comp(r, k), len, vecs, f);
      if (k == pc) driver.release(pub);
    }
    add_public_at0(r, fr.sub(fr.sub(eval_pi, fr.mul(e3, fr.add(pr.eval_c, gamma_))), e4));  // r0
    if (trace) record("poly_r", r.a, len);
    // ---- compute_wxi (round5.rs:253-312)
    FieldShareVec wxi = driver.alloc_share(len);
    for (int k = 0; k < K; k++) {
      std::vector<std::pair<const void*, size_t>> vecs = {{comp(r, k).p, len}, {comp(poly_[0], k).p, n + 2}, {comp(poly_[1], k).p, n + 2}, {comp(poly_[2], k).p, n + 2}};
      std::vector<Fr> f = {one, v_[0], v_[1], v_[2]};
      if (k == pc) {
        vecs.push_back({zk.sig_coef[0], n});
        vecs.push_back({zk.sig_coef[1], n});
        f.push_back(v_[3]);
        f.push_back(v_[4]);
      }
      lincomb(comp(wxi, k), len, vecs, f);
    }
    driver.release(r);
    Fr c0 = fr.add(fr.add(fr.mul(v_[0], pr.eval_a), fr.mul(v_[1], pr.eval_b)), fr.add(fr.mul(v_[2], pr.eval_c), fr.add(fr.mul(v_[3], pr.eval_s1), fr.mul(v_[4], pr.eval_s2))));
    add_public_at0(wxi, fr.neg(c0));
    div_by_linear(wxi, len, xi_);
    if (trace) record("wxi", wxi.a, len - 1);
    // ---- compute_wxiw (round5.rs:315-330)
    FieldShareVec wxiw = extend(poly_z_, len, n + 3);
    add_public_at0(wxiw, fr.neg(pr.eval_zw));
    div_by_linear(wxiw, n + 3, fr.mul(xi_, dom_.group_gen));
    zero_tail(wxiw, n + 2, len);  // the quotient has n + 2 coefficients; slot n + 2 holds the (vanishing) remainder term
    std::vector<Point> cw = commit_many({&wxi, &wxiw}, len - 1);
    pr.wxi = cw[0];
    pr.wxiw = cw[1];
    driver.release(wxi);
    driver.release(wxiw);
  }
};

}  // namespace cohost
