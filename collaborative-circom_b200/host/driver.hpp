// MPC drivers over the cocg C ABI: the C++ mirror of the reference's trait implementations
//   PlainDriver   /root/reference/mpc-core/src/protocols/plain.rs:111-416
//   Rep3Protocol  /root/reference/mpc-core/src/protocols/rep3.rs:491-947
// for the traits PrimeFieldMpcProtocol, EcMpcProtocol, PairingEcMpcProtocol, FFTProvider, MSMProvider
// (/root/reference/mpc-core/src/traits.rs:43-223, 472-568).  Method names, argument meaning and the party-id rules are
// the reference's; the bodies enqueue kernels on the driver's cocg context instead of looping on the CPU.  Share
// vectors live in HBM between calls; only the MPC network rounds cross PCIe.
#pragma once
#include <future>
#include <map>
#include <memory>

#include "../csrc/ec.cuh"  // host branch: Fr / point arithmetic for the O(1) bookkeeping
#include "network.hpp"

namespace cohost {

// ---------------------------------------------------------------- host scalar-field helpers (curve chosen at run time)
struct FrOps {
  int curve;
  template <class Fn>
  Fr un(const Fr& a, Fn fn) const {
    Fr r;
    if (curve == COCG_BN254) { cocg::Bn254Fr x; memcpy(x.l, a.l, 32); auto y = fn(x); memcpy(r.l, y.l, 32); }
    else { cocg::Bls381Fr x; memcpy(x.l, a.l, 32); auto y = fn(x); memcpy(r.l, y.l, 32); }
    return r;
  }
  template <class Fn>
  Fr bin(const Fr& a, const Fr& b, Fn fn) const {
    Fr r;
    if (curve == COCG_BN254) { cocg::Bn254Fr x, y; memcpy(x.l, a.l, 32); memcpy(y.l, b.l, 32); auto z = fn(x, y); memcpy(r.l, z.l, 32); }
    else { cocg::Bls381Fr x, y; memcpy(x.l, a.l, 32); memcpy(y.l, b.l, 32); auto z = fn(x, y); memcpy(r.l, z.l, 32); }
    return r;
  }
  Fr add(const Fr& a, const Fr& b) const { return bin(a, b, [](auto x, auto y) { return cocg::fp_add(x, y); }); }
  Fr sub(const Fr& a, const Fr& b) const { return bin(a, b, [](auto x, auto y) { return cocg::fp_sub(x, y); }); }
  Fr mul(const Fr& a, const Fr& b) const { return bin(a, b, [](auto x, auto y) { return cocg::fp_mul(x, y); }); }
  Fr from_mont(const Fr& a) const { return un(a, [](auto x) { return cocg::fp_from_mont(x); }); }
  Fr neg(const Fr& a) const { return un(a, [](auto x) { return cocg::fp_neg(x); }); }
  Fr sqr(const Fr& a) const { return un(a, [](auto x) { return cocg::fp_sqr(x); }); }
  Fr inv(const Fr& a) const { return un(a, [](auto x) { return cocg::fp_inv(x); }); }
  bool is_zero(const Fr& a) const { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
  Fr from_u64(uint64_t v) const {  // Montgomery form of a small integer
    Fr acc = zero(), base = one();
    for (; v; v >>= 1) { if (v & 1) acc = add(acc, base); base = add(base, base); }
    return acc;
  }
  Fr zero() const { Fr z; memset(z.l, 0, 32); return z; }
  Fr one() const { return un(zero(), [](auto x) { return decltype(x)::one(); }); }
};

// ---------------------------------------------------------------- shared plumbing of both drivers
class DeviceDriver {
 public:
  DeviceDriver(int curve, int device) : fr{curve} {
    if (cocg_create(&ctx, device, curve)) throw Error(std::string("cocg_create: ") + cocg_last_error(nullptr));
    lq = curve == COCG_BN254 ? 4 : 6;
  }
  // A second context on the same device (own stream, own scratch): single-GPU proving runs the four MSMs over the witness on it from
  // a helper thread while this driver's context computes the witness map (CoGroth16::prove) -- the MSM kernels then fill the bubbles
  // of the witness map's exchange rounds and, end to end, of its PCIe staging.
  cocg_ctx* aux_ctx = nullptr;
  void enable_aux_ctx(int device) {
    if (!aux_ctx && cocg_create(&aux_ctx, device, fr.curve)) throw Error(std::string("cocg_create (aux): ") + cocg_last_error(nullptr));
  }
  virtual ~DeviceDriver() {
    if (aux_ctx) cocg_destroy(aux_ctx);
    if (ctx) {
      for (auto& kv : free_) for (void* p : kv.second) cocg_free(ctx, p);
      for (auto& kv : pinned_free_) for (void* p : kv.second) cocg_host_free(ctx, p);
      cocg_destroy(ctx);
    }
  }
  DeviceDriver(const DeviceDriver&) = delete;

  // size-bucketed free lists: share vectors of one proof have a handful of distinct lengths and are recycled across
  // proofs (cudaFree would serialise the device)
  DevVec alloc(size_t n) {
    DevVec v;
    v.n = n;
    auto& fl = free_[n];
    if (!fl.empty()) { v.p = fl.back(); fl.pop_back(); return v; }
    check(ctx, cocg_malloc(ctx, (n ? n : 1) * 32, &v.p), "cocg_malloc");
    return v;
  }
  void release(DevVec& v) {
    if (v.p) free_[v.n].push_back(v.p);
    v.p = nullptr;
    v.n = 0;
  }
  void release(FieldShareVec& v) { release(v.a); release(v.b); }
  std::shared_ptr<void> pinned(size_t bytes) {
    void* p = nullptr;
    {
      std::lock_guard<std::mutex> lk(pin_mu_);
      auto& fl = pinned_free_[bytes];
      if (!fl.empty()) { p = fl.back(); fl.pop_back(); }
    }
    if (!p) check(ctx, cocg_host_alloc(ctx, bytes, &p), "cocg_host_alloc");
    // the receiver may drop the buffer from another thread: return it to the pool under the lock
    return std::shared_ptr<void>(p, [this, bytes](void* q) {
      std::lock_guard<std::mutex> lk(pin_mu_);
      pinned_free_[bytes].push_back(q);
    });
  }
  DevVec upload(const void* host, size_t n) {
    DevVec v = alloc(n);
    check(ctx, cocg_h2d(ctx, v.p, host, n * 32), "cocg_h2d");
    return v;
  }
  void download(const DevVec& v, void* host) { check(ctx, cocg_d2h(ctx, host, v.p, v.n * 32), "cocg_d2h"); }
  // non-owning view of a sub-range (never released)
  static DevVec slice(const DevVec& v, size_t off, size_t n) {
    if (off + n > v.n) throw Error("slice: range");
    return DevVec{v.at(off), n};
  }
  static FieldShareVec slice(const FieldShareVec& v, size_t off, size_t n) {
    FieldShareVec o;
    o.a = slice(v.a, off, n);
    if (v.b.p) o.b = slice(v.b, off, n);
    return o;
  }
  // ---- public (opened) vectors in HBM: the CoPlonk rounds keep them on the device between steps
  DevVec pub_mul(const DevVec& a, const DevVec& b) { DevVec o = alloc(a.n); check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, a.p, b.p, o.p, a.n), "cocg_vec_op"); return o; }
  void pub_scan_mul(DevVec& v) { check(ctx, cocg_vec_scan(ctx, COCG_OP_MUL, v.p, v.p, v.n), "cocg_vec_scan"); }
  // 1 / v element-wise; a zero raises the reference's error (rep3.rs:549-554, plain.rs inverse().unwrap())
  void pub_inv(DevVec& v) {
    size_t zeros = 0;
    check(ctx, cocg_vec_inv(ctx, v.p, v.p, v.n, &zeros), "cocg_vec_inv");
    if (zeros) throw Error("During execution of inverse in MPC: cannot compute inverse of zero");
  }
  Fr eval_public(const DevVec& coeffs, size_t n, const Fr& point) {
    Fr r;
    check(ctx, cocg_poly_eval(ctx, coeffs.p, n, point.l, r.l), "cocg_poly_eval");
    return r;
  }

  // ---- O(1) group operations on host Jacobian points (K7)
  Point ec(int group, int op, const Point* a, const void* b = nullptr) {
    Point r;
    check(ctx, cocg_ec_op(ctx, group, op, a ? a->l : nullptr, b, r.l), "cocg_ec_op");
    return r;
  }
  Point ec_add(int group, const Point& a, const Point& b) { return ec(group, 0, &a, b.l); }
  Point ec_neg(int group, const Point& a) { return ec(group, 4, &a); }
  Point ec_sub(int group, const Point& a, const Point& b) { Point nb = ec_neg(group, b); return ec_add(group, a, nb); }
  Point ec_mul(int group, const Point& a, const Fr& k_mont) { Fr k = fr.from_mont(k_mont); return ec(group, 1, &a, k.l); }
  Point from_affine(int group, const Point& a) { return ec(group, 3, &a); }
  Point to_affine(int group, const Point& a) { return ec(group, 2, &a); }
  Point generator(int group) { return ec(group, 6, nullptr); }
  Point infinity(int group) { Point z; return from_affine(group, z); }

  cocg_ctx* ctx = nullptr;
  FrOps fr;
  int lq = 4;

 protected:
  std::map<size_t, std::vector<void*>> free_;
  std::map<size_t, std::vector<void*>> pinned_free_;
  std::mutex pin_mu_;
};

// Values a test may inject in place of the drivers' randomness (the reference draws them from entropy and so pins no
// proof bytes, plain.rs:202-205 / rep3.rs:595-598).  Consumed in call order.
struct InjectedRandomness {
  std::vector<FieldShare> rand;               // results of rand()
  std::vector<Fr> mul_masks;                  // masking_field_element for mul()
  std::vector<const void*> mul_vec_masks;     // HOST vectors, one per mul_vec call
  std::vector<Point> ec_masks;                // Jacobian, masking_ec_element for scalar_mul()
  size_t i_rand = 0, i_mul = 0, i_vec = 0, i_ec = 0;
};

// ---------------------------------------------------------------- PlainDriver
class PlainDriver : public DeviceDriver {
 public:
  PlainDriver(int curve, int device) : DeviceDriver(curve, device) {}
  static constexpr int kComponents = 1;
  InjectedRandomness* injected = nullptr;
  uint8_t seed[32] = {};
  uint32_t ctr = 0;

  FieldShare rand() {  // plain.rs:202-205
    if (injected && injected->i_rand < injected->rand.size()) return injected->rand[injected->i_rand++];
    FieldShare r;
    cocg_prf_field_host(fr.curve, seed, ctr++, 0, r.a.l);
    r.b = fr.zero();
    return r;
  }
  FieldShare mul(const FieldShare& a, const FieldShare& b) { return FieldShare{fr.mul(a.a, b.a), fr.zero()}; }  // plain.rs:119-121
  FieldShareVec share_vec_from_host(const void* a, const void*, size_t n) { return FieldShareVec{upload(a, n), DevVec{}}; }

  FieldShareVec evaluate_constraints(uint64_t csr, size_t rows, size_t out_len, const DevVec& public_inputs, const FieldShareVec& witness) {
    FieldShareVec o{alloc(out_len), DevVec{}};  // plain.rs:243-251, one row per thread
    check(ctx, cocg_spmv(ctx, csr, public_inputs.p, public_inputs.n, witness.a.p, o.a.p), "cocg_spmv");
    if (out_len > rows) check(ctx, cocg_memset0(ctx, o.a.at(rows), (out_len - rows) * 32), "cocg_memset0");
    return o;
  }
  FieldShareVec promote_to_trivial_shares(const DevVec& pub) {  // plain.rs:231-233
    FieldShareVec o{alloc(pub.n), DevVec{}};
    check(ctx, cocg_d2d(ctx, o.a.p, pub.p, pub.n * 32), "cocg_d2d");
    return o;
  }
  void clone_from_slice(FieldShareVec& dst, const FieldShareVec& src, size_t dst_off, size_t src_off, size_t len) {
    if (dst.len() < dst_off + len || src.len() < src_off + len || len == 0) throw Error("clone_from_slice: range");  // plain.rs:253-265 asserts
    check(ctx, cocg_d2d(ctx, dst.a.at(dst_off), src.a.at(src_off), len * 32), "cocg_d2d");
  }
  FieldShareVec mul_vec(const FieldShareVec& a, const FieldShareVec& b) {  // plain.rs:220-229
    FieldShareVec o{alloc(a.len()), DevVec{}};
    check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, a.a.p, b.a.p, o.a.p, a.len()), "cocg_vec_op");
    return o;
  }
  void sub_assign_vec(FieldShareVec& a, const FieldShareVec& b) { check(ctx, cocg_vec_op(ctx, COCG_OP_SUB, a.a.p, b.a.p, a.a.p, a.len()), "cocg_vec_op"); }
  void distribute_powers_and_mul_by_const(FieldShareVec& v, const Fr& g, const Fr& c) {
    check(ctx, cocg_vec_scale_powers(ctx, v.a.p, v.len(), g.l, c.l), "cocg_vec_scale_powers");
  }
  // FFTProvider (plain.rs:369-400).  coset_g != nullptr fuses the distribute_powers_and_mul_by_const(g, 1) call that
  // follows ifft (precedes fft) in witness_map_from_matrices into the transform.
  void fft_in_place(FieldShareVec& v, const Domain& d, const Fr* coset_g = nullptr) { ntt(v, d, 0, coset_g); }
  void ifft_in_place(FieldShareVec& v, const Domain& d, const Fr* coset_g = nullptr) { ntt(v, d, 1, coset_g); }
  // the same transform of several share vectors as ONE launch sequence (the grid of a single 2^20 transform is 3.5 waves)
  void fft_many(const std::vector<FieldShareVec*>& vs, const Domain& d, const Fr* coset_g = nullptr) { ntt_many(vs, d, 0, coset_g); }
  void ifft_many(const std::vector<FieldShareVec*>& vs, const Domain& d, const Fr* coset_g = nullptr) { ntt_many(vs, d, 1, coset_g); }
  // MSMProvider (plain.rs:408-416)
  PointShare msm_public_points(int group, uint64_t bases, size_t off, size_t n, const FieldShareVec& scalars, size_t scalar_off = 0) {
    PointShare r;
    const void* sc[1] = {scalars.a.at(scalar_off)};
    check(ctx, cocg_msm(ctx, bases, off, n, sc, 1, 1, r.a.l), "cocg_msm");
    return r;
  }
  // several queries times the same scalars: one digit sort shared (cocg_msm_multi)
  std::vector<PointShare> msm_public_points_multi(const std::vector<int>& groups, const std::vector<uint64_t>& bases, const std::vector<size_t>& offs,
                                                  size_t n, const FieldShareVec& scalars, size_t scalar_off = 0, cocg_ctx* on = nullptr) {
    cocg_ctx* c = on ? on : ctx;  // `on`: run on the aux context (the handles in `bases` must be that context's)
    const int nq = (int)bases.size();
    std::vector<PointShare> r(nq);
    std::vector<void*> outs(nq);
    for (int q = 0; q < nq; q++) outs[q] = r[q].a.l;
    const void* sc[1] = {scalars.a.at(scalar_off)};
    check(c, cocg_msm_multi(c, bases.data(), offs.data(), nq, n, sc, 1, 1, outs.data()), "cocg_msm_multi");
    (void)groups;
    return r;
  }
  // EcMpcProtocol (plain.rs:287-357)
  void add_assign_points(int g, PointShare& a, const PointShare& b) { a.a = ec_add(g, a.a, b.a); }
  void sub_assign_points(int g, PointShare& a, const PointShare& b) { a.a = ec_sub(g, a.a, b.a); }
  void add_assign_points_public(int g, PointShare& a, const Point& b) { a.a = ec_add(g, a.a, b); }
  void add_assign_points_public_affine(int g, PointShare& a, const Point& b_aff) { a.a = ec_add(g, a.a, from_affine(g, b_aff)); }
  PointShare scalar_mul_public_point(int g, const Point& a, const FieldShare& b) { PointShare r; r.a = ec_mul(g, a, b.a); return r; }
  PointShare scalar_mul(int g, const PointShare& a, const FieldShare& b) { PointShare r; r.a = ec_mul(g, a.a, b.a); return r; }
  Point open_point(int, const PointShare& a) { return a.a; }
  std::pair<Point, Point> open_two_points(const PointShare& a, const PointShare& b) { return {a.a, b.a}; }

  // ---- CoPlonk (co-plonk/src/round*.rs) driver surface: plain.rs:111-285
  int pub_comp() const { return 0; }
  // several scalar vectors against the SAME bases in one call (the commitments of one Plonk round over p_tau): their bucket sets are
  // reduced together and the call synchronises once
  std::vector<PointShare> msm_public_points_many(int group, uint64_t bases, size_t n, const std::vector<const FieldShareVec*>& scalars) {
    const size_t m = scalars.size(), nl = 3 * group * lq;
    std::vector<const void*> sc(m);
    for (size_t j = 0; j < m; j++) sc[j] = scalars[j]->a.p;
    std::vector<uint64_t> packed(m * nl);
    check(ctx, cocg_msm(ctx, bases, 0, n, sc.data(), (int)m, 1, packed.data()), "cocg_msm");
    std::vector<PointShare> r(m);
    for (size_t j = 0; j < m; j++) memcpy(r[j].a.l, packed.data() + j * nl, nl * 8);
    return r;
  }
  std::vector<Point> open_point_many(int, const std::vector<PointShare>& a) { std::vector<Point> r; for (auto& x : a) r.push_back(x.a); return r; }
  std::vector<FieldShare> mul_many(const std::vector<FieldShare>& a, const std::vector<FieldShare>& b) {  // plain.rs:123-131
    std::vector<FieldShare> r(a.size());
    for (size_t i = 0; i < a.size(); i++) r[i] = mul(a[i], b[i]);
    return r;
  }
  std::vector<Fr> open_many(const std::vector<FieldShare>& a) { std::vector<Fr> r; for (auto& x : a) r.push_back(x.a); return r; }
  FieldShareVec alloc_share(size_t n) { return FieldShareVec{alloc(n), DevVec{}}; }
  FieldShareVec rand_vec(size_t n) {
    FieldShareVec o = alloc_share(n);
    check(ctx, cocg_prf_fill(ctx, seed, ctr++, o.a.p, n), "cocg_prf_fill");
    return o;
  }
  // several independent products: the plain driver has no network round to share
  std::vector<FieldShareVec> mul_vec_many(const std::vector<std::pair<const FieldShareVec*, const FieldShareVec*>>& ops) {
    std::vector<FieldShareVec> r;
    for (auto& o : ops) r.push_back(mul_vec(*o.first, *o.second));
    return r;
  }
  void release_many(std::vector<FieldShareVec>& v) { for (auto& x : v) release(x); v.clear(); }
  DevVec mul_open_many(const FieldShareVec& a, const FieldShareVec& b) { return pub_mul(a.a, b.a); }  // plain.rs:267-285
  FieldShareVec inv_many(const FieldShareVec& a) {  // plain.rs:141-152
    FieldShareVec o = alloc_share(a.len());
    check(ctx, cocg_d2d(ctx, o.a.p, a.a.p, a.len() * 32), "cocg_d2d");
    pub_inv(o.a);
    return o;
  }
  void mul_assign_public_vec(FieldShareVec& a, const DevVec& pub) { check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, a.a.p, pub.p, a.a.p, a.len()), "cocg_vec_op"); }
  // array_prod_mul! (round2.rs:17-42): prefix products; without a network the blinding by random r is pointless, the value is the same
  FieldShareVec array_prod_mul(const FieldShareVec& inp) {
    FieldShareVec o = alloc_share(inp.len());
    check(ctx, cocg_vec_scan(ctx, COCG_OP_MUL, inp.a.p, o.a.p, inp.len()), "cocg_vec_scan");
    return o;
  }
  FieldShare evaluate_poly_public(const FieldShareVec& poly, size_t n, const Fr& point) { return FieldShare{eval_public(poly.a, n, point), fr.zero()}; }
  // the party's additive output of a fused kernel becomes the share itself
  FieldShareVec reshare(DevVec local) { return FieldShareVec{local, DevVec{}}; }
  const uint8_t* seed_own() const { return seed; }
  const uint8_t* seed_prev() const { return seed; }
  uint32_t take_ctr(uint32_t k) { uint32_t c = ctr; ctr += k; return c; }
 private:
  void ntt(FieldShareVec& v, const Domain& d, int inverse, const Fr* coset_g) {
    if (v.len() != d.size()) throw Error("fft: vector length != domain size");
    void* vecs[1] = {v.a.p};
    check(ctx, cocg_ntt(ctx, vecs, 1, d.log_n, d.group_gen.l, inverse, coset_g ? coset_g->l : nullptr), "cocg_ntt");
  }
  void ntt_many(const std::vector<FieldShareVec*>& vs, const Domain& d, int inverse, const Fr* coset_g) {
    std::vector<void*> vecs;
    for (FieldShareVec* v : vs) {
      if (v->len() != d.size()) throw Error("fft: vector length != domain size");
      vecs.push_back(v->a.p);
    }
    check(ctx, cocg_ntt(ctx, vecs.data(), (int)vecs.size(), d.log_n, d.group_gen.l, inverse, coset_g ? coset_g->l : nullptr), "cocg_ntt");
  }
};

// ---------------------------------------------------------------- Rep3Protocol
class Rep3Protocol : public DeviceDriver {
 public:
  static constexpr int kComponents = 2;
  // Rep3Protocol::new: PRF setup = send own seed to next, receive prev's (rep3.rs:343-349)
  Rep3Protocol(int curve, int device, Rep3Network* network, const uint8_t own_seed[32]) : DeviceDriver(curve, device), net(network) {
    memcpy(seed1, own_seed, 32);
    net->send_next_bytes(seed1, 32);
  }
  void finish_setup() { net->recv_prev_bytes(seed2, 32); }  // split so three drivers can be built on one thread

  Rep3Network* net;
  DeviceBridge* bridge = nullptr;  // block-mode multi-GPU: share vectors to / from parties whose witness map runs on another rank
  InjectedRandomness* injected = nullptr;
  uint8_t seed1[32], seed2[32];
  uint32_t ctr = 0;  // advanced in lock-step by the three parties (every consumer below is called by all of them)
  int id() const { return net->get_id(); }

  // ---- PrimeFieldMpcProtocol
  FieldShare rand() {  // rep3.rs:595-598 -> Rep3Rand::random_fes
    if (injected && injected->i_rand < injected->rand.size()) return injected->rand[injected->i_rand++];
    FieldShare r;
    cocg_prf_field_host(fr.curve, seed1, ctr, 0, r.a.l);
    cocg_prf_field_host(fr.curve, seed2, ctr, 0, r.b.l);
    ctr++;
    return r;
  }
  Fr masking_field_element() {
    if (injected && injected->i_mul < injected->mul_masks.size()) return injected->mul_masks[injected->i_mul++];
    FieldShare r = rand();
    return fr.sub(r.a, r.b);
  }
  FieldShare mul(const FieldShare& a, const FieldShare& b) {  // rep3.rs:503-511
    Fr local_a = fr.add(fr.add(fr.mul(a.a, fr.add(b.a, b.b)), fr.mul(a.b, b.a)), masking_field_element());
    net->send_next_bytes(local_a.l, 32);
    FieldShare r;
    r.a = local_a;
    net->recv_prev_bytes(r.b.l, 32);
    return r;
  }
  FieldShareVec share_vec_from_host(const void* a, const void* b, size_t n) { return FieldShareVec{upload(a, n), upload(b, n)}; }

  // evaluate_constraint for every row of a matrix (rep3.rs:690-708): public terms enter party 0's `a` and party 1's `b`
  // only (add_with_public, rep3.rs:600-608)
  FieldShareVec evaluate_constraints(uint64_t csr, size_t rows, size_t out_len, const DevVec& public_inputs, const FieldShareVec& witness) {
    FieldShareVec o{alloc(out_len), alloc(out_len)};
    check(ctx, cocg_spmv(ctx, csr, id() == 0 ? public_inputs.p : nullptr, public_inputs.n, witness.a.p, o.a.p), "cocg_spmv");
    check(ctx, cocg_spmv(ctx, csr, id() == 1 ? public_inputs.p : nullptr, public_inputs.n, witness.b.p, o.b.p), "cocg_spmv");
    if (out_len > rows) {
      check(ctx, cocg_memset0(ctx, o.a.at(rows), (out_len - rows) * 32), "cocg_memset0");
      check(ctx, cocg_memset0(ctx, o.b.at(rows), (out_len - rows) * 32), "cocg_memset0");
    }
    return o;
  }
  FieldShareVec promote_to_trivial_shares(const DevVec& pub) {  // fieldshare.rs:254-283: (x,0) / (0,x) / (0,0)
    FieldShareVec o{alloc(pub.n), alloc(pub.n)};
    if (id() == 0) check(ctx, cocg_d2d(ctx, o.a.p, pub.p, pub.n * 32), "cocg_d2d");
    else check(ctx, cocg_memset0(ctx, o.a.p, pub.n * 32), "cocg_memset0");
    if (id() == 1) check(ctx, cocg_d2d(ctx, o.b.p, pub.p, pub.n * 32), "cocg_d2d");
    else check(ctx, cocg_memset0(ctx, o.b.p, pub.n * 32), "cocg_memset0");
    return o;
  }
  void clone_from_slice(FieldShareVec& dst, const FieldShareVec& src, size_t dst_off, size_t src_off, size_t len) {  // rep3.rs:710-725
    if (dst.len() < dst_off + len || src.len() < src_off + len || len == 0) throw Error("clone_from_slice: range");
    check(ctx, cocg_d2d(ctx, dst.a.at(dst_off), src.a.at(src_off), len * 32), "cocg_d2d");
    check(ctx, cocg_d2d(ctx, dst.b.at(dst_off), src.b.at(src_off), len * 32), "cocg_d2d");
  }
  // rep3.rs:650-670: local product + zero-mask on the GPU, then send_next_many / recv_prev_many over the host network
  FieldShareVec mul_vec(const FieldShareVec& a, const FieldShareVec& b) {
    size_t n = a.len();
    if (b.len() != n) throw Error("mul_vec: length mismatch");
    FieldShareVec o{alloc(n), DevVec{}};
    if (injected && injected->i_vec < injected->mul_vec_masks.size()) {
      DevVec m = upload(injected->mul_vec_masks[injected->i_vec++], n);
      check(ctx, cocg_rep3_mul_local(ctx, a.a.p, a.b.p, b.a.p, b.b.p, m.p, o.a.p, n), "cocg_rep3_mul_local");
      check(ctx, cocg_sync(ctx), "cocg_sync");
      release(m);
    } else {
      check(ctx, cocg_rep3_mul_local_prf(ctx, a.a.p, a.b.p, b.a.p, b.b.p, seed1, seed2, ctr++, o.a.p, n), "cocg_rep3_mul_local_prf");
    }
    o.b = exchange_next(o.a);  // send_next_many / recv_prev_many (in HBM, through pinned host memory, or across GPUs)
    return o;
  }
  void sub_assign_vec(FieldShareVec& a, const FieldShareVec& b) {  // rep3.rs:672-679
    check(ctx, cocg_vec_op(ctx, COCG_OP_SUB, a.a.p, b.a.p, a.a.p, a.len()), "cocg_vec_op");
    check(ctx, cocg_vec_op(ctx, COCG_OP_SUB, a.b.p, b.b.p, a.b.p, a.len()), "cocg_vec_op");
  }
  void distribute_powers_and_mul_by_const(FieldShareVec& v, const Fr& g, const Fr& c) {  // rep3.rs:681-688
    check(ctx, cocg_vec_scale_powers(ctx, v.a.p, v.len(), g.l, c.l), "cocg_vec_scale_powers");
    check(ctx, cocg_vec_scale_powers(ctx, v.b.p, v.len(), g.l, c.l), "cocg_vec_scale_powers");
  }
  // ---- FFTProvider (rep3.rs:880-921): both components in one launch sequence
  void fft_in_place(FieldShareVec& v, const Domain& d, const Fr* coset_g = nullptr) { ntt(v, d, 0, coset_g); }
  void ifft_in_place(FieldShareVec& v, const Domain& d, const Fr* coset_g = nullptr) { ntt(v, d, 1, coset_g); }
  void fft_many(const std::vector<FieldShareVec*>& vs, const Domain& d, const Fr* coset_g = nullptr) { ntt_many(vs, d, 0, coset_g); }
  void ifft_many(const std::vector<FieldShareVec*>& vs, const Domain& d, const Fr* coset_g = nullptr) { ntt_many(vs, d, 1, coset_g); }
  // ---- MSMProvider (rep3.rs:934-947): a and b components against the same resident bases
  PointShare msm_public_points(int group, uint64_t bases, size_t off, size_t n, const FieldShareVec& scalars, size_t scalar_off = 0) {
    Point out[2];
    const void* sc[2] = {scalars.a.at(scalar_off), scalars.b.at(scalar_off)};
    // Point is wider than a Jacobian: gather the two results from a packed buffer
    std::vector<uint64_t> packed(2 * 3 * group * lq);
    check(ctx, cocg_msm(ctx, bases, off, n, sc, 2, 1, packed.data()), "cocg_msm");
    memcpy(out[0].l, packed.data(), 3 * group * lq * 8);
    memcpy(out[1].l, packed.data() + 3 * group * lq, 3 * group * lq * 8);
    return PointShare{out[0], out[1]};
  }
  std::vector<PointShare> msm_public_points_multi(const std::vector<int>& groups, const std::vector<uint64_t>& bases, const std::vector<size_t>& offs,
                                                  size_t n, const FieldShareVec& scalars, size_t scalar_off = 0, cocg_ctx* on = nullptr) {
    cocg_ctx* c = on ? on : ctx;
    const int nq = (int)bases.size();
    std::vector<PointShare> r(nq);
    std::vector<std::vector<uint64_t>> packed(nq);
    std::vector<void*> outs(nq);
    for (int q = 0; q < nq; q++) {
      packed[q].resize(2 * 3 * groups[q] * lq);
      outs[q] = packed[q].data();
    }
    const void* sc[2] = {scalars.a.at(scalar_off), scalars.b.at(scalar_off)};
    check(c, cocg_msm_multi(c, bases.data(), offs.data(), nq, n, sc, 2, 1, outs.data()), "cocg_msm_multi");
    for (int q = 0; q < nq; q++) {
      const size_t nl = 3 * groups[q] * lq;
      memcpy(r[q].a.l, packed[q].data(), nl * 8);
      memcpy(r[q].b.l, packed[q].data() + nl, nl * 8);
    }
    return r;
  }
  // one share component of several queries (block mode: the {l, a, b_g1} bundle or the b_g2 MSM of one component on this rank)
  std::vector<Point> msm_public_points_multi_comp(const std::vector<int>& groups, const std::vector<uint64_t>& bases, const std::vector<size_t>& offs,
                                                  size_t n, const FieldShareVec& scalars, int comp, size_t scalar_off = 0) {
    const int nq = (int)bases.size();
    std::vector<Point> r(nq);
    std::vector<void*> outs(nq);
    for (int q = 0; q < nq; q++) outs[q] = r[q].l;
    const void* sc[1] = {comp == 0 ? scalars.a.at(scalar_off) : scalars.b.at(scalar_off)};
    check(ctx, cocg_msm_multi(ctx, bases.data(), offs.data(), nq, n, sc, 1, 1, outs.data()), "cocg_msm_multi");
    (void)groups;
    return r;
  }
  // ---- EcMpcProtocol (rep3.rs:769-862)
  void add_assign_points(int g, PointShare& a, const PointShare& b) { a.a = ec_add(g, a.a, b.a); a.b = ec_add(g, a.b, b.b); }
  void sub_assign_points(int g, PointShare& a, const PointShare& b) { a.a = ec_sub(g, a.a, b.a); a.b = ec_sub(g, a.b, b.b); }
  void add_assign_points_public(int g, PointShare& a, const Point& b) {
    if (id() == 0) a.a = ec_add(g, a.a, b);
    else if (id() == 1) a.b = ec_add(g, a.b, b);
  }
  void add_assign_points_public_affine(int g, PointShare& a, const Point& b_aff) { add_assign_points_public(g, a, from_affine(g, b_aff)); }
  // The host scalar multiplications of the proof assembly (0.2 ms each) sit on the critical path after the last MSM: independent ones
  // run on helper threads (cocg_ec_op is pure host arithmetic on its arguments).
  PointShare scalar_mul_public_point(int g, const Point& a, const FieldShare& b) {
    auto fb = std::async(std::launch::async, [&] { return ec_mul(g, a, b.b); });
    Point pa = ec_mul(g, a, b.a);
    return PointShare{pa, fb.get()};
  }
  Point masking_ec_element(int g) {  // rngs.rs:48-57; a PRF scalar times the generator instead of C::rand
    if (injected && injected->i_ec < injected->ec_masks.size()) return injected->ec_masks[injected->i_ec++];
    FieldShare r = rand();
    return ec_mul(g, generator(g), fr.sub(r.a, r.b));
  }
  PointShare scalar_mul(int g, const PointShare& a, const FieldShare& b) {  // rep3.rs:835-847, pointshare.rs Mul
    const Point mask_base = generator(g);
    const bool mask_injected = injected && injected->i_ec < injected->ec_masks.size();
    Fr mask_scalar = fr.zero();
    if (!mask_injected) {  // masking_ec_element's PRF draw stays on this thread (the counter is not shared); only the multiplication moves
      FieldShare rr = rand();
      mask_scalar = fr.sub(rr.a, rr.b);
    }
    auto f1 = std::async(std::launch::async, [&] { return ec_mul(g, a.b, b.a); });
    auto f2 = std::async(std::launch::async, [&] { return mask_injected ? Point{} : ec_mul(g, mask_base, mask_scalar); });
    Point local_a = ec_add(g, ec_mul(g, a.a, fr.add(b.a, b.b)), f1.get());
    const Point mask_computed = f2.get();
    local_a = ec_add(g, local_a, mask_injected ? masking_ec_element(g) : mask_computed);
    size_t nb = 3 * g * lq * 8;
    net->send_next_bytes(local_a.l, nb);
    PointShare r;
    r.a = local_a;
    net->recv_prev_bytes(r.b.l, nb);
    return r;
  }
  Point open_point(int g, const PointShare& a) {  // rep3.rs:849-853
    size_t nb = 3 * g * lq * 8;
    net->send_next_bytes(a.b.l, nb);
    Point c;
    net->recv_prev_bytes(c.l, nb);
    return ec_add(g, ec_add(g, a.a, a.b), c);
  }
  std::pair<Point, Point> open_two_points(const PointShare& a, const PointShare& b) {  // rep3.rs:864-878
    size_t n1 = 3 * lq * 8, n2 = 6 * lq * 8;
    std::vector<uint8_t> buf(n1 + n2);
    memcpy(buf.data(), a.b.l, n1);
    memcpy(buf.data() + n1, b.b.l, n2);
    net->send_next_bytes(buf.data(), buf.size());
    net->recv_prev_bytes(buf.data(), buf.size());
    Point r1, r2;
    memcpy(r1.l, buf.data(), n1);
    memcpy(r2.l, buf.data() + n1, n2);
    return {ec_add(1, r1, ec_add(1, a.a, a.b)), ec_add(2, r2, ec_add(2, b.a, b.b))};
  }

  // ---- CoPlonk (co-plonk/src/round*.rs) driver surface: rep3.rs:503-760
  int pub_comp() const { return id() == 0 ? 0 : id() == 1 ? 1 : -1; }
  // several share vectors against the SAME bases in one call (k = 2 components each, at most 4 vectors): one launch sequence, one sync
  std::vector<PointShare> msm_public_points_many(int group, uint64_t bases, size_t n, const std::vector<const FieldShareVec*>& scalars) {
    const size_t m = scalars.size(), nl = 3 * group * lq;
    if (2 * m > 8) throw Error("msm_public_points_many: at most 4 share vectors per call");
    std::vector<const void*> sc(2 * m);
    for (size_t j = 0; j < m; j++) { sc[2 * j] = scalars[j]->a.p; sc[2 * j + 1] = scalars[j]->b.p; }
    std::vector<uint64_t> packed(2 * m * nl);
    check(ctx, cocg_msm(ctx, bases, 0, n, sc.data(), (int)(2 * m), 1, packed.data()), "cocg_msm");
    std::vector<PointShare> r(m);
    for (size_t j = 0; j < m; j++) {
      memcpy(r[j].a.l, packed.data() + 2 * j * nl, nl * 8);
      memcpy(r[j].b.l, packed.data() + (2 * j + 1) * nl, nl * 8);
    }
    return r;
  }
  std::vector<Point> open_point_many(int g, const std::vector<PointShare>& a) {  // rep3.rs:855-861: one message for all of them
    const size_t nb = 3 * g * lq * 8;
    std::vector<uint8_t> snd(a.size() * nb), rcv(a.size() * nb);
    for (size_t j = 0; j < a.size(); j++) memcpy(snd.data() + j * nb, a[j].b.l, nb);
    net->send_next_bytes(snd.data(), snd.size());
    net->recv_prev_bytes(rcv.data(), rcv.size());
    std::vector<Point> r(a.size());
    for (size_t j = 0; j < a.size(); j++) {
      Point c;
      memcpy(c.l, rcv.data() + j * nb, nb);
      r[j] = ec_add(g, ec_add(g, a[j].a, a[j].b), c);
    }
    return r;
  }
  const uint8_t* seed_own() const { return seed1; }
  const uint8_t* seed_prev() const { return seed2; }
  uint32_t take_ctr(uint32_t k) { uint32_t c = ctr; ctr += k; return c; }
  FieldShareVec alloc_share(size_t n) { return FieldShareVec{alloc(n), alloc(n)}; }
  // mul_many on a handful of scalars (rep3.rs:513-528): one message for all of them
  std::vector<FieldShare> mul_many(const std::vector<FieldShare>& a, const std::vector<FieldShare>& b) {
    std::vector<Fr> loc(a.size()), rcv(a.size());
    for (size_t i = 0; i < a.size(); i++)
      loc[i] = fr.add(fr.add(fr.mul(a[i].a, fr.add(b[i].a, b[i].b)), fr.mul(a[i].b, b[i].a)), masking_field_element());
    net->send_next_bytes(loc.data(), loc.size() * 32);
    net->recv_prev_bytes(rcv.data(), rcv.size() * 32);
    std::vector<FieldShare> r(a.size());
    for (size_t i = 0; i < a.size(); i++) r[i] = FieldShare{loc[i], rcv[i]};
    return r;
  }
  std::vector<Fr> open_many(const std::vector<FieldShare>& a) {  // rep3.rs:620-628
    std::vector<Fr> bs(a.size()), cs(a.size());
    for (size_t i = 0; i < a.size(); i++) bs[i] = a[i].b;
    net->send_next_bytes(bs.data(), bs.size() * 32);
    net->recv_prev_bytes(cs.data(), cs.size() * 32);
    for (size_t i = 0; i < a.size(); i++) cs[i] = fr.add(cs[i], fr.add(a[i].a, a[i].b));
    return cs;
  }
  FieldShareVec rand_vec(size_t n) {  // n x rand(): (F(seed1), F(seed2)) under one vector counter
    FieldShareVec o = alloc_share(n);
    check(ctx, cocg_prf_fill(ctx, seed1, ctr, o.a.p, n), "cocg_prf_fill");
    check(ctx, cocg_prf_fill(ctx, seed2, ctr, o.b.p, n), "cocg_prf_fill");
    ctr++;
    return o;
  }
  // A mul_vec that another rank executes for this party (block mode): keep the PRF counter / injected masks in lock-step
  void skip_mul_vec(int k) {
    if (injected && injected->i_vec < injected->mul_vec_masks.size()) injected->i_vec += k;
    else ctr += k;
  }
  // send `local` to the next party, receive the previous party's vector of the same length (send_next_many / recv_prev_many)
  DevVec exchange_next(const DevVec& local) {
    const size_t n = local.n;
    if (bridge && !(bridge->local((id() + 1) % 3) && bridge->local((id() + 2) % 3))) {
      const int nxt = (id() + 1) % 3, prv = (id() + 2) % 3;
      check(ctx, cocg_sync(ctx), "cocg_sync");  // the payload is complete before anyone reads it
      if (bridge->local(nxt)) {  // the neighbour's witness map runs on this GPU: hand a copy over in HBM
        DevVec copy = alloc(n);
        check(ctx, cocg_d2d(ctx, copy.p, local.p, n * 32), "cocg_d2d");
        check(ctx, cocg_sync(ctx), "cocg_sync");
        Message s;
        s.bytes = n * 32;
        s.device = copy.p;
        net->send_next(std::move(s));
      } else {
        bridge->post(CommOp{0, bridge->owner(nxt), local.p, n * 32});
      }
      DevVec r;
      if (bridge->local(prv)) {
        Message m = net->recv_prev();
        if (m.bytes != n * 32 || !m.device) throw Error("During execution of mul_vec in MPC: Invalid number of elements received");
        r = DevVec{m.device, n};
      } else {
        r = alloc(n);
        bridge->post(CommOp{1, bridge->owner(prv), r.p, n * 32});
      }
      bridge->flush();
      return r;
    }
    if (net->device_exchange()) {
      DevVec copy = alloc(n);
      check(ctx, cocg_d2d(ctx, copy.p, local.p, n * 32), "cocg_d2d");
      check(ctx, cocg_sync(ctx), "cocg_sync");
      Message s;
      s.bytes = n * 32;
      s.device = copy.p;
      net->send_next(std::move(s));
      Message m = net->recv_prev();
      if (m.bytes != n * 32 || !m.device) throw Error("During execution of mul_vec in MPC: Invalid number of elements received");
      return DevVec{m.device, n};
    }
    std::shared_ptr<void> buf = pinned(n * 32);
    check(ctx, cocg_d2h(ctx, buf.get(), local.p, n * 32), "cocg_d2h");
    net->send_next(Message{buf, n * 32});
    buf.reset();
    Message m = net->recv_prev();
    if (m.bytes != n * 32) throw Error("During execution of mul_vec in MPC: Invalid number of elements received");
    DevVec r = alloc(n);
    check(ctx, cocg_h2d(ctx, r.p, m.data.get(), n * 32), "cocg_h2d");
    return r;
  }
  FieldShareVec reshare(DevVec local) { DevVec b = exchange_next(local); return FieldShareVec{local, b}; }
  // Several independent products in ONE network round: the local steps write slices of one buffer, one message carries them all.
  // Results are views into two pooled buffers; release_many() returns those.
  std::vector<FieldShareVec> mul_vec_many(const std::vector<std::pair<const FieldShareVec*, const FieldShareVec*>>& ops) {
    size_t total = 0;
    for (auto& o : ops) { if (o.first->len() != o.second->len()) throw Error("mul_vec: length mismatch"); total += o.first->len(); }
    DevVec loc = alloc(total);
    size_t off = 0;
    for (auto& o : ops) {
      const size_t n = o.first->len();
      check(ctx, cocg_rep3_mul_local_prf(ctx, o.first->a.p, o.first->b.p, o.second->a.p, o.second->b.p, seed1, seed2, ctr++, loc.at(off), n), "cocg_rep3_mul_local_prf");
      off += n;
    }
    DevVec rcv = exchange_next(loc);
    std::vector<FieldShareVec> r;
    off = 0;
    for (auto& o : ops) {
      const size_t n = o.first->len();
      r.push_back(FieldShareVec{slice(loc, off, n), slice(rcv, off, n)});
      off += n;
    }
    many_owned_.push_back({r.empty() ? nullptr : r[0].a.p, loc, rcv});
    return r;
  }
  void release_many(std::vector<FieldShareVec>& v) {
    if (!v.empty())
      for (size_t k = 0; k < many_owned_.size(); k++)
        if (many_owned_[k].key == v[0].a.p) {
          release(many_owned_[k].loc);
          release(many_owned_[k].rcv);
          many_owned_.erase(many_owned_.begin() + k);
          break;
        }
    v.clear();
  }
  // mul_open_many (rep3.rs:738-757): local product + mask to BOTH other parties, sum of the three
  DevVec mul_open_many(const FieldShareVec& a, const FieldShareVec& b) {
    const size_t n = a.len();
    DevVec loc = alloc(n);
    check(ctx, cocg_rep3_mul_local_prf(ctx, a.a.p, a.b.p, b.a.p, b.b.p, seed1, seed2, ctr++, loc.p, n), "cocg_rep3_mul_local_prf");
    DevVec from_prev, from_next;
    const int nxt = (id() + 1) % 3, prv = (id() + 2) % 3;
    if (net->device_exchange()) {
      for (int k = 0; k < 2; k++) {
        DevVec copy = alloc(n);
        check(ctx, cocg_d2d(ctx, copy.p, loc.p, n * 32), "cocg_d2d");
        check(ctx, cocg_sync(ctx), "cocg_sync");
        Message s;
        s.bytes = n * 32;
        s.device = copy.p;
        net->send(k == 0 ? nxt : prv, std::move(s));
      }
      Message m1 = net->recv(prv), m2 = net->recv(nxt);
      if (m1.bytes != n * 32 || m2.bytes != n * 32 || !m1.device || !m2.device) throw Error("mul_open_many: Invalid number of elements received");
      from_prev = DevVec{m1.device, n};
      from_next = DevVec{m2.device, n};
    } else {
      std::shared_ptr<void> buf = pinned(n * 32);
      check(ctx, cocg_d2h(ctx, buf.get(), loc.p, n * 32), "cocg_d2h");
      net->send(nxt, Message{buf, n * 32});
      net->send(prv, Message{buf, n * 32});
      buf.reset();
      Message m1 = net->recv(prv), m2 = net->recv(nxt);
      if (m1.bytes != n * 32 || m2.bytes != n * 32) throw Error("mul_open_many: Invalid number of elements received");
      from_prev = alloc(n);
      from_next = alloc(n);
      check(ctx, cocg_h2d(ctx, from_prev.p, m1.data.get(), n * 32), "cocg_h2d");
      check(ctx, cocg_h2d(ctx, from_next.p, m2.data.get(), n * 32), "cocg_h2d");
    }
    check(ctx, cocg_vec_op(ctx, COCG_OP_ADD, loc.p, from_prev.p, loc.p, n), "cocg_vec_op");
    check(ctx, cocg_vec_op(ctx, COCG_OP_ADD, loc.p, from_next.p, loc.p, n), "cocg_vec_op");
    release(from_prev);
    release(from_next);
    return loc;
  }
  void mul_assign_public_vec(FieldShareVec& a, const DevVec& pub) {  // mul_with_public per element
    check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, a.a.p, pub.p, a.a.p, a.len()), "cocg_vec_op");
    check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, a.b.p, pub.p, a.b.p, a.len()), "cocg_vec_op");
  }
  FieldShareVec inv_many(const FieldShareVec& a) {  // rep3.rs:544-558: r random, y = open(a * r), a^-1 = r / y
    FieldShareVec r = rand_vec(a.len());
    DevVec y = mul_open_many(a, r);
    pub_inv(y);
    mul_assign_public_vec(r, y);
    release(y);
    return r;
  }
  // array_prod_mul! (round2.rs:17-42, after Ozdemir-Boneh): prefix products of a shared vector in a constant number of rounds
  FieldShareVec array_prod_mul(const FieldShareVec& inp) {
    const size_t len = inp.len();
    FieldShareVec r = rand_vec(len + 1);
    FieldShareVec r_inv = inv_many(r);
    Fr h[2];
    check(ctx, cocg_d2h(ctx, h[0].l, r_inv.a.p, 32), "cocg_d2h");
    check(ctx, cocg_d2h(ctx, h[1].l, r_inv.b.p, 32), "cocg_d2h");
    FieldShareVec r_inv0 = alloc_share(len);  // vec![r_inv[0]; len]
    check(ctx, cocg_vec_fill(ctx, r_inv0.a.p, len, h[0].l), "cocg_vec_fill");
    check(ctx, cocg_vec_fill(ctx, r_inv0.b.p, len, h[1].l), "cocg_vec_fill");
    const FieldShareVec r_tail = slice(r, 1, len), r_head = slice(r, 0, len), r_inv_tail = slice(r_inv, 1, len);
    std::vector<FieldShareVec> pr = mul_vec_many({{&r_inv0, &r_tail}, {&r_head, &inp}});  // unblind | mul
    DevVec open = mul_open_many(pr[1], r_inv_tail);
    pub_scan_mul(open);
    FieldShareVec unblind = alloc_share(len);
    check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, pr[0].a.p, open.p, unblind.a.p, len), "cocg_vec_op");
    check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, pr[0].b.p, open.p, unblind.b.p, len), "cocg_vec_op");
    release(open);
    release_many(pr);
    release(r_inv0);
    release(r);
    release(r_inv);
    return unblind;
  }
  FieldShare evaluate_poly_public(const FieldShareVec& poly, size_t n, const Fr& point) {  // rep3.rs:923-928
    return FieldShare{eval_public(poly.a, n, point), eval_public(poly.b, n, point)};
  }
 private:
  struct ManyOwned { void* key; DevVec loc, rcv; };
  std::vector<ManyOwned> many_owned_;
  void ntt(FieldShareVec& v, const Domain& d, int inverse, const Fr* coset_g) {
    if (v.len() != d.size()) throw Error("fft: vector length != domain size");
    void* vecs[2] = {v.a.p, v.b.p};
    check(ctx, cocg_ntt(ctx, vecs, 2, d.log_n, d.group_gen.l, inverse, coset_g ? coset_g->l : nullptr), "cocg_ntt");
  }
  void ntt_many(const std::vector<FieldShareVec*>& vs, const Domain& d, int inverse, const Fr* coset_g) {
    std::vector<void*> vecs;
    for (FieldShareVec* v : vs) {
      if (v->len() != d.size()) throw Error("fft: vector length != domain size");
      vecs.push_back(v->a.p);
      vecs.push_back(v->b.p);
    }
    check(ctx, cocg_ntt(ctx, vecs.data(), (int)vecs.size(), d.log_n, d.group_gen.l, inverse, coset_g ? coset_g->l : nullptr), "cocg_ntt");
  }
};

}  // namespace cohost
