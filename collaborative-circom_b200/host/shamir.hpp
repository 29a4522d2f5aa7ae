// ShamirProtocol: the semi-honest n-party Shamir driver over the cocg C ABI, mirror of
// /root/reference/mpc-core/src/protocols/shamir.rs (struct + new :196-245, degree_reduce* :252-438, PrimeFieldMpcProtocol
// :459-712, EcMpcProtocol / PairingEcMpcProtocol :714-824, FFTProvider :826-871, MSMProvider :1027-1039, ShamirRng :873-1025)
// and shamir/shamir_core.rs:8-118 (share, lagrange_from_coeff, reconstruct).  Party i evaluates at the point i + 1; the king
// (party 0) interpolates the degree-2t products and re-shares them with degree t.  A share vector is ONE array in HBM
// (ShamirPrimeFieldShareVec{a}, shamir/fieldshare.rs:152-155); every vector step -- local product, masking with the
// double-random pair, the king's Lagrange combination, polynomial re-sharing, the Vandermonde extraction of the
// preprocessing -- is a cocg kernel (cocg_vec_op, cocg_vec_axpy, cocg_prf_fill); only the messages cross PCIe.
// Differences from the reference, none visible in opened values: random polynomial coefficients come from the counter-addressed
// ChaCha12 PRF (csrc/prf.cuh) instead of a sequential ChaCha12 stream; the (t+1) x amount pairs of one preprocessing batch are
// laid out row-major instead of interleaved; a batch is sized to the vector that needs it instead of 1024.
#pragma once
#include <array>

#include "driver.hpp"

namespace cohost {

class ShamirNetwork {  // shamir/network.rs:14-59
 public:
  virtual ~ShamirNetwork() {}
  virtual int get_id() const = 0;
  virtual int get_num_parties() const = 0;
  virtual void send(int target, Message m) = 0;
  virtual Message recv(int from) = 0;
  // true when all parties live in one process on ONE GPU and the caller opted in: share vectors are then handed over as HBM buffers
  // instead of being staged through pinned host memory (the default: what a party that has to reach a NIC does)
  virtual bool device_exchange() const { return false; }
  void send_bytes(int target, const void* p, size_t n) {
    std::shared_ptr<void> buf(new uint8_t[n ? n : 1], [](void* q) { delete[] (uint8_t*)q; });
    memcpy(buf.get(), p, n);
    send(target, Message{buf, n});
  }
  void recv_bytes(int from, void* p, size_t n) {
    Message m = recv(from);
    if (m.bytes != n) throw Error("network: invalid number of bytes received");
    memcpy(p, m.data.get(), n);
  }
  // send to the next num-1 parties, receive from the previous num-1; result[0] = own data, result[r] = from party id - r
  std::vector<std::vector<uint8_t>> broadcast_next(const void* data, size_t n, int num) {  // network.rs:233-270
    const int id = get_id(), np = get_num_parties();
    for (int s = 1; s < num; s++) send_bytes((id + s) % np, data, n);
    std::vector<std::vector<uint8_t>> res(num, std::vector<uint8_t>(n));
    memcpy(res[0].data(), data, n);
    for (int r = 1; r < num; r++) recv_bytes((id + np - r) % np, res[r].data(), n);
    return res;
  }
};

class ShamirTestNetwork;  // in-process: one channel per ordered pair (tests/src/shamir_network.rs)
class ShamirPartyTestNetwork : public ShamirNetwork {
 public:
  ShamirPartyTestNetwork(ShamirTestNetwork* net, int id, int n) : net_(net), id_(id), n_(n) {}
  int get_id() const override { return id_; }
  int get_num_parties() const override { return n_; }
  void send(int target, Message m) override;
  Message recv(int from) override;
  bool device_exchange() const override;

 private:
  ShamirTestNetwork* net_;
  int id_, n_;
};
class ShamirTestNetwork {
 public:
  explicit ShamirTestNetwork(int n) : n_(n), ch_((size_t)n * n) {
    for (int i = 0; i < n; i++) parties_.emplace_back(new ShamirPartyTestNetwork(this, i, n));
  }
  ShamirPartyTestNetwork* party(int i) { return parties_[i].get(); }
  Channel& chan(int from, int to) { return ch_[(size_t)from * n_ + to]; }
  void close_all() { for (auto& c : ch_) c.close(); }
  bool device_exchange = false;

 private:
  int n_;
  std::vector<Channel> ch_;
  std::vector<std::unique_ptr<ShamirPartyTestNetwork>> parties_;
};
inline void ShamirPartyTestNetwork::send(int target, Message m) { net_->chan(id_, target).send(std::move(m)); }
inline Message ShamirPartyTestNetwork::recv(int from) { return net_->chan(from, id_).recv(); }
inline bool ShamirPartyTestNetwork::device_exchange() const { return net_->device_exchange; }

class ShamirProtocol : public DeviceDriver {
 public:
  static constexpr int kComponents = 1;
  static constexpr int KING_ID = 0;

  ShamirProtocol(int curve, int device, int threshold, ShamirNetwork* network, const uint8_t own_seed[32])
      : DeviceDriver(curve, device), threshold(threshold), net(network) {
    const int np = net->get_num_parties(), id = net->get_id();
    if (2 * threshold + 1 > np) throw Error("Threshold too large for number of parties");  // shamir.rs:213-215
    memcpy(seed, own_seed, 32);
    std::vector<int> pts;
    for (int i = 0; i <= threshold; i++) pts.push_back((id + np - i) % np + 1);  // we send in circles: receive from the previous parties
    open_lagrange_t = lagrange_from_coeff(pts);
    pts.clear();
    for (int i = 0; i <= 2 * threshold; i++) pts.push_back((id + np - i) % np + 1);  // open of a degree-2t sharing (mul_open, shamir.rs:676-712)
    open_lagrange_2t = lagrange_from_coeff(pts);
    pts.clear();
    for (int i = 1; i <= 2 * threshold + 1; i++) pts.push_back(i);
    mul_lagrange_2t = lagrange_from_coeff(pts);
  }
  ~ShamirProtocol() override {
    for (auto& b : batches_) { release(b.r_t); release(b.r_2t); }
  }

  int threshold;
  ShamirNetwork* net;
  uint8_t seed[32];
  uint32_t ctr = 0;
  std::vector<Fr> open_lagrange_t, open_lagrange_2t, mul_lagrange_2t;
  int id() const { return net->get_id(); }

  // ---- ShamirCore (shamir/shamir_core.rs)
  Fr from_u64(uint64_t v) const {  // Montgomery form of a small integer: v * R = v * one
    Fr acc = fr.zero(), base = fr.one();
    for (; v; v >>= 1) { if (v & 1) acc = fr.add(acc, base); base = fr.add(base, base); }
    return acc;
  }
  Fr inverse(const Fr& a) const { return fr.un(a, [](auto x) { return cocg::fp_inv(x); }); }
  std::vector<Fr> lagrange_from_coeff(const std::vector<int>& coeffs) const {  // shamir_core.rs:53-72
    std::vector<Fr> res;
    for (int i : coeffs) {
      Fr num = fr.one(), den = fr.one(), i_ = from_u64(i);
      for (int j : coeffs)
        if (i != j) { Fr j_ = from_u64(j); num = fr.mul(num, j_); den = fr.mul(den, fr.sub(j_, i_)); }
      res.push_back(fr.mul(num, inverse(den)));
    }
    return res;
  }

  // ---- double-random preprocessing (ShamirRng, shamir.rs:873-1025): pairs (r_t, r_2t) sharing the same random value
  struct Batch { DevVec r_t, r_2t; size_t used = 0; };
  // share(secret vector, degree): out[p] = secret + sum_k coeff_k * (p + 1)^k, coefficients from the PRF (shamir_core.rs:8-33)
  std::vector<DevVec> share_vec(const DevVec& secret, int degree) {
    const int np = net->get_num_parties();
    const size_t n = secret.n;
    std::vector<DevVec> out(np);
    for (int p = 0; p < np; p++) {
      out[p] = alloc(n);
      check(ctx, cocg_d2d(ctx, out[p].p, secret.p, n * 32), "cocg_d2d");
    }
    DevVec coeff = alloc(n);
    for (int k = 1; k <= degree; k++) {
      check(ctx, cocg_prf_fill(ctx, seed, ctr++, coeff.p, n), "cocg_prf_fill");
      for (int p = 0; p < np; p++) {
        Fr xp = fr.one(), x = from_u64(p + 1);
        for (int q = 0; q < k; q++) xp = fr.mul(xp, x);
        check(ctx, cocg_vec_axpy(ctx, xp.l, coeff.p, out[p].p, out[p].p, n), "cocg_vec_axpy");
      }
    }
    release(coeff);
    return out;
  }
  void send_vec(int target, const DevVec& v) {
    if (net->device_exchange()) {  // co-located parties: the receiver adopts a copy in HBM
      DevVec copy = alloc(v.n);
      check(ctx, cocg_d2d(ctx, copy.p, v.p, v.n * 32), "cocg_d2d");
      check(ctx, cocg_sync(ctx), "cocg_sync");
      Message m;
      m.bytes = v.n * 32;
      m.device = copy.p;
      net->send(target, std::move(m));
      return;
    }
    std::shared_ptr<void> buf = pinned(v.n * 32);
    check(ctx, cocg_d2h(ctx, buf.get(), v.p, v.n * 32), "cocg_d2h");
    net->send(target, Message{buf, v.n * 32});
  }
  DevVec recv_vec(int from, size_t n, const char* what) {
    Message m = net->recv(from);
    if (m.bytes != n * 32) throw Error(std::string("During execution of ") + what + " in MPC: Invalid number of elements received");
    if (m.device) return DevVec{m.device, n};
    return upload(m.data.get(), n);
  }
  // buffer_triples (shamir.rs:925-1010): every party shares `amount` random values with degree t and 2t, all parties apply the
  // (t+1) x n Vandermonde matrix to what they received -> (t+1) * amount pairs
  void buffer_triples(size_t amount) {
    const int np = net->get_num_parties(), me = id(), t = threshold;
    DevVec rnd = alloc(amount);
    check(ctx, cocg_prf_fill(ctx, seed, ctr++, rnd.p, amount), "cocg_prf_fill");
    std::vector<DevVec> st = share_vec(rnd, t), s2t = share_vec(rnd, 2 * t);
    release(rnd);
    for (int p = 0; p < np; p++)
      if (p != me) { send_vec(p, st[p]); send_vec(p, s2t[p]); }
    std::vector<DevVec> rt(np), r2t(np);
    for (int p = 0; p < np; p++) {
      if (p == me) { rt[p] = st[p]; r2t[p] = s2t[p]; continue; }
      rt[p] = recv_vec(p, amount, "buffer_triples");
      r2t[p] = recv_vec(p, amount, "buffer_triples");
      release(st[p]);
      release(s2t[p]);
    }
    Batch b;
    b.r_t = alloc(amount * (t + 1));
    b.r_2t = alloc(amount * (t + 1));
    for (int row = 0; row <= t; row++) {  // vandermonde_mul, shamir.rs:904-920: res[row] = sum_j (j + 1)^row * in[j]
      for (int j = 0; j < np; j++) {
        Fr w = fr.one(), x = from_u64(j + 1);
        for (int q = 0; q < row; q++) w = fr.mul(w, x);
        void* dt = b.r_t.at(row * amount);
        void* d2 = b.r_2t.at(row * amount);
        check(ctx, cocg_vec_axpy(ctx, w.l, rt[j].p, j ? dt : nullptr, dt, amount), "cocg_vec_axpy");
        check(ctx, cocg_vec_axpy(ctx, w.l, r2t[j].p, j ? d2 : nullptr, d2, amount), "cocg_vec_axpy");
      }
    }
    for (int p = 0; p < np; p++) { release(rt[p]); release(r2t[p]); }
    batches_.push_back(b);
  }
  // n consecutive pairs as device views (valid until the batch is recycled at the next preprocess of the same size)
  std::pair<DevVec, DevVec> get_pairs(size_t n) {
    for (auto& b : batches_)
      if (b.r_t.n - b.used >= n) {
        std::pair<DevVec, DevVec> r{DevVec{b.r_t.at(b.used), n}, DevVec{b.r_2t.at(b.used), n}};
        b.used += n;
        return r;
      }
    for (auto& b : batches_) { release(b.r_t); release(b.r_2t); }  // none of the buffered batches can serve n pairs: recycle them
    batches_.clear();
    buffer_triples((n + threshold) / (threshold + 1));
    return get_pairs(n);
  }
  std::pair<Fr, Fr> get_pair() {  // shamir.rs:1012-1025; single pairs come from a small host-side pool
    if (host_pool_.empty()) {
      constexpr size_t kPool = 64;
      auto pr = get_pairs(kPool);
      std::vector<Fr> a(kPool), b(kPool);
      check(ctx, cocg_d2h(ctx, a.data(), pr.first.p, kPool * 32), "cocg_d2h");
      check(ctx, cocg_d2h(ctx, b.data(), pr.second.p, kPool * 32), "cocg_d2h");
      for (size_t i = 0; i < kPool; i++) host_pool_.push_back({a[i], b[i]});
    }
    auto r = host_pool_.back();
    host_pool_.pop_back();
    return r;
  }
  void preprocess(size_t amount) { buffer_triples(amount); }  // shamir.rs:247-250

  // ---- degree reduction (shamir.rs:252-384)
  FieldShare degree_reduce(Fr input) {
    auto pr = get_pair();
    input = fr.add(input, pr.second);
    const int np = net->get_num_parties(), me = id();
    Fr my_share;
    if (me == KING_ID) {
      Fr acc = fr.zero();
      for (int other = 0; other <= 2 * threshold; other++) {
        Fr v = input;
        if (other != KING_ID) net->recv_bytes(other, v.l, 32);
        acc = fr.add(acc, fr.mul(v, mul_lagrange_2t[other]));
      }
      std::vector<Fr> coeffs(threshold);  // ShamirCore::share
      for (auto& c : coeffs) cocg_prf_field_host(fr.curve, seed, ctr++, 0, c.l);
      for (int p = 0; p < np; p++) {
        Fr share = acc, x = from_u64(p + 1), xp = x;
        for (auto& c : coeffs) { share = fr.add(share, fr.mul(xp, c)); xp = fr.mul(xp, x); }
        if (p == me) my_share = share; else net->send_bytes(p, share.l, 32);
      }
    } else {
      if (me <= 2 * threshold) net->send_bytes(KING_ID, input.l, 32);  // only send if my items are required
      net->recv_bytes(KING_ID, my_share.l, 32);
    }
    return FieldShare{fr.sub(my_share, pr.first), fr.zero()};
  }
  FieldShareVec degree_reduce_vec(DevVec inputs) {  // takes ownership of `inputs`
    const size_t n = inputs.n;
    auto pr = get_pairs(n);
    check(ctx, cocg_vec_op(ctx, COCG_OP_ADD, inputs.p, pr.second.p, inputs.p, n), "cocg_vec_op");
    const int np = net->get_num_parties(), me = id();
    DevVec mine;
    if (me == KING_ID) {
      DevVec acc = alloc(n);
      for (int other = 0; other <= 2 * threshold; other++) {
        DevVec v = other == KING_ID ? inputs : recv_vec(other, n, "degree_reduce_vec");
        check(ctx, cocg_vec_axpy(ctx, mul_lagrange_2t[other].l, v.p, other ? acc.p : nullptr, acc.p, n), "cocg_vec_axpy");
        if (other != KING_ID) release(v);
      }
      std::vector<DevVec> shares = share_vec(acc, threshold);
      release(acc);
      for (int p = 0; p < np; p++) {
        if (p == me) { mine = shares[p]; continue; }
        send_vec(p, shares[p]);
        release(shares[p]);
      }
    } else {
      if (me <= 2 * threshold) send_vec(KING_ID, inputs);
      mine = recv_vec(KING_ID, n, "degree_reduce_vec");
    }
    release(inputs);
    check(ctx, cocg_vec_op(ctx, COCG_OP_SUB, mine.p, pr.first.p, mine.p, n), "cocg_vec_op");
    return FieldShareVec{mine, DevVec{}};
  }
  Point degree_reduce_point(int g, Point input) {  // shamir.rs:386-438
    auto pr = get_pair();
    Point gen = generator(g);
    Point r_t = ec_mul(g, gen, pr.first), r_2t = ec_mul(g, gen, pr.second);
    input = ec_add(g, input, r_2t);
    const int np = net->get_num_parties(), me = id();
    const size_t nb = 3 * g * lq * 8;
    Point my_share;
    if (me == KING_ID) {
      Point acc = infinity(g);
      for (int other = 0; other <= 2 * threshold; other++) {
        Point v = input;
        if (other != KING_ID) net->recv_bytes(other, v.l, nb);
        acc = ec_add(g, acc, ec_mul(g, v, mul_lagrange_2t[other]));
      }
      std::vector<Point> coeffs(threshold);  // share_point: random points = PRF scalars times the generator
      for (auto& c : coeffs) { Fr k; cocg_prf_field_host(fr.curve, seed, ctr++, 0, k.l); c = ec_mul(g, gen, k); }
      for (int p = 0; p < np; p++) {
        Point share = acc;
        Fr x = from_u64(p + 1), xp = x;
        for (auto& c : coeffs) { share = ec_add(g, share, ec_mul(g, c, xp)); xp = fr.mul(xp, x); }
        if (p == me) my_share = share; else net->send_bytes(p, share.l, nb);
      }
    } else {
      if (me <= 2 * threshold) net->send_bytes(KING_ID, input.l, nb);
      net->recv_bytes(KING_ID, my_share.l, nb);
    }
    return ec_sub(g, my_share, r_t);
  }

  // ---- PrimeFieldMpcProtocol (shamir.rs:459-712)
  FieldShare rand() { return FieldShare{get_pair().first, fr.zero()}; }                                   // :570-573
  FieldShare mul(const FieldShare& a, const FieldShare& b) { return degree_reduce(fr.mul(a.a, b.a)); }  // :482-485
  FieldShareVec share_vec_from_host(const void* a, const void*, size_t n) { return FieldShareVec{upload(a, n), DevVec{}}; }
  FieldShareVec evaluate_constraints(uint64_t csr, size_t rows, size_t out_len, const DevVec& public_inputs, const FieldShareVec& witness) {
    FieldShareVec o{alloc(out_len), DevVec{}};  // :642-660, add_with_public: every party adds the public terms (:471-473)
    check(ctx, cocg_spmv(ctx, csr, public_inputs.p, public_inputs.n, witness.a.p, o.a.p), "cocg_spmv");
    if (out_len > rows) check(ctx, cocg_memset0(ctx, o.a.at(rows), (out_len - rows) * 32), "cocg_memset0");
    return o;
  }
  FieldShareVec promote_to_trivial_shares(const DevVec& pub) {  // :629-632
    FieldShareVec o{alloc(pub.n), DevVec{}};
    check(ctx, cocg_d2d(ctx, o.a.p, pub.p, pub.n * 32), "cocg_d2d");
    return o;
  }
  void clone_from_slice(FieldShareVec& dst, const FieldShareVec& src, size_t dst_off, size_t src_off, size_t len) {  // :662-674
    if (dst.len() < dst_off + len || src.len() < src_off + len || len == 0) throw Error("clone_from_slice: range");
    check(ctx, cocg_d2d(ctx, dst.a.at(dst_off), src.a.at(src_off), len * 32), "cocg_d2d");
  }
  FieldShareVec mul_vec(const FieldShareVec& a, const FieldShareVec& b) {  // :609-623
    if (a.len() != b.len()) throw Error("mul_vec: length mismatch");
    DevVec muls = alloc(a.len());
    check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, a.a.p, b.a.p, muls.p, a.len()), "cocg_vec_op");
    return degree_reduce_vec(muls);
  }
  void sub_assign_vec(FieldShareVec& a, const FieldShareVec& b) { check(ctx, cocg_vec_op(ctx, COCG_OP_SUB, a.a.p, b.a.p, a.a.p, a.len()), "cocg_vec_op"); }
  void distribute_powers_and_mul_by_const(FieldShareVec& v, const Fr& g, const Fr& c) {  // :634-640
    check(ctx, cocg_vec_scale_powers(ctx, v.a.p, v.len(), g.l, c.l), "cocg_vec_scale_powers");
  }
  // ---- FFTProvider (:826-871), MSMProvider (:1027-1039)
  void fft_in_place(FieldShareVec& v, const Domain& d, const Fr* coset_g = nullptr) { ntt(v, d, 0, coset_g); }
  void ifft_in_place(FieldShareVec& v, const Domain& d, const Fr* coset_g = nullptr) { ntt(v, d, 1, coset_g); }
  void fft_many(const std::vector<FieldShareVec*>& vs, const Domain& d, const Fr* coset_g = nullptr) { ntt_many(vs, d, 0, coset_g); }
  void ifft_many(const std::vector<FieldShareVec*>& vs, const Domain& d, const Fr* coset_g = nullptr) { ntt_many(vs, d, 1, coset_g); }
  PointShare msm_public_points(int, uint64_t bases, size_t off, size_t n, const FieldShareVec& scalars, size_t scalar_off = 0) {
    PointShare r;
    const void* sc[1] = {scalars.a.at(scalar_off)};
    check(ctx, cocg_msm(ctx, bases, off, n, sc, 1, 1, r.a.l), "cocg_msm");
    return r;
  }
  std::vector<PointShare> msm_public_points_multi(const std::vector<int>&, const std::vector<uint64_t>& bases, const std::vector<size_t>& offs, size_t n,
                                                  const FieldShareVec& scalars, size_t scalar_off = 0, cocg_ctx* on = nullptr) {
    cocg_ctx* c = on ? on : ctx;
    const int nq = (int)bases.size();
    std::vector<PointShare> r(nq);
    std::vector<void*> outs(nq);
    for (int q = 0; q < nq; q++) outs[q] = r[q].a.l;
    const void* sc[1] = {scalars.a.at(scalar_off)};
    check(c, cocg_msm_multi(c, bases.data(), offs.data(), nq, n, sc, 1, 1, outs.data()), "cocg_msm_multi");
    return r;
  }
  // ---- EcMpcProtocol (:714-806): public points are added by every party
  void add_assign_points(int g, PointShare& a, const PointShare& b) { a.a = ec_add(g, a.a, b.a); }
  void sub_assign_points(int g, PointShare& a, const PointShare& b) { a.a = ec_sub(g, a.a, b.a); }
  void add_assign_points_public(int g, PointShare& a, const Point& b) { a.a = ec_add(g, a.a, b); }
  void add_assign_points_public_affine(int g, PointShare& a, const Point& b_aff) { a.a = ec_add(g, a.a, from_affine(g, b_aff)); }
  PointShare scalar_mul_public_point(int g, const Point& a, const FieldShare& b) { PointShare r; r.a = ec_mul(g, a, b.a); return r; }
  PointShare scalar_mul(int g, const PointShare& a, const FieldShare& b) {  // :769-776
    PointShare r;
    r.a = degree_reduce_point(g, ec_mul(g, a.a, b.a));
    return r;
  }
  Point reconstruct_point(int g, const std::vector<std::vector<uint8_t>>& rcv, size_t off, size_t nb) {  // shamir_core.rs:106-117
    Point res = infinity(g);
    for (size_t i = 0; i < rcv.size(); i++) {
      Point s;
      memcpy(s.l, rcv[i].data() + off, nb);
      res = ec_add(g, res, ec_mul(g, s, open_lagrange_t[i]));
    }
    return res;
  }
  Point open_point(int g, const PointShare& a) {  // :778-782
    const size_t nb = 3 * g * lq * 8;
    auto rcv = net->broadcast_next(a.a.l, nb, threshold + 1);
    return reconstruct_point(g, rcv, 0, nb);
  }
  std::pair<Point, Point> open_two_points(const PointShare& a, const PointShare& b) {  // :808-823
    const size_t n1 = 3 * lq * 8, n2 = 6 * lq * 8;
    std::vector<uint8_t> buf(n1 + n2);
    memcpy(buf.data(), a.a.l, n1);
    memcpy(buf.data() + n1, b.a.l, n2);
    auto rcv = net->broadcast_next(buf.data(), buf.size(), threshold + 1);
    return {reconstruct_point(1, rcv, 0, n1), reconstruct_point(2, rcv, n1, n2)};
  }
  Fr open(const FieldShare& a) {  // :575-579
    auto rcv = net->broadcast_next(a.a.l, 32, threshold + 1);
    Fr res = fr.zero();
    for (size_t i = 0; i < rcv.size(); i++) {
      Fr s;
      memcpy(s.l, rcv[i].data(), 32);
      res = fr.add(res, fr.mul(s, open_lagrange_t[i]));
    }
    return res;
  }

  // ---- CoPlonk (co-plonk/src/round*.rs) driver surface: shamir.rs:459-712, 865-871
  int pub_comp() const { return 0; }  // add_with_public: every party adds the constant (shamir.rs:471-473)
  const uint8_t* seed_own() const { return seed; }
  const uint8_t* seed_prev() const { return seed; }
  uint32_t take_ctr(uint32_t k) { uint32_t c = ctr; ctr += k; return c; }
  FieldShareVec alloc_share(size_t n) { return FieldShareVec{alloc(n), DevVec{}}; }
  std::vector<FieldShare> mul_many(const std::vector<FieldShare>& a, const std::vector<FieldShare>& b) {  // :490-502
    std::vector<Fr> prod(a.size());
    for (size_t i = 0; i < a.size(); i++) prod[i] = fr.mul(a[i].a, b[i].a);
    FieldShareVec red = degree_reduce_vec(upload(prod.data(), prod.size()));
    download(red.a, prod.data());
    release(red);
    std::vector<FieldShare> r(a.size());
    for (size_t i = 0; i < a.size(); i++) r[i] = FieldShare{prod[i], fr.zero()};
    return r;
  }
  std::vector<Fr> open_many(const std::vector<FieldShare>& a) {  // :581-607
    std::vector<Fr> mine(a.size());
    for (size_t i = 0; i < a.size(); i++) mine[i] = a[i].a;
    auto rcv = net->broadcast_next(mine.data(), mine.size() * 32, threshold + 1);
    std::vector<Fr> res(a.size(), fr.zero());
    for (size_t j = 0; j < rcv.size(); j++)
      for (size_t i = 0; i < a.size(); i++) {
        Fr s;
        memcpy(s.l, rcv[j].data() + 32 * i, 32);
        res[i] = fr.add(res[i], fr.mul(s, open_lagrange_t[j]));
      }
    return res;
  }
  FieldShareVec rand_vec(size_t n) {  // n x rand(): the degree-t halves of fresh double-random pairs (:570-573)
    auto pr = get_pairs(n);
    FieldShareVec o = alloc_share(n);
    check(ctx, cocg_d2d(ctx, o.a.p, pr.first.p, n * 32), "cocg_d2d");
    return o;
  }
  FieldShareVec reshare(DevVec local) { return degree_reduce_vec(local); }  // a degree-2t local result back to degree t
  // several independent products, ONE degree reduction (one king round) for all of them
  std::vector<FieldShareVec> mul_vec_many(const std::vector<std::pair<const FieldShareVec*, const FieldShareVec*>>& ops) {
    size_t total = 0;
    for (auto& o : ops) { if (o.first->len() != o.second->len()) throw Error("mul_vec: length mismatch"); total += o.first->len(); }
    DevVec loc = alloc(total);
    size_t off = 0;
    for (auto& o : ops) {
      check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, o.first->a.p, o.second->a.p, loc.at(off), o.first->len()), "cocg_vec_op");
      off += o.first->len();
    }
    FieldShareVec red = degree_reduce_vec(loc);
    std::vector<FieldShareVec> r;
    off = 0;
    for (auto& o : ops) {
      r.push_back(FieldShareVec{slice(red.a, off, o.first->len()), DevVec{}});
      off += o.first->len();
    }
    many_owned_.push_back({r.empty() ? nullptr : r[0].a.p, red.a});
    return r;
  }
  void release_many(std::vector<FieldShareVec>& v) {
    if (!v.empty())
      for (size_t k = 0; k < many_owned_.size(); k++)
        if (many_owned_[k].first == v[0].a.p) {
          release(many_owned_[k].second);
          many_owned_.erase(many_owned_.begin() + k);
          break;
        }
    v.clear();
  }
  // mul_open_many (:684-712): the degree-2t products go to the next 2t parties, each reconstructs from its own and the 2t it receives
  DevVec mul_open_many(const FieldShareVec& a, const FieldShareVec& b) {
    const size_t n = a.len();
    const int np = net->get_num_parties(), me = id(), cnt = 2 * threshold + 1;
    DevVec mul = alloc(n);
    check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, a.a.p, b.a.p, mul.p, n), "cocg_vec_op");
    for (int s = 1; s < cnt; s++) send_vec((me + s) % np, mul);
    DevVec res = alloc(n);
    check(ctx, cocg_vec_axpy(ctx, open_lagrange_2t[0].l, mul.p, nullptr, res.p, n), "cocg_vec_axpy");
    for (int r = 1; r < cnt; r++) {
      DevVec v = recv_vec((me + np - r) % np, n, "mul_open_many");
      check(ctx, cocg_vec_axpy(ctx, open_lagrange_2t[r].l, v.p, res.p, res.p, n), "cocg_vec_axpy");
      release(v);
    }
    release(mul);
    return res;
  }
  void mul_assign_public_vec(FieldShareVec& a, const DevVec& pub) { check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, a.a.p, pub.p, a.a.p, a.len()), "cocg_vec_op"); }
  FieldShareVec inv_many(const FieldShareVec& a) {  // :521-535
    FieldShareVec r = rand_vec(a.len());
    DevVec y = mul_open_many(a, r);
    pub_inv(y);
    mul_assign_public_vec(r, y);
    release(y);
    return r;
  }
  FieldShareVec array_prod_mul(const FieldShareVec& inp) {  // round2.rs:17-42
    const size_t len = inp.len();
    FieldShareVec r = rand_vec(len + 1);
    FieldShareVec r_inv = inv_many(r);
    Fr h;
    check(ctx, cocg_d2h(ctx, h.l, r_inv.a.p, 32), "cocg_d2h");
    FieldShareVec r_inv0 = alloc_share(len);
    check(ctx, cocg_vec_fill(ctx, r_inv0.a.p, len, h.l), "cocg_vec_fill");
    const FieldShareVec r_tail = slice(r, 1, len), r_head = slice(r, 0, len), r_inv_tail = slice(r_inv, 1, len);
    std::vector<FieldShareVec> pr = mul_vec_many({{&r_inv0, &r_tail}, {&r_head, &inp}});
    DevVec open = mul_open_many(pr[1], r_inv_tail);
    pub_scan_mul(open);
    FieldShareVec unblind = alloc_share(len);
    check(ctx, cocg_vec_op(ctx, COCG_OP_MUL, pr[0].a.p, open.p, unblind.a.p, len), "cocg_vec_op");
    release(open);
    release_many(pr);
    release(r_inv0);
    release(r);
    release(r_inv);
    return unblind;
  }
  FieldShare evaluate_poly_public(const FieldShareVec& poly, size_t n, const Fr& point) { return FieldShare{eval_public(poly.a, n, point), fr.zero()}; }  // :865-871
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
  std::vector<Point> open_point_many(int g, const std::vector<PointShare>& a) {  // :784-806
    const size_t nb = 3 * g * lq * 8;
    std::vector<uint8_t> buf(a.size() * nb);
    for (size_t j = 0; j < a.size(); j++) memcpy(buf.data() + j * nb, a[j].a.l, nb);
    auto rcv = net->broadcast_next(buf.data(), buf.size(), threshold + 1);
    std::vector<Point> r(a.size());
    for (size_t j = 0; j < a.size(); j++) r[j] = reconstruct_point(g, rcv, j * nb, nb);
    return r;
  }

 private:
  std::vector<std::pair<void*, DevVec>> many_owned_;
  void ntt_many(const std::vector<FieldShareVec*>& vs, const Domain& d, int inverse, const Fr* coset_g) {
    std::vector<void*> vecs;
    for (FieldShareVec* v : vs) {
      if (v->len() != d.size()) throw Error("fft: vector length != domain size");
      vecs.push_back(v->a.p);
    }
    check(ctx, cocg_ntt(ctx, vecs.data(), (int)vecs.size(), d.log_n, d.group_gen.l, inverse, coset_g ? coset_g->l : nullptr), "cocg_ntt");
  }
  void ntt(FieldShareVec& v, const Domain& d, int inverse, const Fr* coset_g) {
    if (v.len() != d.size()) throw Error("fft: vector length != domain size");
    void* vecs[1] = {v.a.p};
    check(ctx, cocg_ntt(ctx, vecs, 1, d.log_n, d.group_gen.l, inverse, coset_g ? coset_g->l : nullptr), "cocg_ntt");
  }
  std::vector<Batch> batches_;
  std::vector<std::pair<Fr, Fr>> host_pool_;
};

}  // namespace cohost
