// CoGroth16<T>: the collaborative Groth16 prover, generic over the MPC driver, as in the reference
// (/root/reference/co-circom/co-groth16/src/groth16.rs:80-326).  Same call sequence, same names; the driver methods
// enqueue sm_100a kernels, so between two MPC network rounds everything stays in HBM:
//   prove                           groth16.rs:113-139
//   witness_map_from_matrices       groth16.rs:141-204   (SpMV -> mul_vec -> 3x [iNTT, coset scale, NTT] -> mul_vec -> sub)
//   calculate_coeff                 groth16.rs:206-235
//   create_proof_with_assignment    groth16.rs:237-326   (5 MSM sites + O(1) group operations + 3 openings)
// Differences from the reference, all result-preserving:
//   * evaluate_constraint is called once per matrix (a CSR SpMV) instead of once per row;
//   * distribute_powers_and_mul_by_const(g, 1) is folded into the ifft that precedes it (the driver's coset_g argument);
//   * the five secret-scalar MSMs are issued before the first opening (they depend on h and the witness only), so a
//     multi-GPU caller can combine all partial sums with ONE all-gather per proof (`MsmCombine`).
#pragma once
#include <chrono>
#include <functional>
#include <future>

#include "driver.hpp"

namespace cohost {

// Groth16 domain generator and coset shift with the snarkjs root convention
// (/root/reference/co-circom/co-circom-snarks/src/lib.rs:208-221 roots_of_unity; groth16.rs:57-77).
struct Groth16Roots {
  Fr omega, coset;
};
inline Groth16Roots root_of_unity_for_groth16(int curve, size_t pow) {
  using namespace cocg;
  Groth16Roots out;
  auto run = [&](auto tag, int two_adicity) {
    using F = decltype(tag);
    constexpr int N = F::N;
    // smallest quadratic non-residue q: q^((r-1)/2) == -1
    uint32_t e[N], t[N];
    for (int i = 0; i < N; i++) e[i] = F::Params::mod(i);
    e[0] -= 1;                                    // r - 1 (r is odd)
    for (int i = 0; i < N; i++) t[i] = e[i];
    auto shr = [&](uint32_t* v, int k) {
      for (int s = 0; s < k; s++) {
        for (int i = 0; i < N - 1; i++) v[i] = (v[i] >> 1) | (v[i + 1] << 31);
        v[N - 1] >>= 1;
      }
    };
    auto pow_big = [&](F b, const uint32_t* ex) {
      F r = F::one();
      for (int i = 32 * N - 1; i >= 0; i--) {
        r = fp_sqr(r);
        if ((ex[i >> 5] >> (i & 31)) & 1) r = fp_mul(r, b);
      }
      return r;
    };
    uint32_t half[N];
    for (int i = 0; i < N; i++) half[i] = e[i];
    shr(half, 1);
    F minus_one = fp_neg(F::one());
    F q = F::one();
    F one = F::one();
    for (;;) {
      if (pow_big(q, half) == minus_one) break;
      q = fp_add(q, one);
    }
    shr(t, two_adicity);                          // odd part of r - 1
    F z = pow_big(q, t);                          // order 2^s
    // roots[k] = z^(2^(s-k))
    auto root = [&](size_t k) {
      F r = z;
      for (size_t i = 0; i < (size_t)two_adicity - k; i++) r = fp_sqr(r);
      return r;
    };
    if (pow > (size_t)two_adicity) throw Error("domain larger than the field's two-adicity");
    F om = root(pow);
    F cs = (pow == (size_t)two_adicity) ? fp_sqr(q) : root(pow + 1);
    memcpy(out.omega.l, om.l, 32);
    memcpy(out.coset.l, cs.l, 32);
  };
  if (curve == COCG_BN254) run(Bn254Fr{}, 28);
  else run(Bls381Fr{}, 32);
  return out;
}

// Hook for the multi-GPU path: receives this rank's partial MSM results of one proof (packed PointShares) and must
// return them summed over ranks (one all-gather + local fold).  Null = single GPU.
struct MsmPartials {
  PointShare h_acc, l_acc, a_acc, b1_acc, b2_acc;  // G1, G1, G1, G1, G2
};
using MsmCombine = std::function<void(int party, MsmPartials&)>;

// Which slice of every query this rank accumulates (index-range sharding of the bases).
struct MsmShard {
  int rank = 0, world = 1;
  void range(size_t n, size_t& off, size_t& len) const {
    size_t per = (n + world - 1) / world;
    off = std::min(n, per * rank);
    len = std::min(n - off, per);
  }
};

template <class T>
class CoGroth16 {
 public:
  explicit CoGroth16(T& driver) : driver(driver) {}
  T& driver;
  MsmShard shard;
  MsmCombine combine;
  // block mode (types.hpp BlockPlan): this rank runs whole blocks of the proof instead of an index range of every MSM
  const BlockPlan* blocks = nullptr;
  int block_rank = 0;
  FieldShareVec last_h;  // kept for parity tests (released by the next prove)
  FieldShare last_r, last_s;
  // host wall-clock of the last prove per phase, seconds (the reference logs "Proof generation took {} ms", co-circom.rs:503-506;
  // here split so that multi-GPU scaling can be read): witness map | MSMs incl. their syncs | wait for the all-gather | assembly
  double phase_s[4] = {0, 0, 0, 0};
  static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

  // bases/CSR handles of `zkey` valid in driver.ctx (aliases made by the session)
  struct Handles {
    uint64_t a_query, b_g1_query, b_g2_query, h_query, l_query, csr_a, csr_b;
  };
  // Single GPU, whole queries resident: handles of the same queries in driver.aux_ctx.  When set, the four MSMs over the witness
  // (they do not depend on h) start on the aux context as soon as the witness is resident and run WHILE the witness map is computed.
  const Handles* aux_hd = nullptr;
  using AuxMsms = std::future<std::vector<PointShare>>;

  // public_inputs: num_inputs = l + 1 elements (leading 1 included) in HBM; witness: n_aux share elements
  Groth16Proof prove(const ZKey& zkey, const Handles& hd, const DevVec& public_inputs, const std::vector<Fr>& public_inputs_host,
                     const FieldShareVec& private_witness) {
    driver.release(last_h);
    const double t0 = now();
    // r, s and r*s are drawn FIRST (the reference draws them after the witness map, groth16.rs:128-131; they are independent of it):
    // every group operation of create_proof_with_assignment that depends only on them and on the key -- 8 G1 and 3 G2 scalar
    // multiplications of delta and of the public-input part of the queries, ~2 ms of host arithmetic -- then runs on a helper thread
    // WHILE the GPU computes the witness map and the MSMs, instead of after them
    FieldShare r = driver.rand();
    FieldShare s = driver.rand();
    last_r = r;
    last_s = s;
    FieldShare rs = driver.mul(r, s);
    std::future<Precomputed> pre = std::async(std::launch::async, [this, &zkey, r, s, rs, &public_inputs_host] { return precompute(zkey, r, s, rs, public_inputs_host); });
    AuxMsms aux;
    if (aux_hd && driver.aux_ctx && !blocks && shard.world == 1 && zkey.world == 1) {
      const size_t l = zkey.n_public, n_aux = zkey.n_aux();
      aux = std::async(std::launch::async, [this, l, n_aux, &private_witness] {
        return driver.msm_public_points_multi({1, 1, 1, 2}, {aux_hd->l_query, aux_hd->a_query, aux_hd->b_g1_query, aux_hd->b_g2_query},
                                              {0, 1 + l, 1 + l, 1 + l}, n_aux, private_witness, 0, driver.aux_ctx);
      });
    }
    FieldShareVec h;
    try {
      if (!blocks || blocks->wm[party_id()] == block_rank) h = witness_map_from_matrices(zkey, hd, public_inputs, private_witness);
      else skip_witness_map();  // another rank runs this party's witness map: stay in step with its randomness
      check(driver.ctx, cocg_sync(driver.ctx), "cocg_sync");
    } catch (...) {
      pre.wait();
      if (aux.valid()) aux.wait();
      throw;
    }
    phase_s[0] = now() - t0;
    Groth16Proof p = create_proof_with_assignment(zkey, hd, r, s, h, public_inputs_host, private_witness, &pre, &aux);
    last_h = h;
    return p;
  }

  FieldShareVec witness_map_from_matrices(const ZKey& zkey, const Handles& hd, const DevVec& public_inputs, const FieldShareVec& private_witness) {
    const size_t num_constraints = zkey.num_constraints, num_inputs = zkey.num_inputs();
    Domain domain;
    domain.log_n = (unsigned)zkey.pow;
    if (num_constraints + num_inputs > domain.size()) throw Error("PolynomialDegreeTooLarge");
    Groth16Roots roots = root_of_unity_for_groth16(zkey.curve, zkey.pow);
    domain.group_gen = roots.omega;
    const Fr root_of_unity = roots.coset;
    const size_t domain_size = domain.size();

    FieldShareVec a = driver.evaluate_constraints(hd.csr_a, num_constraints, domain_size, public_inputs, private_witness);
    FieldShareVec b = driver.evaluate_constraints(hd.csr_b, num_constraints, domain_size, public_inputs, private_witness);
    FieldShareVec promoted_public = driver.promote_to_trivial_shares(public_inputs);
    driver.clone_from_slice(a, promoted_public, num_constraints, 0, num_inputs);
    driver.release(promoted_public);

    // The three [ifft, coset scale, fft] pipelines (groth16.rs:177-199) are independent: a, b and c (every share component) go
    // through each transform as ONE launch sequence.  The order of the two mul_vec rounds -- and with it every PRF counter -- is
    // the reference's, so the proof bytes do not depend on this batching.
    FieldShareVec c = driver.mul_vec(a, b);
    driver.ifft_many({&a, &b, &c}, domain, &root_of_unity);  // + distribute_powers_and_mul_by_const(., root_of_unity, 1)
    driver.fft_many({&a, &b, &c}, domain);
    FieldShareVec ab = driver.mul_vec(a, b);
    driver.release(a);
    driver.release(b);
    driver.sub_assign_vec(ab, c);
    driver.release(c);
    return ab;
  }

  // pub_acc = msm_unchecked(query[1..=l], input_assignment): l is tiny, done on the host
  Point public_acc(int g, const std::vector<Point>& query_head, const std::vector<Fr>& input_assignment) {
    Point pub_acc = driver.infinity(g);
    for (size_t i = 0; i < input_assignment.size(); i++)
      pub_acc = driver.ec_add(g, pub_acc, driver.ec_mul(g, driver.from_affine(g, query_head[1 + i]), input_assignment[i]));
    return pub_acc;
  }
  // what create_proof_with_assignment needs from (r, s, r*s, the key, the public inputs) alone: host arithmetic, no device, no network
  struct Precomputed {
    Point delta_g1, delta_g2, pub_a, pub_b1, pub_b2;
    PointShare r_s_delta_g1, r_g1, s_g1, s_g2;
  };
  Precomputed precompute(const ZKey& zkey, const FieldShare& r, const FieldShare& s, const FieldShare& rs, const std::vector<Fr>& public_inputs_host) {
    std::vector<Fr> input_assignment(public_inputs_host.begin() + 1, public_inputs_host.end());
    Precomputed p;
    p.delta_g1 = driver.from_affine(1, zkey.delta_g1);
    p.delta_g2 = driver.from_affine(2, zkey.delta_g2);
    p.r_s_delta_g1 = driver.scalar_mul_public_point(1, p.delta_g1, rs);
    p.r_g1 = driver.scalar_mul_public_point(1, p.delta_g1, r);
    p.s_g1 = driver.scalar_mul_public_point(1, p.delta_g1, s);
    p.s_g2 = driver.scalar_mul_public_point(2, p.delta_g2, s);
    p.pub_a = public_acc(1, zkey.a_head, input_assignment);
    p.pub_b1 = public_acc(1, zkey.b_g1_head, input_assignment);
    p.pub_b2 = public_acc(2, zkey.b_g2_head, input_assignment);
    return p;
  }
  PointShare calculate_coeff(int g, const PointShare& initial, const std::vector<Point>& query_head, const Point& vk_param, const Point& pub_acc,
                             const PointShare& priv_acc) {
    PointShare res = initial;
    driver.add_assign_points_public_affine(g, res, query_head[0]);
    driver.add_assign_points_public_affine(g, res, vk_param);
    driver.add_assign_points_public(g, res, pub_acc);
    driver.add_assign_points(g, res, priv_acc);
    return res;
  }

  Groth16Proof create_proof_with_assignment(const ZKey& zkey, const Handles& hd, const FieldShare& r, const FieldShare& s, const FieldShareVec& h,
                                            const std::vector<Fr>& public_inputs_host, const FieldShareVec& aux_assignment,
                                            std::future<Precomputed>* pre_async = nullptr, AuxMsms* aux_async = nullptr) {
    const double t_start = now();
    const size_t l = zkey.n_public, n_aux = zkey.n_aux();
    // ---- all secret-scalar MSMs first (msm_public_points at groth16.rs:248, 251-255 and inside calculate_coeff :221-225)
    MsmPartials m;
    size_t off, len;
    if (blocks) {
      block_msms(zkey, hd, h, aux_assignment, m);
    } else if (aux_async && aux_async->valid()) {  // the four witness MSMs have been running on the aux context since before the witness map
      try {
        m.h_acc = driver.msm_public_points(1, hd.h_query, 0, std::min(h.len(), zkey.domain_size()), h, 0);
      } catch (...) {
        aux_async->wait();
        throw;
      }
      std::vector<PointShare> r4 = aux_async->get();
      m.l_acc = r4[0];
      m.a_acc = r4[1];
      m.b1_acc = r4[2];
      m.b2_acc = r4[3];
    } else {
    if (zkey.world != 1 && (zkey.world != shard.world || zkey.rank != shard.rank)) throw Error("the zkey holds another rank's shard of the queries");
    shard.range(std::min(h.len(), zkey.domain_size()), off, len);
    m.h_acc = driver.msm_public_points(1, hd.h_query, off - zkey.h_first, len, h, off);
    shard.range(n_aux, off, len);
    {  // the four queries multiplied by aux_assignment share one digit sort per share component
      std::vector<PointShare> r = driver.msm_public_points_multi(
          {1, 1, 1, 2}, {hd.l_query, hd.a_query, hd.b_g1_query, hd.b_g2_query},
          {off - zkey.l_first, 1 + l + off - zkey.a_first, 1 + l + off - zkey.b_g1_first, 1 + l + off - zkey.b_g2_first}, len, aux_assignment, off);
      m.l_acc = r[0];
      m.a_acc = r[1];
      m.b1_acc = r[2];
      m.b2_acc = r[3];
    }
    }
    const double t_msm = now();
    phase_s[1] = t_msm - t_start;
    if (combine) combine(party_id(), m);
    const double t_comb = now();
    phase_s[2] = t_comb - t_msm;

    Precomputed pc;
    if (pre_async) pc = pre_async->get();
    else pc = precompute(zkey, r, s, driver.mul(r, s), public_inputs_host);

    PointShare g_a = calculate_coeff(1, pc.r_g1, zkey.a_head, zkey.alpha_g1, pc.pub_a, m.a_acc);
    Point g_a_opened = driver.open_point(1, g_a);
    PointShare s_g_a = driver.scalar_mul_public_point(1, g_a_opened, s);

    PointShare g1_b = calculate_coeff(1, pc.s_g1, zkey.b_g1_head, zkey.beta_g1, pc.pub_b1, m.b1_acc);
    PointShare r_g1_b = driver.scalar_mul(1, g1_b, r);

    PointShare g2_b = calculate_coeff(2, pc.s_g2, zkey.b_g2_head, zkey.beta_g2, pc.pub_b2, m.b2_acc);

    PointShare g_c = s_g_a;
    driver.add_assign_points(1, g_c, r_g1_b);
    driver.sub_assign_points(1, g_c, pc.r_s_delta_g1);
    driver.add_assign_points(1, g_c, m.l_acc);
    driver.add_assign_points(1, g_c, m.h_acc);

    auto opened = driver.open_two_points(g_c, g2_b);
    Groth16Proof proof;
    proof.pi_a = driver.to_affine(1, g_a_opened);
    proof.pi_b = driver.to_affine(2, opened.second);
    proof.pi_c = driver.to_affine(1, opened.first);
    phase_s[3] = now() - t_comb;
    return proof;
  }

 private:
  // ---- block mode helpers (REP3 only: the plan speaks of two share components)
  template <class U = T>
  auto skip_witness_map_impl(int) -> decltype(std::declval<U&>().skip_mul_vec(2)) { driver.skip_mul_vec(2); }
  void skip_witness_map_impl(long) {}
  void skip_witness_map() { skip_witness_map_impl(0); }
  template <class U = T>
  auto block_msms_impl(int, const ZKey& zkey, const Handles& hd, const FieldShareVec& h, const FieldShareVec& aux, MsmPartials& m)
      -> decltype(std::declval<U&>().msm_public_points_multi_comp({}, {}, {}, 0, aux, 0), void()) {
    const int p = party_id();
    const size_t l = zkey.n_public, n_aux = zkey.n_aux();
    const Point inf1 = driver.infinity(1), inf2 = driver.infinity(2);
    m.h_acc = m.l_acc = m.a_acc = m.b1_acc = PointShare{inf1, inf1};
    m.b2_acc = PointShare{inf2, inf2};
    if (blocks->wm[p] == block_rank) m.h_acc = driver.msm_public_points(1, hd.h_query, 0, std::min(h.len(), zkey.domain_size()), h, 0);
    for (int c = 0; c < 2; c++) {  // the MSMs of one share component that run here share one digit sort
      std::vector<int> groups, which;
      std::vector<uint64_t> bases;
      std::vector<size_t> offs;
      const uint64_t handle[3] = {hd.l_query, hd.a_query, hd.b_g1_query};
      for (int k = 0; k < 3; k++)
        if (blocks->g1[p][c][k] == block_rank) { groups.push_back(1); bases.push_back(handle[k]); offs.push_back(k == 0 ? 0 : 1 + l); which.push_back(k); }
      if (blocks->g2[p][c] == block_rank) { groups.push_back(2); bases.push_back(hd.b_g2_query); offs.push_back(1 + l); which.push_back(3); }
      if (groups.empty()) continue;
      std::vector<Point> r = driver.msm_public_points_multi_comp(groups, bases, offs, n_aux, aux, c);
      for (size_t i = 0; i < which.size(); i++) {
        PointShare& dst = which[i] == 0 ? m.l_acc : which[i] == 1 ? m.a_acc : which[i] == 2 ? m.b1_acc : m.b2_acc;
        (c ? dst.b : dst.a) = r[i];
      }
    }
  }
  void block_msms_impl(long, const ZKey&, const Handles&, const FieldShareVec&, const FieldShareVec&, MsmPartials&) {
    throw Error("block-mode multi-GPU proving is implemented for the REP3 driver");
  }
  void block_msms(const ZKey& zkey, const Handles& hd, const FieldShareVec& h, const FieldShareVec& aux, MsmPartials& m) { block_msms_impl(0, zkey, hd, h, aux, m); }
  template <class U = T>
  auto party_id_impl(int) -> decltype(std::declval<U&>().id()) { return driver.id(); }
  int party_id_impl(long) { return 0; }
  int party_id() { return party_id_impl(0); }
};

}  // namespace cohost
