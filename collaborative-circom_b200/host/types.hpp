// Host-side data carriers of the proving path: the C++ mirror of the reference's share / point / zkey types.
//   Rep3PrimeFieldShare{a,b}          mpc-core/src/protocols/rep3/fieldshare.rs:14-17
//   Rep3PrimeFieldShareVec{a,b}       mpc-core/src/protocols/rep3/fieldshare.rs:232-236   (SoA; here: two HBM arrays)
//   Rep3PointShare{a,b}               mpc-core/src/protocols/rep3/pointshare.rs:11-14
//   ZKey<P>                           co-circom/circom-types/src/groth16/zkey.rs:47-71
// (paths under /root/reference).  Everything numeric is little-endian u64 limbs in Montgomery form, exactly the bytes
// the C ABI (include/cocg.h) takes, so nothing is converted between this layer and the kernels.
#pragma once
#include <stdint.h>
#include <string.h>

#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/cocg.h"

namespace cohost {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct Fr {
  uint64_t l[4];
  bool operator==(const Fr& o) const { return memcmp(l, o.l, 32) == 0; }
};

// A group element: Jacobian (X, Y, Z), or packed affine (x, y); G1 or G2; coordinates of 4 (BN254) or 6 (BLS12-381)
// limbs (G2: two of those per coordinate).  36 limbs covers Jacobian G2 over BLS12-381.
struct Point {
  uint64_t l[36];
  Point() { memset(l, 0, sizeof(l)); }
};

inline void check(cocg_ctx* ctx, int rc, const char* what) {
  if (rc) throw Error(std::string(what) + ": " + cocg_last_error(ctx));
}

// n Fr elements in HBM, owned by one driver's context; freed back to the driver's pool by the owner.
struct DevVec {
  void* p = nullptr;
  size_t n = 0;
  void* at(size_t off) const { return (char*)p + off * 32; }
};

struct FieldShare {  // b unused by the plain driver
  Fr a, b;
};
struct FieldShareVec {  // b.p == nullptr for the plain driver
  DevVec a, b;
  size_t len() const { return a.n; }
};
struct PointShare {
  Point a, b;
};

// The evaluation domain a caller hands to fft/ifft: size 2^log_n and its generator (`domain.group_gen`, overridden to
// the snarkjs root by root_of_unity_for_groth16, co-circom/co-groth16/src/groth16.rs:57-77).
struct Domain {
  unsigned log_n = 0;
  Fr group_gen;
  size_t size() const { return (size_t)1 << log_n; }
};

// Multi-GPU work distribution of one REP3 Groth16 proof by BLOCKS.  A proof is 15 blocks of nearly equal device time
// (B200, 2^20 constraints: 9.8 / 9.1 / 8.8 ms): per party its witness map together with the two h-query MSMs; per (party, share
// component) the {l, a, b_g1} MSMs, which share one digit sort; per (party, share component) the b_g2 MSM.  Every block runs at FULL
// size on one rank -- the same kernels, window width and occupancy as the single-GPU proof -- instead of every MSM being cut into
// `world` index ranges (smaller windows, more additions per scalar, launch-latency-bound reductions).  The partial results still meet in
// one all-gather per proof.  The plan is a pure function of `world`, so all ranks agree on it without talking.
struct BlockPlan {
  int world = 1;
  int wm[3] = {0, 0, 0};                 // rank of party p's witness map + h MSMs
  int g1[3][2][3] = {};                  // rank of the l / a / b_g1 MSM (index 0 / 1 / 2) of (party, component)
  int g2[3][2] = {{0, 0}, {0, 0}, {0, 0}};  // rank of the b_g2 MSM of (party, component)
  // Costs in ms at 2^20 on a B200 (profiles/r02_launches_*): a witness map with its two h MSMs 9.7, a b_g2 MSM 8.3 with its digit
  // sort, a G1 MSM 2.5 plus 0.4 for the digit sort of its (party, component) scalars when no other MSM of that vector runs on the
  // rank.  Largest first onto the least-loaded rank; the G1 MSMs are placed ONE BY ONE (round 2 placed the {l, a, b_g1} bundles:
  // 15 blocks on 2 / 4 / 8 ranks left the slowest rank 5 - 10 % above the mean).  A pure function of `world`: ranks agree without talking.
  static BlockPlan make(int world) {
    BlockPlan p;
    p.world = world;
    std::vector<double> load(world, 0.0);
    std::vector<char> sorted_here((size_t)world * 6, 0);  // [rank][party * 2 + component]: that scalar vector is sorted on the rank
    auto least = [&](auto cost) {
      int best = 0;
      for (int r = 1; r < world; r++)
        if (load[r] + cost(r) < load[best] + cost(best) - 1e-9) best = r;
      load[best] += cost(best);
      return best;
    };
    for (int q = 0; q < 3; q++) p.wm[q] = least([](int) { return 9.7; });
    for (int q = 0; q < 3; q++)
      for (int c = 0; c < 2; c++) {
        const int r = least([](int) { return 8.3; });
        p.g2[q][c] = r;
        sorted_here[(size_t)r * 6 + q * 2 + c] = 1;
      }
    for (int k = 0; k < 3; k++)
      for (int q = 0; q < 3; q++)
        for (int c = 0; c < 2; c++) {
          const int r = least([&](int rr) { return sorted_here[(size_t)rr * 6 + q * 2 + c] ? 2.5 : 2.9; });
          p.g1[q][c][k] = r;
          sorted_here[(size_t)r * 6 + q * 2 + c] = 1;
        }
    return p;
  }
  bool has_wm(int rank) const { return wm[0] == rank || wm[1] == rank || wm[2] == rank; }
  bool has_g1(int rank, int k) const {  // k: 0 l_query, 1 a_query, 2 b_g1_query
    for (int q = 0; q < 3; q++)
      for (int c = 0; c < 2; c++)
        if (g1[q][c][k] == rank) return true;
    return false;
  }
  bool has_g2(int rank) const {
    for (int q = 0; q < 3; q++)
      for (int c = 0; c < 2; c++)
        if (g2[q][c] == rank) return true;
    return false;
  }
  // does `rank` read share component c of party q's witness?
  bool reads_witness(int rank, int q, int c) const {
    return wm[q] == rank || g2[q][c] == rank || g1[q][c][0] == rank || g1[q][c][1] == rank || g1[q][c][2] == rank;
  }
};

// Device-resident proving key.  Query arrays live in HBM behind bases handles of `owner`; drivers alias them
// (cocg_bases_share), as the reference's three in-process drivers borrow one &ZKey.
struct ZKey {
  int curve = 0;
  cocg_ctx* owner = nullptr;
  size_t n_public = 0;         // l  (public inputs without the leading 1)
  size_t n_vars = 0;           // m
  size_t pow = 0;              // log2(domain size)
  size_t num_constraints = 0;
  size_t domain_size() const { return (size_t)1 << pow; }
  size_t num_inputs() const { return n_public + 1; }  // matrices.num_instance_variables
  size_t n_aux() const { return n_vars - n_public - 1; }
  uint64_t a_query = 0, b_g1_query = 0, b_g2_query = 0, h_query = 0, l_query = 0;  // handles in `owner`
  // Multi-GPU: a rank may hold only the slice of each query it accumulates (index-range sharding); `*_first` is
  // the index, in the full query, of the resident table's entry 0 (0 when the whole query is resident).
  int rank = 0, world = 1;
  size_t a_first = 0, b_g1_first = 0, b_g2_first = 0, h_first = 0, l_first = 0;
  // block mode (BlockPlan): whole queries are resident, but only those of the blocks this rank runs (handle 0 = not resident)
  bool blocks = false;
  BlockPlan plan;
  uint64_t csr_a = 0, csr_b = 0;
  // the first 1 + l points of each coefficient query stay on the host as well (calculate_coeff groth16.rs:219-231)
  std::vector<Point> a_head, b_g1_head, b_g2_head;  // packed affine
  Point alpha_g1, beta_g1, delta_g1, beta_g2, delta_g2;  // packed affine
};

struct Groth16Proof {  // packed affine, Montgomery coordinates; groth16/proof.rs:7-29
  Point pi_a, pi_b, pi_c;
};

}  // namespace cohost
