// Counter-addressed ChaCha12 PRF -> uniform field elements, shared by the device kernels and the host driver.
//
// The reference's REP3 correlated randomness is two ChaCha12 streams per party (RngType = ChaCha12Rng,
// /root/reference/mpc-core/src/lib.rs:10, protocols/rep3/rngs.rs:25-46): rng1 seeded with the party's own seed,
// rng2 with the previous party's (rep3.rs:343-349); a zero-mask is F::rand(rng1) - F::rand(rng2) and F::rand is
// rejection sampling on bit-masked 256-bit draws (ark-ff 0.4.2).  A sequential stream with a variable number of
// draws per element cannot be evaluated in parallel, so element i of mask-vector `ctr` is addressed directly:
//   block(j) = ChaCha12(key = seed, counter = (ctr, i), nonce = j)      j = 0, 1, ...
//   candidates = the two 256-bit halves of block(0), block(1), ... with the top (256 - BITS) bits cleared;
//   the first candidate < modulus is the canonical value, taken as the Montgomery representative directly
//   (a uniform residue is uniform in either form).
// Masks cancel on opening (sum over the three parties is zero), so proofs do not depend on this choice; what
// matters is that party i's rng2 and party i-1's rng1 agree, which they do because both run this function.
#pragma once
#include "fp.cuh"

namespace cocg {

struct PrfKey {
  uint32_t k[8];
};

COCG_HD uint32_t prf_rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

#define COCG_QR(a, b, c, d)  \
  a += b; d ^= a; d = prf_rotl(d, 16); \
  c += d; b ^= c; b = prf_rotl(b, 12); \
  a += b; d ^= a; d = prf_rotl(d, 8);  \
  c += d; b ^= c; b = prf_rotl(b, 7);

// One 64-byte ChaCha12 block: words 12,13 = 64-bit element index, 14 = vector counter, 15 = retry index.
COCG_HD void chacha12_block(const PrfKey& key, uint64_t idx, uint32_t ctr, uint32_t retry, uint32_t out[16]) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3],
                    key.k[4],    key.k[5],    key.k[6],    key.k[7],    (uint32_t)idx, (uint32_t)(idx >> 32), ctr, retry};
  uint32_t x0 = s[0], x1 = s[1], x2 = s[2], x3 = s[3], x4 = s[4], x5 = s[5], x6 = s[6], x7 = s[7];
  uint32_t x8 = s[8], x9 = s[9], x10 = s[10], x11 = s[11], x12 = s[12], x13 = s[13], x14 = s[14], x15 = s[15];
#pragma unroll
  for (int r = 0; r < 6; r++) {
    COCG_QR(x0, x4, x8, x12) COCG_QR(x1, x5, x9, x13) COCG_QR(x2, x6, x10, x14) COCG_QR(x3, x7, x11, x15)
    COCG_QR(x0, x5, x10, x15) COCG_QR(x1, x6, x11, x12) COCG_QR(x2, x7, x8, x13) COCG_QR(x3, x4, x9, x14)
  }
  out[0] = x0 + s[0]; out[1] = x1 + s[1]; out[2] = x2 + s[2]; out[3] = x3 + s[3];
  out[4] = x4 + s[4]; out[5] = x5 + s[5]; out[6] = x6 + s[6]; out[7] = x7 + s[7];
  out[8] = x8 + s[8]; out[9] = x9 + s[9]; out[10] = x10 + s[10]; out[11] = x11 + s[11];
  out[12] = x12 + s[12]; out[13] = x13 + s[13]; out[14] = x14 + s[14]; out[15] = x15 + s[15];
}
#undef COCG_QR

template <class P>
COCG_HD bool prf_below_modulus(const uint32_t* v) {
  for (int i = P::N - 1; i >= 0; i--) {
    if (v[i] < P::mod(i)) return true;
    if (v[i] > P::mod(i)) return false;
  }
  return false;
}

// Uniform element of the 8-limb field P for (key, vector counter, element index).
template <class P>
COCG_HD Fp<P> prf_field(const PrfKey& key, uint32_t ctr, uint64_t idx) {
  static_assert(P::N == 8, "scalar fields are 8 x 32 bit");
  constexpr uint32_t top_mask = (P::BITS % 32) ? ((1u << (P::BITS % 32)) - 1u) : 0xffffffffu;
  Fp<P> r;
  for (uint32_t retry = 0;; retry++) {
    uint32_t blk[16];
    chacha12_block(key, idx, ctr, retry, blk);
    for (int half = 0; half < 2; half++) {
      uint32_t v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = blk[8 * half + i];
      v[7] &= top_mask;
      if (prf_below_modulus<P>(v)) {
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = v[i];
        return r;
      }
    }
  }
}

}  // namespace cocg
