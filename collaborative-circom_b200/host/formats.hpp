// snarkjs on-disk formats feeding the path, read straight into the device layout:
//   binfile container     /root/reference/co-circom/circom-types/src/binfile.rs:52-105
//   Groth16 .zkey         /root/reference/co-circom/circom-types/src/groth16/zkey.rs:139-316 (sections 2-9; public-input rows dropped :196-204)
//   field / point readers /root/reference/co-circom/circom-types/src/traits.rs:47-69 (Fr), :107-155 (G1/G2: Montgomery x|y, (0,0) = infinity)
//   .wtns                 /root/reference/co-circom/circom-types/src/witness.rs:51-92 (canonical little-endian values)
// Query points are stored by snarkjs exactly in the layout the kernels use (packed Montgomery affine), so sections 5-9 are
// uploaded as they are; matrix coefficients are stored times R^2 and leave one Montgomery factor on the way in
// (from_reader_for_groth16_zkey, traits.rs:65-67), witness values enter Montgomery form -- both conversions run on the GPU.
// Not done here: the per-point on-curve / subgroup checks of the reference's readers (traits.rs:121-123).
#pragma once
#include <cstdio>
#include <map>

#include "types.hpp"

namespace cohost {

struct BinFile {
  uint32_t version = 0;
  std::map<uint32_t, std::pair<const uint8_t*, size_t>> sections;  // first occurrence wins
  static uint32_t u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
  static uint64_t u64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
  BinFile(const uint8_t* data, size_t len, const char magic[4]) {
    if (len < 12 || memcmp(data, magic, 4) != 0) throw Error(std::string("invalid file header, expected ") + std::string(magic, 4));
    version = u32(data + 4);
    const uint32_t nsec = u32(data + 8);
    size_t off = 12;
    for (uint32_t i = 0; i < nsec; i++) {
      if (off + 12 > len) throw Error("binfile: truncated section header");
      const uint32_t id = u32(data + off);
      const uint64_t sl = u64(data + off + 4);
      off += 12;
      if (sl > len - off) throw Error("binfile: truncated section");
      sections.emplace(id, std::make_pair(data + off, (size_t)sl));
      off += sl;
    }
  }
  std::pair<const uint8_t*, size_t> take(uint32_t id) const {
    auto it = sections.find(id);
    if (it == sections.end()) throw Error("binfile: missing section " + std::to_string(id));
    return it->second;
  }
};

inline std::vector<uint8_t> read_file(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) throw Error(std::string("cannot open ") + path);
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)(n > 0 ? n : 0));
  if (n > 0 && fread(buf.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); throw Error(std::string("short read on ") + path); }
  fclose(f);
  return buf;
}

static const uint8_t kBn254Q[32] = {0x47, 0xfd, 0x7c, 0xd8, 0x16, 0x8c, 0x20, 0x3c, 0x8d, 0xca, 0x71, 0x68, 0x91, 0x6a, 0x81, 0x97,
                                    0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};
static const uint8_t kBls381Q[48] = {0xab, 0xaa, 0xff, 0xff, 0xff, 0xff, 0xfe, 0xb9, 0xff, 0xff, 0x53, 0xb1, 0xfe, 0xff, 0xab, 0x1e,
                                     0x24, 0xf6, 0xb0, 0xf6, 0xa0, 0xd2, 0x30, 0x67, 0xbf, 0x12, 0x85, 0xf3, 0x84, 0x4b, 0x77, 0x64,
                                     0xd7, 0xac, 0x4b, 0x43, 0xb6, 0xa7, 0x1b, 0x4b, 0x9a, 0xe6, 0x7f, 0x39, 0xea, 0x11, 0x01, 0x1a};
static const uint8_t kBn254R[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                                    0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};
static const uint8_t kBls381R[32] = {0x01, 0x00, 0x00, 0x00, 0xff, 0xff, 0xff, 0xff, 0xfe, 0x5b, 0xfe, 0xff, 0x02, 0xa4, 0xbd, 0x53,
                                     0x05, 0xd8, 0xa1, 0x09, 0x08, 0xd8, 0x39, 0x33, 0x48, 0x7d, 0x9d, 0x29, 0x53, 0xa7, 0xed, 0x73};

// Parsed view of a Groth16 zkey: pointers into the file image plus the CSR form of A and B.
struct Groth16ZKeyFile {
  int curve = 0;
  size_t n8q = 0, n_vars = 0, n_public = 0, domain_size = 0, pow = 0, num_constraints = 0;
  const uint8_t *alpha_g1 = nullptr, *beta_g1 = nullptr, *beta_g2 = nullptr, *gamma_g2 = nullptr, *delta_g1 = nullptr, *delta_g2 = nullptr;
  const uint8_t *ic = nullptr, *a_query = nullptr, *b_g1_query = nullptr, *b_g2_query = nullptr, *l_query = nullptr, *h_query = nullptr;
  std::vector<uint32_t> rowptr[2], col[2];
  std::vector<uint8_t> coeff_raw[2];  // 32 bytes per entry as stored (value * R^2): one Montgomery reduction away from the device form

  Groth16ZKeyFile(const uint8_t* data, size_t len) {
    BinFile bf(data, len, "zkey");
    auto s1 = bf.take(1);
    if (s1.second < 4 || BinFile::u32(s1.first) != 1) throw Error("zkey: not a groth16 key (protocol id != 1)");
    auto h = bf.take(2);
    const uint8_t* p = h.first;
    const uint8_t* end = h.first + h.second;
    auto need = [&](size_t k) { if ((size_t)(end - p) < k) throw Error("zkey: truncated header"); };
    need(4);
    n8q = BinFile::u32(p); p += 4;
    if (n8q != 32 && n8q != 48) throw Error("zkey: unexpected base field byte size " + std::to_string(n8q));
    need(n8q);
    curve = n8q == 32 ? COCG_BN254 : COCG_BLS12_381;
    if (memcmp(p, curve == COCG_BN254 ? kBn254Q : kBls381Q, n8q) != 0) throw Error("zkey: invalid prime in header");
    p += n8q;
    need(4);
    const uint32_t n8r = BinFile::u32(p); p += 4;
    if (n8r != 32) throw Error("zkey: unexpected scalar field byte size " + std::to_string(n8r));
    need(32);
    if (memcmp(p, curve == COCG_BN254 ? kBn254R : kBls381R, 32) != 0) throw Error("zkey: invalid prime in header");
    p += 32;
    need(12);
    n_vars = BinFile::u32(p); n_public = BinFile::u32(p + 4); domain_size = BinFile::u32(p + 8); p += 12;
    if (domain_size == 0 || (domain_size & (domain_size - 1))) throw Error("zkey: domain size must be a power of two");
    while (((size_t)1 << pow) < domain_size) pow++;
    if (n_vars < n_public + 1) throw Error("zkey: n_vars < n_public + 1");
    const size_t g1 = 2 * n8q, g2 = 4 * n8q;
    need(3 * g1 + 3 * g2);
    alpha_g1 = p; p += g1;
    beta_g1 = p; p += g1;
    beta_g2 = p; p += g2;
    gamma_g2 = p; p += g2;
    delta_g1 = p; p += g1;
    delta_g2 = p; p += g2;
    auto sec = [&](uint32_t id, size_t bytes, const char* what) {
      auto s = bf.take(id);
      if (s.second < bytes) throw Error(std::string("zkey: section too short: ") + what);
      return s.first;
    };
    ic = sec(3, (n_public + 1) * g1, "IC");
    a_query = sec(5, n_vars * g1, "A");
    b_g1_query = sec(6, n_vars * g1, "B1");
    b_g2_query = sec(7, n_vars * g2, "B2");
    l_query = sec(8, (n_vars - n_public - 1) * g1, "L");
    h_query = sec(9, domain_size * g1, "H");
    // section 4: u32 n; n x (u32 matrix, u32 constraint, u32 signal, Fr * R^2)
    auto m = bf.take(4);
    if (m.second < 4) throw Error("zkey: truncated coefficient section");
    const uint32_t ncoef = BinFile::u32(m.first);
    if (m.second < 4 + (size_t)ncoef * 44) throw Error("zkey: truncated coefficient section");
    uint32_t max_row = 0;
    for (uint32_t i = 0; i < ncoef; i++) {
      const uint8_t* e = m.first + 4 + (size_t)i * 44;
      if (BinFile::u32(e) > 1) throw Error("zkey: matrix index out of range");
      if (BinFile::u32(e + 4) >= domain_size) throw Error("zkey: constraint index out of range");
      max_row = std::max(max_row, BinFile::u32(e + 4));
    }
    if (max_row < n_public) throw Error("zkey: fewer constraints than public inputs");
    num_constraints = max_row - n_public;  // the public-input rows at the tail are dropped (zkey.rs:196-204)
    for (int k = 0; k < 2; k++) rowptr[k].assign(num_constraints + 1, 0);
    for (uint32_t i = 0; i < ncoef; i++) {
      const uint8_t* e = m.first + 4 + (size_t)i * 44;
      const uint32_t k = BinFile::u32(e), row = BinFile::u32(e + 4);
      if (row < num_constraints) rowptr[k][row + 1]++;
    }
    for (int k = 0; k < 2; k++) {
      for (size_t r = 0; r < num_constraints; r++) rowptr[k][r + 1] += rowptr[k][r];
      col[k].resize(rowptr[k][num_constraints]);
      coeff_raw[k].resize((size_t)rowptr[k][num_constraints] * 32);
    }
    std::vector<uint32_t> fill[2] = {std::vector<uint32_t>(rowptr[0].begin(), rowptr[0].end() - 1),
                                     std::vector<uint32_t>(rowptr[1].begin(), rowptr[1].end() - 1)};
    for (uint32_t i = 0; i < ncoef; i++) {  // entries keep their file order within a row, as the reference's push does
      const uint8_t* e = m.first + 4 + (size_t)i * 44;
      const uint32_t k = BinFile::u32(e), row = BinFile::u32(e + 4), sig = BinFile::u32(e + 8);
      if (row >= num_constraints) continue;
      // the reference indexes the witness with this value and panics on the Rust bounds check; here it would become an
      // out-of-bounds gather on the device (spmv.cu)
      if (sig >= n_vars) throw Error("zkey: signal index out of range in the coefficient section");
      const uint32_t at = fill[k][row]++;
      col[k][at] = sig;
      memcpy(coeff_raw[k].data() + (size_t)at * 32, e + 12, 32);
    }
  }
};

// .r1cs header (circom-types/src/r1cs.rs:100-215): only what split-witness needs -- num_inputs = 1 + n_pub_out + n_pub_in (r1cs.rs:201).
struct R1csHeader {
  int curve = 0;
  size_t n_wires = 0, n_pub_out = 0, n_pub_in = 0, n_prv_in = 0, n_constraints = 0;
  size_t num_inputs() const { return 1 + n_pub_out + n_pub_in; }
  R1csHeader(const uint8_t* data, size_t len) {
    BinFile bf(data, len, "r1cs");
    auto h = bf.take(1);
    const uint8_t* p = h.first;
    if (h.second < 4) throw Error("r1cs: truncated header");
    const uint32_t n8 = BinFile::u32(p);
    if (n8 != 32 || h.second < 4 + 32 + 4 * 4 + 8 + 4) throw Error("r1cs: wrong scalar field");
    if (memcmp(p + 4, kBn254R, 32) == 0) curve = COCG_BN254;
    else if (memcmp(p + 4, kBls381R, 32) == 0) curve = COCG_BLS12_381;
    else throw Error("r1cs: wrong scalar field");
    p += 36;
    n_wires = BinFile::u32(p); n_pub_out = BinFile::u32(p + 4); n_pub_in = BinFile::u32(p + 8); n_prv_in = BinFile::u32(p + 12);
    n_constraints = BinFile::u32(p + 24);
    if (num_inputs() > n_wires) throw Error("r1cs: more inputs than wires");
  }
};

// .wtns: returns the canonical little-endian values (32 bytes each); the caller moves them to Montgomery form on the GPU
struct WitnessFile {
  int curve = 0;
  size_t n = 0;
  const uint8_t* values = nullptr;
  WitnessFile(const uint8_t* data, size_t len) {
    if (len < 12 || memcmp(data, "wtns", 4) != 0) throw Error("invalid file header, expected wtns");
    if (BinFile::u32(data + 4) > 2) throw Error("wtns: version not supported");
    if (BinFile::u32(data + 8) > 2) throw Error("wtns: invalid number of sections");
    size_t off = 12 + 12;  // first section header (id, length) is skipped like the reference does
    if (len < off + 4) throw Error("wtns: truncated");
    const uint32_t n8 = BinFile::u32(data + off);
    off += 4;
    if (n8 != 32 || len < off + 32 + 4 + 12) throw Error("wtns: wrong scalar field");
    if (memcmp(data + off, kBn254R, 32) == 0) curve = COCG_BN254;
    else if (memcmp(data + off, kBls381R, 32) == 0) curve = COCG_BLS12_381;
    else throw Error("wtns: wrong scalar field");
    off += 32;
    n = BinFile::u32(data + off);
    off += 4 + 12;
    if (len < off + n * 32) throw Error("wtns: truncated values");
    values = data + off;
  }
};

}  // namespace cohost
