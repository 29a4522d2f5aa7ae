/* cohost -- C entry points of the host layer above the cocg kernels (libcohost.so).
 *
 * libcocg.so (include/cocg.h) is the drop-in boundary: it replaces the arithmetic behind the reference's driver traits.
 * This header exposes the layer ABOVE it, written in C++ because the reference's own host code is compiled Rust and no
 * Rust toolchain exists in this image: the reference's prover structure -- CoGroth16<T>::prove
 * (/root/reference/co-circom/co-groth16/src/groth16.rs:113-326) over PlainDriver (mpc-core/src/protocols/plain.rs) or
 * three Rep3Protocol drivers (mpc-core/src/protocols/rep3.rs) on three threads joined by an in-process network
 * (tests/src/rep3_network.rs) -- so tests and bench.py can run whole proofs.  All numbers are little-endian u64 limbs in
 * Montgomery form; affine points are packed x|y (G2: x.c0|x.c1|y.c0|y.c1), all-zero = infinity.
 * Every function returns 0 on success; cohost_last_error() describes the last failure on the calling thread. */
#ifndef COHOST_H
#define COHOST_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define COHOST_API __attribute__((visibility("default")))
#else
#define COHOST_API
#endif
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cohost_zkey cohost_zkey;
typedef struct cohost_plain_session cohost_plain_session;
typedef struct cohost_rep3_session cohost_rep3_session;
typedef struct cohost_shamir_session cohost_shamir_session;
typedef struct cohost_plonk_zkey cohost_plonk_zkey;

/* A parsed Groth16 proving key (circom-types/src/groth16/zkey.rs:47-71), HOST pointers; copied to HBM once. */
typedef struct cohost_zkey_desc {
  int curve;                 /* COCG_BN254 | COCG_BLS12_381 */
  int device;
  size_t n_public;           /* l: public inputs without the constant 1 */
  size_t n_vars;             /* m */
  size_t pow;                /* log2(domain size) */
  size_t num_constraints;    /* rows of A and B (public-input identity rows already dropped, zkey.rs:196-204) */
  const uint32_t *a_rowptr, *a_col; const void* a_coeff; size_t a_nnz;   /* CSR, coefficients Montgomery Fr */
  const uint32_t *b_rowptr, *b_col; const void* b_coeff; size_t b_nnz;
  const void *a_query, *b_g1_query, *b_g2_query;   /* m points each (G1, G1, G2) */
  const void *h_query;                             /* 2^pow G1 points */
  const void *l_query;                             /* m - l - 1 G1 points */
  const void *alpha_g1, *beta_g1, *delta_g1, *beta_g2, *delta_g2;
  /* NULL, or 32 bytes: every query / vk pointer above that is NULL is filled with synthetic curve points generated in HBM
   * (cocg_bases_generate) -- the shape-faithful 2^20-constraint benchmark key, for which no zkey ships. */
  const void* synthetic_seed;
  int rank, world;           /* world > 1: only this rank's part of the queries becomes resident */
  int coeff_form;            /* COCG_FORM_MONT (0, default) | COCG_FORM_R2 (as stored in a zkey) | COCG_FORM_CANONICAL */
  int shard_mode;            /* world > 1: 0 = every MSM cut into `world` index ranges; 1 = whole blocks of the proof per rank (REP3):
                                the witness map + h MSMs of a party, the {l, a, b_g1} MSMs or the b_g2 MSM of a (party, component) */
} cohost_zkey_desc;

typedef struct cohost_zkey_info {
  int curve;
  size_t n_public, n_vars, pow, num_constraints;
} cohost_zkey_info;

/* Injected randomness for parity tests (the reference draws all of it from entropy). */
typedef struct cohost_rep3_randomness {
  const void* r;         /* 3 x (a | b) Fr: party i's replicated share of r */
  const void* s;         /* 3 x (a | b) Fr */
  const void* mask_rs;   /* 3 Fr summing to zero: masks of mul(r, s) */
  const void* mask_pt;   /* 3 G1 Jacobian points summing to infinity: masks of scalar_mul(g1_b, r) */
  const void* masks1[3]; /* HOST vectors (domain size) summing to zero: first mul_vec */
  const void* masks2[3]; /* second mul_vec */
} cohost_rep3_randomness;

COHOST_API const char* cohost_last_error(void);
COHOST_API int cohost_zkey_create(const cohost_zkey_desc* desc, cohost_zkey** out);
COHOST_API void cohost_zkey_destroy(cohost_zkey* z);
/* snarkjs files straight into the device layout (circom-types/src/groth16/zkey.rs:109-251, binfile.rs:52-105, witness.rs:51-92). */
COHOST_API int cohost_zkey_load(const void* data, size_t len, int device, cohost_zkey** out);
COHOST_API int cohost_zkey_load_file(const char* path, int device, cohost_zkey** out);
COHOST_API int cohost_zkey_get_info(cohost_zkey* z, cohost_zkey_info* info);
/* which: 0 a_query, 1 b_g1_query, 2 b_g2_query, 3 h_query, 4 l_query -> n packed affine Montgomery points from `off` */
COHOST_API int cohost_zkey_query_download(cohost_zkey* z, int which, size_t off, size_t n, void* out);
/* which: 0 A, 1 B; rowptr / col / coeff may be NULL; *nnz is always set */
COHOST_API int cohost_zkey_matrix_download(cohost_zkey* z, int which, uint32_t* rowptr, uint32_t* col, void* coeff, size_t* nnz);
/* alpha_g1 | beta_g1 | delta_g1 | beta_g2 | delta_g2, packed affine */
COHOST_API int cohost_zkey_vk_download(cohost_zkey* z, void* out);
/* .wtns values as Montgomery limbs; out may be NULL to query the count */
COHOST_API int cohost_wtns_load_file(cohost_zkey* z, const char* path, void* out, size_t* n_values);

/* CoGroth16<PlainDriver>::prove.  public_inputs: l + 1 Fr (leading 1); witness: m - l - 1 Fr; r, s: one Fr each or NULL
 * (PRF); proof_out: A | B | C packed affine; h_out: NULL or 2^pow Fr. */
COHOST_API int cohost_plain_session_create(cohost_zkey* z, cohost_plain_session** out);
COHOST_API void cohost_plain_session_destroy(cohost_plain_session* s);
COHOST_API int cohost_plain_prove(cohost_plain_session* s, const void* public_inputs, const void* witness, const void* r, const void* s_rand,
                                  void* proof_out, void* h_out);

/* Three CoGroth16<Rep3Protocol> provers.  seeds: 3 x 32 bytes (each party's PRF seed, rep3.rs:343-349).
 * rank/world: index-range sharding of every MSM over `world` GPUs (one process per GPU); world = 1 for a single GPU. */
COHOST_API int cohost_rep3_session_create(cohost_zkey* z, const uint8_t* seeds, int rank, int world, cohost_rep3_session** out);
/* Block mode (zkey made with shard_mode = 1): a proof is 15 blocks of equal device time spread over the ranks; each block runs at full
 * size with the single-GPU kernels.  The witness maps of the three parties then live on different GPUs and their two mul_vec payloads
 * per proof travel GPU to GPU: `comm` receives the transfers of one round of this rank and must issue them as ONE grouped NCCL call
 * (dir 0 = send, 1 = receive; dptr = device buffer of this rank) and return 0 once they have completed.  It is called from a party
 * thread while the caller waits in cohost_rep3_prove_partials.  Partial sums still meet in one all-gather per proof (partials / combine). */
typedef struct cohost_comm_op { int dir; int peer; void* dptr; size_t bytes; } cohost_comm_op;
typedef int (*cohost_comm_cb)(void* user, const cohost_comm_op* ops, int nops);
COHOST_API int cohost_rep3_session_create_blocks(cohost_zkey* z, const uint8_t* seeds, int rank, int world, cohost_comm_cb comm, void* user,
                                                 cohost_rep3_session** out);
/* The block plan for `world` ranks (no GPU needed): out[27] = wm[3] | g2[3][2] | g1[3][2][3] ranks (party, share component, query l / a / b_g1). */
COHOST_API int cohost_block_plan(int world, int* out);
COHOST_API void cohost_rep3_session_destroy(cohost_rep3_session* s);
/* One proof = begin [-> partials -> (caller all-gathers) -> combine] -> end.  wit_a[i] / wit_b[i]: party i's HOST share
 * components (m - l - 1 Fr each).  rnd: NULL for PRF-derived randomness. */
COHOST_API int cohost_rep3_prove_begin(cohost_rep3_session* s, const void* public_inputs, const void* const* wit_a, const void* const* wit_b,
                                       const cohost_rep3_randomness* rnd);
/* Same with the witness share components already resident in HBM (DEVICE pointers on the session's device). */
COHOST_API int cohost_rep3_prove_begin_device(cohost_rep3_session* s, const void* public_inputs, const void* const* wit_a,
                                              const void* const* wit_b, const cohost_rep3_randomness* rnd);
/* Per-kernel-class device time summed over the three drivers (cocg_profile_*; classes COCG_PROF_*). */
COHOST_API int cohost_rep3_profile_enable(cohost_rep3_session* s, int on);
COHOST_API int cohost_rep3_profile_read(cohost_rep3_session* s, int cls, double* total_ms, uint64_t* scopes);
COHOST_API int cohost_rep3_profile_reset(cohost_rep3_session* s);
COHOST_API size_t cohost_rep3_partial_bytes(cohost_rep3_session* s);
COHOST_API int cohost_rep3_prove_partials(cohost_rep3_session* s, void* out);
COHOST_API int cohost_rep3_prove_combine(cohost_rep3_session* s, const void* gathered);
/* proofs_out: 3 x (A | B | C); h_a / h_b: NULL or 3 HOST buffers of 2^pow Fr receiving each party's share of h. */
COHOST_API int cohost_rep3_prove_end(cohost_rep3_session* s, void* proofs_out, void* const* h_a, void* const* h_b);
COHOST_API uint64_t cohost_rep3_launch_count(cohost_rep3_session* s);
/* mul_vec payloads between the three co-located parties: device = 0 staged through pinned host memory (default: what a party that
 * must reach a NIC does; also selected by COHOST_MPC_EXCHANGE=host), device = 1 handed over in HBM (the parties share one GPU). */
COHOST_API int cohost_rep3_set_mpc_exchange(cohost_rep3_session* s, int device);
/* Host wall-clock per phase of the last proof, seconds: out[party * 4 + k], k = witness map | MSMs | all-gather wait | assembly. */
COHOST_API int cohost_rep3_phase_times(cohost_rep3_session* s, double* out);
/* num_parties CoGroth16<ShamirProtocol> provers (mpc-core/src/protocols/shamir.rs), threshold t with 2t + 1 <= num_parties, on
 * one thread each over an in-process network; seeds: num_parties x 32 bytes.  wit[i]: party i's HOST share vector; proofs_out:
 * num_parties x (A | B | C); rs_out: NULL or num_parties x (share of r | share of s) for tests. */
/* Multi-GPU Shamir (BASELINE configs[4]): every MSM is sharded by index range over `world` ranks (the zkey created with the same rank /
 * world keeps only this rank's slice of each query resident); the partial sums of all parties travel in ONE all-gather per proof, which
 * the caller supplies: gather(user, local, bytes, gathered) fills gathered[world x bytes] in rank order and returns 0. */
typedef int (*cohost_gather_cb)(void* user, const void* local, size_t bytes, void* gathered);
COHOST_API int cohost_shamir_session_set_shard(cohost_shamir_session* s, int rank, int world, cohost_gather_cb gather, void* user);
COHOST_API int cohost_shamir_set_mpc_exchange(cohost_shamir_session* s, int device);  /* as cohost_rep3_set_mpc_exchange */
COHOST_API int cohost_shamir_session_create(cohost_zkey* z, int num_parties, int threshold, const uint8_t* seeds, cohost_shamir_session** out);
COHOST_API void cohost_shamir_session_destroy(cohost_shamir_session* s);
COHOST_API int cohost_shamir_prove(cohost_shamir_session* s, const void* public_inputs, const void* const* wit, void* proofs_out, void* rs_out);
/* CoPlonk::prove (co-plonk/src/lib.rs:80-99, round1.rs .. round5.rs) on the GPU.
 *  - cohost_plonk_zkey_load_file: the Plonk zkey reader (circom-types/src/plonk/zkey.rs:160-420): header incl. the verifying-key tail,
 *    additions, wire maps, selector / sigma / Lagrange polynomials and p_tau, all resident in HBM afterwards.
 *    info[6] = curve, n_vars, n_public, domain_size, n_additions, n_constraints.
 *  - cohost_plonk_zkey_create_synthetic: a shape-faithful key for the 2^18-gate benchmark (BASELINE configs[3]): caller's wire maps,
 *    PRF-filled polynomials, generated curve points; seed = 32 bytes.  Its proofs exercise every kernel but are not expected to verify.
 *  - sessions: protocol 0 = CoPlonk<PlainDriver> (seeds: 32 bytes), 1 = three CoPlonk<Rep3Protocol> provers on three threads over the
 *    in-process network (seeds: 3 x 32 bytes).
 *  - cohost_plonk_prove: public_inputs = n_public + 1 Fr (the leading entry is ignored, PlonkWitness::new types.rs:105-108);
 *    wit_a / wit_b: one pointer per party to its share components of the private witness (n_vars - n_additions - n_public - 1 Fr; wit_b
 *    NULL for the plain driver), HOST memory or, with wit_on_device, HBM; deterministic != 0: the reference's KAT blinders b_i = i
 *    (round1.rs:101-108).  proofs_out: per party cohost_plonk_proof_limbs() u64 = A B C Z T1 T2 T3 Wxi Wxiw (packed affine
 *    Montgomery) | eval_a eval_b eval_c eval_s1 eval_s2 eval_zw (Montgomery Fr) -- the field order of PlonkProof (plonk/proof.rs).
 *  - cohost_plonk_round1_*: round 1 alone, commitments [a]_1 | [b]_1 | [c]_1 (per party for REP3). */
typedef struct cohost_plonk_session cohost_plonk_session;
COHOST_API int cohost_plonk_zkey_load_file(const char* path, int device, cohost_plonk_zkey** out);
COHOST_API int cohost_plonk_zkey_create_synthetic(int curve, int device, size_t log_n, size_t n_public, size_t n_vars, size_t n_constraints,
                                                  const uint32_t* map_a, const uint32_t* map_b, const uint32_t* map_c, const uint8_t* seed,
                                                  cohost_plonk_zkey** out);
COHOST_API void cohost_plonk_zkey_destroy(cohost_plonk_zkey* z);
COHOST_API int cohost_plonk_zkey_get_info(cohost_plonk_zkey* z, size_t* info);
COHOST_API size_t cohost_plonk_proof_limbs(cohost_plonk_zkey* z);
COHOST_API int cohost_plonk_session_create(cohost_plonk_zkey* z, int protocol, const uint8_t* seeds, cohost_plonk_session** out);
/* CoPlonk over Shamir (num_parties, threshold) shares -- co-plonk is generic over the MPC protocol (co-plonk/src/plonk.rs:50-77); the
 * driver surface is mpc-core/src/protocols/shamir.rs:459-712.  seeds: num_parties x 32 bytes; prove() takes one share per party in wit_a. */
COHOST_API int cohost_plonk_session_create_shamir(cohost_plonk_zkey* z, int num_parties, int threshold, const uint8_t* seeds, cohost_plonk_session** out);
COHOST_API int cohost_plonk_session_parties(cohost_plonk_session* s);
COHOST_API void cohost_plonk_session_destroy(cohost_plonk_session* s);
COHOST_API int cohost_plonk_prove(cohost_plonk_session* s, const void* public_inputs, const void* const* wit_a, const void* const* wit_b,
                                  int deterministic, int wit_on_device, void* proofs_out);
COHOST_API int cohost_plonk_set_mpc_exchange(cohost_plonk_session* s, int device);
COHOST_API uint64_t cohost_plonk_launch_count(cohost_plonk_session* s);
COHOST_API int cohost_plonk_profile_enable(cohost_plonk_session* s, int on);
COHOST_API int cohost_plonk_profile_reset(cohost_plonk_session* s);
COHOST_API int cohost_plonk_profile_read(cohost_plonk_session* s, int cls, double* total_ms, uint64_t* scopes);
/* Host wall-clock per round of the last proof, seconds: out[party * 5 + round]. */
COHOST_API int cohost_plonk_round_times(cohost_plonk_session* s, double* out);
/* Test hook: record component a of named intermediate vectors of the following proofs (buffer_a, poly_a, eval_a, buffer_z, poly_z,
 * t_evals, tz_evals, t1, t2, t3, poly_r, wxi); summed over the three parties they are the plain values.  get with out == NULL
 * returns the element count in *n. */
COHOST_API int cohost_plonk_trace_enable(cohost_plonk_session* s, int on);
COHOST_API int cohost_plonk_trace_get(cohost_plonk_session* s, int party, const char* name, void* out, size_t cap, size_t* n);
COHOST_API int cohost_plonk_round1_plain(cohost_plonk_zkey* z, const void* public_inputs, const void* witness, int deterministic, void* commits_out);
COHOST_API int cohost_plonk_round1_rep3(cohost_plonk_zkey* z, const void* public_inputs, const void* const* wit_a, const void* const* wit_b,
                                        const uint8_t* seeds, int deterministic, void* commits_out);
/* serde_json of PlonkProof (circom-types/src/plonk/proof.rs:7-87) from the block cohost_plonk_prove writes; same writer protocol as below. */
COHOST_API int cohost_plonk_proof_to_json(int curve, const void* proof, char* out, size_t cap, size_t* len);
/* Output-side formats of the path, byte-compatible with what `co-circom generate-proof` / `split-witness` write.
 * Writers fill `out` (capacity `cap`) and set *len; out == NULL only queries the length.  Strings carry no terminating NUL.
 *  - cohost_proof_to_json: serde_json of Groth16Proof (circom-types/src/groth16/proof.rs:7-29, traits.rs:186-233); `proof` is the
 *    A | B | C packed affine Montgomery block the prove calls return.
 *  - cohost_public_inputs_to_json: decimal strings of pub[1..count) (co-circom/src/bin/co-circom.rs:611-629).
 *  - cohost_shared_witness_{encode,decode}: the bincode + ark-serialize image of SharedWitness (co-circom-snarks/src/lib.rs:24-41,
 *    serde_compat.rs:5-24); k = 2 components for REP3 (a, b), 1 for Shamir.  decode with pub == NULL and comps == NULL returns the counts.
 *  - cohost_split_witness_rep3: SharedWitness::share_rep3 (co-circom-snarks/src/lib.rs:149-173) with the random shares drawn by the
 *    GPU PRF from seed[64]; witness / outputs are HOST vectors of n Montgomery Fr (needs a GPU). */
COHOST_API int cohost_proof_to_json(int curve, const void* proof, char* out, size_t cap, size_t* len);
COHOST_API int cohost_public_inputs_to_json(int curve, const void* pub, size_t count, char* out, size_t cap, size_t* len);
COHOST_API int cohost_shared_witness_encode(int curve, const void* pub, size_t n_pub, const void* const* comps, int k, size_t n, void* out,
                                            size_t cap, size_t* len);
COHOST_API int cohost_shared_witness_decode(int curve, const void* data, size_t len, int k, size_t* n_pub, size_t* n, void* pub,
                                            void* const* comps);
COHOST_API int cohost_split_witness_rep3(int curve, int device, const void* witness, size_t n, const uint8_t* seed, void* const* out_a,
                                         void* const* out_b);
/* .r1cs header: info[6] = curve, n_wires, n_pub_out, n_pub_in, n_constraints, num_inputs (circom-types/src/r1cs.rs:100-215).  No GPU. */
COHOST_API int cohost_r1cs_info(const char* path, size_t* info);
/* `co-circom split-witness` (co-circom/src/bin/co-circom.rs:160-256): writes <out_dir>/<witness file name>.<i>.shared for every party.
 * protocol 0 = REP3 (threshold 1, 3 parties), 1 = Shamir; the random part comes from the GPU PRF keyed by seed (32 * max(2, threshold) bytes). */
COHOST_API int cohost_split_witness_files(const char* witness_path, const char* r1cs_path, int protocol, int curve, int threshold, int num_parties,
                                          const uint8_t* seed, const char* out_dir, int device);
/* Groth16 verification on the host (pairing over BN254 / BLS12-381; no GPU needed): the check the reference runs after every proof
 * (co-groth16/src/verifier.rs:23-43) and behind `co-circom verify` (co-circom/src/bin/co-circom.rs:640-720).  *ok = 1 accepted, 0 rejected;
 * a non-zero return means malformed input (counts, a point off the curve or outside the subgroup, unparsable JSON).
 * vk: alpha_g1 | beta_g2 | gamma_g2 | delta_g2 packed affine Montgomery; ic: n_ic G1 points; proof: A | B | C; pub: n_ic - 1 Montgomery Fr. */
COHOST_API int cohost_groth16_verify(int curve, const void* vk, const void* ic, size_t n_ic, const void* proof, const void* pub, int* ok);
COHOST_API int cohost_groth16_verify_json(const char* vk_json, size_t vk_len, const char* proof_json, size_t proof_len, const char* public_json,
                                          size_t public_len, int* ok);
/* Plonk verification on the host (co-plonk/src/plonk.rs:123-283: Keccak-256 transcript, challenges, r0 / D / E / F, one pairing equation).
 * challenges_out: NULL, or room for 6 Montgomery Fr = alpha, beta, gamma, xi, v[0], u (the reference's challenge KAT, plonk.rs:285-350). */
COHOST_API int cohost_plonk_verify_json(const char* vk_json, size_t vk_len, const char* proof_json, size_t proof_len, const char* public_json,
                                        size_t public_len, void* challenges_out, int* ok);
/* Plonk zkey header, host only: info[7] = curve, n_vars, n_public, domain_size, n_additions, n_constraints, parts mask (1 = verifying-key
 * tail, 2 = selector sections, 4 = sigma, 8 = Lagrange); k1k2: 2 Montgomery Fr; vk_g1: Qm Ql Qr Qo Qc S1 S2 S3 packed affine Montgomery;
 * x_2: G2 (circom-types/src/plonk/zkey.rs:329-420).  Output pointers may be NULL. */
COHOST_API int cohost_plonk_zkey_header(const char* path, size_t* info, void* k1k2, void* vk_g1, void* x_2);
/* Batched witness-extension arithmetic: the field opcodes of the reference's MPC VM (circom-mpc-vm/src/mpc_vm.rs:508-546:
 * Add Sub Mul Neg Div, dispatched as Rep3VmType::{add, sub, mul, neg, div}, mpc-core/src/protocols/rep3/witness_extension_impl.rs:81-200)
 * as SIMD over a batch of `batch` independent inputs of one circuit: every register holds `batch` Fr elements in HBM, public or
 * secret-shared; one opcode = one kernel over the batch; a shared multiplication / inversion = ONE network round for the whole batch.
 * protocol 0 = plain (seeds: 32 bytes), 1 = three REP3 parties on three threads (seeds: 3 x 32 bytes).  A straight-line program over
 * registers stands in for the circom bytecode (no circom front end exists here). */
typedef struct cohost_vm cohost_vm;
typedef struct cohost_vm_instr { int op; int dst; int lhs; int rhs; } cohost_vm_instr;  /* op: 0 Add 1 Sub 2 Mul 3 Neg 4 Div */
COHOST_API int cohost_vm_create(int curve, int device, int protocol, const uint8_t* seeds, size_t batch, int n_regs, cohost_vm** out);
COHOST_API void cohost_vm_destroy(cohost_vm* v);
COHOST_API int cohost_vm_set_public(cohost_vm* v, int reg, const void* values);                            /* batch Montgomery Fr */
COHOST_API int cohost_vm_set_shared(cohost_vm* v, int reg, int party, const void* a, const void* b);       /* the party's components */
COHOST_API int cohost_vm_run(cohost_vm* v, const cohost_vm_instr* prog, size_t n);
/* *kind: 1 public (a_out), 2 shared (a_out | b_out = the party's components) */
COHOST_API int cohost_vm_get(cohost_vm* v, int reg, int party, int* kind, void* a_out, void* b_out);
/* out[0] = kernel launches so far, out[1] = network rounds so far */
COHOST_API int cohost_vm_stats(cohost_vm* v, uint64_t* out);
/* (offset, length) of the slice of an n-term MSM that `rank` of `world` accumulates (index-range sharding; needs no GPU). */
COHOST_API int cohost_msm_shard_range(size_t n, int rank, int world, size_t* off, size_t* len);

#ifdef __cplusplus
}
#endif
#endif /* COHOST_H */
