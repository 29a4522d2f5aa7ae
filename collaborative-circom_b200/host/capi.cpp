// C entry points of the host layer (libcohost.so): device-resident zkeys and proving sessions that run the reference's
// prover structure (CoGroth16<T> over PlainDriver / three Rep3Protocol drivers on three threads with an in-process
// Rep3TestNetwork, exactly how /root/reference/tests/tests/circom/e2e_tests/mod.rs:55-70 and
// tests/benches/poseidon_hash2.rs:197-222 run it) on top of libcocg.so.  Declared in include/cohost.h.
#include <algorithm>
#include <array>
#include <cstdlib>
#include <thread>

#include "../../include/cohost.h"
#include "groth16.hpp"
#include "shamir.hpp"
#include "formats.hpp"
#include "plonk.hpp"
#include "serialize.hpp"
#include "verify_json.hpp"
#include "plonk_verify.hpp"
#include "vm.hpp"

using namespace cohost;

namespace {
thread_local std::string g_err;
int fail(const std::string& m) {
  g_err = m;
  return 1;
}
template <class Fn>
int guarded(Fn fn) {
  try {
    fn();
    return 0;
  } catch (const std::exception& e) {
    return fail(e.what());
  }
}
Point load_point(const void* p, size_t limbs) {
  Point r;
  memcpy(r.l, p, limbs * 8);
  return r;
}
}  // namespace

struct cohost_zkey {
  ZKey zk;
  int device = 0;
  size_t lq = 4;
};

struct cohost_rep3_session {
  cohost_zkey* zkey = nullptr;
  std::unique_ptr<Rep3TestNetwork> net;
  std::unique_ptr<Rep3Protocol> drv[3];
  std::unique_ptr<CoGroth16<Rep3Protocol>> prover[3];
  CoGroth16<Rep3Protocol>::Handles hd[3];
  CoGroth16<Rep3Protocol>::Handles hd_aux[3];  // the same queries in the parties' aux contexts (single GPU: MSMs overlap the witness map)
  DevVec pub[3];
  std::unique_ptr<DeviceBridge> bridge;  // block mode: cross-GPU leg of the witness maps' mul_vec rounds
  // multi-GPU two-phase state
  int world = 1, rank = 0;
  std::mutex mu, upload_mu;
  std::condition_variable cv;
  int partials_ready = 0;
  bool combined_ready = false;
  bool abort_wait = false;
  MsmPartials partials[3];
  // in-flight prove
  std::thread th[3];
  std::string errs[3];
  Groth16Proof proofs[3];
  bool running = false;
  bool failed = false;
};

struct cohost_plain_session {
  cohost_zkey* zkey = nullptr;
  std::unique_ptr<PlainDriver> drv;
  std::unique_ptr<CoGroth16<PlainDriver>> prover;
  CoGroth16<PlainDriver>::Handles hd;
};

extern "C" const char* cohost_last_error(void) { return g_err.c_str(); }

extern "C" int cohost_zkey_create(const cohost_zkey_desc* d, cohost_zkey** out) {
  if (!d || !out) return fail("cohost_zkey_create: null argument");
  *out = nullptr;
  return guarded([&] {
    std::unique_ptr<cohost_zkey> z(new cohost_zkey());
    ZKey& zk = z->zk;
    zk.curve = d->curve;
    z->device = d->device;
    z->lq = d->curve == COCG_BN254 ? 4 : 6;
    const size_t lq = z->lq;
    if (cocg_create(&zk.owner, d->device, d->curve)) throw Error(std::string("cocg_create: ") + cocg_last_error(nullptr));
    zk.n_public = d->n_public;
    zk.n_vars = d->n_vars;
    zk.pow = d->pow;
    zk.num_constraints = d->num_constraints;
    if (zk.n_vars < zk.n_public + 1) throw Error("zkey: n_vars < n_public + 1");
    // spmv_kernel gathers z[col] without a bound: a column index past the assignment must be refused here (the reference panics on
    // the Rust bounds check, co-groth16/src/groth16.rs:159-166)
    for (size_t k = 0; k < d->a_nnz; k++) if (d->a_col[k] >= zk.n_vars) throw Error("zkey: column index of matrix A out of range");
    for (size_t k = 0; k < d->b_nnz; k++) if (d->b_col[k] >= zk.n_vars) throw Error("zkey: column index of matrix B out of range");
    const size_t m = zk.n_vars, l = zk.n_public;
    const size_t g1b = 2 * lq * 8, g2b = 4 * lq * 8;
    cocg_ctx* c = zk.owner;
    // a query is uploaded from the caller's array, or generated in HBM when the pointer is NULL and a seed is given; with
    // world > 1 only this rank's slice [first, first + n) becomes resident (its window table is sized for the slice)
    auto bases = [&](int group, const void* pts, size_t first, size_t n, int salt, uint64_t* handle, const char* what) {
      const size_t pb = group == COCG_G1 ? g1b : g2b;
      if (pts) {
        check(c, cocg_bases_upload(c, group, (const char*)pts + first * pb, n, pb, 1, handle), what);
      } else {
        if (!d->synthetic_seed) throw Error(std::string("zkey: ") + what + " is NULL and no synthetic_seed was given");
        uint8_t sd[32];
        memcpy(sd, d->synthetic_seed, 32);
        sd[31] ^= (uint8_t)salt;
        check(c, cocg_bases_generate_range(c, group, first, n, sd, handle), what);
      }
    };
    zk.rank = d->world > 1 ? d->rank : 0;
    zk.world = d->world > 1 ? d->world : 1;
    if (zk.rank < 0 || zk.rank >= zk.world) throw Error("zkey: bad rank / world");
    zk.blocks = zk.world > 1 && d->shard_mode == 1;
    if (zk.blocks) zk.plan = BlockPlan::make(zk.world);
    size_t len_h = zk.domain_size(), len_x = zk.n_aux();
    if (zk.world > 1 && !zk.blocks) {
      MsmShard sh;
      sh.rank = zk.rank;
      sh.world = zk.world;
      size_t off_x;
      sh.range(zk.domain_size(), zk.h_first, len_h);
      sh.range(zk.n_aux(), off_x, len_x);
      zk.l_first = off_x;
      zk.a_first = zk.b_g1_first = zk.b_g2_first = 1 + l + off_x;
    }
    const size_t len_q = (zk.world > 1 && !zk.blocks) ? len_x : m;  // a / b queries: whole (heads included) or the aux slice
    // block mode: whole queries, but only those of the blocks this rank runs (types.hpp BlockPlan)
    const bool want_l = !zk.blocks || zk.plan.has_g1(zk.rank, 0), want_a = !zk.blocks || zk.plan.has_g1(zk.rank, 1), want_b1 = !zk.blocks || zk.plan.has_g1(zk.rank, 2);
    const bool want_g2 = !zk.blocks || zk.plan.has_g2(zk.rank), want_h = !zk.blocks || zk.plan.has_wm(zk.rank);
    if (want_a) bases(COCG_G1, d->a_query, zk.a_first, len_q, 1, &zk.a_query, "a_query");
    if (want_b1) bases(COCG_G1, d->b_g1_query, zk.b_g1_first, len_q, 2, &zk.b_g1_query, "b_g1_query");
    if (want_g2) bases(COCG_G2, d->b_g2_query, zk.b_g2_first, len_q, 3, &zk.b_g2_query, "b_g2_query");
    if (want_h) bases(COCG_G1, d->h_query, zk.h_first, len_h, 4, &zk.h_query, "h_query");
    if (want_l) bases(COCG_G1, d->l_query, zk.l_first, len_x, 5, &zk.l_query, "l_query");
    check(c, cocg_csr_upload_form(c, d->a_rowptr, d->a_col, d->a_coeff, zk.num_constraints, d->a_nnz, d->coeff_form, &zk.csr_a), "csr_a");
    check(c, cocg_csr_upload_form(c, d->b_rowptr, d->b_col, d->b_coeff, zk.num_constraints, d->b_nnz, d->coeff_form, &zk.csr_b), "csr_b");
    zk.a_head.resize(l + 1);
    zk.b_g1_head.resize(l + 1);
    zk.b_g2_head.resize(l + 1);
    {  // the first 1 + l points of the coefficient queries, host copies (from the array, or a tiny generated prefix)
      std::vector<uint64_t> buf((l + 1) * 4 * lq);
      auto heads = [&](int group, const void* pts, int salt, size_t limbs, std::vector<Point>& dst, const char* what) {
        if (pts) {
          memcpy(buf.data(), pts, (l + 1) * limbs * 8);
        } else {
          uint64_t hh = 0;
          bases(group, nullptr, 0, l + 1, salt, &hh, what);
          check(c, cocg_bases_download(c, hh, 0, l + 1, buf.data()), what);
          cocg_bases_free(c, hh);
        }
        for (size_t i = 0; i <= l; i++) dst[i] = load_point(buf.data() + i * limbs, limbs);
      };
      heads(COCG_G1, d->a_query, 1, 2 * lq, zk.a_head, "a_query head");
      heads(COCG_G1, d->b_g1_query, 2, 2 * lq, zk.b_g1_head, "b_g1_query head");
      heads(COCG_G2, d->b_g2_query, 3, 4 * lq, zk.b_g2_head, "b_g2_query head");
    }
    // vk points: given, or three G1 / two G2 synthetic points
    uint64_t vk1[3][12], vk2[2][24];
    if (!d->alpha_g1 || !d->beta_g1 || !d->delta_g1 || !d->beta_g2 || !d->delta_g2) {
      uint64_t h1 = 0, h2 = 0;
      bases(COCG_G1, nullptr, 0, 3, 6, &h1, "vk g1");
      bases(COCG_G2, nullptr, 0, 2, 7, &h2, "vk g2");
      std::vector<uint64_t> b1(3 * 2 * lq), b2(2 * 4 * lq);
      check(c, cocg_bases_download(c, h1, 0, 3, b1.data()), "vk g1");
      check(c, cocg_bases_download(c, h2, 0, 2, b2.data()), "vk g2");
      for (int i = 0; i < 3; i++) memcpy(vk1[i], b1.data() + i * 2 * lq, 2 * lq * 8);
      for (int i = 0; i < 2; i++) memcpy(vk2[i], b2.data() + i * 4 * lq, 4 * lq * 8);
      cocg_bases_free(c, h1);
      cocg_bases_free(c, h2);
    }
    zk.alpha_g1 = load_point(d->alpha_g1 ? d->alpha_g1 : vk1[0], 2 * lq);
    zk.beta_g1 = load_point(d->beta_g1 ? d->beta_g1 : vk1[1], 2 * lq);
    zk.delta_g1 = load_point(d->delta_g1 ? d->delta_g1 : vk1[2], 2 * lq);
    zk.beta_g2 = load_point(d->beta_g2 ? d->beta_g2 : vk2[0], 4 * lq);
    zk.delta_g2 = load_point(d->delta_g2 ? d->delta_g2 : vk2[1], 4 * lq);
    *out = z.release();
  });
}

extern "C" void cohost_zkey_destroy(cohost_zkey* z) {
  if (!z) return;
  if (z->zk.owner) cocg_destroy(z->zk.owner);
  delete z;
}

template <class H>
static void share_handles(cocg_ctx* ctx, const ZKey& zk, H& hd) {
  hd.a_query = hd.b_g1_query = hd.b_g2_query = hd.h_query = hd.l_query = 0;  // 0 = not resident on this rank (block mode)
  if (zk.a_query) check(ctx, cocg_bases_share(ctx, zk.owner, zk.a_query, &hd.a_query), "share a_query");
  if (zk.b_g1_query) check(ctx, cocg_bases_share(ctx, zk.owner, zk.b_g1_query, &hd.b_g1_query), "share b_g1_query");
  if (zk.b_g2_query) check(ctx, cocg_bases_share(ctx, zk.owner, zk.b_g2_query, &hd.b_g2_query), "share b_g2_query");
  if (zk.h_query) check(ctx, cocg_bases_share(ctx, zk.owner, zk.h_query, &hd.h_query), "share h_query");
  if (zk.l_query) check(ctx, cocg_bases_share(ctx, zk.owner, zk.l_query, &hd.l_query), "share l_query");
  check(ctx, cocg_csr_share(ctx, zk.owner, zk.csr_a, &hd.csr_a), "share csr_a");
  check(ctx, cocg_csr_share(ctx, zk.owner, zk.csr_b, &hd.csr_b), "share csr_b");
}

// ------------------------------------------------------------------------------------------------ plain driver
extern "C" int cohost_plain_session_create(cohost_zkey* z, cohost_plain_session** out) {
  if (!z || !out) return fail("cohost_plain_session_create: null argument");
  return guarded([&] {
    std::unique_ptr<cohost_plain_session> s(new cohost_plain_session());
    s->zkey = z;
    s->drv.reset(new PlainDriver(z->zk.curve, z->device));
    s->prover.reset(new CoGroth16<PlainDriver>(*s->drv));
    share_handles(s->drv->ctx, z->zk, s->hd);
    *out = s.release();
  });
}
extern "C" void cohost_plain_session_destroy(cohost_plain_session* s) {
  if (!s) return;
  s->prover->driver.release(s->prover->last_h);
  delete s;
}
extern "C" int cohost_plain_prove(cohost_plain_session* s, const void* public_inputs, const void* witness, const void* r, const void* s_rand,
                                  void* proof_out, void* h_out) {
  if (!s || !public_inputs || !proof_out) return fail("cohost_plain_prove: null argument");
  return guarded([&] {
    const ZKey& zk = s->zkey->zk;
    PlainDriver& d = *s->drv;
    const size_t lq = s->zkey->lq;
    InjectedRandomness inj;
    if (r && s_rand) {
      FieldShare fr_, fs_;
      memcpy(fr_.a.l, r, 32);
      memcpy(fs_.a.l, s_rand, 32);
      fr_.b = fs_.b = d.fr.zero();
      inj.rand = {fr_, fs_};
      d.injected = &inj;
    }
    std::vector<Fr> pub_host(zk.num_inputs());
    memcpy(pub_host.data(), public_inputs, zk.num_inputs() * 32);
    DevVec pub = d.upload(public_inputs, zk.num_inputs());
    FieldShareVec wit = d.share_vec_from_host(witness, nullptr, zk.n_aux());
    Groth16Proof p = s->prover->prove(zk, s->hd, pub, pub_host, wit);
    d.injected = nullptr;
    if (h_out) d.download(s->prover->last_h.a, h_out);
    d.release(pub);
    d.release(wit);
    uint64_t* o = (uint64_t*)proof_out;
    memcpy(o, p.pi_a.l, 2 * lq * 8);
    memcpy(o + 2 * lq, p.pi_b.l, 4 * lq * 8);
    memcpy(o + 6 * lq, p.pi_c.l, 2 * lq * 8);
  });
}

// ------------------------------------------------------------------------------------------------ REP3, three parties
static int rep3_session_create(cohost_zkey* z, const uint8_t* seeds, int rank, int world, cohost_comm_cb comm, void* comm_user, cohost_rep3_session** out);
extern "C" int cohost_rep3_session_create(cohost_zkey* z, const uint8_t* seeds /* 3 x 32 */, int rank, int world, cohost_rep3_session** out) {
  if (z && z->zk.blocks) return fail("cohost_rep3_session_create: the zkey was made for block mode; use cohost_rep3_session_create_blocks");
  return rep3_session_create(z, seeds, rank, world, nullptr, nullptr, out);
}
// Block mode (types.hpp BlockPlan): `comm` carries the witness maps' mul_vec payloads between GPUs -- the caller issues the listed sends
// and receives as ONE grouped NCCL call (ops[i].dptr are device buffers of this rank) and returns 0 when they have completed.
extern "C" int cohost_rep3_session_create_blocks(cohost_zkey* z, const uint8_t* seeds, int rank, int world, cohost_comm_cb comm, void* user,
                                                 cohost_rep3_session** out) {
  if (!z || !z->zk.blocks) return fail("cohost_rep3_session_create_blocks: the zkey was not made with shard_mode = 1");
  if (!comm) return fail("cohost_rep3_session_create_blocks: a transfer callback is required");
  return rep3_session_create(z, seeds, rank, world, comm, user, out);
}
static int rep3_session_create(cohost_zkey* z, const uint8_t* seeds, int rank, int world, cohost_comm_cb comm, void* comm_user, cohost_rep3_session** out) {
  if (!z || !out || !seeds) return fail("cohost_rep3_session_create: null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail("cohost_rep3_session_create: bad rank/world");
  if (z->zk.world != 1 && (z->zk.world != world || z->zk.rank != rank)) return fail("cohost_rep3_session_create: the zkey holds another rank's shard");
  return guarded([&] {
    std::unique_ptr<cohost_rep3_session> s(new cohost_rep3_session());
    s->zkey = z;
    s->rank = rank;
    s->world = world;
    s->net.reset(new Rep3TestNetwork());
    {  // COHOST_MPC_EXCHANGE=device: the three co-located parties pass share vectors in HBM (default: pinned host staging)
      const char* ex = getenv("COHOST_MPC_EXCHANGE");
      s->net->device_exchange = ex && std::string(ex) == "device";
    }
    for (int i = 0; i < 3; i++) s->drv[i].reset(new Rep3Protocol(z->zk.curve, z->device, s->net->party(i), seeds + 32 * i));
    if (z->zk.blocks) {
      static_assert(sizeof(CommOp) == sizeof(cohost_comm_op), "CommOp must mirror cohost_comm_op");
      s->bridge.reset(new DeviceBridge(rank, z->zk.plan.wm, (CommCallback)comm, comm_user));
      s->net->device_exchange = true;  // co-located witness maps hand their payloads over in HBM
    }
    for (int i = 0; i < 3; i++) {
      s->drv[i]->finish_setup();
      s->prover[i].reset(new CoGroth16<Rep3Protocol>(*s->drv[i]));
      s->prover[i]->shard.rank = rank;
      s->prover[i]->shard.world = world;
      if (z->zk.blocks) {
        s->drv[i]->bridge = s->bridge.get();
        s->prover[i]->blocks = &z->zk.plan;
        s->prover[i]->block_rank = rank;
        // the party whose witness map runs here has the longest dependent chain on this rank (witness map, lock-step with two other
        // GPUs, then its h MSMs): its stream's thread blocks go ahead of the other parties' pending MSM blocks.  Only with three or
        // more ranks, where a rank runs at most one witness map.  (world = 2 shows occasional steps whose NCCL exchange round stalls
        // behind the MSM launches, with or without this priority -- DESIGN.md section 6; the priority is kept off there.)
        const char* pr = getenv("COHOST_WM_PRIORITY");
        const bool want = pr ? std::string(pr) == "1" : world >= 3;
        if (z->zk.plan.wm[i] == rank && want) check(s->drv[i]->ctx, cocg_set_stream_priority(s->drv[i]->ctx, 1), "cocg_set_stream_priority");
      }
      share_handles(s->drv[i]->ctx, z->zk, s->hd[i]);
      // single GPU, whole queries resident: the four MSMs over the witness run on a second context per party, concurrently with
      // that party's witness map (COHOST_AUX_MSM=0 keeps everything on one context)
      const char* ax = getenv("COHOST_AUX_MSM");
      if (world == 1 && z->zk.world == 1 && !z->zk.blocks && !(ax && std::string(ax) == "0")) {
        s->drv[i]->enable_aux_ctx(z->device);
        share_handles(s->drv[i]->aux_ctx, z->zk, s->hd_aux[i]);
        s->prover[i]->aux_hd = &s->hd_aux[i];
      }
    }
    cohost_rep3_session* sp = s.get();
    if (world > 1) {
      for (int i = 0; i < 3; i++)
        s->prover[i]->combine = [sp](int party, MsmPartials& m) {
          std::unique_lock<std::mutex> lk(sp->mu);
          sp->partials[party] = m;
          sp->partials_ready++;
          sp->cv.notify_all();
          sp->cv.wait(lk, [&] { return sp->combined_ready || sp->abort_wait; });
          if (sp->abort_wait) throw Error("prove aborted while waiting for the multi-GPU combine");
          m = sp->partials[party];
        };
    }
    *out = s.release();
  });
}

extern "C" void cohost_rep3_session_destroy(cohost_rep3_session* s) {
  if (!s) return;
  if (s->running) {
    {
      std::lock_guard<std::mutex> lk(s->mu);
      s->abort_wait = true;
    }
    s->cv.notify_all();
    s->net->close_all();
    for (auto& t : s->th)
      if (t.joinable()) t.join();
  }
  for (int i = 0; i < 3; i++) {
    if (s->prover[i]) s->drv[i]->release(s->prover[i]->last_h);
    s->drv[i]->release(s->pub[i]);
  }
  delete s;
}

static size_t partial_limbs(size_t lq) { return 8 * 3 * lq + 2 * 6 * lq; }
static void pack_partials(const MsmPartials& m, size_t lq, uint64_t* o) {
  const PointShare* g1[4] = {&m.h_acc, &m.l_acc, &m.a_acc, &m.b1_acc};
  for (int i = 0; i < 4; i++) {
    memcpy(o, g1[i]->a.l, 3 * lq * 8); o += 3 * lq;
    memcpy(o, g1[i]->b.l, 3 * lq * 8); o += 3 * lq;
  }
  memcpy(o, m.b2_acc.a.l, 6 * lq * 8); o += 6 * lq;
  memcpy(o, m.b2_acc.b.l, 6 * lq * 8);
}

// Starts one proof on the three party threads.  wit_a[i] / wit_b[i]: HOST share components of party i (n_aux elements).
// rnd may be NULL (PRF-derived randomness).
static int rep3_prove_begin(cohost_rep3_session* s, const void* public_inputs, const void* const* wit_a, const void* const* wit_b,
                            const cohost_rep3_randomness* rnd, bool wit_on_device);
extern "C" int cohost_rep3_prove_begin(cohost_rep3_session* s, const void* public_inputs, const void* const* wit_a, const void* const* wit_b,
                                       const cohost_rep3_randomness* rnd) {
  return rep3_prove_begin(s, public_inputs, wit_a, wit_b, rnd, false);
}
extern "C" int cohost_rep3_prove_begin_device(cohost_rep3_session* s, const void* public_inputs, const void* const* wit_a,
                                              const void* const* wit_b, const cohost_rep3_randomness* rnd) {
  return rep3_prove_begin(s, public_inputs, wit_a, wit_b, rnd, true);
}
static int rep3_prove_begin(cohost_rep3_session* s, const void* public_inputs, const void* const* wit_a, const void* const* wit_b,
                            const cohost_rep3_randomness* rnd, bool wit_on_device) {
  if (!s || !public_inputs || !wit_a || !wit_b) return fail("cohost_rep3_prove_begin: null argument");
  if (s->running) return fail("cohost_rep3_prove_begin: a proof is already in flight");
  if (s->failed) return fail("cohost_rep3_prove_begin: the session's network is closed after an earlier failure");
  const ZKey& zk = s->zkey->zk;
  const size_t lq = s->zkey->lq;
  s->partials_ready = 0;
  s->combined_ready = false;
  s->abort_wait = false;
  s->running = true;
  // the caller's pointer arrays need not outlive this call: copy them before the party threads start
  const void* wa[3] = {wit_a[0], wit_a[1], wit_a[2]};
  const void* wb[3] = {wit_b[0], wit_b[1], wit_b[2]};
  cohost_rep3_randomness rnd_copy;
  if (rnd) rnd_copy = *rnd;
  const bool have_rnd = rnd != nullptr;
  for (int i = 0; i < 3; i++) {
    s->errs[i].clear();
    const void* wit_a_i = wa[i];
    const void* wit_b_i = wb[i];
    s->th[i] = std::thread([=] {
      const cohost_rep3_randomness* rnd = have_rnd ? &rnd_copy : nullptr;
      try {
        Rep3Protocol& d = *s->drv[i];
        InjectedRandomness inj;
        if (rnd) {
          FieldShare r, sv;
          memcpy(r.a.l, (const char*)rnd->r + 64 * i, 32);
          memcpy(r.b.l, (const char*)rnd->r + 64 * i + 32, 32);
          memcpy(sv.a.l, (const char*)rnd->s + 64 * i, 32);
          memcpy(sv.b.l, (const char*)rnd->s + 64 * i + 32, 32);
          inj.rand = {r, sv};
          Fr mrs;
          memcpy(mrs.l, (const char*)rnd->mask_rs + 32 * i, 32);
          inj.mul_masks = {mrs};
          inj.mul_vec_masks = {rnd->masks1[i], rnd->masks2[i]};
          inj.ec_masks = {load_point((const char*)rnd->mask_pt + 3 * lq * 8 * i, 3 * lq)};
          d.injected = &inj;
        }
        std::vector<Fr> pub_host(zk.num_inputs());
        memcpy(pub_host.data(), public_inputs, zk.num_inputs() * 32);
        d.release(s->pub[i]);
        s->pub[i] = d.upload(public_inputs, zk.num_inputs());
        FieldShareVec wit;
        bool need[2] = {true, true};  // block mode: only the components of this party that a block of this rank reads
        if (zk.blocks)
          for (int c = 0; c < 2; c++) need[c] = zk.plan.reads_witness(s->rank, i, c);
        if (wit_on_device) {  // borrowed: the caller keeps ownership
          wit.a = DevVec{const_cast<void*>(wit_a_i), zk.n_aux()};
          wit.b = DevVec{const_cast<void*>(wit_b_i), zk.n_aux()};
        } else {
          // one party at a time: the three uploads share one PCIe link anyway, and the party that finishes first can start its
          // witness MSMs (aux context) while the others are still uploading
          std::lock_guard<std::mutex> up(s->upload_mu);
          if (need[0]) wit.a = d.upload(wit_a_i, zk.n_aux());
          if (need[1]) wit.b = d.upload(wit_b_i, zk.n_aux());
        }
        s->proofs[i] = s->prover[i]->prove(zk, s->hd[i], s->pub[i], pub_host, wit);
        if (!wit_on_device) d.release(wit);
        d.injected = nullptr;
      } catch (const std::exception& e) {
        s->errs[i] = e.what();
        s->drv[i]->injected = nullptr;
        s->net->close_all();  // unblock the peers (a dead party aborts the proof)
        if (s->bridge) s->bridge->abort();
        {
          std::lock_guard<std::mutex> lk(s->mu);
          s->abort_wait = true;
        }
        s->cv.notify_all();
      }
    });
  }
  return 0;
}

// Multi-GPU: blocks until the three parties have produced this rank's partial MSM sums; copies them out
// (3 x cohost_rep3_partial_bytes).  The caller all-gathers these over ranks (one NCCL all-gather per proof).
extern "C" size_t cohost_rep3_partial_bytes(cohost_rep3_session* s) { return s ? 3 * partial_limbs(s->zkey->lq) * 8 : 0; }
extern "C" int cohost_rep3_prove_partials(cohost_rep3_session* s, void* out) {
  if (!s || !out) return fail("cohost_rep3_prove_partials: null argument");
  if (!s->running || s->world == 1) return fail("cohost_rep3_prove_partials: no sharded proof in flight");
  std::unique_lock<std::mutex> lk(s->mu);
  s->cv.wait(lk, [&] { return s->partials_ready == 3 || s->abort_wait; });
  if (s->abort_wait) return fail("cohost_rep3_prove_partials: a party failed");
  const size_t pl = partial_limbs(s->zkey->lq);
  for (int i = 0; i < 3; i++) pack_partials(s->partials[i], s->zkey->lq, (uint64_t*)out + pl * i);
  return 0;
}
// gathered: world x (3 x partial) buffers in rank order.  Folds them (group additions) and releases the party threads.
extern "C" int cohost_rep3_prove_combine(cohost_rep3_session* s, const void* gathered) {
  if (!s || !gathered) return fail("cohost_rep3_prove_combine: null argument");
  return guarded([&] {
    const size_t lq = s->zkey->lq, pl = partial_limbs(lq);
    Rep3Protocol& d = *s->drv[0];
    std::unique_lock<std::mutex> lk(s->mu);
    for (int i = 0; i < 3; i++) {
      MsmPartials& m = s->partials[i];
      PointShare* g1[4] = {&m.h_acc, &m.l_acc, &m.a_acc, &m.b1_acc};
      for (int q = 0; q < 5; q++) {
        const int g = q < 4 ? 1 : 2;
        PointShare* ps = q < 4 ? g1[q] : &m.b2_acc;
        const size_t off = q < 4 ? (size_t)q * 6 * lq : 24 * lq, nl = 3 * g * lq;
        for (int comp = 0; comp < 2; comp++) {
          Point acc = d.infinity(g);
          for (int r = 0; r < s->world; r++) {
            Point p = load_point((const uint64_t*)gathered + ((size_t)r * 3 + i) * pl + off + comp * nl, nl);
            acc = d.ec_add(g, acc, p);
          }
          (comp ? ps->b : ps->a) = acc;
        }
      }
    }
    s->combined_ready = true;
    s->cv.notify_all();
  });
}

// Joins the party threads.  proofs_out: 3 x (A | B | C) packed affine; h_a/h_b: optional HOST buffers (3 each) for h.
extern "C" int cohost_rep3_prove_end(cohost_rep3_session* s, void* proofs_out, void* const* h_a, void* const* h_b) {
  if (!s || !proofs_out) return fail("cohost_rep3_prove_end: null argument");
  if (!s->running) return fail("cohost_rep3_prove_end: no proof in flight");
  for (auto& t : s->th)
    if (t.joinable()) t.join();
  s->running = false;
  for (int i = 0; i < 3; i++)
    if (!s->errs[i].empty()) {
      std::string first = s->errs[i];
      s->failed = true;  // channels were closed: the session is not reusable after a failure
      return fail("party " + std::to_string(i) + ": " + first);
    }
  return guarded([&] {
    const size_t lq = s->zkey->lq;
    for (int i = 0; i < 3; i++) {
      uint64_t* o = (uint64_t*)proofs_out + (size_t)i * 8 * lq;
      memcpy(o, s->proofs[i].pi_a.l, 2 * lq * 8);
      memcpy(o + 2 * lq, s->proofs[i].pi_b.l, 4 * lq * 8);
      memcpy(o + 6 * lq, s->proofs[i].pi_c.l, 2 * lq * 8);
      if (h_a && h_a[i]) s->drv[i]->download(s->prover[i]->last_h.a, h_a[i]);
      if (h_b && h_b[i]) s->drv[i]->download(s->prover[i]->last_h.b, h_b[i]);
    }
  });
}

extern "C" uint64_t cohost_rep3_launch_count(cohost_rep3_session* s) {
  uint64_t t = 0;
  if (s)
    for (int i = 0; i < 3; i++) t += cocg_launch_count(s->drv[i]->ctx) + (s->drv[i]->aux_ctx ? cocg_launch_count(s->drv[i]->aux_ctx) : 0);
  return t;
}

extern "C" int cohost_rep3_profile_enable(cohost_rep3_session* s, int on) {
  if (!s) return fail("null session");
  for (int i = 0; i < 3; i++) {
    cocg_profile_enable(s->drv[i]->ctx, on);
    if (s->drv[i]->aux_ctx) cocg_profile_enable(s->drv[i]->aux_ctx, on);
  }
  return 0;
}
extern "C" int cohost_rep3_profile_read(cohost_rep3_session* s, int cls, double* total_ms, uint64_t* scopes) {
  if (!s) return fail("null session");
  double t = 0;
  uint64_t n = 0;
  for (int i = 0; i < 3; i++) {
    double ms = 0;
    uint64_t k = 0;
    if (cocg_profile_read(s->drv[i]->ctx, cls, &ms, &k)) return fail(cocg_last_error(s->drv[i]->ctx));
    t += ms;
    n += k;
    if (s->drv[i]->aux_ctx) {
      if (cocg_profile_read(s->drv[i]->aux_ctx, cls, &ms, &k)) return fail(cocg_last_error(s->drv[i]->aux_ctx));
      t += ms;
      n += k;
    }
  }
  if (total_ms) *total_ms = t;
  if (scopes) *scopes = n;
  return 0;
}
extern "C" int cohost_rep3_profile_reset(cohost_rep3_session* s) {
  if (!s) return fail("null session");
  for (int i = 0; i < 3; i++) {
    cocg_profile_reset(s->drv[i]->ctx);
    if (s->drv[i]->aux_ctx) cocg_profile_reset(s->drv[i]->aux_ctx);
  }
  return 0;
}

extern "C" int cohost_block_plan(int world, int* out) {
  if (world < 1 || !out) return fail("cohost_block_plan: bad argument");
  BlockPlan p = BlockPlan::make(world);
  for (int q = 0; q < 3; q++) out[q] = p.wm[q];
  for (int q = 0; q < 3; q++)
    for (int c = 0; c < 2; c++) {
      out[3 + 2 * q + c] = p.g2[q][c];
      for (int k = 0; k < 3; k++) out[9 + 6 * q + 3 * c + k] = p.g1[q][c][k];
    }
  return 0;
}
// The index-range partition of an n-term MSM over `world` ranks (MsmShard::range); no GPU needed.
extern "C" int cohost_msm_shard_range(size_t n, int rank, int world, size_t* off, size_t* len) {
  if (!off || !len || world < 1 || rank < 0 || rank >= world) return fail("cohost_msm_shard_range: bad argument");
  MsmShard sh;
  sh.rank = rank;
  sh.world = world;
  sh.range(n, *off, *len);
  return 0;
}

// ------------------------------------------------------------------------------------------------ Shamir, n parties
struct cohost_shamir_session {
  cohost_zkey* zkey = nullptr;
  int n = 3, t = 1;
  std::unique_ptr<ShamirTestNetwork> net;
  std::vector<std::unique_ptr<ShamirProtocol>> drv;
  std::vector<std::unique_ptr<CoGroth16<ShamirProtocol>>> prover;
  std::vector<CoGroth16<ShamirProtocol>::Handles> hd;
  bool failed = false;
  // multi-GPU (index-range sharded MSMs, SURVEY 8(e) / BASELINE configs[4]): the parties' partial sums are exchanged by ONE all-gather per
  // proof, performed by the caller's callback on the thread of the last party to arrive
  int rank = 0, world = 1;
  cohost_gather_cb gather = nullptr;
  void* gather_user = nullptr;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  bool combined = false, abort_wait = false;
  std::vector<MsmPartials> partials;
};

extern "C" int cohost_shamir_session_create(cohost_zkey* z, int num_parties, int threshold, const uint8_t* seeds, cohost_shamir_session** out) {
  if (!z || !out || !seeds) return fail("cohost_shamir_session_create: null argument");
  if (num_parties < 3 || num_parties > 64) return fail("Shamir protocol requires at least 3 parties");  // shamir/network.rs:75-77
  return guarded([&] {
    std::unique_ptr<cohost_shamir_session> s(new cohost_shamir_session());
    s->zkey = z;
    s->n = num_parties;
    s->t = threshold;
    s->net.reset(new ShamirTestNetwork(num_parties));
    s->hd.resize(num_parties);
    for (int i = 0; i < num_parties; i++) {
      s->drv.emplace_back(new ShamirProtocol(z->zk.curve, z->device, threshold, s->net->party(i), seeds + 32 * i));
      s->prover.emplace_back(new CoGroth16<ShamirProtocol>(*s->drv[i]));
      share_handles(s->drv[i]->ctx, z->zk, s->hd[i]);
    }
    *out = s.release();
  });
}
// Index-range sharding of every MSM over `world` GPUs (one process per GPU; the zkey must have been created with the same rank / world).
// gather(user, local, bytes, gathered): must fill `gathered` (world x bytes, rank order) with every rank's `local` -- one all-gather.
extern "C" int cohost_shamir_session_set_shard(cohost_shamir_session* s, int rank, int world, cohost_gather_cb gather, void* user) {
  if (!s) return fail("cohost_shamir_session_set_shard: null session");
  if (world < 1 || rank < 0 || rank >= world) return fail("cohost_shamir_session_set_shard: bad rank / world");
  if (world > 1 && !gather) return fail("cohost_shamir_session_set_shard: a gather callback is required when world > 1");
  if (s->zkey->zk.world != 1 && (s->zkey->zk.world != world || s->zkey->zk.rank != rank)) return fail("cohost_shamir_session_set_shard: the zkey holds another rank's shard");
  s->rank = rank;
  s->world = world;
  s->gather = gather;
  s->gather_user = user;
  s->partials.assign(s->n, MsmPartials());
  cohost_shamir_session* sp = s;
  for (int i = 0; i < s->n; i++) {
    s->prover[i]->shard.rank = rank;
    s->prover[i]->shard.world = world;
    if (world == 1) { s->prover[i]->combine = nullptr; continue; }
    s->prover[i]->combine = [sp](int party, MsmPartials& m) {
      const size_t lq = sp->zkey->lq, pl = 18 * lq;  // h, l, a, b1 (G1 Jacobian) | b2 (G2 Jacobian), one component
      std::unique_lock<std::mutex> lk(sp->mu);
      sp->partials[party] = m;
      if (++sp->arrived == sp->n) {
        std::vector<uint64_t> local((size_t)sp->n * pl), all((size_t)sp->world * sp->n * pl);
        for (int i = 0; i < sp->n; i++) {
          const MsmPartials& q = sp->partials[i];
          const Point* g1[4] = {&q.h_acc.a, &q.l_acc.a, &q.a_acc.a, &q.b1_acc.a};
          for (int k = 0; k < 4; k++) memcpy(local.data() + i * pl + k * 3 * lq, g1[k]->l, 3 * lq * 8);
          memcpy(local.data() + i * pl + 12 * lq, q.b2_acc.a.l, 6 * lq * 8);
        }
        int rc = sp->gather(sp->gather_user, local.data(), local.size() * 8, all.data());
        if (rc) {
          sp->abort_wait = true;
          sp->cv.notify_all();
          throw Error("the multi-GPU gather callback failed");
        }
        ShamirProtocol& d = *sp->drv[0];
        for (int i = 0; i < sp->n; i++) {
          MsmPartials& q = sp->partials[i];
          Point* g1[4] = {&q.h_acc.a, &q.l_acc.a, &q.a_acc.a, &q.b1_acc.a};
          for (int k = 0; k < 5; k++) {
            const int g = k < 4 ? 1 : 2;
            Point acc = d.infinity(g);
            for (int r = 0; r < sp->world; r++)
              acc = d.ec_add(g, acc, load_point(all.data() + ((size_t)r * sp->n + i) * pl + (k < 4 ? k * 3 * lq : 12 * lq), 3 * g * lq));
            (k < 4 ? *g1[k] : q.b2_acc.a) = acc;
          }
        }
        sp->combined = true;
        sp->cv.notify_all();
      } else {
        sp->cv.wait(lk, [&] { return sp->combined || sp->abort_wait; });
        if (sp->abort_wait) throw Error("prove aborted while waiting for the multi-GPU combine");
      }
      m = sp->partials[party];
    };
  }
  return 0;
}
// 0 = share vectors staged through pinned host memory (default), 1 = handed over in HBM (the parties share the session's GPU)
extern "C" int cohost_shamir_set_mpc_exchange(cohost_shamir_session* s, int device) {
  if (!s) return fail("cohost_shamir_set_mpc_exchange: null session");
  s->net->device_exchange = device != 0;
  return 0;
}
extern "C" void cohost_shamir_session_destroy(cohost_shamir_session* s) {
  if (!s) return;
  for (size_t i = 0; i < s->drv.size(); i++) s->drv[i]->release(s->prover[i]->last_h);
  delete s;
}
// wit[i]: party i's HOST share vector (m - l - 1 Fr).  proofs_out: n x (A | B | C); rs_out: NULL or n x (r share | s share) --
// the parties' degree-t shares of the prover randomness, so a test can reconstruct (r, s) and compare with the plain prover.
extern "C" int cohost_shamir_prove(cohost_shamir_session* s, const void* public_inputs, const void* const* wit, void* proofs_out, void* rs_out) {
  if (!s || !public_inputs || !wit || !proofs_out) return fail("cohost_shamir_prove: null argument");
  if (s->failed) return fail("cohost_shamir_prove: the session's network is closed after an earlier failure");
  const ZKey& zk = s->zkey->zk;
  const size_t lq = s->zkey->lq;
  std::vector<std::thread> th(s->n);
  std::vector<std::string> errs(s->n);
  std::vector<Groth16Proof> proofs(s->n);
  std::vector<const void*> w(wit, wit + s->n);
  s->arrived = 0;
  s->combined = false;
  s->abort_wait = false;
  for (int i = 0; i < s->n; i++) {
    th[i] = std::thread([&, i] {
      try {
        ShamirProtocol& d = *s->drv[i];
        std::vector<Fr> pub_host(zk.num_inputs());
        memcpy(pub_host.data(), public_inputs, zk.num_inputs() * 32);
        DevVec pub = d.upload(public_inputs, zk.num_inputs());
        FieldShareVec witv = d.share_vec_from_host(w[i], nullptr, zk.n_aux());
        proofs[i] = s->prover[i]->prove(zk, s->hd[i], pub, pub_host, witv);
        d.release(witv);
        d.release(pub);
      } catch (const std::exception& e) {
        errs[i] = e.what();
        s->net->close_all();
        {
          std::lock_guard<std::mutex> lk(s->mu);
          s->abort_wait = true;
        }
        s->cv.notify_all();
      }
    });
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < s->n; i++)
    if (!errs[i].empty()) {
      s->failed = true;
      return fail("party " + std::to_string(i) + ": " + errs[i]);
    }
  for (int i = 0; i < s->n; i++) {
    uint64_t* o = (uint64_t*)proofs_out + (size_t)i * 8 * lq;
    memcpy(o, proofs[i].pi_a.l, 2 * lq * 8);
    memcpy(o + 2 * lq, proofs[i].pi_b.l, 4 * lq * 8);
    memcpy(o + 6 * lq, proofs[i].pi_c.l, 2 * lq * 8);
    if (rs_out) {
      memcpy((char*)rs_out + 64 * i, s->prover[i]->last_r.a.l, 32);
      memcpy((char*)rs_out + 64 * i + 32, s->prover[i]->last_s.a.l, 32);
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ snarkjs files
namespace {
template <class CP>
void check_vk_points_impl(const void* const* g1s, int n1, const void* const* g2s, int n2) {
  using E = typename CP::E;
  for (int i = 0; i < n1; i++) {
    typename E::G1 p;
    memcpy(&p, g1s[i], sizeof(p));
    if (p.is_inf()) continue;
    if (!on_curve(p, CP::b1()) || !scalar_mul_affine(p, CP::order()).is_inf()) throw Error("zkey: InvalidData: a verifying-key G1 point is not on the curve / in the subgroup");
  }
  for (int i = 0; i < n2; i++) {
    typename E::G2 p;
    memcpy(&p, g2s[i], sizeof(p));
    if (p.is_inf()) continue;
    if (!on_curve(p, CP::b2()) || !scalar_mul_affine(p, CP::order()).is_inf()) throw Error("zkey: InvalidData: a verifying-key G2 point is not on the curve / in the subgroup");
  }
}
void check_vk_points(int curve, const void* a1, const void* b1, const void* d1, const void* b2, const void* d2) {
  const void* g1s[3] = {a1, b1, d1};
  const void* g2s[2] = {b2, d2};
  if (curve == COCG_BN254) check_vk_points_impl<Bn254Pairing>(g1s, 3, g2s, 2);
  else check_vk_points_impl<Bls381Pairing>(g1s, 3, g2s, 2);
}
}  // namespace

// ZKey::from_reader (circom-types/src/groth16/zkey.rs:109, :139-251): parse a Groth16 .zkey image and make it resident in HBM.
extern "C" int cohost_zkey_load(const void* data, size_t len, int device, cohost_zkey** out) {
  if (!data || !out) return fail("cohost_zkey_load: null argument");
  *out = nullptr;
  return guarded([&] {
    Groth16ZKeyFile f((const uint8_t*)data, len);
    cohost_zkey_desc d;
    memset(&d, 0, sizeof(d));
    d.curve = f.curve;
    d.device = device;
    d.n_public = f.n_public;
    d.n_vars = f.n_vars;
    d.pow = f.pow;
    d.num_constraints = f.num_constraints;
    d.a_rowptr = f.rowptr[0].data(); d.a_col = f.col[0].data(); d.a_coeff = f.coeff_raw[0].data(); d.a_nnz = f.col[0].size();
    d.b_rowptr = f.rowptr[1].data(); d.b_col = f.col[1].data(); d.b_coeff = f.coeff_raw[1].data(); d.b_nnz = f.col[1].size();
    d.coeff_form = COCG_FORM_R2;
    d.a_query = f.a_query; d.b_g1_query = f.b_g1_query; d.b_g2_query = f.b_g2_query; d.h_query = f.h_query; d.l_query = f.l_query;
    d.alpha_g1 = f.alpha_g1; d.beta_g1 = f.beta_g1; d.delta_g1 = f.delta_g1; d.beta_g2 = f.beta_g2; d.delta_g2 = f.delta_g2;
    if (cohost_zkey_create(&d, out)) throw Error(cohost_last_error());
    // g1_from_bytes / g2_from_bytes (circom-types/src/traits.rs:107-155) reject every point that is off the curve or outside the
    // prime-order subgroup while parsing; here the query arrays are already in HBM and one kernel per query checks them
    try {
      const ZKey& zk = (*out)->zk;
      const std::pair<const char*, uint64_t> qs[5] = {{"A", zk.a_query}, {"B1", zk.b_g1_query}, {"B2", zk.b_g2_query}, {"H", zk.h_query}, {"L", zk.l_query}};
      for (const auto& q : qs) {
        size_t bad = 0, first = 0;
        check(zk.owner, cocg_bases_check(zk.owner, q.second, 1, &bad, &first), "cocg_bases_check");
        if (bad)
          throw Error(std::string("zkey: InvalidData: point ") + std::to_string(first) + " of section " + q.first + " is not on the curve or not in the "
                      "prime-order subgroup (" + std::to_string(bad) + " such points)");
      }
      check_vk_points(f.curve, f.alpha_g1, f.beta_g1, f.delta_g1, f.beta_g2, f.delta_g2);
    } catch (...) {
      cohost_zkey_destroy(*out);
      *out = nullptr;
      throw;
    }
  });
}
extern "C" int cohost_zkey_load_file(const char* path, int device, cohost_zkey** out) {
  if (!path || !out) return fail("cohost_zkey_load_file: null argument");
  std::vector<uint8_t> buf;
  int rc = guarded([&] { buf = read_file(path); });
  if (rc) return rc;
  return cohost_zkey_load(buf.data(), buf.size(), device, out);
}
extern "C" int cohost_zkey_get_info(cohost_zkey* z, cohost_zkey_info* info) {
  if (!z || !info) return fail("cohost_zkey_get_info: null argument");
  info->curve = z->zk.curve;
  info->n_public = z->zk.n_public;
  info->n_vars = z->zk.n_vars;
  info->pow = z->zk.pow;
  info->num_constraints = z->zk.num_constraints;
  return 0;
}
// which: 0 a_query, 1 b_g1_query, 2 b_g2_query, 3 h_query, 4 l_query; packed affine Montgomery points
extern "C" int cohost_zkey_query_download(cohost_zkey* z, int which, size_t off, size_t n, void* out) {
  if (!z || !out) return fail("cohost_zkey_query_download: null argument");
  const uint64_t h[5] = {z->zk.a_query, z->zk.b_g1_query, z->zk.b_g2_query, z->zk.h_query, z->zk.l_query};
  if (which < 0 || which > 4) return fail("cohost_zkey_query_download: unknown query");
  if (z->zk.world != 1) return fail("cohost_zkey_query_download: only this rank's shard of the queries is resident");
  if (cocg_bases_download(z->zk.owner, h[which], off, n, out)) return fail(cocg_last_error(z->zk.owner));
  return 0;
}
// which: 0 = A, 1 = B.  Any of rowptr (rows + 1), col (nnz), coeff (nnz Fr, Montgomery) may be NULL; nnz is always returned.
extern "C" int cohost_zkey_matrix_download(cohost_zkey* z, int which, uint32_t* rowptr, uint32_t* col, void* coeff, size_t* nnz) {
  if (!z) return fail("cohost_zkey_matrix_download: null argument");
  if (which < 0 || which > 1) return fail("cohost_zkey_matrix_download: unknown matrix");
  if (cocg_csr_download(z->zk.owner, which ? z->zk.csr_b : z->zk.csr_a, rowptr, col, coeff, nnz)) return fail(cocg_last_error(z->zk.owner));
  return 0;
}
// alpha_g1 | beta_g1 | delta_g1 (G1 affine) | beta_g2 | delta_g2 (G2 affine)
extern "C" int cohost_zkey_vk_download(cohost_zkey* z, void* out) {
  if (!z || !out) return fail("cohost_zkey_vk_download: null argument");
  const size_t lq = z->lq;
  uint64_t* o = (uint64_t*)out;
  memcpy(o, z->zk.alpha_g1.l, 2 * lq * 8);
  memcpy(o + 2 * lq, z->zk.beta_g1.l, 2 * lq * 8);
  memcpy(o + 4 * lq, z->zk.delta_g1.l, 2 * lq * 8);
  memcpy(o + 6 * lq, z->zk.beta_g2.l, 4 * lq * 8);
  memcpy(o + 10 * lq, z->zk.delta_g2.l, 4 * lq * 8);
  return 0;
}
// Witness::from_reader (witness.rs:51-92): values -> Montgomery limbs (the conversion runs on the zkey's device).
// out: NULL (only count) or n_values x 32 bytes.
extern "C" int cohost_wtns_load_file(cohost_zkey* z, const char* path, void* out, size_t* n_values) {
  if (!z || !path || !n_values) return fail("cohost_wtns_load_file: null argument");
  return guarded([&] {
    std::vector<uint8_t> buf = read_file(path);
    WitnessFile w(buf.data(), buf.size());
    if (w.curve != z->zk.curve) throw Error("wtns: wrong scalar field");
    *n_values = w.n;
    if (!out || w.n == 0) return;
    cocg_ctx* c = z->zk.owner;
    void* d = nullptr;
    check(c, cocg_malloc(c, w.n * 32, &d), "cocg_malloc");
    check(c, cocg_h2d(c, d, w.values, w.n * 32), "cocg_h2d");
    check(c, cocg_vec_op(c, COCG_OP_TO_MONT, d, nullptr, d, w.n), "cocg_vec_op");
    check(c, cocg_d2h(c, out, d, w.n * 32), "cocg_d2h");
    check(c, cocg_free(c, d), "cocg_free");
  });
}

// Host wall-clock per phase of the last proof, seconds: out[party][4] = witness map | MSMs | wait for the all-gather | assembly.
extern "C" int cohost_rep3_phase_times(cohost_rep3_session* s, double* out) {
  if (!s || !out) return fail("cohost_rep3_phase_times: null argument");
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 4; k++) out[4 * i + k] = s->prover[i]->phase_s[k];
  return 0;
}

// ------------------------------------------------------------------------------------------------ Plonk
struct cohost_plonk_zkey {
  PlonkZKey zk;
  size_t lq = 4;
  ~cohost_plonk_zkey() {  // also runs when loading fails half-way (a rejected point): nothing stays allocated on the device
    if (zk.owner) {
      for (void* p : zk.owned) cocg_free(zk.owner, p);
      cocg_destroy(zk.owner);
    }
  }
};

namespace {
void* plonk_dev_alloc(PlonkZKey& zk, size_t bytes) {
  void* p = nullptr;
  check(zk.owner, cocg_malloc(zk.owner, bytes, &p), "cocg_malloc");
  zk.owned.push_back(p);
  return p;
}
void* plonk_dev_upload(PlonkZKey& zk, const void* host, size_t bytes) {
  void* p = plonk_dev_alloc(zk, bytes);
  check(zk.owner, cocg_h2d(zk.owner, p, host, bytes), "cocg_h2d");
  return p;
}
void plonk_upload_maps(PlonkZKey& zk) {
  const std::vector<uint32_t>* maps[3] = {&zk.map_a, &zk.map_b, &zk.map_c};
  for (int k = 0; k < 3; k++) {
    std::vector<uint32_t> padded(zk.domain_size, 0xffffffffu);  // gates past n_constraints read zero (round1.rs:150-166)
    std::copy(maps[k]->begin(), maps[k]->end(), padded.begin());
    zk.d_map[k] = (uint32_t*)plonk_dev_upload(zk, padded.data(), padded.size() * 4);
  }
}
void plonk_proof_pack(const PlonkProof& p, size_t lq, uint64_t* o) {
  const Point* pts[9] = {&p.a, &p.b, &p.c, &p.z, &p.t1, &p.t2, &p.t3, &p.wxi, &p.wxiw};
  for (int i = 0; i < 9; i++) memcpy(o + (size_t)i * 2 * lq, pts[i]->l, 2 * lq * 8);
  const Fr* ev[6] = {&p.eval_a, &p.eval_b, &p.eval_c, &p.eval_s1, &p.eval_s2, &p.eval_zw};
  for (int i = 0; i < 6; i++) memcpy(o + 18 * lq + (size_t)i * 4, ev[i]->l, 32);
}
}  // namespace

extern "C" int cohost_plonk_zkey_load_file(const char* path, int device, cohost_plonk_zkey** out) {
  if (!path || !out) return fail("cohost_plonk_zkey_load_file: null argument");
  *out = nullptr;
  return guarded([&] {
    std::vector<uint8_t> buf = read_file(path);
    PlonkZKeyFile f(buf.data(), buf.size());
    std::unique_ptr<cohost_plonk_zkey> z(new cohost_plonk_zkey());
    PlonkZKey& zk = z->zk;
    zk.curve = f.curve;
    zk.device = device;
    z->lq = f.curve == COCG_BN254 ? 4 : 6;
    zk.n_vars = f.n_vars; zk.n_public = f.n_public; zk.domain_size = f.domain_size; zk.pow = f.pow;
    zk.n_additions = f.n_additions; zk.n_constraints = f.n_constraints;
    zk.additions.resize(f.n_additions);
    for (size_t i = 0; i < f.n_additions; i++) {
      const uint8_t* e = f.additions + i * 72;
      zk.additions[i].s1 = BinFile::u32(e);
      zk.additions[i].s2 = BinFile::u32(e + 4);
      memcpy(zk.additions[i].f1.l, e + 8, 32);   // already Montgomery limbs (montgomery_bigint_from_reader)
      memcpy(zk.additions[i].f2.l, e + 40, 32);
    }
    auto rd_map = [&](const uint8_t* p, std::vector<uint32_t>& m) {
      m.resize(f.n_constraints);
      memcpy(m.data(), p, f.n_constraints * 4);
    };
    rd_map(f.map_a, zk.map_a);
    rd_map(f.map_b, zk.map_b);
    rd_map(f.map_c, zk.map_c);
    if (cocg_create(&zk.owner, device, f.curve)) throw Error(std::string("cocg_create: ") + cocg_last_error(nullptr));
    check(zk.owner, cocg_bases_upload(zk.owner, COCG_G1, f.p_tau, f.domain_size + 6, 2 * f.n8q, 1, &zk.p_tau), "p_tau");
    {  // taus() -> g1_vec_from_reader checks every point (plonk/zkey.rs:222-225, traits.rs:107-126)
      size_t bad = 0, first = 0;
      check(zk.owner, cocg_bases_check(zk.owner, zk.p_tau, 1, &bad, &first), "cocg_bases_check");
      if (bad) throw Error("zkey: InvalidData: point " + std::to_string(first) + " of the p_tau section is not on the curve or not in the prime-order subgroup");
    }
    plonk_upload_maps(zk);
    // rounds 2-5: selector / sigma / Lagrange polynomials (n coefficients | 4n evaluations each, Montgomery) and the vk tail
    if (f.k1 && f.sigma && f.lagrange && f.sel[0] && f.sel[1] && f.sel[2] && f.sel[3] && f.sel[4]) {
      const size_t n = f.domain_size;
      memcpy(zk.k1.l, f.k1, 32);
      memcpy(zk.k2.l, f.k2, 32);
      for (int i = 0; i < 8; i++) memcpy(zk.vk_g1[i].l, f.vk_g1 + (size_t)i * 2 * f.n8q, 2 * f.n8q);
      for (int k = 0; k < 5; k++) {
        char* d = (char*)plonk_dev_upload(zk, f.sel[k], 5 * n * 32);
        zk.sel_coef[k] = d;
        zk.sel_eval[k] = d + n * 32;
      }
      for (int k = 0; k < 3; k++) {
        char* d = (char*)plonk_dev_upload(zk, f.sigma + (size_t)k * 5 * n * 32, 5 * n * 32);
        zk.sig_coef[k] = d;
        zk.sig_eval[k] = d + n * 32;
      }
      zk.n_lagrange = f.n_public > 1 ? f.n_public : 1;
      char* lg = (char*)plonk_dev_alloc(zk, zk.n_lagrange * 4 * n * 32);
      for (size_t j = 0; j < zk.n_lagrange; j++)
        check(zk.owner, cocg_h2d(zk.owner, lg + j * 4 * n * 32, f.lagrange + j * 5 * n * 32 + n * 32, 4 * n * 32), "lagrange");
      zk.lagrange = lg;
      zk.full = true;
    }
    *out = z.release();
  });
}
// A shape-faithful synthetic key (BASELINE configs[3]: 2^18 gates; no circom / snarkjs exists here to make a real one): the wire maps
// come from the caller, every polynomial is filled by the device PRF, p_tau and the vk points are valid curve points generated in HBM.
// The resulting proofs exercise every kernel of the five rounds at full size; they are not expected to verify.
extern "C" int cohost_plonk_zkey_create_synthetic(int curve, int device, size_t log_n, size_t n_public, size_t n_vars, size_t n_constraints,
                                                  const uint32_t* map_a, const uint32_t* map_b, const uint32_t* map_c, const uint8_t* seed,
                                                  cohost_plonk_zkey** out) {
  if (!out || !seed || !map_a || !map_b || !map_c) return fail("cohost_plonk_zkey_create_synthetic: null argument");
  *out = nullptr;
  return guarded([&] {
    const size_t n = (size_t)1 << log_n;
    if (n_constraints > n || n_vars < n_public + 1) throw Error("plonk synthetic key: inconsistent sizes");
    std::unique_ptr<cohost_plonk_zkey> z(new cohost_plonk_zkey());
    PlonkZKey& zk = z->zk;
    zk.curve = curve;
    zk.device = device;
    z->lq = curve == COCG_BN254 ? 4 : 6;
    zk.n_vars = n_vars; zk.n_public = n_public; zk.domain_size = n; zk.pow = log_n; zk.n_additions = 0; zk.n_constraints = n_constraints;
    for (const uint32_t* m : {map_a, map_b, map_c})
      for (size_t i = 0; i < n_constraints; i++)
        if (m[i] >= n_vars) throw Error("plonk synthetic key: wire map index out of range");
    zk.map_a.assign(map_a, map_a + n_constraints);
    zk.map_b.assign(map_b, map_b + n_constraints);
    zk.map_c.assign(map_c, map_c + n_constraints);
    if (cocg_create(&zk.owner, device, curve)) throw Error(std::string("cocg_create: ") + cocg_last_error(nullptr));
    uint8_t sd[32];
    memcpy(sd, seed, 32);
    check(zk.owner, cocg_bases_generate(zk.owner, COCG_G1, n + 6, sd, &zk.p_tau), "p_tau");
    plonk_upload_maps(zk);
    uint32_t ctr = 1;
    auto filled = [&](size_t elems) {
      void* p = plonk_dev_alloc(zk, elems * 32);
      check(zk.owner, cocg_prf_fill(zk.owner, sd, ctr++, p, elems), "cocg_prf_fill");
      return p;
    };
    for (int k = 0; k < 5; k++) { zk.sel_coef[k] = filled(n); zk.sel_eval[k] = filled(4 * n); }
    for (int k = 0; k < 3; k++) { zk.sig_coef[k] = filled(n); zk.sig_eval[k] = filled(4 * n); }
    zk.n_lagrange = n_public > 1 ? n_public : 1;
    zk.lagrange = filled(zk.n_lagrange * 4 * n);
    cocg_prf_field_host(curve, sd, 1000, 0, zk.k1.l);
    cocg_prf_field_host(curve, sd, 1000, 1, zk.k2.l);
    {
      uint64_t h = 0;
      sd[31] ^= 0x5a;
      check(zk.owner, cocg_bases_generate(zk.owner, COCG_G1, 8, sd, &h), "vk points");
      std::vector<uint64_t> b(8 * 2 * z->lq);
      check(zk.owner, cocg_bases_download(zk.owner, h, 0, 8, b.data()), "vk points");
      cocg_bases_free(zk.owner, h);
      for (int i = 0; i < 8; i++) memcpy(zk.vk_g1[i].l, b.data() + (size_t)i * 2 * z->lq, 2 * z->lq * 8);
    }
    check(zk.owner, cocg_sync(zk.owner), "cocg_sync");
    zk.full = true;
    *out = z.release();
  });
}
extern "C" void cohost_plonk_zkey_destroy(cohost_plonk_zkey* z) { delete z; }
// info[6] = curve, n_vars, n_public, domain_size, n_additions, n_constraints
extern "C" int cohost_plonk_zkey_get_info(cohost_plonk_zkey* z, size_t* info) {
  if (!z || !info) return fail("cohost_plonk_zkey_get_info: null argument");
  const PlonkZKey& k = z->zk;
  const size_t v[6] = {(size_t)k.curve, k.n_vars, k.n_public, k.domain_size, k.n_additions, k.n_constraints};
  memcpy(info, v, sizeof(v));
  return 0;
}

// CoPlonk sessions: protocol 0 = CoPlonk<PlainDriver> (one party), 1 = three CoPlonk<Rep3Protocol> provers on three threads over the
// in-process network (tests/tests/circom/e2e_tests/mod.rs:55-70 runs the reference's Plonk provers the same way).
struct cohost_plonk_session {
  cohost_plonk_zkey* zkey = nullptr;
  int parties = 1, protocol = 0;
  std::unique_ptr<PlainDriver> plain;
  std::unique_ptr<Rep3TestNetwork> net;
  std::unique_ptr<Rep3Protocol> drv[3];
  std::unique_ptr<ShamirTestNetwork> snet;  // protocol 2: `parties` CoPlonk<ShamirProtocol> provers
  std::vector<std::unique_ptr<ShamirProtocol>> sdrv;
  std::vector<uint64_t> h_tau;
  bool trace_on = false;
  std::vector<std::map<std::string, std::vector<Fr>>> traces;
  std::vector<std::array<double, 5>> round_s;
  bool failed = false;
  DeviceDriver* driver(int i) { return protocol == 0 ? (DeviceDriver*)plain.get() : protocol == 1 ? (DeviceDriver*)drv[i].get() : (DeviceDriver*)sdrv[i].get(); }
  void size_for(int n) {
    parties = n;
    h_tau.assign(n, 0);
    traces.assign(n, {});
    round_s.assign(n, std::array<double, 5>{});
  }
  void close_all() {
    if (net) net->close_all();
    if (snet) snet->close_all();
  }
};

extern "C" int cohost_plonk_session_create(cohost_plonk_zkey* z, int protocol, const uint8_t* seeds, cohost_plonk_session** out) {
  if (!z || !out || !seeds) return fail("cohost_plonk_session_create: null argument");
  if (protocol != 0 && protocol != 1) return fail("cohost_plonk_session_create: protocol must be 0 (plain) or 1 (REP3); Shamir sessions come from cohost_plonk_session_create_shamir");
  *out = nullptr;
  return guarded([&] {
    std::unique_ptr<cohost_plonk_session> s(new cohost_plonk_session());
    s->zkey = z;
    s->protocol = protocol;
    s->size_for(protocol == 0 ? 1 : 3);
    if (protocol == 0) {
      s->plain.reset(new PlainDriver(z->zk.curve, z->zk.device));
      memcpy(s->plain->seed, seeds, 32);
      check(s->plain->ctx, cocg_bases_share(s->plain->ctx, z->zk.owner, z->zk.p_tau, &s->h_tau[0]), "share p_tau");
    } else {
      s->net.reset(new Rep3TestNetwork());
      const char* ex = getenv("COHOST_MPC_EXCHANGE");
      s->net->device_exchange = ex && std::string(ex) == "device";
      for (int i = 0; i < 3; i++) s->drv[i].reset(new Rep3Protocol(z->zk.curve, z->zk.device, s->net->party(i), seeds + 32 * i));
      for (int i = 0; i < 3; i++) {
        s->drv[i]->finish_setup();
        check(s->drv[i]->ctx, cocg_bases_share(s->drv[i]->ctx, z->zk.owner, z->zk.p_tau, &s->h_tau[i]), "share p_tau");
      }
    }
    *out = s.release();
  });
}
// CoPlonk over Shamir (num_parties, threshold) shares: ShamirProtocol implements the same driver surface (co-plonk is generic over
// PrimeFieldMpcProtocol + ..., plonk.rs:50-77; mpc-core/src/protocols/shamir.rs:459-712).  seeds: num_parties x 32 bytes.
extern "C" int cohost_plonk_session_create_shamir(cohost_plonk_zkey* z, int num_parties, int threshold, const uint8_t* seeds, cohost_plonk_session** out) {
  if (!z || !out || !seeds) return fail("cohost_plonk_session_create_shamir: null argument");
  if (num_parties < 3 || num_parties > 64) return fail("Shamir protocol requires at least 3 parties");  // shamir/network.rs:75-77
  *out = nullptr;
  return guarded([&] {
    std::unique_ptr<cohost_plonk_session> s(new cohost_plonk_session());
    s->zkey = z;
    s->protocol = 2;
    s->size_for(num_parties);
    s->snet.reset(new ShamirTestNetwork(num_parties));
    const char* ex = getenv("COHOST_MPC_EXCHANGE");
    s->snet->device_exchange = ex && std::string(ex) == "device";
    for (int i = 0; i < num_parties; i++) {
      s->sdrv.emplace_back(new ShamirProtocol(z->zk.curve, z->zk.device, threshold, s->snet->party(i), seeds + 32 * i));
      check(s->sdrv[i]->ctx, cocg_bases_share(s->sdrv[i]->ctx, z->zk.owner, z->zk.p_tau, &s->h_tau[i]), "share p_tau");
    }
    *out = s.release();
  });
}
extern "C" void cohost_plonk_session_destroy(cohost_plonk_session* s) { delete s; }
extern "C" int cohost_plonk_session_parties(cohost_plonk_session* s) { return s ? s->parties : 0; }

// One proof.  public_inputs: n_public + 1 Fr; wit_a / wit_b: `parties` pointers each (wit_b may be NULL for the one-component plain and
// Shamir drivers) to the parties' share components of the private witness, HOST memory or -- wit_on_device -- HBM of the session's device.
// proofs_out: parties x cohost_plonk_proof_limbs() u64.  rounds = 1 stops after round 1 (commitments a, b, c only).
template <class T, class GetDriver>
static void plonk_prove_parties(cohost_plonk_session* s, GetDriver get, const void* public_inputs, const void* const* wit_a, const void* const* wit_b,
                                bool deterministic, bool wit_on_device, bool round1_only, uint64_t* proofs_out) {
  const PlonkZKey& zk = s->zkey->zk;
  const size_t lq = s->zkey->lq, pl = 18 * lq + 24;
  const int n = s->parties;
  std::vector<std::thread> th(n);
  std::vector<std::string> errs(n);
  std::vector<PlonkProof> proofs(n);
  for (int i = 0; i < n; i++) {
    const void* a = wit_a[i];
    const void* b = wit_b ? wit_b[i] : nullptr;
    th[i] = std::thread([&, i, a, b] {
      try {
        CoPlonk<T> pv(*get(i));
        if (s->trace_on) { s->traces[i].clear(); pv.trace = &s->traces[i]; }
        proofs[i] = pv.prove(zk, s->h_tau[i], (const Fr*)public_inputs, a, b, deterministic, wit_on_device, round1_only);
        memcpy(s->round_s[i].data(), pv.round_s, sizeof(pv.round_s));
      } catch (const std::exception& e) {
        errs[i] = e.what();
        s->close_all();
      }
    });
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < n; i++)
    if (!errs[i].empty()) {
      s->failed = true;
      throw Error("party " + std::to_string(i) + ": " + errs[i]);
    }
  for (int i = 0; i < n; i++) plonk_proof_pack(proofs[i], lq, proofs_out + (size_t)i * pl);
}
static int plonk_prove_impl(cohost_plonk_session* s, const void* public_inputs, const void* const* wit_a, const void* const* wit_b, int deterministic,
                            int wit_on_device, int rounds, void* proofs_out) {
  if (!s || !public_inputs || !wit_a || !proofs_out) return fail("cohost_plonk_prove: null argument");
  if (s->protocol == 1 && !wit_b) return fail("cohost_plonk_prove: REP3 needs both share components");
  if (s->failed) return fail("cohost_plonk_prove: the session's network is closed after an earlier failure");
  return guarded([&] {
    const PlonkZKey& zk = s->zkey->zk;
    const size_t lq = s->zkey->lq;
    if (s->protocol == 0) {
      CoPlonk<PlainDriver> pv(*s->plain);
      if (s->trace_on) { s->traces[0].clear(); pv.trace = &s->traces[0]; }
      PlonkProof p = pv.prove(zk, s->h_tau[0], (const Fr*)public_inputs, wit_a[0], nullptr, deterministic != 0, wit_on_device != 0, rounds == 1);
      memcpy(s->round_s[0].data(), pv.round_s, sizeof(pv.round_s));
      plonk_proof_pack(p, lq, (uint64_t*)proofs_out);
    } else if (s->protocol == 1) {
      plonk_prove_parties<Rep3Protocol>(s, [s](int i) { return s->drv[i].get(); }, public_inputs, wit_a, wit_b, deterministic != 0, wit_on_device != 0, rounds == 1,
                                        (uint64_t*)proofs_out);
    } else {
      plonk_prove_parties<ShamirProtocol>(s, [s](int i) { return s->sdrv[i].get(); }, public_inputs, wit_a, nullptr, deterministic != 0, wit_on_device != 0,
                                          rounds == 1, (uint64_t*)proofs_out);
    }
  });
}
extern "C" int cohost_plonk_prove(cohost_plonk_session* s, const void* public_inputs, const void* const* wit_a, const void* const* wit_b, int deterministic,
                                  int wit_on_device, void* proofs_out) {
  return plonk_prove_impl(s, public_inputs, wit_a, wit_b, deterministic, wit_on_device, 5, proofs_out);
}
extern "C" size_t cohost_plonk_proof_limbs(cohost_plonk_zkey* z) { return z ? 18 * z->lq + 24 : 0; }
extern "C" int cohost_plonk_set_mpc_exchange(cohost_plonk_session* s, int device) {
  if (!s) return fail("cohost_plonk_set_mpc_exchange: null session");
  if (s->net) s->net->device_exchange = device != 0;
  if (s->snet) s->snet->device_exchange = device != 0;
  return 0;
}
extern "C" uint64_t cohost_plonk_launch_count(cohost_plonk_session* s) {
  uint64_t t = 0;
  if (s)
    for (int i = 0; i < s->parties; i++) t += cocg_launch_count(s->driver(i)->ctx);
  return t;
}
extern "C" int cohost_plonk_profile_enable(cohost_plonk_session* s, int on) {
  if (!s) return fail("null session");
  for (int i = 0; i < s->parties; i++) cocg_profile_enable(s->driver(i)->ctx, on);
  return 0;
}
extern "C" int cohost_plonk_profile_reset(cohost_plonk_session* s) {
  if (!s) return fail("null session");
  for (int i = 0; i < s->parties; i++) cocg_profile_reset(s->driver(i)->ctx);
  return 0;
}
extern "C" int cohost_plonk_profile_read(cohost_plonk_session* s, int cls, double* total_ms, uint64_t* scopes) {
  if (!s) return fail("null session");
  double t = 0;
  uint64_t n = 0;
  for (int i = 0; i < s->parties; i++) {
    double ms = 0;
    uint64_t k = 0;
    if (cocg_profile_read(s->driver(i)->ctx, cls, &ms, &k)) return fail(cocg_last_error(s->driver(i)->ctx));
    t += ms;
    n += k;
  }
  if (total_ms) *total_ms = t;
  if (scopes) *scopes = n;
  return 0;
}
// Host wall-clock per round of the last proof, seconds: out[party * 5 + round].
extern "C" int cohost_plonk_round_times(cohost_plonk_session* s, double* out) {
  if (!s || !out) return fail("cohost_plonk_round_times: null argument");
  for (int i = 0; i < s->parties; i++) memcpy(out + 5 * i, s->round_s[i].data(), 5 * sizeof(double));
  return 0;
}
// Test hook: keep component a of named intermediate vectors of the next proofs (buffer_z, poly_z, t_evals, tz_evals, t1, t2, t3, poly_r,
// wxi, ...); the sum over the three parties is the plain value.  get: out == NULL queries the length.
extern "C" int cohost_plonk_trace_enable(cohost_plonk_session* s, int on) {
  if (!s) return fail("null session");
  s->trace_on = on != 0;
  return 0;
}
extern "C" int cohost_plonk_trace_get(cohost_plonk_session* s, int party, const char* name, void* out, size_t cap, size_t* n) {
  if (!s || !name || !n || party < 0 || party >= s->parties) return fail("cohost_plonk_trace_get: bad argument");
  auto it = s->traces[party].find(name);
  if (it == s->traces[party].end()) return fail(std::string("cohost_plonk_trace_get: no vector named ") + name);
  *n = it->second.size();
  if (out) {
    if (cap < it->second.size()) return fail("cohost_plonk_trace_get: buffer too small");
    memcpy(out, it->second.data(), it->second.size() * 32);
  }
  return 0;
}

// Round 1 only (kept from the first build of this path): commitments [a]_1 | [b]_1 | [c]_1, packed affine.
extern "C" int cohost_plonk_round1_plain(cohost_plonk_zkey* z, const void* public_inputs, const void* witness, int deterministic, void* commits_out) {
  if (!z || !public_inputs || !witness || !commits_out) return fail("cohost_plonk_round1_plain: null argument");
  cohost_plonk_session* s = nullptr;
  uint8_t seed[32] = {};
  if (cohost_plonk_session_create(z, 0, seed, &s)) return 1;
  std::vector<uint64_t> pr(18 * z->lq + 24);
  const void* w[1] = {witness};
  int rc = plonk_prove_impl(s, public_inputs, w, nullptr, deterministic, 0, 1, pr.data());
  cohost_plonk_session_destroy(s);
  if (rc) return rc;
  memcpy(commits_out, pr.data(), 6 * z->lq * 8);
  return 0;
}
extern "C" int cohost_plonk_round1_rep3(cohost_plonk_zkey* z, const void* public_inputs, const void* const* wit_a, const void* const* wit_b,
                                        const uint8_t* seeds, int deterministic, void* commits_out) {
  if (!z || !public_inputs || !wit_a || !wit_b || !seeds || !commits_out) return fail("cohost_plonk_round1_rep3: null argument");
  cohost_plonk_session* s = nullptr;
  if (cohost_plonk_session_create(z, 1, seeds, &s)) return 1;
  const size_t pl = 18 * z->lq + 24;
  std::vector<uint64_t> pr(3 * pl);
  int rc = plonk_prove_impl(s, public_inputs, wit_a, wit_b, deterministic, 0, 1, pr.data());
  cohost_plonk_session_destroy(s);
  if (rc) return rc;
  for (int i = 0; i < 3; i++) memcpy((uint64_t*)commits_out + (size_t)i * 6 * z->lq, pr.data() + (size_t)i * pl, 6 * z->lq * 8);
  return 0;
}

// ------------------------------------------------------------------------------------------------ batched witness-extension arithmetic
struct cohost_vm {
  int parties = 1;
  size_t batch = 0;
  std::unique_ptr<PlainDriver> plain;
  std::unique_ptr<Rep3TestNetwork> net;
  std::unique_ptr<Rep3Protocol> drv[3];
  std::unique_ptr<BatchedWitnessVm<PlainDriver>> vm_plain;
  std::unique_ptr<BatchedWitnessVm<Rep3Protocol>> vm_rep3[3];
  bool failed = false;
};
extern "C" int cohost_vm_create(int curve, int device, int protocol, const uint8_t* seeds, size_t batch, int n_regs, cohost_vm** out) {
  if (!out || !seeds) return fail("cohost_vm_create: null argument");
  if (protocol != 0 && protocol != 1) return fail("cohost_vm_create: protocol must be 0 (plain) or 1 (REP3)");
  if (n_regs <= 0 || n_regs > 65536) return fail("cohost_vm_create: bad register count");
  *out = nullptr;
  return guarded([&] {
    std::unique_ptr<cohost_vm> v(new cohost_vm());
    v->parties = protocol == 0 ? 1 : 3;
    v->batch = batch;
    if (protocol == 0) {
      v->plain.reset(new PlainDriver(curve, device));
      memcpy(v->plain->seed, seeds, 32);
      v->vm_plain.reset(new BatchedWitnessVm<PlainDriver>(*v->plain, batch, n_regs));
    } else {
      v->net.reset(new Rep3TestNetwork());
      v->net->device_exchange = true;  // the three parties of this process share the GPU
      for (int i = 0; i < 3; i++) v->drv[i].reset(new Rep3Protocol(curve, device, v->net->party(i), seeds + 32 * i));
      for (int i = 0; i < 3; i++) {
        v->drv[i]->finish_setup();
        v->vm_rep3[i].reset(new BatchedWitnessVm<Rep3Protocol>(*v->drv[i], batch, n_regs));
      }
    }
    *out = v.release();
  });
}
extern "C" void cohost_vm_destroy(cohost_vm* v) {
  if (!v) return;
  v->vm_plain.reset();
  for (auto& m : v->vm_rep3) m.reset();  // registers go back to the drivers' pools before the drivers die
  delete v;
}
extern "C" int cohost_vm_set_public(cohost_vm* v, int reg, const void* values) {
  if (!v || !values) return fail("cohost_vm_set_public: null argument");
  return guarded([&] {
    if (v->parties == 1) v->vm_plain->set_public(reg, values);
    else for (int i = 0; i < 3; i++) v->vm_rep3[i]->set_public(reg, values);
  });
}
extern "C" int cohost_vm_set_shared(cohost_vm* v, int reg, int party, const void* a, const void* b) {
  if (!v || !a || party < 0 || party >= v->parties || (v->parties == 3 && !b)) return fail("cohost_vm_set_shared: bad argument");
  return guarded([&] {
    if (v->parties == 1) v->vm_plain->set_shared(reg, a, nullptr);
    else v->vm_rep3[party]->set_shared(reg, a, b);
  });
}
extern "C" int cohost_vm_run(cohost_vm* v, const cohost_vm_instr* prog, size_t n) {
  if (!v || (n && !prog)) return fail("cohost_vm_run: null argument");
  if (v->failed) return fail("cohost_vm_run: the network is closed after an earlier failure");
  static_assert(sizeof(cohost_vm_instr) == sizeof(VmInstr), "cohost_vm_instr must mirror VmInstr");
  const VmInstr* p = reinterpret_cast<const VmInstr*>(prog);
  return guarded([&] {
    if (v->parties == 1) { v->vm_plain->run(p, n); return; }
    std::thread th[3];
    std::string errs[3];
    for (int i = 0; i < 3; i++)
      th[i] = std::thread([&, i] {
        try {
          v->vm_rep3[i]->run(p, n);
        } catch (const std::exception& e) {
          errs[i] = e.what();
          v->net->close_all();
        }
      });
    for (auto& t : th) t.join();
    for (int i = 0; i < 3; i++)
      if (!errs[i].empty()) {
        v->failed = true;
        throw Error("party " + std::to_string(i) + ": " + errs[i]);
      }
  });
}
// kind: 1 public (a_out receives the values, b_out untouched), 2 shared (a_out | b_out: the party's components; b_out unused by plain)
extern "C" int cohost_vm_get(cohost_vm* v, int reg, int party, int* kind, void* a_out, void* b_out) {
  if (!v || !kind || party < 0 || party >= v->parties) return fail("cohost_vm_get: bad argument");
  return guarded([&] {
    auto get = [&](auto& vm) {
      VmValue& val = vm.at(reg);
      *kind = (int)val.kind;
      if (val.kind == VmValue::PUBLIC && a_out) vm.driver.download(val.pub, a_out);
      if (val.kind == VmValue::SHARED) {
        if (a_out) vm.driver.download(val.sh.a, a_out);
        if (b_out && val.sh.b.p) vm.driver.download(val.sh.b, b_out);
      }
    };
    if (v->parties == 1) get(*v->vm_plain);
    else get(*v->vm_rep3[party]);
  });
}
// out[0] = kernel launches, out[1] = network rounds of party 0 (shared multiplications and inversions: ONE exchange each per batch)
extern "C" int cohost_vm_stats(cohost_vm* v, uint64_t* out) {
  if (!v || !out) return fail("cohost_vm_stats: null argument");
  out[0] = 0;
  if (v->parties == 1) { out[0] = cocg_launch_count(v->plain->ctx); out[1] = v->vm_plain->network_rounds; }
  else { for (int i = 0; i < 3; i++) out[0] += cocg_launch_count(v->drv[i]->ctx); out[1] = v->vm_rep3[0]->network_rounds; }
  return 0;
}

// How the three co-located parties of a session move share vectors in the mul_vec rounds: 0 = staged through pinned host memory
// (default; what a party that has to reach a NIC does), 1 = handed over in HBM (all parties share the session's GPU).
extern "C" int cohost_rep3_set_mpc_exchange(cohost_rep3_session* s, int device) {
  if (!s) return fail("cohost_rep3_set_mpc_exchange: null session");
  if (s->running) return fail("cohost_rep3_set_mpc_exchange: a proof is in flight");
  s->net->device_exchange = device != 0;
  return 0;
}

// ------------------------------------------------------------------------------------------------ output formats (serialize.hpp)
namespace {
int copy_out(const void* src, size_t n, void* out, size_t cap, size_t* len, const char* what) {
  if (len) *len = n;
  if (!out) return 0;  // size query
  if (cap < n) return fail(std::string(what) + ": output buffer too small");
  memcpy(out, src, n);
  return 0;
}
}  // namespace

// proof: A | B | C packed affine Montgomery (as written by cohost_*_prove).  out == NULL: only *len is set.  No terminating NUL.
extern "C" int cohost_proof_to_json(int curve, const void* proof, char* out, size_t cap, size_t* len) {
  if (!proof) return fail("cohost_proof_to_json: null argument");
  if (curve != COCG_BN254 && curve != COCG_BLS12_381) return fail("cohost_proof_to_json: unknown curve");
  int rc = 0;
  int g = guarded([&] {
    std::string s = proof_to_json(curve, (const uint64_t*)proof);
    rc = copy_out(s.data(), s.size(), out, cap, len, "cohost_proof_to_json");
  });
  return g ? g : rc;
}
// pub: count Montgomery Fr, pub[0] = the constant 1 (skipped in the output like the reference's writer)
extern "C" int cohost_plonk_proof_to_json(int curve, const void* proof, char* out, size_t cap, size_t* len) {
  if (!proof || !len) return fail("cohost_plonk_proof_to_json: null argument");
  if (curve != COCG_BN254 && curve != COCG_BLS12_381) return fail("cohost_plonk_proof_to_json: unknown curve");
  return guarded([&] {
    std::string s = plonk_proof_to_json(curve, (const uint64_t*)proof);
    *len = s.size();
    if (!out) return;
    if (cap < s.size()) throw Error("cohost_plonk_proof_to_json: buffer too small");
    memcpy(out, s.data(), s.size());
  });
}
extern "C" int cohost_public_inputs_to_json(int curve, const void* pub, size_t count, char* out, size_t cap, size_t* len) {
  if (!pub && count) return fail("cohost_public_inputs_to_json: null argument");
  if (curve != COCG_BN254 && curve != COCG_BLS12_381) return fail("cohost_public_inputs_to_json: unknown curve");
  int rc = 0;
  int g = guarded([&] {
    std::string s = public_inputs_to_json(curve, (const uint64_t*)pub, count);
    rc = copy_out(s.data(), s.size(), out, cap, len, "cohost_public_inputs_to_json");
  });
  return g ? g : rc;
}
// SharedWitness file image of one party.  comps: k = 2 (REP3: a, b) or 1 (Shamir) HOST vectors of n Montgomery Fr.
extern "C" int cohost_shared_witness_encode(int curve, const void* pub, size_t n_pub, const void* const* comps, int k, size_t n, void* out,
                                            size_t cap, size_t* len) {
  if ((!pub && n_pub) || !comps || (k != 1 && k != 2)) return fail("cohost_shared_witness_encode: bad argument");
  if (curve != COCG_BN254 && curve != COCG_BLS12_381) return fail("cohost_shared_witness_encode: unknown curve");
  const size_t need = 8 + 8 + 32 * n_pub + 8 + (size_t)k * (8 + 32 * n);
  if (len) *len = need;
  if (!out) return 0;
  if (cap < need) return fail("cohost_shared_witness_encode: output buffer too small");
  for (int j = 0; j < k; j++)
    if (n && !comps[j]) return fail("cohost_shared_witness_encode: null share component");
  return guarded([&] {
    const uint64_t* c[2] = {(const uint64_t*)comps[0], k == 2 ? (const uint64_t*)comps[1] : nullptr};
    std::vector<uint8_t> o = shared_witness_encode(curve, (const uint64_t*)pub, n_pub, c, k, n);
    memcpy(out, o.data(), o.size());
  });
}
// Two-step decode: with pub == NULL only the element counts are returned; then the caller passes buffers of that size.
extern "C" int cohost_shared_witness_decode(int curve, const void* data, size_t len, int k, size_t* n_pub, size_t* n, void* pub,
                                            void* const* comps) {
  if (!data || (k != 1 && k != 2) || !n_pub || !n) return fail("cohost_shared_witness_decode: bad argument");
  if (curve != COCG_BN254 && curve != COCG_BLS12_381) return fail("cohost_shared_witness_decode: unknown curve");
  return guarded([&] {
    SharedWitnessData w = shared_witness_decode(curve, (const uint8_t*)data, len, k);
    *n_pub = w.public_inputs.size();
    *n = w.comps[0].size();
    if (!pub && !comps) return;
    if (pub && *n_pub) memcpy(pub, w.public_inputs.data(), *n_pub * 32);
    if (comps)
      for (int j = 0; j < k; j++)
        if (comps[j] && *n) memcpy(comps[j], w.comps[j].data(), *n * 32);
  });
}
// SharedWitness::share_rep3 (co-circom-snarks/src/lib.rs:149-173) on the GPU: a, b <- PRF streams of seed[0..32) / seed[32..64),
// c = x - a - b; party 0 = (a, c), party 1 = (b, a), party 2 = (c, b)  (rep3::utils::share_field_elements, mpc-core/src/protocols/rep3.rs:124-150;
// the reference draws a and b from the caller's RNG, so the shares themselves are not reproducible across implementations).
// witness: n Montgomery Fr on the HOST (values[num_pub_inputs..]); out_a / out_b: 3 HOST vectors each.
extern "C" int cohost_split_witness_rep3(int curve, int device, const void* witness, size_t n, const uint8_t* seed, void* const* out_a,
                                         void* const* out_b) {
  if ((!witness && n) || !seed || !out_a || !out_b) return fail("cohost_split_witness_rep3: null argument");
  return guarded([&] {
    DeviceDriver d(curve, device);
    cocg_ctx* c = d.ctx;
    if (n == 0) return;
    DevVec x = d.alloc(n), a = d.alloc(n), b = d.alloc(n);
    check(c, cocg_h2d(c, x.p, witness, n * 32), "cocg_h2d");
    check(c, cocg_prf_fill(c, seed, 0, a.p, n), "cocg_prf_fill");
    check(c, cocg_prf_fill(c, seed + 32, 0, b.p, n), "cocg_prf_fill");
    check(c, cocg_vec_op(c, COCG_OP_SUB, x.p, a.p, x.p, n), "cocg_vec_op");
    check(c, cocg_vec_op(c, COCG_OP_SUB, x.p, b.p, x.p, n), "cocg_vec_op");  // x is now c
    const void* src_a[3] = {a.p, b.p, x.p};
    const void* src_b[3] = {x.p, a.p, b.p};
    for (int i = 0; i < 3; i++) {
      check(c, cocg_d2h(c, out_a[i], src_a[i], n * 32), "cocg_d2h");
      check(c, cocg_d2h(c, out_b[i], src_b[i], n * 32), "cocg_d2h");
    }
    d.release(x); d.release(a); d.release(b);
  });
}

// info[6] = curve, n_wires, n_pub_out, n_pub_in, n_constraints, num_inputs (= 1 + n_pub_out + n_pub_in, r1cs.rs:201).  No GPU needed.
extern "C" int cohost_r1cs_info(const char* path, size_t* info) {
  if (!path || !info) return fail("cohost_r1cs_info: null argument");
  return guarded([&] {
    std::vector<uint8_t> buf = read_file(path);
    R1csHeader h(buf.data(), buf.size());
    info[0] = (size_t)h.curve; info[1] = h.n_wires; info[2] = h.n_pub_out; info[3] = h.n_pub_in; info[4] = h.n_constraints; info[5] = h.num_inputs();
  });
}

// `co-circom split-witness` (co-circom/src/bin/co-circom.rs:160-256): witness.wtns + circuit.r1cs -> <out_dir>/<witness file name>.<i>.shared,
// one SharedWitness file per party.  protocol 0 = REP3 (threshold 1, 3 parties), 1 = Shamir(threshold, num_parties): share i is the value of
// x + sum_k c_k (i + 1)^k, k = 1..threshold (shamir/shamir_core.rs:8-33).  The random shares / coefficients come from the GPU PRF keyed by
// seed[32 * j ..) (REP3: 2 streams, Shamir: `threshold` streams; seed holds 32 * max(2, threshold) bytes).
extern "C" int cohost_split_witness_files(const char* witness_path, const char* r1cs_path, int protocol, int curve, int threshold, int num_parties,
                                          const uint8_t* seed, const char* out_dir, int device) {
  if (!witness_path || !r1cs_path || !seed || !out_dir) return fail("cohost_split_witness_files: null argument");
  if (protocol == 0 && threshold != 1) return fail("REP3 only allows the threshold to be 1");
  if (protocol == 0 && num_parties != 3) return fail("REP3 only allows the number of parties to be 3");
  if (protocol != 0 && protocol != 1) return fail("cohost_split_witness_files: unknown protocol");
  if (protocol == 1 && (threshold < 1 || num_parties <= threshold)) return fail("Shamir needs 1 <= threshold < num_parties");
  return guarded([&] {
    std::vector<uint8_t> wbuf = read_file(witness_path), rbuf = read_file(r1cs_path);
    WitnessFile w(wbuf.data(), wbuf.size());
    R1csHeader r(rbuf.data(), rbuf.size());
    if (w.curve != curve || r.curve != curve) throw Error("split-witness: the files are over a different curve");
    if (r.num_inputs() > w.n) throw Error("split-witness: fewer witness values than public inputs");
    const size_t n_pub = r.num_inputs(), n = w.n - n_pub;
    DeviceDriver d(curve, device);
    cocg_ctx* c = d.ctx;
    DevVec all = d.alloc(w.n);
    check(c, cocg_h2d(c, all.p, w.values, w.n * 32), "cocg_h2d");
    check(c, cocg_vec_op(c, COCG_OP_TO_MONT, all.p, nullptr, all.p, w.n), "cocg_vec_op");
    std::vector<uint64_t> pub(4 * n_pub);
    check(c, cocg_d2h(c, pub.data(), all.p, n_pub * 32), "cocg_d2h");
    void* x = all.at(n_pub);
    const int parties = protocol == 0 ? 3 : num_parties, k = protocol == 0 ? 2 : 1;
    std::vector<std::vector<uint64_t>> comp_a(parties, std::vector<uint64_t>(4 * n)), comp_b(k == 2 ? parties : 0, std::vector<uint64_t>(4 * n));
    if (n) {
      if (protocol == 0) {
        DevVec a = d.alloc(n), b = d.alloc(n), cc = d.alloc(n);
        check(c, cocg_prf_fill(c, seed, 0, a.p, n), "cocg_prf_fill");
        check(c, cocg_prf_fill(c, seed + 32, 0, b.p, n), "cocg_prf_fill");
        check(c, cocg_vec_op(c, COCG_OP_SUB, x, a.p, cc.p, n), "cocg_vec_op");
        check(c, cocg_vec_op(c, COCG_OP_SUB, cc.p, b.p, cc.p, n), "cocg_vec_op");
        const void* sa[3] = {a.p, b.p, cc.p};   // share_field_elements rep3.rs:124-150: (a, c), (b, a), (c, b)
        const void* sb[3] = {cc.p, a.p, b.p};
        for (int i = 0; i < 3; i++) {
          check(c, cocg_d2h(c, comp_a[i].data(), sa[i], n * 32), "cocg_d2h");
          check(c, cocg_d2h(c, comp_b[i].data(), sb[i], n * 32), "cocg_d2h");
        }
        d.release(a); d.release(b); d.release(cc);
      } else {
        std::vector<DevVec> coeff(threshold);
        for (int j = 0; j < threshold; j++) {
          coeff[j] = d.alloc(n);
          check(c, cocg_prf_fill(c, seed + 32 * j, 0, coeff[j].p, n), "cocg_prf_fill");
        }
        DevVec acc = d.alloc(n);
        for (int i = 0; i < parties; i++) {
          Fr xi = d.fr.zero(), one = d.fr.one(), pw;
          for (int t = 0; t <= i; t++) xi = d.fr.add(xi, one);  // i + 1 in Montgomery form
          pw = xi;
          const void* cur = x;
          for (int j = 0; j < threshold; j++) {  // acc = x + sum_j coeff_j * (i+1)^(j+1)
            check(c, cocg_vec_axpy(c, pw.l, coeff[j].p, cur, acc.p, n), "cocg_vec_axpy");
            cur = acc.p;
            pw = d.fr.mul(pw, xi);
          }
          check(c, cocg_d2h(c, comp_a[i].data(), acc.p, n * 32), "cocg_d2h");
        }
        d.release(acc);
        for (auto& v : coeff) d.release(v);
      }
    }
    d.release(all);
    std::string base = witness_path;
    size_t slash = base.find_last_of('/');
    if (slash != std::string::npos) base = base.substr(slash + 1);
    for (int i = 0; i < parties; i++) {
      const uint64_t* comps[2] = {comp_a[i].data(), k == 2 ? comp_b[i].data() : nullptr};
      std::vector<uint8_t> img = shared_witness_encode(curve, pub.data(), n_pub, comps, k, n);
      std::string path = std::string(out_dir) + "/" + base + "." + std::to_string(i) + ".shared";
      FILE* f = fopen(path.c_str(), "wb");
      if (!f) throw Error("cannot create " + path);
      size_t wr = fwrite(img.data(), 1, img.size(), f);
      fclose(f);
      if (wr != img.size()) throw Error("short write on " + path);
    }
  });
}

// ------------------------------------------------------------------------------------------------ Groth16 verification (pairing.hpp)
// *ok = 1 accepted, 0 rejected by the pairing check.  A non-zero return is a malformed input (bad counts, a point off the curve or
// outside the subgroup, unparsable JSON) -- where the reference fails while deserialising, before Groth16::verify runs.  No GPU needed.
// vk: alpha_g1 | beta_g2 | gamma_g2 | delta_g2 packed affine Montgomery; ic: n_ic G1 points; proof: A | B | C; pub: n_ic - 1 Montgomery Fr.
extern "C" int cohost_groth16_verify(int curve, const void* vk, const void* ic, size_t n_ic, const void* proof, const void* pub, int* ok) {
  if (!vk || !ic || !proof || !ok || (n_ic > 1 && !pub)) return fail("cohost_groth16_verify: null argument");
  if (curve != COCG_BN254 && curve != COCG_BLS12_381) return fail("cohost_groth16_verify: unknown curve");
  return guarded([&] {
    const size_t lq = curve == COCG_BN254 ? 4 : 6;
    const uint64_t* v = (const uint64_t*)vk;
    VerifyInput in{v, v + 2 * lq, v + 6 * lq, v + 10 * lq, (const uint64_t*)ic, n_ic, (const uint64_t*)proof, (const uint64_t*)pub};
    *ok = groth16_verify(curve, in) ? 1 : 0;
  });
}
// `co-circom verify groth16 --vk verification_key.json --proof proof.json --public-input public.json` (co-circom.rs:640-720)
extern "C" int cohost_groth16_verify_json(const char* vk_json, size_t vk_len, const char* proof_json, size_t proof_len, const char* public_json,
                                          size_t public_len, int* ok) {
  if (!vk_json || !proof_json || !public_json || !ok) return fail("cohost_groth16_verify_json: null argument");
  return guarded([&] { *ok = groth16_verify_json(vk_json, vk_len, proof_json, proof_len, public_json, public_len) ? 1 : 0; });
}

// `co-circom verify plonk` (co-plonk/src/plonk.rs:123-283).  challenges_out: NULL, or 6 Montgomery Fr = alpha, beta, gamma, xi, v[0], u.
extern "C" int cohost_plonk_verify_json(const char* vk_json, size_t vk_len, const char* proof_json, size_t proof_len, const char* public_json,
                                        size_t public_len, void* challenges_out, int* ok) {
  if (!vk_json || !proof_json || !public_json || !ok) return fail("cohost_plonk_verify_json: null argument");
  return guarded([&] { *ok = plonk_verify_json(vk_json, vk_len, proof_json, proof_len, public_json, public_len, (uint64_t*)challenges_out) ? 1 : 0; });
}

// Plonk zkey header without touching a GPU: info[7] = curve, n_vars, n_public, domain_size, n_additions, n_constraints, and a bit mask of
// the optional parts present (1 = verifying-key tail, 2 = selector sections, 4 = sigma, 8 = Lagrange).  k1k2: 2 Montgomery Fr;
// vk_g1: Qm Ql Qr Qo Qc S1 S2 S3 packed affine Montgomery; x_2: one G2 point.  Output pointers may be NULL.
extern "C" int cohost_plonk_zkey_header(const char* path, size_t* info, void* k1k2, void* vk_g1, void* x_2) {
  if (!path || !info) return fail("cohost_plonk_zkey_header: null argument");
  return guarded([&] {
    std::vector<uint8_t> buf = read_file(path);
    PlonkZKeyFile f(buf.data(), buf.size());
    info[0] = (size_t)f.curve; info[1] = f.n_vars; info[2] = f.n_public; info[3] = f.domain_size; info[4] = f.n_additions; info[5] = f.n_constraints;
    bool sels = true;
    for (int k = 0; k < 5; k++) sels = sels && f.sel[k];
    info[6] = (f.k1 ? 1 : 0) | (sels ? 2 : 0) | (f.sigma ? 4 : 0) | (f.lagrange ? 8 : 0);
    if (!f.k1) return;
    if (k1k2) { memcpy(k1k2, f.k1, 32); memcpy((char*)k1k2 + 32, f.k2, 32); }
    if (vk_g1) memcpy(vk_g1, f.vk_g1, 8 * 2 * f.n8q);
    if (x_2) memcpy(x_2, f.x_2, 4 * f.n8q);
  });
}
