#!/usr/bin/env python3
"""Groth16 proofs/s on a synthetic 2^20-constraint BN254 circuit, REP3 (three parties in process), on N B200s.

One "step" = one complete collaborative proof: the three REP3 parties (threads of this process, joined by the in-process
network, as the reference's own bench runs them -- tests/benches/poseidon_hash2.rs:197-222) each run CoGroth16::prove
(witness_map_from_matrices: 2 SpMV pairs, 2 mul_vec rounds, 12 NTTs; create_proof_with_assignment: 8 G1 + 2 G2 MSMs of ~2^20)
on the GPU.  With N > 1 every MSM is sharded by index range over the N ranks (one process per GPU) and the partial sums are
combined with ONE NCCL all-gather per proof; the NTT pipeline is replicated.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 20] [--impl reference]

--impl reference times the CPU restatement of the reference's path (oracle/, OpenMP over all host cores) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BN254_R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
BLS381_R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
SEED = 0xC0C12C0D20240001
PRF_SEEDS = bytes(range(96))  # the parties' PRF seeds: fixed HERE ONLY so that every run (and every rank) produces the same proofs


def limbs_of(v):
    return np.array([(v >> (64 * i)) & (2**64 - 1) for i in range(4)], dtype=np.uint64)


def rand_fr(n, rng):
    """uniform 252-bit values: valid Montgomery residues of BN254 Fr"""
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def synthetic_r1cs(log_n, rng):
    """Shape-faithful synthetic constraint system: n = 2^log_n domain, l = 1 public input,
    num_constraints = n - 2, m = n variables, A and B with 2 non-zeros per row at columns with poseidon-like locality."""
    n = 1 << log_n
    n_public, n_vars, rows = 1, n, n - 2

    def mat():
        rowptr = (2 * np.arange(rows + 1)).astype(np.uint32)
        base = np.repeat(np.arange(rows, dtype=np.int64), 2)
        col = (base + rng.integers(-64, 64, size=2 * rows)) % n_vars
        return rowptr, col.astype(np.uint32), rand_fr(2 * rows, rng)

    return n_public, n_vars, rows, mat(), mat()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def ncu_traffic(world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (msm_accumulate_kernel, weighted 4 G1 : 1 G2
    like the launches of one share component) from the committed `ncu --set full` capture summary, single-GPU shape only."""
    p = os.path.join(ROOT, "profiles", "r02_msm_accumulate_traffic.json")
    if world != 1 or not os.path.exists(p):
        return None
    t = json.load(open(p))
    if t.get("g2_bytes_per_launch") is None:
        return t["g1_bytes_per_launch"]  # only the G1 instantiation was captured
    return (4 * t["g1_bytes_per_launch"] + t["g2_bytes_per_launch"]) / 5


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU restatement
def host_threads():
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not inherit that
    (round-1 verdict: the N > 1 reference runs were measured on ONE core), so the OpenMP pool is sized explicitly."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuGroth16Slice:
    """The reference's CPU path for ONE (party, share component) slice of a REP3 Groth16 proof, on the C oracle (oracle/c/cocg_oracle.c:
    Pippenger with arkworks' window rule, radix-2 NTT, OpenMP over all host cores).  A proof is exactly SIX such slices of identical
    work (3 parties x 2 share components), each: 2 SpMV (A and B rows on one component), 1 mul_vec local step, 3 coset transforms
    (6 NTTs + 3 coset scalings), 1 sub, and the 5 MSMs (h, l, a, b_g1 in G1; b_g2 in G2) at the full 2^log_n size -- nothing is
    scaled down, so proofs/s = 1 / (6 x slice time).  Inputs are built once, outside the timed steps."""

    SLICES_PER_PROOF = 6

    def __init__(self, log_n):
        from oracle import cref, ntt as ontt
        from oracle.curves import BN254 as C

        self.cref, self.C, self.log_n = cref, C, log_n
        L = cref.lib()
        L.orc_set_threads(host_threads())
        self.cores = L.orc_num_threads()
        rng = np.random.default_rng(SEED)
        n = 1 << log_n
        self.n = n
        n_public, n_vars, self.rows, self.A, self.B = synthetic_r1cs(log_n, rng)
        n_aux = n_vars - n_public - 1

        def chain(group, cnt, k):
            p0 = cref.g_to_mont(C, [C.mul(C.gen(group), 1000 + k, group)], group)[0]
            q = cref.g_to_mont(C, [C.mul(C.gen(group), 77 + k, group)], group)[0]
            return cref.gen_chain(C, group, p0, q, cnt)

        self.h_q, self.l_q, self.a_q, self.b1_q = chain(1, n, 1), chain(1, n_aux, 2), chain(1, n_aux, 3), chain(1, n_aux, 4)
        self.b2_q = chain(2, n_aux, 5)
        self.wa, self.wb = rand_fr(n_aux, rng), rand_fr(n_aux, rng)
        self.pub = rand_fr(n_public + 1, rng)
        omega, g = ontt.groth16_roots(C, log_n)
        self.om, self.omi = cref.fr_to_mont(C, [omega]), cref.fr_to_mont(C, [pow(omega, -1, C.r)])
        self.gm, self.one = cref.fr_to_mont(C, [g]), cref.fr_to_mont(C, [1])

    def step(self):
        """One slice; returns seconds."""
        cref, C, n, rows = self.cref, self.C, self.n, self.rows
        t0 = time.perf_counter()
        z = np.concatenate([self.pub, self.wa])

        def rows_of(M):
            out = np.zeros((n, 4), dtype=np.uint64)
            out[:rows] = cref.spmv(C, M[0], M[1], M[2], z)
            return out

        a, b = rows_of(self.A), rows_of(self.B)
        c = cref.rep3_mul_local(C, a, b, b, a, None)  # the local step reads both components of both operands

        def coset(v):
            v = cref.ntt(C, v, self.omi, inverse=True)
            v = cref.distribute_powers(C, v, self.gm, self.one)
            return cref.ntt(C, v, self.om)

        a, b, c = coset(a), coset(b), coset(c)
        h = cref.fr_vec_op(C, cref.OP_SUB, a, c)
        cref.msm(C, 1, self.h_q, h)
        cref.msm(C, 1, self.l_q, self.wa)
        cref.msm(C, 1, self.a_q, self.wa)
        cref.msm(C, 1, self.b1_q, self.wa)
        cref.msm(C, 2, self.b2_q, self.wa)
        return time.perf_counter() - t0

    def sample_text(self, steps):
        return (f"each step = 1 of the 6 identical (party, share component) slices of one 2^{self.log_n} REP3 proof at full size "
                f"(2 SpMV, 1 mul_vec local step, 6 NTTs + 3 coset scalings, 4 G1 + 1 G2 MSMs of 2^{self.log_n}); C oracle (port of the "
                f"reference's arkworks algorithms: the Rust reference cannot be built in this image), OpenMP on {self.cores} host threads; "
                f"proofs/s = 1 / (6 x mean step time) over {steps} timed steps")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "msm":
        return run_reference_msm(args)
    if args.workload == "plonk":
        return run_reference_plonk(args)
    sl = CpuGroth16Slice(args.log_n)
    for _ in range(args.warmup):
        sl.step()
    times = [sl.step() for _ in range(args.steps)]
    t = float(np.mean(times))
    value = 1.0 / (sl.SLICES_PER_PROOF * t)
    print(json.dumps({
        "impl": "reference", "metric": "groth16_proofs_per_sec", "value": value, "unit": "proofs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64 limbs (256-bit Montgomery)", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": sl.cores, "kind": "port", "sample": sl.sample_text(args.steps)},
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world):
    n = 1 << args.log_n
    default = args.curve == "bn254" and args.protocol == "rep3" and args.log_n == 20
    comps = 2 if args.protocol == "rep3" else 1
    return {"workload": f"Groth16 prove, synthetic 2^{args.log_n}-constraint R1CS, {args.curve.upper()}, {args.protocol.upper()} 3-party in-process"
                        + (" (BASELINE configs[2])" if default else " (non-default configuration)"),
            "curve": args.curve, "protocol": args.protocol, "domain_size": n, "n_vars": n, "n_public": 1, "nnz_per_row": 2,
            "msm_per_proof": f"3 parties x {comps} components x (4 G1 + 1 G2)", "ntt_per_proof": 18 * comps,
            "parallelism": "single GPU" if world == 1 else (
                f"27 units of the proof (per party: witness map + h MSMs; per party and share component: the b_g2 MSM and each of the l / a / b_g1 "
                f"MSMs) placed largest-first on {world} GPUs at full size, mul_vec payloads GPU to GPU over NCCL, 1 all-gather/proof"
                if args.protocol == "rep3" and getattr(args, "shard_mode", "blocks") == "blocks" else
                f"MSM bases sharded by index range over {world} GPUs, NTT replicated, 1 all-gather/proof"),
            "l2_policy": "inputs_exceed_l2 (>= 0.9 GB of bases + share vectors streamed per proof vs 126 MB L2)"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_own(args):
    import torch
    import torch.distributed as dist

    import cocg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL's kernels on a high-priority stream: a send / receive of a mul_vec round must not queue behind the thousands of pending
        # thread blocks of an MSM launch on the same GPU (the witness maps of three ranks advance in lock-step through these rounds)
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    rng = np.random.default_rng(SEED)
    log_n = args.log_n
    n = 1 << log_n
    n_public, n_vars, rows, A, B = synthetic_r1cs(log_n, rng)
    n_aux = n_vars - n_public - 1
    seed_bytes = SEED.to_bytes(8, "little") * 4
    t_setup = time.perf_counter()
    curve_id = cocg.BN254 if args.curve == "bn254" else cocg.BLS12_381
    modulus = BN254_R if args.curve == "bn254" else BLS381_R
    if args.protocol == "shamir":
        return run_own_shamir(args, cocg, torch, curve_id, modulus, local, (n_public, n_vars, rows, A, B), seed_bytes, rng)
    from importlib import import_module
    distmod = import_module("collaborative-circom_b200.distributed")
    mode = args.shard_mode if world > 1 else "ranges"
    zk = cocg.Groth16ZKey(curve_id, n_public, n_vars, log_n, rows, A, B, device=local, synthetic_seed=seed_bytes, rank=rank, world=world, shard_mode=mode)
    sess = cocg.Rep3Session(zk, seeds=PRF_SEEDS, rank=rank, world=world, comm=distmod.make_p2p(torch.device("cuda", local)) if mode == "blocks" else None)
    # witness: x = x0 + x1 + x2, party i holds (x_i, x_{i-1}) (rep3.rs:57-68); pinned host copies + resident device copies
    xs = []
    for i in range(3):
        t = torch.empty(n_aux * 4, dtype=torch.int64).pin_memory()
        t.numpy().view(np.uint64).reshape(n_aux, 4)[:] = rand_fr(n_aux, rng)
        xs.append(t)
    host_a = [xs[i].data_ptr() for i in range(3)]
    host_b = [xs[(i - 1) % 3].data_ptr() for i in range(3)]
    ctx = cocg.Context(curve_id, local)
    dev = [ctx.upload(xs[i].numpy().view(np.uint64).reshape(n_aux, 4)) for i in range(3)]
    dev_a = [dev[i].ptr for i in range(3)]
    dev_b = [dev[(i - 1) % 3].ptr for i in range(3)]
    r1 = pow(2, 256, modulus)
    pub = np.stack([limbs_of(r1), limbs_of(12345 * r1 % modulus)])
    setup_s = time.perf_counter() - t_setup

    all_gather = distmod.make_all_gather(world, torch.device("cuda", local))

    def step(device_resident):
        if device_resident:
            return sess.prove(pub, dev_a, dev_b, all_gather=all_gather, device_ptrs=True)
        return sess.prove(pub, host_a, host_b, all_gather=all_gather)

    step_walls, outliers = {}, {}

    def timed(device_resident, steps, sample_clocks=False):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        walls, step_phases = [], []
        for _ in range(steps):
            tw = time.perf_counter()
            proofs = step(device_resident)
            walls.append((time.perf_counter() - tw) * 1e3)
            if world > 1:
                step_phases.append(sess.phase_times() * 1e3)
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        step_walls[device_resident] = [round(float(np.min(walls)), 2), round(float(np.median(walls)), 2), round(float(np.max(walls)), 2), int(np.argmax(walls))]
        if world > 1:  # a step far above the median: every rank's per-party phase times of THAT step (which rank, which phase stalled)
            worst = torch.tensor([int(np.argmax(walls)) if np.max(walls) > 1.4 * np.median(walls) else -1], device="cuda")
            dist.broadcast(worst, 0)
            w = int(worst.item())
            rows = [None] * world
            dist.all_gather_object(rows, None if w < 0 else {"rank": rank, "wall_ms": round(walls[w], 2),
                                                              "party_phase_ms": [[round(float(x), 2) for x in r] for r in step_phases[w]]})
            outliers[device_resident] = None if w < 0 else {"step": w, "ranks": rows}
        return float(ms.item()), clocks, proofs

    # `value` leg: everything resident in HBM -- witness shares AND the mul_vec payloads the three co-located parties exchange;
    # `e2e` leg: witness shares from pinned host memory and the MPC payloads staged through pinned host memory, as a party that has
    # to reach a NIC would (north_star: the MPC rounds stay on the host network stack)
    sess.set_mpc_exchange("device")
    for _ in range(args.warmup):
        step(True)
    sess.profile(True)
    sess.profile_reset()
    launches0 = sess.launch_count()
    ms, clocks, proofs = timed(True, args.steps, sample_clocks=True)
    launches = sess.launch_count() - launches0
    prof = sess.profile_read()
    sess.profile(False)
    assert np.array_equal(proofs[0], proofs[1]) and np.array_equal(proofs[1], proofs[2]), "the three parties disagree on the proof"
    phases_value = sess.phase_times().max(axis=0) * 1e3  # last proof of the value leg, slowest party per phase
    # N > 1 diagnostics (outside the timed region): per rank, the host wall time of each API call of one more value-leg proof and
    # the parties' phase times -- shows which rank the all-gather waits for and what surrounds it
    rank_diag = None
    if world > 1:
        api = np.zeros(5)
        reps = 4
        for _ in range(reps):
            dist.barrier()
            t0 = time.perf_counter()
            sess.begin(pub, dev_a, dev_b, None, True)
            t1 = time.perf_counter()
            part = sess.partials()
            t2 = time.perf_counter()
            gathered = all_gather(part)
            t3 = time.perf_counter()
            sess.combine(gathered)
            t4 = time.perf_counter()
            sess.end()
            t5 = time.perf_counter()
            api += np.array([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4]) * 1e3 / reps
        mine = {"rank": rank, "api_ms": {k: round(float(v), 2) for k, v in zip(("begin", "partials_wait", "all_gather", "combine", "end"), api)},
                "party_phase_ms": [[round(float(x) * 1e3, 2) for x in row] for row in sess.phase_times()]}
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        rank_diag = everyone
    # The PRF seeds are fixed and every mask / blinder is addressed by a counter the parties advance in lock-step, so proof number
    # warmup + steps of this leg is the same byte string in every run and at every N (sharding only changes who adds which points).
    import hashlib
    proof_sha256 = hashlib.sha256(np.ascontiguousarray(proofs[0]).tobytes()).hexdigest()
    sess.set_mpc_exchange("host")
    for _ in range(min(args.warmup, 2)):
        step(False)
    ms_e2e, _, _ = timed(False, args.steps)
    phases = sess.phase_times().max(axis=0) * 1e3  # last e2e proof, slowest party per phase
    # supplementary, N > 1: the same N GPUs as independent replicas (one whole proof per rank per step, no collective) -- the
    # throughput-optimal deployment; the headline `value` stays the sharded, one-all-gather-per-proof mode
    replicas = None
    if world > 1:
        zk1 = cocg.Groth16ZKey(curve_id, n_public, n_vars, log_n, rows, A, B, device=local, synthetic_seed=seed_bytes)
        sess1 = cocg.Rep3Session(zk1, seeds=PRF_SEEDS)
        for _ in range(2):
            sess1.prove(pub, host_a, host_b)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            sess1.prove(pub, host_a, host_b)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        replicas = {"value": world * args.steps / (float(t.item()) / 1e3), "unit": "proofs/s", "scaling": "weak",
                    "note": "e2e leg (witness uploaded per proof), one independent 3-party proof per GPU per step, no collective"}
        sess1.close()
        zk1.close()

    # The accumulate kernel timed ALONE (one context, one stream, nothing else on the GPU): the in-situ scope times above contain the
    # time slices the GPU gives to the other parties' kernels.  One G1 and one G2 query of the size a rank runs (the full query in
    # block mode and at N = 1; this rank's index range in `ranges` mode), one share component each.
    alone = None
    if rank == 0:
        per_alone = n_aux if mode == "blocks" or world == 1 else (n_aux + world - 1) // world
        alone = {"terms": per_alone}
        for grp, key in ((1, "g1"), (2, "g2")):
            hb = ctx.bases_generate(grp, per_alone, bytes([77 + grp] * 32))
            ctx.msm(hb, [dev[0]], n=per_alone)
            ctx.profile(True)
            ctx.profile_reset()
            for _ in range(5):
                ctx.msm(hb, [dev[0]], n=per_alone)
            ctx.sync()
            pa = ctx.profile_read()
            ctx.profile(False)
            ctx.bases_free(hb)
            alone[key + "_ms"] = pa["msm_accumulate"][0] / max(pa["msm_accumulate"][1], 1)
        alone["window_bits"], alone["windows"] = cocg.msm_plan(curve_id, per_alone)

    value = args.steps / (ms / 1e3)
    e2e = args.steps / (ms_e2e / 1e3)
    peak, peak_src = measured_peak()
    # roofline of the dominant kernel: MSM bucket accumulation.  Algorithmic bytes per launch = one read of every point + one read of
    # its scalar: G1 96 B/term, G2 160 B/term.  A proof launches it 4 x G1 : 1 x G2 per party and share component, so
    # the reported figure is (4 x 96 + 160) x terms bytes over (4 x t_G1 + t_G2), with the kernel times measured alone (above);
    # `in_situ` is the same ratio from the event scopes inside the timed proofs (they overlap with the other parties' kernels).
    per = n_aux if mode == "blocks" or world == 1 else (n_aux + world - 1) // world
    perh = n if mode == "blocks" or world == 1 else (n + world - 1) // world
    bytes_per_proof = 3 * 2 * (perh * 96 + 3 * per * 96 + per * 160)
    acc_ms, acc_n = prof["msm_accumulate"]
    in_situ = None
    if acc_n and world == 1:
        in_situ = bytes_per_proof / (acc_ms / args.steps * 1e-3) / 1e9
    achieved = 0.0
    if alone:
        achieved = alone["terms"] * (4 * 96 + 160) / ((4 * alone["g1_ms"] + alone["g2_ms"]) * 1e-3) / 1e9
    kernels = {}
    alg = {"msm_sort": bytes_per_proof, "msm_accumulate": bytes_per_proof, "msm_reduce": bytes_per_proof,
           "ntt": 36 * 64 * n, "vec": 3 * (2 * 160 + 2 * 96) * n, "spmv": 3 * 2 * 2 * (2 * rows * 68 + rows * 36)}
    for name, (tms, cnt) in prof.items():
        per_proof_ms = tms / args.steps
        kernels[name] = {"ms_per_proof_summed_over_parties": round(per_proof_ms, 3), "scopes": cnt,
                         "algorithmic_GBps": round(alg[name] / (per_proof_ms * 1e-3) / 1e9, 1) if per_proof_ms else None}
    if rank == 0:
        out = {
            "metric": "groth16_proofs_per_sec", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 limbs (256-bit Montgomery integers; no floating point)", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": e2e, "unit": "proofs/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": 3 * 2 * n_aux * 32 + 3 * 2 * 32, "d2h_bytes_per_step": 3 * 8 * 4 * 8,
                    "mpc_exchange_bytes_per_step": 2 * 3 * 2 * n * 32,
                    "note": ("witness shares in pinned host memory uploaded every step, proofs read back, and the two mul_vec rounds of each "
                             "party (n x 32 B out + in per round) staged through pinned host memory; the `value` leg keeps all of that in HBM")
                            if world == 1 or mode != "blocks" else
                            ("witness share components in pinned host memory uploaded every step by the ranks whose blocks read them, proofs read "
                             "back; the parties' witness maps run on different GPUs, so their mul_vec payloads travel GPU to GPU over NVLink "
                             "(NCCL send / recv) in both legs; the `value` leg keeps the witness in HBM")},
            "gpu_launches": int(launches),
            "proof_sha256": proof_sha256,
            "proof_sha256_of": f"proof number {args.warmup + args.steps} of the value leg (A | B | C packed affine Montgomery limbs), fixed PRF seeds",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(world),
                         "kernel": "msm_accumulate_kernel: algorithmic bytes of the 4 G1 + 1 G2 launches of one party and share component over their "
                                   "durations, each instantiation timed alone with CUDA events on its launching stream in this run",
                         "in_situ": in_situ, "peak_source": peak_src,
                         "note": "MSM is bound by 32-bit integer multiply-add issue, not HBM (DESIGN.md section 4): see `issue`"},
            "replicas": replicas, "rank_diag": rank_diag,
            "step_wall_ms_rank0": {"value_leg_min_median_max_argmax": step_walls.get(True), "e2e_leg_min_median_max_argmax": step_walls.get(False),
                                   "outlier_value_leg": outliers.get(True), "outlier_e2e_leg": outliers.get(False)},
            "kernels": kernels, "setup_s": round(setup_s, 2),
            "host_phases_ms": {"witness_map": round(float(phases[0]), 2), "msm": round(float(phases[1]), 2),
                               "all_gather_wait": round(float(phases[2]), 2), "assembly": round(float(phases[3]), 2),
                               "value_leg": [round(float(x), 2) for x in phases_value]},
        }
        if alone:
            # issue roofline of the same kernel: 10 Fq multiplications (8M + 2S) per table point added
            fq_mul = alone["terms"] * alone["windows"] * 10 / (alone["g1_ms"] * 1e-3) / 1e9
            # measured in this run: a pure dependent chain of the library's own Montgomery product on every SM (cocg_fp_mul_ceiling);
            # the hardware figure beside it is 148 SMs x 32 IMAD.WIDE/clk x sm_max_mhz / 128 wide products per 8-limb multiplication
            ceiling = ctx.fp_mul_ceiling(base_field=True)
            hw = 148 * 32 * ((clocks or {}).get("sm_max_mhz") or 1965.0) * 1e6 / 128 / 1e9
            ceiling_src = f"fp_mul chain measured in this run (cocg_fp_mul_ceiling); IMAD.WIDE hardware bound {hw:.1f} G/s"
            out["roofline"].update({
                "g1": {"ms": alone["g1_ms"], "GBps": alone["terms"] * 96 / (alone["g1_ms"] * 1e-3) / 1e9},
                "g2": {"ms": alone["g2_ms"], "GBps": alone["terms"] * 160 / (alone["g2_ms"] * 1e-3) / 1e9},
                "terms_per_launch": alone["terms"],
                "issue": {"unit": "G Fq-mul/s", "achieved": fq_mul, "peak": ceiling, "frac": fq_mul / ceiling,
                          "note": f"G1 accumulate alone: {alone['terms']} terms x {alone['windows']} windows (c = {alone['window_bits']}) x 10 "
                                  f"Montgomery products; peak = {ceiling_src}"}})
        if world == 1 and not args.no_cpu_baseline:
            sl = CpuGroth16Slice(log_n)
            k_cpu = 2
            t = float(np.mean([sl.step() for _ in range(k_cpu)]))
            out["cpu_baseline"] = {"value": 1.0 / (sl.SLICES_PER_PROOF * t), "unit": "proofs/s", "cores": sl.cores, "kind": "port",
                                   "sample": sl.sample_text(k_cpu)}
        print(json.dumps(out))
    sess.close()
    zk.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ BASELINE configs[1]: one MSM
def msm_config(args, world):
    return {"workload": f"BN254 G1 MSM, 2^{args.log_n} random scalars x synthetic points, one share component (BASELINE configs[1])",
            "curve": "bn254", "group": "G1", "terms": 1 << args.log_n,
            "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (one MSM per GPU per step, no collective)",
            "l2_policy": "inputs_exceed_l2 (window table 0.87 GB + 32 MiB of scalars per MSM vs 126 MB L2)"}


def run_reference_msm(args):
    from oracle import cref
    from oracle.curves import BN254 as C
    L = cref.lib()
    L.orc_set_threads(host_threads())
    cores = L.orc_num_threads()
    n = 1 << args.log_n
    rng = np.random.default_rng(SEED)
    p0 = cref.g_to_mont(C, [C.mul(C.gen(1), 1001, 1)], 1)[0]
    q = cref.g_to_mont(C, [C.mul(C.gen(1), 78, 1)], 1)[0]
    pts = cref.gen_chain(C, 1, p0, q, n)
    sc = rand_fr(n, rng)
    for _ in range(args.warmup):
        cref.msm(C, 1, pts, sc)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cref.msm(C, 1, pts, sc)
    t = (time.perf_counter() - t0) / args.steps
    sample = f"one full 2^{args.log_n}-term G1 MSM per step on the C oracle (Pippenger, arkworks' window rule, OpenMP on {cores} host threads)"
    print(json.dumps({
        "impl": "reference", "metric": "msm_per_sec", "value": 1.0 / t, "unit": "MSM/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 limbs (256-bit Montgomery)", "data": "synthetic", "config": msm_config(args, 1),
        "cpu_baseline": {"value": 1.0 / t, "unit": "MSM/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": 1.0 / t, "unit": "MSM/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_own_msm(args):
    import torch
    import torch.distributed as dist

    import cocg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << args.log_n
    rng = np.random.default_rng(SEED + rank)
    ctx = cocg.Context(cocg.BN254, local)
    h = ctx.bases_generate(1, n, SEED.to_bytes(8, "little") * 4)
    host = torch.empty(n * 4, dtype=torch.int64).pin_memory()
    sc = host.numpy().view(np.uint64).reshape(n, 4)
    sc[:] = rand_fr(n, rng)
    dev = ctx.upload(sc)

    def timed(fn, steps, sample_clocks=False):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks, out

    # cocg kernels run on the context's own stream and every entry point used here synchronises before returning, so the torch
    # events on the current stream bracket them; the per-kernel times below come from CUDA events on the launching stream itself
    resident = lambda: ctx.msm(h, [dev])
    e2e_fn = lambda: ctx.msm_host(h, [sc])
    for _ in range(args.warmup):
        resident()
    ctx.profile(True)
    ctx.profile_reset()
    l0 = ctx.launch_count()
    ms, clocks, out = timed(resident, args.steps, sample_clocks=True)
    launches = ctx.launch_count() - l0
    prof = ctx.profile_read()
    ctx.profile(False)
    for _ in range(min(args.warmup, 2)):
        e2e_fn()
    ms_e2e, _, out2 = timed(e2e_fn, args.steps)
    # Jacobian representatives depend on the (atomic) order in which a bucket's points were added: compare the affine point
    aff, aff2 = ctx.ec_op(1, cocg.EC_TO_AFFINE, out[0]), ctx.ec_op(1, cocg.EC_TO_AFFINE, out2[0])
    assert np.array_equal(aff, aff2), "host-pointer and device-pointer entry points disagree"
    if rank == 0:
        peak, peak_src = measured_peak()
        acc_ms = prof["msm_accumulate"][0] / max(prof["msm_accumulate"][1], 1)
        c_bits, nwin = cocg.msm_plan(cocg.BN254, n)
        ceiling = ctx.fp_mul_ceiling(base_field=True)
        fq_mul = n * nwin * 10 / (acc_ms * 1e-3) / 1e9
        value = world * args.steps / (ms / 1e3)
        o = {
            "metric": "msm_per_sec", "value": value, "unit": "MSM/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (256-bit Montgomery integers; no floating point)", "data": "synthetic", "config": msm_config(args, world),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": world * args.steps / (ms_e2e / 1e3), "unit": "MSM/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": n * 32,
                    "d2h_bytes_per_step": 96, "note": "cocg_msm_host: scalars in pinned host memory uploaded every step, result point read back"},
            "result_sha256": __import__("hashlib").sha256(np.ascontiguousarray(aff).tobytes()).hexdigest(),
            "roofline": {"bound": "hbm", "achieved": n * 96 / (acc_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": n * 96 / (acc_ms * 1e-3) / 1e9 / peak, "traffic": ncu_traffic(1), "peak_source": peak_src,
                         "kernel": "msm_accumulate_kernel, one launch per MSM, timed with CUDA events on its launching stream",
                         "whole_msm": {"achieved": n * 96 / (ms / args.steps * 1e-3) / 1e9, "frac": n * 96 / (ms / args.steps * 1e-3) / 1e9 / peak},
                         "issue": {"unit": "G Fq-mul/s", "achieved": fq_mul, "peak": ceiling, "frac": fq_mul / ceiling,
                                   "note": f"{n} terms x {nwin} windows (c = {c_bits}) x 10 Montgomery products per mixed addition; peak = fp_mul "
                                           "chain measured in this run (cocg_fp_mul_ceiling)"}},
            "kernels": {k: {"ms_per_msm": round(v[0] / max(args.steps, 1), 4), "scopes": v[1]} for k, v in prof.items() if v[1]},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cref
            from oracle.curves import BN254 as C
            L = cref.lib()
            L.orc_set_threads(host_threads())
            pts = ctx.bases_download(h, 0, n)
            k_cpu = 5
            t0 = time.perf_counter()
            for _ in range(k_cpu):
                want = cref.msm(C, 1, pts, sc)
            t = (time.perf_counter() - t0) / k_cpu
            o["cpu_baseline"] = {"value": 1.0 / t, "unit": "MSM/s", "cores": L.orc_num_threads(), "kind": "port",
                                 "sample": f"{k_cpu} full 2^{args.log_n}-term G1 MSMs on the C oracle (Pippenger, arkworks' window rule, OpenMP), same points and scalars"}
            o["matches_cpu_oracle"] = bool(cref.jac_from_mont(C, out[0], 1) == cref.jac_from_mont(C, want, 1))
        print(json.dumps(o))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ BASELINE configs[3]: Plonk
def plonk_config(args, world):
    n = 1 << args.log_n
    return {"workload": f"Plonk prove, synthetic 2^{args.log_n}-gate circuit, BN254, REP3 3-party in-process (BASELINE configs[3])",
            "curve": "bn254", "protocol": "rep3", "domain_size": n, "extended_domain": 4 * n, "n_public": 1,
            "per_party": "2 components x (4 iNTT(n) + 4 NTT(4n) + 2 iNTT(4n) + 9 G1 MSMs of ~n over p_tau), 2 fused quotient kernels on 4n, "
                         "prefix-product / inverse / Horner scans on n",
            "mpc_exchanges_per_party": "8 vectors of 4n (round 3, two rounds) + ~16 vectors of n (round 2)",
            "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (one proof per GPU per step, no collective)",
            "l2_policy": "inputs_exceed_l2 (>= 1.5 GB of evaluation vectors and tables streamed per proof vs 126 MB L2)"}


def plonk_synthetic_maps(log_n, rng):
    n = 1 << log_n
    n_public, n_vars, n_constraints = 1, n - 6, n - 4
    base = np.arange(n_constraints, dtype=np.int64)
    maps = [((base + rng.integers(-64, 64, size=n_constraints)) % n_vars).astype(np.uint32) for _ in range(3)]
    return n_public, n_vars, n_constraints, maps


class CpuPlonkSlice:
    """The reference's CPU work for ONE (party, share component) slice of a REP3 Plonk proof on the C oracle (OpenMP, all host threads):
    4 iNTT(n) + 4 NTT(4n) + 2 iNTT(4n), 9 MSMs over p_tau, and the element-wise work of compute_z / compute_t as the reference does it:
    half of its 52 + 8 mul_vec local steps (3 products + mask each: 26 on 4n, 4 on n) and ~20 mul_with_public passes over 4n
    (round3.rs:268-431).  A proof is six such slices; proofs/s = 1 / (6 x slice time)."""

    SLICES_PER_PROOF = 6

    def __init__(self, log_n):
        from oracle import cref, ntt as ontt
        from oracle.curves import BN254 as C
        self.cref, self.C, self.log_n = cref, C, log_n
        L = cref.lib()
        L.orc_set_threads(host_threads())
        self.cores = L.orc_num_threads()
        rng = np.random.default_rng(SEED)
        n = 1 << log_n
        self.n = n
        p0 = cref.g_to_mont(C, [C.mul(C.gen(1), 1001, 1)], 1)[0]
        q = cref.g_to_mont(C, [C.mul(C.gen(1), 78, 1)], 1)[0]
        self.p_tau = cref.gen_chain(C, 1, p0, q, n + 6)
        self.v = [rand_fr(n, rng) for _ in range(4)]
        self.e = [rand_fr(4 * n, rng) for _ in range(4)]
        _, roots = ontt.roots_of_unity(C)
        f = lambda x: cref.fr_to_mont(C, [x])
        self.w_n, self.w_ni = f(roots[log_n]), f(pow(roots[log_n], -1, C.r))
        self.w_4n, self.w_4ni = f(roots[log_n + 2]), f(pow(roots[log_n + 2], -1, C.r))

    def step(self):
        cref, C, n = self.cref, self.C, self.n
        t0 = time.perf_counter()
        for k in range(4):                       # a, b, c, z: coefficients and extended evaluations
            cref.ntt(C, self.v[k], self.w_ni, inverse=True)
            cref.ntt(C, self.e[k], self.w_4n)
        for _ in range(4):                       # round 2 products on n
            cref.rep3_mul_local(C, self.v[0], self.v[1], self.v[2], self.v[3], None)
        for _ in range(26):                      # half of the 52 mul_vec local steps of compute_t
            cref.rep3_mul_local(C, self.e[0], self.e[1], self.e[2], self.e[3], None)
        for k in range(20):                      # mul_with_public / add_mul_public passes over the 4n domain
            cref.fr_vec_op(C, cref.OP_MUL, self.e[k % 4], self.e[(k + 1) % 4])
        for k in range(2):                       # t, tz -> coefficients
            cref.ntt(C, self.e[k], self.w_4ni, inverse=True)
        for k in range(9):                       # a b c z t1 t2 t3 Wxi Wxiw
            cref.msm(C, 1, self.p_tau, self.v[k % 4])
        return time.perf_counter() - t0

    def sample_text(self, steps):
        return (f"each step = 1 of the 6 (party, share component) slices of one 2^{self.log_n}-gate REP3 Plonk proof: 4 iNTT(n) + 4 NTT(4n) + 2 iNTT(4n), "
                f"9 G1 MSMs of 2^{self.log_n}, 30 mul_vec local steps and 20 public-multiplication passes (the reference's compute_z / compute_t op "
                f"count, round3.rs:268-431); C oracle (port; the Rust reference cannot be built here), OpenMP on {self.cores} host threads; "
                f"proofs/s = 1 / (6 x mean step time) over {steps} timed steps")


def run_reference_plonk(args):
    sl = CpuPlonkSlice(args.log_n)
    for _ in range(args.warmup):
        sl.step()
    t = float(np.mean([sl.step() for _ in range(args.steps)]))
    value = 1.0 / (sl.SLICES_PER_PROOF * t)
    print(json.dumps({
        "impl": "reference", "metric": "plonk_proofs_per_sec", "value": value, "unit": "proofs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 limbs (256-bit Montgomery)", "data": "synthetic", "config": plonk_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": sl.cores, "kind": "port", "sample": sl.sample_text(args.steps)},
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_own_plonk(args):
    import hashlib

    import torch
    import torch.distributed as dist

    import cocg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    log_n = args.log_n
    n = 1 << log_n
    rng = np.random.default_rng(SEED)
    n_public, n_vars, n_constraints, maps = plonk_synthetic_maps(log_n, rng)
    zk = cocg.PlonkZKey.synthetic(cocg.BN254, log_n, n_public, n_vars, maps, SEED.to_bytes(8, "little") * 4, device=local)
    sess = cocg.PlonkSession(zk, "rep3", seeds=PRF_SEEDS)
    n_wit = zk.n_witness
    xs = []
    for i in range(3):
        t = torch.empty(n_wit * 4, dtype=torch.int64).pin_memory()
        t.numpy().view(np.uint64).reshape(n_wit, 4)[:] = rand_fr(n_wit, rng)
        xs.append(t)
    host_a = [xs[i].data_ptr() for i in range(3)]
    host_b = [xs[(i - 1) % 3].data_ptr() for i in range(3)]
    ctx = cocg.Context(cocg.BN254, local)
    dev = [ctx.upload(xs[i].numpy().view(np.uint64).reshape(n_wit, 4)) for i in range(3)]
    dev_a = [dev[i].ptr for i in range(3)]
    dev_b = [dev[(i - 1) % 3].ptr for i in range(3)]
    r1 = pow(2, 256, BN254_R)
    pub = np.stack([limbs_of(r1), limbs_of(12345 * r1 % BN254_R)])

    def timed(fn, steps, sample_clocks=False):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks, out

    resident = lambda: sess.prove(pub, dev_a, dev_b, device_ptrs=True)
    e2e_fn = lambda: sess.prove(pub, host_a, host_b)
    sess.set_mpc_exchange("device")
    for _ in range(args.warmup):
        resident()
    sess.profile(True)
    sess.profile_reset()
    l0 = sess.launch_count()
    ms, clocks, proofs = timed(resident, args.steps, sample_clocks=True)
    launches = sess.launch_count() - l0
    prof = sess.profile_read()
    sess.profile(False)
    assert np.array_equal(proofs[0], proofs[1]) and np.array_equal(proofs[1], proofs[2]), "the three parties disagree on the proof"
    rounds = sess.round_times().max(axis=0) * 1e3
    sess.set_mpc_exchange("host")
    for _ in range(min(args.warmup, 2)):
        e2e_fn()
    ms_e2e, _, _ = timed(e2e_fn, args.steps)
    rounds_e2e = sess.round_times().max(axis=0) * 1e3
    if rank == 0:
        peak, peak_src = measured_peak()
        # algorithmic bytes per proof: MSM N x (64 + 32); NTT 64 B per element per transform
        msm_terms = 3 * 2 * (3 * (n + 2) + (n + 3) + 2 * (n + 1) + (n + 6) + (n + 5) + (n + 2))
        ntt_elems = 3 * 2 * (4 * n + 4 * 4 * n + 2 * 4 * n)
        alg = {"msm_sort": msm_terms * 96, "msm_accumulate": msm_terms * 96, "msm_reduce": msm_terms * 96, "ntt": ntt_elems * 64}
        kernels = {}
        total_ms = sum(v[0] for v in prof.values()) or 1.0
        for name, (tms, cnt) in prof.items():
            per = tms / args.steps
            kernels[name] = {"ms_per_proof_summed_over_parties": round(per, 3), "scopes": cnt, "share_of_timed_scopes": round(tms / total_ms, 3),
                             "algorithmic_GBps": round(alg[name] / (per * 1e-3) / 1e9, 1) if per and name in alg else None}
        dom = max(("ntt", "msm_accumulate"), key=lambda k: prof[k][0])
        dom_ms = prof[dom][0] / args.steps
        achieved = alg[dom] / (dom_ms * 1e-3) / 1e9
        value = world * args.steps / (ms / 1e3)
        exch = 3 * (8 * 4 * n + 16 * n) * 32
        o = {
            "metric": "plonk_proofs_per_sec", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (256-bit Montgomery integers; no floating point)", "data": "synthetic", "config": plonk_config(args, world),
            "clocks": clocks, "gpu_launches": int(launches),
            "proof_sha256": hashlib.sha256(np.ascontiguousarray(proofs[0]).tobytes()).hexdigest(),
            "e2e": {"value": world * args.steps / (ms_e2e / 1e3), "unit": "proofs/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": 3 * 2 * n_wit * 32 + 64, "d2h_bytes_per_step": int(3 * zk.proof_limbs * 8),
                    "mpc_exchange_bytes_per_step": int(2 * exch),
                    "note": "witness shares uploaded from pinned host memory every step, proofs read back, every MPC payload staged through pinned "
                            "host memory (D2H + H2D); the `value` leg keeps witness and payloads in HBM"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": f"{dom} class (largest share of device time), CUDA events on the launching streams, three parties running concurrently",
                         "peak_source": peak_src},
            "kernels": kernels,
            "host_round_ms": {"value_leg": [round(float(x), 2) for x in rounds], "e2e_leg": [round(float(x), 2) for x in rounds_e2e]},
        }
        if world == 1 and not args.no_cpu_baseline:
            sl = CpuPlonkSlice(log_n)
            k_cpu = 2
            t = float(np.mean([sl.step() for _ in range(k_cpu)]))
            o["cpu_baseline"] = {"value": 1.0 / (sl.SLICES_PER_PROOF * t), "unit": "proofs/s", "cores": sl.cores, "kind": "port", "sample": sl.sample_text(k_cpu)}
        print(json.dumps(o))
    sess.close()
    zk.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_own_shamir(args, cocg, torch, curve_id, modulus, local, r1cs, seed_bytes, rng):
    """BASELINE configs[4] flavour: CoGroth16<ShamirProtocol>, 3 parties, threshold 1.  With N > 1 ranks every MSM is sharded by index
    range and the partial sums of all parties travel in one all-gather per proof (the witness map is replicated).  The double-random
    preprocessing of the two mul_vec rounds (shamir.rs:923-1010) is inside the timed region, as in the reference."""
    import torch.distributed as dist
    from importlib import import_module
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    n_public, n_vars, rows, A, B = r1cs
    log_n = args.log_n
    n_aux = n_vars - n_public - 1
    zk = cocg.Groth16ZKey(curve_id, n_public, n_vars, log_n, rows, A, B, device=local, synthetic_seed=seed_bytes, rank=rank, world=world)
    all_gather = import_module("collaborative-circom_b200.distributed").make_all_gather(world, torch.device("cuda", local))
    sess = cocg.ShamirSession(zk, 3, 1, seeds=PRF_SEEDS, rank=rank, world=world, all_gather=all_gather if world > 1 else None)
    ctx = cocg.Context(curve_id, local)
    v, c = ctx.upload(rand_fr(n_aux, rng)), ctx.upload(rand_fr(n_aux, rng))
    r1 = pow(2, 256, modulus)
    shares = []
    for p in range(3):  # degree-1 sharing: x_p = v + (p + 1) * c
        t = torch.empty(n_aux * 4, dtype=torch.int64).pin_memory()
        t.numpy().view(np.uint64).reshape(n_aux, 4)[:] = ctx.vec_axpy(limbs_of((p + 1) * r1 % modulus), c, v).to_host()
        shares.append(t)
    wit = [t.numpy().view(np.uint64).reshape(n_aux, 4) for t in shares]
    pub = np.stack([limbs_of(r1), limbs_of(12345 * r1 % modulus)])
    def leg(mode, warm, sample):
        sess.set_mpc_exchange(mode)
        for _ in range(warm):
            sess.prove(pub, wit)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local) if sample else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            proofs, _ = sess.prove(pub, wit)
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert np.array_equal(proofs[0], proofs[1]) and np.array_equal(proofs[1], proofs[2]), "the parties disagree on the proof"
        return float(t.item()), clocks, proofs

    # `value`: the MPC payloads of the three co-located parties handed over in HBM; `e2e`: staged through pinned host memory, as a
    # party that has to reach a NIC would.  Both legs upload the witness shares from pinned host memory every step.
    ms, clocks, proofs = leg("device", args.warmup, True)
    ms_e2e, _, _ = leg("host", min(args.warmup, 2), False)
    value = args.steps / (ms / 1e3)
    if rank == 0:
        import hashlib
        print(json.dumps({
            "metric": "groth16_proofs_per_sec", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 limbs (256/384-bit Montgomery integers; no floating point)", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "proof_sha256": hashlib.sha256(np.ascontiguousarray(proofs[0]).tobytes()).hexdigest(),
            "e2e": {"value": args.steps / (ms_e2e / 1e3), "unit": "proofs/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": 3 * n_aux * 32,
                    "d2h_bytes_per_step": 3 * 8 * (4 if args.curve == "bn254" else 6) * 8,
                    "note": "witness shares uploaded from pinned host memory every step, proofs read back, every Shamir network message (double-random "
                            "preprocessing, the king's reconstruct / re-share of both mul_vec rounds) staged through pinned host memory"},
        }))
    sess.close()
    zk.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log-n", type=int, default=None, help="default: 20 (groth16, msm), 18 (plonk)")
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="groth16", choices=["groth16", "msm", "plonk"],
                    help="groth16 = BASELINE configs[2] (the headline); msm = configs[1] (BN254 G1 MSM 2^20); plonk = configs[3] (Plonk 2^18 gates, REP3)")
    ap.add_argument("--curve", default="bn254", choices=["bn254", "bls12_381"], help="non-default: BLS12-381 (BASELINE configs[4] flavour)")
    ap.add_argument("--protocol", default="rep3", choices=["rep3", "shamir"], help="non-default: Shamir (3,1), single GPU only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard-mode", default="blocks", choices=["blocks", "ranges"],
                    help="N > 1, groth16 REP3: blocks = whole witness maps / MSM bundles per rank (default); ranges = every MSM cut into N index ranges")
    args = ap.parse_args()
    if args.log_n is None:
        args.log_n = 18 if args.workload == "plonk" else 20
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "msm":
        run_own_msm(args)
    elif args.workload == "plonk":
        run_own_plonk(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
