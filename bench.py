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
    p = os.path.join(ROOT, "profiles", "r01_msm_accumulate_traffic.json")
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
def cpu_party_time(log_n, threads=None):
    """Times ONE REP3 party's share of one proof on the host cores with the C oracle (oracle/c/cocg_oracle.c: Pippenger with
    arkworks' window rule, radix-2 NTT, OpenMP): 2 SpMV pairs, 2 mul_vec local steps, 12 NTTs + 6 coset scalings, 1 sub,
    8 G1 + 2 G2 MSMs.  Returns (seconds, cores)."""
    from oracle import cref, ntt as ontt
    from oracle.curves import BN254 as C

    L = cref.lib()
    if threads:
        L.orc_set_threads(threads)
    cores = L.orc_num_threads()
    rng = np.random.default_rng(SEED)
    n = 1 << log_n
    n_public, n_vars, rows, A, B = synthetic_r1cs(log_n, rng)
    n_aux = n_vars - n_public - 1

    def chain(group, cnt, k):
        p0 = cref.g_to_mont(C, [C.mul(C.gen(group), 1000 + k, group)], group)[0]
        q = cref.g_to_mont(C, [C.mul(C.gen(group), 77 + k, group)], group)[0]
        return cref.gen_chain(C, group, p0, q, cnt)

    h_q, l_q, a_q, b1_q = chain(1, n, 1), chain(1, n_aux, 2), chain(1, n_aux, 3), chain(1, n_aux, 4)
    b2_q = chain(2, n_aux, 5)
    wa, wb = rand_fr(n_aux, rng), rand_fr(n_aux, rng)
    pub = rand_fr(n_public + 1, rng)
    omega, g = ontt.groth16_roots(C, log_n)
    om, omi = cref.fr_to_mont(C, [omega]), cref.fr_to_mont(C, [pow(omega, -1, C.r)])
    gm, one = cref.fr_to_mont(C, [g]), cref.fr_to_mont(C, [1])
    t0 = time.perf_counter()
    za, zb = np.concatenate([pub, wa]), np.concatenate([np.zeros_like(pub), wb])

    def rows_of(M, z):
        out = np.zeros((n, 4), dtype=np.uint64)
        out[:rows] = cref.spmv(C, M[0], M[1], M[2], z)
        return out

    a = [rows_of(A, za), rows_of(A, zb)]
    b = [rows_of(B, za), rows_of(B, zb)]
    c0 = cref.rep3_mul_local(C, a[0], a[1], b[0], b[1], None)
    c = [c0, c0.copy()]  # the received component has the same cost profile

    def coset(v):
        v = cref.ntt(C, v, omi, inverse=True)
        v = cref.distribute_powers(C, v, gm, one)
        return cref.ntt(C, v, om)

    a = [coset(v) for v in a]
    b = [coset(v) for v in b]
    ab0 = cref.rep3_mul_local(C, a[0], a[1], b[0], b[1], None)
    c = [coset(v) for v in c]
    h = [cref.fr_vec_op(C, cref.OP_SUB, ab0, c[0]), cref.fr_vec_op(C, cref.OP_SUB, ab0, c[1])]
    for comp in range(2):
        w = wa if comp == 0 else wb
        cref.msm(C, 1, h_q, h[comp])
        cref.msm(C, 1, l_q, w)
        cref.msm(C, 1, a_q, w)
        cref.msm(C, 1, b1_q, w)
        cref.msm(C, 2, b2_q, w)
    return time.perf_counter() - t0, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(min(args.warmup, 1)):
        cpu_party_time(min(args.log_n, 14))
    times = []
    budget_s = 150.0
    t_start = time.perf_counter()
    cores = 0
    for i in range(args.steps):
        t, cores = cpu_party_time(args.log_n)
        times.append(t)
        if time.perf_counter() - t_start + t > budget_s:
            break
    t = float(np.median(times))
    value = 1.0 / (3.0 * t)
    sample = (f"one of the three REP3 parties' share of one 2^{args.log_n} proof per step (2 SpMV pairs, 2 mul_vec local steps, 12 NTTs, "
              f"8 G1 + 2 G2 MSMs), proofs/s = 1 / (3 x median step time); {len(times)} timed steps")
    print(json.dumps({
        "impl": "reference", "metric": "groth16_proofs_per_sec", "value": value, "unit": "proofs/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": 3e3 * t, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64 limbs (256-bit Montgomery)", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": cores, "kind": "port", "sample": sample},
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
            "parallelism": "single GPU" if world == 1 else f"MSM bases sharded by index range over {world} GPUs, NTT replicated, 1 all-gather/proof",
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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
    zk = cocg.Groth16ZKey(curve_id, n_public, n_vars, log_n, rows, A, B, device=local, synthetic_seed=seed_bytes, rank=rank, world=world)
    sess = cocg.Rep3Session(zk, seeds=PRF_SEEDS, rank=rank, world=world)
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

    from importlib import import_module
    all_gather = import_module("collaborative-circom_b200.distributed").make_all_gather(world, torch.device("cuda", local))

    def step(device_resident):
        if device_resident:
            return sess.prove(pub, dev_a, dev_b, all_gather=all_gather, device_ptrs=True)
        return sess.prove(pub, host_a, host_b, all_gather=all_gather)

    def timed(device_resident, steps, sample_clocks=False):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            proofs = step(device_resident)
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
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

    # The same accumulate kernel timed ALONE (one context, one stream, no other party competing for the SMs): the in-situ scope times
    # above include the time slices the GPU gives to the other two parties' kernels.  One G1 query of this rank's shard size, k = 1.
    alone = None
    if world == 1:
        bits = 254 if args.curve == "bn254" else 255
        per_alone = n_aux
        h_alone = ctx.bases_generate(1, per_alone, bytes([77] * 32))
        ctx.msm(h_alone, [dev[0]], n=per_alone)
        ctx.profile(True)
        ctx.profile_reset()
        for _ in range(5):
            ctx.msm(h_alone, [dev[0]], n=per_alone)
        ctx.sync()
        pa = ctx.profile_read()
        ctx.profile(False)
        ctx.bases_free(h_alone)
        lg = (per_alone + per_alone // 2).bit_length() - 1
        c_bits = min(max(lg, 4), 20)
        nwin = (bits + c_bits) // c_bits
        alone = {"ms": pa["msm_accumulate"][0] / max(pa["msm_accumulate"][1], 1), "terms": per_alone, "window_bits": c_bits, "windows": nwin}

    value = args.steps / (ms / 1e3)
    e2e = args.steps / (ms_e2e / 1e3)
    peak, peak_src = measured_peak()
    # roofline of the dominant kernel: MSM bucket accumulation.  Algorithmic bytes per launch scope = one read of every point of
    # the rank's slice + one read of its scalar: G1 96 B/term, G2 160 B/term; one scope = one share component.
    per = (n_aux + world - 1) // world
    perh = (n + world - 1) // world
    scopes_per_proof = 3 * 2 * 5
    bytes_per_proof = 3 * 2 * (perh * 96 + 3 * per * 96 + per * 160)
    acc_ms, acc_n = prof["msm_accumulate"]
    avg_ms = acc_ms / max(acc_n, 1)
    achieved = (bytes_per_proof / scopes_per_proof) / (avg_ms * 1e-3) / 1e9 if acc_n else 0.0
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
                    "note": "witness shares in pinned host memory uploaded every step, proofs read back, and the two mul_vec rounds of each "
                            "party (n x 32 B out + in per round) staged through pinned host memory; the `value` leg keeps all of that in HBM"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(world),
                         "kernel": "msm_accumulate_kernel (+ msm_heavy_kernel), per share-component launch, timed in situ with the three "
                                   "parties' streams running concurrently", "peak_source": peak_src,
                         "note": "MSM is bound by 32-bit integer multiply-add issue, not HBM (DESIGN.md)"},
            "replicas": replicas,
            "kernels": kernels, "setup_s": round(setup_s, 2),
            "host_phases_ms": {"witness_map": round(float(phases[0]), 2), "msm": round(float(phases[1]), 2),
                               "all_gather_wait": round(float(phases[2]), 2), "assembly": round(float(phases[3]), 2)},
        }
        if alone:
            # issue roofline of the same kernel: 10 Fq multiplications (8M + 2S) per table point added; ceiling = 148 SMs x 32
            # IMAD.WIDE/clk x sm_max_mhz / 144 multiply-pipe instructions per 8-limb Montgomery product (SASS count)
            fq_mul = alone["terms"] * alone["windows"] * 10 / (alone["ms"] * 1e-3) / 1e9
            ceiling, ceiling_src = 148 * 32 * (clocks or {}).get("sm_max_mhz", 1965.0) * 1e6 / 144 / 1e9, "IMAD.WIDE issue ceiling (148 SMs x 32/clk / 144 per product)"
            cpath = os.path.join(ROOT, "profiles", "r01_fp_mul_ceiling.json")
            if os.path.exists(cpath):  # measured: a pure chain of the same Montgomery product on every SM
                ceiling = float(json.load(open(cpath))["G_mul_per_s"]["bn254_fq" if args.curve == "bn254" else "bls381_fq"])
                ceiling_src = "measured fp_mul chain, profiles/r01_fp_mul_ceiling.json"
            out["roofline"].update({
                "achieved_alone": alone["terms"] * 96 / (alone["ms"] * 1e-3) / 1e9, "frac_alone": alone["terms"] * 96 / (alone["ms"] * 1e-3) / 1e9 / peak,
                "alone_ms": alone["ms"],
                "issue": {"unit": "G Fq-mul/s", "achieved": fq_mul, "peak": ceiling, "frac": fq_mul / ceiling,
                          "note": f"G1 accumulate alone: {alone['terms']} terms x {alone['windows']} windows (c = {alone['window_bits']}) x 10 "
                                  f"Montgomery products; peak = {ceiling_src}"}})
        if world == 1 and not args.no_cpu_baseline:
            t, cores = cpu_party_time(log_n)
            out["cpu_baseline"] = {"value": 1.0 / (3.0 * t), "unit": "proofs/s", "cores": cores, "kind": "port",
                                   "sample": f"one of the three REP3 parties' share of one 2^{log_n} proof on the C oracle (OpenMP), {t:.1f} s; "
                                             "proofs/s = 1 / (3 x that)"}
        print(json.dumps(out))
    sess.close()
    zk.close()
    if world > 1:
        dist.destroy_process_group()


def run_own_shamir(args, cocg, torch, curve_id, modulus, local, r1cs, seed_bytes, rng):
    """Non-default configuration: CoGroth16<ShamirProtocol>, 3 parties, threshold 1, one GPU (BASELINE configs[4] flavour).  The
    double-random preprocessing of the two mul_vec rounds (shamir.rs:923-1010) is inside the timed region, as in the reference."""
    n_public, n_vars, rows, A, B = r1cs
    log_n = args.log_n
    n_aux = n_vars - n_public - 1
    zk = cocg.Groth16ZKey(curve_id, n_public, n_vars, log_n, rows, A, B, device=local, synthetic_seed=seed_bytes)
    sess = cocg.ShamirSession(zk, 3, 1, seeds=PRF_SEEDS)
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
    for _ in range(args.warmup):
        sess.prove(pub, wit)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        proofs, _ = sess.prove(pub, wit)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    assert np.array_equal(proofs[0], proofs[1]) and np.array_equal(proofs[1], proofs[2]), "the parties disagree on the proof"
    value = args.steps / (ms / 1e3)
    print(json.dumps({
        "metric": "groth16_proofs_per_sec", "value": value, "unit": "proofs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32 limbs (256/384-bit Montgomery integers; no floating point)", "data": "synthetic", "config": workload_config(args, 1),
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 3 * n_aux * 32, "d2h_bytes_per_step": 3 * 8 * (4 if args.curve == "bn254" else 6) * 8,
                "note": "witness shares uploaded from pinned host memory every step (this configuration has no device-resident leg)"},
    }))
    sess.close()
    zk.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--curve", default="bn254", choices=["bn254", "bls12_381"], help="non-default: BLS12-381 (BASELINE configs[4] flavour)")
    ap.add_argument("--protocol", default="rep3", choices=["rep3", "shamir"], help="non-default: Shamir (3,1), single GPU only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
