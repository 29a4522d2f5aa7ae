#!/usr/bin/env python3
"""Profiling aid: rank 0's share of a `world`-way sharded Groth16 REP3 proof on ONE GPU (the other ranks' partial sums are replaced by
copies of rank 0's, so the proof is meaningless but every kernel and its size are those of a real sharded run).
  python tools/shard_profile.py --world 8 --steps 2            (run under ncu for a launch list)"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import cocg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--log-n", type=int, default=20)
a = ap.parse_args()
rng = np.random.default_rng(bench.SEED)
n_public, n_vars, rows, A, B = bench.synthetic_r1cs(a.log_n, rng)
n_aux = n_vars - n_public - 1
zk = cocg.Groth16ZKey(cocg.BN254, n_public, n_vars, a.log_n, rows, A, B, synthetic_seed=bench.SEED.to_bytes(8, "little") * 4, rank=0, world=a.world)
sess = cocg.Rep3Session(zk, seeds=bench.PRF_SEEDS, rank=0, world=a.world)
sess.set_mpc_exchange("device")
ctx = cocg.Context(cocg.BN254, 0)
dev = [ctx.upload(bench.rand_fr(n_aux, rng)) for _ in range(3)]
da, db = [d.ptr for d in dev], [dev[(i - 1) % 3].ptr for i in range(3)]
r1 = pow(2, 256, bench.BN254_R)
pub = np.stack([bench.limbs_of(r1), bench.limbs_of(12345 * r1 % bench.BN254_R)])
gather = (lambda p: np.concatenate([p] * a.world)) if a.world > 1 else None
for _ in range(2):
    sess.prove(pub, da, db, all_gather=gather, device_ptrs=True)
t0 = time.perf_counter()
for _ in range(a.steps):
    sess.prove(pub, da, db, all_gather=gather, device_ptrs=True)
dt = (time.perf_counter() - t0) / a.steps
ph = sess.phase_times().max(axis=0) * 1e3
print(f"world {a.world}: {dt * 1e3:.2f} ms per proof (rank 0's share); phases ms: witness map {ph[0]:.2f}, msm {ph[1]:.2f}, gather wait {ph[2]:.2f}, assembly {ph[3]:.2f}")
