#!/usr/bin/env python3
"""Short workloads for `ncu --set full -k regex:<kernel>` captures (B200_PROFILING.md): each mode launches the named kernels a few
times at the benchmark's size and nothing else heavy.
  python tools/ncu_targets.py msm_g1 | msm_g2 | msm_multi | ntt | plonk"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import cocg  # noqa: E402

mode = sys.argv[1]
rng = np.random.default_rng(bench.SEED)
if mode in ("msm_g1", "msm_g2"):
    ctx = cocg.Context(cocg.BN254, 0)
    n = 1 << 20
    h = ctx.bases_generate(1 if mode == "msm_g1" else 2, n, bytes([9] * 32))
    s = ctx.upload(bench.rand_fr(n, rng))
    for _ in range(3):
        ctx.msm(h, [s])
elif mode == "msm_multi":  # the Groth16 aux shape: 3 G1 queries + 1 G2 query x 2 share components, one call
    ctx = cocg.Context(cocg.BN254, 0)
    n = 1 << 20
    hs = [ctx.bases_generate(1, n, bytes([10 + i] * 32)) for i in range(3)] + [ctx.bases_generate(2, n, bytes([13] * 32))]
    sc = [ctx.upload(bench.rand_fr(n, rng)) for _ in range(2)]
    for _ in range(2):
        ctx.msm_multi(hs, [0, 0, 0, 0], sc)
elif mode == "ntt":
    from oracle import ntt as ontt, cref
    from oracle.curves import BN254
    ctx = cocg.Context(cocg.BN254, 0)
    omega, g = ontt.groth16_roots(BN254, 20)
    v = [ctx.upload(bench.rand_fr(1 << 20, rng)) for _ in range(2)]
    for _ in range(3):
        ctx.ntt(v, 20, cref.fr_to_mont(BN254, [omega]))
elif mode == "plonk":
    log_n = 18
    n_public, n_vars, n_constraints, maps = bench.plonk_synthetic_maps(log_n, rng)
    zk = cocg.PlonkZKey.synthetic(cocg.BN254, log_n, n_public, n_vars, maps, bytes(range(32)))
    sess = cocg.PlonkSession(zk, "rep3", seeds=bench.PRF_SEEDS)
    sess.set_mpc_exchange("device")
    x = [bench.rand_fr(zk.n_witness, rng) for _ in range(3)]
    r1 = pow(2, 256, bench.BN254_R)
    pub = np.stack([bench.limbs_of(r1), bench.limbs_of(12345 * r1 % bench.BN254_R)])
    sess.prove(pub, x, [x[2], x[0], x[1]])
else:
    raise SystemExit("unknown mode")
