#!/usr/bin/env python3
"""Per-kernel-class timings on one B200, one context, one stream (cocg_profile_*: CUDA events on the launching stream).
Development aid; bench.py is the contract."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cocg  # noqa: E402

R1 = pow(2, 256, 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001)


def rand_fr(n, seed):
    a = np.random.default_rng(seed).integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def prof(ctx, fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ctx.profile(True)
    ctx.profile_reset()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    ctx.sync()
    wall = (time.perf_counter() - t0) / iters * 1e3
    out = {k: round(v[0] / iters, 4) for k, v in ctx.profile_read().items() if v[1]}
    ctx.profile(False)
    out["wall_ms"] = round(wall, 4)
    return out


def main():
    res = {}
    ctx = cocg.Context(cocg.BN254, 0)
    sizes = [int(x) for x in sys.argv[1:]] or [20]
    for logn in sizes:
        n = 1 << logn
        for group in (1, 2):
            h = ctx.bases_generate(group, n, bytes([group] * 32))
            sc = [ctx.upload(rand_fr(n, 9)), ctx.upload(rand_fr(n, 10))]
            r = prof(ctx, lambda: ctx.msm(h, sc[:1]))
            r["alg_GBps"] = round(n * (64 * group + 32) / r["wall_ms"] / 1e6, 2)
            res[f"msm_g{group}_2^{logn}_k1"] = r
            res[f"msm_g{group}_2^{logn}_k2"] = prof(ctx, lambda: ctx.msm(h, sc))
            ctx.bases_free(h)
            for v in sc:
                v.free()
        # the Groth16 aux shape: 3 G1 queries + 1 G2 query x 2 share components, one call (shared sorts, batched reductions)
        hs = [ctx.bases_generate(1, n, bytes([10 + i] * 32)) for i in range(3)] + [ctx.bases_generate(2, n, bytes([13] * 32))]
        sc = [ctx.upload(rand_fr(n, 9)), ctx.upload(rand_fr(n, 10))]
        res[f"msm_multi_3g1_1g2_2^{logn}_k2"] = prof(ctx, lambda: ctx.msm_multi(hs, [0, 0, 0, 0], sc))
        for h in hs:
            ctx.bases_free(h)
        for v in sc:
            v.free()
        # NTT: root of unity is irrelevant for timing; use any element
        vs = [ctx.upload(rand_fr(n, 3 + i)) for i in range(2)]
        om = np.array([[R1 & (2**64 - 1), (R1 >> 64) & (2**64 - 1), (R1 >> 128) & (2**64 - 1), R1 >> 192]], dtype=np.uint64)
        r = prof(ctx, lambda: ctx.ntt(vs, logn, om))
        r["alg_GBps"] = round(2 * n * 64 / r["ntt"] / 1e6, 1)
        res[f"ntt_2^{logn}_k2"] = r
        o = ctx.zeros(n)
        r = prof(ctx, lambda: ctx.vec_op(cocg.OP_SUB, vs[0], vs[1], out=o))
        r["alg_GBps"] = round(n * 96 / r["vec"] / 1e6, 1)
        res[f"vec_sub_2^{logn}"] = r
        r = prof(ctx, lambda: ctx.rep3_mul_local(vs[0], vs[1], vs[1], vs[0], None, out=o))
        r["alg_GBps"] = round(n * 160 / r["vec"] / 1e6, 1)
        res[f"rep3_mul_local_2^{logn}"] = r
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/microbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
