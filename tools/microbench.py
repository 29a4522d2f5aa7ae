#!/usr/bin/env python3
"""Per-kernel timings on one B200 (CUDA events on the context's stream).  Development aid; bench.py is the contract."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cocg  # noqa: E402
from oracle import cref, ntt as ontt  # noqa: E402
from oracle.curves import BN254  # noqa: E402


def rand_fr(n, seed):
    a = np.random.default_rng(seed).integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(min(ts))


def main():
    res = {}
    torch.cuda.init()
    ctx = cocg.Context(cocg.BN254, 0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    c = BN254
    # ---- element-wise, 2^24 elements = 512 MiB per vector (>> L2)
    n = 1 << 24
    a, b = ctx.upload(rand_fr(n, 1)), ctx.upload(rand_fr(n, 2))
    o = ctx.zeros(n)
    for name, op, nbytes in (("add", cocg.OP_ADD, 96), ("sub", cocg.OP_SUB, 96), ("mul", cocg.OP_MUL, 96), ("neg", cocg.OP_NEG, 64)):
        med, mn = timeit(lambda: ctx.vec_op(op, a, b, out=o))
        res["vec_" + name] = {"ms": med, "GBps": n * nbytes / med / 1e6, "Gelem_s": n / med / 1e6}
    med, mn = timeit(lambda: ctx.rep3_mul_local(a, b, b, a, None, out=o))
    res["rep3_mul_local"] = {"ms": med, "GBps": n * 160 / med / 1e6, "Gmul_s": 2 * n / med / 1e6}
    for v in (a, b, o):
        v.free()
    # ---- NTT 2^20, 2 components (and 2^22, 2^24 single)
    for logn, k in ((20, 2), (20, 1), (22, 1), (24, 1)):
        n = 1 << logn
        omega, g = ontt.groth16_roots(c, logn)
        om = cref.fr_to_mont(c, [omega])
        vs = [ctx.upload(rand_fr(n, 3 + i)) for i in range(k)]
        med, mn = timeit(lambda: ctx.ntt(vs, logn, om))
        res[f"ntt_2^{logn}_k{k}"] = {"ms": med, "GBps_algorithmic": k * n * 64 / med / 1e6}
        for v in vs:
            v.free()
    # ---- MSM G1 / G2
    for group, logn in ((1, 16), (1, 18), (1, 20), (2, 18), (2, 20)):
        n = 1 << logn
        p0 = cref.g_to_mont(c, [c.mul(c.gen(group), 12345, group)], group)
        q = cref.g_to_mont(c, [c.mul(c.gen(group), 6789, group)], group)
        t0 = time.time()
        pts = cref.gen_chain(c, group, p0[0], q[0], n)
        tgen = time.time() - t0
        h = ctx.bases_upload(group, pts)
        sc = ctx.upload(rand_fr(n, 9))
        pb = 64 * group + 32
        med, mn = timeit(lambda: ctx.msm(h, [sc]), iters=5, warm=2)
        res[f"msm_g{group}_2^{logn}"] = {"ms": med, "min_ms": mn, "GBps_algorithmic": n * pb / med / 1e6, "gen_s": tgen}
        ctx.bases_free(h)
        sc.free()
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/microbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
