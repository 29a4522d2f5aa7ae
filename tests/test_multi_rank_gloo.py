"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path -- index-range sharding of an MSM and the single
all-gather + fold per proof (collaborative-circom_b200/distributed.py).  The group sum is emulated with the C oracle's MSM on
each rank's slice (the GPU kernels need a device; their sharded result is checked in tests/test_gpu_groth16.py)."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module
    d = import_module("collaborative-circom_b200.distributed")
    from oracle import cref
    from oracle.curves import BN254 as C
    n = 3001
    rng = np.random.default_rng(5)
    sc = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    sc[:, 3] &= np.uint64((1 << 60) - 1)
    p0 = cref.g_to_mont(C, [C.mul(C.gen(1), 11, 1)], 1)[0]
    qq = cref.g_to_mont(C, [C.mul(C.gen(1), 13, 1)], 1)[0]
    pts = cref.gen_chain(C, 1, p0, qq, n)
    off, ln = d.shard_range(n, rank, world)
    partial = cref.msm(C, 1, pts[off:off + ln], sc[off:off + ln])           # this rank's partial sum (Jacobian limbs)
    gathered = d.make_all_gather(world)(partial).reshape(world, -1)          # ONE collective
    acc = gathered[0]
    for r in range(1, world):
        acc = cref.ec_op(C, 1, 0, acc, gathered[r])
    want = cref.msm(C, 1, pts, sc)
    ok = cref.jac_from_mont(C, acc, 1) == cref.jac_from_mont(C, want, 1) and np.array_equal(gathered[rank], partial)
    q.put((rank, bool(ok), (off, ln)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_msm_all_gather_fold_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2][0] == 0 and res[0][2][0] + res[0][2][1] == res[1][2][0] and res[1][2][0] + res[1][2][1] == 3001
