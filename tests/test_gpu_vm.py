"""Batched witness-extension arithmetic (host/vm.hpp): the field opcodes of the reference's MPC VM over a batch of independent inputs,
against the semantics of CircomWitnessExtensionProtocol for the plain and REP3 drivers
(/root/reference/mpc-core/src/protocols/plain.rs:421-445, rep3/witness_extension_impl.rs:81-200; dispatch
/root/reference/co-circom/circom-mpc-vm/src/mpc_vm.rs:508-546).  Values are checked on Python ints, per instance."""
import random

import numpy as np
import pytest

from oracle import cref, groth16
from oracle.curves import BN254, BLS12_381

pytestmark = pytest.mark.gpu


def _program():
    """A Poseidon-style round on registers: r0 = x (shared), r1 = k (public round constant), r2 = y (shared), r3 = c (public);
    t = (x + k)^5 (three shared multiplications); u = t * c - y; v = u / (y + c) ; w = -(v * v) + k / c ; z = c * k (public only)."""
    A, S, M, N, D = range(5)
    return [
        (A, 4, 0, 1),    # r4 = x + k            shared + public
        (M, 5, 4, 4),    # r5 = r4^2             shared * shared
        (M, 5, 5, 5),    # r5 = r4^4
        (M, 5, 5, 4),    # r5 = r4^5
        (M, 6, 5, 3),    # r6 = t * c            shared * public
        (S, 6, 6, 2),    # r6 = t c - y          shared - shared
        (A, 7, 2, 3),    # r7 = y + c
        (D, 8, 6, 7),    # r8 = u / (y + c)      shared / shared
        (M, 9, 8, 8),    # r9 = v^2
        (N, 9, 9, 0),    # r9 = -v^2
        (D, 10, 1, 3),   # r10 = k / c           public / public
        (A, 9, 9, 10),   # r9 = -v^2 + k / c     shared + public
        (M, 11, 3, 1),   # r11 = c * k           public * public
        (S, 12, 1, 0),   # r12 = k - x           public - shared
        (D, 13, 1, 2),   # r13 = k / y           public / shared
        (D, 14, 0, 3),   # r14 = x / c           shared / public
    ]


def _expected(r, x, k, y, c):
    t = pow((x + k) % r, 5, r)
    u = (t * c - y) % r
    v = u * pow((y + c) % r, -1, r) % r
    return {9: (-(v * v) + k * pow(c, -1, r)) % r, 11: c * k % r, 12: (k - x) % r, 13: k * pow(y, -1, r) % r, 14: x * pow(c, -1, r) % r, 8: v}


@pytest.mark.parametrize("curve", [BN254, BLS12_381], ids=["bn254", "bls12_381"])
@pytest.mark.parametrize("batch", [1, 1000, (1 << 16) + 7])
def test_batched_vm_plain_and_rep3(cocg, curve, batch):
    r = curve.r
    rng = random.Random(batch)
    cid = cocg.BN254 if curve is BN254 else cocg.BLS12_381
    x, k, y, c = ([rng.randrange(1, r) for _ in range(batch)] for _ in range(4))
    want = [_expected(r, *vals) for vals in zip(x, k, y, c)]
    f = lambda vals: cref.fr_to_mont(curve, vals)
    # plain driver
    vm = cocg.BatchedVm(cid, "plain", batch, 16, seeds=bytes(32))
    vm.set_shared(0, 0, f(x)); vm.set_public(1, f(k)); vm.set_shared(2, 0, f(y)); vm.set_public(3, f(c))
    vm.run(_program())
    for reg in (8, 9, 11, 12, 13, 14):
        got = vm.get(reg)
        assert got[0] == ("public" if reg == 11 else "shared")
        assert cref.fr_from_mont(curve, got[1]) == [w[reg] for w in want], reg
    vm.close()
    # three REP3 parties: the sum of the three a components is the value, b of party i is a of party i - 1
    vm = cocg.BatchedVm(cid, "rep3", batch, 16, seeds=bytes(range(96)))
    sx, sy = groth16.share_rep3(x, rng, r), groth16.share_rep3(y, rng, r)
    for p in range(3):
        vm.set_shared(0, p, f(sx[p][0]), f(sx[p][1]))
        vm.set_shared(2, p, f(sy[p][0]), f(sy[p][1]))
    vm.set_public(1, f(k)); vm.set_public(3, f(c))
    vm.run(_program())
    for reg in (8, 9, 12, 13, 14):
        parts = [vm.get(reg, p) for p in range(3)]
        assert all(q[0] == "shared" for q in parts)
        a = [cref.fr_from_mont(curve, q[1]) for q in parts]
        b = [cref.fr_from_mont(curve, q[2]) for q in parts]
        assert [(u + v + w) % r for u, v, w in zip(*a)] == [w[reg] for w in want], reg
        for p in range(3):
            assert b[p] == a[(p - 1) % 3]
    kind, z = vm.get(11, 1)
    assert kind == "public" and cref.fr_from_mont(curve, z) == [w[11] for w in want]
    st = vm.stats()
    # 5 shared multiplications + 2 shared inversions (an inversion is rand, mul-open): one network round each for the WHOLE batch
    assert st["network_rounds"] == 7 and st["launches"] > 0
    vm.close()


def test_batched_vm_division_by_zero_is_an_error(cocg):
    c = BN254
    vm = cocg.BatchedVm(cocg.BN254, "plain", 4, 4, seeds=bytes(32))
    vm.set_public(0, cref.fr_to_mont(c, [1, 2, 3, 4]))
    vm.set_public(1, cref.fr_to_mont(c, [5, 0, 7, 8]))
    with pytest.raises(cocg.CocgError, match="Cannot invert zero"):   # witness_extension_impl.rs:186-188
        vm.run([(4, 2, 0, 1)])
    vm.close()
