"""Round 1 of the Plonk prover through the C++ host layer (host/plonk.hpp) against the reference's bit-exact KATs
(co-plonk/src/round1.rs:344-427: PlainDriver, deterministic blinders), and the same commitments from three REP3 parties."""
import json
import os
import random

import numpy as np
import pytest

from oracle import cref, formats, groth16
from oracle.curves import BN254, BLS12_381

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("curve,circ", [("bn254", "multiplier2"), ("bls12_381", "poseidon")])
def test_round1_kat_plain_and_rep3(cocg, curve, circ):
    d = os.path.join(G, "plonk", curve, circ)
    kat = json.load(open(os.path.join(G, "plonk_round1_kats.json")))[curve + "/" + circ]
    want = [(int(kat[k][0]), int(kat[k][1])) for k in ("commit_a", "commit_b", "commit_c")]
    zk = cocg.PlonkZKey(os.path.join(d, "circuit.round1.zkey"))
    ozk = formats.parse_plonk_zkey(open(os.path.join(d, "circuit.round1.zkey"), "rb").read())
    assert (zk.n_vars, zk.n_public, zk.domain_size, zk.n_additions, zk.n_constraints) == \
        (ozk.n_vars, ozk.n_public, ozk.domain_size, ozk.n_additions, ozk.n_constraints)
    c = ozk.curve
    _, wt = formats.parse_wtns(open(os.path.join(d, "witness.wtns"), "rb").read())
    ell = zk.n_public
    pub = cref.fr_to_mont(c, wt[:ell + 1])
    wit = [v % c.r for v in wt[ell + 1:]]
    got = zk.round1_plain(pub, cref.fr_to_mont(c, wit))
    assert [cref.g_from_mont(c, g, 1)[0] for g in got] == want
    # three REP3 parties, trivial shares of the deterministic blinders: every party opens the same literal commitments
    rng = random.Random(3)
    shares = groth16.share_rep3(wit, rng, c.r)
    out = zk.round1_rep3(pub, [cref.fr_to_mont(c, s[0]) for s in shares], [cref.fr_to_mont(c, s[1]) for s in shares])
    for party in range(3):
        assert [cref.g_from_mont(c, g, 1)[0] for g in out[party]] == want
    # random blinders: still agree between parties, differ from the KAT
    out2 = zk.round1_rep3(pub, [cref.fr_to_mont(c, s[0]) for s in shares], [cref.fr_to_mont(c, s[1]) for s in shares], deterministic=False)
    assert np.array_equal(out2[0], out2[1]) and np.array_equal(out2[1], out2[2]) and not np.array_equal(out2[0], out[0])
    zk.close()
