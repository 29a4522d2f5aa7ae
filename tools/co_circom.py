#!/usr/bin/env python3
"""The slice of the reference's `co-circom` command line that sits either side of the proving path, same
sub-commands and flag names as /root/reference/co-circom/co-circom/src/lib.rs:105-444 and src/bin/co-circom.rs:

  split-witness   --witness W.wtns --r1cs C.r1cs --protocol REP3|SHAMIR --curve BN254|BLS12-381 --out-dir DIR [-t T] [-n N]
                  -> DIR/<W>.<i>.shared                                   (co-circom.rs:160-256)
  generate-proof  groth16|plonk --witness S0.shared S1.shared S2.shared [...] --zkey K.zkey --protocol REP3|SHAMIR [-t T] --curve ...
                  --out proof.json [--public-input public.json]             (co-circom.rs:455-636; plonk: REP3)
  verify          groth16|plonk --proof proof.json --vk verification_key.json --public-input public.json --curve ...
                  exit code 0 = accepted, 1 = rejected                      (co-circom.rs:640-720; host pairing, no GPU)

Differences, by design: the reference runs ONE party per process and joins them over QUIC (mpc-net, out of scope); here
`generate-proof` takes the three parties' share files and runs them as three threads of one process on one B200, joined by the
in-process network the reference's own tests use (tests/tests/circom/e2e_tests/mod.rs:55-70).  All arithmetic and all file codecs
are in libcocg.so / libcohost.so; this file only parses flags.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cocg  # noqa: E402

CURVES = {"BN254": cocg.BN254, "BLS12-381": cocg.BLS12_381}


def split_witness(a):
    if not os.path.isdir(a.out_dir):
        sys.exit(f"directory {a.out_dir} does not exist")
    paths = cocg.split_witness_files(a.witness, a.r1cs, a.protocol, CURVES[a.curve], a.out_dir, a.threshold, a.num_parties)
    for i, p in enumerate(paths):
        print(f"Wrote witness share {i} to file {p}")


def generate_proof_plonk(a):
    """CoPlonk::prove over three REP3 parties or n Shamir parties (co-circom.rs:455-636 with proof system plonk)."""
    rep3 = a.protocol == "REP3"
    if rep3 and len(a.witness) != 3:
        sys.exit("REP3 needs the three parties' share files (one process plays all three parties)")
    if not rep3 and len(a.witness) <= 2 * a.threshold:
        sys.exit("Shamir needs the share files of all n > 2t parties (one process plays all of them)")
    curve = CURVES[a.curve]
    zk = cocg.PlonkZKey(a.zkey)
    pubs, wa, wb = [], [], []
    for path in a.witness:
        pub, comps = cocg.shared_witness_decode(curve, open(path, "rb").read(), 2 if rep3 else 1)
        pubs.append(pub)
        wa.append(comps[0])
        wb.append(comps[-1])
    if any(p.shape != pubs[0].shape or not (p == pubs[0]).all() for p in pubs[1:]):
        sys.exit("the share files disagree on the public inputs")
    # a SharedWitness of the Groth16 flow carries every signal after the public ones; the Plonk prover reads the first
    # n_vars - n_additions - n_public - 1 of them (the additions are recomputed, round1.rs:213-242)
    need = zk.n_witness
    if any(x.shape[0] < need for x in wa + wb):
        sys.exit("the share files hold fewer witness elements than the zkey expects")
    if rep3:
        sess = cocg.PlonkSession(zk, "rep3", seeds=os.urandom(96))
        proofs = sess.prove(pubs[0], [x[:need] for x in wa], [x[:need] for x in wb])
    else:
        sess = cocg.PlonkSession(zk, "shamir", seeds=os.urandom(32 * len(a.witness)), num_parties=len(a.witness), threshold=a.threshold)
        proofs = sess.prove(pubs[0], [x[:need] for x in wa])
    if any(not (p == proofs[0]).all() for p in proofs[1:]):
        sys.exit("the parties opened different proofs")
    with open(a.out, "w") as f:
        f.write(cocg.plonk_proof_to_json(curve, proofs[0]))
    print(f"Wrote proof to file {a.out}")
    if a.public_input:
        with open(a.public_input, "w") as f:
            f.write(cocg.public_inputs_to_json(curve, pubs[0]))
        print(f"Wrote public inputs to file {a.public_input}")
    sess.close()
    zk.close()


def generate_proof(a):
    if a.proof_system == "plonk":
        return generate_proof_plonk(a)
    rep3 = a.protocol == "REP3"
    if rep3 and len(a.witness) != 3:
        sys.exit("REP3 needs the three parties' share files (one process plays all three parties)")
    if not rep3 and len(a.witness) <= 2 * a.threshold:
        sys.exit("Shamir needs the share files of all n > 2t parties (one process plays all of them)")
    curve = CURVES[a.curve]
    zk = cocg.Groth16ZKey.from_file(a.zkey)
    pubs, wa, wb = [], [], []
    for path in a.witness:
        pub, comps = cocg.shared_witness_decode(curve, open(path, "rb").read(), 2 if rep3 else 1)
        pubs.append(pub)
        wa.append(comps[0])
        wb.append(comps[-1])
    if any(p.shape != pubs[0].shape or not (p == pubs[0]).all() for p in pubs[1:]):
        sys.exit("the share files disagree on the public inputs")
    if rep3:
        sess = cocg.Rep3Session(zk, seeds=os.urandom(96))
        proofs = sess.prove(pubs[0], wa, wb)
    else:
        sess = cocg.ShamirSession(zk, len(a.witness), a.threshold, seeds=os.urandom(32 * len(a.witness)))
        proofs, _ = sess.prove(pubs[0], wa)
    if any(not (p == proofs[0]).all() for p in proofs[1:]):
        sys.exit("the parties opened different proofs")
    with open(a.out, "w") as f:
        f.write(cocg.proof_to_json(curve, proofs[0]))
    print(f"Wrote proof to file {a.out}")
    if a.public_input:
        with open(a.public_input, "w") as f:
            f.write(cocg.public_inputs_to_json(curve, pubs[0]))
        print(f"Wrote public inputs to file {a.public_input}")
    sess.close()
    zk.close()


def verify(a):
    vk, proof, pub = (open(p).read() for p in (a.vk, a.proof, a.public_input))
    import json
    if json.loads(vk).get("curve") != {"BN254": "bn128", "BLS12-381": "bls12381"}[a.curve]:
        sys.exit("the verification key is over a different curve")
    check = cocg.groth16_verify_json if a.proof_system == "groth16" else cocg.plonk_verify_json
    if check(vk, proof, pub):
        print("Proof verified successfully")
        return
    print("Proof verification failed")
    sys.exit(1)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="co-circom")
    sub = ap.add_subparsers(dest="cmd", required=True)
    sw = sub.add_parser("split-witness")
    sw.add_argument("--witness", required=True)
    sw.add_argument("--r1cs", required=True)
    sw.add_argument("--protocol", required=True, choices=["REP3", "SHAMIR"])
    sw.add_argument("--curve", required=True, choices=list(CURVES))
    sw.add_argument("--out-dir", required=True)
    sw.add_argument("-t", "--threshold", type=int, default=1)
    sw.add_argument("-n", "--num-parties", type=int, default=3)
    sw.set_defaults(fn=split_witness)
    gp = sub.add_parser("generate-proof")
    gp.add_argument("proof_system", choices=["groth16", "plonk"])
    gp.add_argument("--witness", required=True, nargs="+")
    gp.add_argument("--zkey", required=True)
    gp.add_argument("--protocol", required=True, choices=["REP3", "SHAMIR"])
    gp.add_argument("--curve", required=True, choices=list(CURVES))
    gp.add_argument("--out", required=True)
    gp.add_argument("--public-input")
    gp.add_argument("-t", "--threshold", type=int, default=1)
    gp.set_defaults(fn=generate_proof)
    vf = sub.add_parser("verify")
    vf.add_argument("proof_system", choices=["groth16", "plonk"])
    vf.add_argument("--proof", required=True)
    vf.add_argument("--vk", required=True)
    vf.add_argument("--public-input", required=True)
    vf.add_argument("--curve", required=True, choices=list(CURVES))
    vf.set_defaults(fn=verify)
    a = ap.parse_args(argv)
    try:
        a.fn(a)
    except cocg.CocgError as e:
        sys.exit(str(e))


if __name__ == "__main__":
    main()
