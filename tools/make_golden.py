#!/usr/bin/env python3
"""Regenerate tests/golden/ from the reference checkout (run in the build container, where /root/reference exists).

Two kinds of artefacts are produced:
  * data fixtures copied verbatim from the reference's test_vectors/ (zkey, wtns, snarkjs proofs, verification keys) --
    these are inputs/outputs of snarkjs, not reference source code;
  * known-answer literals lifted out of the reference's own unit tests into JSON:
      - circom-types/src/groth16/zkey.rs:335-585  (every point of the multiplier2 zkeys, both curves)
      - mpc-core/tests/protocols/rep3.rs:242-350  (rep3_mul_vec_bn: x, y, x*y)
The GPU box has no /root/reference; tests read only tests/golden/.
"""
import json
import os
import re
import shutil

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def copy_fixtures():
    for curve in ("bn254", "bls12_381"):
        for circ in ("multiplier2", "poseidon"):
            src = os.path.join(REF, "test_vectors", "Groth16", curve, circ)
            dst = os.path.join(OUT, "groth16", curve, circ)
            os.makedirs(dst, exist_ok=True)
            for f in ("circuit.zkey", "witness.wtns", "circom.proof", "public.json", "verification_key.json", "circuit.r1cs"):
                shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
                os.chmod(os.path.join(dst, f), 0o644)


def _fn_body(text, name):
    i = text.index("fn %s()" % name)
    j = text.index("\n    }\n", i)
    return text[i:j]


_ITEM = re.compile(
    r'to_g1_\w+!\(\s*"(\d+)"\s*,\s*"(\d+)"\s*\)'
    r'|to_g2_\w+!\(\s*\{\s*"(\d+)"\s*,\s*"(\d+)"\s*\}\s*,\s*\{\s*"(\d+)"\s*,\s*"(\d+)"\s*\}\s*\)'
    r'|(G1Affine::identity\(\))|(G2Affine::identity\(\))')


def zkey_kats():
    text = open(os.path.join(REF, "co-circom/circom-types/src/groth16/zkey.rs")).read()
    out = {}
    for name, key in (("can_deser_bls12_381_mult2_key", "bls12_381"), ("can_deser_bn254_mult2_key", "bn254")):
        body = _fn_body(text, name)
        kat = {}
        for m in re.finditer(r"let (\w+) = (.*?);\n", body, flags=re.S):
            var, val = m.group(1), m.group(2)
            items = []
            for it in _ITEM.finditer(val):
                if it.group(1):
                    items.append([it.group(1), it.group(2)])
                elif it.group(3):
                    items.append([[it.group(3), it.group(4)], [it.group(5), it.group(6)]])
                else:
                    items.append(None)
            if items:
                kat[var] = items if val.lstrip().startswith("vec!") else items[0]
        out[key] = kat
    return out


def rep3_mul_kat():
    text = open(os.path.join(REF, "mpc-core/tests/protocols/rep3.rs")).read()
    body = _fn_body(text, "rep3_mul_vec_bn").replace("\n    }\n", "")
    i = text.index("fn rep3_mul_vec_bn()")
    body = text[i:text.index("let mut x_shares1", i)]
    nums = re.findall(r'from_str\(\s*"(\d+)"', body)
    assert len(nums) == 12
    return {"x": nums[0:4], "y": nums[4:8], "xy": nums[8:12]}


def plonk_round1():
    """co-plonk/src/round1.rs:344-427: fixtures + the literal commitments.  The 6.4 MB poseidon zkey is trimmed to the sections
    round 1 reads (header, additions, wire maps, p_tau) -- still a valid binfile."""
    import struct
    text = open(os.path.join(REF, "co-circom/co-plonk/src/round1.rs")).read()
    out = {}
    for fn, curve, circ in (("test_round1_multiplier2", "bn254", "multiplier2"), ("test_round1_poseidon_bls12_381", "bls12_381", "poseidon")):
        i = text.index("fn %s()" % fn)
        j = text.index("\n    }\n", i)
        pts = re.findall(r'from_xy!\(\s*"(\d+)",\s*"(\d+)"\s*\)', text[i:j])
        assert len(pts) == 3
        out[curve + "/" + circ] = {"commit_a": pts[0], "commit_b": pts[1], "commit_c": pts[2]}
        src = os.path.join(REF, "test_vectors", "Plonk", curve, circ)
        dst = os.path.join(OUT, "plonk", curve, circ)
        os.makedirs(dst, exist_ok=True)
        shutil.copyfile(os.path.join(src, "witness.wtns"), os.path.join(dst, "witness.wtns"))
        data = open(os.path.join(src, "circuit.zkey"), "rb").read()
        nsec = struct.unpack_from("<I", data, 8)[0]
        off, keep = 12, []
        for _ in range(nsec):
            sid, slen = struct.unpack_from("<IQ", data, off)
            if sid in (1, 2, 3, 4, 5, 6, 14):
                keep.append(data[off:off + 12 + slen])
            off += 12 + slen
        with open(os.path.join(dst, "circuit.round1.zkey"), "wb") as f:
            f.write(data[:8] + struct.pack("<I", len(keep)) + b"".join(keep))
        for f in ("witness.wtns", "circuit.round1.zkey"):
            os.chmod(os.path.join(dst, f), 0o644)
    return out


def plonk_round2():
    """Bit-exact round-2 material of the reference: commit_z of test_round2_multiplier2 (co-plonk/src/round2.rs:326-355, blinders
    b_i = i), the transcript KAT (types.rs:190-226) and the verifier's beta / gamma / alpha / xi for the shipped snarkjs proof
    (plonk.rs:285-330).  Copies the full multiplier2 Plonk zkey (14.7 KB: sigma evaluations and vk points are needed from round 2 on),
    its snarkjs proof and public inputs."""
    out = {}
    text = open(os.path.join(REF, "co-circom/co-plonk/src/round2.rs")).read()
    i = text.index("fn test_round2_multiplier2()")
    pts = re.findall(r'from_xy!\(\s*"(\d+)",\s*"(\d+)"\s*\)', text[i:])
    out["commit_z"] = pts[0]
    text = open(os.path.join(REF, "co-circom/co-plonk/src/types.rs")).read()
    i = text.index("fn test_keccak_transcript()")
    out["transcript"] = {"points": re.findall(r'to_g1_bn254!\(\s*"(\d+)",\s*"(\d+)"\s*\)', text[i:]),
                         "scalars_and_challenge": re.findall(r'from_str\(\s*"(\d+)",?\s*\)', text[i:])}
    text = open(os.path.join(REF, "co-circom/co-plonk/src/plonk.rs")).read()
    i = text.index("fn calculate_verifier_challenges()")
    vals = re.findall(r'challenges\.(\w+),\s*ark_bn254::Fr::from_str\(\s*"(\d+)"', text[i:])
    out["verifier_challenges"] = {k: v for k, v in vals}
    src = os.path.join(REF, "test_vectors", "Plonk", "bn254", "multiplier2")
    dst = os.path.join(OUT, "plonk", "bn254", "multiplier2")
    for f in ("circuit.zkey", "circom.proof", "public.json", "circuit.r1cs"):
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
        os.chmod(os.path.join(dst, f), 0o644)
    # snarkjs Plonk proofs + keys of all four fixtures: the known-answer test of the verifier (co-plonk/src/lib.rs:255-275)
    for curve in ("bn254", "bls12_381"):
        for circ in ("multiplier2", "poseidon"):
            src = os.path.join(REF, "test_vectors", "Plonk", curve, circ)
            dst = os.path.join(OUT, "plonk", curve, circ)
            os.makedirs(dst, exist_ok=True)
            for f in ("circom.proof", "public.json", "verification_key.json"):
                shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
                os.chmod(os.path.join(dst, f), 0o644)
    # rounds 3-5 (round3.rs:553-596, round4.rs:181-247, round5.rs:391-429): [t1..3]_1, the six evaluations, [Wxi]_1, [Wxiw]_1
    t3 = open(os.path.join(REF, "co-circom/co-plonk/src/round3.rs")).read()
    out["commit_t"] = re.findall(r'g1_from_xy!\(\s*"(\d+)",\s*"(\d+)"\s*\)', t3[t3.index("fn test_round3_multiplier2"):])[:3]
    t4 = open(os.path.join(REF, "co-circom/co-plonk/src/round4.rs")).read()
    out["evals"] = dict(re.findall(r'round5\.proof\.(eval_\w+),\s*ark_bn254::Fr::from_str\(\s*"(\d+)"', t4[t4.index("fn test_round4_multiplier2"):]))
    t5 = open(os.path.join(REF, "co-circom/co-plonk/src/round5.rs")).read()
    out["commit_w"] = re.findall(r'g1_from_xy!\(\s*"(\d+)",\s*"(\d+)"\s*\)', t5[t5.index("fn test_round5_multiplier2"):])[:2]
    src = os.path.join(REF, "test_vectors", "Plonk", "bls12_381", "multiplier2")
    dst = os.path.join(OUT, "plonk", "bls12_381", "multiplier2")
    os.makedirs(dst, exist_ok=True)
    for f in ("circuit.zkey", "witness.wtns"):
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
        os.chmod(os.path.join(dst, f), 0o644)
    v = re.search(r'challenges\.v\.to_vec\(\),\s*vec!\[(.*?)\]', text[i:], flags=re.S)
    if v:
        out["verifier_challenges"]["v"] = re.findall(r'"(\d+)"', v.group(1))
    return out


def plonk_poseidon_full():
    """The full BN254 poseidon Plonk zkey (n = 4096, 2228 additions, 6.3 MB; xz-compressed here) + witness: the mid-size VALID circuit
    the GPU tests of rounds 2-5 prove and verify (co-plonk/src/lib.rs:232-275 test_poseidon_bn254)."""
    import lzma
    src = os.path.join(REF, "test_vectors", "Plonk", "bn254", "poseidon")
    dst = os.path.join(OUT, "plonk", "bn254", "poseidon")
    os.makedirs(dst, exist_ok=True)
    with open(os.path.join(src, "circuit.zkey"), "rb") as f, lzma.open(os.path.join(dst, "circuit.zkey.xz"), "wb", preset=9) as g:
        g.write(f.read())
    shutil.copyfile(os.path.join(src, "witness.wtns"), os.path.join(dst, "witness.wtns"))
    os.chmod(os.path.join(dst, "witness.wtns"), 0o644)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    copy_fixtures()
    plonk_poseidon_full()
    with open(os.path.join(OUT, "plonk_round1_kats.json"), "w") as f:
        json.dump(plonk_round1(), f, indent=1)
    with open(os.path.join(OUT, "plonk_round2_kats.json"), "w") as f:
        json.dump(plonk_round2(), f, indent=1)
    with open(os.path.join(OUT, "zkey_kats.json"), "w") as f:
        json.dump(zkey_kats(), f, indent=1)
    with open(os.path.join(OUT, "rep3_mul_vec_bn.json"), "w") as f:
        json.dump(rep3_mul_kat(), f, indent=1)
    print("wrote", OUT)
