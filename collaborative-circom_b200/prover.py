"""ctypes binding of libcohost.so (include/cohost.h): device-resident Groth16 zkeys and proving sessions.

The host layer itself is C++ (collaborative-circom_b200/host): CoGroth16<T>::prove over PlainDriver / three Rep3Protocol
drivers, mirroring /root/reference/co-circom/co-groth16/src/groth16.rs and mpc-core/src/protocols/{plain,rep3}.rs.  This
file only marshals numpy arrays (uint64 limbs, Montgomery form) across the C boundary; nothing is computed in Python.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib
from ._lib import CocgError

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libcohost.so")
HOST_SYMBOLS = [
    "cohost_last_error", "cohost_zkey_create", "cohost_zkey_destroy", "cohost_plain_session_create",
    "cohost_plain_session_destroy", "cohost_plain_prove", "cohost_rep3_session_create", "cohost_rep3_session_destroy",
    "cohost_rep3_prove_begin", "cohost_rep3_partial_bytes", "cohost_rep3_prove_partials", "cohost_rep3_prove_combine",
    "cohost_rep3_prove_end", "cohost_rep3_launch_count", "cohost_rep3_prove_begin_device", "cohost_rep3_profile_enable",
    "cohost_rep3_profile_read", "cohost_rep3_profile_reset", "cohost_msm_shard_range", "cohost_shamir_session_create",
    "cohost_shamir_session_destroy", "cohost_shamir_prove", "cohost_zkey_load", "cohost_zkey_load_file", "cohost_zkey_get_info",
    "cohost_zkey_query_download", "cohost_zkey_matrix_download", "cohost_zkey_vk_download", "cohost_wtns_load_file", "cohost_rep3_phase_times", "cohost_plonk_zkey_load_file", "cohost_plonk_zkey_destroy",
    "cohost_plonk_zkey_get_info", "cohost_plonk_round1_plain", "cohost_plonk_round1_rep3", "cohost_rep3_set_mpc_exchange",
    "cohost_proof_to_json", "cohost_public_inputs_to_json", "cohost_shared_witness_encode", "cohost_shared_witness_decode",
    "cohost_split_witness_rep3", "cohost_r1cs_info", "cohost_split_witness_files",
    "cohost_groth16_verify", "cohost_groth16_verify_json", "cohost_plonk_verify_json", "cohost_plonk_zkey_header",
    "cohost_vm_create", "cohost_vm_destroy", "cohost_vm_set_public", "cohost_vm_set_shared", "cohost_vm_run", "cohost_vm_get", "cohost_vm_stats",
    "cohost_shamir_session_set_shard", "cohost_shamir_set_mpc_exchange", "cohost_rep3_session_create_blocks", "cohost_block_plan", "cohost_plonk_zkey_create_synthetic", "cohost_plonk_proof_limbs", "cohost_plonk_session_create", "cohost_plonk_session_create_shamir", "cohost_plonk_session_parties", "cohost_plonk_session_destroy",
    "cohost_plonk_prove", "cohost_plonk_set_mpc_exchange", "cohost_plonk_launch_count", "cohost_plonk_profile_enable",
    "cohost_plonk_profile_reset", "cohost_plonk_profile_read", "cohost_plonk_round_times", "cohost_plonk_trace_enable",
    "cohost_plonk_trace_get", "cohost_plonk_proof_to_json",
]
PROF_CLASSES = ["msm_sort", "msm_accumulate", "msm_reduce", "ntt", "vec", "spmv"]

vp, sz, ci, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64


class ZKeyDesc(ctypes.Structure):
    _fields_ = [("curve", ci), ("device", ci), ("n_public", sz), ("n_vars", sz), ("pow", sz), ("num_constraints", sz),
                ("a_rowptr", vp), ("a_col", vp), ("a_coeff", vp), ("a_nnz", sz),
                ("b_rowptr", vp), ("b_col", vp), ("b_coeff", vp), ("b_nnz", sz),
                ("a_query", vp), ("b_g1_query", vp), ("b_g2_query", vp), ("h_query", vp), ("l_query", vp),
                ("alpha_g1", vp), ("beta_g1", vp), ("delta_g1", vp), ("beta_g2", vp), ("delta_g2", vp), ("synthetic_seed", vp), ("rank", ci), ("world", ci), ("coeff_form", ci), ("shard_mode", ci)]


class CommOp(ctypes.Structure):
    _fields_ = [("dir", ci), ("peer", ci), ("dptr", vp), ("bytes", sz)]


class ZKeyInfo(ctypes.Structure):
    _fields_ = [("curve", ci), ("n_public", sz), ("n_vars", sz), ("pow", sz), ("num_constraints", sz)]


class Rep3Randomness(ctypes.Structure):
    _fields_ = [("r", vp), ("s", vp), ("mask_rs", vp), ("mask_pt", vp), ("masks1", vp * 3), ("masks2", vp * 3)]


GATHER_CB = ctypes.CFUNCTYPE(ci, vp, vp, sz, vp)
COMM_CB = ctypes.CFUNCTYPE(ci, vp, ctypes.POINTER(CommOp), ci)

_host = None


def load_host():
    global _host
    if _host is not None:
        return _host
    _lib.load()  # libcocg.so first (libcohost.so links it through $ORIGIN)
    if not os.path.exists(HOST_LIB_PATH):
        raise CocgError(f"{HOST_LIB_PATH} is missing: build it with `make -C {_HERE}/host`")
    L = ctypes.CDLL(HOST_LIB_PATH)
    pvp = ctypes.POINTER(vp)
    L.cohost_last_error.restype = ctypes.c_char_p
    L.cohost_zkey_create.argtypes = [ctypes.POINTER(ZKeyDesc), pvp]
    L.cohost_zkey_destroy.argtypes = [vp]
    L.cohost_zkey_destroy.restype = None
    L.cohost_plain_session_create.argtypes = [vp, pvp]
    L.cohost_plain_session_destroy.argtypes = [vp]
    L.cohost_plain_session_destroy.restype = None
    L.cohost_plain_prove.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.cohost_rep3_session_create.argtypes = [vp, vp, ci, ci, pvp]
    L.cohost_vm_create.argtypes = [ci, ci, ci, vp, sz, ci, pvp]
    L.cohost_vm_destroy.argtypes = [vp]
    L.cohost_vm_destroy.restype = None
    L.cohost_vm_set_public.argtypes = [vp, ci, vp]
    L.cohost_vm_set_shared.argtypes = [vp, ci, ci, vp, vp]
    L.cohost_vm_run.argtypes = [vp, vp, sz]
    L.cohost_vm_get.argtypes = [vp, ci, ci, ctypes.POINTER(ci), vp, vp]
    L.cohost_vm_stats.argtypes = [vp, ctypes.POINTER(u64)]
    L.cohost_rep3_session_create_blocks.argtypes = [vp, vp, ci, ci, COMM_CB, vp, pvp]
    L.cohost_block_plan.argtypes = [ci, ctypes.POINTER(ci)]
    L.cohost_rep3_session_destroy.argtypes = [vp]
    L.cohost_rep3_session_destroy.restype = None
    L.cohost_rep3_prove_begin.argtypes = [vp, vp, pvp, pvp, ctypes.POINTER(Rep3Randomness)]
    L.cohost_rep3_prove_begin_device.argtypes = [vp, vp, pvp, pvp, ctypes.POINTER(Rep3Randomness)]
    L.cohost_rep3_profile_enable.argtypes = [vp, ci]
    L.cohost_rep3_profile_read.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64)]
    L.cohost_rep3_profile_reset.argtypes = [vp]
    L.cohost_rep3_partial_bytes.argtypes = [vp]
    L.cohost_rep3_partial_bytes.restype = sz
    L.cohost_rep3_prove_partials.argtypes = [vp, vp]
    L.cohost_rep3_prove_combine.argtypes = [vp, vp]
    L.cohost_rep3_prove_end.argtypes = [vp, vp, pvp, pvp]
    L.cohost_rep3_launch_count.argtypes = [vp]
    L.cohost_rep3_launch_count.restype = u64
    L.cohost_zkey_load.argtypes = [vp, sz, ci, pvp]
    L.cohost_zkey_load_file.argtypes = [ctypes.c_char_p, ci, pvp]
    L.cohost_zkey_get_info.argtypes = [vp, ctypes.POINTER(ZKeyInfo)]
    L.cohost_zkey_query_download.argtypes = [vp, ci, sz, sz, vp]
    L.cohost_zkey_matrix_download.argtypes = [vp, ci, vp, vp, vp, ctypes.POINTER(sz)]
    L.cohost_zkey_vk_download.argtypes = [vp, vp]
    L.cohost_wtns_load_file.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(sz)]
    L.cohost_plonk_zkey_load_file.argtypes = [ctypes.c_char_p, ci, pvp]
    L.cohost_plonk_zkey_destroy.argtypes = [vp]
    L.cohost_plonk_zkey_destroy.restype = None
    L.cohost_plonk_zkey_get_info.argtypes = [vp, ctypes.POINTER(sz)]
    L.cohost_plonk_round1_plain.argtypes = [vp, vp, vp, ci, vp]
    L.cohost_plonk_round1_rep3.argtypes = [vp, vp, pvp, pvp, vp, ci, vp]
    L.cohost_plonk_zkey_create_synthetic.argtypes = [ci, ci, sz, sz, sz, sz, vp, vp, vp, vp, pvp]
    L.cohost_plonk_proof_limbs.argtypes = [vp]
    L.cohost_plonk_proof_limbs.restype = sz
    L.cohost_plonk_session_create.argtypes = [vp, ci, vp, pvp]
    L.cohost_plonk_session_create_shamir.argtypes = [vp, ci, ci, vp, pvp]
    L.cohost_plonk_session_parties.argtypes = [vp]
    L.cohost_plonk_session_destroy.argtypes = [vp]
    L.cohost_plonk_session_destroy.restype = None
    L.cohost_plonk_prove.argtypes = [vp, vp, pvp, pvp, ci, ci, vp]
    L.cohost_plonk_set_mpc_exchange.argtypes = [vp, ci]
    L.cohost_plonk_launch_count.argtypes = [vp]
    L.cohost_plonk_launch_count.restype = u64
    L.cohost_plonk_profile_enable.argtypes = [vp, ci]
    L.cohost_plonk_profile_reset.argtypes = [vp]
    L.cohost_plonk_profile_read.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64)]
    L.cohost_plonk_round_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.cohost_plonk_trace_enable.argtypes = [vp, ci]
    L.cohost_plonk_trace_get.argtypes = [vp, ci, ctypes.c_char_p, vp, sz, ctypes.POINTER(sz)]
    L.cohost_plonk_proof_to_json.argtypes = [ci, vp, vp, sz, ctypes.POINTER(sz)]
    L.cohost_rep3_set_mpc_exchange.argtypes = [vp, ci]
    L.cohost_rep3_phase_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.cohost_shamir_session_create.argtypes = [vp, ci, ci, vp, pvp]
    L.cohost_shamir_session_destroy.argtypes = [vp]
    L.cohost_shamir_session_destroy.restype = None
    L.cohost_shamir_prove.argtypes = [vp, vp, pvp, vp, vp]
    L.cohost_shamir_session_set_shard.argtypes = [vp, ci, ci, GATHER_CB, vp]
    L.cohost_shamir_set_mpc_exchange.argtypes = [vp, ci]
    L.cohost_msm_shard_range.argtypes = [sz, ci, ci, ctypes.POINTER(sz), ctypes.POINTER(sz)]
    L.cohost_proof_to_json.argtypes = [ci, vp, vp, sz, ctypes.POINTER(sz)]
    L.cohost_public_inputs_to_json.argtypes = [ci, vp, sz, vp, sz, ctypes.POINTER(sz)]
    L.cohost_shared_witness_encode.argtypes = [ci, vp, sz, pvp, ci, sz, vp, sz, ctypes.POINTER(sz)]
    L.cohost_shared_witness_decode.argtypes = [ci, vp, sz, ci, ctypes.POINTER(sz), ctypes.POINTER(sz), vp, pvp]
    L.cohost_split_witness_rep3.argtypes = [ci, ci, vp, sz, vp, pvp, pvp]
    L.cohost_r1cs_info.argtypes = [ctypes.c_char_p, ctypes.POINTER(sz)]
    L.cohost_groth16_verify.argtypes = [ci, vp, vp, sz, vp, vp, ctypes.POINTER(ci)]
    L.cohost_groth16_verify_json.argtypes = [ctypes.c_char_p, sz, ctypes.c_char_p, sz, ctypes.c_char_p, sz, ctypes.POINTER(ci)]
    L.cohost_plonk_zkey_header.argtypes = [ctypes.c_char_p, ctypes.POINTER(sz), vp, vp, vp]
    L.cohost_plonk_verify_json.argtypes = [ctypes.c_char_p, sz, ctypes.c_char_p, sz, ctypes.c_char_p, sz, vp, ctypes.POINTER(ci)]
    L.cohost_split_witness_files.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ci, ci, ci, ci, vp, ctypes.c_char_p, ci]
    _host = L
    return L


def _ck(rc):
    if rc:
        raise CocgError(load_host().cohost_last_error().decode())


def _c(a, dtype=np.uint64):
    return np.ascontiguousarray(a, dtype=dtype)


def _need(cond, msg):
    """Argument validation at the C boundary: the C side reads raw pointers, so a wrong length must raise (never `assert`, which
    `python -O` removes)."""
    if not cond:
        raise CocgError(msg)


def _sized(call):
    """Two-step writer protocol of include/cohost.h: query the length, then fill a buffer of that size."""
    n = sz(0)
    _ck(call(None, 0, ctypes.byref(n)))
    buf = ctypes.create_string_buffer(max(n.value, 1))
    _ck(call(buf, n.value, ctypes.byref(n)))
    return buf.raw[:n.value]


def proof_to_json(curve: int, proof) -> str:
    """snarkjs / serde_json text of a proof block (A | B | C packed affine Montgomery) as `co-circom generate-proof` writes it."""
    L = load_host()
    p = _c(proof)
    return _sized(lambda out, cap, n: L.cohost_proof_to_json(curve, p.ctypes.data, out, cap, n)).decode()


def public_inputs_to_json(curve: int, pub) -> str:
    """JSON array of decimal strings of pub[1:] (pub[0] is the constant 1), Montgomery Fr in."""
    L = load_host()
    p = _c(pub).reshape(-1, 4)
    return _sized(lambda out, cap, n: L.cohost_public_inputs_to_json(curve, p.ctypes.data, p.shape[0], out, cap, n)).decode()


def shared_witness_encode(curve: int, pub, comps) -> bytes:
    """SharedWitness file image (bincode + ark-serialize); comps = [a, b] for REP3, [a] for Shamir; Montgomery Fr in."""
    L = load_host()
    p = _c(pub).reshape(-1, 4)
    cs = [_c(c).reshape(-1, 4) for c in comps]
    arr = (vp * len(cs))(*[c.ctypes.data for c in cs])
    return _sized(lambda out, cap, n: L.cohost_shared_witness_encode(curve, p.ctypes.data, p.shape[0], arr, len(cs), cs[0].shape[0], out, cap, n))


def shared_witness_decode(curve: int, data: bytes, k: int):
    """-> (public_inputs, [components]) as Montgomery Fr arrays."""
    L = load_host()
    npub, n = sz(0), sz(0)
    buf = ctypes.create_string_buffer(data, len(data))
    _ck(L.cohost_shared_witness_decode(curve, buf, len(data), k, ctypes.byref(npub), ctypes.byref(n), None, None))
    pub = np.zeros((npub.value, 4), dtype=np.uint64)
    comps = [np.zeros((n.value, 4), dtype=np.uint64) for _ in range(k)]
    arr = (vp * k)(*[c.ctypes.data for c in comps])
    _ck(L.cohost_shared_witness_decode(curve, buf, len(data), k, ctypes.byref(npub), ctypes.byref(n), pub.ctypes.data, arr))
    return pub, comps


def split_witness_rep3(curve: int, witness, seed: bytes, device: int = 0):
    """SharedWitness::share_rep3 on the GPU: three (a, b) pairs of Montgomery Fr arrays; seed = 64 bytes."""
    L = load_host()
    _need(len(seed) == 64, "split_witness_rep3: seed must be 64 bytes")
    w = _c(witness).reshape(-1, 4)
    n = w.shape[0]
    oa = [np.zeros((n, 4), dtype=np.uint64) for _ in range(3)]
    ob = [np.zeros((n, 4), dtype=np.uint64) for _ in range(3)]
    A = (vp * 3)(*[o.ctypes.data for o in oa])
    B = (vp * 3)(*[o.ctypes.data for o in ob])
    sb = ctypes.create_string_buffer(seed, 64)
    _ck(L.cohost_split_witness_rep3(curve, device, w.ctypes.data, n, sb, A, B))
    return list(zip(oa, ob))


def groth16_verify_json(vk_json: str, proof_json: str, public_json: str) -> bool:
    """`co-circom verify groth16`: True = accepted, False = rejected by the pairing check; CocgError for malformed input."""
    ok = ci(0)
    a, b, c = vk_json.encode(), proof_json.encode(), public_json.encode()
    _ck(load_host().cohost_groth16_verify_json(a, len(a), b, len(b), c, len(c), ctypes.byref(ok)))
    return bool(ok.value)


def plonk_verify_json(vk_json: str, proof_json: str, public_json: str, want_challenges: bool = False):
    """`co-circom verify plonk`.  With want_challenges also the Montgomery Fr rows alpha, beta, gamma, xi, v[0], u."""
    ok = ci(0)
    a, b, c = vk_json.encode(), proof_json.encode(), public_json.encode()
    ch = np.zeros((6, 4), dtype=np.uint64)
    _ck(load_host().cohost_plonk_verify_json(a, len(a), b, len(b), c, len(c), ch.ctypes.data if want_challenges else None, ctypes.byref(ok)))
    return (bool(ok.value), ch) if want_challenges else bool(ok.value)


def groth16_verify(curve: int, vk, ic, proof, pub) -> bool:
    """Binary form: vk = alpha_g1 | beta_g2 | gamma_g2 | delta_g2, ic = (n_ic, 2 lq), proof = A | B | C, pub = (n_ic - 1, 4), all Montgomery."""
    ok = ci(0)
    v, i, p, u = _c(vk), _c(ic), _c(proof), _c(pub)
    lq = 4 if curve == _lib.BN254 else 6
    _ck(load_host().cohost_groth16_verify(curve, v.ctypes.data, i.ctypes.data, i.size // (2 * lq), p.ctypes.data, u.ctypes.data, ctypes.byref(ok)))
    return bool(ok.value)


def plonk_zkey_header(path: str) -> dict:
    """Header of a Plonk zkey read by the product's C++ reader (no GPU): counts, k1 / k2, the eight G1 commitments, X_2 (Montgomery)."""
    info = (sz * 7)()
    _ck(load_host().cohost_plonk_zkey_header(path.encode(), info, None, None, None))
    lq = 4 if int(info[0]) == _lib.BN254 else 6
    k = np.zeros((2, 4), dtype=np.uint64)
    g1 = np.zeros((8, 2 * lq), dtype=np.uint64)
    x2 = np.zeros((1, 4 * lq), dtype=np.uint64)
    _ck(load_host().cohost_plonk_zkey_header(path.encode(), info, k.ctypes.data, g1.ctypes.data, x2.ctypes.data))
    out = dict(zip(("curve", "n_vars", "n_public", "domain_size", "n_additions", "n_constraints", "parts"), [int(x) for x in info]))
    out.update(k=k, vk_g1=g1, x_2=x2)
    return out


def block_plan(world: int) -> dict:
    """The block-mode work distribution for `world` ranks: {"wm": [rank x 3 parties], "g2": [[a, b] x 3],
    "g1": [[[l, a, b_g1] x 2 components] x 3 parties]}."""
    out = (ci * 27)()
    _ck(load_host().cohost_block_plan(world, out))
    v = [int(x) for x in out]
    return {"wm": v[:3], "g2": [v[3 + 2 * q:5 + 2 * q] for q in range(3)],
            "g1": [[v[9 + 6 * q + 3 * c:12 + 6 * q + 3 * c] for c in range(2)] for q in range(3)]}


def r1cs_info(path: str) -> dict:
    info = (sz * 6)()
    _ck(load_host().cohost_r1cs_info(path.encode(), info))
    return dict(zip(("curve", "n_wires", "n_pub_out", "n_pub_in", "n_constraints", "num_inputs"), [int(x) for x in info]))


def split_witness_files(witness: str, r1cs: str, protocol: str, curve: int, out_dir: str, threshold: int = 1, num_parties: int = 3,
                        seed: bytes | None = None, device: int = 0):
    """`co-circom split-witness`: writes <out_dir>/<witness name>.<i>.shared; seed defaults to os.urandom."""
    proto = {"REP3": 0, "SHAMIR": 1}[protocol.upper()]
    need = 32 * max(2, threshold)
    seed = os.urandom(need) if seed is None else seed
    _need(len(seed) >= need, f"split_witness_files: seed must be at least {need} bytes")
    sb = ctypes.create_string_buffer(seed, len(seed))
    _ck(load_host().cohost_split_witness_files(witness.encode(), r1cs.encode(), proto, curve, threshold, num_parties, sb, out_dir.encode(), device))
    n = 3 if proto == 0 else num_parties
    return [os.path.join(out_dir, f"{os.path.basename(witness)}.{i}.shared") for i in range(n)]


class Groth16ZKey:
    """A Groth16 proving key resident in HBM (query arrays as MSM bases, A/B matrices as CSR)."""

    def __init__(self, curve: int, n_public: int, n_vars: int, pow_: int, num_constraints: int, a_csr, b_csr,
                 a_query=None, b_g1_query=None, b_g2_query=None, h_query=None, l_query=None, alpha_g1=None, beta_g1=None,
                 delta_g1=None, beta_g2=None, delta_g2=None, device: int = 0, synthetic_seed: bytes | None = None, rank: int = 0,
                 world: int = 1, shard_mode: str = "ranges"):
        """Query / vk arrays left as None are generated in HBM from `synthetic_seed` (32 bytes).  world > 1: only rank's part of the
        queries becomes resident (host arrays are still passed whole): shard_mode "ranges" = its index range of every query,
        "blocks" = the whole queries of the blocks of the proof it runs (REP3; see block_plan)."""
        _need(shard_mode in ("ranges", "blocks"), "zkey: shard_mode must be ranges or blocks")
        self.shard_mode = shard_mode if world > 1 else "ranges"
        L = load_host()
        self.curve, self.lq = curve, (4 if curve == _lib.BN254 else 6)
        self.n_public, self.n_vars, self.pow, self.num_constraints = n_public, n_vars, pow_, num_constraints
        self.n_aux = n_vars - n_public - 1
        keep = []

        def P(a, dtype=np.uint64):
            a = _c(a, dtype)
            keep.append(a)
            return a.ctypes.data

        d = ZKeyDesc()
        d.curve, d.device, d.n_public, d.n_vars, d.pow, d.num_constraints = curve, device, n_public, n_vars, pow_, num_constraints
        d.rank, d.world = rank, world
        d.shard_mode = 1 if shard_mode == "blocks" else 0
        d.a_rowptr, d.a_col, d.a_coeff, d.a_nnz = P(a_csr[0], np.uint32), P(a_csr[1], np.uint32), P(a_csr[2]), len(a_csr[1])
        d.b_rowptr, d.b_col, d.b_coeff, d.b_nnz = P(b_csr[0], np.uint32), P(b_csr[1], np.uint32), P(b_csr[2]), len(b_csr[1])
        _need(len(a_csr[0]) == num_constraints + 1 and len(b_csr[0]) == num_constraints + 1, "zkey: rowptr must hold num_constraints + 1 entries")
        _need(len(a_csr[1]) * 4 == _c(a_csr[2]).size and len(b_csr[1]) * 4 == _c(b_csr[2]).size, "zkey: one coefficient per column index")
        lq = self.lq
        for name, arr, n, w in (("a_query", a_query, n_vars, 2), ("b_g1_query", b_g1_query, n_vars, 2), ("b_g2_query", b_g2_query, n_vars, 4),
                                ("h_query", h_query, 1 << pow_, 2), ("l_query", l_query, self.n_aux, 2)):
            if arr is None:
                _need(synthetic_seed is not None, f"{name} missing and no synthetic_seed")
                continue
            arr = _c(arr)
            _need(arr.size == n * w * lq, f"{name}: expected {n} points")
            setattr(d, name, P(arr))
        for name, arr in (("alpha_g1", alpha_g1), ("beta_g1", beta_g1), ("delta_g1", delta_g1), ("beta_g2", beta_g2), ("delta_g2", delta_g2)):
            if arr is not None:
                setattr(d, name, P(arr))
        if synthetic_seed is not None:
            _need(len(synthetic_seed) == 32, "synthetic_seed must be 32 bytes")
            d.synthetic_seed = P(np.frombuffer(synthetic_seed, dtype=np.uint8), np.uint8)
        h = vp()
        _ck(L.cohost_zkey_create(ctypes.byref(d), ctypes.byref(h)))
        self.h = h

    @classmethod
    def from_file(cls, path: str, device: int = 0) -> "Groth16ZKey":
        """ZKey::from_reader: parse a snarkjs Groth16 .zkey in the C++ host layer and make it resident in HBM."""
        L = load_host()
        h = vp()
        _ck(L.cohost_zkey_load_file(path.encode(), device, ctypes.byref(h)))
        self = cls.__new__(cls)
        self.h = h
        self.shard_mode = "ranges"
        info = ZKeyInfo()
        _ck(L.cohost_zkey_get_info(h, ctypes.byref(info)))
        self.curve, self.lq = info.curve, (4 if info.curve == _lib.BN254 else 6)
        self.n_public, self.n_vars, self.pow, self.num_constraints = info.n_public, info.n_vars, info.pow, info.num_constraints
        self.n_aux = self.n_vars - self.n_public - 1
        return self

    QUERIES = {"a_query": (0, 1), "b_g1_query": (1, 1), "b_g2_query": (2, 2), "h_query": (3, 1), "l_query": (4, 1)}

    def query(self, name: str) -> np.ndarray:
        which, group = self.QUERIES[name]
        n = {"h_query": 1 << self.pow, "l_query": self.n_aux}.get(name, self.n_vars)
        out = np.zeros((n, 2 * group * self.lq), dtype=np.uint64)
        _ck(load_host().cohost_zkey_query_download(self.h, which, 0, n, out.ctypes.data))
        return out

    def matrix(self, which: int):
        nnz = sz()
        _ck(load_host().cohost_zkey_matrix_download(self.h, which, None, None, None, ctypes.byref(nnz)))
        rowptr = np.zeros(self.num_constraints + 1, dtype=np.uint32)
        col = np.zeros(nnz.value, dtype=np.uint32)
        coeff = np.zeros((nnz.value, 4), dtype=np.uint64)
        _ck(load_host().cohost_zkey_matrix_download(self.h, which, rowptr.ctypes.data, col.ctypes.data, coeff.ctypes.data, ctypes.byref(nnz)))
        return rowptr, col, coeff

    def vk(self) -> dict:
        lq = self.lq
        out = np.zeros(14 * lq, dtype=np.uint64)
        _ck(load_host().cohost_zkey_vk_download(self.h, out.ctypes.data))
        return {"alpha_g1": out[:2 * lq], "beta_g1": out[2 * lq:4 * lq], "delta_g1": out[4 * lq:6 * lq], "beta_g2": out[6 * lq:10 * lq],
                "delta_g2": out[10 * lq:14 * lq]}

    def load_witness(self, path: str) -> np.ndarray:
        """Witness::from_reader: .wtns values as (n, 4) Montgomery limbs."""
        n = sz()
        _ck(load_host().cohost_wtns_load_file(self.h, path.encode(), None, ctypes.byref(n)))
        out = np.zeros((n.value, 4), dtype=np.uint64)
        _ck(load_host().cohost_wtns_load_file(self.h, path.encode(), out.ctypes.data, ctypes.byref(n)))
        return out

    def close(self):
        if self.h:
            load_host().cohost_zkey_destroy(self.h)
            self.h = None


class PlainSession:
    """CoGroth16<PlainDriver>."""

    def __init__(self, zkey: Groth16ZKey):
        self.zkey = zkey
        h = vp()
        _ck(load_host().cohost_plain_session_create(zkey.h, ctypes.byref(h)))
        self.h = h

    def prove(self, public_inputs, witness, r=None, s=None, want_h=False):
        zk = self.zkey
        pub, wit = _c(public_inputs), _c(witness)
        _need(pub.size == 4 * (zk.n_public + 1), f"prove: expected {zk.n_public + 1} public inputs")
        _need(wit.size == 4 * zk.n_aux, f"prove: expected {zk.n_aux} witness elements")
        proof = np.zeros(8 * zk.lq, dtype=np.uint64)
        hbuf = np.zeros((1 << zk.pow, 4), dtype=np.uint64) if want_h else None
        rr = None if r is None else _c(r)
        ss = None if s is None else _c(s)
        _ck(load_host().cohost_plain_prove(self.h, pub.ctypes.data, wit.ctypes.data, None if rr is None else rr.ctypes.data,
                                           None if ss is None else ss.ctypes.data, proof.ctypes.data,
                                           None if hbuf is None else hbuf.ctypes.data))
        return (proof, hbuf) if want_h else proof

    def close(self):
        if self.h:
            load_host().cohost_plain_session_destroy(self.h)
            self.h = None


class Rep3Session:
    """Three CoGroth16<Rep3Protocol> provers on three threads over an in-process network (one GPU, or one MSM shard of `world`)."""

    def __init__(self, zkey: Groth16ZKey, seeds: bytes | None = None, rank: int = 0, world: int = 1, comm=None):
        """comm (zkeys made with shard_mode="blocks" only): f(ops) performing the cross-GPU transfers of one mul_vec round, ops = list of
        (dir, peer_rank, device_ptr, nbytes) with dir 0 = send / 1 = receive -- see distributed.make_p2p.
        seeds: 3 x 32 bytes, the parties' PRF seeds (Rep3Protocol::new draws them from entropy, rep3.rs:343-349).  Every mask and
        the Groth16 blinders r, s derive from them, so the default is os.urandom; fixed seeds are for tests and the benchmark only.
        In a multi-GPU run every rank must pass the SAME seeds (the ranks replay the same three parties)."""
        seeds = os.urandom(96) if seeds is None else seeds
        _need(len(seeds) == 96, "Rep3Session: seeds must be 3 x 32 bytes")
        self.zkey, self.rank, self.world = zkey, rank, world
        self._seeds = np.frombuffer(seeds, dtype=np.uint8).copy()
        h = vp()
        self._comm_cb = None
        if getattr(zkey, "shard_mode", "ranges") == "blocks":
            _need(comm is not None, "Rep3Session: a block-mode zkey needs a comm function (distributed.make_p2p)")

            def cb(_user, ops, nops):
                try:
                    comm([(ops[i].dir, ops[i].peer, ops[i].dptr, ops[i].bytes) for i in range(nops)])
                    return 0
                except Exception:  # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return 1

            self._comm_cb = COMM_CB(cb)
            _ck(load_host().cohost_rep3_session_create_blocks(zkey.h, self._seeds.ctypes.data, rank, world, self._comm_cb, None, ctypes.byref(h)))
        else:
            _ck(load_host().cohost_rep3_session_create(zkey.h, self._seeds.ctypes.data, rank, world, ctypes.byref(h)))
        self.h = h

    def partial_bytes(self) -> int:
        return int(load_host().cohost_rep3_partial_bytes(self.h))

    def begin(self, public_inputs, wit_a, wit_b, rnd: dict | None = None, device_ptrs: bool = False):
        """wit_a / wit_b: three host arrays each (numpy, or raw HOST addresses of pinned buffers given as int), or -- with
        device_ptrs=True -- three DEVICE addresses each."""
        zk = self.zkey
        self._keep = [_c(public_inputs)]
        pub = self._keep[0]
        _need(pub.size == 4 * (zk.n_public + 1), f"prove: expected {zk.n_public + 1} public inputs")
        _need(len(wit_a) == 3 and len(wit_b) == 3, "prove: three share components per side")

        def addr(x):
            if isinstance(x, int):
                return x
            a = _c(x)
            _need(a.size == 4 * zk.n_aux, f"prove: expected {zk.n_aux} witness share elements")
            self._keep.append(a)
            return a.ctypes.data

        A = (vp * 3)(*[addr(x) for x in wit_a])
        B = (vp * 3)(*[addr(x) for x in wit_b])
        R = None
        if rnd is not None:
            R = Rep3Randomness()
            for k in ("r", "s", "mask_rs", "mask_pt"):
                a = _c(rnd[k])
                self._keep.append(a)
                setattr(R, k, a.ctypes.data)
            for k in ("masks1", "masks2"):
                arrs = [_c(x) for x in rnd[k]]
                self._keep += arrs
                setattr(R, k, (vp * 3)(*[x.ctypes.data for x in arrs]))
            self._keep.append(R)
        fn = load_host().cohost_rep3_prove_begin_device if device_ptrs else load_host().cohost_rep3_prove_begin
        _ck(fn(self.h, pub.ctypes.data, A, B, None if R is None else ctypes.byref(R)))

    def partials(self) -> np.ndarray:
        out = np.zeros(self.partial_bytes() // 8, dtype=np.uint64)
        _ck(load_host().cohost_rep3_prove_partials(self.h, out.ctypes.data))
        return out

    def combine(self, gathered: np.ndarray):
        g = _c(gathered)
        _need(g.nbytes == self.world * self.partial_bytes(), "combine: gathered buffer has the wrong size")
        _ck(load_host().cohost_rep3_prove_combine(self.h, g.ctypes.data))

    def end(self, want_h=False):
        zk = self.zkey
        proofs = np.zeros((3, 8 * zk.lq), dtype=np.uint64)
        ha = hb = None
        HA = HB = None
        if want_h:
            ha = [np.zeros((1 << zk.pow, 4), dtype=np.uint64) for _ in range(3)]
            hb = [np.zeros((1 << zk.pow, 4), dtype=np.uint64) for _ in range(3)]
            HA = (vp * 3)(*[x.ctypes.data for x in ha])
            HB = (vp * 3)(*[x.ctypes.data for x in hb])
        _ck(load_host().cohost_rep3_prove_end(self.h, proofs.ctypes.data, HA, HB))
        self._keep = None
        return (proofs, ha, hb) if want_h else proofs

    def profile(self, on: bool = True):
        _ck(load_host().cohost_rep3_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        _ck(load_host().cohost_rep3_profile_reset(self.h))

    def profile_read(self) -> dict:
        out = {}
        for i, name in enumerate(PROF_CLASSES):
            ms, k = ctypes.c_double(), u64()
            _ck(load_host().cohost_rep3_profile_read(self.h, i, ctypes.byref(ms), ctypes.byref(k)))
            out[name] = (ms.value, int(k.value))
        return out

    def prove(self, public_inputs, wit_a, wit_b, rnd=None, want_h=False, all_gather=None, device_ptrs=False):
        """One proof.  all_gather(partials: np.ndarray) -> concatenation over ranks; required when world > 1."""
        self.begin(public_inputs, wit_a, wit_b, rnd, device_ptrs)
        if self.world > 1:
            self.combine(all_gather(self.partials()))
        return self.end(want_h)

    def launch_count(self) -> int:
        return int(load_host().cohost_rep3_launch_count(self.h))

    def set_mpc_exchange(self, mode: str):
        """'host': mul_vec payloads staged through pinned host memory (default); 'device': handed over in HBM (co-located parties)."""
        _need(mode in ("host", "device"), "set_mpc_exchange: mode must be host or device")
        _ck(load_host().cohost_rep3_set_mpc_exchange(self.h, 1 if mode == "device" else 0))

    def phase_times(self) -> np.ndarray:
        """(3, 4) seconds of the last proof per party: witness map | MSMs | all-gather wait | assembly."""
        out = (ctypes.c_double * 12)()
        _ck(load_host().cohost_rep3_phase_times(self.h, out))
        return np.array(out).reshape(3, 4)

    def close(self):
        if self.h:
            load_host().cohost_rep3_session_destroy(self.h)
            self.h = None


class ShamirSession:
    """n CoGroth16<ShamirProtocol> provers (threshold t) on n threads over an in-process network."""

    def __init__(self, zkey: Groth16ZKey, num_parties: int = 3, threshold: int = 1, seeds: bytes | None = None, rank: int = 0, world: int = 1,
                 all_gather=None):
        """world > 1: MSMs sharded by index range over `world` ranks; all_gather(np.ndarray uint64) -> concatenation over ranks (one
        NCCL all-gather per proof).  Every rank must pass the same seeds."""
        self.zkey, self.n, self.t = zkey, num_parties, threshold
        # every double-random pair, and through rand() the Groth16 blinders r and s, derive from these seeds: entropy by default
        # (ShamirProtocol::new seeds its RngType from entropy, shamir.rs:196-245); fixed seeds are for tests and the benchmark only
        seeds = os.urandom(32 * num_parties) if seeds is None else seeds
        _need(len(seeds) == 32 * num_parties, "ShamirSession: seeds must be num_parties x 32 bytes")
        self._seeds = np.frombuffer(seeds, dtype=np.uint8).copy()
        h = vp()
        _ck(load_host().cohost_shamir_session_create(zkey.h, num_parties, threshold, self._seeds.ctypes.data, ctypes.byref(h)))
        self.h = h
        self._cb = None
        if world > 1:
            _need(all_gather is not None, "ShamirSession: world > 1 needs an all_gather function")

            def cb(_user, local, nbytes, gathered):
                try:
                    loc = np.ctypeslib.as_array(ctypes.cast(local, ctypes.POINTER(ctypes.c_uint64)), shape=(nbytes // 8,))
                    out = np.ascontiguousarray(all_gather(loc.copy()), dtype=np.uint64)
                    if out.size != world * (nbytes // 8):
                        return 1
                    ctypes.memmove(gathered, out.ctypes.data, out.nbytes)
                    return 0
                except Exception:  # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return 1

            self._cb = GATHER_CB(cb)
            _ck(load_host().cohost_shamir_session_set_shard(self.h, rank, world, self._cb, None))

    def set_mpc_exchange(self, mode: str):
        """'host': share vectors staged through pinned host memory (default); 'device': handed over in HBM (co-located parties)."""
        _need(mode in ("host", "device"), "set_mpc_exchange: mode must be host or device")
        _ck(load_host().cohost_shamir_set_mpc_exchange(self.h, 1 if mode == "device" else 0))

    def prove(self, public_inputs, wit):
        """wit: n host share vectors.  Returns (proofs (n, A|B|C), rs (n, 2, 4): each party's shares of r and s)."""
        zk = self.zkey
        pub = _c(public_inputs)
        ws = [_c(x) for x in wit]
        _need(len(ws) == self.n, f"prove: expected {self.n} share vectors")
        _need(pub.size == 4 * (zk.n_public + 1), f"prove: expected {zk.n_public + 1} public inputs")
        _need(all(x.size == 4 * zk.n_aux for x in ws), f"prove: expected {zk.n_aux} witness share elements per party")
        W = (vp * self.n)(*[x.ctypes.data for x in ws])
        proofs = np.zeros((self.n, 8 * zk.lq), dtype=np.uint64)
        rs = np.zeros((self.n, 2, 4), dtype=np.uint64)
        _ck(load_host().cohost_shamir_prove(self.h, pub.ctypes.data, W, proofs.ctypes.data, rs.ctypes.data))
        return proofs, rs

    def close(self):
        if self.h:
            load_host().cohost_shamir_session_destroy(self.h)
            self.h = None


def plonk_proof_to_json(curve: int, proof) -> str:
    """snarkjs / serde_json text of a Plonk proof block (9 points | 6 evaluations) as `co-circom generate-proof plonk` writes it."""
    L = load_host()
    p = _c(proof)
    return _sized(lambda out, cap, n: L.cohost_plonk_proof_to_json(curve, p.ctypes.data, out, cap, n)).decode()


class PlonkZKey:
    """A snarkjs Plonk proving key resident in HBM: wire maps, selector / sigma / Lagrange polynomials, p_tau."""

    def __init__(self, path: str | None = None, device: int = 0, _handle=None):
        if _handle is None:
            h = vp()
            _ck(load_host().cohost_plonk_zkey_load_file(path.encode(), device, ctypes.byref(h)))
        else:
            h = _handle
        self.h = h
        info = (sz * 6)()
        _ck(load_host().cohost_plonk_zkey_get_info(h, info))
        self.curve, self.n_vars, self.n_public, self.domain_size, self.n_additions, self.n_constraints = (int(x) for x in info)
        self.lq = 4 if self.curve == _lib.BN254 else 6
        self.n_witness = self.n_vars - self.n_additions - self.n_public - 1
        self.proof_limbs = int(load_host().cohost_plonk_proof_limbs(h))

    @classmethod
    def synthetic(cls, curve: int, log_n: int, n_public: int, n_vars: int, maps, seed: bytes, device: int = 0) -> "PlonkZKey":
        """Shape-faithful benchmark key: maps = (map_a, map_b, map_c) uint32 arrays of n_constraints wire indices < n_vars."""
        _need(len(seed) == 32, "synthetic plonk key: seed must be 32 bytes")
        ma, mb, mc = (_c(m, np.uint32) for m in maps)
        _need(ma.size == mb.size == mc.size, "synthetic plonk key: the three wire maps must have the same length")
        sd = np.frombuffer(seed, dtype=np.uint8).copy()
        h = vp()
        _ck(load_host().cohost_plonk_zkey_create_synthetic(curve, device, log_n, n_public, n_vars, ma.size, ma.ctypes.data, mb.ctypes.data, mc.ctypes.data,
                                                           sd.ctypes.data, ctypes.byref(h)))
        return cls(_handle=h)

    def round1_plain(self, public_inputs, witness, deterministic=True) -> np.ndarray:
        pub, wit = _c(public_inputs), _c(witness)
        _need(pub.size == 4 * (self.n_public + 1) and wit.size == 4 * self.n_witness, "round1: wrong input length")
        out = np.zeros((3, 2 * self.lq), dtype=np.uint64)
        _ck(load_host().cohost_plonk_round1_plain(self.h, pub.ctypes.data, wit.ctypes.data, 1 if deterministic else 0, out.ctypes.data))
        return out

    def round1_rep3(self, public_inputs, wit_a, wit_b, seeds: bytes = bytes(range(96)), deterministic=True) -> np.ndarray:
        pub = _c(public_inputs)
        wa, wb = [_c(x) for x in wit_a], [_c(x) for x in wit_b]
        _need(pub.size == 4 * (self.n_public + 1) and all(x.size == 4 * self.n_witness for x in wa + wb), "round1: wrong input length")
        A = (vp * 3)(*[x.ctypes.data for x in wa])
        B = (vp * 3)(*[x.ctypes.data for x in wb])
        sd = np.frombuffer(seeds, dtype=np.uint8).copy()
        out = np.zeros((3, 3, 2 * self.lq), dtype=np.uint64)
        _ck(load_host().cohost_plonk_round1_rep3(self.h, pub.ctypes.data, A, B, sd.ctypes.data, 1 if deterministic else 0, out.ctypes.data))
        return out

    def close(self):
        if self.h:
            load_host().cohost_plonk_zkey_destroy(self.h)
            self.h = None


class PlonkSession:
    """CoPlonk::prove: protocol 'plain' (one party), 'rep3' (three parties on three threads, in-process network) or 'shamir'
    (num_parties parties, threshold t; one share component per party)."""

    def __init__(self, zkey: PlonkZKey, protocol: str = "plain", seeds: bytes | None = None, num_parties: int = 3, threshold: int = 1):
        _need(protocol in ("plain", "rep3", "shamir"), "PlonkSession: protocol must be plain, rep3 or shamir")
        self.zkey, self.protocol = zkey, protocol
        self.parties = 1 if protocol == "plain" else 3 if protocol == "rep3" else num_parties
        seeds = os.urandom(32 * self.parties) if seeds is None else seeds  # blinders and masks derive from these: entropy by default
        _need(len(seeds) == 32 * self.parties, "PlonkSession: seeds must be 32 bytes per party")
        self._seeds = np.frombuffer(seeds, dtype=np.uint8).copy()
        h = vp()
        if protocol == "shamir":
            _ck(load_host().cohost_plonk_session_create_shamir(zkey.h, num_parties, threshold, self._seeds.ctypes.data, ctypes.byref(h)))
        else:
            _ck(load_host().cohost_plonk_session_create(zkey.h, 0 if protocol == "plain" else 1, self._seeds.ctypes.data, ctypes.byref(h)))
        self.h = h

    def prove(self, public_inputs, wit_a, wit_b=None, deterministic=False, device_ptrs=False) -> np.ndarray:
        """wit_a / wit_b: one array (or raw address) per party.  Returns (parties, proof_limbs) uint64."""
        zk = self.zkey
        pub = _c(public_inputs)
        _need(pub.size == 4 * (zk.n_public + 1), f"prove: expected {zk.n_public + 1} public inputs")
        keep = []

        def addr(x):
            if isinstance(x, int):
                return x
            a = _c(x)
            _need(a.size == 4 * zk.n_witness, f"prove: expected {zk.n_witness} witness share elements")
            keep.append(a)
            return a.ctypes.data

        _need(len(wit_a) == self.parties and (self.protocol != "rep3" or (wit_b is not None and len(wit_b) == 3)), "prove: one share per party")
        if self.protocol != "rep3":
            wit_b = None
        A = (vp * self.parties)(*[addr(x) for x in wit_a])
        B = None if wit_b is None else (vp * self.parties)(*[addr(x) for x in wit_b])
        out = np.zeros((self.parties, zk.proof_limbs), dtype=np.uint64)
        _ck(load_host().cohost_plonk_prove(self.h, pub.ctypes.data, A, B, 1 if deterministic else 0, 1 if device_ptrs else 0, out.ctypes.data))
        return out

    def set_mpc_exchange(self, mode: str):
        _need(mode in ("host", "device"), "set_mpc_exchange: mode must be host or device")
        _ck(load_host().cohost_plonk_set_mpc_exchange(self.h, 1 if mode == "device" else 0))

    def launch_count(self) -> int:
        return int(load_host().cohost_plonk_launch_count(self.h))

    def profile(self, on: bool = True):
        _ck(load_host().cohost_plonk_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        _ck(load_host().cohost_plonk_profile_reset(self.h))

    def profile_read(self) -> dict:
        out = {}
        for i, name in enumerate(PROF_CLASSES):
            ms, k = ctypes.c_double(), u64()
            _ck(load_host().cohost_plonk_profile_read(self.h, i, ctypes.byref(ms), ctypes.byref(k)))
            out[name] = (ms.value, int(k.value))
        return out

    def round_times(self) -> np.ndarray:
        out = (ctypes.c_double * (5 * self.parties))()
        _ck(load_host().cohost_plonk_round_times(self.h, out))
        return np.array(out).reshape(self.parties, 5)

    def trace(self, on: bool = True):
        _ck(load_host().cohost_plonk_trace_enable(self.h, 1 if on else 0))

    def trace_get(self, party: int, name: str) -> np.ndarray:
        n = sz(0)
        _ck(load_host().cohost_plonk_trace_get(self.h, party, name.encode(), None, 0, ctypes.byref(n)))
        out = np.zeros((n.value, 4), dtype=np.uint64)
        _ck(load_host().cohost_plonk_trace_get(self.h, party, name.encode(), out.ctypes.data, n.value, ctypes.byref(n)))
        return out

    def close(self):
        if self.h:
            load_host().cohost_plonk_session_destroy(self.h)
            self.h = None


VM_ADD, VM_SUB, VM_MUL, VM_NEG, VM_DIV = range(5)


class BatchedVm:
    """The field opcodes of the MPC witness-extension VM over a batch of independent inputs (host/vm.hpp)."""

    def __init__(self, curve: int, protocol: str, batch: int, n_regs: int, seeds: bytes | None = None, device: int = 0):
        _need(protocol in ("plain", "rep3"), "BatchedVm: protocol must be plain or rep3")
        self.parties = 1 if protocol == "plain" else 3
        self.batch = batch
        seeds = os.urandom(32 * self.parties) if seeds is None else seeds
        _need(len(seeds) == 32 * self.parties, "BatchedVm: seeds must be 32 bytes per party")
        sd = np.frombuffer(seeds, dtype=np.uint8).copy()
        h = vp()
        _ck(load_host().cohost_vm_create(curve, device, 0 if protocol == "plain" else 1, sd.ctypes.data, batch, n_regs, ctypes.byref(h)))
        self.h = h

    def set_public(self, reg: int, values):
        v = _c(values)
        _need(v.size == 4 * self.batch, "set_public: one value per instance")
        _ck(load_host().cohost_vm_set_public(self.h, reg, v.ctypes.data))

    def set_shared(self, reg: int, party: int, a, b=None):
        a = _c(a)
        b = None if b is None else _c(b)
        _need(a.size == 4 * self.batch and (b is None or b.size == 4 * self.batch), "set_shared: one share per instance")
        _ck(load_host().cohost_vm_set_shared(self.h, reg, party, a.ctypes.data, None if b is None else b.ctypes.data))

    def run(self, program):
        """program: iterable of (op, dst, lhs, rhs)"""
        prog = np.ascontiguousarray(np.array(list(program), dtype=np.int32).reshape(-1, 4))
        _ck(load_host().cohost_vm_run(self.h, prog.ctypes.data, prog.shape[0]))

    def get(self, reg: int, party: int = 0):
        """-> ("public", values) or ("shared", a, b)"""
        kind = ci(0)
        a = np.zeros((self.batch, 4), dtype=np.uint64)
        b = np.zeros((self.batch, 4), dtype=np.uint64)
        _ck(load_host().cohost_vm_get(self.h, reg, party, ctypes.byref(kind), a.ctypes.data, b.ctypes.data))
        if kind.value == 1:
            return ("public", a)
        _need(kind.value == 2, "get: empty register")
        return ("shared", a, b)

    def stats(self) -> dict:
        out = (u64 * 2)()
        _ck(load_host().cohost_vm_stats(self.h, out))
        return {"launches": int(out[0]), "network_rounds": int(out[1])}

    def close(self):
        if self.h:
            load_host().cohost_vm_destroy(self.h)
            self.h = None
