"""cocg -- B200 (sm_100a) kernels behind collaborative-circom's MPC proving hot path.

The directory name carries a hyphen (it mirrors the reference's name), so import it with
``importlib.import_module("collaborative-circom_b200")`` or through the ``cocg`` shim at the repo root.
"""
from ._lib import (BN254, BLS12_381, G1, G2, OP_MUL, OP_ADD, OP_SUB, OP_NEG, OP_TO_MONT, OP_FROM_MONT,  # noqa: F401
                   EC_ADD, EC_MUL, EC_TO_AFFINE, EC_FROM_AFFINE, EC_NEG, EC_DBL, SYMBOLS, LIB_PATH, CocgError, load, msm_plan)
from .context import Context, DeviceVec  # noqa: F401
from .prover import Groth16ZKey, PlainSession, Rep3Session, ShamirSession, PlonkZKey, PlonkSession, plonk_proof_to_json, block_plan, BatchedVm, VM_ADD, VM_SUB, VM_MUL, VM_NEG, VM_DIV, HOST_SYMBOLS, HOST_LIB_PATH, load_host, proof_to_json, public_inputs_to_json, shared_witness_encode, shared_witness_decode, split_witness_rep3, r1cs_info, split_witness_files, groth16_verify, groth16_verify_json, plonk_verify_json, plonk_zkey_header  # noqa: F401,E402
