"""CPU oracle for the cocg hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (Python big-int + a C library under
``oracle/c``) of the reference's Groth16 / REP3 / MSM / NTT path.  It exists so
that the CUDA product path in ``collaborative-circom_b200`` can be checked
bit-for-bit.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product path
never does (and fails loudly when its CUDA library is missing).

Parity pinning: the reference (pure Rust, arkworks 0.4.x un-vendored, no
cargo in this image) cannot be compiled here, so the oracle is pinned against
the reference's own fixtures and known-answer tests instead -- see
``tests/test_oracle_pinned.py`` and DESIGN.md section "Oracle".
"""
