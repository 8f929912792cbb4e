"""`-m "not gpu"` coverage of the product's host orchestration and kernel bodies: the C ABI library compiled for
the CPU by tests/hostsim (TEST HARNESS ONLY -- it is not a product path) must agree with the oracle bit for bit.
The same checks run against the real CUDA library in tests/test_gpu_parity.py."""
import pytest

import oracle as O
import parity_common as PC
import workloads as W
from elastic_elgamal_b200 import Engine
from hostsim.build_hostsim import build


@pytest.fixture(scope="module")
def env():
    e = Engine(lib_path=build())
    sk, pk = W.receiver()
    e.set_receiver(pk)
    yield e, sk, pk
    e.close()


def test_group_helpers(env):
    PC.check_group_helpers(env[0], n=6)


def test_ciphertexts_sum(env):
    PC.check_ciphertexts_sum(env[0], env[2])


def test_verify_zero(env):
    PC.check_verify_zero(env[0], env[2], n=14)


def test_verify_bool(env):
    PC.check_verify_bool(env[0], env[2], n=24)


def test_verify_choice_single(env):
    PC.check_verify_choice(env[0], env[2], options=5, n=16, single=True, frac=0.5)


def test_verify_choice_multi(env):
    PC.check_verify_choice(env[0], env[2], options=3, n=8, single=False, frac=0.5)


def test_per_equation_pipeline(env):
    """Ring mode 1 (what the library picks for small chunks on the GPU) on the CPU-compiled bodies."""
    e = env[0]
    e.set_ring_mode(1)
    try:
        PC.check_verify_bool(e, env[2], n=24)
        PC.check_verify_choice(e, env[2], options=3, n=8, single=True, frac=0.5)
        PC.check_verify_range(e, env[2], 21, n=6, frac=0.3)
    finally:
        e.set_ring_mode(0)


def test_pair_engine(env):
    """Ring mode 3 (two lanes per ring, for small chunks): the per-side evaluation, single-point encoding and transcript
    hand-over of the CUDA kernel, run lane after lane on the CPU-compiled bodies -- bool, choice (deferred terminal
    encodings + sum proof), range proofs with rings of different lengths, QV ballots."""
    e = env[0]
    e.set_ring_mode(3)
    try:
        PC.check_verify_bool(e, env[2], n=24)
        PC.check_verify_choice(e, env[2], options=3, n=8, single=True, frac=0.5)
        PC.check_verify_choice(e, env[2], options=2, n=6, single=False, frac=0.5)
        PC.check_verify_range(e, env[2], 21, n=6, frac=0.3)
        PC.check_verify_range(e, env[2], 100, n=4, frac=0.5)
    finally:
        e.set_ring_mode(0)


def test_choice_tally_round_trip(env):
    PC.check_choice_tally_decrypts(env[0], env[2], env[1], options=3, n=7)


def test_empty_and_tiny(env):
    e, sk, pk = env
    PC.check_empty_and_tiny(e, pk)


def test_receiver_validation(env):
    e, sk, pk = env
    from elastic_elgamal_b200 import EngineError
    from elastic_elgamal_b200 import _ffi
    with pytest.raises(EngineError) as ei:
        e.set_receiver(bytes(32))                       # keys/mod.rs:168-169 IdentityKey
    assert ei.value.status == _ffi.ERR_IDENTITY_KEY
    with pytest.raises(EngineError) as ei:
        e.set_receiver(W.BAD_POINT)                     # keys/mod.rs:166-167 InvalidGroupElement
    assert ei.value.status == _ffi.ERR_INVALID_ELEMENT
    with pytest.raises(EngineError) as ei:
        import numpy as np
        e.verify_bool(np.zeros((1, 64), np.uint8), np.zeros((1, 96), np.uint8))
    assert ei.value.status == _ffi.ERR_NO_RECEIVER
    e.set_receiver(pk)


def test_range_decomposition(env):
    PC.check_range_decomposition(env[0])


@pytest.mark.parametrize("ub", [2, 5, 21, 100])
def test_verify_range(env, ub):
    PC.check_verify_range(env[0], env[2], ub, n=8, frac=0.5)


def test_verify_qv(env):
    PC.check_verify_qv(env[0], env[2], env[1], n=8)


def test_verify_qv_other_shape(env):
    PC.check_verify_qv(env[0], env[2], env[1], n=3, options=3, credits=15)


def test_shares_and_decrypt(env):
    PC.check_shares_and_decrypt(env[0], n=8)


def test_encrypt_bool(env):
    PC.check_encrypt_bool(env[0], env[2], n=6)


def test_encrypt_choice(env):
    PC.check_encrypt_choice(env[0], env[2], options=3, n=5)
    PC.check_encrypt_multi_choice(env[0], env[2], options=3, n=4)


def test_verifiers_chunk_pipeline(env):
    PC.check_verifiers_chunked(env[0], env[2], chunk=5, n=24)


def test_provers_chunk_pipeline(env):
    PC.check_provers_chunked(env[0], env[2], chunk=3, n=8)


def test_identity_commitments(env):
    PC.check_identity_commitments(env[0], env[2], options=3, n=12)
    PC.check_identity_commitments(env[0], env[2], options=1, n=4)


def test_encrypt_reference_snapshots(env):
    PC.check_encrypt_against_reference_snapshots(env[0])


def test_commitment_equivalence(env):
    PC.check_commitment_equiv(env[0], env[2], n=12)


def test_commitment_equivalence_reference_snapshot(env):
    try:
        PC.check_commitment_equiv_snapshot(env[0])
    finally:
        env[0].set_receiver(env[2])


@pytest.mark.parametrize("k", [1, 5])
def test_proof_of_possession(env, k):
    PC.check_possession(env[0], n=8, keys_per_proof=k)


def test_base64url_wire_format(env):
    PC.check_base64url(env[0], n=20)


@pytest.mark.parametrize("ub", [2, 5, 21, 100])
def test_encrypt_range(env, ub):
    PC.check_encrypt_range(env[0], env[2], ub, n=6)


def test_encrypt_range_reference_snapshot(env):
    try:
        PC.check_encrypt_range_reference_snapshot(env[0])
    finally:
        env[0].set_receiver(env[2])


def test_encrypt_qv(env):
    PC.check_encrypt_qv(env[0], env[2], env[1], n=4)
    PC.check_encrypt_qv(env[0], env[2], env[1], n=2, options=3, credits=15)


def test_encrypt_qv_reference_snapshot(env):
    try:
        PC.check_encrypt_qv_reference_snapshot(env[0])
    finally:
        env[0].set_receiver(env[2])


def test_encrypt_plain_and_zero(env):
    PC.check_encrypt_plain_and_zero(env[0], env[2], env[1], n=8)
    try:
        PC.check_encrypt_plain_and_zero_snapshots(env[0])
    finally:
        env[0].set_receiver(env[2])


def test_ciphertext_ops(env):
    PC.check_ciphertext_ops(env[0], env[2], n=6)


def test_multi_mul(env):
    PC.check_multi_mul(env[0], n=6)


@pytest.mark.parametrize("seed", [1])
def test_fuzz_differential(env, seed):
    PC.check_fuzz_differential(env[0], env[2], n=8, seed=seed)


@pytest.mark.parametrize("shares,threshold", [(5, 3), (4, 4), (3, 1)])
def test_keysets_validate(env, shares, threshold):
    PC.check_keysets_validate(env[0], n_sets=6, shares=shares, threshold=threshold)


def test_seeded_provers(env):
    PC.check_seeded_provers(env[0], env[2], env[1], n=6)


def test_seeded_reference_snapshots(env):
    PC.check_seeded_reference_snapshots(env[0])


def test_constant_time_prover_mode(env):
    PC.check_constant_time_prover_mode(env[0], env[2], n=4)


def test_single_choice_validation(env):
    PC.check_single_choice_validation(env[0], env[2])


def test_verify_sumsq(env):
    PC.check_verify_sumsq(env[0], env[2], n=9, count=3)


def test_verify_sumsq_reference_snapshot(env):
    PC.check_verify_sumsq_reference_snapshot(env[0])


def test_verify_decryption_custom_key(env):
    PC.check_verify_decryption(env[0], n=9)


def test_key_tables_forced(env):
    """The per-call fixed-base tables of the keys shares are checked against (default: from 32 768 tallies per call on),
    forced for tiny batches: one table for the custom-key form (the CPU build of a 48 MiB table takes seconds, hence one)."""
    e = env[0]
    e.set_key_table_min(0)
    try:
        PC.check_verify_decryption(e, n=9)
    finally:
        e.set_key_table_min(32768)


def test_wire_objects(env):
    PC.check_wire_objects(env[0], env[2])


@pytest.mark.parametrize("ub", [5, 100])
def test_prove_range_from_ciphertext(env, ub):
    PC.check_prove_range_from_ciphertext(env[0], env[2], ub, n=5)


def test_prover_dev_forms(env):
    import numpy as np

    class Buf:          # the harness has no device: "device" buffers are host arrays
        def __init__(self, a):
            self.a = np.ascontiguousarray(a)
            self.ptr = self.a.ctypes.data
    PC.check_prover_dev_forms(env[0], env[2], lambda a: Buf(a), lambda b: b.a, lambda shape: Buf(np.zeros(shape, np.uint8)), n=6)
