"""Multi-device context (eg_ctx_create_multi) on the CPU-compiled library: the sharding of every host-pointer batch
entry point across child contexts and the in-library combine of the partial tallies (comm.inc / api_multi.inc) must give
the same verdicts and tallies as the oracle, for device counts that do and do not divide the batch (TEST HARNESS ONLY:
the harness has one pretend device, so the "devices" are separate child contexts; the NCCL all-gather itself is covered
by the `-m gpu` tests on a multi-GPU box)."""
import numpy as np
import pytest

import oracle as O
import parity_common as PC
import workloads as W
from elastic_elgamal_b200 import Engine, EngineError, _ffi
from hostsim.build_hostsim import build


@pytest.fixture(scope="module", params=[1, 3])
def env(request):
    e = Engine(lib_path=build(), devices=[0] * request.param)
    sk, pk = W.receiver()
    e.set_receiver(pk)
    assert e.comm_info()["devices"] == request.param
    yield e, sk, pk
    e.close()


def test_multi_verify_bool_and_zero(env):
    PC.check_verify_bool(env[0], env[2], n=24)
    PC.check_verify_zero(env[0], env[2], n=14)


def test_multi_verify_choice_and_tally(env):
    PC.check_verify_choice(env[0], env[2], options=3, n=11, single=True, frac=0.4)
    PC.check_choice_tally_decrypts(env[0], env[2], env[1], options=3, n=7)


def test_multi_fewer_items_than_devices(env):
    """Shards of zero items: verdicts stay consistent and an all-empty shard contributes the identity to the tally."""
    PC.check_verify_choice(env[0], env[2], options=2, n=2, single=True, frac=0.0)
    PC.check_empty_and_tiny(env[0], env[2])


def test_multi_verify_range(env):
    PC.check_verify_range(env[0], env[2], 21, n=7, frac=0.3)


def test_multi_verify_qv(env):
    PC.check_verify_qv(env[0], env[2], env[1], n=5, options=3, credits=9)


def test_multi_shares_and_decrypt(env):
    PC.check_shares_and_decrypt(env[0], n=7, shares=5, threshold=3, used=(0, 2, 4), table_hi=32)


def test_multi_helpers_run_on_first_device(env):
    PC.check_group_helpers(env[0], n=4)
    PC.check_ciphertexts_sum(env[0], env[2])
    PC.check_ciphertext_ops(env[0], env[2], n=5)


def test_multi_knobs_reach_every_child(env):
    """Tuning knobs set on the parent apply to all children: the pair engine pinned on every shard gives the same verdicts."""
    e = env[0]
    e.set_ring_mode(3)
    try:
        PC.check_verify_bool(e, env[2], n=24)
        PC.check_verify_range(e, env[2], 21, n=7, frac=0.3)
    finally:
        e.set_ring_mode(0)


def test_multi_rejects_device_pointer_calls_and_attach(env):
    e = env[0]
    st = e.lib.eg_verify_bool_batch_dev(e.h, 1, 8, 8, 8)
    assert st == _ffi.ERR_INVALID_ARG
    with pytest.raises(EngineError):
        e.attach_comm(np.zeros(128, np.uint8), 0, 2)


def test_multi_launch_counts_accumulate(env):
    e = env[0]
    before = e.kernel_launches
    PC.check_verify_bool(e, env[2], n=24)
    assert e.kernel_launches > before


def test_multi_provers_shard_with_their_own_randomness(env):
    """Provers on a multi-device context: item i keeps block stream i (seeded) / its own slice of the caller's blocks, so the
    outputs are byte-identical to the oracle's whatever the device count."""
    PC.check_seeded_provers(env[0], env[2], env[1], n=7)
    PC.check_encrypt_choice(env[0], env[2], options=3, n=7)
    PC.check_encrypt_range(env[0], env[2], 21, n=5)
    PC.check_encrypt_qv(env[0], env[2], env[1], n=4, options=3, credits=9)
    PC.check_encrypt_plain_and_zero(env[0], env[2], env[1], n=7)
    PC.check_prove_range_from_ciphertext(env[0], env[2], 21, n=5)
