"""`-m gpu` parity tests proper: the CUDA library (libeg_b200.so, through its C ABI) against the CPU oracle on
the same seeded inputs -- bit-exact verdicts, encodings and tallies -- plus size-independent properties at
BASELINE.json sizes (tally round trip, chunking invariance)."""
import random

import numpy as np
import pytest

import oracle as O
import parity_common as PC
import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from elastic_elgamal_b200 import Engine
    e = Engine(device=0)           # fails loudly without a GPU / without the built extension
    sk, pk = W.receiver()
    e.set_receiver(pk)
    e.set_ring_mode(2)             # the batches below are small: pin the large-batch engine (k_ring), which the library's
    yield e, sk, pk                # automatic choice would reserve for chunks of >= 2 waves (test_ring_mode_auto)
    e.close()


def test_field_selftest(env):
    assert env[0].selftest_field(n=1 << 20, seed=7) == 0


def test_group_helpers(env):
    PC.check_group_helpers(env[0], n=300)


def test_ciphertexts_sum(env):
    PC.check_ciphertexts_sum(env[0], env[2])


def test_verify_zero(env):
    PC.check_verify_zero(env[0], env[2], n=200)


def test_verify_bool_config1_shape(env):
    # BASELINE config 1 (benches/basics.rs:60-82 shape) at a size the oracle checks in seconds
    PC.check_verify_bool(env[0], env[2], n=2000, seed=31)


@pytest.mark.parametrize("options", [2, 3, 5, 10, 15])          # tests/integration/basic.rs:164
def test_verify_choice_sizes(env, options):
    PC.check_verify_choice(env[0], env[2], options=options, n=300, single=True, frac=0.2)


def test_verify_choice_multi(env):
    PC.check_verify_choice(env[0], env[2], options=5, n=300, single=False, frac=0.2)


def test_verify_choice_headline_shape(env):
    v = PC.check_verify_choice(env[0], env[2], options=5, n=4000, single=True, frac=0.01)
    assert set(v.tolist()) >= {O.OK}


def test_choice_tally_round_trip(env):
    PC.check_choice_tally_decrypts(env[0], env[2], env[1], options=5, n=1000)


def test_empty_and_ragged(env):
    PC.check_empty_and_tiny(env[0], env[2])


def test_chunking_invariance(env):
    e, sk, pk = env
    cts, rings, sums = O.gen_choice_batch(pk, 5, W.SEED_CHOICE, 777)
    cts, rings, sums = cts.copy(), rings.copy(), sums.copy()
    W.tamper_choice(cts, rings, sums, random.Random(5), frac=0.05)
    base_v, base_t = e.verify_choice(5, cts, rings, sums)
    for chunk in (100, 256, 1000):
        e.set_chunk_items(chunk)
        v, t = e.verify_choice(5, cts, rings, sums)
        assert (v == base_v).all() and (t == base_t).all()
    e.set_chunk_items(0)
    ov, ot = O.verify_choice_batch(pk, 5, True, cts, rings, sums)
    assert (base_v == ov).all() and (base_t == ot).all()


def test_large_batch_properties(env):
    """At 256k ballots the oracle is too slow to check every item; use size-independent properties: a tiled batch
    must give tiled verdicts, and its tally must decrypt to the per-option counts of the accepted ballots."""
    e, sk, pk = env
    base = 4096
    cts, rings, sums = O.gen_choice_batch(pk, 5, W.SEED_CHOICE, base)
    cts, rings, sums = cts.copy(), rings.copy(), sums.copy()
    W.tamper_choice(cts, rings, sums, random.Random(9), frac=0.01)
    ov, _ = O.verify_choice_batch(pk, 5, True, cts, rings, sums)
    reps = 64
    big = (np.tile(cts, (reps, 1, 1)), np.tile(rings, (reps, 1, 1)), np.tile(sums, (reps, 1)))
    v, t = e.verify_choice(5, *big)
    assert (v.reshape(reps, base) == ov[None, :]).all()
    table = O.DlogTable(0, reps * base + 1)
    accepted = (ov == 0)
    for k in range(5):
        expect = reps * int(np.count_nonzero(accepted[k::5]))
        assert table.get(O.decrypt_to_element(sk, bytes(t[k]))) == expect


def test_range_decomposition(env):
    PC.check_range_decomposition(env[0])


@pytest.mark.parametrize("ub", [2, 5, 16, 21, 100, 1000, 65536])      # 65536 = BASELINE config 4 (8 rings x 4)
def test_verify_range(env, ub):
    PC.check_verify_range(env[0], env[2], ub, n=120 if ub < 65536 else 64, frac=0.2)


def test_verify_qv_config3_shape(env):
    PC.check_verify_qv(env[0], env[2], env[1], n=96)                   # 5 options / 20 credits


def test_verify_qv_other_shapes(env):
    PC.check_verify_qv(env[0], env[2], env[1], n=16, options=3, credits=15)
    PC.check_verify_qv(env[0], env[2], env[1], n=16, options=8, credits=30)


def test_shares_config5_shape(env):
    PC.check_shares_and_decrypt(env[0], n=200, shares=5, threshold=3, used=(0, 2, 4), table_hi=1 << 12)


def test_shares_with_key_tables(env):
    """PublicKeySet::verify_share / CandidateDecryption::verify with per-call fixed-base tables for the participant keys
    (the library's choice from 32 768 tallies per call on; forced here): same verdicts and decrypted values as the oracle, for
    several key sets in a row on one context (the table cache is keyed by the key bytes), then back on the default path."""
    e = env[0]
    e.set_key_table_min(0)
    try:
        PC.check_shares_and_decrypt(e, n=200, shares=5, threshold=3, used=(0, 2, 4), table_hi=1 << 12)
        PC.check_shares_and_decrypt(e, n=16, shares=10, threshold=7, used=(1, 2, 3, 5, 6, 8, 9), table_hi=100)
        PC.check_shares_and_decrypt(e, n=33, shares=5, threshold=3, used=(4, 1, 0), table_hi=64, seed=8)
        PC.check_shares_and_decrypt(e, n=200, shares=5, threshold=3, used=(0, 2, 4), table_hi=1 << 12)
        PC.check_verify_decryption(e, n=40)
    finally:
        e.set_key_table_min(32768)
    PC.check_shares_and_decrypt(e, n=16, shares=5, threshold=3, used=(0, 2, 4), table_hi=64)


def test_shares_other_params(env):
    PC.check_shares_and_decrypt(env[0], n=16, shares=10, threshold=7, used=(1, 2, 3, 5, 6, 8, 9), table_hi=100)
    PC.check_shares_and_decrypt(env[0], n=16, shares=2, threshold=1, used=(1,), table_hi=100)


def test_dlog_table_large(env):
    e, sk, pk = env
    table = e.dlog_table(0, 1 << 20)
    rng = O.rng_from_seed(bytes([9] * 32))
    ks, secrets = O.dealer_new(3, 2, rng)
    shared = bytes(ks.shared_key)
    values = [0, 1, (1 << 20) - 1, 1 << 20, 777777, 123456]
    cts = [O.encrypt(shared, v, rng) for v in values]
    sh = [[O.decrypt_share(ks, i, secrets[i], ct, rng)[0] for i in (0, 2)] for ct in cts]
    vals, found = e.combine_decrypt([0, 2], np.frombuffer(b"".join(cts), np.uint8), np.frombuffer(b"".join(b"".join(r) for r in sh), np.uint8), table)
    assert found.tolist() == [1, 1, 1, 0, 1, 1] and [int(v) for v in vals[[0, 1, 2, 4, 5]]] == [0, 1, (1 << 20) - 1, 777777, 123456]
    table.close()


def test_ring_mode_1_still_matches(env):
    """The per-equation launch pipeline (ring mode 1) stays available as an A/B knob and must agree with the oracle."""
    e, sk, pk = env
    e.set_ring_mode(1)
    try:
        PC.check_verify_bool(e, pk, n=200, seed=5)
        PC.check_verify_choice(e, pk, options=5, n=200, single=True, frac=0.2)
        PC.check_verify_range(e, pk, 100, n=60, frac=0.2)
    finally:
        e.set_ring_mode(2)


def test_ring_mode_3_pair_engine(env):
    """The pair engine (ring mode 3: lanes 2p / 2p + 1 evaluate the two sides of ring p and swap their encodings by shuffle)
    agrees with the oracle on every ring shape: bool, choice (terminal points + sum proof), multi-choice, range proofs with
    rings of 2 to 16 equations (pairs of one warp in different equations), QV ballots, the reference's tamper patterns."""
    e, sk, pk = env
    e.set_ring_mode(3)
    try:
        PC.check_verify_bool(e, pk, n=333, seed=7)
        PC.check_verify_bool(e, pk, n=31, seed=8)          # a partial warp: 62 lanes, the last pair next to exited lanes
        PC.check_verify_choice(e, pk, options=5, n=203, single=True, frac=0.2)
        PC.check_verify_choice(e, pk, options=3, n=37, single=False, frac=0.3)
        for bound in (2, 5, 100, 1000, 65536):
            PC.check_verify_range(e, pk, bound, n=45, frac=0.25)
        PC.check_verify_qv(e, pk, sk, n=12)
        PC.check_fuzz_differential(e, pk, n=600, seed=303)
    finally:
        e.set_ring_mode(2)


def test_ring_mode_auto(env):
    """Default engine choice: small chunks go through the pair engine (rings only) or the per-equation pipeline (choice
    ballots: rings + sum proof), large ones through k_ring; a batch
    that mixes both (a 30 310-ballot ramp-up chunk and a full 257 638-ballot chunk through k_ring + a 5 052-ballot remainder,
    a third of a wave, through the pipeline) must still give tiled verdicts and the tally of the accepted ballots."""
    e, sk, pk = env
    e.set_ring_mode(0)
    try:
        PC.check_verify_bool(e, pk, n=300, seed=6)
        PC.check_verify_choice(e, pk, options=5, n=256, single=True, frac=0.1)
        PC.check_verify_range(e, pk, 1000, n=50, frac=0.2)
        PC.check_verify_qv(e, pk, sk, n=12)
        base, total = 1024, 293000
        cts, rings, sums = O.gen_choice_batch(pk, 5, W.SEED_CHOICE, base)
        cts, rings, sums = cts.copy(), rings.copy(), sums.copy()
        W.tamper_choice(cts, rings, sums, random.Random(10), frac=0.02)
        ov, _ = O.verify_choice_batch(pk, 5, True, cts, rings, sums)
        reps = (total + base - 1) // base
        big = [np.ascontiguousarray(np.tile(x, (reps,) + (1,) * (x.ndim - 1))[:total]) for x in (cts, rings, sums)]
        v, t = e.verify_choice(5, *big)
        exp = np.tile(ov, reps)[:total]
        assert (v == exp).all()
        table = O.DlogTable(0, total + 1)
        choice = np.arange(total) % base % 5            # gen_choice_batch: item i votes for option i mod 5
        for k in range(5):
            assert table.get(O.decrypt_to_element(sk, bytes(t[k]))) == int(np.count_nonzero((exp == 0) & (choice == k)))
    finally:
        e.set_ring_mode(2)


def test_encrypt_bool_config1_shape(env):
    PC.check_encrypt_bool(env[0], env[2], n=500)


def test_encrypt_choice(env):
    PC.check_encrypt_choice(env[0], env[2], options=5, n=300)
    PC.check_encrypt_choice(env[0], env[2], options=2, n=50)
    PC.check_encrypt_multi_choice(env[0], env[2], options=6, n=100)


def test_verifiers_chunk_pipeline(env):
    PC.check_verifiers_chunked(env[0], env[2], chunk=41, n=400)


def test_verify_range_several_chunks(env):
    """2600 range proofs with a 1024-proof chunk: three chunks through the copy / compute pipeline."""
    e, sk, pk = env
    e.set_chunk_items(100)          # verify_range clamps its chunk to >= 1024 proofs
    try:
        PC.check_verify_range(e, pk, 21, n=2600, frac=0.05)
        PC.check_verify_qv(e, pk, sk, n=2300)          # verify_qv clamps to >= 1024 ballots per chunk as well
    finally:
        e.set_chunk_items(0)


def test_provers_chunk_pipeline(env):
    PC.check_provers_chunked(env[0], env[2], chunk=37, n=300)


def test_identity_commitments(env):
    PC.check_identity_commitments(env[0], env[2], options=5, n=200)
    PC.check_identity_commitments(env[0], env[2], options=1, n=20)


def test_encrypt_reference_snapshots(env):
    PC.check_encrypt_against_reference_snapshots(env[0])


def test_encrypt_verify_round_trip_large(env):
    """Size-independent property at a size the oracle does not check item by item: everything the GPU prover emits from
    arbitrary randomness is accepted by the GPU verifier, and the tally decrypts to the per-option counts."""
    e, sk, pk = env
    n, m = 20000, 5
    rs = np.random.RandomState(3)
    values = np.zeros((n, m), np.uint8)
    values[np.arange(n), rs.randint(0, m, n)] = 1
    wide = rs.randint(0, 256, (n, 3 * m + 1, 64)).astype(np.uint8)
    cts, rings, sums = e.encrypt_choice(m, values, wide, single=True)
    v, t = e.verify_choice(m, cts, rings, sums)
    assert (v == 0).all()
    table = O.DlogTable(0, n + 1)
    assert [table.get(O.decrypt_to_element(sk, bytes(t[k]))) for k in range(m)] == values.sum(axis=0).tolist()


def test_commitment_equivalence(env):
    PC.check_commitment_equiv(env[0], env[2], n=400)


def test_commitment_equivalence_reference_snapshot(env):
    try:
        PC.check_commitment_equiv_snapshot(env[0])
    finally:
        env[0].set_receiver(env[2])


@pytest.mark.parametrize("k", [1, 2, 5, 41, 64])
def test_proof_of_possession(env, k):
    PC.check_possession(env[0], n=60 if k <= 5 else 12, keys_per_proof=k)


def test_base64url_wire_format(env):
    PC.check_base64url(env[0], n=3000)


@pytest.mark.parametrize("ub", [2, 5, 16, 21, 100, 1000, 65536])
def test_encrypt_range(env, ub):
    PC.check_encrypt_range(env[0], env[2], ub, n=40)


def test_encrypt_range_reference_snapshot(env):
    try:
        PC.check_encrypt_range_reference_snapshot(env[0])
    finally:
        env[0].set_receiver(env[2])


def test_encrypt_qv(env):
    PC.check_encrypt_qv(env[0], env[2], env[1], n=48)
    PC.check_encrypt_qv(env[0], env[2], env[1], n=12, options=3, credits=15)


def test_encrypt_qv_reference_snapshot(env):
    try:
        PC.check_encrypt_qv_reference_snapshot(env[0])
    finally:
        env[0].set_receiver(env[2])


def test_encrypt_plain_and_zero(env):
    PC.check_encrypt_plain_and_zero(env[0], env[2], env[1], n=64)
    try:
        PC.check_encrypt_plain_and_zero_snapshots(env[0])
    finally:
        env[0].set_receiver(env[2])


def test_ciphertext_ops(env):
    PC.check_ciphertext_ops(env[0], env[2], n=64)


def test_multi_mul(env):
    PC.check_multi_mul(env[0], n=64)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz_differential(env, seed):
    PC.check_fuzz_differential(env[0], env[2], n=400, seed=seed)


@pytest.mark.parametrize("shares,threshold", [(5, 3), (10, 7), (4, 4), (3, 1), (6, 2), (40, 16), (64, 1)])
def test_keysets_validate(env, shares, threshold):
    PC.check_keysets_validate(env[0], n_sets=40, shares=shares, threshold=threshold)


def test_device_pointer_forms_match_host_forms(env):
    """eg_*_batch_dev (inputs already in HBM) give the same bytes as the host-pointer forms: QV + tally, decryption shares,
    combine + decrypt (range / bool / choice are exercised by bench.py every round)."""
    import ctypes as C

    import torch
    e, sk, pk = env
    dev = torch.device("cuda", 0)

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    # QuadraticVotingBallot::verify
    p, ep = O.qv_params(5, 20), e.qv_params(5, 20)
    n = 40
    votes = np.array([PC.QV_VOTES[i % 4] for i in range(n)], np.uint64)
    ballots = O.gen_qv_batch(pk, p, W.SEED_QV, votes).copy()
    ballots[3, -1] = 0xff
    ballots[5, 32:64] = np.frombuffer(O.point_add(bytes(ballots[5, 32:64]), W.G_ENC), np.uint8)
    hv, ht = e.verify_qv(ep, ballots)
    d_b, d_v, d_t = up(ballots), torch.empty(n, dtype=torch.uint8, device=dev), torch.empty((5, 64), dtype=torch.uint8, device=dev)
    e._check(e.lib.eg_verify_qv_batch_dev(e.h, C.byref(ep), n, d_b.data_ptr(), d_v.data_ptr(), d_t.data_ptr()))
    assert (d_v.cpu().numpy() == hv).all() and (d_t.cpu().numpy() == ht).all()
    ov, ot = O.verify_qv_batch(pk, p, ballots)
    assert (hv == ov).all() and (ht == ot).all()
    # verify_share + combine_shares + decrypt
    rng = O.rng_from_seed(bytes([9] * 32))
    ks, secrets = O.dealer_new(5, 3, rng)
    eks = PC.as_engine_keyset(ks)
    used, n, hi = (0, 2, 4), 50, 128
    rnd = random.Random(11)
    values = [rnd.randrange(hi + 10) for _ in range(n)]
    cts = [O.encrypt(bytes(ks.shared_key), v, rng) for v in values]
    rows = [[O.decrypt_share(ks, i, secrets[i], ct, rng) for i in used] for ct in cts]
    cts_a = np.frombuffer(b"".join(cts), np.uint8).reshape(n, 64).copy()
    sh_a = np.frombuffer(b"".join(b"".join(r[0] for r in row) for row in rows), np.uint8).reshape(n, 3, 32).copy()
    pr_a = np.frombuffer(b"".join(b"".join(r[1] for r in row) for row in rows), np.uint8).reshape(n, 3, 64).copy()
    pr_a[1, 0] = pr_a[2, 0]
    sh_a[4, 2] = np.frombuffer(W.BAD_POINT, np.uint8)
    hv = e.verify_shares(eks, list(used), cts_a, sh_a, pr_a)
    table = e.dlog_table(0, hi)
    hvals, hfound = e.combine_decrypt(list(used), cts_a, sh_a, table)
    idx = (C.c_uint32 * 3)(*used)
    d_c, d_s, d_p = up(cts_a), up(sh_a), up(pr_a)
    d_v = torch.empty((n, 3), dtype=torch.uint8, device=dev)
    d_vals, d_found = torch.empty(n, dtype=torch.int64, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    e._check(e.lib.eg_verify_shares_batch_dev(e.h, C.byref(eks), n, 3, idx, d_c.data_ptr(), d_s.data_ptr(), d_p.data_ptr(), d_v.data_ptr()))
    e._check(e.lib.eg_combine_decrypt_batch_dev(e.h, 3, idx, n, 3, d_c.data_ptr(), d_s.data_ptr(), table.h, d_vals.data_ptr(), d_found.data_ptr()))
    assert (d_v.cpu().numpy() == hv).all()
    f = d_found.cpu().numpy()
    assert (f == hfound).all() and (d_vals.cpu().numpy().view(np.uint64)[f == 1] == hvals[hfound == 1]).all()
    assert hv[1, 0] == O.CHALLENGE_MISMATCH and hv[4, 2] == O.MALFORMED and hfound[4] == 2
    table.close()


def test_seeded_provers(env):
    PC.check_seeded_provers(env[0], env[2], env[1], n=300)


def test_seeded_reference_snapshots(env):
    PC.check_seeded_reference_snapshots(env[0])


def test_constant_time_prover_mode(env):
    PC.check_constant_time_prover_mode(env[0], env[2], n=6)


def test_single_choice_validation(env):
    PC.check_single_choice_validation(env[0], env[2])


@pytest.mark.parametrize("count", [1, 5, 14])            # (the oracle restates up to 14 terms)
def test_verify_sumsq(env, count):
    PC.check_verify_sumsq(env[0], env[2], n=60, count=count)


def test_verify_sumsq_reference_snapshot(env):
    PC.check_verify_sumsq_reference_snapshot(env[0])


def test_verify_decryption_custom_key(env):
    PC.check_verify_decryption(env[0], n=200)


def test_wire_objects(env):
    PC.check_wire_objects(env[0], env[2])


@pytest.mark.parametrize("ub", [2, 100, 65536])
def test_prove_range_from_ciphertext(env, ub):
    PC.check_prove_range_from_ciphertext(env[0], env[2], ub, n=40)


def test_prover_dev_forms(env):
    import torch
    dev = torch.device("cuda", 0)

    class Buf:
        def __init__(self, t):
            self.t = t
            self.ptr = t.data_ptr()
    PC.check_prover_dev_forms(env[0], env[2], lambda a: Buf(torch.from_numpy(np.ascontiguousarray(a)).to(dev)),
                              lambda b: b.t.cpu().numpy(), lambda shape: Buf(torch.zeros(shape, dtype=torch.uint8, device=dev)), n=200)
