"""`-m gpu` size-independent properties at (a quarter of) the BASELINE.json sizes for the configs the oracle cannot check item
by item in seconds: RangeProof [0, 2^16), QuadraticVotingBallot 5/20 and 3-of-5 decryption shares at 262 144 units each,
plus a 20 000-item differential fuzz of six object kinds under both ring engines.  Properties: a tiled batch gives tiled
verdicts (every tile position is compared with the oracle's verdict of the unique item), tallies decrypt to the sums over
the accepted ballots, results do not depend on the chunking, and what the GPU prover emits at that size the GPU verifier
accepts (bench.py asserts the same for every timed 1 M batch)."""
import random

import numpy as np
import pytest

import oracle as O
import parity_common as PC
import workloads as W

pytestmark = pytest.mark.gpu
N = 1 << 18


@pytest.fixture(scope="module")
def env():
    from elastic_elgamal_b200 import Engine
    e = Engine(device=0)
    sk, pk = W.receiver()
    e.set_receiver(pk)
    yield e, sk, pk
    e.close()


def tile(a, n):
    reps = (n + a.shape[0] - 1) // a.shape[0]
    return np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1))[:n])


def test_range_2_16_at_262k(env):
    e, sk, pk = env
    base = 768
    spec = O.range_optimal(65536)
    espec = PC.to_engine_range(e, spec)
    values = (np.arange(base, dtype=np.uint64) * 40503) % 65536
    cts, partials, rings = O.gen_range_batch(pk, spec, "ciphertext_range", W.SEED_CHOICE, values)
    cts, partials, rings = cts.copy(), partials.copy(), rings.copy()
    PC.tamper_range(cts, partials, rings, random.Random(4), 0.02)
    ov = O.verify_range_batch(pk, spec, "ciphertext_range", cts, partials, rings)
    assert 0 < int((ov != 0).sum()) < base // 10
    n = N + 77                                                  # not a multiple of the tile, the warp or the wave
    v = e.verify_range(espec, "ciphertext_range", tile(cts, n), tile(partials, n), tile(rings, n))
    assert (v == tile(ov, n)).all()
    # chunking invariance on a slice, and prover -> verifier at size with in-kernel randomness
    e.set_chunk_items(5000)
    try:
        v2 = e.verify_range(espec, "ciphertext_range", tile(cts, 23001), tile(partials, 23001), tile(rings, 23001))
    finally:
        e.set_chunk_items(0)
    assert (v2 == tile(ov, 23001)).all()
    vals = (np.arange(1 << 16, dtype=np.uint64) * 7919) % 65536
    c, p, r = e.encrypt_range(espec, "ciphertext_range", vals, seed=bytes(range(32)))
    assert (e.verify_range(espec, "ciphertext_range", c, p, r) == 0).all()
    oc, op, orr = O.gen_range_batch(pk, spec, "ciphertext_range", bytes(range(32)), vals[:64])
    assert (c[:64] == oc).all() and (p[:64] == op).all() and (r[:64] == orr).all()


def test_qv_5_20_at_262k(env):
    e, sk, pk = env
    base = 512
    p, ep = O.qv_params(5, 20), e.qv_params(5, 20)
    votes = np.array([PC.QV_VOTES[i % 4] for i in range(base)], np.uint64)
    ballots = O.gen_qv_batch(pk, p, W.SEED_QV, votes).copy()
    rnd = random.Random(3)
    for k, i in enumerate(sorted(rnd.sample(range(base), base // 50))):
        if k % 3 == 0:
            ballots[i, 32:64] = np.frombuffer(O.point_add(bytes(ballots[i, 32:64]), W.G_ENC), np.uint8)
        elif k % 3 == 1:
            ballots[i, -32 * 12:] = ballots[(i + 1) % base, -32 * 12:]
        else:
            ballots[i, -1] = 0xff
    ov, _ = O.verify_qv_batch(pk, p, ballots)
    assert {int(x) for x in ov} >= {0, 1, O.QV_CREDIT_EQUIV}
    n = N + 13
    v, t = e.verify_qv(ep, tile(ballots, n))
    ev = tile(ov, n)
    assert (v == ev).all()
    table = O.DlogTable(0, 4 * n + 1)
    vt = tile(votes, n)
    for k in range(5):
        assert table.get(O.decrypt_to_element(sk, bytes(t[k]))) == int(vt[ev == 0, k].sum())
    made = e.encrypt_qv(ep, vt[:1 << 15], seed=bytes([3] * 32))
    mv, mt = e.verify_qv(ep, made)
    assert (mv == 0).all()
    for k in range(5):
        assert table.get(O.decrypt_to_element(sk, bytes(mt[k]))) == int(vt[:1 << 15, k].sum())


def test_shares_3_of_5_at_262k(env):
    e, sk, pk = env
    base, hi = 256, 1 << 16
    rng = O.rng_from_seed(bytes([9] * 32))
    ks, secrets = O.dealer_new(5, 3, rng)
    eks = PC.as_engine_keyset(ks)
    used = (0, 2, 4)
    rnd = random.Random(5)
    values = np.array([rnd.randrange(hi) for _ in range(base)], np.uint64)
    cts = [O.encrypt(bytes(ks.shared_key), int(v), rng) for v in values]
    rows = [[O.decrypt_share(ks, i, secrets[i], ct, rng) for i in used] for ct in cts]
    cts_a = np.frombuffer(b"".join(cts), np.uint8).reshape(base, 64).copy()
    sh_a = np.frombuffer(b"".join(b"".join(r[0] for r in row) for row in rows), np.uint8).reshape(base, 3, 32).copy()
    pr_a = np.frombuffer(b"".join(b"".join(r[1] for r in row) for row in rows), np.uint8).reshape(base, 3, 64).copy()
    for k, i in enumerate(sorted(rnd.sample(range(base), 8))):
        pr_a[i, k % 3] = pr_a[(i + 1) % base, k % 3]
    ov = np.array([[O.verify_share(ks, used[j], bytes(cts_a[i]), bytes(sh_a[i, j]), bytes(pr_a[i, j])) for j in range(3)]
                   for i in range(base)], np.uint8)
    assert int((ov != 0).sum()) == 8
    n = N + 5
    C, S, P = tile(cts_a, n), tile(sh_a, n), tile(pr_a, n)
    v = e.verify_shares(eks, list(used), C, S, P)
    assert (v == tile(ov, n)).all()
    table = e.dlog_table(0, hi)
    vals, found = e.combine_decrypt(list(used), C, S, table)
    assert (found == 1).all() and (vals == tile(values, n)).all()
    table.close()


@pytest.mark.parametrize("mode", [2, 1, 3])
def test_fuzz_differential_20k(env, mode):
    """Random bit flips anywhere in the inputs of EncryptedChoice / bool / RangeProof / QV ballot / CommitmentEquivalenceProof /
    ProofOfPossession batches of 20 000 items: every GPU verdict and tally equals the oracle's, under each of the three ring engines."""
    e, sk, pk = env
    e.set_ring_mode(mode)
    try:
        PC.check_fuzz_differential(e, pk, n=20000, seed=200 + mode)
    finally:
        e.set_ring_mode(0)


@pytest.mark.parametrize("n", [1, 31, 33, 127, 129, 148 * 128 - 1, 148 * 128 + 1, 40001])
def test_tally_reduction_shapes(env, n):
    """The masked tally reduction (k_tally_partial: per-thread accumulation, warp-shuffle tree, shared-memory fold, one
    partial per CTA; k_tally_final) at ballot counts that are not multiples of the warp, the CTA or the grid, with chunks
    whose ballots are ALL rejected and chunk boundaries inside the batch: tally == the oracle's."""
    e, sk, pk = env
    base = min(n, 600)
    cts, rings, sums = O.gen_choice_batch(pk, 3, W.SEED_CHOICE, base)
    cts, rings, sums = cts.copy(), rings.copy(), sums.copy()
    rnd = random.Random(n)
    for i in rnd.sample(range(base), max(1, base // 3)) if base > 2 else []:
        sums[i, 0] ^= 1                                          # a wrong sum proof: rejected, must not reach the tally
    if base >= 300:
        sums[90:230, 0] ^= 2                                     # 140 consecutive rejections: whole warps, a whole CTA and (chunk = 97) a whole chunk masked out
    ov, _ = O.verify_choice_batch(pk, 3, True, cts, rings, sums)
    C, R, S, ev = tile(cts, n), tile(rings, n), tile(sums, n), tile(ov, n)
    # expected tally from per-option counts (accepted ballots of option k encrypt 1 there, 0 elsewhere)
    table = O.DlogTable(0, n + 1)
    idx = np.arange(n) % base
    for chunk in (0, 97) if n < 50000 else (0, 14999):
        e.set_chunk_items(chunk)
        e.set_ring_mode(2)
        try:
            v, t = e.verify_choice(3, C, R, S)
        finally:
            e.set_chunk_items(0)
            e.set_ring_mode(0)
        assert (v == ev).all()
        for k in range(3):
            assert table.get(O.decrypt_to_element(sk, bytes(t[k]))) == int(np.count_nonzero((ev == 0) & (idx % 3 == k)))
    if n <= 600:
        _, ot = O.verify_choice_batch(pk, 3, True, C, R, S)
        assert (t == ot).all()
