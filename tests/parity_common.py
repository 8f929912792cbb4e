"""Parity checks shared by the GPU tests (-m gpu, through libeg_b200.so) and the host-simulation tests
(-m "not gpu", the same C ABI compiled for the CPU by tests/hostsim).  Every check compares the engine with the
CPU oracle on the same seeded inputs: bit-exact verdicts, encodings and tallies."""
import random

import numpy as np

import oracle as O
import workloads as W

L = W.L


def sc(x):
    return (x % L).to_bytes(32, "little")


def check_group_helpers(e, n=24):
    rnd = random.Random(11)
    # elements: valid multiples, the RFC 9496 invalid vectors, identity
    valid = [O.point_mul_generator(sc(rnd.randrange(L))) for _ in range(n)] + [bytes(32), W.G_ENC]
    invalid = [W.BAD_POINT, W.BAD_POINT2, b"\xff" * 32,
               bytes.fromhex("edffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f"),
               bytes.fromhex("ecffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f"),
               bytes.fromhex("3eb858e78f5a7254d8c9731174a94f76755fd3941c0ac93735c07ba14579630e")]
    enc = np.frombuffer(b"".join(valid + invalid), np.uint8)
    ok = e.elements_validate(enc)
    assert ok.tolist() == [True] * len(valid) + [False] * len(invalid)
    assert ok.tolist() == [O.point_valid(x) for x in valid + invalid]
    # scalars
    scal = [sc(0), sc(1), sc(L - 1), L.to_bytes(32, "little"), (L + 1).to_bytes(32, "little"), b"\xff" * 32,
            (2**252).to_bytes(32, "little"), (2**253).to_bytes(32, "little")]
    ok = e.scalars_validate(np.frombuffer(b"".join(scal), np.uint8))
    assert ok.tolist() == [O.scalar_is_canonical(x) for x in scal]
    # wide reduction
    wide = [rnd.getrandbits(512).to_bytes(64, "little") for _ in range(n)] + [b"\xff" * 64, bytes(64)]
    out = e.scalars_from_wide(np.frombuffer(b"".join(wide), np.uint8))
    for i, w in enumerate(wide):
        assert bytes(out[i]) == O.scalar_reduce_wide(w)
    # [k]G and [a]A + [b]G
    ks = [sc(rnd.randrange(L)) for _ in range(n)] + [sc(0), sc(1), sc(L - 1), sc(8), sc(2**252)]
    out, ok = e.mul_generator(np.frombuffer(b"".join(ks), np.uint8))
    assert ok.all()
    for i, k in enumerate(ks):
        assert bytes(out[i]) == O.point_mul_generator(k), i
    a = [sc(rnd.randrange(L)) for _ in range(n)] + [sc(0), sc(L - 1), sc(1)]
    b = [sc(rnd.randrange(L)) for _ in range(n)] + [sc(5), sc(0), sc(L - 1)]
    A = [O.point_mul_generator(sc(rnd.randrange(L))) for _ in range(n)] + [bytes(32), W.G_ENC, W.G_ENC]
    out, ok = e.double_mul_generator(np.frombuffer(b"".join(a), np.uint8), np.frombuffer(b"".join(A), np.uint8),
                                     np.frombuffer(b"".join(b), np.uint8))
    assert ok.all()
    for i in range(len(a)):
        expect = O.point_add(O.point_mul(a[i], A[i]), O.point_mul_generator(b[i]))
        assert bytes(out[i]) == expect, i
    # malformed inputs are reported, not computed
    out, ok = e.double_mul_generator(np.frombuffer(sc(1) + L.to_bytes(32, "little"), np.uint8),
                                     np.frombuffer(W.BAD_POINT + W.G_ENC, np.uint8), np.frombuffer(sc(1) + sc(1), np.uint8))
    assert ok.tolist() == [False, False]


def check_ciphertexts_sum(e, pk):
    rng = O.rng_from_seed(bytes([3] * 32))
    parts = np.frombuffer(b"".join(O.encrypt(pk, v, rng) for v in range(12)), np.uint8).reshape(4, 3, 64)
    out, ok = e.ciphertexts_sum(parts)
    assert ok
    for c in range(3):
        r = b = bytes(32)
        for p in range(4):
            r = O.point_add(r, bytes(parts[p, c, :32]))
            b = O.point_add(b, bytes(parts[p, c, 32:]))
        assert bytes(out[c]) == r + b
    bad = parts.copy()
    bad[1, 1, :32] = np.frombuffer(W.BAD_POINT, np.uint8)
    assert not e.ciphertexts_sum(bad)[1]


def check_verify_zero(e, pk, n=40):
    rng = O.rng_from_seed(bytes([4] * 32))
    items = [O.encrypt_zero(pk, rng) for _ in range(n)]
    cts = np.frombuffer(b"".join(i[0] for i in items), np.uint8).reshape(n, 64).copy()
    proofs = np.frombuffer(b"".join(i[1] for i in items), np.uint8).reshape(n, 64).copy()
    cts[3, 32:] = np.frombuffer(O.point_add(bytes(cts[3, 32:]), W.G_ENC), np.uint8)
    proofs[5] = proofs[6]
    proofs[7, 32:] = np.frombuffer(W.BAD_SCALAR, np.uint8)
    cts[9, :32] = np.frombuffer(W.BAD_POINT, np.uint8)
    proofs[11, 1] ^= 4
    expected = [O.verify_zero(pk, bytes(cts[i]), bytes(proofs[i])) for i in range(n)]
    got = e.verify_zero(cts, proofs)
    assert got.tolist() == expected
    assert expected[3] == O.CHALLENGE_MISMATCH and expected[7] == O.MALFORMED and expected[9] == O.MALFORMED and expected[0] == O.OK


def check_verify_bool(e, pk, n=100, seed=21):
    cts, proofs = O.gen_bool_batch(pk, W.SEED_CHOICE, n)
    cts, proofs = cts.copy(), proofs.copy()
    tampered = W.tamper_bool(cts, proofs, random.Random(seed), frac=0.25)
    expected = O.verify_bool_batch(pk, cts, proofs)
    got = e.verify_bool(cts, proofs)
    assert got.tolist() == expected.tolist()
    assert (expected[tampered] != 0).all() and (np.delete(expected, tampered) == 0).all()
    assert set(expected.tolist()) >= {O.OK, O.MALFORMED, O.CHALLENGE_MISMATCH}


def check_verify_choice(e, pk, options=5, n=60, single=True, seed=22, frac=0.3):
    if single:
        cts, rings, sums = O.gen_choice_batch(pk, options, W.SEED_CHOICE, n)
    else:
        rng = O.rng_from_seed(bytes([8] * 32))
        rnd = random.Random(seed)
        items = [O.choice_new(pk, [rnd.random() < 0.5 for _ in range(options)], False, rng) for _ in range(n)]
        cts = np.frombuffer(b"".join(i[0] for i in items), np.uint8).reshape(n, options, 64)
        rings = np.frombuffer(b"".join(i[1] for i in items), np.uint8).reshape(n, 1 + 2 * options, 32)
        sums = None
    cts, rings = cts.copy(), rings.copy()
    sums = sums.copy() if sums is not None else None
    tampered = W.tamper_choice(cts, rings, sums, random.Random(seed), frac=frac) if frac else []
    expected, exp_tally = O.verify_choice_batch(pk, options, single, cts, rings, sums)
    got, tally = e.verify_choice(options, cts, rings, sums, single=single, tally=True)
    assert got.tolist() == expected.tolist()
    assert (tally == exp_tally).all()
    if frac:
        assert (expected[tampered] != 0).all() and (np.delete(expected, tampered) == 0).all()
    return expected


def check_choice_tally_decrypts(e, pk, sk, options=5, n=50):
    """encode -> verify -> tally -> decrypt round trip: every option must count its voters (voting.rs:122-177)."""
    cts, rings, sums = O.gen_choice_batch(pk, options, W.SEED_CHOICE, n)
    verdicts, tally = e.verify_choice(options, cts, rings, sums, single=True, tally=True)
    assert (verdicts == 0).all()
    table = O.DlogTable(0, n + 1)
    counts = [table.get(O.decrypt_to_element(sk, bytes(tally[k]))) for k in range(options)]
    assert counts == [len(range(k, n, options)) for k in range(options)]


def check_empty_and_tiny(e, pk):
    z = np.zeros((0, 64), np.uint8)
    assert e.verify_bool(z, np.zeros((0, 96), np.uint8)).shape == (0,)
    v, t = e.verify_choice(3, np.zeros((0, 3, 64), np.uint8), np.zeros((0, 7, 32), np.uint8), np.zeros((0, 64), np.uint8))
    assert v.shape == (0,) and not t.any()          # empty tally = identity ciphertexts (all-zero encodings)
    for n in (1, 2, 31, 33):
        check_verify_choice(e, pk, options=2, n=n, frac=0)
