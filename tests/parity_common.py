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


def wide_window_edge_scalars():
    """Scalars that sit on the digit boundaries of the signed fixed-base windows (ge.cuh sc_wide_digit: W-bit windows,
    digits in [-2^(W-1), 2^(W-1)), carries rippling through runs of all-ones / 0x7f..f / 0x80..0 chunks) for every window
    width the library is built with (24 on the device, 16 in the CPU harness, 20 as a build option), all canonical (< l)."""
    out = []
    for w in (16, 20, 24):
        top, ones = 1 << (w - 1), (1 << w) - 1
        for chunk in (top, top - 1, ones, top + 1, 1, ones - 1):
            x = sum(chunk << (w * i) for i in range(256 // w + 1)) % L
            out += [sc(x), sc(L - x)]
        for k in (w - 1, w, w + 1, 2 * w - 1, 2 * w, 3 * w - 1, 3 * w, 5 * w, 240 - w, 240 - 1, 240):
            out += [sc(2**k), sc(2**k - 1), sc(L - 2**k), sc((2**k) + top), sc(((((top - 1) << (3 * w)) | (top << (2 * w)) | (ones << w)) << k) % L)]
        out += [sc(2**252 + top), sc(2**252 - top)]
    for k in (31, 32, 63, 64, 127, 128, 247, 251, 252):
        out += [sc(2**k), sc(2**k - 1), sc(L - 2**k)]
    out += [sc(L - 2), sc((L - 1) // 2), sc((L + 1) // 2)]
    return out


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
    ks = [sc(rnd.randrange(L)) for _ in range(n)] + [sc(0), sc(1), sc(L - 1), sc(8), sc(2**252)] + wide_window_edge_scalars()
    out, ok = e.mul_generator(np.frombuffer(b"".join(ks), np.uint8))
    assert ok.all()
    for i, k in enumerate(ks):
        assert bytes(out[i]) == O.point_mul_generator(k), i
    edge = wide_window_edge_scalars()[::5]
    a = [sc(rnd.randrange(L)) for _ in range(n)] + [sc(0), sc(L - 1), sc(1)] + edge[::-1]
    b = [sc(rnd.randrange(L)) for _ in range(n)] + [sc(5), sc(0), sc(L - 1)] + edge
    A = [O.point_mul_generator(sc(rnd.randrange(L))) for _ in range(n)] + [bytes(32), W.G_ENC, W.G_ENC] + [W.G_ENC] * len(edge)
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


def to_engine_range(e, spec):
    """oracle Range -> engine Range (same rings); also checks RangeDecomposition::optimal parity."""
    r = e.range_optimal(int(O.lib().eo_range_upper_bound(O.C.byref(spec))))
    assert r.rings == spec.rings and e.range_display(r) == O.range_display(spec)
    return r


def check_range_decomposition(e):
    for ub in (2, 5, 16, 17, 21, 42, 60, 100, 101, 256, 1000, 12345, 65536, 777777):
        to_engine_range(e, O.range_optimal(ub))


def tamper_range(cts, partials, rings, rnd, frac):
    n = cts.shape[0]
    idx = sorted(rnd.sample(range(n), max(1, int(n * frac)))) if n else []
    for k, i in enumerate(idx):
        kind = k % 5
        if kind == 0:
            cts[i, 32:] = np.frombuffer(O.point_add(bytes(cts[i, 32:]), W.G_ENC), np.uint8)
        elif kind == 1 and partials.shape[1] > 0:
            partials[i, 0, 32:] = np.frombuffer(O.point_add(bytes(partials[i, 0, 32:]), W.G_ENC), np.uint8)
        elif kind == 2:
            rings[i, -1] = np.frombuffer(W.BAD_SCALAR, np.uint8)
        elif kind == 3:
            cts[i, :32] = np.frombuffer(W.BAD_POINT2, np.uint8)
        else:
            a, b = rings[i, 1].copy(), rings[i, 2].copy()
            rings[i, 1], rings[i, 2] = b, a
    return idx


def check_verify_range(e, pk, upper_bound, n=40, label="ciphertext_range", frac=0.25, seed=3):
    spec = O.range_optimal(upper_bound)
    rng = to_engine_range(e, spec)
    rnd = random.Random(seed)
    values = [0, upper_bound - 1] + [rnd.randrange(upper_bound) for _ in range(n - 2)]
    cts, partials, rings = O.gen_range_batch(pk, spec, label, W.SEED_QV, np.array(values[:n], np.uint64))
    cts, partials, rings = cts.copy(), partials.copy(), rings.copy()
    tampered = tamper_range(cts, partials, rings, rnd, frac) if frac else []
    expected = O.verify_range_batch(pk, spec, label, cts, partials, rings)
    got = e.verify_range(rng, label, cts, partials, rings)
    assert got.tolist() == expected.tolist()
    if frac:
        assert (expected[tampered] != 0).all() and (np.delete(expected, tampered) == 0).all()
    # a proof is bound to its transcript label (range.rs:732-794)
    other = e.verify_range(rng, "another_label", cts[:4], partials[:4], rings[:4])
    assert (other != 0).all()


def as_engine_keyset(ks):
    from elastic_elgamal_b200 import _ffi
    out = _ffi.KeySet()
    O.C.memmove(O.C.byref(out), O.C.byref(ks), O.C.sizeof(_ffi.KeySet))
    return out


def as_engine_qv(p):
    from elastic_elgamal_b200 import _ffi
    out = _ffi.QvParams()
    out.options, out.credits = p.options, p.credits
    for dst, src in ((out.vote_range, p.vote_range), (out.credit_range, p.credit_range)):
        dst.n_rings = src.n_rings
        for i in range(src.n_rings):
            dst.size[i], dst.step[i] = src.size[i], src.step[i]
    return out


QV_VOTES = [[4, 0, 0, 1, 1], [1, 3, 0, 3, 1], [0, 0, 0, 0, 0], [2, 2, 2, 2, 2]]      # quadratic_voting.rs:188 and friends


def check_verify_qv(e, pk, sk=None, n=12, options=5, credits=20, seed=4):
    p = O.qv_params(options, credits)
    ep = e.qv_params(options, credits)
    assert ep.vote_range.rings == p.vote_range.rings and ep.credit_range.rings == p.credit_range.rings
    bsz = O.qv_ballot_size(p)
    assert e.qv_ballot_size(ep) == bsz
    rnd = random.Random(seed)
    if options == 5 and credits == 20:
        votes = [QV_VOTES[i % 4] for i in range(n)]
    else:
        mv = int(O.lib().eo_isqrt(credits))
        votes = []
        for _ in range(n):
            while True:
                v = [rnd.randrange(mv + 1) for _ in range(options)]
                if sum(x * x for x in v) <= credits:
                    break
            votes.append(v)
    ballots = O.gen_qv_batch(pk, p, W.SEED_QV, np.array(votes, np.uint64)).copy()
    vsz = 64 + 64 * (p.vote_range.n_rings - 1) + 32 * (1 + p.vote_range.rings_size)
    csz = 64 + 64 * (p.credit_range.n_rings - 1) + 32 * (1 + p.credit_range.rings_size)
    if n >= 8:
        # the reference's tamper patterns (quadratic_voting.rs:433-464) at fixed positions
        o = 2 % options
        ballots[1, vsz * o + 32:vsz * o + 64] = np.frombuffer(O.point_add(bytes(ballots[1, vsz * o + 32:vsz * o + 64]), W.G_ENC), np.uint8)
        ballots[2, vsz * options:vsz * options + 64] = ballots[3, vsz * options:vsz * options + 64]
        ballots[4, -32 * (2 * options + 2):] = ballots[5, -32 * (2 * options + 2):]
        ballots[6, -1] = 0xff
        ballots[7, 0:32] = np.frombuffer(W.BAD_POINT, np.uint8)
    expected, exp_tally = O.verify_qv_batch(pk, p, ballots)
    got, tally = e.verify_qv(ep, ballots)
    assert got.tolist() == expected.tolist(), (got, expected)
    assert (tally == exp_tally).all()
    if n >= 8:
        assert expected[1] == O.QV_VARIANT_BASE + o and expected[2] == O.QV_CREDIT_RANGE and expected[6] == O.MALFORMED
        assert expected[7] == O.MALFORMED and expected[0] == O.OK
    if sk is not None:
        table = O.DlogTable(0, 4 * n + 1)
        ok = [i for i in range(n) if expected[i] == 0]
        for k in range(options):
            assert table.get(O.decrypt_to_element(sk, bytes(tally[k]))) == sum(votes[i][k] for i in ok)


def check_shares_and_decrypt(e, n=10, shares=5, threshold=3, used=(0, 2, 4), table_hi=64, seed=6):
    rng = O.rng_from_seed(bytes([9] * 32))
    ks, secrets = O.dealer_new(shares, threshold, rng)
    eks = as_engine_keyset(ks)
    shared = bytes(ks.shared_key)
    rnd = random.Random(seed)
    values = [0, table_hi - 1, table_hi + 5] + [rnd.randrange(table_hi) for _ in range(n - 3)]
    cts = [O.encrypt(shared, v, rng) for v in values[:n]]
    sh, pr = [], []
    for ct in cts:
        row = [O.decrypt_share(ks, i, secrets[i], ct, rng) for i in used]
        sh.append([r[0] for r in row]); pr.append([r[1] for r in row])
    cts_a = np.frombuffer(b"".join(cts), np.uint8).reshape(n, 64).copy()
    sh_a = np.frombuffer(b"".join(b"".join(r) for r in sh), np.uint8).reshape(n, len(used), 32).copy()
    pr_a = np.frombuffer(b"".join(b"".join(r) for r in pr), np.uint8).reshape(n, len(used), 64).copy()
    # tamper: proof of another tally, share of another participant, malformed share, malformed scalar, malformed ciphertext
    tamper = n >= 8 and len(used) >= 3
    if tamper:
        pr_a[1, 0] = pr_a[2, 0]
        sh_a[3, 1] = sh_a[3, 2]
        sh_a[4, 2] = np.frombuffer(W.BAD_POINT, np.uint8)
        pr_a[5, 1, 32:] = np.frombuffer(W.BAD_SCALAR, np.uint8)
        cts_a[6, :32] = np.frombuffer(W.BAD_POINT2, np.uint8)
    expected = np.array([[O.verify_share(ks, used[j], bytes(cts_a[i]), bytes(sh_a[i, j]), bytes(pr_a[i, j])) for j in range(len(used))]
                         for i in range(n)], np.uint8)
    got = e.verify_shares(eks, list(used), cts_a, sh_a, pr_a)
    assert got.tolist() == expected.tolist(), (got, expected)
    if tamper:
        assert expected[1, 0] == O.CHALLENGE_MISMATCH and expected[4, 2] == O.MALFORMED and (expected[6] == O.MALFORMED).all()
        assert (expected[0] == 0).all()
    # combine + decrypt + table lookup (sharing/mod.rs:302-325, decryption.rs:138-144, encryption.rs:287-297)
    table = e.dlog_table(0, table_hi)
    otable = O.DlogTable(0, table_hi)
    vals, found = e.combine_decrypt(list(used[:threshold]), cts_a, sh_a, table)
    for i in range(n):
        rc, elem = O.combine_decrypt(list(used[:threshold]), [bytes(sh_a[i, j]) for j in range(threshold)], bytes(cts_a[i]))
        if rc != 0:
            assert found[i] == 2
            continue
        ov = otable.get(elem)
        if ov is None:
            assert found[i] == 0
        else:
            assert found[i] == 1 and int(vals[i]) == ov
    assert found[0] == 1 and vals[0] == 0 and found[1] == 1 and vals[1] == table_hi - 1 and found[2] == 0
    table.close()


def item_blocks(seed, index, count):
    """`count` 64-byte keystream blocks of item `index`'s ChaCha20 stream (block counter = index << 20)."""
    rng = O.rng_from_seed(seed, first_block=index << 20)
    return b"".join(O.rng_block(rng) for _ in range(count))


def check_encrypt_bool(e, pk, n=16, seed=W.SEED_CHOICE):
    """eg_encrypt_bool_batch fed with the oracle's per-item ChaCha20 blocks reproduces encrypt_bool byte for byte."""
    ocs, ops = O.gen_bool_batch(pk, seed, n)
    values = np.array([i & 1 for i in range(n)], np.uint8)
    wide = np.frombuffer(b"".join(item_blocks(seed, i, 3) for i in range(n)), np.uint8)
    cts, proofs = e.encrypt_bool(values, wide)
    assert (cts == ocs).all() and (proofs == ops).all()
    assert (e.verify_bool(cts, proofs) == 0).all()


def check_encrypt_choice(e, pk, options=5, n=12, seed=W.SEED_CHOICE):
    ocs, ors, oss = O.gen_choice_batch(pk, options, seed, n)
    values = np.zeros((n, options), np.uint8)
    for i in range(n):
        values[i, i % options] = 1
    wide = np.frombuffer(b"".join(item_blocks(seed, i, 3 * options + 1) for i in range(n)), np.uint8)
    cts, rings, sums = e.encrypt_choice(options, values, wide, single=True)
    assert (cts == ocs).all() and (rings == ors).all() and (sums == oss).all()
    v, _ = e.verify_choice(options, cts, rings, sums)
    assert (v == 0).all()


def check_identity_commitments(e, pk, options=5, n=20):
    """Valid proofs whose commitments are the identity element (all-zero encodings): the prover's nonces are caller
    supplied, so all-zero 64-byte blocks give x = 0, i.e. [x]G = [x]K = O, in the real equation of a ring (first
    equation for value 0: encoded inside k_ring; last equation for value 1: deferred to k_terminal) and in the sum
    proof (k_commit -> k_terminal).  The oracle must accept what the engine produced, and the engine its own output."""
    draws = 3 * options + 1
    rnd = random.Random(5)
    wide = np.frombuffer(rnd.randbytes(n * draws * 64), np.uint8).reshape(n, draws, 64).copy()
    values = np.zeros((n, options), np.uint8)
    for i in range(n):
        c = i % options
        values[i, c] = 1
        pos = 0
        for k in range(options):                     # ring k draws r_k, x_k and, for value 0, the forged response
            if (i // options + k) % 2 == 0:
                wide[i, pos + 1] = 0                 # x_k = 0
            pos += 2 + (0 if k == c else 1)
        if i % 3 != 1:
            wide[i, draws - 1] = 0                   # nonce of the sum proof
    cts, rings, sums = e.encrypt_choice(options, values, wide, single=True)
    expected, exp_tally = O.verify_choice_batch(pk, options, True, cts, rings, sums)
    assert (expected == 0).all()
    got, tally = e.verify_choice(options, cts, rings, sums, single=True, tally=True)
    assert (got == 0).all() and (tally == exp_tally).all()
    # and a flipped response still fails the same way on both sides
    rings = rings.copy()
    rings[0, 1, 0] ^= 1
    expected, _ = O.verify_choice_batch(pk, options, True, cts, rings, sums)
    got, _ = e.verify_choice(options, cts, rings, sums, single=True, tally=False)
    assert got.tolist() == expected.tolist() and got[0] != 0


def check_verifiers_chunked(e, pk, chunk=5, n=23):
    """verify_bool / verify_range run the same double-buffered chunk pipeline: several full chunks and a remainder must
    give the oracle's verdicts (range chunks are clamped to >= 1024 proofs unless n is larger, so the range batch is
    sized by the caller)."""
    e.set_chunk_items(chunk)
    try:
        check_verify_bool(e, pk, n=max(24, n))
        check_shares_and_decrypt(e, n=max(10, n // 2))
    finally:
        e.set_chunk_items(0)


def check_provers_chunked(e, pk, chunk=3, n=11):
    """The provers run a double-buffered chunk pipeline (copies of chunks c + 1 / c - 1 overlap the kernels of chunk c):
    with a chunk size that splits the batch into several full chunks and a remainder, every output byte must still equal
    the oracle's."""
    e.set_chunk_items(chunk)
    try:
        check_encrypt_bool(e, pk, n=n)
        check_encrypt_choice(e, pk, options=3, n=n)
        check_encrypt_range(e, pk, 21, n=n)
        check_encrypt_qv(e, pk, n=max(4, n // 2))
    finally:
        e.set_chunk_items(0)


def check_encrypt_multi_choice(e, pk, options=4, n=10, seed=b"\x09" * 32):
    """EncryptedChoice::new with MultiChoice (choice.rs:313-349): any 0/1 pattern, no sum proof."""
    rnd = random.Random(17)
    values = np.array([[rnd.randrange(2) for _ in range(options)] for _ in range(n)], np.uint8)
    values[0, :] = 0
    values[1, :] = 1
    exp_c, exp_r = [], []
    for i in range(n):
        rng = O.rng_from_seed(seed, first_block=i << 20)
        c, r, _ = O.choice_new(pk, [bool(x) for x in values[i]], False, rng)
        exp_c.append(c); exp_r.append(r)
    wide = np.frombuffer(b"".join(item_blocks(seed, i, 3 * options) for i in range(n)), np.uint8)
    cts, rings, sums = e.encrypt_choice(options, values, wide, single=False)
    assert sums is None
    assert cts.tobytes() == b"".join(exp_c) and rings.tobytes() == b"".join(exp_r)
    v, _ = e.verify_choice(options, cts, rings, None, single=False)
    assert (v == 0).all()


def check_encrypt_against_reference_snapshots(e):
    """The engine's prover, fed with the ChaCha20 blocks of `ChaChaRng::seed_from_u64(12345)`, regenerates the
    reference's own golden snapshots (tests/snapshots.rs:73-130: bool-encryption, encrypted-choice,
    encrypted-multi-choice) byte for byte, and the engine's verifier accepts them."""
    import json
    import pathlib
    gold = json.loads((pathlib.Path(__file__).parent / "golden" / "ristretto_snapshots.json").read_text())
    H = bytes.fromhex

    def fresh(blocks):
        rng = O.rng_from_u64(12345)
        sk, pk = O.keypair(rng)                      # snapshots.rs:32-33: the first draw is the secret key
        return pk, np.frombuffer(b"".join(O.rng_block(rng) for _ in range(blocks)), np.uint8)

    def ctb(d):
        return H(d["random_element"]) + H(d["blinded_element"])

    try:
        pk, wide = fresh(3)
        e.set_receiver(pk)
        cts, proofs = e.encrypt_bool(np.array([1], np.uint8), wide)
        g = gold["bool-encryption"]
        assert cts.tobytes() == ctb(g["ciphertext"]) and proofs.tobytes() == H(gold["bool-encryption-bin"])
        assert e.verify_bool(cts, proofs).tolist() == [0]

        pk, wide = fresh(16)
        cts, rings, sums = e.encrypt_choice(5, np.array([[0, 0, 0, 1, 0]], np.uint8), wide, single=True)
        g = gold["encrypted-choice"]
        assert cts.tobytes() == b"".join(ctb(c) for c in g["choices"])
        assert rings.tobytes() == H(g["range_proof"]["common_challenge"]) + b"".join(H(x) for x in g["range_proof"]["ring_responses"])
        assert sums.tobytes() == H(g["sum_proof"]["challenge"]) + H(g["sum_proof"]["response"])
        assert e.verify_choice(5, cts, rings, sums)[0].tolist() == [0]

        pk, wide = fresh(15)
        cts, rings, sums = e.encrypt_choice(5, np.array([[0, 1, 1, 0, 1]], np.uint8), wide, single=False)
        g = gold["encrypted-multi-choice"]
        assert cts.tobytes() == b"".join(ctb(c) for c in g["choices"])
        assert rings.tobytes() == H(g["range_proof"]["common_challenge"]) + b"".join(H(x) for x in g["range_proof"]["ring_responses"])
        assert e.verify_choice(5, cts, rings, None, single=False)[0].tolist() == [0]
    finally:
        e.set_receiver(W.receiver()[1])


# ---------------------------------------------------------------- CommitmentEquivalenceProof / ProofOfPossession

# Blinding base used in Bulletproofs (tests/snapshots.rs:253-257)
BLINDING_BASE = bytes([140, 146, 64, 180, 86, 169, 230, 220, 101, 195, 119, 161, 4, 141, 116, 95, 148, 160,
                       140, 219, 127, 68, 203, 205, 123, 70, 243, 64, 72, 135, 17, 52])


def check_commitment_equiv(e, pk, n=24, label="test", seed=W.SEED_CHOICE):
    from elastic_elgamal_b200 import EngineError, _ffi
    e.set_blinding_base(BLINDING_BASE)
    values = (np.arange(n, dtype=np.uint64) * 977) % 100000
    cts, coms, proofs = O.gen_ceq_batch(pk, BLINDING_BASE, label, seed, values)
    cts, coms, proofs = cts.copy(), coms.copy(), proofs.copy()
    if n >= 10:
        # commitment.rs:300-333 negative cases + malformed inputs
        cts[1] = cts[2]                                                                   # proof for another ciphertext
        coms[3] = np.frombuffer(O.point_add(bytes(coms[3]), W.G_ENC), np.uint8)           # commitment + G
        proofs[4, 32:64], proofs[4, 64:96] = proofs[4, 64:96].copy(), proofs[4, 32:64].copy()
        proofs[5, 96:] = np.frombuffer(W.BAD_SCALAR, np.uint8)
        coms[6] = np.frombuffer(W.BAD_POINT2, np.uint8)
        cts[7, 32:] = np.frombuffer(W.BAD_POINT, np.uint8)
        proofs[8, 0] ^= 1
    expected = O.verify_ceq_batch(pk, BLINDING_BASE, label, cts, coms, proofs)
    got = e.verify_commitment_equiv(label, cts, coms, proofs)
    assert got.tolist() == expected.tolist(), (got, expected)
    if n >= 10:
        assert expected[0] == O.OK and expected[1] == O.CHALLENGE_MISMATCH and expected[3] == O.CHALLENGE_MISMATCH
        assert expected[5] == O.MALFORMED and expected[6] == O.MALFORMED and expected[7] == O.MALFORMED
    # another transcript label rejects everything that was accepted (commitment.rs:322-332)
    other = e.verify_commitment_equiv("other_" + label, cts, coms, proofs)
    assert other.tolist() == O.verify_ceq_batch(pk, BLINDING_BASE, "other_" + label, cts, coms, proofs).tolist()
    assert not (other == O.OK).any()
    assert e.verify_commitment_equiv(label, cts[:0], coms[:0], proofs[:0]).shape == (0,)
    for bad, status in ((bytes(32), _ffi.ERR_IDENTITY_KEY), (W.BAD_POINT, _ffi.ERR_INVALID_ELEMENT)):
        try:
            e.set_blinding_base(bad)
            raise AssertionError("invalid blinding base accepted")
        except EngineError as exc:
            assert exc.status == status
    try:
        e.verify_commitment_equiv(label, cts, coms, proofs)
        raise AssertionError("verification without a blinding base")
    except EngineError as exc:
        assert exc.status == _ffi.ERR_NO_RECEIVER
    e.set_blinding_base(BLINDING_BASE)


def check_commitment_equiv_snapshot(e):
    """tests/snapshots.rs:163-189: the reference's own ciphertext / commitment / proof for seed 12345 verifies."""
    import json
    import pathlib
    gold = json.loads((pathlib.Path(__file__).parent / "golden" / "ristretto_snapshots.json").read_text())["commitment-equiv-proof"]
    rng = O.rng_from_u64(12345)
    sk, pk = O.keypair(rng)
    e.set_receiver(pk)
    e.set_blinding_base(BLINDING_BASE)
    ct = bytes.fromhex(gold["ciphertext"]["random_element"]) + bytes.fromhex(gold["ciphertext"]["blinded_element"])
    pr = gold["proof"]
    proof = b"".join(bytes.fromhex(pr[k]) for k in ("challenge", "randomness_response", "value_response", "commitment_response"))
    com = bytes.fromhex(gold["commitment"])
    as_np = lambda b: np.frombuffer(b, np.uint8)
    assert e.verify_commitment_equiv("test", as_np(ct), as_np(com), as_np(proof)).tolist() == [O.OK]
    assert e.verify_commitment_equiv("tesT", as_np(ct), as_np(com), as_np(proof)).tolist() == [O.CHALLENGE_MISMATCH]


def check_possession(e, n=12, keys_per_proof=5, label="test_multi_PoP", seed=b"\x0b" * 32):
    k = keys_per_proof
    keys, proofs = O.gen_pop_batch(k, label, seed, n)
    keys, proofs = keys.copy(), proofs.copy()
    if n >= 8:
        keys[1] = keys[1, ::-1].copy() if k > 1 else keys[2]            # possession.rs:55-66: keys in another order
        proofs[2] = proofs[3]                                           # proof of other keys
        proofs[4, 1] = np.frombuffer(W.BAD_SCALAR2, np.uint8)
        keys[5, k - 1] = np.frombuffer(W.BAD_POINT, np.uint8)
        keys[6, 0] = 0                                                  # identity key
        proofs[7, 0, 0] ^= 1
    expected = O.verify_pop_batch(label, keys, proofs)
    got = e.verify_possession(label, keys, proofs)
    assert got.tolist() == expected.tolist(), (got, expected)
    if n >= 8:
        assert expected[0] == O.OK and expected[2] == O.CHALLENGE_MISMATCH and expected[4] == O.MALFORMED
        assert expected[5] == O.MALFORMED and expected[6] == O.MALFORMED and expected[7] in (O.CHALLENGE_MISMATCH, O.MALFORMED)
        if k > 1:
            assert expected[1] == O.CHALLENGE_MISMATCH
    assert not (e.verify_possession(label + "x", keys, proofs) == O.OK).any()
    assert e.verify_possession(label, keys[:0], proofs[:0]).shape == (0,)


# ---------------------------------------------------------------- wire format

def check_base64url(e, n=200):
    """Against Python's base64 module (RFC 4648 section 5, no padding = base64ct::Base64UrlUnpadded) and the reference's
    own serde vectors (serde.rs:402-429; tests/snapshots/*.snap strings)."""
    import base64
    import json
    import pathlib
    rs = np.random.RandomState(11)
    for size in (1, 2, 3, 31, 32, 33, 64, 96, 128, 352):
        raw = rs.randint(0, 256, (n, size)).astype(np.uint8)
        raw[0] = 0
        raw[1] = 0xff
        text = e.base64url_encode(raw)
        expect = [base64.urlsafe_b64encode(bytes(r)).rstrip(b"=") for r in raw]
        assert [bytes(t) for t in text] == expect
        back, ok = e.base64url_decode(text, size)
        assert ok.all() and (back == raw).all()
        # strictness: padding, the standard alphabet's '+' and '/', whitespace, non-canonical trailing bits
        bad = text.copy()
        bad[2, 0] = ord("=")
        bad[3, -1] = ord("+")
        bad[4, 0] = ord("/")
        bad[5, 0] = ord(" ")
        bad[6, -1] = 0x80
        _, ok = e.base64url_decode(bad, size)
        assert ok.tolist()[:8] == [True, True, False, False, False, False, False, True]
        if size % 3:
            nc = text.copy()
            v = nc[7, -1]
            alphabet = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_"
            nc[7, -1] = alphabet[alphabet.index(bytes([v])) | 1]       # set an unused trailing bit
            _, ok = e.base64url_decode(nc, size)
            assert not ok[7] and ok[8:].all()
    # serde.rs:402-404 / :427-429: these strings decode (the *values* are then rejected by the element / scalar checks)
    for s_, valid_el, valid_sc in (("tNDkeYUVQWgh34d-RqaElOk7yFB8d2qCh5f4Vi2euT0", False, None),
                                   ("nN3xf7lSOX0_zs6QPBwWHYi0Dkx2Ln_z1MPwnbzaM_8", None, False)):
        raw, ok = e.base64url_decode(np.frombuffer(s_.encode(), np.uint8), 32)
        assert ok.all() and bytes(raw[0]) == base64.urlsafe_b64decode(s_ + "=")
        if valid_el is not None:
            assert e.elements_validate(raw).tolist() == [valid_el]
        if valid_sc is not None:
            assert e.scalars_validate(raw).tolist() == [valid_sc]
    assert e.base64url_decode(np.zeros((0, 43), np.uint8), 32)[0].shape == (0, 32)


# ---------------------------------------------------------------- RangeProof::new on the GPU

def check_encrypt_range(e, pk, upper_bound, n=12, label="ciphertext_range", seed=W.SEED_CHOICE):
    """Same ChaCha blocks as the oracle's prover (item i: stream i << 20) => byte-identical ciphertexts, partial
    ciphertexts and ring proofs; the GPU verifier accepts them."""
    spec = O.range_optimal(upper_bound)
    espec = to_engine_range(e, spec)
    rnd = random.Random(upper_bound)
    values = [0, upper_bound - 1] + [rnd.randrange(upper_bound) for _ in range(max(0, n - 2))]
    values = np.array(values[:n], np.uint64)
    cts, partials, rings = O.gen_range_batch(pk, spec, label, seed, values)
    draws = e.lib.eg_range_prover_draws(O.C.byref(espec))
    assert draws == spec.n_rings + spec.rings_size
    wide = np.frombuffer(b"".join(item_blocks(seed, i, draws) for i in range(n)), np.uint8).reshape(n, draws, 64)
    g_cts, g_partials, g_rings = e.encrypt_range(espec, label, values, wide)
    assert (g_cts == cts).all()
    assert (g_partials == partials).all()
    assert (g_rings == rings).all()
    assert (e.verify_range(espec, label, g_cts, g_partials, g_rings) == 0).all()
    from elastic_elgamal_b200 import EngineError, _ffi
    try:
        e.encrypt_range(espec, label, np.array([upper_bound], np.uint64), wide[:1])
        raise AssertionError("out-of-range value accepted")
    except EngineError as exc:
        assert exc.status == _ffi.ERR_INVALID_ARG


def check_encrypt_range_reference_snapshot(e):
    """tests/snapshots.rs:96-109: encrypt_range(optimal(100), 42) from seed 12345, byte for byte."""
    import json
    import pathlib
    gold = json.loads((pathlib.Path(__file__).parent / "golden" / "ristretto_snapshots.json").read_text())["range-encryption"]
    rng = O.rng_from_u64(12345)
    sk, pk = O.keypair(rng)
    e.set_receiver(pk)
    espec = e.range_optimal(100)
    draws = e.lib.eg_range_prover_draws(O.C.byref(espec))
    wide = np.frombuffer(b"".join(O.rng_block(rng) for _ in range(draws)), np.uint8).reshape(1, draws, 64)
    cts, partials, rings = e.encrypt_range(espec, "ciphertext_range", np.array([42], np.uint64), wide)
    hx = bytes.fromhex
    assert bytes(cts[0]) == hx(gold["ciphertext"]["random_element"]) + hx(gold["ciphertext"]["blinded_element"])
    pr = gold["proof"]
    assert bytes(partials[0].reshape(-1)) == b"".join(hx(c["random_element"]) + hx(c["blinded_element"]) for c in pr["partial_ciphertexts"])
    assert bytes(rings[0].reshape(-1)) == hx(pr["common_challenge"]) + b"".join(hx(x) for x in pr["ring_responses"])


# ---------------------------------------------------------------- QuadraticVotingBallot::new on the GPU

def check_encrypt_qv(e, pk, sk=None, n=8, options=5, credits=20, seed=W.SEED_QV):
    p, ep = O.qv_params(options, credits), e.qv_params(options, credits)
    rnd = random.Random(n * options + credits)
    if options == 5 and credits == 20:
        votes = [QV_VOTES[i % 4] for i in range(n)]
    else:
        mv = int(O.lib().eo_isqrt(credits))
        votes = []
        for _ in range(n):
            while True:
                v = [rnd.randrange(mv + 1) for _ in range(options)]
                if sum(x * x for x in v) <= credits:
                    break
            votes.append(v)
    votes = np.array(votes, np.uint64)
    ballots = O.gen_qv_batch(pk, p, seed, votes)
    draws = e.lib.eg_qv_prover_draws(O.C.byref(ep))
    wide = np.frombuffer(b"".join(item_blocks(seed, i, draws) for i in range(n)), np.uint8).reshape(n, draws, 64)
    got = e.encrypt_qv(ep, votes, wide)
    assert got.shape == ballots.shape
    bad = np.nonzero((got != ballots).any(axis=1))[0]
    assert bad.size == 0, (bad[:4], np.nonzero(got[bad[0]] != ballots[bad[0]])[0][:8])
    v, t = e.verify_qv(ep, got)
    assert (v == 0).all()
    if sk is not None:
        table = O.DlogTable(0, 8 * n + 1)
        assert [table.get(O.decrypt_to_element(sk, bytes(t[k]))) for k in range(options)] == votes.sum(axis=0).tolist()
    from elastic_elgamal_b200 import EngineError, _ffi
    too_many = votes[:1].copy()
    too_many[0, 0] = 1000
    try:
        e.encrypt_qv(ep, too_many, wide[:1])
        raise AssertionError("out-of-range vote accepted")
    except EngineError as exc:
        assert exc.status == _ffi.ERR_INVALID_ARG


def check_encrypt_qv_reference_snapshot(e):
    """tests/snapshots.rs:153-161: QuadraticVotingBallot::new(params(pk, 5, 15), [3, 0, 1, 0, 2]) from seed 12345."""
    import json
    import pathlib
    gold = json.loads((pathlib.Path(__file__).parent / "golden" / "ristretto_snapshots.json").read_text())["qv-ballot"]
    rng = O.rng_from_u64(12345)
    sk, pk = O.keypair(rng)
    e.set_receiver(pk)
    ep = e.qv_params(5, 15)
    draws = e.lib.eg_qv_prover_draws(O.C.byref(ep))
    wide = np.frombuffer(b"".join(O.rng_block(rng) for _ in range(draws)), np.uint8).reshape(1, draws, 64)
    ballot = bytes(e.encrypt_qv(ep, np.array([[3, 0, 1, 0, 2]], np.uint64), wide)[0])
    hx = bytes.fromhex

    def rp(d):     # CiphertextWithRangeProof -> ct | partial ciphertexts | common challenge | responses
        pr = d["range_proof"]
        return (hx(d["ciphertext"]["random_element"]) + hx(d["ciphertext"]["blinded_element"])
                + b"".join(hx(c["random_element"]) + hx(c["blinded_element"]) for c in pr["partial_ciphertexts"])
                + hx(pr["common_challenge"]) + b"".join(hx(x) for x in pr["ring_responses"]))
    ce = gold["credit_equivalence_proof"]
    expected = (b"".join(rp(v) for v in gold["votes"]) + rp(gold["credit"]) + hx(ce["challenge"])
                + b"".join(hx(x) for x in ce["ciphertext_responses"]) + hx(ce["sum_response"]))
    assert ballot == expected


# ---------------------------------------------------------------- encrypt / encrypt_zero / vartime_multi_mul

def check_encrypt_plain_and_zero(e, pk, sk, n=20):
    """Same ChaCha blocks as the oracle (one sequential stream) => identical ciphertexts and zero proofs; includes the
    reference's `ciphertext` and `zero-encryption` snapshots (tests/snapshots.rs:31-71)."""
    rng = O.rng_from_seed(bytes([6] * 32))
    rng2 = O.rng_from_seed(bytes([6] * 32))
    values = [0, 1, 2**32 + 5, 2**63] + list(range(100, 100 + n - 4))
    expected = [O.encrypt(pk, v, rng) for v in values]
    wide = np.frombuffer(b"".join(O.rng_block(rng2) for _ in range(n)), np.uint8).reshape(n, 64)
    got = e.encrypt(np.array(values, np.uint64), wide)
    assert [bytes(g) for g in got] == expected
    table = O.DlogTable(0, 200)
    assert table.get(O.decrypt_to_element(sk, bytes(got[5]))) == 101
    expected = [O.encrypt_zero(pk, rng) for _ in range(n)]
    wide = np.frombuffer(b"".join(O.rng_block(rng2) for _ in range(2 * n)), np.uint8).reshape(n, 2, 64)
    cts, proofs = e.encrypt_zero(wide)
    assert [(bytes(c), bytes(p)) for c, p in zip(cts, proofs)] == expected
    assert (e.verify_zero(cts, proofs) == 0).all()
    assert e.encrypt(np.zeros(0, np.uint64), np.zeros((0, 64), np.uint8)).shape == (0, 64)


def check_encrypt_plain_and_zero_snapshots(e):
    import json
    import pathlib
    gold = json.loads((pathlib.Path(__file__).parent / "golden" / "ristretto_snapshots.json").read_text())
    hx = bytes.fromhex
    for name in ("ciphertext", "zero-encryption"):
        rng = O.rng_from_u64(12345)
        sk, pk = O.keypair(rng)
        e.set_receiver(pk)
        if name == "ciphertext":
            wide = np.frombuffer(O.rng_block(rng), np.uint8).reshape(1, 64)
            ct = bytes(e.encrypt(np.array([42], np.uint64), wide)[0])
            assert ct == hx(gold[name]["random_element"]) + hx(gold[name]["blinded_element"]) == hx(gold["ciphertext-bin"])
        else:
            wide = np.frombuffer(O.rng_block(rng) + O.rng_block(rng), np.uint8).reshape(1, 2, 64)
            cts, proofs = e.encrypt_zero(wide)
            g = gold[name]
            assert bytes(cts[0]) == hx(g["ciphertext"]["random_element"]) + hx(g["ciphertext"]["blinded_element"])
            assert bytes(proofs[0]) == hx(g["proof"]["challenge"]) + hx(g["proof"]["response"]) == hx(gold["zero-encryption-bin"])


def check_multi_mul(e, n=12):
    rnd = random.Random(17)
    rng = O.rng_from_seed(bytes([8] * 32))
    for terms in (1, 2, 3, 7, 16):
        scalars = np.zeros((n, terms, 32), np.uint8)
        points = np.zeros((n, terms, 32), np.uint8)
        expected = []
        for i in range(n):
            acc = bytes(32)
            for j in range(terms):
                s_ = O.scalar_reduce_wide(O.rng_block(rng)) if (i + j) % 5 else sc(rnd.choice([0, 1, W.L - 1]))
                p_ = O.point_mul_generator(O.scalar_reduce_wide(O.rng_block(rng))) if (i * j) % 7 != 3 else bytes(32)
                scalars[i, j] = np.frombuffer(s_, np.uint8)
                points[i, j] = np.frombuffer(p_, np.uint8)
                acc = O.point_add(acc, O.point_mul(s_, p_))
            expected.append(acc)
        if n >= 4:
            points[1, terms - 1] = np.frombuffer(W.BAD_POINT, np.uint8)
            scalars[2, 0] = np.frombuffer(W.BAD_SCALAR, np.uint8)
        out, ok = e.multi_mul(scalars, points)
        for i in range(n):
            if n >= 4 and i in (1, 2):
                assert not ok[i] and bytes(out[i]) == bytes(32)
            else:
                assert ok[i] and bytes(out[i]) == expected[i], (terms, i)


def check_ciphertext_ops(e, pk, n=12):
    """Ciphertext Add / Sub / Neg / Mul<&Scalar> (encryption.rs:160-226) through eg_ciphertexts_lincomb_batch against the
    oracle's point arithmetic, incl. a - a = the identity ciphertext, an undecodable element and a non-canonical scalar."""
    rng = O.rng_from_seed(bytes([11] * 32))
    one, minus_one = sc(1), sc(W.L - 1)
    cts = [O.encrypt(pk, i, rng) for i in range(2 * n)]
    ks = [O.scalar_reduce_wide(O.rng_block(rng)) for _ in range(n)]

    def lin(terms):                                     # [(scalar, ct)] -> 64 bytes
        r, b = bytes(32), bytes(32)
        for s_, ct in terms:
            r = O.point_add(r, O.point_mul(s_, ct[:32]))
            b = O.point_add(b, O.point_mul(s_, ct[32:]))
        return r + b

    def run(rows):
        t = len(rows[0])
        S = np.frombuffer(b"".join(s_ for row in rows for s_, _ in row), np.uint8).reshape(len(rows), t, 32)
        C_ = np.frombuffer(b"".join(ct for row in rows for _, ct in row), np.uint8).reshape(len(rows), t, 64)
        out, ok = e.ciphertexts_lincomb(S, C_)
        return [bytes(o) for o in out], ok

    add = [[(one, cts[2 * i]), (one, cts[2 * i + 1])] for i in range(n)]
    sub = [[(one, cts[2 * i]), (minus_one, cts[2 * i + 1])] for i in range(n)]
    sub[0] = [(one, cts[0]), (minus_one, cts[0])]       # a - a
    neg = [[(minus_one, cts[i])] for i in range(n)]
    mul = [[(ks[i], cts[i])] for i in range(n)]
    mul[0] = [(sc(0), cts[0])]
    for rows in (add, sub, neg, mul):
        out, ok = run(rows)
        assert ok.all() and out == [lin(row) for row in rows]
    assert run(sub)[0][0] == bytes(64) and run(mul)[0][0] == bytes(64)
    # a + b decrypts to the sum of the plaintexts: homomorphism (encryption.rs:96-112), checked on the encodings
    bad = [row[:] for row in add]
    bad[1][0] = (one, W.BAD_POINT + cts[2][32:])
    bad[2][1] = (W.BAD_SCALAR, cts[5])
    out, ok = run(bad)
    assert not ok[1] and not ok[2] and out[1] == bytes(64) and out[2] == bytes(64) and ok[0] and all(ok[3:])


# ---------------------------------------------------------------- randomized differential tests

def _flip_random(arrs, rnd, items, flips_per_item=1):
    """Flips random bits at random byte positions of random items across the given (n, ...) uint8 arrays."""
    flat = [a.reshape(a.shape[0], -1) for a in arrs]
    sizes = [f.shape[1] for f in flat]
    total = sum(sizes)
    for i in items:
        for _ in range(flips_per_item):
            pos = rnd.randrange(total)
            for f, sz in zip(flat, sizes):
                if pos < sz:
                    f[i, pos] ^= 1 << rnd.randrange(8)
                    break
                pos -= sz


def check_fuzz_differential(e, pk, n=64, seed=1):
    """Random single- and multi-bit corruption anywhere in the inputs (points, scalars, proofs): the GPU verdict must
    equal the oracle's for every item -- malformed encodings, non-canonical scalars and plain mismatches alike."""
    rnd = random.Random(seed)
    half = list(range(0, n, 2))
    # EncryptedChoice
    cts, rings, sums = O.gen_choice_batch(pk, 3, W.SEED_CHOICE, n)
    cts, rings, sums = cts.copy(), rings.copy(), sums.copy()
    _flip_random([cts, rings, sums], rnd, half, flips_per_item=rnd.choice([1, 2, 5]))
    ov, ot = O.verify_choice_batch(pk, 3, True, cts, rings, sums)
    gv, gt = e.verify_choice(3, cts, rings, sums)
    assert gv.tolist() == ov.tolist() and (gt == ot).all()
    assert (ov[1::2] == 0).all() and (ov[half] != 0).sum() >= len(half) - 2
    # bool
    bc, bp = O.gen_bool_batch(pk, W.SEED_CHOICE, n)
    bc, bp = bc.copy(), bp.copy()
    _flip_random([bc, bp], rnd, half)
    assert e.verify_bool(bc, bp).tolist() == O.verify_bool_batch(pk, bc, bp).tolist()
    # RangeProof
    spec = O.range_optimal(21)
    espec = to_engine_range(e, spec)
    values = np.array([rnd.randrange(21) for _ in range(n)], np.uint64)
    rc, rp, rr = O.gen_range_batch(pk, spec, "ciphertext_range", W.SEED_CHOICE, values)
    rc, rp, rr = rc.copy(), rp.copy(), rr.copy()
    _flip_random([rc, rp, rr], rnd, half)
    assert e.verify_range(espec, "ciphertext_range", rc, rp, rr).tolist() == O.verify_range_batch(pk, spec, "ciphertext_range", rc, rp, rr).tolist()
    # QuadraticVotingBallot
    p, ep = O.qv_params(3, 15), e.qv_params(3, 15)
    votes = np.array([[rnd.randrange(3), rnd.randrange(3), rnd.randrange(2)] for _ in range(n // 2)], np.uint64)
    ballots = O.gen_qv_batch(pk, p, W.SEED_QV, votes).copy()
    _flip_random([ballots], rnd, list(range(0, n // 2, 2)))
    ov, ot = O.verify_qv_batch(pk, p, ballots)
    gv, gt = e.verify_qv(ep, ballots)
    assert gv.tolist() == ov.tolist() and (gt == ot).all()
    # CommitmentEquivalenceProof
    e.set_blinding_base(BLINDING_BASE)
    cc, cm, cp = O.gen_ceq_batch(pk, BLINDING_BASE, "fuzz", W.SEED_CHOICE, np.arange(n, dtype=np.uint64))
    cc, cm, cp = cc.copy(), cm.copy(), cp.copy()
    _flip_random([cc, cm, cp], rnd, half)
    assert e.verify_commitment_equiv("fuzz", cc, cm, cp).tolist() == O.verify_ceq_batch(pk, BLINDING_BASE, "fuzz", cc, cm, cp).tolist()
    # ProofOfPossession
    keys, proofs = O.gen_pop_batch(3, "fuzz_pop", bytes([12] * 32), n)
    keys, proofs = keys.copy(), proofs.copy()
    _flip_random([keys, proofs], rnd, half)
    assert e.verify_possession("fuzz_pop", keys, proofs).tolist() == O.verify_pop_batch("fuzz_pop", keys, proofs).tolist()


# ---------------------------------------------------------------- PublicKeySet::from_participants

def check_keysets_validate(e, n_sets=8, shares=5, threshold=3, seed=13):
    """key_set.rs:238-270: consistent sets restore the dealer's shared key; swapped / shifted keys are
    MalformedParticipantKeys; undecodable or identity keys are malformed."""
    rng = O.rng_from_seed(bytes([seed] * 32))
    keys = np.zeros((n_sets, shares, 32), np.uint8)
    shared_expected = []
    for i in range(n_sets):
        ks, _ = O.dealer_new(shares, threshold, rng)
        for j in range(shares):
            keys[i, j] = np.frombuffer(bytes(ks.participant_keys[j]), np.uint8)
        shared_expected.append(bytes(ks.shared_key))
    if n_sets >= 6 and shares > threshold and threshold > 1:
        keys[1, [0, shares - 1]] = keys[1, [shares - 1, 0]]                                  # order of keys matters
        keys[2, 1] = np.frombuffer(O.point_add(bytes(keys[2, 1]), W.G_ENC), np.uint8)        # one of the first t keys + G
        keys[3, shares - 1] = np.frombuffer(O.point_add(bytes(keys[3, shares - 1]), W.G_ENC), np.uint8)
        keys[4, 2] = np.frombuffer(W.BAD_POINT2, np.uint8)
        keys[5, shares - 1] = 0                                                              # identity key
    expected = [O.keyset_from_participants(shares, threshold, [bytes(k) for k in keys[i]]) for i in range(n_sets)]
    shared, v = e.keysets_validate(shares, threshold, keys)
    assert v.tolist() == [x[0] for x in expected], (v, expected)
    for i in range(n_sets):
        assert bytes(shared[i]) == (expected[i][1] if expected[i][0] == 0 else bytes(32))
    assert expected[0] == (0, shared_expected[0])
    if n_sets >= 6 and shares > threshold and threshold > 1:
        assert [x[0] for x in expected[1:6]] == [O.MALFORMED_PARTICIPANT_KEYS] * 3 + [O.MALFORMED] * 2
    assert e.keysets_validate(shares, threshold, keys[:0])[1].shape == (0,)


# ---------------------------------------------------------------- seeded provers (in-kernel ChaCha20) / constant-time mode

def _gold():
    import json
    import pathlib
    return json.loads((pathlib.Path(__file__).parent / "golden" / "ristretto_snapshots.json").read_text())


def check_seeded_provers(e, pk, sk, n=9, first=3):
    """eg_*_batch_seeded generate their randomness in the kernel (ChaCha20, key = seed, block counter = counter_base +
    (i << 20) + draw): byte-identical to the oracle's provers on the same per-item streams (SURVEY.md 8(d) layout), to the
    caller-supplied form fed with those blocks, and independent of the chunking."""
    seed, base = W.SEED_CHOICE, first << 20
    # encrypt_bool
    ocs, ops = O.gen_bool_batch(pk, seed, n, first=first)
    values = np.array([(first + i) & 1 for i in range(n)], np.uint8)
    cts, proofs = e.encrypt_bool(values, seed=seed, counter_base=base)
    assert (cts == ocs).all() and (proofs == ops).all()
    # EncryptedChoice::single, also with tiny chunks (item0 of later chunks)
    options = 4
    ocs, ors, oss = O.gen_choice_batch(pk, options, seed, n, first=first)
    values = np.zeros((n, options), np.uint8)
    for i in range(n):
        values[i, (first + i) % options] = 1
    for chunk in (0, 4):
        e.set_chunk_items(chunk)
        try:
            cts, rings, sums = e.encrypt_choice(options, values, single=True, seed=seed, counter_base=base)
        finally:
            e.set_chunk_items(0)
        assert (cts == ocs).all() and (rings == ors).all() and (sums == oss).all()
    # RangeProof::new
    spec = O.range_optimal(100)
    espec = to_engine_range(e, spec)
    vals = np.array([(17 * (first + i)) % 100 for i in range(n)], np.uint64)
    oc, op, orr = O.gen_range_batch(pk, spec, "ciphertext_range", seed, vals, first=first)
    for chunk in (0, 4):
        e.set_chunk_items(chunk)
        try:
            c, p, r = e.encrypt_range(espec, "ciphertext_range", vals, seed=seed, counter_base=base)
        finally:
            e.set_chunk_items(0)
        assert (c == oc).all() and (p == op).all() and (r == orr).all()
    # QuadraticVotingBallot::new (records of several proofs: block0 / group addressing)
    p_, ep = O.qv_params(5, 20), e.qv_params(5, 20)
    votes = np.array([QV_VOTES[(first + i) % 4] for i in range(n)], np.uint64)
    ob = O.gen_qv_batch(pk, p_, W.SEED_QV, votes, first=first)
    for chunk in (0, 4):
        e.set_chunk_items(chunk)
        try:
            b = e.encrypt_qv(ep, votes, seed=W.SEED_QV, counter_base=base)
        finally:
            e.set_chunk_items(0)
        assert (b == ob).all()
    assert (e.verify_qv(ep, b)[0] == 0).all()
    # encrypt / encrypt_zero: equal to the caller-supplied form on the same blocks
    vals = np.arange(n, dtype=np.uint64) * 1000
    wide = np.frombuffer(b"".join(item_blocks(seed, first + i, 1) for i in range(n)), np.uint8)
    assert (e.encrypt(vals, seed=seed, counter_base=base) == e.encrypt(vals, wide)).all()
    wide = np.frombuffer(b"".join(item_blocks(seed, first + i, 2) for i in range(n)), np.uint8)
    c1, p1 = e.encrypt_zero(seed=seed, counter_base=base, n=n)
    c2, p2 = e.encrypt_zero(wide)
    assert (c1 == c2).all() and (p1 == p2).all() and (e.verify_zero(c1, p1) == 0).all()


def check_seeded_reference_snapshots(e):
    """counter_base = 1 and the ChaCha20 key of `ChaChaRng::seed_from_u64(12345)` (PCG32 expansion): the
    seeded provers regenerate the reference's own snapshots (tests/snapshots.rs:73-161) -- block 0 of that stream is the
    receiver's secret key, the object's draws follow."""
    gold, hx = _gold(), bytes.fromhex
    rng = O.rng_from_u64(12345)
    key = bytes(rng.key)
    sk, pk = O.keypair(rng)

    def ctb(d):
        return hx(d["random_element"]) + hx(d["blinded_element"])
    try:
        e.set_receiver(pk)
        cts, proofs = e.encrypt_bool(np.array([1], np.uint8), seed=key, counter_base=1)
        assert cts.tobytes() == ctb(gold["bool-encryption"]["ciphertext"]) and proofs.tobytes() == hx(gold["bool-encryption-bin"])
        cts, rings, sums = e.encrypt_choice(5, np.array([[0, 0, 0, 1, 0]], np.uint8), single=True, seed=key, counter_base=1)
        g = gold["encrypted-choice"]
        assert cts.tobytes() == b"".join(ctb(c) for c in g["choices"])
        assert rings.tobytes() == hx(g["range_proof"]["common_challenge"]) + b"".join(hx(x) for x in g["range_proof"]["ring_responses"])
        assert sums.tobytes() == hx(g["sum_proof"]["challenge"]) + hx(g["sum_proof"]["response"])
        cts, rings, _ = e.encrypt_choice(5, np.array([[0, 1, 1, 0, 1]], np.uint8), single=False, seed=key, counter_base=1)
        g = gold["encrypted-multi-choice"]
        assert cts.tobytes() == b"".join(ctb(c) for c in g["choices"])
        assert rings.tobytes() == hx(g["range_proof"]["common_challenge"]) + b"".join(hx(x) for x in g["range_proof"]["ring_responses"])
        espec = e.range_optimal(100)
        c, p, r = e.encrypt_range(espec, "ciphertext_range", np.array([42], np.uint64), seed=key, counter_base=1)
        g = gold["range-encryption"]
        assert bytes(c[0]) == ctb(g["ciphertext"])
        assert bytes(p[0].reshape(-1)) == b"".join(ctb(x) for x in g["proof"]["partial_ciphertexts"])
        assert bytes(r[0].reshape(-1)) == hx(g["proof"]["common_challenge"]) + b"".join(hx(x) for x in g["proof"]["ring_responses"])
        ep = e.qv_params(5, 15)
        ballot = bytes(e.encrypt_qv(ep, np.array([[3, 0, 1, 0, 2]], np.uint64), seed=key, counter_base=1)[0])
        g = gold["qv-ballot"]

        def rp(d):
            pr = d["range_proof"]
            return (ctb(d["ciphertext"]) + b"".join(ctb(x) for x in pr["partial_ciphertexts"]) + hx(pr["common_challenge"])
                    + b"".join(hx(x) for x in pr["ring_responses"]))
        ce = g["credit_equivalence_proof"]
        assert ballot == (b"".join(rp(v) for v in g["votes"]) + rp(g["credit"]) + hx(ce["challenge"])
                          + b"".join(hx(x) for x in ce["ciphertext_responses"]) + hx(ce["sum_response"]))
        ct = bytes(e.encrypt(np.array([42], np.uint64), seed=key, counter_base=1)[0])
        assert ct == hx(gold["ciphertext-bin"])
        cts, proofs = e.encrypt_zero(seed=key, counter_base=1, n=1)
        assert bytes(cts[0]) == ctb(gold["zero-encryption"]["ciphertext"]) and bytes(proofs[0]) == hx(gold["zero-encryption-bin"])
    finally:
        e.set_receiver(W.receiver()[1])


def check_constant_time_prover_mode(e, pk, n=6):
    """eg_ctx_set_prover_mode(1) changes how secret scalars walk the fixed-base tables (64 masked 4-bit windows), never
    the group elements: every prover output is byte-identical to the default mode's."""
    seed = W.SEED_CHOICE
    spec = to_engine_range(e, O.range_optimal(21))
    ep = e.qv_params(3, 9)
    rvals = np.array([(5 * i) % 21 for i in range(n)], np.uint64)
    votes = np.array([[(i + k) % 2 for k in range(3)] for i in range(n)], np.uint64)
    cvals = np.zeros((n, 3), np.uint8)
    for i in range(n):
        cvals[i, i % 3] = 1

    def run():
        return [a for out in (e.encrypt_bool(np.array([i & 1 for i in range(n)], np.uint8), seed=seed),
                              e.encrypt_choice(3, cvals, single=True, seed=seed),
                              e.encrypt_range(spec, "ciphertext_range", rvals, seed=seed),
                              (e.encrypt_qv(ep, votes, seed=seed),),
                              (e.encrypt(np.array([0, 1, 2**40, 7, 8, 9][:n], np.uint64), seed=seed),),
                              e.encrypt_zero(seed=seed, n=n)) for a in out]
    fast = run()
    e.set_prover_mode(True)
    try:
        slow = run()
    finally:
        e.set_prover_mode(False)
    assert len(fast) == len(slow) and all((a == b).all() for a, b in zip(fast, slow))
    assert (e.verify_bool(slow[0], slow[1]) == 0).all()


def check_single_choice_validation(e, pk):
    """eg_encrypt_choice_batch(single != 0) rejects rows that do not mark exactly one option (EncryptedChoice::single cannot
    produce them, choice.rs:288-306) instead of emitting ballots that fail verification."""
    from elastic_elgamal_b200 import EngineError, _ffi
    for row in ([0, 0, 0], [1, 1, 0]):
        try:
            e.encrypt_choice(3, np.array([[1, 0, 0], row], np.uint8), single=True, seed=W.SEED_CHOICE)
        except EngineError as exc:
            assert exc.status == _ffi.ERR_INVALID_ARG
        else:
            raise AssertionError("a malformed single-choice row was accepted")
    e.encrypt_choice(3, np.array([[1, 1, 0]], np.uint8), single=False, seed=W.SEED_CHOICE)     # fine for MultiChoice


# ---------------------------------------------------------------- SumOfSquaresProof::verify / CandidateDecryption::verify

def _sumsq_instance(pk, values, rng, label):
    def enc_with_value(v):
        peek = O.Rng.from_buffer_copy(bytes(rng))
        r = O.scalar_reduce_wide(O.rng_block(peek))        # CiphertextWithValue::new draws exactly one block
        return O.encrypt(pk, v, rng), r
    sum_ct, sum_r = enc_with_value(sum(v * v for v in values))
    pairs = [enc_with_value(v) for v in values]
    proof = O.sumsq_prove(pk, [p[0] for p in pairs], [v.to_bytes(32, "little") for v in values], [p[1] for p in pairs], sum_ct, sum_r,
                          label, rng)
    return b"".join(p[0] for p in pairs), sum_ct, proof


def check_verify_sumsq(e, pk, n=16, count=5, label="test", seed=b"\x0d" * 32):
    """eg_verify_sumsq_batch against the oracle's SumOfSquaresProof::verify, with the reference's tamper patterns
    (mul.rs:332-361,418-437): responses of another proof, a swapped ciphertext, a wrong sum ciphertext, a wrong label."""
    rng = O.rng_from_seed(seed)
    rnd = random.Random(count)
    rows = [_sumsq_instance(pk, [rnd.randrange(6) for _ in range(count)], rng, label) for _ in range(n)]
    cts = np.frombuffer(b"".join(r[0] for r in rows), np.uint8).reshape(n, count, 64).copy()
    sums = np.frombuffer(b"".join(r[1] for r in rows), np.uint8).reshape(n, 64).copy()
    proofs = np.frombuffer(b"".join(r[2] for r in rows), np.uint8).reshape(n, 2 * count + 2, 32).copy()
    if n >= 8:
        proofs[1] = proofs[2]
        if count >= 2:
            cts[3, 0], cts[3, 1] = cts[3, 1].copy(), cts[3, 0].copy()
        sums[4] = sums[5]
        proofs[6, 2] = np.frombuffer(W.BAD_SCALAR, np.uint8)
        cts[7, 0, :32] = np.frombuffer(W.BAD_POINT2, np.uint8)
    expected = [O.sumsq_verify(pk, [bytes(c) for c in cts[i]], bytes(sums[i]), label, bytes(proofs[i].reshape(-1))) for i in range(n)]
    got = e.verify_sumsq(label, cts, sums, proofs)
    assert got.tolist() == expected, (got.tolist(), expected)
    if n >= 8:
        assert expected[0] == O.OK and expected[1] == O.CHALLENGE_MISMATCH and expected[6] == O.MALFORMED and expected[7] == O.MALFORMED
        assert expected[3] in (O.OK, O.CHALLENGE_MISMATCH)      # swapping two equal votes' ciphertexts is still a valid statement
    assert (e.verify_sumsq(label + "x", cts, sums, proofs) != 0).all()
    assert e.verify_sumsq(label, cts[:0], sums[:0], proofs[:0]).shape == (0,)


def check_verify_sumsq_reference_snapshot(e):
    """The reference's `sum-sq-proof` snapshot (tests/snapshots.rs:132-151: values [1, 3, 3, 7, 5], seed 12345) verifies."""
    g = _gold()["sum-sq-proof"]
    rng = O.rng_from_u64(12345)
    sk, pk = O.keypair(rng)
    cts, sum_ct, proof = _sumsq_instance(pk, [1, 3, 3, 7, 5], rng, "test")
    hx = bytes.fromhex
    assert proof == hx(g["challenge"]) + b"".join(hx(x) for x in g["ciphertext_responses"]) + hx(g["sum_response"])
    try:
        e.set_receiver(pk)
        v = e.verify_sumsq("test", np.frombuffer(cts, np.uint8).reshape(1, 5, 64), np.frombuffer(sum_ct, np.uint8).reshape(1, 64),
                           np.frombuffer(proof, np.uint8).reshape(1, 12, 32))
        assert v.tolist() == [0]
    finally:
        e.set_receiver(W.receiver()[1])


def check_verify_decryption(e, n=14, label="custom_key_decryption", seed=b"\x0e" * 32):
    """eg_verify_decryption_batch (CandidateDecryption::verify with a custom key, decryption.rs:189-205) against the oracle."""
    from elastic_elgamal_b200 import EngineError, _ffi
    rng = O.rng_from_seed(seed)
    sk, pk = O.keypair(rng)
    sk2, pk2 = O.keypair(rng)
    cts = [O.encrypt(pk2, 3 * i, rng) for i in range(n)]
    rows = [O.decryption_prove(sk, label, ct, rng) for ct in cts]
    cts_a = np.frombuffer(b"".join(cts), np.uint8).reshape(n, 64).copy()
    dh_a = np.frombuffer(b"".join(r[0] for r in rows), np.uint8).reshape(n, 32).copy()
    pr_a = np.frombuffer(b"".join(r[1] for r in rows), np.uint8).reshape(n, 64).copy()
    if n >= 8:
        pr_a[1] = pr_a[2]                                             # proof of another ciphertext
        dh_a[3] = dh_a[4]                                             # decryption of another ciphertext
        dh_a[5] = np.frombuffer(W.BAD_POINT, np.uint8)                # CandidateDecryption::from_bytes -> None
        pr_a[6, 32:] = np.frombuffer(W.BAD_SCALAR2, np.uint8)
        cts_a[7, 32:] = np.frombuffer(O.point_add(bytes(cts_a[7, 32:]), W.G_ENC), np.uint8)     # B is not committed: still verifies
    expected = [O.decryption_verify(pk, label, bytes(cts_a[i]), bytes(dh_a[i]), bytes(pr_a[i])) for i in range(n)]
    got = e.verify_decryption(label, pk, cts_a, dh_a, pr_a)
    assert got.tolist() == expected, (got.tolist(), expected)
    if n >= 8:
        assert expected[0] == O.OK and expected[1] == O.CHALLENGE_MISMATCH and expected[3] == O.CHALLENGE_MISMATCH
        assert expected[5] == O.MALFORMED and expected[6] == O.MALFORMED and expected[7] == O.OK
    assert (e.verify_decryption(label, pk2, cts_a, dh_a, pr_a) != 0).all()          # another key
    assert (e.verify_decryption(label + "2", pk, cts_a, dh_a, pr_a) != 0).all()     # another transcript
    for bad in (W.BAD_POINT, bytes(32)):                                           # undecodable / identity key
        try:
            e.verify_decryption(label, bad, cts_a, dh_a, pr_a)
        except EngineError as exc:
            assert exc.status == _ffi.ERR_INVALID_ELEMENT
        else:
            raise AssertionError("an invalid key was accepted")


# ---------------------------------------------------------------- struct-level wire format (VecHelper bounds)

def check_wire_objects(e, pk):
    """eg_wire_fields / eg_wire_decode_batch: the base64url strings of the reference's own human-readable snapshots
    (tests/snapshots/*.snap, held in tests/golden/) decode, object by object, to the flat layouts the batch entry points
    take -- and those verify; the VecHelper minimum lengths (serde.rs:303-355) are enforced; one bad field rejects its
    whole object."""
    import base64
    from elastic_elgamal_b200 import EngineError, _ffi
    F = e.wire_fields
    assert F(_ffi.WIRE_CIPHERTEXT) == 2 and F(_ffi.WIRE_LOG_EQUALITY_PROOF) == 2 and F(_ffi.WIRE_DECRYPTION) == 1
    assert F(_ffi.WIRE_COMMITMENT_EQUIV_PROOF) == 4
    assert [F(_ffi.WIRE_RING_PROOF, c) for c in (0, 1, 2, 10)] == [0, 0, 3, 11]             # ring.rs:285: at least 2 scalars
    assert [F(_ffi.WIRE_POSSESSION_PROOF, c) for c in (0, 1, 5)] == [0, 2, 6]              # possession.rs:74: at least 1
    assert [F(_ffi.WIRE_SUMSQ_PROOF, c) for c in (0, 1, 2, 10)] == [0, 0, 4, 12]           # mul.rs:89: at least 2
    assert F(99, 5) == 0
    try:
        e.wire_decode(np.zeros((1, 0), np.uint8), 0)
    except EngineError as exc:
        assert exc.status == _ffi.ERR_LEN_MISMATCH
    else:
        raise AssertionError("zero fields accepted")

    def b64(hexstr):
        return base64.urlsafe_b64encode(bytes.fromhex(hexstr)).rstrip(b"=")
    gold = _gold()
    g = gold["encrypted-choice"]
    # EncryptedChoice (5 options) = 5 ciphertexts + ring proof (10 responses) + log-equality proof: 10 + 11 + 2 fields
    fields = 5 * F(_ffi.WIRE_CIPHERTEXT) + F(_ffi.WIRE_RING_PROOF, 10) + F(_ffi.WIRE_LOG_EQUALITY_PROOF)
    strings = ([s for c in g["choices"] for s in (b64(c["random_element"]), b64(c["blinded_element"]))]
               + [b64(g["range_proof"]["common_challenge"])] + [b64(x) for x in g["range_proof"]["ring_responses"]]
               + [b64(g["sum_proof"]["challenge"]), b64(g["sum_proof"]["response"])])
    assert len(strings) == fields == 23 and all(len(s) == 43 for s in strings)
    text = np.frombuffer(b"".join(strings), np.uint8).reshape(1, fields * 43)
    batch = np.tile(text, (4, 1)).copy()
    batch[1, 43 * 7 + 5] = ord("=")                  # padding inside field 7 of object 1
    batch[2, 43 * 22 + 42] = ord("B")                # non-zero trailing bits in the last field of object 2
    raw, ok = e.wire_decode(batch, fields)
    assert ok.tolist() == [True, False, False, True]
    assert (e.wire_encode(raw[[0, 3]], fields) == batch[[0, 3]]).all()
    rng = O.rng_from_u64(12345)
    sk, snap_pk = O.keypair(rng)
    try:
        e.set_receiver(snap_pk)
        obj = raw[[0, 3]]
        v, t = e.verify_choice(5, obj[:, :320], obj[:, 320:672], obj[:, 672:])
        assert v.tolist() == [0, 0]
    finally:
        e.set_receiver(pk)


def check_prove_range_from_ciphertext(e, pk, upper_bound=100, n=9, label="ciphertext_range", seed=b"\x0f" * 32):
    """eg_prove_range_batch (RangeProof::from_ciphertext, range.rs:482-534): given the randomness RangeProof::new would have
    drawn for the ciphertext (block 0 of the item's stream) and the remaining blocks, the proof is byte-identical to
    RangeProof::new's -- which is `CiphertextWithValue::new` followed by `from_ciphertext` (range.rs:462-473)."""
    spec = O.range_optimal(upper_bound)
    espec = to_engine_range(e, spec)
    values = np.array([(37 * i) % upper_bound for i in range(n)], np.uint64)
    oc, op, orr = O.gen_range_batch(pk, spec, label, seed, values)
    draws = e.lib.eg_range_prover_draws(O.C.byref(espec))
    blocks = [item_blocks(seed, i, draws) for i in range(n)]
    ct_r = np.frombuffer(b"".join(O.scalar_reduce_wide(b[:64]) for b in blocks), np.uint8).reshape(n, 32)
    wide = np.frombuffer(b"".join(b[64:] for b in blocks), np.uint8)
    c, p, r = e.prove_range(espec, label, values, ct_r, wide)
    assert (c == oc).all() and (p == op).all() and (r == orr).all()
    c2, p2, r2 = e.prove_range(espec, label, values, ct_r, seed=seed, counter_base=1, want_cts=False)      # draws start at block 1
    assert c2 is None and (p2 == op).all() and (r2 == orr).all()
    assert (e.verify_range(espec, label, oc, p2, r2) == 0).all()
    from elastic_elgamal_b200 import EngineError, _ffi
    bad = ct_r.copy()
    bad[0] = np.frombuffer(W.BAD_SCALAR, np.uint8)
    try:
        e.prove_range(espec, label, values, bad, wide)
    except EngineError as exc:
        assert exc.status == _ffi.ERR_INVALID_ARG
    else:
        raise AssertionError("non-canonical ciphertext randomness accepted")


# ---------------------------------------------------------------- device-pointer provers

def check_prover_dev_forms(e, pk, up, down, new, n=9):
    """eg_encrypt_{bool,choice,range}_batch_dev: the same bytes as the host forms, with the seed and with blocks resident
    in "device" memory.  up(array) -> device buffer, new(shape) -> empty uint8 device buffer, down(buffer) -> numpy; a buffer
    exposes its address as `.ptr` (torch on the GPU, numpy in the CPU-compiled harness where device == host)."""
    seed = _u8s(W.SEED_CHOICE)
    # bool, seeded
    values = np.array([i & 1 for i in range(n)], np.uint8)
    hc, hp = e.encrypt_bool(values, seed=W.SEED_CHOICE, counter_base=5 << 20)
    d_v, d_c, d_p = up(values), new((n, 64)), new((n, 96))
    e._check(e.lib.eg_encrypt_bool_batch_dev(e.h, n, d_v.ptr, None, seed.ctypes.data, 5 << 20, d_c.ptr, d_p.ptr))
    assert (down(d_c) == hc).all() and (down(d_p) == hp).all()
    # choice (single), blocks resident on the device
    m = 4
    cv = np.zeros((n, m), np.uint8)
    cv[np.arange(n), np.arange(n) % m] = 1
    wide = np.frombuffer(b"".join(item_blocks(W.SEED_CHOICE, i, 3 * m + 1) for i in range(n)), np.uint8).reshape(n, 3 * m + 1, 64)
    hc, hr, hs = e.encrypt_choice(m, cv, wide, single=True)
    d_v, d_w, d_c, d_r, d_s = up(cv), up(wide), new((n, m, 64)), new((n, 1 + 2 * m, 32)), new((n, 64))
    e._check(e.lib.eg_encrypt_choice_batch_dev(e.h, n, m, 1, d_v.ptr, d_w.ptr, None, 0, d_c.ptr, d_r.ptr, d_s.ptr))
    assert (down(d_c) == hc).all() and (down(d_r) == hr).all() and (down(d_s) == hs).all()
    oc, orr, oss = O.gen_choice_batch(pk, m, W.SEED_CHOICE, n)
    assert (hc == oc).all() and (hr == orr).all() and (hs == oss).all()
    # range, seeded, tiny chunks
    spec = to_engine_range(e, O.range_optimal(100))
    rv = np.array([(29 * i) % 100 for i in range(n)], np.uint64)
    hc, hp, hr = e.encrypt_range(spec, "ciphertext_range", rv, seed=W.SEED_CHOICE)
    d_v, d_c, d_p, d_r = up(rv.view(np.uint8)), new(hc.shape), new(hp.shape), new(hr.shape)
    e.set_chunk_items(4)
    try:
        e._check(e.lib.eg_encrypt_range_batch_dev(e.h, O.C.byref(spec), b"ciphertext_range", n, d_v.ptr, None, seed.ctypes.data, 0,
                                                  d_c.ptr, d_p.ptr, d_r.ptr))
    finally:
        e.set_chunk_items(0)
    assert (down(d_c) == hc).all() and (down(d_p) == hp).all() and (down(d_r) == hr).all()
    assert (e.verify_range(spec, "ciphertext_range", hc, hp, hr) == 0).all()


def _u8s(b):
    return np.frombuffer(bytes(b), np.uint8).copy()
