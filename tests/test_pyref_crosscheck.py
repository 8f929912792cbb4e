"""Cross-check of the C oracle against tests/pyref.py, an independent pure-Python restatement, for the two objects that
have no reference-held golden vector (ProofOfPossession, PublicKeySet::from_participants): same proof bytes from the same
ChaCha20 blocks, same verdicts on valid and tampered inputs, same reconstructed shared keys.  The primitives pyref is
built on are pinned first (RFC 9496 vectors, Merlin's published test vector, a reference snapshot)."""
import json
import pathlib
import random

import numpy as np

import oracle as O
import pyref as R
import workloads as W

GOLD = json.loads((pathlib.Path(__file__).parent / "golden" / "ristretto_snapshots.json").read_text())


def test_pyref_primitives_are_pinned():
    # RFC 9496 A.1: multiples of the generator
    assert R.Point.identity().encode() == bytes(32)
    assert R.G.encode().hex() == "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76"
    assert (R.G * 2).encode().hex() == "6a493210f7499cd17fecb510ae0cea23a110e8d5b901f8acadd3095c73a3b919"
    assert (R.G * 15).encode().hex() == "e0c418f7c8d9c4cdd7395b93ea124f3ad99021bb681dfc3302a9d99a2e53e64e"
    assert R.Point.decode(bytes.fromhex("e0c418f7c8d9c4cdd7395b93ea124f3ad99021bb681dfc3302a9d99a2e53e64e")) == R.G * 15
    # RFC 9496 A.3: invalid encodings
    for bad in ("00ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff", "0100000000000000000000000000000000000000000000000000000000000000",
                "26948d35ca62e643e26a83177332e6b6afeb9d08e4268b650f1f5bbd8d81d371", "ed57ffd8c914fb201471d1c3d245ce3c746fcbe63a3679d51b6a516ebebe0e20"):
        assert R.Point.decode(bytes.fromhex(bad)) is None
    # merlin 3.0.0 `test_transcript::equivalence_simple`
    t = R.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    # and a reference snapshot that exercises group + transcript together: the `zero-encryption` proof verifies
    g = GOLD["zero-encryption"]
    rng = O.rng_from_u64(12345)
    sk, pk = O.keypair(rng)
    Rp, Bp = R.Point.decode(bytes.fromhex(g["ciphertext"]["random_element"])), R.Point.decode(bytes.fromhex(g["ciphertext"]["blinded_element"]))
    c, s = int(bytes.fromhex(g["proof"]["challenge"])[::-1].hex(), 16), int(bytes.fromhex(g["proof"]["response"])[::-1].hex(), 16)
    K = R.Point.decode(pk)
    t = R.Transcript(b"zero_encryption")                      # keys/impls.rs:67 -> log_equality.rs:153-180
    t.start_proof(b"log_eq")
    t.append_message(b"K", pk)
    t.append_element(b"[r]G", Rp)
    t.append_element(b"[r]K", Bp)
    t.append_element(b"[x]G", Rp * ((-c) % R.L) + R.G * s)
    t.append_element(b"[x]K", Bp * ((-c) % R.L) + K * s)
    assert t.challenge_scalar(b"c") == c


def test_proof_of_possession_oracle_equals_pyref():
    label = "test_multi_PoP"
    for k in (1, 2, 5):
        rng = O.rng_from_seed(bytes([11] * 32), first_block=k << 20)
        pairs = [O.keypair(rng) for _ in range(k)]
        secrets, keys = [p[0] for p in pairs], [p[1] for p in pairs]
        peek = O.Rng.from_buffer_copy(bytes(rng))
        blocks = [O.rng_block(peek) for _ in range(k)]
        proof = O.pop_prove(secrets, keys, label, rng)
        assert proof == R.pop_prove([int.from_bytes(s, "little") for s in secrets], keys, label.encode(), blocks)
        assert O.pop_verify(keys, label, proof) == R.pop_verify(keys, label.encode(), proof) == R.OK
        rnd = random.Random(k)
        for _ in range(6):                                    # possession.rs:55-66 negative cases + random corruption
            bad_keys, bad_proof = [bytearray(x) for x in keys], bytearray(proof)
            kind = rnd.randrange(5)
            if kind == 0:
                bad_proof[rnd.randrange(len(bad_proof))] ^= 1 << rnd.randrange(8)
            elif kind == 1:
                bad_keys[rnd.randrange(k)] = bytearray(O.keypair(rng)[1])
            elif kind == 2:
                bad_keys[rnd.randrange(k)] = bytearray(W.BAD_POINT2)
            elif kind == 3:
                bad_keys[rnd.randrange(k)] = bytearray(32)    # the identity is not a public key
            else:
                bad_proof[32:64] = W.BAD_SCALAR
            bk = [bytes(x) for x in bad_keys]
            assert O.pop_verify(bk, label, bytes(bad_proof)) == R.pop_verify(bk, label.encode(), bytes(bad_proof)) != R.OK
        assert O.pop_verify(keys, label + "x", proof) == R.pop_verify(keys, (label + "x").encode(), proof) == R.CHALLENGE_MISMATCH


def test_keyset_from_participants_oracle_equals_pyref():
    for shares, threshold in ((3, 2), (5, 3), (5, 5), (6, 4), (4, 1)):
        rng = O.rng_from_seed(bytes([13] * 32), first_block=(shares * 16 + threshold) << 20)
        ks, secrets = O.dealer_new(shares, threshold, rng)
        keys = [bytes(ks.participant_keys[i]) for i in range(shares)]
        v, shared = R.keyset_from_participants(shares, threshold, keys)
        ov, oshared = O.keyset_from_participants(shares, threshold, keys)
        assert v == ov == R.OK and shared == oshared == bytes(ks.shared_key)
        if 1 < threshold < shares:            # (threshold 1: a constant polynomial, all participant keys are equal)
            swapped = keys[:]
            swapped[-1] = swapped[0]              # (a plain swap of the end points of a line is again a line)
            assert R.keyset_from_participants(shares, threshold, swapped)[0] == O.keyset_from_participants(shares, threshold, swapped)[0] \
                == R.MALFORMED_PARTICIPANT_KEYS
        broken = keys[:]
        broken[shares - 1] = W.BAD_POINT
        assert R.keyset_from_participants(shares, threshold, broken)[0] == O.keyset_from_participants(shares, threshold, broken)[0] == R.MALFORMED
