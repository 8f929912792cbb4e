"""A non-Python consumer of the C ABI: tests/cabi/consumer.c is compiled with gcc against include/eg_b200.h, linked to
libeg_b200.so and replays the reference's `encrypted-choice` snapshot plus a seeded config-2 mini batch.  Without a GPU
(`-m "not gpu"`) the program must build, link, load and refuse to compute (exit code 77: no CPU fallback)."""
import json
import pathlib
import random
import struct
import subprocess

import numpy as np
import pytest

import oracle as O
import workloads as W

ROOT = pathlib.Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "cabi" / "consumer.c"


def build_consumer(tmp_path):
    from elastic_elgamal_b200 import build
    lib = build.build()
    exe = tmp_path / "consumer"
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"), "-o", str(exe), str(SRC),
                    str(lib), f"-Wl,-rpath,{lib.parent}"], check=True)
    return exe


def case(options, single, key, cts, rings, sums, verdicts, tally):
    n = cts.shape[0]
    return (struct.pack("<IIQ", options, 1 if single else 0, n) + bytes(key) + cts.tobytes() + rings.tobytes()
            + (sums.tobytes() if sums is not None else bytes(64 * n)) + verdicts.astype(np.uint8).tobytes() + tally.tobytes())


def vectors(tmp_path):
    hx = bytes.fromhex
    g = json.loads((ROOT / "tests" / "golden" / "ristretto_snapshots.json").read_text())["encrypted-choice"]
    rng = O.rng_from_u64(12345)
    sk, pk = O.keypair(rng)                       # tests/snapshots.rs:32-34
    cts = np.frombuffer(b"".join(hx(c["random_element"]) + hx(c["blinded_element"]) for c in g["choices"]), np.uint8).reshape(1, 5, 64)
    rings = np.frombuffer(hx(g["range_proof"]["common_challenge"]) + b"".join(hx(x) for x in g["range_proof"]["ring_responses"]),
                          np.uint8).reshape(1, 11, 32)
    sums = np.frombuffer(hx(g["sum_proof"]["challenge"]) + hx(g["sum_proof"]["response"]), np.uint8).reshape(1, 64)
    blob = case(5, True, pk, cts, rings, sums, np.zeros(1, np.uint8), cts[0])       # accepted; tally of one ballot = the ballot
    sk2, pk2 = W.receiver()
    n = 257
    c, r, s = O.gen_choice_batch(pk2, 5, W.SEED_CHOICE, n)
    c, r, s = c.copy(), r.copy(), s.copy()
    W.tamper_choice(c, r, s, random.Random(7), frac=0.1)
    v, t = O.verify_choice_batch(pk2, 5, True, c, r, s)
    blob += case(5, True, pk2, c, r, s, v, t)
    c, r, _ = O.gen_choice_batch(pk2, 3, W.SEED_CHOICE, 0)                         # an empty batch: identity tally
    v, t = O.verify_choice_batch(pk2, 3, True, c, r, np.zeros((0, 64), np.uint8))
    blob += case(3, True, pk2, c, r, np.zeros((0, 64), np.uint8), v, t)
    path = tmp_path / "vectors.bin"
    path.write_bytes(blob)
    return path


def test_consumer_builds_links_and_refuses_without_a_gpu(tmp_path):
    import torch
    exe = build_consumer(tmp_path)
    res = subprocess.run([str(exe), str(vectors(tmp_path))], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert "eg_b200" in res.stdout and "sm_100a" in res.stdout
    if torch.cuda.is_available():
        assert res.returncode == 0, res.stderr
    else:
        assert res.returncode == 77, (res.returncode, res.stdout, res.stderr)


@pytest.mark.gpu
def test_consumer_replays_snapshot_and_mini_batch(tmp_path):
    exe = build_consumer(tmp_path)
    res = subprocess.run([str(exe), str(vectors(tmp_path))], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert res.returncode == 0, (res.stdout, res.stderr)
    assert "3 cases, 0 failures" in res.stdout
