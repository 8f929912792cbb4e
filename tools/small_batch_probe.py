#!/usr/bin/env python3
"""Latency of small batches (BASELINE configs[0]: 10 000 Boolean ciphertexts) under the three ring-engine settings:
0 = chosen per chunk (library default), 1 = per-equation pipeline, 2 = one thread per ring (k_ring), 3 = two lanes per ring
(k_ring_pair)."""
import pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import numpy as np
import oracle as O, workloads as W, parity_common as PC
from elastic_elgamal_b200 import Engine

e = Engine(device=0)
sk, pk = W.receiver()
e.set_receiver(pk)


def tile(a, n):
    reps = (n + a.shape[0] - 1) // a.shape[0]
    return np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1))[:n])


def run(name, fn, sizes):
    for n in sizes:
        row = []
        for mode in (0, 1, 2, 3):
            e.set_ring_mode(mode)
            fn(n)
            t0 = time.perf_counter()
            for _ in range(5):
                fn(n)
            row.append((time.perf_counter() - t0) / 5)
        print("%-28s n=%7d  auto %.2f ms  pipeline %.2f ms  k_ring %.2f ms  pair %.2f ms  (%.0f items/s auto)" % (name, n, row[0] * 1e3, row[1] * 1e3, row[2] * 1e3, row[3] * 1e3, n / row[0]))
    e.set_ring_mode(0)


bc, bp = O.gen_bool_batch(pk, W.SEED_CHOICE, 1024)
run("verify_bool", lambda n: e.verify_bool(tile(bc, n), tile(bp, n)), (1000, 5000, 10000, 20000, 30000, 40000, 60000, 80000, 100000, 140000, 200000))
cc, cr, cs = O.gen_choice_batch(pk, 5, W.SEED_CHOICE, 1024)
run("verify_choice (5 options)", lambda n: e.verify_choice(5, tile(cc, n), tile(cr, n), tile(cs, n)), (1000, 2500, 5000, 7500, 10000, 15000, 20000, 30000, 60000))
spec = O.range_optimal(65536)
rng = PC.to_engine_range(e, spec)
rc, rp, rr = O.gen_range_batch(pk, spec, "range", W.SEED_QV, (np.arange(256, dtype=np.uint64) * 40503) % 65536)
run("verify_range [0, 2^16)", lambda n: e.verify_range(rng, "range", tile(rc, n), tile(rp, n), tile(rr, n)), (500, 1000, 2000, 3000, 5000, 10000, 20000, 40000))
