#!/usr/bin/env python3
"""One-off large differential fuzz of the GPU verifiers against the oracle (tests/parity_common.check_fuzz_differential
at sizes beyond the test suite), under every ring-engine setting.  usage: python tools/big_fuzz.py [n] [seeds]"""
import pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import workloads as W, parity_common as PC
from elastic_elgamal_b200 import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
e = Engine(device=0)
sk, pk = W.receiver()
e.set_receiver(pk)
for mode in (0, 2, 1):
    e.set_ring_mode(mode)
    for seed in range(100, 100 + seeds):
        t0 = time.perf_counter()
        PC.check_fuzz_differential(e, pk, n=n, seed=seed + 10 * mode)
        print("ring mode %d seed %d n=%d: all verdicts and tallies equal the oracle's (%.1f s)" % (mode, seed + 10 * mode, n, time.perf_counter() - t0), flush=True)
