#!/usr/bin/env python3
"""verify_range over 47 360 proofs for [0, 2^16) (5 waves of the long-ring k_ring shape); run under ncu with
-k 'regex:^k_ring$' to capture k_ring<256, 2, 8> (profiles/r1_k_ring_long_range_s9.txt)."""
import sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import numpy as np
import oracle as O, workloads as W, parity_common as PC
from elastic_elgamal_b200 import Engine
e = Engine(device=0); sk, pk = W.receiver(); e.set_receiver(pk)
spec = O.range_optimal(65536); rng = PC.to_engine_range(e, spec)
rc, rp, rr = O.gen_range_batch(pk, spec, "range", W.SEED_QV, (np.arange(256, dtype=np.uint64) * 40503) % 65536)
n = 47360
t = lambda a: np.ascontiguousarray(np.tile(a, (n // 256,) + (1,) * (a.ndim - 1)))
c, p, r = t(rc), t(rp), t(rr)
for _ in range(2):
    t0 = time.perf_counter(); v = e.verify_range(rng, "range", c, p, r); print(n / (time.perf_counter() - t0), int((v == 0).sum()))
