#!/usr/bin/env python3
"""Wall time of eg_encrypt_range_batch (host buffers) for one batch; run under
`ncu --metrics gpu__time_duration.sum --clock-control none --csv` to compare with the sum of its kernels."""
import ctypes as C, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import numpy as np
import workloads as W
from elastic_elgamal_b200 import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
e = Engine(device=0)
sk, pk = W.receiver()
e.set_receiver(pk)
rng = e.range_optimal(65536)
draws = e.lib.eg_range_prover_draws(C.byref(rng))
wide = np.random.default_rng(1).integers(0, 256, (n, draws, 64), dtype=np.uint8)
values = (np.arange(n, dtype=np.uint64) * 40503) % 65536
e.encrypt_range(rng, "range", values[:1024], wide[:1024])
for _ in range(2):
    t0 = time.perf_counter()
    cts, partials, rings = e.encrypt_range(rng, "range", values, wide)
    dt = time.perf_counter() - t0
    print("encrypt_range n=%d: %.1f ms wall, %.0f proofs/s, %d draws/item" % (n, dt * 1e3, n / dt, draws))
