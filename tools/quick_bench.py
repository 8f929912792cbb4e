#!/usr/bin/env python3
"""Device-resident throughput of eg_verify_choice_batch_dev for one or more library builds (tuning experiments).
usage: python tools/quick_bench.py [--ballots N] lib1.so [lib2.so ...]"""
import argparse, json, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import numpy as np, torch
import oracle as O, workloads as W
from elastic_elgamal_b200 import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--ballots", type=int, default=1 << 18)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("libs", nargs="+")
a = ap.parse_args()
sk, pk = W.receiver()
cts, rings, sums = O.gen_choice_batch(pk, 5, W.SEED_CHOICE, 2048)
ov, ot = O.verify_choice_batch(pk, 5, True, cts, rings, sums)
reps = a.ballots // 2048
d = [torch.from_numpy(np.tile(x, (reps,) + (1,) * (x.ndim - 1))).cuda() for x in (cts, rings, sums)]
B = reps * 2048
dv = torch.empty(B, dtype=torch.uint8, device="cuda"); dt = torch.empty((5, 64), dtype=torch.uint8, device="cuda")
for lib in a.libs:
    e = Engine(device=0, lib_path=lib); e.set_receiver(pk)
    run = lambda: e.verify_choice_dev(B, 5, True, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), dv.data_ptr(), dt.data_ptr())
    run(); run()
    t0 = time.perf_counter()
    cm = 0.0
    for _ in range(a.steps):
        run(); cm += e.last_commit_stats()["ms"]
    torch.cuda.synchronize(); dtm = time.perf_counter() - t0
    ok = bool((dv.cpu().numpy().reshape(reps, 2048) == ov[None]).all())
    print(json.dumps({"lib": pathlib.Path(lib).name, "ballots_per_s": B * a.steps / dtm, "ms_per_step": 1e3 * dtm / a.steps,
                      "commit_ms_per_step": cm / a.steps, "verdicts_ok": ok, "timings": e.last_timings()}))
    e.close()
