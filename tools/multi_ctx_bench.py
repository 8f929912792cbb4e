#!/usr/bin/env python3
"""tools/multi_ctx_bench.py -- ONE process, all GPUs of the box through eg_ctx_create_multi: the host hands the whole
batch (host buffers) to eg_verify_choice_batch; the library shards it across its per-device child contexts (one host
thread each), combines the partial tallies with a grouped ncclAllGather + point-add kernel and returns verdicts in input
order.  This is what a single Rust process replacing the loop of examples/voting.rs:188-204 gets.  End-to-end numbers only
(H2D + D2H inside), pinned and pageable host buffers.

    python tools/multi_ctx_bench.py --gpus 8 [--ballots-per-gpu 1048576] [--total 0] [--steps 3]
"""
import argparse
import json
import pathlib
import random
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import numpy as np  # noqa: E402

import oracle as O  # noqa: E402
import workloads as W  # noqa: E402
from elastic_elgamal_b200 import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--ballots-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--total", type=int, default=0, help="total ballots (strong scaling); 0 = ballots-per-gpu x gpus")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--unique", type=int, default=4096)
    args = ap.parse_args()
    import torch
    n = args.total or args.ballots_per_gpu * args.gpus
    sk, pk = W.receiver()
    cts, rings, sums = O.gen_choice_batch(pk, 5, W.SEED_CHOICE, args.unique, threads=O.hw_threads())
    cts, rings, sums = cts.copy(), rings.copy(), sums.copy()
    W.tamper_choice(cts, rings, sums, random.Random(2), frac=0.01)
    ov, _ = O.verify_choice_batch(pk, 5, True, cts, rings, sums, threads=O.hw_threads())
    reps = (n + args.unique - 1) // args.unique
    big = [np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1))[:n]) for a in (cts, rings, sums)]
    ev = np.tile(ov, reps)[:n]
    e = Engine(devices=list(range(args.gpus)))
    e.set_receiver(pk)
    e.set_ring_mode(2)

    def pin(a):
        t = torch.empty(a.shape, dtype=torch.uint8, pin_memory=True)
        t.numpy()[...] = a
        return t
    pinned = [pin(a) for a in big]
    out = {}
    for name, arrs in (("pinned", [t.numpy() for t in pinned]), ("pageable", big)):
        v, t = e.verify_choice(5, *arrs)          # warm-up (tables, scratch, NCCL channels)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            v, t = e.verify_choice(5, *arrs)
        dt = (time.perf_counter() - t0) / args.steps
        assert (v == ev).all()
        out[name] = n / dt
    table = O.DlogTable(0, n + 1) if n <= (1 << 23) else None
    if table is not None:
        idx = np.arange(n)
        for k in range(5):
            assert table.get(O.decrypt_to_element(sk, bytes(t[k]))) == int(np.count_nonzero((ev == 0) & ((idx % args.unique) % 5 == k)))
    print(json.dumps({"what": "eg_ctx_create_multi: one process, host buffers in, verdicts + combined tally out", "n_gpus": args.gpus,
                      "ballots": n, "e2e_pinned_ballots_per_s": out["pinned"], "e2e_pageable_ballots_per_s": out["pageable"],
                      "steps": args.steps, "kernel_launches": e.kernel_launches, "comm": e.comm_info()}))


if __name__ == "__main__":
    main()
