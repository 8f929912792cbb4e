#!/usr/bin/env python3
"""Device-memory footprint of a context (INTEGRATION.md "What a context costs"): free memory before / after eg_ctx_create,
eg_ctx_set_receiver, a 262 144-ballot verification and a share verification with key tables."""
import pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import numpy as np
import torch
import oracle as O, workloads as W
from elastic_elgamal_b200 import Engine

def used():
    free, total = torch.cuda.mem_get_info(0)
    return (total - free) / 2**30

torch.cuda.init()
base = used()
t0 = time.perf_counter(); e = Engine(device=0); t1 = time.perf_counter()
print("after eg_ctx_create: +%.2f GiB (%.0f ms)" % (used() - base, 1e3 * (t1 - t0)))
sk, pk = W.receiver()
t0 = time.perf_counter(); e.set_receiver(pk); t1 = time.perf_counter()
print("after eg_ctx_set_receiver: +%.2f GiB (%.0f ms)" % (used() - base, 1e3 * (t1 - t0)))
cc, cr, cs = O.gen_choice_batch(pk, 5, W.SEED_CHOICE, 512)
tile = lambda a, n: np.ascontiguousarray(np.tile(a, ((n + a.shape[0] - 1) // a.shape[0],) + (1,) * (a.ndim - 1))[:n])
n = 600000
e.verify_choice(5, tile(cc, n), tile(cr, n), tile(cs, n))
print("after verifying %d five-option ballots (two chunks + a remainder): +%.2f GiB" % (n, used() - base))
e.close()
print("after eg_ctx_destroy: +%.2f GiB" % (used() - base))
