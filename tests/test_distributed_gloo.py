"""World-size-2 `gloo` test of the sharding logic (elastic_elgamal_b200/distributed.py) on CPU.

Each rank drives the host-compiled harness build of the library (tests/hostsim, TEST HARNESS ONLY) exactly as it would
drive the CUDA library on its own GPU: verify a contiguous slice, all_gather the partial tallies, add them.  The
result must equal the oracle's single-process verdicts and tally bit for bit."""
import os
import pathlib
import random
import socket
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, options, out_dir):
    for p in (str(ROOT), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    import oracle as O
    import workloads as W
    from elastic_elgamal_b200 import Engine
    from elastic_elgamal_b200 import distributed as D
    from hostsim.build_hostsim import build
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        e = Engine(lib_path=build())
        sk, pk = W.receiver()
        e.set_receiver(pk)
        cts, rings, sums = O.gen_choice_batch(pk, options, W.SEED_CHOICE, n)
        cts, rings, sums = cts.copy(), rings.copy(), sums.copy()
        W.tamper_choice(cts, rings, sums, random.Random(11), frac=0.3)
        v, (lo, hi), total = D.verify_choice_sharded(e, options, cts, rings, sums, single=True, dist=dist)
        allv = D.gather_verdicts(v, n, dist)
        np.save(os.path.join(out_dir, f"v{rank}.npy"), allv)
        np.save(os.path.join(out_dir, f"t{rank}.npy"), total)
        np.save(os.path.join(out_dir, f"b{rank}.npy"), np.array([lo, hi]))
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    from elastic_elgamal_b200.distributed import shard_bounds
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_two_rank_choice_tally(tmp_path):
    import torch.multiprocessing as mp

    import oracle as O
    import workloads as W
    from hostsim.build_hostsim import build
    build()                                    # compile once, before the workers race for it
    n, options, world = 11, 3, 2               # odd size: ragged shards (6 + 5)
    mp.spawn(_worker, args=(world, _free_port(), n, options, str(tmp_path)), nprocs=world, join=True)
    sk, pk = W.receiver()
    cts, rings, sums = O.gen_choice_batch(pk, options, W.SEED_CHOICE, n)
    cts, rings, sums = cts.copy(), rings.copy(), sums.copy()
    W.tamper_choice(cts, rings, sums, random.Random(11), frac=0.3)
    ov, ot = O.verify_choice_batch(pk, options, True, cts, rings, sums)
    assert len(set(ov.tolist())) > 1            # the tampering produced rejected ballots
    for r in range(world):
        assert (np.load(tmp_path / f"v{r}.npy") == ov).all()
        assert (np.load(tmp_path / f"t{r}.npy") == ot).all()
    assert np.load(tmp_path / "b0.npy").tolist() == [0, 6] and np.load(tmp_path / "b1.npy").tolist() == [6, 11]
