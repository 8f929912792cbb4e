"""One rank of the one-process-per-GPU path (eg_ctx_attach_comm), started by tests/test_gpu_multi.py and by
tools/strong_scaling.py: python comm_worker.py RANK WORLD ID_FILE OUT_FILE N OPTIONS.  Rank 0 creates the NCCL unique id
inside the library (eg_comm_unique_id) and publishes it through ID_FILE -- the "any channel the host already has" of
include/eg_b200.h; no torch.distributed anywhere.  Each rank verifies its contiguous slice of the same seeded batch and
writes its verdicts and the (global) tally to OUT_FILE."""
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    rank, world, id_file, out_file, n, options = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
    import oracle as O
    import workloads as W
    from elastic_elgamal_b200 import Engine
    from elastic_elgamal_b200.distributed import shard_bounds
    import random
    e = Engine(device=rank)
    if rank == 0:
        uid = e.comm_unique_id()
        tmp = id_file + ".tmp"
        with open(tmp, "wb") as f:
            f.write(uid.tobytes())
        os.replace(tmp, id_file)
    else:
        for _ in range(1200):
            if os.path.exists(id_file):
                break
            time.sleep(0.05)
        uid = np.frombuffer(open(id_file, "rb").read(), np.uint8)
    e.attach_comm(uid, rank, world)
    assert e.comm_info() == {"rank": rank, "world": world, "devices": 1}
    sk, pk = W.receiver()
    e.set_receiver(pk)
    e.set_ring_mode(2)
    cts, rings, sums = O.gen_choice_batch(pk, options, W.SEED_CHOICE, n)
    cts, rings, sums = cts.copy(), rings.copy(), sums.copy()
    W.tamper_choice(cts, rings, sums, random.Random(5), frac=0.1)
    lo, hi = shard_bounds(n, world, rank)
    v, t = e.verify_choice(options, cts[lo:hi], rings[lo:hi], sums[lo:hi])
    # a second collective call with an empty slice on the last rank: shards of zero items contribute the identity
    lo2, hi2 = (lo, hi) if rank + 1 < world or world == 1 else (hi, hi)
    v2, t2 = e.verify_choice(options, cts[lo2:hi2], rings[lo2:hi2], sums[lo2:hi2])
    np.savez(out_file, verdicts=v, tally=t, lo=lo, hi=hi, tally2=t2, lo2=lo2, hi2=hi2)
    e.close()


if __name__ == "__main__":
    main()
