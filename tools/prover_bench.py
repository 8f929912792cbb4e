#!/usr/bin/env python3
"""tools/prover_bench.py -- the creation side through the host C ABI: caller-supplied randomness blocks (64 B per draw over
PCIe) vs in-kernel ChaCha20 (eg_*_batch_seeded) vs the constant-time prover mode, for encrypt_bool, EncryptedChoice::single
(5 options), RangeProof::new [0, 2^16) and QuadraticVotingBallot::new (5 / 20).  Every emitted batch is verified by the GPU
verifier; the first items are compared byte for byte with the oracle's provers on the same streams.

    python tools/prover_bench.py [--items 262144] [--out profiles/rN_provers.json]
"""
import argparse
import ctypes as C
import json
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import numpy as np  # noqa: E402

import oracle as O  # noqa: E402
import parity_common as PC  # noqa: E402
import workloads as W  # noqa: E402
from elastic_elgamal_b200 import Engine  # noqa: E402


def timed(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=1 << 18)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    n = args.items
    e = Engine(device=0)
    sk, pk = W.receiver()
    e.set_receiver(pk)
    seed = W.SEED_CHOICE
    rs = np.random.default_rng(1)
    res = {}

    def run(name, draws, fn_wide, fn_seed, verify, oracle_check):
        wide = rs.integers(0, 256, (n, draws, 64), dtype=np.uint8)
        t_wide, _ = timed(lambda: fn_wide(wide))
        t_seed, out = timed(fn_seed)
        assert verify(out)
        oracle_check(out)
        e.set_prover_mode(True)
        try:
            t_ct, out_ct = timed(fn_seed, reps=1)
        finally:
            e.set_prover_mode(False)
        assert all((a == b).all() for a, b in zip(out, out_ct))
        res[name] = {"items": n, "draws_per_item": draws, "randomness_bytes_per_item": draws * 64,
                     "caller_blocks_per_s": n / t_wide, "seeded_per_s": n / t_seed, "seeded_constant_time_per_s": n / t_ct}
        print(name, json.dumps(res[name]), flush=True)

    values = (np.arange(n) & 1).astype(np.uint8)
    ob = O.gen_bool_batch(pk, seed, 32)
    run("encrypt_bool", 3, lambda w: e.encrypt_bool(values, w), lambda: e.encrypt_bool(values, seed=seed),
        lambda o: (e.verify_bool(*o) == 0).all(), lambda o: (o[0][:32] == ob[0]).all() and (o[1][:32] == ob[1]).all() or (_ for _ in ()).throw(AssertionError("bool")))
    cv = np.zeros((n, 5), np.uint8)
    cv[np.arange(n), np.arange(n) % 5] = 1
    oc = O.gen_choice_batch(pk, 5, seed, 32)
    run("encrypted_choice_single_5", 16, lambda w: e.encrypt_choice(5, cv, w), lambda: e.encrypt_choice(5, cv, seed=seed),
        lambda o: (e.verify_choice(5, *o)[0] == 0).all(), lambda o: all((a[:32] == b).all() for a, b in zip(o, oc)) or (_ for _ in ()).throw(AssertionError("choice")))
    spec = O.range_optimal(65536)
    espec = PC.to_engine_range(e, spec)
    rv = (np.arange(n, dtype=np.uint64) * 40503) % 65536
    orr = O.gen_range_batch(pk, spec, "ciphertext_range", seed, rv[:32])
    run("range_proof_2_16", e.lib.eg_range_prover_draws(C.byref(espec)), lambda w: e.encrypt_range(espec, "ciphertext_range", rv, w),
        lambda: e.encrypt_range(espec, "ciphertext_range", rv, seed=seed),
        lambda o: (e.verify_range(espec, "ciphertext_range", *o) == 0).all(),
        lambda o: all((a[:32] == b).all() for a, b in zip(o, orr)) or (_ for _ in ()).throw(AssertionError("range")))
    p, ep = O.qv_params(5, 20), e.qv_params(5, 20)
    votes = np.array([PC.QV_VOTES[i % 4] for i in range(n)], np.uint64)
    oq = O.gen_qv_batch(pk, p, W.SEED_QV, votes[:32])
    run("qv_ballot_5_20", e.lib.eg_qv_prover_draws(C.byref(ep)), lambda w: (e.encrypt_qv(ep, votes, w),),
        lambda: (e.encrypt_qv(ep, votes, seed=W.SEED_QV),), lambda o: (e.verify_qv(ep, o[0])[0] == 0).all(),
        lambda o: (o[0][:32] == oq).all() or (_ for _ in ()).throw(AssertionError("qv")))
    # device-resident forms (values in, objects out, all in HBM; in-kernel randomness): the kernels' own rate
    import torch
    dev = torch.device("cuda", 0)
    key = np.frombuffer(seed, np.uint8).copy()

    def dev_rate(fn):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        return n * 3 / (time.perf_counter() - t0)
    d_vals = torch.from_numpy(values).to(dev)
    d_c, d_p = torch.empty((n, 64), dtype=torch.uint8, device=dev), torch.empty((n, 96), dtype=torch.uint8, device=dev)
    res["encrypt_bool"]["seeded_device_resident_per_s"] = dev_rate(lambda: e._check(e.lib.eg_encrypt_bool_batch_dev(
        e.h, n, d_vals.data_ptr(), None, key.ctypes.data, 0, d_c.data_ptr(), d_p.data_ptr())))
    assert (e.verify_bool(d_c.cpu().numpy(), d_p.cpu().numpy()) == 0).all()
    d_cv = torch.from_numpy(cv).to(dev)
    d_cc, d_cr, d_cs = (torch.empty(s_, dtype=torch.uint8, device=dev) for s_ in ((n, 5, 64), (n, 11, 32), (n, 64)))
    res["encrypted_choice_single_5"]["seeded_device_resident_per_s"] = dev_rate(lambda: e._check(e.lib.eg_encrypt_choice_batch_dev(
        e.h, n, 5, 1, d_cv.data_ptr(), None, key.ctypes.data, 0, d_cc.data_ptr(), d_cr.data_ptr(), d_cs.data_ptr())))
    d_rv = torch.from_numpy(rv.view(np.int64)).to(dev)
    d_rc, d_rp, d_rr = (torch.empty(s_, dtype=torch.uint8, device=dev) for s_ in ((n, 64), (n, 7, 64), (n, 33, 32)))
    res["range_proof_2_16"]["seeded_device_resident_per_s"] = dev_rate(lambda: e._check(e.lib.eg_encrypt_range_batch_dev(
        e.h, C.byref(espec), b"ciphertext_range", n, d_rv.data_ptr(), None, key.ctypes.data, 0, d_rc.data_ptr(), d_rp.data_ptr(), d_rr.data_ptr())))
    assert (e.verify_range(espec, "ciphertext_range", d_rc.cpu().numpy(), d_rp.cpu().numpy(), d_rr.cpu().numpy()) == 0).all()
    print("device-resident:", {k: v.get("seeded_device_resident_per_s") for k, v in res.items()}, flush=True)
    if args.out:
        pathlib.Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
