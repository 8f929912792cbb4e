#!/usr/bin/env python3
"""tools/bench_configs.py -- throughput of every BASELINE.json config through the host C ABI (the `e2e` shape of
bench.py: host buffers in, verdicts / tallies / values out, H2D + D2H inside the timed region).

    python tools/bench_configs.py [--configs 1,2,3,4,5] [--scale 1.0] [--out profiles/rN_configs.json]

bench.py stays the driver-facing headline (config 2).  This tool gives the other configs the same treatment:
    1  encrypt_bool + RingProof verify of Boolean ciphertexts            (160 B/item)
    2  EncryptedChoice::single, 5 options, verify + tally                 (736 B/ballot)
    3  QuadraticVotingBallot, 5 options / 20 credits, verify + tally      (2144 B/ballot)
    4  RangeProof for [0, 2^16) (8 rings x 4), verify                     (1568 B/proof)
    5  3-of-5 decryption shares (LogEqualityProof) verify + combine + DiscreteLogTable lookup (352 B/tally)
`unique` oracle-generated items (1 % tampered where a tamper helper exists) are tiled to the batch size; verdicts of
the timed batch are compared with the oracle's verdicts of the unique items, tallies/values with their expected
values.  The CPU column is the oracle port on one thread and on all threads (a reported baseline only).
"""
import argparse
import json
import pathlib
import random
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

import oracle as O  # noqa: E402
import parity_common as PC  # noqa: E402
import workloads as W  # noqa: E402
from elastic_elgamal_b200 import Engine  # noqa: E402

# reference-equivalent field operations per unit
REF_FIELD_OPS = {1: 12.1e3, 2: 66.8e3, 3: 256e3, 4: 185.5e3, 5: 27e3}


PINNED = False      # --pinned: input batches live in page-locked host memory (what a serving process would use)
_keep = []


def tile(a, n):
    reps = (n + a.shape[0] - 1) // a.shape[0]
    out = np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1))[:n])
    if PINNED and out.dtype == np.uint8 and out.nbytes >= (1 << 20):
        import torch
        t = torch.empty(out.shape, dtype=torch.uint8, pin_memory=True)
        t.numpy()[...] = out
        _keep.append(t)
        return t.numpy()
    return out


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fn()
    return (time.perf_counter() - t0) / steps, out


def cpu_rates(fn_single, n_single, fn_all, n_all):
    t0 = time.perf_counter(); fn_single(); t1 = time.perf_counter(); fn_all(); t2 = time.perf_counter()
    return n_single / (t1 - t0), n_all / (t2 - t1)


def config1(e, pk, sk, n, unique, steps, warmup, threads):
    cts, proofs = O.gen_bool_batch(pk, W.SEED_CHOICE, unique, threads=threads)
    cts, proofs = cts.copy(), proofs.copy()
    W.tamper_bool(cts, proofs, random.Random(1), frac=0.01)
    ov = O.verify_bool_batch(pk, cts, proofs, threads=threads)
    C, P = tile(cts, n), tile(proofs, n)
    dt, v = timed(lambda: e.verify_bool(C, P), steps, warmup)
    assert (v == tile(ov, n)).all()
    # the encryption half of the config: GPU prover on the same count, checked by the GPU verifier
    rs = np.random.RandomState(1)
    values = (np.arange(n) & 1).astype(np.uint8)
    wide = rs.randint(0, 256, (n, 3, 64)).astype(np.uint8)
    dt_enc, (ec, ep) = timed(lambda: e.encrypt_bool(values, wide), max(1, steps // 2), 1)
    assert (e.verify_bool(ec, ep) == 0).all()
    s1, sa = cpu_rates(lambda: O.verify_bool_batch(pk, cts[:256], proofs[:256], threads=1), 256,
                       lambda: O.verify_bool_batch(pk, cts, proofs, threads=threads), unique)
    return {"unit": "verified ciphertexts/s", "bytes_per_item": 160, "gpu_e2e": n / dt, "gpu_encrypt_bool_e2e": n / dt_enc,
            "cpu_1t": s1, "cpu_all": sa, "rejected": int((ov != 0).sum())}


def config2(e, pk, sk, n, unique, steps, warmup, threads):
    cts, rings, sums = O.gen_choice_batch(pk, 5, W.SEED_CHOICE, unique, threads=threads)
    cts, rings, sums = cts.copy(), rings.copy(), sums.copy()
    W.tamper_choice(cts, rings, sums, random.Random(2), frac=0.01)
    ov, _ = O.verify_choice_batch(pk, 5, True, cts, rings, sums, threads=threads)
    C, R, S = tile(cts, n), tile(rings, n), tile(sums, n)
    dt, (v, t) = timed(lambda: e.verify_choice(5, C, R, S), steps, warmup)
    ev = tile(ov, n)
    assert (v == ev).all()
    table = O.DlogTable(0, n + 1)
    idx = np.arange(n)
    for k in range(5):
        assert table.get(O.decrypt_to_element(sk, bytes(t[k]))) == int(np.count_nonzero((ev == 0) & ((idx % unique) % 5 == k)))
    s1, sa = cpu_rates(lambda: O.verify_choice_batch(pk, 5, True, cts[:128], rings[:128], sums[:128], threads=1), 128,
                       lambda: O.verify_choice_batch(pk, 5, True, cts, rings, sums, threads=threads), unique)
    return {"unit": "verified ballots/s", "bytes_per_item": 736, "gpu_e2e": n / dt, "cpu_1t": s1, "cpu_all": sa,
            "rejected": int((ov != 0).sum())}


def config3(e, pk, sk, n, unique, steps, warmup, threads):
    p, ep = O.qv_params(5, 20), e.qv_params(5, 20)
    votes = np.array([PC.QV_VOTES[i % 4] for i in range(unique)], np.uint64)
    ballots = O.gen_qv_batch(pk, p, W.SEED_QV, votes, threads=threads).copy()
    bsz = ballots.shape[1]
    rnd = random.Random(3)
    for k, i in enumerate(sorted(rnd.sample(range(unique), max(1, unique // 100)))):
        if k % 3 == 0:
            ballots[i, 32:64] = np.frombuffer(O.point_add(bytes(ballots[i, 32:64]), W.G_ENC), np.uint8)
        elif k % 3 == 1:
            ballots[i, -32 * 12:] = ballots[(i + 1) % unique, -32 * 12:]
        else:
            ballots[i, -1] = 0xff
    ov, _ = O.verify_qv_batch(pk, p, ballots, threads=threads)
    Bt = tile(ballots, n)
    dt, (v, t) = timed(lambda: e.verify_qv(ep, Bt), steps, warmup)
    ev = tile(ov, n)
    assert (v == ev).all()
    table = O.DlogTable(0, 4 * n + 1)
    vt = tile(votes, n)
    for k in range(5):
        assert table.get(O.decrypt_to_element(sk, bytes(t[k]))) == int(vt[ev == 0, k].sum())
    s1, sa = cpu_rates(lambda: O.verify_qv_batch(pk, p, ballots[:32], threads=1), 32,
                       lambda: O.verify_qv_batch(pk, p, ballots, threads=threads), unique)
    # the creation side (QuadraticVotingBallot::new) on a quarter of the count, accepted by the GPU verifier
    ne = max(64, n // 4)
    draws = e.lib.eg_qv_prover_draws(O.C.byref(ep))
    wide = np.random.RandomState(3).randint(0, 256, (ne, draws, 64)).astype(np.uint8)
    dt_enc, made = timed(lambda: e.encrypt_qv(ep, vt[:ne], wide), 1, 1)
    assert (e.verify_qv(ep, made)[0] == 0).all()
    t0 = time.perf_counter(); O.gen_qv_batch(pk, p, W.SEED_QV, votes, threads=threads); cpu_enc = unique / (time.perf_counter() - t0)
    return {"unit": "verified QV ballots/s", "bytes_per_item": int(bsz), "gpu_e2e": n / dt, "cpu_1t": s1, "cpu_all": sa,
            "rejected": int((ov != 0).sum()), "gpu_encrypt_qv_e2e": ne / dt_enc, "cpu_all_encrypt_qv": cpu_enc,
            "prover_randomness_bytes_per_item": int(draws * 64)}


def config4(e, pk, sk, n, unique, steps, warmup, threads):
    spec = O.range_optimal(65536)
    espec = PC.to_engine_range(e, spec)
    values = (np.arange(unique, dtype=np.uint64) * 40503) % 65536
    cts, partials, rings = O.gen_range_batch(pk, spec, "ciphertext_range", W.SEED_CHOICE, values, threads=threads)
    cts, partials, rings = cts.copy(), partials.copy(), rings.copy()
    PC.tamper_range(cts, partials, rings, random.Random(4), 0.01)
    ov = O.verify_range_batch(pk, spec, "ciphertext_range", cts, partials, rings, threads=threads)
    C, P, R = tile(cts, n), tile(partials, n), tile(rings, n)
    dt, v = timed(lambda: e.verify_range(espec, "ciphertext_range", C, P, R), steps, warmup)
    assert (v == tile(ov, n)).all()
    s1, sa = cpu_rates(lambda: O.verify_range_batch(pk, spec, "ciphertext_range", cts[:32], partials[:32], rings[:32], threads=1), 32,
                       lambda: O.verify_range_batch(pk, spec, "ciphertext_range", cts, partials, rings, threads=threads), unique)
    # the creation side (encrypt_range = RangeProof::new) on a quarter of the count, accepted by the GPU verifier
    ne = max(64, n // 4)
    draws = e.lib.eg_range_prover_draws(O.C.byref(espec))
    wide = np.random.RandomState(4).randint(0, 256, (ne, draws, 64)).astype(np.uint8)
    ev = (np.arange(ne, dtype=np.uint64) * 40503) % 65536
    dt_enc, (c2, p2, r2) = timed(lambda: e.encrypt_range(espec, "ciphertext_range", ev, wide), 1, 1)
    assert (e.verify_range(espec, "ciphertext_range", c2, p2, r2) == 0).all()
    t0 = time.perf_counter(); O.gen_range_batch(pk, spec, "ciphertext_range", W.SEED_CHOICE, values, threads=threads); cpu_enc = unique / (time.perf_counter() - t0)
    return {"unit": "verified range proofs/s", "bytes_per_item": 1568, "gpu_e2e": n / dt, "cpu_1t": s1, "cpu_all": sa,
            "rejected": int((ov != 0).sum()), "decomposition": O.range_display(spec), "gpu_encrypt_range_e2e": ne / dt_enc,
            "cpu_all_encrypt_range": cpu_enc, "prover_randomness_bytes_per_item": int(draws * 64)}


def config5(e, pk, sk, n, unique, steps, warmup, threads):
    rng = O.rng_from_seed(bytes([9] * 32))
    ks, secrets = O.dealer_new(5, 3, rng)
    eks = PC.as_engine_keyset(ks)
    shared = bytes(ks.shared_key)
    used = (0, 2, 4)
    table_hi = 1 << 20
    unique = min(unique, 1024)          # the share prover runs item by item through ctypes
    rnd = random.Random(5)
    values = [rnd.randrange(table_hi) for _ in range(unique)]
    cts = [O.encrypt(shared, v, rng) for v in values]
    rows = [[O.decrypt_share(ks, i, secrets[i], ct, rng) for i in used] for ct in cts]
    cts_a = np.frombuffer(b"".join(cts), np.uint8).reshape(unique, 64).copy()
    sh_a = np.frombuffer(b"".join(b"".join(r[0] for r in row) for row in rows), np.uint8).reshape(unique, 3, 32).copy()
    pr_a = np.frombuffer(b"".join(b"".join(r[1] for r in row) for row in rows), np.uint8).reshape(unique, 3, 64).copy()
    for k, i in enumerate(sorted(rnd.sample(range(unique), max(1, unique // 100)))):
        pr_a[i, k % 3] = pr_a[(i + 1) % unique, k % 3]
    t0 = time.perf_counter()
    ov = np.array([[O.verify_share(ks, used[j], bytes(cts_a[i]), bytes(sh_a[i, j]), bytes(pr_a[i, j])) for j in range(3)]
                   for i in range(unique)], np.uint8)
    t_cpu = time.perf_counter() - t0
    C, S, P = tile(cts_a, n), tile(sh_a, n), tile(pr_a, n)
    t0 = time.perf_counter()
    table = e.dlog_table(0, table_hi)
    t_table = time.perf_counter() - t0

    def step():
        v = e.verify_shares(eks, list(used), C, S, P)
        vals, found = e.combine_decrypt(list(used), C, S, table)
        return v, vals, found
    dt, (v, vals, found) = timed(step, steps, warmup)
    assert (v == tile(ov, n)).all()
    assert (found == 1).all() and (vals == tile(np.array(values, np.uint64), n)).all()
    table.close()
    return {"unit": "tallies/s (3 shares verified + combined + dlog lookup)", "bytes_per_item": 352, "gpu_e2e": n / dt,
            "cpu_1t": unique / t_cpu, "cpu_all": None, "cpu_note": "share verification only, through per-item ctypes calls",
            "dlog_table_entries": table_hi, "dlog_table_build_s": t_table, "rejected_shares": int((ov != 0).sum())}


CONFIGS = {1: (config1, 1 << 20), 2: (config2, 1 << 20), 3: (config3, 1 << 18), 4: (config4, 1 << 18), 5: (config5, 1 << 20)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--scale", type=float, default=1.0, help="multiplies the default batch sizes")
    ap.add_argument("--unique", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--out", default="")
    ap.add_argument("--pinned", action="store_true", help="keep the input batches in page-locked host memory")
    args = ap.parse_args()
    global PINNED
    PINNED = args.pinned
    e = Engine(device=0)
    sk, pk = W.receiver()
    e.set_receiver(pk)
    threads = O.hw_threads()
    results = {}
    for c in [int(x) for x in args.configs.split(",")]:
        fn, n0 = CONFIGS[c]
        n = max(64, int(n0 * args.scale))
        l0 = e.kernel_launches
        r = fn(e, pk, sk, n, args.unique, args.steps, args.warmup, threads)
        r.update({"config": c, "items": n, "unique": args.unique, "cpu_threads": threads, "pinned_inputs": bool(args.pinned),
                  "gpu_launches": e.kernel_launches - l0,
                  "ref_equiv_field_ops_per_item": REF_FIELD_OPS[c], "ref_equiv_field_ops_per_s": REF_FIELD_OPS[c] * r["gpu_e2e"],
                  "host_gbs": r["gpu_e2e"] * r["bytes_per_item"] / 1e9})
        results[c] = r
        print(json.dumps(r), flush=True)
    if args.out:
        pathlib.Path(args.out).write_text(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
