#!/usr/bin/env python3
"""bench.py -- headline benchmark: verified 5-option EncryptedChoice ballots / s (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W                     (N > 1: launched under torchrun by the driver)
    python bench.py --impl reference --gpus N --steps K --warmup W    (the CPU restatement of the reference path)

Workload = BASELINE.json configs[1]: "single-choice polling: EncryptedChoice::single 5 options, batch verify 1M
ballots + homomorphic tally".  One step = one pass of the hot path (decode -> sum proof + ring proof verification
-> verdicts -> masked tally) over one batch of `--ballots` ballots per GPU (weak scaling: ballots are independent,
the batch shards across ranks; the only exchange is the per-rank partial tally, combined after an all_gather).

Synthetic data: `--unique` distinct ballots are produced by the oracle's prover from the seeded ChaCha streams of
SURVEY.md 8(d) (1 % tampered with the reference's tamper patterns) and tiled to the batch size; verification cost
does not depend on ballot contents (uniform control flow), and verdicts/tally are checked against the oracle.

JSON keys follow the driver contract; `value` = device-resident throughput, `e2e` = through the host C ABI with
pinned host buffers (H2D + D2H inside the timed region), `roofline` = the dominant kernel (k_ring: one thread per ring
proof, all of its equations) against the INT32 multiply-add issue rate measured live by
tools/microbench/int_pipe_bench, `cpu_baseline` = the oracle port timed on this box's host cores.
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

OPTIONS = 5
BALLOT_BYTES = OPTIONS * 64 + (1 + 2 * OPTIONS) * 32 + 64        # 736, SURVEY.md 8(a) a12
# Reference-equivalent algorithmic work: one verification-equation side = half of a
# (double-base + 2-term) pair, 4956 / 2 field operations, plus one compression of 280; 144 IMAD-class
# instructions per field operation.
FIELD_OPS_PER_COMMIT = 4956 / 2 + 280
FIELD_OPS_PER_BALLOT = 4956 * 11 + 280 * (34 + 10)               # 66 836
# What k_ring actually executes per equation side of a two-equation ring, counted from the formulas in
# ge.cuh: 2 x 1603 table build (192 doublings + 28 additions per point) + 4 sides x (435 for 60 doublings + 497 for 64
# per-item additions + 112 for 16 fixed-base additions from the wide table) + 112 for the [e a]G term + 304 for encoding
# the first equation's pair (the last equation's points are encoded by k_terminal, outside this kernel).
EXECUTED_FIELD_OPS_PER_RING_SIDE = (2 * 1603 + 4 * (435 + 497 + 112) + 112 + 304) / 4      # 1949.5
IMAD_PER_FIELD_OP = 144
METRIC = "verified ballots/sec (5-option choice)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ballots", type=int, default=1 << 20, help="ballots per GPU per step")
    ap.add_argument("--unique", type=int, default=4096, help="distinct oracle-generated ballots that are tiled")
    ap.add_argument("--cpu-sample", type=int, default=0, help="ballots in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ring-mode", type=int, default=2, choices=[1, 2], help="2: k_ring (default); 1: per-equation k_commit launches")
    ap.add_argument("--chunk", type=int, default=0, help="ballots per internal chunk (0 = library default)")
    return ap.parse_args()


def make_workload(unique, threads=0):
    import random

    import numpy as np

    import oracle as O
    import workloads as W
    sk, pk = W.receiver()
    cts, rings, sums = O.gen_choice_batch(pk, OPTIONS, W.SEED_CHOICE, unique, threads=threads)
    cts, rings, sums = cts.copy(), rings.copy(), sums.copy()
    W.tamper_choice(cts, rings, sums, random.Random(2), frac=0.01)
    return sk, pk, cts, rings, sums


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_baseline(pk, cts, rings, sums, sample, threads):
    """Times the oracle (CPU restatement of the reference path) on `sample` ballots with `threads` host threads."""
    import numpy as np

    import oracle as O
    reps = max(1, sample // cts.shape[0])
    c, r, s = (np.tile(cts, (reps, 1, 1))[:sample], np.tile(rings, (reps, 1, 1))[:sample], np.tile(sums, (reps, 1))[:sample])
    t0 = time.perf_counter()
    v, t = O.verify_choice_batch(pk, OPTIONS, True, c, r, s, threads=threads)
    dt = time.perf_counter() - t0
    return c.shape[0] / dt, dt, v, t


def run_reference(args):
    """`--impl reference`: the reference crate is Rust and cannot be built in this image (no cargo/rustc), so the
    arm times the oracle port of the same path on all host cores (cpu_baseline.kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    threads = O.hw_threads()
    sk, pk, cts, rings, sums = make_workload(min(args.unique, 2048), threads)
    # bounded sample per step: ~2 s of work on all cores
    single, _, _, _ = cpu_baseline(pk, cts, rings, sums, 256, 1)
    sample = args.cpu_sample or max(256, int(single * threads * 2.0) // 64 * 64)
    for _ in range(args.warmup):
        cpu_baseline(pk, cts, rings, sums, max(64, sample // 8), threads)
    t_tot, n_tot = 0.0, 0
    for _ in range(args.steps):
        rate, dt, v, t = cpu_baseline(pk, cts, rings, sums, sample, threads)
        t_tot += dt
        n_tot += sample
    value = n_tot / t_tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "ballots/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (GF(2^255-19))", "data": "synthetic seeded ballots (oracle prover)",
        "config": {"workload": "EncryptedChoice::single 5 options: verify + homomorphic tally (BASELINE configs[1])",
                   "ballots_per_step": sample, "note": "bounded sample of the 1M-ballot workload; units are independent"},
        "cpu_baseline": {"value": value, "unit": "ballots/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} ballots/step x {args.steps} steps, {threads} threads; single thread {single:.1f} ballots/s"},
        "e2e": {"value": value, "unit": "ballots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


def hbm_roofline(sides, kernel_ms, traffic_per_launch, launches, step_stream_gbs):
    """{"bound": "hbm", ...} for the dominant kernel against MEASURED_PEAKS.json (fallback 6650 GB/s, B200_PROFILING.md)."""
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
        src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        pass
    bytes_per_side = (2 * 128 + 2 * 32 + 2.2 * 32 + 2 * 32) / 4.0          # per equation side (4 sides per ring)
    achieved = sides * bytes_per_side / (kernel_ms * 1e-3) / 1e9 if kernel_ms else None
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
            "traffic": traffic_per_launch, "peak_source": src,
            "actual_dram_gbs": (traffic_per_launch * launches / (kernel_ms * 1e-3) / 1e9) if traffic_per_launch and kernel_ms else None,
            "input_streaming_gbs_whole_step": step_stream_gbs,
            "note": "the kernel is integer-pipe bound: HBM sits at a few percent of its peak even counting the window-table spill traffic"}


def int32_peak():
    """Measured INT32 multiply-add issue rate (lane-ops / s) of this GPU: tools/microbench/int_pipe_bench."""
    exe = ROOT / "tools" / "microbench" / "int_pipe_bench"
    try:
        if not exe.exists():      # built by __graft_entry__.build(); rebuilt here if the binary did not travel
            subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-o", str(exe),
                            str(exe) + ".cu"], check=True, timeout=600)
        out = subprocess.run([str(exe), "8192"], check=True, stdout=subprocess.PIPE, text=True, timeout=120).stdout
        data = json.loads(out.strip().splitlines()[-1])
        return data
    except Exception as exc:      # the bench line then carries peak = None
        return {"error": repr(exc), "tests": {}}


_REAL_STDOUT = None


def emit(line):
    """The one JSON line of the contract goes to the process's original stdout."""
    data = (line + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line + "\n")
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    # Libraries print to stdout behind our back (NCCL's version banner when NCCL_DEBUG=VERSION, for one): route file
    # descriptor 1 to stderr for the whole run and keep the original stdout for the JSON line alone.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch

    import oracle as O
    from elastic_elgamal_b200 import Engine
    from elastic_elgamal_b200 import build as eg_build

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not eg_build.LIB.exists() and "EG_B200_LIB" not in os.environ:
        # the CUDA library normally travels with the repo snapshot; compile it (nvcc, sm_100a) if it did not.  Rank 0
        # of a node builds, the others wait for the file.  There is no other implementation to fall back to.
        if local_rank == 0:
            eg_build.build()
        else:
            for _ in range(1200):
                if eg_build.LIB.exists():
                    break
                time.sleep(0.5)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    e = Engine(device=local_rank)
    threads = O.hw_threads()
    sk, pk, cts, rings, sums = make_workload(args.unique, max(1, threads // max(1, min(world, 8))))
    e.set_receiver(pk)
    e.set_ring_mode(args.ring_mode)
    if args.chunk:
        e.set_chunk_items(args.chunk)
    ov, ot = O.verify_choice_batch(pk, OPTIONS, True, cts, rings, sums, threads=max(1, threads // max(1, min(world, 8))))

    B = args.ballots
    reps = (B + args.unique - 1) // args.unique
    h_cts = np.tile(cts, (reps, 1, 1))[:B]
    h_rings = np.tile(rings, (reps, 1, 1))[:B]
    h_sums = np.tile(sums, (reps, 1))[:B]
    expected_v = np.tile(ov, reps)[:B]

    dev = torch.device("cuda", local_rank)
    d_cts = torch.from_numpy(h_cts).to(dev)
    d_rings = torch.from_numpy(h_rings).to(dev)
    d_sums = torch.from_numpy(h_sums).to(dev)
    d_verdicts = torch.empty(B, dtype=torch.uint8, device=dev)
    d_tally = torch.empty((OPTIONS, 64), dtype=torch.uint8, device=dev)
    gathered = torch.empty((world, OPTIONS, 64), dtype=torch.uint8, device=dev) if world > 1 else None
    d_total = torch.empty((OPTIONS, 64), dtype=torch.uint8, device=dev) if world > 1 else None
    stream = torch.cuda.ExternalStream(e.stream, device=dev)

    def step_device():
        e.verify_choice_dev(B, OPTIONS, True, d_cts.data_ptr(), d_rings.data_ptr(), d_sums.data_ptr(),
                            d_verdicts.data_ptr(), d_tally.data_ptr())
        if world > 1:
            # the only exchange step: per-rank partial tallies (options x 64 B); NCCL has no EC-add reduction, so
            # all_gather + a local point-add kernel
            dist.all_gather_into_tensor(gathered, d_tally)
            e.ciphertexts_sum_dev(world, OPTIONS, gathered.data_ptr(), d_total.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = e.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the dominant kernel: k_ring in ring mode 2 (kind 1), k_commit in the per-equation A/B mode (kind 0)
    dom_kind = 1 if args.ring_mode == 2 else 0
    commit_ms, commit_tasks, commit_launches = 0.0, 0, 0
    other_ms, other_tasks, other_launches = 0.0, 0, 0
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step_device()
            st = e.last_kernel_stats(dom_kind)
            commit_ms += st["ms"]; commit_tasks += st["tasks"]; commit_launches += st["launches"]
            if dom_kind == 1:
                st = e.last_kernel_stats(0)
                other_ms += st["ms"]; other_tasks += st["tasks"]; other_launches += st["launches"]
        ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = e.kernel_launches - launches0
    sampler.stop_flag = True
    sampler.join()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())

    # ---------------- correctness of what was timed
    v = d_verdicts.cpu().numpy()
    assert (v == expected_v).all(), "verdict mismatch against the oracle"
    if world > 1:
        total = d_total.cpu().numpy()
        check, ok = e.ciphertexts_sum(gathered.cpu().numpy())
        assert ok and (check == total).all()
    else:
        total = d_tally.cpu().numpy()
    if rank == 0:
        import workloads as W
        accepted = (expected_v == 0)
        # tally of the timed batch must decrypt to the number of accepted ballots per option, over all ranks
        table_hi = world * B + 1
        table = O.DlogTable(0, table_hi) if table_hi <= (1 << 22) + 1 else None
        if table is not None:
            for k in range(OPTIONS):
                idx = np.arange(B)
                expect = world * int(np.count_nonzero(accepted & ((idx % args.unique) % OPTIONS == k)))
                got = table.get(O.decrypt_to_element(sk, bytes(total[k])))
                assert got == expect, (k, got, expect)

    # ---------------- end-to-end through the host C ABI (`e2e`): pinned host buffers, H2D + D2H inside
    p_cts = torch.from_numpy(h_cts).pin_memory()
    p_rings = torch.from_numpy(h_rings).pin_memory()
    p_sums = torch.from_numpy(h_sums).pin_memory()
    p_verdicts = torch.empty(B, dtype=torch.uint8).pin_memory()
    p_tally = torch.empty((OPTIONS, 64), dtype=torch.uint8).pin_memory()

    def step_e2e():
        st = e.lib.eg_verify_choice_batch(e.h, B, OPTIONS, 1, p_cts.data_ptr(), p_rings.data_ptr(), p_sums.data_ptr(),
                                          p_verdicts.data_ptr(), p_tally.data_ptr())
        e._check(st)
        if world > 1:
            dist.all_gather_into_tensor(gathered, p_tally.to(dev, non_blocking=True))

    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert (p_verdicts.numpy() == expected_v).all()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel + CPU baseline (rank 0)
    peak = int32_peak()
    imad = peak.get("tests", {}).get("imad", {})
    clk = sampler.summary()
    # peak = IMAD lane-ops/s measured with CUDA events by the microbenchmark, rescaled by (SM clock sampled during the
    # timed region) / (SM clock probed in the microbenchmark: clock64 ticks per globaltimer ns) when the two differ
    peak_ops = None
    if imad.get("per_s"):
        peak_ops = imad["per_s"]
        if clk.get("sm_mhz") and peak.get("sm_clock_mhz_probed"):
            peak_ops *= clk["sm_mhz"] / peak["sm_clock_mhz_probed"]
    achieved_ops = commit_tasks * FIELD_OPS_PER_COMMIT * IMAD_PER_FIELD_OP / (commit_ms * 1e-3) if commit_ms > 0 else None
    traffic = None
    tfile = ROOT / "profiles" / ("k_ring_traffic.json" if args.ring_mode == 2 else "k_commit_traffic.json")
    if tfile.exists():
        try:
            # dram bytes of one ncu-captured launch, rescaled to this run's equation sides per launch
            t = json.loads(tfile.read_text())
            traffic = t["dram_bytes_per_launch"] * (commit_tasks / max(1, commit_launches)) / t["equation_sides_per_launch"]
        except Exception:
            traffic = None
    roofline = {
        "kernel": "k_ring" if args.ring_mode == 2 else "k_commit", "bound": "int32",
        "achieved": achieved_ops / 1e12 if achieved_ops else None, "peak": peak_ops / 1e12 if peak_ops else None,
        "unit": "T int32 multiply-add lane-ops/s",
        "frac": (achieved_ops / peak_ops) if achieved_ops and peak_ops else None,
        "peak_source": "measured live: tools/microbench/int_pipe_bench `imad` lane-ops/s (CUDA events; = 63 of the nominal 64 "
                       "lanes/clk/SM) at the SM clock sampled during the timed region (MEASURED_PEAKS.json has no integer peak)",
        "traffic": traffic,
        "launches": commit_launches, "avg_launch_ms": commit_ms / max(1, commit_launches),
        "share_of_step": commit_ms / dev_ms if dev_ms else None,
        "ncu": "profiles/r1_k_ring_full_s7.txt: fmaheavy pipe 88.4 % busy, issue slots 45.6 %, 15.8 warps/SM, top stalls wait / math_pipe_throttle (ncu --set full, same command)",
        "second_kernel": {"kernel": "k_commit (sum proof)", "launches": other_launches, "equation_sides": other_tasks,
                          "ms": other_ms, "share_of_step": other_ms / dev_ms if dev_ms else None} if dom_kind == 1 else None,
        "algorithmic": {"field_ops_per_equation_side": FIELD_OPS_PER_COMMIT, "imad_per_field_op": IMAD_PER_FIELD_OP,
                        "equation_sides_per_launch": commit_tasks / max(1, commit_launches)},
        # `achieved` counts the reference's algorithm; the engine does less work per side (shared doublings,
        # chunked tables, inversion-only encoding), so the fraction of the pipe it really keeps busy is the lower one:
        "executed": ({"field_ops_per_equation_side": EXECUTED_FIELD_OPS_PER_RING_SIDE,
                      "frac": achieved_ops * EXECUTED_FIELD_OPS_PER_RING_SIDE / FIELD_OPS_PER_COMMIT / peak_ops}
                     if achieved_ops and peak_ops and dom_kind == 1 else None),
        "field_ops_per_s": commit_tasks * FIELD_OPS_PER_COMMIT / (commit_ms * 1e-3) if commit_ms > 0 else None,
        # the HBM view of the same kernel in the contract's own shape (north_star asks for achieved GB/s too): algorithmic
        # bytes = 2 points x 128 B + 2 encodings x 32 B + 2.2 scalars x 32 B in, 2 x 32 B out per ring of a 5-option ballot
        "hbm": hbm_roofline(commit_tasks, commit_ms, traffic, commit_launches,
                            world * B * (BALLOT_BYTES + 1) * args.steps / (dev_ms_max * 1e-3) / 1e9),
        "microbench": {k: v.get("per_clk_per_sm") for k, v in peak.get("tests", {}).items()},
    }
    cpu = None
    if not args.no_cpu_baseline:
        single, _, _, _ = cpu_baseline(pk, cts, rings, sums, 512, 1)
        sample = args.cpu_sample or max(512, int(single * threads * 6.0) // 64 * 64)
        rate, dt, _, _ = cpu_baseline(pk, cts, rings, sums, sample, threads)
        cpu = {"value": rate, "unit": "ballots/s", "cores": threads, "kind": "port",
               "sample": f"{sample} ballots of the same workload on {threads} threads ({dt:.1f} s); single thread {single:.1f} ballots/s",
               "single_thread": single}

    value = world * B * args.steps / (dev_ms_max * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "ballots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (GF(2^255-19), integers mod l)", "data": f"synthetic: {args.unique} seeded oracle-proved ballots (1% tampered) tiled to {B}/GPU",
        "config": {"workload": "EncryptedChoice::single 5 options: batch verify + homomorphic tally (BASELINE configs[1])",
                   "ballots_per_gpu": B, "global_ballots": world * B, "options": OPTIONS, "bytes_per_ballot": BALLOT_BYTES,
                   "l2": "inputs (736 B x ballots) and scratch exceed the 126 MB L2", "parallelism": f"dp{world} over ballots"},
        "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "ballots/s", "h2d_bytes_per_step": B * BALLOT_BYTES,
                "d2h_bytes_per_step": B + OPTIONS * 64, "steps": e2e_steps},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "reference_equivalent_field_ops_per_ballot": FIELD_OPS_PER_BALLOT,
    }
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
