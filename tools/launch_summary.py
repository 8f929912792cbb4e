#!/usr/bin/env python3
"""Summarises an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`)
into per-kernel launch counts, total time and share of the step.
usage: python tools/launch_summary.py launches.csv [header comment]"""
import csv, collections, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).strip()
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
    tot[name] += v; cnt[name] += 1
micro = lambda n: any(x in n for x in ("k_pipe", "k_field", "k_clock_probe"))
step = sum(v for n, v in tot.items() if not micro(n))
if len(sys.argv) > 2: print("#", sys.argv[2])
print("# unit=ns; shares exclude the integer-pipe microbenchmark that bench.py runs for the roofline denominator (listed last)")
for n, v in tot.most_common():
    if not micro(n): print("%-34s launches=%3d total=%15.1f share=%5.1f%%" % (n, cnt[n], v, 100 * v / step))
for n, v in tot.most_common():
    if micro(n): print("%-34s launches=%3d total=%15.1f (microbenchmark)" % (n, cnt[n], v))
