#!/usr/bin/env python3
"""Aggregates an ncu source page (SASS view) into address ranges split at RET instructions (= device functions), with
sample counts, executed instructions and the dominant stall reasons per function.
usage: python tools/ncu_hotspots.py report.ncu-rep [kernel-name]"""
import csv, subprocess, sys
rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if len(sys.argv) > 2: cmd += ["-k", sys.argv[2]]
out = subprocess.run(cmd, stdout=subprocess.PIPE, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
funcs, cur = [], []
for r in data:
    cur.append(r)
    toks = r[col["Source"]].split()
    if any(t.startswith("RET") for t in toks) or (toks and toks[0].startswith("EXIT")):
        funcs.append(cur); cur = []
if cur: funcs.append(cur)
tot = sum(int(r[col["# Samples"]] or 0) for r in data) or 1
toti = sum(int(r[col["Instructions Executed"]] or 0) for r in data) or 1
print("%-4s %-8s %7s %8s %8s  %s" % ("fn", "instrs", "size", "samples%", "inst%", "top stalls / mix"))
for k, f in enumerate(funcs):
    s = sum(int(r[col["# Samples"]] or 0) for r in f)
    ie = sum(int(r[col["Instructions Executed"]] or 0) for r in f)
    st = sorted(((sum(int(r[col[h]] or 0) for r in f), h) for h in stalls), reverse=True)[:4]
    ops = {}
    for r in f:
        op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
        if op.startswith("@"): op = r[col["Source"]].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[col["Instructions Executed"]] or 0)
    top_ops = sorted(ops.items(), key=lambda x: -x[1])[:5]
    print("%-4d %-8d %7d %7.1f%% %7.1f%%  %s | %s" % (k, len(f), len(f) * 16, 100.0 * s / tot, 100.0 * ie / toti,
          " ".join("%s=%.0f%%" % (h.replace("stall_", ""), 100.0 * v / max(1, s)) for v, h in st),
          " ".join("%s:%.0f%%" % (o, 100.0 * c / max(1, ie)) for o, c in top_ops)))
