#!/usr/bin/env python
"""Per-source-line summary of an ncu report:
    python profiles/srcprof.py report.ncu-rep [top_n] [--kernel SUBSTR] [line_lo line_hi ...]
(uses `ncu --page source --print-source cuda,sass --csv`; sums samples / instructions / shared wavefronts per line)."""
import csv, subprocess, sys, io
args = sys.argv[1:]
kern = None
if "--kernel" in args:
    i = args.index("--kernel"); kern = args[i + 1]; del args[i:i + 2]
rep = args[0]; top = int(args[1]) if len(args) > 1 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
if kern:
    cmd += ["-k", "regex:" + kern]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
h = heads[0]
end = heads[1] if len(heads) > 1 else len(rows)
ix = {n: i for i, n in enumerate(rows[h])}
def num(x):
    try: return float(x)
    except ValueError: return 0.0
L = []
for r in rows[h + 1:end]:
    if len(r) < 10 or r[2] != "-": continue
    L.append((int(r[0]), r[1][:110], num(r[ix["# Samples"]]), num(r[ix["Instructions Executed"]]), num(r[ix["L1 Wavefronts Shared"]]), num(r[ix["L1 Wavefronts Shared Ideal"]])))
ts, ti, tw = (sum(l[k] for l in L) or 1 for k in (2, 3, 4))
print(f"total samples {ts:.0f} inst {ti:.3e} shared wavefronts {tw:.3e}")
if len(args) > 3:
    bounds = [int(v) for v in args[2:]]
    for a, b in zip(bounds, bounds[1:]):
        sel = [l for l in L if a <= l[0] < b]
        print(f"lines {a:4d}-{b - 1:4d}: samples {sum(l[2] for l in sel) / ts * 100:5.1f}%  inst {sum(l[3] for l in sel) / ti * 100:5.1f}%  wf {sum(l[4] for l in sel) / tw * 100:5.1f}% (ideal {sum(l[5] for l in sel) / tw * 100:5.1f}%)")
for l in sorted(L, key=lambda l: -l[2])[:top]:
    print(f"{l[0]:4d} {l[2] / ts * 100:5.1f}% inst {l[3] / ti * 100:5.1f}% wf {l[4] / tw * 100:5.1f}% ideal {l[5] / tw * 100:5.1f}% | {l[1]}")
