#!/usr/bin/env python
"""Summarise one profiling round (tools/profile_round.sh TAG) into profiles/:
    python tools/ncu_summary.py TAG
writes profiles/TAG_launches.txt (per-kernel share of the step from the ncu launch list), profiles/TAG_ncu_summary.md
(key `--set full` metrics per kernel) and profiles/traffic.json (DRAM bytes per launch, read by bench.py for `roofline.traffic`)."""
import csv, io, json, os, subprocess, sys, collections
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "gpurun_out")
prof = os.path.join(root, "profiles")

# ---- launch list
rows = []
with open(os.path.join(out, f"launches_{tag}.csv")) as f:
    lines = [l for l in f if l.startswith('"')]
rd = list(csv.reader(io.StringIO("".join(lines))))
h = rd[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.OrderedDict()
for r in rd[1:]:
    name = r[ki].split("(")[0].replace("void ", "").strip()
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += v
allms = sum(t[1] for t in tot.values())
with open(os.path.join(prof, f"{tag}_launches.txt"), "w") as f:
    f.write(f"# ncu launch list of `python bench.py --no-e2e --no-cpu --no-extra --steps 1 --warmup 1` (warm-up step + timed step, all launches),\n"
            f"# --metrics gpu__time_duration.sum --clock-control none; per-launch times are serialised and cold-cache: compare SHARES\n")
    f.write(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'ms/launch':>10s} {'share':>7s}\n")
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:60]:60s} {n:8d} {ms:10.3f} {ms / n:10.3f} {ms / allms:7.1%}\n")
print(open(os.path.join(prof, f"{tag}_launches.txt")).read())

# ---- full captures
want = [("gpu__time_duration.sum", "duration"), ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "registers / thread"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "FMA pipe %")]
unit_scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic = {}
md = [f"# {tag}: `ncu --set full --clock-control none --import-source on`, first launch of each kernel, full cmip6_1deg grid (64 800 cells)\n",
      "Reports: gpurun_out/prof_%s_{net,hot,scan}.ncu-rep (not tracked); numbers copied from `ncu -i ... --page raw --csv`.\n" % tag]
for short in ("net", "cand", "seg", "hot", "scan"):
    rep = os.path.join(out, f"prof_{tag}_{short}.ncu-rep")
    raw_csv = os.path.join(out, f"raw_{tag}_{short}.csv")          # written on the GPU box by profile_round.sh (the report itself is dropped there)
    if os.path.exists(raw_csv):
        txt = open(raw_csv).read()
    elif os.path.exists(rep):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        continue
    rr = list(csv.reader(io.StringIO(txt)))
    hh, units, vals = rr[0], rr[1], rr[2]
    name = vals[hh.index("Kernel Name")].split("(")[0].replace("void ", "")
    md.append(f"\n## {name}  (grid {vals[hh.index('Grid Size')]}, block {vals[hh.index('Block Size')]})\n")
    md.append("| metric | value |\n|---|---|\n")
    rd_b = wr_b = 0.0
    for key, label in want:
        if key not in hh:
            continue
        i = hh.index(key)
        md.append(f"| {label} (`{key}`) | {vals[i]} {units[i]} |\n")
        if key == "dram__bytes_read.sum": rd_b = float(vals[i].replace(",", "")) * unit_scale.get(units[i], 1.0)
        if key == "dram__bytes_write.sum": wr_b = float(vals[i].replace(",", "")) * unit_scale.get(units[i], 1.0)
    stall = [(float(vals[i] or 0), n) for i, n in enumerate(hh) if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio")]
    top = sorted(stall, reverse=True)[:5]
    md.append("| top stalls (warps per issue-active cycle) | " + ", ".join(f"{n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for v, n in top) + " |\n")
    key_name = name.split("<")[0].strip()
    if key_name.startswith("hdp::"):
        key_name = key_name[5:]
    if key_name.startswith("k_thr_net"):           # k_thr_net_tm (tensor-memory variant) is timed under the same kernel id
        key_name = "k_thr_net"
    traffic[key_name] = {"dram_bytes_read": rd_b, "dram_bytes_write": wr_b, "dram_bytes": rd_b + wr_b, "cells": 64800, "capture": f"prof_{tag}_{short}.ncu-rep"}
open(os.path.join(prof, f"{tag}_ncu_summary.md"), "w").write("".join(md))
json.dump({"tag": tag, "kernels": traffic}, open(os.path.join(prof, "traffic.json"), "w"), indent=1)
print("".join(md))
