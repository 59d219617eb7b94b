#!/usr/bin/env python
"""Per-phase summary of one kernel in an ncu report (needs --import-source on and -lineinfo):

    python profiles/phaseprof.py report.ncu-rep FILE_SUBSTR MARKER_REGEX [first_line last_line]

Source lines of the file whose name contains FILE_SUBSTR are grouped into phases: a phase starts at every line matching
MARKER_REGEX (e.g. '// ---- [0-9A-Z]') inside [first_line, last_line].  Prints, per phase, the share of warp-stall samples,
executed warp instructions and shared-memory wavefronts (with the ideal, i.e. conflict-free, count)."""
import csv, io, os, re, subprocess, sys
rep, fsub, marker = sys.argv[1], sys.argv[2], re.compile(sys.argv[3])
lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 10 ** 9)
if rep.endswith(".csv"):                                           # an exported `--page source --print-source cuda,sass --csv`
    txt = open(rep).read()
else:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
def num(x):
    try: return float(x)
    except ValueError: return 0.0
sections, cur = [], None
for r in rows:
    if r and r[0] == "File Path" and len(r) > 1:
        cur = {"file": r[1], "rows": [], "ix": None}; sections.append(cur)
    elif r and r[0] == "Line No" and cur is not None:
        cur["ix"] = {}
        for i, n in enumerate(r):
            cur["ix"].setdefault(n, i)
    elif cur is not None and cur["ix"] and len(r) > 5 and r[2] == "-":     # per-source-line totals (SASS rows carry an address)
        cur["rows"].append(r)
sec = [s for s in sections if fsub in s["file"] and s["rows"]]
if not sec:
    sys.exit("no such file in the report: " + ", ".join(s["file"] for s in sections))
sec = max(sec, key=lambda s: len(s["rows"]))
ix = sec["ix"]
# the report lists code lines only: marker (comment) lines are located in the source file itself
marks = []
src_path = sec["file"]
if not os.path.exists(src_path) and "/hdp_b200/" in src_path:      # captured on the GPU box: same tree, other root
    src_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hdp_b200", src_path.split("/hdp_b200/", 1)[1])
try:
    with open(src_path) as f:
        marks = [(i + 1, l.strip()[:100]) for i, l in enumerate(f) if marker.search(l)]
except OSError:
    pass
acc = {}
order = []
for r in sec["rows"]:
    line = int(num(r[ix["Line No"]]))
    if not (lo <= line <= hi):
        continue
    name = "(before the first marker)"
    for ml, mt in marks:
        if ml <= line:
            name = f"{ml}: {mt}"
    if name not in acc:
        acc[name] = [0.0, 0.0, 0.0, 0.0]; order.append(name)
    a = acc[name]
    a[0] += num(r[ix["# Samples"]]); a[1] += num(r[ix["Instructions Executed"]])
    a[2] += num(r[ix["L1 Wavefronts Shared"]]); a[3] += num(r[ix["L1 Wavefronts Shared Ideal"]])
tot = [sum(acc[n][k] for n in order) or 1.0 for k in range(4)]
print(f"{sec['file']} lines {lo}..{hi}: samples {tot[0]:.0f}, warp instructions {tot[1]:.3e}, shared wavefronts {tot[2]:.3e} (ideal {tot[3]:.3e})")
print(f"{'samples':>8s} {'inst':>7s} {'wavefr.':>8s} {'excess':>7s}  phase")
for n in order:
    a = acc[n]
    print(f"{a[0] / tot[0]:8.1%} {a[1] / tot[1]:7.1%} {a[2] / tot[2]:8.1%} {(a[2] - a[3]) / tot[2]:7.1%}  {n}")
