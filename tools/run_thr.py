#!/usr/bin/env python
"""tools/run_thr.py [cells] [reps] [workload] - thresholds only on the workload's grid (quick kernel iterations / ncu captures).
Prints per-kernel CUDA-event times."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdp_b200 import _core, synth, workloads

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = workloads.get(sys.argv[3] if len(sys.argv) > 3 else "cmip6_1deg")
lat, _ = synth.grid_latitudes(wl.n_lat, wl.n_lon)
if cells:
    lat = lat[np.linspace(0, wl.cells - 1, cells).astype(np.int64)]
base = synth.gridded_field(lat, wl.base_axis().dayofyr, seed=1234, offset=5.0, device="cuda")
wt = wl.window_tables()
out = torch.empty((lat.size, wt.n_doy, len(wl.percentiles)), dtype=torch.float64, device="cuda")
_core.thresholds_array(base, wt, wl.percentiles, out=out)
torch.cuda.synchronize()
_core.timing_enable(True); _core.timing_read()
for _ in range(reps):
    _core.thresholds_array(base, wt, wl.percentiles, out=out)
torch.cuda.synchronize()
by = {}
for n, ms in _core.timing_read():
    by.setdefault(n, []).append(ms)
print({k: round(float(np.mean(v)), 4) for k, v in by.items()}, "cells", lat.size, "checksum", float(out.sum()))
