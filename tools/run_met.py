#!/usr/bin/env python
"""tools/run_met.py [cells] [reps] [workload] - the metric sweep only (thresholds computed once): per-kernel CUDA-event times."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdp_b200 import _core, synth, workloads, _tables as tb

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = workloads.get(sys.argv[3] if len(sys.argv) > 3 else "cmip6_1deg")
lat, _ = synth.grid_latitudes(wl.n_lat, wl.n_lon)
if cells:
    lat = lat[np.linspace(0, wl.cells - 1, cells).astype(np.int64)]
base = synth.gridded_field(lat, wl.base_axis().dayofyr, seed=1234, offset=5.0, device="cuda")
run = synth.gridded_field(lat, wl.run_axis().dayofyr, seed=1235, offset=5.0, trend=4.0, device="cuda")
wt, st, dm = wl.window_tables(), wl.seasons(), tb.doy_map(wl.run_axis().dayofyr)
south = torch.as_tensor((lat < 0).astype(np.uint8), device="cuda")
thr = _core.thresholds_array(base, wt, wl.percentiles)
del base
out = _core.metrics_array(run, thr, dm, wl.defs, st.north, st.south, south)
torch.cuda.synchronize()
_core.timing_enable(True); _core.timing_read()
for _ in range(reps):
    _core.metrics_array(run, thr, dm, wl.defs, st.north, st.south, south, out=out)
torch.cuda.synchronize()
by = {}
for n, ms in _core.timing_read():
    by.setdefault(n, []).append(ms)
print({k: round(float(np.mean(v)), 4) for k, v in by.items()}, "cells", lat.size, "checksum", int(out.to(torch.int64).sum()))
