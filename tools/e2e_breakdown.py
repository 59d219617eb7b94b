"""Wall time of the two host-buffer entry points on one measure of cmip6_1deg (pinned host buffers).
    python tools/e2e_breakdown.py [cells] [--pageable]      (--pageable: ordinary NumPy memory instead of pinned buffers)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hdp_b200 import _core, _tables as tb, synth, workloads

wl = workloads.get("cmip6_1deg")
wt, st = wl.window_tables(), wl.seasons()
dm = tb.doy_map(wl.run_axis().dayofyr)
lat, _ = synth.grid_latitudes(wl.n_lat, wl.n_lon)
PAGEABLE = "--pageable" in sys.argv
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
if argv:
    lat = lat[np.linspace(0, wl.cells - 1, int(argv[0])).astype(np.int64)]
C = lat.size
base = synth.gridded_field(lat, wl.base_axis().dayofyr, seed=1, device="cuda")
run = synth.gridded_field(lat, wl.run_axis().dayofyr, seed=2, trend=4.0, device="cuda")
P, D, Y = len(wl.percentiles), len(wl.defs), st.n_years
pin = not PAGEABLE
h_base = torch.empty(base.shape, dtype=torch.float32, pin_memory=pin); h_base.copy_(base)
h_run = torch.empty(run.shape, dtype=torch.float32, pin_memory=pin); h_run.copy_(run)
h_thr = torch.empty((C, wt.n_doy, P), dtype=torch.float64, pin_memory=pin)
h_out = torch.empty((4, P, D, Y, C), dtype=torch.uint16, pin_memory=pin)
print("host buffers:", "pageable" if PAGEABLE else "pinned", flush=True)
del base, run
torch.cuda.empty_cache()
south = (lat < 0).astype(np.uint8)
for rep in range(3):
    t0 = time.perf_counter()
    _core.thresholds_host(h_base.numpy(), wt, wl.percentiles, out=h_thr.numpy())
    t1 = time.perf_counter()
    _core.metrics_host(h_run.numpy(), h_thr.numpy(), dm, wl.defs, st.north, st.south, south, out=h_out.numpy())
    t2 = time.perf_counter()
    b1 = h_base.numel() * 4 + h_thr.numel() * 8
    b2 = h_run.numel() * 4 + h_thr.numel() * 8 + h_out.numel() * 2
    print(f"rep {rep}: thresholds_host {1e3 * (t1 - t0):.1f} ms ({b1 / (t1 - t0) / 1e9:.1f} GB/s moved), "
          f"metrics_host {1e3 * (t2 - t1):.1f} ms ({b2 / (t2 - t1) / 1e9:.1f} GB/s moved)", flush=True)
