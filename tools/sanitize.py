#!/usr/bin/env python
"""tools/sanitize.py - one small pass through every kernel of the library, for compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py

README-sized grid plus a few dozen cells with NaN / inf samples (so that k_thr_net hands cells to k_thr_seg and k_thr_cand hands
segments over), every threshold path (force modes 0..5), the metric sweep with the run filter on and off, the host pipeline with
small chunks, and the measure pre-pass.  Results are checked against the oracle so that the run is known to be a real one."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HDP_B200_HOST_CHUNK_CELLS", "32")
import torch
import oracle
from hdp_b200 import _core, _lib, _tables as tb

rng = np.random.default_rng(5)
base_ax = tb.TimeAxis.date_range("1961-01-01", "1990-12-31", "noleap")
run_ax = tb.TimeAxis.date_range("2015-01-01", "2022-12-31", "noleap")
C = 75
season = lambda ax: 15 + 10 * np.sin(2 * np.pi * (ax.dayofyr[:, None] - 110) / 365)
xb = (season(base_ax) + 3 * rng.standard_normal((len(base_ax), C))).astype(np.float32)
xr = (season(run_ax) + 2 + 3 * rng.standard_normal((len(run_ax), C))).astype(np.float32)
xb[rng.integers(0, len(base_ax), 8), rng.integers(0, C, 8)] = np.nan
xb[rng.integers(0, len(base_ax), 8), rng.integers(0, C, 8)] = np.inf
xb[:, 3] = 2.5
wt = tb.window_tables(base_ax.dayofyr, 7)
q = np.arange(0.9, 1.0, 0.01)
st, dm = tb.hemisphere_ranges(run_ax), tb.doy_map(run_ax.dayofyr)
defs = [[3, 0, 0], [3, 1, 1], [4, 0, 0], [4, 1, 1], [5, 0, 0], [5, 1, 1]]
south = (rng.random(C) < 0.5).astype(np.uint8)
want_thr = oracle.thresholds_batch(xb, wt.window_samples(), q)
L = _lib.lib()
d_xb, d_xr = torch.as_tensor(xb).cuda(), torch.as_tensor(xr).cuda()
eq = lambda a, b: np.array_equal((a + 0.0).view(np.uint64), (b + 0.0).view(np.uint64))
for mode in (0, 5, 4, 3, 2, 1):
    L.hdp_b200_thresholds_force_generic(mode)
    thr = _core.thresholds_array(d_xb, wt, q)
    assert eq(thr.cpu().numpy(), want_thr), mode
L.hdp_b200_thresholds_force_generic(0)
q2 = np.array([0.0, 0.5, 0.99, 1.0])                                   # low quantiles: the segment kernel without candidate filter
assert eq(_core.thresholds_array(d_xb, wt, q2).cpu().numpy(), oracle.thresholds_batch(xb, wt.window_samples(), q2))
thr_clean = np.nan_to_num(want_thr, nan=1e9)
d_thr = torch.as_tensor(thr_clean).cuda()
want_met = oracle.metrics_batch(xr, thr_clean, dm, defs, st.north, st.south, south)
for filt in (1, 0):
    L.hdp_b200_metrics_run_filter(filt)
    out = _core.metrics_array(d_xr, d_thr, dm, defs, st.north, st.south, south)
    assert np.array_equal(out.cpu().numpy().astype(np.int64).transpose(1, 2, 4, 0, 3), want_met), filt
L.hdp_b200_metrics_run_filter(1)
_core.hot_days_array(d_xr, d_thr, dm)
thr_h = _core.thresholds_host(xb, wt, q, keep=True)
assert eq(thr_h, want_thr)
out_h = _core.metrics_host(xr, thr_clean, dm, defs, st.north, st.south, south)
assert np.array_equal(out_h.astype(np.int64).transpose(1, 2, 4, 0, 3), want_met)
t = torch.as_tensor((xr[:2000] * 1.8 + 32).astype(np.float32)).cuda()
rh = torch.as_tensor(rng.uniform(5, 95, t.shape).astype(np.float32)).cuda()
_core.heat_index_array(t, rh); _core.heat_index_measure_array(torch.as_tensor(xr[:2000]).cuda(), rh); _core.to_celsius_array(t, "degF")
torch.cuda.synchronize()
_core.host_release()
print("sanitize pass ok: launches", _core.launch_count())
