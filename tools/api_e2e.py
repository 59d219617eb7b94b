"""Wall time of the drop-in Python API (hdp_b200.threshold.compute_thresholds + hdp_b200.metric.compute_group_metrics)
on one measure of cmip6_1deg, labelled arrays in HDP's (lon, lat, time) order, host memory in / host memory out.
    python tools/api_e2e.py [n_lat n_lon]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hdp_b200 import _tables as tb, measure, metric, threshold, workloads, xr, synth

wl = workloads.get("cmip6_1deg")
n_lat, n_lon = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (wl.n_lat, wl.n_lon)
lat = -90 + (np.arange(n_lat) + 0.5) * (180.0 / n_lat)
lon = (np.arange(n_lon) + 0.5) * (360.0 / n_lon)
cell_lat = np.repeat(lat, n_lon)


def field(axis, seed, trend):
    x = synth.gridded_field(cell_lat, axis.dayofyr, seed=seed, trend=trend, device="cuda").cpu().numpy()     # [T, C]
    v = np.ascontiguousarray(x.T).reshape(n_lat, n_lon, len(axis)).transpose(1, 0, 2)                         # (lon, lat, time) view
    return xr.DataArray(np.ascontiguousarray(v), dims=["lon", "lat", "time"], coords={"lon": lon, "lat": lat, "time": axis},
                        name="tas", attrs={"units": "degC"})


t = time.perf_counter()
base, run = field(wl.base_axis(), 1, 0.0), field(wl.run_axis(), 2, 4.0)
print(f"synthetic inputs: {time.perf_counter() - t:.1f} s", flush=True)
for rep in range(2):
    t0 = time.perf_counter()
    base_m = measure.format_standard_measures([base])
    run_m = measure.format_standard_measures([run])
    t1 = time.perf_counter()
    thr = threshold.compute_thresholds(base_m, wl.percentiles)
    t2 = time.perf_counter()
    met = metric.compute_group_metrics(run_m, thr, wl.defs)
    t3 = time.perf_counter()
    cy = n_lat * n_lon * (wl.base_years + wl.run_years)
    print(f"rep {rep}: format_standard_measures {t1 - t0:.2f} s, compute_thresholds {t2 - t1:.2f} s, compute_group_metrics {t3 - t2:.2f} s "
          f"-> {cy / (t3 - t0):.3g} cell-years/s through the Python API", flush=True)
name = "tas.tas_threshold.HWF"
print(name, met[name].shape, met[name].dtype, float(np.asarray(xr.values_of(met[name])).mean()))
# the same with the library's native metric encoding (no int64 widening)
metric.METRIC_DTYPE = np.uint16
t2 = time.perf_counter()
met = metric.compute_group_metrics(run_m, thr, wl.defs)
print(f"METRIC_DTYPE = uint16: compute_group_metrics {time.perf_counter() - t2:.2f} s", met[name].dtype, flush=True)
