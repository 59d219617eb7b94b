"""Path-in / path-out wrappers (reference ``compute_threshold_io`` hdp/threshold.py:232-289, ``compute_metrics_io``
hdp/metric.py:526-590) as an out-of-core stream through the GPU.

The reference reads a netCDF / zarr store into a Dask-backed DataArray and writes the result back; its wrappers are broken as
shipped (``Path.isdir``, undefined ``makedirs`` / ``overwrite``) and neither xarray, zarr nor netCDF4 can be installed in this
image.  What is kept here is the data flow - measures that never fit in host memory at once go from disk to the device and
results back to disk - on the one container that needs no library: ``.npy`` files opened with ``numpy.memmap``.  A memory-mapped
array is pageable host memory, so the chunked three-stream pipeline of ``hdp_b200_*_host`` (pinned staging rings filled by a
thread pool while the copy engines move the previous blocks) reads it cell chunk by cell chunk straight from the page cache and
writes the outputs the same way: at no time is more than a few chunks of the measure resident in host RAM.  With xarray present,
``.nc`` / ``.zarr`` paths are opened with it instead and handed to the Dataset-level functions (same semantics as the reference).
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np

from . import _core, _tables, xr


def _check_output(path: str, overwrite: bool) -> None:
    if os.path.exists(path) and not overwrite:
        raise FileExistsError(f"Overwrite parameter set to False and file exists at '{path}'.")        # threshold.py:267-268
    parent = os.path.dirname(os.path.abspath(path))
    if not os.path.isdir(parent):
        if overwrite:
            os.makedirs(parent, exist_ok=True)
        else:
            raise FileExistsError(f"Overwrite parameter set to False and directory '{parent}' does not exist.")   # :270-274


def _open_measure(path: str) -> np.ndarray:
    a = np.load(path, mmap_mode="r")
    if a.ndim != 2 or a.dtype != np.float32:
        raise ValueError(f"'{path}' must hold a float32 [time, cells] array")
    return a


def compute_threshold_io(baseline_path: str, time_axis: _tables.TimeAxis, output_path: str, percentiles: Sequence[float],
                         rolling_window_size: int = 7, overwrite: bool = False, units: str = None) -> None:
    """``baseline_path``: ``.npy`` float32 ``[time, cells]`` (any size: memory-mapped).  Writes float64 ``[cells, doy, percentile]``
    to ``output_path`` (``.npy``).  ``time_axis`` carries the calendar fields of the time dimension (a file of plain numbers has
    no cftime coordinate)."""
    if not str(output_path).endswith(".npy") or not str(baseline_path).endswith(".npy"):
        if xr.HAVE_XARRAY:                                                   # pragma: no cover - no xarray in the build image
            import xarray
            from . import threshold as _thr
            src = xarray.open_zarr(baseline_path) if str(baseline_path).endswith(".zarr") else xarray.open_dataset(baseline_path)
            ds = _thr.compute_thresholds(src, percentiles, rolling_window_size=rolling_window_size)
            _check_output(output_path, overwrite)
            ds.to_zarr(output_path) if str(output_path).endswith(".zarr") else ds.to_netcdf(output_path)
            return
        raise ValueError(f"File type of '{baseline_path}' / '{output_path}' not supported without xarray: use .npy")   # :276-277
    _check_output(output_path, overwrite)
    x = _open_measure(baseline_path)
    if x.shape[0] != len(time_axis):
        raise ValueError("time_axis does not match the file's time dimension")
    tables = _tables.window_tables(time_axis.dayofyr, rolling_window_size)
    q = np.asarray(percentiles, dtype=np.float64)
    out = np.lib.format.open_memmap(output_path, mode="w+", dtype=np.float64, shape=(x.shape[1], tables.n_doy, q.size))
    _core.thresholds_host(x, tables, q, out=out, units=units)
    out.flush()


def compute_metrics_io(measure_path: str, time_axis: _tables.TimeAxis, threshold_path: str, cell_lat: np.ndarray, output_path: str,
                       hw_definitions: Sequence[Sequence[int]], overwrite: bool = False, units: str = None) -> None:
    """``measure_path`` float32 ``[time, cells]`` and ``threshold_path`` float64 ``[cells, doy, percentile]`` (both ``.npy``,
    memory-mapped) -> uint16 ``[4, percentile, definition, year, cells]`` at ``output_path`` (planes HWF, HWN, HWD, HWA)."""
    _check_output(output_path, overwrite)
    x = _open_measure(measure_path)
    thr = np.load(threshold_path, mmap_mode="r")
    if x.shape[0] != len(time_axis) or thr.shape[0] != x.shape[1]:
        raise ValueError("measure, thresholds and time_axis do not match")
    st = _tables.hemisphere_ranges(time_axis)
    defs = np.asarray(hw_definitions, dtype=np.int64).reshape(-1, 3)
    out = np.lib.format.open_memmap(output_path, mode="w+", dtype=np.uint16, shape=(4, thr.shape[2], defs.shape[0], st.n_years, x.shape[1]))
    _core.metrics_host(x, thr, _tables.doy_map(time_axis.dayofyr), defs, st.north, st.south, _tables.is_south(np.asarray(cell_lat)),
                       out=out, units=units)
    out.flush()
