"""Between labelled N-d arrays (time + arbitrary cell dims) and the kernels' [T, C] view."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def to_time_cells(values: np.ndarray, dims: Tuple[str, ...], require_float32: bool = False) -> Tuple[np.ndarray, List[str], List[int]]:
    """float32 [T, C] VIEW (or copy when no view exists) of an N-d array with a ``time`` dim; cells are the remaining
    dims flattened in their original order.  Time-first arrays become cell-contiguous views, time-last arrays become
    time-contiguous (ld_t == 1) views - both are handled natively by the C ABI, so no host-side transposition."""
    if "time" not in dims:
        raise ValueError("measure has no 'time' dimension")
    axis = dims.index("time")
    cell_dims = [d for d in dims if d != "time"]
    cell_shape = [values.shape[i] for i, d in enumerate(dims) if d != "time"]
    v = np.asarray(values)
    if v.dtype != np.float32:
        # The kernels read float32 samples.  Thresholds: the reference casts too (hdp/threshold.py:121).  Metrics: the
        # reference compares whatever dtype it is given in float64 (hdp/metric.py:280-301), so a float64 measure that
        # did not go through format_standard_measures (which casts to float32, hdp/measure.py:166) can differ from the
        # reference on hot days within one float32 ulp of the threshold - callers are told once.
        if require_float32 and v.dtype == np.float64:
            import warnings
            warnings.warn("hdp_b200 compares float32 samples with the thresholds: this float64 measure is rounded to float32 "
                          "first (pass it through format_standard_measures, like the reference workflow does)", stacklevel=3)
        v = v.astype(np.float32)
    T = v.shape[axis]
    C = int(np.prod(cell_shape)) if cell_shape else 1
    if axis == 0:
        out = np.ascontiguousarray(v).reshape(T, C)
    elif axis == v.ndim - 1:
        out = np.ascontiguousarray(v).reshape(C, T).T
    else:
        out = np.ascontiguousarray(np.moveaxis(v, axis, 0)).reshape(T, C)
    return out, cell_dims, cell_shape


def cell_latitudes(lat_values: np.ndarray, cell_dims: List[str], cell_shape: List[int]) -> np.ndarray:
    """Latitude of every flattened cell (lat broadcast over the other cell dims)."""
    if "lat" not in cell_dims:
        raise ValueError("measure has no 'lat' dimension")
    shape = [1] * len(cell_dims)
    i = cell_dims.index("lat")
    shape[i] = cell_shape[i]
    return np.broadcast_to(np.asarray(lat_values, dtype=np.float64).reshape(shape), cell_shape).reshape(-1)
