"""hdp.threshold on B200 (reference hdp/threshold.py).

Same function names, arguments, output variables, coordinates and attributes as the reference; the per-cell Numba
gufunc + Dask dispatch underneath (`compute_percentiles`, threshold.py:52-93; `xarray.map_blocks`, :161) is replaced
by one call into libhdp_b200.so.  Results are eager (NumPy-backed): ``.compute()`` on them is a no-op.
"""
from __future__ import annotations

import numpy as np

from . import _core, _layout, _tables, xr
from .utils import add_history, get_version, time_axis_of


def datetimes_to_windows(datetimes, window_radius: int) -> np.ndarray:
    """The reference's ``int[n_doy, (2r+1) * n_y]`` window table (hdp/threshold.py:12-49), quirks included."""
    axis = datetimes if isinstance(datetimes, _tables.TimeAxis) else _tables.TimeAxis.from_datetimes(list(datetimes))
    return _tables.window_tables(axis.dayofyr, window_radius).window_samples()


def compute_percentiles(temperatures: np.ndarray, window_samples: np.ndarray, percentiles: np.ndarray) -> np.ndarray:
    """Array-level seam of the reference gufunc '(t),(d,b),(p)->(d,p)' (hdp/threshold.py:52-78) on the GPU:
    ``temperatures`` float32 ``[..., t]`` -> float64 ``[..., d, p]``.  ``window_samples`` is factored back into the
    (time_index, win_rows) pair the kernel takes; it must come from :func:`datetimes_to_windows` (row-structured)."""
    temps = np.asarray(temperatures, dtype=np.float32)
    lead = temps.shape[:-1]
    win = np.asarray(window_samples, dtype=np.int64)
    tables = _tables.factor_window_samples(win)
    x = temps.reshape(-1, temps.shape[-1]).T                       # [T, C] time-contiguous view
    out = _core.thresholds_host(x, tables, np.asarray(percentiles, dtype=np.float64))
    return out.reshape(*lead, win.shape[0], -1)


def compute_percentiles_wrapper(baseline_data, rolling_windows, percentiles):
    """hdp/threshold.py:81-93: the gufunc applied over a labelled array - core dims ``time`` / ``(doy, t_index)`` / ``percentile`` ->
    ``(<other dims>, doy, percentile)``, attributes of ``baseline_data`` kept (``keep_attrs="override"``)."""
    dims = tuple(baseline_data.dims)
    values = np.moveaxis(xr.values_of(baseline_data), dims.index("time"), -1)
    q = np.asarray(getattr(percentiles, "values", percentiles), dtype=np.float64)
    out = compute_percentiles(values, np.asarray(getattr(rolling_windows, "values", rolling_windows)), q)
    coords = dict(xr.coords_of(baseline_data, skip=("time",)))
    if hasattr(percentiles, "coords") and "percentile" in percentiles.coords:
        coords["percentile"] = xr.coord_values(percentiles, "percentile")
    return xr.DataArray(out, dims=[*(d for d in dims if d != "time"), "doy", "percentile"], coords=coords,
                        name=baseline_data.name, attrs=dict(baseline_data.attrs))


def compute_threshold(baseline_data, percentiles, no_season: bool = False, rolling_window_size: int = 7, fixed_value: float = None):
    """hdp/threshold.py:96-204.  ``no_season`` and ``fixed_value`` are accepted and echoed into the attributes only,
    exactly like the reference (they are not implemented there either, :105-110, :182-184)."""
    values = xr.values_of(baseline_data)
    dims = tuple(baseline_data.dims)
    time_axis = time_axis_of(baseline_data)
    coords = xr.coords_of(baseline_data, skip=("time", "member") if "member" in dims else ("time",))
    if "member" in dims:                                            # pool the ensemble into the sample, :114-119
        m_axis = dims.index("member")
        t_axis = dims.index("time")
        moved = np.moveaxis(values, (m_axis, t_axis), (0, 1))       # [member, time, ...]
        values = moved.reshape((moved.shape[0] * moved.shape[1],) + moved.shape[2:])
        dims = ("time",) + tuple(d for d in dims if d not in ("member", "time"))
        n_member = moved.shape[0]
        time_axis = _tables.TimeAxis(*(np.tile(getattr(time_axis, f), n_member) for f in ("year", "month", "day", "dayofyr")),
                                     calendar=time_axis.calendar)

    percentiles = np.array(percentiles)
    x, cell_dims, cell_shape = _layout.to_time_cells(values, dims)  # .astype(float32), :121
    tables = _tables.window_tables(time_axis.dayofyr, rolling_window_size)   # datetimes_to_windows, :125
    thr = _core.thresholds_host(x, tables, percentiles.astype(np.float64), keep=True)   # [C, n_doy, P] float64 (+ device copy)
    n_doy = tables.n_doy

    da_coords = dict(coords)
    da_coords["doy"] = np.arange(n_doy)
    da_coords["percentile"] = percentiles
    threshold_da = xr.DataArray(thr.reshape(*cell_shape, n_doy, percentiles.size), dims=[*cell_dims, "doy", "percentile"],
                                coords=da_coords)

    add_history(threshold_da, f"Threshold data computed by HDP v{get_version()}.\n")
    stamps = time_axis.date_strings()
    threshold_da.attrs.update({
        "long_name": f"Percentile threshold values for baseline variable '{baseline_data.name}'",
        "baseline_variable": baseline_data.name,
        "baseline_start_time": stamps[0],
        "baseline_end_time": stamps[-1],
        "baseline_calendar": f"{time_axis.calendar}",
        "param_percentiles": str(percentiles),
        "param_noseason": str(no_season),
        "param_rolling_window_size": str(rolling_window_size),
        "param_fixed_value": str(fixed_value),
        "hdp_type": "threshold"
    })

    ds = xr.Dataset(
        data_vars={f"{baseline_data.name}_threshold": threshold_da},
        coords=dict(
            lon=(["lon"], np.asarray(xr.coord_values(baseline_data, "lon"))),
            lat=(["lat"], np.asarray(xr.coord_values(baseline_data, "lat"))),
            doy=np.arange(0, n_doy),
            percentile=percentiles
        ),
        attrs=dict(
            description=f"Extreme heat threshold dataset generated by Heatwave Diagnostics Package (HDP v{get_version()})",
            hdp_version=get_version(),
        )
    )
    xr.set_coord_attrs(ds, "doy", dict(units="day_of_year", baseline_calendar=str(time_axis.calendar)), replace=True)
    return ds


def compute_thresholds(baseline_dataset, percentiles, no_season: bool = False, rolling_window_size: int = 7, fixed_value: float = None):
    """hdp/threshold.py:207-229: one threshold variable per data variable, merged."""
    threshold_datasets = []
    for var_name in baseline_dataset:
        threshold_datasets.append(compute_threshold(baseline_dataset[var_name], percentiles, no_season, rolling_window_size, fixed_value))
    return xr.merge(threshold_datasets)


def compute_threshold_io(baseline_path: str, baseline_var: str, output_path: str, percentiles, no_season: bool = False,
                         rolling_window_size: int = 7, fixed_value: float = None, overwrite: bool = False) -> None:
    """hdp/threshold.py:232-289: thresholds from a netCDF file / zarr store, written back to disk.  The reference's wrapper does not
    run as shipped (``Path.isdir``, undefined ``makedirs``); this one does what it sets out to do, with its errors
    (``FileExistsError``, ``ValueError`` for an unsupported suffix).  netCDF / zarr need xarray; without it (this image) use
    :func:`hdp_b200.io.compute_threshold_io`, the same flow on memory-mapped ``.npy`` files."""
    import os
    from pathlib import Path
    output_path, baseline_path = Path(output_path), Path(baseline_path)
    if output_path.exists() and not overwrite:
        raise FileExistsError(f"Overwrite parameter set to False and file exists at '{output_path}'.")
    if not output_path.parent.exists():
        if overwrite:
            os.makedirs(output_path.parent, exist_ok=True)
        else:
            raise FileExistsError(f"Overwrite parameter set to False and directory '{output_path.parent}' does not exist.")
    if output_path.suffix not in [".zarr", ".nc"]:
        raise ValueError(f"File type '{output_path.suffix}' from '{output_path}' not supported.")
    if not xr.HAVE_XARRAY:
        raise RuntimeError("reading netCDF / zarr needs xarray; hdp_b200.io.compute_threshold_io streams .npy files without it")
    import xarray                                                   # pragma: no cover - no xarray in the build image
    if baseline_path.suffix == ".zarr" and baseline_path.is_dir():  # pragma: no cover
        baseline_data = xarray.open_zarr(baseline_path)[baseline_var]
    else:                                                           # pragma: no cover
        baseline_data = xarray.open_dataset(baseline_path)[baseline_var]
    baseline_data.attrs["baseline_source"] = str(baseline_path)     # pragma: no cover
    threshold_ds = compute_threshold(baseline_data, percentiles, no_season, rolling_window_size, fixed_value)   # pragma: no cover
    if output_path.suffix == ".zarr":                               # pragma: no cover
        threshold_ds.to_zarr(output_path, mode="w" if overwrite else "w-")
    else:                                                           # pragma: no cover
        threshold_ds.to_netcdf(output_path)
