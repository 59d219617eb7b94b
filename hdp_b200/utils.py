"""Provenance helpers and test-data generators (reference hdp/utils.py)."""
from __future__ import annotations

import datetime
from time import time

import numpy as np

from . import synth, xr
from ._tables import TimeAxis

HDP_COMPAT_VERSION = "1.0.2"          # the reference release whose behaviour this package reproduces


def get_time_stamp() -> str:
    return datetime.datetime.fromtimestamp(time()).strftime('%Y-%m-%d %H:%M')


def get_version() -> str:
    """Reference: importlib.metadata.version('hdp_python') (hdp/utils.py:23-24).  If the reference package is
    installed its version is reported, otherwise the release this package mirrors."""
    try:
        from importlib.metadata import version
        return version('hdp_python')
    except Exception:                  # noqa: BLE001
        return HDP_COMPAT_VERSION


def add_history(ds, msg):
    """hdp/utils.py:14-20 - same strings, same order."""
    if "history" in ds.attrs:
        ds.attrs["history"] += f"({get_time_stamp()}) {msg}\n"
    else:
        ds.attrs["history"] = f"({get_time_stamp()}) History metadata initialized by HDP v{get_version()}.\n"
        ds.attrs["history"] += f"({get_time_stamp()}) {msg}\n"
    return ds


def get_func_description(func) -> str:
    """hdp/utils.py:27-36: the prose of a docstring up to its first ``:param`` line, on one line."""
    words = []
    for line in func.__doc__.split("\n"):
        if ":param" in line:
            break
        if line.strip():
            words.append(line.strip() + " ")
    return "".join(words)


def time_axis_of(obj) -> TimeAxis:
    """Integer calendar fields of the ``time`` coordinate (cftime objects with xarray, a TimeAxis with the stand-in)."""
    t = xr.coord_values(obj, "time")
    if isinstance(t, TimeAxis):
        return t
    return TimeAxis.from_datetimes(list(np.asarray(t).ravel()))


def _sample_to_dataarray(sample: synth.SampleData):
    if xr.HAVE_XARRAY:                 # pragma: no cover
        import xarray
        time_values = xarray.date_range(start=sample.time.date_strings()[0][:10], end=sample.time.date_strings()[-1][:10],
                                        freq="D", calendar=sample.time.calendar, use_cftime=True)
    else:
        time_values = sample.time
    return xr.DataArray(sample.values, dims=["lon", "lat", "time"],
                        coords={"lon": sample.lon, "lat": sample.lat, "time": time_values},
                        name=sample.name, attrs={"units": sample.units})


def generate_test_control_dataarray(start_date="1700-01-01", end_date="1749-12-31", grid_shape=(2, 3), add_noise=False, seed=0):
    """hdp/utils.py:53-92 (eager; the reference returns the same values Dask-chunked)."""
    return _sample_to_dataarray(synth.sample_control(start_date, end_date, grid_shape, add_noise, seed))


def generate_test_warming_dataarray(start_date="2000-01-01", end_date="2049-12-31", grid_shape=(2, 3), warming_period=100, add_noise=False):
    """hdp/utils.py:39-42"""
    return _sample_to_dataarray(synth.sample_warming(start_date, end_date, grid_shape, warming_period, add_noise))


def generate_test_rh_dataarray(start_date="2000-01-01", end_date="2049-12-31", grid_shape=(2, 3)):
    """hdp/utils.py:45-50"""
    base = generate_test_control_dataarray(start_date=start_date, end_date=end_date, grid_shape=grid_shape)
    vals = xr.values_of(base)
    vals = np.abs(vals / vals.max() - 0.3)
    return xr.with_values(base, vals, attrs={"units": "g/g"}, name="test_rh_data")
