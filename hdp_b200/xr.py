"""Labelled-array layer used by hdp_b200.threshold / .metric / .measure.

With xarray installed, everything here IS xarray: ``DataArray``/``Dataset``/``merge`` are the real classes and the
public functions take and return real xarray objects, exactly like the reference (hdp/threshold.py, hdp/metric.py,
hdp/measure.py).  This image has no xarray/cftime; there, and only there, the names are bound to the
small stand-in of ``hdp_b200/_xr_standin.py`` so that the host layer stays importable and testable.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

try:                                       # pragma: no cover - not importable in the build image
    import xarray as _xarray
    HAVE_XARRAY = True
except Exception:                          # noqa: BLE001
    _xarray = None
    HAVE_XARRAY = False


if not HAVE_XARRAY:
    from ._xr_standin import MiniDataArray, MiniDataset, _CoordView, _is_time_like, mini_merge     # noqa: F401


if HAVE_XARRAY:                            # pragma: no cover
    DataArray = _xarray.DataArray
    Dataset = _xarray.Dataset
    merge = _xarray.merge
else:
    DataArray = MiniDataArray
    Dataset = MiniDataset
    merge = mini_merge


# ---------------------------------------------------------------------------------------------- helpers
def values_of(obj) -> np.ndarray:
    """Eager NumPy values (computes Dask-backed xarray objects)."""
    return np.asarray(obj.values)


def coord_values(obj, name: str):
    c = obj.coords[name]
    return getattr(c, "values", c)


def set_coord_attrs(ds, name: str, attrs: dict, replace: bool = False) -> None:
    target = ds[name].attrs
    if replace:
        target.clear()
    target.update(attrs)


def coords_of(da, skip=()) -> "OrderedDict[str, object]":
    """The coordinates of ``da`` in a form ``DataArray(coords=...)`` accepts back.

    With real xarray the coordinate objects themselves are passed on: they carry their own dims, so non-dimension
    coordinates that are not scalar (2-D ``lat``/``lon`` on curvilinear grids) survive, like the reference's
    ``{**measure.coords}``.  Coordinates that live on a skipped dim (``time``, ``member``) are dropped with it.
    The stand-in has dimension coordinates only and hands back their values."""
    out = OrderedDict()
    for name in da.coords:
        c = da.coords[name]
        if name in skip or any(d in skip for d in getattr(c, "dims", ())):
            continue
        out[name] = c if HAVE_XARRAY else coord_values(da, name)
    return out


def non_time_coords(da) -> "OrderedDict[str, object]":
    return coords_of(da, skip=("time",))


def with_values(da, values, attrs=None, name=None):
    """``da`` with new data of the same shape (same dims and coordinates), optionally new attrs / name."""
    if HAVE_XARRAY:                        # pragma: no cover
        out = da.copy(deep=False, data=values)
        out.attrs = dict(da.attrs if attrs is None else attrs)
        if name is not None:
            out = out.rename(name)
        return out
    return DataArray(values, dims=list(da.dims), coords=coords_of(da), name=da.name if name is None else name,
                     attrs=dict(da.attrs if attrs is None else attrs))
