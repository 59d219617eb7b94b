"""Labelled-array layer used by hdp_b200.threshold / .metric / .measure.

With xarray installed, everything here IS xarray: ``DataArray``/``Dataset``/``merge`` are the real classes and the
public functions take and return real xarray objects, exactly like the reference (hdp/threshold.py, hdp/metric.py,
hdp/measure.py).  This image has no xarray/cftime, so a small stand-in with the same
attribute surface (``dims``, ``shape``, ``values``, ``coords``, ``attrs``, ``name``, ``ds[name]``, iteration over
data variables, ``merge``) keeps the host layer importable and testable; it implements only what HDP uses.
"""
from __future__ import annotations

import copy
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

try:                                       # pragma: no cover - not importable in the build image
    import xarray as _xarray
    HAVE_XARRAY = True
except Exception:                          # noqa: BLE001
    _xarray = None
    HAVE_XARRAY = False


class MiniDataArray:
    """Stand-in for xarray.DataArray: eager NumPy data + named dims + coords + attrs."""

    def __init__(self, data, dims: Sequence[str] = None, coords: Dict[str, object] = None, name: str = None, attrs: dict = None):
        self.values = np.asarray(data)
        if dims is None:
            dims = list(coords.keys()) if coords is not None else [f"dim_{i}" for i in range(self.values.ndim)]
        self.dims = tuple(dims)
        if len(self.dims) != self.values.ndim:
            raise ValueError(f"{len(self.dims)} dims for {self.values.ndim}-d data")
        self.coords = OrderedDict()
        for k, v in (coords or {}).items():
            self.coords[k] = v if _is_time_like(v) else np.asarray(v)
        self.name = name
        self.attrs = dict(attrs or {})

    shape = property(lambda self: self.values.shape)
    dtype = property(lambda self: self.values.dtype)
    size = property(lambda self: self.values.size)
    chunks = None

    def __getattr__(self, item):           # da.lat, da.time ... like xarray
        coords = self.__dict__.get("coords", {})
        if item in coords:
            return _CoordView(coords[item])
        raise AttributeError(item)

    def copy(self, deep: bool = True):
        out = copy.copy(self)
        out.values = self.values.copy() if deep else self.values
        out.coords = OrderedDict(self.coords)
        out.attrs = dict(self.attrs)
        return out

    def astype(self, dtype):
        out = self.copy(deep=False)
        out.values = self.values.astype(dtype)
        return out

    def rename(self, name):
        out = self.copy(deep=False)
        out.name = name
        return out

    def compute(self):
        return self

    def transpose(self, *dims):
        order = [self.dims.index(d) for d in dims]
        out = self.copy(deep=False)
        out.values = self.values.transpose(order)
        out.dims = tuple(dims)
        return out


class _CoordView:
    def __init__(self, v):
        self.values = v if _is_time_like(v) else np.asarray(v)
        self.attrs = {}

    @property
    def size(self):
        return len(self.values)


def _is_time_like(v) -> bool:
    return hasattr(v, "dayofyr") and hasattr(v, "calendar") and not isinstance(v, np.ndarray)


class MiniDataset:
    """Stand-in for xarray.Dataset: ordered data variables sharing coords, plus attrs."""

    def __init__(self, data_vars: Dict[str, MiniDataArray] = None, coords: Dict[str, object] = None, attrs: dict = None):
        self.data_vars = OrderedDict()
        self.coords = OrderedDict()
        self.coord_attrs: Dict[str, dict] = {}
        self.attrs = dict(attrs or {})
        for k, v in (coords or {}).items():
            if isinstance(v, tuple):       # (dims, values) form used by the reference at threshold.py:193
                v = v[1]
            self.coords[k] = v if _is_time_like(v) else np.asarray(getattr(v, "values", v))
        for k, v in (data_vars or {}).items():
            self[k] = v

    def __setitem__(self, key, da: MiniDataArray):
        da = da.rename(key)
        self.data_vars[key] = da
        for c, v in da.coords.items():
            self.coords.setdefault(c, v)

    def __getitem__(self, key):
        if key in self.data_vars:
            return self.data_vars[key]
        if key in self.coords:
            view = _CoordView(self.coords[key])
            view.attrs = self.coord_attrs.setdefault(key, {})
            return view
        raise KeyError(key)

    def __iter__(self):
        return iter(self.data_vars)

    def __contains__(self, key):
        return key in self.data_vars or key in self.coords

    def __len__(self):
        return len(self.data_vars)

    def keys(self):
        return self.data_vars.keys()

    def rename(self, mapping: Dict[str, str]):
        out = MiniDataset(attrs=self.attrs)
        out.coords = OrderedDict(self.coords)
        out.coord_attrs = {k: dict(v) for k, v in self.coord_attrs.items()}
        for k, v in self.data_vars.items():
            out.data_vars[mapping.get(k, k)] = v.rename(mapping.get(k, k))
        return out

    def compute(self):
        return self


def mini_merge(objects: Iterable) -> MiniDataset:
    """xarray.merge for the cases HDP produces: same-coordinate variables; attrs of the first Dataset win."""
    out = MiniDataset()
    first = True
    for obj in objects:
        if isinstance(obj, MiniDataArray):
            if obj.name is None:
                raise ValueError("cannot merge an unnamed DataArray")
            obj = MiniDataset({obj.name: obj})
        if first:
            out.attrs = dict(obj.attrs)
            first = False
        for c, v in obj.coords.items():
            if c in out.coords and not _is_time_like(v):
                if not np.array_equal(np.asarray(out.coords[c]), np.asarray(v)):
                    raise ValueError(f"conflicting values for coordinate '{c}' (hdp_b200.xr supports exact joins only)")
            out.coords.setdefault(c, v)
        for c, a in obj.coord_attrs.items():
            out.coord_attrs.setdefault(c, {}).update(a)
        for k, v in obj.data_vars.items():
            out.data_vars[k] = v
    return out


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
