"""Stand-in for the handful of xarray classes HDP's host layer touches - used ONLY where xarray is not installed (this build image).

Not a labelled-array library: eager NumPy data + named dims + coords + attrs, exact-join merge; exactly the attribute surface
``hdp_b200.{threshold,metric,measure,utils}`` use (``dims``, ``shape``, ``values``, ``coords``, ``attrs``, ``name``, ``ds[name]``,
iteration over data variables).  With xarray present ``hdp_b200.xr`` binds the real classes and nothing here is imported."""
from __future__ import annotations

import copy
from collections import OrderedDict
from typing import Dict, Iterable, Sequence

import numpy as np


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
