"""Import the UNMODIFIED reference Numba kernels: from oracle/_ref/ (the pip-installed copy that
oracle/build.py:build_ref() makes; it travels to the GPU box) or, failing that, from /root/reference
(build container only).

TEST INFRASTRUCTURE - not product code.  Users: tests/golden/make_golden.py, the cross-check tests, and
bench.py's CPU legs (cpu_baseline / --impl reference), which time these kernels as the reference baseline.
Nothing on the GPU box reads /root/reference: there the only source is oracle/_ref/.

The reference's hdp/threshold.py and hdp/metric.py import xarray, dask.array, cftime and
tqdm at module level; none of those are installed in this image and they are only used by
the xarray/Dask wrapper layer, never by the Numba kernels.  We register permissive stub
modules for them, then import the reference modules as they are.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_INSTALLED = os.path.join(_HERE, "_ref")


def _root() -> str:
    if os.path.isfile(os.path.join(_INSTALLED, "hdp", "metric.py")):
        return _INSTALLED
    return os.environ.get("HDP_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _root()


class _Stub(types.ModuleType):
    """Module whose every attribute is a dummy class (annotations such as
    ``xarray.DataArray`` are evaluated at def time, reference hdp/metric.py:212)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hdp"))


def load():
    """Return (hdp.threshold, hdp.metric, hdp.measure) of the reference, stubbing absent host modules."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("xarray", "dask", "dask.array", "cftime", "tqdm", "tqdm.auto"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    if "dask" in sys.modules and isinstance(sys.modules["dask"], _Stub):
        sys.modules["dask"].array = sys.modules["dask.array"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "hdp" or k.startswith("hdp.")}
    try:
        thr = importlib.import_module("hdp.threshold")
        met = importlib.import_module("hdp.metric")
        mea = importlib.import_module("hdp.measure")
    finally:
        # keep the reference modules reachable only through the returned handles
        for k in [k for k in sys.modules if k == "hdp" or k.startswith("hdp.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)
        if REFERENCE_ROOT in sys.path:
            sys.path.remove(REFERENCE_ROOT)
    return thr, met, mea


class FakeDate:
    """Duck-typed stand-in for a cftime datetime: exposes the attributes the reference's
    table builders dereference (threshold.py:30, metric.py:193-199, metric.py:276)."""

    __slots__ = ("year", "month", "day", "dayofyr", "calendar")

    def __init__(self, year, month, day, dayofyr, calendar):
        self.year, self.month, self.day, self.dayofyr, self.calendar = year, month, day, dayofyr, calendar

    def __repr__(self):
        return f"{self.year:04d}-{self.month:02d}-{self.day:02d}"
