"""hdp.measure on the host (reference hdp/measure.py): unit conversion, NWS heat index, measure Dataset assembly.

This layer stays host-side Python by design (BASELINE.json north_star): it is an elementwise pre-pass, not one of
the two data-parallel hot paths.  Arithmetic follows the reference operation by operation, including where it
happens in float32 (`temp -= 273.15` on a float32 array) and where Numba promotes to float64 (heat_index).
"""
from __future__ import annotations

import numpy as np

from . import xr
from .utils import add_history, get_version

TEMPERATURE_UNITS = ['degC', 'degK', 'degF', 'C', 'K', 'F']
HUMIDITY_UNITS = ["%", "g/g"]


def _with_values(da, values, attrs=None, name=None):
    return xr.with_values(da, values, attrs, name)


def kelvin_to_celsius(temp):
    """hdp/measure.py:10-24: in-place float32 subtraction."""
    attrs = dict(temp.attrs)
    vals = xr.values_of(temp).copy()
    vals -= 273.15
    attrs["units"] = "degC"
    return add_history(_with_values(temp, vals, attrs), "HDP converted units from Kelvin to Celsius.")


def fahrenheit_to_celsius(temp):
    """hdp/measure.py:27-41"""
    attrs = dict(temp.attrs)
    vals = (xr.values_of(temp) - 32) / 1.8
    attrs["units"] = "degC"
    return add_history(_with_values(temp, vals, attrs), "HDP converted units from Fahrenheit to Celsius.")


def celsius_to_fahrenheit(temp):
    """hdp/measure.py:44-58"""
    attrs = dict(temp.attrs)
    vals = (xr.values_of(temp) * 1.8) + 32
    attrs["units"] = "degF"
    return add_history(_with_values(temp, vals, attrs), "HDP converted units from Celsius to Fahrenheit.")


def heat_index(temp, rel_humid) -> np.ndarray:
    """NWS heat-index regression, hdp/measure.py:61-94 (`@nb.vectorize([float32(float32, float32)])`).

    Inside the Numba kernel the float32 arguments meet float64 literals, so sums and products with a literal are
    float64 - but ``temp**2``, ``rel_humid**2`` and ``rel_humid*temp`` involve float32 operands only and are rounded to
    float32 first (``(rel_humid*temp)**2`` twice).  Only with exactly that mix does this NumPy restatement reproduce
    the reference kernel bit for bit (tests/test_host_api.py pins it against outputs of the reference kernel)."""
    t32 = np.asarray(temp, dtype=np.float32)
    rh32 = np.asarray(rel_humid, dtype=np.float32)
    t32, rh32 = np.broadcast_arrays(t32, rh32)
    t, rh = t32.astype(np.float64), rh32.astype(np.float64)
    t_sq = (t32 * t32).astype(np.float64)
    rh_sq = (rh32 * rh32).astype(np.float64)
    rt = rh32 * t32
    rt_sq = (rt * rt).astype(np.float64)
    hi = 0.5 * (t + 61.0 + ((t - 68.0) * 1.2) + (rh * 0.094))
    full = np.full(t.shape, -42.379)
    full = full + 2.04901523 * t
    full = full + 10.14333127 * rh
    full = full + -0.22475541 * t * rh
    full = full + -0.00683783 * t_sq
    full = full + -0.05481717 * rh_sq
    full = full + 0.00122874 * t_sq * rh
    full = full + 0.00085282 * t * rh_sq
    full = full + -0.00000199 * rt_sq
    low = (rh < 13) & (t >= 80) & (t <= 112)
    with np.errstate(invalid="ignore"):
        adj_low = ((13 - rh) / 4) * np.sqrt((np.abs(17 - np.abs(t - 95)) / 17))
    high = (~low) & (rh > 85) & (t >= 80) & (t <= 87)
    adj_high = ((rh - 85) / 10) * ((87 - t) / 5)
    full = np.where(low, full - adj_low, np.where(high, full + adj_high, full))
    return np.where(hi > 80, full, hi).astype(np.float32)


def heat_index_map_wrapper(ds):
    """hdp/measure.py:97-108: ``heat_index`` over the ``temp`` (degF) and ``rh`` (%) variables of a dataset, float32."""
    temp = ds["temp"]
    return _with_values(temp, heat_index(xr.values_of(temp).astype(np.float32), xr.values_of(ds["rh"]).astype(np.float32)))


def apply_heat_index(temp, rh):
    """hdp/measure.py:111-133"""
    assert temp.attrs["units"] == "degF"
    assert rh.attrs["units"] == "%"
    vals = heat_index(xr.values_of(temp).astype(np.float32), xr.values_of(rh).astype(np.float32))
    hi_da = _with_values(temp, vals, name=f"{temp.name}_hi")
    hi_da.attrs["baseline_variable"] = hi_da.name
    return add_history(hi_da, f"Converted to heat index using '{rh.name}' relative humidity, renamed from '{temp.name}' to '{hi_da.name}'.")


def convert_temp_units(temp_ds):
    """hdp/measure.py:136-149"""
    if temp_ds.attrs["units"] == "K" or temp_ds.attrs["units"] == "degK":
        temp_ds = kelvin_to_celsius(temp_ds)
    elif temp_ds.attrs["units"] == "F" or temp_ds.attrs["units"] == "degF":
        temp_ds = fahrenheit_to_celsius(temp_ds)
    return temp_ds


def _as_float32_copy(da):
    """``da.copy(deep=True).astype(np.float32)`` (reference hdp/measure.py:166,182) with ONE pass over the data instead of two
    (the cast already yields a new array), split across threads for GB-sized fields (NumPy releases the GIL in the copy)."""
    src = np.asarray(xr.values_of(da))
    dst = np.empty(src.shape, dtype=np.float32)
    n = src.shape[0] if src.ndim else 0
    if src.size < (1 << 22) or n < 2:
        np.copyto(dst, src, casting="unsafe")
    else:
        import os
        from concurrent.futures import ThreadPoolExecutor
        workers = max(1, min(n, 16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
        bounds = np.linspace(0, n, workers + 1).astype(int)
        with ThreadPoolExecutor(workers) as pool:
            list(pool.map(lambda k: np.copyto(dst[bounds[k]:bounds[k + 1]], src[bounds[k]:bounds[k + 1]], casting="unsafe"), range(workers)))
    return _with_values(da, dst)


def format_standard_measures(temp_datasets: list, rh=None):
    """hdp/measure.py:152-203 - same checks, attrs, variable names and merge order."""
    measures = []
    for temp_ds in temp_datasets:
        temp_ds = _as_float32_copy(temp_ds)                        # copy(deep=True).astype(float32), measure.py:166
        assert "units" in temp_ds.attrs, f"Attribute 'units' not found in '{temp_ds.name}' dataset."
        assert temp_ds.attrs["units"] in TEMPERATURE_UNITS, f"Units for '{temp_ds.name}' must be one of the following: {TEMPERATURE_UNITS}"
        temp_ds.attrs.update({
            "hdp_type": "measure",
            "input_variable": temp_ds.name,
            "baseline_variable": temp_ds.name
        })
        measures.append(convert_temp_units(temp_ds))

    if rh is not None:
        rh = _as_float32_copy(rh)
        assert "units" in rh.attrs, "Attribute 'units' not found in rh dataset."
        assert rh.attrs["units"] in HUMIDITY_UNITS, f"Units for rh must be one of the following: {HUMIDITY_UNITS}"
        if rh.attrs["units"] == "g/g":
            attrs = dict(rh.attrs)
            attrs["units"] = "%"
            rh = _with_values(rh, xr.values_of(rh) * np.float32(100), attrs)
        heat_index_datasets = []
        for measure in measures:
            ftemp_ds = celsius_to_fahrenheit(measure.copy(deep=True))
            heat_index_datasets.append(fahrenheit_to_celsius(apply_heat_index(ftemp_ds, rh)))
        measures.extend(heat_index_datasets)

    agg_ds = xr.merge(measures)
    agg_ds.attrs = {
        "description": f"Heat measurement dataset generated by Heatwave Diagnostics Package (HDP v{get_version()})",
        "hdp_version": get_version(),
    }
    return add_history(agg_ds, f"Dataset aggregated by HDP with measures: {[ds.name for ds in measures]}")
