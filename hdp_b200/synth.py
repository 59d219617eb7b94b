"""Synthetic inputs for tests and bench.py (no network, no datasets).

* :func:`sample_control` / :func:`sample_warming` restate the reference's test-data generators
  (reference hdp/utils.py:39-92) without xarray: same formulae, same ``np.random.seed`` use, same
  ``(lon, lat, time)`` order.  They define BASELINE config 1 (README quick-start).
* :func:`gridded_field` is the seeded generator for the CMIP6-like / CESM2-LENS-like / ERA5-like
  configs: seasonal cycle x latitude gradient + trend + AR(1) noise, produced
  directly on the device as a time-major, cell-contiguous float32 ``[T, C]`` array.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

from ._tables import TimeAxis


@dataclass
class SampleData:
    values: np.ndarray      # float64 (lon, lat, time), like the reference DataArray
    lon: np.ndarray
    lat: np.ndarray
    time: TimeAxis
    units: str = "degC"
    name: str = "test_temperature_data"

    def as_time_major(self) -> np.ndarray:
        """float32 [T, C] with cells flattened in (lon, lat) order (the reference's non-time dim order)."""
        v = self.values.astype(np.float32)
        return np.ascontiguousarray(v.reshape(-1, v.shape[-1]).T)

    def cell_lat(self) -> np.ndarray:
        return np.broadcast_to(self.lat[None, :], self.values.shape[:2]).reshape(-1)


def sample_control(start_date="1700-01-01", end_date="1749-12-31", grid_shape=(2, 3), add_noise=False, seed=0) -> SampleData:
    """generate_test_control_dataarray (reference hdp/utils.py:53-92)."""
    time = TimeAxis.date_range(start_date, end_date, "noleap")
    n = len(time)
    t = np.arange(n, dtype=float)
    north = 20 + 2 * np.sin(2 * np.pi * ((270 + t) / 365))
    south = 20 + 2 * np.sin(2 * np.pi * ((90 + t) / 365))
    vals = np.zeros((grid_shape[0], grid_shape[1], n))
    vals[:, grid_shape[1] // 2:, :] = north
    vals[:, :grid_shape[1] // 2, :] = south
    if add_noise:
        np.random.seed(seed)
        vals += np.random.random(vals.shape) * (np.std(vals) / 2)
    lat = np.linspace(-90, 90, grid_shape[1], dtype=float)
    lat_grad = np.broadcast_to(np.abs(lat) / 90, grid_shape)
    vals = vals - 10 * np.broadcast_to(lat_grad[:, :, None], vals.shape)
    return SampleData(vals, np.linspace(-180, 180, grid_shape[0], dtype=float), lat, time)


def sample_warming(start_date="2000-01-01", end_date="2049-12-31", grid_shape=(2, 3), warming_period=100, add_noise=False) -> SampleData:
    """generate_test_warming_dataarray (reference hdp/utils.py:39-42)."""
    base = sample_control(start_date, end_date, grid_shape, add_noise)
    base.values = base.values + np.arange(len(base.time)) / (365 * warming_period)
    return base


def grid_latitudes(n_lat: int, n_lon: int) -> Tuple[np.ndarray, np.ndarray]:
    """Cell-centre latitudes/longitudes of a regular global grid, cells flattened in (lat, lon) order."""
    lat = -90 + (np.arange(n_lat) + 0.5) * (180.0 / n_lat)
    lon = (np.arange(n_lon) + 0.5) * (360.0 / n_lon)
    return np.repeat(lat, n_lon), np.tile(lon, n_lat)


def gridded_field(cell_lat, dayofyr: np.ndarray, *, seed: int, trend: float = 0.0, offset: float = 0.0,
                  sigma: float = 3.0, ar: float = 0.7, device="cuda", chunk_days: int = 2048, cols=None, out=None):
    """float32 ``[T, C]`` on ``device``:
    ``15 + offset + 12 cos(lat) sin(2 pi (doy-110)/365) sgn(lat) - 25 |lat|/90 + trend t/T + sigma AR(1)``.

    The AR(1) term is the stationary filter ``sqrt(1-ar^2) * sum_k ar^k e[t-k]`` truncated at 32 lags
    (ar^32 ~ 1e-5), so every time chunk can be generated independently and reproducibly from the seed.
    ``cols = (g0, g1)`` keeps only cells g0..g1 of the field (a shard): the values are those of the full field, whoever
    generates them, because the noise is always drawn for the whole grid.  ``out``: optional ``[T, g1 - g0]`` float32
    tensor (any strides, e.g. a column block of a larger array) that receives the field.
    """
    import torch
    dev = torch.device(device)
    lat = torch.as_tensor(np.asarray(cell_lat, dtype=np.float32), device=dev)
    C = lat.numel()
    T = int(len(dayofyr))
    doy = torch.as_tensor(np.asarray(dayofyr, dtype=np.float32), device=dev)
    amp = 12.0 * torch.cos(torch.deg2rad(lat)) * torch.where(lat < 0, -1.0, 1.0)
    base = 15.0 + offset - 25.0 * lat.abs() / 90.0
    g0, g1 = (0, C) if cols is None else (int(cols[0]), int(cols[1]))
    if out is None:
        out = torch.empty((T, g1 - g0), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (T, g1 - g0) or out.dtype != torch.float32:
        raise ValueError("out must be float32 [T, cells]")
    lags = 32
    w = (ar ** torch.arange(lags, device=dev, dtype=torch.float32)) * float(np.sqrt(1 - ar * ar)) * sigma
    for t0 in range(0, T, chunk_days):
        t1 = min(T, t0 + chunk_days)
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(seed) * 1_000_003 + t0)
        lo = max(0, t0 - (lags - 1))
        # noise rows lo..t1: regenerate the overlap with the previous chunk from that chunk's seed
        e = torch.empty((t1 - lo, C), dtype=torch.float32, device=dev)
        if lo < t0:
            gprev = torch.Generator(device=dev)
            prev0 = (t0 - 1) // chunk_days * chunk_days
            gprev.manual_seed(int(seed) * 1_000_003 + prev0)
            eprev = torch.randn((t0 - prev0, C), generator=gprev, dtype=torch.float32, device=dev)
            e[: t0 - lo] = eprev[lo - prev0:]
            del eprev
        e[t0 - lo:] = torch.randn((t1 - t0, C), generator=gen, dtype=torch.float32, device=dev)
        if cols is not None:
            e = e[:, g0:g1].contiguous()
        seg = out[t0:t1]
        seg.zero_()
        for k in range(lags):
            a = t0 - lo - k
            if a < 0:
                a0 = -a                      # rows before the start of the series have no noise history
                seg[a0:] += w[k] * e[: t1 - t0 - a0]
            else:
                seg += w[k] * e[a: a + (t1 - t0)]
        tt = torch.arange(t0, t1, device=dev, dtype=torch.float32)
        season = torch.sin(2 * np.pi * (doy[t0:t1] - 110.0) / 365.0)
        seg += base[None, g0:g1] + season[:, None] * amp[None, g0:g1] + (trend * tt / T)[:, None]
        del e
    return out
