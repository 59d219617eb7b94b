"""Generate tests/golden/*.npz by running the UNMODIFIED reference (AgentOxygen/HDP) Numba kernels.

Run in the build container only (needs /root/reference and numba):

    python tests/golden/make_golden.py

The reference's wrapper layer (xarray/Dask/cftime) is absent from this image; the Numba kernels
and the pure-Python table builders are imported through oracle/ref_numba.py (stub modules for the
absent host packages, reference sources untouched) and driven with duck-typed date objects.
Everything written here is OUTPUT OF THE REFERENCE on seeded inputs; nothing is computed by
hdp_b200 or by the oracle.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_numba  # noqa: E402
from hdp_b200._tables import TimeAxis  # noqa: E402  (only used to enumerate calendar dates)


def fake_dates(axis: TimeAxis):
    return np.array([ref_numba.FakeDate(int(y), int(m), int(d), int(j), axis.calendar)
                     for y, m, d, j in zip(axis.year, axis.month, axis.day, axis.dayofyr)], dtype=object)


def seasonal_series(rng, axis: TimeAxis, n_cells: int, trend: float = 0.0, noise: float = 3.0) -> np.ndarray:
    """[T, C] float32: seasonal cycle + AR(1) noise + linear trend."""
    T = len(axis)
    t = np.arange(T)
    out = np.empty((T, n_cells), np.float32)
    for c in range(n_cells):
        phase = 110 if c % 2 == 0 else 290
        e = rng.standard_normal(T)
        ar = np.empty(T)
        ar[0] = e[0]
        for i in range(1, T):
            ar[i] = 0.7 * ar[i - 1] + e[i]
        out[:, c] = (15 + 10 * np.sin(2 * np.pi * (axis.dayofyr - phase) / 365.0) + noise * ar * 0.714
                     + trend * t / T).astype(np.float32)
    return out


def main():
    thr_mod, met_mod, mea_mod = ref_numba.load()
    import numba
    meta = dict(numba=numba.__version__, numpy=np.__version__)
    print("reference loaded; numba", numba.__version__)
    rng = np.random.default_rng(20261018)

    # ---------------------------------------------------------------- tables
    tables = {}
    axes = {
        "noleap3": TimeAxis.daily((1990, 1, 1), 3 * 365, "noleap"),
        "std5": TimeAxis.date_range("1999-01-01", "2003-12-31", "standard"),
        "d360_2": TimeAxis.daily((2000, 1, 1), 720, "360_day"),
        "noleap_mid": TimeAxis.daily((2001, 7, 10), 4 * 365 + 100, "noleap"),
        "allleap2": TimeAxis.daily((2000, 1, 1), 2 * 366, "all_leap"),
    }
    for name, ax in axes.items():
        dates = fake_dates(ax)
        tables[f"{name}.year"] = ax.year
        tables[f"{name}.month"] = ax.month
        tables[f"{name}.day"] = ax.day
        tables[f"{name}.dayofyr"] = ax.dayofyr
        for r in (0, 1, 7, 15):
            tables[f"{name}.windows_r{r}"] = thr_mod.datetimes_to_windows(dates, r)
        tables[f"{name}.doy_map"] = met_mod.build_doy_map(dates)
        tables[f"{name}.north"] = met_mod.get_range_indices(dates, (5, 1), (10, 1))
        tables[f"{name}.south"] = met_mod.get_range_indices(dates, (11, 1), (4, 1))
    np.savez_compressed(os.path.join(HERE, "tables.npz"), **tables)
    print("tables.npz", len(tables))

    # ---------------------------------------------------------------- percentiles (path 1)
    pct = {}
    cases = [
        ("noleap6_r7", TimeAxis.daily((1961, 1, 1), 6 * 365, "noleap"), 7, np.arange(0.9, 1.0, 0.01), 4),
        ("std9_r15", TimeAxis.date_range("1991-01-01", "1999-12-31", "standard"), 15, np.linspace(0.80, 0.99, 20), 3),
        ("d360_r2", TimeAxis.daily((2000, 1, 1), 4 * 360, "360_day"), 2, np.array([0.0, 0.25, 0.5, 0.9, 1.0]), 3),
        ("noleap3_r0", TimeAxis.daily((1990, 1, 1), 3 * 365, "noleap"), 0, np.array([0.1, 0.5, 0.95]), 2),
        ("noleap30_r7", TimeAxis.daily((1961, 1, 1), 30 * 365, "noleap"), 7, np.arange(0.9, 1.0, 0.01), 2),
    ]
    for name, ax, r, q, ncell in cases:
        dates = fake_dates(ax)
        win = thr_mod.datetimes_to_windows(dates, r)
        x = seasonal_series(rng, ax, ncell)
        if name == "d360_r2":
            x = np.round(x * 2) / 2                      # heavy ties
            x = x.astype(np.float32)
        out = np.stack([thr_mod.compute_percentiles(np.ascontiguousarray(x[:, c]), win, q.astype(np.float64))
                        for c in range(ncell)])
        pct[f"{name}.dayofyr"] = ax.dayofyr
        pct[f"{name}.radius"] = np.int64(r)
        pct[f"{name}.q"] = q.astype(np.float64)
        pct[f"{name}.x"] = x
        pct[f"{name}.out"] = out
    # special values: NaN / +-inf inside some windows
    ax = TimeAxis.daily((1990, 1, 1), 4 * 365, "noleap")
    dates = fake_dates(ax)
    win = thr_mod.datetimes_to_windows(dates, 3)
    x = seasonal_series(rng, ax, 4)
    x[100, 0] = np.nan
    x[500, 1] = np.inf
    x[900, 1] = np.inf
    x[40, 2] = -np.inf
    x[700, 3] = -np.inf
    x[705, 3] = np.inf
    q = np.array([0.0, 0.3, 0.9, 0.99, 1.0])
    pct["special_r3.dayofyr"] = ax.dayofyr
    pct["special_r3.radius"] = np.int64(3)
    pct["special_r3.q"] = q
    pct["special_r3.x"] = x
    pct["special_r3.out"] = np.stack([thr_mod.compute_percentiles(np.ascontiguousarray(x[:, c]), win, q) for c in range(4)])
    np.savez_compressed(os.path.join(HERE, "percentiles.npz"), **pct)
    print("percentiles.npz", len(pct))

    # ---------------------------------------------------------------- metrics (path 2)
    met = {}
    # (a) index_heatwaves on random masks
    masks, defs_used, ids = [], [], []
    for i in range(60):
        T = int(rng.integers(1, 200))
        frac = rng.uniform(0.05, 0.95)
        m = rng.random(T) < frac
        d = (int(rng.integers(0, 7)), int(rng.integers(0, 4)), int(rng.integers(0, 4)))
        out = met_mod.index_heatwaves(m, *d)
        masks.append(np.pad(m.astype(np.uint8), (0, 200 - T)))
        defs_used.append((T,) + d)
        ids.append(np.pad(out, (0, 200 - T)))
    met["index.masks"] = np.stack(masks)
    met["index.defs"] = np.array(defs_used, np.int64)        # [T, min_duration, max_break, max_subs]
    met["index.ids"] = np.stack(ids).astype(np.int64)
    # (b) full compute_heatwave_metrics sweeps
    sweeps = [
        ("noleap12", TimeAxis.daily((2000, 1, 1), 12 * 365, "noleap"), 6, 2.0,
         np.arange(0.9, 1.0, 0.01), [[3, 0, 0], [3, 1, 1], [4, 0, 0], [4, 1, 1], [5, 0, 0], [5, 1, 1]]),
        ("std7", TimeAxis.date_range("2015-01-01", "2021-12-31", "standard"), 4, 6.0,
         np.array([0.5, 0.8, 0.95]), [[1, 0, 0], [2, 3, 2], [0, 0, 1], [6, 2, 1], [3, 1, 3]]),
        ("noleap_mid5", TimeAxis.daily((2001, 7, 10), 5 * 365 + 30, "noleap"), 3, 3.0,
         np.array([0.85, 0.9]), [[3, 1, 1], [2, 0, 0]]),
    ]
    for name, ax, ncell, trend, q, defs in sweeps:
        dates = fake_dates(ax)
        # thresholds from a Jan-1 aligned baseline of the same calendar (reference percentiles)
        base_ax = TimeAxis.daily((1980, 1, 1), len(TimeAxis.date_range("1980-01-01", "1987-12-31", ax.calendar)), ax.calendar)
        base = seasonal_series(rng, base_ax, ncell)
        win = thr_mod.datetimes_to_windows(fake_dates(base_ax), 7)
        thr = np.stack([thr_mod.compute_percentiles(np.ascontiguousarray(base[:, c]), win, q.astype(np.float64))
                        for c in range(ncell)])                      # [C, n_doy, P]
        x = seasonal_series(rng, ax, ncell, trend=trend)
        doy_map = met_mod.build_doy_map(dates)
        north = met_mod.get_range_indices(dates, (5, 1), (10, 1))
        south = met_mod.get_range_indices(dates, (11, 1), (4, 1))
        # trim like compute_hemisphere_ranges would (keep rows where all four end points are known)
        keep = np.array([(-1 not in np.concatenate([north[i], south[i]])) for i in range(north.shape[0])])
        lo = int(np.argmax(keep))
        hi = lo
        while hi < keep.size and keep[hi]:
            hi += 1
        north, south = north[lo:hi], south[lo:hi]
        is_south = (np.arange(ncell) % 2).astype(np.uint8)
        defs = np.array(defs, np.int64)
        out = np.zeros((q.size, defs.shape[0], ncell, 4, north.shape[0]), np.int64)
        for p in range(q.size):
            for k in range(defs.shape[0]):
                for c in range(ncell):
                    rng_tab = south if is_south[c] else north
                    out[p, k, c] = met_mod.compute_heatwave_metrics(
                        np.ascontiguousarray(x[:, c]), np.ascontiguousarray(thr[c, :, p]), doy_map,
                        int(defs[k, 0]), int(defs[k, 1]), int(defs[k, 2]), rng_tab)
        for key, val in dict(year=ax.year, month=ax.month, day=ax.day, dayofyr=ax.dayofyr, x=x, thr=thr, q=q,
                             doy_map=doy_map, north=north, south=south, is_south=is_south, defs=defs, out=out).items():
            met[f"{name}.{key}"] = val
        print(name, "hot fraction p0:", float(np.mean(x > thr[np.arange(ncell)[None, :], doy_map[:, None], 0])))
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **met)
    print("metrics.npz", len(met))
    # ---------------------------------------------------------------- heat index (measure.py:61-94)
    t = rng.uniform(40, 125, 20000).astype(np.float32)
    rh = rng.uniform(0, 100, 20000).astype(np.float32)
    t[:2000] = rng.uniform(79, 113, 2000).astype(np.float32)      # the two adjustment regimes
    rh[:1000] = rng.uniform(0, 14, 1000).astype(np.float32)
    rh[1000:2000] = rng.uniform(84, 100, 1000).astype(np.float32)
    t[1000:2000] = rng.uniform(79, 88, 1000).astype(np.float32)
    t[2000:2010] = [80, 87, 112, 95, 80.00001, 79.99999, 68, 0, -40, 150]
    rh[2000:2010] = [13, 85, 12.9, 0, 85.1, 50, 100, 50, 50, 100]
    np.savez_compressed(os.path.join(HERE, "measure.npz"), t=t, rh=rh, hi=mea_mod.heat_index(t, rh))
    print("measure.npz")
    with open(os.path.join(HERE, "VERSIONS.txt"), "w") as f:
        f.write("golden fixtures generated by tests/golden/make_golden.py from the unmodified reference\n")
        f.write("reference: AgentOxygen/HDP v1.0.2 at /root/reference\n")
        for k, v in meta.items():
            f.write(f"{k} {v}\n")


if __name__ == "__main__":
    main()
