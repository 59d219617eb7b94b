"""CPU oracle for the two HDP hot paths - TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package; ``hdp_b200`` never does.

Two independent restatements of the reference algorithm live here:

* ``hdp_oracle.c`` (loaded through ctypes below) - the primary checker and the timed CPU baseline;
* the pure NumPy/Python functions at the bottom of this file - a slow second opinion for small cases.

Both are pinned against the reference's own known-answer tests and against fixtures produced by
the unmodified reference Numba kernels (``tests/golden/``).  Parity status: PINNED.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Sequence

import numpy as np

from . import build as _build

_i64 = ctypes.c_int64
_p = ctypes.c_void_p
_LIB: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(_build.build())
        L.hdp_oracle_percentiles.argtypes = [_p, _i64, _p, _i64, _i64, _p, ctypes.c_int, _p]
        L.hdp_oracle_thresholds_batch.argtypes = [_p, _i64, _i64, _i64, _i64, _p, _i64, _i64, _p, ctypes.c_int, _p, ctypes.c_int]
        L.hdp_oracle_indicate_hot_days.argtypes = [_p, _p, _p, _i64, _p]
        L.hdp_oracle_indicate_hot_days.restype = None
        L.hdp_oracle_index_heatwaves.argtypes = [_p, _i64, _i64, _i64, _i64, _p]
        for name in ("frequency", "number", "duration", "average"):
            getattr(L, f"hdp_oracle_heatwave_{name}").argtypes = [_p, _i64, _p, _i64, _p]
        L.hdp_oracle_heatwave_frequency.restype = None
        L.hdp_oracle_heatwave_metrics.argtypes = [_p, _p, _p, _i64, _i64, _i64, _i64, _p, _i64, _p]
        L.hdp_oracle_metrics_batch.argtypes = [_p, _i64, _i64, _i64, _i64, _p, _i64, ctypes.c_int, _p, _p, ctypes.c_int,
                                               _p, _p, _i64, _p, _p, ctypes.c_int]
        L.hdp_oracle_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def max_threads() -> int:
    return int(lib().hdp_oracle_max_threads())


# ----------------------------------------------------------------------------------------------
# C oracle wrappers (same signatures as the reference functions they restate)
# ----------------------------------------------------------------------------------------------

def compute_percentiles(temperatures, window_samples, percentiles) -> np.ndarray:
    """reference hdp/threshold.py:52-78 for one cell: f32[T], i64[d,b], f64[P] -> f64[d,P]."""
    temps = _c(temperatures, np.float32)
    win = _c(window_samples, np.int64)
    q = _c(percentiles, np.float64)
    out = np.empty((win.shape[0], q.size), np.float64)
    rc = lib().hdp_oracle_percentiles(_ptr(temps), temps.size, _ptr(win), win.shape[0], win.shape[1], _ptr(q), q.size, _ptr(out))
    assert rc == 0
    return out


def thresholds_batch(temps_tc, window_samples, percentiles, threads: int = 0) -> np.ndarray:
    """All cells of a ``[T, C]`` array -> f64 ``[C, n_doy, P]`` (threshold.py:81-93 loop)."""
    temps = np.asarray(temps_tc)
    assert temps.dtype == np.float32 and temps.ndim == 2
    T, C = temps.shape
    win = _c(window_samples, np.int64)
    q = _c(percentiles, np.float64)
    out = np.empty((C, win.shape[0], q.size), np.float64)
    ld_t, ld_c = (s // 4 for s in temps.strides)
    rc = lib().hdp_oracle_thresholds_batch(_ptr(temps), C, T, ld_t, ld_c, _ptr(win), win.shape[0], win.shape[1],
                                           _ptr(q), q.size, _ptr(out), threads)
    assert rc == 0
    return out


def indicate_hot_days(measure, threshold, doy_map) -> np.ndarray:
    m, th, dm = _c(measure, np.float32), _c(threshold, np.float64), _c(doy_map, np.int64)
    out = np.empty(m.size, np.uint8)
    lib().hdp_oracle_indicate_hot_days(_ptr(m), _ptr(th), _ptr(dm), m.size, _ptr(out))
    return out.astype(bool)


def index_heatwaves(hot_days_ts, min_duration, max_break, max_subs) -> np.ndarray:
    hot = _c(np.asarray(hot_days_ts).astype(bool), np.uint8)
    out = np.empty(hot.size, np.int64)
    rc = lib().hdp_oracle_index_heatwaves(_ptr(hot), hot.size, int(min_duration), int(max_break), int(max_subs), _ptr(out))
    assert rc == 0
    return out


def _season_fn(name, hw_ts, season_ranges, out_dtype):
    hw = _c(hw_ts, np.int64)
    rng = _c(season_ranges, np.int64)
    out = np.empty(rng.shape[0], out_dtype)
    getattr(lib(), f"hdp_oracle_heatwave_{name}")(_ptr(hw), hw.size, _ptr(rng), rng.shape[0], _ptr(out))
    return out


def heatwave_frequency(hw_ts, season_ranges):
    return _season_fn("frequency", hw_ts, season_ranges, np.int64)


def heatwave_number(hw_ts, season_ranges):
    return _season_fn("number", hw_ts, season_ranges, np.int64)


def heatwave_duration(hw_ts, season_ranges):
    return _season_fn("duration", hw_ts, season_ranges, np.int64)


def heatwave_average(hw_ts, season_ranges):
    return _season_fn("average", hw_ts, season_ranges, np.float64)


def compute_heatwave_metrics(measure, threshold, doy_map, min_duration, max_break, max_subs, season_ranges) -> np.ndarray:
    """reference hdp/metric.py:304-341: -> int64[4, Y] = [HWF, HWN, HWD, HWA]."""
    m, th, dm = _c(measure, np.float32), _c(threshold, np.float64), _c(doy_map, np.int64)
    rng = _c(season_ranges, np.int64)
    out = np.empty((4, rng.shape[0]), np.int64)
    rc = lib().hdp_oracle_heatwave_metrics(_ptr(m), _ptr(th), _ptr(dm), m.size, int(min_duration), int(max_break),
                                           int(max_subs), _ptr(rng), rng.shape[0], _ptr(out))
    assert rc == 0
    return out


def metrics_batch(measure_tc, thresholds_cdp, doy_map, defs, seasons_north, seasons_south, is_south, threads: int = 0) -> np.ndarray:
    """The (percentile, definition, cell) sweep of metric.py:344-369 -> int64 ``[P, D, C, 4, Y]``."""
    meas = np.asarray(measure_tc)
    assert meas.dtype == np.float32 and meas.ndim == 2
    T, C = meas.shape
    thr = _c(thresholds_cdp, np.float64)
    assert thr.shape[0] == C
    n_doy, P = thr.shape[1], thr.shape[2]
    dm = _c(doy_map, np.int64)
    df = _c(defs, np.int64).reshape(-1, 3)
    sn, ss = _c(seasons_north, np.int64), _c(seasons_south, np.int64)
    south = np.zeros(C, np.uint8) if is_south is None else _c(is_south, np.uint8)
    Y = sn.shape[0]
    out = np.empty((P, df.shape[0], C, 4, Y), np.int64)
    ld_t, ld_c = (s // 4 for s in meas.strides)
    rc = lib().hdp_oracle_metrics_batch(_ptr(meas), C, T, ld_t, ld_c, _ptr(thr), n_doy, P, _ptr(dm), _ptr(df), df.shape[0],
                                        _ptr(sn), _ptr(ss), Y, _ptr(south), _ptr(out), threads)
    assert rc == 0
    return out


# ----------------------------------------------------------------------------------------------
# Second opinion: pure NumPy / Python restatement (slow; small cases only)
# ----------------------------------------------------------------------------------------------

def np_quantile_row(samples_f64: np.ndarray, q: Sequence[float]) -> np.ndarray:
    """numba/np/arraymath.py:1655-1704 + :1754-1768 evaluated with separately rounded f64 ops."""
    a = np.sort(np.asarray(samples_f64, dtype=np.float64))
    n = a.size
    q = np.asarray(q, dtype=np.float64)
    if np.isnan(a).any() or n == 0 or (n == 1 and not np.isfinite(a[0])):
        return np.full(q.size, np.nan)
    if n == 1:
        return np.full(q.size, a[0])
    out = np.empty(q.size, np.float64)
    for i, qq in enumerate(q):
        pct = np.float64(qq) * np.float64(100.0)
        if pct == 100:
            val = a[-1]
            if not np.all(np.isfinite(a)) and not np.isfinite(val):
                val = np.nan
        elif pct == 0:
            val = a[0]
            if not np.all(np.isfinite(a)):
                npos, nneg = int(np.sum(a == np.inf)), int(np.sum(a == -np.inf))
                nfin = n - (npos + nneg)
                if nfin == 0:
                    val = np.nan
                if npos == 1 and n == 2:
                    val = np.nan
                if nneg > 1:
                    val = np.nan
                if nfin == 1 and npos > 1 and nneg != 1:
                    val = np.nan
        else:
            rank = np.float64(1.0) + np.float64(n - 1) * (pct / np.float64(100.0))
            f = math.floor(rank)
            m = np.float64(rank - f)
            k = int(f - 1)
            if k >= n - 1:
                lower = upper = a[n - 1]
            else:
                lower, upper = a[k], a[k + 1]
            with np.errstate(invalid="ignore"):
                val = lower * (np.float64(1.0) - m) + upper * m
        out[i] = val
    return out


def np_compute_percentiles(temperatures, window_samples, percentiles) -> np.ndarray:
    temps = np.asarray(temperatures, np.float32)
    win = np.asarray(window_samples, np.int64)
    return np.stack([np_quantile_row(temps[win[d]].astype(np.float64), percentiles) for d in range(win.shape[0])])


def py_streaming_metrics(hot: np.ndarray, min_duration: int, max_break: int, max_subs: int, ranges: np.ndarray) -> np.ndarray:
    """Run-streaming restatement of A4-A10: never builds the per-day id array.
    ``ranges`` must already be clamped to 0 <= lo <= hi <= T.  Returns int64[4, Y]."""
    hot = np.asarray(hot).astype(bool)
    T = hot.size
    Y = len(ranges)
    hwf = np.zeros(Y, np.int64)
    hwn = np.zeros(Y, np.int64)
    hwd = np.zeros(Y, np.int64)
    cur_id = np.zeros(Y, np.int64)     # id whose in-season count is being accumulated
    cur_cnt = np.zeros(Y, np.int64)
    in_hw, hw_id, sub = False, 0, 0
    prev_end = None
    t = 0
    while t < T:
        if not hot[t]:
            t += 1
            continue
        s = t
        while t < T and hot[t]:
            t += 1
        e = t
        if prev_end is not None and s - prev_end > max_break:
            in_hw = False                                        # branch B on the preceding -1 transition
        prev_end = e
        length = e - s
        labelled = False
        if length >= min_duration and not in_hw:                 # A
            hw_id += 1
            in_hw = True
            labelled = True
        elif in_hw and sub < max_subs:                           # C
            sub += 1
            labelled = True
        elif in_hw and sub >= max_subs:                          # D
            if length >= min_duration:
                hw_id += 1
                labelled = True
            else:
                in_hw = False
            sub = 0
        if not labelled:
            continue
        for y in range(Y):
            a, b = int(ranges[y][0]), int(ranges[y][1])
            c = min(e, b) - max(s, a)
            if c <= 0:
                continue
            hwf[y] += c
            if cur_id[y] != hw_id:
                cur_id[y] = hw_id
                cur_cnt[y] = 0
                hwn[y] += 1
            cur_cnt[y] += c
            hwd[y] = max(hwd[y], cur_cnt[y])
    hwa = np.where(hwn > 0, hwf // np.maximum(hwn, 1), 0)
    return np.stack([hwf, hwn, hwd, hwa])
