"""Host-side index tables that feed the two CUDA kernels (NumPy only, no xarray/cftime needed).

These restate the reference's L0 table builders so that the tables become kernel arguments:

* :func:`window_tables`      <- ``datetimes_to_windows``      (reference hdp/threshold.py:12-49)
* :func:`doy_map`            <- ``build_doy_map``             (reference hdp/metric.py:265-277)
* :func:`range_indices`      <- ``get_range_indices``         (reference hdp/metric.py:175-209)
* :func:`hemisphere_ranges`  <- ``compute_hemisphere_ranges`` (reference hdp/metric.py:212-262)

The reference walks arrays of cftime objects; here the same logic runs on integer field arrays
(``year``, ``month``, ``day``, ``dayofyr``) held by :class:`TimeAxis`, which can be built either
from real cftime/duck-typed date objects or from a calendar name without cftime.
Quirks of the reference are reproduced on purpose.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Sequence, Tuple

import numpy as np

_DAYS_NOLEAP = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])
_DAYS_LEAP = np.array([31, 29, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])

_CAL_ALIASES = {
    "noleap": "noleap", "365_day": "noleap",
    "all_leap": "all_leap", "366_day": "all_leap",
    "360_day": "360_day",
    "standard": "standard", "gregorian": "standard",
    "proleptic_gregorian": "proleptic_gregorian",
    "julian": "julian",
}


def _is_leap(year: int, calendar: str) -> bool:
    cal = _CAL_ALIASES[calendar]
    if cal == "noleap" or cal == "360_day":
        return False
    if cal == "all_leap":
        return True
    if cal == "julian" or (cal == "standard" and year < 1583):
        return year % 4 == 0
    return (year % 4 == 0 and year % 100 != 0) or (year % 400 == 0)


def _month_lengths(year: int, calendar: str) -> np.ndarray:
    if _CAL_ALIASES[calendar] == "360_day":
        return np.full(12, 30)
    return _DAYS_LEAP if _is_leap(year, calendar) else _DAYS_NOLEAP


@dataclass
class TimeAxis:
    """Integer calendar fields of a daily time axis (what the reference reads off cftime objects)."""

    year: np.ndarray
    month: np.ndarray
    day: np.ndarray
    dayofyr: np.ndarray      # 1-based, like cftime's ``dayofyr``
    calendar: str

    def __len__(self) -> int:
        return int(self.year.shape[0])

    @staticmethod
    def from_datetimes(times: Iterable) -> "TimeAxis":
        """From cftime (or duck-typed) objects exposing year/month/day/dayofyr/calendar."""
        times = list(times)
        return TimeAxis(
            year=np.fromiter((t.year for t in times), dtype=np.int64, count=len(times)),
            month=np.fromiter((t.month for t in times), dtype=np.int64, count=len(times)),
            day=np.fromiter((t.day for t in times), dtype=np.int64, count=len(times)),
            dayofyr=np.fromiter((t.dayofyr for t in times), dtype=np.int64, count=len(times)),
            calendar=str(getattr(times[0], "calendar", "")) if times else "",
        )

    @staticmethod
    def daily(start: Tuple[int, int, int], n_days: int, calendar: str = "noleap") -> "TimeAxis":
        """Consecutive days from ``start=(year, month, day)`` in a CF calendar, without cftime."""
        if calendar not in _CAL_ALIASES:
            raise ValueError(f"unsupported calendar '{calendar}'")
        year = np.empty(n_days, np.int64)
        month = np.empty(n_days, np.int64)
        day = np.empty(n_days, np.int64)
        doy = np.empty(n_days, np.int64)
        y, m, d = start
        filled = 0
        while filled < n_days:
            ml = _month_lengths(y, calendar)
            cum = np.concatenate([[0], np.cumsum(ml)])
            first = int(cum[m - 1]) + (d - 1)           # 0-based day of year of (m, d)
            n = min(int(cum[-1]) - first, n_days - filled)
            dd = np.arange(first, first + n)
            mm = np.searchsorted(cum, dd, side="right")  # 1-based month
            year[filled:filled + n] = y
            month[filled:filled + n] = mm
            day[filled:filled + n] = dd - cum[mm - 1] + 1
            doy[filled:filled + n] = dd + 1
            filled += n
            y, m, d = y + 1, 1, 1
        return TimeAxis(year, month, day, doy, calendar)

    @staticmethod
    def date_range(start: str, end: str, calendar: str = "noleap") -> "TimeAxis":
        """Inclusive daily range between ISO dates ``YYYY-MM-DD`` (like xarray.date_range(freq='D'))."""
        ys, ms, ds = (int(v) for v in start.split("-"))
        ye, me, de = (int(v) for v in end.split("-"))
        n = 0
        for y in range(ys, ye + 1):
            n += int(_month_lengths(y, calendar).sum())
        head = int(_month_lengths(ys, calendar)[:ms - 1].sum()) + (ds - 1)
        tail_full = int(_month_lengths(ye, calendar).sum())
        tail = tail_full - (int(_month_lengths(ye, calendar)[:me - 1].sum()) + de)
        return TimeAxis.daily((ys, ms, ds), n - head - tail, calendar)

    def date_strings(self) -> list:
        """``str(cftime_obj)`` style stamps (used for the baseline_start/end_time attrs)."""
        return [f"{y:04d}-{m:02d}-{d:02d} 00:00:00" for y, m, d in zip(self.year, self.month, self.day)]


# ----------------------------------------------------------------------------------------------
# Path 1 tables
# ----------------------------------------------------------------------------------------------

@dataclass
class WindowTables:
    """Factored form of the reference's ``window_samples`` table.

    ``window_samples[d] == time_index[win_rows[d]].ravel()`` (reference threshold.py:41-49).
    ``time_index`` keeps the reference's ``-1`` pads (a pad reads the LAST sample of the series,
    because the Numba gufunc indexes ``temperatures[-1]``, threshold.py:77).
    """

    time_index: np.ndarray   # int64 [n_doy, n_y]; row i = time indices of the i-th distinct dayofyr (first-appearance order)
    win_rows: np.ndarray     # int64 [n_doy, W];  rows of time_index pooled by the window of row d
    radius: int

    @property
    def n_doy(self) -> int:
        return int(self.time_index.shape[0])

    @property
    def n_y(self) -> int:
        return int(self.time_index.shape[1])

    @property
    def width(self) -> int:
        return int(self.win_rows.shape[1])

    def window_samples(self) -> np.ndarray:
        """The reference's full ``int64[n_doy, W*n_y]`` table (used by the oracle)."""
        return self.time_index[self.win_rows].reshape(self.n_doy, self.width * self.n_y)


def window_tables(dayofyr: np.ndarray, window_radius: int) -> WindowTables:
    """Restates ``datetimes_to_windows`` (reference hdp/threshold.py:12-49).

    * rows are the distinct ``dayofyr`` values in first-appearance order (:28-33);
    * short rows are padded with -1 (:35-39);
    * the argument the reference calls ``rolling_window_size`` is a RADIUS: width = 2r+1 (:41);
    * slot k of row d takes row ``s = d + r - k``; ``s >= n_doy`` becomes ``n_doy - s`` (:44-47),
      i.e. the upper wrap is mirrored (``s = n_doy + j`` -> row ``-j``), and negative ``s`` index
      from the end exactly like the NumPy fancy index at :48.
    """
    dayofyr = np.asarray(dayofyr, dtype=np.int64)
    r = int(window_radius)
    if r < 0:
        raise ValueError("window radius must be >= 0")
    uniq, first_idx, inverse = np.unique(dayofyr, return_index=True, return_inverse=True)
    order = np.argsort(first_idx, kind="stable")            # unique-slot -> appearance rank
    rank_of_slot = np.empty_like(order)
    rank_of_slot[order] = np.arange(order.size)
    row_of_t = rank_of_slot[inverse]                        # row index of every time step
    n_doy = int(uniq.size)
    counts = np.bincount(row_of_t, minlength=n_doy)
    n_y = int(counts.max()) if n_doy else 0
    time_index = np.full((n_doy, n_y), -1, dtype=np.int64)
    by_row = np.argsort(row_of_t, kind="stable")            # time indices grouped by row, in time order
    col = np.arange(by_row.size) - np.repeat(np.cumsum(counts) - counts, counts)
    time_index[row_of_t[by_row], col] = by_row
    if r > n_doy:
        raise IndexError(f"window radius {r} exceeds the number of distinct days of year {n_doy}")
    d = np.arange(n_doy)[:, None]
    k = np.arange(2 * r + 1)[None, :]
    s = d + r - k
    s = np.where(s >= n_doy, n_doy - s, s)
    s = np.where(s < 0, s + n_doy, s)                       # NumPy negative indexing
    return WindowTables(time_index=time_index, win_rows=s.astype(np.int64), radius=r)


# ----------------------------------------------------------------------------------------------
# Path 2 tables
# ----------------------------------------------------------------------------------------------

def no_season_tables(n_time: int) -> WindowTables:
    """The window of the reference's declared-but-unimplemented ``no_season`` option (hdp/threshold.py:105-106: "calculate a
    single percentile for the entire year"): ONE row that pools every time step.  The kernels take it like any other table
    (``n_doy = 1``, ``n_y = n_time``, ``W = 1``); the matching day-of-year map is all zeros."""
    return WindowTables(np.arange(int(n_time), dtype=np.int64).reshape(1, -1), np.zeros((1, 1), np.int64), 0)


def doy_map(dayofyr: np.ndarray) -> np.ndarray:
    """``build_doy_map`` (reference hdp/metric.py:265-277): ``dayofyr - 1`` per time step."""
    return np.asarray(dayofyr, dtype=np.int64) - 1


def range_indices(axis: TimeAxis, start: Tuple[int, int], end: Tuple[int, int]) -> np.ndarray:
    """``get_range_indices`` (reference hdp/metric.py:175-209).

    Season n is ``[index of the n-th start date, index of the next end date)``; the table has
    ``last.year - first.year + 1`` rows, unfound entries stay -1, and a season still open at the
    end of the series closes at ``T`` in the LAST row (:206-207).
    """
    T = len(axis)
    num_years = int(axis.year[-1] - axis.year[0] + 1)
    ranges = np.zeros((num_years, 2), dtype=np.int64) - 1
    is_start = (axis.month == start[0]) & (axis.day == start[1])
    is_end = (axis.month == end[0]) & (axis.day == end[1])
    n = 0
    looking_for_start = True
    # walk only the candidate dates, in time order (same transitions as the reference's scan)
    for t in np.flatnonzero(is_start | is_end):
        if looking_for_start:
            if is_start[t]:
                looking_for_start = False
                ranges[n, 0] = t
        else:
            if is_end[t]:
                looking_for_start = True
                ranges[n, 1] = t
                n += 1
    if not looking_for_start:
        ranges[-1, -1] = T
    return ranges


@dataclass
class SeasonTables:
    north: np.ndarray    # int64 [Y, 2]  May 1 -> Oct 1  (reference metric.py:221)
    south: np.ndarray    # int64 [Y, 2]  Nov 1 -> Apr 1 of the next year (reference metric.py:222)
    years: np.ndarray    # int64 [Y]     calendar years kept after trimming

    @property
    def n_years(self) -> int:
        return int(self.years.shape[0])


def hemisphere_ranges(axis: TimeAxis) -> SeasonTables:
    """The time-axis part of ``compute_hemisphere_ranges`` (reference hdp/metric.py:221-243).

    Leading/trailing years in which any of the four end points is -1 are trimmed with the
    reference's exact loop, including ``slice_start = year_index`` (not +1) at :231 and the initial
    ``slice_end = north_ranges.size`` at :225.
    """
    north = range_indices(axis, (5, 1), (10, 1))
    south = range_indices(axis, (11, 1), (4, 1))
    slice_start = 0
    slice_end = north.size
    start_identified = False
    for year_index in range(north.shape[0]):
        end_points = np.concatenate([north[year_index], south[year_index]])
        if -1 in end_points and not start_identified:
            slice_start = year_index
            continue
        elif not start_identified:
            start_identified = True
        if start_identified and -1 in end_points:
            slice_end = year_index
            break
    years = np.arange(int(axis.year[0]), int(axis.year[-1]) + 1, 1, dtype=np.int64)
    return SeasonTables(north=north[slice_start:slice_end], south=south[slice_start:slice_end],
                        years=years[slice_start:slice_end])


def is_south(lat: np.ndarray) -> np.ndarray:
    """Hemisphere flag per latitude: ``lat < 0`` -> South, else North (reference metric.py:247-252)."""
    return (np.asarray(lat) < 0).astype(np.uint8)


def clamp_ranges(ranges: np.ndarray, T: int) -> np.ndarray:
    """Resolve ``hw_ts[a:b]`` Python-slice semantics (negative / out-of-range end points) into plain
    ``0 <= lo <= hi <= T`` pairs, so that the kernel never sees a negative index."""
    r = np.asarray(ranges, dtype=np.int64).copy().reshape(-1, 2)
    r = np.where(r < 0, np.maximum(r + T, 0), r)
    r = np.minimum(r, T)
    r[:, 1] = np.maximum(r[:, 1], r[:, 0])
    return r


def definition_labels(hw_definitions: Sequence[Sequence[int]]) -> list:
    """Coordinate strings of the ``definition`` dimension (reference hdp/metric.py:347-350)."""
    return [f"{d[0]}-{d[1]}-{d[2]}" for d in hw_definitions]


def factor_window_samples(window_samples: np.ndarray) -> WindowTables:
    """Recover (time_index, win_rows) from the reference's flat ``int[n_doy, W*n_y]`` table.

    ``datetimes_to_windows`` builds every window row as a concatenation of whole ``time_index`` rows, and the centre
    slot (k = r, sample_index = d) of row d is ``time_index[d]`` itself (reference hdp/threshold.py:41-48).  The number
    of years per row, n_y, is the largest divisor of the row length for which that structure holds."""
    win = np.asarray(window_samples, dtype=np.int64)
    n_doy, b = win.shape
    for W in range(b if b % 2 else b - 1, 0, -2):            # widest window (fewest years per row) first
        if b % W:
            continue
        n_y = b // W
        blocks = win.reshape(n_doy, W, n_y)
        time_index = blocks[:, W // 2, :]
        lookup = {tuple(row): i for i, row in enumerate(time_index)}
        rows = np.full((n_doy, W), -1, dtype=np.int64)
        ok = True
        for d in range(n_doy):
            for k in range(W):
                i = lookup.get(tuple(blocks[d, k]))
                if i is None:
                    ok = False
                    break
                rows[d, k] = i
            if not ok:
                break
        if ok and len(lookup) == n_doy:
            return WindowTables(time_index=time_index.copy(), win_rows=rows, radius=W // 2)
    raise ValueError("window_samples is not a table produced by datetimes_to_windows")
