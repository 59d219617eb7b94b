"""The oracle and the host table builders against the UNMODIFIED reference, run live (oracle/_ref: the reference package as
oracle/build.py:build_ref() installs it; its Numba kernels and pure-Python table builders are imported through
oracle/ref_numba.py with the absent wrapper packages stubbed).  The committed golden fixtures pin fixed cases; here the same
comparison runs on fresh random inputs, and the answers that other tests quote as "probed from the reference" are re-probed.
Skipped where neither /root/reference nor an earlier install exists."""
import numpy as np
import pytest

import oracle
from conftest import bits_equal
from hdp_b200 import _tables as tb

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    pytest.importorskip("numba")
    from oracle import build as obuild, ref_numba
    if obuild.build_ref() is None or not ref_numba.available():
        pytest.skip("no reference tree and no earlier install")
    thr_mod, met_mod, mea_mod = ref_numba.load()

    class Ref:
        threshold, metric, measure = thr_mod, met_mod, mea_mod

        @staticmethod
        def dates(axis):
            return np.array([ref_numba.FakeDate(int(y), int(m), int(d), int(j), axis.calendar)
                             for y, m, d, j in zip(axis.year, axis.month, axis.day, axis.dayofyr)], dtype=object)
    return Ref


def test_quoted_answers_are_the_references(ref):
    """What tests/test_seams_host.py and tests/test_reference_units.py assert as "probed from the reference"."""
    m = ref.metric
    b = lambda v: np.array(v, dtype=bool)
    assert m.index_heatwaves(b([1, 1, 1, 0, 1, 1]), 3, -1, 5).tolist() == [1, 1, 1, 0, 0, 0]
    assert m.index_heatwaves(b([1, 1, 1, 0, 1, 1]), -2, 0, -1).tolist() == [1, 1, 1, 0, 2, 2]
    assert m.index_heatwaves(np.array([2.5, 0, 0, 1]), 1, 1, 1).tolist() == [1, 0, 0, 2]
    assert m.index_heatwaves(np.zeros(0, dtype=bool), 1, 1, 1).size == 0
    fns = (m.heatwave_frequency, m.heatwave_number, m.heatwave_duration, m.heatwave_average)
    for v, want in (([1, 1, 2, 2, 2], (5, 2, 3, 3.0)), ([3, 1, 1, 2], (4, 3, 1, 1.0)), ([-1, 0, 1, 1], (2, 2, 2, 1.0)),
                    ([5, 5, 5], (3, 1, 3, 3.0)), ([0, 0], (0, 0, 0, 0.0)), ([2, 1, 2, 1, 0], (4, 2, 2, 2.0))):
        assert tuple(fn(np.array(v, dtype=np.int64), np.array([[0, len(v)]], dtype=np.int64))[0] for fn in fns) == want, v
    hw = np.array([0, 1, 1, 0, 2, 2, 2, 0], dtype=np.int64)
    ranges = np.array([[-3, 100], [0, 4], [4, 8], [1, 3], [4, 7], [1, 7]], dtype=np.int64)
    assert [fn(hw, ranges).tolist() for fn in fns] == [[2, 2, 3, 2, 3, 5], [1, 1, 1, 1, 1, 2], [2, 2, 3, 2, 3, 3],
                                                       [2.0, 2.0, 3.0, 2.0, 3.0, 2.5]]
    empty = np.array([[3, 3]], dtype=np.int64)
    assert m.heatwave_frequency(hw, empty).tolist() == [0] and m.heatwave_number(hw, empty).tolist() == [0]
    with pytest.raises(ValueError):
        m.heatwave_duration(hw, empty)
    with pytest.raises(ZeroDivisionError):
        m.heatwave_average(hw, empty)


def test_oracle_building_blocks_fuzz_vs_reference(ref):
    m = ref.metric
    rng = np.random.default_rng(2026)
    for trial in range(120):
        T = int(rng.integers(1, 300))
        hot = rng.random(T) < rng.choice([0.1, 0.4, 0.7, 0.95])
        definition = (int(rng.integers(-1, 7)), int(rng.integers(-1, 4)), int(rng.integers(-1, 4)))
        hw = m.index_heatwaves(hot, *definition)
        assert np.array_equal(oracle.index_heatwaves(hot, *definition), hw), (trial, definition)
        ids = hw if trial % 3 else rng.integers(-2, 6, T).astype(np.int64)         # what index_heatwaves yields / arbitrary ids
        lo = rng.integers(-T - 2, T, 4)
        ranges = np.stack([lo, lo + rng.integers(1, T + 3, 4)], axis=1).astype(np.int64)
        sl = [ids[a:b] for a, b in ranges]
        ranges = ranges[[len(s) > 0 for s in sl]]                                   # the reference raises on empty slices
        if not len(ranges):
            continue
        assert np.array_equal(oracle.heatwave_frequency(ids, ranges), m.heatwave_frequency(ids, ranges)), trial
        assert np.array_equal(oracle.heatwave_number(ids, ranges), m.heatwave_number(ids, ranges)), trial
        assert np.array_equal(oracle.heatwave_duration(ids, ranges), m.heatwave_duration(ids, ranges)), trial
        assert np.array_equal(oracle.heatwave_average(ids, ranges), m.heatwave_average(ids, ranges)), trial


def test_oracle_paths_fuzz_vs_reference(ref):
    """compute_percentiles (with Numba's np.quantile) and compute_heatwave_metrics on fresh random series, bit for bit."""
    rng = np.random.default_rng(7)
    for trial, (calendar, years, radius) in enumerate((("noleap", 4, 7), ("standard", 5, 2), ("360_day", 3, 15))):
        ax = tb.TimeAxis.date_range("1999-01-01", f"{1998 + years}-12-30" if calendar == "360_day" else f"{1998 + years}-12-31", calendar)
        n_doy = int(ax.dayofyr.max())
        x = (14 + 9 * np.sin(2 * np.pi * (ax.dayofyr - 100) / n_doy) + 4 * rng.standard_normal(len(ax))).astype(np.float32)
        if trial == 1:
            x[rng.integers(0, len(ax), 3)] = np.nan
        win = ref.threshold.datetimes_to_windows(ref.dates(ax), radius)
        q = np.sort(np.r_[rng.random(4), 0.0, 1.0])
        want = np.empty((win.shape[0], q.size))
        ref.threshold.compute_percentiles(x, win, q, want)
        assert bits_equal(oracle.compute_percentiles(x, win, q), want), calendar
        thr = np.where(np.isnan(want[:, 3]), 1e9, want[:, 3])
        dm = ref.metric.build_doy_map(ref.dates(ax))
        seasons = ref.metric.get_range_indices(ref.dates(ax), (5, 1), (10, 1))
        seasons = seasons[(seasons >= 0).all(axis=1)]
        for definition in ((3, 0, 0), (3, 1, 1), (2, 2, 3)):
            got = oracle.compute_heatwave_metrics(x, thr, dm, *definition, seasons)
            assert np.array_equal(got, ref.metric.compute_heatwave_metrics(x, thr, np.asarray(dm), *definition, seasons)), (calendar, definition)


@pytest.mark.parametrize("calendar", ["noleap", "standard", "360_day", "all_leap"])
def test_table_builders_fuzz_vs_reference(ref, calendar):
    """datetimes_to_windows / build_doy_map / get_range_indices of the reference on random time axes (any start date, any
    length, any radius) against hdp_b200/_tables.py."""
    rng = np.random.default_rng(len(calendar))
    for trial in range(6):
        start = (int(rng.integers(1990, 2010)), 1, 1) if trial % 2 == 0 else (int(rng.integers(1990, 2010)), int(rng.integers(1, 13)), int(rng.integers(1, 28)))
        ax = tb.TimeAxis.daily(start, int(rng.integers(400, 1900)), calendar)
        dates = ref.dates(ax)
        r = int(rng.choice([0, 1, 3, 7, 15]))
        assert np.array_equal(tb.window_tables(ax.dayofyr, r).window_samples(), ref.threshold.datetimes_to_windows(dates, r)), (start, r)
        assert np.array_equal(tb.doy_map(ax.dayofyr), ref.metric.build_doy_map(dates))
        for s, e in (((5, 1), (10, 1)), ((11, 1), (4, 1))):
            assert np.array_equal(tb.range_indices(ax, s, e), ref.metric.get_range_indices(dates, s, e)), (start, s, e)


def test_heat_index_fuzz_vs_reference(ref):
    """The NWS heat index ufunc (hdp/measure.py:61-94, float64 inside Numba, float32 out) over every branch of the regression:
    the host mirror is bit-identical on fresh random inputs (the device pre-pass is checked against the same function's golden)."""
    from hdp_b200 import measure
    rng = np.random.default_rng(5)
    t = np.concatenate([rng.uniform(-20, 130, 20000), rng.uniform(79, 113, 20000), rng.uniform(79.5, 87.5, 10000)]).astype(np.float32)
    rh = np.concatenate([rng.uniform(0, 100, 20000), rng.uniform(0, 14, 20000), rng.uniform(84, 100, 10000)]).astype(np.float32)
    t[:5] = [80.0, 87.0, 112.0, np.nan, 95.0]
    rh[:5] = [13.0, 85.0, 12.99, 50.0, np.nan]
    want = ref.measure.heat_index(t, rh)
    got = measure.heat_index(t, rh)
    assert got.dtype == np.float32 and want.dtype == np.float32
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan) and np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32))
