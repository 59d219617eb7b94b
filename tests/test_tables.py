"""hdp_b200._tables (NumPy table builders) against tables produced by the reference's own builders."""
import numpy as np
import pytest

from hdp_b200 import _tables as tb

AXES = {
    "noleap3": lambda: tb.TimeAxis.daily((1990, 1, 1), 3 * 365, "noleap"),
    "std5": lambda: tb.TimeAxis.date_range("1999-01-01", "2003-12-31", "standard"),
    "d360_2": lambda: tb.TimeAxis.daily((2000, 1, 1), 720, "360_day"),
    "noleap_mid": lambda: tb.TimeAxis.daily((2001, 7, 10), 4 * 365 + 100, "noleap"),
    "allleap2": lambda: tb.TimeAxis.daily((2000, 1, 1), 2 * 366, "all_leap"),
}


@pytest.mark.parametrize("name", list(AXES))
def test_time_axis_fields(golden_tables, name):
    ax = AXES[name]()
    for f in ("year", "month", "day", "dayofyr"):
        assert np.array_equal(getattr(ax, f), golden_tables[f"{name}.{f}"])


@pytest.mark.parametrize("name", list(AXES))
@pytest.mark.parametrize("r", [0, 1, 7, 15])
def test_window_tables(golden_tables, name, r):
    wt = tb.window_tables(golden_tables[f"{name}.dayofyr"], r)
    assert np.array_equal(wt.window_samples(), golden_tables[f"{name}.windows_r{r}"])


@pytest.mark.parametrize("name", list(AXES))
def test_doy_map_and_ranges(golden_tables, name):
    ax = AXES[name]()
    assert np.array_equal(tb.doy_map(ax.dayofyr), golden_tables[f"{name}.doy_map"])
    assert np.array_equal(tb.range_indices(ax, (5, 1), (10, 1)), golden_tables[f"{name}.north"])
    assert np.array_equal(tb.range_indices(ax, (11, 1), (4, 1)), golden_tables[f"{name}.south"])


def test_survey_recorded_ranges():
    # SURVEY.md 8a/A12, recorded from the reference: noleap, 3 years
    ax = tb.TimeAxis.daily((1990, 1, 1), 3 * 365, "noleap")
    assert tb.range_indices(ax, (5, 1), (10, 1)).tolist() == [[120, 273], [485, 638], [850, 1003]]
    assert tb.range_indices(ax, (11, 1), (4, 1)).tolist() == [[304, 455], [669, 820], [1034, 1095]]


def test_hemisphere_trim():
    ax = tb.TimeAxis.daily((1990, 1, 1), 3 * 365, "noleap")
    st = tb.hemisphere_ranges(ax)
    assert st.years.tolist() == [1990, 1991, 1992] and st.north.shape == (3, 2)
    ax = tb.TimeAxis.daily((2001, 7, 10), 4 * 365 + 100, "noleap")   # first year has no May 1
    st = tb.hemisphere_ranges(ax)
    assert (st.north >= 0).all() and (st.south >= 0).all()
    assert st.years.size == st.north.shape[0] == st.south.shape[0]


def test_upper_wrap_is_mirrored():
    # doy 364 with r=7 pools rows 359..364 twice and never rows 1..6
    wt = tb.window_tables(np.tile(np.arange(1, 366), 2), 7)
    assert sorted(wt.win_rows[364].tolist()) == sorted([357, 358] + [359, 360, 361, 362, 363, 364] * 2 + [0])
