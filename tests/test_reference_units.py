"""The reference's own unit tests, replayed against the drop-in functions of the same names on the GPU.

hdp/tests/test_index_heatwaves.py:6-40 and hdp/tests/test_heatwave_{frequency,number,duration,average}.py:6-54 import
``index_heatwaves`` / ``heatwave_*`` from ``hdp.metric`` and call them on hand-written vectors; the first tests below are those
tests with ``hdp`` replaced by ``hdp_b200`` (vectors transcribed in tests/kat.py).  After them: answers probed from the
reference for inputs its tests do not cover, the batched device entry points against the oracle on random series, and the
other array-level functions of ``hdp.metric`` (``indicate_hot_days``, ``compute_heatwave_metrics_wrapper``)."""
import numpy as np
import pytest

import oracle
from kat import INDEX_KAT, SEASON_KAT

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def metric():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hdp_b200 import metric
    return metric


@pytest.mark.parametrize("case", range(len(INDEX_KAT)))
def test_index_heatwaves_reference_cases(metric, case):
    """hdp/tests/test_index_heatwaves.py:7-40 (null series, full series, case1 - case3, each under the definitions (1, 1, 1),
    (1, 0, 1), (0, 0, 1)) and the two sub-event carry-over vectors recorded in SURVEY.md section 8a."""
    mask, answers = INDEX_KAT[case]
    for definition, want in answers:
        got = metric.index_heatwaves(mask, *definition)
        assert got.dtype == np.int64 and np.array_equal(got, want), definition


@pytest.mark.parametrize("name,column", [("heatwave_frequency", 2), ("heatwave_number", 3), ("heatwave_duration", 4), ("heatwave_average", 5)])
def test_season_metric_reference_cases(metric, name, column):       # hdp/tests/test_heatwave_*.py, 8 cases each
    fn = getattr(metric, name)
    for case in SEASON_KAT:
        hw, ranges, want = np.asarray(case[0]), np.array(case[1], dtype=int), case[column]
        got = fn(hw, ranges)
        assert np.array_equal(got, np.asarray(want)), (name, case[0], ranges)
        assert got.dtype == (np.float64 if name == "heatwave_average" else np.int64)
    assert np.array_equal(fn(np.ones(100, dtype=bool), np.array([[0, 100]])), [100 if column != 3 else 1])   # boolean input, full case


def test_probed_edge_cases(metric):
    """Recorded from the reference's Numba functions in the build container."""
    assert metric.index_heatwaves(np.array([1, 1, 1, 0, 1, 1]), 3, -1, 5).tolist() == [1, 1, 1, 0, 0, 0]
    assert metric.index_heatwaves(np.array([1, 1, 1, 0, 1, 1]), -2, 0, -1).tolist() == [1, 1, 1, 0, 2, 2]
    assert metric.index_heatwaves(np.array([2.5, 0, 0, 1]), 1, 1, 1).tolist() == [1, 0, 0, 2]
    assert metric.index_heatwaves(np.zeros(0, dtype=bool), 1, 1, 1).size == 0
    fns = (metric.heatwave_frequency, metric.heatwave_number, metric.heatwave_duration, metric.heatwave_average)
    for v, want in (([1, 1, 2, 2, 2], (5, 2, 3, 3.0)), ([3, 1, 1, 2], (4, 3, 1, 1.0)), ([-1, 0, 1, 1], (2, 2, 2, 1.0)),
                    ([5, 5, 5], (3, 1, 3, 3.0)), ([0, 0], (0, 0, 0, 0.0)), ([2, 1, 2, 1, 0], (4, 2, 2, 2.0))):
        assert tuple(fn(np.array(v), np.array([[0, len(v)]]))[0] for fn in fns) == want, v
    hw = np.array([0, 1, 1, 0, 2, 2, 2, 0])
    ranges = np.array([[-3, 100], [0, 4], [4, 8], [1, 3], [4, 7], [1, 7]])
    assert [fn(hw, ranges).tolist() for fn in fns] == [[2, 2, 3, 2, 3, 5], [1, 1, 1, 1, 1, 2], [2, 2, 3, 2, 3, 3],
                                                       [2.0, 2.0, 3.0, 2.0, 3.0, 2.5]]
    for empty in ([[3, 3]], [[0, 4], [5, 2]]):                       # an empty season slice: the reference raises where it reduces
        assert metric.heatwave_frequency(hw, np.array(empty))[-1] == 0 and metric.heatwave_number(hw, np.array(empty))[-1] == 0
        with pytest.raises(ValueError):
            metric.heatwave_duration(hw, np.array(empty))
        with pytest.raises(ZeroDivisionError):
            metric.heatwave_average(hw, np.array(empty))


@pytest.mark.parametrize("T", [1, 31, 32, 33, 100, 1000, 4097])
def test_index_heatwaves_batched_vs_oracle(T):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hdp_b200 import _core
    rng = np.random.default_rng(T)
    S = 37
    hot = rng.random((S, T)) < rng.choice([0.05, 0.3, 0.5, 0.8, 0.97], S)[:, None]
    hot[3] = True
    hot[4] = False
    hot[5:9, -1] = True                                              # series that end hot
    defs = [[3, 0, 0], [3, 1, 1], [0, 0, 1], [1, 2, 0], [5, 1, 3], [2, 3, 2], [6, 0, 1]]
    got = _core.index_heatwaves_array(torch.as_tensor(hot).cuda(), defs).cpu().numpy()
    assert got.shape == (S, len(defs), T) and got.dtype == np.int64
    for s in range(S):
        for d, definition in enumerate(defs):
            assert np.array_equal(got[s, d], oracle.index_heatwaves(hot[s], *definition)), (s, definition)


def test_season_metrics_batched_vs_oracle():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hdp_b200 import _core
    rng = np.random.default_rng(3)
    for T, Y in ((1, 1), (45, 3), (365 * 3, 6), (700, 40)):
        S = 12
        hw = np.empty((S, T), np.int64)
        for s in range(S):
            if s % 3 == 0:
                hw[s] = oracle.index_heatwaves(rng.random(T) < 0.5, int(rng.integers(0, 5)), int(rng.integers(0, 3)), int(rng.integers(0, 3)))
            elif s % 3 == 1:
                hw[s] = rng.integers(0 if s % 2 else 1, 5, T)
            else:
                hw[s] = rng.integers(-3, 40, T)
        ranges = np.sort(rng.integers(-T - 3, T + 4, (Y, 2)), axis=1)
        ranges[0] = [0, T]
        got = {k: v.cpu().numpy() for k, v in _core.season_metrics_array(torch.as_tensor(hw).cuda(), ranges).items()}
        assert got["HWA"].dtype == np.float64 and got["HWF"].shape == (S, Y)
        for s in range(S):
            assert np.array_equal(got["HWF"][s], oracle.heatwave_frequency(hw[s], ranges))
            assert np.array_equal(got["HWN"][s], oracle.heatwave_number(hw[s], ranges))
            assert np.array_equal(got["HWD"][s], oracle.heatwave_duration(hw[s], ranges))
            assert np.array_equal(got["HWA"][s], oracle.heatwave_average(hw[s], ranges))


def test_building_blocks_compose_to_the_fused_kernel(metric):
    """indicate_hot_days -> index_heatwaves -> heatwave_* (reference compute_heatwave_metrics, hdp/metric.py:329-340) gives what the
    fused path (compute_heatwave_metrics = hdp_b200_metrics) gives, and what the oracle gives."""
    from hdp_b200 import _tables as tb
    rng = np.random.default_rng(8)
    ax = tb.TimeAxis.daily((2001, 1, 1), 6 * 365, "noleap")
    x = (15 + 9 * np.sin(2 * np.pi * (ax.dayofyr - 110) / 365) + 4 * rng.standard_normal(len(ax))).astype(np.float32)
    x[100] = np.nan
    thr = 17 + 7 * np.sin(2 * np.pi * (np.arange(365) - 109) / 365)
    dm = metric.build_doy_map(ax)
    seasons = metric.get_range_indices(ax, (5, 1), (10, 1))
    hot = metric.indicate_hot_days(x, thr, dm)
    assert hot.dtype == bool and np.array_equal(hot, oracle.indicate_hot_days(x, thr, dm))
    for definition in ((3, 0, 0), (3, 1, 1), (2, 2, 2)):
        hw = metric.index_heatwaves(hot, *definition)
        parts = [metric.heatwave_frequency(hw, seasons), metric.heatwave_number(hw, seasons), metric.heatwave_duration(hw, seasons),
                 metric.heatwave_average(hw, seasons).astype(np.int64)]                        # stored into int64: truncation, :340
        fused = metric.compute_heatwave_metrics(x, thr, dm, *definition, seasons)
        assert np.array_equal(np.stack(parts), fused)
        assert np.array_equal(fused, oracle.compute_heatwave_metrics(x, thr, dm, *definition, seasons))


def test_compute_heatwave_metrics_wrapper(metric):
    """hdp/metric.py:344-369: dims (percentile, definition, <cells>, metric, year), int64, same numbers as compute_group_metrics."""
    from hdp_b200 import measure, threshold, utils, xr
    grid = (2, 3)
    base = measure.format_standard_measures([utils.generate_test_control_dataarray(grid_shape=grid, add_noise=True)])
    test = measure.format_standard_measures([utils.generate_test_warming_dataarray(grid_shape=grid, add_noise=True)])
    defs = [[3, 0, 0], [3, 1, 1], [4, 1, 1]]
    thresholds = threshold.compute_thresholds(base, np.array([0.9, 0.95]))
    name = "test_temperature_data"
    m, t = test[name], thresholds[f"{name}_threshold"]
    da = metric.compute_heatwave_metrics_wrapper(m, t, metric.build_doy_map(utils.time_axis_of(m)), defs)
    assert tuple(da.dims) == ("percentile", "definition", "lon", "lat", "metric", "year")
    assert da.shape == (2, 3, 2, 3, 4, 50) and da.dtype == np.int64
    assert list(np.asarray(xr.coord_values(da, "definition"))) == ["3-0-0", "3-1-1", "4-1-1"]
    group = metric.compute_group_metrics(test, thresholds, defs)
    for i, short in enumerate(("HWF", "HWN", "HWD", "HWA")):
        assert np.array_equal(xr.values_of(da)[..., i, :], xr.values_of(group[f"{name}.{name}_threshold.{short}"]))


def test_compute_percentiles_wrapper():
    """hdp/threshold.py:81-93: the gufunc over a labelled array, (<cells>, doy, percentile), attributes kept."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from conftest import bits_equal
    from hdp_b200 import _tables as tb, threshold, xr
    rng = np.random.default_rng(4)
    ax = tb.TimeAxis.daily((1990, 1, 1), 5 * 365, "noleap")
    vals = (10 + 5 * rng.standard_normal((2, len(ax), 3))).astype(np.float32)          # time in the middle
    da = xr.DataArray(vals, dims=["lon", "time", "lat"], coords={"lon": [0.0, 1.0], "time": ax, "lat": [-1.0, 0.0, 1.0]},
                      name="tas", attrs={"units": "degC"})
    win = threshold.datetimes_to_windows(ax, 2)
    q = np.array([0.25, 0.9, 1.0])
    got = threshold.compute_percentiles_wrapper(da, win, q)
    assert tuple(got.dims) == ("lon", "lat", "doy", "percentile") and got.shape == (2, 3, 365, 3) and got.attrs["units"] == "degC"
    for i in range(2):
        for j in range(3):
            assert bits_equal(xr.values_of(got)[i, j], oracle.compute_percentiles(vals[i, :, j], win, q))
