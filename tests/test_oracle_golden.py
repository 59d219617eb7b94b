"""Pins the CPU oracle (oracle/hdp_oracle.c and its NumPy second opinion) against
(a) the reference's own known-answer tests and (b) fixtures produced by the unmodified reference."""
import numpy as np
import pytest

import oracle
from hdp_b200 import _tables as tb
from conftest import bits_equal
from kat import INDEX_KAT, SEASON_KAT


@pytest.mark.parametrize("case", range(len(INDEX_KAT)))
def test_index_heatwaves_kat(case):
    mask, expectations = INDEX_KAT[case]
    for d, want in expectations:
        assert np.array_equal(oracle.index_heatwaves(mask, *d), np.asarray(want))


@pytest.mark.parametrize("case", range(len(SEASON_KAT)))
def test_season_metrics_kat(case):
    hw, ranges, hwf, hwn, hwd, hwa = SEASON_KAT[case]
    assert np.array_equal(oracle.heatwave_frequency(hw, ranges), hwf)
    assert np.array_equal(oracle.heatwave_number(hw, ranges), hwn)
    assert np.array_equal(oracle.heatwave_duration(hw, ranges), hwd)
    assert np.allclose(oracle.heatwave_average(hw, ranges), hwa, rtol=0, atol=1e-12)


def test_index_heatwaves_golden(golden_metrics):
    g = golden_metrics
    for mask, (T, a, b, c), ids in zip(g["index.masks"], g["index.defs"], g["index.ids"]):
        got = oracle.index_heatwaves(mask[:T], a, b, c)
        assert np.array_equal(got, ids[:T])


@pytest.mark.parametrize("name", ["noleap12", "std7", "noleap_mid5"])
def test_metrics_golden(golden_metrics, name):
    g = golden_metrics
    got = oracle.metrics_batch(g[f"{name}.x"], g[f"{name}.thr"], g[f"{name}.doy_map"], g[f"{name}.defs"],
                               g[f"{name}.north"], g[f"{name}.south"], g[f"{name}.is_south"])
    assert np.array_equal(got, g[f"{name}.out"])
    # the run-streaming restatement (what the CUDA kernel implements) agrees as well
    x, thr, dm = g[f"{name}.x"], g[f"{name}.thr"], g[f"{name}.doy_map"]
    T = x.shape[0]
    for c in range(x.shape[1]):
        ranges = tb.clamp_ranges(g[f"{name}.south"] if g[f"{name}.is_south"][c] else g[f"{name}.north"], T)
        for p in range(thr.shape[2]):
            hot = oracle.indicate_hot_days(x[:, c], thr[c, :, p], dm)
            for k, d in enumerate(g[f"{name}.defs"]):
                got1 = oracle.py_streaming_metrics(hot, int(d[0]), int(d[1]), int(d[2]), ranges)
                assert np.array_equal(got1, g[f"{name}.out"][p, k, c]), (c, p, k)


@pytest.mark.parametrize("name", ["noleap6_r7", "std9_r15", "d360_r2", "noleap3_r0", "noleap30_r7", "special_r3"])
def test_percentiles_golden(golden_percentiles, name):
    g = golden_percentiles
    wt = tb.window_tables(g[f"{name}.dayofyr"], int(g[f"{name}.radius"]))
    x, q, want = g[f"{name}.x"], g[f"{name}.q"], g[f"{name}.out"]
    got = oracle.thresholds_batch(x, wt.window_samples(), q)
    assert bits_equal(got, want)            # 0 ulp against the reference's Numba np.quantile
    if name != "noleap30_r7":               # the NumPy second opinion is slow
        got2 = np.stack([oracle.np_compute_percentiles(x[:, c], wt.window_samples(), q) for c in range(x.shape[1])])
        assert bits_equal(got2, want)


def test_hot_days_semantics():
    # f32 measure vs f64 threshold compared in double; NaN on either side -> False (metric.py:280-301)
    thr = np.array([1.0000000001, np.nan, 2.0])
    x = np.array([np.float32(1.0), 5.0, np.nan, np.float32(1.0000001), 2.0, 2.0000002], np.float32)
    dm = np.array([0, 1, 2, 0, 2, 2])
    assert oracle.indicate_hot_days(x, thr, dm).tolist() == [False, False, False, True, False, True]
