"""The per-run / per-season logic of the building-block kernels (hdp_b200/csrc/seams.h: index_heatwaves, heatwave_frequency /
number / duration / average, reference hdp/metric.py:11-172) WITHOUT a GPU: tests/seams_host.cpp replays the kernels' lane loops
on the CPU around the same functions the kernels call, and the results are compared with the reference's known answers
(hdp/tests/test_index_heatwaves.py, test_heatwave_*.py via tests/kat.py), with answers probed from the reference for the
inputs its own tests do not cover, and with the oracle on random series.  The kernels themselves are checked by
tests/test_reference_units.py on the GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from kat import INDEX_KAT, SEASON_KAT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("seams") / "libseams_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-shared", "-fPIC", "-I", os.path.join(ROOT, "hdp_b200", "csrc"),
                    os.path.join(ROOT, "tests", "seams_host.cpp"), "-o", so], check=True)
    L = ctypes.CDLL(so)
    p, i64 = ctypes.c_void_p, ctypes.c_int64
    L.seams_host_index_heatwaves.argtypes = [p, i64, i64, i64, i64, p]
    L.seams_host_season_metrics.argtypes = [p, i64, p, ctypes.c_int, p, p, p, p]

    class Host:
        @staticmethod
        def index(hot, definition):
            hot = np.ascontiguousarray(np.asarray(hot) != 0, dtype=np.uint8)
            out = np.full(hot.size, -7, np.int64)
            L.seams_host_index_heatwaves(hot.ctypes.data, hot.size, *[int(v) for v in definition], out.ctypes.data)
            return out

        @staticmethod
        def seasons(hw, ranges):
            hw = np.ascontiguousarray(hw, dtype=np.int64)
            rng = np.ascontiguousarray(ranges, dtype=np.int64).reshape(-1, 2)
            Y = rng.shape[0]
            f, n, d = (np.full(Y, -7, np.int64) for _ in range(3))
            a = np.full(Y, -7.0, np.float64)
            L.seams_host_season_metrics(hw.ctypes.data, hw.size, rng.ctypes.data, Y, f.ctypes.data, n.ctypes.data, d.ctypes.data, a.ctypes.data)
            return f, n, d, a
    return Host


def test_index_heatwaves_reference_kat(host):
    for mask, cases in INDEX_KAT:
        for definition, want in cases:
            assert np.array_equal(host.index(mask, definition), np.asarray(want)), (definition, mask)


def test_index_heatwaves_probed_edge_cases(host):
    # answers recorded from the reference's Numba function (negative / zero parameters, truthiness of non-boolean input)
    assert host.index(np.array([1, 1, 1, 0, 1, 1]), (3, -1, 5)).tolist() == [1, 1, 1, 0, 0, 0]
    assert host.index(np.array([1, 1, 1, 0, 1, 1]), (-2, 0, -1)).tolist() == [1, 1, 1, 0, 2, 2]
    assert host.index(np.array([2.5, 0, 0, 1]), (1, 1, 1)).tolist() == [1, 0, 0, 2]
    assert host.index(np.zeros(0), (1, 1, 1)).size == 0


@pytest.mark.parametrize("T", [1, 31, 32, 33, 64, 65, 500, 4097])
def test_index_heatwaves_random_vs_oracle(host, T):
    rng = np.random.default_rng(T)
    for trial in range(40):
        frac = rng.choice([0.05, 0.3, 0.5, 0.8, 0.97])
        hot = rng.random(T) < frac
        if trial % 7 == 0:
            hot[-1] = True                                       # the series ends hot: the pad day closes the run
        if trial % 11 == 0:
            hot[:] = True
        definition = (int(rng.integers(0, 7)), int(rng.integers(0, 4)), int(rng.integers(0, 4)))
        assert np.array_equal(host.index(hot, definition), oracle.index_heatwaves(hot, *definition)), (definition, trial)


def test_season_metrics_reference_kat(host):
    for hw, ranges, hwf, hwn, hwd, hwa in SEASON_KAT:
        f, n, d, a = host.seasons(hw, ranges)
        assert f.tolist() == hwf and n.tolist() == hwn and d.tolist() == hwd
        assert np.array_equal(a, np.asarray(hwa, np.float64))


def test_season_metrics_probed_edge_cases(host):
    # recorded from the reference: id series without a cold day, unordered ids, negative ids, Python slice semantics
    whole = lambda v: [[0, len(v)]]
    for v, want in (([1, 1, 2, 2, 2], (5, 2, 3, 3.0)), ([3, 1, 1, 2], (4, 3, 1, 1.0)), ([-1, 0, 1, 1], (2, 2, 2, 1.0)),
                    ([5, 5, 5], (3, 1, 3, 3.0)), ([0, 0], (0, 0, 0, 0.0)), ([2, 1, 2, 1, 0], (4, 2, 2, 2.0))):
        f, n, d, a = host.seasons(v, whole(v))
        assert (f[0], n[0], d[0], a[0]) == want, v
    hw = [0, 1, 1, 0, 2, 2, 2, 0]
    f, n, d, a = host.seasons(hw, [[-3, 100], [0, 4], [4, 8], [1, 3], [4, 7], [1, 7]])
    assert f.tolist() == [2, 2, 3, 2, 3, 5] and n.tolist() == [1, 1, 1, 1, 1, 2]
    assert d.tolist() == [2, 2, 3, 2, 3, 3] and a.tolist() == [2.0, 2.0, 3.0, 2.0, 3.0, 2.5]
    f, n, d, a = host.seasons(hw, [[3, 3], [5, 2]])             # empty slices: 0 (the Python mirror raises like the reference)
    assert f.tolist() == [0, 0] and n.tolist() == [0, 0] and d.tolist() == [0, 0] and a.tolist() == [0.0, 0.0]


def test_season_metrics_random_vs_oracle(host):
    rng = np.random.default_rng(11)
    for trial in range(150):
        T = int(rng.integers(1, 400))
        kind = trial % 3
        if kind == 0:                                            # what index_heatwaves produces
            hw = oracle.index_heatwaves(rng.random(T) < rng.choice([0.2, 0.5, 0.9]), int(rng.integers(0, 5)), int(rng.integers(0, 3)),
                                        int(rng.integers(0, 3)))
        elif kind == 1:                                          # arbitrary small ids, cold days or not
            hw = rng.integers(0 if trial % 2 else 1, 5, T)
        else:                                                    # anything, negative ids included
            hw = rng.integers(-3, 40, T)
        Y = int(rng.integers(1, 6))
        ranges = np.sort(rng.integers(-T - 3, T + 4, (Y, 2)), axis=1)
        if trial % 5 == 0:
            ranges[0] = [0, T]
        f, n, d, a = host.seasons(hw, ranges)
        assert np.array_equal(f, oracle.heatwave_frequency(hw, ranges)), trial
        assert np.array_equal(n, oracle.heatwave_number(hw, ranges)), trial
        assert np.array_equal(d, oracle.heatwave_duration(hw, ranges)), trial
        assert np.array_equal(a, oracle.heatwave_average(hw, ranges)), trial
