"""Host logic that needs no GPU: which threshold kernel the library picks for a table / quantile set (the planner of
csrc/threshold.cu + csrc/thr_net.cu through hdp_b200_thresholds_kernel_choice), the compare-exchange network generator, and the
reference's own Numba kernels (oracle/_ref) under the process pool that bench.py times as the CPU baseline."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from hdp_b200 import _lib, _tables as tb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GENERIC, RANKED, SEG, CAND, NET = range(5)


def choice(calendar, years, radius, q):
    end = f"{1960 + years}-12-30" if calendar == "360_day" else f"{1960 + years}-12-31"
    ax = tb.TimeAxis.date_range("1961-01-01", end, calendar)
    wt = tb.window_tables(ax.dayofyr, radius)
    ti = np.ascontiguousarray(wt.time_index, np.int32)
    wr = np.ascontiguousarray(wt.win_rows, np.int32)
    qq = np.ascontiguousarray(q, np.float64)
    info = (ctypes.c_int * 8)()
    rc = _lib.lib().hdp_b200_thresholds_kernel_choice(ti.ctypes.data_as(ctypes.c_void_p), wr.ctypes.data_as(ctypes.c_void_p), len(ax),
                                                      wt.n_doy, wt.n_y, wt.width, qq.ctypes.data_as(ctypes.c_void_p), qq.size, info)
    assert rc == 0
    return list(info)


def test_kernel_choice_bench_shape():
    # cmip6_1deg / lens: 30 years, 15-day window, q = 0.90 .. 0.99 -> the network kernel, 30 samples per row, K = 48 >= 46,
    # three blocks of five rows, the seven mirrored year-end days irregular, 12-warp CTAs with three suffix lists in tensor memory
    info = choice("noleap", 30, 7, np.arange(0.9, 1.0, 0.01))
    assert info == [NET, 30, 48, 3, 5, 7, 12, 3]


@pytest.mark.parametrize("calendar,years,radius,q,want", [
    ("noleap", 30, 7, np.linspace(0.80, 0.99, 20), CAND),        # wide_sweep: positions up to 91 from the top: beyond K = 64
    ("noleap", 30, 7, np.array([0.0, 0.5, 0.9]), SEG),           # the minimum / the median: not among the largest
    ("standard", 30, 15, np.arange(0.9, 1.0, 0.01), RANKED),     # era5: 31-day window (K = 94, W not a multiple of 3 and > 13)
    ("noleap", 30, 10, np.array([0.9, 0.95]), NET),              # W = 21, K = 64
    ("360_day", 20, 2, np.array([0.95]), NET),                   # W = 5: one block per window
    ("noleap", 40, 7, np.array([0.95]), CAND),                   # 40 samples per row: more than a lane's row
])
def test_kernel_choice(calendar, years, radius, q, want):
    assert choice(calendar, years, radius, q)[0] == want


def test_kernel_choice_geometry():
    info = choice("noleap", 30, 10, np.array([0.9, 0.95]))        # W = 21 -> blocks of 7, K = 64: 4-warp CTAs, all six suffix lists in TMEM
    assert info[1:5] == [30, 64, 3, 7] and info[6] == 4 and info[7] == 6
    info = choice("standard", 12, 7, np.arange(0.9, 1.0, 0.01))   # leap calendar, 12 years: rows padded to 16, K = 32
    assert info[:5] == [NET, 16, 32, 3, 5] and info[5] == 7
    info = choice("noleap", 9, 0, np.array([0.5, 0.9]))           # W = 1
    assert info[:5] == [NET, 16, 16, 1, 1] and info[5] == 0


def test_network_generator_verifies():
    # every sorting / merge network of thr_net_gen.cuh: 0/1 principle over all sorted input pairs for the merges, random and 0/1
    # inputs for the sorts; and the committed header is what the generator produces
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_networks.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0 and "all networks verified" in r.stdout, r.stdout + r.stderr
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import gen_networks
        import tempfile
        with tempfile.NamedTemporaryFile("r", suffix=".cuh") as f:
            gen_networks.emit(f.name)
            assert open(f.name).read() == open(os.path.join(ROOT, "hdp_b200", "csrc", "thr_net_gen.cuh")).read()
    finally:
        sys.path.pop(0)


def test_reference_pool_matches_oracle():
    # oracle/_ref (the unmodified reference package, installed by oracle/build.py:build_ref) driven like bench.py drives it:
    # its Numba kernels under a process pool reproduce the oracle port bit for bit
    import oracle
    from oracle import build as obuild, ref_numba, ref_pool
    if obuild.build_ref() is None or not ref_numba.available():
        pytest.skip("no reference tree and no earlier install")
    pytest.importorskip("numba")
    rng = np.random.default_rng(2)
    base_ax = tb.TimeAxis.date_range("1961-01-01", "1966-12-31", "noleap")
    run_ax = tb.TimeAxis.date_range("2001-01-01", "2004-12-31", "noleap")
    C = 6
    xb = (15 + 8 * np.sin(2 * np.pi * base_ax.dayofyr[:, None] / 365) + 3 * rng.standard_normal((len(base_ax), C))).astype(np.float32)
    xr = (17 + 8 * np.sin(2 * np.pi * run_ax.dayofyr[:, None] / 365) + 3 * rng.standard_normal((len(run_ax), C))).astype(np.float32)
    wt = tb.window_tables(base_ax.dayofyr, 7)
    q = np.array([0.9, 0.95])
    defs = [[3, 0, 0], [3, 1, 1]]
    st, dm = tb.hemisphere_ranges(run_ax), tb.doy_map(run_ax.dayofyr)
    south = np.array([0, 1, 0, 1, 0, 1], np.uint8)
    pool = ref_pool.RefPool(2)
    try:
        thr, _ = pool.thresholds(xb, wt.window_samples(), q)
        met, _ = pool.metrics(xr, thr, dm, defs, st.north, st.south, south)
    finally:
        pool.close()
    want_thr = oracle.thresholds_batch(xb, wt.window_samples(), q)
    assert np.array_equal(thr.view(np.uint64), want_thr.view(np.uint64))
    assert np.array_equal(met, oracle.metrics_batch(xr, want_thr, dm, defs, st.north, st.south, south))
