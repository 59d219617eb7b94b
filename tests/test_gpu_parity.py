"""GPU parity: the CUDA paths (through the C ABI) against the reference's golden fixtures and the CPU oracle.

Integer metrics and the hot-day mask must be bit-exact; thresholds are compared at 0 ulp in float64
(stated tolerance: bit-exact up to the sign of zero; NaN positions identical)."""
import numpy as np
import pytest

import oracle
from conftest import bits_equal
from kat import INDEX_KAT

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def core():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hdp_b200 import _core
    return _core


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def ref_layout(out_u16):
    """uint16 [4, P, D, Y, C] -> the reference's int64 [P, D, C, 4, Y]."""
    return out_u16.cpu().numpy().astype(np.int64).transpose(1, 2, 4, 0, 3)


# ------------------------------------------------------------------------------------------ path 1
@pytest.fixture(params=["default", "cand", "seg_all", "seg_heavy", "ranked", "generic"])
def thr_path(request, core):
    """Every threshold path: the default (k_thr_net where the tables and quantiles allow, else the segment kernels), the
    segment kernels without k_thr_net (k_thr_cand + k_thr_seg), k_thr_seg with and without its candidate filter,
    k_thr_ranked, and the generic gather+sort fallback."""
    from hdp_b200 import _lib
    _lib.lib().hdp_b200_thresholds_force_generic({"default": 0, "generic": 1, "ranked": 2, "seg_all": 3, "seg_heavy": 4, "cand": 5}[request.param])
    yield request.param
    _lib.lib().hdp_b200_thresholds_force_generic(0)


@pytest.mark.parametrize("name", ["noleap6_r7", "std9_r15", "d360_r2", "noleap3_r0", "noleap30_r7", "special_r3"])
def test_thresholds_golden(core, golden_percentiles, name, thr_path):
    from hdp_b200 import _tables as tb
    g = golden_percentiles
    wt = tb.window_tables(g[f"{name}.dayofyr"], int(g[f"{name}.radius"]))
    got = core.thresholds_array(dev(g[f"{name}.x"]), wt, g[f"{name}.q"]).cpu().numpy()
    assert bits_equal(got, g[f"{name}.out"])


def test_thresholds_layouts_and_host(core, golden_percentiles):
    from hdp_b200 import _tables as tb
    g, name = golden_percentiles, "noleap6_r7"
    wt = tb.window_tables(g[f"{name}.dayofyr"], 7)
    x, q, want = g[f"{name}.x"], g[f"{name}.q"], g[f"{name}.out"]
    xt = dev(x.T.copy()).t()                                  # [T, C] view of a time-contiguous [C, T] array
    assert xt.stride(0) == 1
    assert bits_equal(core.thresholds_array(xt, wt, q).cpu().numpy(), want)
    assert bits_equal(core.thresholds_host(x, wt, q), want)
    assert bits_equal(core.thresholds_host(np.ascontiguousarray(x.T).T, wt, q), want)


def test_thresholds_random_vs_oracle(core, thr_path):
    from hdp_b200 import _tables as tb
    rng = np.random.default_rng(7)
    ax = tb.TimeAxis.daily((1961, 1, 1), 5 * 365, "noleap")
    wt = tb.window_tables(ax.dayofyr, 7)
    C = 203                                                   # ragged: not a multiple of any tile
    x = (15 + 10 * np.sin(2 * np.pi * ax.dayofyr[:, None] / 365) + 3 * rng.standard_normal((len(ax), C))).astype(np.float32)
    x[rng.integers(0, len(ax), 20), rng.integers(0, C, 20)] = np.nan
    x[rng.integers(0, len(ax), 20), rng.integers(0, C, 20)] = np.inf
    x[rng.integers(0, len(ax), 20), rng.integers(0, C, 20)] = -np.inf
    x[:, 5] = np.round(x[:, 5])                               # heavy ties
    x[:, 6] = 1.5                                             # constant series
    x[100, 7] = 1e30                                          # one outlier squeezes every other sample into one bucket
    x[200, 8], x[300, 8] = -3e38, 3e38                        # finite range overflows float
    x[:, 9] = np.where(rng.random(len(ax)) < 0.5, 1e20, x[:, 9])   # half fill values
    x[:, 10] = np.float32(20.0) + np.arange(len(ax), dtype=np.float32) * np.float32(2e-6)   # many distinct values per bucket
    q = np.array([0.0, 0.1, 0.5, 0.9, 0.95, 0.999, 1.0])
    want = oracle.thresholds_batch(x, wt.window_samples(), q)
    got = core.thresholds_array(dev(x), wt, q).cpu().numpy()
    assert bits_equal(got, want)


def test_thresholds_many_percentiles_and_leap(core, thr_path):
    # 20 percentiles (two select rounds), leap calendar with -1 pads and a 31-day window (ERA5-like shape)
    from hdp_b200 import _tables as tb
    rng = np.random.default_rng(17)
    ax = tb.TimeAxis.date_range("1991-01-01", "2000-12-31", "standard")
    wt = tb.window_tables(ax.dayofyr, 15)
    assert (wt.time_index < 0).any()
    x = (10 * np.sin(2 * np.pi * ax.dayofyr[:, None] / 365) + 3 * rng.standard_normal((len(ax), 37))).astype(np.float32)
    q = np.linspace(0.80, 0.99, 20)
    want = oracle.thresholds_batch(x, wt.window_samples(), q)
    assert bits_equal(core.thresholds_array(dev(x), wt, q).cpu().numpy(), want)


@pytest.mark.parametrize("calendar,years,radius,q", [
    ("noleap", 30, 7, np.arange(0.9, 1.0, 0.01)),                      # the README sweep: 4 candidates per row
    ("noleap", 12, 7, np.array([0.75, 0.9, 0.999, 1.0])),              # the filter's loosest setting / the maximum
    ("standard", 9, 15, np.linspace(0.85, 0.99, 15)),                  # leap rows with -1 pads, 31-day window
    ("360_day", 20, 2, np.array([0.95, 0.97])),
])
def test_thresholds_high_quantiles_candidate_filter(core, thr_path, calendar, years, radius, q):
    # high quantiles only: k_thr_seg orders just the samples above a per-segment bound (candidate filter); ties at the
    # bound, constant series, outliers, steps and trends must not change a single bit
    from hdp_b200 import _tables as tb
    rng = np.random.default_rng(29)
    ax = tb.TimeAxis.date_range("1961-01-01", f"{1960 + years}-12-30" if calendar == "360_day" else f"{1960 + years}-12-31", calendar)
    wt = tb.window_tables(ax.dayofyr, radius)
    T, C = len(ax), 67
    doy = ax.dayofyr[:, None]
    x = (15 + 12 * np.sin(2 * np.pi * (doy - 110) / 365) + 3 * rng.standard_normal((T, C))).astype(np.float32)
    x[:, 0] = np.round(x[:, 0])                               # heavy ties
    x[:, 1] = 2.5                                             # constant
    x[:, 2] = np.where(rng.random(T) < 0.5, 1.0, 2.0)         # two values
    x[:, 3] = np.arange(T, dtype=np.float32) * np.float32(1e-3)   # a trend: the last year holds every window's top
    x[:, 4] = -np.arange(T, dtype=np.float32)
    x[T // 2, 5] = 1e30                                       # one outlier squeezes the candidates into one bucket
    x[:, 6] = np.where(rng.random(T) < 0.9, -5.0, x[:, 6])    # 90 % fill value below the data
    x[:, 7] = np.where(rng.random(T) < 0.3, 99.0, x[:, 7])    # 30 % fill value above the data: ties inside the top
    x[:, 8] = (np.arange(T) % 7).astype(np.float32)
    x[:, 9] = np.float32(20.0) + np.arange(T, dtype=np.float32) * np.float32(2e-6)
    # non-finite samples: the light kernel hands these segments over to k_thr_seg (as it does columns 1, 2 and 7, whose
    # candidates outnumber what it orders)
    x[rng.integers(0, T, 6), 11] = np.nan
    x[rng.integers(0, T, 6), 12] = np.inf
    x[rng.integers(0, T, 6), 13] = -np.inf
    x[T // 3, 14], x[T // 3 + 1, 14] = np.inf, np.nan
    want = oracle.thresholds_batch(x, wt.window_samples(), q)
    got = core.thresholds_array(dev(x), wt, q).cpu().numpy()
    assert bits_equal(got, want)


def test_thresholds_many_handed_over_segments(core):
    # high quantiles on a grid where every other cell has a NaN / an infinity somewhere and a few are constant: thousands of
    # (cell, segment) pairs go from k_thr_cand onto the hand-over list, more than the small k_thr_seg grid behind it has blocks
    # (every block takes several turns); the rest stays on the light kernel
    from hdp_b200 import _tables as tb
    rng = np.random.default_rng(61)
    ax = tb.TimeAxis.date_range("1961-01-01", "1990-12-31", "noleap")
    wt = tb.window_tables(ax.dayofyr, 7)
    T, C = len(ax), 12000
    x = (15 + 12 * np.sin(2 * np.pi * (ax.dayofyr[:, None] - 110) / 365) + 3 * rng.standard_normal((T, C), dtype=np.float32)).astype(np.float32)
    bad = np.arange(0, C, 2)
    x[rng.integers(0, T, bad.size), bad] = np.where(rng.random(bad.size) < 0.5, np.nan, np.inf).astype(np.float32)
    x[:, 5::97] = 3.25
    q = np.arange(0.9, 1.0, 0.01)
    got = core.thresholds_array(dev(x), wt, q).cpu().numpy()
    sel = np.unique(np.concatenate([np.arange(0, 64), rng.integers(0, C, 192), [C - 1]]))
    want = oracle.thresholds_batch(np.ascontiguousarray(x[:, sel]), wt.window_samples(), q)
    assert bits_equal(got[sel], want)
    # NaN rows exactly where a window holds the cell's NaN; every cell without one is finite
    assert np.isfinite(got[1::2]).all()


def _net_launches(core, fn):
    """Runs fn() and returns the names of the threshold kernels it launched."""
    core.timing_enable(True)
    core.timing_read()
    try:
        fn()
    finally:
        names = [n for n, _ in core.timing_read()]
        core.timing_enable(False)
    return names


@pytest.mark.parametrize("calendar,years,radius,q,C", [
    ("noleap", 30, 7, np.arange(0.9, 1.0, 0.01), 203),                 # the bench shape: NY = 30, K = 48, three blocks per window
    ("noleap", 30, 7, np.array([0.95, 1.0]), 64),                      # the maximum (q == 1) beside an interpolated position
    ("noleap", 11, 7, np.array([0.8, 0.9, 0.999]), 97),                # short rows (NY = 32 with pads), K = 48 of 165
    ("standard", 12, 7, np.arange(0.9, 1.0, 0.01), 70),                # leap calendar: 366 rows, -1 pads in the last row
    ("360_day", 20, 2, np.array([0.9, 0.95, 0.97]), 45),               # W = 5: one block per window
    ("noleap", 24, 4, np.array([0.85, 0.9, 0.99]), 33),                # W = 9: three blocks of three rows
    ("noleap", 32, 1, np.array([0.6, 0.9]), 40),                       # W = 3: blocks of one row
    ("noleap", 9, 0, np.array([0.5, 0.9]), 40),                        # W = 1: a window is one row
    ("all_leap", 16, 3, np.array([0.75, 0.9]), 50),                    # W = 7, 366-day years
    ("noleap", 30, 10, np.array([0.9, 0.95]), 70),                     # W = 21: blocks of seven rows, K = 64
    ("noleap", 6, 7, np.array([0.9, 0.99]), 40),                       # NY = 8, K = 16
])
def test_thresholds_network_kernel(core, calendar, years, radius, q, C):
    # k_thr_net (lane = cell, sorting / merge networks): regular windows through the block decomposition, the mirrored year-end
    # days row by row, ragged tiles, several chunks per tile, ties / constants / trends, and cells with NaN / +-inf samples
    # that it must hand to k_thr_seg.  Bit-exact against the oracle, and identical to the segment kernels.
    from hdp_b200 import _lib, _tables as tb
    rng = np.random.default_rng(101 + years)
    end = f"{1960 + years}-12-30" if calendar == "360_day" else f"{1960 + years}-12-31"
    ax = tb.TimeAxis.date_range("1961-01-01", end, calendar)
    wt = tb.window_tables(ax.dayofyr, radius)
    T = len(ax)
    doy = ax.dayofyr[:, None]
    x = (15 + 12 * np.sin(2 * np.pi * (doy - 110) / 365) + 3 * rng.standard_normal((T, C))).astype(np.float32)
    x[:, 0] = np.round(x[:, 0])                               # heavy ties
    x[:, 1] = 2.5                                             # constant
    x[:, 2] = np.arange(T, dtype=np.float32) * np.float32(1e-3)   # a trend: one row holds every window's top
    x[:, 3] = -np.arange(T, dtype=np.float32)
    x[T // 2, 4] = 3e38                                       # near the pad value's magnitude
    x[:, 5] = -3.402823466e+38                                # samples EQUAL to the kernel's pad value
    x[:, 6] = np.where(rng.random(T) < 0.3, 99.0, x[:, 6])    # ties inside the top
    x[rng.integers(0, T, 3), 11] = np.nan
    x[rng.integers(0, T, 3), 12] = np.inf
    x[rng.integers(0, T, 3), 13] = -np.inf
    x[T - 1, 14] = np.nan                                     # the sample every -1 pad reads
    x[0, C - 1] = np.inf                                      # last cell of a ragged tile
    want = oracle.thresholds_batch(x, wt.window_samples(), q)
    _lib.lib().hdp_b200_thresholds_force_generic(0)
    names = _net_launches(core, lambda: core.thresholds_array(dev(x), wt, q))
    assert "k_thr_net" in names, names
    got = core.thresholds_array(dev(x), wt, q).cpu().numpy()
    assert bits_equal(got, want)
    xt = dev(x.T.copy()).t()                                  # time-contiguous input: transposed on the device first
    assert bits_equal(core.thresholds_array(xt, wt, q).cpu().numpy(), want)
    try:
        _lib.lib().hdp_b200_thresholds_force_generic(6)      # every list in shared memory (no tensor-memory variant)
        names = _net_launches(core, lambda: core.thresholds_array(dev(x), wt, q))
        assert "k_thr_net" in names
        assert bits_equal(core.thresholds_array(dev(x), wt, q).cpu().numpy(), want)
        _lib.lib().hdp_b200_thresholds_force_generic(5)      # the segment kernels instead
        names = _net_launches(core, lambda: core.thresholds_array(dev(x), wt, q))
        assert "k_thr_net" not in names
        assert bits_equal(core.thresholds_array(dev(x), wt, q).cpu().numpy(), want)
    finally:
        _lib.lib().hdp_b200_thresholds_force_generic(0)


def test_thresholds_network_kernel_many_tiles_and_bad_cells(core):
    # thousands of cells, every third one with a NaN / an infinity somewhere: k_thr_net marks them and k_thr_seg's small grid
    # works through more hand-over entries than it has blocks; all other cells stay on the network kernel
    from hdp_b200 import _tables as tb
    rng = np.random.default_rng(67)
    ax = tb.TimeAxis.date_range("1961-01-01", "1990-12-31", "noleap")
    wt = tb.window_tables(ax.dayofyr, 7)
    T, C = len(ax), 9001
    x = (15 + 12 * np.sin(2 * np.pi * (ax.dayofyr[:, None] - 110) / 365) + 3 * rng.standard_normal((T, C), dtype=np.float32)).astype(np.float32)
    bad = np.arange(0, C, 3)
    x[rng.integers(0, T, bad.size), bad] = np.where(rng.random(bad.size) < 0.5, np.nan, -np.inf).astype(np.float32)
    q = np.arange(0.9, 1.0, 0.01)
    got = core.thresholds_array(dev(x), wt, q).cpu().numpy()
    sel = np.unique(np.concatenate([np.arange(0, 96), rng.integers(0, C, 160), [C - 2, C - 1]]))
    want = oracle.thresholds_batch(np.ascontiguousarray(x[:, sel]), wt.window_samples(), q)
    assert bits_equal(got[sel], want)
    assert np.isfinite(got[1::3]).all() and np.isfinite(got[2::3]).all()


@pytest.mark.parametrize("units", ["degK", "degF"])
def test_fused_unit_conversion(core, units, thr_path):
    # raw Kelvin / Fahrenheit samples handed straight to the kernels (input_unit of the C ABI): bit-identical to converting first
    # with the host mirror of hdp.measure (pinned against the reference kernel outputs in tests/golden/measure.npz), on every
    # threshold path, for the hot-day comparison and through the host pipeline; NaN / inf cells included
    from hdp_b200 import _tables as tb, measure as hm, xr
    rng = np.random.default_rng(77)
    base_ax = tb.TimeAxis.date_range("1961-01-01", "1990-12-31", "noleap")
    run_ax = tb.TimeAxis.date_range("2001-01-01", "2006-12-31", "noleap")
    C = 71
    season = lambda ax: 15 + 10 * np.sin(2 * np.pi * (ax.dayofyr[:, None] - 110) / 365)          # noqa: E731
    to_raw = (lambda c: c + 273.15) if units == "degK" else (lambda c: c * 1.8 + 32)
    xb = to_raw(season(base_ax) + 3 * rng.standard_normal((len(base_ax), C))).astype(np.float32)
    xr_ = to_raw(season(run_ax) + 2 + 3 * rng.standard_normal((len(run_ax), C))).astype(np.float32)
    xb[rng.integers(0, len(base_ax), 3), 5] = np.nan
    xb[rng.integers(0, len(base_ax), 3), 6] = np.inf
    conv = hm.kelvin_to_celsius if units == "degK" else hm.fahrenheit_to_celsius
    as_da = lambda a: xr.DataArray(a, dims=["time", "cell"], coords={}, name="t", attrs={"units": units})   # noqa: E731
    cb, cr = xr.values_of(conv(as_da(xb))), xr.values_of(conv(as_da(xr_)))
    assert cb.dtype == np.float32
    wt = tb.window_tables(base_ax.dayofyr, 7)
    q = np.arange(0.9, 1.0, 0.01)
    want = oracle.thresholds_batch(np.ascontiguousarray(cb), wt.window_samples(), q)
    got = core.thresholds_array(dev(xb), wt, q, units=units)
    assert bits_equal(got.cpu().numpy(), want)
    assert bits_equal(core.thresholds_array(dev(xb.T.copy()).t(), wt, q, units=units).cpu().numpy(), want)    # transposed + converted
    if thr_path != "default":
        return
    assert bits_equal(core.thresholds_host(xb, wt, q, units=units), want)
    st = tb.hemisphere_ranges(run_ax)
    thr = np.nan_to_num(want, nan=1e9, posinf=1e9)
    args = (tb.doy_map(run_ax.dayofyr), [[3, 0, 0], [3, 1, 1], [4, 1, 1]], st.north, st.south, (np.arange(C) % 2).astype(np.uint8))
    want_met = oracle.metrics_batch(np.ascontiguousarray(cr), thr, *args)
    out = core.metrics_array(dev(xr_), dev(thr), *args, units=units)
    assert np.array_equal(ref_layout(out), want_met)
    assert np.array_equal(core.metrics_host(xr_, thr, *args, units=units).astype(np.int64).transpose(1, 2, 4, 0, 3), want_met)


def test_thresholds_errors(core):
    from hdp_b200 import _tables as tb, _lib
    ax = tb.TimeAxis.daily((1961, 1, 1), 2 * 365, "noleap")
    wt = tb.window_tables(ax.dayofyr, 1)
    x = torch.zeros((len(ax), 4), dtype=torch.float32, device="cuda")
    with pytest.raises(_lib.HdpB200Error) as e:
        core.thresholds_array(x, wt, [0.5, 1.5])              # reference: ValueError('Quantiles must be in the range [0, 1]')
    assert e.value.code == -1
    with pytest.raises(_lib.HdpB200Error):
        core.thresholds_array(x, wt, [float("nan")])
    assert core.thresholds_array(x[:, :0], wt, [0.5]).shape == (0, 365, 1)


# ------------------------------------------------------------------------------------------ path 2
@pytest.fixture(params=["filter", "nofilter"], autouse=True)
def scan_filter(request, core):
    """Every test below runs twice: with k_scan's run filter (short runs after long breaks are dropped before the state
    machines see them) and with every run fed through; both must match the reference bit for bit."""
    from hdp_b200 import _lib
    _lib.lib().hdp_b200_metrics_run_filter(1 if request.param == "filter" else 0)
    yield request.param
    _lib.lib().hdp_b200_metrics_run_filter(1)


@pytest.mark.parametrize("name", ["noleap12", "std7", "noleap_mid5"])
def test_metrics_golden(core, golden_metrics, name):
    g = golden_metrics
    args = (g[f"{name}.doy_map"], g[f"{name}.defs"], g[f"{name}.north"], g[f"{name}.south"], g[f"{name}.is_south"])
    out = core.metrics_array(dev(g[f"{name}.x"]), dev(g[f"{name}.thr"]), *args)
    assert np.array_equal(ref_layout(out), g[f"{name}.out"])
    # host-buffer entry point, and a time-contiguous device layout
    out_h = core.metrics_host(g[f"{name}.x"], g[f"{name}.thr"], *args)
    assert np.array_equal(out_h, out.cpu().numpy())
    xt = dev(g[f"{name}.x"].T.copy()).t()
    assert np.array_equal(core.metrics_array(xt, dev(g[f"{name}.thr"]), *args).cpu().numpy(), out_h)


@pytest.mark.parametrize("name", ["noleap12", "std7", "noleap_mid5"])
def test_hot_days_golden_inputs(core, golden_metrics, name):
    g = golden_metrics
    x, thr, dm = g[f"{name}.x"], g[f"{name}.thr"], g[f"{name}.doy_map"]
    got = core.hot_days_array(dev(x), dev(thr), dm).cpu().numpy()
    for c in range(x.shape[1]):
        for p in range(thr.shape[2]):
            assert np.array_equal(got[p, :, c].astype(bool), oracle.indicate_hot_days(x[:, c], thr[c, :, p], dm))


def test_hot_days_f32_vs_f64_boundary(core):
    # thresholds a hair above / below / equal to representable f32 values: the f32 round-down trick must
    # agree with the reference's double-precision compare ("f32-vs-f64 compare")
    rng = np.random.default_rng(3)
    n_doy, T, C = 40, 80, 64
    base = rng.standard_normal((n_doy, C)).astype(np.float32)
    thr = base.astype(np.float64)
    thr[::3] = np.nextafter(thr[::3], np.inf)
    thr[1::3] = np.nextafter(thr[1::3], -np.inf)
    thr[5] = np.nan
    thr[6] = np.inf
    thr[7] = -np.inf
    thr[8] = 1e300
    thr[9] = -1e300
    dm = np.arange(T) % n_doy
    x = base[dm].copy()
    x[::2] = np.nextafter(x[::2], np.float32(np.inf))
    x[3] = np.nan
    x[4] = np.inf
    x[5 + n_doy] = -np.inf
    thr_cdp = np.ascontiguousarray(thr.T[:, :, None])         # [C, n_doy, 1]
    got = core.hot_days_array(dev(x), dev(thr_cdp), dm).cpu().numpy()[0]
    want = np.stack([oracle.indicate_hot_days(x[:, c], thr_cdp[c, :, 0], dm) for c in range(C)], axis=1)
    assert np.array_equal(got.astype(bool), want)


def _mask_case(core, masks, defs, north, south, is_south):
    """Feed 0/1 masks as the measure against a 0.5 threshold: the hot-day mask IS the input."""
    masks = np.asarray(masks)
    T, C = masks.shape
    x = masks.astype(np.float32)
    thr = np.full((C, 1, 1), 0.5)
    dm = np.zeros(T, np.int64)
    out = ref_layout(core.metrics_array(dev(x), dev(thr), dm, defs, north, south, is_south))
    want = oracle.metrics_batch(x, thr, dm, defs, north, south, is_south)
    assert np.array_equal(out, want)
    return out


def test_reference_kat_masks(core):
    # the reference's own index_heatwaves vectors (hdp/tests/test_index_heatwaves.py), pushed through the
    # metric kernel: HWF/HWN/HWD/HWA over one whole-series season and over a two-season split
    for mask, expectations in INDEX_KAT:
        T = mask.size
        defs = [list(d) for d, _ in expectations]
        for ranges in ([[0, T]], [[0, 8], [8, T]], [[2, 5], [5, 6], [9, T - 1]]):
            out = _mask_case(core, mask[:, None].repeat(3, 1), defs, ranges, ranges, [0, 1, 0])
            for k, (d, ids) in enumerate(expectations):
                ids = np.asarray(ids)
                assert np.array_equal(out[0, k, 0, 0], oracle.heatwave_frequency(ids, ranges))
                assert np.array_equal(out[0, k, 0, 1], oracle.heatwave_number(ids, ranges))
                assert np.array_equal(out[0, k, 0, 2], oracle.heatwave_duration(ids, ranges))


def test_overlapping_and_odd_seasons(core):
    rng = np.random.default_rng(11)
    T, C = 300, 40
    masks = rng.random((T, C)) < 0.45
    defs = [[1, 1, 1], [3, 0, 0], [2, 2, 3], [0, 0, 1]]
    north = [[0, 5], [0, 10], [20, 30], [42, 50], [100, 100], [-50, -1], [250, 400]]     # overlapping, empty, negative, beyond T
    south = [[10, 60], [60, 61], [61, 200], [150, 260], [0, 300], [299, 300], [5, 5]]
    _mask_case(core, masks, defs, north, south, rng.integers(0, 2, C))


def test_long_runs_span_seasons(core):
    T, C = 2000, 33
    masks = np.ones((T, C), bool)
    masks[700, :] = False
    masks[1500:1503, 1::2] = False
    seasons = [[100, 250], [465, 615], [830, 980], [1195, 1345], [1560, 1710]]
    _mask_case(core, masks, [[3, 0, 0], [3, 1, 1], [5, 2, 0]], seasons, seasons, None)


def test_metrics_random_sweep_vs_oracle(core):
    from hdp_b200 import _tables as tb
    rng = np.random.default_rng(5)
    ax = tb.TimeAxis.date_range("2001-01-01", "2009-12-31", "standard")
    T, C, P = len(ax), 171, 10
    doy = ax.dayofyr
    x = (15 + 10 * np.sin(2 * np.pi * (doy[:, None] - 110) / 365) + 4 * rng.standard_normal((T, C))
         + 3 * np.arange(T)[:, None] / T).astype(np.float32)
    x[rng.integers(0, T, 50), rng.integers(0, C, 50)] = np.nan
    thr = 15 + 10 * np.sin(2 * np.pi * (np.arange(366)[None, :, None] - 110) / 365) + np.linspace(1, 9, P)[None, None, :] \
        + rng.standard_normal((C, 366, 1))
    st = tb.hemisphere_ranges(ax)
    defs = [[a, b, c] for a in (3, 4, 5, 6) for b in (0, 1, 2) for c in (0, 1)]       # the 24-definition grid
    is_south = (np.arange(C) < C // 2).astype(np.uint8)
    args = (tb.doy_map(doy), defs, st.north, st.south, is_south)
    out = core.metrics_array(dev(x), dev(thr), *args)
    want = oracle.metrics_batch(x, thr, *args)
    assert np.array_equal(ref_layout(out), want)
    hwf, hwn, hwd, hwa = (out[i].cpu().numpy().astype(np.int64) for i in range(4))
    assert (hwf >= hwd).all() and (hwd >= hwa).all()           # reference invariant, hdp/tests/test_workflow.py:52-53
    assert np.array_equal(hwa, np.where(hwn > 0, hwf // np.maximum(hwn, 1), 0))


@pytest.mark.parametrize("D", [1, 2, 7, 8, 9, 17, 25, 32])
def test_metrics_definition_counts_and_ranges(core, D):
    # every kernel instantiation (definition pairs x sub-event counter planes), incl. large max_subs / min_duration
    rng = np.random.default_rng(100 + D)
    T, C = 1500, 70
    masks = rng.random((T, C)) < rng.uniform(0.2, 0.8, C)[None, :]
    big = [1, 3, 9, 200, 5000][D % 5]
    defs = np.stack([rng.integers(0, 9, D), rng.integers(0, 4, D), rng.integers(0, big + 1, D)], axis=1)
    defs[0] = [40, 0, 0] if D > 1 else defs[0]
    defs[-1] = [2, 30, big]
    seasons_n = [[i * 250 + 20, i * 250 + 170] for i in range(6)]
    seasons_s = [[i * 250 + 150, i * 250 + 260] for i in range(6)]
    _mask_case(core, masks, defs.tolist(), seasons_n, seasons_s, rng.integers(0, 2, C))


@pytest.mark.parametrize("seed", range(8))
def test_run_filter_random_definitions(core, seed):
    # definitions that switch k_scan's run filter ON (every min_duration >= 2) with many (lmin, bmax) pairs, hot-day
    # densities from sparse to dense, a calendar whose day-of-year words are short (leap years / 360-day), both hemispheres
    from hdp_b200 import _tables as tb
    rng = np.random.default_rng(1000 + seed)
    cal = ["noleap", "standard", "360_day", "standard"][seed % 4]
    ax = tb.TimeAxis.date_range("2001-01-01", "2012-12-30" if cal == "360_day" else "2012-12-31", cal)
    T, C = len(ax), 97
    lmin, bmax = int(rng.integers(2, 13)), int(rng.integers(0, 10))
    D = int(rng.integers(1, 12))
    defs = np.stack([rng.integers(lmin, lmin + 6, D), rng.integers(0, bmax + 1, D), rng.integers(0, 4, D)], axis=1)
    defs[rng.integers(0, D)] = [lmin, bmax, int(rng.integers(0, 3))]
    dens = rng.uniform(0.02, 0.9, C)
    # AR(1)-like persistence so that long runs exist at every density
    z = rng.standard_normal((T, C))
    for t in range(1, T):
        z[t] = 0.8 * z[t - 1] + 0.6 * z[t]
    thr_q = np.array([np.quantile(z[:, c], 1 - dens[c]) for c in range(C)])
    masks = z > thr_q[None, :]
    masks[:, 0] = True                                        # one run over the whole series
    masks[:, 1] = False
    masks[:, 2] = (np.arange(T) % (lmin + bmax + 1)) < lmin   # exactly lmin hot days, then bmax + 1 cold: every run starts a heatwave
    masks[:, 3] = (np.arange(T) % (lmin + bmax)) < lmin - 1   # always one day short, breaks exactly bmax + 1 ... bmax
    st = tb.hemisphere_ranges(ax)
    x = masks.astype(np.float32)
    thr = np.full((C, int(ax.dayofyr.max()), 1), 0.5)
    args = (tb.doy_map(ax.dayofyr), defs.tolist(), st.north, st.south, rng.integers(0, 2, C).astype(np.uint8))
    out = ref_layout(core.metrics_array(dev(x), dev(thr), *args))
    want = oracle.metrics_batch(x, thr, *args)
    assert np.array_equal(out, want)


def test_scan_queue_pressure_with_unrelated_season_tables(core):
    # northern and southern tables whose seasons lie thousands of days apart, in one warp: lanes that wait for their next
    # season sit on full word queues while the others starve (k_scan's blocked path: close the waiting lanes' seasons
    # first), lanes run out of seasons long before the series ends (their queues are dropped), seasons of unequal counts
    rng = np.random.default_rng(321)
    T, C = 6000, 50
    masks = rng.random((T, C)) < rng.uniform(0.3, 0.9, C)[None, :]
    masks[:, 0] = True
    defs = [[3, 0, 0], [3, 1, 1], [2, 2, 3], [4, 1, 0], [6, 2, 1]]
    north = [[0, 90], [100, 130], [3000, 3200], [3200, 3210], [5900, 6000]]
    south = [[1400, 2100], [2500, 2600], [2600, 2601], [4000, 5000], [5990, 6000]]
    _mask_case(core, masks, defs, north, south, rng.integers(0, 2, C))
    _mask_case(core, masks, defs, north[:2] + [[150, 160], [170, 180], [190, 200]], south, rng.integers(0, 2, C))
    _mask_case(core, masks, defs, north, south, np.zeros(C, np.uint8))
    _mask_case(core, masks, defs, north, south, np.ones(C, np.uint8))


def test_metrics_tiny_and_empty(core):
    one = _mask_case(core, [[1]], [[1, 0, 0], [2, 0, 0]], [[0, 1]], [[0, 1]], None)
    assert one[0, :, 0, 0, 0].tolist() == [1, 0]
    x = torch.zeros((10, 0), dtype=torch.float32, device="cuda")
    thr = torch.zeros((0, 1, 2), dtype=torch.float64, device="cuda")
    out = core.metrics_array(x, thr, np.zeros(10, int), [[3, 0, 0]], [[0, 10]], [[0, 10]])
    assert tuple(out.shape) == (4, 2, 1, 1, 0)


# ------------------------------------------------------------------------------------------ limits
def test_maximum_percentiles_and_definitions(core):
    # HDP_B200_MAX_PERCENTILES = HDP_B200_MAX_DEFINITIONS = 32: the largest sweep one call takes, and the error beyond it
    from hdp_b200 import _tables as tb, _lib
    rng = np.random.default_rng(77)
    base_ax = tb.TimeAxis.date_range("1961-01-01", "1968-12-31", "noleap")
    run_ax = tb.TimeAxis.date_range("2001-01-01", "2006-12-31", "noleap")
    C = 45
    wt = tb.window_tables(base_ax.dayofyr, 7)
    xb = (12 + 9 * np.sin(2 * np.pi * (base_ax.dayofyr[:, None] - 110) / 365) + 3 * rng.standard_normal((len(base_ax), C))).astype(np.float32)
    xr_ = (13 + 9 * np.sin(2 * np.pi * (run_ax.dayofyr[:, None] - 110) / 365) + 3 * rng.standard_normal((len(run_ax), C))).astype(np.float32)
    q = np.linspace(0.5, 0.995, 32)
    thr = core.thresholds_array(dev(xb), wt, q)
    want_thr = oracle.thresholds_batch(xb, wt.window_samples(), q)
    assert bits_equal(thr.cpu().numpy(), want_thr)
    defs = [[int(a), int(b), int(c)] for a, b, c in zip(rng.integers(1, 7, 32), rng.integers(0, 3, 32), rng.integers(0, 3, 32))]
    st = tb.hemisphere_ranges(run_ax)
    args = (tb.doy_map(run_ax.dayofyr), defs, st.north, st.south, (rng.random(C) < 0.5).astype(np.uint8))
    out = core.metrics_array(dev(xr_), thr, *args)
    assert tuple(out.shape) == (4, 32, 32, st.n_years, C)
    assert np.array_equal(ref_layout(out), oracle.metrics_batch(xr_, want_thr, *args))
    with pytest.raises(_lib.HdpB200Error) as e:
        core.thresholds_array(dev(xb), wt, np.linspace(0.5, 0.99, 33))
    assert e.value.code == -2                                  # HDP_B200_ERR_UNSUPPORTED
    with pytest.raises(_lib.HdpB200Error) as e:
        core.metrics_array(dev(xr_), thr, args[0], defs + [[3, 0, 0]], *args[2:])
    assert e.value.code == -2


# ------------------------------------------------------------------------------------------ host pipeline
@pytest.mark.parametrize("chunk", ["32", "64", "96"])
def test_host_pipeline_many_chunks(core, monkeypatch, chunk):
    # the *_host entry points with cell chunks far smaller than the grid: more chunks than pipeline slots, a ragged last
    # chunk, tables uploaded with the first chunk only; both host layouts; must equal the device-resident path bit for bit
    from hdp_b200 import _tables as tb
    monkeypatch.setenv("HDP_B200_HOST_CHUNK_CELLS", chunk)
    rng = np.random.default_rng(23)
    base_ax = tb.TimeAxis.date_range("1961-01-01", "1966-12-31", "noleap")
    run_ax = tb.TimeAxis.date_range("2001-01-01", "2008-12-31", "noleap")
    wt = tb.window_tables(base_ax.dayofyr, 7)
    C = 299
    season = lambda ax: 15 + 10 * np.sin(2 * np.pi * (ax.dayofyr[:, None] - 110) / 365)
    xb = (season(base_ax) + 3 * rng.standard_normal((len(base_ax), C))).astype(np.float32)
    xr = (season(run_ax) + 3 * rng.standard_normal((len(run_ax), C)) + 2).astype(np.float32)
    q = np.arange(0.9, 1.0, 0.01)
    thr_d = core.thresholds_array(dev(xb), wt, q)
    thr_h = core.thresholds_host(xb, wt, q)
    assert bits_equal(thr_h, thr_d.cpu().numpy())
    assert bits_equal(core.thresholds_host(np.ascontiguousarray(xb.T).T, wt, q), thr_h)
    st = tb.hemisphere_ranges(run_ax)
    is_south = (rng.random(C) < 0.5).astype(np.uint8)
    args = (tb.doy_map(run_ax.dayofyr), [[3, 0, 0], [3, 1, 1], [4, 0, 0], [4, 1, 1], [5, 0, 0], [5, 1, 1]], st.north, st.south, is_south)
    out_d = core.metrics_array(dev(xr), thr_d, *args).cpu().numpy()
    assert np.array_equal(core.metrics_host(xr, thr_h, *args), out_d)
    assert np.array_equal(core.metrics_host(np.ascontiguousarray(xr.T).T, thr_h, *args), out_d)
    assert np.array_equal(core.metrics_host(xr, thr_h, *args[:-1], None), core.metrics_array(dev(xr), thr_d, *args[:-1], None).cpu().numpy())
    # thresholds that stay on the device between the two host calls (the compute_thresholds -> compute_group_metrics chain)
    import torch
    thr_k = core.thresholds_host(xb, wt, q, keep=True)
    assert bits_equal(thr_k, thr_h) and not thr_k.flags.writeable
    d_thr = core.resident_thresholds(thr_k)
    assert d_thr is not None and torch.equal(d_thr, thr_d)
    assert core.resident_thresholds(thr_k.reshape(C, -1, q.size)) is d_thr         # a same-memory view still hits
    assert core.resident_thresholds(thr_k.copy()) is None and core.resident_thresholds(thr_h) is None
    assert np.array_equal(core.metrics_host(xr, thr_k, *args), out_d)              # resident copy used (no upload)
    assert np.array_equal(core.metrics_host(xr, thr_d, *args), out_d)              # explicit device tensor
    keep = torch.empty_like(thr_d)
    core.thresholds_host(xb, wt, q, out=np.empty_like(thr_h), keep=keep)
    assert torch.equal(keep, thr_d)
    core.host_release()
    assert core.resident_thresholds(thr_k) is None
    assert bits_equal(core.thresholds_host(xb, wt, q), thr_h)          # the context comes back after a release


# ------------------------------------------------------------------------------------------ hdp.measure pre-pass (next row)
def _bits32(a):
    """float32 bit patterns with every NaN mapped to one pattern (CUDA's single-precision operations return the canonical
    NaN 0x7fffffff where x86 propagates the operand's payload; which NaN it is carries no meaning)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.where(np.isnan(a), np.uint32(0x7fc00000), a.view(np.uint32))


def test_heat_index_kernel_matches_reference_kernel(core):
    # outputs of the reference's own Numba heat_index (tests/golden/measure.npz), bit for bit in float32
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "measure.npz"))
    got = core.heat_index_array(dev(g["t"].astype(np.float32).ravel()), dev(g["rh"].astype(np.float32).ravel())).cpu().numpy()
    assert np.array_equal(_bits32(got), _bits32(g["hi"].ravel()))


def test_measure_prepass_matches_host_mirror(core):
    # dense sweep over every branch of the regression (simple formula, full regression, both adjustments, NaN / inf),
    # odd length (scalar tail + misaligned views), against the NumPy mirror that is itself pinned to the reference kernel
    from hdp_b200 import measure
    rng = np.random.default_rng(41)
    n = 1_000_003
    tf = rng.uniform(20, 125, n).astype(np.float32)
    rh = rng.uniform(0, 100, n).astype(np.float32)
    tf[:7] = [np.nan, np.inf, -np.inf, 80.0, 87.0, 112.0, 95.0]
    rh[:7] = [50.0, 50.0, 50.0, 13.0, 85.0, 12.999, np.nan]
    want = measure.heat_index(tf, rh)
    got = core.heat_index_array(dev(tf), dev(rh)).cpu().numpy()
    assert np.array_equal(_bits32(got), _bits32(want))
    # misaligned (offset by one element): the scalar path
    d_t, d_r = dev(tf), dev(rh)
    got1 = core.heat_index_array(d_t[1:].contiguous()[:], d_r[1:].contiguous()).cpu().numpy()
    assert np.array_equal(_bits32(got1), _bits32(want[1:]))
    # the `{name}_hi` chain of format_standard_measures: C -> F, heat index, F -> C, rh in % and in g/g
    tc = rng.uniform(-10, 50, n).astype(np.float32)
    for frac in (False, True):
        r = (rh / np.float32(100)).astype(np.float32) if frac else rh
        pct = r * np.float32(100) if frac else r
        chain = (measure.heat_index((tc * 1.8) + 32, pct) - 32) / 1.8
        assert chain.dtype == np.float32
        got = core.heat_index_measure_array(dev(tc), dev(r), rh_is_fraction=frac).cpu().numpy()
        assert np.array_equal(_bits32(got), _bits32(chain))
    # unit conversions (float32 array arithmetic in the reference)
    k = (tc + np.float32(273.15)).astype(np.float32)
    want_k = k.copy(); want_k -= 273.15
    assert np.array_equal(_bits32(core.to_celsius_array(dev(k), "K").cpu().numpy()), _bits32(want_k))
    f = ((tc * 1.8) + 32).astype(np.float32)
    assert np.array_equal(_bits32(core.to_celsius_array(dev(f), "degF").cpu().numpy()), _bits32((f - 32) / 1.8))
    assert np.array_equal(_bits32(core.to_celsius_array(dev(tc), "degC").cpu().numpy()), _bits32(tc))
    assert core.heat_index_array(dev(tf[:0]), dev(rh[:0])).numel() == 0


def test_measure_prepass_full_size_throughput(core, capsys):
    # one measure of cmip6_1deg (64 800 x 31 390 float32): in-place chain, result spot-checked; prints the achieved GB/s
    n = 64800 * 31390
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    tc = torch.empty(n, dtype=torch.float32, device="cuda").uniform_(-10, 50, generator=g)
    rh = torch.empty(n, dtype=torch.float32, device="cuda").uniform_(0, 100, generator=g)
    out = torch.empty_like(tc)
    core.heat_index_measure_array(tc, rh, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        core.heat_index_measure_array(tc, rh, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    from hdp_b200 import measure
    idx = torch.randint(0, n, (200_000,), device="cuda")
    t_s, r_s = tc[idx].cpu().numpy(), rh[idx].cpu().numpy()
    want = (measure.heat_index((t_s * 1.8) + 32, r_s) - 32) / 1.8
    assert np.array_equal(_bits32(out[idx].cpu().numpy()), _bits32(want))
    with capsys.disabled():
        print(f"\n[k_measure] {n} elements, {ms:.2f} ms, {12 * n / ms / 1e6:.0f} GB/s (12 B per element)")


def test_host_pipeline_pageable_memory_many_ring_blocks(core):
    # ordinary (pageable) NumPy arrays large enough that every copy runs through several 32 MB blocks of the pinned staging
    # rings: pitched sample rows, threshold chunks that are one contiguous row longer than a block, pitched result rows;
    # both host layouts; against the device-resident path bit for bit
    from hdp_b200 import _tables as tb
    rng = np.random.default_rng(77)
    base_ax = tb.TimeAxis.date_range("1961-01-01", "1964-12-31", "noleap")
    run_ax = tb.TimeAxis.date_range("2001-01-01", "2007-12-31", "noleap")
    wt = tb.window_tables(base_ax.dayofyr, 7)
    C = 40001
    season = lambda ax: 15 + 10 * np.sin(2 * np.pi * (ax.dayofyr[:, None] - 110) / 365)
    xb = (season(base_ax) + 3 * rng.standard_normal((len(base_ax), C), dtype=np.float32)).astype(np.float32)
    xr_ = (season(run_ax) + 2 + 3 * rng.standard_normal((len(run_ax), C), dtype=np.float32)).astype(np.float32)
    q = np.array([0.5, 0.8, 0.9, 0.95, 0.975, 0.99])
    thr_d = core.thresholds_array(dev(xb), wt, q)
    thr_h = core.thresholds_host(xb, wt, q)                                  # 58 MB in, 700 MB out
    assert bits_equal(thr_h, thr_d.cpu().numpy())
    st = tb.hemisphere_ranges(run_ax)
    defs = [[a, b, c] for a in (3, 4) for b in (0, 1) for c in (0, 1)]
    is_south = (rng.random(C) < 0.5).astype(np.uint8)
    args = (tb.doy_map(run_ax.dayofyr), defs, st.north, st.south, is_south)
    out_d = core.metrics_array(dev(xr_), thr_d, *args).cpu().numpy()
    assert np.array_equal(core.metrics_host(xr_, thr_h, *args), out_d)       # 409 MB + 700 MB in, 107 MB out
    xr_t = np.ascontiguousarray(xr_.T).T                                     # time-contiguous host layout
    assert np.array_equal(core.metrics_host(xr_t, thr_h, *args), out_d)
    core.host_release()


# ------------------------------------------------------------------------------------------ declared options, reductions, file IO
def test_no_season_and_fixed_value_variants(core):
    # the options the reference declares but does not implement (hdp/threshold.py:105-110) as explicit variants on the same
    # kernels: one window that pools the whole baseline / one constant threshold, with a day-of-year axis of length 1
    from hdp_b200 import _tables as tb
    rng = np.random.default_rng(3)
    ax = tb.TimeAxis.date_range("1961-01-01", "1968-12-31", "noleap")
    run_ax = tb.TimeAxis.date_range("2001-01-01", "2005-12-31", "noleap")
    C = 45
    x = (15 + 8 * np.sin(2 * np.pi * ax.dayofyr[:, None] / 365) + 3 * rng.standard_normal((len(ax), C))).astype(np.float32)
    xr_ = (17 + 8 * np.sin(2 * np.pi * run_ax.dayofyr[:, None] / 365) + 3 * rng.standard_normal((len(run_ax), C))).astype(np.float32)
    q = np.array([0.5, 0.9, 0.99])
    thr = core.thresholds_no_season_array(dev(x), q)
    assert tuple(thr.shape) == (C, 1, 3)
    want = oracle.thresholds_batch(x, tb.no_season_tables(len(ax)).window_samples(), q)
    assert bits_equal(thr.cpu().numpy(), want)
    st = tb.hemisphere_ranges(run_ax)
    zeros = np.zeros(len(run_ax), np.int64)
    args = ([[3, 0, 0], [3, 1, 1]], st.north, st.south, (np.arange(C) % 2).astype(np.uint8))
    out = core.metrics_array(dev(xr_), thr, zeros, *args)
    assert np.array_equal(ref_layout(out), oracle.metrics_batch(xr_, want, zeros, *args))
    fixed = core.fixed_thresholds(C, 24.5)
    out = core.metrics_array(dev(xr_), fixed, zeros, *args)
    assert np.array_equal(ref_layout(out), oracle.metrics_batch(xr_, np.full((C, 1, 1), 24.5), zeros, *args))


def test_weighted_spatial_mean(core):
    # compute_weighted_spatial_mean of the reference's figure deck (hdp/graphics/figure.py:14-15): cos(lat)-weighted mean over cells
    rng = np.random.default_rng(8)
    C = 180 * 36
    lat = np.repeat(-89.5 + np.arange(180), 36)
    w = np.cos(np.deg2rad(lat))
    m = torch.as_tensor(rng.integers(0, 154, (4, 3, 2, 7, C)).astype(np.uint16)).cuda()
    got = core.weighted_spatial_mean(m, w).cpu().numpy()
    want = (m.cpu().numpy().astype(np.float64) * w).sum(-1) / w.sum()
    assert got.shape == (4, 3, 2, 7)
    assert np.allclose(got, want, rtol=1e-12, atol=0)              # float64 sums in a different (fixed) order: tolerance, not bits


def test_streaming_io_npy(core, tmp_path):
    # path in / path out (reference compute_threshold_io / compute_metrics_io): memory-mapped .npy files streamed through the host
    # pipeline; equals the in-memory path bit for bit; refuses to overwrite like the reference
    from hdp_b200 import _tables as tb, io as hio
    rng = np.random.default_rng(4)
    ax = tb.TimeAxis.date_range("1961-01-01", "1972-12-31", "noleap")
    run_ax = tb.TimeAxis.date_range("2001-01-01", "2004-12-31", "noleap")
    C = 130
    lat = np.linspace(-60, 60, C)
    xb = (288 + 8 * np.sin(2 * np.pi * ax.dayofyr[:, None] / 365) + 3 * rng.standard_normal((len(ax), C))).astype(np.float32)     # Kelvin
    xr_ = (290 + 8 * np.sin(2 * np.pi * run_ax.dayofyr[:, None] / 365) + 3 * rng.standard_normal((len(run_ax), C))).astype(np.float32)
    np.save(tmp_path / "base.npy", xb); np.save(tmp_path / "run.npy", xr_)
    q = np.arange(0.9, 1.0, 0.02)
    defs = [[3, 0, 0], [4, 1, 1]]
    hio.compute_threshold_io(str(tmp_path / "base.npy"), ax, str(tmp_path / "thr.npy"), q, units="degK")
    with pytest.raises(FileExistsError):
        hio.compute_threshold_io(str(tmp_path / "base.npy"), ax, str(tmp_path / "thr.npy"), q, units="degK")
    hio.compute_metrics_io(str(tmp_path / "run.npy"), run_ax, str(tmp_path / "thr.npy"), lat, str(tmp_path / "met.npy"), defs, units="degK")
    thr = core.thresholds_array(dev(xb), tb.window_tables(ax.dayofyr, 7), q, units="degK")
    assert bits_equal(np.load(tmp_path / "thr.npy"), thr.cpu().numpy())
    st = tb.hemisphere_ranges(run_ax)
    out = core.metrics_array(dev(xr_), thr, tb.doy_map(run_ax.dayofyr), defs, st.north, st.south, tb.is_south(lat), units="degK")
    assert np.array_equal(np.load(tmp_path / "met.npy"), out.cpu().numpy())
