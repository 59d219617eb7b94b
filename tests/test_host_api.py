"""Host-side mirror of the reference API (hdp_b200.measure / .threshold / .metric / .utils).

CPU part: everything that does not need the CUDA library (measure arithmetic against outputs of the reference's
own Numba ``heat_index``, generators, layout helpers, the labelled-array stand-in).  GPU part (``-m gpu``): the
reference's end-to-end workflow test (hdp/tests/test_workflow.py) replayed through the drop-in functions and
checked value by value against the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits_equal
from hdp_b200 import _layout, _tables as tb, measure, utils, xr


# ------------------------------------------------------------------------------------------ CPU
def test_heat_index_matches_reference_kernel():
    g = np.load(os.path.join(GOLDEN, "measure.npz"))
    got = measure.heat_index(g["t"], g["rh"])
    assert got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), g["hi"].view(np.uint32))       # bit-exact float32


def test_generators_like_reference_test_utils():
    # hdp/tests/test_utils.py:6-56
    ctl = utils.generate_test_control_dataarray(grid_shape=(2, 3))
    assert tuple(ctl.dims) == ("lon", "lat", "time") and ctl.shape == (2, 3, 50 * 365)
    assert ctl.attrs["units"] == "degC"
    t = utils.time_axis_of(ctl)
    assert t.calendar == "noleap" and len(t) == 50 * 365
    v = xr.values_of(ctl)
    slope = np.polyfit(np.arange(v.shape[-1]), v[0, 0], 1)[0]
    assert abs(slope) < 0.01                                         # hdp/tests/test_utils.py:25
    warm = xr.values_of(utils.generate_test_warming_dataarray(grid_shape=(2, 3)))
    assert np.polyfit(np.arange(warm.shape[-1]), warm[0, 0], 1)[0] > 0
    rh = xr.values_of(utils.generate_test_rh_dataarray(grid_shape=(2, 3)))
    assert rh.min() >= 0 and rh.max() <= 1


def test_format_standard_measures_units_and_attrs():
    ctl = utils.generate_test_control_dataarray(add_noise=True)
    kel = ctl.copy()
    kel.values = xr.values_of(ctl) + 273.15
    kel.attrs["units"] = "K"
    kel.name = "tas_k"
    ds = measure.format_standard_measures([ctl, kel])
    assert list(ds.keys()) == ["test_temperature_data", "tas_k"]
    a, b = ds["test_temperature_data"], ds["tas_k"]
    assert a.dtype == np.float32 and b.dtype == np.float32
    assert b.attrs["units"] == "degC" and b.attrs["hdp_type"] == "measure"
    assert b.attrs["input_variable"] == "tas_k" and b.attrs["baseline_variable"] == "tas_k"
    # the conversion happens in float32, like the reference's in-place `temp -= 273.15` on a float32 array
    want = (xr.values_of(ctl) + 273.15).astype(np.float32)
    want -= 273.15
    assert np.array_equal(xr.values_of(b), want)
    assert "hdp_version" in ds.attrs and "history" in ds.attrs
    with pytest.raises(AssertionError):
        bad = ctl.copy()
        bad.attrs["units"] = "furlongs"
        measure.format_standard_measures([bad])


def test_format_standard_measures_with_rh_adds_heat_index():
    ctl = utils.generate_test_warming_dataarray(add_noise=True)
    rh = utils.generate_test_rh_dataarray()
    ds = measure.format_standard_measures([ctl], rh=rh)
    assert list(ds.keys()) == ["test_temperature_data", "test_temperature_data_hi"]
    hi = ds["test_temperature_data_hi"]
    assert hi.attrs["baseline_variable"] == "test_temperature_data_hi" and hi.attrs["units"] == "degC"
    # same chain as the reference: C -> F (f32), heat_index(f32, f32 %), F -> C
    f = xr.values_of(ds["test_temperature_data"]) * 1.8 + 32
    want = (measure.heat_index(f.astype(np.float32), (xr.values_of(rh).astype(np.float32) * np.float32(100))) - 32) / 1.8
    assert np.array_equal(xr.values_of(hi), want)


def test_factor_window_samples_round_trip():
    for cal, days, r in (("noleap", 4 * 365, 7), ("standard", 1461 + 365, 15), ("360_day", 720, 2), ("noleap", 730, 0)):
        ax = tb.TimeAxis.daily((1999, 1, 1), days, cal)
        wt = tb.window_tables(ax.dayofyr, r)
        back = tb.factor_window_samples(wt.window_samples())
        assert np.array_equal(back.window_samples(), wt.window_samples())
        assert back.width == wt.width and back.n_y == wt.n_y
    with pytest.raises(ValueError):
        tb.factor_window_samples(np.array([[0, 1], [0, 1]]))             # duplicate rows: not a doy grouping


def test_to_time_cells_layouts():
    rng = np.random.default_rng(0)
    v = rng.standard_normal((3, 4, 11))
    x, dims, shape = _layout.to_time_cells(v, ("lon", "lat", "time"))
    assert x.shape == (11, 12) and dims == ["lon", "lat"] and shape == [3, 4] and x.dtype == np.float32
    assert x.strides[0] == 4                                         # time-contiguous view, no transposition on the host
    assert np.array_equal(x[:, 5], v[1, 1].astype(np.float32))
    x2, dims2, _ = _layout.to_time_cells(np.moveaxis(v, -1, 0), ("time", "lon", "lat"))
    assert x2.strides[1] == 4 and np.array_equal(x2, x) and dims2 == ["lon", "lat"]
    x3, dims3, shape3 = _layout.to_time_cells(np.moveaxis(v, -1, 1), ("lon", "time", "lat"))
    assert np.array_equal(x3, x) and dims3 == ["lon", "lat"] and shape3 == [3, 4]
    lat = _layout.cell_latitudes(np.array([-45.0, -1.0, 0.0, 30.0]), ["lon", "lat"], [3, 4])
    assert lat.tolist() == [-45.0, -1.0, 0.0, 30.0] * 3
    assert tb.is_south(lat).tolist() == [1, 1, 0, 0] * 3             # equator counts as North (metric.py:247-252)


def test_mini_merge_exact_join():
    a = xr.MiniDataArray(np.zeros((2, 3)), dims=["x", "y"], coords={"x": [0, 1], "y": [0, 1, 2]}, name="a")
    b = xr.MiniDataArray(np.ones((2, 3)), dims=["x", "y"], coords={"x": [0, 1], "y": [0, 1, 2]}, name="b")
    ds = xr.mini_merge([a, b])
    assert list(ds.keys()) == ["a", "b"] and len(ds) == 2
    c = xr.MiniDataArray(np.ones((2, 3)), dims=["x", "y"], coords={"x": [0, 2], "y": [0, 1, 2]}, name="c")
    with pytest.raises(ValueError):
        xr.mini_merge([a, c])


def test_mirror_exports_every_function_of_the_reference_modules():
    """Every public function of hdp/{threshold,metric,measure,utils}.py exists under the same name (reference line in brackets)."""
    from hdp_b200 import metric, threshold
    surface = {
        threshold: ["datetimes_to_windows", "compute_percentiles", "compute_percentiles_wrapper", "compute_threshold",       # :12 :59 :81 :96
                    "compute_thresholds", "compute_threshold_io"],                                                          # :207 :232
        metric: ["index_heatwaves", "heatwave_number", "heatwave_frequency", "heatwave_duration", "heatwave_average",        # :12 :64 :86 :106 :141
                 "get_range_indices", "compute_hemisphere_ranges", "build_doy_map", "indicate_hot_days",                     # :175 :212 :265 :281
                 "compute_heatwave_metrics", "compute_heatwave_metrics_wrapper", "compute_individual_metrics",               # :305 :344 :372
                 "compute_group_metrics", "compute_metrics_io"],                                                            # :509 :526
        measure: ["kelvin_to_celsius", "fahrenheit_to_celsius", "celsius_to_fahrenheit", "heat_index", "heat_index_map_wrapper",   # :10 :27 :44 :62 :97
                  "apply_heat_index", "convert_temp_units", "format_standard_measures"],                                    # :111 :136 :152
        utils: ["get_time_stamp", "add_history", "get_version", "get_func_description", "generate_test_warming_dataarray",   # :10 :14 :23 :27 :39
                "generate_test_rh_dataarray", "generate_test_control_dataarray"],                                           # :45 :53
    }
    for module, names in surface.items():
        for name in names:
            assert callable(getattr(module, name, None)), f"{module.__name__}.{name}"


def test_io_wrappers_keep_the_reference_errors(tmp_path):
    """hdp/threshold.py:264-277, hdp/metric.py:566-576: FileExistsError / ValueError before anything is read."""
    from hdp_b200 import metric, threshold
    existing = tmp_path / "there.nc"
    existing.write_bytes(b"x")
    with pytest.raises(FileExistsError):
        threshold.compute_threshold_io("in.nc", "tas", str(existing), [0.9])
    with pytest.raises(FileExistsError):
        threshold.compute_threshold_io("in.nc", "tas", str(tmp_path / "missing_dir" / "out.nc"), [0.9])
    with pytest.raises(ValueError):
        threshold.compute_threshold_io("in.nc", "tas", str(tmp_path / "out.txt"), [0.9])
    with pytest.raises(FileExistsError):
        metric.compute_metrics_io(str(existing), "m.nc", "tas", "t.nc", [[3, 0, 0]])
    with pytest.raises(ValueError):
        metric.compute_metrics_io(str(tmp_path / "out.csv"), "m.nc", "tas", "t.nc", [[3, 0, 0]])
    if not xr.HAVE_XARRAY:
        with pytest.raises(RuntimeError, match="hdp_b200.io"):
            threshold.compute_threshold_io("in.nc", "tas", str(tmp_path / "out.nc"), [0.9])


def test_get_func_description_and_heat_index_map_wrapper():
    def documented():
        """First line.

        Second line,
        continued.

        :param x: not part of the description
        """
    assert utils.get_func_description(documented) == "First line. Second line, continued. "       # hdp/utils.py:27-36
    g = np.load(os.path.join(GOLDEN, "measure.npz"))
    t, rh = np.asarray(g["t"]).ravel()[:600].reshape(20, 30), np.asarray(g["rh"]).ravel()[:600].reshape(20, 30)
    coords = {"lat": np.arange(20.0), "lon": np.arange(30.0)}
    ds = xr.Dataset({"temp": xr.DataArray(t.astype(np.float64), dims=["lat", "lon"], coords=coords, name="temp", attrs={"units": "degF"}),
                     "rh": xr.DataArray(rh, dims=["lat", "lon"], coords=coords, name="rh", attrs={"units": "%"})})
    hi = measure.heat_index_map_wrapper(ds)                                                        # hdp/measure.py:97-108
    assert tuple(hi.dims) == ("lat", "lon") and hi.dtype == np.float32
    assert np.array_equal(xr.values_of(hi).ravel().view(np.uint32), np.asarray(g["hi"]).ravel()[:600].view(np.uint32))


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_reference_workflow_through_drop_in_api():
    """hdp/tests/test_workflow.py:15-63 (2x3 grid, RH -> two measures) through hdp_b200, values against the oracle."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import oracle
    from hdp_b200 import metric, threshold

    grid = (2, 3)
    base = utils.generate_test_control_dataarray(grid_shape=grid, add_noise=True)
    test = utils.generate_test_warming_dataarray(grid_shape=grid, add_noise=True)
    rh = utils.generate_test_rh_dataarray(grid_shape=grid)
    base_measures = measure.format_standard_measures([base], rh=rh)
    test_measures = measure.format_standard_measures([test], rh=rh)
    percentiles = np.arange(0.9, 1.0, 0.01)
    defs = [[3, 0, 0], [3, 1, 1], [4, 0, 0], [4, 1, 1], [5, 0, 0], [5, 1, 1]]

    thresholds = threshold.compute_thresholds(base_measures, percentiles)
    assert len(thresholds.data_vars) == 2
    assert np.array_equal(np.asarray(xr.coord_values(thresholds, "percentile")), percentiles)    # bit-for-bit, :39
    thr_da = thresholds["test_temperature_data_threshold"]
    assert tuple(thr_da.dims) == ("lon", "lat", "doy", "percentile") and thr_da.dtype == np.float64
    assert thr_da.attrs["hdp_type"] == "threshold" and thr_da.attrs["baseline_calendar"] == "noleap"
    assert thr_da.attrs["param_rolling_window_size"] == "7"

    metrics = metric.compute_group_metrics(test_measures, thresholds, defs)
    names = list(metrics.keys())
    assert len(names) == 8 and "test_temperature_data.test_temperature_data_threshold.HWF" in names
    assert metrics.attrs["variable_naming_delimeter"] == "."
    units = {"HWF": "heatwave days", "HWD": "heatwave days", "HWN": "heatwave events", "HWA": "heatwave events"}
    for name in names:
        da = metrics[name]
        assert tuple(da.dims) == ("percentile", "definition", "lon", "lat", "time")               # :56
        assert da.shape == (10, 6, 2, 3, 50) and da.dtype == np.int64                             # :57
        assert da.attrs["units"] == units[name.rsplit(".", 1)[1]]
    assert list(np.asarray(xr.coord_values(metrics, "definition"))) == ["3-0-0", "3-1-1", "4-0-0", "4-1-1", "5-0-0", "5-1-1"]

    # values: the oracle on the same measures
    for mname in ("test_temperature_data", "test_temperature_data_hi"):
        b = xr.values_of(base_measures[mname]).reshape(6, -1).T
        r = xr.values_of(test_measures[mname]).reshape(6, -1).T
        t_axis = utils.time_axis_of(test_measures[mname])
        wt = tb.window_tables(utils.time_axis_of(base_measures[mname]).dayofyr, 7)
        thr_ref = oracle.thresholds_batch(np.ascontiguousarray(b), wt.window_samples(), percentiles)
        assert bits_equal(xr.values_of(thresholds[f"{mname}_threshold"]).reshape(6, 365, 10), thr_ref)
        st = tb.hemisphere_ranges(t_axis)
        south = tb.is_south(np.tile(np.linspace(-90, 90, 3), 2))
        want = oracle.metrics_batch(np.ascontiguousarray(r), thr_ref, tb.doy_map(t_axis.dayofyr), defs, st.north, st.south, south)
        for i, short in enumerate(("HWF", "HWN", "HWD", "HWA")):
            got = xr.values_of(metrics[f"{mname}.{mname}_threshold.{short}"]).reshape(10, 6, 6, 50)
            assert np.array_equal(got, want[:, :, :, i, :])
    hwf = xr.values_of(metrics["test_temperature_data.test_temperature_data_threshold.HWF"]).mean()
    hwd = xr.values_of(metrics["test_temperature_data.test_temperature_data_threshold.HWD"]).mean()
    hwa = xr.values_of(metrics["test_temperature_data.test_temperature_data_threshold.HWA"]).mean()
    assert hwf >= hwd >= hwa                                                                      # :52-53


@pytest.mark.gpu
def test_array_level_seams_match_reference_signatures():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import oracle
    from hdp_b200 import metric, threshold
    rng = np.random.default_rng(2)
    ax = tb.TimeAxis.daily((1990, 1, 1), 4 * 365, "noleap")
    temps = (10 + 5 * rng.standard_normal((2, 3, len(ax)))).astype(np.float32)
    win = threshold.datetimes_to_windows(ax, 3)
    q = np.array([0.5, 0.9])
    got = threshold.compute_percentiles(temps, win, q)                   # '(t),(d,b),(p)->(d,p)' with leading dims
    assert got.shape == (2, 3, 365, 2)
    for i in range(2):
        for j in range(3):
            assert bits_equal(got[i, j], oracle.compute_percentiles(temps[i, j], win, q))
    dm = metric.build_doy_map(ax)
    seasons = metric.get_range_indices(ax, (5, 1), (10, 1))
    out = metric.compute_heatwave_metrics(temps[0, 0], got[0, 0, :, 0], dm, 3, 1, 1, seasons)
    assert out.shape == (4, 4) and out.dtype == np.int64
    assert np.array_equal(out, oracle.compute_heatwave_metrics(temps[0, 0], got[0, 0, :, 0], dm, 3, 1, 1, seasons))


@pytest.mark.gpu
@pytest.mark.parametrize("members,years", [(3, 4), (40, 20)])
def test_member_dimension_is_pooled_into_the_sample(members, years):
    """hdp/threshold.py:114-119: a `member` dimension is concatenated along time, so every window pools
    W x years x members samples (CESM2-LENS style).  (3, 4): the segment kernel; (40, 20): 15 x 800 = 12 000 samples per
    window, which only the generic gather + sort kernel takes."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import oracle
    from hdp_b200 import threshold
    rng = np.random.default_rng(members)
    ax = tb.TimeAxis.date_range("1961-01-01", f"{1960 + years}-12-31", "noleap")
    lon, lat = np.array([0.0, 120.0]), np.array([-30.0, 0.0, 30.0])
    vals = (15 + 8 * np.sin(2 * np.pi * (ax.dayofyr - 110) / 365)[None, None, None, :]
            + 3 * rng.standard_normal((members, 2, 3, len(ax)))).astype(np.float32)
    da = xr.DataArray(vals, dims=["member", "lon", "lat", "time"],
                      coords={"member": np.arange(members), "lon": lon, "lat": lat, "time": ax},
                      name="tas", attrs={"units": "degC", "hdp_type": "measure", "baseline_variable": "tas"})
    q = np.array([0.5, 0.9, 0.95, 0.99])
    thr = threshold.compute_threshold(da, q)["tas_threshold"]
    assert tuple(thr.dims) == ("lon", "lat", "doy", "percentile") and thr.shape == (2, 3, 365, 4)
    pooled = np.concatenate([vals[m] for m in range(members)], axis=-1).reshape(6, -1).T     # members one after another along time
    wt = tb.window_tables(np.tile(ax.dayofyr, members), 7)
    assert wt.n_y == members * years
    want = oracle.thresholds_batch(np.ascontiguousarray(pooled), wt.window_samples(), q)
    assert bits_equal(xr.values_of(thr).reshape(6, 365, 4), want)
