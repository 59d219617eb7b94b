"""GPU parity at the FULL sizes of BASELINE.json's configs: size-independent properties over every cell, plus a bit-exact
comparison with the CPU oracle on cells sampled across the whole grid (the oracle finishes those in well under a second).

    cmip6_1deg    180 x 360, 30-year baseline + 86-year run, 10 percentiles x 6 definitions        (configs[1], the bench)
    lens_member   192 x 288, one CESM2-LENS-like member                                            (configs[2], per GPU)
    era5_025deg   0.25 deg, 30-year baseline, 31-day window, standard calendar: the whole 721 x 1440 grid     (configs[3])
    wide_sweep    1 deg, 20 percentiles x 24 definitions                                           (configs[4])
"""
import numpy as np
import pytest

import oracle
from conftest import bits_equal

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def core():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hdp_b200 import _core
    yield _core
    _core.release_workspaces()
    torch.cuda.empty_cache()


def _fields(wl, lat, seed, run=True):
    from hdp_b200 import synth
    base = synth.gridded_field(lat, wl.base_axis().dayofyr, seed=seed, device="cuda")
    warm = synth.gridded_field(lat, wl.run_axis().dayofyr, seed=seed + 1, trend=4.0, device="cuda") if run and wl.run_years else None
    return base, warm


def _check_threshold_properties(base, thr, q):
    # quantiles of one sample set are non-decreasing in q and lie inside the range of the series
    assert not torch.isnan(thr).any()
    order = np.argsort(q)
    t = thr[:, :, torch.as_tensor(order, device=thr.device)]
    assert bool((t[:, :, 1:] >= t[:, :, :-1]).all())
    lo, hi = base.amin(dim=0).double(), base.amax(dim=0).double()
    assert bool((thr >= lo[:, None, None]).all()) and bool((thr <= hi[:, None, None]).all())


def _check_metric_properties(out, defs, seasons_len_max):
    hwf, hwn, hwd, hwa = (out[i].to(torch.int32) for i in range(4))
    assert bool((hwf >= hwd).all()) and bool((hwd >= hwa).all())          # reference invariant, hdp/tests/test_workflow.py:52-53
    assert bool((hwn <= hwf).all()) and bool((hwf <= seasons_len_max).all())
    assert bool(((hwn == 0) == (hwf == 0)).all())
    assert bool((hwa == torch.where(hwn > 0, hwf // hwn.clamp(min=1), torch.zeros_like(hwf))).all())   # trunc(mean), metric.py:340
    # definitions without breaks: a higher percentile only removes hot days, so no season gains heatwave days
    plain = [i for i, d in enumerate(defs) if d[1] == 0 and d[2] == 0]
    if plain:
        f = hwf[:, plain]
        assert bool((f[1:] <= f[:-1]).all())


def _oracle_sample(core, wl, lat, base, warm, thr, out, n_sample, seed=0):
    from hdp_b200 import _tables as tb
    C = lat.size
    sel = np.unique(np.concatenate([[0, C - 1], np.random.default_rng(seed).integers(0, C, n_sample)]))
    sel_t = torch.as_tensor(sel, device="cuda")
    wt = wl.window_tables()
    b = base[:, sel_t].cpu().numpy()
    thr_ref = oracle.thresholds_batch(b, wt.window_samples(), wl.percentiles)
    assert bits_equal(thr[sel_t].cpu().numpy(), thr_ref)
    if warm is not None:
        st = wl.seasons()
        dm = tb.doy_map(wl.run_axis().dayofyr)
        south = (lat[sel] < 0).astype(np.uint8)
        want = oracle.metrics_batch(warm[:, sel_t].cpu().numpy(), thr_ref, dm, wl.defs, st.north, st.south, south)
        got = out.view(torch.int16)[..., sel_t].cpu().numpy().view(np.uint16).astype(np.int64).transpose(1, 2, 4, 0, 3)
        assert np.array_equal(got, want)


def _run(core, wl, lat, seed, n_sample):
    from hdp_b200 import _tables as tb
    base, warm = _fields(wl, lat, seed)
    wt = wl.window_tables()
    thr = core.thresholds_array(base, wt, wl.percentiles)
    _check_threshold_properties(base, thr, np.asarray(wl.percentiles))
    out = None
    if warm is not None:
        st = wl.seasons()
        south = (lat < 0).astype(np.uint8)
        out = core.metrics_array(warm, thr, tb.doy_map(wl.run_axis().dayofyr), wl.defs, st.north, st.south, south)
        longest = int(max((np.asarray(st.north)[:, 1] - np.asarray(st.north)[:, 0]).max(),
                          (np.asarray(st.south)[:, 1] - np.asarray(st.south)[:, 0]).max()))
        _check_metric_properties(out, wl.defs, longest)
    _oracle_sample(core, wl, lat, base, warm, thr, out, n_sample, seed)
    return base, warm, thr, out


def test_cmip6_1deg_full_grid(core):
    from hdp_b200 import synth, workloads
    wl = workloads.get("cmip6_1deg")
    lat, _ = synth.grid_latitudes(wl.n_lat, wl.n_lon)
    base, warm, thr, out = _run(core, wl, lat, 4321, 96)
    # whole years of a noleap baseline can be permuted without changing any window's sample set: thresholds are bit-identical
    n_y, T_b, C = wl.base_years, base.shape[0], base.shape[1]
    perm = torch.as_tensor(np.random.default_rng(1).permutation(n_y), device="cuda")
    shuffled = base.view(n_y, T_b // n_y, C)[perm].reshape(T_b, C)
    thr2 = core.thresholds_array(shuffled, wl.window_tables(), wl.percentiles)
    assert torch.equal(thr2.view(torch.int64), thr.view(torch.int64))
    # a threshold nothing exceeds: no heatwave anywhere; one everything exceeds: every season is one heatwave of its full length
    from hdp_b200 import _tables as tb
    st, dm = wl.seasons(), tb.doy_map(wl.run_axis().dayofyr)
    south = (lat < 0).astype(np.uint8)
    sub = slice(0, 4096)
    never = torch.full((4096, thr.shape[1], 1), float("inf"), dtype=torch.float64, device="cuda")
    z = core.metrics_array(warm[:, sub], never, dm, wl.defs, st.north, st.south, south[sub])
    assert int(z.to(torch.int32).abs().sum()) == 0
    always = -never
    a = core.metrics_array(warm[:, sub], always, dm, wl.defs, st.north, st.south, south[sub]).to(torch.int32)
    T = warm.shape[0]
    for tab, mask in ((np.asarray(st.north), south[sub] == 0), (np.asarray(st.south), south[sub] == 1)):
        if mask.any():
            length = np.clip(np.minimum(tab[:, 1], T) - np.clip(tab[:, 0], 0, T), 0, None)
            m = torch.as_tensor(np.nonzero(mask)[0], device="cuda")
            want = torch.as_tensor(length, device="cuda", dtype=torch.int32)[None, None, :, None]
            assert bool((a[0][..., m] == want).all())                 # HWF = season length for every definition
            assert bool((a[1][..., m] == (want > 0).to(torch.int32)).all())   # one heatwave id
            assert bool((a[2][..., m] == want).all())


def test_lens_member_full_grid(core):
    from hdp_b200 import synth, workloads
    wl = workloads.get("lens_member")
    lat, _ = synth.grid_latitudes(wl.n_lat, wl.n_lon)
    _run(core, wl, lat, 99, 48)


def test_era5_full_grid_thresholds(core):
    # the whole 721 x 1440 grid (1 038 240 cells): 45.5 GB of samples in, 30.4 GB of thresholds out, standard calendar
    # (366 day-of-year rows, -1 pads), 31-day window.  Properties are checked band by band to bound the temporaries.
    from hdp_b200 import synth, workloads
    wl = workloads.get("era5_025deg")
    lat, _ = synth.grid_latitudes(wl.n_lat, wl.n_lon)
    base, _ = _fields(wl, lat, 7, run=False)
    thr = core.thresholds_array(base, wl.window_tables(), wl.percentiles)
    q = np.asarray(wl.percentiles)
    step = 64 * wl.n_lon
    for c0 in range(0, lat.size, step):
        _check_threshold_properties(base[:, c0:c0 + step], thr[c0:c0 + step], q)
    _oracle_sample(core, wl, lat, base, None, thr, None, 24, 7)


def test_run_sharded_on_gpu_world1(core):
    # the product's shard entry point with the CUDA kernels (world size 1: the partition is the whole range, no collective),
    # full-array and local-shard forms, against the oracle; the N > 1 partition + gather logic is covered over gloo in test_shard.py
    from hdp_b200 import _tables as tb, shard, synth, workloads
    wl = workloads.get("lens_member")
    lat, _ = synth.grid_latitudes(wl.n_lat, wl.n_lon)
    lat = lat[::37][:1000]
    base, warm = _fields(wl, lat, 11)
    st, dm = wl.seasons(), tb.doy_map(wl.run_axis().dayofyr)
    south = torch.as_tensor((lat < 0).astype(np.uint8), device="cuda")
    thr, out = shard.run_sharded(base, warm, wl.window_tables(), wl.percentiles, dm, wl.defs, st.north, st.south, south)
    _oracle_sample(core, wl, lat, base, warm, thr, out, 16, 3)
    thr2 = torch.empty_like(thr); out2 = torch.empty_like(out)
    shard.run_sharded(base, warm, wl.window_tables(), wl.percentiles, dm, wl.defs, st.north, st.south, south,
                      gather=False, local_of=lat.size, out=(thr2, out2))
    assert torch.equal(thr2.view(torch.int64), thr.view(torch.int64)) and torch.equal(out2.view(torch.int16), out.view(torch.int16))
    for m, g0, g1 in shard.member_pieces(100, 900, 400):               # pieces of a member-sharded range meet their thresholds
        piece = core.metrics_array(warm[:, g0:g1], thr[g0:g1], dm, wl.defs, st.north, st.south, south[g0:g1])
        assert torch.equal(piece.view(torch.int16), out.view(torch.int16)[..., g0:g1])


def test_wide_sweep_full_grid(core):
    from hdp_b200 import synth, workloads
    wl = workloads.get("wide_sweep")
    lat, _ = synth.grid_latitudes(wl.n_lat, wl.n_lon)
    _run(core, wl, lat, 555, 24)
