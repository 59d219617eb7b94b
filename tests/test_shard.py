"""Multi-GPU host logic on CPU: cell partition + output gather with world_size 2 and 3 over gloo.

The CUDA kernels are replaced by the oracle (allowed here: tests/ is one of the places that may call it) so that the
partition, the shard-local chaining thresholds -> metrics and the padded all_gather are exercised without a GPU."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from hdp_b200 import _tables as tb, shard  # noqa: E402


def test_cell_ranges_cover_and_align():
    for C in (0, 1, 31, 32, 33, 1000, 64800, 55296 * 50):
        for world in (1, 2, 3, 4, 8):
            r = shard.all_ranges(C, world)
            assert r[0][0] == 0 and r[-1][1] == C
            for (a0, a1), (b0, b1) in zip(r, r[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(a % shard.CELL_ALIGN == 0 for a, _ in r if a < C)
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) < 2 * shard.CELL_ALIGN        # one block of imbalance + a ragged last block
    with pytest.raises(ValueError):
        shard.cell_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_kernels():
    import oracle

    def thresholds(x, wt, q):
        return torch.from_numpy(oracle.thresholds_batch(np.ascontiguousarray(x.numpy()), wt.window_samples(), q))

    def metrics(x, thr, dm, defs, north, south, is_south):
        out = oracle.metrics_batch(np.ascontiguousarray(x.numpy()), thr.numpy(), dm, defs, north, south, is_south)   # [P, D, C, 4, Y]
        return torch.from_numpy(np.ascontiguousarray(out.transpose(3, 0, 1, 4, 2)).astype(np.uint16))                # [4, P, D, Y, C]

    return thresholds, metrics


def _case(C):
    rng = np.random.default_rng(C)
    base_ax = tb.TimeAxis.daily((1990, 1, 1), 3 * 365, "noleap")
    run_ax = tb.TimeAxis.daily((2000, 1, 1), 4 * 365, "noleap")
    season = lambda ax: 10 * np.sin(2 * np.pi * (ax.dayofyr[:, None] - 110) / 365)          # noqa: E731
    base = (season(base_ax) + 3 * rng.standard_normal((len(base_ax), C))).astype(np.float32)
    run = (season(run_ax) + 1 + 3 * rng.standard_normal((len(run_ax), C))).astype(np.float32)
    st = tb.hemisphere_ranges(run_ax)
    return dict(base=base, run=run, wt=tb.window_tables(base_ax.dayofyr, 2), q=np.array([0.8, 0.9]),
                dm=tb.doy_map(run_ax.dayofyr), defs=[[3, 0, 0], [2, 1, 1]], north=st.north, south=st.south,
                is_south=(np.arange(C) % 3 == 0).astype(np.uint8))


def _worker(rank, world, port, C, result_dir, local=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k = _case(C)
        if local:                       # every rank holds only its shard of a problem no rank holds whole (bench.py --gpus N)
            c0, c1 = shard.cell_range(C, rank, world)
            thr, met = shard.run_sharded(torch.from_numpy(k["base"][:, c0:c1].copy()), torch.from_numpy(k["run"][:, c0:c1].copy()),
                                         k["wt"], k["q"], k["dm"], k["defs"], k["north"], k["south"], k["is_south"][c0:c1],
                                         kernels=_oracle_kernels(), local_of=C)
        else:
            thr, met = shard.run_sharded(torch.from_numpy(k["base"]), torch.from_numpy(k["run"]), k["wt"], k["q"], k["dm"], k["defs"],
                                         k["north"], k["south"], k["is_south"], kernels=_oracle_kernels())
        np.savez(os.path.join(result_dir, f"rank{rank}.npz"), thr=thr.numpy(), met=met.numpy().astype(np.int64))
        # a gathered tensor must also round-trip for 1-D per-cell vectors
        c0, c1 = shard.cell_range(C, rank, world)
        ids = shard.gather_cells(torch.arange(c0, c1, dtype=torch.int64), C, dim=0)
        assert ids.tolist() == list(range(C))
        # the rank-major form: every shard as it lies in memory, trimmed by its range
        local = met[..., c0:c1].contiguous()
        bucket, ranges = shard.gather_shards(local, C, dim=-1)
        rows = local.numel() // max(c1 - c0, 1) if c1 > c0 else int(np.prod(met.shape[:-1]))
        for r, (a, b) in enumerate(ranges):
            blk = bucket[r, : (b - a) * rows * 2].view(torch.uint16).view(*met.shape[:-1], b - a)
            assert torch.equal(blk.view(torch.int16), met[..., a:b].contiguous().view(torch.int16))
    finally:
        dist.destroy_process_group()


def test_member_pieces():
    assert shard.member_pieces(0, 10, 10) == [(0, 0, 10)]
    assert shard.member_pieces(5, 27, 10) == [(0, 5, 10), (1, 0, 10), (2, 0, 7)]
    assert shard.member_pieces(7, 7, 10) == []
    grid, members, world = 55296, 50, 8
    seen = []
    for r in range(world):
        c0, c1 = shard.cell_range(grid * members, r, world)
        seen += [(m * grid + g0, m * grid + g1) for m, g0, g1 in shard.member_pieces(c0, c1, grid)]
    assert seen[0][0] == 0 and seen[-1][1] == grid * members and all(a[1] == b[0] for a, b in zip(seen, seen[1:]))


@pytest.mark.parametrize("world,C,local", [(2, 70, False), (3, 101, False), (2, 20, False), (2, 128, False), (2, 128, True), (3, 101, True)])
def test_sharded_run_equals_single_process(tmp_path, world, C, local):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, C, str(tmp_path), local), nprocs=world, join=True)
    k = _case(C)
    thresholds, metrics = _oracle_kernels()
    thr = thresholds(torch.from_numpy(k["base"]), k["wt"], k["q"])
    met = metrics(torch.from_numpy(k["run"]), thr, k["dm"], k["defs"], k["north"], k["south"], k["is_south"])
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        assert np.array_equal(got["thr"].view(np.uint64), thr.numpy().view(np.uint64))
        assert np.array_equal(got["met"], met.numpy().astype(np.int64))
