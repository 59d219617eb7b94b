"""The reference's own CPU path for the two hot loops, one worker process per host core - TEST INFRASTRUCTURE.

This is what bench.py times as the CPU baseline (``cpu_baseline.kind = "reference"``, ``--impl reference``): the
UNMODIFIED Numba kernels of AgentOxygen/HDP (imported from oracle/_ref/ through oracle/ref_numba.py), driven
exactly like the reference drives them,

* thresholds: the ``compute_percentiles`` gufunc over the cells of a block (hdp/threshold.py:52-93),
* metrics: ``for perc: for hw_def: for cell: compute_heatwave_metrics(...)`` (hdp/metric.py:357-367; the
  ``apply_ufunc(vectorize=True)`` there is one Python -> Numba call per cell),

with a ``multiprocessing`` pool standing in for ``LocalCluster(processes=True)`` (the Numba kernels hold the GIL,
docs/examples.rst:122; Dask and xarray cannot be installed in this image).  Cells are dealt out in contiguous
blocks, one block per worker, like Dask's spatial chunks.  Workers import and JIT-compile the reference once, in
``RefPool.__init__`` (not timed).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from typing import Optional, Sequence

import numpy as np

_REF = None


def _init_worker():
    global _REF
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import ref_numba
    thr, met, _ = ref_numba.load()
    _REF = (thr, met)
    # warm the lazy @njit kernels (the gufunc is compiled eagerly at import)
    x = np.linspace(0, 1, 40, dtype=np.float32)
    thr.compute_percentiles(x[None, :], np.arange(40, dtype=np.int64).reshape(4, 10), np.array([0.5]))
    met.compute_heatwave_metrics(x, np.full(10, 0.5), (np.arange(40) % 10).astype(np.int64), 1, 0, 0,
                                 np.array([[0, 20], [20, 40]], dtype=np.int64))


def _ready(_):
    return os.getpid()


def _thr_block(args):
    temps_ct, win, q = args
    t0 = time.perf_counter()
    out = _REF[0].compute_percentiles(temps_ct, win, q)                    # [cells, n_doy, P] float64
    return out, time.perf_counter() - t0


def _met_block(args):
    run_ct, thr, doy_map, defs, north, south, is_south = args
    fn = _REF[1].compute_heatwave_metrics
    n, P, D, Y = run_ct.shape[0], thr.shape[2], len(defs), north.shape[0]
    out = np.empty((P, D, n, 4, Y), np.int64)
    t0 = time.perf_counter()
    for p in range(P):                                                      # hdp/metric.py:357
        for d, (a, b, c) in enumerate(defs):                                # :358
            for i in range(n):                                              # apply_ufunc(vectorize=True), :360-366
                rng = south if is_south[i] else north
                out[p, d, i] = fn(run_ct[i], np.ascontiguousarray(thr[i, :, p]), doy_map, int(a), int(b), int(c), rng)
    return out, time.perf_counter() - t0


class RefPool:
    def __init__(self, workers: Optional[int] = None):
        from oracle import ref_numba
        if not ref_numba.available():
            raise RuntimeError("the reference is not installed (oracle/_ref/ is made by oracle/build.py:build_ref())")
        self.workers = workers or max(1, len(os.sched_getaffinity(0)))
        ctx = mp.get_context("spawn")                                      # no forked CUDA / OpenMP state in the workers
        self.pool = ctx.Pool(self.workers, initializer=_init_worker)
        self.pool.map(_ready, range(4 * self.workers))                     # all workers imported + compiled

    def close(self):
        self.pool.terminate()
        self.pool.join()

    def _blocks(self, n: int):
        edges = np.linspace(0, n, min(self.workers, n) + 1).astype(int)
        return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]

    def thresholds(self, base_tc: np.ndarray, window_samples: np.ndarray, q: Sequence[float]):
        """[T, C] float32 -> ([C, n_doy, P] float64, wall seconds)."""
        ct = np.ascontiguousarray(base_tc.T, dtype=np.float32)
        win = np.ascontiguousarray(window_samples, dtype=np.int64)
        qq = np.ascontiguousarray(q, dtype=np.float64)
        t0 = time.perf_counter()
        parts = self.pool.map(_thr_block, [(ct[a:b], win, qq) for a, b in self._blocks(ct.shape[0])])
        dt = time.perf_counter() - t0
        return np.concatenate([p[0] for p in parts], axis=0), dt

    def metrics(self, run_tc: np.ndarray, thr: np.ndarray, doy_map, defs, north, south, is_south):
        """[T, C] float32, [C, n_doy, P] float64 -> (int64 [P, D, C, 4, Y], wall seconds)."""
        ct = np.ascontiguousarray(run_tc.T, dtype=np.float32)
        dm = np.ascontiguousarray(doy_map, dtype=np.int64)
        nn, ss = np.ascontiguousarray(north, dtype=np.int64), np.ascontiguousarray(south, dtype=np.int64)
        df = [tuple(int(v) for v in d) for d in defs]
        sth = np.zeros(ct.shape[0], np.uint8) if is_south is None else np.asarray(is_south, dtype=np.uint8)
        t0 = time.perf_counter()
        parts = self.pool.map(_met_block, [(ct[a:b], np.ascontiguousarray(thr[a:b]), dm, df, nn, ss, sth[a:b])
                                           for a, b in self._blocks(ct.shape[0])])
        dt = time.perf_counter() - t0
        return np.concatenate([p[0] for p in parts], axis=2), dt
