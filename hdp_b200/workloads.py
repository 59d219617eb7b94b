"""The named workloads of BASELINE.json (`configs`), as data: grid, calendars, sweep parameters.

bench.py measures ``cmip6_1deg`` (configs[1]); the others are parity-test / exploration cases.
Algorithmic bytes follow SURVEY.md section 8d:

    B_thr = C * (4*T_b + 8*n_doy*P)                     read each f32 sample once, write f64 thresholds once
    B_met = C * (4*T + 8*n_doy*P + 2*4*Y*P*D)           read samples + thresholds once, write 4 u16 metrics
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import _tables as tb

README_DEFS = [[3, 0, 0], [3, 1, 1], [4, 0, 0], [4, 1, 1], [5, 0, 0], [5, 1, 1]]      # reference README.md:51
WIDE_DEFS = [[a, b, c] for a in (3, 4, 5, 6) for b in (0, 1, 2) for c in (0, 1)]


@dataclass
class Workload:
    name: str
    n_lat: int
    n_lon: int
    base_years: int
    run_years: int
    measures: int
    percentiles: np.ndarray
    defs: List[Sequence[int]]
    radius: int = 7
    calendar: str = "noleap"
    base_start: int = 1961
    run_start: int = 2015
    description: str = ""
    members: int = 1                   # ensemble members of the metric run that share ONE member-free threshold table
    _cache: dict = field(default_factory=dict, repr=False)

    @property
    def cells(self) -> int:
        return self.n_lat * self.n_lon

    def base_axis(self) -> tb.TimeAxis:
        if "base" not in self._cache:
            self._cache["base"] = tb.TimeAxis.date_range(f"{self.base_start}-01-01", f"{self.base_start + self.base_years - 1}-12-31",
                                                         self.calendar)
        return self._cache["base"]

    def run_axis(self) -> tb.TimeAxis:
        if "run" not in self._cache:
            self._cache["run"] = tb.TimeAxis.date_range(f"{self.run_start}-01-01", f"{self.run_start + self.run_years - 1}-12-31",
                                                        self.calendar)
        return self._cache["run"]

    def window_tables(self) -> tb.WindowTables:
        if "wt" not in self._cache:
            self._cache["wt"] = tb.window_tables(self.base_axis().dayofyr, self.radius)
        return self._cache["wt"]

    def seasons(self) -> tb.SeasonTables:
        if "st" not in self._cache:
            self._cache["st"] = tb.hemisphere_ranges(self.run_axis())
        return self._cache["st"]

    # ---- algorithmic bytes per measure
    def bytes_thresholds(self, cells: int = None) -> int:
        C = self.cells if cells is None else cells
        wt = self.window_tables()
        return C * (4 * len(self.base_axis()) + 8 * wt.n_doy * len(self.percentiles))

    def bytes_metrics(self, cells: int = None) -> int:
        if not self.run_years:
            return 0
        C = self.cells if cells is None else cells
        n_doy = self.window_tables().n_doy
        P, D, Y = len(self.percentiles), len(self.defs), self.seasons().n_years
        return C * (4 * len(self.run_axis()) + 8 * n_doy * P + 2 * 4 * Y * P * D)

    def cell_years(self, cells: int = None) -> int:
        C = self.cells if cells is None else cells
        return C * (self.base_years + self.members * self.run_years) * self.measures


def get(name: str) -> Workload:
    if name == "cmip6_1deg":
        return Workload(name, 180, 360, 30, 86, 3, np.arange(0.9, 1.0, 0.01), README_DEFS,
                        description="CMIP6-like 1 deg grid 180x360, 30-year baseline thresholds + 86-year metric run, "
                                    "tmax/tmin/tavg, 10 percentiles x 6 definitions, 15-day window, noleap")
    if name == "wide_sweep":
        return Workload(name, 180, 360, 30, 86, 3, np.linspace(0.80, 0.99, 20), WIDE_DEFS,
                        description="1 deg grid, 3 measures x 20 percentiles x 24 definitions")
    if name == "era5_025deg":
        return Workload(name, 721, 1440, 30, 0, 1, np.arange(0.9, 1.0, 0.01), README_DEFS, radius=15, calendar="standard",
                        base_start=1991, description="ERA5-like 0.25 deg grid 721x1440, 30-year baseline, 31-day window, thresholds only")
    if name == "lens_member":
        return Workload(name, 192, 288, 30, 86, 1, np.arange(0.9, 1.0, 0.01), README_DEFS,
                        description="one CESM2-LENS-like member: 192x288, 30-year baseline + 86-year run")
    if name == "lens50":
        return Workload(name, 192, 288, 30, 86, 1, np.arange(0.9, 1.0, 0.01), README_DEFS, members=50,
                        description="CESM2-LENS-like 192x288, 50 ensemble members x 86 years against one 30-year baseline, "
                                    "sharded by flattened (member, cell) index across the GPUs")
    raise KeyError(name)
