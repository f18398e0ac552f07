"""Synthetic cell populations for the configurations in BASELINE.json / SURVEY.md section 8(d).

Every per-cell draw is a pure function of (seed, cell_id, stream) through a counter-based hash
(splitmix64), so any shard of any population can be regenerated independently on any rank.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
    return z ^ (z >> np.uint64(31))


def uniform01(seed, cell_ids, stream):
    """U[0,1) doubles keyed by (seed, cell_id, stream)."""
    with np.errstate(over="ignore"):
        ids = np.asarray(cell_ids, dtype=np.uint64)
        h = _splitmix64(ids * np.uint64(0x100000001B3) + np.uint64(stream) * np.uint64(0xD6E8FEB86659FD93)
                        + _splitmix64(np.full_like(ids, np.uint64(seed))))
        h = _splitmix64(h)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


@dataclass
class Population:
    """Host-side inputs of a population (what initKF/initMPC/OB_step's first call take per cell)."""
    SOC0: np.ndarray        # [B] percent (plant truth, runMPC.m:13)
    ekf_SOC0: np.ndarray    # [B] percent (initKF's SOC0, runMPC.m:20)
    T: np.ndarray           # [B] deg C
    SigmaX0: np.ndarray     # [B,6,6]
    limits: np.ndarray      # [B,7] u_max,u_min,du_max,du_min,v_max,phise_min,z_max(+z_tol)

    def __len__(self):
        return len(self.SOC0)

    def slice(self, lo, hi):
        return Population(self.SOC0[lo:hi], self.ekf_SOC0[lo:hi], self.T[lo:hi], self.SigmaX0[lo:hi], self.limits[lo:hi])


def default_limits(Q, B=1, Crate=2.0, u_max=2.0, du_max=50.0, du_min=-50.0, v_max=4.1, phise_min=0.08, z_max=0.95, z_tol=0.0):
    """runMPC.m:36-44 and initMPC.m:66-67 (u_min = -Q*Crate)."""
    lim = np.empty((B, 7))
    lim[:, 0] = u_max; lim[:, 1] = -Q * np.asarray(Crate); lim[:, 2] = du_max; lim[:, 3] = du_min
    lim[:, 4] = v_max; lim[:, 5] = phise_min; lim[:, 6] = z_max + z_tol
    return lim


def single_cell(Q):
    """C1: runMPC.m's own configuration (runMPC.m:8-44)."""
    return Population(np.array([10.0]), np.array([10.0]), np.array([25.0]),
                      np.diag([1.0] * 5 + [2e6])[None].copy(), default_limits(Q))


def make_population(Q, B, seed, heterogeneous=False, first_cell=0, ekf_soc_error=0.0):
    """C2/C3/C4 (heterogeneous=False) and C5 (True) distributions of SURVEY.md section 8(d)."""
    ids = np.arange(first_cell, first_cell + B, dtype=np.uint64)
    u = lambda s: uniform01(seed, ids, s)
    SOC0 = 5.0 + 25.0 * u(0)
    s = np.exp(np.log(0.1) + (np.log(10.0) - np.log(0.1)) * u(1))
    s0 = np.exp(np.log(2e5) + (np.log(2e7) - np.log(2e5)) * u(2))
    SigmaX0 = np.zeros((B, 6, 6))
    for j in range(5):
        SigmaX0[:, j, j] = s
    SigmaX0[:, 5, 5] = s0
    T = np.full(B, 25.0)
    lim = default_limits(Q, B)
    if heterogeneous:
        Crate = 1.0 + 2.0 * u(3)
        lim[:, 1] = -Q * Crate
        lim[:, 4] = 4.05 + 0.15 * u(4)
        lim[:, 5] = 0.05 + 0.07 * u(5)
        dumax = 20.0 + 60.0 * u(6)
        lim[:, 2] = dumax; lim[:, 3] = -dumax
        T = 15.0 + 20.0 * u(7)
    ekf = SOC0 + ekf_soc_error * (2.0 * u(8) - 1.0)
    return Population(SOC0, ekf, T, SigmaX0, lim)
