"""Population analytics.

The reference ends a run with plot4Figures.m: current, voltage, SOC and side-reaction overpotential of ONE cell against
time, with the limits drawn in (plot4Figures.m:29-75).  For a population the same four quantities become envelopes over the
logged cells, and the per-cell outcomes (time to the SOC target, lowest overpotential, highest voltage, Hildreth effort)
become distributions.  Everything here is host-side NumPy on what the CUDA path produced (CellBatch.logs(),
CellBatch.cell_summaries()); histograms are additive, so shards on several GPUs are merged with one all-reduce of the
bin counts (reduce_histograms) -- the same "gather of summary statistics" as mek/dist.py, nothing on the data path.
"""
from __future__ import annotations

import numpy as np

QUANTILES = (0.0, 0.05, 0.5, 0.95, 1.0)


def trajectory_envelopes(logs, quantiles=QUANTILES):
    """The four panels of plot4Figures.m for a population: per sample instant, the given quantiles over the logged cells of
    u (current, A), V, SOC and phise.  Lost cells (NaN, quirk Q19) are ignored.  Returns {name: [len(quantiles), N]}."""
    out = {}
    for k in ("u", "V", "SOC", "phise"):
        a = np.asarray(logs[k], float)
        with np.errstate(all="ignore"):
            out[k] = np.nanquantile(a, quantiles, axis=1) if a.shape[1] else np.full((len(quantiles), a.shape[0]), np.nan)
    return out


def constraint_blocks(Np, Nc, constraints=(1, 1, 1), use_zmax=True):
    """Row ranges of the stacked constraints in the reference's order (constraintsMPC.m:23-101):
    u_max, u_min, du_max, du_min (Nc rows each), voltage, eta, z_max (Np rows each)."""
    blocks, r = {}, 0
    if constraints[0]:
        for nm in ("u_max", "u_min", "du_max", "du_min"):
            blocks[nm] = (r, r + Nc); r += Nc
    for nm, on in (("v_max", constraints[1]), ("phise_min", constraints[2]), ("z_max", use_zmax)):
        if on:
            blocks[nm] = (r, r + Np); r += Np
    return blocks, r


def constraint_activity(logs, Np=5, Nc=2, constraints=(1, 1, 1), use_zmax=True):
    """Constraint-activity timeline: per sample instant, the fraction of logged cells whose QP ended with a positive
    multiplier in each constraint block (the active set on exit of hildreth.m).  Returns {block: [N]}."""
    act = np.asarray(logs["active"], np.uint32)                  # [N, cells, 4]
    blocks, nC = constraint_blocks(Np, Nc, constraints, use_zmax)
    bits = np.zeros(act.shape[:2] + (nC,), bool)
    for i in range(nC):
        bits[:, :, i] = (act[:, :, i >> 5] >> np.uint32(i & 31)) & np.uint32(1)
    ncell = max(act.shape[1], 1)
    return {nm: bits[:, :, a:b].any(axis=2).sum(axis=1) / ncell for nm, (a, b) in blocks.items()}


def outcome_histograms(cs, nsim, Ts=1.0, bins=64, v_range=(3.9, 4.3), phise_range=(0.0, 0.2), sweep_range=(0.0, 100.0)):
    """Distributions of per-cell outcomes from CellBatch.cell_summaries(): time to reach the SOC target (s; cells that
    never got there are counted in `not_reached`), lowest side-reaction overpotential, highest voltage and mean Hildreth
    sweeps per QP.  Fixed bin edges (so that shards can be added): returns {name: (counts[bins], edges[bins+1])} and counters."""
    t = np.asarray(cs["t_reach"], float)
    ok = t >= 0
    nqp = np.maximum(np.asarray(cs["n_qp"], float), 1.0)
    h = {
        "t_reach_s": np.histogram(t[ok] * Ts, bins=bins, range=(0.0, nsim * Ts)),
        "min_phise": np.histogram(np.asarray(cs["min_phise"], float), bins=bins, range=phise_range),
        "max_V": np.histogram(np.asarray(cs["max_V"], float), bins=bins, range=v_range),
        "sweeps_per_qp": np.histogram(np.asarray(cs["sum_nexec"], float) / nqp, bins=bins, range=sweep_range),
    }
    st = np.asarray(cs["status"], float).astype(np.int64)
    counters = {"cells": int(len(t)), "not_reached": int((~ok).sum()), "lost": int(((st & 1) != 0).sum()),
                "q18": int(((st & 2) != 0).sum()), "nonfinite": int(((st & 4) != 0).sum()),
                "cap_hits": float(np.asarray(cs["n_cap"], float).sum())}
    return {k: (c.astype(np.int64), e) for k, (c, e) in h.items()}, counters


def reduce_histograms(hists, counters, device=None, group=None):
    """Sum histograms and counters over the process group (NCCL on GPUs, gloo on CPU); a no-op for a single process."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return hists, counters
    names, ckeys = sorted(hists), sorted(counters)
    flat = np.concatenate([hists[k][0].astype(np.float64) for k in names] + [np.array([float(counters[k]) for k in ckeys])])
    t = torch.tensor(flat, dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    flat = t.cpu().numpy()
    out, o = {}, 0
    for k in names:
        n = len(hists[k][0])
        out[k] = (np.rint(flat[o:o + n]).astype(np.int64), hists[k][1]); o += n
    cnt = {k: (float(flat[o + i]) if k == "cap_hits" else int(round(flat[o + i]))) for i, k in enumerate(ckeys)}
    return out, cnt


def quantile_from_histogram(counts, edges, q):
    """Quantile of a histogrammed quantity (linear inside the bin)."""
    c = np.cumsum(counts).astype(float)
    if c[-1] <= 0:
        return float("nan")
    target = q * c[-1]
    i = int(np.searchsorted(c, target, side="left"))
    i = min(i, len(counts) - 1)
    prev = c[i - 1] if i > 0 else 0.0
    frac = 0.0 if counts[i] == 0 else (target - prev) / counts[i]
    return float(edges[i] + frac * (edges[i + 1] - edges[i]))
