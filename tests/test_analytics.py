"""Population analytics (mek/analytics.py): host-side logic on synthetic logs (CPU), the per-cell summary read-back on
the GPU, and the histogram reduction over a gloo process group (world size 2)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_logs(N=50, B=7):
    rng = np.random.default_rng(0)
    lg = dict(u=rng.uniform(-60, 0, (N, B)), V=rng.uniform(3.5, 4.1, (N, B)), SOC=np.sort(rng.uniform(0.1, 0.95, (N, B)), axis=0),
              phise=rng.uniform(0.08, 0.15, (N, B)), active=np.zeros((N, B, 4), np.uint32))
    lg["u"][10:, 3] = np.nan                                   # a lost cell
    lg["active"][5:, :4, 0] |= np.uint32(1 << 2)               # u_min row 1 (row 2) for four cells
    lg["active"][20:, 1, 0] |= np.uint32(1 << 14)              # phise row 2 (row 8+5+1 = 14)
    return lg


def test_envelopes_and_constraint_activity():
    from mek import analytics as A
    lg = _fake_logs()
    env = A.trajectory_envelopes(lg)
    assert env["u"].shape == (5, 50)
    assert np.all(env["V"][0] <= env["V"][2]) and np.all(env["V"][2] <= env["V"][4])
    assert np.allclose(env["u"][4, 12], np.nanmax(lg["u"][12]))          # NaN cells ignored
    blocks, nC = A.constraint_blocks(5, 2)
    assert nC == 23 and blocks["u_min"] == (2, 4) and blocks["v_max"] == (8, 13) and blocks["phise_min"] == (13, 18) and blocks["z_max"] == (18, 23)
    act = A.constraint_activity(lg)
    assert act["u_min"][4] == 0 and abs(act["u_min"][5] - 4 / 7) < 1e-12
    assert act["phise_min"][19] == 0 and abs(act["phise_min"][20] - 1 / 7) < 1e-12
    assert not act["v_max"].any() and not act["z_max"].any()
    b2, n2 = A.constraint_blocks(20, 5, (1, 1, 0))
    assert n2 == 60 and b2["z_max"] == (40, 60) and "phise_min" not in b2


def test_outcome_histograms_and_quantiles():
    from mek import analytics as A
    B = 1000
    rng = np.random.default_rng(1)
    cs = dict(t_reach=rng.integers(2000, 2900, B).astype(float), min_phise=rng.uniform(0.079, 0.081, B), max_V=rng.uniform(4.09, 4.101, B),
              sum_nexec=rng.uniform(1e4, 3e4, B), n_qp=np.full(B, 3001.0), n_cap=np.full(B, 4.0), status=np.zeros(B))
    cs["t_reach"][:10] = -1.0; cs["status"][3] = 1.0
    h, cnt = A.outcome_histograms(cs, nsim=3001)
    assert cnt == {"cells": B, "not_reached": 10, "lost": 1, "q18": 0, "nonfinite": 0, "cap_hits": 4000.0}
    assert h["t_reach_s"][0].sum() == B - 10 and h["max_V"][0].sum() == B
    med = A.quantile_from_histogram(*h["t_reach_s"], 0.5)
    assert abs(med - np.median(cs["t_reach"][10:])) < 3001 / 64
    # additivity: two shards add up to the whole
    a = {k: v[:500] for k, v in cs.items()}; b = {k: v[500:] for k, v in cs.items()}
    ha, ca = A.outcome_histograms(a, 3001); hb, cbb = A.outcome_histograms(b, 3001)
    for k in h:
        assert np.array_equal(ha[k][0] + hb[k][0], h[k][0])


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "mpc-ekf4fastcharge_b200"))
    import torch.distributed as dist
    from mek import analytics as A
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    B = 300 + 17 * rank
    cs = dict(t_reach=rng.integers(1500, 2900, B).astype(float), min_phise=rng.uniform(0.07, 0.09, B), max_V=rng.uniform(4.0, 4.11, B),
              sum_nexec=rng.uniform(1e4, 3e4, B), n_qp=np.full(B, 3001.0), n_cap=np.full(B, float(rank + 1)), status=np.zeros(B))
    h, c = A.outcome_histograms(cs, 3001)
    hr, cr = A.reduce_histograms(h, c)
    q.put((rank, {k: v[0].tolist() for k, v in h.items()}, c, {k: v[0].tolist() for k, v in hr.items()}, cr))
    dist.barrier()
    dist.destroy_process_group()


def test_histogram_reduction_gloo_world2():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted([q.get(timeout=120) for _ in ps])
    for p in ps: p.join(60)
    (_, h0, c0, r0, cr0), (_, h1, c1, r1, cr1) = res
    assert r0 == r1 and cr0 == cr1
    for k in h0:
        assert r0[k] == (np.array(h0[k]) + np.array(h1[k])).tolist()
    assert cr0["cells"] == c0["cells"] + c1["cells"] and cr0["cap_hits"] == c0["cap_hits"] + c1["cap_hits"]


@pytest.mark.gpu
def test_cell_summaries_match_logs(bank):
    from mek import analytics as A, api, population as P
    B, N = 200, 1500
    pop = P.make_population(bank.params[2], B, seed=9, heterogeneous=True)
    cb = api.CellBatch(api.RomBank(bank), B)
    cb.set_logging(B, N)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    cb.run(0, 700); cb.run(700, N)
    lg, cs = cb.logs(0, N), cb.cell_summaries()
    assert np.array_equal(cs["sum_nexec"], lg["nexec"].sum(axis=0)) and np.array_equal(cs["n_qp"], (lg["nexec"] > 0).sum(axis=0))
    assert np.array_equal(cs["min_phise"], lg["phise"].min(axis=0)) and np.array_equal(cs["max_V"], lg["V"].max(axis=0))
    assert np.array_equal(cs["last_soc"], lg["SOC"][-1]) and np.all(cs["steps"] == N)
    zmax = pop.limits[:, 6]
    first = np.array([np.argmax(lg["SOC"][:, c] >= zmax[c] - 1e-3) if (lg["SOC"][:, c] >= zmax[c] - 1e-3).any() else -1 for c in range(B)])
    assert np.array_equal(cs["t_reach"], first.astype(float))
    s = cb.summary()
    h, cnt = A.outcome_histograms(cs, N)
    assert cnt["cells"] == B and cnt["not_reached"] == B - int(s["n_reached"]) and cnt["cap_hits"] == s["n_cap"]
    act = A.constraint_activity(lg)
    assert act["u_min"][50:300].mean() > 0.9                    # constant-current phase: the lower current bound is active
    env = A.trajectory_envelopes(lg)
    assert np.all(env["SOC"][2, 1:] >= env["SOC"][2, :-1] - 1e-6)  # median SOC never decreases during a charge


def test_plot4figures_svg(tmp_path):
    """mek/plots.py: the four panels of plot4Figures.m (current, voltage, SOC, overpotential with their limit lines) as an SVG
    file, for one cell and for a population (median + 5-95 % band), from logs shaped like CellBatch.logs()."""
    import xml.etree.ElementTree as ET
    from mek import plots
    rng = np.random.default_rng(3)
    N, B = 400, 24
    t = np.arange(N)[:, None]
    logs = dict(u=-60.0 + 40.0 * (t / N) + rng.normal(0, 1, (N, B)), V=3.5 + 0.6 * (t / N) + rng.normal(0, 0.005, (N, B)),
                SOC=0.1 + 0.85 * (t / N) * np.ones((1, B)), phise=0.2 - 0.12 * (t / N) + rng.normal(0, 0.002, (N, B)))
    logs["V"][300:, 5] = np.nan                                      # a lost cell (Q19) must not break the figure
    lim = dict(u_max=2.0, u_min=-59.7, v_max=4.1, v_min=3.4, phise_min=0.08)
    for cell, f in ((None, "pop.svg"), (3, "cell.svg")):
        p = str(tmp_path / f)
        svg = plots.plot4figures_svg(logs, p, limits=lim, ref=95.0, cell=cell)
        root = ET.fromstring(open(p).read())
        assert root.tag.endswith("svg") and svg.startswith("<svg")
        ns = "{http://www.w3.org/2000/svg}"
        panels = [g.get("id") for g in root.iter(ns + "g")]
        assert panels == ["panel-u", "panel-V", "panel-SOC", "panel-phise"]
        assert len([l for l in root.iter(ns + "line") if l.get("class") == "limit"]) == 2 + 2 + 1 + 2   # u, V limits, target, phi_min + zero line
        assert (len(list(root.iter(ns + "polygon"))) == 4) == (cell is None)
