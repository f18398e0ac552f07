"""GPU tests at BASELINE.json's full size (config C3: 1 Mi cells, one GPU): parity on a sub-sample of cells against the
oracle plus size-independent properties of the whole population."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

B_FULL = 1 << 20
N_SIM = 3001


def test_c3_one_million_cells_subsample_parity_and_population_properties(bank, coracle_mod):
    from mek import api, population as P
    Q = bank.params[2]
    pop = P.make_population(Q, B_FULL, seed=2)
    rom = api.RomBank(bank)
    cb = api.CellBatch(rom, B_FULL)
    nlog = 4096                                   # parity on a 4 096-cell sub-sample
    cb.set_logging(nlog, N_SIM)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    cb.run(0, N_SIM)
    g = cb.logs()
    s = cb.summary()
    # (1) the first 4 096 cells (caller's indices; they are scattered over the SOC-ordered state block) match the oracle bit for bit
    sub = pop.slice(0, nlog)
    co = coracle_mod.COracle(bank)
    co.init(sub.SOC0, sub.T, sub.SigmaX0, sub.limits, sub.ekf_SOC0)
    c = co.run(0, N_SIM, nthreads=max(8, os.cpu_count() or 8))
    for k in ("u", "V", "SOC", "phise", "iT", "iZ", "nexec", "active"):
        assert np.array_equal(g[k], c[k]), k
    # (2) population-level properties
    assert s["cells"] == B_FULL and s["cellsteps"] == float(B_FULL) * N_SIM
    assert s["n_lost"] == 0 and s["n_nonfinite"] == 0 and s["n_q18"] == 0
    assert s["n_reached"] == B_FULL                                  # every cell reaches 95 % within 3000 s
    assert s["n_qp"] == s["cellsteps"]                               # Q9: Hildreth runs on every step
    assert 0.0785 < s["min_phise"] < 0.08 and 4.1 <= s["max_V"] < 4.102   # constraints ride their limits
    assert abs(s["sum_soc"] / B_FULL - 0.95) < 1e-4                  # SOC held at z_max
    mean_sweeps = s["sum_nexec"] / s["cellsteps"]
    assert abs(mean_sweeps - c["nexec"].mean()) < 0.15               # sub-sample is representative
    # (3) determinism and restartability: a second charge split in two launches reproduces every statistic exactly
    cb.reset()
    cb.run(0, 1000)
    cb.run(1000, N_SIM)
    s2 = cb.summary()
    for k in s:
        if k == "sum_soc":                      # floating-point sum over 1 Mi cells, accumulated with atomics: order is not fixed
            assert abs(s[k] - s2[k]) < 1e-6
        else:
            assert s[k] == s2[k], k
    g2 = cb.logs()
    assert np.array_equal(g["u"], g2["u"]) and np.array_equal(g["active"], g2["active"])


def test_c5_heterogeneous_population_at_scale(bank, coracle_mod):
    """262 144 heterogeneous cells (per-cell limits, T in 15..35 C): sub-sample parity and summary consistency."""
    from mek import api, population as P
    B = 1 << 18
    pop = P.make_population(bank.params[2], B, seed=4, heterogeneous=True)
    rom = api.RomBank(bank)
    cb = api.CellBatch(rom, B)
    cb.set_logging(256, 2000)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    cb.run(0, 2000)
    g = cb.logs()
    sub = pop.slice(0, 256)
    co = coracle_mod.COracle(bank)
    co.init(sub.SOC0, sub.T, sub.SigmaX0, sub.limits, sub.ekf_SOC0)
    c = co.run(0, 2000, nthreads=8)
    for k in ("u", "V", "SOC", "phise", "iT", "iZ", "nexec", "active"):
        assert np.array_equal(g[k], c[k]), k
    s = cb.summary()
    assert s["cells"] == B and s["n_lost"] == 0 and s["n_nonfinite"] == 0
    assert len(np.unique(g["iT"])) >= 2                               # the temperature axis of the bank is exercised
    # per-cell limits are honoured: the command never goes below each cell's own u_min (to QP tolerance)
    assert np.all(g["u"] >= sub.limits[:, 1][None, :] - 1e-3)
