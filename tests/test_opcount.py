"""The frozen flop table (profiles/r01_flop_tally.json) against a fresh count on a small case,
and the op-counting build of the oracle against the timed build (same bits: counting must not change the arithmetic)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_counting_build_reproduces_the_oracle_and_the_frozen_table(bank, coracle_mod):
    from mek import population as P
    from oracle import opcount_run
    frozen = json.load(open(os.path.join(ROOT, "profiles", "r01_flop_tally.json")))["configs"]
    for name, cells, nsim in (("C3", 3, 400), ("C3v", 2, 300), ("C4", 1, 120)):
        cfg = dict(opcount_run.CONFIGS[name], cells=cells)
        t = opcount_run.tally(bank, cfg, nsim=nsim)
        f = frozen[name]
        # structural constants: the fixed part, the QP set-up and one sweep cost the same number of flops on every step
        assert t["nC"] == f["nC"]
        assert abs(t["F_sweep"] - f["F_sweep"]) < 1e-9
        assert abs(t["F_qp_setup"] - f["F_qp_setup"]) < 1e-9
        assert abs(t["F_fixed"] - f["F_fixed"]) < 1e-2 * f["F_fixed"]   # the first instants of a charge (no QP yet, ramp) weigh more in a short run
        assert abs(t["flops_per_cellstep"] - (t["F_fixed"] + t["n_qp"] / t["cellsteps"] * t["F_qp_setup"]
                                               + t["mean_sweeps_per_step"] * t["F_sweep"])) < 1e-6
    # the counting build computes the same bits as the timed build
    pop = P.make_population(bank.params[2], 3, 2)
    co = coracle_mod.COracle(bank)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    ref = co.run(0, 400)
    coracle_mod._SO, coracle_mod._lib = opcount_run.count_lib(), None
    try:
        cc = coracle_mod.COracle(bank)
        cc.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
        got = cc.run(0, 400)
    finally:
        coracle_mod._lib = None
        coracle_mod._SO = os.path.join(os.path.dirname(coracle_mod.__file__), "_build",
                                       "libmek_oracle_fma.so" if coracle_mod._has_fma() else "libmek_oracle.so")
    for k in ("u", "V", "SOC", "phise", "iT", "iZ", "nexec", "active"):
        assert np.array_equal(ref[k], got[k]), k
