"""TEST / MEASUREMENT INFRASTRUCTURE (oracle) -- exact FP64 operation tallies of the control loop.

Runs the op-counting build of the C restatement (oracle/opcount.cpp) on a sample of each named configuration and
writes the per-block tallies SURVEY.md 8(d) asks for (flops per cell-step, per QP and per Hildreth sweep) to
profiles/r01_flop_tally.json.  bench.py reads that JSON (it never executes this file).

    python -m oracle.opcount_run            # from the repo root
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mpc-ekf4fastcharge_b200"))

BLOCKS = ["plant OB_step/simStep", "EKF time update (all M models)", "EKF measurement path (getXind x2, getVariables x2, getChatV, gains, 4 updates)",
          "PSD check / repair (iterEKF.m:143-145)", "EKFmatsHandler", "iterMPC fixed part (predMat SOC row, F, G'G, Ru, E, solve, violation test)",
          "constraintsMPC (+ predMat V and phise rows)", "QP set-up + recovery (inv(E), E\\M', H_ii, K, dU)", "Hildreth sweeps", "other"]
KINDS = ["add", "mul", "fma", "div", "sqrt", "cmp"]
FLOPS = np.array([1, 1, 2, 1, 1, 0])          # add/mul = 1, FMA = 2, div/sqrt = 1; comparisons free
INSTR = np.array([1, 1, 1, 1, 1, 0])          # FP64-pipe instructions (one DFMA is one issue slot)

CONFIGS = {
    "C3":  dict(seed=2, Np=5, Nc=2, constraints=(1, 1, 1), het=False, cells=64),
    "C3v": dict(seed=2, Np=5, Nc=2, constraints=(1, 1, 0), het=False, cells=64),
    "C4":  dict(seed=3, Np=20, Nc=5, constraints=(1, 1, 1), het=False, cells=16),
    "C5":  dict(seed=4, Np=5, Nc=2, constraints=(1, 1, 1), het=True, cells=64),
}


def count_lib():
    so = os.path.join(HERE, "_build", "libmek_oracle_count.so")
    srcs = [os.path.join(HERE, f) for f in ("opcount.cpp", "mek_oracle.c")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s", "_build/libmek_oracle_count.so"])
    return so


def tally(bank, cfg, nsim=3001):
    from mek import population as P
    from oracle import coracle
    coracle._SO, coracle._lib = count_lib(), None          # bind the counting build instead of the timed one
    try:
        L = coracle.lib()
        pop = P.make_population(bank.params[2], cfg["cells"], cfg["seed"], heterogeneous=cfg["het"])
        co = coracle.COracle(bank, Np=cfg["Np"], Nc=cfg["Nc"], constraints=cfg["constraints"])
        co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
        L.mko_set_zero_weight_skip(0)                       # the table counts the reference's algorithm term by term
        L.mko_opc_reset()
        lg = co.run(0, nsim, logs=True, nthreads=1)
        raw = np.zeros((len(BLOCKS), len(KINDS)), np.uint64)
        L.mko_opc_get(raw.ctypes.data_as(C.c_void_p))
    finally:
        coracle._lib = None
        coracle._SO = os.path.join(HERE, "_build", "libmek_oracle_fma.so" if coracle._has_fma() else "libmek_oracle.so")
    raw = raw.astype(np.float64)
    steps = float(cfg["cells"]) * nsim
    n_qp = float((lg["nexec"] > 0).sum())
    sweeps = float(lg["nexec"].sum())
    fl = raw @ FLOPS
    ins = raw @ INSTR
    out = {"cells": cfg["cells"], "nsim": nsim, "cellsteps": steps, "n_qp": n_qp, "sum_nexec": sweeps,
           "mean_sweeps_per_step": sweeps / steps, "nC": co.nC,
           "blocks": [{"block": BLOCKS[b], "flops_per_cellstep": fl[b] / steps, "fp64_instr_per_cellstep": ins[b] / steps,
                       **{k: raw[b, i] / steps for i, k in enumerate(KINDS)}} for b in range(len(BLOCKS))],
           "F_fixed": float(fl[:7].sum() + fl[9]) / steps,            # everything but the QP itself
           "F_qp_setup": float(fl[7]) / max(n_qp, 1.0),
           "F_sweep": float(fl[8]) / max(sweeps, 1.0),
           "flops_per_cellstep": float(fl.sum()) / steps,
           "fp64_instr_per_cellstep": float(ins.sum()) / steps}
    return out


def main():
    from mek import bank as Bk
    bank = Bk.Bank.load(Bk.default_bank_path())
    res = {"rule": "add/sub = 1, mul = 1, fma = 2, div = 1, sqrt = 1; negation, fabs, floor, comparisons, selects free",
           "source": "oracle/opcount.cpp: oracle/mek_oracle.c compiled with a counting scalar type; full 3001-instant charges",
           "bank": "canonical synthetic bank (n = 5, nz = 30, M = 30)", "configs": {}}
    for name, cfg in CONFIGS.items():
        res["configs"][name] = tally(bank, cfg)
        r = res["configs"][name]
        print("%-4s F_fixed %.1f  F_qp_setup %.1f  F_sweep %.1f  sweeps/step %.2f  flops/cell-step %.1f  fp64 instr/cell-step %.1f"
              % (name, r["F_fixed"], r["F_qp_setup"], r["F_sweep"], r["mean_sweeps_per_step"], r["flops_per_cellstep"], r["fp64_instr_per_cellstep"]))
    with open(os.path.join(ROOT, "profiles", "r01_flop_tally.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
