"""Timing of the fused loop under different knobs on one population (full charge).
usage: fused_probe.py CELLS NSIM "k=v,k=v" ["k=v" ...]   each argument is one configuration of CellBatch keyword arguments
env: MEK_LIB_PATH = alternative libmek build, MEK_HET=1 heterogeneous population (C5), MEK_CONS=1,1,0"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
import numpy as np
from mek import api, bank as B, population as P
cells, nsim = int(sys.argv[1]), int(sys.argv[2])
cfgs = sys.argv[3:] or [""]
het = os.environ.get("MEK_HET", "0") == "1"
cons = tuple(int(c) for c in os.environ.get("MEK_CONS", "1,1,1").split(","))
Np, Nc = int(os.environ.get("MEK_NP", "5")), int(os.environ.get("MEK_NC", "2"))
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
pop = P.make_population(bk.params[2], cells, seed=4 if het else (3 if Np > 5 else 2), heterogeneous=het)
ref = None
for cfg in cfgs:
    kw = {k: int(v) for k, v in (kv.split("=") for kv in cfg.split(",") if kv)}
    cb = api.CellBatch(rom, cells, Np=Np, Nc=Nc, constraints=cons, **kw)
    cb.set_logging(min(cells, 1024), nsim)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits)
    cb.run(0, min(200, nsim))            # warm-up (clocks, caches), then the timed full charge from a fresh state
    cb.reset()
    ms = cb.run_timed(0, nsim)
    s = cb.summary()
    lg = cb.logs(0, nsim)
    sig = (float(np.nansum(lg["u"])), int(lg["nexec"].sum()), s["sum_nexec"], s["min_phise"])
    if ref is None:
        ref = sig
    print("%-40s %9.1f ms  %7.1f M cell-steps/s  sweeps %.2f  same_as_first %s" %
          (cfg or "(default)", ms, cells * nsim / ms / 1e3, s["sum_nexec"] / s["cellsteps"], sig == ref), flush=True)
    cb.close()
