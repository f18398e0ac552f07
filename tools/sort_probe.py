"""Probe: how much does ordering the cells by initial SOC (so that a warp's 32 cells are in the same charge phase) buy?"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
import numpy as np
from mek import api, bank as B, population as P
cells = int(sys.argv[1]); nsim = 3001
bk = B.Bank.load(B.default_bank_path()); rom = api.RomBank(bk)
pop = P.make_population(bk.params[2], cells, seed=2)
for mode in ("random", "sorted"):
    idx = np.argsort(pop.SOC0, kind="stable") if mode == "sorted" else np.arange(cells)
    cb = api.CellBatch(rom, cells)
    cb.init(pop.SOC0[idx], pop.T[idx], pop.SigmaX0[idx], pop.limits[idx])
    cb.run(0, 200); cb.reset(); cb.sync()
    ms = cb.run_timed(0, nsim)
    print(mode, "%.1f ms %.2f M cell-steps/s" % (ms, cells * nsim / ms / 1e3), flush=True)
    cb.close()
