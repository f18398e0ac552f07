"""A/B of the lean vs general Hildreth sweep (test hook general_sweep_test).  usage: lean_probe.py CELLS [NSIM]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
from mek import api, bank as B, population as P
cells = int(sys.argv[1]); nsim = int(sys.argv[2]) if len(sys.argv) > 2 else 3001
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
pop = P.make_population(bk.params[2], cells, seed=2)
for hook in (True, False, True, False):
    cb = api.CellBatch(rom, cells, general_sweep_test=hook)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits)
    cb.run(0, 200); cb.reset()
    ms = cb.run_timed(0, nsim)
    s = cb.summary()
    print("general" if hook else "lean   ", "cells %d nsim %d: %.1f ms  %.2f M cell-steps/s  mean sweeps %.2f" % (cells, nsim, ms, cells * nsim / ms / 1e3, s["sum_nexec"] / s["cellsteps"]), flush=True)
    cb.close()
