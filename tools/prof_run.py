"""Short driver for ncu captures: one population, one full charge.  usage: prof_run.py CELLS NSIM [NP NC]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
from mek import api, bank as B, population as P
cells = int(sys.argv[1]); nsim = int(sys.argv[2])
Np = int(sys.argv[3]) if len(sys.argv) > 3 else 5
Nc = int(sys.argv[4]) if len(sys.argv) > 4 else 2
bps = int(sys.argv[5]) if len(sys.argv) > 5 else 0
tpb = int(sys.argv[6]) if len(sys.argv) > 6 else 0
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
het = os.environ.get("MEK_HET", "0") == "1"
pop = P.make_population(bk.params[2], cells, seed=4 if het else 2, heterogeneous=het)
cb = api.CellBatch(rom, cells, Np=Np, Nc=Nc, blocks_per_sm=0, threads_per_block=(bps if bps > 100 else 0), generic_path=(bps == 9), sync_steps=(tpb != 2))
cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits)
ksplit = int(sys.argv[7]) if len(sys.argv) > 7 else 0
if ksplit:
    cb.run(0, ksplit)          # first launch: bring the population to sample instant ksplit (skipped by ncu -s 1)
ms = cb.run_timed(ksplit, nsim)
s = cb.summary()
print("bps %d tpb %d" % (bps, tpb), "cells %d nsim %d: %.1f ms  %.2f M cell-steps/s  mean sweeps %.2f" % (cells, nsim, ms, cells * (nsim - ksplit) / ms / 1e3, s["sum_nexec"] / s["cellsteps"]))
