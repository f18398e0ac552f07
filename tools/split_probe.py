"""Timing of the phase-split form against the fused kernel on one population.
usage: split_probe.py CELLS K_WARM K_END [est,mpc ...]   (thread pairs; 0,0 = fused)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
import numpy as np
from mek import api, bank as B, population as P
cells, kw, ke = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pairs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[4:]] or [(0, 0), (256, 256)]
het = os.environ.get("MEK_HET", "0") == "1"
cons = tuple(int(c) for c in os.environ.get("MEK_CONS", "1,1,1").split(","))
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
pop = P.make_population(bk.params[2], cells, seed=4 if het else 2, heterogeneous=het)
ref = None
for te, tm in pairs:
    cb = api.CellBatch(rom, cells, constraints=cons, split_phases=(1 if te else -1), split_threads_est=te, split_threads_mpc=tm)
    cb.set_logging(min(cells, 2048), ke)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits)
    if kw:
        cb.run(0, kw)
    ms = cb.run_timed(kw, ke)
    s = cb.summary()
    lg = cb.logs(0, ke)
    sig = (float(np.nansum(lg["u"])), int(lg["nexec"].sum()), s["sum_nexec"], s["min_phise"])
    if ref is None:
        ref = sig
    print("est %3d mpc %3d split %d: %8.1f ms  %7.1f M cell-steps/s  sweeps %.2f  same_as_first %s" %
          (te, tm, cb.split, ms, cells * (ke - kw) / ms / 1e3, s["sum_nexec"] / s["cellsteps"], sig == ref), flush=True)
    cb.close()
