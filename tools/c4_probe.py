"""C4 (Np=20, Nc=5) throughput of the long-horizon instance at different CTA sizes: usage c4_probe.py CELLS NSIM [threads,...]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
import numpy as np
from mek import api, bank as B, population as P
cells = int(sys.argv[1]); nsim = int(sys.argv[2])
tp = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [64, 128, 256]
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
pop = P.make_population(bk.params[2], cells, seed=3)
ref = None
for t in tp:
    cb = api.CellBatch(rom, cells, Np=20, Nc=5, threads_per_block=t, generic_path=(t == 0))
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits); cb.sync()
    ms = cb.run_timed(0, nsim)
    s = cb.summary()
    st = cb.read_state(); st[:, 29] = 0
    same = "ref" if ref is None else ("state bit-identical" if np.array_equal(st, ref, equal_nan=True) else "STATE DIFFERS")
    if ref is None: ref = st
    print("threads %3d: %.1f ms  %.2f M cell-steps/s  mean sweeps %.2f  %s" % (t, ms, cells * nsim / ms / 1e3, s["sum_nexec"] / s["cellsteps"], same), flush=True)
    cb.close()
