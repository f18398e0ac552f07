"""A/B of the deferral block length on one population: usage defer_probe.py CELLS NSIM [K,K,...] [threads]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
import numpy as np
from mek import api, bank as B, population as P
cells = int(sys.argv[1]); nsim = int(sys.argv[2])
ks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [-1, 8, 16, 32]
tpb = int(sys.argv[4]) if len(sys.argv) > 4 else 0
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
pop = P.make_population(bk.params[2], cells, seed=2)
ref = None
for k in ks:
    for stag in ((False, True) if k > 0 else (False,)):
        cb = api.CellBatch(rom, cells, defer_block=k, defer_stagger=stag, threads_per_block=tpb)
        cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits); cb.sync()
        cb.run_timed(0, 64)
        cb.reset(); cb.sync()
        ms = cb.run_timed(0, nsim)
        s = cb.summary()
        st = cb.read_state()
        same = "ref" if ref is None else ("state bit-identical" if np.array_equal(st, ref, equal_nan=True) else "STATE DIFFERS (max abs %.3e)" % np.nanmax(np.abs(st - ref)))
        if ref is None: ref = st
        print("defer_block %3d stagger %d: %.1f ms  %.2f M cell-steps/s  mean sweeps %.3f  %s" % (k, stag, ms, cells * nsim / ms / 1e3, s["sum_nexec"] / s["cellsteps"], same), flush=True)
        cb.close()
