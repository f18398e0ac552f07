"""C5 (heterogeneous population) throughput against the regrouping period: usage c5_probe.py CELLS NSIM [every,...]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
import numpy as np
from mek import api, bank as B, population as P
cells = int(sys.argv[1]); nsim = int(sys.argv[2])
ev = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [-1, 256, 128, 64, 32]
het = os.environ.get("MEK_HET", "1") == "1"
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
pop = P.make_population(bk.params[2], cells, seed=4 if het else 2, heterogeneous=het)
ref = None
for e in ev:
    cb = api.CellBatch(rom, cells, regroup_every=e)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits); cb.sync()
    cb.run_timed(0, 64); cb.reset(); cb.sync()
    l0 = cb.launches
    ms = cb.run_timed(0, nsim)
    s = cb.summary()
    st = cb.read_state(); st[:, 29] = 0
    same = "ref" if ref is None else ("state bit-identical" if np.array_equal(st, ref, equal_nan=True) else "STATE DIFFERS")
    if ref is None: ref = st
    print("regroup_every %4d: %.1f ms  %.2f M cell-steps/s  mean sweeps %.2f  launches %d  %s" % (e, ms, cells * nsim / ms / 1e3, s["sum_nexec"] / s["cellsteps"], cb.launches - l0, same), flush=True)
    cb.close()
