import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'mpc-ekf4fastcharge_b200'))
import numpy as np
from mek import api, bank as B, population as P
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
for nb in (4096, 65536, 262144):
    pop = P.make_population(bk.params[2], nb, seed=1)
    cb = api.CellBatch(rom, nb)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits); cb.sync()
    ms = cb.run_timed(0, 300)
    cb.reset(); cb.sync()
    ms = cb.run_timed(0, 3001)
    s = cb.summary()
    print("B=%d: %.1f ms, %.2f M cell-steps/s, mean nexec %.2f, reached %d" % (nb, ms, nb*3001/ms/1e3, s['sum_nexec']/s['cellsteps'], s['n_reached']), flush=True)
    cb.close()
