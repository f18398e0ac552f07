"""Where the end-to-end leg of bench.py spends its time.  usage: e2e_probe.py [CELLS]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
import numpy as np, torch
from mek import api, bank as B, population as P
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
pop = P.make_population(bk.params[2], cells, seed=2)
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
h = [pin(pop.SOC0), pin(pop.ekf_SOC0), pin(pop.T), pin(np.swapaxes(pop.SigmaX0, 1, 2)), pin(pop.limits)]
cb = api.CellBatch(rom, cells)
cb.set_logging(4096, 3001)
for it in range(3):
    t0 = time.perf_counter(); cb.init_raw(*[t.data_ptr() for t in h]); cb.sync(); t1 = time.perf_counter()
    cb.run(0, 3001, sync=False); cb.sync(); t2 = time.perf_counter()
    s = cb.summary(); t3 = time.perf_counter()
    lg = cb.logs(0, 3001, reuse=True); t4 = time.perf_counter()
    print("init %.1f ms  run %.1f ms  summary %.1f ms  logs %.1f ms  total %.1f ms" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)), flush=True)
