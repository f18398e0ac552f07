"""Throughput against the deferral block length (shared-memory history columns vs L1 size).  usage: defer_l1_probe.py [CELLS]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
from mek import api, bank as B, population as P
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
pop = P.make_population(bk.params[2], cells, seed=2)
for kb in (16, 20, 24, 32, 16):
    cb = api.CellBatch(rom, cells, defer_block=kb)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits)
    cb.run(0, 200); cb.reset()
    ms = cb.run_timed(0, 3001)
    print("defer_block %2d (in use %d): %.1f ms  %.2f M cell-steps/s" % (kb, cb.defer_block, ms, cells * 3001 / ms / 1e3), flush=True)
    cb.close()
