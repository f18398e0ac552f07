"""A/B of the state header in shared memory (mek_config.smem_header).  usage: hdr_probe.py CELLS"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
from mek import api, bank as B, population as P
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
pop = P.make_population(bk.params[2], cells, seed=2)
for hdr, kb in ((0, 0), (1, 0), (0, 12), (1, 0)):
    cb = api.CellBatch(rom, cells, smem_header=hdr, defer_block=kb)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits)
    cb.run(0, 200); cb.reset()
    ms = cb.run_timed(0, 3001)
    print("smem_header %2d defer_block %2d: %.1f ms  %.2f M cell-steps/s" % (hdr, cb.defer_block, ms, cells * 3001 / ms / 1e3), flush=True)
    cb.close()
