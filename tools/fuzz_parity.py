"""Randomised parity sweep: CUDA path vs the C oracle over several seeds / configurations (bit-exact).  usage: fuzz_parity.py [NSEEDS [NP NC [MAXCELLS MAXN]]]
(NP NC = 20 5 runs the long-horizon static-shape instance)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
import numpy as np
from mek import api, bank as B, population as P
from oracle import coracle
coracle.build()
bk = B.Bank.load(B.default_bank_path())
rom = api.RomBank(bk)
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
Np, Nc = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (5, 2)
maxcells, maxn = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (900, 2400)
bad = 0
for seed in range(100, 100 + nseeds):
    rng = np.random.default_rng(seed)
    cells = int(rng.integers(200, maxcells)); N = int(rng.integers(600, maxn))
    cons = [(1, 1, 1), (1, 1, 0)][seed % 2]
    pop = P.make_population(bk.params[2], cells, seed=seed, heterogeneous=bool(seed % 3), ekf_soc_error=float(rng.uniform(0, 4)))
    if seed % 4 == 0:
        pop.SOC0[::7] = rng.uniform(80, 94, len(pop.SOC0[::7])); pop.ekf_SOC0[::7] = pop.SOC0[::7] + 0.5
    cb = api.CellBatch(rom, cells, Np=Np, Nc=Nc, constraints=cons)
    cb.set_logging(cells, N)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    cuts = sorted(set([0, N] + [int(x) for x in rng.integers(1, N, 2)]))
    for a, b in zip(cuts[:-1], cuts[1:]):
        cb.run(a, b)
    g = cb.logs(0, N)
    co = coracle.COracle(bk, Np=Np, Nc=Nc, constraints=cons)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    c = co.run(0, N, nthreads=os.cpu_count() or 8)
    ok = all(np.array_equal(g[k], c[k], equal_nan=k in ("u", "V", "SOC", "phise")) for k in ("u", "V", "SOC", "phise", "iT", "iZ", "nexec", "active"))
    s = cb.summary()
    print("Np %d Nc %d seed %d cells %d N %d cons %s cuts %s: %s  (cap hits %d, mean sweeps %.2f)" % (Np, Nc, seed, cells, N, cons, cuts, "bit-exact" if ok else "MISMATCH", s["n_cap"], s["sum_nexec"] / s["cellsteps"]), flush=True)
    bad += (not ok)
    cb.close()
print("mismatches:", bad)
sys.exit(1 if bad else 0)
