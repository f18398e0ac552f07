import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "mpc-ekf4fastcharge_b200"))
import numpy as np
from mek import api, bank as B, population as P
from oracle import coracle
bank = B.Bank.load(B.default_bank_path())
Bn = 32
pop = P.make_population(bank.params[2], Bn, seed=11, heterogeneous=True)
rom = api.RomBank(bank); cb = api.CellBatch(rom, Bn)
cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
co = coracle.COracle(bank); co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
rng = np.random.default_rng(0)
names = ['lost','iT','iZ'] + ['xhat%d'%i for i in range(6)] + ['a%d'%i for i in range(5)] + ['Cv%d'%i for i in range(5)] + ['Dv'] + ['Cphi%d'%i for i in range(5)] + ['Dphi','bv','bphi','vhat','soc','rsvd']
nbad = 0
for k in range(25):
    I = rng.uniform(-60, 5, Bn)
    V = cb.ob_step(I)
    Vc = np.array([co.ob_step(c, I[c]) for c in range(Bn)])
    zk, mats = cb.iter_ekf(V, I)
    for c in range(Bn):
        zkc, matsc = co.iter_ekf(c, Vc[c], I[c])
        bad = [(names[i], mats[c][i], matsc[i]) for i in range(31) if not (mats[c][i] == matsc[i] or (np.isnan(mats[c][i]) and np.isnan(matsc[i])))]
        if bad and nbad < 6:
            nbad += 1
            print("step", k, "cell", c, "T", pop.T[c], bad[:6], "zk equal", np.array_equal(zk[c], zkc, equal_nan=True))
print("done, mismatching (step,cell) printed:", nbad)
