"""Generates tests/golden/*.npz with the literal NumPy oracle (oracle/literal.py).

The reference has no golden vectors and cannot be executed here, so these
fixtures pin the *restatement*: they freeze the literal oracle's trajectories for a list of
(config, cell) cases so that the C oracle, the CUDA path and any later edit of the literal oracle
are all checked against the same numbers.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mpc-ekf4fastcharge_b200"))

from mek import bank as B, population as P  # noqa: E402
from oracle import literal as L  # noqa: E402

CASES = [
    # name, population kind, seed, cell ids, Np, Nc, constraints, N, ekf_soc_error
    ("c1_111", "single", 0, [0], 5, 2, (1, 1, 1), 3001, 0.0),
    ("c1_110", "single", 0, [0], 5, 2, (1, 1, 0), 3001, 0.0),
    ("c2_seed1", "hom", 1, [0, 1, 2, 3], 5, 2, (1, 1, 1), 3001, 0.0),
    ("c5_seed4", "het", 4, [0, 1, 2, 3], 5, 2, (1, 1, 1), 3001, 5.0),
    ("c4_seed3", "hom", 3, [0], 20, 5, (1, 1, 1), 1500, 0.0),
]


def run_case(bank, case):
    name, kind, seed, cells, Np, Nc, cons, N, err = case
    Q = bank.params[2]
    ncell = max(cells) + 1
    pop = P.single_cell(Q) if kind == "single" else P.make_population(Q, ncell, seed, heterogeneous=(kind == "het"), ekf_soc_error=err)
    out = dict(SOC0=pop.SOC0[cells], ekf_SOC0=pop.ekf_SOC0[cells], T=pop.T[cells], SigmaX0=pop.SigmaX0[cells], limits=pop.limits[cells],
               Np=Np, Nc=Nc, constraints=np.array(cons), N=N)
    keys = ("u", "V", "SOC", "phise", "iT", "iZ", "nexec", "active")
    acc = {k: [] for k in keys}
    for i in cells:
        lim = pop.limits[i]
        o = L.run_mpc(bank, SOC0=pop.SOC0[i], TC=pop.T[i], N=N, Np=Np, Nc=Nc, constraints=cons, Crate=-lim[1] / Q,
                      limits=dict(u_max=lim[0], du_max=lim[2], du_min=lim[3], v_max=lim[4], phise_min=lim[5], z_max=lim[6]),
                      SigmaX0=pop.SigmaX0[i], ekf_SOC0=pop.ekf_SOC0[i])
        for k in keys:
            acc[k].append(o[k])
        print(name, "cell", i, "final SOC %.4f" % o["SOC"][-1], "mean nexec %.2f" % o["nexec"].mean(), flush=True)
    for k in keys:
        out[k] = np.stack(acc[k], axis=1)        # [step][cell](...)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    bank = B.Bank.load(B.default_bank_path())
    only = sys.argv[1:]
    for case in CASES:
        if not only or case[0] in only:
            run_case(bank, case)
