"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

NumPy/libm evaluation of the synthetic ``cellData.function.*`` closed forms, presented with the
call surface the reference uses on MATLAB function handles:
``fn.neg.Uocp(theta, T)``, ``fn.pos.soc(z, T)``, ``fn.const.Q()`` ...  Uses numpy's own
exp/log/sqrt (independent of the in-house device math in csrc/mek_math.cuh).
"""
from types import SimpleNamespace

import numpy as np

# parameter-vector layout (restated here; must equal mek/bank.py and csrc/mek_params.h)
P_F, P_R, P_Q, P_RC, P_TREF = 0, 1, 2, 3, 4
P_NEG, P_POS = 5, 16
(E_THETA0, E_THETA100, E_K0REF, E_K0SLOPE, E_EA_K0, E_RFREF, E_RFSLOPE, E_EA_RF,
 E_WDL, E_CDL, E_NDL) = range(11)
P_UN, P_UP, P_UP_COEF = 27, 33, 36


def _arr(p, T, Ea):
    return np.exp(Ea / p[P_R] * (1.0 / p[P_TREF] - 1.0 / T))


def _electrode(p, base, uocp, duocp):
    e = SimpleNamespace()
    th0, th100 = p[base + E_THETA0], p[base + E_THETA100]
    e.theta0 = lambda: th0
    e.theta100 = lambda: th100
    e.soc = lambda z, T=None: th0 + z * (th100 - th0)
    e.Uocp = lambda th, T=None: uocp(th)
    e.dUocp = lambda th, T=None: duocp(th)
    e.k0 = lambda th, T: p[base + E_K0REF] * (1.0 + p[base + E_K0SLOPE] * (th - 0.5)) * _arr(p, T, p[base + E_EA_K0])
    e.Rf = lambda th, T: p[base + E_RFREF] * (1.0 + p[base + E_RFSLOPE] * th) * _arr(p, T, -p[base + E_EA_RF])
    e.wDL = lambda th, T: p[base + E_WDL]
    e.Cdl = lambda th, T: p[base + E_CDL]
    e.nDL = lambda: p[base + E_NDL]
    return e


def _interp1(grid_lo, grid_hi, y):
    """MATLAB's interp1(theta, y, x, 'linear') on a uniform grid with x clamped to the grid (the tabulated form of a handle)."""
    n = len(y)
    def f(th, T=None):
        th = float(th)
        if th != th:
            return th
        t = (th - grid_lo) * ((n - 1) / (grid_hi - grid_lo))
        t = min(max(t, 0.0), float(n - 1))
        i = min(int(t), n - 2)
        return y[i] + (t - i) * (y[i + 1] - y[i])
    return f


def make_cell_data(params, lut=None, lut_range=(0.0, 1.0)):
    """lut: optional [8,n] tables (order: Uocp_neg, dUocp_neg, Uocp_pos, dUocp_pos, k0_neg, k0_pos, Rf_neg, Rf_pos) replacing the
    closed forms of those handles; k0 / Rf tables are at the reference temperature and keep the Arrhenius factor."""
    p = np.asarray(params, dtype=np.float64)
    c = p[P_UN:P_UN + 6]

    def un(th):
        x = (th - c[4]) / c[5]
        return c[0] + c[1] / ((th + c[2]) * (th + c[2])) + c[3] * (1.0 - x / np.sqrt(1.0 + x * x))

    def dun(th):
        x = (th - c[4]) / c[5]
        q = 1.0 + x * x
        return -2.0 * c[1] / ((th + c[2]) ** 3) - (c[3] / c[5]) / (q * np.sqrt(q))
    pc, pw, nc = p[P_UP], p[P_UP + 1], int(p[P_UP + 2])
    coef = p[P_UP_COEF:P_UP_COEF + nc]

    def up(th):
        x = (th - pc) / pw
        return np.polynomial.polynomial.polyval(x, coef)

    def dup(th):
        x = (th - pc) / pw
        return np.polynomial.polynomial.polyval(x, np.polynomial.polynomial.polyder(coef)) / pw
    if lut is not None:
        T = [_interp1(lut_range[0], lut_range[1], np.asarray(row, dtype=np.float64)) for row in lut]
        un, dun, up, dup = T[0], T[1], T[2], T[3]
    cd = SimpleNamespace()
    cd.const = SimpleNamespace(F=p[P_F], R=p[P_R])
    cd.function = SimpleNamespace(
        const=SimpleNamespace(Q=lambda *_: p[P_Q], Rc=lambda *_: p[P_RC]),
        neg=_electrode(p, P_NEG, un, dun),
        pos=_electrode(p, P_POS, up, dup),
    )
    if lut is not None:
        for e, base, k0t, rft in ((cd.function.neg, P_NEG, T[4], T[6]), (cd.function.pos, P_POS, T[5], T[7])):
            e.k0 = (lambda k0t, base: lambda th, Tk: k0t(th) * _arr(p, Tk, p[base + E_EA_K0]))(k0t, base)
            e.Rf = (lambda rft, base: lambda th, Tk: rft(th) * _arr(p, Tk, -p[base + E_EA_RF]))(rft, base)
    return cd
