"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Literal NumPy restatement of the reference's hot path, function by function, dense algebra,
LAPACK svd/solve, all quirks of SURVEY.md Appendix B kept (Q1..Q24).  One cell, slow, trusted.

PARITY UNPINNED: the reference ships no tests/golden vectors and neither MATLAB nor Octave nor
the ROM .mat file exist in this environment.  This file follows the
reference source line by line; each function cites the lines it restates.  Indices are 0-based
here (MATLAB: 1-based); ``iT``/``iZ`` in logs are reported 1-based like MPC.iT/MPC.iZ.

Files restated (under /root/reference/src):
  MPC-EKF4FastCharge/OB_step.m, UTILITY/initKF.m, UTILITY/iterEKF.m (OB branches),
  MPC-EKF4FastCharge/EKFmatsHandler.m, predMat.m, constraintsMPC.m, iterMPC.m, hildreth.m,
  initMPC.m, runMPC.m (loop, order, logging).
"""
from __future__ import annotations

import copy
from types import SimpleNamespace as NS

import numpy as np

from . import cellfn

EPS = np.finfo(float).eps


# ------------------------------------------------------------------------------------------
# L0: MATLAB-shaped ROM struct from a bank
# ------------------------------------------------------------------------------------------
def build_rom(bank):
    """ROM struct with the field surface of ROM_NMC30_HRA.mat."""
    mdls = []
    for it in range(bank.nT):
        row = []
        for iz in range(bank.nZ):
            row.append(NS(A=np.diag(bank.A[it, iz]).copy(), B=np.ones((bank.n + 1, 1)),
                          C=bank.C[it, iz].copy(), D=bank.D[it, iz].reshape(-1, 1).copy(),
                          T=float(bank.T_K[it, iz]), SOC=float(bank.Z[it, iz])))
        mdls.append(row)
    return NS(ROMmdls=mdls,
              tfData=NS(names=list(bank.names), xLoc=np.asarray(bank.xloc, float).copy()),
              xraData=NS(T=bank.Tpts_C.copy(), SOC=bank.Zpts_pct.copy(), Tsamp=float(bank.Tsamp)),
              cellData=cellfn.make_cell_data(bank.params, getattr(bank, "lut", None), getattr(bank, "lut_range", (0.0, 1.0))))


def _find(names, nm):
    return [i for i, s in enumerate(names) if s == nm]


def _msort(v):
    """MATLAB sort: ascending, stable, NaN last -> indices."""
    return np.argsort(np.asarray(v, float), kind="stable")


# ------------------------------------------------------------------------------------------
# OB_step.m
# ------------------------------------------------------------------------------------------
def OB_step(Iapp, Tc, cellState, ROM, initCfg=None):
    """[Vcell, obs, cellState] = OB_step(Iapp, Tc, cellState, ROM, initCfg)   (OB_step.m:1-80)"""
    if cellState is None or not getattr(cellState, "initialized", False):      # :39
        if initCfg is None or "SOC0" not in initCfg:                           # :40
            raise ValueError("First call requires initCfg.SOC0 (in %).")       # :41
        cellState = NS()                                                       # :43
        cellState.initialized = True                                           # :44
        cellState.warnOff = bool(initCfg.get("warnOff", False))                # :45
        cellState.ROMmdls = copy.deepcopy(ROM.ROMmdls)                         # :48 (value copy)
        cellState.cellData = ROM.cellData                                      # :49
        cellState.Ts = ROM.xraData.Tsamp                                       # :50
        cellState.ROM = ROM                                                    # :53
        _ob_setupIndsLocs(cellState)                                           # :54
        _ob_setupBlend(cellState, ROM)                                         # :55-60
        Tk1 = Tc + 273.15                                                      # :63
        fn = cellState.cellData.function
        SOC0n = fn.neg.soc(initCfg["SOC0"] / 100, Tk1)                         # :64
        SOC0p = fn.pos.soc(initCfg["SOC0"] / 100, Tk1)                         # :65
        cellState.SOCnAvg = SOC0n; cellState.SOCpAvg = SOC0p                   # :66-67
        cellState.SOC0n = SOC0n; cellState.SOC0p = SOC0p                       # :70-71
        cellState.warnings = 0
    Vcell, new, obs = _ob_simStep(Iapp, Tc + 273.15, cellState)                # :75
    cellState.bigX = new.bigX                                                  # :78-80
    cellState.SOCnAvg = new.SOCnAvg
    cellState.SOCpAvg = new.SOCpAvg
    return Vcell, obs, cellState


def _ob_setupIndsLocs(cs):
    """OB_step.m:86-163"""
    names = cs.ROM.tfData.names; locs = np.asarray(cs.ROM.tfData.xLoc, float)            # :87-89
    cs.tfLocs = locs                                                                     # :90
    ind = {}; loc = {}
    for nm in ("negIfdl", "negIf", "negIdl", "negPhis", "negPhise", "negThetass", "posIfdl", "posIf",
               "posIdl", "posPhis", "posPhise", "posThetass", "negPhie", "sepPhie", "posPhie",
               "negThetae", "sepThetae", "posThetae"):
        ind[nm] = _find(names, nm)                                                       # :93-98,101-106,109-111,120-122
    loc["Phie"] = [locs[i] for i in ind["negPhie"] + ind["sepPhie"] + ind["posPhie"]]    # :112-114,117
    ind["Phie"] = ind["negPhie"] + ind["sepPhie"] + ind["posPhie"]                       # :116
    ind["Thetae"] = ind["negThetae"] + ind["sepThetae"] + ind["posThetae"]               # :127
    loc["Thetae"] = [locs[i] for i in ind["Thetae"]]                                     # :123-125,128

    def at(nm, x):
        return [i for i in _find(names, nm) if locs[i] == x]
    ind["negIfdl0"] = at("negIfdl", 0); ind["posIfdl3"] = at("posIfdl", 3)               # :131-137
    ind["negIf0"] = at("negIf", 0); ind["posIf3"] = at("posIf", 3)
    ind["negThetass0"] = at("negThetass", 0); ind["posThetass3"] = at("posThetass", 3)
    ind["negPhise0"] = at("negPhise", 0)
    if not ind["negIfdl0"]: raise ValueError("Requires ifdl at negative collector")      # :140-150
    if not ind["posIfdl3"]: raise ValueError("Requires ifdl at positive collector")
    if not ind["negIf0"]: raise ValueError("Requires if at negative collector")
    if not ind["posIf3"]: raise ValueError("Requires if at positive collector")
    if loc["Thetae"][0] > 0: raise ValueError("Requires thetae at negative collector")
    if loc["Thetae"][-1] > 3 + EPS or loc["Thetae"][-1] < 3 - EPS:
        raise ValueError("Requires thetae at positive collector")
    if not ind["negThetass0"]: raise ValueError("Requires thetass at negative collector")
    if not ind["posThetass3"]: raise ValueError("Requires thetass at positive collector")
    if not ind["negPhise0"]: raise ValueError("Requires phise at negative collector")
    if loc["Phie"][0] == 0:                                                              # :151-155
        ind["Phie"] = ind["Phie"][1:]; loc["Phie"] = loc["Phie"][1:]
    if loc["Phie"][-1] > 3 + EPS or loc["Phie"][-1] < 3 - EPS:                           # :156-158
        raise ValueError("Requires phie at positive collector")
    cs.ind = ind; cs.loc = loc                                                           # :161-162


def _ob_setupBlend(cs, ROM):
    """OB_step.m:165-186"""
    cs.Tspts = np.sort(np.asarray(ROM.xraData.T, float)) + 273.15                        # :166
    cs.Zspts = np.sort(np.asarray(ROM.xraData.SOC, float) / 100)                         # :167
    TT, ZZ = len(cs.ROMmdls), len(cs.ROMmdls[0])                                         # :168
    nA = cs.ROMmdls[0][0].A.shape[0]
    bigA = np.zeros((nA, TT * ZZ))                                                       # :171
    for tt in range(TT):                                                                 # :172-173
        for zz in range(ZZ):
            col = tt * ZZ + zz                                                           # :174
            R = cs.ROMmdls[tt][zz]                                                       # :175,182 (a reference here: stored back in place)
            bigA[:, col] = np.real(np.diag(R.A))                                         # :176
            R.C[cs.ind["negPhise"], -1] = 0                                              # :178
            if cs.ind["posPhise"]:
                R.C[cs.ind["posPhise"], -1] = 0                                          # :179-181
    cs.bigA = bigA; cs.bigX = np.zeros_like(bigA); cs.ZZ = ZZ                            # :185 (and OB_step.m:56-60)


def _ob_warn(cs, msg):
    """OB_step.m:359-370 (shortWarn of the plant)"""
    # not restated :360-368 (warnOff early-out -- runMPC.m:73 sets it --, the 'on'/'off' switch and the console print: only a count is kept here)
    cs.warnings += 1


def _ob_simStep(Iapp, T, cs):
    """OB_step.m:188-357"""
    # not restated :289-290,294-295,299-300 (the vector entries negIfdl/posIfdl, negIf/posIf, negIdl/posIdl of obs: runMPC reads nothing of obs)
    ind = cs.ind; Ts = cs.Ts; ROMmdls = cs.ROMmdls; bigX = cs.bigX                       # :190-195 (loc :191 is not used by simStep)
    Tspts, Zspts, ZZ = cs.Tspts, cs.Zspts, cs.ZZ                                         # :196-198
    cd = cs.cellData; fn = cd.function                                                   # :199
    SOC0n, SOC0p = cs.SOC0n, cs.SOC0p                                                    # :200-201
    F, R = cd.const.F, cd.const.R                                                        # :203-204
    Q = fn.const.Q(); Rc = fn.const.Rc()                                                 # :205-206
    theta0n, theta0p = fn.neg.theta0(), fn.pos.theta0()                                  # :207-208
    theta100n, theta100p = fn.neg.theta100(), fn.pos.theta100()                          # :209-210
    wDLn = fn.neg.wDL(SOC0n, T); wDLp = fn.pos.wDL(SOC0p, T)                             # :212-219
    Cdln = fn.neg.Cdl(SOC0n, T); Cdlp = fn.pos.Cdl(SOC0p, T)
    nDLn = fn.neg.nDL(); nDLp = fn.pos.nDL()
    Cdleffn = (Cdln ** (2 - nDLn)) * (wDLn ** (nDLn - 1))
    Cdleffp = (Cdlp ** (2 - nDLp)) * (wDLp ** (nDLp - 1))
    SOCnAvg = cs.SOCnAvg; SOCpAvg = cs.SOCpAvg                                           # :222-223
    obs = NS()
    obs.negSOC = SOCnAvg; obs.posSOC = SOCpAvg                                           # :226-228
    obs.cellSOC = (SOCnAvg - theta0n) / (theta100n - theta0n)
    dUocpnAvg = fn.neg.dUocp(SOCnAvg, T); dUocppAvg = fn.pos.dUocp(SOCpAvg, T)           # :231-236
    dQn = abs(theta100n - theta0n); dQp = abs(theta100p - theta0p)
    res0n = -dQn / (3600 * Q - Cdleffn * dQn * dUocpnAvg)
    res0p = dQp / (3600 * Q - Cdleffp * dQp * dUocppAvg)
    SOCnAvg = SOCnAvg + res0n * Iapp * Ts                                                # :239-244
    SOCpAvg = SOCpAvg + res0p * Iapp * Ts
    if SOCnAvg < 0: _ob_warn(cs, "Average SOCn < 0"); SOCnAvg = 0
    if SOCnAvg > 1: _ob_warn(cs, "Average SOCn > 1"); SOCnAvg = 1
    if SOCpAvg < 0: _ob_warn(cs, "Average SOCp < 0"); SOCpAvg = 0
    if SOCpAvg > 1: _ob_warn(cs, "Average SOCp > 1"); SOCpAvg = 1
    iZ = _msort(np.abs(obs.cellSOC - Zspts))                                             # :251-259
    if len(iZ) > 1:
        iZupper = max(iZ[0], iZ[1]); iZlower = min(iZ[0], iZ[1])
        Zupper = Zspts[iZupper]; Zlower = Zspts[iZlower]
    else:
        iZupper = iZlower = iZ[0]; Zupper = Zlower = Zspts[0]
    iT = _msort(np.abs(T - Tspts))                                                       # :261-269
    if len(iT) > 1:
        iTupper = max(iT[0], iT[1]); iTlower = min(iT[0], iT[1])
        Tupper = Tspts[iTupper]; Tlower = Tspts[iTlower]
    else:
        iTupper = iTlower = iT[0]; Tupper = Tlower = Tspts[0]

    def corner(it, iz):                                                                  # :272-275
        m = ROMmdls[it][iz]
        return m.C @ bigX[:, it * ZZ + iz] + m.D[:, 0] * Iapp
    yk1 = corner(iTlower, iZlower); yk2 = corner(iTlower, iZupper)                       # :272-273
    yk3 = corner(iTupper, iZlower); yk4 = corner(iTupper, iZupper)                       # :274-275
    bigX = cs.bigA * bigX + Iapp                                                         # :278
    alphaZ = 0.0; alphaT = 0.0                                                           # :281-285
    if Zupper != Zlower: alphaZ = (obs.cellSOC - Zlower) / (Zupper - Zlower)
    if Tupper != Tlower: alphaT = (T - Tlower) / (Tupper - Tlower)
    yk = (1 - alphaT) * ((1 - alphaZ) * yk1 + alphaZ * yk2) + alphaT * ((1 - alphaZ) * yk3 + alphaZ * yk4)
    obs.negIfdl0 = yk[ind["negIfdl0"][0]]; obs.posIfdl3 = yk[ind["posIfdl3"][0]]         # :289-297
    obs.negIf0 = yk[ind["negIf0"][0]]; obs.posIf3 = yk[ind["posIf3"][0]]
    obs.negThetass = yk[ind["negThetass"]] + SOC0n                                       # :302-310
    obs.posThetass = yk[ind["posThetass"]] + SOC0p
    obs.negThetass0 = yk[ind["negThetass0"][0]] + SOC0n
    obs.posThetass3 = yk[ind["posThetass3"][0]] + SOC0p
    clamp = lambda v: np.minimum(np.maximum(v, 1e-6), 1 - 1e-6)
    obs.negThetass = clamp(obs.negThetass); obs.posThetass = clamp(obs.posThetass)
    obs.negThetass0 = float(clamp(obs.negThetass0)); obs.posThetass3 = float(clamp(obs.posThetass3))
    UocpnAvg = fn.neg.Uocp(obs.negSOC, T); UocppAvg = fn.pos.Uocp(obs.posSOC, T)         # :313-317
    obs.negPhise = yk[ind["negPhise"]] + UocpnAvg
    obs.posPhise = yk[ind["posPhise"]] + UocppAvg
    obs.negPhise0 = yk[ind["negPhise0"][0]] + UocpnAvg
    obs.Phie = np.zeros(len(ind["Phie"]) + 1)                                            # :320-322
    obs.Phie[0] = -obs.negPhise0
    obs.Phie[1:] = yk[ind["Phie"]] - obs.negPhise0
    obs.Thetae = np.maximum(yk[ind["Thetae"]] + 1, 1e-6)                                 # :325-326
    k0n = fn.neg.k0(obs.negSOC, T); k0p = fn.pos.k0(obs.posSOC, T)                       # :329-334
    i0n = k0n * np.sqrt(obs.Thetae[0] * (1 - obs.negThetass0) * obs.negThetass0)
    i0p = k0p * np.sqrt(obs.Thetae[-1] * (1 - obs.posThetass3) * obs.posThetass3)
    negEta0 = 2 * R * T / F * np.arcsinh(obs.negIf0 / (2 * i0n))
    posEta3 = 2 * R * T / F * np.arcsinh(obs.posIf3 / (2 * i0p))
    Uocpn0 = fn.neg.Uocp(obs.negThetass0, T); Uocpp3 = fn.pos.Uocp(obs.posThetass3, T)   # :337-344
    Rfn = fn.neg.Rf(obs.negSOC, T); Rfp = fn.pos.Rf(obs.posSOC, T)
    Vcell = posEta3 - negEta0 + yk[ind["Phie"][-1]] + Uocpp3 - Uocpn0 + (Rfp * obs.posIfdl3 - Rfn * obs.negIfdl0)
    Vcell = Vcell - Rc * Iapp
    obs.negPhis = yk[ind["negPhis"]]; obs.posPhis = yk[ind["posPhis"]] + Vcell           # :347-348
    obs.Vcell = Vcell                                                                    # :356
    return float(Vcell), NS(bigX=bigX, SOCnAvg=SOCnAvg, SOCpAvg=SOCpAvg), obs            # :351-353


# ------------------------------------------------------------------------------------------
# initKF.m
# ------------------------------------------------------------------------------------------
def initKF(SOC0, T0, SigmaX0, SigmaV, SigmaW, blend, ROMs):
    """kfData = initKF(SOC0,T0,SigmaX0,SigmaV,SigmaW,blend,ROMs)   (initKF.m:30-136)"""
    # not restated :32 (clear iterSPKF: another filter's persistent variables)
    # not restated :42 (kfData.T = T0 + 273.15 is overwritten by the set-point grid at :124 before anything reads it)
    # not restated :108-118 (sigma-point weights of the SPKF variant; the EKF never reads alpham / alphac / h)
    kf = NS()
    kf.SOC = SOC0 / 100                                                                  # :37
    if T0 > 100: T0 = T0 - 273.15                                                        # :38-41
    b = str(blend).upper()
    if b in ("OUTB", "OB"): kf.method = "OB"                                             # :44-53
    elif b in ("MDLB", "MB"): kf.method = "MB"
    else: kf.method = "OB"
    mdls = ROMs.ROMmdls
    nT, nZ = len(mdls), len(mdls[0])
    T = np.array([[m.T for m in row] for row in mdls]); Z = np.array([[m.SOC for m in row] for row in mdls])   # :57-59
    n = mdls[0][0].A.shape[0] - 1                                                        # :60
    nz = mdls[0][0].C.shape[0]                                                           # :61
    SigmaX0 = np.asarray(SigmaX0, float)
    if SigmaX0.shape != (n + 1, n + 1):                                                  # :66-69
        raise ValueError("Input argument SigmaX0 has wrong dimension. Should be %d by %d" % (n + 1, n + 1))
    SigmaX = SigmaX0[:-1, :-1]                                                           # :70
    kf.M = [[None] * nZ for _ in range(nT)]
    for theT in range(nT):                                                               # :71-72
        for theZ in range(nZ):
            A = mdls[theT][theZ].A                                                       # :73
            if A[-1, -1] != 1: raise ValueError("A does not have integrator state")      # :74-77
            if not np.array_equal(np.diag(np.diag(A)), A): raise ValueError("A is not diagonal")   # :78-81
            B = mdls[theT][theZ].B                                                       # :84
            if np.prod(B) != 1: raise ValueError("B is not all units values")            # :85-88
            kf.M[theT][theZ] = NS(A=np.diag(A[:n, :n]).copy(),                            # :82
                                  C=mdls[theT][theZ].C[:, :-1].copy(),                    # :91
                                  D=mdls[theT][theZ].D.copy(),                            # :92
                                  xhat=np.zeros(n), SigmaX=SigmaX.copy())                 # :94-95
    kf.x0 = 0.0; kf.SigmaX0 = float(SigmaX0[-1, -1])                                     # :98-99
    kf.xhat = np.zeros(n + 1); kf.SigmaX = SigmaX0.copy()                                # :100-101 (MB only)
    kf.SigmaV = SigmaV; kf.SigmaW = SigmaW                                               # :104-105
    kf.priorI = 0.0                                                                      # :121
    kf.T = T; kf.Z = Z; kf.n = n; kf.nz = nz                                             # :124-127
    kf.tfData = ROMs.tfData; kf.Ts = ROMs.xraData.Tsamp; kf.SOC0 = SOC0 / 100            # :130-133
    kf.cellData = ROMs.cellData                                                          # :135
    kf._persist = None                    # per-instance stand-in for iterEKF's persistent variables, which :31 clears   # :31
    return kf


# ------------------------------------------------------------------------------------------
# iterEKF.m (both blend methods: 'OB' output blend -- what runMPC.m:16 selects -- and 'MB' model blend)
# ------------------------------------------------------------------------------------------
class EKFLost(Exception):
    """warnCount > 10 (iterEKF.m:55-59): the reference returns NaNs and EKFmatsHandler then fails."""


def _ekf_setupIndsLocs(ekfData):
    """iterEKF.m:610-735"""
    # not restated :689-690 (ind.Thetae0 / ind.Thetae3: nothing reads them)
    names = ekfData.tfData.names; locs = np.asarray(ekfData.tfData.xLoc, float)          # :612-613
    ind = {}; loc = {}
    for nm in ("negIfdl", "negIf", "negPhis", "negPhise", "negThetass", "posIfdl", "posIf", "posPhis",
               "posPhise", "posThetass", "negPhie", "sepPhie", "posPhie", "negThetae", "sepThetae", "posThetae"):
        ind[nm] = _find(names, nm); loc[nm] = [locs[i] for i in ind[nm]]                 # :616-625,628-637,640-645,648-653
    for key, parts in (("Ifdl", ("negIfdl", "posIfdl")), ("If", ("negIf", "posIf")), ("Phis", ("negPhis", "posPhis")),
                       ("Phise", ("negPhise", "posPhise")), ("Thetass", ("negThetass", "posThetass")),
                       ("Phie", ("negPhie", "sepPhie", "posPhie")), ("Thetae", ("negThetae", "sepThetae", "posThetae"))):
        ind[key] = sum((ind[p] for p in parts), []); loc[key] = [locs[i] for i in ind[key]]   # :656-670

    def at(key, x):
        return [i for i in ind[key] if locs[i] == x]
    ind["Ifdl0"], ind["Ifdl3"] = at("Ifdl", 0), at("Ifdl", 3)                            # :682-688
    ind["If0"], ind["If3"] = at("If", 0), at("If", 3)
    ind["Thetass0"], ind["Thetass3"] = at("Thetass", 0), at("Thetass", 3)
    ind["Phise0"] = at("Phise", 0)
    if not ind["Ifdl0"]: raise ValueError("Simulation requires ifdl at negative-collector!")   # :692-734
    if not ind["Ifdl3"]: raise ValueError("Simulation requires ifdl at positive-collector!")
    if not ind["If0"]: raise ValueError("Simulation requires if at negative-collector!")
    if not ind["If3"]: raise ValueError("Simulation requires if at positive-collector!")
    if loc["Thetae"][0] > 0: raise ValueError("Simulation requires thetae at negative-collector!")
    if loc["Thetae"][-1] > 3 + EPS or loc["Thetae"][-1] < 3 - EPS:
        raise ValueError("Simulation requires thetae at positive-collector!")
    if not ind["Thetass0"]: raise ValueError("Simulation requires thetass at negative-collector!")
    if not ind["Thetass3"]: raise ValueError("Simulation requires thetass at positive-collector!")
    if not ind["Phise0"]: raise ValueError("Simulation requires phise at negative-collector!")
    warned = 0
    if loc["Phie"][0] == 0:
        warned = 1                                     # shortWarn -> warnCount+1 (:729, :749)
        ind["Phie"] = ind["Phie"][1:]; loc["Phie"] = loc["Phie"][1:]
    if loc["Phie"][-1] > 3 + EPS or loc["Phie"][-1] < 3 - EPS:
        raise ValueError("Simulation requires phie at positive-collector!")
    return ind, loc, warned


def iterEKF(vk, ik, Tk, ekfData):
    """[zk,boundzk,ekfData,Xind] = iterEKF(vk,ik,Tk,ekfData)   (iterEKF.m:30-210, ekfData.method 'OB' or 'MB')"""
    # not restated :47-52 (SOC0 guessed from the voltage when ekfData.SOC0 is empty: initKF.m:133 always sets it)
    P = ekfData._persist                                                                 # :31
    if P is None:                                                                        # :32-45
        ind, loc, warned = _ekf_setupIndsLocs(ekfData)
        P = NS(ind=ind, loc=loc, Tpts=np.unique(ekfData.T), Zpts=np.unique(ekfData.Z), n=ekfData.n,
               warnCount=0, q18=0, cellData=ekfData.cellData, F=ekfData.cellData.const.F, R=ekfData.cellData.const.R)
        # NB the reference zeroes warnCount (:42) *after* setupIndsLocs (:33), so a set-up warning is dropped
        P.numT, P.numZ = len(P.Tpts), len(P.Zpts)
        ekfData.ind = ind; ekfData.loc = loc
        ekfData.Q = ekfData.cellData.function.const.Q()                                  # :41
        ekfData._persist = P
    ind, loc, n, cellData, F, R = P.ind, P.loc, P.n, P.cellData, P.F, P.R
    fn = cellData.function
    Tpts, Zpts = P.Tpts, P.Zpts

    # not restated :742-748 (the 'on'/'off' switch of shortWarn, never used, and the console print)
    def shortWarn(msg):                                                                  # :741-751
        P.warnCount += 1                                                                 # :749

    if P.warnCount > 10:                                                                 # :55-59
        zk = np.full(ekfData.nz + 2, np.nan)
        Xind = NS(gamma=np.full(4, np.nan), theT=np.full(4, np.nan), theZ=np.full(4, np.nan))
        return zk, zk.copy(), ekfData, Xind
    if Tk > 100: pass                                                                    # :62-66
    else: Tk = Tk + 273.15

    M = ekfData.M
    MB = ekfData.method == "MB"
    res_per_amp = ekfData.Ts / (3600 * ekfData.Q)

    # ---- helpers (nested functions of iterEKF.m) ----
    def getXind(Tk, SOC):                                                                # :219-255
        dZ = np.abs(SOC - Zpts); iZ = _msort(dZ)                                         # :221
        Zupper = Zlower = Zpts[0]; iZupper = iZlower = iZ[0]                             # :222-223
        if len(iZ) > 1:                                                                  # :224-231
            Zupper = Zpts[iZ[0]]; iZupper = iZ[0]
            Zlower = Zpts[iZ[1]]; iZlower = iZ[1]
            if Zupper < Zlower:
                Zupper = Zpts[iZ[1]]; iZupper = iZ[1]
                Zlower = Zpts[iZ[0]]; iZlower = iZ[0]
        dT = np.abs(Tk - Tpts); iT = _msort(dT)                                          # :234
        Tupper = Tlower = Tpts[0]; iTupper = iTlower = iT[0]                             # :235-236
        if len(iT) > 1:                                                                  # :237-244
            Tupper = Tpts[iT[0]]; iTupper = iT[0]
            Tlower = Tpts[iT[1]]; iTlower = iT[1]
            if Tupper < Tlower:
                Tupper = Tpts[iT[1]]; iTupper = iT[1]
                Tlower = Tpts[iT[0]]; iTlower = iT[0]
        alphaZ = 0.0; alphaT = 0.0                                                       # :246
        if len(iZ) > 1: alphaZ = (SOC - Zlower) / (Zupper - Zlower)                      # :247 (Q14)
        if len(iT) > 1: alphaT = (Tk - Tlower) / (Tupper - Tlower)                       # :248
        # :249-254
        gamma = np.array([(1 - alphaT) * (1 - alphaZ), (1 - alphaT) * alphaZ, alphaT * (1 - alphaZ), alphaT * alphaZ])
        return NS(gamma=gamma, theT=np.array([iTlower, iTlower, iTupper, iTupper]),
                  theZ=np.array([iZlower, iZupper, iZlower, iZupper]))

    def getVariables(ik, Xind, T):                                                       # :259-417
        if MB:                                                                           # :265-275
            xhat = ekfData.xhat
            xk = [xhat[:n]] * 4                                                          # :310-312: every model reads xhat(1:n)
            x0 = xhat[-1]
        else:
            xk = [M[Xind.theT[m]][Xind.theZ[m]].xhat for m in range(4)]
            x0 = ekfData.x0
        xSOC = ekfData.SOC0 - x0 * res_per_amp                                           # :277
        SOCnAvg = fn.neg.soc(xSOC, T); SOCpAvg = fn.pos.soc(xSOC, T)                     # :282-295
        if SOCnAvg < 0: shortWarn("SOCnAvg < 0"); SOCnAvg = 1e-6
        if SOCnAvg > 1: shortWarn("SOCnAvg > 1"); SOCnAvg = 1 - 1e-6
        if SOCpAvg < 0: shortWarn("SOCpAvg < 0"); SOCpAvg = 1e-6
        if SOCpAvg > 0.998: shortWarn("SOCpAvg > 1"); SOCpAvg = 0.998
        cols = []                                                                        # :301 (Z preallocated)
        for m in range(4):                                                               # :302-317 (both methods: xk holds xhat(1:n) four times for 'MB', :314-316)
            mdl = M[Xind.theT[m]][Xind.theZ[m]]
            cols.append(mdl.C @ xk[m] + mdl.D[:, 0] * ik)
        Z = np.column_stack(cols) @ Xind.gamma
        If0 = Z[ind["If0"][0]]; If3 = Z[ind["If3"][0]]                                   # :324-325
        ii = ind["negThetass"]
        Z[ii] = Z[ii] + SOCnAvg                                                          # :331-343
        if np.any(Z[ii] < 0):
            shortWarn("negThetass < 0"); v = Z[ii]; v[v < 0] = 1e-6; Z[ii] = v
        if np.any(Z[ii] > 1):
            shortWarn("negThetass > 1"); v = Z[ii]; v[v > 1] = 1 - 1e-6; Z[ii] = v
        ii = ind["posThetass"]
        Z[ii] = Z[ii] + SOCpAvg                                                          # :345-357
        if np.any(Z[ii] < 0):
            shortWarn("posThetass < 0"); v = Z[ii]; v[v < 0] = 1e-6; Z[ii] = v
        if np.any(Z[ii] > 0.998):
            shortWarn("posThetass > 1"); v = Z[ii]; v[v > 0.998] = 0.998; Z[ii] = v
        UocpnAvg = fn.neg.Uocp(SOCnAvg, T); UocppAvg = fn.pos.Uocp(SOCpAvg, T)           # :362-367
        Z[ind["negPhise"]] = Z[ind["negPhise"]] + UocpnAvg
        if ind["posPhise"]:
            Z[ind["posPhise"]] = Z[ind["posPhise"]] + UocppAvg
        PhieTilde3 = Z[ind["Phie"][-1]]                                                  # :371-379
        Phise0 = Z[ind["Phise0"][0]]
        for k, i in enumerate(ind["Phie"]):
            Z[i] = (0 - Phise0) if loc["Phie"][k] == 0 else (Z[i] - Phise0)
        Z[ind["Thetae"]] = Z[ind["Thetae"]] + 1                                          # :383-389
        if np.any(Z[ind["Thetae"]] < 0):
            # Q18: the reference branch has typos and would throw; intended behaviour restated
            shortWarn("Thetae < 0"); P.q18 += 1
            v = Z[ind["Thetae"]]; v[v < 0] = 1e-6; Z[ind["Thetae"]] = v
        k0n = fn.neg.k0(SOCnAvg, T); k0p = fn.pos.k0(SOCpAvg, T)                         # :392-401
        th0 = Z[ind["Thetass0"][0]]; th3 = Z[ind["Thetass3"][0]]
        i0n = k0n * np.sqrt(Z[ind["Thetae"][0]] * (1 - th0) * th0)
        i0p = k0p * np.sqrt(Z[ind["Thetae"][-1]] * (1 - th3) * th3)
        negEta0 = 2 * R * T / F * np.arcsinh(If0 / (2 * i0n))
        posEta3 = 2 * R * T / F * np.arcsinh(If3 / (2 * i0p))
        Uocpn0 = fn.neg.Uocp(th0, T); Uocpp3 = fn.pos.Uocp(th3, T)                       # :404-410
        Rfn = fn.neg.Rf(SOCnAvg, T); Rfp = fn.pos.Rf(SOCpAvg, T)
        Vcell = posEta3 - negEta0 + PhieTilde3 + Uocpp3 - Uocpn0 + (Rfp * Z[ind["Ifdl3"][0]] - Rfn * Z[ind["Ifdl0"][0]])
        Z[ind["posPhis"]] = Z[ind["posPhis"]] + Vcell                                    # :413
        Zsoc = ekfData.SOC0 - x0 * res_per_amp                                           # :416
        return float(Vcell), Z, float(Zsoc)

    def getChatV(Xind, Z, Tk):                                                           # :421-519
        Cs = [Xind.gamma[m] * M[Xind.theT[m]][Xind.theZ[m]].C for m in range(4)]         # :423-426
        x0 = ekfData.xhat[-1] if MB else ekfData.x0                                      # :435-437 / :449
        xSOC = ekfData.SOC0 - x0 * res_per_amp                                           # :438 / :450
        SOCnAvg = fn.neg.soc(xSOC, Tk); SOCpAvg = fn.pos.soc(xSOC, Tk)                   # :439-440 / :451-452
        Rfn = fn.neg.Rf(SOCnAvg, Tk); Rfp = fn.pos.Rf(SOCpAvg, Tk)                       # :441-442 / :453-454
        if MB:                                                                           # :448-460,475-479,489-491,512-518; the switches at :469-470,483-484,502-503
            rsum = lambda key: np.sum(np.vstack([C[ind[key][0 if key != "Phie" else -1], :] for C in Cs]), axis=0)
            Chat = Rfp * rsum("Ifdl3") - Rfn * rsum("Ifdl0")
            k0n = fn.neg.k0(SOCnAvg, Tk); k0p = fn.pos.k0(SOCpAvg, Tk)
            th0 = Z[ind["Thetass0"][0]]; th3 = Z[ind["Thetass3"][0]]
            i0n = k0n * np.sqrt(Z[ind["Thetae"][0]] * (1 - th0) * th0)
            i0p = k0p * np.sqrt(Z[ind["Thetae"][-1]] * (1 - th3) * th3)
            Rctn = R * Tk / (F * i0n); Rctp = R * Tk / (F * i0p)
            Chat = Chat + Rctp * rsum("If3") - Rctn * rsum("If0")
            Chat = Chat + rsum("Phie")
            dUocpn0 = fn.neg.dUocp(th0, Tk); dUocpp3 = fn.pos.dUocp(th3, Tk)
            res0n = -dUocpn0 * ekfData.Ts * (fn.neg.soc(1, Tk) - fn.neg.soc(0, Tk)) / (3600 * ekfData.Q)
            res0p = -dUocpp3 * ekfData.Ts * (fn.pos.soc(1, Tk) - fn.pos.soc(0, Tk)) / (3600 * ekfData.Q)
            Chat0 = res0p - res0n
            Chat = Chat + dUocpp3 * rsum("Thetass3") - dUocpn0 * rsum("Thetass0")
            return np.concatenate([Chat, [Chat0]]), Chat0                                # :517 augmented with the integrator
        Chat = [Rfp * C[ind["Ifdl3"][0], :] - Rfn * C[ind["Ifdl0"][0], :] for C in Cs]   # :444-447
        k0n = fn.neg.k0(SOCnAvg, Tk); k0p = fn.pos.k0(SOCpAvg, Tk)                       # :463-468
        th0 = Z[ind["Thetass0"][0]]; th3 = Z[ind["Thetass3"][0]]
        i0n = k0n * np.sqrt(Z[ind["Thetae"][0]] * (1 - th0) * th0)
        i0p = k0p * np.sqrt(Z[ind["Thetae"][-1]] * (1 - th3) * th3)
        Rctn = R * Tk / (F * i0n); Rctp = R * Tk / (F * i0p)
        Chat = [Chat[m] + Rctp * Cs[m][ind["If3"][0], :] - Rctn * Cs[m][ind["If0"][0], :] for m in range(4)]   # :471-474
        Chat = [Chat[m] + Cs[m][ind["Phie"][-1], :] for m in range(4)]                   # :485-488
        dUocpn0 = fn.neg.dUocp(th0, Tk); dUocpp3 = fn.pos.dUocp(th3, Tk)                 # :495-501
        res0n = -dUocpn0 * ekfData.Ts * (fn.neg.soc(1, Tk) - fn.neg.soc(0, Tk)) / (3600 * ekfData.Q)
        res0p = -dUocpp3 * ekfData.Ts * (fn.pos.soc(1, Tk) - fn.pos.soc(0, Tk)) / (3600 * ekfData.Q)
        Chat0 = res0p - res0n
        Chat = [Chat[m] + (dUocpp3 * Cs[m][ind["Thetass3"][0], :] - dUocpn0 * Cs[m][ind["Thetass0"][0], :])
                for m in range(4)]                                                       # :504-511
        return Chat, Chat0

    def getChatZ(Xind, Z, Tk):                                                           # :523-602
        Cs = [Xind.gamma[m] * M[Xind.theT[m]][Xind.theZ[m]].C for m in range(4)]         # :525-528
        if MB:                                                                           # :534-535,537-538,544-545,554-558,570-575,587-588,596-600 (the method switches)
            Chat = Cs[0] + Cs[1] + Cs[2] + Cs[3]
            Chat0 = np.zeros(Cs[0].shape[0])
            ChatV, _ = getChatV(Xind, Z, Tk)
            for r in ind["posPhis"]:
                Chat[r, :] = Chat[r, :] + ChatV[:-1]
            Chat0[ind["posPhis"]] = ChatV[-1]
            res0n = -ekfData.Ts * (fn.neg.soc(1, Tk) - fn.neg.soc(0, Tk)) / (3600 * ekfData.Q)
            res0p = -ekfData.Ts * (fn.pos.soc(1, Tk) - fn.pos.soc(0, Tk)) / (3600 * ekfData.Q)
            Chat0[ind["negThetass"]] = res0n; Chat0[ind["posThetass"]] = res0p
            x0 = ekfData.xhat[-1]
            xSOC = ekfData.SOC0 - x0 * res_per_amp
            SOCnAvg = fn.neg.soc(xSOC, Tk); SOCpAvg = fn.pos.soc(xSOC, Tk)
            dUocpn = fn.neg.dUocp(SOCnAvg, Tk); dUocpp = fn.pos.dUocp(SOCpAvg, Tk)
            Chat0[ind["negPhise"]] = dUocpn * res0n
            Chat0[ind["posPhise"]] = dUocpp * res0p
            Chat0[ind["Phie"]] = -dUocpn * res0n
            for r in ind["Phie"]:
                Chat[r, :] = Chat[r, :] - Chat[ind["Phise0"][0], :]
            return np.column_stack([Chat, Chat0]), ChatV, None, None
        Chat = [C.copy() for C in Cs]                                                    # :536
        Chat0 = np.zeros(Cs[0].shape[0])                                                 # :540
        ChatV, ChatV0 = getChatV(Xind, Z, Tk)                                            # :543
        for kk in range(4):                                                              # :546-553
            for r in ind["posPhis"]:
                Chat[kk][r, :] = Chat[kk][r, :] + ChatV[kk]
        Chat0[ind["posPhis"]] = ChatV0
        res0n = -ekfData.Ts * (fn.neg.soc(1, Tk) - fn.neg.soc(0, Tk)) / (3600 * ekfData.Q)   # :562-567
        res0p = -ekfData.Ts * (fn.pos.soc(1, Tk) - fn.pos.soc(0, Tk)) / (3600 * ekfData.Q)
        Chat0[ind["negThetass"]] = res0n; Chat0[ind["posThetass"]] = res0p
        x0 = ekfData.x0
        xSOC = ekfData.SOC0 - x0 * res_per_amp                                           # :576-583
        SOCnAvg = fn.neg.soc(xSOC, Tk); SOCpAvg = fn.pos.soc(xSOC, Tk)
        dUocpn = fn.neg.dUocp(SOCnAvg, Tk); dUocpp = fn.pos.dUocp(SOCpAvg, Tk)
        Chat0[ind["negPhise"]] = dUocpn * res0n
        Chat0[ind["posPhise"]] = dUocpp * res0p
        Chat0[ind["Phie"]] = -dUocpn * res0n                                             # :586
        for kk in range(4):                                                              # :589-595
            C = Chat[kk]
            for r in ind["Phie"]:
                C[r, :] = C[r, :] - C[ind["Phise0"][0], :]
        return Chat, ChatV, Chat0, ChatV0

    if MB:
        # ---- Steps 1a/1b, model blend: one blended model (:90-102) ----
        SOC = ekfData.SOC0 - ekfData.xhat[-1] * res_per_amp                              # :92
        Xind = getXind(Tk, SOC)                                                          # :93
        As = np.column_stack([M[Xind.theT[m]][Xind.theZ[m]].A for m in range(4)])       # :94-97
        AMB = np.concatenate([As @ Xind.gamma, [1.0]])                                   # :98
        ekfData.xhat = AMB * ekfData.xhat + ekfData.priorI                               # :99
        ekfData.SigmaX = np.diag(AMB) @ ekfData.SigmaX @ np.diag(AMB) + ekfData.SigmaW   # :101
        SOC = ekfData.SOC0 - ekfData.xhat[-1] * res_per_amp                              # :102
        Xind = getXind(Tk, SOC)                                                          # :107-108
        vhat, Z, _ = getVariables(ik, Xind, Tk)
        ChatV, _ = getChatV(Xind, Z, Tk)                                                 # :125-128
        SigmaVtilde = ChatV @ ekfData.SigmaX @ ChatV + ekfData.SigmaV
        L = ekfData.SigmaX @ ChatV / SigmaVtilde
        residual = vk - vhat                                                             # :133
        ekfData.xhat = ekfData.xhat + L * residual                                       # :160-161
        SigmaX = ekfData.SigmaX - np.outer(L * SigmaVtilde, L)                           # :162-163
        _, SS, VVt = np.linalg.svd(SigmaX)                                               # :164-166
        HH = VVt.T @ np.diag(SS) @ VVt
        SigmaX = (SigmaX + SigmaX.T + HH + HH.T) / 4
        nbump = 0
        if residual ** 2 > 9 * SigmaVtilde:                                              # :169-172
            SigmaX = SigmaX * 2; nbump = 1
        ekfData.SigmaX = SigmaX                                                          # :173
        SOC = ekfData.SOC0 - ekfData.xhat[-1] * res_per_amp                              # :176
        Xind = getXind(Tk, SOC)                                                          # :181-183
        vhat, zk, Zsoc = getVariables(ik, Xind, Tk)
        zk = np.concatenate([zk, [vhat, Zsoc]])
        res = -ekfData.Ts / (3600 * ekfData.Q)
        ChatZ, ChatVb, _, _ = getChatZ(Xind, zk, Tk)                                     # :199-203
        SigmaZ = ChatZ @ ekfData.SigmaX @ ChatZ.T
        SigmaVb = ChatVb @ ekfData.SigmaX @ ChatVb
        SigmaSOC = res * ekfData.SigmaX[-1, -1] * res
        with np.errstate(invalid="ignore"):
            boundzk = 3 * np.sqrt(np.concatenate([np.diag(SigmaZ), [SigmaVb, SigmaSOC]]))    # :205
        ekfData.priorI = ik                                                              # :210
        ekfData.nbump_last = nbump
        return zk, boundzk, ekfData, Xind
    # ---- Steps 1a/1b: time update of ALL models (:70-89; the switch at :70-71 / :90 is the `if MB` above) ----
    for theT in range(P.numT):
        for theZ in range(P.numZ):
            mdl = M[theT][theZ]
            A = mdl.A
            mdl.xhat = A * mdl.xhat + ekfData.priorI                                     # :77
            mdl.SigmaX = np.diag(A) @ mdl.SigmaX @ np.diag(A) + ekfData.SigmaW           # :81 (Q12)
    ekfData.x0 = ekfData.x0 + ekfData.priorI                                             # :85
    ekfData.SigmaX0 = ekfData.SigmaX0 + ekfData.SigmaW                                   # :86
    SOC = ekfData.SOC0 - ekfData.x0 * res_per_amp                                        # :89
    # ---- Step 1c (:107-108) ----
    Xind = getXind(Tk, SOC)
    vhat, Z, _ = getVariables(ik, Xind, Tk)
    # ---- Step 2a (:112-124) ----
    ChatV, C0 = getChatV(Xind, Z, Tk)
    SigmaVtilde = [None] * 4; L = [None] * 4
    for m in range(4):
        SigmaX = M[Xind.theT[0]][Xind.theZ[0]].SigmaX                                    # :117 (Q1)
        SigmaVtilde[m] = ChatV[m] @ SigmaX @ ChatV[m] + ekfData.SigmaV
        L[m] = SigmaX @ ChatV[m] / SigmaVtilde[m]
    SigmaX0 = ekfData.SigmaX0
    SigmaVtilde0 = C0 * SigmaX0 * C0 + ekfData.SigmaV
    L0 = SigmaX0 * C0 / SigmaVtilde0
    # ---- Steps 2b/2c (:133-159) ----
    residual = vk - vhat
    nbump = 0
    for m in range(4):
        mdl = M[Xind.theT[m]][Xind.theZ[m]]
        mdl.xhat = mdl.xhat + L[m] * residual                                            # :138
        SigmaX = mdl.SigmaX - np.outer(L[m] * SigmaVtilde[m], L[m])                      # :142
        _, SS, VVt = np.linalg.svd(SigmaX)                                               # :143
        HH = VVt.T @ np.diag(SS) @ VVt                                                   # :144
        SigmaX = (SigmaX + SigmaX.T + HH + HH.T) / 4                                     # :145 (Q13)
        if residual ** 2 > 9 * SigmaVtilde[m]:                                           # :148-151
            SigmaX = SigmaX * 2; nbump += 1
        mdl.SigmaX = SigmaX
    ekfData.x0 = ekfData.x0 + L0 * residual                                              # :155
    ekfData.SigmaX0 = ekfData.SigmaX0 - L0 * SigmaVtilde0 * L0                           # :156
    SOC = ekfData.SOC0 - ekfData.x0 * res_per_amp                                        # :159
    # ---- Steps 3a/3b (:181-205) ----
    Xind = getXind(Tk, SOC)
    vhat, zk, Zsoc = getVariables(ik, Xind, Tk)
    zk = np.concatenate([zk, [vhat, Zsoc]])                                              # :183
    res = -ekfData.Ts / (3600 * ekfData.Q)
    SigmaZ = np.zeros((ekfData.nz, ekfData.nz)); SigmaVb = 0.0
    ChatZ, ChatVb, ChatZ0, ChatV0 = getChatZ(Xind, zk, Tk)
    for m in range(4):
        SigmaX = M[Xind.theT[0]][Xind.theZ[0]].SigmaX                                    # :192 (Q4)
        SigmaZ = SigmaZ + ChatZ[m] @ SigmaX @ ChatZ[m].T
        SigmaVb = SigmaVb + ChatVb[m] @ SigmaX @ ChatVb[m]
    SigmaZ = SigmaZ + np.outer(ChatZ0 * ekfData.SigmaX0, ChatZ0)
    SigmaVb = SigmaVb + ChatV0 * ekfData.SigmaX0 * ChatV0
    SigmaSOC = res * ekfData.SigmaX0 * res
    with np.errstate(invalid="ignore"):
        boundzk = 3 * np.sqrt(np.concatenate([np.diag(SigmaZ), [SigmaVb, SigmaSOC]]))    # :205
    ekfData.priorI = ik                                                                  # :210
    ekfData.nbump_last = nbump
    return zk, boundzk, ekfData, Xind


# ------------------------------------------------------------------------------------------
# EKFmatsHandler.m
# ------------------------------------------------------------------------------------------
def EKFmatsHandler(ekfData, Xind, zk, Tk):
    """[MPC, xhat] = EKFmatsHandler(ekfData, Xind, zk, Tk)   (EKFmatsHandler.m:1-115)"""
    g = np.asarray(Xind.gamma, float).ravel()
    if np.all(np.isnan(g)):
        raise EKFLost("Xind is NaN (iterEKF.m:55-59); MATLAB would fail indexing M(NaN,NaN)")
    imax = int(np.nanargmax(g))                      # first maximum, NaN ignored   # :26
    iT = int(Xind.theT[imax]); iZ = int(Xind.theZ[imax])                                 # :27-28
    mdl = ekfData.M[iT][iZ]                                                              # :30
    xhat = np.concatenate([mdl.xhat, [ekfData.xhat[-1]]])                                # :33 (Q3)
    Ad = np.concatenate([mdl.A, [1.0]])                                                  # :35
    nx = len(Ad)                                                                         # :36
    A = np.diag(Ad); B = np.ones((nx, 1)); C = mdl.C; D = mdl.D                          # :37-40
    r = -ekfData.Ts / (3600 * ekfData.Q)                                                 # :43-45
    Csoc = np.concatenate([np.zeros(nx - 1), [r]]); Dsoc = 0.0
    ind = ekfData.ind; cellData = ekfData.cellData; fn = cellData.function               # :49-50
    F = cellData.const.F; R = cellData.const.R                                           # :51-52
    TK = Tk + 273.15                                                                     # :53
    SOCavg = zk[-1]                                                                      # :56-58
    SOCnAvg = fn.neg.soc(SOCavg, TK); SOCpAvg = fn.pos.soc(SOCavg, TK)
    k0n = fn.neg.k0(SOCnAvg, TK); k0p = fn.pos.k0(SOCpAvg, TK)                           # :60-66
    th0 = zk[ind["Thetass0"][0]]; th3 = zk[ind["Thetass3"][0]]
    i0n = k0n * np.sqrt(zk[ind["Thetae"][0]] * (1 - th0) * th0)
    i0p = k0p * np.sqrt(zk[ind["Thetae"][-1]] * (1 - th3) * th3)
    Rfn = fn.neg.Rf(SOCnAvg, TK); Rfp = fn.pos.Rf(SOCpAvg, TK)                           # :68-69
    Cv = Rfp * C[ind["Ifdl3"][0], :] - Rfn * C[ind["Ifdl0"][0], :] + C[ind["Phie"][-1], :]      # :75-76
    Dv = Rfp * D[ind["Ifdl3"][0], 0] - Rfn * D[ind["Ifdl0"][0], 0] + D[ind["Phie"][-1], 0]      # :79-80
    Upos = fn.pos.Uocp(th3, TK); Uneg = fn.neg.Uocp(th0, TK)                             # :84-85
    If0 = zk[ind["If0"][0]]; If3 = zk[ind["If3"][0]]                                     # :87
    negEta0 = 2 * R * TK / F * np.arcsinh(If0 / (2 * i0n))                               # :88-89
    posEta3 = 2 * R * TK / F * np.arcsinh(If3 / (2 * i0p))
    b_phi = 0.01 * 0                                                                     # :91
    bv = (Upos - Uneg) + (posEta3 - negEta0) + b_phi                                     # :92
    Uocpn = fn.neg.Uocp(SOCnAvg)                                                         # :96 (Q21)
    Cphi = C[ind["negPhise"][1], :]; Dphi = D[ind["negPhise"][1], 0]                     # :97-98
    MPC = NS(A=A, B=B, Csoc=Csoc, Dsoc=Dsoc, Cv=np.concatenate([Cv, [0.0]]), Dv=Dv,       # :101-114
             Cphi=np.concatenate([Cphi, [0.0]]), Dphi=Dphi, bv=bv, bphi=Uocpn,
             iT=iT + 1, iZ=iZ + 1, pickWeight=g[imax])
    return MPC, xhat


# ------------------------------------------------------------------------------------------
# predMat.m / constraintsMPC.m / hildreth.m / initMPC.m / iterMPC.m
# ------------------------------------------------------------------------------------------
def predMat(A, B, C, D, Np, Nc):
    """[Phi, G, aug] = predMat(A, B, C, D, Np, Nc)   (predMat.m:1-54)"""
    A = np.atleast_2d(np.asarray(A, float)); B = np.asarray(B, float).reshape(A.shape[0], -1)
    C = np.atleast_2d(np.asarray(C, float)); D = np.asarray(D, float).reshape(C.shape[0], -1)
    nx = A.shape[0]; m = B.shape[1]; q = C.shape[0]                                      # :12-14
    Np = max(1, int(round(Np))); Nc = max(1, min(int(round(Nc)), Np))                    # :15-16
    Abar = np.block([[A, B], [np.zeros((m, nx)), np.eye(m)]])                            # :20-22
    Bbar = np.vstack([np.zeros((nx, m)), np.eye(m)])
    Cbar = np.hstack([C, D])
    nx_aug = nx + m                                                                      # :24
    Phi = np.zeros((q * Np, nx_aug)); G = np.zeros((q * Np, m * Nc))                     # :27-28
    H = []; X = Bbar                                                                     # :31-36
    for _ in range(Np):
        H.append(Cbar @ X); X = Abar @ X
    Ap = np.eye(nx_aug)                                                                  # :39-50
    for i in range(1, Np + 1):
        Ap = Abar @ Ap
        r = slice((i - 1) * q, i * q)
        Phi[r, :] = Cbar @ Ap
        for j in range(1, min(i, Nc) + 1):
            G[r, (j - 1) * m:j * m] = H[i - j]
    return Phi, G, NS(A=Abar, B=Bbar, C=Cbar, nx_aug=nx_aug, m=m, q=q)                   # :53


def constraintsMPC(x_aug_k, cellState, mpcData):
    """[M, gamma] = constraintsMPC(x_aug_k, cellState, mpcData)   (constraintsMPC.m:1-113)"""
    Nc, Np = mpcData.Nc, mpcData.Np                                                      # :15
    cons = np.asarray(mpcData.const.constraints).ravel()
    useCurrent = len(cons) >= 1 and cons[0] == 1                                         # :18-20
    useVoltage = len(cons) >= 2 and cons[1] == 1
    useEta = len(cons) >= 3 and cons[2] == 1
    Ms = []; gs = []                                                                     # :16
    c = mpcData.const
    mpc = cellState.MPC
    if useCurrent:                                                                       # :23-42
        Cu = np.tril(np.ones((Nc, Nc)))
        uk = mpcData.uk_1
        Ms += [Cu, -Cu]
        gs += [(c.u_max - uk) * np.ones(Nc), -(c.u_min - uk) * np.ones(Nc)]
        Ms += [np.eye(Nc), -np.eye(Nc)]
        gs += [c.du_max * np.ones(Nc), -c.du_min * np.ones(Nc)]
    if useVoltage:                                                                       # :45-63
        Phi_v, G_v, _ = predMat(mpc.A, mpc.B, mpc.Cv, mpc.Dv, Np, Nc)
        nrows = Phi_v.shape[0]
        v_max = np.asarray(c.v_max, float).ravel()
        if v_max.size == 1: v_max = np.repeat(v_max, nrows)
        rhs_v = Phi_v @ x_aug_k + mpc.bv * np.ones(nrows)
        Ms.append(G_v); gs.append(v_max - rhs_v)
    if useEta:                                                                           # :66-85
        Phi_e, G_e, _ = predMat(mpc.A, mpc.B, mpc.Cphi, mpc.Dphi, Np, Nc)
        nrows = Phi_e.shape[0]
        eta_min = np.asarray(c.phise_min, float).ravel()
        if eta_min.size == 1: eta_min = np.repeat(eta_min, nrows)
        assert eta_min.size == nrows, "phise_min must be scalar or length(size(Phi_eta,1))."
        rhs_e = Phi_e @ x_aug_k + mpc.bphi * np.ones(nrows)
        Ms.append(-G_e); gs.append(-eta_min + rhs_e)
    if hasattr(c, "z_max"):                                                              # :89-101
        Phi_s = mpcData.Phi_soc; G_s = mpcData.G_soc
        rhs_s = Phi_s @ x_aug_k + mpcData.SOCk_1 * np.ones(Phi_s.shape[0])
        zmax = c.z_max
        if hasattr(c, "z_tol"): zmax = zmax + c.z_tol
        Ms.append(G_s); gs.append(zmax * np.ones(rhs_s.shape) - rhs_s)
    if getattr(c, "z_target_terminal", False):                                           # :104-112
        Phi_s = mpcData.Phi_soc; G_s = mpcData.G_soc
        rN = Phi_s[-1, :] @ x_aug_k
        Ms.append(-G_s[-1:, :]); gs.append(np.array([-c.z_max + rN]))
    if not Ms:
        return np.zeros((0, Nc)), np.zeros(0)
    return np.vstack(Ms), np.concatenate(gs)


def hildreth(E, F, M, gamma, lambda0=None, maxIter=100):
    """[DU, lambda, nexec] = hildreth(E, F, M, gamma, lambda0, maxIter)   (hildreth.m:1-47)"""
    E = np.atleast_2d(np.asarray(E, float)); F = np.asarray(F, float).ravel()
    M = np.atleast_2d(np.asarray(M, float)); gamma = np.asarray(gamma, float).ravel()
    nConstr = M.shape[0]                                                                 # :17,23-25 (maxIter's default: the signature's)
    if lambda0 is None or np.size(lambda0) == 0:                                         # :18-22
        lam = np.zeros(nConstr)
    else:
        lam = np.array(lambda0, float).ravel().copy()
    with np.errstate(all="ignore"):
        H = M @ np.linalg.solve(E, M.T)                                                  # :28
        K = M @ np.linalg.solve(E, F) + gamma                                            # :29
        k = 0
        for k in range(1, int(maxIter) + 1):                                             # :32-42
            lam_old = lam.copy()
            for i in range(nConstr):
                w = -(K[i] + H[i, :] @ lam - H[i, i] * lam[i]) / H[i, i]                 # :35
                lam[i] = np.fmax(0.0, w)                                                 # :36  max(0,NaN)=0
            if np.max(np.abs(lam - lam_old)) < 1e-6:                                     # :39  NaN-propagating
                break
        nexec = k                                                                        # :43
        DU = -np.linalg.solve(E, F + M.T @ lam)                                          # :46
    return DU, lam, nexec


def initMPC(SOC0, Np, Nc, targetSOC, opts=None):
    """mpcData = initMPC(SOC0, Np, Nc, targetSOC, opts)   (initMPC.m:1-84)"""
    # :76-79 getfield_with_default is dict.get below; :81-84 copyfields is the setattr loop at the end
    opts = opts or {}                                                                    # :29
    Np = max(1, int(round(Np))); Nc = max(1, min(int(round(Nc)), Np))                    # :32-33
    d = NS()
    d.Np, d.Nc = Np, Nc                                                                  # :36-37
    d.Ts = opts.get("Ts", 1)                                                              # :38
    d.ref = targetSOC                                                                    # :39 (Q9)
    d.Sigma = np.tril(np.ones((Nc, Nc)))                                                 # :40
    d.Nsim = opts.get("Nsim", np.nan)                                                    # :41
    d.maxHild = 100                                                                      # :47
    d.lambda_ = None                                                                     # :48  []
    d.DUk_1 = np.zeros(Nc); d.xk_1 = None; d.uk_1 = 0.0; d.SOCk_1 = 0.0; d.SOC0 = SOC0   # :49-53
    cons = np.asarray(opts.get("constraints", [1, 1, 1])).ravel() != 0                   # :56-59
    d.const = NS(constraints=cons.astype(int))
    L = dict(opts.get("limits", {}))                                                     # :62-63
    Imax = -opts["ROM"].cellData.function.const.Q(0.5) * opts["Crate"]                   # :66-67
    L["u_min"] = Imax
    for k, v in L.items():                                                               # :72
        setattr(d.const, k, v)
    return d


def iterMPC(xk, cellState, mpcData, diagnostics=True):
    """[uk, mpcData] = iterMPC(xk,cellState,mpcData)   (iterMPC.m:1-96)"""
    Np, Nc = mpcData.Np, mpcData.Nc                                                      # :17-18
    uk_1 = mpcData.uk_1                                                                  # :19
    Ref = mpcData.ref * np.ones(Np)                                                      # :20
    mpc = cellState.MPC                                                                  # :23-26
    dx = np.concatenate([np.asarray(xk, float).ravel(), [uk_1]])                         # :29
    Phi_soc, G_soc, Aug = predMat(mpc.A, mpc.B, mpc.Csoc, mpc.Dsoc, Np, Nc)              # :33
    mpcData.Phi_soc = Phi_soc; mpcData.G_soc = G_soc                                     # :34
    F = -2 * G_soc.T @ (Ref - Phi_soc @ dx)                                              # :38
    S = np.linalg.svd(G_soc.T @ G_soc, compute_uv=False)                                 # :41
    Ru = (np.linalg.norm(F, 2) / (2 * mpcData.const.du_max * np.sqrt(Nc))) - S[-1]       # :42-43
    Ru = Ru * np.eye(Nc)                                                                 # :44
    mpcData.Ru = Ru                                                                      # :45
    E = 2 * (G_soc.T @ G_soc + Ru)                                                       # :47
    with np.errstate(all="ignore"):
        DU = -np.linalg.solve(E, F)                                                      # :48
    e = Ref - Phi_soc @ dx - G_soc @ DU
    J_unc = e @ e + DU @ Ru @ DU                                                         # :50-51
    if diagnostics:                                                                      # :54-60 (unused downstream)
        Kmpc = np.linalg.solve(E, G_soc.T @ Phi_soc)[0:1, :]
        CL = Aug.A - Aug.B @ Kmpc
        mpcData.poles = np.linalg.eigvals(CL)
        mpcData.sv = np.linalg.svd(CL, compute_uv=False)
    M, gamma = constraintsMPC(dx, cellState, mpcData)                                    # :63
    mpcData.M_last, mpcData.gamma_last = M, gamma
    if np.sum(M @ DU - gamma > 0) > 0:                                                   # :66 (Q8)
        DU, lam, nexec = hildreth(E, F, M, gamma, mpcData.lambda_, mpcData.maxHild)      # :67
        mpcData.lambda_ = lam                                                            # :68 (Q11)
        mpcData.nexec = nexec                                                            # :69
        mpcData.active = lam > 0
    else:
        mpcData.nexec = 0                                                                # :72
        mpcData.active = np.zeros(M.shape[0], bool)
    du = DU[0]
    uk = du + mpcData.uk_1                                                               # :75-76
    mpcData.uk_1 = uk; mpcData.xk_1 = xk; mpcData.DUk_1 = DU                             # :78-80
    vres = M @ DU - gamma                                                                # :82
    nviol = int(np.sum(vres > 1e-9))                                                     # :83
    e = Ref - Phi_soc @ dx - G_soc @ DU
    J_fin = e @ e + DU @ Ru @ DU                                                         # :85-86
    if hasattr(mpcData, "cost"):                                                         # :88-95
        k = mpcData.k - 1
        cst = mpcData.cost
        cst["t"][k] = (mpcData.k - 1) * mpcData.Ts
        cst["J_uncon"][k] = J_unc; cst["J_final"][k] = J_fin
        cst["norm_DU"][k] = np.linalg.norm(DU, 2); cst["viol"][k] = nviol; cst["nexec"][k] = mpcData.nexec
    return float(uk), mpcData


# ------------------------------------------------------------------------------------------
# runMPC.m (script) as a function of its hard-coded configuration
# ------------------------------------------------------------------------------------------
DEFAULT_LIMITS = dict(u_max=2.0, du_min=-50.0, du_max=50.0, v_min=3.4, v_max=4.1, phise_min=0.08)


def noise(seed, cell, k):
    """Measurement-noise sample for (seed, cell, 0-based instant): Irwin-Hall sum of twelve splitmix64 uniforms minus 6 (the
    same integer recipe as include/mek.h documents; Python integers, IEEE additions in the same order)."""
    M64 = (1 << 64) - 1
    x = (seed ^ (cell * 0x9E3779B97F4A7C15) ^ ((k + 1) * 0xD1B54A32D192ED03)) & M64
    s = 0.0
    for _ in range(12):
        x = (x + 0x9E3779B97F4A7C15) & M64
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        z = z ^ (z >> 31)
        s = s + float(z >> 11) * 2.0 ** -53
    return s - 6.0


def cell_uniform(seed, cell):
    """per-cell uniform on [0,1) of the plant-model mismatch scenario (include/mek.h): one splitmix64 step of a key of (seed, cell)"""
    M64 = (1 << 64) - 1
    z = ((seed ^ (cell * 0x9E3779B97F4A7C15) ^ 0xA5A5A5A5A5A5A5A5) + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z = z ^ (z >> 31)
    return float(z >> 11) * 2.0 ** -53


def run_mpc(bank, SOC0=10.0, TC=25.0, N=3001, Np=5, Nc=2, targetSOC=95.0, constraints=(1, 1, 1),
            Crate=2.0, limits=None, SigmaX0=None, SigmaV=1e-3, SigmaW=1e2, ekf_SOC0=None,
            diagnostics=False, keep_zk=False, blend="OutB", meas_sigma=0.0, noise_seed=0, cell_id=0,
            plant_q_scale=1.0, plant_q_spread=0.0, plant_rc_scale=1.0):
    """runMPC.m:3-115 with its literals exposed as arguments.  Returns a dict of per-step stores."""
    # not restated :3 (clear; clc; close all)
    # not restated :115 (plot4Figures: MATLAB graphics)
    # :8,10-13,16,18-19,28-30,33,36 are this function's arguments, with the script's values as defaults
    # not restated :21-22 (ekfData.maxWarn / .trueSOC0: fields nothing reads)
    # not restated :76-78 (seed entries of the stores; the loop overwrites them at k = 1)
    ROM = build_rom(bank)                                                                # :4-5
    Ts = 1                                                                               # :9
    if SigmaX0 is None:
        SigmaX0 = np.diag([1.0] * 5 + [2e6])                                             # :17
    ekfData = initKF(1 * (SOC0 if ekf_SOC0 is None else ekf_SOC0), TC, SigmaX0, SigmaV, SigmaW, blend, ROM)    # :20 (runMPC.m:16 says 'OutB')
    lim = dict(DEFAULT_LIMITS)                                                           # :37-42
    lim["z_max"] = targetSOC / 100; lim["z_tol"] = 0 * 5 / 100                           # :43-44
    if limits: lim.update(limits)
    opts = dict(Ts=Ts, constraints=list(constraints), limits=lim, ROM=ROM, Crate=Crate, Nsim=N - 1)   # :45-50
    mpcData = initMPC(SOC0, Np, Nc, targetSOC, opts)                                     # :52
    out = dict(u=np.full(N, np.nan), V=np.full(N, np.nan), SOC=np.full(N, np.nan), phise=np.full(N, np.nan),   # :55-60
               x=np.full((6, N), np.nan), iT=np.zeros(N, np.int32), iZ=np.zeros(N, np.int32),
               nexec=np.zeros(N, np.int32), active=np.zeros((N, 4), np.uint32), nbump=np.zeros(N, np.int32),
               u_applied=np.full(N, np.nan), status=0, lost_at=-1)
    if keep_zk:                                                                          # :24-25
        out["zk"] = np.full((ekfData.nz + 2, N), np.nan); out["zb"] = np.full((ekfData.nz + 2, N), np.nan)
    mpcData.cost = dict(t=np.arange(N, dtype=float), J_uncon=np.full(N, np.nan), J_final=np.full(N, np.nan),
                        norm_DU=np.full(N, np.nan), viol=np.full(N, np.nan), nexec=np.full(N, np.nan))   # :63-69
    uk = 0.0                                                                             # :72
    ROMp = ROM                                               # the plant's ROM struct: the same one in the reference's script
    if plant_q_scale != 1.0 or plant_q_spread != 0.0 or plant_rc_scale != 1.0:          # scenario: an aged cell under a nominal filter
        qs = plant_q_scale if plant_q_spread == 0.0 else plant_q_scale * (1.0 + plant_q_spread * (2.0 * cell_uniform(noise_seed, cell_id) - 1.0))
        ROMp = copy.copy(ROM); ROMp.cellData = copy.copy(ROM.cellData); ROMp.cellData.function = copy.copy(ROM.cellData.function)
        Qn, Rcn = ROM.cellData.function.const.Q(), ROM.cellData.function.const.Rc()
        ROMp.cellData.function.const = NS(Q=lambda *_: Qn * qs, Rc=lambda *_: Rcn * plant_rc_scale)
    voltage, obs, cellState = OB_step(uk, TC, None, ROMp, dict(SOC0=SOC0, warnOff=True))  # :73-74 (Q20)
    for k in range(1, N + 1):                                                            # :83-112
        voltage, obs, cellState = OB_step(uk, TC, cellState, ROMp)                       # :85
        vk = voltage if meas_sigma == 0.0 else voltage + meas_sigma * noise(noise_seed, cell_id, k - 1)   # scenario: a noisy sensor
        zk, zbk, ekfData, Xind = iterEKF(vk, uk, TC, ekfData)                            # :88
        try:
            MPC, xhat = EKFmatsHandler(ekfData, Xind, zk, TC)                            # :92
        except EKFLost:
            out["status"] |= 1; out["lost_at"] = k
            break
        cellState.MPC = MPC                                                              # :93
        out["u_applied"][k - 1] = uk
        out["phise"][k - 1] = MPC.Cphi @ xhat + uk * MPC.Dphi + MPC.bphi                 # :94-95 (Q24)
        mpcData.SOCk_1 = zk[-1]; mpcData.k = k                                           # :99,102
        uk, mpcData = iterMPC(xhat, cellState, mpcData, diagnostics=diagnostics)         # :103
        out["u"][k - 1] = uk; out["V"][k - 1] = voltage                                  # :106-111
        out["x"][:, k - 1] = xhat; out["SOC"][k - 1] = zk[-1]
        out["iT"][k - 1] = MPC.iT; out["iZ"][k - 1] = MPC.iZ
        out["nexec"][k - 1] = mpcData.nexec
        out["nbump"][k - 1] = ekfData.nbump_last
        for i, a in enumerate(mpcData.active):
            if a: out["active"][k - 1, i >> 5] |= np.uint32(1 << (i & 31))
        if keep_zk:
            out["zk"][:, k - 1] = zk; out["zb"][:, k - 1] = zbk
    out["cost"] = mpcData.cost
    out["warnCount"] = ekfData._persist.warnCount
    out["q18"] = ekfData._persist.q18
    out["plant_warnings"] = cellState.warnings
    out["final"] = dict(cellState=cellState, ekfData=ekfData, mpcData=mpcData)
    return out
