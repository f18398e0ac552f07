"""ROM table bank (layer L0): container, binary file format and the canonical synthetic NMC30-like bank.

The reference loads ``ROM_NMC30_HRA.mat`` (runMPC.m:4-5), which is not shipped
(/root/reference/.MISSING_LARGE_BLOBS).  This module supplies the same field surface the hot path
reads:

  ROM.ROMmdls(iT,iZ).{A,B,C,D,T,SOC}   -> Bank.A (diag incl. integrator), Bank.C, Bank.D, Bank.T_K, Bank.Z
  ROM.tfData.{names,xLoc}              -> Bank.names, Bank.xloc
  ROM.xraData.{T,SOC,Tsamp}            -> Bank.Tpts_C, Bank.Zpts_pct, Bank.Tsamp
  ROM.cellData.const / .function.*     -> Bank.params (closed-form parameter vector, see P_* indices)

The closed forms behind ``cellData.function.*`` are *data* in the reference (MATLAB function handles
stored in the .mat file).  Here they are a parameter vector evaluated by (a) the CUDA library
(csrc/mek_cellfn.cuh), (b) the C oracle (oracle/mek_oracle.c) and (c) the literal NumPy oracle
(oracle/cellfn.py), each with its own code.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

MAGIC = b"MEKBANK1"
VERSION = 1                      # written when the bank has no tables (files of earlier builds stay valid)
VERSION_LUT = 2                  # + [int32 n][f64 lo][f64 hi][f64 tables[8][n]] behind the parameter vector
LUT_NAMES = ("Uocp_neg", "dUocp_neg", "Uocp_pos", "dUocp_pos", "k0_neg", "k0_pos", "Rf_neg", "Rf_pos")   # table order (include/mek.h)

# ---- parameter-vector layout (mirrored by csrc/mek_params.h and oracle/mek_oracle.c) ----
P_F, P_R, P_Q, P_RC, P_TREF = 0, 1, 2, 3, 4
P_NEG, P_POS = 5, 16            # electrode blocks, 11 entries each
(E_THETA0, E_THETA100, E_K0REF, E_K0SLOPE, E_EA_K0, E_RFREF, E_RFSLOPE, E_EA_RF,
 E_WDL, E_CDL, E_NDL) = range(11)
P_UN = 27                        # neg Uocp: c0..c5
P_UP = 33                        # pos Uocp: centre, width, ncoef, coef[0..7]
P_UP_COEF = 36
N_PARAMS = 44
UP_MAXCOEF = 8

ROW_NAMES_SYNTH = [
    ("negIfdl", 0), ("negIfdl", 1), ("negIf", 0), ("negIf", 1), ("negIdl", 0), ("negIdl", 1),
    ("negPhis", 1), ("negPhise", 0), ("negPhise", 1), ("negThetass", 0), ("negThetass", 1),
    ("posIfdl", 2), ("posIfdl", 3), ("posIf", 2), ("posIf", 3), ("posIdl", 2), ("posIdl", 3),
    ("posPhis", 2), ("posPhis", 3), ("posPhise", 2), ("posPhise", 3), ("posThetass", 2),
    ("posThetass", 3), ("negPhie", 1), ("posPhie", 2), ("posPhie", 3), ("negThetae", 0),
    ("negThetae", 1), ("posThetae", 2), ("posThetae", 3),
]


class BankError(ValueError):
    """Structural problem in a bank (replaces MATLAB error() at OB_step.m:140-157, initKF.m:67-87)."""


@dataclass
class Bank:
    n: int                      # transient states per local model (5; runMPC.m:17, initKF.m:60)
    nz: int                     # ROM outputs
    Tpts_C: np.ndarray          # [nT] xraData.T, deg C, ascending
    Zpts_pct: np.ndarray        # [nZ] xraData.SOC, percent, ascending
    Tsamp: float                # xraData.Tsamp
    names: list                 # tfData.names
    xloc: np.ndarray            # tfData.xLoc
    A: np.ndarray               # [nT,nZ,n+1] diag of ROMmdls.A (last entry == 1)
    C: np.ndarray               # [nT,nZ,nz,n+1] ROMmdls.C (last column = integrator residue res0)
    D: np.ndarray               # [nT,nZ,nz]
    params: np.ndarray          # [N_PARAMS]
    T_K: np.ndarray = field(default=None)   # [nT,nZ] ROMmdls.T (K)
    Z: np.ndarray = field(default=None)     # [nT,nZ] ROMmdls.SOC (0..1)
    lut: np.ndarray = field(default=None)   # optional [8,n] tables of cellData.function.* on a uniform stoichiometry grid (LUT_NAMES)
    lut_range: tuple = (0.0, 1.0)           # the grid's end points

    def __post_init__(self):
        self.Tpts_C = np.asarray(self.Tpts_C, dtype=np.float64)
        self.Zpts_pct = np.asarray(self.Zpts_pct, dtype=np.float64)
        self.xloc = np.asarray(self.xloc, dtype=np.float64)
        self.params = np.asarray(self.params, dtype=np.float64)
        nT, nZ = len(self.Tpts_C), len(self.Zpts_pct)
        if self.T_K is None:
            self.T_K = np.repeat((self.Tpts_C + 273.15)[:, None], nZ, axis=1)
        if self.Z is None:
            self.Z = np.repeat((self.Zpts_pct / 100.0)[None, :], nT, axis=0)

    @property
    def nT(self):
        return len(self.Tpts_C)

    @property
    def nZ(self):
        return len(self.Zpts_pct)

    @property
    def M(self):
        return self.nT * self.nZ

    # -- structural checks the reference performs, as one validation pass --------------------
    def validate(self):
        nT, nZ, n, nz = self.nT, self.nZ, self.n, self.nz
        if self.A.shape != (nT, nZ, n + 1) or self.C.shape != (nT, nZ, nz, n + 1) or self.D.shape != (nT, nZ, nz):
            raise BankError("bank array shapes inconsistent")
        if np.any(self.A[:, :, -1] != 1.0):                     # initKF.m:74-77
            raise BankError("A does not have integrator state")
        if np.any(np.diff(self.Tpts_C) <= 0) or np.any(np.diff(self.Zpts_pct) <= 0):
            raise BankError("set-point grids must be strictly ascending (SURVEY A.1)")
        if len(self.params) != N_PARAMS:
            raise BankError("parameter vector has wrong length")
        resolve_indices(self.names, self.xloc)                   # OB_step.m:140-158, iterEKF.m:692-734
        if self.lut is not None:
            self.lut = np.ascontiguousarray(self.lut, dtype=np.float64)
            if self.lut.ndim != 2 or self.lut.shape[0] != len(LUT_NAMES) or self.lut.shape[1] < 2 or not np.all(np.isfinite(self.lut)):
                raise BankError("cell-function tables must be a finite [8, n >= 2] array")
            if not (self.lut_range[1] > self.lut_range[0]):
                raise BankError("cell-function tables need an ascending stoichiometry range")
        return self

    def with_tables(self, n=2001, lo=0.0, hi=1.0):
        """A copy of this bank whose cellData.function.{Uocp,dUocp,k0,Rf} are TABLES sampled from its own closed forms on a
        uniform grid (what matlab/mek_sample_celldata.m does to the reference's function handles)."""
        import copy
        b = copy.copy(self)
        b.lut = sample_closed_forms(self.params, np.linspace(lo, hi, int(n)))
        b.lut_range = (float(lo), float(hi))
        return b

    # -- binary file ----------------------------------------------------------------------------
    def save(self, path):
        nT, nZ = self.nT, self.nZ
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write(struct.pack("<6i", VERSION if self.lut is None else VERSION_LUT, self.n, self.nz, nT, nZ, len(self.params)))
            f.write(struct.pack("<d", self.Tsamp))
            f.write(np.ascontiguousarray(self.Tpts_C, "<f8").tobytes())
            f.write(np.ascontiguousarray(self.Zpts_pct, "<f8").tobytes())
            for nm, xl in zip(self.names, self.xloc):
                f.write(nm.encode("ascii").ljust(16, b"\0")[:16])
                f.write(struct.pack("<d", float(xl)))
            for a in (self.T_K, self.Z, self.A, self.C, self.D, self.params):
                f.write(np.ascontiguousarray(a, "<f8").tobytes())
            if self.lut is not None:
                f.write(struct.pack("<idd", self.lut.shape[1], float(self.lut_range[0]), float(self.lut_range[1])))
                f.write(np.ascontiguousarray(self.lut, "<f8").tobytes())

    @staticmethod
    def load(path):
        with open(path, "rb") as f:
            buf = f.read()
        if buf[:8] != MAGIC:
            raise BankError("not a MEKBANK1 file")
        off = 8
        ver, n, nz, nT, nZ, npar = struct.unpack_from("<6i", buf, off); off += 24
        if ver not in (VERSION, VERSION_LUT):
            raise BankError("unsupported bank version %d" % ver)
        (Tsamp,) = struct.unpack_from("<d", buf, off); off += 8

        def take(shape):
            nonlocal off
            cnt = int(np.prod(shape))
            a = np.frombuffer(buf, "<f8", cnt, off).reshape(shape).copy(); off += 8 * cnt
            return a
        Tpts = take((nT,)); Zpts = take((nZ,))
        names, xloc = [], []
        for _ in range(nz):
            names.append(buf[off:off + 16].rstrip(b"\0").decode("ascii")); off += 16
            xloc.append(struct.unpack_from("<d", buf, off)[0]); off += 8
        T_K = take((nT, nZ)); Z = take((nT, nZ))
        A = take((nT, nZ, n + 1)); C = take((nT, nZ, nz, n + 1)); D = take((nT, nZ, nz))
        params = take((npar,))
        lut, rng = None, (0.0, 1.0)
        if ver >= VERSION_LUT:
            ln, lo, hi = struct.unpack_from("<idd", buf, off); off += 20
            if ln > 0:
                lut, rng = take((len(LUT_NAMES), ln)), (lo, hi)
        return Bank(n, nz, Tpts, Zpts, Tsamp, names, np.array(xloc), A, C, D, params, T_K, Z, lut, rng)


def resolve_indices(names, xloc):
    """0-based output-row indices used by the hot path; same rules and the same errors as
    OB_step.m:93-158 (plant) and iterEKF.m:616-734 (EKF)."""
    names = list(names); xloc = np.asarray(xloc, dtype=np.float64)

    def find(nm, loc=None):
        return [i for i, (s, x) in enumerate(zip(names, xloc)) if s == nm and (loc is None or x == loc)]
    ind = {}
    for nm in ("negIfdl", "negIf", "negIdl", "negPhis", "negPhise", "negThetass", "posIfdl", "posIf", "posIdl",
               "posPhis", "posPhise", "posThetass", "negPhie", "sepPhie", "posPhie", "negThetae", "sepThetae",
               "posThetae"):
        ind[nm] = find(nm)
    ind["Phie"] = ind["negPhie"] + ind["sepPhie"] + ind["posPhie"]
    ind["Thetae"] = ind["negThetae"] + ind["sepThetae"] + ind["posThetae"]
    ind["Ifdl"] = ind["negIfdl"] + ind["posIfdl"]
    ind["If"] = ind["negIf"] + ind["posIf"]
    ind["Thetass"] = ind["negThetass"] + ind["posThetass"]
    ind["Phise"] = ind["negPhise"] + ind["posPhise"]

    def at(key, loc):
        return [i for i in ind[key] if xloc[i] == loc]
    ind["Ifdl0"], ind["Ifdl3"] = at("Ifdl", 0), at("Ifdl", 3)
    ind["If0"], ind["If3"] = at("If", 0), at("If", 3)
    ind["Thetass0"], ind["Thetass3"] = at("Thetass", 0), at("Thetass", 3)
    ind["Phise0"] = at("Phise", 0)
    if not [i for i in ind["negIfdl"] if xloc[i] == 0]:
        raise BankError("Requires ifdl at negative collector")
    if not [i for i in ind["posIfdl"] if xloc[i] == 3]:
        raise BankError("Requires ifdl at positive collector")
    if not [i for i in ind["negIf"] if xloc[i] == 0]:
        raise BankError("Requires if at negative collector")
    if not [i for i in ind["posIf"] if xloc[i] == 3]:
        raise BankError("Requires if at positive collector")
    eps = np.finfo(float).eps
    if not ind["Thetae"] or xloc[ind["Thetae"][0]] > 0:
        raise BankError("Requires thetae at negative collector")
    if xloc[ind["Thetae"][-1]] > 3 + eps or xloc[ind["Thetae"][-1]] < 3 - eps:
        raise BankError("Requires thetae at positive collector")
    if not [i for i in ind["negThetass"] if xloc[i] == 0]:
        raise BankError("Requires thetass at negative collector")
    if not [i for i in ind["posThetass"] if xloc[i] == 3]:
        raise BankError("Requires thetass at positive collector")
    if not [i for i in ind["negPhise"] if xloc[i] == 0]:
        raise BankError("Requires phise at negative collector")
    if ind["Phie"] and xloc[ind["Phie"][0]] == 0:          # OB_step.m:151-155 / iterEKF.m:728-731
        ind["Phie"] = ind["Phie"][1:]
        ind["Phie_dropped_first"] = True
    if not ind["Phie"] or xloc[ind["Phie"][-1]] > 3 + eps or xloc[ind["Phie"][-1]] < 3 - eps:
        raise BankError("Requires phie at positive collector")
    if len(ind["negPhise"]) < 2:                           # EKFmatsHandler.m:97 indexes negPhise(2)
        raise BankError("EKFmatsHandler needs a second negPhise output (ind.negPhise(2))")
    return ind


def sample_closed_forms(params, theta):
    """The eight tabulated functions (LUT_NAMES order) of the closed-form parameter vector at stoichiometries `theta`; k0 and Rf
    at the reference temperature (Arrhenius factor 1)."""
    p = np.asarray(params, dtype=np.float64); th = np.asarray(theta, dtype=np.float64)
    c = p[P_UN:P_UN + 6]
    x = (th - c[4]) / c[5]
    un = c[0] + c[1] / ((th + c[2]) ** 2) + c[3] * (1.0 - x / np.sqrt(1.0 + x * x))
    dun = -2.0 * c[1] / ((th + c[2]) ** 3) - (c[3] / c[5]) / ((1.0 + x * x) ** 1.5)
    nc = int(p[P_UP + 2]); coef = p[P_UP_COEF:P_UP_COEF + nc]
    xp = (th - p[P_UP]) / p[P_UP + 1]
    up = np.polynomial.polynomial.polyval(xp, coef)
    dup = np.polynomial.polynomial.polyval(xp, np.polynomial.polynomial.polyder(coef)) / p[P_UP + 1]
    k0 = [p[b + E_K0REF] * (1.0 + p[b + E_K0SLOPE] * (th - 0.5)) for b in (P_NEG, P_POS)]
    rf = [p[b + E_RFREF] * (1.0 + p[b + E_RFSLOPE] * th) for b in (P_NEG, P_POS)]
    return np.stack([un, dun, up, dup, k0[0], k0[1], rf[0], rf[1]])


def default_params():
    p = np.zeros(N_PARAMS)
    p[P_F], p[P_R], p[P_Q], p[P_RC], p[P_TREF] = 96485.3321, 8.314462618, 29.86, 2.0e-5, 298.15
    neg = [0.02, 0.85, 1000.0, 0.25, 30000.0, 3.0e-4, 0.15, 20000.0, 1.0e-3, 60.0, 0.95]
    pos = [0.95, 0.40, 1200.0, -0.20, 25000.0, 2.5e-4, 0.10, 18000.0, 2.0e-3, 45.0, 0.93]
    p[P_NEG:P_NEG + 11] = neg
    p[P_POS:P_POS + 11] = pos
    # neg Uocp: c0 + c1/(th+c2)^2 + c3*(1 - x/sqrt(1+x^2)), x=(th-c4)/c5
    p[P_UN:P_UN + 6] = [0.086, 0.0015, 0.05, 0.020, 0.47, 0.035]
    # pos Uocp: polynomial in x=(th-pc)/pw fitted so that OCV(z)=Up(thp(z))-Un(thn(z)) is NMC-like
    z = np.array([0.0, 0.05, 0.10, 0.20, 0.33, 0.50, 0.60, 0.71, 0.80, 0.90, 0.95, 1.00])
    ocv = np.array([3.40, 3.47, 3.52, 3.59, 3.65, 3.70, 3.74, 3.82, 3.885, 3.99, 4.065, 4.13])
    thn = neg[0] + z * (neg[1] - neg[0])
    thp = pos[0] + z * (pos[1] - pos[0])
    c = p[P_UN:P_UN + 6]
    x = (thn - c[4]) / c[5]
    un = c[0] + c[1] / (thn + c[2]) ** 2 + c[3] * (1 - x / np.sqrt(1 + x * x))
    pc, pw = 0.675, 0.275
    xs = (thp - pc) / pw
    coef = np.polynomial.polynomial.polyfit(xs, ocv + un, 6)
    coef = np.array([float(np.float32(v)) for v in coef])      # short, exactly representable literals
    p[P_UP], p[P_UP + 1], p[P_UP + 2] = pc, pw, len(coef)
    p[P_UP_COEF:P_UP_COEF + len(coef)] = coef
    return p


def make_synthetic_bank():
    """Canonical synthetic NMC30-like bank: n=5, nT=3 (15/25/35 C),
    nZ=10 (5:10:95 %), nz=30 outputs.  Fully deterministic closed forms (no RNG)."""
    n, nz = 5, len(ROW_NAMES_SYNTH)
    Tpts = np.array([15.0, 25.0, 35.0]); Zpts = np.arange(5.0, 100.0, 10.0)
    nT, nZ = len(Tpts), len(Zpts)
    names = [r[0] for r in ROW_NAMES_SYNTH]; xloc = np.array([float(r[1]) for r in ROW_NAMES_SYNTH])
    params = default_params()
    Q = params[P_Q]
    dQn = abs(params[P_NEG + E_THETA100] - params[P_NEG + E_THETA0])
    dQp = abs(params[P_POS + E_THETA100] - params[P_POS + E_THETA0])
    tau0 = np.array([2.0, 11.0, 60.0, 420.0, 3000.0])
    # (feed-through d, DC gains g[0..4] of the five poles, integrator residue) per row, per ampere
    m = 1.0e-3
    rows = {
        ("negIfdl", 0): (0.90, [0.03, 0.04, 0.03, 0.0, 0.0], 0.0),
        ("negIfdl", 1): (1.15, [-0.05, -0.06, -0.04, 0.0, 0.0], 0.0),
        ("negIf", 0): (0.78, [0.12, 0.05, 0.05, 0.0, 0.0], 0.0),
        ("negIf", 1): (1.00, [0.10, -0.05, -0.05, 0.0, 0.0], 0.0),
        ("negIdl", 0): (0.12, [-0.09, -0.01, -0.02, 0.0, 0.0], 0.0),
        ("negIdl", 1): (0.15, [-0.15, -0.01, 0.01, 0.0, 0.0], 0.0),
        ("negPhis", 1): (1.0e-5, [2e-6, 1e-6, 0.0, 0.0, 0.0], 0.0),
        ("negPhise", 0): (0.10 * m, [0.015 * m, 0.03 * m, 0.045 * m, 0.06 * m, 0.03 * m], -2.0e-7),
        ("negPhise", 1): (0.13 * m, [0.02 * m, 0.04 * m, 0.06 * m, 0.08 * m, 0.04 * m], -2.5e-7),
        ("negThetass", 0): (-1.0e-4, [-1e-4, -2.5e-4, -4.5e-4, -3.5e-4, -1.0e-4], -dQn / (3600.0 * Q)),
        ("negThetass", 1): (-1.6e-4, [-2e-4, -4e-4, -6e-4, -4.5e-4, -1.2e-4], -dQn / (3600.0 * Q)),
        ("posIfdl", 2): (-1.15, [0.05, 0.06, 0.04, 0.0, 0.0], 0.0),
        ("posIfdl", 3): (-0.90, [-0.03, -0.04, -0.03, 0.0, 0.0], 0.0),
        ("posIf", 2): (-1.00, [-0.10, 0.05, 0.05, 0.0, 0.0], 0.0),
        ("posIf", 3): (-0.78, [-0.12, -0.05, -0.05, 0.0, 0.0], 0.0),
        ("posIdl", 2): (-0.15, [0.15, 0.01, -0.01, 0.0, 0.0], 0.0),
        ("posIdl", 3): (-0.12, [0.09, 0.01, 0.02, 0.0, 0.0], 0.0),
        ("posPhis", 2): (-1.2e-5, [-2e-6, -1e-6, 0.0, 0.0, 0.0], 0.0),
        ("posPhis", 3): (-2.0e-5, [-3e-6, -2e-6, 0.0, 0.0, 0.0], 0.0),
        ("posPhise", 2): (-0.25 * m, [-0.03 * m, -0.05 * m, -0.06 * m, -0.04 * m, -0.01 * m], 3.0e-7),
        ("posPhise", 3): (-0.18 * m, [-0.02 * m, -0.04 * m, -0.05 * m, -0.03 * m, -0.01 * m], 2.5e-7),
        ("posThetass", 2): (1.2e-4, [1.5e-4, 3e-4, 4.5e-4, 3e-4, 0.8e-4], dQp / (3600.0 * Q)),
        ("posThetass", 3): (0.8e-4, [1e-4, 2e-4, 3.2e-4, 2.4e-4, 0.6e-4], dQp / (3600.0 * Q)),
        ("negPhie", 1): (-0.03 * m, [-0.005 * m, -0.01 * m, -0.015 * m, -0.005 * m, 0.0], 0.0),
        ("posPhie", 2): (-0.06 * m, [-0.01 * m, -0.02 * m, -0.03 * m, -0.015 * m, 0.0], 0.0),
        ("posPhie", 3): (-0.15 * m, [-0.015 * m, -0.03 * m, -0.045 * m, -0.02 * m, 0.0], 0.0),
        ("negThetae", 0): (0.0, [0.0, 2e-4, 9e-4, 1.0e-3, 2e-4], 0.0),
        ("negThetae", 1): (0.0, [0.0, 1e-4, 4e-4, 5e-4, 1e-4], 0.0),
        ("posThetae", 2): (0.0, [0.0, -1e-4, -3e-4, -4e-4, -1e-4], 0.0),
        ("posThetae", 3): (0.0, [0.0, -2e-4, -7e-4, -8e-4, -2e-4], 0.0),
    }
    A = np.zeros((nT, nZ, n + 1)); C = np.zeros((nT, nZ, nz, n + 1)); D = np.zeros((nT, nZ, nz))
    for it, Tc in enumerate(Tpts):
        for iz, zp in enumerate(Zpts):
            z = zp / 100.0
            dT = Tc - 25.0
            tau = tau0 * (1.0 + 0.35 * (z - 0.5) + 0.2 * (z - 0.5) ** 2) * np.exp(-0.018 * dT)
            a = np.exp(-1.0 / tau)
            A[it, iz, :n] = a; A[it, iz, n] = 1.0
            # resistive/diffusive gains: lower when warm, bath-tub in SOC
            gs = (1.0 + 0.9 * (z - 0.55) ** 2 + 0.25 * (z - 0.5)) * np.exp(-0.014 * dT)
            for r, key in enumerate(ROW_NAMES_SYNTH):
                d, g, res0 = rows[key]
                is_flux = key[0].endswith(("Ifdl", "If", "Idl"))
                s = (1.0 + 0.05 * (z - 0.5)) if is_flux else gs
                ph = 1.0 + 0.03 * np.sin(0.7 * r + 3.0 * z)       # row-specific smooth wobble
                C[it, iz, r, :n] = np.array(g) * s * ph * (1.0 - a)
                C[it, iz, r, n] = res0
                D[it, iz, r] = d * (1.0 if is_flux else s)
    return Bank(n, nz, Tpts, Zpts, 1.0, names, xloc, A, C, D, params).validate()


def default_bank_path():
    import os
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "bank_nmc30_synth.bin")


# ------------------------------------------------------------------------------------------------
# MATLAB .mat import / export of the numeric part of the reference's ROM file
# ------------------------------------------------------------------------------------------------
def from_mat(path, params=None, var="ROM", lut_mat=None):
    """Build a Bank from the reference's ``ROM_NMC30_HRA.mat`` (loaded at runMPC.m:4-5) or any file with the same
    structure (MATLAB v5..v7; v7.3/HDF5 is not readable without h5py).

    Read: ``ROM.ROMmdls(iT,iZ).{A,B,C,D}`` (OB_step.m:171-181, initKF.m:60-92), ``ROM.xraData.{T,SOC,Tsamp}``
    (OB_step.m:50,166-167; initKF.m:132), ``ROM.tfData.{names,xLoc}`` (OB_step.m:88-89) and, when present,
    ``ROM.cellData.const.{F,R}``.  The structural checks are the reference's own (initKF.m:74-88: integrator state,
    B = ones; diagonal A as every line of the hot path assumes).

    NOT read: ``cellData.function.*`` -- MATLAB function handles, which a .mat file stores as opaque objects.  The closed
    forms are supplied as the parameter vector ``params`` (layout P_*/E_* above; ``default_params()`` when None, with F and
    R taken from the file).  Fitting that vector to a real cell's handles is the user's step: sample the handles in
    MATLAB, fit, pass the result here -- or, without fitting, sample them with matlab/mek_sample_celldata.m,
    ``save('celldata_lut.mat','-struct','lut')`` and pass that file as ``lut_mat``: Uocp, dUocp, k0 and Rf of both electrodes
    are then tables (``Bank.lut``, linear interpolation), and only the scalars of ``params`` (F, R, Q, Rc, Tref, the
    stoichiometry windows, activation energies and double-layer constants) remain to be supplied.
    """
    import scipy.io
    try:
        m = scipy.io.loadmat(path, struct_as_record=False, squeeze_me=False)
    except NotImplementedError as e:                              # v7.3
        raise BankError("cannot read %s: %s (save with -v7, or convert)" % (path, e))
    if var not in m:
        raise BankError("variable %r not found in %s" % (var, path))

    def one(x):                                                   # 1x1 struct array -> its element
        while isinstance(x, np.ndarray) and x.dtype == object and x.size == 1:
            x = x.reshape(-1)[0]
        return x
    rom = one(m[var])
    for f in ("ROMmdls", "xraData", "tfData"):
        if not hasattr(rom, f):
            raise BankError("ROM.%s missing" % f)
    xra, tf = one(rom.xraData), one(rom.tfData)
    Tpts = np.asarray(xra.T, float).reshape(-1); Zpts = np.asarray(xra.SOC, float).reshape(-1)
    Tsamp = float(np.asarray(xra.Tsamp).reshape(-1)[0])
    names = [str(one(s)) if not isinstance(one(s), np.ndarray) else "".join(one(s).reshape(-1).tolist())
             for s in np.asarray(tf.names, dtype=object).reshape(-1)]
    xloc = np.asarray(tf.xLoc, float).reshape(-1)
    mdls = np.asarray(rom.ROMmdls, dtype=object)
    nT, nZ = len(Tpts), len(Zpts)
    if mdls.size != nT * nZ:
        raise BankError("ROMmdls has %d models, xraData grids give %d x %d" % (mdls.size, nT, nZ))
    mdls = mdls.reshape(nT, nZ)
    A0 = np.asarray(mdls[0, 0].A, float)
    n = A0.shape[0] - 1
    nz = np.asarray(mdls[0, 0].C, float).shape[0]
    if len(names) != nz or len(xloc) != nz:
        raise BankError("tfData.names/xLoc do not match the %d output rows of C" % nz)
    A = np.zeros((nT, nZ, n + 1)); C = np.zeros((nT, nZ, nz, n + 1)); D = np.zeros((nT, nZ, nz))
    T_K = np.zeros((nT, nZ)); Z = np.zeros((nT, nZ))
    for i in range(nT):
        for j in range(nZ):
            md = mdls[i, j]
            Am = np.asarray(md.A, float)
            if Am.shape != (n + 1, n + 1) or np.any(Am - np.diag(np.diag(Am)) != 0.0):
                raise BankError("ROMmdls(%d,%d).A is not diagonal %dx%d" % (i + 1, j + 1, n + 1, n + 1))
            if Am[-1, -1] != 1.0:                                 # initKF.m:74-77
                raise BankError("A does not have integrator state")
            if hasattr(md, "B") and np.any(np.asarray(md.B, float).reshape(-1) != 1.0):   # initKF.m:84-88
                raise BankError("B is not all ones")
            A[i, j] = np.diag(Am)
            C[i, j] = np.asarray(md.C, float).reshape(nz, n + 1)
            D[i, j] = np.asarray(md.D, float).reshape(nz)
            T_K[i, j] = float(np.asarray(md.T).reshape(-1)[0]) if hasattr(md, "T") else Tpts[i] + 273.15
            Z[i, j] = float(np.asarray(md.SOC).reshape(-1)[0]) if hasattr(md, "SOC") else Zpts[j] / 100.0
    if params is None:
        params = default_params()
        cd = one(getattr(rom, "cellData", None))
        cst = one(getattr(cd, "const", None)) if cd is not None else None
        if cst is not None and hasattr(cst, "F") and hasattr(cst, "R"):
            params[P_F] = float(np.asarray(cst.F).reshape(-1)[0]); params[P_R] = float(np.asarray(cst.R).reshape(-1)[0])
    lut, rng = None, (0.0, 1.0)
    if lut_mat is not None:                                       # matlab/mek_sample_celldata.m: theta [1,n] uniform, tables [8,n]
        lm = scipy.io.loadmat(lut_mat)
        if "theta" not in lm or "tables" not in lm:
            raise BankError("%s: expected the variables 'theta' and 'tables' (mek_sample_celldata.m)" % lut_mat)
        th = np.asarray(lm["theta"], float).reshape(-1); lut = np.asarray(lm["tables"], float)
        if lut.shape != (len(LUT_NAMES), len(th)) or len(th) < 2:
            raise BankError("tables must be %d x numel(theta)" % len(LUT_NAMES))
        if np.max(np.abs(np.diff(th) - (th[-1] - th[0]) / (len(th) - 1))) > 1e-12 * max(1.0, abs(th[-1] - th[0])):
            raise BankError("theta must be a uniform grid")
        rng = (float(th[0]), float(th[-1]))
    return Bank(n, nz, Tpts, Zpts, Tsamp, names, xloc, A, C, D, np.asarray(params, float), T_K, Z, lut, rng).validate()


def to_mat(bank, path, var="ROM"):
    """Write the numeric part of a Bank in the structure ``from_mat`` reads (and the reference's loader expects), so a
    MATLAB user can inspect the synthetic bank; cellData.function.* cannot be written (function handles)."""
    import scipy.io
    nT, nZ, n, nz = bank.nT, bank.nZ, bank.n, bank.nz
    dt = [(k, object) for k in ("A", "B", "C", "D", "T", "SOC")]
    mdls = np.empty((nT, nZ), dtype=dt)
    for i in range(nT):
        for j in range(nZ):
            mdls[i, j] = (np.diag(bank.A[i, j]), np.ones((n + 1, 1)), bank.C[i, j], bank.D[i, j].reshape(nz, 1),
                          float(bank.T_K[i, j]), float(bank.Z[i, j]))
    rom = {"ROMmdls": mdls,
           "xraData": {"T": bank.Tpts_C.reshape(1, -1), "SOC": bank.Zpts_pct.reshape(1, -1), "Tsamp": float(bank.Tsamp)},
           "tfData": {"names": np.array(list(bank.names), dtype=object).reshape(-1, 1), "xLoc": bank.xloc.reshape(-1, 1)},
           "cellData": {"const": {"F": float(bank.params[P_F]), "R": float(bank.params[P_R])}}}
    scipy.io.savemat(path, {var: rom}, oned_as="column")
