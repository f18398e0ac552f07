"""TEST INFRASTRUCTURE (oracle) -- ctypes binding of oracle/mek_oracle.c (the structured C restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from types import SimpleNamespace as NS

import numpy as np

from . import literal

_HERE = os.path.dirname(os.path.abspath(__file__))
def _has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in line + " "
    except OSError:
        pass
    return False


_SO = os.path.join(_HERE, "_build", "libmek_oracle_fma.so" if _has_fma() else "libmek_oracle.so")
MAXROWS = 64

# state-vector field offsets (enum in mek_oracle.c)
(S_SOCN, S_SOCP, S_SOC0N, S_SOC0P, S_CDLN, S_CDLP, S_TC, S_ARRK0N, S_ARRK0P, S_ARRRFN, S_ARRRFP,
 S_X0, S_SIGX0, S_PRIORI, S_EKFSOC0, S_WARN, S_UK, S_LAMVALID, S_STATUS, S_PLANTWARN,
 S_UMAX, S_UMIN, S_DUMAX, S_DUMIN, S_VMAX, S_PHIMIN, S_ZMAX, S_NJAC, S_NBUMP, S_SPARE) = range(30)
S_LAMBDA = 30


class MkoBank(C.Structure):
    _fields_ = [("n", C.c_int), ("nz", C.c_int), ("nT", C.c_int), ("nZ", C.c_int), ("Ts", C.c_double),
                ("Tpts_K", C.POINTER(C.c_double)), ("Zpts", C.POINTER(C.c_double)), ("A", C.POINTER(C.c_double)),
                ("C", C.POINTER(C.c_double)), ("D", C.POINTER(C.c_double)), ("params", C.POINTER(C.c_double)),
                ("r_ifdl0", C.c_int), ("r_ifdl3", C.c_int), ("r_if0", C.c_int), ("r_if3", C.c_int), ("r_tss0", C.c_int),
                ("r_tss3", C.c_int), ("r_phie_last", C.c_int), ("r_phise0", C.c_int), ("r_phi", C.c_int),
                ("n_negtss", C.c_int), ("n_postss", C.c_int), ("n_thetae", C.c_int), ("n_negphise", C.c_int),
                ("n_posphise", C.c_int), ("n_phie", C.c_int), ("n_posphis", C.c_int),
                ("negtss", C.c_int * MAXROWS), ("postss", C.c_int * MAXROWS), ("thetae", C.c_int * MAXROWS),
                ("negphise", C.c_int * MAXROWS), ("posphise", C.c_int * MAXROWS), ("phie", C.c_int * MAXROWS),
                ("posphis", C.c_int * MAXROWS)]


class MkoConfig(C.Structure):
    _fields_ = [("Np", C.c_int), ("Nc", C.c_int), ("maxHild", C.c_int), ("useCurrent", C.c_int), ("useVoltage", C.c_int),
                ("useEta", C.c_int), ("useZmax", C.c_int), ("ref", C.c_double), ("SigmaV", C.c_double),
                ("SigmaW", C.c_double), ("force_jacobi", C.c_int), ("z_target_terminal", C.c_int), ("z_tol", C.c_double), ("ekf_method", C.c_int),
                ("meas_sigma", C.c_double), ("noise_seed", C.c_uint64), ("first_cell", C.c_int64),
                ("plant_q_scale", C.c_double), ("plant_q_spread", C.c_double), ("plant_rc_scale", C.c_double)]


def build(force=False):
    src = os.path.join(_HERE, "mek_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.mko_state_len.restype = C.c_int
        L.mko_ncon_max.restype = C.c_int
        L.mko_sigma_min_gtg.restype = C.c_double
        L.mko_sigma_min_gtg.argtypes = [C.c_int, C.c_int, C.c_double]
        L.mko_cell_uniform.restype = C.c_double
        L.mko_cell_uniform.argtypes = [C.c_uint64, C.c_uint64]
        L.mko_noise.restype = C.c_double
        L.mko_noise.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.mko_ob_step_s.restype = C.c_double
        L.mko_ob_step.restype = C.c_double
        L.mko_ob_step.argtypes = [C.POINTER(MkoBank), C.c_void_p, C.c_double]
        for f in ("mko_test_log", "mko_test_exp", "mko_test_asinh"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class COracle:
    """Population runner over the C restatement."""

    def __init__(self, bank, Np=5, Nc=2, constraints=(1, 1, 1), ref=95.0, SigmaV=1e-3, SigmaW=1e2, maxHild=100,
                 force_jacobi=False, z_target_terminal=False, z_tol=0.0, ekf_method="OutB", meas_sigma=0.0, noise_seed=0, first_cell=0,
                 plant_q_scale=1.0, plant_q_spread=0.0, plant_rc_scale=1.0):
        self.bank = bank
        self.L = lib()
        self._keep = dict(
            Tpts=np.ascontiguousarray(bank.Tpts_C + 273.15), Zpts=np.ascontiguousarray(bank.Zpts_pct / 100.0),
            A=np.ascontiguousarray(bank.A.reshape(-1)), Cm=np.ascontiguousarray(bank.C.reshape(-1)),
            D=np.ascontiguousarray(bank.D.reshape(-1)), params=self._params_x(bank))
        # row indices from the literal restatement's own set-up code (iterEKF.m:610-735)
        ind, loc, _ = literal._ekf_setupIndsLocs(NS(tfData=NS(names=list(bank.names), xLoc=bank.xloc)))
        b = MkoBank()
        b.n, b.nz, b.nT, b.nZ, b.Ts = bank.n, bank.nz, bank.nT, bank.nZ, float(bank.Tsamp)
        k = self._keep
        b.Tpts_K, b.Zpts, b.A, b.C, b.D, b.params = _dp(k["Tpts"]), _dp(k["Zpts"]), _dp(k["A"]), _dp(k["Cm"]), _dp(k["D"]), _dp(k["params"])
        b.r_ifdl0, b.r_ifdl3 = ind["Ifdl0"][0], ind["Ifdl3"][0]
        b.r_if0, b.r_if3 = ind["If0"][0], ind["If3"][0]
        b.r_tss0, b.r_tss3 = ind["Thetass0"][0], ind["Thetass3"][0]
        b.r_phie_last, b.r_phise0, b.r_phi = ind["Phie"][-1], ind["Phise0"][0], ind["negPhise"][1]
        for nm, key in (("negtss", "negThetass"), ("postss", "posThetass"), ("thetae", "Thetae"), ("negphise", "negPhise"),
                        ("posphise", "posPhise"), ("phie", "Phie"), ("posphis", "posPhis")):
            lst = ind[key]
            assert len(lst) <= MAXROWS
            setattr(b, "n_" + nm, len(lst))
            arr = getattr(b, nm)
            for i, v in enumerate(lst):
                arr[i] = v
        self.b = b
        cf = MkoConfig()
        cf.Np, cf.Nc, cf.maxHild = Np, Nc, maxHild
        cf.useCurrent, cf.useVoltage, cf.useEta = [int(bool(c)) for c in constraints]
        cf.useZmax = 1
        cf.ref, cf.SigmaV, cf.SigmaW = ref, SigmaV, SigmaW
        cf.force_jacobi = int(force_jacobi)
        cf.z_target_terminal, cf.z_tol = int(bool(z_target_terminal)), float(z_tol)
        cf.ekf_method = 1 if str(ekf_method).upper() in ("MDLB", "MB") else 0          # initKF.m:44-53
        cf.meas_sigma, cf.noise_seed, cf.first_cell = float(meas_sigma), int(noise_seed), int(first_cell)
        cf.plant_q_scale, cf.plant_q_spread, cf.plant_rc_scale = float(plant_q_scale), float(plant_q_spread), float(plant_rc_scale)
        self.cf = cf
        self.Np, self.Nc = Np, Nc
        self.nC = (4 * Nc if cf.useCurrent else 0) + Np * (cf.useVoltage + cf.useEta + 1) + cf.z_target_terminal
        self.SL = self.L.mko_state_len(bank.M)

    def _params_x(self, bank):
        """the parameter vector + the descriptor of the optional cell-function tables (mek_oracle.c: P_LUT_*)"""
        import struct
        px = np.zeros(len(bank.params) + 4)
        px[:len(bank.params)] = bank.params
        lut = getattr(bank, "lut", None)
        if lut is not None:
            self._lut = np.ascontiguousarray(lut, dtype=np.float64)
            n = self._lut.shape[1]
            lo, hi = bank.lut_range
            px[-4], px[-3], px[-2] = float(n), float(lo), float(n - 1) / (float(hi) - float(lo))
            px[-1] = struct.unpack("<d", struct.pack("<Q", self._lut.ctypes.data))[0]
        return px

    def init(self, SOC0pct, Tc, SigmaX0, limits, ekf_SOC0pct=None):
        """SigmaX0: [B,6,6]; limits: [B,7] = u_max,u_min,du_max,du_min,v_max,phise_min,z_max(+z_tol)."""
        SOC0pct = np.atleast_1d(np.asarray(SOC0pct, float)); B = len(SOC0pct)
        ekf = SOC0pct if ekf_SOC0pct is None else np.atleast_1d(np.asarray(ekf_SOC0pct, float))
        Tc = np.broadcast_to(np.asarray(Tc, float), (B,))
        SigmaX0 = np.broadcast_to(np.asarray(SigmaX0, float), (B, 6, 6))
        limits = np.broadcast_to(np.asarray(limits, float), (B, 7))
        st = np.zeros((B, self.SL))
        for c in range(B):
            sx = np.asfortranarray(SigmaX0[c]).ravel(order="F").copy()
            lm = np.ascontiguousarray(limits[c]).copy()
            self.L.mko_cell_init(C.byref(self.b), C.byref(self.cf), C.c_void_p(st[c].ctypes.data), C.c_double(SOC0pct[c]),
                                 C.c_double(ekf[c]), C.c_double(Tc[c]), _dp(sx), _dp(lm))
        self.state = st
        return st

    def run(self, k0, k1, logs=True, nthreads=1):
        B = self.state.shape[0]; n = k1 - k0
        if logs:
            out = dict(u=np.empty((n, B)), V=np.empty((n, B)), SOC=np.empty((n, B)), phise=np.empty((n, B)),
                       iT=np.empty((n, B), np.int32), iZ=np.empty((n, B), np.int32), nexec=np.empty((n, B), np.int32),
                       active=np.empty((n, B, 4), np.uint32))
            args = [_dp(out[k]) for k in ("u", "V", "SOC", "phise")] + \
                   [out[k].ctypes.data_as(C.c_void_p) for k in ("iT", "iZ", "nexec", "active")]
        else:
            out = {}
            args = [None] * 8
        rc = self.L.mko_run(C.byref(self.b), C.byref(self.cf), B, C.c_void_p(self.state.ctypes.data), k0, k1, *args, nthreads)
        assert rc == 0
        return out

    def ob_step(self, c, Iapp):
        return self.L.mko_ob_step_s(C.byref(self.b), C.byref(self.cf), C.c_void_p(self.state[c].ctypes.data), C.c_double(Iapp), C.c_int64(c))

    def iter_ekf(self, c, vk, ik):
        zk = np.empty(self.bank.nz + 2); mats = np.empty(31)
        self.L.mko_iter_ekf(C.byref(self.b), C.byref(self.cf), C.c_void_p(self.state[c].ctypes.data), C.c_double(vk),
                            C.c_double(ik), _dp(zk), _dp(mats))
        return zk, mats

    def iter_ekf_mb(self, c, vk, ik):
        """iterEKF with the model-blend method ('MdlB') + EKFmatsHandler for cell c: zk, boundzk, mats."""
        nzk = self.bank.nz + 2
        zk = np.empty(nzk); bz = np.empty(nzk); mats = np.empty(31)
        self.L.mko_iter_ekf_mb(C.byref(self.b), C.byref(self.cf), C.c_void_p(self.state[c].ctypes.data), C.c_double(vk),
                               C.c_double(ik), _dp(zk), _dp(bz), _dp(mats))
        return zk, bz, mats

    def hildreth(self, E, F, M, gamma, lambda0=None, maxIter=100):
        E = np.ascontiguousarray(E, float); F = np.ascontiguousarray(F, float)
        M = np.ascontiguousarray(M, float); gamma = np.ascontiguousarray(gamma, float)
        nC, Nc = M.shape
        DU = np.empty(Nc); lam = np.empty(nC); nexec = C.c_int(0)
        l0 = None if lambda0 is None or np.size(lambda0) == 0 else _dp(np.ascontiguousarray(lambda0, float))
        rc = self.L.mko_hildreth(Nc, nC, _dp(E), _dp(F), _dp(M), _dp(gamma), l0, int(maxIter), _dp(DU), _dp(lam), C.byref(nexec))
        assert rc == 0
        return DU, lam, nexec.value


def default_limits(bank, Crate=2.0, u_max=2.0, du_max=50.0, du_min=-50.0, v_max=4.1, phise_min=0.08, z_max=0.95, z_tol=0.0):
    """runMPC.m:36-44 + initMPC.m:66-67"""
    Q = bank.params[2]
    return np.array([u_max, -Q * Crate, du_max, du_min, v_max, phise_min, z_max + z_tol])
