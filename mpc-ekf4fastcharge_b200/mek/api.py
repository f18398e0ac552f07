"""Batched host API over libmek (the C ABI in include/mek.h).

``RomBank``     <- load(ROMfile,'ROM')                      runMPC.m:4-5
``CellBatch``   <- cellState + ekfData + mpcData of B cells (OB_step.m:43-71, initKF.m:37-135, initMPC.m:36-72)
  .init()       <- initKF + initMPC + first OB_step call    runMPC.m:20,52,72-74
  .run(k0,k1)   <- the k-loop                               runMPC.m:83-112
  .ob_step / .iter_ekf / .iter_mpc  <- one line of the loop body each (runMPC.m:85,88-92,103)
Everything executes in the CUDA library; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MekConfig, MekError, NMATS, NSTAT, STAT_NAMES, check
from .bank import Bank, default_bank_path


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), a.shape))
    return a


class RomBank:
    """Device-side ROM bank handle (mek_bank)."""

    def __init__(self, bank: Bank | str | None = None):
        L = _lib.lib()
        self.h = C.c_void_p()
        if bank is None:
            bank = default_bank_path()
        if isinstance(bank, str):
            self.bank = Bank.load(bank)
            check(L.mek_bank_load(bank.encode(), C.byref(self.h)))
        else:
            self.bank = bank
            b = bank
            keep = dict(T=_f64(b.Tpts_C), Z=_f64(b.Zpts_pct), x=_f64(b.xloc), A=_f64(b.A.reshape(-1)), Cm=_f64(b.C.reshape(-1)),
                        D=_f64(b.D.reshape(-1)), p=_f64(b.params))
            names = (C.c_char_p * b.nz)(*[s.encode() for s in b.names])
            dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
            d = _lib.MekBankDesc(b.n, b.nz, b.nT, b.nZ, float(b.Tsamp), dp(keep["T"]), dp(keep["Z"]), names, dp(keep["x"]),
                                 dp(keep["A"]), dp(keep["Cm"]), dp(keep["D"]), dp(keep["p"]), len(b.params))
            check(L.mek_bank_create(C.byref(d), C.byref(self.h)))
        if self.bank.lut is not None and L.mek_bank_lut_points(self.h) == 0:    # tabulated cellData.function.* (mek_bank_set_lut)
            t = _f64(self.bank.lut.reshape(-1))
            check(L.mek_bank_set_lut(self.h, int(self.bank.lut.shape[1]), float(self.bank.lut_range[0]), float(self.bank.lut_range[1]),
                                     t.ctypes.data_as(C.POINTER(C.c_double))))
        self.Q = L.mek_bank_capacity_Ah(self.h)

    def close(self):
        if self.h:
            _lib.lib().mek_bank_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CellBatch:
    """B cells on one GPU (mek_pop): plant, EKF and MPC state of every cell in one SoA block."""

    def __init__(self, rom: RomBank, B, Np=5, Nc=2, constraints=(1, 1, 1), ref=95.0, SigmaV=1e-3, SigmaW=1e2, maxHild=100,
                 use_zmax=True, force_jacobi=False, device=0, threads_per_block=0, blocks_per_sm=0, generic_path=False, sync_steps=True, keep_cell_order=False,
                 defer_block=0, defer_stagger=False, defer_narrow_test=False, regroup_every=0, general_sweep_test=0, ekf_method="OutB",
                 z_target_terminal=False, z_tol=0.0, meas_sigma=0.0, noise_seed=0, first_cell=0,
                 plant_q_scale=1.0, plant_q_spread=0.0, plant_rc_scale=1.0):
        L = _lib.lib()
        self.rom = rom
        self.B = int(B)
        cfg = MekConfig()
        cfg.Np, cfg.Nc, cfg.maxHild = int(Np), int(Nc), int(maxHild)
        for i in range(3):
            cfg.constraints[i] = int(constraints[i]) if i < len(constraints) else 0
        cfg.use_zmax = int(bool(use_zmax))
        cfg.ref, cfg.SigmaV, cfg.SigmaW = float(ref), float(SigmaV), float(SigmaW)
        cfg.force_jacobi = int(bool(force_jacobi))
        cfg.threads_per_block, cfg.blocks_per_sm = int(threads_per_block), int(blocks_per_sm)
        cfg.generic_path = int(bool(generic_path))
        cfg.sync_steps = 1 if sync_steps else -1
        cfg.keep_cell_order = int(bool(keep_cell_order))
        cfg.defer_block, cfg.defer_stagger = int(defer_block), int(bool(defer_stagger))
        cfg.defer_narrow_test = int(bool(defer_narrow_test))
        cfg.general_sweep_test = int(general_sweep_test)
        cfg.regroup_every = int(regroup_every)
        cfg.z_target_terminal, cfg.z_tol = int(bool(z_target_terminal)), float(z_tol)
        cfg.meas_sigma, cfg.noise_seed, cfg.first_cell = float(meas_sigma), int(noise_seed), int(first_cell)   # noisy-sensor scenario (include/mek.h)
        cfg.plant_q_scale, cfg.plant_q_spread, cfg.plant_rc_scale = float(plant_q_scale), float(plant_q_spread), float(plant_rc_scale)   # plant-model mismatch
        blend = str(ekf_method).upper()                      # initKF.m:44-53: unknown strings fall back to output blending
        cfg.ekf_method = 1 if blend in ("MDLB", "MB") else 0
        self.cfg = cfg
        self._log_cache, self._pinned_ptrs = {}, []
        self.h = C.c_void_p()
        check(L.mek_pop_create(rom.h, C.byref(cfg), self.B, int(device), C.byref(self.h)))
        self.nC = L.mek_pop_ncon(self.h)
        self.defer_block = L.mek_pop_defer_block(self.h)      # what the library settled on (0 = off)
        self.state_len = L.mek_pop_state_len(self.h)
        self.nz = rom.bank.nz
        self.nlog = 0
        self.log_steps = 0

    def close(self):
        self._log_cache = {}
        for ptr in getattr(self, "_pinned_ptrs", []):
            _lib.lib().mek_host_free(ptr)
        self._pinned_ptrs = []
        if self.h:
            _lib.lib().mek_pop_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- set-up -------------------------------------------------------------------------------
    def set_logging(self, n_cells, max_steps):
        check(_lib.lib().mek_pop_set_logging(self.h, int(n_cells), int(max_steps)))
        self.nlog = min(int(n_cells), self.B)
        self.log_steps = int(max_steps)

    def init(self, SOC0, T, SigmaX0, limits, ekf_SOC0=None):
        """SOC0 [B] %, T [B] degC, SigmaX0 [B,6,6] (or [6,6]), limits [B,7] (or [7])."""
        B = self.B
        SOC0 = _f64(np.broadcast_to(np.asarray(SOC0, float), (B,)))
        T = _f64(np.broadcast_to(np.asarray(T, float), (B,)))
        sx = np.broadcast_to(np.asarray(SigmaX0, float), (B, 6, 6))
        if not np.array_equal(sx, np.swapaxes(sx, 1, 2)):
            raise ValueError("SigmaX0 must be symmetric")
        sx = _f64(np.swapaxes(sx, 1, 2))          # column-major 6x6 per cell
        lim = _f64(np.broadcast_to(np.asarray(limits, float), (B, 7)))
        ekf = None if ekf_SOC0 is None else _f64(np.broadcast_to(np.asarray(ekf_SOC0, float), (B,)))
        check(_lib.lib().mek_pop_init(self.h, _p(SOC0), _p(ekf), _p(T), _p(sx), _p(lim)))

    def init_raw(self, SOC0_ptr, ekf_ptr, T_ptr, sx_ptr, lim_ptr):
        """Same as init() from raw host addresses (e.g. pinned torch tensors) -- used by bench.py's e2e leg."""
        check(_lib.lib().mek_pop_init(self.h, C.c_void_p(SOC0_ptr), C.c_void_p(ekf_ptr) if ekf_ptr else None,
                                      C.c_void_p(T_ptr), C.c_void_p(sx_ptr), C.c_void_p(lim_ptr)))

    def reset(self):
        check(_lib.lib().mek_pop_reset(self.h))

    # -- the loop -----------------------------------------------------------------------------
    def run(self, k0, k1, sync=True):
        check(_lib.lib().mek_run(self.h, int(k0), int(k1)))
        if sync:
            self.sync()

    def run_timed(self, k0, k1):
        ms = C.c_float(0.0)
        check(_lib.lib().mek_run_timed(self.h, int(k0), int(k1), C.byref(ms)))
        return ms.value

    def sync(self):
        check(_lib.lib().mek_sync(self.h))

    @property
    def launches(self):
        return int(_lib.lib().mek_pop_launches(self.h))

    @property
    def regroups(self):
        """Regroupings of the state block so far (scheduling only; results do not depend on it)."""
        return int(_lib.lib().mek_pop_regroups(self.h))

    @property
    def stream(self):
        return _lib.lib().mek_pop_stream(self.h)

    def _pinned(self, shape, dtype):
        """A page-locked NumPy array owned by this batch (freed in close())."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        check(_lib.lib().mek_host_alloc(max(n, 1), C.byref(ptr)))
        self._pinned_ptrs.append(ptr)
        buf = (C.c_char * max(n, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def logs(self, k0=0, k1=None, reuse=False):
        """Trajectory logs of the first `nlog` cells for instants [k0, k1): dict of [step][cell] arrays.
        reuse=True returns page-locked arrays owned by this batch (allocated once per range, overwritten by the next call
        with the same range): the device-to-host copies then run at link speed with no page faults on fresh memory."""
        k1 = self.log_steps if k1 is None else k1
        n, nl = k1 - k0, self.nlog
        spec = (("u", (n, nl), np.float64), ("V", (n, nl), np.float64), ("SOC", (n, nl), np.float64), ("phise", (n, nl), np.float64),
                ("iT", (n, nl), np.int32), ("iZ", (n, nl), np.int32), ("nexec", (n, nl), np.int32), ("active", (n, nl, 4), np.uint32))
        if reuse:
            key = (int(k0), int(k1), nl)
            if key not in self._log_cache:
                self._log_cache[key] = {k: self._pinned(sh, dt) for k, sh, dt in spec}
            out = self._log_cache[key]
        else:
            out = {k: np.empty(sh, dt) for k, sh, dt in spec}
        check(_lib.lib().mek_read_logs(self.h, int(k0), int(k1), *[_p(out[k]) for k in ("u", "V", "SOC", "phise", "iT", "iZ", "nexec", "active")]))
        return out

    def read_state(self):
        st = np.empty((self.B, self.state_len))
        check(_lib.lib().mek_read_state(self.h, _p(st)))
        return st

    def write_state(self, st):
        st = _f64(st, (self.B, self.state_len))
        check(_lib.lib().mek_write_state(self.h, _p(st)))

    CELLSUM_NAMES = ("status", "warn_count", "sum_nexec", "n_qp", "n_cap", "min_phise", "max_V", "t_reach", "steps",
                     "last_soc", "last_V", "last_phise")

    def cell_summaries(self):
        """Per-cell summary columns (caller's cell order) as a dict of [B] arrays; see mek_read_cell_summaries."""
        a = np.empty((self.B, len(self.CELLSUM_NAMES)))
        check(_lib.lib().mek_read_cell_summaries(self.h, _p(a)))
        return {k: a[:, i].copy() for i, k in enumerate(self.CELLSUM_NAMES)}

    def summary(self):
        t = np.empty(NSTAT)
        check(_lib.lib().mek_summary(self.h, _p(t)))
        return dict(zip(STAT_NAMES, t.tolist()))

    # -- function-level entry points (one line of runMPC's loop body each) ----------------------
    def ob_step(self, Iapp):
        I = _f64(np.broadcast_to(np.asarray(Iapp, float), (self.B,)))
        V = np.empty(self.B)
        check(_lib.lib().mek_pop_ob_step(self.h, _p(I), _p(V)))
        return V

    def iter_ekf(self, vk, ik, want_zk=True, want_bounds=False):
        """[zk, boundzk, ekfData, Xind] = iterEKF(vk, ik, Tk, ekfData) + EKFmatsHandler: returns (zk, mats) or (zk, boundzk, mats)."""
        v = _f64(np.broadcast_to(np.asarray(vk, float), (self.B,)))
        i = _f64(np.broadcast_to(np.asarray(ik, float), (self.B,)))
        zk = np.empty((self.B, self.nz + 2)) if (want_zk or want_bounds) else None
        bz = np.empty((self.B, self.nz + 2)) if want_bounds else None
        mats = np.empty((self.B, NMATS))
        check(_lib.lib().mek_pop_iter_ekf(self.h, _p(v), _p(i), _p(zk), _p(bz), _p(mats)))
        return (zk, bz, mats) if want_bounds else (zk, mats)

    def iter_mpc(self, mats, uk_applied, want_cost=False):
        """iterMPC.m for every cell.  want_cost: also return mpcData.cost's columns J_uncon, J_final, norm_DU, viol [B,4]
        (iterMPC.m:50-51,82-95)."""
        mats = _f64(mats, (self.B, NMATS))
        ua = _f64(np.broadcast_to(np.asarray(uk_applied, float), (self.B,)))
        uk = np.empty(self.B); ph = np.empty(self.B); nx = np.empty(self.B, np.int32); act = np.empty((self.B, 4), np.uint32)
        if want_cost:
            cost = np.empty((self.B, 4))
            check(_lib.lib().mek_pop_iter_mpc_cost(self.h, _p(mats), _p(ua), _p(uk), _p(ph), _p(nx), _p(act), _p(cost)))
            return uk, ph, nx, act, cost
        check(_lib.lib().mek_pop_iter_mpc(self.h, _p(mats), _p(ua), _p(uk), _p(ph), _p(nx), _p(act)))
        return uk, ph, nx, act


# -- stateless batched entry points ---------------------------------------------------------------
def hildreth_batch(E, F, M, gamma, lambda0=None, maxIter=100, device=0):
    """hildreth.m over B problems: E [B,Nc,Nc], F [B,Nc], M [B,nC,Nc], gamma [B,nC] -> DU, lambda, nexec."""
    E = _f64(E); F = _f64(F); M = _f64(M); gamma = _f64(gamma)
    B, nC, Nc = M.shape
    l0 = None if lambda0 is None or np.size(lambda0) == 0 else _f64(lambda0, (B, nC))
    DU = np.empty((B, Nc)); lam = np.empty((B, nC)); nexec = np.empty(B, np.int32)
    check(_lib.lib().mek_hildreth(B, Nc, nC, _p(E), _p(F), _p(M), _p(gamma), _p(l0), int(maxIter), _p(DU), _p(lam), _p(nexec), int(device)))
    return DU, lam, nexec


def pred_mat_batch(a, c, d, Np, Nc, device=0):
    """predMat.m for scalar outputs with diagonal A (a [B,6]) and B = ones: c [B,6], d [B] -> Phi [B,Np,7], G [B,Np,Nc]."""
    a = _f64(a); c = _f64(c); d = _f64(d)
    B = a.shape[0]
    Np = max(1, int(round(Np))); Nc = max(1, min(int(round(Nc)), Np))
    Phi = np.empty((B, Np, 7)); G = np.empty((B, Np, Nc))
    check(_lib.lib().mek_pred_mat(B, _p(a), _p(c), _p(d), Np, Nc, _p(Phi), _p(G), int(device)))
    return Phi, G


def constraints_batch(x_aug, mats, SOCk_1, limits, Np, Nc, constraints, r, ref=95.0, use_zmax=True, device=0, z_target_terminal=False):
    """constraintsMPC.m (+ iterMPC.m:38-48) over B problems -> M [B,nC,Nc], gamma [B,nC], E, F, DU."""
    x_aug = _f64(x_aug); mats = _f64(mats); SOCk_1 = _f64(SOCk_1); limits = _f64(limits)
    B = x_aug.shape[0]
    Np = max(1, int(round(Np))); Nc = max(1, min(int(round(Nc)), Np))
    cons = (C.c_int * 3)(*[int(c) for c in constraints])
    nC = (4 * Nc if cons[0] == 1 else 0) + Np * (int(cons[1] == 1) + int(cons[2] == 1) + int(bool(use_zmax))) + int(bool(z_target_terminal))
    M = np.empty((B, nC, Nc)); g = np.empty((B, nC)); E = np.empty((B, Nc, Nc)); F = np.empty((B, Nc)); DU = np.empty((B, Nc))
    check(_lib.lib().mek_constraints(B, Np, Nc, cons, int(bool(use_zmax)) | (2 if z_target_terminal else 0), _p(x_aug), _p(mats), _p(SOCk_1), _p(limits), _p(M), _p(g),
                                     float(r), float(ref), _p(E), _p(F), _p(DU), int(device)))
    return M, g, E, F, DU


def selftest_spec(x, y, device=0):
    """The kernels' speculative quotient / root (csrc/mek_math.cuh) beside the standard operators on the same operands:
    returns (q_spec, q_std, s_spec, s_std, fast) for x / y and sqrt(x)."""
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    n = x.size
    out = [np.empty(n) for _ in range(4)]
    fast = np.empty(n, dtype=np.int32)
    check(_lib.lib().mek_selftest_spec(int(device), n, _p(x), _p(y), _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), fast.ctypes.data_as(C.c_void_p)))
    return out[0], out[1], out[2], out[3], fast


def fp64_peak_tflops(device=0):
    """Measured DFMA throughput (TFLOP/s, 2 flops per FMA) -- denominator of the FP64-pipe roofline."""
    t = C.c_double(0.0)
    check(_lib.lib().mek_fp64_peak(int(device), C.byref(t)))
    return t.value
