"""ctypes binding of libmek.so (include/mek.h).  There is no CPU fallback: if the CUDA library is
missing or no GPU is present, calls fail loudly (MekError)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MEK_LIB_PATH") or os.path.join(_HERE, "libmek.so")   # MEK_LIB_PATH: an alternative build (tuning runs)
NMATS = 31
NSTAT = 16
STAT_NAMES = ["cells", "cellsteps", "sum_nexec", "n_qp", "n_cap", "n_reached", "sum_t_reach", "min_phise", "max_V",
              "n_lost", "n_nonfinite", "sum_soc", "n_jacobi", "n_bump", "n_warn", "n_q18"]

# every symbol include/mek.h declares (tests check that the library exports all of them)
SYMBOLS = ["mek_bank_create", "mek_bank_load", "mek_bank_destroy", "mek_bank_dims", "mek_bank_capacity_Ah", "mek_bank_set_lut", "mek_bank_lut_points",
           "mek_pop_create", "mek_pop_destroy", "mek_pop_set_logging", "mek_pop_init", "mek_pop_reset", "mek_run",
           "mek_run_timed", "mek_sync", "mek_pop_stream", "mek_pop_launches", "mek_pop_ncon", "mek_pop_defer_block", "mek_pop_regroups", "mek_pop_state_len",
           "mek_read_logs", "mek_read_state", "mek_write_state", "mek_summary", "mek_pop_ob_step", "mek_pop_iter_ekf",
           "mek_pop_iter_mpc", "mek_pop_iter_mpc_cost", "mek_read_cell_summaries", "mek_host_alloc", "mek_host_free", "mek_hildreth", "mek_pred_mat", "mek_constraints", "mek_fp64_peak", "mek_selftest_spec", "mek_noise", "mek_last_error", "mek_version"]


class MekError(RuntimeError):
    pass


class MekConfig(C.Structure):
    _fields_ = [("Np", C.c_int), ("Nc", C.c_int), ("maxHild", C.c_int), ("constraints", C.c_int * 3), ("use_zmax", C.c_int),
                ("ref", C.c_double), ("SigmaV", C.c_double), ("SigmaW", C.c_double), ("force_jacobi", C.c_int),
                ("threads_per_block", C.c_int), ("blocks_per_sm", C.c_int), ("generic_path", C.c_int), ("sync_steps", C.c_int), ("keep_cell_order", C.c_int),
                ("defer_block", C.c_int), ("defer_stagger", C.c_int), ("regroup_every", C.c_int), ("defer_narrow_test", C.c_int), ("general_sweep_test", C.c_int),
                ("z_target_terminal", C.c_int), ("z_tol", C.c_double), ("ekf_method", C.c_int),
                ("meas_sigma", C.c_double), ("noise_seed", C.c_uint64), ("first_cell", C.c_int64),
                ("plant_q_scale", C.c_double), ("plant_q_spread", C.c_double), ("plant_rc_scale", C.c_double)]


class MekBankDesc(C.Structure):
    _fields_ = [("n", C.c_int), ("nz", C.c_int), ("nT", C.c_int), ("nZ", C.c_int), ("Tsamp", C.c_double),
                ("Tpts_C", C.POINTER(C.c_double)), ("Zpts_pct", C.POINTER(C.c_double)), ("names", C.POINTER(C.c_char_p)),
                ("xloc", C.POINTER(C.c_double)), ("A", C.POINTER(C.c_double)), ("C", C.POINTER(C.c_double)),
                ("D", C.POINTER(C.c_double)), ("params", C.POINTER(C.c_double)), ("nparams", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MekError("libmek.so is not built (%s); run __graft_entry__.build() or make -C mpc-ekf4fastcharge_b200/csrc. "
                           "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.mek_last_error.restype = C.c_char_p
        L.mek_bank_capacity_Ah.restype = C.c_double
        L.mek_bank_capacity_Ah.argtypes = [C.c_void_p]
        L.mek_pop_stream.restype = C.c_void_p
        L.mek_pop_stream.argtypes = [C.c_void_p]
        L.mek_pop_launches.restype = C.c_int64
        L.mek_pop_launches.argtypes = [C.c_void_p]
        L.mek_pop_ncon.argtypes = [C.c_void_p]
        L.mek_pop_defer_block.argtypes = [C.c_void_p]
        L.mek_pop_regroups.restype = C.c_int64
        L.mek_pop_regroups.argtypes = [C.c_void_p]
        L.mek_pop_state_len.argtypes = [C.c_void_p]
        L.mek_noise.restype = C.c_double
        L.mek_noise.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.mek_bank_set_lut.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.mek_bank_lut_points.argtypes = [C.c_void_p]
        L.mek_bank_destroy.argtypes = [C.c_void_p]
        L.mek_bank_destroy.restype = None
        L.mek_pop_destroy.argtypes = [C.c_void_p]
        L.mek_pop_destroy.restype = None
        L.mek_bank_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.mek_bank_create.argtypes = [C.POINTER(MekBankDesc), C.POINTER(C.c_void_p)]
        L.mek_bank_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.mek_pop_create.argtypes = [C.c_void_p, C.POINTER(MekConfig), C.c_int64, C.c_int, C.POINTER(C.c_void_p)]
        L.mek_pop_set_logging.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        L.mek_pop_init.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.mek_pop_reset.argtypes = [C.c_void_p]
        L.mek_run.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.mek_run_timed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.mek_sync.argtypes = [C.c_void_p]
        L.mek_read_logs.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8
        L.mek_read_state.argtypes = [C.c_void_p, C.c_void_p]
        L.mek_write_state.argtypes = [C.c_void_p, C.c_void_p]
        L.mek_summary.argtypes = [C.c_void_p, C.c_void_p]
        L.mek_pop_ob_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mek_pop_iter_ekf.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.mek_pop_iter_mpc.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.mek_pop_iter_mpc_cost.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.mek_read_cell_summaries.argtypes = [C.c_void_p, C.c_void_p]
        L.mek_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.mek_host_free.argtypes = [C.c_void_p]
        L.mek_hildreth.argtypes = [C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int]
        L.mek_pred_mat.argtypes = [C.c_int64] + [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 2 + [C.c_int]
        L.mek_constraints.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int] + [C.c_void_p] * 6 + [C.c_double, C.c_double] + [C.c_void_p] * 3 + [C.c_int]
        L.mek_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.mek_selftest_spec.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 7
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise MekError("libmek error %d: %s" % (rc, lib().mek_last_error().decode("utf-8", "replace")))
