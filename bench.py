#!/usr/bin/env python
"""bench.py -- cell-steps/s of the fused ROM+EKF+MPC loop (BASELINE.json metric).

One "step" = one full charge (N_SIM sample instants, runMPC.m:83-112) of every cell of the
population: a device-side re-initialisation (initKF+initMPC+first OB_step) followed by ONE launch of
the persistent kernel.  Workload at any N: C3 of SURVEY.md 8(d) -- `--cells` cells PER GPU (weak
scaling, cells are independent), C2 distributions, constraints [1 1 1], Np=5, Nc=2.

  python bench.py --gpus 1 --steps K --warmup W            (N>1: launched under torchrun by the driver)
  python bench.py --impl reference ...                     the CPU restatement on the host cores

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mpc-ekf4fastcharge_b200"))

import numpy as np  # noqa: E402

N_SIM = 3001            # runMPC.m:10-12
METRIC = "cell-steps/sec (ROM+EKF+MPC)"
UNIT = "cell-steps/s"
SEED = 2                # C3

# The configurations BASELINE.json names.  The default bench line is C3 = the configuration the metric is
# quoted on ("cell-steps/sec at 1M cells"); the others are run with --config for the results table in BASELINE.md.
CONFIGS = {
    "C3":   dict(kernel="k_run_population<5,2,15,256,1>", profile="r02_c3_ncu_summary.csv", seed=2, Np=5, Nc=2, constraints=(1, 1, 1), het=False, cells=1 << 20, what="limits 2C/4.1V/80mV, constraints [1 1 1]"),
    "C3v":  dict(kernel="k_run_population<5,2,11,256,1>", profile=None, seed=2, Np=5, Nc=2, constraints=(1, 1, 0), het=False, cells=1 << 20, what="limits 2C/4.1V, constraints [1 1 0] (no side-reaction constraint)"),
    "C4":   dict(kernel="k_run_population<20,5,15,256,1>", profile="r02_c4_ncu_summary.csv", seed=3, Np=20, Nc=5, constraints=(1, 1, 1), het=False, cells=65536, what="extended horizon, 80 constraint rows, limits 2C/4.1V/80mV"),
    "C5":   dict(kernel="k_run_population<5,2,15,256,1>", profile=None, seed=4, Np=5, Nc=2, constraints=(1, 1, 1), het=True, cells=1 << 20, what="heterogeneous per-cell limits and temperature 15..35 C"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mek", choices=["mek", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--cells", type=int, default=0, help="cells per GPU (default: the configuration's size)")
    ap.add_argument("--nsim", type=int, default=N_SIM)
    ap.add_argument("--log-cells", type=int, default=4096, help="cells per GPU whose full trajectory is logged and read back in the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-cells-per-core", type=int, default=1024, help="CPU sample: cells per host core (full charges); ~10-15 s of CPU work on the GPU box")
    a = ap.parse_args()
    a.cfg = CONFIGS[a.config]
    if a.cells <= 0:
        a.cells = a.cfg["cells"]
    return a


def workload_name(a):
    c = a.cfg
    return ("%s: %d NMC30 cells per GPU x %d sample instants (5..30%% -> 95%% SOC), %s, Np=%d, Nc=%d, seed %d"
            % (a.config, a.cells, a.nsim, c["what"], c["Np"], c["Nc"], c["seed"]))


def config_dict(a, bank, world):
    """The `config` object of the JSON line -- the same for both arms (the driver compares them)."""
    c = a.cfg
    nC = 4 * c["Nc"] * int(c["constraints"][0] == 1) + c["Np"] * (int(c["constraints"][1] == 1) + int(c["constraints"][2] == 1) + 1)
    state_bytes = 8 * (26 * bank.M + 40 + nC)                      # SoA state block per cell
    return {"workload": workload_name(a), "cells_total": a.cells * world, "nsim": a.nsim,
            "l2": "inputs larger than L2: per-GPU state block %.2f GB of HBM-resident per-cell state" % (state_bytes * a.cells / 1e9)}


# ------------------------------------------------------------------------------------------------
# clocks: sampled during the timed region through NVML
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.sm_max = [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._loop, daemon=True)

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, attr in names.items():
                    if r & getattr(nv, attr, 0):
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        if self.nv:
            self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.nv and self.t.is_alive():
            self.t.join(1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement on the host cores (the only CPU implementation that exists:
# the reference is MATLAB and cannot run here)
# ------------------------------------------------------------------------------------------------
def cpu_run(bank, cells, nsim, nthreads, cfg, first_cell=0):
    from mek import population as P
    from oracle import coracle
    pop = P.make_population(bank.params[2], cells, cfg["seed"], heterogeneous=cfg["het"], first_cell=first_cell)
    kw = dict(Np=cfg["Np"], Nc=cfg["Nc"], constraints=cfg["constraints"])
    warm = coracle.COracle(bank, **kw)                             # spin the OpenMP team up outside the timed region
    warm.init(pop.SOC0[:nthreads], pop.T[:nthreads], pop.SigmaX0[:nthreads], pop.limits[:nthreads])
    warm.run(0, 20, logs=False, nthreads=nthreads)
    co = coracle.COracle(bank, **kw)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    t0 = time.perf_counter()
    co.run(0, nsim, logs=False, nthreads=nthreads)
    return cells * nsim / (time.perf_counter() - t0)


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mek import bank as Bk
    bank = Bk.Bank.load(Bk.default_bank_path())
    cores = os.cpu_count() or 1
    cells = max(cores * a.cpu_cells_per_core // (8 if a.cfg["Np"] > 5 else 1), cores)
    vals = []
    for i in range(a.warmup + a.steps):
        v = cpu_run(bank, cells, a.nsim, cores, a.cfg)
        if i >= a.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    sample = "%d cells x %d instants per step (first cells of the same population), C restatement oracle/mek_oracle.c, OpenMP over cells" % (cells, a.nsim)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": 1e3 * cells * a.nsim / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": config_dict(a, bank, a.gpus),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "the reference is MATLAB (no MATLAB/Octave here, ROM .mat missing): this arm times the plain-C restatement of its loop"}
    print(json.dumps(out), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main_mek(a):
    import torch
    import torch.distributed as dist
    from mek import api, bank as Bk, dist as D, population as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl mek needs a CUDA device: libmek has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    bank = Bk.Bank.load(Bk.default_bank_path())
    rom = api.RomBank(bank)
    B = a.cells
    cfg = a.cfg
    pop = P.make_population(bank.params[2], B, cfg["seed"], heterogeneous=cfg["het"], first_cell=rank * B)   # weak scaling: B cells on every rank
    # pinned host inputs (what a user hands to the ABI)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    h_soc, h_ekf, h_T = pin(pop.SOC0), pin(pop.ekf_SOC0), pin(pop.T)
    h_sx, h_lim = pin(np.swapaxes(pop.SigmaX0, 1, 2)), pin(pop.limits)
    h2d = sum(t.numel() * 8 for t in (h_soc, h_ekf, h_T, h_sx, h_lim))
    cb = api.CellBatch(rom, B, Np=cfg["Np"], Nc=cfg["Nc"], constraints=cfg["constraints"], device=local)
    nlog = min(a.log_cells, B)
    cb.set_logging(nlog, a.nsim)
    cb.init_raw(h_soc.data_ptr(), h_ekf.data_ptr(), h_T.data_ptr(), h_sx.data_ptr(), h_lim.data_ptr())
    cb.sync()

    def step():
        cb.reset()
        return cb.run_timed(0, a.nsim)

    for _ in range(a.warmup):
        step()
    # ---- timed region: K steps, device time (CUDA events on the library's stream), max over ranks ----
    sampler = ClockSampler(local)
    ext = torch.cuda.ExternalStream(cb.stream, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = cb.launches
    barrier()
    sampler.start()
    ev0.record(ext)
    kern_ms = [step() for _ in range(a.steps)]
    ev1.record(ext)
    ev1.synchronize()
    barrier()
    clocks = sampler.stop()
    launches = cb.launches - l0
    total_ms = D.max_over_ranks(ev0.elapsed_time(ev1), device=dev)
    summ = D.reduce_summaries(cb.summary(), device=dev)           # the only collective: ~16 scalars (NCCL all-reduce)
    cellsteps = float(B) * a.nsim * world
    value = cellsteps * a.steps / (total_ms * 1e-3)
    kern_avg_ms = D.max_over_ranks(float(np.mean(kern_ms)), device=dev)

    # ---- e2e: the public API with HOST buffers, H2D of the inputs and D2H of the results inside the timed region ----
    d2h = 16 * 8 + nlog * a.nsim * (4 * 8 + 3 * 4 + 16)
    cb.logs(0, a.nsim, reuse=True)                                  # one-time allocation of the page-locked result buffers (like `pin` above)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cb.init_raw(h_soc.data_ptr(), h_ekf.data_ptr(), h_T.data_ptr(), h_sx.data_ptr(), h_lim.data_ptr())
        cb.run(0, a.nsim, sync=False)
        s_e2e = cb.summary()
        lg = cb.logs(0, a.nsim, reuse=True)          # page-locked result buffers owned by the batch
    barrier()
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, device=dev)
    e2e_value = cellsteps * a.steps / e2e_s
    assert s_e2e["cellsteps"] == B * a.nsim and np.isfinite(lg["u"]).all()

    if rank == 0:
        nC = cb.nC
        M = bank.M
        state_bytes = 8 * (26 * M + 40 + nC)                       # SoA state block per cell
        kdef = cb.defer_block                                      # sample instants per deferral block (0: every model every instant)
        streamed_bytes = 2 * state_bytes + 40                      # instant-by-instant design: state once in, once out, + log record
        # with deferral every model row crosses HBM at most once each way per block of kdef instants
        bytes_per_cellstep = (2 * 8 * 26 * M) / kdef + 40 if kdef > 0 else streamed_bytes
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = bytes_per_cellstep * float(B) * a.nsim / (kern_avg_ms * 1e-3) / 1e9
        # DRAM bytes per cell-step and issue-slot / FP64-pipe utilisation measured by `ncu --set full` on this kernel
        # (profiles/r02_*_ncu_summary.csv); traffic is scaled to the units one launch of this run processes
        traffic, prof = None, {}
        try:
            for line in open(os.path.join(ROOT, "profiles", (cfg["profile"] or "none") if kdef > 0 else "r01_final_ncu_summary.csv")):
                if line.startswith("# DRAM traffic per cell-step"):
                    traffic = float(line.split("=")[1].split("B")[0].replace(" ", "")) * float(B) * a.nsim
                elif not line.startswith("#") and line.count(",") == 2:
                    k_, _, v_ = line.strip().split(",")
                    try:
                        prof[k_] = float(v_)
                    except ValueError:
                        pass
        except Exception:
            pass
        fp64_tf = api.fp64_peak_tflops(local)
        qps = summ["n_qp"] / summ["cellsteps"]
        sweeps = summ["sum_nexec"] / summ["cellsteps"]
        # FP64 roofline: exact algorithmic flops from the frozen op-count table (profiles/r01_flop_tally.json, made by
        # oracle/opcount_run.py) x this run's own QP and sweep counters, over the kernel time
        fp64 = {"peak_tflops_measured": fp64_tf}
        try:
            tl = json.load(open(os.path.join(ROOT, "profiles", "r01_flop_tally.json")))["configs"][a.config]
            fl = tl["F_fixed"] + qps * tl["F_qp_setup"] + sweeps * tl["F_sweep"]
            ins = tl["fp64_instr_per_cellstep"] + (sweeps - tl["mean_sweeps_per_step"]) * tl["F_sweep"] * (tl["blocks"][8]["fp64_instr_per_cellstep"] / tl["blocks"][8]["flops_per_cellstep"])
            rate = float(B) * a.nsim / (kern_avg_ms * 1e-3)
            fp64.update({"flops_per_cellstep": fl, "achieved_tflops": fl * rate / 1e12, "frac": fl * rate / 1e12 / fp64_tf,
                         "fp64_instr_per_cellstep": ins, "pipe_frac": ins * rate / (fp64_tf * 1e12 / 2),
                         "F_fixed": tl["F_fixed"], "F_qp_setup": tl["F_qp_setup"], "F_sweep": tl["F_sweep"],
                         "rule": "add/mul/div/sqrt = 1, fma = 2 (SURVEY.md 8d); pipe_frac counts one pipe slot per FP64 instruction against peak/2 DFMA/s"})
        except Exception:
            pass
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": config_dict(a, bank, world),
               "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "what": "mek_pop_init (pinned host inputs) + mek_run + mek_summary + mek_read_logs of %d cells' trajectories" % nlog},
               "gpu_launches": launches,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                            "traffic": traffic, "kernel": cfg["kernel"], "peak_source": "measured" if peaks else "fallback",
                            "bytes_per_cellstep": bytes_per_cellstep, "kernel_ms": kern_avg_ms, "defer_block": kdef,
                            "note": "deferral blocks took the kernel off the HBM roofline (see roofline_streamed and issue): it is bound by "
                                    "FP64 dependency latency at 2 warps per scheduler, not by bytes"},
               # the HBM bound of the instant-by-instant form of the same loop (every model streamed every instant)
               "roofline_streamed": {"bytes_per_cellstep": streamed_bytes, "bound_cellsteps_per_s": hbm_peak * 1e9 / streamed_bytes,
                                     "value_over_bound": (float(B) * a.nsim / (kern_avg_ms * 1e-3)) / (hbm_peak * 1e9 / streamed_bytes)},
               "issue": {"issue_slots_active_pct": prof.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                         "fp64_pipe_active_pct": prof.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                         "source": "ncu capture in profiles/ (37888-cell stretch of the same kernel), not measured in this run"},
               "fp64": fp64,
               "qp_solves_per_s": value * qps, "hildreth_sweeps_per_s": value * sweeps, "mean_sweeps_per_step": sweeps,
               "summary": {k: summ[k] for k in ("n_reached", "n_cap", "n_lost", "n_nonfinite", "min_phise", "max_V")}}
        if not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cells = max(cores * a.cpu_cells_per_core // (8 if cfg["Np"] > 5 else 1), cores)
            v = cpu_run(bank, cells, a.nsim, cores, cfg)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": "first %d cells of the same population x %d instants, oracle/mek_oracle.c with OpenMP over cells" % (cells, a.nsim)}
        print(json.dumps(out), file=_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that libraries write to file descriptor 1 (NCCL prints its version
    banner there) goes to stderr.  Returns a text stream on the original stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


if __name__ == "__main__":
    args = parse()
    _OUT = _claim_stdout()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_mek(args)
