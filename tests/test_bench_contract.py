"""CPU test of bench.py's reference arm and of the JSON contract (no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-cells-per-core", "1", "--nsim", "40"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("cell-steps/sec") and d["unit"] == "cell-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_side_files_exist_and_parse():
    """bench.py reads two committed artefacts at run time: the ncu summary named by each configuration (DRAM traffic, issue
    and FP64-pipe utilisation) and the frozen flop table.  A stale name would silently null those fields on the GPU box."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    tally = json.load(open(os.path.join(ROOT, "profiles", "r01_flop_tally.json")))["configs"]
    for name, cfg in bench.CONFIGS.items():
        assert name in tally and tally[name]["F_sweep"] > 0 and tally[name]["F_fixed"] > 0
        if cfg["profile"]:
            txt = open(os.path.join(ROOT, "profiles", cfg["profile"])).read()
            assert "# DRAM traffic per cell-step" in txt
            line = [l for l in txt.splitlines() if l.startswith("# DRAM traffic per cell-step")][0]
            assert float(line.split("=")[1].split("B")[0].replace(" ", "")) > 0
            assert "smsp__issue_active.avg.pct_of_peak_sustained_active" in txt
    # the committed bench lines carry every key of the contract
    for f in ("r02_bench_C3_1Mcells.json", "r02_bench_C3v_1Mcells.json", "r02_bench_C4_65536cells.json", "r02_bench_C5_1Mcells.json"):
        d = json.load(open(os.path.join(ROOT, "profiles", f)))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                  "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
            assert k in d, (f, k)
        assert d["roofline"]["bound"] in ("hbm", "fp64"), f
        assert 0 < d["roofline"]["frac"] < 1, f
        assert d["roofline"]["traffic"] is None or d["roofline"]["traffic"] > 0, f
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
