"""GPU parity: the CUDA path (through the C ABI) against the C oracle on the same seeded inputs.
The bar is bit-exact for every logged quantity (integers and doubles): both sides evaluate the same
correctly rounded IEEE-754 operations in the same order (numerics contract)."""
import os

import numpy as np
import pytest

from conftest import SCALE, rel_err

pytestmark = pytest.mark.gpu


def _run_both(bank, coracle_mod, pop, N, Np=5, Nc=2, constraints=(1, 1, 1), force_jacobi=False, chunks=None, generic_path=False, **kw):
    from mek import api
    rom = api.RomBank(bank)
    cb = api.CellBatch(rom, len(pop), Np=Np, Nc=Nc, constraints=constraints, force_jacobi=force_jacobi, generic_path=generic_path, **kw)
    cb.set_logging(len(pop), N)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    if chunks is None:
        cb.run(0, N)
    else:
        for k0, k1 in chunks:
            cb.run(k0, k1)
    g = cb.logs(0, N)
    co = coracle_mod.COracle(bank, Np=Np, Nc=Nc, constraints=constraints, force_jacobi=force_jacobi)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    c = co.run(0, N, nthreads=max(8, os.cpu_count() or 8))
    return cb, g, co, c


def _assert_bit_exact(g, c):
    for k in ("iT", "iZ", "nexec"):
        assert np.array_equal(g[k], c[k]), k
    assert np.array_equal(g["active"], c["active"])
    for k in ("u", "V", "SOC", "phise"):
        same = (g[k] == c[k]) | (np.isnan(g[k]) & np.isnan(c[k]))
        assert same.all(), "%s differs in %d entries, max rel %.3e" % (k, (~same).sum(), rel_err(g[k], c[k], SCALE[k]))


def test_c1_single_cell_full_charge(bank, coracle_mod):
    from mek import population as P
    pop = P.single_cell(bank.params[2])
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 3001)
    _assert_bit_exact(g, c)
    assert g["SOC"][-1, 0] > 0.949


def test_c2_population_full_charge(bank, coracle_mod):
    """Config C2 of SURVEY.md 8(d) in full: all 4 096 cells, all 3 001 instants, every logged quantity bit for bit."""
    from mek import population as P
    B = 4096
    pop = P.make_population(bank.params[2], B, seed=1)
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 3001)
    _assert_bit_exact(g, c)
    s = cb.summary()
    assert s["cells"] == B and s["cellsteps"] == B * 3001
    assert s["sum_nexec"] == c["nexec"].sum()
    assert s["n_reached"] == B and s["n_lost"] == 0


def test_c5_heterogeneous(bank, coracle_mod):
    from mek import population as P
    pop = P.make_population(bank.params[2], 192, seed=4, heterogeneous=True, ekf_soc_error=5.0)
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 1500)
    _assert_bit_exact(g, c)


def test_c3_without_phise_constraint(bank, coracle_mod):
    from mek import population as P
    pop = P.make_population(bank.params[2], 128, seed=2)
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 2200, constraints=(1, 1, 0))
    _assert_bit_exact(g, c)


def test_c4_long_horizon(bank, coracle_mod):
    from mek import population as P
    pop = P.make_population(bank.params[2], 64, seed=3)
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 1200, Np=20, Nc=5)
    _assert_bit_exact(g, c)


def test_force_jacobi_path(bank, coracle_mod):
    from mek import population as P
    pop = P.make_population(bank.params[2], 64, seed=5)
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 400, force_jacobi=True)
    _assert_bit_exact(g, c)
    assert cb.summary()["n_jacobi"] == 64 * 400 * 2    # 25 degC sits on a set-point of the bank: two of the four neighbours carry a zero weight and are not repaired


def test_resume_in_chunks_matches_single_launch(bank, coracle_mod):
    from mek import population as P
    pop = P.make_population(bank.params[2], 96, seed=6)
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 900, chunks=[(0, 1), (1, 300), (300, 900)])
    _assert_bit_exact(g, c)


def test_generic_instance_matches_too(bank, coracle_mod):
    """The generic (runtime-shape, local-memory) instance of the fused kernel, and constraint sets without a fast instance."""
    from mek import population as P
    pop = P.make_population(bank.params[2], 96, seed=8)
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 1000, generic_path=True)
    _assert_bit_exact(g, c)
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 600, constraints=(0, 1, 1))
    _assert_bit_exact(g, c)
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 600, Np=7, Nc=3)
    _assert_bit_exact(g, c)


def test_cell_ordering_is_invisible_to_the_caller(bank):
    """mek_pop_init orders cells by (T, SOC0) inside the state block; logs, state read-back and summaries keep the caller's indices."""
    from mek import api, population as P
    pop = P.make_population(bank.params[2], 200, seed=13, heterogeneous=True)
    rom = api.RomBank(bank)
    outs = []
    for keep in (False, True):
        cb = api.CellBatch(rom, 200, keep_cell_order=keep)
        cb.set_logging(150, 500)
        cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
        cb.run(0, 500)
        outs.append((cb.logs(), cb.read_state(), cb.summary()))
    for k in ("u", "V", "SOC", "phise", "iT", "iZ", "nexec", "active"):
        assert np.array_equal(outs[0][0][k], outs[1][0][k]), k
    assert outs[0][0]["u"].shape == (500, 150)
    assert np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2]["sum_nexec"] == outs[1][2]["sum_nexec"] and outs[0][2]["min_phise"] == outs[1][2]["min_phise"]


def test_lost_filter_latch_q19(bank, coracle_mod):
    """Q19: more than 10 EKF warnings latch the filter 'lost' (iterEKF.m:55-59): NaN outputs from then on, status bit set,
    the cell idles; healthy cells in the same warp are unaffected."""
    from mek import api, population as P
    pop = P.make_population(bank.params[2], 64, seed=14)
    pop.ekf_SOC0[::7] = 150.0                       # EKF initialised at 150 % SOC: SOCnAvg > 1 on every getVariables call
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 60)
    _assert_bit_exact(g, c)
    lost = np.isnan(g["u"][-1])
    assert lost[::7].all() and not lost[1::7].any()
    first_nan = np.argmax(np.isnan(g["u"][:, 0]))
    assert 3 <= first_nan <= 8                       # two warnings per iterEKF call -> latched after ~6 calls
    s = cb.summary()
    assert s["n_lost"] == lost.sum() and s["n_warn"] >= 11 * lost.sum()
    st = cb.read_state()
    assert ((st[:, 18].astype(int) & 1) == lost).all()          # S_STATUS bit MEK_ST_LOST


# ---- deferral blocks (temporal blocking of the model stream) ----
def _oracle_state_in_gpu_layout(co, nC, M):
    """The C oracle's per-cell state rearranged into mek_read_state's layout (header 0..28, lambda, bigX, xhat, SigmaX)."""
    ncm = co.L.mko_ncon_max()
    st = co.state
    return np.concatenate([st[:, :29], st[:, 30:30 + nC], st[:, 30 + ncm:30 + ncm + 26 * M]], axis=1)


def _gpu_state_comparable(cb, M):
    st = cb.read_state()
    return np.concatenate([st[:, :29], st[:, 40:40 + cb.nC], st[:, 40 + cb.nC:40 + cb.nC + 26 * M]], axis=1)


def test_deferral_final_state_of_every_model_matches_oracle(bank, coracle_mod):
    """Logs only see the models a cell reads; the state block after the run holds all M models of every cell.  With
    deferral blocks the models outside the hot set are advanced by replay -- their final bigX / xhat / SigmaX must
    equal the instant-by-instant oracle's bit for bit (also across launches that cut blocks short)."""
    from mek import population as P
    pop = P.make_population(bank.params[2], 160, seed=21, heterogeneous=True, ekf_soc_error=3.0)
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 700, chunks=[(0, 5), (5, 37), (37, 38), (38, 700)])
    assert cb.defer_block == 16
    _assert_bit_exact(g, c)
    a, b = _gpu_state_comparable(cb, bank.M), _oracle_state_in_gpu_layout(co, cb.nC, bank.M)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    assert same.all(), "state differs in %d entries (first columns %s)" % ((~same).sum(), np.unique(np.nonzero(~same)[1])[:10])


@pytest.mark.parametrize("kblock", [2, 5, 16, 32])
def test_deferral_block_length_does_not_change_anything(bank, kblock):
    from mek import api, population as P
    pop = P.make_population(bank.params[2], 224, seed=22, heterogeneous=True, ekf_soc_error=4.0)
    rom = api.RomBank(bank)
    outs = []
    for kb in (-1, kblock):
        cb = api.CellBatch(rom, len(pop), defer_block=kb)
        assert cb.defer_block == (0 if kb < 0 else kb)
        cb.set_logging(len(pop), 450)
        cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
        cb.run(0, 211)
        cb.run(211, 450)
        st = cb.read_state()
        st[:, 29] = 0.0                               # S_SPARE: diagnostic counter of on-demand catch-ups
        outs.append((cb.logs(0, 450), st))
    for k in ("u", "V", "SOC", "phise", "iT", "iZ", "nexec", "active"):
        assert np.array_equal(outs[0][0][k], outs[1][0][k], equal_nan=k in ("u", "V", "SOC", "phise")), k
    assert np.array_equal(outs[0][1], outs[1][1], equal_nan=True)


def test_deferral_on_demand_catch_up(bank, coracle_mod):
    """Models outside the block's hot set that a cell needs mid-block are replayed on the spot (S_SPARE counts those).
    The test hook shrinks the hot set to the EKF estimate's own neighbours; the EKF starts 12-35 % SOC away from the
    truth, so the plant's set-point neighbours -- and whatever the estimate walks into -- all take that path."""
    from mek import population as P
    pop = P.make_population(bank.params[2], 96, seed=23, heterogeneous=True)
    pop.ekf_SOC0[:] = pop.SOC0 + np.where(np.arange(96) % 2 == 0, 35.0, 12.0)
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 400, chunks=[(0, 3), (3, 400)], defer_narrow_test=True)
    _assert_bit_exact(g, c)
    a, b = _gpu_state_comparable(cb, bank.M), _oracle_state_in_gpu_layout(co, cb.nC, bank.M)
    assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()
    n_on_demand = cb.read_state()[:, 29]
    assert n_on_demand.sum() >= 20, "the on-demand path was not exercised"


# ---- dynamic regrouping of a heterogeneous population ----
def test_regrouping_is_invisible_to_the_caller(bank, coracle_mod):
    """mek_run re-sorts the state block by (temperature bracket, SOC, C-rate) between launches of the fused kernel; logs,
    state read-back, summaries and a later reset keep the caller's indices and stay bit-identical to the oracle."""
    from mek import api, population as P
    pop = P.make_population(bank.params[2], 333, seed=31, heterogeneous=True, ekf_soc_error=2.0)
    rom = api.RomBank(bank)
    outs = []
    for every in (-1, 37, 0):                       # off, every 37 instants, automatic (128: the population is heterogeneous)
        cb = api.CellBatch(rom, len(pop), regroup_every=every)
        cb.set_logging(len(pop), 600)
        cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
        cb.run(0, 250)
        cb.run(250, 600)
        st = cb.read_state(); st[:, 29] = 0.0
        outs.append((cb.logs(0, 600), st, cb.summary()))
        if every == 37:                             # reset goes back to the order chosen at init and reproduces the run
            cb.reset()
            cb.run(0, 600)
            g2 = cb.logs(0, 600)
            for k in ("u", "V", "SOC", "phise", "nexec"):
                assert np.array_equal(g2[k], outs[-1][0][k], equal_nan=True), k
    for o in outs[1:]:
        for k in ("u", "V", "SOC", "phise", "iT", "iZ", "nexec", "active"):
            assert np.array_equal(outs[0][0][k], o[0][k], equal_nan=k in ("u", "V", "SOC", "phise")), k
        assert np.array_equal(outs[0][1], o[1], equal_nan=True)
        assert outs[0][2]["sum_nexec"] == o[2]["sum_nexec"] and outs[0][2]["min_phise"] == o[2]["min_phise"]
    co = coracle_mod.COracle(bank)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    _assert_bit_exact(outs[1][0], co.run(0, 600, nthreads=8))


# ---- lean Hildreth sweep of the static-shape instances ----
@pytest.mark.parametrize("constraints", [(1, 1, 1), (1, 1, 0)])
def test_lean_sweep_equals_general_sweep(bank, coracle_mod, constraints):
    """The static-shape instances run a lean sweep when the only all-zero row of M is the first z_max row, its lambda is
    finite and its K is not negative, and the general masked sweep otherwise (a QP that drives the zero row to +Inf: the
    cap-hit cases).  Both must give the oracle's bits; the test hook forces the general sweep everywhere for comparison.
    The population includes cells that start above z_max - a few percent so that the +Inf / NaN pattern occurs."""
    from mek import population as P
    pop = P.make_population(bank.params[2], 160, seed=41, heterogeneous=True, ekf_soc_error=1.0)
    pop.SOC0[::5] = 93.5                                  # reach z_max within the run: K of the zero row changes sign
    pop.ekf_SOC0[::5] = 94.5
    outs = []
    for hook in (0, 1, 2):                                # lean sweep with idle-row skipping, general sweep, lean without skipping
        cb, g, co, c = _run_both(bank, coracle_mod, pop, 700, constraints=constraints, general_sweep_test=hook)
        _assert_bit_exact(g, c)
        outs.append((g, cb.summary()))
    assert outs[0][1]["n_cap"] > 0, "no QP hit the sweep cap: the general sweep was not exercised by the lean build"
    for o in outs[1:]:
        for k in ("u", "nexec", "active"):
            assert np.array_equal(outs[0][0][k], o[0][k], equal_nan=(k == "u")), k


def test_regrouping_also_happens_when_the_run_is_chunked(bank, coracle_mod):
    """The instants since the last regrouping are counted across mek_run calls: a caller that advances the charge in chunks
    shorter than regroup_every is regrouped like one that runs it in one call (it used to never regroup)."""
    from mek import api, population as P
    pop = P.make_population(bank.params[2], 200, seed=32, heterogeneous=True)
    rom = api.RomBank(bank)
    cb = api.CellBatch(rom, len(pop), regroup_every=50)
    cb.set_logging(len(pop), 400)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    for k in range(0, 400, 20):
        cb.run(k, k + 20)
    assert cb.regroups == 7                          # after instants 50, 100, ..., 350
    one = api.CellBatch(rom, len(pop), regroup_every=50)
    one.set_logging(len(pop), 400)
    one.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    one.run(0, 400)
    assert one.regroups == 7
    co = coracle_mod.COracle(bank)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    c = co.run(0, 400, nthreads=8)
    _assert_bit_exact(cb.logs(0, 400), c)
    _assert_bit_exact(one.logs(0, 400), c)
    cb.reset()                                       # a reset starts the count again
    cb.run(0, 49)
    assert cb.regroups == 7
    cb.run(49, 51)
    assert cb.regroups == 8


# ---- the product against the LITERAL restatement (committed golden trajectories made by oracle/literal.py) ----
import glob as _glob
_GOLD = sorted(_glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", _GOLD, ids=[os.path.basename(p)[:-4] for p in _GOLD])
def test_cuda_matches_literal_golden(bank, path):
    """CUDA (through the C ABI) against the line-by-line NumPy restatement of the reference, without the C oracle in
    between: <= 1e-9 relative on u, V, SOC, phi_se (north_star's FP64 tolerance), identical local-ROM indices, Hildreth sweep
    counts and active sets.  The golden files were made by tests/golden/make_golden.py from oracle/literal.py."""
    from mek import api
    g = np.load(path)
    B, N = len(g["SOC0"]), int(g["N"])
    cb = api.CellBatch(api.RomBank(bank), B, Np=int(g["Np"]), Nc=int(g["Nc"]), constraints=tuple(int(c) for c in g["constraints"]))
    cb.set_logging(B, N)
    cb.init(g["SOC0"], g["T"], g["SigmaX0"], g["limits"], g["ekf_SOC0"])
    cb.run(0, N)
    o = cb.logs(0, N)
    for k in ("iT", "iZ", "nexec"):
        assert np.array_equal(o[k], g[k]), "%s: %d mismatches" % (k, (o[k] != g[k]).sum())
    assert np.array_equal(o["active"], g["active"])
    for k in ("u", "V", "SOC", "phise"):
        assert rel_err(o[k], g[k], SCALE[k]) <= 1e-9, k



def test_terminal_constraint_row(bank, coracle_mod):
    """The optional terminal row of constraintsMPC.m:104-112 (generic kernel instance: 24 rows at Np 5 / Nc 2)."""
    from mek import population as P
    pop = P.make_population(bank.params[2], 48, seed=61)
    cb, g, co, c = None, None, None, None
    from mek import api
    rom = api.RomBank(bank)
    cb = api.CellBatch(rom, len(pop), z_target_terminal=True)
    assert cb.nC == 24
    cb.set_logging(len(pop), 200)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    cb.run(0, 200)
    g = cb.logs(0, 200)
    co = coracle_mod.COracle(bank, z_target_terminal=True)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    c = co.run(0, 200, nthreads=8)
    _assert_bit_exact(g, c)
    assert ((g["active"][:, :, 0] >> 23) & 1).any()


def test_zero_weight_neighbours_mixed_temperatures(bank, coracle_mod):
    """Cells at temperatures ON set-points of the bank (15 / 25 / 35 degC: one temperature row of neighbours carries weights that are
    exactly zero and is left out -- gv_rows_n<2>, the plant's single row, the bump-only update) mixed, lane by lane, with cells
    between set-points (all four neighbours); a large initial SOC error makes the first residuals trip the Q-bump, which the
    zero-weight neighbours still receive; one cell starts exactly ON an SOC set-point (a zero weight on the SOC axis)."""
    from mek import population as P
    pop = P.make_population(bank.params[2], 160, seed=71, ekf_soc_error=15.0)
    pop.T[:] = np.resize(np.array([25.0, 15.0, 20.0, 35.0, 25.0, 31.5, 25.0]), len(pop.T))
    pop.SOC0[3] = 15.0; pop.ekf_SOC0[3] = 15.0
    pop.SOC0[4] = 25.0; pop.ekf_SOC0[4] = 25.0
    cb, g, co, c = _run_both(bank, coracle_mod, pop, 900, chunks=[(0, 7), (7, 900)])
    _assert_bit_exact(g, c)
    s = cb.summary()
    assert s["n_bump"] > 100          # 175 in the oracle, most of them on zero-weight neighbours (svt = SigmaV there)
    cb2, g2, co2, c2 = _run_both(bank, coracle_mod, pop, 300, generic_path=True, regroup_every=64)
    _assert_bit_exact(g2, c2)


def test_model_blend_fused_loop(bank, coracle_mod):
    """mek_run with the model-blend filter (mek_config.ekf_method = 1: a script that says initKF(...,'MdlB',...)): kernel
    k_run_population_mb against the C restatement's loop in the same mode, bit for bit; resumed in chunks."""
    from mek import population as P
    for het, seed in ((False, 81), (True, 82)):
        pop = P.make_population(bank.params[2], 96, seed=seed, heterogeneous=het, ekf_soc_error=2.0)
        from mek import api
        cb = api.CellBatch(api.RomBank(bank), len(pop), ekf_method="MdlB")
        N = 700
        cb.set_logging(len(pop), N)
        cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
        cb.run(0, 5); cb.run(5, N)
        g = cb.logs(0, N)
        co = coracle_mod.COracle(bank, ekf_method="MdlB")
        co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
        c = co.run(0, N, nthreads=8)
        _assert_bit_exact(g, c)
        assert cb.summary()["n_qp"] > 0


def test_tabulated_cell_functions(bank, coracle_mod, tmp_path):
    """A bank whose cellData.function.{Uocp,dUocp,k0,Rf} are tables (mek_bank_set_lut / bank file version 2): the fused kernels
    (k_fused_fast_lut.cu and the runtime-shape instances) against the C restatement, bit for bit; through the descriptor +
    set_lut and through a file."""
    from mek import population as P, api
    bt = bank.with_tables(n=1001)
    pop = P.make_population(bank.params[2], 96, seed=91, heterogeneous=True, ekf_soc_error=2.0)
    N = 900
    co = coracle_mod.COracle(bt)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    c = co.run(0, N, nthreads=8)
    f = str(tmp_path / "tab.bin")
    bt.save(f)
    for src, generic in ((bt, False), (f, False), (bt, True)):      # static-shape instances with the look-up compiled in; runtime-shape ones
        rom = api.RomBank(src)
        cb = api.CellBatch(rom, len(pop), generic_path=generic)
        cb.set_logging(len(pop), N)
        cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
        cb.run(0, N)
        _assert_bit_exact(cb.logs(0, N), c)
        cb.close()
    # the closed forms of the same bank give another trajectory: the tables are what ran
    cb0, g0, co0, c0 = _run_both(bank, coracle_mod, pop, N)
    assert np.nanmax(np.abs(g0["V"] - c["V"])) > 0


def test_noisy_sensor_scenario(bank, coracle_mod):
    """mek_config.meas_sigma: the filter sees V + sigma*g(seed, first_cell + cell, k).  CUDA (full-featured static-shape instances,
    runtime-shape ones and the model-blend loop) against the C restatement, bit for bit; a shard that starts at cell 40 of the
    population reproduces those cells' trajectories exactly."""
    from mek import population as P, api
    lib = api._lib.lib()
    assert lib.mek_noise(3, 4, 5) == coracle_mod.lib().mko_noise(3, 4, 5)
    pop = P.make_population(bank.params[2], 96, seed=95, ekf_soc_error=2.0)
    kw = dict(meas_sigma=1.5e-3, noise_seed=424242)
    N = 800
    cb, g, co, c = _run_both_kw(bank, coracle_mod, pop, N, {}, kw)
    _assert_bit_exact(g, c)
    cb2, g2, _, _ = _run_both_kw(bank, coracle_mod, pop, N, dict(generic_path=True), kw, oracle=False)
    _assert_bit_exact(g2, c)
    cb3, g3, co3, c3 = _run_both_kw(bank, coracle_mod, pop, 300, dict(ekf_method="MdlB"), dict(kw, ekf_method="MdlB"))
    _assert_bit_exact(g3, c3)
    sub = pop.slice(40, 96)
    cb4, g4, _, _ = _run_both_kw(bank, coracle_mod, sub, N, {}, dict(kw, first_cell=40), oracle=False)
    for k in ("u", "V", "SOC", "phise", "nexec"):
        assert np.array_equal(g4[k], g[k][:, 40:], equal_nan=(g[k].dtype.kind == "f")), k
    cb0, g0, _, _ = _run_both_kw(bank, coracle_mod, pop, N, {}, {}, oracle=False)
    assert np.nanmax(np.abs(g0["SOC"] - g["SOC"])) > 1e-6


def _run_both_kw(bank, coracle_mod, pop, N, cuda_only, both, oracle=True):
    from mek import api
    cb = api.CellBatch(api.RomBank(bank), len(pop), **dict(both, **cuda_only))
    cb.set_logging(len(pop), N)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    cb.run(0, N)
    g = cb.logs(0, N)
    co = c = None
    if oracle:
        co = coracle_mod.COracle(bank, **both)
        co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
        c = co.run(0, N, nthreads=8)
    return cb, g, co, c


def test_plant_model_mismatch_scenario(bank, coracle_mod):
    """mek_config.plant_q_scale / plant_q_spread / plant_rc_scale (an aged population under a nominal filter): CUDA -- the
    full-featured static-shape instances, the runtime-shape ones, the step-at-a-time entry -- against the C restatement,
    bit for bit; a shard reproduces its cells."""
    from mek import population as P, api
    pop = P.make_population(bank.params[2], 96, seed=97, heterogeneous=True)
    kw = dict(plant_q_scale=0.92, plant_q_spread=0.06, plant_rc_scale=1.4, noise_seed=5150)
    N = 800
    cb, g, co, c = _run_both_kw(bank, coracle_mod, pop, N, {}, kw)
    _assert_bit_exact(g, c)
    cb2, g2, _, _ = _run_both_kw(bank, coracle_mod, pop, N, dict(generic_path=True), kw, oracle=False)
    _assert_bit_exact(g2, c)
    sub = pop.slice(32, 96)
    cb3, g3, _, _ = _run_both_kw(bank, coracle_mod, sub, N, {}, dict(kw, first_cell=32), oracle=False)
    for k in ("u", "V", "SOC", "phise", "nexec"):
        assert np.array_equal(g3[k], g[k][:, 32:], equal_nan=(g[k].dtype.kind == "f")), k
    cb0, g0, _, _ = _run_both_kw(bank, coracle_mod, pop, N, {}, {}, oracle=False)
    assert np.nanmax(np.abs(g0["V"] - g["V"])) > 1e-3
    # step-at-a-time entry: the plant of mek_pop_ob_step is the mismatched one too
    cb4 = api.CellBatch(api.RomBank(bank), len(pop), **kw)
    cb4.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    V = cb4.ob_step(np.full(len(pop), -20.0))
    co4 = coracle_mod.COracle(bank, **kw)
    co4.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    assert all(V[i] == co4.ob_step(i, -20.0) for i in range(len(pop)))
