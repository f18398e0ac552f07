"""CPU tests: the C oracle against the committed golden trajectories (made by the literal NumPy
oracle, tests/golden/make_golden.py) and the literal oracle re-run live on short horizons.

Tolerance (north_star: 1e-9 relative in FP64): |a-b| <= 1e-9 * max(|a|, scale) with the signal
scales in conftest.SCALE; discrete outcomes (local-ROM indices iT/iZ, Hildreth sweep counts,
active-set masks) must be identical."""
import glob
import os

import numpy as np
import pytest

from conftest import SCALE, rel_err

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
TOL = 1e-9


def _run_c(bank, coracle_mod, g, N=None):
    N = int(g["N"]) if N is None else N
    co = coracle_mod.COracle(bank, Np=int(g["Np"]), Nc=int(g["Nc"]), constraints=tuple(int(c) for c in g["constraints"]))
    co.init(g["SOC0"], g["T"], g["SigmaX0"], g["limits"], g["ekf_SOC0"])
    return co, co.run(0, N, nthreads=4)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_c_oracle_matches_golden(bank, coracle_mod, path):
    g = np.load(path)
    co, c = _run_c(bank, coracle_mod, g)
    for k in ("iT", "iZ", "nexec"):
        assert np.array_equal(c[k], g[k]), "%s: %d mismatches" % (k, (c[k] != g[k]).sum())
    assert np.array_equal(c["active"], g["active"])
    for k in ("u", "V", "SOC", "phise"):
        assert rel_err(c[k], g[k], SCALE[k]) <= TOL, k


def test_golden_has_the_qualitative_pins():
    """CC at u_min from the second sample, phi_se rides 0.08 V, V <= 4.1 V, SOC holds 0.95."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_111.npz"))
    u, V, soc, ph = (g[k][:, 0] for k in ("u", "V", "SOC", "phise"))
    assert abs(u[0] + 50) < 10 and abs(u[5] + 59.72) < 1e-3 and np.all(np.abs(u[5:600] + 59.72) < 1e-3)
    assert ph[1000:2400].min() > 0.079 and np.median(np.abs(ph[1000:2400] - 0.08)) < 1e-4
    assert V.max() < 4.105 and abs(soc[-1] - 0.95) < 1e-4 and soc.max() < 0.9502
    g2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_110.npz"))
    t95 = lambda s: int(np.argmax(s >= 0.9495))
    assert t95(g2["SOC"][:, 0]) < t95(soc)          # no plating constraint => faster charge (README.md:85)


def test_literal_live_matches_golden_prefix(bank):
    """Guards oracle/literal.py against edits: re-run 150 steps and compare with the stored run."""
    from oracle import literal as L
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c2_seed1.npz"))
    Q = bank.params[2]
    lim = g["limits"][1]
    o = L.run_mpc(bank, SOC0=g["SOC0"][1], TC=g["T"][1], N=150, Crate=-lim[1] / Q,
                  limits=dict(u_max=lim[0], du_max=lim[2], du_min=lim[3], v_max=lim[4], phise_min=lim[5], z_max=lim[6]),
                  SigmaX0=g["SigmaX0"][1], ekf_SOC0=g["ekf_SOC0"][1])
    for k in ("u", "V", "SOC", "phise"):
        assert np.array_equal(o[k], g[k][:150, 1]), k
    assert np.array_equal(o["nexec"], g["nexec"][:150, 1])


def test_fast_psd_path_equals_always_jacobi(bank, coracle_mod):
    """The LDL^T fast path (skip the eigen-decomposition when SigmaX is PD) must not move trajectories."""
    from mek import population as P
    pop = P.make_population(bank.params[2], 6, seed=9, heterogeneous=True, ekf_soc_error=5.0)
    outs = []
    for fj in (False, True):
        co = coracle_mod.COracle(bank, force_jacobi=fj)
        co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
        outs.append(co.run(0, 1200, nthreads=4))
    for k in ("iT", "iZ", "nexec"):
        assert np.array_equal(outs[0][k], outs[1][k])
    for k in ("u", "V", "SOC", "phise"):
        assert rel_err(outs[0][k], outs[1][k], SCALE[k]) <= TOL


def test_function_level_c_vs_literal(bank, coracle_mod):
    """OB_step / iterEKF / EKFmatsHandler one call at a time: C restatement vs literal, incl. the full zk vector."""
    from oracle import literal as L
    ROM = L.build_rom(bank)
    SigmaX0 = np.diag([1.0] * 5 + [2e6])
    kf = L.initKF(22.0, 30.0, SigmaX0, 1e-3, 1e2, "OutB", ROM)
    V, obs, cs = L.OB_step(0.0, 30.0, None, ROM, dict(SOC0=22.0, warnOff=True))
    co = coracle_mod.COracle(bank)
    co.init([22.0], 30.0, SigmaX0, coracle_mod.default_limits(bank))
    rng = np.random.default_rng(0)
    for k in range(40):
        I = float(rng.uniform(-60, 5))
        V, obs, cs = L.OB_step(I, 30.0, cs, ROM)
        Vc = co.ob_step(0, I)
        assert abs(V - Vc) <= 1e-12 * 4
        zk, zb, kf, X = L.iterEKF(V, I, 30.0, kf)
        MPC, xhat = L.EKFmatsHandler(kf, X, zk, 30.0)
        zkc, mats = co.iter_ekf(0, Vc, I)
        assert np.allclose(zk, zkc, rtol=1e-10, atol=1e-12)
        assert mats[1] == MPC.iT and mats[2] == MPC.iZ
        assert np.allclose(mats[3:9], xhat, rtol=1e-9, atol=1e-9)
        assert np.allclose(mats[14:19], MPC.Cv[:5], rtol=1e-10) and np.isclose(mats[19], MPC.Dv, rtol=1e-10)
        assert np.allclose(mats[20:25], MPC.Cphi[:5], rtol=1e-12) and mats[25] == MPC.Dphi
        assert np.isclose(mats[26], MPC.bv, rtol=1e-10) and np.isclose(mats[27], MPC.bphi, rtol=1e-12)


def test_model_blend_ekf_c_vs_literal(bank, coracle_mod):
    """iterEKF's other `blend` method ('MdlB', iterEKF.m:90-102,125-128,160-176,199-203 + the MB branches of getVariables /
    getChatV / getChatZ): one blended 6-state model per cell.  C restatement vs the literal one on a driven cell: full zk,
    the 3-sigma bounds and EKFmatsHandler's output (whose integrator entry ekfData.xhat(end) is non-zero in this mode, Q3)."""
    from oracle import literal as L
    ROM = L.build_rom(bank)
    SigmaX0 = np.diag([1.0] * 5 + [2e6])
    kf = L.initKF(22.0, 30.0, SigmaX0, 1e-3, 1e2, "MdlB", ROM)
    assert kf.method == "MB"
    V, obs, cs = L.OB_step(0.0, 30.0, None, ROM, dict(SOC0=20.0, warnOff=True))
    co = coracle_mod.COracle(bank)
    co.init([20.0], 30.0, SigmaX0, coracle_mod.default_limits(bank), [22.0])
    rng = np.random.default_rng(0)
    for k in range(250):
        I = float(rng.uniform(-60, 5)) if k < 180 else -30.0
        V, obs, cs = L.OB_step(I, 30.0, cs, ROM)
        Vc = co.ob_step(0, I)
        zk, zb, kf, X = L.iterEKF(V, I, 30.0, kf)
        MPC, xhat = L.EKFmatsHandler(kf, X, zk, 30.0)
        zkc, zbc, mats = co.iter_ekf_mb(0, Vc, I)
        assert np.allclose(zk, zkc, rtol=1e-9, atol=1e-12), k
        assert np.allclose(zb, zbc, rtol=1e-9, atol=1e-14), k
        assert mats[1] == MPC.iT and mats[2] == MPC.iZ
        assert np.allclose(mats[3:9], xhat, rtol=1e-9, atol=1e-9) and xhat[5] != 0.0 and np.all(xhat[:5] == 0.0)
        assert np.allclose(mats[14:19], MPC.Cv[:5], rtol=1e-10) and np.isclose(mats[26], MPC.bv, rtol=1e-9)
    th0, th100 = bank.params[5], bank.params[6]                   # neg.theta0 / theta100: electrode stoichiometry -> cell SOC
    assert abs(zk[-1] - (cs.SOCnAvg - th0) / (th100 - th0)) < 0.03    # the filter (started 2 % off) tracks the driven plant


def test_terminal_constraint_row_c_vs_literal(bank, coracle_mod):
    """mpcData.const.z_target_terminal (constraintsMPC.m:104-112; runMPC never sets it): one more row demanding SOC(Np) >= z_max,
    with Phi_s(end,:)*x_aug but WITHOUT the SOC_k baseline the z_max rows add (:90).  From 10 % SOC that QP is infeasible and
    the loop answers with kA-sized commands -- the reference's behaviour, reproduced, not fixed."""
    from oracle import literal as L
    N = 120
    o = L.run_mpc(bank, SOC0=10.0, N=N, limits=dict(z_target_terminal=True))
    co = coracle_mod.COracle(bank, z_target_terminal=True)
    assert co.nC == 24
    co.init([10.0], 25.0, np.diag([1.0] * 5 + [2e6]), coracle_mod.default_limits(bank))
    c = co.run(0, N, nthreads=1)
    assert np.array_equal(o["nexec"][:N], c["nexec"][:, 0]) and np.array_equal(o["active"][:N], c["active"][:, 0])
    assert (c["active"][:, 0, 0] >> 23 & 1).any()                  # the terminal row (index 23) becomes active
    for k in ("u", "V", "SOC", "phise"):
        assert rel_err(c[k][:, 0], o[k][:N], SCALE[k]) <= TOL, k
    assert np.nanmax(np.abs(c["u"])) > 1e3


def test_model_blend_loop_c_vs_literal(bank, coracle_mod):
    """The whole loop of runMPC.m:83-112 with initKF's other blend ('MdlB' instead of runMPC.m:16's 'OutB'): C restatement
    (mko_config.ekf_method = 1) vs the literal one.  In this mode EKFmatsHandler hands the MPC xhat = [0 0 0 0 0, integrator]
    (M(iT,iZ).xhat is never written by the MB filter, Q3) -- reproduced, not fixed."""
    from oracle import literal as L
    N = 1500
    o = L.run_mpc(bank, SOC0=10.0, N=N, blend="MdlB")
    assert o["status"] == 0
    co = coracle_mod.COracle(bank, ekf_method="MdlB")
    co.init([10.0], 25.0, np.diag([1.0] * 5 + [2e6]), coracle_mod.default_limits(bank))
    c = co.run(0, N, nthreads=1)
    assert np.array_equal(o["nexec"][:N], c["nexec"][:, 0]) and np.array_equal(o["active"][:N], c["active"][:, 0])
    assert np.array_equal(o["iT"][:N], c["iT"][:, 0]) and np.array_equal(o["iZ"][:N], c["iZ"][:, 0])
    for k in ("u", "V", "SOC", "phise"):
        assert rel_err(c[k][:, 0], o[k][:N], SCALE[k]) <= TOL, k
    assert c["SOC"][-1, 0] > 0.5                                    # the charge proceeds


def test_tabulated_cell_functions_c_vs_literal(bank, coracle_mod):
    """cellData.function.{Uocp,dUocp,k0,Rf} as TABLES (what a user of the real ROM file supplies: the handles sampled on a
    stoichiometry grid, mek_bank_set_lut) instead of the closed forms: C restatement vs literal restatement of the whole loop,
    and -- sanity of the interpolation -- a finely tabulated bank stays close to its closed forms."""
    from oracle import literal as L
    N = 1200
    bt = bank.with_tables(n=257)                                    # coarse on purpose: the interpolation error is visible
    o = L.run_mpc(bt, SOC0=10.0, N=N)
    co = coracle_mod.COracle(bt)
    co.init([10.0], 25.0, np.diag([1.0] * 5 + [2e6]), coracle_mod.default_limits(bt))
    c = co.run(0, N, nthreads=1)
    assert np.array_equal(o["nexec"][:N], c["nexec"][:, 0]) and np.array_equal(o["active"][:N], c["active"][:, 0])
    assert np.array_equal(o["iT"][:N], c["iT"][:, 0]) and np.array_equal(o["iZ"][:N], c["iZ"][:, 0])
    for k in ("u", "V", "SOC", "phise"):
        assert rel_err(c[k][:, 0], o[k][:N], SCALE[k]) <= TOL, k
    co0 = coracle_mod.COracle(bank)
    co0.init([10.0], 25.0, np.diag([1.0] * 5 + [2e6]), coracle_mod.default_limits(bank))
    c0 = co0.run(0, N, nthreads=1)
    assert 0 < np.abs(c["V"] - c0["V"]).max() < 5e-3                # tables are in use, and they are a fair copy
    cf = coracle_mod.COracle(bank.with_tables(n=20001))
    cf.init([10.0], 25.0, np.diag([1.0] * 5 + [2e6]), coracle_mod.default_limits(bank))
    c1 = cf.run(0, N, nthreads=1)
    assert np.abs(c1["V"] - c0["V"]).max() < 1e-5 and np.abs(c1["SOC"] - c0["SOC"]).max() < 1e-6


def test_noisy_sensor_scenario_c_vs_literal(bank, coracle_mod):
    """Measurement noise on the voltage the filter sees (a scenario beyond runMPC.m:88, which feeds the noiseless plant voltage):
    the counter-based sequence agrees bit for bit between the C and the Python code, and the loop with it agrees within TOL."""
    from oracle import literal as L
    lib = coracle_mod.lib()
    rng = np.random.default_rng(5)
    for _ in range(500):
        s, c, k = (int(x) for x in rng.integers(0, 2 ** 62, 3))
        assert lib.mko_noise(s, c, k) == L.noise(s, c, k)
    g = np.array([lib.mko_noise(11, c, k) for c in range(100) for k in range(300)])
    assert abs(g.mean()) < 0.02 and abs(g.std() - 1.0) < 0.02 and np.abs(g).max() <= 6.0
    N = 900
    o = L.run_mpc(bank, SOC0=10.0, N=N, meas_sigma=2e-3, noise_seed=77, cell_id=1005)
    co = coracle_mod.COracle(bank, meas_sigma=2e-3, noise_seed=77, first_cell=1005)
    co.init([10.0], 25.0, np.diag([1.0] * 5 + [2e6]), coracle_mod.default_limits(bank))
    c = co.run(0, N, nthreads=1)
    assert np.array_equal(o["nexec"][:N], c["nexec"][:, 0]) and np.array_equal(o["active"][:N], c["active"][:, 0])
    for k in ("u", "V", "SOC", "phise"):
        assert rel_err(c[k][:, 0], o[k][:N], SCALE[k]) <= TOL, k
    c0 = coracle_mod.COracle(bank)
    c0.init([10.0], 25.0, np.diag([1.0] * 5 + [2e6]), coracle_mod.default_limits(bank))
    q = c0.run(0, N, nthreads=1)
    assert np.abs(q["SOC"] - c["SOC"]).max() > 1e-6                 # the estimate is perturbed ...
    assert np.abs(q["SOC"] - c["SOC"]).max() < 5e-3                 # ... and the filter keeps tracking


def test_plant_model_mismatch_scenario_c_vs_literal(bank, coracle_mod):
    """An aged cell under a nominal filter (the reference: OB_step given another ROM struct than initKF / initMPC): the plant's
    capacity is Q*0.9*(1 +- 5 %) per cell and its collector resistance 1.5*Rc; the filter and the MPC keep the bank's values.
    C restatement vs literal restatement; and the scenario does what it says (the plant's SOC runs ahead of the estimate)."""
    from oracle import literal as L
    lib = coracle_mod.lib()
    rng = np.random.default_rng(6)
    for _ in range(300):
        s, c = (int(x) for x in rng.integers(0, 2 ** 62, 2))
        assert lib.mko_cell_uniform(s, c) == L.cell_uniform(s, c)
    u = np.array([lib.mko_cell_uniform(9, c) for c in range(20000)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    N = 1500
    kw = dict(plant_q_scale=0.9, plant_q_spread=0.05, plant_rc_scale=1.5, noise_seed=31)
    o = L.run_mpc(bank, SOC0=10.0, N=N, cell_id=7, **kw)
    co = coracle_mod.COracle(bank, first_cell=7, **kw)
    co.init([10.0], 25.0, np.diag([1.0] * 5 + [2e6]), coracle_mod.default_limits(bank))
    c = co.run(0, N, nthreads=1)
    assert np.array_equal(o["nexec"][:N], c["nexec"][:, 0]) and np.array_equal(o["active"][:N], c["active"][:, 0])
    for k in ("u", "V", "SOC", "phise"):
        assert rel_err(c[k][:, 0], o[k][:N], SCALE[k]) <= TOL, k
    c0 = coracle_mod.COracle(bank)
    c0.init([10.0], 25.0, np.diag([1.0] * 5 + [2e6]), coracle_mod.default_limits(bank))
    q = c0.run(0, N, nthreads=1)
    assert np.abs(q["V"] - c["V"]).max() > 1e-3                     # another plant
    th0, th100 = bank.params[5], bank.params[6]
    soc_plant = (co.state[0, 0] - th0) / (th100 - th0)              # S_SOCN -> the plant's true cell SOC
    assert soc_plant > c["SOC"][-1, 0] + 0.02                       # a smaller cell fills faster than the nominal model believes
