"""GPU parity of the function-level C-ABI entry points (one line of runMPC's loop body per call) and of
the stateless batched hildreth / predMat / constraintsMPC, against the oracles."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(bank, B, seed=11, **kw):
    from mek import api, population as P
    pop = P.make_population(bank.params[2], B, seed=seed, heterogeneous=True)
    rom = api.RomBank(bank)
    cb = api.CellBatch(rom, B, **kw)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    return pop, cb


def test_stepwise_calls_reproduce_fused_run(bank):
    """OB_step -> iterEKF/EKFmatsHandler -> iterMPC called one at a time == k_run_population."""
    N, B = 120, 64
    pop, cb = _mk(bank, B)
    cb.set_logging(B, N)
    cb.run(0, N)
    ref = cb.logs(0, N)
    pop, cb2 = _mk(bank, B)
    uk = np.zeros(B)
    for k in range(N):
        V = cb2.ob_step(uk)
        zk, mats = cb2.iter_ekf(V, uk)
        unew, ph, nx, act = cb2.iter_mpc(mats, uk)
        assert np.array_equal(V, ref["V"][k]) and np.array_equal(unew, ref["u"][k])
        assert np.array_equal(ph, ref["phise"][k]) and np.array_equal(nx, ref["nexec"][k])
        assert np.array_equal(zk[:, -1], ref["SOC"][k]) and np.array_equal(act, ref["active"][k])
        assert np.array_equal(mats[:, 1].astype(np.int32), ref["iT"][k]) and np.array_equal(mats[:, 2].astype(np.int32), ref["iZ"][k])
        uk = unew


def test_function_level_vs_c_oracle_and_full_zk(bank, coracle_mod):
    B = 32
    pop, cb = _mk(bank, B)
    co = coracle_mod.COracle(bank)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    rng = np.random.default_rng(0)
    for k in range(25):
        I = rng.uniform(-60, 5, B)
        V = cb.ob_step(I)
        Vc = np.array([co.ob_step(c, I[c]) for c in range(B)])
        assert np.array_equal(V, Vc)
        zk, mats = cb.iter_ekf(V, I)
        for c in range(B):
            zkc, matsc = co.iter_ekf(c, Vc[c], I[c])
            assert np.array_equal(zk[c], zkc), (k, c)
            assert np.array_equal(mats[c], matsc), (k, c)
    st = cb.read_state()
    # state block round trip (checkpoint / resume): write it back into a fresh batch and continue identically
    pop2, cb2 = _mk(bank, B)
    cb2.write_state(st)
    assert np.array_equal(cb2.read_state(), st)
    I = rng.uniform(-60, 5, B)
    assert np.array_equal(cb.ob_step(I), cb2.ob_step(I))


def test_hildreth_batch_vs_oracles(bank, coracle_mod):
    from mek import api
    from oracle import literal as L
    co = coracle_mod.COracle(bank)
    rng = np.random.default_rng(3)
    for Nc, nC in ((2, 23), (2, 18), (5, 80), (3, 11), (1, 4), (8, 100)):
        Bq = 64
        E = np.empty((Bq, Nc, Nc)); F = np.empty((Bq, Nc)); M = np.empty((Bq, nC, Nc)); g = np.empty((Bq, nC)); l0 = np.abs(rng.normal(size=(Bq, nC)))
        for q in range(Bq):
            A = rng.normal(size=(Nc, Nc)); E[q] = A @ A.T + Nc * np.eye(Nc)
            F[q] = rng.normal(size=Nc) * 3
            M[q] = rng.normal(size=(nC, Nc)) * np.exp(rng.uniform(-3, 1, size=(nC, 1)))
            if q % 4 == 0:
                M[q, rng.integers(nC)] = 0.0                        # Q7 zero row
            g[q] = np.abs(rng.normal(size=nC)) * 0.5 + M[q] @ (-np.linalg.solve(E[q], F[q])) * rng.uniform(0.2, 0.9)
            if q % 8 == 0:
                g[q, np.all(M[q] == 0, axis=1)] = -0.1
        for warm in (None, l0):
            DU, lam, nexec = api.hildreth_batch(E, F, M, g, warm, 100)
            for q in range(Bq):
                DUc, lc, nc = co.hildreth(E[q], F[q], M[q], g[q], None if warm is None else warm[q], 100)
                assert nexec[q] == nc
                assert np.array_equal(DU[q], DUc, equal_nan=True) and np.array_equal(lam[q], lc, equal_nan=True)
            q = 1
            DUl, ll, nl = L.hildreth(E[q], F[q], M[q], g[q], None if warm is None else warm[q], 100)
            assert nl == nexec[q] and np.allclose(DUl, DU[q], rtol=1e-9, atol=1e-11)


def test_pred_mat_batch_vs_literal(bank):
    from mek import api
    from oracle import literal as L
    rng = np.random.default_rng(4)
    Bq = 40
    a = np.concatenate([rng.uniform(0.1, 0.9999, (Bq, 5)), np.ones((Bq, 1))], axis=1)
    c = rng.normal(size=(Bq, 6)); d = rng.normal(size=Bq)
    for Np, Nc in ((5, 2), (20, 5), (1, 1), (7, 9)):
        Phi, G = api.pred_mat_batch(a, c, d, Np, Nc)
        for q in range(0, Bq, 7):
            Pl, Gl, _ = L.predMat(np.diag(a[q]), np.ones((6, 1)), c[q][None], np.array([[d[q]]]), Np, Nc)
            assert Phi[q].shape == Pl.shape and G[q].shape == Gl.shape
            assert np.allclose(Phi[q], Pl, rtol=1e-12, atol=1e-14) and np.allclose(G[q], Gl, rtol=1e-12, atol=1e-14)


def test_constraints_batch_vs_literal(bank):
    """constraintsMPC + the pre-QP part of iterMPC from the same linearisation, vs the literal restatement."""
    from types import SimpleNamespace as NS
    from mek import api
    from oracle import literal as L
    B = 16
    pop, cb = _mk(bank, B)
    V = cb.ob_step(np.full(B, -30.0))
    zk, mats = cb.iter_ekf(V, np.full(B, -30.0))
    Q = bank.params[2]; r = -1.0 / (3600 * Q)
    x_aug = np.concatenate([mats[:, 3:9], np.full((B, 1), -30.0)], axis=1)
    for cons in ((1, 1, 1), (1, 1, 0), (0, 1, 1)):
        M, g, E, F, DU = api.constraints_batch(x_aug, mats, zk[:, -1], pop.limits, 5, 2, cons, r)
        for q in range(B):
            m = mats[q]; lim = pop.limits[q]
            MPC = NS(A=np.diag(np.concatenate([m[9:14], [1.0]])), B=np.ones((6, 1)), Csoc=np.array([0, 0, 0, 0, 0, r]), Dsoc=0.0,
                     Cv=np.concatenate([m[14:19], [0.0]]), Dv=m[19], Cphi=np.concatenate([m[20:25], [0.0]]), Dphi=m[25], bv=m[26], bphi=m[27])
            md = NS(Np=5, Nc=2, uk_1=-30.0, ref=95.0, SOCk_1=zk[q, -1], maxHild=100, lambda_=None, Ts=1, k=1,
                    const=NS(constraints=np.array(cons), u_max=lim[0], u_min=lim[1], du_max=lim[2], du_min=lim[3], v_max=lim[4],
                             phise_min=lim[5], z_max=lim[6], z_tol=0.0))
            cs = NS(MPC=MPC)
            L.iterMPC(m[3:9], cs, md, diagnostics=False)
            assert M[q].shape == md.M_last.shape
            assert np.allclose(M[q], md.M_last, rtol=1e-11, atol=1e-16) and np.allclose(g[q], md.gamma_last, rtol=1e-10, atol=1e-12)
            assert np.allclose(E[q], 2 * (md.G_soc.T @ md.G_soc + md.Ru), rtol=1e-10)


def test_summary_matches_logs(bank):
    N, B = 200, 96
    pop, cb = _mk(bank, B)
    cb.set_logging(B, N)
    cb.run(0, N)
    lg = cb.logs()
    s = cb.summary()
    assert s["cells"] == B and s["cellsteps"] == B * N
    assert s["sum_nexec"] == lg["nexec"].sum() and s["n_qp"] == (lg["nexec"] > 0).sum() and s["n_cap"] == (lg["nexec"] >= 100).sum()
    assert s["min_phise"] == lg["phise"].min() and s["max_V"] == lg["V"].max()
    assert abs(s["sum_soc"] - lg["SOC"][-1].sum()) < 1e-9
    assert cb.launches >= 3
    assert api_fp64_ok()


def api_fp64_ok():
    from mek import api
    tf = api.fp64_peak_tflops(0)
    return 5.0 < tf < 100.0


def test_boundzk_three_sigma_bounds_vs_literal(bank):
    """iterEKF's second output (iterEKF.m:185-205, getChatZ) from mek_pop_iter_ekf against the literal restatement."""
    from mek import api, population as P
    from oracle import literal as L
    ROM = L.build_rom(bank)
    SigmaX0 = np.diag([1.0] * 5 + [2e6])
    for SOC0, T in ((22.0, 30.0), (10.0, 25.0), (61.0, 17.0)):
        kf = L.initKF(SOC0, T, SigmaX0, 1e-3, 1e2, "OutB", ROM)
        V, obs, cs = L.OB_step(0.0, T, None, ROM, dict(SOC0=SOC0, warnOff=True))
        rom = api.RomBank(bank)
        cb = api.CellBatch(rom, 1)
        cb.init([SOC0], [T], SigmaX0, P.default_limits(bank.params[2]))
        rng = np.random.default_rng(1)
        for k in range(30):
            I = float(rng.uniform(-60, 2))
            V, obs, cs = L.OB_step(I, T, cs, ROM)
            zk, zb, kf, X = L.iterEKF(V, I, T, kf)
            cb.ob_step([I])
            gz, gb, mats = cb.iter_ekf([V], [I], want_bounds=True)
            assert np.allclose(gz[0], zk, rtol=1e-9, atol=1e-12)
            assert np.all(np.isfinite(gb[0])) and np.all(gb[0] >= 0)
            assert np.allclose(gb[0], zb, rtol=1e-8, atol=1e-13), (k, np.max(np.abs(gb[0] - zb) / np.maximum(zb, 1e-300)))


def test_iter_mpc_cost_log_vs_literal(bank):
    """mpcData.cost (iterMPC.m:50-51,82-95: J_uncon, J_final, norm_DU, viol) from mek_pop_iter_mpc_cost against the literal
    restatement's run of runMPC, through the ramp, the constant-current phase and the first constraint-riding steps."""
    from mek import api, population as P
    from oracle import literal as L
    N = 900
    lit = L.run_mpc(bank, SOC0=10.0, TC=25.0, N=N)
    rom = api.RomBank(bank)
    cb = api.CellBatch(rom, 1)
    cb.init([10.0], [25.0], np.diag([1.0] * 5 + [2e6]), P.default_limits(bank.params[2]))
    uk = np.zeros(1)
    cost = np.zeros((N, 4)); nx = np.zeros(N, np.int32); u = np.zeros(N)
    for k in range(N):
        V = cb.ob_step(uk)
        zk, mats = cb.iter_ekf(V, uk)
        uk, ph, n1, act, cj = cb.iter_mpc(mats, uk, want_cost=True)
        cost[k] = cj[0]; nx[k] = n1[0]; u[k] = uk[0]
    c = lit["cost"]
    assert np.array_equal(nx, lit["nexec"]) and np.allclose(u, lit["u"], rtol=1e-9, atol=1e-9)
    assert (nx > 0).sum() > 100 and (c["viol"][:N] > 0).sum() >= 0
    for j, key in enumerate(("J_uncon", "J_final", "norm_DU")):
        ref = c[key][:N]
        assert np.allclose(cost[:, j], ref, rtol=1e-9, atol=1e-12 * np.max(np.abs(ref))), (key, np.max(np.abs(cost[:, j] - ref) / np.abs(ref)))
    assert np.array_equal(cost[:, 3], c["viol"][:N])
    assert np.all(cost[:, 1] >= cost[:, 0] * (1 - 1e-12))          # a constrained optimum never beats the unconstrained one


def test_full_zk_with_more_than_16_rows_of_one_kind(bank, coracle_mod):
    """A bank with 19 posPhis rows (the per-kind row lists used to be cut at 16 entries): every row of zk must still get
    its correction (+vhat for posPhis, iterEKF.m:411-413) and its 3-sigma bound, bit for bit like the oracle."""
    from mek import api, bank as Bk, population as P
    extra = 17
    src = bank.names.index("posPhis") if "posPhis" in bank.names else None
    assert src is not None
    names = list(bank.names) + ["posPhis"] * extra
    xloc = np.concatenate([bank.xloc, np.full(extra, 3.0)])
    scale = 1.0 + 0.01 * np.arange(1, extra + 1)
    C = np.concatenate([bank.C, bank.C[:, :, src:src + 1, :] * scale[None, None, :, None]], axis=2)
    D = np.concatenate([bank.D, bank.D[:, :, src:src + 1] * scale[None, None, :]], axis=2)
    big = Bk.Bank(bank.n, bank.nz + extra, bank.Tpts_C, bank.Zpts_pct, bank.Tsamp, names, xloc, bank.A, C, D, bank.params).validate()
    B = 16
    pop = P.make_population(big.params[2], B, seed=12, heterogeneous=True)
    cb = api.CellBatch(api.RomBank(big), B)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    co = coracle_mod.COracle(big)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    rng = np.random.default_rng(5)
    for k in range(6):
        I = rng.uniform(-60, 5, B)
        V = cb.ob_step(I)
        zk, mats = cb.iter_ekf(V, I)
        for c in range(B):
            assert V[c] == co.ob_step(c, I[c])
            zkc, matsc = co.iter_ekf(c, V[c], I[c])
            assert np.array_equal(zk[c], zkc), (k, c, np.nonzero(zk[c] != zkc)[0])
    # the duplicated rows differ from their source row only by the linear part's scale: all carry +vhat
    vhat = zk[:, big.nz]
    lin_src = zk[:, src] - vhat
    for j in range(extra):
        assert np.allclose(zk[:, bank.nz + j] - vhat, lin_src * scale[j], rtol=1e-9, atol=1e-15)


def test_kelvin_temperature_is_rejected(bank):
    """initKF.m:38-41 / iterEKF.m:62-66 guess Kelvin when T > 100 (OB_step.m:75 and EKFmatsHandler.m:53 do not): libmek takes
    degrees Celsius only and refuses anything else instead of silently running a different filter than the reference."""
    from mek import api, population as P
    pop = P.make_population(bank.params[2], 8, seed=3)
    cb = api.CellBatch(api.RomBank(bank), 8)
    T = pop.T.copy(); T[5] = 298.15
    with pytest.raises(api.MekError, match="Celsius"):
        cb.init(pop.SOC0, T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)        # the batch is still usable
    assert np.isfinite(cb.ob_step(np.zeros(8))).all()


def test_model_blend_ekf_vs_c_oracle(bank, coracle_mod):
    """mek_pop_iter_ekf of a population created with ekf_method = 'MdlB' (initKF's blend argument): zk, boundzk and mats
    bit for bit like the C restatement of iterEKF.m's MB branches; the fused loop refuses this method."""
    from mek import api, population as P
    B = 24
    pop = P.make_population(bank.params[2], B, seed=15, heterogeneous=True, ekf_soc_error=3.0)
    cb = api.CellBatch(api.RomBank(bank), B, ekf_method="MdlB")
    cb.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    co = coracle_mod.COracle(bank)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    rng = np.random.default_rng(6)
    for k in range(40):
        I = rng.uniform(-60, 5, B)
        V = cb.ob_step(I)
        zk, bz, mats = cb.iter_ekf(V, I, want_bounds=True)
        for c in range(B):
            assert V[c] == co.ob_step(c, I[c])
            zkc, bzc, matsc = co.iter_ekf_mb(c, V[c], I[c])
            assert np.array_equal(zk[c], zkc), (k, c)
            assert np.array_equal(bz[c], bzc, equal_nan=True), (k, c)
            assert np.array_equal(mats[c], matsc), (k, c)
    assert np.all(mats[:, 8] != 0.0) and np.all(mats[:, 3:8] == 0.0)     # EKFmatsHandler.m:33 in MB mode
    # the state block carries the blended filter behind the model rows; a round trip preserves it
    st = cb.read_state()
    assert st.shape[1] == cb.state_len and cb.state_len == 40 + cb.nC + 26 * bank.M + 27
    cb.run(0, 5)                                          # the loop runs in this mode too (tests/test_gpu_parity.py::test_model_blend_fused_loop)
    # and the method names of initKF.m:44-53
    assert api.CellBatch(api.RomBank(bank), 2, ekf_method="nonsense").cfg.ekf_method == 0


def test_speculative_quotient_and_root_are_the_standard_operators():
    """csrc/mek_math.cuh div_spec / sqrt_spec (the fast sequences the kernels run side by side, with a replay by the standard
    operators when their exponent tests reject the operands) against `/` and `sqrt` on the device: bit for bit on operands over
    the whole exponent range, and on zeros, infinities, NaNs, subnormals and negative radicands."""
    from mek import api
    rng = np.random.default_rng(20261019)
    n = 1 << 20
    def wide(m):
        return np.ldexp(rng.uniform(1.0, 2.0, m), rng.integers(-1070, 1023, m)) * rng.choice([-1.0, 1.0], m)
    x = np.concatenate([wide(n), rng.uniform(-5, 5, n), rng.uniform(0, 1, n)])
    y = np.concatenate([wide(n), rng.uniform(-5, 5, n), rng.uniform(0.5, 4.2, n)])
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                   1e-310, 3.0, 1.0 / 3.0, 2.0 ** -1022, 2.0 ** 1023, 2.0 ** -537, 2.0 ** 600])
    xs, ys = np.meshgrid(sp, sp)
    x = np.concatenate([x, xs.ravel()]); y = np.concatenate([y, ys.ravel()])
    q_spec, q_std, s_spec, s_std, fast = api.selftest_spec(x, y)
    for a, b, what in ((q_spec, q_std, "x / y"), (s_spec, s_std, "sqrt(x)")):
        nan = np.isnan(b)
        assert np.array_equal(np.isnan(a), nan), what
        bad = np.flatnonzero((a.view(np.uint64) != b.view(np.uint64)) & ~nan)
        assert bad.size == 0, (what, bad[:5], x[bad[:5]], y[bad[:5]], a[bad[:5]], b[bad[:5]])
    with np.errstate(all="ignore"):                     # and the device's standard operators are IEEE: numpy's on the host agree
        assert np.array_equal(q_std, x / y, equal_nan=True) and np.array_equal(s_std, np.sqrt(x), equal_nan=True)
    mid = slice(n, 3 * n)                               # ordinary operands: the fast sequences carry (almost) all of them
    assert np.mean((fast[mid] & 1) != 0) > 0.99 and np.mean((fast[2 * n:3 * n] & 2) != 0) > 0.99
    assert not np.any(fast[3 * n:][np.isnan(x[3 * n:]) | np.isnan(y[3 * n:])] & 1)      # a NaN operand is never accepted
