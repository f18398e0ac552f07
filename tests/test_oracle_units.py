"""CPU tests of the oracles: unit/property checks of the literal NumPy restatement and the C
restatement against each other and against independent mathematics."""
import ctypes as C
import itertools

import numpy as np
import pytest

from oracle import literal as L


# ---------------------------------------------------------------- elementary functions
def test_inhouse_math_matches_libm(coracle_mod):
    lib = coracle_mod.lib()
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.exp(rng.uniform(-30, 30, 4000)), [1.0, 0.5, 2.0, 1e-300, 1e300]])
    for x in xs:
        assert abs(lib.mko_test_log(x) - np.log(x)) <= 4e-16 * max(1.0, abs(np.log(x)))
    for x in np.concatenate([rng.uniform(-50, 50, 4000), [0.0, 1e-10, -1e-10, 700.0, -700.0]]):
        assert abs(lib.mko_test_exp(x) - np.exp(x)) <= 6e-16 * np.exp(x)
    for x in np.concatenate([rng.uniform(-5, 5, 3000), rng.uniform(-1e-3, 1e-3, 1000), np.exp(rng.uniform(0, 40, 500)), [0.0, 1e-300]]):
        assert abs(lib.mko_test_asinh(x) - np.arcsinh(x)) <= 6e-16 * max(abs(np.arcsinh(x)), 1e-300)
    assert np.isnan(lib.mko_test_asinh(float("nan"))) and np.isinf(lib.mko_test_asinh(float("inf")))
    assert lib.mko_test_asinh(-0.3) == -lib.mko_test_asinh(0.3)


# ---------------------------------------------------------------- predMat
def test_predmat_equals_explicit_simulation():
    rng = np.random.default_rng(1)
    for _ in range(20):
        a = np.concatenate([rng.uniform(0.2, 0.999, 5), [1.0]])
        A = np.diag(a); B = np.ones((6, 1)); Cr = rng.normal(size=(1, 6)); D = rng.normal(size=(1, 1))
        Np, Nc = int(rng.integers(1, 9)), int(rng.integers(1, 5))
        Phi, G, aug = L.predMat(A, B, Cr, D, Np, Nc)
        Nc = min(Nc, Np)
        x0 = rng.normal(size=7); dU = rng.normal(size=Nc)
        # simulate x_aug(k+1) = Abar x_aug(k) + Bbar du(k), y(k) = Cbar x_aug(k), y stacked for k = 1..Np
        x = x0.copy(); ys = []
        for k in range(Np):
            x = aug.A @ x + aug.B[:, 0] * (dU[k] if k < Nc else 0.0)
            ys.append((aug.C @ x)[0])
        # the reference places H{i-j+1} = Cbar*Abar^(i-j)*Bbar: the input du(j) acts on y(i) through i-j steps
        pred = Phi @ x0 + G @ dU
        x = x0.copy(); ys2 = []
        for k in range(Np):
            ys2.append((aug.C @ (aug.A @ x))[0] + sum((aug.C @ np.linalg.matrix_power(aug.A, k - j) @ aug.B)[0, 0] * dU[j] for j in range(min(k + 1, Nc))))
            x = aug.A @ x
        assert np.allclose(pred, ys2, rtol=1e-10, atol=1e-12)


def test_predmat_soc_closed_form():
    r = -1.0 / (3600 * 29.86)
    a = np.array([0.5, 0.6, 0.7, 0.8, 0.9, 1.0])
    Phi, G, _ = L.predMat(np.diag(a), np.ones((6, 1)), np.array([[0, 0, 0, 0, 0, r]]), np.zeros((1, 1)), 5, 2)
    for i in range(1, 6):
        assert np.array_equal(Phi[i - 1], [0, 0, 0, 0, 0, r, r * i])
        for j in range(1, 3):
            assert G[i - 1, j - 1] == (r * (i - j) if j <= i else 0.0)
    assert np.all(G[0] == 0)          # Q7: first SOC row of M is structurally zero


def test_predmat_c_oracle_matches_literal(coracle_mod):
    lib = coracle_mod.lib()
    rng = np.random.default_rng(2)
    for _ in range(50):
        a5 = rng.uniform(0.1, 0.9999, 5); c6 = rng.normal(size=6); d = float(rng.normal()); x6 = rng.normal(size=6)
        Np = int(rng.integers(1, 21))
        h = np.empty(Np + 1); phix = np.empty(Np)
        dp = lambda v: v.ctypes.data_as(C.POINTER(C.c_double))
        lib.mko_pred_row(dp(a5), dp(c6), C.c_double(d), dp(x6), Np, dp(h), dp(phix))
        Phi, G, _ = L.predMat(np.diag(np.concatenate([a5, [1.0]])), np.ones((6, 1)), c6[None, :], np.array([[d]]), Np, min(3, Np))
        assert np.allclose(h[1:], Phi[:, 6], rtol=1e-12, atol=1e-14)
        assert np.allclose(phix, Phi[:, :6] @ x6, rtol=1e-11, atol=1e-13)
        assert np.allclose(G[:, 0], h[:Np], rtol=1e-12, atol=1e-14)


# ---------------------------------------------------------------- Hildreth
def _random_qp(rng, Nc, nC, scale_rows=True):
    Aq = rng.normal(size=(Nc, Nc)); E = Aq @ Aq.T + Nc * np.eye(Nc)
    F = rng.normal(size=Nc) * 3
    M = rng.normal(size=(nC, Nc))
    if scale_rows:
        M *= np.exp(rng.uniform(-3, 1, size=(nC, 1)))
    gamma = np.abs(rng.normal(size=nC)) * 0.5 + M @ (-np.linalg.solve(E, F)) * rng.uniform(0.2, 0.9)
    return E, F, M, gamma


def test_hildreth_c_rank_form_matches_literal_dense(bank, coracle_mod):
    co = coracle_mod.COracle(bank)
    rng = np.random.default_rng(3)
    n_conv = 0
    for t in range(300):
        Nc = int(rng.integers(1, 6)); nC = int(rng.integers(1, 40))
        E, F, M, g = _random_qp(rng, Nc, nC)
        lam0 = None if t % 3 else np.abs(rng.normal(size=nC))
        DU1, l1, n1 = L.hildreth(E, F, M, g, lam0, 100)
        DU2, l2, n2 = co.hildreth(E, F, M, g, lam0, 100)
        assert n1 == n2
        assert np.allclose(DU1, DU2, rtol=1e-9, atol=1e-11)
        assert np.allclose(l1, l2, rtol=1e-8, atol=1e-10)
        assert np.array_equal(l1 > 0, l2 > 0)
        n_conv += n1 < 100
    assert n_conv > 100


def test_hildreth_kkt_and_bruteforce_active_set(bank, coracle_mod):
    co = coracle_mod.COracle(bank)
    rng = np.random.default_rng(4)
    checked = 0
    for _ in range(60):
        Nc, nC = 2, int(rng.integers(2, 8))
        E, F, M, g = _random_qp(rng, Nc, nC, scale_rows=False)
        DU, lam, nexec = co.hildreth(E, F, M, g, None, 5000)
        if nexec >= 5000:
            continue
        # KKT: stationarity is exact by construction; primal feasibility, dual feasibility, complementarity
        assert np.all(lam >= 0)
        assert np.all(M @ DU - g <= 1e-5)
        assert np.all(np.abs(lam * (M @ DU - g)) <= 1e-5 * (1 + np.abs(lam)))
        # brute force over active sets: the QP optimum is the feasible KKT point of smallest cost
        best = None
        for k in range(0, Nc + 1):
            for S in itertools.combinations(range(nC), k):
                S = list(S)
                if S:
                    Ms = M[S]
                    KKT = np.block([[E, Ms.T], [Ms, np.zeros((len(S), len(S)))]])
                    try:
                        sol = np.linalg.solve(KKT, np.concatenate([-F, g[S]]))
                    except np.linalg.LinAlgError:
                        continue
                    x, mu = sol[:Nc], sol[Nc:]
                    if np.any(mu < -1e-9):
                        continue
                else:
                    x = -np.linalg.solve(E, F)
                if np.all(M @ x - g <= 1e-9):
                    J = 0.5 * x @ E @ x + F @ x
                    if best is None or J < best[0]:
                        best = (J, x)
        assert best is not None
        assert np.allclose(DU, best[1], rtol=1e-4, atol=1e-5)
        checked += 1
    assert checked > 30


def test_hildreth_zero_row_inf_nan_semantics(bank, coracle_mod):
    """Q7: an all-zero M row gives H_ii = 0 => w = -K/0; with gamma_i < 0 lambda_i flips Inf <-> 0 every
    sweep, poisons every other row with NaN (-> 0) and the cap is hit; both restatements must agree."""
    co = coracle_mod.COracle(bank)
    rng = np.random.default_rng(5)
    for trial in range(60):
        Nc = int(rng.integers(1, 4)); nC = int(rng.integers(3, 16))
        E, F, M, g = _random_qp(rng, Nc, nC)
        zr = rng.choice(nC, size=int(rng.integers(1, 3)), replace=False)
        M[zr] = 0.0
        g[zr] = rng.choice([-0.3, 0.0, 0.4], size=len(zr))
        for maxIter in (100, 7):
            DU1, l1, n1 = L.hildreth(E, F, M, g, None, maxIter)
            DU2, l2, n2 = co.hildreth(E, F, M, g, None, maxIter)
            assert n1 == n2, (trial, maxIter)
            assert np.array_equal(np.isnan(DU1), np.isnan(DU2)) and np.array_equal(np.isinf(l1), np.isinf(l2))
            fin = np.isfinite(DU1)
            assert np.allclose(DU1[fin], DU2[fin], rtol=1e-9, atol=1e-11)
            lf = np.isfinite(l1)
            assert np.allclose(l1[lf], l2[lf], rtol=1e-8, atol=1e-10)
            if np.any(g[zr] < 0):
                assert n1 == maxIter                      # never converges
    # the runMPC situation: SOC already above z_max -> first SOC row has gamma < 0, even cap => finite DU
    E = np.array([[3e-4, 1e-10], [1e-10, 3e-4]]); F = np.array([0.0177, 0.0106])
    r = -9.3e-6
    M = np.array([[1, 0], [0, 0], [r, 0], [2 * r, r]]); g = np.array([50.0, -1e-4, -1e-4 + r, -1e-4 + 2 * r])
    DU1, l1, n1 = L.hildreth(E, F, M, g, None, 100)
    DU2, l2, n2 = co.hildreth(E, F, M, g, None, 100)
    assert n1 == n2 == 100 and np.all(np.isfinite(DU1)) and np.allclose(DU1, DU2, rtol=1e-9)


# ---------------------------------------------------------------- index choice
def test_two_nearest_matches_stable_sort(coracle_mod):
    lib = coracle_mod.lib()
    rng = np.random.default_rng(6)
    grids = [np.arange(0.05, 1.0, 0.1), np.array([288.15, 298.15, 308.15]), np.array([0.5]), np.array([0.2, 0.8])]
    for pts in grids:
        vals = list(rng.uniform(pts.min() - 0.3, pts.max() + 0.3, 200)) + list(pts) + list((pts[:-1] + pts[1:]) / 2)
        for v in vals:
            i1, i2 = C.c_int(), C.c_int()
            lib.mko_test_two_nearest(C.c_double(v), pts.ctypes.data_as(C.POINTER(C.c_double)), len(pts), C.byref(i1), C.byref(i2))
            order = L._msort(np.abs(v - pts))
            assert i1.value == order[0]
            assert i2.value == (order[1] if len(pts) > 1 else order[0])


def test_getxind_weights_sum_to_one_and_extrapolate(bank):
    ROM = L.build_rom(bank)
    kf = L.initKF(10, 25, np.diag([1.0] * 5 + [2e6]), 1e-3, 1e2, "OutB", ROM)
    zk, zb, kf, X = L.iterEKF(3.5, 0.0, 25, kf)
    assert abs(X.gamma.sum() - 1) < 1e-12
    # Q2/Q15: T exactly on a grid point -> stable sort picks the lower neighbour, alphaT = 1
    assert list(X.theT) == [0, 0, 1, 1] and X.gamma[0] == 0 and X.gamma[1] == 0


# ---------------------------------------------------------------- PSD repair
def test_psd_repair_fast_path_and_projection(coracle_mod):
    lib = coracle_mod.lib()
    rng = np.random.default_rng(7)
    iu = np.triu_indices(5)

    def run(S, force):
        v = np.ascontiguousarray(S[iu])
        ran = lib.mko_test_psd(v.ctypes.data_as(C.POINTER(C.c_double)), int(force))
        out = np.zeros((5, 5)); out[iu] = v
        return out + np.triu(out, 1).T, ran

    for _ in range(200):
        Aq = rng.normal(size=(5, 5)) * np.exp(rng.uniform(-2, 3, size=(5, 1)))
        S = Aq @ Aq.T
        if rng.random() < 0.5:                       # make it indefinite
            w, V = np.linalg.eigh(S); w[: int(rng.integers(1, 3))] *= -rng.uniform(0.1, 1.0); S = (V * w) @ V.T; S = (S + S.T) / 2
        _, SS, VVt = np.linalg.svd(S)
        HH = VVt.T @ np.diag(SS) @ VVt
        lit = (S + S.T + HH + HH.T) / 4              # iterEKF.m:143-145
        fast, ran = run(S, False)
        full, _ = run(S, True)
        scale = np.abs(S).max()
        assert np.allclose(fast, lit, rtol=0, atol=2e-12 * scale)
        assert np.allclose(full, lit, rtol=0, atol=2e-12 * scale)
        assert ran == (np.linalg.eigvalsh(S).min() <= 0) or abs(np.linalg.eigvalsh(S).min()) < 1e-10 * scale
        assert np.linalg.eigvalsh(fast).min() >= -1e-10 * scale


# ---------------------------------------------------------------- bank checks
def test_bank_structural_checks(bank):
    import copy
    from mek.bank import BankError, resolve_indices
    b = copy.deepcopy(bank); b.A[0, 0, -1] = 0.9
    with pytest.raises(BankError):
        b.validate()
    names = list(bank.names); names[names.index("negPhise")] = "xx"
    with pytest.raises(BankError):
        resolve_indices(names, bank.xloc)
    ind = resolve_indices(bank.names, bank.xloc)
    assert ind["Thetass0"] == [9] and ind["Thetass3"] == [22] and ind["Phie"] == [23, 24, 25] and ind["negPhise"][1] == 8
    # the literal restatement's own set-up code resolves the same rows
    from types import SimpleNamespace as NS
    ind2, _, _ = L._ekf_setupIndsLocs(NS(tfData=NS(names=list(bank.names), xLoc=bank.xloc)))
    for k in ("Ifdl0", "Ifdl3", "If0", "If3", "Thetass0", "Thetass3", "Phise0", "Phie", "Thetae"):
        assert ind[k] == ind2[k]


def test_initkf_rejects_bad_sigma(bank):
    ROM = L.build_rom(bank)
    with pytest.raises(ValueError):
        L.initKF(10, 25, np.eye(5), 1e-3, 1e2, "OutB", ROM)      # initKF.m:66-69
    with pytest.raises(ValueError):
        L.OB_step(0.0, 25, None, ROM, {})                        # OB_step.m:40-42


def test_quirks_q3_q9_q20(bank):
    """Q3: MPC integrator state is 0; Q9: Hildreth runs on essentially every step; Q20: init call is a state no-op."""
    o = L.run_mpc(bank, N=60)
    assert np.all(o["x"][5, :] == 0.0)
    assert (o["nexec"] > 0).all()
    ROM = L.build_rom(bank)
    V, obs, cs = L.OB_step(0.0, 25, None, ROM, dict(SOC0=10, warnOff=True))
    assert np.all(cs.bigX == 0) and cs.SOCnAvg == cs.SOC0n
    # u ramps at du_min per sample down to u_min = -Q*Crate (runMPC.m:36-38, initMPC.m:66-67)
    assert abs(o["u"][5] + 2 * 29.86) < 1e-3 and abs(o["u"][0] + 50) > 1.0


def test_portable_and_fma_builds_of_the_c_oracle_agree_bitwise(bank, coracle_mod):
    """oracle/Makefile builds mek_oracle.c twice (libm fma vs -mfma); fma is correctly rounded either way."""
    import ctypes as C
    import os
    from mek import population as P
    d = os.path.join(os.path.dirname(coracle_mod.__file__), "_build")
    coracle_mod.build()
    if not coracle_mod._has_fma():
        pytest.skip("no hardware FMA on this host")
    pop = P.make_population(bank.params[2], 3, seed=12, heterogeneous=True)
    outs = []
    for so in ("libmek_oracle.so", "libmek_oracle_fma.so"):
        co = coracle_mod.COracle(bank)
        co.L = C.CDLL(os.path.join(d, so))
        co.L.mko_state_len.restype = C.c_int
        co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
        outs.append(co.run(0, 900))
    for k in ("u", "V", "SOC", "phise", "nexec", "iT", "iZ", "active"):
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_zero_weight_rule_equals_the_term_by_term_restatement(bank, coracle_mod):
    """The zero-weight rule of the C restatement and the CUDA kernels (a neighbour model with bilinear weight exactly 0 is left out
    of the blends and of the measurement update) against the same oracle with the rule switched off, i.e.
    the reference's term-by-term arithmetic: identical logs, bit for bit, on cells at 25 degC (on a set-point of the bank: two
    zero-weight neighbours at every instant) and on cells at temperatures between set-points (no zero weight: same code path)."""
    from mek import population as P
    L = coracle_mod.lib()
    for het, seed in ((False, 11), (True, 12)):
        pop = P.make_population(bank.params[2], 24, seed=seed, heterogeneous=het)
        runs = []
        for on in (1, 0):
            L.mko_set_zero_weight_skip(on)
            try:
                co = coracle_mod.COracle(bank)
                co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
                runs.append(co.run(0, 3001, nthreads=8))
            finally:
                L.mko_set_zero_weight_skip(1)
        a, b = runs
        for k in ("u", "V", "SOC", "phise", "nexec", "iT", "iZ", "active"):
            assert np.array_equal(a[k], b[k], equal_nan=True) if a[k].dtype.kind == "f" else np.array_equal(a[k], b[k]), (het, k)
