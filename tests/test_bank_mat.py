"""Real-ROM import path: the numeric part of a MATLAB ROM file round-trips through mek/bank.py, and
the reference's structural errors (initKF.m:74-88, OB_step.m:140-158) are raised on malformed files."""
import numpy as np
import pytest


def test_mat_round_trip_is_exact(bank, tmp_path):
    from mek import bank as Bk
    p = str(tmp_path / "ROM_synth.mat")
    Bk.to_mat(bank, p)
    b2 = Bk.from_mat(p, params=bank.params)
    assert (b2.n, b2.nz, b2.nT, b2.nZ, b2.Tsamp) == (bank.n, bank.nz, bank.nT, bank.nZ, bank.Tsamp)
    assert list(b2.names) == list(bank.names)
    for k in ("Tpts_C", "Zpts_pct", "xloc", "A", "C", "D", "T_K", "Z", "params"):
        assert np.array_equal(getattr(b2, k), getattr(bank, k)), k
    # without an explicit parameter vector the closed forms fall back to the defaults, F and R come from the file
    b3 = Bk.from_mat(p)
    assert b3.params[Bk.P_F] == bank.params[Bk.P_F] and b3.params[Bk.P_R] == bank.params[Bk.P_R]
    # and the imported bank serialises to the same MEKBANK1 bytes as the original
    q1, q2 = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    bank.save(q1); b2.save(q2)
    assert open(q1, "rb").read() == open(q2, "rb").read()


def test_mat_structural_errors(bank, tmp_path):
    import copy
    from mek import bank as Bk
    bad = copy.deepcopy(bank)
    bad.A[1, 2, -1] = 0.5                                      # initKF.m:74-77
    p = str(tmp_path / "bad1.mat")
    Bk.to_mat(bad, p)
    with pytest.raises(Bk.BankError, match="integrator"):
        Bk.from_mat(p)
    bad = copy.deepcopy(bank)
    bad.names[0] = "negIdl"                                    # drops negIfdl at x = 0: OB_step.m:140-142
    p = str(tmp_path / "bad2.mat")
    Bk.to_mat(bad, p)
    with pytest.raises(Bk.BankError, match="ifdl at negative collector"):
        Bk.from_mat(p)
    import scipy.io
    scipy.io.savemat(str(tmp_path / "empty.mat"), {"x": 1.0})
    with pytest.raises(Bk.BankError, match="not found"):
        Bk.from_mat(str(tmp_path / "empty.mat"))


def test_bank_with_tables_round_trip(tmp_path):
    """Tabulated cellData.function.* (mek_bank_set_lut): file format version 2 carries the tables; version-1 files stay valid."""
    from mek import bank as B
    bk = B.Bank.load(B.default_bank_path())
    assert bk.lut is None
    bt = bk.with_tables(n=501, lo=0.0, hi=1.0).validate()
    assert bt.lut.shape == (len(B.LUT_NAMES), 501)
    f = str(tmp_path / "t.bin")
    bt.save(f)
    b2 = B.Bank.load(f)
    assert np.array_equal(b2.lut, bt.lut) and b2.lut_range == (0.0, 1.0) and np.array_equal(b2.C, bk.C)
    bk.save(f)                                                       # no tables: still a version-1 file
    assert B.Bank.load(f).lut is None
    import struct
    assert struct.unpack_from("<i", open(f, "rb").read(), 8)[0] == 1
    bad = bk.with_tables(n=11)
    bad.lut[2, 3] = np.nan
    with pytest.raises(B.BankError):
        bad.validate()


def test_from_mat_with_sampled_cell_functions(bank, tmp_path):
    """from_mat(..., lut_mat=...): the file matlab/mek_sample_celldata.m writes (theta, tables) becomes Bank.lut."""
    import scipy.io
    from mek import bank as Bk
    rom = str(tmp_path / "ROM.mat"); lutf = str(tmp_path / "lut.mat")
    Bk.to_mat(bank, rom)
    th = np.linspace(0.0, 1.0, 301)
    scipy.io.savemat(lutf, {"theta": th.reshape(1, -1), "tables": Bk.sample_closed_forms(bank.params, th)})
    b = Bk.from_mat(rom, params=bank.params, lut_mat=lutf)
    assert b.lut.shape == (8, 301) and b.lut_range == (0.0, 1.0)
    assert np.array_equal(b.lut, bank.with_tables(n=301).lut)
    scipy.io.savemat(lutf, {"theta": (th ** 2).reshape(1, -1), "tables": Bk.sample_closed_forms(bank.params, th)})
    with pytest.raises(Bk.BankError, match="uniform"):
        Bk.from_mat(rom, params=bank.params, lut_mat=lutf)
