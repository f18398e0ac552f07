"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/mek.h
declares, bank handling through the ABI (host-only code), loud failure without a GPU, the
population generator and the sharding helpers.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _decl_names():
    txt = open(os.path.join(ROOT, "include", "mek.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mek_[a-z_A-Z0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from mek import _lib
    L = _lib.lib()
    names = _decl_names()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libmek.so does not export " + n
    assert sorted(_lib.SYMBOLS) == names
    assert L.mek_version() >= 100


def test_bank_through_the_abi(bank, tmp_path):
    from mek import _lib, api
    from mek.bank import default_bank_path
    rom = api.RomBank(default_bank_path())
    n, nz, nT, nZ = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _lib.check(_lib.lib().mek_bank_dims(rom.h, C.byref(n), C.byref(nz), C.byref(nT), C.byref(nZ)))
    assert (n.value, nz.value, nT.value, nZ.value) == (5, 30, 3, 10)
    assert rom.Q == bank.params[2]
    rom2 = api.RomBank(bank)                      # from arrays (mek_bank_create)
    assert rom2.Q == rom.Q
    # structural errors surface as MEK_ERR_BANK with the reference's messages
    import copy
    bad = copy.deepcopy(bank); bad.A[1, 2, -1] = 0.5
    with pytest.raises(_lib.MekError, match="integrator"):
        api.RomBank(bad)
    bad = copy.deepcopy(bank); bad.names = ["negIf" if s == "negIfdl" else s for s in bad.names]
    with pytest.raises(_lib.MekError, match="ifdl at negative collector"):
        api.RomBank(bad)
    p = tmp_path / "junk.bin"; p.write_bytes(b"not a bank")
    h = C.c_void_p()
    assert _lib.lib().mek_bank_load(str(p).encode(), C.byref(h)) == -2          # MEK_ERR_BANK
    assert b"MEKBANK1" in _lib.lib().mek_last_error()


def test_no_gpu_fails_loudly(bank):
    """The product has no CPU path: without a device every compute entry point returns MEK_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mek import _lib, api
    rom = api.RomBank(bank)
    with pytest.raises(_lib.MekError, match="no CUDA device"):
        api.CellBatch(rom, 4)
    with pytest.raises(_lib.MekError, match="no CUDA device"):
        api.hildreth_batch(np.eye(2)[None], np.ones((1, 2)), np.ones((1, 3, 2)), np.ones((1, 3)))


def test_population_generator_is_shard_invariant(bank):
    from mek import population as P
    Q = bank.params[2]
    full = P.make_population(Q, 1000, seed=4, heterogeneous=True, ekf_soc_error=2.0)
    part = P.make_population(Q, 300, seed=4, heterogeneous=True, ekf_soc_error=2.0, first_cell=450)
    for f in ("SOC0", "ekf_SOC0", "T", "SigmaX0", "limits"):
        assert np.array_equal(getattr(full, f)[450:750], getattr(part, f))
    assert full.SOC0.min() >= 5 and full.SOC0.max() <= 30 and full.T.min() >= 15 and full.T.max() <= 35
    assert np.all(full.limits[:, 1] <= -Q) and np.all(full.limits[:, 1] >= -3 * Q)
    c1 = P.single_cell(Q)
    assert c1.SOC0[0] == 10 and abs(c1.limits[0, 1] + 59.72) < 1e-9 and c1.SigmaX0[0, 5, 5] == 2e6
    other = P.make_population(Q, 10, seed=5)
    assert not np.array_equal(other.SOC0, P.make_population(Q, 10, seed=4).SOC0)


def test_shard_ranges_cover_population():
    from mek.dist import shard_range
    for B in (1, 7, 4096, 1000003):
        for world in (1, 2, 3, 8):
            rs = [shard_range(B, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == B
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1


def test_mex_gateway_sources_compile_against_stub():
    """No MATLAB here: the MEX gateway is syntax-checked against matlab/mex_stub.h with the real include/mek.h."""
    import subprocess
    d = os.path.join(ROOT, "mpc-ekf4fastcharge_b200", "matlab")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-DMEK_MEX_STUB", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(d, "mek_mex.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for shim in ("runMPC", "initKF", "iterEKF", "iterMPC", "hildreth", "OB_step", "EKFmatsHandler", "predMat", "initMPC"):
        assert os.path.exists(os.path.join(d, shim + ".m"))
