"""CPU: the C-ABI library loads, exports every symbol include/pysco_b200.h declares, the ctypes
table matches the header, and argument validation fails loudly -- no compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "pysco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(psc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from pysco_b200 import build
    build.build()
    from pysco_b200 import _lib
    return _lib


def test_header_symbols_exported(lib):
    names = header_functions()
    assert len(names) >= 35
    raw = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/pysco_b200.h but not exported"
    assert sorted(lib.SIGNATURES) == names, "ctypes signature table out of sync with the header"
    assert lib.load().psc_version() >= 100


def test_argument_validation_without_gpu(lib):
    L = lib.load()
    # invalid arguments are rejected before any CUDA call
    assert L.psc_gradient(None, None, 0.0, 0, 4, 0, 16, None, 3, None) == -1
    assert b"order" in L.psc_last_error()
    assert L.psc_deposit(None, 10, 16, 7, 1.0, 1.0, 0.0, None, None) == -1
    assert L.psc_green(None, 16, 0, 0, 1.0, None) == -1
    assert L.psc_mond_rhs(None, None, 16, 1.0, 9, 1.0, None) == -1
    with pytest.raises(ValueError):
        lib.check(-1)
    with pytest.raises(lib.PyscoCudaError):
        lib.check(-2)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pysco_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pysco_b200.mesh.TSC(np.zeros((4, 3), np.float32), 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pysco_b200.laplacian.operator(np.zeros((8, 8, 8), np.float32))


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or /root/reference)."""
    pkg = os.path.join(ROOT, "pysco_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "/root/reference" not in txt, f
                assert "pysco_oracle" not in txt, f


def test_host_side_helpers():
    from pysco_b200 import cosmology, iostream, utils
    c = cosmology.Flatw0waCDM(H0=72, Om0=0.25733, Tcmb0=2.726, Neff=3.044, w0=-1.0, wa=0.0)
    assert abs(c.Ogamma0 + c.Onu0 - 8.0763e-05) < 5e-8          # reference examples/INFOS: 0.000080763
    assert abs(c.efunc(0.0) - 1.0) < 1e-12
    assert abs(c.Om(0.0) + c.Ogamma(0.0) + c.Onu(0.0) + c.Ode(0.0) - 1.0) < 1e-12
    p = {"H0": 72, "aexp": 0.5, "boxlen": 100, "Om_m": 0.25733, "npart": 128 ** 3}
    utils.set_units(p)
    assert abs(p["unit_t"] / 1.0714158269067246e17 - 1) < 1e-12   # value printed by the reference
    assert abs(p["unit_l"] / 2.1428316538134494e21 - 1) < 1e-12
    path = os.path.join(ROOT, "examples", "param.ini")
    s = iostream.read_param_file(path)
    assert s["npart"] == 128 ** 3 and s["theory"] == "newton" and s["fixed_ICS"] is False
    assert s["z_out"] == "[10, 5, 2, 1, 0.5, 0]" and s["Courant_factor"] == 1.0
    assert iostream.parse_z_out(s) == [10, 5, 2, 1, 0.5, 0]


def test_cosmotable_tables():
    import pandas as pd
    from pysco_b200 import cosmotable
    param = pd.Series({"H0": 72, "Om_m": 0.25733, "T_cmb": 2.726, "N_eff": 3.044, "w0": -1.0, "wa": 0.0,
                       "theory": "newton", "base": ""})
    tabs = cosmotable.generate(param)
    assert len(tabs) == 13
    lna = np.log(0.5)
    t = float(tabs[1](lna))
    assert t < 0 and abs(float(tabs[0](t)) - lna) < 1e-8        # lna(t) inverts t(lna)
    assert abs(float(tabs[1](0.0))) < 1e-12                      # t = 0 today
    d1 = float(tabs[3](0.0)) / float(tabs[3](lna))
    assert 1.5 < d1 < 2.0                                        # LCDM growth between a = 0.5 and 1
    assert 0.4 < float(tabs[4](0.0)) < 0.6                       # f = dlnD/dlna ~ Om^0.55 ~ 0.47
    assert abs(param["Om_lambda"] + param["Om_m"] + param["Om_r"] - 1) < 1e-12


def test_cosmotable_vs_reference_golden():
    """pysco_b200.cosmotable.generate against the reference's cosmotable.generate (cosmotable.py:18-110) at 24 scale
    factors: a(t), t(a), H(a), D1, f1, D2, f2, D3a .. f3c for LCDM, w0-wa and the parametrized theory
    (tests/golden/cosmotable.npz, make_golden.py cosmotable).  Provenance: astropy is not installed where the vectors
    were made, so the reference ran with the repo's Flatw0waCDM restatement as its astropy.cosmology class -- the
    vectors pin the time integration and the growth ODEs, NOT astropy's E(a) (parity unpinned for that class)."""
    import os
    import sys
    import pandas as pd
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    from pysco_b200 import cosmotable
    g = np.load(os.path.join(ROOT, "tests", "golden", "cosmotable.npz"))
    lna = np.log(g["aexp"])
    for name, over in cases.COSMO_CASES.items():
        param = pd.Series(cases.cosmo_param(**over))
        tabs = cosmotable.generate(param)
        ref = g[f"{name}_tables"]
        t = tabs[1](lna)
        mine = np.array([tabs[0](t)] + [tb(lna) for tb in tabs[1:]], dtype=np.float64)
        assert mine.shape == ref.shape == (13, len(lna))
        np.testing.assert_allclose(mine, ref, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose([param["Om_r"], param["Om_lambda"]], g[f"{name}_Om_r_Om_lambda"], rtol=1e-12)


def test_dt_weak_variation_memo_follows_table_and_parameters():
    """integration.dt_weak_variation answers the second of its two calls per step from a one-entry memo: the memo must
    follow the table object, a(t) and max_aexp_stepping (integration.py:329-358 has no state)"""
    import numpy as np
    import pandas as pd
    from scipy.interpolate import interp1d
    from pysco_b200 import integration

    def table(scale):
        lna = np.linspace(-6.0, 0.5, 64)
        return interp1d(lna, scale * np.exp(1.5 * lna), fill_value="extrapolate")

    def direct(f, p):
        fac = 1.0 + 0.01 * p["max_aexp_stepping"]
        return np.float32(f(np.log(fac * p["aexp"])) - f(np.log(p["aexp"])))

    p = pd.Series({"aexp": 0.1, "max_aexp_stepping": 10})
    f1, f2 = table(1.0), table(2.0)
    for f in (f1, f1, f2, f1):
        assert integration.dt_weak_variation(f, p) == direct(f, p)
    p["aexp"] = 0.2
    assert integration.dt_weak_variation(f1, p) == direct(f1, p)
    p["max_aexp_stepping"] = 5
    assert integration.dt_weak_variation(f1, p) == direct(f1, p)
    del f2
    f3 = table(3.0)      # may reuse the identity of the deleted table
    assert integration.dt_weak_variation(f3, p) == direct(f3, p)
