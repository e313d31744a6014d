"""Initial conditions (pysco_b200/initial_conditions.py) against the unmodified reference's output
(tests/golden/ics.npz, made by tests/golden/make_golden.py ics): the host-side white noise / density spectrum on
CPU, the whole 1LPT / 2LPT / 3LPT generation (incl. fixed + paired, dealiased, edge lattice) on the GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
import cases  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "ics.npz"))


@pytest.mark.parametrize("N,seed", [(8, 5), (12, 6)])
def test_white_noise_matches_reference(N, seed):
    from pysco_b200 import initial_conditions as ic
    wn = ic.white_noise_fourier(N, np.random.default_rng(seed))
    assert wn.shape == (N, N, N // 2 + 1) and wn.dtype == np.complex64
    assert np.max(np.abs(wn - G[f"wn_N{N}"])) < 2e-6          # float32 libm differences only
    wf = ic.white_noise_fourier_fixed(N, np.random.default_rng(seed), True)
    assert np.max(np.abs(wf - G[f"wnfixed_N{N}"])) < 2e-6


def test_density_fourier_host_part_matches_reference(tmp_path):
    from pysco_b200 import initial_conditions as ic
    import pandas as pd
    param = pd.Series(cases.ic_param(str(tmp_path), npart=8 ** 3, seed=9))
    transfer = ic.get_transfer_grid(param)
    wn = ic.white_noise_fourier(8, np.random.default_rng(9))
    d = (wn * transfer).astype(np.complex64)
    ref = G["density_fourier_N8"]
    assert np.max(np.abs(d - ref)) < 2e-6 * np.max(np.abs(ref))


def _generate_and_check(name, tmp_path, device):
    import pandas as pd
    from pysco_b200 import initial_conditions as ic, utils
    param = pd.Series(cases.ic_param(str(tmp_path), **cases.IC_CASES[name]))
    param["aexp"] = 1.0 / (1 + param["z_start"])
    utils.set_units(param)
    assert abs(param["unit_t"] - G[f"{name}_unit_t"][0]) <= 1e-12 * abs(param["unit_t"])
    t = G[f"{name}_tables"]   # [H(lna), D1(0), D1(lna), f1, D2, f2, D3a, f3a, D3b, f3b, D3c, f3c]
    tables = [None, None, lambda x: t[0], lambda x: t[1] if x == 0 else t[2]] + \
             [(lambda v: (lambda x: v))(v) for v in t[3:]]
    os.makedirs(os.path.join(str(tmp_path), "output_00000"), exist_ok=True)
    pos, vel = ic.generate(param, tables, device=device)
    pos, vel = pos.cpu().numpy(), vel.cpu().numpy()
    rpos, rvel = G[f"{name}_pos"], G[f"{name}_vel"]
    d = np.abs(pos - rpos)
    d = np.minimum(d, 1 - d)
    assert d.max() < 2e-6, d.max()                                   # box units (cell = 1/16)
    assert np.max(np.abs(vel - rvel)) < 2e-5 * np.sqrt(np.mean(rvel.astype(np.float64) ** 2)) + 1e-9
    assert os.path.exists(os.path.join(str(tmp_path), "output_00000", "particles_test.parquet"))


@pytest.mark.parametrize("name", list(cases.IC_CASES))
def test_generate_matches_reference_cpu(name, tmp_path):
    """the LPT chain is torch ops only: the same code on CPU tensors against the reference's particles"""
    _generate_and_check(name, tmp_path, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(cases.IC_CASES))
def test_generate_matches_reference(name, tmp_path):
    _generate_and_check(name, tmp_path, None)
