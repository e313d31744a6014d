"""Initial conditions generated per x-slab (pysco_b200/initial_conditions.py: white_noise_fourier_block, SlabLayout,
generate_slab) against the unmodified reference's particles (tests/golden/ics.npz): every rank draws only its block
of the white noise -- bit-identical to the slice of the full draw -- and runs the LPT chain on its planes / its
transposed spectrum block; the union of the ranks' particles, put back in lattice order by their ids, is the
reference's particle set.  P = 1, 2, 4 virtual ranks (ThreadComm) on CPU tensors, world_size-2 gloo, and through
slab.run from a parameter file."""
import os
import sys
import threading

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
import cases  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "ics.npz"))


@pytest.mark.parametrize("N,seed", [(8, 5), (12, 6), (32, 1)])
def test_noise_block_is_the_slice_of_the_full_draw(N, seed):
    """random access into the PCG64 stream: same bits as drawing everything (pinned on the reference by
    tests/test_ics.py::test_white_noise_matches_reference)"""
    from pysco_b200 import initial_conditions as ic
    full = ic.white_noise_fourier(N, np.random.default_rng(seed))
    fixed = ic.white_noise_fourier_fixed(N, np.random.default_rng(seed), True)
    unpaired = ic.white_noise_fourier_fixed(N, np.random.default_rng(seed), False)
    for P in (1, 2, 4):
        nyl = N // P
        for r in range(P):
            sl = slice(r * nyl, (r + 1) * nyl)
            assert np.array_equal(ic.white_noise_fourier_block(N, seed, r * nyl, nyl).view(np.float32),
                                  np.ascontiguousarray(full[:, sl]).view(np.float32))
            assert np.array_equal(ic.white_noise_fourier_block(N, seed, r * nyl, nyl, True, True).view(np.float32),
                                  np.ascontiguousarray(fixed[:, sl]).view(np.float32))
            assert np.array_equal(ic.white_noise_fourier_block(N, seed, r * nyl, nyl, True, False).view(np.float32),
                                  np.ascontiguousarray(unpaired[:, sl]).view(np.float32))


def test_transfer_block_is_the_slice(tmp_path):
    import pandas as pd
    from pysco_b200 import initial_conditions as ic
    param = pd.Series(cases.ic_param(str(tmp_path), npart=16 ** 3, seed=9))
    full = ic.get_transfer_grid(param)
    for y0, nyl in ((0, 16), (4, 4), (8, 8)):
        assert np.array_equal(ic.get_transfer_grid_block(param, y0, nyl), full[:, y0:y0 + nyl])


@pytest.mark.parametrize("P", [1, 2, 4])
def test_distributed_transforms_match_rfftn(P):
    from pysco_b200 import initial_conditions as ic
    from pysco_b200.slab import ThreadComm
    N = 16
    x = torch.from_numpy(cases.scalar_grid(N, seed=3))
    spec = torch.fft.rfftn(x, dim=(0, 1, 2))
    out = {}

    def rank(comm):
        L = ic.SlabLayout(comm, N)
        mine = x[L.x0:L.x0 + L.nxl].clone()
        s = L.fft(mine)
        out[comm.rank] = (s, L.ifft(s.clone()), L.y0, L.nyl, L.x0, L.nxl)
    ts = [threading.Thread(target=rank, args=(c,)) for c in ThreadComm.world(P)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    assert len(out) == P
    scale = float(spec.abs().max())
    for r, (s, back, y0, nyl, x0, nxl) in out.items():
        assert float((s - spec[:, y0:y0 + nyl]).abs().max()) < 1e-5 * scale
        assert float((back - x[x0:x0 + nxl]).abs().max()) < 1e-5 * float(x.abs().max())


def _tables_of(name):
    t = G[f"{name}_tables"]   # [H(lna), D1(0), D1(lna), f1, D2, f2, D3a, f3a, D3b, f3b, D3c, f3c]
    return [None, None, lambda x: t[0], lambda x: t[1] if x == 0 else t[2]] + \
           [(lambda v: (lambda x: v))(v) for v in t[3:]]


def _param_of(name, tmp_path):
    import pandas as pd
    from pysco_b200 import utils
    param = pd.Series(cases.ic_param(str(tmp_path), **cases.IC_CASES[name]))
    param["aexp"] = 1.0 / (1 + param["z_start"])
    utils.set_units(param)
    return param


def _check_against_reference(name, parts):
    rpos, rvel = G[f"{name}_pos"], G[f"{name}_vel"]
    ids = np.concatenate([p[2] for p in parts])
    assert np.array_equal(np.sort(ids), np.arange(len(rpos)))
    pos, vel = np.empty_like(rpos), np.empty_like(rvel)
    pos[ids] = np.concatenate([p[0] for p in parts])
    vel[ids] = np.concatenate([p[1] for p in parts])
    d = np.abs(pos - rpos)
    d = np.minimum(d, 1 - d)
    assert d.max() < 2e-6, d.max()                                   # box units (cell = 1/16): test_ics.py's bar
    assert np.max(np.abs(vel - rvel)) < 2e-5 * np.sqrt(np.mean(rvel.astype(np.float64) ** 2)) + 1e-9
    assert pos.min() >= 0 and pos.max() < 1


@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("name", ["lpt1_edge", "lpt2", "lpt3", "lpt2_fixed_paired"])
def test_generate_slab_matches_reference(name, P, tmp_path):
    from pysco_b200 import initial_conditions as ic
    from pysco_b200.slab import ThreadComm
    out, err = {}, []
    param0 = _param_of(name, tmp_path)      # written once (the power-spectrum file), copied per rank

    def rank(comm):
        try:
            param = param0.copy()
            pos, vel, ids = ic.generate_slab(param, _tables_of(name), comm, device="cpu")
            n3 = 16 ** 3 // P
            assert pos.shape == (n3, 3) and ids[0] == comm.rank * n3 and ids[-1] == (comm.rank + 1) * n3 - 1
            out[comm.rank] = (pos.numpy(), vel.numpy(), ids.numpy())
        except Exception as e:   # noqa: BLE001
            err.append(e)
            raise
    ts = [threading.Thread(target=rank, args=(c,)) for c in ThreadComm.world(P)]
    [t.start() for t in ts]
    [t.join(120) for t in ts]
    assert not err and len(out) == P, err
    _check_against_reference(name, [out[r] for r in range(P)])
    # the layout is per thread and gone afterwards: the single-domain generator is untouched
    assert ic._layout() is ic._WHOLE


def test_generate_slab_dealiased_raises(tmp_path):
    from pysco_b200 import initial_conditions as ic
    from pysco_b200.slab import SelfComm
    param = _param_of("lpt3_dealiased", tmp_path)
    with pytest.raises(NotImplementedError):
        ic.generate_slab(param, _tables_of("lpt3_dealiased"), SelfComm(), device="cpu")
    assert ic._layout() is ic._WHOLE
