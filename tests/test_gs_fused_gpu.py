"""The plane-marching red-black sweep (csrc/gs_fused.cu: shared-memory ring of four planes, TMA bulk copies or plain
loads) against the two-launch sweep of csrc/multigrid.cu -- which the golden vectors of the reference pin at 16^3 /
32^3 (tests/test_gpu_parity.py) -- and against the oracle: Laplacian, cubic (f(R) n = 1) and quartic (n = 2) smoothers,
with and without the FAS right-hand side.  Same per-cell arithmetic and association: the Laplacian sweep is BIT-IDENTICAL;
the f(R) sweeps agree to a few ulps (the closed-form roots are evaluated in float64 by two separately compiled copies
of the same statements, whose multiply-adds the compiler may contract differently)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def psc():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import pysco_b200
    pysco_b200._lib.load()
    return pysco_b200


def _fields(N, kind, seed=5):
    import torch
    if kind == 0:
        g = torch.Generator(device="cuda").manual_seed(seed)
        b = torch.randn((N, N, N), generator=g, device="cuda")
        x = torch.randn((N, N, N), generator=g, device="cuda") * 1e-3
        return x, b, 0.0, None
    # u ~ 1 with h^2 b << 1: the regime of the golden f(R) kernel cases (every root on a real branch)
    u, b, q, rhs = cases.fr_kernel_case(N, kind)
    return tuple(torch.from_numpy(a).cuda() for a in (u, b)) + (float(q), torch.from_numpy(rhs).cuda())


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("tma", [1, 0])
@pytest.mark.parametrize("N", [128, 192])
def test_fused_sweep_is_bit_identical_to_two_launch_sweep(psc, kind, tma, N):
    import torch
    lib, L = psc._lib, psc._lib.load()
    assert L.psc_gauss_seidel_fused_supported(N) and not L.psc_gauss_seidel_fused_supported(96)
    x, b, q, rhs = _fields(N, kind)
    for r in ((None,) if kind == 0 else (None, rhs)):
        ref = x.clone()
        lib.check(L.psc_gauss_seidel(lib.ptr(ref), lib.ptr(b), float(q), lib.ptr(r), N, kind, 1.25, lib.stream()))
        out = torch.full_like(x, float("nan"))
        lib.check(L.psc_gauss_seidel_fused(lib.ptr(x), lib.ptr(b), float(q), lib.ptr(r), N, kind, 1.25, lib.ptr(out),
                                           tma, lib.stream()))
        torch.cuda.synchronize()
        assert bool(torch.isfinite(out).all())
        if kind == 0:
            assert torch.equal(out, ref), f"tma {tma}: max diff {(out - ref).abs().max().item():.3e}"
        else:
            err = ((out - ref).abs().max() / ref.abs().max()).item()
            assert err < 1e-6, f"kind {kind} tma {tma}: max relative diff {err:.3e}"


def test_fused_sweep_rejects_in_place_and_small_grids(psc):
    import torch
    lib, L = psc._lib, psc._lib.load()
    x = torch.zeros((64, 64, 64), device="cuda")
    with pytest.raises(ValueError):
        lib.check(L.psc_gauss_seidel_fused(lib.ptr(x), lib.ptr(x), 0.0, None, 64, 0, 1.25, lib.ptr(x.clone()), 1,
                                           lib.stream()))
    y = torch.zeros((128, 128, 128), device="cuda")
    with pytest.raises(ValueError):
        lib.check(L.psc_gauss_seidel_fused(lib.ptr(y), lib.ptr(y), 0.0, None, 128, 0, 1.25, lib.ptr(y), 1, lib.stream()))


def test_smoothing_pairs_vs_oracle(psc):
    """laplacian.smoothing / cubic.smoothing with 2 and 3 sweeps at 128^3 (pairs of fused sweeps + an in-place tail)
    against the oracle's sweeps"""
    import oracle
    oracle.build()
    from conftest import assert_close
    N = 128
    x = cases.scalar_grid(N, seed=21, smooth=True)
    b = cases.density_contrast_rhs(N, seed=22)
    for n in (2, 3):
        y = x.copy()
        psc.laplacian.smoothing(y, b, n)
        z = x.copy()
        oracle.laplacian.smoothing(z, b, n)
        assert_close(y, z, 1e-5, f"laplacian smoothing x{n}")
    u, bb, q, rhs = cases.fr_kernel_case(N, 1)
    for n in (2, 3):
        y = u.copy()
        psc.cubic.smoothing(y, bb, q, n)
        z = u.copy()
        oracle.cubic.smoothing(z, bb, q, n)
        assert_close(y, z, 1e-5, f"cubic smoothing x{n}")
        y = u.copy()
        psc.cubic.smoothing_with_rhs(y, bb, q, n, rhs)
        z = u.copy()
        oracle.cubic.smoothing_with_rhs(z, bb, q, n, rhs)
        assert_close(y, z, 1e-5, f"cubic smoothing_with_rhs x{n}")


def test_smoothing_512_fused_pairs_equal_two_launch_sweeps(psc, monkeypatch):
    """laplacian.sweeps at 512^3 (where the product uses the fused pairs) against the two-launch sweeps: same bits"""
    import torch
    N = 512
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((N, N, N), generator=g, device="cuda") * 1e-3
    b = torch.randn((N, N, N), generator=g, device="cuda")
    y = x.clone()
    psc.laplacian.sweeps(y, b, 3)
    monkeypatch.setenv("PSC_NO_FUSED_GS", "1")
    z = x.clone()
    psc.laplacian.sweeps(z, b, 3)
    assert torch.equal(y, z)
