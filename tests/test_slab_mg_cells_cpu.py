"""The per-cell bodies of the slab multigrid / MOND kernels (pysco_b200/csrc/slab_mg_cells.cuh), run on the host through
tests/slab_mg_harness.cpp, against the oracle's single-domain functions -- and inside guard bands: every array sits
between two bands of NaN, so a read outside the array poisons the result and a write outside it shows in the band."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle  # noqa: E402
from slab_oracle_ops import OracleOps  # noqa: E402

N = 16
GUARD = 3 * N * N


def guarded(a):
    """a copy of `a` placed between two NaN bands; returns (view, whole buffer)"""
    a = np.ascontiguousarray(a, dtype=np.float32)
    buf = torch.full((a.size + 2 * GUARD,), float("nan"), dtype=torch.float32)
    view = buf[GUARD:GUARD + a.size].view(*a.shape)
    view.copy_(torch.from_numpy(a))
    return view, buf


def bands_intact(buf, size):
    return bool(torch.isnan(buf[:GUARD]).all() and torch.isnan(buf[GUARD + size:]).all())


def ghosted(x):
    """periodic cube -> [N + 2, N, N] with the wrapped planes as ghosts (a one-rank slab)"""
    return np.concatenate([x[-1:], x, x[:1]])


@pytest.fixture(scope="module")
def fields():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((N, N, N)).astype(np.float32)
    b = rng.standard_normal((N, N, N)).astype(np.float32)
    return x, b


def test_operator_restrictions_prolongation_in_guard_bands(fields):
    x, b = fields
    ops = OracleOps(N, 1, 0)
    xg, xbuf = guarded(ghosted(x))
    bg, bbuf = guarded(b)
    out = ops.mg_operator(xg, N, N)
    assert np.allclose(out.numpy(), oracle.laplacian.operator(x), rtol=0, atol=1e-3)
    rr = ops.mg_restrict_residual(xg, bg, N, N)
    assert np.allclose(rr.numpy(), oracle.laplacian.restrict_residual(x, b), rtol=0, atol=1e-3)
    r = ops.mg_restriction(bg, N, N, -1.0)
    assert np.allclose(r.numpy(), oracle.mesh.minus_restriction(b), rtol=0, atol=1e-6)
    # prolongation: the fine array is written, bands included in the check
    c = np.random.default_rng(4).standard_normal((N // 2,) * 3).astype(np.float32)
    cg, cbuf = guarded(ghosted(c))
    fine, fbuf = guarded(ghosted(x))
    ops.mg_add_prolongation(fine, cg, N // 2, N // 2)
    want = x.copy()
    oracle.mesh.add_prolongation(want, c)
    assert np.allclose(fine[1:N + 1].numpy(), want, rtol=0, atol=1e-5)
    assert np.array_equal(fine[0].numpy(), x[-1]) and np.array_equal(fine[N + 1].numpy(), x[0])   # ghosts untouched
    for buf, size in ((xbuf, xg.numel()), (bbuf, bg.numel()), (cbuf, cg.numel()), (fbuf, fine.numel())):
        assert bands_intact(buf, size)


def test_red_black_sweep_in_guard_bands(fields):
    x, b = fields
    ops = OracleOps(N, 1, 0)
    xg, xbuf = guarded(ghosted(x))
    bg, bbuf = guarded(b)
    for colour in (1, 0):
        xg[0].copy_(xg[N])
        xg[N + 1].copy_(xg[1])
        ops.mg_gs_colour(xg, bg, N, N, 0, colour, np.float32(1.25))
    want = x.copy()
    oracle.laplacian.gauss_seidel(want, b, np.float32(1.25))
    assert np.allclose(xg[1:N + 1].numpy(), want, rtol=0, atol=1e-5)
    assert bands_intact(xbuf, xg.numel()) and bands_intact(bbuf, bg.numel())


@pytest.mark.parametrize("fn,alpha", [("simple", 1.0), ("n", 2.0), ("beta", 1.5), ("gamma", 1.5), ("delta", 2.0)])
def test_mond_rhs_in_guard_bands(fields, fn, alpha):
    from pysco_b200 import _lib
    x, _ = fields
    phi = (0.05 * x).astype(np.float32)
    ops = OracleOps(N, 1, 0)
    pg, pbuf = guarded(ghosted(phi))
    out, obuf = guarded(np.zeros((N, N, N), np.float32))
    g0 = 0.7
    ops.mond_rhs(pg, out, N, N, g0, _lib.MOND_FN[fn], alpha)
    want = np.empty_like(phi)
    kw = {} if fn == "simple" else {fn: alpha}
    getattr(oracle.mond, f"rhs_{fn}")(phi, want, np.float32(g0), **kw)
    assert np.isfinite(out.numpy()).all()
    scale = np.sqrt(np.mean(want.astype(np.float64) ** 2))
    assert np.abs(out.numpy() - want).max() < 2e-5 * scale
    assert bands_intact(pbuf, pg.numel()) and bands_intact(obuf, out.numel())


@pytest.mark.parametrize("kind,mod", [(1, "cubic"), (2, "quartic")])
def test_scalaron_cells_in_guard_bands(fields, kind, mod):
    """cubic / quartic operator, first guess and one red-black sweep (with and without the FAS right-hand side)"""
    x, b = fields
    m = getattr(oracle, mod)
    q = np.float32(-0.8)
    u = (1.0 + 0.2 * x).astype(np.float32)           # positive scalaron, O(1)
    dens = (0.5 * b).astype(np.float32)
    rhs = (1e-4 * b[::-1]).astype(np.float32).copy()
    ops = OracleOps(N, 1, 0)
    ug, ubuf = guarded(ghosted(u))
    dg, dbuf = guarded(dens)
    rg, rbuf = guarded(rhs)
    L = ops.mg_operator_fr(ug, dg, q, N, N, kind)
    want = m.operator(u, dens, q)
    assert np.abs(L.numpy() - want).max() < 2e-5 * np.sqrt(np.mean(want.astype(np.float64) ** 2))
    first, fbuf = guarded(np.zeros_like(u))
    pos_dens = (np.abs(dens) + 0.1).astype(np.float32)   # the first guess is defined for the physical sign only
    pg, pbuf = guarded(pos_dens)
    ops.mg_init_fr(pg, q, N, N, kind, first)
    want = m.initialise_potential(pos_dens, q)
    assert np.isfinite(want).all() and np.allclose(first.numpy(), want, rtol=2e-6, atol=0)
    for with_rhs in (False, True):
        ug, ubuf = guarded(ghosted(u))
        for colour in (1, 0):
            ug[0].copy_(ug[N])
            ug[N + 1].copy_(ug[1])
            ops.mg_gs_colour_fr(ug, dg, rg if with_rhs else None, q, N, N, 0, colour, np.float32(1.25), kind)
        want = u.copy()
        if with_rhs:
            m.gauss_seidel_with_rhs(want, dens, q, rhs, np.float32(1.25))
        else:
            m.gauss_seidel(want, dens, q, np.float32(1.25))
        assert np.isfinite(want).all()
        assert np.abs(ug[1:N + 1].numpy() - want).max() < 5e-6 * np.abs(want).max()
        assert bands_intact(ubuf, ug.numel())
    for buf, size in ((dbuf, dg.numel()), (rbuf, rg.numel()), (fbuf, first.numel()), (pbuf, pg.numel())):
        assert bands_intact(buf, size)
