"""Mirror of pysco/laplacian.py (7-point Laplacian multigrid kernels)."""
import numpy as np
import torch

from . import _lib

K = _lib.OP_LAPLACIAN


def _sumsq(fn_call):
    out = _lib.zeros((1,), torch.float64)
    fn_call(out)
    return np.float32(np.sqrt(out.item()))


def operator(x, _kind=K, b=None, q=0.0):
    """laplacian.py:12-54"""
    c = _lib.Ctx()
    tx, tb = c.dev(x), c.dev(b)
    out = torch.empty_like(tx)
    _lib.check(_lib.load().psc_operator(_lib.ptr(tx), _lib.ptr(tb), float(np.float32(q)), tx.shape[0], _kind,
                                        _lib.ptr(out), _lib.stream()))
    return c.ret(out)


def residual(x, b):
    """laplacian.py:63-117: b - Lx"""
    c = _lib.Ctx()
    tx, tb = c.dev(x), c.dev(b)
    out = torch.empty_like(tx)
    _lib.check(_lib.load().psc_residual(_lib.ptr(tx), _lib.ptr(tb), 0.0, None, tx.shape[0], K, _lib.ptr(out),
                                        _lib.stream()))
    return c.ret(out)


def restrict_residual(x, b):
    """laplacian.py:125-226: R(b - Lx), fused"""
    c = _lib.Ctx()
    tx, tb = c.dev(x), c.dev(b)
    N = tx.shape[0]
    out = _lib.empty((N // 2,) * 3)
    _lib.check(_lib.load().psc_restrict_residual(_lib.ptr(tx), _lib.ptr(tb), N, _lib.ptr(out), _lib.stream()))
    return c.ret(out)


def residual_error(x, b, _kind=K, q=0.0):
    """laplacian.py:327-381: sqrt(sum (b - Lx)^2) (synchronises)"""
    c = _lib.Ctx()
    tx, tb = c.dev(x), c.dev(b)
    return _sumsq(lambda out: _lib.check(_lib.load().psc_residual_sumsq(
        _lib.ptr(tx), _lib.ptr(tb), float(np.float32(q)), tx.shape[0], _kind, _lib.ptr(out), _lib.stream())))


def truncation_error(x):
    """laplacian.py:502-533: || R(Lx) - L(Rx) ||"""
    c = _lib.Ctx()
    tx = c.dev(x)
    from . import mesh
    RLx = mesh.restriction(operator(tx))
    LRx = operator(mesh.restriction(tx))
    return _sumsq(lambda out: _lib.check(_lib.load().psc_diff_sumsq(
        _lib.ptr(RLx), 1.0, _lib.ptr(LRx), RLx.numel(), _lib.ptr(out), _lib.stream())))


def initialise_potential(b, _kind=K, q=0.0):
    """laplacian.py:765-796: -h^2/6 * b"""
    c = _lib.Ctx()
    tb = c.dev(b)
    out = torch.empty_like(tb)
    _lib.check(_lib.load().psc_initialise_potential(_lib.ptr(tb), float(np.float32(q)), tb.shape[0], _kind,
                                                    _lib.ptr(out), _lib.stream()))
    return c.ret(out)


def gauss_seidel(x, b, f_relax, _kind=K, q=0.0, rhs=None) -> None:
    """laplacian.py:844-1022: one red-black SOR sweep, in place"""
    c = _lib.Ctx()
    tx, tb, tr = c.dev(x, inplace=True), c.dev(b), c.dev(rhs)
    _lib.check(_lib.load().psc_gauss_seidel(_lib.ptr(tx), _lib.ptr(tb), float(np.float32(q)), _lib.ptr(tr),
                                            tx.shape[0], _kind, float(f_relax), _lib.stream()))
    c.finish()


def fused_sweeps_enabled():
    import os
    return not os.environ.get("PSC_NO_FUSED_GS")


def sweeps(tx, tb, n, _kind=K, q=0.0, tr=None, f_relax=np.float32(1.25)) -> None:
    """n red-black SOR sweeps on device tensors, in place on tx.  Laplacian sweeps on grids of 512^3 and more run in
    PAIRS through psc_gauss_seidel_fused (x -> scratch -> x: 13 B per cell and sweep); an odd last sweep, every sweep
    of a smaller grid and the f(R) smoothers are the two-launch in-place psc_gauss_seidel (24 B per cell).  Same bits
    either way."""
    lib = _lib.load()
    N, n = tx.shape[0], int(n)
    args = (float(np.float32(q)), _lib.ptr(tr), N, _kind, float(f_relax))
    # measured (tools/bench_multigrid.py, profiles/r02_multigrid_{512,256}.txt): the fused sweep wins at 512^3 for the
    # Laplacian (0.693 vs 0.722 ms) but loses at 256^3 (32 column CTAs for 148 SMs: 0.218 vs 0.103 ms) and for the
    # f(R) smoothers, whose arithmetic hides the second pass over the data (1.51 vs 1.29 ms)
    if n >= 2 and _kind == 0 and N >= 512 and fused_sweeps_enabled() and lib.psc_gauss_seidel_fused_supported(N):
        tmp = torch.empty_like(tx)
        tma = 0 if __import__("os").environ.get("PSC_GS_NO_TMA") else 1
        for _ in range(n // 2):
            _lib.check(lib.psc_gauss_seidel_fused(_lib.ptr(tx), _lib.ptr(tb), *args, _lib.ptr(tmp), tma, _lib.stream()))
            _lib.check(lib.psc_gauss_seidel_fused(_lib.ptr(tmp), _lib.ptr(tb), *args, _lib.ptr(tx), tma, _lib.stream()))
        n -= 2 * (n // 2)
    for _ in range(n):
        _lib.check(lib.psc_gauss_seidel(_lib.ptr(tx), _lib.ptr(tb), *args, _lib.stream()))


def smoothing(x, b, n_smoothing) -> None:
    """laplacian.py:1026-1055"""
    c = _lib.Ctx()
    tx, tb = c.dev(x, inplace=True), c.dev(b)
    sweeps(tx, tb, n_smoothing)
    c.finish()
