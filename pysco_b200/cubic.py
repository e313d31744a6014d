"""Mirror of pysco/cubic.py (f(R) Hu-Sawicki n = 1: u^3 + p u + q = 0)."""
import numpy as np
import torch

from . import _lib, laplacian, mesh

K = _lib.OP_CUBIC


def operator(x, b, q):
    """cubic.operator"""
    return laplacian.operator(x, K, b, q)


def residual_with_rhs(x, b, q, rhs):
    """cubic.residual_with_rhs: rhs - L(x)"""
    c = _lib.Ctx()
    tx, tb, tr = c.dev(x), c.dev(b), c.dev(rhs)
    out = torch.empty_like(tx)
    _lib.check(_lib.load().psc_residual(_lib.ptr(tx), _lib.ptr(tb), float(np.float32(q)), _lib.ptr(tr),
                                        tx.shape[0], K, _lib.ptr(out), _lib.stream()))
    return c.ret(out)


def initialise_potential(b, q):
    """cubic.initialise_potential"""
    return laplacian.initialise_potential(b, K, q)


def gauss_seidel(x, b, q, f_relax) -> None:
    """cubic.gauss_seidel: one red-black nonlinear SOR sweep, closed-form root in float64"""
    laplacian.gauss_seidel(x, b, f_relax, K, q)


def gauss_seidel_with_rhs(x, b, q, rhs, f_relax) -> None:
    """cubic.gauss_seidel_with_rhs (FAS coarse levels)"""
    laplacian.gauss_seidel(x, b, f_relax, K, q, rhs)


def smoothing(x, b, q, n_smoothing) -> None:
    c = _lib.Ctx()
    tx, tb = c.dev(x, inplace=True), c.dev(b)
    laplacian.sweeps(tx, tb, n_smoothing, K, q)
    c.finish()


def smoothing_with_rhs(x, b, q, n_smoothing, rhs) -> None:
    c = _lib.Ctx()
    tx, tb, tr = c.dev(x, inplace=True), c.dev(b), c.dev(rhs)
    laplacian.sweeps(tx, tb, n_smoothing, K, q, tr)
    c.finish()


def residual_error(x, b, q):
    """cubic.residual_error: sqrt(sum L(x)^2)"""
    return laplacian.residual_error(x, b, K, q)


def truncation_error(x, b, q):
    """cubic.truncation_error: || 4 R(L x) - L(R x; R b) ||"""
    c = _lib.Ctx()
    tx, tb = c.dev(x), c.dev(b)
    RLx = mesh.restriction(operator(tx, tb, q))
    LRx = operator(mesh.restriction(tx), mesh.restriction(tb), q)
    out = _lib.zeros((1,), torch.float64)
    _lib.check(_lib.load().psc_diff_sumsq(_lib.ptr(RLx), 4.0, _lib.ptr(LRx), RLx.numel(), _lib.ptr(out),
                                          _lib.stream()))
    return np.float32(np.sqrt(out.item()))
