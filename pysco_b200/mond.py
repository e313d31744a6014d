"""Mirror of pysco/mond.py: QUMOND right-hand side  div( nu(|grad phi_N|/g0) grad phi_N )."""
import numpy as np

from . import _lib


def _rhs(fn, potential, out, g0, alpha):
    c = _lib.Ctx()
    tp, to = c.dev(potential), c.dev(out, inplace=True)
    _lib.check(_lib.load().psc_mond_rhs(_lib.ptr(tp), _lib.ptr(to), tp.shape[0], float(np.float32(g0)),
                                        _lib.MOND_FN[fn], float(alpha), _lib.stream()))
    c.finish()


def rhs_simple(potential, out, g0) -> None:
    """mond.py:171-316"""
    _rhs("simple", potential, out, g0, 1.0)


def rhs_n(potential, out, g0, n) -> None:
    """mond.py:322-470"""
    _rhs("n", potential, out, g0, n)


def rhs_beta(potential, out, g0, beta) -> None:
    """mond.py:476-624"""
    _rhs("beta", potential, out, g0, beta)


def rhs_gamma(potential, out, g0, gamma) -> None:
    """mond.py:630-778"""
    _rhs("gamma", potential, out, g0, gamma)


def rhs_delta(potential, out, g0, delta) -> None:
    """mond.py:784-932"""
    _rhs("delta", potential, out, g0, delta)
