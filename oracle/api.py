"""oracle/api.py -- TEST INFRASTRUCTURE ONLY (see oracle/pysco_oracle.c header).

ctypes/NumPy front-end of the CPU restatement, exposing the reference's module/function names
(``mesh.TSC``, ``fourier.inverse_laplacian_compensated``, ``laplacian.gauss_seidel`` ...) so
parity tests read like calls into PySCo itself.  FFTs use ``numpy.fft`` exactly as the reference
does when pyfftw is absent (fourier.py:129-130, 276-277).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.
"""
import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpysco_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/pysco_oracle.c -> oracle/_build/libpysco_oracle.so (gcc, OpenMP)."""
    src = os.path.join(_HERE, "pysco_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_max_abs.restype = C.c_float
        _lib.orc_lap_residual_error.restype = C.c_float
        _lib.orc_lap_truncation_error.restype = C.c_float
        _lib.orc_fr_residual_error.restype = C.c_float
        _lib.orc_fr_truncation_error.restype = C.c_float
        _lib.orc_solution_cubic_equation.restype = C.c_float
        _lib.orc_solution_quartic_equation.restype = C.c_float
        _lib.orc_solution_cubic_equation.argtypes = [C.c_float, C.c_float]
        _lib.orc_solution_quartic_equation.argtypes = [C.c_float, C.c_float]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _chk(a, dtype=np.float32):
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous, (
        f"expected C-contiguous {dtype}, got {type(a)} {getattr(a, 'dtype', None)}")
    return a


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(int(n)))


def num_threads() -> int:
    return int(lib().orc_num_threads())


# --------------------------------------------------------------------------------------------
# morton / utils
# --------------------------------------------------------------------------------------------
def _positions_to_keys(position):
    pos = _chk(position)
    keys = np.empty(pos.shape[0], dtype=np.int64)
    lib().orc_morton_keys(_p(pos), C.c_int64(pos.shape[0]), _p(keys))
    return keys


morton = SimpleNamespace(positions_to_keys=_positions_to_keys)


def _gather3(idx, a):
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.empty_like(a)
    lib().orc_gather3(_p(idx), _p(_chk(a)), _p(out), C.c_int64(a.shape[0]))
    return out


def _reorder_particles(position, velocity=None, acceleration=None):
    """utils.py:1019-1075 with nthreads == 1 semantics (global argsort); stable sort so that the
    order is defined also for equal keys (the reference's np.argsort is unstable on ties)."""
    keys = _positions_to_keys(position)
    arg = np.argsort(keys, kind="stable")
    if acceleration is not None:
        return _gather3(arg, position), _gather3(arg, velocity), _gather3(arg, acceleration)
    if velocity is not None:
        return _gather3(arg, position), _gather3(arg, velocity)
    return _gather3(arg, position)


def _add_vector_scalar_inplace(y, x, a):
    _chk(y), _chk(x)
    n = C.c_int64(y.size)
    if isinstance(a, np.float32):
        lib().orc_add_vector_scalar_inplace_f32(_p(y), _p(x), C.c_float(float(a)), n)
    else:
        lib().orc_add_vector_scalar_inplace_f64(_p(y), _p(x), C.c_double(float(a)), n)


def _prod_vector_scalar_inplace(y, a):
    lib().orc_prod_vector_scalar_inplace(_p(_chk(y)), C.c_float(float(np.float32(a))), C.c_int64(y.size))


def _linear_operator(x, f1, f2):
    out = np.empty_like(_chk(x))
    lib().orc_linear_operator(_p(x), C.c_float(float(np.float32(f1))), C.c_float(float(np.float32(f2))),
                              _p(out), C.c_int64(x.size))
    return out


def _linear_operator_inplace(x, f1, f2):
    lib().orc_linear_operator(_p(_chk(x)), C.c_float(float(np.float32(f1))),
                              C.c_float(float(np.float32(f2))), _p(x), C.c_int64(x.size))


def _linear_operator_vectors_inplace(x, f1, y, f2):
    lib().orc_linear_operator_vectors_inplace(_p(_chk(x)), C.c_float(float(np.float32(f1))), _p(_chk(y)),
                                              C.c_float(float(np.float32(f2))), C.c_int64(x.size))


def _max_abs(x):
    return np.float32(lib().orc_max_abs(_p(_chk(x)), C.c_int64(x.size)))


def _periodic_wrap(position):
    lib().orc_periodic_wrap(_p(_chk(position)), C.c_int64(position.size))


utils = SimpleNamespace(
    reorder_particles=_reorder_particles,
    injection_with_indices=_gather3,
    add_vector_scalar_inplace=_add_vector_scalar_inplace,
    prod_vector_scalar_inplace=_prod_vector_scalar_inplace,
    linear_operator=_linear_operator,
    linear_operator_inplace=_linear_operator_inplace,
    linear_operator_vectors_inplace=_linear_operator_vectors_inplace,
    max_abs=_max_abs,
    periodic_wrap=_periodic_wrap,
)


# --------------------------------------------------------------------------------------------
# mesh
# --------------------------------------------------------------------------------------------
def _deposit(fn, position, ncells_1d, parallel):
    pos = _chk(position)
    N = int(ncells_1d)
    rho = np.empty((N, N, N), dtype=np.float32)
    getattr(lib(), fn)(_p(pos), C.c_int64(pos.shape[0]), C.c_int(N), _p(rho), C.c_int(parallel))
    return rho


def _inv(fn, grid, position):
    g = _chk(grid)
    pos = _chk(position)
    N = g.shape[0]
    ncomp = 3 if g.ndim == 4 else 1
    out = np.empty((pos.shape[0], 3) if ncomp == 3 else (pos.shape[0],), dtype=np.float32)
    getattr(lib(), fn)(_p(g), _p(pos), C.c_int64(pos.shape[0]), C.c_int(N), C.c_int(ncomp), _p(out))
    return out


def _derivative(a, order, b=None, f=0.0, fr_n=0):
    if order not in (2, 3, 5, 7):
        raise NotImplementedError(f"Unsupported: gradient_order={order}")
    a = _chk(a)
    N = a.shape[0]
    out = np.empty((N, N, N, 3), dtype=np.float32)
    lib().orc_derivative(_p(a), _p(b) if b is not None else None, C.c_float(float(np.float32(f))),
                         C.c_int(fr_n), C.c_int(order), C.c_int(0), C.c_int(N), _p(out))
    return out


def _derivative_fR(a, b, f, fR_n, order):
    if fR_n not in (1, 2):
        raise NotImplementedError(f"Unsupported: fR_n={fR_n}")
    return _derivative(a, order, _chk(b), f, fR_n)


def _add_derivative_fR(force, b, f, fR_n, order):
    if fR_n not in (1, 2):
        raise NotImplementedError(f"Unsupported: fR_n={fR_n}")
    if order not in (2, 3, 5, 7):
        raise NotImplementedError(f"Unsupported: gradient_order={order}")
    N = b.shape[0]
    lib().orc_derivative(None, _p(_chk(b)), C.c_float(float(np.float32(f))), C.c_int(fR_n),
                         C.c_int(order), C.c_int(1), C.c_int(N), _p(_chk(force)))


def _restriction(x, sign=1.0):
    x = _chk(x)
    N = x.shape[0]
    out = np.empty((N // 2,) * 3, dtype=np.float32)
    lib().orc_restriction(_p(x), C.c_int(N), C.c_float(sign), _p(out))
    return out


def _prolongation(x):
    x = _chk(x)
    Nc = x.shape[0]
    y = np.empty((2 * Nc,) * 3, dtype=np.float32)
    lib().orc_prolongation(_p(y), _p(x), C.c_int(Nc), C.c_int(0))
    return y


def _add_prolongation(y, x):
    lib().orc_prolongation(_p(_chk(y)), _p(_chk(x)), C.c_int(x.shape[0]), C.c_int(1))


def _deposit_f64(position, ncells_1d, scheme):
    """float64-accumulated deposit (exact-sum yardstick for tests; scheme 1 = CIC, 2 = TSC)."""
    pos = _chk(position)
    N = int(ncells_1d)
    rho = np.empty((N, N, N), dtype=np.float64)
    lib().orc_deposit_f64(_p(pos), C.c_int64(pos.shape[0]), C.c_int(N), C.c_int(scheme), _p(rho))
    return rho


mesh = SimpleNamespace(
    deposit_f64=_deposit_f64,
    NGP=lambda position, n: _deposit("orc_ngp", position, n, 0),
    CIC=lambda position, n: _deposit("orc_cic", position, n, 0),
    TSC=lambda position, n: _deposit("orc_tsc", position, n, 1),
    TSC_seq=lambda position, n: _deposit("orc_tsc", position, n, 0),
    CIC_par=lambda position, n: _deposit("orc_cic", position, n, 1),
    invNGP=lambda grid, position: _inv("orc_inv_ngp", grid, position),
    invNGP_vec=lambda grid, position: _inv("orc_inv_ngp", grid, position),
    invCIC=lambda grid, position: _inv("orc_inv_cic", grid, position),
    invCIC_vec=lambda grid, position: _inv("orc_inv_cic", grid, position),
    invTSC=lambda grid, position: _inv("orc_inv_tsc", grid, position),
    invTSC_vec=lambda grid, position: _inv("orc_inv_tsc", grid, position),
    derivative=_derivative,
    derivative_fR=_derivative_fR,
    add_derivative_fR=_add_derivative_fR,
    restriction=_restriction,
    minus_restriction=lambda x: _restriction(x, -1.0),
    prolongation=_prolongation,
    add_prolongation=_add_prolongation,
)


# --------------------------------------------------------------------------------------------
# fourier
# --------------------------------------------------------------------------------------------
def _fft_3D_real(x, threads=1):
    return np.ascontiguousarray(np.fft.rfftn(x).astype(np.complex64))


def _ifft_3D_real(x, threads=1):
    return np.ascontiguousarray(np.fft.irfftn(x).astype(np.float32))


def _ifft_3D_real_grad(x, threads=1):
    return np.ascontiguousarray(np.fft.irfftn(x, axes=(0, 1, 2)).astype(np.float32))


def _green(kind):
    def f(x, p=0):
        _chk(x, np.complex64)
        N = x.shape[0]
        if kind == "plain":
            lib().orc_green_plain(_p(x), C.c_int(N))
        elif kind == "comp":
            lib().orc_green_compensated(_p(x), C.c_int(N), C.c_int(int(p)))
        else:
            lib().orc_green_7pt(_p(x), C.c_int(N))
    return f


def _grad_green(x, p=0):
    _chk(x, np.complex64)
    N = x.shape[0]
    out = np.empty((N, N, N // 2 + 1, 3), dtype=np.complex64)
    lib().orc_grad_green(_p(x), C.c_int(N), C.c_int(int(p)), _p(out))
    return out


def _fourier_grid_to_Pk(density_k, p):
    _chk(density_k, np.complex64)
    N = density_k.shape[0]
    n = int(lib().orc_pk_len(C.c_int(N)))
    k = np.empty(n, dtype=np.float32)
    pk = np.empty(n, dtype=np.float32)
    nm = np.empty(n, dtype=np.float32)
    lib().orc_pk(_p(density_k), C.c_int(N), C.c_int(int(p)), _p(k), _p(pk), _p(nm))
    return k, pk, nm


fourier = SimpleNamespace(
    fft_3D_real=_fft_3D_real,
    ifft_3D_real=_ifft_3D_real,
    ifft_3D_real_grad=_ifft_3D_real_grad,
    inverse_laplacian=_green("plain"),
    inverse_laplacian_compensated=_green("comp"),
    inverse_laplacian_7pt=_green("7pt"),
    gradient_inverse_laplacian=lambda x: _grad_green(x, 0),
    gradient_inverse_laplacian_compensated=_grad_green,
    fourier_grid_to_Pk=_fourier_grid_to_Pk,
)


# --------------------------------------------------------------------------------------------
# laplacian / cubic / quartic / mond
# --------------------------------------------------------------------------------------------
def _lap_operator(x):
    out = np.empty_like(_chk(x))
    lib().orc_lap_operator(_p(x), C.c_int(x.shape[0]), _p(out))
    return out


def _lap_residual(x, b):
    out = np.empty_like(_chk(x))
    lib().orc_lap_residual(_p(x), _p(_chk(b)), C.c_int(x.shape[0]), _p(out))
    return out


def _lap_restrict_residual(x, b):
    N = x.shape[0]
    out = np.empty((N // 2,) * 3, dtype=np.float32)
    lib().orc_lap_restrict_residual(_p(_chk(x)), _p(_chk(b)), C.c_int(N), _p(out))
    return out


def _lap_init(b):
    out = np.empty_like(_chk(b))
    lib().orc_lap_initialise_potential(_p(b), C.c_int(b.shape[0]), _p(out))
    return out


def _lap_gs(x, b, f_relax):
    lib().orc_lap_gauss_seidel(_p(_chk(x)), _p(_chk(b)), C.c_int(x.shape[0]), C.c_float(float(f_relax)))


def _lap_smoothing(x, b, n_smoothing):
    for _ in range(int(n_smoothing)):
        _lap_gs(x, b, np.float32(1.25))


laplacian = SimpleNamespace(
    operator=_lap_operator,
    residual=_lap_residual,
    restrict_residual=_lap_restrict_residual,
    residual_error=lambda x, b: np.float32(
        lib().orc_lap_residual_error(_p(_chk(x)), _p(_chk(b)), C.c_int(x.shape[0]))),
    truncation_error=lambda x: np.float32(lib().orc_lap_truncation_error(_p(_chk(x)), C.c_int(x.shape[0]))),
    initialise_potential=_lap_init,
    gauss_seidel=_lap_gs,
    smoothing=_lap_smoothing,
)


def _fr_ns(kind):
    def operator(x, b, q):
        out = np.empty_like(_chk(x))
        lib().orc_fr_operator(_p(x), _p(_chk(b)), C.c_float(float(q)), C.c_int(x.shape[0]), C.c_int(kind), _p(out))
        return out

    def residual_with_rhs(x, b, q, rhs):
        out = np.empty_like(_chk(x))
        lib().orc_fr_residual_with_rhs(_p(x), _p(_chk(b)), C.c_float(float(q)), _p(_chk(rhs)),
                                       C.c_int(x.shape[0]), C.c_int(kind), _p(out))
        return out

    def initialise_potential(b, q):
        out = np.empty_like(_chk(b))
        lib().orc_fr_initialise_potential(_p(b), C.c_float(float(q)), C.c_int(b.shape[0]), C.c_int(kind), _p(out))
        return out

    def gauss_seidel(x, b, q, f_relax):
        lib().orc_fr_gauss_seidel(_p(_chk(x)), _p(_chk(b)), C.c_float(float(q)), None, C.c_int(x.shape[0]),
                                  C.c_int(kind), C.c_float(float(f_relax)))

    def gauss_seidel_with_rhs(x, b, q, rhs, f_relax):
        lib().orc_fr_gauss_seidel(_p(_chk(x)), _p(_chk(b)), C.c_float(float(q)), _p(_chk(rhs)),
                                  C.c_int(x.shape[0]), C.c_int(kind), C.c_float(float(f_relax)))

    def smoothing(x, b, q, n_smoothing):
        for _ in range(int(n_smoothing)):
            gauss_seidel(x, b, q, np.float32(1.25))

    def smoothing_with_rhs(x, b, q, n_smoothing, rhs):
        for _ in range(int(n_smoothing)):
            gauss_seidel_with_rhs(x, b, q, rhs, np.float32(1.25))

    def residual_error(x, b, q):
        return np.float32(lib().orc_fr_residual_error(_p(_chk(x)), _p(_chk(b)), C.c_float(float(q)),
                                                      C.c_int(x.shape[0]), C.c_int(kind)))

    def truncation_error(x, b, q):
        return np.float32(lib().orc_fr_truncation_error(_p(_chk(x)), _p(_chk(b)), C.c_float(float(q)),
                                                        C.c_int(x.shape[0]), C.c_int(kind)))

    ns = SimpleNamespace(operator=operator, residual_with_rhs=residual_with_rhs,
                         initialise_potential=initialise_potential, gauss_seidel=gauss_seidel,
                         gauss_seidel_with_rhs=gauss_seidel_with_rhs, smoothing=smoothing,
                         smoothing_with_rhs=smoothing_with_rhs, residual_error=residual_error,
                         truncation_error=truncation_error)
    if kind == 1:
        ns.solution_cubic_equation = lambda p, d1: np.float32(
            lib().orc_solution_cubic_equation(float(np.float32(p)), float(np.float32(d1))))
    else:
        ns.solution_quartic_equation = lambda p, q: np.float32(
            lib().orc_solution_quartic_equation(float(np.float32(p)), float(np.float32(q))))
    return ns


cubic = _fr_ns(1)
quartic = _fr_ns(2)


def _mond_rhs(fn):
    def f(potential, out, g0, alpha=1.0, **kw):
        if kw:
            alpha = list(kw.values())[0]
        lib().orc_mond_rhs(_p(_chk(potential)), _p(_chk(out)), C.c_int(potential.shape[0]),
                           C.c_float(float(np.float32(g0))), C.c_int(fn), C.c_float(float(alpha)))
    return f


mond = SimpleNamespace(rhs_simple=_mond_rhs(0), rhs_n=_mond_rhs(1), rhs_beta=_mond_rhs(2),
                       rhs_gamma=_mond_rhs(3), rhs_delta=_mond_rhs(4))
