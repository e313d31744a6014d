"""Mirror of the hot-path part of pysco/fourier.py: cuFFT transforms + hand-written spectral kernels."""
import numpy as np
import torch

from . import _lib


def fft_3D_real(x, threads=1):
    """fourier.py:104-147: rfftn -> [N,N,N/2+1] complex64 (unnormalised)."""
    c = _lib.Ctx()
    tx = c.dev(x)
    N = tx.shape[0]
    spec = _lib.empty((N, N, N // 2 + 1), torch.complex64)
    _lib.check(_lib.load().psc_fft_r2c(_lib.fft_plan(N), _lib.ptr(tx), _lib.ptr(spec), _lib.stream()))
    return c.ret(spec)


def ifft_3D_real(x, threads=1, out=None, prescaled=False):
    """fourier.py:251-294: irfftn with the 1/N^3 normalisation.  The input spectrum is consumed
    (cuFFT C2R overwrites its input); NumPy inputs are left untouched (they are copied to the device).
    prescaled=True means 1/N^3 was already folded into the Green multiply."""
    c = _lib.Ctx()
    ts = c.dev(x, torch.complex64)
    N = ts.shape[0]
    res = out if out is not None else _lib.empty((N, N, N))
    lib = _lib.load()
    _lib.check(lib.psc_fft_c2r(_lib.fft_plan(N), _lib.ptr(ts), _lib.ptr(res), _lib.stream()))
    if not prescaled:
        _lib.check(lib.psc_linear_operator(_lib.ptr(res), 1.0 / float(N) ** 3, 0.0, _lib.ptr(res), res.numel(),
                                           _lib.stream()))
    return c.ret(res)


def ifft_3D_real_grad(x, threads=1, prescaled=False):
    """fourier.py:372-410: [N,N,N/2+1,3] complex64 -> AoS [N,N,N,3] float32."""
    c = _lib.Ctx()
    ts = c.dev(x, torch.complex64)
    N = ts.shape[0]
    res = _lib.empty((N, N, N, 3))
    lib = _lib.load()
    _lib.check(lib.psc_fft_c2r_vec3(_lib.fft_plan(N), _lib.ptr(ts), _lib.ptr(res), _lib.stream()))
    if not prescaled:
        _lib.check(lib.psc_linear_operator(_lib.ptr(res), 1.0 / float(N) ** 3, 0.0, _lib.ptr(res), res.numel(),
                                           _lib.stream()))
    return c.ret(res)


def _green(x, kind, p, scale=1.0):
    c = _lib.Ctx()
    ts = c.dev(x, torch.complex64, inplace=True)
    _lib.check(_lib.load().psc_green(_lib.ptr(ts), ts.shape[0], kind, int(p), float(scale), _lib.stream()))
    c.finish()


def inverse_laplacian(x, scale=1.0) -> None:
    """fourier.py:460-491 (in place)"""
    _green(x, _lib.GREEN_PLAIN, 0, scale)


def inverse_laplacian_compensated(x, p, scale=1.0) -> None:
    """fourier.py:502-544 (in place)"""
    _green(x, _lib.GREEN_COMPENSATED, p, scale)


def inverse_laplacian_7pt(x, scale=1.0) -> None:
    """fourier.py:555-595 (in place)"""
    _green(x, _lib.GREEN_7PT, 0, scale)


def gradient_inverse_laplacian_compensated(x, p, scale=1.0):
    """fourier.py:664-719"""
    c = _lib.Ctx()
    ts = c.dev(x, torch.complex64)
    N = ts.shape[0]
    out = _lib.empty((N, N, N // 2 + 1, 3), torch.complex64)
    _lib.check(_lib.load().psc_grad_green(_lib.ptr(ts), N, int(p), float(scale), _lib.ptr(out), _lib.stream()))
    return c.ret(out)


def gradient_inverse_laplacian(x, scale=1.0):
    """fourier.py:606-653"""
    return gradient_inverse_laplacian_compensated(x, 0, scale)


def fourier_grid_to_Pk(density_k, p):
    """fourier.py:22-100 -> (k, Pk, Nmodes) float32 arrays of length int(2*(N/2)/3) - 1.
    Like the reference, zeroes density_k[0,0,0]."""
    c = _lib.Ctx()
    ts = c.dev(density_k, torch.complex64, inplace=True)
    N = ts.shape[0]
    bins = _lib.empty((3, N), torch.float64)
    _lib.check(_lib.load().psc_pk(_lib.ptr(ts), N, int(p), _lib.ptr(bins), _lib.stream()))
    c.finish()
    b = bins.cpu().numpy()
    kmax = int(2 * (N // 2) / 3)
    nm = b[2, 1:kmax]
    with np.errstate(invalid="ignore", divide="ignore"):
        k = (b[0, 1:kmax] / nm).astype(np.float32)
        pk = (b[1, 1:kmax] / nm).astype(np.float32)
    return k, pk, nm.astype(np.float32)
