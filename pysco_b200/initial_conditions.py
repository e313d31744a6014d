"""Mirror of pysco/initial_conditions.py: LPT initial conditions (1LPT / 2LPT / 3LPT) on the device.

SURVEY 8(f) rank 1 -- the row after the hot path: it is what `main.run` needs to start from the shipped
`examples/param.ini` (`initial_conditions = 2LPT`).  One-off setup, not part of the per-step path, so it is built
from device-resident torch tensors and cuFFT through `torch.fft` (no hand-written kernels): the white noise is drawn
on the host with the SAME NumPy generator calls, in the same order, as the reference (initial_conditions.py:615-616,
636-655), everything after it -- transfer multiply, inverse Laplacian, spectral gradients / Hessians, the 2LPT and
3LPT sources, displacements -- runs on the GPU in complex64 / float32 like the reference.

    generate                      initial_conditions.py:25-213
    generate_density_fourier      :401-445        get_transfer_grid :530-576
    white_noise_fourier[_fixed]   :579-722
    compute_2ndorder_rhs          :976-1039       compute_3{a,b,c_Ax,c_Ay,c_Az}_{rhs,displacement} :1042-1632
    initialise_1LPT_{center,edge} :1681-1799      add_nLPT :1802-1855     pad / trim :1858-1927
    fourier.gradient / hessian / sum_of_hessian / diff_of_hessian / gradient_inverse_laplacian (fourier.py:606-960)
"""
import logging
import math
import threading

import numpy as np
import torch

from . import _lib, iostream, utils

_PC = 3.0856775814913673e16  # astropy.constants.pc (m)


# ------------------------------------------------------------------------------------------ white noise (host)
def _hermitian_fill(upper, middle, N):
    """The full [N,N,N] array the reference's sequential loops leave behind (initial_conditions.py:620-635):
    density[i,j,k] = upper[i,j,k] and density[-i,-j,-k] = conj(upper[i,j,k]) for i in 0..middle, later writes
    winning.  Only the half k <= middle is used downstream; it is returned as [N, N, middle + 1]."""
    nz = middle + 1
    out = np.empty((N, N, nz), dtype=np.complex64)
    # planes 0 < i < middle: written once as "upper"; planes N - i: conj(upper[i, -j, -k])
    out[1:middle] = upper[1:middle, :, :nz]
    neg = (-np.arange(N)) % N
    for i in range(1, middle):
        out[N - i] = np.conj(upper[i][neg][:, neg][:, :nz])
    # self-conjugate planes i = 0 and i = middle: entry e = (j, k) and its partner p = (-j, -k) are both written at
    # iteration e and at iteration p; the later iteration (lexicographic j, k) wins and leaves upper at its own
    # index, conj(upper) at the partner (for e == p the conjugate is written last)
    jj, kk = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    order_e = jj * N + kk
    order_p = neg[jj] * N + neg[kk]
    for i0 in (0, middle):
        u = upper[i0]
        conj_partner = np.conj(u[neg][:, neg])
        plane = np.where(order_e > order_p, u, conj_partner)
        out[i0] = plane[:, :nz]
    return out


def white_noise_fourier(N, rng):
    """initial_conditions.py:585-655 (Rayleigh amplitudes, uniform phases); half-spectrum [N, N, N/2+1]"""
    middle = N // 2
    twopi = np.float32(2 * math.pi)
    one = np.float32(1)
    amp = rng.random((middle + 1, N, N), dtype=np.float32)
    pha = rng.random((middle + 1, N, N), dtype=np.float32)
    phase = twopi * pha
    amplitude = np.sqrt(-np.log(one - amp))
    upper = (amplitude * np.cos(phase) + 1j * (amplitude * np.sin(phase))).astype(np.complex64)
    del amp, pha, phase, amplitude
    d = _hermitian_fill(upper, middle, N)
    d[0, 0, 0] = 0
    for idx in ((0, 0, middle), (0, middle, 0), (0, middle, middle), (middle, 0, 0), (middle, 0, middle),
                (middle, middle, 0), (middle, middle, middle)):
        v = np.float32(math.sqrt(-math.log(one - rng.random(dtype=np.float32))))
        if idx[2] <= middle:
            d[idx] = v
    return d


def white_noise_fourier_fixed(N, rng, is_paired):
    """initial_conditions.py:664-722 (unit amplitudes; paired: phases shifted by pi)"""
    middle = N // 2
    twopi = np.float32(2 * np.pi)
    shift = np.float32(math.pi) if is_paired else np.float32(0)
    pha = rng.random((middle + 1, N, N), dtype=np.float32)
    phase = twopi * pha + shift
    upper = (np.cos(phase) + 1j * np.sin(phase)).astype(np.complex64)
    d = _hermitian_fill(upper, middle, N)
    d[0, 0, 0] = 0
    for idx in ((0, 0, middle), (0, middle, 0), (0, middle, middle), (middle, 0, 0), (middle, 0, middle),
                (middle, middle, 0), (middle, middle, middle)):
        d[idx] = 1
    return d


def get_transfer_grid(param):
    """initial_conditions.py:531-576: sqrt(P(k)) N^3 / L^1.5 interpolated on the grid (float64), half grid only"""
    k, Pk = np.loadtxt(param["power_spectrum_file"]).T
    N = int(round(param["npart"] ** (1.0 / 3)))
    if param["npart"] != N ** 3:
        raise ValueError(f"{math.cbrt(param['npart'])=}, should be integer")
    kf = 2 * np.pi / param["boxlen"]
    k_dimensionless = k / kf
    sqrtPk = (np.sqrt(Pk / param["boxlen"] ** 3) * N ** 3).astype(np.float32)
    k_1d = np.fft.fftfreq(N, 1 / N)
    kz = k_1d[: N // 2 + 1]      # the reference builds the full grid; only k <= N/2 is read afterwards
    k_grid = np.sqrt(kz[np.newaxis, np.newaxis, :] ** 2 + k_1d[:, np.newaxis, np.newaxis] ** 2
                     + k_1d[np.newaxis, :, np.newaxis] ** 2)
    return np.interp(k_grid, k_dimensionless, sqrtPk)


# ---- the same white noise, one y block of the transposed spectrum at a time (x-slab runs: SlabLayout)
class _NoiseStream:
    """Random access into the float32 draws of numpy.random.default_rng(seed): Generator.random(dtype=float32) takes
    the 32-bit halves of PCG64's 64-bit outputs in order (low half first), so element e of a draw that started on a
    64-bit boundary comes from output e // 2 -- PCG64.advance() jumps there in O(log) time.  Every rank reproduces
    exactly the numbers the reference draws for its block without drawing the rest."""

    def __init__(self, seed):
        self.bg = np.random.PCG64(seed)
        self.start = self.bg.state

    def floats(self, first, count):
        """elements [first, first + count) of the float32 stream (first even)"""
        assert first % 2 == 0
        self.bg.state = self.start
        self.bg.advance(first // 2)
        return np.random.Generator(self.bg).random(count, dtype=np.float32)


def _rows_of_stream(stream, offset, i, rows, N):
    """rows `rows` (ascending) of plane i of an [.., N, N] float32 draw that starts at element `offset`: [len, N]"""
    out = np.empty((len(rows), N), dtype=np.float32)
    a = 0
    while a < len(rows):            # contiguous runs of rows: one jump each
        b = a + 1
        while b < len(rows) and rows[b] == rows[b - 1] + 1:
            b += 1
        out[a:b] = stream.floats(offset + (i * N + int(rows[a])) * N, (b - a) * N).reshape(b - a, N)
        a = b
    return out


def white_noise_fourier_block(N, seed, y0, nyl, fixed=False, paired=False):
    """The block [:, y0:y0+nyl, :] of white_noise_fourier(N, default_rng(seed)) (or of white_noise_fourier_fixed),
    bit for bit, from 2 nyl / N of the random numbers: plane i needs the rows of the block and, for the conjugate
    half (planes N - i and the self-conjugate planes 0, N/2), the rows -j of the block.  complex64 [N, nyl, N/2+1]"""
    middle = N // 2
    nz = middle + 1
    stream = _NoiseStream(seed)
    plane = (middle + 1) * N * N                       # elements of one (middle + 1, N, N) draw
    twopi, one = np.float32(2 * math.pi), np.float32(1)
    shift = np.float32(math.pi) if paired else np.float32(0)
    J = np.arange(y0, y0 + nyl)
    negJ = (-J) % N
    rows = np.unique(np.concatenate([J, negJ]))        # ascending: what _rows_of_stream wants
    at_J, at_negJ = np.searchsorted(rows, J), np.searchsorted(rows, negJ)
    K = np.arange(nz)
    negK = (-K) % N
    later = (J[:, None] * N + K[None, :]) > (negJ[:, None] * N + negK[None, :])    # the write that wins (planes 0, N/2)
    out = np.empty((N, nyl, nz), dtype=np.complex64)
    for i in range(middle + 1):
        if fixed:
            phase = twopi * _rows_of_stream(stream, 0, i, rows, N) + shift
            upper = (np.cos(phase) + 1j * np.sin(phase)).astype(np.complex64)
        else:
            amp = _rows_of_stream(stream, 0, i, rows, N)
            phase = twopi * _rows_of_stream(stream, plane, i, rows, N)
            amplitude = np.sqrt(-np.log(one - amp))
            upper = (amplitude * np.cos(phase) + 1j * (amplitude * np.sin(phase))).astype(np.complex64)
        own = upper[at_J][:, :nz]
        partner = np.conj(upper[at_negJ][:, negK])
        if i in (0, middle):
            out[i] = np.where(later, own, partner)
        else:
            out[i] = own
            out[N - i] = partner
    # the eight real modes (initial_conditions.py:636-655): drawn after the two big draws, in this order
    specials = ((0, 0, middle), (0, middle, 0), (0, middle, middle), (middle, 0, 0), (middle, 0, middle),
                (middle, middle, 0), (middle, middle, middle))
    if fixed:
        values = [np.float32(1)] * 7
    else:
        stream.bg.state = stream.start
        stream.bg.advance(plane)                       # two draws of `plane` floats = plane 64-bit outputs
        g = np.random.Generator(stream.bg)
        values = [np.float32(math.sqrt(-math.log(one - g.random(dtype=np.float32)))) for _ in specials]
    for (i, j, k), v in zip(specials, values):
        if y0 <= j < y0 + nyl:
            out[i, j - y0, k] = v
    if y0 == 0:
        out[0, 0, 0] = 0
    return out


def get_transfer_grid_block(param, y0, nyl, x0=0, nx=None):
    """get_transfer_grid for the block [x0:x0+nx, y0:y0+nyl, :] (same float64 expressions element by element)"""
    k, Pk = np.loadtxt(param["power_spectrum_file"]).T
    N = int(round(param["npart"] ** (1.0 / 3)))
    if param["npart"] != N ** 3:
        raise ValueError(f"{math.cbrt(param['npart'])=}, should be integer")
    kf = 2 * np.pi / param["boxlen"]
    sqrtPk = (np.sqrt(Pk / param["boxlen"] ** 3) * N ** 3).astype(np.float32)
    k_1d = np.fft.fftfreq(N, 1 / N)
    kz = k_1d[: N // 2 + 1]
    ky = k_1d[y0:y0 + nyl]
    kx = k_1d[x0:N if nx is None else x0 + nx]
    k_grid = np.sqrt(kz[np.newaxis, np.newaxis, :] ** 2 + kx[:, np.newaxis, np.newaxis] ** 2
                     + ky[np.newaxis, :, np.newaxis] ** 2)
    return np.interp(k_grid, k / kf, sqrtPk)


def _periodic_wrap(x):
    """utils.periodic_wrap (utils.py:1120-1149) with torch ops (device-agnostic; the step's CUDA kernel does the same)"""
    eps = -2.98023223876953125e-08 * (1.0 + 1e-6)
    neg = torch.where(x.double() > eps, torch.zeros_like(x), x + 1.0)
    return torch.where(x < 0, neg, torch.where(x >= 1.0, x - 1.0, x))


def generate_density_fourier(param, device=None):
    """initial_conditions.py:402-445 -> device complex64 [N, N, N/2+1]"""
    L = _layout()
    if not L.whole:
        # this rank's block [N, nyl, N/2+1] of the same spectrum (bit-identical to the slice of the full one)
        seed = int(param["seed"])
        if seed < 0:            # "random": every rank must still draw from the same stream
            t = torch.tensor([float(np.random.default_rng().integers(0, 2 ** 24))],
                             device=_lib.device() if device is None else device)
            L.comm.allreduce_max_(t)
            seed = int(t[0])
        d = white_noise_fourier_block(L.N, seed, L.y0, L.nyl, bool(param["fixed_ICS"]), bool(param["paired_ICS"]))
        step = max(1, (1 << 22) // (L.nyl * L.nz))     # kx planes per chunk: float64 temporaries of ~32 MB
        for a in range(0, L.N, step):
            b = min(L.N, a + step)
            d[a:b] = (d[a:b] * get_transfer_grid_block(param, L.y0, L.nyl, a, b - a)).astype(np.complex64)
        return torch.from_numpy(np.ascontiguousarray(d)).to(_lib.device() if device is None else device)
    transfer = get_transfer_grid(param)
    N = transfer.shape[0]
    seed = param["seed"]
    rng = np.random.default_rng() if seed < 0 else np.random.default_rng(seed)
    if param["fixed_ICS"]:
        d = white_noise_fourier_fixed(N, rng, param["paired_ICS"])
    else:
        d = white_noise_fourier(N, rng)
    d = (d * transfer).astype(np.complex64)   # complex64 *= float64, rounded once like the reference
    return torch.from_numpy(np.ascontiguousarray(d)).to(_lib.device() if device is None else device)


# ------------------------------------------------------------------------------------------ spectral operators
def _kvec(N, dev):
    """signed integer wave numbers with index >= N/2 -> index - N (fourier.py:754-766), as float32 [N], [N/2+1]"""
    i = torch.arange(N, device=dev)
    kfull = torch.where(i >= N // 2, i - N, i).to(torch.float32)
    kz = torch.arange(N // 2 + 1, device=dev, dtype=torch.float32)
    return kfull, kz


class _WholeLayout:
    """One process holds the whole half-spectrum [N, N, N/2+1] and the whole real grid [N, N, N]."""
    has_dc = True       # element [0, 0, 0] of a spectrum is the k = 0 mode
    whole = True

    def k_of(self, axis, N, dev):
        kfull, kz = _kvec(N, dev)
        if axis == 0:
            return kfull[:, None, None]
        if axis == 1:
            return kfull[None, :, None]
        return kz[None, None, :]

    def fft(self, x):
        return torch.fft.rfftn(x, dim=(0, 1, 2))

    def ifft(self, x):
        N = x.shape[0]
        return torch.fft.irfftn(x, s=(N, N, N), dim=(0, 1, 2))


class SlabLayout:
    """The grids of an x-slab decomposition over the P ranks of `comm` (pysco_b200/slab.py): real fields are this
    rank's planes [nxl, N, N] (x0 = rank nxl), spectra are TRANSPOSED blocks [N, nyl, N/2+1] -- all of kx, the y block
    [y0, y0 + nyl) of ky, half of kz -- exactly the layout of Slab.fft_poisson.  The transforms are torch.fft along
    (y, z) on the planes, one equal-size all-to-all, torch.fft along x (and back): one-off setup work, so library FFTs
    and the comm's generic all-to-all rather than the peer-memory kernels of the time loop."""
    whole = False

    def __init__(self, comm, N):
        self.comm, self.N, self.P, self.rank = comm, int(N), comm.size, comm.rank
        if self.N % self.P:
            raise ValueError(f"N = {N} planes cannot be split over {self.P} ranks")
        self.nxl = self.nyl = self.N // self.P
        self.x0 = self.y0 = self.rank * self.nxl
        self.nz = self.N // 2 + 1
        self.has_dc = self.rank == 0

    def k_of(self, axis, N, dev):
        """N: the grid the spectrum belongs to (the particle grid or its 3/2 dealiasing grid)"""
        kfull, kz = _kvec(N, dev)
        if axis == 0:
            return kfull[:, None, None]
        if axis == 1:
            nyl = N // self.P
            return kfull[None, self.rank * nyl:(self.rank + 1) * nyl, None]
        return kz[None, None, :]

    def _all_to_all(self, send):
        """send [P, ...] complex: block q goes to rank q; returns [P, ...] with block q received from rank q"""
        a = torch.view_as_real(send.contiguous()).reshape(self.P, -1)
        out = torch.empty_like(a)
        self.comm.all_to_all_equal(a, out)
        return torch.view_as_complex(out.reshape(tuple(send.shape) + (2,)))

    def fft(self, x):
        """real planes [n/P, n, n] -> transposed spectrum block [n, n/P, n/2+1] (n: any grid divisible by P)"""
        P, nxl, n = self.P, x.shape[0], x.shape[1]
        nyl, nz = n // P, n // 2 + 1
        a = torch.fft.rfft2(x, dim=(1, 2))                                        # [nxl, n, nz]
        a = a.reshape(nxl, P, nyl, nz).permute(1, 0, 2, 3)                        # [P (y block), nxl, nyl, nz]
        b = self._all_to_all(a)                                                   # [P (x block), nxl, nyl, nz]
        return torch.fft.fft(b.reshape(n, nyl, nz), dim=0)

    def ifft(self, x):
        """transposed spectrum block [n, n/P, n/2+1] -> real planes [n/P, n, n]"""
        P, n, nyl, nz = self.P, x.shape[0], x.shape[1], x.shape[2]
        nxl = n // P
        a = torch.fft.ifft(x, dim=0).reshape(P, nxl, nyl, nz)                     # [P (x block), nxl, nyl, nz]
        b = self._all_to_all(a)                                                   # [P (y block), nxl, nyl, nz]
        b = b.permute(1, 0, 2, 3).reshape(nxl, n, nz)
        return torch.fft.irfft2(b, s=(n, n), dim=(1, 2))

    def regrid(self, x, n_to):
        """pad (n_to = 3 n / 2) or trim (n_to = 2 n / 3) of a transposed spectrum block (initial_conditions.py:1858-1927,
        the Orszag 3/2 rule): the modes |k| < m = min(n, n_to) / 2 of every axis keep their signed wave numbers on
        the other grid, the rest is zero.  Along x and z that is local index arithmetic; the ky rows change owner
        (the y blocks of the two grids differ): one all-to-all-v of rows.  The mapping of kept rows is monotonic, so
        the rows a rank sends to each destination, and receives from each source, are contiguous and in order."""
        P, r = self.P, self.rank
        n, nyl = x.shape[0], x.shape[1]
        if n_to % P:
            raise ValueError(f"dealiasing grid {n_to} cannot be split over {P} ranks")
        m = min(n, n_to) // 2
        nyl_to, nz_to = n_to // P, n_to // 2 + 1
        dev = x.device

        def kept(lo, hi, size):
            """indices in [lo, hi) of a `size` grid that hold a kept mode, and the same modes' indices on the other
            grid (index < m: as is; index > size - m: from the end)"""
            idx = torch.arange(lo, hi, device=dev)
            keep = (idx < m) | (idx > size - m)
            idx = idx[keep]
            other = n_to if size == n else n
            return idx, torch.where(idx < m, idx, idx + (other - size))
        # x and z: local
        xi, xo = kept(0, n, n)
        rows_here, rows_there = kept(r * nyl, (r + 1) * nyl, n)                  # my ky rows that survive, their new index
        part = torch.zeros((len(rows_here), n_to, nz_to), dtype=x.dtype, device=dev)
        src = x[:, rows_here - r * nyl, :m].permute(1, 0, 2)                     # [rows, n, m]
        part[:, xo, :m] = src[:, xi, :]
        dest = torch.div(rows_there, nyl_to, rounding_mode="floor")
        send_counts = [int((dest == q).sum()) for q in range(P)]
        recv_counts = self.comm.exchange_counts(send_counts)
        flat = torch.view_as_real(part.contiguous()).reshape(len(rows_here), -1)
        got = self.comm.all_to_all_v(flat, send_counts, recv_counts)
        got = torch.view_as_complex(got.reshape(-1, n_to, nz_to, 2).contiguous())
        # the rows I now own, in the order they arrive (ascending on the new grid)
        mine, _ = kept(r * nyl_to, (r + 1) * nyl_to, n_to)
        assert got.shape[0] == len(mine), (got.shape, len(mine))
        out = torch.zeros((n_to, nyl_to, nz_to), dtype=x.dtype, device=dev)
        out[:, mine - r * nyl_to, :] = got.permute(1, 0, 2)
        return out


_WHOLE = _WholeLayout()
_tls = threading.local()     # the ranks of a ThreadComm are threads of one process: one layout per thread


def _layout():
    return getattr(_tls, "layout", _WHOLE)


def _k_of(axis, N, dev):
    return _layout().k_of(axis, N, dev)


def _k2(N, dev):
    """kx^2 + ky^2 + kz^2 on the current layout, with the k = 0 mode (if this process holds it) set to 1"""
    k2 = _k_of(0, N, dev) ** 2 + _k_of(1, N, dev) ** 2 + _k_of(2, N, dev) ** 2
    if _layout().has_dc:
        k2[0, 0, 0] = 1.0
    return k2


def inverse_laplacian(x):
    """fourier.inverse_laplacian (fourier.py:460-491) on a half-spectrum, in place"""
    k2 = _k2(x.shape[0], x.device)
    invpi2 = np.float32(-0.25 / np.pi ** 2)
    x *= (invpi2 / k2)
    if _layout().has_dc:
        x[0, 0, 0] = 0
    return x


def gradient(x):
    """fourier.gradient (fourier.py:730-770): i 2 pi k_j x -> [N, N, N/2+1, 3]"""
    N = x.shape[0]
    tmp = x * torch.complex(torch.zeros((), device=x.device), torch.tensor(np.float32(2 * np.pi), device=x.device))
    return torch.stack([tmp * _k_of(a, N, x.device) for a in range(3)], dim=-1)


def hessian(x, ij):
    """fourier.hessian (fourier.py:784-832): -k_n k_m 4 pi^2 x"""
    N = x.shape[0]
    fourpi2 = np.float32(4 * np.pi ** 2)
    return x * (-(_k_of(ij[0], N, x.device) * _k_of(ij[1], N, x.device)) * fourpi2)


def sum_of_hessian(x, ij1, ij2, sign=1.0):
    """fourier.sum_of_hessian / diff_of_hessian (fourier.py:842-960): -(k k +- k k) 4 pi^2 x"""
    N = x.shape[0]
    fourpi2 = np.float32(4 * np.pi ** 2)
    dev = x.device
    kk = _k_of(ij1[0], N, dev) * _k_of(ij1[1], N, dev) + sign * (_k_of(ij2[0], N, dev) * _k_of(ij2[1], N, dev))
    return x * (-kk * fourpi2)


def diff_of_hessian(x, ij1, ij2):
    return sum_of_hessian(x, ij1, ij2, sign=-1.0)


def gradient_inverse_laplacian(x):
    """fourier.gradient_inverse_laplacian (fourier.py:606-653): -i k_j / (2 pi k^2) x, DC = 0"""
    N = x.shape[0]
    dev = x.device
    k2 = _k2(N, dev)
    invtwopi = np.float32(0.5 / np.pi)
    tmp = x * torch.complex(torch.zeros((), device=dev), -(invtwopi / k2))
    out = torch.stack([tmp * _k_of(a, N, dev) for a in range(3)], dim=-1)
    if _layout().has_dc:
        out[0, 0, 0, :] = 0
    return out


def ifft_3D_real(x):
    return _layout().ifft(x)


def fft_3D_real(x):
    return _layout().fft(x)


def ifft_3D_real_grad(x):
    """the three components of a spectral vector field [..., 3] -> real [..., 3]"""
    L = _layout()
    if L.whole:
        N = x.shape[0]
        return torch.fft.irfftn(x, s=(N, N, N), dim=(0, 1, 2)).contiguous()
    return torch.stack([L.ifft(x[..., c]) for c in range(3)], dim=-1)


def pad(x):
    """initial_conditions.py:1859-1893 (Orszag 3/2 rule)"""
    N = x.shape[0]
    if not _layout().whole:
        return _layout().regrid(x, 3 * N // 2)
    Ne, m = 3 * N // 2, N // 2
    out = torch.zeros((Ne, Ne, Ne // 2 + 1), dtype=x.dtype, device=x.device)
    out[:m, :m, :m] = x[:m, :m, :m]
    out[-m + 1:, :m, :m] = x[-m + 1:, :m, :m]
    out[:m, -m + 1:, :m] = x[:m, -m + 1:, :m]
    out[-m + 1:, -m + 1:, :m] = x[-m + 1:, -m + 1:, :m]
    return out


def trim(x):
    """initial_conditions.py:1897-1927"""
    Ne = x.shape[0]
    N = 2 * Ne // 3
    if not _layout().whole:
        return _layout().regrid(x, N)
    m = N // 2
    out = torch.zeros((N, N, m + 1), dtype=x.dtype, device=x.device)
    out[:m, :m, :m] = x[:m, :m, :m]
    out[-m + 1:, :m, :m] = x[-m + 1:, :m, :m]
    out[:m, -m + 1:, :m] = x[:m, -m + 1:, :m]
    out[-m + 1:, -m + 1:, :m] = x[-m + 1:, -m + 1:, :m]
    return out


def _real(x, ij):
    return ifft_3D_real(hessian(x, ij))


def _dealias_in(param, *fields):
    L = _layout()
    if param["dealiased_ICS"] and not L.whole and (3 * L.N // 2) % L.P:
        raise NotImplementedError(f"dealiased_ICS on x-slabs: the 3/2 grid {3 * L.N // 2} must split over {L.P} ranks")
    return [pad(f) for f in fields] if param["dealiased_ICS"] else list(fields)


def _dealias_out(param, phi, power):
    if param["dealiased_ICS"]:
        phi = ifft_3D_real(trim(fft_3D_real(phi)))
        phi = phi * np.float32(1.5 ** power)
    return phi


# ------------------------------------------------------------------------------------------ LPT sources
def compute_2ndorder_rhs(phi1, param):
    """initial_conditions.py:976-1039"""
    (p1,) = _dealias_in(param, phi1)
    phi2 = _real(p1, (0, 0)) * ifft_3D_real(sum_of_hessian(p1, (1, 1), (2, 2)))
    phi2 += _real(p1, (1, 1)) * _real(p1, (2, 2))
    for ij in ((0, 1), (0, 2), (1, 2)):
        t = _real(p1, ij)
        phi2 -= t * t
    return _dealias_out(param, phi2, 3)


def compute_3a_rhs(phi1, param):
    """initial_conditions.py:1042-1121"""
    (p1,) = _dealias_in(param, phi1)
    h = {ij: _real(p1, ij) for ij in ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))}
    phi = h[0, 0] * h[1, 1] * h[2, 2]
    phi += np.float32(2) * h[0, 1] * h[0, 2] * h[1, 2]
    phi -= h[1, 2] * h[1, 2] * h[0, 0]
    phi -= h[0, 2] * h[0, 2] * h[1, 1]
    phi -= h[0, 1] * h[0, 1] * h[2, 2]
    return _dealias_out(param, phi, 6)


def compute_3b_rhs(phi1, phi2, param):
    """initial_conditions.py:1162-1246"""
    p1, p2 = _dealias_in(param, phi1, phi2)
    half = np.float32(0.5)
    phi = _real(p1, (0, 0)) * (half * ifft_3D_real(sum_of_hessian(p2, (1, 1), (2, 2))))
    phi += half * _real(p1, (1, 1)) * ifft_3D_real(sum_of_hessian(p2, (0, 0), (2, 2)))
    phi += half * _real(p1, (2, 2)) * ifft_3D_real(sum_of_hessian(p2, (0, 0), (1, 1)))
    for ij in ((0, 1), (0, 2), (1, 2)):
        phi -= _real(p1, ij) * _real(p2, ij)
    return _dealias_out(param, phi, 3)


def _compute_3c_rhs(phi1, phi2, param, a, b, c, d1, d2):
    """Common form of compute_3c_A{x,y,z}_rhs (initial_conditions.py:1290-1591):
    H1[a] H2[b] - H2[a] H1[b] + H1[c] (H2[d1] - H2[d2]) - H2[c] (H1[d1] - H1[d2])"""
    p1, p2 = _dealias_in(param, phi1, phi2)
    phi = _real(p1, a) * _real(p2, b)
    phi -= _real(p2, a) * _real(p1, b)
    phi += _real(p1, c) * ifft_3D_real(diff_of_hessian(p2, d1, d2))
    phi -= _real(p2, c) * ifft_3D_real(diff_of_hessian(p1, d1, d2))
    return _dealias_out(param, phi, 3)


def compute_3c_Ax_rhs(phi1, phi2, param):
    return _compute_3c_rhs(phi1, phi2, param, (0, 2), (0, 1), (1, 2), (1, 1), (2, 2))


def compute_3c_Ay_rhs(phi1, phi2, param):
    return _compute_3c_rhs(phi1, phi2, param, (0, 1), (1, 2), (0, 2), (2, 2), (0, 0))


def compute_3c_Az_rhs(phi1, phi2, param):
    return _compute_3c_rhs(phi1, phi2, param, (1, 2), (0, 2), (0, 1), (0, 0), (1, 1))


def _displacement(rhs):
    """fft -> gradient_inverse_laplacian -> inverse fft (initial_conditions.py:1124-1160 and siblings)"""
    return ifft_3D_real_grad(gradient_inverse_laplacian(fft_3D_real(rhs)))


# ------------------------------------------------------------------------------------------ particles
def initialise_1LPT(psi, dplus_1, fH, param):
    """initial_conditions.py:1635-1799: lattice (cell centres or edges) minus D1 psi; velocity -D1 f H psi"""
    POSITION = param["position_ICS"].casefold()
    if POSITION not in ("center", "edge"):
        raise NotImplementedError(f"{POSITION=}, should be 'center' or 'edge'")
    N = psi.shape[1]
    h = np.float32(1.0 / N)
    half_h = np.float32(0.5 / N) if POSITION == "center" else np.float32(0)
    ax = half_h + torch.arange(N, device=psi.device, dtype=torch.float32) * h
    L = _layout()
    axx = ax if L.whole else ax[L.x0:L.x0 + L.nxl]       # x-slab: the lattice planes this rank owns
    grid = torch.stack(torch.meshgrid(axx, ax, ax, indexing="ij"), dim=-1)
    dfH = np.float32(dplus_1 * fH)
    position = grid + np.float32(dplus_1) * (-psi)
    velocity = dfH * (-psi)
    return position, velocity


def add_nLPT(position, velocity, psi, dplus_n, fH_n):
    """initial_conditions.py:1809-1855"""
    position += np.float32(dplus_n) * psi
    velocity += np.float32(dplus_n * fH_n) * psi


def finalise_initial_conditions(position, velocity, param, do_reorder):
    """initial_conditions.py:216-280: wrap (+ reorder) and write snapshot 0"""
    if "base" not in param:
        raise ValueError(f"{param.index=}, should contain 'base'")
    position.copy_(_periodic_wrap(position))
    if do_reorder:
        position, velocity = utils.reorder_particles(position, velocity)
    fmt = param["output_snapshot_format"].casefold()
    if fmt == "parquet":
        snap_name = f"{param['base']}/output_00000/particles_{param['extra']}.parquet"
        iostream.write_snapshot_particles_parquet(snap_name, position, velocity)
        param.to_csv(f"{param['base']}/output_00000/param_{param['extra']}.txt", sep="=", header=False)
    elif fmt == "hdf5":
        raise NotImplementedError("output_snapshot_format = HDF5 needs h5py (absent in this image); use parquet")
    else:
        raise NotImplementedError(f"{param['output_snapshot_format']=}, should be 'parquet' or 'hdf5'")
    logging.warning(f"Write initial snapshot...{snap_name=}")
    return position, velocity


def generate(param, tables, write_snapshot=True, device=None):
    """initial_conditions.py:25-213 for initial_conditions in {1LPT, 2LPT, 3LPT}.  Returns device tensors
    (position, velocity) [Npart, 3] in the reference's lattice (lexicographic) order.  `device` (default: the current
    CUDA device) exists for the CPU test tier: this row is built from torch ops only, so the same code is checked
    against the reference's output without a GPU too."""
    IC = param["initial_conditions"]
    if not (isinstance(IC, str) and "lpt" in IC.casefold()):
        raise ValueError(f"{IC=}, should be 1LPT, 2LPT or 3LPT")
    order = IC.casefold()
    if order not in ("1lpt", "2lpt", "3lpt"):
        raise ValueError(f"INITIAL_CONDITIONS={IC!r}, should be 1LPT, 2LPT or 3LPT")
    a_start = 1.0 / (1 + param["z_start"])
    lna_start = np.log(a_start)
    logging.warning(f"{param['z_start']=}")
    Hz = tables[2](lna_start)
    mpc_to_km = 1e3 * _PC
    Hz = Hz * param["unit_t"] / mpc_to_km  # km/s/Mpc to BU

    phi1 = inverse_laplacian(generate_density_fourier(param, device))
    psi1 = ifft_3D_real_grad(gradient(phi1))
    logging.warning("Compute 1LPT contribution")
    dplus_1_z0 = tables[3](0)
    dplus_1 = np.float32(tables[3](lna_start) / dplus_1_z0)
    fH_1 = np.float32(tables[4](lna_start) * Hz)
    position, velocity = initialise_1LPT(psi1, dplus_1, fH_1, param)
    del psi1

    def done():
        pos = position.reshape(-1, 3).contiguous()
        vel = velocity.reshape(-1, 3).contiguous()
        if write_snapshot:
            finalise_initial_conditions(pos, vel, param, do_reorder=False)
        else:
            pos = _periodic_wrap(pos)
        return pos, vel

    if order == "1lpt":
        return done()
    logging.warning("Compute 2LPT contribution")
    phi2 = inverse_laplacian(fft_3D_real(compute_2ndorder_rhs(phi1, param)))
    psi2 = ifft_3D_real_grad(gradient(phi2))
    dplus_2 = np.float32(tables[5](lna_start) / dplus_1_z0 ** 2)
    fH_2 = np.float32(tables[6](lna_start) * Hz)
    add_nLPT(position, velocity, psi2, dplus_2, fH_2)
    del psi2
    if order == "2lpt":
        return done()
    dplus_3a = -np.float32(tables[7](lna_start) / dplus_1_z0 ** 3)
    fH_3a = np.float32(tables[8](lna_start) * Hz)
    dplus_3b = -np.float32(tables[9](lna_start) / dplus_1_z0 ** 3)
    fH_3b = np.float32(tables[10](lna_start) * Hz)
    dplus_3c = -np.float32(tables[11](lna_start) / dplus_1_z0 ** 3)
    fH_3c = np.float32(tables[12](lna_start) * Hz)
    logging.warning("Compute 3LPT contributions")
    add_nLPT(position, velocity, _displacement(compute_3a_rhs(phi1, param)), dplus_3a, fH_3a)
    add_nLPT(position, velocity, _displacement(compute_3b_rhs(phi1, phi2, param)), dplus_3b, fH_3b)
    for rhs in (compute_3c_Ax_rhs, compute_3c_Ay_rhs, compute_3c_Az_rhs):
        add_nLPT(position, velocity, _displacement(rhs(phi1, phi2, param)), dplus_3c, fH_3c)
    return done()


def generate_slab(param, tables, comm, device=None):
    """generate() for an x-slab run (SURVEY 8f rank 1 at the sizes of BASELINE config 5): every rank produces only
    the particles of its own lattice planes [x0, x0 + N/P) -- its y block of the white noise (bit-identical to the
    reference's draw, white_noise_fourier_block), the LPT chain on transposed spectrum blocks / real planes with the
    distributed transforms of SlabLayout -- so no rank ever holds a global array.  One call per rank of `comm`
    (collective).  Returns (position [n, 3], velocity [n, 3], ids [n] int64): wrapped positions, ids = the particle's
    row in the reference's lexicographic lattice order.  A displaced particle may lie outside the slab:
    Slab.set_particles routes it to its owner.  dealiased_ICS: the 3/2 grid must split over the ranks too (SlabLayout.regrid)."""
    N = int(round(param["npart"] ** (1.0 / 3)))
    if param["npart"] != N ** 3:
        raise ValueError(f"{math.cbrt(param['npart'])=}, should be integer")
    L = SlabLayout(comm, N)
    _tls.layout = L
    try:
        pos, vel = generate(param, tables, write_snapshot=False, device=device)
    finally:
        del _tls.layout
    n = pos.shape[0]
    ids = torch.arange(n, dtype=torch.int64, device=pos.device) + L.x0 * N * N
    return pos, vel, ids
