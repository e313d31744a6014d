"""CPU stand-in for pysco_b200.slab.CudaOps, built on the oracle (test infrastructure only).

It lets the world_size-2 gloo test run the slab decomposition's HOST logic (migration bookkeeping, ghost-plane
exchanges, the y-block all-to-all of the transposed FFT, time-step reduction) without a GPU: every kernel call is
answered by the C/NumPy oracle working on the full periodic grid and cut down to the slab.
"""
import numpy as np
import torch

import oracle
from oracle import api as oapi

REC = 8

_HX = None


def _harness():
    """tests/slab_mg_harness.cpp (the per-cell bodies of csrc/slab_mg.cu run on the host), built on first use"""
    global _HX
    if _HX is None:
        import ctypes as C
        import os
        import subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        src = os.path.join(here, "slab_mg_harness.cpp")
        hdr = os.path.join(here, "..", "pysco_b200", "csrc", "slab_mg_cells.cuh")
        out = os.path.join(here, "_build", "libslab_mg_harness.so")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            tmp = f"{out}.{os.getpid()}.tmp"
            subprocess.check_call(["g++", "-O2", "-x", "c++", "-shared", "-fPIC", "-ffp-contract=off", "-o", tmp, src])
            os.replace(tmp, out)
        _HX = C.CDLL(out)
        for f in ("hx_gs_colour", "hx_operator", "hx_restrict_residual", "hx_restriction", "hx_add_prolongation",
                  "hx_mond_rhs", "hx_gs_colour_fr", "hx_operator_fr", "hx_init_fr"):
            getattr(_HX, f).restype = None
    return _HX


def _fp(t):
    import ctypes as C
    assert t.dtype == torch.float32 and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def _np(t):
    return t.detach().cpu().numpy()


class OracleOps:
    def __init__(self, N, P, rank):
        self.N, self.P, self.rank = N, P, rank
        self.nxl = N // P
        self.x0 = rank * self.nxl
        self.nyl = N // P
        self.y0 = rank * self.nyl
        self.dev = torch.device("cpu")
        self._binned = None

    # -- particles
    def kick_drift_wrap(self, pos, vel, acc, half_dt, dt, dt_is_f64):
        p, v, a = _np(pos), _np(vel), _np(acc)
        oracle.utils.add_vector_scalar_inplace(v, a, -np.float32(half_dt))
        oracle.utils.add_vector_scalar_inplace(p, v, dt if dt_is_f64 else np.float32(dt))
        oracle.utils.periodic_wrap(p)

    def kick_drift_wrap_detect(self, pos, vel, acc, half_dt, dt, dt_is_f64):
        self.kick_drift_wrap(pos, vel, acc, half_dt, dt, dt_is_f64)
        own = self._owner(pos)
        rows = np.nonzero(own != self.rank)[0].astype(np.int64)
        counts = np.bincount(own[rows], minlength=self.P).astype(np.int64)
        # a deliberately small list every third call exercises the overflow fallback (full scan)
        self._calls = getattr(self, "_calls", 0) + 1
        cap = max(1, rows.size // 2) if self._calls % 3 == 0 else rows.size + 7
        lst = np.full(cap, -1, dtype=np.int64)
        lst[:min(cap, rows.size)] = rows[::-1][:cap]
        return torch.from_numpy(np.concatenate([counts, [rows.size]])), torch.from_numpy(lst)

    def pack_rows(self, pos, vel, ids, rows, offsets, nout):
        r = np.sort(_np(rows))
        own = self._owner(pos)[r]
        r = r[np.argsort(own, kind="stable")]
        assert r.size == nout
        rec = np.empty((nout, REC), dtype=np.float32)
        rec[:, 0:3] = _np(pos)[r]
        rec[:, 3:6] = _np(vel)[r]
        rec[:, 6:8] = _np(ids)[r].astype(np.int64).view(np.float32).reshape(-1, 2)
        return torch.from_numpy(rec), torch.from_numpy(r.astype(np.int64))

    def pack_fixed(self, pos, vel, ids, detected, cap):
        own = self._owner(pos)
        left, right = (self.rank - 1) % self.P, (self.rank + 1) % self.P
        status = np.zeros(3, np.int64)
        send = [np.zeros((cap + 1, REC), np.float32), np.zeros((cap + 1, REC), np.float32)]
        holes = [np.zeros(cap, np.int64), np.zeros(cap, np.int64)]
        for side, dest in ((0, left), (1, right)):
            if side == 1 and right == left:
                break
            rows = np.nonzero(own == dest)[0] if dest != self.rank else np.zeros(0, np.int64)
            status[side] = rows.size
            rows = rows[:cap]
            send[side][1:1 + rows.size, 0:3] = _np(pos)[rows]
            send[side][1:1 + rows.size, 3:6] = _np(vel)[rows]
            send[side][1:1 + rows.size, 6:8] = _np(ids)[rows].astype(np.int64).view(np.float32).reshape(-1, 2)
            holes[side][:rows.size] = rows
        status[2] = int(((own != self.rank) & (own != left) & (own != right)).sum())
        for side in (0, 1):
            send[side][0, 0:2] = np.array([status[side]], np.int64).view(np.float32)
        return (torch.from_numpy(send[0]), torch.from_numpy(send[1]), torch.from_numpy(holes[0]),
                torch.from_numpy(holes[1]), torch.from_numpy(status))

    def _owner(self, pos):
        i = (_np(pos)[:, 0] * np.float32(self.N)).astype(np.int64)
        return np.clip(i // self.nxl, 0, self.P - 1)

    def count_owners(self, pos):
        return torch.from_numpy(np.bincount(self._owner(pos), minlength=self.P).astype(np.int64))

    def pack_leavers(self, pos, vel, ids, offsets, nout):
        own = self._owner(pos)
        rows = np.nonzero(own != self.rank)[0]
        rows = rows[np.argsort(own[rows], kind="stable")]
        assert rows.size == nout
        rec = np.empty((nout, REC), dtype=np.float32)
        rec[:, 0:3] = _np(pos)[rows]
        rec[:, 3:6] = _np(vel)[rows]
        rec[:, 6:8] = _np(ids)[rows].astype(np.int64).view(np.float32).reshape(-1, 2)
        return torch.from_numpy(rec), torch.from_numpy(rows.astype(np.int64))

    def unpack_rows(self, recvbuf, rows, pos, vel, ids):
        rec = _np(recvbuf)
        r = _np(rows)
        _np(pos)[r] = rec[:, 0:3]
        _np(vel)[r] = rec[:, 3:6]
        _np(ids)[r] = np.ascontiguousarray(rec[:, 6:8]).view(np.int64).reshape(-1)

    def move_rows(self, src, dst, pos, vel, ids):
        s, d = _np(src), _np(dst)
        for a in (pos, vel, ids):
            _np(a)[d] = _np(a)[s]

    def max_abs(self, x):
        return torch.tensor([float(np.abs(_np(x)).max()) if x.numel() else 0.0], dtype=torch.float32)

    def morton_order(self, pos):
        keys = oracle.morton.positions_to_keys(np.ascontiguousarray(_np(pos)))
        return torch.from_numpy(np.argsort(keys, kind="stable").astype(np.int64))

    def gather_rows(self, idx, a, out=None):
        r = a[idx]
        if out is None:
            return r
        out.copy_(r)
        return out

    # -- particles <-> mesh (oracle on the full periodic grid, cut to the slab + ghosts)
    def bin(self, pos):
        self._binned = np.ascontiguousarray(_np(pos))
        own = self._owner(pos)
        assert (own == self.rank).all(), "binning particles that are not in the slab"
        return pos.shape[0]

    # the bin-ordered layout: a stable sort by (bin, micro-block) key -- the CUDA sort is not stable inside a key,
    # which no consumer depends on
    sorted_layout = True

    def sort_by_bin(self, pos, vel, ids, pos_out, vel_out, ids_out, src_rows=None):
        own = self._owner(pos)
        assert (own == self.rank).all(), "sorting particles that are not in the slab"
        N = self.N
        c = np.minimum((_np(pos) * np.float32(N)).astype(np.int64), N - 1)
        NB = N // 8
        i = c[:, 0] - self.x0
        b = ((i >> 3) * NB + (c[:, 1] >> 3)) * NB + (c[:, 2] >> 3)
        ii, jj, kk = (i >> 1) & 3, (c[:, 1] >> 1) & 3, (c[:, 2] >> 1) & 3
        mb = ((ii >> 1) << 5) | ((jj >> 1) << 4) | ((kk >> 1) << 3) | ((ii & 1) << 2) | ((jj & 1) << 1) | (kk & 1)
        order = torch.from_numpy(np.argsort(b * 64 + mb, kind="stable"))
        pos_out.copy_(pos[order])
        vel_out.copy_(vel[order])
        ids_out.copy_(ids[order])
        return pos.shape[0]

    def deposit_sorted(self, pos, scheme):
        self._binned = np.ascontiguousarray(_np(pos))
        return self.deposit(pos.shape[0], scheme)

    def interp_kick_phi_sorted(self, phi_g, ghost, order, pos, vel, acc, scheme, half_dt, u_g=None, f=0.0, fr_n=0):
        self._binned = np.ascontiguousarray(_np(pos))
        return self.interp_kick_phi(phi_g, ghost, order, pos.shape[0], vel, acc, scheme, half_dt, u_g, f, fr_n)

    def _planes(self, ghost):
        return np.arange(self.x0 - ghost, self.x0 + self.nxl + ghost) % self.N

    def deposit(self, binned, scheme):
        fn = {1: oracle.mesh.CIC, 2: oracle.mesh.TSC_seq}[scheme]
        full = fn(self._binned, self.N) if binned else np.zeros((self.N,) * 3, np.float32)
        if self.P == 1:
            # ghosts of a periodic slab: what fell on the wrapped planes is counted in the ghosts only
            out = np.zeros((self.nxl + 2, self.N, self.N), np.float32)
            out[1:-1] = full
            return torch.from_numpy(out)
        # particles of this slab only touch planes x0-1 .. x0+nxl, each exactly once when P >= 2 and nxl >= 2
        return torch.from_numpy(np.ascontiguousarray(full[self._planes(1)]))

    def affine(self, x, f1, f2):
        a = _np(x)
        oracle.utils.linear_operator_inplace(a.reshape(-1), np.float32(f1), np.float32(f2))

    def interp_kick_phi(self, phi_g, ghost, order, binned, vel, acc, scheme, half_dt, u_g=None, f=0.0, fr_n=0):
        full = np.zeros((self.N,) * 3, np.float32)
        full[self._planes(ghost)] = _np(phi_g)
        if u_g is not None:
            full_u = np.zeros((self.N,) * 3, np.float32)
            full_u[self._planes(ghost)] = _np(u_g)
            force = oracle.mesh.derivative_fR(full, full_u, np.float32(f), fr_n, order)
        else:
            force = oracle.mesh.derivative(full, order)
        fn = {1: oracle.mesh.invCIC_vec, 2: oracle.mesh.invTSC_vec}[scheme]
        a = fn(force, self._binned) if binned else np.zeros((0, 3), np.float32)
        _np(acc)[:] = a
        mv = 0.0
        if vel is not None:
            v = _np(vel)
            oracle.utils.add_vector_scalar_inplace(v, a, -np.float32(half_dt))
            mv = float(np.abs(v).max()) if v.size else 0.0
        return torch.tensor([float(np.abs(a).max()) if a.size else 0.0, mv], dtype=torch.float32)

    # -- transposed FFT
    def spectrum_buffer(self):
        return torch.empty((self.nxl * self.N * (self.N // 2 + 1), 2), dtype=torch.float32)

    def _c(self, t, shape):
        return _np(t).reshape(-1).view(np.complex64).reshape(shape)

    def fft2d_r2c(self, planes, spec2d):
        self._c(spec2d, (self.nxl, self.N, self.N // 2 + 1))[:] = np.fft.rfftn(_np(planes), axes=(1, 2))

    def fft2d_c2r(self, spec2d, planes):
        s = self._c(spec2d, (self.nxl, self.N, self.N // 2 + 1))
        # unnormalised, like cuFFT
        _np(planes)[:] = np.fft.irfftn(s, s=(self.N, self.N), axes=(1, 2)) * (self.N * self.N)

    def yblocks(self, src, dst, to_blocks):
        nz = self.N // 2 + 1
        if to_blocks:
            a = self._c(src, (self.nxl, self.P, self.nyl, nz))
            self._c(dst, (self.P, self.nxl, self.nyl, nz))[:] = a.transpose(1, 0, 2, 3)
        else:
            a = self._c(src, (self.P, self.nxl, self.nyl, nz))
            self._c(dst, (self.nxl, self.P, self.nyl, nz))[:] = a.transpose(1, 0, 2, 3)

    def fft_x(self, spec_t, inverse):
        s = self._c(spec_t, (self.N, self.nyl, self.N // 2 + 1))
        s[:] = np.fft.ifft(s, axis=0) * self.N if inverse else np.fft.fft(s, axis=0)

    def pk_bins(self, spec_t, p):
        """fourier.fourier_grid_to_Pk (fourier.py:22-100) on this rank's y-block of the transposed spectrum: per-bin
        sums of |k|, |delta_k W^-p|^2 and mode counts, [3][N] float64 (what psc_pk_slab returns)"""
        N, nz = self.N, self.N // 2 + 1
        s = self._c(spec_t, (N, self.nyl, nz))
        fold = lambda i: np.where(i >= N // 2, i - N, i).astype(np.float32)   # noqa: E731
        kx = fold(np.arange(N))[:, None, None]
        ky = fold(np.arange(self.y0, self.y0 + self.nyl))[None, :, None]
        kz = np.arange(nz, dtype=np.float32)[None, None, :]
        w = (np.sinc(kx / np.float32(N)) * np.sinc(ky / np.float32(N)) * np.sinc(kz / np.float32(N))).astype(np.float64)
        d2 = (np.abs(s.astype(np.complex128)) ** 2) * w ** (-2.0 * p)
        knorm = np.sqrt(kx * kx + ky * ky + kz * kz).astype(np.float32)
        ki = (knorm + np.float32(0.5)).astype(np.int64)
        keep = np.ones(ki.shape, bool)
        if self.y0 == 0:
            keep[0, 0, 0] = False    # the DC mode is skipped
        bins = np.stack([np.bincount(ki[keep], weights=a[keep].astype(np.float64), minlength=N)[:N]
                         for a in (knorm, d2, np.ones_like(d2))])
        return torch.from_numpy(bins)

    def green(self, spec_t, kind, p, scale):
        ones = np.ones((self.N, self.N, self.N // 2 + 1), dtype=np.complex64)
        if kind == 0:
            oapi.fourier.inverse_laplacian(ones)
        elif kind == 1:
            oapi.fourier.inverse_laplacian_compensated(ones, p)
        else:
            oapi.fourier.inverse_laplacian_7pt(ones)
        s = self._c(spec_t, (self.N, self.nyl, self.N // 2 + 1))
        s *= ones[:, self.y0:self.y0 + self.nyl, :] * np.float32(scale)

    # -- multigrid on the slab: the per-cell code of csrc/slab_mg.cu through the host harness; gathered coarse levels
    #    through the oracle's single-domain V-cycle
    def mg_gs_colour(self, xg, b, nxl, n, x0, colour, f_relax):
        import ctypes as C
        _harness().hx_gs_colour(_fp(xg), _fp(b), nxl, n, x0, colour, C.c_float(float(f_relax)))

    def mg_operator(self, xg, nxl, n):
        out = torch.empty((nxl, n, n), dtype=torch.float32)
        _harness().hx_operator(_fp(xg), nxl, n, _fp(out))
        return out

    def mg_restrict_residual(self, xg, b, nxl, n):
        out = torch.empty((nxl // 2, n // 2, n // 2), dtype=torch.float32)
        _harness().hx_restrict_residual(_fp(xg), _fp(b), nxl, n, _fp(out))
        return out

    def mg_restriction(self, fine, nxl, n, sign, out=None):
        import ctypes as C
        if out is None:
            out = torch.empty((nxl // 2, n // 2, n // 2), dtype=torch.float32)
        _harness().hx_restriction(_fp(fine), nxl, n, C.c_float(float(sign)), _fp(out))
        return out

    def mg_add_prolongation(self, fine_g, coarse_g, nxlc, nc):
        _harness().hx_add_prolongation(_fp(fine_g), _fp(coarse_g), nxlc, nc)

    def mg_diff_sumsq(self, a, fa, b):
        d = np.float32(fa) * _np(a).astype(np.float32) - _np(b)
        return torch.tensor([float(np.sum(d.astype(np.float64) ** 2))], dtype=torch.float64)

    def mg_gs_colour_fr(self, xg, b, rhs, q, nxl, n, x0, colour, f_relax, kind):
        import ctypes as C
        _harness().hx_gs_colour_fr(_fp(xg), _fp(b), _fp(rhs) if rhs is not None else None, C.c_float(float(q)), nxl, n,
                                   x0, colour, C.c_float(float(f_relax)), kind)

    def mg_operator_fr(self, xg, b, q, nxl, n, kind):
        import ctypes as C
        out = torch.empty((nxl, n, n), dtype=torch.float32)
        _harness().hx_operator_fr(_fp(xg), _fp(b), C.c_float(float(q)), nxl, n, kind, _fp(out))
        return out

    def mg_init_fr(self, b, q, nxl, n, kind, out):
        import ctypes as C
        _harness().hx_init_fr(_fp(b), C.c_float(float(q)), nxl, n, kind, _fp(out))

    def lincomb(self, x, f1, y, f2):
        oracle.utils.linear_operator_vectors_inplace(_np(x).reshape(-1), np.float32(f1), _np(y).reshape(-1),
                                                     np.float32(f2))

    def axpy(self, y, x, a):
        oracle.utils.add_vector_scalar_inplace(_np(y).reshape(-1, 1), _np(x).reshape(-1, 1), np.float32(a))

    def mg_cube_solve_fas(self, x_c, b_c, res_c, param, nlevel, coarsest):
        from oracle import host
        x_c, b_c, res_c = (np.ascontiguousarray(_np(t)) for t in (x_c, b_c, res_c))
        q = np.float32(param["fR_q"])
        L_c = host._fr_mod(param).operator(x_c, b_c, q)
        oracle.utils.linear_operator_vectors_inplace(res_c, np.float32(4), L_c, np.float32(1))
        corr = x_c.copy()
        if coarsest:
            host._fas_smooth(corr, b_c, param["Npre"], param, res_c)
        else:
            host._cycle_FAS("V", corr, b_c, param, nlevel + 1, res_c)
        oracle.utils.add_vector_scalar_inplace(corr, x_c, np.float32(-1))
        return torch.from_numpy(corr)

    def mond_rhs(self, phig, out, nxl, n, g0, fn, alpha):
        import ctypes as C
        _harness().hx_mond_rhs(_fp(phig), _fp(out), nxl, n, C.c_float(float(np.float32(g0))), int(fn),
                               C.c_float(float(alpha)))

    def mg_cube_solve(self, res_cube, param, nlevel, coarsest):
        from oracle import host
        res = np.ascontiguousarray(_np(res_cube))
        x = oracle.laplacian.initialise_potential(res)
        if coarsest:
            oracle.laplacian.smoothing(x, res, param["Npre"])
        else:
            host._cycle("V", x, res, param, nlevel + 1)
        return torch.from_numpy(x)

    def close(self):
        pass
