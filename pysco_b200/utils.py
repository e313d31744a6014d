"""Mirror of the hot-path part of pysco/utils.py."""
import numpy as np
import torch

from . import _lib

# astropy.constants values the reference reads (utils.py:11): CODATA-2018 / IAU-2015
_PC = 3.0856775814913673e16
_G = 6.6743e-11


def set_units(param) -> None:
    """utils.py:167-196 (host scalars)."""
    mpc_to_km = 1e3 * _PC
    g = _G * 1e-9
    H0 = param["H0"] / mpc_to_km
    rhoc = 3.0 * H0 ** 2 / (8.0 * np.pi * g)
    param["unit_l"] = param["aexp"] * param["boxlen"] * 100.0 / H0
    param["unit_t"] = param["aexp"] ** 2 / H0
    param["unit_d"] = param["Om_m"] * rhoc / param["aexp"] ** 3
    param["mpart"] = param["unit_d"] * param["unit_l"] ** 3 / param["npart"]


def max_abs(x):
    """utils.py:220-240 -> np.float32 (synchronises)."""
    c = _lib.Ctx()
    t = c.dev(x)
    out = _lib.zeros((1,))
    _lib.check(_lib.load().psc_max_abs(_lib.ptr(t), t.numel(), _lib.ptr(out), _lib.stream()))
    return np.float32(out.item())


def add_vector_scalar_inplace(y, x, a) -> None:
    """utils.py:264-297: y += a*x.  The arithmetic type follows the scalar's type as in Numba:
    np.float32 -> float32 FMA; Python float / np.float64 -> float64 intermediate."""
    c = _lib.Ctx()
    ty, tx = c.dev(y, inplace=True), c.dev(x)
    is64 = 0 if isinstance(a, np.float32) else 1
    _lib.check(_lib.load().psc_axpy(_lib.ptr(ty), _lib.ptr(tx), float(a), is64, ty.numel(), _lib.stream()))
    c.finish()


def prod_vector_scalar_inplace(y, a) -> None:
    """utils.py:410-430: y *= a"""
    c = _lib.Ctx()
    ty = c.dev(y, inplace=True)
    _lib.check(_lib.load().psc_linear_operator(_lib.ptr(ty), float(np.float32(a)), 0.0, _lib.ptr(ty),
                                               ty.numel(), _lib.stream()))
    c.finish()


def linear_operator(x, f1, f2):
    """utils.py:644-680: f1*x + f2 (new array)"""
    c = _lib.Ctx()
    tx = c.dev(x)
    out = torch.empty_like(tx)
    _lib.check(_lib.load().psc_linear_operator(_lib.ptr(tx), float(np.float32(f1)), float(np.float32(f2)),
                                               _lib.ptr(out), tx.numel(), _lib.stream()))
    return c.ret(out)


def linear_operator_inplace(x, f1, f2) -> None:
    """utils.py:684-717"""
    c = _lib.Ctx()
    tx = c.dev(x, inplace=True)
    _lib.check(_lib.load().psc_linear_operator(_lib.ptr(tx), float(np.float32(f1)), float(np.float32(f2)),
                                               _lib.ptr(tx), tx.numel(), _lib.stream()))
    c.finish()


def linear_operator_vectors_inplace(x, f1, y, f2) -> None:
    """utils.py:721-755: x = f1*x + f2*y"""
    c = _lib.Ctx()
    tx, ty = c.dev(x, inplace=True), c.dev(y)
    _lib.check(_lib.load().psc_lincomb(_lib.ptr(tx), float(np.float32(f1)), _lib.ptr(ty),
                                       float(np.float32(f2)), tx.numel(), _lib.stream()))
    c.finish()


def periodic_wrap(position) -> None:
    """utils.py:1120-1149"""
    c = _lib.Ctx()
    t = c.dev(position, inplace=True)
    _lib.check(_lib.load().psc_periodic_wrap(_lib.ptr(t), t.numel(), _lib.stream()))
    c.finish()


def injection_with_indices(idx, a):
    """utils.py:894-924: out[i] = a[idx[i]] (rows of [Np,3])"""
    c = _lib.Ctx()
    ti, ta = c.dev(idx, torch.int64), c.dev(a)
    out = torch.empty_like(ta)
    _lib.check(_lib.load().psc_gather3(_lib.ptr(ti), _lib.ptr(ta), _lib.ptr(out), ta.shape[0], _lib.stream()))
    return c.ret(out)


def argsort_keys(keys):
    """Stable argsort of Morton keys on the device (= np.argsort(keys, kind='stable'))."""
    lib = _lib.load()
    n = keys.shape[0]
    idx = _lib.empty((n,), torch.int64)
    nbytes = int(lib.psc_argsort_workspace_bytes(n))
    scratch = _lib.empty((nbytes,), torch.uint8)
    _lib.check(lib.psc_argsort_keys(_lib.ptr(keys), n, _lib.ptr(idx), _lib.ptr(scratch), nbytes, _lib.stream()))
    return idx


# ------------------------------------------------------------------ particle order of the device-resident time loop
# integration.leapfrog keeps device-resident particle arrays in BIN order and re-sorts them every step (csrc/binned.cu:
# psc_step_sort); the particle in row n of such arrays is the one the reference has in row ids[n].  The ids travel in
# this registry, keyed by the tensors leapfrog returned; anything that needs the reference's row order (snapshots,
# row-wise comparisons) calls reference_order().  Arrays nobody registered are in the reference's order already.
import weakref

_order_ids = {}


def set_particle_ids(tensors, ids) -> None:
    dead = [k for k, (ref, _) in _order_ids.items() if ref() is None]
    for k in dead:
        del _order_ids[k]
    for t in tensors:
        if isinstance(t, torch.Tensor):
            _order_ids[id(t)] = (weakref.ref(t), ids)


def particle_ids(t):
    """int32 device tensor: row n of `t` is the particle of the reference's row ids[n]; None = reference order"""
    e = _order_ids.get(id(t)) if isinstance(t, torch.Tensor) else None
    return e[1] if e is not None and e[0]() is t else None


def reference_order(*arrays):
    """The given [np, 3] arrays (position, velocity, acceleration of the device-resident time loop) in the reference's
    particle order: out[ids[n]] = in[n].  Arrays that already are in that order come back unchanged."""
    out = []
    for a in arrays:
        ids = particle_ids(a)
        if ids is None:
            out.append(a)
            continue
        b = torch.empty_like(a)
        _lib.check(_lib.load().psc_scatter3_by_id(_lib.ptr(ids), _lib.ptr(a), _lib.ptr(b), a.shape[0], _lib.stream()))
        out.append(b)
    return out[0] if len(out) == 1 else tuple(out)


def _morton_relabel(position, velocity, acceleration):
    """reorder_particles for the bin-ordered arrays of the device-resident time loop: the arrays stay where they are
    (the step re-sorts them into bins anyway) and only their ids change -- ids[n] becomes the row the particle has in
    the reference after its Morton reorder (csrc/binned.cu psc_morton_ids_sorted: a 512-key radix sort per bin in shared
    memory, ~1 ms at 512^3 against 34 ms for the global sort and the gathers).  Returns the SAME tensors, or None when
    the arrays are not the bin-ordered output of the last step (then the caller sorts globally)."""
    import os
    from . import mesh
    if not isinstance(position, torch.Tensor) or particle_ids(position) is None or os.environ.get("PSC_NO_RELABEL"):
        return None
    sb = mesh.sorted_bins_of(position)
    if sb is None or sb.N & (sb.N - 1):
        return None
    n = position.shape[0]
    ids = _lib.empty((n,), torch.int32)
    flag = _lib.empty((1,), torch.int32)
    _lib.check(_lib.load().psc_morton_ids_sorted(_lib.ptr(position), _lib.ptr(sb.scratch), sb.scratch.numel(), sb.table,
                                                 n, sb.N, _lib.ptr(ids), _lib.ptr(flag), _lib.stream()))
    sb.predicted = None       # the relabelling used the count table as scratch
    if int(flag.item()):      # a bin beyond the in-kernel sort's capacity (2048 particles): global sort
        return None
    group = [t for t in (position, velocity, acceleration) if t is not None]
    set_particle_ids(group, ids)
    return group[0] if len(group) == 1 else tuple(group)


def reorder_particles(position, velocity=None, acceleration=None):
    """utils.py:1019-1075: Morton keys -> (global, stable) argsort -> gathers; returns NEW arrays.
    Matches the reference's nthreads == 1 path (np.argsort); rows of equal key keep their order.
    Bin-ordered device arrays (integration.leapfrog's) are relabelled instead of moved: see _morton_relabel."""
    relabelled = _morton_relabel(position, velocity, acceleration)
    if relabelled is not None:
        return relabelled
    c = _lib.Ctx()
    pos = c.dev(position)
    vel = c.dev(velocity)
    acc = c.dev(acceleration)
    lib = _lib.load()
    n = pos.shape[0]
    keys = _lib.empty((n,), torch.int64)
    _lib.check(lib.psc_morton_keys(_lib.ptr(pos), n, _lib.ptr(keys), _lib.stream()))
    idx = argsort_keys(keys)
    del keys

    def g(a):
        out = torch.empty_like(a)
        _lib.check(lib.psc_gather3(_lib.ptr(idx), _lib.ptr(a), _lib.ptr(out), n, _lib.stream()))
        return out

    if acc is not None:
        return c.ret(g(pos)), c.ret(g(vel)), c.ret(g(acc))
    if vel is not None:
        return c.ret(g(pos)), c.ret(g(vel))
    return c.ret(g(pos))
