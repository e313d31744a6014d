"""Mirror of pysco/mesh.py: mass assignment, inverse interpolation, finite-difference gradients,
restriction / prolongation.  Same names and argument meaning as the reference."""
import numpy as np
import torch

from . import _lib


class Binned:
    """Per-step shadow binning of the particles into 8^3-cell bins (csrc/binned.cu): per-bin first record / fill, a
    binned copy of the positions and the source row of every binned particle, in one device scratch buffer.
    ready: the scratch holds the fill of a finished binning, which the next step's kick+drift+wrap can use to drop its
    records straight into the bins (direct scatter, psc_kick_drift_wrap_count mode 1)."""

    def __init__(self, scratch, np_, N):
        self.scratch, self.np, self.N = scratch, np_, N
        self.ready = False
        self.mode = 0


def can_bin(N, np_):
    return N >= 8 and N % 8 == 0 and 0 < np_ < 2 ** 31


def alloc_binned(np_, ncells_1d):
    """Scratch of a binning whose counts are produced elsewhere (psc_kick_drift_wrap_count)."""
    N = int(ncells_1d)
    nbytes = int(_lib.load().psc_bin_workspace_bytes(int(np_), N))
    return Binned(_lib.empty((nbytes,), torch.uint8), int(np_), N)


class SortedBins:
    """Bin table of particle arrays that are themselves stored in bin order (psc_step_sort): first row / count of
    every 8^3-cell bin and the list of bins split into parts; no binned copy of the positions."""

    def __init__(self, scratch, np_, N):
        self.scratch, self.np, self.N = scratch, np_, N
        # the scratch holds two tables: `table` describes the arrays the last step_sort returned (`owner`, a weak
        # reference to its position tensor); the next sort reads that table and writes the other one
        self.table, self.owner = 0, None
        # speculative count pass (csrc/binned.cu BinPredict): predict_next = (half_dt, dt, dt_is_f64) asks the next
        # interp_kick_phi to count the bins of the next step under that time step; predicted = what it counted for
        self.predict_next = None
        self.predicted = None
        self.miss_streak = 0      # consecutive predictions the host did not use (another criterion set the step)
        self.rest = 0             # steps left without prediction after a streak of misses
        self.counts_skipped = 0   # sorts that used a prediction

    def describes(self, pos):
        return self.owner is not None and self.owner() is pos


_step_sorted = {}


def step_sorted(np_, ncells_1d):
    """persistent SortedBins scratch of the time loop for (device, N, np)"""
    key = (torch.cuda.current_device(), int(ncells_1d), int(np_))
    b = _step_sorted.get(key)
    if b is None:
        _step_sorted.clear()
        nbytes = int(_lib.load().psc_sorted_workspace_bytes(int(np_), int(ncells_1d)))
        b = _step_sorted[key] = SortedBins(_lib.empty((nbytes,), torch.uint8), int(np_), int(ncells_1d))
    return b


def _stamp(*tensors):
    """identity and in-place version of device tensors (the library's own kernels write through raw pointers and do not
    bump versions; anything the caller does through torch does)"""
    return tuple((id(t), t._version) for t in tensors)


def sorted_bins_of(pos):
    """the SortedBins whose current table describes the bin-ordered tensor `pos`, or None"""
    for sb in _step_sorted.values():
        if sb.describes(pos):
            return sb
    return None


def step_sort(pos, vel, acc, ids, half_dt, dt, dt_is_f64, sb):
    """integration.py:250-258 (first half-kick, drift, wrap) fused with the re-sort of the particle arrays into bin
    order: returns NEW (position, velocity, ids) in bin order; `sb` then describes the bins of these arrays.

    When `pos` is the bin-ordered array the previous call returned (sb.describes(pos)), every CTA sorts the particles
    of one source bin in shared memory (csrc/binned.cu step_sort_local_kernel); otherwise -- first step, arrays
    reordered by the caller, PSC_NO_LOCAL_SORT=1 -- the sort goes through one global atomic per particle."""
    import os
    import weakref
    n = pos.shape[0]
    pos2, vel2 = torch.empty_like(pos), torch.empty_like(vel)
    ids2 = _lib.empty((n,), torch.int32)
    src = sb.table if (sb.describes(pos) and not os.environ.get("PSC_NO_LOCAL_SORT")) else -1
    # the count pass was done by the previous interpolation kernel if it predicted exactly this time step for exactly
    # these arrays (same bits: float(half_dt), float(dt) and the float64 flag are what both calls hand to the library)
    ready = int(src >= 0 and sb.predicted == (float(half_dt), float(dt), int(dt_is_f64), _stamp(pos, vel, acc)))
    if sb.predicted is not None:
        sb.miss_streak = 0 if ready else sb.miss_streak + 1
    sb.predicted = None
    _lib.check(_lib.load().psc_step_sort(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(acc), _lib.ptr(ids), n, float(half_dt),
                                         float(dt), int(dt_is_f64), sb.N, src, ready, _lib.ptr(sb.scratch),
                                         sb.scratch.numel(), _lib.ptr(pos2), _lib.ptr(vel2), _lib.ptr(ids2),
                                         _lib.stream()))
    sb.counts_skipped += ready
    sb.table = 0 if src < 0 else 1 - src
    sb.owner = weakref.ref(pos2)
    return pos2, vel2, ids2


_step_binned = {}


def step_binned(np_, ncells_1d):
    """The persistent Binned of the time loop for (device, N, np): kept from step to step so that the fill of one
    step's binning sizes the bins of the next step's direct scatter.  PSC_NO_DIRECT_SCATTER=1 restores the count ->
    scan -> scatter binning of every step."""
    import os
    key = (torch.cuda.current_device(), int(ncells_1d), int(np_))
    b = _step_binned.get(key)
    if b is None:
        _step_binned.clear()          # one time loop at a time: do not hold scratch of finished runs
        b = _step_binned[key] = alloc_binned(np_, ncells_1d)
    if os.environ.get("PSC_NO_DIRECT_SCATTER"):
        b.ready = False
    return b


def kick_drift_wrap_count(pos, vel, acc, half_dt, dt, dt_is_f64, binned, zero_counts=True, row0=0):
    """integration.py:250-258 on (a chunk of) the particles + binning of the new positions into `binned`: per-bin counts
    (binned.ready False) or the records themselves (direct scatter, binned.ready True)."""
    if zero_counts:
        binned.mode = 1 if binned.ready else 0
    _lib.check(_lib.load().psc_kick_drift_wrap_count(
        _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(acc), pos.shape[0], float(half_dt), float(dt), int(dt_is_f64),
        binned.N, binned.np, _lib.ptr(binned.scratch), binned.scratch.numel(), int(zero_counts), binned.mode, int(row0),
        _lib.stream()))


def finish_binning(position, binned):
    """the rest of the binning started by kick_drift_wrap_count: scan + scatter from the counts (mode 0), or the
    heavy-bin list and the on-device overflow fall-back of the direct scatter (mode 1)"""
    _lib.check(_lib.load().psc_bin_particles_counted(_lib.ptr(position), position.shape[0], binned.N,
                                                     _lib.ptr(binned.scratch), binned.scratch.numel(), binned.mode,
                                                     _lib.stream()))
    binned.ready = True
    return binned


def bin_particles(position, ncells_1d):
    """Bin a device position array; returns a Binned handle for deposit_rhs / interp_kick."""
    pos = _lib.Ctx().dev(position)
    N, n = int(ncells_1d), pos.shape[0]
    lib = _lib.load()
    nbytes = int(lib.psc_bin_workspace_bytes(n, N))
    scratch = _lib.empty((nbytes,), torch.uint8)
    _lib.check(lib.psc_bin_particles(_lib.ptr(pos), n, N, _lib.ptr(scratch), nbytes, _lib.stream()))
    return Binned(scratch, n, N)


def _deposit(position, ncells_1d, scheme, scale=1.0, f1=1.0, f2=0.0, binned=None):
    c = _lib.Ctx()
    pos = c.dev(position)
    N = int(ncells_1d)
    rho = _lib.empty((N, N, N))
    lib = _lib.load()
    if binned is None and can_bin(N, pos.shape[0]):
        binned = bin_particles(pos, N)
    if isinstance(binned, SortedBins):
        _lib.check(lib.psc_deposit_sorted(_lib.ptr(pos), _lib.ptr(binned.scratch), binned.scratch.numel(), binned.table,
                                          binned.np, N, scheme, float(scale), float(f1), float(f2), _lib.ptr(rho),
                                          _lib.stream()))
    elif binned is not None:
        _lib.check(lib.psc_deposit_binned(_lib.ptr(binned.scratch), binned.scratch.numel(), binned.np, N, scheme,
                                          float(scale), float(f1), float(f2), _lib.ptr(rho), _lib.stream()))
    else:
        _lib.check(lib.psc_deposit(_lib.ptr(pos), pos.shape[0], N, scheme, float(scale), float(f1),
                                   float(f2), _lib.ptr(rho), _lib.stream()))
    return c.ret(rho)


def NGP(position, ncells_1d):
    """mesh.py:2240-2278"""
    return _deposit(position, ncells_1d, _lib.NGP)


def CIC(position, ncells_1d):
    """mesh.py:2284-2358"""
    return _deposit(position, ncells_1d, _lib.CIC)


def TSC(position, ncells_1d):
    """mesh.py:2468-2595"""
    return _deposit(position, ncells_1d, _lib.TSC)


TSC_seq = TSC  # mesh.py:2363-2462: same result, the reference's sequential variant


def deposit_rhs(position, ncells_1d, scheme, scale, f1, f2, binned=None):
    """Fused mass assignment + density rescale + Poisson right-hand side (solver.py:80-116, 444-449):
    f1 * (scale * deposit) + f2.  binned: a Binned handle of the same positions (shared with interp_kick)."""
    return _deposit(position, ncells_1d, scheme, scale, f1, f2, binned)


def _interp(grid, position, scheme):
    c = _lib.Ctx()
    g, pos = c.dev(grid), c.dev(position)
    N = g.shape[0]
    ncomp = 3 if g.dim() == 4 else 1
    n = pos.shape[0]
    out = _lib.empty((n, 3) if ncomp == 3 else (n,))
    _lib.check(_lib.load().psc_interp(_lib.ptr(g), _lib.ptr(pos), n, N, ncomp, scheme, _lib.ptr(out),
                                      _lib.stream()))
    return c.ret(out)


def invNGP(grid, position):
    return _interp(grid, position, _lib.NGP)


def invCIC(grid, position):
    return _interp(grid, position, _lib.CIC)


def invTSC(grid, position):
    return _interp(grid, position, _lib.TSC)


invNGP_vec, invCIC_vec, invTSC_vec = invNGP, invCIC, invTSC  # mesh.py:2627-3088 (AoS [N,N,N,3] grid)


def interp_kick(force, position, velocity, scheme, half_dt, binned=None):
    """inv{CIC,TSC}_vec fused with v -= half_dt*a and the max|a|, max|v| reductions.
    Returns (acceleration, maxima[2] device tensor).  velocity may be None (plain interpolation).
    force is AoS [N,N,N,3] (reference layout) or float4-padded [N,N,N,4] (derivative(..., padded=True))."""
    c = _lib.Ctx()
    g, pos = c.dev(force), c.dev(position)
    vel = c.dev(velocity, inplace=True)
    n = pos.shape[0]
    acc = _lib.empty((n, 3))
    mx = _lib.zeros((2,))
    if binned is not None and g.shape[-1] == 4 and scheme != _lib.NGP:
        _lib.check(_lib.load().psc_interp_kick4_binned(
            _lib.ptr(g), _lib.ptr(binned.scratch), binned.scratch.numel(), _lib.ptr(vel), _lib.ptr(acc), n,
            g.shape[0], scheme, float(half_dt), _lib.ptr(mx), _lib.stream()))
    else:
        fn = _lib.load().psc_interp_kick4 if g.shape[-1] == 4 else _lib.load().psc_interp_kick
        _lib.check(fn(_lib.ptr(g), _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(acc), n,
                      g.shape[0], scheme, float(half_dt), _lib.ptr(mx), _lib.stream()))
    c.finish()
    return c.ret(acc), mx


def interp_kick_phi(potential, u, f, fr_n, order, position, velocity, scheme, half_dt, binned):
    """derivative[_fR](potential[, u]) + inv{CIC,TSC}_vec + half-kick + max reductions in one kernel per bin:
    the force tile of every 8^3 bin is formed in shared memory from the potential (no force grid in HBM).
    Returns (acceleration, maxima[2] device tensor)."""
    if fr_n not in (0, 1, 2):
        raise NotImplementedError(f"Unsupported: fR_n={fr_n}")
    if order not in (2, 3, 5, 7):
        raise NotImplementedError(f"Unsupported: gradient_order={order}")
    c = _lib.Ctx()
    phi, tu, pos = c.dev(potential), c.dev(u), c.dev(position)
    vel = c.dev(velocity, inplace=True)
    n = pos.shape[0]
    acc = _lib.empty((n, 3))
    mx = _lib.zeros((2,))
    if isinstance(binned, SortedBins):
        nxt, binned.predict_next, binned.predicted = binned.predict_next, None, None
        predict = nxt is not None and vel is not None and binned.describes(pos)
        h2, d2, f2 = nxt if predict else (0.0, 0.0, 0)
        _lib.check(_lib.load().psc_interp_kick_phi_sorted(
            _lib.ptr(phi), _lib.ptr(tu), float(np.float32(f)), fr_n, order, _lib.ptr(pos), _lib.ptr(binned.scratch),
            binned.scratch.numel(), binned.table, _lib.ptr(vel), _lib.ptr(acc), n, phi.shape[0], scheme,
            float(half_dt), _lib.ptr(mx), int(predict), float(h2), float(d2), int(f2), _lib.stream()))
        if predict:
            # the guess holds for exactly these tensors in exactly this state: an in-place torch operation of the
            # caller on any of them before the next step bumps its version and voids the prediction
            binned.predicted = (float(h2), float(d2), int(f2), _stamp(pos, vel, acc))
    else:
        _lib.check(_lib.load().psc_interp_kick_phi_binned(
            _lib.ptr(phi), _lib.ptr(tu), float(np.float32(f)), fr_n, order, _lib.ptr(binned.scratch),
            binned.scratch.numel(), _lib.ptr(vel), _lib.ptr(acc), n, phi.shape[0], scheme, float(half_dt), _lib.ptr(mx),
            _lib.stream()))
    c.finish()
    return c.ret(acc), mx


def _gradient(a, b, f, fr_n, order, add, force=None, padded=False):
    if fr_n not in (0, 1, 2):
        raise NotImplementedError(f"Unsupported: fR_n={fr_n}")
    if order not in (2, 3, 5, 7):
        raise NotImplementedError(f"Unsupported: gradient_order={order}")
    c = _lib.Ctx()
    ta, tb = c.dev(a), c.dev(b)
    N = (ta if ta is not None else tb).shape[0]
    out = c.dev(force, inplace=True) if add else _lib.empty((N, N, N, 4 if padded else 3))
    _lib.check(_lib.load().psc_gradient(_lib.ptr(ta), _lib.ptr(tb), float(np.float32(f)), fr_n, order,
                                        1 if add else 0, N, _lib.ptr(out), out.shape[-1], _lib.stream()))
    c.finish()
    return None if add else c.ret(out)


def derivative(a, gradient_order, padded=False):
    """mesh.py:2072-2109.  padded=True returns the float4-padded [N,N,N,4] layout used inside solver.pm."""
    return _gradient(a, None, 0.0, 0, gradient_order, False, padded=padded)


def derivative_fR(a, b, f, fR_n, gradient_order, padded=False):
    """mesh.py:2112-2174"""
    if fR_n not in (1, 2):
        raise NotImplementedError(f"Unsupported: {fR_n=}")
    return _gradient(a, b, f, fR_n, gradient_order, False, padded=padded)


def add_derivative_fR(force, b, f, fR_n, gradient_order) -> None:
    """mesh.py:2177-2237 (in place on force)"""
    if fR_n not in (1, 2):
        raise NotImplementedError(f"Unsupported: {fR_n=}")
    _gradient(None, b, f, fR_n, gradient_order, True, force)


def _restriction(x, sign):
    c = _lib.Ctx()
    tx = c.dev(x)
    N = tx.shape[0]
    out = _lib.empty((N // 2,) * 3)
    _lib.check(_lib.load().psc_restriction(_lib.ptr(tx), N, sign, _lib.ptr(out), _lib.stream()))
    return c.ret(out)


def restriction(x):
    """mesh.py:14-60"""
    return _restriction(x, 1.0)


def minus_restriction(x):
    """mesh.py:62-108"""
    return _restriction(x, -1.0)


def prolongation(x):
    """mesh.py:180-330"""
    c = _lib.Ctx()
    tx = c.dev(x)
    Nc = tx.shape[0]
    out = _lib.empty((2 * Nc,) * 3)
    _lib.check(_lib.load().psc_prolongation(_lib.ptr(out), _lib.ptr(tx), Nc, 0, _lib.stream()))
    return c.ret(out)


def add_prolongation(y, x) -> None:
    """mesh.py:334-453: y += P(x), in place"""
    c = _lib.Ctx()
    ty, tx = c.dev(y, inplace=True), c.dev(x)
    _lib.check(_lib.load().psc_prolongation(_lib.ptr(ty), _lib.ptr(tx), tx.shape[0], 1, _lib.stream()))
    c.finish()
