"""Mirror of pysco/integration.py: time-step selection and the leapfrog / Euler integrators.

integrate :17-118, euler :121-189, leapfrog :192-264, dt_CFL_maxacc :267-295, dt_CFL_maxvel :298-326,
dt_weak_variation :329-358.

B200 fusions (same arithmetic): kick+drift+wrap is one kernel (60 B/particle); the second half-kick
and the max|a|, max|v| reductions of the NEXT step's dt are folded into the force interpolation.
"""
import logging
import os
import weakref

import numpy as np
import torch

from . import _lib, distributed, mesh, solver, utils

# maxima produced by the fused interpolation kernel, valid for exactly the (acceleration, velocity)
# tensors returned by the last leapfrog step AND only as long as nobody has written to them since (the tensor's
# version counter is stored with the weak reference: an in-place change by the caller is a cache miss)
_maxima_cache = {"acc": None, "vel": None, "max": None}


def _remember_maxima(acc, vel, mx):
    _maxima_cache.update(acc=(weakref.ref(acc), acc._version), vel=(weakref.ref(vel), vel._version), max=mx)


def _cached_max(x, which):
    entry = _maxima_cache[which]
    if entry is not None and isinstance(x, torch.Tensor) and entry[0]() is x and entry[1] == x._version \
            and _maxima_cache["max"] is not None:
        return np.float32(_maxima_cache["max"][0 if which == "acc" else 1])
    m = utils.max_abs(x)
    if distributed.is_active():  # particle-parallel: the time step is global
        t = torch.tensor([float(m)], dtype=torch.float32, device=_lib.device())
        m = np.float32(distributed.allreduce_max_(t).item())
    return m


def _finite(m, what):
    """The fused interpolation kernel lets a NaN win its max reduction: a non-finite force (e.g. the f(R) cubic root
    without a real branch, cubic.py:196-197, where the reference stops with a math domain error) must stop the run
    instead of silently turning every position into NaN."""
    if not np.isfinite(m):
        raise ValueError(f"math domain error: max|{what}| is {m} (non-finite {what} after the force computation)")
    return m


def dt_CFL_maxacc(acceleration, param):
    """integration.py:267-295 (free fall)"""
    dx = np.float32(0.5 ** param["ncoarse"])
    max_acc = _finite(_cached_max(acceleration, "acc"), "acceleration")
    return np.float32(param["Courant_factor"]) * np.sqrt(dx / max_acc)


def dt_CFL_maxvel(velocity, param):
    """integration.py:298-326"""
    dx = np.float32(0.5 ** param["ncoarse"])
    max_vel = _finite(_cached_max(velocity, "vel"), "velocity")
    return np.float32(param["Courant_factor"]) * dx / max_vel


_dt3_memo = [None, None]


def dt_weak_variation(func_t_a, param):
    """integration.py:329-358.  The value depends on a(t) only; leapfrog asks for it right after it has advanced the
    clock (the guess of the speculative bin count) and integrate() asks again at the start of the next step, on the
    critical path between two steps: the second call is answered from a one-entry memo."""
    aexp, stepping = param["aexp"], param["max_aexp_stepping"]
    key = (float(aexp), float(stepping))
    owner = _dt3_memo[0]
    if owner is not None and owner[0]() is func_t_a and owner[1] == key:
        return _dt3_memo[1]
    aexp_factor = 1.0 + 0.01 * stepping
    val = np.float32(func_t_a(np.log(aexp_factor * aexp)) - func_t_a(np.log(aexp)))
    try:
        _dt3_memo[0], _dt3_memo[1] = (weakref.ref(func_t_a), key), val
    except TypeError:      # a table object that cannot be weakly referenced: no memo
        _dt3_memo[0] = None
    return val


def _advance_clock(dt, tables, param):
    param["t"] += dt
    param["aexp_old"] = param["aexp"]
    param["aexp"] = np.exp(tables[0](param["t"]))
    logging.info(f"{param['t']=} {param['aexp']=}")
    utils.set_units(param)


def _to_device(c, position, velocity, acceleration, potential, additional_field):
    pos, vel, acc = c.dev(position, inplace=True), c.dev(velocity, inplace=True), c.dev(acceleration)
    pot = c.dev(potential) if len(potential) else potential
    add = c.dev(additional_field) if len(additional_field) else additional_field
    return pos, vel, acc, pot, add


def _from_device(c, position, velocity, pos, vel, acc, pot, add):
    if c.np_mode:
        c.finish()  # position / velocity were updated in place, like the reference
        return (position, velocity, c.ret(acc), c.ret(pot) if len(pot) else pot, c.ret(add) if len(add) else add)
    return pos, vel, acc, pot, add


def _binned_for(pos, param):
    """Bin scratch for the step if the binned particle<->mesh kernels will be used (solver._pm_device)."""
    N = 2 ** param["ncoarse"]
    n = pos.shape[0]
    if mesh.can_bin(N, n) and pos.data_ptr() % 16 == 0:
        return mesh.step_binned(n, N)
    return None


_copy_streams = {}


def _streams():
    d = torch.cuda.current_device()
    if d not in _copy_streams:
        _copy_streams[d] = (torch.cuda.Stream(), torch.cuda.Stream())
    return _copy_streams[d]


_pinned_pool = []


def _pinned_empty(shape):
    """Pinned host staging buffer for a result.  cudaHostAlloc of a 1.6 GB array costs more than the copy that fills
    it, so buffers are pooled; one is handed out again only when the caller has dropped every reference to it (the
    reference returns fresh arrays each step: a result the caller still holds is never overwritten)."""
    import sys
    shape = tuple(shape)
    for t in _pinned_pool:
        # references: the pool's list, the loop variable, getrefcount's argument
        if tuple(t.shape) == shape and sys.getrefcount(t) <= 3:
            return t
    t = torch.empty(shape, dtype=torch.float32, pin_memory=True)
    _pinned_pool.append(t)
    if len(_pinned_pool) > 8:   # shapes no longer in use (another run in the same process)
        _pinned_pool[:] = [u for u in _pinned_pool if sys.getrefcount(u) > 3 or tuple(u.shape) == shape][-8:]
    return t


def _is_pinned_host(*ts):
    return all(isinstance(t, torch.Tensor) and not t.is_cuda and t.is_pinned() and t.dtype == torch.float32
               and t.is_contiguous() for t in ts)


def _leapfrog_pinned(position, velocity, acceleration, potential, additional_field, dt, tables, param, nchunk=8):
    """leapfrog() for PINNED host tensors: the PCIe transfers are pipelined with the kernels instead of being
    serialised around them.  The particle arrays are uploaded in chunks on a copy stream; each chunk is
    kicked / drifted / wrapped as soon as it has landed and its final positions start flowing back on a second
    copy stream while later chunks are still arriving (full-duplex PCIe).  After the solve, velocity (in place),
    acceleration and potential are downloaded on the copy stream.  Same arithmetic as leapfrog()."""
    lib = _lib.load()
    dev = _lib.device()
    n = position.shape[0]
    cur = torch.cuda.current_stream()
    s_in, s_out = _streams()
    pos, vel, acc = (torch.empty((n, 3), dtype=torch.float32, device=dev) for _ in range(3))
    for t in (pos, vel, acc):
        t.record_stream(s_in)
        t.record_stream(s_out)
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    half_dt = np.float32(0.5 * dt)
    dt_is_f64 = 0 if isinstance(dt, np.float32) else 1
    solver_name = param["linear_newton_solver"].casefold()
    pot = potential
    if len(potential) and solver_name == "multigrid":      # the previous potential is only a multigrid first guess
        pot = torch.empty(potential.shape, dtype=torch.float32, device=dev)
        pot.record_stream(s_in)
        with torch.cuda.stream(s_in):
            pot.copy_(potential, non_blocking=True)
    elif len(potential):
        pot = torch.empty(0, dtype=torch.float32, device=dev)
    add = additional_field
    if len(additional_field):
        add = additional_field.to(dev, non_blocking=True)
    # rows per chunk: a multiple of 4 keeps every chunk 16-byte aligned
    step = max(4, ((n + nchunk - 1) // nchunk + 3) // 4 * 4)
    counted = _binned_for(pos, param)
    for a in range(0, n, step):
        b = min(n, a + step)
        with torch.cuda.stream(s_in):
            pos[a:b].copy_(position[a:b], non_blocking=True)
            vel[a:b].copy_(velocity[a:b], non_blocking=True)
            acc[a:b].copy_(acceleration[a:b], non_blocking=True)
            landed = torch.cuda.Event()
            landed.record(s_in)
        cur.wait_event(landed)
        if counted is not None:
            mesh.kick_drift_wrap_count(pos[a:b], vel[a:b], acc[a:b], half_dt, dt, dt_is_f64, counted,
                                       zero_counts=(a == 0), row0=a)
        else:
            _lib.check(lib.psc_kick_drift_wrap(_lib.ptr(pos[a:b]), _lib.ptr(vel[a:b]), _lib.ptr(acc[a:b]), b - a,
                                               float(half_dt), float(dt), dt_is_f64, _lib.stream()))
        drifted = torch.cuda.Event()
        drifted.record(cur)
        s_out.wait_event(drifted)
        with torch.cuda.stream(s_out):
            position[a:b].copy_(pos[a:b], non_blocking=True)
    _advance_clock(dt, tables, param)
    cur.wait_stream(s_in)
    del acc
    acc, pot, add, maxima = solver._pm_device(pos, param, pot, add, tables, kick=(vel, half_dt), counted=counted)
    distributed.allreduce_max_(maxima)
    acc_h = _pinned_empty((n, 3))
    pot_h = _pinned_empty(pot.shape) if len(pot) else pot
    add_h = add
    s_out.wait_stream(cur)
    for t in (acc, pot, add):
        if isinstance(t, torch.Tensor) and t.is_cuda:
            t.record_stream(s_out)
    with torch.cuda.stream(s_out):
        velocity.copy_(vel, non_blocking=True)
        acc_h.copy_(acc, non_blocking=True)
        if len(pot):
            pot_h.copy_(pot, non_blocking=True)
        if len(add):
            add_h = _pinned_empty(add.shape)
            add_h.copy_(add, non_blocking=True)
    mx = maxima.cpu().numpy()
    s_out.synchronize()
    # the cached maxima belong to exactly these host tensors
    _remember_maxima(acc_h, velocity, mx)
    return position, velocity, acc_h, pot_h, add_h


def leapfrog(position, velocity, acceleration, potential, additional_field, dt, tables, param):
    """integration.py:192-264 (kick-drift-kick)"""
    if _is_pinned_host(position, velocity, acceleration) and (not len(potential) or _is_pinned_host(potential)) \
            and (not len(additional_field) or _is_pinned_host(additional_field)) and position.shape[0] >= 4096:
        return _leapfrog_pinned(position, velocity, acceleration, potential, additional_field, dt, tables, param)
    c = _lib.Ctx()
    pos, vel, acc, pot, add = _to_device(c, position, velocity, acceleration, potential, additional_field)
    half_dt = np.float32(0.5 * dt)
    dt_is_f64 = 0 if isinstance(dt, np.float32) else 1
    N = 2 ** param["ncoarse"]
    if _bin_ordered_loop(c, pos, vel, acc, param):
        # Device-resident arrays: the step re-sorts them into bin order while it kicks and drifts (no binned copy, no
        # source-row indirection in the deposit / interpolation).  The arrays returned are NEW tensors in bin order;
        # utils.particle_ids / utils.reference_order give the reference's rows back.
        sb = mesh.step_sorted(pos.shape[0], N)
        pos, vel, ids = mesh.step_sort(pos, vel, acc, utils.particle_ids(pos), half_dt, dt, dt_is_f64, sb)
        _advance_clock(dt, tables, param)
        if not os.environ.get("PSC_NO_PREDICTED_COUNT"):
            # The next time step is min(free fall, Courant, scale-factor variation) (integrate, above); the last of the
            # three depends on a(t) only and is the one that binds through most of a cosmological run.  Let the
            # interpolation kernel count the next sort's bins under that guess; step_sort checks the guess.
            # After three unused guesses in a row (another criterion sets the step) it pauses for eight steps.
            if sb.miss_streak >= 3:
                sb.miss_streak, sb.rest = 0, 8
            if sb.rest > 0:
                sb.rest -= 1
            else:
                dt_guess = dt_weak_variation(tables[1], param)
                sb.predict_next = (np.float32(0.5 * dt_guess), dt_guess, 0 if isinstance(dt_guess, np.float32) else 1)
        acc, pot, add, maxima = solver._pm_device(pos, param, pot, add, tables, kick=(vel, half_dt), counted=sb)
        utils.set_particle_ids((pos, vel, acc), ids)
        mx = maxima.cpu().numpy()
        _remember_maxima(acc, vel, mx)
        return pos, vel, acc, pot, add
    counted = _binned_for(pos, param)
    if counted is not None:   # first pass of the binning folded into the kick-drift-wrap
        mesh.kick_drift_wrap_count(pos, vel, acc, half_dt, dt, dt_is_f64, counted)
    else:
        _lib.check(_lib.load().psc_kick_drift_wrap(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(acc), pos.shape[0],
                                                   float(half_dt), float(dt), dt_is_f64, _lib.stream()))
    _advance_clock(dt, tables, param)
    acc, pot, add, maxima = solver._pm_device(pos, param, pot, add, tables, kick=(vel, half_dt), counted=counted)
    distributed.allreduce_max_(maxima)
    mx = maxima.cpu().numpy()  # one 8-byte read: max|a|, max|v| for the next integrate()
    _remember_maxima(acc, vel, mx)
    return _from_device(c, position, velocity, pos, vel, acc, pot, add)


def _bin_ordered_loop(c, pos, vel, acc, param):
    """The bin-ordered time loop applies to device tensors (a NumPy / host caller gets its arrays updated in place, row
    for row, like the reference), meshes the binned kernels take, the fused gradient + interpolation (not full_fft) and
    one process per problem.  param["particle_order"] = "reference" keeps the rows where they are (shadow binning)."""
    if c.np_mode or distributed.is_active():
        return False
    if str(param["particle_order"] if "particle_order" in param.index else "bins").casefold() == "reference":
        return False
    N = 2 ** param["ncoarse"]
    ok = all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
             and t.data_ptr() % 16 == 0 for t in (pos, vel, acc))
    return (ok and mesh.can_bin(N, pos.shape[0]) and N >= 16
            and param["linear_newton_solver"].casefold() != "full_fft"
            and param["mass_scheme"].casefold() in ("cic", "tsc"))


def euler(position, velocity, acceleration, potential, additional_field, dt, tables, param):
    """integration.py:121-189"""
    c = _lib.Ctx()
    pos, vel, acc, pot, add = _to_device(c, position, velocity, acceleration, potential, additional_field)
    utils.add_vector_scalar_inplace(pos, vel, dt)
    _advance_clock(dt, tables, param)
    utils.periodic_wrap(pos)
    utils.add_vector_scalar_inplace(vel, acc, -dt)
    acc, pot, add, _ = solver._pm_device(pos, param, pot, add, tables)
    _maxima_cache.update(acc=None, vel=None, max=None)
    return _from_device(c, position, velocity, pos, vel, acc, pot, add)


def integrate(position, velocity, acceleration, potential, additional_field, tables, param,
              t_snap_next=np.float32(0)):
    """integration.py:17-118: one step; dt = min(free-fall, velocity Courant, scale-factor variation),
    shortened to land on the next snapshot time (sets param["write_snapshot"])."""
    dt1 = dt_CFL_maxacc(acceleration, param)
    dt2 = dt_CFL_maxvel(velocity, param)
    dt3 = dt_weak_variation(tables[1], param)
    dt = np.min([dt1, dt2, dt3])
    if (param["t"] + dt) > t_snap_next:
        dt = t_snap_next - param["t"]
        param["write_snapshot"] = True
    else:
        param["write_snapshot"] = False
    logging.info(f"Conditions: velocity {dt1=}, acceleration {dt2=}, scale factor {dt3=}")
    INTEGRATOR = param["integrator"].casefold()
    if INTEGRATOR == "leapfrog":
        return leapfrog(position, velocity, acceleration, potential, additional_field, dt, tables, param)
    if INTEGRATOR == "euler":
        return euler(position, velocity, acceleration, potential, additional_field, dt, tables, param)
    raise NotImplementedError("ERROR: Integrator must be 'leapfrog' or 'euler'")
