"""Mirror of pysco/integration.py: time-step selection and the leapfrog / Euler integrators.

integrate :17-118, euler :121-189, leapfrog :192-264, dt_CFL_maxacc :267-295, dt_CFL_maxvel :298-326,
dt_weak_variation :329-358.

B200 fusions (same arithmetic): kick+drift+wrap is one kernel (60 B/particle); the second half-kick
and the max|a|, max|v| reductions of the NEXT step's dt are folded into the force interpolation.
"""
import logging
import weakref

import numpy as np
import torch

from . import _lib, distributed, solver, utils

# maxima produced by the fused interpolation kernel, valid for exactly the (acceleration, velocity)
# tensors returned by the last leapfrog step
_maxima_cache = {"acc": None, "vel": None, "max": None}


def _cached_max(x, which):
    ref = _maxima_cache[which]
    if ref is not None and isinstance(x, torch.Tensor) and ref() is x and _maxima_cache["max"] is not None:
        return np.float32(_maxima_cache["max"][0 if which == "acc" else 1])
    m = utils.max_abs(x)
    if distributed.is_active():  # particle-parallel: the time step is global
        t = torch.tensor([float(m)], dtype=torch.float32, device=_lib.device())
        m = np.float32(distributed.allreduce_max_(t).item())
    return m


def dt_CFL_maxacc(acceleration, param):
    """integration.py:267-295 (free fall)"""
    dx = np.float32(0.5 ** param["ncoarse"])
    max_acc = _cached_max(acceleration, "acc")
    return np.float32(param["Courant_factor"]) * np.sqrt(dx / max_acc)


def dt_CFL_maxvel(velocity, param):
    """integration.py:298-326"""
    dx = np.float32(0.5 ** param["ncoarse"])
    max_vel = _cached_max(velocity, "vel")
    return np.float32(param["Courant_factor"]) * dx / max_vel


def dt_weak_variation(func_t_a, param):
    """integration.py:329-358"""
    aexp_factor = 1.0 + 0.01 * param["max_aexp_stepping"]
    return np.float32(func_t_a(np.log(aexp_factor * param["aexp"])) - func_t_a(np.log(param["aexp"])))


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


def leapfrog(position, velocity, acceleration, potential, additional_field, dt, tables, param):
    """integration.py:192-264 (kick-drift-kick)"""
    c = _lib.Ctx()
    pos, vel, acc, pot, add = _to_device(c, position, velocity, acceleration, potential, additional_field)
    half_dt = np.float32(0.5 * dt)
    dt_is_f64 = 0 if isinstance(dt, np.float32) else 1
    _lib.check(_lib.load().psc_kick_drift_wrap(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(acc), pos.shape[0],
                                               float(half_dt), float(dt), dt_is_f64, _lib.stream()))
    _advance_clock(dt, tables, param)
    acc, pot, add, maxima = solver._pm_device(pos, param, pot, add, tables, kick=(vel, half_dt))
    distributed.allreduce_max_(maxima)
    mx = maxima.cpu().numpy()  # one 8-byte read: max|a|, max|v| for the next integrate()
    _maxima_cache.update(acc=weakref.ref(acc), vel=weakref.ref(vel), max=mx)
    return _from_device(c, position, velocity, pos, vel, acc, pot, add)


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
