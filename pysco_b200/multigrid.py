"""Mirror of pysco/multigrid.py: host sequencing of the multigrid solvers over the CUDA kernels.

linear / V_cycle / F_cycle / W_cycle      multigrid.py:23-83, 474-517, 583-638, 722-776
FAS / V_cycle_FAS / F_cycle_FAS / W_cycle_FAS   multigrid.py:88-140, 521-579, 642-718, 780-858
All cycles update ``x`` in place (and ``linear`` / ``FAS`` also return it), like the reference.
"""
import logging

import numpy as np
import torch

from . import _lib, cubic, laplacian, mesh, quartic, utils

_EMPTY = np.empty(0, dtype=np.float32)


def _fr_module(param):
    n = param["fR_n"]
    if n == 1:
        return cubic
    if n == 2:
        return quartic
    raise NotImplementedError(f"Only f(R) with n = 1 and 2, currently {param['fR_n']=}")


def _is_scalaron(param) -> bool:
    return bool(param["compute_additional_field"]) and "fr" == param["theory"].casefold()


def _require_scalaron(param):
    if not _is_scalaron(param):
        # reference: laplacian_reformulated (multigrid.py:116-121 doctest usage only)
        raise NotImplementedError(
            "FAS on a linear problem (laplacian_reformulated) is not part of the B200 hot path; "
            "use multigrid.linear / V_cycle for the Newtonian potential")


def _coarsest(param, nlevel) -> bool:
    return nlevel >= (param["ncoarse"] - 3)


# ------------------------------------------------------------------------------ linear cycles
def _cycle(kind, x, b, param, nlevel):
    laplacian.smoothing(x, b, param["Npre"])
    visits = (kind,) if kind == "V" else (kind, "V" if kind == "F" else "W")
    for n, sub in enumerate(visits):
        if n:
            laplacian.smoothing(x, b, param["Npre"])
        res_c = laplacian.restrict_residual(x, b)
        x_corr_c = laplacian.initialise_potential(res_c)
        if _coarsest(param, nlevel):
            laplacian.smoothing(x_corr_c, res_c, param["Npre"])
        else:
            _cycle(sub, x_corr_c, res_c, param, nlevel + 1)
        mesh.add_prolongation(x, x_corr_c)
    laplacian.smoothing(x, b, param["Npost"])


# A V-cycle at 256^3 is ~45 kernels, most of them on coarse levels where a launch costs more than the kernel: the
# whole cycle is captured once per (grid size, smoothing counts) into a CUDA graph working on persistent buffers and
# replayed by linear() -- one launch per cycle instead of one per kernel.
_vcycle_graphs = {}


def _graphs_enabled():
    import os
    return not os.environ.get("PSC_NO_GRAPHS")


def _vcycle_graph(N, param):
    key = (torch.cuda.current_device(), int(N), int(param["Npre"]), int(param["Npost"]), int(param["ncoarse"]))
    if key not in _vcycle_graphs:
        gx, gb = _lib.zeros((N, N, N)), _lib.zeros((N, N, N))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up outside the capture (allocator, lazy module loading)
            _cycle("V", gx, gb, param, 0)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            _cycle("V", gx, gb, param, 0)
        _vcycle_graphs[key] = (graph, gx, gb)
    return _vcycle_graphs[key]


def _run_cycle(kind, x, b, param, nlevel):
    c = _lib.Ctx()
    tx, tb = c.dev(x, inplace=True), c.dev(b)
    _cycle(kind, tx, tb, param, nlevel)
    c.finish()


def V_cycle(x, b, param, nlevel=0) -> None:
    _run_cycle("V", x, b, param, nlevel)


def F_cycle(x, b, param, nlevel=0) -> None:
    _run_cycle("F", x, b, param, nlevel)


def W_cycle(x, b, param, nlevel=0) -> None:
    _run_cycle("W", x, b, param, nlevel)


def linear(x, b, param):
    """multigrid.py:23-83"""
    THEORY = param["theory"].casefold()
    if param["compute_additional_field"] and "fr" == THEORY:
        raise ValueError("Linear should not be used for scalaron field")
    c = _lib.Ctx()
    tx, tb = c.dev(x, inplace=True), c.dev(b)
    mond_pass = (not param["compute_additional_field"]) and "mond" == THEORY
    if ("tolerance" not in param) or (param["nsteps"] % 3) == 0:
        logging.info("Compute Truncation error")
        tolerance = param["epsrel"] * laplacian.truncation_error(tx)
        if mond_pass:
            param["tolerance_mond"] = tolerance
        else:
            param["tolerance"] = tolerance
    tolerance = param["tolerance_mond"] if mond_pass else param["tolerance"]
    logging.info("Start linear Multigrid")
    residual_err = 1e30
    N = tx.shape[0]
    graph = None
    if _graphs_enabled() and N >= 32 and not torch.cuda.is_current_stream_capturing():
        graph, gx, gb = _vcycle_graph(N, param)
        gx.copy_(tx)
        gb.copy_(tb)
        wx, wb = gx, gb
    else:
        wx, wb = tx, tb
    while residual_err > tolerance:
        if graph is not None:
            graph.replay()
        else:
            _cycle("V", wx, wb, param, 0)
        residual_error_tmp = laplacian.residual_error(wx, wb)
        logging.info(f"{residual_error_tmp=} {tolerance=}")
        if residual_error_tmp < tolerance or residual_err / residual_error_tmp < 2:
            break
        residual_err = residual_error_tmp
    if graph is not None:
        tx.copy_(gx)
    c.finish()
    return x if c.np_mode else tx


# --------------------------------------------------------------------------------- FAS cycles
def truncation_error(x, param, b=_EMPTY):
    """multigrid.py:143-189"""
    _require_scalaron(param)
    return _fr_module(param).truncation_error(x, b, np.float32(param["fR_q"]))


def normalisation_residual(param):
    return np.float32(4)


def residual_error(x, b, param):
    """multigrid.py:231-270"""
    _require_scalaron(param)
    return _fr_module(param).residual_error(x, b, np.float32(param["fR_q"]))


def restrict_residual(x, b, param, rhs=_EMPTY):
    """multigrid.py:287-348"""
    _require_scalaron(param)
    m = _fr_module(param)
    q = np.float32(param["fR_q"])
    if len(rhs) == 0:
        return mesh.minus_restriction(m.operator(x, b, q))
    return mesh.restriction(m.residual_with_rhs(x, b, q, rhs))


def smoothing(x, b, n_smoothing, param, rhs=_EMPTY) -> None:
    """multigrid.py:351-410"""
    _require_scalaron(param)
    m = _fr_module(param)
    q = np.float32(param["fR_q"])
    if len(rhs) == 0:
        m.smoothing(x, b, q, n_smoothing)
    else:
        m.smoothing_with_rhs(x, b, q, n_smoothing, rhs)


def operator(x, param, b=_EMPTY):
    """multigrid.py:413-470"""
    _require_scalaron(param)
    return _fr_module(param).operator(x, b, np.float32(param["fR_q"]))


def _cycle_FAS(kind, x, b, param, nlevel, rhs):
    smoothing(x, b, param["Npre"], param, rhs)
    visits = (kind,) if kind == "V" else (kind, "V" if kind == "F" else "W")
    b_c = mesh.restriction(b)
    for n, sub in enumerate(visits):
        if n:
            smoothing(x, b, param["Npre"], param, rhs)
        res_c = restrict_residual(x, b, param, rhs)
        x_c = mesh.restriction(x)
        x_corr_c = x_c.clone()
        L_c = operator(x_c, param, b_c)
        utils.linear_operator_vectors_inplace(res_c, normalisation_residual(param), L_c, np.float32(1))
        del L_c
        if _coarsest(param, nlevel):
            smoothing(x_corr_c, b_c, param["Npre"], param, res_c)
        else:
            _cycle_FAS(sub, x_corr_c, b_c, param, nlevel + 1, res_c)
        utils.add_vector_scalar_inplace(x_corr_c, x_c, np.float32(-1))
        mesh.add_prolongation(x, x_corr_c)
    smoothing(x, b, param["Npost"], param, rhs)


_fas_graphs = {}


def _fas_graph(N, param):
    """CUDA graph of one FAS V-cycle on persistent buffers.  q (param["fR_q"]) changes every step, so the f(R) kernels
    read it from a device scalar while the graph is captured (psc_mg_set_q_device) and the host rewrites that scalar
    before each replay."""
    key = (torch.cuda.current_device(), int(N), int(param["Npre"]), int(param["Npost"]), int(param["ncoarse"]),
           int(param["fR_n"]))
    if key not in _fas_graphs:
        lib = _lib.load()
        gx, gb, gq = _lib.zeros((N, N, N)), _lib.zeros((N, N, N)), _lib.zeros((1,))
        gx.fill_(1.0)
        gq.fill_(float(param["fR_q"]))
        _lib.check(lib.psc_mg_set_q_device(_lib.ptr(gq)))
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                _cycle_FAS("V", gx, gb, param, 0, _EMPTY)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                _cycle_FAS("V", gx, gb, param, 0, _EMPTY)
        finally:
            _lib.check(lib.psc_mg_set_q_device(None))
        _fas_graphs[key] = (graph, gx, gb, gq)
    return _fas_graphs[key]


def _run_cycle_FAS(kind, x, b, param, nlevel, rhs):
    c = _lib.Ctx()
    tx, tb = c.dev(x, inplace=True), c.dev(b)
    tr = c.dev(rhs) if len(rhs) else _EMPTY
    _cycle_FAS(kind, tx, tb, param, nlevel, tr)
    c.finish()


def V_cycle_FAS(x, b, param, nlevel=0, rhs=_EMPTY) -> None:
    _run_cycle_FAS("V", x, b, param, nlevel, rhs)


def F_cycle_FAS(x, b, param, nlevel=0, rhs=_EMPTY) -> None:
    _run_cycle_FAS("F", x, b, param, nlevel, rhs)


def W_cycle_FAS(x, b, param, nlevel=0, rhs=_EMPTY) -> None:
    _run_cycle_FAS("W", x, b, param, nlevel, rhs)


def FAS(x, b, param):
    """multigrid.py:88-140"""
    c = _lib.Ctx()
    tx, tb = c.dev(x, inplace=True), c.dev(b)
    if ("tolerance_FAS" not in param) or (param["nsteps"] % 3) == 0:
        logging.info("Compute FAS Truncation error")
        param["tolerance_FAS"] = param["epsrel"] * truncation_error(tx, param, tb)
    tolerance = param["tolerance_FAS"]
    logging.info("Start Full-Approximation Storage Multigrid")
    residual_err = 1e30
    N = tx.shape[0]
    graph = None
    if _graphs_enabled() and N >= 32 and not torch.cuda.is_current_stream_capturing():
        graph, gx, gb, gq = _fas_graph(N, param)
        gq.fill_(float(np.float32(param["fR_q"])))
        gx.copy_(tx)
        gb.copy_(tb)
        wx, wb = gx, gb
    else:
        wx, wb = tx, tb
    while residual_err > tolerance:
        if graph is not None:
            graph.replay()
        else:
            _cycle_FAS("V", wx, wb, param, 0, _EMPTY)
        residual_error_tmp = residual_error(wx, wb, param)
        logging.info(f"{residual_error_tmp=} {tolerance=}")
        if residual_error_tmp < tolerance or residual_err / residual_error_tmp < 2:
            break
        residual_err = residual_error_tmp
    if graph is not None:
        tx.copy_(gx)
    c.finish()
    return x if c.np_mode else tx
