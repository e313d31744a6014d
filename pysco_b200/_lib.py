"""ctypes binding of libpysco_b200.so (include/pysco_b200.h) and the small amount of device-buffer
plumbing the mirrored PySCo modules share.

PyTorch is used for device memory and streams only.  There is NO CPU fallback: if the CUDA library
is missing or no GPU is visible, every compute entry point raises.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpysco_b200.so")

_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/pysco_b200.h
SIGNATURES = {
    "psc_last_error": [],
    "psc_version": [],
    "psc_launch_count": [],
    "psc_morton_keys": [_vp, _i64, _vp, _vp],
    "psc_argsort_workspace_bytes": [_i64],
    "psc_argsort_keys": [_vp, _i64, _vp, _vp, _sz, _vp],
    "psc_gather3": [_vp, _vp, _vp, _i64, _vp],
    "psc_axpy": [_vp, _vp, _d, _i, _i64, _vp],
    "psc_periodic_wrap": [_vp, _i64, _vp],
    "psc_max_abs": [_vp, _i64, _vp, _vp],
    "psc_kick_drift_wrap": [_vp, _vp, _vp, _i64, _f, _d, _i, _vp],
    "psc_deposit": [_vp, _i64, _i, _i, _f, _f, _f, _vp, _vp],
    "psc_interp": [_vp, _vp, _i64, _i, _i, _i, _vp, _vp],
    "psc_interp_kick": [_vp, _vp, _vp, _vp, _i64, _i, _i, _f, _vp, _vp],
    "psc_bin_workspace_bytes": [_i64, _i],
    "psc_bin_particles": [_vp, _i64, _i, _vp, _sz, _vp],
    "psc_kick_drift_wrap_count": [_vp, _vp, _vp, _i64, _f, _d, _i, _i, _i64, _vp, _sz, _i, _i, _i64, _vp],
    "psc_bin_particles_counted": [_vp, _i64, _i, _vp, _sz, _i, _vp],
    "psc_deposit_binned": [_vp, _sz, _i64, _i, _i, _f, _f, _f, _vp, _vp],
    "psc_interp_kick4_binned": [_vp, _vp, _sz, _vp, _vp, _i64, _i, _i, _f, _vp, _vp],
    "psc_interp_kick_phi_binned": [_vp, _vp, _f, _i, _i, _vp, _sz, _vp, _vp, _i64, _i, _i, _f, _vp, _vp],
    "psc_sorted_workspace_bytes": [_i64, _i],
    "psc_morton_ids_sorted": [_vp, _vp, _sz, _i, _i64, _i, _vp, _vp, _vp],
    "psc_step_sort": [_vp, _vp, _vp, _vp, _i64, _f, _d, _i, _i, _i, _i, _vp, _sz, _vp, _vp, _vp, _vp],
    "psc_deposit_sorted": [_vp, _vp, _sz, _i, _i64, _i, _i, _f, _f, _f, _vp, _vp],
    "psc_interp_kick_phi_sorted": [_vp, _vp, _f, _i, _i, _vp, _vp, _sz, _i, _vp, _vp, _i64, _i, _i, _f, _vp, _i, _f, _d,
                                   _i, _vp],
    "psc_scatter3_by_id": [_vp, _vp, _vp, _i64, _vp],
    "psc_sorted_workspace_bytes_slab": [_i64, _i, _i],
    "psc_sort_by_bin_slab": [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _i64, _vp, _sz, _vp, _vp, _vp, _vp],
    "psc_deposit_sorted_slab": [_vp, _vp, _sz, _i, _i64, _i, _i, _i, _i, _vp, _vp],
    "psc_interp_kick_phi_sorted_slab": [_vp, _vp, _f, _i, _i, _i, _i, _i, _vp, _vp, _sz, _i, _vp, _vp, _i64, _i, _i, _f,
                                        _vp, _vp],
    "psc_bin_workspace_bytes_slab": [_i64, _i, _i],
    "psc_bin_particles_slab": [_vp, _i64, _i, _i, _i, _vp, _sz, _vp],
    "psc_deposit_binned_slab": [_vp, _sz, _i64, _i, _i, _i, _i, _vp, _vp],
    "psc_interp_kick_phi_binned_slab": [_vp, _vp, _f, _i, _i, _i, _i, _i, _vp, _sz, _vp, _vp, _i64, _i, _i, _f, _vp,
                                        _vp],
    "psc_kick_drift_wrap_slab": [_vp, _vp, _vp, _i64, _f, _d, _i, _i, _i, _i, _i, _vp, _vp, _i64, _vp],
    "psc_slab_pack_rows": [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "psc_slab_pack_fixed": [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "psc_slab_count": [_vp, _i64, _i, _i, _i, _i, _vp, _vp],
    "psc_slab_pack_leavers": [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "psc_slab_unpack_rows": [_vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "psc_slab_move_rows": [_vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "psc_slab_fft_plan_create": [_i, _i, _i, C.POINTER(_vp)],
    "psc_slab_fft_plan_destroy": [_vp],
    "psc_slab_fft_workspace_bytes": [_vp],
    "psc_slab_fft_set_workspace": [_vp, _vp],
    "psc_slab_fft_r2c_planes": [_vp, _vp, _vp, _vp],
    "psc_slab_fft_c2r_planes": [_vp, _vp, _vp, _vp],
    "psc_slab_fft_x": [_vp, _vp, _i, _vp],
    "psc_slab_transpose_put": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "psc_slab_yblocks": [_vp, _vp, _i, _i, _i, _i, _vp],
    "psc_pk_slab": [_vp, _i, _i, _i, _i, _vp, _vp],
    "psc_green_slab": [_vp, _i, _i, _i, _i, _i, _f, _vp],
    "psc_mg_set_q_device": [_vp],
    "psc_linear_operator": [_vp, _f, _f, _vp, _i64, _vp],
    "psc_lincomb": [_vp, _f, _vp, _f, _i64, _vp],
    "psc_gradient": [_vp, _vp, _f, _i, _i, _i, _i, _vp, _i, _vp],
    "psc_interp_kick4": [_vp, _vp, _vp, _vp, _i64, _i, _i, _f, _vp, _vp],
    "psc_fft_plan_create": [_i, C.POINTER(_vp)],
    "psc_fft_plan_destroy": [_vp],
    "psc_fft_plan_workspace_bytes": [_vp],
    "psc_fft_r2c": [_vp, _vp, _vp, _vp],
    "psc_fft_c2r": [_vp, _vp, _vp, _vp],
    "psc_fft_poisson_supported": [_i],
    "psc_fft_poisson": [_vp, _vp, _vp, _vp, _i, _i, _f, _vp],
    "psc_xfft_green_slab": [_vp, _i, _i, _i, _i, _i, _f, _vp],
    "psc_fft_c2r_vec3": [_vp, _vp, _vp, _vp],
    "psc_green": [_vp, _i, _i, _i, _f, _vp],
    "psc_grad_green": [_vp, _i, _i, _f, _vp, _vp],
    "psc_pk": [_vp, _i, _i, _vp, _vp],
    "psc_operator": [_vp, _vp, _f, _i, _i, _vp, _vp],
    "psc_residual": [_vp, _vp, _f, _vp, _i, _i, _vp, _vp],
    "psc_restrict_residual": [_vp, _vp, _i, _vp, _vp],
    "psc_residual_sumsq": [_vp, _vp, _f, _i, _i, _vp, _vp],
    "psc_diff_sumsq": [_vp, _f, _vp, _i64, _vp, _vp],
    "psc_initialise_potential": [_vp, _f, _i, _i, _vp, _vp],
    "psc_gauss_seidel": [_vp, _vp, _f, _vp, _i, _i, _f, _vp],
    "psc_gauss_seidel_fused_supported": [_i],
    "psc_gauss_seidel_fused": [_vp, _vp, _f, _vp, _i, _i, _f, _vp, _i, _vp],
    "psc_mg_q_device_ptr": [],
    "psc_restriction": [_vp, _i, _f, _vp, _vp],
    "psc_prolongation": [_vp, _vp, _i, _i, _vp],
    "psc_mond_rhs": [_vp, _vp, _i, _f, _i, _f, _vp],
    "psc_box_gauss_seidel_colour": [_vp, _vp, _i, _i, _i, _i, _f, _vp],
    "psc_box_operator": [_vp, _i, _i, _vp, _vp],
    "psc_box_restrict_residual": [_vp, _vp, _i, _i, _vp, _vp],
    "psc_box_restriction": [_vp, _i, _i, _f, _vp, _vp],
    "psc_box_add_prolongation": [_vp, _vp, _i, _i, _vp],
    "psc_box_mond_rhs": [_vp, _vp, _i, _i, _f, _i, _f, _vp],
    "psc_box_gauss_seidel_colour_fr": [_vp, _vp, _vp, _f, _i, _i, _i, _i, _f, _i, _vp],
    "psc_box_operator_fr": [_vp, _vp, _f, _i, _i, _i, _vp, _vp],
    "psc_box_initialise_potential_fr": [_vp, _f, _i, _i, _i, _vp, _vp],
}
_RESTYPES = {"psc_mg_q_device_ptr": C.c_void_p, "psc_last_error": C.c_char_p, "psc_launch_count": _i64, "psc_argsort_workspace_bytes": _sz,
             "psc_bin_workspace_bytes": _sz, "psc_bin_workspace_bytes_slab": _sz, "psc_sorted_workspace_bytes": _sz, "psc_sorted_workspace_bytes_slab": _sz,
             "psc_slab_fft_workspace_bytes": _sz,
             "psc_fft_plan_workspace_bytes": _sz}

NGP, CIC, TSC = 0, 1, 2
GREEN_PLAIN, GREEN_COMPENSATED, GREEN_7PT = 0, 1, 2
OP_LAPLACIAN, OP_CUBIC, OP_QUARTIC = 0, 1, 2
MOND_FN = {"simple": 0, "n": 1, "beta": 2, "gamma": 3, "delta": 4}

_lib = None
_timing = None  # when enabled: list of (name, start_event, end_event)


def enable_timing(on: bool = True):
    """bench.py hook: record a CUDA-event pair around every C-ABI call on the launching stream."""
    global _timing
    _timing = [] if on else None


def timing_records():
    return _timing


class _Timed:
    __slots__ = ("fn", "name")

    def __init__(self, fn, name):
        self.fn, self.name = fn, name

    def __call__(self, *args):
        if _timing is None or torch.cuda.is_current_stream_capturing():
            return self.fn(*args)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = self.fn(*args)
        e1.record()
        _timing.append((self.name, e0, e1))
        return rc


_TIMED = ("psc_morton_keys", "psc_argsort_keys", "psc_gather3", "psc_axpy", "psc_periodic_wrap", "psc_max_abs",
          "psc_kick_drift_wrap", "psc_kick_drift_wrap_count", "psc_bin_particles_counted", "psc_step_sort", "psc_deposit_sorted", "psc_interp_kick_phi_sorted", "psc_scatter3_by_id", "psc_morton_ids_sorted", "psc_deposit", "psc_bin_particles", "psc_deposit_binned", "psc_interp_kick4_binned", "psc_interp_kick_phi_binned", "psc_interp", "psc_interp_kick", "psc_interp_kick4", "psc_linear_operator",
          "psc_lincomb", "psc_gradient", "psc_fft_r2c", "psc_fft_c2r", "psc_fft_c2r_vec3", "psc_fft_poisson", "psc_green",
          "psc_grad_green", "psc_pk", "psc_operator", "psc_residual", "psc_restrict_residual",
          "psc_residual_sumsq", "psc_diff_sumsq", "psc_initialise_potential", "psc_gauss_seidel", "psc_gauss_seidel_fused",
          "psc_restriction", "psc_prolongation", "psc_mond_rhs",
          "psc_bin_particles_slab", "psc_deposit_binned_slab", "psc_interp_kick_phi_binned_slab", "psc_sort_by_bin_slab", "psc_deposit_sorted_slab", "psc_interp_kick_phi_sorted_slab", "psc_slab_count",
          "psc_slab_pack_leavers", "psc_kick_drift_wrap_slab", "psc_slab_pack_rows", "psc_slab_pack_fixed", "psc_slab_unpack_rows", "psc_slab_move_rows", "psc_slab_fft_r2c_planes",
          "psc_slab_fft_c2r_planes", "psc_slab_fft_x", "psc_slab_transpose_put", "psc_slab_yblocks", "psc_green_slab", "psc_xfft_green_slab")


def load():
    """Load libpysco_b200.so (building it is __graft_entry__.build()'s / pysco_b200.build's job)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m pysco_b200.build` (nvcc, sm_100a). "
                "pysco_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _i)
        for name in _TIMED:
            setattr(lib, name, _Timed(getattr(lib, name), name))
        _lib = lib
    return _lib


class PyscoCudaError(RuntimeError):
    pass


def check(rc: int):
    if rc == 0:
        return
    msg = load().psc_last_error().decode()
    if rc == -1:
        raise ValueError(msg)
    raise PyscoCudaError(f"libpysco_b200 error {rc}: {msg}")


_cuda_ok = False


def device() -> torch.device:
    global _cuda_ok
    if not _cuda_ok:      # torch.cuda.is_available() costs microseconds on every call of a step's critical path
        if not torch.cuda.is_available():
            raise RuntimeError("pysco_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _cuda_ok = True
    return torch.device("cuda", torch.cuda.current_device())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream() -> int:
    """handle of torch's current stream on the current device (what every C-ABI call is launched on)"""
    if _raw_stream is not None:   # ~1 us instead of ~15 us for torch.cuda.current_stream()
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def launch_count() -> int:
    return int(load().psc_launch_count())


class Ctx:
    """Per-call adapter of the drop-in surface.  Arguments may be
      * torch CUDA tensors  -> used in place, results are CUDA tensors (no copies);
      * NumPy arrays        -> uploaded, results downloaded to new ndarrays, in-place arguments
                               written back (what an unmodified PySCo script passes);
      * torch CPU tensors   -> same as NumPy but results come back as pinned CPU tensors and the
                               copies are asynchronous when the source is pinned (bench.py's e2e leg).
    """

    def __init__(self):
        self.np_mode = False      # any host-side argument seen
        self.host_tensor = False  # host arguments were torch CPU tensors
        self._writeback = []

    def dev(self, a, dtype=torch.float32, inplace=False):
        if a is None:
            return None
        if isinstance(a, torch.Tensor) and a.is_cuda:
            if a.dtype != dtype or not a.is_contiguous():
                if inplace:
                    raise TypeError(f"in-place argument must be contiguous {dtype}")
                a = a.to(dtype).contiguous()
            return a
        self.np_mode = True
        if isinstance(a, torch.Tensor):
            self.host_tensor = True
            src = a
        else:
            if inplace and not isinstance(a, np.ndarray):
                raise TypeError("in-place argument must be an ndarray or a tensor")
            src = torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
        t = src.to(device(), non_blocking=True).to(dtype).contiguous()
        if inplace:
            self._writeback.append((a, t))
        return t

    def ret(self, t):
        """Return value conversion (device tensor -> host object of the kind the caller passed)."""
        if t is None or not self.np_mode:
            return t
        if self.host_tensor:
            out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            out.copy_(t, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out
        return t.cpu().numpy()

    def finish(self):
        for dst, t in self._writeback:
            if isinstance(dst, torch.Tensor):
                dst.copy_(t.reshape(dst.shape), non_blocking=True)
            else:
                np.copyto(dst, t.cpu().numpy().reshape(dst.shape))
        if self._writeback and self.host_tensor:
            torch.cuda.current_stream().synchronize()
        self._writeback = []


def empty(shape, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype=torch.float32):
    return torch.zeros(shape, dtype=dtype, device=device())


# ------------------------------------------------------------------ cuFFT plan cache (per device, N)
_plans = {}


def fft_plan(N: int):
    key = (torch.cuda.current_device(), int(N))
    if key not in _plans:
        h = _vp()
        check(load().psc_fft_plan_create(int(N), C.byref(h)))
        _plans[key] = h
    return _plans[key]


def free_plans():
    for h in _plans.values():
        load().psc_fft_plan_destroy(h)
    _plans.clear()
