"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL / NVLink).

Round-1 decomposition: PARTICLE-PARALLEL, MESH-REPLICATED.  Every rank owns a contiguous index range of
the particle arrays and a full copy of the N^3 mesh.  One step exchanges
  * one all-reduce(sum) of the deposited density grid  (the only data-path collective), and
  * one all-reduce(max) of two floats (max|a|, max|v| for the next time step);
everything else (kick/drift, deposit, FFT solve, gradient, interpolation) is rank-local.  The particle
kernels -- 75% of the single-GPU step -- scale with 1/P; the FFT solve is replicated.  This covers
meshes that fit one GPU (<= 1024^3); the x-slab decomposition with ghost-plane exchange, particle
migration and a transposed FFT that 2048^3 needs is the next row and is laid out in
DESIGN.md.  The reference has no distributed path at all (README.md:49).

The helpers work on CPU tensors with the gloo backend too (host-logic tests, world_size 2).
"""
import os

import torch
import torch.distributed as dist


def is_active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size() -> int:
    return dist.get_world_size() if is_active() else 1


def rank() -> int:
    return dist.get_rank() if is_active() else 0


def init_from_env(backend: str = None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT).  No-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or (dist.is_available() and dist.is_initialized()):
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)


def local_range(n_global: int, r: int = None, w: int = None):
    """Contiguous index range [lo, hi) of the particle arrays owned by rank r (balanced to +-1)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, rem = divmod(int(n_global), w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (density grid after the local deposits)."""
    if is_active():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_max_(t: torch.Tensor) -> torch.Tensor:
    """In-place max over ranks (max|a|, max|v|; residual / timing maxima)."""
    if is_active():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t
