"""GPU parity of theory = mond (QUMOND: psc_box_mond_rhs between two slab solves) and theory = fr (scalaron by the FAS
cycle on ghosted slabs, psc_box_*_fr; fifth force through the slab interpolation kernel) on x-slabs against the
oracle's single-process step, on P = 1, 2, 4 virtual ranks sharing cuda:0.

The host sequencing and the per-cell kernel code are covered on the CPU tier (tests/test_slab_cpu.py,
tests/test_slab_mg_cells_cpu.py); first B200 run: the driver's round-1 GPU tier (GPUTEST_r01.json, all cases passed)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import test_slab_cpu as cpu  # noqa: E402
import test_slab_gpu as gpu  # noqa: E402

pytestmark = pytest.mark.gpu


def _case(P, solver, overrides, N=32):
    ref, ref_t = cpu._reference(N, solver, **dict(overrides))
    out = gpu._threads(P, lambda c, o: gpu._run_rank_cuda(N, c, o, 5, reorder_at=1, solver=solver, **dict(overrides)))
    cpu._check(out, ref, ref_t, P)


@pytest.mark.parametrize("P,solver,overrides", cpu.MOND_CASES)
def test_slab_cuda_mond_vs_oracle(P, solver, overrides):
    _case(P, solver, overrides)


@pytest.mark.parametrize("P,solver,overrides", cpu.FR_CASES)
def test_slab_cuda_fr_vs_oracle(P, solver, overrides):
    _case(P, solver, overrides)


def test_slab_cuda_fr_8_ranks_gathered_fas_levels():
    _case(8, "fft", dict(theory="fr", fR_n=1, aexp=0.05), N=64)
