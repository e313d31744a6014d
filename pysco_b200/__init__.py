"""pysco_b200 -- B200-native (sm_100a) particle-mesh gravity step behind PySCo's module API.

Drop-in surface: ``pysco_b200.run(param)`` and the flat modules ``solver``, ``mesh``,
``multigrid``, ``integration``, ``fourier``, ``laplacian``, ``cubic``, ``quartic``, ``mond``,
``morton``, ``utils`` with the reference's function names, argument meaning and error behaviour.
Arrays may be NumPy (uploaded per call, results returned as NumPy) or torch CUDA tensors (zero-copy;
results are CUDA tensors).  All arithmetic runs in hand-written CUDA kernels of libpysco_b200.so
(cuFFT for the transforms); there is no CPU fallback.
"""
__version__ = "0.1.0"

import os as _os

# CUDA loads kernels lazily by default: the first launch of every template instantiation (a float64 time step, the
# other bin table, ...) then costs milliseconds in the middle of a run -- measured as a 3 ms hiccup in the step where
# the time-step criterion changes.  Ask for eager loading unless the user has decided otherwise; it only takes effect
# if the CUDA driver has not been initialised yet.
_os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

from . import _lib  # noqa: F401
from . import cubic, fourier, integration, laplacian, mesh, mond, morton, multigrid, quartic, solver, utils  # noqa: F401,E402
from .main import run  # noqa: F401,E402
