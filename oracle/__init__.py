"""oracle -- CPU restatement of the PySCo PM hot path.  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
reference legs).  Nothing under pysco_b200/ may import this package.
"""
from . import api, host  # noqa: F401
from .api import (build, cubic, fourier, laplacian, mesh, mond, morton, num_threads, quartic,  # noqa: F401
                  set_num_threads, utils)
