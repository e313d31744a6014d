"""Mirror of pysco/morton.py (hot-path part: positions_to_keys, morton.py:42-137)."""
import torch

from . import _lib


def positions_to_keys(positions):
    """int64 Morton keys, 21 bits per axis, x most significant (morton.py:113-137)."""
    c = _lib.Ctx()
    pos = c.dev(positions)
    n = pos.shape[0]
    keys = _lib.empty((n,), torch.int64)
    _lib.check(_lib.load().psc_morton_keys(_lib.ptr(pos), n, _lib.ptr(keys), _lib.stream()))
    return c.ret(keys)
