import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from pysco_b200 import _lib, utils
N = 512
lib = _lib.load(); raw = C.CDLL(_lib.LIB_PATH)
raw.psc_interp_kick4.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
pos_lex, vel = bench.synthetic_ics_device(N)
pos_mor = utils.reorder_particles(pos_lex)
force4 = torch.randn((N, N, N, 4), device="cuda")
acc = torch.empty_like(pos_mor); mx = torch.zeros(2, device="cuda")
for p in (pos_mor, pos_lex, pos_mor, pos_lex):
    raw.psc_interp_kick4(force4.data_ptr(), p.data_ptr(), vel.data_ptr(), acc.data_ptr(), p.shape[0], N, 2, 0.01, mx.data_ptr(), None)
torch.cuda.synchronize(); print("done")
