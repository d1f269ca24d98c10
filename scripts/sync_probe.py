import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrr_b200 import lib
L = lib(); ctx = L.context(0)
buf = torch.zeros(2 * 1184 * 112 + 16, dtype=torch.float64, device="cuda"); out = torch.zeros(4, dtype=torch.float64, device="cuda")
f = L.dll.lrr_debug_sync_probe; f.restype = C.c_int
for mode, name in ((0, "grid.sync only"), (1, "reduce<5> (block reduce + row write + grid.sync + redundant totals)")):
    for reps in (1, 101):
        ms = C.c_float(0)
        rc = f(C.c_void_p(buf.data_ptr()), C.c_void_p(out.data_ptr()), reps, mode, C.byref(ms))
        print(name, "reps", reps, "rc", rc, "ms", round(ms.value, 4))
