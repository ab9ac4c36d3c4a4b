"""Time bvg_tsse_draws on a c2-A-sized draw buffer ([4096 chains][5000 iterations][21 rows]); run under
`ncu --metrics gpu__time_duration.sum` for per-kernel times."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bvartools_b200._lib import check, lib

chains, iters, rows = 4096, 5000, 21
g = torch.Generator(device="cuda").manual_seed(1)
d = torch.randn((chains, iters, rows), dtype=torch.float64, device="cuda", generator=g)
out = np.empty(rows)
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    check(lib().bvg_tsse_draws(C.c_void_p(d.data_ptr()), chains, iters, rows, out.ctypes.data_as(C.POINTER(C.c_double)), None))
    print("tsse ms", (time.perf_counter() - t0) * 1e3, out[:2] * np.sqrt(chains * iters))
