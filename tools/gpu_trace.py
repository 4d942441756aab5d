#!/usr/bin/env python3
"""Per-item timeline of the persistent top-of-tree SpTRSV kernels (solver_b200_debug_trace): for every tree level of the
persistent region, when its items started, when their dependencies were satisfied and when they finished.
Usage: python tools/gpu_trace.py [grid=1000] > gpurun_out/trace.txt"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import russell_b200 as rb  # noqa: E402
from russell_b200._lib import p_i32, ptr  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
coo = helpers.laplacian_2d_coo(k)
n = coo.nrow
sol = rb.SolverB200()
sol.set_option("trace", 1)
for kv in sys.argv[2:]:
    key, val = kv.split("=")
    sol.set_option(key, float(val))
sol.factorize(coo)
b, x = np.ones(n), np.zeros(n)
for _ in range(4):
    sol.solve(x, b)
st = sol.device_stats()
lib = sol._lib
nit = lib.solver_b200_debug_trace(sol.solver, None, None, 0)
out = np.zeros(8 * nit, dtype=np.uint64)
desc = np.zeros(4 * nit, dtype=np.int32)
lib.solver_b200_debug_trace(sol.solver, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), ptr(desc, p_i32), nit)
desc = desc.reshape(nit, 4)
print("items", nit, "sptrsv_ms", st["ms_sptrsv_device"], "levels", st["nlevels"])
for name, tr in (("fwd", out[: 4 * nit].reshape(nit, 4)), ("bwd", out[4 * nit:].reshape(nit, 4))):
    tr = tr.astype(np.int64)
    t0 = tr[:, 0][tr[:, 0] > 0].min()
    start, dep, end = (tr[:, 0] - t0) / 1e3, (tr[:, 1] - t0) / 1e3, (tr[:, 2] - t0) / 1e3
    mid = (tr[:, 3] - t0) / 1e3  # (forward LL kernel: all dependent values in shared memory)
    print("==", name, "kernel span %.1f us" % (end.max() - start.min()))
    print("level nfronts nitems  first_start  first_dep  last_dep  last_end | med(dep->end) max(dep->end) | level_latency")
    levels = np.unique(desc[:, 1])
    order = levels if name == "fwd" else levels[::-1]
    prev_end = None
    for lv in order:
        m = desc[:, 1] == lv
        nf = len(np.unique(desc[m, 0]))
        work = end[m] - dep[m]
        lat = (end[m].max() - prev_end) if prev_end is not None else float("nan")
        gath = np.median(mid[m] - dep[m]) if (tr[:, 3][m] > 0).all() else float("nan")
        print("%4d %6d %6d   %9.1f %9.1f %9.1f %9.1f | %6.2f %6.2f | %6.2f | dep->gathered %5.2f" % (lv, nf, m.sum(), start[m].min(), dep[m].min(), dep[m].max(), end[m].max(),
                                                                            np.median(work), work.max(), lat, gath))
        prev_end = end[m].max()
