#!/usr/bin/env python3
"""one factorization + a few solves of config 2 (or k given): the process to put under ncu for the SpTRSV kernels"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, helpers, russell_b200 as rb
k = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nsolve = int(sys.argv[2]) if len(sys.argv) > 2 else 3
coo = helpers.laplacian_2d_coo(k)
b = np.ones(coo.nrow); x = np.zeros(coo.nrow)
sol = rb.SolverB200()
for kv in sys.argv[3:]:
    key, val = kv.split("="); sol.set_option(key, float(val))
sol.factorize(coo)
for _ in range(nsolve): sol.solve(x, b)
st = sol.device_stats()
print("sptrsv %.4f ms, %.0f GB/s, residual %.2e" % (st["ms_sptrsv_device"], st["sptrsv_bytes"] / st["ms_sptrsv_device"] / 1e6, sol.residual(x, b)))
