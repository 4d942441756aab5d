#!/usr/bin/env python3
"""A/B of the subtree solve kernels at config 2: variant (0 = CTA per subtree, bulk-staged; 1 = warp per subtree) x budget."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, helpers, russell_b200 as rb
cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(1, 8192), (1, 4096), (1, 16384), (1, 32768), (0, 3072)]
coo = helpers.laplacian_2d_coo(1000)
b = np.ones(coo.nrow)
for variant, budget in cases:
    x = np.zeros(coo.nrow)
    sol = rb.SolverB200(); sol.set_option("subtree_budget", budget)
    par = rb.LinSolParams(); par.verbose = True
    sol.factorize(coo, par)
    best = 1e9
    for _ in range(5):
        sol.solve(x, b); best = min(best, sol.device_stats()["ms_sptrsv_device"])
    st = sol.device_stats()
    print("variant %d budget %6d: sptrsv %.4f ms (%.0f GB/s)  residual %.2e" % (variant, budget, best, st["sptrsv_bytes"] / best / 1e6, sol.residual(x, b)), flush=True)
    del sol
