#!/usr/bin/env python3
"""Option sweeps at config-2 size: python tools/gpu_sweep.py key=v1,v2,... [key2=...]  (one line per setting:
device ms of the last of 3 refactorizations and of the solve sweep)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import russell_b200 as rb
import helpers

coo = helpers.laplacian_2d_coo(1000)
b = np.ones(coo.nrow)
x = np.zeros(coo.nrow)
settings = []
for arg in sys.argv[1:]:
    if "+" in arg:  # a combination: key=v+key2=v2
        settings.append([kv.split("=") for kv in arg.split("+")])
    else:
        key, vals = arg.split("=")
        settings += [[(key, v)] for v in vals.split(",")]
for combo in settings:
    if True:
        sol = rb.SolverB200()
        for key, v in combo:
            sol.set_option(key, float(v))
        key, v = "+".join(k for k, _ in combo), "+".join(x for _, x in combo)
        best_f, best_s = 1e9, 1e9
        for r in range(4):
            sol.factorize(coo)
            sol.solve(x, b)
            st = sol.device_stats()
            if r > 0:
                best_f, best_s = min(best_f, st["ms_factorize_device"]), min(best_s, st["ms_sptrsv_device"])
        print(f"{key}={v}: fact {best_f:.3f} ms  sweep {best_s:.3f} ms  sum {best_f + best_s:.3f}  levels {st['nlevels']:.0f} nnzLU {st['nnz_l'] + st['nnz_u']:.3e} "
              f"flops {st['flops']:.3e} resid {st['last_rel_residual']:.1e}", flush=True)
        del sol
