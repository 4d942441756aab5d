#!/usr/bin/env python3
"""Size sample of the headline workload (5-point Laplacian k x k): k = 447 is the "1M-nnz" reading of BASELINE.json's metric
string (n = 199,809, nnz = 997,257), k = 1000 is configs[1], k = 2000 shows where the bandwidth goes at 4 M dof.
Usage: python tools/gpu_sizes.py [k ...]   (one line per size: device ms of the last of 3 refactorizations / sweeps)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import russell_b200 as rb
import helpers

for k in [int(a) for a in sys.argv[1:]] or [447, 1000, 2000]:
    coo = helpers.laplacian_2d_coo(k)
    b, x = np.ones(coo.nrow), np.zeros(coo.nrow)
    sol = rb.SolverB200()
    for _ in range(3):
        sol.factorize(coo)
        sol.solve(x, b)
    st = sol.device_stats()
    step = st["ms_factorize_device"] + st["ms_solve_device"]
    print(f"k={k} n={coo.nrow} nnz={5 * coo.nrow - 4 * k}: init {sol.get_ns_init() / 1e9:.2f} s, factorize {st['ms_factorize_device']:.3f} ms "
          f"({st['flops'] / st['ms_factorize_device'] / 1e9:.2f} TFLOP/s), solve {st['ms_solve_device']:.3f} ms (sweep {st['ms_sptrsv_device']:.3f} ms = "
          f"{st['sptrsv_bytes'] / st['ms_sptrsv_device'] / 1e6:.0f} GB/s), {1e3 / step:.1f} systems/s device-resident, levels {st['nlevels']:.0f}, "
          f"nnz(L+U) {st['nnz_l'] + st['nnz_u']:.3e}, residual {sol.residual(x, b):.1e}", flush=True)
    del sol
