#!/usr/bin/env python3
"""BASELINE.json configs[2] (af_shell10, 1.5M x 1.5M, 52M nnz) is not in the tree and there is no network; this runs the
stand-in SURVEY 8d names for it -- 27-point 3D Laplacian on a k^3 grid (k = 115: n = 1,520,875, 40.4 M nonzeros) with the
index-seeded skew perturbation -- through the general (full-pattern) path, like `solve_matrix_market` runs af_shell10
with MakeItFull (russell_sparse/src/bin/solve_matrix_market.rs:97-305).  If a file af_shell10.mtx is given, it is read
instead.  The residual is evaluated on the HOST (scipy).  Usage: python tools/gpu_config3.py [k | path.mtx] [reps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import russell_b200 as rb  # noqa: E402

arg = sys.argv[1] if len(sys.argv) > 1 else "115"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t0 = time.perf_counter()
if arg.endswith(".mtx"):
    coo = rb.read_matrix_market(arg, rb.MMsym.MakeItFull)
    n, ai, aj, ax = coo.nrow, coo.indices_i[: coo.nnz], coo.indices_j[: coo.nnz], coo.values[: coo.nnz]
    name = os.path.basename(arg)
else:
    k = int(arg)
    n, ai, aj, ax = helpers.laplacian_3d_27pt_triplets(k, skew=1e-3)
    coo = rb.CooMatrix.from_triplets(n, n, ai, aj, ax, rb.Sym.No)
    name = "standin_27pt_%d^3_skew1e-3" % k
t_build = time.perf_counter() - t0
b = np.ones(n)  # protocol of solve_matrix_market.rs:179
x = np.zeros(n)
sol = rb.SolverB200()
out = {"matrix": name, "n": int(n), "nnz": int(len(ax)), "build_s": t_build, "runs": []}
for r in range(reps):
    t0 = time.perf_counter()
    sol.factorize(coo)
    t1 = time.perf_counter()
    sol.solve(x, b)
    t2 = time.perf_counter()
    st = sol.device_stats()
    out["runs"].append({"factorize_wall_s": t1 - t0, "solve_wall_s": t2 - t1, "factorize_dev_ms": st["ms_factorize_device"],
                        "solve_dev_ms": st["ms_solve_device"], "sptrsv_ms": st["ms_sptrsv_device"], "refine": st["last_refine_steps"],
                        "rel_residual_device": st["last_rel_residual"]})
    if r == 0:
        out["initialize_s"] = sol.get_ns_init() / 1e9
        out["symbolic"] = {k2: st[k2] for k2 in ("nnodes", "nlevels", "nnz_l", "nnz_u", "flops", "max_front", "fac_bytes", "cb_bytes",
                                                 "sptrsv_bytes", "t_order_s", "t_symbolic_s")}
out["rel_residual_host"] = helpers.host_rel_residual(n, ai, aj, ax, x, b)
vf = rb.VerifyLinSys.from_(coo, x, b, sol)
out["verify_relative_error"] = vf.relative_error
last = out["runs"][-1]
out["factorize_tflops"] = out["symbolic"]["flops"] / (last["factorize_dev_ms"] * 1e-3) / 1e12
out["sptrsv_gbs"] = out["symbolic"]["sptrsv_bytes"] / (last["sptrsv_ms"] * 1e-3) / 1e9
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "config3_%s.json" % name.replace("^", "")), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
