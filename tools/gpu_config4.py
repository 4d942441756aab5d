#!/usr/bin/env python3
"""BASELINE.json configs[3] at full size on one GPU: the two Newton matrices Radau5 factorizes for the Brusselator PDE
with npoint = 500 (ndim = 500,000; real K and complex K), a geometric sweep of step sizes h (x1.5 per refactorization,
SURVEY 8d) -- every step refactorizes both systems and solves each once.  Prints one JSON object (also written to
gpurun_out/config4.json).  Usage: python tools/gpu_config4.py [npoint] [nsteps]"""
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

npoint = int(sys.argv[1]) if len(sys.argv) > 1 else 500
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ndim, ai, aj, kr, kc = helpers.brusselator_radau5_triplets(npoint, h=1e-4)
rcoo = rb.CooMatrix.from_triplets(ndim, ndim, ai, aj, kr)
ccoo = rb.ComplexCooMatrix.from_triplets(ndim, ndim, ai, aj, kc)
rsol, csol = rb.SolverB200(), rb.ComplexSolverB200()
b = np.ones(ndim)
bz = np.ones(ndim, dtype=np.complex128) * (1.0 + 0.5j)
x, z = np.zeros(ndim), np.zeros(ndim, dtype=np.complex128)
out = {"npoint": npoint, "ndim": ndim, "nnz_coo": int(len(ai)), "steps": []}
h = 1e-4
for it in range(nsteps):
    _, _, _, kr, kc = helpers.brusselator_radau5_triplets(npoint, h=h)
    rcoo.values[:] = kr
    ccoo.values[:] = kc
    t0 = time.perf_counter()
    rsol.factorize(rcoo)
    t1 = time.perf_counter()
    rsol.solve(x, b)
    t2 = time.perf_counter()
    csol.factorize(ccoo)
    t3 = time.perf_counter()
    csol.solve(z, bz)
    t4 = time.perf_counter()
    rs, cs = rsol.device_stats(), csol.device_stats()
    out["steps"].append({
        "h": h,
        "real": {"fact_wall_ms": (t1 - t0) * 1e3, "solve_wall_ms": (t2 - t1) * 1e3, "fact_dev_ms": rs["ms_factorize_device"],
                 "solve_dev_ms": rs["ms_solve_device"], "sptrsv_ms": rs["ms_sptrsv_device"], "rel_residual": rsol.residual(x, b),
                 "refine": rs["last_refine_steps"], "perturbed": rs["n_perturbed"]},
        "complex": {"fact_wall_ms": (t3 - t2) * 1e3, "solve_wall_ms": (t4 - t3) * 1e3, "fact_dev_ms": cs["ms_factorize_device"],
                    "solve_dev_ms": cs["ms_solve_device"], "sptrsv_ms": cs["ms_sptrsv_device"], "rel_residual": csol.residual(z, bz),
                    "refine": cs["last_refine_steps"], "perturbed": cs["n_perturbed"]},
    })
    if it == 0:
        out["init_wall_s"] = {"real": rsol.get_ns_init() / 1e9, "complex": csol.get_ns_init() / 1e9}
        for nm, st in (("real", rs), ("complex", cs)):
            out[nm + "_symbolic"] = {k: st[k] for k in ("nnodes", "nlevels", "nnz_l", "nnz_u", "flops", "max_front", "fac_bytes", "cb_bytes", "sptrsv_bytes")}
    h *= 1.5
last = out["steps"][-1]
for nm in ("real", "complex"):
    f = out[nm + "_symbolic"]["flops"]
    out[nm + "_fact_tflops"] = f / (last[nm]["fact_dev_ms"] * 1e-3) / 1e12
    out[nm + "_sptrsv_gbs"] = out[nm + "_symbolic"]["sptrsv_bytes"] / (last[nm]["sptrsv_ms"] * 1e-3) / 1e9
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "config4_n%d.json" % npoint), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
