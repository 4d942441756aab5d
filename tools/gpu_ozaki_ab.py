#!/usr/bin/env python3
"""A/B of the Schur-complement kernels on the config-3 stand-in (27-point 3D grid, big separator fronts):
schur_variant 1 = f64 DMMA (mma.sync m8n8k4) vs 2 = tcgen05 int8 Ozaki kernel for fronts with u >= ozaki_min_u.
Also times the standalone GEMM entry.  Usage: python tools/gpu_ozaki_ab.py [k=64]"""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, helpers, russell_b200 as rb
from russell_b200._lib import p_f64, ptr

k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
out = {"grid": k, "cases": [], "gemm": []}
lib = rb._lib.load()
for u, kk in ((4096, 64), (8192, 64), (8192, 128), (8192, 256)):
    rng = np.random.default_rng(1)
    a = np.asfortranarray(rng.standard_normal((u, kk))); b = np.asfortranarray(rng.standard_normal((u, kk)))
    c = np.zeros((u, u), order="F")
    ms = ctypes.c_double(0.0)
    for _ in range(2):
        lib.solver_b200_ozaki_gemm(u, kk, ptr(a, p_f64), ptr(b, p_f64), ptr(c, p_f64), ctypes.byref(ms))
    out["gemm"].append({"u": u, "k": kk, "ms_split_plus_gemm": ms.value, "tflops_f64_equivalent": 2.0 * u * u * kk / (ms.value * 1e-3) / 1e12})
    print(out["gemm"][-1], flush=True)
n, ai, aj, ax = helpers.laplacian_3d_27pt_triplets(k, skew=1e-3)
coo = rb.CooMatrix.from_triplets(n, n, ai, aj, ax, rb.Sym.No)
b = np.ones(n)
for name, opts in (("dmma_fused", {"schur_variant": 1}), ("dmma_unfused", {"schur_variant": 1, "fuse_chain": 0}),
                   ("ozaki_1024", {"schur_variant": 2, "ozaki_min_u": 1024}), ("ozaki_2048", {"schur_variant": 2, "ozaki_min_u": 2048}),
                   ("ozaki_512", {"schur_variant": 2, "ozaki_min_u": 512})):
    sol = rb.SolverB200()
    for kk, v in opts.items(): sol.set_option(kk, v)
    x = np.zeros(n)
    sol.factorize(coo); sol.factorize(coo); sol.factorize(coo)
    sol.solve(x, b)
    st = sol.device_stats()
    res = helpers.host_rel_residual(n, ai, aj, ax, x, b)
    out["cases"].append({"name": name, "opts": opts, "factorize_ms": st["ms_factorize_device"], "tflops": st["flops"] / st["ms_factorize_device"] / 1e9,
                         "rel_residual_host": res, "refine_steps": st["last_refine_steps"], "launches": st["launches_factorize"]})
    print(out["cases"][-1], flush=True)
    del sol
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ozaki_ab_%d.json" % k), "w"), indent=1)
