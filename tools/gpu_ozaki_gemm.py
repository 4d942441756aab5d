#!/usr/bin/env python3
"""standalone run of the tcgen05 Schur kernel (for ncu): python tools/gpu_ozaki_gemm.py [u=4096] [k=64] [reps=2]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, russell_b200 as rb
from russell_b200._lib import p_f64, ptr
u = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lib = rb._lib.load()
rng = np.random.default_rng(1)
a = np.asfortranarray(rng.standard_normal((u, k))); b = np.asfortranarray(rng.standard_normal((u, k)))
c = np.zeros((u, u), order="F")
ms = ctypes.c_double(0.0)
for _ in range(reps):
    lib.solver_b200_ozaki_gemm(u, k, ptr(a, p_f64), ptr(b, p_f64), ptr(c, p_f64), ctypes.byref(ms))
print("u %d k %d: %.3f ms = %.2f TFLOP/s f64-equivalent" % (u, k, ms.value, 2.0 * u * u * k / ms.value / 1e9))
