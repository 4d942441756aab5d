#!/usr/bin/env python3
"""CPU-only view of the front tree the host analysis builds (no GPU needed): per level the number of fronts, pivot and
update-row statistics and stored entries; the separator chains (supernodes split into panels); the subtree partition the
solve phase would use.  Usage: python tests/dev/plan_stats.py [grid=1000]  (5-point Laplacian)  |  --brusselator N"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from oracle import oracle  # noqa: E402  (lives under tests/: only test infrastructure may use oracle/)

oracle.build()
lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_mf.so"))
ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
lib.oracle_plan_create.restype = ctypes.c_void_p
lib.oracle_plan_create.argtypes = [ctypes.c_int, ip, ip, dp] + [ctypes.c_int] * 5 + [ip]
lib.oracle_plan_nodes.argtypes = [ctypes.c_void_p, ip, ip, ip, ip]
lib.oracle_plan_free.argtypes = [ctypes.c_void_p]

if len(sys.argv) > 2 and sys.argv[1] == "--brusselator":
    n, ai, aj, ax, _ = helpers.brusselator_radau5_triplets(int(sys.argv[2]))
else:
    n, ai, aj, ax = helpers.laplacian_2d_triplets(int(sys.argv[1]) if len(sys.argv) > 1 else 1000)
bp, bj, bx = oracle.coo_to_csr(n, n, ai, aj, ax)
nn = ctypes.c_int(0)
h = lib.oracle_plan_create(n, bp.ctypes.data_as(ip), bj.ctypes.data_as(ip), bx.ctypes.data_as(dp), 0, 0, 2, 0, 0, ctypes.byref(nn))
N = nn.value
p, u, lev, par = (np.zeros(N, dtype=np.int32) for _ in range(4))
lib.oracle_plan_nodes(h, p.ctypes.data_as(ip), u.ctypes.data_as(ip), lev.ctypes.data_as(ip), par.ctypes.data_as(ip))
lib.oracle_plan_free(h)
ent = p.astype(np.int64) * (p + 2 * u.astype(np.int64))
print("n = %d, fronts = %d, levels = %d, stored entries = %.2f M" % (n, N, lev.max() + 1, ent.sum() / 1e6))
print("level  fronts   p(mean/max)   u(mean/max)   entries(M)")
for l in range(lev.max() + 1):
    m = lev == l
    print("%5d %7d   %5.1f /%3d   %6.1f /%4d   %8.3f" % (l, m.sum(), p[m].mean(), p[m].max(), u[m].mean(), u[m].max(), ent[m].sum() / 1e6))
# subtree partition of the solve phase (solver_b200.cu: every front f <= 96, p <= 32, <= 16384 entries per subtree)
sub, size, elig = ent.copy(), np.ones(N, dtype=np.int64), (p + u <= 96) & (p <= 32)
for v in range(N):
    if sub[v] > 16384:
        elig[v] = False
    if par[v] >= 0:
        sub[par[v]] += sub[v]
        size[par[v]] += size[v]
        if not elig[v]:
            elig[par[v]] = False
root = elig & ~np.where(par >= 0, elig[np.maximum(par, 0)], False)
print("subtrees: %d roots covering %d fronts (%.1f %%) and %.1f %% of the entries; deepest root at level %d"
      % (root.sum(), size[root].sum(), 100.0 * size[root].sum() / N, 100.0 * sub[root].sum() / ent.sum(), lev[root].max() if root.any() else -1))
