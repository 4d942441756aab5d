#!/usr/bin/env python3
"""Cold-start probe: phases of solver_b200_initialize (verbose) at config-2 size for (1) a first handle (plan cache miss),
(2) a second handle of the same pattern (cache hit), (3) a handle with the plan cache off, (4) one with the host analysis
forced serial.  Usage: python tools/gpu_init_probe.py [grid]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, helpers, russell_b200 as rb

coo = helpers.laplacian_2d_coo(int(sys.argv[1]) if len(sys.argv) > 1 else 1000)
keep = []
for label, env in (("first handle", {}), ("second handle (cache hit)", {}), ("plan cache off", {"B200_PLAN_CACHE": "0"}),
                   ("plan cache off, serial analysis", {"B200_PLAN_CACHE": "0", "B200_ND_SERIAL": "1"})):
    os.environ.pop("B200_PLAN_CACHE", None), os.environ.pop("B200_ND_SERIAL", None)
    os.environ.update(env)
    print("----", label, flush=True)
    sys.stderr.flush()
    sol = rb.SolverB200(coo_boundary=False)
    par = rb.LinSolParams(); par.verbose = True
    t = time.time(); sol.factorize(coo, par)
    print("factorize wall %.3f s, initialize %.3f s" % (time.time() - t, sol.get_ns_init() / 1e9), flush=True)
    keep.append(sol)
