import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, helpers, russell_b200 as rb
coo = helpers.laplacian_2d_coo(1000)
sol = rb.SolverB200(coo_boundary=False)
par = rb.LinSolParams(); par.verbose = True
t=time.time(); sol.factorize(coo, par); print("factorize wall", time.time()-t, "init ns", sol.get_ns_init()/1e9)
