#!/usr/bin/env python3
"""First-contact GPU diagnostics: runs the C-ABI solver on small and large systems and compares every stage
with the CPU checkers.  Prints one line per check; never raises, so that one gpurun call reports everything."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import russell_b200 as rb
from russell_b200 import _lib
from russell_b200._lib import ptr, p_f64, p_i32
from oracle import oracle
import helpers

def relerr(a, b):
    d = np.abs(a - b).max(); s = max(np.abs(b).max(), 1e-300)
    return d / s

def compare_factors(coo, opts, label):
    """GPU factor panels vs the scalar walk of the same plan"""
    try:
        sol = rb.SolverB200()
        for k, v in opts.items(): sol.set_option(k, v)
        sol.factorize(coo)
        st = sol.device_stats()
        lib = _lib.load()
        nfac = int(st["fac_bytes"] / 8)
        csr = rb.CsrMatrix.from_coo(coo)
        sym_lower = coo.symmetric == rb.Sym.YesLower
        h = oracle.MfHandle(csr.nrow, csr.pointers, csr.indices, csr.values[:csr.nnz], sym_lower=sym_lower, ordering=0, matching=2,
                            panel_width=int(opts.get("panel_width", 0)), nd_leaf=int(opts.get("nd_leaf", 0)))
        hfac, hdinv, hperm = h.factors()
        fac = np.zeros(h.fac_size); dinv = np.zeros(h.dinv_size); lperm = np.zeros(h.n, dtype=np.int32)
        rc = lib.solver_b200_debug_copy_factors(sol.solver, ptr(fac, p_f64), len(fac), ptr(dinv, p_f64), len(dinv), ptr(lperm, p_i32), len(lperm))
        same_perm = bool(np.array_equal(lperm, hperm))
        print(f"[factors {label}] rc={rc} sizes gpu={nfac} host={h.fac_size} perm_equal={same_perm} "
              f"fac_relerr={relerr(fac, hfac):.3e} dinv_relerr={relerr(dinv, hdinv):.3e} perturbed gpu={st['n_perturbed']} host={h.n_perturbed}")
        if relerr(fac, hfac) > 1e-8:
            bad = np.argmax(np.abs(fac - hfac)); print("   first big diff at", bad, fac[bad], hfac[bad], " nonfinite:", np.count_nonzero(~np.isfinite(fac)))
    except Exception as e:
        print(f"[factors {label}] EXCEPTION {e!r}"); traceback.print_exc()

def solve_check(coo, b, label, opts=None, xref=None, params=None):
    try:
        sol = rb.SolverB200()
        for k, v in (opts or {}).items(): sol.set_option(k, v)
        t0 = time.time(); sol.factorize(coo, params); t1 = time.time()
        x = np.zeros(coo.nrow); sol.solve(x, b); t2 = time.time()
        st = sol.device_stats()
        res = sol.residual(x, b)
        msg = (f"[solve {label}] n={coo.nrow} init={sol.get_ns_init()/1e6:.1f}ms fact={sol.get_ns_fact()/1e6:.2f}ms (dev {st['ms_factorize_device']:.3f}) "
               f"solve={sol.get_ns_solve()/1e6:.2f}ms (dev {st['ms_solve_device']:.3f}, sptrsv {st['ms_sptrsv_device']:.3f}, spmv {st['ms_spmv_device']:.4f}) "
               f"resid={res:.2e} last={st['last_rel_residual']:.2e} refine={st['last_refine_steps']:.0f} perturbed={st['n_perturbed']:.0f} "
               f"levels={st['nlevels']:.0f} nodes={st['nnodes']:.0f} nnzLU={st['nnz_l']+st['nnz_u']:.3e} launches={st['launches_factorize']:.0f}/{st['launches_solve']:.0f}")
        if xref is not None: msg += f" xerr={relerr(x, xref):.2e}"
        print(msg)
        # second factorize+solve (graph replay path)
        t0 = time.time(); sol.factorize(coo); t1 = time.time(); sol.solve(x, b); t2 = time.time()
        st = sol.device_stats()
        print(f"   again: fact={1e3*(t1-t0):.2f}ms (dev {st['ms_factorize_device']:.3f}) solve={1e3*(t2-t1):.2f}ms (dev {st['ms_solve_device']:.3f}, sptrsv {st['ms_sptrsv_device']:.3f}) resid={sol.residual(x, b):.2e}")
        return sol, x
    except Exception as e:
        print(f"[solve {label}] EXCEPTION {e!r}"); traceback.print_exc()
        return None, None

def main():
    print("version:", _lib.load().solver_b200_version())
    # 1. reference KATs
    S = helpers.load_samples()
    kats = [("umfpack_unsymmetric_5x5", [8, 45, -3, 3, 19], [1, 2, 3, 4, 5]),
            ("mkl_symmetric_5x5_full", [1, 2, 3, 4, 5], [-979 / 3, 983, 1961 / 12, 398, 123 / 2]),
            ("mkl_symmetric_5x5_lower(true,true)", [1, 2, 3, 4, 5], [-979 / 3, 983, 1961 / 12, 398, 123 / 2]),
            ("mkl_positive_definite_5x5_lower", [1, 2, 3, 4, 5], [-979 / 3, 983, 1961 / 12, 398, 123 / 2])]
    for name, b, xc in kats:
        coo, _ = helpers.sample_coo(name)
        solve_check(coo, np.array(b, float), name, xref=np.array(xc, float))
    # bfwb62
    coo = rb.read_matrix_market(helpers.mm_path("bfwb62.mtx"), rb.MMsym.LeaveAsLower)
    solve_check(coo, np.ones(62), "bfwb62-lower", xref=helpers.load_bfwb62_x())
    coo = rb.read_matrix_market(helpers.mm_path("bfwb62.mtx"), rb.MMsym.MakeItFull)
    solve_check(coo, np.ones(62), "bfwb62-full", xref=helpers.load_bfwb62_x())
    # 2. factor comparison vs host walk, both schur variants
    small = "--small" in sys.argv
    for k in ((6, 30) if small else (6, 30, 100)):
        for var in (0, 1):
            compare_factors(helpers.laplacian_2d_coo(k), {"schur_variant": var, "use_graph": 0}, f"lap{k}-v{var}")
    n, ai, aj, ax = helpers.convection_diffusion_triplets(60)
    compare_factors(rb.CooMatrix.from_triplets(n, n, ai, aj, ax), {"schur_variant": 1}, "convdiff60")
    compare_factors(helpers.laplacian_2d_coo(60, lower=True), {"schur_variant": 1}, "lap60-lower")
    n, ai, aj, ax = helpers.saddle_point_triplets(20)
    compare_factors(rb.CooMatrix.from_triplets(n, n, ai, aj, ax), {"schur_variant": 1}, "saddle20")
    if small:
        return
    # 3. solves at growing size
    for k in (100, 300, 1000):
        coo = helpers.laplacian_2d_coo(k)
        b = np.ones(coo.nrow)
        for var, graph in ((1, 1), (0, 0)):
            solve_check(coo, b, f"lap{k}-v{var}-g{graph}", opts={"schur_variant": var, "use_graph": graph})
    coo = helpers.laplacian_2d_coo(1000, lower=True)
    solve_check(coo, np.ones(coo.nrow), "lap1000-lower")

if __name__ == "__main__":
    main()
