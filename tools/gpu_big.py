#!/usr/bin/env python3
"""Large-size GPU run: 2D Laplacians through the C-ABI; --profile does ONE factorize+solve without graphs (for ncu)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import russell_b200 as rb
import helpers

def run(k, lower, opts, reps=3):
    coo = helpers.laplacian_2d_coo(k, lower=lower)
    b = np.ones(coo.nrow)
    sol = rb.SolverB200()
    for kk, v in opts.items(): sol.set_option(kk, v)
    x = np.zeros(coo.nrow)
    for r in range(reps):
        t0 = time.time(); sol.factorize(coo); t1 = time.time(); sol.solve(x, b); t2 = time.time()
        st = sol.device_stats()
        print(f"[lap{k}{'L' if lower else ''} {opts} rep{r}] init={sol.get_ns_init()/1e6:.0f}ms fact={1e3*(t1-t0):.2f}ms (dev {st['ms_factorize_device']:.3f}) "
              f"solve={1e3*(t2-t1):.2f}ms (dev {st['ms_solve_device']:.3f} sptrsv {st['ms_sptrsv_device']:.3f} spmv {st['ms_spmv_device']:.4f}) "
              f"resid={st['last_rel_residual']:.2e} refine={st['last_refine_steps']:.0f} launches={st['launches_factorize']:.0f}/{st['launches_solve']:.0f} "
              f"levels={st['nlevels']:.0f} nnzLU={st['nnz_l']+st['nnz_u']:.3e} flops={st['flops']:.3e}", flush=True)
    gbs = st['sptrsv_bytes'] / (st['ms_sptrsv_device'] * 1e-3) / 1e9
    print(f"   sptrsv {gbs:.1f} GB/s ; spmv {st['spmv_bytes']/(st['ms_spmv_device']*1e-3)/1e9:.1f} GB/s ; factor {st['flops']/(st['ms_factorize_device']*1e-3)/1e12:.3f} TFLOP/s", flush=True)
    return sol

if __name__ == "__main__":
    if "--sweep" in sys.argv:
        # tuning sweep over a runtime option at config-2 size
        key = sys.argv[sys.argv.index("--sweep") + 1]
        for val in sys.argv[sys.argv.index("--sweep") + 2:]:
            run(1000, False, {key: float(val)}, reps=2)
    elif "--profile" in sys.argv:
        k = int(sys.argv[sys.argv.index("--profile") + 1])
        run(k, False, {"use_graph": 0}, reps=1)
    else:
        for k in (300, 1000):
            run(k, False, {"schur_variant": 1, "use_graph": 1})
            run(k, False, {"schur_variant": 0, "use_graph": 0}, reps=2)
        run(1000, True, {})
        run(2000, False, {}, reps=2)
