#!/usr/bin/env python3
"""solve_matrix_market for the B200 backend: the reference's benchmark driver
(russell_sparse/src/bin/solve_matrix_market.rs:97-305) with `--genie b200`, printing the reference's StatsLinSol JSON
(russell_sparse/src/stats_lin_sol.rs:14-115: main / matrix / requests / output / determinant / verify / time_human /
time_nanoseconds / mumps_stats), so its numbers drop into the reference's own comparison tables (zscripts, README).

Protocol of the reference: read the matrix (LeaveAsLower for a symmetric file, like Genie::Cudss -- B200 takes
Sym::YesLower), rhs = ones (complex: 1+1i), `nrun` times { new solver, factorize, solve, VerifyLinSys }, averaged times;
bfwb62 is additionally checked against the 62 golden values (solve_matrix_market.rs:217-230, 307-372).

    python tools/solve_matrix_market.py tests/golden/matrix_market/bfwb62.mtx [-r 3] [-d] [-o Metis] [-v]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import russell_b200 as rb  # noqa: E402


def format_nanoseconds(ns):
    """russell_lab::format_nanoseconds (russell_lab/src/base/formatters.rs:60-100)"""
    ns = int(ns)
    if ns == 0:
        return "0ns"
    if ns < 1_000_000_000:
        if ns < 1_000:
            return "%dns" % ns
        if ns < 1_000_000:
            return "%sµs" % repr(ns / 1e3)
        return "%sms" % repr(ns / 1e6)
    out, value = "", ns
    hours, value = divmod(value, 3_600_000_000_000)
    if hours:
        out += "%dh" % hours
    minutes, value = divmod(value, 60_000_000_000)
    if minutes:
        out += "%dm" % minutes
    if value > 0:
        sec = value / 1e9
        out += ("%ds" % sec) if sec == int(sec) else ("%ss" % repr(sec))
    return out


def main():
    ap = argparse.ArgumentParser(description="solve_matrix_market (B200 backend), StatsLinSol JSON on stdout")
    ap.add_argument("matrix_market_file")
    ap.add_argument("-g", "--genie", default="B200")
    ap.add_argument("-o", "--ordering", default="Auto")
    ap.add_argument("-s", "--scaling", default="Auto")
    ap.add_argument("--matching-sym", default="None")
    ap.add_argument("--matching-gen", default="Auto")
    ap.add_argument("-p", "--positive-definite", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("-d", "--determinant", action="store_true")
    ap.add_argument("--hide-json", action="store_true")
    ap.add_argument("-r", "--nrun", type=int, default=1)
    opt = ap.parse_args()
    if rb.Genie.from_str(opt.genie) != rb.Genie.B200:
        raise SystemExit("this driver only knows Genie::B200 (the CPU backends belong to the reference)")

    params = rb.LinSolParams()
    params.ordering = rb.Ordering[opt.ordering]
    params.scaling = rb.Scaling[opt.scaling]
    params.positive_definite = opt.positive_definite
    params.compute_determinant = opt.determinant
    params.verbose = opt.verbose

    t0 = time.perf_counter_ns()
    coo = rb.read_matrix_market(opt.matrix_market_file, rb.MMsym.LeaveAsLower)
    t_read = time.perf_counter_ns() - t0
    is_complex = coo.values.dtype == np.complex128
    nrow, ncol, nnz, sym = coo.get_info()
    matching = opt.matching_sym if sym != rb.Sym.No else opt.matching_gen
    params.matching = rb.Matching["None_" if matching == "None" else matching]

    name = os.path.splitext(os.path.basename(opt.matrix_market_file))[0]
    times = {"initialize_array": [], "factorize_array": [], "solve_array": [], "total_ifs_array": []}
    verify, t_verify, last = None, 0, None
    stats = rb.StatsLinSol()
    rhs = np.full(nrow, 1.0 + 1.0j if is_complex else 1.0)
    x = np.zeros(nrow, dtype=rhs.dtype)
    out_of_memory = False
    for _ in range(max(1, opt.nrun)):
        solver = rb.ComplexSolverB200() if is_complex else rb.SolverB200()
        try:
            solver.factorize(coo, params)
        except rb.StrError as e:
            if any(w in str(e) for w in ("cudaMalloc", "MALLOC", "ALLOC_FAILED")):  # stats_lin_sol.rs:334-340
                out_of_memory = True
                break
            raise
        solver.solve(x, rhs, opt.verbose)
        solver.update_stats(stats)
        t0 = time.perf_counter_ns()
        v = rb.verify_from_complex(coo, x, rhs, solver) if is_complex else rb.VerifyLinSys.from_(coo, x, rhs, solver)
        t_verify = time.perf_counter_ns() - t0
        if verify is None or v.relative_error > verify.relative_error:  # max over runs (stats_lin_sol.rs max_relative_error)
            verify = v
        ti, tf, ts = solver.get_ns_init(), solver.get_ns_fact(), solver.get_ns_solve()
        times["initialize_array"].append(ti), times["factorize_array"].append(tf), times["solve_array"].append(ts)
        times["total_ifs_array"].append(ti + tf + ts)
        last = solver
    if name == "bfwb62" and not is_complex and not out_of_memory:  # solve_matrix_market.rs:217-230
        gold = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "bfwb62_x.json"))))
        if np.max(np.abs(x - gold)) > 1e-10:
            raise SystemExit("bfwb62: the solution differs from the reference's golden values")

    avg = lambda a: int(sum(a) // max(1, len(a)))
    tn = {"read_matrix": t_read, "initialize_array": times["initialize_array"], "initialize": avg(times["initialize_array"]),
          "factorize_array": times["factorize_array"], "factorize": avg(times["factorize_array"]),
          "solve_array": times["solve_array"], "solve": avg(times["solve_array"]),
          "total_ifs_array": times["total_ifs_array"], "total_ifs": avg(times["total_ifs_array"]), "verify": t_verify}
    th = {k: ([format_nanoseconds(v) for v in val] if isinstance(val, list) else format_nanoseconds(val)) for k, val in tn.items()}
    det = stats.determinant if (opt.determinant and not is_complex) else (0.0, 0.0)
    doc = {
        "main": {"platform": "Russell", "blas_lib": "none (CUDA kernels, sm_100a)", "solver": "B200", "local_sparse": False,
                 "out_of_memory": out_of_memory},
        "matrix": {"name": name, "nrow": nrow, "ncol": ncol, "nnz": nnz, "nnz_actual": nnz,
                   "complex": bool(is_complex), "symmetric": sym.name},
        "requests": {"ordering": opt.ordering, "scaling": opt.scaling, "matching": matching, "pivoting": "Auto", "mumps_num_threads": 0,
                     "positive_definite": bool(opt.positive_definite), "hybrid_memory_factor": None},
        "output": {"effective_ordering": stats.effective_ordering, "effective_scaling": stats.effective_scaling,
                   "effective_matching": stats.effective_matching, "effective_pivoting": stats.effective_pivoting,
                   "effective_mumps_num_threads": 0, "openmp_num_threads": 0, "umfpack_strategy": "Unknown",
                   "umfpack_rcond_estimate": stats.rcond_estimate},
        "determinant": {"mantissa_real": det[0], "mantissa_imag": 0.0, "base": 10.0, "exponent": det[1]},
        "verify": {"max_abs_a": verify.max_abs_a, "max_abs_ax": verify.max_abs_ax, "max_abs_diff": verify.max_abs_diff,
                   "relative_error": verify.relative_error} if verify else {},
        "time_human": th,
        "time_nanoseconds": tn,
        "mumps_stats": {"inf_norm_a": 0.0, "inf_norm_x": 0.0, "scaled_residual": 0.0, "backward_error_omega1": 0.0,
                        "backward_error_omega2": 0.0, "normalized_delta_x": 0.0, "condition_number1": 0.0, "condition_number2": 0.0},
        # not part of the reference's schema: what the device reports (fronts, flops, kernel times, residual, perturbed pivots)
        "b200_device": stats.device,
    }
    if not opt.hide_json:
        print(json.dumps(doc, indent=2, ensure_ascii=False))


if __name__ == "__main__":
    main()
