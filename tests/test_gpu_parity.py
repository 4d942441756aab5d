"""GPU parity tests (run with `-m gpu` on the B200 box): every call goes through the C ABI of
libsolver_b200.so via the russell_b200 mirror; the oracle (oracle/) is only the checker.

Bars:  integer/index work (local pivot permutations) bit-exact against the scalar walk of the same plan;
       f64 results within the tolerance the reference's own tests use (1e-14 / 1e-10, cited per test) and
       ||b - A x|| / ||b|| <= 1e-10 (north star) at every size.
"""
import threading

import numpy as np
import pytest

import helpers
import russell_b200 as rb
from oracle import oracle
from test_oracle import KATS, NEWTON_REF, newton_jacobian_triplets, newton_residual, run_newton

pytestmark = pytest.mark.gpu
TOL_RESIDUAL = 1e-10  # BASELINE.json north_star


def solve_through_abi(coo, b, params=None, opts=None):
    sol = rb.SolverB200()
    for k, v in (opts or {}).items():
        sol.set_option(k, v)
    sol.factorize(coo, params)
    x = np.zeros(coo.nrow)
    sol.solve(x, np.asarray(b, dtype=float))
    return sol, x


# ---- the reference's known-answer tests, through the C ABI -------------------------------------------------
@pytest.mark.parametrize("name,rhs,xc,tol,src", KATS)
def test_known_answers(name, rhs, xc, tol, src):
    coo, _ = helpers.sample_coo(name)
    sol, x = solve_through_abi(coo, rhs)
    assert np.max(np.abs(x - np.array(xc))) <= tol * max(1.0, np.max(np.abs(xc))), src
    # calling solve again works (solver_umfpack.rs:673-675)
    x2 = np.zeros(5)
    sol.solve(x2, np.array(rhs))
    assert np.array_equal(x, x2)  # deterministic kernels: bit-identical
    stats = rb.StatsLinSol()
    sol.update_stats(stats)
    assert stats.solver == "B200" and len(stats.factorize_array) == 1
    assert sol.get_ns_init() > 0 and sol.get_ns_fact() > 0 and sol.get_ns_solve() > 0


def test_unsymmetric_5x5_default_params_do_not_inherit_the_cudss_weakness():
    # solver_cudss.rs:664-671 documents x[3] = 4.00124 with default parameters on zero diagonals; we must get 4.0
    coo, _ = helpers.sample_coo("umfpack_unsymmetric_5x5")
    sol, x = solve_through_abi(coo, [8.0, 45.0, -3.0, 3.0, 19.0], rb.LinSolParams())
    assert np.max(np.abs(x - np.arange(1.0, 6.0))) <= 1e-14
    assert sol.effective_matching == 5  # the zero diagonals triggered the max-product matching
    # explicit requests from the cuDSS tests (solver_cudss.rs:703-797)
    for setup in ("matching", "colamd", "pivot"):
        par = rb.LinSolParams()
        if setup == "matching":
            par.matching = rb.Matching.Auto
        elif setup == "colamd":
            par.ordering = rb.Ordering.Colamd
        else:
            par.pivot_epsilon, par.refinement_nstep = 1e-12, 1
        _, x = solve_through_abi(coo, [8.0, 45.0, -3.0, 3.0, 19.0], par)
        assert np.max(np.abs(x - np.arange(1.0, 6.0))) <= 1e-12


def test_cudss_example_systems():
    # solver_cudss.rs:827-857 (SPD lower) and :860-892 (unsymmetric), tol 1e-10
    coo = rb.CooMatrix(5, 5, 8, rb.Sym.YesLower)
    for i, j, v in [(0, 0, 4.0), (1, 1, 3.0), (2, 0, 1.0), (2, 1, 2.0), (2, 2, 5.0), (3, 3, 1.0), (4, 2, 1.0), (4, 4, 2.0)]:
        coo.put(i, j, v)
    par = rb.LinSolParams()
    par.positive_definite = True
    _, x = solve_through_abi(coo, [7.0, 12.0, 25.0, 4.0, 13.0], par)
    assert np.max(np.abs(x - np.arange(1.0, 6.0))) <= 1e-10
    par.hybrid_memory_factor = 0.5  # accepted and ignored (solver_cudss.rs:895-925)
    _, x = solve_through_abi(coo, [7.0, 12.0, 25.0, 4.0, 13.0], par)
    assert np.max(np.abs(x - np.arange(1.0, 6.0))) <= 1e-10
    coo = rb.CooMatrix(5, 5, 13)
    for i, j, v in [(0, 0, 5.0), (0, 1, 1.0), (0, 4, 3.0), (1, 0, 2.0), (1, 1, 6.0), (1, 3, 4.0), (2, 2, 7.0), (2, 3, 2.0),
                    (3, 1, 1.0), (3, 2, 3.0), (3, 3, 8.0), (4, 0, 4.0), (4, 4, 9.0)]:
        coo.put(i, j, v)
    _, x = solve_through_abi(coo, [22.0, 30.0, 29.0, 43.0, 49.0])
    assert np.max(np.abs(x - np.arange(1.0, 6.0))) <= 1e-10


def test_lin_solver_compute_lower_and_full():
    # lin_solver.rs:241-270
    xc = np.array([-979.0 / 3.0, 983.0, 1961.0 / 12.0, 398.0, 123.0 / 2.0])
    for name in ("mkl_symmetric_5x5_lower(true,false)", "mkl_symmetric_5x5_full"):
        coo, _ = helpers.sample_coo(name)
        x = np.zeros(5)
        rb.LinSolver.compute(rb.Genie.B200, x, coo, np.arange(1.0, 6.0))
        assert np.max(np.abs(x - xc)) <= 1e-10 * np.max(np.abs(xc))


def test_bfwb62_golden_solution():
    # solve_matrix_market.rs:217-230: 62 golden values, abs tol 1e-10 (|x| ~ 1e5); both symmetric handlings
    xg = helpers.load_bfwb62_x()
    for handling in (rb.MMsym.LeaveAsLower, rb.MMsym.MakeItFull):
        coo = rb.read_matrix_market(helpers.mm_path("bfwb62.mtx"), handling)
        sol, x = solve_through_abi(coo, np.ones(62))
        assert np.max(np.abs(x - xg)) <= 1e-10
        v = rb.VerifyLinSys.from_(coo, x, np.ones(62), sol)
        assert v.relative_error <= 1e-14 and v.max_abs_a > 0
        ov = oracle.verify(62, coo.indices_i[: coo.nnz], coo.indices_j[: coo.nnz], coo.values[: coo.nnz], x, np.ones(62),
                           mirror=coo.symmetric.triangular())
        assert abs(v.max_abs_ax - ov["max_abs_ax"]) <= 1e-12 * ov["max_abs_ax"]


def test_diagonal_10x10():
    # tests/test_umfpack.rs, tol 1e-14
    n = 10
    coo = rb.CooMatrix(n, n, n)
    xc = np.arange(n, dtype=float)
    rhs = np.zeros(n)
    for k in range(n):
        akk = 10.0 + k * (n / 10.0)
        coo.put(k, k, akk)
        rhs[k] = akk * xc[k]
    _, x = solve_through_abi(coo, rhs)
    assert np.max(np.abs(x - xc)) <= 1e-14


def test_newton_refactorization_loop():
    # tests/test_nonlinear_system.rs:60-129: one solver, repeated factorize(&jj, None)/solve, exactly 5 iterations
    solver = rb.LinSolver(rb.Genie.B200)
    jj = rb.CooMatrix(4, 4, 16)

    def solve(trip, r, it):
        jj.reset()
        for i, j, v in zip(*trip):
            jj.put(i, j, v)
        solver.actual.factorize(jj, None)
        mdu = np.zeros(4)
        solver.actual.solve(mdu, r)
        return mdu

    assert run_newton(solve) == 5


# ---- error paths (solver_umfpack.rs:534-582,609-657; solver_cudss.rs:577-624) -------------------------------------
def test_factorize_handles_errors():
    solver = rb.SolverB200()
    assert not solver.factorized
    coo, _ = helpers.sample_coo("rectangular_1x7")
    with pytest.raises(rb.StrError, match="the matrix must be square"):
        solver.factorize(coo)
    with pytest.raises(rb.StrError, match="the COO matrix must have at least one non-zero value"):
        solver.factorize(rb.CooMatrix(1, 1, 1))
    coo, _ = helpers.sample_coo("mkl_symmetric_5x5_upper(true,false)")
    with pytest.raises(rb.StrError, match="B200 requires Sym::YesLower or Sym::YesFull for symmetric matrices"):
        solver.factorize(coo)
    coo = rb.CooMatrix(2, 2, 2)
    coo.put(0, 0, 1.0), coo.put(1, 1, 2.0)
    solver.factorize(coo)
    c2 = rb.CooMatrix(2, 2, 2, rb.Sym.YesFull)
    c2.put(0, 0, 1.0), c2.put(1, 1, 2.0)
    with pytest.raises(rb.StrError, match=r"subsequent factorizations must use the same matrix \(symmetric differs\)"):
        solver.factorize(c2)
    c3 = rb.CooMatrix(1, 1, 1)
    c3.put(0, 0, 1.0)
    with pytest.raises(rb.StrError, match=r"subsequent factorizations must use the same matrix \(ndim differs\)"):
        solver.factorize(c3)
    c4 = rb.CooMatrix(2, 2, 1)
    c4.put(0, 0, 1.0)
    with pytest.raises(rb.StrError, match=r"subsequent factorizations must use the same matrix \(nnz differs\)"):
        solver.factorize(c4)
    with pytest.raises(rb.StrError, match="subsequent factorizations must not change LinSolParams"):
        solver.factorize(coo, rb.LinSolParams())
    solver.factorize(coo)  # calling factorize again with None works


def test_factorize_fails_on_singular_matrix():
    # solver_umfpack.rs:624-630
    solver = rb.SolverB200()
    coo = rb.CooMatrix(2, 2, 2)
    coo.put(0, 0, 1.0), coo.put(1, 1, 0.0)
    with pytest.raises(rb.StrError) as e:
        solver.factorize(coo)
    assert e.value.msg == "Error(1): Matrix is singular"


def test_solve_handles_errors():
    coo = rb.CooMatrix(2, 2, 2)
    coo.put(0, 0, 123.0), coo.put(1, 1, 456.0)
    solver = rb.SolverB200()
    with pytest.raises(rb.StrError, match="the function factorize must be called before solve"):
        solver.solve(np.zeros(2), np.zeros(2))
    solver.factorize(coo)
    with pytest.raises(rb.StrError, match="the dimension of the vector of unknown values x is incorrect"):
        solver.solve(np.zeros(1), np.zeros(2))
    with pytest.raises(rb.StrError, match="the dimension of the right-hand side vector is incorrect"):
        solver.solve(np.zeros(2), np.zeros(1))


def test_c_abi_state_machine_codes():
    from russell_b200 import _lib
    from russell_b200._lib import p_f64, p_i32, ptr

    lib = _lib.load()
    h = lib.solver_b200_new()
    assert h
    v = np.ones(1)
    assert lib.solver_b200_factorize(h, None, None, 0, ptr(v, p_f64)) == 500000  # interface_cudss.cu:416-418
    assert lib.solver_b200_solve(h, ptr(v, p_f64), ptr(v, p_f64), 0) == 600000   # interface_cudss.cu:518-520
    rp, ci = np.array([0, 1], dtype=np.int32), np.array([0], dtype=np.int32)
    assert lib.solver_b200_initialize(h, 0, 0, 0, -1.0, -1, -1.0, 0, 0, 0, 1, ptr(rp, p_i32), ptr(ci, p_i32), ptr(v, p_f64)) == 0
    assert lib.solver_b200_initialize(h, 0, 0, 0, -1.0, -1, -1.0, 0, 0, 0, 1, ptr(rp, p_i32), ptr(ci, p_i32), ptr(v, p_f64)) == 700000
    assert lib.solver_b200_solve(h, ptr(v, p_f64), ptr(v, p_f64), 0) == 600000
    assert lib.solver_b200_factorize(h, None, None, 0, ptr(v, p_f64)) == 0
    bad = np.array([np.nan])
    assert lib.solver_b200_factorize(h, None, None, 0, ptr(bad, p_f64)) == 802
    lib.solver_b200_drop(h)


# ---- kernels vs the scalar walk of the same plan: factor panels array by array -------------------------------------
def _factor_compare(coo, opts):
    from russell_b200 import _lib
    from russell_b200._lib import p_f64, p_i32, ptr

    sol = rb.SolverB200()
    for k, v in opts.items():
        sol.set_option(k, v)
    sol.factorize(coo)
    csr = rb.CsrMatrix.from_coo(coo)
    h = oracle.MfHandle(csr.nrow, csr.pointers, csr.indices, csr.values[: csr.nnz], sym_lower=(coo.symmetric == rb.Sym.YesLower),
                        matching=2, panel_width=int(opts.get("panel_width", 0)), nd_leaf=int(opts.get("nd_leaf", 0)))
    hfac, hdinv, hperm = h.factors()
    fac, dinv, lperm = np.zeros(h.fac_size), np.zeros(h.dinv_size), np.zeros(h.n, dtype=np.int32)
    rc = _lib.load().solver_b200_debug_copy_factors(sol.solver, ptr(fac, p_f64), len(fac), ptr(dinv, p_f64), len(dinv), ptr(lperm, p_i32), len(lperm))
    assert rc == 0
    assert int(sol.device_stats()["fac_bytes"] / 8) == h.fac_size
    assert np.array_equal(lperm, hperm)  # pivot choices: bit-exact index work
    scale = np.max(np.abs(hfac))
    assert np.max(np.abs(fac - hfac)) <= 1e-11 * scale
    assert np.max(np.abs(dinv - hdinv)) <= 1e-10 * max(1.0, np.max(np.abs(hdinv)))
    return sol


@pytest.mark.parametrize("schur_variant", [0, 1])
@pytest.mark.parametrize("k", [1, 2, 5, 17, 64, 130])
def test_factor_panels_match_host_walk_laplacian(k, schur_variant):
    _factor_compare(helpers.laplacian_2d_coo(k), {"schur_variant": schur_variant})


@pytest.mark.parametrize("panel_width,nd_leaf", [(4, 4), (8, 16), (16, 8), (32, 50), (64, 300)])
def test_factor_panels_match_host_walk_panel_sizes(panel_width, nd_leaf):
    n, ai, aj, ax = helpers.convection_diffusion_triplets(48)
    _factor_compare(rb.CooMatrix.from_triplets(n, n, ai, aj, ax), {"panel_width": panel_width, "nd_leaf": nd_leaf})


def test_factor_panels_match_host_walk_lower_and_saddle():
    _factor_compare(helpers.laplacian_2d_coo(70, lower=True), {})
    n, ai, aj, ax = helpers.saddle_point_triplets(30)
    _factor_compare(rb.CooMatrix.from_triplets(n, n, ai, aj, ax), {})


@pytest.mark.parametrize("opts", [{"diag_variant": 0}, {"use_fused": 0}, {"use_top": 0}, {"fuse_chain": 0}, {"schur_variant": 0, "use_fused": 0},
                                  {"panel_width": 16, "use_fused": 0}, {"diag_variant": 0, "use_fused": 0}, {"diag_variant": 0, "panel_width": 20, "use_fused": 0},
                                  {"diag_variant": 0, "panel_width": 37, "use_fused": 0}, {"panel_variant": 1},
                                  {"panel_variant": 1, "use_fused": 0, "panel_width": 37}, {"panel_variant": 1, "use_fused": 0, "panel_width": 8},
                                  {"panel_variant": 0}, {"panel_variant": 0, "use_fused": 0, "panel_width": 37}, {"use_graph": 0}, {"diag_variant": 4},
                                  {"fused_variant": 0}, {"fused_variant": 1, "fused_maxf": 64}, {"fused_variant": 1, "fused_maxf": 20, "panel_width": 13},
                                  {"diag_variant": 4, "use_fused": 0, "panel_width": 37}, {"diag_variant": 4, "use_fused": 0, "panel_width": 5},
                                  {"asm_variant": 0}, {"asm_variant": 0, "use_fused": 0, "panel_width": 13}, {"asm_variant": 1, "use_fused": 0, "panel_width": 13},
                                  {"nd_leaf": 96}, {"schur_variant": 2, "ozaki_min_u": 64}, {"schur_variant": 2, "ozaki_min_u": 100, "panel_width": 37},
                                  {"schur_variant": 2, "ozaki_min_u": 1, "use_fused": 0, "panel_width": 20},
                                  {"schur_front_nt": 2}, {"schur_front_nt": 1, "use_fused": 0, "panel_width": 37}, {"schur_front_nt": 1000}])
def test_kernel_variants_match_host_walk(opts):
    # every alternative code path (shared-memory vs register-resident pivot-block LU, fused vs multi-kernel fronts,
    # persistent vs per-level sweeps) against the scalar walk, on a grid with fronts above the fused limit
    n, ai, aj, ax = helpers.convection_diffusion_triplets(140)
    coo = rb.CooMatrix.from_triplets(n, n, ai, aj, ax)
    sol = _factor_compare(coo, opts)
    b = np.sin(np.arange(n) + 1.0)
    x = np.zeros(n)
    sol.solve(x, b)
    a = oracle.full_scipy_matrix(n, n, ai, aj, ax)
    assert np.linalg.norm(b - a @ x) / np.linalg.norm(b) <= TOL_RESIDUAL


def _raw_factors(coo, opts):
    from russell_b200 import _lib
    from russell_b200._lib import p_f64, p_i32, ptr

    sol = rb.SolverB200()
    for k, v in opts.items():
        sol.set_option(k, v)
    sol.factorize(coo)
    st = sol.device_stats()
    fac = np.zeros(int(st["fac_bytes"] / 8))
    lperm = np.zeros(coo.nrow, dtype=np.int32)
    dinv = np.zeros(1)
    rc = _lib.load().solver_b200_debug_copy_factors(sol.solver, ptr(fac, p_f64), len(fac), ptr(dinv, p_f64), 0, ptr(lperm, p_i32), len(lperm))
    assert rc == 0
    return fac, lperm


def test_register_kernels_are_bit_identical_to_the_shared_memory_fallbacks():
    # k_diag_w8 (register-resident, one warp per column group) performs the same operations per entry in the same order as
    # the shared-memory LU k_diag; k_front_fused_w8 vs k_front_fused and k_panel_warp vs k_panel likewise: same bits
    rng = np.random.default_rng(3)
    n, ai, aj, ax = helpers.convection_diffusion_triplets(120)
    ax = ax * (1.0 + 0.3 * rng.standard_normal(len(ax)))
    coo = rb.CooMatrix.from_triplets(n, n, ai, aj, ax)
    for extra in ({}, {"use_fused": 0}, {"use_fused": 0, "panel_width": 29}):
        f1, p1 = _raw_factors(coo, dict(extra, diag_variant=0))
        f6, p6 = _raw_factors(coo, dict(extra, diag_variant=4))
        assert np.array_equal(p1, p6)
        assert np.array_equal(f1, f6)
    for extra in ({}, {"fused_maxf": 64}, {"fused_maxf": 30, "panel_width": 11}):
        f7, p7 = _raw_factors(coo, dict(extra, fused_variant=0))
        f8, p8 = _raw_factors(coo, dict(extra, fused_variant=1))
        assert np.array_equal(p7, p8)
        assert np.array_equal(f7, f8)
        # thread-per-row triangular panel solves: same operation order as the tile kernel
        f3, p3 = _raw_factors(coo, dict(extra, panel_variant=0))
        f4, p4 = _raw_factors(coo, dict(extra, panel_variant=1))
        assert np.array_equal(p3, p4)
        assert np.array_equal(f3, f4)


def test_round1_kernels_are_bit_identical_to_the_ones_they_replace():
    # k_panel_row (warp per four rows) vs k_panel_warp (thread per row), k_leaf_reg (leaf fronts in registers) vs
    # k_front_fused: same operations per entry in the same order => same bits, same pivots
    rng = np.random.default_rng(5)
    n, ai, aj, ax = helpers.convection_diffusion_triplets(150)
    ax = ax * (1.0 + 0.3 * rng.standard_normal(len(ax)))
    coo = rb.CooMatrix.from_triplets(n, n, ai, aj, ax)
    for extra in ({}, {"use_fused": 0, "panel_width": 29}, {"panel_width": 64, "nd_leaf": 300}):
        f0, p0 = _raw_factors(coo, dict(extra, panel_row_max=0, use_leaf_reg=0, asm_variant=0))
        f1, p1 = _raw_factors(coo, dict(extra, panel_row_max=100000, use_leaf_reg=0))
        f2, p2 = _raw_factors(coo, dict(extra, panel_row_max=0, use_leaf_reg=1))
        assert np.array_equal(p0, p1) and np.array_equal(f0, f1)
        assert np.array_equal(p0, p2) and np.array_equal(f0, f2)
        f3, p3 = _raw_factors(coo, dict(extra, panel_row_max=0, use_leaf_reg=0, asm_variant=1))  # shared-memory extend-add tile
        assert np.array_equal(p0, p3) and np.array_equal(f0, f3)


@pytest.mark.parametrize("k,lower", [(150, False), (400, True)])
def test_sweep_variants_agree(k, lower):
    # per-level launches, persistent top-of-tree kernels, with and without the one-CTA-per-subtree kernels.
    # The persistent kernels give the same bits with and without graph replay; the per-level and
    # subtree kernels sum the pivot-block GEMV in a different (also fixed) order, so they agree to rounding only.
    coo = helpers.laplacian_2d_coo(k, lower=lower)
    b = np.sin(0.1 * np.arange(coo.nrow)) + 1.0
    ref, ref_top = None, None
    for opts in ({"use_top": 0, "use_subtree": 0}, {"use_subtree": 0}, {"use_subtree": 1}, {"use_top": 0, "use_subtree": 1},
                 {"use_subtree": 1, "subtree_budget": 2000, "subtree_maxf": 40}, {"top_max_nodes": 8},
                 {"use_subtree": 0, "use_graph": 0}):
        sol, x = solve_through_abi(coo, b, opts=opts)
        assert sol.residual(x, b) <= TOL_RESIDUAL, opts
        x2 = np.zeros_like(x)
        sol.solve(x2, b)  # epoch counters of the persistent kernels: a second sweep must behave like the first
        assert np.array_equal(x, x2), opts
        if ref is None:
            ref = x
        assert np.max(np.abs(x - ref)) <= 1e-12 * np.max(np.abs(ref)), opts
        if opts.get("use_subtree") == 0 and "use_top" not in opts:
            if ref_top is None:
                ref_top = x
            assert np.array_equal(x, ref_top), opts


def test_graph_replay_equals_direct_launches():
    coo = helpers.laplacian_2d_coo(90)
    b = np.cos(np.arange(coo.nrow))
    _, xg = solve_through_abi(coo, b, opts={"use_graph": 1})
    _, xd = solve_through_abi(coo, b, opts={"use_graph": 0})
    assert np.array_equal(xg, xd)


# ---- SpMV / residual kernel vs csr_matrix.rs:709-729 restatement -------------------------------------------------
@pytest.mark.parametrize("lower", [False, True])
def test_spmv_kernel_matches_oracle(lower):
    rng = np.random.default_rng(2)
    for k in (3, 40, 257):
        coo = helpers.laplacian_2d_coo(k, lower=lower)
        coo.values[: coo.nnz] *= 1.0 + 0.1 * rng.standard_normal(coo.nnz) if not lower else 1.0
        sol = rb.SolverB200()
        sol.factorize(coo)
        u = rng.standard_normal(coo.nrow)
        y = sol.mat_vec_mul(u)
        csr = rb.CsrMatrix.from_coo(coo)
        yo = oracle.csr_matvec(csr.pointers, csr.indices[: csr.nnz], csr.values[: csr.nnz], u, mirror=lower)
        assert np.max(np.abs(y - yo)) <= 1e-14 * np.max(np.abs(yo))
        x = rng.standard_normal(coo.nrow)
        res = sol.residual(x, u)
        ro = u - oracle.csr_matvec(csr.pointers, csr.indices[: csr.nnz], csr.values[: csr.nnz], x, mirror=lower)
        assert abs(res - np.linalg.norm(ro) / np.linalg.norm(u)) <= 1e-13 * res


def test_spmv_long_rows_and_empty_rows():
    # a dense row longer than one SpMV block, an arrow matrix, and rows with a single entry
    n = 5000
    ai = np.concatenate([np.zeros(n, np.int32), np.arange(1, n, dtype=np.int32), np.arange(n, dtype=np.int32)])
    aj = np.concatenate([np.arange(n, dtype=np.int32), np.zeros(n - 1, np.int32), np.arange(n, dtype=np.int32)])
    ax = np.concatenate([np.full(n, 0.01), np.full(n - 1, 0.02), np.full(n, 10.0)])
    coo = rb.CooMatrix.from_triplets(n, n, ai, aj, ax)
    sol, x = solve_through_abi(coo, np.ones(n))
    assert sol.residual(x, np.ones(n)) <= TOL_RESIDUAL
    u = np.sin(np.arange(n))
    bp, bj, bx = oracle.coo_to_csr(n, n, ai, aj, ax)
    assert np.max(np.abs(sol.mat_vec_mul(u) - oracle.csr_matvec(bp, bj, bx, u))) <= 1e-12


# ---- solves against the independent CPU LU + size-independent properties ------------------------------------------
@pytest.mark.parametrize("k,lower", [(10, False), (64, True), (200, False), (300, True)])
def test_laplacian_vs_cpu_lu(k, lower):
    coo = helpers.laplacian_2d_coo(k, lower=lower)
    n = coo.nrow
    b = np.ones(n)  # solve_matrix_market.rs:179
    sol, x = solve_through_abi(coo, b)
    a = oracle.full_scipy_matrix(n, n, coo.indices_i[: coo.nnz], coo.indices_j[: coo.nnz], coo.values[: coo.nnz], coo.symmetric.name)
    xs = oracle.lu_solve(a, b)
    assert np.linalg.norm(b - a @ x) / np.linalg.norm(b) <= TOL_RESIDUAL
    assert np.max(np.abs(x - xs)) <= 1e-8 * np.max(np.abs(xs))  # agreement with the CPU LU
    assert sol.device_stats()["last_rel_residual"] <= TOL_RESIDUAL


def test_config2_full_size_properties():
    # BASELINE config 2: 1000 x 1000 grid, 1M dof.  No CPU factorization at this size in the test (13.8 s);
    # parity is held through properties: residual bound, forward error on a manufactured solution, linearity,
    # refactorization idempotence.
    k = 1000
    coo = helpers.laplacian_2d_coo(k)
    n = coo.nrow
    sol = rb.SolverB200()
    sol.factorize(coo)
    st = sol.device_stats()
    assert st["n_perturbed"] == 0
    b1 = np.ones(n)
    x1 = np.zeros(n)
    sol.solve(x1, b1)
    assert sol.residual(x1, b1) <= TOL_RESIDUAL
    xstar = np.sin(np.arange(n, dtype=float))  # SURVEY 8d forward-error check
    b2 = sol.mat_vec_mul(xstar)
    x2 = np.zeros(n)
    sol.solve(x2, b2)
    assert np.max(np.abs(x2 - xstar)) <= 1e-8
    x3 = np.zeros(n)
    sol.solve(x3, 2.0 * b1 - 3.0 * b2)
    assert np.max(np.abs(x3 - (2.0 * x1 - 3.0 * x2))) <= 1e-7 * np.max(np.abs(x1))
    sol.factorize(coo)  # same values again: identical factors, identical solution
    x4 = np.zeros(n)
    sol.solve(x4, b1)
    assert np.array_equal(x4, x1)


def test_unsymmetric_values_and_refactorization_with_new_values():
    n, ai, aj, ax = helpers.convection_diffusion_triplets(150)
    coo = rb.CooMatrix.from_triplets(n, n, ai, aj, ax)
    sol = rb.SolverB200()
    b = np.ones(n)
    x = np.zeros(n)
    for scale in (1.0, 1.5, 0.3):  # the Radau5 pattern: same structure, new values (radau5.rs:260-296)
        coo.values[: coo.nnz] = ax * scale + (ai == aj) * (1.0 / scale)
        sol.factorize(coo)
        sol.solve(x, b)
        a = oracle.full_scipy_matrix(n, n, ai, aj, coo.values[: coo.nnz])
        assert np.linalg.norm(b - a @ x) / np.linalg.norm(b) <= TOL_RESIDUAL
        xs = oracle.lu_solve(a, b)
        assert np.max(np.abs(x - xs)) <= 1e-9 * np.max(np.abs(xs))


@pytest.mark.parametrize("lower", [False, True])
def test_coo_boundary_device_conversion_is_bit_identical_to_host_conversion(lower):
    # solver_b200_factorize_coo (device duplicate summation) against solver_b200_factorize fed by the host
    # CsrMatrix::update_from_coo clone: identical CSR values => identical factors => identical solutions, bit for bit
    rng = np.random.default_rng(21)
    n, ai, aj, ax = helpers.convection_diffusion_triplets(60)
    if lower:
        n, ai, aj, ax = helpers.laplacian_2d_triplets(60, lower=True)
    # finite-element style duplicates: split every entry into 1..4 pieces at random positions of the triplet list
    reps = rng.integers(1, 5, len(ax))
    di, dj = np.repeat(ai, reps), np.repeat(aj, reps)
    perm = rng.permutation(len(di))
    di, dj = di[perm].astype(np.int32), dj[perm].astype(np.int32)
    sym = rb.Sym.YesLower if lower else rb.Sym.No
    b = rng.standard_normal(n)
    sols = {}
    for coo_boundary in (True, False):
        sol = rb.SolverB200(coo_boundary=coo_boundary)
        coo = rb.CooMatrix.from_triplets(n, n, di, dj, np.zeros(len(di)), sym)
        xs = []
        for trial in range(3):  # refactorizations with new values reuse the map
            r2 = np.random.default_rng(100 + trial)
            w = 1.0 + 0.2 * (r2.random(len(di)) - 0.5)  # stays diagonally dominant
            coo.values[: coo.nnz] = (np.repeat(ax / reps, reps)[perm]) * w * (1.0 + trial)
            sol.factorize(coo)
            x = np.zeros(n)
            sol.solve(x, b)
            xs.append(x.copy())
            a = oracle.full_scipy_matrix(n, n, di, dj, coo.values[: coo.nnz], "YesLower" if lower else "No")
            assert np.linalg.norm(b - a @ x) / np.linalg.norm(b) <= TOL_RESIDUAL
        sols[coo_boundary] = xs
    for xa, xb in zip(sols[True], sols[False]):
        assert np.array_equal(xa, xb)


def test_coo_boundary_errors():
    import ctypes
    from russell_b200 import _lib

    lib = _lib.load()
    P = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
    i = np.array([0, 1, 0], np.int32)
    j = np.array([0, 1, 1], np.int32)
    v = np.array([1.0, 2.0, 3.0])
    s = lib.solver_b200_new()
    args = lambda ii, jj, sym: (s, 0, 0, 0, -1.0, -1, -1.0, 0, sym, 0, 2, 3, P(ii, ctypes.c_int32), P(jj, ctypes.c_int32), P(v, ctypes.c_double))
    assert lib.solver_b200_factorize_coo(s, None, None, 0, P(v, ctypes.c_double)) == 500000  # need initialization
    assert lib.solver_b200_initialize_coo(*args(i, j, 1)) == 704  # upper entry under Sym::YesLower
    bad = np.array([0, 5, 0], np.int32)
    assert lib.solver_b200_initialize_coo(*args(bad, j, 0)) == 703
    assert lib.solver_b200_initialize_coo(*args(i, j, 0)) == 0
    assert lib.solver_b200_initialize_coo(*args(i, j, 0)) == 700000  # already initialized
    assert lib.solver_b200_factorize_coo(s, None, None, 0, None) == 100000
    assert lib.solver_b200_factorize_coo(s, None, None, 0, P(v, ctypes.c_double)) == 0
    x, b = np.zeros(2), np.array([4.0, 2.0])
    assert lib.solver_b200_solve(s, P(x, ctypes.c_double), P(b, ctypes.c_double), 0) == 0
    assert np.allclose(x, [1.0, 1.0], atol=1e-15)
    lib.solver_b200_drop(s)
    # a handle initialised through the CSR entry point has no triplet map
    s2 = lib.solver_b200_new()
    rp, ci = np.array([0, 2, 3], np.int32), np.array([0, 1, 1], np.int32)
    assert lib.solver_b200_initialize(s2, 0, 0, 0, -1.0, -1, -1.0, 0, 0, 0, 2, P(rp, ctypes.c_int32), P(ci, ctypes.c_int32), P(v[[0, 2, 1]].copy(), ctypes.c_double)) == 0
    assert lib.solver_b200_factorize_coo(s2, None, None, 0, P(v, ctypes.c_double)) == 500000
    lib.solver_b200_drop(s2)


def test_saddle_point_and_random_zero_diagonal():
    n, ai, aj, ax = helpers.saddle_point_triplets(40)
    coo = rb.CooMatrix.from_triplets(n, n, ai, aj, ax)
    b = np.sin(np.arange(n) + 1.0)
    sol, x = solve_through_abi(coo, b)
    a = oracle.full_scipy_matrix(n, n, ai, aj, ax)
    assert np.linalg.norm(b - a @ x) / np.linalg.norm(b) <= TOL_RESIDUAL
    assert sol.effective_matching == 5
    import scipy.sparse as sp

    rng = np.random.default_rng(5)
    m = 400
    r = sp.random(m, m, density=0.02, random_state=5, format="coo")
    perm = rng.permutation(m)
    ri = np.concatenate([r.row, np.arange(m)]).astype(np.int32)
    rj = np.concatenate([r.col, perm]).astype(np.int32)
    rx = np.concatenate([r.data, 10.0 + rng.random(m)])
    keep = ri != rj
    coo = rb.CooMatrix.from_triplets(m, m, ri[keep], rj[keep], rx[keep])
    b = np.ones(m)
    sol, x = solve_through_abi(coo, b)
    a = oracle.full_scipy_matrix(m, m, ri[keep], rj[keep], rx[keep])
    assert np.linalg.norm(b - a @ x) / np.linalg.norm(b) <= TOL_RESIDUAL


def test_determinant():
    # solver_umfpack.rs:585-606: det(umfpack_unsymmetric_5x5) = 114 @1e-13; samples carry their determinants
    for name in ("umfpack_unsymmetric_5x5", "mkl_unsymmetric_5x5", "unsymmetric_3x3(false,false)", "block_unsymmetric_5x5(true,true)",
                 "mkl_symmetric_5x5_full", "positive_definite_3x3_lower", "tiny_1x1"):
        coo, s = helpers.sample_coo(name)
        sol = rb.SolverB200()
        sol.factorize(coo)
        c, e = sol.determinant()
        assert abs(c * 10.0**e - s["det"]) <= 1e-12 * abs(s["det"]), name


def test_two_handles_on_two_threads():
    # radau5.rs:270-296 drives a real and a complex solver from two scoped threads; handles must be independent
    results = {}

    def work(tag, k):
        coo = helpers.laplacian_2d_coo(k)
        b = np.ones(coo.nrow)
        sol, x = solve_through_abi(coo, b)
        results[tag] = sol.residual(x, b)

    ts = [threading.Thread(target=work, args=(i, 60 + 10 * i)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(results) == 4 and max(results.values()) <= TOL_RESIDUAL


# ---- 3D problems: large separator fronts (the shape of BASELINE.json configs[2], whose file is not in the tree) -----------
def test_laplacian_3d_vs_cpu_lu():
    n, ai, aj, ax = helpers.laplacian_3d_triplets(22, skew=1e-3)  # 10,648 dof, unsymmetric values, top separator 22^2
    coo = rb.CooMatrix.from_triplets(n, n, ai, aj, ax)
    b = np.ones(n)
    sol, x = solve_through_abi(coo, b)
    a = oracle.full_scipy_matrix(n, n, ai, aj, ax)
    xs = oracle.lu_solve(a, b)
    assert np.linalg.norm(b - a @ x) / np.linalg.norm(b) <= TOL_RESIDUAL
    assert np.max(np.abs(x - xs)) <= 1e-9 * np.max(np.abs(xs))
    assert sol.device_stats()["max_front"] >= 22 * 22


def test_laplacian_3d_60_properties():
    # 216,000 dof, 1.5 M nonzeros, separator fronts of ~3600 columns (56 chain panels): held by properties
    n, ai, aj, ax = helpers.laplacian_3d_triplets(60, skew=1e-3)
    coo = rb.CooMatrix.from_triplets(n, n, ai, aj, ax)
    sol = rb.SolverB200()
    sol.factorize(coo)
    st = sol.device_stats()
    assert st["n_perturbed"] == 0 and st["max_front"] >= 3000
    xstar = np.cos(0.01 * np.arange(n))
    b = sol.mat_vec_mul(xstar)
    x = np.zeros(n)
    sol.solve(x, b)
    assert sol.residual(x, b) <= TOL_RESIDUAL
    assert np.max(np.abs(x - xstar)) <= 1e-8


def test_radau5_real_and_complex_handles_on_two_threads():
    # russell_ode/src/radau5.rs:270-296: with concurrency on, the real and the complex system are factorized (and solved)
    # on two threads at the same time, each through its own handle (own stream, own buffers)
    ndim, ai, aj, kr, kc = helpers.brusselator_radau5_triplets(40, h=1e-3)
    rcoo = rb.CooMatrix.from_triplets(ndim, ndim, ai, aj, kr)
    ccoo = rb.ComplexCooMatrix.from_triplets(ndim, ndim, ai, aj, kc)
    rsol, csol = rb.SolverB200(), rb.ComplexSolverB200()
    b = np.ones(ndim)
    bz = np.ones(ndim, dtype=np.complex128) * (1.0 - 2.0j)
    x, z = np.zeros(ndim), np.zeros(ndim, dtype=np.complex128)
    errs = []

    def run_real():
        try:
            for _ in range(3):
                rsol.factorize(rcoo)
                rsol.solve(x, b)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    def run_complex():
        try:
            for _ in range(3):
                csol.factorize(ccoo)
                csol.solve(z, bz)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    t1, t2 = threading.Thread(target=run_real), threading.Thread(target=run_complex)
    t1.start(), t2.start()
    t1.join(), t2.join()
    assert not errs, errs
    assert rsol.residual(x, b) <= TOL_RESIDUAL and csol.residual(z, bz) <= TOL_RESIDUAL
    # same answers as a serial run
    x2, z2 = np.zeros(ndim), np.zeros(ndim, dtype=np.complex128)
    rsol.solve(x2, b)
    csol.solve(z2, bz)
    assert np.array_equal(x, x2) and np.array_equal(z, z2)
