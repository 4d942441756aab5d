"""GPU parity tests of the Complex64 twin (run with `-m gpu` on the B200 box): the reference's complex solver tests
(russell_sparse/src/complex_solver_cudss.rs:438-806, complex_solver_umfpack.rs:463-612, complex_lin_solver.rs:191-228)
re-stated against `ComplexSolverB200`, every call through the C ABI `complex_solver_b200_*`; plus the two Radau5
Newton matrices of the Brusselator PDE (BASELINE.json configs[3]) against the independent CPU LU."""
import numpy as np
import pytest

import helpers
import russell_b200 as rb
from oracle import oracle
from test_complex_cpu import COMPLEX_KATS, dense_of

pytestmark = pytest.mark.gpu
TOL_RESIDUAL = 1e-10  # BASELINE.json north_star


def csolve(coo, b, params=None):
    sol = rb.ComplexSolverB200()
    sol.factorize(coo, params)
    x = np.zeros(coo.nrow, dtype=np.complex128)
    sol.solve(x, np.asarray(b, dtype=np.complex128))
    return sol, x


@pytest.mark.parametrize("name,rhs,xc,tol,src", COMPLEX_KATS)
def test_complex_known_answers(name, rhs, xc, tol, src):
    coo, _ = helpers.complex_sample_coo(name)
    sol, x = csolve(coo, rhs)
    assert np.max(np.abs(x - np.array(xc))) <= tol * max(1.0, np.max(np.abs(xc))), src
    x2 = np.zeros(coo.nrow, dtype=np.complex128)  # calling solve again (complex_solver_cudss.rs:531-538)
    sol.solve(x2, np.asarray(rhs, dtype=np.complex128))
    assert np.array_equal(x, x2)
    stats = rb.StatsLinSol()
    sol.update_stats(stats)
    assert stats.solver == "B200" and len(stats.initialize_array) == 1 and len(stats.solve_array) == 1
    assert sol.get_ns_init() > 0 and sol.get_ns_fact() > 0 and sol.get_ns_solve() > 0


def test_complex_unsymmetric_5x5_parameter_variants():
    # complex_solver_cudss.rs:488-650: default (cuDSS loses x[3] to 1.2e-3 there; we must not), colamd, matching, pivot
    coo, _ = helpers.complex_sample_coo("umfpack_complex_unsymmetric_5x5")
    rhs = [8.0, 45.0, -3.0, 3.0, 19.0]
    xc = np.arange(1.0, 6.0)
    for setup in ("default", "colamd", "matching", "pivot"):
        par = rb.LinSolParams()
        if setup == "colamd":
            par.ordering = rb.Ordering.Colamd
        elif setup == "matching":
            par.matching = rb.Matching.Auto
        elif setup == "pivot":
            par.pivot_epsilon, par.refinement_nstep = 1e-12, 1
        sol, x = csolve(coo, rhs, par)
        assert np.max(np.abs(x - xc)) <= 1e-12, setup
        with pytest.raises(rb.StrError, match="subsequent factorizations must not change LinSolParams"):
            sol.factorize(coo, par)


def test_complex_cudss_example_systems():
    # complex_solver_cudss.rs:680-745 (and hybrid_memory_works :748-778): SPD lower and unsymmetric, tol 1e-10
    spd = [(0, 0, 4.0), (1, 1, 3.0), (2, 0, 1.0), (2, 1, 2.0), (2, 2, 5.0), (3, 3, 1.0), (4, 2, 1.0), (4, 4, 2.0)]
    for hybrid in (None, 0.5):
        coo = rb.ComplexCooMatrix(5, 5, 8, rb.Sym.YesLower)
        for i, j, v in spd:
            coo.put(i, j, complex(v, 0.0))
        par = rb.LinSolParams()
        par.positive_definite = True
        par.hybrid_memory_factor = hybrid
        _, x = csolve(coo, [7.0, 12.0, 25.0, 4.0, 13.0], par)
        assert np.max(np.abs(x - np.arange(1.0, 6.0))) <= 1e-10
    uns = [(0, 0, 5.0), (0, 1, 1.0), (0, 4, 3.0), (1, 0, 2.0), (1, 1, 6.0), (1, 3, 4.0), (2, 2, 7.0), (2, 3, 2.0),
           (3, 1, 1.0), (3, 2, 3.0), (3, 3, 8.0), (4, 0, 4.0), (4, 4, 9.0)]
    coo = rb.ComplexCooMatrix(5, 5, 13, rb.Sym.No)
    for i, j, v in uns:
        coo.put(i, j, complex(v, 0.0))
    _, x = csolve(coo, [22.0, 30.0, 29.0, 43.0, 49.0])
    assert np.max(np.abs(x - np.arange(1.0, 6.0))) <= 1e-10


def test_complex_lin_solver_compute():
    # complex_lin_solver.rs:195-227
    for name in ("complex_symmetric_3x3_lower", "complex_symmetric_3x3_full"):
        coo, _ = helpers.complex_sample_coo(name)
        x = np.zeros(3, dtype=np.complex128)
        rb.ComplexLinSolver.compute(rb.Genie.B200, x, coo, np.array([-3 + 3j, 2 - 2j, 9 + 7j]))
        assert np.max(np.abs(x - np.array([1 + 1j, 2 - 2j, 3 + 3j]))) <= 1e-14


def test_complex_factorize_handles_errors():
    # complex_solver_cudss.rs:438-486
    sol = rb.ComplexSolverB200()
    assert not sol.factorized
    coo, _ = helpers.complex_sample_coo("complex_rectangular_4x3")
    with pytest.raises(rb.StrError, match="the matrix must be square"):
        sol.factorize(coo)
    with pytest.raises(rb.StrError, match="the COO matrix must have at least one non-zero value"):
        sol.factorize(rb.ComplexCooMatrix(1, 1, 1, rb.Sym.No))
    coo, _ = helpers.complex_sample_coo("complex_symmetric_3x3_upper")
    with pytest.raises(rb.StrError, match="B200 requires Sym::YesLower or Sym::YesFull for symmetric matrices"):
        sol.factorize(coo)
    coo = rb.ComplexCooMatrix(2, 2, 2, rb.Sym.No)
    coo.put(0, 0, 1.0)
    coo.put(1, 1, 2.0)
    sol.factorize(coo)
    bad = rb.ComplexCooMatrix(2, 2, 2, rb.Sym.YesFull)
    bad.put(0, 0, 1.0)
    bad.put(1, 1, 2.0)
    with pytest.raises(rb.StrError, match=r"subsequent factorizations must use the same matrix \(symmetric differs\)"):
        sol.factorize(bad)
    bad = rb.ComplexCooMatrix(1, 1, 1, rb.Sym.No)
    bad.put(0, 0, 1.0)
    with pytest.raises(rb.StrError, match=r"subsequent factorizations must use the same matrix \(ndim differs\)"):
        sol.factorize(bad)
    bad = rb.ComplexCooMatrix(2, 2, 1, rb.Sym.No)
    bad.put(0, 0, 1.0)
    with pytest.raises(rb.StrError, match=r"subsequent factorizations must use the same matrix \(nnz differs\)"):
        sol.factorize(bad)


def test_complex_solve_handles_errors():
    # complex_solver_cudss.rs:781-805
    coo = rb.ComplexCooMatrix(2, 2, 2, rb.Sym.No)
    coo.put(0, 0, 123.0)
    coo.put(1, 1, 456.0)
    sol = rb.ComplexSolverB200()
    x, rhs = np.zeros(2, dtype=np.complex128), np.zeros(2, dtype=np.complex128)
    with pytest.raises(rb.StrError, match="the function factorize must be called before solve"):
        sol.solve(x, rhs)
    sol.factorize(coo)
    with pytest.raises(rb.StrError, match="the dimension of the vector of unknown values x is incorrect"):
        sol.solve(np.zeros(1, dtype=np.complex128), rhs)
    with pytest.raises(rb.StrError, match="the dimension of the right-hand side vector is incorrect"):
        sol.solve(x, np.zeros(1, dtype=np.complex128))


def test_complex_c_abi_state_machine_codes():
    from russell_b200 import _lib
    from russell_b200._lib import p_f64, p_i32, ptr

    lib = _lib.load()
    h = lib.complex_solver_b200_new()
    assert h
    v = np.array([1.0, 0.5, 2.0, -0.5])  # diag(1+0.5i, 2-0.5i)
    rp, ci = np.array([0, 1, 2], dtype=np.int32), np.array([0, 1], dtype=np.int32)
    x, b = np.zeros(4), np.array([1.0, 0.5, 2.0, -0.5])
    assert lib.complex_solver_b200_factorize(h, None, None, 0, ptr(v, p_f64)) == 500000  # interface_complex_cudss.cu NEED_INITIALIZATION
    assert lib.complex_solver_b200_solve(h, ptr(x, p_f64), ptr(b, p_f64), 0) == 600000
    args = (0, 0, 0, -1.0, -1, -1.0, 0, 0, 0, 2, ptr(rp, p_i32), ptr(ci, p_i32), ptr(v, p_f64))
    assert lib.complex_solver_b200_initialize(h, *args) == 0
    assert lib.complex_solver_b200_initialize(h, *args) == 700000
    assert lib.complex_solver_b200_solve(h, ptr(x, p_f64), ptr(b, p_f64), 0) == 600000
    assert lib.complex_solver_b200_factorize(h, None, None, 0, ptr(v, p_f64)) == 0
    assert lib.complex_solver_b200_solve(h, ptr(x, p_f64), ptr(b, p_f64), 0) == 0
    assert np.allclose(x, [1.0, 0.0, 1.0, 0.0], atol=1e-15)
    lib.complex_solver_b200_drop(h)
    lib.complex_solver_b200_drop(None)


def test_complex_verify_and_spmv_kernel_match_oracle():
    # VerifyLinSys::from_complex (verify_lin_sys.rs:104-146) with A·x through the CUDA SpMV of the embedded matrix
    rng = np.random.default_rng(3)
    n, ai, aj, ax = helpers.convection_diffusion_triplets(40)
    az = ax * (1.0 + 0.3j) + (ai == aj) * 2.0j + 0.1j * np.sin(ai + 2.0 * aj)
    coo = rb.ComplexCooMatrix.from_triplets(n, n, ai, aj, az)
    sol = rb.ComplexSolverB200()
    sol.factorize(coo)
    u = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    want = oracle.complex_coo_matvec(n, ai, aj, az, u)
    assert np.max(np.abs(sol.mat_vec_mul(u) - want)) <= 1e-12 * np.max(np.abs(want))
    x = np.zeros(n, dtype=np.complex128)
    sol.solve(x, want)
    assert np.max(np.abs(x - u)) <= 1e-10
    got = rb.verify_from_complex(coo, x, want, sol)
    ref = oracle.complex_verify(n, ai, aj, az, x, want)
    assert abs(got.max_abs_a - ref["max_abs_a"]) <= 1e-15 * ref["max_abs_a"]
    assert abs(got.max_abs_ax - ref["max_abs_ax"]) <= 1e-12 * ref["max_abs_ax"]
    assert got.relative_error <= 1e-12 and ref["relative_error"] <= 1e-12
    assert sol.residual(x, want) <= TOL_RESIDUAL


def test_complex_symmetric_lower_equals_full():
    # Sym::YesLower input (mirrored on the host of the library) gives the same solution as the full matrix
    rng = np.random.default_rng(5)
    n, ai, aj, ax = helpers.laplacian_2d_triplets(30, lower=True)
    az = ax * (1.0 - 0.2j) + (ai == aj) * (0.5 + 1.0j)
    lo = rb.ComplexCooMatrix.from_triplets(n, n, ai, aj, az, rb.Sym.YesLower)
    off = ai != aj
    fi, fj, fz = np.concatenate([ai, aj[off]]), np.concatenate([aj, ai[off]]), np.concatenate([az, az[off]])
    fu = rb.ComplexCooMatrix.from_triplets(n, n, fi, fj, fz, rb.Sym.No)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    s1, x1 = csolve(lo, b)
    s2, x2 = csolve(fu, b)
    assert np.max(np.abs(x1 - x2)) <= 1e-12 * np.max(np.abs(x2))
    import scipy.sparse as sp

    a = sp.coo_matrix((fz, (fi, fj)), shape=(n, n)).tocsc()
    xs = oracle.lu_solve(a, b)
    assert np.max(np.abs(x1 - xs)) <= 1e-10 * np.max(np.abs(xs))
    assert s1.residual(x1, b) <= TOL_RESIDUAL and s2.residual(x2, b) <= TOL_RESIDUAL


@pytest.mark.parametrize("npoint", [9, 60])
def test_radau5_brusselator_newton_matrices_vs_cpu_lu(npoint):
    # BASELINE.json configs[3]: the real and the complex system Radau5 factorizes per Jacobian (radau5.rs:226-296),
    # refactorized for a sequence of step sizes h (same structure, new values), against the independent CPU LU
    import scipy.sparse as sp

    ndim, ai, aj, kr, kc = helpers.brusselator_radau5_triplets(npoint, h=1e-4)
    rcoo = rb.CooMatrix.from_triplets(ndim, ndim, ai, aj, kr)
    ccoo = rb.ComplexCooMatrix.from_triplets(ndim, ndim, ai, aj, kc)
    rsol, csol = rb.SolverB200(), rb.ComplexSolverB200()
    rng = np.random.default_rng(11)
    b = rng.standard_normal(ndim)
    bz = b + 1j * rng.standard_normal(ndim)
    x, z = np.zeros(ndim), np.zeros(ndim, dtype=np.complex128)
    for h in (1e-4, 1.5e-4, 1e-2, 0.3):
        _, _, _, kr, kc = helpers.brusselator_radau5_triplets(npoint, h=h)
        rcoo.values[:] = kr
        ccoo.values[:] = kc
        rsol.factorize(rcoo)
        csol.factorize(ccoo)
        rsol.solve(x, b)
        csol.solve(z, bz)
        ar = sp.coo_matrix((kr, (ai, aj)), shape=(ndim, ndim)).tocsc()
        ac = sp.coo_matrix((kc, (ai, aj)), shape=(ndim, ndim)).tocsc()
        assert np.linalg.norm(b - ar @ x) / np.linalg.norm(b) <= TOL_RESIDUAL
        assert np.linalg.norm(bz - ac @ z) / np.linalg.norm(bz) <= TOL_RESIDUAL
        xs, zs = oracle.lu_solve(ar, b), oracle.lu_solve(ac, bz)
        assert np.max(np.abs(x - xs)) <= 1e-9 * np.max(np.abs(xs))
        assert np.max(np.abs(z - zs)) <= 1e-9 * np.max(np.abs(zs))
    assert csol.device_stats()["n_perturbed"] == 0


def test_radau5_brusselator_n200_properties():
    # a mid-size instance (ndim = 80,000; embedded order 160,000) held by size-independent properties
    npoint = 200
    ndim, ai, aj, kr, kc = helpers.brusselator_radau5_triplets(npoint, h=1e-3)
    csol = rb.ComplexSolverB200()
    csol.factorize(rb.ComplexCooMatrix.from_triplets(ndim, ndim, ai, aj, kc))
    zstar = np.sin(np.arange(ndim)) + 1j * np.cos(0.5 * np.arange(ndim))
    bz = csol.mat_vec_mul(zstar)
    z = np.zeros(ndim, dtype=np.complex128)
    csol.solve(z, bz)
    assert np.max(np.abs(z - zstar)) <= 1e-9
    assert csol.residual(z, bz) <= TOL_RESIDUAL
    z2 = np.zeros(ndim, dtype=np.complex128)
    csol.solve(z2, (2.0 - 1.0j) * bz)  # linearity over the complex field
    assert np.max(np.abs(z2 - (2.0 - 1.0j) * z)) <= 1e-9


def test_complex_coo_boundary_is_bit_identical_to_host_conversion():
    # complex_solver_b200_factorize_coo (duplicates summed on the device) against complex_solver_b200_factorize fed by the
    # host ComplexCsrMatrix::update_from_coo clone: identical CSR values => identical factors => identical solutions
    ndim, ai, aj, _, kc = helpers.brusselator_radau5_triplets(30, h=2e-3)  # 16 triplets per grid point, with duplicates
    coo = rb.ComplexCooMatrix.from_triplets(ndim, ndim, ai, aj, kc)
    rng = np.random.default_rng(1)
    b = rng.standard_normal(ndim) + 1j * rng.standard_normal(ndim)
    xs = []
    for flag in (True, False):
        sol = rb.ComplexSolverB200(coo_boundary=flag)
        sol.factorize(coo)
        x = np.zeros(ndim, dtype=np.complex128)
        sol.solve(x, b)
        assert sol.residual(x, b) <= TOL_RESIDUAL
        coo.values[: coo.nnz] = kc * (1.0 + 0.25j)  # refactorize with new values, same structure
        sol.factorize(coo)
        x2 = np.zeros(ndim, dtype=np.complex128)
        sol.solve(x2, b)
        assert sol.residual(x2, b) <= TOL_RESIDUAL
        coo.values[: coo.nnz] = kc
        xs.append((x, x2))
    assert np.array_equal(xs[0][0], xs[1][0]) and np.array_equal(xs[0][1], xs[1][1])
    assert np.max(np.abs(xs[0][0] - xs[0][1] * (1.0 + 0.25j))) <= 1e-12 * np.max(np.abs(xs[0][0]))  # (cA) x = b  =>  x scales by 1/c


def test_complex_coo_boundary_errors():
    from russell_b200 import _lib
    from russell_b200._lib import p_f64, p_i32, ptr

    lib = _lib.load()
    h = lib.complex_solver_b200_new()
    ii, jj = np.array([0, 1, 0], dtype=np.int32), np.array([0, 1, 1], dtype=np.int32)
    vv = np.array([1.0, 0.0, 2.0, 0.5, 0.25, 0.0])
    args = lambda sym, i, j, nnz: (0, 0, 0, -1.0, -1, -1.0, 0, sym, 0, 2, nnz, ptr(i, p_i32), ptr(j, p_i32), ptr(vv, p_f64))
    assert lib.complex_solver_b200_factorize_coo(h, None, None, 0, ptr(vv, p_f64)) == 500000
    assert lib.complex_solver_b200_initialize_coo(h, *args(1, ii, jj, 3)) == 704          # j > i with Sym::YesLower
    bad = np.array([0, 2, 0], dtype=np.int32)
    assert lib.complex_solver_b200_initialize_coo(h, *args(0, bad, jj, 3)) == 703         # index out of range
    assert lib.complex_solver_b200_initialize_coo(h, *args(0, ii, jj, 3)) == 0
    assert lib.complex_solver_b200_initialize_coo(h, *args(0, ii, jj, 3)) == 700000
    assert lib.complex_solver_b200_factorize_coo(h, None, None, 0, ptr(vv, p_f64)) == 0
    x, b = np.zeros(4), np.array([1.25, 0.0, 2.0, 0.5])  # A = [[1, 0.25], [0, 2+0.5i]], x = (1, 1)
    assert lib.complex_solver_b200_solve(h, ptr(x, p_f64), ptr(b, p_f64), 0) == 0
    assert np.allclose(x, [1.0, 0.0, 1.0, 0.0], atol=1e-14)
    lib.complex_solver_b200_drop(h)
