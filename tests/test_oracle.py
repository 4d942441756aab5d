"""CPU tests (no GPU): pin the oracle against every golden vector the reference's own tests hold for this path
(SURVEY.md 8c), and check the product's HOST code (COO->CSR/CSC converter, Matrix Market reader, analysis)
against the oracle.  Fixtures were generated from the reference by tests/golden/make_golden.py."""
import numpy as np
import pytest

import helpers
from oracle import oracle

SAMPLES = helpers.load_samples()
SQUARE = [k for k, v in SAMPLES.items() if v["nrow"] == v["ncol"]]


# ---- host formats: oracle vs the reference's hand-written CSR/CSC arrays (samples.rs) ---------------------------
@pytest.mark.parametrize("name", sorted(SAMPLES))
def test_oracle_coo_to_csr_csc_match_reference_samples(name):
    s = SAMPLES[name]
    bp, bj, bx = oracle.coo_to_csr(s["nrow"], s["ncol"], s["coo_i"], s["coo_j"], s["coo_v"])
    assert bp.tolist() == s["row_pointers"]
    assert bj.tolist() == s["col_indices"]
    assert bx.tolist() == s["csr_values"]  # bit-exact: duplicates are summed in order of appearance
    cp, ci, cx = oracle.coo_to_csc(s["nrow"], s["ncol"], s["coo_i"], s["coo_j"], s["coo_v"])
    assert cp.tolist() == s["col_pointers"]
    assert ci.tolist() == s["row_indices"]
    assert cx.tolist() == s["csc_values"]


@pytest.mark.parametrize("name", sorted(SAMPLES))
def test_product_converter_matches_reference_samples(name):
    import russell_b200 as rb

    coo, s = helpers.sample_coo(name)
    csr = rb.CsrMatrix.from_coo(coo)
    n = csr.nnz
    assert csr.row_pointers.tolist() == s["row_pointers"]
    assert csr.col_indices[:n].tolist() == s["col_indices"]
    assert csr.values[:n].tolist() == s["csr_values"]
    assert len(csr.values) == coo.nnz  # buffer keeps nnz(dup) slots (csr_matrix.rs:339-340)
    csc = rb.CscMatrix.from_coo(coo)
    n = csc.nnz
    assert csc.col_pointers.tolist() == s["col_pointers"]
    assert csc.row_indices[:n].tolist() == s["row_indices"]
    assert csc.values[:n].tolist() == s["csc_values"]


def test_product_converter_random_with_duplicates_bit_exact():
    import russell_b200 as rb

    rng = np.random.default_rng(11)
    for nrow, ncol, nnz in [(1, 1, 5), (7, 3, 40), (50, 50, 900), (200, 180, 5000), (1000, 1000, 20000)]:
        ai = rng.integers(0, nrow, nnz).astype(np.int32)
        aj = rng.integers(0, ncol, nnz).astype(np.int32)
        ax = rng.standard_normal(nnz) * 10.0 ** rng.integers(-8, 8, nnz)
        coo = rb.CooMatrix.from_triplets(nrow, ncol, ai, aj, ax)
        csr = rb.CsrMatrix.from_coo(coo)
        bp, bj, bx = oracle.coo_to_csr(nrow, ncol, ai, aj, ax)
        n = csr.nnz
        assert np.array_equal(csr.row_pointers, bp)
        assert np.array_equal(csr.col_indices[:n], bj)
        assert np.array_equal(csr.values[:n], bx)  # bit-exact, same summation order
        # update_from_coo with new values keeps the structure
        coo.values[:] = rng.standard_normal(nnz)
        csr.update_from_coo(coo)
        bp2, bj2, bx2 = oracle.coo_to_csr(nrow, ncol, ai, aj, coo.values)
        assert np.array_equal(csr.row_pointers, bp2) and np.array_equal(csr.values[:n], bx2)


def test_coo_slot_map_reproduces_the_conversion_bit_exact():
    # the triplet -> CSR-slot map behind solver_b200_initialize_coo / k_coo_to_csr_values: summing the mapped
    # triplets of every slot in map order must equal CsrMatrix::update_from_coo (csr_matrix.rs:431-459) bit for bit
    import ctypes
    from russell_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(5)
    for n, nnz in [(1, 4), (9, 60), (120, 4000), (2000, 30000)]:
        ai = rng.integers(0, n, nnz).astype(np.int32)
        aj = rng.integers(0, n, nnz).astype(np.int32)
        ax = rng.standard_normal(nnz) * 10.0 ** rng.integers(-8, 8, nnz)
        ptr_, idx, val = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        sp, si = np.zeros(nnz + 1, np.int32), np.zeros(nnz, np.int32)
        P = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
        rc = lib.b200_coo_to_csr_map(n, n, nnz, P(ai, ctypes.c_int32), P(aj, ctypes.c_int32), P(ax, ctypes.c_double),
                                     P(ptr_, ctypes.c_int32), P(idx, ctypes.c_int32), P(val, ctypes.c_double),
                                     P(sp, ctypes.c_int32), P(si, ctypes.c_int32))
        assert rc == 0
        bp, bj, bx = oracle.coo_to_csr(n, n, ai, aj, ax)
        m = int(ptr_[n])
        assert np.array_equal(ptr_, bp) and np.array_equal(idx[:m], bj) and np.array_equal(val[:m], bx)
        assert sp[m] == nnz and sorted(si.tolist()) == list(range(nnz))  # every triplet lands in exactly one slot
        ax2 = rng.standard_normal(nnz)
        _, _, bx2 = oracle.coo_to_csr(n, n, ai, aj, ax2)
        for sl in range(m):
            acc = ax2[si[sp[sl]]]
            for t in range(sp[sl] + 1, sp[sl + 1]):
                acc += ax2[si[t]]
            assert acc == bx2[sl]
            assert np.all(np.diff(si[sp[sl]:sp[sl + 1]]) > 0)  # order of appearance


def test_converter_errors_mirror_reference():
    import russell_b200 as rb

    coo = rb.CooMatrix(2, 2, 3)
    with pytest.raises(rb.StrError, match="COO to CSR requires nnz > 0"):
        rb.CsrMatrix.from_coo(coo)
    coo.put(0, 0, 1.0)
    coo.put(1, 1, 1.0)
    csr = rb.CsrMatrix.from_coo(rb.SolverB200._trim(coo))
    other = rb.CooMatrix(3, 3, 2)
    other.put(0, 0, 1.0), other.put(1, 1, 1.0)
    with pytest.raises(rb.StrError, match="coo.nrow must be equal to csr.nrow"):
        csr.update_from_coo(other)
    with pytest.raises(rb.StrError, match="max number of items has been reached"):
        c = rb.CooMatrix(1, 1, 1)
        c.put(0, 0, 1.0)
        c.put(0, 0, 1.0)
    with pytest.raises(rb.StrError, match="j > i is incorrect for lower triangular storage"):
        rb.CooMatrix(2, 2, 2, rb.Sym.YesLower).put(0, 1, 1.0)
    with pytest.raises(rb.StrError, match="symmetric storage requires a square matrix"):
        rb.CooMatrix(2, 3, 2, rb.Sym.YesFull)


# ---- mat-vec + VerifyLinSys restatements ------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(SAMPLES))
def test_oracle_matvec_matches_dense(name):
    s = SAMPLES[name]
    rng = np.random.default_rng(3)
    u = rng.standard_normal(s["ncol"])
    mirror = s["sym"] in ("YesLower", "YesUpper")
    dense = np.zeros((s["nrow"], s["ncol"]))
    for i, j, v in zip(s["coo_i"], s["coo_j"], s["coo_v"]):
        dense[i, j] += v
        if mirror and i != j:
            dense[j, i] += v
    v1 = oracle.coo_matvec(s["nrow"], s["coo_i"], s["coo_j"], s["coo_v"], u, mirror)
    v2 = oracle.csr_matvec(s["row_pointers"], s["col_indices"], s["csr_values"], u, mirror)
    assert np.allclose(v1, dense @ u, rtol=1e-14, atol=1e-14)
    assert np.allclose(v2, dense @ u, rtol=1e-14, atol=1e-14)


def test_oracle_verify_matches_reference_doc_example():
    # verify_lin_sys.rs:31-57: a = [[1,0,4],[0,2,0],[0,0,3]], x = ones, rhs = [5,2,3]
    out = oracle.verify(3, [0, 0, 1, 2], [0, 2, 1, 2], [1.0, 4.0, 2.0, 3.0], np.ones(3), np.array([5.0, 2.0, 3.0]))
    assert out["max_abs_a"] == 4.0 and out["max_abs_ax"] == 5.0
    assert out["max_abs_diff"] == 0.0 and out["relative_error"] == 0.0


# ---- Matrix Market: product reader (C++) and oracle reader (python) --------------------------------------------
MM_BAD = {
    "bad_empty_file.mtx": "the file is empty",
    "bad_wrong_header.mtx": "after %%MatrixMarket, the first option must be \"matrix\"",
    "bad_wrong_dims.mtx": "found invalid (zero or negative) dimensions",
    "bad_missing_data.mtx": "not all values have been found",
    "bad_many_lines.mtx": "there are more values than specified",
    "bad_symmetric_rectangular.mtx": "MatrixMarket data is invalid: the number of rows must equal the number of columns for symmetric matrices",
    # this fixture uses 0-based indices; the reference's parser (read_matrix_market.rs:173-178) rejects it the same way
    "ok_rectangular.mtx": "found an invalid index",
}


@pytest.mark.parametrize("fn", sorted(MM_BAD))
def test_matrix_market_bad_files(fn):
    import russell_b200 as rb

    with pytest.raises(rb.StrError) as e:
        rb.read_matrix_market(helpers.mm_path(fn), rb.MMsym.LeaveAsLower)
    assert e.value.msg == MM_BAD[fn]
    if fn != "bad_wrong_header.mtx":  # the oracle reader only distinguishes the data-section errors
        with pytest.raises(ValueError) as e2:
            oracle.read_matrix_market(helpers.mm_path(fn))
        assert str(e2.value) == MM_BAD[fn]


@pytest.mark.parametrize("fn", ["ok_general.mtx", "ok_symmetric.mtx", "ok_symmetric_small.mtx", "ok_simple_general.mtx",
                                "ok_simple_symmetric.mtx", "umfpack_di_demo.mtx", "bfwb62.mtx"])
@pytest.mark.parametrize("handling", ["LeaveAsLower", "SwapToUpper", "MakeItFull"])
def test_matrix_market_ok_files_product_vs_oracle(fn, handling):
    import russell_b200 as rb

    coo = rb.read_matrix_market(helpers.mm_path(fn), rb.MMsym[handling])
    m, n, sym, ai, aj, ax = oracle.read_matrix_market(helpers.mm_path(fn), handling)
    assert (coo.nrow, coo.ncol, coo.symmetric.name, coo.nnz) == (m, n, sym, len(ax))
    assert np.array_equal(coo.indices_i[: coo.nnz], ai)
    assert np.array_equal(coo.indices_j[: coo.nnz], aj)
    assert np.array_equal(coo.values[: coo.nnz], ax)


def test_matrix_market_bfwb62_counts():
    # 202 stored lower entries -> 342 when made full
    import russell_b200 as rb

    assert rb.read_matrix_market(helpers.mm_path("bfwb62.mtx"), rb.MMsym.LeaveAsLower).nnz == 202
    full = rb.read_matrix_market(helpers.mm_path("bfwb62.mtx"), rb.MMsym.MakeItFull)
    assert full.nnz == 342 and full.symmetric == rb.Sym.YesFull


def test_matrix_market_simple_symmetric_doc_example():
    # read_matrix_market.rs:309-345 (ok_simple_symmetric.mtx): lower storage gives [[1,2,0],[2,3,4],[0,4,0]]
    import russell_b200 as rb

    coo = rb.read_matrix_market(helpers.mm_path("ok_simple_symmetric.mtx"), rb.MMsym.LeaveAsLower)
    assert coo.symmetric == rb.Sym.YesLower and coo.nnz == 4
    assert np.array_equal(coo.as_dense(), np.array([[1.0, 2, 0], [2, 3, 4], [0, 4, 0]]))


# ---- the solver oracle (SuperLU stand-in) pinned to the reference's known answers ---------------------------------
KATS = [  # (sample, rhs, x_correct, tol, source)
    ("umfpack_unsymmetric_5x5", [8.0, 45.0, -3.0, 3.0, 19.0], [1.0, 2.0, 3.0, 4.0, 5.0], 1e-14, "solver_umfpack.rs:660-672"),
    ("mkl_symmetric_5x5_full", [1.0, 2.0, 3.0, 4.0, 5.0], [-979.0 / 3.0, 983.0, 1961.0 / 12.0, 398.0, 123.0 / 2.0], 1e-10, "solver_umfpack.rs:689-702"),
    ("mkl_symmetric_5x5_lower(true,false)", [1.0, 2.0, 3.0, 4.0, 5.0], [-979.0 / 3.0, 983.0, 1961.0 / 12.0, 398.0, 123.0 / 2.0], 1e-10, "lin_solver.rs:241-248"),
    ("mkl_positive_definite_5x5_lower", [1.0, 2.0, 3.0, 4.0, 5.0], [-979.0 / 3.0, 983.0, 1961.0 / 12.0, 398.0, 123.0 / 2.0], 1e-10, "solver_cudss.rs:800-824"),
]


def _kat_system(name):
    s = SAMPLES[name]
    a = oracle.full_scipy_matrix(s["nrow"], s["ncol"], s["coo_i"], s["coo_j"], s["coo_v"], s["sym"])
    return s, a


@pytest.mark.parametrize("name,rhs,xc,tol,src", KATS)
def test_oracle_lu_reproduces_reference_known_answers(name, rhs, xc, tol, src):
    s, a = _kat_system(name)
    x = oracle.lu_solve(a, np.array(rhs))
    assert np.max(np.abs(x - np.array(xc))) <= tol * max(1.0, np.max(np.abs(xc))), src


@pytest.mark.parametrize("name,rhs,xc,tol,src", KATS)
def test_host_walk_reproduces_reference_known_answers(name, rhs, xc, tol, src):
    s = SAMPLES[name]
    bp, bj, bx = oracle.coo_to_csr(s["nrow"], s["ncol"], s["coo_i"], s["coo_j"], s["coo_v"])
    rc, x, st = oracle.mf_solve(s["nrow"], bp, bj, bx, np.array(rhs), sym_lower=(s["sym"] == "YesLower"))
    assert rc == 0
    assert np.max(np.abs(x - np.array(xc))) <= tol * max(1.0, np.max(np.abs(xc))), src


def test_oracle_bfwb62_golden_x():
    # solve_matrix_market.rs:217-230,307-372: 62 golden values, abs tol 1e-10 on |x| ~ 1e5
    xg = helpers.load_bfwb62_x()
    for handling in ("MakeItFull", "LeaveAsLower"):
        m, n, sym, ai, aj, ax = oracle.read_matrix_market(helpers.mm_path("bfwb62.mtx"), handling)
        a = oracle.full_scipy_matrix(m, n, ai, aj, ax, sym)
        x = oracle.lu_solve(a, np.ones(62))
        assert np.max(np.abs(x - xg)) <= 1e-10
        bp, bj, bx = oracle.coo_to_csr(m, n, ai, aj, ax)
        rc, xw, st = oracle.mf_solve(n, bp, bj, bx, np.ones(62), sym_lower=(sym == "YesLower"))
        assert rc == 0 and np.max(np.abs(xw - xg)) <= 1e-10
        v = oracle.verify(m, ai, aj, ax, xw, np.ones(62), mirror=(sym == "YesLower"))
        assert v["relative_error"] <= 1e-14  # README run reports 5.55e-16 (russell_sparse/README.md:266-271)


def test_oracle_diagonal_10x10():
    # tests/test_umfpack.rs: a_kk = 10 + k, x = k, tol 1e-14
    n = 10
    akk = 10.0 + np.arange(n) * (n / 10.0)
    xc = np.arange(n, dtype=float)
    a = oracle.full_scipy_matrix(n, n, np.arange(n), np.arange(n), akk)
    assert np.max(np.abs(oracle.lu_solve(a, akk * xc) - xc)) <= 1e-14
    bp, bj, bx = oracle.coo_to_csr(n, n, np.arange(n), np.arange(n), akk)
    rc, x, _ = oracle.mf_solve(n, bp, bj, bx, akk * xc)
    assert rc == 0 and np.max(np.abs(x - xc)) <= 1e-14


def newton_residual(u):
    d1, d2, d3, d4 = u
    return np.array([
        2.0 * d1 + d1**4 + d2 + 3.0 * d1 * d2 * d2 - 9.0 * d4 + d4**4 - 0.2,
        d1 + 3.0 * d1 * d1 * d2 + 10.0 * d2 + 4.0 * d2 * d2 + 2.0 * d2 * d3 - 8.0 * d3 + 7.0 * d4 + 0.1,
        -8.0 * d2 + d2 * d2 + 3.0 * d3 + d3 * d3 + 2.0 * d4,
        -9.0 * d1 + 4.0 * d1 * d4**3 + 7.0 * d2 + 2.0 * d3 + 5.0 * d4 - 0.5,
    ])


def newton_jacobian_triplets(u):
    d1, d2, d3, d4 = u
    rows = [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3]
    cols = [0, 1, 2, 3] * 4
    vals = [2.0 + 4.0 * d1**3 + 3.0 * d2 * d2, 1.0 + 6.0 * d1 * d2, 0.0, -9.0 + 4.0 * d4**3,
            1.0 + 6.0 * d1 * d2, 10.0 + 3.0 * d1 * d1 + 8.0 * d2 + 2.0 * d3, -8.0 + 2.0 * d2, 7.0,
            0.0, -8.0 + 2.0 * d2, 3.0 + 2.0 * d3, 2.0,
            -9.0 + 4.0 * d4**3, 7.0, 2.0, 5.0 + 12.0 * d1 * d4 * d4]
    return rows, cols, vals


NEWTON_REF = [  # tests/test_nonlinear_system.rs:72-79
    [0.000000, 0.000000, 0.000000, 0.000000],
    [-0.236393, -0.106230, -0.225574, -0.086557],
    [-0.196773, -0.079071, -0.171604, -0.074904],
    [-0.194395, -0.077412, -0.168376, -0.074249],
    [-0.194386, -0.077406, -0.168364, -0.074246],
    [-0.194386, -0.077406, -0.168364, -0.074246],
]


def run_newton(solve):
    """tests/test_nonlinear_system.rs:60-111: returns the iteration count; asserts the iterate table at 1e-6"""
    u = np.zeros(4)
    norm0 = 1.0
    it = 0
    while it < 10:
        r = newton_residual(u)
        err = 1.0 if it == 0 else np.linalg.norm(r) / norm0
        if it == 0:
            norm0 = np.linalg.norm(r)
        assert np.max(np.abs(u - np.array(NEWTON_REF[it]))) <= 1e-6
        if err < 1e-13:
            break
        mdu = solve(newton_jacobian_triplets(u), r, it)
        u = u - mdu
        it += 1
    return it


def test_oracle_newton_iterate_table():
    def solve(trip, r, it):
        rows, cols, vals = trip
        return oracle.lu_solve(oracle.full_scipy_matrix(4, 4, rows, cols, vals), r)

    assert run_newton(solve) == 5


def test_host_walk_newton_iterate_table():
    def solve(trip, r, it):
        rows, cols, vals = trip
        bp, bj, bx = oracle.coo_to_csr(4, 4, rows, cols, vals)
        rc, x, _ = oracle.mf_solve(4, bp, bj, bx, r)
        assert rc == 0
        return x

    assert run_newton(solve) == 5


# ---- host analysis + front walk vs the independent LU on varied systems -------------------------------------
def _walk_vs_lu(n, ai, aj, ax, sym="No", **kw):
    bp, bj, bx = oracle.coo_to_csr(n, n, ai, aj, ax)
    b = np.sin(np.arange(n) + 1.0)
    rc, x, st = oracle.mf_solve(n, bp, bj, bx, b, sym_lower=(sym == "YesLower"), **kw)
    a = oracle.full_scipy_matrix(n, n, ai, aj, ax, sym)
    xs = oracle.lu_solve(a, b)
    res = np.linalg.norm(b - a @ x) / np.linalg.norm(b)
    return rc, x, xs, res, st


@pytest.mark.parametrize("k", [1, 2, 3, 7, 16, 33, 64, 120])
@pytest.mark.parametrize("lower", [False, True])
def test_host_walk_laplacian(k, lower):
    n, ai, aj, ax = helpers.laplacian_2d_triplets(k, lower)
    rc, x, xs, res, st = _walk_vs_lu(n, ai, aj, ax, "YesLower" if lower else "No")
    assert rc == 0 and res <= 1e-12
    assert np.max(np.abs(x - xs)) <= 1e-9 * np.max(np.abs(xs))


@pytest.mark.parametrize("panel_width,nd_leaf", [(4, 4), (8, 16), (16, 8), (64, 96), (64, 400)])
def test_host_walk_panel_and_leaf_sizes(panel_width, nd_leaf):
    n, ai, aj, ax = helpers.convection_diffusion_triplets(40)
    rc, x, xs, res, st = _walk_vs_lu(n, ai, aj, ax, panel_width=panel_width, nd_leaf=nd_leaf)
    assert rc == 0 and res <= 1e-12
    assert np.max(np.abs(x - xs)) <= 1e-9 * np.max(np.abs(xs))


@pytest.mark.parametrize("ordering", [0, 1, 2])
def test_host_walk_orderings(ordering):
    n, ai, aj, ax = helpers.laplacian_2d_triplets(25)
    rc, x, xs, res, st = _walk_vs_lu(n, ai, aj, ax, ordering=ordering)
    assert rc == 0 and res <= 1e-12


def test_host_walk_saddle_point_needs_matching():
    n, ai, aj, ax = helpers.saddle_point_triplets(24)
    rc, x, xs, res, st = _walk_vs_lu(n, ai, aj, ax, matching=2)
    assert rc == 0 and res <= 1e-10
    assert np.max(np.abs(x - xs)) <= 1e-8 * np.max(np.abs(xs))


def test_host_walk_random_unsymmetric_zero_diagonal():
    rng = np.random.default_rng(5)
    n = 300
    import scipy.sparse as sp

    a = sp.random(n, n, density=0.02, random_state=5, format="coo")
    perm = rng.permutation(n)  # a hidden permutation carries the large entries: the diagonal itself is empty
    ai = np.concatenate([a.row, np.arange(n)])
    aj = np.concatenate([a.col, perm])
    ax = np.concatenate([a.data, 10.0 + rng.random(n)])
    keep = ai != aj
    rc, x, xs, res, st = _walk_vs_lu(n, ai[keep].astype(np.int32), aj[keep].astype(np.int32), ax[keep], matching=2)
    assert rc == 0 and res <= 1e-10


def test_host_walk_singular_is_reported():
    # solver_umfpack.rs:624-630: diag(1, 0) -> "Error(1): Matrix is singular"
    bp, bj, bx = oracle.coo_to_csr(2, 2, [0, 1], [0, 1], [1.0, 0.0])
    rc, x, st = oracle.mf_solve(2, bp, bj, bx, np.ones(2))
    assert rc == 1


def test_host_walk_disconnected_and_diagonal():
    n = 500
    rc, x, xs, res, st = _walk_vs_lu(n, np.arange(n, dtype=np.int32), np.arange(n, dtype=np.int32), 1.0 + np.arange(n))
    assert rc == 0 and res <= 1e-14
    # two independent grids
    n1, ai, aj, ax = helpers.laplacian_2d_triplets(12)
    ai2 = np.concatenate([ai, ai + n1])
    aj2 = np.concatenate([aj, aj + n1])
    rc, x, xs, res, st = _walk_vs_lu(2 * n1, ai2, aj2, np.concatenate([ax, 2 * ax]))
    assert rc == 0 and res <= 1e-12


def test_analysis_stats_are_sane_at_scale():
    # 250 x 250 grid: fill must stay in the range of a good minimum-degree ordering (SuperLU MMD 3.2e6)
    import russell_b200 as rb

    coo = helpers.laplacian_2d_coo(250)
    csr = rb.CsrMatrix.from_coo(coo)
    rc, st = oracle.mf_analyze(csr.nrow, csr.pointers, csr.indices, csr.values)
    assert rc == 0
    assert st["nnz_l"] + st["nnz_u"] < 6.0e6
    assert st["max_front"] < 600


def test_threaded_nested_dissection_equals_serial():
    # ordering.cpp orders the two halves of the first three dissection levels on separate threads; the permutation must
    # not depend on the schedule.  The serial run happens in a child process (the switch is read once per process).
    import os
    import subprocess
    import sys

    HERE = os.path.dirname(os.path.abspath(__file__))
    ROOT = os.path.dirname(HERE)
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, helpers\nfrom oracle import oracle\n"
            "n, ai, aj, ax = helpers.convection_diffusion_triplets(260)\n"
            "bp, bj, bx = oracle.coo_to_csr(n, n, ai, aj, ax)\n"
            "rc, x, st = oracle.mf_solve(n, bp, bj, bx, np.ones(n))\n"
            "print(rc, float(x.sum()).hex(), float(np.abs(x).max()).hex(), int(st['nnz_l']), int(st['nlevels']))\n"
            % (ROOT, HERE))
    outs = []
    # B200_PAR_FLOOR = 0 forces every threaded stage of symbolic.cpp (row-parallel loops, subtree tasks of the row structures,
    # the helper thread of the contribution-block allocator) at this size
    for env_extra in ({}, {"B200_ND_SERIAL": "1"}, {"B200_PAR_FLOOR": "0"}):
        env = dict(os.environ, **env_extra)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip())
    assert outs[0] == outs[1] == outs[2], outs
    assert outs[0].startswith("0 ")


def test_threaded_analysis_builds_the_same_plan_bit_for_bit():
    # every array of the plan (permutations, front tree, row lists, relative indices, storage offsets incl. the contribution
    # arena of the helper-thread allocator, scatter map) hashed in child processes: default threading, everything serial,
    # every threaded stage forced (B200_PAR_FLOOR = 0)
    import os
    import subprocess
    import sys

    HERE = os.path.dirname(os.path.abspath(__file__))
    ROOT = os.path.dirname(HERE)
    code = ("import sys, ctypes; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, helpers\nfrom oracle import oracle\n"
            "oracle.build()\n"
            "lib = ctypes.CDLL(%r)\n"
            "ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)\n"
            "lib.oracle_plan_create.restype = ctypes.c_void_p\n"
            "lib.oracle_plan_create.argtypes = [ctypes.c_int, ip, ip, dp] + [ctypes.c_int] * 5 + [ip]\n"
            "lib.oracle_plan_hash.restype = ctypes.c_ulonglong\n"
            "lib.oracle_plan_hash.argtypes = [ctypes.c_void_p]\n"
            "out = []\n"
            "for gen in (lambda: helpers.convection_diffusion_triplets(300), lambda: helpers.laplacian_3d_triplets(30), lambda: helpers.saddle_point_triplets(120)):\n"
            "    n, ai, aj, ax = gen()\n"
            "    bp, bj, bx = oracle.coo_to_csr(n, n, ai, aj, ax)\n"
            "    nn = ctypes.c_int(0)\n"
            "    h = lib.oracle_plan_create(n, bp.ctypes.data_as(ip), bj.ctypes.data_as(ip), bx.ctypes.data_as(dp), 0, 0, 2, 0, 0, ctypes.byref(nn))\n"
            "    assert h\n"
            "    out.append('%%d:%%016x' %% (nn.value, lib.oracle_plan_hash(h)))\n"
            "print(' '.join(out))\n"
            % (ROOT, HERE, os.path.join(ROOT, "oracle", "_build", "liboracle_mf.so")))
    outs = []
    for env_extra in ({}, {"B200_ND_SERIAL": "1"}, {"B200_PAR_FLOOR": "0"}):
        env = dict(os.environ, **env_extra)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip())
    assert outs[0] == outs[1] == outs[2] and len(outs[0].split()) == 3, outs


def test_host_walk_unstructured_graphs():
    # irregular graphs (k-nearest-neighbour meshes in 2D and 3D): the level-set separators are shrunk by the
    # minimum-vertex-cover refinement (ordering.cpp: shrink_separator_by_cover); the walk must still solve the system
    import scipy.sparse as sp
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(1)
    for npts, dim in ((6000, 2), (4000, 3)):
        pts = rng.random((npts, dim))
        _, idx = cKDTree(pts).query(pts, k=7)
        rows, cols = np.repeat(np.arange(npts), 6), idx[:, 1:].ravel()
        g = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(npts, npts))
        g = (g + g.T).tocsr()
        g.data[:] = -1.0
        a = (g + sp.diags(np.asarray(-g.sum(axis=1)).ravel() + 1.0)).tocsr()
        a.sort_indices()
        b = np.ones(npts)
        rc, x, st = oracle.mf_solve(npts, a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data, b)
        assert rc == 0
        assert np.linalg.norm(b - a @ x) / np.linalg.norm(b) <= 1e-10
        xs = oracle.lu_solve(a.tocsc(), b)
        assert np.max(np.abs(x - xs)) <= 1e-9 * np.max(np.abs(xs))
