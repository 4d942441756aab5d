"""CPU tests of the Complex64 twin: host formats against the reference's complex fixtures,
the oracle's complex functions pinned to the reference's complex known answers, the real embedding the CUDA path
factorizes (checked densely and by the scalar walk of the front plan), and the Radau5/Brusselator matrix generator.
No CUDA compute here."""
import ctypes

import numpy as np
import pytest

import helpers
import russell_b200 as rb
from oracle import oracle
from russell_b200 import _lib
from russell_b200._lib import p_f64, p_i32, p_i64, ptr

CS = helpers.load_complex_samples()

# (sample, rhs, x_correct, tolerance, reference test)
COMPLEX_KATS = [
    ("complex_symmetric_3x3_lower", [-3 + 3j, 2 - 2j, 9 + 7j], [1 + 1j, 2 - 2j, 3 + 3j], 1e-14,
     "complex_lin_solver.rs:199-204"),
    ("complex_symmetric_3x3_full", [-3 + 3j, 2 - 2j, 9 + 7j], [1 + 1j, 2 - 2j, 3 + 3j], 1e-14,
     "complex_solver_umfpack.rs:600-610, complex_lin_solver.rs:221-226"),
    ("umfpack_complex_unsymmetric_5x5", [8.0, 45.0, -3.0, 3.0, 19.0], [1.0, 2.0, 3.0, 4.0, 5.0], 1e-12,
     "complex_solver_cudss.rs:556-585 (colamd/matching variants reach 1e-12)"),
    ("mkl_complex_positive_definite_5x5_lower", [1.0, 2.0, 3.0, 4.0, 5.0],
     [-979.0 / 3.0, 983.0, 1961.0 / 12.0, 398.0, 123.0 / 2.0], 1e-10, "complex_solver_cudss.rs:653-677"),
]


def dense_of(s):
    a = np.zeros((s["nrow"], s["ncol"]), dtype=np.complex128)
    for i, j, v in zip(s["coo_i"], s["coo_j"], s["coo_v"]):
        a[i, j] += complex(*v)
        if s["sym"] in ("YesLower", "YesUpper") and i != j:
            a[j, i] += complex(*v)
    return a


def embed(csr, lower):
    """calls the product's b200_complex_embed (the host half of complex_solver_b200_initialize)"""
    lib = _lib.load()
    n, nnz = csr.nrow, csr.nnz
    vals = np.ascontiguousarray(csr.values[:nnz])
    info = np.zeros(2, dtype=np.int64)
    rc = lib.b200_complex_embed(n, ptr(csr.pointers, p_i32), ptr(csr.indices, p_i32), ptr(vals, p_f64), int(lower),
                                ptr(info, p_i64), None, None, None, None)
    assert rc == 0
    nreal = int(info[1])
    rptr = np.zeros(2 * n + 1, dtype=np.int32)
    rcol = np.zeros(nreal, dtype=np.int32)
    code = np.zeros(nreal, dtype=np.int32)
    rval = np.zeros(nreal)
    rc = lib.b200_complex_embed(n, ptr(csr.pointers, p_i32), ptr(csr.indices, p_i32), ptr(vals, p_f64), int(lower),
                                ptr(info, p_i64), ptr(rptr, p_i32), ptr(rcol, p_i32), ptr(code, p_i32), ptr(rval, p_f64))
    assert rc == 0
    return rptr, rcol, code, rval, vals


@pytest.mark.parametrize("name", sorted(CS))
def test_complex_coo_to_csr_csc_match_reference_fixtures_bit_exactly(name):
    coo, s = helpers.complex_sample_coo(name)
    want_csr = np.array([complex(*v) for v in s["csr_values"]])
    want_csc = np.array([complex(*v) for v in s["csc_values"]])
    csr = rb.ComplexCsrMatrix.from_coo(coo)
    n = csr.nnz
    assert list(csr.row_pointers) == s["row_pointers"] and list(csr.col_indices[:n]) == s["col_indices"]
    assert np.array_equal(csr.values[:n], want_csr)
    csc = rb.ComplexCscMatrix.from_coo(coo)
    assert list(csc.col_pointers) == s["col_pointers"] and list(csc.row_indices[:n]) == s["row_indices"]
    assert np.array_equal(csc.values[:n], want_csc)
    # the oracle's restatement agrees with the same fixtures (pins the oracle)
    bp, bj, bx = oracle.complex_coo_to_csr(s["nrow"], s["ncol"], s["coo_i"], s["coo_j"], [complex(*v) for v in s["coo_v"]])
    assert list(bp) == s["row_pointers"] and list(bj) == s["col_indices"] and np.array_equal(bx, want_csr)


def test_complex_coo_guards():
    # same guards and messages as the real CooMatrix (coo_matrix.rs:173-198,324-352)
    with pytest.raises(rb.StrError, match="nrow must be ≥ 1"):
        rb.ComplexCooMatrix(0, 1, 1)
    coo = rb.ComplexCooMatrix(2, 2, 1, rb.Sym.YesLower)
    with pytest.raises(rb.StrError, match="j > i is incorrect for lower triangular storage"):
        coo.put(0, 1, 1j)
    coo.put(1, 0, 2 + 1j)
    with pytest.raises(rb.StrError, match="max number of items has been reached"):
        coo.put(1, 1, 1.0)
    assert coo.get_values()[0] == 2 + 1j and coo.values.dtype == np.complex128


def test_oracle_complex_verify_matches_reference_numbers():
    # verify_lin_sys.rs:253-275 (new_complex_matrix_works)
    s = CS["complex_rectangular_4x3"]
    vals = [complex(*v) for v in s["coo_v"]]
    x = np.array([1 + 2j, 2 - 1j, 1j])
    rhs = np.array([-6 + 14j, -1 + 2j, 14 + 6j, 1 + 2j])
    v = oracle.complex_verify(4, s["coo_i"], s["coo_j"], vals, x, rhs)
    assert abs(v["max_abs_a"] - 7.0710678118654755) <= 1e-15 and abs(v["max_abs_ax"] - 15.231546211727817) <= 1e-15
    assert v["max_abs_diff"] <= 1e-15 and v["relative_error"] <= 1e-15
    rhs[3] = 1.0
    v = oracle.complex_verify(4, s["coo_i"], s["coo_j"], vals, x, rhs)
    assert abs(v["max_abs_diff"] - 2.0) <= 1e-15 and abs(v["relative_error"] - 2.0 / (7.0710678118654755 + 1.0)) <= 1e-15


@pytest.mark.parametrize("name,rhs,xc,tol,src", COMPLEX_KATS)
def test_oracle_complex_lu_is_pinned_to_the_reference_known_answers(name, rhs, xc, tol, src):
    s = CS[name]
    a = dense_of(s)
    import scipy.sparse as sp

    x = oracle.lu_solve(sp.csc_matrix(a), np.array(rhs, dtype=np.complex128))
    assert np.max(np.abs(x - np.array(xc))) <= tol * max(1.0, np.max(np.abs(xc))), src
    # determinant fixture of the sample (samples.rs) agrees with the dense matrix
    assert abs(np.linalg.det(a) - complex(*s["det"])) <= 1e-12 * abs(complex(*s["det"]))


@pytest.mark.parametrize("name", [k for k in sorted(CS) if CS[k]["nrow"] == CS[k]["ncol"] and CS[k]["sym"] != "YesUpper"])
def test_embedding_is_the_2x2_block_matrix(name):
    coo, s = helpers.complex_sample_coo(name)
    csr = rb.ComplexCsrMatrix.from_coo(coo)
    rptr, rcol, code, rval, vals = embed(csr, s["sym"] == "YesLower")
    a = dense_of(s)
    n2 = 2 * s["nrow"]
    want = np.zeros((n2, n2))
    want[0::2, 0::2], want[0::2, 1::2], want[1::2, 0::2], want[1::2, 1::2] = a.real, -a.imag, a.imag, a.real
    got = np.zeros((n2, n2))
    for i in range(n2):
        cols = rcol[rptr[i]:rptr[i + 1]]
        assert np.all(np.diff(cols) > 0)  # sorted rows, no duplicates: the CSR contract of solver_b200_initialize
        got[i, cols] = rval[rptr[i]:rptr[i + 1]]
    assert np.array_equal(got, want)
    # the slot map k_complex_expand applies on the device reproduces the values bit for bit
    flat = vals.view(np.float64)
    k, src = code & 3, code >> 2
    assert np.array_equal(np.where(k == 1, -flat[2 * src + 1], np.where(k == 2, flat[2 * src + 1], flat[2 * src])), rval)


def test_embed_rejects_bad_csr():
    lib = _lib.load()
    info = np.zeros(2, dtype=np.int64)
    rp = np.array([0, 2, 3], dtype=np.int32)
    v = np.zeros(6)
    for ci in ([1, 0, 1], [0, 0, 1], [0, 2, 1]):  # unsorted, duplicate, out of range
        c = np.array(ci, dtype=np.int32)
        assert lib.b200_complex_embed(2, ptr(rp, p_i32), ptr(c, p_i32), ptr(v, p_f64), 0, ptr(info, p_i64), None, None, None, None) == -1
    c = np.array([0, 1, 1], dtype=np.int32)  # j > i with the lower-triangle promise
    assert lib.b200_complex_embed(2, ptr(rp, p_i32), ptr(c, p_i32), ptr(v, p_f64), 1, ptr(info, p_i64), None, None, None, None) == -1
    assert lib.b200_complex_embed(2, ptr(rp, p_i32), ptr(c, p_i32), ptr(v, p_f64), 0, ptr(info, p_i64), None, None, None, None) == 0
    assert list(info) == [3, 12]


@pytest.mark.parametrize("name,rhs,xc,tol,src", COMPLEX_KATS)
def test_scalar_walk_of_the_embedded_system_reproduces_the_known_answers(name, rhs, xc, tol, src):
    # the same host analysis + front plan the CUDA kernels execute, walked on the CPU (oracle/mf_host.cpp), applied to
    # the embedded real system: validates matching/ordering on matrices with explicit-zero block entries
    coo, s = helpers.complex_sample_coo(name)
    csr = rb.ComplexCsrMatrix.from_coo(coo)
    rptr, rcol, _, rval, _ = embed(csr, s["sym"] == "YesLower")
    b = np.array(rhs, dtype=np.complex128)
    rc, x, st = oracle.mf_solve(2 * s["nrow"], rptr, rcol, rval, b.view(np.float64))
    assert rc == 0
    assert np.max(np.abs(x.view(np.complex128) - np.array(xc))) <= tol * max(1.0, np.max(np.abs(xc))), src


def test_brusselator_radau5_matrices():
    # BASELINE.json configs[3] generator: sizes from SURVEY 8 (C4) and K = c I - J structure
    npoint = 9
    ndim, ai, aj, kr, kc = helpers.brusselator_radau5_triplets(npoint, h=1e-4)
    s = npoint * npoint
    assert ndim == 2 * s and len(ai) == 14 * s + ndim
    bp, bj, bx = oracle.coo_to_csr(ndim, ndim, ai, aj, kr)
    assert bp[-1] == 6 * ndim  # 6 unique entries per row (~3.0M unique at npoint 500)
    cp, cj, cx = oracle.complex_coo_to_csr(ndim, ndim, ai, aj, kc)
    assert np.array_equal(cp, bp) and np.array_equal(cj, bj)
    # same -J, different shift: K_comp - K_real = ((alpha - gamma) + i beta)/h on the diagonal only
    d = cx - bx
    diag = bj == np.repeat(np.arange(ndim), 6)
    assert np.allclose(d[~diag], 0.0) and np.allclose(d[diag], complex(helpers.RADAU5_ALPHA - helpers.RADAU5_GAMMA, helpers.RADAU5_BETA) / 1e-4)
    # periodic 5-point rows sum: row sums of -J's diffusion part vanish, so K_real row sum = gamma/h - reaction terms
    import scipy.sparse as sp

    a = sp.csr_matrix((bx, bj, bp), shape=(ndim, ndim))
    y = np.linspace(0, 1, npoint)
    um = np.tile(22.0 * y * (1 - y) ** 1.5, (npoint, 1)).T.ravel()  # U depends on y = j*dy, m = i + j*nx
    vm = np.tile(27.0 * y * (1 - y) ** 1.5, (npoint, 1)).ravel()    # V depends on x = i*dx
    rs = np.asarray(a.sum(axis=1)).ravel()
    assert np.allclose(rs[:s], helpers.RADAU5_GAMMA / 1e-4 - (-4.4 + 2 * um * vm) - um * um, rtol=1e-12)
    # both systems are solved by the scalar walk (real) / embedded scalar walk (complex) to the north-star residual
    b = np.ones(ndim)
    rc, x, st = oracle.mf_solve(ndim, bp, bj, bx, b)
    assert rc == 0 and np.linalg.norm(b - a @ x) / np.linalg.norm(b) <= 1e-10


# ---- complex Matrix Market files (read_matrix_market.rs:401-437; tests :613-660, :684-816) ----------------------------------
def test_read_matrix_market_complex_general():
    coo_real, coo = rb.read_matrix_market_pair(helpers.mm_path("ok_complex_general.mtx"), rb.MMsym.LeaveAsLower)
    assert coo_real is None and coo.symmetric == rb.Sym.No
    assert (coo.nrow, coo.ncol, coo.nnz, coo.max_nnz) == (5, 5, 12, 12)
    assert list(coo.indices_i) == [0, 1, 0, 2, 4, 1, 2, 3, 4, 2, 1, 4]
    assert list(coo.indices_j) == [0, 0, 1, 1, 1, 2, 2, 2, 2, 3, 4, 4]
    want = [2 - 1j, 3 - 8j, 3 + 80j, -1 + 30j, 4 + 33j, 4 + 60j, -3 + 6j, 1 + 8j, 2 + 3j, 2 + 1j, 6 + 9j, 1 - 2j]
    assert np.array_equal(coo.values, np.array(want))
    real, cpx = rb.read_matrix_market_pair(helpers.mm_path("ok_general.mtx"), rb.MMsym.LeaveAsLower)
    assert cpx is None and real.nnz == 12


@pytest.mark.parametrize("handling,sym,ii,jj", [
    (rb.MMsym.LeaveAsLower, rb.Sym.YesLower, [0, 1, 2, 3, 3, 4, 4], [0, 0, 1, 2, 3, 1, 4]),
    (rb.MMsym.SwapToUpper, rb.Sym.YesUpper, [0, 0, 1, 2, 3, 1, 4], [0, 1, 2, 3, 3, 4, 4]),
])
def test_read_matrix_market_complex_symmetric(handling, sym, ii, jj):
    coo = rb.read_matrix_market(helpers.mm_path("ok_complex_symmetric_small.mtx"), handling)
    assert coo.symmetric == sym and (coo.nrow, coo.ncol, coo.nnz, coo.max_nnz) == (5, 5, 7, 7)
    assert list(coo.indices_i) == ii and list(coo.indices_j) == jj
    assert np.array_equal(coo.values, np.array([2 + 1j, 3 + 2j, -1 + 3j, 2 + 4j, 3 + 5j, 6 + 6j, 1 + 7j]))


def test_read_matrix_market_complex_make_it_full_and_hermitian():
    coo = rb.read_matrix_market(helpers.mm_path("ok_complex_symmetric_small.mtx"), rb.MMsym.MakeItFull)
    assert coo.symmetric == rb.Sym.YesFull and coo.max_nnz == 14 and coo.nnz == 11  # 3 diagonal + 2 * 4 off-diagonal entries
    a = coo.as_dense()
    assert np.array_equal(a, a.T) and a[1, 0] == 3 + 2j
    # "Hermitian" files are read like general ones (read_matrix_market.rs:86-93): the stored entries, no mirroring
    h = rb.read_matrix_market(helpers.mm_path("ok_complex_hermitian.mtx"), rb.MMsym.LeaveAsLower)
    assert h.symmetric == rb.Sym.No and h.values.dtype == np.complex128 and h.nnz >= 1


@pytest.mark.parametrize("name,msg", [
    ("bad_wrong_dims_complex.mtx", "found invalid (zero or negative) dimensions"),
    ("bad_missing_data_complex.mtx", "not all values have been found"),
    ("bad_many_lines_complex.mtx", "there are more values than specified"),
    ("bad_symmetric_rectangular_complex.mtx",
     "MatrixMarket data is invalid: the number of rows must equal the number of columns for symmetric matrices"),
    ("bad_not_complex_hermitian.mtx", '"Hermitian" keyword can only be used with the "complex" type'),
])
def test_read_matrix_market_complex_bad_files(name, msg):
    with pytest.raises(rb.StrError) as e:
        rb.read_matrix_market(helpers.mm_path(name), rb.MMsym.LeaveAsLower)
    assert str(e.value) == msg or msg in str(e.value)


def test_read_matrix_market_complex_value_errors(tmp_path):
    # parse_values (read_matrix_market.rs:166-171, tests :598-607): "cannot read bij" / "cannot parse bij"
    for body, msg in (("1 1 1.0\n", "cannot read bij"), ("1 1 1.0 wrong\n", "cannot parse bij")):
        f = tmp_path / "c.mtx"
        f.write_text("%%MatrixMarket matrix coordinate complex general\n2 2 1\n" + body)
        with pytest.raises(rb.StrError, match=msg):
            rb.read_matrix_market(str(f), rb.MMsym.LeaveAsLower)


# ---- randomized checks ----------------------------------------------------------------------------------------------
def test_complex_converter_random_with_duplicates_bit_exact():
    # product (C++ stable sort + in-order sums) vs the oracle's restatement on random triplets with many duplicates
    rng = np.random.default_rng(11)
    for nrow, ncol, nnz in ((7, 5, 60), (40, 40, 900), (1, 9, 30), (300, 300, 5000)):
        ai = rng.integers(0, nrow, nnz).astype(np.int32)
        aj = rng.integers(0, ncol, nnz).astype(np.int32)
        av = rng.standard_normal(nnz) + 1j * rng.standard_normal(nnz)
        coo = rb.ComplexCooMatrix.from_triplets(nrow, ncol, ai, aj, av)
        csr = rb.ComplexCsrMatrix.from_coo(coo)
        bp, bj, bx = oracle.complex_coo_to_csr(nrow, ncol, ai, aj, av)
        k = csr.nnz
        assert np.array_equal(csr.pointers, bp) and np.array_equal(csr.indices[:k], bj) and np.array_equal(csr.values[:k], bx)
        csc = rb.ComplexCscMatrix.from_coo(coo)
        cp, cj, cx = oracle.complex_coo_to_csr(ncol, nrow, aj, ai, av)  # columns of A = rows of A^T
        assert np.array_equal(csc.pointers, cp) and np.array_equal(csc.indices[:k], cj) and np.array_equal(csc.values[:k], cx)
        # and the mat-vec of the oracle agrees with the dense product
        u = rng.standard_normal(ncol) + 1j * rng.standard_normal(ncol)
        dense = np.zeros((nrow, ncol), dtype=np.complex128)
        np.add.at(dense, (ai, aj), av)
        assert np.allclose(oracle.complex_coo_matvec(nrow, ai, aj, av, u), dense @ u, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("seed,lower", [(0, False), (1, True), (2, False)])
def test_scalar_walk_of_random_complex_systems(seed, lower):
    # random sparse complex matrices (unsymmetric with an empty diagonal, or complex symmetric in lower storage): embedding
    # + host analysis + scalar walk against the complex CPU LU
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    n = 120
    a = sp.random(n, n, density=0.04, random_state=seed, format="coo")
    ai, aj = a.row.astype(np.int32), a.col.astype(np.int32)
    av = a.data + 1j * rng.standard_normal(len(a.data))
    if lower:
        keep = aj < ai
        ai, aj, av = ai[keep], aj[keep], av[keep]
        ai = np.concatenate([ai, np.arange(n, dtype=np.int32)])
        aj = np.concatenate([aj, np.arange(n, dtype=np.int32)])
        av = np.concatenate([av, 4.0 + 2.0j + rng.standard_normal(n)])
        sym = rb.Sym.YesLower
    else:
        perm = rng.permutation(n).astype(np.int32)  # the heavy entries sit on a hidden permutation, not on the diagonal
        keep = ai != aj
        ai = np.concatenate([ai[keep], np.arange(n, dtype=np.int32)])
        aj = np.concatenate([aj[keep], perm])
        av = np.concatenate([av[keep], (8.0 + rng.random(n)) * np.exp(1j * rng.random(n))])
        keep = ai != aj
        ai, aj, av = ai[keep], aj[keep], av[keep]
        sym = rb.Sym.No
    coo = rb.ComplexCooMatrix.from_triplets(n, n, ai, aj, av, sym)
    csr = rb.ComplexCsrMatrix.from_coo(coo)
    rptr, rcol, _, rval, _ = embed(csr, lower)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    rc, x, st = oracle.mf_solve(2 * n, rptr, rcol, rval, b.view(np.float64))
    assert rc == 0
    dense = coo.as_dense()
    z = x.view(np.complex128)
    assert np.linalg.norm(b - dense @ z) / np.linalg.norm(b) <= 1e-10
    zs = oracle.lu_solve(sp.csc_matrix(dense), b)
    assert np.max(np.abs(z - zs)) <= 1e-8 * np.max(np.abs(zs))
