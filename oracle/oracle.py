"""oracle.py -- TEST INFRASTRUCTURE ONLY.  CPU checker for the B200 sparse-solver hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
The product (russell_b200 + libsolver_b200.so) never does.

What the oracle restates (each function cites the reference lines it follows):
  * host formats  -> oracle/formats.c   (COO->CSR with duplicate summation, CSR/COO mat-vec, VerifyLinSys)
  * Matrix Market -> read_matrix_market() below (read_matrix_market.rs:44-178,346-475)
  * sparse LU     -> the reference's arithmetic lives in SuiteSparse/UMFPACK, an UNVENDORED, UNPINNED dependency
                     (russell_sparse/src/util.rs:49; call sites interface_umfpack.c:109,167,229) that is not
                     installed here.  The independent CPU LU used as the solver oracle is SuperLU through
                     scipy.sparse.linalg.splu -- a different algorithm, so parity is on x and on the residual,
                     never on L/U/pivot sequences.  It is pinned against every known-answer vector the reference's
                     own tests hold for this path (tests/test_oracle.py: samples 5x5, bfwb62's 62 golden values,
                     10x10 diagonal, Newton iterate table).
  * front walk    -> oracle/mf_host.cpp (scalar CPU execution of the SAME front plan the CUDA kernels run;
                     compares factor panels array-by-array on the GPU box).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def build():
    """compiles the C/C++ checkers (gcc/g++ only, seconds)"""
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load(name):
    path = os.path.join(_BUILD, name)
    if not os.path.exists(path):
        build()
    return ctypes.CDLL(path)


_fmt = None
_mf = None


def fmt():
    global _fmt
    if _fmt is None:
        lib = _load("liboracle_fmt.so")
        lib.oracle_coo_to_csr.argtypes = [ctypes.c_int32] * 3 + [_i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
        lib.oracle_coo_to_csr.restype = ctypes.c_int32
        lib.oracle_csr_matvec.argtypes = [ctypes.c_int32, _i32p, _i32p, _f64p, ctypes.c_int32, ctypes.c_double, _f64p, _f64p]
        lib.oracle_coo_matvec.argtypes = [ctypes.c_int32, ctypes.c_int32, _i32p, _i32p, _f64p, ctypes.c_int32, ctypes.c_double, _f64p, _f64p]
        lib.oracle_verify.argtypes = [ctypes.c_int32, ctypes.c_int32, _i32p, _i32p, _f64p, ctypes.c_int32, _f64p, _f64p, _f64p]
        lib.oracle_complex_coo_to_csr.argtypes = [ctypes.c_int32] * 3 + [_i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
        lib.oracle_complex_coo_to_csr.restype = ctypes.c_int32
        lib.oracle_complex_coo_matvec.argtypes = [ctypes.c_int32, ctypes.c_int32, _i32p, _i32p, _f64p, ctypes.c_int32, _f64p, _f64p]
        lib.oracle_complex_verify.argtypes = [ctypes.c_int32, ctypes.c_int32, _i32p, _i32p, _f64p, ctypes.c_int32, _f64p, _f64p, _f64p]
        _fmt = lib
    return _fmt


def mf():
    global _mf
    if _mf is None:
        lib = _load("liboracle_mf.so")
        lib.oracle_mf_solve.argtypes = [ctypes.c_int, _i32p, _i32p, _f64p] + [ctypes.c_int] * 6 + [ctypes.c_double, _f64p, _f64p, _f64p, ctypes.c_int]
        lib.oracle_mf_solve.restype = ctypes.c_int
        lib.oracle_mf_analyze.argtypes = [ctypes.c_int, _i32p, _i32p, _f64p] + [ctypes.c_int] * 5 + [_f64p, ctypes.c_int]
        lib.oracle_mf_analyze.restype = ctypes.c_int
        lib.oracle_mf_create.argtypes = [ctypes.c_int, _i32p, _i32p, _f64p] + [ctypes.c_int] * 5 + [ctypes.c_double, ctypes.POINTER(ctypes.c_int)]
        lib.oracle_mf_create.restype = ctypes.c_void_p
        lib.oracle_mf_sizes.argtypes = [ctypes.c_void_p, _i64p]
        lib.oracle_mf_get.argtypes = [ctypes.c_void_p, _f64p, _f64p, _i32p]
        lib.oracle_mf_handle_solve.argtypes = [ctypes.c_void_p, _f64p, _f64p, _f64p, ctypes.c_int, _f64p]
        lib.oracle_mf_free.argtypes = [ctypes.c_void_p]
        _mf = lib
    return _mf


def _p(a, t):
    return a.ctypes.data_as(t)


# ---- host formats --------------------------------------------------------------------------------------
def coo_to_csr(nrow, ncol, ai, aj, ax):
    """csr_matrix.rs:359-480 -> (row_pointers, col_indices[:nnz], values[:nnz])"""
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    ax = np.ascontiguousarray(ax, dtype=np.float64)
    nnz = len(ax)
    bp = np.zeros(nrow + 1, dtype=np.int32)
    bj = np.zeros(nnz, dtype=np.int32)
    bx = np.zeros(nnz, dtype=np.float64)
    rc = fmt().oracle_coo_to_csr(nrow, ncol, nnz, _p(ai, _i32p), _p(aj, _i32p), _p(ax, _f64p), _p(bp, _i32p), _p(bj, _i32p), _p(bx, _f64p))
    if rc != 0:
        raise ValueError("oracle_coo_to_csr failed: %d" % rc)
    n = bp[-1]
    return bp, bj[:n].copy(), bx[:n].copy()


def coo_to_csc(nrow, ncol, ai, aj, ax):
    """csc_matrix.rs:365-505: the column form is the row form of the transpose (same duplicate-sum order)"""
    return coo_to_csr(ncol, nrow, aj, ai, ax)


def csr_matvec(bp, bj, bx, u, mirror=False, alpha=1.0):
    """csr_matrix.rs:709-729"""
    nrow = len(bp) - 1
    bp = np.ascontiguousarray(bp, dtype=np.int32)
    bj = np.ascontiguousarray(bj, dtype=np.int32)
    bx = np.ascontiguousarray(bx, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.zeros(nrow)
    fmt().oracle_csr_matvec(nrow, _p(bp, _i32p), _p(bj, _i32p), _p(bx, _f64p), 1 if mirror else 0, alpha, _p(u, _f64p), _p(v, _f64p))
    return v


def coo_matvec(nrow, ai, aj, ax, u, mirror=False, alpha=1.0):
    """coo_matrix.rs:547-565"""
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    ax = np.ascontiguousarray(ax, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.zeros(nrow)
    fmt().oracle_coo_matvec(nrow, len(ax), _p(ai, _i32p), _p(aj, _i32p), _p(ax, _f64p), 1 if mirror else 0, alpha, _p(u, _f64p), _p(v, _f64p))
    return v


def verify(nrow, ai, aj, ax, x, rhs, mirror=False):
    """verify_lin_sys.rs:60-96 -> dict(max_abs_a, max_abs_ax, max_abs_diff, relative_error)"""
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    ax = np.ascontiguousarray(ax, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    out = np.zeros(4)
    fmt().oracle_verify(nrow, len(ax), _p(ai, _i32p), _p(aj, _i32p), _p(ax, _f64p), 1 if mirror else 0, _p(x, _f64p), _p(rhs, _f64p), _p(out, _f64p))
    return dict(max_abs_a=out[0], max_abs_ax=out[1], max_abs_diff=out[2], relative_error=out[3])


# ---- Complex64 twins --------------------------------------------------------------------------------------
def complex_coo_to_csr(nrow, ncol, ai, aj, ax):
    """ComplexCsrMatrix::update_from_coo (csr_matrix.rs:359-480 over Complex64) -> (row_pointers, col_indices, values)"""
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    ax = np.ascontiguousarray(ax, dtype=np.complex128)
    nnz = len(ax)
    bp = np.zeros(nrow + 1, dtype=np.int32)
    bj = np.zeros(nnz, dtype=np.int32)
    bx = np.zeros(nnz, dtype=np.complex128)
    rc = fmt().oracle_complex_coo_to_csr(nrow, ncol, nnz, _p(ai, _i32p), _p(aj, _i32p), _p(ax, _f64p), _p(bp, _i32p), _p(bj, _i32p), _p(bx, _f64p))
    if rc != 0:
        raise ValueError("oracle_complex_coo_to_csr failed: %d" % rc)
    n = bp[-1]
    return bp, bj[:n].copy(), bx[:n].copy()


def complex_coo_matvec(nrow, ai, aj, ax, u, mirror=False):
    """ComplexCooMatrix::mat_vec_mul (coo_matrix.rs:547-565)"""
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    ax = np.ascontiguousarray(ax, dtype=np.complex128)
    u = np.ascontiguousarray(u, dtype=np.complex128)
    v = np.zeros(nrow, dtype=np.complex128)
    fmt().oracle_complex_coo_matvec(nrow, len(ax), _p(ai, _i32p), _p(aj, _i32p), _p(ax, _f64p), 1 if mirror else 0, _p(u, _f64p), _p(v, _f64p))
    return v


def complex_verify(nrow, ai, aj, ax, x, rhs, mirror=False):
    """VerifyLinSys::from_complex (verify_lin_sys.rs:104-146)"""
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    ax = np.ascontiguousarray(ax, dtype=np.complex128)
    x = np.ascontiguousarray(x, dtype=np.complex128)
    rhs = np.ascontiguousarray(rhs, dtype=np.complex128)
    out = np.zeros(4)
    fmt().oracle_complex_verify(nrow, len(ax), _p(ai, _i32p), _p(aj, _i32p), _p(ax, _f64p), 1 if mirror else 0, _p(x, _f64p), _p(rhs, _f64p), _p(out, _f64p))
    return dict(max_abs_a=out[0], max_abs_ax=out[1], max_abs_diff=out[2], relative_error=out[3])


def read_matrix_market(path, handling="LeaveAsLower"):
    """read_matrix_market.rs:346-475 (real files) -> (nrow, ncol, sym, ai, aj, ax); raises ValueError(message)"""
    with open(path) as f:
        lines = f.read().split("\n")
    if not lines or (len(lines) == 1 and lines[0] == ""):
        raise ValueError("the file is empty")
    hdr = lines[0].split()
    if not hdr:
        raise ValueError("cannot find the keyword %%MatrixMarket on the first line")
    if hdr[0] != "%%MatrixMarket":
        raise ValueError("the header (first line) must start with %%MatrixMarket")
    if len(hdr) < 5 or hdr[1] != "matrix" or hdr[2] != "coordinate" or hdr[3] not in ("real", "complex"):
        raise ValueError("bad header")
    if hdr[3] == "complex":
        raise ValueError("complex")
    if hdr[4] not in ("general", "symmetric"):
        raise ValueError("bad header")
    symmetric = hdr[4] == "symmetric"
    k = 1
    while True:
        t = lines[k].split()
        k += 1
        if not t or t[0].startswith("%"):
            continue
        m, n, nnz = int(t[0]), int(t[1]), int(t[2])
        break
    if m < 1 or n < 1 or nnz < 1:
        raise ValueError("found invalid (zero or negative) dimensions")
    if symmetric and m != n:
        raise ValueError("MatrixMarket data is invalid: the number of rows must equal the number of columns for symmetric matrices")
    ai, aj, ax = [], [], []
    pos = 0
    for line in lines[k:]:
        t = line.split()
        if not t or t[0].startswith("%"):
            continue
        if pos == nnz:
            raise ValueError("there are more values than specified")
        i, j, a = int(t[0]) - 1, int(t[1]) - 1, float(t[2])
        if i < 0 or i >= m or j < 0 or j >= n:
            raise ValueError("found an invalid index")
        pos += 1
        if symmetric and handling == "SwapToUpper":
            ai.append(j), aj.append(i), ax.append(a)
        else:
            ai.append(i), aj.append(j), ax.append(a)
            if symmetric and handling == "MakeItFull" and i != j:
                ai.append(j), aj.append(i), ax.append(a)
    if pos != nnz:
        raise ValueError("not all values have been found")
    sym = "No"
    if symmetric:
        sym = {"LeaveAsLower": "YesLower", "SwapToUpper": "YesUpper", "MakeItFull": "YesFull"}[handling]
    return m, n, sym, np.array(ai, dtype=np.int32), np.array(aj, dtype=np.int32), np.array(ax)


# ---- independent CPU LU (SuperLU stand-in for UMFPACK; see module docstring) -----------------------------------
def full_scipy_matrix(nrow, ncol, ai, aj, ax, sym="No"):
    import scipy.sparse as sp

    a = sp.coo_matrix((ax, (ai, aj)), shape=(nrow, ncol)).tocsr()  # sums duplicates
    if sym in ("YesLower", "YesUpper"):
        d = sp.diags(a.diagonal())
        a = a + a.T - d
    return a.tocsc()


def lu_factorize(a_csc, permc_spec="MMD_AT_PLUS_A"):
    import scipy.sparse.linalg as spla

    return spla.splu(a_csc, permc_spec=permc_spec)


def lu_solve(a_csc, b, permc_spec="COLAMD"):
    """factorize + solve + one step of iterative refinement (UMFPACK's solve refines too, interface_umfpack.c:229)"""
    lu = lu_factorize(a_csc, permc_spec)
    x = lu.solve(b)
    r = b - a_csc @ x
    x = x + lu.solve(r)
    return x


# ---- the scalar walk of the front plan (mf_host.cpp) -----------------------------------------------------------
def mf_solve(n, rowptr, colidx, vals, b, sym_lower=False, ordering=0, matching=2, panel_width=0, nd_leaf=0, nrefine=2, verbose=0):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(n)
    st = np.zeros(10)
    rc = mf().oracle_mf_solve(n, _p(rowptr, _i32p), _p(colidx, _i32p), _p(vals, _f64p), 1 if sym_lower else 0, ordering, matching,
                              panel_width, nd_leaf, nrefine, 0.0, _p(b, _f64p), _p(x, _f64p), _p(st, _f64p), verbose)
    return rc, x, dict(nnodes=st[0], nlevels=st[1], nnz_l=st[2], nnz_u=st[3], flops=st[4], n_perturbed=st[5],
                       rel_residual=st[6], max_front=st[7], t_order=st[8], t_symbolic=st[9])


def mf_analyze(n, rowptr, colidx, vals, sym_lower=False, ordering=0, matching=0, panel_width=0, nd_leaf=0, verbose=0):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    st = np.zeros(10)
    rc = mf().oracle_mf_analyze(n, _p(rowptr, _i32p), _p(colidx, _i32p), _p(vals, _f64p), 1 if sym_lower else 0, ordering, matching,
                                panel_width, nd_leaf, _p(st, _f64p), verbose)
    return rc, dict(nnodes=st[0], nlevels=st[1], nnz_l=st[2], nnz_u=st[3], flops=st[4], cb_size=st[5], fac_size=st[6],
                    max_front=st[7], t_order=st[8], t_symbolic=st[9])


class MfHandle:
    """analyze + factorize on the host; exposes the factor arrays for array-by-array comparison with the GPU"""

    def __init__(self, n, rowptr, colidx, vals, sym_lower=False, ordering=0, matching=2, panel_width=0, nd_leaf=0):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)
        st = ctypes.c_int(0)
        self.h = mf().oracle_mf_create(n, _p(self.rowptr, _i32p), _p(self.colidx, _i32p), _p(self.vals, _f64p), 1 if sym_lower else 0,
                                       ordering, matching, panel_width, nd_leaf, 0.0, ctypes.byref(st))
        self.status = st.value
        if not self.h:
            raise ValueError("oracle_mf_create failed: %d" % st.value)
        sz = np.zeros(5, dtype=np.int64)
        mf().oracle_mf_sizes(self.h, _p(sz, _i64p))
        self.fac_size, self.dinv_size, self.n, self.nnodes, self.n_perturbed = (int(v) for v in sz)

    def factors(self):
        fac = np.zeros(self.fac_size)
        dinv = np.zeros(self.dinv_size)
        lperm = np.zeros(self.n, dtype=np.int32)
        mf().oracle_mf_get(self.h, _p(fac, _f64p), _p(dinv, _f64p), _p(lperm, _i32p))
        return fac, dinv, lperm

    def solve(self, b, nrefine=2):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(self.n)
        res = ctypes.c_double(0.0)
        mf().oracle_mf_handle_solve(self.h, _p(self.vals, _f64p), _p(b, _f64p), _p(x, _f64p), nrefine, ctypes.byref(res))
        return x, res.value

    def __del__(self):
        try:
            if self.h:
                mf().oracle_mf_free(self.h)
                self.h = None
        except Exception:
            pass
