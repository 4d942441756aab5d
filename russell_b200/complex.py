"""Complex64 twin of the host-side mirror.

    ComplexCooMatrix / ComplexCsrMatrix / ComplexCscMatrix <- NumCooMatrix<Complex64> etc. (russell_sparse/src/aliases.rs,
                                                               coo_matrix.rs:21-73, csr_matrix.rs:332-480)
    ComplexSolverB200 (ComplexLinSolTrait)                  <- russell_sparse/src/complex_solver_cudss.rs:101-420 (cloned shape)
    ComplexLinSolver                                        <- russell_sparse/src/complex_lin_solver.rs:66-185
    verify_from_complex                                     <- russell_sparse/src/verify_lin_sys.rs:104-146

Vectors are numpy complex128 arrays (the memory layout of russell_lab::ComplexVector: interleaved re/im f64 pairs).
All arithmetic of factorize/solve/A·x runs in the CUDA kernels behind complex_solver_b200_* (include/solver_b200.h).
"""
import ctypes
import time

import numpy as np

from . import _lib
from ._lib import p_f64, p_i32, ptr
from . import (CooMatrix, LinSolParams, StrError, Sym, Genie, VerifyLinSys, _to_i32, b200_matching, b200_ordering,
               b200_pivoting, handle_b200_error_code, SolverB200)

__all__ = ["ComplexCooMatrix", "ComplexCsrMatrix", "ComplexCscMatrix", "ComplexSolverB200", "ComplexLinSolver",
           "verify_from_complex"]


class ComplexCooMatrix(CooMatrix):
    """COO triplets with Complex64 values; same guards and messages as CooMatrix (coo_matrix.rs:173-198,324-352)"""

    def __init__(self, nrow, ncol, max_nnz, symmetric=Sym.No):
        super().__init__(nrow, ncol, max_nnz, symmetric)
        self.values = np.zeros(max_nnz, dtype=np.complex128)

    @staticmethod
    def from_triplets(nrow, ncol, ii, jj, vv, symmetric=Sym.No):
        vv = np.asarray(vv, dtype=np.complex128)
        real = CooMatrix.from_triplets(nrow, ncol, ii, jj, np.zeros(len(vv)), symmetric)  # index validation
        coo = ComplexCooMatrix(nrow, ncol, max(1, len(vv)), symmetric)
        coo.indices_i[:], coo.indices_j[:] = real.indices_i, real.indices_j
        coo.values[: len(vv)] = vv
        coo.nnz = len(vv)
        return coo

    def as_dense(self):
        a = np.zeros((self.nrow, self.ncol), dtype=np.complex128)
        for p in range(self.nnz):
            i, j = self.indices_i[p], self.indices_j[p]
            a[i, j] += self.values[p]
            if self.symmetric.triangular() and i != j:
                a[j, i] += self.values[p]
        return a


class _ComplexCompressed:
    _fn = None

    def __init__(self, coo):
        if coo.nnz < 1:
            raise StrError(self._empty_msg)
        self.symmetric = coo.symmetric
        self.nrow, self.ncol = coo.nrow, coo.ncol
        nmajor = coo.nrow if self._fn == "b200_complex_coo_to_csr" else coo.ncol
        self.pointers = np.zeros(nmajor + 1, dtype=np.int32)
        self.indices = np.zeros(coo.nnz, dtype=np.int32)
        self.values = np.zeros(coo.nnz, dtype=np.complex128)
        self.update_from_coo(coo)

    def update_from_coo(self, coo):  # csr_matrix.rs:359-480 over Complex64
        kind = "csr" if self._fn == "b200_complex_coo_to_csr" else "csc"
        if coo.symmetric != self.symmetric:
            raise StrError("coo.symmetric must be equal to %s.symmetric" % kind)
        if coo.nrow != self.nrow:
            raise StrError("coo.nrow must be equal to %s.nrow" % kind)
        if coo.ncol != self.ncol:
            raise StrError("coo.ncol must be equal to %s.ncol" % kind)
        if coo.nnz != len(self.values):
            raise StrError("coo.nnz must be equal to nnz(dup)")
        lib = _lib.load()
        ci = np.ascontiguousarray(coo.indices_i[: coo.nnz])
        cj = np.ascontiguousarray(coo.indices_j[: coo.nnz])
        cv = np.ascontiguousarray(coo.values[: coo.nnz], dtype=np.complex128)
        rc = getattr(lib, self._fn)(coo.nrow, coo.ncol, coo.nnz, ptr(ci, p_i32), ptr(cj, p_i32), ptr(cv, p_f64),
                                    ptr(self.pointers, p_i32), ptr(self.indices, p_i32), ptr(self.values, p_f64))
        if rc != 0:
            raise StrError("COO conversion failed (code %d)" % rc)

    @property
    def nnz(self):
        return int(self.pointers[-1])


class ComplexCsrMatrix(_ComplexCompressed):
    _fn = "b200_complex_coo_to_csr"
    _empty_msg = "COO to CSR requires nnz > 0"

    @staticmethod
    def from_coo(coo):
        return ComplexCsrMatrix(coo)

    row_pointers = property(lambda self: self.pointers)
    col_indices = property(lambda self: self.indices)


class ComplexCscMatrix(_ComplexCompressed):
    _fn = "b200_complex_coo_to_csc"
    _empty_msg = "COO to CSC requires nnz > 0"

    @staticmethod
    def from_coo(coo):
        return ComplexCscMatrix(coo)

    col_pointers = property(lambda self: self.pointers)
    row_indices = property(lambda self: self.indices)


class ComplexSolverB200:
    """Clone of ComplexSolverCUDSS's state machine (complex_solver_cudss.rs:101-420) over complex_solver_b200_*."""

    STAT_NAMES = SolverB200.STAT_NAMES

    def __init__(self, coo_boundary=True):
        """coo_boundary=True: triplet structure analysed once (complex_solver_b200_initialize_coo), every later factorize
        ships the raw triplet values and the duplicates are summed on the device; False: the reference's data flow, host
        ComplexCsrMatrix::update_from_coo on every call (complex_solver_cudss.rs:221)."""
        self._lib = _lib.load()
        self.solver = self._lib.complex_solver_b200_new()
        if not self.solver:
            raise StrError("c-code failed to allocate the B200 solver")
        self.coo_boundary = bool(coo_boundary)
        self.csr = None
        self.initialized = False
        self.factorized = False
        self.initialized_sym = Sym.No
        self.initialized_ndim = 0
        self.initialized_nnz = 0
        self.effective_matching = 0
        self.effective_pivoting = 0
        self.time_initialize_ns = 0
        self.time_factorize_ns = 0
        self.time_solve_ns = 0

    def __del__(self):  # Drop (complex_solver_cudss.rs:146-153)
        try:
            if getattr(self, "solver", None):
                self._lib.complex_solver_b200_drop(self.solver)
                self.solver = None
        except Exception:
            pass

    def set_option(self, key, value):
        rc = self._lib.complex_solver_b200_set_option(self.solver, key.encode(), float(value))
        if rc != 0:
            raise StrError(handle_b200_error_code(rc))

    def factorize(self, mat, params=None):  # complex_solver_cudss.rs:206-326
        if self.initialized:
            if mat.symmetric != self.initialized_sym:
                raise StrError("subsequent factorizations must use the same matrix (symmetric differs)")
            if mat.nrow != self.initialized_ndim:
                raise StrError("subsequent factorizations must use the same matrix (ndim differs)")
            if mat.nnz != self.initialized_nnz:
                raise StrError("subsequent factorizations must use the same matrix (nnz differs)")
            if params is not None:
                raise StrError("subsequent factorizations must not change LinSolParams")
            if not self.coo_boundary:
                self.csr.update_from_coo(mat)
        else:
            if mat.nrow != mat.ncol:
                raise StrError("the matrix must be square")
            if mat.nnz < 1:
                raise StrError("the COO matrix must have at least one non-zero value")
            if mat.symmetric == Sym.YesUpper:
                raise StrError("B200 requires Sym::YesLower or Sym::YesFull for symmetric matrices")
            self.initialized_sym = mat.symmetric
            self.initialized_ndim = mat.nrow
            self.initialized_nnz = mat.nnz
            if not self.coo_boundary:
                self.csr = ComplexCsrMatrix.from_coo(mat)
        csr = self.csr
        if self.coo_boundary:
            coo_i = np.ascontiguousarray(mat.indices_i[: mat.nnz], dtype=np.int32)
            coo_j = np.ascontiguousarray(mat.indices_j[: mat.nnz], dtype=np.int32)
            coo_v = np.ascontiguousarray(mat.values[: mat.nnz], dtype=np.complex128)
        par = params if params is not None else LinSolParams()
        pivot_epsilon = par.pivot_epsilon if par.pivot_epsilon is not None else -1.0
        refinement_nstep = par.refinement_nstep if par.refinement_nstep is not None else -1
        if par.hybrid_memory_factor is not None:
            v = par.hybrid_memory_factor
            if v < 0.01 or v > 0.99:
                raise StrError("the hybrid memory factor must be in [0.01, 0.99]")
            hybrid = v
        else:
            hybrid = -1.0
        verbose = 1 if par.verbose else 0
        general_symmetric = 1 if mat.symmetric == Sym.YesLower else 0
        positive_definite = 1 if (par.positive_definite and mat.symmetric == Sym.YesLower) else 0
        if not self.initialized:
            t0 = time.perf_counter_ns()
            if self.coo_boundary:
                status = self._lib.complex_solver_b200_initialize_coo(
                    self.solver, b200_ordering(par.ordering), b200_matching(par.matching), b200_pivoting(par.pivoting),
                    pivot_epsilon, refinement_nstep, hybrid, verbose, general_symmetric, positive_definite,
                    _to_i32(mat.nrow), _to_i32(mat.nnz), ptr(coo_i, p_i32), ptr(coo_j, p_i32), ptr(coo_v, p_f64))
            else:
                status = self._lib.complex_solver_b200_initialize(
                    self.solver, b200_ordering(par.ordering), b200_matching(par.matching), b200_pivoting(par.pivoting),
                    pivot_epsilon, refinement_nstep, hybrid, verbose, general_symmetric, positive_definite,
                    _to_i32(csr.nrow), ptr(csr.pointers, p_i32), ptr(csr.indices, p_i32), ptr(csr.values, p_f64))
            if status != 0:
                raise StrError(handle_b200_error_code(status))
            self.time_initialize_ns = time.perf_counter_ns() - t0
            self.initialized = True
        em, ep = _lib.c_i32(0), _lib.c_i32(0)
        t0 = time.perf_counter_ns()
        if self.coo_boundary:
            status = self._lib.complex_solver_b200_factorize_coo_checked(self.solver, ctypes.byref(em), ctypes.byref(ep), verbose,
                                                                         _to_i32(mat.nnz), ptr(coo_i, p_i32), ptr(coo_j, p_i32),
                                                                         ptr(coo_v, p_f64))
        else:
            status = self._lib.complex_solver_b200_factorize(self.solver, ctypes.byref(em), ctypes.byref(ep), verbose,
                                                             ptr(csr.values, p_f64))
        if status != 0:
            raise StrError(handle_b200_error_code(status))
        self.time_factorize_ns = time.perf_counter_ns() - t0
        self.effective_matching, self.effective_pivoting = em.value, ep.value
        self.factorized = True

    def solve(self, x, rhs, verbose=False):  # complex_solver_cudss.rs:345-374
        if not self.factorized:
            raise StrError("the function factorize must be called before solve")
        if len(x) != self.initialized_ndim:
            raise StrError("the dimension of the vector of unknown values x is incorrect")
        if len(rhs) != self.initialized_ndim:
            raise StrError("the dimension of the right-hand side vector is incorrect")
        assert x.dtype == np.complex128 and x.flags.c_contiguous
        rhs = np.ascontiguousarray(rhs, dtype=np.complex128)
        t0 = time.perf_counter_ns()
        status = self._lib.complex_solver_b200_solve(self.solver, ptr(x, p_f64), ptr(rhs, p_f64), 1 if verbose else 0)
        if status != 0:
            raise StrError(handle_b200_error_code(status))
        self.time_solve_ns = time.perf_counter_ns() - t0

    # --- extensions -------------------------------------------------------------------------------------
    def device_stats(self):
        out = np.zeros(len(self.STAT_NAMES))
        rc = self._lib.complex_solver_b200_get_stats(self.solver, ptr(out, p_f64), len(out))
        if rc != 0:
            raise StrError(handle_b200_error_code(rc))
        return dict(zip(self.STAT_NAMES, out.tolist()))

    def residual(self, x, rhs):
        """||rhs - A x||_2 / ||rhs||_2 evaluated by the CUDA SpMV kernel"""
        out = ctypes.c_double(0.0)
        x = np.ascontiguousarray(x, dtype=np.complex128)
        rhs = np.ascontiguousarray(rhs, dtype=np.complex128)
        rc = self._lib.complex_solver_b200_residual(self.solver, ptr(x, p_f64), ptr(rhs, p_f64), ctypes.byref(out))
        if rc != 0:
            raise StrError(handle_b200_error_code(rc))
        return out.value

    def mat_vec_mul(self, x):
        """A x on the device (CSR SpMV kernel over the embedded matrix)"""
        x = np.ascontiguousarray(x, dtype=np.complex128)
        y = np.zeros_like(x)
        rc = self._lib.complex_solver_b200_spmv(self.solver, ptr(y, p_f64), ptr(x, p_f64))
        if rc != 0:
            raise StrError(handle_b200_error_code(rc))
        return y

    def update_stats(self, stats):  # complex_solver_cudss.rs:376-405
        stats.solver = "B200"
        stats.initialize_array.append(self.time_initialize_ns)
        stats.factorize_array.append(self.time_factorize_ns)
        stats.solve_array.append(self.time_solve_ns)
        names = {0: "None", 5: "MaxDiagProduct", 6: "Auto"}
        stats.effective_matching = names.get(self.effective_matching, "Unknown")
        stats.effective_pivoting = {5: "LocalBlock"}.get(self.effective_pivoting, "Unknown")
        if self.initialized:
            stats.device = self.device_stats()

    def get_ns_init(self):
        return self.time_initialize_ns

    def get_ns_fact(self):
        return self.time_factorize_ns

    def get_ns_solve(self):
        return self.time_solve_ns


class ComplexLinSolver:
    """complex_lin_solver.rs:66-185"""

    def __init__(self, genie):
        if genie == Genie.B200:
            self.actual = ComplexSolverB200()
        elif genie == Genie.Cudss:
            raise StrError("cuDSS solver is not available")
        elif genie == Genie.Mumps:
            raise StrError("MUMPS solver is not available")
        else:
            raise StrError("UMFPACK solver is not available")

    @staticmethod
    def compute(genie, x, mat, rhs, params=None):  # complex_lin_solver.rs:170-182
        solver = ComplexLinSolver(genie)
        solver.actual.factorize(mat, params)
        solver.actual.solve(x, rhs, False)
        return solver


def verify_from_complex(mat, x, rhs, solver):
    """VerifyLinSys::from_complex (verify_lin_sys.rs:104-146); A·x goes through the CUDA SpMV of `solver`"""
    nrow, ncol, _, _ = mat.get_info()
    if len(x) != ncol:
        raise StrError("x.dim() must be equal to ncol")
    if len(rhs) != nrow:
        raise StrError("rhs.dim() must be equal to nrow")
    values = mat.get_values()
    if len(values) < 1:
        raise StrError("matrix is empty")
    max_abs_a = float(np.max(np.abs(values)))
    ax = solver.mat_vec_mul(x)
    max_abs_ax = float(np.max(np.abs(ax)))
    max_abs_diff = float(np.max(np.abs(ax - rhs)))
    return VerifyLinSys(max_abs_a, max_abs_ax, max_abs_diff, max_abs_diff / (max_abs_a + 1.0))
