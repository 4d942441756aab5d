"""russell_b200 -- host-side mirror of russell_sparse's solver interface for the B200 backend.

The reference is Rust; this container has no rustc, so the host logic above the C ABI lives in
libsolver_b200.so (C++: analysis, COO->CSR, Matrix Market) and this thin Python layer only reproduces the
*interface* (names, argument meaning, error strings, state machine) so that the parity tests read like the
reference's own tests:

    CooMatrix / CsrMatrix / CscMatrix   <- russell_sparse/src/{coo,csr,csc}_matrix.rs
    Sym, Genie, MMsym, Ordering, ...     <- russell_sparse/src/enums.rs
    LinSolParams                         <- russell_sparse/src/lin_sol_params.rs:5-110
    SolverB200 (LinSolTrait)             <- russell_sparse/src/solver_cudss.rs:92-390 (cloned shape)
    LinSolver                            <- russell_sparse/src/lin_solver.rs:105-224
    VerifyLinSys                         <- russell_sparse/src/verify_lin_sys.rs:60-96
    read_matrix_market                   <- russell_sparse/src/read_matrix_market.rs:346-475

Nothing here computes on the CPU what the GPU is supposed to compute: factorize/solve/SpMV/residual all go
through the CUDA kernels; the product fails loudly when the extension or the device is missing.
"""
import enum
import time

import numpy as np

from . import _lib
from ._lib import p_f64, p_i32, p_i64, ptr

__all__ = [
    "StrError", "Sym", "Genie", "MMsym", "Ordering", "Scaling", "Matching", "Pivoting", "LinSolParams",
    "CooMatrix", "CsrMatrix", "CscMatrix", "SolverB200", "LinSolver", "VerifyLinSys", "StatsLinSol",
    "read_matrix_market", "read_matrix_market_pair", "handle_b200_error_code",
]


class StrError(Exception):
    """Mirrors russell's `StrError = &'static str` error channel."""

    def __init__(self, msg):
        super().__init__(msg)
        self.msg = msg


class Sym(enum.Enum):  # enums.rs:27-39
    No = 0
    YesFull = 1
    YesLower = 2
    YesUpper = 3

    def triangular(self):
        return self in (Sym.YesLower, Sym.YesUpper)


class MMsym(enum.Enum):  # enums.rs:45-67
    LeaveAsLower = 0
    SwapToUpper = 1
    MakeItFull = 2


class Genie(enum.Enum):  # enums.rs:5-20 plus the new arm
    Cudss = "cudss"
    Mumps = "mumps"
    Umfpack = "umfpack"
    B200 = "b200"

    @staticmethod
    def from_str(name):
        name = name.lower()
        for g in Genie:
            if g.value == name:
                return g
        return Genie.Umfpack  # reference default (enums.rs:338-345)

    def to_string(self):
        return self.value

    def get_sym(self, symmetric):
        if not symmetric:
            return Sym.No
        return Sym.YesFull if self == Genie.Umfpack else Sym.YesLower


Ordering = enum.Enum("Ordering", "Amd Amf Auto Best BtfColamd Cholmod Colamd Metis No Pord Qamd Scotch")
Scaling = enum.Enum("Scaling", "Auto Column Diagonal Max No RowCol RowColIter RowColRig Sum")
Matching = enum.Enum("Matching", "None_ Auto MaxDiagCount MaxMinDiag MaxMinDiagAlt MaxDiagSum MaxDiagProduct")
Pivoting = enum.Enum("Pivoting", "Auto None_ GlobalCol GlobalRow Diagonal LocalBlock")

# integer codes sent over the C ABI: identical to the cuDSS maps (solver_cudss.rs:393-466)
_ORDERING_CODE = {"Amd": 3, "BtfColamd": 1, "Colamd": 2, "Metis": 4, "No": 5}
_MATCHING_CODE = {"None_": 0, "MaxDiagCount": 1, "MaxMinDiag": 2, "MaxMinDiagAlt": 3, "MaxDiagSum": 4,
                  "MaxDiagProduct": 5, "Auto": 6}
_PIVOTING_CODE = {"Auto": 0, "None_": 1, "GlobalCol": 2, "GlobalRow": 3, "Diagonal": 4, "LocalBlock": 5}


def b200_ordering(o):
    return _ORDERING_CODE.get(o.name, 0)


def b200_matching(m):
    return _MATCHING_CODE[m.name]


def b200_pivoting(p):
    return _PIVOTING_CODE[p.name]


class LinSolParams:
    """lin_sol_params.rs:5-110 (fields that do not apply to this backend are kept for interface parity)"""

    def __init__(self):
        self.ordering = Ordering.Auto
        self.scaling = Scaling.Auto
        self.matching = Matching.None_
        self.pivoting = Pivoting.Auto
        self.pivot_epsilon = None
        self.refinement_nstep = None
        self.hybrid_memory_factor = None
        self.positive_definite = False
        self.compute_determinant = False
        self.compute_error_estimates = False
        self.compute_condition_numbers = False
        self.verbose = False


def _to_i32(x):
    if x > 2147483647:
        raise OverflowError("index does not fit in i32")  # constants.rs:20-22 panics
    return int(x)


class CooMatrix:
    """COO triplets with duplicates allowed (coo_matrix.rs:21-73)."""

    def __init__(self, nrow, ncol, max_nnz, symmetric=Sym.No):  # coo_matrix.rs:173-198
        if nrow < 1:
            raise StrError("nrow must be ≥ 1")
        if ncol < 1:
            raise StrError("ncol must be ≥ 1")
        if max_nnz < 1:
            raise StrError("max_nnz must be ≥ 1")
        if symmetric != Sym.No and nrow != ncol:
            raise StrError("symmetric storage requires a square matrix")
        self.symmetric = symmetric
        self.nrow, self.ncol = nrow, ncol
        self.nnz, self.max_nnz = 0, max_nnz
        self.indices_i = np.zeros(max_nnz, dtype=np.int32)
        self.indices_j = np.zeros(max_nnz, dtype=np.int32)
        self.values = np.zeros(max_nnz, dtype=np.float64)

    @staticmethod
    def from_triplets(nrow, ncol, ii, jj, vv, symmetric=Sym.No):
        """bulk constructor (coo_matrix.rs:200-290 `from`): validates like `put` but vectorised"""
        ii = np.asarray(ii, dtype=np.int32)
        jj = np.asarray(jj, dtype=np.int32)
        vv = np.asarray(vv, dtype=np.float64)
        coo = CooMatrix(nrow, ncol, max(1, len(vv)), symmetric)
        if len(vv):
            if ii.min() < 0 or ii.max() >= nrow:
                raise StrError("COO matrix: index of row is outside range")
            if jj.min() < 0 or jj.max() >= ncol:
                raise StrError("COO matrix: index of column is outside range")
            if symmetric == Sym.YesLower and np.any(jj > ii):
                raise StrError("COO matrix: j > i is incorrect for lower triangular storage")
            if symmetric == Sym.YesUpper and np.any(jj < ii):
                raise StrError("COO matrix: j < i is incorrect for upper triangular storage")
        coo.indices_i[:], coo.indices_j[:], coo.values[:] = ii, jj, vv
        coo.nnz = len(vv)
        return coo

    def put(self, i, j, aij):  # coo_matrix.rs:324-352
        if i >= self.nrow or i < 0:
            raise StrError("COO matrix: index of row is outside range")
        if j >= self.ncol or j < 0:
            raise StrError("COO matrix: index of column is outside range")
        if self.nnz >= self.max_nnz:
            raise StrError("COO matrix: max number of items has been reached")
        if self.symmetric == Sym.YesLower and j > i:
            raise StrError("COO matrix: j > i is incorrect for lower triangular storage")
        if self.symmetric == Sym.YesUpper and j < i:
            raise StrError("COO matrix: j < i is incorrect for upper triangular storage")
        self.indices_i[self.nnz] = _to_i32(i)
        self.indices_j[self.nnz] = _to_i32(j)
        self.values[self.nnz] = aij
        self.nnz += 1

    def reset(self):  # coo_matrix.rs:388-390
        self.nnz = 0

    def get_info(self):
        return self.nrow, self.ncol, self.nnz, self.symmetric

    def get_values(self):
        return self.values[: self.nnz]

    def as_dense(self):
        a = np.zeros((self.nrow, self.ncol))
        for p in range(self.nnz):
            i, j = self.indices_i[p], self.indices_j[p]
            a[i, j] += self.values[p]
            if self.symmetric.triangular() and i != j:
                a[j, i] += self.values[p]
        return a


class _Compressed:
    """shared part of CsrMatrix / CscMatrix: conversion through the C++ host converter"""

    _fn = None

    def __init__(self, coo):
        if coo.nnz < 1:
            raise StrError(self._empty_msg)
        self.symmetric = coo.symmetric
        self.nrow, self.ncol = coo.nrow, coo.ncol
        nmajor = coo.nrow if self._fn == "b200_coo_to_csr" else coo.ncol
        self.pointers = np.zeros(nmajor + 1, dtype=np.int32)
        self.indices = np.zeros(coo.nnz, dtype=np.int32)
        self.values = np.zeros(coo.nnz, dtype=np.float64)
        self.update_from_coo(coo)

    def update_from_coo(self, coo):  # csr_matrix.rs:359-480 / csc_matrix.rs:365-505
        kind = "csr" if self._fn == "b200_coo_to_csr" else "csc"
        if coo.symmetric != self.symmetric:
            raise StrError("coo.symmetric must be equal to %s.symmetric" % kind)
        if coo.nrow != self.nrow:
            raise StrError("coo.nrow must be equal to %s.nrow" % kind)
        if coo.ncol != self.ncol:
            raise StrError("coo.ncol must be equal to %s.ncol" % kind)
        if coo.nnz != len(self.values):
            raise StrError("coo.nnz must be equal to nnz(dup)")
        lib = _lib.load()
        rc = getattr(lib, self._fn)(coo.nrow, coo.ncol, coo.nnz, ptr(coo.indices_i, p_i32), ptr(coo.indices_j, p_i32),
                                    ptr(coo.values, p_f64), ptr(self.pointers, p_i32), ptr(self.indices, p_i32),
                                    ptr(self.values, p_f64))
        if rc != 0:
            raise StrError("COO conversion failed (code %d)" % rc)

    @property
    def nnz(self):
        return int(self.pointers[-1])


class CsrMatrix(_Compressed):
    _fn = "b200_coo_to_csr"
    _empty_msg = "COO to CSR requires nnz > 0"

    @staticmethod
    def from_coo(coo):  # csr_matrix.rs:332-350
        return CsrMatrix(coo)

    row_pointers = property(lambda self: self.pointers)
    col_indices = property(lambda self: self.indices)


class CscMatrix(_Compressed):
    _fn = "b200_coo_to_csc"
    _empty_msg = "COO to CSC requires nnz > 0"

    @staticmethod
    def from_coo(coo):  # csc_matrix.rs:337-356
        return CscMatrix(coo)

    col_pointers = property(lambda self: self.pointers)
    row_indices = property(lambda self: self.indices)


_MM_ERRORS = {  # read_matrix_market.rs messages, keyed by the codes of host_formats.cpp
    1: "cannot open file",
    2: "the file is empty",
    3: "the header (first line) must start with %%MatrixMarket",
    4: "cannot find the keyword %%MatrixMarket on the first line",
    5: "after %%MatrixMarket, the first option must be \"matrix\"",
    6: "cannot find the first option in the header line",
    7: "after %%MatrixMarket, the second option must be \"coordinate\"",
    8: "cannot find the second option in the header line",
    9: "after %%MatrixMarket, the third option must be \"real\" or \"complex\"",
    10: "cannot find the third option in the header line",
    11: "after %%MatrixMarket, the fourth option must be \"general\", \"symmetric\", or \"Hermitian\"",
    12: "cannot find the fourth option in the header line",
    13: "\"Hermitian\" keyword can only be used with the \"complex\" type",
    14: "cannot parse number of rows",
    15: "cannot read number of columns",
    16: "cannot parse number of columns",
    17: "cannot read number of non-zeros",
    18: "cannot parse number of non-zeros",
    19: "found invalid (zero or negative) dimensions",
    20: "there are more values than specified",
    21: "cannot parse i",
    22: "cannot read j",
    23: "cannot parse j",
    24: "cannot read aij",
    25: "cannot parse aij",
    26: "found an invalid index",
    27: "not all values have been found",
    28: "MatrixMarket data is invalid: the number of rows must equal the number of columns for symmetric matrices",
    29: "complex MatrixMarket files are not supported by the B200 backend yet",
    30: "cannot find the dimensions line",
    31: "cannot read bij",
    32: "cannot parse bij",
}


def read_matrix_market(full_path, symmetric_handling=MMsym.LeaveAsLower):
    """read_matrix_market.rs:346-475: coordinate files, real or complex, general / symmetric (/ Hermitian read as general).
    Returns a CooMatrix for a real file and a ComplexCooMatrix for a complex one (read_matrix_market_pair gives the
    reference's `(Option<CooMatrix>, Option<ComplexCooMatrix>)` shape)."""
    lib = _lib.load()
    info = np.zeros(8, dtype=np.int64)
    path = str(full_path).encode()
    rc = lib.b200_mm_read(path, symmetric_handling.value, ptr(info, p_i64), None, None, None, 0)
    if rc != 0:
        raise StrError(_MM_ERRORS.get(rc, "MatrixMarket error %d" % rc))
    m, n, _, is_sym, cap = (int(v) for v in info[:5])
    is_complex = bool(info[6])
    if is_sym:
        sym = {MMsym.LeaveAsLower: Sym.YesLower, MMsym.SwapToUpper: Sym.YesUpper, MMsym.MakeItFull: Sym.YesFull}[symmetric_handling]
    else:
        sym = Sym.No
    if is_complex:
        from .complex import ComplexCooMatrix

        coo = ComplexCooMatrix(m, n, cap, sym)
    else:
        coo = CooMatrix(m, n, cap, sym)
    rc = lib.b200_mm_read(path, symmetric_handling.value, ptr(info, p_i64), ptr(coo.indices_i, p_i32),
                          ptr(coo.indices_j, p_i32), ptr(coo.values, p_f64), cap)
    if rc != 0:
        raise StrError(_MM_ERRORS.get(rc, "MatrixMarket error %d" % rc))
    coo.nnz = int(info[5])
    return coo


def read_matrix_market_pair(full_path, symmetric_handling=MMsym.LeaveAsLower):
    """the reference's return shape: (coo_real or None, coo_complex or None)"""
    coo = read_matrix_market(full_path, symmetric_handling)
    return (None, coo) if coo.values.dtype == np.complex128 else (coo, None)


_B200_ERRORS = {  # same spirit as handle_cudss_error_code (solver_cudss.rs:501-558); OOM wording kept for
    # stats_lin_sol.rs:334-340 `is_memory_error` (substring "cudaMalloc" / "MALLOC")
    1: "Error(1): Matrix is singular",
    100: "cudaMalloc failed in the C code (B200)",
    200: "cudaMemcpy failed in the C code (B200)",
    300: "cudaStreamSynchronize failed in the C code (B200)",
    701: "B200 analysis failed: matrix is structurally singular",
    702: "B200 analysis failed: invalid CSR structure",
    703: "B200 analysis failed: invalid COO structure (index out of range or empty)",
    704: "B200 analysis failed: Sym::YesLower requires triplets with j <= i",
    705: "subsequent factorizations must use the same matrix (the COO structure differs from the analysed one)",
    801: "B200 numeric factorization failed: kernel launch failure",
    802: "B200 numeric factorization failed: matrix values are not finite",
    901: "B200 solve failed: kernel launch failure",
    907: "B200 solve failed: iterative refinement failed",
    100000: "B200 failed due to NULL POINTER error",
    200000: "B200 failed due to MALLOC error",
    300000: "B200 failed due to VERSION error",
    400000: "B200 solver is not AVAILABLE (no CUDA device)",
    500000: "B200 failed because INITIALIZATION is needed",
    600000: "B200 failed because FACTORIZATION is needed",
    700000: "B200 failed because INITIALIZATION has been completed already",
}


def handle_b200_error_code(err):
    return _B200_ERRORS.get(err, "Error: unknown error returned by c-code (B200)")


class StatsLinSol:
    """the slice of stats_lin_sol.rs:105-115 this backend fills"""

    def __init__(self):
        self.solver = "Unknown"
        self.effective_matching = "Unknown"
        self.effective_pivoting = "Unknown"
        self.effective_ordering = "Unknown"  # stats_lin_sol.rs:50-60 (StatsLinSolOutput)
        self.effective_scaling = "Unknown"
        self.rcond_estimate = 0.0
        self.determinant = (0.0, 0.0)  # (mantissa, exponent), base 10
        self.n_perturbed = 0
        self.rel_residual = -1.0
        self.initialize_array, self.factorize_array, self.solve_array = [], [], []
        self.device = {}


class SolverB200:
    """Clone of SolverCUDSS's state machine (solver_cudss.rs:92-390) over solver_b200_*."""

    STAT_NAMES = ["nnodes", "nlevels", "nnz_l", "nnz_u", "flops", "fac_bytes", "cb_bytes", "max_front", "t_order_s",
                  "t_symbolic_s", "n_perturbed", "last_rel_residual", "last_refine_steps", "ms_factorize_device",
                  "ms_solve_device", "ms_sptrsv_device", "ms_spmv_device", "launches_factorize", "launches_solve",
                  "sptrsv_bytes", "spmv_bytes", "matched", "t_match_s", "last_backward_error", "effective_ordering",
                  "effective_scaling", "rcond", "t_initialize_host_s", "plan_cache_hit"]

    def __init__(self, coo_boundary=True):
        """coo_boundary=True: the triplet structure is analysed once (solver_b200_initialize_coo) and every later
        factorize ships only the raw triplet values, summed into CSR slots on the device (solver_b200_factorize_coo);
        False: the reference's data flow, host CsrMatrix::update_from_coo on every call (solver_cudss.rs:209)."""
        self._lib = _lib.load()
        self.solver = self._lib.solver_b200_new()
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

    def __del__(self):  # Drop (solver_cudss.rs:133-140)
        try:
            if getattr(self, "solver", None):
                self._lib.solver_b200_drop(self.solver)
                self.solver = None
        except Exception:
            pass

    def set_option(self, key, value):
        rc = self._lib.solver_b200_set_option(self.solver, key.encode(), float(value))
        if rc != 0:
            raise StrError(handle_b200_error_code(rc))

    def factorize(self, mat, params=None):  # solver_cudss.rs:194-311
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
                self.csr.update_from_coo(self._trim(mat))
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
                self.csr = CsrMatrix.from_coo(self._trim(mat))
        csr = self.csr
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
        import ctypes
        if self.coo_boundary:
            tm = self._trim(mat)
            coo_i = np.ascontiguousarray(tm.indices_i[: mat.nnz], dtype=np.int32)
            coo_j = np.ascontiguousarray(tm.indices_j[: mat.nnz], dtype=np.int32)
            coo_v = np.ascontiguousarray(tm.values[: mat.nnz], dtype=np.float64)
        if not self.initialized:
            t0 = time.perf_counter_ns()
            if self.coo_boundary:
                status = self._lib.solver_b200_initialize_coo(
                    self.solver, b200_ordering(par.ordering), b200_matching(par.matching), b200_pivoting(par.pivoting),
                    pivot_epsilon, refinement_nstep, hybrid, verbose, general_symmetric, positive_definite,
                    _to_i32(mat.nrow), _to_i32(mat.nnz), ptr(coo_i, p_i32), ptr(coo_j, p_i32), ptr(coo_v, p_f64))
            else:
                status = self._lib.solver_b200_initialize(
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
            # the triplet indices travel with the values, like CsrMatrix::update_from_coo sees them on every call
            # (solver_cudss.rs:209): a CooMatrix refilled in another triplet order must not land in the wrong slots
            status = self._lib.solver_b200_factorize_coo_checked(self.solver, ctypes.byref(em), ctypes.byref(ep), verbose,
                                                                 _to_i32(mat.nnz), ptr(coo_i, p_i32), ptr(coo_j, p_i32), ptr(coo_v, p_f64))
        else:
            status = self._lib.solver_b200_factorize(self.solver, ctypes.byref(em), ctypes.byref(ep), verbose, ptr(csr.values, p_f64))
        if status != 0:
            raise StrError(handle_b200_error_code(status))
        self.time_factorize_ns = time.perf_counter_ns() - t0
        self.effective_matching, self.effective_pivoting = em.value, ep.value
        self.factorized = True

    @staticmethod
    def _trim(mat):
        """view of the live prefix of the triplet arrays (the Rust code indexes 0..nnz)"""
        if mat.nnz == mat.max_nnz:
            return mat
        v = CooMatrix.__new__(CooMatrix)
        v.symmetric, v.nrow, v.ncol, v.nnz, v.max_nnz = mat.symmetric, mat.nrow, mat.ncol, mat.nnz, mat.nnz
        v.indices_i = np.ascontiguousarray(mat.indices_i[: mat.nnz])
        v.indices_j = np.ascontiguousarray(mat.indices_j[: mat.nnz])
        v.values = np.ascontiguousarray(mat.values[: mat.nnz])
        return v

    def solve(self, x, rhs, verbose=False):  # solver_cudss.rs:333-360
        if not self.factorized:
            raise StrError("the function factorize must be called before solve")
        if len(x) != self.initialized_ndim:
            raise StrError("the dimension of the vector of unknown values x is incorrect")
        if len(rhs) != self.initialized_ndim:
            raise StrError("the dimension of the right-hand side vector is incorrect")
        assert x.dtype == np.float64 and x.flags.c_contiguous
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        t0 = time.perf_counter_ns()
        status = self._lib.solver_b200_solve(self.solver, ptr(x, p_f64), ptr(rhs, p_f64), 1 if verbose else 0)
        if status != 0:
            raise StrError(handle_b200_error_code(status))
        self.time_solve_ns = time.perf_counter_ns() - t0

    # --- extensions -------------------------------------------------------------------------------------
    def device_stats(self):
        out = np.zeros(len(self.STAT_NAMES))
        rc = self._lib.solver_b200_get_stats(self.solver, ptr(out, p_f64), len(out))
        if rc != 0:
            raise StrError(handle_b200_error_code(rc))
        return dict(zip(self.STAT_NAMES, out.tolist()))

    def residual(self, x, rhs):
        """||rhs - A x||_2 / ||rhs||_2 evaluated by the CUDA SpMV kernel"""
        import ctypes
        out = ctypes.c_double(0.0)
        x = np.ascontiguousarray(x, dtype=np.float64)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        rc = self._lib.solver_b200_residual(self.solver, ptr(x, p_f64), ptr(rhs, p_f64), ctypes.byref(out))
        if rc != 0:
            raise StrError(handle_b200_error_code(rc))
        return out.value

    def mat_vec_mul(self, x):
        """A x on the device (CSR SpMV kernel)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        rc = self._lib.solver_b200_spmv(self.solver, ptr(y, p_f64), ptr(x, p_f64))
        if rc != 0:
            raise StrError(handle_b200_error_code(rc))
        return y

    def determinant(self):
        """(coefficient, exponent) with det = coefficient * 10^exponent (solver_umfpack.rs:141-152)"""
        import ctypes
        c, e = ctypes.c_double(0.0), ctypes.c_double(0.0)
        rc = self._lib.solver_b200_determinant(self.solver, ctypes.byref(c), ctypes.byref(e))
        if rc != 0:
            raise StrError(handle_b200_error_code(rc))
        return c.value, e.value

    def rcond(self):
        """UMFPACK's reciprocal condition number estimate min|U_kk| / max|U_kk| (interface_umfpack.c:179-184)"""
        import ctypes
        out = ctypes.c_double(0.0)
        rc = self._lib.solver_b200_rcond(self.solver, ctypes.byref(out))
        if rc != 0:
            raise StrError(handle_b200_error_code(rc))
        return out.value

    def update_stats(self, stats):  # solver_cudss.rs:362-390 + the UMFPACK output fields (solver_umfpack.rs:392-422)
        stats.solver = "B200"
        if self.factorized:
            stats.rcond_estimate = self.rcond()
            stats.determinant = self.determinant()
        stats.initialize_array.append(self.time_initialize_ns)
        stats.factorize_array.append(self.time_factorize_ns)
        stats.solve_array.append(self.time_solve_ns)
        names = {0: "None", 1: "MaxDiagCount", 2: "MaxMinDiag", 3: "MaxMinDiagAlt", 4: "MaxDiagSum", 5: "MaxDiagProduct", 6: "Auto"}
        stats.effective_matching = names.get(self.effective_matching, "Unknown")
        piv = {0: "Auto", 1: "None", 2: "GlobalCol", 3: "GlobalRow", 4: "Diagonal", 5: "LocalBlock"}
        stats.effective_pivoting = piv.get(self.effective_pivoting, "Unknown")
        if self.initialized:
            stats.device = self.device_stats()
            stats.effective_ordering = {0: "Metis", 4: "Metis", 3: "Amd", 5: "No"}.get(int(stats.device["effective_ordering"]), "Unknown")
            stats.effective_scaling = "Max" if stats.device["effective_scaling"] else "No"
            stats.n_perturbed = int(stats.device["n_perturbed"])
            stats.rel_residual = stats.device["last_rel_residual"]

    def get_ns_init(self):
        return self.time_initialize_ns

    def get_ns_fact(self):
        return self.time_factorize_ns

    def get_ns_solve(self):
        return self.time_solve_ns


class LinSolver:
    """lin_solver.rs:105-224: `solver.actual.factorize(...)`, `solver.actual.solve(...)`"""

    def __init__(self, genie):
        if genie == Genie.B200:
            self.actual = SolverB200()
        elif genie == Genie.Cudss:
            raise StrError("cuDSS solver is not available")
        elif genie == Genie.Mumps:
            raise StrError("MUMPS solver is not available")
        else:
            raise StrError("UMFPACK solver is not available")  # CPU backends are not part of this package

    @staticmethod
    def compute(genie, x, mat, rhs, params=None):  # lin_solver.rs:212-224
        solver = LinSolver(genie)
        solver.actual.factorize(mat, params)
        solver.actual.solve(x, rhs, False)
        return solver


class VerifyLinSys:
    """verify_lin_sys.rs:60-96 -- russell's own accuracy metric; A·x goes through the CUDA SpMV of `solver`."""

    def __init__(self, max_abs_a, max_abs_ax, max_abs_diff, relative_error):
        self.max_abs_a, self.max_abs_ax = max_abs_a, max_abs_ax
        self.max_abs_diff, self.relative_error = max_abs_diff, relative_error

    @staticmethod
    def from_(mat, x, rhs, solver):
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


from .complex import (ComplexCooMatrix, ComplexCscMatrix, ComplexCsrMatrix, ComplexLinSolver,  # noqa: E402
                      ComplexSolverB200, verify_from_complex)

__all__ += ["ComplexCooMatrix", "ComplexCsrMatrix", "ComplexCscMatrix", "ComplexSolverB200", "ComplexLinSolver",
            "verify_from_complex"]
