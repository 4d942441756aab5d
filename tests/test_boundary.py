"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/solver_b200.h
declares, the status codes match the reference's constants.h, and -- without a CUDA device -- the product fails
loudly instead of falling back to a CPU path."""
import ctypes
import os
import re

import pytest

import russell_b200 as rb
from russell_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_text():
    with open(os.path.join(ROOT, "include", "solver_b200.h")) as f:
        return f.read()


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = re.findall(r"\b((?:complex_solver_b200|solver_b200|b200)_\w+)\s*\(", header_text())
    names = sorted(set(n for n in names if not n.startswith("B200_")))
    assert {"solver_b200_new", "solver_b200_drop", "solver_b200_initialize", "solver_b200_factorize",
            "solver_b200_solve", "complex_solver_b200_new", "complex_solver_b200_drop", "complex_solver_b200_initialize",
            "complex_solver_b200_factorize", "complex_solver_b200_solve"} <= set(names)
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    # and the ctypes prototypes cover the header
    for n in names:
        assert n in _lib.SIGNATURES, "no ctypes prototype for " + n


def test_five_entry_points_have_the_cudss_shim_shape():
    # solver_cudss.rs:25-52: initialize takes 14 arguments, factorize 5, solve 4
    sig = _lib.SIGNATURES
    assert len(sig["solver_b200_initialize"][1]) == 14
    assert len(sig["solver_b200_factorize"][1]) == 5
    assert len(sig["solver_b200_solve"][1]) == 4
    assert sig["solver_b200_new"][1] == [] and len(sig["solver_b200_drop"][1]) == 1
    # complex_solver_cudss.rs:32-64: the complex twin has the same shapes
    assert len(sig["complex_solver_b200_initialize"][1]) == 14
    assert len(sig["complex_solver_b200_factorize"][1]) == 5
    assert len(sig["complex_solver_b200_solve"][1]) == 4


def test_status_codes_match_reference_constants():
    # russell_sparse/c_code/constants.h:5-36
    h = header_text()
    want = {"B200_SUCCESSFUL_EXIT": 0, "B200_ERROR_NULL_POINTER": 100000, "B200_ERROR_MALLOC": 200000,
            "B200_ERROR_VERSION": 300000, "B200_ERROR_NOT_AVAILABLE": 400000, "B200_ERROR_NEED_INITIALIZATION": 500000,
            "B200_ERROR_NEED_FACTORIZATION": 600000, "B200_ERROR_ALREADY_INITIALIZED": 700000, "B200_ERROR_CUDA_MALLOC": 100,
            "B200_ERROR_CUDA_MEMCPY": 200, "B200_ERROR_CUDA_SYNCHRONIZE": 300, "B200_ERROR_SINGULAR": 1}
    for k, v in want.items():
        m = re.search(r"#define\s+%s\s+(\d+)" % k, h)
        assert m and int(m.group(1)) == v, k


def test_null_handle_is_rejected_not_dereferenced():
    lib = _lib.load()
    assert lib.solver_b200_factorize(None, None, None, 0, None) == 100000
    assert lib.solver_b200_solve(None, None, None, 0) == 100000
    assert lib.solver_b200_initialize(None, 0, 0, 0, -1.0, -1, -1.0, 0, 0, 0, 1, None, None, None) == 100000
    assert lib.complex_solver_b200_factorize(None, None, None, 0, None) == 100000
    assert lib.complex_solver_b200_solve(None, None, None, 0) == 100000
    lib.complex_solver_b200_drop(None)
    lib.solver_b200_drop(None)  # NULL-safe like solver_cudss_drop (interface_cudss.cu:126-129)


def test_error_messages():
    assert rb.handle_b200_error_code(1) == "Error(1): Matrix is singular"  # solver_umfpack.rs:492
    assert "cudaMalloc" in rb.handle_b200_error_code(100)                  # stats_lin_sol.rs:334-340 OOM substring
    assert "MALLOC" in rb.handle_b200_error_code(200000)
    assert rb.handle_b200_error_code(123) == "Error: unknown error returned by c-code (B200)"
    for c in (100, 200, 300, 701, 702, 801, 802, 901, 907, 100000, 200000, 300000, 400000, 500000, 600000, 700000):
        assert rb.handle_b200_error_code(c) != rb.handle_b200_error_code(123)


def test_genie_and_enum_maps():
    assert rb.Genie.from_str("B200") == rb.Genie.B200 and rb.Genie.B200.to_string() == "b200"
    assert rb.Genie.from_str("unknown") == rb.Genie.Umfpack  # enums.rs:338-345 default
    assert rb.Genie.B200.get_sym(True) == rb.Sym.YesLower and rb.Genie.B200.get_sym(False) == rb.Sym.No
    assert rb.Genie.Umfpack.get_sym(True) == rb.Sym.YesFull
    # the integers crossing the ABI are the cuDSS ones (solver_cudss.rs:393-466)
    assert rb.b200_ordering(rb.Ordering.Amd) == 3 and rb.b200_ordering(rb.Ordering.Metis) == 4
    assert rb.b200_ordering(rb.Ordering.No) == 5 and rb.b200_ordering(rb.Ordering.Auto) == 0
    assert rb.b200_matching(rb.Matching.None_) == 0 and rb.b200_matching(rb.Matching.Auto) == 6
    assert rb.b200_matching(rb.Matching.MaxDiagProduct) == 5
    assert rb.b200_pivoting(rb.Pivoting.Auto) == 0 and rb.b200_pivoting(rb.Pivoting.LocalBlock) == 5
    par = rb.LinSolParams()  # lin_sol_params.rs:86-110 defaults
    assert par.ordering == rb.Ordering.Auto and par.matching == rb.Matching.None_ and par.pivot_epsilon is None


def test_other_genies_are_not_available_here():
    for g, msg in ((rb.Genie.Cudss, "cuDSS solver is not available"), (rb.Genie.Mumps, "MUMPS solver is not available")):
        with pytest.raises(rb.StrError, match=msg):  # lin_solver.rs:125,132
            rb.LinSolver(g)


def _no_device():
    lib = _lib.load()
    h = lib.solver_b200_new()
    if h:
        lib.solver_b200_drop(h)
        return False
    return True


@pytest.mark.skipif(not _no_device(), reason="a CUDA device is present")
def test_no_device_means_no_solver_not_a_cpu_fallback():
    assert _lib.load().solver_b200_new() is None
    with pytest.raises(rb.StrError, match="c-code failed to allocate the B200 solver"):
        rb.SolverB200()
    with pytest.raises(rb.StrError):
        rb.LinSolver(rb.Genie.B200)
    assert _lib.load().complex_solver_b200_new() is None
    with pytest.raises(rb.StrError, match="c-code failed to allocate the B200 solver"):
        rb.ComplexSolverB200()


def test_product_does_not_import_the_oracle():
    # the oracle is test infrastructure: nothing under russell_b200/ may reference it
    for top in ("russell_b200", "tools", "bindings", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for fn in files:
                if fn.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h", ".rs")):
                    with open(os.path.join(dirpath, fn)) as f:
                        txt = f.read()
                    assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("// ", ""), fn
                    assert "liboracle" not in txt, fn
