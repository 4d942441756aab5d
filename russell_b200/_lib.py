"""ctypes binding of libsolver_b200.so (the C ABI declared in include/solver_b200.h).

Plays the role of the `unsafe extern "C"` block of the reference's Rust wrapper
(russell_sparse/src/solver_cudss.rs:25-52).  There is no fallback: if the shared library is missing the
import fails loudly, and without a CUDA device `solver_b200_new` returns NULL.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsolver_b200.so")

c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_f64 = ctypes.c_double
p_i32 = ctypes.POINTER(ctypes.c_int32)
p_i64 = ctypes.POINTER(ctypes.c_int64)
p_f64 = ctypes.POINTER(ctypes.c_double)
p_void = ctypes.c_void_p

# every symbol include/solver_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "solver_b200_new": (p_void, []),
    "solver_b200_drop": (None, [p_void]),
    "solver_b200_initialize": (c_i32, [p_void, c_i32, c_i32, c_i32, c_f64, c_i32, c_f64, c_i32, c_i32, c_i32, c_i32,
                                       p_i32, p_i32, p_f64]),
    "solver_b200_factorize": (c_i32, [p_void, p_i32, p_i32, c_i32, p_f64]),
    "solver_b200_solve": (c_i32, [p_void, p_f64, p_f64, c_i32]),
    "solver_b200_factorize_device": (c_i32, [p_void, p_void]),
    "solver_b200_initialize_coo": (c_i32, [p_void, c_i32, c_i32, c_i32, c_f64, c_i32, c_f64, c_i32, c_i32, c_i32, c_i32,
                                           c_i32, p_i32, p_i32, p_f64]),
    "solver_b200_factorize_coo": (c_i32, [p_void, p_i32, p_i32, c_i32, p_f64]),
    "solver_b200_factorize_coo_device": (c_i32, [p_void, p_void]),
    "solver_b200_factorize_coo_checked": (c_i32, [p_void, p_i32, p_i32, c_i32, c_i32, p_i32, p_i32, p_f64]),
    "solver_b200_rcond": (c_i32, [p_void, p_f64]),
    "solver_b200_ozaki_gemm": (c_i32, [c_i32, c_i32, p_f64, p_f64, p_f64, p_f64]),
    "solver_b200_solve_device": (c_i32, [p_void, p_void, p_void]),
    "solver_b200_residual": (c_i32, [p_void, p_f64, p_f64, p_f64]),
    "solver_b200_spmv": (c_i32, [p_void, p_f64, p_f64]),
    "solver_b200_determinant": (c_i32, [p_void, p_f64, p_f64]),
    "solver_b200_get_stats": (c_i32, [p_void, p_f64, c_i32]),
    "solver_b200_set_option": (c_i32, [p_void, ctypes.c_char_p, c_f64]),
    "solver_b200_debug_copy_factors": (c_i32, [p_void, p_f64, c_i64, p_f64, c_i64, p_i32, c_i64]),
    "solver_b200_debug_trace": (c_i32, [p_void, ctypes.POINTER(ctypes.c_uint64), p_i32, c_i32]),
    "solver_b200_version": (ctypes.c_char_p, []),
    "solver_b200_get_stream": (p_void, [p_void]),
    "solver_b200_copy_h2d": (c_i32, [p_void, p_void, p_void, ctypes.c_int64]),
    "solver_b200_copy_d2h": (c_i32, [p_void, p_void, p_void, ctypes.c_int64]),
    "solver_b200_get_device": (c_i32, [p_void]),
    # Complex64 twin (russell_b200/csrc/complex_b200.cu)
    "complex_solver_b200_new": (p_void, []),
    "complex_solver_b200_drop": (None, [p_void]),
    "complex_solver_b200_initialize": (c_i32, [p_void, c_i32, c_i32, c_i32, c_f64, c_i32, c_f64, c_i32, c_i32, c_i32, c_i32,
                                               p_i32, p_i32, p_f64]),
    "complex_solver_b200_factorize": (c_i32, [p_void, p_i32, p_i32, c_i32, p_f64]),
    "complex_solver_b200_solve": (c_i32, [p_void, p_f64, p_f64, c_i32]),
    "complex_solver_b200_initialize_coo": (c_i32, [p_void, c_i32, c_i32, c_i32, c_f64, c_i32, c_f64, c_i32, c_i32, c_i32, c_i32,
                                                   c_i32, p_i32, p_i32, p_f64]),
    "complex_solver_b200_factorize_coo": (c_i32, [p_void, p_i32, p_i32, c_i32, p_f64]),
    "complex_solver_b200_factorize_coo_checked": (c_i32, [p_void, p_i32, p_i32, c_i32, c_i32, p_i32, p_i32, p_f64]),
    "complex_solver_b200_factorize_device": (c_i32, [p_void, p_void]),
    "complex_solver_b200_solve_device": (c_i32, [p_void, p_void, p_void]),
    "complex_solver_b200_spmv": (c_i32, [p_void, p_f64, p_f64]),
    "complex_solver_b200_residual": (c_i32, [p_void, p_f64, p_f64, p_f64]),
    "complex_solver_b200_get_stats": (c_i32, [p_void, p_f64, c_i32]),
    "complex_solver_b200_set_option": (c_i32, [p_void, ctypes.c_char_p, c_f64]),
    "complex_solver_b200_real_handle": (p_void, [p_void]),
    # host formats (russell_b200/csrc/host_formats.cpp)
    "b200_coo_to_csr": (c_i32, [c_i32, c_i32, c_i32, p_i32, p_i32, p_f64, p_i32, p_i32, p_f64]),
    "b200_coo_to_csc": (c_i32, [c_i32, c_i32, c_i32, p_i32, p_i32, p_f64, p_i32, p_i32, p_f64]),
    "b200_complex_coo_to_csr": (c_i32, [c_i32, c_i32, c_i32, p_i32, p_i32, p_f64, p_i32, p_i32, p_f64]),
    "b200_complex_coo_to_csc": (c_i32, [c_i32, c_i32, c_i32, p_i32, p_i32, p_f64, p_i32, p_i32, p_f64]),
    "b200_complex_embed": (c_i32, [c_i32, p_i32, p_i32, p_f64, c_i32, p_i64, p_i32, p_i32, p_i32, p_f64]),
    "b200_mm_read": (c_i32, [ctypes.c_char_p, c_i32, p_i64, p_i32, p_i32, p_f64, c_i64]),
}

_lib = None


def load():
    """Loads the shared library once and sets the prototypes. Raises OSError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(
                "libsolver_b200.so has not been built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make` at the repo root. There is no CPU fallback." % LIB_PATH
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def ptr(arr, typ):
    return arr.ctypes.data_as(typ)
