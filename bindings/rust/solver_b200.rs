//! `Genie::B200` -- Rust side of the B200 backend for russell_sparse (drop-in for the factorize -> solve path).
//!
//! NOT COMPILED IN THE AUTHORING IMAGE (no rustc/cargo there): reviewed source, kept next to the C ABI it binds
//! (`include/solver_b200.h`).  It implements `LinSolTrait` (russell_sparse/src/lin_solver.rs:12-64) with the guards and
//! error strings of the reference's GPU wrapper (russell_sparse/src/solver_cudss.rs:194-360), but uses the COO-level entry
//! points: the triplet structure is analysed once, afterwards only `mat.values` crosses the FFI per refactorization and
//! the duplicate summation of `CsrMatrix::update_from_coo` (csr_matrix.rs:431-459) runs on the device.
//! `russell_b200/__init__.py: SolverB200` is the tested Python twin of this file (same calls, same order, same strings).
//!
//! Registration: see INTEGRATION.md section 2 (enums.rs, lin_solver.rs, lib.rs, Cargo.toml, build.rs, solve_matrix_market.rs).

use super::{CooMatrix, LinSolParams, LinSolTrait, Matching, Ordering, StatsLinSol, Sym};
use crate::constants::*;
use crate::StrError;
use russell_lab::{Stopwatch, Vector};

#[repr(C)]
struct InterfaceB200 {
    _data: [u8; 0],
    _marker: core::marker::PhantomData<(*mut u8, core::marker::PhantomPinned)>,
}
unsafe impl Send for InterfaceB200 {}
unsafe impl Send for SolverB200 {}

#[rustfmt::skip]
unsafe extern "C" {
    fn solver_b200_new() -> *mut InterfaceB200;
    fn solver_b200_drop(solver: *mut InterfaceB200);
    fn solver_b200_initialize_coo(solver: *mut InterfaceB200, ordering: i32, matching: i32, pivoting: i32, pivot_epsilon: f64,
        refinement_nstep: i32, hybrid_memory_factor: f64, verbose: CcBool, general_symmetric: CcBool, positive_definite: CcBool,
        ndim: i32, nnz_coo: i32, indices_i: *const i32, indices_j: *const i32, values: *const f64) -> i32;
    // the triplet indices travel with the values, like CsrMatrix::update_from_coo sees them on every call
    // (russell_sparse/src/solver_cudss.rs:209): same indices -> values only; re-ordered triplets -> slot map rebuilt;
    // another pattern -> 705
    fn solver_b200_factorize_coo_checked(solver: *mut InterfaceB200, effective_matching: *mut i32, effective_pivoting: *mut i32,
        verbose: CcBool, nnz_coo: i32, indices_i: *const i32, indices_j: *const i32, coo_values: *const f64) -> i32;
    fn solver_b200_rcond(solver: *mut InterfaceB200, rcond: *mut f64) -> i32;
    fn solver_b200_determinant(solver: *mut InterfaceB200, coefficient: *mut f64, exponent: *mut f64) -> i32;
    fn solver_b200_get_stats(solver: *mut InterfaceB200, out: *mut f64, n_out: i32) -> i32;
    fn solver_b200_solve(solver: *mut InterfaceB200, x: *mut f64, rhs: *const f64, verbose: CcBool) -> i32;
}

/// Same integers as `cudss_ordering` (russell_sparse/src/solver_cudss.rs:425-441) and as the tested Python twin: the C side
/// maps 3 to minimum degree, 5 to the natural order and everything else to nested dissection (include/solver_b200.h)
pub(crate) fn b200_ordering(ordering: Ordering) -> i32 {
    match ordering {
        Ordering::BtfColamd => 1,
        Ordering::Colamd => 2,
        Ordering::Amd => 3,
        Ordering::Metis => 4,
        Ordering::No => 5,
        _ => 0,
    }
}

/// Same integers as `cudss_matching` (solver_cudss.rs:459-469): 0 none (upgraded to "auto" by the library: zero diagonals get a
/// max-product matching + scaling), 6 auto, 1..5 force the matching
pub(crate) fn b200_matching(matching: Matching) -> i32 {
    match matching {
        Matching::None => 0,
        Matching::MaxDiagCount => 1,
        Matching::MaxMinDiag => 2,
        Matching::MaxMinDiagAlt => 3,
        Matching::MaxDiagSum => 4,
        Matching::MaxDiagProduct => 5,
        Matching::Auto => 6,
    }
}

pub(crate) fn handle_b200_error_code(err: i32) -> StrError {
    match err {
        1 => "Error(1): Matrix is singular",
        100 => "cudaMalloc failed in the C code (B200)",
        200 => "cudaMemcpy failed in the C code (B200)",
        300 => "cudaStreamSynchronize failed in the C code (B200)",
        701 => "B200 analysis failed: matrix is structurally singular",
        702 => "B200 analysis failed: invalid CSR structure",
        907 => "B200 solve failed: iterative refinement failed",
        705 => "subsequent factorizations must use the same matrix (the COO structure differs from the analysed one)",
        703 => "B200 analysis failed: invalid COO structure (index out of range or empty)",
        704 => "B200 analysis failed: Sym::YesLower requires triplets with j <= i",
        801 => "B200 numeric factorization failed: kernel launch failure",
        802 => "B200 numeric factorization failed: matrix values are not finite",
        901 => "B200 solve failed: kernel launch failure",
        ERROR_NULL_POINTER => "B200 failed due to NULL POINTER error",
        ERROR_MALLOC => "B200 failed due to MALLOC error",
        ERROR_NOT_AVAILABLE => "B200 solver is not AVAILABLE (no CUDA device)",
        ERROR_NEED_INITIALIZATION => "B200 failed because INITIALIZATION is needed",
        ERROR_NEED_FACTORIZATION => "B200 failed because FACTORIZATION is needed",
        ERROR_ALREADY_INITIALIZED => "B200 failed because INITIALIZATION has been completed already",
        _ => "Error: unknown error returned by c-code (B200)",
    }
}

/// What the first `factorize` froze: the structure must not change afterwards (lin_solver.rs:18-27)
#[derive(Clone, Copy)]
struct Frozen {
    sym: Sym,
    ndim: usize,
    nnz: usize,
}

/// Sparse direct solver running on one NVIDIA B200 (multifrontal LU, f64)
pub struct SolverB200 {
    handle: *mut InterfaceB200,
    frozen: Option<Frozen>,
    factorized: bool,
    effective_matching: i32,
    effective_pivoting: i32,
    stopwatch: Stopwatch,
    ns_init: u128,
    ns_fact: u128,
    ns_solve: u128,
}

impl Drop for SolverB200 {
    fn drop(&mut self) {
        unsafe { solver_b200_drop(self.handle) }
    }
}

impl SolverB200 {
    pub fn new() -> Result<Self, StrError> {
        let handle = unsafe { solver_b200_new() };
        if handle.is_null() {
            return Err("c-code failed to allocate the B200 solver");
        }
        Ok(SolverB200 {
            handle,
            frozen: None,
            factorized: false,
            effective_matching: 0,
            effective_pivoting: 0,
            stopwatch: Stopwatch::new(),
            ns_init: 0,
            ns_fact: 0,
            ns_solve: 0,
        })
    }
}

impl LinSolTrait for SolverB200 {
    fn factorize(&mut self, mat: &CooMatrix, params: Option<LinSolParams>) -> Result<(), StrError> {
        match self.frozen {
            Some(fz) => {
                if mat.symmetric != fz.sym {
                    return Err("subsequent factorizations must use the same matrix (symmetric differs)");
                }
                if mat.nrow != fz.ndim {
                    return Err("subsequent factorizations must use the same matrix (ndim differs)");
                }
                if mat.nnz != fz.nnz {
                    return Err("subsequent factorizations must use the same matrix (nnz differs)");
                }
                if params.is_some() {
                    return Err("subsequent factorizations must not change LinSolParams");
                }
            }
            None => {
                if mat.nrow != mat.ncol {
                    return Err("the matrix must be square");
                }
                if mat.nnz < 1 {
                    return Err("the COO matrix must have at least one non-zero value");
                }
                if mat.symmetric == Sym::YesUpper {
                    return Err("B200 requires Sym::YesLower or Sym::YesFull for symmetric matrices");
                }
                let par = params.unwrap_or_else(LinSolParams::new);
                let hybrid = match par.hybrid_memory_factor {
                    Some(v) if v < 0.01 || v > 0.99 => return Err("the hybrid memory factor must be in [0.01, 0.99]"),
                    Some(v) => v,
                    None => -1.0,
                };
                let lower = mat.symmetric == Sym::YesLower;
                self.stopwatch.reset();
                let status = unsafe {
                    solver_b200_initialize_coo(
                        self.handle,
                        b200_ordering(par.ordering),
                        b200_matching(par.matching),
                        0,
                        par.pivot_epsilon.unwrap_or(-1.0),
                        par.refinement_nstep.unwrap_or(-1),
                        hybrid,
                        if par.verbose { 1 } else { 0 },
                        if lower { 1 } else { 0 },
                        if par.positive_definite && lower { 1 } else { 0 },
                        to_i32(mat.nrow),
                        to_i32(mat.nnz),
                        mat.indices_i.as_ptr(),
                        mat.indices_j.as_ptr(),
                        mat.values.as_ptr(),
                    )
                };
                if status != SUCCESSFUL_EXIT {
                    return Err(handle_b200_error_code(status));
                }
                self.ns_init = self.stopwatch.stop();
                self.frozen = Some(Frozen { sym: mat.symmetric, ndim: mat.nrow, nnz: mat.nnz });
            }
        }
        self.factorized = false;
        self.stopwatch.reset();
        let status = unsafe {
            solver_b200_factorize_coo_checked(self.handle, &mut self.effective_matching, &mut self.effective_pivoting, 0,
                to_i32(mat.nnz), mat.indices_i.as_ptr(), mat.indices_j.as_ptr(), mat.values.as_ptr())
        };
        if status != SUCCESSFUL_EXIT {
            return Err(handle_b200_error_code(status));
        }
        self.ns_fact = self.stopwatch.stop();
        self.factorized = true;
        Ok(())
    }

    fn solve(&mut self, x: &mut Vector, rhs: &Vector, verbose: bool) -> Result<(), StrError> {
        let ndim = match (self.factorized, self.frozen) {
            (true, Some(fz)) => fz.ndim,
            _ => return Err("the function factorize must be called before solve"),
        };
        if x.dim() != ndim {
            return Err("the dimension of the vector of unknown values x is incorrect");
        }
        if rhs.dim() != ndim {
            return Err("the dimension of the right-hand side vector is incorrect");
        }
        self.stopwatch.reset();
        let status = unsafe {
            solver_b200_solve(self.handle, x.as_mut_data().as_mut_ptr(), rhs.as_data().as_ptr(), if verbose { 1 } else { 0 })
        };
        if status != SUCCESSFUL_EXIT {
            return Err(handle_b200_error_code(status));
        }
        self.ns_solve = self.stopwatch.stop();
        Ok(())
    }

    fn update_stats(&self, stats: &mut StatsLinSol) {
        stats.main.solver = "B200".to_string();
        stats.time_nanoseconds.initialize_array.push(self.ns_init);
        stats.time_nanoseconds.factorize_array.push(self.ns_fact);
        stats.time_nanoseconds.solve_array.push(self.ns_solve);
        stats.output.effective_matching = if self.effective_matching == 5 { "MaxDiagProduct" } else { "None" }.to_string();
        stats.output.effective_pivoting = "LocalBlock".to_string();
        // what SolverUMFPACK::update_stats fills from UMFPACK's Info array (solver_umfpack.rs:392-422)
        let mut st = [0.0_f64; 29]; // B200_STAT_COUNT
        if self.factorized && unsafe { solver_b200_get_stats(self.handle, st.as_mut_ptr(), 29) } == SUCCESSFUL_EXIT {
            stats.output.effective_ordering = match st[24] as i32 { 3 => "Amd", 5 => "No", _ => "Metis" }.to_string(); // B200_STAT_EFFECTIVE_ORDERING
            stats.output.effective_scaling = if st[25] != 0.0 { "Max" } else { "No" }.to_string();                      // B200_STAT_EFFECTIVE_SCALING
            let mut rcond = 0.0;
            if unsafe { solver_b200_rcond(self.handle, &mut rcond) } == SUCCESSFUL_EXIT {
                stats.output.umfpack_rcond_estimate = rcond; // UMFPACK's own definition: min|U_kk| / max|U_kk|
            }
            let (mut c, mut e) = (0.0, 0.0);
            if unsafe { solver_b200_determinant(self.handle, &mut c, &mut e) } == SUCCESSFUL_EXIT {
                stats.determinant.mantissa_real = c;
                stats.determinant.base = 10.0;
                stats.determinant.exponent = e;
            }
        }
    }

    fn get_ns_init(&self) -> u128 {
        self.ns_init
    }

    fn get_ns_fact(&self) -> u128 {
        self.ns_fact
    }

    fn get_ns_solve(&self) -> u128 {
        self.ns_solve
    }
}
