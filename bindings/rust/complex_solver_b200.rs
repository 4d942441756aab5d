//! Complex64 twin of `solver_b200.rs`: implements `ComplexLinSolTrait` (russell_sparse/src/complex_lin_solver.rs:12-64) over
//! `complex_solver_b200_*` (include/solver_b200.h), the counterpart of russell_sparse/src/complex_solver_cudss.rs.
//! NOT COMPILED IN THE AUTHORING IMAGE; `russell_b200/complex.py: ComplexSolverB200` is the tested Python twin.
//! Radau5 (russell_ode/src/radau5.rs:260-296) holds one `SolverB200` and one `ComplexSolverB200`; both may be driven from
//! different threads at the same time (independent handles, streams and device buffers).

use super::{ComplexCooMatrix, ComplexLinSolTrait, LinSolParams, StatsLinSol, Sym};
use crate::constants::*;
use crate::solver_b200::{b200_matching, b200_ordering, handle_b200_error_code};
use crate::StrError;
use russell_lab::{Complex64, ComplexVector, Stopwatch};

#[repr(C)]
struct InterfaceComplexB200 {
    _data: [u8; 0],
    _marker: core::marker::PhantomData<(*mut u8, core::marker::PhantomPinned)>,
}
unsafe impl Send for InterfaceComplexB200 {}
unsafe impl Send for ComplexSolverB200 {}

#[rustfmt::skip]
unsafe extern "C" {
    fn complex_solver_b200_new() -> *mut InterfaceComplexB200;
    fn complex_solver_b200_drop(solver: *mut InterfaceComplexB200);
    fn complex_solver_b200_initialize_coo(solver: *mut InterfaceComplexB200, ordering: i32, matching: i32, pivoting: i32,
        pivot_epsilon: f64, refinement_nstep: i32, hybrid_memory_factor: f64, verbose: CcBool, general_symmetric: CcBool,
        positive_definite: CcBool, ndim: i32, nnz_coo: i32, indices_i: *const i32, indices_j: *const i32,
        values: *const Complex64) -> i32;
    fn complex_solver_b200_factorize_coo_checked(solver: *mut InterfaceComplexB200, effective_matching: *mut i32,
        effective_pivoting: *mut i32, verbose: CcBool, nnz_coo: i32, indices_i: *const i32, indices_j: *const i32,
        coo_values: *const Complex64) -> i32;
    fn complex_solver_b200_solve(solver: *mut InterfaceComplexB200, x: *mut Complex64, rhs: *const Complex64, verbose: CcBool) -> i32;
}

pub struct ComplexSolverB200 {
    handle: *mut InterfaceComplexB200,
    frozen: Option<(Sym, usize, usize)>, // (symmetric, ndim, nnz) of the first factorize
    factorized: bool,
    effective_matching: i32,
    effective_pivoting: i32,
    stopwatch: Stopwatch,
    ns: [u128; 3], // initialize, factorize, solve
}

impl Drop for ComplexSolverB200 {
    fn drop(&mut self) {
        unsafe { complex_solver_b200_drop(self.handle) }
    }
}

impl ComplexSolverB200 {
    pub fn new() -> Result<Self, StrError> {
        let handle = unsafe { complex_solver_b200_new() };
        if handle.is_null() {
            return Err("c-code failed to allocate the B200 solver");
        }
        Ok(ComplexSolverB200 { handle, frozen: None, factorized: false, effective_matching: 0, effective_pivoting: 0,
                               stopwatch: Stopwatch::new(), ns: [0; 3] })
    }
}

impl ComplexLinSolTrait for ComplexSolverB200 {
    fn factorize(&mut self, mat: &ComplexCooMatrix, params: Option<LinSolParams>) -> Result<(), StrError> {
        if let Some((sym, ndim, nnz)) = self.frozen {
            if mat.symmetric != sym {
                return Err("subsequent factorizations must use the same matrix (symmetric differs)");
            }
            if mat.nrow != ndim {
                return Err("subsequent factorizations must use the same matrix (ndim differs)");
            }
            if mat.nnz != nnz {
                return Err("subsequent factorizations must use the same matrix (nnz differs)");
            }
            if params.is_some() {
                return Err("subsequent factorizations must not change LinSolParams");
            }
        } else {
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
            let lower = mat.symmetric == Sym::YesLower; // complex SYMMETRIC (A = A^T), like the reference
            self.stopwatch.reset();
            let status = unsafe {
                complex_solver_b200_initialize_coo(self.handle, b200_ordering(par.ordering), b200_matching(par.matching), 0,
                    par.pivot_epsilon.unwrap_or(-1.0), par.refinement_nstep.unwrap_or(-1), hybrid,
                    if par.verbose { 1 } else { 0 }, if lower { 1 } else { 0 },
                    if par.positive_definite && lower { 1 } else { 0 }, to_i32(mat.nrow), to_i32(mat.nnz),
                    mat.indices_i.as_ptr(), mat.indices_j.as_ptr(), mat.values.as_ptr())
            };
            if status != SUCCESSFUL_EXIT {
                return Err(handle_b200_error_code(status));
            }
            self.ns[0] = self.stopwatch.stop();
            self.frozen = Some((mat.symmetric, mat.nrow, mat.nnz));
        }
        self.factorized = false;
        self.stopwatch.reset();
        let status = unsafe {
            complex_solver_b200_factorize_coo_checked(self.handle, &mut self.effective_matching, &mut self.effective_pivoting, 0,
                                                      to_i32(mat.nnz), mat.indices_i.as_ptr(), mat.indices_j.as_ptr(), mat.values.as_ptr())
        };
        if status != SUCCESSFUL_EXIT {
            return Err(handle_b200_error_code(status));
        }
        self.ns[1] = self.stopwatch.stop();
        self.factorized = true;
        Ok(())
    }

    fn solve(&mut self, x: &mut ComplexVector, rhs: &ComplexVector, verbose: bool) -> Result<(), StrError> {
        let ndim = match (self.factorized, self.frozen) {
            (true, Some((_, ndim, _))) => ndim,
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
            complex_solver_b200_solve(self.handle, x.as_mut_data().as_mut_ptr(), rhs.as_data().as_ptr(), if verbose { 1 } else { 0 })
        };
        if status != SUCCESSFUL_EXIT {
            return Err(handle_b200_error_code(status));
        }
        self.ns[2] = self.stopwatch.stop();
        Ok(())
    }

    fn update_stats(&self, stats: &mut StatsLinSol) {
        stats.main.solver = "B200".to_string();
        stats.time_nanoseconds.initialize_array.push(self.ns[0]);
        stats.time_nanoseconds.factorize_array.push(self.ns[1]);
        stats.time_nanoseconds.solve_array.push(self.ns[2]);
        stats.output.effective_matching = if self.effective_matching == 5 { "MaxDiagProduct" } else { "None" }.to_string();
        stats.output.effective_pivoting = "LocalBlock".to_string();
    }

    fn get_ns_init(&self) -> u128 {
        self.ns[0]
    }

    fn get_ns_fact(&self) -> u128 {
        self.ns[1]
    }

    fn get_ns_solve(&self) -> u128 {
        self.ns[2]
    }
}
