/* solver_b200.h -- C ABI of the B200-native sparse direct solver backend for russell_sparse.
 *
 * Drop-in seam: these five entry points have exactly the shape of the reference's cuDSS shim
 *   russell_sparse/src/solver_cudss.rs:25-52        (the Rust `extern "C"` block that binds them)
 *   russell_sparse/c_code/interface_cudss.cu:62-566 (solver_cudss_{new,drop,initialize,factorize,solve})
 * so a `SolverB200` Rust wrapper cloned from `SolverCUDSS` (see bindings/rust/solver_b200.rs and
 * INTEGRATION.md) binds them without any change to LinSolTrait callers (russell_ode Radau5/BwEuler,
 * russell_pde, russell_nonlin).  Status codes follow russell_sparse/c_code/constants.h:5-36.
 *
 * Matrix layout (same as the reference's CSR contract, russell_sparse/src/csr_matrix.rs:359-480):
 *   CSR, 0-based int32 indices, duplicates already summed, nnz = row_pointers[ndim]; `values` may be longer
 *   than nnz (only the first nnz entries are read).  With general_symmetric=1 only the LOWER triangle is given.
 * Host arrays are borrowed for the duration of a call; nothing is retained.  All calls are synchronous.
 * One handle is used by one thread at a time; distinct handles may be driven concurrently from different
 * threads (russell_ode/src/radau5.rs:270-296) and a handle may migrate between threads between calls.
 */
#ifndef SOLVER_B200_H
#define SOLVER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------------------ */
#define B200_SUCCESSFUL_EXIT 0                   /* constants.h:5  SUCCESSFUL_EXIT */
#define B200_ERROR_SINGULAR 1                    /* parity with UMFPACK status 1 "Matrix is singular" (solver_umfpack.rs:492) */
#define B200_ERROR_NULL_POINTER 100000           /* constants.h:6  */
#define B200_ERROR_MALLOC 200000                 /* constants.h:7  */
#define B200_ERROR_VERSION 300000                /* constants.h:8  */
#define B200_ERROR_NOT_AVAILABLE 400000          /* constants.h:9  (no CUDA device / kernels not loadable) */
#define B200_ERROR_NEED_INITIALIZATION 500000    /* constants.h:10 */
#define B200_ERROR_NEED_FACTORIZATION 600000     /* constants.h:11 */
#define B200_ERROR_ALREADY_INITIALIZED 700000    /* constants.h:12 */
#define B200_ERROR_CUDA_MALLOC 100               /* constants.h:22 */
#define B200_ERROR_CUDA_MEMCPY 200               /* constants.h:23 */
#define B200_ERROR_CUDA_SYNCHRONIZE 300          /* constants.h:24 */
#define B200_ERROR_ANALYSIS 700                  /* phase offset like ERROR_CUDSS_SYM_FACTORIZATION; +1 structurally singular, +2 invalid CSR, +3 invalid COO, +4 COO not lower, +5 COO structure differs from the analysed one */
#define B200_ERROR_NUM_FACTORIZATION 800         /* phase offset like ERROR_CUDSS_NUM_FACTORIZATION; +1 kernel launch failure, +2 non-finite values */
#define B200_ERROR_SOLVE 900                     /* phase offset like ERROR_CUDSS_SOLVE; +1 kernel launch failure, +7 refinement failed (residual NaN, or above 10 x ir_tol with a backward error above rounding level) */

/* ---- option values (same integers the Rust maps for cuDSS send, solver_cudss.rs:393-466) ------------------- */
#define B200_ORDERING_DEFAULT 0  /* nested dissection */
#define B200_ORDERING_AMD 3      /* minimum degree family (small problems: exact minimum degree) */
#define B200_ORDERING_ND 4
#define B200_ORDERING_NONE 5     /* natural order */
#define B200_MATCHING_NONE 0
#define B200_MATCHING_MAX_DIAG_PRODUCT 5
#define B200_MATCHING_AUTO 6

struct InterfaceB200; /* opaque: stream, device buffers, symbolic plan, numeric factors */

/* replaces solver_cudss_new (interface_cudss.cu:62-123): NULL on failure (no device, out of memory) */
struct InterfaceB200 *solver_b200_new(void);

/* replaces solver_cudss_drop (interface_cudss.cu:126-171): NULL-safe, releases everything */
void solver_b200_drop(struct InterfaceB200 *solver);

/* replaces solver_cudss_initialize (interface_cudss.cu:190-396): once per structure.
 * Runs the host analysis (matching/scaling, ordering, elimination tree, supernodes, front plan) and
 * uploads the plan.  pivot_epsilon<=0, refinement_nstep<0 and hybrid_memory_factor<=0 select defaults
 * (hybrid memory is accepted and ignored: 180 GB of HBM3e hold the factors). */
int32_t solver_b200_initialize(struct InterfaceB200 *solver,
                               int32_t ordering, int32_t matching, int32_t pivoting,
                               double pivot_epsilon, int32_t refinement_nstep, double hybrid_memory_factor,
                               int32_t verbose, int32_t general_symmetric, int32_t positive_definite,
                               int32_t ndim, const int32_t *row_pointers, const int32_t *col_indices,
                               const double *values);

/* replaces solver_cudss_factorize (interface_cudss.cu:406-501): numeric LU of new values, same pattern */
int32_t solver_b200_factorize(struct InterfaceB200 *solver, int32_t *effective_matching,
                              int32_t *effective_pivoting, int32_t verbose, const double *values);

/* replaces solver_cudss_solve (interface_cudss.cu:510-566): x <- A^{-1} rhs with iterative refinement */
int32_t solver_b200_solve(struct InterfaceB200 *solver, double *x, const double *rhs, int32_t verbose);

/* ---- extensions (not part of the reference's five; used by bench.py, tests and the stats block) -------------- */

/* device-resident variants: pointers are DEVICE pointers on the handle's device; no host<->device copies */
int32_t solver_b200_factorize_device(struct InterfaceB200 *solver, const double *d_values);
int32_t solver_b200_solve_device(struct InterfaceB200 *solver, double *d_x, const double *d_rhs);

/* COO-level boundary extension.  The reference's wrapper re-runs CsrMatrix::update_from_coo on the
 * HOST on every factorize (russell_sparse/src/solver_cudss.rs:209, csr_matrix.rs:359-480: counting sort + duplicate
 * summation, tens of ms at 5M triplets).  Here the triplet STRUCTURE is analysed once (CSR built on the host with the
 * same duplicate-summation order, plus a triplet->CSR-slot map), and every refactorization only ships the raw triplet
 * VALUES: a device kernel sums the duplicates of every CSR slot in their order of appearance (bit-identical to the host
 * conversion) and the numeric factorization follows.
 * initialize_coo takes the same option arguments as solver_b200_initialize; with general_symmetric (Sym::YesLower)
 * the triplets must satisfy j <= i (B200_ERROR_ANALYSIS+4 otherwise); +3 = index out of range / empty.
 * factorize_coo_device takes a DEVICE pointer to nnz_coo values (e.g. an assembly kernel's output). */
int32_t solver_b200_initialize_coo(struct InterfaceB200 *solver,
                                   int32_t ordering, int32_t matching, int32_t pivoting,
                                   double pivot_epsilon, int32_t refinement_nstep, double hybrid_memory_factor,
                                   int32_t verbose, int32_t general_symmetric, int32_t positive_definite,
                                   int32_t ndim, int32_t nnz_coo, const int32_t *indices_i, const int32_t *indices_j,
                                   const double *values);
int32_t solver_b200_factorize_coo(struct InterfaceB200 *solver, int32_t *effective_matching,
                                  int32_t *effective_pivoting, int32_t verbose, const double *coo_values);
int32_t solver_b200_factorize_coo_device(struct InterfaceB200 *solver, const double *d_coo_values);
/* factorize_coo that also receives the triplet indices, like CsrMatrix::update_from_coo does on every call in the reference
 * (solver_cudss.rs:209): identical indices -> same as factorize_coo (the comparison runs on a helper thread underneath the
 * copy and the kernels); same pattern, triplets in another order -> the slot map is rebuilt; another pattern or nnz ->
 * B200_ERROR_ANALYSIS+5 (the structure is frozen after the first call, solver_cudss.rs:196-208) */
int32_t solver_b200_factorize_coo_checked(struct InterfaceB200 *solver, int32_t *effective_matching,
                                          int32_t *effective_pivoting, int32_t verbose, int32_t nnz_coo,
                                          const int32_t *indices_i, const int32_t *indices_j, const double *coo_values);

/* residual r = rhs - A x with the CSR SpMV kernel; returns ||r||_2 / ||rhs||_2 through *rel_residual
 * (russell's VerifyLinSys / the north-star accuracy metric).  Host pointers. */
int32_t solver_b200_residual(struct InterfaceB200 *solver, const double *x, const double *rhs, double *rel_residual);

/* y <- A x on the device (host pointers; SpMV kernel parity tests) */
int32_t solver_b200_spmv(struct InterfaceB200 *solver, double *y, const double *x);

/* reciprocal condition number estimate, UMFPACK's definition: min|U_kk| / max|U_kk| (Info[UMFPACK_RCOND],
 * interface_umfpack.c:179-184 -> StatsLinSol.output.umfpack_rcond_estimate) */
int32_t solver_b200_rcond(struct InterfaceB200 *solver, double *rcond);

/* determinant as mantissa * 10^exponent (StatsLinSol / solver_umfpack.rs:141-152 convention) */
int32_t solver_b200_determinant(struct InterfaceB200 *solver, double *coefficient, double *exponent);

/* stats: fills out[0..n_out-1]; see B200_STAT_* */
#define B200_STAT_NNODES 0
#define B200_STAT_NLEVELS 1
#define B200_STAT_NNZ_L 2
#define B200_STAT_NNZ_U 3
#define B200_STAT_FLOPS 4
#define B200_STAT_FAC_BYTES 5
#define B200_STAT_CB_BYTES 6
#define B200_STAT_MAX_FRONT 7
#define B200_STAT_T_ORDER_S 8
#define B200_STAT_T_SYMBOLIC_S 9
#define B200_STAT_N_PERTURBED 10
#define B200_STAT_LAST_REL_RESIDUAL 11
#define B200_STAT_LAST_REFINE_STEPS 12
#define B200_STAT_MS_FACTORIZE_DEVICE 13   /* CUDA-event time of the last numeric factorization (kernels only) */
#define B200_STAT_MS_SOLVE_DEVICE 14       /* CUDA-event time of the last solve incl. refinement (kernels only) */
#define B200_STAT_MS_SPTRSV_DEVICE 15      /* CUDA-event time of the last forward+backward sweep */
#define B200_STAT_MS_SPMV_DEVICE 16        /* CUDA-event time of the last residual SpMV */
#define B200_STAT_LAUNCHES_FACTORIZE 17    /* kernels launched per factorization */
#define B200_STAT_LAUNCHES_SOLVE 18        /* kernels launched by the last solve */
#define B200_STAT_SPTRSV_BYTES 19          /* algorithmic bytes of one forward+backward sweep */
#define B200_STAT_SPMV_BYTES 20            /* algorithmic bytes of one SpMV */
#define B200_STAT_MATCHED 21
#define B200_STAT_T_MATCH_S 22
#define B200_STAT_LAST_BACKWARD_ERROR 23 /* max_i |r_i| / (|A||x|+|b|)_i of the last solve */
#define B200_STAT_EFFECTIVE_ORDERING 24   /* B200_ORDERING_ND / _AMD / _NONE: what the analysis actually ran (UMFPACK_ORDERING_USED, interface_umfpack.c:180) */
#define B200_STAT_EFFECTIVE_SCALING 25    /* 0 = none, 1 = row/column scaling from the max-product matching (UMFPACK_SCALE, interface_umfpack.c:181) */
#define B200_STAT_RCOND 26                /* last solver_b200_rcond value, -1 when not computed for the current factors */
#define B200_STAT_T_INITIALIZE_HOST_S 27  /* wall time of the host analysis (matching + ordering + symbolic + plan) */
#define B200_STAT_PLAN_CACHE_HIT 28        /* 1: the plan was served from the process-wide cache of analysed patterns */
#define B200_STAT_COUNT 29
int32_t solver_b200_get_stats(struct InterfaceB200 *solver, double *out, int32_t n_out);

/* tuning knobs, to be set before initialize ("ir_tol", "refinement_nstep" and "strict_residual" also later).  Host analysis:
 * "panel_width", "nd_leaf", "relax_small", "relax_z1", "relax_z2", "relax_z3", "force_no_matching", "use_subtree", "subtree_maxf",
 * "subtree_budget".  Execution: "device", "use_graph", "ir_tol", "refinement_nstep", "strict_residual", "trace".  One fallback
 * per stage is kept next to the default kernel: "diag_variant" (4 = register-resident k_diag_w8, 0 = shared-memory
 * k_diag), "panel_variant" (1 = k_panel_warp / k_panel_row, 0 = k_panel), "panel_row_max", "schur_variant" (1 = f64 DMMA, 0 = FMA,
 * 2 = DMMA + the tcgen05 int8 Ozaki kernel for fronts with at least "ozaki_min_u" update rows), "fuse_chain", "use_fused",
 * "fused_variant", "fused_maxf", "fused_w8_max", "use_front_warp" (1 = one warp per front for the large fused launches, 0 = k_front_fused), "use_level_fork" (1 = the independent
 * launches of a tree level run on parallel branches of the graph), "inv_overlap" (1 = the pivot-block inverses of the fronts below
 * the chain levels run on a low-priority branch underneath those levels), "schur_front_nt" (fronts with at least this many rows of
 * 64 x 64 Schur tiles get a launch of their own whose grid enumerates the tiles; smaller fronts share a launch over a tile list),
 * "staged_copy" (1 = large pageable host buffers are staged through pinned memory by a few threads),
 * "use_leaf_reg", "asm_variant", "use_top", "top_max_nodes".  Unknown keys return
 * B200_ERROR_NOT_AVAILABLE. */
int32_t solver_b200_set_option(struct InterfaceB200 *solver, const char *key, double value);

/* debug/parity: copies factor panels (fac), pivot-block inverses (dinv) and local pivots to host buffers
 * sized by get_stats (FAC_BYTES/8 ...); any pointer may be NULL */
int32_t solver_b200_debug_copy_factors(struct InterfaceB200 *solver, double *fac, int64_t fac_len,
                                       double *dinv, int64_t dinv_len, int32_t *lperm, int64_t n);

/* debug/profiling: per-item %globaltimer timestamps (ns) of the last persistent top-of-tree sweeps -- 4 per item (start,
 * dependency satisfied, end, unused), the forward items first, then the backward ones (2 * 4 * n values); desc gets
 * (front, level, slice, rows) per item.  Needs set_option("trace", 1) before initialize.  Returns the item count. */
int32_t solver_b200_debug_trace(struct InterfaceB200 *solver, unsigned long long *out, int32_t *desc, int32_t cap);

/* the CUDA stream (cudaStream_t) every kernel of this handle is launched on, and its device ordinal: lets a
 * caller bracket calls with its own CUDA events (bench.py) or order its own work after ours */
void *solver_b200_get_stream(struct InterfaceB200 *solver);

/* extension: transfers between a caller's host buffer and device memory on the handle's stream.  Pinned host memory goes
 * straight to the copy engine (asynchronous, like cudaMemcpyAsync); large pageable buffers -- what a Rust Vec<f64> or a numpy
 * array is -- are striped over a few threads and staged through pinned buffers of the handle (H2D: ordered on the stream;
 * D2H: complete on return).  solver_b200_factorize / _factorize_coo / _solve and the complex twins use it for values, rhs and x. */
int32_t solver_b200_copy_h2d(struct InterfaceB200 *solver, void *dst_device, const void *src_host, int64_t bytes);
int32_t solver_b200_copy_d2h(struct InterfaceB200 *solver, void *dst_host, const void *src_device, int64_t bytes);
int32_t solver_b200_get_device(struct InterfaceB200 *solver);

/* ---- Complex64 twin ---------------------------------------------------------------------
 * Replaces the reference's complex cuDSS shim, same shapes:
 *   russell_sparse/src/complex_solver_cudss.rs:32-64          (Rust `extern "C"` block)
 *   russell_sparse/c_code/interface_complex_cudss.cu:60-567   (complex_solver_cudss_{new,drop,initialize,factorize,solve})
 * `values`, `x`, `rhs` point to Complex64 data = interleaved (re, im) f64 pairs (russell_lab::Complex64 is
 * num_complex::Complex<f64>, #[repr(C)]): 2*nnz / 2*ndim doubles.  CSR contract as above (sorted columns, duplicates
 * summed, lower triangle only when general_symmetric=1: complex SYMMETRIC, not Hermitian, like the reference).
 * Implementation: equivalent real system of order 2n in interleaved unknowns (russell_b200/csrc/complex_b200.cu). */
struct InterfaceComplexB200;
struct InterfaceComplexB200 *complex_solver_b200_new(void);
void complex_solver_b200_drop(struct InterfaceComplexB200 *solver);
int32_t complex_solver_b200_initialize(struct InterfaceComplexB200 *solver,
                                       int32_t ordering, int32_t matching, int32_t pivoting,
                                       double pivot_epsilon, int32_t refinement_nstep, double hybrid_memory_factor,
                                       int32_t verbose, int32_t general_symmetric, int32_t positive_definite,
                                       int32_t ndim, const int32_t *row_pointers, const int32_t *col_indices,
                                       const double *values /* Complex64[nnz] */);
int32_t complex_solver_b200_factorize(struct InterfaceComplexB200 *solver, int32_t *effective_matching,
                                      int32_t *effective_pivoting, int32_t verbose, const double *values /* Complex64[nnz] */);
int32_t complex_solver_b200_solve(struct InterfaceComplexB200 *solver, double *x /* Complex64[ndim] out */,
                                  const double *rhs /* Complex64[ndim] */, int32_t verbose);
/* COO-level boundary for complex triplets (the complex twin of solver_b200_initialize_coo / _factorize_coo): the triplet
 * structure is analysed once; every refactorization ships the raw Complex64 triplet values and the duplicates are summed
 * per CSR slot on the device, in their order of appearance (ComplexCsrMatrix::update_from_coo, csr_matrix.rs:431-459) */
int32_t complex_solver_b200_initialize_coo(struct InterfaceComplexB200 *solver,
                                           int32_t ordering, int32_t matching, int32_t pivoting,
                                           double pivot_epsilon, int32_t refinement_nstep, double hybrid_memory_factor,
                                           int32_t verbose, int32_t general_symmetric, int32_t positive_definite,
                                           int32_t ndim, int32_t nnz_coo, const int32_t *indices_i, const int32_t *indices_j,
                                           const double *values /* Complex64[nnz_coo] */);
int32_t complex_solver_b200_factorize_coo(struct InterfaceComplexB200 *solver, int32_t *effective_matching,
                                          int32_t *effective_pivoting, int32_t verbose, const double *coo_values);
int32_t complex_solver_b200_factorize_coo_checked(struct InterfaceComplexB200 *solver, int32_t *effective_matching,
                                                  int32_t *effective_pivoting, int32_t verbose, int32_t nnz_coo,
                                                  const int32_t *indices_i, const int32_t *indices_j, const double *coo_values);
/* extensions, as for the real solver: device-resident variants, A x and residual through the SpMV kernel, stats and
 * options of the underlying order-2n real handle */
int32_t complex_solver_b200_factorize_device(struct InterfaceComplexB200 *solver, const double *d_values);
int32_t complex_solver_b200_solve_device(struct InterfaceComplexB200 *solver, double *d_x, const double *d_rhs);
int32_t complex_solver_b200_spmv(struct InterfaceComplexB200 *solver, double *y, const double *x);
int32_t complex_solver_b200_residual(struct InterfaceComplexB200 *solver, const double *x, const double *rhs, double *rel_residual);
int32_t complex_solver_b200_get_stats(struct InterfaceComplexB200 *solver, double *out, int32_t n_out);
int32_t complex_solver_b200_set_option(struct InterfaceComplexB200 *solver, const char *key, double value);
struct InterfaceB200 *complex_solver_b200_real_handle(struct InterfaceComplexB200 *solver);

/* the tcgen05 Schur-complement kernel on its own (tests / profiling): C (u x u, column-major) <- C - A * B^T, A and B u x k
 * column-major, all HOST pointers.  Same operand split (8 signed 7-bit slices per entry), tile layout and kernel
 * (tcgen05.mma.kind::i8, TMEM accumulators, TMA bulk loads, f64 recombination) as schur_variant = 2 of the factorization:
 * russell_b200/csrc/ozaki_tc.cuh.  *ms = device time of split + GEMM. */
int32_t solver_b200_ozaki_gemm(int32_t u, int32_t k, const double *a, const double *b, double *c, double *ms);

/* library identification string (static storage) */
const char *solver_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SOLVER_B200_H */
