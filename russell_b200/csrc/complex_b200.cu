// complex_b200.cu -- Complex64 twin of the solver.
//
// Replaces the reference's complex cuDSS shim
//   russell_sparse/src/complex_solver_cudss.rs:32-64        (the Rust `extern "C"` block)
//   russell_sparse/c_code/interface_complex_cudss.cu:60-567 (complex_solver_cudss_{new,drop,initialize,factorize,solve})
// for the callers that need a complex system: russell_ode's Radau5 solves one real and one complex n x n system per
// Newton iteration (russell_ode/src/radau5.rs:45,51,264-301).
//
// Design: the complex system A z = c (A = Ar + i Ai) is solved as the equivalent REAL system of order 2n in
// *interleaved* unknowns, every complex entry becoming the 2x2 block
//        | ar  -ai |
//        | ai   ar |        rows (2i, 2i+1) = (Re, Im) of equation i, columns (2j, 2j+1) = (Re, Im) of unknown j,
// so a Complex64 vector (re, im pairs, the layout of russell_lab::ComplexVector) IS the real vector of the embedded
// system: rhs and x cross the boundary without any conversion, and the whole real pipeline (matching, nested
// dissection, multifrontal LU kernels, SpTRSV, SpMV residual, iterative refinement) is reused unchanged.
// A 2x2 real block multiply-add costs the same 8 flops as a complex multiply-add; the price of the embedding is
// 2x the factor bytes of a native complex factorization (32 B vs 16 B per entry).  Complex *symmetric* input
// (Sym::YesLower; A = A^T, not Hermitian) is mirrored to the full pattern first: its real embedding is not symmetric.
//
// Per refactorization the host ships only the nnz Complex64 values (16 B each); `k_complex_expand` writes the
// 4*nnz_full real CSR values on the device from a slot map built once at initialize.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <new>
#include <vector>

#include "../../include/solver_b200.h"

#include "coo_guard.hpp"
#include <atomic>
#include <thread>

// host side of the COO conversion and of the embedding (host_formats.cpp)
extern "C" int32_t b200_complex_embed(int32_t n, const int32_t* rp, const int32_t* ci, const double* values, int32_t lower,
                                      int64_t* info, int32_t* rptr, int32_t* rcol, int32_t* code, double* rval);

namespace {

// real CSR slot r of the embedded matrix <- complex slot (code >> 2), component by (code & 3):
//   0: +re   1: -im   2: +im   3: +re
__global__ void __launch_bounds__(256) k_complex_expand(long long nreal, const int* __restrict__ code, const double2* __restrict__ cv,
                                                        double* __restrict__ rv) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nreal; r += (long long)gridDim.x * blockDim.x) {
        const int c = __ldg(code + r);
        const double2 z = __ldg(cv + (c >> 2));
        const int k = c & 3;
        rv[r] = (k == 1) ? -z.y : (k == 2) ? z.y : z.x;
    }
}

// COO-level boundary: every CSR slot sums its triplets in their order of appearance (ComplexCsrMatrix::update_from_coo,
// csr_matrix.rs:431-459 over Complex64), on the device
__global__ void __launch_bounds__(256) k_complex_coo_to_csr_values(int nslots, const int* __restrict__ seg_ptr,
                                                                   const int* __restrict__ seg_idx, const double2* __restrict__ coo,
                                                                   double2* __restrict__ csr) {
    for (int sl = blockIdx.x * blockDim.x + threadIdx.x; sl < nslots; sl += gridDim.x * blockDim.x) {
        const int a = seg_ptr[sl], b = seg_ptr[sl + 1];
        double2 acc = __ldg(coo + seg_idx[a]);
        for (int e = a + 1; e < b; e++) {
            const double2 v = __ldg(coo + seg_idx[e]);
            acc.x += v.x, acc.y += v.y;
        }
        csr[sl] = acc;
    }
}

} // namespace

struct InterfaceComplexB200 {
    InterfaceB200* real = nullptr; // the order-2n real solver that does all the work
    int n = 0;                     // complex dimension
    int nnz = 0;                   // complex CSR entries the caller passes (lower triangle only for symmetric input)
    long long nreal = 0;           // entries of the embedded real CSR matrix (4 per full complex entry)
    bool initialized = false;
    int* d_code = nullptr;
    double2* d_cvals = nullptr;
    double* d_rvals = nullptr;
    // COO-level boundary (complex_solver_b200_initialize_coo)
    int nnz_coo = 0;
    int *d_seg_ptr = nullptr, *d_seg_idx = nullptr;
    double2* d_coo_cvals = nullptr;
    b200::CooGuard coo_guard;
};

#define CB_CUDA_TRY(call, code)             \
    do {                                    \
        if ((call) != cudaSuccess) {        \
            cudaGetLastError();             \
            return (code);                  \
        }                                   \
    } while (0)

extern "C" {

struct InterfaceComplexB200* complex_solver_b200_new(void) {
    InterfaceComplexB200* s = new (std::nothrow) InterfaceComplexB200();
    if (!s) return nullptr;
    s->real = solver_b200_new();
    if (!s->real) { // no device: no CPU fallback
        delete s;
        return nullptr;
    }
    return s;
}

void complex_solver_b200_drop(struct InterfaceComplexB200* s) {
    if (!s) return;
    if (s->real) {
        cudaSetDevice(solver_b200_get_device(s->real));
        cudaStream_t st = (cudaStream_t)solver_b200_get_stream(s->real);
        if (st) cudaStreamSynchronize(st);
    }
    if (s->d_code) cudaFree(s->d_code);
    if (s->d_cvals) cudaFree(s->d_cvals);
    if (s->d_rvals) cudaFree(s->d_rvals);
    if (s->d_seg_ptr) cudaFree(s->d_seg_ptr);
    if (s->d_seg_idx) cudaFree(s->d_seg_idx);
    if (s->d_coo_cvals) cudaFree(s->d_coo_cvals);
    solver_b200_drop(s->real);
    delete s;
}

int32_t complex_solver_b200_set_option(struct InterfaceComplexB200* s, const char* key, double value) {
    if (!s) return B200_ERROR_NULL_POINTER;
    return solver_b200_set_option(s->real, key, value);
}

struct InterfaceB200* complex_solver_b200_real_handle(struct InterfaceComplexB200* s) { return s ? s->real : nullptr; }

int32_t complex_solver_b200_initialize(struct InterfaceComplexB200* s, int32_t ordering, int32_t matching, int32_t pivoting,
                                       double pivot_epsilon, int32_t refinement_nstep, double hybrid_memory_factor,
                                       int32_t verbose, int32_t general_symmetric, int32_t positive_definite, int32_t ndim,
                                       const int32_t* row_pointers, const int32_t* col_indices, const double* values) {
    if (!s || !row_pointers || !col_indices || !values) return B200_ERROR_NULL_POINTER;
    if (s->initialized) return B200_ERROR_ALREADY_INITIALIZED;
    if (ndim < 1) return B200_ERROR_ANALYSIS + 2;
    const int n = ndim;
    const int nnz = row_pointers[n];
    const int lower = (general_symmetric != 0 || positive_definite != 0) ? 1 : 0;
    int64_t info[2] = {0, 0};
    if (b200_complex_embed(n, row_pointers, col_indices, values, lower, info, nullptr, nullptr, nullptr, nullptr) != 0)
        return B200_ERROR_ANALYSIS + 2;
    const long long nreal = info[1];
    std::vector<int32_t> rptr((size_t)2 * n + 1), rcol((size_t)nreal), code((size_t)nreal);
    std::vector<double> rval((size_t)nreal);
    if (b200_complex_embed(n, row_pointers, col_indices, values, lower, info, rptr.data(), rcol.data(), code.data(), rval.data()) != 0)
        return B200_ERROR_ANALYSIS + 2;
    (void)positive_definite; // the embedded matrix is handled as a general one (matching decides by itself)
    int32_t rc = solver_b200_initialize(s->real, ordering, matching, pivoting, pivot_epsilon, refinement_nstep,
                                        hybrid_memory_factor, verbose, 0, 0, 2 * n, rptr.data(), rcol.data(), rval.data());
    if (rc != B200_SUCCESSFUL_EXIT) return rc;
    CB_CUDA_TRY(cudaSetDevice(solver_b200_get_device(s->real)), B200_ERROR_NOT_AVAILABLE);
    cudaStream_t st = (cudaStream_t)solver_b200_get_stream(s->real);
    CB_CUDA_TRY(cudaMalloc(&s->d_code, (size_t)nreal * sizeof(int)), B200_ERROR_CUDA_MALLOC);
    CB_CUDA_TRY(cudaMalloc(&s->d_cvals, (size_t)nnz * sizeof(double2)), B200_ERROR_CUDA_MALLOC);
    CB_CUDA_TRY(cudaMalloc(&s->d_rvals, (size_t)nreal * sizeof(double)), B200_ERROR_CUDA_MALLOC);
    CB_CUDA_TRY(cudaMemcpyAsync(s->d_code, code.data(), (size_t)nreal * sizeof(int), cudaMemcpyHostToDevice, st), B200_ERROR_CUDA_MEMCPY);
    CB_CUDA_TRY(cudaStreamSynchronize(st), B200_ERROR_CUDA_SYNCHRONIZE);
    s->n = n, s->nnz = nnz, s->nreal = nreal;
    s->initialized = true;
    if (verbose)
        printf("complex_solver_b200_initialize: n = %d complex (%d real) unknowns, %d complex entries -> %lld real entries\n",
               n, 2 * n, nnz, nreal);
    return B200_SUCCESSFUL_EXIT;
}

int32_t complex_solver_b200_factorize_device(struct InterfaceComplexB200* s, const double* d_values) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->initialized) return B200_ERROR_NEED_INITIALIZATION;
    if (!d_values) return B200_ERROR_NULL_POINTER;
    CB_CUDA_TRY(cudaSetDevice(solver_b200_get_device(s->real)), B200_ERROR_NOT_AVAILABLE);
    cudaStream_t st = (cudaStream_t)solver_b200_get_stream(s->real);
    long long blocks = (s->nreal + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_complex_expand<<<(int)blocks, 256, 0, st>>>(s->nreal, s->d_code, (const double2*)d_values, s->d_rvals);
    if (cudaGetLastError() != cudaSuccess) return B200_ERROR_NUM_FACTORIZATION + 1;
    return solver_b200_factorize_device(s->real, s->d_rvals);
}

int32_t complex_solver_b200_factorize(struct InterfaceComplexB200* s, int32_t* effective_matching, int32_t* effective_pivoting,
                                      int32_t verbose, const double* values) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->initialized) return B200_ERROR_NEED_INITIALIZATION;
    if (!values) return B200_ERROR_NULL_POINTER;
    CB_CUDA_TRY(cudaSetDevice(solver_b200_get_device(s->real)), B200_ERROR_NOT_AVAILABLE);
    cudaStream_t st = (cudaStream_t)solver_b200_get_stream(s->real);
    if (solver_b200_copy_h2d(s->real, s->d_cvals, values, (int64_t)s->nnz * (int64_t)sizeof(double2)) != 0) return B200_ERROR_CUDA_MEMCPY; // (staged when pageable)
    int32_t rc = complex_solver_b200_factorize_device(s, (const double*)s->d_cvals);
    double st8[B200_STAT_COUNT];
    if (solver_b200_get_stats(s->real, st8, B200_STAT_COUNT) == 0) {
        if (effective_matching) *effective_matching = st8[B200_STAT_MATCHED] != 0.0 ? B200_MATCHING_MAX_DIAG_PRODUCT : B200_MATCHING_NONE;
        if (rc == 0 && verbose)
            printf("complex_solver_b200_factorize: numeric factorization completed in %.3f ms (device)\n", st8[B200_STAT_MS_FACTORIZE_DEVICE]);
    }
    if (effective_pivoting) *effective_pivoting = 5; // LocalBlock (solver_cudss.rs:393-466 numbering)
    return rc;
}

// ---- COO-level boundary for complex triplets: structure analysed once, duplicates summed on the device per call ------
int32_t complex_solver_b200_initialize_coo(struct InterfaceComplexB200* s, int32_t ordering, int32_t matching, int32_t pivoting,
                                           double pivot_epsilon, int32_t refinement_nstep, double hybrid_memory_factor,
                                           int32_t verbose, int32_t general_symmetric, int32_t positive_definite, int32_t ndim,
                                           int32_t nnz_coo, const int32_t* indices_i, const int32_t* indices_j, const double* values) {
    if (!s || !indices_i || !indices_j || !values) return B200_ERROR_NULL_POINTER;
    if (s->initialized) return B200_ERROR_ALREADY_INITIALIZED;
    if (ndim < 1 || nnz_coo < 1) return B200_ERROR_ANALYSIS + 3;
    if (general_symmetric || positive_definite)
        for (int32_t k = 0; k < nnz_coo; k++)
            if (indices_j[k] > indices_i[k]) return B200_ERROR_ANALYSIS + 4; // Sym::YesLower promised: j <= i
    std::vector<int32_t> ptr((size_t)ndim + 1), idx((size_t)nnz_coo), seg_ptr((size_t)nnz_coo + 1), seg_idx((size_t)nnz_coo);
    std::vector<double> re((size_t)nnz_coo), dummy((size_t)nnz_coo);
    for (int32_t k = 0; k < nnz_coo; k++) re[k] = values[2 * (size_t)k];
    if (b200_coo_to_csr_map(ndim, ndim, nnz_coo, indices_i, indices_j, re.data(), ptr.data(), idx.data(), dummy.data(), seg_ptr.data(),
                            seg_idx.data()) != 0)
        return B200_ERROR_ANALYSIS + 3;
    const int nslots = ptr[ndim];
    std::vector<double> csr((size_t)2 * nslots);
    for (int sl = 0; sl < nslots; sl++) { // the same sums the device kernel forms later (order of appearance)
        double ar = values[2 * (size_t)seg_idx[seg_ptr[sl]]], ai = values[2 * (size_t)seg_idx[seg_ptr[sl]] + 1];
        for (int e = seg_ptr[sl] + 1; e < seg_ptr[sl + 1]; e++) ar += values[2 * (size_t)seg_idx[e]], ai += values[2 * (size_t)seg_idx[e] + 1];
        csr[2 * (size_t)sl] = ar, csr[2 * (size_t)sl + 1] = ai;
    }
    int32_t rc = complex_solver_b200_initialize(s, ordering, matching, pivoting, pivot_epsilon, refinement_nstep, hybrid_memory_factor,
                                                verbose, general_symmetric, positive_definite, ndim, ptr.data(), idx.data(), csr.data());
    if (rc != B200_SUCCESSFUL_EXIT) return rc;
    cudaStream_t st = (cudaStream_t)solver_b200_get_stream(s->real);
    s->nnz_coo = nnz_coo;
    CB_CUDA_TRY(cudaMalloc(&s->d_seg_ptr, ((size_t)nslots + 1) * sizeof(int)), B200_ERROR_CUDA_MALLOC);
    CB_CUDA_TRY(cudaMalloc(&s->d_seg_idx, (size_t)nnz_coo * sizeof(int)), B200_ERROR_CUDA_MALLOC);
    CB_CUDA_TRY(cudaMalloc(&s->d_coo_cvals, (size_t)nnz_coo * sizeof(double2)), B200_ERROR_CUDA_MALLOC);
    CB_CUDA_TRY(cudaMemcpyAsync(s->d_seg_ptr, seg_ptr.data(), ((size_t)nslots + 1) * sizeof(int), cudaMemcpyHostToDevice, st), B200_ERROR_CUDA_MEMCPY);
    CB_CUDA_TRY(cudaMemcpyAsync(s->d_seg_idx, seg_idx.data(), (size_t)nnz_coo * sizeof(int), cudaMemcpyHostToDevice, st), B200_ERROR_CUDA_MEMCPY);
    CB_CUDA_TRY(cudaStreamSynchronize(st), B200_ERROR_CUDA_SYNCHRONIZE);
    s->coo_guard.remember(ndim, nnz_coo, indices_i, indices_j, ptr, idx);
    return B200_SUCCESSFUL_EXIT;
}

int32_t complex_solver_b200_factorize_coo(struct InterfaceComplexB200* s, int32_t* effective_matching, int32_t* effective_pivoting,
                                          int32_t verbose, const double* coo_values) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->initialized || !s->d_seg_ptr) return B200_ERROR_NEED_INITIALIZATION;
    if (!coo_values) return B200_ERROR_NULL_POINTER;
    CB_CUDA_TRY(cudaSetDevice(solver_b200_get_device(s->real)), B200_ERROR_NOT_AVAILABLE);
    cudaStream_t st = (cudaStream_t)solver_b200_get_stream(s->real);
    if (solver_b200_copy_h2d(s->real, s->d_coo_cvals, coo_values, (int64_t)s->nnz_coo * (int64_t)sizeof(double2)) != 0) return B200_ERROR_CUDA_MEMCPY;
    int blocks = (s->nnz + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_complex_coo_to_csr_values<<<blocks, 256, 0, st>>>(s->nnz, s->d_seg_ptr, s->d_seg_idx, s->d_coo_cvals, s->d_cvals);
    if (cudaGetLastError() != cudaSuccess) return B200_ERROR_NUM_FACTORIZATION + 1;
    int32_t rc = complex_solver_b200_factorize_device(s, (const double*)s->d_cvals);
    double st8[B200_STAT_COUNT];
    if (solver_b200_get_stats(s->real, st8, B200_STAT_COUNT) == 0) {
        if (effective_matching) *effective_matching = st8[B200_STAT_MATCHED] != 0.0 ? B200_MATCHING_MAX_DIAG_PRODUCT : B200_MATCHING_NONE;
        if (rc == 0 && verbose)
            printf("complex_solver_b200_factorize_coo: numeric factorization completed in %.3f ms (device)\n", st8[B200_STAT_MS_FACTORIZE_DEVICE]);
    }
    if (effective_pivoting) *effective_pivoting = 5;
    return rc;
}

int32_t complex_solver_b200_factorize_coo_checked(struct InterfaceComplexB200* s, int32_t* effective_matching, int32_t* effective_pivoting,
                                                  int32_t verbose, int32_t nnz_coo, const int32_t* indices_i, const int32_t* indices_j,
                                                  const double* coo_values) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->initialized || !s->d_seg_ptr || !s->coo_guard.armed()) return B200_ERROR_NEED_INITIALIZATION;
    if (!indices_i || !indices_j || !coo_values) return B200_ERROR_NULL_POINTER;
    if (nnz_coo != s->nnz_coo) return B200_ERROR_ANALYSIS + 5;
    std::atomic<int> same{1};
    std::thread checker([&] { same.store(s->coo_guard.same(indices_i, indices_j) ? 1 : 0); });
    int32_t rc = complex_solver_b200_factorize_coo(s, effective_matching, effective_pivoting, verbose, coo_values);
    checker.join();
    if (same.load()) return rc;
    std::vector<int32_t> seg_ptr, seg_idx;
    if (s->coo_guard.remap(indices_i, indices_j, seg_ptr, seg_idx) != 0) return B200_ERROR_ANALYSIS + 5;
    CB_CUDA_TRY(cudaSetDevice(solver_b200_get_device(s->real)), B200_ERROR_NOT_AVAILABLE);
    cudaStream_t st = (cudaStream_t)solver_b200_get_stream(s->real);
    CB_CUDA_TRY(cudaMemcpyAsync(s->d_seg_ptr, seg_ptr.data(), ((size_t)s->nnz + 1) * sizeof(int), cudaMemcpyHostToDevice, st), B200_ERROR_CUDA_MEMCPY);
    CB_CUDA_TRY(cudaMemcpyAsync(s->d_seg_idx, seg_idx.data(), (size_t)nnz_coo * sizeof(int), cudaMemcpyHostToDevice, st), B200_ERROR_CUDA_MEMCPY);
    CB_CUDA_TRY(cudaStreamSynchronize(st), B200_ERROR_CUDA_SYNCHRONIZE);
    return complex_solver_b200_factorize_coo(s, effective_matching, effective_pivoting, verbose, coo_values);
}

int32_t complex_solver_b200_solve(struct InterfaceComplexB200* s, double* x, const double* rhs, int32_t verbose) {
    if (!s) return B200_ERROR_NULL_POINTER;
    return solver_b200_solve(s->real, x, rhs, verbose); // Complex64[n] == f64[2n] in the interleaved embedding
}

int32_t complex_solver_b200_solve_device(struct InterfaceComplexB200* s, double* d_x, const double* d_rhs) {
    if (!s) return B200_ERROR_NULL_POINTER;
    return solver_b200_solve_device(s->real, d_x, d_rhs);
}

int32_t complex_solver_b200_spmv(struct InterfaceComplexB200* s, double* y, const double* x) {
    if (!s) return B200_ERROR_NULL_POINTER;
    return solver_b200_spmv(s->real, y, x);
}

int32_t complex_solver_b200_residual(struct InterfaceComplexB200* s, const double* x, const double* rhs, double* rel_residual) {
    if (!s) return B200_ERROR_NULL_POINTER;
    return solver_b200_residual(s->real, x, rhs, rel_residual); // ||.||_2 of a complex vector = ||.||_2 of its (re, im) pairs
}

int32_t complex_solver_b200_get_stats(struct InterfaceComplexB200* s, double* out, int32_t n_out) {
    if (!s) return B200_ERROR_NULL_POINTER;
    return solver_b200_get_stats(s->real, out, n_out);
}

} // extern "C"
