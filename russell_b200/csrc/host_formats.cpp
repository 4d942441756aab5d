// host_formats.cpp -- host-side data formats either side of the solver (C ABI, no CUDA needed).
//
// These are what the reference's Rust wrapper does before crossing the FFI:
//   b200_coo_to_csr / b200_coo_to_csc  <- CsrMatrix::update_from_coo (russell_sparse/src/csr_matrix.rs:359-480)
//                                         CscMatrix::update_from_coo (russell_sparse/src/csc_matrix.rs:365-505)
//       contract: duplicates (i,j) are SUMMED in their order of appearance, rows are sorted by column,
//       final nnz = pointers[n] may be smaller than the triplet count.
//   b200_mm_read_*                     <- read_matrix_market (russell_sparse/src/read_matrix_market.rs:346-475)
//       coordinate real general/symmetric, 1-based -> 0-based, MMsym handling of symmetric files.
// The implementation is a stable two-key counting/merge sort, not the reference's workspace scheme.
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

// Converts triplets (with duplicates) to compressed rows; W doubles per value (1 = real, 2 = interleaved complex).
template <int W>
static int32_t coo_to_csr_impl(int32_t nrow, int32_t ncol, int32_t nnz, const int32_t* ai, const int32_t* aj, const double* ax,
                               int32_t* ptr, int32_t* idx, double* val) {
    if (nnz < 1) return -2;
    std::vector<int32_t> start(nrow + 1, 0);
    for (int32_t k = 0; k < nnz; k++) {
        if (ai[k] < 0 || ai[k] >= nrow || aj[k] < 0 || aj[k] >= ncol) return -1;
        start[ai[k] + 1]++;
    }
    for (int32_t i = 0; i < nrow; i++) start[i + 1] += start[i];
    // stable bucket by row: order[] lists triplet ids row by row, in order of appearance
    std::vector<int32_t> order(nnz), fill(start.begin(), start.end() - 1);
    for (int32_t k = 0; k < nnz; k++) order[fill[ai[k]]++] = k;
    int32_t out = 0;
    ptr[0] = 0;
    for (int32_t i = 0; i < nrow; i++) {
        int32_t* b = order.data() + start[i];
        int32_t* e = order.data() + start[i + 1];
        std::stable_sort(b, e, [&](int32_t x, int32_t y) { return aj[x] < aj[y]; });
        for (int32_t* q = b; q != e;) {
            const int32_t j = aj[*q];
            double s[W];
            for (int c = 0; c < W; c++) s[c] = ax[(size_t)W * *q + c];
            for (++q; q != e && aj[*q] == j; ++q) // duplicates: summed in order of appearance
                for (int c = 0; c < W; c++) s[c] += ax[(size_t)W * *q + c];
            idx[out] = j;
            for (int c = 0; c < W; c++) val[(size_t)W * out + c] = s[c];
            out++;
        }
        ptr[i + 1] = out;
    }
    return 0;
}

extern "C" {

// Real triplets -> CSR.  ptr has nmajor+1 entries; idx/val have nnz slots, of which the first ptr[nmajor] are valid on
// return.  Returns 0, or a negative error: -1 index out of range, -2 nnz < 1
int32_t b200_coo_to_csr(int32_t nrow, int32_t ncol, int32_t nnz, const int32_t* ai, const int32_t* aj, const double* ax,
                        int32_t* ptr, int32_t* idx, double* val) {
    return coo_to_csr_impl<1>(nrow, ncol, nnz, ai, aj, ax, ptr, idx, val);
}

// Complex twin (ComplexCsrMatrix::update_from_coo is the same generic code, csr_matrix.rs:359-480 over Complex64):
// ax / val hold interleaved (re, im) pairs, 2*nnz doubles.
int32_t b200_complex_coo_to_csr(int32_t nrow, int32_t ncol, int32_t nnz, const int32_t* ai, const int32_t* aj,
                                const double* ax, int32_t* ptr, int32_t* idx, double* val) {
    return coo_to_csr_impl<2>(nrow, ncol, nnz, ai, aj, ax, ptr, idx, val);
}

// Same conversion, additionally returning the triplet -> slot map: seg_ptr[nslots+1] / seg_idx[nnz] list, for every
// CSR slot, the triplets summed into it in order of appearance.  (row_pointers, col_indices, values) as above.
int32_t b200_coo_to_csr_map(int32_t nrow, int32_t ncol, int32_t nnz, const int32_t* ai, const int32_t* aj, const double* ax,
                            int32_t* ptr, int32_t* idx, double* val, int32_t* seg_ptr, int32_t* seg_idx) {
    if (nnz < 1) return -2;
    std::vector<int32_t> start(nrow + 1, 0);
    for (int32_t k = 0; k < nnz; k++) {
        if (ai[k] < 0 || ai[k] >= nrow || aj[k] < 0 || aj[k] >= ncol) return -1;
        start[ai[k] + 1]++;
    }
    for (int32_t i = 0; i < nrow; i++) start[i + 1] += start[i];
    std::vector<int32_t> order(nnz), fill(start.begin(), start.end() - 1);
    for (int32_t k = 0; k < nnz; k++) order[fill[ai[k]]++] = k;
    int32_t out = 0, w = 0;
    ptr[0] = 0;
    seg_ptr[0] = 0;
    for (int32_t i = 0; i < nrow; i++) {
        int32_t* b = order.data() + start[i];
        int32_t* e = order.data() + start[i + 1];
        std::stable_sort(b, e, [&](int32_t x, int32_t y) { return aj[x] < aj[y]; });
        for (int32_t* q = b; q != e;) {
            const int32_t j = aj[*q];
            double s = ax[*q];
            seg_idx[w++] = *q;
            for (++q; q != e && aj[*q] == j; ++q) s += ax[*q], seg_idx[w++] = *q;
            idx[out] = j;
            val[out] = s;
            out++;
            seg_ptr[out] = w;
        }
        ptr[i + 1] = out;
    }
    return 0;
}

int32_t b200_coo_to_csc(int32_t nrow, int32_t ncol, int32_t nnz, const int32_t* ai, const int32_t* aj, const double* ax,
                        int32_t* ptr, int32_t* idx, double* val) {
    // columns of A are the rows of A^T; the duplicate-summation order (order of appearance) is unchanged
    return b200_coo_to_csr(ncol, nrow, nnz, aj, ai, ax, ptr, idx, val);
}

int32_t b200_complex_coo_to_csc(int32_t nrow, int32_t ncol, int32_t nnz, const int32_t* ai, const int32_t* aj,
                                const double* ax, int32_t* ptr, int32_t* idx, double* val) {
    return b200_complex_coo_to_csr(ncol, nrow, nnz, aj, ai, ax, ptr, idx, val);
}

// ---- Complex64 -> real embedding (used by complex_b200.cu; exported so that the CPU tests can check it) -------------
// The complex system A z = c becomes the real system of order 2n in interleaved unknowns: every complex entry a = ar + i ai
// at (i, j) becomes the 2x2 block [ar -ai; ai ar] at rows (2i, 2i+1), columns (2j, 2j+1).  Input: complex CSR (sorted
// columns, no duplicates; `lower` = only j <= i given, complex SYMMETRIC A = A^T, mirrored here).  Output: real CSR of
// order 2n plus, per real slot, code = 4*source_slot + k with k: 0 -> +re, 1 -> -im, 2 -> +im, 3 -> +re.
// Query mode (rptr == NULL): info[0] = full complex entries, info[1] = real entries.  rval may be NULL.
// Returns 0, -1 invalid CSR, -2 the embedded matrix does not fit int32 indices.
int32_t b200_complex_embed(int32_t n, const int32_t* rp, const int32_t* ci, const double* values, int32_t lower,
                           int64_t* info, int32_t* rptr, int32_t* rcol, int32_t* code, double* rval) {
    if (n < 1 || !rp || !ci || !info || rp[0] != 0 || rp[n] < 1) return -1;
    if (n > (1 << 30) - 1) return -2;
    std::vector<int64_t> fptr((size_t)n + 1, 0);
    for (int32_t i = 0; i < n; i++) {
        if (rp[i + 1] < rp[i]) return -1;
        for (int32_t p = rp[i]; p < rp[i + 1]; p++) {
            const int32_t j = ci[p];
            if (j < 0 || j >= n || (lower && j > i)) return -1;
            if (p > rp[i] && ci[p - 1] >= j) return -1; // sorted, no duplicates
            fptr[i + 1]++;
            if (lower && j != i) fptr[j + 1]++;
        }
    }
    for (int32_t i = 0; i < n; i++) fptr[i + 1] += fptr[i];
    const int64_t nfull = fptr[n];
    info[0] = nfull, info[1] = 4 * nfull;
    if (4 * nfull > 2147483647LL) return -2;
    if (!rptr) return 0;
    if (!rcol || !code) return -1;
    std::vector<int32_t> fcol((size_t)nfull), fsrc((size_t)nfull);
    std::vector<int64_t> fill(fptr.begin(), fptr.end() - 1);
    // row i of the full matrix: its own entries (ascending), then the mirrored ones (j > i; ascending because the source
    // rows are visited in ascending order)
    for (int32_t i = 0; i < n; i++)
        for (int32_t p = rp[i]; p < rp[i + 1]; p++) fcol[fill[i]] = ci[p], fsrc[fill[i]] = p, fill[i]++;
    if (lower)
        for (int32_t i = 0; i < n; i++)
            for (int32_t p = rp[i]; p < rp[i + 1]; p++) {
                const int32_t j = ci[p];
                if (j != i) fcol[fill[j]] = i, fsrc[fill[j]] = p, fill[j]++;
            }
    int64_t w = 0;
    rptr[0] = 0;
    for (int32_t i = 0; i < n; i++)
        for (int half = 0; half < 2; half++) { // half 0: real part of equation i, half 1: imaginary part
            for (int64_t q = fptr[i]; q < fptr[i + 1]; q++) {
                const int32_t j = fcol[q], src = fsrc[q];
                const double re = values ? values[2 * (size_t)src] : 0.0, im = values ? values[2 * (size_t)src + 1] : 0.0;
                rcol[w] = 2 * j, code[w] = 4 * src + (half ? 2 : 0);
                if (rval) rval[w] = half ? im : re;
                w++;
                rcol[w] = 2 * j + 1, code[w] = 4 * src + (half ? 3 : 1);
                if (rval) rval[w] = half ? re : -im;
                w++;
            }
            rptr[2 * i + half + 1] = (int32_t)w;
        }
    return 0;
}

// ---- Matrix Market ----------------------------------------------------------------------------------
// error codes (the Python/Rust side maps them to the reference's messages):
enum {
    MM_OK = 0,
    MM_CANNOT_OPEN = 1,
    MM_EMPTY = 2,
    MM_HDR_START = 3,      // "the header (first line) must start with %%MatrixMarket"
    MM_HDR_NO_KEYWORD = 4, // "cannot find the keyword %%MatrixMarket on the first line"
    MM_HDR_OPT1 = 5,       // first option must be "matrix"
    MM_HDR_OPT1_MISSING = 6,
    MM_HDR_OPT2 = 7, // second option must be "coordinate"
    MM_HDR_OPT2_MISSING = 8,
    MM_HDR_OPT3 = 9, // third option must be "real" or "complex"
    MM_HDR_OPT3_MISSING = 10,
    MM_HDR_OPT4 = 11, // fourth option must be general/symmetric/Hermitian
    MM_HDR_OPT4_MISSING = 12,
    MM_HDR_HERMITIAN_REAL = 13,
    MM_DIM_ROWS = 14, // cannot parse number of rows
    MM_DIM_NO_COLS = 15,
    MM_DIM_COLS = 16,
    MM_DIM_NO_NNZ = 17,
    MM_DIM_NNZ = 18,
    MM_DIM_INVALID = 19, // found invalid (zero or negative) dimensions
    MM_VAL_TOO_MANY = 20,
    MM_VAL_I = 21,
    MM_VAL_NO_J = 22,
    MM_VAL_J = 23,
    MM_VAL_NO_A = 24,
    MM_VAL_A = 25,
    MM_VAL_INDEX = 26,     // found an invalid index
    MM_VAL_MISSING = 27,   // not all values have been found
    MM_SYM_RECT = 28,      // symmetric matrices must be square
    MM_COMPLEX = 29,       // (unused since complex files are read: kept so that the numbering stays stable)
    MM_NO_DIMS = 30,
    MM_VAL_NO_B = 31,      // cannot read bij
    MM_VAL_B = 32,         // cannot parse bij
};

static bool parse_i64(const std::string& s, long long& out) {
    if (s.empty()) return false;
    char* end = nullptr;
    out = strtoll(s.c_str(), &end, 10);
    return end && *end == '\0';
}
static bool parse_f64(const std::string& s, double& out) {
    if (s.empty()) return false;
    char* end = nullptr;
    out = strtod(s.c_str(), &end);
    return end && *end == '\0';
}
static std::vector<std::string> split_ws(const std::string& line) {
    std::vector<std::string> t;
    std::istringstream is(line);
    std::string w;
    while (is >> w) t.push_back(w);
    return t;
}

static int parse_header(const std::string& line, bool& is_complex, bool& is_sym) {
    std::vector<std::string> t = split_ws(line);
    if (t.empty()) return MM_HDR_NO_KEYWORD;
    if (t[0] != "%%MatrixMarket") return MM_HDR_START;
    if (t.size() < 2) return MM_HDR_OPT1_MISSING;
    if (t[1] != "matrix") return MM_HDR_OPT1;
    if (t.size() < 3) return MM_HDR_OPT2_MISSING;
    if (t[2] != "coordinate") return MM_HDR_OPT2;
    if (t.size() < 4) return MM_HDR_OPT3_MISSING;
    if (t[3] == "real") is_complex = false;
    else if (t[3] == "complex") is_complex = true;
    else return MM_HDR_OPT3;
    if (t.size() < 5) return MM_HDR_OPT4_MISSING;
    if (t[4] == "general") is_sym = false;
    else if (t[4] == "symmetric") is_sym = true;
    else if (t[4] == "Hermitian") {
        if (!is_complex) return MM_HDR_HERMITIAN_REAL;
        is_sym = false;
    } else return MM_HDR_OPT4;
    return MM_OK;
}

// Reads a coordinate file, real or complex.  handling: 0 LeaveAsLower, 1 SwapToUpper, 2 MakeItFull (enums.rs:45-67).
// Pass ai=aj=NULL to query: info[0..4] = nrow, ncol, nnz_file, symmetric(0/1), max_entries (capacity to allocate),
// info[6] = 1 for a complex file (values are then interleaved (re, im) pairs: ax needs 2 * capacity doubles).
// With arrays: fills triplets, info[5] = number of triplets written.  info must hold 8 entries.
int32_t b200_mm_read(const char* path, int32_t handling, int64_t* info, int32_t* ai, int32_t* aj, double* ax, int64_t cap) {
    std::ifstream in(path);
    if (!in.good()) return MM_CANNOT_OPEN;
    std::string line;
    if (!std::getline(in, line)) return MM_EMPTY;
    bool is_complex = false, is_sym = false;
    int rc = parse_header(line, is_complex, is_sym);
    if (rc != MM_OK) return rc;
    long long m = 0, n = 0, nnz = 0;
    bool have_dims = false;
    while (std::getline(in, line)) {
        std::vector<std::string> t = split_ws(line);
        if (t.empty() || t[0][0] == '%') continue;
        if (!parse_i64(t[0], m)) return MM_DIM_ROWS;
        if (t.size() < 2) return MM_DIM_NO_COLS;
        if (!parse_i64(t[1], n)) return MM_DIM_COLS;
        if (t.size() < 3) return MM_DIM_NO_NNZ;
        if (!parse_i64(t[2], nnz)) return MM_DIM_NNZ;
        if (m < 1 || n < 1 || nnz < 1) return MM_DIM_INVALID;
        have_dims = true;
        break;
    }
    if (!have_dims) return MM_NO_DIMS;
    if (is_sym && m != n) return MM_SYM_RECT;
    const int W = is_complex ? 2 : 1;
    long long maxent = (is_sym && handling == 2) ? 2 * nnz : nnz;
    info[0] = m, info[1] = n, info[2] = nnz, info[3] = is_sym ? 1 : 0, info[4] = maxent, info[6] = is_complex ? 1 : 0;
    if (!ai || !aj || !ax) return MM_OK;
    long long pos = 0, w = 0;
    while (std::getline(in, line)) {
        std::vector<std::string> t = split_ws(line);
        if (t.empty() || t[0][0] == '%') continue;
        if (pos == nnz) return MM_VAL_TOO_MANY;
        long long i, j;
        double a, bim = 0.0;
        if (!parse_i64(t[0], i)) return MM_VAL_I;
        if (t.size() < 2) return MM_VAL_NO_J;
        if (!parse_i64(t[1], j)) return MM_VAL_J;
        if (t.size() < 3) return MM_VAL_NO_A;
        if (!parse_f64(t[2], a)) return MM_VAL_A;
        if (is_complex) {
            if (t.size() < 4) return MM_VAL_NO_B;
            if (!parse_f64(t[3], bim)) return MM_VAL_B;
        }
        i -= 1, j -= 1;
        if (i < 0 || i >= m || j < 0 || j >= n) return MM_VAL_INDEX;
        pos++;
        if (w + 2 > cap && w + 1 > cap) return MM_VAL_TOO_MANY;
        auto put = [&](long long r, long long c) {
            ai[w] = (int32_t)r, aj[w] = (int32_t)c, ax[W * w] = a;
            if (W == 2) ax[W * w + 1] = bim;
            w++;
        };
        if (is_sym && handling == 1) {
            put(j, i);
        } else {
            put(i, j);
            if (is_sym && handling == 2 && i != j) put(j, i);
        }
    }
    if (pos != nnz) return MM_VAL_MISSING;
    info[5] = w;
    return MM_OK;
}

} // extern "C"
