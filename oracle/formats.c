/* formats.c -- TEST INFRASTRUCTURE ONLY (oracle/): plain-C restatement of the reference's host formats, used as
 * the checker for the product's converters and for the CUDA SpMV / residual kernels.
 *
 *   oracle_coo_to_csr   restates CsrMatrix::update_from_coo  (russell_sparse/src/csr_matrix.rs:359-480):
 *                       bucket triplets by row, sum duplicates in order of appearance, sort each row by column.
 *   oracle_csr_matvec   restates CsrMatrix::mat_vec_mul      (russell_sparse/src/csr_matrix.rs:709-729),
 *                       including the mirrored update for triangular (Sym::YesLower / YesUpper) storage.
 *   oracle_coo_matvec   restates CooMatrix::mat_vec_mul      (russell_sparse/src/coo_matrix.rs:547-565).
 *   oracle_verify       restates VerifyLinSys::from          (russell_sparse/src/verify_lin_sys.rs:60-96).
 * Pinned against the reference's own fixtures (tests/golden/samples.json) by tests/test_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int32_t oracle_coo_to_csr(int32_t nrow, int32_t ncol, int32_t nnz, const int32_t *ai, const int32_t *aj, const double *ax,
                          int32_t *bp, int32_t *bj, double *bx) {
    if (nnz < 1) return -2;
    int32_t *cnt = (int32_t *)calloc((size_t)nrow + 1, sizeof(int32_t));
    int32_t *rj = (int32_t *)malloc((size_t)nnz * sizeof(int32_t));
    double *rx = (double *)malloc((size_t)nnz * sizeof(double));
    int32_t *pos = (int32_t *)malloc((size_t)(ncol > 0 ? ncol : 1) * sizeof(int32_t));
    if (!cnt || !rj || !rx || !pos) return -3;
    for (int32_t k = 0; k < nnz; k++) cnt[ai[k] + 1]++;
    for (int32_t i = 0; i < nrow; i++) cnt[i + 1] += cnt[i];
    int32_t *w = (int32_t *)malloc((size_t)(nrow > 0 ? nrow : 1) * sizeof(int32_t));
    memcpy(w, cnt, (size_t)nrow * sizeof(int32_t));
    for (int32_t k = 0; k < nnz; k++) {
        int32_t p = w[ai[k]]++;
        rj[p] = aj[k];
        rx[p] = ax[k];
    }
    for (int32_t j = 0; j < ncol; j++) pos[j] = -1;
    int32_t out = 0;
    bp[0] = 0;
    for (int32_t i = 0; i < nrow; i++) {
        int32_t start = out;
        for (int32_t p = cnt[i]; p < cnt[i + 1]; p++) {
            int32_t j = rj[p];
            if (pos[j] >= start) {
                bx[pos[j]] += rx[p];
            } else {
                pos[j] = out;
                bj[out] = j;
                bx[out] = rx[p];
                out++;
            }
        }
        /* insertion sort of the (short) row by column */
        for (int32_t a = start + 1; a < out; a++) {
            int32_t cj = bj[a];
            double cx = bx[a];
            int32_t b = a - 1;
            while (b >= start && bj[b] > cj) {
                bj[b + 1] = bj[b];
                bx[b + 1] = bx[b];
                b--;
            }
            bj[b + 1] = cj;
            bx[b + 1] = cx;
        }
        for (int32_t a = start; a < out; a++) pos[bj[a]] = -1;
        bp[i + 1] = out;
    }
    free(cnt), free(rj), free(rx), free(pos), free(w);
    return 0;
}

/* v = alpha * A * u for CSR; mirror != 0 when only one triangle is stored */
void oracle_csr_matvec(int32_t nrow, const int32_t *bp, const int32_t *bj, const double *bx, int32_t mirror, double alpha,
                       const double *u, double *v) {
    for (int32_t i = 0; i < nrow; i++) v[i] = 0.0;
    for (int32_t i = 0; i < nrow; i++)
        for (int32_t p = bp[i]; p < bp[i + 1]; p++) {
            int32_t j = bj[p];
            v[i] += alpha * bx[p] * u[j];
            if (mirror && i != j) v[j] += alpha * bx[p] * u[i];
        }
}

void oracle_coo_matvec(int32_t nrow, int32_t nnz, const int32_t *ai, const int32_t *aj, const double *ax, int32_t mirror,
                       double alpha, const double *u, double *v) {
    for (int32_t i = 0; i < nrow; i++) v[i] = 0.0;
    for (int32_t p = 0; p < nnz; p++) {
        int32_t i = ai[p], j = aj[p];
        v[i] += alpha * ax[p] * u[j];
        if (mirror && i != j) v[j] += alpha * ax[p] * u[i];
    }
}

/* out[0..3] = max_abs_a, max_abs_ax, max_abs_diff, relative_error */
void oracle_verify(int32_t nrow, int32_t nnz, const int32_t *ai, const int32_t *aj, const double *ax, int32_t mirror,
                   const double *x, const double *rhs, double *out) {
    double *v = (double *)malloc((size_t)nrow * sizeof(double));
    double maxa = 0.0, maxax = 0.0, maxd = 0.0;
    for (int32_t p = 0; p < nnz; p++)
        if (fabs(ax[p]) > maxa) maxa = fabs(ax[p]);
    oracle_coo_matvec(nrow, nnz, ai, aj, ax, mirror, 1.0, x, v);
    for (int32_t i = 0; i < nrow; i++) {
        if (fabs(v[i]) > maxax) maxax = fabs(v[i]);
        if (fabs(v[i] - rhs[i]) > maxd) maxd = fabs(v[i] - rhs[i]);
    }
    out[0] = maxa, out[1] = maxax, out[2] = maxd, out[3] = maxd / (maxa + 1.0);
    free(v);
}

/* ---- Complex64 twins (values are interleaved (re, im) pairs) --------------------------------------------------------
 *   oracle_complex_coo_to_csr  restates ComplexCsrMatrix::update_from_coo (the same generic code over Complex64,
 *                              russell_sparse/src/csr_matrix.rs:359-480; alias russell_sparse/src/aliases.rs)
 *   oracle_complex_coo_matvec  restates ComplexCooMatrix::mat_vec_mul     (russell_sparse/src/coo_matrix.rs:547-565)
 *   oracle_complex_verify      restates VerifyLinSys::from_complex        (russell_sparse/src/verify_lin_sys.rs:104-146)
 * Pinned against tests/golden/complex_samples.json and the complex known answers by tests/test_complex_cpu.py. */
int32_t oracle_complex_coo_to_csr(int32_t nrow, int32_t ncol, int32_t nnz, const int32_t *ai, const int32_t *aj,
                                  const double *ax, int32_t *bp, int32_t *bj, double *bx) {
    if (nnz < 1) return -2;
    /* first-seen slot per (row, column), accumulate in order of appearance, then order every row by column */
    int32_t *cnt = (int32_t *)calloc((size_t)nrow + 2, sizeof(int32_t));
    int32_t *ord = (int32_t *)malloc((size_t)nnz * sizeof(int32_t));
    if (!cnt || !ord) return -3;
    for (int32_t k = 0; k < nnz; k++) cnt[ai[k] + 2]++;
    for (int32_t i = 0; i < nrow; i++) cnt[i + 2] += cnt[i + 1];
    for (int32_t k = 0; k < nnz; k++) ord[cnt[ai[k] + 1]++] = k; /* cnt[i+1] ends at the end of row i, cnt[i] is its start */
    int32_t out = 0;
    bp[0] = 0;
    for (int32_t i = 0; i < nrow; i++) {
        int32_t start = out;
        for (int32_t p = cnt[i]; p < cnt[i + 1]; p++) {
            int32_t k = ord[p], hit = -1;
            for (int32_t a = start; a < out; a++)
                if (bj[a] == aj[k]) hit = a;
            if (hit < 0) hit = out++, bj[hit] = aj[k], bx[2 * hit] = 0.0, bx[2 * hit + 1] = 0.0;
            bx[2 * hit] += ax[2 * k], bx[2 * hit + 1] += ax[2 * k + 1];
        }
        for (int32_t a = start + 1; a < out; a++) { /* insertion sort by column */
            int32_t cj = bj[a], b = a - 1;
            double cr = bx[2 * a], cim = bx[2 * a + 1];
            while (b >= start && bj[b] > cj) bj[b + 1] = bj[b], bx[2 * b + 2] = bx[2 * b], bx[2 * b + 3] = bx[2 * b + 1], b--;
            bj[b + 1] = cj, bx[2 * b + 2] = cr, bx[2 * b + 3] = cim;
        }
        bp[i + 1] = out;
    }
    (void)ncol;
    free(cnt), free(ord);
    return 0;
}

/* v = A u (complex), mirror != 0 when one triangle of a complex SYMMETRIC matrix is stored */
void oracle_complex_coo_matvec(int32_t nrow, int32_t nnz, const int32_t *ai, const int32_t *aj, const double *ax,
                               int32_t mirror, const double *u, double *v) {
    for (int32_t i = 0; i < 2 * nrow; i++) v[i] = 0.0;
    for (int32_t p = 0; p < nnz; p++) {
        int32_t i = ai[p], j = aj[p];
        double ar = ax[2 * p], am = ax[2 * p + 1];
        v[2 * i] += ar * u[2 * j] - am * u[2 * j + 1];
        v[2 * i + 1] += ar * u[2 * j + 1] + am * u[2 * j];
        if (mirror && i != j) {
            v[2 * j] += ar * u[2 * i] - am * u[2 * i + 1];
            v[2 * j + 1] += ar * u[2 * i + 1] + am * u[2 * i];
        }
    }
}

void oracle_complex_verify(int32_t nrow, int32_t nnz, const int32_t *ai, const int32_t *aj, const double *ax, int32_t mirror,
                           const double *x, const double *rhs, double *out) {
    double *v = (double *)malloc((size_t)2 * nrow * sizeof(double));
    double maxa = 0.0, maxax = 0.0, maxd = 0.0;
    for (int32_t p = 0; p < nnz; p++) {
        double a = hypot(ax[2 * p], ax[2 * p + 1]);
        if (a > maxa) maxa = a;
    }
    oracle_complex_coo_matvec(nrow, nnz, ai, aj, ax, mirror, x, v);
    for (int32_t i = 0; i < nrow; i++) {
        double a = hypot(v[2 * i], v[2 * i + 1]), d = hypot(v[2 * i] - rhs[2 * i], v[2 * i + 1] - rhs[2 * i + 1]);
        if (a > maxax) maxax = a;
        if (d > maxd) maxd = d;
    }
    out[0] = maxa, out[1] = maxax, out[2] = maxd, out[3] = maxd / (maxa + 1.0);
    free(v);
}
