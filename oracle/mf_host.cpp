// mf_host.cpp -- TEST INFRASTRUCTURE ONLY (oracle/): a plain scalar CPU walk through the SAME front plan the
// CUDA kernels execute (russell_b200/csrc/plan.hpp).  It exists so that
//   (1) the host analysis (ordering, supernodes, relative indices, scatter map) can be validated in the
//       GPU-less authoring container against SuperLU / the reference's known-answer tests, and
//   (2) the device kernels can be compared array-by-array (factor panels, pivots, solution) on the GPU box.
// The product (libsolver_b200.so) never links or calls this file; it fails loudly without a CUDA device.
//
// The arithmetic restated here is the standard multifrontal LU the reference reaches through
// umfpack_di_numeric / umfpack_di_solve (russell_sparse/c_code/interface_umfpack.c:167,229), specialised to a
// symmetric-pattern front tree with pivoting restricted to each front's pivot block.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../russell_b200/csrc/plan.hpp"

using namespace b200;

namespace {

struct HostSolver {
    Plan P;
    std::vector<double> fac, cb, dinv;
    std::vector<int> piv; // piv[c0+k] = local row swapped with k at step k
    int n_perturbed = 0, n_zero_pivot = 0;
    bool singular = false;
    double anorm = 0;
};

void host_factorize(HostSolver& S, const double* vals, double pivot_eps) {
    const Plan& P = S.P;
    S.fac.assign(P.fac_size, 0.0);
    S.cb.assign(P.cb_size, 0.0);
    S.dinv.assign(P.dinv_size, 0.0);
    S.piv.assign(P.n, 0);
    S.n_perturbed = S.n_zero_pivot = 0;
    S.singular = false;
    double amax = 0;
    for (size_t k = 0; k < P.a_dst.size(); k++) {
        double v = vals[P.a_src[k]] * (P.a_scl.empty() ? 1.0 : P.a_scl[k]);
        S.fac[P.a_dst[k]] = v;
        amax = std::max(amax, std::fabs(v));
    }
    S.anorm = amax;
    const double tiny = pivot_eps * (amax > 0 ? amax : 1.0);
    for (int v = 0; v < P.nnodes; v++) { // postorder == a valid level order
        const int p = P.p[v], u = P.u[v], f = p + u;
        double* L = &S.fac[P.Loff[v]];
        double* U = u ? &S.fac[P.Uoff[v]] : nullptr;
        double* C = u ? &S.cb[P.Coff[v]] : nullptr;
        // extend-add of the children
        for (int e = P.child_ptr[v]; e < P.child_ptr[v + 1]; e++) {
            const int c = P.child_idx[e];
            const int uc = P.u[c];
            const double* Cc = &S.cb[P.Coff[c]];
            const int* rel = &P.rel[P.rows_ptr[c]];
            for (int j = 0; j < uc; j++) {
                const int tj = rel[j];
                for (int i = 0; i < uc; i++) {
                    const int ti = rel[i];
                    const double val = Cc[i + (size_t)j * uc];
                    if (tj < p) L[ti + (size_t)tj * f] += val;
                    else if (ti < p) U[(tj - p) + (size_t)ti * u] += val;
                    else C[(ti - p) + (size_t)(tj - p) * u] += val;
                }
            }
        }
        // LU of the pivot block, partial pivoting restricted to the block
        int* piv = &S.piv[P.c0[v]];
        for (int k = 0; k < p; k++) {
            int r = k;
            double best = std::fabs(L[k + (size_t)k * f]);
            for (int i = k + 1; i < p; i++) {
                double a = std::fabs(L[i + (size_t)k * f]);
                if (a > best) best = a, r = i;
            }
            piv[k] = r;
            if (r != k) {
                for (int j = 0; j < p; j++) std::swap(L[k + (size_t)j * f], L[r + (size_t)j * f]);
                for (int j = 0; j < u; j++) std::swap(U[j + (size_t)k * u], U[j + (size_t)r * u]);
            }
            double d = L[k + (size_t)k * f];
            if (!(std::fabs(d) >= tiny)) {
                if (d == 0.0 || d != d) {
                    S.n_zero_pivot++;
                    if (u == 0) S.singular = true;
                }
                S.n_perturbed++;
                d = (d < 0 ? -tiny : tiny);
                if (d == 0.0) d = 1e-300;
                L[k + (size_t)k * f] = d;
            }
            const double inv = 1.0 / d;
            for (int i = k + 1; i < p; i++) L[i + (size_t)k * f] *= inv;
            for (int j = k + 1; j < p; j++) {
                const double ukj = L[k + (size_t)j * f];
                if (ukj != 0.0)
                    for (int i = k + 1; i < p; i++) L[i + (size_t)j * f] -= L[i + (size_t)k * f] * ukj;
            }
        }
        // explicit inverses of the triangular pivot factors
        double* D = &S.dinv[P.Doff[v]];
        for (int j = 0; j < p; j++) { // column j of inv(L11) (unit lower): solve L x = e_j
            for (int i = j + 1; i < p; i++) {
                double s = -L[i + (size_t)j * f];
                for (int m = j + 1; m < i; m++) s -= L[i + (size_t)m * f] * D[m + (size_t)j * p];
                D[i + (size_t)j * p] = s;
            }
        }
        // inv(U11) by back substitution, column by column: solve U x = e_j (x has entries 0..j)
        {
            std::vector<double> X((size_t)p * p, 0.0);
            for (int j = 0; j < p; j++) {
                for (int i = j; i >= 0; i--) {
                    double s = (i == j) ? 1.0 : 0.0;
                    for (int m = i + 1; m <= j; m++) s -= L[i + (size_t)m * f] * X[m + (size_t)j * p];
                    X[i + (size_t)j * p] = s / L[i + (size_t)i * f];
                }
            }
            for (int j = 0; j < p; j++)
                for (int i = 0; i <= j; i++) D[i + (size_t)j * p] = X[i + (size_t)j * p];
        }
        if (u == 0) continue;
        // panels: L21 <- F21 * inv(U11) ;  Upanel <- Upanel * inv(L11)^T
        std::vector<double> row(p);
        for (int i = p; i < f; i++) {
            for (int j = 0; j < p; j++) {
                double s = 0;
                for (int k = 0; k <= j; k++) s += L[i + (size_t)k * f] * D[k + (size_t)j * p];
                row[j] = s;
            }
            for (int j = 0; j < p; j++) L[i + (size_t)j * f] = row[j];
        }
        for (int j = 0; j < u; j++) {
            for (int k = 0; k < p; k++) {
                double s = U[j + (size_t)k * u];
                for (int m = 0; m < k; m++) s += D[k + (size_t)m * p] * U[j + (size_t)m * u];
                row[k] = s;
            }
            for (int k = 0; k < p; k++) U[j + (size_t)k * u] = row[k];
        }
        // Schur complement
        for (int j = 0; j < u; j++)
            for (int k = 0; k < p; k++) {
                const double ujk = U[j + (size_t)k * u];
                if (ujk != 0.0)
                    for (int i = 0; i < u; i++) C[i + (size_t)j * u] -= L[(p + i) + (size_t)k * f] * ujk;
            }
    }
}

// solves A'' xp = y in the permuted/scaled space; y is overwritten by the forward result, xp receives the answer
void host_solve_permuted(const HostSolver& S, std::vector<double>& y, std::vector<double>& xp) {
    const Plan& P = S.P;
    std::vector<double> wv(P.rows_ptr[P.nnodes] + 1, 0.0), t, z;
    for (int v = 0; v < P.nnodes; v++) {
        const int p = P.p[v], u = P.u[v], f = p + u, c0 = P.c0[v];
        const double* L = &S.fac[P.Loff[v]];
        const double* D = &S.dinv[P.Doff[v]];
        t.assign(f, 0.0);
        for (int k = 0; k < p; k++) t[k] = y[c0 + k];
        for (int e = P.child_ptr[v]; e < P.child_ptr[v + 1]; e++) {
            const int c = P.child_idx[e];
            const int* rel = &P.rel[P.rows_ptr[c]];
            const double* wc = &wv[P.rows_ptr[c]];
            for (int i = 0; i < P.u[c]; i++) t[rel[i]] += wc[i];
        }
        const int* piv = &S.piv[c0];
        for (int k = 0; k < p; k++)
            if (piv[k] != k) std::swap(t[k], t[piv[k]]);
        z.assign(p, 0.0);
        for (int k = 0; k < p; k++) {
            double s = t[k];
            for (int m = 0; m < k; m++) s += D[k + (size_t)m * p] * t[m];
            z[k] = s;
        }
        for (int k = 0; k < p; k++) y[c0 + k] = z[k];
        double* w = &wv[P.rows_ptr[v]];
        for (int i = 0; i < u; i++) {
            double s = t[p + i];
            for (int k = 0; k < p; k++) s -= L[(p + i) + (size_t)k * f] * z[k];
            w[i] = s;
        }
    }
    xp.assign(P.n, 0.0);
    for (int v = P.nnodes - 1; v >= 0; v--) {
        const int p = P.p[v], u = P.u[v], c0 = P.c0[v];
        const double* U = u ? &S.fac[P.Uoff[v]] : nullptr;
        const double* D = &S.dinv[P.Doff[v]];
        const int* rows = &P.rows[P.rows_ptr[v]];
        t.assign(p, 0.0);
        for (int k = 0; k < p; k++) {
            double s = y[c0 + k];
            for (int j = 0; j < u; j++) s -= U[j + (size_t)k * u] * xp[rows[j]];
            t[k] = s;
        }
        for (int k = 0; k < p; k++) {
            double s = 0;
            for (int m = k; m < p; m++) s += D[k + (size_t)m * p] * t[m];
            xp[c0 + k] = s;
        }
    }
}

void host_spmv_full(const Plan& P, const double* vals, const double* x, double* yv) {
    for (int i = 0; i < P.n; i++) {
        double s = 0;
        for (int k = P.full_ptr[i]; k < P.full_ptr[i + 1]; k++)
            s += vals[P.full_src.empty() ? k : P.full_src[k]] * x[P.full_col[k]];
        yv[i] = s;
    }
}

void host_solve(const HostSolver& S, const double* vals, const double* b, double* x, int nrefine, double* resid_out) {
    const Plan& P = S.P;
    const int n = P.n;
    auto apply = [&](const double* rhs, double* sol) {
        std::vector<double> y(n), xp;
        for (int k = 0; k < n; k++) {
            int r = P.rowperm[k];
            y[k] = rhs[r] * (P.rscale.empty() ? 1.0 : P.rscale[r]);
        }
        host_solve_permuted(S, y, xp);
        for (int k = 0; k < n; k++) {
            int c = P.colperm[k];
            sol[c] = xp[k] * (P.cscale.empty() ? 1.0 : P.cscale[c]);
        }
    };
    apply(b, x);
    std::vector<double> r(n), d(n), ax(n);
    double bn = 0;
    for (int i = 0; i < n; i++) bn += b[i] * b[i];
    bn = std::sqrt(bn);
    double rn = 0;
    for (int it = 0; it <= nrefine; it++) {
        host_spmv_full(P, vals, x, ax.data());
        rn = 0;
        for (int i = 0; i < n; i++) r[i] = b[i] - ax[i], rn += r[i] * r[i];
        rn = std::sqrt(rn);
        if (it == nrefine || rn <= 1e-15 * bn) break;
        apply(r.data(), d.data());
        for (int i = 0; i < n; i++) x[i] += d[i];
    }
    if (resid_out) *resid_out = bn > 0 ? rn / bn : rn;
}

} // namespace

extern "C" {

// One-shot: analyze + factorize + solve on the host.  Returns 0, 1 (singular), or <0 (analysis failure).
// stats[0..9] = nnodes, nlevels, nnz_L, nnz_U, flops, n_perturbed, rel.residual, max_front, t_order, t_symbolic
int oracle_mf_solve(int n, const int* rowptr, const int* colidx, const double* vals, int sym_lower, int ordering,
                    int matching, int panel_width, int nd_leaf, int nrefine, double pivot_eps, const double* b,
                    double* x, double* stats, int verbose) {
    HostSolver S;
    AnalyzeOptions opt;
    opt.ordering = ordering;
    opt.matching = matching;
    if (panel_width > 0) opt.panel_width = panel_width;
    if (nd_leaf > 0) opt.nd_leaf = nd_leaf;
    opt.verbose = verbose;
    opt.cb_reuse = false; // this walk runs in postorder and accumulates into zero-initialised private blocks
    int rc = analyze(n, rowptr, colidx, vals, sym_lower != 0, opt, S.P);
    if (rc != 0) return rc;
    host_factorize(S, vals, pivot_eps > 0 ? pivot_eps : 1e-13);
    double resid = 0;
    host_solve(S, vals, b, x, nrefine, &resid);
    if (stats) {
        stats[0] = S.P.nnodes, stats[1] = S.P.nlevels, stats[2] = (double)S.P.nnz_L, stats[3] = (double)S.P.nnz_U;
        stats[4] = S.P.flops, stats[5] = S.n_perturbed, stats[6] = resid, stats[7] = S.P.max_front;
        stats[8] = S.P.t_order, stats[9] = S.P.t_symbolic;
    }
    return S.singular ? 1 : 0;
}

// analysis only: fills stats like above (no numeric work) -- used to size-check large problems quickly
int oracle_mf_analyze(int n, const int* rowptr, const int* colidx, const double* vals, int sym_lower, int ordering,
                      int matching, int panel_width, int nd_leaf, double* stats, int verbose) {
    Plan P;
    AnalyzeOptions opt;
    opt.ordering = ordering;
    opt.matching = matching;
    if (panel_width > 0) opt.panel_width = panel_width;
    if (nd_leaf > 0) opt.nd_leaf = nd_leaf;
    opt.verbose = verbose;
    int rc = analyze(n, rowptr, colidx, vals, sym_lower != 0, opt, P);
    if (rc != 0) return rc;
    stats[0] = P.nnodes, stats[1] = P.nlevels, stats[2] = (double)P.nnz_L, stats[3] = (double)P.nnz_U;
    stats[4] = P.flops, stats[5] = (double)P.cb_size, stats[6] = (double)P.fac_size, stats[7] = P.max_front;
    stats[8] = P.t_order, stats[9] = P.t_symbolic;
    return 0;
}

// ---- handle API (used by the GPU parity tests to compare factor panels array-by-array) --------------------
void* oracle_mf_create(int n, const int* rowptr, const int* colidx, const double* vals, int sym_lower, int ordering,
                       int matching, int panel_width, int nd_leaf, double pivot_eps, int* status) {
    HostSolver* S = new HostSolver();
    AnalyzeOptions opt;
    opt.ordering = ordering;
    opt.matching = matching;
    if (panel_width > 0) opt.panel_width = panel_width;
    if (nd_leaf > 0) opt.nd_leaf = nd_leaf;
    opt.cb_reuse = false;
    int rc = analyze(n, rowptr, colidx, vals, sym_lower != 0, opt, S->P);
    if (rc != 0) {
        *status = rc;
        delete S;
        return nullptr;
    }
    host_factorize(*S, vals, pivot_eps > 0 ? pivot_eps : 1e-13);
    *status = S->singular ? 1 : 0;
    return S;
}
void oracle_mf_sizes(void* h, int64_t* out) { // fac, dinv, n, nnodes, perturbed
    HostSolver* S = (HostSolver*)h;
    out[0] = S->P.fac_size, out[1] = S->P.dinv_size, out[2] = S->P.n, out[3] = S->P.nnodes, out[4] = S->n_perturbed;
}
// lperm: composed local permutation per front (row at position k after pivoting = original local row lperm[k])
void oracle_mf_get(void* h, double* fac, double* dinv, int* lperm) {
    HostSolver* S = (HostSolver*)h;
    if (fac) std::copy(S->fac.begin(), S->fac.end(), fac);
    if (dinv) std::copy(S->dinv.begin(), S->dinv.end(), dinv);
    if (lperm) {
        const Plan& P = S->P;
        for (int v = 0; v < P.nnodes; v++) {
            const int p = P.p[v], c0 = P.c0[v];
            for (int k = 0; k < p; k++) lperm[c0 + k] = k;
            for (int k = 0; k < p; k++) std::swap(lperm[c0 + k], lperm[c0 + S->piv[c0 + k]]);
        }
    }
}
void oracle_mf_handle_solve(void* h, const double* vals, const double* b, double* x, int nrefine, double* resid) {
    host_solve(*(HostSolver*)h, vals, b, x, nrefine, resid);
}
void oracle_mf_free(void* h) { delete (HostSolver*)h; }

// analysis-only handle: node shapes for studying the front tree (p, u, level, parent per node)
void* oracle_plan_create(int n, const int* rowptr, const int* colidx, const double* vals, int sym_lower, int ordering,
                         int matching, int panel_width, int nd_leaf, int* nnodes) {
    Plan* P = new Plan();
    AnalyzeOptions opt;
    opt.ordering = ordering;
    opt.matching = matching;
    if (panel_width > 0) opt.panel_width = panel_width;
    if (nd_leaf > 0) opt.nd_leaf = nd_leaf;
    if (analyze(n, rowptr, colidx, vals, sym_lower != 0, opt, *P) != 0) {
        delete P;
        return nullptr;
    }
    *nnodes = P->nnodes;
    return P;
}
void oracle_plan_nodes(void* h, int* p, int* u, int* level, int* parent) {
    Plan* P = (Plan*)h;
    for (int v = 0; v < P->nnodes; v++) p[v] = P->p[v], u[v] = P->u[v], level[v] = P->level[v], parent[v] = P->parent[v];
}
// FNV-1a over every array of the plan the device consumes: two runs of the analysis (serial / threaded) must agree bit for bit
unsigned long long oracle_plan_hash(void* h) {
    const Plan& P = *(Plan*)h;
    unsigned long long x = 1469598103934665603ull;
    auto fnv = [&x](const void* data, size_t bytes) {
        const unsigned char* b = (const unsigned char*)data;
        for (size_t i = 0; i < bytes; i++) x = (x ^ b[i]) * 1099511628211ull;
    };
#define HV(v) fnv((v).data(), (v).size() * sizeof((v)[0]))
    HV(P.rowperm), HV(P.colperm), HV(P.c0), HV(P.p), HV(P.u), HV(P.parent), HV(P.level), HV(P.Loff), HV(P.Uoff), HV(P.Coff), HV(P.Doff);
    HV(P.rows_ptr), HV(P.rows), HV(P.rel), HV(P.child_ptr), HV(P.child_idx), HV(P.level_ptr), HV(P.level_nodes), HV(P.in_sub);
    HV(P.st_first), HV(P.st_root), HV(P.a_src), HV(P.a_dst), HV(P.a_scl), HV(P.full_ptr), HV(P.full_col), HV(P.full_src), HV(P.rscale), HV(P.cscale);
#undef HV
    fnv(&P.fac_size, sizeof(P.fac_size)), fnv(&P.cb_size, sizeof(P.cb_size)), fnv(&P.dinv_size, sizeof(P.dinv_size));
    return x;
}
void oracle_plan_free(void* h) { delete (Plan*)h; }
}
