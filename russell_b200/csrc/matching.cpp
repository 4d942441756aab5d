// matching.cpp -- maximum-product bipartite matching with row/column scaling (Duff & Koster, MC64 job 5 idea),
// written from scratch as a sparse shortest-augmenting-path (Hungarian) method on the costs
//     c(i,j) = log(max_i |a_ij|) - log |a_ij|   >= 0.
//
// Why it exists: the GPU factorization pivots only inside a front's pivot block (static structure), so zero
// or tiny diagonals must be removed beforehand.  The reference documents exactly this weakness for cuDSS on
// Samples::umfpack_unsymmetric_5x5 (russell_sparse/src/solver_cudss.rs:664-671: error 1.2e-3 without matching)
// and fixes it with Matching::Auto (solver_cudss.rs:735-755).  UMFPACK itself does not need it because it
// pivots across the whole front.
//
// Output: rowmatch[j] = row matched to column j;  rscale/cscale such that the scaled, row-permuted matrix has
// |diagonal| = 1 and |off-diagonal| <= 1.
#include "plan.hpp"

#include <cmath>
#include <limits>
#include <queue>

namespace b200 {

static int matching_impl(int n, const int* rowptr, const int* colidx, const double* vals_in, bool zeros_as_tiny,
                         std::vector<int>& rowmatch, std::vector<double>& rscale, std::vector<double>& cscale) {
    const double INF = std::numeric_limits<double>::infinity();
    const int nnz = rowptr[n];
    // structural mode: stored entries that are exactly zero stay usable, at a prohibitive (but finite) cost
    std::vector<double> tmp;
    const double* vals = vals_in;
    if (zeros_as_tiny) {
        tmp.assign(vals_in, vals_in + nnz);
        for (int k = 0; k < nnz; k++)
            if (tmp[k] == 0.0) tmp[k] = 1e-150;
        vals = tmp.data();
    }
    // CSC copy with costs
    std::vector<int> cp(n + 1, 0), ri(nnz);
    std::vector<double> cost(nnz);
    std::vector<double> colmax(n, 0.0);
    for (int i = 0; i < n; i++)
        for (int k = rowptr[i]; k < rowptr[i + 1]; k++) {
            double a = std::fabs(vals[k]);
            if (a > 0.0 && std::isfinite(a)) {
                cp[colidx[k] + 1]++;
                if (a > colmax[colidx[k]]) colmax[colidx[k]] = a;
            }
        }
    for (int j = 0; j < n; j++) cp[j + 1] += cp[j];
    {
        std::vector<int> fill(cp.begin(), cp.end() - 1);
        for (int i = 0; i < n; i++)
            for (int k = rowptr[i]; k < rowptr[i + 1]; k++) {
                double a = std::fabs(vals[k]);
                if (a > 0.0 && std::isfinite(a)) {
                    int j = colidx[k];
                    int d = fill[j]++;
                    ri[d] = i;
                    cost[d] = std::log(colmax[j]) - std::log(a);
                }
            }
    }
    std::vector<double> u(n, INF), v(n, INF);
    // initial duals: u[i] = min_j c(i,j);  v[j] = min_i (c(i,j) - u[i])
    for (int j = 0; j < n; j++)
        for (int k = cp[j]; k < cp[j + 1]; k++)
            if (cost[k] < u[ri[k]]) u[ri[k]] = cost[k];
    for (int i = 0; i < n; i++)
        if (u[i] == INF) u[i] = 0.0; // empty row: will fail to match below
    for (int j = 0; j < n; j++) {
        for (int k = cp[j]; k < cp[j + 1]; k++) {
            double r = cost[k] - u[ri[k]];
            if (r < v[j]) v[j] = r;
        }
        if (v[j] == INF) v[j] = 0.0;
    }
    std::vector<int> mrow(n, -1); // mrow[j] = row matched to column j
    std::vector<int> mcol(n, -1); // mcol[i] = column matched to row i
    std::vector<double> mcost(n, 0.0); // cost of the matched edge of row i
    // greedy start on tight edges
    int matched = 0;
    for (int j = 0; j < n; j++)
        for (int k = cp[j]; k < cp[j + 1]; k++) {
            int i = ri[k];
            if (mcol[i] < 0 && cost[k] - u[i] - v[j] <= 1e-14) {
                mrow[j] = i, mcol[i] = j, mcost[i] = cost[k];
                matched++;
                break;
            }
        }
    // shortest augmenting paths
    std::vector<double> d(n, INF), pcost(n, 0.0);
    std::vector<int> pred(n, -1);
    std::vector<char> scanned(n, 0);
    std::vector<int> touched, done;
    typedef std::pair<double, int> QE;
    for (int j0 = 0; j0 < n; j0++) {
        if (mrow[j0] >= 0) continue;
        std::priority_queue<QE, std::vector<QE>, std::greater<QE>> heap;
        touched.clear();
        done.clear();
        for (int k = cp[j0]; k < cp[j0 + 1]; k++) {
            int i = ri[k];
            double nd = cost[k] - u[i] - v[j0];
            if (nd < 0) nd = 0;
            if (nd < d[i]) {
                if (d[i] == INF) touched.push_back(i);
                d[i] = nd, pred[i] = j0, pcost[i] = cost[k];
                heap.push(QE(nd, i));
            }
        }
        int iend = -1;
        double dend = 0;
        while (!heap.empty()) {
            QE top = heap.top();
            heap.pop();
            int i = top.second;
            if (scanned[i] || top.first > d[i]) continue;
            scanned[i] = 1;
            done.push_back(i);
            if (mcol[i] < 0) {
                iend = i;
                dend = d[i];
                break;
            }
            int j = mcol[i];
            double base = d[i];
            for (int k = cp[j]; k < cp[j + 1]; k++) {
                int i2 = ri[k];
                if (scanned[i2]) continue;
                double r = cost[k] - u[i2] - v[j];
                if (r < 0) r = 0;
                double nd = base + r;
                if (nd < d[i2]) {
                    if (d[i2] == INF) touched.push_back(i2);
                    d[i2] = nd, pred[i2] = j, pcost[i2] = cost[k];
                    heap.push(QE(nd, i2));
                }
            }
        }
        if (iend >= 0) {
            // dual update on the scanned rows
            for (int i : done) u[i] += d[i] - dend;
            // augment along the predecessor chain
            int i = iend;
            while (true) {
                int j = pred[i];
                int inext = mrow[j];
                mrow[j] = i, mcol[i] = j, mcost[i] = pcost[i];
                if (j == j0) break;
                i = inext;
            }
            // column duals: every tree column is now matched to a scanned row; keep those edges tight
            for (int i2 : done) v[mcol[i2]] = mcost[i2] - u[i2];
            matched++;
        }
        for (int i : touched) d[i] = INF, pred[i] = -1, scanned[i] = 0;
    }
    rowmatch.assign(n, -1);
    for (int j = 0; j < n; j++) rowmatch[j] = mrow[j];
    if (matched < n) return matched;
    rscale.resize(n);
    cscale.resize(n);
    bool sane = !zeros_as_tiny;
    for (int i = 0; i < n && sane; i++)
        if (!(std::fabs(u[i]) < 200.0) || !(std::fabs(v[i]) < 200.0)) sane = false;
    if (sane) {
        for (int i = 0; i < n; i++) rscale[i] = std::exp(u[i]);
        for (int j = 0; j < n; j++) cscale[j] = std::exp(v[j]) / colmax[j];
    } else { // keep the permutation, drop the scaling (it would overflow / be dominated by the placeholder entries)
        rscale.clear();
        cscale.clear();
    }
    return matched;
}

int max_product_matching(int n, const int* rowptr, const int* colidx, const double* vals,
                         std::vector<int>& rowmatch, std::vector<double>& rscale, std::vector<double>& cscale) {
    int m = matching_impl(n, rowptr, colidx, vals, false, rowmatch, rscale, cscale);
    if (m == n) return m;
    // no perfect matching on the numerically nonzero entries: the values seen at analysis time may contain
    // explicit zeros that later become nonzero (Jacobian structures), so retry on the stored PATTERN
    return matching_impl(n, rowptr, colidx, vals, true, rowmatch, rscale, cscale);
}

} // namespace b200
