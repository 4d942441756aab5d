// symbolic.cpp -- elimination tree, column counts, relaxed supernodes, front plan (host, once per structure).
//
// Reference role: umfpack_di_symbolic (russell_sparse/c_code/interface_umfpack.c:109) /
// CUDSS_PHASE_ANALYSIS (russell_sparse/c_code/interface_cudss.cu:361).  The algorithms are the textbook
// ones (Liu's elimination tree with path compression, Gilbert-Ng-Peyton column counts, supernode
// amalgamation with an explicit-zero budget); the code is written from scratch for the flat, GPU-facing
// Plan layout in plan.hpp.
#include "plan.hpp"
#include <iterator>
#include <map>
#include <set>

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <numeric>
#include <atomic>
#include <cstdlib>
#include <thread>

namespace b200 {

// problem sizes below which the threaded variants of the analysis run inline; B200_PAR_FLOOR overrides them (the tests force
// the threaded code paths on small matrices and compare the plan with the serial one)
static int par_floor(int dflt) {
    static const int forced = getenv("B200_PAR_FLOOR") ? atoi(getenv("B200_PAR_FLOOR")) : -1;
    return forced >= 0 ? forced : dflt;
}

// splits [0, n) into contiguous chunks, one per hardware thread (at most 16); small ranges run inline
template <class F>
static void parallel_rows(int n, F fn) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt > 16) nt = 16;
    if (nt < 2 || n < par_floor(50000) || getenv("B200_ND_SERIAL")) {
        fn(0, n);
        return;
    }
    std::vector<std::thread> th;
    const int chunk = (n + (int)nt - 1) / (int)nt;
    for (unsigned t = 1; t < nt; t++) {
        const int a = (int)t * chunk, b = std::min(n, a + chunk);
        if (a < b) th.emplace_back([=, &fn]() { fn(a, b); });
    }
    fn(0, std::min(n, chunk));
    for (auto& x : th) x.join();
}


namespace {
double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}
inline int64_t round_up4(int64_t x) { return (x + 3) & ~int64_t(3); }
} // namespace

// Liu's algorithm on a symmetric graph (uses the neighbours i < k of every vertex k)
void etree_symmetric(const Graph& g, std::vector<int>& parent) {
    const int n = g.n;
    parent.assign(n, -1);
    std::vector<int> anc(n, -1);
    for (int k = 0; k < n; k++) {
        for (int e = g.ptr[k]; e < g.ptr[k + 1]; e++) {
            int i = g.adj[e];
            while (i != -1 && i < k) {
                int nxt = anc[i];
                anc[i] = k;
                if (nxt == -1) parent[i] = k;
                i = nxt;
            }
        }
    }
}

// post[k] = vertex visited k-th; children are visited in increasing `weight` (heaviest child last)
void postorder_tree(const std::vector<int>& parent, const std::vector<int>& weight, std::vector<int>& post) {
    const int n = (int)parent.size();
    std::vector<int> cptr(n + 2, 0), cidx(n);
    // bucket children under their parent; virtual root = n
    for (int v = 0; v < n; v++) cptr[(parent[v] < 0 ? n : parent[v]) + 1]++;
    for (int v = 0; v <= n; v++) cptr[v + 1] += cptr[v];
    {
        std::vector<int> fill(cptr.begin(), cptr.end() - 1);
        for (int v = 0; v < n; v++) cidx[fill[parent[v] < 0 ? n : parent[v]]++] = v;
    }
    if (!weight.empty()) {
        parallel_rows(n + 1, [&](int v0, int v1) { // every vertex sorts its own child list
            for (int v = v0; v < v1; v++) {
                int a = cptr[v], b = cptr[v + 1];
                if (b - a > 1)
                    std::stable_sort(cidx.begin() + a, cidx.begin() + b, [&](int x, int y) { return weight[x] < weight[y]; });
            }
        });
    }
    // No traversal is needed: a vertex is emitted right after its subtree, so its position is (start of its subtree) + (size
    // of its subtree) - 1, and the subtrees of the children of one vertex follow each other in list order.  Sizes bottom-up
    // and starts top-down are plain loops because parents carry larger indices than their children (elimination trees,
    // supernode trees); the general case falls back to an explicit depth-first traversal.
    bool increasing = true;
    for (int v = 0; v < n && increasing; v++) increasing = parent[v] < 0 || parent[v] > v;
    post.assign(n, 0);
    if (increasing) {
        std::vector<int> size(n + 1, 1), start(n + 1, 0);
        size[n] = 0;
        for (int v = 0; v < n; v++) size[parent[v] < 0 ? n : parent[v]] += size[v];
        for (int v = n; v >= 0; v--) {
            int at = start[v];
            for (int e = cptr[v]; e < cptr[v + 1]; e++) start[cidx[e]] = at, at += size[cidx[e]];
        }
        parallel_rows(n, [&](int a, int b) { for (int v = a; v < b; v++) post[start[v] + size[v] - 1] = v; });
        return;
    }
    post.clear();
    post.reserve(n);
    // iterative DFS
    std::vector<int> stack, it(n + 1);
    for (int v = 0; v <= n; v++) it[v] = cptr[v];
    stack.push_back(n);
    while (!stack.empty()) {
        int v = stack.back();
        if (it[v] < cptr[v + 1]) {
            stack.push_back(cidx[it[v]++]);
        } else {
            stack.pop_back();
            if (v != n) post.push_back(v);
        }
    }
}

// Gilbert-Ng-Peyton column counts of the Cholesky factor of a symmetric pattern (diagonal included)
static void column_counts_post(const Graph& g, const std::vector<int>& parent, const std::vector<int>& post,
                               std::vector<int>& cc) {
    const int n = g.n;
    std::vector<int> first(n, -1), maxfirst(n, -1), prevleaf(n, -1), anc(n);
    cc.assign(n, 0);
    for (int k = 0; k < n; k++) {
        int j = post[k];
        cc[j] = (first[j] == -1) ? 1 : 0;
        for (; j != -1 && first[j] == -1; j = parent[j]) first[j] = k;
    }
    std::iota(anc.begin(), anc.end(), 0);
    for (int k = 0; k < n; k++) {
        int j = post[k];
        if (parent[j] != -1) cc[parent[j]]--;
        for (int e = g.ptr[j]; e < g.ptr[j + 1]; e++) {
            int i = g.adj[e];
            if (i <= j || first[j] <= maxfirst[i]) continue;
            maxfirst[i] = first[j];
            int jprev = prevleaf[i];
            prevleaf[i] = j;
            if (jprev == -1) {
                cc[j]++;
            } else {
                int q = jprev;
                while (q != anc[q]) q = anc[q];
                for (int s = jprev; s != q;) {
                    int sp = anc[s];
                    anc[s] = q;
                    s = sp;
                }
                cc[j]++;
                cc[q]--;
            }
        }
        if (parent[j] != -1) anc[j] = parent[j];
    }
    for (int j = 0; j < n; j++)
        if (parent[j] != -1) cc[parent[j]] += cc[j];
}

void column_counts(const Graph& g, const std::vector<int>& parent, std::vector<int>& cc) {
    std::vector<int> post, none;
    postorder_tree(parent, none, post);
    column_counts_post(g, parent, post, cc);
}

// relabels a symmetric graph: vertex old -> inv[old]
static void relabel_graph(const Graph& g, const std::vector<int>& perm /*new->old*/, Graph& out) {
    const int n = g.n;
    std::vector<int> inv(n);
    for (int k = 0; k < n; k++) inv[perm[k]] = k;
    out.n = n;
    out.ptr.assign(n + 1, 0);
    for (int k = 0; k < n; k++) out.ptr[k + 1] = out.ptr[k] + (g.ptr[perm[k] + 1] - g.ptr[perm[k]]);
    out.adj.resize(g.adj.size());
    parallel_rows(n, [&](int kbeg, int kend) { // every vertex fills and sorts its own adjacency slice
        for (int k = kbeg; k < kend; k++) {
            int o = perm[k], d = out.ptr[k];
            for (int e = g.ptr[o]; e < g.ptr[o + 1]; e++) out.adj[d++] = inv[g.adj[e]];
            std::sort(out.adj.begin() + out.ptr[k], out.adj.begin() + out.ptr[k + 1]);
        }
    });
}

int analyze(int n, const int* rowptr, const int* colidx, const double* vals, bool sym_lower,
            const AnalyzeOptions& opt, Plan& P) {
    P = Plan();
    P.opt = opt;
    P.n = n;
    P.sym_lower = sym_lower;
    if (n < 1 || rowptr == nullptr || colidx == nullptr) return -2;
    const int nnz = rowptr[n];
    P.nnz_in = nnz;
    if (rowptr[0] != 0 || nnz < 1) return -2;
    for (int i = 0; i < n; i++)
        if (rowptr[i + 1] < rowptr[i]) return -2;
    {
        std::atomic<int> bad{0};
        parallel_rows(n, [&](int ibeg, int iend) {
            for (int i = ibeg; i < iend; i++)
                for (int k = rowptr[i]; k < rowptr[i + 1]; k++) {
                    const int j = colidx[k];
                    // the CSR contract (csr_matrix.rs:359-480): columns ascending, duplicates already summed.  A duplicate would make
                    // the value scatter last-writer-wins while the SpMV sums it, so it is rejected rather than tolerated.
                    if (j < 0 || j >= n || (sym_lower && j > i) || (!sym_lower && k > rowptr[i] && j <= colidx[k - 1])) {
                        bad = 1;
                        return;
                    }
                }
        });
        if (bad) return -2;
    }
    const int W = std::max(4, std::min(opt.panel_width, 128));

    // ---- full (mirrored) CSR of the original matrix -------------------------------------------------
    std::vector<int>&fptr = P.full_ptr, &fcol = P.full_col, &fsrc = P.full_src;
    if (sym_lower) {
        fptr.assign(n + 1, 0);
        for (int i = 0; i < n; i++)
            for (int k = rowptr[i]; k < rowptr[i + 1]; k++) {
                fptr[i + 1]++;
                if (colidx[k] != i) fptr[colidx[k] + 1]++;
            }
        for (int i = 0; i < n; i++) fptr[i + 1] += fptr[i];
        fcol.resize(fptr[n]);
        fsrc.resize(fptr[n]);
        std::vector<int> fill(fptr.begin(), fptr.end() - 1);
        // rows are visited in increasing order, so each row's columns come out sorted:
        // mirrored entries (j,i) with i>j land after row j's own lower entries (cols <= j)
        for (int i = 0; i < n; i++)
            for (int k = rowptr[i]; k < rowptr[i + 1]; k++) {
                int j = colidx[k];
                fcol[fill[i]] = j, fsrc[fill[i]] = k, fill[i]++;
            }
        // second sweep for the mirrored part keeps ascending column order per row
        for (int i = 0; i < n; i++)
            for (int k = rowptr[i]; k < rowptr[i + 1]; k++) {
                int j = colidx[k];
                if (j != i) fcol[fill[j]] = i, fsrc[fill[j]] = k, fill[j]++;
            }
        for (int i = 0; i < n; i++) { // lower entries of a row might be unsorted in the input: sort pairs
            int a = fptr[i], b = fptr[i + 1];
            bool sorted = true;
            for (int k = a + 1; k < b; k++)
                if (fcol[k] < fcol[k - 1]) sorted = false;
            if (!sorted) {
                std::vector<std::pair<int, int>> tmp(b - a);
                for (int k = a; k < b; k++) tmp[k - a] = {fcol[k], fsrc[k]};
                std::sort(tmp.begin(), tmp.end());
                for (int k = a; k < b; k++) fcol[k] = tmp[k - a].first, fsrc[k] = tmp[k - a].second;
            }
            for (int k = a + 1; k < b; k++)
                if (fcol[k] == fcol[k - 1]) return -2; // duplicate entry
        }
    } else {
        fptr.assign(rowptr, rowptr + n + 1);
        fcol.assign(colidx, colidx + nnz);
        fsrc.clear(); // identity
    }
    const int fnnz = fptr[n];
    auto src_of = [&](int k) { return fsrc.empty() ? k : fsrc[k]; };

    // ---- matching / scaling ----------------------------------------------------------------------------
    double t0 = now_s();
    std::vector<int> rowmatch; // rowmatch[j] = row matched to column j
    bool want_match = (opt.matching == 1);
    if (opt.matching == 2 && vals != nullptr) {
        // auto: only when some diagonal entry is structurally or numerically zero
        for (int i = 0; i < n && !want_match; i++) {
            bool ok = false;
            for (int k = fptr[i]; k < fptr[i + 1]; k++)
                if (fcol[k] == i && vals[src_of(k)] != 0.0) ok = true;
            if (!ok) want_match = true;
        }
    }
    if (want_match && vals != nullptr) {
        std::vector<double> fv(fnnz);
        for (int k = 0; k < fnnz; k++) fv[k] = vals[src_of(k)];
        int matched = max_product_matching(n, fptr.data(), fcol.data(), fv.data(), rowmatch, P.rscale, P.cscale);
        if (matched < n) return -1; // structurally singular
        P.matched = false;
        for (int j = 0; j < n; j++)
            if (rowmatch[j] != j) P.matched = true;
    } else {
        rowmatch.resize(n);
        std::iota(rowmatch.begin(), rowmatch.end(), 0);
    }
    std::vector<int> colmatch(n); // colmatch[i] = column matched to row i
    for (int j = 0; j < n; j++) colmatch[rowmatch[j]] = j;
    P.t_match = now_s() - t0;

    // ---- graph of A' + A'^T (A' = row-permuted A, vertex = column index) --------------------------------
    t0 = now_s();
    Graph g0;
    g0.n = n;
    // Fast path (the usual case: PDE / Jacobian patterns): without a row matching and with a structurally symmetric pattern
    // the graph is the pattern itself minus the diagonal -- checked and copied by row-parallel threads.
    bool pattern_is_graph = !P.matched;
    if (pattern_is_graph)
        for (int j = 0; j < n && pattern_is_graph; j++) pattern_is_graph = rowmatch[j] == j;
    if (pattern_is_graph) {
        std::atomic<int> asym{0};
        g0.ptr.assign(n + 1, 0);
        parallel_rows(n, [&](int ibeg, int iend) {
            for (int i = ibeg; i < iend && !asym.load(std::memory_order_relaxed); i++) {
                int cnt = 0;
                for (int k = fptr[i]; k < fptr[i + 1]; k++) {
                    const int j = fcol[k];
                    if (j == i) continue;
                    cnt++;
                    if (!std::binary_search(fcol.begin() + fptr[j], fcol.begin() + fptr[j + 1], i)) { asym = 1; break; }
                }
                g0.ptr[i + 1] = cnt;
            }
        });
        pattern_is_graph = asym == 0;
    }
    if (pattern_is_graph) {
        for (int v = 0; v < n; v++) g0.ptr[v + 1] += g0.ptr[v];
        g0.adj.resize(g0.ptr[n]);
        parallel_rows(n, [&](int ibeg, int iend) {
            for (int i = ibeg; i < iend; i++) {
                int d = g0.ptr[i];
                for (int k = fptr[i]; k < fptr[i + 1]; k++)
                    if (fcol[k] != i) g0.adj[d++] = fcol[k]; // (rows are sorted and duplicate-free: checked above)
            }
        });
    } else
    {
        std::vector<int> deg(n + 1, 0);
        for (int i = 0; i < n; i++) {
            int ip = colmatch[i];
            for (int k = fptr[i]; k < fptr[i + 1]; k++) {
                int j = fcol[k];
                if (j != ip) deg[ip + 1]++, deg[j + 1]++;
            }
        }
        for (int i = 0; i < n; i++) deg[i + 1] += deg[i];
        std::vector<int> adj(deg[n]);
        std::vector<int> fill(deg.begin(), deg.end() - 1);
        for (int i = 0; i < n; i++) {
            int ip = colmatch[i];
            for (int k = fptr[i]; k < fptr[i + 1]; k++) {
                int j = fcol[k];
                if (j != ip) adj[fill[ip]++] = j, adj[fill[j]++] = ip;
            }
        }
        // every vertex sorts its own slice and counts its distinct neighbours; a prefix sum places the compacted lists
        g0.ptr.assign(n + 1, 0);
        parallel_rows(n, [&](int vbeg, int vend) {
            for (int v = vbeg; v < vend; v++) {
                std::sort(adj.begin() + deg[v], adj.begin() + deg[v + 1]);
                g0.ptr[v + 1] = (int)(std::unique(adj.begin() + deg[v], adj.begin() + deg[v + 1]) - (adj.begin() + deg[v]));
            }
        });
        for (int v = 0; v < n; v++) g0.ptr[v + 1] += g0.ptr[v];
        g0.adj.resize(g0.ptr[n]);
        parallel_rows(n, [&](int vbeg, int vend) {
            for (int v = vbeg; v < vend; v++) std::copy(adj.begin() + deg[v], adj.begin() + deg[v] + (g0.ptr[v + 1] - g0.ptr[v]), g0.adj.begin() + g0.ptr[v]);
        });
    }
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .g0 at %.3f s\n", now_s() - t0);

    // ---- fill-reducing ordering ----------------------------------------------------------------------
    std::vector<int> q; // new -> old
    if (opt.ordering == ORDERING_NATURAL) {
        q.resize(n);
        std::iota(q.begin(), q.end(), 0);
    } else if (opt.ordering == ORDERING_MINDEG) {
        order_minimum_degree(g0, q);
    } else {
        order_nested_dissection(g0, opt.nd_leaf, q);
    }
    P.t_order = now_s() - t0;

    // ---- elimination tree, postorder (heaviest child last), final labelling -----------------------------
    t0 = now_s();
    Graph g;
    std::vector<int> parent, cc;
    {
        Graph g1;
        relabel_graph(g0, q, g1);
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .relabel1 at %.3f s\n", now_s() - t0);
        std::vector<int> par1, cc1, post1, none;
        etree_symmetric(g1, par1);
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .etree1 at %.3f s\n", now_s() - t0);
        postorder_tree(par1, none, post1);
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .post1 at %.3f s\n", now_s() - t0);
        column_counts_post(g1, par1, post1, cc1);
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .cc1 at %.3f s\n", now_s() - t0);
        std::vector<int> post;
        postorder_tree(par1, cc1, post);
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .post2 at %.3f s\n", now_s() - t0);
        std::vector<int> q2(n);
        parallel_rows(n, [&](int a, int b) { for (int k = a; k < b; k++) q2[k] = q[post[k]]; });
        q.swap(q2);
        relabel_graph(g0, q, g);
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .relabel2 at %.3f s\n", now_s() - t0);
        // A postorder only renames the vertices of the elimination tree: parents and column counts of the final
        // labelling follow from the first pass by renaming (no second etree / column-count computation).
        std::vector<int> newid(n);
        parallel_rows(n, [&](int a, int b) { for (int k = a; k < b; k++) newid[post[k]] = k; });
        parent.assign(n, -1);
        cc.assign(n, 0);
        parallel_rows(n, [&](int a, int b) {
            for (int k = a; k < b; k++) {
                const int old = post[k];
                parent[k] = par1[old] < 0 ? -1 : newid[par1[old]];
                cc[k] = cc1[old];
            }
        });
    }
    g0 = Graph();
    std::vector<int> invq(n);
    P.colperm = q;
    P.rowperm.resize(n);
    parallel_rows(n, [&](int a, int b) {
        for (int k = a; k < b; k++) invq[q[k]] = k, P.rowperm[k] = rowmatch[q[k]];
    });

    if (opt.verbose >= 2) fprintf(stderr, "b200 analyze:   phase etree done at %.3f s\n", now_s() - t0);
    // ---- fundamental supernodes, then amalgamation over the supernode tree -------------------------------
    // Any child may be merged into its parent (not only a contiguous last child): columns of different
    // subtrees are independent, so the merged children's columns can be moved right in front of the parent's
    // columns without changing the fill.  The budget is counted in STORED entries p*(2f-p) of the L/U panels.
    struct Grp {
        int first, ncols;
        int64_t f;     // front order (pivots + update rows)
        int64_t zeros; // explicit zeros accumulated in the stored panels
    };
    std::vector<Grp> grp;
    {
        std::vector<Grp> fund;
        for (int j = 0; j < n; j++) {
            bool join = j > 0 && parent[j - 1] == j && cc[j] == cc[j - 1] - 1;
            if (join) fund.back().ncols++;
            else fund.push_back({j, 1, cc[j], 0});
        }
        const int nf = (int)fund.size();
        P.nsuper_fundamental = nf;
        std::vector<int> c2s(n);
        parallel_rows(nf, [&](int s0, int s1) {
            for (int s = s0; s < s1; s++)
                for (int j = fund[s].first; j < fund[s].first + fund[s].ncols; j++) c2s[j] = s;
        });
        std::vector<int> fpar(nf, -1);
        parallel_rows(nf, [&](int s0, int s1) {
            for (int s = s0; s < s1; s++) {
                int lastc = fund[s].first + fund[s].ncols - 1;
                fpar[s] = parent[lastc] < 0 ? -1 : c2s[parent[lastc]];
            }
        });
        // children lists of the fundamental supernode tree
        std::vector<int> cptr(nf + 1, 0), cidx(nf);
        for (int s = 0; s < nf; s++)
            if (fpar[s] >= 0) cptr[fpar[s] + 1]++;
        for (int s = 0; s < nf; s++) cptr[s + 1] += cptr[s];
        {
            std::vector<int> fill(cptr.begin(), cptr.end() - 1);
            for (int s = 0; s < nf; s++)
                if (fpar[s] >= 0) cidx[fill[fpar[s]]++] = s;
        }
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .fundtree at %.3f s\n", now_s() - t0);
        // bottom-up greedy merging; nc/ff/zz describe the merged group rooted at s
        std::vector<int64_t> nc(nf), ff(nf), zz(nf, 0);
        std::vector<char> merged_into_parent(nf, 0);
        for (int s = 0; s < nf; s++) nc[s] = fund[s].ncols, ff[s] = fund[s].f;
        auto stor = [](int64_t c, int64_t f) { return c * (2 * f - c); };
        std::vector<int> kids;
        for (int s = 0; s < nf; s++) { // children have smaller indices: already final
            kids.assign(cidx.begin() + cptr[s], cidx.begin() + cptr[s + 1]);
            // cheapest candidates first: small children with tall fronts add the fewest zeros
            std::sort(kids.begin(), kids.end(), [&](int a, int b) {
                int64_t da = nc[a] * (ff[s] - ff[a] + nc[a]), db = nc[b] * (ff[s] - ff[b] + nc[b]);
                return da != db ? da < db : a < b;
            });
            for (int c : kids) {
                const int64_t nm = nc[s] + nc[c], fm = ff[s] + nc[c];
                const int64_t delta = stor(nm, fm) - stor(nc[s], ff[s]) - stor(nc[c], ff[c]);
                const int64_t z = zz[s] + zz[c] + delta;
                const double ratio = (double)z / (double)stor(nm, fm);
                bool merge;
                if (nm <= opt.relax_small) merge = true;
                else if (nm <= 32) merge = ratio <= opt.relax_z1;
                else if (nm <= W) merge = ratio <= opt.relax_z2;
                else merge = ratio <= opt.relax_z3;
                if (!merge) continue;
                merged_into_parent[c] = 1;
                nc[s] = nm, ff[s] = fm, zz[s] = z;
            }
        }
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .merge at %.3f s\n", now_s() - t0);
        // new elimination order: for every group, first the (unmerged) child groups, then the group's columns
        // (merged children's columns before their parent's columns)
        std::vector<int> order;
        order.reserve(n);
        std::vector<int> roots;
        for (int s = 0; s < nf; s++)
            if (fpar[s] < 0) roots.push_back(s);
        // iterative two-phase traversal: phase 0 = emit unmerged descendants' groups, phase 1 = emit own columns
        struct Frame {
            int s, phase;
        };
        std::vector<Frame> st;
        std::vector<int> members; // supernodes of the current group in emission order
        std::vector<int> mem_store, mem_off(nf, 0), mem_end(nf, 0); // the member lists, kept from phase 0 for phase 1
        mem_store.reserve(nf);
        // collect(s): members of the group rooted at s = collect(merged children)... + s
        std::vector<std::pair<int, int>> stack;
        auto collect = [&](int root, std::vector<int>& out) {
            // post-order over merged-children edges only
            stack.clear();
            stack.push_back({root, cptr[root]});
            while (!stack.empty()) {
                auto& top = stack.back();
                int sidx = top.first;
                bool pushed = false;
                while (top.second < cptr[sidx + 1]) {
                    int c = cidx[top.second++];
                    if (merged_into_parent[c]) {
                        stack.push_back({c, cptr[c]});
                        pushed = true;
                        break;
                    }
                }
                if (pushed) continue;
                out.push_back(sidx);
                stack.pop_back();
            }
        };
        for (int r : roots) st.push_back({r, 0});
        std::reverse(st.begin(), st.end());
        std::vector<int> tmp;
        while (!st.empty()) {
            Frame fr = st.back();
            st.pop_back();
            if (fr.phase == 1) { // the members were listed when the group was scheduled (phase 0)
                Grp gnew{(int)order.size(), 0, ff[fr.s], zz[fr.s]};
                for (int q2 = mem_off[fr.s]; q2 < mem_end[fr.s]; q2++) {
                    const int m = mem_store[q2];
                    for (int j = fund[m].first; j < fund[m].first + fund[m].ncols; j++) order.push_back(j);
                }
                gnew.ncols = (int)order.size() - gnew.first;
                grp.push_back(gnew);
                continue;
            }
            // phase 0: schedule own emission after all unmerged child groups of every member
            st.push_back({fr.s, 1});
            members.clear();
            collect(fr.s, members);
            mem_off[fr.s] = (int)mem_store.size();
            mem_store.insert(mem_store.end(), members.begin(), members.end());
            mem_end[fr.s] = (int)mem_store.size();
            tmp.clear();
            for (int m : members)
                for (int e = cptr[m]; e < cptr[m + 1]; e++)
                    if (!merged_into_parent[cidx[e]]) tmp.push_back(cidx[e]);
            for (auto it = tmp.rbegin(); it != tmp.rend(); ++it) st.push_back({*it, 0});
        }
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .emit at %.3f s\n", now_s() - t0);
        // relabel everything to the new order
        bool identity = true;
        for (int k = 0; k < n; k++)
            if (order[k] != k) identity = false;
        if (!identity) {
            std::vector<int> q2(n);
            parallel_rows(n, [&](int a, int b) { for (int k = a; k < b; k++) q2[k] = q[order[k]]; });
            q.swap(q2);
            Graph g2;
            relabel_graph(g, order, g2);
            g.ptr.swap(g2.ptr);
            g.adj.swap(g2.adj);
            // `order` lists every vertex before its elimination-tree parent (merged children before their parent's columns,
            // child groups before parent groups): an equivalent reordering, so the tree of the new labelling is the old one renamed
            {
                std::vector<int> newid(n), par2(n);
                parallel_rows(n, [&](int a, int b) { for (int k = a; k < b; k++) newid[order[k]] = k; });
                parallel_rows(n, [&](int a, int b) {
                    for (int k = a; k < b; k++) par2[k] = parent[order[k]] < 0 ? -1 : newid[parent[order[k]]];
                });
                parent.swap(par2);
            }
            P.colperm = q;
            parallel_rows(n, [&](int a, int b) {
                for (int k = a; k < b; k++) invq[q[k]] = k, P.rowperm[k] = rowmatch[q[k]];
            });
        }
    }
    const int ns = (int)grp.size();
    P.nsuper_relaxed = ns;
    std::vector<int> col2sn(n);
    parallel_rows(ns, [&](int s0, int s1) {
        for (int s = s0; s < s1; s++)
            for (int j = grp[s].first; j < grp[s].first + grp[s].ncols; j++) col2sn[j] = s;
    });
    std::vector<int> sparent(ns, -1);
    parallel_rows(ns, [&](int s0, int s1) {
        for (int s = s0; s < s1; s++) {
            int lastc = grp[s].first + grp[s].ncols - 1;
            sparent[s] = parent[lastc] < 0 ? -1 : col2sn[parent[lastc]];
        }
    });

    if (opt.verbose >= 2) fprintf(stderr, "b200 analyze:   phase amalgamation done at %.3f s\n", now_s() - t0);
    // ---- row structure of every supernode (rows beyond its last column, ascending) ----------------------
    std::vector<std::vector<int>> srows(ns);
    {
        std::vector<int> mark(n, -1);
        std::vector<int> scptr(ns + 1, 0), scidx(ns);
        for (int s = 0; s < ns; s++)
            if (sparent[s] >= 0) scptr[sparent[s] + 1]++;
        for (int s = 0; s < ns; s++) scptr[s + 1] += scptr[s];
        {
            std::vector<int> fill(scptr.begin(), scptr.end() - 1);
            for (int s = 0; s < ns; s++)
                if (sparent[s] >= 0) scidx[fill[sparent[s]]++] = s;
        }
        // one supernode: union of the adjacency of its columns and of its children's rows, beyond its last column
        auto row_struct = [&](int s, std::vector<int>& mk) {
            const int first = grp[s].first, last = first + grp[s].ncols - 1;
            std::vector<int>& r = srows[s];
            for (int j = first; j <= last; j++)
                for (int e = g.ptr[j]; e < g.ptr[j + 1]; e++) {
                    int i = g.adj[e];
                    if (i > last && mk[i] != s) mk[i] = s, r.push_back(i);
                }
            for (int e = scptr[s]; e < scptr[s + 1]; e++) {
                int c = scidx[e];
                for (int i : srows[c])
                    if (i > last && mk[i] != s) mk[i] = s, r.push_back(i);
            }
            std::sort(r.begin(), r.end());
            if ((int64_t)r.size() + grp[s].ncols != grp[s].f) {
                if (opt.verbose)
                    fprintf(stderr, "b200 analyze: supernode %d predicted front %lld, found %lld\n", s,
                            (long long)grp[s].f, (long long)(r.size() + grp[s].ncols));
                grp[s].f = (int64_t)r.size() + grp[s].ncols; // the enumerated structure is authoritative
            }
        };
        // A supernode needs its children only: disjoint subtrees (contiguous ranges, the supernodes are in postorder) are
        // independent tasks for a pool of threads, each with its own mark array; the few supernodes above them follow serially.
        unsigned nt = std::thread::hardware_concurrency();
        if (nt > 16) nt = 16;
        std::vector<char> done(ns, 0);
        if (nt >= 2 && ns >= par_floor(20000) && !getenv("B200_ND_SERIAL")) {
            std::vector<int64_t> wsub(ns);
            std::vector<int> cnt(ns, 1);
            for (int s = 0; s < ns; s++) wsub[s] = grp[s].f;
            for (int s = 0; s < ns; s++)
                if (sparent[s] >= 0) wsub[sparent[s]] += wsub[s], cnt[sparent[s]] += cnt[s];
            int64_t total = 0;
            for (int s = 0; s < ns; s++)
                if (sparent[s] < 0) total += wsub[s];
            const int64_t target = std::max<int64_t>(total / (8 * (int64_t)nt), 1);
            std::vector<std::pair<int64_t, int>> tasks; // (weight, root)
            for (int s = 0; s < ns; s++)
                if (wsub[s] <= target && (sparent[s] < 0 || wsub[sparent[s]] > target)) tasks.push_back({wsub[s], s});
            std::sort(tasks.begin(), tasks.end(), [](const std::pair<int64_t, int>& x, const std::pair<int64_t, int>& y) {
                return x.first != y.first ? x.first > y.first : x.second < y.second;
            });
            std::atomic<size_t> next{0};
            auto worker = [&]() {
                std::vector<int> mk(n, -1);
                for (size_t t = next++; t < tasks.size(); t = next++) {
                    const int root = tasks[t].second;
                    for (int s2 = root - cnt[root] + 1; s2 <= root; s2++) row_struct(s2, mk), done[s2] = 1;
                }
            };
            std::vector<std::thread> th;
            for (unsigned t = 1; t < nt; t++) th.emplace_back(worker);
            worker();
            for (auto& x : th) x.join();
        }
        for (int s = 0; s < ns; s++)
            if (!done[s]) row_struct(s, mark);
    }
    g = Graph();

    if (opt.verbose >= 2) fprintf(stderr, "b200 analyze:   phase rowstruct done at %.3f s\n", now_s() - t0);
    // ---- split wide supernodes into panels: the front tree ------------------------------------------------
    std::vector<int> sn_firstnode(ns), sn_npieces(ns);
    int nnodes = 0;
    for (int s = 0; s < ns; s++) {
        int K = (grp[s].ncols + W - 1) / W;
        sn_firstnode[s] = nnodes;
        sn_npieces[s] = K;
        nnodes += K;
    }
    P.nnodes = nnodes;
    P.c0.resize(nnodes), P.p.resize(nnodes), P.u.resize(nnodes), P.parent.assign(nnodes, -1), P.level.assign(nnodes, 0);
    P.Loff.resize(nnodes), P.Uoff.resize(nnodes), P.Coff.resize(nnodes), P.Doff.resize(nnodes);
    P.rows_ptr.assign(nnodes + 1, 0);
    std::vector<int> col2node(n);
    for (int s = 0; s < ns; s++) {
        const int K = sn_npieces[s], nc = grp[s].ncols;
        int off = 0;
        for (int k = 0; k < K; k++) {
            int w = nc / K + (k < nc % K ? 1 : 0);
            int v = sn_firstnode[s] + k;
            P.c0[v] = grp[s].first + off;
            P.p[v] = w;
            P.u[v] = (nc - off - w) + (int)srows[s].size();
            P.parent[v] = (k + 1 < K) ? v + 1 : (sparent[s] >= 0 ? sn_firstnode[sparent[s]] : -1);
            for (int j = 0; j < w; j++) col2node[P.c0[v] + j] = v;
            off += w;
            P.rows_ptr[v + 1] = P.rows_ptr[v] + P.u[v];
        }
    }
    // children lists + levels
    P.child_ptr.assign(nnodes + 1, 0);
    P.child_idx.resize(nnodes);
    for (int v = 0; v < nnodes; v++)
        if (P.parent[v] >= 0) P.child_ptr[P.parent[v] + 1]++;
    for (int v = 0; v < nnodes; v++) P.child_ptr[v + 1] += P.child_ptr[v];
    {
        std::vector<int> fill(P.child_ptr.begin(), P.child_ptr.end() - 1);
        for (int v = 0; v < nnodes; v++)
            if (P.parent[v] >= 0) P.child_idx[fill[P.parent[v]]++] = v;
    }
    P.child_idx.resize(P.child_ptr[nnodes]);
    int nlev = 0;
    for (int v = 0; v < nnodes; v++) { // children precede parents
        if (P.parent[v] >= 0) P.level[P.parent[v]] = std::max(P.level[P.parent[v]], P.level[v] + 1);
        nlev = std::max(nlev, P.level[v] + 1);
    }
    P.nlevels = nlev;
    P.level_ptr.assign(nlev + 1, 0);
    for (int v = 0; v < nnodes; v++) P.level_ptr[P.level[v] + 1]++;
    for (int l = 0; l < nlev; l++) P.level_ptr[l + 1] += P.level_ptr[l];
    P.level_nodes.resize(nnodes);
    {
        std::vector<int> fill(P.level_ptr.begin(), P.level_ptr.end() - 1);
        for (int v = 0; v < nnodes; v++) P.level_nodes[fill[P.level[v]]++] = v;
    }
    // The contribution-block allocator needs only the tree (levels, u, children): it runs on a helper thread underneath the
    // row lists, relative indices, subtree layout and scatter map (it was 0.1 s of a 0.55 s symbolic phase at 1M dof).
    int64_t co_reuse = 0;
    P.Coff.assign(nnodes, 0);
    auto alloc_cb = [&]() {
        // Contribution blocks live only from the level that first writes them to the level of their parent, and the
        // device executes the tree level by level: blocks whose lifetimes do not overlap share storage.  (One block per
        // front for the whole factorization costs sum(u^2): 159 GB for a 64^3 27-point grid, TBs at 115^3.)
        //   born(v) = level(v), or the level of its only child when that child may write v's block from its Schur
        //             epilogue (chain links of a split supernode);   dies(v) = level(parent(v)).
        std::vector<int> born(nnodes), dies(nnodes);
        for (int v = 0; v < nnodes; v++) {
            born[v] = P.level[v];
            if (P.child_ptr[v + 1] - P.child_ptr[v] == 1) born[v] = std::min(born[v], P.level[P.child_idx[P.child_ptr[v]]]);
            dies[v] = P.parent[v] >= 0 ? P.level[P.parent[v]] : P.level[v];
        }
        std::vector<std::vector<int>> born_at(nlev), dies_at(nlev);
        for (int v = 0; v < nnodes; v++)
            if (P.u[v] > 0) born_at[born[v]].push_back(v), dies_at[dies[v]].push_back(v);
        std::map<int64_t, int64_t> free_by_off;            // offset -> size (coalesced)
        std::set<std::pair<int64_t, int64_t>> free_by_size; // (size, offset)
        auto release = [&](int64_t off, int64_t size) {
            auto nx = free_by_off.lower_bound(off);
            if (nx != free_by_off.begin()) {
                auto pv = std::prev(nx);
                if (pv->first + pv->second == off) { off = pv->first, size += pv->second; free_by_size.erase({pv->second, pv->first}); free_by_off.erase(pv); }
            }
            if (nx != free_by_off.end() && off + size == nx->first) { size += nx->second; free_by_size.erase({nx->second, nx->first}); free_by_off.erase(nx); }
            free_by_off[off] = size;
            free_by_size.insert({size, off});
        };
        std::vector<std::pair<int64_t, int64_t>> dying;
        for (int l = 0; l < nlev; l++) {
            // largest first: big blocks take the big holes
            std::sort(born_at[l].begin(), born_at[l].end(), [&](int a, int b) { return P.u[a] != P.u[b] ? P.u[a] > P.u[b] : a < b; });
            for (int v : born_at[l]) {
                const int64_t need = round_up4((int64_t)P.u[v] * P.u[v]);
                auto it = free_by_size.lower_bound({need, (int64_t)0}); // best fit
                if (it != free_by_size.end()) {
                    const int64_t size = it->first, off = it->second;
                    free_by_size.erase(it);
                    free_by_off.erase(off);
                    P.Coff[v] = off;
                    if (size > need) release(off + need, size - need);
                } else {
                    // grow the arena: a free block that ends at the top is extended instead of wasted
                    int64_t off = co_reuse;
                    if (!free_by_off.empty()) {
                        auto last = std::prev(free_by_off.end());
                        if (last->first + last->second == co_reuse) { off = last->first; free_by_size.erase({last->second, last->first}); free_by_off.erase(last); }
                    }
                    P.Coff[v] = off;
                    co_reuse = off + need;
                }
            }
            // reusable from the NEXT level on.  The blocks that die together are merged among themselves first (sorted by
            // offset, one linear pass): the free set after a batch of releases does not depend on their order, and the
            // balanced trees then see one operation per run instead of one per block (the bottom levels free ~80k blocks each)
            dying.clear();
            for (int v : dies_at[l]) dying.push_back({P.Coff[v], round_up4((int64_t)P.u[v] * P.u[v])});
            std::sort(dying.begin(), dying.end());
            for (size_t i = 0; i < dying.size();) {
                int64_t off = dying[i].first, size = dying[i].second;
                size_t j = i + 1;
                while (j < dying.size() && dying[j].first == off + size) size += dying[j].second, j++;
                release(off, size);
                i = j;
            }
        }
    };
    std::thread cb_thread;
    struct Joiner { // (early error returns below must not leave a running thread behind)
        std::thread& t;
        void finish() { if (t.joinable()) t.join(); }
        ~Joiner() { finish(); }
    } cb_join{cb_thread};
    if (opt.cb_reuse) {
        if (nnodes >= par_floor(20000) && !getenv("B200_ND_SERIAL")) cb_thread = std::thread(alloc_cb);
        else alloc_cb();
    }
    P.rows.resize(P.rows_ptr[nnodes]);
    P.rel.assign(P.rows_ptr[nnodes], -1);
    for (int s = 0; s < ns; s++) {
        const int K = sn_npieces[s];
        const int lastcol = grp[s].first + grp[s].ncols - 1;
        for (int k = 0; k < K; k++) {
            int v = sn_firstnode[s] + k;
            int* r = &P.rows[P.rows_ptr[v]];
            int t = 0;
            for (int j = P.c0[v] + P.p[v]; j <= lastcol; j++) r[t++] = j; // remaining columns of the supernode
            for (int i : srows[s]) r[t++] = i;
            assert(t == P.u[v]);
        }
    }
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .rows at %.3f s\n", now_s() - t0);
    // relative indices into the parent front
    for (int s = 0; s < ns; s++) {
        const int K = sn_npieces[s];
        for (int k = 0; k < K; k++) {
            int v = sn_firstnode[s] + k;
            int* rel = &P.rel[P.rows_ptr[v]];
            if (k + 1 < K) {
                for (int i = 0; i < P.u[v]; i++) rel[i] = i; // the next panel's front is exactly this update set
            } else if (sparent[s] >= 0) {
                int t = sparent[s];
                const int tfirst = grp[t].first, tn = grp[t].ncols;
                const std::vector<int>& tr = srows[t];
                size_t w = 0;
                const int* r = &P.rows[P.rows_ptr[v]];
                for (int i = 0; i < P.u[v]; i++) {
                    int gi = r[i];
                    if (gi < tfirst + tn) {
                        rel[i] = gi - tfirst;
                    } else {
                        while (w < tr.size() && tr[w] < gi) w++;
                        if (w >= tr.size() || tr[w] != gi) return -2; // structure inconsistency (should not happen)
                        rel[i] = tn + (int)w;
                    }
                }
            }
        }
    }
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .rel at %.3f s\n", now_s() - t0);
    // storage offsets + stats
    if (opt.verbose >= 2) fprintf(stderr, "b200 analyze:   phase fronttree done at %.3f s\n", now_s() - t0);
    // ---- solve-phase subtrees: maximal subtrees whose fronts are all small (nodes are in postorder: the subtree rooted at
    //      v is the contiguous range [v - size + 1, v])
    P.in_sub.assign(nnodes, 0);
    if (opt.st_enable) {
        std::vector<int64_t> ent(nnodes, 0);
        std::vector<int> size(nnodes, 1), cols(nnodes, 0);
        std::vector<char> elig(nnodes, 1);
        for (int v = 0; v < nnodes; v++) { // children precede parents: the sums of v are final when v is visited
            ent[v] += round_up4((int64_t)P.p[v] * (P.p[v] + P.u[v]));
            cols[v] += P.p[v];
            if (P.p[v] + P.u[v] > opt.st_maxf || P.p[v] > opt.st_pmax || ent[v] > opt.st_budget || cols[v] + P.u[v] > opt.st_maxcols)
                elig[v] = 0;
            const int par = P.parent[v];
            if (par >= 0) {
                ent[par] += ent[v], size[par] += size[v], cols[par] += cols[v];
                if (!elig[v]) elig[par] = 0;
            }
        }
        int count = 0;
        for (int v = 0; v < nnodes; v++)
            if (elig[v] && (P.parent[v] < 0 || !elig[P.parent[v]])) count++;
        if (count >= opt.st_min_count)
            for (int v = 0; v < nnodes; v++)
                if (elig[v] && (P.parent[v] < 0 || !elig[P.parent[v]])) {
                    P.st_first.push_back(v - size[v] + 1), P.st_root.push_back(v);
                    for (int w = v - size[v] + 1; w <= v; w++) P.in_sub[w] = 1;
                }
    }
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .subtrees at %.3f s\n", now_s() - t0);
    int64_t fo = 0, co = 0, dof = 0;
    {
        size_t si = 0;
        for (int v = 0; v < nnodes;) {
            if (si < P.st_first.size() && P.st_first[si] == v) { // a subtree: all L panels, then all U panels
                const int r = P.st_root[si++];
                for (int w = v; w <= r; w++) P.Loff[w] = fo, fo = round_up4(fo + (int64_t)(P.p[w] + P.u[w]) * P.p[w]);
                for (int w = v; w <= r; w++) P.Uoff[w] = fo, fo = round_up4(fo + (int64_t)P.u[w] * P.p[w]);
                v = r + 1;
            } else {
                const int64_t p = P.p[v], u = P.u[v];
                P.Loff[v] = fo, fo = round_up4(fo + (p + u) * p);
                P.Uoff[v] = fo, fo = round_up4(fo + u * p);
                v++;
            }
        }
    }
    for (int v = 0; v < nnodes; v++) {
        int64_t p = P.p[v], u = P.u[v], f = p + u;
        if (!opt.cb_reuse) P.Coff[v] = co, co = round_up4(co + u * u);
        P.Doff[v] = dof, dof = round_up4(dof + p * p);
        P.nnz_L += f * p;
        P.nnz_U += u * p;
        P.flops += (2.0 / 3.0) * p * p * p + 2.0 * p * p * u + 2.0 * (double)p * u * u;
        P.max_front = std::max<int>(P.max_front, (int)f);
    }
    if (opt.verbose >= 3) fprintf(stderr, "b200 analyze:     .layout at %.3f s\n", now_s() - t0);
    P.fac_size = fo, P.dinv_size = dof;

    if (opt.verbose >= 2) fprintf(stderr, "b200 analyze:   phase subtrees+layout+cb done at %.3f s\n", now_s() - t0);
    // ---- value scatter map ---------------------------------------------------------------------------
    const bool scaled = !P.rscale.empty();
    P.a_src.resize(fnnz);
    P.a_dst.resize(fnnz);
    if (scaled) P.a_scl.resize(fnnz);
    std::atomic<int> scatter_bad{0};
    parallel_rows(n, [&](int ibeg, int iend) { // rows are independent: every entry writes its own slot of the map
    for (int i = ibeg; i < iend; i++) {
        const int kr = invq[colmatch[i]];
        for (int k = fptr[i]; k < fptr[i + 1]; k++) {
            const int j = fcol[k];
            const int kc = invq[j];
            int64_t dst;
            if (kc <= kr) { // lower triangle (incl. diagonal): column kc's node, L panel
                int v = col2node[kc];
                int64_t f = P.p[v] + P.u[v];
                int rpos;
                if (kr < P.c0[v] + P.p[v]) rpos = kr - P.c0[v];
                else {
                    const int* r = &P.rows[P.rows_ptr[v]];
                    const int* it = std::lower_bound(r, r + P.u[v], kr);
                    if (it == r + P.u[v] || *it != kr) { scatter_bad = 1; return; }
                    rpos = P.p[v] + (int)(it - r);
                }
                dst = P.Loff[v] + rpos + (int64_t)(kc - P.c0[v]) * f;
            } else { // upper triangle: row kr's node
                int v = col2node[kr];
                int64_t f = P.p[v] + P.u[v];
                if (kc < P.c0[v] + P.p[v]) dst = P.Loff[v] + (kr - P.c0[v]) + (int64_t)(kc - P.c0[v]) * f;
                else {
                    const int* r = &P.rows[P.rows_ptr[v]];
                    const int* it = std::lower_bound(r, r + P.u[v], kc);
                    if (it == r + P.u[v] || *it != kc) { scatter_bad = 1; return; }
                    dst = P.Uoff[v] + (int64_t)(it - r) + (int64_t)(kr - P.c0[v]) * P.u[v];
                }
            }
            P.a_src[k] = src_of(k);
            P.a_dst[k] = dst;
            if (scaled) P.a_scl[k] = P.rscale[i] * P.cscale[j];
        }
    }
    });
    cb_join.finish();
    P.cb_size = opt.cb_reuse ? co_reuse : co;
    if (scatter_bad) return -2;
    if (opt.verbose >= 2) fprintf(stderr, "b200 analyze:   phase scattermap done at %.3f s\n", now_s() - t0);
    P.t_symbolic = now_s() - t0;
    if (opt.verbose) {
        fprintf(stderr,
                "b200 analyze: n=%d nnz=%d nodes=%d (fund %d, relaxed %d) levels=%d nnz(L)=%lld nnz(U)=%lld "
                "flops=%.3e maxfront=%d cb=%.1f MB  t(match,order,symb)=%.3f %.3f %.3f s\n",
                n, fnnz, nnodes, P.nsuper_fundamental, P.nsuper_relaxed, P.nlevels, (long long)P.nnz_L,
                (long long)P.nnz_U, P.flops, P.max_front, P.cb_size * 8e-6, P.t_match, P.t_order, P.t_symbolic);
    }
    return 0;
}

} // namespace b200
