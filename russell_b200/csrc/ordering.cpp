// ordering.cpp -- fill-reducing orderings written from scratch (no METIS/AMD library is linked).
//
// Role in the reference: UMFPACK picks AMD/COLAMD inside umfpack_di_symbolic
// (russell_sparse/c_code/interface_umfpack.c:104-109); cuDSS defaults to METIS nested dissection
// (russell_sparse/src/solver_cudss.rs:401-404).  We use automatic nested dissection (BFS level-structure
// separators from a pseudo-peripheral vertex, trimmed and greedily refined) down to small subdomains that
// are ordered by exact minimum degree.  Nested dissection gives the wide, balanced assembly tree the GPU wants.
#include "plan.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

namespace b200 {

namespace {

struct NdCtx {
    const Graph& g;
    int leaf;
    std::vector<int> label; // partition id of every vertex
    std::vector<int> loc;   // local index scratch
    std::atomic<int> next_id{1}; // (the two halves of a dissection may be ordered by two threads)
    int* out = nullptr;     // perm (new -> old)
    explicit NdCtx(const Graph& gg, int lf) : g(gg), leaf(lf), label(gg.n, 0), loc(gg.n, -1) {}
};

// BFS restricted to label == id.  `order` receives the visit order, `lptr` the level offsets.
// A visited vertex carries the label id ^ VISITED until clear_levels() restores it: one random read per edge (the label)
// instead of two (label + level), and the level boundaries follow from the positions in `order`.  Vertices of this
// subdomain are only ever touched by the thread that owns it (see the fork in nd_component), so the temporary labels are private.
constexpr int VISITED = 0x40000000;
void bfs_levels(NdCtx& c, int id, int start, std::vector<int>& order, std::vector<int>& lptr) {
    order.clear();
    lptr.clear();
    const int seen = id ^ VISITED;
    order.push_back(start);
    c.label[start] = seen;
    lptr.push_back(0);
    size_t head = 0, level_end = 1;
    const int* const ptr = c.g.ptr.data();
    const int* const adj = c.g.adj.data();
    int* const label = c.label.data();
    while (head < order.size()) {
        if (head == level_end) {
            lptr.push_back((int)head);
            level_end = order.size();
        }
        const int v = order[head++];
        for (int e = ptr[v]; e < ptr[v + 1]; e++) {
            const int w = adj[e];
            if (label[w] == id) {
                label[w] = seen;
                order.push_back(w);
            }
        }
    }
    lptr.push_back((int)order.size());
}

inline void clear_levels(NdCtx& c, const std::vector<int>& order) {
    for (int v : order) c.label[v] &= ~VISITED;
}

// exact minimum degree on the subgraph induced by `verts` (all carrying label `id`), bitset adjacency
void local_min_degree(NdCtx& c, const std::vector<int>& verts, int id, int* out) {
    const int m = (int)verts.size();
    if (m == 0) return;
    if (m == 1) {
        out[0] = verts[0];
        return;
    }
    if (m > 4096) { // too large for the dense bitset: cheap static heuristic (increasing degree)
        std::vector<int> idx(verts);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
            return (c.g.ptr[a + 1] - c.g.ptr[a]) < (c.g.ptr[b + 1] - c.g.ptr[b]);
        });
        std::copy(idx.begin(), idx.end(), out);
        return;
    }
    const int W = (m + 63) / 64;
    std::vector<uint64_t> bits((size_t)m * W, 0);
    for (int i = 0; i < m; i++) c.loc[verts[i]] = i;
    for (int i = 0; i < m; i++) {
        int v = verts[i];
        uint64_t* row = &bits[(size_t)i * W];
        for (int e = c.g.ptr[v]; e < c.g.ptr[v + 1]; e++) {
            int w = c.g.adj[e];
            if (c.label[w] == id) {
                int j = c.loc[w];
                if (j >= 0 && j != i) row[j >> 6] |= (1ull << (j & 63));
            }
        }
    }
    for (int i = 0; i < m; i++) c.loc[verts[i]] = -1;
    std::vector<int> deg(m);
    std::vector<char> alive(m, 1);
    for (int i = 0; i < m; i++) {
        int d = 0;
        for (int w = 0; w < W; w++) d += __builtin_popcountll(bits[(size_t)i * W + w]);
        deg[i] = d;
    }
    std::vector<int> nb;
    for (int step = 0; step < m; step++) {
        int best = -1, bd = 1 << 30;
        for (int i = 0; i < m; i++)
            if (alive[i] && deg[i] < bd) {
                bd = deg[i];
                best = i;
            }
        out[step] = verts[best];
        alive[best] = 0;
        uint64_t* rb = &bits[(size_t)best * W];
        nb.clear();
        for (int w = 0; w < W; w++) {
            uint64_t x = rb[w];
            while (x) {
                int b = __builtin_ctzll(x);
                x &= x - 1;
                nb.push_back(w * 64 + b);
            }
        }
        for (int j : nb) {
            uint64_t* rj = &bits[(size_t)j * W];
            int d = 0;
            for (int w = 0; w < W; w++) {
                rj[w] |= rb[w];
            }
            rj[j >> 6] &= ~(1ull << (j & 63));
            rj[best >> 6] &= ~(1ull << (best & 63));
            for (int w = 0; w < W; w++) d += __builtin_popcountll(rj[w]);
            deg[j] = d;
        }
    }
}

void nd_component(NdCtx& c, std::vector<int>& verts, int id, int* out, std::vector<int>& order, std::vector<int>& lptr, int depth);

// `verts` all carry label `id`; writes verts.size() entries at out
void nd_rec(NdCtx& c, std::vector<int>& verts, int id, int* out, int depth = 0) {
    const int m = (int)verts.size();
    if (m == 0) return;
    if (m <= c.leaf) {
        local_min_degree(c, verts, id, out);
        return;
    }
    // split into connected components (iteratively, so that a diagonal matrix does not recurse n deep)
    std::vector<int> order, lptr;
    bfs_levels(c, id, verts[0], order, lptr);
    if ((int)order.size() == m) {
        nd_component(c, verts, id, out, order, lptr, depth);
        return;
    }
    clear_levels(c, order);
    // several components: give each its own label and order them one after another
    int pos = 0;
    std::vector<int> comp;
    std::vector<int> small; // vertices of tiny components are ordered together (they do not interact)
    for (int s : verts) {
        if (c.label[s] != id) continue;
        bfs_levels(c, id, s, order, lptr);
        clear_levels(c, order);
        int cid = c.next_id++;
        for (int v : order) c.label[v] = cid;
        if ((int)order.size() <= c.leaf) {
            comp = order;
            local_min_degree(c, comp, cid, out + pos);
        } else {
            comp = order;
            nd_rec(c, comp, cid, out + pos, depth);
        }
        pos += (int)order.size();
    }
}

// Minimum-vertex-cover refinement of a separator `sep` (label `id`) between sides idA and idB.
// X = separator vertices with a neighbour in B, Y = B vertices with a neighbour in the separator.  A maximum matching of
// the bipartite graph (X, Y) is found by augmenting paths; with Z = everything reachable from the unmatched X vertices by
// alternating paths, (X \ Z) + (Y & Z) covers every X-Y edge.  The new separator is that cover; X & Z and the separator
// vertices without B neighbours go to A.
void shrink_separator_by_cover(NdCtx& c, std::vector<int>& sep, int id, int idA, int idB, long& na, long& nb) {
    std::vector<int> X, Y;
    for (int v : sep) {
        bool touches_b = false;
        for (int e = c.g.ptr[v]; e < c.g.ptr[v + 1] && !touches_b; e++) touches_b = c.label[c.g.adj[e]] == idB;
        if (touches_b) X.push_back(v);
    }
    if (X.size() < 8) return;
    for (size_t i = 0; i < X.size(); i++) c.loc[X[i]] = (int)i;
    for (int v : X)
        for (int e = c.g.ptr[v]; e < c.g.ptr[v + 1]; e++) {
            const int w = c.g.adj[e];
            if (c.label[w] == idB && c.loc[w] < 0) c.loc[w] = (int)Y.size(), Y.push_back(w);
        }
    const int nx = (int)X.size(), ny = (int)Y.size();
    std::vector<int> mx(nx, -1), my(ny, -1), seen(ny, -1);
    // iterative augmenting-path search from every X vertex
    std::vector<int> stack_x, stack_e, parent_y(ny, -1);
    int matched = 0;
    for (int s0 = 0; s0 < nx; s0++) {
        stack_x.assign(1, s0);
        stack_e.assign(1, c.g.ptr[X[s0]]);
        int found = -1;
        while (!stack_x.empty() && found < 0) {
            const int x = stack_x.back();
            int& e = stack_e.back();
            bool pushed = false;
            for (; e < c.g.ptr[X[x] + 1]; e++) {
                const int w = c.g.adj[e];
                if (c.label[w] != idB) continue;
                const int y = c.loc[w];
                if (seen[y] == s0) continue;
                seen[y] = s0;
                parent_y[y] = x;
                if (my[y] < 0) {
                    found = y;
                    break;
                }
                e++;
                stack_x.push_back(my[y]);
                stack_e.push_back(c.g.ptr[X[my[y]]]);
                pushed = true;
                break;
            }
            if (found >= 0 || pushed) continue;
            stack_x.pop_back();
            stack_e.pop_back();
        }
        if (found >= 0) { // flip the path
            int y = found;
            while (y >= 0) {
                const int x = parent_y[y];
                const int prev = mx[x];
                mx[x] = y, my[y] = x;
                y = prev;
            }
            matched++;
        }
    }
    if (matched < nx) { // the cover is smaller than X: apply it
        std::vector<char> zx(nx, 0), zy(ny, 0);
        std::vector<int> q;
        for (int x = 0; x < nx; x++)
            if (mx[x] < 0) zx[x] = 1, q.push_back(x);
        while (!q.empty()) {
            const int x = q.back();
            q.pop_back();
            for (int e = c.g.ptr[X[x]]; e < c.g.ptr[X[x] + 1]; e++) {
                const int w = c.g.adj[e];
                if (c.label[w] != idB) continue;
                const int y = c.loc[w];
                if (zy[y]) continue;
                zy[y] = 1; // reached through a non-matching edge
                const int x2 = my[y];
                if (x2 >= 0 && !zx[x2]) zx[x2] = 1, q.push_back(x2); // and back along the matching edge
            }
        }
        long pulled = 0;
        for (int y = 0; y < ny; y++) pulled += zy[y] ? 1 : 0;
        if (nb - pulled < 1 || 2 * (nb - pulled) < nb) { // side B must survive (and keep at least half of its vertices)
            for (int v : X) c.loc[v] = -1;
            for (int w : Y) c.loc[w] = -1;
            return;
        }
        for (int x = 0; x < nx; x++)
            if (zx[x]) c.label[X[x]] = idA, na++; // covered by its B neighbours, which join the separator
        for (int y = 0; y < ny; y++)
            if (zy[y]) c.label[Y[y]] = id, nb--;
        std::vector<int> nsep;
        for (int v : sep)
            if (c.label[v] == id) nsep.push_back(v);
        for (int y = 0; y < ny; y++)
            if (zy[y]) nsep.push_back(Y[y]);
        sep.swap(nsep);
    }
    for (int v : X) c.loc[v] = -1;
    for (int w : Y) c.loc[w] = -1;
}

// connected subgraph; `order`/`lptr` hold a BFS from verts[0] (its vertices still carry the VISITED bit)
void nd_component(NdCtx& c, std::vector<int>& verts, int id, int* out, std::vector<int>& order, std::vector<int>& lptr, int depth) {
    const int m = (int)verts.size();
    // pseudo-peripheral start: walk to the far end, at most twice.  (Up to four walks until round 2: the third and fourth never
    // changed the separators of the grids measured -- 2D 447..1500, 3D 40^3/60^3 -- or made them ~1 % worse, and every walk is a
    // BFS over the whole subdomain, 60 % of the ordering time.)
    int nl = (int)lptr.size() - 1;
    for (int it = 0; it < 2; it++) {
        int best = -1, bd = 1 << 30;
        for (int k = lptr[nl - 1]; k < lptr[nl]; k++) {
            int v = order[k];
            int d = c.g.ptr[v + 1] - c.g.ptr[v];
            if (d < bd) {
                bd = d;
                best = v;
            }
        }
        clear_levels(c, order);
        std::vector<int> order2, lptr2;
        bfs_levels(c, id, best, order2, lptr2);
        int nl2 = (int)lptr2.size() - 1;
        bool better = nl2 > nl;
        order.swap(order2);
        lptr.swap(lptr2);
        nl = nl2;
        if (!better) break;
    }
    if (nl < 3) { // a clique-like blob: nothing to dissect
        clear_levels(c, order);
        local_min_degree(c, verts, id, out);
        return;
    }
    // choose the separator level: smallest level whose two sides both keep >= 1/3 of the vertices (the numeric phases are
    // bound by the length of the dependent chain of separator panels: at config 2 the 1/4 rule gives the same flops but 73
    // instead of 62 levels; fully balanced bisection gives 62 levels but 23 % more flops)
    int ls = -1;
    {
        long bestsz = -1;
        long bestbal = 0;
        for (int l = 1; l <= nl - 2; l++) {
            long before = lptr[l], sep = lptr[l + 1] - lptr[l], after = m - lptr[l + 1];
            if (std::min(before, after) * 3 < m - sep) continue; // >= 1/3 on each side: same flops as 1/4, 15 % fewer tree levels
            long bal = std::labs(before - after);
            if (ls < 0 || sep < bestsz || (sep == bestsz && bal < bestbal)) {
                ls = l;
                bestsz = sep;
                bestbal = bal;
            }
        }
        if (ls < 0) { // no balanced level: take the one closest to the median
            long bestd = -1;
            for (int l = 1; l <= nl - 2; l++) {
                long before = lptr[l], after = m - lptr[l + 1];
                long d = std::labs(before - after);
                if (ls < 0 || d < bestd) {
                    ls = l;
                    bestd = d;
                }
            }
        }
    }
    const int idA = c.next_id++, idB = c.next_id++;
    long na = 0, nb = 0;
    for (int k = 0; k < lptr[ls]; k++) c.label[order[k]] = idA, na++;
    for (int k = lptr[ls + 1]; k < m; k++) c.label[order[k]] = idB, nb++;
    std::vector<int> sep(order.begin() + lptr[ls], order.begin() + lptr[ls + 1]);
    clear_levels(c, order);
    order.clear();
    order.shrink_to_fit();
    // A BFS level is a wide separator on irregular graphs: shrink it to a minimum vertex cover of the edges that join it
    // to side B (Koenig: |cover| = |maximum matching| of the bipartite boundary graph).  Separator vertices that are
    // not needed move to A, covering B vertices join the separator.  Applied only when it strictly shrinks the
    // separator (on regular grids every level-set vertex is matched: nothing changes there).
    shrink_separator_by_cover(c, sep, id, idA, idB, na, nb);

    // refinement: trim vertices that touch only one side, then zero-gain balancing moves
    for (int pass = 0; pass < 3; pass++) {
        bool changed = false;
        for (size_t si = 0; si < sep.size(); si++) { // index loop: the balancing moves below append to `sep`
            const int v = sep[si];
            if (c.label[v] != id) continue;
            int ca = 0, cb = 0;
            for (int e = c.g.ptr[v]; e < c.g.ptr[v + 1]; e++) {
                int w = c.g.adj[e];
                if (c.label[w] == idA) ca++;
                else if (c.label[w] == idB) cb++;
            }
            if (cb == 0 && (ca > 0 || na <= nb)) {
                c.label[v] = idA, na++, changed = true;
            } else if (ca == 0) {
                c.label[v] = idB, nb++, changed = true;
            } else if (pass > 0 && cb == 1 && na + 1 < nb) {
                // move v to A and pull its single B neighbour into the separator: |S| unchanged, better balance
                for (int e = c.g.ptr[v]; e < c.g.ptr[v + 1]; e++) {
                    int w = c.g.adj[e];
                    if (c.label[w] == idB) {
                        c.label[w] = id;
                        sep.push_back(w);
                        nb--;
                        break;
                    }
                }
                c.label[v] = idA, na++, changed = true;
            } else if (pass > 0 && ca == 1 && nb + 1 < na) {
                for (int e = c.g.ptr[v]; e < c.g.ptr[v + 1]; e++) {
                    int w = c.g.adj[e];
                    if (c.label[w] == idA) {
                        c.label[w] = id;
                        sep.push_back(w);
                        na--;
                        break;
                    }
                }
                c.label[v] = idB, nb++, changed = true;
            }
        }
        if (!changed) break;
    }
    std::vector<int> A, B, S;
    A.reserve(na);
    B.reserve(nb);
    for (int v : verts) {
        int l = c.label[v];
        if (l == idA) A.push_back(v);
        else if (l == idB) B.push_back(v);
        else S.push_back(v);
    }
    verts.clear();
    verts.shrink_to_fit();
    const int sa = (int)A.size(), sb = (int)B.size();
    // the separator is eliminated last
    std::copy(S.begin(), S.end(), out + sa + sb);
    {
        int sid = c.next_id++;
        for (int v : S) c.label[v] = sid; // take separator vertices out of play for the sub-problems
    }
    S.clear();
    S.shrink_to_fit();
    // The two halves are independent: the separator (already relabelled) keeps every vertex of A away from every vertex
    // of B, so the per-vertex scratch arrays (label, loc) are touched at disjoint indices and the halves write
    // disjoint slices of `out`.  The first three dissection levels of a large graph are ordered by two threads each
    // (up to 8 concurrent); the resulting permutation does not depend on the schedule (labels are only compared for equality).
    static const bool serial = getenv("B200_ND_SERIAL") != nullptr; // (tests compare the threaded and the serial ordering)
    static const int tdepth = []() { // levels that fork: 2^tdepth concurrent halves, about twice the hardware threads
        if (const char* e = getenv("B200_ND_THREAD_DEPTH")) return atoi(e);
        const unsigned hc = std::thread::hardware_concurrency();
        int d = 0;
        while ((1u << d) < 2 * hc && d < 6) d++;
        return d;
    }();
    if (!serial && depth < tdepth && sa + sb > 20000) {
        std::thread other([&c, &A, idA, out, depth]() { nd_rec(c, A, idA, out, depth + 1); });
        nd_rec(c, B, idB, out + sa, depth + 1);
        other.join();
    } else {
        nd_rec(c, A, idA, out, depth + 1);
        nd_rec(c, B, idB, out + sa, depth + 1);
    }
}

} // namespace

void order_nested_dissection(const Graph& g, int leaf_size, std::vector<int>& perm) {
    perm.assign(g.n, 0);
    if (g.n == 0) return;
    NdCtx c(g, std::max(leaf_size, 4));
    c.out = perm.data();
    std::vector<int> verts(g.n);
    std::iota(verts.begin(), verts.end(), 0);
    nd_rec(c, verts, 0, perm.data());
}

void order_minimum_degree(const Graph& g, std::vector<int>& perm) {
    perm.assign(g.n, 0);
    if (g.n == 0) return;
    if (g.n > 4096) { // the bitset minimum degree is meant for small graphs: fall back to dissection with small leaves
        order_nested_dissection(g, 32, perm);
        return;
    }
    NdCtx c(g, g.n);
    std::vector<int> verts(g.n);
    std::iota(verts.begin(), verts.end(), 0);
    local_min_degree(c, verts, 0, perm.data());
}

} // namespace b200
