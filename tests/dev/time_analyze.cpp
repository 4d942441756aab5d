// Development tool: times b200::analyze (host analysis) on the k x k 5-point Laplacian, phases printed by verbose = 2.
// Build: g++ -O2 -std=c++17 -pthread tests/dev/time_analyze.cpp build/symbolic.o build/ordering.o build/matching.o -o build/time_analyze
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../russell_b200/csrc/plan.hpp"
// FNV-1a over everything the device consumes: two builds of the analysis must print the same hash (bit-identical plans)
template <class T>
static void fnv(uint64_t& h, const std::vector<T>& v) {
    const unsigned char* b = (const unsigned char*)v.data();
    for (size_t i = 0; i < v.size() * sizeof(T); i++) h = (h ^ b[i]) * 1099511628211ull;
}
static uint64_t plan_hash(const b200::Plan& P) {
    uint64_t h = 1469598103934665603ull;
    fnv(h, P.rowperm), fnv(h, P.colperm), fnv(h, P.c0), fnv(h, P.p), fnv(h, P.u), fnv(h, P.parent), fnv(h, P.level);
    fnv(h, P.Loff), fnv(h, P.Uoff), fnv(h, P.Coff), fnv(h, P.Doff), fnv(h, P.rows_ptr), fnv(h, P.rows), fnv(h, P.rel);
    fnv(h, P.child_ptr), fnv(h, P.child_idx), fnv(h, P.level_ptr), fnv(h, P.level_nodes), fnv(h, P.in_sub), fnv(h, P.st_first);
    fnv(h, P.st_root), fnv(h, P.a_src), fnv(h, P.a_dst), fnv(h, P.a_scl), fnv(h, P.full_ptr), fnv(h, P.full_col), fnv(h, P.full_src);
    return h;
}
int main(int argc, char** argv) {
    const int k = argc > 1 ? atoi(argv[1]) : 1000;
    const int n = k * k;
    std::vector<int> ptr(n + 1, 0), col;
    std::vector<double> val;
    for (int i = 0; i < k; i++)
        for (int j = 0; j < k; j++) {
            const int r = i * k + j;
            if (i > 0) col.push_back(r - k), val.push_back(-1);
            if (j > 0) col.push_back(r - 1), val.push_back(-1);
            col.push_back(r), val.push_back(4);
            if (j < k - 1) col.push_back(r + 1), val.push_back(-1);
            if (i < k - 1) col.push_back(r + k), val.push_back(-1);
            ptr[r + 1] = (int)col.size();
        }
    if (getenv("G3")) { // 7-point Laplacian on a k^3 grid instead
        const int k3 = atoi(getenv("G3"));
        ptr.assign(1, 0), col.clear(), val.clear();
        for (int z = 0; z < k3; z++)
            for (int y = 0; y < k3; y++)
                for (int x = 0; x < k3; x++) {
                    const int r = (z * k3 + y) * k3 + x;
                    if (z > 0) col.push_back(r - k3 * k3), val.push_back(-1);
                    if (y > 0) col.push_back(r - k3), val.push_back(-1);
                    if (x > 0) col.push_back(r - 1), val.push_back(-1);
                    col.push_back(r), val.push_back(6);
                    if (x < k3 - 1) col.push_back(r + 1), val.push_back(-1);
                    if (y < k3 - 1) col.push_back(r + k3), val.push_back(-1);
                    if (z < k3 - 1) col.push_back(r + k3 * k3), val.push_back(-1);
                    ptr.push_back((int)col.size());
                }
    }
    if (getenv("SKEW")) { // unsymmetric pattern: drop the west neighbour of every third row (exercises the A+A^T graph)
        std::vector<int> p2(n + 1, 0), c2;
        std::vector<double> v2;
        for (int r = 0; r < n; r++) {
            for (int e = ptr[r]; e < ptr[r + 1]; e++)
                if (!(r % 3 == 0 && col[e] == r - 1)) c2.push_back(col[e]), v2.push_back(val[e]);
            p2[r + 1] = (int)c2.size();
        }
        ptr.swap(p2), col.swap(c2), val.swap(v2);
    }
    for (int rep = 0; rep < (argc > 2 ? atoi(argv[2]) : 2); rep++) {
        b200::AnalyzeOptions opt;
        opt.matching = 2;
        if (getenv("NDLEAF")) opt.nd_leaf = atoi(getenv("NDLEAF"));
        if (getenv("Z2")) opt.relax_z2 = atof(getenv("Z2"));
        opt.verbose = getenv("V") ? atoi(getenv("V")) : 2;
        b200::Plan P;
        auto t0 = std::chrono::steady_clock::now();
        int rc = b200::analyze((int)ptr.size() - 1, ptr.data(), col.data(), val.data(), false, opt, P);
        double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("analyze rc=%d: %.3f s (match %.3f, order %.3f, symbolic %.3f), fronts %d, nnz %lld, flops %.3e\n", rc, t, P.t_match, P.t_order,
               P.t_symbolic, P.nnodes, (long long)(P.nnz_L + P.nnz_U), P.flops);
        printf("plan hash %016llx  fac %lld cb %lld dinv %lld\n", (unsigned long long)plan_hash(P), (long long)P.fac_size, (long long)P.cb_size, (long long)P.dinv_size);
    }
    return 0;
}
