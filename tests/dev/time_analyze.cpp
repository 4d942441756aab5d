// Development tool: times b200::analyze (host analysis) on the k x k 5-point Laplacian, phases printed by verbose = 2.
// Build: g++ -O2 -std=c++17 -pthread tests/dev/time_analyze.cpp build/symbolic.o build/ordering.o build/matching.o -o build/time_analyze
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../russell_b200/csrc/plan.hpp"
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
    for (int rep = 0; rep < (argc > 2 ? atoi(argv[2]) : 2); rep++) {
        b200::AnalyzeOptions opt;
        opt.matching = 2;
        opt.verbose = 2;
        b200::Plan P;
        auto t0 = std::chrono::steady_clock::now();
        int rc = b200::analyze(n, ptr.data(), col.data(), val.data(), false, opt, P);
        double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("analyze rc=%d: %.3f s (match %.3f, order %.3f, symbolic %.3f), fronts %d, nnz %lld, flops %.3e\n", rc, t, P.t_match, P.t_order,
               P.t_symbolic, P.nnodes, (long long)(P.nnz_L + P.nnz_U), P.flops);
    }
    return 0;
}
