// Development micro-benchmark (not product code): the 64 x 64 pivot-block LU alone, one CTA, per-step clock stamps;
// k_diag_w8 (register-resident, lean owner path) against the shared-memory fallback k_diag (bit-identical factors), and
// rcp_fast against __drcp_rn.
// Build: nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a tests/dev/diag_bench.cu -o build/diag_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
__device__ long long g_stamps[80];
#define B200_LU_STAMP(k) if (lane == 0) g_stamps[k] = clock64()
#include "../../russell_b200/csrc/kernels.cuh"
using namespace b200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void k_empty() {}
__global__ void k_rcp_check(unsigned long long seed, unsigned long long* nbad, int iters) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    unsigned long long bad = 0;
    for (int i = 0; i < iters; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        unsigned long long bits = x;
        unsigned ex = (unsigned)((bits >> 52) & 0x7ff);
        if (i & 1) { ex = 1023 - 40 + (ex % 80); } // half of the samples near 1
        if (ex < 4 || ex > 2042) ex = 1023;       // rcp_fast's range
        bits = (bits & 0x800fffffffffffffull) | ((unsigned long long)ex << 52);
        if (i % 7 == 0) bits &= 0xfffffffff0000000ull; // short mantissas (grid values)
        const double d = __longlong_as_double((long long)bits);
        const double r1 = __drcp_rn(d), r2 = b200::rcp_fast(d);
        if (__double_as_longlong(r1) != __double_as_longlong(r2)) bad++;
    }
    if (bad) atomicAdd(nbad, bad);
}

int main(int argc, char** argv) {
    const int P = argc > 1 ? atoi(argv[1]) : 64;
    const int NREP = 200;
    // NREP independent fronts (p = P, u = 0), Laplacian-like Schur blocks + noise
    std::vector<double> h((size_t)NREP * P * P);
    srand(7);
    for (int r = 0; r < NREP; r++)
        for (int j = 0; j < P; j++)
            for (int i = 0; i < P; i++) {
                double v = (rand() / (double)RAND_MAX - 0.5);
                if (i == j) v += (r & 1) ? 0.0 : 3.0; // odd fronts: pivoting really moves rows
                h[(size_t)r * P * P + i + (size_t)j * P] = v;
            }
    std::vector<NodeDev> nd(NREP);
    std::vector<int> list(NREP);
    for (int r = 0; r < NREP; r++) {
        nd[r] = NodeDev{};
        nd[r].p = P, nd[r].u = 0, nd[r].c0 = r * P, nd[r].Loff = (long long)r * P * P;
        list[r] = r;
    }
    double *d_fac, *d_fac2, *d_upiv, *d_upiv2;
    int *d_list, *d_lperm, *d_lperm2, *d_cnt;
    NodeDev* d_nd;
    unsigned long long* d_amax;
    size_t bytes = h.size() * 8;
    CK(cudaMalloc(&d_fac, bytes)); CK(cudaMalloc(&d_fac2, bytes));
    CK(cudaMalloc(&d_upiv, NREP * P * 8)); CK(cudaMalloc(&d_upiv2, NREP * P * 8));
    CK(cudaMalloc(&d_lperm, NREP * P * 4)); CK(cudaMalloc(&d_lperm2, NREP * P * 4));
    CK(cudaMalloc(&d_list, NREP * 4)); CK(cudaMalloc(&d_nd, NREP * sizeof(NodeDev)));
    CK(cudaMalloc(&d_cnt, 64)); CK(cudaMalloc(&d_amax, 8));
    CK(cudaMemset(d_cnt, 0, 64));
    double one = 4.0;
    CK(cudaMemcpy(d_amax, &one, 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_list, list.data(), NREP * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_nd, nd.data(), NREP * sizeof(NodeDev), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    { unsigned long long* d_bad; CK(cudaMalloc(&d_bad, 8)); CK(cudaMemset(d_bad, 0, 8));
      k_rcp_check<<<1480, 256>>>(12345ull, d_bad, 4000); CK(cudaDeviceSynchronize());
      unsigned long long nb; CK(cudaMemcpy(&nb, d_bad, 8, cudaMemcpyDeviceToHost));
      printf("rcp_fast vs __drcp_rn: %llu mismatches in %.2e samples\n", nb, 1480.0 * 256 * 4000); }
    // launch overhead
    for (int i = 0; i < 20; i++) k_empty<<<1, 256>>>();
    cudaEventRecord(e0);
    for (int i = 0; i < NREP; i++) k_empty<<<1, 256>>>();
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("empty launch: %.2f us\n", ms * 1e3 / NREP);
    const size_t sm_diag = (size_t)(B200_MAXP * (B200_MAXP + 1)) * sizeof(double) + B200_MAXP * sizeof(int);
    CK(cudaFuncSetAttribute(k_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_diag));
    for (int rep = 0; rep < 2; rep++) { // the shared-memory fallback
        CK(cudaMemcpy(d_fac, h.data(), bytes, cudaMemcpyHostToDevice));
        cudaEventRecord(e0);
        for (int i = 0; i < NREP; i++) k_diag<<<1, 512, sm_diag>>>(d_list + i, d_nd, d_fac, d_lperm, d_upiv, d_amax, 1e-13, d_cnt);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf("k_diag (p=%d): %.2f us per launch\n", P, ms * 1e3 / NREP);
    }
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaMemcpy(d_fac2, h.data(), bytes, cudaMemcpyHostToDevice));
        cudaEventRecord(e0);
        for (int i = 0; i < NREP; i++) k_diag_w8<<<1, 256>>>(d_list + i, d_nd, d_fac2, d_lperm2, d_upiv2, d_amax, 1e-13, d_cnt);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf("k_diag_w8 (p=%d): %.2f us per launch\n", P, ms * 1e3 / NREP);
    }
    long long st[80];
    CK(cudaMemcpyFromSymbol(st, g_stamps, sizeof(st)));
    printf("k_diag_w8 step clocks:");
    for (int k = 1; k < P; k++) printf(" %lld", st[k] - st[k - 1]);
    printf("\n  total %lld cycles for %d steps\n", st[P - 1] - st[0], P - 1);
    // compare
    std::vector<double> f1(h.size()), f2(h.size()), u1(NREP * P), u2(NREP * P);
    std::vector<int> p1(NREP * P), p2(NREP * P);
    CK(cudaMemcpy(f1.data(), d_fac, bytes, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(f2.data(), d_fac2, bytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(u1.data(), d_upiv, NREP * P * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(u2.data(), d_upiv2, NREP * P * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(p1.data(), d_lperm, NREP * P * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(p2.data(), d_lperm2, NREP * P * 4, cudaMemcpyDeviceToHost));
    size_t nbad = 0, nbp = 0, nbu = 0;
    double maxd = 0;
    for (size_t i = 0; i < f1.size(); i++) { if (memcmp(&f1[i], &f2[i], 8)) nbad++; maxd = fmax(maxd, fabs(f1[i] - f2[i])); }
    for (int i = 0; i < NREP * P; i++) { nbp += p1[i] != p2[i]; nbu += memcmp(&u1[i], &u2[i], 8) != 0; }
    printf("compare: %zu factor entries differ (max |d| %.3e), %zu pivots differ, %zu diagonal entries differ\n", nbad, maxd, nbp, nbu);
    int cnt[4];
    CK(cudaMemcpy(cnt, d_cnt, 16, cudaMemcpyDeviceToHost));
    printf("counters %d %d %d\n", cnt[0], cnt[1], cnt[2]);
    return 0;
}
