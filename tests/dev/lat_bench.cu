// Development micro-benchmark: dependent-issue latencies (cycles) of the instructions on the pivot-block critical path.
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__device__ __forceinline__ long long clk_after(double dep) { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "d"(dep) : "memory"); return t; }
#define START(a) t0 = clk_after(a); a += (t0 == 123456789ll) ? 1.0 : 0.0;
__global__ void k_lat(double* out, long long* cyc, double x, int src) {
    const int lane = threadIdx.x & 31;
    double a = x + lane, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    int i = 0;
    // DFMA chain
    START(a)
#pragma unroll
    for (int k = 0; k < N; k++) a = fma(a, b, c);
    t1 = clk_after(a); cyc[i++] = t1 - t0;
    // DMUL chain
    START(a)
#pragma unroll
    for (int k = 0; k < N; k++) a = a * b;
    t1 = clk_after(a); cyc[i++] = t1 - t0;
    // DADD chain
    START(a)
#pragma unroll
    for (int k = 0; k < N; k++) a = a + c;
    t1 = clk_after(a); cyc[i++] = t1 - t0;
    // SHFL (64-bit = 2 SHFL in parallel) chain
    t0 = clock64();
#pragma unroll
    for (int k = 0; k < N; k++) a = __shfl_sync(0xffffffffu, a, (src + k) & 31);
    t1 = clock64(); cyc[i++] = t1 - t0;
    // REDUX chain (u32 max)
    unsigned u = (unsigned)__double2hiint(a) + lane;
    t0 = clock64();
#pragma unroll
    for (int k = 0; k < N; k++) u = __reduce_max_sync(0xffffffffu, u ^ (unsigned)k) + lane;
    t1 = clock64(); cyc[i++] = t1 - t0;
    // ballot + ffs chain
    t0 = clock64();
#pragma unroll
    for (int k = 0; k < N; k++) u = __ffs(__ballot_sync(0xffffffffu, (u + k) & 1 || lane == 31)) + u;
    t1 = clock64(); cyc[i++] = t1 - t0;
    // drcp chain
    a = a + u;
    START(a)
#pragma unroll
    for (int k = 0; k < N; k++) a = __drcp_rn(a) + 1.5;
    t1 = clk_after(a); cyc[i++] = t1 - t0;
    // fabs by DADD vs by integer AND, each followed by a DADD
    START(a)
#pragma unroll
    for (int k = 0; k < N; k++) a = fabs(a) + c;
    t1 = clk_after(a); cyc[13] = t1 - t0;
    START(a)
#pragma unroll
    for (int k = 0; k < N; k++) a = __longlong_as_double(__double_as_longlong(a) & 0x7fffffffffffffffll) + c;
    t1 = clk_after(a); cyc[14] = t1 - t0;
    // DFMA x2 / x4 independent
    { double r0 = a, r1 = a + 1, r2 = a + 2, r3 = a + 3;
    START(r0)
#pragma unroll
    for (int k = 0; k < N; k++) { r0 = fma(r0, b, c); r1 = fma(r1, b, c); }
    t1 = clk_after(r0 + r1); cyc[15] = t1 - t0;
    START(r0)
#pragma unroll
    for (int k = 0; k < N; k++) { r0 = fma(r0, b, c); r1 = fma(r1, b, c); r2 = fma(r2, b, c); r3 = fma(r3, b, c); }
    t1 = clk_after(r0 + r1 + r2 + r3); cyc[16] = t1 - t0; a = r0 + r1 + r2 + r3; }
    // float FFMA chain for comparison
    { float g = (float)a;
    t0 = clk_after((double)g); g += (t0 == 123456789ll) ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < N; k++) g = fmaf(g, 1.0000001f, 1e-9f);
    t1 = clk_after((double)g); cyc[17] = t1 - t0; a += g; }
    // shared memory store -> load chain
    __shared__ double sm[64];
    t0 = clock64();
#pragma unroll
    for (int k = 0; k < N; k++) { sm[lane] = a; __syncwarp(); a = sm[(lane + 1) & 31] + 1.0; __syncwarp(); }
    t1 = clock64(); cyc[i++] = t1 - t0;
    // FSEL/IMAD-style integer chain
    int v = u;
    t0 = clock64();
#pragma unroll
    for (int k = 0; k < N; k++) v = (v > k) ? v - lane : v + 3;
    t1 = clock64(); cyc[i++] = t1 - t0;
    // 64-bit unsigned compare+select chain (argmax merge)
    unsigned long long w = (unsigned long long)__double_as_longlong(a), w2 = w ^ 0x5555ull;
    t0 = clock64();
#pragma unroll
    for (int k = 0; k < N; k++) { w = (w > w2) ? w + k : w2 - k; }
    t1 = clock64(); cyc[i++] = t1 - t0;
    // SHFL 32-bit chain
    t0 = clock64();
#pragma unroll
    for (int k = 0; k < N; k++) v = __shfl_sync(0xffffffffu, v, (v + k) & 31);
    t1 = clock64(); cyc[i++] = t1 - t0;
    // DFMA throughput: 8 independent chains
    double r[8];
#pragma unroll
    for (int q = 0; q < 8; q++) r[q] = a + q;
    t0 = clock64();
#pragma unroll
    for (int k = 0; k < N; k++)
#pragma unroll
        for (int q = 0; q < 8; q++) r[q] = fma(r[q], b, c);
    t1 = clock64(); cyc[i++] = t1 - t0;
    // MUFU.RCP64H alone is not reachable from CUDA C; rsqrt(double) for comparison
    t0 = clock64();
#pragma unroll
    for (int k = 0; k < N; k++) a = (double)__frcp_rn((float)a) + 1.5;
    t1 = clock64(); cyc[i++] = t1 - t0;
    double s = a + (double)v + (double)w;
#pragma unroll
    for (int q = 0; q < 8; q++) s += r[q];
    out[threadIdx.x] = s + u;
}
int main() {
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 8 * 256); cudaMalloc(&d_cyc, 8 * 32);
    const char* names[] = {"DFMA", "DMUL", "DADD", "SHFL64", "REDUX.MAX(+IADD)", "BALLOT+FFS(+ops)", "DRCP(+DADD)", "STS->LDS(+DADD,2 syncwarp)", "ISETP+SEL int", "u64 cmp+sel", "SHFL32(+iadd)", "DFMA x8 independent (per 8)", "F2F+FRCP+F2F+DADD", "fabs+DADD", "AND-abs+DADD", "DFMA x2 independent (per 2)", "DFMA x4 independent (per 4)", "FFMA"};
    for (int warps = 1; warps <= 8; warps *= 8) {
        k_lat<<<1, 32 * warps>>>(d_out, d_cyc, 1.25, 3);
        cudaDeviceSynchronize();
        k_lat<<<1, 32 * warps>>>(d_out, d_cyc, 1.25, 3);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        long long c[32];
        cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
        printf("-- %d warp(s) in the CTA (last writer's numbers)\n", warps);
        for (int i = 0; i < 18; i++) printf("%-32s %7.1f cycles per op\n", names[i], c[i] / (double)N);
    }
    return 0;
}
