// Development micro-benchmark: issue throughput of DFMA / SHFL / FSEL per SM as a function of the warps in the CTA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
template <int MODE>
__global__ void k_tput(double* out, long long* cyc, double x, int src) {
    const int lane = threadIdx.x & 31;
    double r[16];
#pragma unroll
    for (int q = 0; q < 16; q++) r[q] = x + q + lane;
    double b = 1.0000001 + x * 1e-9, c = x * 1e-9;
    __syncthreads();
    const long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < 64; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int q = 0; q < 16; q++) r[q] = fma(r[q], b, c);
        } else if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < 16; q++) r[q] = __shfl_sync(0xffffffffu, r[q], src); // 2 SHFL each
        } else if (MODE == 2) { // the consumer's mix: 16 x (SHFL64 + DFMA)
            double u[8];
#pragma unroll
            for (int q = 0; q < 8; q++) u[q] = __shfl_sync(0xffffffffu, r[q], src);
#pragma unroll
            for (int q = 0; q < 8; q++) r[q] = fma(-b, u[q], r[q]), r[q + 8] = fma(-c, u[q], r[q + 8]);
        } else if (MODE == 3) { // DFMA with distinct operand registers
#pragma unroll
            for (int q = 0; q < 16; q++) r[q] = fma(r[q], r[(q + 5) & 15], r[(q + 11) & 15]);
        }
    }
    const long long t1 = clk();
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int q = 0; q < 16; q++) s += r[q];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 8 * 1024); cudaMalloc(&d_cyc, 8 * 32);
    const char* names[] = {"DFMA (16 chains, shared b,c)", "SHFL64 (16 independent = 32 SHFL)", "consumer mix 8 SHFL64 + 16 DFMA", "DFMA (distinct operands)"};
    const int per_iter[] = {16, 32, 32, 16};
    for (int mode = 0; mode < 4; mode++)
        for (int warps = 1; warps <= 16; warps *= 2) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) k_tput<0><<<1, 32 * warps>>>(d_out, d_cyc, 1.25, 3);
                if (mode == 1) k_tput<1><<<1, 32 * warps>>>(d_out, d_cyc, 1.25, 3);
                if (mode == 2) k_tput<2><<<1, 32 * warps>>>(d_out, d_cyc, 1.25, 3);
                if (mode == 3) k_tput<3><<<1, 32 * warps>>>(d_out, d_cyc, 1.25, 3);
                cudaDeviceSynchronize();
            }
            long long c;
            cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-36s warps=%2d: %6lld cycles, %.2f cycles per warp-instruction per warp, %.2f per SM\n", names[mode], warps, c,
                   c / (64.0 * per_iter[mode]), c / (64.0 * per_iter[mode] * warps));
        }
    return 0;
}
