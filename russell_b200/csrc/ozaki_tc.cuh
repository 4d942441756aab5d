// ozaki_tc.cuh -- the Schur-complement update C <- C - L21 * U12 of a big front on the 5th-generation tensor cores.
//
// Role in the reference: the dense BLAS-3 update inside umfpack_di_numeric / cudssExecute(FACTORIZATION)
// (russell_sparse/c_code/interface_umfpack.c:167, interface_cudss.cu:439).
//
// tcgen05.mma has no f64 kind (f16 / tf32 / f8f6f4 / i8 / mx formats only), and the north star needs f64 residuals.  The
// legacy f64 path (mma.sync.m8n8k4, k_schur_dmma) is exact but tops out at the DMMA rate.  This file takes the other road
// (SURVEY 7 hard part 1, option B): an Ozaki-style error-free split of both operands into signed 7-bit slices, exact int32
// products on `tcgen05.mma.kind::i8` with the accumulators in TENSOR MEMORY, and an f64 recombination in the epilogue:
//
//     a_ik = 2^ea_i * sum_s A_s[i,k] * 2^(-7s)      (row exponent ea_i, slices A_s in [-127, 127], s = 1..S)
//     b_kj = 2^eb_j * sum_t B_t[j,k] * 2^(-7t)      (the U panel is stored as U12^T: row j = update column j)
//     c_ij -= 2^(ea_i + eb_j) * sum_{g = 2..S+1} 2^(-7g) * G_g[i,j],      G_g = sum_{s+t=g} A_s B_t^T   (exact in int32)
//
// With S = 8 slices the dropped terms are below 2^-63 relative to (max_k |a_ik|) (max_k |b_kj|) K: the update is as
// accurate as an f64 GEMM whose rows / columns are well scaled, and iterative refinement covers the rest (tested).
//
// Data path (one CTA = one 128 x 64 tile of C, 128 threads):
//   * k_ozaki_split writes every operand tile, slice by slice, in the canonical K-major no-swizzle UMMA layout (8-row x
//     16-byte core matrices), so that ONE bulk copy per slice (cp.async.bulk = TMA, completion on an mbarrier) brings it
//     to shared memory in exactly the form the tensor core reads: no tensor map, no repacking;
//   * one elected thread issues the 36 x (K / 32) tcgen05.mma of the tile: the products of group g accumulate in their own
//     64 TMEM columns (8 groups x 64 columns = all 512 columns of the SM's tensor memory), completion -> tcgen05.commit;
//   * the four warps read the accumulators back with tcgen05.ld (thread = row), recombine the groups by Horner's rule in
//     f64, scale by 2^(ea_i + eb_j) and subtract from C (coalesced along the rows of a column-major C).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "sweep_sub.cuh" // ldg_if

namespace b200 {

#define OZ_S 8        // slices per operand
#define OZ_BM 128     // tile rows (TMEM lanes)
#define OZ_BN 64      // tile columns: (OZ_S groups) x 64 = 512 TMEM columns
#define OZ_KC 64      // K per shared-memory stage
#define OZ_GROUPS OZ_S // groups g = 2 .. S+1

struct OzakiItem {
    long long a_tiles;  // first byte of this front's sliced A tiles: [row tile][k chunk][slice] blocks of OZ_BM x OZ_KC bytes
    long long b_tiles;  // same for B: blocks of OZ_BN x OZ_KC bytes
    long long a_scale;  // offset of 2^ea_i (u doubles) in the scale arena
    long long b_scale;
    long long c_off;    // C in the contribution arena (u x u, column-major)
    int u, kchunks;     // update rows, K chunks of OZ_KC
    int ti, tj;         // tile coordinates
};
struct OzakiSplitItem {
    long long src;      // first element of the operand in fac (row 0, column 0)
    long long tiles;    // first byte of its sliced tiles
    long long scale;    // offset of its scales
    int rows, k, ld;    // rows (u), columns (p), leading dimension of the source
    int tile_rows;      // OZ_BM (A) or OZ_BN (B)
};

// ---- operand split --------------------------------------------------------------------------------------------------------
// one thread per operand row: exponent of the row, then S slices of 7 bits, written in the tile-canonical layout
//   byte(m, k) of a (tile, k chunk, slice) block = (k / 16) * (tile_rows / 8) * 128 + (m / 8) * 128 + (m % 8) * 16 + (k % 16)
__global__ void __launch_bounds__(128) k_ozaki_split(const OzakiSplitItem* __restrict__ items, const double* __restrict__ fac,
                                                     signed char* __restrict__ tiles, double* __restrict__ scales) {
    const OzakiSplitItem it = items[blockIdx.y];
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int nrt = (it.rows + it.tile_rows - 1) / it.tile_rows;
    if (row >= nrt * it.tile_rows) return;
    const int kchunks = (it.k + OZ_KC - 1) / OZ_KC;
    const int tr = row / it.tile_rows, m = row % it.tile_rows;
    const size_t blk = (size_t)it.tile_rows * OZ_KC; // bytes of one block
    signed char* base = tiles + it.tiles + ((size_t)tr * kchunks * OZ_S) * blk + (size_t)(m >> 3) * 128 + (size_t)(m & 7) * 16;
    const size_t kstride = (size_t)(it.tile_rows >> 3) * 128; // between 16-byte K chunks
    if (row >= it.rows) { // padding rows of the last tile: zeros
        for (int kc = 0; kc < kchunks; kc++)
            for (int s = 0; s < OZ_S; s++)
                for (int k16 = 0; k16 < OZ_KC / 16; k16++)
                    *reinterpret_cast<int4*>(base + ((size_t)kc * OZ_S + s) * blk + k16 * kstride) = make_int4(0, 0, 0, 0);
        return;
    }
    const double* src = fac + it.src + row;
    double amax = 0.0;
    for (int k = 0; k < it.k; k++) amax = fmax(amax, fabs(src[(size_t)k * it.ld]));
    int e = 0;
    if (amax > 0.0 && amax < 1.7e308) {
        frexp(amax, &e);              // amax = f * 2^e, f in [0.5, 1)
        if (ldexp(amax, -e) >= 0.9999999999999999) e += 1; // keep |a| 2^-e strictly below 1 after rounding
    }
    scales[it.scale + row] = ldexp(1.0, e);
    for (int kc = 0; kc < kchunks; kc++) {
        for (int k16 = 0; k16 < OZ_KC / 16; k16++) {
            signed char q[OZ_S][16];
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                const int k = kc * OZ_KC + k16 * 16 + kk;
                double v = (k < it.k) ? ldexp(src[(size_t)k * it.ld], -e) : 0.0; // exact scaling, |v| < 1
#pragma unroll
                for (int s = 0; s < OZ_S; s++) {
                    v *= 128.0;                 // exact
                    const double t = trunc(v);  // |t| <= 127
                    q[s][kk] = (signed char)(int)t;
                    v -= t;                     // exact
                }
            }
#pragma unroll
            for (int s = 0; s < OZ_S; s++)
                *reinterpret_cast<int4*>(base + ((size_t)kc * OZ_S + s) * blk + k16 * kstride) = *reinterpret_cast<const int4*>(q[s]);
        }
    }
}

// ---- tcgen05 / TMEM / TMA helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned oz_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// K-major, no swizzle: 8-row core matrices of 128 bytes; LBO = distance of the two 16-byte K chunks of one MMA, SBO = distance
// of consecutive 8-row groups (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ unsigned long long oz_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((saddr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor of kind::i8: D = s32, A = B = signed 8 bit, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ unsigned oz_idesc(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void oz_mma_i8(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void oz_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem(bar)) : "memory");
}
__device__ __forceinline__ void oz_ld16(unsigned taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

#define OZ_A_BYTES (OZ_BM * OZ_KC)                          // one A slice block: 8 KB
#define OZ_B_BYTES (OZ_BN * OZ_KC)                          // one B slice block: 4 KB
#define OZ_SMEM ((size_t)120 * 1024) // 96 KB of operand slices + alignment slack; more than half an SM's shared memory on purpose: ONE CTA per
                                     // SM, because each CTA allocates all 512 columns of the SM's tensor memory

__global__ void __launch_bounds__(128, 1) k_schur_ozaki(const OzakiItem* __restrict__ items, int nitems, const signed char* __restrict__ tiles,
                                                        const double* __restrict__ scales, double* __restrict__ cb) {
    extern __shared__ unsigned char oz_raw[];
    unsigned char* sm = (unsigned char*)(((size_t)oz_raw + 1023) & ~(size_t)1023);
    unsigned char* As = sm;                         // [slice][OZ_BM x OZ_KC]
    unsigned char* Bs = sm + OZ_S * OZ_A_BYTES;     // [slice][OZ_BN x OZ_KC]
    __shared__ __align__(8) unsigned long long bar_full, bar_mma;
    __shared__ unsigned tmem_base_s;
    __shared__ double bsc[OZ_BN];
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) { // one warp allocates all 512 columns of tensor memory for the life of the CTA
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&bar_full, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_base_s;
    const unsigned idesc = oz_idesc(OZ_BM, OZ_BN);
    unsigned ph_full = 0, ph_mma = 0;
    for (int itx = blockIdx.x; itx < nitems; itx += gridDim.x) {
        const OzakiItem it = items[itx];
        const int i0 = it.ti * OZ_BM, j0 = it.tj * OZ_BN;
        for (int kc = 0; kc < it.kchunks; kc++) {
            if (tid == 0) {
                // ---- TMA: the S slices of this K chunk of the A tile and of the B tile, one bulk copy each
                const unsigned bytes = OZ_S * (OZ_A_BYTES + OZ_B_BYTES);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_smem(&bar_full)), "r"(bytes) : "memory");
                const signed char* ga = tiles + it.a_tiles + ((size_t)it.ti * it.kchunks + kc) * (size_t)OZ_S * OZ_A_BYTES;
                const signed char* gb = tiles + it.b_tiles + ((size_t)it.tj * it.kchunks + kc) * (size_t)OZ_S * OZ_B_BYTES;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(oz_smem(As)), "l"(ga),
                             "r"((unsigned)(OZ_S * OZ_A_BYTES)), "r"(oz_smem(&bar_full))
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(oz_smem(Bs)), "l"(gb),
                             "r"((unsigned)(OZ_S * OZ_B_BYTES)), "r"(oz_smem(&bar_full))
                             : "memory");
                mbar_wait(&bar_full, ph_full);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // ---- the 36 slice products of this chunk: group g = s + t accumulates in TMEM columns [64 (g - 2), 64 (g - 1))
                const unsigned a0 = oz_smem(As), b0 = oz_smem(Bs);
                for (int g = 2; g <= OZ_S + 1; g++) {
                    const unsigned td = tmem_base + (unsigned)(g - 2) * OZ_BN;
                    bool first = (kc == 0);
                    for (int s = 1; s < g; s++) {
                        const int t = g - s;
                        if (s > OZ_S || t > OZ_S) continue;
                        for (int ks = 0; ks < OZ_KC / 32; ks++) { // one MMA = K 32 = two 16-byte chunks
                            const unsigned long long ad = oz_desc(a0 + (s - 1) * OZ_A_BYTES + ks * 2 * (OZ_BM / 8) * 128, (OZ_BM / 8) * 128, 128);
                            const unsigned long long bd = oz_desc(b0 + (t - 1) * OZ_B_BYTES + ks * 2 * (OZ_BN / 8) * 128, (OZ_BN / 8) * 128, 128);
                            oz_mma_i8(td, ad, bd, idesc, first ? 0u : 1u);
                            first = false;
                        }
                    }
                }
                oz_commit(&bar_mma); // arrives when every MMA above has completed (and has finished reading shared memory)
            }
            ph_full ^= 1;
            mbar_wait(&bar_mma, ph_mma); // all threads: the operand buffers are free again, the accumulators of this chunk are final
            ph_mma ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: thread = row of the tile (TMEM lane); 16 columns at a time, groups recombined by Horner's rule
        if (tid < OZ_BN) bsc[tid] = (j0 + tid < it.u) ? scales[it.b_scale + j0 + tid] : 0.0;
        __syncthreads();
        const int i = i0 + tid;
        const double rs = (i < it.u) ? scales[it.a_scale + i] : 0.0;
        double* C = cb + it.c_off;
        const unsigned lane_base = tmem_base + ((unsigned)(warp * 32) << 16);
        for (int c16 = 0; c16 < OZ_BN / 16; c16++) {
            // the 16 entries of C this thread updates are requested first (independent loads, in flight under the TMEM reads):
            // a load -> subtract -> store loop per entry costs one DRAM round trip per entry (measured: 40 us per tile)
            double cold[16];
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int j = j0 + c16 * 16 + e;
                cold[e] = ldg_if(C + (size_t)i + (size_t)j * it.u, i < it.u && j < it.u);
            }
            double acc[16];
#pragma unroll
            for (int e = 0; e < 16; e++) acc[e] = 0.0;
            // acc = acc * 2^-7 + G_g, smallest scale first; four groups are read out per tcgen05.wait
#pragma unroll
            for (int gb = OZ_S + 1; gb >= 2; gb -= 4) {
                int v[4][16];
#pragma unroll
                for (int h = 0; h < 4; h++) oz_ld16(lane_base + (unsigned)(gb - h - 2) * OZ_BN + c16 * 16, v[h]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int h = 0; h < 4; h++)
#pragma unroll
                    for (int e = 0; e < 16; e++) acc[e] = fma(acc[e], 0.0078125, (double)v[h][e]);
            }
            if (i < it.u) {
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    const int j = j0 + c16 * 16 + e;
                    if (j < it.u) C[(size_t)i + (size_t)j * it.u] = cold[e] - (acc[e] * 6.103515625e-05 /* 2^-14: g starts at 2 */) * rs * bsc[c16 * 16 + e];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads(); // the accumulators are reused by the next tile
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

} // namespace b200
