// sweep_sub.cuh -- triangular solves over the BOTTOM of the front tree: one WARP per small subtree.
//
// Role in the reference: the inside of umfpack_di_solve / cudssExecute(SOLVE)
// (russell_sparse/c_code/interface_umfpack.c:229, interface_cudss.cu:530) for the ~95 % of the fronts that are tiny.
//
// The round-1 kernels (k_fwd_subtree / k_bwd_subtree) walked a subtree front by front with the multifrontal vector scheme:
// every front gathered its children's update vectors through relative indices, built its own update vector and wrote it
// back to global memory.  ncu showed them bound by instruction issue (~900 warp instructions per front for ~370 stored
// entries) and, in the backward direction, by p short column segments per front read out of the f x p L panels (1.42x the
// algorithmic DRAM traffic).  These kernels are right-looking on a shared-memory segment of the solution vector instead:
//   * the columns of a subtree are contiguous (postorder) -> xs[0 .. ncols) holds them, xs[ncols .. ncols + u_root) holds the
//     update rows of the subtree's root (everything that leaves the subtree); a host-built 16-bit target index per update
//     row says where a row of any front of the subtree lives in xs.  No child lists, no relative indices, no per-front
//     update vectors, no global traffic between the fronts of a subtree;
//   * the host analysis lays the subtree's L panels out contiguously, then its U panels; a post-pass of the factorization
//     (k_pack_pivot_blocks) packs the pivot blocks next to each other: the forward walk streams the L region, the backward
//     walk the U region + the packed pivot blocks -- the algorithmic bytes, in full lines (measured: 196 MB of DRAM traffic
//     per direction for 173 MB of panels at config 2);
//   * a subtree needs only xs in shared memory (<= 8 KB), so 24-28 warps = subtrees are resident per SM; a warp requests all
//     loads of a front at once (coalesced: lanes = rows), the other warps hide their latency, and nothing but __syncwarp
//     orders the walk: ~100 instructions per front.
// A first version staged a whole subtree (24 KB on average) in shared memory with ONE bulk copy per array (cp.async.bulk on
// an mbarrier, one CTA per subtree): correct, but 4-5 CTAs per SM were too few fronts in flight -- 1.10 ms per sweep against
// 0.92 ms for round 1 and 0.80 ms for the warp-per-subtree walk (profiles/r2c_*, r2e_*); it was removed.
#pragma once
#include "kernels.cuh"

namespace b200 {

struct SubtreeDev {          // built once by the host
    long long Lbeg, Ubeg, Dbeg; // first double of the subtree's L region / U region in fac, of its pivot-block copies in dinv
    long long tgt_beg;          // first entry of its target indices (uint16)
    long long root_rows;        // offset of the root's update-row list in rows[] / wv[]
    int Lcount, Ucount, Dcount; // doubles (multiples of 4)
    int tgt_count;              // uint16 entries, padded to a multiple of 8
    int pu_beg;                 // first (p, u) pair of its fronts inside the pu array
    int nfr;                    // fronts
    int cbeg, ncols;            // first column and number of columns
    int next;                   // update rows of the root
    int pan;                    // max(Lcount, Ucount + Dcount)
};

__device__ __forceinline__ int sub_round4(int x) { return (x + 3) & ~3; }

// predicated global load that the compiler cannot sink next to its first use: ncu (profiles/r2e) showed the backward walk
// issuing load -> multiply -> load -> multiply (eight exposed DRAM latencies per front) because the front end had merged the
// batched loads back into the loop that consumes them.  asm volatile keeps the loads in program order, back to back.
__device__ __forceinline__ double ldg_if(const double* p, const bool pred) {
    double v;
    asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n mov.f64 %0, 0d0000000000000000;\n @q ld.global.nc.f64 %0, [%1];\n}"
                 : "=d"(v)
                 : "l"(p), "r"((int)pred));
    return v;
}

#define B200_SUBW_WARPS 4
#define B200_SUBW_XS 1024 // doubles of shared memory per warp: columns of the subtree + update rows of its root

// sums v[0..7] over the 32 lanes with 9 shuffles (instead of 40): after the three halving steps lane l holds the total of
// column ((l >> 4) & 1) * 4 + ((l >> 3) & 1) * 2 + ((l >> 2) & 1), complete after the last two steps.  Fixed order: deterministic.
__device__ __forceinline__ double warp_reduce8(double (&v)[8], const int lane) {
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const double send = h16 ? v[q] : v[q + 4];
        const double keep = h16 ? v[q + 4] : v[q];
        v[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const double send = h8 ? v[q] : v[q + 2];
        const double keep = h8 ? v[q + 2] : v[q];
        v[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const double send = h4 ? v[0] : v[1];
        const double keep = h4 ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}
// the lane that holds column c (0..7) after warp_reduce8
__device__ __forceinline__ int reduce8_lane_of(int c) { return ((c & 4) ? 16 : 0) | ((c & 2) ? 8 : 0) | ((c & 1) ? 4 : 0); }

__global__ void __launch_bounds__(32 * B200_SUBW_WARPS) k_fwd_stree_w(const SubtreeDev* __restrict__ trees, int ntrees,
                                                                       const double* __restrict__ fac, const unsigned short* __restrict__ tgt_all,
                                                                       const uchar2* __restrict__ pu_all, const int* __restrict__ lperm,
                                                                       const double* __restrict__ y, double* __restrict__ zv,
                                                                       double* __restrict__ wv) {
    __shared__ double xs_all[B200_SUBW_WARPS][B200_SUBW_XS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* xs = xs_all[warp];
    for (int sidx = blockIdx.x * B200_SUBW_WARPS + warp; sidx < ntrees; sidx += gridDim.x * B200_SUBW_WARPS) {
        const SubtreeDev st = trees[sidx];
        for (int i = lane; i < st.ncols; i += 32) xs[i] = y[st.cbeg + i];
        for (int i = lane; i < st.next; i += 32) xs[st.ncols + i] = 0.0;
        __syncwarp();
        const double* L = fac + st.Lbeg;
        const unsigned short* tg = tgt_all + st.tgt_beg;
        const uchar2* pu = pu_all + st.pu_beg;
        const int* lp = lperm + st.cbeg;
        int c0 = 0;
        uchar2 q = pu[0];
        for (int fr = 0; fr < st.nfr; fr++) {
            const int p = q.x, u = q.y, f = p + u;
            if (fr + 1 < st.nfr) q = pu[fr + 1]; // the next front's shape, one iteration ahead
            // ---- everything this front reads from global memory is requested here, before the first dependent use
            const int lpv = lane < p ? lp[c0 + lane] : 0;
            int t0 = 0, t1 = 0, t2 = 0;
            if (lane < u) t0 = tg[lane];
            if (u > 32) { // warp-uniform
                if (lane + 32 < u) t1 = tg[lane + 32];
                if (lane + 64 < u) t2 = tg[lane + 64];
            }
            double l11[8], a[8];
#pragma unroll
            for (int m = 0; m < 8; m++) l11[m] = ldg_if(L + lane + m * f, lane < p && m < lane);
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = ldg_if(L + p + lane + k * f, lane < u && k < p);
            // ---- z1 = inv(L11) P t1: lane = row, L11 unit lower triangular, eight columns of L11 at a time
            double tv = lane < p ? xs[c0 + lpv] : 0.0;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const double zm = __shfl_sync(0xffffffffu, tv, m);
                if (m + 1 < p) tv -= l11[m] * zm; // (l11[m] = 0 for the lanes at or above the diagonal)
            }
            for (int m0 = 8; m0 + 1 < p; m0 += 8) { // p > 9: rare at the bottom of the tree
#pragma unroll
                for (int m = 0; m < 8; m++) l11[m] = ldg_if(L + lane + (m0 + m) * f, lane < p && m0 + m < lane);
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    const double zm = __shfl_sync(0xffffffffu, tv, m0 + m);
                    if (m0 + m + 1 < p) tv -= l11[m] * zm;
                }
            }
            __syncwarp();
            if (lane < p) xs[c0 + lane] = tv;
            // ---- update rows: lane = row (three row blocks when u > 64), z broadcast from the registers of lanes 0..p-1
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) acc += a[k] * __shfl_sync(0xffffffffu, tv, k);
            for (int k0 = 8; k0 < p; k0 += 8) {
#pragma unroll
                for (int k = 0; k < 8; k++) a[k] = ldg_if(L + p + lane + (k0 + k) * f, lane < u && k0 + k < p);
#pragma unroll
                for (int k = 0; k < 8; k++) acc += a[k] * __shfl_sync(0xffffffffu, tv, (k0 + k) & 31);
            }
            __syncwarp();
            if (lane < u) xs[t0] -= acc; // the rows of one front are distinct: no conflicting updates
            for (int rb = 1; rb * 32 < u; rb++) { // rows 32.. (u <= 96)
                const int r = lane + 32 * rb;
                double acc2 = 0.0;
                for (int k0 = 0; k0 < p; k0 += 8) {
#pragma unroll
                    for (int k = 0; k < 8; k++) a[k] = ldg_if(L + p + r + (k0 + k) * f, r < u && k0 + k < p);
#pragma unroll
                    for (int k = 0; k < 8; k++) acc2 += a[k] * __shfl_sync(0xffffffffu, tv, (k0 + k) & 31);
                }
                if (r < u) xs[rb == 1 ? t1 : t2] -= acc2;
            }
            __syncwarp();
            L += sub_round4(f * p), tg += u, c0 += p;
        }
        for (int i = lane; i < st.ncols; i += 32) zv[st.cbeg + i] = xs[i];
        for (int i = lane; i < st.next; i += 32) wv[st.root_rows + i] = xs[st.ncols + i];
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32 * B200_SUBW_WARPS) k_bwd_stree_w(const SubtreeDev* __restrict__ trees, int ntrees,
                                                                       const double* __restrict__ fac, const double* __restrict__ dpack,
                                                                       const unsigned short* __restrict__ tgt_all, const uchar2* __restrict__ pu_all,
                                                                       const int* __restrict__ rows_all, const double* __restrict__ zv,
                                                                       double* __restrict__ xp) {
    __shared__ double xs_all[B200_SUBW_WARPS][B200_SUBW_XS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* xs = xs_all[warp];
    for (int sidx = blockIdx.x * B200_SUBW_WARPS + warp; sidx < ntrees; sidx += gridDim.x * B200_SUBW_WARPS) {
        const SubtreeDev st = trees[sidx];
        for (int i = lane; i < st.ncols; i += 32) xs[i] = zv[st.cbeg + i];
        for (int i = lane; i < st.next; i += 32) xs[st.ncols + i] = xp[rows_all[st.root_rows + i]];
        __syncwarp();
        const uchar2* pu = pu_all + st.pu_beg;
        // offsets one past the LAST front: the walk runs from the root down to the leaves
        const double* U = fac + st.Ubeg + st.Ucount;
        const double* D = dpack + st.Dbeg + st.Dcount;
        int to = 0;
        for (int fr = lane; fr < st.nfr; fr += 32) to += pu[fr].y;
        for (int off = 16; off > 0; off >>= 1) to += __shfl_xor_sync(0xffffffffu, to, off);
        const unsigned short* tg = tgt_all + st.tgt_beg + to;
        int c0 = st.ncols;
        uchar2 q = pu[st.nfr - 1];
        for (int fr = st.nfr - 1; fr >= 0; fr--) {
            const int p = q.x, u = q.y;
            if (fr > 0) q = pu[fr - 1];
            U -= sub_round4(u * p), D -= sub_round4(p * p), tg -= u, c0 -= p;
            // ---- every global load of a front with p <= 8 (the common case) is requested here, before the first dependent
            //      use: target indices, the first eight columns of the U panel, the last eight columns of U11 and its diagonal
            const int mlast = (p - 1) & ~7;
            int t0 = 0, t1 = 0, t2 = 0;
            if (lane < u) t0 = tg[lane];
            if (u > 32) { // warp-uniform
                if (lane + 32 < u) t1 = tg[lane + 32];
                if (lane + 64 < u) t2 = tg[lane + 64];
            }
            double d[8], v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = ldg_if(U + lane + k * u, lane < u && k < p);
#pragma unroll
            for (int m = 0; m < 8; m++) d[m] = ldg_if(D + lane + (mlast + m) * p, mlast + m < p && lane < mlast + m);
            double dg = ldg_if(D + lane + lane * p, lane < p);
            if (lane >= p) dg = 1.0;
            double x2a = 0.0, x2b = 0.0, x2c = 0.0;
            if (lane < u) x2a = xs[t0];
            if (u > 32) {
                if (lane + 32 < u) x2b = xs[t1];
                if (lane + 64 < u) x2c = xs[t2];
            }
            double tv = lane < p ? xs[c0 + lane] : 0.0;
            // ---- t = z1 - U12 x2: lanes = rows of the U panel (coalesced), eight columns at a time, one 9-shuffle reduction
            for (int k0 = 0; k0 < p; k0 += 8) {
                if (k0 > 0) {
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = ldg_if(U + lane + (k0 + k) * u, lane < u && k0 + k < p);
                }
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] *= x2a;
                if (u > 32) {
                    double vb[8], vc[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) vb[k] = ldg_if(U + lane + 32 + (k0 + k) * u, lane + 32 < u && k0 + k < p);
#pragma unroll
                    for (int k = 0; k < 8; k++) vc[k] = ldg_if(U + lane + 64 + (k0 + k) * u, lane + 64 < u && k0 + k < p);
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] += vb[k] * x2b + vc[k] * x2c;
                }
                const double tot = warp_reduce8(v, lane);
                // lane k0 + c takes the total of column c
                const double mine = __shfl_sync(0xffffffffu, tot, reduce8_lane_of((lane - k0) & 7));
                if (lane >= k0 && lane < k0 + 8 && lane < p && u > 0) tv -= mine;
            }
            // ---- x1 = inv(U11) t by backward substitution: lane = row, eight columns of U11 at a time (last chunk first)
            const double rd = __drcp_rn(dg);
            for (int m0 = mlast; m0 >= 0; m0 -= 8) {
                if (m0 != mlast) {
#pragma unroll
                    for (int m = 0; m < 8; m++) d[m] = ldg_if(D + lane + (m0 + m) * p, lane < m0 + m);
                }
#pragma unroll
                for (int m = 7; m >= 0; m--) {
                    const int mm = m0 + m;
                    const double xm = __shfl_sync(0xffffffffu, tv * rd, mm & 31); // lane mm: its row is complete
                    if (mm < p) {
                        if (lane == mm) tv = xm;
                        tv -= d[m] * xm; // (d[m] = 0 for the lanes at or below the diagonal)
                    }
                }
            }
            __syncwarp();
            if (lane < p) xs[c0 + lane] = tv;
            __syncwarp();
        }
        for (int i = lane; i < st.ncols; i += 32) xp[st.cbeg + i] = xs[i];
        __syncwarp();
    }
}

// post-pass of the factorization: the pivot blocks L11\U11 of the subtree fronts (top p rows of the f x p L panels), packed
// p x p per front into the pivot-block arena, where the other fronts keep their explicit inverses
__global__ void __launch_bounds__(256) k_pack_pivot_blocks(const int* __restrict__ nodelist, int nn, const NodeDev* __restrict__ nodes,
                                                           const double* __restrict__ fac, double* __restrict__ dinv) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    for (int e = warp; e < nn; e += nw) {
        const NodeDev nd = nodes[nodelist[e]];
        const int p = nd.p, f = nd.p + nd.u;
        const double* L = fac + nd.Loff;
        double* D = dinv + nd.Doff;
        for (int idx = lane; idx < p * p; idx += 32) {
            const int i = idx % p, m = idx / p;
            D[idx] = L[i + m * f];
        }
    }
}

} // namespace b200
