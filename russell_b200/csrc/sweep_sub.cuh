// sweep_sub.cuh -- triangular solves over the BOTTOM of the front tree: one CTA per small subtree, everything in shared memory.
//
// Role in the reference: the inside of umfpack_di_solve / cudssExecute(SOLVE)
// (russell_sparse/c_code/interface_umfpack.c:229, interface_cudss.cu:530) for the ~95 % of the fronts that are tiny.
//
// The first version of these kernels (k_fwd_subtree / k_bwd_subtree, round 1) walked a subtree front by front with the
// multifrontal vector scheme: every front gathered its children's update vectors through relative indices, built its own
// update vector and wrote it back to global memory.  ncu showed them bound by instruction issue (~900 warp instructions per
// front for ~370 stored entries) and, in the backward direction, by p short column segments per front read out of the f x p
// L panels (1.42x the algorithmic DRAM traffic).  This version is right-looking on a shared-memory segment of the solution
// vector instead:
//   * the columns of a subtree are contiguous (postorder) -> xs[0 .. ncols) holds them, xs[ncols .. ncols + u_root) holds the
//     update rows of the subtree's root (everything that leaves the subtree); a host-built 16-bit target index per update
//     row says where a row of any front of the subtree lives in xs.  No child lists, no relative indices, no per-front
//     update vectors, no global traffic between the fronts of a subtree;
//   * the host analysis lays the subtree's L panels out contiguously, then its U panels; a post-pass of the factorization
//     packs the pivot blocks (needed by both directions) next to each other.  ONE bulk-copy transaction per array
//     (cp.async.bulk = TMA, completion on an mbarrier) stages what a direction needs: the forward kernel reads the L
//     region, the backward kernel the U region + the packed pivot blocks -- exactly the algorithmic bytes, in full lines;
//   * per front: two block barriers, a p-step substitution by shuffles inside warp 0, and one multiply-add per stored entry.
#pragma once
#include "kernels.cuh"

namespace b200 {

struct SubtreeDev {          // 80 bytes, built once by the host
    long long Lbeg, Ubeg, Dbeg; // first double of the subtree's L region / U region in fac, of its pivot-block copies in dinv
    long long tgt_beg;          // first entry of its target indices (uint16) -- multiple of 8 entries (16 bytes)
    long long root_rows;        // offset of the root's update-row list in rows[] / wv[]
    int Lcount, Ucount, Dcount; // doubles (multiples of 4)
    int tgt_count;              // uint16 entries, padded to a multiple of 8
    int pu_beg;                 // first (p, u) pair of its fronts inside the pu array (multiple of 8 pairs = 16 bytes)
    int nfr;                    // fronts
    int cbeg, ncols;            // first column and number of columns
    int next;                   // update rows of the root
    int pan;                    // doubles reserved for the panel region of this subtree's CTA: max(Lcount, Ucount + Dcount)
};

#define B200_SUB_THREADS 64

__device__ __forceinline__ void sub_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void sub_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ int sub_round4(int x) { return (x + 3) & ~3; }

// dynamic shared memory of one CTA: panels (pan doubles) | xs (ncols + next, rounded up to even) | tgt (tgt_count uint16) |
// pu (nfr pairs, rounded up to 8) | lps (ncols bytes); every region starts 16-byte aligned.  The launch reserves the largest
// total over all subtrees (sub_smem_bytes on the host).
__host__ __device__ __forceinline__ int sub_xs_doubles(int ncols, int next) { return (ncols + next + 1) & ~1; }
__host__ __device__ __forceinline__ size_t sub_smem_bytes(int pan, int ncols, int next, int tgt_count, int nfr) {
    return (size_t)pan * 8 + (size_t)sub_xs_doubles(ncols, next) * 8 + (size_t)tgt_count * 2 + (size_t)((nfr + 7) & ~7) * 2 + (size_t)((ncols + 15) & ~15);
}

// forward:  z = L^{-1} P y  on the subtree's columns; the root's update vector goes to wv (read by the front above)
__global__ void __launch_bounds__(B200_SUB_THREADS) k_fwd_stree(const SubtreeDev* __restrict__ trees, const double* __restrict__ fac,
                                                                const unsigned short* __restrict__ tgt_all, const uchar2* __restrict__ pu_all,
                                                                const int* __restrict__ lperm, const double* __restrict__ y,
                                                                double* __restrict__ zv, double* __restrict__ wv) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ __align__(8) unsigned long long bar;
    const SubtreeDev st = trees[blockIdx.x];
    double* Ls = reinterpret_cast<double*>(smraw);
    double* xs = Ls + st.pan;
    unsigned short* tg = reinterpret_cast<unsigned short*>(xs + sub_xs_doubles(st.ncols, st.next));
    uchar2* pu = reinterpret_cast<uchar2*>(tg + st.tgt_count);
    unsigned char* lps = reinterpret_cast<unsigned char*>(pu + ((st.nfr + 7) & ~7));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned nb_pu = (unsigned)(((st.nfr + 7) & ~7) * 2);
        sub_expect_tx(&bar, (unsigned)st.Lcount * 8u + (unsigned)st.tgt_count * 2u + nb_pu);
        sub_bulk_g2s(Ls, fac + st.Lbeg, (unsigned)st.Lcount * 8u, &bar);
        if (st.tgt_count > 0) sub_bulk_g2s(tg, tgt_all + st.tgt_beg, (unsigned)st.tgt_count * 2u, &bar);
        sub_bulk_g2s(pu, pu_all + st.pu_beg, nb_pu, &bar);
    }
    // the subtree's segment of the right-hand side and of the local pivot permutations, underneath the bulk copies
    for (int i = tid; i < st.ncols; i += B200_SUB_THREADS) xs[i] = y[st.cbeg + i], lps[i] = (unsigned char)lperm[st.cbeg + i];
    for (int i = tid; i < st.next; i += B200_SUB_THREADS) xs[st.ncols + i] = 0.0;
    __syncthreads(); // (also publishes the mbarrier initialisation)
    mbar_wait(&bar, 0);
    int Lo = 0, c0 = 0, to = 0;
    for (int fr = 0; fr < st.nfr; fr++) {
        const uchar2 q = pu[fr];
        const int p = q.x, u = q.y, f = p + u;
        const double* L = Ls + Lo;
        if (warp == 0) { // z1 = inv(L11) P t1 by forward substitution: lane = row, L11 unit lower triangular
            double tv = lane < p ? xs[c0 + lps[c0 + lane]] : 0.0;
            for (int m = 0; m + 1 < p; m++) {
                const double zm = __shfl_sync(0xffffffffu, tv, m);
                if (lane > m && lane < p) tv -= L[lane + m * f] * zm;
            }
            __syncwarp();
            if (lane < p) xs[c0 + lane] = tv;
        }
        __syncthreads();
        // update rows: one thread per row (u <= 96: at most two rows per thread), one multiply-add per stored entry
        for (int i = tid; i < u; i += B200_SUB_THREADS) {
            const double* row = L + p + i;
            double a0 = 0.0, a1 = 0.0;
            int k = 0;
            for (; k + 1 < p; k += 2) a0 += row[k * f] * xs[c0 + k], a1 += row[(k + 1) * f] * xs[c0 + k + 1];
            if (k < p) a0 += row[k * f] * xs[c0 + k];
            xs[tg[to + i]] -= a0 + a1; // the rows of one front are distinct: no conflicting updates
        }
        __syncthreads();
        Lo += sub_round4(f * p), c0 += p, to += u;
    }
    for (int i = tid; i < st.ncols; i += B200_SUB_THREADS) zv[st.cbeg + i] = xs[i];
    for (int i = tid; i < st.next; i += B200_SUB_THREADS) wv[st.root_rows + i] = xs[st.ncols + i];
}

// backward:  x = U^{-1} z  on the subtree's columns, given the solution at the root's update rows (fronts above)
__global__ void __launch_bounds__(B200_SUB_THREADS) k_bwd_stree(const SubtreeDev* __restrict__ trees, const double* __restrict__ fac,
                                                                const double* __restrict__ dpack, const unsigned short* __restrict__ tgt_all,
                                                                const uchar2* __restrict__ pu_all, const int* __restrict__ rows_all,
                                                                const double* __restrict__ zv, double* __restrict__ xp) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ double red[8][32];
    const SubtreeDev st = trees[blockIdx.x];
    double* Us = reinterpret_cast<double*>(smraw); // U region, then the packed pivot blocks
    double* xs = Us + st.pan;
    unsigned short* tg = reinterpret_cast<unsigned short*>(xs + sub_xs_doubles(st.ncols, st.next));
    uchar2* pu = reinterpret_cast<uchar2*>(tg + st.tgt_count);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* Ds = Us + st.Ucount;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned nb_pu = (unsigned)(((st.nfr + 7) & ~7) * 2);
        sub_expect_tx(&bar, (unsigned)(st.Ucount + st.Dcount) * 8u + (unsigned)st.tgt_count * 2u + nb_pu);
        if (st.Ucount > 0) sub_bulk_g2s(Us, fac + st.Ubeg, (unsigned)st.Ucount * 8u, &bar);
        sub_bulk_g2s(Ds, dpack + st.Dbeg, (unsigned)st.Dcount * 8u, &bar);
        if (st.tgt_count > 0) sub_bulk_g2s(tg, tgt_all + st.tgt_beg, (unsigned)st.tgt_count * 2u, &bar);
        sub_bulk_g2s(pu, pu_all + st.pu_beg, nb_pu, &bar);
    }
    for (int i = tid; i < st.ncols; i += B200_SUB_THREADS) xs[i] = zv[st.cbeg + i];
    for (int i = tid; i < st.next; i += B200_SUB_THREADS) xs[st.ncols + i] = xp[rows_all[st.root_rows + i]];
    __syncthreads();
    mbar_wait(&bar, 0);
    // offsets of the LAST front: the walk runs from the root down to the leaves
    int Uo = st.Ucount, Do = st.Dcount, c0 = st.ncols, to = 0;
    for (int fr = 0; fr < st.nfr; fr++) to += pu[fr].y; // (a few hundred fronts at most; every thread keeps its own copy)
    for (int fr = st.nfr - 1; fr >= 0; fr--) {
        const uchar2 q = pu[fr];
        const int p = q.x, u = q.y;
        Uo -= sub_round4(u * p), Do -= sub_round4(p * p), c0 -= p, to -= u;
        const double* U = Us + Uo; // u x p, column-major: U[j + k*u] = U12(k, j)
        const double* D = Ds + Do; // p x p pivot block: upper triangle incl. diagonal = U11
        // t = z1 - U12 x2: thread (k, part) sums the entries j = part, part + nparts, ... of column k.  Lanes run over k, so a
        // plain j would make them stride u doubles through shared memory (a 32-way bank conflict for u = 16): every column
        // starts its walk at a different row (skew) so that the lane stride u + skew is odd
        int p2 = 1;
        while (p2 < p) p2 <<= 1;
        const int nparts = min(B200_SUB_THREADS / p2, 8);
        const int k = tid & (p2 - 1), part = tid / p2;
        if (k < p && part < nparts && u > 0) {
            const int skew = (u & 1) ? 0 : 1;
            double acc = 0.0;
            int jj = (part + k * skew) % u;
            for (int it = part; it < u; it += nparts) {
                acc += U[jj + k * u] * xs[tg[to + jj]];
                jj += nparts;
                if (jj >= u) jj -= u;
            }
            red[part][k] = acc;
        }
        __syncthreads();
        if (warp == 0) { // x1 = inv(U11) t by backward substitution: lane = row
            double tv = 0.0;
            if (lane < p) {
                tv = xs[c0 + lane];
                if (u > 0)
                    for (int q2 = 0; q2 < nparts; q2++) tv -= red[q2][lane]; // fixed order: deterministic
            }
            const double rd = lane < p ? __drcp_rn(D[lane + lane * p]) : 0.0;
            for (int m = p - 1; m >= 0; m--) {
                const double xm = __shfl_sync(0xffffffffu, tv * rd, m); // lane m: its row is complete
                if (lane == m) tv = xm;
                if (lane < m) tv -= D[lane + m * p] * xm;
            }
            if (lane < p) xs[c0 + lane] = tv;
        }
        __syncthreads();
    }
    for (int i = tid; i < st.ncols; i += B200_SUB_THREADS) xp[st.cbeg + i] = xs[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// v2 (default): ONE WARP per subtree, panels straight from global memory.
// The bulk-staged kernels above keep a whole subtree (24 KB on average, 45 KB at most) in shared memory: 4-5 CTAs = 8-10
// warps per SM, and every front is a chain of dependent shared-memory round trips, two block barriers and a p-step
// substitution -- measured 1.10 ms per sweep at config 2 against 0.92 ms for the round-1 kernels; a smaller subtree budget
// (more CTAs per SM) was faster, i.e. the walk is bound by the number of fronts in flight per SM, not by bytes.  Here a
// subtree needs only its solution segment xs in shared memory (<= 8 KB), so 28 warps = 28 subtrees are resident per SM; a
// warp issues all loads of a front at once (coalesced: lanes = rows), the other warps hide their latency, and nothing but
// __syncwarp orders the walk.  Per front and warp: ~100 instructions (the round-1 kernels: ~900 on two warps).
// ---------------------------------------------------------------------------------------------------------------------
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
