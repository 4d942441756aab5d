// sweep_chain.cuh -- triangular solves over the TOP of the front tree: pipelined supernode chains, one CTA per 64-row block.
//
// Role in the reference: the inside of umfpack_di_solve / cudssExecute(SOLVE)
// (russell_sparse/c_code/interface_umfpack.c:229, interface_cudss.cu:530) for the large separator fronts.
//
// A wide supernode is stored as a CHAIN of panels (<= 64 pivots each), every panel a front whose update set is the next
// panel's front.  The earlier persistent kernels (sweep_top.cuh) treated every (panel, 128-row slice) as an item: each panel
// of a chain was one cross-CTA hop, and the per-item trace showed ~3.0 us (forward) / ~5.2 us (backward) per hop -- of which
// only ~0.4 us is the hop itself (LL lines) and the rest barriers, gathers and two dependent GEMVs that every slice repeated.
// 57 chain links at config 2 = 0.5 of the 0.8 ms sweep.
//
// Here the unit of work is a ROW BLOCK of a chain: block j < K owns the pivot rows of panel j, the blocks after them own 64
// update rows each, and a block keeps its rows of the right-hand side in shared memory for the whole chain.
//   forward : block j applies panels 0 .. min(j, K)-1 to its rows as their z arrive (z travels as LL lines, one 16-byte line
//             per entry, value and epoch tag in one store), then -- if it is a pivot block -- solves its 64 x 64 triangle and
//             publishes z_j.  The dependent path per panel is: z_{j-1} lands -> one 64 x 64 GEMV -> one triangular GEMV ->
//             publish; everything else (the other blocks' updates) runs beside it, the classic look-ahead of a dense
//             triangular solve.  Update blocks publish their final rows as LL lines for the parent chain.
//   backward: an update block knows its x (ancestors' columns) from the start and publishes its partial dot products for ALL
//             panels at streaming speed; pivot block j sums the partials of the blocks after it (fixed order), solves its
//             triangle, publishes x_j, then contributes to the panels before it.  Dependent path per panel: x_{j+1} -> one
//             64 x 64 GEMV -> publish -> sum -> triangular GEMV.
// Items are handed out by atomic ticket in dependency order (see sweep_top.cuh for the progress argument); panel slices are
// double-buffered in shared memory with cp.async so the HBM stream runs ahead of the dependency wave.
#pragma once
#include "sweep_top.cuh"

namespace b200 {

#define B200_CH_B 64 // rows per block = pivots per panel (B200_MAXP)

struct ChainDev {
    int panel_ptr, K;       // first entry in the panel table, number of panels
    int P, U;               // pivots of the whole chain, update rows (of its last panel)
    int nblocks;            // K + ceil(U / 64)
    int last_node;          // front of the last panel (its update-row list names the chain's update rows)
    long long wll_off;      // first LL line of the chain's update vector (U lines)
    long long pbase;        // first 64-line group of the chain's partial dot products: group (k * nblocks + jb)
};
struct ChainPanel {
    int node, off; // front of this panel, its first row inside the chain's row space
};
struct ChainItem {
    int chain, block, row0, nrows; // rows [row0, row0 + nrows) of the chain's row space
    int rng, nch;                  // forward: child records (8 ints each) of the chain's first front that touch these rows
};
// child record: [0] a, [1] b (positions in the child's update list), [2],[3] offset of that list in rel[] / wv[] (int64),
// [4],[5] first LL line of the child's update vector (int64, -1: the child lies below the region, plain wv), [6],[7] unused
#define B200_CH_REC 8

#define B200_CHF_SMEM ((size_t)2 * B200_CH_B * B200_CH_B * sizeof(double))
#define B200_CHB_LD (B200_CH_B + 1)
#define B200_CHB_SMEM ((size_t)2 * B200_CH_B * B200_CHB_LD * sizeof(double))

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait_n() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__global__ void __launch_bounds__(256, 3) k_fwd_chain(const ChainItem* __restrict__ items, int nitems, const ChainDev* __restrict__ chains,
                                                   const ChainPanel* __restrict__ panels, const NodeDev* __restrict__ nodes,
                                                   const int* __restrict__ rel_all, const double* __restrict__ fac,
                                                   const double* __restrict__ dinv, const int* __restrict__ lperm,
                                                   const int* __restrict__ ranges, const double* __restrict__ y, double* __restrict__ zv,
                                                   const double* __restrict__ wv, ulonglong2* __restrict__ wll, ulonglong2* __restrict__ zll,
                                                   int* __restrict__ epoch_ptr, int* __restrict__ abort_flag,
                                                   unsigned long long* __restrict__ trace) {
    const unsigned epoch = (unsigned)*(volatile int*)epoch_ptr;
    extern __shared__ double smt[]; // two panel slices: Ls[buf][kk * 64 + r]
    __shared__ double w[B200_CH_B], zs[B200_CH_B], part[4][B200_CH_B];
    __shared__ int s_next;
    const int tid = threadIdx.x;
    const int r = tid & (B200_CH_B - 1), q = tid >> 6; // GEMV layout of the panel updates: row r, columns kk = q (mod 4)
    const int gk = tid >> 2, gpart = tid & 3;           // GEMV layout of the triangular solve: four threads per row
    int* ticket = epoch_ptr + 1;
    if (tid == 0) s_next = atomicAdd(ticket, 1);
    __syncthreads();
    for (int itx = s_next; itx < nitems; itx = s_next) {
        int nxt = 0;
        if (tid == 0) nxt = atomicAdd(ticket, 1); // the next item's ticket is fetched underneath this item's work
        const ChainItem it = items[itx];
        const ChainDev ch = chains[it.chain];
        const int j = it.block, nr = it.nrows;
        const bool piv = j < ch.K;
        const int niter = piv ? j : ch.K;
        const ChainPanel* pn = panels + ch.panel_ptr;
        if (trace && tid == 0) trace[4 * (long long)itx] = gtime();
        // slice of panel k for my rows -> buffer (k & 1)
        auto stage = [&](const int k, const NodeDev& ndk, const int offk) {
            double* dst = smt + (k & 1) * (B200_CH_B * B200_CH_B);
            const int fk = ndk.p + ndk.u;
            if (r < nr) {
                const double* src = fac + ndk.Loff + (it.row0 - offk) + r;
                for (int kk = q; kk < ndk.p; kk += 4) cp_async8(dst + kk * B200_CH_B + r, src + (long long)kk * fk);
            }
            cp_async_commit();
        };
        NodeDev ndk;
        int offk = 0;
        if (niter > 0) {
            const ChainPanel p0 = pn[0];
            ndk = nodes[p0.node], offk = p0.off;
            stage(0, ndk, offk);
        }
        // my own panel: rows of inv(L11) in registers, local pivot permutation, right-hand side
        int c0j = 0, lp = 0;
        double dreg[16];
#pragma unroll
        for (int m = 0; m < 16; m++) dreg[m] = 0.0;
        if (piv) {
            const NodeDev ndj = nodes[pn[j].node];
            c0j = ndj.c0;
            const double* D = dinv + ndj.Doff;
#pragma unroll
            for (int m4 = 0; m4 < 16; m4++) {
                const int m = gpart + 4 * m4;
                if (gk < nr && m < gk) dreg[m4] = D[gk + (long long)m * nr];
            }
            if (tid < nr) lp = lperm[c0j + tid];
        }
        if (tid < B200_CH_B) w[tid] = (piv && tid < nr) ? y[c0j + tid] : 0.0;
        __syncthreads();
        // ---- contributions of the children of the chain's first front (complete long before, as a rule)
        int ok = 1;
        for (int e = 0; e < it.nch; e++) {
            const int* rg = ranges + it.rng + B200_CH_REC * e;
            const int a = rg[0], b = rg[1];
            const long long wofs = ((long long)rg[3] << 32) | (long long)(unsigned)rg[2];
            const long long llo = ((long long)rg[5] << 32) | (long long)(unsigned)rg[4];
            if (llo >= 0) { // one thread waits politely for the first line, then everybody fetches
                if (tid == 0) {
                    double dummy;
                    ok &= ll_wait(wll + llo + a, epoch, dummy, abort_flag, true) ? 1 : 0;
                }
                if (!__syncthreads_and(ok)) return;
            }
            for (int i = a + tid; i < b; i += 256) { // (b - a <= 64: the rows of one child are distinct)
                double val;
                if (llo >= 0) ok &= ll_wait(wll + llo + i, epoch, val, abort_flag, false) ? 1 : 0;
                else val = __ldcg(wv + wofs + i);
                w[rel_all[wofs + i] - it.row0] += val;
            }
            __syncthreads(); // one child after the other: fixed order of additions
        }
        // ---- panels before this block, as their z arrive
        for (int k = 0; k < niter; k++) {
            const NodeDev ndc = ndk; // panel k
            if (k + 1 < niter) {
                const ChainPanel pnx = pn[k + 1];
                ndk = nodes[pnx.node], offk = pnx.off;
                stage(k + 1, ndk, offk);
            }
            const int pk = ndc.p;
            if (tid == 0) {
                double dummy;
                ok &= ll_wait(zll + ndc.c0, epoch, dummy, abort_flag, true) ? 1 : 0;
            }
            if (!__syncthreads_and(ok)) return;
            if (trace && tid == 0 && k == niter - 1) trace[4 * (long long)itx + 1] = gtime();
            if (tid < pk) {
                double v;
                ok &= ll_wait(zll + ndc.c0 + tid, epoch, v, abort_flag, false) ? 1 : 0;
                zs[tid] = v;
            }
            if (k + 1 < niter) cp_async_wait_n<1>();
            else cp_async_wait_n<0>();
            __syncthreads();
            {
                const double* Ls = smt + (k & 1) * (B200_CH_B * B200_CH_B);
                double s0 = 0.0, s1 = 0.0;
                int kk = q;
                for (; kk + 4 < pk; kk += 8) s0 += Ls[kk * B200_CH_B + r] * zs[kk], s1 += Ls[(kk + 4) * B200_CH_B + r] * zs[kk + 4];
                if (kk < pk) s0 += Ls[kk * B200_CH_B + r] * zs[kk];
                part[q][r] = s0 + s1;
            }
            __syncthreads();
            if (tid < nr) w[tid] -= (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
            __syncthreads(); // zs, part and the slice buffer are reused
        }
        if (!__syncthreads_and(ok)) return;
        if (piv) {
            if (trace && tid == 0 && niter == 0) trace[4 * (long long)itx + 1] = gtime();
            const double tp = tid < nr ? w[lp] : 0.0;
            __syncthreads();
            if (tid < nr) w[tid] = tp;
            __syncthreads();
            double s = 0.0;
#pragma unroll
            for (int m4 = 0; m4 < 16; m4++) {
                const int m = gpart + 4 * m4;
                if (gk < nr && m < gk) s += dreg[m4] * w[m];
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (gk < nr && gpart == 0) {
                s += w[gk];
                ll_store(zll + c0j + gk, epoch, s); // the blocks after this one may go on
                zv[c0j + gk] = s;
            }
        } else if (tid < nr) {
            ll_store(wll + ch.wll_off + (it.row0 - ch.P) + tid, epoch, w[tid]); // final rows of the chain's update vector
        }
        if (trace && tid == 0) trace[4 * (long long)itx + 2] = gtime();
        if (tid == 0) s_next = nxt;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256, 3) k_bwd_chain(const ChainItem* __restrict__ items, int nitems, const ChainDev* __restrict__ chains,
                                                   const ChainPanel* __restrict__ panels, const NodeDev* __restrict__ nodes,
                                                   const int* __restrict__ rows_all, const double* __restrict__ fac,
                                                   const double* __restrict__ dinv, const double* __restrict__ zv, double* __restrict__ xp,
                                                   ulonglong2* __restrict__ xll, ulonglong2* __restrict__ pll, int* __restrict__ epoch_ptr,
                                                   int* __restrict__ abort_flag, unsigned long long* __restrict__ trace) {
    const unsigned epoch = (unsigned)*(volatile int*)epoch_ptr;
    extern __shared__ double smt[]; // two U-panel slices: Us[buf][i * 65 + r]
    __shared__ double xs[B200_CH_B], t[B200_CH_B], part[4][B200_CH_B];
    __shared__ int s_next;
    const int tid = threadIdx.x;
    const int i = tid & (B200_CH_B - 1), q = tid >> 6; // dot-product layout: pivot i, rows r = q (mod 4)
    const int gk = tid >> 2, gpart = tid & 3;
    int* ticket = epoch_ptr + 2;
    if (tid == 0) s_next = nitems - 1 - atomicAdd(ticket, 1);
    __syncthreads();
    for (int itx = s_next; itx >= 0; itx = s_next) {
        int nxt = 0;
        if (tid == 0) nxt = nitems - 1 - atomicAdd(ticket, 1);
        const ChainItem it = items[itx];
        const ChainDev ch = chains[it.chain];
        const int j = it.block, nr = it.nrows;
        const bool piv = j < ch.K;
        const int niter = piv ? j : ch.K;
        const ChainPanel* pn = panels + ch.panel_ptr;
        if (trace && tid == 0) trace[4 * (long long)itx] = gtime();
        // rows [row0, row0 + nr) of panel k's U panel (u_k x p_k, row jj = chain row off_k + p_k + jj) -> buffer (k & 1)
        auto stage = [&](const int k, const NodeDev& ndk, const int offk) {
            double* dst = smt + (k & 1) * (B200_CH_B * B200_CHB_LD);
            const int rr = tid & (B200_CH_B - 1);
            if (rr < nr) {
                const double* src = fac + ndk.Uoff + (it.row0 - offk - ndk.p) + rr;
                for (int ii = q; ii < ndk.p; ii += 4) cp_async8(dst + ii * B200_CHB_LD + rr, src + (long long)ii * ndk.u);
            }
            cp_async_commit();
        };
        NodeDev ndk;
        int offk = 0;
        if (niter > 0) {
            const ChainPanel p0 = pn[niter - 1];
            ndk = nodes[p0.node], offk = p0.off;
            stage(niter - 1, ndk, offk);
        }
        int ok = 1;
        if (!piv) {
            // update block: its x are solution entries of ancestor columns
            int col = -1;
            if (tid < nr) col = rows_all[nodes[ch.last_node].rows_ptr + (it.row0 - ch.P) + tid];
            if (tid == 0) {
                double dummy;
                ok &= ll_wait(xll + col, epoch, dummy, abort_flag, true) ? 1 : 0;
            }
            if (!__syncthreads_and(ok)) return;
            if (tid < nr) {
                double v;
                ok &= ll_wait(xll + col, epoch, v, abort_flag, false) ? 1 : 0;
                xs[tid] = v;
            }
            if (trace && tid == 0) trace[4 * (long long)itx + 1] = gtime();
        } else {
            // pivot block: x_j = inv(U11) (z_j - sum of the partial dot products of the blocks after it)
            const NodeDev ndj = nodes[pn[j].node];
            const int c0j = ndj.c0;
            double dreg[16];
            {
                const double* D = dinv + ndj.Doff;
#pragma unroll
                for (int m4 = 0; m4 < 16; m4++) {
                    const int m = gk + gpart + 4 * m4;
                    dreg[m4] = (gk < nr && m < nr) ? D[gk + (long long)m * nr] : 0.0;
                }
            }
            const double zj = tid < nr ? zv[c0j + tid] : 0.0;
            const ulonglong2* grp = pll + (ch.pbase + (long long)j * ch.nblocks) * B200_CH_B;
            if (j + 1 < ch.nblocks) { // the block right after this one publishes last
                if (tid == 0) {
                    double dummy;
                    ok &= ll_wait(grp + (long long)(j + 1) * B200_CH_B, epoch, dummy, abort_flag, true) ? 1 : 0;
                }
                if (!__syncthreads_and(ok)) return;
            }
            if (trace && tid == 0) trace[4 * (long long)itx + 1] = gtime();
            double acc = 0.0;
            if (i < nr)
                for (int jb = j + 1 + q; jb < ch.nblocks; jb += 4) { // fixed order per thread, fixed order of the four threads below
                    double v;
                    ok &= ll_wait(grp + (long long)jb * B200_CH_B + i, epoch, v, abort_flag, false) ? 1 : 0;
                    acc += v;
                }
            part[q][i] = acc;
            __syncthreads();
            if (tid < nr) t[tid] = zj - ((part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]));
            __syncthreads();
            double s = 0.0;
#pragma unroll
            for (int m4 = 0; m4 < 16; m4++) {
                const int m = gk + gpart + 4 * m4;
                if (gk < nr && m < nr) s += dreg[m4] * t[m];
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (gk < nr && gpart == 0) {
                xs[gk] = s;
                ll_store(xll + c0j + gk, epoch, s); // for the update blocks of the chains below
                xp[c0j + gk] = s;                   // for the subtree kernels and the caller
            }
        }
        if (!__syncthreads_and(ok)) return;
        // ---- partial dot products of my rows for the panels before this block (last panel first)
        for (int k = niter - 1; k >= 0; k--) {
            const NodeDev ndc = ndk;
            if (k > 0) {
                const ChainPanel pnx = pn[k - 1];
                ndk = nodes[pnx.node], offk = pnx.off;
                stage(k - 1, ndk, offk);
                cp_async_wait_n<1>();
            } else cp_async_wait_n<0>();
            __syncthreads();
            const int pk = ndc.p;
            {
                const double* Us = smt + (k & 1) * (B200_CH_B * B200_CHB_LD) + i * B200_CHB_LD;
                double s0 = 0.0, s1 = 0.0;
                if (i < pk) {
                    int rr = q;
                    for (; rr + 4 < nr; rr += 8) s0 += Us[rr] * xs[rr], s1 += Us[rr + 4] * xs[rr + 4];
                    if (rr < nr) s0 += Us[rr] * xs[rr];
                }
                part[q][i] = s0 + s1;
            }
            __syncthreads();
            if (tid < pk)
                ll_store(pll + (ch.pbase + (long long)k * ch.nblocks + j) * B200_CH_B + tid, epoch,
                         (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]));
            __syncthreads(); // part and the slice buffer are reused
        }
        if (trace && tid == 0) trace[4 * (long long)itx + 2] = gtime();
        if (tid == 0) s_next = nxt;
        __syncthreads();
    }
}

} // namespace b200
