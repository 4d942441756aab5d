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
#include "sweep_sub.cuh"

namespace b200 {

#define B200_CH_B 64 // rows per block = pivots per panel (B200_MAXP)

// one panel of a chain, self-contained (no front-descriptor lookup on the way)
struct ChainPanel {
    long long Loff, Uoff; // its L panel (f x p) and U panel (u x p) in fac
    int off;              // its first row inside the chain's row space
    int p, f, u, c0;      // pivots, front order, update rows, first column
    int pad;
};
// one row block of a chain: everything the kernels need in ONE record
struct ChainItem {
    int chain, block, row0, nrows; // rows [row0, row0 + nrows) of the chain's row space
    int rng, nch;                  // forward: child records (8 ints each) of the chain's first front that touch these rows
    int K, nblocks;                // panels and row blocks of the chain (block < K: pivot block of panel `block`)
    int panel_ptr, c0j;            // first entry of the chain in the panel table; first column of a pivot block
    long long Doff;                // pivot block: its inverse pair in dinv
    long long out;                 // update block: first LL line of its rows in wll (forward), offset of its rows in rows[] (backward)
    long long rows_off;
    long long pgrp;                // first 64-line group of the chain's partial dot products: group (k * nblocks + jb)
};
// child record: [0] a, [1] b (positions in the child's update list), [2],[3] offset of that list in rel[] / wv[] (int64),
// [4],[5] first LL line of the child's update vector (int64, -1: the child lies below the region, plain wv), [6],[7] unused
#define B200_CH_REC 8

// non-blocking read of one LL line
__device__ __forceinline__ void ll_peek(const ulonglong2* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ bool ll_decode(const unsigned long long a, const unsigned long long b, const unsigned epoch, double& out) {
    if ((unsigned)(a >> 32) != epoch || (unsigned)(b >> 32) != epoch) return false;
    out = __longlong_as_double((long long)((b << 32) | (a & 0xffffffffull)));
    return true;
}

// sums v[0..15] over the 32 lanes with 16 shuffles: afterwards lane l holds the total of column
// 8 * bit4(l) + 4 * bit3(l) + 2 * bit2(l) + bit1(l) (both lanes of a pair hold it).  Fixed order: deterministic.
__device__ __forceinline__ double warp_reduce16(double (&v)[16], const int lane) {
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const double send = h16 ? v[c] : v[c + 8];
        const double keep = h16 ? v[c + 8] : v[c];
        v[c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const double send = h8 ? v[c] : v[c + 4];
        const double keep = h8 ? v[c + 4] : v[c];
        v[c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const double send = h4 ? v[c] : v[c + 2];
        const double keep = h4 ? v[c + 2] : v[c];
        v[c] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const double send = h2 ? v[0] : v[1];
        const double keep = h2 ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

__global__ void __launch_bounds__(256, 2) k_fwd_chain(const ChainItem* __restrict__ items, int nitems, const ChainPanel* __restrict__ panels,
                                                   const int* __restrict__ rel_all, const double* __restrict__ fac,
                                                   const double* __restrict__ dinv, const int* __restrict__ lperm,
                                                   const int* __restrict__ ranges, const double* __restrict__ y, double* __restrict__ zv,
                                                   const double* __restrict__ wv, ulonglong2* __restrict__ wll, ulonglong2* __restrict__ zll,
                                                   int* __restrict__ epoch_ptr, int* __restrict__ abort_flag,
                                                   unsigned long long* __restrict__ trace) {
    const unsigned epoch = (unsigned)*(volatile int*)epoch_ptr;
    __shared__ double w[B200_CH_B], zs[2][B200_CH_B], part[4][B200_CH_B];
    __shared__ int s_next;
    const int tid = threadIdx.x;
    const int r = tid & (B200_CH_B - 1), q = tid >> 6; // panel updates: row r, columns kk = q + 4 m held in registers
    const int gk = tid >> 2, gpart = tid & 3;           // triangular solve: four threads per row
    int* ticket = epoch_ptr + 1;
    if (tid == 0) s_next = atomicAdd(ticket, 1);
    __syncthreads();
    for (int itx = s_next; itx < nitems; itx = s_next) {
        int nxt = 0;
        if (tid == 0) nxt = atomicAdd(ticket, 1); // the next item's ticket is fetched underneath this item's work
        const ChainItem it = items[itx];
        const int j = it.block, nr = it.nrows;
        const bool piv = j < it.K, rowok = r < nr;
        const int niter = piv ? j : it.K;
        const ChainPanel* pn = panels + it.panel_ptr;
        if (trace && tid == 0) trace[4 * (long long)itx] = gtime();
        // my rows of panel k, columns q, q+4, ...: sixteen registers, requested one panel ahead of their use
        auto load_slice = [&](double (&a)[16], const ChainPanel& pk) {
            const double* src = fac + pk.Loff + (it.row0 - pk.off) + r;
#pragma unroll
            for (int m = 0; m < 16; m++) a[m] = ldg_if(src + (long long)(q + 4 * m) * pk.f, rowok && q + 4 * m < pk.p);
        };
        double a0[16], a1[16];
        ChainPanel pk;
        unsigned long long za = 0, zb = 0;
        if (niter > 0) {
            pk = pn[0];
            load_slice(a0, pk);
            if (tid < pk.p) ll_peek(zll + pk.c0 + tid, za, zb);
        }
        // my own panel: rows of inv(L11) in registers, local pivot permutation, right-hand side
        int lp = 0;
        double dreg[16];
#pragma unroll
        for (int m = 0; m < 16; m++) dreg[m] = 0.0;
        if (piv) {
            const double* D = dinv + it.Doff;
#pragma unroll
            for (int m4 = 0; m4 < 16; m4++) {
                const int m = gpart + 4 * m4;
                dreg[m4] = ldg_if(D + gk + (long long)m * nr, gk < nr && m < gk);
            }
            if (tid < nr) lp = lperm[it.c0j + tid];
        }
        if (tid < B200_CH_B) w[tid] = (piv && tid < nr) ? y[it.c0j + tid] : 0.0;
        __syncthreads();
        // ---- contributions of the children of the chain's first front (complete long before, as a rule)
        int ok = 1;
        for (int e = 0; e < it.nch; e++) {
            const int* rg = ranges + it.rng + B200_CH_REC * e;
            const int a = rg[0], b = rg[1];
            const long long wofs = ((long long)rg[3] << 32) | (long long)(unsigned)rg[2];
            const long long llo = ((long long)rg[5] << 32) | (long long)(unsigned)rg[4];
            const int i = a + tid; // (b - a <= 64: the rows of one child are distinct)
            double val = 0.0;
            int idx = -1;
            if (i < b) idx = rel_all[wofs + i] - it.row0;
            if (llo >= 0) {
                bool have = true;
                if (i < b) {
                    unsigned long long ca, cb;
                    ll_peek(wll + llo + i, ca, cb);
                    have = ll_decode(ca, cb, epoch, val);
                }
                if (!__syncthreads_and(have)) { // not all there yet: one thread waits politely, then the missing lines are fetched
                    if (tid == 0) {
                        double dummy;
                        ok &= ll_wait(wll + llo + a, epoch, dummy, abort_flag, true) ? 1 : 0;
                    }
                    __syncthreads();
                    if (i < b && !have) ok &= ll_wait(wll + llo + i, epoch, val, abort_flag, false) ? 1 : 0;
                }
            } else if (i < b) val = __ldcg(wv + wofs + i);
            if (idx >= 0) w[idx] += val;
            __syncthreads(); // one child after the other: fixed order of additions
        }
        // ---- panels before this block, as their z arrive
        double acc = 0.0;
        auto step = [&](const int k, const double (&acur)[16], double (&anext)[16]) {
            const ChainPanel pc = pk; // panel k
            if (k + 1 < niter) {
                pk = pn[k + 1];
                load_slice(anext, pk);
            }
            double zval = 0.0;
            bool have = true;
            if (tid < pc.p) {
                have = ll_decode(za, zb, epoch, zval);
                if (have) zs[k & 1][tid] = zval;
            }
            if (!__syncthreads_and(have)) { // z_k is not complete yet
                if (tid == 0) {
                    double dummy;
                    ok &= ll_wait(zll + pc.c0, epoch, dummy, abort_flag, true) ? 1 : 0;
                }
                __syncthreads();
                if (tid < pc.p && !have) {
                    ok &= ll_wait(zll + pc.c0 + tid, epoch, zval, abort_flag, false) ? 1 : 0;
                    zs[k & 1][tid] = zval;
                }
                __syncthreads();
            }
            if (k + 1 < niter && tid < pk.p) ll_peek(zll + pk.c0 + tid, za, zb); // next panel's z, one iteration ahead
            if (trace && tid == 0 && k == niter - 1) trace[4 * (long long)itx + 1] = gtime();
            const double* zk = zs[k & 1];
#pragma unroll
            for (int m = 0; m < 16; m++)
                if (q + 4 * m < pc.p) acc += acur[m] * zk[q + 4 * m];
        };
        for (int k = 0; k < niter; k += 2) {
            step(k, a0, a1);
            if (k + 1 < niter) step(k + 1, a1, a0);
        }
        part[q][r] = acc;
        if (!__syncthreads_and(ok)) return;
        if (tid < nr) w[tid] -= (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
        __syncthreads();
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
                ll_store(zll + it.c0j + gk, epoch, s); // the blocks after this one may go on
                zv[it.c0j + gk] = s;
            }
        } else if (tid < nr) {
            ll_store(wll + it.out + tid, epoch, w[tid]); // final rows of the chain's update vector
        }
        if (trace && tid == 0) trace[4 * (long long)itx + 2] = gtime();
        if (tid == 0) s_next = nxt;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256, 2) k_bwd_chain(const ChainItem* __restrict__ items, int nitems, const ChainPanel* __restrict__ panels,
                                                   const int* __restrict__ rows_all, const double* __restrict__ fac,
                                                   const double* __restrict__ dinv, const double* __restrict__ zv, double* __restrict__ xp,
                                                   ulonglong2* __restrict__ xll, ulonglong2* __restrict__ pll, int* __restrict__ epoch_ptr,
                                                   int* __restrict__ abort_flag, unsigned long long* __restrict__ trace) {
    const unsigned epoch = (unsigned)*(volatile int*)epoch_ptr;
    __shared__ double xs[B200_CH_B], t[B200_CH_B], part[4][B200_CH_B], pr[2][2][B200_CH_B];
    __shared__ int s_next;
    const int tid = threadIdx.x, lane = tid & 31;
    const int r = tid & (B200_CH_B - 1), q = tid >> 6; // dot products: row r (coalesced loads), pivots i = q + 4 m in registers
    const int half = (tid >> 5) & 1;                    // rows 0..31 / 32..63 of the block
    const int gk = tid >> 2, gpart = tid & 3;
    const int mycol = ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0); // after warp_reduce16
    int* ticket = epoch_ptr + 2;
    if (tid == 0) s_next = nitems - 1 - atomicAdd(ticket, 1);
    __syncthreads();
    for (int itx = s_next; itx >= 0; itx = s_next) {
        int nxt = 0;
        if (tid == 0) nxt = nitems - 1 - atomicAdd(ticket, 1);
        const ChainItem it = items[itx];
        const int j = it.block, nr = it.nrows;
        const bool piv = j < it.K, rowok = r < nr;
        const int niter = piv ? j : it.K;
        const ChainPanel* pn = panels + it.panel_ptr;
        if (trace && tid == 0) trace[4 * (long long)itx] = gtime();
        // my rows of panel k's U panel (u_k x p_k, row jj = chain row off_k + p_k + jj), pivots q, q+4, ...
        auto load_slice = [&](double (&a)[16], const ChainPanel& pk) {
            const double* src = fac + pk.Uoff + (it.row0 - pk.off - pk.p) + r;
#pragma unroll
            for (int m = 0; m < 16; m++) a[m] = ldg_if(src + (long long)(q + 4 * m) * pk.u, rowok && q + 4 * m < pk.p);
        };
        double a0[16], a1[16];
        ChainPanel pk;
        if (niter > 0) {
            pk = pn[niter - 1];
            load_slice(a0, pk);
        }
        int ok = 1;
        if (!piv) {
            // update block: its x are solution entries of ancestor columns
            int col = 0;
            double v = 0.0;
            bool have = true;
            if (tid < nr) {
                col = rows_all[it.rows_off + tid];
                unsigned long long ca, cb;
                ll_peek(xll + col, ca, cb);
                have = ll_decode(ca, cb, epoch, v);
            }
            if (!__syncthreads_and(have)) {
                if (tid == 0) {
                    double dummy;
                    ok &= ll_wait(xll + col, epoch, dummy, abort_flag, true) ? 1 : 0;
                }
                __syncthreads();
                if (tid < nr && !have) ok &= ll_wait(xll + col, epoch, v, abort_flag, false) ? 1 : 0;
            }
            if (tid < nr) xs[tid] = v;
            if (trace && tid == 0) trace[4 * (long long)itx + 1] = gtime();
        } else {
            // pivot block: x_j = inv(U11) (z_j - sum of the partial dot products of the blocks after it)
            double dreg[16];
            {
                const double* D = dinv + it.Doff;
#pragma unroll
                for (int m4 = 0; m4 < 16; m4++) {
                    const int m = gk + gpart + 4 * m4;
                    dreg[m4] = ldg_if(D + gk + (long long)m * nr, gk < nr && m < nr);
                }
            }
            const double zj = tid < nr ? zv[it.c0j + tid] : 0.0;
            const ulonglong2* grp = pll + (it.pgrp + (long long)j * it.nblocks) * B200_CH_B;
            // thread (r, q) sums the lines of blocks j+1+q, j+5+q, ... of pivot r: eight lines are requested at once, whatever
            // is missing (as a rule only the line of block j+1, published last) is waited for afterwards
            double acc = 0.0;
            for (int jbase = j + 1; jbase < it.nblocks; jbase += 32) { // (uniform trip count: there are barriers inside)
                const int jb0 = jbase + q;
                unsigned long long ca[8], cb[8];
#pragma unroll
                for (int e = 0; e < 8; e++)
                    if (rowok && jb0 + 4 * e < it.nblocks) ll_peek(grp + (long long)(jb0 + 4 * e) * B200_CH_B + r, ca[e], cb[e]);
                bool miss = false;
                double v[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    v[e] = 0.0;
                    if (rowok && jb0 + 4 * e < it.nblocks && !ll_decode(ca[e], cb[e], epoch, v[e])) miss = true;
                }
                if (__syncthreads_or(miss)) {
                    if (jbase == j + 1 && tid == 0) { // the block right after this one publishes last
                        double dummy;
                        ok &= ll_wait(grp + (long long)(j + 1) * B200_CH_B, epoch, dummy, abort_flag, true) ? 1 : 0;
                    }
                    __syncthreads();
#pragma unroll
                    for (int e = 0; e < 8; e++)
                        if (rowok && jb0 + 4 * e < it.nblocks && !ll_decode(ca[e], cb[e], epoch, v[e]))
                            ok &= ll_wait(grp + (long long)(jb0 + 4 * e) * B200_CH_B + r, epoch, v[e], abort_flag, false) ? 1 : 0;
                }
#pragma unroll
                for (int e = 0; e < 8; e++) acc += v[e]; // fixed order
            }
            if (trace && tid == 0) trace[4 * (long long)itx + 1] = gtime();
            part[q][r] = acc;
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
                ll_store(xll + it.c0j + gk, epoch, s); // for the update blocks of the chains below
                xp[it.c0j + gk] = s;                   // for the subtree kernels and the caller
            }
        }
        if (!__syncthreads_and(ok)) return;
        const double xr = rowok ? xs[r] : 0.0;
        // ---- partial dot products of my rows for the panels before this block (last panel first)
        auto step = [&](const int k, const int par, const double (&acur)[16], double (&anext)[16]) {
            const ChainPanel pc = pk; // panel k
            if (k > 0) {
                pk = pn[k - 1];
                load_slice(anext, pk);
            }
            double v[16];
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = acur[m] * xr;
            const double tot = warp_reduce16(v, lane); // over the 32 rows of this warp, pivot q + 4 * mycol
            if ((lane & 1) == 0) pr[par][half][q + 4 * mycol] = tot;
            __syncthreads();
            if (tid < pc.p) ll_store(pll + (it.pgrp + (long long)k * it.nblocks + j) * B200_CH_B + tid, epoch, pr[par][0][tid] + pr[par][1][tid]);
        };
        for (int k = niter - 1; k >= 0; k -= 2) {
            step(k, 0, a0, a1);
            if (k - 1 >= 0) step(k - 1, 1, a1, a0);
        }
        if (trace && tid == 0) trace[4 * (long long)itx + 2] = gtime();
        if (tid == 0) s_next = nxt;
        __syncthreads();
    }
}

} // namespace b200
