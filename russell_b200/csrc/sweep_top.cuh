// sweep_top.cuh -- persistent, dependency-driven triangular-solve kernels for the TOP of the front tree.
//
// The upper levels of the assembly tree hold few, large fronts (the chain-split separators): per-level kernel
// launches leave the GPU idle between ~60 dependent launches.  Here ONE co-resident grid walks all (front, row
// slice) items of those levels in level order; instead of kernel boundaries, a slice waits on per-front completion
// counters in global memory (release: __threadfence + atomicAdd, acquire: volatile load + __threadfence), and the
// front's panel slice and pivot-block inverse are staged into shared memory with cp.async BEFORE the wait, so the
// HBM stream runs ahead of the dependency wave.
//
// Progress guarantee: items are sorted so that every dependency has a smaller index, and CTAs take items from an ATOMIC
// TICKET counter in increasing order (backward sweep: decreasing).  The unfinished item with the smallest ticket was
// therefore handed to a CTA that is running (it executed the atomic), and all of its dependencies are finished or owned by
// running CTAs as well: the wave always advances, whatever part of the grid is resident -- two handles sweeping at the same
// time, or a sweep next to another handle's factorization (Radau5's real + complex systems), cannot starve each other.
// Spins are bounded all the same: on timeout an abort flag makes every CTA leave, and the host reports B200_ERROR_SOLVE+1.
//
// Role in the reference: the inside of umfpack_di_solve / cudssExecute(SOLVE)
// (russell_sparse/c_code/interface_umfpack.c:229, interface_cudss.cu:530).
#pragma once
#include "kernels.cuh"
#include "sweep_sub.cuh" // warp_reduce8

namespace b200 {

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
}

// bounded acquire-spin: returns false (and raises *abort_flag) on timeout or when another CTA aborted
__device__ __forceinline__ bool wait_counter_ge(const int* counter, int target, int* abort_flag) {
    const volatile int* vc = counter;
    const volatile int* va = abort_flag;
    for (int it = 0; it < (1 << 24); it++) { // ~0.5 us per probe: a few seconds before giving up
        if (*vc >= target) {
            __threadfence();
            return true;
        }
        if ((it & 255) == 255 && *va) return false;
    }
    atomicExch(abort_flag, 1);
    return false;
}

// dynamic shared memory: one panel slice (SLICE x MAXP doubles); the pivot-block inverses live in registers
#define B200_TOP3_SMEM ((size_t)(B200_SLICE * B200_MAXP) * sizeof(double)) // k_bwd_top3 stages the U slice only

// per (item, child) record of the forward sweep, built on the host: head count, slice begin, slice end (positions in the
// child's update list), child front, offset of the child's update list in rel[] / wv[] (two halves of an int64), number
// of row slices of the child (the completion-counter target per epoch), pad
#define B200_TOP_REC 8
#define B200_TOP_EC 4 // children whose gather is kept in registers across the dependency wait (the rest take the slow loop)

// optional per-item timestamps (debug / profiling): trace[4*item + {0,1,2}] = item start, dependency satisfied, item end
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ int slices_of(int u) { return u > 0 ? (u + B200_SLICE - 1) / B200_SLICE : 1; }

// ---- forward sweep: nothing but the dependent data itself is loaded after the dependency wait -------------------------
// The first version of these kernels paid 4-5 serialized L2 round trips per chain link AFTER the wait (child descriptor -> relative
// indices -> values -> local permutation; children polled one after the other).  Here every descriptor, index, the local
// permutation and z come from host-built records or are loaded before the wait; the wait itself is one poll per child in
// parallel; after it there is ONE batch of value loads (all children in flight together).  Fronts with a single row slice
// skip the cross-CTA ticket reduction of the backward sweep.  Items are handed out through an atomic ticket counter (see the progress guarantee above).
__global__ void __launch_bounds__(256, 3) k_fwd_top2(const SolveItem* __restrict__ items, int nitems, const NodeDev* __restrict__ nodes,
                                                  const int* __restrict__ rel_all, const double* __restrict__ fac,
                                                  const double* __restrict__ dinv, const int* __restrict__ lperm,
                                                  const int* __restrict__ ranges, const double* __restrict__ y, double* __restrict__ zv,
                                                  double* __restrict__ wv, int* __restrict__ cdone, const int* __restrict__ epoch_ptr,
                                                  int* __restrict__ abort_flag, unsigned long long* __restrict__ trace) {
    const int epoch = *epoch_ptr;
    extern __shared__ double smt[];
    double* Ps = smt; // this slice of L21: Ps[k * SLICE + r]  (inv(L11) is kept in registers: B200_TOP3_SMEM, 3 CTAs per SM)
    __shared__ double t1[B200_MAXP], z[B200_MAXP], wloc[B200_SLICE], wpart[B200_SLICE];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int gk = tid >> 2, gpart = tid & 3; // GEMV layout: four threads per row
    __shared__ int s_next;
    int* ticket = const_cast<int*>(epoch_ptr) + 1;
    if (tid == 0) s_next = atomicAdd(ticket, 1);
    __syncthreads();
    for (int itx = s_next; itx < nitems; itx = s_next) {
        int nxt = 0;
        if (tid == 0) nxt = atomicAdd(ticket, 1); // the next item's ticket is fetched underneath this item's work
        const SolveItem it = items[itx];
        const NodeDev nd = nodes[it.node];
        const int p = nd.p, u = nd.u, nchild = nd.nchild;
        const long long f = (long long)p + u;
        {
            const int r = tid & (B200_SLICE - 1), g = tid >> 7;
            if (r < it.nrows) {
                const double* src = fac + nd.Loff + p + it.r0 + r;
                for (int k = g; k < p; k += 2) cp_async8(Ps + k * B200_SLICE + r, src + (long long)k * f);
            }
        }
        double dreg[16]; // row gk of inv(L11), columns gpart, gpart+4, ... (< gk): loaded before the wait
        {
            const double* D = dinv + nd.Doff;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int m = gpart + 4 * q;
                dreg[q] = (gk < p && m < gk) ? D[gk + (long long)m * p] : 0.0;
            }
        }
        if (trace && tid == 0) trace[4 * (long long)itx] = gtime();
        const int lp = tid < p ? lperm[nd.c0 + tid] : 0;
        if (tid < p) t1[tid] = y[nd.c0 + tid];
        if (tid < B200_SLICE) wloc[tid] = 0.0;
        const int lo = p + it.r0;
        const int* rg0 = ranges + it.rng;
        int ih[B200_TOP_EC], is[B200_TOP_EC];
        const double *ph[B200_TOP_EC], *ps[B200_TOP_EC];
#pragma unroll
        for (int e = 0; e < B200_TOP_EC; e++) {
            ih[e] = -1, is[e] = -1, ph[e] = nullptr, ps[e] = nullptr;
            if (e < nchild) {
                const int* rg = rg0 + B200_TOP_REC * e;
                const int nhead = rg[0], a = rg[1], b = rg[2];
                const long long wofs = ((long long)rg[5] << 32) | (long long)(unsigned)rg[4];
                if (tid < nhead) ih[e] = rel_all[wofs + tid], ph[e] = wv + wofs + tid;
                if (a + tid < b) is[e] = rel_all[wofs + a + tid] - lo, ps[e] = wv + wofs + a + tid;
            }
        }
        int ok = 1;
        for (int e = tid; e < nchild; e += nt) {
            const int* rg = rg0 + B200_TOP_REC * e;
            ok &= wait_counter_ge(&cdone[rg[3]], epoch * rg[6], abort_flag) ? 1 : 0;
        }
        if (!__syncthreads_and(ok)) return;
        if (trace && tid == 0) trace[4 * (long long)itx + 1] = gtime();
        double vh[B200_TOP_EC], vs[B200_TOP_EC];
#pragma unroll
        for (int e = 0; e < B200_TOP_EC; e++) {
            vh[e] = ih[e] >= 0 ? __ldcg(ph[e]) : 0.0; // written by other SMs: read through L2
            vs[e] = is[e] >= 0 ? __ldcg(ps[e]) : 0.0;
        }
#pragma unroll
        for (int e = 0; e < B200_TOP_EC; e++)
            if (e < nchild) { // children are applied one after the other (fixed order: deterministic sums)
                if (ih[e] >= 0) t1[ih[e]] += vh[e];
                if (is[e] >= 0) wloc[is[e]] += vs[e];
                __syncthreads();
            }
        for (int e = B200_TOP_EC; e < nchild; e++) { // fronts with many children (rare at the top of the tree)
            const int* rg = rg0 + B200_TOP_REC * e;
            const int nhead = rg[0], a = rg[1], b = rg[2];
            const long long wofs = ((long long)rg[5] << 32) | (long long)(unsigned)rg[4];
            const int* rel = rel_all + wofs;
            const double* wc = wv + wofs;
            for (int i = tid; i < nhead; i += nt) t1[rel[i]] += __ldcg(wc + i);
            for (int i = a + tid; i < b; i += nt) wloc[rel[i] - lo] += __ldcg(wc + i);
            __syncthreads();
        }
        double tp = 0.0;
        if (tid < p) tp = t1[lp];
        cp_async_wait_all();
        __syncthreads();
        if (tid < p) t1[tid] = tp;
        __syncthreads();
        {   // z = inv(L11) t1: four threads per row (fixed partition + fixed shuffle order: deterministic)
            const int k = gk, part = gpart;
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int m = gpart + 4 * q;
                if (gk < p && m < gk) s += dreg[q] * t1[m];
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (k < p && part == 0) {
                s += t1[k];
                z[k] = s;
                if (it.slice == 0) zv[nd.c0 + k] = s;
            }
        }
        __syncthreads();
        {
            const int r = tid & (B200_SLICE - 1), h = tid >> 7;
            const int kh = (p + 1) >> 1;
            const int kbeg = h * kh, kend = min(p, kbeg + kh);
            double s = 0.0;
            if (r < it.nrows) { // four independent chains (the dependent-add chain of 32 FMAs was the longest leg of an item)
                double s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int k = kbeg;
                for (; k + 3 < kend; k += 4) {
                    s += Ps[k * B200_SLICE + r] * z[k], s1 += Ps[(k + 1) * B200_SLICE + r] * z[k + 1];
                    s2 += Ps[(k + 2) * B200_SLICE + r] * z[k + 2], s3 += Ps[(k + 3) * B200_SLICE + r] * z[k + 3];
                }
                for (; k < kend; k++) s += Ps[k * B200_SLICE + r] * z[k];
                s = (s + s1) + (s2 + s3);
            }
            if (h == 1) wpart[r] = s;
            __syncthreads();
            if (h == 0 && r < it.nrows) wv[nd.rows_ptr + it.r0 + r] = wloc[r] - (s + wpart[r]);
        }
        if (tid == 0) s_next = nxt; // (every thread read the previous value before this item's first barrier)
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(&cdone[it.node], 1);
            if (trace) trace[4 * (long long)itx + 2] = gtime();
        }
    }
}

// ---- v3 backward sweep (default): the CONSUMER finishes its parent's pivot block ---------------------------------------
// In k_bwd_top/k_bwd_top2 a chain link costs two cross-CTA hops: slices -> (ticket) -> last slice reduces the partial
// dot products, applies inv(U11), publishes x1 -> children.  The trace (profiles/r01m_trace_*) shows 7.5 us per level for
// it against 3.5 us for the single-hop forward sweep.  Here a front's slices only publish their partial dot products
// (release: fence + atomicAdd on bdone[front]); every slice CTA of every CHILD waits for that counter, loads the nsl x p
// partials and redundantly evaluates x1 = inv(U11) (z - sum of partials) for its PARENT with inv(U11) held in registers
// (loaded before the wait) -- the same arithmetic in the same order in every CTA, so the values are identical.  One hop
// per level.  x1 is written to xp by slice 0 of each child before that child releases its own counter (so that deeper
// descendants, which gather it from xp, see it) and by the front's own last slice (which covers fronts whose children
// live outside the persistent region).
__global__ void __launch_bounds__(256, 3) k_bwd_top3(const SolveItem* __restrict__ items, int nitems, const NodeDev* __restrict__ nodes,
                                                  const int* __restrict__ rows_all, const double* __restrict__ fac,
                                                  const double* __restrict__ dinv, const double* __restrict__ zv,
                                                  double* __restrict__ xp, double* __restrict__ scratch,
                                                  const int* __restrict__ node_slot, int* __restrict__ bdone,
                                                  const int* __restrict__ epoch_ptr, int* __restrict__ abort_flag,
                                                  unsigned long long* __restrict__ trace) {
    const int epoch = *epoch_ptr;
    extern __shared__ double smt[];
    double* Ps = smt; // this slice of the U panel: Ps[k * SLICE + r]
    __shared__ double t[B200_MAXP], xpar[B200_MAXP], x2[B200_SLICE];
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int gk = tid >> 2, gpart = tid & 3; // GEMV layout: four threads per row
    __shared__ int s_next;
    int* ticket = const_cast<int*>(epoch_ptr) + 2;
    if (tid == 0) s_next = nitems - 1 - atomicAdd(ticket, 1);
    __syncthreads();
    for (int itx = s_next; itx >= 0; itx = s_next) {
        int nxt = 0;
        if (tid == 0) nxt = nitems - 1 - atomicAdd(ticket, 1);
        const SolveItem it = items[itx];
        const NodeDev nd = nodes[it.node];
        const int p = nd.p, u = nd.u;
        const int nsl = slices_of(u);
        {
            const int r = tid & (B200_SLICE - 1), g = tid >> 7;
            if (r < it.nrows) {
                const double* src = fac + nd.Uoff + it.r0 + r;
                for (int k = g; k < p; k += 2) cp_async8(Ps + k * B200_SLICE + r, src + (long long)k * u);
            }
        }
        // ---- everything static is loaded before the wait: own row indices and slot, the parent's descriptor, z and inv(U11)
        const int par = nd.pad;
        const int rr = tid < it.nrows ? rows_all[nd.rows_ptr + it.r0 + tid] : -1;
        const int slot = node_slot[it.node];
        int pp = 0, nslp = 0, slotp = 0, c0p = 0;
        double dreg[16], zp = 0.0;
#pragma unroll
        for (int q = 0; q < 16; q++) dreg[q] = 0.0;
        if (par >= 0) {
            const NodeDev pd = nodes[par];
            pp = pd.p, nslp = slices_of(pd.u), slotp = node_slot[par], c0p = pd.c0;
            const double* Dp = dinv + pd.Doff;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int m = gk + gpart + 4 * q;
                if (gk < pp && m < pp) dreg[q] = Dp[gk + (long long)m * pp];
            }
            if (tid < pp) zp = zv[c0p + tid];
        }
        if (tid == 0) {
            if (trace) trace[4 * (long long)itx] = gtime();
            s_flag = (par < 0) ? 1 : (wait_counter_ge(&bdone[par], epoch * nslp, abort_flag) ? 1 : 0);
            if (trace) trace[4 * (long long)itx + 1] = gtime();
        }
        __syncthreads();
        if (!s_flag) return;
        // ---- one batch of dependent loads: the parent's partials and the solution entries of older ancestors
        const bool in_par = rr >= c0p && rr < c0p + pp; // (pp == 0 without a parent)
        double xv = 0.0;
        if (rr >= 0 && !in_par) xv = __ldcg(xp + rr);
        if (par >= 0) {
            if (tid < pp) {
                double sacc = zp;
                const double* base = scratch + (long long)slotp * B200_MAXP + tid;
                for (int sl0 = 0; sl0 < nslp; sl0 += 8) { // eight loads in flight, subtracted in slice order
                    double v[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) v[q] = sl0 + q < nslp ? __ldcg(base + (long long)(sl0 + q) * B200_MAXP) : 0.0;
#pragma unroll
                    for (int q = 0; q < 8; q++) sacc -= v[q];
                }
                t[tid] = sacc;
            }
            __syncthreads();
            {   // x1(parent) = inv(U11) t, four threads per row, inv(U11) from registers (same order as the shared-memory GEMV)
                double sacc = 0.0;
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int m = gk + gpart + 4 * q;
                    if (gk < pp && m < pp) sacc += dreg[q] * t[m];
                }
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
                if (gk < pp && gpart == 0) {
                    xpar[gk] = sacc;
                    if (it.slice == 0) xp[c0p + gk] = sacc; // visible to deeper descendants through this front's release
                }
            }
            __syncthreads();
        }
        if (rr >= 0) x2[tid] = in_par ? xpar[rr - c0p] : xv;
        cp_async_wait_all();
        __syncthreads();
        double* part = scratch + ((long long)slot + it.slice) * B200_MAXP;
        {   // warp w owns the columns 8w .. 8w+7: lane-local products over its four rows, then ONE 9-shuffle butterfly for the
            // eight columns (40 shuffles with a reduction per column)
            double v8[8];
            const bool ra = lane < it.nrows, rb = lane + 32 < it.nrows, rc = lane + 64 < it.nrows, rd = lane + 96 < it.nrows;
            const double xa = ra ? x2[lane] : 0.0, xb = rb ? x2[lane + 32] : 0.0, xc = rc ? x2[lane + 64] : 0.0, xd = rd ? x2[lane + 96] : 0.0;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const double* col = Ps + (8 * warp + c) * B200_SLICE;
                double acc8 = 0.0;
                if (8 * warp + c < p) { // (rows beyond the slice were never staged: they must not be read)
                    if (ra) acc8 = col[lane] * xa;
                    if (rb) acc8 += col[lane + 32] * xb;
                    if (rc) acc8 += col[lane + 64] * xc;
                    if (rd) acc8 += col[lane + 96] * xd;
                }
                v8[c] = acc8;
            }
            const double tot = warp_reduce8(v8, lane);
            const int c = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); // the column this lane holds
            if ((lane & 3) == 0 && 8 * warp + c < p) part[8 * warp + c] = tot;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int old = atomicAdd(&bdone[it.node], 1); // release: the children of this front may start
            s_flag = (old + 1 == epoch * nsl);
            if (trace) trace[4 * (long long)itx + 2] = gtime();
        }
        __syncthreads();
        if (s_flag) { // last slice of this front: write its own x1 (off the critical path of the children)
            __threadfence();
            if (tid < p) {
                double sacc = zv[nd.c0 + tid];
                const double* base = scratch + (long long)slot * B200_MAXP + tid;
                for (int sl0 = 0; sl0 < nsl; sl0 += 8) {
                    double v[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) v[q] = sl0 + q < nsl ? __ldcg(base + (long long)(sl0 + q) * B200_MAXP) : 0.0;
#pragma unroll
                    for (int q = 0; q < 8; q++) sacc -= v[q];
                }
                t[tid] = sacc;
            }
            const double* Dv = dinv + nd.Doff;
            double dv[16];
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int m = gk + gpart + 4 * q;
                dv[q] = (gk < p && m < p) ? Dv[gk + (long long)m * p] : 0.0;
            }
            __syncthreads();
            double sacc = 0.0;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int m = gk + gpart + 4 * q;
                if (gk < p && m < p) sacc += dv[q] * t[m];
            }
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            if (gk < p && gpart == 0) xp[nd.c0 + gk] = sacc;
        }
        if (tid == 0) s_next = nxt;
        __syncthreads(); // shared buffers are reused by the next item
    }
}

} // namespace b200
