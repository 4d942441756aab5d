// solver_b200.cu -- host driver + C ABI (include/solver_b200.h) of the B200-native sparse direct solver.
//
// Mirrors the reference's cuDSS shim function by function (russell_sparse/c_code/interface_cudss.cu):
//   solver_b200_new        <- solver_cudss_new        (:62-123)   stream + handle
//   solver_b200_drop       <- solver_cudss_drop       (:126-171)  frees everything, NULL-safe
//   solver_b200_initialize <- solver_cudss_initialize (:190-396)  H2D of the structure + ANALYSIS phase
//   solver_b200_factorize  <- solver_cudss_factorize  (:406-501)  H2D of values + FACTORIZATION phase
//   solver_b200_solve      <- solver_cudss_solve      (:510-566)  H2D rhs, SOLVE phase (+ refinement), D2H x
// but the phases are our own kernels (kernels.cuh) instead of cudssExecute().
//
// There is deliberately NO CPU fallback: without a CUDA device solver_b200_new() returns NULL and every
// entry point reports B200_ERROR_NOT_AVAILABLE.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/solver_b200.h"
#include "kernels.cuh"
#include "sweep_top.cuh"
#include "sweep_sub.cuh"
#include "ozaki_tc.cuh"
#include "plan.hpp"
#include "coo_guard.hpp"
#include <atomic>
#include <chrono>
#include <thread>
#include <memory>
#include <mutex>
#include <condition_variable>
#include <functional>

using namespace b200;

#define CUDA_TRY(expr, code)                                                                        \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            if (s && s->verbose) fprintf(stderr, "solver_b200: %s failed: %s\n", #expr, cudaGetErrorString(_e)); \
            return (code);                                                                          \
        }                                                                                           \
    } while (0)

// front-size classes of the fused kernel: upper bound on f and CTA size; class NFC = "big" (multi-kernel path)
#define B200_ASM_SMEM_MAX (160 * 1024) // a one-column tile of a front of order 20,000 still fits
static const int NFC = 6;
static const int FC_MAXF[NFC] = {16, 32, 48, 64, 96, B200_FUSED_MAXF};
static const int FC_THREADS[NFC] = {32, 64, 64, 128, 256, 256};
// pivot-count classes of the batched inverse post-pass
static const int NIC = 3;
static const int IC_MAXP[NIC] = {16, 32, B200_MAXP};
// solve classes: CTA size by front order
static const int NSC = 3;
static const int SC_MAXF[NSC] = {32, 256, 1 << 30};
static const int SC_THREADS[NSC] = {32, 128, 256};
static const int SC_MAXP[NSC] = {32, B200_MAXP, B200_MAXP};

// A few persistent helper threads per handle for the striped host <-> device transfers (spawning threads per transfer cost
// ~0.15 ms, three times per step).  run(n, f) executes f(0) .. f(n-1): f(n-1) on the calling thread, the rest on the workers.
struct StripePool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<void(int)> job;
    int generation = 0, njobs = 0, pending = 0;
    bool stop = false;
    void ensure(int n) {
        while ((int)workers.size() < n) {
            const int t = (int)workers.size();
            workers.emplace_back([this, t]() {
                int seen = 0;
                for (;;) {
                    std::function<void(int)> f;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv_go.wait(lk, [&]() { return stop || generation != seen; });
                        if (stop) return;
                        seen = generation;
                        if (t >= njobs) continue;
                        f = job;
                    }
                    f(t);
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        if (--pending == 0) cv_done.notify_all();
                    }
                }
            });
        }
    }
    void run(int n, const std::function<void(int)>& f) {
        if (n > 1) {
            ensure(n - 1);
            {
                std::lock_guard<std::mutex> lk(mu);
                job = f, njobs = n - 1, pending = n - 1, generation++;
            }
            cv_go.notify_all();
        }
        f(n - 1);
        if (n > 1) {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&]() { return pending == 0; });
        }
    }
    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_go.notify_all();
        for (auto& w : workers) w.join();
        workers.clear();
        stop = false;
    }
};

// Device memory of a handle comes from a few large chunks instead of ~70 separate cudaMalloc calls: on the B200 boxes a
// cudaMalloc costs 0.1 ms on a quiet device but ~5-10 ms next to live handles / instantiated graphs (profiles/r4f_init_probe.txt:
// 0.28-0.41 s of a 0.9 s initialize went into cudaMalloc alone).  A request that does not fit the chunk being filled gets a chunk
// of its own when it is at least half a chunk, a fresh chunk otherwise.
struct SlabSet {
    struct Chunk {
        char* base;
        size_t cap, used;
    };
    std::vector<Chunk> chunks; // the chunk that is being filled is the last one
    size_t chunk_default = (size_t)1 << 20;
    int n_malloc = 0;
    cudaError_t alloc(void** out, size_t bytes) {
        bytes = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
        const bool fits = !chunks.empty() && chunks.back().used + bytes <= chunks.back().cap;
        const bool own = !fits && 2 * bytes >= chunk_default;
        if (!fits) {
            const size_t cap = own ? bytes : chunk_default;
            char* b = nullptr;
            cudaError_t e = cudaMalloc((void**)&b, cap);
            n_malloc++;
            if (e != cudaSuccess) return e;
            if (own) { // full from the start: goes in front of the chunk that is being filled
                chunks.insert(chunks.empty() ? chunks.end() : chunks.end() - 1, Chunk{b, cap, cap});
                *out = b;
                return cudaSuccess;
            }
            chunks.push_back(Chunk{b, cap, 0});
        }
        Chunk& c = chunks.back();
        *out = c.base + c.used;
        c.used += bytes;
        return cudaSuccess;
    }
    bool owns(const void* p) const {
        for (const Chunk& c : chunks)
            if ((const char*)p >= c.base && (const char*)p < c.base + c.cap) return true;
        return false;
    }
    void release() {
        for (Chunk& c : chunks) cudaFree(c.base);
        chunks.clear();
    }
};

struct LevelLists {
    // offsets (nlevels+1) into the concatenated device item arrays (big fronts only)
    std::vector<int> asm_ptr, panel_ptr, schur_ptr;
    // fact_ptr[l*(NFC+1)+c .. +1]: nodes of level l and factorization class c inside d_fact_nodes (class NFC = big)
    std::vector<int> fact_ptr;
    // solve_ptr[l*NSC+c .. +1] inside d_solve_nodes
    std::vector<int> solve_ptr;
    std::vector<int> solve_threads, solve_pmax; // per level: block size and largest pivot count of the small launch
    std::vector<int> big_ptr; // nlevels+1: slices of the big solve class inside d_big_items
    std::vector<int> inv_ptr; // NIC+1: fronts by pivot-count class inside d_inv_nodes
    std::vector<int> inv_mid; // NIC: first front of the class that waits for the end of the level loop (the ones before it lie below inv_split_level)
    std::vector<size_t> fused_smem; // per (level, class): dynamic shared memory of the fused launch
    // k_front_warp: the fronts of a (level, class) range are sorted by order and cut into size buckets, one launch each
    // (shared memory per warp follows the bucket's largest front: small fronts get more resident warps)
    struct FwBatch {
        int start, count, need; // range inside d_fact_nodes; doubles of shared memory of the largest front ((f | 1) * f)
    };
    std::vector<FwBatch> fw_batches;
    std::vector<int> fw_ptr; // per (level, class) + 1: batches inside fw_batches
    std::vector<size_t> asm_smem;   // per level: largest tile (bytes) of k_assemble_tile
    // schur_variant 2: fronts with u >= ozaki_min_u take the tcgen05 kernel (ozaki_tc.cuh)
    std::vector<int> oz_split_ptr, oz_item_ptr, oz_rows_max; // nlevels+1 / nlevels+1 / nlevels
    std::vector<int> diag_cnt; // per level: big fronts that need their own pivot-block launch (listed first in their class)
    struct SchurFront { // a front whose Schur tiles are enumerated by the launch grid (k_schur_dmma_front)
        int node, nt, parent, lookahead;
    };
    std::vector<SchurFront> schur_fronts;
    std::vector<int> schur_front_ptr; // per level
    size_t oz_tile_bytes = 0, oz_scales = 0;                 // arena sizes (largest level)
};

struct InterfaceB200 {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool initialized = false, factorized = false;
    int verbose = 0;

    // options
    int opt_panel_width = 64, opt_nd_leaf = 32; // (nd_leaf: measured at config 2 -- 96: 8.68 ms, 48: 8.46, 32: 8.27, 24: 8.29, 12: 8.20 ms per factorization)
    int use_graph = 1;
    int schur_variant = 1; // 0 = FMA, 1 = DMMA (mma.sync f64), 2 = DMMA + tcgen05 int8 Ozaki kernel for fronts with u >= ozaki_min_u
    int ozaki_min_u = 1024;
    OzakiSplitItem* d_oz_split = nullptr;
    OzakiItem* d_oz_items = nullptr;
    signed char* d_oz_tiles = nullptr;
    double* d_oz_scales = nullptr;
    int relax_small = -1;                                // supernode amalgamation knobs of the host analysis (plan.hpp); < 0: defaults
    double relax_z1 = -1.0, relax_z2 = -1.0, relax_z3 = -1.0;
    int asm_variant = 1;     // 0 = k_assemble (read-modify-write in global memory), 1 = k_assemble_tile (tile in shared memory)
    int use_leaf_reg = 1;    // leaf fronts of order <= 32: k_leaf_reg (one warp per front, registers only)
    int panel_row_max = 160; // launches of at most this many 128-row panel items use k_panel_row (one warp per four rows)
    int panel_variant = 1; // 0 = k_panel (32-row tiles, barrier per column), 1 = k_panel_warp (thread per row, 128-row items)
    int use_fused = 1;     // fronts with f <= B200_FUSED_MAXF go through k_front_fused
    int fused_maxf = 48;   // fronts above this order take the multi-kernel path (measured optimum at config 2)
    int fuse_chain = 1;    // chain links receive their child's Schur complement directly (no k_assemble pass)
    int fuse_diag = 1;     // ... and the Schur CTA of tile (0, 0) factorizes the parent's pivot block right away (look-ahead)
    int diag_variant = 4;  // 0 = shared-memory LU (k_diag), 1 = register-resident LU with implicit pivoting (k_diag_reg)
    int nrefine = 2;
    double ir_tol = 1e-11;
    double pivot_eps = 1e-13;
    int force_no_matching = 0;
    int strict_residual = 0; // 1: solve returns B200_ERROR_SOLVE+7 whenever ||b-Ax||/||b|| > 10 ir_tol after refinement

    // Host buffers of the callers (Rust Vec<f64>, numpy arrays) are pageable: a plain cudaMemcpyAsync from them is staged by the
    // driver at a fraction of the PCIe rate.  Large pageable transfers are striped over a few (pooled) threads, each copying its stripe
    // through its own pair of pinned staging buffers (memcpy of piece k+1 overlaps the DMA of piece k); pinned buffers (the
    // benchmark's) go straight to the copy engine.  Option "staged_copy" / B200_STAGED_COPY = 0 restores the plain copies.
    static const int NST = 4;                     // stripes (threads): the host's memcpy rate, not the thread count, bounds the staged path
                                                  // (8 stripes: 10.0 ms per step, 4: 10.1 ms, plain cudaMemcpyAsync: 10.9-11.6 ms; pinned: 7.7 ms)
    static const size_t STAGE_PIECE = (size_t)1 << 20; // bytes per staging buffer (8 MB of pinned memory per handle, allocated at the first staged transfer)
    char* h_stage = nullptr;                      // NST x 2 pinned buffers
    cudaEvent_t ev_stage[2 * NST] = {};
    StripePool stripe_pool;
    int staged_copy = 1;
    SlabSet slabs; // device memory of the plan arrays, work vectors and arenas (released as a whole)
    std::shared_ptr<Plan> plan_sp = std::make_shared<Plan>(); // shared (read-only) with the plan cache and with other handles of the same pattern
    bool plan_shared = false;
    LevelLists lv;
    int n = 0, nnz_in = 0, fnnz = 0;
    bool sym_lower = false;
    int effective_matching = 0, effective_pivoting = 5;

    // device: plan
    NodeDev* d_nodes = nullptr;
    int *d_rows = nullptr, *d_rel = nullptr, *d_child_idx = nullptr, *d_fact_nodes = nullptr, *d_solve_nodes = nullptr, *d_inv_nodes = nullptr;
    AsmItem* d_asm = nullptr;
    PanelItem* d_panel = nullptr;
    SchurItem* d_schur = nullptr;
    SolveItem* d_big_items = nullptr;
    int* d_big_slot = nullptr;
    // persistent top-of-tree sweep (sweep_top.cuh)
    int top_max_nodes = 1600; // (measured optimum at config 2) levels with at most this many fronts belong to the persistent sweep region
    // bottom of the tree: one warp per small subtree (sweep_sub.cuh)
    int use_subtree = 1, subtree_maxf = 96, subtree_budget = 0; // (0: the variant's default)
    // eligibility: every front f <= maxf, p <= 32, stored L entries of the subtree <= budget
    std::vector<char> in_sub;   // per front: handled by a subtree CTA in the solve phase
    int n_subtrees = 0;
    SubtreeDev* d_subtrees = nullptr; // descriptors, largest first (sweep_sub.cuh)
    unsigned short* d_st_tgt = nullptr;
    uchar2* d_st_pu = nullptr;
    bool fac_cleared = false;      // the host entry points clear the factor arena on the side stream, under their H2D copy
    std::vector<int> inv_skip_ptr; // NIC+1: fronts whose inverses are skipped, by class, inside d_inv_skip
    int* d_inv_skip = nullptr;
    unsigned long long* d_trace = nullptr; // optional per-item timestamps of the persistent sweeps (option "trace")
    int want_trace = 0;
    int *d_node_slot = nullptr, *d_bdone = nullptr; // k_bwd_top3: scratch slot base per front, partial-products-done counters
    int use_top = 1, ltop = 0, n_top_items = 0, top_grid = 0, top_grid_b = 0, n_slots = 0; // (the backward kernel needs less shared memory: its own, larger co-resident grid)
    bool sweep_dirty = false;          // a persistent sweep aborted: counters must be re-armed
    std::vector<int> cdone_init;       // host copy of the initial completion counters
    SolveItem* d_top_items = nullptr;
    int *d_top_ranges = nullptr, *d_top_slot = nullptr, *d_cdone = nullptr, *d_xdone = nullptr, *d_epoch = nullptr, *d_abort = nullptr;
    int *d_asm_ranges = nullptr, *d_big_ranges = nullptr;
    double* d_big_scratch = nullptr;
    int* d_big_tickets = nullptr;
    int* d_a_src = nullptr;
    long long* d_a_dst = nullptr;
    double* d_a_scl = nullptr;
    int *d_rowperm = nullptr, *d_colperm = nullptr;
    double *d_rscale = nullptr, *d_cscale = nullptr;
    int *d_full_ptr = nullptr, *d_full_col = nullptr, *d_full_src = nullptr, *d_rowblk = nullptr;
    int n_rowblk = 0;
    // device: numeric
    double *d_vals = nullptr, *d_fullvals = nullptr;
    // COO-level boundary (solver_b200_initialize_coo): triplet -> CSR-slot map and a staging buffer for raw triplet values
    int nnz_coo = 0;
    int *d_seg_ptr = nullptr, *d_seg_idx = nullptr;
    double* d_coo_vals = nullptr;
    CooGuard coo_guard; // host copy of the triplet indices the slot map was built from (solver_b200_factorize_coo_checked)
    const double* spmv_vals = nullptr; // mirrored values (symmetric input) or d_vals (general input)
    double *d_fac = nullptr, *d_cb = nullptr, *d_dinv = nullptr, *d_upiv = nullptr;
    int* d_lperm = nullptr;
    int* d_counters = nullptr; // 4 ints
    unsigned long long* d_amax = nullptr;
    // device: vectors
    double *d_b = nullptr, *d_x = nullptr, *d_r = nullptr, *d_y = nullptr, *d_z = nullptr, *d_xp = nullptr, *d_wv = nullptr;
    double *d_partial = nullptr, *d_norms = nullptr;
    double* h_norms = nullptr; // pinned, 2 doubles + counters
    int* h_counters = nullptr; // pinned

    cudaGraphExec_t g_fact = nullptr, g_sweep = nullptr;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t side = nullptr;          // side stream: the factor arena is cleared there, under the H2D copy of the new values
    cudaEvent_t ev_clr0 = nullptr, ev_clr1 = nullptr; // fork / join of the factor-arena clear that runs under the H2D copy
    // The launches of one tree level are independent of each other (fused size classes / buckets, and the big-front sequence
    // assembly -> pivot block -> panels -> Schur): they are enqueued on parallel branches (captured into the graph as a fork /
    // join per level), so that a launch of a few dozen fronts -- which costs the latency of one front, 15-35 us, however
    // small it is -- runs beside the large launches of its level instead of in front of them.
    static const int NLS = 3;
    cudaStream_t lvl_side[NLS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[NLS] = {nullptr, nullptr, nullptr};
    int use_level_fork = 1;
    // The explicit pivot-block inverses (and the packed pivot blocks of the subtree fronts) are needed by the solve phase
    // only.  Those of the fronts below `inv_split_level` -- the level from which every level holds just a few big fronts, the
    // latency-bound chain at the top of the tree -- are computed on a low-priority branch forked at that level, in the
    // shadow of the chain; the fronts of the chain itself follow after the level loop as before.
    cudaStream_t inv_side = nullptr;
    cudaEvent_t ev_inv0 = nullptr, ev_inv1 = nullptr;
    int inv_overlap = 1;      // option "inv_overlap" / B200_INV_OVERLAP
    int schur_front_nt = 64;  // option "schur_front_nt": fronts with at least this many rows of 64 x 64 tiles (u > 4032: 4096+ tiles, a full
                              // GPU by themselves) get a Schur launch of their own.  (16 was measured at config 2: its top chain links then
                              // leave the shared launch of their level, 5.70 -> 6.08 ms, profiles/r4p_schur_front.txt)
    int inv_split_level = -1; // -1: no early branch
    bool pack_early = false;  // the subtree fronts all lie below the split level
    int fused_variant = 2;  // 0 = shared-memory LU (k_front_fused), 1 = register-resident (k_front_fused_w8) for f <= 64,
                            // 2 = register-resident only for launches of at most fused_w8_max fronts (measured crossover)
    int fused_w8_max = 2000;
    int use_front_warp = 1; // fused launches of more than fused_w8_max fronts of order <= 64: one warp per front (k_front_warp)

    // stats
    int n_perturbed = 0;
    double rcond = -1.0;      // last value computed by solver_b200_rcond (-1: not computed for these factors)
    double t_init_host = 0.0; // wall time of the host analysis inside initialize
    int plan_cache_hit = 0;   // 1: this handle's plan came from the process-wide plan cache
    double last_rel_residual = -1.0, last_backward_error = -1.0;
    int last_refine_steps = 0;
    float ms_factorize = 0, ms_solve = 0, ms_sptrsv = 0, ms_spmv = 0;
    int launches_factorize = 0, launches_solve = 0, sweep_launches = 0;
    double sptrsv_bytes = 0, spmv_bytes = 0;
};

namespace {

// ---- plan cache: identical sparsity patterns are analysed once per process ("symbolic analysis shared on the
// host when patterns are identical" -- a sweep of LinSolver::new over one mesh, the handles of a parameter study, one handle
// per GPU driven from one process).  Key = two independent 64-bit hashes of (n, row pointers, column indices) + the analysis
// options; only plans WITHOUT matching / scaling are shared (those depend on the pattern alone), and a request is served from
// the cache only if its own values would not trigger the matching either.  B200_PLAN_CACHE=0 turns it off.
struct PlanCacheEntry {
    uint64_t h1, h2;
    int n, nnz;
    bool sym_lower;
    AnalyzeOptions opt;
    std::shared_ptr<const Plan> plan;
};
std::mutex g_plan_mu;
std::vector<PlanCacheEntry> g_plan_cache; // most recent last, at most 3 entries

// runs fn(chunk) for chunk = 0 .. nchunks-1 on a few threads (chunk boundaries are fixed by the caller, so results that are
// combined in chunk order do not depend on the number of threads)
template <class F>
void for_chunks(int nchunks, F fn) {
    unsigned nt = std::min<unsigned>(std::thread::hardware_concurrency(), 8u);
    if (nt < 2 || nchunks < 4) {
        for (int c = 0; c < nchunks; c++) fn(c);
        return;
    }
    std::atomic<int> next{0};
    auto work = [&]() {
        for (int c = next++; c < nchunks; c = next++) fn(c);
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
}

// two independent 64-bit hashes of (n, row pointers, column indices): fixed chunks of 2^18 entries are hashed independently
// (in parallel for large patterns) and their results chained in chunk order
void hash_pattern(int n, const int* rp, const int* ci, uint64_t& h1, uint64_t& h2) {
    struct H {
        uint64_t a, b;
        void mix(uint64_t v) {
            a = (a ^ v) * 0xff51afd7ed558ccdull, a ^= a >> 29;
            b = (b + v) * 0xc4ceb9fe1a85ec53ull, b ^= b >> 31;
        }
    };
    const int CH = 1 << 18;
    const int nnz = rp[n];
    const int c_rp = (int)(((long long)n + CH) / CH), c_ci = (int)(((long long)nnz + CH - 1) / CH);
    std::vector<H> part((size_t)c_rp + c_ci);
    for_chunks(c_rp + c_ci, [&](int c) {
        H h{0x9e3779b97f4a7c15ull ^ (uint64_t)c, 0xc2b2ae3d27d4eb4full + (uint64_t)c};
        if (c < c_rp) {
            const int lo = c * CH, hi = (int)std::min<long long>((long long)n + 1, (long long)lo + CH);
            for (int i = lo; i < hi; i++) h.mix((uint32_t)rp[i]);
        } else {
            const int lo = (c - c_rp) * CH, hi = (int)std::min<long long>(nnz, (long long)lo + CH);
            int k = lo;
            for (; k + 1 < hi; k += 2) h.mix(((uint64_t)(uint32_t)ci[k] << 32) | (uint32_t)ci[k + 1]);
            if (k < hi) h.mix((uint32_t)ci[k]);
        }
        part[c] = h;
    });
    H tot{0x9e3779b97f4a7c15ull ^ (uint64_t)n, 0xc2b2ae3d27d4eb4full + (uint64_t)n};
    for (const H& h : part) tot.mix(h.a), tot.mix(h.b);
    h1 = tot.a, h2 = tot.b;
}
bool same_options(const AnalyzeOptions& x, const AnalyzeOptions& y) {
    return x.ordering == y.ordering && x.panel_width == y.panel_width && x.nd_leaf == y.nd_leaf && x.relax_small == y.relax_small &&
           x.relax_z1 == y.relax_z1 && x.relax_z2 == y.relax_z2 && x.relax_z3 == y.relax_z3 && x.st_enable == y.st_enable &&
           x.st_maxf == y.st_maxf && x.st_pmax == y.st_pmax && x.st_budget == y.st_budget && x.st_maxcols == y.st_maxcols &&
           x.st_min_count == y.st_min_count && x.cb_reuse == y.cb_reuse;
}
// would these values make the analysis run the matching?  (plan.hpp: matching = 1 always, 2 = only if a diagonal entry is
// structurally or numerically zero)
bool values_need_matching(int matching, int n, const int* rp, const int* ci, const double* vals) {
    if (matching == 0) return false;
    if (matching == 1) return true;
    const int CH = 1 << 16;
    std::atomic<int> weak{0};
    for_chunks((n + CH - 1) / CH, [&](int c) {
        const int lo = c * CH, hi = (int)std::min<long long>(n, (long long)lo + CH);
        for (int i = lo; i < hi && !weak.load(std::memory_order_relaxed); i++) {
            bool ok = false;
            for (int k = rp[i]; k < rp[i + 1] && !ok; k++) ok = ci[k] == i && vals[k] != 0.0;
            if (!ok) weak = 1;
        }
    });
    return weak != 0;
}

// time spent in the allocations / copies of upload() by the calling thread (printed by initialize when verbose)
thread_local double g_up_malloc_s = 0.0, g_up_copy_s = 0.0, g_up_bytes = 0.0;
// the slab set device allocations of the calling thread go to (set for the duration of solver_b200_initialize / release_device)
thread_local SlabSet* g_slab = nullptr;
struct SlabScope {
    explicit SlabScope(SlabSet* sl) { g_slab = sl; }
    ~SlabScope() { g_slab = nullptr; }
};
cudaError_t device_alloc(void** out, size_t bytes) { return g_slab ? g_slab->alloc(out, bytes) : cudaMalloc(out, bytes); }

template <typename T>
cudaError_t upload(T** dptr, const std::vector<T>& h) {
    *dptr = nullptr;
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = device_alloc((void**)dptr, bytes);
    const auto t1 = std::chrono::steady_clock::now();
    g_up_malloc_s += std::chrono::duration<double>(t1 - t0).count();
    if (e != cudaSuccess) return e;
    if (!h.empty()) e = cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    g_up_copy_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    g_up_bytes += (double)bytes;
    return e;
}

template <typename T>
void dfree(T*& p) {
    if (p && !(g_slab && g_slab->owns(p))) cudaFree(p); // (chunk memory is released as a whole)
    p = nullptr;
}

void release_device(InterfaceB200* s) {
    SlabScope scope(&s->slabs);
    if (s->g_fact) cudaGraphExecDestroy(s->g_fact), s->g_fact = nullptr;
    if (s->g_sweep) cudaGraphExecDestroy(s->g_sweep), s->g_sweep = nullptr;
    dfree(s->d_nodes), dfree(s->d_rows), dfree(s->d_rel), dfree(s->d_child_idx), dfree(s->d_fact_nodes), dfree(s->d_solve_nodes), dfree(s->d_inv_nodes);
    dfree(s->d_asm), dfree(s->d_panel), dfree(s->d_schur);
    dfree(s->d_oz_split), dfree(s->d_oz_items), dfree(s->d_oz_tiles), dfree(s->d_oz_scales);
    dfree(s->d_trace);
    dfree(s->d_subtrees), dfree(s->d_st_tgt), dfree(s->d_st_pu);
    dfree(s->d_node_slot), dfree(s->d_bdone);
    dfree(s->d_inv_skip);
    dfree(s->d_big_items), dfree(s->d_big_slot), dfree(s->d_top_items), dfree(s->d_top_ranges), dfree(s->d_top_slot),
    dfree(s->d_cdone), dfree(s->d_xdone), dfree(s->d_epoch), dfree(s->d_abort), dfree(s->d_asm_ranges), dfree(s->d_big_ranges), dfree(s->d_big_scratch), dfree(s->d_big_tickets);
    dfree(s->d_a_src), dfree(s->d_a_dst), dfree(s->d_a_scl);
    dfree(s->d_rowperm), dfree(s->d_colperm), dfree(s->d_rscale), dfree(s->d_cscale);
    dfree(s->d_full_ptr), dfree(s->d_full_col), dfree(s->d_full_src), dfree(s->d_rowblk);
    dfree(s->d_vals), dfree(s->d_fullvals);
    dfree(s->d_seg_ptr), dfree(s->d_seg_idx), dfree(s->d_coo_vals);
    dfree(s->d_fac), dfree(s->d_cb), dfree(s->d_dinv), dfree(s->d_upiv), dfree(s->d_lperm);
    dfree(s->d_counters), dfree(s->d_amax);
    dfree(s->d_b), dfree(s->d_x), dfree(s->d_r), dfree(s->d_y), dfree(s->d_z), dfree(s->d_xp), dfree(s->d_wv);
    dfree(s->d_partial), dfree(s->d_norms);
    if (s->h_norms) cudaFreeHost(s->h_norms), s->h_norms = nullptr;
    if (s->h_counters) cudaFreeHost(s->h_counters), s->h_counters = nullptr;
    s->slabs.release();
    s->stripe_pool.shutdown();
    if (s->h_stage) cudaFreeHost(s->h_stage), s->h_stage = nullptr;
    for (int i = 0; i < 2 * InterfaceB200::NST; i++)
        if (s->ev_stage[i]) cudaEventDestroy(s->ev_stage[i]), s->ev_stage[i] = nullptr;
}

int grid_for(long long work, int block = 256, int cap = 148 * 16) {
    long long g = (work + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

// ---- work-item lists --------------------------------------------------------------------------------
size_t smem_fused(int f, int p) { return ((size_t)(f | 1) * f) * sizeof(double) + (size_t)p * sizeof(int); }

void build_work_lists(InterfaceB200* s, std::vector<AsmItem>& asm_items, std::vector<PanelItem>& panel_items,
                      std::vector<SchurItem>& schur_items, std::vector<int>& fact_nodes, std::vector<int>& solve_nodes,
                      std::vector<int>& asm_ranges, std::vector<OzakiSplitItem>& oz_split, std::vector<OzakiItem>& oz_items) {
    const Plan& P = (*s->plan_sp);
    LevelLists& lv = s->lv;
    lv.schur_fronts.clear();
    lv.schur_front_ptr.assign(P.nlevels + 1, 0);
    lv.asm_ptr.assign(P.nlevels + 1, 0);
    lv.panel_ptr.assign(P.nlevels + 1, 0);
    lv.schur_ptr.assign(P.nlevels + 1, 0);
    lv.fact_ptr.assign((size_t)P.nlevels * (NFC + 1) + 1, 0);
    lv.solve_ptr.assign((size_t)P.nlevels * NSC + 1, 0);
    lv.solve_threads.assign(P.nlevels, 32);
    lv.solve_pmax.assign(P.nlevels, 1);
    lv.fused_smem.assign((size_t)P.nlevels * NFC, 0);
    lv.fw_batches.clear();
    lv.fw_ptr.assign(1, 0);
    lv.asm_smem.assign(P.nlevels, 0);
    lv.oz_split_ptr.assign(P.nlevels + 1, 0), lv.oz_item_ptr.assign(P.nlevels + 1, 0), lv.oz_rows_max.assign(P.nlevels, 0);
    lv.oz_tile_bytes = 0, lv.oz_scales = 0;
    auto ozaki = [&](int v) { return s->schur_variant == 2 && P.u[v] >= s->ozaki_min_u; };
    auto fclass = [&](int f) {
        if (!s->use_fused || f > s->fused_maxf) return NFC;
        for (int c = 0; c < NFC; c++)
            if (f <= FC_MAXF[c]) return c;
        return NFC;
    };
    // a front is fed by the fused Schur epilogue of its single child when it is a later panel of a split supernode
    // (update set of the child == this front, identity relative indices) and both run on the multi-kernel path
    auto chain_fused = [&](int v) {
        if (!s->fuse_chain || s->schur_variant < 1) return false;
        if (P.child_ptr[v + 1] - P.child_ptr[v] != 1) return false;
        const int c = P.child_idx[P.child_ptr[v]];
        if (ozaki(c)) return false; // the tcgen05 kernel updates its own contribution block; the parent assembles it
        if (P.u[c] != P.p[v] + P.u[v]) return false;
        if (fclass(P.p[v] + P.u[v]) != NFC || fclass(P.p[c] + P.u[c]) != NFC) return false;
        const int* rel = &P.rel[P.rows_ptr[c]];
        for (int i = 0; i < P.u[c]; i++)
            if (rel[i] != i) return false;
        return true;
    };
    auto sclass = [&](int f) {
        for (int c = 0; c < NSC; c++)
            if (f <= SC_MAXF[c]) return c;
        return NSC - 1;
    };
    std::vector<char> diag_fused(P.nnodes, 0); // pivot block factorized by the chain child's Schur CTA (set at the child's level)
    lv.diag_cnt.assign(P.nlevels, 0);
    for (int l = 0; l < P.nlevels; l++) {
        size_t lvl_tiles = 0, lvl_scales = 0; // the sliced operands live for one level only
        for (int c = 0; c <= NFC; c++) {
            for (int pass = 0; pass < (c == NFC ? 2 : 1); pass++) // big fronts: those that need a pivot-block launch first
            for (int e = P.level_ptr[l]; e < P.level_ptr[l + 1]; e++) {
                const int v = P.level_nodes[e];
                const int f = P.p[v] + P.u[v];
                if (fclass(f) != c) continue;
                if (c == NFC && (diag_fused[v] != 0) != (pass == 1)) continue;
                if (c == NFC && pass == 0) lv.diag_cnt[l]++;
                fact_nodes.push_back(v);
                if (c < NFC) {
                    size_t& sm = lv.fused_smem[(size_t)l * NFC + c];
                    sm = std::max(sm, smem_fused(f, P.p[v]));
                }
            }
            lv.fact_ptr[(size_t)l * (NFC + 1) + c + 1] = (int)fact_nodes.size();
            if (c < NFC) { // size buckets of the warp-per-front kernel (fronts of one launch are independent: any order will do)
                const int beg = lv.fact_ptr[(size_t)l * (NFC + 1) + c], end = (int)fact_nodes.size();
                std::stable_sort(fact_nodes.begin() + beg, fact_nodes.begin() + end,
                                 [&P](int a, int b) { return P.p[a] + P.u[a] < P.p[b] + P.u[b]; });
                const size_t first = lv.fw_batches.size();
                int b0 = beg;
                while (b0 < end) {
                    const int f0 = P.p[fact_nodes[b0]] + P.u[fact_nodes[b0]];
                    const int cap = ((f0 + 3) / 4) * 4 + (f0 > 32 ? 4 : 0); // buckets of 4 (8 above 32)
                    int b1 = b0, fmax = f0;
                    while (b1 < end && P.p[fact_nodes[b1]] + P.u[fact_nodes[b1]] <= cap) fmax = P.p[fact_nodes[b1]] + P.u[fact_nodes[b1]], b1++;
                    // a launch costs the latency of one front (tens of microseconds) however few fronts it has: a sliver joins
                    // the bucket before it (whose shared memory then follows the sliver's larger fronts)
                    if (b1 - b0 < 1024 && lv.fw_batches.size() > first) {
                        lv.fw_batches.back().count += b1 - b0;
                        lv.fw_batches.back().need = (fmax | 1) * fmax;
                    } else
                        lv.fw_batches.push_back({b0, b1 - b0, (fmax | 1) * fmax});
                    b0 = b1;
                }
            }
            if (c < NFC) lv.fw_ptr.push_back((int)lv.fw_batches.size());
        }
        {
            // solve phase: ONE launch per level for all fronts below the "big" threshold (block size chosen from the
            // level's largest such front), plus the sliced launch for big fronts
            int maxf = 0, maxp = 0;
            for (int e = P.level_ptr[l]; e < P.level_ptr[l + 1]; e++) {
                const int v = P.level_nodes[e];
                const int f = P.p[v] + P.u[v];
                if (s->in_sub[v]) continue; // walked by a subtree CTA
                if (sclass(f) < NSC - 1) maxf = std::max(maxf, f), maxp = std::max(maxp, P.p[v]);
            }
            lv.solve_threads[l] = maxf <= 32 ? 32 : (maxf <= 64 ? 64 : 128);
            lv.solve_pmax[l] = std::max(maxp, 1);
            for (int c = 0; c < NSC; c++) {
                for (int e = P.level_ptr[l]; e < P.level_ptr[l + 1]; e++) {
                    const int v = P.level_nodes[e];
                    const int cls = sclass(P.p[v] + P.u[v]) < NSC - 1 ? 0 : NSC - 1; // everything small goes to slot 0
                    if (cls == c && !s->in_sub[v]) solve_nodes.push_back(v);
                }
                lv.solve_ptr[(size_t)l * NSC + c + 1] = (int)solve_nodes.size();
            }
        }
        for (int e = P.level_ptr[l]; e < P.level_ptr[l + 1]; e++) {
            const int v = P.level_nodes[e];
            const int p = P.p[v], u = P.u[v], f = p + u;
            if (fclass(f) != NFC) continue; // fused fronts need no work items
            const int nch = P.child_ptr[v + 1] - P.child_ptr[v];
            if (!chain_fused(v)) {
                double total = (double)u * u * 0.25; // clearing the contribution block counts as work too
                for (int c = P.child_ptr[v]; c < P.child_ptr[v + 1]; c++) {
                    double uc = P.u[P.child_idx[c]];
                    total += uc * uc;
                }
                int ntiles = (int)std::ceil(total / 4096.0);
                ntiles = std::max(1, std::min(ntiles, f));
                int tw = (f + ntiles - 1) / ntiles;
                if (s->asm_variant == 1) tw = std::max(1, std::min(tw, B200_ASM_TILE / f)); // the tile lives in shared memory
                lv.asm_smem[l] = std::max(lv.asm_smem[l], (size_t)tw * f * sizeof(double));
                for (int t0 = 0; t0 < f; t0 += tw) {
                    const int t1 = std::min(f, t0 + tw);
                    asm_items.push_back({v, t0, t1, (int)asm_ranges.size()});
                    for (int c = P.child_ptr[v]; c < P.child_ptr[v + 1]; c++) { // child columns landing in [t0, t1)
                        const int ch = P.child_idx[c];
                        const int* rel = &P.rel[P.rows_ptr[ch]];
                        asm_ranges.push_back((int)(std::lower_bound(rel, rel + P.u[ch], t0) - rel));
                        asm_ranges.push_back((int)(std::lower_bound(rel, rel + P.u[ch], t1) - rel));
                    }
                }
            }
            if (u > 0) {
                const int TR = (s->panel_variant >= 1) ? B200_PW_ROWS : B200_TR;
                for (int r0 = 0; r0 < u; r0 += TR) panel_items.push_back({v, r0, std::min(TR, u - r0), 0});
                for (int r0 = 0; r0 < u; r0 += TR) panel_items.push_back({v, r0, std::min(TR, u - r0), 1});
                if (ozaki(v) && fclass(f) == NFC) {
                    // operands split into int8 slices (tile-canonical layout), then one item per 128 x 64 tile of C
                    const int kch = (p + OZ_KC - 1) / OZ_KC;
                    const int nta = (u + OZ_BM - 1) / OZ_BM, ntb = (u + OZ_BN - 1) / OZ_BN;
                    const long long a_t = (long long)lvl_tiles, b_t = a_t + (long long)nta * kch * OZ_S * OZ_A_BYTES;
                    lvl_tiles = (size_t)b_t + (size_t)ntb * kch * OZ_S * OZ_B_BYTES;
                    const long long a_s = (long long)lvl_scales, b_s = a_s + u;
                    lvl_scales += 2 * (size_t)u;
                    oz_split.push_back({P.Loff[v] + p, a_t, a_s, u, p, (int)f, OZ_BM});
                    oz_split.push_back({P.Uoff[v], b_t, b_s, u, p, u, OZ_BN});
                    lv.oz_rows_max[l] = std::max(lv.oz_rows_max[l], nta * OZ_BM);
                    for (int tj = 0; tj < ntb; tj++)
                        for (int ti = 0; ti < nta; ti++) oz_items.push_back({a_t, b_t, a_s, b_s, P.Coff[v], u, kch, ti, tj});
                } else {
                    int nt = (u + B200_TS - 1) / B200_TS;
                    const int par = P.parent[v];
                    const int fuse_into = (par >= 0 && chain_fused(par)) ? par : -1;
                    const bool la = fuse_into >= 0 && s->fuse_diag && s->diag_variant == 4 && s->schur_variant >= 1;
                    if (la) diag_fused[par] = 1;
                    if (nt >= s->schur_front_nt && s->schur_variant >= 1) // a launch of its own, tiles from the grid
                        lv.schur_fronts.push_back({v, nt, fuse_into, la ? 1 : 0});
                    else
                        for (int tj = 0; tj < nt; tj++)
                            for (int ti = 0; ti < nt; ti++)
                                schur_items.push_back({v, ti, tj, (la && ti == 0 && tj == 0) ? (fuse_into | B200_SCHUR_DIAG) : fuse_into});
                }
            }
        }
        lv.asm_ptr[l + 1] = (int)asm_items.size();
        lv.panel_ptr[l + 1] = (int)panel_items.size();
        lv.schur_ptr[l + 1] = (int)schur_items.size();
        lv.schur_front_ptr[l + 1] = (int)lv.schur_fronts.size();
        lv.oz_split_ptr[l + 1] = (int)oz_split.size(), lv.oz_item_ptr[l + 1] = (int)oz_items.size();
        lv.oz_tile_bytes = std::max(lv.oz_tile_bytes, lvl_tiles), lv.oz_scales = std::max(lv.oz_scales, lvl_scales);
    }
}

size_t smem_diag(int) { return (size_t)(B200_MAXP * (B200_MAXP + 1)) * sizeof(double) + B200_MAXP * sizeof(int); }
size_t smem_invert(int pmax) { return (size_t)(2 * pmax * (pmax | 1)) * sizeof(double); }
size_t smem_panel(int W) { return (size_t)(W * W + B200_TR * W) * sizeof(double) + W * sizeof(int); }
size_t smem_schur_fma(int W) { return (size_t)2 * W * B200_TS * sizeof(double); }
size_t smem_schur_dmma() { return (size_t)2 * B200_MAXP * (B200_TS + 8) * sizeof(double); }

// enqueue the per-level numeric kernels: fused fronts (one launch per size class), then the big-front path
// (assembly -> pivot block -> panels -> Schur complement)
int enqueue_levels(InterfaceB200* s, int* launches) {
    const Plan& P = (*s->plan_sp);
    const LevelLists& lv = s->lv;
    const int W = s->opt_panel_width;
    int cnt = 0;
    auto invert_range = [&](int c, int a, int b, cudaStream_t st) {
        if (b > a) {
            k_invert_col<<<b - a, 2 * IC_MAXP[c], smem_invert(IC_MAXP[c]), st>>>(s->d_inv_nodes + a, s->d_nodes, s->d_fac, s->d_dinv, IC_MAXP[c], b - a);
            cnt++;
        }
    };
    auto pack_blocks = [&](cudaStream_t st) {
        if (s->n_subtrees > 0 && !s->inv_skip_ptr.empty() && s->inv_skip_ptr[NIC] > 0) {
            // pivot blocks of the subtree fronts, packed for the subtree solve kernels (they keep no explicit inverses)
            const int nn = s->inv_skip_ptr[NIC];
            k_pack_pivot_blocks<<<std::min((nn + 7) / 8, 148 * 8), 256, 0, st>>>(s->d_inv_skip, nn, s->d_nodes, s->d_fac, s->d_dinv);
            cnt++;
        }
    };
    const bool early = s->inv_split_level > 0 && s->inv_split_level < P.nlevels && (int)lv.inv_mid.size() == NIC;
    for (int l = 0; l < P.nlevels; l++) {
        if (early && l == s->inv_split_level) {
            // every front below this level is final (a pivot block is factorized at its own level at the latest): their
            // inverses run on a low-priority branch while the chain levels above proceed
            cudaEventRecord(s->ev_inv0, s->stream);
            cudaStreamWaitEvent(s->inv_side, s->ev_inv0, 0);
            for (int c = NIC - 1; c >= 0; c--) invert_range(c, lv.inv_ptr[c], lv.inv_mid[c], s->inv_side); // the long class first
            if (s->pack_early) pack_blocks(s->inv_side);
            cudaEventRecord(s->ev_inv1, s->inv_side);
        }
        const int* fp = &lv.fact_ptr[(size_t)l * (NFC + 1)];
        // independent launch groups of this level: fused classes / buckets + the big-front sequence
        int ngroups = (fp[NFC + 1] - fp[NFC] > 0) ? 1 : 0;
        for (int c = 0; c < NFC; c++)
            if (fp[c + 1] - fp[c] > 0) ngroups += std::max(1, lv.fw_ptr[(size_t)l * NFC + c + 1] - lv.fw_ptr[(size_t)l * NFC + c]);
        const bool fork = s->use_level_fork && ngroups >= 2;
        unsigned used = 0;
        int rr = 0;
        if (fork) cudaEventRecord(s->ev_fork, s->stream);
        auto fstream = [&]() -> cudaStream_t { // the branch of the next fused launch
            if (!fork) return s->stream;
            const int i = rr++ % InterfaceB200::NLS;
            if (!(used & (1u << i))) cudaStreamWaitEvent(s->lvl_side[i], s->ev_fork, 0), used |= 1u << i;
            return s->lvl_side[i];
        };
        auto join = [&]() {
            for (int i = 0; i < InterfaceB200::NLS; i++)
                if (used & (1u << i)) cudaEventRecord(s->ev_join[i], s->lvl_side[i]), cudaStreamWaitEvent(s->stream, s->ev_join[i], 0);
        };
        for (int c = 0; c < NFC; c++) {
            int nn = fp[c + 1] - fp[c];
            if (nn > 0) {
                if (FC_MAXF[c] <= 32 && l == 0 && s->use_leaf_reg) {
                    // leaves: the whole front in one warp's registers
                    const int gridw = (nn + B200_LEAF_WARPS - 1) / B200_LEAF_WARPS;
                    if (FC_MAXF[c] <= 16)
                        k_small_reg<16><<<gridw, 32 * B200_LEAF_WARPS, 0, fstream()>>>(s->d_fact_nodes + fp[c], nn, s->d_nodes, s->d_child_idx,
                                                                                      s->d_rel, s->d_fac, s->d_cb, s->d_lperm, s->d_upiv,
                                                                                      s->d_amax, s->pivot_eps, s->d_counters);
                    else
                        k_small_reg<32><<<gridw, 32 * B200_LEAF_WARPS, 0, fstream()>>>(s->d_fact_nodes + fp[c], nn, s->d_nodes, s->d_child_idx,
                                                                                      s->d_rel, s->d_fac, s->d_cb, s->d_lperm, s->d_upiv,
                                                                                      s->d_amax, s->pivot_eps, s->d_counters);
                }
                // register-resident LU (one warp per 8 front columns): wins when the launch is latency bound (few fronts),
                // loses to the leaner shared-memory kernel when tens of thousands of fronts compete for thread slots
                else if (FC_MAXF[c] <= 64 && (s->fused_variant == 1 || (s->fused_variant == 2 && nn <= s->fused_w8_max)))
                    k_front_fused_w8<<<nn, 32 * ((FC_MAXF[c] + 7) / 8), lv.fused_smem[(size_t)l * NFC + c], fstream()>>>(
                        s->d_fact_nodes + fp[c], s->d_nodes, s->d_child_idx, s->d_rel, s->d_fac, s->d_cb, s->d_lperm,
                        s->d_upiv, s->d_amax, s->pivot_eps, s->d_counters);
                else if (FC_MAXF[c] <= 64 && s->use_front_warp) {
                    // one warp per front, the front in the warp's slice of shared memory: one launch per size bucket
                    const int R = FC_MAXF[c] <= 32 ? 1 : 2;
                    for (int bi = lv.fw_ptr[(size_t)l * NFC + c]; bi < lv.fw_ptr[(size_t)l * NFC + c + 1]; bi++) {
                        const LevelLists::FwBatch& B = lv.fw_batches[bi];
                        const int wstride = ((B.need + 1) & ~1) + 32 * R; // front + permutation + one child's relative indices
                        const int gridw = (B.count + B200_FW_WARPS - 1) / B200_FW_WARPS;
                        const size_t smw = (size_t)wstride * B200_FW_WARPS * sizeof(double);
                        if (R == 1)
                            k_front_warp<1><<<gridw, 32 * B200_FW_WARPS, smw, fstream()>>>(s->d_fact_nodes + B.start, B.count, s->d_nodes, s->d_child_idx,
                                                                                        s->d_rel, s->d_fac, s->d_cb, s->d_lperm, s->d_upiv, s->d_amax,
                                                                                        s->pivot_eps, s->d_counters, wstride);
                        else
                            k_front_warp<2><<<gridw, 32 * B200_FW_WARPS, smw, fstream()>>>(s->d_fact_nodes + B.start, B.count, s->d_nodes, s->d_child_idx,
                                                                                        s->d_rel, s->d_fac, s->d_cb, s->d_lperm, s->d_upiv, s->d_amax,
                                                                                        s->pivot_eps, s->d_counters, wstride);
                        cnt++;
                    }
                    cnt--;
                } else
                    k_front_fused<<<nn, FC_THREADS[c], lv.fused_smem[(size_t)l * NFC + c], fstream()>>>(
                        s->d_fact_nodes + fp[c], s->d_nodes, s->d_child_idx, s->d_rel, s->d_fac, s->d_cb, s->d_lperm,
                        s->d_upiv, s->d_amax, s->pivot_eps, s->d_counters);
                cnt++;
            }
        }
        int nbig = fp[NFC + 1] - fp[NFC];
        if (nbig == 0) {
            join();
            continue;
        }
        int na = lv.asm_ptr[l + 1] - lv.asm_ptr[l];
        if (na > 0) {
            if (s->asm_variant == 1 && lv.asm_smem[l] <= (size_t)B200_ASM_SMEM_MAX) // tile in shared memory
                k_assemble_tile<<<na, 256, lv.asm_smem[l], s->stream>>>(s->d_asm + lv.asm_ptr[l], s->d_nodes, s->d_child_idx, s->d_rel,
                                                                      s->d_asm_ranges, s->d_fac, s->d_cb);
            else
                k_assemble<<<na, 256, 0, s->stream>>>(s->d_asm + lv.asm_ptr[l], s->d_nodes, s->d_child_idx, s->d_rel, s->d_asm_ranges, s->d_fac, s->d_cb);
            cnt++;
        }
        const int ndiag = lv.diag_cnt[l]; // big fronts whose pivot block was NOT factorized by their chain child's Schur CTA (they come first)
        if (ndiag > 0) {
            if (s->diag_variant == 4)
                k_diag_w8<<<ndiag, 256, 0, s->stream>>>(s->d_fact_nodes + fp[NFC], s->d_nodes, s->d_fac, s->d_lperm, s->d_upiv, s->d_amax,
                                                        s->pivot_eps, s->d_counters);
            else
                k_diag<<<ndiag, 512, smem_diag(W), s->stream>>>(s->d_fact_nodes + fp[NFC], s->d_nodes, s->d_fac,
                                                                s->d_lperm, s->d_upiv, s->d_amax, s->pivot_eps, s->d_counters);
            cnt++;
        }
        int np = lv.panel_ptr[l + 1] - lv.panel_ptr[l];
        if (np > 0) {
            if (s->panel_variant == 1 && np <= s->panel_row_max) // few rows: the launch is latency bound
                k_panel_row<<<np * 4, 256, 0, s->stream>>>(s->d_panel + lv.panel_ptr[l], s->d_nodes, s->d_fac, s->d_lperm);
            else if (s->panel_variant == 1)
                k_panel_warp<<<np, 128, B200_PW_SMEM, s->stream>>>(s->d_panel + lv.panel_ptr[l], s->d_nodes, s->d_fac, s->d_lperm);
            else
                k_panel<<<np, 256, smem_panel(W), s->stream>>>(s->d_panel + lv.panel_ptr[l], s->d_nodes, s->d_fac, s->d_lperm);
            cnt++;
        }
        const int noz = lv.oz_item_ptr[l + 1] - lv.oz_item_ptr[l];
        if (noz > 0) { // tcgen05 path: split the two panels into int8 slices, then the Schur tiles on the tensor cores
            const int nsp = lv.oz_split_ptr[l + 1] - lv.oz_split_ptr[l];
            k_ozaki_split<<<dim3((lv.oz_rows_max[l] + 127) / 128, nsp), 128, 0, s->stream>>>(s->d_oz_split + lv.oz_split_ptr[l], s->d_fac, s->d_oz_tiles,
                                                                                          s->d_oz_scales);
            k_schur_ozaki<<<std::min(noz, 148), 128, OZ_SMEM, s->stream>>>(s->d_oz_items + lv.oz_item_ptr[l], noz, s->d_oz_tiles, s->d_oz_scales, s->d_cb);
            cnt += 2;
        }
        int nsch = lv.schur_ptr[l + 1] - lv.schur_ptr[l];
        if (nsch > 0) {
            if (s->schur_variant >= 1)
                k_schur_dmma<<<nsch, 256, smem_schur_dmma(), s->stream>>>(s->d_schur + lv.schur_ptr[l], s->d_nodes, s->d_fac, s->d_cb, s->d_lperm, s->d_upiv,
                                                                          s->d_amax, s->pivot_eps, s->d_counters);
            else
                k_schur_fma<<<nsch, 256, smem_schur_fma(W), s->stream>>>(s->d_schur + lv.schur_ptr[l], s->d_nodes, s->d_fac, s->d_cb);
            cnt++;
        }
        for (int e = lv.schur_front_ptr[l]; e < lv.schur_front_ptr[l + 1]; e++) { // large fronts: one launch each, tiles from the grid
            const LevelLists::SchurFront& sf = lv.schur_fronts[e];
            k_schur_dmma_front<<<dim3(sf.nt, sf.nt), 256, smem_schur_dmma(), s->stream>>>(sf.node, sf.parent, sf.lookahead, s->d_nodes, s->d_fac, s->d_cb,
                                                                                          s->d_lperm, s->d_upiv, s->d_amax, s->pivot_eps, s->d_counters);
            cnt++;
        }
        join();
    }
    // explicit inverses of the pivot blocks, all fronts outside the subtree region in one batched launch per size class:
    // only the solve phase needs them (with the early branch: the fronts of the top levels only)
    for (int c = 0; c < NIC; c++) invert_range(c, early ? lv.inv_mid[c] : lv.inv_ptr[c], lv.inv_ptr[c + 1], s->stream);
    if (!(early && s->pack_early)) pack_blocks(s->stream);
    if (early) cudaStreamWaitEvent(s->stream, s->ev_inv1, 0); // join
    if (launches) *launches = cnt;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// one sweep = one epoch of the completion counters; the item tickets of both persistent kernels restart at zero
__global__ void k_bump_epoch(int* epoch) { epoch[0] += 1, epoch[1] = 0, epoch[2] = 0; }
// min and max of |U_kk| (IEEE bit patterns of non-negative doubles order like unsigned integers): UMFPACK's rcond estimate
__global__ void __launch_bounds__(256) k_minmax_abs(int n, const double* __restrict__ v, unsigned long long* __restrict__ mm) {
    unsigned long long lo = ~0ull, hi = 0ull;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v[k]));
        lo = b < lo ? b : lo, hi = b > hi ? b : hi;
    }
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long l2 = __shfl_down_sync(0xffffffffu, lo, off), h2 = __shfl_down_sync(0xffffffffu, hi, off);
        lo = l2 < lo ? l2 : lo, hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(mm, lo), atomicMax(mm + 1, hi);
}
void k_fwd_top_launch(InterfaceB200* s) {
    k_fwd_top2<<<s->top_grid, 256, B200_TOP3_SMEM, s->stream>>>(s->d_top_items, s->n_top_items, s->d_nodes, s->d_rel, s->d_fac, s->d_dinv,
                                                             s->d_lperm, s->d_top_ranges, s->d_y, s->d_z, s->d_wv, s->d_cdone, s->d_epoch,
                                                             s->d_abort, s->d_trace);
}
void k_bwd_top_launch(InterfaceB200* s) {
    k_bwd_top3<<<s->top_grid_b, 256, B200_TOP3_SMEM, s->stream>>>(s->d_top_items, s->n_top_items, s->d_nodes, s->d_rows, s->d_fac, s->d_dinv,
                                                              s->d_z, s->d_xp, s->d_big_scratch, s->d_node_slot, s->d_bdone, s->d_epoch,
                                                              s->d_abort, s->d_trace ? s->d_trace + 4 * (size_t)s->n_top_items : nullptr);
}

int enqueue_sweep_levels(InterfaceB200* s, int* launches) {
    const Plan& P = (*s->plan_sp);
    const LevelLists& lv = s->lv;
    int cnt = 0;
    const int lsplit = s->n_top_items > 0 ? s->ltop : P.nlevels; // levels >= lsplit run in the persistent kernels
    if (s->n_subtrees > 0) {
        k_fwd_stree_w<<<(s->n_subtrees + B200_SUBW_WARPS - 1) / B200_SUBW_WARPS, 32 * B200_SUBW_WARPS, 0, s->stream>>>(
            s->d_subtrees, s->n_subtrees, s->d_fac, s->d_st_tgt, s->d_st_pu, s->d_lperm, s->d_y, s->d_z, s->d_wv);
        cnt++;
    }
    for (int l = 0; l < lsplit; l++) {
        for (int c = 0; c < NSC - 1; c++) {
            int a = lv.solve_ptr[(size_t)l * NSC + c], b = lv.solve_ptr[(size_t)l * NSC + c + 1];
            if (b > a) {
                k_fwd<<<b - a, lv.solve_threads[l], (size_t)lv.solve_pmax[l] * lv.solve_pmax[l] * sizeof(double), s->stream>>>(s->d_solve_nodes + a, s->d_nodes, s->d_child_idx, s->d_rel, s->d_fac,
                                                             s->d_dinv, s->d_lperm, s->d_y, s->d_z, s->d_wv);
                cnt++;
            }
        }
        int nb = lv.big_ptr[l + 1] - lv.big_ptr[l];
        if (nb > 0) {
            k_fwd_big<<<nb, 256, 0, s->stream>>>(s->d_big_items + lv.big_ptr[l], s->d_nodes, s->d_child_idx, s->d_rel, s->d_fac, s->d_dinv,
                                                  s->d_lperm, s->d_big_ranges, s->d_y, s->d_z, s->d_wv);
            cnt++;
        }
    }
    if (s->n_top_items > 0) {
        k_bump_epoch<<<1, 1, 0, s->stream>>>(s->d_epoch);
        k_fwd_top_launch(s);
        k_bwd_top_launch(s);
        cnt += 3;
    }
    for (int l = lsplit - 1; l >= 0; l--) {
        for (int c = 0; c < NSC - 1; c++) {
            int a = lv.solve_ptr[(size_t)l * NSC + c], b = lv.solve_ptr[(size_t)l * NSC + c + 1];
            if (b > a) {
                k_bwd<<<b - a, lv.solve_threads[l], (size_t)lv.solve_pmax[l] * lv.solve_pmax[l] * sizeof(double), s->stream>>>(s->d_solve_nodes + a, s->d_nodes, s->d_rows, s->d_fac, s->d_dinv, s->d_z, s->d_xp);
                cnt++;
            }
        }
        int nb = lv.big_ptr[l + 1] - lv.big_ptr[l];
        if (nb > 0) {
            k_bwd_big<<<nb, 256, 0, s->stream>>>(s->d_big_items + lv.big_ptr[l], s->d_nodes, s->d_rows, s->d_fac, s->d_dinv, s->d_z, s->d_xp,
                                                  s->d_big_scratch, s->d_big_tickets, s->d_big_slot + lv.big_ptr[l]);
            cnt++;
        }
    }
    if (s->n_subtrees > 0) {
        k_bwd_stree_w<<<(s->n_subtrees + B200_SUBW_WARPS - 1) / B200_SUBW_WARPS, 32 * B200_SUBW_WARPS, 0, s->stream>>>(
            s->d_subtrees, s->n_subtrees, s->d_fac, s->d_dinv, s->d_st_tgt, s->d_st_pu, s->d_rows, s->d_z, s->d_xp);
        cnt++;
    }
    if (launches) *launches = cnt;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// runs `enqueue` either directly or through a captured graph (captured on first use)
template <typename F>
int run_maybe_graph(InterfaceB200* s, cudaGraphExec_t* exec, F enqueue, int* launches) {
    if (!s->use_graph) return enqueue(s, launches);
    if (*exec == nullptr) {
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            return enqueue(s, launches);
        }
        int rc = enqueue(s, launches);
        cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
        if (rc != 0 || e != cudaSuccess || graph == nullptr) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            s->use_graph = 0;
            if (s->verbose) fprintf(stderr, "solver_b200: graph capture failed, falling back to direct launches\n");
            return enqueue(s, launches);
        }
        e = cudaGraphInstantiate(exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            cudaGetLastError();
            *exec = nullptr;
            s->use_graph = 0;
            return enqueue(s, launches);
        }
    }
    return cudaGraphLaunch(*exec, s->stream) == cudaSuccess ? 0 : 1;
}

int sweep(InterfaceB200* s, const double* d_rhs, double* d_out, int accumulate, bool time_it) {
    const int n = s->n;
    k_permute_in<<<grid_for(n), 256, 0, s->stream>>>(n, s->d_rowperm, s->d_rscale, d_rhs, s->d_y);
    if (time_it) cudaEventRecord(s->ev[4], s->stream);
    int launches = 0;
    int rc = run_maybe_graph(s, &s->g_sweep, enqueue_sweep_levels, &launches);
    if (time_it) cudaEventRecord(s->ev[5], s->stream);
    k_permute_out<<<grid_for(n), 256, 0, s->stream>>>(n, s->d_colperm, s->d_cscale, s->d_xp, d_out, accumulate);
    if (launches > 0) s->sweep_launches = launches;
    s->launches_solve += 2 + s->sweep_launches;
    return rc;
}

int residual(InterfaceB200* s, const double* d_xv, const double* d_rhs, double* d_rout, bool time_it) {
    if (time_it) cudaEventRecord(s->ev[6], s->stream);
    k_spmv_stream<<<s->n_rowblk, 256, 0, s->stream>>>(s->d_rowblk, s->d_full_ptr, s->d_full_col, s->spmv_vals, d_xv, d_rhs,
                                                       d_rout, s->d_partial, 1);
    if (time_it) cudaEventRecord(s->ev[7], s->stream);
    k_reduce_partials<<<1, 256, 0, s->stream>>>(s->n_rowblk, s->d_partial, s->d_norms);
    s->launches_solve += 2;
    if (cudaMemcpyAsync(s->h_norms, s->d_norms, 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) return 1;
    return 0;
}

} // namespace

// =========================================================================================================
extern "C" {

void* solver_b200_get_stream(struct InterfaceB200* s) { return s ? (void*)s->stream : nullptr; }

int32_t solver_b200_get_device(struct InterfaceB200* s) { return s ? s->device : -1; }

const char* solver_b200_version(void) { return "solver_b200 0.1 (sm_100a, multifrontal LU f64)"; }

static bool create_streams(InterfaceB200* s) {
    int lo = 0, hi = 0; // numerically lower = higher priority
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, hi) != cudaSuccess) return false;
    if (cudaStreamCreateWithPriority(&s->side, cudaStreamNonBlocking, lo) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&s->ev_clr0, cudaEventDisableTiming) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&s->ev_clr1, cudaEventDisableTiming) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess) return false;
    if (cudaStreamCreateWithPriority(&s->inv_side, cudaStreamNonBlocking, lo) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&s->ev_inv0, cudaEventDisableTiming) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&s->ev_inv1, cudaEventDisableTiming) != cudaSuccess) return false;
    for (int i = 0; i < InterfaceB200::NLS; i++) {
        if (cudaStreamCreateWithPriority(&s->lvl_side[i], cudaStreamNonBlocking, hi) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&s->ev_join[i], cudaEventDisableTiming) != cudaSuccess) return false;
    }
    return true;
}
static void destroy_streams(InterfaceB200* s) {
    if (s->ev_clr0) cudaEventDestroy(s->ev_clr0), s->ev_clr0 = nullptr;
    if (s->ev_clr1) cudaEventDestroy(s->ev_clr1), s->ev_clr1 = nullptr;
    if (s->ev_fork) cudaEventDestroy(s->ev_fork), s->ev_fork = nullptr;
    if (s->ev_inv0) cudaEventDestroy(s->ev_inv0), s->ev_inv0 = nullptr;
    if (s->ev_inv1) cudaEventDestroy(s->ev_inv1), s->ev_inv1 = nullptr;
    if (s->inv_side) cudaStreamDestroy(s->inv_side), s->inv_side = nullptr;
    for (int i = 0; i < InterfaceB200::NLS; i++) {
        if (s->ev_join[i]) cudaEventDestroy(s->ev_join[i]), s->ev_join[i] = nullptr;
        if (s->lvl_side[i]) cudaStreamDestroy(s->lvl_side[i]), s->lvl_side[i] = nullptr;
    }
    if (s->side) cudaStreamDestroy(s->side), s->side = nullptr;
    if (s->stream) cudaStreamDestroy(s->stream), s->stream = nullptr;
}

struct InterfaceB200* solver_b200_new(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return nullptr; // no CPU fallback, by design
    }
    InterfaceB200* s = new (std::nothrow) InterfaceB200();
    if (!s) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const char* edev = getenv("B200_DEVICE"); // one process per GPU: the launcher may pin the device explicitly
    if (edev && atoi(edev) >= 0 && atoi(edev) < ndev) dev = atoi(edev);
    s->device = dev;
    if (cudaSetDevice(dev) != cudaSuccess || !create_streams(s)) {
        destroy_streams(s);
        delete s;
        return nullptr;
    }
    for (int i = 0; i < 8; i++)
        if (cudaEventCreate(&s->ev[i]) != cudaSuccess) {
            solver_b200_drop(s);
            return nullptr;
        }
    const char* e;
    if ((e = getenv("B200_NO_GRAPH")) && atoi(e)) s->use_graph = 0;
    if ((e = getenv("B200_SCHUR_VARIANT"))) s->schur_variant = atoi(e);
    if ((e = getenv("B200_PANEL_VARIANT"))) s->panel_variant = atoi(e);
    if ((e = getenv("B200_PANEL_ROW_MAX"))) s->panel_row_max = atoi(e);
    if ((e = getenv("B200_USE_FRONT_WARP"))) s->use_front_warp = atoi(e);
    if ((e = getenv("B200_USE_LEAF_REG"))) s->use_leaf_reg = atoi(e);
    if ((e = getenv("B200_USE_LEVEL_FORK"))) s->use_level_fork = atoi(e);
    if ((e = getenv("B200_INV_OVERLAP"))) s->inv_overlap = atoi(e);
    if ((e = getenv("B200_STAGED_COPY"))) s->staged_copy = atoi(e);
    if ((e = getenv("B200_FUSED_W8_MAX"))) s->fused_w8_max = atoi(e);
    if ((e = getenv("B200_USE_LEAF_REG"))) s->use_leaf_reg = atoi(e);
    if ((e = getenv("B200_ASM_VARIANT"))) s->asm_variant = atoi(e);
    if ((e = getenv("B200_FUSED_VARIANT"))) s->fused_variant = atoi(e);
    if ((e = getenv("B200_USE_FUSED"))) s->use_fused = atoi(e);
    if ((e = getenv("B200_USE_TOP"))) s->use_top = atoi(e);
    if ((e = getenv("B200_USE_SUBTREE"))) s->use_subtree = atoi(e);
    if ((e = getenv("B200_SUBTREE_MAXF"))) s->subtree_maxf = atoi(e);
    if ((e = getenv("B200_SUBTREE_BUDGET"))) s->subtree_budget = atoi(e);
    if ((e = getenv("B200_DIAG_VARIANT"))) s->diag_variant = atoi(e);
    if ((e = getenv("B200_FUSE_CHAIN"))) s->fuse_chain = atoi(e);
    if ((e = getenv("B200_FUSED_MAXF"))) s->fused_maxf = std::max(0, std::min(atoi(e), B200_FUSED_MAXF));
    if ((e = getenv("B200_PANEL_WIDTH"))) s->opt_panel_width = atoi(e);
    if ((e = getenv("B200_ND_LEAF"))) s->opt_nd_leaf = atoi(e);
    return s;
}

void solver_b200_drop(struct InterfaceB200* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    release_device(s);
    for (int i = 0; i < 8; i++)
        if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    destroy_streams(s);
    delete s;
}

int32_t solver_b200_set_option(struct InterfaceB200* s, const char* key, double value) {
    if (!s || !key) return B200_ERROR_NULL_POINTER;
    std::string k(key);
    if (k == "ir_tol") { s->ir_tol = value; return 0; }
    if (k == "refinement_nstep") { s->nrefine = (int)value; return 0; }
    if (k == "strict_residual") { s->strict_residual = value != 0.0; return 0; }
    if (k == "staged_copy") { s->staged_copy = value != 0.0; return 0; }
    if (s->initialized) return B200_ERROR_ALREADY_INITIALIZED;
    if (k == "panel_width") s->opt_panel_width = std::max(4, std::min((int)value, B200_MAXP));
    else if (k == "nd_leaf") s->opt_nd_leaf = std::max(4, (int)value);
    else if (k == "use_graph") s->use_graph = value != 0.0;
    else if (k == "schur_variant") s->schur_variant = (int)value;
    else if (k == "ozaki_min_u") s->ozaki_min_u = std::max(1, (int)value);
    else if (k == "panel_variant") s->panel_variant = (int)value;
    else if (k == "panel_row_max") s->panel_row_max = (int)value;
    else if (k == "use_leaf_reg") s->use_leaf_reg = value != 0.0;
    else if (k == "asm_variant") s->asm_variant = (int)value;
    else if (k == "relax_small") s->relax_small = (int)value;
    else if (k == "relax_z1") s->relax_z1 = value;
    else if (k == "relax_z2") s->relax_z2 = value;
    else if (k == "relax_z3") s->relax_z3 = value;
    else if (k == "fused_variant") s->fused_variant = (int)value;
    else if (k == "fused_w8_max") s->fused_w8_max = (int)value;
    else if (k == "use_front_warp") s->use_front_warp = value != 0.0;
    else if (k == "use_level_fork") s->use_level_fork = value != 0.0;
    else if (k == "inv_overlap") s->inv_overlap = value != 0.0;
    else if (k == "schur_front_nt") s->schur_front_nt = std::max(1, (int)value);
    else if (k == "use_fused") s->use_fused = value != 0.0;
    else if (k == "use_top") s->use_top = value != 0.0;
    else if (k == "trace") s->want_trace = value != 0.0;
    else if (k == "use_subtree") s->use_subtree = value != 0.0;
    else if (k == "subtree_maxf") s->subtree_maxf = std::max(1, (int)value);
    else if (k == "subtree_budget") s->subtree_budget = std::max(0, (int)value);
    else if (k == "diag_variant") s->diag_variant = (int)value;
    else if (k == "fuse_chain") s->fuse_chain = value != 0.0;
    else if (k == "fuse_diag") s->fuse_diag = value != 0.0;
    else if (k == "top_max_nodes") s->top_max_nodes = std::max(1, (int)value);
    else if (k == "fused_maxf") s->fused_maxf = std::max(0, std::min((int)value, B200_FUSED_MAXF));
    else if (k == "force_no_matching") s->force_no_matching = value != 0.0;
    else if (k == "device") {
        // re-home the handle: the stream and the timing events belong to a device
        int ndev = 0, dev = (int)value;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || dev < 0 || dev >= ndev) return B200_ERROR_NOT_AVAILABLE;
        if (dev != s->device) {
            cudaSetDevice(s->device);
            for (int i = 0; i < 8; i++)
                if (s->ev[i]) cudaEventDestroy(s->ev[i]), s->ev[i] = nullptr;
            destroy_streams(s);
            s->device = dev;
            if (cudaSetDevice(dev) != cudaSuccess || !create_streams(s)) return B200_ERROR_NOT_AVAILABLE;
            for (int i = 0; i < 8; i++)
                if (cudaEventCreate(&s->ev[i]) != cudaSuccess) return B200_ERROR_NOT_AVAILABLE;
        }
    }
    else return B200_ERROR_NOT_AVAILABLE;
    return 0;
}

int32_t solver_b200_initialize(struct InterfaceB200* s, int32_t ordering, int32_t matching, int32_t pivoting,
                               double pivot_epsilon, int32_t refinement_nstep, double hybrid_memory_factor,
                               int32_t verbose, int32_t general_symmetric, int32_t positive_definite, int32_t ndim,
                               const int32_t* row_pointers, const int32_t* col_indices, const double* values) {
    (void)pivoting;
    (void)hybrid_memory_factor;
    if (!s || !row_pointers || !col_indices || !values) return B200_ERROR_NULL_POINTER;
    if (s->initialized) return B200_ERROR_ALREADY_INITIALIZED;
    s->verbose = verbose;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    if (pivot_epsilon > 0.0) s->pivot_eps = pivot_epsilon;
    if (refinement_nstep >= 0) s->nrefine = refinement_nstep;
    s->opt_panel_width = std::max(4, std::min(s->opt_panel_width, B200_MAXP));

    AnalyzeOptions opt;
    opt.panel_width = s->opt_panel_width;
    opt.nd_leaf = s->opt_nd_leaf;
    if (s->relax_small >= 0) opt.relax_small = s->relax_small;
    if (s->relax_z1 >= 0.0) opt.relax_z1 = s->relax_z1;
    if (s->relax_z2 >= 0.0) opt.relax_z2 = s->relax_z2;
    if (s->relax_z3 >= 0.0) opt.relax_z3 = s->relax_z3;
    opt.verbose = verbose;
    opt.st_enable = s->use_subtree;
    opt.st_maxf = std::max(1, std::min(s->subtree_maxf, 200));
    opt.st_budget = std::max(16, s->subtree_budget > 0 ? s->subtree_budget : 8192);
    opt.st_maxcols = B200_SUBW_XS;
    if (ordering == B200_ORDERING_NONE) opt.ordering = ORDERING_NATURAL;
    else if (ordering == B200_ORDERING_AMD) opt.ordering = ORDERING_MINDEG;
    else opt.ordering = ORDERING_ND;
    // Matching::None is upgraded to "auto": unlike cuDSS (solver_cudss.rs:664-671) we must not lose accuracy on
    // zero diagonals; a symmetric positive-definite promise skips it.
    if (positive_definite || s->force_no_matching) opt.matching = 0;
    else if (matching == B200_MATCHING_NONE || matching == B200_MATCHING_AUTO) opt.matching = 2;
    else opt.matching = 1;

    const auto t_host0 = std::chrono::steady_clock::now();
    const bool want_sym = general_symmetric != 0 || positive_definite != 0;
    int rc = 0;
    {
        const char* pc = getenv("B200_PLAN_CACHE");
        const bool cache_on = !(pc && atoi(pc) == 0) && ndim >= 4096 && row_pointers[0] == 0 && row_pointers[ndim] > 0;
        uint64_t h1 = 0, h2 = 0;
        std::shared_ptr<const Plan> hit;
        s->plan_cache_hit = 0;
        if (cache_on && !values_need_matching(opt.matching, ndim, row_pointers, col_indices, values)) {
            hash_pattern(ndim, row_pointers, col_indices, h1, h2);
            std::lock_guard<std::mutex> lock(g_plan_mu);
            for (const PlanCacheEntry& e : g_plan_cache)
                if (e.h1 == h1 && e.h2 == h2 && e.n == ndim && e.nnz == row_pointers[ndim] && e.sym_lower == want_sym && same_options(e.opt, opt)) hit = e.plan;
        }
        if (hit) {
            s->plan_sp = std::const_pointer_cast<Plan>(hit); // no copy: a shared plan is never written again
            s->plan_shared = true;
            s->plan_cache_hit = 1;
            if (verbose) fprintf(stderr, "solver_b200_initialize:   plan served from the cache (identical pattern analysed before)\n");
        } else {
            s->plan_sp = std::make_shared<Plan>();
            s->plan_shared = false;
            rc = analyze(ndim, row_pointers, col_indices, values, want_sym, opt, *s->plan_sp);
            if (rc == 0 && cache_on && !s->plan_sp->matched && s->plan_sp->rscale.empty()) {
                if (h1 == 0 && h2 == 0) hash_pattern(ndim, row_pointers, col_indices, h1, h2);
                s->plan_shared = true; // the cache and this handle hold the same object (copying ~150 MB at 1M dof cost tens of ms)
                std::lock_guard<std::mutex> lock(g_plan_mu);
                if (g_plan_cache.size() >= 3) g_plan_cache.erase(g_plan_cache.begin());
                g_plan_cache.push_back({h1, h2, ndim, row_pointers[ndim], want_sym, opt, s->plan_sp});
            }
        }
    }
    if (rc == -1) return B200_ERROR_SINGULAR;
    if (rc != 0) return B200_ERROR_ANALYSIS + 2;
    if (verbose) fprintf(stderr, "solver_b200_initialize:   analysis done at %.3f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host0).count());
    Plan& P = (*s->plan_sp);
    s->n = P.n;
    s->nnz_in = P.nnz_in;
    s->sym_lower = P.sym_lower;
    s->fnnz = (int)P.full_col.size();
    s->effective_matching = P.matched || !P.rscale.empty() ? B200_MATCHING_MAX_DIAG_PRODUCT : B200_MATCHING_NONE;

    // node descriptors
    std::vector<NodeDev> nodes(P.nnodes);
    for (int v = 0; v < P.nnodes; v++) {
        NodeDev& d = nodes[v];
        d.p = P.p[v], d.u = P.u[v], d.c0 = P.c0[v];
        d.nchild = P.child_ptr[v + 1] - P.child_ptr[v];
        d.child_ptr = P.child_ptr[v];
        d.Loff = P.Loff[v], d.Uoff = P.Uoff[v], d.Coff = P.Coff[v], d.Doff = P.Doff[v], d.rows_ptr = P.rows_ptr[v];
        d.pad = P.parent[v]; // parent front (used by the persistent backward sweep)
    }
    // solve phase, bottom of the tree: the subtrees found (and laid out contiguously) by the analysis, largest first.
    // Per subtree: a descriptor, the 16-bit target index of every update row of every front (its place in the CTA's
    // shared-memory solution segment), and the (p, u) pairs of its fronts.
    std::vector<SubtreeDev> subtrees;
    std::vector<unsigned short> st_tgt;
    std::vector<uchar2> st_pu;
    s->in_sub = P.in_sub;
    {
        std::vector<std::pair<int64_t, int>> order;
        for (size_t i = 0; i < P.st_first.size(); i++) {
            int64_t ent = 0;
            for (int w = P.st_first[i]; w <= P.st_root[i]; w++) ent += (int64_t)P.p[w] * (P.p[w] + P.u[w]);
            order.push_back({-ent, (int)i});
        }
        std::sort(order.begin(), order.end()); // the long subtrees start early
        for (auto& o : order) {
            const int first = P.st_first[o.second], root = P.st_root[o.second];
            SubtreeDev d;
            d.Lbeg = P.Loff[first], d.Ubeg = P.Uoff[first], d.Dbeg = P.Doff[first];
            d.root_rows = P.rows_ptr[root];
            d.cbeg = P.c0[first], d.ncols = P.c0[root] + P.p[root] - P.c0[first], d.next = P.u[root];
            d.nfr = root - first + 1;
            d.tgt_beg = (long long)st_tgt.size(), d.pu_beg = (int)st_pu.size();
            int64_t lc = 0, uc = 0, dc = 0;
            const int cend = d.cbeg + d.ncols;
            const int* rroot = &P.rows[P.rows_ptr[root]];
            for (int w = first; w <= root; w++) {
                const int p = P.p[w], u = P.u[w];
                lc += (((int64_t)(p + u) * p + 3) & ~3), uc += (((int64_t)u * p + 3) & ~3), dc += (((int64_t)p * p + 3) & ~3);
                st_pu.push_back(make_uchar2((unsigned char)p, (unsigned char)u));
                const int* r = &P.rows[P.rows_ptr[w]];
                for (int i = 0; i < u; i++) {
                    int t;
                    if (r[i] < cend) t = r[i] - d.cbeg; // a column of this subtree
                    else {                               // leaves the subtree: a row of the root's update set
                        const int* it = std::lower_bound(rroot, rroot + d.next, r[i]);
                        if (it == rroot + d.next || *it != r[i]) return B200_ERROR_ANALYSIS + 2;
                        t = d.ncols + (int)(it - rroot);
                    }
                    st_tgt.push_back((unsigned short)t);
                }
            }
            while (st_tgt.size() & 7) st_tgt.push_back(0);
            while (st_pu.size() & 7) st_pu.push_back(make_uchar2(0, 0));
            d.Lcount = (int)lc, d.Ucount = (int)uc, d.Dcount = (int)dc;
            d.tgt_count = (int)(st_tgt.size() - (size_t)d.tgt_beg);
            d.pan = (int)std::max(lc, uc + dc);
            subtrees.push_back(d);
        }
    }
    s->n_subtrees = (int)subtrees.size();
    if (verbose) fprintf(stderr, "solver_b200_initialize:   subtree descriptors built at %.3f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host0).count());
    std::vector<AsmItem> asm_items;
    std::vector<PanelItem> panel_items;
    std::vector<SchurItem> schur_items;
    std::vector<int> fact_nodes, solve_nodes, asm_ranges, big_ranges;
    std::vector<OzakiSplitItem> oz_split;
    std::vector<OzakiItem> oz_items;
    build_work_lists(s, asm_items, panel_items, schur_items, fact_nodes, solve_nodes, asm_ranges, oz_split, oz_items);
    // big solve class: row slices of the update set; a node with no update rows still needs one (head-only) item
    std::vector<SolveItem> big_items;
    std::vector<int> big_slot;
    int nslots = 0;
    s->lv.big_ptr.assign(P.nlevels + 1, 0);
    for (int l = 0; l < P.nlevels; l++) {
        int a = s->lv.solve_ptr[(size_t)l * NSC + NSC - 1], b = s->lv.solve_ptr[(size_t)l * NSC + NSC];
        for (int e = a; e < b; e++) {
            const int v = solve_nodes[e];
            const int u = P.u[v];
            const int nsl = std::max(1, (u + B200_SLICE - 1) / B200_SLICE);
            for (int sl = 0; sl < nsl; sl++) {
                int r0 = sl * B200_SLICE;
                const int nrows = std::max(0, std::min(B200_SLICE, u - r0));
                big_items.push_back({v, r0, nrows, sl, (int)big_ranges.size(), 0});
                big_slot.push_back(nslots);
                for (int c = P.child_ptr[v]; c < P.child_ptr[v + 1]; c++) {
                    const int ch = P.child_idx[c];
                    const int* rel = &P.rel[P.rows_ptr[ch]];
                    const int pv = P.p[v];
                    big_ranges.push_back((int)(std::lower_bound(rel, rel + P.u[ch], pv) - rel));
                    big_ranges.push_back((int)(std::lower_bound(rel, rel + P.u[ch], pv + r0) - rel));
                    big_ranges.push_back((int)(std::lower_bound(rel, rel + P.u[ch], pv + r0 + nrows) - rel));
                }
            }
            nslots += nsl;
        }
        s->lv.big_ptr[l + 1] = (int)big_items.size();
    }

    if (verbose) fprintf(stderr, "solver_b200_initialize:   work lists built at %.3f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host0).count());
    // SpMV row blocks (rows never split; at most B200_SPMV_NNZ nonzeros and 1024 rows per block)
    std::vector<int> rowblk;
    {
        rowblk.push_back(0);
        int r = 0;
        const std::vector<int>& ptr = P.full_ptr;
        while (r < P.n) {
            int r1 = r + 1;
            while (r1 < P.n && r1 - r < 1024 && ptr[r1 + 1] - ptr[r] <= B200_SPMV_NNZ) r1++;
            rowblk.push_back(r1);
            r = r1;
        }
    }
    s->n_rowblk = (int)rowblk.size() - 1;

    // persistent top-of-tree sweep: all levels above the last "wide" level (more than 96 fronts)
    std::vector<SolveItem> top_items;
    std::vector<int> top_ranges, top_slot, cdone_init(P.nnodes, 1 << 30), node_slot(P.nnodes, -1);
    s->ltop = P.nlevels;
    if (s->use_top) {
        std::vector<int> cnt(P.nlevels, 0); // fronts per level outside the subtree region
        for (int v = 0; v < P.nnodes; v++)
            if (!s->in_sub[v]) cnt[P.level[v]]++;
        int l = P.nlevels;
        while (l > 0 && cnt[l - 1] <= s->top_max_nodes) l--;
        if (P.nlevels - l >= 4) s->ltop = l; // worth it only when a real chain of levels is replaced
    }
    for (int l = s->ltop; l < P.nlevels; l++)
        for (int e = P.level_ptr[l]; e < P.level_ptr[l + 1]; e++) {
            const int v = P.level_nodes[e];
            if (s->in_sub[v]) continue; // done by the subtree kernel before the persistent sweep starts (counter pre-set)
            const int u = P.u[v], pv = P.p[v];
            const int nsl = std::max(1, (u + B200_SLICE - 1) / B200_SLICE);
            cdone_init[v] = 0;
            node_slot[v] = nslots;
            for (int sl = 0; sl < nsl; sl++) {
                const int r0 = sl * B200_SLICE;
                const int nrows = std::max(0, std::min(B200_SLICE, u - r0));
                top_items.push_back({v, r0, nrows, sl, (int)top_ranges.size(), 0});
                top_slot.push_back(nslots);
                for (int c = P.child_ptr[v]; c < P.child_ptr[v + 1]; c++) {
                    const int ch = P.child_idx[c];
                    const int* rel = &P.rel[P.rows_ptr[ch]];
                    // B200_TOP_REC ints per (item, child): see sweep_top.cuh
                    top_ranges.push_back((int)(std::lower_bound(rel, rel + P.u[ch], pv) - rel));
                    top_ranges.push_back((int)(std::lower_bound(rel, rel + P.u[ch], pv + r0) - rel));
                    top_ranges.push_back((int)(std::lower_bound(rel, rel + P.u[ch], pv + r0 + nrows) - rel));
                    top_ranges.push_back(ch);
                    top_ranges.push_back((int)(uint32_t)((uint64_t)P.rows_ptr[ch] & 0xffffffffu));
                    top_ranges.push_back((int)(uint32_t)((uint64_t)P.rows_ptr[ch] >> 32));
                    top_ranges.push_back(std::max(1, (P.u[ch] + B200_SLICE - 1) / B200_SLICE));
                    top_ranges.push_back(0);
                }
            }
            nslots += nsl;
        }
    s->n_top_items = (int)top_items.size();


    if (verbose) fprintf(stderr, "solver_b200_initialize:   top items built at %.3f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host0).count());
    g_up_malloc_s = g_up_copy_s = g_up_bytes = 0.0;
    SlabScope slab_scope(&s->slabs);
    {   // one chunk for everything but the big arenas: plan arrays (~28 B per matrix entry, ~26 B per front row) + work vectors
        const double est = 36.0 * (double)s->fnnz + 28.0 * (double)P.rows_ptr[P.nnodes] + 96.0 * (double)P.n + 256.0 * (double)P.nnodes +
                           16.0 * (double)(asm_items.size() + panel_items.size() + schur_items.size() + top_items.size() + big_items.size()) +
                           4.0 * (double)(asm_ranges.size() + big_ranges.size() + top_ranges.size());
        s->slabs.chunk_default = (size_t)std::min(std::max(est * 1.1, 1048576.0), 1024.0 * 1048576.0);
    }
#define UP(dst, vec) CUDA_TRY(upload(&s->dst, vec), B200_ERROR_CUDA_MALLOC)
    UP(d_nodes, nodes);
    UP(d_rows, P.rows);
    UP(d_rel, P.rel);
    UP(d_child_idx, P.child_idx);
    UP(d_fact_nodes, fact_nodes);
    UP(d_solve_nodes, solve_nodes);
    {
        // fronts that need explicit pivot-block inverses (everything outside the subtree region), by pivot-count class;
        // within a class first the fronts below the split level (inverted on the early branch), then the rest
        s->inv_split_level = -1;
        s->pack_early = false;
        if (s->inv_overlap) {
            std::vector<int> nbig(P.nlevels, 0); // fronts per level outside the subtree region
            int sub_top = -1;
            for (int v = 0; v < P.nnodes; v++) {
                if (s->in_sub[v]) sub_top = std::max(sub_top, P.level[v]);
                else nbig[P.level[v]]++;
            }
            int l = P.nlevels;
            while (l > 0 && nbig[l - 1] <= 16) l--;
            if (l >= 1 && P.nlevels - l >= 8) s->inv_split_level = l, s->pack_early = sub_top < l;
        }
        std::vector<int> inv_nodes;
        s->lv.inv_ptr.assign(NIC + 1, 0);
        s->lv.inv_mid.assign(NIC, 0);
        for (int c = 0; c < NIC; c++) {
            for (int pass = 0; pass < 2; pass++) {
                if (pass == 1) s->lv.inv_mid[c] = (int)inv_nodes.size();
                for (int v = 0; v < P.nnodes; v++) {
                    int cls = 0;
                    while (cls < NIC - 1 && P.p[v] > IC_MAXP[cls]) cls++;
                    if (s->in_sub[v]) continue; // the subtree kernels substitute with L11 / U11 directly (their slots hold the packed pivot blocks)
                    const bool is_early = s->inv_split_level > 0 && P.level[v] < s->inv_split_level;
                    if (cls == c && is_early == (pass == 0)) inv_nodes.push_back(v);
                }
            }
            s->lv.inv_ptr[c + 1] = (int)inv_nodes.size();
        }
        UP(d_inv_nodes, inv_nodes);
        // the fronts skipped above, by class: their inverses are only computed on demand (solver_b200_debug_copy_factors)
        std::vector<int> skip_nodes;
        s->inv_skip_ptr.assign(NIC + 1, 0);
        for (int c = 0; c < NIC; c++) {
            for (int v = 0; v < P.nnodes; v++) {
                int cls = 0;
                while (cls < NIC - 1 && P.p[v] > IC_MAXP[cls]) cls++;
                if (cls == c && s->in_sub[v]) skip_nodes.push_back(v);
            }
            s->inv_skip_ptr[c + 1] = (int)skip_nodes.size();
        }
        UP(d_inv_skip, skip_nodes);
    }
    UP(d_asm, asm_items);
    UP(d_panel, panel_items);
    UP(d_schur, schur_items);
    UP(d_oz_split, oz_split);
    UP(d_oz_items, oz_items);
    UP(d_big_items, big_items);
    UP(d_big_slot, big_slot);
    UP(d_asm_ranges, asm_ranges);
    UP(d_big_ranges, big_ranges);
    UP(d_subtrees, subtrees);
    UP(d_st_tgt, st_tgt);
    UP(d_st_pu, st_pu);
    UP(d_top_items, top_items);
    UP(d_top_ranges, top_ranges);
    UP(d_top_slot, top_slot);
    UP(d_node_slot, node_slot);
    UP(d_cdone, cdone_init);
    s->cdone_init = cdone_init;
    s->n_slots = nslots;
    UP(d_a_src, P.a_src);
    static_assert(sizeof(long long) == sizeof(int64_t), "a_dst is uploaded as it is");
    CUDA_TRY(upload(reinterpret_cast<int64_t**>(&s->d_a_dst), P.a_dst), B200_ERROR_CUDA_MALLOC);
    if (!P.a_scl.empty()) UP(d_a_scl, P.a_scl);
    UP(d_rowperm, P.rowperm);
    UP(d_colperm, P.colperm);
    if (!P.rscale.empty()) {
        UP(d_rscale, P.rscale);
        UP(d_cscale, P.cscale);
    }
    UP(d_full_ptr, P.full_ptr);
    UP(d_full_col, P.full_col);
    if (!P.full_src.empty()) UP(d_full_src, P.full_src);
    UP(d_rowblk, rowblk);
#undef UP
    if (verbose)
        fprintf(stderr, "solver_b200_initialize:   plan uploaded at %.3f s (%.1f MB: cudaMalloc %.3f s, cudaMemcpy %.3f s)\n",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host0).count(), g_up_bytes * 1e-6, g_up_malloc_s, g_up_copy_s);
#define DM(ptr, count, type) CUDA_TRY(device_alloc((void**)&s->ptr, std::max<size_t>((size_t)(count), 1) * sizeof(type)), B200_ERROR_CUDA_MALLOC)
    DM(d_vals, P.nnz_in, double);
    if (P.sym_lower) DM(d_fullvals, s->fnnz, double);
    DM(d_fac, P.fac_size, double);
    DM(d_cb, P.cb_size, double);
    DM(d_dinv, P.dinv_size, double);
    if (!oz_items.empty()) {
        DM(d_oz_tiles, s->lv.oz_tile_bytes, signed char);
        DM(d_oz_scales, s->lv.oz_scales, double);
        CUDA_TRY(cudaFuncSetAttribute(k_schur_ozaki, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM), B200_ERROR_NOT_AVAILABLE);
    }
    DM(d_upiv, P.n, double);
    DM(d_lperm, P.n, int);
    DM(d_counters, 4, int);
    DM(d_amax, 1, unsigned long long);
    DM(d_b, P.n, double);
    DM(d_x, P.n, double);
    DM(d_r, P.n, double);
    DM(d_y, P.n, double);
    DM(d_z, P.n, double);
    DM(d_xp, P.n, double);
    DM(d_wv, P.rows_ptr[P.nnodes] + 1, double);
    DM(d_partial, 3 * (size_t)s->n_rowblk, double);
    DM(d_big_scratch, (size_t)std::max(nslots, 1) * B200_MAXP, double);
    DM(d_big_tickets, std::max(nslots, 1), int);
    CUDA_TRY(cudaMemset(s->d_big_tickets, 0, (size_t)std::max(nslots, 1) * sizeof(int)), B200_ERROR_CUDA_MALLOC);
    DM(d_norms, 4, double);
    DM(d_xdone, P.nnodes, int);
    DM(d_bdone, P.nnodes, int);
    CUDA_TRY(cudaMemset(s->d_bdone, 0, (size_t)std::max(P.nnodes, 1) * sizeof(int)), B200_ERROR_CUDA_MALLOC);
    DM(d_epoch, 4, int); // [0] sweep epoch, [1] / [2] item tickets of the forward / backward persistent kernels
    DM(d_abort, 1, int);
    CUDA_TRY(cudaMemset(s->d_xdone, 0, (size_t)std::max(P.nnodes, 1) * sizeof(int)), B200_ERROR_CUDA_MALLOC);
    CUDA_TRY(cudaMemset(s->d_epoch, 0, 4 * sizeof(int)), B200_ERROR_CUDA_MALLOC);
    CUDA_TRY(cudaMemset(s->d_abort, 0, sizeof(int)), B200_ERROR_CUDA_MALLOC);
#undef DM
    if (s->want_trace && s->n_top_items > 0) { // must precede the first graph capture of the sweep (the pointer is a kernel argument)
        const size_t nt = (size_t)s->n_top_items;
        CUDA_TRY(cudaMalloc((void**)&s->d_trace, 8 * nt * sizeof(unsigned long long)), B200_ERROR_CUDA_MALLOC);
        CUDA_TRY(cudaMemset(s->d_trace, 0, 8 * nt * sizeof(unsigned long long)), B200_ERROR_CUDA_MALLOC);
    }
    CUDA_TRY(cudaMallocHost((void**)&s->h_norms, 4 * sizeof(double)), B200_ERROR_MALLOC);
    CUDA_TRY(cudaMallocHost((void**)&s->h_counters, 16 * sizeof(int)), B200_ERROR_MALLOC);

    if (verbose)
        fprintf(stderr, "solver_b200_initialize:   arenas allocated at %.3f s (%d cudaMalloc calls for this handle)\n",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host0).count(), s->slabs.n_malloc);
    // kernels that need more than 48 KB of dynamic shared memory
    const int W = s->opt_panel_width;
    CUDA_TRY(cudaFuncSetAttribute(k_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_diag(B200_MAXP)), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaFuncSetAttribute(k_invert_col, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_invert(B200_MAXP)), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaFuncSetAttribute(k_front_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fused(B200_FUSED_MAXF, B200_MAXP)), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaFuncSetAttribute(k_front_fused_w8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fused(64, B200_MAXP)), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaFuncSetAttribute(k_front_warp<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(B200_FW_WARPS * (smem_fused(32, 0) + 32 * 8 + 16))), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaFuncSetAttribute(k_front_warp<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(B200_FW_WARPS * (smem_fused(64, 0) + 64 * 8 + 16))), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaFuncSetAttribute(k_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_panel(B200_MAXP)), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaFuncSetAttribute(k_panel_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B200_PW_SMEM), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaFuncSetAttribute(k_schur_fma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_schur_fma(B200_MAXP)), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaFuncSetAttribute(k_schur_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaFuncSetAttribute(k_schur_dmma_front, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaFuncSetAttribute(k_assemble_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, B200_ASM_SMEM_MAX), B200_ERROR_NOT_AVAILABLE);
    (void)W;
    if (s->n_top_items > 0) {
        CUDA_TRY(cudaFuncSetAttribute(k_fwd_top2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B200_TOP3_SMEM), B200_ERROR_NOT_AVAILABLE);
        CUDA_TRY(cudaFuncSetAttribute(k_bwd_top3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B200_TOP3_SMEM), B200_ERROR_NOT_AVAILABLE);
        int occ_f = 0, occ_b = 0, nsm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, k_fwd_top2, 256, B200_TOP3_SMEM), B200_ERROR_NOT_AVAILABLE);
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, k_bwd_top3, 256, B200_TOP3_SMEM), B200_ERROR_NOT_AVAILABLE);
        CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, s->device), B200_ERROR_NOT_AVAILABLE);
        int occ = std::min(occ_f, occ_b);
        if (occ < 1) s->n_top_items = 0; // the kernels cannot run at all: fall back to per-level launches
        s->top_grid = std::max(1, std::min(s->n_top_items, occ_f * nsm)); // (a performance choice only: items are handed out by ticket)
        s->top_grid_b = std::max(1, std::min(s->n_top_items, occ_b * nsm));
    }

    // algorithmic bytes: SpTRSV streams every stored factor entry once (+ the pivot-block inverses)
    // and touches the vectors; SpMV = 12 B per nonzero + row pointers + x and y
    s->sptrsv_bytes = 8.0 * ((double)P.nnz_L + (double)P.nnz_U) + 16.0 * P.n;
    s->spmv_bytes = 12.0 * s->fnnz + 4.0 * (P.n + 1) + 16.0 * P.n;

    // host-side plan arrays that are no longer needed (a plan that lives in the cache keeps them for the next handle)
    if (!s->plan_shared) {
        std::vector<int>().swap(P.a_src);
        std::vector<int64_t>().swap(P.a_dst);
        std::vector<double>().swap(P.a_scl);
        std::vector<int>().swap(P.full_col);
        std::vector<int>().swap(P.full_src);
        std::vector<int>().swap(P.rel);
    }

    CUDA_TRY(cudaStreamSynchronize(s->stream), B200_ERROR_CUDA_SYNCHRONIZE);
    if (verbose) {
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        printf("solver_b200_initialize: analysis done: %d fronts, %d levels, nnz(L+U)=%lld, %.3e flops, device memory used %.2f GB\n",
               P.nnodes, P.nlevels, (long long)(P.nnz_L + P.nnz_U), P.flops, (tot - fr) / 1e9);
        long long nsubfr = 0;
        for (int v = 0; v < P.nnodes; v++) nsubfr += s->in_sub[v];
        printf("solver_b200_initialize: solve phase: %d subtrees (one warp each, %lld fronts), %d persistent items above level %d\n",
               s->n_subtrees, nsubfr, s->n_top_items, s->ltop);
    }
    s->t_init_host = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host0).count();
    s->initialized = true;
    return B200_SUCCESSFUL_EXIT;
}

int32_t solver_b200_factorize_device(struct InterfaceB200* s, const double* d_values) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->initialized) return B200_ERROR_NEED_INITIALIZATION;
    if (!d_values) return B200_ERROR_NULL_POINTER;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    const Plan& P = (*s->plan_sp);
    s->factorized = false;
    s->rcond = -1.0;
    if (s->sweep_dirty) { // re-arm the dependency counters of the persistent sweep after an aborted solve
        CUDA_TRY(cudaMemcpyAsync(s->d_cdone, s->cdone_init.data(), s->cdone_init.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream), B200_ERROR_CUDA_MEMCPY);
        CUDA_TRY(cudaMemsetAsync(s->d_xdone, 0, (size_t)std::max(P.nnodes, 1) * sizeof(int), s->stream), B200_ERROR_CUDA_MEMCPY);
        CUDA_TRY(cudaMemsetAsync(s->d_bdone, 0, (size_t)std::max(P.nnodes, 1) * sizeof(int), s->stream), B200_ERROR_CUDA_MEMCPY);
        CUDA_TRY(cudaMemsetAsync(s->d_epoch, 0, 4 * sizeof(int), s->stream), B200_ERROR_CUDA_MEMCPY);
        CUDA_TRY(cudaMemsetAsync(s->d_abort, 0, sizeof(int), s->stream), B200_ERROR_CUDA_MEMCPY);
        CUDA_TRY(cudaMemsetAsync(s->d_big_tickets, 0, (size_t)std::max(s->n_slots, 1) * sizeof(int), s->stream), B200_ERROR_CUDA_MEMCPY);
        CUDA_TRY(cudaStreamSynchronize(s->stream), B200_ERROR_CUDA_SYNCHRONIZE);
        s->sweep_dirty = false;
    }
    if (d_values != s->d_vals)
        CUDA_TRY(cudaMemcpyAsync(s->d_vals, d_values, (size_t)s->nnz_in * sizeof(double), cudaMemcpyDeviceToDevice, s->stream), B200_ERROR_CUDA_MEMCPY);
    cudaEventRecord(s->ev[0], s->stream);
    if (s->fac_cleared) // the host entry point cleared the arena on the side stream, under its H2D copy: join
        CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_clr1, 0), B200_ERROR_NUM_FACTORIZATION + 1);
    else
        CUDA_TRY(cudaMemsetAsync(s->d_fac, 0, (size_t)P.fac_size * sizeof(double), s->stream), B200_ERROR_NUM_FACTORIZATION + 1);
    s->fac_cleared = false;
    CUDA_TRY(cudaMemsetAsync(s->d_counters, 0, 4 * sizeof(int), s->stream), B200_ERROR_NUM_FACTORIZATION + 1);
    CUDA_TRY(cudaMemsetAsync(s->d_amax, 0, sizeof(unsigned long long), s->stream), B200_ERROR_NUM_FACTORIZATION + 1);
    k_scatter_values<<<grid_for(s->fnnz), 256, 0, s->stream>>>(s->fnnz, s->d_a_src, s->d_a_dst, s->d_a_scl, s->d_vals, s->d_fac, s->d_amax);
    int extra = 1;
    if (s->sym_lower) {
        k_gather<<<grid_for(s->fnnz), 256, 0, s->stream>>>(s->fnnz, s->d_full_src, s->d_vals, s->d_fullvals);
        extra++;
    }
    int launches = 0;
    int rc = run_maybe_graph(s, &s->g_fact, enqueue_levels, &launches);
    if (launches > 0) s->launches_factorize = launches + extra;
    cudaEventRecord(s->ev[1], s->stream);
    if (rc != 0) return B200_ERROR_NUM_FACTORIZATION + 1;
    CUDA_TRY(cudaMemcpyAsync(s->h_counters, s->d_counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaMemcpyAsync(s->h_counters + 4, s->d_amax, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream), B200_ERROR_CUDA_MEMCPY);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) {
        if (s->verbose) fprintf(stderr, "solver_b200_factorize: %s\n", cudaGetErrorString(e));
        return B200_ERROR_CUDA_SYNCHRONIZE;
    }
    cudaEventElapsedTime(&s->ms_factorize, s->ev[0], s->ev[1]);
    unsigned long long bits;
    memcpy(&bits, s->h_counters + 4, sizeof(bits));
    double amax;
    memcpy(&amax, &bits, sizeof(double));
    if (!(amax <= 1.79e308)) return B200_ERROR_NUM_FACTORIZATION + 2; // NaN / Inf in the input values
    s->n_perturbed = s->h_counters[0];
    if (s->h_counters[2] != 0) return B200_ERROR_SINGULAR;
    s->spmv_vals = s->sym_lower ? s->d_fullvals : s->d_vals;
    s->factorized = true;
    return B200_SUCCESSFUL_EXIT;
}

// host entry points: the factor arena (hundreds of MB) is cleared on the side stream WHILE the copy engine brings the new
// values over PCIe on the main stream; the main stream then waits for the clear before the numeric kernels start
static void clear_fac_under_h2d(InterfaceB200* s) {
    if (!s->side || !s->ev_clr0 || !s->ev_clr1) return;
    if (cudaEventRecord(s->ev_clr0, s->stream) != cudaSuccess) return; // (orders the clear after the previous solve's reads)
    cudaStreamWaitEvent(s->side, s->ev_clr0, 0);
    if (cudaMemsetAsync(s->d_fac, 0, (size_t)(*s->plan_sp).fac_size * sizeof(double), s->side) != cudaSuccess) return;
    cudaEventRecord(s->ev_clr1, s->side);
    s->fac_cleared = true;
}

// ---- host <-> device transfers of the callers' buffers ------------------------------------------------------------
namespace {
bool host_pointer_is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
bool staging_ready(InterfaceB200* s) {
    if (s->h_stage) return true;
    if (cudaMallocHost((void**)&s->h_stage, 2 * InterfaceB200::NST * InterfaceB200::STAGE_PIECE) != cudaSuccess) {
        cudaGetLastError();
        s->h_stage = nullptr;
        return false;
    }
    for (int i = 0; i < 2 * InterfaceB200::NST; i++)
        if (cudaEventCreateWithFlags(&s->ev_stage[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            cudaFreeHost(s->h_stage), s->h_stage = nullptr;
            s->staged_copy = 0; // plain copies from now on
            return false;
        }
    return true;
}
// one stripe [lo, hi) of a transfer between pageable host memory and the device, through the stripe's two staging buffers
cudaError_t stripe_copy(InterfaceB200* s, int t, char* dev, char* host, size_t lo, size_t hi, bool to_device) {
    cudaError_t e = cudaSetDevice(s->device);
    const size_t piece = InterfaceB200::STAGE_PIECE;
    char* buf[2] = {s->h_stage + (size_t)(2 * t) * piece, s->h_stage + (size_t)(2 * t + 1) * piece};
    cudaEvent_t ev[2] = {s->ev_stage[2 * t], s->ev_stage[2 * t + 1]};
    size_t prev_off = 0, prev_n = 0;
    int k = 0;
    for (size_t off = lo; off < hi && e == cudaSuccess; off += piece, k++) {
        const size_t n = std::min(piece, hi - off);
        const int b = k & 1;
        if (to_device) {
            e = cudaEventSynchronize(ev[b]); // the DMA that last read this buffer is done (a fresh event is complete)
            if (e != cudaSuccess) break;
            memcpy(buf[b], host + off, n);
            e = cudaMemcpyAsync(dev + off, buf[b], n, cudaMemcpyHostToDevice, s->stream);
            if (e == cudaSuccess) e = cudaEventRecord(ev[b], s->stream);
        } else {
            e = cudaMemcpyAsync(buf[b], dev + off, n, cudaMemcpyDeviceToHost, s->stream);
            if (e == cudaSuccess) e = cudaEventRecord(ev[b], s->stream);
            if (e == cudaSuccess && k > 0) { // unload the previous piece while this one is in flight
                e = cudaEventSynchronize(ev[b ^ 1]);
                if (e == cudaSuccess) memcpy(host + prev_off, buf[b ^ 1], prev_n);
            }
            prev_off = off, prev_n = n;
        }
    }
    if (!to_device && e == cudaSuccess && k > 0) {
        e = cudaEventSynchronize(ev[(k - 1) & 1]);
        if (e == cudaSuccess) memcpy(host + prev_off, buf[(k - 1) & 1], prev_n);
    }
    return e;
}
// H2D: enqueued on the handle's stream (ordered like a cudaMemcpyAsync).  D2H: pinned destination -> enqueued; pageable
// destination -> complete when the call returns.
cudaError_t transfer(InterfaceB200* s, void* dev, void* host, size_t bytes, bool to_device) {
    const bool plain = !s->staged_copy || bytes < 4 * InterfaceB200::STAGE_PIECE || host_pointer_is_pinned(host) || !staging_ready(s);
    if (plain)
        return to_device ? cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s->stream)
                         : cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s->stream);
    // stripes of at least one piece (small vectors use fewer threads), 256-byte aligned
    const int nt = (int)std::max<size_t>(1, std::min<size_t>(InterfaceB200::NST, bytes / InterfaceB200::STAGE_PIECE));
    const size_t stripe = ((bytes + nt - 1) / nt + 255) & ~(size_t)255;
    cudaError_t err[InterfaceB200::NST];
    for (int t = 0; t < InterfaceB200::NST; t++) err[t] = cudaSuccess;
    s->stripe_pool.run(nt, [&](int t) {
        const size_t lo = std::min(bytes, (size_t)t * stripe), hi = std::min(bytes, lo + stripe);
        if (lo < hi) err[t] = stripe_copy(s, t, (char*)dev, (char*)host, lo, hi, to_device);
    });
    for (int t = 0; t < nt; t++)
        if (err[t] != cudaSuccess) return err[t];
    return cudaSuccess;
}
} // namespace

int32_t solver_b200_copy_h2d(struct InterfaceB200* s, void* dst_device, const void* src_host, int64_t bytes) {
    if (!s || !dst_device || !src_host) return B200_ERROR_NULL_POINTER;
    if (bytes <= 0) return B200_SUCCESSFUL_EXIT;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(transfer(s, dst_device, const_cast<void*>(src_host), (size_t)bytes, true), B200_ERROR_CUDA_MEMCPY);
    return B200_SUCCESSFUL_EXIT;
}

int32_t solver_b200_copy_d2h(struct InterfaceB200* s, void* dst_host, const void* src_device, int64_t bytes) {
    if (!s || !dst_host || !src_device) return B200_ERROR_NULL_POINTER;
    if (bytes <= 0) return B200_SUCCESSFUL_EXIT;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(transfer(s, const_cast<void*>(src_device), dst_host, (size_t)bytes, false), B200_ERROR_CUDA_MEMCPY);
    return B200_SUCCESSFUL_EXIT;
}

int32_t solver_b200_factorize(struct InterfaceB200* s, int32_t* effective_matching, int32_t* effective_pivoting,
                              int32_t verbose, const double* values) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->initialized) return B200_ERROR_NEED_INITIALIZATION;
    if (!values) return B200_ERROR_NULL_POINTER;
    s->verbose = verbose;
    s->factorized = false; // the factor arena is about to be cleared: a failed copy must not leave a "factorized" handle
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    clear_fac_under_h2d(s);
    CUDA_TRY(transfer(s, s->d_vals, const_cast<double*>(values), (size_t)s->nnz_in * sizeof(double), true), B200_ERROR_CUDA_MEMCPY);
    int32_t rc = solver_b200_factorize_device(s, s->d_vals);
    if (effective_matching) *effective_matching = s->effective_matching;
    if (effective_pivoting) *effective_pivoting = s->effective_pivoting;
    if (rc == 0 && verbose) {
        if (s->n_perturbed > 0)
            printf("solver_b200_factorize: WARNING: %d pivot(s) perturbed (matrix may be (nearly) singular)\n", s->n_perturbed);
        printf("solver_b200_factorize: numeric factorization completed in %.3f ms (device)\n", s->ms_factorize);
    }
    return rc;
}

// ---- COO-level boundary: what CsrMatrix::update_from_coo + solver_cudss_factorize do together
// (russell_sparse/src/solver_cudss.rs:195-290), with the per-factorize conversion moved to the device.
int32_t solver_b200_initialize_coo(struct InterfaceB200* s, int32_t ordering, int32_t matching, int32_t pivoting,
                                   double pivot_epsilon, int32_t refinement_nstep, double hybrid_memory_factor,
                                   int32_t verbose, int32_t general_symmetric, int32_t positive_definite, int32_t ndim,
                                   int32_t nnz_coo, const int32_t* indices_i, const int32_t* indices_j, const double* values) {
    if (!s || !indices_i || !indices_j || !values) return B200_ERROR_NULL_POINTER;
    if (s->initialized) return B200_ERROR_ALREADY_INITIALIZED;
    if (ndim < 1 || nnz_coo < 1) return B200_ERROR_ANALYSIS + 3;
    if (general_symmetric || positive_definite)
        for (int32_t k = 0; k < nnz_coo; k++)
            if (indices_j[k] > indices_i[k]) return B200_ERROR_ANALYSIS + 4; // Sym::YesLower promised: j <= i
    std::vector<int32_t> ptr((size_t)ndim + 1), idx((size_t)nnz_coo), seg_ptr((size_t)nnz_coo + 1), seg_idx((size_t)nnz_coo);
    std::vector<double> val((size_t)nnz_coo);
    if (b200_coo_to_csr_map(ndim, ndim, nnz_coo, indices_i, indices_j, values, ptr.data(), idx.data(), val.data(),
                            seg_ptr.data(), seg_idx.data()) != 0)
        return B200_ERROR_ANALYSIS + 3;
    int32_t rc = solver_b200_initialize(s, ordering, matching, pivoting, pivot_epsilon, refinement_nstep, hybrid_memory_factor,
                                        verbose, general_symmetric, positive_definite, ndim, ptr.data(), idx.data(), val.data());
    if (rc != B200_SUCCESSFUL_EXIT) return rc;
    const int nslots = ptr[ndim];
    s->nnz_coo = nnz_coo;
    {
        SlabScope slab_scope(&s->slabs); // (one chunk for the three arrays)
        s->slabs.chunk_default = std::max<size_t>((size_t)1 << 20, ((size_t)nslots + 1) * sizeof(int) + (size_t)nnz_coo * (sizeof(int) + sizeof(double)) + 1024);
        if (!s->slabs.chunks.empty()) s->slabs.chunks.back().used = s->slabs.chunks.back().cap; // start a fresh chunk of that size
        cudaError_t e = device_alloc((void**)&s->d_seg_ptr, ((size_t)nslots + 1) * sizeof(int));
        if (e == cudaSuccess) e = device_alloc((void**)&s->d_seg_idx, (size_t)nnz_coo * sizeof(int));
        if (e == cudaSuccess) e = device_alloc((void**)&s->d_coo_vals, (size_t)nnz_coo * sizeof(double));
        if (e != cudaSuccess) return B200_ERROR_CUDA_MALLOC;
    }
    CUDA_TRY(cudaMemcpyAsync(s->d_seg_ptr, seg_ptr.data(), ((size_t)nslots + 1) * sizeof(int), cudaMemcpyHostToDevice, s->stream), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaMemcpyAsync(s->d_seg_idx, seg_idx.data(), (size_t)nnz_coo * sizeof(int), cudaMemcpyHostToDevice, s->stream), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaStreamSynchronize(s->stream), B200_ERROR_CUDA_SYNCHRONIZE);
    s->coo_guard.remember(ndim, nnz_coo, indices_i, indices_j, ptr, idx);
    return B200_SUCCESSFUL_EXIT;
}

int32_t solver_b200_factorize_coo_device(struct InterfaceB200* s, const double* d_coo_values) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->initialized || !s->d_seg_ptr) return B200_ERROR_NEED_INITIALIZATION;
    if (!d_coo_values) return B200_ERROR_NULL_POINTER;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    k_coo_to_csr_values<<<grid_for(s->nnz_in), 256, 0, s->stream>>>(s->nnz_in, s->d_seg_ptr, s->d_seg_idx, d_coo_values, s->d_vals);
    return solver_b200_factorize_device(s, s->d_vals);
}

int32_t solver_b200_factorize_coo(struct InterfaceB200* s, int32_t* effective_matching, int32_t* effective_pivoting,
                                  int32_t verbose, const double* coo_values) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->initialized || !s->d_seg_ptr) return B200_ERROR_NEED_INITIALIZATION;
    if (!coo_values) return B200_ERROR_NULL_POINTER;
    s->verbose = verbose;
    s->factorized = false;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    clear_fac_under_h2d(s);
    CUDA_TRY(transfer(s, s->d_coo_vals, const_cast<double*>(coo_values), (size_t)s->nnz_coo * sizeof(double), true), B200_ERROR_CUDA_MEMCPY);
    int32_t rc = solver_b200_factorize_coo_device(s, s->d_coo_vals);
    if (effective_matching) *effective_matching = s->effective_matching;
    if (effective_pivoting) *effective_pivoting = s->effective_pivoting;
    if (rc == 0 && verbose) printf("solver_b200_factorize_coo: numeric factorization completed in %.3f ms (device)\n", s->ms_factorize);
    return rc;
}

// what SolverCUDSS::factorize does on every call (update_from_coo over the caller's indices, solver_cudss.rs:209), at the
// price of two memcmp that run on a helper thread underneath the copy and the kernels
int32_t solver_b200_factorize_coo_checked(struct InterfaceB200* s, int32_t* effective_matching, int32_t* effective_pivoting,
                                          int32_t verbose, int32_t nnz_coo, const int32_t* indices_i, const int32_t* indices_j,
                                          const double* coo_values) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->initialized || !s->d_seg_ptr || !s->coo_guard.armed()) return B200_ERROR_NEED_INITIALIZATION;
    if (!indices_i || !indices_j || !coo_values) return B200_ERROR_NULL_POINTER;
    if (nnz_coo != s->nnz_coo) return B200_ERROR_ANALYSIS + 5;
    std::atomic<int> same{1};
    std::thread checker([&] { same.store(s->coo_guard.same(indices_i, indices_j) ? 1 : 0); });
    int32_t rc = solver_b200_factorize_coo(s, effective_matching, effective_pivoting, verbose, coo_values);
    checker.join();
    if (same.load()) return rc;
    // the triplets were refilled in another order: rebuild the slot map (or refuse a different pattern) and redo the work
    s->factorized = false;
    std::vector<int32_t> seg_ptr, seg_idx;
    if (s->coo_guard.remap(indices_i, indices_j, seg_ptr, seg_idx) != 0) return B200_ERROR_ANALYSIS + 5;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaMemcpyAsync(s->d_seg_ptr, seg_ptr.data(), ((size_t)s->nnz_in + 1) * sizeof(int), cudaMemcpyHostToDevice, s->stream), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaMemcpyAsync(s->d_seg_idx, seg_idx.data(), (size_t)nnz_coo * sizeof(int), cudaMemcpyHostToDevice, s->stream), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaStreamSynchronize(s->stream), B200_ERROR_CUDA_SYNCHRONIZE);
    if (verbose) printf("solver_b200_factorize_coo_checked: triplet order changed, slot map rebuilt\n");
    return solver_b200_factorize_coo(s, effective_matching, effective_pivoting, verbose, coo_values);
}

int32_t solver_b200_solve_device(struct InterfaceB200* s, double* d_xout, const double* d_rhs) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->factorized) return B200_ERROR_NEED_FACTORIZATION;
    if (!d_xout || !d_rhs) return B200_ERROR_NULL_POINTER;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    s->launches_solve = 0;
    cudaEventRecord(s->ev[2], s->stream);
    int rc = sweep(s, d_rhs, d_xout, 0, true);
    double prev = -1.0, prev_omega = -1.0;
    int steps = 0;
    for (int it = 0; rc == 0; it++) {
        if (residual(s, d_xout, d_rhs, s->d_r, it == 0) != 0) {
            rc = 1;
            break;
        }
        const double rr = s->h_norms[0], bb = s->h_norms[1], omega = s->h_norms[2];
        const double rel = bb > 0.0 ? std::sqrt(rr / bb) : std::sqrt(rr);
        s->last_rel_residual = rel;
        s->last_backward_error = omega;
        if (!(rel == rel)) break;                        // NaN: nothing to refine
        if (it >= s->nrefine) break;                     // out of steps
        if (rel <= s->ir_tol) break;                     // requested residual reached
        if (omega <= 4.0 * 2.220446049250313e-16) break; // componentwise backward error at machine precision:
                                                         // the residual itself is rounding noise from here on
        if (prev >= 0.0 && rel > 0.5 * prev && omega > 0.5 * prev_omega) break; // stagnation
        prev = rel;
        prev_omega = omega;
        rc = sweep(s, s->d_r, d_xout, 1, false);
        steps++;
    }
    cudaEventRecord(s->ev[3], s->stream);
    s->last_refine_steps = steps;
    if (s->n_top_items > 0) cudaMemcpyAsync(s->h_counters + 8, s->d_abort, sizeof(int), cudaMemcpyDeviceToHost, s->stream);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (rc != 0 || e != cudaSuccess) {
        if (s->verbose) fprintf(stderr, "solver_b200_solve: %s\n", cudaGetErrorString(e));
        return B200_ERROR_SOLVE + 1;
    }
    if (s->n_top_items > 0 && s->h_counters[8] != 0) { // a dependency wait timed out: reset the sweep state and report
        s->sweep_dirty = true;
        s->factorized = false;
        if (s->verbose) fprintf(stderr, "solver_b200_solve: persistent sweep aborted (dependency wait timed out)\n");
        return B200_ERROR_SOLVE + 1;
    }
    cudaEventElapsedTime(&s->ms_solve, s->ev[2], s->ev[3]);
    cudaEventElapsedTime(&s->ms_sptrsv, s->ev[4], s->ev[5]);
    cudaEventElapsedTime(&s->ms_spmv, s->ev[6], s->ev[7]);
    // accuracy gate (UMFPACK reports a failed solve through its status, solver_umfpack.rs:380-387; cuDSS does not): the
    // refinement loop above may stop on stagnation or on the step limit, so the answer is checked here.  A solve FAILED
    // when the residual is not a number, or when it stays above 10 x ir_tol although the componentwise backward error is
    // not at rounding level either (a backward-stable solve of an ill-conditioned system is not a failure); with
    // "strict_residual" the first condition alone decides.
    {
        const double rel = s->last_rel_residual, omega = s->last_backward_error;
        const double bar = 10.0 * s->ir_tol;
        const bool nan = !(rel == rel) || !(omega == omega);
        const bool unstable = rel > bar && (s->strict_residual || omega > 1e3 * 2.220446049250313e-16);
        if (nan || unstable) {
            if (s->verbose)
                fprintf(stderr, "solver_b200_solve: refinement failed: ||b-Ax||/||b|| = %.3e, backward error %.3e, %d perturbed pivot(s)\n",
                        rel, omega, s->n_perturbed);
            return B200_ERROR_SOLVE + 7;
        }
    }
    return B200_SUCCESSFUL_EXIT;
}

int32_t solver_b200_solve(struct InterfaceB200* s, double* x, const double* rhs, int32_t verbose) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->factorized) return B200_ERROR_NEED_FACTORIZATION;
    if (!x || !rhs) return B200_ERROR_NULL_POINTER;
    s->verbose = verbose;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(transfer(s, s->d_b, const_cast<double*>(rhs), (size_t)s->n * sizeof(double), true), B200_ERROR_CUDA_MEMCPY);
    int32_t rc = solver_b200_solve_device(s, s->d_x, s->d_b);
    if (rc != 0) return rc;
    CUDA_TRY(transfer(s, s->d_x, x, (size_t)s->n * sizeof(double), false), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaStreamSynchronize(s->stream), B200_ERROR_CUDA_SYNCHRONIZE);
    if (verbose)
        printf("solver_b200_solve: solution completed: %d refinement step(s), ||b-Ax||/||b|| = %.3e, %.3f ms (device)\n",
               s->last_refine_steps, s->last_rel_residual, s->ms_solve);
    return B200_SUCCESSFUL_EXIT;
}

int32_t solver_b200_residual(struct InterfaceB200* s, const double* x, const double* rhs, double* rel_residual) {
    if (!s || !x || !rhs || !rel_residual) return B200_ERROR_NULL_POINTER;
    if (!s->factorized) return B200_ERROR_NEED_FACTORIZATION;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaMemcpyAsync(s->d_b, rhs, (size_t)s->n * sizeof(double), cudaMemcpyHostToDevice, s->stream), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaMemcpyAsync(s->d_x, x, (size_t)s->n * sizeof(double), cudaMemcpyHostToDevice, s->stream), B200_ERROR_CUDA_MEMCPY);
    int rc = residual(s, s->d_x, s->d_b, s->d_r, true);
    if (rc != 0) return B200_ERROR_SOLVE + 1;
    double rr = s->h_norms[0], bb = s->h_norms[1];
    *rel_residual = bb > 0.0 ? std::sqrt(rr / bb) : std::sqrt(rr);
    cudaEventElapsedTime(&s->ms_spmv, s->ev[6], s->ev[7]);
    return B200_SUCCESSFUL_EXIT;
}

int32_t solver_b200_spmv(struct InterfaceB200* s, double* y, const double* x) {
    if (!s || !x || !y) return B200_ERROR_NULL_POINTER;
    if (!s->factorized) return B200_ERROR_NEED_FACTORIZATION;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    CUDA_TRY(cudaMemcpyAsync(s->d_x, x, (size_t)s->n * sizeof(double), cudaMemcpyHostToDevice, s->stream), B200_ERROR_CUDA_MEMCPY);
    cudaEventRecord(s->ev[6], s->stream);
    k_spmv_stream<<<s->n_rowblk, 256, 0, s->stream>>>(s->d_rowblk, s->d_full_ptr, s->d_full_col, s->spmv_vals, s->d_x, nullptr, s->d_r, s->d_partial, 0);
    cudaEventRecord(s->ev[7], s->stream);
    CUDA_TRY(cudaMemcpyAsync(y, s->d_r, (size_t)s->n * sizeof(double), cudaMemcpyDeviceToHost, s->stream), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaStreamSynchronize(s->stream), B200_ERROR_CUDA_SYNCHRONIZE);
    cudaEventElapsedTime(&s->ms_spmv, s->ev[6], s->ev[7]);
    return B200_SUCCESSFUL_EXIT;
}

int32_t solver_b200_determinant(struct InterfaceB200* s, double* coefficient, double* exponent) {
    if (!s || !coefficient || !exponent) return B200_ERROR_NULL_POINTER;
    if (!s->factorized) return B200_ERROR_NEED_FACTORIZATION;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    const Plan& P = (*s->plan_sp);
    const int n = s->n;
    std::vector<double> up(n);
    std::vector<int> lp(n);
    CUDA_TRY(cudaMemcpy(up.data(), s->d_upiv, n * sizeof(double), cudaMemcpyDeviceToHost), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaMemcpy(lp.data(), s->d_lperm, n * sizeof(int), cudaMemcpyDeviceToHost), B200_ERROR_CUDA_MEMCPY);
    // det(A) = det(Dr)^-1 det(Dc)^-1 sign(rowperm) sign(colperm) sign(local pivots) prod(U_kk)
    double mant = 1.0, ex = 0.0;
    auto mul = [&](double v) {
        if (v == 0.0) { mant = 0.0; return; }
        mant *= v;
        double a = std::fabs(mant);
        if (a != 0.0 && (a >= 1e150 || a < 1e-150)) {
            double e10 = std::floor(std::log10(a));
            mant /= std::pow(10.0, e10);
            ex += e10;
        }
    };
    for (int k = 0; k < n; k++) mul(up[k]);
    if (!P.rscale.empty())
        for (int i = 0; i < n; i++) mul(1.0 / (P.rscale[i] * P.cscale[i]));
    auto perm_sign = [&](const std::vector<int>& pm) {
        std::vector<char> seen(pm.size(), 0);
        int sign = 1;
        for (size_t i = 0; i < pm.size(); i++) {
            if (seen[i]) continue;
            size_t j = i;
            int len = 0;
            while (!seen[j]) seen[j] = 1, j = pm[j], len++;
            if ((len & 1) == 0) sign = -sign;
        }
        return sign;
    };
    int sign = perm_sign(P.rowperm) * perm_sign(P.colperm);
    for (int v = 0; v < P.nnodes; v++) { // local pivot permutations (per front)
        std::vector<int> loc(lp.begin() + P.c0[v], lp.begin() + P.c0[v] + P.p[v]);
        sign *= perm_sign(loc);
    }
    mant *= sign;
    double a = std::fabs(mant);
    if (a != 0.0) {
        double e10 = std::floor(std::log10(a));
        mant /= std::pow(10.0, e10);
        ex += e10;
    }
    *coefficient = mant;
    *exponent = ex;
    return B200_SUCCESSFUL_EXIT;
}

// standalone entry of the tcgen05 Schur kernel (tests / profiling): C (u x u, column-major, host) <- C - A * B^T with A, B
// u x k column-major host arrays -- the operation k_schur_ozaki performs for one front, through the same split kernel,
// the same tile layout and the same launch as the factorization uses.  Returns the kernel time in *ms_out.
int32_t solver_b200_ozaki_gemm(int32_t u, int32_t k, const double* a, const double* b, double* c, double* ms_out) {
    if (!a || !b || !c || u < 1 || k < 1) return B200_ERROR_NULL_POINTER;
    InterfaceB200* s = nullptr;
    const int kch = (k + OZ_KC - 1) / OZ_KC;
    const int nta = (u + OZ_BM - 1) / OZ_BM, ntb = (u + OZ_BN - 1) / OZ_BN;
    const size_t a_bytes = (size_t)nta * kch * OZ_S * OZ_A_BYTES, b_bytes = (size_t)ntb * kch * OZ_S * OZ_B_BYTES;
    double *d_ab = nullptr, *d_c = nullptr, *d_sc = nullptr;
    signed char* d_tiles = nullptr;
    OzakiSplitItem* d_split = nullptr;
    OzakiItem* d_items = nullptr;
    CUDA_TRY(cudaMalloc(&d_ab, (size_t)2 * u * k * sizeof(double)), B200_ERROR_CUDA_MALLOC);
    CUDA_TRY(cudaMalloc(&d_c, (size_t)u * u * sizeof(double)), B200_ERROR_CUDA_MALLOC);
    CUDA_TRY(cudaMalloc(&d_sc, (size_t)2 * u * sizeof(double)), B200_ERROR_CUDA_MALLOC);
    CUDA_TRY(cudaMalloc(&d_tiles, a_bytes + b_bytes), B200_ERROR_CUDA_MALLOC);
    CUDA_TRY(cudaMemcpy(d_ab, a, (size_t)u * k * sizeof(double), cudaMemcpyHostToDevice), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaMemcpy(d_ab + (size_t)u * k, b, (size_t)u * k * sizeof(double), cudaMemcpyHostToDevice), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaMemcpy(d_c, c, (size_t)u * u * sizeof(double), cudaMemcpyHostToDevice), B200_ERROR_CUDA_MEMCPY);
    std::vector<OzakiSplitItem> split = {{0, 0, 0, u, k, u, OZ_BM}, {(long long)u * k, (long long)a_bytes, u, u, k, u, OZ_BN}};
    std::vector<OzakiItem> items;
    for (int tj = 0; tj < ntb; tj++)
        for (int ti = 0; ti < nta; ti++) items.push_back({0, (long long)a_bytes, 0, u, 0, u, kch, ti, tj});
    CUDA_TRY(cudaMalloc(&d_split, split.size() * sizeof(OzakiSplitItem)), B200_ERROR_CUDA_MALLOC);
    CUDA_TRY(cudaMalloc(&d_items, items.size() * sizeof(OzakiItem)), B200_ERROR_CUDA_MALLOC);
    CUDA_TRY(cudaMemcpy(d_split, split.data(), split.size() * sizeof(OzakiSplitItem), cudaMemcpyHostToDevice), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaMemcpy(d_items, items.data(), items.size() * sizeof(OzakiItem), cudaMemcpyHostToDevice), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaFuncSetAttribute(k_schur_ozaki, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM), B200_ERROR_NOT_AVAILABLE);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_ozaki_split<<<dim3((nta * OZ_BM + 127) / 128, 2), 128>>>(d_split, d_ab, d_tiles, d_sc);
    k_schur_ozaki<<<std::min((int)items.size(), 148), 128, OZ_SMEM>>>(d_items, (int)items.size(), d_tiles, d_sc, d_c);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms;
    if (err == cudaSuccess) err = cudaMemcpy(c, d_c, (size_t)u * u * sizeof(double), cudaMemcpyDeviceToHost);
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    cudaFree(d_ab), cudaFree(d_c), cudaFree(d_sc), cudaFree(d_tiles), cudaFree(d_split), cudaFree(d_items);
    if (err != cudaSuccess) {
        fprintf(stderr, "solver_b200_ozaki_gemm: %s\n", cudaGetErrorString(err));
        return B200_ERROR_NUM_FACTORIZATION + 1;
    }
    return B200_SUCCESSFUL_EXIT;
}

// reciprocal condition number estimate with UMFPACK's definition (Info[UMFPACK_RCOND] = min|U_kk| / max|U_kk|, the number
// interface_umfpack.c:179-184 hands to StatsLinSol.output.umfpack_rcond_estimate), taken over the scaled, permuted factors
int32_t solver_b200_rcond(struct InterfaceB200* s, double* rcond) {
    if (!s || !rcond) return B200_ERROR_NULL_POINTER;
    if (!s->factorized) return B200_ERROR_NEED_FACTORIZATION;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    unsigned long long* mm = (unsigned long long*)s->d_norms; // 4 doubles of scratch
    const unsigned long long init[2] = {~0ull, 0ull};
    CUDA_TRY(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, s->stream), B200_ERROR_CUDA_MEMCPY);
    k_minmax_abs<<<grid_for(s->n), 256, 0, s->stream>>>(s->n, s->d_upiv, mm);
    unsigned long long out[2];
    CUDA_TRY(cudaMemcpyAsync(out, mm, sizeof(out), cudaMemcpyDeviceToHost, s->stream), B200_ERROR_CUDA_MEMCPY);
    CUDA_TRY(cudaStreamSynchronize(s->stream), B200_ERROR_CUDA_SYNCHRONIZE);
    double lo, hi;
    memcpy(&lo, &out[0], 8), memcpy(&hi, &out[1], 8);
    s->rcond = (hi > 0.0 && hi == hi && lo == lo) ? lo / hi : 0.0;
    *rcond = s->rcond;
    return B200_SUCCESSFUL_EXIT;
}

int32_t solver_b200_get_stats(struct InterfaceB200* s, double* out, int32_t n_out) {
    if (!s || !out) return B200_ERROR_NULL_POINTER;
    if (!s->initialized) return B200_ERROR_NEED_INITIALIZATION;
    const Plan& P = (*s->plan_sp);
    double v[B200_STAT_COUNT];
    v[B200_STAT_NNODES] = P.nnodes;
    v[B200_STAT_NLEVELS] = P.nlevels;
    v[B200_STAT_NNZ_L] = (double)P.nnz_L;
    v[B200_STAT_NNZ_U] = (double)P.nnz_U;
    v[B200_STAT_FLOPS] = P.flops;
    v[B200_STAT_FAC_BYTES] = 8.0 * P.fac_size;
    v[B200_STAT_CB_BYTES] = 8.0 * P.cb_size;
    v[B200_STAT_MAX_FRONT] = P.max_front;
    v[B200_STAT_T_ORDER_S] = P.t_order;
    v[B200_STAT_T_SYMBOLIC_S] = P.t_symbolic;
    v[B200_STAT_N_PERTURBED] = s->n_perturbed;
    v[B200_STAT_LAST_REL_RESIDUAL] = s->last_rel_residual;
    v[B200_STAT_LAST_REFINE_STEPS] = s->last_refine_steps;
    v[B200_STAT_MS_FACTORIZE_DEVICE] = s->ms_factorize;
    v[B200_STAT_MS_SOLVE_DEVICE] = s->ms_solve;
    v[B200_STAT_MS_SPTRSV_DEVICE] = s->ms_sptrsv;
    v[B200_STAT_MS_SPMV_DEVICE] = s->ms_spmv;
    v[B200_STAT_LAUNCHES_FACTORIZE] = s->launches_factorize;
    v[B200_STAT_LAUNCHES_SOLVE] = s->launches_solve;
    v[B200_STAT_SPTRSV_BYTES] = s->sptrsv_bytes;
    v[B200_STAT_SPMV_BYTES] = s->spmv_bytes;
    v[B200_STAT_MATCHED] = P.matched ? 1.0 : 0.0;
    v[B200_STAT_T_MATCH_S] = P.t_match;
    v[B200_STAT_LAST_BACKWARD_ERROR] = s->last_backward_error;
    v[B200_STAT_EFFECTIVE_ORDERING] = P.opt.ordering == ORDERING_NATURAL ? B200_ORDERING_NONE : P.opt.ordering == ORDERING_MINDEG ? B200_ORDERING_AMD : B200_ORDERING_ND;
    v[B200_STAT_EFFECTIVE_SCALING] = P.rscale.empty() ? 0.0 : 1.0;
    v[B200_STAT_RCOND] = s->rcond;
    v[B200_STAT_T_INITIALIZE_HOST_S] = s->t_init_host;
    v[B200_STAT_PLAN_CACHE_HIT] = s->plan_cache_hit;
    for (int i = 0; i < n_out && i < B200_STAT_COUNT; i++) out[i] = v[i];
    return B200_SUCCESSFUL_EXIT;
}

// debug/profiling: per-item timestamps (ns, %globaltimer) of the last persistent sweeps, 4 per item (start, dependency
// satisfied, end, unused), forward items first then backward; item descriptors (node, level, slice, nrows) in `desc`.
// Needs set_option("trace", 1) before initialize.  Returns the number of items (<= cap) or a negative value.
int32_t solver_b200_debug_trace(struct InterfaceB200* s, unsigned long long* out, int32_t* desc, int32_t cap) {
    if (!s || !s->d_trace || !s->initialized) return -1;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    const int n = std::min(cap, s->n_top_items);
    if (out) {
        cudaMemcpy(out, s->d_trace, 4 * (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        cudaMemcpy(out + 4 * (size_t)n, s->d_trace + 4 * (size_t)s->n_top_items, 4 * (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    }
    if (desc) {
        std::vector<SolveItem> items(s->n_top_items);
        cudaMemcpy(items.data(), s->d_top_items, items.size() * sizeof(SolveItem), cudaMemcpyDeviceToHost);
        for (int i = 0; i < n; i++) {
            desc[4 * i] = items[i].node, desc[4 * i + 1] = (*s->plan_sp).level[items[i].node];
            desc[4 * i + 2] = items[i].slice, desc[4 * i + 3] = items[i].nrows;
        }
    }
    return s->n_top_items;
}

int32_t solver_b200_debug_copy_factors(struct InterfaceB200* s, double* fac, int64_t fac_len, double* dinv,
                                       int64_t dinv_len, int32_t* lperm, int64_t n) {
    if (!s) return B200_ERROR_NULL_POINTER;
    if (!s->factorized) return B200_ERROR_NEED_FACTORIZATION;
    CUDA_TRY(cudaSetDevice(s->device), B200_ERROR_NOT_AVAILABLE);
    if (dinv && !s->inv_skip_ptr.empty() && s->inv_skip_ptr[NIC] > 0) { // inverses the solve phase never needed: compute them now
        for (int c = 0; c < NIC; c++) {
            const int nn = s->inv_skip_ptr[c + 1] - s->inv_skip_ptr[c];
            if (nn > 0)
                k_invert_col<<<nn, 2 * IC_MAXP[c], smem_invert(IC_MAXP[c]), s->stream>>>(s->d_inv_skip + s->inv_skip_ptr[c], s->d_nodes, s->d_fac,
                                                                                          s->d_dinv, IC_MAXP[c], nn);
        }
        CUDA_TRY(cudaStreamSynchronize(s->stream), B200_ERROR_CUDA_SYNCHRONIZE);
    }
    if (dinv) {
        CUDA_TRY(cudaMemcpy(dinv, s->d_dinv, std::min<int64_t>(dinv_len, (*s->plan_sp).dinv_size) * sizeof(double), cudaMemcpyDeviceToHost), B200_ERROR_CUDA_MEMCPY);
        if (s->n_subtrees > 0 && !s->inv_skip_ptr.empty() && s->inv_skip_ptr[NIC] > 0) { // give the subtree fronts their packed pivot blocks back
            const int nn = s->inv_skip_ptr[NIC];
            k_pack_pivot_blocks<<<std::min((nn + 7) / 8, 148 * 8), 256, 0, s->stream>>>(s->d_inv_skip, nn, s->d_nodes, s->d_fac, s->d_dinv);
            CUDA_TRY(cudaStreamSynchronize(s->stream), B200_ERROR_CUDA_SYNCHRONIZE);
        }
    }
    if (fac) CUDA_TRY(cudaMemcpy(fac, s->d_fac, std::min<int64_t>(fac_len, (*s->plan_sp).fac_size) * sizeof(double), cudaMemcpyDeviceToHost), B200_ERROR_CUDA_MEMCPY);
    if (lperm) CUDA_TRY(cudaMemcpy(lperm, s->d_lperm, std::min<int64_t>(n, s->n) * sizeof(int), cudaMemcpyDeviceToHost), B200_ERROR_CUDA_MEMCPY);
    return B200_SUCCESSFUL_EXIT;
}

} // extern "C"
