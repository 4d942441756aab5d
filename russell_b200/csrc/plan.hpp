// plan.hpp -- host-side analysis ("symbolic phase") of the B200 multifrontal LU.
//
// This is the part that `umfpack_di_symbolic` (russell_sparse/c_code/interface_umfpack.c:109) and
// `cudssExecute(CUDSS_PHASE_ANALYSIS)` (russell_sparse/c_code/interface_cudss.cu:361) perform for the
// reference: it runs ONCE per sparsity structure, on the CPU, and produces everything the device
// kernels need as flat arrays (no pointers), so a numeric re-factorization is pure kernel launches.
//
// Pipeline:  CSR pattern -> (optional) max-product matching + scaling -> A+A^T graph ->
//            nested dissection + local minimum degree -> elimination tree -> postorder ->
//            column counts -> relaxed supernodes -> row structures -> panel splitting ->
//            assembly-tree levels, relative indices, value scatter map, work-item lists.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace b200 {

// ---- options -------------------------------------------------------------------------------

enum OrderingKind {
    ORDERING_ND = 0,      // nested dissection + local minimum degree (default)
    ORDERING_MINDEG = 1,  // minimum degree only (small problems / comparison)
    ORDERING_NATURAL = 2, // identity (Ordering::No)
};

struct AnalyzeOptions {
    int ordering = ORDERING_ND;
    int matching = 0;          // 0 = none, 1 = max-product matching + scaling, 2 = auto (only if a diagonal is weak)
    int panel_width = 64;      // max pivots per front node (wide supernodes are split into chains)
    int nd_leaf = 32;          // dissection stops below this many vertices; local minimum degree takes over
    int relax_small = 8;       // always merge a last child into its parent while the merged width <= this
    double relax_z1 = 0.3;     // merged width <= 32 : allowed fraction of explicit zeros (0.6 until r01w: with the leaf/subtree
                               // kernels small fronts are cheap, 0.3 stores 12 % fewer entries: 9.37 -> 8.84 ms per step at config 2)
    double relax_z2 = 0.25;    // merged width <= panel_width : allowed fraction of explicit zeros
    double relax_z3 = 0.05;    // wider
    // solve phase, bottom of the tree: maximal subtrees made of small fronts are walked by ONE CTA each with the subtree's
    // segment of the solution vector in shared memory (sweep_sub.cuh).  Their L panels, U panels and pivot-block copies are
    // laid out contiguously per subtree so that one bulk copy (TMA, cp.async.bulk) stages everything a CTA needs.
    int st_enable = 1;
    int st_maxf = 96;          // largest front order inside a subtree
    int st_pmax = 32;          // most pivots per front inside a subtree
    int st_budget = 8192;      // stored L entries (padded) per subtree
    int st_maxcols = 1024;     // columns of a subtree + update rows of its root (shared-memory solution segment of one warp)
    int st_min_count = 64;     // fewer subtrees than this are not worth a launch of their own
    bool cb_reuse = true;      // contribution blocks with disjoint lifetimes (tree levels) share storage; a walk in
                               // postorder (the test suite's scalar host walk) needs private, zero-initialised blocks: false
    int verbose = 0;
};

// ---- front tree ------------------------------------------------------------------------------
//
// A *node* is one frontal matrix with `p` pivot columns (first column `c0` in the permuted order) and
// `u` update rows.  Front order f = p + u.  Device storage per node (column-major):
//   L panel : f x p  at fac[Loff]   rows 0..p-1 hold the pivot block (L11\U11 after factorization),
//                                    rows p..f-1 hold L21
//   U panel : u x p  at fac[Uoff]   holds U12^T (row j = update column j)
//   C block : u x u  at cb[Coff]    Schur complement / contribution block handed to the parent
//   D block : p x p  at dinv[Doff]  strictly-lower part = inv(L11) (unit diag implied), upper = inv(U11)
struct Plan {
    int n = 0;
    int nnz_in = 0;           // nnz of the caller's CSR (lower triangle only when sym_lower)
    bool sym_lower = false;   // caller passed the lower triangle of a symmetric matrix
    AnalyzeOptions opt;

    // permutation + scaling:  A''[k][l] = rscale[rowperm[k]] * A[rowperm[k]][colperm[l]] * cscale[colperm[l]]
    std::vector<int> rowperm, colperm;       // new -> old
    std::vector<double> rscale, cscale;      // indexed by ORIGINAL row / column; empty => all ones
    bool matched = false;                    // a non-identity row matching is in use

    // nodes in postorder (children before parents; a node's columns follow its descendants')
    int nnodes = 0;
    std::vector<int> c0, p, u, parent, level;
    std::vector<int64_t> Loff, Uoff, Coff, Doff;
    std::vector<int64_t> rows_ptr;           // nnodes+1 : offsets into rows[] / rel[]
    std::vector<int> rows;                   // update-row global (permuted) indices, ascending per node
    std::vector<int> rel;                    // position of each update row inside the PARENT's front (0..f_parent-1)
    std::vector<int> child_ptr, child_idx;   // children lists
    int nlevels = 0;
    std::vector<int> level_ptr, level_nodes; // nodes grouped by level (level 0 = leaves)

    int64_t fac_size = 0, cb_size = 0, dinv_size = 0;

    // subtree region of the solve phase: in_sub[v] = front v is walked by a subtree CTA; st_first/st_root = node ranges
    // (a subtree is the contiguous postorder range [first, root]); within a subtree the L panels are contiguous in `fac`
    // (first..root), then the U panels; Doff is contiguous in node order anyway
    std::vector<char> in_sub;
    std::vector<int> st_first, st_root;

    // value scatter map (user CSR slot -> fac offset); symmetric-lower input contributes two entries
    std::vector<int> a_src;
    std::vector<int64_t> a_dst;
    std::vector<double> a_scl;               // empty => no scaling

    // full (mirrored) CSR of the ORIGINAL matrix for residual SpMV; full_src maps to the caller's slot
    std::vector<int> full_ptr, full_col, full_src; // full_src empty => identity (general input)

    // stats
    int64_t nnz_L = 0, nnz_U = 0;            // stored entries incl. explicit zeros (L incl. pivot blocks)
    double flops = 0.0;                      // LU flops of the front tree
    int nsuper_fundamental = 0, nsuper_relaxed = 0;
    int max_front = 0;
    double t_match = 0, t_order = 0, t_symbolic = 0;
};

// Runs the whole analysis.  `vals` is used only by the matching/scaling step (may be null when matching = 0).
// Returns 0 or a negative error (‑1 structurally singular, ‑2 invalid input).
int analyze(int n, const int* rowptr, const int* colidx, const double* vals, bool sym_lower,
            const AnalyzeOptions& opt, Plan& plan);

// ---- pieces (exposed for unit tests) -----------------------------------------------------------

struct Graph {
    int n = 0;
    std::vector<int> ptr, adj; // symmetric, no self loops, sorted neighbours
};

// fill-reducing ordering of a symmetric graph; perm is new -> old
void order_nested_dissection(const Graph& g, int leaf_size, std::vector<int>& perm);
void order_minimum_degree(const Graph& g, std::vector<int>& perm);

// maximum-product bipartite matching with scaling (Duff & Koster style) on a CSR matrix.
// rowmatch[j] = row matched to column j.  Returns the number of matched columns (== n on success).
int max_product_matching(int n, const int* rowptr, const int* colidx, const double* vals,
                         std::vector<int>& rowmatch, std::vector<double>& rscale, std::vector<double>& cscale);

void etree_symmetric(const Graph& g, std::vector<int>& parent);
void postorder_tree(const std::vector<int>& parent, const std::vector<int>& weight, std::vector<int>& post);
void column_counts(const Graph& g, const std::vector<int>& parent, std::vector<int>& cc);

} // namespace b200
