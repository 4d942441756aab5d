"""Shared test helpers: golden fixtures + synthetic matrix generators (restating the reference's generators)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def load_samples():
    with open(os.path.join(GOLDEN, "samples.json")) as f:
        return json.load(f)


def load_bfwb62_x():
    with open(os.path.join(GOLDEN, "bfwb62_x.json")) as f:
        return np.array(json.load(f))


def mm_path(name):
    return os.path.join(GOLDEN, "matrix_market", name)


def sample_coo(name):
    """builds a russell_b200.CooMatrix from a golden sample (Samples::<name>, russell_sparse/src/samples.rs)"""
    from russell_b200 import CooMatrix, Sym

    s = load_samples()[name]
    coo = CooMatrix(s["nrow"], s["ncol"], s["max_nnz"], Sym[s["sym"]])
    for i, j, v in zip(s["coo_i"], s["coo_j"], s["coo_v"]):
        coo.put(i, j, v)
    return coo, s


def laplacian_2d_triplets(k, lower=False):
    """5-point Laplacian on a k x k interior grid, A = I (x) T + T (x) I, T = tridiag(-1, 2, -1), row m = i + j*k.

    Triplets come in the reference's stencil order (cur, left, right, bottom, top per node:
    russell_pde/src/fdm_2d.rs:944-979, loop_over_bandwidth), Dirichlet neighbours dropped like get_matrices_sps
    does (fdm_2d.rs:603-649); with `lower` the entries above the diagonal are skipped (Sym::YesLower)."""
    n = k * k
    m = np.arange(n, dtype=np.int64)
    i, j = m % k, m // k
    cols = np.stack([m, m - 1, m + 1, m - k, m + k], axis=1)
    ok = np.stack([np.ones(n, bool), i > 0, i < k - 1, j > 0, j < k - 1], axis=1)
    vals = np.tile(np.array([4.0, -1.0, -1.0, -1.0, -1.0]), (n, 1))
    rows = np.repeat(m[:, None], 5, axis=1)
    if lower:
        ok &= cols <= rows
    sel = ok.ravel()
    return n, rows.ravel()[sel].astype(np.int32), cols.ravel()[sel].astype(np.int32), vals.ravel()[sel]


def laplacian_2d_coo(k, lower=False):
    from russell_b200 import CooMatrix, Sym

    n, ai, aj, ax = laplacian_2d_triplets(k, lower)
    return CooMatrix.from_triplets(n, n, ai, aj, ax, Sym.YesLower if lower else Sym.No)


def convection_diffusion_triplets(k, peclet=0.4):
    """unsymmetric 5-point operator (upwinded convection): same pattern as the Laplacian, unsymmetric values"""
    n, ai, aj, ax = laplacian_2d_triplets(k)
    ax = ax.copy()
    off = ai != aj
    left = off & (aj == ai - 1)
    right = off & (aj == ai + 1)
    ax[left] -= peclet
    ax[right] += peclet
    return n, ai, aj, ax


def saddle_point_triplets(k, ncon=None):
    """[K C^T; C 0] with K the k x k-grid Laplacian and C selecting/averaging a few unknowns: the LMM shape of
    russell_pde (fdm_2d.rs:651-662), zero diagonal block -> needs matching or pivoting"""
    n, ai, aj, ax = laplacian_2d_triplets(k)
    if ncon is None:
        ncon = max(1, k // 2)
    rng = np.random.default_rng(7)
    rows, cols, vals = list(ai), list(aj), list(ax)
    picks = rng.choice(n, size=ncon, replace=False)
    for c, m in enumerate(picks):
        r = n + c
        for mm, v in ((m, 1.0), ((m + 1) % n, 0.5)):
            rows += [r, mm]
            cols += [mm, r]
            vals += [v, v]
    nt = n + ncon
    return nt, np.array(rows, dtype=np.int32), np.array(cols, dtype=np.int32), np.array(vals)
