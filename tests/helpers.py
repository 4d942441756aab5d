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


def load_complex_samples():
    with open(os.path.join(GOLDEN, "complex_samples.json")) as f:
        return json.load(f)


def complex_sample_coo(name):
    """russell_b200.ComplexCooMatrix from a golden complex sample (Samples::<name>, russell_sparse/src/samples.rs)"""
    from russell_b200 import ComplexCooMatrix, Sym

    s = load_complex_samples()[name]
    coo = ComplexCooMatrix(s["nrow"], s["ncol"], s["max_nnz"], Sym[s["sym"]])
    for i, j, v in zip(s["coo_i"], s["coo_j"], s["coo_v"]):
        coo.put(i, j, complex(*v))
    return coo, s


# Radau5 constants (russell_ode/src/radau5.rs:697-699)
RADAU5_ALPHA = 2.6810828736277521338957907432111121010270319565630
RADAU5_BETA = 3.0504301992474105694263776247875679044407041991795
RADAU5_GAMMA = 3.6378342527444957322084185135777757979459360868739


def brusselator_radau5_triplets(npoint, h=1e-4, alpha=0.1):
    """The two Newton matrices Radau5 factorizes for the Brusselator PDE (BASELINE.json configs[3]):
    K_real = (gamma/h) I - J and K_comp = ((alpha_r + i beta_r)/h) I - J, as COO triplets WITH duplicates in the
    reference's order of `put` calls.

    J restates Samples::brusselator_pde(alpha, npoint, second_book=true, ignore_diffusion=false)'s Jacobian at y0
    (russell_ode/src/samples.rs:549-571,598-602): per grid point m (m = i + j*nx) the four reaction entries, then the
    five-point molecule {2(kx/dx^2+ky/dy^2), -kx/dx^2 x2, -ky/dy^2 x2} with kx = ky = -alpha on a periodic grid
    (russell_pde/src/fdm_2d.rs:944-979) for the U block and the V block; Radau5 then appends ndim diagonal entries
    (russell_ode/src/radau5.rs:226-237).  Returns (ndim, ai, aj, k_real_values, k_comp_values): 16*npoint^2 triplets."""
    nx = ny = npoint
    s = nx * ny
    ndim = 2 * s
    dx = 1.0 / (nx - 1)
    dy = 1.0 / (ny - 1)
    kx = ky = -alpha
    mol = np.array([2.0 * (kx / dx**2 + ky / dy**2), -kx / dx**2, -kx / dx**2, -ky / dy**2, -ky / dy**2])
    m = np.arange(s, dtype=np.int64)
    i, j = m % nx, m // nx
    x, y = i * dx, j * dy
    um = 22.0 * y * np.power(1.0 - y, 1.5)
    vm = 27.0 * x * np.power(1.0 - x, 1.5)
    um2 = um * um
    fin_x, fin_y = nx - 1, ny - 1
    nn = np.stack([m, np.where(i != 0, m - 1, m + fin_x), np.where(i != fin_x, m + 1, m - fin_x),
                   np.where(j != 0, m - nx, m + fin_y * nx), np.where(j != fin_y, m + nx, m - fin_y * nx)], axis=1)
    # per point: 4 reaction entries, then (m, n) and (s+m, s+n) interleaved for the 5 molecule entries
    rows = np.empty((s, 14), dtype=np.int64)
    cols = np.empty((s, 14), dtype=np.int64)
    vals = np.empty((s, 14))
    rows[:, 0], cols[:, 0], vals[:, 0] = m, m, -4.4 + 2.0 * um * vm
    rows[:, 1], cols[:, 1], vals[:, 1] = m, s + m, um2
    rows[:, 2], cols[:, 2], vals[:, 2] = s + m, m, 3.4 - 2.0 * um * vm
    rows[:, 3], cols[:, 3], vals[:, 3] = s + m, s + m, -um2
    for b in range(5):
        rows[:, 4 + 2 * b], cols[:, 4 + 2 * b], vals[:, 4 + 2 * b] = m, nn[:, b], mol[b]
        rows[:, 5 + 2 * b], cols[:, 5 + 2 * b], vals[:, 5 + 2 * b] = s + m, s + nn[:, b], mol[b]
    d = np.arange(ndim, dtype=np.int64)
    ai = np.concatenate([rows.ravel(), d]).astype(np.int32)
    aj = np.concatenate([cols.ravel(), d]).astype(np.int32)
    jv = -vals.ravel()  # K = -J + ...
    k_real = np.concatenate([jv, np.full(ndim, RADAU5_GAMMA / h)])
    k_comp = np.concatenate([jv.astype(np.complex128), np.full(ndim, complex(RADAU5_ALPHA / h, RADAU5_BETA / h))])
    return ndim, ai, aj, k_real, k_comp


def laplacian_3d_triplets(k, skew=0.0):
    """7-point Laplacian on a k^3 grid (Dirichlet), optionally with an index-seeded skew perturbation of the off-diagonal
    entries (+-skew*sin(i*j)): the small-scale analogue of the stand-in SURVEY 8d proposes for af_shell10 (BASELINE.json
    configs[2], which is not in the tree): large separator fronts (k^2 vertices), unsymmetric values."""
    n = k * k * k
    m = np.arange(n, dtype=np.int64)
    i, j, l = m % k, (m // k) % k, m // (k * k)
    cols = np.stack([m, m - 1, m + 1, m - k, m + k, m - k * k, m + k * k], axis=1)
    ok = np.stack([np.ones(n, bool), i > 0, i < k - 1, j > 0, j < k - 1, l > 0, l < k - 1], axis=1)
    vals = np.tile(np.array([6.0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0]), (n, 1))
    rows = np.repeat(m[:, None], 7, axis=1)
    sel = ok.ravel()
    ai, aj, ax = rows.ravel()[sel], cols.ravel()[sel], vals.ravel()[sel].copy()
    if skew != 0.0:
        off = ai != aj
        ax[off] += skew * np.sin((ai[off] + 1.0) * (aj[off] + 1.0))
    return n, ai.astype(np.int32), aj.astype(np.int32), ax


def laplacian_3d_27pt_triplets(k, skew=1e-3):
    """SURVEY 8d's named stand-in for BASELINE.json configs[2] (af_shell10 is not in the tree and there is no network):
    27-point Laplacian on a k^3 grid (centre 26, every one of the 26 neighbours -1, Dirichlet neighbours dropped) plus the
    index-seeded skew perturbation +-skew*sin((i+1)(j+1)) of the off-diagonal entries, so the VALUES are unsymmetric like
    the general (MakeItFull) path af_shell10 is run through.  k = 115: n = 1,520,875, nnz = 40.4 M."""
    n = k * k * k
    m = np.arange(n, dtype=np.int64)
    i, j, l = m % k, (m // k) % k, m // (k * k)
    ai, aj, ax = [], [], []
    for dl in (-1, 0, 1):
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                ok = (i + di >= 0) & (i + di < k) & (j + dj >= 0) & (j + dj < k) & (l + dl >= 0) & (l + dl < k)
                r = m[ok]
                c = r + di + dj * k + dl * k * k
                v = np.full(len(r), 26.0 if (di == 0 and dj == 0 and dl == 0) else -1.0)
                if skew != 0.0 and not (di == 0 and dj == 0 and dl == 0):
                    v += skew * np.sin((r + 1.0) * (c + 1.0))
                ai.append(r), aj.append(c), ax.append(v)
    ai, aj, ax = np.concatenate(ai), np.concatenate(aj), np.concatenate(ax)
    order = np.argsort(ai, kind="stable")  # row-major triplets, the 27 stencil entries of a row in (dl, dj, di) order
    return n, ai[order].astype(np.int32), aj[order].astype(np.int32), ax[order]


def host_rel_residual(n, ai, aj, ax, x, b):
    """||b - A x||_2 / ||b||_2 evaluated on the HOST with scipy (duplicates summed): the independent check of the
    product's own SpMV-based residual"""
    import scipy.sparse as sp

    a = sp.coo_matrix((ax, (ai, aj)), shape=(n, n)).tocsr()
    r = b - a @ x
    return float(np.linalg.norm(r) / np.linalg.norm(b))
