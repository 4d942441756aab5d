#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 sparse direct solver backend (BASELINE.json metric).

A *step* is one numeric re-factorization + one solve (with iterative refinement) of BASELINE.json configs[1]:
the 5-point 2D Laplacian on a 1000 x 1000 grid (1M dof, 4,996,000 nnz, f64, b = ones), i.e. what a Newton /
Radau5 loop issues per Jacobian update (structure analysed once, outside the timed region, exactly like the
reference's `initialize`).  Metric: factorize+solve per second, whole job (all GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--grid 1000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference   # the CPU path (SuperLU stand-in for UMFPACK, see oracle/oracle.py)

Prints ONE JSON line (rank 0).  `value` = device-resident loop; `e2e` = the same loop through the five
reference-facing C-ABI calls with pinned HOST buffers (H2D of values+rhs and D2H of x inside the timed region).
Multi-GPU: one independent system per rank (weak scaling), no data-path collective; the solutions are gathered
with one NCCL all_gather per step.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def laplacian_csr(k, lower=False):
    """CSR of the 5-point Laplacian through the product's own host converter (COO in stencil order -> CSR)"""
    import helpers
    import russell_b200 as rb

    coo = helpers.laplacian_2d_coo(k, lower=lower)
    csr = rb.CsrMatrix.from_coo(coo)
    n = csr.nnz
    return coo, csr.pointers.copy(), csr.indices[:n].copy(), csr.values[:n].copy()


# the kernels one SpTRSV sweep launches at HEAD (library defaults): a committed traffic capture counts only if it is a
# capture of exactly these kernels -- a capture older than the kernels is refused (roofline.traffic = null)
SWEEP_KERNELS = ["k_permute_in", "k_fwd_stree_w", "k_fwd_top2", "k_bwd_top3", "k_bwd_stree_w", "k_permute_out"]


def measured_traffic(grid):
    """DRAM bytes (read + write) of one SpTRSV sweep from the committed ncu capture of the same workload
    (profiles/*_sptrsv_traffic.json, written by tools/traffic_json.py from `ncu --metrics dram__bytes_read.sum,
    dram__bytes_write.sum`); None when no capture of this grid size AND of the current kernels is committed."""
    import glob
    best = (None, None)
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_sptrsv_traffic.json"))):
        try:
            with open(f) as fh:
                d = json.load(fh)
            names = [k for k in d.get("kernels", {}) if "spmv" not in k]
            if int(d.get("grid", -1)) == int(grid) and sorted(names) == sorted(SWEEP_KERNELS):
                best = (float(d["sptrsv_sweep_traffic_bytes"]), "profiles/" + os.path.basename(f))
        except Exception:
            pass
    return best


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for s in self.samples:
            t = [x.strip() for x in s.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0])), smax.append(float(t[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU arm -------------------------------------------------------------------------------------------------------
# The reference's CPU path is UMFPACK (russell_sparse/c_code/interface_umfpack.c:109,167,229).  UMFPACK is neither in
# /root/reference (un-vendored SuiteSparse) nor on the GPU box (probed: profiles/r2_probe_gpu_box_umfpack.txt -- no
# umfpack.h, no libumfpack, no scikits.umfpack), so oracle/_ref cannot exist and the arm runs the oracle's CPU LU: SuperLU
# through scipy, a STAND-IN with a different algorithm.  To compare like with like the arm is timed the way the GPU arm is:
#   warm = numeric refactorization + solve with the column ordering computed once outside the timed region (SuperLU has
#          no separate symbolic phase to reuse, so its symbolic work stays inside: this favours the GPU arm a little);
#   cold = ordering + factorization + solve (against the GPU arm's `e2e_cold` = initialize + factorize + solve).
# With --gpus N the arm solves N systems at once in N processes, like the GPU arm solves one per GPU.
def cpu_system(k, scale=1.0):
    from oracle import oracle
    import helpers

    n, ai, aj, ax = helpers.laplacian_2d_triplets(k)
    return oracle.full_scipy_matrix(n, n, ai, aj, ax * scale), np.ones(n)


def cpu_reference_step(k, state={}):
    """one COLD factorize+solve (ordering inside) -- also what `cpu_baseline` of the GPU line reports"""
    from oracle import oracle

    if k not in state:
        state[k] = cpu_system(k)
    a, b = state[k]
    t0 = time.perf_counter()
    lu = oracle.lu_factorize(a, "MMD_AT_PLUS_A")
    x = lu.solve(b)
    t1 = time.perf_counter()
    res = float(np.linalg.norm(b - a @ x) / np.linalg.norm(b))
    return t1 - t0, res, lu.perm_c


def _cpu_worker(args):
    """one process of the reference arm: `warmup` + `steps` warm refactorize+solve of its own system"""
    k, idx, steps, warmup = args
    import scipy.sparse.linalg as spl

    a, b = cpu_system(k, 1.0 + 0.01 * idx)
    t0 = time.perf_counter()
    lu = spl.splu(a, permc_spec="MMD_AT_PLUS_A")
    x = lu.solve(b)
    t_cold = time.perf_counter() - t0
    # Pr A Pc = L U with Pc[j, perm_c[j]] = 1: column j of A is eliminated at position perm_c[j].  Put the columns in that
    # order once; the refactorizations below skip the ordering (permc_spec NATURAL)
    ap = a[:, np.argsort(lu.perm_c)].tocsc()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        lu2 = spl.splu(ap, permc_spec="NATURAL")
        y = lu2.solve(b)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    x2 = y[lu.perm_c]  # x = Pc y
    res = float(np.linalg.norm(b - a @ x2) / np.linalg.norm(b))
    return t_cold, times, res


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the CPU path timed on the box's host cores (rank 0 only; N systems in N processes)"""
    if rank != 0:
        return
    import multiprocessing as mp

    k = args.grid
    nsys = max(1, args.gpus)
    budget_s = 150.0
    # size the run from one cold step: K warm steps per process, bounded to ~budget seconds of wall time
    t_probe, _, _ = cpu_reference_step(k)
    steps = max(1, min(args.steps, int(budget_s / max(t_probe, 1e-3)) - 1))
    warmup = 1 if steps > 1 and args.warmup > 0 else 0
    steps = max(1, steps - warmup)
    w0 = time.perf_counter()
    if nsys == 1:
        results = [_cpu_worker((k, 0, steps, warmup))]
    else:
        with mp.get_context("spawn").Pool(nsys) as pool:
            results = pool.map(_cpu_worker, [(k, i, steps, warmup) for i in range(nsys)])
    wall = time.perf_counter() - w0
    per_step = max(float(np.mean(r[1])) for r in results)  # the slowest process sets the pace, like max over ranks
    val = nsys / per_step
    cold = max(r[0] for r in results)
    res = max(r[2] for r in results)
    line = {
        "impl": "reference", "metric": "factorize+solve/sec", "value": val, "unit": "systems/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(k, args.gpus),
        "rel_residual": res,
        "cold": {"value": nsys / cold, "unit": "systems/s", "ms_per_step": 1e3 * cold,
                 "what": "ordering (MMD on A'+A) + factorization + solve, one shot: compare with the GPU arm's e2e_cold"},
        "cpu_baseline": {"value": val, "unit": "systems/s", "cores": nsys, "kind": "port",
                         "sample": "%d process(es) x %d warm refactorize+solve (column ordering precomputed, splu NATURAL) of the %dx%d "
                                   "Laplacian with scipy SuperLU, sequential per process; STAND-IN for UMFPACK, which is not installed on "
                                   "this box (profiles/r2_probe_gpu_box_umfpack.txt); host cores available: %d"
                                   % (nsys, steps, k, k, host_threads())},
        "e2e": {"value": val, "unit": "systems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "requested steps=%d warmup=%d; CPU steps bounded to ~%.0f s of work; wall %.1f s" % (args.steps, args.warmup, budget_s, wall),
    }
    print(json.dumps(line), flush=True)


def workload_config(k, gpus):
    n = k * k
    return {"workload": "5-point 2D Laplacian %dx%d (n=%d, nnz=%d), b=ones, numeric refactorization + solve per step "
                        "(BASELINE.json configs[1]); one independent system per GPU" % (k, k, n, 5 * n - 4 * k),
            "grid": k, "n": n, "nnz": 5 * n - 4 * k, "systems_per_step": gpus, "parallelism": "one-matrix-per-gpu x%d" % gpus,
            "l2_policy": "working set (factors + contribution blocks > 2 GB) exceeds the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--grid", type=int, default=1000)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import russell_b200 as rb
    from russell_b200 import _lib, batch
    from russell_b200._lib import p_f64, p_i32, ptr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W = max(args.warmup, 3)
    K = args.steps
    k = args.grid

    # ---- set-up (untimed): structure analysis + device plan, like the reference's `initialize` -----------------
    coo, rp, ci, vals = laplacian_csr(k)
    n, nnz = len(rp) - 1, len(vals)
    vals = vals * (1.0 + 0.01 * rank)  # every rank owns a different system
    lib = _lib.load()
    h = lib.solver_b200_new()
    if not h:
        raise SystemExit("solver_b200_new failed: no CUDA device")
    assert lib.solver_b200_set_option(h, b"device", float(local_rank)) == 0
    t0 = time.perf_counter()
    rc = lib.solver_b200_initialize(h, 0, 0, 0, -1.0, -1, -1.0, 0, 0, 0, n, ptr(rp, p_i32), ptr(ci, p_i32), ptr(vals, p_f64))
    if rc != 0:
        raise SystemExit("solver_b200_initialize failed: %d" % rc)
    t_init = time.perf_counter() - t0
    # the one-shot cost a caller like Fdm2d::solve_sps pays (russell_pde/src/fdm_2d.rs:439-462): first factorize + solve
    # through the five-call API, pageable host buffers, graph capture included
    x_first = np.zeros(n)
    rhs_first = np.ones(n)
    em0, ep0 = ctypes.c_int32(0), ctypes.c_int32(0)
    t0 = time.perf_counter()
    assert lib.solver_b200_factorize(h, ctypes.byref(em0), ctypes.byref(ep0), 0, ptr(vals, p_f64)) == 0
    assert lib.solver_b200_solve(h, ptr(x_first, p_f64), ptr(rhs_first, p_f64), 0) == 0
    t_first = time.perf_counter() - t0
    stream = torch.cuda.ExternalStream(lib.solver_b200_get_stream(h), device=torch.device("cuda", local_rank))

    # pinned host buffers (what the Rust wrapper's Vec<f64> would be, page-locked) and device-resident copies
    h_vals = torch.from_numpy(vals).pin_memory()
    h_rhs = torch.ones(n, dtype=torch.float64).pin_memory()
    h_x = torch.zeros(n, dtype=torch.float64).pin_memory()
    d_vals = h_vals.cuda()
    d_rhs = h_rhs.cuda()
    # one system per rank: the solve writes x straight into this rank's slot of the preallocated gather buffer
    gatherer = batch.SolutionGatherer(world, world, rank, n, torch.float64, torch.device("cuda", local_rank))
    d_x = gatherer.block[0]
    em, ep = ctypes.c_int32(0), ctypes.c_int32(0)

    def step_device():
        rc = lib.solver_b200_factorize_device(h, d_vals.data_ptr())
        assert rc == 0, rc
        rc = lib.solver_b200_solve_device(h, d_x.data_ptr(), d_rhs.data_ptr())
        assert rc == 0, rc
        if world > 1:
            gatherer.gather()  # system r lives on rank r; (the solve above returned synchronised: x is complete)
            stream.wait_stream(torch.cuda.current_stream())  # the timing events live on the solver's stream: they must see the gather

    def step_e2e():
        rc = lib.solver_b200_factorize(h, ctypes.byref(em), ctypes.byref(ep), 0, ctypes.cast(h_vals.data_ptr(), p_f64))
        assert rc == 0, rc
        rc = lib.solver_b200_solve(h, ctypes.cast(h_x.data_ptr(), p_f64), ctypes.cast(h_rhs.data_ptr(), p_f64), 0)
        assert rc == 0, rc
        if world > 1:
            d_x.copy_(h_x, non_blocking=True)
            gatherer.gather()
            stream.wait_stream(torch.cuda.current_stream())

    def get_stats():
        out = np.zeros(len(rb.SolverB200.STAT_NAMES))
        lib.solver_b200_get_stats(h, ptr(out, p_f64), len(out))
        return dict(zip(rb.SolverB200.STAT_NAMES, out.tolist()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; CUDA events on the solver's own stream; max over ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        e0.record(stream)
        acc = {"fact": 0.0, "solve": 0.0, "sptrsv": 0.0, "spmv": 0.0}
        for _ in range(steps):
            fn()
            st = get_stats()
            acc["fact"] += st["ms_factorize_device"]
            acc["solve"] += st["ms_solve_device"]
            acc["sptrsv"] += st["ms_sptrsv_device"]
            acc["spmv"] += st["ms_spmv_device"]
        e1.record(stream)
        barrier()
        w1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, 1e3 * (w1 - w0)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), {k2: v / steps for k2, v in acc.items()}

    for _ in range(W):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, wall_dev, parts = timed(step_device, K)
    clocks = sampler.stop() if rank == 0 else None
    st = get_stats()
    for _ in range(W):
        step_e2e()
    ms_e2e, wall_e2e, parts_e2e = timed(step_e2e, K)
    res_e2e = get_stats()["last_rel_residual"]

    # the same five-call loop with PAGEABLE host buffers (plain numpy arrays -- what a Rust Vec<f64> is): with the library's
    # striped staging through pinned buffers (default) and with plain cudaMemcpyAsync from the caller's pages
    pageable = None
    if world == 1:
        pv, pr, px = vals.copy(), np.ones(n), np.zeros(n)

        def step_pageable():
            rc = lib.solver_b200_factorize(h, ctypes.byref(em), ctypes.byref(ep), 0, ptr(pv, p_f64))
            assert rc == 0, rc
            rc = lib.solver_b200_solve(h, ptr(px, p_f64), ptr(pr, p_f64), 0)
            assert rc == 0, rc

        pageable = {"unit": "systems/s", "what": "e2e loop with pageable (numpy) host buffers; staged = striped copies through the "
                                                 "handle's pinned staging buffers (default), plain = cudaMemcpyAsync from the caller's pages"}
        for label, flag in (("staged", 1.0), ("plain", 0.0)):
            assert lib.solver_b200_set_option(h, b"staged_copy", flag) == 0
            for _ in range(2):
                step_pageable()
            ms_p, wall_p, _ = timed(step_pageable, min(K, 10))
            pageable[label] = {"value": min(K, 10) / (wall_p * 1e-3), "wall_ms_per_step": wall_p / min(K, 10), "ms_per_step": ms_p / min(K, 10)}
        assert lib.solver_b200_set_option(h, b"staged_copy", 1.0) == 0
        pageable["x_equals_pinned_run"] = bool(np.array_equal(px, h_x.numpy()))

    # accuracy (north star): ||b - A x|| / ||b|| of the last device-resident solve, evaluated in f64 by the SpMV kernel
    rel_res = st["last_rel_residual"]
    t_res = torch.tensor([rel_res, res_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
    rel_res, res_e2e = float(t_res[0]), float(t_res[1])

    # independent accuracy check on the HOST (scipy SpMV, not the product's kernel): x of the last e2e step
    res_host = None
    if rank == 0:
        import scipy.sparse as sp

        a_host = sp.csr_matrix((vals, ci, rp), shape=(n, n))
        xh = h_x.numpy()
        bh = h_rhs.numpy()
        res_host = float(np.linalg.norm(bh - a_host @ xh) / np.linalg.norm(bh))
    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        sptrsv_gbs = st["sptrsv_bytes"] / (parts["sptrsv"] * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(k)
        spmv_gbs = st["spmv_bytes"] / (parts["spmv"] * 1e-3) / 1e9 if parts["spmv"] > 0 else None
        line = {
            "metric": "factorize+solve/sec", "value": world * K / (ms_dev * 1e-3), "unit": "systems/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(k, world),
            "rel_residual": rel_res,
            "rel_residual_host_checked": res_host,
            "e2e": {"value": world * K / (ms_e2e * 1e-3), "unit": "systems/s", "h2d_bytes_per_step": 8 * nnz + 8 * n,
                    "d2h_bytes_per_step": 8 * n, "ms_per_step": ms_e2e / K, "wall_ms_per_step": wall_e2e / K,
                    "rel_residual": res_e2e, "api": "solver_b200_factorize + solver_b200_solve (pinned host buffers)"},
            "e2e_pageable": pageable,
            "e2e_cold": {"value": world / (t_init + t_first), "unit": "systems/s", "initialize_s": t_init, "first_factorize_solve_s": t_first,
                         "what": "initialize (host analysis + plan upload) + first factorize + solve, one shot per system; compare with the reference arm's `cold`"},
            "gpu_launches": int(K * (st["launches_factorize"] + st["launches_solve"])),
            "clocks": clocks,
            "roofline": {"kernel": "SpTRSV sweep = k_fwd_stree_w + k_fwd_top2 + k_bwd_top3 + k_bwd_stree_w (one forward+backward solve over the whole front tree; algorithmic bytes = 8 B per stored factor entry + 16 B per unknown)",
                         "bound": "hbm", "achieved": sptrsv_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": sptrsv_gbs / hbm_peak,
                         "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes": st["sptrsv_bytes"], "ms": parts["sptrsv"]},
            "roofline_factorize": {"kernel": "numeric factorization (all kernels of one refactorization, CUDA events)", "flops": st["flops"],
                                   "achieved_tflops": st["flops"] / (parts["fact"] * 1e-3) / 1e12, "peak_tflops": 40.0,
                                   "peak_source": "nominal B200 FP64 (no measured FP64 figure in MEASURED_PEAKS.json)",
                                   "frac": st["flops"] / (parts["fact"] * 1e-3) / 1e12 / 40.0,
                                   "hbm_lower_bound_bytes": 12.0 * nnz + 8.0 * (st["nnz_l"] + st["nnz_u"]),
                                   "hbm_lower_bound_frac": (12.0 * nnz + 8.0 * (st["nnz_l"] + st["nnz_u"])) / (parts["fact"] * 1e-3) / 1e9 / hbm_peak,
                                   "ms": parts["fact"], "note": "latency bound: ~60 dependent tree levels of small kernels, neither bound is tight (DESIGN.md 6)"},
            "phases_ms": {"factorize_device": parts["fact"], "solve_device": parts["solve"], "sptrsv_sweep": parts["sptrsv"],
                          "residual_spmv": parts["spmv"], "initialize_once_s": t_init,
                          "factorize_tflops": st["flops"] / (parts["fact"] * 1e-3) / 1e12,
                          "spmv_gbs": spmv_gbs, "refine_steps": st["last_refine_steps"]},
            "symbolic": {"fronts": st["nnodes"], "levels": st["nlevels"], "nnz_LU": st["nnz_l"] + st["nnz_u"], "flops": st["flops"],
                         "max_front": st["max_front"]},
            "wall_ms_per_step": wall_dev / K,
        }
        if not args.no_cpu_baseline and world == 1:
            t_cold, t_warm, res_cpu = _cpu_worker((k, 0, 2, 0))
            t_warm = float(np.mean(t_warm))
            line["cpu_baseline"] = {"value": 1.0 / t_warm, "unit": "systems/s", "cores": 1, "kind": "port",
                                    "cold_value": 1.0 / t_cold,
                                    "sample": "2 warm refactorize+solve (column ordering precomputed) of the same %dx%d Laplacian with scipy SuperLU, "
                                              "sequential: %.2f s each (cold, with ordering: %.2f s), rel.residual %.1e; STAND-IN for UMFPACK, which "
                                              "is not installed on the box" % (k, k, t_warm, t_cold, res_cpu)}
        print(json.dumps(line), flush=True)
    # tear down in dependency order: the gather buffers were used on the solver's stream (torch's allocator records an event
    # on that stream when it frees them), and NCCL's last collectives were enqueued there -- both must go before
    # solver_b200_drop destroys the stream
    import gc

    torch.cuda.synchronize()
    del d_x
    gatherer.block = gatherer.out = gatherer.index = None
    del gatherer
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    torch.cuda.synchronize()
    lib.solver_b200_drop(h)


if __name__ == "__main__":
    main()
