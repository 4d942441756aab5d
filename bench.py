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


def measured_traffic(grid):
    """DRAM bytes (read + write) of one SpTRSV sweep from the committed ncu capture of the same workload
    (profiles/*_sptrsv_traffic.json, written from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`);
    None when no capture of this grid size is committed."""
    import glob
    best = (None, None)
    for f in sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "*_sptrsv_traffic.json"))):
        try:
            with open(f) as fh:
                d = json.load(fh)
            if int(d.get("grid", -1)) == int(grid):
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


def cpu_reference_step(k, csc_cache={}):
    """one factorize+solve of the same workload on the CPU: SuperLU (scipy) stand-in for the reference's UMFPACK path"""
    from oracle import oracle
    import helpers

    if k not in csc_cache:
        n, ai, aj, ax = helpers.laplacian_2d_triplets(k)
        csc_cache[k] = (oracle.full_scipy_matrix(n, n, ai, aj, ax), np.ones(n))
    a, b = csc_cache[k]
    t0 = time.perf_counter()
    lu = oracle.lu_factorize(a, "MMD_AT_PLUS_A")
    x = lu.solve(b)
    t1 = time.perf_counter()
    res = float(np.linalg.norm(b - a @ x) / np.linalg.norm(b))
    return t1 - t0, res


def host_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return 1


def run_reference(args, rank, world):
    """--impl reference: the CPU path timed on the box's host cores (rank 0 only)"""
    if rank != 0:
        return
    k = args.grid
    budget_s = 170.0
    t_first, res = cpu_reference_step(k)
    times = [t_first]
    steps = max(1, min(args.steps, int(budget_s / max(t_first, 1e-3))))
    for _ in range(steps - 1):
        times.append(cpu_reference_step(k)[0])
    per = float(np.mean(times))
    val = 1.0 / per
    line = {
        "impl": "reference", "metric": "factorize+solve/sec", "value": val, "unit": "systems/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": 0, "ms_per_step": 1e3 * per, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(k, args.gpus),
        "rel_residual": res,
        "cpu_baseline": {"value": val, "unit": "systems/s", "cores": 1, "kind": "port",
                         "sample": "%d full factorize+solve of the %dx%d Laplacian with scipy SuperLU (MMD_AT_PLUS_A), sequential; "
                                   "stand-in for UMFPACK which is not installed (oracle/oracle.py); BLAS threads available: %d"
                                   % (len(times), k, k, host_threads())},
        "e2e": {"value": val, "unit": "systems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "requested steps=%d warmup=%d; CPU steps bounded to ~%.0f s of work" % (args.steps, args.warmup, budget_s),
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
    stream = torch.cuda.ExternalStream(lib.solver_b200_get_stream(h), device=torch.device("cuda", local_rank))

    # pinned host buffers (what the Rust wrapper's Vec<f64> would be, page-locked) and device-resident copies
    h_vals = torch.from_numpy(vals).pin_memory()
    h_rhs = torch.ones(n, dtype=torch.float64).pin_memory()
    h_x = torch.zeros(n, dtype=torch.float64).pin_memory()
    d_vals = h_vals.cuda()
    d_rhs = h_rhs.cuda()
    d_x = torch.zeros(n, dtype=torch.float64, device="cuda")
    em, ep = ctypes.c_int32(0), ctypes.c_int32(0)

    def step_device():
        rc = lib.solver_b200_factorize_device(h, d_vals.data_ptr())
        assert rc == 0, rc
        rc = lib.solver_b200_solve_device(h, d_x.data_ptr(), d_rhs.data_ptr())
        assert rc == 0, rc
        if world > 1:
            batch.gather_solutions(d_x[None, :], world, world, rank)  # system r lives on rank r

    def step_e2e():
        rc = lib.solver_b200_factorize(h, ctypes.byref(em), ctypes.byref(ep), 0, ctypes.cast(h_vals.data_ptr(), p_f64))
        assert rc == 0, rc
        rc = lib.solver_b200_solve(h, ctypes.cast(h_x.data_ptr(), p_f64), ctypes.cast(h_rhs.data_ptr(), p_f64), 0)
        assert rc == 0, rc
        if world > 1:
            batch.gather_solutions(d_x.copy_(h_x, non_blocking=True)[None, :], world, world, rank)

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

    # accuracy (north star): ||b - A x|| / ||b|| of the last device-resident solve, evaluated in f64 by the SpMV kernel
    rel_res = st["last_rel_residual"]
    t_res = torch.tensor([rel_res, res_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
    rel_res, res_e2e = float(t_res[0]), float(t_res[1])

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
            "e2e": {"value": world * K / (ms_e2e * 1e-3), "unit": "systems/s", "h2d_bytes_per_step": 8 * nnz + 8 * n,
                    "d2h_bytes_per_step": 8 * n, "ms_per_step": ms_e2e / K, "wall_ms_per_step": wall_e2e / K,
                    "rel_residual": res_e2e, "api": "solver_b200_factorize + solver_b200_solve (pinned host buffers)"},
            "gpu_launches": int(K * (st["launches_factorize"] + st["launches_solve"])),
            "clocks": clocks,
            "roofline": {"kernel": "SpTRSV sweep = k_fwd_subtree + k_fwd_top2 + k_bwd_top3 + k_bwd_subtree (one forward+backward solve over the whole front tree; algorithmic bytes = 8 B per stored factor entry + 16 B per unknown)",
                         "bound": "hbm", "achieved": sptrsv_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": sptrsv_gbs / hbm_peak,
                         "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes": st["sptrsv_bytes"], "ms": parts["sptrsv"]},
            "phases_ms": {"factorize_device": parts["fact"], "solve_device": parts["solve"], "sptrsv_sweep": parts["sptrsv"],
                          "residual_spmv": parts["spmv"], "initialize_once_s": t_init,
                          "factorize_tflops": st["flops"] / (parts["fact"] * 1e-3) / 1e12,
                          "spmv_gbs": spmv_gbs, "refine_steps": st["last_refine_steps"]},
            "symbolic": {"fronts": st["nnodes"], "levels": st["nlevels"], "nnz_LU": st["nnz_l"] + st["nnz_u"], "flops": st["flops"],
                         "max_front": st["max_front"]},
            "wall_ms_per_step": wall_dev / K,
        }
        if not args.no_cpu_baseline and world == 1:
            t_cpu, res_cpu = cpu_reference_step(k)
            line["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": "systems/s", "cores": 1, "kind": "port",
                                    "sample": "1 full factorize+solve of the same %dx%d Laplacian with scipy SuperLU (MMD_AT_PLUS_A, "
                                              "sequential; UMFPACK not installed), %.1f s, rel.residual %.1e" % (k, k, t_cpu, res_cpu)}
        print(json.dumps(line), flush=True)
    lib.solver_b200_drop(h)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
