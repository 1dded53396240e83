#!/usr/bin/env python
"""Benchmark of the Newton-step linear solve (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c3|c2|c4|small]

One "step" = one pass of the hot path: ILU0 factorisation + ILU0/BiCGStab solve to
linear_solver_reduction = 1e-2 of one synthetic three-phase black-oil Jacobian system
(SURVEY.md §8d), i.e. what NewtonIterationBlackoilInterleaved::computeNewtonIncrement does
between formInterleavedSystem and the de-interleave.

* `value` / `ms_per_step`: inputs resident in HBM, opmgpu_solve_bcrs3_dev, CUDA events on the
  launching stream, max over ranks.
* `e2e`: the same solve through the adapter-facing C-ABI call opmgpu_solve_from_csc_blocks with
  HOST buffers (nine CSC value arrays + residual in pinned memory in, increment out), host<->device
  copies and the device-side interleave inside the timed region.
* `roofline`: the dominant kernel class (the ILU0 apply = pipelined lower+upper sweeps), timed
  live with CUDA events around every apply inside the timed solves.
* `cpu_baseline`: the CPU oracle (a port of the reference's dune-istl path, 1 thread like the
  reference) on the same system, on rank 0.
* `--impl reference`: the CPU oracle alone, same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dims, permeability)
    "small": ((24, 20, 12), "lognormal"),
    "c2": ((100, 100, 50), "homogeneous"),
    "c3": ((100, 100, 100), "lognormal"),     # 1M cells: the configuration the metric is quoted on
    "c4": ((200, 200, 200), "lognormal"),     # 8M cells: strong-scaling configuration
}
METRIC = "linear_solve_ms_per_newton_step"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML (in-process, every 2 ms) during the
    timed regions; falls back to one nvidia-smi query when NVML cannot be loaded."""

    def __init__(self, index=0):
        self.index = index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._nv = None

    def _phys_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._nv = nv
            self._h = nv.nvmlDeviceGetHandleByIndex(self._phys_index())
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM))
        except Exception:
            self._nv = None
            return
        self._sample()
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def _sample(self):
        nv = self._nv
        try:
            self.sm.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                              ("hw_thermal_slowdown", 0x40)):
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _loop(self):
        while not self._stop.wait(0.002):
            self._sample()

    def stop(self):
        if self._nv is None:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
                a, b = (int(v) for v in out.strip().split(","))
                return {"sm_mhz": a, "sm_max_mhz": b, "samples": 1, "reasons": [], "how": "nvidia-smi after the timed region (NVML unavailable)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        self._sample()
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "reasons": sorted(self.reasons), "how": "NVML, every 2 ms during the timed region"}


def build_system(workload):
    from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian
    dims, perm = WORKLOADS[workload]
    return synth_blackoil_jacobian(*dims, perm=perm)


def config_of(workload, s, n_gpus):
    dims, perm = WORKLOADS[workload]
    return {"workload": f"{workload}: synthetic 3-phase black-oil Jacobian, {dims[0]}x{dims[1]}x{dims[2]} "
                        f"Cartesian 7-point stencil, {perm} permeability, 3x3 BCRS",
            "cells": s.N, "nnzb": s.nnzb, "linear_solver_reduction": 1e-2, "linear_solver_maxiter": 150,
            "ilu_relaxation": 0.9, "partition": "single GPU" if n_gpus == 1 else f"{n_gpus} slabs along the weakest-coupling axis, block-Jacobi ILU0",
            "l2": "inputs (>= 0.5 GB matrix) larger than the 126 MB L2; no explicit flush"}


def run_reference(args, rank):
    """CPU arm: the oracle (port of the reference's dune-istl path; the reference itself cannot be
    built here), single-threaded like the reference's solver."""
    if rank != 0:
        return
    from oracle import oracle_py as O
    s = build_system(args.workload)
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    res = None
    for _ in range(args.warmup):
        _, res = O.solve_bcrs(rp, ci, v, b)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        _, res = O.solve_bcrs(rp, ci, v, b)
        ts.append((time.perf_counter() - t0) * 1e3)
    ms = sum(ts) / len(ts)
    line = {"impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args.workload, s, 1), "iterations": res["iterations"],
            "cpu_baseline": {"value": ms, "unit": "ms", "cores": 1, "kind": "port",
                             "sample": f"{args.steps} full solves of the same system (ILU0 factor + BiCGStab to 1e-2)"},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_gpu(args, rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from opm_simulators_legacy_b200.solver import GpuLinearSolver, make_params

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from opm_simulators_legacy_b200.distributed import DistributedSolver
    s = build_system(args.workload)
    params = make_params()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    if world == 1:
        g = GpuLinearSolver(local)
        g.use_torch_stream()
        t0 = time.perf_counter()
        g.set_pattern(s.rowptr.numpy(), s.colidx.numpy())
        analysis_ms = (time.perf_counter() - t0) * 1e3
        vals = s.vals.cuda()
        rhs = s.rhs.cuda()
        x = torch.zeros_like(rhs)
        solve = lambda: g.solve_bcrs_dev(vals, rhs, x, params=params)          # noqa: E731
    else:
        g = DistributedSolver(s, local)
        analysis_ms = g.analysis_ms
        x = None
        solve = lambda: g.solve(params)                                          # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = None
    for _ in range(args.warmup):
        res = solve()
    g.set_profiling(True)
    l0 = g.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        res = solve()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms = float(ms_total) / args.steps
    launches = g.launch_count() - l0
    prof = g.profile()
    g.set_profiling(False)

    # ---- end to end through the C-ABI call with host buffers (rank-local system)
    e2e = None
    if world > 1:
        vals_h = g.vals.cpu().pin_memory().numpy()
        rhs_h = g.rhs.cpu().pin_memory().numpy()
        for _ in range(2):
            g.solve_bcrs(vals_h, rhs_h, params=params)
        barrier()
        ev0.record()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            g.solve_bcrs(vals_h, rhs_h, params=params)
        ev1.record()
        barrier()
        t = torch.tensor([max((time.perf_counter() - t0) * 1e3, ev0.elapsed_time(ev1)) / args.steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": float(t), "unit": "ms", "h2d_bytes_per_step": int(vals_h.nbytes + rhs_h.nbytes),
               "d2h_bytes_per_step": int(rhs_h.nbytes), "call": "opmgpu_solve_bcrs3 (pinned host buffers, this rank's rows; bytes are per rank)"}
    if world == 1:
        blocks = s.csc_blocks()
        pinned = []
        for cp, ri, v in blocks:
            t = torch.from_numpy(v).pin_memory()
            pinned.append((cp, ri, t.numpy()))
        rhs_h = s.rhs_eqmajor_unscaled.clone().pin_memory().numpy()
        dx_h = torch.empty(3 * s.N, dtype=torch.float64).pin_memory().numpy()      # page-locked result buffer
        h2d = sum(b[2].nbytes for b in pinned) + rhs_h.nbytes
        d2h = rhs_h.nbytes
        g.solve_from_csc_blocks(s.N, pinned, s.matbalscale, rhs_h, params=params, out=dx_h)       # pattern analysis, untimed
        for _ in range(max(1, args.warmup // 2)):
            g.solve_from_csc_blocks(s.N, pinned, s.matbalscale, rhs_h, params=params, out=dx_h)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(args.steps):
            dx, r2 = g.solve_from_csc_blocks(s.N, pinned, s.matbalscale, rhs_h, params=params, out=dx_h)
        ev1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / args.steps
        e2e = {"value": max(wall, ev0.elapsed_time(ev1) / args.steps), "unit": "ms", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "call": "opmgpu_solve_from_csc_blocks (pinned host buffers)",
               "breakdown_ms": {k: r2[k] for k in ("ms_h2d", "ms_interleave", "ms_factor", "ms_solve", "ms_d2h")}}

    if rank != 0:
        return
    # ---- roofline of the dominant kernel class: ILU0 apply (lower + upper sweep)
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        peak = json.load(f).get("hbm_gbs") if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
    peak_src = "measured (MEASURED_PEAKS.json)"
    if not peak:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # per launch = per rank: rank 0's rows (its diagonal block for the block-Jacobi ILU0)
    N, nnzb = (s.N, s.nnzb) if world == 1 else (g.N, g.nnzb)
    nnzb_ilu = nnzb if world == 1 else g.nnzb_diag      # blocks of the rank's diagonal block (block-Jacobi ILU0)
    b_ilu = 76 * (nnzb_ilu - N) + 176 * N       # SURVEY.md §8d, bytes per apply
    b_spmv = 76 * nnzb + 52 * N
    ap_ms, ap_n = prof["ilu_apply"]
    sp_ms, sp_n = prof["spmv"]
    ach = b_ilu * ap_n / (ap_ms * 1e-3) / 1e9 if ap_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get("ilu_apply_dram_bytes")
    roofline = {"kernel": "ILU0 apply = ilu0_sweep_pipe_kernel<lower> + <upper> (right-hand side permuted by the producing vector kernel)", "bound": "hbm",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                "traffic": traffic, "algorithmic_bytes_per_launch": b_ilu, "launches_timed": ap_n,
                "avg_launch_us": ap_ms * 1e3 / max(ap_n, 1),
                "share_of_step": ap_ms / (ms * args.steps),
                "spmv": {"achieved": b_spmv * sp_n / (sp_ms * 1e-3) / 1e9 if sp_ms > 0 else 0.0,
                         "avg_launch_us": sp_ms * 1e3 / max(sp_n, 1), "algorithmic_bytes_per_launch": b_spmv,
                         "share_of_step": sp_ms / (ms * args.steps)},
                "factor_share_of_step": prof["factor"][0] / (ms * args.steps),
                "vector_share_of_step": prof["vector"][0] / (ms * args.steps)}

    # ---- CPU baseline on this box's host cores (bounded sample: full solves of the same system)
    from oracle import oracle_py as O
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    t0 = time.perf_counter()
    x_ref, ref = O.solve_bcrs(rp, ci, v, b)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    parity = None
    if x is not None:
        xg = x.cpu().numpy()
        parity = {"iterations_gpu": res["iterations"], "iterations_cpu_oracle": ref["iterations"],
                  "max_rel_diff_increment": float((np.abs(xg - x_ref).max(0) / np.abs(x_ref).max(0)).max())}
    line = {"metric": METRIC, "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_of(args.workload, s, world), "clocks": clocks,
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": {"value": cpu_ms, "unit": "ms", "cores": 1, "kind": "port",
                             "sample": "1 full solve of the same system (ILU0 factor + BiCGStab to 1e-2), "
                                       "single thread like the reference's sequential dune-istl solver",
                             "host_cpus": os.cpu_count()},
            "iterations": res["iterations"], "reduction": res["reduction"], "analysis_ms_once_per_pattern": analysis_ms,
            "solve_breakdown_ms": {"factor": res["ms_factor"], "bicgstab": res["ms_solve"]}, "parity": parity}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_gpu(args, rank, world)


if __name__ == "__main__":
    main()
