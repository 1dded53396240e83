#!/usr/bin/env python
"""Benchmark of the Newton-step linear solve (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c3|c2|c4|small] [--dtype f64|f32]

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
* `--dtype f32`: the whole line for the single-precision instance (the reference's Impl<3,float>,
  selected by LinearisedBlackoilResidual::singlePrecision), checked against the float build of the
  oracle.  The default (f64) line carries a compact `f32` sub-measurement of the same workload.
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


def build_system(workload, rank=0, world=1):
    """The synthetic system of a workload.  Multi-GPU runs build it once, on rank 0 (CPU, so every
    rank sees bit-identical inputs), and broadcast the arrays through NCCL: a rank never holds the
    generator's temporaries (15 GB at 8M cells)."""
    from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, SynthSystem
    dims, perm = WORKLOADS[workload]
    if world == 1:
        return synth_blackoil_jacobian(*dims, perm=perm)
    import torch
    import torch.distributed as dist
    s = synth_blackoil_jacobian(*dims, perm=perm) if rank == 0 else None
    N = dims[0] * dims[1] * dims[2]
    nnzb = 7 * N - 2 * (dims[0] * dims[1] + dims[1] * dims[2] + dims[0] * dims[2])
    spec = [("rowptr", (N + 1,), torch.int32), ("colidx", (nnzb,), torch.int32), ("vals_unscaled", (nnzb, 9), torch.float64),
            ("sat_present", (nnzb,), torch.bool), ("rhs_unscaled", (N, 3), torch.float64), ("xstar", (N, 3), torch.float64)]
    got = {}
    for name, shape, dt in spec:
        t = getattr(s, name).cuda() if rank == 0 else torch.empty(shape, dtype=dt, device="cuda")
        if dt == torch.bool:
            t = t.to(torch.uint8)
        dist.broadcast(t, 0)
        got[name] = (t.to(torch.bool) if dt == torch.bool else t).cpu()
        del t
    torch.cuda.empty_cache()
    if rank == 0:
        return s
    return SynthSystem(dims, N, nnzb, got["rowptr"], got["colidx"], got["vals_unscaled"], got["sat_present"],
                       got["rhs_unscaled"], got["xstar"])


def config_of(workload, s, n_gpus, axis=None):
    dims, perm = WORKLOADS[workload]
    part = "single GPU" if n_gpus == 1 else (f"{n_gpus} slabs along grid axis {'ijk'[axis] if axis is not None else '?'} "
                                              "(the weakest coupling), block-Jacobi ILU0")
    return {"workload": f"{workload}: synthetic 3-phase black-oil Jacobian, {dims[0]}x{dims[1]}x{dims[2]} "
                        f"Cartesian 7-point stencil, {perm} permeability, 3x3 BCRS",
            "cells": s.N, "nnzb": s.nnzb, "linear_solver_reduction": 1e-2, "linear_solver_maxiter": 150,
            "ilu_relaxation": 0.9, "partition": part,
            "l2": "inputs (>= 0.5 GB matrix) larger than the 126 MB L2; no explicit flush"}


def run_reference(args, rank):
    """CPU arm: the oracle (port of the reference's dune-istl path; the reference itself cannot be
    built here), single-threaded like the reference's solver."""
    if rank != 0:
        return
    from oracle import oracle_py
    O = oracle_py.instance(args.dtype == "f32")
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
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": config_of(args.workload, s, 1), "iterations": res["iterations"],
            "cpu_baseline": {"value": ms, "unit": "ms", "cores": 1, "kind": "port",
                             "sample": f"{args.steps} full solves of the same system (ILU0 factor + BiCGStab to 1e-2)"},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def measure(workload, steps, warmup, rank, world, local, sample_clocks=True, single=False, with_allcores=True):
    """One workload on `world` GPUs: device-timed solve, end-to-end solve, roofline of the ILU0
    apply, parity against the CPU oracle.  single: the float instance (Impl<3,float>) against the
    float oracle.  Returns the pieces of the JSON line (rank 0) or None."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from opm_simulators_legacy_b200.solver import GpuLinearSolver, make_params

    s = build_system(workload, rank, world)
    params = make_params()

    if world == 1:
        g = GpuLinearSolver(local)
        g.use_torch_stream()
        t0 = time.perf_counter()
        g.set_pattern(s.rowptr.numpy(), s.colidx.numpy())
        analysis_ms = (time.perf_counter() - t0) * 1e3
        vals = s.vals.cuda()
        rhs = s.rhs.cuda()
        x = torch.zeros_like(rhs)
        if single:
            # matrix resident in HBM as the instance stores it (float): rounded once, outside the timed region
            g.set_precision(True)
            g.set_values_dev(vals)
            solve = lambda: g.solve_bcrs_dev(None, rhs, x, params=params)      # noqa: E731
        else:
            solve = lambda: g.solve_bcrs_dev(vals, rhs, x, params=params)      # noqa: E731
        axis = None
    else:
        from opm_simulators_legacy_b200.distributed import DistributedSolver
        g = DistributedSolver(s, local)
        analysis_ms = g.analysis_ms
        x = g.x
        if single:
            g.set_precision(True)
            g.set_values_dev(g.vals)
            solve = lambda: g.solve_bcrs_dev(None, g.rhs, g.x, params=params)    # noqa: E731
        else:
            solve = lambda: g.solve(params)                                      # noqa: E731
        axis = g.axis

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # distributed SpMV against the oracle's global one (bit-exact: entries keep their natural order)
    y_nat = None
    if world > 1:
        xs = s.xstar.numpy()
        g.set_values_dev(g.vals)
        y_loc = torch.from_numpy(g.spmv(xs[g.perm[g.lo:g.hi]])).cuda()
        parts = [torch.zeros((int(g.offsets[r + 1] - g.offsets[r]), 3), dtype=torch.float64, device="cuda") for r in range(world)]
        dist.all_gather(parts, y_loc)
        if rank == 0:
            y_nat = g.to_natural(torch.cat(parts).cpu().numpy())
        del parts

    res = None
    for _ in range(warmup):
        res = solve()
    g.set_profiling(True)
    l0 = g.launch_count()
    sampler = ClockSampler(local)
    if rank == 0 and sample_clocks:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        res = solve()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 and sample_clocks else None
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms = float(ms_total) / steps
    launches = g.launch_count() - l0
    prof = g.profile()
    g.set_profiling(False)

    # the increment in natural cell order on rank 0
    if world > 1:
        parts = [torch.zeros((int(g.offsets[r + 1] - g.offsets[r]), 3), dtype=torch.float64, device="cuda") for r in range(world)]
        dist.all_gather(parts, g.x)
        x_nat = g.to_natural(torch.cat(parts).cpu().numpy()) if rank == 0 else None
        del parts
    else:
        x_nat = x.cpu().numpy()

    # ---- end to end through the C-ABI call with host buffers
    e2e = None
    if world > 1:
        vals_h = g.vals.cpu().pin_memory().numpy()
        rhs_h = g.rhs.cpu().pin_memory().numpy()
        for _ in range(2):
            g.solve_bcrs(vals_h, rhs_h, params=params)
        barrier()
        ev0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            g.solve_bcrs(vals_h, rhs_h, params=params)
        ev1.record()
        barrier()
        t = torch.tensor([max((time.perf_counter() - t0) * 1e3, ev0.elapsed_time(ev1)) / steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": float(t), "unit": "ms", "h2d_bytes_per_step": int(vals_h.nbytes + rhs_h.nbytes),
               "d2h_bytes_per_step": int(rhs_h.nbytes), "call": "opmgpu_solve_bcrs3 (pinned host buffers, this rank's rows; bytes are per rank)"}
    else:
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
        for _ in range(max(1, warmup // 2)):
            g.solve_from_csc_blocks(s.N, pinned, s.matbalscale, rhs_h, params=params, out=dx_h)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            dx, r2 = g.solve_from_csc_blocks(s.N, pinned, s.matbalscale, rhs_h, params=params, out=dx_h)
        ev1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / steps
        e2e = {"value": max(wall, ev0.elapsed_time(ev1) / steps), "unit": "ms", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "call": "opmgpu_solve_from_csc_blocks (pinned host buffers)",
               "breakdown_ms": {k: r2[k] for k in ("ms_h2d", "ms_interleave", "ms_factor", "ms_solve", "ms_d2h")}}
        del pinned, blocks

    if rank != 0:
        g.close()
        return None
    # ---- roofline of the dominant kernel class: ILU0 apply (lower + upper sweep)
    peak, peak_src = None, "measured (MEASURED_PEAKS.json)"
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f).get("hbm_gbs")
    if not peak:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # per launch = per rank: rank 0's rows (its diagonal block for the block-Jacobi ILU0)
    N, nnzb = (s.N, s.nnzb) if world == 1 else (g.N, g.nnzb)
    nnzb_ilu = nnzb if world == 1 else g.nnzb_diag      # blocks of the rank's diagonal block (block-Jacobi ILU0)
    b_ilu = 76 * (nnzb_ilu - N) + 176 * N       # SURVEY.md §8d, bytes per apply (the float instance keeps 8-byte factor containers)
    b_spmv = (40 * nnzb + 28 * N) if single else (76 * nnzb + 52 * N)          # 36-byte blocks, 4-byte vector entries in float
    ap_ms, ap_n = prof["ilu_apply"]
    sp_ms, sp_n = prof["spmv"]
    ach = b_ilu * ap_n / (ap_ms * 1e-3) / 1e9 if ap_ms > 0 else 0.0
    # DRAM traffic of one apply from an `ncu --set full` capture: only quoted for the workload and
    # GPU count it was captured on (profiles/r02_ncu_traffic.json), null otherwise
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{workload}_n{world}", {}).get("ilu_apply_dram_bytes")
    roofline = {"kernel": "ILU0 apply = lower + upper sweep kernel (right-hand side permuted by the producing vector kernel)", "bound": "hbm",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                "traffic": traffic, "algorithmic_bytes_per_launch": b_ilu, "launches_timed": ap_n,
                "avg_launch_us": ap_ms * 1e3 / max(ap_n, 1),
                "share_of_step": ap_ms / (ms * steps),
                "spmv": {"achieved": b_spmv * sp_n / (sp_ms * 1e-3) / 1e9 if sp_ms > 0 else 0.0,
                         "avg_launch_us": sp_ms * 1e3 / max(sp_n, 1), "algorithmic_bytes_per_launch": b_spmv,
                         "share_of_step": sp_ms / (ms * steps)},
                "factor_share_of_step": prof["factor"][0] / (ms * steps),
                "vector_share_of_step": prof["vector"][0] / (ms * steps)}

    # ---- CPU baseline on this box's host cores (bounded sample: one full solve of the same system)
    from oracle import oracle_py
    O = oracle_py.instance(single)
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    t0 = time.perf_counter()
    x_ref, ref = O.solve_bcrs(rp, ci, v, b)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    x_ref = x_ref.astype(np.float64)
    true_red = float(np.linalg.norm(b - oracle_py.spmv(rp, ci, v, x_nat)) / np.linalg.norm(b))
    if world == 1:
        parity = {"iterations_gpu": res["iterations"], "iterations_cpu_oracle": ref["iterations"],
                  "max_rel_diff_increment": float((np.abs(x_nat - x_ref).max(0) / np.abs(x_ref).max(0)).max()),
                  "true_residual_reduction": true_red}
    else:
        # partitioned parity (BASELINE.json north_star): the true residual of the gathered increment
        # is reduced by linear_solver_reduction; iteration counts side by side
        parity = {"iterations_partitioned": res["iterations"], "iterations_oracle_unpartitioned": ref["iterations"],
                  "true_residual_reduction": true_red, "tolerance": 1e-2, "within_tolerance": bool(true_red <= 1e-2 * 1.0001),
                  "spmv_bit_exact": bool(np.array_equal(y_nat, O.spmv(rp, ci, v, s.xstar.numpy()).astype(np.float64))),
                  "max_rel_diff_vs_unpartitioned": float((np.abs(x_nat - x_ref).max(0) / np.abs(x_ref).max(0)).max())}
    # Baseline B: the same algorithm with OpenMP on every host core, level-scheduled
    # ILU0 -- labelled, because the reference's solver is sequential (rank 0 only; torchrun pins
    # OMP_NUM_THREADS to 1 per rank, so the thread count is set explicitly)
    allcores = None
    try:
        if single or not with_allcores:
            raise RuntimeError("not measured for this sub-line (the OpenMP baseline exists for the double instance)")
        ncpu = len(os.sched_getaffinity(0))
        x_b, rb = oracle_py.solve_bcrs_openmp(rp, ci, v, b, nthreads=ncpu)          # warm-up (page faults, thread pool)
        x_b, rb = oracle_py.solve_bcrs_openmp(rp, ci, v, b, nthreads=ncpu)
        allcores = {"value": rb["ms_factor"] + rb["ms_solve"], "unit": "ms", "cores": rb["threads"], "kind": "port-openmp",
                    "note": "NOT the reference (its solver is sequential): oracle arithmetic, level-scheduled ILU0 factor / "
                            "sweeps, row-parallel SpMV, OpenMP reductions; level sets excluded like the GPU's pattern analysis",
                    "iterations": rb["iterations"],
                    "max_rel_diff_vs_oracle": float((np.abs(x_b - x_ref).max(0) / np.abs(x_ref).max(0)).max())}
    except Exception as e:      # no OpenMP-capable compiler in the image
        allcores = {"unavailable": str(e)[:200]}
    out = {"ms": ms, "clocks": clocks, "e2e": e2e, "launches": launches, "roofline": roofline, "cpu_baseline_allcores": allcores,
           "cpu_baseline": {"value": cpu_ms, "unit": "ms", "cores": 1, "kind": "port",
                            "sample": "1 full solve of the same system (ILU0 factor + BiCGStab to 1e-2), "
                                      "single thread like the reference's sequential dune-istl solver",
                            "host_cpus": os.cpu_count()},
           "iterations": res["iterations"], "reduction": res["reduction"], "analysis_ms": analysis_ms,
           "breakdown": {"factor": res["ms_factor"], "bicgstab": res["ms_solve"]}, "parity": parity,
           "config": config_of(workload, s, world, axis)}
    g.close()
    return out


def measure_multicolour(workload, steps, warmup, local, lines=False):
    """The flagged multicolour-ILU0 variant on one GPU (OPMGPU_ILU_MULTICOLOUR): a DIFFERENT
    preconditioner (ILU0 of P A P^T), so its iteration count stands beside the natural-order one and is
    checked against the oracle run on the permuted system, never against the reference's count."""
    import numpy as np
    import torch
    from opm_simulators_legacy_b200.solver import GpuLinearSolver, make_params
    from oracle import oracle_py

    s = build_system(workload)
    params = make_params()
    g = GpuLinearSolver(local)
    g.use_torch_stream()
    g.set_ilu_ordering("lines" if lines else True)
    t0 = time.perf_counter()
    g.set_pattern(s.rowptr.numpy(), s.colidx.numpy())
    analysis_ms = (time.perf_counter() - t0) * 1e3
    ncolours, n2p = g.ilu_permutation()
    vals, rhs = s.vals.cuda(), s.rhs.cuda()
    x = torch.zeros_like(rhs)
    for _ in range(warmup):
        res = g.solve_bcrs_dev(vals, rhs, x, params=params)
    g.set_profiling(True)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        res = g.solve_bcrs_dev(vals, rhs, x, params=params)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    prof = g.profile()
    x_nat = x.cpu().numpy()
    g.close()
    peak = 6650.0
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f).get("hbm_gbs") or peak
    b_ilu = 76 * (s.nnzb - s.N) + 176 * s.N
    ap_ms, ap_n = prof["ilu_apply"]
    ach = b_ilu * ap_n / (ap_ms * 1e-3) / 1e9 if ap_ms > 0 else 0.0
    # oracle on P A P^T (rows sorted by colour)
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    rows = np.repeat(np.arange(s.N), np.diff(rp))
    pr, pc = n2p[rows], n2p[ci]
    order = np.lexsort((pc, pr))
    prp = np.zeros(s.N + 1, dtype=np.int32)
    np.cumsum(np.bincount(pr, minlength=s.N), out=prp[1:])
    p2n = np.argsort(n2p)
    xp, ref = oracle_py.solve_bcrs(prp, pc[order].astype(np.int32), np.ascontiguousarray(v[order]), b.reshape(-1, 3)[p2n].reshape(-1))
    x_ref = np.empty((s.N, 3)); x_ref[p2n] = xp
    true_red = float(np.linalg.norm(b - oracle_py.spmv(rp, ci, v, x_nat)) / np.linalg.norm(b))
    return {"value": ms, "unit": "ms", "steps": steps, "colours": ncolours, "iterations": res["iterations"],
            "solve_breakdown_ms": {"factor": res["ms_factor"], "bicgstab": res["ms_solve"]},
            "ilu_apply_us": ap_ms * 1e3 / max(ap_n, 1), "ilu_apply_gbs": ach, "ilu_apply_frac": ach / peak,
            "ilu_apply_algorithmic_bytes": b_ilu, "analysis_ms_once_per_pattern": analysis_ms,
            "parity_vs_oracle_on_permuted_system": {
                "iterations_gpu": res["iterations"], "iterations_cpu_oracle": ref["iterations"],
                "max_rel_diff_increment": float((np.abs(x_nat - x_ref).max(0) / np.abs(x_ref).max(0)).max()),
                "true_residual_reduction": true_red},
            "ordering": "k-lines: red-black over the (i,j) columns, natural order along k" if lines else "red-black points (greedy colouring)",
            "note": "FLAGGED VARIANT, not the reference's preconditioner: ILU0 of the colour-sorted permutation P A P^T; "
                    "its iteration count is not comparable with the natural-order (reference) count above"}


def run_gpu(args, rank, world):
    import torch
    import torch.distributed as dist

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    single = args.dtype == "f32"
    m = measure(args.workload, args.steps, args.warmup, rank, world, local, single=single)
    # the float instance (the reference's default for time steps below 20 days,
    # BlackoilModelBase_impl.hpp:284) beside the double headline
    f32 = None
    if not single and not args.no_f32:
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        mf = measure(args.workload, args.steps, 3, rank, world, local, sample_clocks=False, single=True, with_allcores=False)
        if mf is not None:
            f32 = {"value": mf["ms"], "unit": "ms", "dtype": "f32", "steps": args.steps, "iterations": mf["iterations"],
                   "e2e": mf["e2e"], "solve_breakdown_ms": mf["breakdown"],
                   "ilu_apply_us": mf["roofline"]["avg_launch_us"], "ilu_apply_frac_8byte_containers": mf["roofline"]["frac"],
                   "spmv_us": mf["roofline"]["spmv"]["avg_launch_us"], "spmv_gbs": mf["roofline"]["spmv"]["achieved"],
                   "spmv_algorithmic_bytes": mf["roofline"]["spmv"]["algorithmic_bytes_per_launch"],
                   "cpu_baseline_float_oracle_ms": mf["cpu_baseline"]["value"], "parity_vs_float_oracle": mf["parity"],
                   "note": "Impl<3,float>: matrix values and vectors in float (SpMV 40 B/block + 28 B/row), ILU0 factors and sweep "
                           "records keep 8-byte containers (float arithmetic, bit-exact against the float oracle)"}
    # the multicolour-ILU0 variant (north star: "level-set (or multicolour) scheduling"), flagged
    mc = mc_lines = None
    if world == 1 and not single and not args.no_multicolour:
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        mc = measure_multicolour(args.workload, args.steps, 3, local)
        gc.collect()
        torch.cuda.empty_cache()
        mc_lines = measure_multicolour(args.workload, args.steps, 3, local, lines=True)
    # the strong-scaling configuration (BASELINE.json config 4: 8M cells) beside the headline
    c4 = None
    if args.workload == "c3" and not args.no_c4:
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        k4 = max(2, min(args.steps, 3))
        m4 = measure("c4", k4, 3, rank, world, local, sample_clocks=False, single=single)
        if m4 is not None:
            c4 = {"value": m4["ms"], "unit": "ms", "steps": k4, "warmup": 3, "iterations": m4["iterations"],
                  "e2e": m4["e2e"], "ilu_apply_frac": m4["roofline"]["frac"], "ilu_apply_us": m4["roofline"]["avg_launch_us"],
                  "spmv_gbs": m4["roofline"]["spmv"]["achieved"], "cpu_baseline_ms": m4["cpu_baseline"]["value"],
                  "cpu_baseline_allcores": m4["cpu_baseline_allcores"],
                  "analysis_ms_once_per_pattern": m4["analysis_ms"], "parity": m4["parity"],
                  "config": m4["config"]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {"metric": METRIC, "value": m["ms"], "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["ms"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "config": m["config"], "clocks": m["clocks"],
            "e2e": m["e2e"], "gpu_launches": m["launches"], "roofline": m["roofline"],
            "cpu_baseline": m["cpu_baseline"], "cpu_baseline_allcores": m["cpu_baseline_allcores"],
            "iterations": m["iterations"], "reduction": m["reduction"], "analysis_ms_once_per_pattern": m["analysis_ms"],
            "solve_breakdown_ms": m["breakdown"], "parity": m["parity"], "c4": c4, "f32": f32,
            "multicolour_variant": mc, "multicolour_lines_variant": mc_lines}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-c4", action="store_true", help="skip the 8M-cell sub-measurement of the default (c3) run")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="instance measured on the headline: Impl<3,double> (default) or Impl<3,float> (singlePrecision)")
    ap.add_argument("--no-multicolour", action="store_true", help="skip the sub-measurement of the flagged multicolour-ILU0 variant")
    ap.add_argument("--no-f32", action="store_true", help="skip the float-instance sub-measurement of the default (f64) run")
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
