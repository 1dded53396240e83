"""Kernel-level timings (CUDA events on the launching stream) for SpMV, ILU0 apply, ILU0
factor and the full solve on a synthetic Cartesian black-oil Jacobian.
Usage: python tools/microbench.py NX NY NZ [perm] [reps]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from opm_simulators_legacy_b200.solver import GpuLinearSolver, make_params  # noqa: E402


def timed(fn, reps, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.add_(1.0)          # > L2: evict
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    nx, ny, nz = (int(a) for a in sys.argv[1:4])
    perm = sys.argv[4] if len(sys.argv) > 4 else "lognormal"
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
    t0 = time.time()
    s = synth_blackoil_jacobian(nx, ny, nz, perm=perm)
    tgen = time.time() - t0
    g = GpuLinearSolver(0)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    g.use_torch_stream()
    if os.environ.get("MB_MULTICOLOUR"):          # the flagged multicolour-ILU0 variant
        g.set_ilu_ordering("lines" if os.environ["MB_MULTICOLOUR"] == "lines" else True)
    t0 = time.time()
    g.set_pattern(s.rowptr.numpy(), s.colidx.numpy())
    tan = time.time() - t0
    vals = s.vals.cuda(); rhs = s.rhs.cuda(); x = torch.zeros_like(rhs); y = torch.zeros_like(rhs)
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")     # 256 MB
    g.set_values_dev(vals)
    N, nnzb = s.N, s.nnzb
    b_spmv = 76 * nnzb + 52 * N
    b_ilu = 76 * (nnzb - N) + 176 * N
    out = {"dims": [nx, ny, nz], "N": N, "nnzb": nnzb, "levels": g.num_levels(), "gen_s": tgen, "analysis_s": tan}
    for _ in range(3):
        g.spmv_dev(rhs, y)
    med, best = timed(lambda: g.spmv_dev(rhs, y), reps, flush)
    out["spmv_us"] = med * 1e3; out["spmv_gbs"] = b_spmv / med / 1e6; out["spmv_best_gbs"] = b_spmv / best / 1e6
    t0 = time.time(); bad = g.ilu0_factor(); torch.cuda.synchronize(); out["factor_first_ms"] = (time.time() - t0) * 1e3
    assert bad == -1
    med, best = timed(lambda: g.ilu0_factor(), max(3, reps // 4), flush)
    out["factor_ms"] = med
    for _ in range(3):
        g.ilu0_apply_dev(0.9, rhs, y)
    med, best = timed(lambda: g.ilu0_apply_dev(0.9, rhs, y), reps, flush)
    out["ilu_apply_us"] = med * 1e3; out["ilu_apply_gbs"] = b_ilu / med / 1e6; out["ilu_apply_best_gbs"] = b_ilu / best / 1e6
    if os.environ.get("MB_NOSOLVE"):
        print(json.dumps(out)); return
    p = make_params()
    for _ in range(2):
        res = g.solve_bcrs_dev(vals, rhs, x, params=p)
    ts = []
    for _ in range(5):
        flush.add_(1.0)
        torch.cuda.synchronize(); t0 = time.time()
        res = g.solve_bcrs_dev(vals, rhs, x, params=p)
        torch.cuda.synchronize(); ts.append((time.time() - t0) * 1e3)
    out["solve_ms"] = sorted(ts)[len(ts) // 2]
    out["solve"] = {k: res[k] for k in ("iterations", "half_steps", "reduction", "ms_factor", "ms_solve")}
    out["launches"] = g.launch_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
