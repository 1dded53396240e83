"""C5 sweep beyond 2*10^7 cells (BASELINE.json config 5: 1e4 ... 1e8): the matrix is generated ON THE
DEVICE (a homogeneous 7-point block stencil with block-diagonally dominant 3x3 blocks; SpMV and sweep
times do not depend on the values), the host only sees the index arrays.
    python tools/c5_large.py N [spmv]      N = cells per axis; "spmv": operator only (no ILU0 analysis)
Prints one JSON line like tools/microbench.py."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.solver import GpuLinearSolver  # noqa: E402


def device_stencil(n):
    """BCRS of the 7-point stencil on an n^3 grid, natural ordering, built with torch ops on the GPU."""
    dev = "cuda"
    N = n * n * n
    idx = torch.arange(N, device=dev, dtype=torch.int64)
    i = idx % n; j = (idx // n) % n; k = idx // (n * n)
    offs = [-n * n, -n, -1, 0, 1, n, n * n]
    ok = [k > 0, j > 0, i > 0, torch.ones_like(i, dtype=torch.bool), i < n - 1, j < n - 1, k < n - 1]
    del i, j, k
    mask = torch.stack(ok, dim=1)                                   # [N, 7]
    del ok
    cnt = mask.sum(dim=1)
    rowptr = torch.zeros(N + 1, device=dev, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(cnt, 0)
    nlow = mask[:, :3].sum(dim=1)                                   # blocks before the diagonal
    diag_slot = rowptr[:-1] + nlow
    del cnt, nlow
    cols = (idx[:, None] + torch.tensor(offs, device=dev, dtype=torch.int64)[None, :])[mask].to(torch.int32)
    del mask, idx
    nnzb = int(rowptr[-1])
    off = torch.tensor([[-1.0, -0.05, 0.02], [0.03, -1.0, -0.04], [-0.02, 0.05, -0.5]], device=dev, dtype=torch.float64).reshape(9)
    dia = torch.tensor([[6.5, 0.2, -0.1], [0.1, 6.8, 0.3], [-0.2, 0.1, 3.6]], device=dev, dtype=torch.float64).reshape(9)
    vals = off.repeat(nnzb, 1)
    vals[diag_slot] = dia
    del diag_slot
    return N, nnzb, rowptr.to(torch.int32), cols, vals


def timed(fn, reps, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.add_(1.0)
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    n = int(sys.argv[1])
    spmv_only = len(sys.argv) > 2 and sys.argv[2] == "spmv"
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    t0 = time.time()
    N, nnzb, rowptr, cols, vals = device_stencil(n)
    torch.cuda.synchronize()
    tgen = time.time() - t0
    g = GpuLinearSolver(0)
    g.use_torch_stream()
    rp, ci = rowptr.cpu().numpy(), cols.cpu().numpy()
    del rowptr, cols
    torch.cuda.empty_cache()
    t0 = time.time()
    if spmv_only:
        g.set_pattern_operator_only(rp, ci)
    else:
        g.set_pattern(rp, ci)
    tan = time.time() - t0
    del rp, ci
    x = torch.rand(N, 3, device="cuda", dtype=torch.float64); y = torch.zeros_like(x)
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    g.set_values_dev(vals)
    out = {"dims": [n, n, n], "N": N, "nnzb": nnzb, "gen_s_on_device": tgen, "analysis_s": tan, "operator_only": spmv_only}
    b_spmv = 76 * nnzb + 52 * N
    b_ilu = 76 * (nnzb - N) + 176 * N
    for _ in range(3):
        g.spmv_dev(x, y)
    med, best = timed(lambda: g.spmv_dev(x, y), 11, flush)
    out["spmv_us"] = med * 1e3; out["spmv_gbs"] = b_spmv / med / 1e6; out["spmv_best_gbs"] = b_spmv / best / 1e6
    if not spmv_only:
        out["levels"] = g.num_levels()
        t0 = time.time(); bad = g.ilu0_factor(); torch.cuda.synchronize(); out["factor_first_ms"] = (time.time() - t0) * 1e3
        assert bad == -1
        med, best = timed(lambda: g.ilu0_factor(), 3, flush)
        out["factor_ms"] = med
        for _ in range(2):
            g.ilu0_apply_dev(0.9, x, y)
        med, best = timed(lambda: g.ilu0_apply_dev(0.9, x, y), 7, flush)
        out["ilu_apply_us"] = med * 1e3; out["ilu_apply_gbs"] = b_ilu / med / 1e6; out["ilu_apply_best_gbs"] = b_ilu / best / 1e6
    out["hbm_gb_in_use"] = torch.cuda.memory_allocated() / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
