"""Multi-GPU parity check, run under torchrun: the row-partitioned solve (block-Jacobi ILU0,
NCCL halo exchange + all-reduce) against the CPU oracle's unpartitioned solve.
Partitioned parity per BASELINE.json: residual reduction <= linear_solver_reduction, iteration
counts reported side by side."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from opm_simulators_legacy_b200.distributed import DistributedSolver  # noqa: E402
from opm_simulators_legacy_b200.solver import make_params  # noqa: E402

dims = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (24, 20, 16)
singular = "--singular" in sys.argv        # one rank's slab holds a singular pivot: every rank must return, none may hang
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
s = synth_blackoil_jacobian(*dims, perm="lognormal")
g = DistributedSolver(s, local)
if singular:
    # zero the diagonal block of the first row of the LAST rank's slab (row 0 of a slab has no lower
    # blocks inside the slab, so its pivot is the block itself)
    from opm_simulators_legacy_b200.solver import NumericalIssue
    if rank == world - 1:
        # local pattern: the diagonal of local row 0 is the entry whose global column equals the slab's first row
        import numpy as _np
        from opm_simulators_legacy_b200.distributed import local_rows_permuted, local_rows
        g_rp, g_ci, g_v, g_b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
        rp_l, cg_l, _, _ = (local_rows(g_rp, g_ci, g_v, g_b, g.lo, g.hi) if g.axis == 2
                            else local_rows_permuted(g_rp, g_ci, g_v, g_b, g.perm, g.lo, g.hi))
        k = int(_np.nonzero(cg_l[rp_l[0]:rp_l[1]] == g.lo)[0][0])
        g.vals[k] = 0.0
    status = "no exception"
    try:
        g.solve(make_params())
    except NumericalIssue as e:
        status = "NumericalIssue"
    codes = [None] * world
    dist.all_gather_object(codes, status)
    if rank == 0:
        ok = all(c == "NumericalIssue" for c in codes)
        print(json.dumps({"singular_slab": codes, "ok": ok}))
        if not ok:
            sys.exit(1)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
# distributed SpMV against the global one
xs = s.xstar.numpy()
g.set_values_dev(g.vals)
y_loc = g.spmv(xs[g.perm[g.lo:g.hi]])
out = {}
for red in (1e-2, 1e-8):
    res = g.solve(make_params(linear_solver_reduction=red, linear_solver_maxiter=400))
    parts = [torch.zeros((int(g.offsets[r + 1] - g.offsets[r]), 3), dtype=torch.float64, device="cuda") for r in range(world)]
    dist.all_gather(parts, g.x)
    x = g.to_natural(torch.cat(parts).cpu().numpy())
    out[red] = (res, x)
ys = [torch.zeros((int(g.offsets[r + 1] - g.offsets[r]), 3), dtype=torch.float64, device="cuda") for r in range(world)]
dist.all_gather(ys, torch.from_numpy(y_loc).cuda())
if rank == 0:
    from oracle import oracle_py as O
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    y_ref = O.spmv(rp, ci, v, xs)
    y = g.to_natural(torch.cat(ys).cpu().numpy())
    report = {"dims": dims, "world": world, "slab_axis": "ijk"[g.axis], "spmv_bit_exact": bool(np.array_equal(y, y_ref))}
    for red, (res, x) in out.items():
        r = b - O.spmv(rp, ci, v, x)
        x_ref, ref = O.solve_bcrs(rp, ci, v, b, reduction=red, maxiter=400)
        report[f"red_{red:g}"] = {"iterations_partitioned": res["iterations"], "iterations_oracle_unpartitioned": ref["iterations"],
                                  "true_residual_reduction": float(np.linalg.norm(r) / np.linalg.norm(b)),
                                  "reported_reduction": res["reduction"], "converged": res["converged"],
                                  "max_rel_diff_vs_unpartitioned": float((np.abs(x - x_ref).max(0) / np.abs(x_ref).max(0)).max())}
    ok = report["spmv_bit_exact"] and all(report[f"red_{red:g}"]["true_residual_reduction"] < red * 1.0001 for red in out)
    report["ok"] = bool(ok)
    print(json.dumps(report))
    if not ok:
        sys.exit(1)
dist.barrier()
dist.destroy_process_group()
