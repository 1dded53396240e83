"""Small driver for profiling: factor once, then a few ILU0 applies and SpMVs.
Usage: python tools/apply_only.py NX NY NZ [napply] [f32]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from opm_simulators_legacy_b200.solver import GpuLinearSolver  # noqa: E402

nx, ny, nz = (int(a) for a in sys.argv[1:4])
napply = int(sys.argv[4]) if len(sys.argv) > 4 else 4
single = len(sys.argv) > 5 and sys.argv[5] == "f32"
s = synth_blackoil_jacobian(nx, ny, nz, perm="lognormal")
g = GpuLinearSolver(0)
g.set_precision(single)
if os.environ.get("MB_MULTICOLOUR"):          # the flagged multicolour-ILU0 variant
    g.set_ilu_ordering("lines" if os.environ["MB_MULTICOLOUR"] == "lines" else True)
g.set_pattern(s.rowptr.numpy(), s.colidx.numpy())
vals = s.vals.cuda(); rhs = s.rhs.cuda(); y = torch.zeros_like(rhs)
g.set_values_dev(vals)
assert g.ilu0_factor() == -1
for _ in range(napply):
    g.ilu0_apply_dev(0.9, rhs, y)
    g.spmv_dev(rhs, y)
torch.cuda.synchronize()
print("ok", float(y.abs().sum()))
