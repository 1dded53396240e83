"""Small driver for profiling the ILU0 factorisation: a few factorisations of one matrix.
Usage: python tools/factor_only.py NX NY NZ [nfactor]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from opm_simulators_legacy_b200.solver import GpuLinearSolver  # noqa: E402

nx, ny, nz = (int(a) for a in sys.argv[1:4])
nf = int(sys.argv[4]) if len(sys.argv) > 4 else 4
s = synth_blackoil_jacobian(nx, ny, nz, perm="lognormal")
g = GpuLinearSolver(0)
g.set_pattern(s.rowptr.numpy(), s.colidx.numpy())
vals = s.vals.cuda()
g.set_values_dev(vals)
for _ in range(nf):
    assert g.ilu0_factor() == -1
torch.cuda.synchronize()
print("ok")
