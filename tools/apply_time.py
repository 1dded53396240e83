"""Times the ILU0 apply (both sweeps) and the SpMV with CUDA events.  Usage: apply_time.py NX NY NZ [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from opm_simulators_legacy_b200.solver import GpuLinearSolver  # noqa: E402

nx, ny, nz = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 50
s = synth_blackoil_jacobian(nx, ny, nz, perm="lognormal")
g = GpuLinearSolver(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); g.use_torch_stream()
g.set_pattern(s.rowptr.numpy(), s.colidx.numpy())
vals = s.vals.cuda(); rhs = s.rhs.cuda(); y = torch.zeros_like(rhs)
g.set_values_dev(vals)
assert g.ilu0_factor() == -1
for _ in range(5):
    g.ilu0_apply_dev(0.9, rhs, y)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.ilu0_apply_dev(0.9, rhs, y)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / reps * 1e3)
nnzb = s.colidx.numel(); N = s.N
B = 76 * (nnzb - N) + 176 * N
t = min(ts)
print(f"{nx}x{ny}x{nz}: ILU0 apply (incl. permute kernel) {t:.1f} us (runs {' '.join(f'{v:.1f}' for v in ts)}) = {B / t / 1e3:.0f} GB/s")
