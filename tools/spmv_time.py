"""Times the SpMV of the double and the float instance.  Usage: spmv_time.py NX NY NZ"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from opm_simulators_legacy_b200.solver import GpuLinearSolver  # noqa: E402

nx, ny, nz = (int(a) for a in sys.argv[1:4])
s = synth_blackoil_jacobian(nx, ny, nz, perm="lognormal")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for single in (False, True):
    g = GpuLinearSolver(0)
    g.use_torch_stream()
    g.set_precision(single)
    g.set_pattern_operator_only(s.rowptr.numpy(), s.colidx.numpy())
    vals = s.vals.cuda(); x = s.rhs.cuda(); y = torch.zeros_like(x)
    g.set_values_dev(vals)
    for _ in range(5):
        g.spmv_dev(x, y)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            g.spmv_dev(x, y)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 50 * 1e3)
    nnzb, N = s.nnzb, s.N
    B = (40 * nnzb + 28 * N) if single else (76 * nnzb + 52 * N)
    conv = " (incl. two conversion kernels double <-> float at the ABI)" if single else ""
    print(f"{nx}x{ny}x{nz} {'f32' if single else 'f64'} ROWT={os.environ.get('OPMGPU_SPMV_ROWT', 'default')}: SpMV {min(ts):.1f} us{conv} = {B / min(ts) / 1e3:.0f} GB/s")
    g.close()
