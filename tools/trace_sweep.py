"""Debug: per-step clock64 trace of one CTA of the pipelined ILU0 sweeps.
Stamps per step: [0] static data in registers, [1] turn barrier passed, [5] pushed inputs ready,
[2] chain done, [3] stores issued, [4] rows."""
import ctypes as C
import os
import sys

import numpy as np

os.environ.setdefault("OPMGPU_CLUSTER", "0")      # the traced kernel variants exist without clusters only
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from opm_simulators_legacy_b200.solver import GpuLinearSolver  # noqa: E402

G = 3
nx, ny, nz = (int(a) for a in sys.argv[1:4])
ctas = [int(a) for a in sys.argv[4:]] or [0, 70, 142]
s = synth_blackoil_jacobian(nx, ny, nz, perm="lognormal")
g = GpuLinearSolver(0, experiments=True)          # tracing entry points: experiments build
g.set_pattern(s.rowptr.numpy(), s.colidx.numpy())
vals = s.vals.cuda(); rhs = s.rhs.cuda(); y = torch.zeros_like(rhs)
g.set_values_dev(vals)
assert g.ilu0_factor() == -1
for _ in range(3):
    g.ilu0_apply_dev(0.9, rhs, y)
torch.cuda.synchronize()
f = g.lib.opmgpu_debug_trace_apply
f.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
f.restype = C.c_int
for cta in ctas:
    out = np.zeros((2, 512, 16), dtype=np.int64)
    rc = f(g.h, cta, 0.9, C.c_void_p(rhs.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(out.ctypes.data))
    assert rc == 0, g.error()
    for sw, name in ((0, "L"), (1, "U")):
        t = out[sw]
        n = int((t[:500, 0] > 0).sum())
        if n < 8:
            print(f"cta {cta} {name}: {n} steps"); continue
        med = lambda v: float(np.median(v))
        print(f"cta {cta} {name}: {n} steps, total {t[n-1,2]-t[0,0]} cycles; median cycles per step: "
              f"prev chain done -> turn barrier passed {med(t[1:n,1]-t[:n-1,2]):.0f} | barrier -> pushed inputs ok {med(t[1:n,5]-t[1:n,1]):.0f} | "
              f"chain {med(t[1:n,2]-t[1:n,5]):.0f} | chain done -> chain done {med(np.diff(t[:n,2])):.0f} || off path: stores {med(t[:n,3]-t[:n,2]):.0f} | "
              f"stores -> next static loaded {med(t[G:n,0]-t[:n-G,3]):.0f} | static loaded -> barrier passed (idle) {med(t[1:n,1]-t[1:n,0]):.0f}")
        idx = list(range(0, n, max(1, n // 12)))
        print("   step rows | prev-done->bar  bar->ext  chain  stores")
        for i in idx[1:]:
            print(f"   {i:4d} {t[i,4]:4d} | {t[i,1]-t[i-1,2]:6d} {t[i,5]-t[i,1]:6d} {t[i,2]-t[i,5]:6d} {t[i,3]-t[i,2]:6d}")
