"""Debug: per-step clock64 trace of one CTA of the pipelined ILU0 sweeps."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from opm_simulators_legacy_b200.solver import GpuLinearSolver  # noqa: E402

nx, ny, nz = (int(a) for a in sys.argv[1:4])
ctas = [int(a) for a in sys.argv[4:]] or [0, 70, 143]
s = synth_blackoil_jacobian(nx, ny, nz, perm="lognormal")
g = GpuLinearSolver(0)
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
        if n < 3:
            print(f"cta {cta} {name}: {n} steps"); continue
        t0 = t[0, 0]
        kentry = t[511, 15]
        print(f"cta {cta} {name}: first step chain done {t[0,2]-kentry} cyc after kernel entry; step 8 done {t[8,2]-kentry}; step 16 done {t[16,2]-kentry}; last {t[n-1,2]-kentry}")
        wait = (t[:n, 1] - t[:n, 0]); comp = (t[:n, 2] - t[:n, 1]); bar = (t[:n, 3] - t[:n, 2])
        step = np.diff(t[:n, 0])
        print(f"cta {cta} {name}: steps {n} total {t[n-1,3]-t0} cyc; per-step median: period {np.median(step):.0f} "
              f"wait {np.median(wait):.0f} compute {np.median(comp):.0f} barrier {np.median(bar):.0f}; "
              f"mean wait {wait.mean():.0f} comp {comp.mean():.0f} bar {bar.mean():.0f}")
        own = t[2:n, 5] - t[:n-2, 3]     # this group's: end of stores (step s-2) -> start of step s
        print("   ping-pong (median cycles): own loop overhead %d | full-wait+static load %d | barrier wait %d | chain %d | stores %d | step-to-step (same group) %d"
              % (np.median(own), np.median(t[:n,0]-t[:n,5]), np.median(t[:n,1]-t[:n,0]), np.median(t[:n,2]-t[:n,1]), np.median(t[:n,3]-t[:n,2]), np.median(t[2:n,5]-t[:n-2,5])))
        idx = list(range(0, n, max(1, n // 12)))
        print("   step  nrows  enter   wait  comp  bar | bulk_issue-enter  gather_issue-enter")
        for i in idx:
            print(f"   {i:4d} {t[i,4]:5d} {t[i,0]-t0:8d} {wait[i]:6d} {comp[i]:5d} {bar[i]:4d} | {t[i,5]-t[i,0]:8d} {t[i,6]-t[i,0]:8d}")
        print(f"   helper: loop start {t[509,8]-kentry}, first delivery at {t[509,9]-kentry} after {t[509,10]} polls, n={t[509,11]}; kentry {kentry}")
        print("   raw (rel. kernel entry): step: looptop static_done bar_issued chain_done stores_done")
        for i in range(0, min(n, 6)):
            print(f"      {i}: {t[i,5]-kentry} {t[i,0]-kentry} {t[i,1]-kentry} {t[i,2]-kentry} {t[i,3]-kentry}")
