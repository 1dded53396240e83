"""Debug: %globaltimer trace of every CTA of the pipelined ILU0 sweeps as launched in production
(thread-block clusters).  Prints, per CTA in order of its first step: start, steps, median step
period, median barrier wait / pushed-input wait / chain (the chain includes waits for results
delivered through distributed shared memory)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from opm_simulators_legacy_b200.solver import GpuLinearSolver  # noqa: E402

nx, ny, nz = (int(a) for a in sys.argv[1:4])
STEPS = 256
s = synth_blackoil_jacobian(nx, ny, nz, perm="lognormal")
g = GpuLinearSolver(0, experiments=True)          # tracing entry points: experiments build
st = torch.cuda.Stream(); torch.cuda.set_stream(st); g.use_torch_stream()
g.set_pattern(s.rowptr.numpy(), s.colidx.numpy())
vals = s.vals.cuda(); rhs = s.rhs.cuda(); y = torch.zeros_like(rhs)
g.set_values_dev(vals)
assert g.ilu0_factor() == -1
for _ in range(3):
    g.ilu0_apply_dev(0.9, rhs, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.ilu0_apply_dev(0.9, rhs, y)
e1.record(); torch.cuda.synchronize()
print(f"ILU0 apply {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
f = g.lib.opmgpu_debug_gtrace_apply
f.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
f.restype = C.c_int
P = (C.c_int * 2)()
assert f(g.h, STEPS, 0.9, None, None, None, P) == 0
PL, PU = P[0], P[1]
PER = STEPS * 8 + 2048
out = np.zeros(((PL + PU) * PER,), dtype=np.int64)
rc = f(g.h, STEPS, 0.9, C.c_void_p(rhs.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(out.ctypes.data), P)
assert rc == 0, g.error()
tr = {"L": out[: PL * STEPS * 8].reshape(PL, STEPS, 8), "U": out[PL * PER: PL * PER + PU * STEPS * 8].reshape(PU, STEPS, 8)}
for name in ("L", "U"):
    t = tr[name]
    entry = t[:, STEPS - 1, 0]
    act = entry > 0
    T0 = entry[act].min()
    nst = (t[:, : STEPS - 1, 3] > 0).sum(axis=1)
    end = np.array([t[c, nst[c] - 1, 3] if nst[c] else 0 for c in range(t.shape[0])])
    print(f"== {name}: {act.sum()} CTAs, entry skew {entry[act].max() - T0} ns, last chain done at {end.max() - T0} ns")
    order = sorted([c for c in range(t.shape[0]) if nst[c] > 4], key=lambda c: t[c, 0, 3])
    print("   cta  first-done  last-done  steps | median ns: period  prev-done->bar  bar->ext  chain(+dsmem wait) | p90 period | sum period>2*median")
    allper = []
    for c in order:
        n = nst[c]; tt = t[c, :n]
        per = np.diff(tt[:, 3]); allper.append(per)
        med = np.median(per)
        print(f"   {c:3d} {tt[0,3]-T0:9d} {tt[-1,3]-T0:9d} {n:5d} | {med:6.0f} {np.median(tt[1:,1]-tt[:-1,3]):6.0f} {np.median(tt[1:,2]-tt[1:,1]):6.0f} {np.median(tt[1:,3]-tt[1:,2]):6.0f} | {np.percentile(per,90):6.0f} | {per[per>2*med].sum():7.0f}")
    ap = np.concatenate(allper)
    print(f"   all CTAs: step period median {np.median(ap):.0f} mean {ap.mean():.0f} p10 {np.percentile(ap,10):.0f} p90 {np.percentile(ap,90):.0f}")
    # the CTA that finishes last: its full step series
    c = int(np.argmax(end)); n = nst[c]; tt = t[c, :n]
    print(f"   last CTA {c}: periods " + " ".join(str(int(v)) for v in np.diff(tt[:, 3])))
    c = order[0]; n = nst[c]; tt = t[c, :n]
    print(f"   first CTA {c}: periods " + " ".join(str(int(v)) for v in np.diff(tt[:, 3])))
    print(f"   first CTA {c}: chain   " + " ".join(str(int(v)) for v in (tt[:, 3]-tt[:, 2])))
    print(f"   first CTA {c}: bar     " + " ".join(str(int(v)) for v in (tt[1:, 1]-tt[:-1, 3])))
    print(f"   first CTA {c}: rows    " + " ".join(str(int(v)) for v in tt[:, 5]))
