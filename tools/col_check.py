"""Column-owned sweeps on the GPU: bit parity of the ILU0 apply against the CPU oracle on a few
grids, then apply timing.  Usage: python tools/col_check.py [time NX NY NZ [reps]]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from opm_simulators_legacy_b200.solver import GpuLinearSolver  # noqa: E402
from oracle import oracle_py  # noqa: E402


def parity(dims):
    s = synth_blackoil_jacobian(*dims, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    g = GpuLinearSolver(0, experiments=True)          # column-owned sweeps: experiments build
    g.set_pattern(rp, ci)
    g.set_values(v)
    assert g.ilu0_factor() == -1
    lu, bad = oracle_py.ilu0_factor(rp, ci, v)
    ok = True
    for w in (0.9, 1.0):
        for rep in range(3):
            got = g.ilu0_apply(w, b)
            ref = oracle_py.ilu0_apply(rp, ci, lu, w, b)
            same = np.array_equal(got, ref)
            ok = ok and same
            if not same:
                bad_rows = np.nonzero((got != ref).any(1))[0]
                print(f"  MISMATCH dims={dims} w={w} rep={rep}: {len(bad_rows)} rows, first {bad_rows[:8]}, max abs diff {np.nanmax(np.abs(got - ref)):.3e}")
    x, res = g.solve_bcrs(v, b)
    x_ref, ref = oracle_py.solve_bcrs(rp, ci, v, b)
    ok = ok and res["iterations"] == ref["iterations"]
    print(f"parity {dims}: {'ok' if ok else 'FAILED'} (iterations {res['iterations']} / {ref['iterations']})", flush=True)
    g.close()
    return ok


def timing(dims, reps):
    s = synth_blackoil_jacobian(*dims, perm="lognormal")
    g = GpuLinearSolver(0, experiments=True)          # column-owned sweeps: experiments build
    t0 = time.time()
    g.set_pattern(s.rowptr.numpy(), s.colidx.numpy())
    t1 = time.time()
    vals = s.vals.cuda(); rhs = s.rhs.cuda(); y = torch.zeros_like(rhs)
    g.set_values_dev(vals)
    assert g.ilu0_factor() == -1
    for _ in range(3):
        g.ilu0_apply_dev(0.9, rhs, y)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    st = torch.cuda.current_stream()
    g.use_torch_stream()
    ev[0].record(st)
    for _ in range(reps):
        g.ilu0_apply_dev(0.9, rhs, y)
    ev[1].record(st)
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) * 1e3 / reps
    N = s.N; nnzb = len(s.colidx)
    algo = 76 * (nnzb - N) + 176 * N
    print(f"apply {dims}: {us:.1f} us per apply (incl. permute kernel), {algo / us * 1e-3:.0f} GB/s algorithmic, frac of 6550 = {algo / us * 1e-3 / 6550.4:.3f}; analysis {t1 - t0:.2f} s", flush=True)
    g.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        timing(tuple(int(a) for a in sys.argv[2:5]), int(sys.argv[5]) if len(sys.argv) > 5 else 20)
    else:
        allok = True
        for dims in [(10, 10, 3), (24, 20, 12), (64, 1, 1), (30, 17, 1), (7, 6, 40), (40, 40, 20), (120, 125, 6), (100, 100, 100)]:
            allok = parity(dims) and allok
        print("ALL OK" if allok else "FAILURES")
