"""Small workload for compute-sanitizer (tools/sanitize.sh): the CSC front end, the pipelined
factorisation, the pipelined sweeps with thread-block clusters, several tiles per CTA, the
general (non-stencil) kernels, GMRES -- every kernel family of the library once, on grids small
enough for the sanitizer's slow-down."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, random_bcrs  # noqa: E402
from opm_simulators_legacy_b200.solver import GpuLinearSolver  # noqa: E402

g = GpuLinearSolver(0)
for dims in [(16, 12, 8), (40, 40, 6)]:
    s = synth_blackoil_jacobian(*dims, perm="lognormal")
    dx, res = g.solve_from_csc_blocks(s.N, s.csc_blocks(), s.matbalscale, s.rhs_eqmajor_unscaled.numpy())
    print("csc solve", dims, res["iterations"], res["status"], flush=True)
    x, res = g.solve_bcrs(s.vals.numpy(), s.rhs.numpy(), newton_use_gmres=True, linear_solver_restart=5)
    print("gmres", dims, res["iterations"], res["status"], flush=True)
rp, ci, v = random_bcrs(600, 3, seed=7, dense_group=12)
g.set_pattern(rp, ci)
b = np.random.default_rng(1).standard_normal((600, 3))
x, res = g.solve_bcrs(v, b, raise_on_failure=False)
print("general pattern", res["iterations"], res["status"], flush=True)
g.close()
print("done")
