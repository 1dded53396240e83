"""Multi-GPU parity on real hardware (`-m gpu`, needs >= 2 GPUs on the box): the row-partitioned
solve under torchrun (tests/dist_check.py is the per-rank program) against the CPU oracle's
unpartitioned solve.  Partitioned parity per BASELINE.json north_star: the true residual of the
gathered increment is reduced by linear_solver_reduction, the distributed SpMV is bit-exact,
iteration counts are reported side by side.  The reference's counterpart is the MPI branch of
ISTLSolver::solve (opm/autodiff/ISTLSolver.hpp:286-298)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, *script_args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_check.py"), *[str(a) for a in script_args]]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    return json.loads(lines[-1])


def _need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs on one box (gpurun --gpus {n})")


@pytest.mark.parametrize("dims", [(24, 20, 16), (40, 40, 40)])
def test_partitioned_solve_two_gpus(dims):
    _need(2)
    rep = _torchrun(2, *dims)
    assert rep["ok"] and rep["spmv_bit_exact"]
    for red in ("red_0.01", "red_1e-08"):
        r = rep[red]
        assert r["converged"] == 1 and r["true_residual_reduction"] <= float(red[4:]) * 1.0001
        assert r["iterations_partitioned"] >= 1 and r["iterations_oracle_unpartitioned"] >= 1


def test_partitioned_solve_four_gpus():
    _need(4)
    rep = _torchrun(4, 40, 36, 24)
    assert rep["ok"] and rep["spmv_bit_exact"]


def test_singular_slab_ends_the_solve_on_every_rank():
    """A singular ILU0 pivot on one rank only: the status is agreed on with an all-reduce, so every
    rank raises NumericalIssue instead of the healthy ranks waiting forever in the next collective."""
    _need(2)
    rep = _torchrun(2, 24, 20, 16, "--singular", timeout=300)
    assert rep["ok"], rep


# ---- multi-GPU beneath the C ABI: ONE process, opmgpu_create_multi ----
def _true_reduction(oracle, rp, ci, v, b, x):
    import numpy as np
    return float(np.linalg.norm(b - oracle.spmv(rp, ci, v, x)) / np.linalg.norm(b))


@pytest.mark.parametrize("ngpus", [1, 2, 4])
def test_multi_handle_cartesian(oracle, ngpus):
    """Global arrays in, global increment out; the handle partitions (weakest-coupling slabs),
    builds the halo plan and runs one worker thread per GPU.  One GPU: identical to the plain
    handle's answer; several: partitioned parity (true residual within the tolerance)."""
    import numpy as np
    from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian
    from opm_simulators_legacy_b200.solver import GpuLinearSolver
    _need(ngpus)
    s = synth_blackoil_jacobian(24, 20, 16, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    g = GpuLinearSolver.multi(range(ngpus))
    try:
        g.set_pattern(rp, ci)
        for red in (1e-2, 1e-8):
            x, res = g.solve_bcrs(v, b, linear_solver_reduction=red, linear_solver_maxiter=400)
            assert res["converged"] == 1
            assert _true_reduction(oracle, rp, ci, v, b, x) <= red * 1.0001
            x_ref, ref = oracle.solve_bcrs(rp, ci, v, b, reduction=red, maxiter=400)
            if ngpus == 1:            # one GPU: the same arithmetic as the plain handle, bit for bit
                plain = GpuLinearSolver(0)
                plain.set_pattern(rp, ci)
                x_p, res_p = plain.solve_bcrs(v, b, linear_solver_reduction=red, linear_solver_maxiter=400)
                plain.close()
                assert res["iterations"] == ref["iterations"] == res_p["iterations"]
                assert np.array_equal(x, x_p)
        axis, offs = g.multi_partition()
        assert offs[0] == 0 and offs[-1] == s.N and all(offs[i] < offs[i + 1] for i in range(ngpus))
        assert axis == (-1 if ngpus == 1 else axis) and axis in (-1, 0, 1, 2)
        # the CSC front end (formInterleavedSystem on the host side of the multi handle)
        dx, res = g.solve_from_csc_blocks(s.N, s.csc_blocks(), s.matbalscale, s.rhs_eqmajor_unscaled.numpy())
        assert res["converged"] == 1 and _true_reduction(oracle, rp, ci, v, b, dx.reshape(3, -1).T) <= 1e-2 * 1.0001
        # ... and again with the cached pattern
        dx2, res2 = g.solve_from_csc_blocks(s.N, s.csc_blocks(), s.matbalscale, s.rhs_eqmajor_unscaled.numpy())
        assert np.array_equal(dx, dx2) and res2["iterations"] == res["iterations"]
    finally:
        g.close()


@pytest.mark.parametrize("ngpus", [1, 2])
def test_multi_handle_general_pattern(oracle, ngpus):
    """No grid behind the pattern (random couplings plus a dense group, as Schur fill of a well
    produces): contiguous row blocks, halo plan from the pattern alone."""
    import numpy as np
    from opm_simulators_legacy_b200.jacobian import random_bcrs
    from opm_simulators_legacy_b200.solver import GpuLinearSolver
    _need(ngpus)
    rp, ci, v = random_bcrs(900, 3, seed=11, dense_group=12)
    b = np.random.default_rng(3).standard_normal((900, 3))
    g = GpuLinearSolver.multi(range(ngpus))
    try:
        g.set_pattern(rp, ci)
        x, res = g.solve_bcrs(v, b, linear_solver_reduction=1e-8, linear_solver_maxiter=400)
        assert res["converged"] == 1 and _true_reduction(oracle, rp, ci, v, b, x) <= 1e-8 * 1.0001
        assert g.multi_partition()[0] == -1
    finally:
        g.close()


def test_cpp_host_mirror_on_two_gpus():
    """host_selftest drives the same NewtonIterationBlackoilGPU class with gpu_devices=0,1: wells on
    the host, cells on two GPUs of one process."""
    import subprocess
    _need(2)
    exe = os.path.join(ROOT, "opm_simulators_legacy_b200", "host_selftest")
    out = subprocess.run([exe, "0,1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.parametrize("ngpus", [1, 2])
def test_multi_handle_single_precision(oracle, ngpus):
    """The float instance (Impl<3,float>) on a multi handle: every GPU runs in float (halo exchange of
    float vectors); one GPU reproduces the float oracle's iteration count, two satisfy the partitioned
    tolerance rule on the true residual."""
    import numpy as np
    from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian
    from opm_simulators_legacy_b200.solver import GpuLinearSolver
    _need(ngpus)
    s = synth_blackoil_jacobian(24, 20, 16, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy()
    g = GpuLinearSolver.multi(range(ngpus))
    try:
        g.set_precision(True)
        g.set_pattern(rp, ci)
        x, res = g.solve_bcrs(v, b)
        assert res["converged"] == 1
        assert np.array_equal(x, x.astype(np.float32).astype(np.float64))
        assert _true_reduction(oracle, rp, ci, v, b, x) <= 1e-2 * 1.001
        if ngpus == 1:
            assert res["iterations"] == oracle.f32.solve_bcrs(rp, ci, v, b)[1]["iterations"]
        dx, res = g.solve_from_csc_blocks(s.N, s.csc_blocks(), s.matbalscale, s.rhs_eqmajor_unscaled.numpy())
        assert res["converged"] == 1 and _true_reduction(oracle, rp, ci, v, b, dx.reshape(3, -1).T) <= 1e-2 * 1.001
    finally:
        g.close()
