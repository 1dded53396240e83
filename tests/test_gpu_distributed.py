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
