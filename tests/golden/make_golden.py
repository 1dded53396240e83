"""Generates the committed golden fixtures (tests/golden/*.npz) from the CPU oracle.

The reference's own implementation of this path cannot be run here and its
tests hold no result vectors for it, so the fixtures pin the ORACLE: they make
any later change of the oracle's arithmetic visible, and give the GPU tests inputs/outputs that
do not depend on the oracle being rebuilt identically on the GPU box.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian, random_bcrs  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dump_f32(name, rp, ci, v, b, x_probe):
    """The same quantities from the float instance of the oracle (Impl<3,float>): float32 arrays."""
    F = O.f32
    os.makedirs(os.path.join(HERE, "f32"), exist_ok=True)
    lu, bad = F.ilu0_factor(rp, ci, v)
    assert bad == -1
    x, res = F.solve_bcrs(rp, ci, v, b)
    np.savez_compressed(os.path.join(HERE, "f32", name + ".npz"), rowptr=rp, colidx=ci, vals=v, rhs=b, x_probe=x_probe,
                        spmv=F.spmv(rp, ci, v, x_probe), lu=lu, apply_w09=F.ilu0_apply(rp, ci, lu, 0.9, b),
                        apply_w1=F.ilu0_apply(rp, ci, lu, 1.0, b), x=x, iterations=res["iterations"],
                        half_steps=res["half_steps"], reduction=res["reduction"])
    print("f32/" + name, "iterations", res["iterations"])


def dump(name, rp, ci, v, b, x_probe):
    dump_f32(name, rp, ci, v, b, x_probe)
    lu, bad = O.ilu0_factor(rp, ci, v)
    assert bad == -1
    x, res = O.solve_bcrs(rp, ci, v, b)
    x5, res5 = O.solve_bcrs(rp, ci, v, b, reduction=1e-30, max_half_steps=5)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), rowptr=rp, colidx=ci, vals=v, rhs=b, x_probe=x_probe,
                        spmv=O.spmv(rp, ci, v, x_probe), lu=lu, apply_w09=O.ilu0_apply(rp, ci, lu, 0.9, b),
                        apply_w1=O.ilu0_apply(rp, ci, lu, 1.0, b), x=x, iterations=res["iterations"],
                        half_steps=res["half_steps"], reduction=res["reduction"], x_5half=x5)
    print(name, "N", len(rp) - 1, "nnzb", len(ci), "iterations", res["iterations"])


s = synth_blackoil_jacobian(10, 10, 3)                      # C1: the SPE1-shaped case
dump("c1_spe1_shape", s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy(), s.xstar.numpy())
s = synth_blackoil_jacobian(12, 9, 7, perm="lognormal")
dump("lognormal_12x9x7", s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy(), s.xstar.numpy())
rp, ci, v = random_bcrs(240, extra_per_row=3, seed=11, dense_group=9)
rng = np.random.default_rng(4)
dump("general_wells_240", rp, ci, v, rng.standard_normal((240, 3)), rng.standard_normal((240, 3)))
