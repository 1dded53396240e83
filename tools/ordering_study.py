"""Host-only study (CPU oracle): BiCGStab iteration counts of the ILU0 of P A P^T for several
parallel orderings of a Cartesian black-oil system.  Natural order is the reference's; red-black is
what the multicolour variant of the library builds; "cubes of B" is red-black over B x B x B cubes with
the natural order kept inside a cube (3B-2 dependency levels per cube, all cubes of a colour in parallel).
Usage: python tools/ordering_study.py NX NY NZ"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def permute_bcrs(rp, ci, v, n2p):
    N = rp.size - 1
    rows = np.repeat(np.arange(N), np.diff(rp))
    pr, pc = n2p[rows], n2p[ci]
    order = np.lexsort((pc, pr))
    prp = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(np.bincount(pr, minlength=N), out=prp[1:])
    return prp, pc[order].astype(np.int32), np.ascontiguousarray(v[order])


def main():
    nx, ny, nz = (int(a) for a in sys.argv[1:4])
    s = synth_blackoil_jacobian(nx, ny, nz, perm="lognormal")
    rp, ci, v, b = s.rowptr.numpy(), s.colidx.numpy(), s.vals.numpy(), s.rhs.numpy().reshape(-1, 3)
    N = s.N
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()          # natural cell index = i + nx (j + ny k)
    nat = np.arange(N)
    orderings = {"natural": nat}
    orderings["red-black"] = np.lexsort((nat, (i + j + k) % 2))
    for B in (2, 4, 8, 16):
        colour = (i // B + j // B + k // B) % 2
        cube = (i // B) + ((nx + B - 1) // B) * ((j // B) + ((ny + B - 1) // B) * (k // B))
        orderings[f"cubes of {B}"] = np.lexsort((nat, cube, colour))
    orderings["k-lines red-black"] = np.lexsort((nat, (i + j) % 2))
    orderings["k-planes red-black"] = np.lexsort((nat, k % 2))
    print(f"{nx}x{ny}x{nz}: iterations / half steps to 1e-2 (ILU0 of P A P^T, relaxation 0.9)")
    for name, p2n in orderings.items():
        n2p = np.empty(N, dtype=np.int64)
        n2p[p2n] = np.arange(N)
        prp, pci, pv = permute_bcrs(rp, ci, v, n2p)
        x, r = O.solve_bcrs(prp, pci, pv, b[p2n].reshape(-1))
        print(f"  {name:22s} {r['iterations']:3d} / {r['half_steps']:3d}")


if __name__ == "__main__":
    main()
