"""Synthetic three-phase black-oil Jacobians on Cartesian grids.

Input synthesis for tests and benchmarks -- not part of the solver.  The blocks are shaped
like the reference's assembly (opm/autodiff/BlackoilModelBase_impl.hpp:848-913
assembleMassBalanceEq, :712-751 computeAccum, :1499-1511 flux; TPFA transmissibilities as
opm/autodiff/GeoProps.hpp:121-153): rows = [water, oil, gas] equations, columns =
[p, sw, sg]; the pressure column of every off-diagonal block is non-zero, the saturation
columns only when the neighbour is upwind, so the pattern of the interleaved system is the
union of the pressure-derivative patterns, exactly what formInterleavedSystem assumes
(opm/autodiff/NewtonIterationBlackoilInterleaved.cpp:116-123).

Written against torch so the same code builds a 300-cell case on the CPU and a 1e8-cell
case directly in HBM.  Random fields come from numpy's default_rng with the seeds the
survey fixes (20190401 fields, 20190402 x*), so CPU and GPU builds see identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

MATBALSCALE = (1.1169, 1.0031, 0.0031)   # BlackoilModelBase_impl.hpp:139

_MD = 9.869233e-16      # m^2 per millidarcy
_BAR = 1.0e5
_DAY = 86400.0


@dataclass
class SynthSystem:
    """A·dx = residual in both of the layouts that cross the solver boundary."""
    dims: tuple
    N: int
    nnzb: int
    rowptr: torch.Tensor          # int32 [N+1]
    colidx: torch.Tensor          # int32 [nnzb], ascending per row
    vals_unscaled: torch.Tensor   # f64 [nnzb, 9], block row-major [eq][var]
    sat_present: torch.Tensor     # bool [nnzb]: saturation columns structurally present
    rhs_unscaled: torch.Tensor    # f64 [N, 3] = A_unscaled · xstar  (cell-major)
    xstar: torch.Tensor           # f64 [N, 3]
    matbalscale: tuple = MATBALSCALE
    _vals: Optional[torch.Tensor] = field(default=None, repr=False)

    @property
    def vals(self) -> torch.Tensor:
        """Scaled BCRS values, i.e. what formInterleavedSystem stores after :234-236."""
        if self._vals is None:
            s = torch.tensor(self.matbalscale, dtype=torch.float64, device=self.vals_unscaled.device)
            self._vals = (self.vals_unscaled.view(-1, 3, 3) * s.view(1, 3, 1)).reshape(-1, 9).contiguous()
        return self._vals

    @property
    def rhs(self) -> torch.Tensor:
        """Scaled, cell-major right-hand side (istlb at :263-269)."""
        s = torch.tensor(self.matbalscale, dtype=torch.float64, device=self.rhs_unscaled.device)
        return (self.rhs_unscaled * s.view(1, 3)).contiguous()

    @property
    def rhs_eqmajor_unscaled(self) -> torch.Tensor:
        """Equation-major residual values as LinearisedBlackoilResidual carries them."""
        return self.rhs_unscaled.t().contiguous().view(-1)

    def csc_blocks(self):
        """The nine scalar Jacobian blocks d(eq p1)/d(var p2) in Eigen's column-major
        compressed layout (AutoDiffMatrix::getSparse), unscaled.  Returns a list of nine
        (colptr int32 [N+1], rowidx int32 [nnz], val f64 [nnz]) numpy triples, index p1*3+p2."""
        rowptr = self.rowptr.cpu().numpy().astype(np.int64)
        col = self.colidx.cpu().numpy()
        row = np.repeat(np.arange(self.N, dtype=np.int32), np.diff(rowptr))
        order = np.argsort(col, kind="stable")          # by column, rows stay ascending
        vals = self.vals_unscaled.cpu().numpy()
        present = self.sat_present.cpu().numpy()
        out = []
        for p1 in range(3):
            for p2 in range(3):
                sel = order if p2 == 0 else order[present[order]]
                c = col[sel]
                colptr = np.zeros(self.N + 1, dtype=np.int32)
                np.cumsum(np.bincount(c, minlength=self.N), out=colptr[1:])
                out.append((colptr, np.ascontiguousarray(row[sel]),
                            np.ascontiguousarray(vals[sel, p1 * 3 + p2])))
        return out


def cartesian_pattern(nx: int, ny: int, nz: int, device="cpu"):
    """7-point stencil in natural ordering c = i + nx (j + ny k), ascending columns.
    Returns rowptr, colidx and the slot of each of the 7 positions (or -1)."""
    N = nx * ny * nz
    c = torch.arange(N, device=device, dtype=torch.int64)
    i = c % nx
    j = (c // nx) % ny
    k = c // (nx * ny)
    has = [k > 0, j > 0, i > 0, None, i < nx - 1, j < ny - 1, k < nz - 1]
    off = [-nx * ny, -nx, -1, 0, 1, nx, nx * ny]
    cnt = torch.ones(N, device=device, dtype=torch.int64)
    for h in has:
        if h is not None:
            cnt += h
    rowptr = torch.zeros(N + 1, device=device, dtype=torch.int64)
    torch.cumsum(cnt, 0, out=rowptr[1:])
    slots = []
    pos = rowptr[:-1].clone()
    for h in has:
        if h is None:
            slots.append(pos.clone())
            pos += 1
        else:
            slots.append(torch.where(h, pos, torch.full_like(pos, -1)))
            pos += h
    nnzb = int(rowptr[-1])
    colidx = torch.empty(nnzb, device=device, dtype=torch.int32)
    for s, o in zip(slots, off):
        m = s >= 0
        colidx[s[m]] = (c[m] + o).to(torch.int32)
    return rowptr.to(torch.int32), colidx, slots


def synth_blackoil_jacobian(nx: int, ny: int, nz: int, perm: str = "homogeneous",
                            sigma: float = 2.0, seed: int = 20190401,
                            device="cpu", xstar: str = "correlated",
                            corr_len: float = 8.0) -> SynthSystem:
    """xstar: "correlated" (default) draws the increment the rhs is built from as a
    Gaussian-correlated field (correlation length corr_len cells) plus 1 % white noise --
    a Newton increment is smooth away from fronts, and a smooth x* is what makes ILU0 +
    BiCGStab take a realistic 5-20 iterations to reduce the residual by 1e-2; "white" is
    the survey's i.i.d. field (1-2 iterations), "none" skips x*/rhs (kernel sweeps)."""
    dev = torch.device(device)
    f64 = torch.float64
    N = nx * ny * nz
    rng = np.random.default_rng(seed)
    g = rng.standard_normal(N)
    u1 = rng.random(N)
    u2 = rng.random(N)
    u3 = rng.random(N)
    rngx = np.random.default_rng(seed + 1)        # 20190402 for the default seed
    if xstar == "none":
        xs = np.zeros((1, 3))
    else:
        xs = rngx.standard_normal((N, 3))
        if xstar == "correlated":
            from scipy.ndimage import gaussian_filter
            white = rngx.standard_normal(N)
            for q in range(3):
                f = gaussian_filter(xs[:, q].reshape(nz, ny, nx), corr_len, mode="reflect").ravel()
                xs[:, q] = f / f.std()
            xs[:, 0] += 0.01 * white
        elif xstar != "white":
            raise ValueError(xstar)

    def T(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev, f64)

    g, u1, u2, u3 = T(g), T(u1), T(u2), T(u3)
    xstar_t = T(xs) * torch.tensor([1.0 * _BAR, 0.01, 0.01], dtype=f64, device=dev)

    dx, dy, dz = 20.0, 20.0, 2.0
    phi, dt = 0.2, 10.0 * _DAY
    c = torch.arange(N, device=dev, dtype=torch.int64)
    ci = c % nx
    cj = (c // nx) % ny
    ck = c // (nx * ny)

    K = torch.full((N,), 100.0 * _MD, dtype=f64, device=dev)
    if perm == "lognormal":
        K = K * torch.exp(sigma * g)
    elif perm != "homogeneous":
        raise ValueError(perm)
    Kdir = (K, K, 0.1 * K)
    area = (dy * dz, dx * dz, dx * dy)
    dist = (dx, dy, dz)

    # state: hydrostatic + smooth 5-bar potential (fixes upwind directions) + jitter (no ties)
    xx = (ci.to(f64) + 0.5) / nx
    yy = (cj.to(f64) + 0.5) / ny
    zz = (ck.to(f64) + 0.5) / nz
    Phi = 5.0 * _BAR * (0.5 * torch.sin(2 * np.pi * xx) * torch.cos(np.pi * yy)
                        + 0.3 * torch.cos(3 * np.pi * zz) * torch.sin(np.pi * xx + 0.3)
                        + 0.2 * yy) + 1.0e3 * u3
    p0 = 250.0 * _BAR
    p = p0 + 800.0 * 9.81 * (ck.to(f64) + 0.5) * dz + Phi
    sw = 0.2 + 0.1 * u1
    sg = 0.1 + 0.1 * u2
    so = 1.0 - sw - sg

    mu = (0.5e-3, 1.0e-3, 0.02e-3)
    cw, co, cg = 4e-5 / _BAR, 1e-4 / _BAR, 5e-3 / _BAR
    bw0, bo0, bg0 = 1.0, 0.85, 200.0
    bw, dbw = bw0 * (1 + cw * (p - p0)), torch.full_like(p, bw0 * cw)
    bo, dbo = bo0 * (1 + co * (p - p0)), torch.full_like(p, bo0 * co)
    bg, dbg = bg0 * (1 + cg * (p - p0)), torch.full_like(p, bg0 * cg)
    rs, drs = 0.5 * p / _BAR, torch.full_like(p, 0.5 / _BAR)

    z = torch.zeros_like(p)
    # mobility-like terms M[eq] and dM[eq][var] per cell
    lw, lo, lg = sw * sw / mu[0], so * so / mu[1], sg * sg / mu[2]
    dlw, dlo, dlg = 2 * sw / mu[0], -2 * so / mu[1], 2 * sg / mu[2]   # d/dsw, d/d(sw|sg), d/dsg
    M = torch.stack([bw * lw, bo * lo, bg * lg + rs * bo * lo], 1)            # [N,3]
    dM = torch.stack([
        torch.stack([dbw * lw, bw * dlw, z], 1),
        torch.stack([dbo * lo, bo * dlo, bo * dlo], 1),
        torch.stack([dbg * lg + (drs * bo + rs * dbo) * lo, rs * bo * dlo, bg * dlg + rs * bo * dlo], 1),
    ], 1)                                                                      # [N,3,3]
    a = phi * dx * dy * dz / dt
    dAcc = a * torch.stack([
        torch.stack([dbw * sw, bw, z], 1),
        torch.stack([dbo * so, -bo, -bo], 1),
        torch.stack([dbg * sg + (drs * bo + rs * dbo) * so, -rs * bo, bg - rs * bo], 1),
    ], 1)

    rowptr, colidx, slots = cartesian_pattern(nx, ny, nz, dev)
    nnzb = int(rowptr[-1])
    vals = torch.zeros((nnzb, 3, 3), dtype=f64, device=dev)
    present = torch.zeros(nnzb, dtype=torch.bool, device=dev)
    s_km, s_jm, s_im, s_d, s_ip, s_jp, s_kp = slots
    vals[s_d] = dAcc
    present[s_d] = True
    ep = torch.tensor([1.0, 0.0, 0.0], dtype=f64, device=dev).view(1, 1, 3)

    strides = (1, nx, nx * ny)
    lowmask = (ci < nx - 1, cj < ny - 1, ck < nz - 1)
    up_slot = (s_ip, s_jp, s_kp)
    dn_slot = (s_im, s_jm, s_km)
    for d in range(3):
        c1 = c[lowmask[d]]
        c2 = c1 + strides[d]
        t1 = Kdir[d][c1] * area[d] / (0.5 * dist[d])
        t2 = Kdir[d][c2] * area[d] / (0.5 * dist[d])
        Tf = 1.0 / (1.0 / t1 + 1.0 / t2)
        dPhi = Phi[c1] - Phi[c2]
        up1 = dPhi >= 0
        upc = torch.where(up1, c1, c2)
        TM = (Tf.view(-1, 1) * M[upc]).view(-1, 3, 1) * ep            # T M(up) e_p
        TdM = (Tf * dPhi).view(-1, 1, 1) * dM[upc]                    # T dPhi dM(up)
        G1 = TM + TdM * up1.view(-1, 1, 1)                             # dF(c1->c2)/d vars(c1)
        G2 = -TM + TdM * (~up1).view(-1, 1, 1)                         # dF(c1->c2)/d vars(c2)
        vals.index_add_(0, s_d[c1], G1)
        vals.index_add_(0, s_d[c2], -G2)
        vals[up_slot[d][c1]] = G2
        vals[dn_slot[d][c2]] = -G1
        present[up_slot[d][c1]] = ~up1
        present[dn_slot[d][c2]] = up1

    vals = vals.reshape(nnzb, 9).contiguous()
    sysm = SynthSystem((nx, ny, nz), N, nnzb, rowptr, colidx, vals, present,
                       torch.empty(0), xstar_t)
    if xstar != "none":
        sysm.rhs_unscaled = bcrs_matvec(rowptr, colidx, vals, xstar_t)
    return sysm


def bcrs_matvec(rowptr, colidx, vals, x):
    """Plain torch y = A x for 3x3 BCRS (input synthesis / property tests only)."""
    N = rowptr.numel() - 1
    counts = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(N, device=vals.device), counts)
    prod = torch.einsum("kab,kb->ka", vals.view(-1, 3, 3), x.view(-1, 3)[colidx.to(torch.int64)])
    y = torch.zeros((N, 3), dtype=vals.dtype, device=vals.device)
    y.index_add_(0, rows, prod)
    return y


def random_bcrs(N: int, extra_per_row: int = 3, seed: int = 1, dense_group: int = 0,
                diag_boost: float = 4.0):
    """General (non-stencil) block-diagonally-dominant BCRS for edge-case tests: a random
    structurally symmetric pattern plus, optionally, one densely coupled group of cells
    such as the Schur complement of a multi-perforation well produces
    (opm/autodiff/NewtonIterationUtilities.cpp:98-115).  numpy arrays."""
    rng = np.random.default_rng(seed)
    pairs = set()
    for r in range(N):
        pairs.add((r, r))
        for cc in rng.integers(0, N, size=extra_per_row):
            pairs.add((r, int(cc)))
            pairs.add((int(cc), r))
        if r + 1 < N:
            pairs.add((r, r + 1)); pairs.add((r + 1, r))
    if dense_group > 1:
        grp = rng.choice(N, size=min(dense_group, N), replace=False)
        for a_ in grp:
            for b_ in grp:
                pairs.add((int(a_), int(b_)))
    pairs = sorted(pairs)
    rows = np.array([q[0] for q in pairs]); cols = np.array([q[1] for q in pairs], dtype=np.int32)
    rowptr = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=N), out=rowptr[1:])
    vals = rng.standard_normal((len(pairs), 9))
    absrow = np.zeros(N)
    np.add.at(absrow, rows, np.abs(vals).sum(1))
    dmask = rows == cols
    vals[dmask] += (np.eye(3).reshape(1, 9)) * (diag_boost * absrow[rows[dmask]]).reshape(-1, 1) / 3.0
    return rowptr, cols, np.ascontiguousarray(vals)


def block_system_np(rowptr, colidx, np_: int, seed: int = 1, diag_boost: float = 3.0):
    """Values and right-hand side of an np_ x np_ block system on a given BCRS pattern (block sizes
    4..6: the reference's Impl<np,Scalar> for the polymer / solvent extensions,
    opm/autodiff/NewtonIterationBlackoilInterleaved.cpp:467-487): random blocks, block rows made
    diagonally dominant.  Returns vals[nnzb, np_*np_], rhs[N, np_], xstar[N, np_] (numpy)."""
    rowptr = np.asarray(rowptr); colidx = np.asarray(colidx)
    N = rowptr.size - 1
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(N), np.diff(rowptr))
    vals = rng.standard_normal((colidx.size, np_ * np_))
    absrow = np.zeros(N)
    np.add.at(absrow, rows, np.abs(vals).sum(1))
    dmask = rows == colidx
    vals[dmask] += np.eye(np_).reshape(1, -1) * (diag_boost * absrow[rows[dmask]]).reshape(-1, 1) / np_
    xstar = rng.standard_normal((N, np_))
    rhs = np.zeros((N, np_))
    prod = np.einsum("kij,kj->ki", vals.reshape(-1, np_, np_), xstar[colidx])
    np.add.at(rhs, rows, prod)
    return np.ascontiguousarray(vals), rhs, xstar
