"""Debug: %globaltimer trace of every CTA of the pipelined ILU0 sweeps; walks the critical path
back from the last step and splits it into in-tile steps and tile crossings."""
import ctypes as C
import os
import sys

import numpy as np

os.environ.setdefault("OPMGPU_CLUSTER", "0")      # the traced kernel variants exist without clusters only
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402
from opm_simulators_legacy_b200.solver import GpuLinearSolver  # noqa: E402

nx, ny, nz = (int(a) for a in sys.argv[1:4])
pa, pb = (int(a) for a in os.environ["OPMGPU_TILING"].split("x"))
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
print(f"tiling {pa}x{pb}: ILU0 apply {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
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
hl = {"L": out[PL * STEPS * 8: PL * PER].reshape(PL, 512, 4), "U": out[PL * PER + PU * STEPS * 8:].reshape(PU, 512, 4)}

i0 = [min(i for i in range(nx) if i * pa // nx == a) for a in range(pa)] + [nx]
j0 = [min(j for j in range(ny) if j * pb // ny == b) for b in range(pb)] + [ny]

for name in ("L", "U"):
    t = tr[name]
    entry = t[:, STEPS - 1, 0]
    act = entry > 0
    T0 = entry[act].min()
    nst = (t[:, : STEPS - 1, 3] > 0).sum(axis=1)
    end = np.array([t[c, nst[c] - 1, 3] if nst[c] else 0 for c in range(t.shape[0])])
    print(f"== {name}: {act.sum()} CTAs, entry skew {entry[act].max() - T0} ns, last chain done at {end.max() - T0} ns, steps/CTA {nst[act].min()}..{nst.max()}")
    # per-CTA sums
    tot = dict(land=0.0, bar=0.0, ext=0.0, chain=0.0)
    for c in range(t.shape[0]):
        n = nst[c]
        if n < 2: continue
        tot["bar"] += (t[c, :n, 1] - t[c, :n, 0]).sum(); tot["ext"] += (t[c, :n, 2] - t[c, :n, 1]).sum(); tot["chain"] += (t[c, :n, 3] - t[c, :n, 2]).sum()
        tot["land"] += (t[c, 1:n, 0] - t[c, : n - 1, 3]).sum()
    na = act.sum()
    print("   mean per CTA (ns): after-prev-chain->record landed+static %.0f | bar wait %.0f | ext wait %.0f | chain %.0f" % (tot["land"] / na, tot["bar"] / na, tot["ext"] / na, tot["chain"] / na))
    # tile coordinates; for U the sweep runs from the far corner
    def tile(c): return c % pa, c // pa
    def cta(a, b): return a + pa * b
    up = name == "U"
    # critical path walk: at every step find what the chain start was waiting for last
    def deliv_time(c, q):
        h = hl[name][c]; nd = int((h[:, 0] > 0).sum())
        need = t[c, q, 4]
        if need <= 0 or nd == 0: return 0
        k = int(np.searchsorted(h[:nd, 1] & 0xffffffff, need))
        return h[k, 0] if k < nd else 0
    def upstream(c, q):
        a, b = tile(c); cands = []
        if not up:
            if a > 0: cands.append((cta(a - 1, b), q + (i0[a] - i0[a - 1]) - 1))
            if b > 0: cands.append((cta(a, b - 1), q + (j0[b] - j0[b - 1]) - 1))
        else:
            if a < pa - 1: cands.append((cta(a + 1, b), q + (i0[a + 2] - i0[a + 1]) - 1))
            if b < pb - 1: cands.append((cta(a, b + 1), q + (j0[b + 2] - j0[b + 1]) - 1))
        cands = [(cc, ss) for cc, ss in cands if 0 <= ss < nst[cc]]
        return max(cands, key=lambda x: t[x[0], x[1], 3]) if cands else None
    c = int(np.argmax(end)); q = nst[c] - 1
    acc = dict(chain=0, prev=0, landed=0, cross=0); cnt = dict(prev=0, landed=0, cross=0)
    lam_push = []; lam_wake = []
    tend = t[c, q, 3]
    while True:
        acc["chain"] += t[c, q, 3] - t[c, q, 2]
        start = t[c, q, 2]
        tprev = t[c, q - 1, 3] if q > 0 else 0
        tland = t[c, q, 0]
        tdel = deliv_time(c, q)
        u = upstream(c, q) if tdel >= max(tprev, tland) else None
        if u is not None:
            cnt["cross"] += 1; acc["cross"] += start - t[u[0], u[1], 3]
            lam_push.append(int(tdel - t[u[0], u[1], 3])); lam_wake.append(int(start - tdel))
            if os.environ.get("VERBOSE"):
                h = hl[name][c]; nd = int((h[:, 0] > 0).sum()); k = int(np.searchsorted(h[:nd, 1] & 0xffffffff, t[c, q, 4]))
                seq = " ".join(f"{int(h[m,0]-t[u[0],u[1],3])}:{int(h[m,1]&0xffffffff)}:{int(h[m,1]>>32)}" for m in range(max(0,k-3), min(nd,k+2)))
                a2, b2 = tile(u[0]); a1, b1 = tile(c)
                print(f"      binding crossing CTA {u[0]} ({a2},{b2}) step {u[1]} -> CTA {c} ({a1},{b1}) step {q} rows {t[c,q,5]} ext_end {t[c,q,4]}: push->delivered {int(tdel - t[u[0], u[1], 3])}, ->start {int(start - tdel)}; prev step done {int(tprev - t[u[0],u[1],3]) if tprev else None}; helper deliveries (ns rel push : ext_ready : polls) {seq}")
            c, q = u
        elif tland > tprev or q == 0:
            # record stream: go back to when the previous step of the same group finished (stage
            # recycling is not traced; treat it as the cause)
            cnt["landed"] += 1
            if q == 0: break
            acc["landed"] += start - tprev; q -= 1
        else:
            cnt["prev"] += 1; acc["prev"] += start - tprev; q -= 1
    print(f"   critical path ({tend - T0} ns): chain {acc['chain']} ns | hand-over inside the CTA {acc['prev']} ns ({cnt['prev']} steps) | "
          f"waiting for own record/static load {acc['landed']} ns ({cnt['landed']} steps) | crossings {acc['cross']} ns ({cnt['cross']}; "
          f"push->delivered median {np.median(lam_push) if lam_push else 0:.0f}, delivered->chain start median {np.median(lam_wake) if lam_wake else 0:.0f})")
    crossings = []
    # first-delivery per CTA
    fd = t[:, STEPS - 1, 1]; polls = t[:, STEPS - 1, 2]
    sel = [c for c in (1, 2, pa, pa + 1, 70) if c < t.shape[0] and fd[c] > 0]
    print("   helper first delivery (ns after own entry, polls): " + ", ".join(f"CTA {c}: {fd[c] - entry[c]} ({polls[c]})" for c in sel))
    print("   step period (chain done -> chain done, ns) of CTA 70: " + " ".join(str(int(v)) for v in np.diff(t[70, : nst[70], 3])[:: max(1, nst[70] // 24)]))

    # timeline of one crossing pair on the critical path: last crossing
    if crossings:
        for (cd, sd, cu, su, lam, ew) in crossings[:: max(1, len(crossings) // 3)][:3]:
            h = hl[name][cd]; nd = int((h[:, 0] > 0).sum())
            ht = h[:nd, 0]; hready = h[:nd, 1] & 0xffffffff; hpolls = h[:nd, 1] >> 32
            print(f"   crossing CTA {cu} step {su} -> CTA {cd} step {sd} (lambda {lam} ns): timeline of CTA {cd} around it (ns rel. upstream chain done)")
            for q in range(max(0, sd - 2), min(nst[cd], sd + 3)):
                need = t[cd, q, 4]
                k = int(np.searchsorted(hready, need))          # first delivery reaching ext_end
                tu = t[cu, q - sd + su, 3] if 0 <= q - sd + su < nst[cu] else 0
                hd = ht[k] - tu if k < nd else -1
                hp = (hpolls[k] - (hpolls[k - 1] if k > 0 else 0)) if k < nd else -1
                prev = (ht[k - 1] - tu) if 0 < k < nd else -1
                print(f"      step {q}: rows {t[cd,q,5]} ext_end {need}; landed {t[cd,q,0]-tu} bar {t[cd,q,1]-tu} helper delivered {hd} (prev delivery {prev}, polls since {hp}) ext-ready seen {t[cd,q,2]-tu} chain done {t[cd,q,3]-tu}")
    for cc in (0, 70, 142):
        n = nst[cc]
        if n < 8: continue
        tt = t[cc, :n]
        G = 3
        offp = tt[G:, 0] - tt[:-G, 3]        # own previous chain done -> next own record staged in registers
        onp = tt[:, 3] - tt[:, 0]            # staged -> chain done (barrier wait + ext wait + chain)
        per = tt[G:, 3] - tt[:-G, 3]
        print(f"   CTA {cc}: per step (ns, median): prev chain done -> barrier passed {np.median(tt[1:,1]-tt[:-1,3]):.0f}; barrier -> ext ok {np.median(tt[1:,2]-tt[1:,1]):.0f}; chain {np.median(tt[1:,3]-tt[1:,2]):.0f}; chain done -> chain done {np.median(np.diff(tt[:,3])):.0f}")
        print(f"   CTA {cc}: per own step (ns, median/mean): chain done -> next static loaded {np.median(offp):.0f}/{offp.mean():.0f}; static loaded -> chain done {np.median(onp):.0f}/{onp.mean():.0f}; own period {np.median(per):.0f}/{per.mean():.0f}")

    for cc in (70,):
        h = hl[name][cc]; nd = int((h[:, 0] > 0).sum())
        if nd < 8: continue
        gap = (h[:nd, 2] >> 40) & 0xfffff; ld = (h[:nd, 2] >> 20) & 0xfffff; dl = h[:nd, 2] & 0xfffff
        pol = np.diff(h[:nd, 1] >> 32)
        print(f"   helper of CTA {cc}: {nd} deliveries; cycles median: end of previous delivery -> this poll's start {np.median(gap):.0f}, poll start -> loads back {np.median(ld):.0f}, loads back -> delivered {np.median(dl):.0f}; entries per delivery {np.median(h[:nd,3]):.0f}; polls per delivery {np.median(pol):.0f}; delivery period {np.median(np.diff(h[:nd,0])):.0f} ns")

    rbv = []
    for cc in range(t.shape[0]):
        n = nst[cc]
        sel = t[cc, :n, 6] > 0
        rbv += list((t[cc, :n, 6] - t[cc, :n, 3])[sel])
    if rbv:
        rbv = np.array(rbv)
        print(f"   push -> read back through L2 by the pushing thread (ns): median {np.median(rbv):.0f}, p10 {np.percentile(rbv,10):.0f}, p90 {np.percentile(rbv,90):.0f}, max {rbv.max()} ({len(rbv)} samples)")
