"""Kernel micro-benchmarks through the per-box C ABI: one 256^3 box spanning the periodic domain,
CUDA events on the launching stream, warm-up + repeated launches (arrays >> L2 where a single
array is 134 MB).  Prints us per call and the ALGORITHMIC GB/s (DESIGN.md bytes per unit).
Usage: python scripts/kbench.py [n] [reps]"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import iamr_b200 as ix
from util import box_of, d3, stream_of

spec = "--spec" in sys.argv     # SURVEY.md 8d inputs (hashed phi / rhs, sinusoidal rho, Taylor-Green velocities) instead of uniform noise
argv = [a for a in sys.argv if a != "--spec"]
n = int(argv[1]) if len(argv) > 1 else 256
reps = int(argv[2]) if len(argv) > 2 else 20
dev = 'cuda:0'
lib = ix.load()
s = stream_of(dev)
DXINV = (float(n),) * 3


def fab(shape_valid, ng, ncomp=1, fill=None):
    shp = (ncomp,) + tuple(m + 2 * ng for m in reversed(shape_valid))
    t = torch.rand(shp, dtype=torch.float64, device=dev) if fill is None else torch.full(shp, fill, dtype=torch.float64, device=dev)
    return t, ix.fab_of(t, [-ng] * 3)


def timeit(name, fn, algo_bytes, reps=reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name:44s} {us:10.1f} us  {algo_bytes / us / 1e3:8.1f} GB/s (algorithmic)", flush=True)
    return us


cells = (n, n, n)
nodes = (n + 1, n + 1, n + 1)
bx = box_of((0, 0, 0), (n - 1, n - 1, n - 1))
nbx = box_of((0, 0, 0), (n, n, n))
N = n ** 3
NN = (n + 1) ** 3

# ---- nodal ----
tphi, fphi = fab(nodes, 1)
tphi2, fphi2 = fab(nodes, 1)
trhs, frhs = fab(nodes, 1)
tsig, fsig = fab(cells, 1)
tsig.add_(1.0)


def gs8():
    for c in range(8):
        lib.check(lib.iamrx_nodal_gs_box(C.byref(nbx), C.byref(fphi), C.byref(frhs), C.byref(fsig), d3(DXINV), c, s))


if hasattr(lib, 'iamrx_nodal_gs_sweep_box'):
    timeit("nodal GS fused sweep (2 launches)",
           lambda: lib.check(lib.iamrx_nodal_gs_sweep_box(C.byref(nbx), C.byref(fphi2), C.byref(fphi), C.byref(frhs), C.byref(fsig), d3(DXINV), s)),
           32.0 * NN)
timeit("nodal GS 8 colour launches (no wrap)", gs8, 32.0 * NN)
tout, fout = fab(nodes, 1)
timeit("nodal adotx residual", lambda: lib.check(lib.iamrx_nodal_adotx_box(C.byref(nbx), C.byref(fout), C.byref(fphi), C.byref(frhs), C.byref(fsig), d3(DXINV), s)), 32.0 * NN)

# ---- ABec ----
tp, fp = fab(cells, 1)
tr, fr = fab(cells, 0)
to, fo = fab(cells, 0)
tb = [fab(tuple(m + (1 if d == q else 0) for q, m in enumerate(cells)), 0) for d in range(3)]
for t, _ in tb:
    t.add_(0.5)
fb = [f for _, f in tb]
ta, fa = fab(cells, 0)
if spec:   # K1 of SURVEY.md 8d
    sys.path.insert(0, 'scripts')
    import spec_inputs as si
    k1 = si.k1_gsrb(n, dev)
    tp.copy_(si.ghosted(k1["phi"], 1)); tr.copy_(si.ghosted(k1["rhs"], 0)); ta.copy_(si.ghosted(k1["rho"], 0))
    for d in range(3):
        tb[d][0].copy_(si.ghosted(k1["beta"][d], 0, tuple(1 if q == d else 0 for q in range(3))))
    del k1


def gsrb(a):
    def f():
        for rb in range(2):
            lib.check(lib.iamrx_abec_gsrb_box(C.byref(bx), C.byref(fp), C.byref(fr), a, 1.0, C.byref(fa) if a else None, C.byref(fb[0]),
                                              C.byref(fb[1]), C.byref(fb[2]), d3(DXINV), 1.15, rb, 1, s))
    return f


timeit("ABec GSRB red+black a=0 (2 launches)", gsrb(0.0), 96.0 * N)
timeit("ABec GSRB red+black a=1 (2 launches)", gsrb(1.0), 112.0 * N)
if hasattr(lib, 'iamrx_abec_gsrb_sweep_box'):
    tp2, fp2 = fab(cells, 1)
    for a in (0.0, 1.0):
        timeit(f"ABec GSRB fused sweep a={a:g}",
               lambda: lib.check(lib.iamrx_abec_gsrb_sweep_box(C.byref(bx), C.byref(fp2), C.byref(fp), C.byref(fr), a, 1.0, C.byref(fa) if a else None,
                                                               C.byref(fb[0]), C.byref(fb[1]), C.byref(fb[2]), d3(DXINV), 1.15, 1, s)),
               (48.0 + (8.0 if a else 0.0)) * N)
timeit("ABec residual a=0", lambda: lib.check(lib.iamrx_abec_apply_box(C.byref(bx), C.byref(fo), C.byref(fp), C.byref(fr), 0.0, 1.0, None, C.byref(fb[0]),
                                                                       C.byref(fb[1]), C.byref(fb[2]), d3(DXINV), 1, s)), 48.0 * N)
del tp, tr, to, tb, ta, tphi, tphi2, trhs, tsig, tout
torch.cuda.empty_cache()

# ---- Godunov ----
g = ix.Geom.make(cells)
tS, fS = fab(cells, 3, 3)
tS.mul_(0.5)
tF, fF = fab(cells, 1, 3)
tum = [fab(tuple(m + (1 if d == q else 0) for q, m in enumerate(cells)), 1) for d in range(3)]
for t, _ in tum:
    t.sub_(0.5)
fum = [f for _, f in tum]
tA, fA = fab(cells, 0, 3)
icons = (C.c_int * 3)(0, 0, 0)
dt = 0.7 / n
if spec:   # K2 of SURVEY.md 8d: Taylor-Green velocities at cell and face centres, zero forcing
    import spec_inputs as si
    k2 = si.k2_advection(n, dev)
    tS.copy_(si.ghosted(k2["vel"], 3)); tF.zero_()
    for d, key in enumerate(("umac", "vmac", "wmac")):
        tum[d][0].copy_(si.ghosted(k2[key], 1, tuple(1 if q == d else 0 for q in range(3))))
    del k2
timeit("ComputeAofs 3 comps (velocity)",
       lambda: lib.check(lib.iamrx_compute_aofs_box(C.byref(bx), C.byref(fA), 0, C.byref(fS), 0, 3, C.byref(fF), 0, None, C.byref(fum[0]), C.byref(fum[1]),
                                                    C.byref(fum[2]), None, None, None, None, None, None, None, None, None, icons, None, C.byref(g), dt, ix.ADV_IS_VELOCITY if hasattr(ix, 'ADV_IS_VELOCITY') else 4, s)),
       104.0 * N, reps=5)
timeit("ComputeAofs 3 comps staged kernels",
       lambda: lib.check(lib.iamrx_compute_aofs_box(C.byref(bx), C.byref(fA), 0, C.byref(fS), 0, 3, C.byref(fF), 0, None, C.byref(fum[0]), C.byref(fum[1]),
                                                    C.byref(fum[2]), None, None, None, None, None, None, None, None, None, icons, None, C.byref(g), dt, 4 | 32, s)),
       104.0 * N, reps=3)
tmac = [fab(tuple(m + (1 if d == q else 0) for q, m in enumerate(cells)), 1) for d in range(3)]
fmac = [f for _, f in tmac]
timeit("ExtrapVelToFaces",
       lambda: lib.check(lib.iamrx_extrap_vel_to_faces_box(C.byref(bx), C.byref(fS), C.byref(fF), C.byref(fmac[0]), C.byref(fmac[1]), C.byref(fmac[2]),
                                                           None, C.byref(g), dt, 0, s)),
       72.0 * N, reps=5)
