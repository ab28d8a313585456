"""Launch one kernel family a few times at 256^3 (for ncu captures).  Usage: python scripts/kb_one.py <gs_sweep|adotx|gsrb|gsrb_sweep|aofs|extrap> [n] [reps]"""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import iamr_b200 as ix
from util import box_of, d3, stream_of

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = 'cuda:0'
lib = ix.load()
s = stream_of(dev)
DXINV = (float(n),) * 3
cells, nodes = (n, n, n), (n + 1, n + 1, n + 1)
bx, nbx = box_of((0, 0, 0), (n - 1, n - 1, n - 1)), box_of((0, 0, 0), (n, n, n))


def fab(shape_valid, ng, ncomp=1):
    t = torch.rand((ncomp,) + tuple(m + 2 * ng for m in reversed(shape_valid)), dtype=torch.float64, device=dev)
    return t, ix.fab_of(t, [-ng] * 3)


def faces(ng):
    return [fab(tuple(m + (1 if d == q else 0) for q, m in enumerate(cells)), ng) for d in range(3)]


if what in ('gs_sweep', 'adotx'):
    tphi, fphi = fab(nodes, 1); tphi2, fphi2 = fab(nodes, 1); trhs, frhs = fab(nodes, 1); tsig, fsig = fab(cells, 1)
    tsig.add_(1.0)
    for _ in range(reps):
        if what == 'gs_sweep':
            lib.check(lib.iamrx_nodal_gs_sweep_box(C.byref(nbx), C.byref(fphi2), C.byref(fphi), C.byref(frhs), C.byref(fsig), d3(DXINV), s))
        else:
            lib.check(lib.iamrx_nodal_adotx_box(C.byref(nbx), C.byref(fphi2), C.byref(fphi), C.byref(frhs), C.byref(fsig), d3(DXINV), s))
elif what in ('gsrb', 'gsrb_sweep'):
    tp, fp = fab(cells, 1); tp2, fp2 = fab(cells, 1); tr, fr = fab(cells, 0)
    tb = faces(0)
    for t, _ in tb:
        t.add_(0.5)
    fb = [f for _, f in tb]
    for _ in range(reps):
        if what == 'gsrb':
            for rb in range(2):
                lib.check(lib.iamrx_abec_gsrb_box(C.byref(bx), C.byref(fp), C.byref(fr), 0.0, 1.0, None, C.byref(fb[0]), C.byref(fb[1]), C.byref(fb[2]),
                                                  d3(DXINV), 1.15, rb, 1, s))
        else:
            lib.check(lib.iamrx_abec_gsrb_sweep_box(C.byref(bx), C.byref(fp2), C.byref(fp), C.byref(fr), 0.0, 1.0, None, C.byref(fb[0]), C.byref(fb[1]),
                                                    C.byref(fb[2]), d3(DXINV), 1.15, 1, s))
else:
    g = ix.Geom.make(cells)
    tS, fS = fab(cells, 3, 3); tS.mul_(0.5)
    tF, fF = fab(cells, 1, 3)
    tum = faces(1)
    for t, _ in tum:
        t.sub_(0.5)
    fum = [f for _, f in tum]
    dt = 0.7 / n
    if what == 'aofs':
        tA, fA = fab(cells, 0, 3)
        icons = (C.c_int * 3)(0, 0, 0)
        for _ in range(reps):
            lib.check(lib.iamrx_compute_aofs_box(C.byref(bx), C.byref(fA), 0, C.byref(fS), 0, 3, C.byref(fF), 0, None, C.byref(fum[0]), C.byref(fum[1]),
                                                 C.byref(fum[2]), None, None, None, None, None, None, None, None, None, icons, None, C.byref(g), dt, 4, s))
    else:
        tm = faces(1)
        fm = [f for _, f in tm]
        for _ in range(reps):
            lib.check(lib.iamrx_extrap_vel_to_faces_box(C.byref(bx), C.byref(fS), C.byref(fF), C.byref(fm[0]), C.byref(fm[1]), C.byref(fm[2]), None, C.byref(g), dt, 0, s))
torch.cuda.synchronize()
print("done", what)
