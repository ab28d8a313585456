"""Time one kernel family at n^3 with CUDA events (like kbench.py, one line).  Usage: python scripts/kb_time.py adotx [n] [reps]"""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import iamr_b200 as ix
from util import box_of, d3, stream_of

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = 'cuda:0'
lib = ix.load()
s = stream_of(dev)
DXINV = (float(n),) * 3


def fab(shape_valid, ng, ncomp=1):
    t = torch.rand((ncomp,) + tuple(m + 2 * ng for m in reversed(shape_valid)), dtype=torch.float64, device=dev)
    return t, ix.fab_of(t, [-ng] * 3)


nodes, cells = (n + 1,) * 3, (n,) * 3
nbx = box_of((0, 0, 0), (n, n, n))
if what == 'adotx':
    tphi, fphi = fab(nodes, 1); tout, fout = fab(nodes, 1); trhs, frhs = fab(nodes, 1); tsig, fsig = fab(cells, 1)
    fn = lambda: lib.check(lib.iamrx_nodal_adotx_box(C.byref(nbx), C.byref(fout), C.byref(fphi), C.byref(frhs), C.byref(fsig), d3(DXINV), s))
    nbytes = 32.0 * (n + 1) ** 3
elif what == 'gs_sweep':
    tphi, fphi = fab(nodes, 1); tout, fout = fab(nodes, 1); trhs, frhs = fab(nodes, 1); tsig, fsig = fab(cells, 1)
    tsig.add_(1.0)
    fn = lambda: lib.check(lib.iamrx_nodal_gs_sweep_box(C.byref(nbx), C.byref(fout), C.byref(fphi), C.byref(frhs), C.byref(fsig), d3(DXINV), s))
    nbytes = 32.0 * (n + 1) ** 3
elif what in ('gsrb', 'gsrb_sweep'):
    tp, fp = fab(cells, 1); tp2, fp2 = fab(cells, 1); tr, fr = fab(cells, 0)
    tb = [fab(tuple(m + (1 if d == q else 0) for q, m in enumerate(cells)), 0) for d in range(3)]
    for t, _ in tb:
        t.add_(0.5)
    fb = [f for _, f in tb]
    bx = box_of((0, 0, 0), (n - 1, n - 1, n - 1))
    if what == 'gsrb':
        def fn():
            for rb in range(2):
                lib.check(lib.iamrx_abec_gsrb_box(C.byref(bx), C.byref(fp), C.byref(fr), 0.0, 1.0, None, C.byref(fb[0]), C.byref(fb[1]), C.byref(fb[2]),
                                                  d3(DXINV), 1.15, rb, 1, s))
    else:
        fn = lambda: lib.check(lib.iamrx_abec_gsrb_sweep_box(C.byref(bx), C.byref(fp2), C.byref(fp), C.byref(fr), 0.0, 1.0, None, C.byref(fb[0]),
                                                             C.byref(fb[1]), C.byref(fb[2]), d3(DXINV), 1.15, 1, s))
    nbytes = 96.0 * n ** 3
else:
    raise SystemExit("unknown kernel")
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
print(f"{what} {us:9.1f} us  {nbytes / us / 1e3:8.1f} GB/s (algorithmic)")
