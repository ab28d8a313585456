"""TaylorGreen on ONE GPU with the domain split into boxes (an AMR-like layout: no in-kernel wrap, ghost exchange between
boxes): per-step time, per-kernel table, MG iteration counts.  Diagnostic (not the bench.py contract).
usage: multibox_bench.py [n=256] [nbx nby nbz = 2 2 2] [steps=5]"""
import ctypes as C
import sys
import time
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iamr_b200 as ix
from util import split_boxes

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nb = tuple(int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (2, 2, 2)
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
n = (nx, nx, nx)
lib = ix.load()
dev = 'cuda:0'
g = ix.Geom.make(n)
lev = ix.Level(lib, g, split_boxes(n, nb))
ns = ix.NavierStokes(lib, lev, dev, visc_coef=1e-4, cfl=0.7)
ns.init_prob(11, [1.0, 1.0, 1.0, 1.0, 1.0])
ns.post_init()
for _ in range(2):
    ns.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    ns.step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps
print(f"TaylorGreen {n} as {nb} boxes: {wall * 1e3:.2f} ms/step, {nx ** 3 / wall / 1e6:.1f} Mcells/s, iters {ns.last_iters()}")
lib.iamrx_prof_all(1)
ns.step(); torch.cuda.synchronize()
lib.iamrx_prof_all(0)
buf = C.create_string_buffer(1 << 16)
lib.iamrx_prof_dump(buf, len(buf))
rows = []
for line in buf.value.decode().splitlines():
    name, cnt, ms = line.rsplit(' ', 2)
    rows.append((float(ms), int(cnt), name.replace(' ', '')))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
for ms, cnt, name in rows[:14]:
    print(f"{name:34s} {cnt:7d} {ms:10.3f} ms {100 * ms / tot:5.1f}%")
S = ns.field(0)
print("checksum max|u|", float(S[0].abs().max()), "sum u^2", float((S[0] ** 2).sum()))
