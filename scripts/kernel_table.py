"""Per-kernel time table of the timed step (CUDA events around every launch, iamrx_prof_all):
which kernels the step spends its time in.  Usage: python scripts/kernel_table.py [n] [steps]"""
import ctypes as C
import sys
import time
import torch
sys.path.insert(0, '.')
import iamr_b200 as ix

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nzb = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # number of z slabs (boxes) the domain is cut into, all on this GPU
lib = ix.load()
dev = 'cuda:0'
g = ix.Geom.make((n, n, n))
lev = ix.Level(lib, g, [((0, 0, b * (n // nzb)), (n - 1, n - 1, (b + 1) * (n // nzb) - 1)) for b in range(nzb)])
ns = ix.NavierStokes(lib, lev, dev, visc_coef=1e-4, cfl=0.7)
ns.init_prob(11, [1.0, 1.0, 0.0, 1.0, 1.0])
ns.post_init()
for _ in range(2):
    ns.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    ns.step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps * 1e3
lib.iamrx_prof_all(1)
lib.iamrx_launch_count_reset()
for _ in range(steps):
    ns.step()
torch.cuda.synchronize()
lib.iamrx_prof_all(0)
buf = C.create_string_buffer(1 << 16)
lib.iamrx_prof_dump(buf, len(buf))
rows = []
for line in buf.value.decode().splitlines():
    name, cnt, ms = line.rsplit(' ', 2)
    name = name.replace(' ', '')
    rows.append((float(ms) / steps, int(cnt) // steps, name))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"# TaylorGreen {n}^3, per step: wall {wall:.2f} ms (untimed run), sum of kernel times {tot:.2f} ms, "
      f"launches {lib.iamrx_launch_count() // steps}, iters {ns.last_iters()}")
print(f"{'kernel':28s} {'launches':>9s} {'ms/step':>10s} {'share':>7s}")
for ms, cnt, name in rows:
    print(f"{name:28s} {cnt:9d} {ms:10.3f} {100 * ms / tot:6.1f}%")
