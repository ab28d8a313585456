"""RayleighTaylor 3-D single level (BASELINE.json configs[3] geometry, one level): 256 x 256 x 512 cells on one GPU, periodic x/y,
slip walls in z, gravity; per-step time, per-kernel table, MG iteration counts.  Diagnostic (not the bench.py contract)."""
import ctypes as C
import sys
import time
import torch
sys.path.insert(0, '.')
import iamr_b200 as ix

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = (nx, nx, 2 * nx)
lib = ix.load()
dev = 'cuda:0'
g = ix.Geom.make(n, (0.0, 0.0, 0.0), (0.5, 0.5, 1.0), periodic=(1, 1, 0))
lev = ix.Level(lib, g, [((0, 0, 0), (n[0] - 1, n[1] - 1, n[2] - 1))])
ns = ix.NavierStokes(lib, lev, dev, lo_bc=(0, 0, 4), hi_bc=(0, 0, 4), visc_coef=0.0, cfl=0.7, gravity=-1.0)
ns.init_prob(10, [1.0, 2.0, 1.0, 0.0, 0.01, 0.005])
t0 = time.perf_counter(); ns.post_init(); torch.cuda.synchronize()
print(f"post_init {time.perf_counter() - t0:.2f} s")
for _ in range(2):
    ns.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    ns.step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps
cells = n[0] * n[1] * n[2]
print(f"RayleighTaylor {n}: {wall * 1e3:.2f} ms/step, {cells / wall / 1e6:.1f} Mcells/s, iters {ns.last_iters()}")
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
for ms, cnt, name in rows[:25]:
    print(f"{name:34s} {cnt:7d} {ms:10.3f} ms {100 * ms / tot:5.1f}%")
S = ns.field(0)
print("max |u|", [float(S[c].abs().max()) for c in range(3)], "mass", float(S[3].sum()))
