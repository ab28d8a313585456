"""Quick GPU probe: Taylor-Green on the CUDA library, per-step timings and analytic error."""
import math, sys, time
import torch
sys.path.insert(0, '.')
import iamr_b200 as ix

lib = ix.load()
dev = 'cuda:0'
torch.cuda.init()
for n, nsteps in [(32, 3), (64, 3), (128, 3), (256, 3)]:
    g = ix.Geom.make((n, n, n))
    lev = ix.Level(lib, g, [((0, 0, 0), (n - 1, n - 1, n - 1))])
    ns = ix.NavierStokes(lib, lev, dev, visc_coef=1e-4, cfl=0.7, verbose=0, mg_verbose=0)
    ns.init_prob(11, [1.0, 1.0, 0.0, 1.0, 1.0])
    torch.cuda.synchronize(); t0 = time.time()
    dt0 = ns.post_init()
    torch.cuda.synchronize(); print(f'n={n} dt0={dt0:.6e} init {time.time()-t0:.3f}s', flush=True)
    for s in range(nsteps):
        lib.iamrx_launch_count_reset()
        torch.cuda.synchronize(); t0 = time.time()
        dt = ns.step()
        torch.cuda.synchronize(); w = time.time() - t0
        print(f'  step {s} dt {dt:.6e} iters {ns.last_iters()} wall {w*1e3:.1f} ms launches {lib.iamrx_launch_count()} '
              f'Mcells/s {n**3/w/1e6:.1f}', flush=True)
    S = ns.field(0)
    t = ns.time
    x = (torch.arange(n, dtype=torch.float64, device=dev) + 0.5) / n
    X = x.view(1, 1, n); Y = x.view(1, n, 1)
    dec = math.exp(-8 * math.pi ** 2 * 1e-4 * t)
    ue = torch.sin(2 * math.pi * X) * torch.cos(2 * math.pi * Y) * dec
    print(f'  time {t:.5f} err u {(S[0]-ue).abs().max().item():.3e} w {S[2].abs().max().item():.2e} '
          f'mem {torch.cuda.memory_allocated()/1e9:.2f} GB(torch)', flush=True)
    ns.close(); lev.close()
