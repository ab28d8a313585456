"""Diagnostic: 2-D DoubleShearLayer as a two-layer box, multigrid histories (usage: dsl2d_diag.py N [verbose])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iamr_b200 as ix
lib = ix.load(); dev = "cuda:0"
N = int(sys.argv[1]); v = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = (N, N, 2)
g = ix.Geom.make(n, (-1., -1., -1.), (1., 1., 1.))
lev = ix.Level(lib, g, [((0, 0, 0), (N - 1, N - 1, 1))])
ns = ix.NavierStokes(lib, lev, dev, visc_coef=0.0, cfl=0.5, mg_verbose=v)
ns.init_prob(5, [1.0, 1.0, 0.0, 0.0, 0.0, 0.4])
try:
    print(ns.post_init()); print(ns.step(), ns.last_iters())
except Exception as e:
    print("FAILED", e)
