"""Diagnostic: HIT (probtype 20) post_init at 512^3 on ONE GPU as 2x2x2 boxes of 256^3, multigrid history printed (mg_verbose 2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import torch, iamr_b200 as ix
from util import split_boxes
lib = ix.load(); dev = "cuda:0"
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
vd = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
n = (2 * nb,) * 3
g = ix.Geom.make(n, (0, 0, 0), (2.0, 2.0, 2.0))
lev = ix.Level(lib, g, split_boxes(n, (2, 2, 2)))
ns = ix.NavierStokes(lib, lev, dev, visc_coef=1e-4, cfl=0.7, proj_tol=1e-10, mac_tol=float(sys.argv[3]) if len(sys.argv) > 3 else 1e-12, mg_verbose=int(sys.argv[4]) if len(sys.argv) > 4 else 2)
ns.init_prob(20, [1.0, 1.0, vd])
print("mem GB", torch.cuda.memory_allocated() / 1e9, flush=True)
try:
    print("post_init dt", ns.post_init())
    for _ in range(3):
        print("step", ns.step(), ns.last_iters())
except Exception as e:
    print("FAILED", e)
