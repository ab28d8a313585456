"""Generates the fixtures in this directory.  Run from the repo root: python tests/golden/make_golden.py

1. taylor_vortex_exact.npz -- the reference's ONLY in-repo known answer for the path: the analytic decaying Taylor
   vortex of Tutorials/TaylorGreen/benchmarks/EXACT_3D.F:75-118 (unifdir = 2: u = sin 2pi x cos 2pi y e^{-8 pi^2 nu t},
   v = -cos 2pi x sin 2pi y e^{-8 pi^2 nu t}, w = 0, rho = 1), sampled at cell centres exactly as that routine does
   (x = xlo + hx (i - lo + 1/2)), for the inputs.3d.taylorgreen parameters (nu = 1e-4) at t = 0.05, n = 16 and 32.
   It pins results to the O(h^2) truncation error only ("parity unpinned" at 1e-10).
2. oracle_snapshot_16.npz -- the CPU oracle's own state after post_init + 2 steps of a variable-density 16^3 problem.
   NOT a reference output: it freezes the checker, so that an accidental change to oracle/ shows up as a failure
   instead of silently moving the target the CUDA path is compared with.
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def exact_taylor(n, t, nu):
    x = (np.arange(n) + 0.5) / n
    Z, Y, X = np.meshgrid(x, x, x, indexing="ij")
    e = math.exp(-2.0 * 4.0 * math.pi ** 2 * nu * t)
    s = np.zeros((4, n, n, n))
    s[0] = np.sin(2 * math.pi * X) * np.cos(2 * math.pi * Y) * e
    s[1] = -np.cos(2 * math.pi * X) * np.sin(2 * math.pi * Y) * e
    s[3] = 1.0
    return s


def main():
    out = {"nu": 1.0e-4, "time": 0.05}
    for n in (16, 32):
        out[f"state_{n}"] = exact_taylor(n, 0.05, 1.0e-4)
    np.savez_compressed(os.path.join(HERE, "taylor_vortex_exact.npz"), **out)

    import orc
    kw = dict(visc_coef=1e-3, cfl=0.7, gravity=-0.5)
    o = orc.OracleNS((16, 16, 16), **kw)
    o.init_prob(100, [1.0, 1.0, 1.0, 1.0, 1.0])
    dts = [o.post_init()]
    for _ in range(2):
        dts.append(o.step())
    np.savez_compressed(os.path.join(HERE, "oracle_snapshot_16.npz"), state=o.get(0), press=o.get(1), dts=np.array(dts),
                        iters=np.array(o.last_iters()) if hasattr(o, "last_iters") else np.zeros(3))
    o.close()
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
