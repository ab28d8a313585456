"""Committed fixtures (tests/golden/, generator make_golden.py):
  * the analytic Taylor vortex of the reference's Tutorials/TaylorGreen/benchmarks/EXACT_3D.F -- the only known answer
    the reference holds for this path; pins the oracle (CPU) and the CUDA path (GPU) to the O(h^2) truncation error;
  * a snapshot of the oracle itself, which freezes the checker the CUDA parity tests compare against."""
import math
import os

import numpy as np
import pytest

import iamr_b200 as ix

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("n,tol", [(16, 2.5e-2), (32, 6.5e-3)])
def test_oracle_matches_exact_taylor_vortex(oracle, n, tol):
    g = np.load(os.path.join(GOLD, "taylor_vortex_exact.npz"))
    o = oracle.OracleNS((n, n, n), visc_coef=float(g["nu"]), cfl=0.7)
    o.init_prob(11, [1.0, 1.0, 0.0, 1.0, 1.0])
    o.post_init()
    t = 0.0
    while t < float(g["time"]):
        t += o.step()
    S = o.get(0)
    # the fixture is sampled at t = 0.05; the run overshoots by less than one step: rescale the decay factor
    ex = g[f"state_{n}"].copy()
    ex[:2] *= math.exp(-8 * math.pi ** 2 * float(g["nu"]) * (t - float(g["time"])))
    assert np.abs(S[:2] - ex[:2]).max() < tol                 # O(h^2): 4x smaller at 32 than at 16
    # rho sees dt * div(u_mac), which the MAC solve drives to mac_tol * |rhs| (1e-12 relative), not to rounding
    assert np.abs(S[2]).max() < 1e-13 and np.abs(S[3] - 1.0).max() < 2e-12
    o.close()


def test_oracle_snapshot_is_frozen(oracle):
    g = np.load(os.path.join(GOLD, "oracle_snapshot_16.npz"))
    o = oracle.OracleNS((16, 16, 16), visc_coef=1e-3, cfl=0.7, gravity=-0.5)
    o.init_prob(100, [1.0, 1.0, 1.0, 1.0, 1.0])
    dts = [o.post_init()] + [o.step() for _ in range(2)]
    assert np.allclose(dts, g["dts"], rtol=1e-12, atol=0)
    # identical source + compiler flags reproduce it bit for bit; allow for a different OpenMP reduction order
    assert np.abs(o.get(0) - g["state"]).max() < 1e-11
    assert np.abs(o.get(1) - g["press"]).max() < 1e-9
    o.close()


@pytest.mark.gpu
def test_cuda_matches_exact_taylor_vortex_and_snapshot(cuda_lib):
    import torch
    lib, dev = cuda_lib, "cuda:0"
    g = np.load(os.path.join(GOLD, "taylor_vortex_exact.npz"))
    n = 32
    lev = ix.Level(lib, ix.Geom.make((n, n, n)), [((0, 0, 0), (n - 1, n - 1, n - 1))])
    ns = ix.NavierStokes(lib, lev, dev, visc_coef=float(g["nu"]), cfl=0.7)
    ns.init_prob(11, [1.0, 1.0, 0.0, 1.0, 1.0])
    ns.post_init()
    while ns.time < float(g["time"]):
        ns.step()
    S = ns.field(0).cpu().numpy()
    ex = g[f"state_{n}"].copy()
    ex[:2] *= math.exp(-8 * math.pi ** 2 * float(g["nu"]) * (ns.time - float(g["time"])))
    assert np.abs(S[:2] - ex[:2]).max() < 6.5e-3
    ns.close(); lev.close()
    # the frozen oracle snapshot: two boxes on the GPU against the committed CPU result
    s = np.load(os.path.join(GOLD, "oracle_snapshot_16.npz"))
    boxes = [((0, 0, 0), (15, 15, 7)), ((0, 0, 8), (15, 15, 15))]
    lev = ix.Level(lib, ix.Geom.make((16, 16, 16)), boxes)
    ns = ix.NavierStokes(lib, lev, dev, visc_coef=1e-3, cfl=0.7, gravity=-0.5)
    ns.init_prob(100, [1.0, 1.0, 1.0, 1.0, 1.0])
    dts = [ns.post_init()] + [ns.step() for _ in range(2)]
    assert np.allclose(dts, s["dts"], rtol=1e-11, atol=0)
    for il, (lo, hi) in enumerate(boxes):
        t = ns.field(0, il).cpu().numpy()
        assert np.abs(t - s["state"][:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]).max() <= 1e-10
    ns.close(); lev.close()
    torch.cuda.synchronize()
