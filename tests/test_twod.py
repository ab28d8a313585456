"""2-D problems (Exec/run2d, Tutorials/DoubleShearLayer) run as TWO-LAYER 3-D problems: n = (nx, ny, 2), periodic and uniform in z.
With w = 0 and no z variation the corner-coupled Godunov scheme, the 7-point operators and the Q1 nodal operator reduce exactly
to their 2-D forms; the multigrid hierarchies never coarsen the thin direction (semi-coarsening, mlmg.cu thin_mask) and the
ghost-cell plans reach across more than one period.  The layers are made THICK (z extent = x extent) so that the coupling across
them never dominates the point smoothers.  Reference for every run: the 3-D oracle on an n^3 box with the
same z-uniform initial data -- same answer, independent code path, full coarsening."""
import numpy as np
import pytest

import iamr_b200 as ix
from util import split_boxes

LID = [[0.0] * 5 for _ in range(6)]
LID[4][0] = 1.0   # yhi.velocity = 1 0 0: the lid of the 2-D cavity (Tutorials/LidDrivenCavity/inputs.2d.lid_driven_cavity)

RUNS = [
    # TaylorGreen with prob.c = 0: the 2-D Taylor vortex (inputs.2d.taylorgreen)
    dict(probtype=11, pp=[1.0, 1.0, 0.0, 1.0, 1.0], lo=(0, 0, 0), hi=(1, 1, 1), per=(1, 1, 1), kw=dict(visc_coef=1e-3, cfl=0.7), ncmp=5),
    # DoubleShearLayer (inputs.2d.double_shear_layer), conservative tracer; the 3-D blob is a sphere, so the tracer is not compared
    dict(probtype=5, pp=[1.0, 1.0, 0.0, 0.0, 0.0, 0.4], lo=(-1, -1, -1), hi=(1, 1, 1), per=(1, 1, 1),
         kw=dict(visc_coef=1e-3, cfl=0.7, conservative_tracer=1), ncmp=4),
    # 2-D lid-driven cavity: no-slip walls in x and y, moving lid at y-hi
    dict(probtype=1, pp=[0.0], lo=(0, 0, 0), hi=(1, 1, 1), per=(0, 0, 1), lo_bc=(5, 5, 0), hi_bc=(5, 5, 0), bcv=LID,
         kw=dict(visc_coef=0.01, cfl=0.7, init_shrink=0.3, init_iter=3, fixed_dt=0.0140625), ncmp=5),
]


@pytest.mark.parametrize("run", RUNS, ids=["taylor_vortex_2d", "double_shear_layer_2d", "lid_driven_cavity_2d"])
@pytest.mark.parametrize("nb", [(1, 1, 1), (2, 2, 1)])
def test_two_layer_run_matches_z_uniform_oracle(backend, oracle, run, nb):
    lib, dev = backend
    N = 32 if run["per"] == (1, 1, 1) else 16
    lo, hi, per = run["lo"], run["hi"], run["per"]
    wall = dict(lo_bc=run["lo_bc"], hi_bc=run["hi_bc"]) if "lo_bc" in run else {}
    okw = dict(per=per, phys_lo=run["lo_bc"], phys_hi=run["hi_bc"], bcv=run.get("bcv")) if wall else {}
    o = oracle.OracleNS((N, N, N), lo, hi, **okw, **run["kw"])
    o.init_prob(run["probtype"], run["pp"])
    dto = [o.post_init()] + [o.step() for _ in range(3)]
    So = o.get(0)
    assert np.abs(So[2]).max() < 1e-12                                      # w stays zero
    assert np.abs(So[:4] - So[:4, :1]).max() < 1e-10                        # and the 3-D solution stays z-uniform
    n = (N, N, 2)
    g = ix.Geom.make(n, lo, hi, periodic=per)                                 # two layers as thick as half the box
    boxes = split_boxes(n, nb)
    lev = ix.Level(lib, g, boxes)
    ns = ix.NavierStokes(lib, lev, dev, **wall, **(dict(bc_vals=run["bcv"]) if "bcv" in run else {}), **run["kw"])
    ns.init_prob(run["probtype"], run["pp"])
    dts = [ns.post_init()] + [ns.step() for _ in range(3)]
    assert np.allclose(dts, dto, rtol=1e-10, atol=0)
    err = 0.0
    c = run["ncmp"]
    for il, (blo, bhi) in enumerate(boxes):
        t = ns.field(0, il).cpu().numpy()
        ny, nx = bhi[1] - blo[1] + 1, bhi[0] - blo[0] + 1
        ref = So[:c, :1, blo[1]:bhi[1] + 1, blo[0]:bhi[0] + 1]
        err = max(err, np.abs(t[:c, :2, :ny, :nx] - ref).max())               # both layers against the oracle's (z-uniform) plane
    assert err <= 1e-10
    it = ns.last_iters()
    assert it[0] <= 14 and it[2] <= 14                                        # multigrid convergence is not degraded by the thin direction
    ns.close(); o.close(); lev.close()
