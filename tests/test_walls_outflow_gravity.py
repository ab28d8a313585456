"""Outflow boundaries under gravity: Projection::set_outflow_bcs -> computeRhoG (Projection.cpp:1721-2379) in the step
driver -- the hydrostatic pressure on the nodes of outflow faces, in initialPressureProject and in every level_project -- against the
oracle's twin and against a known answer.  (A separate file, late in pytest's order: these tests were added after the round's GPU
budget was spent and have only run through the host-emulation build.)"""
import numpy as np
import pytest

import iamr_b200 as ix
from test_bc import INFLOW, run_walls_case

OUTFLOW_GRAVITY_RUNS = [
    # the channel under gravity with a variable density: Projection::set_outflow_bcs puts the hydrostatic pressure of the density
    # next to the outflow face on its nodes (computeRhoG, Projection.cpp:1933-2379; transverse direction periodic), in
    # initialPressureProject and in every level_project
    dict(n=(32, 16, 16), hi=(2.0, 1.0, 1.0), per=(0, 1, 0), lo_bc=(1, 0, 4), hi_bc=(2, 0, 4), probtype=101, pp=[1.0, 1.0, 0.3], bcv=INFLOW, iters_slack=1,
         kw=dict(visc_coef=0.01, cfl=0.5, gravity=-0.5)),
    # outflow faces in x (high) and y (low) under gravity whose transverse edges meet an inflow side (ext_dir density: the ghost
    # column) and an outflow / wall side (foextrap: the first column); the two planes share a vertical edge
    dict(n=(16, 16, 16), hi=(1.0, 1.0, 1.0), per=(0, 0, 0), lo_bc=(1, 2, 4), hi_bc=(2, 5, 4), probtype=101, pp=[1.0, 1.0, 0.3],
         bcv=[[0.3, 0.0, 0.0, 1.2, 0.5], [0.0] * 5, [0.0] * 5, [0.0] * 5, [0.0] * 5, [0.0] * 5], iters_slack=1,
         kw=dict(visc_coef=0.01, cfl=0.5, gravity=-0.5)),
]


@pytest.mark.parametrize("run", OUTFLOW_GRAVITY_RUNS, ids=["channel_gravity_outflow", "two_outflow_faces_gravity"])
@pytest.mark.parametrize("nb", [(1, 1, 1), (2, 2, 2)])
def test_step_with_outflow_under_gravity_matches_oracle(backend, oracle, run, nb):
    """post_init + 3 steps: velocity, scalars, pressure and grad(p) against the oracle, L-inf <= 1e-10."""
    run_walls_case(backend, oracle, run, nb)


def test_hydrostatic_equilibrium_with_an_outflow_side(backend):
    """Known answer (no oracle needed): a flat stratification at rest in a box with an OUTFLOW side face.  Projection::set_outflow_bcs
    must put exactly the discrete hydrostatic pressure on the nodes of that face (computeRhoG: phi(k) = -g dz sum of the densities
    of the cell layers above; the extrapolation (3 rho_1 - rho_2) / 2 of a horizontally uniform density is that density) -- with
    phi = 0 there instead, the projection would drive a flow through the face.  The fluid stays at rest to round-off."""
    lib, dev = backend
    n = (8, 8, 32)
    g = ix.Geom.make(n, (0.0, 0.0, 0.0), (0.25, 0.25, 1.0), periodic=(0, 1, 0))
    lev = ix.Level(lib, g, [((0, 0, 0), (7, 7, 15)), ((0, 0, 16), (7, 7, 31))])
    ns = ix.NavierStokes(lib, lev, dev, lo_bc=(4, 0, 4), hi_bc=(2, 0, 4), visc_coef=0.0, gravity=-1.0, fixed_dt=0.01)
    ns.init_prob(10, [1.0, 2.0, 1.0, 0.0, 0.05, 0.0])   # perturbation amplitude 0
    ns.post_init()
    for _ in range(3):
        ns.step()
    for il in range(2):
        S, G, P = ns.field(0, il), ns.field(2, il), ns.field(1, il)
        assert float(S[:3].abs().max()) < 1e-11
        assert float((G[2] + S[3]).abs().max()) < 1e-10   # dp/dz = rho * g with g = -1
    assert float(ns.field(1, 0).abs().max()) > 0.5        # the pressure is hydrostatic, not zero, on the outflow face too
    ns.close(); lev.close()
