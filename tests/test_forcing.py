"""The HIT tutorial's turbulent forcing (Tutorials/HIT/NS_getForce.cpp:205-640, TurbulentForcing_def.H): the per-box kernel against a
numpy restatement of the reference's exact path, and the forced time step against the oracle."""
import ctypes as C

import numpy as np
import pytest

import iamr_b200 as ix
from util import split_boxes, to_fab, from_fabs, stream_of, sync

AS = 33   # TurbulentForcing::array_size


def make_forcedata(L, nmodes, mode_start, div_free, seed=111397):
    """A table with the structure TurbulentForcing::init_turbulent_forcing builds (TurbulentForcing_def.H:141-230): frequencies,
    phases and amplitudes on the active modes, spectrum_type 2, moderate_zero_modes.  (numpy's generator, not DepRand: parity is
    kernel vs restatement on the SAME table; the table itself is an input of the ABI.)"""
    rng = np.random.default_rng(seed)
    fd = np.zeros((17, AS, AS, AS))
    Lmin = min(L)
    step = [int(l / Lmin + 0.5) for l in L]
    kmax = nmodes / Lmin + 1e-8
    def fill(kx, ky, kz):
        kappa = np.sqrt((kx / L[0]) ** 2 + (ky / L[1]) ** 2 + (kz / L[2]) ** 2)
        if kappa > kmax:
            return
        fd[0, kz, ky, kx] = (1.0 + 1.0 * rng.random()) * 2 * np.pi
        fd[1:5, kz, ky, kx] = rng.random(4) * 2 * np.pi
        fd[8:17, kz, ky, kx] = rng.random(9) * 2 * np.pi
        th, ph = rng.random() * 2 * np.pi, rng.random() * np.pi
        p = np.array([np.cos(th) * np.sin(ph), np.sin(th) * np.sin(ph), np.cos(ph)])
        if kappa < 1e-6:
            return
        e = 1.0 / kappa ** 2 / (kappa if div_free else 1.0)
        for kk in (kx, ky, kz):
            if kk == 0:
                e /= 2.0
        fd[5:8, kz, ky, kx] = p * e / (p @ p)
    for kz in range(mode_start * step[2], nmodes * step[2] + 1, step[2]):
        for ky in range(mode_start * step[1], nmodes * step[1] + 1, step[1]):
            for kx in range(mode_start * step[0], nmodes * step[0] + 1, step[0]):
                fill(kx, ky, kz)
    for kz in range(1, step[2]):
        for ky in range(mode_start, nmodes * step[1] + 1):
            for kx in range(mode_start, nmodes * step[0] + 1):
                fill(kx, ky, kz)
    return fd


def numpy_force(n, lo, L, t, fd, nmodes, mode_start, div_free):
    """f(x, t) at the cell centres: NS_getForce.cpp:541-621 (the exact path), vectorised over cells."""
    x, y, z = [lo[d] + L[d] / n[d] * (np.arange(n[d]) + 0.5) for d in range(3)]
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    f = np.zeros((3,) + X.shape)
    Lmin = min(L)
    step = [int(l / Lmin + 0.5) for l in L]
    kmax = nmodes / Lmin + 1e-8
    tp = 2 * np.pi
    modes = [(kx, ky, kz) for kz in range(mode_start * step[2], nmodes * step[2] + 1, step[2])
             for ky in range(mode_start * step[1], nmodes * step[1] + 1, step[1])
             for kx in range(mode_start * step[0], nmodes * step[0] + 1, step[0])]
    modes += [(kx, ky, kz) for kz in range(1, step[2]) for ky in range(mode_start, nmodes * step[1] + 1)
              for kx in range(mode_start, nmodes * step[0] + 1)]
    for kx, ky, kz in modes:
        if np.sqrt((kx / L[0]) ** 2 + (ky / L[1]) ** 2 + (kz / L[2]) ** 2) > kmax:
            continue
        g = lambda a: fd[a, kz, ky, kx]
        xT = np.cos(g(0) * t + g(1))
        ax, ay, az = tp * kx * X / L[0], tp * ky * Y / L[1], tp * kz * Z / L[2]
        if div_free:
            f[0] += xT * (g(7) * tp * (ky / L[1]) * np.sin(ax + g(14)) * np.cos(ay + g(15)) * np.sin(az + g(16))
                          - g(6) * tp * (kz / L[2]) * np.sin(ax + g(11)) * np.sin(ay + g(12)) * np.cos(az + g(13)))
            f[1] += xT * (g(5) * tp * (kz / L[2]) * np.sin(ax + g(8)) * np.sin(ay + g(9)) * np.cos(az + g(10))
                          - g(7) * tp * (kx / L[0]) * np.cos(ax + g(14)) * np.sin(ay + g(15)) * np.sin(az + g(16)))
            f[2] += xT * (g(6) * tp * (kx / L[0]) * np.cos(ax + g(11)) * np.sin(ay + g(12)) * np.sin(az + g(13))
                          - g(5) * tp * (ky / L[1]) * np.sin(ax + g(8)) * np.cos(ay + g(9)) * np.sin(az + g(10)))
        else:
            f[0] += xT * g(5) * np.cos(ax + g(2)) * np.sin(ay + g(3)) * np.sin(az + g(4))
            f[1] += xT * g(6) * np.sin(ax + g(2)) * np.cos(ay + g(3)) * np.sin(az + g(4))
            f[2] += xT * g(7) * np.sin(ax + g(2)) * np.sin(ay + g(3)) * np.cos(az + g(4))
    return f


CASES = [((16, 16, 16), (1.0, 1.0, 1.0), 4, 0, 1),    # inputs.3d.forced: turb.nmodes = 4, defaults mode_start 0, div_free_force
         ((16, 16, 16), (1.0, 1.0, 1.0), 3, 1, 0),    # plain (not divergence-free) form, zero modes skipped
         ((8, 8, 16), (1.0, 1.0, 2.0), 2, 0, 1)]      # Lz = 2 Lx: mode steps (1, 1, 2) and the extra symmetry-breaking modes


@pytest.mark.parametrize("n,L,nmodes,mode_start,div_free", CASES)
@pytest.mark.parametrize("nb", [(1, 1, 1), (2, 2, 2)])
def test_turbulent_force_box(backend, n, L, nmodes, mode_start, div_free, nb):
    lib, dev = backend
    lo = (-0.5, -0.5, -0.5)
    fd = make_forcedata(L, nmodes, mode_start, div_free)
    t = 0.37
    ref = numpy_force(n, lo, L, t, fd, nmodes, mode_start, div_free)
    assert np.abs(ref).max() > 1e-2
    x = [(np.arange(m) + 0.5) / m for m in n]
    Zc, Yc, Xc = np.meshgrid(x[2], x[1], x[0], indexing="ij")
    rho = (1.0 + 0.3 * np.sin(2 * np.pi * Xc) * np.cos(2 * np.pi * Yc) * np.sin(2 * np.pi * Zc))[None]
    base = np.stack([0.1 * Xc, -0.2 * Yc, 0.3 * Zc])          # the kernel ACCUMULATES into the force array
    g = ix.Geom.make(n, lo, tuple(lo[d] + L[d] for d in range(3)))
    boxes = split_boxes(n, nb)
    F = [to_fab(base, b, 0, ix.CELL, dev) for b in boxes]
    R = [to_fab(rho, b, 0, ix.CELL, dev) for b in boxes]
    for (blo, bhi), f, r in zip(boxes, F, R):
        bx = ix.Box((C.c_int * 3)(*blo), (C.c_int * 3)(*bhi))
        lib.check(lib.iamrx_turbulent_force_box(C.byref(bx), C.byref(f[1]), C.byref(r[1]), C.byref(g), t, nmodes, mode_start, div_free,
                                                AS, fd.ctypes.data_as(C.POINTER(C.c_double)), stream_of(dev)))
    sync(dev)
    got, _ = from_fabs([f[0] for f in F], boxes, 0, ix.CELL, n, 3)
    assert np.abs(got - (base + rho * ref)).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    if div_free:   # the divergence-free form is the curl of a vector potential: its analytic divergence vanishes; the centred
        h = [L[d] / n[d] for d in range(3)]   # difference of the sampled field is O(h^2) small next to the field's own gradient
        div = sum((np.roll(ref[d], -1, 2 - d) - np.roll(ref[d], 1, 2 - d)) / (2 * h[d]) for d in range(3))
        grad = np.abs((np.roll(ref[0], -1, 2) - np.roll(ref[0], 1, 2)) / (2 * h[0])).max()
        assert np.abs(div).max() < 0.35 * grad


@pytest.mark.parametrize("n,L,nmodes,mode_start,div_free", CASES)
@pytest.mark.parametrize("nb", [(1, 1, 1), (2, 2, 2)])
def test_forced_step_matches_oracle(backend, oracle, n, L, nmodes, mode_start, div_free, nb):
    """HIT initial field + turbulent forcing (USE_TURBULENT_FORCING): post_init + 3 steps vs the oracle, L-inf <= 1e-10.  The force
    enters predict_velocity / velocity_advection at t^n, the velocity update at t^n + dt/2 (with rho_half) and estTimeStep at t^n+1."""
    lib, dev = backend
    lo = (-0.5, -0.5, -0.5)
    hi = tuple(lo[d] + L[d] for d in range(3))
    fd = make_forcedata(L, nmodes, mode_start, div_free)
    boxes = split_boxes(n, nb)
    lev = ix.Level(lib, ix.Geom.make(n, lo, hi), boxes)
    kw = dict(visc_coef=1e-3, cfl=0.7)
    ns = ix.NavierStokes(lib, lev, dev, **kw)
    o = oracle.OracleNS(n, lo, hi, **kw)
    ns.set_turbulent_forcing(nmodes, mode_start, div_free, fd); o.set_turbulent_forcing(nmodes, mode_start, div_free, fd)
    ns.init_prob(20, [1.0, 1.0, 0.5]); o.init_prob(20, [1.0, 1.0, 0.5])
    d1, d2 = ns.post_init(), o.post_init()
    assert abs(d1 - d2) <= 1e-12 * d2
    unforced = oracle.OracleNS(n, lo, hi, **kw)
    unforced.init_prob(20, [1.0, 1.0, 0.5]); unforced.post_init()
    for _ in range(3):
        a, b = ns.step(), o.step()
        unforced.step()
        assert abs(a - b) <= 1e-11 * b
    So = o.get(0)
    err = 0.0
    for il, (blo, bhi) in enumerate(boxes):
        t = ns.field(0, il).cpu().numpy()
        nz, ny, nx = bhi[2] - blo[2] + 1, bhi[1] - blo[1] + 1, bhi[0] - blo[0] + 1
        err = max(err, np.abs(t[:, :nz, :ny, :nx] - So[:, blo[2]:bhi[2] + 1, blo[1]:bhi[1] + 1, blo[0]:bhi[0] + 1]).max())
    assert err <= 1e-10
    assert np.abs(So[:3] - unforced.get(0)[:3]).max() > 1e-3    # the forcing did something
    ns.close(); o.close(); unforced.close(); lev.close()
