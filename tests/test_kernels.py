"""Per-box kernels through the C ABI vs the CPU oracle on the same seeded inputs.
Tolerances: these are single stencil evaluations in fp64; the two sides may differ only by
FMA contraction / summation order, so max|diff| <= 1e-13 * scale is asserted."""
import ctypes as C

import numpy as np
import pytest

import iamr_b200 as ix
from util import (hash_uniform, smooth_field, split_boxes, to_fab, from_fabs, d3, box_of, stream_of, sync)

N = (16, 12, 8)
DX = (1.0 / 16, 1.0 / 12, 1.0 / 8)
DXINV = tuple(1.0 / h for h in DX)
RTOL = 1e-13


def _coeffs(seed, n=N, ncomp=1):
    nz, ny, nx = n[2], n[1], n[0]
    bx = 1.0 + 0.5 * hash_uniform(seed + 1, (ncomp, nz, ny, nx))
    by = 1.0 + 0.5 * hash_uniform(seed + 2, (ncomp, nz, ny, nx))
    bz = 1.0 + 0.5 * hash_uniform(seed + 3, (ncomp, nz, ny, nx))
    alpha = 1.5 + 0.5 * hash_uniform(seed + 4, (1, nz, ny, nx))
    return alpha, bx, by, bz


def _scale(a):
    return max(1.0, float(np.abs(a).max()))


@pytest.mark.parametrize("nb", [(1, 1, 1), (2, 2, 2)])
@pytest.mark.parametrize("a,ncomp,bn", [(0.0, 1, 1), (1.0, 1, 1), (1.0, 3, 3)])
def test_abec_gsrb_and_apply(backend, oracle, nb, a, ncomp, bn):
    lib, dev = backend
    b = 0.37
    alpha, bx, by, bz = _coeffs(10, ncomp=bn)
    phi = hash_uniform(1, (ncomp, N[2], N[1], N[0]))
    rhs = hash_uniform(2, (ncomp, N[2], N[1], N[0]))
    boxes = split_boxes(N, nb)
    s = stream_of(dev)
    for rb in (0, 1):
        ref = oracle.abec_gsrb(DXINV, a, b, alpha if a else None, bx, by, bz, rhs, 1.15, rb, phi)
        outs = []
        for box in boxes:
            tp, fp = to_fab(phi, box, 1, ix.CELL, dev)
            tr, fr = to_fab(rhs, box, 0, ix.CELL, dev)
            ta, fa = to_fab(alpha, box, 0, ix.CELL, dev)
            tbx, fbx = to_fab(bx, box, 0, ix.XFACE, dev)
            tby, fby = to_fab(by, box, 0, ix.YFACE, dev)
            tbz, fbz = to_fab(bz, box, 0, ix.ZFACE, dev)
            bb = box_of(*box)
            lib.check(lib.iamrx_abec_gsrb_box(C.byref(bb), C.byref(fp), C.byref(fr), a, b, C.byref(fa) if a else None,
                                              C.byref(fbx), C.byref(fby), C.byref(fbz), d3(DXINV), 1.15, rb, ncomp, s))
            outs.append(tp)
        sync(dev)
        got, _ = from_fabs(outs, boxes, 1, ix.CELL, N, ncomp)
        assert np.abs(got - ref).max() <= RTOL * _scale(ref) * 100  # division by gamma amplifies rounding
        # cells of the other colour are untouched (bit-exact)
        kk, jj, ii = np.meshgrid(np.arange(N[2]), np.arange(N[1]), np.arange(N[0]), indexing="ij")
        other = ((ii + jj + kk + rb) % 2) == 1
        assert np.array_equal(got[:, other], phi[:, other])
    # apply and residual
    ref = oracle.abec_apply(DXINV, a, b, alpha if a else None, bx, by, bz, phi)
    for with_rhs in (False, True):
        outs = []
        for box in boxes:
            tp, fp = to_fab(phi, box, 1, ix.CELL, dev)
            to, fo = to_fab(np.zeros_like(phi), box, 0, ix.CELL, dev)
            tr, fr = to_fab(rhs, box, 0, ix.CELL, dev)
            ta, fa = to_fab(alpha, box, 0, ix.CELL, dev)
            tbx, fbx = to_fab(bx, box, 0, ix.XFACE, dev)
            tby, fby = to_fab(by, box, 0, ix.YFACE, dev)
            tbz, fbz = to_fab(bz, box, 0, ix.ZFACE, dev)
            bb = box_of(*box)
            lib.check(lib.iamrx_abec_apply_box(C.byref(bb), C.byref(fo), C.byref(fp), C.byref(fr) if with_rhs else None, a, b,
                                               C.byref(fa) if a else None, C.byref(fbx), C.byref(fby), C.byref(fbz),
                                               d3(DXINV), ncomp, s))
            outs.append(to)
        sync(dev)
        got, _ = from_fabs(outs, boxes, 0, ix.CELL, N, ncomp)
        want = (rhs - ref) if with_rhs else ref
        assert np.abs(got - want).max() <= RTOL * _scale(ref)


@pytest.mark.parametrize("N3", [(8, 8, 8), (16, 12, 8), (72, 20, 12), (136, 10, 8), (60, 14, 24), (64, 30, 10)])
@pytest.mark.parametrize("a,ncomp,bn", [(0.0, 1, 1), (1.0, 1, 1), (1.0, 3, 3)])
def test_abec_gsrb_fused_sweep(backend, oracle, N3, a, ncomp, bn):
    """One fused red+black launch (out of place, in-kernel periodic wrap, z-marching tiles) against two
    oracle colour passes; sizes exercise partial tiles in x and y and several z chunks."""
    lib, dev = backend
    dxinv = tuple(float(m) for m in N3)
    alpha, bx, by, bz = _coeffs(11, N3, bn)
    nz, ny, nx = N3[2], N3[1], N3[0]
    rhs = hash_uniform(21, (ncomp, nz, ny, nx))
    phi = hash_uniform(22, (ncomp, nz, ny, nx))
    b = 0.37
    ref = phi
    for sweep in range(2):
        for rb in range(2):
            ref = oracle.abec_gsrb(dxinv, a, b, alpha if a else None, bx, by, bz, rhs, 1.15, rb, ref)
        if sweep == 0:
            ref1 = ref
    box = ((0, 0, 0), (nx - 1, ny - 1, nz - 1))
    tp, fp = to_fab(phi, box, 1, ix.CELL, dev, fill_ghost=False)      # ghost cells must not be read
    to, fo = to_fab(np.zeros_like(phi), box, 1, ix.CELL, dev)
    tr, fr = to_fab(rhs, box, 0, ix.CELL, dev)
    ta, fa = to_fab(alpha, box, 0, ix.CELL, dev)
    tbx, fbx = to_fab(bx, box, 0, ix.XFACE, dev)
    tby, fby = to_fab(by, box, 0, ix.YFACE, dev)
    tbz, fbz = to_fab(bz, box, 0, ix.ZFACE, dev)
    bb = box_of(*box)
    s = stream_of(dev)
    args = (C.byref(fr), a, b, C.byref(fa) if a else None, C.byref(fbx), C.byref(fby), C.byref(fbz), d3(dxinv), 1.15, ncomp, s)
    lib.check(lib.iamrx_abec_gsrb_sweep_box(C.byref(bb), C.byref(fo), C.byref(fp), *args))
    sync(dev)
    got, _ = from_fabs([to], [box], 1, ix.CELL, N3, ncomp)
    assert np.abs(got - ref1).max() <= RTOL * _scale(ref1)
    # second sweep back into the first buffer (ping-pong as the multigrid smoother does)
    lib.check(lib.iamrx_abec_gsrb_sweep_box(C.byref(bb), C.byref(fp), C.byref(fo), *args))
    sync(dev)
    got, _ = from_fabs([tp], [box], 1, ix.CELL, N3, ncomp)
    assert np.abs(got - ref).max() <= RTOL * _scale(ref)


def test_tensor_cross(backend, oracle):
    lib, dev = backend
    vel = smooth_field(N, 5, 3)
    ex, ey, ez = (1.0 + 0.3 * hash_uniform(20 + d, (1, N[2], N[1], N[0])) for d in range(3))
    out0 = hash_uniform(30, (3, N[2], N[1], N[0]))
    ref = oracle.tensor_cross(DXINV, -0.8, ex, ey, ez, vel, out0)
    boxes = split_boxes(N, (2, 1, 2))
    outs = []
    for box in boxes:
        tv, fv = to_fab(vel, box, 1, ix.CELL, dev)
        to, fo = to_fab(out0, box, 0, ix.CELL, dev)
        te = [to_fab(e, box, 0, t, dev) for e, t in ((ex, ix.XFACE), (ey, ix.YFACE), (ez, ix.ZFACE))]
        bb = box_of(*box)
        lib.check(lib.iamrx_tensor_cross_box(C.byref(bb), C.byref(fo), C.byref(fv), C.byref(te[0][1]), C.byref(te[1][1]),
                                             C.byref(te[2][1]), -0.8, d3(DXINV), stream_of(dev)))
        outs.append(to)
    sync(dev)
    got, _ = from_fabs(outs, boxes, 0, ix.CELL, N, 3)
    assert np.abs(got - ref).max() <= RTOL * _scale(ref)


@pytest.mark.parametrize("nb,N", [((1, 1, 1), N), ((2, 2, 1), N), ((1, 1, 1), (72, 20, 12)), ((2, 1, 2), (136, 10, 8))])
def test_nodal_kernels(backend, oracle, nb, N):
    """The last two shapes are wide enough for the shared-memory tile kernels on the GPU
    (partial tiles in x and y, boxes with and without neighbours)."""
    lib, dev = backend
    DX = tuple(1.0 / m for m in N)
    DXINV = tuple(1.0 / h for h in DX)
    sig = 1.0 + 0.5 * hash_uniform(40, (1, N[2], N[1], N[0]))
    phi = hash_uniform(41, (1, N[2], N[1], N[0]))
    rhs = hash_uniform(42, (1, N[2], N[1], N[0]))
    vel = hash_uniform(43, (3, N[2], N[1], N[0]))
    boxes = split_boxes(N, nb)
    s = stream_of(dev)
    # divergence
    ref = oracle.nodal_divu(DXINV, vel)[None]
    outs = []
    for box in boxes:
        tv, fv = to_fab(vel, box, 1, ix.CELL, dev)
        tr, fr = to_fab(np.zeros_like(phi), box, 0, ix.NODE, dev)
        nbx = box_of(box[0], tuple(h + 1 for h in box[1]))
        lib.check(lib.iamrx_nodal_divu_box(C.byref(nbx), C.byref(fr), C.byref(fv), d3(DXINV), s))
        outs.append(tr)
    sync(dev)
    got, dup = from_fabs(outs, boxes, 0, ix.NODE, N, 1)
    assert dup == 0.0 and np.abs(got - ref).max() <= RTOL * _scale(ref)
    # A*phi and residual
    ref = oracle.nodal_adotx(DXINV, sig, phi)
    for with_rhs in (False, True):
        outs = []
        for box in boxes:
            tp, fp = to_fab(phi, box, 1, ix.NODE, dev)
            ts, fs = to_fab(sig, box, 1, ix.CELL, dev)
            tr, fr = to_fab(rhs, box, 0, ix.NODE, dev)
            to, fo = to_fab(np.zeros_like(phi), box, 0, ix.NODE, dev)
            nbx = box_of(box[0], tuple(h + 1 for h in box[1]))
            lib.check(lib.iamrx_nodal_adotx_box(C.byref(nbx), C.byref(fo), C.byref(fp), C.byref(fr) if with_rhs else None,
                                                C.byref(fs), d3(DXINV), s))
            outs.append(to)
        sync(dev)
        got, dup = from_fabs(outs, boxes, 0, ix.NODE, N, 1)
        want = (rhs - ref) if with_rhs else ref
        assert dup == 0.0 and np.abs(got - want).max() <= RTOL * _scale(ref)
    # Gauss-Seidel colours
    for color in range(8):
        ref = oracle.nodal_gs(DXINV, sig, rhs, color, phi)
        outs = []
        for box in boxes:
            tp, fp = to_fab(phi, box, 1, ix.NODE, dev)
            ts, fs = to_fab(sig, box, 1, ix.CELL, dev)
            tr, fr = to_fab(rhs, box, 0, ix.NODE, dev)
            nbx = box_of(box[0], tuple(h + 1 for h in box[1]))
            lib.check(lib.iamrx_nodal_gs_box(C.byref(nbx), C.byref(fp), C.byref(fr), C.byref(fs), d3(DXINV), color, s))
            outs.append(tp)
        sync(dev)
        got, dup = from_fabs(outs, boxes, 1, ix.NODE, N, 1)
        assert dup == 0.0 and np.abs(got - ref).max() <= 1e-12 * _scale(ref)
    # mknewu / gradient
    vref, gref = oracle.nodal_mknewu(DXINV, sig, phi, vel)
    ov, og = [], []
    for box in boxes:
        tp, fp = to_fab(phi, box, 1, ix.NODE, dev)
        ts, fs = to_fab(sig, box, 1, ix.CELL, dev)
        tv, fv = to_fab(vel, box, 1, ix.CELL, dev)
        tg, fg = to_fab(np.zeros_like(vel), box, 0, ix.CELL, dev)
        bb = box_of(*box)
        lib.check(lib.iamrx_nodal_mknewu_box(C.byref(bb), C.byref(fv), C.byref(fg), C.byref(fp), C.byref(fs), d3(DXINV), s))
        ov.append(tv); og.append(tg)
    sync(dev)
    gv, _ = from_fabs(ov, boxes, 1, ix.CELL, N, 3)
    gg, _ = from_fabs(og, boxes, 0, ix.CELL, N, 3)
    assert np.abs(gv - vref).max() <= RTOL * _scale(vref) and np.abs(gg - gref).max() <= RTOL * _scale(gref)


@pytest.mark.parametrize("N", [(2, 2, 2), (4, 6, 2), (16, 16, 16), (72, 20, 12), (136, 10, 8), (58, 30, 6), (112, 28, 4)])
def test_nodal_gs_fused_sweep(backend, oracle, N):
    """Fused out-of-place 8-colour sweep on a single box spanning the periodic domain == eight
    sequential colour passes of the oracle (shapes cover single/multiple/partial tiles and
    domains smaller than the tile halo, where the halo wraps around several times)."""
    lib, dev = backend
    DXINV = tuple(float(m) for m in N)
    sig = 1.0 + 0.5 * hash_uniform(50, (1, N[2], N[1], N[0]))
    phi = hash_uniform(51, (1, N[2], N[1], N[0]))
    rhs = hash_uniform(52, (1, N[2], N[1], N[0]))
    ref = phi
    for color in range(8):
        ref = oracle.nodal_gs(DXINV, sig, rhs, color, ref)
    box = ((0, 0, 0), tuple(m - 1 for m in N))
    s = stream_of(dev)
    tp, fp = to_fab(phi, box, 1, ix.NODE, dev, fill_ghost=False)   # ghost nodes must not be read
    to, fo = to_fab(np.zeros_like(phi), box, 1, ix.NODE, dev)
    ts, fs = to_fab(sig, box, 1, ix.CELL, dev)
    tr, fr = to_fab(rhs, box, 0, ix.NODE, dev)
    nbx = box_of(box[0], tuple(h + 1 for h in box[1]))
    lib.check(lib.iamrx_nodal_gs_sweep_box(C.byref(nbx), C.byref(fo), C.byref(fp), C.byref(fr), C.byref(fs), d3(DXINV), s))
    sync(dev)
    got, dup = from_fabs([to], [box], 1, ix.NODE, N, 1)
    assert dup == 0.0 and np.abs(got - ref).max() <= 1e-12 * _scale(ref)
    # a second sweep back into the first buffer (ping-pong as the multigrid smoother does)
    for color in range(8):
        ref = oracle.nodal_gs(DXINV, sig, rhs, color, ref)
    lib.check(lib.iamrx_nodal_gs_sweep_box(C.byref(nbx), C.byref(fp), C.byref(fo), C.byref(fr), C.byref(fs), d3(DXINV), s))
    sync(dev)
    got, dup = from_fabs([tp], [box], 1, ix.NODE, N, 1)
    assert dup == 0.0 and np.abs(got - ref).max() <= 1e-12 * _scale(ref)


def _adv_inputs(n, ncomp, seed):
    vel = smooth_field(n, seed, 3, amp=0.4)
    q = smooth_field(n, seed + 7, ncomp, amp=0.5) + 1.0
    # sharpen one component so the limiters are active
    q[0] += 0.3 * np.sign(smooth_field(n, seed + 9, 1)[0])
    f = smooth_field(n, seed + 11, max(ncomp, 3), amp=0.2)
    dx = tuple(1.0 / m for m in n)
    # face velocities: average of cells + perturbation (not divergence free on purpose)
    um = 0.5 * (vel[0] + np.roll(vel[0], 1, axis=2)) + 0.05 * smooth_field(n, seed + 1, 1)[0]
    vm = 0.5 * (vel[1] + np.roll(vel[1], 1, axis=1)) + 0.05 * smooth_field(n, seed + 2, 1)[0]
    wm = 0.5 * (vel[2] + np.roll(vel[2], 1, axis=0)) + 0.05 * smooth_field(n, seed + 3, 1)[0]
    return vel, q, f, (um, vm, wm), dx


@pytest.mark.parametrize("nb", [(1, 1, 1), (2, 2, 2)])
@pytest.mark.parametrize("fit,ppm", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_extrap_vel_to_faces(backend, oracle, nb, fit, ppm):
    lib, dev = backend
    n = (16, 16, 8)
    vel, _, f, _, dx = _adv_inputs(n, 3, 100)
    vel[2] += 0.2  # make sure every branch of the upwinding sees both signs
    dt = 0.5 * min(dx) / np.abs(vel).max()
    ref = oracle.extrap_vel_to_faces(dx, dt, vel, f[:3].copy(), fit, ppm)
    g = ix.Geom.make(n)
    boxes = split_boxes(n, nb)
    outs = [[], [], []]
    for box in boxes:
        tv, fv = to_fab(vel, box, 3, ix.CELL, dev)
        tf, ff = to_fab(f[:3], box, 1, ix.CELL, dev)
        macs = [to_fab(np.zeros((1, n[2], n[1], n[0])), box, 1, t, dev) for t in (ix.XFACE, ix.YFACE, ix.ZFACE)]
        bb = box_of(*box)
        lib.check(lib.iamrx_extrap_vel_to_faces_box(C.byref(bb), C.byref(fv), C.byref(ff), C.byref(macs[0][1]),
                                                    C.byref(macs[1][1]), C.byref(macs[2][1]), None, C.byref(g), dt,
                                                    (2 if fit else 0) | (1 if ppm else 0), stream_of(dev)))
        for d in range(3):
            outs[d].append(macs[d][0])
    sync(dev)
    for d, t in enumerate((ix.XFACE, ix.YFACE, ix.ZFACE)):
        got, dup = from_fabs(outs[d], boxes, 1, t, n, 1)
        assert dup == 0.0
        assert np.abs(got[0] - ref[d]).max() <= RTOL * 10


@pytest.mark.parametrize("nb", [(1, 1, 1), (2, 2, 2), (2, 2, 1)])
@pytest.mark.parametrize("ncomp,iconserv,fit,ppm", [(3, (0, 0, 0), 0, 0), (2, (1, 0), 0, 0), (2, (1, 1), 1, 0),
                                                    (3, (0, 0, 0), 0, 1), (2, (1, 0), 1, 1)])   # ppm: Godunov_PPM (staged kernels)
def test_compute_aofs(backend, oracle, nb, ncomp, iconserv, fit, ppm):
    # (1,1,1) and (2,2,1): boxes are whole 8^3 tiles -> fused tile kernel on the GPU; (2,2,2): 8x8x4 boxes -> staged kernels
    lib, dev = backend
    n = (16, 16, 8)
    _, q, f, (um, vm, wm), dx = _adv_inputs(n, ncomp, 200)
    dt = 0.5 * min(dx) / max(np.abs(um).max(), np.abs(vm).max(), np.abs(wm).max())
    ref, (rfx, rfy, rfz, rxe, rye, rze) = oracle.compute_aofs(dx, dt, q, f[:ncomp].copy(), um, vm, wm, iconserv, fit, want_fluxes=True, ppm=ppm)
    g = ix.Geom.make(n)
    boxes = split_boxes(n, nb)
    ic = (C.c_int * ncomp)(*iconserv)
    out_a, out_f, out_e = [], [[], [], []], [[], [], []]
    for box in boxes:
        tq, fq = to_fab(q, box, 3, ix.CELL, dev)
        tf, ff = to_fab(f[:ncomp], box, 1, ix.CELL, dev)
        ta, fa = to_fab(np.zeros_like(q), box, 0, ix.CELL, dev)
        macs = [to_fab(m[None], box, 1, t, dev) for m, t in ((um, ix.XFACE), (vm, ix.YFACE), (wm, ix.ZFACE))]
        fl = [to_fab(np.zeros_like(q), box, 0, t, dev) for t in (ix.XFACE, ix.YFACE, ix.ZFACE)]
        ed = [to_fab(np.zeros_like(q), box, 0, t, dev) for t in (ix.XFACE, ix.YFACE, ix.ZFACE)]
        bb = box_of(*box)
        flags = (2 if fit else 0) | 8 | (1 if ppm else 0)
        lib.check(lib.iamrx_compute_aofs_box(C.byref(bb), C.byref(fa), 0, C.byref(fq), 0, ncomp, C.byref(ff), 0, None,
                                             C.byref(macs[0][1]), C.byref(macs[1][1]), C.byref(macs[2][1]), None, None, None,
                                             C.byref(fl[0][1]), C.byref(fl[1][1]), C.byref(fl[2][1]),
                                             C.byref(ed[0][1]), C.byref(ed[1][1]), C.byref(ed[2][1]),
                                             ic, None, C.byref(g), dt, flags, stream_of(dev)))
        out_a.append(ta)
        for d in range(3):
            out_f[d].append(fl[d][0]); out_e[d].append(ed[d][0])
    sync(dev)
    got, _ = from_fabs(out_a, boxes, 0, ix.CELL, n, ncomp)
    scale = _scale(ref)
    assert np.abs(got - ref).max() <= 1e-12 * scale
    for d, (t, rf, re) in enumerate(((ix.XFACE, rfx, rxe), (ix.YFACE, rfy, rye), (ix.ZFACE, rfz, rze))):
        gf, dup = from_fabs(out_f[d], boxes, 0, t, n, ncomp)
        ge, dup2 = from_fabs(out_e[d], boxes, 0, t, n, ncomp)
        assert dup == 0.0 and dup2 == 0.0
        assert np.abs(ge - re).max() <= RTOL * 10 and np.abs(gf - rf).max() <= RTOL * 10


@pytest.mark.gpu
@pytest.mark.parametrize("nb", [(1, 1, 1), (2, 1, 2)])
@pytest.mark.parametrize("ncomp,iconserv,fit,sync_flag,with_divu", [
    (3, (0, 0, 0), 0, 0, 0), (2, (1, 0), 1, 0, 1), (2, (1, 1), 0, 16, 1), (1, (0,), 1, 16, 0)])
def test_compute_aofs_fused_tile_vs_staged_gpu(cuda_lib, nb, ncomp, iconserv, fit, sync_flag, with_divu):
    """The fused shared-memory tile kernel (boxes made of whole 8^3 tiles) against the staged
    global-scratch kernels (IAMRX_ADV_STAGED) on the same inputs: same algorithm regrouped per cell,
    so only last-bit differences are allowed.  Covers divu, conservative comps, the sync sign and
    the flux / edge-state outputs, which the oracle comparison above does not."""
    lib, dev = cuda_lib, "cuda:0"
    n = (32, 24, 16)
    _, q, f, (um, vm, wm), dx = _adv_inputs(n, ncomp, 300)
    divu = smooth_field(n, 77, 1, amp=0.3)
    dt = 0.5 * min(dx) / max(np.abs(um).max(), np.abs(vm).max(), np.abs(wm).max())
    g = ix.Geom.make(n)
    boxes = split_boxes(n, nb)
    ic = (C.c_int * ncomp)(*iconserv)
    res = {}
    for staged in (0, 32):
        out_a, out_f, out_e = [], [[], [], []], [[], [], []]
        for box in boxes:
            tq, fq = to_fab(q, box, 3, ix.CELL, dev)
            tf, ff = to_fab(f[:ncomp], box, 1, ix.CELL, dev)
            td, fd = to_fab(divu, box, 1, ix.CELL, dev)
            ta, fa = to_fab(0.25 * np.ones_like(q), box, 0, ix.CELL, dev)
            macs = [to_fab(m[None], box, 1, t, dev) for m, t in ((um, ix.XFACE), (vm, ix.YFACE), (wm, ix.ZFACE))]
            fl = [to_fab(np.zeros_like(q), box, 0, t, dev) for t in (ix.XFACE, ix.YFACE, ix.ZFACE)]
            ed = [to_fab(np.zeros_like(q), box, 0, t, dev) for t in (ix.XFACE, ix.YFACE, ix.ZFACE)]
            bb = box_of(*box)
            flags = (2 if fit else 0) | 8 | sync_flag | staged
            lib.check(lib.iamrx_compute_aofs_box(C.byref(bb), C.byref(fa), 0, C.byref(fq), 0, ncomp, C.byref(ff), 0,
                                                 C.byref(fd) if with_divu else None,
                                                 C.byref(macs[0][1]), C.byref(macs[1][1]), C.byref(macs[2][1]), None, None, None,
                                                 C.byref(fl[0][1]), C.byref(fl[1][1]), C.byref(fl[2][1]),
                                                 C.byref(ed[0][1]), C.byref(ed[1][1]), C.byref(ed[2][1]),
                                                 ic, None, C.byref(g), dt, flags, stream_of(dev)))
            out_a.append(ta)
            for d in range(3):
                out_f[d].append(fl[d][0]); out_e[d].append(ed[d][0])
        sync(dev)
        r = [from_fabs(out_a, boxes, 0, ix.CELL, n, ncomp)[0]]
        for d, t in enumerate((ix.XFACE, ix.YFACE, ix.ZFACE)):
            gf, dup = from_fabs(out_f[d], boxes, 0, t, n, ncomp)
            ge, dup2 = from_fabs(out_e[d], boxes, 0, t, n, ncomp)
            assert dup == 0.0 and dup2 == 0.0
            r += [gf, ge]
        res[staged] = r
    for x, y in zip(res[0], res[32]):
        assert np.abs(x - y).max() <= 1e-12 * _scale(y)


def test_bad_arguments(backend):
    lib, dev = backend
    rc = lib.iamrx_abec_gsrb_box(None, None, None, 0.0, 1.0, None, None, None, None, None, 1.0, 0, 1, None)
    assert rc == -1 and b"null" in lib.iamrx_last_error()
    n = (8, 8, 8)
    g = ix.Geom.make(n)
    box = ((0, 0, 0), (7, 7, 7))
    z = np.zeros((2, 8, 8, 8))
    tv, fv = to_fab(z, box, 3, ix.CELL, dev)          # only 2 components: ExtrapVelToFaces needs all three velocities
    macs = [to_fab(z[:1], box, 1, t, dev) for t in (ix.XFACE, ix.YFACE, ix.ZFACE)]
    bb = box_of(*box)
    rc = lib.iamrx_extrap_vel_to_faces_box(C.byref(bb), C.byref(fv), None, C.byref(macs[0][1]), C.byref(macs[1][1]),
                                           C.byref(macs[2][1]), None, C.byref(g), 0.1, 0, stream_of(dev))
    assert rc == -1 and b"3 components" in lib.iamrx_last_error()


@pytest.mark.parametrize("nb", [(1, 1, 1), (2, 2, 2)])   # whole 8^3 tiles -> fused tile kernel on the GPU; 8x8x4 boxes -> staged kernels
@pytest.mark.parametrize("case", ["divu", "sync_ucorr", "known_edges", "sync_known"])
def test_compute_aofs_callsite_arguments(backend, oracle, nb, case):
    """The arguments of the ComputeFluxesOnBoxFromState call site (NSB.cpp:4701-4717) that a plain advection call does
    not exercise, each against the ORACLE (not against another kernel variant): divu != 0 with conservative components
    (edge-state -dt/2 q divu term), the sync call (is_sync: aofs -= update, fluxes built with a distinct U_corr while the
    edge states still use u_mac, NSB.cpp:4672-4677,4834) and known_edge_state (MacProj.cpp:776-785)."""
    lib, dev = backend
    n = (16, 16, 8)
    ncomp, iconserv = 2, (1, 0)
    _, q, f, (um, vm, wm), dx = _adv_inputs(n, ncomp, 400)
    dt = 0.5 * min(dx) / max(np.abs(um).max(), np.abs(vm).max(), np.abs(wm).max())
    divu = smooth_field(n, 77, 1, amp=0.3) if case == "divu" else None
    is_sync = case in ("sync_ucorr", "sync_known")
    known = case in ("known_edges", "sync_known")
    ucorr = [0.1 * smooth_field(n, 500 + d, 1)[0] for d in range(3)] if is_sync else None
    aofs0 = 0.25 + 0.1 * smooth_field(n, 600, ncomp) if is_sync else None
    edges = [q + 0.05 * smooth_field(n, 700 + d, ncomp) for d in range(3)] if known else None
    ref, rfl, red = oracle.compute_aofs2(dx, dt, q, f[:ncomp].copy(), um, vm, wm, iconserv, fit=0, divu=divu, uflux=ucorr,
                                         is_sync=is_sync, aofs_in=aofs0, known_edges=edges)
    g = ix.Geom.make(n)
    boxes = split_boxes(n, nb)
    ic = (C.c_int * ncomp)(*iconserv)
    out_a, out_f, out_e = [], [[], [], []], [[], [], []]
    for box in boxes:
        tq, fq = to_fab(q, box, 3, ix.CELL, dev)
        tf, ff = to_fab(f[:ncomp], box, 1, ix.CELL, dev)
        ta, fa = to_fab(aofs0 if is_sync else np.zeros_like(q), box, 0, ix.CELL, dev)
        macs = [to_fab(m[None], box, 1, t, dev) for m, t in ((um, ix.XFACE), (vm, ix.YFACE), (wm, ix.ZFACE))]
        ucs = [to_fab(m[None], box, 0, t, dev) for m, t in zip(ucorr, (ix.XFACE, ix.YFACE, ix.ZFACE))] if is_sync else None
        fl = [to_fab(np.zeros_like(q), box, 0, t, dev) for t in (ix.XFACE, ix.YFACE, ix.ZFACE)]
        ed = [to_fab(edges[d] if known else np.zeros_like(q), box, 0, t, dev) for d, t in enumerate((ix.XFACE, ix.YFACE, ix.ZFACE))]
        td, fd = to_fab(divu, box, 1, ix.CELL, dev) if divu is not None else (None, None)
        bb = box_of(*box)
        flags = ix.ADV_WRITE_FLUXES | (ix.ADV_IS_SYNC if is_sync else 0) | (ix.ADV_KNOWN_EDGE_STATE if known else 0)
        lib.check(lib.iamrx_compute_aofs_box(C.byref(bb), C.byref(fa), 0, C.byref(fq), 0, ncomp, C.byref(ff), 0,
                                             C.byref(fd) if fd is not None else None,
                                             C.byref(macs[0][1]), C.byref(macs[1][1]), C.byref(macs[2][1]),
                                             *([C.byref(u[1]) for u in ucs] if is_sync else [None, None, None]),
                                             C.byref(fl[0][1]), C.byref(fl[1][1]), C.byref(fl[2][1]),
                                             C.byref(ed[0][1]), C.byref(ed[1][1]), C.byref(ed[2][1]),
                                             ic, None, C.byref(g), dt, flags, stream_of(dev)))
        out_a.append(ta)
        for d in range(3):
            out_f[d].append(fl[d][0]); out_e[d].append(ed[d][0])
    sync(dev)
    got, _ = from_fabs(out_a, boxes, 0, ix.CELL, n, ncomp)
    assert np.abs(got - ref).max() <= 1e-12 * _scale(ref)
    for d, t in enumerate((ix.XFACE, ix.YFACE, ix.ZFACE)):
        gf, dup = from_fabs(out_f[d], boxes, 0, t, n, ncomp)
        ge, dup2 = from_fabs(out_e[d], boxes, 0, t, n, ncomp)
        assert dup == 0.0 and dup2 == 0.0
        assert np.abs(gf - rfl[d]).max() <= RTOL * 10 and np.abs(ge - red[d]).max() <= RTOL * 10


def test_advection_argument_checks(backend):
    """Fabs that do not cover the stencil are rejected instead of read out of bounds (S needs 3 ghost cells,
    force / the MAC velocities 1: NSB.cpp:4539-4552, NavierStokesBase.H:810)."""
    lib, dev = backend
    n = (8, 8, 8)
    g = ix.Geom.make(n)
    box = ((0, 0, 0), (7, 7, 7))
    z = np.zeros((3, 8, 8, 8))
    bb = box_of(*box)
    ic = (C.c_int * 3)(0, 0, 0)
    macs = [to_fab(z[:1], box, 1, t, dev) for t in (ix.XFACE, ix.YFACE, ix.ZFACE)]
    ta, fa = to_fab(z, box, 0, ix.CELL, dev)
    for ng_s, ng_f, ng_m, msg in ((2, 1, 1, b"grown by 3"), (3, 0, 1, b"force"), (3, 1, 0, b"ghost face layer")):
        tq, fq = to_fab(z, box, ng_s, ix.CELL, dev)
        tf, ff = to_fab(z, box, ng_f, ix.CELL, dev)
        mm = [to_fab(z[:1], box, ng_m, t, dev) for t in (ix.XFACE, ix.YFACE, ix.ZFACE)]
        rc = lib.iamrx_compute_aofs_box(C.byref(bb), C.byref(fa), 0, C.byref(fq), 0, 3, C.byref(ff), 0, None, C.byref(mm[0][1]),
                                        C.byref(mm[1][1]), C.byref(mm[2][1]), None, None, None, None, None, None, None, None, None,
                                        ic, None, C.byref(g), 0.1, 0, stream_of(dev))
        assert rc == -1 and msg in lib.iamrx_last_error(), lib.iamrx_last_error()
    tq, fq = to_fab(z, box, 2, ix.CELL, dev)
    rc = lib.iamrx_extrap_vel_to_faces_box(C.byref(bb), C.byref(fq), None, C.byref(macs[0][1]), C.byref(macs[1][1]),
                                           C.byref(macs[2][1]), None, C.byref(g), 0.1, 0, stream_of(dev))
    assert rc == -1 and b"grown by 3" in lib.iamrx_last_error()
    # one flux velocity without the others, an unknown BC code
    tq, fq = to_fab(z, box, 3, ix.CELL, dev)
    tf, ff = to_fab(z, box, 1, ix.CELL, dev)
    rc = lib.iamrx_compute_aofs_box(C.byref(bb), C.byref(fa), 0, C.byref(fq), 0, 3, C.byref(ff), 0, None, C.byref(macs[0][1]),
                                    C.byref(macs[1][1]), C.byref(macs[2][1]), C.byref(macs[0][1]), None, None, None, None, None, None, None,
                                    None, ic, None, C.byref(g), 0.1, 0, stream_of(dev))
    assert rc == -1 and b"all three" in lib.iamrx_last_error()
    bad = (ix.BCRec * 3)(*[ix.BCRec.make((7, 0, 0), (0, 0, 0))] * 3)
    rc = lib.iamrx_compute_aofs_box(C.byref(bb), C.byref(fa), 0, C.byref(fq), 0, 3, C.byref(ff), 0, None, C.byref(macs[0][1]),
                                    C.byref(macs[1][1]), C.byref(macs[2][1]), None, None, None, None, None, None, None, None,
                                    None, ic, bad, C.byref(g), 0.1, 0, stream_of(dev))
    assert rc == -1 and b"BCRec" in lib.iamrx_last_error()
