"""Two-level coupling building blocks through the C ABI vs the oracle: average_down (cells / faces / nodes),
the FillPatchTwoLevels interpolaters (cell_cons_interp, node_bilinear_interp, face_linear_interp) and the advective flux
register (CrseAdd / FineAdd / Reflux), plus the properties they exist for: conservation and the telescoping of fluxes."""
import ctypes as C

import numpy as np
import pytest

import iamr_b200 as ix
from util import hash_uniform, smooth_field, to_fab, from_fabs, split_boxes, fab_array, box_of, stream_of, sync

NC = (8, 8, 8)
NF = (16, 16, 16)


@pytest.mark.parametrize("ixtype", [ix.CELL, ix.XFACE, ix.YFACE, ix.ZFACE, ix.NODE])
def test_average_down(backend, oracle, ixtype):
    lib, dev = backend
    fine = hash_uniform(11 + ixtype, (2,) + NF[::-1])
    ref = oracle.average_down(NC, ixtype, fine)
    outs, cboxes = [], split_boxes(NC, (2, 1, 2))
    for cb in cboxes:
        fb = (tuple(2 * l for l in cb[0]), tuple(2 * h + 1 for h in cb[1]))
        tf, ff = to_fab(fine, fb, 1, ixtype, dev)
        tc, fc = to_fab(np.zeros((2,) + NC[::-1]), cb, 0, ixtype, dev)
        bb = box_of(*cb)
        lib.check(lib.iamrx_average_down_box(C.byref(bb), C.byref(fc), C.byref(ff), 2, ixtype, stream_of(dev)))
        outs.append(tc)
    sync(dev)
    got, dup = from_fabs(outs, cboxes, 0, ixtype, NC, 2)
    assert dup == 0.0
    assert np.abs(got - ref).max() <= 1e-15


@pytest.mark.parametrize("kind,ixtype", [(0, ix.CELL), (1, ix.NODE), (2, ix.XFACE), (3, ix.YFACE), (4, ix.ZFACE)])
def test_interpolaters(backend, oracle, kind, ixtype):
    lib, dev = backend
    crse = smooth_field(NC, 21 + kind, 2)
    crse[0] += 0.4 * np.sign(smooth_field(NC, 29, 1)[0])      # steep: the limiters are active
    ref = oracle.interp(kind, NC, crse)
    fboxes = split_boxes(NF, (2, 2, 1))
    outs = []
    for fb in fboxes:
        cb = (tuple(l // 2 for l in fb[0]), tuple(h // 2 for h in fb[1]))
        tcr, fcr = to_fab(crse, cb, 1, ixtype, dev)
        tfi, ffi = to_fab(np.zeros((2,) + NF[::-1]), fb, 0, ixtype, dev)
        bb = box_of(*fb)
        lib.check(lib.iamrx_interp_box(kind, C.byref(bb), C.byref(ffi), C.byref(fcr), 2, stream_of(dev)))
        outs.append(tfi)
    sync(dev)
    got, dup = from_fabs(outs, fboxes, 0, ixtype, NF, 2)
    assert dup <= 1e-15
    assert np.abs(got - ref).max() <= 1e-14
    if kind == 0:
        # conservative: the children average back to the parent; and no new extrema (range of the 3^3 coarse neighbourhood)
        assert np.abs(oracle.average_down(NC, ix.CELL, got) - crse).max() <= 1e-14
        lo = np.min([np.roll(crse, (a, b, c), (1, 2, 3)) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)], axis=0)
        hi = np.max([np.roll(crse, (a, b, c), (1, 2, 3)) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)], axis=0)
        up = lambda a: np.repeat(np.repeat(np.repeat(a, 2, 1), 2, 2), 2, 3)
        assert (got >= up(lo) - 1e-14).all() and (got <= up(hi) + 1e-14).all()
    if kind == 1:
        assert np.array_equal(got[:, ::2, ::2, ::2], crse)   # coincident nodes are injected


def test_cell_cons_interp_is_exact_for_linear_data(backend):
    lib, dev = backend
    z, y, x = np.meshgrid(*[np.arange(-1, 9) + 0.5] * 3, indexing="ij")
    crse = (0.3 * x - 0.7 * y + 1.1 * z)[None]
    import torch
    tcr = torch.from_numpy(crse.copy()).to(dev)
    fcr = ix.fab_of(tcr, [-1, -1, -1])
    tfi, ffi = ix.alloc_fab((0, 0, 0), (15, 15, 15), 1, 0, dev)
    bb = box_of((0, 0, 0), (15, 15, 15))
    lib.check(lib.iamrx_interp_box(0, C.byref(bb), C.byref(ffi), C.byref(fcr), 1, stream_of(dev)))
    sync(dev)
    zf, yf, xf = np.meshgrid(*[(np.arange(16) + 0.5) / 2] * 3, indexing="ij")
    assert np.abs(tfi.cpu().numpy()[0] - (0.3 * xf - 0.7 * yf + 1.1 * zf)).max() <= 1e-13


def _level_pair(lib, fine_boxes_c):
    """coarse level 8^3 (two boxes) + fine level = the refinement of the given coarse-index boxes"""
    gc = ix.Geom.make(NC)
    gf = ix.Geom.make(NF)
    cboxes = split_boxes(NC, (2, 1, 1))
    fboxes = [(tuple(2 * l for l in lo), tuple(2 * h + 1 for h in hi)) for lo, hi in fine_boxes_c]
    return ix.Level(lib, gc, cboxes), ix.Level(lib, gf, fboxes), cboxes, fboxes


FINE_LAYOUTS = [[((2, 2, 2), (5, 5, 5))],                                   # one fine box in the middle
                [((2, 0, 2), (3, 7, 5)), ((4, 0, 2), (5, 7, 5))],           # two abutting boxes spanning the periodic y direction
                [((0, 2, 2), (1, 5, 5)), ((6, 2, 2), (7, 5, 5))]]           # two boxes that touch through the periodic x boundary


@pytest.mark.parametrize("layout", FINE_LAYOUTS)
def test_flux_register(backend, oracle, layout):
    lib, dev = backend
    ncomp = 2
    clev, flev, cboxes, fboxes = _level_pair(lib, layout)
    mask = np.zeros(NC[::-1], dtype=bool)
    for lo, hi in layout:
        mask[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = True
    cflux = [hash_uniform(31 + d, (ncomp,) + NC[::-1]) for d in range(3)]
    fflux = [hash_uniform(41 + d, (ncomp,) + NF[::-1]) for d in range(3)]
    dt, vol = 0.05, (1.0 / 8) ** 3
    ref = oracle.fluxreg(NC, mask, cflux, fflux, dt, vol)
    reg = C.c_void_p()
    lib.check(lib.iamrx_fluxreg_create(clev.h, flev.h, ncomp, C.byref(reg)))
    assert lib.iamrx_fluxreg_num_patches(reg) > 0
    CF = [[to_fab(cflux[d], b, 0, t, dev) for b in cboxes] for d, t in enumerate((ix.XFACE, ix.YFACE, ix.ZFACE))]
    FF = [[to_fab(fflux[d], b, 0, t, dev) for b in fboxes] for d, t in enumerate((ix.XFACE, ix.YFACE, ix.ZFACE))]
    fa = lambda L: fab_array([p[1] for p in L])
    lib.check(lib.iamrx_fluxreg_reset(reg, stream_of(dev)))
    lib.check(lib.iamrx_fluxreg_crse_add(reg, fa(CF[0]), fa(CF[1]), fa(CF[2]), dt, vol, stream_of(dev)))
    lib.check(lib.iamrx_fluxreg_fine_add(reg, fa(FF[0]), fa(FF[1]), fa(FF[2]), dt, vol, stream_of(dev)))
    sync(dev)
    got = np.zeros((ncomp,) + NC[::-1])
    for il, (lo, hi) in enumerate(cboxes):
        f = ix.Fab()
        lib.check(lib.iamrx_fluxreg_field(reg, il, C.byref(f)))
        t = ix.tensor_of(f, dev).cpu().numpy()
        got[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = t[:, :hi[2] - lo[2] + 1, :hi[1] - lo[1] + 1, :hi[0] - lo[0] + 1]
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(got[:, mask]).max() == 0.0            # nothing under the fine grids
    # reflux: state += register
    state = hash_uniform(51, (3,) + NC[::-1])
    ST = [to_fab(state, b, 0, ix.CELL, dev) for b in cboxes]
    lib.check(lib.iamrx_fluxreg_reflux(reg, fa(ST), 1, 1.0, stream_of(dev)))
    sync(dev)
    snew, _ = from_fabs([p[0] for p in ST], cboxes, 0, ix.CELL, NC, 3)
    assert np.abs(snew[1:] - (state[1:] + ref)).max() <= 1e-13 and np.array_equal(snew[0], state[0])
    lib.iamrx_fluxreg_destroy(reg)
    clev.close(); flev.close()


def test_flux_register_restores_conservation(backend, oracle):
    """The reason the register exists: advance a coarse level and a fine patch with their own (inconsistent) fluxes, average the
    fine result down, reflux -- the composite total changes exactly by nothing (periodic domain)."""
    lib, dev = backend
    layout = FINE_LAYOUTS[1]
    clev, flev, cboxes, fboxes = _level_pair(lib, layout)
    mask = np.zeros(NC[::-1], dtype=bool)
    for lo, hi in layout:
        mask[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = True
    dxc, dxf = 1.0 / 8, 1.0 / 16
    cflux = [hash_uniform(61 + d, (1,) + NC[::-1]) * dxc ** 2 for d in range(3)]     # area-weighted
    fflux = [hash_uniform(71 + d, (1,) + NF[::-1]) * dxf ** 2 for d in range(3)]
    dt = 0.01
    div = lambda f, h: sum((np.roll(f[d], -1, 3 - d) - f[d]) for d in range(3)) / h ** 3
    sc = 1.0 + 0.1 * hash_uniform(81, (1,) + NC[::-1])
    sf = np.repeat(np.repeat(np.repeat(sc, 2, 1), 2, 2), 2, 3)
    total0 = (sc * (~mask)).sum() * dxc ** 3 + (sf * np.repeat(np.repeat(np.repeat(mask, 2, 0), 2, 1), 2, 2)).sum() * dxf ** 3
    sc1 = sc - dt * div(cflux, dxc)
    sf1 = sf - dt * div(fflux, dxf)      # (the fine patch's outer faces use its own fine fluxes)
    reg = C.c_void_p()
    lib.check(lib.iamrx_fluxreg_create(clev.h, flev.h, 1, C.byref(reg)))
    CF = [[to_fab(cflux[d], b, 0, t, dev) for b in cboxes] for d, t in enumerate((ix.XFACE, ix.YFACE, ix.ZFACE))]
    FF = [[to_fab(fflux[d], b, 0, t, dev) for b in fboxes] for d, t in enumerate((ix.XFACE, ix.YFACE, ix.ZFACE))]
    fa = lambda L: fab_array([p[1] for p in L])
    lib.check(lib.iamrx_fluxreg_crse_add(reg, fa(CF[0]), fa(CF[1]), fa(CF[2]), dt, dxc ** 3, stream_of(dev)))
    lib.check(lib.iamrx_fluxreg_fine_add(reg, fa(FF[0]), fa(FF[1]), fa(FF[2]), dt, dxc ** 3, stream_of(dev)))
    ST = [to_fab(sc1, b, 0, ix.CELL, dev) for b in cboxes]
    lib.check(lib.iamrx_fluxreg_reflux(reg, fa(ST), 0, 1.0, stream_of(dev)))
    sync(dev)
    sc2, _ = from_fabs([p[0] for p in ST], cboxes, 0, ix.CELL, NC, 1)
    fm = np.repeat(np.repeat(np.repeat(mask, 2, 0), 2, 1), 2, 2)
    total_noreflux = (sc1[0] * (~mask)).sum() * dxc ** 3 + (sf1[0] * fm).sum() * dxf ** 3
    total_reflux = (sc2[0] * (~mask)).sum() * dxc ** 3 + (sf1[0] * fm).sum() * dxf ** 3
    assert abs(total_noreflux - total0) > 1e-6          # the mismatch is real
    assert abs(total_reflux - total0) <= 1e-14 * abs(total0) * 10
    lib.iamrx_fluxreg_destroy(reg)
    clev.close(); flev.close()


@pytest.mark.parametrize("layout", FINE_LAYOUTS)
@pytest.mark.parametrize("ngrow", [1, 3])
def test_fillpatch_two_levels(backend, oracle, layout, ngrow):
    """amrex::FillPatchTwoLevels (cell data, conservative linear interpolation, linear in time): ghost cells under a fine
    neighbour (or its periodic image) take the fine data, the rest the interpolated coarse data at `time`.  Reference =
    the oracle's whole-domain cell_cons_interp of the time-interpolated coarse field + the fine data where the fine level exists."""
    lib, dev = backend
    ncomp = 2
    clev, flev, cboxes, fboxes = _level_pair(lib, layout)
    cmask = np.zeros(NC[::-1], dtype=bool)
    for lo, hi in layout:
        cmask[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = True
    fmask = np.repeat(np.repeat(np.repeat(cmask, 2, 0), 2, 1), 2, 2)
    c_old = smooth_field(NC, 91, ncomp)
    c_new = smooth_field(NC, 92, ncomp)
    c_new[0] += 0.3 * np.sign(smooth_field(NC, 93, 1)[0])      # steep: limited slopes
    fdat = hash_uniform(94, (ncomp,) + NF[::-1])
    t_old, t_new, time = 0.5, 0.9, 0.62
    w = (time - t_old) / (t_new - t_old)
    interp = oracle.interp(0, NC, (1.0 - w) * c_old + w * c_new)
    expect = np.where(fmask[None], fdat, interp)
    CO = [to_fab(c_old, b, 0, ix.CELL, dev) for b in cboxes]
    CN = [to_fab(c_new, b, 0, ix.CELL, dev) for b in cboxes]
    FI = [to_fab(fdat, b, ngrow, ix.CELL, dev, fill_ghost=False) for b in fboxes]
    fa = lambda L: fab_array([p[1] for p in L])
    lib.check(lib.iamrx_fillpatch_two_levels(flev.h, clev.h, fa(FI), fa(CO), fa(CN), t_old, t_new, time, ncomp, ngrow,
                                             None, None, stream_of(dev)))
    sync(dev)
    for (t, _), b in zip(FI, fboxes):
        ref, _ = to_fab(expect, b, ngrow, ix.CELL, "cpu")
        assert np.abs(t.cpu().numpy() - ref.numpy()).max() <= 1e-14
    # crse_old = NULL: the new data as they are
    FI = [to_fab(fdat, b, ngrow, ix.CELL, dev, fill_ghost=False) for b in fboxes]
    lib.check(lib.iamrx_fillpatch_two_levels(flev.h, clev.h, fa(FI), None, fa(CN), t_old, t_new, t_new, ncomp, ngrow,
                                             None, None, stream_of(dev)))
    sync(dev)
    expect = np.where(fmask[None], fdat, oracle.interp(0, NC, c_new))
    for (t, _), b in zip(FI, fboxes):
        ref, _ = to_fab(expect, b, ngrow, ix.CELL, "cpu")
        assert np.abs(t.cpu().numpy() - ref.numpy()).max() <= 1e-14
    clev.close(); flev.close()


def test_fillpatch_two_levels_walls(backend, oracle):
    """Non-periodic z (comp 0 reflect_even, comp 1 reflect_odd): a fine box on the low wall.  Cells outside the domain mirror the
    filled interior (physical boundary fill after the interpolation); interior ghost cells away from the wall and from the
    fine box's lateral neighbours equal the periodic-case interpolation two coarse cells away from the wall."""
    lib, dev = backend
    ncomp, ngrow = 2, 2
    gc = ix.Geom.make(NC, periodic=(1, 1, 0))
    gf = ix.Geom.make(NF, periodic=(1, 1, 0))
    cboxes = split_boxes(NC, (2, 1, 1))
    layout = [((2, 2, 0), (5, 5, 3))]
    fboxes = [(tuple(2 * l for l in lo), tuple(2 * h + 1 for h in hi)) for lo, hi in layout]
    clev, flev = ix.Level(lib, gc, cboxes), ix.Level(lib, gf, fboxes)
    c_new = smooth_field(NC, 95, ncomp)
    fdat = hash_uniform(96, (ncomp,) + NF[::-1])
    CN = [to_fab(c_new, b, 0, ix.CELL, dev) for b in cboxes]
    FI = [to_fab(fdat, b, ngrow, ix.CELL, dev, fill_ghost=False) for b in fboxes]
    bcs = (ix.BCRec * ncomp)(ix.BCRec.make((0, 0, ix.BC_REFLECT_EVEN), (0, 0, ix.BC_REFLECT_EVEN)),
                             ix.BCRec.make((0, 0, ix.BC_REFLECT_ODD), (0, 0, ix.BC_REFLECT_ODD)))
    fa = lambda L: fab_array([p[1] for p in L])
    lib.check(lib.iamrx_fillpatch_two_levels(flev.h, clev.h, fa(FI), None, fa(CN), 0.0, 1.0, 1.0, ncomp, ngrow,
                                             bcs, None, stream_of(dev)))
    sync(dev)
    a = FI[0][0].cpu().numpy()                      # (ncomp, 8+4, 8+4, 8+4), origin (4-2, 4-2, 0-2)
    assert np.abs(a).max() < 1e30                   # every ghost cell was written
    g = ngrow
    assert np.array_equal(a[:, g:-g, g:-g, g:-g], fdat[:, 0:8, 4:12, 4:12])
    for m in range(g):                              # below the wall: mirror images (even / odd)
        assert np.array_equal(a[0, g - 1 - m], a[0, g + m])
        assert np.array_equal(a[1, g - 1 - m], -a[1, g + m])
    # lateral ghost cells at z >= 4 fine cells from the wall: the coarse stencil there does not see the wall -> periodic interpolation
    per = oracle.interp(0, NC, c_new)
    ref, _ = to_fab(per, fboxes[0], ngrow, ix.CELL, "cpu")
    r = ref.numpy()
    lat = np.ones(a.shape[1:], dtype=bool)
    lat[:, g:-g, g:-g] = False                      # x / y ghost columns only
    lat[:g + 4] = False
    lat[-g:] = False
    assert lat.any() and np.abs(a[:, lat] - r[:, lat]).max() <= 1e-14
    clev.close(); flev.close()


@pytest.mark.parametrize("layout", FINE_LAYOUTS)
@pytest.mark.parametrize("which,increment", [(ix.SYNC_CELL_CONS, 1), (ix.SYNC_PC, 1), (ix.SYNC_CELL_CONS, 0)])
def test_sync_interp(backend, oracle, layout, which, increment):
    """NavierStokesBase::SyncInterp (NSB.cpp:3071-3255) on a periodic two-level hierarchy: fine[dest..] (+)= dt_clev * I(crse[src..]) with
    pc_interp or cell_cons_interp; the other components of the fine fabs are left alone."""
    lib, dev = backend
    ncomp, src, dest, dt = 2, 1, 2, 0.37
    clev, flev, cboxes, fboxes = _level_pair(lib, layout)
    crse = smooth_field(NC, 101, src + ncomp)
    crse[src] += 0.3 * np.sign(smooth_field(NC, 102, 1)[0])
    fine = hash_uniform(103, (dest + ncomp,) + NF[::-1])
    if which == ix.SYNC_PC:
        interp = np.repeat(np.repeat(np.repeat(crse[src:src + ncomp], 2, 1), 2, 2), 2, 3)
    else:
        interp = oracle.interp(0, NC, crse[src:src + ncomp])
    expect = fine.copy()
    expect[dest:] = fine[dest:] + dt * interp if increment else interp
    CS = [to_fab(crse, b, 0, ix.CELL, dev) for b in cboxes]
    FS = [to_fab(fine, b, 0, ix.CELL, dev) for b in fboxes]
    fa = lambda L: fab_array([p[1] for p in L])
    lib.check(lib.iamrx_sync_interp(flev.h, clev.h, fa(FS), dest, fa(CS), src, ncomp, increment, dt, which, None, stream_of(dev)))
    sync(dev)
    for (t, _), b in zip(FS, fboxes):
        ref, _ = to_fab(expect, b, 0, ix.CELL, "cpu")
        assert np.abs(t.cpu().numpy() - ref.numpy()).max() <= 2e-15
    if which == ix.SYNC_CELL_CONS and increment:
        # conservative: the fine correction integrates to the coarse one under the fine grids
        for (t, _), (lo, hi) in zip(FS, fboxes):
            d = t.cpu().numpy()[dest:] - fine[dest:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
            c = crse[src:src + ncomp, lo[2] // 2:hi[2] // 2 + 1, lo[1] // 2:hi[1] // 2 + 1, lo[0] // 2:hi[0] // 2 + 1]
            assert abs(d.sum() / 8.0 - dt * c.sum()) <= 1e-12 * max(1.0, abs(c.sum()))
    clev.close(); flev.close()


def test_sync_interp_walls_homogeneous_extdir(backend):
    """Non-periodic z with an ext_dir component: SyncInterp fills the coarse ghost cells with the HOMOGENEOUS boundary value
    (HomExtDirFill), so a constant correction is NOT constant in the fine cells next to the wall under cell_cons_interp ... it is:
    ext_dir ghost = 0 gives a slope towards the wall, limited by the neighbours; piecewise-constant interpolation ignores ghosts."""
    lib, dev = backend
    gc = ix.Geom.make(NC, periodic=(1, 1, 0))
    gf = ix.Geom.make(NF, periodic=(1, 1, 0))
    cboxes = split_boxes(NC, (1, 1, 1))
    fboxes = [((4, 4, 0), (11, 11, 7))]
    clev, flev = ix.Level(lib, gc, cboxes), ix.Level(lib, gf, fboxes)
    crse = np.ones((1,) + NC[::-1])
    fine = np.zeros((1,) + NF[::-1])
    bcs = (ix.BCRec * 1)(ix.BCRec.make((0, 0, ix.BC_EXT_DIR), (0, 0, ix.BC_EXT_DIR)))
    fa = lambda L: fab_array([p[1] for p in L])
    for which in (ix.SYNC_PC, ix.SYNC_CELL_CONS):
        CS = [to_fab(crse, b, 0, ix.CELL, dev) for b in cboxes]
        FS = [to_fab(fine, b, 0, ix.CELL, dev) for b in fboxes]
        lib.check(lib.iamrx_sync_interp(flev.h, clev.h, fa(FS), 0, fa(CS), 0, 1, 0, 1.0, which, bcs, stream_of(dev)))
        sync(dev)
        a = FS[0][0].cpu().numpy()[0]
        assert np.abs(a[2:] - 1.0).max() <= 1e-15          # away from the wall: the constant
        pair = a[0:2].mean(axis=0)
        assert np.abs(pair - 1.0).max() <= 1e-15           # conservative in the wall cells as well
        if which == ix.SYNC_PC:
            assert np.abs(a[0:2] - 1.0).max() == 0.0
    clev.close(); flev.close()


@pytest.mark.parametrize("layout", FINE_LAYOUTS)
def test_sync_proj_interp(backend, oracle, layout):
    """NavierStokesBase::SyncProjInterp (NSB.cpp:3258-3336): P_new += I(phi), P_old += I(phi), I = node_bilinear_interp."""
    lib, dev = backend
    clev, flev, cboxes, fboxes = _level_pair(lib, layout)
    phi = smooth_field(NC, 111, 1)
    pn, po = hash_uniform(112, (1,) + NF[::-1]), hash_uniform(113, (1,) + NF[::-1])
    interp = oracle.interp(1, NC, phi)
    PH = [to_fab(phi, b, 0, ix.NODE, dev) for b in cboxes]
    PN = [to_fab(pn, b, 0, ix.NODE, dev) for b in fboxes]
    PO = [to_fab(po, b, 0, ix.NODE, dev) for b in fboxes]
    fa = lambda L: fab_array([p[1] for p in L])
    lib.check(lib.iamrx_sync_proj_interp(flev.h, clev.h, fa(PN), fa(PO), fa(PH), stream_of(dev)))
    sync(dev)
    for arr, base in ((PN, pn), (PO, po)):
        for (t, _), b in zip(arr, fboxes):
            ref, _ = to_fab(base + interp, b, 0, ix.NODE, "cpu")
            assert np.abs(t.cpu().numpy() - ref.numpy()).max() <= 2e-15
    clev.close(); flev.close()


@pytest.mark.parametrize("layout", FINE_LAYOUTS)
def test_fill_coarse_patch_nodal(backend, oracle, layout):
    """AmrLevel::FillCoarsePatch of Press_Type (Projection.cpp:236-239): every fine node = node_bilinear_interp of the coarse pressure of
    the time interval that contains `time` (Press_Type is an Interval quantity: no interpolation in time)."""
    lib, dev = backend
    clev, flev, cboxes, fboxes = _level_pair(lib, layout)
    p0, p1 = smooth_field(NC, 121, 1), smooth_field(NC, 122, 1)
    P0 = [to_fab(p0, b, 0, ix.NODE, dev) for b in cboxes]
    P1 = [to_fab(p1, b, 0, ix.NODE, dev) for b in cboxes]
    PF = [to_fab(hash_uniform(123, (1,) + NF[::-1]), b, 0, ix.NODE, dev) for b in fboxes]
    fa = lambda L: fab_array([p[1] for p in L])
    t0, t1 = 0.3, 0.5                                   # the coarse pressure's new interval; the old one ends at t0
    for t, src in ((0.35, p1), (0.45, p1), (0.3, p1), (0.25, p0)):
        lib.check(lib.iamrx_fill_coarse_patch_nodal(flev.h, clev.h, fa(PF), fa(P0), fa(P1), t0, t1, t, stream_of(dev)))
        sync(dev)
        expect = oracle.interp(1, NC, src)
        for (tt, _), b in zip(PF, fboxes):
            ref, _ = to_fab(expect, b, 0, ix.NODE, "cpu")
            assert np.abs(tt.cpu().numpy() - ref.numpy()).max() <= 2e-15
    assert lib.iamrx_fill_coarse_patch_nodal(flev.h, clev.h, fa(PF), None, fa(P1), t0, t1, 0.25, stream_of(dev)) == -1   # no old data
    assert lib.iamrx_fill_coarse_patch_nodal(flev.h, clev.h, fa(PF), fa(P0), fa(P1), t0, t1, 0.7, stream_of(dev)) == -1  # beyond the new interval
    clev.close(); flev.close()


def test_mac_sync_solve(backend, oracle):
    """MacProj::mac_sync_solve (MacProj.cpp:359-479): right-hand side from the MAC register (reflux with scale -1) + increment, negated;
    coarse-level MAC solve with rhs_scale 2/dt and no div(umac); Ucorr = -(-B grad phi).  Reference: the oracle's flux register and
    MAC projection composed the same way."""
    lib, dev = backend
    layout = FINE_LAYOUTS[1]
    clev, flev, cboxes, fboxes = _level_pair(lib, layout)
    mask = np.zeros(NC[::-1], dtype=bool)
    for lo, hi in layout:
        mask[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = True
    dxc, dxf, dt = 1.0 / 8, 1.0 / 16, 0.05
    vol = dxc ** 3
    cflux = [hash_uniform(121 + d, (1,) + NC[::-1]) * dxc ** 2 for d in range(3)]     # area-weighted face velocities
    fflux = [hash_uniform(131 + d, (1,) + NF[::-1]) * dxf ** 2 for d in range(3)]
    inc = 0.1 * smooth_field(NC, 141, 1)
    rho = 1.0 + 0.3 * smooth_field(NC, 142, 1)
    reg_ref = oracle.fluxreg(NC, mask, cflux, fflux, 1.0, vol)
    rhs_neg = -(-reg_ref + inc)
    z = np.zeros((1,) + NC[::-1])
    mg = oracle.mg_default(rtol=1e-12, atol=1e-14)
    u, v, w, phi, rc, _ = oracle.mac_project((dxc,) * 3, z[0], z[0], z[0], rho[0], rhs_neg[0], z[0], 2.0 / dt, mg)
    assert rc == 0
    reg = C.c_void_p()
    lib.check(lib.iamrx_fluxreg_create(clev.h, flev.h, 1, C.byref(reg)))
    CF = [[to_fab(cflux[d], b, 0, t, dev) for b in cboxes] for d, t in enumerate((ix.XFACE, ix.YFACE, ix.ZFACE))]
    FF = [[to_fab(fflux[d], b, 0, t, dev) for b in fboxes] for d, t in enumerate((ix.XFACE, ix.YFACE, ix.ZFACE))]
    fa = lambda L: fab_array([p[1] for p in L])
    lib.check(lib.iamrx_fluxreg_reset(reg, stream_of(dev)))
    lib.check(lib.iamrx_fluxreg_crse_add(reg, fa(CF[0]), fa(CF[1]), fa(CF[2]), 1.0, vol, stream_of(dev)))
    lib.check(lib.iamrx_fluxreg_fine_add(reg, fa(FF[0]), fa(FF[1]), fa(FF[2]), 1.0, vol, stream_of(dev)))
    RH = [to_fab(rho, b, 1, ix.CELL, dev) for b in cboxes]
    IN = [to_fab(inc, b, 0, ix.CELL, dev) for b in cboxes]
    PH = [to_fab(z, b, 1, ix.CELL, dev) for b in cboxes]
    UC = [[to_fab(z, b, 0, t, dev) for b in cboxes] for t in (ix.XFACE, ix.YFACE, ix.ZFACE)]
    info = ix.MGInfo()
    lib.iamrx_mg_info_default(C.byref(info))
    info.rtol, info.atol = 1e-12, 1e-14
    rc = lib.iamrx_mac_sync_solve(clev.h, reg, fa(RH), fa(IN), fa(UC[0]), fa(UC[1]), fa(UC[2]), fa(PH), dt, None, None, C.byref(info),
                                  stream_of(dev))
    lib.check(rc)
    sync(dev)
    for d, (t, ref) in enumerate(zip((ix.XFACE, ix.YFACE, ix.ZFACE), (u, v, w))):
        got, dup = from_fabs([p[0] for p in UC[d]], cboxes, 0, t, NC, 1)
        assert dup <= 1e-12
        assert np.abs(got[0] + ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())       # Ucorr = -(0 - beta grad phi)
    gp, _ = from_fabs([p[0] for p in PH], cboxes, 1, ix.CELL, NC, 1)
    assert np.abs((gp[0] - gp[0].mean()) - (phi - phi.mean())).max() <= 1e-10 * max(1.0, np.abs(phi).max())
    lib.iamrx_fluxreg_destroy(reg)
    clev.close(); flev.close()


@pytest.mark.parametrize("layout", FINE_LAYOUTS)
@pytest.mark.parametrize("ixtype", [ix.CELL, ix.NODE, ix.XFACE])
def test_average_down_levels(backend, oracle, layout, ixtype):
    """amrex::average_down between two levels (NavierStokesBase::avgDown): coarse data replaced by the averaged fine data under the
    fine grids, untouched elsewhere; components outside [scomp, scomp + ncomp) untouched."""
    lib, dev = backend
    scomp, ncomp = 1, 2
    clev, flev, cboxes, fboxes = _level_pair(lib, layout)
    fine = hash_uniform(151 + ixtype, (scomp + ncomp,) + NF[::-1])
    crse = hash_uniform(161 + ixtype, (scomp + ncomp,) + NC[::-1])
    ref_avg = oracle.average_down(NC, ixtype, fine[scomp:])
    ext = {ix.CELL: (0, 0, 0), ix.NODE: (1, 1, 1), ix.XFACE: (1, 0, 0)}[ixtype]
    FI = [to_fab(fine, b, 0, ixtype, dev) for b in fboxes]
    CR = [to_fab(crse, b, 0, ixtype, dev) for b in cboxes]
    fa = lambda L: fab_array([p[1] for p in L])
    lib.check(lib.iamrx_average_down(flev.h, clev.h, fa(FI), fa(CR), scomp, ncomp, ixtype, stream_of(dev)))
    sync(dev)
    for (t, _), (blo, bhi) in zip(CR, cboxes):
        a = t.cpu().numpy()
        exp, _ = to_fab(crse, (blo, bhi), 0, ixtype, "cpu")
        exp = exp.numpy().copy()
        for lo, hi in layout:     # covered region (points of the coarsened fine box), clipped to this coarse box (periodic images not needed here)
            r_lo = [max(lo[d], blo[d]) for d in range(3)]
            r_hi = [min(hi[d] + ext[d], bhi[d] + ext[d]) for d in range(3)]
            if any(r_hi[d] < r_lo[d] for d in range(3)):
                continue
            sl_loc = tuple(slice(r_lo[d] - blo[d], r_hi[d] - blo[d] + 1) for d in (2, 1, 0))
            idx = [np.arange(r_lo[d], r_hi[d] + 1) % NC[d] for d in (2, 1, 0)]
            exp[(slice(scomp, None),) + sl_loc] = ref_avg[:, idx[0]][:, :, idx[1]][:, :, :, idx[2]]
        assert np.abs(a - exp).max() <= 1e-15
    clev.close(); flev.close()


def umac_grown_expected(filled, fmask, box, n, dx, divu=None):
    """numpy restatement of create_umac_grown's divergence correction (NSB.cpp:1203-1308) on one fine box: `filled` = the three
    face arrays of the whole fine index space after the face_linear FillPatchTwoLevels; returns the box's face arrays with one
    ghost cell ([k][j][i], origin lo - 1) and the number of corrected halo cells."""
    types = (ix.XFACE, ix.YFACE, ix.ZFACE)
    lo, hi = box
    exp = [to_fab(filled[d], (lo, hi), 1, types[d], "cpu")[0].numpy()[0].copy() for d in range(3)]   # [k][j][i], origin lo - 1
    U, V, W = exp
    with_divu = divu is not None
    def m(i, j, k):   # 0 interior, 1 covered, 2 not covered (global cell index, periodic)
        if all(lo[q] <= (i, j, k)[q] <= hi[q] for q in range(3)):
            return 0
        return 1 if fmask[k % n[2], j % n[1], i % n[0]] else 2
    fixed = 0
    for k in range(lo[2] - 1, hi[2] + 2):
        for j in range(lo[1] - 1, hi[1] + 2):
            for i in range(lo[0] - 1, hi[0] + 2):
                if m(i, j, k) != 2:
                    continue
                nb = [(i - 1, j, k), (i + 1, j, k), (i, j - 1, k), (i, j + 1, k), (i, j, k - 1), (i, j, k + 1)]
                if sum(m(*c) in (0, 1) for c in nb) != 1:
                    continue
                a, b, c = i - (lo[0] - 1), j - (lo[1] - 1), k - (lo[2] - 1)      # local cell index in the grown box
                dv = divu[0, k % n[2], j % n[1], i % n[0]] if with_divu else 0.0
                dux = (U[c, b, a + 1] - U[c, b, a]) / dx
                duy = (V[c, b + 1, a] - V[c, b, a]) / dx
                duz = (W[c + 1, b, a] - W[c, b, a]) / dx
                if i < lo[0] and m(i + 1, j, k) != 2:
                    U[c, b, a] = U[c, b, a + 1] + dx * (duy + duz - dv)
                elif i > hi[0] and m(i - 1, j, k) != 2:
                    U[c, b, a + 1] = U[c, b, a] - dx * (duy + duz - dv)
                if j < lo[1] and m(i, j + 1, k) != 2:
                    V[c, b, a] = V[c, b + 1, a] + dx * (dux + duz - dv)
                elif j > hi[1] and m(i, j - 1, k) != 2:
                    V[c, b + 1, a] = V[c, b, a] - dx * (dux + duz - dv)
                if k < lo[2] and m(i, j, k + 1) != 2:
                    W[c, b, a] = W[c + 1, b, a] + dx * (dux + duy - dv)
                elif k > hi[2] and m(i, j, k - 1) != 2:
                    W[c + 1, b, a] = W[c, b, a] - dx * (dux + duy - dv)
                fixed += 1
                div = (U[c, b, a + 1] - U[c, b, a] + V[c, b + 1, a] - V[c, b, a] + W[c + 1, b, a] - W[c, b, a]) / dx
                assert abs(div - dv) <= 1e-12 * max(1.0, abs(dv)) * n[0]
    return exp, fixed


@pytest.mark.parametrize("layout", FINE_LAYOUTS)
@pytest.mark.parametrize("with_divu", [0, 1])
def test_create_umac_grown(backend, oracle, layout, with_divu):
    """NavierStokesBase::create_umac_grown on the fine level (NSB.cpp:1108-1310): FillPatchTwoLevels of the MAC velocities with
    face_linear_interp into one ghost cell, then the divergence correction of the halo cells with exactly one valid / covered
    neighbour (numpy restatement of the reference loop); afterwards those cells satisfy div(u_mac) = divu."""
    lib, dev = backend
    clev, flev, cboxes, fboxes = _level_pair(lib, layout)
    types = (ix.XFACE, ix.YFACE, ix.ZFACE)
    cmask = np.zeros(NC[::-1], dtype=bool)
    for lo, hi in layout:
        cmask[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = True
    fmask = np.repeat(np.repeat(np.repeat(cmask, 2, 0), 2, 1), 2, 2)
    uc = [smooth_field(NC, 171 + d, 1) for d in range(3)]
    uf = [hash_uniform(181 + d, (1,) + NF[::-1]) for d in range(3)]
    divu = 0.3 * smooth_field(NF, 191, 1) if with_divu else None
    dx = 1.0 / NF[0]
    filled = []
    for d in range(3):
        interp = oracle.interp(2 + d, NC, uc[d])
        fface = fmask | np.roll(fmask, 1, 2 - d)            # faces of fine cells: the cell above or below the face is fine
        filled.append(np.where(fface[None], uf[d], interp))
    UC = [[to_fab(uc[d], b, 0, types[d], dev) for b in cboxes] for d in range(3)]
    UF = [[to_fab(uf[d], b, 1, types[d], dev, fill_ghost=False) for b in fboxes] for d in range(3)]
    DV = [to_fab(divu, b, 1, ix.CELL, dev) for b in fboxes] if with_divu else None
    fa = lambda L: fab_array([p[1] for p in L])
    lib.check(lib.iamrx_create_umac_grown(flev.h, clev.h, fa(UF[0]), fa(UF[1]), fa(UF[2]), fa(UC[0]), fa(UC[1]), fa(UC[2]),
                                          fa(DV) if with_divu else None, stream_of(dev)))
    sync(dev)
    for ib, (lo, hi) in enumerate(fboxes):
        exp, fixed = umac_grown_expected(filled, fmask, (lo, hi), NF, dx, divu)
        assert fixed > 0
        for d in range(3):
            got = UF[d][ib][0].cpu().numpy()[0]
            assert np.abs(got - exp[d]).max() <= 1e-13, (d, ib)
    clev.close(); flev.close()
