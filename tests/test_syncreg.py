"""SyncRegister (Source/SyncRegister.cpp) through the C ABI vs oracle/syncreg.py, a restatement in the reference's own
per-face / per-grid data structure: CrseInit, FineAdd (tent restriction, edge / corner weights, wall doubling, periodic images),
InitRHS (outflow planes, the interior mask with the reference's 3-D threshold and with the 2-D expression's generalisation)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import iamr_b200 as ix
from util import hash_uniform, fab_array, stream_of, sync

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import syncreg as orc_syncreg  # noqa: E402

NC, NF = (8, 8, 8), (16, 16, 16)
WHOLE = ((0, 0, 0), (7, 7, 7))

# (periodicity, fine boxes in fine index space, phys_lo, phys_hi)
CASES = [
    ((1, 1, 1), [((4, 4, 4), (11, 11, 11))], (0, 0, 0), (0, 0, 0)),                                   # one interior grid
    ((1, 1, 1), [((4, 4, 4), (7, 11, 11)), ((8, 4, 4), (11, 11, 11))], (0, 0, 0), (0, 0, 0)),          # two grids sharing a plane
    ((1, 1, 1), [((4, 4, 4), (7, 7, 11)), ((8, 4, 4), (11, 7, 11)), ((4, 8, 4), (7, 11, 11))], (0, 0, 0), (0, 0, 0)),   # L shape
    ((0, 0, 0), [((0, 4, 4), (7, 11, 15))], (5, 5, 5), (5, 5, 5)),                                   # against the low x and high z walls
    ((0, 1, 0), [((8, 4, 4), (15, 11, 11))], (1, 0, 4), (2, 0, 4)),                                   # against the outflow side (x high)
    ((1, 1, 1), [((4, 4, 12), (11, 11, 15)), ((4, 4, 0), (11, 11, 3))], (0, 0, 0), (0, 0, 0)),         # two grids joined across the periodic z boundary
]


def _node_fab(arr, lo, dev):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr[None])).to(dev)
    return t, ix.fab_of(t, list(lo))


@pytest.mark.parametrize("per,fboxes,plo,phi", CASES)
@pytest.mark.parametrize("split", [False, True])
def test_sync_register(backend, per, fboxes, plo, phi, split):
    lib, dev = backend
    cboxes = [WHOLE] if not split else [((0, 0, 0), (3, 7, 7)), ((4, 0, 0), (7, 7, 7))]
    clev = ix.Level(lib, ix.Geom.make(NC, periodic=per), cboxes)
    flev = ix.Level(lib, ix.Geom.make(NF, periodic=per), fboxes)
    rc_full = hash_uniform(41, (NC[2] + 1, NC[1] + 1, NC[0] + 1)) - 0.5
    for d in range(3):          # a periodic field: the high node plane repeats the low one
        if per[d]:
            sl_hi = [slice(None)] * 3; sl_hi[2 - d] = NC[d]
            sl_lo = [slice(None)] * 3; sl_lo[2 - d] = 0
            rc_full[tuple(sl_hi)] = rc_full[tuple(sl_lo)]
    rf = [hash_uniform(50 + g, (hi[2] - lo[2] + 2, hi[1] - lo[1] + 2, hi[0] - lo[0] + 2)) - 0.5 for g, (lo, hi) in enumerate(fboxes)]
    m_c, m_f = 1.0, 0.5
    for maxcount in (0.0, 7.5):
        ref = orc_syncreg.SyncRegister(NC, per, fboxes)
        ref.crse_init(rc_full, m_c)
        ref.fine_add(fboxes, rf, m_f)
        expect = ref.init_rhs(plo, phi, maxcount if maxcount > 0 else 26.5)
        h = C.c_void_p()
        lib.check(lib.iamrx_syncreg_create(clev.h, flev.h, maxcount, C.byref(h)))
        CR = [_node_fab(rc_full[lo[2]:hi[2] + 2, lo[1]:hi[1] + 2, lo[0]:hi[0] + 2], lo, dev) for lo, hi in cboxes]
        FR = [_node_fab(a, lo, dev) for a, (lo, hi) in zip(rf, fboxes)]
        RH = [_node_fab(np.full((hi[2] - lo[2] + 2, hi[1] - lo[1] + 2, hi[0] - lo[0] + 2), 9.0), lo, dev) for lo, hi in cboxes]
        fa = lambda L: fab_array([p[1] for p in L])
        st = stream_of(dev)
        lib.check(lib.iamrx_syncreg_crse_init(h, fa(CR), m_c, st))
        lib.check(lib.iamrx_syncreg_fine_add(h, fa(FR), m_f, st))
        lib.check(lib.iamrx_syncreg_init_rhs(h, fa(RH), (C.c_int * 3)(*plo), (C.c_int * 3)(*phi), st))
        sync(dev)
        assert np.abs(expect).max() > 0.1
        for (t, _), (lo, hi) in zip(RH, cboxes):
            got = t.cpu().numpy()[0]
            assert np.abs(got - expect[lo[2]:hi[2] + 2, lo[1]:hi[1] + 2, lo[0]:hi[0] + 2]).max() <= 1e-14
        # the inputs are not scaled in place (the reference's .mult(mult) is not reproduced)
        assert np.array_equal(FR[0][0].cpu().numpy()[0], rf[0])
        lib.check(lib.iamrx_syncreg_destroy(h))
    clev.close(); flev.close()


def test_fine_add_conserves_the_surface_sum(backend):
    """A property of FineAdd that needs no oracle: the tent restriction is conservative.  With a fine residual of 1 on every node of
    one interior grid, each fine SURFACE node ends up counted exactly once -- 1 on a face, 1/2 on each of the two planes an edge
    node lies on, 1/3 on each of the three planes of a corner -- and spreads r_dir / prod(r^2) * (tent sum 4 * 4 / ... ) = 1 / r^3 of
    itself over the coarse nodes: the register sums to (number of fine surface nodes) / 8."""
    lib, dev = backend
    per = (1, 1, 1)
    fboxes = [((4, 4, 4), (11, 11, 11))]
    clev = ix.Level(lib, ix.Geom.make(NC, periodic=per), [WHOLE])
    flev = ix.Level(lib, ix.Geom.make(NF, periodic=per), fboxes)
    h = C.c_void_p()
    lib.check(lib.iamrx_syncreg_create(clev.h, flev.h, 0.0, C.byref(h)))
    ones = np.ones((9, 9, 9))
    FR = [_node_fab(ones, (4, 4, 4), dev)]
    CR = [_node_fab(np.zeros((9, 9, 9)), (0, 0, 0), dev)]
    RH = [_node_fab(np.zeros((9, 9, 9)), (0, 0, 0), dev)]
    fa = lambda L: fab_array([p[1] for p in L])
    st = stream_of(dev)
    lib.check(lib.iamrx_syncreg_crse_init(h, fa(CR), 1.0, st))
    lib.check(lib.iamrx_syncreg_fine_add(h, fa(FR), 1.0, st))
    lib.check(lib.iamrx_syncreg_init_rhs(h, fa(RH), None, None, st))
    sync(dev)
    got = RH[0][0].cpu().numpy()[0]
    ref = orc_syncreg.SyncRegister(NC, per, fboxes)
    ref.crse_init(np.zeros((9, 9, 9)), 1.0)
    ref.fine_add(fboxes, [ones], 1.0)
    assert abs(got.sum() - ref.init_rhs().sum()) <= 1e-12
    n_surface = 9 ** 3 - 7 ** 3
    assert abs(got.sum() - n_surface / 8.0) <= 1e-12
    lib.check(lib.iamrx_syncreg_destroy(h))
    clev.close(); flev.close()
