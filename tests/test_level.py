"""BoxArray / FillBoundary: the copy plan (pure host logic, product library) and the data
movement (emulation on CPU, CUDA on GPU) against numpy periodic wrap."""
import ctypes as C

import numpy as np
import pytest

import iamr_b200 as ix
from util import hash_uniform, split_boxes, to_fab, fab_array, stream_of, sync, IX_EXT


def _plan(lib, lev, ixtype, ng):
    cap = 4096
    db = (C.c_int * cap)(); sb = (C.c_int * cap)(); rg = (C.c_int * (6 * cap))(); sh = (C.c_int * (3 * cap))()
    n = lib.iamrx_debug_fb_plan(lev.h, ixtype, ng, cap, db, sb, rg, sh)
    assert 0 <= n <= cap
    return [(db[r], sb[r], tuple(rg[6 * r:6 * r + 3]), tuple(rg[6 * r + 3:6 * r + 6]), tuple(sh[3 * r:3 * r + 3])) for r in range(n)]


@pytest.mark.parametrize("ixtype", [ix.CELL, ix.XFACE, ix.ZFACE, ix.NODE])
@pytest.mark.parametrize("nb,ng", [((1, 1, 1), 1), ((2, 2, 2), 1), ((2, 1, 2), 3), ((4, 2, 1), 2)])
def test_fb_plan_covers_every_ghost_point_once(ixtype, nb, ng):
    lib = ix.load()  # host logic of the PRODUCT library; needs no device
    n = (16, 8, 8)
    g = ix.Geom.make(n)
    boxes = split_boxes(n, nb)
    lev = ix.Level(lib, g, boxes)
    plan = _plan(lib, lev, ixtype, ng)
    ext = IX_EXT[ixtype]
    for bi, (lo, hi) in enumerate(boxes):
        shape = tuple(hi[d] - lo[d] + 1 + ext[d] + 2 * ng for d in range(3))
        cnt = np.zeros(shape, dtype=int)
        for (d_, s_, rlo, rhi, shf) in plan:
            if d_ != bi:
                continue
            sl = tuple(slice(rlo[d] - (lo[d] - ng), rhi[d] - (lo[d] - ng) + 1) for d in range(3))
            cnt[sl] += 1
            # source region lies inside the source box's valid points
            slo, shi = boxes[s_]
            for d in range(3):
                assert slo[d] <= rlo[d] + shf[d] and rhi[d] + shf[d] <= shi[d] + ext[d]
                assert shf[d] % n[d] == 0
        inner = tuple(slice(ng, shape[d] - ng) for d in range(3))
        assert (cnt[inner] == 0).all()
        cnt[inner] = 1
        assert (cnt == 1).all()  # fully periodic: every ghost point has exactly one source
    lev.close()


@pytest.mark.parametrize("ixtype,ncomp,ng", [(ix.CELL, 3, 3), (ix.CELL, 1, 1), (ix.YFACE, 1, 1), (ix.NODE, 1, 1)])
@pytest.mark.parametrize("nb", [(1, 1, 1), (2, 2, 2), (1, 4, 2)])
def test_fill_boundary(backend, ixtype, ncomp, ng, nb):
    lib, dev = backend
    n = (16, 16, 8)
    g = ix.Geom.make(n)
    boxes = split_boxes(n, nb)
    lev = ix.Level(lib, g, boxes)
    dense = hash_uniform(7, (ncomp, n[2], n[1], n[0]))
    pairs = [to_fab(dense, b, ng, ixtype, dev, fill_ghost=False) for b in boxes]
    want = [to_fab(dense, b, ng, ixtype, "cpu", fill_ghost=True)[0] for b in boxes]
    fabs = fab_array([p[1] for p in pairs])
    lib.check(lib.iamrx_fill_boundary(lev.h, fabs, ixtype, ncomp, ng, stream_of(dev)))
    sync(dev)
    for (t, _), w in zip(pairs, want):
        assert np.array_equal(t.cpu().numpy(), w.numpy())
    lev.close()
