"""Worker of tests/test_multirank.py::test_two_level_blocks_two_ranks: one rank of a world_size-2 gloo job.  The collective two-level
building blocks (FillPatchTwoLevels, average_down, SyncInterp, SyncProjInterp, create_umac_grown) with the coarse and the fine boxes
owned by DIFFERENT ranks, against the same references the single-rank tests use (oracle interpolaters + numpy)."""
import ctypes as C
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import iamr_b200 as ix  # noqa: E402
import orc  # noqa: E402
from mr_worker import EXCH, ARED, exchange, allreduce  # noqa: E402
from util import split_boxes, hash_uniform, smooth_field, to_fab, fab_array, stream_of  # noqa: E402

NC, NF = (8, 8, 8), (16, 16, 16)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = ix.load(os.path.join(ROOT, "tests", "emul", "_build", "libiamrx_emul.so"))
    ex, ar = EXCH(exchange), ARED(allreduce)
    lib.check(lib.iamrx_comm_set_transport(rank, world, C.cast(ex, C.c_void_p), C.cast(ar, C.c_void_p), None))
    orc.lib()
    st = stream_of("cpu")
    # coarse level: two boxes, rank 0 / rank 1; fine level: two abutting boxes spanning periodic y, owned the OTHER way round
    cboxes = split_boxes(NC, (2, 1, 1)); cown = [0, 1]
    layout = [((2, 0, 2), (3, 7, 5)), ((4, 0, 2), (5, 7, 5))]
    fboxes = [(tuple(2 * l for l in lo), tuple(2 * h + 1 for h in hi)) for lo, hi in layout]; fown = [1, 0]
    clev = ix.Level(lib, ix.Geom.make(NC), cboxes, cown)
    flev = ix.Level(lib, ix.Geom.make(NF), fboxes, fown)
    cmine = [i for i, o in enumerate(cown) if o == rank]
    fmine = [i for i, o in enumerate(fown) if o == rank]
    cmask = np.zeros(NC[::-1], dtype=bool)
    for lo, hi in layout:
        cmask[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = True
    fmask = np.repeat(np.repeat(np.repeat(cmask, 2, 0), 2, 1), 2, 2)
    fa = lambda L: fab_array([p[1] for p in L])

    # 1. FillPatchTwoLevels
    ncomp, ngrow = 2, 3
    c_old, c_new = smooth_field(NC, 91, ncomp), smooth_field(NC, 92, ncomp)
    fdat = hash_uniform(94, (ncomp,) + NF[::-1])
    w = 0.3
    expect = np.where(fmask[None], fdat, orc.interp(0, NC, (1.0 - w) * c_old + w * c_new))
    CO = [to_fab(c_old, cboxes[i], 0, ix.CELL, "cpu") for i in cmine]
    CN = [to_fab(c_new, cboxes[i], 0, ix.CELL, "cpu") for i in cmine]
    FI = [to_fab(fdat, fboxes[i], ngrow, ix.CELL, "cpu", fill_ghost=False) for i in fmine]
    lib.check(lib.iamrx_fillpatch_two_levels(flev.h, clev.h, fa(FI), fa(CO), fa(CN), 0.0, 1.0, w, ncomp, ngrow, None, None, st))
    for (t, _), i in zip(FI, fmine):
        ref, _ = to_fab(expect, fboxes[i], ngrow, ix.CELL, "cpu")
        assert np.abs(t.numpy() - ref.numpy()).max() <= 1e-14, ("fillpatch", rank)

    # 2. average_down (cells and nodes)
    for ixtype in (ix.CELL, ix.NODE):
        fine = hash_uniform(151 + ixtype, (2,) + NF[::-1])
        crse = hash_uniform(161 + ixtype, (2,) + NC[::-1])
        avg = orc.average_down(NC, ixtype, fine)
        ext = 1 if ixtype == ix.NODE else 0
        FI = [to_fab(fine, fboxes[i], 0, ixtype, "cpu") for i in fmine]
        CR = [to_fab(crse, cboxes[i], 0, ixtype, "cpu") for i in cmine]
        lib.check(lib.iamrx_average_down(flev.h, clev.h, fa(FI), fa(CR), 0, 2, ixtype, st))
        for (t, _), i in zip(CR, cmine):
            blo, bhi = cboxes[i]
            exp = to_fab(crse, cboxes[i], 0, ixtype, "cpu")[0].numpy().copy()
            for lo, hi in layout:
                r_lo = [max(lo[d], blo[d]) for d in range(3)]
                r_hi = [min(hi[d] + ext, bhi[d] + ext) for d in range(3)]
                if any(r_hi[d] < r_lo[d] for d in range(3)):
                    continue
                sl = tuple(slice(r_lo[d] - blo[d], r_hi[d] - blo[d] + 1) for d in (2, 1, 0))
                idx = [np.arange(r_lo[d], r_hi[d] + 1) % NC[d] for d in (2, 1, 0)]
                exp[(slice(None),) + sl] = avg[:, idx[0]][:, :, idx[1]][:, :, :, idx[2]]
            assert np.abs(t.numpy() - exp).max() <= 1e-15, ("average_down", ixtype, rank)

    # 3. SyncInterp (cell-conservative, increment) and SyncProjInterp
    crse = smooth_field(NC, 101, 2)
    fine = hash_uniform(103, (2,) + NF[::-1])
    dt = 0.37
    expect = fine + dt * orc.interp(0, NC, crse)
    CS = [to_fab(crse, cboxes[i], 0, ix.CELL, "cpu") for i in cmine]
    FS = [to_fab(fine, fboxes[i], 0, ix.CELL, "cpu") for i in fmine]
    lib.check(lib.iamrx_sync_interp(flev.h, clev.h, fa(FS), 0, fa(CS), 0, 2, 1, dt, ix.SYNC_CELL_CONS, None, st))
    for (t, _), i in zip(FS, fmine):
        ref, _ = to_fab(expect, fboxes[i], 0, ix.CELL, "cpu")
        assert np.abs(t.numpy() - ref.numpy()).max() <= 2e-15, ("sync_interp", rank)
    phi = smooth_field(NC, 111, 1)
    pn, po = hash_uniform(112, (1,) + NF[::-1]), hash_uniform(113, (1,) + NF[::-1])
    interp = orc.interp(1, NC, phi)
    PH = [to_fab(phi, cboxes[i], 0, ix.NODE, "cpu") for i in cmine]
    PN = [to_fab(pn, fboxes[i], 0, ix.NODE, "cpu") for i in fmine]
    PO = [to_fab(po, fboxes[i], 0, ix.NODE, "cpu") for i in fmine]
    lib.check(lib.iamrx_sync_proj_interp(flev.h, clev.h, fa(PN), fa(PO), fa(PH), st))
    for arr, base in ((PN, pn), (PO, po)):
        for (t, _), i in zip(arr, fmine):
            ref, _ = to_fab(base + interp, fboxes[i], 0, ix.NODE, "cpu")
            assert np.abs(t.numpy() - ref.numpy()).max() <= 2e-15, ("sync_proj_interp", rank)

    # 4. create_umac_grown: the FillPatchTwoLevels part (ghost faces owned by the fine neighbour on the other rank / interpolated from
    #    coarse data on the other rank); the halo correction itself is local and covered by tests/test_amr.py -- here the cells it
    #    fixes must come out divergence-free
    types = (ix.XFACE, ix.YFACE, ix.ZFACE)
    uc = [smooth_field(NC, 171 + d, 1) for d in range(3)]
    uf = [hash_uniform(181 + d, (1,) + NF[::-1]) for d in range(3)]
    UC = [[to_fab(uc[d], cboxes[i], 0, types[d], "cpu") for i in cmine] for d in range(3)]
    UF = [[to_fab(uf[d], fboxes[i], 1, types[d], "cpu", fill_ghost=False) for i in fmine] for d in range(3)]
    lib.check(lib.iamrx_create_umac_grown(flev.h, clev.h, fa(UF[0]), fa(UF[1]), fa(UF[2]), fa(UC[0]), fa(UC[1]), fa(UC[2]), None, st))
    dx = 1.0 / NF[0]
    for q, i in enumerate(fmine):
        lo, hi = fboxes[i]
        U, V, W = (UF[d][q][0].numpy()[0] for d in range(3))
        assert max(np.abs(U).max(), np.abs(V).max(), np.abs(W).max()) < 1e30, ("create_umac_grown: unfilled ghost faces", rank)
        # valid faces untouched
        for d in range(3):
            ref = to_fab(uf[d], fboxes[i], 0, types[d], "cpu")[0].numpy()[0]
            got = UF[d][q][0].numpy()[0][1:-1, 1:-1, 1:-1]
            assert np.array_equal(got, ref), ("create_umac_grown: valid faces", rank)
        # ghost cells beyond the x faces of the pair that the fine level does not cover (x = lo-1 of the first box / hi+1 of the second):
        # exactly one valid neighbour -> divergence-free after the correction
        a = 0 if i == 0 else U.shape[2] - 2        # local cell index of the ghost column on the uncovered x side
        c, b = slice(1, W.shape[0] - 2), slice(1, V.shape[1] - 2)
        div = (U[c, b, a + 1] - U[c, b, a]) / dx + (V[c, 2:V.shape[1] - 1, a] - V[c, 1:V.shape[1] - 2, a]) / dx \
            + (W[2:W.shape[0] - 1, b, a] - W[1:W.shape[0] - 2, b, a]) / dx
        assert np.abs(div).max() <= 1e-11, ("create_umac_grown: divergence of the corrected halo", rank, float(np.abs(div).max()))

    # 5. the level > 0 MAC solve: setCoarseFineBC (collective: coarse data on the other rank) + coarse-fine Dirichlet sides in the
    #    multigrid, fine boxes on different ranks, against the oracle solving on the patch as its own domain (tests/test_twolevel.py)
    from test_bc import _mg
    CF, PER_ = 5, 0
    flo, fhi = (4, 0, 4), (11, 15, 11)
    n = tuple(fhi[d] - flo[d] + 1 for d in range(3))
    dxf = tuple(1.0 / m for m in NF)
    cov = cmask
    cphi = 0.05 * smooth_field(NC, 191, 1)
    rho = 1.0 + 0.3 * hash_uniform(192, (1,) + NF[::-1])
    macs = [smooth_field(NF, 193 + d, 1) + 0.3 * hash_uniform(196 + d, (1,) + NF[::-1]) for d in range(3)]
    wp = lambda a, g: np.pad(a, ((0, 0), (g, g), (g, g), (g, g)), mode="wrap")
    cut = lambda P, g, ng: np.ascontiguousarray(P[:, flo[2] - ng + g:fhi[2] + ng + g + 1, flo[1] - ng + g:fhi[1] + ng + g + 1, flo[0] - ng + g:fhi[0] + ng + g + 1])
    phi0 = orc.interp_bndry(NC, (1, 1, 1), cphi, cov, flo, fhi, np.zeros((1, n[2] + 2, n[1] + 2, n[0] + 2)))
    mgo = orc.mg_default(rtol=1e-12)
    ru, rv, rw, rphi, rc, mgo = orc.mac_project_bc(n, (0, 1, 0), dxf, cut(wp(macs[0], 2), 2, 1), cut(wp(macs[1], 2), 2, 1), cut(wp(macs[2], 2), 2, 1),
                                                   cut(wp(rho, 1), 1, 1), None, phi0, 3.0, (CF, PER_, CF), (CF, PER_, CF), 4, mgo)
    assert rc == 0
    CP = [to_fab(np.where(cov[None], 1.0e30, cphi), cboxes[i], 0, ix.CELL, "cpu") for i in cmine]
    UF = [[to_fab(macs[d], fboxes[i], 1, types[d], "cpu") for i in fmine] for d in range(3)]
    RH = [to_fab(rho, fboxes[i], 1, ix.CELL, "cpu") for i in fmine]
    PH = [to_fab(np.zeros((1,) + NF[::-1]), fboxes[i], 1, ix.CELL, "cpu") for i in fmine]
    lib.check(lib.iamrx_set_coarse_fine_bc(flev.h, clev.h, fa(PH), fa(CP), 1, st))
    info = _mg(lib, rtol=1e-12, maxorder=4)
    lib.check(lib.iamrx_mac_project(flev.h, fa(UF[0]), fa(UF[1]), fa(UF[2]), fa(RH), None, fa(PH), 3.0, None, None, C.byref(info), st))
    for (t, _), i in zip(PH, fmine):
        lo, hi = fboxes[i]
        got = t.numpy()[:, 1:-1, 1:-1, 1:-1]
        ref = rphi[:, 1 + lo[2] - flo[2]:2 + hi[2] - flo[2], 1 + lo[1] - flo[1]:2 + hi[1] - flo[1], 1 + lo[0] - flo[0]:2 + hi[0] - flo[0]]
        assert np.abs(got - ref).max() <= 2e-10, ("coarse-fine mac_project", rank, float(np.abs(got - ref).max()))

    # 5b. the level > 0 nodal projection with the fine boxes on different ranks: FillCoarsePatch of the pressure (coarse data on the
    #     other rank), interior zeroed, coarse-fine boundary nodes kept as Dirichlet data -- against the oracle on the patch
    DIR_ = 1
    cpress = 0.02 * smooth_field(NC, 221, 1)
    velg = smooth_field(NF, 222, 3) + 0.2 * hash_uniform(223, (3,) + NF[::-1])
    sigg = 1.0 / (1.0 + 0.3 * hash_uniform(224, (1,) + NF[::-1]))
    Gn = wp(orc.interp(1, NC, cpress), 2)
    Pg = np.zeros_like(Gn)
    ilo, ihi = [flo[d] + 2 for d in range(3)], [fhi[d] + 1 + 2 for d in range(3)]
    for d in range(3):
        for pl in (ilo[d], ihi[d]):
            sl = [slice(None), slice(ilo[2], ihi[2] + 1), slice(ilo[1], ihi[1] + 1), slice(ilo[0], ihi[0] + 1)]
            sl[3 - d] = slice(pl, pl + 1)
            Pg[tuple(sl)] = Gn[tuple(sl)]
    cutg = lambda P, g, ng: np.ascontiguousarray(P[:, flo[2] - ng + g:fhi[2] + ng + g + 1, flo[1] - ng + g:fhi[1] + ng + g + 1, flo[0] - ng + g:fhi[0] + ng + g + 1])
    mgn = orc.mg_default(rtol=1e-12)
    rvel, rphi_n, rgp, rcn, mgn = orc.nodal_project_bc(n, (0, 1, 0), dxf, cutg(wp(velg, 1), 1, 1), cutg(sigg, 0, 0), cutg(Pg, 2, 2), (DIR_, PER_, DIR_), (DIR_, PER_, DIR_), mgn)
    assert rcn == 0
    VV = [to_fab(velg, fboxes[i], 1, ix.CELL, "cpu") for i in fmine]
    SG = [to_fab(sigg, fboxes[i], 0, ix.CELL, "cpu") for i in fmine]
    PN = [to_fab(np.zeros((1,) + NF[::-1]), fboxes[i], 1, ix.NODE, "cpu") for i in fmine]
    CPN = [to_fab(cpress, cboxes[i], 0, ix.NODE, "cpu") for i in cmine]
    lib.check(lib.iamrx_fill_coarse_patch_nodal(flev.h, clev.h, fa(PN), None, fa(CPN), 0.0, 1.0, 0.5, st))
    for t, _ in PN:
        t[:, 2:-2, 2:-2, 2:-2] = 0.0
    infon = _mg(lib, rtol=1e-12)
    lib.check(lib.iamrx_nodal_project(flev.h, fa(VV), fa(SG), fa(PN), None, 0, None, None, C.byref(infon), st))
    for (t, _), i in zip(VV, fmine):
        lo, hi = fboxes[i]
        got = t.numpy()[:, 1:-1, 1:-1, 1:-1]
        ref = rvel[:, 1 + lo[2] - flo[2]:2 + hi[2] - flo[2], 1 + lo[1] - flo[1]:2 + hi[1] - flo[1], 1 + lo[0] - flo[0]:2 + hi[0] - flo[0]]
        assert np.abs(got - ref).max() <= 1e-9, ("coarse-fine nodal_project", rank, float(np.abs(got - ref).max()))

    # 5c. the flux register with the coarse and the fine side of the interface on different ranks: CrseAdd is local, FineAdd goes
    #     through one replicated accumulation + all-reduce, Reflux is local -- against the oracle's register
    cflux = [hash_uniform(231 + d, (2,) + NC[::-1]) for d in range(3)]
    fflux = [hash_uniform(241 + d, (2,) + NF[::-1]) for d in range(3)]
    dtr, volr = 0.05, (1.0 / NC[0]) ** 3
    refreg = orc.fluxreg(NC, cmask, cflux, fflux, dtr, volr)
    freg = C.c_void_p()
    lib.check(lib.iamrx_fluxreg_create(clev.h, flev.h, 2, C.byref(freg)))
    CFX = [[to_fab(cflux[d], cboxes[i], 0, types[d], "cpu") for i in cmine] for d in range(3)]
    FFX = [[to_fab(fflux[d], fboxes[i], 0, types[d], "cpu") for i in fmine] for d in range(3)]
    lib.check(lib.iamrx_fluxreg_reset(freg, st))
    lib.check(lib.iamrx_fluxreg_crse_add(freg, fa(CFX[0]), fa(CFX[1]), fa(CFX[2]), dtr, volr, st))
    lib.check(lib.iamrx_fluxreg_fine_add(freg, fa(FFX[0]), fa(FFX[1]), fa(FFX[2]), dtr, volr, st))
    state = hash_uniform(251, (2,) + NC[::-1])
    STR = [to_fab(state, cboxes[i], 0, ix.CELL, "cpu") for i in cmine]
    lib.check(lib.iamrx_fluxreg_reflux(freg, fa(STR), 0, 1.0, st))
    for (t, _), i in zip(STR, cmine):
        lo, hi = cboxes[i]
        exp = (state + refreg)[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
        assert np.abs(t.numpy() - exp).max() <= 1e-12 * max(1.0, np.abs(refreg).max()), ("flux register across ranks", rank)
    assert np.abs(refreg).max() > 0.1
    lib.iamrx_fluxreg_destroy(freg)

    # 6. SyncRegister: CrseInit on the coarse boxes of both ranks, FineAdd from fine boxes on both ranks (one replicated accumulation
    #    + all-reduce), InitRHS -- against oracle/syncreg.py
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import syncreg as orc_syncreg
    import torch
    rc_full = hash_uniform(201, (NC[2] + 1, NC[1] + 1, NC[0] + 1)) - 0.5
    for d in range(3):
        hi_, lo_ = [slice(None)] * 3, [slice(None)] * 3
        hi_[2 - d], lo_[2 - d] = NC[d], 0
        rc_full[tuple(hi_)] = rc_full[tuple(lo_)]
    rf = [hash_uniform(210 + g, (hi[2] - lo[2] + 2, hi[1] - lo[1] + 2, hi[0] - lo[0] + 2)) - 0.5 for g, (lo, hi) in enumerate(fboxes)]
    ref = orc_syncreg.SyncRegister(NC, (1, 1, 1), fboxes)
    ref.crse_init(rc_full, 1.0)
    ref.fine_add(fboxes, rf, 0.5)
    expect = ref.init_rhs((0, 0, 0), (0, 0, 0), 7.5)
    def nfab(arr, lo):
        t = torch.from_numpy(np.ascontiguousarray(arr[None]))
        return t, ix.fab_of(t, list(lo))
    h = C.c_void_p()
    lib.check(lib.iamrx_syncreg_create(clev.h, flev.h, 7.5, C.byref(h)))
    CR = [nfab(rc_full[cboxes[i][0][2]:cboxes[i][1][2] + 2, cboxes[i][0][1]:cboxes[i][1][1] + 2, cboxes[i][0][0]:cboxes[i][1][0] + 2], cboxes[i][0]) for i in cmine]
    FR = [nfab(rf[i], fboxes[i][0]) for i in fmine]
    RH = [nfab(np.zeros((cboxes[i][1][2] - cboxes[i][0][2] + 2, cboxes[i][1][1] - cboxes[i][0][1] + 2, cboxes[i][1][0] - cboxes[i][0][0] + 2)), cboxes[i][0]) for i in cmine]
    lib.check(lib.iamrx_syncreg_crse_init(h, fa(CR), 1.0, st))
    lib.check(lib.iamrx_syncreg_fine_add(h, fa(FR), 0.5, st))
    lib.check(lib.iamrx_syncreg_init_rhs(h, fa(RH), None, None, st))
    for (t, _), i in zip(RH, cmine):
        lo, hi = cboxes[i]
        assert np.abs(t.numpy()[0] - expect[lo[2]:hi[2] + 2, lo[1]:hi[1] + 2, lo[0]:hi[0] + 2]).max() <= 1e-14, ("sync register", rank)
    lib.check(lib.iamrx_syncreg_destroy(h))

    clev.close(); flev.close()
    dist.barrier()
    print(f"rank {rank} ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
