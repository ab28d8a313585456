"""Worker of tests/test_multirank.py: one rank of a world_size-2 gloo job driving the host
emulation build of libiamrx through the host-transport hook (iamrx_comm_set_transport)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import iamr_b200 as ix  # noqa: E402
import orc  # noqa: E402
from util import split_boxes, hash_uniform, to_fab, fab_array, stream_of  # noqa: E402

EXCH = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                   C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p)
ARED = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p)


def _view(ptr, n):
    return torch.from_numpy(np.ctypeslib.as_array((C.c_double * n).from_address(ptr)))


def exchange(ctx, npeers, peers, sbuf, scount, rbuf, rcount, stream):
    try:
        reqs = []
        for i in range(npeers):
            if scount[i] > 0:
                reqs.append(dist.isend(_view(sbuf[i], scount[i]), peers[i]))
            if rcount[i] > 0:
                reqs.append(dist.irecv(_view(rbuf[i], rcount[i]), peers[i]))
        for r in reqs:
            r.wait()
        return 0
    except Exception as e:  # noqa: BLE001
        print("exchange failed:", e, flush=True)
        return 1


def allreduce(ctx, buf, n, op, stream):
    try:
        t = _view(buf, n)
        dist.all_reduce(t, op={0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}[op])
        return 0
    except Exception as e:  # noqa: BLE001
        print("allreduce failed:", e, flush=True)
        return 1


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = ix.load(os.path.join(ROOT, "tests", "emul", "_build", "libiamrx_emul.so"))
    ex, ar = EXCH(exchange), ARED(allreduce)
    lib.check(lib.iamrx_comm_set_transport(rank, world, C.cast(ex, C.c_void_p), C.cast(ar, C.c_void_p), None))
    assert lib.iamrx_comm_rank() == rank and lib.iamrx_comm_size() == world
    n = (16, 16, 16)
    boxes = split_boxes(n, (2, 2, 1))
    owners = [0, 1, 1, 0]
    g = ix.Geom.make(n)
    lev = ix.Level(lib, g, boxes, owners)
    mine = [i for i, o in enumerate(owners) if o == rank]
    assert lev.num_local() == len(mine)

    # 1. FillBoundary across ranks (cell ng=3, nodal ng=1) vs periodic wrap
    for ixtype, ncomp, ng in ((ix.CELL, 2, 3), (ix.NODE, 1, 1), (ix.XFACE, 1, 1)):
        dense = hash_uniform(11 + ixtype, (ncomp, n[2], n[1], n[0]))
        pairs = [to_fab(dense, boxes[i], ng, ixtype, "cpu", fill_ghost=False) for i in mine]
        want = [to_fab(dense, boxes[i], ng, ixtype, "cpu", fill_ghost=True)[0] for i in mine]
        lib.check(lib.iamrx_fill_boundary(lev.h, fab_array([p[1] for p in pairs]), ixtype, ncomp, ng, stream_of("cpu")))
        for (t, _), w in zip(pairs, want):
            assert np.array_equal(t.numpy(), w.numpy()), (rank, ixtype)

    # 2. the full step on 2 ranks vs the single-box oracle
    kw = dict(visc_coef=1e-3, cfl=0.7, gravity=-0.5)
    ns = ix.NavierStokes(lib, lev, "cpu", **kw)
    o = orc.OracleNS(n, **kw)
    pp = [1.0, 1.0, 1.0, 1.0, 1.0]
    ns.init_prob(100, pp); o.init_prob(100, pp)
    d1, d2 = ns.post_init(), o.post_init()
    assert abs(d1 - d2) <= 1e-13 * d2
    for _ in range(2):
        a, b = ns.step(), o.step()
        assert abs(a - b) <= 1e-12 * b
    So = o.get(0)
    err = 0.0
    for il, gi in enumerate(mine):
        lo, hi = boxes[gi]
        t = ns.field(0, il).numpy()
        err = max(err, np.abs(t - So[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]).max())
    assert err <= 1e-10, err
    ns.close(); lev.close()

    # 3. z slabs (the bench.py decomposition): x and y wrap inside the kernels, only z planes are exchanged, and the
    #    nodal smoother runs its two-phase fused sweep with a plane exchange after each phase
    boxes = split_boxes(n, (1, 1, 2))
    owners = [0, 1]
    lev = ix.Level(lib, g, boxes, owners)
    mine = [i for i, o in enumerate(owners) if o == rank]
    ns = ix.NavierStokes(lib, lev, "cpu", **kw)
    o2 = orc.OracleNS(n, **kw)
    ns.init_prob(100, pp); o2.init_prob(100, pp)
    d1, d2 = ns.post_init(), o2.post_init()
    assert abs(d1 - d2) <= 1e-13 * d2
    for _ in range(2):
        a, b = ns.step(), o2.step()
        assert abs(a - b) <= 1e-12 * b
    So = o2.get(0)
    err2 = 0.0
    for il, gi in enumerate(mine):
        lo, hi = boxes[gi]
        t = ns.field(0, il).numpy()
        err2 = max(err2, np.abs(t - So[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]).max())
    assert err2 <= 1e-10, err2
    o2.close()
    ns.close(); lev.close()

    # 4. the empty case: rank 1 owns NO boxes (both slabs on rank 0) -- it must still take part in every exchange,
    #    gather and reduction and report the same dt and solver iterations
    owners = [0, 0]
    lev = ix.Level(lib, g, boxes, owners)
    assert lev.num_local() == (2 if rank == 0 else 0)
    ns = ix.NavierStokes(lib, lev, "cpu", **kw)
    ns.init_prob(100, pp)
    dts = [ns.post_init()] + [ns.step() for _ in range(2)]
    o3 = orc.OracleNS(n, **kw)
    o3.init_prob(100, pp)
    dto = [o3.post_init()] + [o3.step() for _ in range(2)]
    assert np.allclose(dts, dto, rtol=1e-12, atol=0)
    if rank == 0:
        So = o3.get(0)
        for il, (lo, hi) in enumerate(boxes):
            t = ns.field(0, il).numpy()
            err2 = max(err2, np.abs(t - So[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]).max())
        assert err2 <= 1e-10, err2
    o3.close()
    # 5. walls across ranks: single-level RayleighTaylor (periodic x, y; slip walls in z; gravity, initialPressureProject) as two
    #    z slabs (each rank owns one wall) and as 2 x 1 x 2 blocks (deep-ghost nodal sweep in x, walls in z), vs the oracle
    nrt = (16, 16, 32)
    grt = ix.Geom.make(nrt, (0.0, 0.0, 0.0), (0.5, 0.5, 1.0), periodic=(1, 1, 0))
    kwrt = dict(visc_coef=1e-3, cfl=0.7, gravity=-1.0)
    ppr = [1.0, 2.0, 1.0, 0.0, 0.05, 0.02]
    ort = orc.OracleNS(nrt, (0, 0, 0), (0.5, 0.5, 1.0), per=(1, 1, 0), phys_lo=(0, 0, 4), phys_hi=(0, 0, 4), **kwrt)
    ort.init_prob(10, ppr)
    dto = [ort.post_init()] + [ort.step() for _ in range(2)]
    Srt = ort.get(0)
    err3 = 0.0
    for nbk, own in (((1, 1, 2), [0, 1]), ((2, 1, 2), [0, 1, 1, 0])):
        bxs = split_boxes(nrt, nbk)
        lv = ix.Level(lib, grt, bxs, own)
        nsr = ix.NavierStokes(lib, lv, "cpu", lo_bc=(0, 0, 4), hi_bc=(0, 0, 4), **kwrt)
        nsr.init_prob(10, ppr)
        dts = [nsr.post_init()] + [nsr.step() for _ in range(2)]
        assert np.allclose(dts, dto, rtol=1e-11, atol=0), (dts, dto)
        for il, gi in enumerate([i for i, o_ in enumerate(own) if o_ == rank]):
            lo, hi = bxs[gi]
            t = nsr.field(0, il).numpy()
            nz, ny, nx = hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1
            err3 = max(err3, np.abs(t[:, :nz, :ny, :nx] - Srt[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]).max())
        nsr.close(); lv.close()
    assert err3 <= 1e-10, err3
    ort.close()
    # 5b. outflow under gravity across ranks: Projection::set_outflow_bcs gathers the density next to the outflow face from both ranks
    #     (the hydrostatic integral runs down through both z slabs), vs the oracle
    nch = (16, 8, 16)
    gch = ix.Geom.make(nch, (0.0, 0.0, 0.0), (2.0, 1.0, 2.0), periodic=(0, 1, 0))
    kwch = dict(visc_coef=0.01, cfl=0.5, gravity=-0.5)
    bcv = [[0.0] * 5 for _ in range(6)]
    bcv[0] = [1.0, 0.0, 0.0, 1.0, 0.5]
    och = orc.OracleNS(nch, (0, 0, 0), (2.0, 1.0, 2.0), per=(0, 1, 0), phys_lo=(1, 0, 4), phys_hi=(2, 0, 4), bcv=bcv, **kwch)
    och.init_prob(101, [1.0, 1.0, 0.3])
    dto = [och.post_init()] + [och.step() for _ in range(2)]
    Sch = och.get(0)
    bxs = split_boxes(nch, (1, 1, 2))
    lv = ix.Level(lib, gch, bxs, [0, 1])
    nsc = ix.NavierStokes(lib, lv, "cpu", lo_bc=(1, 0, 4), hi_bc=(2, 0, 4), bc_vals=bcv, **kwch)
    nsc.init_prob(101, [1.0, 1.0, 0.3])
    dts = [nsc.post_init()] + [nsc.step() for _ in range(2)]
    assert np.allclose(dts, dto, rtol=1e-10, atol=0), (dts, dto)
    lo, hi = bxs[rank]
    t = nsc.field(0, 0).numpy()
    nz, ny, nx = hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1
    err4 = np.abs(t[:, :nz, :ny, :nx] - Sch[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]).max()
    assert err4 <= 1e-10, err4
    nsc.close(); lv.close(); och.close()
    # 6. reductions agree across ranks
    t = torch.tensor([max(err, err2, err3)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    print(f"rank {rank} ok max_err {t.item():.3e}", flush=True)
    ns.close(); lev.close()
    lib.iamrx_comm_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
