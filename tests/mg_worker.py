"""Worker of tests/test_multigpu.py: one rank of a 2-GPU NCCL job (torchrun-style env).  z-slab decomposition as in
bench.py; two steps of a variable-density problem compared with the single-box CPU oracle."""
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
from util import split_boxes  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    lib = ix.load()
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = C.create_string_buffer(128)
        lib.check(lib.iamrx_comm_unique_id(buf))
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.to(dev)
    dist.broadcast(uid, 0)
    lib.check(lib.iamrx_comm_init(rank, world, bytes(uid.cpu().numpy().tobytes())))
    n = (32, 32, 32)   # 32 x 32 x 16 slabs: one distributed multigrid level, then the consolidated replicated ones
                       # (kept small: both ranks also run the CPU oracle; verified at 64^3 as well, profiles/r01_notes.md)
    boxes = split_boxes(n, (1, 1, world))
    owners = list(range(world))
    lev = ix.Level(lib, ix.Geom.make(n), boxes, owners)
    kw = dict(visc_coef=1e-3, cfl=0.7, gravity=-0.5)
    ns = ix.NavierStokes(lib, lev, dev, **kw)
    pp = [1.0, 1.0, 1.0, 1.0, 1.0]
    ns.init_prob(100, pp)
    dts = [ns.post_init()] + [ns.step() for _ in range(2)]
    torch.cuda.synchronize()
    err = torch.zeros(1, dtype=torch.float64, device=dev)
    if True:
        o = orc.OracleNS(n, **kw)
        o.init_prob(100, pp)
        dto = [o.post_init()] + [o.step() for _ in range(2)]
        assert np.allclose(dts, dto, rtol=1e-11, atol=0), (dts, dto)
        So = o.get(0)
        lo, hi = boxes[rank]
        t = ns.field(0, 0).cpu().numpy()
        err[0] = float(np.abs(t - So[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]).max())
        o.close()
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    assert err.item() <= 1e-10, err.item()
    print(f"rank {rank} ok max_err {err.item():.3e} iters {ns.last_iters()}", flush=True)
    ns.close(); lev.close()
    lib.iamrx_comm_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
