"""Test helpers: deterministic synthetic fields (splitmix64 hash -> uniform(-1,1)),
and conversion between dense periodic arrays [c][k][j][i] and ghosted per-box fabs."""
import ctypes as C

import numpy as np
import torch

import iamr_b200 as ix


def hash_uniform(seed, shape):
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        z = (np.arange(n, dtype=np.uint64) ^ np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    return (2.0 * u - 1.0).reshape(shape)


def smooth_field(n, seed, ncomp=1, amp=1.0):
    """Smooth periodic field: a few low Fourier modes with hashed coefficients."""
    nx, ny, nz = n
    x = (np.arange(nx) + 0.5) / nx
    y = (np.arange(ny) + 0.5) / ny
    z = (np.arange(nz) + 0.5) / nz
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    out = np.zeros((ncomp, nz, ny, nx))
    co = hash_uniform(seed, (ncomp, 6, 4))
    for c in range(ncomp):
        for m in range(6):
            kx, ky, kz = (m % 3) + 1, ((m // 2) % 3), (m % 2) + (m // 4)
            out[c] += amp * co[c, m, 0] * np.sin(2 * np.pi * (kx * X + co[c, m, 1])) * np.cos(2 * np.pi * (ky * Y + co[c, m, 2])) \
                * np.cos(2 * np.pi * (kz * Z + co[c, m, 3]))
    return out


def split_boxes(n, nb):
    """Decompose the domain (nx,ny,nz) into nb[0] x nb[1] x nb[2] boxes -> list of (lo, hi)."""
    boxes = []
    for kb in range(nb[2]):
        for jb in range(nb[1]):
            for ib in range(nb[0]):
                s = [n[d] // nb[d] for d in range(3)]
                lo = (ib * s[0], jb * s[1], kb * s[2])
                hi = (lo[0] + s[0] - 1, lo[1] + s[1] - 1, lo[2] + s[2] - 1)
                boxes.append((lo, hi))
    return boxes


IX_EXT = {ix.CELL: (0, 0, 0), ix.XFACE: (1, 0, 0), ix.YFACE: (0, 1, 0), ix.ZFACE: (0, 0, 1), ix.NODE: (1, 1, 1)}


def to_fab(dense, box, ng, ixtype, device, fill_ghost=True):
    """dense: periodic array (ncomp, nz, ny, nx) of the whole domain.  Returns (tensor, Fab) for
    one box with ng ghost layers; ghost and face/node overlap values come from periodic wrap."""
    ncomp, nz, ny, nx = dense.shape
    lo, hi = box
    ext = IX_EXT[ixtype]
    idx = []
    for d, nd in zip(range(3), (nx, ny, nz)):
        r = np.arange(lo[d] - ng, hi[d] + ext[d] + ng + 1)
        idx.append(r % nd)
    sub = dense[:, idx[2]][:, :, idx[1]][:, :, :, idx[0]]
    if not fill_ghost and ng > 0:
        sub = sub.copy()
        mask = np.ones(sub.shape[1:], dtype=bool)
        mask[ng:-ng, ng:-ng, ng:-ng] = False
        sub[:, mask] = 1.0e40
    t = torch.from_numpy(np.ascontiguousarray(sub)).to(device)
    return t, ix.fab_of(t, [lo[d] - ng for d in range(3)])


def from_fabs(tensors, boxes, ng, ixtype, n, ncomp):
    """Assemble per-box tensors back into the dense periodic array (valid region only; shared
    face/node points are checked to agree and taken from the first box)."""
    nx, ny, nz = n
    out = np.full((ncomp, nz, ny, nx), np.nan)
    ext = IX_EXT[ixtype]
    maxdiff = 0.0
    for t, (lo, hi) in zip(tensors, boxes):
        a = t.detach().cpu().numpy()
        sl = tuple(slice(ng, a.shape[1 + q] - ng) for q in range(3))
        v = a[(slice(None),) + sl]
        iz = np.arange(lo[2], hi[2] + ext[2] + 1) % nz
        iy = np.arange(lo[1], hi[1] + ext[1] + 1) % ny
        ixx = np.arange(lo[0], hi[0] + ext[0] + 1) % nx
        cur = out[:, iz][:, :, iy][:, :, :, ixx]
        have = ~np.isnan(cur)
        if have.any():
            maxdiff = max(maxdiff, float(np.abs(cur[have] - v[have]).max()))
        new = np.where(have, cur, v)
        out[np.ix_(np.arange(ncomp), iz, iy, ixx)] = new
    assert not np.isnan(out).any()
    return out, maxdiff


def fab_array(fabs):
    return (ix.Fab * len(fabs))(*fabs)


def d3(x):
    return (C.c_double * 3)(*[float(v) for v in x])


def i3(x):
    return (C.c_int * 3)(*[int(v) for v in x])


def box_of(lo, hi):
    return ix.Box.make(lo, hi)


def stream_of(device):
    if torch.device(device).type == "cuda":
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    return C.c_void_p(0)


def sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)
