"""Deterministic synthetic inputs of SURVEY.md 8d for the kernel micro-benchmarks (K1 GSRB, K2 advection), generated on the
device with torch: hash = splitmix64(seed xor linear index) -> uniform(-1, 1), identical to tests/util.hash_uniform (numpy)."""
import math

import torch

_M64 = (1 << 64) - 1


def _lsr(z, k):
    """logical shift right of int64 tensors (torch's >> is arithmetic)"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def _wrap(c):
    """python int in [0, 2^64) -> the int64 with the same bits"""
    return c - (1 << 64) if c >= (1 << 63) else c


def hash_uniform(seed, shape, device="cpu"):
    n = 1
    for m in shape:
        n *= m
    z = (torch.arange(n, dtype=torch.int64, device=device) ^ seed) + _wrap(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _wrap(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _wrap(0x94D049BB133111EB)
    z = z ^ _lsr(z, 31)
    u = _lsr(z, 11).to(torch.float64) / float(1 << 53)
    return (2.0 * u - 1.0).reshape(shape)


def centres(n, lo=0.0, hi=1.0, device="cpu"):
    return lo + (torch.arange(n, dtype=torch.float64, device=device) + 0.5) * ((hi - lo) / n)


def k1_gsrb(n, device="cpu", nu=1.0e-4, theta=0.5):
    """K1: phi0 = hash(1), rhs = hash(2), rho = 1 + 1/2 sin 2pi x sin 2pi y sin 2pi z on [0,1]^3, beta_face = (dt/2)/mean(rho_L, rho_R)
    with dt = 0.7/n (the MAC projection's coefficients), periodic.  Returns dense periodic arrays [k][j][i] (faces: the LOW face of
    each cell) plus the (a, b) pairs of the two variants: MAC (a = 0, b = 1) and diffusion (a = 1 with alpha = rho, b = theta dt nu)."""
    x = centres(n, device=device)
    s = torch.sin(2 * math.pi * x)
    rho = 1.0 + 0.5 * s.view(n, 1, 1) * s.view(1, n, 1) * s.view(1, 1, n)          # [k][j][i]
    dt = 0.7 / n
    beta = [(0.5 * dt) / (0.5 * (rho + torch.roll(rho, 1, dims=2 - d))) for d in range(3)]   # d = 0: x faces (roll along i)
    return {"phi": hash_uniform(1, (n, n, n), device), "rhs": hash_uniform(2, (n, n, n), device), "rho": rho, "beta": beta, "dt": dt,
            "mac": (0.0, 1.0), "diffusion": (1.0, theta * dt * nu)}


def k2_advection(n, device="cpu"):
    """K2: q = the Taylor-Green velocity of prob_init.cpp:538-540 (a = b = c = 1, V0 = 1) at cell centres, u_mac = the same field at
    face centres, force = 0, divu = 0, dt = 0.7 dx; scalars for the second call: rho = 1, tracer = the TG pressure-like field."""
    tp = 2 * math.pi
    xc = centres(n, device=device)
    xf = torch.arange(n, dtype=torch.float64, device=device) / n           # low faces

    def tg(x, y, z):
        X, Y, Z = x.view(1, 1, -1), y.view(1, -1, 1), z.view(-1, 1, 1)
        u = torch.sin(tp * X) * torch.cos(tp * Y) * torch.cos(tp * Z)
        v = -torch.cos(tp * X) * torch.sin(tp * Y) * torch.cos(tp * Z)
        return u, v, torch.zeros_like(u)

    u, v, w = tg(xc, xc, xc)
    umac = tg(xf, xc, xc)[0]
    vmac = tg(xc, xf, xc)[1]
    wmac = tg(xc, xc, xf)[2]
    return {"vel": torch.stack([u, v, w]), "umac": umac, "vmac": vmac, "wmac": wmac, "dt": 0.7 / n}


def ghosted(dense, ng, ext=(0, 0, 0)):
    """dense periodic array [(c,) k, j, i] -> the array of one box spanning the domain with ng ghost layers (periodic images) and
    `ext` extra high layers per direction (x, y, z) for face / nodal data (the high duplicate of the periodic low face / node)."""
    d4 = dense if dense.dim() == 4 else dense.unsqueeze(0)
    nz, ny, nx = d4.shape[1:]
    ix = torch.arange(-ng, nx + ext[0] + ng, device=d4.device) % nx
    iy = torch.arange(-ng, ny + ext[1] + ng, device=d4.device) % ny
    iz = torch.arange(-ng, nz + ext[2] + ng, device=d4.device) % nz
    return d4[:, iz][:, :, iy][:, :, :, ix].contiguous()
