"""ctypes binding of the CPU oracle (oracle/oracle.h).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")


class OrcMG(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("max_iter", C.c_int), ("nu1", C.c_int), ("nu2", C.c_int),
                ("bottom_sweeps", C.c_int), ("max_coarsening", C.c_int), ("omega", C.c_double), ("iters", C.c_int),
                ("resnorm0", C.c_double), ("resnorm", C.c_double), ("rhsnorm", C.c_double),
                ("bottom_solver", C.c_int), ("bottom_maxiter", C.c_int), ("bottom_rtol", C.c_double), ("bottom_iters", C.c_int), ("pad_", C.c_int)]


class OrcNSParams(C.Structure):
    _fields_ = [("cfl", C.c_double), ("visc_coef", C.c_double), ("be_cn_theta", C.c_double), ("change_max", C.c_double),
                ("init_shrink", C.c_double), ("fixed_dt", C.c_double), ("gravity", C.c_double), ("visc_tol", C.c_double),
                ("mac_tol", C.c_double), ("mac_abs_tol", C.c_double), ("proj_tol", C.c_double), ("proj_abs_tol", C.c_double),
                ("init_iter", C.c_int), ("init_vel_iter", C.c_int), ("do_init_proj", C.c_int), ("use_forces_in_trans", C.c_int),
                ("conservative_tracer", C.c_int), ("verbose", C.c_int), ("scal_diff_coef", C.c_double), ("use_ppm", C.c_int), ("do_scalminmax", C.c_int), ("do_mom_diff", C.c_int), ("bottom_solver", C.c_int)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_ns_create.restype = C.c_void_p
        _lib.orc_ns_time.restype = C.c_double
        _lib.orc_ns_time.argtypes = [C.c_void_p]
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _i3(n):
    return (C.c_int * 3)(*[int(v) for v in n])


def _d3(x):
    return (C.c_double * 3)(*[float(v) for v in x])


def set_option(opt, value):
    lib().orc_set_option.argtypes = [C.c_int, C.c_double]
    assert lib().orc_set_option(int(opt), float(value)) == 0


def mg_default(**kw):
    m = OrcMG()
    lib().orc_mg_default(C.byref(m))
    for k, v in kw.items():
        setattr(m, k, v)
    return m


def _n_of(a):  # array [..., nz, ny, nx] -> (nx, ny, nz)
    return (a.shape[-1], a.shape[-2], a.shape[-3])


def abec_apply(dxinv, a, b, alpha, bx, by, bz, phi):
    ncomp, bn = phi.shape[0], bx.shape[0]
    out = np.empty_like(phi)
    lib().orc_abec_apply(_i3(_n_of(phi)), _d3(dxinv), C.c_double(a), C.c_double(b), _p(alpha), _p(bx), _p(by), _p(bz),
                         ncomp, bn, _p(phi), _p(out))
    return out


def abec_gsrb(dxinv, a, b, alpha, bx, by, bz, rhs, omega, redblack, phi):
    ncomp, bn = phi.shape[0], bx.shape[0]
    phi = phi.copy()
    lib().orc_abec_gsrb(_i3(_n_of(phi)), _d3(dxinv), C.c_double(a), C.c_double(b), _p(alpha), _p(bx), _p(by), _p(bz),
                        ncomp, bn, _p(rhs), C.c_double(omega), int(redblack), _p(phi))
    return phi


def tensor_cross(dxinv, b, ex, ey, ez, vel, out):
    out = out.copy()
    lib().orc_tensor_cross(_i3(_n_of(vel)), _d3(dxinv), C.c_double(b), _p(ex), _p(ey), _p(ez), _p(vel), _p(out))
    return out


def diffusion_solve(dx, tensor, a, b, alpha, ex, ey, ez, rhs, soln, mg=None):
    mg = mg or mg_default()
    soln = soln.copy()
    rc = lib().orc_diffusion_solve(_i3(_n_of(rhs)), _d3(dx), int(tensor), rhs.shape[0], C.c_double(a), C.c_double(b),
                                   _p(alpha), _p(ex), _p(ey), _p(ez), _p(rhs), _p(soln), C.byref(mg))
    return soln, rc, mg


def diffusion_apply(dx, tensor, a, b, alpha, ex, ey, ez, soln):
    out = np.empty_like(soln)
    lib().orc_diffusion_apply(_i3(_n_of(soln)), _d3(dx), int(tensor), soln.shape[0], C.c_double(a), C.c_double(b),
                              _p(alpha), _p(ex), _p(ey), _p(ez), _p(soln), _p(out))
    return out


def mac_project(dx, umac, vmac, wmac, rho, rhs, phi, rhs_scale, mg=None):
    mg = mg or mg_default()
    u, v, w, phi = umac.copy(), vmac.copy(), wmac.copy(), phi.copy()
    rc = lib().orc_mac_project(_i3(_n_of(rho)), _d3(dx), _p(u), _p(v), _p(w), _p(rho), _p(rhs), _p(phi),
                               C.c_double(rhs_scale), C.byref(mg))
    return u, v, w, phi, rc, mg


def nodal_divu(dxinv, vel):
    out = np.empty(vel.shape[1:], dtype=np.float64)
    lib().orc_nodal_divu(_i3(_n_of(vel)), _d3(dxinv), _p(vel), _p(out))
    return out


def nodal_adotx(dxinv, sigma, phi):
    out = np.empty_like(phi)
    lib().orc_nodal_adotx(_i3(_n_of(phi)), _d3(dxinv), _p(sigma), _p(phi), _p(out))
    return out


def nodal_gs(dxinv, sigma, rhs, color, phi):
    phi = phi.copy()
    lib().orc_nodal_gs(_i3(_n_of(phi)), _d3(dxinv), _p(sigma), _p(rhs), int(color), _p(phi))
    return phi


def nodal_mknewu(dxinv, sigma, phi, vel):
    vel = vel.copy()
    gp = np.empty_like(vel)
    lib().orc_nodal_mknewu(_i3(_n_of(phi)), _d3(dxinv), _p(sigma), _p(phi), _p(vel), _p(gp))
    return vel, gp


def nodal_project(dx, vel, sigma, phi, mg=None):
    mg = mg or mg_default()
    vel, phi = vel.copy(), phi.copy()
    gp = np.zeros_like(vel)
    rc = lib().orc_nodal_project(_i3(_n_of(sigma)), _d3(dx), _p(vel), _p(sigma), _p(phi), _p(gp), 0, C.byref(mg))
    return vel, phi, gp, rc, mg


def extrap_vel_to_faces(dx, dt, vel, force, fit=0, ppm=0):
    shp = vel.shape[1:]
    u, v, w = (np.empty(shp, dtype=np.float64) for _ in range(3))
    lib().orc_extrap_vel_to_faces(_i3(_n_of(vel)), _d3(dx), C.c_double(dt), _p(vel), _p(force), int(fit) | (2 if ppm else 0), _p(u), _p(v), _p(w))
    return u, v, w


def compute_aofs(dx, dt, S, force, umac, vmac, wmac, iconserv, fit=0, divu=None, want_fluxes=False, ppm=0):
    ncomp = S.shape[0]
    aofs = np.empty_like(S)
    ic = (C.c_int * ncomp)(*iconserv)
    outs = [np.empty_like(S) for _ in range(6)] if want_fluxes else [None] * 6
    lib().orc_compute_aofs(_i3(_n_of(S)), _d3(dx), C.c_double(dt), ncomp, _p(S), _p(force), _p(divu), _p(umac), _p(vmac),
                           _p(wmac), ic, int(fit) | (2 if ppm else 0), _p(aofs), *[_p(o) for o in outs])
    return (aofs, outs) if want_fluxes else aofs


def compute_aofs2(dx, dt, S, force, umac, vmac, wmac, iconserv, fit=0, divu=None, ppm=0, uflux=None, is_sync=False,
                  aofs_in=None, known_edges=None):
    """Full call-site argument list: returns (aofs, fluxes[3], edges[3])."""
    ncomp = S.shape[0]
    aofs = np.zeros_like(S) if aofs_in is None else aofs_in.copy()
    ic = (C.c_int * ncomp)(*iconserv)
    fl = [np.empty_like(S) for _ in range(3)]
    ed = [e.copy() for e in known_edges] if known_edges is not None else [np.empty_like(S) for _ in range(3)]
    uf = uflux if uflux is not None else (None, None, None)
    lib().orc_compute_aofs2(_i3(_n_of(S)), _d3(dx), C.c_double(dt), ncomp, _p(S), _p(force), _p(divu), _p(umac), _p(vmac), _p(wmac),
                            _p(uf[0]), _p(uf[1]), _p(uf[2]), ic, int(fit) | (2 if ppm else 0), int(is_sync),
                            int(known_edges is not None), _p(aofs), *[_p(o) for o in fl], *[_p(o) for o in ed])
    return aofs, fl, ed


# ---- non-periodic domains: "padded" arrays carry their ghost layers -------------------------------------------------
def _ia(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    return a.ctypes.data_as(C.POINTER(C.c_int)), a


def padded_shape(n, ncomp, ng):
    return (ncomp, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng)


def fill_physbc(n, per, ng, a, bclo, bchi, bcv=None):
    """a: padded (ncomp, nz+2ng, ...) array with valid data inside; returns it with periodic + physical ghost cells filled."""
    a = np.ascontiguousarray(a).copy()
    plo, _k1 = _ia(bclo); phi, _k2 = _ia(bchi)
    bv = None if bcv is None else np.ascontiguousarray(np.asarray(bcv, dtype=np.float64))
    lib().orc_fill_physbc(_i3(n), _i3(per), int(ng), a.shape[0], plo, phi, _p(bv), _p(a))
    return a


def extrap_vel_to_faces_bc(n, per, dx, dt, vel, force, bclo, bchi, fit=0, ppm=0):
    outs = [np.zeros(padded_shape(n, 1, 1)) for _ in range(3)]
    plo, _k1 = _ia(bclo); phi, _k2 = _ia(bchi)
    lib().orc_extrap_vel_to_faces_bc(_i3(n), _i3(per), _d3(dx), C.c_double(dt), _p(vel), _p(force), int(fit) | (2 if ppm else 0), plo, phi,
                                     *[_p(o) for o in outs])
    return outs


def compute_aofs_bc(n, per, dx, dt, S, force, umac, vmac, wmac, iconserv, bclo, bchi, fit=0, ppm=0, is_velocity=0, divu=None):
    ncomp = S.shape[0]
    aofs = np.zeros((ncomp, n[2], n[1], n[0]))
    fl = [np.zeros(padded_shape(n, ncomp, 1)) for _ in range(3)]
    ed = [np.zeros(padded_shape(n, ncomp, 1)) for _ in range(3)]
    ic = (C.c_int * ncomp)(*iconserv)
    plo, _k1 = _ia(bclo); phi, _k2 = _ia(bchi)
    lib().orc_compute_aofs_bc(_i3(n), _i3(per), _d3(dx), C.c_double(dt), ncomp, _p(S), _p(force), _p(divu), _p(umac), _p(vmac), _p(wmac), ic,
                              int(fit) | (2 if ppm else 0) | (4 if is_velocity else 0), plo, phi, _p(aofs), *[_p(o) for o in fl], *[_p(o) for o in ed])
    return aofs, fl, ed


def mac_project_bc(n, per, dx, umac, vmac, wmac, rho, rhs, phi, rhs_scale, lobc, hibc, maxorder=4, mg=None):
    mg = mg or mg_default()
    u, v, w, phi = umac.copy(), vmac.copy(), wmac.copy(), phi.copy()
    rc = lib().orc_mac_project_bc(_i3(n), _i3(per), _d3(dx), _p(u), _p(v), _p(w), _p(rho), _p(rhs), _p(phi), C.c_double(rhs_scale),
                                  _i3(lobc), _i3(hibc), int(maxorder), C.byref(mg))
    return u, v, w, phi, rc, mg


def nodal_project_bc(n, per, dx, vel, sigma, phi, lobc, hibc, mg=None):
    mg = mg or mg_default()
    vel, phi = vel.copy(), phi.copy()
    gp = np.zeros((3, n[2], n[1], n[0]))
    rc = lib().orc_nodal_project_bc(_i3(n), _i3(per), _d3(dx), _p(vel), _p(sigma), _p(phi), _p(gp), _i3(lobc), _i3(hibc), C.byref(mg))
    return vel, phi, gp, rc, mg


def diffusion_bc(n, per, dx, solve, tensor, a, b, alpha, ex, ey, ez, rhs, soln, lobc, hibc, maxorder=2, mg=None):
    """lobc/hibc: [ncomp][3] LinOpBCType.  solve: returns (soln padded, rc, mg); apply: returns out (dense)."""
    mg = mg or mg_default()
    ncomp = soln.shape[0]
    soln = soln.copy()
    out = np.zeros((ncomp, n[2], n[1], n[0]))
    plo, _k1 = _ia(lobc); phi_, _k2 = _ia(hibc)
    rc = lib().orc_diffusion_bc(_i3(n), _i3(per), _d3(dx), int(solve), int(tensor), ncomp, C.c_double(a), C.c_double(b), _p(alpha),
                                _p(ex), _p(ey), _p(ez), _p(rhs), _p(soln), _p(out), plo, phi_, int(maxorder), C.byref(mg))
    return (soln, rc, mg) if solve else out


def average_down(nc, ixtype, fine):
    ncomp = fine.shape[0]
    crse = np.empty((ncomp, nc[2], nc[1], nc[0]))
    lib().orc_average_down(_i3(nc), ncomp, int(ixtype), _p(fine), _p(crse))
    return crse


def interp(kind, nc, crse):
    ncomp = crse.shape[0]
    fine = np.empty((ncomp, 2 * nc[2], 2 * nc[1], 2 * nc[0]))
    lib().orc_interp(int(kind), _i3(nc), ncomp, _p(crse), _p(fine))
    return fine


def interp_bndry(nc, per, crse, covered, flo, fhi, out):
    """InterpBndryData (order 3) into the face ghost cells of the padded fine box array `out` (ncomp, fn + 2); returns it."""
    out = np.ascontiguousarray(out.copy())
    m = np.ascontiguousarray(covered.astype(np.uint8))
    lib().orc_interp_bndry(_i3(nc), _i3(per), int(crse.shape[0]), _p(np.ascontiguousarray(crse)), m.ctypes.data_as(C.c_void_p), _i3(flo), _i3(fhi),
                           _p(out))
    return out


def fluxreg(nc, mask, cflux, fflux, dt, vol):
    ncomp = cflux[0].shape[0]
    reg = np.empty((ncomp, nc[2], nc[1], nc[0]))
    m = np.ascontiguousarray(mask.astype(np.uint8))
    lib().orc_fluxreg(_i3(nc), ncomp, m.ctypes.data_as(C.c_void_p), _p(cflux[0]), _p(cflux[1]), _p(cflux[2]), _p(fflux[0]), _p(fflux[1]),
                      _p(fflux[2]), C.c_double(dt), C.c_double(vol), _p(reg))
    return reg


class OracleNS:
    def __init__(self, n, prob_lo=(0, 0, 0), prob_hi=(1, 1, 1), per=None, phys_lo=None, phys_hi=None, bcv=None, **params):
        self.n = tuple(n)
        p = OrcNSParams()
        lib().orc_ns_params_default(C.byref(p))
        for k, v in params.items():
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, v)
        self.h = C.c_void_p(lib().orc_ns_create(_i3(n), _d3(prob_lo), _d3(prob_hi), C.byref(p)))
        if per is not None:
            bv = None if bcv is None else np.ascontiguousarray(np.asarray(bcv, dtype=np.float64).reshape(6, 5))
            lib().orc_ns_set_bc(self.h, _i3(per), _i3(phys_lo), _i3(phys_hi), _p(bv))

    def get_padded(self, which):
        nc, ng = {0: (5, 1), 1: (1, 2), 2: (3, 1), 4: (1, 1), 5: (1, 1), 6: (1, 1)}[which]
        out = np.empty(padded_shape(self.n, nc, ng), dtype=np.float64)
        lib().orc_ns_get_padded(self.h, which, _p(out))
        return out

    def set_turbulent_forcing(self, nmodes, mode_start, div_free_force, forcedata):
        fd = np.ascontiguousarray(forcedata, dtype=np.float64)
        lib().orc_ns_set_turbulent_forcing(self.h, nmodes, mode_start, int(div_free_force), fd.shape[1], _p(fd))

    def init_prob(self, probtype, params):
        arr = (C.c_double * len(params))(*params)
        lib().orc_ns_init_prob(self.h, probtype, arr, len(params))

    def post_init(self):
        dt = C.c_double(0)
        rc = lib().orc_ns_post_init(self.h, C.byref(dt))
        assert rc == 0, rc
        return dt.value

    def step(self, dt=-1.0):
        d = C.c_double(dt)
        rc = lib().orc_ns_step(self.h, C.byref(d))
        assert rc == 0, rc
        return d.value

    def get(self, which):
        nc = {0: 5, 1: 1, 2: 3, 4: 1, 5: 1, 6: 1, 7: 5}[which]
        out = np.empty((nc, self.n[2], self.n[1], self.n[0]), dtype=np.float64)
        lib().orc_ns_get(self.h, which, _p(out))
        return out

    def set_state(self, s):
        lib().orc_ns_set_state(self.h, _p(np.ascontiguousarray(s)))

    @property
    def time(self):
        return lib().orc_ns_time(self.h)

    def last_iters(self):
        it = (C.c_int * 3)()
        lib().orc_ns_last_iters(self.h, it)
        return tuple(it)

    def close(self):
        if self.h:
            lib().orc_ns_destroy(self.h)
            self.h = None
