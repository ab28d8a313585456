"""ctypes binding of include/iamrx.h.  Thin by design: structs, argtypes, error
translation, and helpers that describe torch tensors as ``iamrx_fab`` views."""
import ctypes as C
import os

import numpy as np
import torch

CELL, XFACE, YFACE, ZFACE, NODE = 0, 1, 2, 3, 4

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "libiamrx.so")


class IamrxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"iamrx error {code}: {msg}")
        self.code = code


class Box(C.Structure):
    _fields_ = [("lo", C.c_int * 3), ("hi", C.c_int * 3)]

    @staticmethod
    def make(lo, hi):
        b = Box()
        for d in range(3):
            b.lo[d] = int(lo[d])
            b.hi[d] = int(hi[d])
        return b

    def shape(self):
        return tuple(self.hi[d] - self.lo[d] + 1 for d in range(3))


class Fab(C.Structure):
    _fields_ = [("p", C.c_void_p), ("lo", C.c_int * 3), ("hi", C.c_int * 3),
                ("jstride", C.c_int64), ("kstride", C.c_int64), ("nstride", C.c_int64),
                ("ncomp", C.c_int), ("pad_", C.c_int)]


class Geom(C.Structure):
    _fields_ = [("domain", Box), ("dx", C.c_double * 3), ("prob_lo", C.c_double * 3),
                ("periodic", C.c_int * 3), ("pad_", C.c_int)]

    @staticmethod
    def make(ncell, prob_lo=(0.0, 0.0, 0.0), prob_hi=(1.0, 1.0, 1.0), periodic=(1, 1, 1)):
        g = Geom()
        g.domain = Box.make((0, 0, 0), tuple(n - 1 for n in ncell))
        for d in range(3):
            g.dx[d] = (prob_hi[d] - prob_lo[d]) / ncell[d]
            g.prob_lo[d] = prob_lo[d]
            g.periodic[d] = int(periodic[d])
        return g


class BCRec(C.Structure):
    """amrex::BCRec of one component: math BC codes (lo[3], hi[3])."""
    _fields_ = [("lo", C.c_int * 3), ("hi", C.c_int * 3)]

    @staticmethod
    def make(lo, hi):
        b = BCRec()
        for d in range(3):
            b.lo[d] = int(lo[d])
            b.hi[d] = int(hi[d])
        return b


# amrex::BCType math codes (include/iamrx.h IAMRX_BC_*)
BC_INT_DIR, BC_REFLECT_ODD, BC_REFLECT_EVEN, BC_FOEXTRAP, BC_EXT_DIR, BC_HOEXTRAP = 0, -1, 1, 2, 3, 4
SYNC_PC, SYNC_CELL_CONS = 0, 1   # NavierStokesBase::SyncInterpType subset (iamrx_sync_interp)
# iamrx_compute_aofs_box flags
ADV_PPM, ADV_FORCES_IN_TRANS, ADV_IS_VELOCITY, ADV_WRITE_FLUXES, ADV_IS_SYNC, ADV_STAGED, ADV_KNOWN_EDGE_STATE = 1, 2, 4, 8, 16, 32, 64


class MGInfo(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("max_iter", C.c_int),
                ("max_coarsening", C.c_int), ("nu1", C.c_int), ("nu2", C.c_int),
                ("bottom_sweeps", C.c_int), ("verbose", C.c_int), ("omega", C.c_double),
                ("iters", C.c_int), ("maxorder", C.c_int), ("resnorm0", C.c_double),
                ("resnorm", C.c_double), ("rhsnorm", C.c_double),
                ("bottom_solver", C.c_int), ("bottom_maxiter", C.c_int), ("bottom_rtol", C.c_double),
                ("bottom_iters", C.c_int), ("pad_", C.c_int)]


class LinopBC(C.Structure):
    """iamrx_linop_bc: LinOpBCType [comp][dir] of the low / high sides + maxorder."""
    _fields_ = [("lo", (C.c_int * 3) * 3), ("hi", (C.c_int * 3) * 3), ("maxorder", C.c_int), ("pad_", C.c_int)]

    @staticmethod
    def make(lo, hi, maxorder=2):
        b = LinopBC()
        for c in range(3):
            cs = c if c < len(lo) else 0
            for d in range(3):
                b.lo[c][d] = int(lo[cs][d])
                b.hi[c][d] = int(hi[cs][d])
        b.maxorder = maxorder
        return b


OPT_SMALL_VEL, OPT_SLOPE_ORDER, OPT_CORNER_FORM, OPT_EXTDIR_BOTH = 0, 1, 2, 3
LINOP_PERIODIC, LINOP_DIRICHLET, LINOP_NEUMANN, LINOP_REFLECT_ODD, LINOP_INFLOW = 0, 1, 2, 3, 4


class NSParams(C.Structure):
    _fields_ = [("cfl", C.c_double), ("visc_coef", C.c_double), ("scal_diff_coef", C.c_double),
                ("be_cn_theta", C.c_double), ("change_max", C.c_double), ("init_shrink", C.c_double),
                ("fixed_dt", C.c_double), ("gravity", C.c_double), ("visc_tol", C.c_double),
                ("mac_tol", C.c_double), ("mac_abs_tol", C.c_double), ("proj_tol", C.c_double),
                ("proj_abs_tol", C.c_double), ("init_iter", C.c_int), ("init_vel_iter", C.c_int),
                ("do_init_proj", C.c_int), ("use_forces_in_trans", C.c_int), ("verbose", C.c_int),
                ("conservative_tracer", C.c_int), ("mg_verbose", C.c_int), ("godunov_ppm", C.c_int), ("do_scalminmax", C.c_int), ("do_mom_diff", C.c_int), ("bottom_solver", C.c_int),
                ("lo_bc", C.c_int * 3), ("hi_bc", C.c_int * 3), ("bc_vals", (C.c_double * 5) * 6)]


_P = C.POINTER
_vp = C.c_void_p
_d3 = C.c_double * 3
_i3 = C.c_int * 3

# name -> (restype, argtypes); every symbol include/iamrx.h declares
SIGNATURES = {
    "iamrx_last_error": (C.c_char_p, []),
    "iamrx_set_option": (C.c_int, [C.c_int, C.c_double]),
    "iamrx_get_option": (C.c_double, [C.c_int]),
    "iamrx_version": (C.c_int, []),
    "iamrx_launch_count": (C.c_int64, []),
    "iamrx_launch_count_reset": (None, []),
    "iamrx_device_ok": (C.c_int, []),
    "iamrx_debug_fp64_peak": (C.c_int, [_P(C.c_double), _vp]),
    "iamrx_prof_enable": (C.c_int, [C.c_int, C.c_int64]),
    "iamrx_prof_reset": (None, []),
    "iamrx_prof_all": (C.c_int, [C.c_int]),
    "iamrx_prof_dump": (C.c_int, [C.c_char_p, C.c_int]),
    "iamrx_prof_report": (C.c_int, [C.c_int, _P(C.c_double), _P(C.c_int64), _P(C.c_double)]),
    "iamrx_abec_gsrb_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), C.c_double, C.c_double, _P(Fab), _P(Fab), _P(Fab),
                                      _P(Fab), _P(C.c_double), C.c_double, C.c_int, C.c_int, _vp]),
    "iamrx_abec_gsrb_sweep_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), _P(Fab), C.c_double, C.c_double, _P(Fab), _P(Fab),
                                            _P(Fab), _P(Fab), _P(C.c_double), C.c_double, C.c_int, _vp]),
    "iamrx_abec_apply_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), _P(Fab), C.c_double, C.c_double, _P(Fab), _P(Fab),
                                       _P(Fab), _P(Fab), _P(C.c_double), C.c_int, _vp]),
    "iamrx_tensor_cross_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), C.c_double,
                                         _P(C.c_double), _vp]),
    "iamrx_extrap_vel_to_faces_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(BCRec), _P(Geom),
                                                C.c_double, C.c_int, _vp]),
    "iamrx_compute_aofs_box": (C.c_int, [_P(Box), _P(Fab), C.c_int, _P(Fab), C.c_int, C.c_int, _P(Fab), C.c_int,
                                         _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab),
                                         _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(C.c_int), _P(BCRec),
                                         _P(Geom), C.c_double, C.c_int, _vp]),
    "iamrx_nodal_divu_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), _P(C.c_double), _vp]),
    "iamrx_nodal_adotx_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(C.c_double), _vp]),
    "iamrx_nodal_gs_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), _P(Fab), _P(C.c_double), C.c_int, _vp]),
    "iamrx_nodal_gs_sweep_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(C.c_double), _vp]),
    "iamrx_nodal_mknewu_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(C.c_double), _vp]),
    "iamrx_comm_unique_id": (C.c_int, [C.c_char_p]),
    "iamrx_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_char_p]),
    "iamrx_comm_set_transport": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp]),
    "iamrx_comm_finalize": (C.c_int, []),
    "iamrx_debug_fb_plan": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _P(C.c_int), _P(C.c_int), _P(C.c_int), _P(C.c_int)]),
    "iamrx_comm_rank": (C.c_int, []),
    "iamrx_comm_size": (C.c_int, []),
    "iamrx_allreduce": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "iamrx_level_create": (C.c_int, [_P(Geom), C.c_int, _P(Box), _P(C.c_int), _P(_vp)]),
    "iamrx_level_destroy": (C.c_int, [_vp]),
    "iamrx_level_num_local": (C.c_int, [_vp]),
    "iamrx_level_local_box": (C.c_int, [_vp, C.c_int, _P(Box), _P(C.c_int)]),
    "iamrx_fill_boundary": (C.c_int, [_vp, _P(Fab), C.c_int, C.c_int, C.c_int, _vp]),
    "iamrx_fill_physbc": (C.c_int, [_vp, _P(Fab), C.c_int, C.c_int, _P(BCRec), _P(C.c_double), _vp]),
    "iamrx_average_down_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), C.c_int, C.c_int, _vp]),
    "iamrx_interp_box": (C.c_int, [C.c_int, _P(Box), _P(Fab), _P(Fab), C.c_int, _vp]),
    "iamrx_fluxreg_create": (C.c_int, [_vp, _vp, C.c_int, _P(_vp)]),
    "iamrx_fluxreg_destroy": (C.c_int, [_vp]),
    "iamrx_fluxreg_num_patches": (C.c_int, [_vp]),
    "iamrx_fluxreg_reset": (C.c_int, [_vp, _vp]),
    "iamrx_fluxreg_crse_add": (C.c_int, [_vp, _P(Fab), _P(Fab), _P(Fab), C.c_double, C.c_double, _vp]),
    "iamrx_fluxreg_fine_add": (C.c_int, [_vp, _P(Fab), _P(Fab), _P(Fab), C.c_double, C.c_double, _vp]),
    "iamrx_fluxreg_reflux": (C.c_int, [_vp, _P(Fab), C.c_int, C.c_double, _vp]),
    "iamrx_fluxreg_field": (C.c_int, [_vp, C.c_int, _P(Fab)]),
    "iamrx_mg_info_default": (None, [_P(MGInfo)]),
    "iamrx_mac_project": (C.c_int, [_vp, _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), C.c_double,
                                    _P(C.c_int), _P(C.c_int), _P(MGInfo), _vp]),
    "iamrx_mac_get_fluxes": (C.c_int, [_vp, _P(Fab), _P(Fab), _P(Fab), _P(Fab), _vp]),
    "iamrx_nodal_project": (C.c_int, [_vp, _P(Fab), _P(Fab), _P(Fab), _P(Fab), C.c_int, _P(C.c_int), _P(C.c_int),
                                      _P(MGInfo), _vp]),
    "iamrx_diffusion_apply": (C.c_int, [_vp, C.c_int, C.c_int, _P(Fab), _P(Fab), C.c_double, C.c_double, _P(Fab),
                                        _P(Fab), _P(Fab), _P(Fab), _P(LinopBC), _vp]),
    "iamrx_diffusion_solve": (C.c_int, [_vp, C.c_int, C.c_int, _P(Fab), _P(Fab), C.c_double, C.c_double, _P(Fab),
                                        _P(Fab), _P(Fab), _P(Fab), _P(LinopBC), _P(MGInfo), _vp]),
    "iamrx_ns_params_default": (None, [_P(NSParams)]),
    "iamrx_ns_create": (C.c_int, [_vp, _P(NSParams), _P(_vp)]),
    "iamrx_ns_destroy": (C.c_int, [_vp]),
    "iamrx_debug_fb_stats": (None, [_P(C.c_int64), C.c_int]),
    "iamrx_ns_mac_sync_compute": (C.c_int, [_vp, _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), C.c_double]),
    "iamrx_create_umac_grown": (C.c_int, [_vp, _vp, _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _vp]),
    "iamrx_average_down": (C.c_int, [_vp, _vp, _P(Fab), _P(Fab), C.c_int, C.c_int, C.c_int, _vp]),
    "iamrx_mac_sync_solve": (C.c_int, [_vp, _vp, _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), _P(Fab), C.c_double, _P(C.c_int), _P(C.c_int),
                                       _P(MGInfo), _vp]),
    "iamrx_sync_interp": (C.c_int, [_vp, _vp, _P(Fab), C.c_int, _P(Fab), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _P(BCRec), _vp]),
    "iamrx_sync_proj_interp": (C.c_int, [_vp, _vp, _P(Fab), _P(Fab), _P(Fab), _vp]),
    "iamrx_set_coarse_fine_bc": (C.c_int, [_vp, _vp, _P(Fab), _P(Fab), C.c_int, _vp]),
    "iamrx_syncreg_create": (C.c_int, [_vp, _vp, C.c_double, _P(_vp)]),
    "iamrx_syncreg_destroy": (C.c_int, [_vp]),
    "iamrx_syncreg_crse_init": (C.c_int, [_vp, _P(Fab), C.c_double, _vp]),
    "iamrx_syncreg_fine_add": (C.c_int, [_vp, _P(Fab), C.c_double, _vp]),
    "iamrx_syncreg_init_rhs": (C.c_int, [_vp, _P(Fab), _P(C.c_int), _P(C.c_int), _vp]),
    "iamrx_syncreg_field": (C.c_int, [_vp, C.c_int, C.c_int, _P(Fab)]),
    "iamrx_fill_coarse_patch_nodal": (C.c_int, [_vp, _vp, _P(Fab), _P(Fab), _P(Fab), C.c_double, C.c_double, C.c_double, _vp]),
    "iamrx_diffusion_get_fluxes": (C.c_int, [_vp, C.c_int, _P(Fab), _P(Fab), _P(Fab), _P(Fab), C.c_double, _P(Fab), _P(Fab), _P(Fab), C.c_double, _vp]),
    "iamrx_fillpatch_two_levels": (C.c_int, [_vp, _vp, _P(Fab), _P(Fab), _P(Fab), C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                             _P(BCRec), _P(C.c_double), _vp]),
    "iamrx_ns_init_prob": (C.c_int, [_vp, C.c_int, _P(C.c_double), C.c_int]),
    "iamrx_ns_set_turbulent_forcing": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _P(C.c_double)]),
    "iamrx_turbulent_force_box": (C.c_int, [_P(Box), _P(Fab), _P(Fab), _P(Geom), C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                            _P(C.c_double), _vp]),
    "iamrx_ns_post_init": (C.c_int, [_vp, _P(C.c_double)]),
    "iamrx_ns_step": (C.c_int, [_vp, _P(C.c_double)]),
    "iamrx_ns_time": (C.c_double, [_vp]),
    "iamrx_ns_nstep": (C.c_int, [_vp]),
    "iamrx_ns_field": (C.c_int, [_vp, C.c_int, C.c_int, _P(Fab)]),
    "iamrx_ns_step_host": (C.c_int, [_vp, _P(_vp), _P(_vp), _P(C.c_double)]),
    "iamrx_ns_write_plotfile": (C.c_int, [_vp, C.c_char_p]),
    "iamrx_ns_last_iters": (C.c_int, [_vp, _P(C.c_int)]),
    "iamrx_ns_sum_integrated_quantities": (C.c_int, [_vp, _P(C.c_double)]),
}


class Library:
    """A loaded libiamrx (or, in the CPU tests only, the host emulation build of the
    same sources)."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it with `make` (nvcc, sm_100a). iamr_b200 has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path, mode=C.RTLD_LOCAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.dll, name)  # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args

    def __getattr__(self, name):
        return getattr(self.dll, name)

    def check(self, rc, allow_positive=False):
        if rc < 0 or (rc > 0 and not allow_positive):
            raise IamrxError(rc, self.dll.iamrx_last_error().decode())
        return rc


_default = None


def load(path=None):
    """Load the product library (default) or an explicit path."""
    global _default
    if path is None:
        if _default is None:
            _default = Library(lib_path())
        return _default
    return Library(path)


# ---------------------------------------------------------------------------
# torch tensors <-> iamrx_fab
# ---------------------------------------------------------------------------
def fab_of(t, lo):
    """Describe tensor ``t`` of shape (ncomp, nz, ny, nx) (x fastest) whose element
    [n, 0, 0, 0] is index ``lo`` as an iamrx_fab.  ``t`` must be float64 with unit x
    stride; it is NOT copied, the caller keeps it alive."""
    assert t.dtype == torch.float64 and t.dim() == 4 and t.stride(3) == 1
    f = Fab()
    f.p = t.data_ptr()
    ncomp, nz, ny, nx = t.shape
    for d, n in zip(range(3), (nx, ny, nz)):
        f.lo[d] = int(lo[d])
        f.hi[d] = int(lo[d]) + n - 1
    f.jstride = t.stride(2)
    f.kstride = t.stride(1)
    f.nstride = t.stride(0)
    f.ncomp = ncomp
    return f


def alloc_fab(box_lo, box_hi, ncomp, ngrow, device, fill=0.0):
    """Allocate a (ncomp, nz, ny, nx) tensor covering [box_lo-ngrow, box_hi+ngrow] and
    its iamrx_fab view."""
    lo = [box_lo[d] - ngrow for d in range(3)]
    n = [box_hi[d] - box_lo[d] + 1 + 2 * ngrow for d in range(3)]
    t = torch.full((ncomp, n[2], n[1], n[0]), fill, dtype=torch.float64, device=device)
    return t, fab_of(t, lo)


class _CudaView:
    def __init__(self, ptr, nelem):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def tensor_of(fab, device):
    """Zero-copy torch view (ncomp, nz, ny, nx) of a library-owned iamrx_fab."""
    nx, ny, nz = (fab.hi[d] - fab.lo[d] + 1 for d in range(3))
    nelem = fab.nstride * fab.ncomp
    if torch.device(device).type == "cuda":
        flat = torch.as_tensor(_CudaView(fab.p, nelem), device=device)
    else:
        buf = (C.c_double * nelem).from_address(fab.p)
        flat = torch.from_numpy(np.ctypeslib.as_array(buf))
    return torch.as_strided(flat, (fab.ncomp, nz, ny, nx), (fab.nstride, fab.kstride, fab.jstride, 1))


def _stream_ptr(device):
    if torch.device(device).type == "cuda":
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    return C.c_void_p(0)


class Level:
    """amrex BoxArray + DistributionMapping + Geometry of one level (iamrx_level_t)."""

    def __init__(self, lib, geom, boxes, owners=None):
        self.lib, self.geom = lib, geom
        nb = len(boxes)
        arr = (Box * nb)(*[Box.make(lo, hi) for lo, hi in boxes])
        own = (C.c_int * nb)(*(owners if owners is not None else [0] * nb))
        h = C.c_void_p()
        lib.check(lib.iamrx_level_create(C.byref(geom), nb, arr, own, C.byref(h)))
        self.h = h
        self.boxes = boxes

    def num_local(self):
        return self.lib.iamrx_level_num_local(self.h)

    def local_box(self, il):
        b = Box()
        gi = C.c_int()
        self.lib.check(self.lib.iamrx_level_local_box(self.h, il, C.byref(b), C.byref(gi)))
        return (tuple(b.lo), tuple(b.hi)), gi.value

    def close(self):
        if self.h:
            self.lib.iamrx_level_destroy(self.h)
            self.h = None


class NavierStokes:
    """The level time-step object (iamrx_ns_t): NavierStokes::advance / post_init."""

    def __init__(self, lib, level, device, **params):
        self.lib, self.level, self.device = lib, level, device
        p = NSParams()
        lib.iamrx_ns_params_default(C.byref(p))
        for k, v in params.items():
            if not hasattr(p, k):
                raise KeyError(k)
            if k in ("lo_bc", "hi_bc"):
                for d in range(3):
                    getattr(p, k)[d] = int(v[d])
            elif k == "bc_vals":
                for f in range(6):
                    for c in range(5):
                        p.bc_vals[f][c] = float(v[f][c])
            else:
                setattr(p, k, v)
        self.params = p
        h = C.c_void_p()
        lib.check(lib.iamrx_ns_create(level.h, C.byref(p), C.byref(h)))
        self.h = h

    def set_turbulent_forcing(self, nmodes, mode_start, div_free_force, forcedata):
        """forcedata: numpy float64 array [17, as, as, as] indexed [array, kz, ky, kx] (i fastest, TurbulentForcing::forcedata)."""
        import numpy as np
        fd = np.ascontiguousarray(forcedata, dtype=np.float64)
        assert fd.ndim == 4 and fd.shape[0] == 17 and fd.shape[1] == fd.shape[2] == fd.shape[3]
        self.lib.check(self.lib.iamrx_ns_set_turbulent_forcing(self.h, nmodes, mode_start, int(div_free_force), fd.shape[1],
                                                              fd.ctypes.data_as(_P(C.c_double))))

    def init_prob(self, probtype, params):
        arr = (C.c_double * len(params))(*params)
        self.lib.check(self.lib.iamrx_ns_init_prob(self.h, probtype, arr, len(params)))

    def post_init(self):
        dt = C.c_double(0.0)
        self.lib.check(self.lib.iamrx_ns_post_init(self.h, C.byref(dt)))
        return dt.value

    def step(self, dt=-1.0):
        d = C.c_double(dt)
        self.lib.check(self.lib.iamrx_ns_step(self.h, C.byref(d)))
        return d.value

    def step_host(self, host_in, host_out, dt=-1.0):
        """host_in/host_out: lists (one per local box) of CPU float64 tensors of shape
        (5, nz, ny, nx), contiguous (pinned for the bench)."""
        n = len(host_in)
        pin = (C.c_void_p * n)(*[t.data_ptr() for t in host_in])
        pout = (C.c_void_p * n)(*[t.data_ptr() for t in host_out])
        d = C.c_double(dt)
        self.lib.check(self.lib.iamrx_ns_step_host(self.h, pin, pout, C.byref(d)))
        return d.value

    def field(self, which, il=0, valid=True):
        """Zero-copy tensor of a state field; valid=True strips ghost cells."""
        f = Fab()
        self.lib.check(self.lib.iamrx_ns_field(self.h, which, il, C.byref(f)))
        t = tensor_of(f, self.device)
        if not valid:
            return t
        (lo, hi), _ = self.level.local_box(il)
        ext = [0, 0, 0]
        if which == 1:
            ext = [1, 1, 1]
        elif which in (4, 5, 6):
            ext[which - 4] = 1
        sl = []
        for d in (2, 1, 0):
            o = lo[d] - f.lo[d]
            sl.append(slice(o, o + hi[d] - lo[d] + 1 + ext[d]))
        return t[(slice(None),) + tuple(sl)]

    @property
    def time(self):
        return self.lib.iamrx_ns_time(self.h)

    @property
    def nstep(self):
        return self.lib.iamrx_ns_nstep(self.h)

    def last_iters(self):
        it = (C.c_int * 3)()
        self.lib.check(self.lib.iamrx_ns_last_iters(self.h, it))
        return tuple(it)

    def write_plotfile(self, path):
        self.lib.check(self.lib.iamrx_ns_write_plotfile(self.h, str(path).encode()))

    def sums(self):
        """(MASS, TRAC, KINETIC ENERGY) as NavierStokes::sum_integrated_quantities prints them."""
        out = (C.c_double * 3)()
        self.lib.check(self.lib.iamrx_ns_sum_integrated_quantities(self.h, out))
        return tuple(out)

    def close(self):
        if self.h:
            self.lib.iamrx_ns_destroy(self.h)
            self.h = None
