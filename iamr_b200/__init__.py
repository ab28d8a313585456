"""iamr_b200 -- B200-native hot path of AMReX-Fluids/IAMR (Godunov advection, tensor
diffusion, MAC + nodal projection multigrid) behind a C ABI.

The product is ``iamr_b200/libiamrx.so`` (CUDA, sm_100a; sources in ``iamr_b200/csrc``,
ABI in ``include/iamrx.h``).  This package is only the ctypes binding used by the
tests and ``bench.py``: torch provides device memory and ``torch.distributed``, nothing
else.  There is no CPU fallback: loading fails loudly if the library is missing and
every compute entry point returns ``IAMRX_ERR_NO_DEVICE`` without a CUDA device.
"""
from .binding import (  # noqa: F401
    Box, Fab, Geom, MGInfo, NSParams, IamrxError, Library, load, lib_path,
    Level, NavierStokes, fab_of, alloc_fab, tensor_of, CELL, XFACE, YFACE, ZFACE, NODE, BCRec, LinopBC, OPT_SMALL_VEL, OPT_SLOPE_ORDER, OPT_CORNER_FORM, OPT_EXTDIR_BOTH,
    LINOP_PERIODIC, LINOP_DIRICHLET, LINOP_NEUMANN, LINOP_REFLECT_ODD, LINOP_INFLOW,
    BC_INT_DIR, BC_REFLECT_ODD, BC_REFLECT_EVEN, BC_FOEXTRAP, BC_EXT_DIR, BC_HOEXTRAP, SYNC_PC, SYNC_CELL_CONS,
    ADV_PPM, ADV_FORCES_IN_TRANS, ADV_IS_VELOCITY, ADV_WRITE_FLUXES, ADV_IS_SYNC, ADV_STAGED, ADV_KNOWN_EDGE_STATE,
)
