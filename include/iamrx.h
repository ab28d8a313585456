/*
 * iamrx.h -- C ABI of the B200-native IAMR hot path (libiamrx.so).
 *
 * Everything that crosses this boundary is POD: raw DEVICE pointers, index
 * boxes, strides, scalars and a cudaStream_t passed as void*.  No torch, no
 * AMReX, no C++ types.  The layout of a field is the one AMReX's Array4 uses
 * (x unit-stride, then y, then z, component outermost, ghost cells inside the
 * allocation) with the three strides made explicit, so that an adapter inside
 * IAMR can forward `MultiFab::array(mfi)` ({p, begin, end, jstride, kstride,
 * nstride, ncomp}) without a copy.
 *
 * Each entry point names the reference call it stands in for
 * (/root/reference = AMReX-Fluids/IAMR @ f46ba59; "NSB.cpp" =
 * Source/NavierStokesBase.cpp, "NS.cpp" = Source/NavierStokes.cpp).  The
 * arithmetic of those calls lives in AMReX / AMReX-Hydro, which the reference
 * does not vendor (Exec/Make.IAMR:15-19,36-37); see DESIGN.md "Oracle".
 *
 * Error convention (reference: amrex::Abort on bad configuration, "MLMG failed
 * to converge" abort): every function returns int, 0 = ok, <0 = bad argument /
 * CUDA error (iamrx_last_error() gives text), >0 = solver did not converge
 * (value = iterations done).  Nothing throws across the boundary.  There is no
 * CPU fallback: without a usable CUDA device every compute entry returns
 * IAMRX_ERR_NO_DEVICE.
 */
#ifndef IAMRX_H_
#define IAMRX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IAMRX_OK 0
#define IAMRX_ERR_ARG (-1)
#define IAMRX_ERR_CUDA (-2)
#define IAMRX_ERR_NO_DEVICE (-3)
#define IAMRX_ERR_COMM (-4)
#define IAMRX_ERR_NAN (-5)

/* Index-space box, inclusive bounds (amrex::Box smallEnd/bigEnd). */
typedef struct iamrx_box {
  int lo[3];
  int hi[3];
} iamrx_box;

/* One FArrayBox / Array4 view.  `lo`/`hi` bound the ALLOCATED region
 * (valid + ghost) in the index space of the field's own centring; element
 * (i,j,k,n) lives at p[(i-lo[0]) + (j-lo[1])*jstride + (k-lo[2])*kstride +
 * n*nstride]. */
typedef struct iamrx_fab {
  double* p;
  int lo[3];
  int hi[3];
  int64_t jstride;
  int64_t kstride;
  int64_t nstride;
  int ncomp;
  int pad_;
} iamrx_fab;

/* amrex::Geometry subset: domain box (cell indices), mesh spacing, periodicity
 * (NS_setup.cpp / inputs `geometry.is_periodic`). */
typedef struct iamrx_geom {
  iamrx_box domain;
  double dx[3];
  double prob_lo[3];
  int periodic[3];
  int pad_;
} iamrx_geom;

/* amrex::BCType math codes used by BCRec (Source/NS_BC.H:7-55). */
enum {
  IAMRX_BC_INT_DIR = 0,      /* interior / periodic */
  IAMRX_BC_REFLECT_ODD = -1,
  IAMRX_BC_REFLECT_EVEN = 1,
  IAMRX_BC_FOEXTRAP = 2,
  IAMRX_BC_EXT_DIR = 3,
  IAMRX_BC_HOEXTRAP = 4
};

/* amrex::BCRec of ONE component: math BC codes per direction, low and high side (what
 * NavierStokesBase::fetchBCArray hands to the Godunov routines, NSB.cpp:4489,4645,4712). */
typedef struct iamrx_bcrec {
  int lo[3];
  int hi[3];
} iamrx_bcrec;

/* amrex::LinOpBCType subset (MacProj.cpp:1187-1208, Projection.cpp:2436-2464,
 * Diffusion.cpp:1900-1938). */
enum {
  IAMRX_LINOP_PERIODIC = 0,
  IAMRX_LINOP_DIRICHLET = 1,
  IAMRX_LINOP_NEUMANN = 2,
  IAMRX_LINOP_REFLECT_ODD = 3,   /* cell-centred operators only (Diffusion.cpp:1921-1923) */
  IAMRX_LINOP_INFLOW = 4         /* nodal projection only (Projection.cpp:2450,2458) */
};

/* flags of iamrx_compute_aofs_box (arguments of
 * HydroUtils::ComputeFluxesOnBoxFromState, NSB.cpp:4701-4717) */
enum {
  IAMRX_ADV_PPM = 1,               /* godunov_use_ppm: advection_scheme Godunov_PPM (NSB.cpp:552-554,4485); staged kernels */
  IAMRX_ADV_FORCES_IN_TRANS = 2,   /* godunov.use_forces_in_trans, NSB.cpp:556 */
  IAMRX_ADV_IS_VELOCITY = 4,
  IAMRX_ADV_WRITE_FLUXES = 8,      /* also store area-weighted fluxes + edge states */
  IAMRX_ADV_IS_SYNC = 16,          /* aofs -= update, fluxes from U_corr (NSB.cpp:4834) */
  IAMRX_ADV_STAGED = 32,           /* use the staged (global-scratch) kernels instead of the fused tile
                                      kernel: same algorithm, kept for cross-checks and partial tiles */
  IAMRX_ADV_KNOWN_EDGE_STATE = 64  /* known_edge_state (NSB.cpp:4708; MacProj.cpp:776-785): xed/yed/zed are INPUTS,
                                      only fluxes, divergence and the convective term are formed */
};

/* UNVERIFIED-UPSTREAM switches.  The arithmetic of Godunov::ExtrapVelToFaces / ComputeEdgeState lives in
 * AMReX-Hydro, which the reference does not vendor; where this library restates a detail from memory that a site with the
 * upstream source might find different, the choice is a run-time switch (process-wide; mirrored by orc_set_option in the
 * oracle) instead of a constant.  Defaults in brackets. */
enum {
  IAMRX_OPT_SMALL_VEL = 0,     /* [1e-8] |u| below which a face velocity counts as zero when upwinding */
  IAMRX_OPT_SLOPE_ORDER = 1,   /* [4] order of the limited slopes of the PLM trace: 4 or 2 */
  IAMRX_OPT_CORNER_FORM = 2,   /* [0] corner coupling of non-conservative states: 0 flux form - q div(u), 1 advective form */
  IAMRX_OPT_EXTDIR_BOTH = 3    /* [0] ext_dir faces: 1 = both traced states take the boundary value (0: only the outside one,
                                  except for the normal velocity) */
};
int iamrx_set_option(int option, double value);
double iamrx_get_option(int option);

const char* iamrx_last_error(void);
int iamrx_version(void);
/* number of kernels this library has launched since load (bench.py's
 * `gpu_launches`), and reset. */
int64_t iamrx_launch_count(void);
void iamrx_launch_count_reset(void);
int iamrx_device_ok(void);
/* diagnostic: measured fp64 FMA issue rate of the device, in 1e9 thread-instructions/s (bench.py's second roofline
 * for the Godunov kernels, which are FP64-pipe- rather than HBM-bound) */
int iamrx_debug_fp64_peak(double* dp_ginstr_per_s, void* stream);

/* Kernel timing for the roofline report (bench.py): when enabled, every launch of
 * the listed kernel classes that covers >= min_points points is bracketed by CUDA
 * events on its stream.  iamrx_prof_report synchronises and returns the summed
 * device time, launch count and ALGORITHMIC bytes (per-kernel table)
 * of one class since the last reset. */
enum {
  IAMRX_PROF_ABEC_GSRB = 0,   /* one colour pass of the ABec smoother */
  IAMRX_PROF_NODAL_GS = 1,    /* one colour pass of the nodal smoother */
  IAMRX_PROF_AOFS = 2,        /* ComputeAofs on one box (all stages) */
  IAMRX_PROF_EXTRAP = 3,      /* ExtrapVelToFaces on one box (all stages) */
  IAMRX_PROF_ABEC_APPLY = 4,  /* ABec apply / residual */
  IAMRX_PROF_NODAL_ADOTX = 5, /* nodal apply / residual */
  IAMRX_PROF_NCLASS = 6
};
int iamrx_prof_enable(int on, int64_t min_points);
void iamrx_prof_reset(void);
int iamrx_prof_report(int kclass, double* total_ms, int64_t* launches, double* algo_bytes);
/* Time EVERY kernel launch by kernel name (CUDA events; adds two event records per
 * launch while on).  iamrx_prof_dump synchronises, writes "name launches total_ms"
 * lines into buf (truncated to cap) and clears the records; returns the full length. */
int iamrx_prof_all(int on);
int iamrx_prof_dump(char* buf, int cap);

/* ------------------------------------------------------------------------
 * 1. Per-box kernels (one FArrayBox at a time, async on `stream`).
 * ---------------------------------------------------------------------- */

/* MLABecLaplacian red-black Gauss-Seidel colour pass on `bx`
 * (AMReX MLABecLaplacian::Fsmooth, reached from every mlmg.solve in
 * MacProj.cpp:1179, Diffusion.cpp:567,923):
 *   phi += omega/(gamma) * (rhs - (a*alpha*phi - b*div(beta grad phi)))
 * for cells with (i+j+k+redblack) even.  Ghost cells of phi must be filled.
 * acoef may be NULL when a == 0.  ncomp components share bx/by/bz unless
 * bcoef fabs carry ncomp components (MLTensorOp). */
int iamrx_abec_gsrb_box(const iamrx_box* bx, iamrx_fab* phi, const iamrx_fab* rhs,
                        double a, double b, const iamrx_fab* acoef,
                        const iamrx_fab* bcoef_x, const iamrx_fab* bcoef_y,
                        const iamrx_fab* bcoef_z, const double dxinv[3],
                        double omega, int redblack, int ncomp, void* stream);

/* One full red-black sweep (redblack 0, then 1) of the same smoother, phi_in -> phi_out (two
 * different fabs), for a box that spans a fully periodic domain with even extents >= 8
 * (neighbours wrap inside the kernel; ghost cells are not read).  Bit-identical to two
 * iamrx_abec_gsrb_box passes with periodic ghost fills, in one fused launch that reads every
 * array once.  Other boxes: IAMRX_ERR_ARG. */
int iamrx_abec_gsrb_sweep_box(const iamrx_box* bx, iamrx_fab* phi_out, const iamrx_fab* phi_in,
                              const iamrx_fab* rhs, double a, double b, const iamrx_fab* acoef,
                              const iamrx_fab* bcoef_x, const iamrx_fab* bcoef_y,
                              const iamrx_fab* bcoef_z, const double dxinv[3], double omega,
                              int ncomp, void* stream);

/* MLABecLaplacian::Fapply / residual: out = L(phi) (rhs == NULL) or
 * out = rhs - L(phi).  (mlmg.apply Diffusion.cpp:768,1757; residuals inside
 * every solve.) */
int iamrx_abec_apply_box(const iamrx_box* bx, iamrx_fab* out, const iamrx_fab* phi,
                         const iamrx_fab* rhs, double a, double b,
                         const iamrx_fab* acoef, const iamrx_fab* bcoef_x,
                         const iamrx_fab* bcoef_y, const iamrx_fab* bcoef_z,
                         const double dxinv[3], int ncomp, void* stream);

/* MLTensorOp cross terms: out += b * div(cross-flux(eta, vel))
 * (MLTensorOp::apply after the ABec part; Diffusion.cpp:715-768,1708-1757).
 * vel has 3 components and filled ghost cells including edges/corners. */
int iamrx_tensor_cross_box(const iamrx_box* bx, iamrx_fab* out, const iamrx_fab* vel,
                           const iamrx_fab* eta_x, const iamrx_fab* eta_y,
                           const iamrx_fab* eta_z, double b, const double dxinv[3],
                           void* stream);

/* Godunov::ExtrapVelToFaces(vel, forces, u_mac, v_mac, w_mac, h_bcrec, d_bcrec, geom, dt, use_ppm,
 * use_forces_in_trans) (NSB.cpp:4487-4491), one box:
 * vel (3 comps, 3 ghost cells filled incl. the physical-boundary fill), force (3 comps, 1 ghost) ->
 * umac/vmac/wmac on the faces of bx.  bcrec: the BCRec of the three velocity components
 * (h_bcrec/d_bcrec; NULL = interior / periodic everywhere); geom->domain locates the boundary.
 * vel must cover grow(bx,3), force grow(bx,1), the outputs the faces of bx (else IAMRX_ERR_ARG). */
int iamrx_extrap_vel_to_faces_box(const iamrx_box* bx, const iamrx_fab* vel,
                                  const iamrx_fab* force, iamrx_fab* umac,
                                  iamrx_fab* vmac, iamrx_fab* wmac,
                                  const iamrx_bcrec* bcrec,
                                  const iamrx_geom* geom, double dt, int flags,
                                  void* stream);

/* NavierStokesBase::ComputeAofs body for one box (NSB.cpp:4661-4845), argument for argument the
 * HydroUtils::ComputeFluxesOnBoxFromState call of NSB.cpp:4701-4717 followed by
 * ComputeDivergence(mult=-1) -> ComputeConvectiveTerm -> aofs = -update (or aofs -= update, sync).
 * S: ncomp comps, 3 ghosts.  force: ncomp comps, 1 ghost.  divu may be NULL (== 0).
 * umac..wmac: the advecting MAC velocities (1 ghost face layer) that build the edge states;
 * uflux..wflux: the velocities that multiply them into fluxes -- NULL = the MAC velocities, U_corr in
 * the sync call (NSB.cpp:4672-4677).  fx,fy,fz / xed,yed,zed may be NULL unless
 * IAMRX_ADV_WRITE_FLUXES; with IAMRX_ADV_KNOWN_EDGE_STATE xed,yed,zed are inputs.
 * bcrec: BCRec of the ncomp components (NULL = interior).  S must cover grow(bx,3) (bx itself with
 * known edge states), force / divu / the velocities grow(bx,1), aofs bx (else IAMRX_ERR_ARG). */
int iamrx_compute_aofs_box(const iamrx_box* bx, iamrx_fab* aofs, int aofs_comp,
                           const iamrx_fab* S, int s_comp, int ncomp,
                           const iamrx_fab* force, int f_comp, const iamrx_fab* divu,
                           const iamrx_fab* umac, const iamrx_fab* vmac,
                           const iamrx_fab* wmac, const iamrx_fab* uflux,
                           const iamrx_fab* vflux, const iamrx_fab* wflux,
                           iamrx_fab* fx, iamrx_fab* fy,
                           iamrx_fab* fz, iamrx_fab* xed, iamrx_fab* yed,
                           iamrx_fab* zed, const int* iconserv, const iamrx_bcrec* bcrec,
                           const iamrx_geom* geom, double dt, int flags, void* stream);

/* MLNodeLaplacian pieces on one box of NODES (Projection.cpp:2512-2542,
 * NSB.cpp:4106-4118).  sigma is cell-centred with 1 ghost cell. */
int iamrx_nodal_divu_box(const iamrx_box* nbx, iamrx_fab* rhs, const iamrx_fab* vel,
                         const double dxinv[3], void* stream);
int iamrx_nodal_adotx_box(const iamrx_box* nbx, iamrx_fab* out, const iamrx_fab* phi,
                          const iamrx_fab* rhs, const iamrx_fab* sigma,
                          const double dxinv[3], void* stream);
int iamrx_nodal_gs_box(const iamrx_box* nbx, iamrx_fab* phi, const iamrx_fab* rhs,
                       const iamrx_fab* sigma, const double dxinv[3], int color,
                       void* stream);
/* One full sweep of the 8-colour nodal Gauss-Seidel smoother (colours 0..7 in order, the
 * MLNodeLaplacian GPU smoother reached from NodalProjector::project, Projection.cpp:2540),
 * phi_in -> phi_out (two different fabs), for a node box that spans a fully periodic
 * domain with an even number of cells per direction (node hi duplicates node lo; ghost
 * nodes are not read).  Same result as eight iamrx_nodal_gs_box calls with periodic ghost
 * fills in between, from two fused launches.  Other boxes: IAMRX_ERR_ARG. */
int iamrx_nodal_gs_sweep_box(const iamrx_box* nbx, iamrx_fab* phi_out, const iamrx_fab* phi_in,
                             const iamrx_fab* rhs, const iamrx_fab* sigma, const double dxinv[3],
                             void* stream);
/* vel -= sigma*grad(phi); gp = grad(phi) (cell-centred average of the four
 * parallel edge differences) -- NodalProjector::project tail + getGradPhi,
 * MLNodeLaplacian::compGrad (NSB.cpp:4118). Either of vel / gp may be NULL. */
int iamrx_nodal_mknewu_box(const iamrx_box* bx, iamrx_fab* vel, iamrx_fab* gp,
                           const iamrx_fab* phi, const iamrx_fab* sigma,
                           const double dxinv[3], void* stream);

/* ------------------------------------------------------------------------
 * 2. Level objects: a BoxArray + DistributionMapping of one AMR level,
 *    sharded one-rank-per-GPU.
 * ---------------------------------------------------------------------- */
typedef struct iamrx_level_s* iamrx_level_t;
typedef struct iamrx_ns_s* iamrx_ns_t;

/* Communicator.  rank/nranks from the launcher; `nccl_uid` = 128-byte
 * ncclUniqueId created on rank 0 with iamrx_comm_unique_id and broadcast by
 * the caller (torch.distributed in this repo; MPI_Bcast in IAMR).  nranks==1
 * needs no uid and loads no NCCL. */
int iamrx_comm_unique_id(unsigned char uid[128]);
int iamrx_comm_init(int rank, int nranks, const unsigned char uid[128]);
int iamrx_comm_finalize(void);
int iamrx_comm_rank(void);
int iamrx_comm_size(void);
/* Host-supplied transport (optional alternative to iamrx_comm_init).  IAMR's ranks
 * talk MPI through amrex::ParallelDescriptor; a host that owns the
 * communicator registers its (GPU-aware) point-to-point exchange and all-reduce
 * here instead of letting libiamrx open NCCL.  Buffers are DEVICE pointers;
 * counts are in doubles; op as in iamrx_allreduce.  Return 0 on success. */
typedef int (*iamrx_exchange_fn)(void* ctx, int npeers, const int* peers, double* const* sendbuf,
                                 const int64_t* sendcount, double* const* recvbuf,
                                 const int64_t* recvcount, void* stream);
typedef int (*iamrx_allreduce_fn)(void* ctx, double* buf, int n, int op, void* stream);
int iamrx_comm_set_transport(int rank, int nranks, iamrx_exchange_fn exchange,
                             iamrx_allreduce_fn allreduce, void* ctx);
/* ParallelDescriptor::ReduceReal{Min,Max,Sum} on n doubles in DEVICE memory. */
int iamrx_allreduce(double* dev_buf, int n, int op /*0 sum,1 min,2 max*/, void* stream);

/* Level = Geometry + BoxArray + owner rank of each box (DistributionMapping). */
int iamrx_level_create(const iamrx_geom* geom, int nboxes, const iamrx_box* boxes,
                       const int* owner, iamrx_level_t* out);
int iamrx_level_destroy(iamrx_level_t lev);
int iamrx_level_num_local(iamrx_level_t lev);
int iamrx_level_local_box(iamrx_level_t lev, int ilocal, iamrx_box* out, int* global_index);

/* Test hook (pure host logic, no device needed): the FillBoundary copy plan of a
 * level.  Returns the number of regions; fills up to `cap` entries (6 ints per
 * region lo/hi in destination index space, 3 ints per shift: src = dst + shift). */
/* counters of iamrx_fill_boundary-type ghost fills since the last reset: [0] in-place plane exchanges, [1] packed exchanges
 * (pack + send/recv + unpack), [2] local-copy launches next to an in-place exchange, [3] fills with no remote part */
void iamrx_debug_fb_stats(int64_t out[4], int reset);
int iamrx_debug_fb_plan(iamrx_level_t lev, int ixtype, int ng, int cap, int* dst_box, int* src_box,
                        int* region6, int* shift3);

/* ------------------------------------------------------------------------
 * 3. Operators on a level (what Source/MacProj.cpp, Projection.cpp and
 *    Diffusion.cpp call).  Field arguments are arrays of iamrx_fab, one per
 *    LOCAL box in level order; the caller owns all field memory.
 * ---------------------------------------------------------------------- */

/* amrex::IndexType of a field, as the `ixtype` arguments below take it */
enum { IAMRX_IX_CELL = 0, IAMRX_IX_XFACE = 1, IAMRX_IX_YFACE = 2, IAMRX_IX_ZFACE = 3, IAMRX_IX_NODE = 4 };

/* FabArray::FillBoundary(periodicity) on cell (ixtype 0), face-d (1+d) or
 * nodal (4) data (NSB.cpp:1171, MacProj.cpp:1127, Projection.cpp:338). */
int iamrx_fill_boundary(iamrx_level_t lev, iamrx_fab* fabs, int ixtype, int ncomp,
                        int ngrow, void* stream);

/* The HIT tutorial's turbulent forcing block of NavierStokesBase::getForce (Tutorials/HIT/NS_getForce.cpp:205-640, exact path):
 * frc(i,j,k,0..2) += rho(i,j,k) * f(x, time), f = the sum over the low-wavenumber modes of TurbulentForcing::forcedata
 * (17 arrays of array_size^3, TurbulentForcing_def.H:66-82; HOST pointer here), divergence free if div_free_force
 * (turb.div_free_force), modes mode_start .. nmodes per direction (turb.mode_start, turb.nmodes) with kappa <= nmodes / Lmin.
 * rho may be NULL (== 1).  geom: the level geometry (domain lengths and cell centres). */
int iamrx_turbulent_force_box(const iamrx_box* bx, iamrx_fab* frc, const iamrx_fab* rho, const iamrx_geom* geom, double time,
                              int nmodes, int mode_start, int div_free_force, int array_size, const double* forcedata,
                              void* stream);

/* The physical-boundary part of AmrLevel::FillPatch for cell-centred state data (NSB.cpp:4399,4435,3382; the fill
 * functions of NS_bcfill.H:17-167 with the BCRec tables of NS_BC.H:7-55): cells outside the non-periodic sides of the
 * domain.  Call after iamrx_fill_boundary.  bcrec: one BCRec per component; bcvals: the ext_dir face values
 * [x lo, y lo, z lo, x hi, y hi, z hi][ncomp] (NavierStokes::get_bc_values; NULL = 0).  ext_dir ghost cells hold the
 * value ON the face (Docs .../Software.rst:206-213). */
int iamrx_fill_physbc(iamrx_level_t lev, iamrx_fab* fabs, int ncomp, int ngrow,
                      const iamrx_bcrec* bcrec, const double* bcvals, void* stream);

typedef struct iamrx_mg_info {
  double rtol;
  double atol;
  int max_iter;          /* MLMG::setMaxIter, default 200 */
  int max_coarsening;    /* LPInfo::setMaxCoarseningLevel */
  int nu1, nu2;          /* pre/post smooths (AMReX default 2,2) */
  int bottom_sweeps;     /* smoother bottom solve sweeps */
  int verbose;
  double omega;          /* GSRB over-relaxation (AMReX abec_gsrb: 1.15) */
  /* out */
  int iters;
  int maxorder;          /* in: MLLinOp::setMaxOrder -- order of the Dirichlet ghost-cell extrapolation of the cell-centred
                            operators (AMReX default 3; mac_proj.maxorder 4 MacProj.cpp:30; diffuse.max_order 2 Diffusion.cpp:95-96) */
  double resnorm0;
  double resnorm;
  double rhsnorm;
  /* bottom solver (MLMG::setBottomSolver; IAMR's default is "bicgcg", Docs .../RunningProblems.rst:509-516): 0 = bottom_sweeps
   * smoother sweeps [default: exact enough on the 2^3..4^3 coarsest boxes, DESIGN.md 4a], 1 = BiCGStab (MLCGSolver) run to
   * bottom_rtol (AMReX default 1e-4) within bottom_maxiter (200), falling back to the smoother if it breaks down */
  int bottom_solver;
  int bottom_maxiter;
  double bottom_rtol;
  int bottom_iters;      /* out: BiCGStab iterations summed over the V-cycles of the solve */
  int pad_;
} iamrx_mg_info;

void iamrx_mg_info_default(iamrx_mg_info* info);

/* MacProj::mlmg_mac_solve (MacProj.cpp:1084-1184) = Hydro::MacProjector
 * {ctor, setDomainBC, project}: beta_d = (1/rhs_scale)/avg(rho) on faces,
 * solve -div(beta grad phi) = -(div(umac) - rhs), umac -= beta grad phi.
 * rho: cell fabs with >=1 ghost (filled, physical boundaries included); umac[d]: face fabs; phi: cell fab,
 * 1 ghost; rhs may be NULL (divu == 0).  lobc/hibc: LinOpBCType of the six domain sides as set_mac_solve_bc
 * builds them (MacProj.cpp:1187-1208: outflow -> Dirichlet, every other non-periodic side -> Neumann; NULL =
 * periodic); the ghost cells of phi on entry are the level BC (setLevelBC(0, mac_phi), MacProj.cpp:1168) and
 * info->maxorder the extrapolation order (setMaxOrder, :1172).
 * On a level > 0 (boxes that do not tile the domain) call iamrx_set_coarse_fine_bc(lev, crse_lev, phi, cphi, 1, stream) first
 * (setCoarseFineBC(cphi, ratio) :1164-1167): every box side that borders coarse cells is then a coarse-fine Dirichlet side. */
int iamrx_mac_project(iamrx_level_t lev, iamrx_fab* umac, iamrx_fab* vmac,
                      iamrx_fab* wmac, const iamrx_fab* rho, const iamrx_fab* rhs,
                      iamrx_fab* phi, double rhs_scale, const int lobc[3],
                      const int hibc[3], iamrx_mg_info* info, void* stream);

/* Hydro::MacProjector::getFluxes (MacProj.cpp:1181-1183; "fluxes = -B grad phi"): face fluxes of the
 * phi returned by the last iamrx_mac_project on this level, with that call's beta.  MacProj::mac_sync_solve
 * turns them into U_corr (MacProj.cpp:459-468).  fx,fy,fz: face fabs of every local box; phi: 1 ghost. */
int iamrx_mac_get_fluxes(iamrx_level_t lev, iamrx_fab* fx, iamrx_fab* fy, iamrx_fab* fz,
                         iamrx_fab* phi, void* stream);

/* Projection::doMLMGNodalProjection (Projection.cpp:2385-2567) =
 * Hydro::NodalProjector{ctor, setDomainBC, project, getGradPhi}: solve
 * div(sigma grad phi) = div(vel) on nodes, vel -= sigma grad phi,
 * gp (=|+=) grad phi.  vel: 3 comps >=1 ghost; sigma: 1 comp 1 ghost;
 * phi: nodal 1 ghost (initial guess in, solution out); gp: 3 comps.
 * On a level > 0 (Projection::level_project :236-257; boxes that do not tile the domain): the nodes on the boundary of the fine
 * region that border coarse cells are Dirichlet nodes and KEEP the values of phi on entry (the coarse pressure interpolated by
 * FillCoarsePatch: iamrx_fill_coarse_patch_nodal), the interior nodes are the initial guess (IAMR zeroes them).  One rectangular
 * patch of boxes: the boundary planes are cut off the boxes' active node ranges; any other shape (re-entrant edges, partly covered
 * sides): the boundary nodes are found node by node and reset after every kernel that may write them (IAMRX_NODAL_CF_MASK=1 forces
 * this path for rectangular patches too). */
int iamrx_nodal_project(iamrx_level_t lev, iamrx_fab* vel, const iamrx_fab* sigma,
                        iamrx_fab* phi, iamrx_fab* gp, int increment_gp,
                        const int lobc[3], const int hibc[3], iamrx_mg_info* info,
                        void* stream);

/* Diffusion: MLABecLaplacian / MLTensorOp solve and apply
 * (Diffusion.cpp:327-567, 715-768, 858-923, 1708-1757).
 *   (a*acoef - b*div(eta grad))soln = rhs   [+ tensor cross terms if tensor]
 * eta_[xyz]: face fabs, 1 comp.  soln: ncomp comps (3 if tensor), 1 ghost.
 * bc: domain boundary conditions as Diffusion::setDomainBC builds them from the BCRec of each component
 * (Diffusion.cpp:1887-1999: ext_dir -> Dirichlet, foextrap/hoextrap/reflect_even -> Neumann, reflect_odd ->
 * reflect_odd), one set per component for the tensor operator (:711-724), plus setMaxOrder; NULL = periodic.
 * The ghost cells of soln on entry are the level BC (setLevelBC(0, &Soln) :743-744, 886-887): the Dirichlet
 * values ON the faces, as the FillPatch that precedes the call leaves them. */
typedef struct iamrx_linop_bc {
  int lo[3][3];   /* [component][direction] IAMRX_LINOP_* (component 0 only unless ncomp > 1) */
  int hi[3][3];
  int maxorder;   /* diffuse.max_order / diffuse.tensor_max_order = 2 (Diffusion.cpp:95-96,102-103) */
  int pad_;
} iamrx_linop_bc;
int iamrx_diffusion_apply(iamrx_level_t lev, int tensor, int ncomp, iamrx_fab* out,
                          iamrx_fab* soln, double a, double b, const iamrx_fab* acoef,
                          const iamrx_fab* eta_x, const iamrx_fab* eta_y,
                          const iamrx_fab* eta_z, const iamrx_linop_bc* bc, void* stream);
/* Diffusion::computeExtensiveFluxes (Diffusion.cpp:1463-1537) for the scalar operator: f_d = fac * area_d * (-b eta_d dsoln/dx_d) on the
 * faces of every local box (MLMG::getFluxes, then the area weighting) -- the fluxes diffuse_scalar gives to the viscous flux
 * registers (Diffusion.cpp:560-566).  soln: ncomp components, 1 ghost: the cells beyond physical sides as the solve left them
 * (setFinalFillBC), interior / periodic ghost cells are refilled here.  fx, fy, fz: face fabs with ncomp components. */
int iamrx_diffusion_get_fluxes(iamrx_level_t lev, int ncomp, iamrx_fab* fx, iamrx_fab* fy, iamrx_fab* fz, iamrx_fab* soln, double b,
                               const iamrx_fab* eta_x, const iamrx_fab* eta_y, const iamrx_fab* eta_z, double fac, void* stream);
int iamrx_diffusion_solve(iamrx_level_t lev, int tensor, int ncomp, iamrx_fab* soln,
                          const iamrx_fab* rhs, double a, double b,
                          const iamrx_fab* acoef, const iamrx_fab* eta_x,
                          const iamrx_fab* eta_y, const iamrx_fab* eta_z,
                          const iamrx_linop_bc* bc, iamrx_mg_info* info, void* stream);

/* ------------------------------------------------------------------------
 * 3b. Two-level coupling: inter-level transfer operators and the coarse-fine flux register (refinement ratio 2).
 *     Building blocks of NavierStokesBase::avgDown_StatePress (NSB.cpp:4125-4191), FillPatchTwoLevels' interpolaters
 *     (NS_setup.cpp:211,228-230,331; NSB.cpp:1127) and the advective flux register (NSB.cpp:4848-4889,5083-5096;
 *     NS.cpp:1794-1795).  The two-level time stepping itself (subcycling, coarse-fine solver boundaries, mac_sync,
 *     level_sync) is not driven by this library.
 * ---------------------------------------------------------------------- */
/* amrex::average_down (cells, ixtype 0: mean of the 8 children), average_down_faces (ixtype 1..3: mean of the 4 fine
 * faces on the coarse face), average_down_nodal (ixtype 4: injection) on the coarse box cbx (cell index space). */
int iamrx_average_down_box(const iamrx_box* cbx, iamrx_fab* crse, const iamrx_fab* fine, int ncomp, int ixtype, void* stream);
enum {
  IAMRX_INTERP_CELL_CONS = 0,       /* cell_cons_interp (State_Type, NS_setup.cpp:211,228-230): conservative linear, MC-limited
                                       slopes scaled to keep every fine value inside the range of the 3^3 coarse neighbourhood */
  IAMRX_INTERP_NODE_BILINEAR = 1,   /* node_bilinear_interp (Press_Type, NS_setup.cpp:331) */
  IAMRX_INTERP_FACE_LINEAR_X = 2,   /* face_linear_interp (u_mac in create_umac_grown, NSB.cpp:1127) */
  IAMRX_INTERP_FACE_LINEAR_Y = 3,
  IAMRX_INTERP_FACE_LINEAR_Z = 4
};
/* fine (on the cells / nodes / faces of the FINE cell box fbx) = interpolation of crse.  cell_cons needs one filled ghost
 * cell around the coarsened box (interior / periodic; physical-boundary one-sided slopes are not implemented). */
int iamrx_interp_box(int kind, const iamrx_box* fbx, iamrx_fab* fine, const iamrx_fab* crse, int ncomp, void* stream);
/* Flux register between a coarse level and the fine level that refines part of it (with several ranks the coarse and the fine
 * side of an interface may live on different ones: fine_add then sums through one replicated array + all-reduce and is collective): the coarse cells
 * that border the fine grids from outside accumulate dt (sum of fine fluxes - coarse flux) / vol with the sign of the face,
 * for AREA-WEIGHTED fluxes and vol = coarse cell volume (the "dx := volume" convention of NSB.cpp:4878-4889).
 * crse_add / fine_add take the face-flux fabs of every local box of the respective level (x, y, z arrays in local-box
 * order); reflux adds scale * register to the coarse state (NS.cpp:1794-1799). */
/* amrex::FillPatchTwoLevels for cell-centred data with the conservative linear interpolater: the ghost cells of every local fine
 * box (ngrow layers) that no fine box covers are interpolated from the coarse level's data at `time` -- linear in time between
 * crse_old (t_old) and crse_new (t_new); crse_old may be NULL -- ghost cells a fine neighbour (or its periodic image) covers get
 * the fine data, cells outside a non-periodic domain the physical boundary fill (bcrec / bcvals as in iamrx_fill_physbc).  What
 * AmrLevel::FillPatch does for State_Type on a level that does not cover the domain (NSB.cpp:4399,4435; NS_setup.cpp:211).  The
 * valid cells of `fine` hold the fine data at `time` and are not touched.  fine: one fab per local fine box; crse_*: one per local
 * coarse box.  Collective over the ranks of the communicator.  (State_Type is a Point-in-time quantity; for an Interval quantity
 * such as Gradp_Type, NS_setup.cpp:339-341, StateData does not interpolate in time: pass crse_old = NULL and the data of the
 * interval that contains `time` as crse_new.) */
int iamrx_fillpatch_two_levels(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* fine, const iamrx_fab* crse_old,
                               const iamrx_fab* crse_new, double t_old, double t_new, double time, int ncomp, int ngrow,
                               const iamrx_bcrec* bcrec, const double* bcvals, void* stream);

/* amrex::average_down between two levels (NavierStokesBase::avgDown / average_down, NSB.cpp:3913-3990; NS.cpp:1840-1933 for the state and
 * the pressure): crse[scomp .. scomp+ncomp) = average of fine[scomp ..) wherever a fine box covers the coarse level, untouched elsewhere.
 * ixtype CELL: mean of the 8 children; NODE: injection; faces: mean of the 4 fine faces.  fine / crse: one fab per local box of the
 * respective level (no ghost cells needed).  Collective (the fine and the coarse boxes of a region may live on different ranks). */
int iamrx_average_down(iamrx_level_t fine_lev, iamrx_level_t crse_lev, const iamrx_fab* fine, iamrx_fab* crse, int scomp, int ncomp,
                       int ixtype, void* stream);

/* NavierStokesBase::create_umac_grown on a level > 0 (NSB.cpp:1108-1310): the ghost faces (one ghost cell) of the fine MAC velocities from
 * FillPatchTwoLevels with face_linear_interp -- coarse values on coincident faces, the mean of the two coarse faces in between; faces a
 * fine neighbour (or its periodic image) owns take its values -- then the divergence correction: a ghost cell the fine level does not
 * cover and with exactly one face neighbour in the valid / covered region gets its OUTER face velocity from div(u_mac) = divu
 * (0 if divu is NULL); grid edges and corners are left as interpolated.  u/v/wmac_f: face fabs of every local fine box with ONE
 * ghost cell; u/v/wmac_c: face fabs of every local coarse box; divu: fine cell fabs with one ghost cell, or NULL.  Cells outside a
 * non-periodic domain are not touched.  Collective. */
int iamrx_create_umac_grown(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* umac_f, iamrx_fab* vmac_f, iamrx_fab* wmac_f,
                            const iamrx_fab* umac_c, const iamrx_fab* vmac_c, const iamrx_fab* wmac_c, const iamrx_fab* divu, void* stream);

/* MLLinOp::setCoarseFineBC(crse, ratio) as MacProj::mlmg_mac_solve calls it on a level > 0 (MacProj.cpp:1164-1167;
 * Diffusion.cpp:395,518 for the scalar viscous solves, :1608 for getViscTerms; ratio 2): the boundary values of a fine-level cell-centred solve along
 * the coarse-fine interface, InterpBndryData::setBndryValues at order 3 -- the coarse field interpolated quadratically in the two
 * tangential directions (centred differences where the tangential coarse neighbours are neither under the fine level nor outside
 * the domain, one-sided otherwise, plus the mixed term) -- written into the ONE ghost-cell layer of `fine` beyond every box side
 * that touches coarse cells.  A value stands for the coarse cell-centre plane, i.e. one FINE cell beyond the face.  Pass `fine`
 * on to iamrx_mac_project / iamrx_diffusion_solve / iamrx_diffusion_apply on the fine level (setLevelBC(0, phi): "the ghost cells
 * of phi on entry are the level BC"): on a level whose boxes do not tile the domain those solvers treat every box side that is
 * neither a domain face nor covered by another fine box as a coarse-fine Dirichlet side (Lagrange extrapolation of order
 * maxorder through that value and the interior cells, the location fixed in physical space on the coarser multigrid levels),
 * and the operator is non-singular.  fine: cell fabs of every local fine box, >= 1 ghost; crse: cell fabs of every local coarse
 * box (cells under the fine level are never read).  The tensor operator on such a level is not implemented.  Collective. */
int iamrx_set_coarse_fine_bc(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* fine, const iamrx_fab* crse, int ncomp,
                             void* stream);

/* AmrLevel::FillCoarsePatch of Press_Type as Projection::level_project calls it on a level > 0 (Projection.cpp:236-239:
 * LevelData[level]->FillCoarsePatch(P_new, 0, cur_pres_time, Press_Type, 0, 1)): EVERY node of every local fine box takes the coarse
 * pressure of the time INTERVAL that contains `time` -- Press_Type is a StateDescriptor::Interval quantity (NS_setup.cpp:329-331), so
 * StateData picks crse_new if time lies in [t_new_start, t_new_stop] (within 1e-3 of its length), else crse_old (whose interval ends
 * at t_new_start; may be NULL), and does not interpolate in time -- interpolated in space with node_bilinear_interp (NS_setup.cpp:331).  level_project then zeroes the interior nodes of each grid (:241-257), which leaves the
 * coarse-fine boundary data iamrx_nodal_project keeps.  Nodal fabs, no ghost nodes needed.  Collective. */
int iamrx_fill_coarse_patch_nodal(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* fine, const iamrx_fab* crse_old,
                                  const iamrx_fab* crse_new, double t_new_start, double t_new_stop, double time, void* stream);

/* NavierStokesBase::SyncInterp (NSB.cpp:3071-3255): interpolate a coarse-level sync correction (Vsync / Ssync, or the velocity
 * correction of level_sync) onto the fine level -- coarse data with periodic images and the HOMOGENEOUS ext_dir fill of the original
 * quantity's BCRec (HomExtDirFill), interpolated with pc_interp or cell_cons_interp; increment != 0: fine[dest..] += dt_clev * I(crse)
 * (:3209-3236), else fine[dest..] = I(crse).  fine_sync: one fab per local fine box with >= dest_comp + ncomp components; crse_sync: one
 * per local coarse box with >= src_comp + ncomp.  Collective.  (CellConsLin_T / CellConsProt_T are not implemented.) */
enum { IAMRX_SYNC_PC = 0, IAMRX_SYNC_CELL_CONS = 1 };
int iamrx_sync_interp(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* fine_sync, int dest_comp, const iamrx_fab* crse_sync,
                      int src_comp, int ncomp, int increment, double dt_clev, int which_interp, const iamrx_bcrec* bcrec, void* stream);
/* NavierStokesBase::SyncProjInterp (NSB.cpp:3258-3336): the coarse sync-projection correction phi interpolated with
 * node_bilinear_interp and added to BOTH pressure time levels of the fine level.  Nodal fabs, no ghost nodes needed.  Collective. */
int iamrx_sync_proj_interp(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* p_new, iamrx_fab* p_old, const iamrx_fab* phi_crse,
                           void* stream);

typedef struct iamrx_fluxreg_s* iamrx_fluxreg_t;
int iamrx_fluxreg_create(iamrx_level_t crse, iamrx_level_t fine, int ncomp, iamrx_fluxreg_t* out);
int iamrx_fluxreg_destroy(iamrx_fluxreg_t reg);
int iamrx_fluxreg_num_patches(iamrx_fluxreg_t reg);
int iamrx_fluxreg_reset(iamrx_fluxreg_t reg, void* stream);
int iamrx_fluxreg_crse_add(iamrx_fluxreg_t reg, const iamrx_fab* fx, const iamrx_fab* fy, const iamrx_fab* fz, double dt,
                           double vol_crse, void* stream);
int iamrx_fluxreg_fine_add(iamrx_fluxreg_t reg, const iamrx_fab* fx, const iamrx_fab* fy, const iamrx_fab* fz, double dt,
                           double vol_crse, void* stream);
int iamrx_fluxreg_reflux(iamrx_fluxreg_t reg, iamrx_fab* crse_state, int scomp, double scale, void* stream);
int iamrx_fluxreg_field(iamrx_fluxreg_t reg, int ilocal, iamrx_fab* out);

/* SyncRegister (Source/SyncRegister.cpp; used by Projection::level_project :399-426 and Projection::MLsyncProject): the nodal
 * register between a coarse level and the next finer one that collects the divergence residuals of the two level projections
 * and hands their sum to the sync projection as its right-hand side.  Held on the nodes of the coarse level; the register's node
 * set B is the six boundary planes of every coarsened fine grid (the FabSets bndry[face], :20-47) and their periodic images.
 *   crse_init : CrseInit (:288-300)  reg = mult * sync_resid_crse on B (the register is reset first).
 *   fine_add  : FineAdd (:351-607)   reg += on B: for every fine grid, on each boundary plane of its coarsened node box, the fine
 *               residual restricted with the tent weights (r - m)(r - n) r_dir / prod(r^2) of the two tangential directions
 *               (halved on the centre lines); fine nodes on the edges of the grid count 1/2 and on its corners 1/3 (the planes of a
 *               grid overlap there); nodes outside the grid are zero (the residual's ghost nodes, Projection.cpp:375-376);
 *               coarse nodes on non-periodic domain planes are doubled once per direction; periodic images add
 *               (FabSet::plusFrom with periodicity).  mult = 1 / crse_dt_ratio (Projection.cpp:424-425).
 *   init_rhs  : InitRHS (:49-285)    rhs = reg, zero on the node planes of outflow sides (phys_lo / phys_hi: the ns.lo_bc / hi_bc
 *               codes, 2 = outflow; NULL = none) and on nodes whose eight surrounding coarse cells all lie under the fine level
 *               (cells beyond a non-periodic side counted as their mirror images) -- IF that count of covered cells exceeds
 *               mask_maxcount.  The reference's threshold is AMREX_D_TERM(SPACEDIM, *SPACEDIM, *SPACEDIM) - 0.5 (:265): 3.5 in
 *               2-D (all four cells covered) but 26.5 in 3-D, which a count of at most 8 never exceeds, so in 3-D the reference's
 *               mask is all ones.  mask_maxcount <= 0 selects that reference behaviour (26.5); 7.5 is what the 2-D expression
 *               generalises to (nodes interior to the fine level are zeroed).
 * sync_resid_crse / rhs: nodal fabs of every local coarse box; sync_resid_fine: nodal fabs of every local fine box.  Unlike the
 * reference the residual arrays are NOT scaled in place by mult.  Coarse boxes that share nodes must hold the same residual there
 * (each node of B takes it once).  CompAdd (three or more levels) is not implemented.  create / fine_add are collective. */
typedef struct iamrx_syncreg_s* iamrx_syncreg_t;
int iamrx_syncreg_create(iamrx_level_t crse_lev, iamrx_level_t fine_lev, double mask_maxcount, iamrx_syncreg_t* out);
int iamrx_syncreg_destroy(iamrx_syncreg_t reg);
int iamrx_syncreg_crse_init(iamrx_syncreg_t reg, const iamrx_fab* sync_resid_crse, double mult, void* stream);
int iamrx_syncreg_fine_add(iamrx_syncreg_t reg, const iamrx_fab* sync_resid_fine, double mult, void* stream);
int iamrx_syncreg_init_rhs(iamrx_syncreg_t reg, iamrx_fab* rhs, const int phys_lo[3], const int phys_hi[3], void* stream);
/* which: 0 the register, 1 the indicator of B, 2 the InitRHS mask (nodal fab of local coarse box ilocal; for inspection) */
int iamrx_syncreg_field(iamrx_syncreg_t reg, int which, int ilocal, iamrx_fab* out);

/* MacProj::mac_sync_solve (MacProj.cpp:359-479) on the coarse level of a pair: the right-hand side is the MAC register refluxed with
 * scale -1 (SUM{MR / VOL} in the coarse cells next to the fine grids, zero elsewhere and under them; the register must have been
 * filled with area-weighted face velocities: crse_add / fine_add with dt = 1 and vol = the coarse cell volume), plus the optional
 * Rhs_increment, negated; mac_sync_phi is zeroed and solved for with rhs_scale = 2 / dt and no div(umac) term; Ucorr = -(-B grad phi).
 * rho_half: >= 1 filled ghost cell; ucorr / vcorr / wcorr: face fabs of every local box; lobc / hibc / info as iamrx_mac_project. */
int iamrx_mac_sync_solve(iamrx_level_t lev, iamrx_fluxreg_t mac_reg, const iamrx_fab* rho_half, const iamrx_fab* rhs_increment,
                         iamrx_fab* ucorr, iamrx_fab* vcorr, iamrx_fab* wcorr, iamrx_fab* mac_sync_phi, double dt, const int lobc[3],
                         const int hibc[3], iamrx_mg_info* info, void* stream);

/* ------------------------------------------------------------------------
 * 4. The level time step: NavierStokes::advance (NS.cpp:543-691) and the
 *    start-up sequence NavierStokes::post_init (NS.cpp:1254-1432) for a
 *    single-level periodic problem.  This is the caller of sections 1-3 in
 *    this repo (the host driver that IAMR's own NavierStokes class is in the
 *    reference).
 * ---------------------------------------------------------------------- */
typedef struct iamrx_ns_params {
  double cfl;            /* ns.cfl */
  double visc_coef;      /* ns.vel_visc_coef */
  double scal_diff_coef; /* ns.scal_diff_coefs of the tracer: > 0 adds its viscous term to the advection forcing and the
                            Crank-Nicolson solve of NavierStokes::scalar_diffusion_update (NS.cpp:858-1000) ->
                            Diffusion::diffuse_scalar (Diffusion.cpp:207-600), rho_flag 0 (tracer) / 2 (conservative tracer) */
  double be_cn_theta;    /* ns.be_cn_theta, NSB.cpp:124 */
  double change_max;     /* ns.change_max, NSB.cpp:101 */
  double init_shrink;    /* ns.init_shrink */
  double fixed_dt;       /* ns.fixed_dt (<=0: off) */
  double gravity;        /* ns.gravity */
  double visc_tol;       /* ns.visc_tol, NSB.cpp:122 */
  double mac_tol, mac_abs_tol;     /* MacProj.cpp:49-51 */
  double proj_tol, proj_abs_tol;   /* Projection.cpp:19-21 */
  int init_iter;         /* ns.init_iter, NSB.cpp:98 */
  int init_vel_iter;     /* ns.init_vel_iter, NSB.cpp:99 */
  int do_init_proj;
  int use_forces_in_trans;
  int verbose;
  int conservative_tracer; /* ns.do_cons_trac */
  int mg_verbose;
  int godunov_ppm;       /* ns.advection_scheme = Godunov_PPM instead of the default Godunov_PLM (NSB.cpp:169,552-554) */
  int do_scalminmax;     /* ns.do_scalminmax (NSB.cpp:140,2907-2935): clamp the advected tracer to the old 3x3x3 range */
  int do_mom_diff;       /* ns.do_mom_diff (NSB.cpp:167,3358-3470,3609-3616; NS.cpp:606-623,1016): advect and diffuse momentum rho*u */
  int bottom_solver;     /* bottom solver of all three multigrid solves: 0 smoother sweeps, 1 BiCGStab (iamrx_mg_info.bottom_solver) */
  /* physical boundaries (ns.lo_bc / ns.hi_bc, NS.cpp:90-94; codes of inputs.3d.taylorgreen:100-102): 0 interior / periodic,
   * 1 inflow, 2 outflow, 3 symmetry, 4 slip wall, 5 no-slip wall.  A periodic direction must carry 0, a non-periodic one must
   * not.  The step driver implements all of them: walls and symmetry planes (3, 4, 5), inflow (ext_dir values below) and outflow,
   * under gravity with Projection::set_outflow_bcs / computeRhoG (Projection.cpp:1721-2379) for outflow faces in x, y or on top;
   * an outflow face at the bottom together with gravity is refused (the reference aborts, :1957). */
  int lo_bc[3];
  int hi_bc[3];
  /* Dirichlet face values [x lo, y lo, z lo, x hi, y hi, z hi][u, v, w, rho, tracer]: the xlo.velocity / xlo.density ...
   * blocks of NS.cpp:108-237 (e.g. zhi.velocity = 1 0 0, the lid of Tutorials/LidDrivenCavity) */
  double bc_vals[6][5];
} iamrx_ns_params;

void iamrx_ns_params_default(iamrx_ns_params* p);

int iamrx_ns_create(iamrx_level_t lev, const iamrx_ns_params* p, iamrx_ns_t* out);
int iamrx_ns_destroy(iamrx_ns_t ns);
/* prob_init: probtype 11 TaylorGreen (prob_init.cpp:509-560) with (a,b,c, velocity_factor, density); probtype 5
 * DoubleShearLayer (:346-405, direction 1, uniform in z; density, interface_width, blob centre x y z, blob radius);
 * probtype 20 the HIT tutorial's field (Tutorials/HIT/prob_init.cpp:100-131; turb_scale, density [, amplitude of a
 * synthetic density variation]); probtype 100 a synthetic variable-density Taylor-Green field. */
int iamrx_ns_init_prob(iamrx_ns_t ns, int probtype, const double* prob_params, int nparams);
/* Switch on the HIT tutorial's turbulent forcing in the step driver (USE_TURBULENT_FORCING; inputs.3d.forced: turb.nmodes = 4):
 * getForce adds rho * f(x, t) at prev_time (predict_velocity, velocity_advection, initial diffusion update), half_time
 * (velocity update) and cur_time (estTimeStep).  forcedata: TurbulentForcing::forcedata (host or device memory; copied);
 * NULL or nmodes <= 0 switches it off.  Triply periodic domains only. */
int iamrx_ns_set_turbulent_forcing(iamrx_ns_t ns, int nmodes, int mode_start, int div_free_force, int array_size,
                                   const double* forcedata);
/* NavierStokes::post_init: initial velocity projection, initial dt, initial
 * pressure iterations. Returns dt for the first step in *dt0. */
int iamrx_ns_post_init(iamrx_ns_t ns, double* dt0);
/* One coarse time step: computeNewDt (unless dt>0 is forced) + advance.
 * *dt_io: in = forced dt or <=0; out = dt used. */
int iamrx_ns_step(iamrx_ns_t ns, double* dt_io);
double iamrx_ns_time(iamrx_ns_t ns);
int iamrx_ns_nstep(iamrx_ns_t ns);
/* State access: which = 0 State_new (u,v,w,rho,tracer; 1 ghost), 1 Press_new
 * (nodal, 1 ghost), 2 Gradp_new (3 comps, 1 ghost), 3 State_old,
 * 4..6 u_mac (face x,y,z), 7 aofs. */
int iamrx_ns_field(iamrx_ns_t ns, int which, int ilocal, iamrx_fab* out);
/* MacProj::mac_sync_compute (MacProj.cpp:505-731) on the level this object advances, to be called right after iamrx_ns_step: the sync
 * advection with the correction velocities of iamrx_mac_sync_solve -- ComputeAofs(is_sync = true, Ucorr) of the velocity (momenta with
 * do_mom_diff) and of the scalars on the state at prev_time with the forcing of the step, edge states upwinded with the step's u_mac,
 * fluxes taken with Ucorr; the results ACCUMULATE into vsync (3 components) and ssync (density, tracer; may be NULL).  ucorr / vcorr /
 * wcorr: face fabs of every local box; vsync / ssync: cell fabs, no ghost cells.  (The mac-register update of level > 0, :712-730, is
 * the caller's iamrx_fluxreg_fine_add.) */
int iamrx_ns_mac_sync_compute(iamrx_ns_t ns, const iamrx_fab* ucorr, const iamrx_fab* vcorr, const iamrx_fab* wcorr, iamrx_fab* vsync,
                              iamrx_fab* ssync, double dt);
/* Host-buffer variant of one step (the e2e path): copies the caller's HOST
 * valid-region state (5 comps, no ghosts, box order = local boxes) to the
 * device, advances, copies the new state back.  Both copies are inside. */
int iamrx_ns_step_host(iamrx_ns_t ns, const double* const* host_state_in,
                       double* const* host_state_out, double* dt_io);
/* Amr::writePlotFile (main.cpp:135; plotfile type "NavierStokes-V1.1", NavierStokesBase.cpp:3343-3352): the new-time cell-centred
 * state (x_velocity y_velocity z_velocity density tracer gradpx gradpy gradpz) as an AMReX plotfile directory
 * (Header, Level_0/Cell_H, Level_0/Cell_D_<rank>, job_info) that fcompare / yt / Amrvis read -- the way a site with an AMReX build
 * can diff this library against IAMR.  Collective over the ranks of the level; every rank writes its own boxes. */
int iamrx_ns_write_plotfile(iamrx_ns_t ns, const char* dir);
/* solver statistics of the last step: iterations of {mac, visc, nodal}. */
int iamrx_ns_last_iters(iamrx_ns_t ns, int iters[3]);
/* NavierStokes::sum_integrated_quantities (NS.cpp:1046-1080): volume-weighted sums of the new state over the level, all
 * ranks: out = {MASS (density), TRAC (tracer), KINETIC ENERGY (derkeng: 0.5 rho |u|^2, NS_derive.cpp:266-300)} -- the
 * numbers IAMR prints as "TIME= .. MASS= .." every ns.sum_interval steps. */
int iamrx_ns_sum_integrated_quantities(iamrx_ns_t ns, double out[3]);

#ifdef __cplusplus
}
#endif
#endif /* IAMRX_H_ */
