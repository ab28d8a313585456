// kernels.h -- internal C++ interface of the CUDA kernels (one level below the
// C ABI in include/iamrx.h).  All functions are asynchronous on `s` and return
// an IAMRX_* status.
#pragma once
#include "common.h"

namespace ix {
namespace k {

struct Abec {
  double a, b;      // scalars (setScalars)
  C4 acoef;         // may be null if a == 0
  C4 bx, by, bz;    // face coefficients; bncomp comps
  int bncomp;       // 1 or ncomp
  double dxinv[3];
  // constant-coefficient fast path (CellMG detects it when the coefficients are set): every face coefficient of
  // component n in direction d equals cb[n][d] (cc) -- the kernels then read no face-coefficient array; cac: acoef is the
  // constant ca as well (constant viscosity is the rule in IAMR runs; a constant density stops being bitwise constant after
  // the first conservative update)
  int cc = 0, cac = 0;
  double cb[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double ca = 0;
};

// --- cell-centred ABec (abec.cu) -----------------------------------------
// wrapmask bit d: bx spans the periodic domain in direction d -> neighbours are read with
// periodic wrap inside the kernel and phi's ghost cells in that direction are not touched.
// gb (optional): the box touches non-periodic domain faces -- f0[comp][xlo,xhi,ylo,yhi,zlo,zhi] = coefficient of the adjacent
// interior cell in the boundary ghost-cell formula (linop_bc_f0; 0 for sides that are not domain faces)
// even / odd[comp]: bit s = the HOMOGENEOUS ghost cell beyond side s is + / - the adjacent cell (Neumann / reflect_odd or order-2
// Dirichlet): the kernels evaluate it in place and no ghost fill is needed for that side
struct GsBC { double f0[3][6]; int even[3]; int odd[3]; };
int abec_gsrb(const Bx& bx, V4 phi, C4 rhs, const Abec& op, double omega, int redblack,
              int ncomp, cudaStream_t s, int wrapmask = 0, const GsBC* gb = nullptr, bool zero_phi = false);
// zero_phi: phi (ghost cells included) is identically zero before this pass: phi is not read, the cells of the other colour are
// set to zero (no setval needed before the first sweep of a multigrid correction)
// one full red-black sweep (colour rb0, then the other) phi_in -> phi_out (different arrays) on a box that spans
// the periodic domain in all directions with even extents; abec_gsrb_sweep_ok tells whether the box qualifies
bool abec_gsrb_sweep_ok(const Bx& bx, int wrapmask);
// all `nsweeps` red-black sweeps of a small box (coarse multigrid levels) in one single-CTA launch; only when every neighbour is
// reached inside the kernel (wrapmask / mirrored sides of gb cover all sides of the box)
bool abec_gsrb_small_ok(const Bx& bx, int ncomp);
int abec_gsrb_small(const Bx& bx, V4 phi, C4 rhs, const Abec& op, double omega, int ncomp, int nsweeps, bool zero_phi, cudaStream_t s,
                    int wrapmask, const GsBC* gb);
bool abec_gsrb_sweep_enabled();  // multigrid uses the fused sweep only when IAMRX_GSRB_FUSED=1 (see abec.cu)
int abec_gsrb_sweep(const Bx& bx, V4 phi_out, C4 phi_in, C4 rhs, const Abec& op, double omega, int rb0, int ncomp,
                    cudaStream_t s);
// out = L phi (rhs null) or rhs - L phi
int abec_apply(const Bx& bx, V4 out, C4 phi, C4 rhs, const Abec& op, int ncomp, cudaStream_t s,
               int wrapmask = 0, const GsBC* gb = nullptr,    // gb: in-place mirrored sides (homogeneous residuals)
               double* norm_dev = nullptr, bool* norm_fused = nullptr);
// norm_dev (device scalar, initialised by the caller, e.g. reduce_init op 2): *norm_dev = max(*norm_dev, max |out|) in the same
// launch when the fast kernel applies -- *norm_fused tells whether it did (otherwise reduce `out` afterwards)
// face flux_d = -b * beta_d * dphi/dx_d on faces of bx (MLABecLaplacian FFlux)
int abec_flux(const Bx& bx, V4 fx, V4 fy, V4 fz, C4 phi, const Abec& op, int comp, cudaStream_t s);
// crse = mean of 2x2x2 fine (MLCellLinOp restriction / average_down)
// thin (all transfer operators): bit d set = direction d has ratio 1 between the two levels (semi-coarsening of thin domains)
int cc_restrict(const Bx& cbx, V4 crse, C4 fine, int ncomp, cudaStream_t s, int thin = 0);
// fine += crse(i/2,j/2,k/2) (MLCellLinOp piecewise-constant interpolation)
int cc_prolong_add(const Bx& fbx, V4 fine, C4 crse, int ncomp, cudaStream_t s, int thin = 0);
// crse face coefficient = mean of the 4 fine faces (average_down_faces)
int face_restrict(const Bx& cfbx, int dir, V4 crse, C4 fine, int ncomp, cudaStream_t s, int thin = 0);
// beta_d = scale / (0.5*(rho(i-1)+rho(i))) (average_cellcenter_to_face + invert)
int rho_to_beta(const Bx& fbx, int dir, V4 beta, C4 rho, double scale, cudaStream_t s);
// div = fac * sum_d (u_d(i+1)-u_d(i))*dxinv[d]   (computeDivergence)
int mac_divergence(const Bx& bx, V4 div, C4 u, C4 v, C4 w, const double dxinv[3], double fac,
                   C4 minus_rhs, cudaStream_t s);
// u_d -= b*beta_d*(phi(i)-phi(i-1))*dxinv[d]   (umac += getFluxes)
int mac_update(const Bx& bx, V4 u, V4 v, V4 w, C4 phi, const Abec& op, cudaStream_t s);
// tensor cross terms: out += b*div(F_cross(eta, vel))
int tensor_cross(const Bx& bx, V4 out, C4 vel, C4 ex, C4 ey, C4 ez, double b,
                 const double dxinv[3], cudaStream_t s, int wrapmask = 0);   // wrapmask as for abec_apply: periodic images in-kernel

struct LinBC;
// the same on a box that may touch non-periodic domain faces: boundary-aware transverse derivatives (bv: level-BC fab, may be null)
int tensor_cross_bc(const Bx& bx, V4 out, C4 vel, C4 bv, C4 ex, C4 ey, C4 ez, double b, const double dxinv[3], const LinBC& bc,
                    const Bx& dom, const int per[3], cudaStream_t s);

// --- BLAS-1 style level ops (blas.cu) ------------------------------------
int setval(const Bx& bx, V4 dst, int ncomp, double val, cudaStream_t s);
int copy(const Bx& bx, V4 dst, C4 src, int ncomp, cudaStream_t s);
// dst = a*x + b*y  (x or y may alias dst)
int lincomb(const Bx& bx, V4 dst, double a, C4 x, double b, C4 y, int ncomp, cudaStream_t s);
// dst *= c*src  / dst = dst / src etc.
int mult(const Bx& bx, V4 dst, C4 src, int ncomp, int src_ncomp, cudaStream_t s);
int divide(const Bx& bx, V4 dst, C4 src, int ncomp, int src_ncomp, cudaStream_t s);
int scale(const Bx& bx, V4 dst, double c, int ncomp, cudaStream_t s);
int addconst(const Bx& bx, V4 dst, double c, int ncomp, cudaStream_t s);
// region copy with index shift: dst(i,j,k) = src(i+s0, j+s1, k+s2) over bx
int copy_shift(const Bx& bx, V4 dst, C4 src, int s0, int s1, int s2, int ncomp, cudaStream_t s);
// pack / unpack a region to a contiguous buffer (halo exchange)
int pack(const Bx& bx, double* buf, C4 src, int ncomp, cudaStream_t s);
int unpack(const Bx& bx, V4 dst, const double* buf, int ncomp, cudaStream_t s);
// reductions into device scalars: result[n] op= reduce over bx of comp n
// op: 0 sum, 1 min, 2 max|x| (norm0).  `result` must be initialised by the caller
// (reduce_init).  Deterministic order is NOT guaranteed for sum.
int reduce_init(double* result, int n, int op, cudaStream_t s);
int reduce(const Bx& bx, C4 src, int ncomp, int op, double* result, cudaStream_t s);
int reduce_dot(const Bx& bx, C4 x, C4 y, C4 mask, double* result, cudaStream_t s);
// constant-data detection (comp 0): see blas.cu; result = {differs, max first element, min first element}
int const_check(const Bx& bx, C4 src, double* result, cudaStream_t s);

// diagnostic: measured DFMA issue rate of the device in 1e9 instructions/s (thread-level; x2 = flop/s)
int fp64_peak(double* dp_ginstr_per_s, cudaStream_t s);

int godunov_set_option(int opt, double value);
double godunov_get_option(int opt);

// --- Godunov advection (godunov.cu) --------------------------------------
struct AdvGeom { double dx[3]; double dt; };
// physical boundaries seen by the Godunov kernels: the domain's cell bounds and the BCRec (IAMRX_BC_* codes) of every
// component; all-zero lo/hi (int_dir) = periodic / interior everywhere
struct AdvBC {
  int dlo[3], dhi[3];
  int lo[8][3], hi[8][3];
  bool interior() const {
    for (int n = 0; n < 8; ++n) for (int d = 0; d < 3; ++d) if (lo[n][d] != IAMRX_BC_INT_DIR || hi[n][d] != IAMRX_BC_INT_DIR) return false;
    return true;
  }
};
int extrap_vel_to_faces(const Bx& bx, C4 vel, C4 force, V4 umac, V4 vmac, V4 wmac,
                        const AdvGeom& g, int forces_in_trans, cudaStream_t s, int ppm = 0, const AdvBC* bc = nullptr);
struct AofsArgs {
  V4 aofs;           // ncomp
  C4 S, force, divu; // S: ncomp, 3 ghosts; force: ncomp 1 ghost (may be null); divu may be null
  C4 umac, vmac, wmac;
  C4 uflx, vflx, wflx;  // flux velocities (== umac.. unless sync)
  V4 fx, fy, fz, xed, yed, zed;  // optional outputs
  int ncomp;
  int iconserv[8];
  int forces_in_trans, is_velocity, is_sync, write_fluxes;
  int staged = 0;  // force the staged kernels
  int ppm = 0;     // Godunov_PPM (staged kernels only)
  int known_edge_state = 0;   // xed/yed/zed are INPUTS: only fluxes, divergence and the convective term are formed
  AdvBC bc{};      // zero-initialised = interior
};
int compute_aofs(const Bx& bx, const AofsArgs& a, const AdvGeom& g, cudaStream_t s);

// --- nodal Laplacian (nodal.cu) ------------------------------------------
// hide: bit 2d / 2d+1 = the low / high side of nbx in direction d is a Neumann / inflow domain side (tangential ghost velocities unseen)
int nodal_divu(const Bx& nbx, V4 rhs, C4 vel, const double dxinv[3], cudaStream_t s, int hide = 0);
int nodal_adotx(const Bx& nbx, V4 out, C4 phi, C4 rhs, C4 sig, const double dxinv[3], cudaStream_t s,
                int wrapmask = 0, double* norm_dev = nullptr, bool* norm_fused = nullptr);   // norm_dev: as in abec_apply
// all sweeps (8 colours each, in place) of a small single-box level in one single-CTA launch (see nodal.cu)
bool nodal_gs_small_ok(const Bx& nbx);
int nodal_gs_small(const Bx& nbx, V4 phi, C4 rhs, C4 sig, const double dxinv[3], int nsweeps, cudaStream_t s, int wrapmask);
int nodal_gs_color(const Bx& nbx, V4 phi, C4 rhs, C4 sig, const double dxinv[3], int color,
                   cudaStream_t s, int wrapmask = 0);
// full 8-colour Gauss-Seidel sweep phi_in -> phi_out on a box that spans the periodic domain
// (one fused launch per plane parity); nodal_gs_sweep_ok tells whether the box qualifies
bool nodal_gs_sweep_ok(const Bx& nbx, int wrapmask);
// wrapmask flag of nodal_gs_sweep[_ok]: the x / y directions that are NOT wrapped have neighbouring boxes (same level or periodic
// images) and phi_in, phi_out, rhs carry 4 filled ghost nodes, sigma 4 filled ghost cells there: the tiles take their
// halos (and recompute the colour dependences) from the ghost layers, so a block-decomposed level smooths with two
// launches + two ghost exchanges per sweep like a slab instead of eight colour launches + eight exchanges
constexpr int NODAL_DEEP_GHOSTS = 512;
// wrapmask 7 (box spans the periodic domain) or 3 (slab: x, y wrapped, z neighbours read from filled ghost planes).
// phase 0 = even planes (colours 0-3), 1 = odd planes (colours 4-7), -1 = both; with wrapmask 3 the caller exchanges
// the z ghost planes of phi_out between the phases.
int nodal_gs_sweep(const Bx& nbx, V4 phi_out, C4 phi_in, C4 rhs, C4 sig, const double dxinv[3], cudaStream_t s,
                   int wrapmask = 7, int phase = -1);
int nodal_jacobi(const Bx& nbx, V4 out, C4 phi, C4 rhs, C4 sig, const double dxinv[3], double omega,
                 cudaStream_t s);
// wm + fnb: wrap the stencil in the directions (bits) in which the fine node box fnb spans the periodic domain
int nodal_restrict(const Bx& cnbx, V4 crse, C4 fine, cudaStream_t s, int thin = 0, int wm = 0, const Bx* fnb = nullptr);
int nodal_interp_add(const Bx& fnbx, V4 fine, C4 crse, cudaStream_t s, int thin = 0);
int nodal_mknewu(const Bx& bx, V4 vel, V4 gp, int increment_gp, C4 phi, C4 sig,
                 const double dxinv[3], cudaStream_t s);

// --- physical domain boundaries (bc.cu) -----------------------------------
// BCRec (IAMRX_BC_* codes) and ext_dir face values of up to 8 state components: what AmrLevel::FillPatch applies outside the domain
struct PhysBC {
  int lo[8][3], hi[8][3];
  double val[6][8];   // [x lo, y lo, z lo, x hi, y hi, z hi][comp]  (NavierStokes::get_bc_values, NS.cpp:108-237)
};
// fill the cells of `fabbox` (the fab's allocated region) outside the non-periodic sides of the domain
int fill_physbc(const Bx& fabbox, V4 a, int ncomp, const PhysBC& bc, const Bx& dom, const int per[3], cudaStream_t s);
// LinOpBCType of every side for up to 3 components (MLTensorOp has one set per velocity component) + setMaxOrder
struct LinBC {
  int lo[3][3], hi[3][3];   // [comp][dir]
  int maxorder;
  bool any_nonperiodic() const { for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) if (lo[c][d] != IAMRX_LINOP_PERIODIC || hi[c][d] != IAMRX_LINOP_PERIODIC) return true; return false; }
};
int linop_bc_order(int maxorder, int boxlen);
double linop_bc_f0(int code, int maxorder, int boxlen);   // coefficient of the first interior cell in the ghost formula
// MLCellLinOp::applyBC on the domain faces of box vbx: ghost layer of phi beyond every non-periodic side the box touches.
// bv: fab whose ghost cells hold the Dirichlet face values (null: homogeneous).  grow_t: extra transverse cells (the tensor
// cross terms read one cell sideways).  skipmask bit d: leave direction d alone.
int linop_bc_fill(const Bx& vbx, V4 phi, int ncomp, const LinBC& bc, C4 bv, const Bx& dom, const int per[3], int grow_t, int skipmask,
                  cudaStream_t s);
// coarse-fine sides of a fine AMR level (bc.cu): Dirichlet data x0 cell widths from the face (x0 < 0: beyond it)
double linop_cf_f0(int maxorder, int boxlen, double x0);
int linop_cf_fill(const Bx& vbx, V4 phi, int ncomp, int maxorder, C4 bv, int cfmask, const double x0[3], cudaStream_t s);
int mask_zero(const Bx& R, V4 a, C4 m, cudaStream_t s);   // a = 0 where m != 0
struct NodalBC { int lo[3], hi[3]; };   // LinOpBCType per side (Projection.cpp:2436-2464)
int nodal_bc_fill_phi(const Bx& nbx, V4 phi, const NodalBC& bc, const Bx& ndom, const int per[3], int skipmask, cudaStream_t s);
int nodal_bc_fill_sigma(const Bx& cbx, V4 sig, const NodalBC& bc, const Bx& dom, const int per[3], cudaStream_t s, int ngt = 1);
int nodal_bc_scale(const Bx& nbx, V4 a, const NodalBC& bc, const Bx& ndom, const int per[3], double f, cudaStream_t s);

// --- two-level transfer operators and flux register pieces (amr.cu) ---------
int average_down_nodal(const Bx& cnbx, V4 crse, C4 fine, int ncomp, cudaStream_t s);
int cell_cons_interp(const Bx& fbx, V4 fine, C4 crse, int ncomp, cudaStream_t s);
int pc_interp(const Bx& fbx, V4 fine, C4 crse, int ncomp, cudaStream_t s);   // piecewise constant
// InterpBndryData (order 3): coarse-fine boundary values in the ghost layer R beyond side d of a fine box (mask: usable coarse cells)
int cf_bndry_interp(const Bx& R, int d, V4 fine, C4 crse, C4 mask, int ncomp, cudaStream_t s);
// SyncRegister pieces (SyncRegister.cpp): InitRHS's interior mask, FineAdd's restriction of one boundary plane, the periodic gather
int sync_mask(const Bx& nbx, V4 mask, C4 covered, double maxcount, cudaStream_t s);
int sync_fine_add(const Bx& R, V4 acc, C4 fine, const Bx& fnb, int dir, double mult, const Bx& cnd, const int per[3], cudaStream_t s);
int sync_gather(const Bx& nbx, V4 reg, C4 onb, C4 acc, const Bx& cnd, const int plen[3], cudaStream_t s);
// create_umac_grown's divergence correction on the one-cell halo of a fine box (mask: 0 interior, 1 covered, 2 not covered, 3 physbnd)
int umac_divfix(const Bx& vb, C4 mask, V4 u, V4 v, V4 w, C4 divu, const double dx[3], cudaStream_t s);
int node_bilinear_interp(const Bx& fnbx, V4 fine, C4 crse, int ncomp, cudaStream_t s);
int face_linear_interp(const Bx& ffbx, int dir, V4 fine, C4 crse, int ncomp, cudaStream_t s);

// --- pointwise IAMR glue (pointwise.cu) -----------------------------------
// floor NSB.cpp:4530-4534: |v| > 1e-20 ? v : 0
int floor_small(const Bx& bx, V4 f, int ncomp, cudaStream_t s);
// velocity forcing: tf_n = getForce_n + visc_n - gp_n, optionally / rho
// (NSB.cpp:4456-4470, 3445-3466, 1411-1424); getForce = NS_getForce.cpp:117-141
// (buoyancy grav*rho on the z component when |grav| > 1e-4).  visc / gp may be null.
// uf: optional user force per unit mass (the HIT turbulent forcing, 3 comps): f += rho * uf
int force_vel(const Bx& bx, V4 tf, C4 visc, C4 gp, C4 rho, double grav, int div_rho, cudaStream_t s, C4 uf = C4{});

// ---- HIT turbulent forcing (forcing.cu; Tutorials/HIT/NS_getForce.cpp:205-640) -----------------------------------------
struct TurbParams { int nmodes, mode_start, div_free, array_size; };
struct TurbMode { double w[3]; double ph[3][3]; double a[3]; };   // 2 pi k_d / L_d; phases [component][direction]; xT * amplitudes
}  // namespace k
}  // namespace ix
#include <vector>
namespace ix {
namespace k {
int turb_modes(const TurbParams& tp, const double* forcedata, const double L[3], double time, std::vector<TurbMode>& out);
size_t turb_scratch_doubles(const Bx& bx, int nm);
// frc (3 comps) = or += rho * f(x, t) on bx; d_modes / d_tab: device copies of the mode list / table scratch
int turb_force(const Bx& bx, V4 frc, C4 rho, const iamrx_geom& g, const TurbMode* d_modes, int nm, int div_free, double* d_tab,
               int accumulate, cudaStream_t s);
// velocity update NSB.cpp:3607-3626; do_mom_diff when rho_old / rho_new are given (u_new = (rho_old u_old - dt aofs + dt f - dt gp) / rho_new):
//   unew = uold - dt*aofs + dt*(getForce(rho_half) - gp)/rho_half
int vel_update(const Bx& bx, V4 unew, C4 uold, C4 aofs, C4 gp, C4 rhohalf, double grav, double dt,
               int zero_force, cudaStream_t s, C4 rho_old = C4{}, C4 rho_new = C4{}, C4 uf = C4{});
// scalar update NSB.cpp:2761-2765 / 2887-2896 with the default (zero) scalar forcing
int scal_update(const Bx& bx, V4 snew, C4 sold, C4 aofs, double dt, int ncomp, cudaStream_t s);
// ns.do_scalminmax: Conservative / ConvectiveScalMinMax (NSB.cpp:2907-2935, 4256-4370); sold / rhoold need 1 filled ghost cell
int scal_minmax(const Bx& bx, V4 snew, C4 rhonew, C4 sold, C4 rhoold, int conservative, cudaStream_t s);
// diffusion rhs: unew *= rho; rhs += unew (Diffusion.cpp:821-831)
int diff_rhs(const Bx& bx, V4 rhs, V4 unew, C4 rho, int ncomp, cudaStream_t s);
// level_project pre: u = u*dt_inv + gp/rho  (Projection.cpp:273,296-300)
int proj_pre(const Bx& bx, V4 u, C4 gp, C4 rho, double dt_inv, cudaStream_t s);
// Projection::computeRhoG (Projection.cpp:1933-2379) on the part `strip` of an outflow face's node plane: d = direction of the face,
// t = the other horizontal direction, c1 / c2 = the first / second cell layer inside the face, tlo / thi = the transverse node range
// of the domain, code_lo / code_hi = the density's BCRec on the transverse sides, ztop = the top cell layer of the domain
struct OutflowRhoG { int d, t, c1, c2, tlo, thi, code_lo, code_hi, ztop; double gravity, dz; };
int outflow_rhog(const Bx& strip, V4 phi, C4 rho, const OutflowRhoG& a, cudaStream_t s);
// scaleVar: sig = 1/rho (Projection.cpp:1327-1349)
int invert(const Bx& bx, V4 sig, C4 rho, cudaStream_t s);
// prob_init.cpp initial conditions (state: u,v,w,rho,tracer) on bx
int init_prob(const Bx& bx, V4 state, int probtype, const double* params, int nparams,
              const iamrx_geom& g, cudaStream_t s);

}  // namespace k
}  // namespace ix
