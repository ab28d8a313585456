// adapter_sites.cpp -- the call-site replacements of INTEGRATION.md as compilable functions.  Each function takes the
// arguments the IAMR wrapper has in scope at that site (names as in the reference) and forwards them through
// include/IamrxAdapter.H to the C ABI.  Compiled against tests/abi/amrex_stub (this container has no AMReX) by
// tests/test_abi.py::test_integration_snippets_compile, which also checks that every [site] block below appears verbatim
// in INTEGRATION.md -- the document cannot drift from code that compiles against include/iamrx.h.
#define IAMRX_ADAPTER_AMREX_STUB 1
#include <map>
#include "IamrxAdapter.H"

using namespace amrex;

// ---- NavierStokesBase.cpp:4487-4491 --------------------------------------------------------------------------------------
void site1_extrap_vel_to_faces(MultiFab const& Umf, MultiFab const& forcing_term, MultiFab* u_mac, Vector<BCRec> const& h_bcrec,
                               Geometry const& geom, Real dt, bool godunov_use_ppm, bool godunov_use_forces_in_trans) {
// [site 1 begin]
// replaces Godunov::ExtrapVelToFaces(Umf, forcing_term, u_mac[0], u_mac[1], u_mac[2], h_bcrec, d_bcrec, geom, dt, ppm, fit)
auto g = iamrx_adapt::geom(geom);
auto bc = iamrx_adapt::bcrecs(h_bcrec);                     // the three velocity components
int flags = (godunov_use_ppm ? IAMRX_ADV_PPM : 0) | (godunov_use_forces_in_trans ? IAMRX_ADV_FORCES_IN_TRANS : 0);
for (MFIter mfi(Umf); mfi.isValid(); ++mfi) {
  auto bx = iamrx_adapt::box(mfi.validbox());
  auto U = iamrx_adapt::fab(Umf.const_array(mfi)), F = iamrx_adapt::fab(forcing_term.const_array(mfi));
  auto um = iamrx_adapt::fab(u_mac[0].array(mfi)), vm = iamrx_adapt::fab(u_mac[1].array(mfi)), wm = iamrx_adapt::fab(u_mac[2].array(mfi));
  iamrx_adapt::check(iamrx_extrap_vel_to_faces_box(&bx, &U, &F, &um, &vm, &wm, bc.data(), &g, dt, flags, Gpu::gpuStream()));
}
// [site 1 end]
}

// ---- NavierStokesBase.cpp:4661-4845 (MFIter loop body of ComputeAofs) -------------------------------------------------------
void site2_compute_aofs(MultiFab& advc, int a_comp, MultiFab const& S, int s_comp, int ncomp, MultiFab const& forcing, int f_comp,
                        MultiFab const* divu, MultiFab const* u_mac, MultiFab const* U_corr, MultiFab* cfluxes, MultiFab* edgestate,
                        Vector<int> const& iconserv_h, Vector<BCRec> const& bcrec_h, Geometry const& geom, Real dt,
                        bool is_velocity, bool is_sync, bool known_edgestate, bool godunov_use_ppm, bool godunov_use_forces_in_trans) {
// [site 2 begin]
// replaces steps 1-3 of the loop body: ComputeFluxesOnBoxFromState (:4701-4717), ComputeDivergence (:4768-4771),
// div(u_mac) (:4809-4810), ComputeConvectiveTerm (:4813-4820) and the sign / sync accumulation (:4834-4843)
auto g = iamrx_adapt::geom(geom);
auto bc = iamrx_adapt::bcrecs(bcrec_h);                     // fetchBCArray(State_Type, S_comp, ncomp) :4645
int flags = (godunov_use_ppm ? IAMRX_ADV_PPM : 0) | (godunov_use_forces_in_trans ? IAMRX_ADV_FORCES_IN_TRANS : 0)
          | (is_velocity ? IAMRX_ADV_IS_VELOCITY : 0) | (is_sync ? IAMRX_ADV_IS_SYNC : 0)
          | (known_edgestate ? IAMRX_ADV_KNOWN_EDGE_STATE : 0)
          | IAMRX_ADV_WRITE_FLUXES;                         // cfluxes / edgestate feed the flux registers (:4848-4889)
for (MFIter mfi(advc); mfi.isValid(); ++mfi) {
  auto bx = iamrx_adapt::box(mfi.validbox());
  auto A = iamrx_adapt::fab(advc.array(mfi)), Sf = iamrx_adapt::fab(S.const_array(mfi)), Ff = iamrx_adapt::fab(forcing.const_array(mfi));
  iamrx_fab Dv{}, uf[3], um[3], fl[3], ed[3];
  if (divu) Dv = iamrx_adapt::fab(divu->const_array(mfi));
  for (int d = 0; d < 3; ++d) {
    um[d] = iamrx_adapt::fab(u_mac[d].const_array(mfi));
    if (is_sync) uf[d] = iamrx_adapt::fab(U_corr[d].const_array(mfi));   // fluxes use U_corr, edge states u_mac (:4672-4677)
    fl[d] = iamrx_adapt::fab(cfluxes[d].array(mfi));
    ed[d] = iamrx_adapt::fab(edgestate[d].array(mfi));
  }
  iamrx_adapt::check(iamrx_compute_aofs_box(&bx, &A, a_comp, &Sf, s_comp, ncomp, &Ff, f_comp, divu ? &Dv : nullptr,
      &um[0], &um[1], &um[2], is_sync ? &uf[0] : nullptr, is_sync ? &uf[1] : nullptr, is_sync ? &uf[2] : nullptr,
      &fl[0], &fl[1], &fl[2], &ed[0], &ed[1], &ed[2], iconserv_h.data(), bc.data(), &g, dt, flags, Gpu::gpuStream()));
}
// [site 2 end]
}

// ---- MacProj.cpp:1084-1184 ----------------------------------------------------------------------------------------------------
void site3_mlmg_mac_solve(int level, Geometry const& geom, BoxArray const& ba, DistributionMapping const& dm, MultiFab* const* u_mac,
                          MultiFab const* cphi, MultiFab const& rho_half, MultiFab const* Rhs, MultiFab* mac_phi, MultiFab* const* fluxes, Real rhs_scale,
                          Real a_mac_tol, Real a_mac_abs_tol, int max_order, int verbose,
                          Array<LinOpBCType, 3> const& mlmg_lobc, Array<LinOpBCType, 3> const& mlmg_hibc) {
// [site 3 begin]
// replaces the bcoefs build (:1115-1127) + Hydro::MacProjector {ctor, setDomainBC, setLevelBC, setMaxOrder, project, getFluxes}
static std::map<int, iamrx_level_t> lev_cache;              // per AMR level; erase the entry on regrid
auto L = lev_cache.count(level) ? lev_cache[level] : (lev_cache[level] = iamrx_adapt::level(geom, ba, dm));
auto um = iamrx_adapt::fabs(*u_mac[0]), vm = iamrx_adapt::fabs(*u_mac[1]), wm = iamrx_adapt::fabs(*u_mac[2]);
auto rho = iamrx_adapt::fabs(rho_half), phi = iamrx_adapt::fabs(*mac_phi);
std::vector<iamrx_fab> rhs; if (Rhs) rhs = iamrx_adapt::fabs(*Rhs);
int lobc[3], hibc[3];
iamrx_adapt::linop_bc(mlmg_lobc, mlmg_hibc, lobc, hibc);   // as set_mac_solve_bc filled them (:1187-1208)
iamrx_mg_info info; iamrx_mg_info_default(&info);
info.rtol = a_mac_tol; info.atol = a_mac_abs_tol; info.maxorder = max_order; info.verbose = verbose;
if (level > 0 && cphi) {                                    // macproj.setCoarseFineBC(cphi, ratio) (:1164-1167)
  auto cp = iamrx_adapt::fabs(*cphi);                       // -> coarse-fine boundary values in the ghost cells of mac_phi
  iamrx_adapt::check(iamrx_set_coarse_fine_bc(L, lev_cache.at(level - 1), phi.data(), cp.data(), 1, Gpu::gpuStream()));
}
int rc = iamrx_mac_project(L, um.data(), vm.data(), wm.data(), rho.data(), Rhs ? rhs.data() : nullptr, phi.data(), rhs_scale,
                           lobc, hibc, &info, Gpu::gpuStream());
if (rc > 0) Abort("MLMG failed to converge");               // reference behaviour
iamrx_adapt::check(rc);
if (fluxes) {                                               // macproj.getFluxes (:1181-1183), mac_sync_solve only
  auto fx = iamrx_adapt::fabs(*fluxes[0]), fy = iamrx_adapt::fabs(*fluxes[1]), fz = iamrx_adapt::fabs(*fluxes[2]);
  iamrx_adapt::check(iamrx_mac_get_fluxes(L, fx.data(), fy.data(), fz.data(), phi.data(), Gpu::gpuStream()));
}
// [site 3 end]
}

// ---- Projection.cpp:2385-2567 ---------------------------------------------------------------------------------------------------
void site4_nodal_projection(iamrx_level_t L, MultiFab* vel_rebase, MultiFab* sigma_rebase, MultiFab* phi_rebase, MultiFab* Gp,
                            Real rel_tol, Real abs_tol, bool increment_gp,
                            Array<LinOpBCType, 3> const& mlmg_lobc, Array<LinOpBCType, 3> const& mlmg_hibc) {
// [site 4 begin]
// replaces Hydro::NodalProjector {ctor, setDomainBC, getLinOp().setGaussSeidel/HarmonicAverage, project, getGradPhi} (:2512-2567)
auto vel = iamrx_adapt::fabs(*vel_rebase), sig = iamrx_adapt::fabs(*sigma_rebase);
auto phi = iamrx_adapt::fabs(*phi_rebase), gp = iamrx_adapt::fabs(*Gp);
int lobc[3], hibc[3];
iamrx_adapt::linop_bc(mlmg_lobc, mlmg_hibc, lobc, hibc);   // set_boundary_velocity + BC translation (:2436-2464)
iamrx_mg_info info; iamrx_mg_info_default(&info); info.rtol = rel_tol; info.atol = abs_tol;
int rc = iamrx_nodal_project(L, vel.data(), sig.data(), phi.data(), gp.data(), increment_gp ? 1 : 0, lobc, hibc, &info, Gpu::gpuStream());
if (rc > 0) Abort("MLMG failed to converge");
iamrx_adapt::check(rc);
// [site 4 end]
}

// ---- Diffusion.cpp:715-768, 858-923, 1708-1757 --------------------------------------------------------------------------------
void site5_tensor_diffusion(iamrx_level_t L, MultiFab& visc, MultiFab& U, MultiFab& Soln, MultiFab const& Rhs, MultiFab const& rho_half,
                            MultiFab* const* eta, Vector<BCRec> const& velbc, Real be_cn_theta, Real dt, Real visc_tol, Real abs_tol) {
// [site 5 begin]
// Diffusion::setDomainBC for the three velocity components (:1887-1999, :711-724), max_order 2 (:95-96)
iamrx_linop_bc bc{}; bc.maxorder = 2;
for (int c = 0; c < 3; ++c)
  for (int d = 0; d < 3; ++d) {
    auto tr = [](int b) { return b == IAMRX_BC_INT_DIR ? IAMRX_LINOP_PERIODIC : b == IAMRX_BC_EXT_DIR ? IAMRX_LINOP_DIRICHLET
                               : b == IAMRX_BC_REFLECT_ODD ? IAMRX_LINOP_REFLECT_ODD : IAMRX_LINOP_NEUMANN; };
    bc.lo[c][d] = tr(velbc[c].lo(d)); bc.hi[c][d] = tr(velbc[c].hi(d));
  }
auto ex = iamrx_adapt::fabs(*eta[0]), ey = iamrx_adapt::fabs(*eta[1]), ez = iamrx_adapt::fabs(*eta[2]);
// MLTensorOp apply (getTensorViscTerms :1708-1757): visc = div(eta (grad U + grad U^T)); a = 0, b = -1 in the reference's scaling
auto vt = iamrx_adapt::fabs(visc), Uf = iamrx_adapt::fabs(U);
iamrx_adapt::check(iamrx_diffusion_apply(L, /*tensor*/ 1, 3, vt.data(), Uf.data(), 0.0, -1.0, nullptr, ex.data(), ey.data(), ez.data(),
                                         &bc, Gpu::gpuStream()));
// diffuse_tensor_velocity solve (:858-923): (rho_half - theta dt div eta (grad + grad^T)) Soln = Rhs; Soln's ghost cells = level BC
auto so = iamrx_adapt::fabs(Soln), rh = iamrx_adapt::fabs(Rhs), ac = iamrx_adapt::fabs(rho_half);
iamrx_mg_info info; iamrx_mg_info_default(&info); info.rtol = visc_tol; info.atol = abs_tol;
int rc = iamrx_diffusion_solve(L, 1, 3, so.data(), rh.data(), 1.0, be_cn_theta * dt, ac.data(), ex.data(), ey.data(), ez.data(),
                               &bc, &info, Gpu::gpuStream());
if (rc > 0) Abort("MLMG failed to converge");
iamrx_adapt::check(rc);
// [site 5 end]
}

// ---- AmrLevel::FillPatch physical fill (NS_bcfill.H:17-167) + two-level transfer (NSB.cpp:4125-4191, 4848-4889) --------------
void site6_fill_and_transfer(iamrx_level_t Lc, iamrx_level_t Lf, MultiFab& S_fine, MultiFab& S_crse, MultiFab& S_crse_old,
                             Real t_crse_old, Real t_crse_new, Real time, Vector<BCRec> const& bcs,
                             const double* bc_values, MultiFab* const* fine_fluxes, MultiFab* const* crse_fluxes, Real dt_fine,
                             Real dt_crse, int ncomp, Geometry const& crse_geom) {
// [site 6 begin]
auto sf = iamrx_adapt::fabs(S_fine), sc = iamrx_adapt::fabs(S_crse);
auto bc = iamrx_adapt::bcrecs(bcs);
// same-level ghost exchange, then the physical-boundary fill of FillPatch (ext_dir values ON the face)
iamrx_adapt::check(iamrx_fill_boundary(Lf, sf.data(), IAMRX_IX_CELL, ncomp, S_fine.nGrow(), Gpu::gpuStream()));
iamrx_adapt::check(iamrx_fill_physbc(Lf, sf.data(), ncomp, S_fine.nGrow(), bc.data(), bc_values, Gpu::gpuStream()));
// on a level that does not cover the domain: FillPatchTwoLevels (coarse data linear in time between the two coarse states)
auto so = iamrx_adapt::fabs(S_crse_old);
iamrx_adapt::check(iamrx_fillpatch_two_levels(Lf, Lc, sf.data(), so.data(), sc.data(), t_crse_old, t_crse_new, time, ncomp, S_fine.nGrow(),
                                              bc.data(), bc_values, Gpu::gpuStream()));
// advective flux register: CrseInit / FineAdd during the two advances, Reflux afterwards (NS.cpp:1713-1838)
static iamrx_fluxreg_t reg = nullptr;
if (!reg) iamrx_adapt::check(iamrx_fluxreg_create(Lc, Lf, ncomp, &reg));
auto cx = iamrx_adapt::fabs(*crse_fluxes[0]), cy = iamrx_adapt::fabs(*crse_fluxes[1]), cz = iamrx_adapt::fabs(*crse_fluxes[2]);
auto fx = iamrx_adapt::fabs(*fine_fluxes[0]), fy = iamrx_adapt::fabs(*fine_fluxes[1]), fz = iamrx_adapt::fabs(*fine_fluxes[2]);
const double vol_crse = crse_geom.CellSize(0) * crse_geom.CellSize(1) * crse_geom.CellSize(2);   // "dx := volume" (NSB.cpp:4878-4889)
iamrx_adapt::check(iamrx_fluxreg_reset(reg, Gpu::gpuStream()));
iamrx_adapt::check(iamrx_fluxreg_crse_add(reg, cx.data(), cy.data(), cz.data(), dt_crse, vol_crse, Gpu::gpuStream()));
iamrx_adapt::check(iamrx_fluxreg_fine_add(reg, fx.data(), fy.data(), fz.data(), dt_fine, vol_crse, Gpu::gpuStream()));
iamrx_adapt::check(iamrx_fluxreg_reflux(reg, sc.data(), 0, 1.0, Gpu::gpuStream()));
// [site 6 end]
}

int main() { return 0; }
