// solvers.cu -- see solvers.h.
#include "solvers.h"

namespace ix {

#define IX_TRY(call) do { int rc_ = (call); if (rc_ != IAMRX_OK) return rc_; } while (0)

static iamrx_mg_info info_or_default(const iamrx_mg_info* info) {
  iamrx_mg_info d;
  iamrx_mg_info_default(&d);
  return info ? *info : d;
}

// MacProj::mlmg_mac_solve (MacProj.cpp:1084-1184) + Hydro::MacProjector::project:
//   beta_d = (1/rhs_scale) / avg(rho)         MacProj.cpp:1115-1127
//   -div(beta grad phi) = -(div(umac) - rhs)  (MLABecLaplacian, a=0, b=1)
//   umac -= beta grad phi                      (umac += getFluxes)
k::LinBC periodic_linbc() {
  k::LinBC b;
  for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) { b.lo[c][d] = IAMRX_LINOP_PERIODIC; b.hi[c][d] = IAMRX_LINOP_PERIODIC; }
  b.maxorder = 3;
  return b;
}

int mac_project(Level& L, LevelSolvers& sv, MF U[3], const MF& rho, const MF* rhs, MF& phi,
                double rhs_scale, iamrx_mg_info* info, cudaStream_t s, const k::LinBC* bc) {
  iamrx_mg_info mi = info_or_default(info);
  if (!sv.mac || sv.mac_mc != mi.max_coarsening) {
    sv.mac = std::make_unique<CellMG>(&L, 1, false, mi.max_coarsening);
    sv.mac_mc = mi.max_coarsening;
    for (int d = 0; d < 3; ++d) sv.mac_beta[d].define(&L, IX_XFACE + d, 1, 0);
    sv.mac_rhs.define(&L, IX_CELL, 1, 0);
  }
  for (int d = 0; d < 3; ++d)
    for (int il = 0; il < rho.n(); ++il)
      IX_TRY(k::rho_to_beta(sv.mac_beta[d].vbox(il), d, sv.mac_beta[d].v(il), rho.c(il), 1.0 / rhs_scale, s));
  sv.mac->set_scalars(0.0, 1.0);
  sv.mac->set_bc(bc ? *bc : periodic_linbc());
  IX_TRY(sv.mac->set_coeffs(nullptr, &sv.mac_beta[0], &sv.mac_beta[1], &sv.mac_beta[2], s));
  for (int il = 0; il < phi.n(); ++il)
    IX_TRY(k::mac_divergence(sv.mac_rhs.vbox(il), sv.mac_rhs.v(il), U[0].c(il), U[1].c(il), U[2].c(il),
                             L.dxinv, -1.0, rhs ? rhs->c(il) : C4{}, s));
  const int rc = sv.mac->solve(phi, sv.mac_rhs, &mi, s);
  if (info) *info = mi;
  if (rc < 0) return rc;
  for (int il = 0; il < phi.n(); ++il)
    IX_TRY(k::mac_update(L.lbox(il), U[0].v(il), U[1].v(il), U[2].v(il), phi.c(il), sv.mac->op_at(0, il), s));
  return rc;
}

// Hydro::MacProjector::getFluxes (MacProj.cpp:1181-1183): F_d = -beta_d dphi/dx_d on the faces of every local box, with the
// operator (beta) of the last mac_project on this level.  mac_sync_solve turns them into U_corr (MacProj.cpp:459-468).
int mac_get_fluxes(Level& L, LevelSolvers& sv, MF F[3], MF& phi, cudaStream_t s) {
  if (!sv.mac) { set_error("mac_get_fluxes: no MAC projection has been done on this level"); return IAMRX_ERR_ARG; }
  IX_TRY(mf_fill_boundary(phi, 0, 1, 1, s));
  for (int il = 0; il < phi.n(); ++il)
    IX_TRY(k::abec_flux(L.lbox(il), F[0].v(il), F[1].v(il), F[2].v(il), phi.c(il), sv.mac->op_at(0, il), 0, s));
  return IAMRX_OK;
}

// Projection::doMLMGNodalProjection (Projection.cpp:2385-2567) + Hydro::NodalProjector:
//   rhs = FE divergence of vel on nodes; div(sigma grad phi) = rhs;
//   vel -= sigma grad phi; gp (+)= grad phi
int nodal_project(Level& L, LevelSolvers& sv, MF& vel, const MF& sigma, MF& phi, MF* gp, int increment_gp,
                  iamrx_mg_info* info, cudaStream_t s, const k::NodalBC* bc, bool keep_dirichlet) {
  iamrx_mg_info mi = info_or_default(info);
  if (!sv.nodal || sv.nodal_mc != mi.max_coarsening) {
    sv.nodal = std::make_unique<NodeMG>(&L, mi.max_coarsening);
    sv.nodal_mc = mi.max_coarsening;
    sv.nodal_rhs.define(&L, IX_NODE, 1, 1);
  }
  NodeMG& mg = *sv.nodal;
  k::NodalBC pb; for (int d = 0; d < 3; ++d) { pb.lo[d] = IAMRX_LINOP_PERIODIC; pb.hi[d] = IAMRX_LINOP_PERIODIC; }
  mg.set_bc(bc ? *bc : pb);
  IX_TRY(mg.set_sigma(sigma, s));
  IX_TRY(mf_fill_boundary(vel, 0, 3, 1, s));
  if (mg.has_bc()) {
    // Projection::set_boundary_velocity (Projection.cpp:2570-2663): the normal velocity in the ghost cells beyond a
    // non-periodic side is zeroed unless that side is an inflow face (whose ghost cells carry the inflow velocity)
    for (int il = 0; il < vel.n(); ++il)
      for (int d = 0; d < 3; ++d) {
        if (L.geom.periodic[d]) continue;
        for (int side = -1; side <= 1; side += 2) {
          if ((side < 0 ? mg.bc().lo[d] : mg.bc().hi[d]) == IAMRX_LINOP_INFLOW) continue;
          if (side < 0 ? L.lbox(il).lo[d] != L.domain.lo[d] : L.lbox(il).hi[d] != L.domain.hi[d]) continue;
          Bx R = grow(L.lbox(il), 1);
          R.lo[d] = R.hi[d] = side < 0 ? L.domain.lo[d] - 1 : L.domain.hi[d] + 1;
          IX_TRY(k::setval(R, vel.v(il, d), 1, 0.0, s));
        }
      }
    IX_TRY(mf_setval(sv.nodal_rhs, 0.0, 0, 1, 1, s));   // rows ON Dirichlet sides stay zero
  }
  const Bx ndom = ixbox(L.domain, IX_NODE);
  for (int il = 0; il < vel.n(); ++il) {
    IX_TRY(k::nodal_divu(mg.active_nbox(0, il), sv.nodal_rhs.v(il), vel.c(il), L.dxinv, s, mg.neumann_sides(0, il)));
    // mlndlap_impose_neumann_bc: rows ON Neumann / inflow sides are doubled, once per direction
    if (mg.has_bc()) IX_TRY(k::nodal_bc_scale(sv.nodal_rhs.vbox(il), sv.nodal_rhs.v(il), mg.bc(), ndom, L.geom.periodic, 2.0, s));
  }
  IX_TRY(mg.apply_node_mask(0, sv.nodal_rhs, s));   // a fine level of general shape: no equation on its coarse-fine boundary nodes
  if (mg.has_bc() && !keep_dirichlet) {   // nodes ON Dirichlet sides are held at zero
    for (int il = 0; il < phi.n(); ++il) {
      const Bx full = phi.vbox(il), act = mg.active_nbox(0, il, /*with_cf=*/false);   // coarse-fine boundary nodes keep their values
      for (int d = 0; d < 3; ++d) {
        if (act.lo[d] > full.lo[d]) { Bx R = full; R.hi[d] = full.lo[d]; IX_TRY(k::setval(R, phi.v(il), 1, 0.0, s)); }
        if (act.hi[d] < full.hi[d]) { Bx R = full; R.lo[d] = full.hi[d]; IX_TRY(k::setval(R, phi.v(il), 1, 0.0, s)); }
      }
    }
  }
  const int rc = mg.solve(phi, sv.nodal_rhs, &mi, s);
  if (info) *info = mi;
  if (rc < 0) return rc;
  const MF& sig = mg.sigma0();
  for (int il = 0; il < vel.n(); ++il)
    IX_TRY(k::nodal_mknewu(L.lbox(il), vel.v(il), gp ? gp->v(il) : V4{}, increment_gp, phi.c(il), sig.c(il),
                           L.dxinv, s));
  return rc;
}

int comp_grad(Level& L, MF& gp, MF& p, cudaStream_t s) {
  IX_TRY(mf_fill_boundary(p, 0, 1, 1, s));
  for (int il = 0; il < gp.n(); ++il)
    IX_TRY(k::nodal_mknewu(L.lbox(il), V4{}, gp.v(il), 0, p.c(il), C4{}, L.dxinv, s));
  return IAMRX_OK;
}

static CellMG& diff_mg(Level& L, LevelSolvers& sv, bool tensor, int ncomp, int mc) {
  const int key = ncomp * 2 + (tensor ? 1 : 0);
  auto it = sv.diff.find(key);
  if (it == sv.diff.end() || sv.diff_mc[key] != mc) {
    sv.diff[key] = std::make_unique<CellMG>(&L, ncomp, tensor, mc);
    sv.diff_mc[key] = mc;
  }
  return *sv.diff[key];
}

int diffusion_apply(Level& L, LevelSolvers& sv, bool tensor, int ncomp, MF& out, MF& soln, double a, double b,
                    const MF* acoef, MF eta[3], cudaStream_t s, const k::LinBC* bc) {
  iamrx_mg_info mi = info_or_default(nullptr);
  CellMG& mg = diff_mg(L, sv, tensor, ncomp, mi.max_coarsening);
  mg.set_scalars(a, b);
  mg.set_bc(bc ? *bc : periodic_linbc());
  IX_TRY(mg.set_coeffs(acoef, &eta[0], &eta[1], &eta[2], s, /*finest_only=*/true));
  return mg.apply(out, soln, s);
}

int diffusion_solve(Level& L, LevelSolvers& sv, bool tensor, int ncomp, MF& soln, const MF& rhs, double a,
                    double b, const MF* acoef, MF eta[3], iamrx_mg_info* info, cudaStream_t s, const k::LinBC* bc) {
  iamrx_mg_info mi = info_or_default(info);
  CellMG& mg = diff_mg(L, sv, tensor, ncomp, mi.max_coarsening);
  mg.set_scalars(a, b);
  mg.set_bc(bc ? *bc : periodic_linbc());
  IX_TRY(mg.set_coeffs(acoef, &eta[0], &eta[1], &eta[2], s));
  const int rc = mg.solve(soln, rhs, &mi, s);
  if (info) *info = mi;
  return rc;
}

}  // namespace ix
