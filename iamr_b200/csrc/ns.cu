// ns.cu -- the level time step: a host-side restatement of the operator sequence
// of NavierStokes::advance (Source/NavierStokes.cpp:543-691) and of the start-up
// sequence NavierStokes::post_init (NavierStokes.cpp:1254-1432) for a single-level
// periodic problem, calling the level solvers (solvers.h) and kernels (kernels.h).
// This is the caller of the hot path in this repository -- the role IAMR's own
// NavierStokes / NavierStokesBase classes play in the reference.  Each step names
// the reference lines it mirrors.
#include <algorithm>
#include <cmath>
#include "solvers.h"

namespace ix {
Level* level_of(iamrx_level_t h);
LevelSolvers* solvers_of(iamrx_level_t h);
int write_plotfile(Level& L, const std::vector<const MF*>& mfs, const std::vector<int>& comp0, const std::vector<int>& ncomps,
                   const std::vector<std::string>& names, const char* dir, const char* plot_type, double time, int level_steps,
                   cudaStream_t s);
}  // namespace ix

using namespace ix;

#define IX_TRY(call) do { int rc_ = (call); if (rc_ != IAMRX_OK) return rc_; } while (0)
// solver calls return >0 on non-convergence: treat as failure of the step ("MLMG failed to converge" aborts)
#define IX_SOLVE(call) do { int rc_ = (call); if (rc_ != IAMRX_OK) return rc_; } while (0)

enum { Xvel = 0, Yvel = 1, Zvel = 2, Density = 3, Tracer = 4, NUM_STATE = 5, NUM_SCALARS = 2 };

struct iamrx_ns_s {
  iamrx_level_t hlev = nullptr;
  Level* L = nullptr;
  LevelSolvers* sv = nullptr;
  iamrx_ns_params p{};
  cudaStream_t s = nullptr;

  MF S_old, S_new;      // u,v,w,rho,tracer ; 1 ghost (NavierStokesBase.H:741)
  MF P_old, P_new;      // nodal, 1 ghost (NS_setup.cpp:329-334)
  MF Gp_old, Gp_new;    // 3 comps, 1 ghost (NS_setup.cpp:339-360)
  MF umac[3];           // face MFs, 1 ghost (NSB.cpp:663-675)
  MF aofs;              // NUM_STATE comps (NSB.cpp:680)
  MF rho_ptime, rho_ctime, rho_half;  // 1 ghost
  MF Umf, Smf;          // FillPatched copies, 3 ghosts (NSB.cpp:4399,4435; nghost_state :4539-4552)
  MF visc, force;       // 3 comps, 1 ghost (nghost_force, NavierStokesBase.H:810)
  MF sforce;            // scalar forcing, 2 comps, 1 ghost
  MF eta[3];            // face viscosity (getViscosity, NS.cpp:2120-)
  MF soln, rhs3, tmp3;  // diffusion work: Soln (1 ghost), Rhs
  MF sig;               // 1/rho_half, 1 ghost
  MF mac_phi;           // 1 ghost
  MF tf0;               // estTimeStep forces, 0 ghost
  MF seta[3];           // tracer diffusivity on faces (getDiffusivity: constant ns.scal_diff_coefs, NS.cpp:2051-2119)
  MF s1, r1, svisc, ones;  // tracer diffusion work: Soln (1 ghost), Rhs, visc term (1 ghost), alpha = 1
  MF smm;                  // old scalars (density, tracer) with 1 filled ghost for ns.do_scalminmax

  // HIT turbulent forcing (Tutorials/HIT/NS_getForce.cpp:205-640): off unless iamrx_ns_set_turbulent_forcing was called
  bool turb_on = false;
  k::TurbParams turb{};
  std::vector<double> turb_data;   // host copy of TurbulentForcing::forcedata
  MF turbf;                        // f(x, t) per unit mass, 3 comps, 1 ghost
  double* turb_dev = nullptr; size_t turb_dev_n = 0;   // device scratch: mode list + axis tables
  double turb_time = 0.0; int turb_ng = -1; bool turb_time_ok = false;   // what turbf currently holds

  double time = 0.0, dt_level = 0.0, dt_min = 1.0e100;
  int nstep = 0;
  bool initial_step = false, initial_iter = false;
  int it_mac = 0, it_visc = 0, it_nodal = 0;
  double* stage = nullptr; size_t stage_n = 0;  // host<->device staging for step_host
  // step_host: the new density / tracer are final long before the velocity (diffusion solve + nodal projection still
  // to come), so their device->host copy runs on a second stream underneath the rest of the step
#if !defined(IX_EMUL)
  cudaStream_t s2 = nullptr;
  cudaEvent_t ev_scal = nullptr;
#endif
  // ... and the tracer is not read before scalar_advection (after the MAC projection): its host->device copy runs on the second
  // stream underneath predict_velocity and the MAC solve; late_tracer marks a step whose tracer is still in flight
  bool late_tracer = false;
#if !defined(IX_EMUL)
  cudaEvent_t ev_in = nullptr, ev_prev = nullptr;
#endif
  double* early_out = nullptr;   // host destination of comps Density.. (single local box), or null
  double* early_buf = nullptr;   // device staging of the packed scalars

  // ---- physical boundaries (NS_BC.H:7-38: physical type -> math BC of every state component / of grad p) ----------------
  bool walls = false;   // some direction is not periodic
  void state_bc(int comp, int lo[3], int hi[3]) const {
    static const int norm_vel[6] = {IAMRX_BC_INT_DIR, IAMRX_BC_EXT_DIR, IAMRX_BC_FOEXTRAP, IAMRX_BC_REFLECT_ODD, IAMRX_BC_EXT_DIR, IAMRX_BC_EXT_DIR};
    static const int tang_vel[6] = {IAMRX_BC_INT_DIR, IAMRX_BC_EXT_DIR, IAMRX_BC_FOEXTRAP, IAMRX_BC_REFLECT_EVEN, IAMRX_BC_HOEXTRAP, IAMRX_BC_EXT_DIR};
    static const int scalar[6] = {IAMRX_BC_INT_DIR, IAMRX_BC_EXT_DIR, IAMRX_BC_FOEXTRAP, IAMRX_BC_REFLECT_EVEN, IAMRX_BC_FOEXTRAP, IAMRX_BC_FOEXTRAP};
    for (int d = 0; d < 3; ++d) {
      const int* t = comp < 3 ? (comp == d ? norm_vel : tang_vel) : scalar;
      lo[d] = L->geom.periodic[d] ? IAMRX_BC_INT_DIR : t[p.lo_bc[d]];
      hi[d] = L->geom.periodic[d] ? IAMRX_BC_INT_DIR : t[p.hi_bc[d]];
    }
  }
  k::PhysBC phys_bc(int scomp, int ncomp) const {   // BCRec + ext_dir values of state components scomp .. scomp+ncomp-1
    k::PhysBC b{};
    for (int c = 0; c < ncomp; ++c) {
      state_bc(scomp + c, b.lo[c], b.hi[c]);
      for (int f = 0; f < 6; ++f) b.val[f][c] = p.bc_vals[f][scomp + c];
    }
    return b;
  }
  k::PhysBC gradp_bc() const {   // norm_gradp_bc / tang_gradp_bc
    static const int norm_gp[6] = {IAMRX_BC_INT_DIR, IAMRX_BC_FOEXTRAP, IAMRX_BC_FOEXTRAP, IAMRX_BC_REFLECT_ODD, IAMRX_BC_FOEXTRAP, IAMRX_BC_FOEXTRAP};
    static const int tang_gp[6] = {IAMRX_BC_INT_DIR, IAMRX_BC_FOEXTRAP, IAMRX_BC_FOEXTRAP, IAMRX_BC_REFLECT_EVEN, IAMRX_BC_FOEXTRAP, IAMRX_BC_FOEXTRAP};
    k::PhysBC b{};
    for (int c = 0; c < 3; ++c)
      for (int d = 0; d < 3; ++d) {
        const int* t = c == d ? norm_gp : tang_gp;
        b.lo[c][d] = L->geom.periodic[d] ? IAMRX_BC_INT_DIR : t[p.lo_bc[d]];
        b.hi[c][d] = L->geom.periodic[d] ? IAMRX_BC_INT_DIR : t[p.hi_bc[d]];
      }
    return b;
  }
  k::PhysBC foextrap_bc(int ncomp) const {   // Extrapolater::FirstOrderExtrap
    k::PhysBC b{};
    for (int c = 0; c < ncomp; ++c)
      for (int d = 0; d < 3; ++d) { b.lo[c][d] = b.hi[c][d] = L->geom.periodic[d] ? IAMRX_BC_INT_DIR : IAMRX_BC_FOEXTRAP; }
    return b;
  }
  k::AdvBC adv_bc(int scomp, int ncomp) const {   // what fetchBCArray hands to the Godunov routines (NSB.cpp:4645)
    k::AdvBC b{};
    for (int d = 0; d < 3; ++d) { b.dlo[d] = L->domain.lo[d]; b.dhi[d] = L->domain.hi[d]; }
    for (int c = 0; c < ncomp; ++c) state_bc(scomp + c, b.lo[c], b.hi[c]);
    return b;
  }
  static int linop_of(int math) {   // Diffusion::setDomainBC (Diffusion.cpp:1887-1999)
    return math == IAMRX_BC_EXT_DIR ? IAMRX_LINOP_DIRICHLET
         : math == IAMRX_BC_REFLECT_ODD ? IAMRX_LINOP_REFLECT_ODD : math == IAMRX_BC_INT_DIR ? IAMRX_LINOP_PERIODIC : IAMRX_LINOP_NEUMANN;
  }
  k::LinBC diff_bc(int scomp, int ncomp) const {   // diffuse.max_order = diffuse.tensor_max_order = 2 (Diffusion.cpp:95-96,102-103)
    k::LinBC b = periodic_linbc();
    b.maxorder = 2;
    for (int c = 0; c < 3; ++c) {
      int lo[3], hi[3];
      state_bc(scomp + (c < ncomp ? c : 0), lo, hi);
      for (int d = 0; d < 3; ++d) { b.lo[c][d] = linop_of(lo[d]); b.hi[c][d] = linop_of(hi[d]); }
    }
    return b;
  }
  k::LinBC mac_bc() const {   // set_mac_solve_bc (MacProj.cpp:1187-1208); mac_proj.maxorder = 4 (MacProj.cpp:30)
    k::LinBC b = periodic_linbc();
    b.maxorder = 4;
    for (int c = 0; c < 3; ++c)
      for (int d = 0; d < 3; ++d) {
        if (L->geom.periodic[d]) continue;
        b.lo[c][d] = p.lo_bc[d] == 2 ? IAMRX_LINOP_DIRICHLET : IAMRX_LINOP_NEUMANN;
        b.hi[c][d] = p.hi_bc[d] == 2 ? IAMRX_LINOP_DIRICHLET : IAMRX_LINOP_NEUMANN;
      }
    return b;
  }
  k::NodalBC nodal_bc() const {   // Projection.cpp:2436-2464
    k::NodalBC b;
    for (int d = 0; d < 3; ++d) {
      b.lo[d] = L->geom.periodic[d] ? IAMRX_LINOP_PERIODIC : (p.lo_bc[d] == 2 ? IAMRX_LINOP_DIRICHLET : p.lo_bc[d] == 1 ? IAMRX_LINOP_INFLOW : IAMRX_LINOP_NEUMANN);
      b.hi[d] = L->geom.periodic[d] ? IAMRX_LINOP_PERIODIC : (p.hi_bc[d] == 2 ? IAMRX_LINOP_DIRICHLET : p.hi_bc[d] == 1 ? IAMRX_LINOP_INFLOW : IAMRX_LINOP_NEUMANN);
    }
    return b;
  }

  bool diffusive_vel() const { return p.visc_coef > 0.0; }
  bool diffusive_tracer() const { return p.scal_diff_coef > 0.0; }      // is_diffusive[Tracer], NS_setup.cpp:292-295
  int rho_flag() const { return p.conservative_tracer ? 2 : 0; }         // set_rho_flag(diffusionType[Tracer]), NS_setup.cpp:304-308
};

namespace {

// AmrLevel::FillPatch of State_Type on one level: ghost cells of a state MF whose valid cells are in place
int fill_state_ghosts(iamrx_ns_s& ns, MF& m, int mcomp, int scomp, int ncomp, int ng) {
  IX_TRY(mf_fill_boundary(m, mcomp, ncomp, ng, ns.s));
  if (ns.walls) IX_TRY(mf_fill_physbc(m, mcomp, ncomp, ng, ns.phys_bc(scomp, ncomp), ns.s));   // NS_bcfill.H
  return IAMRX_OK;
}
int fillpatch(iamrx_ns_s& ns, MF& dst, const MF& src, int scomp, int ncomp) {
  // single-level FillPatch = copy of the valid region + FillBoundary + the physical boundary fill
  IX_TRY(mf_copy(dst, src, scomp, 0, ncomp, 0, ns.s));
  return fill_state_ghosts(ns, dst, 0, scomp, ncomp, dst.ng);
}
int fill_gradp(iamrx_ns_s& ns, MF& gp) {   // FillPatch of Gradp_Type (NS_setup.cpp:339-360)
  IX_TRY(mf_fill_boundary(gp, 0, 3, 1, ns.s));
  if (ns.walls) IX_TRY(mf_fill_physbc(gp, 0, 3, 1, ns.gradp_bc(), ns.s));
  return IAMRX_OK;
}
int fill_extrap(iamrx_ns_s& ns, MF& m, int ncomp) {   // FillBoundary + Extrapolater::FirstOrderExtrap (NS.cpp:2045-2046)
  IX_TRY(mf_fill_boundary(m, 0, ncomp, 1, ns.s));
  if (ns.walls) IX_TRY(mf_fill_physbc(m, 0, ncomp, 1, ns.foextrap_bc(ncomp), ns.s));
  return IAMRX_OK;
}

iamrx_mg_info mg_info(const iamrx_ns_s& ns, double rtol, double atol) {
  iamrx_mg_info mi;
  iamrx_mg_info_default(&mi);
  mi.rtol = rtol; mi.atol = atol; mi.verbose = ns.p.mg_verbose;
  mi.bottom_solver = ns.p.bottom_solver;
  return mi;
}

// NavierStokes::getViscTerms for velocity (NS.cpp:1960-2049) ->
// Diffusion::getTensorViscTerms (Diffusion.cpp:1655-1777): visc = div(tau(U^n)), a=0, b=-1,
// then FillBoundary on the grow cell (NS.cpp:2046; FirstOrderExtrap is a no-op when periodic)
int get_visc_terms(iamrx_ns_s& ns, MF& visc, const MF& S) {
  if (!ns.diffusive_vel()) return mf_setval(visc, 0.0, 0, 3, 1, ns.s);
  IX_TRY(fillpatch(ns, ns.soln, S, Xvel, 3));
  const k::LinBC vbc = ns.diff_bc(Xvel, 3);
  IX_TRY(diffusion_apply(*ns.L, *ns.sv, true, 3, visc, ns.soln, 0.0, -1.0, nullptr, ns.eta, ns.s, &vbc));
  return fill_extrap(ns, visc, 3);
}

// the user force of getForce at `time` on the ng-ghost boxes -> ns.turbf (per unit mass; the callers weight it with the density
// getForce is handed at that call site)
int turb_eval(iamrx_ns_s& ns, double time, int ng) {
  if (!ns.turb_on) return IAMRX_OK;
  // estTimeStep's cur_time evaluation at the end of a step is the next step's prev_time evaluation: keep it
  if (ns.turb_time_ok && time == ns.turb_time && ng <= ns.turb_ng) return IAMRX_OK;
  ns.turb_time_ok = false;
  Level& L = *ns.L;
  double len[3];
  for (int d = 0; d < 3; ++d) len[d] = (L.domain.hi[d] - L.domain.lo[d] + 1) * L.geom.dx[d];
  std::vector<k::TurbMode> modes;
  IX_TRY(k::turb_modes(ns.turb, ns.turb_data.data(), len, time, modes));
  const int nm = (int)modes.size();
  const size_t mode_doubles = (modes.size() * sizeof(k::TurbMode) + 7) / 8;
  size_t need = mode_doubles + 16;
  for (int il = 0; il < ns.turbf.n(); ++il) need = std::max(need, mode_doubles + 16 + k::turb_scratch_doubles(ns.turbf.gbox(il, ng), nm));
  if (need > ns.turb_dev_n) {
    if (ns.turb_dev) dev_free(ns.turb_dev);
    ns.turb_dev = dev_alloc(need); ns.turb_dev_n = ns.turb_dev ? need : 0;
    if (!ns.turb_dev) return IAMRX_ERR_CUDA;
  }
  if (nm > 0) IX_CUDA(cudaMemcpyAsync(ns.turb_dev, modes.data(), modes.size() * sizeof(k::TurbMode), cudaMemcpyHostToDevice, ns.s));
  IX_CUDA(cudaStreamSynchronize(ns.s));   // `modes` is a local: the copy must have left it
  for (int il = 0; il < ns.turbf.n(); ++il)
    IX_TRY(k::turb_force(ns.turbf.gbox(il, ng), ns.turbf.v(il), C4{}, L.geom, reinterpret_cast<const k::TurbMode*>(ns.turb_dev), nm,
                         ns.turb.div_free, ns.turb_dev + mode_doubles + 8, 0, ns.s));
  ns.turb_time = time; ns.turb_ng = ng; ns.turb_time_ok = true;
  return IAMRX_OK;
}
inline C4 turb_c(const iamrx_ns_s& ns, int il) { return ns.turb_on ? ns.turbf.c(il) : C4{}; }

// velocity forcing on the 1-ghost box: tf = (getForce + visc - gp)/rho  (NSB.cpp:4456-4470 == 3445-3466)
int vel_forcing(iamrx_ns_s& ns) {
  for (int il = 0; il < ns.force.n(); ++il)
    IX_TRY(k::force_vel(ns.force.gbox(il, 1), ns.force.v(il), ns.visc.c(il), ns.Gp_old.c(il), ns.Smf.c(il, 0),
                        ns.p.gravity, 1, ns.s, turb_c(ns, il)));
  return IAMRX_OK;
}

// NavierStokesBase::predict_velocity (NSB.cpp:4376-4512)
int predict_velocity(iamrx_ns_s& ns, double dt, double* dt_test) {
  Level& L = *ns.L;
  IX_TRY(fillpatch(ns, ns.Umf, ns.S_old, Xvel, 3));                       // :4399
  for (int il = 0; il < ns.Umf.n(); ++il) IX_TRY(k::floor_small(ns.Umf.gbox(il, 3), ns.Umf.v(il), 3, ns.s));  // :4403
  double umax[3];
  IX_TRY(mf_norminf_each(ns.Umf, 0, 3, umax, ns.s));                       // :4408 (ghosts are periodic images)
  double cflmax = 0.0;
  for (int d = 0; d < 3; ++d) cflmax = std::max(cflmax, dt * umax[d] / L.geom.dx[d]);
  const double tempdt = (cflmax == 0.0) ? ns.p.change_max : std::min(ns.p.change_max, ns.p.cfl / cflmax);  // :4413
  if (ns.p.be_cn_theta != 1.0) IX_TRY(get_visc_terms(ns, ns.visc, ns.S_old));   // :4426-4433
  else IX_TRY(mf_setval(ns.visc, 0.0, 0, 3, 1, ns.s));
  IX_TRY(fillpatch(ns, ns.Smf, ns.S_old, Density, ns.late_tracer ? 1 : NUM_SCALARS));   // :4435 (a tracer still in flight is patched in scalar_advection)
  IX_TRY(vel_forcing(ns));                                                 // :4456-4470
  k::AdvGeom g; for (int d = 0; d < 3; ++d) g.dx[d] = L.geom.dx[d]; g.dt = dt;
  const k::AdvBC vbc = ns.adv_bc(Xvel, 3);
  for (int il = 0; il < ns.Umf.n(); ++il)                                  // :4487-4491
    IX_TRY(k::extrap_vel_to_faces(L.lbox(il), ns.Umf.c(il), ns.force.c(il), ns.umac[0].v(il), ns.umac[1].v(il),
                                  ns.umac[2].v(il), g, ns.p.use_forces_in_trans, ns.s, ns.p.godunov_ppm, ns.walls ? &vbc : nullptr));
  *dt_test = dt * tempdt;                                                  // :4511
  return IAMRX_OK;
}

// NavierStokesBase::mac_project (NSB.cpp:2070-2109) -> MacProj::mac_project (MacProj.cpp:225-353)
// + create_umac_grown (NSB.cpp:1068-1311; single level: FillPatchSingleLevel == periodic FillBoundary)
int mac_project_step(iamrx_ns_s& ns, double dt) {
  IX_TRY(mf_setval(ns.mac_phi, 0.0, 0, 1, 1, ns.s));                       // MacProj.cpp:253
  iamrx_mg_info mi = mg_info(ns, ns.p.mac_tol, ns.p.mac_abs_tol);
  const k::LinBC mbc = ns.mac_bc();
  IX_SOLVE(mac_project(*ns.L, *ns.sv, ns.umac, ns.rho_ptime, nullptr, ns.mac_phi, 2.0 / dt, &mi, ns.s, ns.walls ? &mbc : nullptr));  // :272,294
  ns.it_mac = mi.iters;
  for (int d = 0; d < 3; ++d) IX_TRY(mf_fill_boundary(ns.umac[d], 0, 1, 1, ns.s));  // NSB.cpp:1102
  return IAMRX_OK;
}

// sync_out / sync_comp / ucorr: the is_sync call of MacProj::mac_sync_compute (MacProj.cpp:681-707) -- the update accumulates into
// sync_out[sync_comp ..] and the fluxes are taken with the correction velocities Ucorr instead of u_mac (NSB.cpp:4672-4677)
int compute_aofs(iamrx_ns_s& ns, int state_comp, int ncomp, const MF& Sq, const MF* forcing, bool is_velocity,
                 double dt, MF* sync_out = nullptr, int sync_comp = 0, const MF* ucorr = nullptr) {
  // NavierStokesBase::ComputeAofs (NSB.cpp:4555-4591 wrapper, :4594-4845 body)
  Level& L = *ns.L;
  k::AdvGeom g; for (int d = 0; d < 3; ++d) g.dx[d] = L.geom.dx[d]; g.dt = dt;
  for (int il = 0; il < ns.aofs.n(); ++il) {
    k::AofsArgs a{};
    a.aofs = sync_out ? sync_out->v(il, sync_comp) : ns.aofs.v(il, state_comp);
    a.S = Sq.c(il);
    a.force = forcing ? forcing->c(il) : C4{};
    a.divu = C4{};  // have_divu == 0: getDivCond returns zeros (NSB.cpp:1577-1590)
    a.umac = ns.umac[0].c(il); a.vmac = ns.umac[1].c(il); a.wmac = ns.umac[2].c(il);
    a.uflx = a.umac; a.vflx = a.vmac; a.wflx = a.wmac;
    if (sync_out) { a.is_sync = 1; a.uflx = ucorr[0].c(il); a.vflx = ucorr[1].c(il); a.wflx = ucorr[2].c(il); }
    a.ncomp = ncomp;
    for (int n = 0; n < ncomp; ++n) {
      const int sc = state_comp + n;
      // advectionType: NS_setup.cpp:285-320 (velocity non-conservative unless do_mom_diff,
      // density conservative, tracer per do_cons_trac)
      a.iconserv[n] = (sc == Density) ? 1 : (sc == Tracer ? (ns.p.conservative_tracer ? 1 : 0) : (ns.p.do_mom_diff ? 1 : 0));
    }
    a.forces_in_trans = ns.p.use_forces_in_trans;
    a.ppm = ns.p.godunov_ppm;   // advection_scheme == Godunov_PPM (NSB.cpp:4609)
    a.is_velocity = is_velocity ? 1 : 0;
    if (ns.walls) a.bc = ns.adv_bc(state_comp, ncomp);
    IX_TRY(k::compute_aofs(L.lbox(il), a, g, ns.s));
  }
  return IAMRX_OK;
}

// NavierStokesBase::velocity_advection (NSB.cpp:3358-3470)
int velocity_advection(iamrx_ns_s& ns, double dt) {
  IX_TRY(fillpatch(ns, ns.Umf, ns.S_old, Xvel, 3));   // :3382 (a fresh, un-floored copy)
  if (ns.p.do_mom_diff) {
    // :3390-3414 the advected state is the momentum rho^n u^n on the 3-ghost box; :3459-3466 the forcing is NOT divided by rho.
    // ns.Smf(0) is the old density with 3 filled ghosts (its floor at 1e-20 never bites a density); ns.visc still holds visc^n.
    for (int il = 0; il < ns.Umf.n(); ++il) {
      IX_TRY(k::mult(ns.Umf.gbox(il, 3), ns.Umf.v(il), ns.Smf.c(il, 0), 3, 1, ns.s));
      IX_TRY(k::force_vel(ns.force.gbox(il, 1), ns.force.v(il), ns.visc.c(il), ns.Gp_old.c(il), ns.Smf.c(il, 0), ns.p.gravity, 0, ns.s, turb_c(ns, il)));
    }
    return compute_aofs(ns, Xvel, 3, ns.Umf, &ns.force, true, dt);  // conservative: advectionType[Xvel..] (NS_setup.cpp:297-299)
  }
  // visc_terms (:3429) and the total forcing (:3445-3466) repeat predict_velocity's
  // arithmetic on the same inputs; ns.force still holds exactly those values.
  return compute_aofs(ns, Xvel, 3, ns.Umf, &ns.force, true, dt);  // :3469
}

// NavierStokes::scalar_advection (NS.cpp:698-812)
int scalar_advection(iamrx_ns_s& ns, double dt) {
#if !defined(IX_EMUL)
  if (ns.late_tracer) {   // step_host: the tracer's upload ran underneath predict_velocity and the MAC solve; first use is here
    IX_CUDA(cudaStreamWaitEvent(ns.s, ns.ev_in, 0));
    IX_TRY(mf_copy(ns.Smf, ns.S_old, Tracer, Tracer - Density, 1, 0, ns.s));
    IX_TRY(fill_state_ghosts(ns, ns.Smf, Tracer - Density, Tracer, 1, ns.Smf.ng));
    ns.late_tracer = false;
  }
#endif
  for (int il = 0; il < ns.Smf.n(); ++il) IX_TRY(k::floor_small(ns.Smf.gbox(il, 3), ns.Smf.v(il), NUM_SCALARS, ns.s));  // :722
  IX_TRY(mf_setval(ns.sforce, 0.0, 0, NUM_SCALARS, 1, ns.s));  // getForce: zero scalar forcing (:755-760)
  if (ns.diffusive_tracer() && ns.p.be_cn_theta != 1.0) {
    // getViscTerms (NS.cpp:2012-2048) -> Diffusion::getViscTerms (Diffusion.cpp:1540-1652): a = 0, b = -1 applied to S (rho_flag 0) or
    // S/rho (rho_flag 2), FillBoundary; then tf = tf/rho + visc or tf + visc with a zero body force (NS.cpp:774-804)
    Level& L = *ns.L;
    IX_TRY(fillpatch(ns, ns.s1, ns.S_old, Tracer, 1));
    if (ns.rho_flag() == 2)
      for (int il = 0; il < ns.s1.n(); ++il) IX_TRY(k::divide(ns.s1.gbox(il, 1), ns.s1.v(il), ns.rho_ptime.c(il), 1, 1, ns.s));
    const k::LinBC tbc = ns.diff_bc(Tracer, 1);
    IX_TRY(diffusion_apply(L, *ns.sv, false, 1, ns.svisc, ns.s1, 0.0, -1.0, nullptr, ns.seta, ns.s, &tbc));
    IX_TRY(fill_extrap(ns, ns.svisc, 1));
    IX_TRY(mf_copy(ns.sforce, ns.svisc, 0, Tracer - Density, 1, 1, ns.s));
  }
  return compute_aofs(ns, Density, NUM_SCALARS, ns.Smf, &ns.sforce, false, dt);  // :811
}

// NavierStokes::scalar_diffusion_update (NS.cpp:858-1000) -> Diffusion::diffuse_scalar (Diffusion.cpp:207-600) for the tracer:
// Crank-Nicolson with rho_flag 0 (S diffuses, alpha = 1) or 2 (S/rho diffuses, alpha = rho_new, result times rho_new)
int tracer_diffusion_update(iamrx_ns_s& ns, double dt) {
  if (!ns.diffusive_tracer()) return IAMRX_OK;
  Level& L = *ns.L;
  const double theta = ns.p.be_cn_theta;
  const int rf = ns.rho_flag();
  if (theta != 1.0) {   // :364-430: Rhs = (1-theta) dt div beta grad (old solution): a = 0, b = -(1-theta) dt
    IX_TRY(fillpatch(ns, ns.s1, ns.S_old, Tracer, 1));
    if (rf == 2)
      for (int il = 0; il < ns.s1.n(); ++il) IX_TRY(k::divide(ns.s1.gbox(il, 1), ns.s1.v(il), ns.rho_ptime.c(il), 1, 1, ns.s));
    const k::LinBC tbc0 = ns.diff_bc(Tracer, 1);
    IX_TRY(diffusion_apply(L, *ns.sv, false, 1, ns.r1, ns.s1, 0.0, -(1.0 - theta) * dt, nullptr, ns.seta, ns.s, &tbc0));
  } else {
    IX_TRY(mf_setval(ns.r1, 0.0, 0, 1, 0, ns.s));
  }
  IX_TRY(mf_lincomb(ns.r1, 0, 1.0, ns.r1, 0, 1.0, ns.S_new, Tracer, 1, 0, ns.s));   // :465-490 Rhs += S_new (no scaling for rho_flag 0, 2)
  double nrm = 0.0;
  IX_TRY(mf_norminf(ns.r1, 0, 1, &nrm, ns.s));
  const double tol_abs = ns.p.visc_tol * nrm;   // get_scaled_abs_tol :193-204
  IX_TRY(fillpatch(ns, ns.s1, ns.S_new, Tracer, 1));   // :520-540 initial guess = new state (/ rho_new)
  if (rf == 2)
    for (int il = 0; il < ns.s1.n(); ++il) IX_TRY(k::divide(ns.s1.gbox(il, 1), ns.s1.v(il), ns.rho_ctime.c(il), 1, 1, ns.s));
  iamrx_mg_info mi = mg_info(ns, ns.p.visc_tol, tol_abs);
  // computeAlpha (:1355-1395): alpha = 1 (rho_flag 0) or rho_new (rho_flag 2); a = 1, b = theta dt
  const k::LinBC tbc = ns.diff_bc(Tracer, 1);
  IX_SOLVE(diffusion_solve(L, *ns.sv, false, 1, ns.s1, ns.r1, 1.0, theta * dt, rf == 2 ? &ns.rho_ctime : &ns.ones, ns.seta, &mi, ns.s, &tbc));
  if (rf == 2)   // :575-590
    for (int il = 0; il < ns.s1.n(); ++il) IX_TRY(k::mult(L.lbox(il), ns.s1.v(il), ns.rho_ctime.c(il), 1, 1, ns.s));
  return mf_copy(ns.S_new, ns.s1, 0, Tracer, 1, 0, ns.s);
}

// Diffusion::diffuse_tensor_velocity (Diffusion.cpp:650-957), rho_flag = 1 (3 with do_mom_diff)
int velocity_diffusion_update(iamrx_ns_s& ns, double dt) {
  if (!ns.diffusive_vel()) return IAMRX_OK;
  Level& L = *ns.L;
  const double theta = ns.p.be_cn_theta;
  if (theta != 1.0) {
    IX_TRY(fillpatch(ns, ns.soln, ns.S_old, Xvel, 3));  // :742
    const k::LinBC vbc0 = ns.diff_bc(Xvel, 3);
    IX_TRY(diffusion_apply(L, *ns.sv, true, 3, ns.rhs3, ns.soln, 0.0, -(1.0 - theta) * dt, nullptr, ns.eta, ns.s, &vbc0));  // :702-768
  } else {
    IX_TRY(mf_setval(ns.rhs3, 0.0, 0, 3, 0, ns.s));
  }
  for (int il = 0; il < ns.rhs3.n(); ++il)  // :821-831
    IX_TRY(k::diff_rhs(L.lbox(il), ns.rhs3.v(il), ns.S_new.v(il, Xvel),   // rho_flag 1: rho_half; 3 (do_mom_diff): the OLD density (:819)
                       ns.p.do_mom_diff ? ns.S_old.c(il, Density) : ns.rho_half.c(il), 3, ns.s));
  // tolerances: visc_tol, get_scaled_abs_tol (Diffusion.cpp:193-204, :846-847)
  double nrm[3];
  IX_TRY(mf_norminf_each(ns.rhs3, 0, 3, nrm, ns.s));
  const double tol_abs = ns.p.visc_tol * (nrm[0] + nrm[1] + nrm[2]) / 3.0;
  IX_TRY(fillpatch(ns, ns.soln, ns.S_new, Xvel, 3));  // :885 initial guess = new-time state
  iamrx_mg_info mi = mg_info(ns, ns.p.visc_tol, tol_abs);
  // :893-897 alpha = rho_half (rho_flag 1) or the NEW density (rho_flag 3, NS.cpp:1016)
  const k::LinBC vbc = ns.diff_bc(Xvel, 3);
  IX_SOLVE(diffusion_solve(L, *ns.sv, true, 3, ns.soln, ns.rhs3, 1.0, theta * dt, ns.p.do_mom_diff ? &ns.rho_ctime : &ns.rho_half, ns.eta,
                           &mi, ns.s, &vbc));  // :895-923
  ns.it_visc = mi.iters;
  return mf_copy(ns.S_new, ns.soln, 0, Xvel, 3, 1, ns.s);  // :928
}

// NavierStokesBase::initial_velocity_diffusion_update (NSB.cpp:3658-3749)
int initial_velocity_diffusion_update(iamrx_ns_s& ns, double time, double dt) {
  if (!ns.diffusive_vel()) return IAMRX_OK;
  Level& L = *ns.L;
  IX_TRY(turb_eval(ns, time, 0));   // getForce at prev_time (:3696)
  if (ns.p.be_cn_theta != 1.0) IX_TRY(get_visc_terms(ns, ns.visc, ns.S_old));
  else IX_TRY(mf_setval(ns.visc, 0.0, 0, 3, 1, ns.s));
  for (int il = 0; il < ns.tf0.n(); ++il) {
    // force = (getForce(rho_old) + visc - gp)/rho_half - aofs ; u_new = u_old + dt*force
    IX_TRY(k::force_vel(L.lbox(il), ns.tf0.v(il), ns.visc.c(il), ns.Gp_old.c(il), ns.S_old.c(il, Density),
                        ns.p.gravity, 0, ns.s, turb_c(ns, il)));
    if (!ns.p.do_mom_diff) IX_TRY(k::divide(L.lbox(il), ns.tf0.v(il), ns.rho_half.c(il), 3, 1, ns.s));
    IX_TRY(k::lincomb(L.lbox(il), ns.tf0.v(il), 1.0, ns.tf0.c(il), -1.0, ns.aofs.c(il, Xvel), 3, ns.s));
    if (ns.p.do_mom_diff) {   // :3743-3744 u_new = (force dt + u_old rho_old) / rho_new
      IX_TRY(k::copy(L.lbox(il), ns.S_new.v(il, Xvel), ns.S_old.c(il, Xvel), 3, ns.s));
      IX_TRY(k::mult(L.lbox(il), ns.S_new.v(il, Xvel), ns.S_old.c(il, Density), 3, 1, ns.s));
      IX_TRY(k::lincomb(L.lbox(il), ns.S_new.v(il, Xvel), 1.0, ns.S_new.c(il, Xvel), dt, ns.tf0.c(il), 3, ns.s));
      IX_TRY(k::divide(L.lbox(il), ns.S_new.v(il, Xvel), ns.S_new.c(il, Density), 3, 1, ns.s));
    } else {
      IX_TRY(k::lincomb(L.lbox(il), ns.S_new.v(il, Xvel), 1.0, ns.S_old.c(il, Xvel), dt, ns.tf0.c(il), 3, ns.s));
    }
  }
  return IAMRX_OK;
}

// Projection::level_project (Projection.cpp:166-450) + doMLMGNodalProjection (:2385-2567)
int inflow_ghost_velocity(iamrx_ns_s& ns, MF& vel, double scale);

// Projection::set_outflow_bcs for LEVEL_PROJ / INITIAL_PRESS without divu (Projection.cpp:1721-1931 -> computeRhoG :1933-2379): under
// gravity the nodes of an outflow face in x or y take the hydrostatic pressure of the density next to the face, integrated down
// from the top of the domain; an outflow face on top keeps phi = 0, one at the bottom is refused (the reference aborts, :1957).
// rho: the density the projection uses (rho_half / the new density), valid cells; its ghost cells are refilled here on a copy.
bool outflow_with_gravity(const iamrx_ns_s& ns) {
  if (std::fabs(ns.p.gravity) == 0.0) return false;
  for (int d = 0; d < 3; ++d) if (!ns.L->geom.periodic[d] && (ns.p.lo_bc[d] == 2 || ns.p.hi_bc[d] == 2)) return true;
  return false;
}
int set_outflow_bcs(iamrx_ns_s& ns, MF& phi, const MF& rho, int rho_comp) {
  if (!outflow_with_gravity(ns)) return IAMRX_OK;
  Level& L = *ns.L;
  // the density as ONE box over the domain with a filled ghost layer (the reference copies a two-cell strip to one FAB, :1891-1893)
  std::vector<Bx> one{L.domain};
  std::vector<int> own{comm().rank};
  std::unique_ptr<Level> RL = make_level(L.geom, one, own);
  RL->replicated = true;
  MF rr(RL.get(), IX_CELL, 1, 1);
  MF rv(&L, IX_CELL, 1, 0);
  IX_TRY(mf_copy(rv, rho, rho_comp, 0, 1, 0, ns.s));
  if (L.replicated || (L.boxes.size() == 1 && L.nlocal() == 1)) IX_TRY(mf_copy(rr, rv, 0, 0, 1, 0, ns.s));
  else IX_TRY(mf_gather_replicate(rr, rv, 1, ns.s));
  IX_TRY(mf_fill_boundary(rr, 0, 1, 1, ns.s));
  const k::PhysBC dbc = ns.phys_bc(Density, 1);
  IX_TRY(mf_fill_physbc(rr, 0, 1, 1, dbc, ns.s));
  for (int d = 0; d < 2; ++d) {
    if (L.geom.periodic[d]) continue;
    for (int side = 0; side < 2; ++side) {
      if ((side == 0 ? ns.p.lo_bc[d] : ns.p.hi_bc[d]) != 2) continue;
      k::OutflowRhoG a{};
      a.d = d; a.t = 1 - d;
      a.c1 = side == 0 ? L.domain.lo[d] : L.domain.hi[d];
      a.c2 = side == 0 ? a.c1 + 1 : a.c1 - 1;
      a.tlo = L.domain.lo[a.t]; a.thi = L.domain.hi[a.t] + 1;
      a.code_lo = L.geom.periodic[a.t] ? IAMRX_BC_INT_DIR : dbc.lo[0][a.t];
      a.code_hi = L.geom.periodic[a.t] ? IAMRX_BC_INT_DIR : dbc.hi[0][a.t];
      a.ztop = L.domain.hi[2];
      a.gravity = ns.p.gravity; a.dz = L.geom.dx[2];
      const int plane = side == 0 ? L.domain.lo[d] : L.domain.hi[d] + 1;
      for (int il = 0; il < phi.n(); ++il) {
        Bx strip = phi.vbox(il);
        if (plane < strip.lo[d] || plane > strip.hi[d]) continue;
        strip.lo[d] = strip.hi[d] = plane;
        IX_TRY(k::outflow_rhog(strip, phi.v(il), rr.c(0), a, ns.s));
      }
    }
  }
  return IAMRX_OK;
}

int level_project(iamrx_ns_s& ns, double dt) {
  Level& L = *ns.L;
  IX_TRY(mf_setval(ns.P_new, 0.0, 0, 1, 1, ns.s));   // :247-256
  IX_TRY(set_outflow_bcs(ns, ns.P_new, ns.rho_half, 0));   // :304-324
  for (int il = 0; il < ns.S_new.n(); ++il)          // :273, :296-300
    IX_TRY(k::proj_pre(L.lbox(il), ns.S_new.v(il, Xvel), ns.Gp_old.c(il), ns.rho_half.c(il), 1.0 / dt, ns.s));
  for (int il = 0; il < ns.sig.n(); ++il)            // scaleVar :332, :1327-1349
    IX_TRY(k::invert(L.lbox(il), ns.sig.v(il), ns.rho_half.c(il), ns.s));
  MF vel; vel.alias(&L, IX_CELL, 3, 1, ns.S_new.fabs.data());  // comps 0..2 of the state
  IX_TRY(inflow_ghost_velocity(ns, vel, 1.0 / dt));
  iamrx_mg_info mi = mg_info(ns, ns.p.proj_tol, ns.p.proj_abs_tol);
  const k::NodalBC nbc = ns.nodal_bc();
  IX_SOLVE(nodal_project(L, *ns.sv, vel, ns.sig, ns.P_new, &ns.Gp_new, 0, &mi, ns.s, ns.walls ? &nbc : nullptr, outflow_with_gravity(ns)));  // :392 ; Gp = grad phi :2542-2563
  ns.it_nodal = mi.iters;
  IX_TRY(fill_gradp(ns, ns.Gp_new));  // :2565
  return mf_scale(ns.S_new, dt, Xvel, 3, 0, ns.s);     // :438 (rescaleVar :434 restores rho_half: ns.sig is separate)
}

// NavierStokes::advance (NS.cpp:543-691)
int advance(iamrx_ns_s& ns, double time, double dt, double* dt_test) {
  Level& L = *ns.L;
  IX_TRY(turb_eval(ns, time, 1));   // getForce at prev_time for predict_velocity / velocity_advection (NSB.cpp:4453, 3448)
  // advance_setup (NSB.cpp:613-741): swap time levels, rho at the previous time
  std::swap(ns.S_old, ns.S_new);
  std::swap(ns.P_old, ns.P_new);
  std::swap(ns.Gp_old, ns.Gp_new);
  IX_TRY(fillpatch(ns, ns.rho_ptime, ns.S_old, Density, 1));   // make_rho_prev_time :703
  IX_TRY(predict_velocity(ns, dt, dt_test));                   // NS.cpp:585
  IX_TRY(mac_project_step(ns, dt));                            // :589-597
  if (!ns.p.do_mom_diff) IX_TRY(velocity_advection(ns, dt));   // :606
  IX_TRY(scalar_advection(ns, dt));                            // :613
  for (int il = 0; il < ns.S_new.n(); ++il)                    // scalar_update(rho) :617 -> NSB.cpp:2761-2765
    IX_TRY(k::scal_update(L.lbox(il), ns.S_new.v(il, Density), ns.S_old.c(il, Density), ns.aofs.c(il, Density), dt, 1, ns.s));
  IX_TRY(fillpatch(ns, ns.rho_ctime, ns.S_new, Density, 1));   // make_rho_curr_time :618
  if (ns.p.do_mom_diff) IX_TRY(velocity_advection(ns, dt));    // :622-623 momenta, once rho^{n+1} exists
  for (int il = 0; il < ns.S_new.n(); ++il)                    // scalar_update(tracer) :627 -> NSB.cpp:2887-2896
    IX_TRY(k::scal_update(L.lbox(il), ns.S_new.v(il, Tracer), ns.S_old.c(il, Tracer), ns.aofs.c(il, Tracer), dt, 1, ns.s));
  if (ns.p.do_scalminmax) {                                    // NSB.cpp:2907-2935 (fresh FillPatch of the old scalars, 1 ghost)
    IX_TRY(fillpatch(ns, ns.smm, ns.S_old, Density, NUM_SCALARS));
    for (int il = 0; il < ns.S_new.n(); ++il)
      IX_TRY(k::scal_minmax(L.lbox(il), ns.S_new.v(il, Tracer), ns.S_new.c(il, Density), ns.smm.c(il, Tracer - Density), ns.smm.c(il, 0),
                            ns.p.conservative_tracer ? 1 : 0, ns.s));
  }
  IX_TRY(tracer_diffusion_update(ns, dt));                     // scalar_update -> scalar_diffusion_update NS.cpp:836-841
#if !defined(IX_EMUL)
  if (ns.early_out && ns.S_new.n() == 1) {   // scalars are final: pack + copy them out underneath the velocity solves
    IX_CUDA(cudaEventRecord(ns.ev_scal, ns.s));
    IX_CUDA(cudaStreamWaitEvent(ns.s2, ns.ev_scal, 0));
    IX_TRY(k::pack(L.lbox(0), ns.early_buf, ns.S_new.c(0, Density), NUM_SCALARS, ns.s2));
    IX_CUDA(cudaMemcpyAsync(ns.early_out, ns.early_buf, (size_t)L.lbox(0).npts() * NUM_SCALARS * sizeof(double),
                            cudaMemcpyDeviceToHost, ns.s2));
  }
#endif
  // velocity_update :645 -> NSB.cpp:3487: rho_half (:1561-1565), advection update (:3523-3655), diffusion
  IX_TRY(mf_lincomb(ns.rho_half, 0, 0.5, ns.rho_ptime, 0, 0.5, ns.rho_ctime, 0, 1, 1, ns.s));
  IX_TRY(turb_eval(ns, time + 0.5 * dt, 0));   // getForce at half_time with the half-time density (NSB.cpp:3581-3583)
  for (int il = 0; il < ns.S_new.n(); ++il)
    IX_TRY(k::vel_update(L.lbox(il), ns.S_new.v(il, Xvel), ns.S_old.c(il, Xvel), ns.aofs.c(il, Xvel), ns.Gp_old.c(il),
                         ns.rho_half.c(il), ns.p.gravity, dt, (ns.initial_iter && ns.diffusive_vel()) ? 1 : 0, ns.s,
                         ns.p.do_mom_diff ? ns.S_old.c(il, Density) : C4{}, ns.p.do_mom_diff ? ns.S_new.c(il, Density) : C4{},
                         turb_c(ns, il)));
  if (!ns.initial_iter) IX_TRY(velocity_diffusion_update(ns, dt));
  else IX_TRY(initial_velocity_diffusion_update(ns, time, dt));
  if (!ns.initial_step) IX_TRY(level_project(ns, dt));         // :650-670
  return IAMRX_OK;
}

// NavierStokesBase::estTimeStep (NSB.cpp:1353-1500)
int est_time_step(iamrx_ns_s& ns, double* out) {
  if (ns.p.fixed_dt > 0.0) { *out = ns.p.fixed_dt; return IAMRX_OK; }
  Level& L = *ns.L;
  const double small = 1.0e-8;
  double estdt = 1.0e20;
  double umax[3], fmax[3];
  IX_TRY(mf_norminf_each(ns.S_new, Xvel, 3, umax, ns.s));
  IX_TRY(turb_eval(ns, ns.time, 1));   // getForce at cur_time (:1410); with the ghost layer the next advance's prev_time call needs
  for (int il = 0; il < ns.tf0.n(); ++il)
    IX_TRY(k::force_vel(L.lbox(il), ns.tf0.v(il), C4{}, ns.Gp_new.c(il), ns.S_new.c(il, Density), ns.p.gravity, 1, ns.s, turb_c(ns, il)));
  IX_TRY(mf_norminf_each(ns.tf0, 0, 3, fmax, ns.s));
  for (int d = 0; d < 3; ++d) {
    if (umax[d] > small) estdt = std::min(estdt, L.geom.dx[d] / umax[d]);
    if (fmax[d] > small) estdt = std::min(estdt, std::sqrt(2.0 * L.geom.dx[d] / fmax[d]));
  }
  if (estdt < 1.0e20) estdt *= ns.p.cfl;
  else { set_error("estTimeStep: zero velocity and forcing; set fixed_dt"); return IAMRX_ERR_ARG; }
  *out = estdt;
  return IAMRX_OK;
}

// Velocity ghost cells beyond INFLOW faces as the nodal projections see them: setPhysBoundaryValues (Projection.cpp:211-217,
// 729, 1096-1097) puts the inflow velocity there and set_boundary_velocity(inflowCorner = true) (:2570-2663) keeps its normal
// component on the face cells and their periodic / interior-neighbour extensions and zeroes it in the corners outside walls.
// `scale`: 1 for initialVelocityProject, 1/dt for level_project (U_new.mult(dt_inv, ..., 1 ghost) :274), 0 for initialSyncProject
// ((U_new - U_old)/dt of a steady inflow value, ConvertUnew on the grown box :1192-1232).
int inflow_ghost_velocity(iamrx_ns_s& ns, MF& vel, double scale) {
  Level& L = *ns.L;
  for (int il = 0; il < vel.n(); ++il)
    for (int d = 0; d < 3; ++d) {
      if (L.geom.periodic[d]) continue;
      for (int side = 0; side < 2; ++side) {
        if ((side == 0 ? ns.p.lo_bc[d] : ns.p.hi_bc[d]) != 1) continue;
        const Bx b = L.lbox(il);
        if (side == 0 ? b.lo[d] != L.domain.lo[d] : b.hi[d] != L.domain.hi[d]) continue;
        const int g = side == 0 ? L.domain.lo[d] - 1 : L.domain.hi[d] + 1;
        Bx R = grow(b, 1); R.lo[d] = R.hi[d] = g;
        IX_TRY(k::setval(R, vel.v(il, d), 1, 0.0, ns.s));
        Bx P = b; P.lo[d] = P.hi[d] = g;
        for (int o = 0; o < 3; ++o) {
          if (o == d) continue;
          if (L.geom.periodic[o] || b.lo[o] != L.domain.lo[o]) P.lo[o] -= 1;
          if (L.geom.periodic[o] || b.hi[o] != L.domain.hi[o]) P.hi[o] += 1;
        }
        IX_TRY(k::setval(P, vel.v(il, d), 1, scale * ns.p.bc_vals[side == 0 ? d : 3 + d][d], ns.s));
      }
    }
  return IAMRX_OK;
}

int project_simple(iamrx_ns_s& ns, MF& vel, const MF& sigma, MF& phi, MF* gp, int incr, int* iters, bool keep_dirichlet = false) {
  iamrx_mg_info mi = mg_info(ns, ns.p.proj_tol, ns.p.proj_abs_tol);
  const k::NodalBC nbc = ns.nodal_bc();
  IX_SOLVE(nodal_project(*ns.L, *ns.sv, vel, sigma, phi, gp, incr, &mi, ns.s, ns.walls ? &nbc : nullptr, keep_dirichlet));
  if (iters) *iters = mi.iters;
  return IAMRX_OK;
}

}  // namespace

extern "C" {

void iamrx_ns_params_default(iamrx_ns_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->cfl = 0.7;
  p->visc_coef = 0.0;
  p->scal_diff_coef = 0.0;
  p->be_cn_theta = 0.5;   // NSB.cpp:124
  p->change_max = 1.1;    // NSB.cpp:101
  p->init_shrink = 1.0;
  p->fixed_dt = -1.0;
  p->gravity = 0.0;
  p->visc_tol = 1.0e-10;  // NSB.cpp:122
  p->mac_tol = 1.0e-12; p->mac_abs_tol = 1.0e-16;    // MacProj.cpp:49-51
  p->proj_tol = 1.0e-12; p->proj_abs_tol = 1.0e-16;  // Projection.cpp:19-21
  p->init_iter = 2;       // NSB.cpp:98
  p->init_vel_iter = 1;   // NSB.cpp:99
  p->do_init_proj = 1;
  p->use_forces_in_trans = 0;
  p->verbose = 0;
  p->conservative_tracer = 0;
  p->mg_verbose = 0;
  p->godunov_ppm = 0;     // ns.advection_scheme = Godunov_PLM (NSB.cpp:169)
  p->do_scalminmax = 0;   // NSB.cpp:140
  p->do_mom_diff = 0;     // NSB.cpp:167
  p->bottom_solver = 0;   // smoother sweeps; 1 = BiCGStab as IAMR's "bicgcg"
}

int iamrx_ns_create(iamrx_level_t lev, const iamrx_ns_params* p, iamrx_ns_t* out) {
  IX_NEED_DEVICE();
  IX_ARG(lev && p && out, "null argument");
  IX_ARG(p->cfl > 0.0 && p->cfl <= 1.0, "ns.cfl must be in (0,1]");
  IX_ARG(p->be_cn_theta >= 0.5 && p->be_cn_theta <= 1.0, "ns.be_cn_theta must be in [0.5,1] (NSB.cpp:506-508)");
  IX_ARG(p->visc_coef >= 0.0, "ns.vel_visc_coef must be >= 0 (NS.cpp:2077)");
  IX_ARG(p->scal_diff_coef >= 0.0, "ns.scal_diff_coefs must be >= 0");
  Level* L = level_of(lev);
  bool walls = false;
  for (int d = 0; d < 3; ++d) {
    if (L->geom.periodic[d]) {
      IX_ARG(p->lo_bc[d] == 0 && p->hi_bc[d] == 0, "ns.lo_bc / ns.hi_bc must be 0 (interior) in a periodic direction");
    } else {
      walls = true;
      for (int v : {p->lo_bc[d], p->hi_bc[d]}) {
        IX_ARG(v >= 1 && v <= 5, "ns.lo_bc / ns.hi_bc of a non-periodic direction must be 1..5 (inputs.3d.taylorgreen:100-102)");
        // Projection::set_outflow_bcs (the hydrostatic pressure on outflow faces under gravity, Projection.cpp:1721-2379; called only
        // when gravity != 0 or have_divu: :309-324, 896-899) is implemented for side and top faces; like the reference (:1957) an
        // outflow face at the bottom under gravity is refused
        IX_ARG(!(d == 2 && p->lo_bc[2] == 2 && p->gravity != 0.0), "outflow at the bottom of the domain together with gravity (Projection::computeRhoG aborts)");
      }
    }
  }
  auto* ns = new iamrx_ns_s();
  ns->hlev = lev; ns->L = L; ns->sv = solvers_of(lev); ns->p = *p; ns->walls = walls;
  ns->S_old.define(L, IX_CELL, NUM_STATE, 1); ns->S_new.define(L, IX_CELL, NUM_STATE, 1);
  ns->P_old.define(L, IX_NODE, 1, 1); ns->P_new.define(L, IX_NODE, 1, 1);
  ns->Gp_old.define(L, IX_CELL, 3, 1); ns->Gp_new.define(L, IX_CELL, 3, 1);
  for (int d = 0; d < 3; ++d) ns->umac[d].define(L, IX_XFACE + d, 1, 1);
  ns->aofs.define(L, IX_CELL, NUM_STATE, 0);
  ns->rho_ptime.define(L, IX_CELL, 1, 1); ns->rho_ctime.define(L, IX_CELL, 1, 1); ns->rho_half.define(L, IX_CELL, 1, 1);
  ns->Umf.define(L, IX_CELL, 3, 3); ns->Smf.define(L, IX_CELL, NUM_SCALARS, 3);
  ns->visc.define(L, IX_CELL, 3, 1); ns->force.define(L, IX_CELL, 3, 1); ns->sforce.define(L, IX_CELL, NUM_SCALARS, 1);
  for (int d = 0; d < 3; ++d) ns->eta[d].define(L, IX_XFACE + d, 1, 0);
  ns->soln.define(L, IX_CELL, 3, 1); ns->rhs3.define(L, IX_CELL, 3, 0);
  ns->sig.define(L, IX_CELL, 1, 1); ns->mac_phi.define(L, IX_CELL, 1, 1); ns->tf0.define(L, IX_CELL, 3, 0);
  if (p->do_scalminmax) ns->smm.define(L, IX_CELL, NUM_SCALARS, 1);
  if (p->scal_diff_coef > 0.0) {
    for (int d = 0; d < 3; ++d) ns->seta[d].define(L, IX_XFACE + d, 1, 0);
    ns->s1.define(L, IX_CELL, 1, 1); ns->r1.define(L, IX_CELL, 1, 0); ns->svisc.define(L, IX_CELL, 1, 1); ns->ones.define(L, IX_CELL, 1, 1);
  }
  cudaStream_t s = ns->s;
  MF* all[] = {&ns->S_old, &ns->S_new, &ns->P_old, &ns->P_new, &ns->Gp_old, &ns->Gp_new, &ns->umac[0], &ns->umac[1],
               &ns->umac[2], &ns->aofs, &ns->rho_ptime, &ns->rho_ctime, &ns->rho_half, &ns->Umf, &ns->Smf, &ns->visc,
               &ns->force, &ns->sforce, &ns->soln, &ns->rhs3, &ns->sig, &ns->mac_phi, &ns->tf0};
  for (MF* m : all) {
    for (int il = 0; il < m->n(); ++il) if (!m->fabs[il].p) { delete ns; return IAMRX_ERR_CUDA; }
    int rc = mf_setval(*m, 0.0, 0, m->ncomp, m->ng, s);
    if (rc) { delete ns; return rc; }
  }
  for (int d = 0; d < 3; ++d) {  // constant viscosity on faces (NS.cpp:2062-2117 calcViscosity/getViscosity)
    int rc = mf_setval(ns->eta[d], p->visc_coef, 0, 1, 0, s);
    if (rc) { delete ns; return rc; }
  }
  if (p->scal_diff_coef > 0.0) {  // constant tracer diffusivity (calcDiffusivity/getDiffusivity NS.cpp:2084-2119); alpha = 1 for rho_flag 0
    int rc = IAMRX_OK;
    for (int d = 0; d < 3 && rc == IAMRX_OK; ++d) rc = mf_setval(ns->seta[d], p->scal_diff_coef, 0, 1, 0, s);
    if (rc == IAMRX_OK) rc = mf_setval(ns->ones, 1.0, 0, 1, 1, s);
    if (rc == IAMRX_OK) rc = mf_setval(ns->s1, 0.0, 0, 1, 1, s);
    if (rc == IAMRX_OK) rc = mf_setval(ns->svisc, 0.0, 0, 1, 1, s);
    if (rc == IAMRX_OK) rc = mf_setval(ns->r1, 0.0, 0, 1, 0, s);
    if (rc) { delete ns; return rc; }
  }
  *out = ns;
  return IAMRX_OK;
}

int iamrx_ns_destroy(iamrx_ns_t ns) {
  if (ns && ns->stage) cudaFreeHost(ns->stage);
  if (ns && ns->turb_dev) dev_free(ns->turb_dev);
#if !defined(IX_EMUL)
  if (ns && ns->s2) { cudaStreamDestroy(ns->s2); cudaEventDestroy(ns->ev_scal); cudaEventDestroy(ns->ev_in); cudaEventDestroy(ns->ev_prev); }
#endif
  delete ns;
  return IAMRX_OK;
}

int iamrx_ns_set_turbulent_forcing(iamrx_ns_t ns, int nmodes, int mode_start, int div_free_force, int array_size, const double* forcedata) {
  IX_NEED_DEVICE();
  IX_ARG(ns, "null argument");
  if (!forcedata || nmodes <= 0) { ns->turb_on = false; return IAMRX_OK; }
  IX_ARG(array_size > 0 && mode_start >= 0 && mode_start <= nmodes, "bad turbulent forcing parameters");
  for (int d = 0; d < 3; ++d) IX_ARG(ns->L->geom.periodic[d], "turbulent forcing needs a triply periodic domain");
  const size_t n = (size_t)17 * array_size * array_size * array_size;
  ns->turb_data.resize(n);
  // the table may live in host or device memory (TurbulentForcing_def.H:350-364 puts it in The_Arena on GPU builds)
  IX_CUDA(cudaMemcpy(ns->turb_data.data(), forcedata, n * sizeof(double), cudaMemcpyDefault));
  ns->turb = k::TurbParams{nmodes, mode_start, div_free_force ? 1 : 0, array_size};
  if (!ns->turbf.ok()) { ns->turbf.define(ns->L, IX_CELL, 3, 1); IX_TRY(mf_setval(ns->turbf, 0.0, 0, 3, 1, ns->s)); }
  // every mode index the loops can reach must lie inside the table
  double len[3];
  for (int d = 0; d < 3; ++d) len[d] = (ns->L->domain.hi[d] - ns->L->domain.lo[d] + 1) * ns->L->geom.dx[d];
  std::vector<k::TurbMode> probe;
  if (k::turb_modes(ns->turb, ns->turb_data.data(), len, 0.0, probe) != IAMRX_OK) {
    set_error("turbulent forcing: mode index beyond array_size");
    return IAMRX_ERR_ARG;
  }
  ns->turb_on = true;
  ns->turb_time_ok = false;
  return IAMRX_OK;
}

int iamrx_ns_init_prob(iamrx_ns_t ns, int probtype, const double* prob_params, int nparams) {
  IX_NEED_DEVICE();
  IX_ARG(ns && prob_params && nparams >= 0, "null argument");
  IX_ARG(probtype == 11 || probtype == 100 || probtype == 5 || probtype == 20 || probtype == 10 || probtype == 1 || probtype == 101,
         "probtype must be 11 (TaylorGreen), 5 (DoubleShearLayer), 10 (RayleighTaylor), 1 (LidDrivenCavity), 20 (HIT), 100 or 101");
  IX_ARG((probtype == 5 || probtype == 10) ? nparams >= 6 : (probtype == 20 ? nparams >= 2 : (probtype == 1 ? nparams >= 0 : (probtype == 101 ? nparams >= 3 : nparams >= 5))),
         "too few prob parameters");
  Level& L = *ns->L;
  // NavierStokes::initData (NS.cpp:335-460): S_new from prob_init, P_new = 0, Gp = 0
  for (int il = 0; il < ns->S_new.n(); ++il)
    IX_TRY(k::init_prob(L.lbox(il), ns->S_new.v(il), probtype, prob_params, nparams, L.geom, ns->s));
  IX_TRY(fill_state_ghosts(*ns, ns->S_new, 0, 0, NUM_STATE, 1));
  IX_TRY(mf_setval(ns->P_new, 0.0, 0, 1, 1, ns->s));
  IX_TRY(mf_setval(ns->P_old, 0.0, 0, 1, 1, ns->s));
  IX_TRY(mf_setval(ns->Gp_new, 0.0, 0, 3, 1, ns->s));
  IX_TRY(mf_setval(ns->Gp_old, 0.0, 0, 3, 1, ns->s));
  ns->time = 0.0; ns->nstep = 0; ns->dt_level = 0.0; ns->dt_min = 1.0e100;
  return IAMRX_OK;
}

// NavierStokes::post_init (NS.cpp:1254-1303): post_init_state (NSB.cpp:2369-2439),
// post_init_estDT (NSB.cpp:2307-2366), post_init_press (NS.cpp:1306-1432)
int iamrx_ns_post_init(iamrx_ns_t nsp, double* dt0) {
  IX_NEED_DEVICE();
  IX_ARG(nsp, "null argument");
  iamrx_ns_s& ns = *nsp;
  Level& L = *ns.L;
  // --- initialVelocityProject (Projection.cpp:615-838): sigma = 1, P/Gp reset to 0 afterwards
  if (ns.p.do_init_proj) {
    for (int it = 0; it < ns.p.init_vel_iter; ++it) {
      IX_TRY(mf_setval(ns.P_old, 0.0, 0, 1, 1, ns.s));
      IX_TRY(mf_setval(ns.sig, 1.0, 0, 1, 1, ns.s));
      MF vel; vel.alias(&L, IX_CELL, 3, 1, ns.S_new.fabs.data());
      IX_TRY(inflow_ghost_velocity(ns, vel, 1.0));
      IX_TRY(project_simple(ns, vel, ns.sig, ns.P_old, nullptr, 0, nullptr));
      IX_TRY(mf_setval(ns.P_old, 0.0, 0, 1, 1, ns.s));
      IX_TRY(mf_setval(ns.P_new, 0.0, 0, 1, 1, ns.s));
      IX_TRY(mf_setval(ns.Gp_old, 0.0, 0, 3, 1, ns.s));
      IX_TRY(mf_setval(ns.Gp_new, 0.0, 0, 3, 1, ns.s));
    }
  }
  ns.initial_step = true;  // NSB.cpp:2403
  if (ns.p.do_init_proj && std::fabs(ns.p.gravity) > 0.0 && ns.walls) {
    // initialPressureProject (NSB.cpp:2421-2431 -> Projection.cpp:841-960): projecting the uniform gravity vector with sigma = 1/rho
    // gives the hydrostatic pressure; P and Gradp old := new.  (On a fully periodic domain div(0,0,g) vanishes: nothing to do.)
    MF gvel(&L, IX_CELL, 3, 1);
    IX_TRY(mf_setval(gvel, 0.0, 0, 2, 1, ns.s));
    IX_TRY(mf_setval(gvel, ns.p.gravity, 2, 1, 1, ns.s));
    for (int il = 0; il < ns.sig.n(); ++il) IX_TRY(k::invert(L.lbox(il), ns.sig.v(il), ns.S_new.c(il, Density), ns.s));
    IX_TRY(mf_setval(ns.P_new, 0.0, 0, 1, 1, ns.s));
    IX_TRY(set_outflow_bcs(ns, ns.P_new, ns.S_new, Density));   // Projection.cpp:893-903 (INITIAL_PRESS)
    IX_TRY(project_simple(ns, gvel, ns.sig, ns.P_new, &ns.Gp_new, 0, nullptr, outflow_with_gravity(ns)));
    IX_TRY(fill_gradp(ns, ns.Gp_new));
    IX_TRY(mf_copy(ns.P_old, ns.P_new, 0, 0, 1, 1, ns.s));
    IX_TRY(mf_copy(ns.Gp_old, ns.Gp_new, 0, 0, 3, 1, ns.s));
  }
  // --- post_init_estDT: dt_init = init_shrink * estTimeStep
  double est = 0.0;
  IX_TRY(est_time_step(ns, &est));
  const double dt_init = ns.p.init_shrink * est;
  const double dt_save = dt_init;
  // --- post_init_press: init_iter x { advance ; initialSyncProject ; resetState }
  if (ns.p.init_iter > 0) {
    ns.initial_iter = true;
    for (int iter = 0; iter < ns.p.init_iter; ++iter) {
      double dt_test = 0.0;
      IX_TRY(advance(ns, ns.time, dt_init, &dt_test));   // NS.cpp:1348-1351
      // initialSyncProject (Projection.cpp:970-1185): vel = (U_new - U_old)/dt, sigma = 1/rho_half,
      // phi = P_old (zeroed scratch), Gp_new += grad phi, P_new += phi
      IX_TRY(mf_setval(ns.P_old, 0.0, 0, 1, 1, ns.s));
      for (int il = 0; il < ns.S_new.n(); ++il)
        IX_TRY(k::lincomb(L.lbox(il), ns.S_new.v(il, Xvel), 1.0 / dt_init, ns.S_new.c(il, Xvel), -1.0 / dt_init,
                          ns.S_old.c(il, Xvel), 3, ns.s));
      for (int il = 0; il < ns.sig.n(); ++il) IX_TRY(k::invert(L.lbox(il), ns.sig.v(il), ns.rho_half.c(il), ns.s));
      MF vel; vel.alias(&L, IX_CELL, 3, 1, ns.S_new.fabs.data());
      IX_TRY(inflow_ghost_velocity(ns, vel, 0.0));
      IX_TRY(project_simple(ns, vel, ns.sig, ns.P_old, &ns.Gp_new, 1, &ns.it_nodal));
      IX_TRY(mf_lincomb(ns.P_new, 0, 1.0, ns.P_new, 0, 1.0, ns.P_old, 0, 1, 1, ns.s));  // :1166-1170
      IX_TRY(fill_gradp(ns, ns.Gp_new));
      // resetState (NSB.cpp:2643-2680): State new := the time-n state; P, Gp old := new
      std::swap(ns.S_old, ns.S_new);
      IX_TRY(mf_copy(ns.P_old, ns.P_new, 0, 0, 1, 1, ns.s));
      IX_TRY(mf_copy(ns.Gp_old, ns.Gp_new, 0, 0, 3, 1, ns.s));
      ns.initial_iter = false;  // NS.cpp:1405
    }
  }
  ns.initial_step = false;  // NS.cpp:1408
  ns.dt_level = dt_save;
  ns.dt_min = 1.0e100;
  if (dt0) *dt0 = dt_save;
  return IAMRX_OK;
}

// Amr::coarseTimeStep for one level: computeNewDt (NSB.cpp:945-1036) unless a dt is
// forced, advance, bookkeeping of dt_min (Amr::timeStep)
int iamrx_ns_step(iamrx_ns_t nsp, double* dt_io) {
  IX_NEED_DEVICE();
  IX_ARG(nsp, "null argument");
  iamrx_ns_s& ns = *nsp;
  double dt = (dt_io && *dt_io > 0.0) ? *dt_io : -1.0;
  if (dt <= 0.0) {
    if (ns.p.fixed_dt > 0.0) dt = ns.p.fixed_dt;
    else if (ns.nstep == 0 && ns.dt_level > 0.0) dt = ns.dt_level;  // computeInitialDt: dt from post_init
    else {
      double est = 0.0;
      IX_TRY(est_time_step(ns, &est));
      dt = std::min(ns.dt_min, est);
      if (ns.dt_level > 0.0) dt = std::min(dt, ns.p.change_max * ns.dt_level);
    }
  }
  double dt_test = 0.0;
  IX_TRY(advance(ns, ns.time, dt, &dt_test));
  ns.dt_min = dt_test;
  ns.dt_level = dt;
  ns.time += dt;
  ns.nstep += 1;
  if (dt_io) *dt_io = dt;
  if (ns.p.verbose)
    fprintf(stderr, "[iamrx] step %d time %.6e dt %.6e iters mac/visc/nodal %d/%d/%d\n", ns.nstep, ns.time, dt, ns.it_mac,
            ns.it_visc, ns.it_nodal);
  return IAMRX_OK;
}

double iamrx_ns_time(iamrx_ns_t ns) { return ns ? ns->time : 0.0; }
int iamrx_ns_nstep(iamrx_ns_t ns) { return ns ? ns->nstep : 0; }

int iamrx_ns_field(iamrx_ns_t ns, int which, int il, iamrx_fab* out) {
  IX_ARG(ns && out, "null argument");
  IX_ARG(il >= 0 && il < ns->L->nlocal(), "local box index");
  MF* m = nullptr;
  switch (which) {
    case 0: m = &ns->S_new; break;
    case 1: m = &ns->P_new; break;
    case 2: m = &ns->Gp_new; break;
    case 3: m = &ns->S_old; break;
    case 4: case 5: case 6: m = &ns->umac[which - 4]; break;
    case 7: m = &ns->aofs; break;
    default: IX_ARG(false, "field selector");
  }
  *out = m->fabs[il];
  return IAMRX_OK;
}

#if !defined(IX_EMUL)
// dense host array (ncomp, nz, ny, nx) <-> the valid region of a ghosted device fab, as one strided DMA per component: no
// staging buffer and no pack / unpack kernel between the PCIe copy and the state (IAMRX_E2E_3D=0: the staged path)
static bool e2e_3d() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("IAMRX_E2E_3D"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}
// IAMRX_E2E_ZC=1: the pack / unpack kernels read and write the pinned host arrays directly over PCIe (zero copy) instead of the DMA
// engines' strided copies -- only when the pointer is pinned memory the device can address
static double* e2e_zero_copy(const void* host) {   // the device's address of the pinned host array, or nullptr
  static int on = -1;
  if (on < 0) { const char* e = getenv("IAMRX_E2E_ZC"); on = (e && e[0] == '1') ? 1 : 0; }
  if (!on) return nullptr;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return a.type == cudaMemoryTypeHost ? (double*)a.devicePointer : nullptr;
}
static int copy3d(const Bx& b, V4 dev, int dcomp, double* host, int ncomp, bool to_device, cudaStream_t s) {
  if (double* hz = e2e_zero_copy(host)) {
    V4 d = dev; d.p = dev.p + dcomp * dev.ns;
    if (to_device) return k::unpack(b, d, hz, ncomp, s);
    return k::pack(b, hz, C4{d.p, d.l0, d.l1, d.l2, d.js, d.ks, d.ns}, ncomp, s);
  }
  if (dev.js <= 0 || dev.ks % dev.js != 0) return IAMRX_ERR_ARG;
  const size_t nx = (size_t)b.nx(), ny = (size_t)b.ny(), nz = (size_t)b.nz();
  for (int n = 0; n < ncomp; ++n) {
    double* d = dev.p + (dcomp + n) * dev.ns + ((b.lo[0] - dev.l0) + (b.lo[1] - dev.l1) * dev.js + (b.lo[2] - dev.l2) * dev.ks);
    cudaMemcpy3DParms p{};
    const cudaPitchedPtr hp = make_cudaPitchedPtr(host + (size_t)n * nx * ny * nz, nx * sizeof(double), nx * sizeof(double), ny);
    const cudaPitchedPtr dp = make_cudaPitchedPtr(d, (size_t)dev.js * sizeof(double), (size_t)dev.js * sizeof(double), (size_t)(dev.ks / dev.js));
    p.srcPtr = to_device ? hp : dp;
    p.dstPtr = to_device ? dp : hp;
    p.extent = make_cudaExtent(nx * sizeof(double), ny, nz);
    p.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    IX_CUDA(cudaMemcpy3DAsync(&p, s));
  }
  return IAMRX_OK;
}
#endif

int iamrx_ns_step_host(iamrx_ns_t nsp, const double* const* host_in, double* const* host_out, double* dt_io) {
  IX_NEED_DEVICE();
  IX_ARG(nsp && host_in && host_out, "null argument");
  iamrx_ns_s& ns = *nsp;
  Level& L = *ns.L;
  size_t maxpts = 0;
  for (int il = 0; il < L.nlocal(); ++il) maxpts = std::max(maxpts, (size_t)L.lbox(il).npts());
  const size_t need = maxpts * NUM_STATE;
  struct Guard { double* p; ~Guard() { dev_free(p); } } dg{dev_alloc(need)}, eg{nullptr};   // released on every exit path
  double* dbuf = dg.p;
  if (!dbuf) return IAMRX_ERR_CUDA;
  bool early = false;
#if !defined(IX_EMUL)
  early = L.nlocal() == 1 && !ns.initial_step && !ns.initial_iter;
  if (early && !ns.s2) {
    IX_CUDA(cudaStreamCreateWithFlags(&ns.s2, cudaStreamNonBlocking));
    IX_CUDA(cudaEventCreateWithFlags(&ns.ev_scal, cudaEventDisableTiming));
    IX_CUDA(cudaEventCreateWithFlags(&ns.ev_in, cudaEventDisableTiming));
    IX_CUDA(cudaEventCreateWithFlags(&ns.ev_prev, cudaEventDisableTiming));
  }
#endif
  // scalminmax / a diffusive tracer / momentum form read the tracer no earlier either, but keep the overlap to the plain path
  const bool late = early && Tracer == NUM_STATE - 1;
  for (int il = 0; il < L.nlocal(); ++il) {
    const size_t npts = (size_t)L.lbox(il).npts();
    const int nfirst = late ? NUM_STATE - 1 : NUM_STATE;
#if !defined(IX_EMUL)
    if (e2e_3d()) {
      IX_TRY(copy3d(L.lbox(il), ns.S_new.v(il), 0, const_cast<double*>(host_in[il]), nfirst, true, ns.s));
    } else
#endif
    {
      IX_CUDA(cudaMemcpyAsync(dbuf, host_in[il], npts * nfirst * sizeof(double), cudaMemcpyHostToDevice, ns.s));
      IX_TRY(k::unpack(L.lbox(il), ns.S_new.v(il), dbuf, nfirst, ns.s));
    }
#if !defined(IX_EMUL)
    if (late) {   // the tracer follows on the second stream, after everything already queued on the main stream (previous readers)
      double* tb = dbuf + npts * (NUM_STATE - 1);
      IX_CUDA(cudaEventRecord(ns.ev_prev, ns.s));
      IX_CUDA(cudaStreamWaitEvent(ns.s2, ns.ev_prev, 0));
      if (e2e_3d()) {
        IX_TRY(copy3d(L.lbox(il), ns.S_new.v(il), Tracer, const_cast<double*>(host_in[il]) + npts * (NUM_STATE - 1), 1, true, ns.s2));
      } else {
        IX_CUDA(cudaMemcpyAsync(tb, host_in[il] + npts * (NUM_STATE - 1), npts * sizeof(double), cudaMemcpyHostToDevice, ns.s2));
        IX_TRY(k::unpack(L.lbox(il), ns.S_new.v(il, Tracer), tb, 1, ns.s2));
      }
      IX_CUDA(cudaEventRecord(ns.ev_in, ns.s2));
      ns.late_tracer = true;
    }
#endif
  }
#if !defined(IX_EMUL)
  if (early) {
    eg.p = ns.early_buf = dev_alloc(maxpts * NUM_SCALARS);
    if (!ns.early_buf) return IAMRX_ERR_CUDA;
    ns.early_out = host_out[0] + (size_t)L.lbox(0).npts() * Density;
  }
#endif
  int rc = iamrx_ns_step(nsp, dt_io);
  ns.early_out = nullptr;
  ns.early_buf = nullptr;
  ns.late_tracer = false;
  if (rc != IAMRX_OK) {
#if !defined(IX_EMUL)
    if (early) cudaStreamSynchronize(ns.s2);
#endif
    return rc;
  }
  const int nout = early ? Density : NUM_STATE;   // velocity only when the scalars already left
  for (int il = 0; il < L.nlocal(); ++il) {
    const size_t n = (size_t)L.lbox(il).npts() * nout;
#if !defined(IX_EMUL)
    if (e2e_3d()) { IX_TRY(copy3d(L.lbox(il), ns.S_new.v(il), 0, host_out[il], nout, false, ns.s)); continue; }
#endif
    IX_TRY(k::pack(L.lbox(il), dbuf, ns.S_new.c(il), nout, ns.s));
    IX_CUDA(cudaMemcpyAsync(host_out[il], dbuf, n * sizeof(double), cudaMemcpyDeviceToHost, ns.s));
  }
  IX_CUDA(cudaStreamSynchronize(ns.s));
#if !defined(IX_EMUL)
  if (early) IX_CUDA(cudaStreamSynchronize(ns.s2));
#endif
  return IAMRX_OK;
}

// Amr::writePlotFile for the level (NavierStokesBase.cpp:3343-3352, NavierStokes.cpp:1080-1196): the cell-centred state components
// AmrLevel::setPlotVariables registers (NS_setup.cpp:253-269 State_Type, :339-360 Gradp_Type), new-time data
int iamrx_ns_write_plotfile(iamrx_ns_t nsp, const char* dir) {
  IX_NEED_DEVICE();
  IX_ARG(nsp && dir && dir[0], "null argument");
  iamrx_ns_s& ns = *nsp;
  const std::vector<std::string> names = {"x_velocity", "y_velocity", "z_velocity", "density", "tracer", "gradpx", "gradpy", "gradpz"};
  return write_plotfile(*ns.L, {&ns.S_new, &ns.Gp_new}, {0, 0}, {NUM_STATE, 3}, names, dir, "NavierStokes-V1.1", ns.time, ns.nstep, ns.s);
}

// MacProj::mac_sync_compute (MacProj.cpp:505-731) for the level this object advances, right after iamrx_ns_step: see iamrx.h
int iamrx_ns_mac_sync_compute(iamrx_ns_t nsp, const iamrx_fab* ucorr, const iamrx_fab* vcorr, const iamrx_fab* wcorr, iamrx_fab* vsync,
                              iamrx_fab* ssync, double dt) {
  IX_NEED_DEVICE();
  IX_ARG(nsp && ucorr && vcorr && wcorr && vsync && dt > 0.0, "mac_sync_compute arguments");
  iamrx_ns_s& ns = *nsp;
  IX_ARG(ns.nstep > 0, "mac_sync_compute follows a time step (it reuses that step's FillPatched state and forcing)");
  Level& L = *ns.L;
  const iamrx_fab* uc[3] = {ucorr, vcorr, wcorr};
  MF U[3];
  for (int d = 0; d < 3; ++d) U[d].alias(&L, IX_XFACE + d, 1, 0, const_cast<iamrx_fab*>(uc[d]));
  // the state at prev_time, its forcing (visc - grad p + body force, /rho unless do_mom_diff; scalar forcing) and u_mac are the
  // ones the step's own advection calls used: ns.Umf / ns.Smf / ns.force / ns.sforce / ns.umac still hold them (:519-660)
  MF Vs; Vs.alias(&L, IX_CELL, 3, 0, vsync);
  IX_TRY(compute_aofs(ns, Xvel, 3, ns.Umf, &ns.force, true, dt, &Vs, 0, U));                       // :681-691
  if (ssync) {
    MF Ss; Ss.alias(&L, IX_CELL, NUM_SCALARS, 0, ssync);
    IX_TRY(compute_aofs(ns, Density, NUM_SCALARS, ns.Smf, &ns.sforce, false, dt, &Ss, 0, U));    // :695-707
  }
  return IAMRX_OK;
}

int iamrx_ns_last_iters(iamrx_ns_t ns, int iters[3]) {
  IX_ARG(ns && iters, "null argument");
  iters[0] = ns->it_mac; iters[1] = ns->it_visc; iters[2] = ns->it_nodal;
  return IAMRX_OK;
}

int iamrx_ns_sum_integrated_quantities(iamrx_ns_t nsp, double out[3]) {
  IX_NEED_DEVICE();
  IX_ARG(nsp && out, "null argument");
  iamrx_ns_s& ns = *nsp;
  Level& L = *ns.L;
  const double vol = L.geom.dx[0] * L.geom.dx[1] * L.geom.dx[2];
  double mass = 0.0, trac = 0.0, ke = 0.0;
  IX_TRY(mf_sum(ns.S_new, Density, &mass, ns.s));
  IX_TRY(mf_sum(ns.S_new, Tracer, &trac, ns.s));
  // derkeng: 0.5 rho (u^2 + v^2 + w^2), built in the estTimeStep scratch from existing pointwise kernels
  IX_TRY(mf_copy(ns.tf0, ns.S_new, Xvel, 0, 3, 0, ns.s));
  for (int il = 0; il < ns.tf0.n(); ++il) {
    const Bx& b = L.lbox(il);
    IX_TRY(k::mult(b, ns.tf0.v(il), ns.tf0.c(il), 3, 3, ns.s));                                        // squares
    IX_TRY(k::lincomb(b, ns.tf0.v(il, 0), 1.0, ns.tf0.c(il, 0), 1.0, ns.tf0.c(il, 1), 1, ns.s));
    IX_TRY(k::lincomb(b, ns.tf0.v(il, 0), 1.0, ns.tf0.c(il, 0), 1.0, ns.tf0.c(il, 2), 1, ns.s));
    IX_TRY(k::mult(b, ns.tf0.v(il, 0), ns.S_new.c(il, Density), 1, 1, ns.s));
  }
  IX_TRY(mf_sum(ns.tf0, 0, &ke, ns.s));
  out[0] = mass * vol; out[1] = trac * vol; out[2] = 0.5 * ke * vol;
  return IAMRX_OK;
}

}  // extern "C"
