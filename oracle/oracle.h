/*
 * oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (plain C++17 + OpenMP, fp64) of the IAMR hot path on ONE
 * fully periodic box: Godunov PLM advection, ABec/tensor operators, cell and
 * nodal multigrid, MAC and nodal projections and the NavierStokes::advance
 * sequence.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load it; nothing under iamr_b200/ does.
 *
 * PARITY UNPINNED at the AMReX / AMReX-Hydro boundary: the kernel arithmetic of
 * this path lives in AMReX (Src/LinearSolvers/MLMG, Src/Base) and AMReX-Hydro
 * (Godunov/, Utils/, Projections/), which /root/reference does not vendor
 * (Exec/Make.IAMR:15-19,36-37; .SUBMODULES.json "submodules": []) and whose
 * versions are not pinned (Test/IAMR-tests.ini:63-76 track development/main;
 * tree date => AMReX ~22.12, AMReX-Hydro main of Dec 2022).  The oracle restates
 * their published algorithms (Almgren, Bell, Colella, Howell, Welcome, JCP 142
 * (1998)) and follows IAMR's own call sites for order,
 * arguments, scalings and tolerances.  It is pinned only by the reference's
 * analytic known-answer (Tutorials/TaylorGreen/benchmarks/EXACT_3D.F:75,114-118,
 * checked to 2nd-order convergence in tests/test_oracle.py) and by discrete
 * identities; the reference ships no golden plotfiles (Test/README.md:23-29).
 *
 * Data layout: every field is a dense periodic array WITHOUT ghost cells,
 * [comp][k][j][i], i fastest, n = {nx, ny, nz} cells.  Face-centred arrays hold
 * the LOW face of each cell (face index i = between cells i-1 and i); nodal
 * arrays hold the LOW corner node of each cell; both are n-periodic.
 */
#ifndef IAMR_ORACLE_H_
#define IAMR_ORACLE_H_
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_mg {
  double rtol, atol;
  int max_iter, nu1, nu2, bottom_sweeps, max_coarsening;
  double omega;
  /* out */
  int iters;
  double resnorm0, resnorm, rhsnorm;
  /* bottom solver: 0 = bottom_sweeps smoother sweeps, 1 = BiCGStab to bottom_rtol (MLCGSolver; IAMR's default "bicgcg") */
  int bottom_solver, bottom_maxiter;
  double bottom_rtol;
  int bottom_iters, pad_;
} orc_mg;
void orc_mg_default(orc_mg* m);

/* MLABecLaplacian (a*alpha - b div beta grad), SURVEY.md A.6.  bncomp = 1 or ncomp. */
void orc_abec_apply(const int n[3], const double dxinv[3], double a, double b, const double* alpha,
                    const double* bx, const double* by, const double* bz, int ncomp, int bncomp,
                    const double* phi, double* out);
void orc_abec_gsrb(const int n[3], const double dxinv[3], double a, double b, const double* alpha,
                   const double* bx, const double* by, const double* bz, int ncomp, int bncomp,
                   const double* rhs, double omega, int redblack, double* phi);
/* MLTensorOp cross terms: out += b * div(F_cross(eta, vel)), SURVEY.md A.8 */
void orc_tensor_cross(const int n[3], const double dxinv[3], double b, const double* ex, const double* ey,
                      const double* ez, const double* vel, double* out);
/* (a*alpha - b div eta grad [+ tensor]) soln = rhs by multigrid; returns 0 / iterations if not converged */
int orc_diffusion_solve(const int n[3], const double dx[3], int tensor, int ncomp, double a, double b,
                        const double* alpha, const double* ex, const double* ey, const double* ez,
                        const double* rhs, double* soln, orc_mg* mg);
void orc_diffusion_apply(const int n[3], const double dx[3], int tensor, int ncomp, double a, double b,
                         const double* alpha, const double* ex, const double* ey, const double* ez,
                         const double* soln, double* out);
/* MacProj::mlmg_mac_solve + Hydro::MacProjector::project (MacProj.cpp:1084-1184) */
int orc_mac_project(const int n[3], const double dx[3], double* umac, double* vmac, double* wmac,
                    const double* rho, const double* rhs, double* phi, double rhs_scale, orc_mg* mg);
/* MLNodeLaplacian pieces, SURVEY.md A.9 */
void orc_nodal_divu(const int n[3], const double dxinv[3], const double* vel, double* rhs);
void orc_nodal_adotx(const int n[3], const double dxinv[3], const double* sigma, const double* phi, double* out);
void orc_nodal_gs(const int n[3], const double dxinv[3], const double* sigma, const double* rhs, int color,
                  double* phi);
void orc_nodal_mknewu(const int n[3], const double dxinv[3], const double* sigma, const double* phi,
                      double* vel, double* gp);
/* Projection::doMLMGNodalProjection (Projection.cpp:2385-2567) */
int orc_nodal_project(const int n[3], const double dx[3], double* vel, const double* sigma, double* phi,
                      double* gp, int increment_gp, orc_mg* mg);
/* Godunov::ExtrapVelToFaces (NSB.cpp:4487-4491); forces_in_trans is a flag word: bit 0 use_forces_in_trans, bit 1 PPM instead of PLM */
void orc_extrap_vel_to_faces(const int n[3], const double dx[3], double dt, const double* vel,
                             const double* force, int forces_in_trans, double* umac, double* vmac,
                             double* wmac);
/* NavierStokesBase::ComputeAofs body (NSB.cpp:4661-4845); fx..zed may be NULL */
void orc_compute_aofs(const int n[3], const double dx[3], double dt, int ncomp, const double* S,
                      const double* force, const double* divu, const double* umac, const double* vmac,
                      const double* wmac, const int* iconserv, int forces_in_trans, double* aofs,
                      double* fx, double* fy, double* fz, double* xed, double* yed, double* zed);

/* the same with the complete argument list of the call site NSB.cpp:4701-4717: flux velocities (NULL = the MAC velocities,
 * U_corr in the sync call), is_sync (aofs in/out: aofs -= update, no convective term), known edge states (xed..zed inputs) */
void orc_compute_aofs2(const int n[3], const double dx[3], double dt, int ncomp, const double* S, const double* force,
                       const double* divu, const double* umac, const double* vmac, const double* wmac, const double* uflux,
                       const double* vflux, const double* wflux, const int* iconserv, int flags, int is_sync, int known,
                       double* aofs, double* fx, double* fy, double* fz, double* xed, double* yed, double* zed);

/* ---- non-periodic domains.  Arrays are exchanged "padded" = with their ghost layers, [comp][nz+2g][ny+2g][nx+2g]; the high
 * face / node of a non-periodic direction sits at index n inside the layer.  bclo/bchi: amrex::BCType codes [ncomp][3]
 * (int_dir 0, reflect_odd -1, reflect_even 1, foextrap 2, ext_dir 3, hoextrap 4); lobc/hibc: LinOpBCType (periodic 0,
 * Dirichlet 1, Neumann 2, reflect_odd 3, inflow 4). ---- */
/* AmrLevel::FillPatch's physical-boundary fill of cell data (filcc + NS_bcfill.H ext_dir values bcv[6][ncomp]) */
void orc_fill_physbc(const int n[3], const int per[3], int ng, int ncomp, const int* bclo, const int* bchi, const double* bcv, double* a);
void orc_extrap_vel_to_faces_bc(const int n[3], const int per[3], const double dx[3], double dt, const double* vel, const double* force,
                                int flags, const int* bclo, const int* bchi, double* umac, double* vmac, double* wmac);
void orc_compute_aofs_bc(const int n[3], const int per[3], const double dx[3], double dt, int ncomp, const double* S, const double* force,
                         const double* divu, const double* umac, const double* vmac, const double* wmac, const int* iconserv, int flags,
                         const int* bclo, const int* bchi, double* aofs, double* fx, double* fy, double* fz, double* xed, double* yed, double* zed);
int orc_mac_project_bc(const int n[3], const int per[3], const double dx[3], double* umac, double* vmac, double* wmac, const double* rho,
                       const double* rhs, double* phi, double rhs_scale, const int lobc[3], const int hibc[3], int maxorder, orc_mg* mg);
int orc_nodal_project_bc(const int n[3], const int per[3], const double dx[3], double* vel, const double* sigma, double* phi, double* gp,
                         const int lobc[3], const int hibc[3], orc_mg* mg);
int orc_diffusion_bc(const int n[3], const int per[3], const double dx[3], int solve, int tensor, int ncomp, double a, double b,
                     const double* alpha, const double* ex, const double* ey, const double* ez, const double* rhs, double* soln, double* out,
                     const int* lobc, const int* hibc, int maxorder, orc_mg* mg);

/* two-level transfer operators and the coarse-fine flux register on a periodic coarse domain nc refined by 2 */
void orc_average_down(const int nc[3], int ncomp, int ixtype, const double* fine, double* crse);
void orc_interp(int kind, const int nc[3], int ncomp, const double* crse, double* fine);
void orc_fluxreg(const int nc[3], int ncomp, const unsigned char* mask, const double* cfx, const double* cfy, const double* cfz,
                 const double* ffx, const double* ffy, const double* ffz, double dt, double vol, double* reg);

/* NavierStokes::advance / post_init on one box (periodic unless orc_ns_set_bc is called) */
typedef struct orc_ns_params {
  double cfl, visc_coef, be_cn_theta, change_max, init_shrink, fixed_dt, gravity, visc_tol;
  double mac_tol, mac_abs_tol, proj_tol, proj_abs_tol;
  int init_iter, init_vel_iter, do_init_proj, use_forces_in_trans, conservative_tracer, verbose;
  double scal_diff_coef;   /* ns.scal_diff_coefs of the tracer (0: non-diffusive) */
  int use_ppm;             /* ns.advection_scheme = Godunov_PPM (NSB.cpp:552-554, 4485) */
  int do_scalminmax;       /* ns.do_scalminmax (NSB.cpp:2907-2935) */
  int do_mom_diff;         /* ns.do_mom_diff (NSB.cpp:3358-3470, 3609-3616; NS.cpp:606-623, 1016) */
  int bottom_solver;       /* orc_mg.bottom_solver of the three solves */
} orc_ns_params;
void orc_ns_params_default(orc_ns_params* p);
typedef struct orc_ns orc_ns;
orc_ns* orc_ns_create(const int n[3], const double prob_lo[3], const double prob_hi[3], const orc_ns_params* p);
void orc_ns_set_bc(orc_ns* ns, const int per[3], const int phys_lo[3], const int phys_hi[3], const double* bcv);
void orc_ns_get_padded(const orc_ns* ns, int which, double* out);
void orc_ns_destroy(orc_ns* ns);
/* Tutorials/HIT turbulent forcing (NS_getForce.cpp:205-640, exact path); forcedata = TurbulentForcing::forcedata (17 x as^3) */
void orc_ns_set_turbulent_forcing(orc_ns* ns, int nmodes, int mode_start, int div_free_force, int array_size, const double* forcedata);
void orc_ns_init_prob(orc_ns* ns, int probtype, const double* params, int nparams);
int orc_ns_post_init(orc_ns* ns, double* dt0);
int orc_ns_step(orc_ns* ns, double* dt_io);
double orc_ns_time(const orc_ns* ns);
/* which: 0 state(5), 1 press(1, nodal), 2 gradp(3), 4..6 umac, 7 aofs(5) */
void orc_ns_get(const orc_ns* ns, int which, double* out);
void orc_ns_set_state(orc_ns* ns, const double* state5);
void orc_ns_last_iters(const orc_ns* ns, int iters[3]);
/* UNVERIFIED-UPSTREAM switches (ids = IAMRX_OPT_* of include/iamrx.h): 0 small_vel, 1 slope order (4|2), 2 corner form (0|1), 3 ext_dir both states */
int orc_set_option(int opt, double value);
int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
