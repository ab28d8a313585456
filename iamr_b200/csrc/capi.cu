// capi.cu -- the extern "C" surface of libiamrx (include/iamrx.h), sections 0-3:
// diagnostics, per-box kernels, communicator, level objects, FillBoundary and
// the level solvers (MAC projection, nodal projection, diffusion).  Section 4
// (the NavierStokes::advance driver) lives in ns.cu.
//
// Every entry point validates its arguments, refuses to run without a CUDA
// device (there is no CPU path) and converts the POD views to the internal
// V4/C4 types; nothing throws across the boundary.
#include "mlmg.h"
#include "solvers.h"

using namespace ix;

namespace ix {
const char* last_error_cstr();
int comm_unique_id(unsigned char uid[128]);
int comm_init(int rank, int nranks, const unsigned char uid[128]);
int comm_finalize();
int prof_enable(int on, int64_t min_points);
int prof_all(int on);
int prof_dump(char* buf, int cap);
void prof_reset();
int prof_report(int kclass, double* total_ms, int64_t* launches, double* algo_bytes);
int comm_set_transport(int rank, int nranks, iamrx_exchange_fn ex, iamrx_allreduce_fn ar, void* ctx);
}  // namespace ix

struct iamrx_level_s {
  std::unique_ptr<Level> lev;
  LevelSolvers solvers;
};

#define IX_TRY(call) do { int rc_ = (call); if (rc_ != IAMRX_OK) return rc_; } while (0)
#define IX_GUARD_BEGIN try {
#define IX_GUARD_END                                                            \
  } catch (const std::exception& e) {                                           \
    ix::set_error(std::string("exception: ") + e.what());                       \
    return IAMRX_ERR_ARG;                                                       \
  } catch (...) {                                                               \
    ix::set_error("unknown exception");                                         \
    return IAMRX_ERR_ARG;                                                       \
  }

static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }

static k::Abec make_abec(double a, double b, const iamrx_fab* acoef, const iamrx_fab* bx,
                         const iamrx_fab* by, const iamrx_fab* bz, const double dxinv[3], int ncomp) {
  k::Abec op;
  op.a = a; op.b = b;
  op.acoef = cview(acoef);
  op.bx = cview(bx); op.by = cview(by); op.bz = cview(bz);
  op.bncomp = (bx && bx->ncomp >= ncomp && ncomp > 1) ? ncomp : 1;
  for (int d = 0; d < 3; ++d) op.dxinv[d] = dxinv[d];
  return op;
}

extern "C" {

const char* iamrx_last_error(void) { return ix::last_error_cstr(); }
int iamrx_set_option(int option, double value) {
  const int rc = k::godunov_set_option(option, value);
  if (rc == IAMRX_ERR_ARG) ix::set_error("bad argument: unknown option or value out of range");
  return rc;
}
double iamrx_get_option(int option) { return k::godunov_get_option(option); }
int iamrx_version(void) { return 100; }
void iamrx_debug_fb_stats(int64_t out[4], int reset) {
  for (int q = 0; q < 4; ++q) { if (out) out[q] = ix::g_fb_stats[q].load(); if (reset) ix::g_fb_stats[q].store(0); }
}
int64_t iamrx_launch_count(void) { return g_launches.load(); }
void iamrx_launch_count_reset(void) { g_launches.store(0); }
int iamrx_device_ok(void) { return device_ok() ? 1 : 0; }
int iamrx_debug_fp64_peak(double* dp_ginstr_per_s, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(dp_ginstr_per_s, "null argument");
  return k::fp64_peak(dp_ginstr_per_s, (cudaStream_t)stream);
}
int iamrx_prof_enable(int on, int64_t min_points) { return prof_enable(on, min_points); }
void iamrx_prof_reset(void) { prof_reset(); }
int iamrx_prof_all(int on) { return prof_all(on); }
int iamrx_prof_dump(char* buf, int cap) { return prof_dump(buf, cap); }
int iamrx_prof_report(int kclass, double* total_ms, int64_t* launches, double* algo_bytes) {
  IX_ARG(kclass >= 0 && kclass < IAMRX_PROF_NCLASS, "kernel class");
  return prof_report(kclass, total_ms, launches, algo_bytes);
}

// ---------------------------------------------------------------------------
// 1. per-box kernels
// ---------------------------------------------------------------------------
int iamrx_abec_gsrb_box(const iamrx_box* bx, iamrx_fab* phi, const iamrx_fab* rhs, double a, double b,
                        const iamrx_fab* acoef, const iamrx_fab* bcoef_x, const iamrx_fab* bcoef_y,
                        const iamrx_fab* bcoef_z, const double dxinv[3], double omega, int redblack,
                        int ncomp, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(bx && phi && rhs && bcoef_x && bcoef_y && bcoef_z && dxinv, "null argument");
  IX_ARG(a == 0.0 || (acoef && acoef->p), "acoef required when a != 0");
  IX_ARG(ncomp >= 1 && ncomp <= phi->ncomp, "ncomp");
  return k::abec_gsrb(mkbx(*bx), view(phi), cview(rhs), make_abec(a, b, acoef, bcoef_x, bcoef_y, bcoef_z, dxinv, ncomp),
                      omega, redblack, ncomp, S(stream));
}

int iamrx_abec_gsrb_sweep_box(const iamrx_box* bx, iamrx_fab* phi_out, const iamrx_fab* phi_in, const iamrx_fab* rhs,
                              double a, double b, const iamrx_fab* acoef, const iamrx_fab* bcoef_x,
                              const iamrx_fab* bcoef_y, const iamrx_fab* bcoef_z, const double dxinv[3], double omega,
                              int ncomp, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(bx && phi_out && phi_in && rhs && bcoef_x && bcoef_y && bcoef_z && dxinv, "null argument");
  IX_ARG(phi_out->p != phi_in->p, "the fused sweep is out of place: phi_out and phi_in must differ");
  IX_ARG(a == 0.0 || (acoef && acoef->p), "acoef required when a != 0");
  IX_ARG(ncomp >= 1 && ncomp <= phi_in->ncomp && ncomp <= phi_out->ncomp, "ncomp");
  IX_ARG(k::abec_gsrb_sweep_ok(mkbx(*bx), 7), "fused sweep needs even extents >= 8 in every direction");
  return k::abec_gsrb_sweep(mkbx(*bx), view(phi_out), cview(phi_in), cview(rhs),
                            make_abec(a, b, acoef, bcoef_x, bcoef_y, bcoef_z, dxinv, ncomp), omega, 0, ncomp, S(stream));
}

int iamrx_abec_apply_box(const iamrx_box* bx, iamrx_fab* out, const iamrx_fab* phi, const iamrx_fab* rhs,
                         double a, double b, const iamrx_fab* acoef, const iamrx_fab* bcoef_x,
                         const iamrx_fab* bcoef_y, const iamrx_fab* bcoef_z, const double dxinv[3],
                         int ncomp, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(bx && out && phi && bcoef_x && bcoef_y && bcoef_z && dxinv, "null argument");
  IX_ARG(a == 0.0 || (acoef && acoef->p), "acoef required when a != 0");
  return k::abec_apply(mkbx(*bx), view(out), cview(phi), cview(rhs),
                       make_abec(a, b, acoef, bcoef_x, bcoef_y, bcoef_z, dxinv, ncomp), ncomp, S(stream));
}

int iamrx_tensor_cross_box(const iamrx_box* bx, iamrx_fab* out, const iamrx_fab* vel, const iamrx_fab* eta_x,
                           const iamrx_fab* eta_y, const iamrx_fab* eta_z, double b, const double dxinv[3],
                           void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(bx && out && vel && eta_x && eta_y && eta_z && dxinv, "null argument");
  IX_ARG(vel->ncomp >= 3 && out->ncomp >= 3, "tensor operator needs 3 components");
  return k::tensor_cross(mkbx(*bx), view(out), cview(vel), cview(eta_x), cview(eta_y), cview(eta_z), b, dxinv,
                         S(stream));
}

// does `f` (all of its allocation) cover box `need`?
static bool covers(const iamrx_fab* f, const Bx& need) {
  for (int d = 0; d < 3; ++d) if (f->lo[d] > need.lo[d] || f->hi[d] < need.hi[d]) return false;
  return true;
}
static k::AdvBC make_advbc(const iamrx_bcrec* bc, int ncomp, const iamrx_geom* geom) {
  k::AdvBC b{};
  for (int d = 0; d < 3; ++d) { b.dlo[d] = geom->domain.lo[d]; b.dhi[d] = geom->domain.hi[d]; }
  if (bc)
    for (int n = 0; n < ncomp && n < 8; ++n)
      for (int d = 0; d < 3; ++d) { b.lo[n][d] = bc[n].lo[d]; b.hi[n][d] = bc[n].hi[d]; }
  return b;
}
static bool bc_codes_ok(const iamrx_bcrec* bc, int ncomp) {
  if (!bc) return true;
  for (int n = 0; n < ncomp; ++n)
    for (int d = 0; d < 3; ++d)
      for (int v : {bc[n].lo[d], bc[n].hi[d]})
        if (v < IAMRX_BC_REFLECT_ODD || v > IAMRX_BC_HOEXTRAP) return false;
  return true;
}

int iamrx_extrap_vel_to_faces_box(const iamrx_box* bx, const iamrx_fab* vel, const iamrx_fab* force,
                                  iamrx_fab* umac, iamrx_fab* vmac, iamrx_fab* wmac, const iamrx_bcrec* bcrec,
                                  const iamrx_geom* geom, double dt, int flags, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(bx && vel && umac && vmac && wmac && geom, "null argument");
  IX_ARG(vel->ncomp >= 3, "vel needs 3 components");
  IX_ARG(bc_codes_ok(bcrec, 3), "unknown BCRec code");
  const Bx b = mkbx(*bx);
  IX_ARG(covers(vel, grow(b, 3)), "vel must cover the box grown by 3 cells (nghost_state, NSB.cpp:4539-4552)");
  IX_ARG(!force || !force->p || (force->ncomp >= 3 && covers(force, grow(b, 1))), "force must have 3 components on the box grown by 1");
  IX_ARG(covers(umac, surrounding(b, 0)) && covers(vmac, surrounding(b, 1)) && covers(wmac, surrounding(b, 2)), "u_mac must cover the faces of the box");
  k::AdvGeom g;
  for (int d = 0; d < 3; ++d) g.dx[d] = geom->dx[d];
  g.dt = dt;
  const k::AdvBC abc = make_advbc(bcrec, 3, geom);
  return k::extrap_vel_to_faces(b, cview(vel), cview(force), view(umac), view(vmac), view(wmac), g,
                                (flags & IAMRX_ADV_FORCES_IN_TRANS) ? 1 : 0, S(stream), (flags & IAMRX_ADV_PPM) ? 1 : 0, &abc);
}

int iamrx_compute_aofs_box(const iamrx_box* bx, iamrx_fab* aofs, int aofs_comp, const iamrx_fab* Sf, int s_comp,
                           int ncomp, const iamrx_fab* force, int f_comp, const iamrx_fab* divu,
                           const iamrx_fab* umac, const iamrx_fab* vmac, const iamrx_fab* wmac,
                           const iamrx_fab* uflux, const iamrx_fab* vflux, const iamrx_fab* wflux, iamrx_fab* fx,
                           iamrx_fab* fy, iamrx_fab* fz, iamrx_fab* xed, iamrx_fab* yed, iamrx_fab* zed,
                           const int* iconserv, const iamrx_bcrec* bcrec, const iamrx_geom* geom, double dt, int flags,
                           void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(bx && aofs && Sf && umac && vmac && wmac && iconserv && geom, "null argument");
  IX_ARG(ncomp >= 1 && ncomp <= 8, "ncomp must be in [1,8]");
  IX_ARG(aofs_comp >= 0 && aofs_comp + ncomp <= aofs->ncomp && s_comp >= 0 && s_comp + ncomp <= Sf->ncomp, "component range");
  const bool wf = (flags & IAMRX_ADV_WRITE_FLUXES) != 0, known = (flags & IAMRX_ADV_KNOWN_EDGE_STATE) != 0;
  IX_ARG(!wf || (fx && fy && fz && xed && yed && zed), "flux/edge outputs required with WRITE_FLUXES");
  IX_ARG(!known || (xed && yed && zed), "edge states required with KNOWN_EDGE_STATE");
  IX_ARG((!uflux && !vflux && !wflux) || (uflux && vflux && wflux), "give all three flux velocities or none");
  IX_ARG(bc_codes_ok(bcrec, ncomp), "unknown BCRec code");
  const Bx b = mkbx(*bx);
  IX_ARG(covers(Sf, known ? b : grow(b, 3)), "S must cover the box grown by 3 cells (nghost_state, NSB.cpp:4539-4552)");
  IX_ARG(!force || !force->p || (f_comp >= 0 && f_comp + ncomp <= force->ncomp && covers(force, grow(b, 1))), "force must cover the box grown by 1");
  IX_ARG(!divu || !divu->p || covers(divu, grow(b, 1)), "divu must cover the box grown by 1");
  IX_ARG(covers(aofs, b), "aofs must cover the box");
  const iamrx_fab* mv[3] = {umac, vmac, wmac};
  const iamrx_fab* fv[3] = {uflux, vflux, wflux};
  iamrx_fab* fo[3] = {fx, fy, fz};
  iamrx_fab* eo[3] = {xed, yed, zed};
  for (int d = 0; d < 3; ++d) {
    // the transverse terms read the MAC velocities one cell outside the box in the other two directions
    IX_ARG(covers(mv[d], known ? surrounding(b, d) : surrounding(grow(b, 1), d)), "MAC velocities need one ghost face layer around the box");
    IX_ARG(!fv[d] || covers(fv[d], surrounding(b, d)), "flux velocities must cover the faces of the box");
    IX_ARG(!(wf || known) || !fo[d] || (fo[d]->ncomp >= ncomp && covers(fo[d], surrounding(b, d))), "flux output must cover the faces of the box");
    IX_ARG(!(wf || known) || !eo[d] || (eo[d]->ncomp >= ncomp && covers(eo[d], surrounding(b, d))), "edge states must cover the faces of the box");
  }
  k::AofsArgs a{};
  a.aofs = view(aofs, aofs_comp);
  a.S = cview(Sf, s_comp);
  a.force = cview(force, f_comp);
  a.divu = cview(divu);
  a.umac = cview(umac); a.vmac = cview(vmac); a.wmac = cview(wmac);
  // flux velocities: u_mac itself except in the sync call, where the fluxes are built with U_corr (NSB.cpp:4672-4677)
  a.uflx = uflux ? cview(uflux) : a.umac; a.vflx = vflux ? cview(vflux) : a.vmac; a.wflx = wflux ? cview(wflux) : a.wmac;
  if (wf || known) { a.fx = view(fx); a.fy = view(fy); a.fz = view(fz); a.xed = view(xed); a.yed = view(yed); a.zed = view(zed); }
  a.ncomp = ncomp;
  for (int n = 0; n < ncomp; ++n) a.iconserv[n] = iconserv[n];
  a.forces_in_trans = (flags & IAMRX_ADV_FORCES_IN_TRANS) ? 1 : 0;
  a.is_velocity = (flags & IAMRX_ADV_IS_VELOCITY) ? 1 : 0;
  a.is_sync = (flags & IAMRX_ADV_IS_SYNC) ? 1 : 0;
  a.write_fluxes = wf ? 1 : 0;
  a.known_edge_state = known ? 1 : 0;
  a.staged = (flags & IAMRX_ADV_STAGED) ? 1 : 0;
  a.ppm = (flags & IAMRX_ADV_PPM) ? 1 : 0;
  a.bc = make_advbc(bcrec, ncomp, geom);
  k::AdvGeom g;
  for (int d = 0; d < 3; ++d) g.dx[d] = geom->dx[d];
  g.dt = dt;
  return k::compute_aofs(b, a, g, S(stream));
}

int iamrx_nodal_divu_box(const iamrx_box* nbx, iamrx_fab* rhs, const iamrx_fab* vel, const double dxinv[3],
                         void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(nbx && rhs && vel && dxinv, "null argument");
  return k::nodal_divu(mkbx(*nbx), view(rhs), cview(vel), dxinv, S(stream));
}
int iamrx_nodal_adotx_box(const iamrx_box* nbx, iamrx_fab* out, const iamrx_fab* phi, const iamrx_fab* rhs,
                          const iamrx_fab* sigma, const double dxinv[3], void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(nbx && out && phi && sigma && dxinv, "null argument");
  return k::nodal_adotx(mkbx(*nbx), view(out), cview(phi), cview(rhs), cview(sigma), dxinv, S(stream));
}
int iamrx_nodal_gs_box(const iamrx_box* nbx, iamrx_fab* phi, const iamrx_fab* rhs, const iamrx_fab* sigma,
                       const double dxinv[3], int color, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(nbx && phi && rhs && sigma && dxinv, "null argument");
  IX_ARG(color >= 0 && color < 8, "colour must be in [0,8)");
  return k::nodal_gs_color(mkbx(*nbx), view(phi), cview(rhs), cview(sigma), dxinv, color, S(stream));
}
int iamrx_nodal_gs_sweep_box(const iamrx_box* nbx, iamrx_fab* phi_out, const iamrx_fab* phi_in, const iamrx_fab* rhs,
                             const iamrx_fab* sigma, const double dxinv[3], void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(nbx && phi_out && phi_in && rhs && sigma && dxinv, "null argument");
  IX_ARG(phi_out->p != phi_in->p, "the fused sweep is out of place");
  const Bx b = mkbx(*nbx);
  for (int d = 0; d < 3; ++d) IX_ARG(((b.hi[d] - b.lo[d]) % 2) == 0 && b.hi[d] - b.lo[d] >= 2, "even number of cells per direction required");
  return k::nodal_gs_sweep(b, view(phi_out), cview(phi_in), cview(rhs), cview(sigma), dxinv, S(stream));
}
int iamrx_nodal_mknewu_box(const iamrx_box* bx, iamrx_fab* vel, iamrx_fab* gp, const iamrx_fab* phi,
                           const iamrx_fab* sigma, const double dxinv[3], void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(bx && phi && dxinv, "null argument");
  IX_ARG(!vel || sigma, "sigma required to update vel");
  return k::nodal_mknewu(mkbx(*bx), view(vel), view(gp), 0, cview(phi), cview(sigma), dxinv, S(stream));
}

// ---------------------------------------------------------------------------
// 2. communicator + level
// ---------------------------------------------------------------------------
int iamrx_comm_unique_id(unsigned char uid[128]) { return comm_unique_id(uid); }
int iamrx_comm_init(int rank, int nranks, const unsigned char uid[128]) {
  IX_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "rank/nranks");
  IX_ARG(nranks == 1 || uid, "uid required for nranks > 1");
  return comm_init(rank, nranks, uid);
}
int iamrx_comm_set_transport(int rank, int nranks, iamrx_exchange_fn exchange, iamrx_allreduce_fn allreduce, void* ctx) {
  IX_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "rank/nranks");
  IX_ARG(nranks == 1 || (exchange && allreduce), "exchange and allreduce callbacks required for nranks > 1");
  return comm_set_transport(rank, nranks, exchange, allreduce, ctx);
}
int iamrx_comm_finalize(void) { return comm_finalize(); }
int iamrx_comm_rank(void) { return comm().rank; }
int iamrx_comm_size(void) { return comm().nranks; }
int iamrx_allreduce(double* dev_buf, int n, int op, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(dev_buf && n > 0 && op >= 0 && op <= 2, "allreduce arguments");
  return comm_allreduce(dev_buf, n, op, S(stream));
}

int iamrx_level_create(const iamrx_geom* geom, int nboxes, const iamrx_box* boxes, const int* owner,
                       iamrx_level_t* out) {
  IX_GUARD_BEGIN
  IX_ARG(geom && boxes && out && nboxes > 0, "null argument");
  std::vector<Bx> bxs; std::vector<int> own;
  const Bx dom = mkbx(geom->domain);
  for (int i = 0; i < nboxes; ++i) {
    Bx b = mkbx(boxes[i]);
    IX_ARG(b.ok(), "empty box");
    IX_ARG(intersect(b, dom).npts() == b.npts(), "box outside the domain");
    for (int j = 0; j < i; ++j) IX_ARG(!intersect(b, bxs[j]).ok(), "boxes overlap");
    bxs.push_back(b);
    const int o = owner ? owner[i] : 0;
    IX_ARG(o >= 0 && o < comm().nranks, "owner rank out of range");
    own.push_back(o);
  }
  for (int d = 0; d < 3; ++d) IX_ARG(geom->dx[d] > 0.0, "dx must be positive");
  auto* h = new iamrx_level_s();
  h->lev = make_level(*geom, bxs, own);
  *out = h;
  return IAMRX_OK;
  IX_GUARD_END
}
int iamrx_level_destroy(iamrx_level_t lev) {
  delete lev;
  return IAMRX_OK;
}
int iamrx_level_num_local(iamrx_level_t lev) { return lev ? lev->lev->nlocal() : IAMRX_ERR_ARG; }
int iamrx_level_local_box(iamrx_level_t lev, int il, iamrx_box* out, int* gi) {
  IX_ARG(lev && out && il >= 0 && il < lev->lev->nlocal(), "local box index");
  const Bx& b = lev->lev->lbox(il);
  for (int d = 0; d < 3; ++d) { out->lo[d] = b.lo[d]; out->hi[d] = b.hi[d]; }
  if (gi) *gi = lev->lev->local[il];
  return IAMRX_OK;
}

// test hook: the FillBoundary copy plan of a level as plain arrays (pure host
// logic; works without a device).  Returns the number of regions; fills up to
// `cap` entries of dst_box/src_box, 6 ints per region, 3 ints per shift.
int iamrx_debug_fb_plan(iamrx_level_t lev, int ixtype, int ng, int cap, int* dst_box, int* src_box,
                        int* region6, int* shift3) {
  IX_GUARD_BEGIN
  IX_ARG(lev, "null level");
  std::vector<int> db, sb, sh; std::vector<Bx> rg;
  build_fb_regions(*lev->lev, ixtype, ng, db, sb, rg, sh);
  const int n = (int)rg.size();
  for (int r = 0; r < n && r < cap; ++r) {
    if (dst_box) dst_box[r] = db[r];
    if (src_box) src_box[r] = sb[r];
    if (region6) for (int d = 0; d < 3; ++d) { region6[6 * r + d] = rg[r].lo[d]; region6[6 * r + 3 + d] = rg[r].hi[d]; }
    if (shift3) for (int d = 0; d < 3; ++d) shift3[3 * r + d] = sh[3 * r + d];
  }
  return n;
  IX_GUARD_END
}

// ---------------------------------------------------------------------------
// 3. level operators
// ---------------------------------------------------------------------------
int iamrx_fill_boundary(iamrx_level_t lev, iamrx_fab* fabs, int ixtype, int ncomp, int ngrow, void* stream) {
  IX_GUARD_BEGIN
  IX_NEED_DEVICE();
  IX_ARG(lev && fabs && ixtype >= 0 && ixtype <= 4 && ncomp >= 1 && ngrow >= 0, "fill_boundary arguments");
  MF m; m.alias(lev->lev.get(), ixtype, ncomp, ngrow, fabs);
  return mf_fill_boundary(m, 0, ncomp, ngrow, S(stream));
  IX_GUARD_END
}

int iamrx_fill_physbc(iamrx_level_t lev, iamrx_fab* fabs, int ncomp, int ngrow, const iamrx_bcrec* bcrec, const double* bcvals,
                      void* stream) {
  IX_GUARD_BEGIN
  IX_NEED_DEVICE();
  IX_ARG(lev && fabs && bcrec && ncomp >= 1 && ncomp <= 8 && ngrow >= 0, "fill_physbc arguments");
  IX_ARG(bc_codes_ok(bcrec, ncomp), "unknown BCRec code");
  k::PhysBC bc{};
  for (int n = 0; n < ncomp; ++n)
    for (int d = 0; d < 3; ++d) { bc.lo[n][d] = bcrec[n].lo[d]; bc.hi[n][d] = bcrec[n].hi[d]; }
  if (bcvals) for (int f = 0; f < 6; ++f) for (int n = 0; n < ncomp; ++n) bc.val[f][n] = bcvals[f * ncomp + n];
  MF m; m.alias(lev->lev.get(), IX_CELL, ncomp, ngrow, fabs);
  return mf_fill_physbc(m, 0, ncomp, ngrow, bc, S(stream));
  IX_GUARD_END
}

int iamrx_turbulent_force_box(const iamrx_box* bx, iamrx_fab* frc, const iamrx_fab* rho, const iamrx_geom* geom, double time,
                              int nmodes, int mode_start, int div_free_force, int array_size, const double* forcedata, void* stream) {
  IX_GUARD_BEGIN
  IX_NEED_DEVICE();
  IX_ARG(bx && frc && geom && forcedata, "null argument");
  IX_ARG(nmodes > 0 && mode_start >= 0 && mode_start <= nmodes && array_size > 0, "bad turbulent forcing parameters");
  const Bx b = mkbx(*bx);
  IX_ARG(frc->ncomp >= 3 && covers(frc, b), "force array does not cover the box");
  IX_ARG(!rho || covers(rho, b), "density array does not cover the box");
  double len[3];
  for (int d = 0; d < 3; ++d) len[d] = (geom->domain.hi[d] - geom->domain.lo[d] + 1) * geom->dx[d];
  const k::TurbParams tp{nmodes, mode_start, div_free_force ? 1 : 0, array_size};
  std::vector<k::TurbMode> modes;
  if (k::turb_modes(tp, forcedata, len, time, modes) != IAMRX_OK) { set_error("turbulent forcing: mode index beyond array_size"); return IAMRX_ERR_ARG; }
  const int nm = (int)modes.size();
  const size_t md = (modes.size() * sizeof(k::TurbMode) + 7) / 8 + 8;
  struct Guard { double* p; ~Guard() { dev_free(p); } } g{dev_alloc(md + k::turb_scratch_doubles(b, nm) + 8)};
  if (!g.p) return IAMRX_ERR_CUDA;
  if (nm > 0) IX_CUDA(cudaMemcpyAsync(g.p, modes.data(), modes.size() * sizeof(k::TurbMode), cudaMemcpyHostToDevice, S(stream)));
  int rc = k::turb_force(b, view(frc, 0), rho ? cview(rho, 0) : C4{}, *geom, reinterpret_cast<const k::TurbMode*>(g.p), nm, tp.div_free, g.p + md, 1, S(stream));
  IX_CUDA(cudaStreamSynchronize(S(stream)));   // the scratch and the host mode list are released on return
  return rc;
  IX_GUARD_END
}

void iamrx_mg_info_default(iamrx_mg_info* info) {
  if (!info) return;
  memset(info, 0, sizeof(*info));
  info->rtol = 1.0e-12;
  info->atol = 1.0e-16;
  info->max_iter = 200;
  info->max_coarsening = 100;
  info->nu1 = 2; info->nu2 = 2;
  info->bottom_sweeps = 8;
  info->verbose = 0;
  info->maxorder = 3;   // MLLinOp default; IAMR sets 4 for the MAC solve, 2 for diffusion
  info->bottom_solver = 0; info->bottom_maxiter = 200; info->bottom_rtol = 1.0e-4;
  info->omega = 1.15;   // AMReX abec_gsrb over-relaxation (AMReX_MLABecLap_3D_K.H); the UNVERIFIED-UPSTREAM table in DESIGN.md
}

// lobc / hibc of a scalar operator -> LinBC; periodic directions must be declared periodic and vice versa
static int make_linbc(const Level& L, const int lobc[3], const int hibc[3], int maxorder, bool nodal, k::LinBC* out) {
  k::LinBC b = periodic_linbc();
  b.maxorder = maxorder > 0 ? maxorder : 3;
  for (int d = 0; d < 3; ++d) {
    const int lo = lobc ? lobc[d] : IAMRX_LINOP_PERIODIC, hi = hibc ? hibc[d] : IAMRX_LINOP_PERIODIC;
    if (L.geom.periodic[d]) {
      IX_ARG(lo == IAMRX_LINOP_PERIODIC && hi == IAMRX_LINOP_PERIODIC, "periodic direction needs periodic BC");
    } else {
      for (int v : {lo, hi}) {
        IX_ARG(v == IAMRX_LINOP_DIRICHLET || v == IAMRX_LINOP_NEUMANN || (!nodal && v == IAMRX_LINOP_REFLECT_ODD) || (nodal && v == IAMRX_LINOP_INFLOW),
               "non-periodic direction needs Dirichlet / Neumann (cell: reflect_odd, nodal: inflow) boundary conditions");
      }
    }
    for (int c = 0; c < 3; ++c) { b.lo[c][d] = lo; b.hi[c][d] = hi; }
  }
  *out = b;
  return IAMRX_OK;
}
static int make_linbc3(const Level& L, const iamrx_linop_bc* bc, int ncomp, k::LinBC* out) {
  k::LinBC b = periodic_linbc();
  if (bc) {
    b.maxorder = bc->maxorder > 0 ? bc->maxorder : 2;
    for (int c = 0; c < 3; ++c) {
      const int cs = c < ncomp ? c : 0;
      for (int d = 0; d < 3; ++d) { b.lo[c][d] = bc->lo[cs][d]; b.hi[c][d] = bc->hi[cs][d]; }
    }
  }
  for (int c = 0; c < ncomp && c < 3; ++c)
    for (int d = 0; d < 3; ++d) {
      if (L.geom.periodic[d]) IX_ARG(b.lo[c][d] == IAMRX_LINOP_PERIODIC && b.hi[c][d] == IAMRX_LINOP_PERIODIC, "periodic direction needs periodic BC");
      else for (int v : {b.lo[c][d], b.hi[c][d]})
        IX_ARG(v == IAMRX_LINOP_DIRICHLET || v == IAMRX_LINOP_NEUMANN || v == IAMRX_LINOP_REFLECT_ODD,
               "non-periodic direction needs Dirichlet / Neumann / reflect_odd boundary conditions");
    }
  *out = b;
  return IAMRX_OK;
}

int iamrx_mac_project(iamrx_level_t lev, iamrx_fab* umac, iamrx_fab* vmac, iamrx_fab* wmac,
                      const iamrx_fab* rho, const iamrx_fab* rhs, iamrx_fab* phi, double rhs_scale,
                      const int lobc[3], const int hibc[3], iamrx_mg_info* info, void* stream) {
  IX_GUARD_BEGIN
  IX_NEED_DEVICE();
  IX_ARG(lev && umac && vmac && wmac && rho && phi, "null argument");
  IX_ARG(rhs_scale != 0.0, "rhs_scale");
  Level* L = lev->lev.get();
  k::LinBC bc;
  IX_TRY(make_linbc(*L, lobc, hibc, info ? info->maxorder : 0, false, &bc));
  iamrx_fab* um[3] = {umac, vmac, wmac};
  MF U[3];
  for (int d = 0; d < 3; ++d) U[d].alias(L, IX_XFACE + d, 1, 0, um[d]);
  MF Rho; Rho.alias(L, IX_CELL, 1, 1, rho);
  MF Phi; Phi.alias(L, IX_CELL, 1, 1, phi);
  MF Rhs; if (rhs) Rhs.alias(L, IX_CELL, 1, 0, rhs);
  return mac_project(*L, lev->solvers, U, Rho, rhs ? &Rhs : nullptr, Phi, rhs_scale, info, S(stream), &bc);
  IX_GUARD_END
}

// MacProj::mac_sync_solve (MacProj.cpp:359-479): see iamrx.h
int iamrx_mac_sync_solve(iamrx_level_t lev, iamrx_fluxreg_t mac_reg, const iamrx_fab* rho_half, const iamrx_fab* rhs_increment,
                         iamrx_fab* ucorr, iamrx_fab* vcorr, iamrx_fab* wcorr, iamrx_fab* mac_sync_phi, double dt, const int lobc[3],
                         const int hibc[3], iamrx_mg_info* info, void* stream) {
  IX_GUARD_BEGIN
  IX_NEED_DEVICE();
  IX_ARG(lev && mac_reg && rho_half && ucorr && vcorr && wcorr && mac_sync_phi && dt > 0.0, "mac_sync_solve arguments");
  Level* L = lev->lev.get();
  cudaStream_t s = S(stream);
  k::LinBC bc;
  IX_TRY(make_linbc(*L, lobc, hibc, info ? info->maxorder : 0, false, &bc));
  // Rhs = SUM{MR / VOL} over the faces of a coarse cell that adjoin the fine grids, sign opposite to a reflux (scale -1), zero
  // elsewhere and under the fine grids (:385-418); + Rhs_increment (:420-423); negated (:445)
  MF Rhs(L, IX_CELL, 1, 0);
  IX_TRY(mf_setval(Rhs, 0.0, 0, 1, 0, s));
  IX_TRY(iamrx_fluxreg_reflux(mac_reg, Rhs.fabs.data(), 0, -1.0, stream));
  if (rhs_increment) {
    MF Inc; Inc.alias(L, IX_CELL, 1, 0, const_cast<iamrx_fab*>(rhs_increment));
    IX_TRY(mf_lincomb(Rhs, 0, 1.0, Rhs, 0, 1.0, Inc, 0, 1, 0, s));
  }
  IX_TRY(mf_scale(Rhs, -1.0, 0, 1, 0, s));
  // mac_sync_phi = 0; solve with a null umac (no div(umac) in the right-hand side), rhs_scale = 2 / dt (:425-452)
  iamrx_fab* um[3] = {ucorr, vcorr, wcorr};
  MF U[3];
  for (int d = 0; d < 3; ++d) { U[d].alias(L, IX_XFACE + d, 1, 0, um[d]); IX_TRY(mf_setval(U[d], 0.0, 0, 1, 0, s)); }
  MF Rho; Rho.alias(L, IX_CELL, 1, 1, const_cast<iamrx_fab*>(rho_half));
  MF Phi; Phi.alias(L, IX_CELL, 1, 1, mac_sync_phi);
  IX_TRY(mf_setval(Phi, 0.0, 0, 1, 1, s));
  const int rc = mac_project(*L, lev->solvers, U, Rho, &Rhs, Phi, 2.0 / dt, info, s, &bc);
  if (rc < 0) return rc;
  // U now holds 0 - beta grad phi = the fluxes "-B grad phi"; Ucorr = -fluxes (:454-459)
  for (int d = 0; d < 3; ++d) IX_TRY(mf_scale(U[d], -1.0, 0, 1, 0, s));
  return rc;
  IX_GUARD_END
}

int iamrx_mac_get_fluxes(iamrx_level_t lev, iamrx_fab* fx, iamrx_fab* fy, iamrx_fab* fz, iamrx_fab* phi, void* stream) {
  IX_GUARD_BEGIN
  IX_NEED_DEVICE();
  IX_ARG(lev && fx && fy && fz && phi, "null argument");
  Level* L = lev->lev.get();
  iamrx_fab* ff[3] = {fx, fy, fz};
  MF F[3];
  for (int d = 0; d < 3; ++d) F[d].alias(L, IX_XFACE + d, 1, 0, ff[d]);
  MF Phi; Phi.alias(L, IX_CELL, 1, 1, phi);
  return mac_get_fluxes(*L, lev->solvers, F, Phi, S(stream));
  IX_GUARD_END
}

// Diffusion::computeExtensiveFluxes (Diffusion.cpp:1463-1537) for the scalar operator: MLMG::getFluxes (-b eta_d dphi/dx_d on the
// faces, with the ghost cells of soln as the solve's final fill left them) times fac * face area -- what diffuse_scalar hands to the
// viscous flux registers (Diffusion.cpp:560-566).
int iamrx_diffusion_get_fluxes(iamrx_level_t lev, int ncomp, iamrx_fab* fx, iamrx_fab* fy, iamrx_fab* fz, iamrx_fab* soln, double b,
                               const iamrx_fab* eta_x, const iamrx_fab* eta_y, const iamrx_fab* eta_z, double fac, void* stream) {
  IX_GUARD_BEGIN
  IX_NEED_DEVICE();
  IX_ARG(lev && fx && fy && fz && soln && eta_x && eta_y && eta_z && ncomp >= 1, "null argument");
  Level* L = lev->lev.get();
  cudaStream_t s = S(stream);
  iamrx_fab* ff[3] = {fx, fy, fz};
  const iamrx_fab* e[3] = {eta_x, eta_y, eta_z};
  MF F[3], E[3];
  for (int d = 0; d < 3; ++d) { F[d].alias(L, IX_XFACE + d, ncomp, 0, ff[d]); E[d].alias(L, IX_XFACE + d, 1, 0, e[d]); }
  MF Sol; Sol.alias(L, IX_CELL, ncomp, 1, soln);
  IX_TRY(mf_fill_boundary(Sol, 0, ncomp, 1, s));   // interior / periodic ghost cells; physical sides keep the caller's (solve's) values
  const double* dx = L->geom.dx;
  const double area[3] = {dx[1] * dx[2], dx[0] * dx[2], dx[0] * dx[1]};
  for (int il = 0; il < Sol.n(); ++il) {
    k::Abec op{};
    op.a = 0.0; op.b = b; op.bncomp = 1;
    op.bx = E[0].c(il); op.by = E[1].c(il); op.bz = E[2].c(il);
    for (int d = 0; d < 3; ++d) op.dxinv[d] = L->dxinv[d];
    for (int n = 0; n < ncomp; ++n) {
      IX_TRY(k::abec_flux(L->lbox(il), F[0].v(il, n), F[1].v(il, n), F[2].v(il, n), Sol.c(il, n), op, 0, s));
      for (int d = 0; d < 3; ++d) IX_TRY(k::scale(F[d].vbox(il), F[d].v(il, n), fac * area[d], 1, s));
    }
  }
  return IAMRX_OK;
  IX_GUARD_END
}

int iamrx_nodal_project(iamrx_level_t lev, iamrx_fab* vel, const iamrx_fab* sigma, iamrx_fab* phi,
                        iamrx_fab* gp, int increment_gp, const int lobc[3], const int hibc[3],
                        iamrx_mg_info* info, void* stream) {
  IX_GUARD_BEGIN
  IX_NEED_DEVICE();
  IX_ARG(lev && vel && sigma && phi, "null argument");
  Level* L = lev->lev.get();
  k::LinBC lb;
  IX_TRY(make_linbc(*L, lobc, hibc, 0, true, &lb));
  k::NodalBC nb;
  for (int d = 0; d < 3; ++d) { nb.lo[d] = lb.lo[0][d]; nb.hi[d] = lb.hi[0][d]; }
  MF Vel; Vel.alias(L, IX_CELL, 3, 1, vel);
  MF Sig; Sig.alias(L, IX_CELL, 1, 0, sigma);
  MF Phi; Phi.alias(L, IX_NODE, 1, 1, phi);
  MF Gp; if (gp) Gp.alias(L, IX_CELL, 3, 0, gp);
  return nodal_project(*L, lev->solvers, Vel, Sig, Phi, gp ? &Gp : nullptr, increment_gp, info, S(stream), &nb);
  IX_GUARD_END
}

int iamrx_diffusion_apply(iamrx_level_t lev, int tensor, int ncomp, iamrx_fab* out, iamrx_fab* soln, double a,
                          double b, const iamrx_fab* acoef, const iamrx_fab* eta_x, const iamrx_fab* eta_y,
                          const iamrx_fab* eta_z, const iamrx_linop_bc* bcin, void* stream) {
  IX_GUARD_BEGIN
  IX_NEED_DEVICE();
  IX_ARG(lev && out && soln && eta_x && eta_y && eta_z, "null argument");
  IX_ARG(!tensor || ncomp == 3, "tensor operator needs ncomp == 3");
  IX_ARG(a == 0.0 || acoef, "acoef required when a != 0");
  Level* L = lev->lev.get();
  MF Out; Out.alias(L, IX_CELL, ncomp, 0, out);
  MF Sol; Sol.alias(L, IX_CELL, ncomp, 1, soln);
  MF A; if (acoef) A.alias(L, IX_CELL, 1, 0, acoef);
  MF E[3];
  const iamrx_fab* e[3] = {eta_x, eta_y, eta_z};
  for (int d = 0; d < 3; ++d) E[d].alias(L, IX_XFACE + d, 1, 0, e[d]);
  k::LinBC bc;
  IX_TRY(make_linbc3(*L, bcin, ncomp, &bc));
  return diffusion_apply(*L, lev->solvers, tensor != 0, ncomp, Out, Sol, a, b, acoef ? &A : nullptr, E, S(stream), &bc);
  IX_GUARD_END
}

int iamrx_diffusion_solve(iamrx_level_t lev, int tensor, int ncomp, iamrx_fab* soln, const iamrx_fab* rhs,
                          double a, double b, const iamrx_fab* acoef, const iamrx_fab* eta_x,
                          const iamrx_fab* eta_y, const iamrx_fab* eta_z, const iamrx_linop_bc* bcin, iamrx_mg_info* info, void* stream) {
  IX_GUARD_BEGIN
  IX_NEED_DEVICE();
  IX_ARG(lev && soln && rhs && eta_x && eta_y && eta_z, "null argument");
  IX_ARG(!tensor || ncomp == 3, "tensor operator needs ncomp == 3");
  IX_ARG(a == 0.0 || acoef, "acoef required when a != 0");
  Level* L = lev->lev.get();
  MF Sol; Sol.alias(L, IX_CELL, ncomp, 1, soln);
  MF Rhs; Rhs.alias(L, IX_CELL, ncomp, 0, rhs);
  MF A; if (acoef) A.alias(L, IX_CELL, 1, 0, acoef);
  MF E[3];
  const iamrx_fab* e[3] = {eta_x, eta_y, eta_z};
  for (int d = 0; d < 3; ++d) E[d].alias(L, IX_XFACE + d, 1, 0, e[d]);
  k::LinBC bc;
  IX_TRY(make_linbc3(*L, bcin, ncomp, &bc));
  return diffusion_solve(*L, lev->solvers, tensor != 0, ncomp, Sol, Rhs, a, b, acoef ? &A : nullptr, E, info,
                         S(stream), &bc);
  IX_GUARD_END
}

}  // extern "C"

// accessors used by ns.cu
namespace ix {
Level* level_of(iamrx_level_t h) { return h ? h->lev.get() : nullptr; }
LevelSolvers* solvers_of(iamrx_level_t h) { return h ? &h->solvers : nullptr; }
}  // namespace ix
