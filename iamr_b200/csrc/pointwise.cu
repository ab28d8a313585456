// pointwise.cu -- the small streaming kernels IAMR itself contributes to the hot
// path: forcing assembly, state updates, projection scaling and
// the problem initial conditions.  Each cites the amrex::ParallelFor lambda in
// the reference it stands in for; the arithmetic is restated, not copied.
#include "kernels.h"

namespace ix {
namespace k {
namespace {

constexpr int TX = 128;
constexpr int TY = 2;
inline dim3 grid_for(const Bx& bx, int nzc) { return dim3(cdiv(bx.nx(), TX), cdiv(bx.ny(), TY), nzc); }
#define IDX3(bx)                                                     \
  const int nz_ = bx.hi[2] - bx.lo[2] + 1;                            \
  const int k = bx.lo[2] + (int)(blockIdx.z % nz_);                   \
  const int n = (int)(blockIdx.z / nz_);                              \
  const int j = bx.lo[1] + blockIdx.y * TY + threadIdx.y;             \
  const int i = bx.lo[0] + blockIdx.x * TX + threadIdx.x;             \
  if (j > bx.hi[1] || i > bx.hi[0]) return;

// Projection::computeRhoG (Projection.cpp:1933-2379): the hydrostatic pressure on the node plane of an outflow face in x or y,
// phi(node, k) = sum over the cell layers k' >= k of -gravity * rhoExt(k') * dz, rhoExt = (3 rho_1 - rho_2) / 2 the density
// extrapolated to the face from the two cell layers next to it, taken at the node's transverse position: the mean of the two
// cell columns beside it, or at the domain's transverse edges what the density's BCRec says (ext_dir: the ghost column,
// hoextrap: linear extrapolation, foextrap: the first column).  One thread per node column, marching down from the top of the
// domain (the reference integrates serially too); only the nodes of the strip (this box's share of the plane) are written.
__global__ void outflow_rhog_kernel(Bx T, int ktop_node, V4 phi, C4 rho, IX_KARG(OutflowRhoG) a) {
  IDX3(T)
  (void)n;
  int q[3] = {i, j, k};
  const int jt = q[a.t];   // transverse node index
  auto R = [&](int c, int jj, int kk) { int p[3]; p[a.d] = c; p[a.t] = jj; p[2] = kk; return rho(p[0], p[1], p[2]); };
  auto col = [&](int c, int kk) -> double {
    if (jt == a.tlo && a.code_lo == IAMRX_BC_EXT_DIR) return R(c, jt - 1, kk);
    if (jt == a.tlo && a.code_lo == IAMRX_BC_HOEXTRAP) return 0.5 * (3.0 * R(c, jt, kk) - R(c, jt + 1, kk));
    if (jt == a.tlo && a.code_lo == IAMRX_BC_FOEXTRAP) return R(c, jt, kk);
    if (jt == a.thi && a.code_hi == IAMRX_BC_EXT_DIR) return R(c, jt, kk);
    if (jt == a.thi && a.code_hi == IAMRX_BC_HOEXTRAP) return 0.5 * (3.0 * R(c, jt - 1, kk) - R(c, jt - 2, kk));
    if (jt == a.thi && a.code_hi == IAMRX_BC_FOEXTRAP) return R(c, jt - 1, kk);
    return 0.5 * (R(c, jt, kk) + R(c, jt - 1, kk));
  };
  double rhog = 0.0;
  for (int kk = a.ztop; kk >= k; --kk) {   // cell layers from the top of the domain down to the bottom of the strip
    const double rho_ext = 0.5 * (3.0 * col(a.c1, kk) - col(a.c2, kk));
    rhog -= a.gravity * rho_ext * a.dz;
    if (kk <= ktop_node) phi(i, j, kk) = rhog;
  }
  if (ktop_node == a.ztop + 1) phi(i, j, ktop_node) = 0.0;   // the top node: the strip starts from zero (Projection.cpp:1888-1889)
}

__global__ void floor_kernel(Bx bx, V4 f) {
  IDX3(bx)
  const double v = f(i, j, k, n);
  f(i, j, k, n) = (fabs(v) > 1.0e-20) ? v : 0.0;
}

IX_D double ext_force(int n, double grav, double rho) {  // NS_getForce.cpp:117-141
  return (n == 2 && fabs(grav) > 1.0e-4) ? grav * rho : 0.0;
}

__global__ void force_vel_kernel(Bx bx, V4 tf, C4 visc, C4 gp, C4 rho, double grav, int div_rho, C4 uf) {
  IDX3(bx)
  const double r = rho(i, j, k);
  double f = ext_force(n, grav, r);
  if (uf.ok()) f += r * uf(i, j, k, n);   // density-weighted user forcing (Tutorials/HIT/NS_getForce.cpp:540, 619-621)
  if (visc.ok()) f += visc(i, j, k, n);
  if (gp.ok()) f -= gp(i, j, k, n);
  if (div_rho) f /= r;
  tf(i, j, k, n) = f;
}

__global__ void vel_update_kernel(Bx bx, V4 unew, C4 uold, C4 aofs, C4 gp, C4 rh, double grav, double dt,
                                  int zero_force, C4 rho_old, C4 rho_new, C4 uf) {
  IDX3(bx)
  const double r = rh(i, j, k);
  const double force = zero_force ? 0.0 : ext_force(n, grav, r) + (uf.ok() ? r * uf(i, j, k, n) : 0.0);
  if (rho_old.ok()) {  // do_mom_diff (NSB.cpp:3609-3616): momentum update, then back to velocity with the new density
    const double m = uold(i, j, k, n) * rho_old(i, j, k) - dt * aofs(i, j, k, n) + dt * force - dt * gp(i, j, k, n);
    unew(i, j, k, n) = m / rho_new(i, j, k);
    return;
  }
  unew(i, j, k, n) = uold(i, j, k, n) - dt * aofs(i, j, k, n) + dt * force / r - dt * gp(i, j, k, n) / r;
}

// NavierStokesBase::ConservativeScalMinMax / ConvectiveScalMinMax (NSB.cpp:4256-4370): clamp the new scalar (per unit mass when
// conservative) to the range of the old one over the 3x3x3 neighbourhood.  smx starts from numeric_limits::min() (the smallest
// POSITIVE double) exactly as in the reference.
__global__ void scal_minmax_kernel(Bx bx, V4 snew, C4 rhonew, C4 sold, C4 rhoold, int conservative) {
  IDX3(bx)
  (void)n;
  double smn = 1.7976931348623157e308, smx = 2.2250738585072014e-308;
  for (int kk = -1; kk <= 1; ++kk)
    for (int jj = -1; jj <= 1; ++jj)
      for (int ii = -1; ii <= 1; ++ii) {
        double v = sold(i + ii, j + jj, k + kk);
        if (conservative) v /= rhoold(i + ii, j + jj, k + kk);
        smn = fmin(smn, v); smx = fmax(smx, v);
      }
  if (conservative) snew(i, j, k) = fmin(fmax(snew(i, j, k) / rhonew(i, j, k), smn), smx) * rhonew(i, j, k);
  else snew(i, j, k) = fmin(fmax(snew(i, j, k), smn), smx);
}

__global__ void scal_update_kernel(Bx bx, V4 snew, C4 sold, C4 aofs, double dt) {
  IDX3(bx)
  snew(i, j, k, n) = sold(i, j, k, n) - dt * aofs(i, j, k, n);
}

__global__ void diff_rhs_kernel(Bx bx, V4 rhs, V4 unew, C4 rho) {
  IDX3(bx)
  const double u = unew(i, j, k, n) * rho(i, j, k);
  unew(i, j, k, n) = u;
  rhs(i, j, k, n) += u;
}

__global__ void proj_pre_kernel(Bx bx, V4 u, C4 gp, C4 rho, double dt_inv) {
  IDX3(bx)
  u(i, j, k, n) = u(i, j, k, n) * dt_inv + gp(i, j, k, n) / rho(i, j, k);
}

__global__ void invert_kernel(Bx bx, V4 sig, C4 rho) {
  IDX3(bx)
  sig(i, j, k, n) = 1.0 / rho(i, j, k, n);
}

struct ProbParams { double v[16]; int n; };

// prob_init.cpp: probtype 11 TaylorGreen (:509-560); probtype 5 DoubleShearLayer
// (:346-405, direction 1, extended uniformly in z); probtype 100 = synthetic
// variable-density Taylor-Green; probtype 20 = the HIT tutorial's initial condition.
__global__ void init_kernel(Bx bx, V4 st, int probtype, ProbParams pp, iamrx_geom g) {
  IDX3(bx)
  (void)n;
  const double twopi = 2.0 * 3.14159265358979323846264338327950288;
  const double x = g.prob_lo[0] + (i - g.domain.lo[0] + 0.5) * g.dx[0];
  const double y = g.prob_lo[1] + (j - g.domain.lo[1] + 0.5) * g.dx[1];
  const double z = g.prob_lo[2] + (k - g.domain.lo[2] + 0.5) * g.dx[2];
  if (probtype == 11 || probtype == 100) {
    const double a = pp.v[0], b = pp.v[1], c = pp.v[2], vx = pp.v[3], dens = pp.v[4];
    st(i, j, k, 0) = vx * sin(a * twopi * x) * cos(b * twopi * y) * cos(c * twopi * z);
    st(i, j, k, 1) = -vx * cos(a * twopi * x) * sin(b * twopi * y) * cos(c * twopi * z);
    st(i, j, k, 2) = 0.0;
    double rho = dens;
    if (probtype == 100) rho = dens * (1.0 + 0.5 * sin(twopi * x) * sin(twopi * y) * sin(twopi * z));
    st(i, j, k, 3) = rho;
    st(i, j, k, 4) = (dens * vx * vx / 16.0) * (2.0 + cos(2.0 * c * twopi * z)) *
                     (cos(2.0 * a * twopi * x) + cos(2.0 * b * twopi * y));
  } else if (probtype == 5) {
    // DoubleShearLayer, direction = 1 (shear layer in y): params = density, interface_width,
    // blob_x, blob_y, blob_z, blob_radius
    const double dens = pp.v[0], width = pp.v[1] > 0.0 ? pp.v[1] : 1.0;
    const double pi = 0.5 * twopi;
    st(i, j, k, 0) = -0.05 * sin(pi * y);
    st(i, j, k, 1) = tanh(30.0 * (0.5 - fabs(x)) / width);
    st(i, j, k, 2) = 0.0;
    st(i, j, k, 3) = dens;
    const double bx0 = pp.v[2], by0 = pp.v[3], bz0 = pp.v[4], br = pp.v[5];
    const double d = sqrt((x - bx0) * (x - bx0) + (y - by0) * (y - by0) + (z - bz0) * (z - bz0));
    st(i, j, k, 4) = (d < br) ? 1.0 : 0.0;
  } else if (probtype == 10) {
    // RayleighTaylor 3-D (prob_init.cpp:447-487): params = rho_1, rho_2, tra_1, tra_2, interface_width, perturbation_amplitude;
    // velocity at rest, the hard-coded random phases of :465-467
    const double pi = 0.5 * twopi;
    const double Lx = (g.domain.hi[0] - g.domain.lo[0] + 1) * g.dx[0], Ly = (g.domain.hi[1] - g.domain.lo[1] + 1) * g.dx[1];
    const double splitz = 0.5 * (g.prob_lo[2] + (g.prob_lo[2] + (g.domain.hi[2] - g.domain.lo[2] + 1) * g.dx[2]));
    const double ranampl = 2. * (0.6544437533747718 - 0.5), ranphse1 = 2. * pi * 0.1556190326530211, ranphse2 = 2. * pi * 0.4196144025537369;
    const double pert = ranampl * sin(2.0 * pi * x / Lx + ranphse1) * sin(2.0 * pi * y / Ly + ranphse2);
    const double pertheight = splitz - pp.v[5] * pert;
    st(i, j, k, 0) = 0.0; st(i, j, k, 1) = 0.0; st(i, j, k, 2) = 0.0;
    st(i, j, k, 3) = pp.v[0] + ((pp.v[1] - pp.v[0]) / 2.0) * (1.0 + tanh((z - pertheight) / pp.v[4]));
    st(i, j, k, 4) = pp.v[2] + ((pp.v[3] - pp.v[2]) / 2.0) * (1.0 + tanh((z - pertheight) / pp.v[4]));
  } else if (probtype == 1) {
    // LidDrivenCavity: start from rest, density 1 (prob_init.cpp:102-109)
    st(i, j, k, 0) = 0.0; st(i, j, k, 1) = 0.0; st(i, j, k, 2) = 0.0; st(i, j, k, 3) = 1.0; st(i, j, k, 4) = 0.0;
  } else if (probtype == 101) {
    // synthetic wall-bounded test field (not in the reference): params = amplitude, density, density variation
    const double pi = 0.5 * twopi, A = pp.v[0], dens = pp.v[1], vd = pp.v[2];
    const double X = (x - g.prob_lo[0]) / ((g.domain.hi[0] - g.domain.lo[0] + 1) * g.dx[0]);
    const double Y = (y - g.prob_lo[1]) / ((g.domain.hi[1] - g.domain.lo[1] + 1) * g.dx[1]);
    const double Z = (z - g.prob_lo[2]) / ((g.domain.hi[2] - g.domain.lo[2] + 1) * g.dx[2]);
    st(i, j, k, 0) = A * sin(pi * X) * cos(twopi * Y) * cos(pi * Z);
    st(i, j, k, 1) = -A * cos(pi * X) * sin(twopi * Y) * cos(twopi * Z) * 0.5;
    st(i, j, k, 2) = A * 0.3 * sin(twopi * X) * sin(twopi * Y) * sin(pi * Z);
    st(i, j, k, 3) = dens * (1.0 + vd * cos(twopi * X) * cos(twopi * Y) * cos(pi * Z));
    st(i, j, k, 4) = exp(-20.0 * ((X - 0.4) * (X - 0.4) + (Y - 0.5) * (Y - 0.5) + (Z - 0.6) * (Z - 0.6)));
  } else if (probtype == 20) {
    // HIT (Tutorials/HIT/prob_init.cpp:100-131): params = turb_scale, density [, amplitude of the synthetic density variation
    // of BASELINE.json's "variable-density HIT"; 0 = the reference's constant density].  Lz is measured from prob_lo[1] as
    // in the reference (:113).
    const double ts = pp.v[0], dens = pp.v[1], vd = pp.n > 2 ? pp.v[2] : 0.0;
    const double Lx = (g.domain.hi[0] - g.domain.lo[0] + 1) * g.dx[0], Ly = (g.domain.hi[1] - g.domain.lo[1] + 1) * g.dx[1];
    const double Lz = g.prob_lo[2] + (g.domain.hi[2] - g.domain.lo[2] + 1) * g.dx[2] - g.prob_lo[1];
    st(i, j, k, 0) = ts * cos(twopi * y / Ly) * cos(twopi * z / Lz);
    st(i, j, k, 1) = ts * cos(twopi * x / Lx) * cos(twopi * z / Lz);
    st(i, j, k, 2) = ts * cos(twopi * x / Lx) * cos(twopi * y / Ly);
    st(i, j, k, 3) = dens * (1.0 + vd * sin(twopi * x / Lx) * sin(twopi * y / Ly) * sin(twopi * z / Lz));
    st(i, j, k, 4) = 1.0;
  }
}

}  // namespace

#define LAUNCH3(kern, bx, ncomp, s, ...)                                                  \
  do {                                                                                    \
    if (!(bx).ok() || (ncomp) <= 0) return IAMRX_OK;                                      \
    IX_LAUNCH(kern, grid_for(bx, (bx).nz() * (ncomp)), dim3(TX, TY, 1), 0, s, bx, __VA_ARGS__);  \
    return check_launch(#kern);                                                           \
  } while (0)

int floor_small(const Bx& bx, V4 f, int ncomp, cudaStream_t s) { LAUNCH3(floor_kernel, bx, ncomp, s, f); }
int force_vel(const Bx& bx, V4 tf, C4 visc, C4 gp, C4 rho, double grav, int div_rho, cudaStream_t s, C4 uf) {
  LAUNCH3(force_vel_kernel, bx, 3, s, tf, visc, gp, rho, grav, div_rho, uf);
}
int vel_update(const Bx& bx, V4 unew, C4 uold, C4 aofs, C4 gp, C4 rhohalf, double grav, double dt, int zero_force,
               cudaStream_t s, C4 rho_old, C4 rho_new, C4 uf) {
  LAUNCH3(vel_update_kernel, bx, 3, s, unew, uold, aofs, gp, rhohalf, grav, dt, zero_force, rho_old, rho_new, uf);
}
int scal_update(const Bx& bx, V4 snew, C4 sold, C4 aofs, double dt, int ncomp, cudaStream_t s) {
  LAUNCH3(scal_update_kernel, bx, ncomp, s, snew, sold, aofs, dt);
}
int scal_minmax(const Bx& bx, V4 snew, C4 rhonew, C4 sold, C4 rhoold, int conservative, cudaStream_t s) {
  LAUNCH3(scal_minmax_kernel, bx, 1, s, snew, rhonew, sold, rhoold, conservative);
}
int diff_rhs(const Bx& bx, V4 rhs, V4 unew, C4 rho, int ncomp, cudaStream_t s) {
  LAUNCH3(diff_rhs_kernel, bx, ncomp, s, rhs, unew, rho);
}
int outflow_rhog(const Bx& strip, V4 phi, C4 rho, const OutflowRhoG& a, cudaStream_t s) {
  Bx T = strip;   // one thread per node column of the strip
  T.lo[2] = T.hi[2] = strip.lo[2];
  IX_LAUNCH(outflow_rhog_kernel, grid_for(T, 1), dim3(TX, TY, 1), 0, s, T, strip.hi[2], phi, rho, a);
  return check_launch("outflow_rhog");
}
int proj_pre(const Bx& bx, V4 u, C4 gp, C4 rho, double dt_inv, cudaStream_t s) {
  LAUNCH3(proj_pre_kernel, bx, 3, s, u, gp, rho, dt_inv);
}
int invert(const Bx& bx, V4 sig, C4 rho, cudaStream_t s) { LAUNCH3(invert_kernel, bx, 1, s, sig, rho); }
int init_prob(const Bx& bx, V4 state, int probtype, const double* params, int nparams, const iamrx_geom& g,
              cudaStream_t s) {
  ProbParams pp{};
  pp.n = nparams < 16 ? nparams : 16;
  for (int q = 0; q < pp.n; ++q) pp.v[q] = params[q];
  LAUNCH3(init_kernel, bx, 1, s, state, probtype, pp, g);
}

}  // namespace k
}  // namespace ix
