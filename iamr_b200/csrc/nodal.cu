// nodal.cu -- nodal (Q1 finite element) Laplacian kernels for sm_100a:
// 27-point div(sigma grad) apply/residual, multi-colour Gauss-Seidel and damped
// Jacobi smoothers, full-weighting restriction, trilinear interpolation, FE
// divergence RHS and the velocity / grad(p) update.
//
// Stands in for AMReX MLNodeLaplacian device code driven by Hydro::NodalProjector
// (IAMR call sites Projection.cpp:2512-2542, NSB.cpp:4106-4118); the stencil is
// the trilinear stiffness matrix with one sigma per cell, which
// the oracle re-derives by element integration (oracle/oracle.cpp nodal_*).
#include "kernels.h"

namespace ix {
namespace k {
namespace {

constexpr int TX = 64;
constexpr int TY = 4;

// A*phi at node (i,j,k) and the diagonal coefficient.
// im/ip, jm/jp, km/kp: indices of the neighbouring NODES (i-1/i+1 ..., or their periodic images)
IX_D double nodal_ax(C4 x, C4 sig, int i, int j, int k, int im, int ip, int jm, int jp, int km, int kp,
                     double facx, double facy, double facz, double& s0) {
  const double s000 = sig(i - 1, j - 1, k - 1), s100 = sig(i, j - 1, k - 1);
  const double s010 = sig(i - 1, j, k - 1), s110 = sig(i, j, k - 1);
  const double s001 = sig(i - 1, j - 1, k), s101 = sig(i, j - 1, k);
  const double s011 = sig(i - 1, j, k), s111 = sig(i, j, k);
  const double fxyz = facx + facy + facz;
  const double fmx2y2z = -facx + 2.0 * facy + 2.0 * facz;
  const double f2xmy2z = 2.0 * facx - facy + 2.0 * facz;
  const double f2x2ymz = 2.0 * facx + 2.0 * facy - facz;
  const double f4xm2ym2z = 4.0 * facx - 2.0 * facy - 2.0 * facz;
  const double fm2x4ym2z = -2.0 * facx + 4.0 * facy - 2.0 * facz;
  const double fm2xm2y4z = -2.0 * facx - 2.0 * facy + 4.0 * facz;
  s0 = (-4.0) * fxyz * (s000 + s100 + s010 + s110 + s001 + s101 + s011 + s111);
  double y = x(i, j, k) * s0;
  y += fxyz * (x(im, jm, km) * s000 + x(ip, jm, km) * s100 +
               x(im, jp, km) * s010 + x(ip, jp, km) * s110 +
               x(im, jm, kp) * s001 + x(ip, jm, kp) * s101 +
               x(im, jp, kp) * s011 + x(ip, jp, kp) * s111);
  y += fmx2y2z * (x(i, jm, km) * (s000 + s100) + x(i, jp, km) * (s010 + s110) +
                  x(i, jm, kp) * (s001 + s101) + x(i, jp, kp) * (s011 + s111));
  y += f2xmy2z * (x(im, j, km) * (s000 + s010) + x(ip, j, km) * (s100 + s110) +
                  x(im, j, kp) * (s001 + s011) + x(ip, j, kp) * (s101 + s111));
  y += f2x2ymz * (x(im, jm, k) * (s000 + s001) + x(ip, jm, k) * (s100 + s101) +
                  x(im, jp, k) * (s010 + s011) + x(ip, jp, k) * (s110 + s111));
  y += f4xm2ym2z * (x(im, j, k) * (s000 + s010 + s001 + s011) +
                    x(ip, j, k) * (s100 + s110 + s101 + s111));
  y += fm2x4ym2z * (x(i, jm, k) * (s000 + s100 + s001 + s101) +
                    x(i, jp, k) * (s010 + s110 + s011 + s111));
  y += fm2xm2y4z * (x(i, j, km) * (s000 + s100 + s010 + s110) +
                    x(i, j, kp) * (s001 + s101 + s011 + s111));
  return y;
}

// neighbour node indices; with wrap bit d set the node box [lo, hi] carries the periodic
// duplicate (node hi == node lo), so lo-1 -> hi-1 and hi+1 -> lo+1
#define NWRAP(bx, wm)                                                                          \
  const int im = (((wm) & 1) && i == bx.lo[0]) ? bx.hi[0] - 1 : i - 1, ip = (((wm) & 1) && i == bx.hi[0]) ? bx.lo[0] + 1 : i + 1; \
  const int jm = (((wm) & 2) && j == bx.lo[1]) ? bx.hi[1] - 1 : j - 1, jp = (((wm) & 2) && j == bx.hi[1]) ? bx.lo[1] + 1 : j + 1; \
  const int km = (((wm) & 4) && k == bx.lo[2]) ? bx.hi[2] - 1 : k - 1, kp = (((wm) & 4) && k == bx.hi[2]) ? bx.lo[2] + 1 : k + 1;

#define NIDX(bx)                                                   \
  const int k = bx.lo[2] + blockIdx.z;                             \
  const int j = bx.lo[1] + blockIdx.y * TY + threadIdx.y;          \
  const int i = bx.lo[0] + blockIdx.x * TX + threadIdx.x;          \
  if (j > bx.hi[1] || i > bx.hi[0]) return;

__global__ void __launch_bounds__(TX* TY)
adotx_kernel(Bx bx, V4 out, C4 phi, C4 rhs, C4 sig, double facx, double facy, double facz, int wm) {
  NIDX(bx)
  double s0;
  NWRAP(bx, wm)
  const double y = nodal_ax(phi, sig, i, j, k, im, ip, jm, jp, km, kp, facx, facy, facz, s0);
  out(i, j, k) = rhs.ok() ? (rhs(i, j, k) - y) : y;
}

__global__ void __launch_bounds__(TX* TY)
jacobi_kernel(Bx bx, V4 out, C4 phi, C4 rhs, C4 sig, double facx, double facy, double facz, double omega) {
  NIDX(bx)
  double s0;
  NWRAP(bx, 0)
  const double y = nodal_ax(phi, sig, i, j, k, im, ip, jm, jp, km, kp, facx, facy, facz, s0);
  out(i, j, k) = phi(i, j, k) + omega * (rhs(i, j, k) - y) / s0;
}

// colour = cx + 2*cy + 4*cz; nodes with (i&1,j&1,k&1) == (cx,cy,cz) are mutually
// uncoupled under the 27-point stencil.
__global__ void __launch_bounds__(TX* TY)
gs_color_kernel(Bx bx, V4 phi, C4 rhs, C4 sig, double facx, double facy, double facz, int i0, int j0,
                int k0, int wm) {
  const int k = k0 + 2 * blockIdx.z;
  const int j = j0 + 2 * (blockIdx.y * TY + threadIdx.y);
  const int i = i0 + 2 * (blockIdx.x * TX + threadIdx.x);
  if (k > bx.hi[2] || j > bx.hi[1] || i > bx.hi[0]) return;
  double s0;
  C4 x{phi.p, phi.l0, phi.l1, phi.l2, phi.js, phi.ks, phi.ns};
  NWRAP(bx, wm)
  const double y = nodal_ax(x, sig, i, j, k, im, ip, jm, jp, km, kp, facx, facy, facz, s0);
  phi(i, j, k) += (rhs(i, j, k) - y) / s0;
}

__global__ void __launch_bounds__(TX* TY) nd_restrict_kernel(Bx cbx, V4 crse, C4 fine) {
  NIDX(cbx)
  const int ii = 2 * i, jj = 2 * j, kk = 2 * k;
  double acc = 0.0;
#pragma unroll
  for (int dk = -1; dk <= 1; ++dk)
#pragma unroll
    for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
      for (int di = -1; di <= 1; ++di) {
        const double w = (double)((di == 0 ? 2 : 1) * (dj == 0 ? 2 : 1) * (dk == 0 ? 2 : 1));
        acc += w * fine(ii + di, jj + dj, kk + dk);
      }
  crse(i, j, k) = acc * (1.0 / 64.0);
}

IX_D int fl2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }

__global__ void __launch_bounds__(TX* TY) nd_interp_kernel(Bx fbx, V4 fine, C4 crse) {
  NIDX(fbx)
  const int ic = fl2(i), jc = fl2(j), kc = fl2(k);
  const int ox = i - 2 * ic, oy = j - 2 * jc, oz = k - 2 * kc;  // 0 or 1
  double acc = 0.0;
  for (int dk = 0; dk <= oz; ++dk)
    for (int dj = 0; dj <= oy; ++dj)
      for (int di = 0; di <= ox; ++di) acc += crse(ic + di, jc + dj, kc + dk);
  const double w = 1.0 / (double)((1 + ox) * (1 + oy) * (1 + oz));
  fine(i, j, k) += w * acc;
}

__global__ void __launch_bounds__(TX* TY)
divu_kernel(Bx bx, V4 rhs, C4 vel, double facx, double facy, double facz) {
  NIDX(bx)
  const double dx = facx * (-vel(i - 1, j - 1, k - 1, 0) + vel(i, j - 1, k - 1, 0) - vel(i - 1, j, k - 1, 0) +
                            vel(i, j, k - 1, 0) - vel(i - 1, j - 1, k, 0) + vel(i, j - 1, k, 0) -
                            vel(i - 1, j, k, 0) + vel(i, j, k, 0));
  const double dy = facy * (-vel(i - 1, j - 1, k - 1, 1) - vel(i, j - 1, k - 1, 1) + vel(i - 1, j, k - 1, 1) +
                            vel(i, j, k - 1, 1) - vel(i - 1, j - 1, k, 1) - vel(i, j - 1, k, 1) +
                            vel(i - 1, j, k, 1) + vel(i, j, k, 1));
  const double dz = facz * (-vel(i - 1, j - 1, k - 1, 2) - vel(i, j - 1, k - 1, 2) - vel(i - 1, j, k - 1, 2) -
                            vel(i, j, k - 1, 2) + vel(i - 1, j - 1, k, 2) + vel(i, j - 1, k, 2) +
                            vel(i - 1, j, k, 2) + vel(i, j, k, 2));
  rhs(i, j, k) = dx + dy + dz;
}

__global__ void __launch_bounds__(TX* TY)
mknewu_kernel(Bx bx, V4 vel, V4 gp, int incr, C4 p, C4 sig, double facx, double facy, double facz) {
  NIDX(bx)
  const double p000 = p(i, j, k), p100 = p(i + 1, j, k), p010 = p(i, j + 1, k), p110 = p(i + 1, j + 1, k);
  const double p001 = p(i, j, k + 1), p101 = p(i + 1, j, k + 1), p011 = p(i, j + 1, k + 1),
               p111 = p(i + 1, j + 1, k + 1);
  const double gx = facx * (-p000 + p100 - p010 + p110 - p001 + p101 - p011 + p111);
  const double gy = facy * (-p000 - p100 + p010 + p110 - p001 - p101 + p011 + p111);
  const double gz = facz * (-p000 - p100 - p010 - p110 + p001 + p101 + p011 + p111);
  if (vel.ok()) {
    const double s = sig(i, j, k);
    vel(i, j, k, 0) -= s * gx;
    vel(i, j, k, 1) -= s * gy;
    vel(i, j, k, 2) -= s * gz;
  }
  if (gp.ok()) {
    if (incr) { gp(i, j, k, 0) += gx; gp(i, j, k, 1) += gy; gp(i, j, k, 2) += gz; }
    else { gp(i, j, k, 0) = gx; gp(i, j, k, 1) = gy; gp(i, j, k, 2) = gz; }
  }
}

inline dim3 grid_for(const Bx& bx) { return dim3(cdiv(bx.nx(), TX), cdiv(bx.ny(), TY), bx.nz()); }
inline void facs(const double dxinv[3], double f[3]) {
  for (int d = 0; d < 3; ++d) f[d] = (1.0 / 36.0) * dxinv[d] * dxinv[d];
}

}  // namespace

int nodal_divu(const Bx& nbx, V4 rhs, C4 vel, const double dxinv[3], cudaStream_t s) {
  if (!nbx.ok()) return IAMRX_OK;
  IX_LAUNCH(divu_kernel, grid_for(nbx), dim3(TX, TY, 1), 0, s, nbx, rhs, vel, 0.25 * dxinv[0], 0.25 * dxinv[1],
                                                        0.25 * dxinv[2]);
  return check_launch("nodal_divu");
}

int nodal_adotx(const Bx& nbx, V4 out, C4 phi, C4 rhs, C4 sig, const double dxinv[3], cudaStream_t s, int wrapmask) {
  if (!nbx.ok()) return IAMRX_OK;
  double f[3]; facs(dxinv, f);
  ProfScope prof_(IAMRX_PROF_NODAL_ADOTX, nbx.npts(), (double)nbx.npts() * (rhs.ok() ? 32.0 : 24.0), s);
  IX_LAUNCH(adotx_kernel, grid_for(nbx), dim3(TX, TY, 1), 0, s, nbx, out, phi, rhs, sig, f[0], f[1], f[2], wrapmask);
  return check_launch("nodal_adotx");
}

int nodal_jacobi(const Bx& nbx, V4 out, C4 phi, C4 rhs, C4 sig, const double dxinv[3], double omega,
                 cudaStream_t s) {
  if (!nbx.ok()) return IAMRX_OK;
  double f[3]; facs(dxinv, f);
  IX_LAUNCH(jacobi_kernel, grid_for(nbx), dim3(TX, TY, 1), 0, s, nbx, out, phi, rhs, sig, f[0], f[1], f[2], omega);
  return check_launch("nodal_jacobi");
}

int nodal_gs_color(const Bx& nbx, V4 phi, C4 rhs, C4 sig, const double dxinv[3], int color,
                   cudaStream_t s, int wrapmask) {
  if (!nbx.ok()) return IAMRX_OK;
  double f[3]; facs(dxinv, f);
  const int c[3] = {color & 1, (color >> 1) & 1, (color >> 2) & 1};
  int o[3], n[3];
  for (int d = 0; d < 3; ++d) {
    o[d] = nbx.lo[d] + ((c[d] - nbx.lo[d]) & 1);
    n[d] = (nbx.hi[d] >= o[d]) ? (nbx.hi[d] - o[d]) / 2 + 1 : 0;
    if (n[d] == 0) return IAMRX_OK;
  }
  ProfScope prof_(IAMRX_PROF_NODAL_GS, nbx.npts(), (double)nbx.npts() * 4.0, s);  // 32 B/node/sweep over 8 colour passes
  dim3 grd(cdiv(n[0], TX), cdiv(n[1], TY), n[2]);
  IX_LAUNCH(gs_color_kernel, grd, dim3(TX, TY, 1), 0, s, nbx, phi, rhs, sig, f[0], f[1], f[2], o[0], o[1], o[2], wrapmask);
  return check_launch("nodal_gs_color");
}

int nodal_restrict(const Bx& cnbx, V4 crse, C4 fine, cudaStream_t s) {
  if (!cnbx.ok()) return IAMRX_OK;
  IX_LAUNCH(nd_restrict_kernel, grid_for(cnbx), dim3(TX, TY, 1), 0, s, cnbx, crse, fine);
  return check_launch("nodal_restrict");
}

int nodal_interp_add(const Bx& fnbx, V4 fine, C4 crse, cudaStream_t s) {
  if (!fnbx.ok()) return IAMRX_OK;
  IX_LAUNCH(nd_interp_kernel, grid_for(fnbx), dim3(TX, TY, 1), 0, s, fnbx, fine, crse);
  return check_launch("nodal_interp_add");
}

int nodal_mknewu(const Bx& bx, V4 vel, V4 gp, int increment_gp, C4 phi, C4 sig, const double dxinv[3],
                 cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  IX_LAUNCH(mknewu_kernel, grid_for(bx), dim3(TX, TY, 1), 0, s, bx, vel, gp, increment_gp, phi, sig,
                                                         0.25 * dxinv[0], 0.25 * dxinv[1], 0.25 * dxinv[2]);
  return check_launch("nodal_mknewu");
}

}  // namespace k
}  // namespace ix
