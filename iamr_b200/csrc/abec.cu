// abec.cu -- cell-centred (a*alpha - b div beta grad) operator kernels for
// sm_100a: red-black GSRB colour pass, apply/residual, flux, restriction,
// prolongation, face-coefficient averaging, MAC divergence/update and the
// MLTensorOp cross terms.
//
// Stands in for AMReX MLABecLaplacian / MLCellLinOp / MLTensorOp device code
// reached from IAMR at MacProj.cpp:1150-1183 and Diffusion.cpp:327-567,715-768,
// 858-923,1708-1757 (arithmetic restated in oracle/).
//
// All kernels are HBM-bound 7-point stencils: x is unit stride, threadIdx.x
// runs along x, each CTA covers a (128 x 4) xy-strip of one z-plane so every
// warp request is a contiguous 256-byte (GSRB: strided 512-byte) span.
#include <cstdlib>
#include "kernels.h"

namespace ix {
namespace k {

namespace {

struct AbecDev {
  double a, b;
  C4 acoef, bx, by, bz;
  int bncomp;
  double dhx, dhy, dhz;  // b * dxinv^2
};

inline AbecDev to_dev(const Abec& op) {
  AbecDev d;
  d.a = op.a; d.b = op.b; d.acoef = op.acoef; d.bx = op.bx; d.by = op.by; d.bz = op.bz;
  d.bncomp = op.bncomp;
  d.dhx = op.b * op.dxinv[0] * op.dxinv[0];
  d.dhy = op.b * op.dxinv[1] * op.dxinv[1];
  d.dhz = op.b * op.dxinv[2] * op.dxinv[2];
  return d;
}

// ---- GSRB -----------------------------------------------------------------
// One thread per updated cell: thread t of a row handles the cell pair
// (2t, 2t+1) relative to the box and picks the one of the right colour.
constexpr int GS_TX = 64;
constexpr int GS_TY = 4;

template <int MINB>
__global__ void __launch_bounds__(GS_TX* GS_TY, MINB)
gsrb_kernel(Bx bx, V4 phi, C4 rhs, IX_KARG(AbecDev) op, double omega, int redblack, int nz, int wm) {
  const int kz = blockIdx.z % nz;
  const int n = blockIdx.z / nz;
  const int k = bx.lo[2] + kz;
  const int j = bx.lo[1] + blockIdx.y * GS_TY + threadIdx.y;
  if (j > bx.hi[1]) return;
  int i = bx.lo[0] + 2 * (blockIdx.x * GS_TX + threadIdx.x);
  // make (i + j + k + redblack) even
  i += (i + j + k + redblack) & 1;
  if (i > bx.hi[0]) return;
  const int nb = (op.bncomp > 1) ? n : 0;

  // 32-bit element offsets from the cell's own address (one address computation per array)
  const double* bxc = op.bx.p + nb * op.bx.ns + ((i - op.bx.l0) + (j - op.bx.l1) * op.bx.js + (k - op.bx.l2) * op.bx.ks);
  const double* byc = op.by.p + nb * op.by.ns + ((i - op.by.l0) + (j - op.by.l1) * op.by.js + (k - op.by.l2) * op.by.ks);
  const double* bzc = op.bz.p + nb * op.bz.ns + ((i - op.bz.l0) + (j - op.bz.l1) * op.bz.js + (k - op.bz.l2) * op.bz.ks);
  const double bxm = bxc[0], bxp = bxc[1];
  const double bym = byc[0], byp = byc[(int)op.by.js];
  const double bzm = bzc[0], bzp = bzc[(int)op.bz.ks];
  double* pc = phi.p + n * phi.ns + ((i - phi.l0) + (j - phi.l1) * phi.js + (k - phi.l2) * phi.ks);
  const int pjs = (int)phi.js, pks = (int)phi.ks;
  const double p0 = pc[0];
  double gamma = op.dhx * (bxm + bxp) + op.dhy * (bym + byp) + op.dhz * (bzm + bzp);
  if (op.a != 0.0) gamma += op.a * op.acoef(i, j, k);
  // periodic wrap inside the kernel when the box spans the domain (no ghost fill needed)
  const int oxm = ((wm & 1) && i == bx.lo[0]) ? bx.hi[0] - i : -1, oxp = ((wm & 1) && i == bx.hi[0]) ? bx.lo[0] - i : 1;
  const int oym = (((wm & 2) && j == bx.lo[1]) ? bx.hi[1] - j : -1) * pjs, oyp = (((wm & 2) && j == bx.hi[1]) ? bx.lo[1] - j : 1) * pjs;
  const int ozm = (((wm & 4) && k == bx.lo[2]) ? bx.hi[2] - k : -1) * pks, ozp = (((wm & 4) && k == bx.hi[2]) ? bx.lo[2] - k : 1) * pks;
  const double rho = op.dhx * (bxm * pc[oxm] + bxp * pc[oxp]) +
                     op.dhy * (bym * pc[oym] + byp * pc[oyp]) +
                     op.dhz * (bzm * pc[ozm] + bzp * pc[ozp]);
  const double res = rhs(i, j, k, n) - (gamma * p0 - rho);
  pc[0] = p0 + omega / gamma * res;
}

// ---- apply / residual ----------------------------------------------------
constexpr int AP_TX = 128;
constexpr int AP_TY = 2;

__global__ void __launch_bounds__(AP_TX* AP_TY)
apply_kernel(Bx bx, V4 out, C4 phi, C4 rhs, IX_KARG(AbecDev) op, int nz, int wm) {
  const int kz = blockIdx.z % nz;
  const int n = blockIdx.z / nz;
  const int k = bx.lo[2] + kz;
  const int j = bx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = bx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > bx.hi[1] || i > bx.hi[0]) return;
  const int nb = (op.bncomp > 1) ? n : 0;
  const double* bxc = op.bx.p + nb * op.bx.ns + ((i - op.bx.l0) + (j - op.bx.l1) * op.bx.js + (k - op.bx.l2) * op.bx.ks);
  const double* byc = op.by.p + nb * op.by.ns + ((i - op.by.l0) + (j - op.by.l1) * op.by.js + (k - op.by.l2) * op.by.ks);
  const double* bzc = op.bz.p + nb * op.bz.ns + ((i - op.bz.l0) + (j - op.bz.l1) * op.bz.js + (k - op.bz.l2) * op.bz.ks);
  const double* pc = phi.p + n * phi.ns + ((i - phi.l0) + (j - phi.l1) * phi.js + (k - phi.l2) * phi.ks);
  const int pjs = (int)phi.js, pks = (int)phi.ks;
  const double p0 = pc[0];
  const int oxm = ((wm & 1) && i == bx.lo[0]) ? bx.hi[0] - i : -1, oxp = ((wm & 1) && i == bx.hi[0]) ? bx.lo[0] - i : 1;
  const int oym = (((wm & 2) && j == bx.lo[1]) ? bx.hi[1] - j : -1) * pjs, oyp = (((wm & 2) && j == bx.hi[1]) ? bx.lo[1] - j : 1) * pjs;
  const int ozm = (((wm & 4) && k == bx.lo[2]) ? bx.hi[2] - k : -1) * pks, ozp = (((wm & 4) && k == bx.hi[2]) ? bx.lo[2] - k : 1) * pks;
  double y = -op.dhx * (bxc[1] * (pc[oxp] - p0) - bxc[0] * (p0 - pc[oxm])) -
             op.dhy * (byc[(int)op.by.js] * (pc[oyp] - p0) - byc[0] * (p0 - pc[oym])) -
             op.dhz * (bzc[(int)op.bz.ks] * (pc[ozp] - p0) - bzc[0] * (p0 - pc[ozm]));
  if (op.a != 0.0) y += op.a * op.acoef(i, j, k) * p0;
  out(i, j, k, n) = rhs.ok() ? (rhs(i, j, k, n) - y) : y;
}

__global__ void flux_kernel(Bx bx, V4 fx, V4 fy, V4 fz, C4 phi, IX_KARG(AbecDev) op, double fxs, double fys,
                            double fzs, int comp) {
  const int k = bx.lo[2] + blockIdx.z;
  const int j = bx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = bx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > bx.hi[1] + 1 || i > bx.hi[0] + 1 || k > bx.hi[2] + 1) return;
  const int nb = (op.bncomp > 1) ? comp : 0;
  const bool ii = i <= bx.hi[0], jj = j <= bx.hi[1], kk = k <= bx.hi[2];
  const double p0 = phi(i, j, k);
  if (fx.ok() && jj && kk) fx(i, j, k) = -fxs * op.bx(i, j, k, nb) * (p0 - phi(i - 1, j, k));
  if (fy.ok() && ii && kk) fy(i, j, k) = -fys * op.by(i, j, k, nb) * (p0 - phi(i, j - 1, k));
  if (fz.ok() && ii && jj) fz(i, j, k) = -fzs * op.bz(i, j, k, nb) * (p0 - phi(i, j, k - 1));
}

__global__ void restrict_kernel(Bx cbx, V4 crse, C4 fine, int nz) {
  const int kz = blockIdx.z % nz;
  const int n = blockIdx.z / nz;
  const int k = cbx.lo[2] + kz;
  const int j = cbx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = cbx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > cbx.hi[1] || i > cbx.hi[0]) return;
  const int ii = 2 * i, jj = 2 * j, kk = 2 * k;
  crse(i, j, k, n) = 0.125 * (fine(ii, jj, kk, n) + fine(ii + 1, jj, kk, n) + fine(ii, jj + 1, kk, n) +
                              fine(ii + 1, jj + 1, kk, n) + fine(ii, jj, kk + 1, n) +
                              fine(ii + 1, jj, kk + 1, n) + fine(ii, jj + 1, kk + 1, n) +
                              fine(ii + 1, jj + 1, kk + 1, n));
}

IX_D int cdiv2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }  // floor(a/2)

__global__ void prolong_kernel(Bx fbx, V4 fine, C4 crse, int nz) {
  const int kz = blockIdx.z % nz;
  const int n = blockIdx.z / nz;
  const int k = fbx.lo[2] + kz;
  const int j = fbx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = fbx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > fbx.hi[1] || i > fbx.hi[0]) return;
  fine(i, j, k, n) += crse(cdiv2(i), cdiv2(j), cdiv2(k), n);
}

__global__ void face_restrict_kernel(Bx cfbx, int dir, V4 crse, C4 fine, int nz) {
  const int kz = blockIdx.z % nz;
  const int n = blockIdx.z / nz;
  const int k = cfbx.lo[2] + kz;
  const int j = cfbx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = cfbx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > cfbx.hi[1] || i > cfbx.hi[0]) return;
  const int ii = 2 * i, jj = 2 * j, kk = 2 * k;
  double v;
  if (dir == 0)
    v = fine(ii, jj, kk, n) + fine(ii, jj + 1, kk, n) + fine(ii, jj, kk + 1, n) + fine(ii, jj + 1, kk + 1, n);
  else if (dir == 1)
    v = fine(ii, jj, kk, n) + fine(ii + 1, jj, kk, n) + fine(ii, jj, kk + 1, n) + fine(ii + 1, jj, kk + 1, n);
  else
    v = fine(ii, jj, kk, n) + fine(ii + 1, jj, kk, n) + fine(ii, jj + 1, kk, n) + fine(ii + 1, jj + 1, kk, n);
  crse(i, j, k, n) = 0.25 * v;
}

__global__ void rho_to_beta_kernel(Bx fbx, int dir, V4 beta, C4 rho, double scale) {
  const int k = fbx.lo[2] + blockIdx.z;
  const int j = fbx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = fbx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > fbx.hi[1] || i > fbx.hi[0]) return;
  const double rm = (dir == 0) ? rho(i - 1, j, k) : (dir == 1) ? rho(i, j - 1, k) : rho(i, j, k - 1);
  beta(i, j, k) = scale / (0.5 * (rm + rho(i, j, k)));
}

__global__ void mac_div_kernel(Bx bx, V4 div, C4 u, C4 v, C4 w, double fx, double fy, double fz,
                               C4 minus_rhs) {
  const int k = bx.lo[2] + blockIdx.z;
  const int j = bx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = bx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > bx.hi[1] || i > bx.hi[0]) return;
  double d = fx * (u(i + 1, j, k) - u(i, j, k)) + fy * (v(i, j + 1, k) - v(i, j, k)) +
             fz * (w(i, j, k + 1) - w(i, j, k));
  if (minus_rhs.ok()) d += minus_rhs(i, j, k);
  div(i, j, k) = d;
}

__global__ void mac_update_kernel(Bx bx, V4 u, V4 v, V4 w, C4 phi, IX_KARG(AbecDev) op, double sx, double sy,
                                  double sz) {
  const int k = bx.lo[2] + blockIdx.z;
  const int j = bx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = bx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > bx.hi[1] + 1 || i > bx.hi[0] + 1 || k > bx.hi[2] + 1) return;
  const bool ii = i <= bx.hi[0], jj = j <= bx.hi[1], kk = k <= bx.hi[2];
  const double p0 = phi(i, j, k);
  if (jj && kk) u(i, j, k) += -sx * op.bx(i, j, k) * (p0 - phi(i - 1, j, k));
  if (ii && kk) v(i, j, k) += -sy * op.by(i, j, k) * (p0 - phi(i, j - 1, k));
  if (ii && jj) w(i, j, k) += -sz * op.bz(i, j, k) * (p0 - phi(i, j, k - 1));
}

// ---- tensor cross terms --------------------------------------------------
// Transverse derivatives on faces (AMReX mltensor_d?_on_?face): average of the
// two centred differences either side of the face.
IX_D double dy_on_x(C4 v, int i, int j, int k, int n, double dyi) {
  return (v(i, j + 1, k, n) + v(i - 1, j + 1, k, n) - v(i, j - 1, k, n) - v(i - 1, j - 1, k, n)) * (0.25 * dyi);
}
IX_D double dz_on_x(C4 v, int i, int j, int k, int n, double dzi) {
  return (v(i, j, k + 1, n) + v(i - 1, j, k + 1, n) - v(i, j, k - 1, n) - v(i - 1, j, k - 1, n)) * (0.25 * dzi);
}
IX_D double dx_on_y(C4 v, int i, int j, int k, int n, double dxi) {
  return (v(i + 1, j, k, n) + v(i + 1, j - 1, k, n) - v(i - 1, j, k, n) - v(i - 1, j - 1, k, n)) * (0.25 * dxi);
}
IX_D double dz_on_y(C4 v, int i, int j, int k, int n, double dzi) {
  return (v(i, j, k + 1, n) + v(i, j - 1, k + 1, n) - v(i, j, k - 1, n) - v(i, j - 1, k - 1, n)) * (0.25 * dzi);
}
IX_D double dx_on_z(C4 v, int i, int j, int k, int n, double dxi) {
  return (v(i + 1, j, k, n) + v(i + 1, j, k - 1, n) - v(i - 1, j, k, n) - v(i - 1, j, k - 1, n)) * (0.25 * dxi);
}
IX_D double dy_on_z(C4 v, int i, int j, int k, int n, double dyi) {
  return (v(i, j + 1, k, n) + v(i, j + 1, k - 1, n) - v(i, j - 1, k, n) - v(i, j - 1, k - 1, n)) * (0.25 * dyi);
}

// cross flux through the x-face i (between cells i-1 and i), comps 0..2
IX_D void cross_fx(C4 vel, C4 ex, int i, int j, int k, double dyi, double dzi, double f[3]) {
  const double dudy = dy_on_x(vel, i, j, k, 0, dyi);
  const double dvdy = dy_on_x(vel, i, j, k, 1, dyi);
  const double dudz = dz_on_x(vel, i, j, k, 0, dzi);
  const double dwdz = dz_on_x(vel, i, j, k, 2, dzi);
  const double divu = dvdy + dwdz;
  const double mu = ex(i, j, k);
  f[0] = -mu * (-(2.0 / 3.0) * divu);
  f[1] = -mu * dudy;
  f[2] = -mu * dudz;
}
IX_D void cross_fy(C4 vel, C4 ey, int i, int j, int k, double dxi, double dzi, double f[3]) {
  const double dudx = dx_on_y(vel, i, j, k, 0, dxi);
  const double dvdx = dx_on_y(vel, i, j, k, 1, dxi);
  const double dvdz = dz_on_y(vel, i, j, k, 1, dzi);
  const double dwdz = dz_on_y(vel, i, j, k, 2, dzi);
  const double divu = dudx + dwdz;
  const double mu = ey(i, j, k);
  f[0] = -mu * dvdx;
  f[1] = -mu * (-(2.0 / 3.0) * divu);
  f[2] = -mu * dvdz;
}
IX_D void cross_fz(C4 vel, C4 ez, int i, int j, int k, double dxi, double dyi, double f[3]) {
  const double dudx = dx_on_z(vel, i, j, k, 0, dxi);
  const double dwdx = dx_on_z(vel, i, j, k, 2, dxi);
  const double dvdy = dy_on_z(vel, i, j, k, 1, dyi);
  const double dwdy = dy_on_z(vel, i, j, k, 2, dyi);
  const double divu = dudx + dvdy;
  const double mu = ez(i, j, k);
  f[0] = -mu * dwdx;
  f[1] = -mu * dwdy;
  f[2] = -mu * (-(2.0 / 3.0) * divu);
}

__global__ void __launch_bounds__(AP_TX* AP_TY)
tensor_cross_kernel(Bx bx, V4 out, C4 vel, C4 ex, C4 ey, C4 ez, double b, double dxi, double dyi,
                    double dzi) {
  const int k = bx.lo[2] + blockIdx.z;
  const int j = bx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = bx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > bx.hi[1] || i > bx.hi[0]) return;
  double fl[3], fh[3], acc[3];
  cross_fx(vel, ex, i, j, k, dyi, dzi, fl);
  cross_fx(vel, ex, i + 1, j, k, dyi, dzi, fh);
  for (int n = 0; n < 3; ++n) acc[n] = dxi * (fh[n] - fl[n]);
  cross_fy(vel, ey, i, j, k, dxi, dzi, fl);
  cross_fy(vel, ey, i, j + 1, k, dxi, dzi, fh);
  for (int n = 0; n < 3; ++n) acc[n] += dyi * (fh[n] - fl[n]);
  cross_fz(vel, ez, i, j, k, dxi, dyi, fl);
  cross_fz(vel, ez, i, j, k + 1, dxi, dyi, fh);
  for (int n = 0; n < 3; ++n) acc[n] += dzi * (fh[n] - fl[n]);
  for (int n = 0; n < 3; ++n) out(i, j, k, n) += b * acc[n];
}

inline dim3 grid_for(const Bx& bx, int tx, int ty, int nz_total) {
  return dim3(cdiv(bx.nx(), tx), cdiv(bx.ny(), ty), nz_total);
}

}  // namespace

int abec_gsrb(const Bx& bx, V4 phi, C4 rhs, const Abec& op, double omega, int redblack, int ncomp,
              cudaStream_t s, int wrapmask) {
  if (!bx.ok()) return IAMRX_OK;
  ProfScope prof_(IAMRX_PROF_ABEC_GSRB, bx.npts(), (double)bx.npts() * ncomp * (op.a != 0.0 ? 56.0 : 48.0), s);
  dim3 blk(GS_TX, GS_TY, 1);
  dim3 grd(cdiv(bx.nx() + 1, 2 * GS_TX), cdiv(bx.ny(), GS_TY), bx.nz() * ncomp);
  static int minb = -1;
  if (minb < 0) { const char* e = getenv("IAMRX_GSRB_MINB"); minb = e ? atoi(e) : 6; }
  if (minb >= 8) IX_LAUNCH(gsrb_kernel<8>, grd, blk, 0, s, bx, phi, rhs, to_dev(op), omega, redblack, bx.nz(), wrapmask);
  else IX_LAUNCH(gsrb_kernel<6>, grd, blk, 0, s, bx, phi, rhs, to_dev(op), omega, redblack, bx.nz(), wrapmask);
  return check_launch("abec_gsrb");
}

int abec_apply(const Bx& bx, V4 out, C4 phi, C4 rhs, const Abec& op, int ncomp, cudaStream_t s, int wrapmask) {
  if (!bx.ok()) return IAMRX_OK;
  ProfScope prof_(IAMRX_PROF_ABEC_APPLY, bx.npts(), (double)bx.npts() * ncomp * ((op.a != 0.0 ? 56.0 : 48.0) + (rhs.ok() ? 0.0 : -8.0)), s);
  IX_LAUNCH(apply_kernel, grid_for(bx, AP_TX, AP_TY, bx.nz() * ncomp), dim3(AP_TX, AP_TY, 1), 0, s, 
      bx, out, phi, rhs, to_dev(op), bx.nz(), wrapmask);
  return check_launch("abec_apply");
}

int abec_flux(const Bx& bx, V4 fx, V4 fy, V4 fz, C4 phi, const Abec& op, int comp, cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  Bx g = bx; g.hi[0]++; g.hi[1]++; g.hi[2]++;
  IX_LAUNCH(flux_kernel, grid_for(g, AP_TX, AP_TY, g.nz()), dim3(AP_TX, AP_TY, 1), 0, s, 
      bx, fx, fy, fz, phi, to_dev(op), op.b * op.dxinv[0], op.b * op.dxinv[1], op.b * op.dxinv[2], comp);
  return check_launch("abec_flux");
}

int cc_restrict(const Bx& cbx, V4 crse, C4 fine, int ncomp, cudaStream_t s) {
  if (!cbx.ok()) return IAMRX_OK;
  IX_LAUNCH(restrict_kernel, grid_for(cbx, AP_TX, AP_TY, cbx.nz() * ncomp), dim3(AP_TX, AP_TY, 1), 0, s, 
      cbx, crse, fine, cbx.nz());
  return check_launch("cc_restrict");
}

int cc_prolong_add(const Bx& fbx, V4 fine, C4 crse, int ncomp, cudaStream_t s) {
  if (!fbx.ok()) return IAMRX_OK;
  IX_LAUNCH(prolong_kernel, grid_for(fbx, AP_TX, AP_TY, fbx.nz() * ncomp), dim3(AP_TX, AP_TY, 1), 0, s, 
      fbx, fine, crse, fbx.nz());
  return check_launch("cc_prolong_add");
}

int face_restrict(const Bx& cfbx, int dir, V4 crse, C4 fine, int ncomp, cudaStream_t s) {
  if (!cfbx.ok()) return IAMRX_OK;
  IX_LAUNCH(face_restrict_kernel, grid_for(cfbx, AP_TX, AP_TY, cfbx.nz() * ncomp), dim3(AP_TX, AP_TY, 1), 0, s, 
      cfbx, dir, crse, fine, cfbx.nz());
  return check_launch("face_restrict");
}

int rho_to_beta(const Bx& fbx, int dir, V4 beta, C4 rho, double scale, cudaStream_t s) {
  if (!fbx.ok()) return IAMRX_OK;
  IX_LAUNCH(rho_to_beta_kernel, grid_for(fbx, AP_TX, AP_TY, fbx.nz()), dim3(AP_TX, AP_TY, 1), 0, s, 
      fbx, dir, beta, rho, scale);
  return check_launch("rho_to_beta");
}

int mac_divergence(const Bx& bx, V4 div, C4 u, C4 v, C4 w, const double dxinv[3], double fac,
                   C4 minus_rhs, cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  IX_LAUNCH(mac_div_kernel, grid_for(bx, AP_TX, AP_TY, bx.nz()), dim3(AP_TX, AP_TY, 1), 0, s, 
      bx, div, u, v, w, fac * dxinv[0], fac * dxinv[1], fac * dxinv[2], minus_rhs);
  return check_launch("mac_divergence");
}

int mac_update(const Bx& bx, V4 u, V4 v, V4 w, C4 phi, const Abec& op, cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  Bx g = bx; g.hi[0]++; g.hi[1]++; g.hi[2]++;
  IX_LAUNCH(mac_update_kernel, grid_for(g, AP_TX, AP_TY, g.nz()), dim3(AP_TX, AP_TY, 1), 0, s, 
      bx, u, v, w, phi, to_dev(op), op.b * op.dxinv[0], op.b * op.dxinv[1], op.b * op.dxinv[2]);
  return check_launch("mac_update");
}

int tensor_cross(const Bx& bx, V4 out, C4 vel, C4 ex, C4 ey, C4 ez, double b, const double dxinv[3],
                 cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  IX_LAUNCH(tensor_cross_kernel, grid_for(bx, AP_TX, AP_TY, bx.nz()), dim3(AP_TX, AP_TY, 1), 0, s, 
      bx, out, vel, ex, ey, ez, b, dxinv[0], dxinv[1], dxinv[2]);
  return check_launch("tensor_cross");
}

}  // namespace k
}  // namespace ix
