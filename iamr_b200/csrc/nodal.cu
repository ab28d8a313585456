// nodal.cu -- nodal (Q1 finite element) Laplacian kernels for sm_100a:
// 27-point div(sigma grad) apply/residual, multi-colour Gauss-Seidel and damped
// Jacobi smoothers, full-weighting restriction, trilinear interpolation, FE
// divergence RHS and the velocity / grad(p) update.
//
// Stands in for AMReX MLNodeLaplacian device code driven by Hydro::NodalProjector
// (IAMR call sites Projection.cpp:2512-2542, NSB.cpp:4106-4118); the stencil is
// the trilinear stiffness matrix with one sigma per cell, which
// the oracle re-derives by element integration (oracle/oracle.cpp nodal_*).
#include <cstdint>
#include <cstdlib>
#include "kernels.h"

namespace ix {
namespace k {
namespace {

constexpr int TX = 64;
constexpr int TY = 4;

// A*phi at node (i,j,k) and the diagonal coefficient.
// A*phi at a node and the diagonal coefficient s0, written against two accessors so that the
// one-thread-per-node kernels and the shared-memory tile kernels evaluate the SAME expression
// (bit-identical results):  X(di,dj,dk) = phi at node offset (di,dj,dk) in {-1,0,1}^3,
//                           S(di,dj,dk) = sigma of the cell at offset (di,dj,dk) in {-1,0}^3.
template <class XA, class SA>
IX_D double nodal_ax_rel(const XA& X, const SA& S, double facx, double facy, double facz, double& s0) {
  const double s000 = S(-1, -1, -1), s100 = S(0, -1, -1);
  const double s010 = S(-1, 0, -1), s110 = S(0, 0, -1);
  const double s001 = S(-1, -1, 0), s101 = S(0, -1, 0);
  const double s011 = S(-1, 0, 0), s111 = S(0, 0, 0);
  const double fxyz = facx + facy + facz;
  const double fmx2y2z = -facx + 2.0 * facy + 2.0 * facz;
  const double f2xmy2z = 2.0 * facx - facy + 2.0 * facz;
  const double f2x2ymz = 2.0 * facx + 2.0 * facy - facz;
  const double f4xm2ym2z = 4.0 * facx - 2.0 * facy - 2.0 * facz;
  const double fm2x4ym2z = -2.0 * facx + 4.0 * facy - 2.0 * facz;
  const double fm2xm2y4z = -2.0 * facx - 2.0 * facy + 4.0 * facz;
  s0 = (-4.0) * fxyz * (s000 + s100 + s010 + s110 + s001 + s101 + s011 + s111);
  double y = X(0, 0, 0) * s0;
  y += fxyz * (X(-1, -1, -1) * s000 + X(1, -1, -1) * s100 + X(-1, 1, -1) * s010 + X(1, 1, -1) * s110 +
               X(-1, -1, 1) * s001 + X(1, -1, 1) * s101 + X(-1, 1, 1) * s011 + X(1, 1, 1) * s111);
  y += fmx2y2z * (X(0, -1, -1) * (s000 + s100) + X(0, 1, -1) * (s010 + s110) +
                  X(0, -1, 1) * (s001 + s101) + X(0, 1, 1) * (s011 + s111));
  y += f2xmy2z * (X(-1, 0, -1) * (s000 + s010) + X(1, 0, -1) * (s100 + s110) +
                  X(-1, 0, 1) * (s001 + s011) + X(1, 0, 1) * (s101 + s111));
  y += f2x2ymz * (X(-1, -1, 0) * (s000 + s001) + X(1, -1, 0) * (s100 + s101) +
                  X(-1, 1, 0) * (s010 + s011) + X(1, 1, 0) * (s110 + s111));
  y += f4xm2ym2z * (X(-1, 0, 0) * (s000 + s010 + s001 + s011) + X(1, 0, 0) * (s100 + s110 + s101 + s111));
  y += fm2x4ym2z * (X(0, -1, 0) * (s000 + s100 + s001 + s101) + X(0, 1, 0) * (s010 + s110 + s011 + s111));
  y += fm2xm2y4z * (X(0, 0, -1) * (s000 + s100 + s010 + s110) + X(0, 0, 1) * (s001 + s101 + s011 + s111));
  return y;
}

// global-memory form: im/ip, jm/jp, km/kp are the neighbouring NODE indices (or their periodic
// images).  Neighbours are addressed as 32-bit element offsets from the node's own address.
IX_D double nodal_ax(C4 x, C4 sig, int i, int j, int k, int im, int ip, int jm, int jp, int km, int kp,
                     double facx, double facy, double facz, double& s0) {
  const double* xc = x.p + ((i - x.l0) + (j - x.l1) * x.js + (k - x.l2) * x.ks);
  const int xjs = (int)x.js, xks = (int)x.ks;
  const int oxm = im - i, oxp = ip - i, oym = (jm - j) * xjs, oyp = (jp - j) * xjs, ozm = (km - k) * xks, ozp = (kp - k) * xks;
  auto X = [&](int di, int dj, int dk) {
    return xc[(di < 0 ? oxm : (di > 0 ? oxp : 0)) + (dj < 0 ? oym : (dj > 0 ? oyp : 0)) + (dk < 0 ? ozm : (dk > 0 ? ozp : 0))];
  };
  const double* sc = sig.p + ((i - sig.l0) + (j - sig.l1) * sig.js + (k - sig.l2) * sig.ks);
  const int sjs = (int)sig.js, sks = (int)sig.ks;
  auto S = [&](int di, int dj, int dk) { return sc[di + dj * sjs + dk * sks]; };
  return nodal_ax_rel(X, S, facx, facy, facz, s0);
}

// neighbour node indices; with wrap bit d set the node box [lo, hi] carries the periodic
// duplicate (node hi == node lo), so lo-1 -> hi-1 and hi+1 -> lo+1
// bits 3..8 of wm (x lo, x hi, y lo, y hi, z lo, z hi): that side of the node box is a Neumann / inflow domain side, whose ghost
// node is the mirror image (lo-1 -> lo+1, hi+1 -> hi-1: mlndlap_applybc) -- evaluated in place, no ghost fill between the colours
#define NWRAP(bx, wm)                                                                          \
  const int im = (((wm) & 1) && i == bx.lo[0]) ? bx.hi[0] - 1 : ((((wm) & 8) && i == bx.lo[0]) ? i + 1 : i - 1),       \
            ip = (((wm) & 1) && i == bx.hi[0]) ? bx.lo[0] + 1 : ((((wm) & 16) && i == bx.hi[0]) ? i - 1 : i + 1);      \
  const int jm = (((wm) & 2) && j == bx.lo[1]) ? bx.hi[1] - 1 : ((((wm) & 32) && j == bx.lo[1]) ? j + 1 : j - 1),      \
            jp = (((wm) & 2) && j == bx.hi[1]) ? bx.lo[1] + 1 : ((((wm) & 64) && j == bx.hi[1]) ? j - 1 : j + 1);      \
  const int km = (((wm) & 4) && k == bx.lo[2]) ? bx.hi[2] - 1 : ((((wm) & 128) && k == bx.lo[2]) ? k + 1 : k - 1),     \
            kp = (((wm) & 4) && k == bx.hi[2]) ? bx.lo[2] + 1 : ((((wm) & 256) && k == bx.hi[2]) ? k - 1 : k + 1);

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
template <int MINB>
__global__ void __launch_bounds__(TX* TY, MINB)
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

// wm / fnb: directions in which the FINE node box fnb spans the periodic domain; the stencil wraps there instead of reading ghosts
__global__ void __launch_bounds__(TX* TY) nd_restrict_kernel(Bx cbx, V4 crse, C4 fine, int thin, int wm, Bx fnb) {
  NIDX(cbx)
  // full weighting (1, 2, 1) / 4 in every coarsened direction, injection in a direction with ratio 1 (thin bit set)
  const int r0 = (thin & 1) ? 1 : 2, r1 = (thin & 2) ? 1 : 2, r2 = (thin & 4) ? 1 : 2;
  const int ii = r0 * i, jj = r1 * j, kk = r2 * k;
  const int e0 = r0 - 1, e1 = r1 - 1, e2 = r2 - 1;   // stencil half-width: 1 or 0
  double acc = 0.0;
  for (int dk = -e2; dk <= e2; ++dk)
    for (int dj = -e1; dj <= e1; ++dj)
      for (int di = -e0; di <= e0; ++di) {
        const double w = (double)(((di == 0 && e0) ? 2 : 1) * ((dj == 0 && e1) ? 2 : 1) * ((dk == 0 && e2) ? 2 : 1));
        int fi = ii + di, fj = jj + dj, fk = kk + dk;
        if (wm & 1) fi = fi < fnb.lo[0] ? fi + (fnb.hi[0] - fnb.lo[0]) : (fi > fnb.hi[0] ? fi - (fnb.hi[0] - fnb.lo[0]) : fi);
        if (wm & 2) fj = fj < fnb.lo[1] ? fj + (fnb.hi[1] - fnb.lo[1]) : (fj > fnb.hi[1] ? fj - (fnb.hi[1] - fnb.lo[1]) : fj);
        if (wm & 4) fk = fk < fnb.lo[2] ? fk + (fnb.hi[2] - fnb.lo[2]) : (fk > fnb.hi[2] ? fk - (fnb.hi[2] - fnb.lo[2]) : fk);
        acc += w * fine(fi, fj, fk);
      }
  crse(i, j, k) = acc / (double)((e0 ? 4 : 1) * (e1 ? 4 : 1) * (e2 ? 4 : 1));
}

IX_D int fl2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }

__global__ void __launch_bounds__(TX* TY) nd_interp_kernel(Bx fbx, V4 fine, C4 crse, int thin) {
  NIDX(fbx)
  const int ic = (thin & 1) ? i : fl2(i), jc = (thin & 2) ? j : fl2(j), kc = (thin & 4) ? k : fl2(k);
  const int ox = (thin & 1) ? 0 : i - 2 * ic, oy = (thin & 2) ? 0 : j - 2 * jc, oz = (thin & 4) ? 0 : k - 2 * kc;  // 0 or 1
  double acc = 0.0;
  for (int dk = 0; dk <= oz; ++dk)
    for (int dj = 0; dj <= oy; ++dj)
      for (int di = 0; di <= ox; ++di) acc += crse(ic + di, jc + dj, kc + dk);
  const double w = 1.0 / (double)((1 + ox) * (1 + oy) * (1 + oz));
  fine(i, j, k) += w * acc;
}

// The same with one thread per COARSE cell (all three directions coarsened): its eight coarse corner nodes are loaded once and
// the 2 x 2 x 2 fine nodes (2 ic + {0,1}, ...) updated from registers; sums in the order of the loop above (bit-identical).
__global__ void __launch_bounds__(TX* TY) nd_interp8_kernel(Bx fbx, Bx cbx, V4 fine, C4 crse) {
  const int kc = cbx.lo[2] + blockIdx.z;
  const int jc = cbx.lo[1] + blockIdx.y * TY + threadIdx.y;
  const int ic = cbx.lo[0] + blockIdx.x * TX + threadIdx.x;
  if (jc > cbx.hi[1] || ic > cbx.hi[0]) return;
  const double* cp = crse.p + ((ic - crse.l0) + (jc - crse.l1) * crse.js + (kc - crse.l2) * crse.ks);
  const int cjs = (int)crse.js, cks = (int)crse.ks;
  // corner nodes beyond the fine box's coarsening are only read for fine nodes inside the box (guards below)
  const int i0 = 2 * ic, j0 = 2 * jc, k0 = 2 * kc;
  const bool xi[2] = {i0 >= fbx.lo[0] && i0 <= fbx.hi[0], i0 + 1 >= fbx.lo[0] && i0 + 1 <= fbx.hi[0]};
  const bool yi[2] = {j0 >= fbx.lo[1] && j0 <= fbx.hi[1], j0 + 1 >= fbx.lo[1] && j0 + 1 <= fbx.hi[1]};
  const bool zi[2] = {k0 >= fbx.lo[2] && k0 <= fbx.hi[2], k0 + 1 >= fbx.lo[2] && k0 + 1 <= fbx.hi[2]};
  double C[2][2][2];
#pragma unroll
  for (int dk = 0; dk < 2; ++dk)
#pragma unroll
    for (int dj = 0; dj < 2; ++dj)
#pragma unroll
      for (int di = 0; di < 2; ++di)
        C[dk][dj][di] = ((di == 0 || xi[1]) && (dj == 0 || yi[1]) && (dk == 0 || zi[1])) ? cp[di + dj * cjs + dk * cks] : 0.0;
  double* fp = fine.p + ((i0 - fine.l0) + (j0 - fine.l1) * fine.js + (k0 - fine.l2) * fine.ks);
  const int fjs = (int)fine.js, fks = (int)fine.ks;
  double F[2][2][2];
#pragma unroll
  for (int oz = 0; oz < 2; ++oz)
#pragma unroll
    for (int oy = 0; oy < 2; ++oy)
#pragma unroll
      for (int ox = 0; ox < 2; ++ox) F[oz][oy][ox] = (xi[ox] && yi[oy] && zi[oz]) ? fp[ox + oy * fjs + oz * fks] : 0.0;
#pragma unroll
  for (int oz = 0; oz < 2; ++oz)
#pragma unroll
    for (int oy = 0; oy < 2; ++oy)
#pragma unroll
      for (int ox = 0; ox < 2; ++ox) {
        if (!(xi[ox] && yi[oy] && zi[oz])) continue;
        double acc = 0.0;
#pragma unroll
        for (int dk = 0; dk <= oz; ++dk)
#pragma unroll
          for (int dj = 0; dj <= oy; ++dj)
#pragma unroll
            for (int di = 0; di <= ox; ++di) acc += C[dk][dj][di];
        const int sh = ox + oy + oz;
        const double w = sh == 0 ? 1.0 : (sh == 1 ? 0.5 : (sh == 2 ? 0.25 : 0.125));   // = 1 / ((1 + ox)(1 + oy)(1 + oz))
        fp[ox + oy * fjs + oz * fks] = F[oz][oy][ox] + w * acc;
      }
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

// the same next to Neumann / inflow domain sides (mlndlap_divu's zero_* factors): bit 2d / 2d+1 of `hide` = the low / high
// side of the node box in direction d is such a side; a node ON it does not see the TANGENTIAL velocities of the cells
// beyond it (the normal component of those ghost cells is used as it is: zero at walls, the inflow value otherwise)
__global__ void __launch_bounds__(TX* TY)
divu_bc_kernel(Bx bx, V4 rhs, C4 vel, double facx, double facy, double facz, int hide) {
  NIDX(bx)
  const int idx[3] = {i, j, k};
  const double fac[3] = {facx, facy, facz};
  bool hlo[3], hhi[3];
  for (int d = 0; d < 3; ++d) { hlo[d] = (hide >> (2 * d) & 1) && idx[d] == bx.lo[d]; hhi[d] = (hide >> (2 * d + 1) & 1) && idx[d] == bx.hi[d]; }
  double r = 0.0;
  for (int cz = 0; cz < 2; ++cz)
    for (int cy = 0; cy < 2; ++cy)
      for (int cx = 0; cx < 2; ++cx) {
        const int off[3] = {cx, cy, cz};
        bool outd[3];
        for (int d = 0; d < 3; ++d) outd[d] = (off[d] == 1 && hlo[d]) || (off[d] == 0 && hhi[d]);
        for (int c = 0; c < 3; ++c) {
          bool hidden = false;
          for (int d = 0; d < 3; ++d) if (d != c && outd[d]) hidden = true;
          if (!hidden) r += (off[c] ? -1.0 : 1.0) * fac[c] * vel(i - cx, j - cy, k - cz, c);
        }
      }
  rhs(i, j, k) = r;
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

#if !defined(IX_EMUL)
// ---- Gauss-Seidel colour pass with 128-bit loads + warp shuffles -------------------------------
// The updated nodes of one colour are 2 apart in x, so a warp's scalar loads of phi(i-1), phi(i),
// phi(i+1) each span 512 B and use half of every sector (ncu: 13 sectors/request, L1 tag stage the
// busiest unit).  Here each lane issues ONE aligned 16-byte load per (row, plane) that returns its
// own node and one x-neighbour; the other neighbour is the adjacent lane's second word (warp
// shuffle).  Same for sigma.  27 + 8 scalar loads become 9 + 4 vector loads + shuffles.
// The stencil is accumulated row by row:
//   y = sum_{dj,dk} [ F1(dj,dk) (Sm xm + Sp xp) + F0(dj,dk) (Sm + Sp) x0 ]
// with Sm/Sp = sums of sigma over the cells on the -x/+x side shared with row (dj,dk) and
//   F(di,dj,dk) = -36 sum_d fac_d s_d prod_{e != d} m_e,  s = +1 (same node) / -1, m = 2/6 (same) / 1/6
// (the Q1 element matrices; identical to the hand-expanded coefficients of nodal_ax_rel).
IX_HD double q1_factor(bool nx, bool ny, bool nz, double facx, double facy, double facz) {
  const double sx = nx ? -1.0 : 1.0, sy = ny ? -1.0 : 1.0, sz = nz ? -1.0 : 1.0;
  const double mx = nx ? 1.0 : 2.0, my = ny ? 1.0 : 2.0, mz = nz ? 1.0 : 2.0;  // x 1/6 each, folded into the -36
  return -(facx * sx * my * mz + facy * sy * mx * mz + facz * sz * mx * my);
}

__global__ void __launch_bounds__(TX* TY)
gs_color_vec_kernel(Bx bx, V4 phi, C4 rhs, C4 sig, double facx, double facy, double facz, int i0, int j0, int k0, int wm,
                    int phi_pair_at_i, int sig_pair_at_i) {
  const int k = k0 + 2 * blockIdx.z;
  const int j = j0 + 2 * (blockIdx.y * TY + threadIdx.y);
  const int i = i0 + 2 * (blockIdx.x * TX + threadIdx.x);
  if (k > bx.hi[2] || j > bx.hi[1]) return;  // uniform per warp (a warp is 32 consecutive x of one row)
  const bool valid = i <= bx.hi[0];
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const bool next_valid = (i + 2) <= bx.hi[0];
  const int ic = valid ? i : bx.hi[0];  // clamp so that address arithmetic stays in range for idle lanes
  double* pc = phi.p + ((ic - phi.l0) + (j - phi.l1) * phi.js + (k - phi.l2) * phi.ks);
  const int pjs = (int)phi.js, pks = (int)phi.ks;
  const bool wx = (wm & 1) != 0;
  const int oxm = (wx && ic == bx.lo[0]) ? (bx.hi[0] - 1 - ic) : -1, oxp = (wx && ic == bx.hi[0]) ? (bx.lo[0] + 1 - ic) : 1;
  const int oy[3] = {(((wm & 2) && j == bx.lo[1]) ? bx.hi[1] - 1 - j : -1) * pjs, 0, (((wm & 2) && j == bx.hi[1]) ? bx.lo[1] + 1 - j : 1) * pjs};
  const int oz[3] = {(((wm & 4) && k == bx.lo[2]) ? bx.hi[2] - 1 - k : -1) * pks, 0, (((wm & 4) && k == bx.hi[2]) ? bx.lo[2] + 1 - k : 1) * pks};
  // sigma: cells (i-1, i) x (j-1, j) x (k-1, k)
  const double* sc = sig.p + ((ic - sig.l0) + (j - sig.l1) * sig.js + (k - sig.l2) * sig.ks);
  const int sjs = (int)sig.js, sks = (int)sig.ks;
  double sm[2][2], sp[2][2];  // [dk+1][dj+1] for dj,dk in {-1,0}: sigma(i-1,..) and sigma(i,..)
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const double* sr = sc + (a - 1) * sjs + (b - 1) * sks;
      if (sig_pair_at_i) {  // aligned pair = (i, i+1): own cell in .x, cell i-1 is the left lane's .y
        double2 v = make_double2(0.0, 0.0);
        if (valid) {  // the last column's pair would end one element past the sigma row
          if (ic < bx.hi[0]) v = *reinterpret_cast<const double2*>(sr); else v.x = sr[0];
        }
        double left = __shfl_up_sync(full, v.y, 1);
        if (lane == 0 && valid) left = sr[-1];
        sp[b][a] = v.x; sm[b][a] = left;
      } else {              // aligned pair = (i-1, i)
        double2 v = make_double2(0.0, 0.0);
        if (valid) v = *reinterpret_cast<const double2*>(sr - 1);
        sm[b][a] = v.x; sp[b][a] = v.y;
      }
    }
  double y = 0.0, x00 = 0.0, s0 = 0.0;
#pragma unroll
  for (int dk = -1; dk <= 1; ++dk)
#pragma unroll
    for (int dj = -1; dj <= 1; ++dj) {
      // sigma sums over the cells shared with this row
      double Sm = 0.0, Sp = 0.0;
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const bool use_a = (dj == 0) || (dj < 0 ? a == 0 : a == 1);
          const bool use_b = (dk == 0) || (dk < 0 ? b == 0 : b == 1);
          if (use_a && use_b) { Sm += sm[b][a]; Sp += sp[b][a]; }
        }
      const double* pr = pc + oy[dj + 1] + oz[dk + 1];
      double xm, x0, xp;
      if (phi_pair_at_i) {  // aligned pair = (i, i+1)
        double2 v = make_double2(0.0, 0.0);
        if (valid) v = *reinterpret_cast<const double2*>(pr);
        x0 = v.x; xp = v.y;
        xm = __shfl_up_sync(full, v.y, 1);
        if (valid && (lane == 0)) xm = pr[oxm];
        if (valid && wx && ic == bx.hi[0]) xp = pr[oxp];
      } else {              // aligned pair = (i-1, i)
        double2 v = make_double2(0.0, 0.0);
        if (valid) v = *reinterpret_cast<const double2*>(pr - 1);
        xm = v.x; x0 = v.y;
        xp = __shfl_down_sync(full, v.x, 1);
        if (valid && (lane == 31 || !next_valid)) xp = pr[oxp];
        if (valid && wx && ic == bx.lo[0]) xm = pr[oxm];
      }
      const double F1 = q1_factor(true, dj != 0, dk != 0, facx, facy, facz);
      const double F0 = q1_factor(false, dj != 0, dk != 0, facx, facy, facz);
      if (dj == 0 && dk == 0) { s0 = F0 * (Sm + Sp); x00 = x0; y += F1 * (Sm * xm + Sp * xp) + s0 * x0; }
      else y += F1 * (Sm * xm + Sp * xp) + F0 * (Sm + Sp) * x0;
    }
  if (valid) pc[0] = x00 + (rhs(i, j, k) - y) / s0;
}
#endif

#if !defined(IX_EMUL)
// ---- shared-memory tile kernels (27-point stencil) ------------------------------------------
// One CTA = TXU x TYU updated nodes of one k-plane.  The phi tile (3 planes x (ST*TYU+1) rows x
// (ST*TXU+1) columns) and the sigma tile (2 x ST*TYU x ST*TXU cells) are staged with cp.async
// (8-byte LDGSTS: rows start at odd element offsets, so 16-byte copies are not possible), every
// global element is requested once per CTA with full-row coalescing, and the 27+8 stencil reads
// come from shared memory.  ST = 2: one Gauss-Seidel colour (updated nodes are 2 apart; the tile
// is stored de-interleaved, even and odd columns apart, so that a warp's stride-2 reads are
// contiguous in shared memory);  ST = 1: apply / residual over every node.
namespace tile {
constexpr int TXU = 64, TYU = 4;
template <int ST> struct Geo {
  static constexpr int NC = ST * (TXU - 1) + 3, NR = ST * (TYU - 1) + 3;  // nodes touched: first-1 .. last+1
  static constexpr int PITCH = (ST == 2) ? 132 : 66;
  static constexpr int SC = ST * (TXU - 1) + 2, SR = ST * (TYU - 1) + 2;  // cells touched by the updated nodes
  IX_D static int col(int c) { return ST == 2 ? ((c & 1) * 66 + (c >> 1)) : c; }
  IX_D static int scol(int c) { return ST == 2 ? ((c & 1) * (SC / 2) + (c >> 1)) : c; }
};
IX_D void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
IX_D int wrap_node(int g, int lo, int hi, bool w) {  // node lo-1 == hi-1, node hi+1 == lo+1 when periodic
  return w ? (g < lo ? hi - 1 : (g > hi ? lo + 1 : g)) : g;
}

template <int ST, bool GS>
__global__ void __launch_bounds__(TXU* TYU)
nodal_tile_kernel(Bx bx, V4 out, C4 phi, C4 rhs, C4 sig, double facx, double facy, double facz, int i0, int j0, int k0,
                  int wm) {
  using G = Geo<ST>;
  __shared__ double sp[3][G::NR][G::PITCH];
  __shared__ double ss[2][G::SR][G::SC];
  const int tid = threadIdx.y * TXU + threadIdx.x;
  const int ib = i0 + ST * TXU * (int)blockIdx.x, jb = j0 + ST * TYU * (int)blockIdx.y, k = k0 + ST * (int)blockIdx.z;
  for (int e = tid; e < 3 * G::NR * G::PITCH; e += TXU * TYU) {
    const int c = e % G::PITCH, rr = e / G::PITCH, r = rr % G::NR, p = rr / G::NR;
    const int gi = ib - 1 + c, gj = jb - 1 + r, gk = k - 1 + p;
    if (c < G::NC && gi <= bx.hi[0] + 1 && gj <= bx.hi[1] + 1)
      cp_async8(&sp[p][r][G::col(c)], &phi.p[(wrap_node(gi, bx.lo[0], bx.hi[0], wm & 1) - phi.l0) +
                                               (wrap_node(gj, bx.lo[1], bx.hi[1], wm & 2) - phi.l1) * phi.js +
                                               (wrap_node(gk, bx.lo[2], bx.hi[2], wm & 4) - phi.l2) * phi.ks]);
  }
  for (int e = tid; e < 2 * G::SR * G::SC; e += TXU * TYU) {
    const int c = e % G::SC, rr = e / G::SC, r = rr % G::SR, p = rr / G::SR;
    const int ci = ib - 1 + c, cj = jb - 1 + r, ck = k - 1 + p;
    if (ci <= bx.hi[0] && cj <= bx.hi[1])
      cp_async8(&ss[p][r][G::scol(c)], &sig.p[(ci - sig.l0) + (cj - sig.l1) * sig.js + (ck - sig.l2) * sig.ks]);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int i = ib + ST * (int)threadIdx.x, j = jb + ST * (int)threadIdx.y;
  if (i > bx.hi[0] || j > bx.hi[1]) return;
  const int c = 1 + ST * (int)threadIdx.x, r = 1 + ST * (int)threadIdx.y;
  auto X = [&](int di, int dj, int dk) { return sp[1 + dk][r + dj][G::col(c + di)]; };
  auto S = [&](int di, int dj, int dk) { return ss[1 + dk][r + dj][G::scol(c + di)]; };
  double s0;
  const double y = nodal_ax_rel(X, S, facx, facy, facz, s0);
  if (GS) out(i, j, k) = X(0, 0, 0) + (rhs(i, j, k) - y) / s0;
  else out(i, j, k) = rhs.ok() ? (rhs(i, j, k) - y) : y;
}
}  // namespace tile
#endif

#if !defined(IX_EMUL)
// ---- fused 8-colour Gauss-Seidel sweep (box spans the periodic domain) --------------------------
// The colour order 0..7 (colour = cx + 2 cy + 4 cz) visits every EVEN plane (cz = 0, colours 0-3)
// before any ODD plane (colours 4-7), and under the 27-point stencil an even plane only couples to
// itself and to the two odd planes next to it.  So one sweep is two phases of mutually independent
// planes: phase A updates all even planes from the old odd planes, phase B all odd planes from the
// new even planes.  Inside a plane the four (cx, cy) colours are run back to back on a shared-memory
// tile; the dependence of colour 3 on 2 on 1 on 0 reaches 3 nodes in x and 1 in y, so each CTA
// recomputes that halo redundantly (identical arithmetic, identical values) instead of waiting for
// its neighbours.  The sweep is out of place (phi_in -> phi_out) because neighbouring CTAs read each
// other's plane-k halo.  Per sweep every phi/sigma/rhs element is fetched from HBM about 1.5x
// (phi planes are read by both phases) instead of 8x by the colour-per-launch kernels.
//
// Tile: 56 x 14 updated nodes; loaded 64 columns x 17 rows x 3 planes of phi and 2 planes of sigma
// (8-byte cp.async; columns stored de-interleaved, even | odd, so that the stride-2 accesses of one
// colour are contiguous in shared memory).  One warp per tile row, one lane per node of the colour.
namespace fused {
constexpr int TXI = 56, TYI = 14, NC = 64, HALF = 32, NR = TYI + 4;

IX_D int wrap_node_any(int g, int lo, int hi) {  // any distance; node hi duplicates node lo
  if (g >= lo && g <= hi) return g;
  const int n = hi - lo;
  int m = (g - lo) % n;
  if (m < 0) m += n;
  return lo + m;
}
IX_D int wrap_cell_any(int c, int lo, int hi) {  // cells lo .. hi-1 (hi = node hi)
  if (c >= lo && c < hi) return c;
  const int n = hi - lo;
  int m = (c - lo) % n;
  if (m < 0) m += n;
  return lo + m;
}
IX_D int col(int c) { return (c & 1) * HALF + (c >> 1); }
// halo index of a tile: periodic image (the box spans the domain in this direction) or, for a box with neighbours
// (ghost = 1), the node / cell itself inside the DEEP ghost layers (depth GD, filled by the caller); indices further out
// only feed halo nodes outside the dependence cone of the box's own nodes, so they are clamped, not computed
constexpr int GD = 4;
IX_D int halo_node(int g, int lo, int hi, int ghost) { return ghost ? min(max(g, lo - GD), hi + GD) : wrap_node_any(g, lo, hi); }
IX_D int halo_cell(int c, int lo, int hi, int ghost) { return ghost ? min(max(c, lo - GD), hi - 1 + GD) : wrap_cell_any(c, lo, hi); }
// the same for boxes at least as wide as a tile (one period at most: no integer division; `big` is uniform per launch)
IX_D int halo_node_f(int g, int lo, int hi, int ghost, bool big) {
  if (!big) return halo_node(g, lo, hi, ghost);
  if (ghost) return min(max(g, lo - GD), hi + GD);
  const int n = hi - lo;
  return g < lo ? g + n : (g > hi ? g - n : g);
}
IX_D int halo_cell_f(int c, int lo, int hi, int ghost, bool big) {
  if (!big) return halo_cell(c, lo, hi, ghost);
  if (ghost) return min(max(c, lo - GD), hi - 1 + GD);
  const int n = hi - lo;
  return c < lo ? c + n : (c >= hi ? c - n : c);
}

struct Q1F { double f0c, f1c, f0j, f1j, f0k, f1k, f0jk, f1jk; };  // q1_factor by row kind

template <int CX, int CY>
IX_D void pass(double (*sp)[NR][NC], double (*ss)[NR][NC], int warp, int lane, double rhsv, const Q1F& q) {
  // nodes of this colour that the later colours (and finally the 56 x 14 interior) depend on
  constexpr int H0 = (CY == 0) ? 1 : 2;                         // first half-column index
  constexpr int NH = (CY == 0) ? (CX == 0 ? 31 : 30) : (CX == 0 ? 29 : 28);
  constexpr int NRW = (CY == 0) ? 8 : 7;
  if (warp < NRW && lane < NH) {
    const int h = H0 + lane;
    const int ty = 2 + CY + 2 * warp;
    // de-interleaved column slots of tx-1, tx, tx+1 for tx = 2h + CX
    const int c0 = CX ? HALF + h : h;
    const int cm = CX ? h : HALF + h - 1;
    const int cp = CX ? h + 1 : HALF + h;
    // A phi accumulated row by row (see gs_color_vec_kernel): for the row at offset (dj, dk)
    //   F1(dj,dk) (Sm xm + Sp xp) + F0(dj,dk) (Sm + Sp) x0,  Sm / Sp = sigma summed over the cells on the
    // -x / +x side that touch the row; rows of the same kind (centre, j-edge, k-edge, corner) share F.
    // Same stencil as nodal_ax_rel with ~55 instead of ~95 fp64 operations (rounding differs in the last bits).
    const double m00 = ss[0][ty - 1][cm], m01 = ss[0][ty][cm], m10 = ss[1][ty - 1][cm], m11 = ss[1][ty][cm];  // [dk+1][dj+1]
    const double p00 = ss[0][ty - 1][c0], p01 = ss[0][ty][c0], p10 = ss[1][ty - 1][c0], p11 = ss[1][ty][c0];
    const double mk0 = m00 + m01, mk1 = m10 + m11, mj0 = m00 + m10, mj1 = m01 + m11, mc = mk0 + mk1;
    const double pk0 = p00 + p01, pk1 = p10 + p11, pj0 = p00 + p10, pj1 = p01 + p11, pc = pk0 + pk1;
#define IXR(P, R) const double xm##P##R = sp[P][ty + R - 1][cm], x0##P##R = sp[P][ty + R - 1][c0], xp##P##R = sp[P][ty + R - 1][cp]
    IXR(0, 0); IXR(0, 1); IXR(0, 2); IXR(1, 0); IXR(1, 1); IXR(1, 2); IXR(2, 0); IXR(2, 1); IXR(2, 2);
#undef IXR
    // corners (dj, dk != 0): one cell on each side
    const double a1jk = m00 * xm00 + p00 * xp00 + m01 * xm02 + p01 * xp02 + m10 * xm20 + p10 * xp20 + m11 * xm22 + p11 * xp22;
    const double a0jk = (m00 + p00) * x000 + (m01 + p01) * x002 + (m10 + p10) * x020 + (m11 + p11) * x022;
    // k-edges (dj = 0, dk != 0) and j-edges (dj != 0, dk = 0): two cells on each side
    const double a1k = mk0 * xm01 + pk0 * xp01 + mk1 * xm21 + pk1 * xp21;
    const double a0k = (mk0 + pk0) * x001 + (mk1 + pk1) * x021;
    const double a1j = mj0 * xm10 + pj0 * xp10 + mj1 * xm12 + pj1 * xp12;
    const double a0j = (mj0 + pj0) * x010 + (mj1 + pj1) * x012;
    const double a1c = mc * xm11 + pc * xp11;
    const double s0 = q.f0c * (mc + pc);
    const double y = s0 * x011 + q.f1c * a1c + q.f1j * a1j + q.f0j * a0j + q.f1k * a1k + q.f0k * a0k + q.f1jk * a1jk + q.f0jk * a0jk;
    sp[1][ty][c0] = x011 + (rhsv - y) / s0;
  }
}

// ---- 2 x 2 node blocks (BLK = true) --------------------------------------------------------------------------------
// The colour passes above read 27 phi + 8 sigma values per node from shared memory and the kernel sits on the shared-memory
// pipe (ncu: l1tex 85 %).  Here a thread owns the 2 x 2 block of nodes (columns 2l, 2l+1; rows 2+2w, 3+2w) = one node of each
// in-plane colour.  The two neighbouring planes never change during the four passes, so their part of A phi is computed ONCE
// for the four nodes from the block's shared 4 x 4 neighbourhood (2 x 16 phi + 2 x 9 sigma loads instead of 4 x (18 + 8)); a
// pass then reads only the node's 3 x 3 in-plane neighbourhood: 86 instead of 140 shared loads per four nodes.
struct Blk {
  double z[3][3];   // sigma summed over the two cell planes: rows b-1 .. b+1, columns a-1 .. a+1 (cells)
  double r[4];      // rhs - (contribution of planes k-1, k+1) of node (cx, cy) at [cx + 2 cy]
};

// one neighbouring plane (phi rows b-1 .. b+2 = r[0..3] of sp, cell rows b-1 .. b+1 of ss) -> B.r, B.z
IX_D void blk_adjacent(const double (*sp)[NC], const double (*ss)[NC], int b, int r3, int sA, int sB, int sC, int sD, const Q1F& q, Blk& B,
                       bool first) {
  double S[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    S[j][0] = ss[b - 1 + j][sA]; S[j][1] = ss[b - 1 + j][sB]; S[j][2] = ss[b - 1 + j][sC];
#pragma unroll
    for (int i = 0; i < 3; ++i) B.z[j][i] = first ? S[j][i] : B.z[j][i] + S[j][i];
  }
  // sums shared by the nodes of the block: H[j][cx] = cells (-x, +x) of node column cx in cell row j; V[cy][i] = cells (-y, +y)
  // of node row cy in cell column i
  double H[3][2], V[2][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { H[j][0] = S[j][0] + S[j][1]; H[j][1] = S[j][1] + S[j][2]; }
#pragma unroll
  for (int i = 0; i < 3; ++i) { V[0][i] = S[0][i] + S[1][i]; V[1][i] = S[1][i] + S[2][i]; }
  const double n1jk = -q.f1jk, n0jk = -q.f0jk, n1k = -q.f1k, n0k = -q.f0k;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int row = (rr == 3) ? r3 : b - 1 + rr;
    const double X[4] = {sp[row][sA], sp[row][sB], sp[row][sC], sp[row][sD]};
#pragma unroll
    for (int cy = 0; cy < 2; ++cy) {
      const int dj = rr - cy - 1;   // this row's offset from the node row b + cy
      if (dj < -1 || dj > 1) continue;
#pragma unroll
      for (int cx = 0; cx < 2; ++cx) {
        const int n = cx + 2 * cy;
        if (dj == 0) {
          const double t1 = V[cy][cx] * X[cx] + V[cy][cx + 1] * X[cx + 2], t0 = (V[cy][cx] + V[cy][cx + 1]) * X[cx + 1];
          B.r[n] = fma(n0k, t0, fma(n1k, t1, B.r[n]));
        } else {
          const int j = (dj < 0) ? cy : cy + 1;   // cell row on that side of the node
          const double t1 = S[j][cx] * X[cx] + S[j][cx + 1] * X[cx + 2], t0 = H[j][cx] * X[cx + 1];
          B.r[n] = fma(n0jk, t0, fma(n1jk, t1, B.r[n]));
        }
      }
    }
  }
}

// colour pass (CX, CY) of the block: node column slot c0 (neighbours cm, cp), row ty.  The update divides by the diagonal
// through a correctly rounded reciprocal (a fraction of the instructions of an IEEE division; last-bit differences).
template <int CX, int CY>
IX_D double blk_pass(double (*sp)[NC], int ty, int cm, int c0, int cp, const Q1F& q, const Blk& B) {
  const double mj0 = B.z[CY][CX], pj0 = B.z[CY][CX + 1], mj1 = B.z[CY + 1][CX], pj1 = B.z[CY + 1][CX + 1];
  const double mc = mj0 + mj1, pc = pj0 + pj1;
  const double s0 = q.f0c * (mc + pc);
  const double x0 = sp[ty][c0];
  const double a1c = mc * sp[ty][cm] + pc * sp[ty][cp];
  const double a1j = mj0 * sp[ty - 1][cm] + pj0 * sp[ty - 1][cp] + mj1 * sp[ty + 1][cm] + pj1 * sp[ty + 1][cp];
  const double a0j = (mj0 + pj0) * sp[ty - 1][c0] + (mj1 + pj1) * sp[ty + 1][c0];
  const double y = s0 * x0 + q.f1c * a1c + q.f1j * a1j + q.f0j * a0j;
  const double v = x0 + (B.r[CX + 2 * CY] - y) * __drcp_rn(s0);
  sp[ty][c0] = v;
  return v;
}

// NRT = rows of the staged tile (NRT - 4 of them updated): 18 (8 warps, 46 KB, four CTAs per SM) or 34 (16 warps, 87 KB, two
// CTAs per SM; 30 of 34 instead of 14 of 18 staged rows are useful).  Dynamic shared memory: 5 planes of NRT x NC doubles.
// FAST: the box is at least a tile wide in x and y and its tile halos are periodic images (no deep ghost layers): halo indices by
// one compare-and-shift, no integer division and no per-variant branches in the staging code (half of the kernel's instructions).
template <bool BLK, int NRT, bool FAST>
__global__ void __launch_bounds__(16 * (NRT - 2), NRT == NR ? 4 : 2)
gs_sweep_kernel(Bx bx, V4 out, C4 pin, C4 padj, C4 rhs, C4 sig, IX_KARG(Q1F) q, int k0, int zwrap, int zmir, int xyg) {
  static_assert(BLK || NRT == NR, "the one-node-per-thread passes are written for the 18-row tile");
  constexpr int NTT = 16 * (NRT - 2), RPP = NTT / 64, TYT = NRT - 4, NWB = (NRT - 2) / 2;   // threads, rows per staging pass, updated rows, warps
  extern __shared__ double smraw[];
  double (*sp)[NRT][NC] = reinterpret_cast<double (*)[NRT][NC]>(smraw);
  double (*ss)[NRT][NC] = sp + 3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool bgx = FAST || bx.hi[0] - bx.lo[0] >= NC, bgy = FAST || bx.hi[1] - bx.lo[1] >= NC;   // a tile reaches less than one period beyond the box
  if (FAST) xyg = 0;
  const int X0 = (bx.lo[0] - (bx.lo[0] & 1)) + TXI * (int)blockIdx.x - 4;   // global node index of tile column 0
  const int Y0 = (bx.lo[1] - (bx.lo[1] & 1)) + TYT * (int)blockIdx.y - 2;   // ... of tile row 0
  const int k = k0 + 2 * (int)blockIdx.z;
  // z neighbours: periodic images when the box spans the domain in z, else the (filled) ghost planes / ghost cells
  // zmir bit 0 / 1: the low / high z side is a Neumann side (mirrored ghost plane, evaluated in place)
  const int km = (zwrap && k == bx.lo[2]) ? bx.hi[2] - 1 : (((zmir & 1) && k == bx.lo[2]) ? k + 1 : k - 1),
            kp = (zwrap && k == bx.hi[2]) ? bx.lo[2] + 1 : (((zmir & 2) && k == bx.hi[2]) ? k - 1 : k + 1);
  const int ckm = zwrap ? wrap_cell_any(k - 1, bx.lo[2], bx.hi[2]) : k - 1, ckp = zwrap ? wrap_cell_any(k, bx.lo[2], bx.hi[2]) : k;
  // stage phi (rows 1..17 of three planes) and sigma (cell rows 1..16 of planes k-1, k): a thread
  // always loads the same tile column, so the x wrap is done once
  {
    // lanes 0-15 of a warp take the even columns of its 32-column segment, lanes 16-31 the odd ones: each
    // half-warp then writes 128 contiguous bytes of shared memory (no bank conflict in the de-interleaved
    // layout) while the warp still reads one contiguous 256-byte global segment
    const int c = (tid & 32) + 2 * (tid & 15) + ((tid >> 4) & 1), r0 = 1 + (tid >> 6), sc = col(c);
    const int gi = halo_node_f(X0 + c, bx.lo[0], bx.hi[0], xyg & 1, bgx);
    const int ci = halo_cell_f(X0 + c, bx.lo[0], bx.hi[0], xyg & 1, bgx);
    const double* pk = pin.p + (gi - pin.l0) + (int64_t)(k - pin.l2) * pin.ks;
    const double* pm = padj.p + (gi - padj.l0) + (int64_t)(km - padj.l2) * padj.ks;
    const double* pp = padj.p + (gi - padj.l0) + (int64_t)(kp - padj.l2) * padj.ks;
    const double* s0p = sig.p + (ci - sig.l0) + (int64_t)(ckm - sig.l2) * sig.ks;
    const double* s1p = sig.p + (ci - sig.l0) + (int64_t)(ckp - sig.l2) * sig.ks;
    const int pjs = (int)pin.js, ajs = (int)padj.js, sjs = (int)sig.js;
#pragma unroll
    for (int m = 0; m < (NRT - 1 + RPP - 1) / RPP; ++m) {
      const int r = r0 + RPP * m;
      if (r <= NRT - 1) {
        const int gj = halo_node_f(Y0 + r, bx.lo[1], bx.hi[1], xyg & 2, bgy);
        tile::cp_async8(&sp[0][r][sc], pm + (gj - padj.l1) * ajs);
        tile::cp_async8(&sp[1][r][sc], pk + (gj - pin.l1) * pjs);
        tile::cp_async8(&sp[2][r][sc], pp + (gj - padj.l1) * ajs);
        if (r <= NRT - 2) {
          const int cj = halo_cell_f(Y0 + r, bx.lo[1], bx.hi[1], xyg & 2, bgy) - sig.l1;
          tile::cp_async8(&ss[0][r][sc], s0p + cj * sjs);
          tile::cp_async8(&ss[1][r][sc], s1p + cj * sjs);
        }
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if constexpr (BLK) {
    // block (warp, lane): node columns a = 2 (lane + 1), a + 1; rows b = 2 + 2 warp, b + 1
    const int lp = lane + 1, a = 2 * lp, b = 2 + 2 * warp;
    const bool act = lane <= 30;
    Blk B;
    {
      const int gi0 = halo_node_f(X0 + a, bx.lo[0], bx.hi[0], xyg & 1, bgx), gi1 = halo_node_f(X0 + min(a + 1, NC - 1), bx.lo[0], bx.hi[0], xyg & 1, bgx);
      const int gj0 = halo_node_f(Y0 + b, bx.lo[1], bx.hi[1], xyg & 2, bgy), gj1 = halo_node_f(Y0 + b + 1, bx.lo[1], bx.hi[1], xyg & 2, bgy);
      const double* rp = rhs.p + (int64_t)(k - rhs.l2) * rhs.ks - rhs.l0;
      const int o0 = (gj0 - rhs.l1) * (int)rhs.js, o1 = (gj1 - rhs.l1) * (int)rhs.js;
      B.r[0] = act ? rp[o0 + gi0] : 0.0; B.r[1] = act ? rp[o0 + gi1] : 0.0;
      B.r[2] = act ? rp[o1 + gi0] : 0.0; B.r[3] = act ? rp[o1 + gi1] : 0.0;
    }
    // de-interleaved slots of columns a-1, a, a+1, a+2 (column 64 and row 18 only feed nodes that are never updated: clamped)
    const int sA = HALF + lp - 1, sB = lp, sC = HALF + lp, sD = min(lp + 1, HALF - 1);
    const int r3 = min(b + 2, NRT - 1);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    double v00 = 0.0, v10 = 0.0, v01 = 0.0, v11 = 0.0;   // the block's new values (every node is updated once: they are final)
    if (act) {
      blk_adjacent(sp[0], ss[0], b, r3, sA, sB, sC, sD, q, B, true);
      blk_adjacent(sp[2], ss[1], b, r3, sA, sB, sC, sD, q, B, false);
      v00 = blk_pass<0, 0>(sp[1], b, sA, sB, sC, q, B);                                  // columns 2 .. 62, rows 2 .. 16
    }
    __syncthreads();
    if (lane <= 29) v10 = blk_pass<1, 0>(sp[1], b, sB, sC, sD, q, B);                    // columns 3 .. 61
    __syncthreads();
    if (lane >= 1 && lane <= 29 && warp <= NWB - 2) v01 = blk_pass<0, 1>(sp[1], b + 1, sA, sB, sC, q, B);   // columns 4 .. 60, rows 3 .. 15
    __syncthreads();
    if (lane >= 1 && lane <= 28 && warp <= NWB - 2) v11 = blk_pass<1, 1>(sp[1], b + 1, sB, sC, sD, q, B);   // columns 5 .. 59
    // the tile's interior (columns 4 .. 59, rows 2 .. NRT - 3) straight from the registers: the two stores of a row pair fill
    // each other's sector halves
    {
      const int gi = X0 + a, gj = Y0 + b;
      const bool c0 = lane >= 1 && lane <= 28 && gi >= bx.lo[0] && gi <= bx.hi[0], c1 = lane >= 1 && lane <= 28 && gi + 1 >= bx.lo[0] && gi + 1 <= bx.hi[0];
      const bool r0 = warp <= NWB - 2 && gj >= bx.lo[1] && gj <= bx.hi[1], r1 = warp <= NWB - 2 && gj + 1 >= bx.lo[1] && gj + 1 <= bx.hi[1];
      double* op = out.p + (gi - out.l0) + (int64_t)(gj - out.l1) * out.js + (int64_t)(k - out.l2) * out.ks;
      if (r0 && c0) op[0] = v00;
      if (r0 && c1) op[1] = v10;
      if (r1 && c0) op[out.js] = v01;
      if (r1 && c1) op[out.js + 1] = v11;
    }
    return;
  } else {
  // right-hand sides of the (up to) four nodes this thread updates, one per colour
  double rv[4];
#pragma unroll
  for (int qq = 0; qq < 4; ++qq) {
    const int cx = qq & 1, cy = qq >> 1;
    const int h = (cy == 0 ? 1 : 2) + lane, ty = 2 + cy + 2 * warp;
    const int gi = halo_node(X0 + 2 * h + cx, bx.lo[0], bx.hi[0], xyg & 1), gj = halo_node(Y0 + ty, bx.lo[1], bx.hi[1], xyg & 2);
    rv[qq] = (warp < 8 && ty < NR) ? rhs(gi, gj, k) : 0.0;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  double (*sp18)[NR][NC] = reinterpret_cast<double (*)[NR][NC]>(smraw);
  double (*ss18)[NR][NC] = sp18 + 3;
  pass<0, 0>(sp18, ss18, warp, lane, rv[0], q);
  __syncthreads();
  pass<1, 0>(sp18, ss18, warp, lane, rv[1], q);
  __syncthreads();
  pass<0, 1>(sp18, ss18, warp, lane, rv[2], q);
  __syncthreads();
  pass<1, 1>(sp18, ss18, warp, lane, rv[3], q);
  __syncthreads();
  for (int e = tid; e < TXI * TYT; e += NTT) {   // the interior from shared memory
    const int tx = 4 + e % TXI, ty = 2 + e / TXI;
    const int gi = X0 + tx, gj = Y0 + ty;
    if (gi >= bx.lo[0] && gi <= bx.hi[0] && gj >= bx.lo[1] && gj <= bx.hi[1]) out(gi, gj, k) = sp[1][ty][col(tx)];
  }
  }
}
}  // namespace fused
#endif

#if !defined(IX_EMUL)
// ---- z-marching apply / residual (27-point stencil, register window) -----------------------------
// One thread owns a node column (i, j) and marches through KB planes keeping the three phi planes (3 x 3 values each)
// and the two sigma planes (2 x 2 cells each) of the current node in registers: per node it loads ONE new phi plane
// (9 values) and one new sigma plane (4 values) instead of 27 + 8, which takes the kernel off the L1 pipe.  The
// arithmetic is the row form of fused::pass (same rounding as the fused smoother).
namespace march {
struct Pl { double v[3][3]; };   // [row j-1, j, j+1][column i-1, i, i+1]
struct Sg { double m0, m1, p0, p1; };   // cells (i-1, j-1), (i-1, j), (i, j-1), (i, j) of one cell plane

template <int KB, int MINB>
__global__ void __launch_bounds__(TX* TY, MINB)
adotx_march_kernel(Bx bx, V4 out, C4 phi, C4 rhs, C4 sig, IX_KARG(fused::Q1F) q, int wm, int nchunk, double* norm) {
  const int i = bx.lo[0] + blockIdx.x * TX + threadIdx.x;
  const int j = bx.lo[1] + blockIdx.y * TY + threadIdx.y;
  if (i > bx.hi[0] || j > bx.hi[1]) return;
  const int kc0 = bx.lo[2] + KB * (int)blockIdx.z;
  const int kc1 = min(kc0 + KB - 1, bx.hi[2]);
  (void)nchunk;
  // neighbour columns / rows (periodic images when the box spans the domain: node hi duplicates node lo)
  const int im = ((wm & 1) && i == bx.lo[0]) ? bx.hi[0] - 1 : (((wm & 8) && i == bx.lo[0]) ? i + 1 : i - 1),
            ip = ((wm & 1) && i == bx.hi[0]) ? bx.lo[0] + 1 : (((wm & 16) && i == bx.hi[0]) ? i - 1 : i + 1);
  const int jm = ((wm & 2) && j == bx.lo[1]) ? bx.hi[1] - 1 : (((wm & 32) && j == bx.lo[1]) ? j + 1 : j - 1),
            jp = ((wm & 2) && j == bx.hi[1]) ? bx.lo[1] + 1 : (((wm & 64) && j == bx.hi[1]) ? j - 1 : j + 1);
  const int pjs = (int)phi.js, sjs = (int)sig.js;
  const double* pcol = phi.p + (i - phi.l0);
  const int ox[3] = {im - i, 0, ip - i};
  const int oy[3] = {(jm - phi.l1) * pjs, (j - phi.l1) * pjs, (jp - phi.l1) * pjs};
  const double* scol = sig.p + (i - sig.l0) + (j - sig.l1) * sjs;
  auto zplane = [&](int k) {  // node plane index with periodic image / Neumann mirror image
    if (wm & 4) return k < bx.lo[2] ? bx.hi[2] - 1 : (k > bx.hi[2] ? bx.lo[2] + 1 : k);
    if ((wm & 128) && k < bx.lo[2]) return bx.lo[2] + 1;
    if ((wm & 256) && k > bx.hi[2]) return bx.hi[2] - 1;
    return k;
  };
  auto load_pl = [&](Pl& P, int k) {
    const double* b = pcol + (int64_t)(zplane(k) - phi.l2) * phi.ks;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) P.v[r][c] = b[oy[r] + ox[c]];
  };
  auto load_sg = [&](Sg& S, int kcell) {  // cell plane kcell (ghost cells of sigma are filled)
    const double* b = scol + (int64_t)(kcell - sig.l2) * sig.ks;
    S.m0 = b[-1 - sjs]; S.m1 = b[-1]; S.p0 = b[-sjs]; S.p1 = b[0];
  };
  // coupling of a node with the plane above / below it through ONE cell layer S: U(X, S) is what plane X contributes to
  // the node on the other side of layer S.  The lower plane's term is computed one iteration early (when that plane and
  // its layer are both in registers) and carried as a scalar, so only two raw planes stay live -- which leaves room to
  // load the NEXT plane and layer one iteration ahead of their use (the loop is latency-, not bandwidth-bound).
  auto U = [&](const Pl& X, const Sg& S) {
    const double mk = S.m0 + S.m1, pk = S.p0 + S.p1;
    return q.f1jk * (S.m0 * X.v[0][0] + S.p0 * X.v[0][2] + S.m1 * X.v[2][0] + S.p1 * X.v[2][2]) +
           q.f0jk * ((S.m0 + S.p0) * X.v[0][1] + (S.m1 + S.p1) * X.v[2][1]) + q.f1k * (mk * X.v[1][0] + pk * X.v[1][2]) +
           q.f0k * (mk + pk) * X.v[1][1];
  };
  Pl B, C, D;
  Sg S0, S1, S2;
  {
    Pl A;
    load_pl(A, kc0 - 1); load_sg(S0, kc0 - 1);
    load_pl(B, kc0); load_sg(S1, kc0);
    load_pl(C, kc0 + 1);
    D = C; S2 = S1;
    double ua = U(A, S0);
    double nrm = 0.0;     // max |out| of this column (norm != nullptr: the residual norm, fused)
    bool isnan_ = false;
    for (int k = kc0; k <= kc1; ++k) {
      if (k < kc1) { load_pl(D, k + 2); load_sg(S2, k + 1); }   // next iteration's plane and layer
      const double mj0 = S0.m0 + S1.m0, mj1 = S0.m1 + S1.m1, pj0 = S0.p0 + S1.p0, pj1 = S0.p1 + S1.p1;
      const double mc = mj0 + mj1, pc = pj0 + pj1;
      const double a1j = mj0 * B.v[0][0] + pj0 * B.v[0][2] + mj1 * B.v[2][0] + pj1 * B.v[2][2];
      const double a0j = (mj0 + pj0) * B.v[0][1] + (mj1 + pj1) * B.v[2][1];
      const double a1c = mc * B.v[1][0] + pc * B.v[1][2];
      const double s0 = q.f0c * (mc + pc);
      const double y = s0 * B.v[1][1] + q.f1c * a1c + q.f1j * a1j + q.f0j * a0j + ua + U(C, S1);
      const double val = rhs.ok() ? (rhs(i, j, k) - y) : y;
      out(i, j, k) = val;
      isnan_ |= (val != val);
      nrm = fmax(nrm, fabs(val));
      ua = U(B, S1);          // plane k seen from node k+1 through layer k
      B = C; C = D; S0 = S1; S1 = S2;
    }
    if (norm) {
      // non-negative doubles order like their bit patterns; a quiet NaN wins (blas.cu nanmax)
      unsigned long long m = isnan_ ? 0x7ff8000000000000ULL : (unsigned long long)__double_as_longlong(nrm);
      const unsigned act = __activemask();
      if (act == 0xffffffffu) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
      }
      if (act != 0xffffffffu || (threadIdx.x & 31) == 0) {
        unsigned long long* a = reinterpret_cast<unsigned long long*>(norm);
        if (m > *reinterpret_cast<volatile unsigned long long*>(a)) atomicMax(a, m);
      }
    }
  }
}
}  // namespace march
#endif

// The tile kernels measured SLOWER than the one-thread-per-node kernels on B200 (nodal GS colour
// pass at 257^3: 186 us vs 105 us; profiles/r01_notes.md), so they are opt-in (IAMRX_NODAL_TILE=1)
// and kept for the parity tests and further tuning.
inline bool use_tile_kernels() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("IAMRX_NODAL_TILE"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
inline dim3 grid_for(const Bx& bx) { return dim3(cdiv(bx.nx(), TX), cdiv(bx.ny(), TY), bx.nz()); }
inline void facs(const double dxinv[3], double f[3]) {
  for (int d = 0; d < 3; ++d) f[d] = (1.0 / 36.0) * dxinv[d] * dxinv[d];
}

}  // namespace

int nodal_divu(const Bx& nbx, V4 rhs, C4 vel, const double dxinv[3], cudaStream_t s, int hide) {
  if (!nbx.ok()) return IAMRX_OK;
  if (hide) {
    IX_LAUNCH(divu_bc_kernel, grid_for(nbx), dim3(TX, TY, 1), 0, s, nbx, rhs, vel, 0.25 * dxinv[0], 0.25 * dxinv[1], 0.25 * dxinv[2], hide);
    return check_launch("nodal_divu_bc");
  }
  IX_LAUNCH(divu_kernel, grid_for(nbx), dim3(TX, TY, 1), 0, s, nbx, rhs, vel, 0.25 * dxinv[0], 0.25 * dxinv[1],
                                                        0.25 * dxinv[2]);
  return check_launch("nodal_divu");
}

int nodal_adotx(const Bx& nbx, V4 out, C4 phi, C4 rhs, C4 sig, const double dxinv[3], cudaStream_t s, int wrapmask,
                double* norm_dev, bool* norm_fused) {
  if (norm_fused) *norm_fused = false;
  if (!nbx.ok()) return IAMRX_OK;
  double f[3]; facs(dxinv, f);
  ProfScope prof_(IAMRX_PROF_NODAL_ADOTX, nbx.npts(), (double)nbx.npts() * (rhs.ok() ? 32.0 : 24.0), s);
#if !defined(IX_EMUL)
  if (use_tile_kernels() && !(wrapmask >> 3) && nbx.nx() >= 32 && out.p != phi.p) {  // boxes wide enough to fill a CTA row
    using namespace tile;
    dim3 grd(cdiv(nbx.nx(), TXU), cdiv(nbx.ny(), TYU), nbx.nz());
    IX_LAUNCH((nodal_tile_kernel<1, false>), grd, dim3(TXU, TYU, 1), 0, s, nbx, out, phi, rhs, sig, f[0], f[1], f[2],
              nbx.lo[0], nbx.lo[1], nbx.lo[2], wrapmask);
    return check_launch("nodal_adotx_tile");
  }
#endif
#if !defined(IX_EMUL)
  {
    static int on = -1;
    if (on < 0) { const char* e = getenv("IAMRX_ADOTX_MARCH"); on = (e && e[0] == '0') ? 0 : 1; }
    if (on && nbx.nz() >= 8 && out.p != phi.p) {
      using namespace fused;
      Q1F q;
      q.f0c = q1_factor(false, false, false, f[0], f[1], f[2]); q.f1c = q1_factor(true, false, false, f[0], f[1], f[2]);
      q.f0j = q1_factor(false, true, false, f[0], f[1], f[2]);  q.f1j = q1_factor(true, true, false, f[0], f[1], f[2]);
      q.f0k = q1_factor(false, false, true, f[0], f[1], f[2]);  q.f1k = q1_factor(true, false, true, f[0], f[1], f[2]);
      q.f0jk = q1_factor(false, true, true, f[0], f[1], f[2]);  q.f1jk = q1_factor(true, true, true, f[0], f[1], f[2]);
      static int kb = -1, minb = -1;   // tuning knobs: planes per thread, resident CTAs per SM the kernel is compiled for
      if (kb < 0) { const char* e = getenv("IAMRX_ADOTX_KB"); kb = e ? atoi(e) : 32; }
      if (minb < 0) { const char* e = getenv("IAMRX_ADOTX_MINB"); minb = e ? atoi(e) : 2; }
      const int KBv = (kb >= 32) ? 32 : 16;
      const int nchunk = cdiv(nbx.nz(), KBv);
      const dim3 grd(cdiv(nbx.nx(), TX), cdiv(nbx.ny(), TY), nchunk), blk(TX, TY, 1);
      if (KBv == 32 && minb >= 3) IX_LAUNCH((march::adotx_march_kernel<32, 3>), grd, blk, 0, s, nbx, out, phi, rhs, sig, q, wrapmask, nchunk, norm_dev);
      else if (KBv == 32) IX_LAUNCH((march::adotx_march_kernel<32, 2>), grd, blk, 0, s, nbx, out, phi, rhs, sig, q, wrapmask, nchunk, norm_dev);
      else if (minb >= 3) IX_LAUNCH((march::adotx_march_kernel<16, 3>), grd, blk, 0, s, nbx, out, phi, rhs, sig, q, wrapmask, nchunk, norm_dev);
      else IX_LAUNCH((march::adotx_march_kernel<16, 2>), grd, blk, 0, s, nbx, out, phi, rhs, sig, q, wrapmask, nchunk, norm_dev);
      if (norm_fused) *norm_fused = norm_dev != nullptr;
      return check_launch("nodal_adotx_march");
    }
  }
#endif
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

#if !defined(IX_EMUL)
// Small node boxes (the coarse multigrid levels): all 8 colours of all sweeps in ONE single-CTA launch, colours separated by
// __syncthreads (in place, like eight nodal_gs_color launches per sweep).  Only for a level that is one box: every neighbour is a
// periodic image or a mirrored node, so nothing is exchanged between the colours.
namespace {
constexpr int GS_SMALL_NT = 1024;
__global__ void __launch_bounds__(GS_SMALL_NT)
gs_small_kernel(Bx bx, V4 phi, C4 rhs, C4 sig, double facx, double facy, double facz, int wm, int nsweeps) {
  C4 x{phi.p, phi.l0, phi.l1, phi.l2, phi.js, phi.ks, phi.ns};
  for (int sw = 0; sw < nsweeps; ++sw)
    for (int color = 0; color < 8; ++color) {
      const int c[3] = {color & 1, (color >> 1) & 1, (color >> 2) & 1};
      int o[3], n[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        o[d] = bx.lo[d] + ((c[d] - bx.lo[d]) & 1);
        n[d] = (bx.hi[d] >= o[d]) ? (bx.hi[d] - o[d]) / 2 + 1 : 0;
      }
      const int total = n[0] * n[1] * n[2];
      for (int idx = threadIdx.x; idx < total; idx += GS_SMALL_NT) {
        const int i = o[0] + 2 * (idx % n[0]), r = idx / n[0];
        const int j = o[1] + 2 * (r % n[1]), k = o[2] + 2 * (r / n[1]);
        double s0;
        NWRAP(bx, wm)
        const double y = nodal_ax(x, sig, i, j, k, im, ip, jm, jp, km, kp, facx, facy, facz, s0);
        phi(i, j, k) += (rhs(i, j, k) - y) / s0;
      }
      __syncthreads();
    }
}
}  // namespace
#endif
bool nodal_gs_small_ok(const Bx& nbx) {
#if defined(IX_EMUL)
  (void)nbx; return false;
#else
  static int on = -1;
  // opt-in: measured neutral (72.8 vs 72.6 ms per TaylorGreen 256^3 step with 247 fewer launches: the tiny launches were already
  // hidden behind each other), so the default keeps the path the host-emulated tests exercise as well
  if (on < 0) { const char* e = getenv("IAMRX_NODAL_SMALL"); on = (e && e[0] == '1') ? 1 : 0; }
  return on && nbx.npts() <= 5000;   // up to 17^3 nodes
#endif
}
int nodal_gs_small(const Bx& nbx, V4 phi, C4 rhs, C4 sig, const double dxinv[3], int nsweeps, cudaStream_t s, int wrapmask) {
  if (!nbx.ok() || nsweeps <= 0) return IAMRX_OK;
#if defined(IX_EMUL)
  (void)phi; (void)rhs; (void)sig; (void)dxinv; (void)s; (void)wrapmask;
  set_error("nodal_gs_small: not available in the host emulation"); return IAMRX_ERR_ARG;
#else
  double f[3]; facs(dxinv, f);
  IX_LAUNCH(gs_small_kernel, 1, GS_SMALL_NT, 0, s, nbx, phi, rhs, sig, f[0], f[1], f[2], wrapmask, nsweeps);
  return check_launch("nodal_gs_small");
#endif
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
#if !defined(IX_EMUL)
  {
    static int use_vec = -1;
    if (use_vec < 0) { const char* e = getenv("IAMRX_GS_VEC"); use_vec = (e && e[0] == '0') ? 0 : 1; }
    // 16-byte loads need even row / plane strides and 8-byte aligned bases whose parity we can read off
    const bool ok = use_vec && !(wrapmask >> 3) && n[0] >= 32 && (phi.js % 2 == 0) && (phi.ks % 2 == 0) && (sig.js % 2 == 0) && (sig.ks % 2 == 0);
    if (ok) {
      const uintptr_t a_phi = (uintptr_t)(phi.p + ((o[0] - phi.l0) + (o[1] - phi.l1) * phi.js + (o[2] - phi.l2) * phi.ks));
      const uintptr_t a_sig = (uintptr_t)(sig.p + ((o[0] - sig.l0) + (o[1] - sig.l1) * sig.js + (o[2] - sig.l2) * sig.ks));
      dim3 vg(cdiv(n[0], TX), cdiv(n[1], TY), n[2]);
      IX_LAUNCH(gs_color_vec_kernel, vg, dim3(TX, TY, 1), 0, s, nbx, phi, rhs, sig, f[0], f[1], f[2], o[0], o[1], o[2], wrapmask,
                (a_phi % 16 == 0) ? 1 : 0, (a_sig % 16 == 0) ? 1 : 0);
      return check_launch("nodal_gs_vec");
    }
  }
  if (use_tile_kernels() && !(wrapmask >> 3) && n[0] >= 16) {
    using namespace tile;
    dim3 tg(cdiv(n[0], TXU), cdiv(n[1], TYU), n[2]);
    C4 pin{phi.p, phi.l0, phi.l1, phi.l2, phi.js, phi.ks, phi.ns};
    IX_LAUNCH((nodal_tile_kernel<2, true>), tg, dim3(TXU, TYU, 1), 0, s, nbx, phi, pin, rhs, sig, f[0], f[1], f[2],
              o[0], o[1], o[2], wrapmask);
    return check_launch("nodal_gs_tile");
  }
#endif
  dim3 grd(cdiv(n[0], TX), cdiv(n[1], TY), n[2]);
  static int minb = -1;  // resident CTAs per SM the kernel is compiled for (register cap): tuning knob
  if (minb < 0) { const char* e = getenv("IAMRX_GS_MINB"); minb = e ? atoi(e) : 4; }
  if (minb >= 5) IX_LAUNCH(gs_color_kernel<5>, grd, dim3(TX, TY, 1), 0, s, nbx, phi, rhs, sig, f[0], f[1], f[2], o[0], o[1], o[2], wrapmask);
  else if (minb == 4) IX_LAUNCH(gs_color_kernel<4>, grd, dim3(TX, TY, 1), 0, s, nbx, phi, rhs, sig, f[0], f[1], f[2], o[0], o[1], o[2], wrapmask);
  else IX_LAUNCH(gs_color_kernel<3>, grd, dim3(TX, TY, 1), 0, s, nbx, phi, rhs, sig, f[0], f[1], f[2], o[0], o[1], o[2], wrapmask);
  return check_launch("nodal_gs_color");
}

// One full 8-colour sweep, phi_in -> phi_out (different arrays), on a node box that spans the
// periodic domain in all three directions with an even number of cells per direction.
bool nodal_gs_sweep_ok(const Bx& nbx, int wrapmask) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("IAMRX_NODAL_FUSED"); on = (e && e[0] == '0') ? 0 : 1; }
  if (!on) return false;
  if (wrapmask & NODAL_DEEP_GHOSTS) {   // x / y sides with neighbours: tile halos come from ghost layers of depth 4 (NodeMG deep levels)
    for (int d = 0; d < 2; ++d) if (!(wrapmask & (1 << d)) && ((nbx.lo[d] & 1) || nbx.hi[d] - nbx.lo[d] < 8)) return false;
  } else if ((wrapmask & 7) != 7 && (wrapmask & 7) != 3) {
    return false;
  }
  for (int d = 0; d < 3; ++d) if (((nbx.hi[d] - nbx.lo[d]) & 1) || nbx.hi[d] - nbx.lo[d] < 2) return false;
  return true;
}

int nodal_gs_sweep(const Bx& nbx, V4 phi_out, C4 phi_in, C4 rhs, C4 sig, const double dxinv[3], cudaStream_t s, int wrapmask,
                   int phase) {
  if (!nbx.ok()) return IAMRX_OK;
  double f[3]; facs(dxinv, f);
  C4 pout{phi_out.p, phi_out.l0, phi_out.l1, phi_out.l2, phi_out.js, phi_out.ks, phi_out.ns};
#if defined(IX_EMUL)
  // host emulation (tests only): the same sweep as eight in-place colour passes on a copy (ghost planes of the
  // exchanged direction included: the even-plane colours read them as old values)
  int rc = IAMRX_OK;
  const bool deep = (wrapmask & NODAL_DEEP_GHOSTS) != 0;
  Bx cb = (wrapmask & 4) ? nbx : grow(nbx, 2, 1);
  if (deep) for (int d = 0; d < 2; ++d) if (!(wrapmask & (1 << d))) cb = grow(cb, d, 4);
  if (phase != 1) rc = copy(cb, phi_out, phi_in, 1, s);
  const int c0 = (phase == 1) ? 4 : 0, c1 = (phase == 0) ? 4 : 8;
  (void)pout;
  for (int color = c0; color < c1 && rc == IAMRX_OK; ++color) {
    // deep ghosts: the halo nodes the later colours depend on are recomputed redundantly, as the tiles of the CUDA kernel do
    // (colour (cx, cy): 3 - cx - 2 cy ... nodes in x, 1 - cy in y; scratch values in the ghost layers, refilled by the caller)
    Bx ub = nbx;
    if (deep) {
      const int cx = color & 1, cy = (color >> 1) & 1;
      if (!(wrapmask & 1)) ub = grow(ub, 0, 3 - cx - 2 * cy);
      if (!(wrapmask & 2)) ub = grow(ub, 1, 1 - cy);
    }
    rc = nodal_gs_color(ub, phi_out, rhs, sig, dxinv, color, s, wrapmask & ~NODAL_DEEP_GHOSTS);
  }
  return rc;
#else
  using namespace fused;
  ProfScope prof_(IAMRX_PROF_NODAL_GS, nbx.npts(), (double)nbx.npts() * (phase < 0 ? 32.0 : 16.0), s);  // phi in + out, rhs, sigma
  Q1F q;
  q.f0c = q1_factor(false, false, false, f[0], f[1], f[2]); q.f1c = q1_factor(true, false, false, f[0], f[1], f[2]);
  q.f0j = q1_factor(false, true, false, f[0], f[1], f[2]);  q.f1j = q1_factor(true, true, false, f[0], f[1], f[2]);
  q.f0k = q1_factor(false, false, true, f[0], f[1], f[2]);  q.f1k = q1_factor(true, false, true, f[0], f[1], f[2]);
  q.f0jk = q1_factor(false, true, true, f[0], f[1], f[2]);  q.f1jk = q1_factor(true, true, true, f[0], f[1], f[2]);
  static int blk = -1, tall = -1;   // IAMRX_NODAL_BLOCK=0: one node per thread and colour (the first fused kernel); IAMRX_NODAL_TALL=1: 34-row tiles
  if (blk < 0) { const char* e = getenv("IAMRX_NODAL_BLOCK"); blk = (e && e[0] == '0') ? 0 : 1; }
  if (tall < 0) { const char* e = getenv("IAMRX_NODAL_TALL"); tall = (e && e[0] == '1') ? 1 : 0; }   // measured slower (293 vs 278 us per 257^3 sweep): opt-in
  constexpr int NRTALL = 34;
  static bool attr_set = false;
  if (!attr_set) {
    IX_CUDA(cudaFuncSetAttribute(gs_sweep_kernel<true, NRTALL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * NRTALL * NC * (int)sizeof(double)));
    attr_set = true;
  }
  const int ylen = nbx.hi[1] - (nbx.lo[1] - (nbx.lo[1] & 1)) + 1;
  const bool use_tall = blk && tall && ylen >= 4 * (NRTALL - 4);
  const int tyi = use_tall ? NRTALL - 4 : TYI;
  const int gx = cdiv(nbx.hi[0] - (nbx.lo[0] - (nbx.lo[0] & 1)) + 1, TXI), gy = cdiv(ylen, tyi);
  for (int cz = 0; cz < 2; ++cz) {
    if (phase >= 0 && phase != cz) continue;
    const int k0 = nbx.lo[2] + ((cz - nbx.lo[2]) & 1);
    if (k0 > nbx.hi[2]) continue;
    const int nk = (nbx.hi[2] - k0) / 2 + 1;
    // phase A (even planes): neighbours = old odd planes; phase B (odd planes): neighbours = new even planes
    const int xyg = (wrapmask & NODAL_DEEP_GHOSTS) ? (((wrapmask & 1) ? 0 : 1) | ((wrapmask & 2) ? 0 : 2)) : 0;
    const bool fast = xyg == 0 && nbx.hi[0] - nbx.lo[0] >= NC && nbx.hi[1] - nbx.lo[1] >= NC;
#define IX_GSW(B, R, F) IX_LAUNCH((gs_sweep_kernel<B, R, F>), dim3(gx, gy, nk), dim3(16 * (R - 2), 1, 1), 5 * R * NC * sizeof(double), s, nbx, phi_out, phi_in, \
                                  cz == 0 ? phi_in : pout, rhs, sig, q, k0, (wrapmask & 4) ? 1 : 0, (wrapmask >> 7) & 3, xyg)
    if (use_tall) IX_GSW(true, NRTALL, false);
    else if (blk && fast) IX_GSW(true, NR, true);
    else if (blk) IX_GSW(true, NR, false);
    else IX_GSW(false, NR, false);
#undef IX_GSW
    const int rc = check_launch("nodal_gs_sweep");
    if (rc != IAMRX_OK) return rc;
  }
  return IAMRX_OK;
#endif
}

int nodal_restrict(const Bx& cnbx, V4 crse, C4 fine, cudaStream_t s, int thin, int wm, const Bx* fnb) {
  if (!cnbx.ok()) return IAMRX_OK;
  IX_LAUNCH(nd_restrict_kernel, grid_for(cnbx), dim3(TX, TY, 1), 0, s, cnbx, crse, fine, thin, fnb ? (wm & 7) : 0, fnb ? *fnb : Bx{});
  return check_launch("nodal_restrict");
}

int nodal_interp_add(const Bx& fnbx, V4 fine, C4 crse, cudaStream_t s, int thin) {
  if (!fnbx.ok()) return IAMRX_OK;
  if (thin == 0) {
    Bx cb;   // coarse cells whose 2 x 2 x 2 fine nodes meet the fine box
    for (int d = 0; d < 3; ++d) { cb.lo[d] = fnbx.lo[d] >= 0 ? fnbx.lo[d] / 2 : -((-fnbx.lo[d] + 1) / 2); cb.hi[d] = fnbx.hi[d] >= 0 ? fnbx.hi[d] / 2 : -((-fnbx.hi[d] + 1) / 2); }
    IX_LAUNCH(nd_interp8_kernel, grid_for(cb), dim3(TX, TY, 1), 0, s, fnbx, cb, fine, crse);
    return check_launch("nodal_interp_add");
  }
  IX_LAUNCH(nd_interp_kernel, grid_for(fnbx), dim3(TX, TY, 1), 0, s, fnbx, fine, crse, thin);
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
