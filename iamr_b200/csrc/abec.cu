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
#include <cstdint>
#include <cstdlib>
#include <type_traits>
#include "kernels.h"

namespace ix {
namespace k {

namespace {

struct AbecDev {
  double a, b;
  C4 acoef, bx, by, bz;
  int bncomp;
  double dhx, dhy, dhz;  // b * dxinv^2
  double cb[3][3], ca;   // constant-coefficient path (Abec::cc, cac): values per component / direction
  int cac;
};

inline AbecDev to_dev(const Abec& op) {
  AbecDev d;
  d.a = op.a; d.b = op.b; d.acoef = op.acoef; d.bx = op.bx; d.by = op.by; d.bz = op.bz;
  d.bncomp = op.bncomp;
  d.dhx = op.b * op.dxinv[0] * op.dxinv[0];
  d.dhy = op.b * op.dxinv[1] * op.dxinv[1];
  d.dhz = op.b * op.dxinv[2] * op.dxinv[2];
  for (int n = 0; n < 3; ++n) for (int e = 0; e < 3; ++e) d.cb[n][e] = op.cb[n][e];
  d.ca = op.ca; d.cac = op.cac;
  return d;
}

// ---- GSRB -----------------------------------------------------------------
// One thread per updated cell: thread t of a row handles the cell pair
// (2t, 2t+1) relative to the box and picks the one of the right colour.
constexpr int GS_TX = 64;
constexpr int GS_TY = 4;

// HASBC: the box touches a non-periodic domain face.  The ghost cell behind such a face is a linear function of the interior
// cells (MLCellLinOp::applyBC); f0 is the coefficient of the adjacent cell, and the smoother removes that self-dependence
// from the diagonal (the delta of AMReX's abec_gsrb: phi += omega/(gamma - delta) * res).
// CONSTB: constant coefficients (Abec::cc) -- the same expression on values taken from the kernel arguments: 24 B/cell per
// colour pass (phi read + write, rhs) instead of 48 / 56.
// ZERO: first colour pass of a sweep on phi == 0 (ghost cells included: the multigrid correction with homogeneous boundary
// conditions).  No phi is read -- phi = omega / (gamma - delta) * rhs, the value the general expression gives -- and the other
// cell of the pair is set to zero, so the caller needs no setval before the sweep.
// the relaxation of ONE cell (i, j, k, n) of the right colour (shared by the colour-pass kernel and the small-box sweep kernel)
template <bool HASBC, bool CONSTB, bool ZERO>
IX_D void gsrb_cell(const Bx& bx, const V4& phi, const C4& rhs, const AbecDev& op, double omega, int wm, const GsBC& gb, int i, int j, int k, int n) {
  const int nb = (op.bncomp > 1) ? n : 0;

  // 32-bit element offsets from the cell's own address (one address computation per array)
  double bxm, bxp, bym, byp, bzm, bzp;
  if (CONSTB) {
    bxm = bxp = (nb == 0) ? op.cb[0][0] : (nb == 1 ? op.cb[1][0] : op.cb[2][0]);
    bym = byp = (nb == 0) ? op.cb[0][1] : (nb == 1 ? op.cb[1][1] : op.cb[2][1]);
    bzm = bzp = (nb == 0) ? op.cb[0][2] : (nb == 1 ? op.cb[1][2] : op.cb[2][2]);
  } else {
    const double* bxc = op.bx.p + nb * op.bx.ns + ((i - op.bx.l0) + (j - op.bx.l1) * op.bx.js + (k - op.bx.l2) * op.bx.ks);
    const double* byc = op.by.p + nb * op.by.ns + ((i - op.by.l0) + (j - op.by.l1) * op.by.js + (k - op.by.l2) * op.by.ks);
    const double* bzc = op.bz.p + nb * op.bz.ns + ((i - op.bz.l0) + (j - op.bz.l1) * op.bz.js + (k - op.bz.l2) * op.bz.ks);
    bxm = bxc[0]; bxp = bxc[1];
    bym = byc[0]; byp = byc[(int)op.by.js];
    bzm = bzc[0]; bzp = bzc[(int)op.bz.ks];
  }
  double* pc = phi.p + n * phi.ns + ((i - phi.l0) + (j - phi.l1) * phi.js + (k - phi.l2) * phi.ks);
  const int pjs = (int)phi.js, pks = (int)phi.ks;
  double gamma = op.dhx * (bxm + bxp) + op.dhy * (bym + byp) + op.dhz * (bzm + bzp);
  if (op.a != 0.0) gamma += op.a * ((CONSTB && op.cac) ? op.ca : op.acoef(i, j, k));
  if (ZERO) {
    double delta = 0.0;
    if (HASBC) {
      const int nc = n < 3 ? n : 0;
      if (i == bx.lo[0]) delta += op.dhx * bxm * gb.f0[nc][0];
      if (i == bx.hi[0]) delta += op.dhx * bxp * gb.f0[nc][1];
      if (j == bx.lo[1]) delta += op.dhy * bym * gb.f0[nc][2];
      if (j == bx.hi[1]) delta += op.dhy * byp * gb.f0[nc][3];
      if (k == bx.lo[2]) delta += op.dhz * bzm * gb.f0[nc][4];
      if (k == bx.hi[2]) delta += op.dhz * bzp * gb.f0[nc][5];
    }
    pc[0] = omega / (HASBC ? (gamma - delta) : gamma) * rhs(i, j, k, n);
    return;
  }
  const double p0 = pc[0];
  // periodic wrap inside the kernel when the box spans the domain (no ghost fill needed)
  int oxm = ((wm & 1) && i == bx.lo[0]) ? bx.hi[0] - i : -1, oxp = ((wm & 1) && i == bx.hi[0]) ? bx.lo[0] - i : 1;
  int oym = (((wm & 2) && j == bx.lo[1]) ? bx.hi[1] - j : -1) * pjs, oyp = (((wm & 2) && j == bx.hi[1]) ? bx.lo[1] - j : 1) * pjs;
  int ozm = (((wm & 4) && k == bx.lo[2]) ? bx.hi[2] - k : -1) * pks, ozp = (((wm & 4) && k == bx.hi[2]) ? bx.lo[2] - k : 1) * pks;
  double rho;
  if (HASBC) {
    // sides whose homogeneous ghost cell is +/- the adjacent cell (Neumann, reflect_odd, order-2 Dirichlet) are evaluated
    // in place: no ghost fill between the colours
    const int nc = n < 3 ? n : 0;
    const int mir = gb.even[nc] | gb.odd[nc], od = gb.odd[nc];
    double sxm = 1.0, sxp = 1.0, sym = 1.0, syp = 1.0, szm = 1.0, szp = 1.0;
    if ((mir & 1) && i == bx.lo[0]) { oxm = 0; if (od & 1) sxm = -1.0; }
    if ((mir & 2) && i == bx.hi[0]) { oxp = 0; if (od & 2) sxp = -1.0; }
    if ((mir & 4) && j == bx.lo[1]) { oym = 0; if (od & 4) sym = -1.0; }
    if ((mir & 8) && j == bx.hi[1]) { oyp = 0; if (od & 8) syp = -1.0; }
    if ((mir & 16) && k == bx.lo[2]) { ozm = 0; if (od & 16) szm = -1.0; }
    if ((mir & 32) && k == bx.hi[2]) { ozp = 0; if (od & 32) szp = -1.0; }
    rho = op.dhx * (bxm * (sxm * pc[oxm]) + bxp * (sxp * pc[oxp])) + op.dhy * (bym * (sym * pc[oym]) + byp * (syp * pc[oyp])) +
          op.dhz * (bzm * (szm * pc[ozm]) + bzp * (szp * pc[ozp]));
  } else {
    rho = op.dhx * (bxm * pc[oxm] + bxp * pc[oxp]) +
          op.dhy * (bym * pc[oym] + byp * pc[oyp]) +
          op.dhz * (bzm * pc[ozm] + bzp * pc[ozp]);
  }
  const double res = rhs(i, j, k, n) - (gamma * p0 - rho);
  if (HASBC) {
    const int nc = n < 3 ? n : 0;
    double delta = 0.0;
    if (i == bx.lo[0]) delta += op.dhx * bxm * gb.f0[nc][0];
    if (i == bx.hi[0]) delta += op.dhx * bxp * gb.f0[nc][1];
    if (j == bx.lo[1]) delta += op.dhy * bym * gb.f0[nc][2];
    if (j == bx.hi[1]) delta += op.dhy * byp * gb.f0[nc][3];
    if (k == bx.lo[2]) delta += op.dhz * bzm * gb.f0[nc][4];
    if (k == bx.hi[2]) delta += op.dhz * bzp * gb.f0[nc][5];
    pc[0] = p0 + omega / (gamma - delta) * res;
  } else {
    pc[0] = p0 + omega / gamma * res;
  }
}

template <int MINB, bool HASBC, bool CONSTB, bool ZERO>
__global__ void __launch_bounds__(GS_TX* GS_TY, MINB)
gsrb_kernel(Bx bx, V4 phi, C4 rhs, IX_KARG(AbecDev) op, double omega, int redblack, int nz, int wm, IX_KARG(GsBC) gb) {
  const int kz = blockIdx.z % nz;
  const int n = blockIdx.z / nz;
  const int k = bx.lo[2] + kz;
  const int j = bx.lo[1] + blockIdx.y * GS_TY + threadIdx.y;
  if (j > bx.hi[1]) return;
  int i = bx.lo[0] + 2 * (blockIdx.x * GS_TX + threadIdx.x);
  const int ipair = i;
  // make (i + j + k + redblack) even
  i += (i + j + k + redblack) & 1;
  if (ZERO) {   // the pair's cell of the other colour (or both, if the coloured one lies beyond the box)
    const int io = ipair + (1 - (i - ipair));
    if (io <= bx.hi[0]) phi(io, j, k, n) = 0.0;
  }
  if (i > bx.hi[0]) return;
  gsrb_cell<HASBC, CONSTB, ZERO>(bx, phi, rhs, op, omega, wm, gb, i, j, k, n);
}

// Small boxes (the coarse multigrid levels: a few thousand cells): ALL sweeps of a smoothing step in one launch of one CTA,
// colours separated by __syncthreads -- 2 nsweeps launches of a few microseconds each become one.  Only for boxes whose
// neighbours are all reached inside the kernel (periodic wrap / mirrored sides): no ghost exchange between the colours.
#if !defined(IX_EMUL)
constexpr int GS_SMALL_T = 512;
template <bool HASBC, bool CONSTB>
__global__ void __launch_bounds__(GS_SMALL_T)
gsrb_small_kernel(Bx bx, V4 phi, C4 rhs, IX_KARG(AbecDev) op, double omega, int ncomp, int wm, IX_KARG(GsBC) gb, int nsweeps, int zero_first) {
  const int hx = (bx.nx() + 1) / 2, ny = bx.ny(), nz = bx.nz();
  const int npairs = hx * ny * nz * ncomp;
  for (int sw = 0; sw < nsweeps; ++sw)
    for (int rb = 0; rb < 2; ++rb) {
      const bool zero = zero_first && sw == 0 && rb == 0;
      for (int idx = threadIdx.x; idx < npairs; idx += GS_SMALL_T) {
        const int ph = idx % hx, r1 = idx / hx;
        const int j = bx.lo[1] + r1 % ny, r2 = r1 / ny;
        const int k = bx.lo[2] + r2 % nz, n = r2 / nz;
        const int ipair = bx.lo[0] + 2 * ph;
        const int i = ipair + ((ipair + j + k + rb) & 1);
        if (zero) {
          const int io = ipair + (1 - (i - ipair));
          if (io <= bx.hi[0]) phi(io, j, k, n) = 0.0;
          if (i <= bx.hi[0]) gsrb_cell<HASBC, CONSTB, true>(bx, phi, rhs, op, omega, wm, gb, i, j, k, n);
        } else if (i <= bx.hi[0]) {
          gsrb_cell<HASBC, CONSTB, false>(bx, phi, rhs, op, omega, wm, gb, i, j, k, n);
        }
      }
      __syncthreads();
    }
}
#endif

// ---- apply / residual ----------------------------------------------------
constexpr int AP_TX = 128;
constexpr int AP_TY = 2;

// mir: per component, bit s (xlo, xhi, ylo, yhi, zlo, zhi) of even / odd = the ghost cell beyond that side of the box is
// + / - the adjacent cell (homogeneous Neumann / reflect_odd / order-2 Dirichlet): evaluated in place, no ghost fill
struct MirBC { int even[3], odd[3]; };
template <bool CONSTB>
__global__ void __launch_bounds__(AP_TX* AP_TY)
apply_kernel(Bx bx, V4 out, C4 phi, C4 rhs, IX_KARG(AbecDev) op, int nz, int wm, IX_KARG(MirBC) mb) {
  const int kz = blockIdx.z % nz;
  const int n = blockIdx.z / nz;
  const int k = bx.lo[2] + kz;
  const int j = bx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = bx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > bx.hi[1] || i > bx.hi[0]) return;
  const int nb = (op.bncomp > 1) ? n : 0;
  double bxm, bxp, bym, byp, bzm, bzp;
  if (CONSTB) {
    bxm = bxp = (nb == 0) ? op.cb[0][0] : (nb == 1 ? op.cb[1][0] : op.cb[2][0]);
    bym = byp = (nb == 0) ? op.cb[0][1] : (nb == 1 ? op.cb[1][1] : op.cb[2][1]);
    bzm = bzp = (nb == 0) ? op.cb[0][2] : (nb == 1 ? op.cb[1][2] : op.cb[2][2]);
  } else {
    const double* bxc = op.bx.p + nb * op.bx.ns + ((i - op.bx.l0) + (j - op.bx.l1) * op.bx.js + (k - op.bx.l2) * op.bx.ks);
    const double* byc = op.by.p + nb * op.by.ns + ((i - op.by.l0) + (j - op.by.l1) * op.by.js + (k - op.by.l2) * op.by.ks);
    const double* bzc = op.bz.p + nb * op.bz.ns + ((i - op.bz.l0) + (j - op.bz.l1) * op.bz.js + (k - op.bz.l2) * op.bz.ks);
    bxm = bxc[0]; bxp = bxc[1];
    bym = byc[0]; byp = byc[(int)op.by.js];
    bzm = bzc[0]; bzp = bzc[(int)op.bz.ks];
  }
  const double* pc = phi.p + n * phi.ns + ((i - phi.l0) + (j - phi.l1) * phi.js + (k - phi.l2) * phi.ks);
  const int pjs = (int)phi.js, pks = (int)phi.ks;
  const double p0 = pc[0];
  int oxm = ((wm & 1) && i == bx.lo[0]) ? bx.hi[0] - i : -1, oxp = ((wm & 1) && i == bx.hi[0]) ? bx.lo[0] - i : 1;
  int oym = (((wm & 2) && j == bx.lo[1]) ? bx.hi[1] - j : -1) * pjs, oyp = (((wm & 2) && j == bx.hi[1]) ? bx.lo[1] - j : 1) * pjs;
  int ozm = (((wm & 4) && k == bx.lo[2]) ? bx.hi[2] - k : -1) * pks, ozp = (((wm & 4) && k == bx.hi[2]) ? bx.lo[2] - k : 1) * pks;
  double sxm = 1.0, sxp = 1.0, sym = 1.0, syp = 1.0, szm = 1.0, szp = 1.0;
  {
    const int nc = n < 3 ? n : 0;
    const int mir = mb.even[nc] | mb.odd[nc], od = mb.odd[nc];
    if (mir) {
      if ((mir & 1) && i == bx.lo[0]) { oxm = 0; if (od & 1) sxm = -1.0; }
      if ((mir & 2) && i == bx.hi[0]) { oxp = 0; if (od & 2) sxp = -1.0; }
      if ((mir & 4) && j == bx.lo[1]) { oym = 0; if (od & 4) sym = -1.0; }
      if ((mir & 8) && j == bx.hi[1]) { oyp = 0; if (od & 8) syp = -1.0; }
      if ((mir & 16) && k == bx.lo[2]) { ozm = 0; if (od & 16) szm = -1.0; }
      if ((mir & 32) && k == bx.hi[2]) { ozp = 0; if (od & 32) szp = -1.0; }
    }
  }
  double y = -op.dhx * (bxp * (sxp * pc[oxp] - p0) - bxm * (p0 - sxm * pc[oxm])) -
             op.dhy * (byp * (syp * pc[oyp] - p0) - bym * (p0 - sym * pc[oym])) -
             op.dhz * (bzp * (szp * pc[ozp] - p0) - bzm * (p0 - szm * pc[ozm]));
  if (op.a != 0.0) y += op.a * ((CONSTB && op.cac) ? op.ca : op.acoef(i, j, k)) * p0;
  out(i, j, k, n) = rhs.ok() ? (rhs(i, j, k, n) - y) : y;
}

#if !defined(IX_EMUL)
// apply / residual, two cells per thread with 128-bit loads and stores (same expression per cell as apply_kernel:
// bit-identical).  Needs an even x extent and 16-byte aligned cell pairs in every array (checked by the launcher).
IX_D double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
// NORM: also max |out| -> *norm (the residual norm of the multigrid iteration, fused: saves a pass over `out`).  Needs whole
// warps (the launcher checks nx / 2 % 32 == 0); a NaN wins the unsigned atomicMax (blas.cu nanmax).
template <bool HASA, bool CONSTB, bool NORM>
__global__ void __launch_bounds__(AP_TX* AP_TY)
apply2_kernel(Bx bx, V4 out, C4 phi, C4 rhs, IX_KARG(AbecDev) op, int nz, int wm, double* norm) {
  const int kz = blockIdx.z % nz;
  const int n = blockIdx.z / nz;
  const int k = bx.lo[2] + kz;
  const int j = bx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = bx.lo[0] + 2 * (blockIdx.x * AP_TX + threadIdx.x);
  if (j > bx.hi[1] || i > bx.hi[0]) return;
  const int nb = (op.bncomp > 1) ? n : 0;
  const double* pc = phi.p + n * phi.ns + ((i - phi.l0) + (j - phi.l1) * phi.js + (k - phi.l2) * phi.ks);
  const int pjs = (int)phi.js, pks = (int)phi.ks;
  const int oxm = ((wm & 1) && i == bx.lo[0]) ? bx.hi[0] - i : -1, oxp = ((wm & 1) && i + 1 == bx.hi[0]) ? bx.lo[0] - i : 2;
  const int oym = (((wm & 2) && j == bx.lo[1]) ? bx.hi[1] - j : -1) * pjs, oyp = (((wm & 2) && j == bx.hi[1]) ? bx.lo[1] - j : 1) * pjs;
  const int ozm = (((wm & 4) && k == bx.lo[2]) ? bx.hi[2] - k : -1) * pks, ozp = (((wm & 4) && k == bx.hi[2]) ? bx.lo[2] - k : 1) * pks;
  const double2 p0 = ld2(pc), pym = ld2(pc + oym), pyp = ld2(pc + oyp), pzm = ld2(pc + ozm), pzp = ld2(pc + ozp);
  const double pxm = pc[oxm], pxp = pc[oxp];
  double2 b01, bym, byp, bzm, bzp;
  double b2;
  if (CONSTB) {
    const double cx = (nb == 0) ? op.cb[0][0] : (nb == 1 ? op.cb[1][0] : op.cb[2][0]);
    const double cy = (nb == 0) ? op.cb[0][1] : (nb == 1 ? op.cb[1][1] : op.cb[2][1]);
    const double cz = (nb == 0) ? op.cb[0][2] : (nb == 1 ? op.cb[1][2] : op.cb[2][2]);
    b01 = make_double2(cx, cx); b2 = cx;
    bym = byp = make_double2(cy, cy);
    bzm = bzp = make_double2(cz, cz);
  } else {
    const double* bxc = op.bx.p + nb * op.bx.ns + ((i - op.bx.l0) + (j - op.bx.l1) * op.bx.js + (k - op.bx.l2) * op.bx.ks);
    const double* byc = op.by.p + nb * op.by.ns + ((i - op.by.l0) + (j - op.by.l1) * op.by.js + (k - op.by.l2) * op.by.ks);
    const double* bzc = op.bz.p + nb * op.bz.ns + ((i - op.bz.l0) + (j - op.bz.l1) * op.bz.js + (k - op.bz.l2) * op.bz.ks);
    b01 = ld2(bxc); b2 = bxc[2];
    bym = ld2(byc); byp = ld2(byc + (int)op.by.js); bzm = ld2(bzc); bzp = ld2(bzc + (int)op.bz.ks);
  }
  double y0 = -op.dhx * (b01.y * (p0.y - p0.x) - b01.x * (p0.x - pxm)) - op.dhy * (byp.x * (pyp.x - p0.x) - bym.x * (p0.x - pym.x)) -
              op.dhz * (bzp.x * (pzp.x - p0.x) - bzm.x * (p0.x - pzm.x));
  double y1 = -op.dhx * (b2 * (pxp - p0.y) - b01.y * (p0.y - p0.x)) - op.dhy * (byp.y * (pyp.y - p0.y) - bym.y * (p0.y - pym.y)) -
              op.dhz * (bzp.y * (pzp.y - p0.y) - bzm.y * (p0.y - pzm.y));
  if (HASA) {
    const double2 ac = (CONSTB && op.cac) ? make_double2(op.ca, op.ca)
                              : ld2(op.acoef.p + ((i - op.acoef.l0) + (j - op.acoef.l1) * op.acoef.js + (k - op.acoef.l2) * op.acoef.ks));
    y0 += op.a * ac.x * p0.x; y1 += op.a * ac.y * p0.y;
  }
  double2 r;
  if (rhs.ok()) {
    const double2 rh = ld2(rhs.p + n * rhs.ns + ((i - rhs.l0) + (j - rhs.l1) * rhs.js + (k - rhs.l2) * rhs.ks));
    r.x = rh.x - y0; r.y = rh.y - y1;
  } else { r.x = y0; r.y = y1; }
  *reinterpret_cast<double2*>(out.p + n * out.ns + ((i - out.l0) + (j - out.l1) * out.js + (k - out.l2) * out.ks)) = r;
  if (NORM) {
    const double ax = fabs(r.x), ay = fabs(r.y);
    const unsigned long long qnan = 0x7ff8000000000000ULL;
    unsigned long long m = (ax != ax || ay != ay) ? qnan : (unsigned long long)__double_as_longlong(ax > ay ? ax : ay);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }   // non-negative doubles order like their bits
    if ((threadIdx.x & 31) == 0) {
      unsigned long long* a = reinterpret_cast<unsigned long long*>(norm);
      if (m > *reinterpret_cast<volatile unsigned long long*>(a)) atomicMax(a, m);
    }
  }
}

// every (lo0 + 2m, j, k, n) element of the view is 16-byte aligned
template <class V> inline bool pairs_aligned(const V& v, const Bx& bx) {
  if (!v.p) return true;
  const double* q = v.p + ((bx.lo[0] - v.l0) + (bx.lo[1] - v.l1) * v.js + (bx.lo[2] - v.l2) * v.ks);
  return ((uintptr_t)q % 16 == 0) && v.js % 2 == 0 && v.ks % 2 == 0 && v.ns % 2 == 0;
}
#endif

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

// thin: bit d set = direction d is NOT coarsened between the two levels (semi-coarsening of thin, "2-D" domains: ratio 1)
__global__ void restrict_kernel(Bx cbx, V4 crse, C4 fine, int nz, int thin) {
  const int kz = blockIdx.z % nz;
  const int n = blockIdx.z / nz;
  const int k = cbx.lo[2] + kz;
  const int j = cbx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = cbx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > cbx.hi[1] || i > cbx.hi[0]) return;
  if (thin == 0) {
    const int ii = 2 * i, jj = 2 * j, kk = 2 * k;
    crse(i, j, k, n) = 0.125 * (fine(ii, jj, kk, n) + fine(ii + 1, jj, kk, n) + fine(ii, jj + 1, kk, n) +
                                fine(ii + 1, jj + 1, kk, n) + fine(ii, jj, kk + 1, n) +
                                fine(ii + 1, jj, kk + 1, n) + fine(ii, jj + 1, kk + 1, n) +
                                fine(ii + 1, jj + 1, kk + 1, n));
    return;
  }
  const int r0 = (thin & 1) ? 1 : 2, r1 = (thin & 2) ? 1 : 2, r2 = (thin & 4) ? 1 : 2;
  double acc = 0.0;
  for (int dk = 0; dk < r2; ++dk) for (int dj = 0; dj < r1; ++dj) for (int di = 0; di < r0; ++di) acc += fine(r0 * i + di, r1 * j + dj, r2 * k + dk, n);
  crse(i, j, k, n) = acc / (double)(r0 * r1 * r2);
}

IX_D int cdiv2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }  // floor(a/2)

// two fine planes per thread (2 m, 2 m + 1 relative to the box: the same coarse plane unless z is not coarsened)
__global__ void prolong_kernel(Bx fbx, V4 fine, C4 crse, int nzh, int thin) {
  const int kz = blockIdx.z % nzh;
  const int n = blockIdx.z / nzh;
  const int j = fbx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = fbx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > fbx.hi[1] || i > fbx.hi[0]) return;
  const int ic = (thin & 1) ? i : cdiv2(i), jc = (thin & 2) ? j : cdiv2(j);
  const int k0 = fbx.lo[2] + 2 * kz, k1 = k0 + 1;
  const bool two = k1 <= fbx.hi[2];
  const double c0 = crse(ic, jc, (thin & 4) ? k0 : cdiv2(k0), n);
  const double c1 = two ? crse(ic, jc, (thin & 4) ? k1 : cdiv2(k1), n) : 0.0;
  const double f0 = fine(i, j, k0, n);
  const double f1 = two ? fine(i, j, k1, n) : 0.0;
  fine(i, j, k0, n) = f0 + c0;
  if (two) fine(i, j, k1, n) = f1 + c1;
}

__global__ void face_restrict_kernel(Bx cfbx, int dir, V4 crse, C4 fine, int nz, int thin) {
  const int kz = blockIdx.z % nz;
  const int n = blockIdx.z / nz;
  const int k = cfbx.lo[2] + kz;
  const int j = cfbx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = cfbx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > cfbx.hi[1] || i > cfbx.hi[0]) return;
  if (thin != 0) {   // mean of the fine faces on the coarse face: the transverse directions that are coarsened contribute 2 each
    const int r[3] = {(thin & 1) ? 1 : 2, (thin & 2) ? 1 : 2, (thin & 4) ? 1 : 2};
    const int c[3] = {i, j, k};
    const int t0 = (dir + 1) % 3, t1 = (dir + 2) % 3;
    double acc = 0.0;
    for (int b = 0; b < r[t1]; ++b)
      for (int a_ = 0; a_ < r[t0]; ++a_) {
        int f[3];
        f[dir] = r[dir] * c[dir]; f[t0] = r[t0] * c[t0] + a_; f[t1] = r[t1] * c[t1] + b;
        acc += fine(f[0], f[1], f[2], n);
      }
    crse(i, j, k, n) = acc / (double)(r[t0] * r[t1]);
    return;
  }
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
// transverse derivatives on a face from the four cells around the face's edge mid-points; the two cells on either side of the
// face along its normal (a = below, b = above) and the neighbour indices (m / p) are passed explicitly so that a caller can hand
// in periodic images (in-kernel wrap) instead of ghost cells
IX_D double dy_on_xw(C4 v, int ia, int ib, int jm, int jp, int k, int n, double dyi) {
  return (v(ib, jp, k, n) + v(ia, jp, k, n) - v(ib, jm, k, n) - v(ia, jm, k, n)) * (0.25 * dyi);
}
IX_D double dz_on_xw(C4 v, int ia, int ib, int j, int km, int kp, int n, double dzi) {
  return (v(ib, j, kp, n) + v(ia, j, kp, n) - v(ib, j, km, n) - v(ia, j, km, n)) * (0.25 * dzi);
}
IX_D double dx_on_yw(C4 v, int im, int ip, int ja, int jb, int k, int n, double dxi) {
  return (v(ip, jb, k, n) + v(ip, ja, k, n) - v(im, jb, k, n) - v(im, ja, k, n)) * (0.25 * dxi);
}
IX_D double dz_on_yw(C4 v, int i, int ja, int jb, int km, int kp, int n, double dzi) {
  return (v(i, jb, kp, n) + v(i, ja, kp, n) - v(i, jb, km, n) - v(i, ja, km, n)) * (0.25 * dzi);
}
IX_D double dx_on_zw(C4 v, int im, int ip, int j, int ka, int kb, int n, double dxi) {
  return (v(ip, j, kb, n) + v(ip, j, ka, n) - v(im, j, kb, n) - v(im, j, ka, n)) * (0.25 * dxi);
}
IX_D double dy_on_zw(C4 v, int i, int jm, int jp, int ka, int kb, int n, double dyi) {
  return (v(i, jp, kb, n) + v(i, jp, ka, n) - v(i, jm, kb, n) - v(i, jm, ka, n)) * (0.25 * dyi);
}
IX_D double dy_on_x(C4 v, int i, int j, int k, int n, double dyi) { return dy_on_xw(v, i - 1, i, j - 1, j + 1, k, n, dyi); }
IX_D double dz_on_x(C4 v, int i, int j, int k, int n, double dzi) { return dz_on_xw(v, i - 1, i, j, k - 1, k + 1, n, dzi); }
IX_D double dx_on_y(C4 v, int i, int j, int k, int n, double dxi) { return dx_on_yw(v, i - 1, i + 1, j - 1, j, k, n, dxi); }
IX_D double dz_on_y(C4 v, int i, int j, int k, int n, double dzi) { return dz_on_yw(v, i, j - 1, j, k - 1, k + 1, n, dzi); }
IX_D double dx_on_z(C4 v, int i, int j, int k, int n, double dxi) { return dx_on_zw(v, i - 1, i + 1, j, k - 1, k, n, dxi); }
IX_D double dy_on_z(C4 v, int i, int j, int k, int n, double dyi) { return dy_on_zw(v, i, j - 1, j + 1, k - 1, k, n, dyi); }

// the cross fluxes with explicit neighbour indices: face between cells (ia, ib) along the normal; mu = the face coefficient
IX_D void cross_fxw(C4 vel, double mu, int ia, int ib, int jm, int j, int jp, int km, int k, int kp, double dyi, double dzi, double f[3]) {
  const double dudy = dy_on_xw(vel, ia, ib, jm, jp, k, 0, dyi);
  const double dvdy = dy_on_xw(vel, ia, ib, jm, jp, k, 1, dyi);
  const double dudz = dz_on_xw(vel, ia, ib, j, km, kp, 0, dzi);
  const double dwdz = dz_on_xw(vel, ia, ib, j, km, kp, 2, dzi);
  const double divu = dvdy + dwdz;
  f[0] = -mu * (-(2.0 / 3.0) * divu);
  f[1] = -mu * dudy;
  f[2] = -mu * dudz;
}
IX_D void cross_fyw(C4 vel, double mu, int im, int i, int ip, int ja, int jb, int km, int k, int kp, double dxi, double dzi, double f[3]) {
  const double dudx = dx_on_yw(vel, im, ip, ja, jb, k, 0, dxi);
  const double dvdx = dx_on_yw(vel, im, ip, ja, jb, k, 1, dxi);
  const double dvdz = dz_on_yw(vel, i, ja, jb, km, kp, 1, dzi);
  const double dwdz = dz_on_yw(vel, i, ja, jb, km, kp, 2, dzi);
  const double divu = dudx + dwdz;
  f[0] = -mu * dvdx;
  f[1] = -mu * (-(2.0 / 3.0) * divu);
  f[2] = -mu * dvdz;
}
IX_D void cross_fzw(C4 vel, double mu, int im, int i, int ip, int jm, int j, int jp, int ka, int kb, double dxi, double dyi, double f[3]) {
  const double dudx = dx_on_zw(vel, im, ip, j, ka, kb, 0, dxi);
  const double dwdx = dx_on_zw(vel, im, ip, j, ka, kb, 2, dxi);
  const double dvdy = dy_on_zw(vel, i, jm, jp, ka, kb, 1, dyi);
  const double dwdy = dy_on_zw(vel, i, jm, jp, ka, kb, 2, dyi);
  const double divu = dudx + dvdy;
  f[0] = -mu * dwdx;
  f[1] = -mu * dwdy;
  f[2] = -mu * (-(2.0 / 3.0) * divu);
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

// MINB = resident CTAs per SM the kernel is compiled for: uncapped it takes 100 registers (two CTAs, 22 % of the warps) and is
// latency-bound (ncu: issue 25 %, l1tex 25 %, 1.9 TB/s)
template <int MINB>
__global__ void __launch_bounds__(AP_TX* AP_TY, MINB)
tensor_cross_kernel(Bx bx, V4 out, C4 vel, C4 ex, C4 ey, C4 ez, double b, double dxi, double dyi,
                    double dzi, int wm) {
  const int k = bx.lo[2] + blockIdx.z;
  const int j = bx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = bx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > bx.hi[1] || i > bx.hi[0]) return;
  // wm bit d: the box spans the periodic domain in direction d -- the neighbours beyond it are its own cells on the other side
  // (no ghost fill of edges and corners needed there)
  const int im = ((wm & 1) && i == bx.lo[0]) ? bx.hi[0] : i - 1, ip = ((wm & 1) && i == bx.hi[0]) ? bx.lo[0] : i + 1;
  const int jm = ((wm & 2) && j == bx.lo[1]) ? bx.hi[1] : j - 1, jp = ((wm & 2) && j == bx.hi[1]) ? bx.lo[1] : j + 1;
  const int km = ((wm & 4) && k == bx.lo[2]) ? bx.hi[2] : k - 1, kp = ((wm & 4) && k == bx.hi[2]) ? bx.lo[2] : k + 1;
  double fl[3], fh[3], acc[3];
  cross_fxw(vel, ex(i, j, k), im, i, jm, j, jp, km, k, kp, dyi, dzi, fl);
  cross_fxw(vel, ex(i + 1, j, k), i, ip, jm, j, jp, km, k, kp, dyi, dzi, fh);
  for (int n = 0; n < 3; ++n) acc[n] = dxi * (fh[n] - fl[n]);
  cross_fyw(vel, ey(i, j, k), im, i, ip, jm, j, km, k, kp, dxi, dzi, fl);
  cross_fyw(vel, ey(i, j + 1, k), im, i, ip, j, jp, km, k, kp, dxi, dzi, fh);
  for (int n = 0; n < 3; ++n) acc[n] += dyi * (fh[n] - fl[n]);
  cross_fzw(vel, ez(i, j, k), im, i, ip, jm, j, jp, km, k, dxi, dyi, fl);
  cross_fzw(vel, ez(i, j, k + 1), im, i, ip, jm, j, jp, k, kp, dxi, dyi, fh);
  for (int n = 0; n < 3; ++n) acc[n] += dzi * (fh[n] - fl[n]);
  for (int n = 0; n < 3; ++n) out(i, j, k, n) += b * acc[n];
}

#if !defined(IX_EMUL)
// z-marching form of tensor_cross_kernel: a CTA owns a 32 x 8 column of cells and walks up in z.  Every thread computes only the
// three LOW-face cross fluxes of its cell (the x / y neighbours' come through shared memory, the z one is carried in registers to
// the next plane), i.e. each face flux is evaluated once instead of twice and the stencil loads per cell halve.
constexpr int TC_X = 32, TC_Y = 8, TC_KB = 32;
__global__ void __launch_bounds__((TC_X + 1) * (TC_Y + 1))
tensor_cross_march_kernel(Bx bx, V4 out, C4 vel, C4 ex, C4 ey, C4 ez, double b, double dxi, double dyi, double dzi) {
  __shared__ double sfx[3][TC_Y + 1][TC_X + 1], sfy[3][TC_Y + 1][TC_X + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = bx.lo[0] + blockIdx.x * TC_X + tx, j = bx.lo[1] + blockIdx.y * TC_Y + ty;
  const int k0 = bx.lo[2] + blockIdx.z * TC_KB, k1 = min(k0 + TC_KB - 1, bx.hi[2]);
  // threads of the extra column / row only provide the x / y flux of the face beyond the tile
  const bool in_x = tx < TC_X && i <= bx.hi[0], in_y = ty < TC_Y && j <= bx.hi[1];
  const bool need_fx = i <= bx.hi[0] + 1 && in_y, need_fy = j <= bx.hi[1] + 1 && in_x, cell = in_x && in_y;
  double fzl[3] = {0.0, 0.0, 0.0}, accxy[3] = {0.0, 0.0, 0.0};
  if (cell) cross_fz(vel, ez, i, j, k0, dxi, dyi, fzl);
  for (int k = k0; k <= k1; ++k) {
    double f[3];
    if (need_fx) { cross_fx(vel, ex, i, j, k, dyi, dzi, f); sfx[0][ty][tx] = f[0]; sfx[1][ty][tx] = f[1]; sfx[2][ty][tx] = f[2]; }
    if (need_fy) { cross_fy(vel, ey, i, j, k, dxi, dzi, f); sfy[0][ty][tx] = f[0]; sfy[1][ty][tx] = f[1]; sfy[2][ty][tx] = f[2]; }
    __syncthreads();
    if (cell) {
#pragma unroll
      for (int n = 0; n < 3; ++n) accxy[n] = dxi * (sfx[n][ty][tx + 1] - sfx[n][ty][tx]) + dyi * (sfy[n][ty + 1][tx] - sfy[n][ty][tx]);
      double fzh[3];
      cross_fz(vel, ez, i, j, k + 1, dxi, dyi, fzh);
#pragma unroll
      for (int n = 0; n < 3; ++n) { out(i, j, k, n) += b * (accxy[n] + dzi * (fzh[n] - fzl[n])); fzl[n] = fzh[n]; }
    }
    __syncthreads();
  }
}
#endif

// tensor cross terms on a box that touches non-periodic domain faces (mltensor_cross_terms_f? with bct / bv?lo / bv?hi): ON such
// a face the transverse derivative of component c comes from the boundary data -- Dirichlet: centred difference of the face
// values (ghost cells of the level-BC fab bv; zero when homogeneous), Neumann: centred difference of the interior cell row,
// reflect_odd: zero -- instead of the two-sided mean.
struct CrossBC { int lo[3][3], hi[3][3]; int dlo[3], dhi[3]; int per[3]; };
template <int D, int T>
IX_D double d_on_face(const C4& vel, const C4& bv, const CrossBC& cb, int c, int i, int j, int k, double dti) {
  const int q[3] = {i, j, k};
  auto V = [&](const C4& A, int oD, int oT) {
    return A(i + oD * (D == 0) + oT * (T == 0), j + oD * (D == 1) + oT * (T == 1), k + oD * (D == 2) + oT * (T == 2), c);
  };
  if (!cb.per[D] && (q[D] == cb.dlo[D] || q[D] == cb.dhi[D] + 1)) {
    const bool low = q[D] == cb.dlo[D];
    const int code = low ? cb.lo[c][D] : cb.hi[c][D];
    const int og = low ? -1 : 0, oi = low ? 0 : -1;
    if (code == IAMRX_LINOP_DIRICHLET) return bv.ok() ? (V(bv, og, 1) - V(bv, og, -1)) * (0.5 * dti) : 0.0;
    if (code == IAMRX_LINOP_NEUMANN) return (V(vel, oi, 1) - V(vel, oi, -1)) * (0.5 * dti);
    return 0.0;
  }
  return (V(vel, 0, 1) + V(vel, -1, 1) - V(vel, 0, -1) - V(vel, -1, -1)) * (0.25 * dti);
}
// cross flux of the D-face whose upper cell is (i,j,k): f[D] = (2/3) eta (d u_t1/d t1 + d u_t2/d t2), f[t] = -eta d u_D/d t
template <int D>
IX_D void cross_flux_bc(const C4& vel, const C4& bv, const C4& eta, const CrossBC& cb, int i, int j, int k, const double dxi[3], double f[3]) {
  constexpr int T1 = (D + 1) % 3, T2 = (D + 2) % 3;
  const double mu = eta(i, j, k);
  f[D] = -mu * (-(2.0 / 3.0) * (d_on_face<D, T1>(vel, bv, cb, T1, i, j, k, dxi[T1]) + d_on_face<D, T2>(vel, bv, cb, T2, i, j, k, dxi[T2])));
  f[T1] = -mu * d_on_face<D, T1>(vel, bv, cb, D, i, j, k, dxi[T1]);
  f[T2] = -mu * d_on_face<D, T2>(vel, bv, cb, D, i, j, k, dxi[T2]);
}
__global__ void __launch_bounds__(AP_TX* AP_TY)
tensor_cross_bc_kernel(Bx bx, V4 out, C4 vel, C4 bv, C4 ex, C4 ey, C4 ez, double b, double dxi0, double dxi1, double dxi2, IX_KARG(CrossBC) cb) {
  const int k = bx.lo[2] + blockIdx.z;
  const int j = bx.lo[1] + blockIdx.y * AP_TY + threadIdx.y;
  const int i = bx.lo[0] + blockIdx.x * AP_TX + threadIdx.x;
  if (j > bx.hi[1] || i > bx.hi[0]) return;
  const double dxi[3] = {dxi0, dxi1, dxi2};
  double fl[3], fh[3], acc[3];
  cross_flux_bc<0>(vel, bv, ex, cb, i, j, k, dxi, fl);
  cross_flux_bc<0>(vel, bv, ex, cb, i + 1, j, k, dxi, fh);
  for (int n = 0; n < 3; ++n) acc[n] = dxi0 * (fh[n] - fl[n]);
  cross_flux_bc<1>(vel, bv, ey, cb, i, j, k, dxi, fl);
  cross_flux_bc<1>(vel, bv, ey, cb, i, j + 1, k, dxi, fh);
  for (int n = 0; n < 3; ++n) acc[n] += dxi1 * (fh[n] - fl[n]);
  cross_flux_bc<2>(vel, bv, ez, cb, i, j, k, dxi, fl);
  cross_flux_bc<2>(vel, bv, ez, cb, i, j, k + 1, dxi, fh);
  for (int n = 0; n < 3; ++n) acc[n] += dxi2 * (fh[n] - fl[n]);
  for (int n = 0; n < 3; ++n) out(i, j, k, n) += b * acc[n];
}

#if !defined(IX_EMUL)
// ---- fused red+black GSRB sweep (box spans the periodic domain) ------------------------------
// One launch = one full sweep (colour `rb0` then the other), phi_in -> phi_out.  A CTA owns an xy tile
// (<= 60 x 14 cells) and marches along z through a chunk of planes.  Per plane r it
//   (1) red-updates plane r on the tile grown by one cell (the ring is recomputed redundantly by the
//       neighbouring CTAs with identical arithmetic) from the OLD values of planes r-1, r, r+1 held in a
//       4-slot shared-memory ring that cp.async fills one plane ahead, and stores the mixed plane
//       (new red, old black) in a 3-slot ring;
//   (2) black-updates plane r-1 on the tile from the mixed planes r-2, r-1, r and writes it out.
// Every phi / rhs / coefficient element is read from HBM once per sweep (plus the tile halo, which
// mostly hits L2) and phi is written once: 56 B/cell (48 + out) instead of 96 for two colour launches.
// A thread owns the cell pair (2p, 2p+1) of one row for the whole march: it loads the coefficients of
// both cells when the pair's plane is red-updated, uses the red cell's set at once and keeps the black
// cell's set in registers for the black update one iteration later.  The update expression is the one
// of gsrb_kernel, so the result is bit-identical to two colour passes.
namespace sweep {
constexpr int PX = 32;              // pairs per row: one warp
constexpr int PW = 2 * PX;          // plane width in shared memory: tile + 2 cells on either side
constexpr int TWMAX = PW - 4;       // 60
constexpr int NP = 4, NM = 3;       // ring depths: old planes / mixed planes
// tile rows TY are a template parameter: 14 (16 warps, one CTA per SM) or 6 (8 warps, two CTAs per SM: two
// independent barrier domains overlap each other's waits at the price of more halo rows)
template <int TY> struct Cfg {
  static constexpr int PH = TY + 4;          // plane rows in shared memory
  static constexpr int NW = TY + 2;          // warps: one per row of the red region
  static constexpr int NT = 32 * NW;
  static constexpr int PLANE = PW * PH;
  static constexpr int SMEM_BYTES = (NP + NM) * PLANE * (int)sizeof(double);
  static constexpr int MINB = (TY <= 6) ? 2 : 1;
};

IX_D int wrapi(int g, int lo, int n) {  // periodic image in [lo, lo+n)
  int m = (g - lo) % n;
  if (m < 0) m += n;
  return lo + m;
}
IX_D void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
// shared-memory slot of plane column c: (c & 1) * PX + (c >> 1) -- even columns first, odd columns after them, so
// that the stride-2 accesses of one colour (and of its x neighbours) are contiguous across a warp

template <bool HASA>
IX_D double relax(const AbecDev& op, double omega, double bxm, double bxp, double bym, double byp, double bzm, double bzp,
                  double rh, double ac, double p0, double pxm, double pxp, double pym, double pyp, double pzm, double pzp) {
  double gamma = op.dhx * (bxm + bxp) + op.dhy * (bym + byp) + op.dhz * (bzm + bzp);
  if (HASA) gamma += op.a * ac;
  const double rho = op.dhx * (bxm * pxm + bxp * pxp) + op.dhy * (bym * pym + byp * pyp) + op.dhz * (bzm * pzm + bzp * pzp);
  const double res = rh - (gamma * p0 - rho);
  return p0 + omega / gamma * res;
}

// raw coefficients of a cell pair on one plane: x faces 0,1,2; y faces (cell 0/1, lower/upper); z likewise; rhs; acoef
struct Raw { double bx0, bx1, bx2, by00, by10, by01, by11, bz00, bz10, bz01, bz11, rh0, rh1, ac0, ac1; };
// the coefficient set of cell F (0/1) of the pair
#define IX_SET(w, F) ((F) ? (w).bx1 : (w).bx0), ((F) ? (w).bx2 : (w).bx1), ((F) ? (w).by10 : (w).by00), ((F) ? (w).by11 : (w).by01), \
                     ((F) ? (w).bz10 : (w).bz00), ((F) ? (w).bz11 : (w).bz01), ((F) ? (w).rh1 : (w).rh0), ((F) ? (w).ac1 : (w).ac0)

struct Ctx {      // per-thread constants of the march
  double* P; double* M;
  int c0, pr, twl;
  bool pair_red, pair_blk;
};

// One plane of the march.  RF = which cell of the pair (0/1) belongs to the FIRST colour on plane r (compile-time:
// it alternates from plane to plane and is warp-uniform, so the caller dispatches once and unrolls by two).
//   cur : coefficients of plane r (loaded one plane ahead)
//   kp  : coefficients of the second-colour cell of plane r-1 (cell RF of the pair as well: the colours swap between planes)
struct Coef { double bxm, bxp, bym, byp, bzm, bzp, rh, ac; };
template <bool HASA, int RF, int PLANE>
IX_D void plane_step(const Ctx& t, const AbecDev& op, double omega, int rr, bool do_black, const Raw& cur, const Coef& kp, double* po) {
  const double* Pm = t.P + ((rr - 1) & (NP - 1)) * PLANE;
  const double* Pc = t.P + (rr & (NP - 1)) * PLANE;
  const double* Pp = t.P + ((rr + 1) & (NP - 1)) * PLANE;
  double* Mc = t.M + (rr % NM) * PLANE;
  const int row = t.pr * PW;
  // slots of the pair's two cells and of the x neighbours of cell RF
  const int s_r = row + RF * PX + (t.c0 >> 1), s_b = row + (1 - RF) * PX + (t.c0 >> 1);
  const int s_xm = RF ? s_b : s_b - 1, s_xp = RF ? s_b + 1 : s_b;   // columns cr-1 / cr+1 live in the other half
  if (t.pair_red) {
    // columns 0 and twl+3 are outside the red region: they are updated from whatever sits next to them, and nothing
    // reads the result (the second colour only looks at columns 1 .. twl+2)
    Mc[s_r] = relax<HASA>(op, omega, IX_SET(cur, RF), Pc[s_r], Pc[s_xm], Pc[s_xp], Pc[s_r - PW], Pc[s_r + PW], Pm[s_r], Pp[s_r]);
    Mc[s_b] = Pc[s_b];
  }
  __syncthreads();
  if (do_black && t.pair_blk) {  // second colour pass on plane r-1: its second-colour cell is cell RF of the pair
    const double* Mm = t.M + ((rr - 2) % NM) * PLANE;
    const double* Mk = t.M + ((rr - 1) % NM) * PLANE;
    const double v = relax<HASA>(op, omega, kp.bxm, kp.bxp, kp.bym, kp.byp, kp.bzm, kp.bzp, kp.rh, kp.ac,
                                 Mk[s_r], Mk[s_xm], Mk[s_xp], Mk[s_r - PW], Mk[s_r + PW], Mm[s_r], Mc[s_r]);
    po[RF] = v;
    po[1 - RF] = Mk[s_b];
  }
}

template <bool HASA, int TY>
__global__ void __launch_bounds__(Cfg<TY>::NT, Cfg<TY>::MINB)
gsrb_sweep_kernel(Bx bx, V4 out, C4 pin, C4 rhs, IX_KARG(AbecDev) op, double omega, int rb0, int tw, int th, int nzc, int nchunk) {
  constexpr int PLANE = Cfg<TY>::PLANE;
  extern __shared__ double sm[];
  Ctx t;
  t.P = sm;                  // [NP][PH][PW] old values
  t.M = sm + NP * PLANE;     // [NM][PH][PW] new first colour / old second colour
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = (int)blockIdx.z / nchunk, chunk = (int)blockIdx.z % nchunk;
  const int nb = (op.bncomp > 1) ? n : 0;
  const int nx = bx.nx(), ny = bx.ny(), nz = bx.nz();
  const int x0 = bx.lo[0] + tw * (int)blockIdx.x, y0 = bx.lo[1] + th * (int)blockIdx.y;
  const int twl = min(tw, bx.hi[0] + 1 - x0), thl = min(th, bx.hi[1] + 1 - y0);   // this tile's extent
  const int kc0 = bx.lo[2] + nzc * chunk, nzl = min(nzc, bx.hi[2] + 1 - kc0);
  // plane coordinates: column c <-> x = x0 - 2 + c (c = 0 .. twl+3), row pr <-> y = y0 - 2 + pr (pr = 0 .. thl+3)
  const int c0 = 2 * lane, pr = warp + 1;
  t.c0 = c0; t.pr = pr; t.twl = twl;
  const bool col_ok = c0 <= twl + 3;          // pair inside the staged plane (twl even: both cells or none)
  const int gx = wrapi(x0 - 2 + c0, bx.lo[0], nx);   // pairs never straddle the wrap (x0 - lo, nx even)
  const int gy = wrapi(y0 - 2 + pr, bx.lo[1], ny);
  // staging: each thread copies its own pair of row pr; warps 0 / 1 also copy rows 0 / thl+3
  const int pr2 = (warp == 0) ? 0 : thl + 3;
  const bool extra = col_ok && warp < 2;
  const int gy2 = wrapi(y0 - 2 + pr2, bx.lo[1], ny);
  const bool own = col_ok && pr <= thl + 3;
  const int r0 = kc0 - 1, nred = nzl + 2;
  // All plane addressing is incremental: a 32-bit element offset per array that advances by the array's plane
  // stride and wraps with the periodic plane index (the host checks that every array has < 2^31 elements).
  const double* pst1 = pin.p + n * pin.ns + (gx - pin.l0) + (gy - pin.l1) * (int)pin.js;
  const double* pst2 = pin.p + n * pin.ns + (gx - pin.l0) + (gy2 - pin.l1) * (int)pin.js;
  const int d1 = pr * PW + (c0 >> 1), d2 = pr2 * PW + (c0 >> 1);
  const int pks = (int)pin.ks;
  int gk_st = wrapi(r0 - 1, bx.lo[2], nz);   // plane being STAGED (two ahead of the plane being updated)
  int ko_st = (gk_st - pin.l2) * pks;
  int st_cnt = 3;                            // its ring counter (plane r0-1 <-> 3)
  auto stage = [&]() {
    double* dst = t.P + (st_cnt & (NP - 1)) * PLANE;
    if (own) { cp_async8(dst + d1, pst1 + ko_st); cp_async8(dst + d1 + PX, pst1 + ko_st + 1); }
    if (extra) { cp_async8(dst + d2, pst2 + ko_st); cp_async8(dst + d2 + PX, pst2 + ko_st + 1); }
    asm volatile("cp.async.commit_group;" ::: "memory");
    ++st_cnt;
    const bool w = gk_st == bx.hi[2];
    gk_st = w ? bx.lo[2] : gk_st + 1;
    ko_st = w ? ko_st - (nz - 1) * pks : ko_st + pks;
    if (own) asm volatile("prefetch.global.L2 [%0];" ::"l"(pst1 + ko_st));   // next plane to be staged -> L2
  };
  stage(); stage(); stage();
  t.pair_red = col_ok && pr >= 1 && pr <= thl + 2 && c0 <= twl + 2;      // at least one cell of the pair in the red region
  t.pair_blk = pr >= 2 && pr <= thl + 1 && c0 >= 2 && c0 + 1 <= twl + 1;   // pair inside the tile
  const double* bxp_ = op.bx.p + nb * op.bx.ns + (gx - op.bx.l0) + (gy - op.bx.l1) * (int)op.bx.js;
  const double* byp_ = op.by.p + nb * op.by.ns + (gx - op.by.l0) + (gy - op.by.l1) * (int)op.by.js;
  const double* bzp_ = op.bz.p + nb * op.bz.ns + (gx - op.bz.l0) + (gy - op.bz.l1) * (int)op.bz.js;
  const double* rhp_ = rhs.p + n * rhs.ns + (gx - rhs.l0) + (gy - rhs.l1) * (int)rhs.js;
  const double* acp_ = HASA ? op.acoef.p + (gx - op.acoef.l0) + (gy - op.acoef.l1) * (int)op.acoef.js : nullptr;
  double* outp_ = out.p + n * out.ns + (gx - out.l0) + (gy - out.l1) * (int)out.js;
  const int byjs = (int)op.by.js;
  const int bxks = (int)op.bx.ks, byks = (int)op.by.ks, bzks = (int)op.bz.ks, rhks = (int)rhs.ks, acks = HASA ? (int)op.acoef.ks : 0;
  // coefficient loads run ONE PLANE AHEAD of the update (registers) and are preceded by an L2 prefetch TWO planes
  // ahead, so that their HBM latency overlaps the updates; gk_ld = plane being loaded
  int gk_ld = wrapi(r0, bx.lo[2], nz);
  int kbx = (gk_ld - op.bx.l2) * bxks, kby = (gk_ld - op.by.l2) * byks, kbz = (gk_ld - op.bz.l2) * bzks, krh = (gk_ld - rhs.l2) * rhks;
  int kac = HASA ? (gk_ld - op.acoef.l2) * acks : 0;
  auto load_raw = [&](Raw& w, const Raw* below) {
    const double* pbx = bxp_ + kbx;
    const double* pby = byp_ + kby;
    const double* pbz = bzp_ + kbz;
    const double* prh = rhp_ + krh;
    w.bx0 = pbx[0]; w.bx1 = pbx[1]; w.bx2 = pbx[2];
    w.by00 = pby[0]; w.by10 = pby[1]; w.by01 = pby[byjs]; w.by11 = pby[byjs + 1];
    if (below) { w.bz00 = below->bz01; w.bz10 = below->bz11; }   // the lower z faces are the upper ones of the plane below
    else { w.bz00 = pbz[0]; w.bz10 = pbz[1]; }
    w.bz01 = pbz[bzks]; w.bz11 = pbz[bzks + 1];
    w.rh0 = prh[0]; w.rh1 = prh[1];
    w.ac0 = 0.0; w.ac1 = 0.0;
    if (HASA) { const double* pac = acp_ + kac; w.ac0 = pac[0]; w.ac1 = pac[1]; }
    const bool wr = gk_ld == bx.hi[2];
    gk_ld = wr ? bx.lo[2] : gk_ld + 1;
    kbx = wr ? kbx - (nz - 1) * bxks : kbx + bxks;
    kby = wr ? kby - (nz - 1) * byks : kby + byks;
    kbz = wr ? kbz - (nz - 1) * bzks : kbz + bzks;
    krh = wr ? krh - (nz - 1) * rhks : krh + rhks;
    if (HASA) kac = wr ? kac - (nz - 1) * acks : kac + acks;
    // the next plane's lines -> L2 (one 16-byte touch per pair covers the warp's 512-byte row segment)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(bxp_ + kbx));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(byp_ + kby));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(bzp_ + kbz + bzks));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(rhp_ + krh));
    if (HASA) asm volatile("prefetch.global.L2 [%0];" ::"l"(acp_ + kac));
  };
  Raw A{}, B{};   // A: coefficients of even iterations' planes, B: of odd iterations' planes
  if (t.pair_red) load_raw(A, nullptr);
  // which cell of the pair is updated first on plane r0 (warp-uniform: 2*lane drops out of the parity)
  const int f0 = (x0 - 2 + c0 + y0 - 2 + pr + rb0 + r0) & 1;
  const int oks = (int)out.ks;
  int gk_out = wrapi(r0 - 1, bx.lo[2], nz);   // plane r-1 of iteration kk (written from kk >= 2)
  int ko_out = (gk_out - out.l2) * oks;
  auto step = [&](auto rf_tag, int kk, const Raw& cur, Raw& nxt) {
    constexpr int RF = decltype(rf_tag)::value;
    // `nxt` still holds plane r-1: save the set its second-colour update needs before the prefetch of plane r+1 reuses it
    const Coef kp = {IX_SET(nxt, RF)};
    if (t.pair_red && kk + 1 < nred) load_raw(nxt, &cur);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (kk + 1 < nred) stage();
    plane_step<HASA, RF, PLANE>(t, op, omega, kk + 4, kk >= 2, cur, kp, outp_ + ko_out);
    const bool w = gk_out == bx.hi[2];
    gk_out = w ? bx.lo[2] : gk_out + 1;
    ko_out = w ? ko_out - (nz - 1) * oks : ko_out + oks;
  };
  // iteration kk updates plane r0+kk with coefficients in A (kk even) / B (kk odd); the previous plane's set is the other one
  if (f0 == 0) {
    for (int kk = 0; kk < nred; kk += 2) {
      step(std::integral_constant<int, 0>{}, kk, A, B);
      if (kk + 1 < nred) step(std::integral_constant<int, 1>{}, kk + 1, B, A);
    }
  } else {
    for (int kk = 0; kk < nred; kk += 2) {
      step(std::integral_constant<int, 1>{}, kk, A, B);
      if (kk + 1 < nred) step(std::integral_constant<int, 0>{}, kk + 1, B, A);
    }
  }
}
#undef IX_SET

inline bool gsrb_sweep_ok(const Bx& bx, int wrapmask) {  // shape requirements of the fused sweep
  if (wrapmask != 7) return false;
  if ((int64_t)(bx.nx() + 18) * (bx.ny() + 2) * (bx.nz() + 2) * 3 >= ((int64_t)1 << 31)) return false;   // 32-bit plane offsets
  return bx.nx() % 2 == 0 && bx.ny() % 2 == 0 && bx.nz() % 2 == 0 && bx.nx() >= 8 && bx.ny() >= 8 && bx.nz() >= 8;
}
}  // namespace sweep
#endif

inline dim3 grid_for(const Bx& bx, int tx, int ty, int nz_total) {
  return dim3(cdiv(bx.nx(), tx), cdiv(bx.ny(), ty), nz_total);
}

}  // namespace

int abec_gsrb(const Bx& bx, V4 phi, C4 rhs, const Abec& op, double omega, int redblack, int ncomp,
              cudaStream_t s, int wrapmask, const GsBC* gb, bool zero_phi) {
  if (!bx.ok()) return IAMRX_OK;
  ProfScope prof_(IAMRX_PROF_ABEC_GSRB, bx.npts(),
                  (double)bx.npts() * ncomp * ((op.cc ? 24.0 : 48.0) + ((op.a != 0.0 && !(op.cc && op.cac)) ? 8.0 : 0.0) - (zero_phi ? 8.0 : 0.0)), s);
  dim3 blk(GS_TX, GS_TY, 1);
  dim3 grd(cdiv(bx.nx() + 1, 2 * GS_TX), cdiv(bx.ny(), GS_TY), bx.nz() * ncomp);
  static int minb = -1;
  if (minb < 0) { const char* e = getenv("IAMRX_GSRB_MINB"); minb = e ? atoi(e) : 6; }
  const GsBC none{};
#define IX_GSRB(M, B, C, Z, G) IX_LAUNCH((gsrb_kernel<M, B, C, Z>), grd, blk, 0, s, bx, phi, rhs, to_dev(op), omega, redblack, bx.nz(), wrapmask, G)
  if (zero_phi) {
    if (op.cc) { if (gb) IX_GSRB(6, true, true, true, *gb); else IX_GSRB(6, false, true, true, none); }
    else { if (gb) IX_GSRB(6, true, false, true, *gb); else IX_GSRB(6, false, false, true, none); }
  }
  else if (op.cc) { if (gb) IX_GSRB(6, true, true, false, *gb); else IX_GSRB(6, false, true, false, none); }
  else if (gb) IX_GSRB(6, true, false, false, *gb);
  else if (minb >= 8) IX_GSRB(8, false, false, false, none);
  else IX_GSRB(6, false, false, false, none);
#undef IX_GSRB
  return check_launch("abec_gsrb");
}

bool abec_gsrb_small_ok(const Bx& bx, int ncomp) { return bx.npts() * ncomp <= 16384; }
int abec_gsrb_small(const Bx& bx, V4 phi, C4 rhs, const Abec& op, double omega, int ncomp, int nsweeps, bool zero_phi, cudaStream_t s,
                    int wrapmask, const GsBC* gb) {
  if (!bx.ok() || nsweeps <= 0) return IAMRX_OK;
#if defined(IX_EMUL)
  // host emulation (tests only; no barriers there): the same sweeps as colour launches
  for (int sw = 0; sw < nsweeps; ++sw)
    for (int rb = 0; rb < 2; ++rb) {
      const int rc = abec_gsrb(bx, phi, rhs, op, omega, rb, ncomp, s, wrapmask, gb, zero_phi && sw == 0 && rb == 0);
      if (rc != IAMRX_OK) return rc;
    }
  return IAMRX_OK;
#else
  const GsBC none{};
#define IX_GSS(B, C, G) IX_LAUNCH((gsrb_small_kernel<B, C>), 1, GS_SMALL_T, 0, s, bx, phi, rhs, to_dev(op), omega, ncomp, wrapmask, G, nsweeps, zero_phi ? 1 : 0)
  if (op.cc) { if (gb) IX_GSS(true, true, *gb); else IX_GSS(false, true, none); }
  else { if (gb) IX_GSS(true, false, *gb); else IX_GSS(false, false, none); }
#undef IX_GSS
  return check_launch("abec_gsrb_small");
#endif
}

// Measured on B200 (profiles/r01_notes.md): the fused sweep halves the DRAM traffic (832 MB vs 1590 MB per sweep at 256^3)
// but its per-plane barriers leave it latency-bound at one CTA per SM -- 252 us vs 264 us for the two colour launches at
// 256^3 and slower on the L2-resident coarser levels -- so the multigrid smoother uses it only on request
// (IAMRX_GSRB_FUSED=1) until the coefficient planes are staged asynchronously as well.
bool abec_gsrb_sweep_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("IAMRX_GSRB_FUSED"); on = (e && e[0] == '1') ? 1 : 0; }
  return on != 0;
}

bool abec_gsrb_sweep_ok(const Bx& bx, int wrapmask) {
#if defined(IX_EMUL)
  (void)bx;
  return wrapmask == 7;
#else
  // IAMRX_GSRB_FUSED_MIN=<cells>: only boxes at least this large take the fused sweep (it wins on HBM-resident levels and
  // loses on the L2-resident coarser ones)
  static int64_t min_cells = -1;
  if (min_cells < 0) { const char* e = getenv("IAMRX_GSRB_FUSED_MIN"); min_cells = e ? atoll(e) : 0; }
  if (bx.npts() < min_cells) return false;
  return sweep::gsrb_sweep_ok(bx, wrapmask);
#endif
}

int abec_gsrb_sweep(const Bx& bx, V4 phi_out, C4 phi_in, C4 rhs, const Abec& op, double omega, int rb0, int ncomp, cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
#if defined(IX_EMUL)
  // host emulation (tests only): the same sweep as two in-place colour passes on a copy
  int rc = copy(bx, phi_out, phi_in, ncomp, s);
  for (int rb = 0; rb < 2 && rc == IAMRX_OK; ++rb) rc = abec_gsrb(bx, phi_out, rhs, op, omega, rb0 ^ rb, ncomp, s, 7);
  return rc;
#else
  using namespace sweep;
  // algorithmic bytes as for two colour passes (SURVEY.md 8d counts no temporal blocking); the kernel moves about 56 B/cell
  ProfScope prof_(IAMRX_PROF_ABEC_GSRB, bx.npts(), (double)bx.npts() * ncomp * 2.0 * (op.a != 0.0 ? 56.0 : 48.0), s);
  static int ty = -1;   // tile rows: 14 (default) or 6 (IAMRX_GSRB_FUSED_TY)
  if (ty < 0) { const char* e = getenv("IAMRX_GSRB_FUSED_TY"); ty = (e && atoi(e) == 6) ? 6 : 14; }
  static bool attr_set = false;
  if (!attr_set) {
    IX_CUDA(cudaFuncSetAttribute(gsrb_sweep_kernel<false, 14>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<14>::SMEM_BYTES));
    IX_CUDA(cudaFuncSetAttribute(gsrb_sweep_kernel<true, 14>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<14>::SMEM_BYTES));
    IX_CUDA(cudaFuncSetAttribute(gsrb_sweep_kernel<false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<6>::SMEM_BYTES));
    IX_CUDA(cudaFuncSetAttribute(gsrb_sweep_kernel<true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<6>::SMEM_BYTES));
    attr_set = true;
  }
  // even tile widths/heights that split the box evenly
  const int ntx = cdiv(bx.nx(), TWMAX), nty = cdiv(bx.ny(), ty);
  const int tw = 2 * cdiv(cdiv(bx.nx(), ntx), 2), th = 2 * cdiv(cdiv(bx.ny(), nty), 2);
  const int gx = cdiv(bx.nx(), tw), gy = cdiv(bx.ny(), th);
  // z chunks: fill whole waves of resident CTAs (each chunk recomputes 2 extra red planes; >= 8 planes per chunk)
  int nchunk = 1;
  {
    const int nxy = gx * gy * ncomp, nslots = 148 * (ty == 6 ? 2 : 1);
    double best = 1e30;
    for (int c = 1; c <= bx.nz() / 8; ++c) {
      const int nzc_ = cdiv(bx.nz(), c), cc = cdiv(bx.nz(), nzc_);
      const double cost = (double)cdiv(nxy * cc, nslots) * (nzc_ + 2);   // waves x iterations per CTA
      if (cost < best - 1e-9) { best = cost; nchunk = cc; }
    }
  }
  const int nzc = cdiv(bx.nz(), nchunk);
  nchunk = cdiv(bx.nz(), nzc);
  const dim3 grd(gx, gy, nchunk * ncomp);
#define IX_SWEEP(HA, T) IX_LAUNCH((gsrb_sweep_kernel<HA, T>), grd, dim3(Cfg<T>::NT, 1, 1), Cfg<T>::SMEM_BYTES, s, bx, phi_out, phi_in, rhs, \
                                  to_dev(op), omega, rb0, tw, th, nzc, nchunk)
  if (ty == 6) { if (op.a != 0.0) IX_SWEEP(true, 6); else IX_SWEEP(false, 6); }
  else { if (op.a != 0.0) IX_SWEEP(true, 14); else IX_SWEEP(false, 14); }
#undef IX_SWEEP
  return check_launch("abec_gsrb_sweep");
#endif
}

int abec_apply(const Bx& bx, V4 out, C4 phi, C4 rhs, const Abec& op, int ncomp, cudaStream_t s, int wrapmask, const GsBC* gb,
               double* norm_dev, bool* norm_fused) {
  if (norm_fused) *norm_fused = false;
  if (!bx.ok()) return IAMRX_OK;
  MirBC mb{};
  bool mirrored = false;
  if (gb) for (int c = 0; c < 3; ++c) { mb.even[c] = gb->even[c]; mb.odd[c] = gb->odd[c]; if (mb.even[c] | mb.odd[c]) mirrored = true; }
  ProfScope prof_(IAMRX_PROF_ABEC_APPLY, bx.npts(), (double)bx.npts() * ncomp * ((op.cc ? 24.0 : 48.0) + ((op.a != 0.0 && !(op.cc && op.cac)) ? 8.0 : 0.0) + (rhs.ok() ? 0.0 : -8.0)), s);
#if !defined(IX_EMUL)
  if (!mirrored && bx.nx() % 2 == 0 && pairs_aligned(out, bx) && pairs_aligned(phi, bx) && pairs_aligned(rhs, bx) && pairs_aligned(op.acoef, bx) &&
      pairs_aligned(op.bx, bx) && pairs_aligned(op.by, bx) && pairs_aligned(op.bz, bx)) {
    const dim3 grd(cdiv(bx.nx() / 2, AP_TX), cdiv(bx.ny(), AP_TY), bx.nz() * ncomp);
    const bool nf = norm_dev != nullptr && (bx.nx() / 2) % 32 == 0;
    if (norm_fused) *norm_fused = nf;
#define IX_AP2(A, C, N) IX_LAUNCH((apply2_kernel<A, C, N>), grd, dim3(AP_TX, AP_TY, 1), 0, s, bx, out, phi, rhs, to_dev(op), bx.nz(), wrapmask, norm_dev)
    if (nf) {
      if (op.cc) { if (op.a != 0.0) IX_AP2(true, true, true); else IX_AP2(false, true, true); }
      else { if (op.a != 0.0) IX_AP2(true, false, true); else IX_AP2(false, false, true); }
    } else {
      if (op.cc) { if (op.a != 0.0) IX_AP2(true, true, false); else IX_AP2(false, true, false); }
      else { if (op.a != 0.0) IX_AP2(true, false, false); else IX_AP2(false, false, false); }
    }
#undef IX_AP2
    return check_launch("abec_apply2");
  }
#endif
  if (op.cc) IX_LAUNCH(apply_kernel<true>, grid_for(bx, AP_TX, AP_TY, bx.nz() * ncomp), dim3(AP_TX, AP_TY, 1), 0, s,
      bx, out, phi, rhs, to_dev(op), bx.nz(), wrapmask, mb);
  else IX_LAUNCH(apply_kernel<false>, grid_for(bx, AP_TX, AP_TY, bx.nz() * ncomp), dim3(AP_TX, AP_TY, 1), 0, s,
      bx, out, phi, rhs, to_dev(op), bx.nz(), wrapmask, mb);
  return check_launch("abec_apply");
}

int abec_flux(const Bx& bx, V4 fx, V4 fy, V4 fz, C4 phi, const Abec& op, int comp, cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  Bx g = bx; g.hi[0]++; g.hi[1]++; g.hi[2]++;
  IX_LAUNCH(flux_kernel, grid_for(g, AP_TX, AP_TY, g.nz()), dim3(AP_TX, AP_TY, 1), 0, s, 
      bx, fx, fy, fz, phi, to_dev(op), op.b * op.dxinv[0], op.b * op.dxinv[1], op.b * op.dxinv[2], comp);
  return check_launch("abec_flux");
}

int cc_restrict(const Bx& cbx, V4 crse, C4 fine, int ncomp, cudaStream_t s, int thin) {
  if (!cbx.ok()) return IAMRX_OK;
  IX_LAUNCH(restrict_kernel, grid_for(cbx, AP_TX, AP_TY, cbx.nz() * ncomp), dim3(AP_TX, AP_TY, 1), 0, s, 
      cbx, crse, fine, cbx.nz(), thin);
  return check_launch("cc_restrict");
}

int cc_prolong_add(const Bx& fbx, V4 fine, C4 crse, int ncomp, cudaStream_t s, int thin) {
  if (!fbx.ok()) return IAMRX_OK;
  const int nzh = cdiv(fbx.nz(), 2);
  IX_LAUNCH(prolong_kernel, grid_for(fbx, AP_TX, AP_TY, nzh * ncomp), dim3(AP_TX, AP_TY, 1), 0, s,
      fbx, fine, crse, nzh, thin);
  return check_launch("cc_prolong_add");
}

int face_restrict(const Bx& cfbx, int dir, V4 crse, C4 fine, int ncomp, cudaStream_t s, int thin) {
  if (!cfbx.ok()) return IAMRX_OK;
  IX_LAUNCH(face_restrict_kernel, grid_for(cfbx, AP_TX, AP_TY, cfbx.nz() * ncomp), dim3(AP_TX, AP_TY, 1), 0, s, 
      cfbx, dir, crse, fine, cfbx.nz(), thin);
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

int tensor_cross_bc(const Bx& bx, V4 out, C4 vel, C4 bv, C4 ex, C4 ey, C4 ez, double b, const double dxinv[3], const LinBC& bc,
                    const Bx& dom, const int per[3], cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  bool touches = false;
  for (int d = 0; d < 3; ++d) if (!per[d] && (bx.lo[d] == dom.lo[d] || bx.hi[d] == dom.hi[d])) touches = true;
  if (!touches) return tensor_cross(bx, out, vel, ex, ey, ez, b, dxinv, s);
  CrossBC cb;
  for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) { cb.lo[c][d] = bc.lo[c][d]; cb.hi[c][d] = bc.hi[c][d]; }
  for (int d = 0; d < 3; ++d) { cb.dlo[d] = dom.lo[d]; cb.dhi[d] = dom.hi[d]; cb.per[d] = per[d]; }
  IX_LAUNCH(tensor_cross_bc_kernel, grid_for(bx, AP_TX, AP_TY, bx.nz()), dim3(AP_TX, AP_TY, 1), 0, s,
      bx, out, vel, bv, ex, ey, ez, b, dxinv[0], dxinv[1], dxinv[2], cb);
  return check_launch("tensor_cross_bc");
}

int tensor_cross(const Bx& bx, V4 out, C4 vel, C4 ex, C4 ey, C4 ez, double b, const double dxinv[3],
                 cudaStream_t s, int wrapmask) {
  if (!bx.ok()) return IAMRX_OK;
#if !defined(IX_EMUL)
  {
    static int on = -1;
    // measured on B200 (profiles/r02_notes.md): 1.22 ms vs 0.86 ms per launch at 256^3 -- the per-plane barriers cost more than the
    // halved stencil loads save -- so the marching form is opt-in (IAMRX_TENSOR_MARCH=1)
    if (on < 0) { const char* e = getenv("IAMRX_TENSOR_MARCH"); on = (e && e[0] == '1') ? 1 : 0; }
    if (on && bx.nz() >= 8 && wrapmask == 0) {
      const dim3 grd(cdiv(bx.nx(), TC_X), cdiv(bx.ny(), TC_Y), cdiv(bx.nz(), TC_KB));
      IX_LAUNCH(tensor_cross_march_kernel, grd, dim3(TC_X + 1, TC_Y + 1, 1), 0, s, bx, out, vel, ex, ey, ez, b, dxinv[0], dxinv[1], dxinv[2]);
      return check_launch("tensor_cross_march");
    }
  }
#endif
  static int minb = -1;
  if (minb < 0) { const char* e = getenv("IAMRX_TC_MINB"); minb = e ? atoi(e) : 2; }   // measured per call at 256^3: 2 -> 0.62 ms, 3 -> 0.70, 4 -> 0.69 (uncapped: 0.85)
#define IX_TC(M) IX_LAUNCH(tensor_cross_kernel<M>, grid_for(bx, AP_TX, AP_TY, bx.nz()), dim3(AP_TX, AP_TY, 1), 0, s, bx, out, vel, ex, ey, ez, b, dxinv[0], dxinv[1], dxinv[2], wrapmask)
  if (minb >= 4) IX_TC(4); else if (minb == 3) IX_TC(3); else IX_TC(2);
#undef IX_TC
  return check_launch("tensor_cross");
}

}  // namespace k
}  // namespace ix
