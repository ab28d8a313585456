// godunov.cu -- Godunov (PLM) advection kernels for sm_100a.
//
//   extrap_vel_to_faces : Godunov::ExtrapVelToFaces        (NavierStokesBase.cpp:4487-4491)
//   compute_aofs        : NavierStokesBase::ComputeAofs body (NavierStokesBase.cpp:4661-4845) =
//                         HydroUtils::ComputeFluxesOnBoxFromState -> ComputeDivergence(mult=-1)
//                         -> ComputeConvectiveTerm -> aofs = -update
//
// Pipeline (one launch per stage over the whole box, all components in the grid):
//   edge   : PLM trace to every face of grow(bx,1), upwinded by the face velocity
//            -> 3 "edge" scratch arrays per component
//   corner : the six corner-coupled transverse states (3-D only)  -> 6 scratch arrays
//   final  : transverse + forcing terms, final upwind -> edge state and area-weighted flux
//   diverg : -div(flux)/vol + qbar*div(umac) for non-conservative comps, sign -> aofs
// The lo/hi traced states are recomputed from q where a later stage needs them (two
// 4th-order slopes) rather than stored: 9 scratch arrays per component instead of 15.
// Physical-boundary (ext_dir/hoextrap) edge treatment is not implemented: the C ABI
// rejects non-periodic configurations before reaching these kernels.
#include "godunov_math.h"
#include "kernels.h"
#include "level.h"

namespace ix {
namespace k {
namespace {

using namespace gd;

constexpr int TX = 64;
constexpr int TY = 4;

struct Comp {  // one component of a C4 view as a 3-index callable
  const double* p; int l0, l1, l2; int64_t js, ks;
  IX_HD double operator()(int i, int j, int k) const { return p[(i - l0) + (j - l1) * js + (k - l2) * ks]; }
};
IX_HD Comp comp(const C4& v, int n) { return Comp{v.p + n * v.ns, v.l0, v.l1, v.l2, v.js, v.ks}; }

// traced states on face (i,j,k) of direction D: lo from the cell below, hi from the
// cell above.  ulo/uhi: the velocity used in the trace (cell-centred normal velocity
// for ExtrapVelToFaces, the MAC velocity of this face for ComputeEdgeState).
template <int D, class Q>
IX_D void trace(const Q& q, int i, int j, int k, double ulo, double uhi, double dtdx, double& lo, double& hi) {
  const int im = i - E<D>::x, jm = j - E<D>::y, km = k - E<D>::z;
  lo = q(im, jm, km) + 0.5 * (1.0 - ulo * dtdx) * slope4<D>(q, im, jm, km);
  hi = q(i, j, k) + 0.5 * (-1.0 - uhi * dtdx) * slope4<D>(q, i, j, k);
}

struct Scratch {  // all on the same grown index box, component-major
  double* p; int l0, l1, l2; int64_t js, ks, as;  // as = stride between arrays
  IX_HD double& operator()(int a, int i, int j, int k) const {
    return p[a * as + (i - l0) + (j - l1) * js + (k - l2) * ks];
  }
};

#define GIDX(R)                                                \
  const int nz_ = R.hi[2] - R.lo[2] + 1;                      \
  const int k = R.lo[2] + (int)(blockIdx.z % nz_);            \
  const int n = (int)(blockIdx.z / nz_);                      \
  const int j = R.lo[1] + blockIdx.y * TY + threadIdx.y;      \
  const int i = R.lo[0] + blockIdx.x * TX + threadIdx.x;      \
  if (j > R.hi[1] || i > R.hi[0]) return;

inline dim3 grid_for(const Bx& r, int ncomp) { return dim3(cdiv(r.nx(), TX), cdiv(r.ny(), TY), r.nz() * ncomp); }

// ===========================================================================
// ComputeEdgeState path
// ===========================================================================
struct EsArgs {
  Bx bx;
  C4 S, force, divu, umac, vmac, wmac, uflx, vflx, wflx;
  int iconserv[8];
  int fit;  // use_forces_in_trans
  double dt, dtdx, dtdy, dtdz;
};

// scratch array ids per component: 0..2 edge x,y,z; 3..8 corner xy,xz,yx,yz,zx,zy; 9..11 flux
enum { A_XE = 0, A_YE, A_ZE, A_XY, A_XZ, A_YX, A_YZ, A_ZX, A_ZY, A_FX, A_FY, A_FZ, A_N };

template <int D>
IX_D void es_lohi(const EsArgs& a, const Comp& q, const C4& mac, int n, int i, int j, int k, double dtdx,
                  double& lo, double& hi) {
  const double u = mac(i, j, k);
  trace<D>(q, i, j, k, u, u, dtdx, lo, hi);
  if (a.fit && a.force.ok()) {
    lo += 0.5 * a.dt * a.force(i - E<D>::x, j - E<D>::y, k - E<D>::z, n);
    hi += 0.5 * a.dt * a.force(i, j, k, n);
  }
}

__global__ void __launch_bounds__(TX* TY) es_edge_kernel(IX_KARG(EsArgs) a, IX_KARG(Scratch) sc, IX_KARG(Bx) R) {
  GIDX(R)
  const Comp q = comp(a.S, n);
  const Bx& b = a.bx;
  const bool inx = i >= b.lo[0] - 1 && i <= b.hi[0] + 1, iny = j >= b.lo[1] - 1 && j <= b.hi[1] + 1,
             inz = k >= b.lo[2] - 1 && k <= b.hi[2] + 1;
  double lo, hi;
  if (i >= b.lo[0] && i <= b.hi[0] + 1 && iny && inz) {  // x-faces lo..hi+1
    es_lohi<0>(a, q, a.umac, n, i, j, k, a.dtdx, lo, hi);
    sc(A_N * n + A_XE, i, j, k) = upwind(lo, hi, a.umac(i, j, k));
  }
  if (j >= b.lo[1] && j <= b.hi[1] + 1 && inx && inz) {
    es_lohi<1>(a, q, a.vmac, n, i, j, k, a.dtdy, lo, hi);
    sc(A_N * n + A_YE, i, j, k) = upwind(lo, hi, a.vmac(i, j, k));
  }
  if (k >= b.lo[2] && k <= b.hi[2] + 1 && inx && iny) {
    es_lohi<2>(a, q, a.wmac, n, i, j, k, a.dtdz, lo, hi);
    sc(A_N * n + A_ZE, i, j, k) = upwind(lo, hi, a.wmac(i, j, k));
  }
}

// corner coupling of the D1-face state at (i,j,k) with the derivative along D2
// (Godunov_corner_couple_<d1><d2>): mac2/edge2 live on D2-faces.
template <int D1, int D2, class Q, class M, class Ed>
IX_D void corner(double& lo1, double& hi1, double lo, double hi, const Q& q, const M& mac2, const Ed& edge2,
                 int i, int j, int k, double dt3dx, bool conserv) {
  const int im = i - E<D1>::x, jm = j - E<D1>::y, km = k - E<D1>::z;  // cell below the face
  const int ip = E<D2>::x, jp = E<D2>::y, kp = E<D2>::z;
  const double mlo_p = mac2(im + ip, jm + jp, km + kp), mlo_m = mac2(im, jm, km);
  const double mhi_p = mac2(i + ip, j + jp, k + kp), mhi_m = mac2(i, j, k);
  lo1 = lo - dt3dx * (edge2(im + ip, jm + jp, km + kp) * mlo_p - edge2(im, jm, km) * mlo_m);
  hi1 = hi - dt3dx * (edge2(i + ip, j + jp, k + kp) * mhi_p - edge2(i, j, k) * mhi_m);
  if (!conserv) {
    lo1 += dt3dx * q(im, jm, km) * (mlo_p - mlo_m);
    hi1 += dt3dx * q(i, j, k) * (mhi_p - mhi_m);
  }
}

struct ScArr {  // one scratch array as a 3-index callable
  const double* p; int l0, l1, l2; int64_t js, ks;
  IX_HD double operator()(int i, int j, int k) const { return p[(i - l0) + (j - l1) * js + (k - l2) * ks]; }
};
IX_HD ScArr arr(const Scratch& s, int a) { return ScArr{s.p + a * s.as, s.l0, s.l1, s.l2, s.js, s.ks}; }

__global__ void __launch_bounds__(TX* TY) es_corner_kernel(IX_KARG(EsArgs) a, IX_KARG(Scratch) sc, IX_KARG(Bx) R) {
  GIDX(R)
  const Comp q = comp(a.S, n);
  const Bx& b = a.bx;
  const bool cs = a.iconserv[n] != 0;
  const int o = A_N * n;
  const ScArr xe = arr(sc, o + A_XE), ye = arr(sc, o + A_YE), ze = arr(sc, o + A_ZE);
  const bool cx = i <= b.hi[0], cy = j <= b.hi[1], cz = k <= b.hi[2];           // cell index inside bx (upper)
  const bool gx = i >= b.lo[0] - 1 && i <= b.hi[0] + 1, gy = j >= b.lo[1] - 1 && j <= b.hi[1] + 1,
             gz = k >= b.lo[2] - 1 && k <= b.hi[2] + 1;
  const bool fx = i >= b.lo[0] && i <= b.hi[0] + 1, fy = j >= b.lo[1] && j <= b.hi[1] + 1,
             fz = k >= b.lo[2] && k <= b.hi[2] + 1;  // face index of bx
  double lo, hi, l1, h1;
  // x-faces
  if (fx && ((fy && cy && gz) || (gy && fz && cz))) {
    es_lohi<0>(a, q, a.umac, n, i, j, k, a.dtdx, lo, hi);
    if (fy && cy && gz) {  // xy: needed for z-faces -> grown in z
      corner<0, 1>(l1, h1, lo, hi, q, a.vmac, ye, i, j, k, a.dtdy / 3.0, cs);
      sc(o + A_XY, i, j, k) = upwind(l1, h1, a.umac(i, j, k));
    }
    if (gy && fz && cz) {  // xz: needed for y-faces -> grown in y
      corner<0, 2>(l1, h1, lo, hi, q, a.wmac, ze, i, j, k, a.dtdz / 3.0, cs);
      sc(o + A_XZ, i, j, k) = upwind(l1, h1, a.umac(i, j, k));
    }
  }
  // y-faces
  if (fy && ((fx && cx && gz) || (gx && fz && cz))) {
    es_lohi<1>(a, q, a.vmac, n, i, j, k, a.dtdy, lo, hi);
    if (fx && cx && gz) {  // yx: needed for z-faces
      corner<1, 0>(l1, h1, lo, hi, q, a.umac, xe, i, j, k, a.dtdx / 3.0, cs);
      sc(o + A_YX, i, j, k) = upwind(l1, h1, a.vmac(i, j, k));
    }
    if (gx && fz && cz) {  // yz: needed for x-faces
      corner<1, 2>(l1, h1, lo, hi, q, a.wmac, ze, i, j, k, a.dtdz / 3.0, cs);
      sc(o + A_YZ, i, j, k) = upwind(l1, h1, a.vmac(i, j, k));
    }
  }
  // z-faces
  if (fz && ((fx && cx && gy) || (gx && fy && cy))) {
    es_lohi<2>(a, q, a.wmac, n, i, j, k, a.dtdz, lo, hi);
    if (fx && cx && gy) {  // zx: needed for y-faces
      corner<2, 0>(l1, h1, lo, hi, q, a.umac, xe, i, j, k, a.dtdx / 3.0, cs);
      sc(o + A_ZX, i, j, k) = upwind(l1, h1, a.wmac(i, j, k));
    }
    if (gx && fy && cy) {  // zy: needed for x-faces
      corner<2, 1>(l1, h1, lo, hi, q, a.vmac, ye, i, j, k, a.dtdy / 3.0, cs);
      sc(o + A_ZY, i, j, k) = upwind(l1, h1, a.wmac(i, j, k));
    }
  }
}

// transverse correction of the D-face state from the two corner-coupled arrays:
// t1 lives on D1-faces, t2 on D2-faces (D1, D2 = the two transverse directions).
template <int D, int D1, int D2, class Q>
IX_D void transverse(double& stl, double& sth, const Q& q, const C4& mac1, const C4& mac2, const ScArr& t1,
                     const ScArr& t2, int i, int j, int k, double dtd1, double dtd2, bool conserv) {
  const int im = i - E<D>::x, jm = j - E<D>::y, km = k - E<D>::z;
  const int i1 = E<D1>::x, j1 = E<D1>::y, k1 = E<D1>::z;
  const int i2 = E<D2>::x, j2 = E<D2>::y, k2 = E<D2>::z;
  if (conserv) {
    stl += -(0.5 * dtd1) * (t1(im + i1, jm + j1, km + k1) * mac1(im + i1, jm + j1, km + k1) - t1(im, jm, km) * mac1(im, jm, km))
           - (0.5 * dtd2) * (t2(im + i2, jm + j2, km + k2) * mac2(im + i2, jm + j2, km + k2) - t2(im, jm, km) * mac2(im, jm, km))
           + (0.5 * dtd1) * q(im, jm, km) * (mac1(im + i1, jm + j1, km + k1) - mac1(im, jm, km))
           + (0.5 * dtd2) * q(im, jm, km) * (mac2(im + i2, jm + j2, km + k2) - mac2(im, jm, km));
    sth += -(0.5 * dtd1) * (t1(i + i1, j + j1, k + k1) * mac1(i + i1, j + j1, k + k1) - t1(i, j, k) * mac1(i, j, k))
           - (0.5 * dtd2) * (t2(i + i2, j + j2, k + k2) * mac2(i + i2, j + j2, k + k2) - t2(i, j, k) * mac2(i, j, k))
           + (0.5 * dtd1) * q(i, j, k) * (mac1(i + i1, j + j1, k + k1) - mac1(i, j, k))
           + (0.5 * dtd2) * q(i, j, k) * (mac2(i + i2, j + j2, k + k2) - mac2(i, j, k));
  } else {
    stl += -(0.25 * dtd1) * (mac1(im + i1, jm + j1, km + k1) + mac1(im, jm, km)) * (t1(im + i1, jm + j1, km + k1) - t1(im, jm, km))
           - (0.25 * dtd2) * (mac2(im + i2, jm + j2, km + k2) + mac2(im, jm, km)) * (t2(im + i2, jm + j2, km + k2) - t2(im, jm, km));
    sth += -(0.25 * dtd1) * (mac1(i + i1, j + j1, k + k1) + mac1(i, j, k)) * (t1(i + i1, j + j1, k + k1) - t1(i, j, k))
           - (0.25 * dtd2) * (mac2(i + i2, j + j2, k + k2) + mac2(i, j, k)) * (t2(i + i2, j + j2, k + k2) - t2(i, j, k));
  }
}

template <int D>
IX_D void es_finish(const EsArgs& a, const Comp& q, int n, int i, int j, int k, double& stl, double& sth) {
  const int im = i - E<D>::x, jm = j - E<D>::y, km = k - E<D>::z;
  if (a.iconserv[n] && a.divu.ok()) {
    stl -= 0.5 * a.dt * q(im, jm, km) * a.divu(im, jm, km);
    sth -= 0.5 * a.dt * q(i, j, k) * a.divu(i, j, k);
  }
  if (!a.fit && a.force.ok()) {
    stl += 0.5 * a.dt * a.force(im, jm, km, n);
    sth += 0.5 * a.dt * a.force(i, j, k, n);
  }
}

struct EsOut {
  V4 fx, fy, fz, xed, yed, zed;  // optional user outputs (area-weighted fluxes, edge states)
  double ax, ay, az;             // face areas
};

__global__ void __launch_bounds__(TX* TY) es_final_kernel(IX_KARG(EsArgs) a, IX_KARG(Scratch) sc, IX_KARG(EsOut) out, IX_KARG(Bx) R) {
  GIDX(R)
  const Comp q = comp(a.S, n);
  const Bx& b = a.bx;
  const bool cs = a.iconserv[n] != 0;
  const int o = A_N * n;
  const bool cx = i <= b.hi[0], cy = j <= b.hi[1], cz = k <= b.hi[2];
  double stl, sth;
  if (cy && cz) {  // x-face
    es_lohi<0>(a, q, a.umac, n, i, j, k, a.dtdx, stl, sth);
    transverse<0, 1, 2>(stl, sth, q, a.vmac, a.wmac, arr(sc, o + A_YZ), arr(sc, o + A_ZY), i, j, k, a.dtdy, a.dtdz, cs);
    es_finish<0>(a, q, n, i, j, k, stl, sth);
    const double st = upwind(stl, sth, a.umac(i, j, k));
    const double f = st * a.uflx(i, j, k) * out.ax;
    sc(o + A_XE, i, j, k) = st;
    sc(o + A_FX, i, j, k) = f;
    if (out.xed.ok()) out.xed(i, j, k, n) = st;
    if (out.fx.ok()) out.fx(i, j, k, n) = f;
  }
  if (cx && cz) {  // y-face
    es_lohi<1>(a, q, a.vmac, n, i, j, k, a.dtdy, stl, sth);
    transverse<1, 0, 2>(stl, sth, q, a.umac, a.wmac, arr(sc, o + A_XZ), arr(sc, o + A_ZX), i, j, k, a.dtdx, a.dtdz, cs);
    es_finish<1>(a, q, n, i, j, k, stl, sth);
    const double st = upwind(stl, sth, a.vmac(i, j, k));
    const double f = st * a.vflx(i, j, k) * out.ay;
    sc(o + A_YE, i, j, k) = st;
    sc(o + A_FY, i, j, k) = f;
    if (out.yed.ok()) out.yed(i, j, k, n) = st;
    if (out.fy.ok()) out.fy(i, j, k, n) = f;
  }
  if (cx && cy) {  // z-face
    es_lohi<2>(a, q, a.wmac, n, i, j, k, a.dtdz, stl, sth);
    transverse<2, 0, 1>(stl, sth, q, a.umac, a.vmac, arr(sc, o + A_XY), arr(sc, o + A_YX), i, j, k, a.dtdx, a.dtdy, cs);
    es_finish<2>(a, q, n, i, j, k, stl, sth);
    const double st = upwind(stl, sth, a.wmac(i, j, k));
    const double f = st * a.wflx(i, j, k) * out.az;
    sc(o + A_ZE, i, j, k) = st;
    sc(o + A_FZ, i, j, k) = f;
    if (out.zed.ok()) out.zed(i, j, k, n) = st;
    if (out.fz.ok()) out.fz(i, j, k, n) = f;
  }
}

// ComputeDivergence(mult=-1, area-weighted) + div(umac) + ComputeConvectiveTerm + sign
__global__ void __launch_bounds__(TX* TY)
es_div_kernel(IX_KARG(EsArgs) a, IX_KARG(Scratch) sc, V4 aofs, double volinv, double dxi, double dyi, double dzi, int is_sync, Bx R) {
  GIDX(R)
  const int o = A_N * n;
  double upd = -volinv * ((sc(o + A_FX, i + 1, j, k) - sc(o + A_FX, i, j, k)) +
                          (sc(o + A_FY, i, j + 1, k) - sc(o + A_FY, i, j, k)) +
                          (sc(o + A_FZ, i, j, k + 1) - sc(o + A_FZ, i, j, k)));
  if (!a.iconserv[n] && !is_sync) {
    const double divum = dxi * (a.umac(i + 1, j, k) - a.umac(i, j, k)) + dyi * (a.vmac(i, j + 1, k) - a.vmac(i, j, k)) +
                         dzi * (a.wmac(i, j, k + 1) - a.wmac(i, j, k));
    double qb = sc(o + A_XE, i, j, k) + sc(o + A_XE, i + 1, j, k) + sc(o + A_YE, i, j, k) + sc(o + A_YE, i, j + 1, k) +
                sc(o + A_ZE, i, j, k) + sc(o + A_ZE, i, j, k + 1);
    qb /= 6.0;
    upd += qb * divum;
  }
  if (is_sync) aofs(i, j, k, n) -= upd;
  else aofs(i, j, k, n) = -upd;
}

// ===========================================================================
// ExtrapVelToFaces path
// ===========================================================================
struct EvArgs {
  Bx bx;
  C4 vel, force;
  int fit;
  double dt, dtdx, dtdy, dtdz;
};
// scratch ids: advective velocities 0..2; transverse edges: XE_V,XE_W (x-faces, comps 1,2),
// YE_U,YE_W, ZE_U,ZE_V; corner: YZ_U, ZY_U (for umac), XZ_V, ZX_V (vmac), XY_W, YX_W (wmac)
enum { B_UAD = 0, B_VAD, B_WAD, B_XE_V, B_XE_W, B_YE_U, B_YE_W, B_ZE_U, B_ZE_V,
       B_YZ_U, B_ZY_U, B_XZ_V, B_ZX_V, B_XY_W, B_YX_W, B_N };

template <int D>
IX_D void ev_lohi(const EvArgs& a, int n, int i, int j, int k, double dtdx, double& lo, double& hi) {
  const int im = i - E<D>::x, jm = j - E<D>::y, km = k - E<D>::z;
  trace<D>(comp(a.vel, n), i, j, k, a.vel(im, jm, km, D), a.vel(i, j, k, D), dtdx, lo, hi);
  if (a.fit && a.force.ok()) {
    lo += 0.5 * a.dt * a.force(im, jm, km, n);
    hi += 0.5 * a.dt * a.force(i, j, k, n);
  }
}

__global__ void __launch_bounds__(TX* TY) ev_edge_kernel(IX_KARG(EvArgs) a, IX_KARG(Scratch) sc, IX_KARG(Bx) R) {
  GIDX(R)
  (void)n;
  const Bx& b = a.bx;
  const bool inx = i >= b.lo[0] - 1 && i <= b.hi[0] + 1, iny = j >= b.lo[1] - 1 && j <= b.hi[1] + 1,
             inz = k >= b.lo[2] - 1 && k <= b.hi[2] + 1;
  double lo, hi;
  if (i >= b.lo[0] && i <= b.hi[0] + 1 && iny && inz) {
    ev_lohi<0>(a, 0, i, j, k, a.dtdx, lo, hi);
    const double uad = riemann(lo, hi);
    sc(B_UAD, i, j, k) = uad;
    ev_lohi<0>(a, 1, i, j, k, a.dtdx, lo, hi);
    sc(B_XE_V, i, j, k) = upwind(lo, hi, uad);
    ev_lohi<0>(a, 2, i, j, k, a.dtdx, lo, hi);
    sc(B_XE_W, i, j, k) = upwind(lo, hi, uad);
  }
  if (j >= b.lo[1] && j <= b.hi[1] + 1 && inx && inz) {
    ev_lohi<1>(a, 1, i, j, k, a.dtdy, lo, hi);
    const double vad = riemann(lo, hi);
    sc(B_VAD, i, j, k) = vad;
    ev_lohi<1>(a, 0, i, j, k, a.dtdy, lo, hi);
    sc(B_YE_U, i, j, k) = upwind(lo, hi, vad);
    ev_lohi<1>(a, 2, i, j, k, a.dtdy, lo, hi);
    sc(B_YE_W, i, j, k) = upwind(lo, hi, vad);
  }
  if (k >= b.lo[2] && k <= b.hi[2] + 1 && inx && iny) {
    ev_lohi<2>(a, 2, i, j, k, a.dtdz, lo, hi);
    const double wad = riemann(lo, hi);
    sc(B_WAD, i, j, k) = wad;
    ev_lohi<2>(a, 0, i, j, k, a.dtdz, lo, hi);
    sc(B_ZE_U, i, j, k) = upwind(lo, hi, wad);
    ev_lohi<2>(a, 1, i, j, k, a.dtdz, lo, hi);
    sc(B_ZE_V, i, j, k) = upwind(lo, hi, wad);
  }
}

__global__ void __launch_bounds__(TX* TY) ev_corner_kernel(IX_KARG(EvArgs) a, IX_KARG(Scratch) sc, IX_KARG(Bx) R) {
  GIDX(R)
  (void)n;
  const Bx& b = a.bx;
  const ScArr uad = arr(sc, B_UAD), vad = arr(sc, B_VAD), wad = arr(sc, B_WAD);
  const bool cx = i <= b.hi[0], cy = j <= b.hi[1], cz = k <= b.hi[2];
  const bool gx = i >= b.lo[0] - 1 && i <= b.hi[0] + 1, gy = j >= b.lo[1] - 1 && j <= b.hi[1] + 1,
             gz = k >= b.lo[2] - 1 && k <= b.hi[2] + 1;
  const bool fx = i >= b.lo[0] && i <= b.hi[0] + 1, fy = j >= b.lo[1] && j <= b.hi[1] + 1,
             fz = k >= b.lo[2] && k <= b.hi[2] + 1;
  double lo, hi, l1, h1;
  // x-face states: comp 2 coupled with y (for wmac), comp 1 coupled with z (for vmac)
  if (fx && fy && cy && gz) {
    ev_lohi<0>(a, 2, i, j, k, a.dtdx, lo, hi);
    corner<0, 1>(l1, h1, lo, hi, comp(a.vel, 2), vad, arr(sc, B_YE_W), i, j, k, a.dtdy / 3.0, false);
    sc(B_XY_W, i, j, k) = upwind(l1, h1, uad(i, j, k));
  }
  if (fx && gy && fz && cz) {
    ev_lohi<0>(a, 1, i, j, k, a.dtdx, lo, hi);
    corner<0, 2>(l1, h1, lo, hi, comp(a.vel, 1), wad, arr(sc, B_ZE_V), i, j, k, a.dtdz / 3.0, false);
    sc(B_XZ_V, i, j, k) = upwind(l1, h1, uad(i, j, k));
  }
  // y-face states: comp 2 coupled with x (for wmac), comp 0 coupled with z (for umac)
  if (fy && fx && cx && gz) {
    ev_lohi<1>(a, 2, i, j, k, a.dtdy, lo, hi);
    corner<1, 0>(l1, h1, lo, hi, comp(a.vel, 2), uad, arr(sc, B_XE_W), i, j, k, a.dtdx / 3.0, false);
    sc(B_YX_W, i, j, k) = upwind(l1, h1, vad(i, j, k));
  }
  if (fy && gx && fz && cz) {
    ev_lohi<1>(a, 0, i, j, k, a.dtdy, lo, hi);
    corner<1, 2>(l1, h1, lo, hi, comp(a.vel, 0), wad, arr(sc, B_ZE_U), i, j, k, a.dtdz / 3.0, false);
    sc(B_YZ_U, i, j, k) = upwind(l1, h1, vad(i, j, k));
  }
  // z-face states: comp 1 coupled with x (for vmac), comp 0 coupled with y (for umac)
  if (fz && fx && cx && gy) {
    ev_lohi<2>(a, 1, i, j, k, a.dtdz, lo, hi);
    corner<2, 0>(l1, h1, lo, hi, comp(a.vel, 1), uad, arr(sc, B_XE_V), i, j, k, a.dtdx / 3.0, false);
    sc(B_ZX_V, i, j, k) = upwind(l1, h1, wad(i, j, k));
  }
  if (fz && gx && fy && cy) {
    ev_lohi<2>(a, 0, i, j, k, a.dtdz, lo, hi);
    corner<2, 1>(l1, h1, lo, hi, comp(a.vel, 0), vad, arr(sc, B_YE_U), i, j, k, a.dtdy / 3.0, false);
    sc(B_ZY_U, i, j, k) = upwind(l1, h1, wad(i, j, k));
  }
}

template <int D, int D1, int D2>
IX_D double ev_final(const EvArgs& a, const Scratch& sc, int A1, int A2, int T1, int T2, int i, int j, int k,
                     double dtdx, double dtd1, double dtd2) {
  const int im = i - E<D>::x, jm = j - E<D>::y, km = k - E<D>::z;
  const int i1 = E<D1>::x, j1 = E<D1>::y, k1 = E<D1>::z;
  const int i2 = E<D2>::x, j2 = E<D2>::y, k2 = E<D2>::z;
  const ScArr ad1 = arr(sc, A1), ad2 = arr(sc, A2), t1 = arr(sc, T1), t2 = arr(sc, T2);
  double stl, sth;
  ev_lohi<D>(a, D, i, j, k, dtdx, stl, sth);
  stl += -(0.25 * dtd1) * (ad1(im + i1, jm + j1, km + k1) + ad1(im, jm, km)) * (t1(im + i1, jm + j1, km + k1) - t1(im, jm, km))
         - (0.25 * dtd2) * (ad2(im + i2, jm + j2, km + k2) + ad2(im, jm, km)) * (t2(im + i2, jm + j2, km + k2) - t2(im, jm, km));
  sth += -(0.25 * dtd1) * (ad1(i + i1, j + j1, k + k1) + ad1(i, j, k)) * (t1(i + i1, j + j1, k + k1) - t1(i, j, k))
         - (0.25 * dtd2) * (ad2(i + i2, j + j2, k + k2) + ad2(i, j, k)) * (t2(i + i2, j + j2, k + k2) - t2(i, j, k));
  if (!a.fit && a.force.ok()) {
    stl += 0.5 * a.dt * a.force(im, jm, km, D);
    sth += 0.5 * a.dt * a.force(i, j, k, D);
  }
  return riemann(stl, sth);
}

__global__ void __launch_bounds__(TX* TY) ev_final_kernel(IX_KARG(EvArgs) a, IX_KARG(Scratch) sc, V4 umac, V4 vmac, V4 wmac, IX_KARG(Bx) R) {
  GIDX(R)
  (void)n;
  const Bx& b = a.bx;
  const bool cx = i <= b.hi[0], cy = j <= b.hi[1], cz = k <= b.hi[2];
  if (cy && cz) umac(i, j, k) = ev_final<0, 1, 2>(a, sc, B_VAD, B_WAD, B_YZ_U, B_ZY_U, i, j, k, a.dtdx, a.dtdy, a.dtdz);
  if (cx && cz) vmac(i, j, k) = ev_final<1, 0, 2>(a, sc, B_UAD, B_WAD, B_XZ_V, B_ZX_V, i, j, k, a.dtdy, a.dtdx, a.dtdz);
  if (cx && cy) wmac(i, j, k) = ev_final<2, 0, 1>(a, sc, B_UAD, B_VAD, B_XY_W, B_YX_W, i, j, k, a.dtdz, a.dtdx, a.dtdy);
}

struct ScratchOwner {
  double* p = nullptr;
  Scratch sc{};
  int init(const Bx& bx, int narrays) {
    const Bx g = grow(bx, 2);
    sc.l0 = g.lo[0]; sc.l1 = g.lo[1]; sc.l2 = g.lo[2];
    sc.js = ((int64_t)g.nx() + 15) / 16 * 16;
    sc.ks = sc.js * g.ny();
    sc.as = sc.ks * g.nz();
    p = dev_alloc((size_t)(sc.as * narrays));
    sc.p = p;
    return p ? IAMRX_OK : IAMRX_ERR_CUDA;
  }
  ~ScratchOwner() { dev_free(p); }  // stream-ordered reuse (single stream per rank)
};

}  // namespace

int compute_aofs(const Bx& bx, const AofsArgs& a, const AdvGeom& g, cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  ProfScope prof_(IAMRX_PROF_AOFS, bx.npts(), (double)bx.npts() * (24.0 * a.ncomp + 32.0), s);
  ScratchOwner so;
  if (so.init(bx, A_N * a.ncomp) != IAMRX_OK) return IAMRX_ERR_CUDA;
  EsArgs e{};
  e.bx = bx;
  e.S = a.S; e.force = a.force; e.divu = a.divu;
  e.umac = a.umac; e.vmac = a.vmac; e.wmac = a.wmac;
  e.uflx = a.uflx; e.vflx = a.vflx; e.wflx = a.wflx;
  for (int n = 0; n < 8; ++n) e.iconserv[n] = (n < a.ncomp) ? a.iconserv[n] : 0;
  e.fit = a.forces_in_trans;
  e.dt = g.dt; e.dtdx = g.dt / g.dx[0]; e.dtdy = g.dt / g.dx[1]; e.dtdz = g.dt / g.dx[2];
  Bx R1 = grow(bx, 1); R1.hi[0]++; R1.hi[1]++; R1.hi[2]++;
  IX_LAUNCH(es_edge_kernel, grid_for(R1, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
  int rc = check_launch("es_edge");
  if (rc) return rc;
  IX_LAUNCH(es_corner_kernel, grid_for(R1, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
  rc = check_launch("es_corner");
  if (rc) return rc;
  Bx R2 = bx; R2.hi[0]++; R2.hi[1]++; R2.hi[2]++;
  EsOut out{};
  if (a.write_fluxes) { out.fx = a.fx; out.fy = a.fy; out.fz = a.fz; out.xed = a.xed; out.yed = a.yed; out.zed = a.zed; }
  out.ax = g.dx[1] * g.dx[2]; out.ay = g.dx[0] * g.dx[2]; out.az = g.dx[0] * g.dx[1];
  IX_LAUNCH(es_final_kernel, grid_for(R2, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, out, R2);
  rc = check_launch("es_final");
  if (rc) return rc;
  IX_LAUNCH(es_div_kernel, grid_for(bx, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, a.aofs,
            1.0 / (g.dx[0] * g.dx[1] * g.dx[2]), 1.0 / g.dx[0], 1.0 / g.dx[1], 1.0 / g.dx[2], a.is_sync, bx);
  return check_launch("es_div");
}

int extrap_vel_to_faces(const Bx& bx, C4 vel, C4 force, V4 umac, V4 vmac, V4 wmac, const AdvGeom& g,
                        int forces_in_trans, cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  ProfScope prof_(IAMRX_PROF_EXTRAP, bx.npts(), (double)bx.npts() * 72.0, s);
  ScratchOwner so;
  if (so.init(bx, B_N) != IAMRX_OK) return IAMRX_ERR_CUDA;
  EvArgs e{};
  e.bx = bx; e.vel = vel; e.force = force; e.fit = forces_in_trans;
  e.dt = g.dt; e.dtdx = g.dt / g.dx[0]; e.dtdy = g.dt / g.dx[1]; e.dtdz = g.dt / g.dx[2];
  Bx R1 = grow(bx, 1); R1.hi[0]++; R1.hi[1]++; R1.hi[2]++;
  IX_LAUNCH(ev_edge_kernel, grid_for(R1, 1), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
  int rc = check_launch("ev_edge");
  if (rc) return rc;
  IX_LAUNCH(ev_corner_kernel, grid_for(R1, 1), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
  rc = check_launch("ev_corner");
  if (rc) return rc;
  Bx R2 = bx; R2.hi[0]++; R2.hi[1]++; R2.hi[2]++;
  IX_LAUNCH(ev_final_kernel, grid_for(R2, 1), dim3(TX, TY, 1), 0, s, e, so.sc, umac, vmac, wmac, R2);
  return check_launch("ev_final");
}

}  // namespace k
}  // namespace ix
