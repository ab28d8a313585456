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
// Physical domain boundaries (BCRec per component, NSB.cpp:4489,4712): one-sided ext_dir / hoextrap slopes, the
// Set{X,Y,Z}EdgeBCs rules after every stage and the outflow clipping of the final states (godunov_math.h) in the
// staged kernels; the fused tile kernel handles boxes whose stencil stays away from non-interior boundaries.
#include <cstdlib>
#include "godunov_math.h"
#include "kernels.h"
#include "level.h"

namespace ix {
namespace k {
namespace {

using namespace gd;

constexpr int TX = 64;
constexpr int TY = 4;

// traced states on the D-face below the cell the cursor `q` sits on: lo from the cell below,
// hi from this cell.  ulo/uhi: the velocity used in the trace (cell-centred normal velocity
// for ExtrapVelToFaces, the MAC velocity of this face for ComputeEdgeState).
// domain bounds + BCRec of every component (amrex::BCRec, NS_BC.H:7-55 via NavierStokesBase::fetchBCArray)
struct BcAll {
  int dlo[3], dhi[3];
  int lo[8][3], hi[8][3];
  int any;   // some side of some component is not int_dir: otherwise every boundary branch is skipped
  IX_HD BcD dir(int n, int d) const { return BcD{lo[n][d], hi[n][d], dlo[d], dhi[d]}; }
};
template <int D> IX_HD int idx_of(int i, int j, int k) { return D == 0 ? i : (D == 1 ? j : k); }

template <int D>
IX_D void trace(const Cur& q, double ulo, double uhi, double dtdx, double& lo, double& hi, int ppm, int c, const BcD* b) {
  const Cur qm = below<D>(q);   // c = index along D of the cell the cursor sits on (== index of the face)
  if (ppm) {  // Godunov_PPM: lo = Ip of the cell below, hi = Im of this cell
    double sm, sp;
    if (b) ppm_parabola_bc(along<D>(qm, -2), along<D>(qm, -1), qm(0, 0, 0), along<D>(qm, 1), along<D>(qm, 2), c - 1, *b, sm, sp);
    else ppm_parabola(along<D>(qm, -2), along<D>(qm, -1), qm(0, 0, 0), along<D>(qm, 1), along<D>(qm, 2), sm, sp);
    lo = ppm_ip(qm(0, 0, 0), sm, sp, ulo, dtdx);
    if (b) ppm_parabola_bc(along<D>(q, -2), along<D>(q, -1), q(0, 0, 0), along<D>(q, 1), along<D>(q, 2), c, *b, sm, sp);
    else ppm_parabola(along<D>(q, -2), along<D>(q, -1), q(0, 0, 0), along<D>(q, 1), along<D>(q, 2), sm, sp);
    hi = ppm_im(q(0, 0, 0), sm, sp, uhi, dtdx);
    return;
  }
  const double sl = b ? slope4_bc_vals(along<D>(qm, -2), along<D>(qm, -1), qm(0, 0, 0), along<D>(qm, 1), along<D>(qm, 2), c - 1, *b) : slope4c<D>(qm);
  const double sh = b ? slope4_bc_vals(along<D>(q, -2), along<D>(q, -1), q(0, 0, 0), along<D>(q, 1), along<D>(q, 2), c, *b) : slope4c<D>(q);
  lo = qm(0, 0, 0) + 0.5 * (1.0 - ulo * dtdx) * sl;
  hi = q(0, 0, 0) + 0.5 * (-1.0 - uhi * dtdx) * sh;
}

struct Scratch {  // all on the same grown index box, component-major
  double* p; int l0, l1, l2; int64_t js, ks, as;  // as = stride between arrays
};
// scratch cursor at (i,j,k): array a, relative offsets
struct SCur {
  double* p; int js, ks; int64_t as;
  IX_D double& operator()(int a, int di, int dj, int dk) const { return p[a * as + (di + dj * js + dk * ks)]; }
};
IX_D SCur scur_at(const Scratch& s, int i, int j, int k) {
  return SCur{s.p + ((i - s.l0) + (j - s.l1) * s.js + (k - s.l2) * s.ks), (int)s.js, (int)s.ks, s.as};
}
struct SArr {  // one scratch array of an SCur as a relative-offset callable
  const double* p; int js, ks;
  IX_D double operator()(int di, int dj, int dk) const { return p[di + dj * js + dk * ks]; }
};
IX_D SArr sarr(const SCur& s, int a) { return SArr{s.p + a * s.as, s.js, s.ks}; }

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
  int ppm;  // Godunov_PPM instead of Godunov_PLM
  int is_velocity;
  double dt, dtdx, dtdy, dtdz;
  BcAll bc;
};

// scratch array ids per component: 0..2 edge x,y,z; 3..8 corner xy,xz,yx,yz,zx,zy; 9..11 flux
enum { A_XE = 0, A_YE, A_ZE, A_XY, A_XZ, A_YX, A_YZ, A_ZX, A_ZY, A_FX, A_FY, A_FZ, A_N };

// lo/hi on the D-face of the cursor's cell, traced with the MAC velocity `u` of that face
// Set{X,Y,Z}EdgeBCs on the states of the D-face with index c for component n
template <int D, bool BC, class A>
IX_D void edge_bc(const A& a, int n, const Cur& q, int c, double& lo, double& hi, bool normal_vel) {
  if (!BC) return;
  set_edge_bc(lo, hi, along<D>(q, -1), q(0, 0, 0), c, a.bc.dir(n, D), normal_vel);
}
template <int D, bool BC>
IX_D void es_lohi(const EsArgs& a, int n, int c, const Cur& q, const Cur& f, double u, double dtdx, double& lo, double& hi) {
  const BcD b = BC ? a.bc.dir(n, D) : BcD{};
  trace<D>(q, u, u, dtdx, lo, hi, a.ppm, c, BC ? &b : nullptr);
  if (a.fit && f.ok()) {
    lo += 0.5 * a.dt * along<D>(f, -1);
    hi += 0.5 * a.dt * f(0, 0, 0);
  }
  edge_bc<D, BC>(a, n, q, c, lo, hi, a.is_velocity && n == D);
}

struct EsCur {  // the cursors one thread needs
  Cur q, f, u, v, w;
  SCur s;
};
IX_D EsCur es_cursors(const EsArgs& a, const Scratch& sc, int n, int i, int j, int k) {
  return EsCur{cur_at(a.S, n, i, j, k), cur_at(a.force, n, i, j, k), cur_at(a.umac, 0, i, j, k),
               cur_at(a.vmac, 0, i, j, k), cur_at(a.wmac, 0, i, j, k), scur_at(sc, i, j, k)};
}

template <bool BC>
__global__ void __launch_bounds__(TX* TY) es_edge_kernel(IX_KARG(EsArgs) a, IX_KARG(Scratch) sc, IX_KARG(Bx) R) {
  GIDX(R)
  const EsCur c = es_cursors(a, sc, n, i, j, k);
  const Bx& b = a.bx;
  const int o = A_N * n;
  const bool inx = i >= b.lo[0] - 1 && i <= b.hi[0] + 1, iny = j >= b.lo[1] - 1 && j <= b.hi[1] + 1,
             inz = k >= b.lo[2] - 1 && k <= b.hi[2] + 1;
  double lo, hi;
  if (i >= b.lo[0] && i <= b.hi[0] + 1 && iny && inz) {  // x-faces lo..hi+1
    const double u = c.u(0, 0, 0);
    es_lohi<0, BC>(a, n, i, c.q, c.f, u, a.dtdx, lo, hi);
    c.s(o + A_XE, 0, 0, 0) = upwind(lo, hi, u);
  }
  if (j >= b.lo[1] && j <= b.hi[1] + 1 && inx && inz) {
    const double v = c.v(0, 0, 0);
    es_lohi<1, BC>(a, n, j, c.q, c.f, v, a.dtdy, lo, hi);
    c.s(o + A_YE, 0, 0, 0) = upwind(lo, hi, v);
  }
  if (k >= b.lo[2] && k <= b.hi[2] + 1 && inx && iny) {
    const double w = c.w(0, 0, 0);
    es_lohi<2, BC>(a, n, k, c.q, c.f, w, a.dtdz, lo, hi);
    c.s(o + A_ZE, 0, 0, 0) = upwind(lo, hi, w);
  }
}

// corner coupling of the D1-face state with the derivative along D2
// (Godunov_corner_couple_<d1><d2>): mac2/edge2 live on D2-faces; all offsets are relative
// to the cell above the D1-face.
template <int D1, int D2, class M, class Ed>
IX_D void corner(double& lo1, double& hi1, double lo, double hi, const Cur& q, const M& mac2, const Ed& edge2,
                 double dt3dx, bool conserv) {
  const double mlo_p = rel2<D1, D2>(mac2, -1, 1), mlo_m = rel2<D1, D2>(mac2, -1, 0);
  const double mhi_p = rel2<D1, D2>(mac2, 0, 1), mhi_m = mac2(0, 0, 0);
  if (!conserv && upopts().corner_adv) {   // UNVERIFIED-UPSTREAM alternative: advective form of the transverse derivative
    lo1 = lo - dt3dx * 0.5 * (mlo_p + mlo_m) * (rel2<D1, D2>(edge2, -1, 1) - rel2<D1, D2>(edge2, -1, 0));
    hi1 = hi - dt3dx * 0.5 * (mhi_p + mhi_m) * (rel2<D1, D2>(edge2, 0, 1) - edge2(0, 0, 0));
    return;
  }
  lo1 = lo - dt3dx * (rel2<D1, D2>(edge2, -1, 1) * mlo_p - rel2<D1, D2>(edge2, -1, 0) * mlo_m);
  hi1 = hi - dt3dx * (rel2<D1, D2>(edge2, 0, 1) * mhi_p - edge2(0, 0, 0) * mhi_m);
  if (!conserv) {
    lo1 += dt3dx * along<D1>(q, -1) * (mlo_p - mlo_m);
    hi1 += dt3dx * q(0, 0, 0) * (mhi_p - mhi_m);
  }
}

template <bool BC>
__global__ void __launch_bounds__(TX* TY) es_corner_kernel(IX_KARG(EsArgs) a, IX_KARG(Scratch) sc, IX_KARG(Bx) R) {
  GIDX(R)
  const EsCur c = es_cursors(a, sc, n, i, j, k);
  const Bx& b = a.bx;
  const bool cs = a.iconserv[n] != 0;
  const int o = A_N * n;
  const SArr xe = sarr(c.s, o + A_XE), ye = sarr(c.s, o + A_YE), ze = sarr(c.s, o + A_ZE);
  const bool cx = i <= b.hi[0], cy = j <= b.hi[1], cz = k <= b.hi[2];           // cell index inside bx (upper)
  const bool gx = i >= b.lo[0] - 1 && i <= b.hi[0] + 1, gy = j >= b.lo[1] - 1 && j <= b.hi[1] + 1,
             gz = k >= b.lo[2] - 1 && k <= b.hi[2] + 1;
  const bool fx = i >= b.lo[0] && i <= b.hi[0] + 1, fy = j >= b.lo[1] && j <= b.hi[1] + 1,
             fz = k >= b.lo[2] && k <= b.hi[2] + 1;  // face index of bx
  double lo, hi, l1, h1;
  // x-faces
  if (fx && ((fy && cy && gz) || (gy && fz && cz))) {
    const double u = c.u(0, 0, 0);
    const bool nv = a.is_velocity && n == 0;
    es_lohi<0, BC>(a, n, i, c.q, c.f, u, a.dtdx, lo, hi);
    if (fy && cy && gz) {  // xy: needed for z-faces -> grown in z
      corner<0, 1>(l1, h1, lo, hi, c.q, c.v, ye, a.dtdy / 3.0, cs);
      edge_bc<0, BC>(a, n, c.q, i, l1, h1, nv);
      c.s(o + A_XY, 0, 0, 0) = upwind(l1, h1, u);
    }
    if (gy && fz && cz) {  // xz: needed for y-faces -> grown in y
      corner<0, 2>(l1, h1, lo, hi, c.q, c.w, ze, a.dtdz / 3.0, cs);
      edge_bc<0, BC>(a, n, c.q, i, l1, h1, nv);
      c.s(o + A_XZ, 0, 0, 0) = upwind(l1, h1, u);
    }
  }
  // y-faces
  if (fy && ((fx && cx && gz) || (gx && fz && cz))) {
    const double v = c.v(0, 0, 0);
    const bool nv = a.is_velocity && n == 1;
    es_lohi<1, BC>(a, n, j, c.q, c.f, v, a.dtdy, lo, hi);
    if (fx && cx && gz) {  // yx: needed for z-faces
      corner<1, 0>(l1, h1, lo, hi, c.q, c.u, xe, a.dtdx / 3.0, cs);
      edge_bc<1, BC>(a, n, c.q, j, l1, h1, nv);
      c.s(o + A_YX, 0, 0, 0) = upwind(l1, h1, v);
    }
    if (gx && fz && cz) {  // yz: needed for x-faces
      corner<1, 2>(l1, h1, lo, hi, c.q, c.w, ze, a.dtdz / 3.0, cs);
      edge_bc<1, BC>(a, n, c.q, j, l1, h1, nv);
      c.s(o + A_YZ, 0, 0, 0) = upwind(l1, h1, v);
    }
  }
  // z-faces
  if (fz && ((fx && cx && gy) || (gx && fy && cy))) {
    const double w = c.w(0, 0, 0);
    const bool nv = a.is_velocity && n == 2;
    es_lohi<2, BC>(a, n, k, c.q, c.f, w, a.dtdz, lo, hi);
    if (fx && cx && gy) {  // zx: needed for y-faces
      corner<2, 0>(l1, h1, lo, hi, c.q, c.u, xe, a.dtdx / 3.0, cs);
      edge_bc<2, BC>(a, n, c.q, k, l1, h1, nv);
      c.s(o + A_ZX, 0, 0, 0) = upwind(l1, h1, w);
    }
    if (gx && fy && cy) {  // zy: needed for x-faces
      corner<2, 1>(l1, h1, lo, hi, c.q, c.v, ye, a.dtdy / 3.0, cs);
      edge_bc<2, BC>(a, n, c.q, k, l1, h1, nv);
      c.s(o + A_ZY, 0, 0, 0) = upwind(l1, h1, w);
    }
  }
}

// transverse correction of the D-face state from the two corner-coupled arrays:
// t1 lives on D1-faces, t2 on D2-faces (D1, D2 = the two transverse directions).
template <int D, int D1, int D2, class M1, class M2, class T1, class T2>
IX_D void transverse(double& stl, double& sth, const Cur& q, const M1& mac1, const M2& mac2, const T1& t1, const T2& t2,
                     double dtd1, double dtd2, bool conserv) {
  // offsets: (a, b) = a*e_D + b*e_Dt ; the cell below the face is a = -1
  const double m1lp = rel2<D, D1>(mac1, -1, 1), m1lm = rel2<D, D1>(mac1, -1, 0), m1hp = rel2<D, D1>(mac1, 0, 1), m1hm = mac1(0, 0, 0);
  const double m2lp = rel2<D, D2>(mac2, -1, 1), m2lm = rel2<D, D2>(mac2, -1, 0), m2hp = rel2<D, D2>(mac2, 0, 1), m2hm = mac2(0, 0, 0);
  const double t1lp = rel2<D, D1>(t1, -1, 1), t1lm = rel2<D, D1>(t1, -1, 0), t1hp = rel2<D, D1>(t1, 0, 1), t1hm = t1(0, 0, 0);
  const double t2lp = rel2<D, D2>(t2, -1, 1), t2lm = rel2<D, D2>(t2, -1, 0), t2hp = rel2<D, D2>(t2, 0, 1), t2hm = t2(0, 0, 0);
  if (conserv) {
    const double ql = along<D>(q, -1), qh = q(0, 0, 0);
    stl += -(0.5 * dtd1) * (t1lp * m1lp - t1lm * m1lm) - (0.5 * dtd2) * (t2lp * m2lp - t2lm * m2lm)
           + (0.5 * dtd1) * ql * (m1lp - m1lm) + (0.5 * dtd2) * ql * (m2lp - m2lm);
    sth += -(0.5 * dtd1) * (t1hp * m1hp - t1hm * m1hm) - (0.5 * dtd2) * (t2hp * m2hp - t2hm * m2hm)
           + (0.5 * dtd1) * qh * (m1hp - m1hm) + (0.5 * dtd2) * qh * (m2hp - m2hm);
  } else {
    stl += -(0.25 * dtd1) * (m1lp + m1lm) * (t1lp - t1lm) - (0.25 * dtd2) * (m2lp + m2lm) * (t2lp - t2lm);
    sth += -(0.25 * dtd1) * (m1hp + m1hm) * (t1hp - t1hm) - (0.25 * dtd2) * (m2hp + m2hm) * (t2hp - t2hm);
  }
}

template <int D>
IX_D void es_finish(const EsArgs& a, const Cur& q, const Cur& f, const Cur& dv, bool conserv, double& stl, double& sth) {
  if (conserv && dv.ok()) {
    stl -= 0.5 * a.dt * along<D>(q, -1) * along<D>(dv, -1);
    sth -= 0.5 * a.dt * q(0, 0, 0) * dv(0, 0, 0);
  }
  if (!a.fit && f.ok()) {
    stl += 0.5 * a.dt * along<D>(f, -1);
    sth += 0.5 * a.dt * f(0, 0, 0);
  }
}

// boundary conditions of the final states: SetEdgeBCs, then the outflow rule (the normal velocity is not advected
// INTO the domain through a foextrap / hoextrap face)
template <int D, bool BC>
IX_D void es_final_bc(const EsArgs& a, int n, const Cur& q, int c, double mac, double& stl, double& sth) {
  if (!BC) return;
  const bool nv = a.is_velocity && n == D;
  set_edge_bc(stl, sth, along<D>(q, -1), q(0, 0, 0), c, a.bc.dir(n, D), nv);
  outflow_bc(stl, sth, c, a.bc.dir(n, D), nv && mac >= 0.0, nv && mac <= 0.0);
}

struct EsOut {
  V4 fx, fy, fz, xed, yed, zed;  // optional user outputs (area-weighted fluxes, edge states)
  double ax, ay, az;             // face areas
};

template <bool BC>
__global__ void __launch_bounds__(TX* TY) es_final_kernel(IX_KARG(EsArgs) a, IX_KARG(Scratch) sc, IX_KARG(EsOut) out, IX_KARG(Bx) R) {
  GIDX(R)
  const EsCur c = es_cursors(a, sc, n, i, j, k);
  const Cur dv = cur_at(a.divu, 0, i, j, k);
  const Bx& b = a.bx;
  const bool cs = a.iconserv[n] != 0;
  const int o = A_N * n;
  const bool cx = i <= b.hi[0], cy = j <= b.hi[1], cz = k <= b.hi[2];
  double stl, sth;
  if (cy && cz) {  // x-face
    const double u = c.u(0, 0, 0);
    es_lohi<0, BC>(a, n, i, c.q, c.f, u, a.dtdx, stl, sth);
    transverse<0, 1, 2>(stl, sth, c.q, c.v, c.w, sarr(c.s, o + A_YZ), sarr(c.s, o + A_ZY), a.dtdy, a.dtdz, cs);
    es_finish<0>(a, c.q, c.f, dv, cs, stl, sth);
    es_final_bc<0, BC>(a, n, c.q, i, u, stl, sth);
    const double st = upwind(stl, sth, u);
    const double f = st * a.uflx(i, j, k) * out.ax;
    c.s(o + A_XE, 0, 0, 0) = st;
    c.s(o + A_FX, 0, 0, 0) = f;
    if (out.xed.ok()) out.xed(i, j, k, n) = st;
    if (out.fx.ok()) out.fx(i, j, k, n) = f;
  }
  if (cx && cz) {  // y-face
    const double v = c.v(0, 0, 0);
    es_lohi<1, BC>(a, n, j, c.q, c.f, v, a.dtdy, stl, sth);
    transverse<1, 0, 2>(stl, sth, c.q, c.u, c.w, sarr(c.s, o + A_XZ), sarr(c.s, o + A_ZX), a.dtdx, a.dtdz, cs);
    es_finish<1>(a, c.q, c.f, dv, cs, stl, sth);
    es_final_bc<1, BC>(a, n, c.q, j, v, stl, sth);
    const double st = upwind(stl, sth, v);
    const double f = st * a.vflx(i, j, k) * out.ay;
    c.s(o + A_YE, 0, 0, 0) = st;
    c.s(o + A_FY, 0, 0, 0) = f;
    if (out.yed.ok()) out.yed(i, j, k, n) = st;
    if (out.fy.ok()) out.fy(i, j, k, n) = f;
  }
  if (cx && cy) {  // z-face
    const double w = c.w(0, 0, 0);
    es_lohi<2, BC>(a, n, k, c.q, c.f, w, a.dtdz, stl, sth);
    transverse<2, 0, 1>(stl, sth, c.q, c.u, c.v, sarr(c.s, o + A_XY), sarr(c.s, o + A_YX), a.dtdx, a.dtdy, cs);
    es_finish<2>(a, c.q, c.f, dv, cs, stl, sth);
    es_final_bc<2, BC>(a, n, c.q, k, w, stl, sth);
    const double st = upwind(stl, sth, w);
    const double f = st * a.wflx(i, j, k) * out.az;
    c.s(o + A_ZE, 0, 0, 0) = st;
    c.s(o + A_FZ, 0, 0, 0) = f;
    if (out.zed.ok()) out.zed(i, j, k, n) = st;
    if (out.fz.ok()) out.fz(i, j, k, n) = f;
  }
}

// known_edge_state (MacProj.cpp:776-785 -> NSB.cpp:4708): the caller's edge states are taken as they are and only the
// fluxes are formed: flux = edge * uflux * area (HydroUtils::ComputeFluxes)
__global__ void __launch_bounds__(TX* TY) es_known_kernel(IX_KARG(EsArgs) a, IX_KARG(Scratch) sc, IX_KARG(EsOut) out, IX_KARG(Bx) R) {
  GIDX(R)
  const SCur s = scur_at(sc, i, j, k);
  const Bx& b = a.bx;
  const int o = A_N * n;
  const bool cx = i <= b.hi[0], cy = j <= b.hi[1], cz = k <= b.hi[2];
  if (cy && cz) {
    const double st = out.xed(i, j, k, n), f = st * a.uflx(i, j, k) * out.ax;
    s(o + A_XE, 0, 0, 0) = st; s(o + A_FX, 0, 0, 0) = f;
    if (out.fx.ok()) out.fx(i, j, k, n) = f;
  }
  if (cx && cz) {
    const double st = out.yed(i, j, k, n), f = st * a.vflx(i, j, k) * out.ay;
    s(o + A_YE, 0, 0, 0) = st; s(o + A_FY, 0, 0, 0) = f;
    if (out.fy.ok()) out.fy(i, j, k, n) = f;
  }
  if (cx && cy) {
    const double st = out.zed(i, j, k, n), f = st * a.wflx(i, j, k) * out.az;
    s(o + A_ZE, 0, 0, 0) = st; s(o + A_FZ, 0, 0, 0) = f;
    if (out.fz.ok()) out.fz(i, j, k, n) = f;
  }
}

// ComputeDivergence(mult=-1, area-weighted) + div(umac) + ComputeConvectiveTerm + sign
__global__ void __launch_bounds__(TX* TY)
es_div_kernel(IX_KARG(EsArgs) a, IX_KARG(Scratch) sc, V4 aofs, double volinv, double dxi, double dyi, double dzi, int is_sync, Bx R) {
  GIDX(R)
  const SCur s = scur_at(sc, i, j, k);
  const int o = A_N * n;
  double upd = -volinv * ((s(o + A_FX, 1, 0, 0) - s(o + A_FX, 0, 0, 0)) + (s(o + A_FY, 0, 1, 0) - s(o + A_FY, 0, 0, 0)) +
                          (s(o + A_FZ, 0, 0, 1) - s(o + A_FZ, 0, 0, 0)));
  if (!a.iconserv[n] && !is_sync) {
    const Cur u = cur_at(a.umac, 0, i, j, k), v = cur_at(a.vmac, 0, i, j, k), w = cur_at(a.wmac, 0, i, j, k);
    const double divum = dxi * (u(1, 0, 0) - u(0, 0, 0)) + dyi * (v(0, 1, 0) - v(0, 0, 0)) + dzi * (w(0, 0, 1) - w(0, 0, 0));
    double qb = s(o + A_XE, 0, 0, 0) + s(o + A_XE, 1, 0, 0) + s(o + A_YE, 0, 0, 0) + s(o + A_YE, 0, 1, 0) +
                s(o + A_ZE, 0, 0, 0) + s(o + A_ZE, 0, 0, 1);
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
  int fit, ppm;
  double dt, dtdx, dtdy, dtdz;
  BcAll bc;
};
// scratch ids: advective velocities 0..2; transverse edges: XE_V,XE_W (x-faces, comps 1,2),
// YE_U,YE_W, ZE_U,ZE_V; corner: YZ_U, ZY_U (for umac), XZ_V, ZX_V (vmac), XY_W, YX_W (wmac)
enum { B_UAD = 0, B_VAD, B_WAD, B_XE_V, B_XE_W, B_YE_U, B_YE_W, B_ZE_U, B_ZE_V,
       B_YZ_U, B_ZY_U, B_XZ_V, B_ZX_V, B_XY_W, B_YX_W, B_N };

struct EvCur {
  Cur q[3], f[3];
  SCur s;
};
IX_D EvCur ev_cursors(const EvArgs& a, const Scratch& sc, int i, int j, int k) {
  EvCur c;
  for (int n = 0; n < 3; ++n) { c.q[n] = cur_at(a.vel, n, i, j, k); c.f[n] = cur_at(a.force, n, i, j, k); }
  c.s = scur_at(sc, i, j, k);
  return c;
}

// lo/hi of component n on the D-face, traced with the cell-centred velocity component D
template <int D, bool BC>
IX_D void ev_lohi(const EvArgs& a, const EvCur& c, int n, int ci, double dtdx, double& lo, double& hi) {
  const BcD b = BC ? a.bc.dir(n, D) : BcD{};
  trace<D>(c.q[n], along<D>(c.q[D], -1), c.q[D](0, 0, 0), dtdx, lo, hi, a.ppm, ci, BC ? &b : nullptr);
  if (a.fit && c.f[n].ok()) {
    lo += 0.5 * a.dt * along<D>(c.f[n], -1);
    hi += 0.5 * a.dt * c.f[n](0, 0, 0);
  }
  edge_bc<D, BC>(a, n, c.q[n], ci, lo, hi, n == D);
}

template <bool BC>
__global__ void __launch_bounds__(TX* TY) ev_edge_kernel(IX_KARG(EvArgs) a, IX_KARG(Scratch) sc, IX_KARG(Bx) R) {
  GIDX(R)
  (void)n;
  const EvCur c = ev_cursors(a, sc, i, j, k);
  const Bx& b = a.bx;
  const bool inx = i >= b.lo[0] - 1 && i <= b.hi[0] + 1, iny = j >= b.lo[1] - 1 && j <= b.hi[1] + 1,
             inz = k >= b.lo[2] - 1 && k <= b.hi[2] + 1;
  double lo, hi;
  if (i >= b.lo[0] && i <= b.hi[0] + 1 && iny && inz) {
    ev_lohi<0, BC>(a, c, 0, i, a.dtdx, lo, hi);
    const double uad = riemann(lo, hi);
    c.s(B_UAD, 0, 0, 0) = uad;
    ev_lohi<0, BC>(a, c, 1, i, a.dtdx, lo, hi);
    c.s(B_XE_V, 0, 0, 0) = upwind(lo, hi, uad);
    ev_lohi<0, BC>(a, c, 2, i, a.dtdx, lo, hi);
    c.s(B_XE_W, 0, 0, 0) = upwind(lo, hi, uad);
  }
  if (j >= b.lo[1] && j <= b.hi[1] + 1 && inx && inz) {
    ev_lohi<1, BC>(a, c, 1, j, a.dtdy, lo, hi);
    const double vad = riemann(lo, hi);
    c.s(B_VAD, 0, 0, 0) = vad;
    ev_lohi<1, BC>(a, c, 0, j, a.dtdy, lo, hi);
    c.s(B_YE_U, 0, 0, 0) = upwind(lo, hi, vad);
    ev_lohi<1, BC>(a, c, 2, j, a.dtdy, lo, hi);
    c.s(B_YE_W, 0, 0, 0) = upwind(lo, hi, vad);
  }
  if (k >= b.lo[2] && k <= b.hi[2] + 1 && inx && iny) {
    ev_lohi<2, BC>(a, c, 2, k, a.dtdz, lo, hi);
    const double wad = riemann(lo, hi);
    c.s(B_WAD, 0, 0, 0) = wad;
    ev_lohi<2, BC>(a, c, 0, k, a.dtdz, lo, hi);
    c.s(B_ZE_U, 0, 0, 0) = upwind(lo, hi, wad);
    ev_lohi<2, BC>(a, c, 1, k, a.dtdz, lo, hi);
    c.s(B_ZE_V, 0, 0, 0) = upwind(lo, hi, wad);
  }
}

template <bool BC>
__global__ void __launch_bounds__(TX* TY) ev_corner_kernel(IX_KARG(EvArgs) a, IX_KARG(Scratch) sc, IX_KARG(Bx) R) {
  GIDX(R)
  (void)n;
  const EvCur c = ev_cursors(a, sc, i, j, k);
  const Bx& b = a.bx;
  const SArr uad = sarr(c.s, B_UAD), vad = sarr(c.s, B_VAD), wad = sarr(c.s, B_WAD);
  const bool cx = i <= b.hi[0], cy = j <= b.hi[1], cz = k <= b.hi[2];
  const bool gx = i >= b.lo[0] - 1 && i <= b.hi[0] + 1, gy = j >= b.lo[1] - 1 && j <= b.hi[1] + 1,
             gz = k >= b.lo[2] - 1 && k <= b.hi[2] + 1;
  const bool fx = i >= b.lo[0] && i <= b.hi[0] + 1, fy = j >= b.lo[1] && j <= b.hi[1] + 1,
             fz = k >= b.lo[2] && k <= b.hi[2] + 1;
  double lo, hi, l1, h1;
  // x-face states: comp 2 coupled with y (for wmac), comp 1 coupled with z (for vmac)
  if (fx && fy && cy && gz) {
    ev_lohi<0, BC>(a, c, 2, i, a.dtdx, lo, hi);
    corner<0, 1>(l1, h1, lo, hi, c.q[2], vad, sarr(c.s, B_YE_W), a.dtdy / 3.0, false);
    edge_bc<0, BC>(a, 2, c.q[2], i, l1, h1, 2 == 0);
    c.s(B_XY_W, 0, 0, 0) = upwind(l1, h1, uad(0, 0, 0));
  }
  if (fx && gy && fz && cz) {
    ev_lohi<0, BC>(a, c, 1, i, a.dtdx, lo, hi);
    corner<0, 2>(l1, h1, lo, hi, c.q[1], wad, sarr(c.s, B_ZE_V), a.dtdz / 3.0, false);
    edge_bc<0, BC>(a, 1, c.q[1], i, l1, h1, 1 == 0);
    c.s(B_XZ_V, 0, 0, 0) = upwind(l1, h1, uad(0, 0, 0));
  }
  // y-face states: comp 2 coupled with x (for wmac), comp 0 coupled with z (for umac)
  if (fy && fx && cx && gz) {
    ev_lohi<1, BC>(a, c, 2, j, a.dtdy, lo, hi);
    corner<1, 0>(l1, h1, lo, hi, c.q[2], uad, sarr(c.s, B_XE_W), a.dtdx / 3.0, false);
    edge_bc<1, BC>(a, 2, c.q[2], j, l1, h1, 2 == 1);
    c.s(B_YX_W, 0, 0, 0) = upwind(l1, h1, vad(0, 0, 0));
  }
  if (fy && gx && fz && cz) {
    ev_lohi<1, BC>(a, c, 0, j, a.dtdy, lo, hi);
    corner<1, 2>(l1, h1, lo, hi, c.q[0], wad, sarr(c.s, B_ZE_U), a.dtdz / 3.0, false);
    edge_bc<1, BC>(a, 0, c.q[0], j, l1, h1, 0 == 1);
    c.s(B_YZ_U, 0, 0, 0) = upwind(l1, h1, vad(0, 0, 0));
  }
  // z-face states: comp 1 coupled with x (for vmac), comp 0 coupled with y (for umac)
  if (fz && fx && cx && gy) {
    ev_lohi<2, BC>(a, c, 1, k, a.dtdz, lo, hi);
    corner<2, 0>(l1, h1, lo, hi, c.q[1], uad, sarr(c.s, B_XE_V), a.dtdx / 3.0, false);
    edge_bc<2, BC>(a, 1, c.q[1], k, l1, h1, 1 == 2);
    c.s(B_ZX_V, 0, 0, 0) = upwind(l1, h1, wad(0, 0, 0));
  }
  if (fz && gx && fy && cy) {
    ev_lohi<2, BC>(a, c, 0, k, a.dtdz, lo, hi);
    corner<2, 1>(l1, h1, lo, hi, c.q[0], vad, sarr(c.s, B_YE_U), a.dtdy / 3.0, false);
    edge_bc<2, BC>(a, 0, c.q[0], k, l1, h1, 0 == 2);
    c.s(B_ZY_U, 0, 0, 0) = upwind(l1, h1, wad(0, 0, 0));
  }
}

template <int D, int D1, int D2, bool BC>
IX_D double ev_final(const EvArgs& a, const EvCur& c, int ci, int A1, int A2, int T1, int T2, double dtdx, double dtd1, double dtd2) {
  double stl, sth;
  ev_lohi<D, BC>(a, c, D, ci, dtdx, stl, sth);
  double fl = 0.0, fh = 0.0;
  if (!a.fit && c.f[D].ok()) { fl = 0.5 * a.dt * along<D>(c.f[D], -1); fh = 0.5 * a.dt * c.f[D](0, 0, 0); }
  transverse<D, D1, D2>(stl, sth, c.q[D], sarr(c.s, A1), sarr(c.s, A2), sarr(c.s, T1), sarr(c.s, T2), dtd1, dtd2, false);
  stl += fl; sth += fh;
  if (BC) {
    set_edge_bc(stl, sth, along<D>(c.q[D], -1), c.q[D](0, 0, 0), ci, a.bc.dir(D, D), true);
    outflow_bc(stl, sth, ci, a.bc.dir(D, D), true, true);
  }
  return riemann(stl, sth);
}

template <bool BC>
__global__ void __launch_bounds__(TX* TY) ev_final_kernel(IX_KARG(EvArgs) a, IX_KARG(Scratch) sc, V4 umac, V4 vmac, V4 wmac, IX_KARG(Bx) R) {
  GIDX(R)
  (void)n;
  const EvCur c = ev_cursors(a, sc, i, j, k);
  const Bx& b = a.bx;
  const bool cx = i <= b.hi[0], cy = j <= b.hi[1], cz = k <= b.hi[2];
  if (cy && cz) umac(i, j, k) = ev_final<0, 1, 2, BC>(a, c, i, B_VAD, B_WAD, B_YZ_U, B_ZY_U, a.dtdx, a.dtdy, a.dtdz);
  if (cx && cz) vmac(i, j, k) = ev_final<1, 0, 2, BC>(a, c, j, B_UAD, B_WAD, B_XZ_V, B_ZX_V, a.dtdy, a.dtdx, a.dtdz);
  if (cx && cy) wmac(i, j, k) = ev_final<2, 0, 1, BC>(a, c, k, B_UAD, B_VAD, B_XY_W, B_YX_W, a.dtdz, a.dtdx, a.dtdy);
}

#if !defined(IX_EMUL)
// ===========================================================================
// Fused ComputeAofs tile kernel
// ===========================================================================
// One CTA = one 8x8x8 tile of cells of ONE component; one thread = one "site" of the tile grown
// by one cell (10^3 = 1000 sites, 1024 threads).  Every intermediate of the Godunov pipeline lives
// in shared memory or registers; global memory sees q (tile grown by 3), the MAC velocities,
// force/divu once per site, and the aofs (and optional flux / edge-state) stores.
//
// The algorithm is the staged one above, regrouped per CELL so that every stage exchanges one
// value per site with its neighbours instead of re-deriving face quantities:
//   site c traces to its two faces per direction:  L_d(c) -> face c+e_d (its "lo" state),
//                                                  H_d(c) -> face c      (its "hi" state)
//   stage 2  e_d(face c)  = upwind(L_d(c-e_d), H_d(c), mac_d(c))
//   stage 3  T_d(c)       = e_d(c+e_d) mac_d(c+e_d) - e_d(c) mac_d(c) [- q (mac_d(c+e_d) - mac_d(c))]
//            (the transverse derivative the corner coupling applies to BOTH faces next to cell c)
//   stage 4  corner state d|t on face c = upwind(L_d(c-e_d) - dt/3dx_t T_t(c-e_d), H_d(c) - dt/3dx_t T_t(c), mac_d(c))
//   stage 5  W_d(c) = transverse + divu + forcing correction of cell c for its two d-faces
//   stage 6  final state on face c = upwind(L_d(c-e_d) - W_d(c-e_d), H_d(c) - W_d(c), mac_d(c))
//   stage 7  aofs(c) from the six final face states
// This is algebraically the staged pipeline; the grouping of the corner/transverse terms per cell
// changes rounding in the last bits only.  Each 4th-order slope is evaluated once per site
// (x1.95 halo redundancy for an 8^3 tile) instead of 18 times per cell in the staged kernels.
namespace tile {
constexpr int TB = 8, G = TB + 2, GG = G * G, NS = G * G * G;
constexpr int QE = TB + 6, QQ = QE * QE, NQ = QE * QE * QE;
constexpr int NT = 1024;
constexpr int PAD = 128;  // neighbour reads of unused sites may run up to GG past an array: keep them inside the allocation
constexpr int SMEM_BYTES = (PAD + NQ + 12 * NS + PAD) * (int)sizeof(double);

// upwind() as a select: identical value for finite inputs (fu is exactly 0 or 1 there), fewer DP instructions
IX_D double upsel(double lo, double hi, double vel) {
  return (fabs(vel) < upopts().small_vel) ? 0.5 * (hi + lo) : ((vel >= 0.0) ? lo : hi);
}
// 32-bit element offset of (i,j,k) in a view (host checks that every fab has < 2^31 elements)
template <class V> IX_D int off32(const V& v, int i, int j, int k) {
  return (i - v.l0) + (j - v.l1) * (int)v.js + (k - v.l2) * (int)v.ks;
}
IX_D void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

// Stages 2-5 run WITHOUT per-site predicates: a site that does not need a quantity computes it anyway
// from whatever its neighbours hold (possibly garbage, always inside the shared allocation).  Values that
// reach a store are only ever derived from needed, fully defined quantities (the dependence cone of the
// tile interior lies inside the site box), so the extra lanes cost nothing and the branches disappear.
template <bool SAMEFLUX>   // the flux velocities are the MAC velocities (everything but the sync call)
__global__ void __launch_bounds__(NT, 1)
aofs_tile_kernel(IX_KARG(EsArgs) a, IX_KARG(EsOut) out, V4 aofs, double volinv, double dxi, double dyi, double dzi,
                 int is_sync, int ncomp) {
  extern __shared__ double sm_raw[];
  double* const Q = sm_raw + PAD;
  double* const AL = Q + NQ;        // L_x, L_y, L_z
  double* const AE = AL + 3 * NS;   // edge states; then corner xy, xz, yx; then final states
  double* const AT = AE + 3 * NS;   // T_x, T_y, T_z; then final lo states
  double* const AC = AT + 3 * NS;   // corner yz, zx, zy
  const int tid = threadIdx.x;
  const int l0 = a.bx.lo[0] + TB * (int)blockIdx.x, l1 = a.bx.lo[1] + TB * (int)blockIdx.y, l2 = a.bx.lo[2] + TB * (int)blockIdx.z;
  // stage 0: q of one component on the tile grown by 3, asynchronously (16 lanes per row of 14, 64 rows per pass)
  const int sjs = (int)a.S.js, sks = (int)a.S.ks;
  const double* Sp0 = a.S.p + off32(a.S, l0 - 3, l1 - 3, l2 - 3);
  auto stage_q = [&](int n) {
    const double* Sp = Sp0 + n * a.S.ns;
    const int x = tid & 15, r0 = tid >> 4;
    if (x < QE) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int r = r0 + 64 * m;
        if (r < QQ) {
          const int z = r / QE, y = r - z * QE;
          cp_async8(&Q[x + r * QE], Sp + x + y * sjs + z * sks);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_q(0);
  const bool act = tid < NS;
  const int t = act ? tid : 0;
  const int sk = t / GG, sj = (t - sk * GG) / G, si = t - sk * GG - sj * G;
  const int i = l0 - 1 + si, j = l1 - 1 + sj, k = l2 - 1 + sk;
  const bool fx = act && si >= 1, fy = act && sj >= 1, fz = act && sk >= 1;          // low face is a face of the tile
  const bool inx = fx && si <= TB, iny = fy && sj <= TB, inz = fz && sk <= TB;      // cell index inside the tile
  const bool hasf = a.force.ok();
  // global inputs of this site that all components share (overlap the q staging): MAC and flux velocities, divu
  const double* pu = a.umac.p + off32(a.umac, i, j, k);
  const double* pv = a.vmac.p + off32(a.vmac, i, j, k);
  const double* pw = a.wmac.p + off32(a.wmac, i, j, k);
  const double um = pu[0], up = pu[1], vm = pv[0], vp = pv[(int)a.vmac.js], wm = pw[0], wp = pw[(int)a.wmac.ks];
  double ufm_ = 0, ufp_ = 0, vfm_ = 0, vfp_ = 0, wfm_ = 0, wfp_ = 0;
  if (!SAMEFLUX) {
    const double* qu = a.uflx.p + off32(a.uflx, i, j, k);
    const double* qv = a.vflx.p + off32(a.vflx, i, j, k);
    const double* qw = a.wflx.p + off32(a.wflx, i, j, k);
    ufm_ = qu[0]; ufp_ = qu[1]; vfm_ = qv[0]; vfp_ = qv[(int)a.vflx.js]; wfm_ = qw[0]; wfp_ = qw[(int)a.wflx.ks];
  }
  const double ufm = SAMEFLUX ? um : ufm_, ufp = SAMEFLUX ? up : ufp_, vfm = SAMEFLUX ? vm : vfm_, vfp = SAMEFLUX ? vp : vfp_,
               wfm = SAMEFLUX ? wm : wfm_, wfp = SAMEFLUX ? wp : wfp_;
  const double dv0 = a.divu.ok() ? a.divu.p[off32(a.divu, i, j, k)] : 0.0;
  const double* pf = hasf ? a.force.p + off32(a.force, i, j, k) : nullptr;
  double fv_next = hasf ? pf[0] : 0.0;
  // The CTA works through the components of its tile one after the other: the velocities above are loaded once, and
  // the next component's q is staged (cp.async) underneath stages 2-7 of the current one (q is dead after stage 1).
  for (int n = 0; n < ncomp; ++n) {
  const bool cs = a.iconserv[n] != 0;
  const double fv = fv_next;
  const double dv = cs ? dv0 : 0.0;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  double q0, Hx, Hy, Hz;
  {  // stage 1: slopes and the six traced states of this cell
    const int qi = (si + 2) + (sj + 2) * QE + (sk + 2) * QQ;
    q0 = Q[qi];
    const double sx = slope4_vals(Q[qi - 2], Q[qi - 1], q0, Q[qi + 1], Q[qi + 2]);
    const double sy = slope4_vals(Q[qi - 2 * QE], Q[qi - QE], q0, Q[qi + QE], Q[qi + 2 * QE]);
    const double sz = slope4_vals(Q[qi - 2 * QQ], Q[qi - QQ], q0, Q[qi + QQ], Q[qi + 2 * QQ]);
    double Lx = q0 + 0.5 * (1.0 - up * a.dtdx) * sx, Ly = q0 + 0.5 * (1.0 - vp * a.dtdy) * sy, Lz = q0 + 0.5 * (1.0 - wp * a.dtdz) * sz;
    Hx = q0 + 0.5 * (-1.0 - um * a.dtdx) * sx; Hy = q0 + 0.5 * (-1.0 - vm * a.dtdy) * sy; Hz = q0 + 0.5 * (-1.0 - wm * a.dtdz) * sz;
    if (a.fit && hasf) {
      const double h = 0.5 * a.dt * fv;
      Lx += h; Ly += h; Lz += h; Hx += h; Hy += h; Hz += h;
    }
    AL[t] = Lx; AL[NS + t] = Ly; AL[2 * NS + t] = Lz;
  }
  __syncthreads();
  if (n + 1 < ncomp) {   // q is consumed: stage the next component into the same buffer, fetch its force
    stage_q(n + 1);
    if (hasf) fv_next = pf[(n + 1) * a.force.ns];
  }
  // stage 2: upwinded edge states on the low faces
  const double lox = AL[t - 1], loy = AL[NS + t - G], loz = AL[2 * NS + t - GG];
  const double xe = upsel(lox, Hx, um), ye = upsel(loy, Hy, vm), ze = upsel(loz, Hz, wm);
  AE[t] = xe; AE[NS + t] = ye; AE[2 * NS + t] = ze;
  __syncthreads();
  {  // stage 3: transverse derivative terms of this cell
    double Tx = AE[t + 1] * up - xe * um, Ty = AE[NS + t + G] * vp - ye * vm, Tz = AE[2 * NS + t + GG] * wp - ze * wm;
    if (!cs) {
      if (upopts().corner_adv) { Tx = 0.5 * (up + um) * (AE[t + 1] - xe); Ty = 0.5 * (vp + vm) * (AE[NS + t + G] - ye); Tz = 0.5 * (wp + wm) * (AE[2 * NS + t + GG] - ze); }
      else { Tx -= q0 * (up - um); Ty -= q0 * (vp - vm); Tz -= q0 * (wp - wm); }
    }
    AT[t] = Tx; AT[NS + t] = Ty; AT[2 * NS + t] = Tz;
    __syncthreads();
    // stage 4: corner-coupled states on the low faces (AE is free: its last readers are behind the barrier)
    const double d3x = a.dtdx / 3.0, d3y = a.dtdy / 3.0, d3z = a.dtdz / 3.0;
    AE[t] = upsel(lox - d3y * AT[NS + t - 1], Hx - d3y * Ty, um);                  // xy
    AE[NS + t] = upsel(lox - d3z * AT[2 * NS + t - 1], Hx - d3z * Tz, um);         // xz
    AE[2 * NS + t] = upsel(loy - d3x * AT[t - G], Hy - d3x * Tx, vm);             // yx
    AC[t] = upsel(loy - d3z * AT[2 * NS + t - G], Hy - d3z * Tz, vm);             // yz
    AC[NS + t] = upsel(loz - d3x * AT[t - GG], Hz - d3x * Tx, wm);                // zx
    AC[2 * NS + t] = upsel(loz - d3y * AT[NS + t - GG], Hz - d3y * Ty, wm);       // zy
  }
  __syncthreads();
  {  // stage 5: transverse / divu / forcing correction of this cell, per direction
    double base = 0.5 * a.dt * q0 * dv;  // es_finish terms, common to the three directions (dv = 0 unless conservative)
    if (!a.fit && hasf) base -= 0.5 * a.dt * fv;
    const double* XY = AE; const double* XZ = AE + NS; const double* YX = AE + 2 * NS;
    const double* YZ = AC; const double* ZX = AC + NS; const double* ZY = AC + 2 * NS;
    double Wx, Wy, Wz;
    if (cs) {
      Wx = base + (0.5 * a.dtdy) * (YZ[t + G] * vp - YZ[t] * vm) + (0.5 * a.dtdz) * (ZY[t + GG] * wp - ZY[t] * wm)
           - (0.5 * a.dtdy) * q0 * (vp - vm) - (0.5 * a.dtdz) * q0 * (wp - wm);
      Wy = base + (0.5 * a.dtdx) * (XZ[t + 1] * up - XZ[t] * um) + (0.5 * a.dtdz) * (ZX[t + GG] * wp - ZX[t] * wm)
           - (0.5 * a.dtdx) * q0 * (up - um) - (0.5 * a.dtdz) * q0 * (wp - wm);
      Wz = base + (0.5 * a.dtdx) * (XY[t + 1] * up - XY[t] * um) + (0.5 * a.dtdy) * (YX[t + G] * vp - YX[t] * vm)
           - (0.5 * a.dtdx) * q0 * (up - um) - (0.5 * a.dtdy) * q0 * (vp - vm);
    } else {
      Wx = base + (0.25 * a.dtdy) * (vp + vm) * (YZ[t + G] - YZ[t]) + (0.25 * a.dtdz) * (wp + wm) * (ZY[t + GG] - ZY[t]);
      Wy = base + (0.25 * a.dtdx) * (up + um) * (XZ[t + 1] - XZ[t]) + (0.25 * a.dtdz) * (wp + wm) * (ZX[t + GG] - ZX[t]);
      Wz = base + (0.25 * a.dtdx) * (up + um) * (XY[t + 1] - XY[t]) + (0.25 * a.dtdy) * (vp + vm) * (YX[t + G] - YX[t]);
    }
    AT[t] = AL[t] - Wx; AT[NS + t] = AL[NS + t] - Wy; AT[2 * NS + t] = AL[2 * NS + t] - Wz;
    Hx -= Wx; Hy -= Wy; Hz -= Wz;
  }
  __syncthreads();
  // stage 6: final states on the low faces (AE is free again: the corner arrays were consumed in stage 5)
  const double xs = upsel(AT[t - 1], Hx, um), ys = upsel(AT[NS + t - G], Hy, vm), zs = upsel(AT[2 * NS + t - GG], Hz, wm);
  AE[t] = xs; AE[NS + t] = ys; AE[2 * NS + t] = zs;
  if (out.xed.ok()) {  // optional outputs: every face is written by exactly one tile
    if (fx && iny && inz && (si <= TB || i == a.bx.hi[0] + 1)) { out.xed(i, j, k, n) = xs; out.fx(i, j, k, n) = xs * ufm * out.ax; }
    if (fy && inx && inz && (sj <= TB || j == a.bx.hi[1] + 1)) { out.yed(i, j, k, n) = ys; out.fy(i, j, k, n) = ys * vfm * out.ay; }
    if (fz && inx && iny && (sk <= TB || k == a.bx.hi[2] + 1)) { out.zed(i, j, k, n) = zs; out.fz(i, j, k, n) = zs * wfm * out.az; }
  }
  __syncthreads();
  // stage 7: ComputeDivergence(mult = -1) + ComputeConvectiveTerm + sign
  if (inx && iny && inz) {
    const double xp = AE[t + 1], yp = AE[NS + t + G], zp = AE[2 * NS + t + GG];
    double upd = -volinv * ((xp * ufp * out.ax - xs * ufm * out.ax) + (yp * vfp * out.ay - ys * vfm * out.ay) +
                            (zp * wfp * out.az - zs * wfm * out.az));
    if (!cs && !is_sync) {
      const double divum = dxi * (up - um) + dyi * (vp - vm) + dzi * (wp - wm);
      double qb = xs + xp + ys + yp + zs + zp;
      qb /= 6.0;
      upd += qb * divum;
    }
    double* pa = aofs.p + n * aofs.ns + off32(aofs, i, j, k);
    if (is_sync) *pa -= upd;
    else *pa = -upd;
  }
  }  // components
}

// ===========================================================================
// Fused ExtrapVelToFaces tile kernel
// ===========================================================================
// The same 8^3 tile / 10^3 site organisation as aofs_tile_kernel, for the whole velocity at once: the advective velocities
// (u_ad, v_ad, w_ad = Riemann states of the NORMAL traces) upwind the transverse traces of the other components, so the three
// components share one CTA.  Per site: 9 limited slopes (3 components x 3 directions, each evaluated once), the 9 high-face traces
// L in shared memory and the 9 low-face traces H in registers; then per site the advective velocities and the 6 transverse edge
// states of its low faces, the 6 transverse-derivative terms, the 6 corner-coupled states, the three normal corrections and the
// three final Riemann states.  Global memory sees the velocity tile grown by 3 (three cp.async stages through one buffer), the
// forcing, and the three face stores: the staged kernels move 5.7x the algorithmic bytes through 15 scratch arrays.
// Interior boxes (no physical boundary within reach) and PLM only; everything else takes the staged kernels.
constexpr int EV_SMEM_BYTES = (PAD + NQ + 18 * NS + PAD) * (int)sizeof(double);
IX_D double riem(double lo, double hi) {
  const double st = ((lo + hi) >= 0.0) ? lo : hi;
  const bool ltm = ((lo <= 0.0 && hi >= 0.0) || (fabs(lo + hi) < upopts().small_vel));
  return ltm ? 0.0 : st;
}
__global__ void __launch_bounds__(NT, 1)
ev_tile_kernel(IX_KARG(EvArgs) a, V4 umac, V4 vmac, V4 wmac) {
  extern __shared__ double sm_raw[];
  double* const Q = sm_raw + PAD;
  double* const AL = Q + NQ;        // L[n][d] at (3 n + d) NS; in stage 5 the slots (n, n) take the corrected normal states
  double* const AB = AL + 9 * NS;   // u_ad, v_ad, w_ad; then 6 transverse edge states -> 6 T terms -> 6 corner states
  const int tid = threadIdx.x;
  const int l0 = a.bx.lo[0] + TB * (int)blockIdx.x, l1 = a.bx.lo[1] + TB * (int)blockIdx.y, l2 = a.bx.lo[2] + TB * (int)blockIdx.z;
  const int sjs = (int)a.vel.js, sks = (int)a.vel.ks;
  const double* Sp0 = a.vel.p + off32(a.vel, l0 - 3, l1 - 3, l2 - 3);
  auto stage_q = [&](int n) {
    const double* Sp = Sp0 + n * a.vel.ns;
    const int x = tid & 15, r0 = tid >> 4;
    if (x < QE) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int r = r0 + 64 * m;
        if (r < QQ) {
          const int z = r / QE, y = r - z * QE;
          cp_async8(&Q[x + r * QE], Sp + x + y * sjs + z * sks);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_q(0);
  const bool act = tid < NS;
  const int t = act ? tid : 0;
  const int sk = t / GG, sj = (t - sk * GG) / G, si = t - sk * GG - sj * G;
  const int i = l0 - 1 + si, j = l1 - 1 + sj, k = l2 - 1 + sk;
  const bool fx = act && si >= 1, fy = act && sj >= 1, fz = act && sk >= 1;          // low face is a face of the tile
  const bool inx = fx && si <= TB, iny = fy && sj <= TB, inz = fz && sk <= TB;      // cell index inside the tile
  const bool hasf = a.force.ok();
  // the cell's own velocity (the trace speeds) and forcing
  const double* pv0 = a.vel.p + off32(a.vel, i, j, k);
  const double cvel[3] = {pv0[0], pv0[a.vel.ns], pv0[2 * a.vel.ns]};
  double frc[3] = {0.0, 0.0, 0.0};
  if (hasf) { const double* pf = a.force.p + off32(a.force, i, j, k); frc[0] = pf[0]; frc[1] = pf[a.force.ns]; frc[2] = pf[2 * a.force.ns]; }
  const double dtd[3] = {a.dtdx, a.dtdy, a.dtdz};
  double H[3][3];   // [component][direction]
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    {  // stage 1: slopes and the six traced states of component n
      const int qi = (si + 2) + (sj + 2) * QE + (sk + 2) * QQ;
      const double q0 = Q[qi];
      const double sl[3] = {slope4_vals(Q[qi - 2], Q[qi - 1], q0, Q[qi + 1], Q[qi + 2]),
                            slope4_vals(Q[qi - 2 * QE], Q[qi - QE], q0, Q[qi + QE], Q[qi + 2 * QE]),
                            slope4_vals(Q[qi - 2 * QQ], Q[qi - QQ], q0, Q[qi + QQ], Q[qi + 2 * QQ])};
      const double h = (a.fit && hasf) ? 0.5 * a.dt * frc[n] : 0.0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double L = q0 + 0.5 * (1.0 - cvel[d] * dtd[d]) * sl[d];
        double Hh = q0 + 0.5 * (-1.0 - cvel[d] * dtd[d]) * sl[d];
        if (a.fit && hasf) { L += h; Hh += h; }
        AL[(3 * n + d) * NS + t] = L;
        H[n][d] = Hh;
      }
    }
    __syncthreads();
    if (n + 1 < 3) stage_q(n + 1);
  }
  // stage 2: advective velocities and transverse edge states on the low faces of this site
  const int off[3] = {1, G, GG};
#define EV_LO(n, d) AL[(3 * (n) + (d)) * NS + t - off[d]]
  const double uad = riem(EV_LO(0, 0), H[0][0]), vad = riem(EV_LO(1, 1), H[1][1]), wad = riem(EV_LO(2, 2), H[2][2]);
  {
    const double xe_v = upsel(EV_LO(1, 0), H[1][0], uad), xe_w = upsel(EV_LO(2, 0), H[2][0], uad);
    const double ye_u = upsel(EV_LO(0, 1), H[0][1], vad), ye_w = upsel(EV_LO(2, 1), H[2][1], vad);
    const double ze_u = upsel(EV_LO(0, 2), H[0][2], wad), ze_v = upsel(EV_LO(1, 2), H[1][2], wad);
    AB[t] = uad; AB[NS + t] = vad; AB[2 * NS + t] = wad;
    AB[3 * NS + t] = xe_v; AB[4 * NS + t] = xe_w; AB[5 * NS + t] = ye_u; AB[6 * NS + t] = ye_w; AB[7 * NS + t] = ze_u; AB[8 * NS + t] = ze_v;
    __syncthreads();
    // stage 3: transverse-derivative terms of this CELL (non-conservative form; godunov.cu corner())
    const double uadp = AB[t + 1], vadp = AB[NS + t + G], wadp = AB[2 * NS + t + GG];
    auto T = [&](double ep, double e, double adp, double ad, double q) {
      return upopts().corner_adv ? 0.5 * (adp + ad) * (ep - e) : (ep * adp - e * ad) - q * (adp - ad);
    };
    const double Tx1 = T(AB[3 * NS + t + 1], xe_v, uadp, uad, cvel[1]), Tx2 = T(AB[4 * NS + t + 1], xe_w, uadp, uad, cvel[2]);
    const double Ty0 = T(AB[5 * NS + t + G], ye_u, vadp, vad, cvel[0]), Ty2 = T(AB[6 * NS + t + G], ye_w, vadp, vad, cvel[2]);
    const double Tz0 = T(AB[7 * NS + t + GG], ze_u, wadp, wad, cvel[0]), Tz1 = T(AB[8 * NS + t + GG], ze_v, wadp, wad, cvel[1]);
    __syncthreads();
    AB[3 * NS + t] = Tx1; AB[4 * NS + t] = Tx2; AB[5 * NS + t] = Ty0; AB[6 * NS + t] = Ty2; AB[7 * NS + t] = Tz0; AB[8 * NS + t] = Tz1;
    __syncthreads();
    // stage 4: corner-coupled states on the low faces (for u_mac: u on y / z faces; v_mac: v on x / z; w_mac: w on x / y)
    const double d3x = a.dtdx / 3.0, d3y = a.dtdy / 3.0, d3z = a.dtdz / 3.0;
    const double yz_u = upsel(EV_LO(0, 1) - d3z * AB[7 * NS + t - G], H[0][1] - d3z * Tz0, vad);
    const double zy_u = upsel(EV_LO(0, 2) - d3y * AB[5 * NS + t - GG], H[0][2] - d3y * Ty0, wad);
    const double xz_v = upsel(EV_LO(1, 0) - d3z * AB[8 * NS + t - 1], H[1][0] - d3z * Tz1, uad);
    const double zx_v = upsel(EV_LO(1, 2) - d3x * AB[3 * NS + t - GG], H[1][2] - d3x * Tx1, wad);
    const double xy_w = upsel(EV_LO(2, 0) - d3y * AB[6 * NS + t - 1], H[2][0] - d3y * Ty2, uad);
    const double yx_w = upsel(EV_LO(2, 1) - d3x * AB[4 * NS + t - G], H[2][1] - d3x * Tx2, vad);
    __syncthreads();
    AB[3 * NS + t] = yz_u; AB[4 * NS + t] = zy_u; AB[5 * NS + t] = xz_v; AB[6 * NS + t] = zx_v; AB[7 * NS + t] = xy_w; AB[8 * NS + t] = yx_w;
    __syncthreads();
    // stage 5: transverse (+ forcing) correction of the normal states of this cell
    const double fb[3] = {(!a.fit && hasf) ? -0.5 * a.dt * frc[0] : 0.0, (!a.fit && hasf) ? -0.5 * a.dt * frc[1] : 0.0,
                          (!a.fit && hasf) ? -0.5 * a.dt * frc[2] : 0.0};
    const double Wx = fb[0] + (0.25 * a.dtdy) * (vadp + vad) * (AB[3 * NS + t + G] - yz_u) + (0.25 * a.dtdz) * (wadp + wad) * (AB[4 * NS + t + GG] - zy_u);
    const double Wy = fb[1] + (0.25 * a.dtdx) * (uadp + uad) * (AB[5 * NS + t + 1] - xz_v) + (0.25 * a.dtdz) * (wadp + wad) * (AB[6 * NS + t + GG] - zx_v);
    const double Wz = fb[2] + (0.25 * a.dtdx) * (uadp + uad) * (AB[7 * NS + t + 1] - xy_w) + (0.25 * a.dtdy) * (vadp + vad) * (AB[8 * NS + t + G] - yx_w);
    // (each thread replaces ITS OWN normal L; the neighbours' reads of the old values are all behind the barriers above)
    AL[0 * NS + t] -= Wx; AL[4 * NS + t] -= Wy; AL[8 * NS + t] -= Wz;
    H[0][0] -= Wx; H[1][1] -= Wy; H[2][2] -= Wz;
  }
  __syncthreads();
  // stage 6: final Riemann states = the face velocities; every face is written by exactly one tile
  if (fx && iny && inz && (si <= TB || i == a.bx.hi[0] + 1)) umac.p[off32(umac, i, j, k)] = riem(AL[0 * NS + t - 1], H[0][0]);
  if (fy && inx && inz && (sj <= TB || j == a.bx.hi[1] + 1)) vmac.p[off32(vmac, i, j, k)] = riem(AL[4 * NS + t - G], H[1][1]);
  if (fz && inx && iny && (sk <= TB || k == a.bx.hi[2] + 1)) wmac.p[off32(wmac, i, j, k)] = riem(AL[8 * NS + t - GG], H[2][2]);
#undef EV_LO
}

inline bool fits32(const C4& v, const Bx& bx, int ng) {  // offsets of grow(bx, ng+1) fit in 32 bits
  return !v.p || ((int64_t)(bx.nz() + 2 * ng + 2) * v.ks < (int64_t)1 << 31);
}
inline bool aofs_tile_ok(const Bx& bx, const AofsArgs& a) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("IAMRX_ADV_TILE"); on = (e && e[0] == '0') ? 0 : 1; }
  if (!on || bx.nx() % TB || bx.ny() % TB || bx.nz() % TB) return false;
  if (a.write_fluxes && !(a.fx.ok() && a.fy.ok() && a.fz.ok() && a.xed.ok() && a.yed.ok() && a.zed.ok())) return false;
  return fits32(a.S, bx, 3) && fits32(a.force, bx, 1) && fits32(a.divu, bx, 1) && fits32(a.umac, bx, 1) && fits32(a.vmac, bx, 1) &&
         fits32(a.wmac, bx, 1) && fits32(a.uflx, bx, 1) && fits32(a.vflx, bx, 1) && fits32(a.wflx, bx, 1) &&
         fits32(C4{a.aofs.p, 0, 0, 0, a.aofs.js, a.aofs.ks, a.aofs.ns}, bx, 0);
}
}  // namespace tile
#endif

// AdvBC -> device form; `any` only if the stencil of `bx` (slopes reach the second cell, faces of the box grown by 1)
// can see a non-interior boundary
inline BcAll to_bcall(const AdvBC* bc, const Bx& bx, int ncomp) {
  BcAll b{};
  if (!bc) return b;
  for (int d = 0; d < 3; ++d) { b.dlo[d] = bc->dlo[d]; b.dhi[d] = bc->dhi[d]; }
  for (int n = 0; n < 8; ++n)
    for (int d = 0; d < 3; ++d) {
      b.lo[n][d] = bc->lo[n][d]; b.hi[n][d] = bc->hi[n][d];
      if (n < ncomp && ((b.lo[n][d] != IAMRX_BC_INT_DIR && bx.lo[d] - 4 <= b.dlo[d]) || (b.hi[n][d] != IAMRX_BC_INT_DIR && bx.hi[d] + 4 >= b.dhi[d])))
        b.any = 1;
    }
  return b;
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
#if !defined(IX_EMUL)
  // A box that touches non-interior domain sides: the 8-cell slabs next to those sides go through the staged kernels (which
  // carry the boundary conditions), the rest of the box keeps the fused tile kernel.
  if (!a.staged && !a.ppm && !a.known_edge_state && !a.bc.interior() && bx.nx() % tile::TB == 0 && bx.ny() % tile::TB == 0 && bx.nz() % tile::TB == 0) {
    Bx rest = bx;
    AofsArgs slab = a; slab.staged = 1;
    for (int d = 0; d < 3; ++d) {
      bool lo_bc = false, hi_bc = false;
      for (int n = 0; n < a.ncomp; ++n) { lo_bc |= a.bc.lo[n][d] != IAMRX_BC_INT_DIR; hi_bc |= a.bc.hi[n][d] != IAMRX_BC_INT_DIR; }
      if (lo_bc && rest.lo[d] - 4 <= a.bc.dlo[d] && rest.hi[d] - rest.lo[d] + 1 > tile::TB) {
        Bx p = rest; p.hi[d] = p.lo[d] + tile::TB - 1; rest.lo[d] += tile::TB;
        const int rc = compute_aofs(p, slab, g, s);
        if (rc) return rc;
      }
      if (hi_bc && rest.hi[d] + 4 >= a.bc.dhi[d] && rest.hi[d] - rest.lo[d] + 1 > tile::TB) {
        Bx p = rest; p.lo[d] = p.hi[d] - tile::TB + 1; rest.hi[d] -= tile::TB;
        const int rc = compute_aofs(p, slab, g, s);
        if (rc) return rc;
      }
    }
    if (rest.lo[0] != bx.lo[0] || rest.hi[0] != bx.hi[0] || rest.lo[1] != bx.lo[1] || rest.hi[1] != bx.hi[1] || rest.lo[2] != bx.lo[2] || rest.hi[2] != bx.hi[2]) {
      AofsArgs in = a;
      if (to_bcall(&a.bc, rest, a.ncomp).any) in.staged = 1;   // still next to a boundary (thin box): staged as well
      else in.bc = AdvBC{};                                      // interior stencil: no boundary logic
      return compute_aofs(rest, in, g, s);
    }
  }
#endif
  ProfScope prof_(IAMRX_PROF_AOFS, bx.npts(), (double)bx.npts() * (24.0 * a.ncomp + 32.0), s);
  EsArgs e{};
  e.bx = bx;
  e.S = a.S; e.force = a.force; e.divu = a.divu;
  e.umac = a.umac; e.vmac = a.vmac; e.wmac = a.wmac;
  e.uflx = a.uflx; e.vflx = a.vflx; e.wflx = a.wflx;
  for (int n = 0; n < 8; ++n) e.iconserv[n] = (n < a.ncomp) ? a.iconserv[n] : 0;
  e.fit = a.forces_in_trans;
  e.ppm = a.ppm;
  e.is_velocity = a.is_velocity;
  e.bc = to_bcall(&a.bc, bx, a.ncomp);
  e.dt = g.dt; e.dtdx = g.dt / g.dx[0]; e.dtdy = g.dt / g.dx[1]; e.dtdz = g.dt / g.dx[2];
  EsOut out{};
  if (a.write_fluxes || a.known_edge_state) { out.fx = a.fx; out.fy = a.fy; out.fz = a.fz; out.xed = a.xed; out.yed = a.yed; out.zed = a.zed; }
  out.ax = g.dx[1] * g.dx[2]; out.ay = g.dx[0] * g.dx[2]; out.az = g.dx[0] * g.dx[1];
#if !defined(IX_EMUL)
  if (!a.staged && !a.ppm && !a.known_edge_state && !e.bc.any && tile::aofs_tile_ok(bx, a)) {   // the tile kernel: PLM, interior stencil
    static bool attr_set = false;
    if (!attr_set) {
      IX_CUDA(cudaFuncSetAttribute(tile::aofs_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile::SMEM_BYTES));
      IX_CUDA(cudaFuncSetAttribute(tile::aofs_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile::SMEM_BYTES));
      attr_set = true;
    }
    const dim3 grd(bx.nx() / tile::TB, bx.ny() / tile::TB, bx.nz() / tile::TB);
    const bool same = (a.uflx.p == a.umac.p) && (a.vflx.p == a.vmac.p) && (a.wflx.p == a.wmac.p);
    if (same) IX_LAUNCH(tile::aofs_tile_kernel<true>, grd, dim3(tile::NT, 1, 1), tile::SMEM_BYTES, s, e, out, a.aofs,
                        1.0 / (g.dx[0] * g.dx[1] * g.dx[2]), 1.0 / g.dx[0], 1.0 / g.dx[1], 1.0 / g.dx[2], a.is_sync, a.ncomp);
    else IX_LAUNCH(tile::aofs_tile_kernel<false>, grd, dim3(tile::NT, 1, 1), tile::SMEM_BYTES, s, e, out, a.aofs,
                   1.0 / (g.dx[0] * g.dx[1] * g.dx[2]), 1.0 / g.dx[0], 1.0 / g.dx[1], 1.0 / g.dx[2], a.is_sync, a.ncomp);
    return check_launch("aofs_tile");
  }
#endif
  ScratchOwner so;
  if (so.init(bx, A_N * a.ncomp) != IAMRX_OK) return IAMRX_ERR_CUDA;
  Bx R1 = grow(bx, 1); R1.hi[0]++; R1.hi[1]++; R1.hi[2]++;
  Bx R2 = bx; R2.hi[0]++; R2.hi[1]++; R2.hi[2]++;
  int rc = IAMRX_OK;
  if (a.known_edge_state) {
    IX_LAUNCH(es_known_kernel, grid_for(R2, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, out, R2);
    rc = check_launch("es_known");
    if (rc) return rc;
  } else {
    if (e.bc.any) IX_LAUNCH(es_edge_kernel<true>, grid_for(R1, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
    else IX_LAUNCH(es_edge_kernel<false>, grid_for(R1, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
    rc = check_launch("es_edge");
    if (rc) return rc;
    if (e.bc.any) IX_LAUNCH(es_corner_kernel<true>, grid_for(R1, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
    else IX_LAUNCH(es_corner_kernel<false>, grid_for(R1, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
    rc = check_launch("es_corner");
    if (rc) return rc;
    if (e.bc.any) IX_LAUNCH(es_final_kernel<true>, grid_for(R2, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, out, R2);
    else IX_LAUNCH(es_final_kernel<false>, grid_for(R2, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, out, R2);
    rc = check_launch("es_final");
    if (rc) return rc;
  }
  IX_LAUNCH(es_div_kernel, grid_for(bx, a.ncomp), dim3(TX, TY, 1), 0, s, e, so.sc, a.aofs,
            1.0 / (g.dx[0] * g.dx[1] * g.dx[2]), 1.0 / g.dx[0], 1.0 / g.dx[1], 1.0 / g.dx[2], a.is_sync, bx);
  return check_launch("es_div");
}

// UNVERIFIED-UPSTREAM switches (include/iamrx.h iamrx_set_option): host copy + the constant-memory copy the kernels read
int extrap_vel_to_faces_staged(const Bx& bx, C4 vel, C4 force, V4 umac, V4 vmac, V4 wmac, const AdvGeom& g,
                               int forces_in_trans, cudaStream_t s, int ppm, const AdvBC* bc);

int godunov_set_option(int opt, double value) {
  UpOpts o = g_upopts_host;
  switch (opt) {
    case IAMRX_OPT_SMALL_VEL: if (!(value >= 0.0)) return IAMRX_ERR_ARG; o.small_vel = value; break;
    case IAMRX_OPT_SLOPE_ORDER: if (value != 2.0 && value != 4.0) return IAMRX_ERR_ARG; o.slope_order = (int)value; break;
    case IAMRX_OPT_CORNER_FORM: if (value != 0.0 && value != 1.0) return IAMRX_ERR_ARG; o.corner_adv = (int)value; break;
    case IAMRX_OPT_EXTDIR_BOTH: if (value != 0.0 && value != 1.0) return IAMRX_ERR_ARG; o.extdir_both = (int)value; break;
    default: return IAMRX_ERR_ARG;
  }
  g_upopts_host = o;
#if !defined(IX_EMUL)
  if (device_ok()) IX_CUDA(cudaMemcpyToSymbol(c_upopts, &o, sizeof(o)));
#endif
  return IAMRX_OK;
}
double godunov_get_option(int opt) {
  const UpOpts& o = g_upopts_host;
  switch (opt) {
    case IAMRX_OPT_SMALL_VEL: return o.small_vel;
    case IAMRX_OPT_SLOPE_ORDER: return o.slope_order;
    case IAMRX_OPT_CORNER_FORM: return o.corner_adv;
    case IAMRX_OPT_EXTDIR_BOTH: return o.extdir_both;
    default: return 0.0;
  }
}

int extrap_vel_to_faces(const Bx& bx, C4 vel, C4 force, V4 umac, V4 vmac, V4 wmac, const AdvGeom& g,
                        int forces_in_trans, cudaStream_t s, int ppm, const AdvBC* bc) {
  if (!bx.ok()) return IAMRX_OK;
#if !defined(IX_EMUL)
  // A box that touches non-interior domain sides: the 8-cell slabs next to those sides go through the staged kernels (which carry
  // the boundary conditions), the rest of the box takes the fused tile kernel (as compute_aofs does).  The faces between a slab and
  // the rest are computed by both; the later launch (the rest) stores last.
  if (!ppm && bc && !bc->interior() && bx.nx() % tile::TB == 0 && bx.ny() % tile::TB == 0 && bx.nz() % tile::TB == 0 &&
      to_bcall(bc, bx, 3).any) {
    Bx rest = bx;
    bool cut = false;
    for (int d = 0; d < 3; ++d) {
      bool lo_bc = false, hi_bc = false;
      for (int n = 0; n < 3; ++n) { lo_bc |= bc->lo[n][d] != IAMRX_BC_INT_DIR; hi_bc |= bc->hi[n][d] != IAMRX_BC_INT_DIR; }
      if (lo_bc && rest.lo[d] - 4 <= bc->dlo[d] && rest.hi[d] - rest.lo[d] + 1 > tile::TB) {
        Bx p = rest; p.hi[d] = p.lo[d] + tile::TB - 1; rest.lo[d] += tile::TB; cut = true;
        const int rc = extrap_vel_to_faces_staged(p, vel, force, umac, vmac, wmac, g, forces_in_trans, s, ppm, bc);
        if (rc) return rc;
      }
      if (hi_bc && rest.hi[d] + 4 >= bc->dhi[d] && rest.hi[d] - rest.lo[d] + 1 > tile::TB) {
        Bx p = rest; p.lo[d] = p.hi[d] - tile::TB + 1; rest.hi[d] -= tile::TB; cut = true;
        const int rc = extrap_vel_to_faces_staged(p, vel, force, umac, vmac, wmac, g, forces_in_trans, s, ppm, bc);
        if (rc) return rc;
      }
    }
    if (cut) {
      if (to_bcall(bc, rest, 3).any) return extrap_vel_to_faces_staged(rest, vel, force, umac, vmac, wmac, g, forces_in_trans, s, ppm, bc);
      return extrap_vel_to_faces(rest, vel, force, umac, vmac, wmac, g, forces_in_trans, s, ppm, nullptr);
    }
  }
#endif
  return extrap_vel_to_faces_staged(bx, vel, force, umac, vmac, wmac, g, forces_in_trans, s, ppm, bc);
}

// one box: the fused tile kernel when it applies, else the three staged kernels
int extrap_vel_to_faces_staged(const Bx& bx, C4 vel, C4 force, V4 umac, V4 vmac, V4 wmac, const AdvGeom& g,
                               int forces_in_trans, cudaStream_t s, int ppm, const AdvBC* bc) {
  if (!bx.ok()) return IAMRX_OK;
  ProfScope prof_(IAMRX_PROF_EXTRAP, bx.npts(), (double)bx.npts() * 72.0, s);
  EvArgs e{};
  e.bx = bx; e.vel = vel; e.force = force; e.fit = forces_in_trans; e.ppm = ppm;
  e.bc = to_bcall(bc, bx, 3);
  e.dt = g.dt; e.dtdx = g.dt / g.dx[0]; e.dtdy = g.dt / g.dx[1]; e.dtdz = g.dt / g.dx[2];
#if !defined(IX_EMUL)
  {
    // fused tile kernel: boxes made of whole 8^3 tiles with no physical boundary within reach of the stencil, PLM
    static int on = -1;
    if (on < 0) { const char* ev = getenv("IAMRX_EV_TILE"); on = (ev && ev[0] == '0') ? 0 : 1; }
    using namespace tile;
    const C4 um{umac.p, umac.l0, umac.l1, umac.l2, umac.js, umac.ks, umac.ns}, vm{vmac.p, vmac.l0, vmac.l1, vmac.l2, vmac.js, vmac.ks, vmac.ns},
        wm{wmac.p, wmac.l0, wmac.l1, wmac.l2, wmac.js, wmac.ks, wmac.ns};
    if (on && !ppm && !e.bc.any && bx.nx() % TB == 0 && bx.ny() % TB == 0 && bx.nz() % TB == 0 && fits32(vel, bx, 3) && fits32(force, bx, 1) &&
        fits32(um, bx, 1) && fits32(vm, bx, 1) && fits32(wm, bx, 1)) {
      static bool attr = false;
      if (!attr) { IX_CUDA(cudaFuncSetAttribute(ev_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EV_SMEM_BYTES)); attr = true; }
      IX_LAUNCH(ev_tile_kernel, dim3(bx.nx() / TB, bx.ny() / TB, bx.nz() / TB), dim3(NT, 1, 1), EV_SMEM_BYTES, s, e, umac, vmac, wmac);
      return check_launch("ev_tile");
    }
  }
#endif
  ScratchOwner so;
  if (so.init(bx, B_N) != IAMRX_OK) return IAMRX_ERR_CUDA;
  Bx R1 = grow(bx, 1); R1.hi[0]++; R1.hi[1]++; R1.hi[2]++;
  if (e.bc.any) IX_LAUNCH(ev_edge_kernel<true>, grid_for(R1, 1), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
    else IX_LAUNCH(ev_edge_kernel<false>, grid_for(R1, 1), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
  int rc = check_launch("ev_edge");
  if (rc) return rc;
  if (e.bc.any) IX_LAUNCH(ev_corner_kernel<true>, grid_for(R1, 1), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
    else IX_LAUNCH(ev_corner_kernel<false>, grid_for(R1, 1), dim3(TX, TY, 1), 0, s, e, so.sc, R1);
  rc = check_launch("ev_corner");
  if (rc) return rc;
  Bx R2 = bx; R2.hi[0]++; R2.hi[1]++; R2.hi[2]++;
  if (e.bc.any) IX_LAUNCH(ev_final_kernel<true>, grid_for(R2, 1), dim3(TX, TY, 1), 0, s, e, so.sc, umac, vmac, wmac, R2);
    else IX_LAUNCH(ev_final_kernel<false>, grid_for(R2, 1), dim3(TX, TY, 1), 0, s, e, so.sc, umac, vmac, wmac, R2);
  return check_launch("ev_final");
}

}  // namespace k
}  // namespace ix
