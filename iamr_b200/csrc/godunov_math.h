// godunov_math.h -- per-point arithmetic of the Godunov (PLM) advection path,
// shared by the staged kernels in godunov.cu and the fused tile kernel.
//
// Restates AMReX-Hydro's hydro_godunov_plm / hydro_godunov_edge_state_3D /
// hydro_godunov_extrap_vel_to_faces_3D / hydro_godunov_corner_couple and
// AMReX_Slopes_K (4th-order limited slopes), which IAMR reaches through
// Godunov::ExtrapVelToFaces (NavierStokesBase.cpp:4487-4491) and
// HydroUtils::ComputeFluxesOnBoxFromState (NavierStokesBase.cpp:4701-4717).
// Those sources are not vendored in the reference tree (Exec/Make.IAMR:15-19);
// the formulas are the published algorithm (Almgren et al., JCP 142 (1998);
// SURVEY.md Appendix A.2-A.5).  Domain boundaries: the ext_dir / hoextrap one-sided slopes
// (amrex_calc_*slope_extdir), the PPM boundary parabolas and hydro_bcs_K.H Set{X,Y,Z}EdgeBCs.
#pragma once
#include "common.h"

namespace ix {
namespace gd {

constexpr double SMALL_VEL = 1.0e-8;

// ---- UNVERIFIED-UPSTREAM switches (include/iamrx.h iamrx_set_option) -------------------------------
// Details of AMReX-Hydro's Godunov that this repository restates from memory of the upstream sources and cannot check (they
// are not vendored in the reference).  Each is a run-time switch, mirrored in the oracle, so that a site with the upstream
// source can flip a choice instead of editing kernels.  Process-wide; device copy in constant memory.
struct UpOpts {
  double small_vel;   // |u| below which a face velocity counts as zero in the upwinding (hydro_constants: 1e-8)
  int slope_order;    // 4 (Godunov default, AMReX_Slopes_K.H order 4) or 2 (monotonised central)
  int corner_adv;     // corner coupling of non-conservative states: 0 flux form minus q div(u) (hydro_godunov_corner_couple.H as
                      // restated), 1 advective form 1/2 (u+ + u-)(q+ - q-)
  int extdir_both;    // ext_dir faces: 0 only the outside state takes the boundary value (tangential components keep the traced
                      // interior state; hydro_bcs_K.H after the turbulent-inflow change), 1 both states take it
};
#if !defined(IX_EMUL)
__constant__ UpOpts c_upopts = {1.0e-8, 4, 0, 0};
#endif
static UpOpts g_upopts_host = {1.0e-8, 4, 0, 0};
IX_HD const UpOpts& upopts() {
#if defined(__CUDA_ARCH__)
  return c_upopts;
#else
  return g_upopts_host;
#endif
}

template <int D> struct E {  // unit offset of direction D
  static constexpr int x = (D == 0), y = (D == 1), z = (D == 2);
};

template <int D, class A>
IX_HD double sh(const A& a, int i, int j, int k, int o) {  // a(idx + o*e_D)
  return a(i + o * E<D>::x, j + o * E<D>::y, k + o * E<D>::z);
}

// (sign(dcen) * x for x >= 0 is written copysign(x, dcen): the same bits, signed zeros included, without the multiplication)
IX_HD double lim2(double dlft, double drgt) {  // limited 2nd-order difference
  const double dcen = 0.5 * (dlft + drgt);
  const double slop = 2.0 * fmin(fabs(dlft), fabs(drgt));
  const double dlim = (dlft * drgt >= 0.0) ? slop : 0.0;
  return copysign(fmin(dlim, fabs(dcen)), dcen);
}

// 4th-order limited slope from the five values q(i-2..i+2) along one direction
// (amrex_calc_{x,y,z}slope, order 4)
IX_HD double slope4_vals(double qm2, double qm, double q0, double qp, double qp2) {
  if (upopts().slope_order == 2) return lim2(q0 - qm, qp - q0);
  const double dxl = lim2(qm - qm2, q0 - qm);
  const double dxr = lim2(qp - q0, qp2 - qp);
  const double dlft = q0 - qm, drgt = qp - q0;
  const double dcen = 0.5 * (dlft + drgt);
  const double slop = 2.0 * fmin(fabs(dlft), fabs(drgt));
  const double dlim = (dlft * drgt >= 0.0) ? slop : 0.0;
  return copysign(fmin(dlim, fabs((4.0 / 3.0) * dcen - (1.0 / 6.0) * (dxl + dxr))), dcen);
}

// 4th-order limited slope of q along D at cell (i,j,k)
template <int D, class A>
IX_HD double slope4(const A& q, int i, int j, int k) {
  return slope4_vals(sh<D>(q, i, j, k, -2), sh<D>(q, i, j, k, -1), q(i, j, k), sh<D>(q, i, j, k, 1), sh<D>(q, i, j, k, 2));
}

// Cursor: a read-only view of one component positioned at a fixed (i,j,k); neighbours are
// addressed by small relative offsets, so the per-access index arithmetic is one 32-bit
// add (the offsets are compile-time constants after inlining) instead of the 64-bit
// (i-l0) + (j-l1)*js + (k-l2)*ks of the absolute views.
struct Cur {
  const double* p;
  int js, ks;
  IX_HD double operator()(int di, int dj, int dk) const { return p[di + dj * js + dk * ks]; }
  IX_HD bool ok() const { return p != nullptr; }
};
template <class V>
IX_HD Cur cur_at(const V& v, int n, int i, int j, int k) {
  if (!v.p) return Cur{nullptr, 0, 0};
  return Cur{v.p + n * v.ns + ((i - v.l0) + (j - v.l1) * v.js + (k - v.l2) * v.ks), (int)v.js, (int)v.ks};
}
template <int D> IX_HD Cur below(const Cur& c) {  // cursor moved one cell down along D
  return Cur{c.p - (E<D>::x + E<D>::y * c.js + E<D>::z * c.ks), c.js, c.ks};
}
template <int D> IX_HD double along(const Cur& c, int o) { return c(o * E<D>::x, o * E<D>::y, o * E<D>::z); }
// value at offset a*e_D1 + b*e_D2
template <int D1, int D2, class C> IX_HD double rel2(const C& c, int a, int b) {
  return c(a * E<D1>::x + b * E<D2>::x, a * E<D1>::y + b * E<D2>::y, a * E<D1>::z + b * E<D2>::z);
}
template <int D> IX_HD double slope4c(const Cur& q) {
  return slope4_vals(along<D>(q, -2), along<D>(q, -1), q(0, 0, 0), along<D>(q, 1), along<D>(q, 2));
}

// ---- PPM (ns.advection_scheme = Godunov_PPM, NSB.cpp:552-554,4485; AMReX-Hydro hydro_godunov_ppm.H restated: van Leer
// limited edge values, Colella-Woodward monotonisation) ----------------------------------------------------------
IX_HD double vanleer(double s0, double sp1, double sm1) {
  const double dsc = 0.5 * (sp1 - sm1), dsl = 2.0 * (s0 - sm1), dsr = 2.0 * (sp1 - s0);
  return (dsl * dsr > 0.0) ? copysign(1.0, dsc) * fmin(fabs(dsc), fmin(fabs(dsl), fabs(dsr))) : 0.0;
}
// parabola edges (sm at the lower, sp at the upper face) of the cell whose five values along the direction are given
IX_HD void ppm_parabola(double sm2, double sm1, double s0, double sp1, double sp2, double& sm, double& sp) {
  const double d0 = vanleer(s0, sp1, sm1), dm = vanleer(sm1, s0, sm2), dp = vanleer(sp1, sp2, s0);
  sm = 0.5 * (s0 + sm1) - (1.0 / 6.0) * (d0 - dm);
  sm = fmin(fmax(sm, fmin(s0, sm1)), fmax(s0, sm1));
  sp = 0.5 * (sp1 + s0) - (1.0 / 6.0) * (dp - d0);
  sp = fmin(fmax(sp, fmin(s0, sp1)), fmax(s0, sp1));
  if ((sp - s0) * (s0 - sm) <= 0.0) { sm = s0; sp = s0; }
  else if (fabs(sp - s0) >= 2.0 * fabs(sm - s0)) sp = 3.0 * s0 - 2.0 * sm;
  else if (fabs(sm - s0) >= 2.0 * fabs(sp - s0)) sm = 3.0 * s0 - 2.0 * sp;
}
// averages of the parabola over the domain of dependence of the upper (Ip) / lower (Im) face for trace velocity v
IX_HD double ppm_ip(double s0, double sm, double sp, double v, double dtdx) {
  if (!(v > upopts().small_vel)) return s0;
  const double sg = fabs(v) * dtdx, s6 = 6.0 * s0 - 3.0 * (sm + sp);
  return sp - 0.5 * sg * ((sp - sm) - (1.0 - (2.0 / 3.0) * sg) * s6);
}
IX_HD double ppm_im(double s0, double sm, double sp, double v, double dtdx) {
  if (!(v < -upopts().small_vel)) return s0;
  const double sg = fabs(v) * dtdx, s6 = 6.0 * s0 - 3.0 * (sm + sp);
  return sm + 0.5 * sg * ((sp - sm) + (1.0 - (2.0 / 3.0) * sg) * s6);
}


// ---- physical domain boundaries ---------------------------------------------------------------------------------
// Boundary description of ONE direction for one component: amrex::BCRec codes (IAMRX_BC_*) of the low / high side and
// the domain's cell bounds in that direction.
struct BcD { int lo, hi, dlo, dhi; };
IX_HD bool bc_extdir_or_ho(int bc) { return bc == IAMRX_BC_EXT_DIR || bc == IAMRX_BC_HOEXTRAP; }
IX_HD double lim_os(double dl, double dr, double d) {   // one-sided limiter of the boundary slopes (dl, dr carry the factor 2)
  const double lim = (dl * dr >= 0.0) ? fmin(fabs(dl), fabs(dr)) : 0.0;
  return copysign(1.0, d) * fmin(lim, fabs(d));
}
// amrex_calc_{x,y,z}slope_extdir, order 4, from the five values along the direction; c = index of the cell along the
// direction.  Next to an ext_dir / hoextrap face the ghost value sits ON the face (Software.rst:206-213), hence the
// one-sided 4-point formula in the first cell and the revised neighbour slope in the second.
IX_HD double slope4_bc_vals(double qm2, double qm, double q0, double qp, double qp2, int c, const BcD& b) {
  const bool edlo = bc_extdir_or_ho(b.lo), edhi = bc_extdir_or_ho(b.hi);
  if (upopts().slope_order == 2) {   // order 2 next to an ext_dir / hoextrap face: the wall value sits half a cell away
    if (edlo && c == b.dlo) return lim_os(2.0 * (q0 - qm), 2.0 * (qp - q0), (qp + 3.0 * q0 - 4.0 * qm) / 3.0);
    if (edhi && c == b.dhi) return lim_os(2.0 * (q0 - qm), 2.0 * (qp - q0), -(qm + 3.0 * q0 - 4.0 * qp) / 3.0);
    return lim2(q0 - qm, qp - q0);
  }
  if (!(edlo && (c == b.dlo || c == b.dlo + 1)) && !(edhi && (c == b.dhi || c == b.dhi - 1)))
    return slope4_vals(qm2, qm, q0, qp, qp2);
  double dfm = lim2(qm - qm2, q0 - qm), dfp = lim2(qp - q0, qp2 - qp);
  const double dlft = q0 - qm, drgt = qp - q0, dcen = 0.5 * (dlft + drgt), dsgn = copysign(1.0, dcen);
  const double dlim = (dlft * drgt >= 0.0) ? 2.0 * fmin(fabs(dlft), fabs(drgt)) : 0.0;
  double s = dsgn * fmin(dlim, fabs((4.0 / 3.0) * dcen - (1.0 / 6.0) * (dfp + dfm)));
  if (edlo && c == b.dlo) {
    s = lim_os(2.0 * (q0 - qm), 2.0 * (qp - q0), -(16.0 / 15.0) * qm + 0.5 * q0 + (2.0 / 3.0) * qp - 0.1 * qp2);
  } else if (edlo && c == b.dlo + 1) {   // the slope of cell dlo, which enters dfm, is the one-sided one
    dfm = lim_os(2.0 * (qm - qm2), 2.0 * (q0 - qm), -(16.0 / 15.0) * qm2 + 0.5 * qm + (2.0 / 3.0) * q0 - 0.1 * qp);
    s = dsgn * fmin(dlim, fabs((4.0 / 3.0) * dcen - (1.0 / 6.0) * (dfp + dfm)));
  }
  if (edhi && c == b.dhi) {
    s = lim_os(2.0 * (q0 - qm), 2.0 * (qp - q0), (16.0 / 15.0) * qp - 0.5 * q0 - (2.0 / 3.0) * qm + 0.1 * qm2);
  } else if (edhi && c == b.dhi - 1) {
    dfp = lim_os(2.0 * (qp - q0), 2.0 * (qp2 - qp), (16.0 / 15.0) * qp2 - 0.5 * qp - (2.0 / 3.0) * q0 + 0.1 * qm);
    s = dsgn * fmin(dlim, fabs((4.0 / 3.0) * dcen - (1.0 / 6.0) * (dfp + dfm)));
  }
  return s;
}
// PPM parabola with the boundary treatment of hydro_godunov_ppm.H (SetXBCs): next to an ext_dir / hoextrap face the
// boundary-side edge value is the face value itself and the other edge the one-sided cubic, clamped
IX_HD double clamp2(double v, double a, double b) { return fmin(fmax(v, fmin(a, b)), fmax(a, b)); }
IX_HD void ppm_parabola_bc(double sm2, double sm1, double s0, double sp1, double sp2, int c, const BcD& b,
                           double& sm, double& sp) {
  ppm_parabola(sm2, sm1, s0, sp1, sp2, sm, sp);
  const bool edlo = bc_extdir_or_ho(b.lo), edhi = bc_extdir_or_ho(b.hi);
  auto mono = [&](double& m, double& p) {
    if ((p - s0) * (s0 - m) <= 0.0) { m = s0; p = s0; }
    else if (fabs(p - s0) >= 2.0 * fabs(m - s0)) p = 3.0 * s0 - 2.0 * m;
    else if (fabs(m - s0) >= 2.0 * fabs(p - s0)) m = 3.0 * s0 - 2.0 * p;
  };
  if (edlo && c == b.dlo) {          // sm1 is the face value
    sp = clamp2(-0.2 * sm1 + 0.75 * s0 + 0.5 * sp1 - 0.05 * sp2, sp1, s0);
    sm = sm1;
  } else if (edlo && c == b.dlo + 1) {   // sm2 is the face value, sm1 the first cell
    const double e1 = clamp2(-0.2 * sm2 + 0.75 * sm1 + 0.5 * s0 - 0.05 * sp1, s0, sm1);
    const double d0 = vanleer(s0, sp1, sm1), dp = vanleer(sp1, sp2, s0);
    const double e2 = clamp2(0.5 * (sp1 + s0) - (1.0 / 6.0) * (dp - d0), s0, sp1);
    sm = e1; sp = e2; mono(sm, sp);
  }
  if (edhi && c == b.dhi) {
    sm = clamp2(-0.2 * sp1 + 0.75 * s0 + 0.5 * sm1 - 0.05 * sm2, sm1, s0);
    sp = sp1;
  } else if (edhi && c == b.dhi - 1) {
    const double e2 = clamp2(-0.2 * sp2 + 0.75 * sp1 + 0.5 * s0 - 0.05 * sm1, s0, sp1);
    const double d0 = vanleer(s0, sp1, sm1), dm = vanleer(sm1, s0, sm2);
    const double e1 = clamp2(0.5 * (s0 + sm1) - (1.0 / 6.0) * (d0 - dm), s0, sm1);
    sm = e1; sp = e2; mono(sm, sp);
  }
}
// hydro_bcs_K.H Set{X,Y,Z}EdgeBCs: boundary conditions on the traced states (lo, hi) of the face with index f along
// the direction.  qbelow / qabove = the cell values either side of the face (the ghost value at an ext_dir face).
// normal_vel: the component is the velocity normal to the face (is_velocity && n == dir): a Dirichlet face then pins
// both states, otherwise the interior state may still carry information out of the domain.
IX_HD void set_edge_bc(double& lo, double& hi, double qbelow, double qabove, int f, const BcD& b, bool normal_vel) {
  if (f == b.dlo) {
    if (b.lo == IAMRX_BC_EXT_DIR) { lo = qbelow; if (normal_vel || upopts().extdir_both) hi = lo; }
    else if (b.lo == IAMRX_BC_FOEXTRAP || b.lo == IAMRX_BC_HOEXTRAP || b.lo == IAMRX_BC_REFLECT_EVEN) lo = hi;
    else if (b.lo == IAMRX_BC_REFLECT_ODD) { lo = 0.0; hi = 0.0; }
  } else if (f == b.dhi + 1) {
    if (b.hi == IAMRX_BC_EXT_DIR) { hi = qabove; if (normal_vel || upopts().extdir_both) lo = hi; }
    else if (b.hi == IAMRX_BC_FOEXTRAP || b.hi == IAMRX_BC_HOEXTRAP || b.hi == IAMRX_BC_REFLECT_EVEN) hi = lo;
    else if (b.hi == IAMRX_BC_REFLECT_ODD) { lo = 0.0; hi = 0.0; }
  }
}
// outflow faces (foextrap / hoextrap) of the final states: no inflow of the normal velocity through an outflow face.
// `clip`: ExtrapVelToFaces always clips; ComputeEdgeState only for the normal velocity component when the MAC
// velocity points into the domain.
IX_HD void outflow_bc(double& lo, double& hi, int f, const BcD& b, bool clip_lo, bool clip_hi) {
  if (f == b.dlo && (b.lo == IAMRX_BC_FOEXTRAP || b.lo == IAMRX_BC_HOEXTRAP)) {
    if (clip_lo) hi = fmin(hi, 0.0);
    lo = hi;
  }
  if (f == b.dhi + 1 && (b.hi == IAMRX_BC_FOEXTRAP || b.hi == IAMRX_BC_HOEXTRAP)) {
    if (clip_hi) lo = fmax(lo, 0.0);
    hi = lo;
  }
}

// upwind the pair (lo, hi) with a given face velocity (ComputeEdgeState /
// transverse states of ExtrapVelToFaces)
IX_HD double upwind(double lo, double hi, double vel) {
  const double st = (vel >= 0.0) ? lo : hi;
  const double fu = (fabs(vel) < upopts().small_vel) ? 0.0 : 1.0;
  return fu * st + (1.0 - fu) * 0.5 * (hi + lo);
}

// Riemann upwinding of a normal velocity by itself (u_ad and the final u_mac of
// ExtrapVelToFaces)
IX_HD double riemann(double lo, double hi) {
  const double st = ((lo + hi) >= 0.0) ? lo : hi;
  const bool ltm = ((lo <= 0.0 && hi >= 0.0) || (fabs(lo + hi) < upopts().small_vel));
  return ltm ? 0.0 : st;
}

}  // namespace gd
}  // namespace ix
