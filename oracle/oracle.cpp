// oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.  See oracle.h for scope and the
// "parity unpinned" statement.  CPU restatement of the IAMR hot path on one fully
// periodic box.  Every function cites the reference call site it follows
// (/root/reference = AMReX-Fluids/IAMR @ f46ba59; NSB.cpp =
// Source/NavierStokesBase.cpp, NS.cpp = Source/NavierStokes.cpp) and, where the
// arithmetic lives in the un-vendored AMReX / AMReX-Hydro, the upstream routine it
// restates.
//
// Implementation notes: fields are stored with ghost layers and refreshed by an
// explicit periodic fill, loops run over the interior with plain indexing (OpenMP
// over k).  The Godunov part is written the way AMReX-Hydro stages it (Im/Ip ->
// lo/hi -> upwinded edge -> corner-coupled -> final), i.e. deliberately NOT the
// way the CUDA kernels are organised.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <array>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// Periodicity of the problem the current extern "C" call works on (geometry.is_periodic).  Every entry point sets it
// (PerScope); Arr::fill_periodic wraps only the periodic directions, physical boundaries are filled explicitly.
int g_per[3] = {1, 1, 1};
struct PerScope {
  int old[3];
  explicit PerScope(const int* per) { for (int d = 0; d < 3; ++d) { old[d] = g_per[d]; g_per[d] = per ? per[d] : 1; } }
  ~PerScope() { for (int d = 0; d < 3; ++d) g_per[d] = old[d]; }
};
// amrex::BCType math codes (AMReX_BC_TYPES.H; NS_BC.H:7-55 maps the physical types onto them)
enum { BC_INT_DIR = 0, BC_REFLECT_ODD = -1, BC_REFLECT_EVEN = 1, BC_FOEXTRAP = 2, BC_EXT_DIR = 3, BC_HOEXTRAP = 4 };
// amrex::LinOpBCType subset (Diffusion.cpp:1887-1999, MacProj.cpp:1187-1208, Projection.cpp:2436-2464)
enum { LO_PERIODIC = 0, LO_DIRICHLET = 1, LO_NEUMANN = 2, LO_REFLECT_ODD = 3, LO_INFLOW = 4,
       LO_COARSE_FINE = 5 };   // a side of a fine AMR level that borders coarse cells: Dirichlet data half a COARSE cell beyond the face
// PhysBCType (inputs ns.lo_bc / ns.hi_bc, inputs.3d.taylorgreen:100-102)
enum { PHYS_INTERIOR = 0, PHYS_INFLOW = 1, PHYS_OUTFLOW = 2, PHYS_SYMMETRY = 3, PHYS_SLIPWALL = 4, PHYS_NOSLIPWALL = 5 };
struct BCRec { int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}; };

struct Arr {
  int n[3] = {0, 0, 0}, ng = 0, nc = 0;
  long sj = 0, sk = 0, sc = 0;
  std::vector<double> d;
  Arr() {}
  Arr(const int n_[3], int nc_, int ng_) { define(n_, nc_, ng_); }
  void define(const int n_[3], int nc_, int ng_) {
    for (int q = 0; q < 3; ++q) n[q] = n_[q];
    nc = nc_; ng = ng_;
    sj = n[0] + 2 * ng; sk = sj * (n[1] + 2 * ng); sc = sk * (n[2] + 2 * ng);
    d.assign((size_t)sc * nc, 0.0);
  }
  inline double& operator()(int i, int j, int k, int c = 0) { return d[(i + ng) + (j + ng) * sj + (k + ng) * sk + c * sc]; }
  inline double operator()(int i, int j, int k, int c = 0) const { return d[(i + ng) + (j + ng) * sj + (k + ng) * sk + c * sc]; }
  void setval(double v) { std::fill(d.begin(), d.end(), v); }
  // FabArray::FillBoundary(periodicity) for a single box covering the domain
  // (ghost cells beyond a NON-periodic side are left alone: the physical-boundary fills own them.  In a non-periodic
  // direction d the index n[d] -- the high face / node of face- and node-centred data -- is ordinary data, so the loops
  // over the other directions include it.)
  void fill_periodic() {
    if (ng == 0) return;
    const int h0 = g_per[0] ? 0 : 1, h1 = g_per[1] ? 0 : 1, h2 = g_per[2] ? 0 : 1;   // extra high index of non-periodic dirs
    for (int c = 0; c < nc; ++c) {
      if (g_per[0]) {
#pragma omp parallel for
        for (int k = 0; k < n[2] + h2; ++k)
          for (int j = 0; j < n[1] + h1; ++j)
            for (int g = 1; g <= ng; ++g) { (*this)(-g, j, k, c) = (*this)(n[0] - g, j, k, c); (*this)(n[0] - 1 + g, j, k, c) = (*this)(g - 1, j, k, c); }
      }
      if (g_per[1]) {
#pragma omp parallel for
        for (int k = 0; k < n[2] + h2; ++k)
          for (int g = 1; g <= ng; ++g)
            for (int i = (g_per[0] ? -ng : 0); i < n[0] + (g_per[0] ? ng : h0); ++i) { (*this)(i, -g, k, c) = (*this)(i, n[1] - g, k, c); (*this)(i, n[1] - 1 + g, k, c) = (*this)(i, g - 1, k, c); }
      }
      if (g_per[2]) {
        for (int g = 1; g <= ng; ++g) {
#pragma omp parallel for
          for (int j = (g_per[1] ? -ng : 0); j < n[1] + (g_per[1] ? ng : h1); ++j)
            for (int i = (g_per[0] ? -ng : 0); i < n[0] + (g_per[0] ? ng : h0); ++i) { (*this)(i, j, -g, c) = (*this)(i, j, n[2] - g, c); (*this)(i, j, n[2] - 1 + g, c) = (*this)(i, j, g - 1, c); }
        }
      }
    }
  }
  void load(const double* src, int c0 = 0, int ncopy = -1) {  // dense [c][k][j][i] -> interior
    if (ncopy < 0) ncopy = nc;
    for (int c = 0; c < ncopy; ++c)
#pragma omp parallel for
      for (int k = 0; k < n[2]; ++k)
        for (int j = 0; j < n[1]; ++j)
          for (int i = 0; i < n[0]; ++i) (*this)(i, j, k, c0 + c) = src[i + (long)n[0] * (j + (long)n[1] * (k + (long)n[2] * c))];
  }
  void store(double* dst, int c0 = 0, int ncopy = -1) const {
    if (ncopy < 0) ncopy = nc;
    for (int c = 0; c < ncopy; ++c)
#pragma omp parallel for
      for (int k = 0; k < n[2]; ++k)
        for (int j = 0; j < n[1]; ++j)
          for (int i = 0; i < n[0]; ++i) dst[i + (long)n[0] * (j + (long)n[1] * (k + (long)n[2] * c))] = (*this)(i, j, k, c0 + c);
  }
  // "padded" exchange format of the non-periodic entry points: the whole array INCLUDING its ng ghost layers,
  // [comp][nz+2ng][ny+2ng][nx+2ng] (the high face / node of a non-periodic direction lives at index n, inside the layer)
  void load_padded(const double* src) { std::copy(src, src + d.size(), d.begin()); }
  void store_padded(double* dst) const { std::copy(d.begin(), d.end(), dst); }
  void copy_from(const Arr& s, int sc0, int dc0, int ncopy) {  // interior only
    for (int c = 0; c < ncopy; ++c)
#pragma omp parallel for
      for (int k = 0; k < n[2]; ++k)
        for (int j = 0; j < n[1]; ++j)
          for (int i = 0; i < n[0]; ++i) (*this)(i, j, k, dc0 + c) = s(i, j, k, sc0 + c);
  }
  double norminf(int c) const {
    double m = 0.0;
#pragma omp parallel for reduction(max : m)
    for (int k = 0; k < n[2]; ++k)
      for (int j = 0; j < n[1]; ++j)
        for (int i = 0; i < n[0]; ++i) m = std::max(m, std::fabs((*this)(i, j, k, c)));
    return m;
  }
  double sum(int c) const {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s)
    for (int k = 0; k < n[2]; ++k)
      for (int j = 0; j < n[1]; ++j)
        for (int i = 0; i < n[0]; ++i) s += (*this)(i, j, k, c);
    return s;
  }
  long ncells() const { return (long)n[0] * n[1] * n[2]; }
};

#define FOR_CELLS(A, i, j, k)                 \
  _Pragma("omp parallel for") for (int k = 0; k < (A).n[2]; ++k) \
    for (int j = 0; j < (A).n[1]; ++j)        \
      for (int i = 0; i < (A).n[0]; ++i)

// faces of direction d (or nodes, d = 3): in a non-periodic direction the high face / node n[d] is ordinary data
#define FOR_IDX(A, hx, hy, hz, i, j, k)                                   \
  _Pragma("omp parallel for") for (int k = 0; k < (A).n[2] + (hz); ++k) \
    for (int j = 0; j < (A).n[1] + (hy); ++j)                            \
      for (int i = 0; i < (A).n[0] + (hx); ++i)
inline int hi_ext(int d, int dir) { return (d == dir || dir == 3) && !g_per[d] ? 1 : 0; }   // dir = face direction, 3 = nodal
#define FOR_FACES(A, dir, i, j, k) FOR_IDX(A, hi_ext(0, dir), hi_ext(1, dir), hi_ext(2, dir), i, j, k)
#define FOR_NODES(A, i, j, k) FOR_IDX(A, hi_ext(0, 3), hi_ext(1, 3), hi_ext(2, 3), i, j, k)

// ===========================================================================
// Physical boundary fill of cell-centred state data: what AmrLevel::FillPatch does outside the domain
// (amrex::GpuBndryFuncFab::ccfcdoit -> filcc_cell, then the user function for ext_dir: NS_bcfill.H:17-167 with the
// constant face values bcv of NavierStokes::get_bc_values, NS.cpp:108-237).  Cells outside the domain in ONE
// non-periodic direction first, then two (domain edges), then three (corners); in every cell the x, y, z rules are
// applied in that order (a later direction overwrites an earlier one), and then the ext_dir values in the same order.
// Ghost values of ext_dir faces therefore sit ON the face (Software.rst:206-213).
// ===========================================================================
void fill_physbc(Arr& a, int c0, int nc, const BCRec* bc, const double* bcv /* [6][nc] or null (= 0) */) {
  if (a.ng == 0 || (g_per[0] && g_per[1] && g_per[2])) return;
  const int ng = a.ng;
  for (int pass = 1; pass <= 3; ++pass) {
    for (int c = 0; c < nc; ++c) {
      const BCRec& b = bc[c];
#pragma omp parallel for
      for (int k = -ng; k < a.n[2] + ng; ++k)
        for (int j = -ng; j < a.n[1] + ng; ++j)
          for (int i = -ng; i < a.n[0] + ng; ++i) {
            const int idx[3] = {i, j, k};
            int nout = 0, side[3] = {0, 0, 0};
            for (int d = 0; d < 3; ++d)
              if (!g_per[d]) { if (idx[d] < 0) { side[d] = -1; ++nout; } else if (idx[d] >= a.n[d]) { side[d] = 1; ++nout; } }
            if (nout != pass) continue;
            double v = a(i, j, k, c0 + c);
            for (int d = 0; d < 3; ++d) {
              if (!side[d]) continue;
              const int code = side[d] < 0 ? b.lo[d] : b.hi[d];
              const int e = side[d] < 0 ? 0 : a.n[d] - 1;        // first interior cell on that side
              const int inw = side[d] < 0 ? 1 : -1;               // direction into the domain
              auto at = [&](int m) { int q[3] = {i, j, k}; q[d] = m; return a(q[0], q[1], q[2], c0 + c); };
              const int dist = side[d] < 0 ? -1 - idx[d] : idx[d] - a.n[d];   // 0 for the first ghost cell
              switch (code) {
                case BC_FOEXTRAP: v = at(e); break;
                case BC_HOEXTRAP:
                  if (dist > 0) v = at(e);
                  else if (a.n[d] >= 3) v = 0.125 * (15.0 * at(e) - 10.0 * at(e + inw) + 3.0 * at(e + 2 * inw));
                  else v = 0.5 * (3.0 * at(e) - at(e + inw));
                  break;
                case BC_REFLECT_EVEN: v = at(e + inw * dist); break;
                case BC_REFLECT_ODD: v = -at(e + inw * dist); break;
                default: break;   // int_dir: nothing; ext_dir: the user function below
              }
              a(i, j, k, c0 + c) = v;   // the next direction's rule may read this cell's row, not the cell itself
            }
            for (int d = 0; d < 3; ++d) {
              if (!side[d]) continue;
              const int code = side[d] < 0 ? b.lo[d] : b.hi[d];
              if (code == BC_EXT_DIR) v = bcv ? bcv[(d + (side[d] > 0 ? 3 : 0)) * nc + c] : 0.0;
            }
            a(i, j, k, c0 + c) = v;
          }
    }
  }
}

// Extrapolater::FirstOrderExtrap (NS.cpp:2046; also FillPatch of Gradp at walls = foextrap, NS_BC.H:27-38): every
// ghost cell outside a non-periodic side copies the nearest valid cell
void first_order_extrap(Arr& a, int c0, int nc) {
  BCRec b; for (int d = 0; d < 3; ++d) { b.lo[d] = BC_FOEXTRAP; b.hi[d] = BC_FOEXTRAP; }
  std::vector<BCRec> v(nc, b);
  fill_physbc(a, c0, nc, v.data(), nullptr);
}

// ===========================================================================
// Domain boundary conditions of the cell-centred linear operators (AMReX MLCellLinOp::applyBC ->
// mllinop_apply_bc_{x,y,z}; A.6).  The ghost cell beyond a non-periodic side is a linear function of the boundary value
// (ON the face) and of the first interior cells:
//   Dirichlet   : Lagrange extrapolation through the face value and min(maxorder, n+1) - 1 interior cells
//   Neumann     : ghost = first interior cell;   reflect_odd : ghost = -first interior cell
// The coefficient of the first interior cell (f0) also enters the smoother's diagonal (the delta term of abec_gsrb).
// ===========================================================================
struct LinBC {
  int lo[3][3], hi[3][3];   // [comp][dir]  LinOpBCType
  int maxorder = 2;
  // coarse-fine sides (MLCellLinOp with setCoarseFineBC; AMReX MLMGBndry::setBoxBC: a side that is not a domain face is
  // Dirichlet with bcloc = ratio * dx / 2): the Dirichlet value sits at x0 cell widths from the face, x0 = -ratio / 2 on the
  // level of the solve and -ratio / 2^(l+1) on multigrid level l (the location is a fixed physical distance: setLOBndryConds
  // is given the finest dx on every multigrid level)
  double cf_x0[3] = {-1.0, -1.0, -1.0};
  LinBC() { for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) { lo[c][d] = LO_PERIODIC; hi[c][d] = LO_PERIODIC; } }
  int code(int c, int d, int side) const { return side < 0 ? lo[c < 3 ? c : 0][d] : hi[c < 3 ? c : 0][d]; }
};
// Lagrange weights at x = -1/2 (ghost centre, in cell widths from the face at 0) for nodes {0 (face), 1/2, 3/2, ...}:
// w[0] multiplies the face value, w[m] the m-th interior cell
inline void dirichlet_weights(int order, double w[5], double x0 = 0.0) {
  double x[5]; x[0] = x0; for (int m = 1; m < order; ++m) x[m] = m - 0.5;
  const double xg = -0.5;
  for (int m = 0; m < order; ++m) {
    double num = 1.0, den = 1.0;
    for (int q = 0; q < order; ++q) if (q != m) { num *= (xg - x[q]); den *= (x[m] - x[q]); }
    w[m] = num / den;
  }
  for (int m = order; m < 5; ++m) w[m] = 0.0;
}
inline int bc_order(const LinBC& bc, int nlen) { return std::max(2, std::min(std::min(bc.maxorder, nlen + 1), 4)); }
// coefficient of the first interior cell in the ghost-cell formula (the f0 of AMReX's undrrelxr)
inline double bc_f0(const LinBC& bc, int c, int d, int side, int nlen) {
  switch (bc.code(c, d, side)) {
    case LO_NEUMANN: return 1.0;
    case LO_REFLECT_ODD: return -1.0;
    case LO_DIRICHLET: { double w[5]; dirichlet_weights(bc_order(bc, nlen), w); return w[1]; }
    case LO_COARSE_FINE: { double w[5]; dirichlet_weights(bc_order(bc, nlen), w, bc.cf_x0[d]); return w[1]; }
    default: return 0.0;
  }
}
// face ghost cells of phi (1 layer) beyond the non-periodic sides; bv = array whose ghost cells hold the Dirichlet face
// values (inhomogeneous: MLLinOp::setLevelBC) or null (homogeneous: the multigrid corrections)
void apply_linop_bc(Arr& phi, int ncomp, const LinBC& bc, const Arr* bv) {
  for (int c = 0; c < ncomp; ++c)
    for (int d = 0; d < 3; ++d) {
      if (g_per[d]) continue;
      const int nl = phi.n[d];
      const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
      double w[5], wd[5], wcf[5];
      dirichlet_weights(bc_order(bc, nl), wd);
      dirichlet_weights(bc_order(bc, nl), wcf, bc.cf_x0[d]);
      for (int side = -1; side <= 1; side += 2) {
        const int code = bc.code(c, d, side);
        for (int m = 0; m < 5; ++m) w[m] = code == LO_COARSE_FINE ? wcf[m] : wd[m];
        const int g = side < 0 ? -1 : nl, e = side < 0 ? 0 : nl - 1, inw = side < 0 ? 1 : -1;
#pragma omp parallel for
        for (int b2 = 0; b2 < phi.n[d2]; ++b2)
          for (int b1 = 0; b1 < phi.n[d1]; ++b1) {
            int q[3]; q[d1] = b1; q[d2] = b2;
            auto P = [&](int m) -> double& { q[d] = m; return phi(q[0], q[1], q[2], c); };
            double v;
            if (code == LO_NEUMANN) v = P(e);
            else if (code == LO_REFLECT_ODD) v = -P(e);
            else {
              q[d] = g;
              v = w[0] * (bv ? (*bv)(q[0], q[1], q[2], c) : 0.0);
              for (int m = 1; m < 5; ++m) if (w[m] != 0.0) v += w[m] * P(e + inw * (m - 1));
            }
            P(g) = v;
          }
      }
    }
}

// ===========================================================================
// MLABecLaplacian  (AMReX MLABecLap_3D_K.H: mlabeclap_adotx / abec_gsrb; A.6)
//   L phi = a*alpha*phi - b * sum_d [ beta_d(i+1)(phi(i+1)-phi(i)) - beta_d(i)(phi(i)-phi(i-1)) ] / h_d^2
// ===========================================================================
struct AbecOp {
  double a = 0, b = 1;
  const Arr* alpha = nullptr;
  const Arr* beta[3] = {nullptr, nullptr, nullptr};  // face arrays with >= 1 ghost (filled)
  int bncomp = 1;
  double dxinv[3];
  const LinBC* bc = nullptr;   // domain BCs of the non-periodic sides (null: fully periodic)
  const Arr* bv = nullptr;     // Dirichlet face values (ghost cells) for the inhomogeneous form, null = homogeneous
};

inline void abec_fill(const AbecOp& op, Arr& phi, int ncomp) {
  phi.fill_periodic();
  if (op.bc) apply_linop_bc(phi, ncomp, *op.bc, op.bv);
}

void abec_apply(const AbecOp& op, Arr& phi, Arr& out, int ncomp) {  // phi ghosts are filled here
  abec_fill(op, phi, ncomp);
  const double hx = op.b * op.dxinv[0] * op.dxinv[0], hy = op.b * op.dxinv[1] * op.dxinv[1], hz = op.b * op.dxinv[2] * op.dxinv[2];
  for (int c = 0; c < ncomp; ++c) {
    const int cb = op.bncomp > 1 ? c : 0;
    const Arr &bx = *op.beta[0], &by = *op.beta[1], &bz = *op.beta[2];
    FOR_CELLS(out, i, j, k) {
      const double p = phi(i, j, k, c);
      double y = -hx * (bx(i + 1, j, k, cb) * (phi(i + 1, j, k, c) - p) - bx(i, j, k, cb) * (p - phi(i - 1, j, k, c)))
                 - hy * (by(i, j + 1, k, cb) * (phi(i, j + 1, k, c) - p) - by(i, j, k, cb) * (p - phi(i, j - 1, k, c)))
                 - hz * (bz(i, j, k + 1, cb) * (phi(i, j, k + 1, c) - p) - bz(i, j, k, cb) * (p - phi(i, j, k - 1, c)));
      if (op.a != 0.0) y += op.a * (*op.alpha)(i, j, k) * p;
      out(i, j, k, c) = y;
    }
  }
}

// one colour of red-black Gauss-Seidel (abec_gsrb): cells with (i+j+k+redblack) even
void abec_gsrb(const AbecOp& op, Arr& phi, const Arr& rhs, int ncomp, double omega, int redblack) {
  abec_fill(op, phi, ncomp);   // AMReX fills the ghost cells (applyBC) before every colour
  const double hx = op.b * op.dxinv[0] * op.dxinv[0], hy = op.b * op.dxinv[1] * op.dxinv[1], hz = op.b * op.dxinv[2] * op.dxinv[2];
  for (int c = 0; c < ncomp; ++c) {
    const int cb = op.bncomp > 1 ? c : 0;
    const Arr &bx = *op.beta[0], &by = *op.beta[1], &bz = *op.beta[2];
#pragma omp parallel for
    for (int k = 0; k < phi.n[2]; ++k)
      for (int j = 0; j < phi.n[1]; ++j) {
        const int ioff = (j + k + redblack) & 1;
        for (int i = ioff; i < phi.n[0]; i += 2) {
          double gamma = hx * (bx(i, j, k, cb) + bx(i + 1, j, k, cb)) + hy * (by(i, j, k, cb) + by(i, j + 1, k, cb)) +
                         hz * (bz(i, j, k, cb) + bz(i, j, k + 1, cb));
          if (op.a != 0.0) gamma += op.a * (*op.alpha)(i, j, k);
          const double rho = hx * (bx(i, j, k, cb) * phi(i - 1, j, k, c) + bx(i + 1, j, k, cb) * phi(i + 1, j, k, c)) +
                             hy * (by(i, j, k, cb) * phi(i, j - 1, k, c) + by(i, j + 1, k, cb) * phi(i, j + 1, k, c)) +
                             hz * (bz(i, j, k, cb) * phi(i, j, k - 1, c) + bz(i, j, k + 1, cb) * phi(i, j, k + 1, c));
          const double res = rhs(i, j, k, c) - (gamma * phi(i, j, k, c) - rho);
          // delta: the part of rho that depends on phi(i,j,k) itself through a boundary ghost cell (mlabeclap_gsrb cf0..cf5)
          double delta = 0.0;
          if (op.bc) {
            const LinBC& B = *op.bc;
            if (!g_per[0]) { if (i == 0) delta += hx * bx(i, j, k, cb) * bc_f0(B, c, 0, -1, phi.n[0]); if (i == phi.n[0] - 1) delta += hx * bx(i + 1, j, k, cb) * bc_f0(B, c, 0, 1, phi.n[0]); }
            if (!g_per[1]) { if (j == 0) delta += hy * by(i, j, k, cb) * bc_f0(B, c, 1, -1, phi.n[1]); if (j == phi.n[1] - 1) delta += hy * by(i, j + 1, k, cb) * bc_f0(B, c, 1, 1, phi.n[1]); }
            if (!g_per[2]) { if (k == 0) delta += hz * bz(i, j, k, cb) * bc_f0(B, c, 2, -1, phi.n[2]); if (k == phi.n[2] - 1) delta += hz * bz(i, j, k + 1, cb) * bc_f0(B, c, 2, 1, phi.n[2]); }
          }
          phi(i, j, k, c) += omega / (gamma - delta) * res;
        }
      }
  }
}

// periodic wrap over the FULL extent (ghost rows of the other directions included): extends wall ghost cells to the
// transverse ghost positions at mixed periodic / physical edges
void fill_periodic_all(Arr& a) {
  const int ng = a.ng;
  for (int c = 0; c < a.nc; ++c)
    for (int d = 0; d < 3; ++d) {
      if (!g_per[d]) continue;
      const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
#pragma omp parallel for
      for (int b2 = -ng; b2 < a.n[d2] + ng; ++b2)
        for (int b1 = -ng; b1 < a.n[d1] + ng; ++b1)
          for (int g = 1; g <= ng; ++g) {
            int q[3], r[3]; q[d1] = r[d1] = b1; q[d2] = r[d2] = b2;
            q[d] = -g; r[d] = a.n[d] - g; a(q[0], q[1], q[2], c) = a(r[0], r[1], r[2], c);
            q[d] = a.n[d] - 1 + g; r[d] = g - 1; a(q[0], q[1], q[2], c) = a(r[0], r[1], r[2], c);
          }
    }
}

// MLTensorOp cross terms (AMReX MLTensor_3D_K.H mltensor_cross_terms_f{x,y,z} with kappa = 0; A.8):
// out += bscalar * div(F), F_x = (-eta*(-2/3)(dv/dy+dw/dz), -eta du/dy, -eta du/dz) on x-faces, etc.
// The transverse derivative of component c on a D-face is the mean of the two centred differences either side of the
// face; ON a non-periodic domain face (mltensor_d?_on_?face with bct / bv?lo / bv?hi) it is taken from the boundary
// data instead: Dirichlet -> centred difference of the FACE values (the ghost cells of the level-BC array; zero in
// the homogeneous form), Neumann -> centred difference of the interior cell row, reflect_odd -> 0.
void tensor_cross(const double dxinv[3], double bscalar, const Arr& ex, const Arr& ey, const Arr& ez, Arr& vel, Arr& out,
                  const LinBC* bc = nullptr, const Arr* bv = nullptr) {
  vel.fill_periodic();
  if (bc) { apply_linop_bc(vel, 3, *bc, bv); fill_periodic_all(vel); }   // face ghosts, then their periodic images
  Arr bvx;   // level-BC values with the periodic images of their wall ghost cells
  if (bc && bv) { bvx = *bv; fill_periodic_all(bvx); }
  const Arr* eta[3] = {&ex, &ey, &ez};
  const int* n = vel.n;
  // derivative along t of component c on the D-face whose upper cell is (i,j,k)
  auto dt_on_face = [&](int D, int t, int c, int i, int j, int k) {
    const double dti = dxinv[t];
    int q[3] = {i, j, k};
    auto V = [&](const Arr& A, int oD, int ot) { int r[3] = {q[0], q[1], q[2]}; r[D] += oD; r[t] += ot; return A(r[0], r[1], r[2], c); };
    if (bc && !g_per[D] && (q[D] == 0 || q[D] == n[D])) {
      const int side = q[D] == 0 ? -1 : 1;
      const int code = bc->code(c, D, side);
      const int og = side < 0 ? -1 : 0, oi = side < 0 ? 0 : -1;   // offsets of the ghost / interior cell along D
      if (code == LO_DIRICHLET) return bv ? (V(bvx, og, 1) - V(bvx, og, -1)) * (0.5 * dti) : 0.0;
      if (code == LO_NEUMANN) return (V(vel, oi, 1) - V(vel, oi, -1)) * (0.5 * dti);
      return 0.0;   // reflect_odd
    }
    return (V(vel, 0, 1) + V(vel, -1, 1) - V(vel, 0, -1) - V(vel, -1, -1)) * (0.25 * dti);
  };
  Arr fl[3] = {Arr(n, 3, 1), Arr(n, 3, 1), Arr(n, 3, 1)};
  const double twoThirds = 2.0 / 3.0;
  for (int D = 0; D < 3; ++D) {
    const int t1 = (D + 1) % 3, t2 = (D + 2) % 3;
    Arr& F = fl[D]; const Arr& E = *eta[D];
    FOR_FACES(F, D, i, j, k) {
      const double mu = E(i, j, k);
      // normal component: -(2/3) eta (d u_t1/d t1 + d u_t2/d t2); tangential component t: eta d u_D / d t
      F(i, j, k, D) = -mu * (-twoThirds * (dt_on_face(D, t1, t1, i, j, k) + dt_on_face(D, t2, t2, i, j, k)));
      F(i, j, k, t1) = -mu * dt_on_face(D, t1, D, i, j, k);
      F(i, j, k, t2) = -mu * dt_on_face(D, t2, D, i, j, k);
    }
    F.fill_periodic();
  }
  const double dxi = dxinv[0], dyi = dxinv[1], dzi = dxinv[2];
  for (int c = 0; c < 3; ++c) {
    FOR_CELLS(out, i, j, k) {
      out(i, j, k, c) += bscalar * (dxi * (fl[0](i + 1, j, k, c) - fl[0](i, j, k, c)) + dyi * (fl[1](i, j + 1, k, c) - fl[1](i, j, k, c)) +
                                    dzi * (fl[2](i, j, k + 1, c) - fl[2](i, j, k, c)));
    }
  }
}

// ===========================================================================
// Cell-centred geometric multigrid (AMReX MLMG V-cycle, A.7; MLCellLinOp restriction =
// mean of 8 children, interpolation = piecewise-constant add; coefficient coarsening =
// average_down (alpha) / average_down_faces (beta))
// ===========================================================================

// ===========================================================================
// BiCGStab (AMReX MLCGSolver::solve_bicgstab, first stage of the "bicgcg" bottom solver IAMR uses by default,
// Docs/sphinx_documentation/source/RunningProblems.rst:509-516): unpreconditioned, x0 = 0, homogeneous BCs, stop when
// |r|_inf <= rtol |r0|_inf.  Returns 0 or the break-down code (1 rho = 0, 2 <rh,v> = 0, 3 <t,t> = 0, 4 omega = 0, 8 cap).
// The coarsest level is tiny: plain serial loops.  nodal: data lives on nodes (index n[d] of a non-periodic direction included).
// ===========================================================================
template <class Apply>
static int bicgstab(Apply&& A, Arr& sol, const Arr& rhs, int ncomp, bool nodal, double rtol, int maxiter, int* iters) {
  int ex[3];
  for (int d = 0; d < 3; ++d) ex[d] = sol.n[d] + ((nodal && !g_per[d]) ? 1 : 0);
  auto each = [&](auto&& f) {
    for (int c = 0; c < ncomp; ++c) for (int k = 0; k < ex[2]; ++k) for (int j = 0; j < ex[1]; ++j) for (int i = 0; i < ex[0]; ++i) f(i, j, k, c);
  };
  auto dot = [&](const Arr& a, const Arr& b) { double s = 0; each([&](int i, int j, int k, int c) { s += a(i, j, k, c) * b(i, j, k, c); }); return s; };
  auto nrm = [&](const Arr& a) { double m = 0; each([&](int i, int j, int k, int c) { m = std::max(m, std::fabs(a(i, j, k, c))); }); return m; };
  Arr r(sol.n, ncomp, sol.ng), rh(sol.n, ncomp, sol.ng), p(sol.n, ncomp, sol.ng), v(sol.n, ncomp, sol.ng), sv(sol.n, ncomp, sol.ng), t(sol.n, ncomp, sol.ng);
  each([&](int i, int j, int k, int c) { r(i, j, k, c) = rhs(i, j, k, c); rh(i, j, k, c) = rhs(i, j, k, c); });
  *iters = 0;
  const double rnorm0 = nrm(r);
  if (rnorm0 == 0.0) return 0;
  const double target = rtol * rnorm0;
  double rho_1 = 0, alpha = 0, omega = 0;
  for (int nit = 1; nit <= maxiter; ++nit) {
    *iters = nit;
    const double rho = dot(rh, r);
    if (rho == 0.0) return 1;
    if (nit == 1) each([&](int i, int j, int k, int c) { p(i, j, k, c) = r(i, j, k, c); });
    else {
      const double beta = (rho / rho_1) * (alpha / omega);
      each([&](int i, int j, int k, int c) { const double q = p(i, j, k, c) - omega * v(i, j, k, c); p(i, j, k, c) = beta * q + r(i, j, k, c); });
    }
    A(v, p);
    const double rhTv = dot(rh, v);
    if (rhTv == 0.0) return 2;
    alpha = rho / rhTv;
    each([&](int i, int j, int k, int c) { sol(i, j, k, c) += alpha * p(i, j, k, c); sv(i, j, k, c) = r(i, j, k, c) - alpha * v(i, j, k, c); });
    { const double q = nrm(sv); if (!(q == q)) return 9; if (q <= target) return 0; }
    A(t, sv);
    const double tt = dot(t, t), ts = dot(t, sv);
    if (tt == 0.0) return 3;
    omega = ts / tt;
    each([&](int i, int j, int k, int c) { sol(i, j, k, c) += omega * sv(i, j, k, c); r(i, j, k, c) = sv(i, j, k, c) - omega * t(i, j, k, c); });
    { const double q = nrm(r); if (!(q == q)) return 9; if (q <= target) return 0; }
    if (omega == 0.0) return 4;
    rho_1 = rho;
  }
  return 8;
}

struct CellMG {
  struct Lev { int n[3]; double dxinv[3]; Arr alpha, beta[3], cor, res, rescor; LinBC bc; };
  std::vector<Lev> lv;
  int ncomp, bncomp;
  bool tensor;
  double a = 0, b = 1;
  const Arr* eta[3] = {nullptr, nullptr, nullptr};
  orc_mg mg;
  bool singular = false;
  LinBC bc; bool has_bc = false;   // domain BCs (setDomainBC + setMaxOrder)
  const Arr* bvals = nullptr;      // setLevelBC: array whose ghost cells hold the Dirichlet face values
  void set_bc(const LinBC& b, const Arr* levelbc) {
    bc = b; has_bc = !(g_per[0] && g_per[1] && g_per[2]); bvals = levelbc;
    for (Lev& L : lv) {   // the coarse-fine Dirichlet location in cell widths of each multigrid level (ratio 2)
      L.bc = b;
      for (int d = 0; d < 3; ++d) L.bc.cf_x0[d] = -0.5 * 2.0 * L.dxinv[d] / lv[0].dxinv[d];
    }
  }

  CellMG(const int n[3], const double dx[3], int ncomp_, bool tensor_, int max_coarsening) : ncomp(ncomp_), tensor(tensor_) {
    bncomp = tensor ? ncomp : 1;
    orc_mg_default(&mg);
    int cur[3] = {n[0], n[1], n[2]};
    double h[3] = {dx[0], dx[1], dx[2]};
    for (int l = 0; l <= max_coarsening; ++l) {
      lv.emplace_back();
      Lev& L = lv.back();
      for (int d = 0; d < 3; ++d) { L.n[d] = cur[d]; L.dxinv[d] = 1.0 / h[d]; }
      L.cor.define(cur, ncomp, 1); L.res.define(cur, ncomp, 0); L.rescor.define(cur, ncomp, 0);
      bool ok = true;
      for (int d = 0; d < 3; ++d) if (cur[d] % 2 != 0 || cur[d] / 2 < 2) ok = false;
      if (!ok) break;
      for (int d = 0; d < 3; ++d) { cur[d] /= 2; h[d] *= 2.0; }
    }
  }
  AbecOp op(int l, bool inhomog = false) const {
    AbecOp o; o.a = a; o.b = b; o.alpha = (a != 0.0) ? &lv[l].alpha : nullptr;
    for (int d = 0; d < 3; ++d) { o.beta[d] = &lv[l].beta[d]; o.dxinv[d] = lv[l].dxinv[d]; }
    o.bncomp = bncomp;
    if (has_bc) { o.bc = &lv[l].bc; o.bv = (inhomog && l == 0) ? bvals : nullptr; }
    return o;
  }
  void set_coeffs(const Arr* alpha, const Arr* e[3]) {
    for (int d = 0; d < 3; ++d) eta[d] = e[d];
    Lev& F = lv[0];
    if (a != 0.0 && alpha) { F.alpha.define(F.n, 1, 0); F.alpha.copy_from(*alpha, 0, 0, 1); }
    for (int d = 0; d < 3; ++d) {
      F.beta[d].define(F.n, bncomp, 1);
      for (int c = 0; c < bncomp; ++c) {
        const double fac = (tensor && c == d) ? 4.0 / 3.0 : 1.0;  // MLTensorOp: (4/3)eta + kappa on the normal component
        Arr& B = F.beta[d]; const Arr& E = *e[d];
        FOR_FACES(B, d, i, j, k) B(i, j, k, c) = fac * E(i, j, k);
      }
      F.beta[d].fill_periodic();
    }
    for (size_t l = 1; l < lv.size(); ++l) {
      Lev& C = lv[l]; Lev& Fi = lv[l - 1];
      if (a != 0.0 && alpha) {
        C.alpha.define(C.n, 1, 0);
        Arr& ca = C.alpha; const Arr& fa = Fi.alpha;
        FOR_CELLS(ca, i, j, k) {
          double s = 0; for (int dk = 0; dk < 2; ++dk) for (int dj = 0; dj < 2; ++dj) for (int di = 0; di < 2; ++di) s += fa(2 * i + di, 2 * j + dj, 2 * k + dk);
          ca(i, j, k) = 0.125 * s;
        }
      }
      for (int d = 0; d < 3; ++d) {
        C.beta[d].define(C.n, bncomp, 1);
        Arr& cb = C.beta[d]; const Arr& fb = Fi.beta[d];
        for (int c = 0; c < bncomp; ++c) {
          FOR_FACES(cb, d, i, j, k) {
            const int ii = 2 * i, jj = 2 * j, kk = 2 * k;
            double s;
            if (d == 0) s = fb(ii, jj, kk, c) + fb(ii, jj + 1, kk, c) + fb(ii, jj, kk + 1, c) + fb(ii, jj + 1, kk + 1, c);
            else if (d == 1) s = fb(ii, jj, kk, c) + fb(ii + 1, jj, kk, c) + fb(ii, jj, kk + 1, c) + fb(ii + 1, jj, kk + 1, c);
            else s = fb(ii, jj, kk, c) + fb(ii + 1, jj, kk, c) + fb(ii, jj + 1, kk, c) + fb(ii + 1, jj + 1, kk, c);
            cb(i, j, k, c) = 0.25 * s;
          }
        }
        cb.fill_periodic();
      }
    }
    singular = (a == 0.0);  // periodic / Neumann everywhere: the operator annihilates constants
    if (has_bc)
      for (int c = 0; c < ncomp && c < 3; ++c) for (int d = 0; d < 3; ++d)
        if (!g_per[d] && (bc.lo[c][d] == LO_DIRICHLET || bc.hi[c][d] == LO_DIRICHLET || bc.lo[c][d] == LO_REFLECT_ODD || bc.hi[c][d] == LO_REFLECT_ODD ||
                          bc.lo[c][d] == LO_COARSE_FINE || bc.hi[c][d] == LO_COARSE_FINE)) singular = false;
  }
  void smooth(int l, Arr& phi, const Arr& rhs, int nsweeps) {
    const AbecOp o = op(l);
    for (int s = 0; s < nsweeps; ++s)
      for (int rb = 0; rb < 2; ++rb) abec_gsrb(o, phi, rhs, ncomp, mg.omega, rb);
  }
  // cross: the top-level residual of the solve (inhomogeneous BCs + tensor cross terms); the cycle's correction
  // residuals are homogeneous and carry no cross terms
  void residual(int l, Arr& out, Arr& phi, const Arr& rhs, bool cross) {
    abec_apply(op(l, cross), phi, out, ncomp);
    if (tensor && l == 0 && cross) tensor_cross(lv[0].dxinv, b, *eta[0], *eta[1], *eta[2], phi, out, has_bc ? &bc : nullptr, bvals);
    for (int c = 0; c < ncomp; ++c) { FOR_CELLS(out, i, j, k) out(i, j, k, c) = rhs(i, j, k, c) - out(i, j, k, c); }
  }
  void apply(Arr& out, Arr& phi) {   // MLMG::apply: inhomogeneous BCs
    abec_apply(op(0, true), phi, out, ncomp);
    if (tensor) tensor_cross(lv[0].dxinv, b, *eta[0], *eta[1], *eta[2], phi, out, has_bc ? &bc : nullptr, bvals);
  }
  void make_solvable(Arr& r) {
    for (int c = 0; c < ncomp; ++c) {
      const double mean = r.sum(c) / (double)r.ncells();
      FOR_CELLS(r, i, j, k) r(i, j, k, c) -= mean;
    }
  }
  void vcycle() {
    const int nl = (int)lv.size();
    for (int l = 0; l < nl - 1; ++l) {
      Lev& L = lv[l];
      L.cor.setval(0.0);
      smooth(l, L.cor, L.res, mg.nu1);
      residual(l, L.rescor, L.cor, L.res, false);
      Arr& cr = lv[l + 1].res; const Arr& fr = L.rescor;
      for (int c = 0; c < ncomp; ++c) {
        FOR_CELLS(cr, i, j, k) {
          double s = 0; for (int dk = 0; dk < 2; ++dk) for (int dj = 0; dj < 2; ++dj) for (int di = 0; di < 2; ++di) s += fr(2 * i + di, 2 * j + dj, 2 * k + dk, c);
          cr(i, j, k, c) = 0.125 * s;
        }
      }
    }
    Lev& B = lv[nl - 1];
    B.cor.setval(0.0);
    if (singular && nl > 1) make_solvable(B.res);
    const int nsm = nl == 1 ? mg.nu1 + mg.nu2 : mg.bottom_sweeps;
    if (mg.bottom_solver == 1 && nl > 1) {
      int its = 0;
      const int ret = bicgstab([&](Arr& out, Arr& in) { abec_apply(op(nl - 1, false), in, out, ncomp); }, B.cor, B.res, ncomp, false,
                               mg.bottom_rtol, mg.bottom_maxiter, &its);
      mg.bottom_iters += its;
      if (ret != 0) { B.cor.setval(0.0); smooth(nl - 1, B.cor, B.res, nsm); }   // MLMG::bottomSolve falls back to smoothing
    } else {
      smooth(nl - 1, B.cor, B.res, nsm);
    }
    for (int l = nl - 2; l >= 0; --l) {
      Arr& fc = lv[l].cor; const Arr& cc = lv[l + 1].cor;
      for (int c = 0; c < ncomp; ++c) { FOR_CELLS(fc, i, j, k) fc(i, j, k, c) += cc(i / 2, j / 2, k / 2, c); }
      smooth(l, fc, lv[l].res, mg.nu2);
    }
  }
  int solve(Arr& sol, const Arr& rhs_in) {
    Arr rhs(lv[0].n, ncomp, 0);
    rhs.copy_from(rhs_in, 0, 0, ncomp);
    if (singular) make_solvable(rhs);
    double rhsnorm = 0; for (int c = 0; c < ncomp; ++c) rhsnorm = std::max(rhsnorm, rhs.norminf(c));
    Arr& res = lv[0].res;
    residual(0, res, sol, rhs, true);
    double r0 = 0; for (int c = 0; c < ncomp; ++c) r0 = std::max(r0, res.norminf(c));
    const double target = std::max(mg.atol, mg.rtol * std::max(rhsnorm, r0));
    double r = r0; int it = 0; int rc = 0;
    if (!(r0 <= target)) {
      rc = mg.max_iter;
      for (it = 1; it <= mg.max_iter; ++it) {
        vcycle();
        Arr& cor = lv[0].cor;
        for (int c = 0; c < ncomp; ++c) { FOR_CELLS(sol, i, j, k) sol(i, j, k, c) += cor(i, j, k, c); }
        residual(0, res, sol, rhs, true);
        r = 0; for (int c = 0; c < ncomp; ++c) r = std::max(r, res.norminf(c));
        if (r <= target) { rc = 0; break; }
      }
    }
    abec_fill(op(0, true), sol, ncomp);   // setFinalFillBC(true)
    mg.iters = it; mg.resnorm0 = r0; mg.resnorm = r; mg.rhsnorm = rhsnorm;
    return rc;
  }
};

// ===========================================================================
// MLNodeLaplacian (A.9): Q1 finite elements, one sigma per cell.  The operator is
// assembled here from the element stiffness matrices (1-D stiffness x 1-D mass
// tensor products) instead of a hand-expanded 27-point formula:
//   (A phi)_node = -(1/vol) * sum_{8 cells} sigma_cell * sum_{corners m} K_cell[node,m] phi_m
// ===========================================================================
struct Q1 {
  double K[8][8];  // element stiffness / volume, corners ordered (cx + 2cy + 4cz)
  explicit Q1(const double dxinv[3]) {
    const double S[2][2] = {{1, -1}, {-1, 1}};         // 1-D stiffness * h
    const double M[2][2] = {{2. / 6, 1. / 6}, {1. / 6, 2. / 6}};  // 1-D mass / h
    for (int a = 0; a < 8; ++a)
      for (int b = 0; b < 8; ++b) {
        const int ax = a & 1, ay = (a >> 1) & 1, az = a >> 2, bx = b & 1, by = (b >> 1) & 1, bz = b >> 2;
        K[a][b] = dxinv[0] * dxinv[0] * S[ax][bx] * M[ay][by] * M[az][bz] + dxinv[1] * dxinv[1] * M[ax][bx] * S[ay][by] * M[az][bz] +
                  dxinv[2] * dxinv[2] * M[ax][bx] * M[ay][by] * S[az][bz];
      }
  }
};

// returns A*phi at node (i,j,k) and the diagonal coefficient; sig has ghosts filled
inline double nodal_ax(const Q1& q, const Arr& sig, const Arr& phi, int i, int j, int k, double& diag) {
  double y = 0.0; diag = 0.0;
  for (int cz = 0; cz < 2; ++cz) for (int cy = 0; cy < 2; ++cy) for (int cx = 0; cx < 2; ++cx) {
    // cell whose corner (cx,cy,cz) is this node
    const int ci = i - cx, cj = j - cy, ck = k - cz;
    const double s = sig(ci, cj, ck);
    const int a = cx + 2 * cy + 4 * cz;
    for (int m = 0; m < 8; ++m) {
      const int mx = m & 1, my = (m >> 1) & 1, mz = m >> 2;
      y -= s * q.K[a][m] * phi(ci + mx, cj + my, ck + mz);
    }
    diag -= s * q.K[a][a];
  }
  return y;
}

// Domain boundary conditions of the nodal operator (Projection.cpp:2436-2464: outflow -> Dirichlet, inflow -> inflow,
// everything else -> Neumann).  AMReX MLNodeLaplacian (A.9): Neumann / inflow sides reflect the ghost node
// (phi(lo-1) = phi(lo+1), mlndlap_applybc) and copy sigma into the ghost cell (mlndlap_fillbc_cc), the right-hand side
// of nodes ON such a side is doubled once per direction (mlndlap_impose_neumann_bc), and nodes ON a Dirichlet side are
// held at zero (dirichlet mask).  Node arrays carry 2 ghost layers: the ghost node of a high side is index n+1.
struct NodalBC {
  int lo[3] = {LO_PERIODIC, LO_PERIODIC, LO_PERIODIC}, hi[3] = {LO_PERIODIC, LO_PERIODIC, LO_PERIODIC};
  bool neu(int d, int side) const { const int c = side < 0 ? lo[d] : hi[d]; return !g_per[d] && (c == LO_NEUMANN || c == LO_INFLOW); }
  bool dir(int d, int side) const { const int c = side < 0 ? lo[d] : hi[d]; return !g_per[d] && c == LO_DIRICHLET; }
};
inline bool nodal_masked(const NodalBC& bc, const int n[3], int i, int j, int k) {   // node on a Dirichlet side
  const int idx[3] = {i, j, k};
  for (int d = 0; d < 3; ++d) if ((idx[d] == 0 && bc.dir(d, -1)) || (idx[d] == n[d] && bc.dir(d, 1))) return true;
  return false;
}
inline double nodal_weight(const NodalBC& bc, const int n[3], int i, int j, int k) {   // mlndlap_set_dot_mask: 1/2 per Neumann side
  const int idx[3] = {i, j, k};
  double w = 1.0;
  for (int d = 0; d < 3; ++d) if ((idx[d] == 0 && bc.neu(d, -1)) || (idx[d] == n[d] && bc.neu(d, 1))) w *= 0.5;
  return w;
}
void nodal_fill_phi(Arr& phi, const NodalBC& bc) {
  phi.fill_periodic();
  for (int d = 0; d < 3; ++d) {
    if (g_per[d]) continue;
    const int d1 = (d + 1) % 3, d2 = (d + 2) % 3, ng = phi.ng;
#pragma omp parallel for
    for (int b2 = -ng; b2 <= phi.n[d2] + ng - 1; ++b2)
      for (int b1 = -ng; b1 <= phi.n[d1] + ng - 1; ++b1) {
        int q[3], r[3]; q[d1] = r[d1] = b1; q[d2] = r[d2] = b2;
        if (bc.neu(d, -1)) { q[d] = -1; r[d] = 1; phi(q[0], q[1], q[2]) = phi(r[0], r[1], r[2]); }
        if (bc.neu(d, 1)) { q[d] = phi.n[d] + 1; r[d] = phi.n[d] - 1; phi(q[0], q[1], q[2]) = phi(r[0], r[1], r[2]); }
      }
  }
}
void nodal_fill_sigma(Arr& sig, const NodalBC& bc) {
  sig.fill_periodic();
  for (int d = 0; d < 3; ++d) {
    if (g_per[d]) continue;
    const int d1 = (d + 1) % 3, d2 = (d + 2) % 3, ng = sig.ng;
#pragma omp parallel for
    for (int b2 = -ng; b2 < sig.n[d2] + ng; ++b2)
      for (int b1 = -ng; b1 < sig.n[d1] + ng; ++b1) {
        int q[3], r[3]; q[d1] = r[d1] = b1; q[d2] = r[d2] = b2;
        q[d] = -1; r[d] = 0; sig(q[0], q[1], q[2]) = bc.neu(d, -1) ? sig(r[0], r[1], r[2]) : 0.0;
        q[d] = sig.n[d]; r[d] = sig.n[d] - 1; sig(q[0], q[1], q[2]) = bc.neu(d, 1) ? sig(r[0], r[1], r[2]) : 0.0;
      }
  }
}
const NodalBC PERIODIC_NBC;

void nodal_adotx(const double dxinv[3], Arr& sig, Arr& phi, Arr& out, const NodalBC& bc = PERIODIC_NBC) {
  nodal_fill_phi(phi, bc);
  const Q1 q(dxinv);
  FOR_NODES(out, i, j, k) {
    double dg;
    out(i, j, k) = nodal_masked(bc, out.n, i, j, k) ? 0.0 : nodal_ax(q, sig, phi, i, j, k, dg);
  }
}

// multi-colour Gauss-Seidel: colour = (i&1) + 2(j&1) + 4(k&1); nodes of one colour are
// mutually uncoupled under the 27-point stencil (even n per direction)
void nodal_gs(const double dxinv[3], Arr& sig, const Arr& rhs, int color, Arr& phi, const NodalBC& bc = PERIODIC_NBC) {
  nodal_fill_phi(phi, bc);
  const Q1 q(dxinv);
  const int c0 = color & 1, c1 = (color >> 1) & 1, c2 = (color >> 2) & 1;
#pragma omp parallel for
  for (int k = c2; k < phi.n[2] + hi_ext(2, 3); k += 2)
    for (int j = c1; j < phi.n[1] + hi_ext(1, 3); j += 2)
      for (int i = c0; i < phi.n[0] + hi_ext(0, 3); i += 2) {
        if (nodal_masked(bc, phi.n, i, j, k)) continue;
        double dg; const double y = nodal_ax(q, sig, phi, i, j, k, dg);
        phi(i, j, k) += (rhs(i, j, k) - y) / dg;
      }
}

// FE divergence of a piecewise-constant velocity (NodalProjector computeRHS / mlndlap_divu).  Beyond a Neumann / inflow
// side the divergence does not see the TANGENTIAL velocities of the ghost cells (the zero_* factors of mlndlap_divu);
// the normal component of the ghost cell is used as it is (zero at walls, the inflow value at inflow faces:
// Projection::set_boundary_velocity, Projection.cpp:2570-2663).  Then mlndlap_impose_neumann_bc doubles the boundary rows.
void nodal_divu(const double dxinv[3], Arr& vel, Arr& rhs, const NodalBC& bc = PERIODIC_NBC) {
  vel.fill_periodic();
  const int* n = rhs.n;
  FOR_NODES(rhs, i, j, k) {
    if (nodal_masked(bc, n, i, j, k)) { rhs(i, j, k) = 0.0; continue; }
    const int idx[3] = {i, j, k};
    double r = 0.0;
    for (int cz = 0; cz < 2; ++cz) for (int cy = 0; cy < 2; ++cy) for (int cx = 0; cx < 2; ++cx) {
      const int off[3] = {cx, cy, cz};
      const int ci = i - cx, cj = j - cy, ck = k - cz;
      // is this cell outside a Neumann / inflow side in direction d?
      bool outd[3];
      for (int d = 0; d < 3; ++d) outd[d] = (idx[d] == 0 && off[d] == 1 && bc.neu(d, -1)) || (idx[d] == n[d] && off[d] == 0 && bc.neu(d, 1));
      for (int c = 0; c < 3; ++c) {
        bool tang_hidden = false;
        for (int d = 0; d < 3; ++d) if (d != c && outd[d]) tang_hidden = true;
        if (tang_hidden) continue;
        r += 0.25 * (off[c] ? -1.0 : 1.0) * dxinv[c] * vel(ci, cj, ck, c);
      }
    }
    double fac = 1.0;
    for (int d = 0; d < 3; ++d) if ((idx[d] == 0 && bc.neu(d, -1)) || (idx[d] == n[d] && bc.neu(d, 1))) fac *= 2.0;
    rhs(i, j, k) = fac * r;
  }
}

// cell-centred gradient of nodal phi (mlndlap_mknewu / compGrad): mean of the 4 parallel edge differences
void nodal_grad(const double dxinv[3], Arr& phi, Arr& g) {
  phi.fill_periodic();
  FOR_CELLS(g, i, j, k) {
    double gx = 0, gy = 0, gz = 0;
    for (int b = 0; b < 2; ++b) for (int a = 0; a < 2; ++a) {
      gx += phi(i + 1, j + a, k + b) - phi(i, j + a, k + b);
      gy += phi(i + a, j + 1, k + b) - phi(i + a, j, k + b);
      gz += phi(i + a, j + b, k + 1) - phi(i + a, j + b, k);
    }
    g(i, j, k, 0) = 0.25 * dxinv[0] * gx; g(i, j, k, 1) = 0.25 * dxinv[1] * gy; g(i, j, k, 2) = 0.25 * dxinv[2] * gz;
  }
}

struct NodeMG {
  struct Lev { int n[3]; double dxinv[3]; Arr sig, cor, res, rescor; };
  std::vector<Lev> lv;
  orc_mg mg;
  NodalBC bc;
  NodeMG(const int n[3], const double dx[3], int max_coarsening, const NodalBC& bc_ = PERIODIC_NBC) : bc(bc_) {
    orc_mg_default(&mg);
    int cur[3] = {n[0], n[1], n[2]}; double h[3] = {dx[0], dx[1], dx[2]};
    for (int l = 0; l <= max_coarsening; ++l) {
      lv.emplace_back();
      Lev& L = lv.back();
      for (int d = 0; d < 3; ++d) { L.n[d] = cur[d]; L.dxinv[d] = 1.0 / h[d]; }
      L.sig.define(cur, 1, 1); L.cor.define(cur, 1, 2); L.res.define(cur, 1, 2); L.rescor.define(cur, 1, 2);
      bool ok = true;
      for (int d = 0; d < 3; ++d) if (cur[d] % 2 != 0 || cur[d] / 2 < 2) ok = false;
      if (!ok) break;
      for (int d = 0; d < 3; ++d) { cur[d] /= 2; h[d] *= 2.0; }
    }
  }
  bool singular() const { for (int d = 0; d < 3; ++d) if (bc.dir(d, -1) || bc.dir(d, 1)) return false; return true; }
  static double norminf_nodes(const Arr& a) {
    double m = 0.0;
    const int n0 = a.n[0] + hi_ext(0, 3), n1 = a.n[1] + hi_ext(1, 3), n2 = a.n[2] + hi_ext(2, 3);
#pragma omp parallel for reduction(max : m)
    for (int k = 0; k < n2; ++k) for (int j = 0; j < n1; ++j) for (int i = 0; i < n0; ++i)
      m = std::max(m, std::fabs(a(i, j, k)));
    return m;
  }
  void set_sigma(const Arr& s) {
    lv[0].sig.copy_from(s, 0, 0, 1); nodal_fill_sigma(lv[0].sig, bc);
    for (size_t l = 1; l < lv.size(); ++l) {  // coarse sigma = mean of the 8 children
      Arr& c = lv[l].sig; const Arr& f = lv[l - 1].sig;
      FOR_CELLS(c, i, j, k) {
        double t = 0; for (int dk = 0; dk < 2; ++dk) for (int dj = 0; dj < 2; ++dj) for (int di = 0; di < 2; ++di) t += f(2 * i + di, 2 * j + dj, 2 * k + dk);
        c(i, j, k) = 0.125 * t;
      }
      nodal_fill_sigma(c, bc);
    }
  }
  void smooth(int l, Arr& phi, const Arr& rhs, int ns) {
    for (int s = 0; s < ns; ++s) for (int c = 0; c < 8; ++c) nodal_gs(lv[l].dxinv, lv[l].sig, rhs, c, phi, bc);
  }
  void residual(int l, Arr& out, Arr& phi, const Arr& rhs) {
    nodal_adotx(lv[l].dxinv, lv[l].sig, phi, out, bc);
    const int* n = out.n;
    FOR_NODES(out, i, j, k) out(i, j, k) = nodal_masked(bc, n, i, j, k) ? 0.0 : rhs(i, j, k) - out(i, j, k);
  }
  void vcycle() {
    const int nl = (int)lv.size();
    for (int l = 0; l < nl - 1; ++l) {
      Lev& L = lv[l];
      L.cor.setval(0.0);
      smooth(l, L.cor, L.res, mg.nu1);
      residual(l, L.rescor, L.cor, L.res);
      nodal_fill_phi(L.rescor, bc);   // MLNodeLaplacian::restriction: applyBC on the fine residual (periodic images, Neumann reflection)
      Arr& cr = lv[l + 1].res; const Arr& fr = L.rescor;
      const int* cn = cr.n;
      FOR_NODES(cr, i, j, k) {  // full weighting (1,2,1)^3 / 64
        double s = 0;
        for (int dk = -1; dk <= 1; ++dk) for (int dj = -1; dj <= 1; ++dj) for (int di = -1; di <= 1; ++di)
          s += (double)((di ? 1 : 2) * (dj ? 1 : 2) * (dk ? 1 : 2)) * fr(2 * i + di, 2 * j + dj, 2 * k + dk);
        cr(i, j, k) = nodal_masked(bc, cn, i, j, k) ? 0.0 : s / 64.0;
      }
    }
    Lev& B = lv[nl - 1];
    B.cor.setval(0.0);
    const int nsm = nl == 1 ? mg.nu1 + mg.nu2 : mg.bottom_sweeps;
    if (mg.bottom_solver == 1 && nl > 1) {
      int its = 0;
      const int* bn = B.cor.n;
      if (singular()) make_solvable(B.res);   // MLMG::bottomSolve: a singular bottom problem is made solvable first
      const int ret = bicgstab([&](Arr& out, Arr& in) {
                                 nodal_adotx(lv[nl - 1].dxinv, lv[nl - 1].sig, in, out, bc);
                                 FOR_NODES(out, i, j, k) if (nodal_masked(bc, bn, i, j, k)) out(i, j, k) = 0.0;
                               }, B.cor, B.res, 1, true, mg.bottom_rtol, mg.bottom_maxiter, &its);
      mg.bottom_iters += its;
      if (ret != 0) { B.cor.setval(0.0); smooth(nl - 1, B.cor, B.res, nsm); }
    } else {
      smooth(nl - 1, B.cor, B.res, nsm);
    }
    for (int l = nl - 2; l >= 0; --l) {
      Arr& fc = lv[l].cor; Arr& cc = lv[l + 1].cor;
      cc.fill_periodic();
      const int* fn = fc.n;
      FOR_NODES(fc, i, j, k) {  // trilinear interpolation (Dirichlet nodes stay zero)
        if (nodal_masked(bc, fn, i, j, k)) continue;
        const int ic = i / 2, jc = j / 2, kc = k / 2, ox = i & 1, oy = j & 1, oz = k & 1;
        double s = 0;
        for (int dk = 0; dk <= oz; ++dk) for (int dj = 0; dj <= oy; ++dj) for (int di = 0; di <= ox; ++di) s += cc(ic + di, jc + dj, kc + dk);
        fc(i, j, k) += s / (double)((1 + ox) * (1 + oy) * (1 + oz));
      }
      smooth(l, fc, lv[l].res, mg.nu2);
    }
  }
  // periodic / Neumann everywhere: make a right-hand side solvable (getSolvabilityOffset / fixSolvabilityByOffset: mean weighted
  // with the dot mask, 1/2 per Neumann side, subtracted from every row)
  void make_solvable(Arr& rhs) {
    const int* n = rhs.n;
    double s1 = 0.0, s2 = 0.0;
    const int n0 = n[0] + hi_ext(0, 3), n1 = n[1] + hi_ext(1, 3), n2 = n[2] + hi_ext(2, 3);
#pragma omp parallel for reduction(+ : s1, s2)
    for (int k = 0; k < n2; ++k) for (int j = 0; j < n1; ++j) for (int i = 0; i < n0; ++i) {
      const double w = nodal_weight(bc, n, i, j, k);
      s1 += w * rhs(i, j, k); s2 += w;
    }
    const double mean = s1 / s2;
    FOR_NODES(rhs, i, j, k) rhs(i, j, k) -= mean;
  }
  int solve(Arr& phi, Arr& rhs) {
    if (singular()) make_solvable(rhs);
    const double rhsnorm = norminf_nodes(rhs);
    Arr& res = lv[0].res;
    residual(0, res, phi, rhs);
    const double r0 = norminf_nodes(res);
    const double target = std::max(mg.atol, mg.rtol * std::max(rhsnorm, r0));
    double r = r0; int it = 0; int rc = 0;
    if (!(r0 <= target)) {
      rc = mg.max_iter;
      for (it = 1; it <= mg.max_iter; ++it) {
        vcycle();
        Arr& cor = lv[0].cor;
        FOR_NODES(phi, i, j, k) phi(i, j, k) += cor(i, j, k);
        residual(0, res, phi, rhs);
        r = norminf_nodes(res);
        if (r <= target) { rc = 0; break; }
      }
    }
    nodal_fill_phi(phi, bc);
    mg.iters = it; mg.resnorm0 = r0; mg.resnorm = r; mg.rhsnorm = rhsnorm;
    return rc;
  }
};

// ===========================================================================
// Godunov PLM (AMReX-Hydro hydro_godunov_plm.cpp, hydro_godunov_edge_state_3D.cpp,
// hydro_godunov_extrap_vel_to_faces_3D.cpp, hydro_godunov_corner_couple.H,
// AMReX_Slopes_K.H; A.2-A.5), periodic boundaries only.
// ===========================================================================
// UNVERIFIED-UPSTREAM switches (orc_set_option; the library's iamrx_set_option mirrors them): details of AMReX-Hydro's Godunov
// restated from memory of the upstream sources, kept switchable
double small_vel = 1.0e-8;   // |u| below which a face velocity counts as zero in the upwinding
int opt_slope_order = 4;     // limited slopes of the PLM trace: 4th order (Godunov default) or 2nd (monotonised central)
int opt_corner_adv = 0;      // corner coupling of non-conservative states: 0 flux form minus q div u, 1 advective form
int opt_extdir_both = 0;     // ext_dir faces: 1 = both traced states take the boundary value

inline double limited2(double dl, double dr) {  // 2nd-order monotonised central difference
  const double dc = 0.5 * (dl + dr);
  const double lim = (dl * dr >= 0.0) ? 2.0 * std::min(std::fabs(dl), std::fabs(dr)) : 0.0;
  return std::copysign(1.0, dc) * std::min(lim, std::fabs(dc));
}
// amrex_calc_{x,y,z}slope, order 4: q values at i-2..i+2 along the direction
inline double slope_order4(double qmm, double qm, double q0, double qp, double qpp) {
  if (opt_slope_order == 2) return limited2(q0 - qm, qp - q0);
  const double dfm = limited2(qm - qmm, q0 - qm);
  const double dfp = limited2(qp - q0, qpp - qp);
  const double dl = q0 - qm, dr = qp - q0, dc = 0.5 * (dl + dr);
  const double lim = (dl * dr >= 0.0) ? 2.0 * std::min(std::fabs(dl), std::fabs(dr)) : 0.0;
  const double d4 = 4.0 / 3.0 * dc - 1.0 / 6.0 * (dfp + dfm);
  return std::copysign(1.0, dc) * std::min(lim, std::fabs(d4));
}
inline int e0(int d) { return d == 0; }
inline int e1(int d) { return d == 1; }
inline int e2(int d) { return d == 2; }

inline double upwind_by(double lo, double hi, double vel) {  // ComputeEdgeState-style upwinding
  const double st = (vel >= 0.0) ? lo : hi;
  const double fu = (std::fabs(vel) < small_vel) ? 0.0 : 1.0;
  return fu * st + (1.0 - fu) * 0.5 * (hi + lo);
}
inline double riemann_self(double lo, double hi) {  // normal velocity upwinded by itself
  const double st = ((lo + hi) >= 0.0) ? lo : hi;
  const bool ltm = ((lo <= 0.0 && hi >= 0.0) || (std::fabs(lo + hi) < small_vel));
  return ltm ? 0.0 : st;
}

// loop over faces of direction-independent "grown by 1" range (all arrays have >= 1 ghost)
#define FOR_G1(A, i, j, k)                                        \
  _Pragma("omp parallel for") for (int k = -1; k <= (A).n[2]; ++k) \
    for (int j = -1; j <= (A).n[1]; ++j)                          \
      for (int i = -1; i <= (A).n[0]; ++i)

// godunov.use_forces_in_trans, advection_scheme == Godunov_PPM (NSB.cpp:4485); bc = BCRec of every advected component
// (null: interior / periodic everywhere), is_velocity as in the ComputeFluxesOnBoxFromState call (NSB.cpp:4701-4717)
struct AdvOpt { bool fit; bool ppm; const BCRec* bc = nullptr; bool is_velocity = false; };

// PPM (AMReX-Hydro hydro_godunov_ppm.H, van Leer limited edges + Colella-Woodward monotonisation): parabola of one cell from
// the five values along the direction; Im / Ip are its averages over the domain of dependence of the lower / upper face
inline double vanleer(double s0, double sp1, double sm1) {
  const double dsc = 0.5 * (sp1 - sm1), dsl = 2.0 * (s0 - sm1), dsr = 2.0 * (sp1 - s0);
  return (dsl * dsr > 0.0) ? std::copysign(1.0, dsc) * std::min(std::fabs(dsc), std::min(std::fabs(dsl), std::fabs(dsr))) : 0.0;
}
inline void ppm_parabola(double sm2, double sm1, double s0, double sp1, double sp2, double& sm, double& sp) {
  const double d0 = vanleer(s0, sp1, sm1), dm = vanleer(sm1, s0, sm2), dp = vanleer(sp1, sp2, s0);
  sm = 0.5 * (s0 + sm1) - (1.0 / 6.0) * (d0 - dm);
  sm = std::min(std::max(sm, std::min(s0, sm1)), std::max(s0, sm1));
  sp = 0.5 * (sp1 + s0) - (1.0 / 6.0) * (dp - d0);
  sp = std::min(std::max(sp, std::min(s0, sp1)), std::max(s0, sp1));
  if ((sp - s0) * (s0 - sm) <= 0.0) { sm = s0; sp = s0; }
  else if (std::fabs(sp - s0) >= 2.0 * std::fabs(sm - s0)) sp = 3.0 * s0 - 2.0 * sm;
  else if (std::fabs(sm - s0) >= 2.0 * std::fabs(sp - s0)) sm = 3.0 * s0 - 2.0 * sp;
}
inline double ppm_ip(double s0, double sm, double sp, double v, double dtdx) {   // state sent to the UPPER face
  if (!(v > small_vel)) return s0;
  const double sg = std::fabs(v) * dtdx, s6 = 6.0 * s0 - 3.0 * (sm + sp);
  return sp - 0.5 * sg * ((sp - sm) - (1.0 - (2.0 / 3.0) * sg) * s6);
}
inline double ppm_im(double s0, double sm, double sp, double v, double dtdx) {   // state sent to the LOWER face
  if (!(v < -small_vel)) return s0;
  const double sg = std::fabs(v) * dtdx, s6 = 6.0 * s0 - 3.0 * (sm + sp);
  return sm + 0.5 * sg * ((sp - sm) + (1.0 - (2.0 / 3.0) * sg) * s6);
}

// ---- physical boundaries of the Godunov states (AMReX_Slopes_K.H amrex_calc_*slope_extdir, hydro_godunov_ppm.H SetXBCs,
// hydro_bcs_K.H Set{X,Y,Z}EdgeBCs) ------------------------------------------------------------------------------------
inline bool ed_or_ho(int code) { return code == BC_EXT_DIR || code == BC_HOEXTRAP; }
inline double one_sided_lim(double dl2, double dr2, double val) {   // dl2, dr2 = 2 x the one-sided differences
  const double lim = (dl2 * dr2 >= 0.0) ? std::min(std::fabs(dl2), std::fabs(dr2)) : 0.0;
  return std::copysign(1.0, val) * std::min(lim, std::fabs(val));
}
// 4th-order limited slope of cell `idx` (0..nd-1 inside the domain) along a direction whose low / high side is an
// ext_dir or hoextrap boundary: the ghost value sits ON the face, so the first cell uses the one-sided 4-point
// difference and the second cell's formula takes that one-sided slope for its neighbour
inline double slope_order4_bc(double qmm, double qm, double q0, double qp, double qpp, int idx, int nd, bool edlo, bool edhi) {
  if (opt_slope_order == 2) {   // A.2: dc = (q(i+1) + 3 q(i) - 4 q_wall)/3 in the first cell
    if (edlo && idx == 0) return one_sided_lim(2.0 * (q0 - qm), 2.0 * (qp - q0), (qp + 3.0 * q0 - 4.0 * qm) / 3.0);
    if (edhi && idx == nd - 1) return one_sided_lim(2.0 * (q0 - qm), 2.0 * (qp - q0), -(qm + 3.0 * q0 - 4.0 * qp) / 3.0);
    return limited2(q0 - qm, qp - q0);
  }
  double dfm = limited2(qm - qmm, q0 - qm), dfp = limited2(qp - q0, qpp - qp);
  const double dl = q0 - qm, dr = qp - q0, dc = 0.5 * (dl + dr);
  const double lim = (dl * dr >= 0.0) ? 2.0 * std::min(std::fabs(dl), std::fabs(dr)) : 0.0;
  auto combine = [&]() { return std::copysign(1.0, dc) * std::min(lim, std::fabs(4.0 / 3.0 * dc - 1.0 / 6.0 * (dfp + dfm))); };
  double sl = combine();
  if (edlo && idx == 0) sl = one_sided_lim(2.0 * (q0 - qm), 2.0 * (qp - q0), -16.0 / 15.0 * qm + 0.5 * q0 + 2.0 / 3.0 * qp - 0.1 * qpp);
  else if (edlo && idx == 1) { dfm = one_sided_lim(2.0 * (qm - qmm), 2.0 * (q0 - qm), -16.0 / 15.0 * qmm + 0.5 * qm + 2.0 / 3.0 * q0 - 0.1 * qp); sl = combine(); }
  if (edhi && idx == nd - 1) sl = one_sided_lim(2.0 * (q0 - qm), 2.0 * (qp - q0), 16.0 / 15.0 * qp - 0.5 * q0 - 2.0 / 3.0 * qm + 0.1 * qmm);
  else if (edhi && idx == nd - 2) { dfp = one_sided_lim(2.0 * (qp - q0), 2.0 * (qpp - qp), 16.0 / 15.0 * qpp - 0.5 * qp - 2.0 / 3.0 * q0 + 0.1 * qm); sl = combine(); }
  return sl;
}
inline double clampv(double v, double a, double b) { return std::min(std::max(v, std::min(a, b)), std::max(a, b)); }
inline void ppm_monotone(double s0, double& sm, double& sp) {
  if ((sp - s0) * (s0 - sm) <= 0.0) { sm = s0; sp = s0; }
  else if (std::fabs(sp - s0) >= 2.0 * std::fabs(sm - s0)) sp = 3.0 * s0 - 2.0 * sm;
  else if (std::fabs(sm - s0) >= 2.0 * std::fabs(sp - s0)) sm = 3.0 * s0 - 2.0 * sp;
}
inline void ppm_parabola_bc(double sm2, double sm1, double s0, double sp1, double sp2, int idx, int nd, bool edlo, bool edhi, double& sm, double& sp) {
  ppm_parabola(sm2, sm1, s0, sp1, sp2, sm, sp);
  if (edlo && idx == 0) { sp = clampv(-0.2 * sm1 + 0.75 * s0 + 0.5 * sp1 - 0.05 * sp2, sp1, s0); sm = sm1; }
  else if (edlo && idx == 1) {
    sm = clampv(-0.2 * sm2 + 0.75 * sm1 + 0.5 * s0 - 0.05 * sp1, s0, sm1);
    sp = clampv(0.5 * (sp1 + s0) - (1.0 / 6.0) * (vanleer(sp1, sp2, s0) - vanleer(s0, sp1, sm1)), s0, sp1);
    ppm_monotone(s0, sm, sp);
  }
  if (edhi && idx == nd - 1) { sm = clampv(-0.2 * sp1 + 0.75 * s0 + 0.5 * sm1 - 0.05 * sm2, sm1, s0); sp = sp1; }
  else if (edhi && idx == nd - 2) {
    sp = clampv(-0.2 * sp2 + 0.75 * sp1 + 0.5 * s0 - 0.05 * sm1, s0, sp1);
    sm = clampv(0.5 * (s0 + sm1) - (1.0 / 6.0) * (vanleer(s0, sp1, sm1) - vanleer(sm1, s0, sm2)), s0, sm1);
    ppm_monotone(s0, sm, sp);
  }
}
// SetEdgeBCs on the pair (lo, hi) of the face with index f (0 = low domain face, nd = high domain face) along d;
// qb / qa: the cell values below / above the face (the ghost value holds the face value at ext_dir sides)
inline void edge_bcs(double& lo, double& hi, double qb, double qa, int f, int nd, int d, const BCRec& bc, bool normal_vel) {
  if (g_per[d]) return;
  if (f == 0) {
    const int c = bc.lo[d];
    if (c == BC_EXT_DIR) { lo = qb; if (normal_vel || opt_extdir_both) hi = lo; }
    else if (c == BC_FOEXTRAP || c == BC_HOEXTRAP || c == BC_REFLECT_EVEN) lo = hi;
    else if (c == BC_REFLECT_ODD) { lo = 0.0; hi = 0.0; }
  } else if (f == nd) {
    const int c = bc.hi[d];
    if (c == BC_EXT_DIR) { hi = qa; if (normal_vel || opt_extdir_both) lo = hi; }
    else if (c == BC_FOEXTRAP || c == BC_HOEXTRAP || c == BC_REFLECT_EVEN) hi = lo;
    else if (c == BC_REFLECT_ODD) { lo = 0.0; hi = 0.0; }
  }
}
// the outflow rule of the FINAL states: through a foextrap / hoextrap face nothing is advected into the domain
inline void outflow_rule(double& lo, double& hi, int f, int nd, int d, const BCRec& bc, bool clip_lo, bool clip_hi) {
  if (g_per[d]) return;
  if (f == 0 && (bc.lo[d] == BC_FOEXTRAP || bc.lo[d] == BC_HOEXTRAP)) { if (clip_lo) hi = std::min(hi, 0.0); lo = hi; }
  if (f == nd && (bc.hi[d] == BC_FOEXTRAP || bc.hi[d] == BC_HOEXTRAP)) { if (clip_hi) lo = std::max(lo, 0.0); hi = lo; }
}
inline int comp_of(int d, int i, int j, int k) { return d == 0 ? i : (d == 1 ? j : k); }

// traced states on d-faces for component c: lo = Ip(cell below), hi = Im(cell above); PLM or PPM.
// trace velocity: umac on the face (edge state) or the cell-centred normal velocity (vel prediction).
void trace_lohi(bool ppm, const Arr& q, int c, int d, double dtdx, const Arr* mac, const Arr* vcc, const Arr* f, int fc, bool fit, double dt,
                Arr& lo, Arr& hi, const BCRec* bc, bool normal_vel) {
  const int nd = q.n[d];
  const bool edlo = bc && !g_per[d] && ed_or_ho(bc->lo[d]), edhi = bc && !g_per[d] && ed_or_ho(bc->hi[d]);
  Arr s(q.n, 1, 1), pm, pp;   // PLM slopes, or PPM parabola edges, of cells -1..n
  if (ppm) { pm.define(q.n, 1, 1); pp.define(q.n, 1, 1); }
#pragma omp parallel for
  for (int k = -1; k <= q.n[2]; ++k)
    for (int j = -1; j <= q.n[1]; ++j)
      for (int i = -1; i <= q.n[0]; ++i) {
        const double qmm = q(i - 2 * e0(d), j - 2 * e1(d), k - 2 * e2(d), c), qm = q(i - e0(d), j - e1(d), k - e2(d), c), q0 = q(i, j, k, c),
                     qp = q(i + e0(d), j + e1(d), k + e2(d), c), qpp = q(i + 2 * e0(d), j + 2 * e1(d), k + 2 * e2(d), c);
        const int idx = comp_of(d, i, j, k);
        if (ppm) {
          if (edlo || edhi) ppm_parabola_bc(qmm, qm, q0, qp, qpp, idx, nd, edlo, edhi, pm(i, j, k), pp(i, j, k));
          else ppm_parabola(qmm, qm, q0, qp, qpp, pm(i, j, k), pp(i, j, k));
        } else {
          s(i, j, k) = (edlo || edhi) ? slope_order4_bc(qmm, qm, q0, qp, qpp, idx, nd, edlo, edhi) : slope_order4(qmm, qm, q0, qp, qpp);
        }
      }
  // faces 0..n along d, -1..n in the transverse directions
#pragma omp parallel for
  for (int k = -1 + e2(d); k <= q.n[2]; ++k)
    for (int j = -1 + e1(d); j <= q.n[1]; ++j)
      for (int i = -1 + e0(d); i <= q.n[0]; ++i) {
    const int im = i - e0(d), jm = j - e1(d), km = k - e2(d);
    // trace velocities: the MAC velocity of THIS face for both sides (edge states), or each cell's own velocity (prediction)
    const double ul = mac ? (*mac)(i, j, k) : (*vcc)(im, jm, km, d);
    const double uh = mac ? (*mac)(i, j, k) : (*vcc)(i, j, k, d);
    double l, h;
    if (ppm) {
      l = ppm_ip(q(im, jm, km, c), pm(im, jm, km), pp(im, jm, km), ul, dtdx);
      h = ppm_im(q(i, j, k, c), pm(i, j, k), pp(i, j, k), uh, dtdx);
    } else {
      l = q(im, jm, km, c) + 0.5 * (1.0 - ul * dtdx) * s(im, jm, km);
      h = q(i, j, k, c) + 0.5 * (-1.0 - uh * dtdx) * s(i, j, k);
    }
    if (fit && f) { l += 0.5 * dt * (*f)(im, jm, km, fc); h += 0.5 * dt * (*f)(i, j, k, fc); }
    if (bc) edge_bcs(l, h, q(im, jm, km, c), q(i, j, k, c), comp_of(d, i, j, k), nd, d, *bc, normal_vel);
    lo(i, j, k) = l; hi(i, j, k) = h;
  }
}

// Godunov_corner_couple_<d1><d2>: d1-face lo/hi states corrected with the d2-derivative of the
// (already upwinded) d2-edge state, boundary conditions of the d1-face re-imposed, then upwinded with the d1 face velocity
void corner_couple(const Arr& lo, const Arr& hi, int d1, int d2, double dt3dx, bool conserv, const Arr& q, int c, const Arr& mac2,
                   const Arr& edge2, const Arr& mac1, Arr& out, const BCRec* bc = nullptr, bool normal_vel = false) {
  const int n0 = q.n[0], n1 = q.n[1], n2 = q.n[2];
#pragma omp parallel for
  for (int k = -1 + e2(d2); k <= n2 - e2(d2); ++k)
    for (int j = -1 + e1(d2); j <= n1 - e1(d2); ++j)
      for (int i = -1 + e0(d2); i <= n0 - e0(d2); ++i) {
        const int im = i - e0(d1), jm = j - e1(d1), km = k - e2(d1);
        if (im < -1 || jm < -1 || km < -1) continue;
        const int ip = e0(d2), jp = e1(d2), kp = e2(d2);
        double l = lo(i, j, k) - dt3dx * (edge2(im + ip, jm + jp, km + kp) * mac2(im + ip, jm + jp, km + kp) - edge2(im, jm, km) * mac2(im, jm, km));
        double h = hi(i, j, k) - dt3dx * (edge2(i + ip, j + jp, k + kp) * mac2(i + ip, j + jp, k + kp) - edge2(i, j, k) * mac2(i, j, k));
        if (!conserv && opt_corner_adv) {   // advective form of the transverse derivative
          l = lo(i, j, k) - dt3dx * 0.5 * (mac2(im + ip, jm + jp, km + kp) + mac2(im, jm, km)) * (edge2(im + ip, jm + jp, km + kp) - edge2(im, jm, km));
          h = hi(i, j, k) - dt3dx * 0.5 * (mac2(i + ip, j + jp, k + kp) + mac2(i, j, k)) * (edge2(i + ip, j + jp, k + kp) - edge2(i, j, k));
        } else if (!conserv) {
          l += dt3dx * q(im, jm, km, c) * (mac2(im + ip, jm + jp, km + kp) - mac2(im, jm, km));
          h += dt3dx * q(i, j, k, c) * (mac2(i + ip, j + jp, k + kp) - mac2(i, j, k));
        }
        if (bc) edge_bcs(l, h, q(im, jm, km, c), q(i, j, k, c), comp_of(d1, i, j, k), q.n[d1], d1, *bc, normal_vel);
        out(i, j, k) = upwind_by(l, h, mac1(i, j, k));
      }
}

// HydroUtils::ComputeFluxesOnBoxFromState -> Godunov::ComputeEdgeState for one component:
// fills the final edge states edge[d] (1 comp) on the low faces of every cell (and the high domain face when not periodic)
void edge_state_comp(const Arr& q, int c, const Arr* f, int fc, const Arr* divu, const Arr* mac[3], bool conserv, AdvOpt opt,
                     const double dx[3], double dt, Arr* edge[3]) {
  const bool fit = opt.fit;
  const int* n = q.n;
  const BCRec* bc = opt.bc ? &opt.bc[c] : nullptr;
  Arr lo[3], hi[3], ed[3];
  for (int d = 0; d < 3; ++d) {
    lo[d].define(n, 1, 2); hi[d].define(n, 1, 2); ed[d].define(n, 1, 2);
    trace_lohi(opt.ppm, q, c, d, dt / dx[d], mac[d], nullptr, f, fc, fit, dt, lo[d], hi[d], bc, opt.is_velocity && c == d);
    Arr& E = ed[d]; const Arr& M = *mac[d]; const Arr &L = lo[d], &H = hi[d];
    FOR_G1(E, i, j, k) E(i, j, k) = upwind_by(L(i, j, k), H(i, j, k), M(i, j, k));
  }
  // six corner-coupled states cc[d1][d2] (d1-face state corrected by the d2 derivative)
  Arr cc[3][3];
  for (int d1 = 0; d1 < 3; ++d1)
    for (int d2 = 0; d2 < 3; ++d2) {
      if (d1 == d2) continue;
      cc[d1][d2].define(n, 1, 2);
      corner_couple(lo[d1], hi[d1], d1, d2, dt / (3.0 * dx[d2]), conserv, q, c, *mac[d2], ed[d2], *mac[d1], cc[d1][d2], bc, opt.is_velocity && c == d1);
    }
  for (int d = 0; d < 3; ++d) {
    const int t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;  // the two transverse directions
    // transverse d/dt1 uses the t1-face state that was corrected by the t2 derivative, and vice versa
    const Arr &A1 = cc[t1][t2], &A2 = cc[t2][t1], &M1 = *mac[t1], &M2 = *mac[t2], &M = *mac[d];
    const double dtd1 = dt / dx[t1], dtd2 = dt / dx[t2];
    const bool nv = opt.is_velocity && c == d;
    Arr& E = *edge[d];
    FOR_FACES(E, d, i, j, k) {
      double st[2];
      for (int side = 0; side < 2; ++side) {  // 0: from the cell below (stl), 1: from the cell above (sth)
        const int ci = i - (side == 0 ? e0(d) : 0), cj = j - (side == 0 ? e1(d) : 0), ck = k - (side == 0 ? e2(d) : 0);
        const int i1 = ci + e0(t1), j1 = cj + e1(t1), k1 = ck + e2(t1), i2 = ci + e0(t2), j2 = cj + e1(t2), k2 = ck + e2(t2);
        double s = (side == 0) ? lo[d](i, j, k) : hi[d](i, j, k);
        if (conserv) {
          s += -(0.5 * dtd1) * (A1(i1, j1, k1) * M1(i1, j1, k1) - A1(ci, cj, ck) * M1(ci, cj, ck))
               - (0.5 * dtd2) * (A2(i2, j2, k2) * M2(i2, j2, k2) - A2(ci, cj, ck) * M2(ci, cj, ck))
               + (0.5 * dtd1) * q(ci, cj, ck, c) * (M1(i1, j1, k1) - M1(ci, cj, ck))
               + (0.5 * dtd2) * q(ci, cj, ck, c) * (M2(i2, j2, k2) - M2(ci, cj, ck));
          if (divu) s -= 0.5 * dt * q(ci, cj, ck, c) * (*divu)(ci, cj, ck);
        } else {
          s += -(0.25 * dtd1) * (M1(i1, j1, k1) + M1(ci, cj, ck)) * (A1(i1, j1, k1) - A1(ci, cj, ck))
               - (0.25 * dtd2) * (M2(i2, j2, k2) + M2(ci, cj, ck)) * (A2(i2, j2, k2) - A2(ci, cj, ck));
        }
        if (!fit && f) s += 0.5 * dt * (*f)(ci, cj, ck, fc);
        st[side] = s;
      }
      if (bc) {
        const int fidx = comp_of(d, i, j, k);
        edge_bcs(st[0], st[1], q(i - e0(d), j - e1(d), k - e2(d), c), q(i, j, k, c), fidx, n[d], d, *bc, nv);
        outflow_rule(st[0], st[1], fidx, n[d], d, *bc, nv && M(i, j, k) >= 0.0, nv && M(i, j, k) <= 0.0);
      }
      E(i, j, k) = upwind_by(st[0], st[1], M(i, j, k));
    }
  }
}

// NavierStokesBase::ComputeAofs body, non-EB (NSB.cpp:4661-4845).  uflux: the velocities that multiply the edge states into
// fluxes (u_mac itself, or U_corr in the sync call :4672-4677); is_sync: aofs -= update and no convective term (:4784,4834);
// known: eds[] are INPUT edge states (known_edge_state :4708), only fluxes / divergence / convective term are formed.
void compute_aofs(const Arr& S, int ncomp, const Arr* force, const Arr* divu, Arr mac[3], const int* iconserv, AdvOpt fit,
                  const double dx[3], double dt, Arr& aofs, int acomp, Arr* fl[3], Arr* eds[3], Arr* uflux = nullptr,
                  bool is_sync = false, bool known = false) {
  const int* n = S.n;
  const Arr* macp[3] = {&mac[0], &mac[1], &mac[2]};
  const double area[3] = {dx[1] * dx[2], dx[0] * dx[2], dx[0] * dx[1]};
  const double vol = dx[0] * dx[1] * dx[2];
  for (int c = 0; c < ncomp; ++c) {
    Arr ed[3] = {Arr(n, 1, 1), Arr(n, 1, 1), Arr(n, 1, 1)};
    Arr* edp[3] = {&ed[0], &ed[1], &ed[2]};
    if (known) { for (int d = 0; d < 3; ++d) { Arr& E = ed[d]; const Arr& K = *eds[d]; FOR_FACES(E, d, i, j, k) E(i, j, k) = K(i, j, k, c); } }
    else edge_state_comp(S, c, force, c, divu, macp, iconserv[c] != 0, fit, dx, dt, edp);
    Arr fx[3] = {Arr(n, 1, 1), Arr(n, 1, 1), Arr(n, 1, 1)};
    for (int d = 0; d < 3; ++d) {
      Arr& F = fx[d]; const Arr& E = ed[d]; const Arr& M = uflux ? uflux[d] : mac[d];
      FOR_FACES(F, d, i, j, k) F(i, j, k) = E(i, j, k) * M(i, j, k) * area[d];  // HydroUtils::ComputeFluxes, area-weighted (NSB.cpp:4651)
      F.fill_periodic(); ed[d].fill_periodic();
    }
    const bool cons = iconserv[c] != 0;
    FOR_CELLS(aofs, i, j, k) {
      // ComputeDivergence(mult = -1, area-weighted fluxes) NSB.cpp:4753-4771
      double upd = -((fx[0](i + 1, j, k) - fx[0](i, j, k)) + (fx[1](i, j + 1, k) - fx[1](i, j, k)) + (fx[2](i, j, k + 1) - fx[2](i, j, k))) / vol;
      if (!cons && !is_sync) {  // ComputeConvectiveTerm NSB.cpp:4784,4809-4820 ("sync is always a conservative update")
        const double divum = (mac[0](i + 1, j, k) - mac[0](i, j, k)) / dx[0] + (mac[1](i, j + 1, k) - mac[1](i, j, k)) / dx[1] +
                             (mac[2](i, j, k + 1) - mac[2](i, j, k)) / dx[2];
        const double qavg = (ed[0](i, j, k) + ed[0](i + 1, j, k) + ed[1](i, j, k) + ed[1](i, j + 1, k) + ed[2](i, j, k) + ed[2](i, j, k + 1)) / 6.0;
        upd += qavg * divum;
      }
      if (is_sync) aofs(i, j, k, acomp + c) -= upd;   // NSB.cpp:4834
      else aofs(i, j, k, acomp + c) = -upd;            // NSB.cpp:4840
    }
    for (int d = 0; d < 3; ++d) {
      if (fl && fl[d]) { Arr& O = *fl[d]; const Arr& F = fx[d]; FOR_FACES(O, d, i, j, k) O(i, j, k, c) = F(i, j, k); }
      if (eds && eds[d] && !known) { Arr& O = *eds[d]; const Arr& E = ed[d]; FOR_FACES(O, d, i, j, k) O(i, j, k, c) = E(i, j, k); }
    }
  }
}

// Godunov::ExtrapVelToFaces (hydro_godunov_extrap_vel_to_faces_3D.cpp), PLM or PPM
void extrap_vel_to_faces(const Arr& vel, const Arr* f, AdvOpt opt, const double dx[3], double dt, Arr umac[3]) {
  const bool fit = opt.fit;
  const int* n = vel.n;
  // traced states of every component on every face direction (boundary conditions applied: is_velocity = true)
  Arr lo[3][3], hi[3][3];  // [dir][comp]
  for (int d = 0; d < 3; ++d)
    for (int c = 0; c < 3; ++c) {
      lo[d][c].define(n, 1, 2); hi[d][c].define(n, 1, 2);
      trace_lohi(opt.ppm, vel, c, d, dt / dx[d], nullptr, &vel, f, c, fit, dt, lo[d][c], hi[d][c], opt.bc ? &opt.bc[c] : nullptr, c == d);
    }
  // advective velocities (ComputeAdvectiveVel) and transverse edge states upwinded by them
  Arr ad[3], ed[3][3];
  for (int d = 0; d < 3; ++d) {
    ad[d].define(n, 1, 2);
    Arr& A = ad[d]; const Arr &L = lo[d][d], &H = hi[d][d];
    FOR_G1(A, i, j, k) A(i, j, k) = riemann_self(L(i, j, k), H(i, j, k));
    for (int c = 0; c < 3; ++c) {
      ed[d][c].define(n, 1, 2);
      Arr& E = ed[d][c]; const Arr &Lc = lo[d][c], &Hc = hi[d][c];
      FOR_G1(E, i, j, k) E(i, j, k) = upwind_by(Lc(i, j, k), Hc(i, j, k), A(i, j, k));
    }
  }
  for (int d = 0; d < 3; ++d) {
    const int t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
    const BCRec* bcd = opt.bc ? &opt.bc[d] : nullptr;
    Arr a1(n, 1, 2), a2(n, 1, 2);  // t1-face state of comp d corrected by t2, and t2-face state corrected by t1
    corner_couple(lo[t1][d], hi[t1][d], t1, t2, dt / (3.0 * dx[t2]), false, vel, d, ad[t2], ed[t2][d], ad[t1], a1, bcd, false);
    corner_couple(lo[t2][d], hi[t2][d], t2, t1, dt / (3.0 * dx[t1]), false, vel, d, ad[t1], ed[t1][d], ad[t2], a2, bcd, false);
    const double dtd1 = dt / dx[t1], dtd2 = dt / dx[t2];
    const Arr &M1 = ad[t1], &M2 = ad[t2];
    Arr& U = umac[d];
    FOR_FACES(U, d, i, j, k) {
      double st[2];
      for (int side = 0; side < 2; ++side) {
        const int ci = i - (side == 0 ? e0(d) : 0), cj = j - (side == 0 ? e1(d) : 0), ck = k - (side == 0 ? e2(d) : 0);
        const int i1 = ci + e0(t1), j1 = cj + e1(t1), k1 = ck + e2(t1), i2 = ci + e0(t2), j2 = cj + e1(t2), k2 = ck + e2(t2);
        double s = (side == 0) ? lo[d][d](i, j, k) : hi[d][d](i, j, k);
        s += -(0.25 * dtd1) * (M1(i1, j1, k1) + M1(ci, cj, ck)) * (a1(i1, j1, k1) - a1(ci, cj, ck))
             - (0.25 * dtd2) * (M2(i2, j2, k2) + M2(ci, cj, ck)) * (a2(i2, j2, k2) - a2(ci, cj, ck));
        if (!fit && f) s += 0.5 * dt * (*f)(ci, cj, ck, d);
        st[side] = s;
      }
      if (bcd) {
        const int fidx = comp_of(d, i, j, k);
        edge_bcs(st[0], st[1], vel(i - e0(d), j - e1(d), k - e2(d), d), vel(i, j, k, d), fidx, n[d], d, *bcd, true);
        outflow_rule(st[0], st[1], fidx, n[d], d, *bcd, true, true);
      }
      U(i, j, k) = riemann_self(st[0], st[1]);
    }
  }
}

// ===========================================================================
// level solvers
// ===========================================================================
// MacProj::mlmg_mac_solve + Hydro::MacProjector (MacProj.cpp:1084-1184).  bc (null: periodic): the LinOpBCType of every
// side from set_mac_solve_bc (:1187-1208) and mac_proj.maxorder (:30,75); phi's ghost cells hold the level BC
// (setLevelBC(0, mac_phi) :1168; mac_phi = 0 on entry :253).  rho: 1 ghost cell, filled (foextrap at walls).
int mac_project(const int n[3], const double dx[3], Arr mac[3], Arr& rho, const Arr* rhs_in, Arr& phi, double rhs_scale, orc_mg* mgp,
                const LinBC* bc = nullptr) {
  // beta = (1/rhs_scale)/avg(rho)  MacProj.cpp:1115-1127
  rho.fill_periodic();
  Arr beta[3];
  for (int d = 0; d < 3; ++d) {
    beta[d].define(n, 1, 1);
    Arr& B = beta[d];
    FOR_FACES(B, d, i, j, k) B(i, j, k) = (1.0 / rhs_scale) / (0.5 * (rho(i - e0(d), j - e1(d), k - e2(d)) + rho(i, j, k)));
    B.fill_periodic();
    mac[d].fill_periodic();
  }
  CellMG mg(n, dx, 1, false, mgp ? mgp->max_coarsening : 100);
  if (mgp) mg.mg = *mgp;
  mg.a = 0.0; mg.b = 1.0;
  Arr levelbc;
  if (bc) { levelbc = phi; mg.set_bc(*bc, &levelbc); }
  const Arr* e[3] = {&beta[0], &beta[1], &beta[2]};
  mg.set_coeffs(nullptr, e);
  Arr rhs(n, 1, 0);
  FOR_CELLS(rhs, i, j, k) {
    double dv = (mac[0](i + 1, j, k) - mac[0](i, j, k)) / dx[0] + (mac[1](i, j + 1, k) - mac[1](i, j, k)) / dx[1] +
                (mac[2](i, j, k + 1) - mac[2](i, j, k)) / dx[2];
    rhs(i, j, k) = -dv + (rhs_in ? (*rhs_in)(i, j, k) : 0.0);
  }
  const int rc = mg.solve(phi, rhs);   // ghost cells of phi are BC-filled on return (setFinalFillBC)
  if (mgp) *mgp = mg.mg;
  for (int d = 0; d < 3; ++d) {  // umac += fluxes = -beta grad phi; a Neumann face has zero flux (its ghost equals the interior cell)
    Arr& U = mac[d]; const Arr& B = beta[d];
    FOR_FACES(U, d, i, j, k) U(i, j, k) -= B(i, j, k) * (phi(i, j, k) - phi(i - e0(d), j - e1(d), k - e2(d))) / dx[d];
    U.fill_periodic();
  }
  return rc;
}

// Projection::set_boundary_velocity (Projection.cpp:2570-2663): the normal velocity in the ghost cells beyond every
// non-periodic side is zeroed unless that side is an inflow face (whose ghost cells carry the inflow velocity)
void set_boundary_velocity(Arr& vel, const NodalBC& bc) {
  for (int d = 0; d < 3; ++d) {
    if (g_per[d]) continue;
    const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
    for (int side = -1; side <= 1; side += 2) {
      if ((side < 0 ? bc.lo[d] : bc.hi[d]) == LO_INFLOW) continue;
      const int g = side < 0 ? -1 : vel.n[d];
#pragma omp parallel for
      for (int b2 = -1; b2 <= vel.n[d2]; ++b2)
        for (int b1 = -1; b1 <= vel.n[d1]; ++b1) { int q[3]; q[d] = g; q[d1] = b1; q[d2] = b2; vel(q[0], q[1], q[2], d) = 0.0; }
    }
  }
}

// Projection::doMLMGNodalProjection + Hydro::NodalProjector (Projection.cpp:2385-2567).  With a non-periodic side phi
// carries 2 ghost layers (the ghost node of a high side is index n+1) and vel's ghost cells beyond inflow sides hold
// the inflow velocity.
int nodal_project(const int n[3], const double dx[3], Arr& vel, Arr& sigma, Arr& phi, Arr* gp, bool increment, orc_mg* mgp,
                  const NodalBC& bc = PERIODIC_NBC) {
  const double dxinv[3] = {1.0 / dx[0], 1.0 / dx[1], 1.0 / dx[2]};
  NodeMG mg(n, dx, mgp ? mgp->max_coarsening : 100, bc);
  if (mgp) mg.mg = *mgp;
  mg.set_sigma(sigma);
  set_boundary_velocity(vel, bc);
  Arr rhs(n, 1, 2);
  nodal_divu(dxinv, vel, rhs, bc);
  const int rc = mg.solve(phi, rhs);
  if (mgp) *mgp = mg.mg;
  Arr g(n, 3, 0);
  nodal_grad(dxinv, phi, g);
  for (int c = 0; c < 3; ++c) {
    FOR_CELLS(g, i, j, k) {
      vel(i, j, k, c) -= sigma(i, j, k) * g(i, j, k, c);
      if (gp) { if (increment) (*gp)(i, j, k, c) += g(i, j, k, c); else (*gp)(i, j, k, c) = g(i, j, k, c); }
    }
  }
  return rc;
}

}  // namespace

// ===========================================================================
// NavierStokes::advance / post_init (single level, periodic)
// ===========================================================================
struct orc_ns {
  int n[3]; double dx[3], prob_lo[3];
  orc_ns_params p;
  enum { Xvel = 0, Density = 3, Tracer = 4, NUM_STATE = 5 };
  Arr S_old, S_new, P_old, P_new, Gp_old, Gp_new, umac[3], aofs, rho_p, rho_c, rho_half, eta[3];
  double time = 0, dt_level = 0, dt_min = 1e100; int nstep = 0;
  bool initial_step = false, initial_iter = false;
  int it[3] = {0, 0, 0};
  Arr force;  // velocity forcing from predict_velocity (1 ghost), reused by velocity_advection
  Arr seta[3];  // tracer diffusivity on faces (getDiffusivity: constant ns.scal_diff_coefs, NS.cpp:2051-2119)
  // ---- physical boundaries (ns.lo_bc / ns.hi_bc, geometry.is_periodic; NS.cpp:90-237) --------------------------------
  int per[3] = {1, 1, 1}, phys_lo[3] = {0, 0, 0}, phys_hi[3] = {0, 0, 0};
  double bcv[6][5] = {};   // Dirichlet face values per face (x lo, y lo, z lo, x hi, y hi, z hi) and state component
  bool has_walls() const { return !(per[0] && per[1] && per[2]); }
  // NS_BC.H:7-38 : physical type -> math BC of every state component / of grad p
  static int math_bc(const int* table, int phys) { return table[phys]; }
  BCRec state_bc(int comp) const {
    static const int norm_vel[6] = {BC_INT_DIR, BC_EXT_DIR, BC_FOEXTRAP, BC_REFLECT_ODD, BC_EXT_DIR, BC_EXT_DIR};
    static const int tang_vel[6] = {BC_INT_DIR, BC_EXT_DIR, BC_FOEXTRAP, BC_REFLECT_EVEN, BC_HOEXTRAP, BC_EXT_DIR};
    static const int scalar[6] = {BC_INT_DIR, BC_EXT_DIR, BC_FOEXTRAP, BC_REFLECT_EVEN, BC_FOEXTRAP, BC_FOEXTRAP};
    BCRec b;
    for (int d = 0; d < 3; ++d) {
      const int* t = comp < 3 ? (comp == d ? norm_vel : tang_vel) : scalar;
      b.lo[d] = per[d] ? BC_INT_DIR : t[phys_lo[d]]; b.hi[d] = per[d] ? BC_INT_DIR : t[phys_hi[d]];
    }
    return b;
  }
  BCRec gradp_bc(int comp) const {
    static const int norm_gp[6] = {BC_INT_DIR, BC_FOEXTRAP, BC_FOEXTRAP, BC_REFLECT_ODD, BC_FOEXTRAP, BC_FOEXTRAP};
    static const int tang_gp[6] = {BC_INT_DIR, BC_FOEXTRAP, BC_FOEXTRAP, BC_REFLECT_EVEN, BC_FOEXTRAP, BC_FOEXTRAP};
    BCRec b;
    for (int d = 0; d < 3; ++d) {
      const int* t = comp == d ? norm_gp : tang_gp;
      b.lo[d] = per[d] ? BC_INT_DIR : t[phys_lo[d]]; b.hi[d] = per[d] ? BC_INT_DIR : t[phys_hi[d]];
    }
    return b;
  }
  // AmrLevel::FillPatch of State_Type on one level: valid copy, periodic images, physical boundaries (NS_bcfill.H)
  void fillpatch(Arr& dst, const Arr& src, int scomp, int ncomp) const {
    dst.copy_from(src, scomp, 0, ncomp);
    fill_state_bc(dst, scomp, ncomp);
  }
  void fill_state_bc(Arr& a, int scomp, int ncomp) const {
    a.fill_periodic();
    if (!has_walls()) return;
    std::vector<BCRec> b(ncomp); std::vector<double> v(6 * ncomp);
    for (int c = 0; c < ncomp; ++c) { b[c] = state_bc(scomp + c); for (int f = 0; f < 6; ++f) v[f * ncomp + c] = bcv[f][scomp + c]; }
    fill_physbc(a, 0, ncomp, b.data(), v.data());
  }
  void fill_gradp(Arr& gp) const {   // FillPatch of Gradp_Type (NS_setup.cpp:339-360)
    gp.fill_periodic();
    if (!has_walls()) return;
    BCRec b[3] = {gradp_bc(0), gradp_bc(1), gradp_bc(2)};
    fill_physbc(gp, 0, 3, b, nullptr);
  }
  static int linop_of(int math) {   // Diffusion::setDomainBC (Diffusion.cpp:1887-1999)
    return math == BC_EXT_DIR ? LO_DIRICHLET : (math == BC_REFLECT_ODD ? LO_REFLECT_ODD : (math == BC_INT_DIR ? LO_PERIODIC : LO_NEUMANN));
  }
  LinBC diff_bc(int scomp, int ncomp, int maxorder = 2) const {   // diffuse.max_order / tensor_max_order = 2 (Diffusion.cpp:95-96)
    LinBC L; L.maxorder = maxorder;
    for (int c = 0; c < ncomp && c < 3; ++c) { const BCRec b = state_bc(scomp + c); for (int d = 0; d < 3; ++d) { L.lo[c][d] = linop_of(b.lo[d]); L.hi[c][d] = linop_of(b.hi[d]); } }
    return L;
  }
  LinBC mac_bc() const {   // set_mac_solve_bc (MacProj.cpp:1187-1208), mac_proj.maxorder = 4 (MacProj.cpp:30)
    LinBC L; L.maxorder = 4;
    for (int d = 0; d < 3; ++d) {
      L.lo[0][d] = per[d] ? LO_PERIODIC : (phys_lo[d] == PHYS_OUTFLOW ? LO_DIRICHLET : LO_NEUMANN);
      L.hi[0][d] = per[d] ? LO_PERIODIC : (phys_hi[d] == PHYS_OUTFLOW ? LO_DIRICHLET : LO_NEUMANN);
    }
    return L;
  }
  NodalBC nodal_bc() const {   // Projection.cpp:2436-2464
    NodalBC B;
    for (int d = 0; d < 3; ++d) {
      B.lo[d] = per[d] ? LO_PERIODIC : (phys_lo[d] == PHYS_OUTFLOW ? LO_DIRICHLET : (phys_lo[d] == PHYS_INFLOW ? LO_INFLOW : LO_NEUMANN));
      B.hi[d] = per[d] ? LO_PERIODIC : (phys_hi[d] == PHYS_OUTFLOW ? LO_DIRICHLET : (phys_hi[d] == PHYS_INFLOW ? LO_INFLOW : LO_NEUMANN));
    }
    return B;
  }
  std::vector<BCRec> adv_bc(int scomp, int ncomp) const { std::vector<BCRec> b(ncomp); for (int c = 0; c < ncomp; ++c) b[c] = state_bc(scomp + c); return b; }

  orc_mg mg(double rtol, double atol) const { orc_mg m; orc_mg_default(&m); m.rtol = rtol; m.atol = atol; m.bottom_solver = p.bottom_solver; return m; }
  bool diffusive() const { return p.visc_coef > 0.0; }
  bool diffusive_tracer() const { return p.scal_diff_coef > 0.0; }   // is_diffusive[Tracer], NS_setup.cpp:292-295
  int rho_flag() const { return p.conservative_tracer ? 2 : 0; }      // Diffusion::set_rho_flag of diffusionType[Tracer], NS_setup.cpp:304-308

  // getViscTerms -> getTensorViscTerms (Diffusion.cpp:1655-1777): a = 0, b = -1
  void visc_terms(const Arr& S, Arr& visc) {
    if (!diffusive()) { visc.setval(0.0); return; }
    Arr u(n, 3, 1); fillpatch(u, S, Xvel, 3);     // Diffusion.cpp:1745: FillPatch'd Soln is the level BC
    Arr lbc = u;
    CellMG op(n, dx, 3, true, 0);
    op.a = 0.0; op.b = -1.0;
    if (has_walls()) op.set_bc(diff_bc(Xvel, 3), &lbc);
    const Arr* e[3] = {&eta[0], &eta[1], &eta[2]};
    op.set_coeffs(nullptr, e);
    op.apply(visc, u);
    visc.fill_periodic();
    if (has_walls()) first_order_extrap(visc, 0, 3);   // NS.cpp:2045-2046
  }
  // ---- HIT turbulent forcing, the exact (all modes in every cell) path of Tutorials/HIT/NS_getForce.cpp:541-621 ------------
  bool turb_on = false;
  int turb_nmodes = 0, turb_mode_start = 0, turb_div_free = 1, turb_as = 33;
  std::vector<double> turb_fd;                  // TurbulentForcing::forcedata: 17 arrays of as^3, kx fastest
  std::vector<std::array<int, 3>> turb_k;       // the modes the two loops of :555-557 and :616-618 visit with kappa <= kappaMax
  double turb_L[3] = {1, 1, 1};
  double fd(int arr, int kx, int ky, int kz) const {
    const size_t ne = (size_t)turb_as * turb_as * turb_as;
    return turb_fd[(size_t)arr * ne + kx + (size_t)turb_as * (ky + (size_t)turb_as * kz)];
  }
  void set_turb(int nmodes, int mode_start, int div_free, int as, const double* data) {
    turb_on = data != nullptr && nmodes > 0;
    if (!turb_on) return;
    turb_nmodes = nmodes; turb_mode_start = mode_start; turb_div_free = div_free; turb_as = as;
    turb_fd.assign(data, data + (size_t)17 * as * as * as);
    for (int d = 0; d < 3; ++d) turb_L[d] = n[d] * dx[d];
    const double Lx = turb_L[0], Ly = turb_L[1], Lz = turb_L[2], Lmin = std::min(Lx, std::min(Ly, Lz));
    const int xstep = (int)(Lx / Lmin + 0.5), ystep = (int)(Ly / Lmin + 0.5), zstep = (int)(Lz / Lmin + 0.5);
    const double kappaMax = nmodes / Lmin + 1.0e-8;
    turb_k.clear();
    auto consider = [&](int kx, int ky, int kz) {
      const double kappa = std::sqrt((kx * kx) / (Lx * Lx) + (ky * ky) / (Ly * Ly) + (kz * kz) / (Lz * Lz));
      if (kappa <= kappaMax) turb_k.push_back({kx, ky, kz});
    };
    for (int kz = mode_start * zstep; kz <= nmodes * zstep; kz += zstep)
      for (int ky = mode_start * ystep; ky <= nmodes * ystep; ky += ystep)
        for (int kx = mode_start * xstep; kx <= nmodes * xstep; kx += xstep) consider(kx, ky, kz);
    for (int kz = 1; kz <= zstep - 1; ++kz)      // high aspect ratio: extra symmetry-breaking modes
      for (int ky = mode_start; ky <= nmodes * ystep; ++ky)
        for (int kx = mode_start; kx <= nmodes * xstep; ++kx) consider(kx, ky, kz);
  }
  double turb(int c, int i, int j, int k, double t) const {   // component c of f at the centre of cell (i, j, k), per unit mass
    const double TwoPi = 2.0 * M_PI, Lx = turb_L[0], Ly = turb_L[1], Lz = turb_L[2];
    const double x = prob_lo[0] + dx[0] * (i + 0.5), y = prob_lo[1] + dx[1] * (j + 0.5), z = prob_lo[2] + dx[2] * (k + 0.5);
    double f = 0.0;
    for (const auto& m : turb_k) {
      const int kx = m[0], ky = m[1], kz = m[2];
      const double xT = std::cos(fd(0, kx, ky, kz) * t + fd(1, kx, ky, kz));
      const double FAX = fd(5, kx, ky, kz), FAY = fd(6, kx, ky, kz), FAZ = fd(7, kx, ky, kz);
      if (turb_div_free) {
        const double FPXX = fd(8, kx, ky, kz), FPXY = fd(9, kx, ky, kz), FPXZ = fd(10, kx, ky, kz);
        const double FPYX = fd(11, kx, ky, kz), FPYY = fd(12, kx, ky, kz), FPYZ = fd(13, kx, ky, kz);
        const double FPZX = fd(14, kx, ky, kz), FPZY = fd(15, kx, ky, kz), FPZZ = fd(16, kx, ky, kz);
        if (c == 0)
          f += xT * (FAZ * TwoPi * (ky / Ly) * std::sin(TwoPi * kx * x / Lx + FPZX) * std::cos(TwoPi * ky * y / Ly + FPZY) * std::sin(TwoPi * kz * z / Lz + FPZZ)
                     - FAY * TwoPi * (kz / Lz) * std::sin(TwoPi * kx * x / Lx + FPYX) * std::sin(TwoPi * ky * y / Ly + FPYY) * std::cos(TwoPi * kz * z / Lz + FPYZ));
        else if (c == 1)
          f += xT * (FAX * TwoPi * (kz / Lz) * std::sin(TwoPi * kx * x / Lx + FPXX) * std::sin(TwoPi * ky * y / Ly + FPXY) * std::cos(TwoPi * kz * z / Lz + FPXZ)
                     - FAZ * TwoPi * (kx / Lx) * std::cos(TwoPi * kx * x / Lx + FPZX) * std::sin(TwoPi * ky * y / Ly + FPZY) * std::sin(TwoPi * kz * z / Lz + FPZZ));
        else
          f += xT * (FAY * TwoPi * (kx / Lx) * std::cos(TwoPi * kx * x / Lx + FPYX) * std::sin(TwoPi * ky * y / Ly + FPYY) * std::sin(TwoPi * kz * z / Lz + FPYZ)
                     - FAX * TwoPi * (ky / Ly) * std::sin(TwoPi * kx * x / Lx + FPXX) * std::cos(TwoPi * ky * y / Ly + FPXY) * std::sin(TwoPi * kz * z / Lz + FPXZ));
      } else {
        const double FPX = fd(2, kx, ky, kz), FPY = fd(3, kx, ky, kz), FPZ = fd(4, kx, ky, kz);
        const double ax = TwoPi * kx * x / Lx + FPX, ay = TwoPi * ky * y / Ly + FPY, az = TwoPi * kz * z / Lz + FPZ;
        if (c == 0) f += xT * FAX * std::cos(ax) * std::sin(ay) * std::sin(az);
        else if (c == 1) f += xT * FAY * std::sin(ax) * std::cos(ay) * std::sin(az);
        else f += xT * FAZ * std::sin(ax) * std::sin(ay) * std::cos(az);
      }
    }
    return f;
  }
  // NavierStokesBase::getForce for a velocity component (NS_getForce.cpp:117-141 gravity; Tutorials/HIT: + rho * turbulent force)
  double ext_force(int c, double rho, int i, int j, int k, double t) const {
    double f = (c == 2 && std::fabs(p.gravity) > 1.0e-4) ? p.gravity * rho : 0.0;
    if (turb_on) f += rho * turb(c, i, j, k, t);
    return f;
  }

  int advance(double dt, double* dt_test) {
    // advance_setup NSB.cpp:613-741
    std::swap(S_old, S_new); std::swap(P_old, P_new); std::swap(Gp_old, Gp_new);
    fillpatch(rho_p, S_old, Density, 1);
    // ---- predict_velocity NSB.cpp:4376-4512
    Arr Umf(n, 3, 3); fillpatch(Umf, S_old, Xvel, 3);
    for (double& v : Umf.d) v = (std::fabs(v) > 1.0e-20) ? v : 0.0;  // floor :4530-4534
    double cflmax = 0.0;
    for (int d = 0; d < 3; ++d) cflmax = std::max(cflmax, dt * Umf.norminf(d) / dx[d]);
    const double tempdt = (cflmax == 0.0) ? p.change_max : std::min(p.change_max, p.cfl / cflmax);
    Arr visc(n, 3, 1);
    if (p.be_cn_theta != 1.0) visc_terms(S_old, visc); else visc.setval(0.0);
    Arr Smf(n, 2, 3); fillpatch(Smf, S_old, Density, 2);
    fill_gradp(Gp_old);
    force.define(n, 3, 1);
    for (int c = 0; c < 3; ++c) {
      FOR_G1(force, i, j, k) force(i, j, k, c) = (ext_force(c, Smf(i, j, k, 0), i, j, k, time) + visc(i, j, k, c) - Gp_old(i, j, k, c)) / Smf(i, j, k, 0);  // :4466-4470
    }
    const std::vector<BCRec> vbc = adv_bc(Xvel, 3), sbc = adv_bc(Density, 2);
    const AdvOpt aopt{p.use_forces_in_trans != 0, p.use_ppm != 0, has_walls() ? vbc.data() : nullptr, true};
    const AdvOpt sopt{p.use_forces_in_trans != 0, p.use_ppm != 0, has_walls() ? sbc.data() : nullptr, false};
    extrap_vel_to_faces(Umf, &force, aopt, dx, dt, umac);  // :4487
    *dt_test = dt * tempdt;
    // ---- mac_project NS.cpp:589-597, MacProj.cpp:225-353
    Arr mac_phi(n, 1, 1);
    orc_mg m1 = mg(p.mac_tol, p.mac_abs_tol);
    const LinBC mbc = mac_bc();
    int rc = mac_project(n, dx, umac, rho_p, nullptr, mac_phi, 2.0 / dt, &m1, has_walls() ? &mbc : nullptr);
    it[0] = m1.iters;
    if (rc) return rc;
    // ---- velocity_advection NSB.cpp:3358-3470 (fresh un-floored FillPatch copy, same forcing); with do_mom_diff it runs AFTER
    //      rho^{n+1} exists (NS.cpp:606-623), advects the momentum rho^n u^n conservatively and its forcing is not divided by rho
    auto velocity_advection = [&]() {
      Arr Umf2(n, 3, 3); fillpatch(Umf2, S_old, Xvel, 3);
      if (p.do_mom_diff) {
        Arr rho3(n, 1, 3); fillpatch(rho3, S_old, Density, 1);
        for (int c = 0; c < 3; ++c) {
#pragma omp parallel for
          for (int k = -3; k < n[2] + 3; ++k) for (int j = -3; j < n[1] + 3; ++j) for (int i = -3; i < n[0] + 3; ++i) Umf2(i, j, k, c) *= rho3(i, j, k);
          FOR_G1(force, i, j, k) force(i, j, k, c) = ext_force(c, rho3(i, j, k), i, j, k, time) + visc(i, j, k, c) - Gp_old(i, j, k, c);   // :3459-3466
        }
        const int ic_mom[3] = {1, 1, 1};   // NS_setup.cpp:297-299
        compute_aofs(Umf2, 3, &force, nullptr, umac, ic_mom, aopt, dx, dt, aofs, Xvel, nullptr, nullptr);
      } else {
        const int ic_vel[3] = {0, 0, 0};
        compute_aofs(Umf2, 3, &force, nullptr, umac, ic_vel, aopt, dx, dt, aofs, Xvel, nullptr, nullptr);
      }
    };
    if (!p.do_mom_diff) velocity_advection();
    // ---- scalar_advection NS.cpp:698-812
    for (double& v : Smf.d) v = (std::fabs(v) > 1.0e-20) ? v : 0.0;
    Arr sforce(n, 2, 1);
    if (diffusive_tracer() && p.be_cn_theta != 1.0) {
      // NavierStokes::getViscTerms (NS.cpp:2012-2048) -> Diffusion::getViscTerms (Diffusion.cpp:1540-1652): a = 0, b = -1 applied to
      // S (rho_flag 0) or S/rho (rho_flag 2), then FillBoundary; tf = tf/rho + visc or tf + visc with zero body force (NS.cpp:774-804)
      Arr s1(n, 1, 1); fillpatch(s1, S_old, Tracer, 1);
      if (rho_flag() == 2) { FOR_G1(s1, i, j, k) s1(i, j, k) /= rho_p(i, j, k); }
      Arr lbc = s1;
      CellMG op(n, dx, 1, false, 0);
      op.a = 0.0; op.b = -1.0;
      if (has_walls()) op.set_bc(diff_bc(Tracer, 1), &lbc);
      const Arr* e[3] = {&seta[0], &seta[1], &seta[2]};
      op.set_coeffs(nullptr, e);
      Arr sv(n, 1, 1);
      op.apply(sv, s1);
      sv.fill_periodic();
      if (has_walls()) first_order_extrap(sv, 0, 1);
      FOR_G1(sforce, i, j, k) sforce(i, j, k, 1) = sv(i, j, k);
    }
    const int ic_scal[2] = {1, p.conservative_tracer ? 1 : 0};
    compute_aofs(Smf, 2, &sforce, nullptr, umac, ic_scal, sopt, dx, dt, aofs, Density, nullptr, nullptr);
    // ---- scalar updates NSB.cpp:2761-2765, 2887-2896
    FOR_CELLS(S_new, i, j, k) S_new(i, j, k, Density) = S_old(i, j, k, Density) - dt * aofs(i, j, k, Density);
    fillpatch(rho_c, S_new, Density, 1);
    if (p.do_mom_diff) velocity_advection();
    FOR_CELLS(S_new, i, j, k) S_new(i, j, k, Tracer) = S_old(i, j, k, Tracer) - dt * aofs(i, j, k, Tracer);
    if (p.do_scalminmax) {   // NSB.cpp:2907-2935 -> Conservative / ConvectiveScalMinMax (:4256-4370) on an un-floored copy of the old scalars
      Arr so(n, 2, 1); fillpatch(so, S_old, Density, 2);
      const bool cons = p.conservative_tracer != 0;
      FOR_CELLS(S_new, i, j, k) {
        double smn = std::numeric_limits<double>::max(), smx = std::numeric_limits<double>::min();   // sic: smallest positive value
        for (int kk = -1; kk <= 1; ++kk) for (int jj = -1; jj <= 1; ++jj) for (int ii = -1; ii <= 1; ++ii) {
          const double v = cons ? so(i + ii, j + jj, k + kk, 1) / so(i + ii, j + jj, k + kk, 0) : so(i + ii, j + jj, k + kk, 1);
          smn = std::min(smn, v); smx = std::max(smx, v);
        }
        const double r = S_new(i, j, k, Density);
        S_new(i, j, k, Tracer) = cons ? std::min(std::max(S_new(i, j, k, Tracer) / r, smn), smx) * r
                                      : std::min(std::max(S_new(i, j, k, Tracer), smn), smx);
      }
    }
    if (diffusive_tracer()) { rc = tracer_diffusion(dt); if (rc) return rc; }   // scalar_update -> scalar_diffusion_update NS.cpp:836-841
    // ---- velocity_update NSB.cpp:3487-3655
    FOR_G1(rho_half, i, j, k) rho_half(i, j, k) = 0.5 * (rho_p(i, j, k) + rho_c(i, j, k));
    const bool zero_force = initial_iter && diffusive();
    for (int c = 0; c < 3; ++c) {
      FOR_CELLS(S_new, i, j, k) {
        const double r = rho_half(i, j, k);
        const double frc = zero_force ? 0.0 : ext_force(c, r, i, j, k, time + 0.5 * dt);   // half_time (NSB.cpp:3581-3583)
        if (p.do_mom_diff)   // NSB.cpp:3609-3616
          S_new(i, j, k, c) = (S_old(i, j, k, c) * S_old(i, j, k, Density) - dt * aofs(i, j, k, c) + dt * frc - dt * Gp_old(i, j, k, c)) / S_new(i, j, k, Density);
        else
          S_new(i, j, k, c) = S_old(i, j, k, c) - dt * aofs(i, j, k, c) + dt * frc / r - dt * Gp_old(i, j, k, c) / r;
      }
    }
    if (!initial_iter) { rc = velocity_diffusion(dt); if (rc) return rc; }
    else initial_velocity_diffusion(dt);
    if (!initial_step) { rc = level_project(dt); if (rc) return rc; }
    return 0;
  }

  // NavierStokes::scalar_diffusion_update (NS.cpp:858-1000) -> Diffusion::diffuse_scalar (Diffusion.cpp:207-600) for the tracer:
  // Crank-Nicolson with rho_flag 0 (S diffuses, alpha = 1) or 2 (S/rho diffuses, alpha = rho_new, result times rho_new)
  int tracer_diffusion(double dt) {
    const double th = p.be_cn_theta;
    const int rf = rho_flag();
    const Arr* e[3] = {&seta[0], &seta[1], &seta[2]};
    Arr rhs(n, 1, 0);
    if (th != 1.0) {   // :364-430: Rhs = -(b) div beta grad (old solution), a = 0, b = -(1-theta) dt
      Arr so(n, 1, 1); fillpatch(so, S_old, Tracer, 1);
      if (rf == 2) { FOR_G1(so, i, j, k) so(i, j, k) /= rho_p(i, j, k); }
      Arr lbc = so;
      CellMG ex(n, dx, 1, false, 0);
      ex.a = 0.0; ex.b = -(1.0 - th) * dt;
      if (has_walls()) ex.set_bc(diff_bc(Tracer, 1), &lbc);
      ex.set_coeffs(nullptr, e);
      ex.apply(rhs, so);
    }
    FOR_CELLS(rhs, i, j, k) rhs(i, j, k) += S_new(i, j, k, Tracer);   // :465-490 (rho_flag 0 and 2: no scaling)
    Arr soln(n, 1, 1), alpha(n, 1, 0);
    fillpatch(soln, S_new, Tracer, 1);   // :520-540 initial guess = FillPatch'd new state (/ rho_new) incl. its ghost cells = the level BC
    if (rf == 2) { FOR_G1(soln, i, j, k) soln(i, j, k) /= rho_c(i, j, k); }
    FOR_CELLS(alpha, i, j, k) alpha(i, j, k) = (rf == 2) ? S_new(i, j, k, Density) : 1.0;   // :1355-1395 computeAlpha
    Arr lbc1 = soln;
    const double tol_abs = p.visc_tol * rhs.norminf(0);   // get_scaled_abs_tol :193-204
    CellMG im(n, dx, 1, false, 100);
    im.mg = mg(p.visc_tol, tol_abs);
    im.a = 1.0; im.b = th * dt;
    if (has_walls()) im.set_bc(diff_bc(Tracer, 1), &lbc1);
    im.set_coeffs(&alpha, e);
    const int rc = im.solve(soln, rhs);
    FOR_CELLS(soln, i, j, k) S_new(i, j, k, Tracer) = (rf == 2) ? soln(i, j, k) * S_new(i, j, k, Density) : soln(i, j, k);   // :575-590
    return rc;
  }

  // Diffusion::diffuse_tensor_velocity Diffusion.cpp:650-957
  int velocity_diffusion(double dt) {
    if (!diffusive()) return 0;
    const double th = p.be_cn_theta;
    const Arr* e[3] = {&eta[0], &eta[1], &eta[2]};
    Arr rhs(n, 3, 0);
    if (th != 1.0) {
      Arr u(n, 3, 1); fillpatch(u, S_old, Xvel, 3);   // Diffusion.cpp:742-743: FillPatch at prev_time = the level BC
      Arr lbc = u;
      CellMG ex(n, dx, 3, true, 0);
      ex.a = 0.0; ex.b = -(1.0 - th) * dt;
      if (has_walls()) ex.set_bc(diff_bc(Xvel, 3), &lbc);
      ex.set_coeffs(nullptr, e);
      ex.apply(rhs, u);
    }
    for (int c = 0; c < 3; ++c) {
      FOR_CELLS(rhs, i, j, k) {   // :814-831: rho_flag 1 -> rho_half, rho_flag 3 (do_mom_diff, NS.cpp:1016) -> the OLD density
        S_new(i, j, k, c) *= p.do_mom_diff ? S_old(i, j, k, Density) : rho_half(i, j, k);
        rhs(i, j, k, c) += S_new(i, j, k, c);
      }
    }
    const double tol_abs = p.visc_tol * (rhs.norminf(0) + rhs.norminf(1) + rhs.norminf(2)) / 3.0;  // get_scaled_abs_tol :193-204
    Arr soln(n, 3, 1); fillpatch(soln, S_new, Xvel, 3);   // :885-886 FillPatch at cur_time (U_new holds rho U* by now) = the level BC
    Arr lbc2 = soln;
    CellMG im(n, dx, 3, true, 100);
    im.mg = mg(p.visc_tol, tol_abs);
    im.a = 1.0; im.b = th * dt;
    if (has_walls()) im.set_bc(diff_bc(Xvel, 3), &lbc2);
    Arr rho_n(n, 1, 0); rho_n.copy_from(S_new, Density, 0, 1);
    im.set_coeffs(p.do_mom_diff ? &rho_n : &rho_half, e);   // :893-897 alpha = rho_half or (rho_flag 3) the NEW density
    const int rc = im.solve(soln, rhs);
    it[1] = im.mg.iters;
    S_new.copy_from(soln, 0, Xvel, 3);
    return rc;
  }
  // NSB.cpp:3658-3749
  void initial_velocity_diffusion(double dt) {
    if (!diffusive()) return;
    Arr visc(n, 3, 1);
    if (p.be_cn_theta != 1.0) visc_terms(S_old, visc); else visc.setval(0.0);
    for (int c = 0; c < 3; ++c) {
      FOR_CELLS(S_new, i, j, k) {
        double f = ext_force(c, S_old(i, j, k, Density), i, j, k, time) + visc(i, j, k, c) - Gp_old(i, j, k, c);   // prev_time (:3696)
        if (!p.do_mom_diff) f /= rho_half(i, j, k);
        f -= aofs(i, j, k, c);
        if (p.do_mom_diff) S_new(i, j, k, c) = (f * dt + S_old(i, j, k, c) * S_old(i, j, k, Density)) / S_new(i, j, k, Density);   // :3743
        else S_new(i, j, k, c) = S_old(i, j, k, c) + f * dt;
      }
    }
  }
  // Velocity ghost cells beyond INFLOW faces for the nodal projections: setPhysBoundaryValues (Projection.cpp:211-217, 729,
  // 1096-1097) + set_boundary_velocity with inflowCorner = true (:2570-2663): normal component = scale * inflow value on the face
  // cells and their periodic extensions, zero in the corners outside walls.  scale: 1 (initialVelocityProject), 1/dt
  // (level_project: U_new.mult(dt_inv, 0, 3, 1) :274), 0 (initialSyncProject of a steady inflow value, ConvertUnew :1192-1232)
  void inflow_ghost_velocity(Arr& vel, double scale) const {
    for (int d = 0; d < 3; ++d) {
      if (per[d]) continue;
      const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
      for (int side = 0; side < 2; ++side) {
        if ((side == 0 ? phys_lo[d] : phys_hi[d]) != PHYS_INFLOW) continue;
        const int g = side == 0 ? -1 : n[d];
        const double v = scale * bcv[side == 0 ? d : 3 + d][d];
        for (int b2 = -1; b2 <= n[d2]; ++b2)
          for (int b1 = -1; b1 <= n[d1]; ++b1) {
            const bool in1 = (b1 >= 0 && b1 < n[d1]) || per[d1], in2 = (b2 >= 0 && b2 < n[d2]) || per[d2];
            int q[3]; q[d] = g; q[d1] = b1; q[d2] = b2;
            vel(q[0], q[1], q[2], d) = (in1 && in2) ? v : 0.0;
          }
      }
    }
  }
  // Projection::level_project Projection.cpp:166-450
  // Projection::set_outflow_bcs for LEVEL_PROJ / INITIAL_PRESS under gravity (Projection.cpp:1721-1931) -> computeRhoG (:1933-2379),
  // kept in the reference's loop structure: interior node columns first, then the two transverse edges by the density's BCRec.
  // rho_valid: component rc of an array whose valid cells hold the density the projection uses.
  void set_outflow_bcs(Arr& phi, const Arr& rho_valid, int rc) const {
    if (!(std::fabs(p.gravity) > 0.0) || !has_walls()) return;
    Arr rho(n, 1, 1);
    rho.copy_from(rho_valid, rc, 0, 1);
    fill_state_bc(rho, Density, 1);   // the one-cell transverse growth of the state strip (:1784-1786, 1891-1893)
    const BCRec db = state_bc(Density);
    const double dh = dx[2];
    for (int od = 0; od < 2; ++od) {       // outDir x or y (top: nothing to do; bottom: the reference aborts, :1949-1958)
      if (per[od]) continue;
      const int td = 1 - od;
      for (int side = 0; side < 2; ++side) {
        if ((side == 0 ? phys_lo[od] : phys_hi[od]) != PHYS_OUTFLOW) continue;
        const int pl = side == 0 ? 0 : n[od];                       // node plane of the face
        const int c1 = side == 0 ? 0 : n[od] - 1, c2 = side == 0 ? 1 : n[od] - 2;   // rho(i), rho(i+1)  /  rho(i-1), rho(i-2)
        auto RHO = [&](int c, int jj, int kk) { return od == 0 ? rho(c, jj, kk) : rho(jj, c, kk); };
        auto PHI = [&](int jj, int kk) -> double& { return od == 0 ? phi(pl, jj, kk) : phi(jj, pl, kk); };
        const int lo_code = per[td] ? BC_INT_DIR : db.lo[td], hi_code = per[td] ? BC_INT_DIR : db.hi[td];
        auto special = [](int code) { return code == BC_EXT_DIR || code == BC_HOEXTRAP || code == BC_FOEXTRAP; };
        const int jlo = special(lo_code) ? 1 : 0, jhi = special(hi_code) ? n[td] - 1 : n[td];
        auto add_rhog = [&](double rho1, double rho2, double& rhog, double& phi_i) {
          const double rhoExt = 0.5 * (3.0 * rho1 - rho2);
          rhog -= p.gravity * rhoExt * dh;
          phi_i += rhog;
        };
        for (int jj = 0; jj <= n[td]; ++jj) for (int kk = 0; kk <= n[2]; ++kk) PHI(jj, kk) = 0.0;   // phi_fine_strip.setVal(0)
        for (int jj = jlo; jj <= jhi; ++jj) {
          double rhog = 0.0;
          for (int kk = n[2] - 1; kk >= 0; --kk)
            add_rhog(0.5 * (RHO(c1, jj, kk) + RHO(c1, jj - 1, kk)), 0.5 * (RHO(c2, jj, kk) + RHO(c2, jj - 1, kk)), rhog, PHI(jj, kk));
        }
        if (special(lo_code)) {
          const int jj = 0; double rhog = 0.0;
          for (int kk = n[2] - 1; kk >= 0; --kk) {
            if (lo_code == BC_EXT_DIR) add_rhog(RHO(c1, jj - 1, kk), RHO(c2, jj - 1, kk), rhog, PHI(jj, kk));
            else if (lo_code == BC_HOEXTRAP) add_rhog(0.5 * (3.0 * RHO(c1, jj, kk) - RHO(c1, jj + 1, kk)), 0.5 * (3.0 * RHO(c2, jj, kk) - RHO(c2, jj + 1, kk)), rhog, PHI(jj, kk));
            else add_rhog(RHO(c1, jj, kk), RHO(c2, jj, kk), rhog, PHI(jj, kk));
          }
        }
        if (special(hi_code)) {
          const int jj = n[td]; double rhog = 0.0;
          for (int kk = n[2] - 1; kk >= 0; --kk) {
            if (hi_code == BC_EXT_DIR) add_rhog(RHO(c1, jj, kk), RHO(c2, jj, kk), rhog, PHI(jj, kk));
            else if (hi_code == BC_HOEXTRAP) add_rhog(0.5 * (3.0 * RHO(c1, jj - 1, kk) - RHO(c1, jj - 2, kk)), 0.5 * (3.0 * RHO(c2, jj - 1, kk) - RHO(c2, jj - 2, kk)), rhog, PHI(jj, kk));
            else add_rhog(RHO(c1, jj - 1, kk), RHO(c2, jj - 1, kk), rhog, PHI(jj, kk));
          }
        }
      }
    }
  }
  int level_project(double dt) {
    P_new.setval(0.0);
    set_outflow_bcs(P_new, rho_half, 0);   // Projection.cpp:304-324
    Arr vel(n, 3, 1), sig(n, 1, 1);
    for (int c = 0; c < 3; ++c) { FOR_CELLS(vel, i, j, k) vel(i, j, k, c) = S_new(i, j, k, c) * (1.0 / dt) + Gp_old(i, j, k, c) / rho_half(i, j, k); }
    FOR_CELLS(sig, i, j, k) sig(i, j, k) = 1.0 / rho_half(i, j, k);
    inflow_ghost_velocity(vel, 1.0 / dt);
    orc_mg m = mg(p.proj_tol, p.proj_abs_tol);
    const int rc = nodal_project(n, dx, vel, sig, P_new, &Gp_new, false, &m, nodal_bc());
    it[2] = m.iters;
    if (rc) return rc;
    fill_gradp(Gp_new);   // Projection.cpp:2565 FillPatch of Gradp
    for (int c = 0; c < 3; ++c) { FOR_CELLS(vel, i, j, k) S_new(i, j, k, c) = vel(i, j, k, c) * dt; }
    return 0;
  }
  // NSB.cpp:1353-1500
  int est_time_step(double* out) {
    if (p.fixed_dt > 0.0) { *out = p.fixed_dt; return 0; }
    double est = 1.0e20;
    for (int d = 0; d < 3; ++d) {
      const double um = S_new.norminf(d);
      double fm = 0.0;
#pragma omp parallel for reduction(max : fm)
      for (int k = 0; k < n[2]; ++k)
        for (int j = 0; j < n[1]; ++j)
          for (int i = 0; i < n[0]; ++i) {
            const double r = S_new(i, j, k, Density);
            fm = std::max(fm, std::fabs((ext_force(d, r, i, j, k, time) - Gp_new(i, j, k, d)) / r));   // cur_time (:1410)
          }
      if (um > 1.0e-8) est = std::min(est, dx[d] / um);
      if (fm > 1.0e-8) est = std::min(est, std::sqrt(2.0 * dx[d] / fm));
    }
    if (est >= 1.0e20) return -1;
    *out = est * p.cfl;
    return 0;
  }
};

extern "C" {

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int orc_set_option(int opt, double value) {   // same option ids as include/iamrx.h IAMRX_OPT_*
  switch (opt) {
    case 0: small_vel = value; return 0;
    case 1: opt_slope_order = (int)value; return 0;
    case 2: opt_corner_adv = (int)value; return 0;
    case 3: opt_extdir_both = (int)value; return 0;
    default: return -1;
  }
}

void orc_set_num_threads(int n) {   // torchrun exports OMP_NUM_THREADS=1: the CPU arm sets its thread count explicitly
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void orc_mg_default(orc_mg* m) {
  m->rtol = 1e-12; m->atol = 1e-16; m->max_iter = 200; m->nu1 = 2; m->nu2 = 2; m->bottom_sweeps = 8; m->max_coarsening = 100;
  m->omega = 1.15; m->iters = 0; m->resnorm0 = m->resnorm = m->rhsnorm = 0.0;
  m->bottom_solver = 0; m->bottom_maxiter = 200; m->bottom_rtol = 1.0e-4; m->bottom_iters = 0; m->pad_ = 0;
}

static void load_faces(const int n[3], const double* bx, const double* by, const double* bz, int bn, Arr b[3]) {
  const double* src[3] = {bx, by, bz};
  for (int d = 0; d < 3; ++d) { b[d].define(n, bn, 1); b[d].load(src[d]); b[d].fill_periodic(); }
}

void orc_abec_apply(const int n[3], const double dxinv[3], double a, double b, const double* alpha, const double* bx,
                    const double* by, const double* bz, int ncomp, int bncomp, const double* phi, double* out) {
  Arr be[3]; load_faces(n, bx, by, bz, bncomp, be);
  Arr al; if (alpha) { al.define(n, 1, 0); al.load(alpha); }
  Arr p(n, ncomp, 1), o(n, ncomp, 0); p.load(phi);
  AbecOp op; op.a = a; op.b = b; op.alpha = alpha ? &al : nullptr; op.bncomp = bncomp;
  for (int d = 0; d < 3; ++d) { op.beta[d] = &be[d]; op.dxinv[d] = dxinv[d]; }
  abec_apply(op, p, o, ncomp);
  o.store(out);
}

void orc_abec_gsrb(const int n[3], const double dxinv[3], double a, double b, const double* alpha, const double* bx,
                   const double* by, const double* bz, int ncomp, int bncomp, const double* rhs, double omega, int redblack,
                   double* phi) {
  Arr be[3]; load_faces(n, bx, by, bz, bncomp, be);
  Arr al; if (alpha) { al.define(n, 1, 0); al.load(alpha); }
  Arr p(n, ncomp, 1), r(n, ncomp, 0); p.load(phi); r.load(rhs);
  AbecOp op; op.a = a; op.b = b; op.alpha = alpha ? &al : nullptr; op.bncomp = bncomp;
  for (int d = 0; d < 3; ++d) { op.beta[d] = &be[d]; op.dxinv[d] = dxinv[d]; }
  abec_gsrb(op, p, r, ncomp, omega, redblack);
  p.store(phi);
}

void orc_tensor_cross(const int n[3], const double dxinv[3], double b, const double* ex, const double* ey, const double* ez,
                      const double* vel, double* out) {
  Arr e[3]; load_faces(n, ex, ey, ez, 1, e);
  Arr v(n, 3, 1), o(n, 3, 0); v.load(vel); o.load(out);
  tensor_cross(dxinv, b, e[0], e[1], e[2], v, o);
  o.store(out);
}

int orc_diffusion_solve(const int n[3], const double dx[3], int tensor, int ncomp, double a, double b, const double* alpha,
                        const double* ex, const double* ey, const double* ez, const double* rhs, double* soln, orc_mg* mgp) {
  Arr e[3]; load_faces(n, ex, ey, ez, 1, e);
  Arr al; if (alpha) { al.define(n, 1, 0); al.load(alpha); }
  CellMG mg(n, dx, ncomp, tensor != 0, mgp ? mgp->max_coarsening : 100);
  if (mgp) mg.mg = *mgp;
  mg.a = a; mg.b = b;
  const Arr* ep[3] = {&e[0], &e[1], &e[2]};
  mg.set_coeffs(alpha ? &al : nullptr, ep);
  Arr s(n, ncomp, 1), r(n, ncomp, 0); s.load(soln); r.load(rhs);
  const int rc = mg.solve(s, r);
  if (mgp) *mgp = mg.mg;
  s.store(soln);
  return rc;
}

void orc_diffusion_apply(const int n[3], const double dx[3], int tensor, int ncomp, double a, double b, const double* alpha,
                         const double* ex, const double* ey, const double* ez, const double* soln, double* out) {
  Arr e[3]; load_faces(n, ex, ey, ez, 1, e);
  Arr al; if (alpha) { al.define(n, 1, 0); al.load(alpha); }
  CellMG mg(n, dx, ncomp, tensor != 0, 0);
  mg.a = a; mg.b = b;
  const Arr* ep[3] = {&e[0], &e[1], &e[2]};
  mg.set_coeffs(alpha ? &al : nullptr, ep);
  Arr s(n, ncomp, 1), o(n, ncomp, 0); s.load(soln);
  mg.apply(o, s);
  o.store(out);
}

int orc_mac_project(const int n[3], const double dx[3], double* umac, double* vmac, double* wmac, const double* rho,
                    const double* rhs, double* phi, double rhs_scale, orc_mg* mg) {
  Arr mac[3]; double* m[3] = {umac, vmac, wmac};
  for (int d = 0; d < 3; ++d) { mac[d].define(n, 1, 1); mac[d].load(m[d]); }
  Arr r(n, 1, 1), p(n, 1, 1), rh; r.load(rho); p.load(phi);
  if (rhs) { rh.define(n, 1, 0); rh.load(rhs); }
  const int rc = mac_project(n, dx, mac, r, rhs ? &rh : nullptr, p, rhs_scale, mg);
  for (int d = 0; d < 3; ++d) mac[d].store(m[d]);
  p.store(phi);
  return rc;
}

void orc_nodal_divu(const int n[3], const double dxinv[3], const double* vel, double* rhs) {
  Arr v(n, 3, 1), r(n, 1, 0); v.load(vel);
  nodal_divu(dxinv, v, r);
  r.store(rhs);
}
void orc_nodal_adotx(const int n[3], const double dxinv[3], const double* sigma, const double* phi, double* out) {
  Arr s(n, 1, 1), p(n, 1, 1), o(n, 1, 0); s.load(sigma); s.fill_periodic(); p.load(phi);
  nodal_adotx(dxinv, s, p, o);
  o.store(out);
}
void orc_nodal_gs(const int n[3], const double dxinv[3], const double* sigma, const double* rhs, int color, double* phi) {
  Arr s(n, 1, 1), p(n, 1, 1), r(n, 1, 0); s.load(sigma); s.fill_periodic(); p.load(phi); r.load(rhs);
  nodal_gs(dxinv, s, r, color, p);
  p.store(phi);
}
void orc_nodal_mknewu(const int n[3], const double dxinv[3], const double* sigma, const double* phi, double* vel, double* gp) {
  Arr p(n, 1, 1), g(n, 3, 0); p.load(phi);
  nodal_grad(dxinv, p, g);
  if (gp) g.store(gp);
  if (vel) {
    Arr s(n, 1, 0), v(n, 3, 0); s.load(sigma); v.load(vel);
    for (int c = 0; c < 3; ++c) { FOR_CELLS(v, i, j, k) v(i, j, k, c) -= s(i, j, k) * g(i, j, k, c); }
    v.store(vel);
  }
}
int orc_nodal_project(const int n[3], const double dx[3], double* vel, const double* sigma, double* phi, double* gp,
                      int increment_gp, orc_mg* mg) {
  Arr v(n, 3, 1), s(n, 1, 1), p(n, 1, 1), g(n, 3, 0); v.load(vel); s.load(sigma); p.load(phi);
  if (gp && increment_gp) g.load(gp);
  const int rc = nodal_project(n, dx, v, s, p, gp ? &g : nullptr, increment_gp != 0, mg);
  v.store(vel); p.store(phi);
  if (gp) g.store(gp);
  return rc;
}

void orc_extrap_vel_to_faces(const int n[3], const double dx[3], double dt, const double* vel, const double* force,
                             int forces_in_trans, double* umac, double* vmac, double* wmac) {
  Arr v(n, 3, 3), f; v.load(vel); v.fill_periodic();
  if (force) { f.define(n, 3, 1); f.load(force); f.fill_periodic(); }
  Arr mac[3] = {Arr(n, 1, 1), Arr(n, 1, 1), Arr(n, 1, 1)};
  extrap_vel_to_faces(v, force ? &f : nullptr, AdvOpt{(forces_in_trans & 1) != 0, (forces_in_trans & 2) != 0}, dx, dt, mac);
  mac[0].store(umac); mac[1].store(vmac); mac[2].store(wmac);
}

void orc_compute_aofs(const int n[3], const double dx[3], double dt, int ncomp, const double* S, const double* force,
                      const double* divu, const double* umac, const double* vmac, const double* wmac, const int* iconserv,
                      int forces_in_trans, double* aofs, double* fx, double* fy, double* fz, double* xed, double* yed, double* zed) {
  Arr q(n, ncomp, 3), f, dv; q.load(S); q.fill_periodic();
  if (force) { f.define(n, ncomp, 1); f.load(force); f.fill_periodic(); }
  if (divu) { dv.define(n, 1, 1); dv.load(divu); dv.fill_periodic(); }
  Arr mac[3]; const double* m[3] = {umac, vmac, wmac};
  for (int d = 0; d < 3; ++d) { mac[d].define(n, 1, 1); mac[d].load(m[d]); mac[d].fill_periodic(); }
  Arr a(n, ncomp, 0);
  Arr fl[3], ed[3]; Arr* flp[3] = {nullptr, nullptr, nullptr}; Arr* edp[3] = {nullptr, nullptr, nullptr};
  double* fo[3] = {fx, fy, fz}; double* eo[3] = {xed, yed, zed};
  for (int d = 0; d < 3; ++d) {
    if (fo[d]) { fl[d].define(n, ncomp, 0); flp[d] = &fl[d]; }
    if (eo[d]) { ed[d].define(n, ncomp, 0); edp[d] = &ed[d]; }
  }
  compute_aofs(q, ncomp, force ? &f : nullptr, divu ? &dv : nullptr, mac, iconserv, AdvOpt{(forces_in_trans & 1) != 0, (forces_in_trans & 2) != 0}, dx, dt, a, 0, flp, edp);
  a.store(aofs);
  for (int d = 0; d < 3; ++d) { if (fo[d]) fl[d].store(fo[d]); if (eo[d]) ed[d].store(eo[d]); }
}

/* the full argument list of the ComputeFluxesOnBoxFromState call site (NSB.cpp:4701-4717): separate flux velocities, the
 * sync sign convention (aofs holds the running Sync on entry) and known edge states (xed..zed are inputs then) */
void orc_compute_aofs2(const int n[3], const double dx[3], double dt, int ncomp, const double* S, const double* force,
                       const double* divu, const double* umac, const double* vmac, const double* wmac, const double* uflux,
                       const double* vflux, const double* wflux, const int* iconserv, int flags, int is_sync, int known,
                       double* aofs, double* fx, double* fy, double* fz, double* xed, double* yed, double* zed) {
  Arr q(n, ncomp, 3), f, dv; q.load(S); q.fill_periodic();
  if (force) { f.define(n, ncomp, 1); f.load(force); f.fill_periodic(); }
  if (divu) { dv.define(n, 1, 1); dv.load(divu); dv.fill_periodic(); }
  Arr mac[3], ufl[3]; const double* m[3] = {umac, vmac, wmac}; const double* uf[3] = {uflux, vflux, wflux};
  for (int d = 0; d < 3; ++d) {
    mac[d].define(n, 1, 1); mac[d].load(m[d]); mac[d].fill_periodic();
    if (uflux) { ufl[d].define(n, 1, 1); ufl[d].load(uf[d]); ufl[d].fill_periodic(); }
  }
  Arr a(n, ncomp, 0); a.load(aofs);
  Arr fl[3], ed[3]; Arr* flp[3] = {nullptr, nullptr, nullptr}; Arr* edp[3] = {nullptr, nullptr, nullptr};
  double* fo[3] = {fx, fy, fz}; double* eo[3] = {xed, yed, zed};
  for (int d = 0; d < 3; ++d) {
    if (fo[d]) { fl[d].define(n, ncomp, 0); flp[d] = &fl[d]; }
    if (eo[d]) { ed[d].define(n, ncomp, 0); edp[d] = &ed[d]; if (known) ed[d].load(eo[d]); }
  }
  compute_aofs(q, ncomp, force ? &f : nullptr, divu ? &dv : nullptr, mac, iconserv, AdvOpt{(flags & 1) != 0, (flags & 2) != 0}, dx, dt, a, 0,
               flp, edp, uflux ? ufl : nullptr, is_sync != 0, known != 0);
  a.store(aofs);
  for (int d = 0; d < 3; ++d) { if (fo[d]) fl[d].store(fo[d]); if (eo[d] && !known) ed[d].store(eo[d]); }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Non-periodic domains.  Arrays are exchanged "padded" (with their ghost layers, see Arr::load_padded); bclo / bchi are
 * BCRec codes [ncomp][3], lobc / hibc LinOpBCType codes.
 * ---------------------------------------------------------------------------------------------------------------- */
static std::vector<BCRec> bcrecs(int ncomp, const int* bclo, const int* bchi) {
  std::vector<BCRec> b(ncomp);
  for (int c = 0; c < ncomp; ++c) for (int d = 0; d < 3; ++d) { b[c].lo[d] = bclo[3 * c + d]; b[c].hi[d] = bchi[3 * c + d]; }
  return b;
}
void orc_fill_physbc(const int n[3], const int per[3], int ng, int ncomp, const int* bclo, const int* bchi, const double* bcv, double* a) {
  PerScope ps(per);
  Arr A(n, ncomp, ng); A.load_padded(a);
  A.fill_periodic();
  const std::vector<BCRec> b = bcrecs(ncomp, bclo, bchi);
  fill_physbc(A, 0, ncomp, b.data(), bcv);
  A.store_padded(a);
}
void orc_extrap_vel_to_faces_bc(const int n[3], const int per[3], const double dx[3], double dt, const double* vel /*ng 3*/,
                                const double* force /*ng 1 or NULL*/, int flags, const int* bclo, const int* bchi,
                                double* umac, double* vmac, double* wmac /*ng 1*/) {
  PerScope ps(per);
  Arr v(n, 3, 3), f; v.load_padded(vel);
  if (force) { f.define(n, 3, 1); f.load_padded(force); }
  const std::vector<BCRec> b = bcrecs(3, bclo, bchi);
  Arr mac[3] = {Arr(n, 1, 1), Arr(n, 1, 1), Arr(n, 1, 1)};
  extrap_vel_to_faces(v, force ? &f : nullptr, AdvOpt{(flags & 1) != 0, (flags & 2) != 0, b.data(), true}, dx, dt, mac);
  mac[0].store_padded(umac); mac[1].store_padded(vmac); mac[2].store_padded(wmac);
}
void orc_compute_aofs_bc(const int n[3], const int per[3], const double dx[3], double dt, int ncomp, const double* S /*ng 3*/,
                         const double* force /*ng 1*/, const double* divu /*ng 1*/, const double* umac, const double* vmac,
                         const double* wmac /*ng 1*/, const int* iconserv, int flags /*1 fit, 2 ppm, 4 is_velocity*/, const int* bclo,
                         const int* bchi, double* aofs /*dense*/, double* fx, double* fy, double* fz, double* xed, double* yed, double* zed /*ng 1 or NULL*/) {
  PerScope ps(per);
  Arr q(n, ncomp, 3), f, dv; q.load_padded(S);
  if (force) { f.define(n, ncomp, 1); f.load_padded(force); }
  if (divu) { dv.define(n, 1, 1); dv.load_padded(divu); }
  Arr mac[3]; const double* m[3] = {umac, vmac, wmac};
  for (int d = 0; d < 3; ++d) { mac[d].define(n, 1, 1); mac[d].load_padded(m[d]); }
  const std::vector<BCRec> b = bcrecs(ncomp, bclo, bchi);
  Arr a(n, ncomp, 0);
  Arr fl[3], ed[3]; Arr* flp[3] = {nullptr, nullptr, nullptr}; Arr* edp[3] = {nullptr, nullptr, nullptr};
  double* fo[3] = {fx, fy, fz}; double* eo[3] = {xed, yed, zed};
  for (int d = 0; d < 3; ++d) {
    if (fo[d]) { fl[d].define(n, ncomp, 1); flp[d] = &fl[d]; }
    if (eo[d]) { ed[d].define(n, ncomp, 1); edp[d] = &ed[d]; }
  }
  compute_aofs(q, ncomp, force ? &f : nullptr, divu ? &dv : nullptr, mac, iconserv, AdvOpt{(flags & 1) != 0, (flags & 2) != 0, b.data(), (flags & 4) != 0},
               dx, dt, a, 0, flp, edp);
  a.store(aofs);
  for (int d = 0; d < 3; ++d) { if (fo[d]) fl[d].store_padded(fo[d]); if (eo[d]) ed[d].store_padded(eo[d]); }
}
static LinBC linbc(int ncomp, const int* lobc, const int* hibc, int maxorder) {
  LinBC L; L.maxorder = maxorder;
  for (int c = 0; c < ncomp && c < 3; ++c) for (int d = 0; d < 3; ++d) { L.lo[c][d] = lobc[3 * c + d]; L.hi[c][d] = hibc[3 * c + d]; }
  return L;
}
int orc_mac_project_bc(const int n[3], const int per[3], const double dx[3], double* umac, double* vmac, double* wmac /*ng 1*/,
                       const double* rho /*ng 1, BC-filled*/, const double* rhs /*dense or NULL*/, double* phi /*ng 1; ghost = level BC*/,
                       double rhs_scale, const int lobc[3], const int hibc[3], int maxorder, orc_mg* mg) {
  PerScope ps(per);
  Arr mac[3]; double* m[3] = {umac, vmac, wmac};
  for (int d = 0; d < 3; ++d) { mac[d].define(n, 1, 1); mac[d].load_padded(m[d]); }
  Arr r(n, 1, 1), p(n, 1, 1), rh; r.load_padded(rho); p.load_padded(phi);
  if (rhs) { rh.define(n, 1, 0); rh.load(rhs); }
  const LinBC L = linbc(1, lobc, hibc, maxorder);
  const int rc = mac_project(n, dx, mac, r, rhs ? &rh : nullptr, p, rhs_scale, mg, &L);
  for (int d = 0; d < 3; ++d) mac[d].store_padded(m[d]);
  p.store_padded(phi);
  return rc;
}
int orc_nodal_project_bc(const int n[3], const int per[3], const double dx[3], double* vel /*ng 1*/, const double* sigma /*dense*/,
                         double* phi /*nodal, ng 2*/, double* gp /*dense, out*/, const int lobc[3], const int hibc[3], orc_mg* mg) {
  PerScope ps(per);
  Arr v(n, 3, 1), s(n, 1, 1), p(n, 1, 2), g(n, 3, 0); v.load_padded(vel); s.load(sigma); p.load_padded(phi);
  NodalBC B; for (int d = 0; d < 3; ++d) { B.lo[d] = lobc[d]; B.hi[d] = hibc[d]; }
  const int rc = nodal_project(n, dx, v, s, p, gp ? &g : nullptr, false, mg, B);
  v.store_padded(vel); p.store_padded(phi);
  if (gp) g.store(gp);
  return rc;
}
/* (a alpha - b div eta grad [+ tensor cross terms]) with domain BCs; soln's ghost cells hold the level BC (Dirichlet face
 * values).  solve != 0: multigrid solve of ... = rhs into soln; else out = L(soln). */
int orc_diffusion_bc(const int n[3], const int per[3], const double dx[3], int solve, int tensor, int ncomp, double a, double b,
                     const double* alpha /*dense*/, const double* ex, const double* ey, const double* ez /*ng 1*/, const double* rhs /*dense*/,
                     double* soln /*ng 1*/, double* out /*dense*/, const int* lobc, const int* hibc, int maxorder, orc_mg* mgp) {
  PerScope ps(per);
  Arr e[3]; const double* es[3] = {ex, ey, ez};
  for (int d = 0; d < 3; ++d) { e[d].define(n, 1, 1); e[d].load_padded(es[d]); }
  Arr al; if (alpha) { al.define(n, 1, 0); al.load(alpha); }
  CellMG mg(n, dx, ncomp, tensor != 0, solve ? (mgp ? mgp->max_coarsening : 100) : 0);
  if (mgp) mg.mg = *mgp;
  mg.a = a; mg.b = b;
  Arr s(n, ncomp, 1); s.load_padded(soln);
  Arr lbc = s;
  const LinBC L = linbc(ncomp, lobc, hibc, maxorder);
  mg.set_bc(L, &lbc);
  const Arr* ep[3] = {&e[0], &e[1], &e[2]};
  mg.set_coeffs(alpha ? &al : nullptr, ep);
  int rc = 0;
  if (solve) {
    Arr r(n, ncomp, 0); r.load(rhs);
    rc = mg.solve(s, r);
    if (mgp) *mgp = mg.mg;
    s.store_padded(soln);
  } else {
    Arr o(n, ncomp, 0);
    mg.apply(o, s);
    o.store(out);
  }
  return rc;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Two-level transfer operators and the coarse-fine flux register, restated on a fully periodic coarse
 * domain nc refined by 2 (dense n-periodic arrays: faces / nodes hold the LOW face / node of each cell).
 * ---------------------------------------------------------------------------------------------------------------- */
/* amrex::average_down (ixtype 0), average_down_faces (1..3), average_down_nodal = injection (4): NSB.cpp:4125-4191 */
void orc_average_down(const int nc[3], int ncomp, int ixtype, const double* fine, double* crse) {
  const int nf[3] = {2 * nc[0], 2 * nc[1], 2 * nc[2]};
  auto F = [&](int i, int j, int k, int c) { return fine[i + (long)nf[0] * (j + (long)nf[1] * (k + (long)nf[2] * c))]; };
  for (int c = 0; c < ncomp; ++c)
    for (int k = 0; k < nc[2]; ++k) for (int j = 0; j < nc[1]; ++j) for (int i = 0; i < nc[0]; ++i) {
      double v = 0.0;
      if (ixtype == 0) { for (int dk = 0; dk < 2; ++dk) for (int dj = 0; dj < 2; ++dj) for (int di = 0; di < 2; ++di) v += F(2 * i + di, 2 * j + dj, 2 * k + dk, c); v *= 0.125; }
      else if (ixtype == 4) v = F(2 * i, 2 * j, 2 * k, c);
      else {
        const int d = ixtype - 1;
        for (int b = 0; b < 2; ++b) for (int a = 0; a < 2; ++a) {
          int o[3]; o[d] = 0; o[(d + 1) % 3] = a; o[(d + 2) % 3] = b;
          v += F(2 * i + o[0], 2 * j + o[1], 2 * k + o[2], c);
        }
        v *= 0.25;
      }
      crse[i + (long)nc[0] * (j + (long)nc[1] * (k + (long)nc[2] * c))] = v;
    }
}
/* The coarse-fine boundary values of a fine-level solve: what MLLinOp::setCoarseFineBC(crse, ratio) (MacProj.cpp:1164-1167) makes
 * MLCellLinOp compute with InterpBndryData::setBndryValues (order 3).  Restated from memory of AMReX_InterpBndryData_3D_K.H
 * interpbndrydata_{x,y,z}_o3 -- AMReX is not in the reference tree: PARITY UNPINNED.  For every ghost cell of the fine box
 * [flo, fhi] (fine indices) that lies in the one-cell layer beyond one of its six sides, is inside the domain (periodic
 * directions: anywhere) and whose coarse parent is not under the fine level:
 *     v = c0 + y dy + y^2 d2y + z dz + z^2 d2z + y z dyz
 * (y, z = the offset of the fine cell centre from the coarse cell centre in coarse cell widths: -1/4 or +1/4 for ratio 2; dy: the
 * centred first difference over the two tangential coarse neighbours when both are usable, the one-sided difference when one is,
 * 0 when none is; d2y: half the second difference, centred stencils only; dyz: a quarter of the mixed difference when all four
 * diagonal neighbours are usable.  usable = inside the domain (periodic wrap) and not covered by the fine level.)
 * The value belongs to the coarse cell-centre plane, i.e. ratio/2 fine cells beyond the face.  crse: dense (ncomp, nc) array;
 * covered: nc bytes, 1 = coarse cell under the fine level; out: (ncomp, fn + 2) padded fine box, only the face ghost cells written. */
void orc_interp_bndry(const int nc[3], const int per[3], int ncomp, const double* crse, const unsigned char* covered, const int flo[3],
                      const int fhi[3], double* out) {
  const int fn[3] = {fhi[0] - flo[0] + 1, fhi[1] - flo[1] + 1, fhi[2] - flo[2] + 1};
  auto wrap = [&](int q[3]) {
    for (int d = 0; d < 3; ++d) {
      if (q[d] < 0 || q[d] >= nc[d]) { if (!per[d]) return false; q[d] = ((q[d] % nc[d]) + nc[d]) % nc[d]; }
    }
    return true;
  };
  auto usable = [&](int i, int j, int k) { int q[3] = {i, j, k}; if (!wrap(q)) return false; return covered[q[0] + (long)nc[0] * (q[1] + (long)nc[1] * q[2])] == 0; };
  auto fl2 = [](int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); };
  for (int c = 0; c < ncomp; ++c) {
    auto CV = [&](int i, int j, int k) { int q[3] = {i, j, k}; wrap(q); return crse[q[0] + (long)nc[0] * (q[1] + (long)nc[1] * (q[2] + (long)nc[2] * c))]; };
    auto O = [&](int i, int j, int k) -> double& {
      return out[(i - flo[0] + 1) + (long)(fn[0] + 2) * ((j - flo[1] + 1) + (long)(fn[1] + 2) * ((k - flo[2] + 1) + (long)(fn[2] + 2) * c))];
    };
    for (int d = 0; d < 3; ++d) {
      const int t1 = (d + 1) % 3, t2 = (d + 2) % 3;
      for (int side = 0; side < 2; ++side) {
        const int g = side == 0 ? flo[d] - 1 : fhi[d] + 1;
        for (int b = flo[t2]; b <= fhi[t2]; ++b)
          for (int a = flo[t1]; a <= fhi[t1]; ++a) {
            int f[3]; f[d] = g; f[t1] = a; f[t2] = b;
            const int I[3] = {fl2(f[0]), fl2(f[1]), fl2(f[2])};
            if (!usable(I[0], I[1], I[2])) continue;   // outside a physical side, or under another fine box
            const double c0 = CV(I[0], I[1], I[2]);
            double v = c0;
            double off[2]; bool um[2], up[2];
            const int tt[2] = {t1, t2};
            for (int q = 0; q < 2; ++q) {
              const int t = tt[q];
              int m[3] = {I[0], I[1], I[2]}, pq[3] = {I[0], I[1], I[2]};
              m[t] -= 1; pq[t] += 1;
              um[q] = usable(m[0], m[1], m[2]); up[q] = usable(pq[0], pq[1], pq[2]);
              const double x = (f[t] - 2 * I[t]) ? 0.25 : -0.25;
              off[q] = x;
              double d1 = 0.0, d2 = 0.0;
              if (um[q] && up[q]) {
                d1 = 0.5 * (CV(pq[0], pq[1], pq[2]) - CV(m[0], m[1], m[2]));
                d2 = 0.5 * (CV(pq[0], pq[1], pq[2]) - 2.0 * c0 + CV(m[0], m[1], m[2]));
              } else if (up[q]) d1 = CV(pq[0], pq[1], pq[2]) - c0;
              else if (um[q]) d1 = c0 - CV(m[0], m[1], m[2]);
              v += x * d1 + x * x * d2;
            }
            {
              bool all = true; double cr[2][2];
              for (int sb = 0; sb < 2; ++sb) for (int sa = 0; sa < 2; ++sa) {
                int q[3] = {I[0], I[1], I[2]}; q[t1] += sa ? 1 : -1; q[t2] += sb ? 1 : -1;
                if (!usable(q[0], q[1], q[2])) all = false; else cr[sb][sa] = CV(q[0], q[1], q[2]);
              }
              if (all) v += off[0] * off[1] * 0.25 * (cr[1][1] - cr[1][0] - cr[0][1] + cr[0][0]);
            }
            O(f[0], f[1], f[2]) = v;
          }
      }
    }
  }
}
/* kind 0 cell_cons_interp (NS_setup.cpp:211: CellConservativeLinear without linear limiting = MC slopes per direction, then one
 * factor per coarse cell that keeps all eight children inside [min, max] of the 3^3 coarse neighbourhood), 1 node_bilinear_interp
 * (NS_setup.cpp:331), 2..4 face_linear_interp for x / y / z faces (NSB.cpp:1127) */
void orc_interp(int kind, const int nc[3], int ncomp, const double* crse, double* fine) {
  const int nf[3] = {2 * nc[0], 2 * nc[1], 2 * nc[2]};
  auto C = [&](int i, int j, int k, int c) {
    i = ((i % nc[0]) + nc[0]) % nc[0]; j = ((j % nc[1]) + nc[1]) % nc[1]; k = ((k % nc[2]) + nc[2]) % nc[2];
    return crse[i + (long)nc[0] * (j + (long)nc[1] * (k + (long)nc[2] * c))];
  };
  auto mc = [](double um, double u0, double up) {
    const double dc = 0.5 * (up - um), df = 2.0 * (up - u0), db = 2.0 * (u0 - um);
    const double lim = (df * db >= 0.0) ? std::min(std::fabs(df), std::fabs(db)) : 0.0;
    return std::copysign(1.0, dc) * std::min(lim, std::fabs(dc));
  };
  for (int c = 0; c < ncomp; ++c)
    for (int K = 0; K < nc[2]; ++K) for (int J = 0; J < nc[1]; ++J) for (int I = 0; I < nc[0]; ++I) {
      const double u = C(I, J, K, c);
      if (kind == 0) {
        const double s[3] = {mc(C(I - 1, J, K, c), u, C(I + 1, J, K, c)), mc(C(I, J - 1, K, c), u, C(I, J + 1, K, c)), mc(C(I, J, K - 1, c), u, C(I, J, K + 1, c))};
        double cmn = u, cmx = u;
        for (int kk = -1; kk <= 1; ++kk) for (int jj = -1; jj <= 1; ++jj) for (int ii = -1; ii <= 1; ++ii) { const double v = C(I + ii, J + jj, K + kk, c); cmn = std::min(cmn, v); cmx = std::max(cmx, v); }
        // the extreme children sit at (+-1/4, +-1/4, +-1/4): the largest excursion is sum |s_d| / 4 either way
        const double dmax = 0.25 * (std::fabs(s[0]) + std::fabs(s[1]) + std::fabs(s[2]));
        double alpha = 1.0;
        if (dmax > 0.0) { if (u + dmax > cmx) alpha = std::min(alpha, (cmx - u) / dmax); if (u - dmax < cmn) alpha = std::min(alpha, (u - cmn) / dmax); }
        for (int dk = 0; dk < 2; ++dk) for (int dj = 0; dj < 2; ++dj) for (int di = 0; di < 2; ++di)
          fine[(2 * I + di) + (long)nf[0] * ((2 * J + dj) + (long)nf[1] * ((2 * K + dk) + (long)nf[2] * c))] =
              u + alpha * (s[0] * (di ? 0.25 : -0.25) + s[1] * (dj ? 0.25 : -0.25) + s[2] * (dk ? 0.25 : -0.25));
      } else if (kind == 1) {
        for (int dk = 0; dk < 2; ++dk) for (int dj = 0; dj < 2; ++dj) for (int di = 0; di < 2; ++di) {
          double acc = 0.0;
          for (int a = 0; a <= dk; ++a) for (int b = 0; b <= dj; ++b) for (int e = 0; e <= di; ++e) acc += C(I + e, J + b, K + a, c);
          fine[(2 * I + di) + (long)nf[0] * ((2 * J + dj) + (long)nf[1] * ((2 * K + dk) + (long)nf[2] * c))] = acc / (double)((1 + di) * (1 + dj) * (1 + dk));
        }
      } else {
        const int d = kind - 2;
        const double up = C(I + (d == 0), J + (d == 1), K + (d == 2), c);
        for (int dk = 0; dk < 2; ++dk) for (int dj = 0; dj < 2; ++dj) for (int di = 0; di < 2; ++di) {
          const int off = d == 0 ? di : (d == 1 ? dj : dk);   // 0: on the coarse face, 1: half way to the next one
          fine[(2 * I + di) + (long)nf[0] * ((2 * J + dj) + (long)nf[1] * ((2 * K + dk) + (long)nf[2] * c))] = off ? 0.5 * (u + up) : u;
        }
      }
    }
}
/* The advective flux register (NSB.cpp:4848-4889, 5083-5096; NS.cpp:1794-1795) restated cell by cell: a coarse cell that is not
 * covered by fine grids looks at its six faces; where the neighbour across a face IS covered, the coarse update used the coarse
 * flux but should have used the sum of the four fine fluxes: reg = dt (sum fine - coarse) / vol, entering (+, low face) or
 * leaving (-, high face) -- area-weighted fluxes, vol = coarse cell volume.  mask: 1 where a coarse cell is covered by fine. */
void orc_fluxreg(const int nc[3], int ncomp, const unsigned char* mask, const double* cfx, const double* cfy, const double* cfz,
                 const double* ffx, const double* ffy, const double* ffz, double dt, double vol, double* reg) {
  const int nf[3] = {2 * nc[0], 2 * nc[1], 2 * nc[2]};
  const double* cf[3] = {cfx, cfy, cfz};
  const double* ff[3] = {ffx, ffy, ffz};
  auto wrap = [](int a, int n) { return ((a % n) + n) % n; };
  auto M = [&](int i, int j, int k) { return mask[wrap(i, nc[0]) + (long)nc[0] * (wrap(j, nc[1]) + (long)nc[1] * wrap(k, nc[2]))] != 0; };
  for (int c = 0; c < ncomp; ++c)
    for (int k = 0; k < nc[2]; ++k) for (int j = 0; j < nc[1]; ++j) for (int i = 0; i < nc[0]; ++i) {
      double r = 0.0;
      if (!M(i, j, k))
        for (int d = 0; d < 3; ++d)
          for (int side = 0; side < 2; ++side) {   // 0: low face, 1: high face
            int q[3] = {i, j, k}; q[d] += side ? 1 : -1;
            if (!M(q[0], q[1], q[2])) continue;
            int fc[3] = {i, j, k}; fc[d] += side;   // coarse face index (low-face convention), periodic
            const double Fc = cf[d][wrap(fc[0], nc[0]) + (long)nc[0] * (wrap(fc[1], nc[1]) + (long)nc[1] * (wrap(fc[2], nc[2]) + (long)nc[2] * c))];
            double Ff = 0.0;
            for (int b = 0; b < 2; ++b) for (int a = 0; a < 2; ++a) {
              int o[3]; o[d] = 0; o[(d + 1) % 3] = a; o[(d + 2) % 3] = b;
              const int fi = wrap(2 * fc[0] + o[0], nf[0]), fj = wrap(2 * fc[1] + o[1], nf[1]), fk = wrap(2 * fc[2] + o[2], nf[2]);
              Ff += ff[d][fi + (long)nf[0] * (fj + (long)nf[1] * (fk + (long)nf[2] * c))];
            }
            r += (side ? -1.0 : 1.0) * dt * (Ff - Fc) / vol;
          }
      reg[i + (long)nc[0] * (j + (long)nc[1] * (k + (long)nc[2] * c))] = r;
    }
}

void orc_ns_set_turbulent_forcing(orc_ns* ns, int nmodes, int mode_start, int div_free_force, int array_size, const double* forcedata) {
  ns->set_turb(nmodes, mode_start, div_free_force, array_size, forcedata);
}

void orc_ns_params_default(orc_ns_params* p) {
  p->cfl = 0.7; p->visc_coef = 0.0; p->be_cn_theta = 0.5; p->change_max = 1.1; p->init_shrink = 1.0; p->fixed_dt = -1.0;
  p->gravity = 0.0; p->visc_tol = 1e-10; p->mac_tol = 1e-12; p->mac_abs_tol = 1e-16; p->proj_tol = 1e-12; p->proj_abs_tol = 1e-16;
  p->init_iter = 2; p->init_vel_iter = 1; p->do_init_proj = 1; p->use_forces_in_trans = 0; p->conservative_tracer = 0; p->verbose = 0; p->scal_diff_coef = 0.0; p->use_ppm = 0; p->do_scalminmax = 0; p->do_mom_diff = 0; p->bottom_solver = 0;
}

orc_ns* orc_ns_create(const int n[3], const double prob_lo[3], const double prob_hi[3], const orc_ns_params* p) {
  orc_ns* ns = new orc_ns();
  for (int d = 0; d < 3; ++d) { ns->n[d] = n[d]; ns->prob_lo[d] = prob_lo[d]; ns->dx[d] = (prob_hi[d] - prob_lo[d]) / n[d]; }
  ns->p = *p;
  ns->S_old.define(n, 5, 1); ns->S_new.define(n, 5, 1); ns->P_old.define(n, 1, 2); ns->P_new.define(n, 1, 2);
  ns->Gp_old.define(n, 3, 1); ns->Gp_new.define(n, 3, 1); ns->aofs.define(n, 5, 0);
  ns->rho_p.define(n, 1, 1); ns->rho_c.define(n, 1, 1); ns->rho_half.define(n, 1, 1);
  for (int d = 0; d < 3; ++d) { ns->umac[d].define(n, 1, 1); ns->eta[d].define(n, 1, 1); ns->eta[d].setval(p->visc_coef); ns->seta[d].define(n, 1, 1); ns->seta[d].setval(p->scal_diff_coef); }
  return ns;
}
/* physical boundaries of the level (geometry.is_periodic, ns.lo_bc / ns.hi_bc, the Dirichlet face values of NS.cpp:108-237):
 * call right after orc_ns_create.  Walls, symmetry planes, inflow and (without gravity: no set_outflow_bcs) outflow. */
void orc_ns_set_bc(orc_ns* ns, const int per[3], const int phys_lo[3], const int phys_hi[3], const double* bcv /*[6][5]*/) {
  for (int d = 0; d < 3; ++d) { ns->per[d] = per[d]; ns->phys_lo[d] = phys_lo[d]; ns->phys_hi[d] = phys_hi[d]; }
  if (bcv) for (int f = 0; f < 6; ++f) for (int c = 0; c < 5; ++c) ns->bcv[f][c] = bcv[5 * f + c];
}
void orc_ns_destroy(orc_ns* ns) { delete ns; }
/* which as orc_ns_get, padded: state ng 1, press ng 2 (nodal), gradp ng 1, umac ng 1 */
void orc_ns_get_padded(const orc_ns* ns, int which, double* out) {
  switch (which) {
    case 0: ns->S_new.store_padded(out); break;
    case 1: ns->P_new.store_padded(out); break;
    case 2: ns->Gp_new.store_padded(out); break;
    case 4: case 5: case 6: ns->umac[which - 4].store_padded(out); break;
    default: break;
  }
}

void orc_ns_init_prob(orc_ns* ns, int probtype, const double* pp, int npp) {
  // Source/prob/prob_init.cpp: 11 TaylorGreen :509-560, 5 DoubleShearLayer :346-405 (direction 1);
  // 100 = synthetic variable-density Taylor-Green (not in the reference)
  const double twopi = 2.0 * 3.14159265358979323846264338327950288;
  Arr& S = ns->S_new;
  const int* n = ns->n;
#pragma omp parallel for
  for (int k = 0; k < n[2]; ++k)
    for (int j = 0; j < n[1]; ++j)
      for (int i = 0; i < n[0]; ++i) {
        const double x = ns->prob_lo[0] + (i + 0.5) * ns->dx[0], y = ns->prob_lo[1] + (j + 0.5) * ns->dx[1], z = ns->prob_lo[2] + (k + 0.5) * ns->dx[2];
        if (probtype == 11 || probtype == 100) {
          const double a = pp[0], b = pp[1], c = pp[2], vx = pp[3], dens = pp[4];
          S(i, j, k, 0) = vx * std::sin(a * twopi * x) * std::cos(b * twopi * y) * std::cos(c * twopi * z);
          S(i, j, k, 1) = -vx * std::cos(a * twopi * x) * std::sin(b * twopi * y) * std::cos(c * twopi * z);
          S(i, j, k, 2) = 0.0;
          S(i, j, k, 3) = (probtype == 100) ? dens * (1.0 + 0.5 * std::sin(twopi * x) * std::sin(twopi * y) * std::sin(twopi * z)) : dens;
          S(i, j, k, 4) = (dens * vx * vx / 16.0) * (2.0 + std::cos(2.0 * c * twopi * z)) * (std::cos(2.0 * a * twopi * x) + std::cos(2.0 * b * twopi * y));
        } else if (probtype == 10) {   // RayleighTaylor 3-D (prob_init.cpp:447-487): rho_1, rho_2, tra_1, tra_2, interface_width, pertamp
          const double pi = 0.5 * twopi;
          const double Lx = n[0] * ns->dx[0], Ly = n[1] * ns->dx[1], splitz = 0.5 * (ns->prob_lo[2] + (ns->prob_lo[2] + n[2] * ns->dx[2]));
          const double ranampl = 2. * (0.6544437533747718 - 0.5), ranphse1 = 2. * pi * 0.1556190326530211, ranphse2 = 2. * pi * 0.4196144025537369;
          const double pert = ranampl * std::sin(2.0 * pi * x / Lx + ranphse1) * std::sin(2.0 * pi * y / Ly + ranphse2);
          const double pertheight = splitz - pp[5] * pert;
          S(i, j, k, 0) = 0.0; S(i, j, k, 1) = 0.0; S(i, j, k, 2) = 0.0;
          S(i, j, k, 3) = pp[0] + ((pp[1] - pp[0]) / 2.0) * (1.0 + std::tanh((z - pertheight) / pp[4]));
          S(i, j, k, 4) = pp[2] + ((pp[3] - pp[2]) / 2.0) * (1.0 + std::tanh((z - pertheight) / pp[4]));
        } else if (probtype == 1) {    // LidDrivenCavity: start from rest, density 1 (prob_init.cpp:102-109)
          S(i, j, k, 0) = 0.0; S(i, j, k, 1) = 0.0; S(i, j, k, 2) = 0.0; S(i, j, k, 3) = 1.0; S(i, j, k, 4) = 0.0;
        } else if (probtype == 101) {  // synthetic wall-bounded test field (not in the reference): pp = amplitude, density, density variation
          const double pi = 0.5 * twopi, A = pp[0], dens = pp[1], vd = pp[2];
          const double X = (x - ns->prob_lo[0]) / (n[0] * ns->dx[0]), Y = (y - ns->prob_lo[1]) / (n[1] * ns->dx[1]), Z = (z - ns->prob_lo[2]) / (n[2] * ns->dx[2]);
          S(i, j, k, 0) = A * std::sin(pi * X) * std::cos(twopi * Y) * std::cos(pi * Z);
          S(i, j, k, 1) = -A * std::cos(pi * X) * std::sin(twopi * Y) * std::cos(twopi * Z) * 0.5;
          S(i, j, k, 2) = A * 0.3 * std::sin(twopi * X) * std::sin(twopi * Y) * std::sin(pi * Z);
          S(i, j, k, 3) = dens * (1.0 + vd * std::cos(twopi * X) * std::cos(twopi * Y) * std::cos(pi * Z));
          S(i, j, k, 4) = std::exp(-20.0 * ((X - 0.4) * (X - 0.4) + (Y - 0.5) * (Y - 0.5) + (Z - 0.6) * (Z - 0.6)));
        } else if (probtype == 20) {   // Tutorials/HIT/prob_init.cpp:100-131 (+ optional synthetic density variation pp[2])
          const double ts = pp[0], dens = pp[1], vd = npp > 2 ? pp[2] : 0.0;
          const double Lx = n[0] * ns->dx[0], Ly = n[1] * ns->dx[1], Lz = ns->prob_lo[2] + n[2] * ns->dx[2] - ns->prob_lo[1];   // :113
          S(i, j, k, 0) = ts * std::cos(twopi * y / Ly) * std::cos(twopi * z / Lz);
          S(i, j, k, 1) = ts * std::cos(twopi * x / Lx) * std::cos(twopi * z / Lz);
          S(i, j, k, 2) = ts * std::cos(twopi * x / Lx) * std::cos(twopi * y / Ly);
          S(i, j, k, 3) = dens * (1.0 + vd * std::sin(twopi * x / Lx) * std::sin(twopi * y / Ly) * std::sin(twopi * z / Lz));
          S(i, j, k, 4) = 1.0;
        } else {
          const double dens = pp[0], width = pp[1] > 0 ? pp[1] : 1.0, pi = 0.5 * twopi;
          S(i, j, k, 0) = -0.05 * std::sin(pi * y);
          S(i, j, k, 1) = std::tanh(30.0 * (0.5 - std::fabs(x)) / width);
          S(i, j, k, 2) = 0.0;
          S(i, j, k, 3) = dens;
          const double dist = std::sqrt((x - pp[2]) * (x - pp[2]) + (y - pp[3]) * (y - pp[3]) + (z - pp[4]) * (z - pp[4]));
          S(i, j, k, 4) = dist < pp[5] ? 1.0 : 0.0;
        }
      }
  ns->P_new.setval(0); ns->P_old.setval(0); ns->Gp_new.setval(0); ns->Gp_old.setval(0);
  ns->time = 0; ns->nstep = 0; ns->dt_level = 0; ns->dt_min = 1e100;
}

int orc_ns_post_init(orc_ns* ns, double* dt0) {
  PerScope ps(ns->per);
  const int* n = ns->n;
  if (ns->p.do_init_proj) {  // initialVelocityProject Projection.cpp:615-838 (sigma = 1)
    for (int it = 0; it < ns->p.init_vel_iter; ++it) {
      Arr vel(n, 3, 1), sig(n, 1, 1), phi(n, 1, 2);
      vel.copy_from(ns->S_new, 0, 0, 3); sig.setval(1.0);
      ns->inflow_ghost_velocity(vel, 1.0);
      orc_mg m = ns->mg(ns->p.proj_tol, ns->p.proj_abs_tol);
      const int rc = nodal_project(n, ns->dx, vel, sig, phi, nullptr, false, &m, ns->nodal_bc());
      if (rc) return rc;
      ns->S_new.copy_from(vel, 0, 0, 3);
      ns->P_old.setval(0); ns->P_new.setval(0); ns->Gp_old.setval(0); ns->Gp_new.setval(0);
    }
  }
  ns->initial_step = true;
  if (ns->p.do_init_proj && std::fabs(ns->p.gravity) > 0.0 && ns->has_walls()) {
    // initialPressureProject (NSB.cpp:2421-2431 -> Projection.cpp:841-960): the projection of the uniform gravity vector with
    // sigma = 1/rho gives the hydrostatic pressure; P and Gradp old := new.  (On a fully periodic domain div(0,0,g) = 0.)
    Arr vel(n, 3, 1), sig(n, 1, 1);
    vel.setval(0.0);
    { const double g = ns->p.gravity; for (int k = -1; k <= n[2]; ++k) for (int j = -1; j <= n[1]; ++j) for (int i = -1; i <= n[0]; ++i) vel(i, j, k, 2) = g; }
    { const Arr& S = ns->S_new; FOR_CELLS(sig, i, j, k) sig(i, j, k) = 1.0 / S(i, j, k, orc_ns::Density); }
    orc_mg m = ns->mg(ns->p.proj_tol, ns->p.proj_abs_tol);
    ns->P_new.setval(0.0);
    ns->set_outflow_bcs(ns->P_new, ns->S_new, orc_ns::Density);   // Projection.cpp:893-903
    const int rc = nodal_project(n, ns->dx, vel, sig, ns->P_new, &ns->Gp_new, false, &m, ns->nodal_bc());
    if (rc) return rc;
    ns->fill_gradp(ns->Gp_new);
    ns->P_old = ns->P_new; ns->Gp_old = ns->Gp_new;
  }
  double est = 0;
  if (ns->est_time_step(&est)) return -1;
  const double dt_init = ns->p.init_shrink * est;  // post_init_estDT NSB.cpp:2307-2366
  if (ns->p.init_iter > 0) {  // post_init_press NS.cpp:1306-1432
    ns->initial_iter = true;
    for (int iter = 0; iter < ns->p.init_iter; ++iter) {
      double dtt;
      int rc = ns->advance(dt_init, &dtt);
      if (rc) return rc;
      // initialSyncProject Projection.cpp:970-1185
      Arr vel(n, 3, 1), sig(n, 1, 1), phi(n, 1, 2);
      for (int c = 0; c < 3; ++c) { FOR_CELLS(vel, i, j, k) vel(i, j, k, c) = ns->S_new(i, j, k, c) * (1.0 / dt_init) + (-1.0 / dt_init) * ns->S_old(i, j, k, c); }
      FOR_CELLS(sig, i, j, k) sig(i, j, k) = 1.0 / ns->rho_half(i, j, k);
      orc_mg m = ns->mg(ns->p.proj_tol, ns->p.proj_abs_tol);
      rc = nodal_project(n, ns->dx, vel, sig, phi, &ns->Gp_new, true, &m, ns->nodal_bc());
      ns->it[2] = m.iters;
      if (rc) return rc;
      { Arr& P = ns->P_new; FOR_G1(P, i, j, k) P(i, j, k) += phi(i, j, k); }
      ns->fill_gradp(ns->Gp_new);
      std::swap(ns->S_old, ns->S_new);  // resetState NSB.cpp:2643-2680
      ns->P_old = ns->P_new; ns->Gp_old = ns->Gp_new;
      ns->initial_iter = false;
    }
  }
  ns->initial_step = false;
  ns->dt_level = dt_init; ns->dt_min = 1e100;
  if (dt0) *dt0 = dt_init;
  return 0;
}

int orc_ns_step(orc_ns* ns, double* dt_io) {
  PerScope ps(ns->per);
  double dt = (dt_io && *dt_io > 0.0) ? *dt_io : -1.0;
  if (dt <= 0.0) {
    if (ns->p.fixed_dt > 0.0) dt = ns->p.fixed_dt;
    else if (ns->nstep == 0 && ns->dt_level > 0.0) dt = ns->dt_level;
    else {  // computeNewDt NSB.cpp:945-1036
      double est; if (ns->est_time_step(&est)) return -1;
      dt = std::min(ns->dt_min, est);
      if (ns->dt_level > 0.0) dt = std::min(dt, ns->p.change_max * ns->dt_level);
    }
  }
  double dtt = 0;
  const int rc = ns->advance(dt, &dtt);
  if (rc) return rc;
  ns->dt_min = dtt; ns->dt_level = dt; ns->time += dt; ns->nstep++;
  if (dt_io) *dt_io = dt;
  return 0;
}

double orc_ns_time(const orc_ns* ns) { return ns->time; }

void orc_ns_get(const orc_ns* ns, int which, double* out) {
  switch (which) {
    case 0: ns->S_new.store(out); break;
    case 1: ns->P_new.store(out); break;
    case 2: ns->Gp_new.store(out); break;
    case 4: case 5: case 6: ns->umac[which - 4].store(out); break;
    case 7: ns->aofs.store(out); break;
    default: break;
  }
}
void orc_ns_set_state(orc_ns* ns, const double* s) { ns->S_new.load(s); }
void orc_ns_last_iters(const orc_ns* ns, int it[3]) { for (int q = 0; q < 3; ++q) it[q] = ns->it[q]; }

}  // extern "C"
