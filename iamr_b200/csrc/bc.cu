// bc.cu -- physical domain boundaries: the ghost-cell fills and boundary rows that the periodic
// round-1 path did not need.
//
//   fill_physbc      : AmrLevel::FillPatch's physical-boundary fill of cell-centred state data
//                      (amrex filcc_cell + the ext_dir user functions of NS_bcfill.H:17-167; BC tables NS_BC.H:7-55)
//   linop_bc_fill    : MLCellLinOp::applyBC at domain faces -- Dirichlet (Lagrange extrapolation through the face value,
//                      order min(maxorder, box length + 1)), Neumann, reflect_odd (Diffusion.cpp:1887-1999,
//                      MacProj.cpp:1187-1208; maxorder MacProj.cpp:30, Diffusion.cpp:95-96)
//   nodal_bc_fill_*  : MLNodeLaplacian's Neumann / inflow sides: reflected ghost node, copied ghost sigma
//                      (Projection.cpp:2436-2464)
//   nodal_bc_scale   : mlndlap_impose_neumann_bc (x2 per Neumann side on the boundary rows of the right-hand side)
// The ghost regions are thin slabs, so every kernel is launched over one slab (one side of one box) at a time.
#include "kernels.h"

namespace ix {
namespace k {
namespace {

constexpr int TX = 64;
constexpr int TY = 4;
inline dim3 grid_for(const Bx& bx, int nzc) { return dim3(cdiv(bx.nx(), TX), cdiv(bx.ny(), TY), nzc); }
#define IDX3(bx)                                                     \
  const int nz_ = bx.hi[2] - bx.lo[2] + 1;                            \
  const int k = bx.lo[2] + (int)(blockIdx.z % nz_);                   \
  const int n = (int)(blockIdx.z / nz_);                              \
  const int j = bx.lo[1] + blockIdx.y * TY + threadIdx.y;             \
  const int i = bx.lo[0] + blockIdx.x * TX + threadIdx.x;             \
  if (j > bx.hi[1] || i > bx.hi[0]) return;

// ---- state fill ---------------------------------------------------------------------------------------------------
// one slab = the cells of the fab outside the domain on `side` of direction `sd`; `pass` = number of non-periodic
// directions the cell must be outside of (1 faces, 2 edges, 3 corners: each pass reads what the previous one wrote).
// A cell outside in several directions belongs to several slabs: it is handled in the slab of its FIRST such direction.
__global__ void __launch_bounds__(TX* TY)
physbc_kernel(Bx R, V4 a, IX_KARG(PhysBC) bc, Bx dom, int p0, int p1, int p2, int sd, int pass) {
  IDX3(R)
  const int per[3] = {p0, p1, p2};
  const int idx[3] = {i, j, k};
  int nout = 0, side[3] = {0, 0, 0}, first = -1;
  for (int d = 0; d < 3; ++d)
    if (!per[d]) {
      if (idx[d] < dom.lo[d]) { side[d] = -1; ++nout; if (first < 0) first = d; }
      else if (idx[d] > dom.hi[d]) { side[d] = 1; ++nout; if (first < 0) first = d; }
    }
  if (nout != pass || first != sd) return;
  double v = a(i, j, k, n);
  for (int d = 0; d < 3; ++d) {
    if (!side[d]) continue;
    const int code = side[d] < 0 ? bc.lo[n][d] : bc.hi[n][d];
    const int e = side[d] < 0 ? dom.lo[d] : dom.hi[d];   // first interior cell on that side
    const int inw = side[d] < 0 ? 1 : -1;
    const int dist = side[d] < 0 ? dom.lo[d] - 1 - idx[d] : idx[d] - dom.hi[d] - 1;   // 0 = first ghost cell
    auto at = [&](int m) { int q[3] = {i, j, k}; q[d] = m; return a(q[0], q[1], q[2], n); };
    const int nd = dom.hi[d] - dom.lo[d] + 1;
    if (code == IAMRX_BC_FOEXTRAP) v = at(e);
    else if (code == IAMRX_BC_HOEXTRAP) {
      if (dist > 0) v = at(e);
      else if (nd >= 3) v = 0.125 * (15.0 * at(e) - 10.0 * at(e + inw) + 3.0 * at(e + 2 * inw));
      else v = 0.5 * (3.0 * at(e) - at(e + inw));
    } else if (code == IAMRX_BC_REFLECT_EVEN) v = at(e + inw * dist);
    else if (code == IAMRX_BC_REFLECT_ODD) v = -at(e + inw * dist);
    a(i, j, k, n) = v;
  }
  for (int d = 0; d < 3; ++d) {
    if (!side[d]) continue;
    const int code = side[d] < 0 ? bc.lo[n][d] : bc.hi[n][d];
    if (code == IAMRX_BC_EXT_DIR) v = bc.val[d + (side[d] > 0 ? 3 : 0)][n];
  }
  a(i, j, k, n) = v;
}

// ---- cell-centred linear operators -----------------------------------------------------------------------------------
// Lagrange weights at the ghost centre (-1/2 cell from the face) for the nodes {face, 1/2, 3/2, ...}
struct DirW { double w[5]; };
// x0: location of the Dirichlet value in cell widths from the face (0: on the face; -ratio/2 scaled to the multigrid level: the
// coarse-fine sides of a fine AMR level, whose data sit at the coarse cell centres)
inline DirW dirichlet_weights(int order, double x0 = 0.0) {
  DirW r;
  double x[5]; x[0] = x0;
  for (int m = 1; m < order; ++m) x[m] = m - 0.5;
  for (int m = 0; m < order; ++m) {
    double num = 1.0, den = 1.0;
    for (int q = 0; q < order; ++q) if (q != m) { num *= (-0.5 - x[q]); den *= (x[m] - x[q]); }
    r.w[m] = num / den;
  }
  for (int m = order; m < 5; ++m) r.w[m] = 0.0;
  return r;
}

// R = the ghost slab (one layer) beyond side `side` of direction d; e = first interior cell index along d
__global__ void __launch_bounds__(TX* TY)
linop_bc_kernel(Bx R, V4 phi, C4 bv, int d, int side, int e, int code0, int code1, int code2, IX_KARG(DirW) w) {
  IDX3(R)
  const int code = n == 0 ? code0 : (n == 1 ? code1 : code2);
  const int inw = side < 0 ? 1 : -1;
  auto P = [&](int m) { int q[3] = {i, j, k}; q[d] = m; return phi(q[0], q[1], q[2], n); };
  double v;
  if (code == IAMRX_LINOP_NEUMANN) v = P(e);
  else if (code == IAMRX_LINOP_REFLECT_ODD) v = -P(e);
  else if (code == IAMRX_LINOP_DIRICHLET) {
    v = w.w[0] * (bv.ok() ? bv(i, j, k, n) : 0.0);
#pragma unroll
    for (int m = 1; m < 5; ++m) if (w.w[m] != 0.0) v += w.w[m] * P(e + inw * (m - 1));
  } else return;
  phi(i, j, k, n) = v;
}

// ---- nodal operator ----------------------------------------------------------------------------------------------------
// ghost node plane g := mirror plane m along d (phi), ghost cell layer := first interior layer (sigma)
__global__ void __launch_bounds__(TX* TY) mirror_kernel(Bx R, V4 a, int d, int src) {
  IDX3(R)
  int q[3] = {i, j, k}; q[d] = src;
  a(i, j, k, n) = a(q[0], q[1], q[2], n);
}
__global__ void __launch_bounds__(TX* TY) scale_plane_kernel(Bx R, V4 a, double f) {
  IDX3(R)
  a(i, j, k, n) *= f;
}

// a(node) = 0 where m(node) != 0 (the Dirichlet nodes of a nodal solve on a fine AMR level of general shape)
__global__ void __launch_bounds__(TX* TY) mask_zero_kernel(Bx R, V4 a, C4 m) {
  IDX3(R)
  if (m(i, j, k) != 0.0) a(i, j, k, n) = 0.0;
}

}  // namespace

int mask_zero(const Bx& R, V4 a, C4 m, cudaStream_t s) {
  IX_LAUNCH(mask_zero_kernel, grid_for(R, R.nz()), dim3(TX, TY, 1), 0, s, R, a, m);
  return check_launch("mask_zero");
}

int fill_physbc(const Bx& fabbox, V4 a, int ncomp, const PhysBC& bc, const Bx& dom, const int per[3], cudaStream_t s) {
  int nnp = 0;
  for (int d = 0; d < 3; ++d) if (!per[d]) ++nnp;
  for (int pass = 1; pass <= nnp; ++pass)
    for (int d = 0; d < 3; ++d) {
      if (per[d]) continue;
      for (int side = -1; side <= 1; side += 2) {
        Bx R = fabbox;
        if (side < 0) R.hi[d] = std::min(R.hi[d], dom.lo[d] - 1); else R.lo[d] = std::max(R.lo[d], dom.hi[d] + 1);
        if (!R.ok()) continue;
        // a pass > 1 slab only holds work if the fab also sticks out in another non-periodic direction
        if (pass > 1) {
          bool other = false;
          for (int q = 0; q < 3; ++q) if (q != d && !per[q] && (fabbox.lo[q] < dom.lo[q] || fabbox.hi[q] > dom.hi[q])) other = true;
          if (!other) continue;
        }
        IX_LAUNCH(physbc_kernel, grid_for(R, R.nz() * ncomp), dim3(TX, TY, 1), 0, s, R, a, bc, dom, per[0], per[1], per[2], d, pass);
        const int rc = check_launch("fill_physbc");
        if (rc) return rc;
      }
    }
  return IAMRX_OK;
}

int linop_bc_order(int maxorder, int boxlen) { return std::max(2, std::min(std::min(maxorder, boxlen + 1), 4)); }

double linop_bc_f0(int code, int maxorder, int boxlen) {
  if (code == IAMRX_LINOP_NEUMANN) return 1.0;
  if (code == IAMRX_LINOP_REFLECT_ODD) return -1.0;
  if (code == IAMRX_LINOP_DIRICHLET) return dirichlet_weights(linop_bc_order(maxorder, boxlen)).w[1];
  return 0.0;
}

int linop_bc_fill(const Bx& vbx, V4 phi, int ncomp, const LinBC& bc, C4 bv, const Bx& dom, const int per[3], int grow_t, int skipmask,
                  cudaStream_t s) {
  for (int d = 0; d < 3; ++d) {
    if (per[d] || (skipmask & (1 << d))) continue;
    for (int side = -1; side <= 1; side += 2) {
      if (side < 0 ? vbx.lo[d] != dom.lo[d] : vbx.hi[d] != dom.hi[d]) continue;   // this side of the box is not a domain face
      Bx R = vbx;
      for (int q = 0; q < 3; ++q) {
        if (q == d) continue;
        // transverse growth (the tensor cross terms read one cell sideways): only into cells that exist -- inside the
        // domain, or periodic images
        R.lo[q] -= grow_t; R.hi[q] += grow_t;
        if (!per[q]) { R.lo[q] = std::max(R.lo[q], dom.lo[q]); R.hi[q] = std::min(R.hi[q], dom.hi[q]); }
      }
      R.lo[d] = R.hi[d] = side < 0 ? dom.lo[d] - 1 : dom.hi[d] + 1;
      const int e = side < 0 ? dom.lo[d] : dom.hi[d];
      const int c0 = side < 0 ? bc.lo[0][d] : bc.hi[0][d], c1 = side < 0 ? bc.lo[1][d] : bc.hi[1][d], c2 = side < 0 ? bc.lo[2][d] : bc.hi[2][d];
      const DirW w = dirichlet_weights(linop_bc_order(bc.maxorder, vbx.hi[d] - vbx.lo[d] + 1));
      IX_LAUNCH(linop_bc_kernel, grid_for(R, R.nz() * ncomp), dim3(TX, TY, 1), 0, s, R, phi, bv, d, side, e, c0, ncomp > 1 ? c1 : c0,
                ncomp > 2 ? c2 : c0, w);
      const int rc = check_launch("linop_bc_fill");
      if (rc) return rc;
    }
  }
  return IAMRX_OK;
}

// ---- coarse-fine sides of a fine AMR level (MLCellLinOp with setCoarseFineBC, MacProj.cpp:1164-1167) -------------------------
// AMReX (MLMGBndry::setBoxBC) treats a box side that is not a domain face as Dirichlet with the value half a COARSE cell beyond the
// face; the ghost cell is the Lagrange extrapolation through that value and the first interior cells, as on a Dirichlet domain side.
double linop_cf_f0(int maxorder, int boxlen, double x0) { return dirichlet_weights(linop_bc_order(maxorder, boxlen), x0).w[1]; }

// cfmask bit 2 d + side: that side of vbx is a coarse-fine side; the whole one-cell layer beyond it is written (the caller's
// FillBoundary then overwrites the cells a fine neighbour covers).  bv: fab whose ghost cells hold the coarse-fine boundary
// values (iamrx_set_coarse_fine_bc); null: homogeneous.
int linop_cf_fill(const Bx& vbx, V4 phi, int ncomp, int maxorder, C4 bv, int cfmask, const double x0[3], cudaStream_t s) {
  for (int d = 0; d < 3; ++d)
    for (int side = 0; side < 2; ++side) {
      if (!(cfmask & (1 << (2 * d + side)))) continue;
      Bx R = vbx;
      R.lo[d] = R.hi[d] = side == 0 ? vbx.lo[d] - 1 : vbx.hi[d] + 1;
      const int e = side == 0 ? vbx.lo[d] : vbx.hi[d];
      const DirW w = dirichlet_weights(linop_bc_order(maxorder, vbx.hi[d] - vbx.lo[d] + 1), x0[d]);
      IX_LAUNCH(linop_bc_kernel, grid_for(R, R.nz() * ncomp), dim3(TX, TY, 1), 0, s, R, phi, bv, d, side == 0 ? -1 : 1, e,
                IAMRX_LINOP_DIRICHLET, IAMRX_LINOP_DIRICHLET, IAMRX_LINOP_DIRICHLET, w);
      const int rc = check_launch("linop_cf_fill");
      if (rc) return rc;
    }
  return IAMRX_OK;
}

// Neumann / inflow sides of a node box: ghost node plane = mirror image (node lo-1 := lo+1, hi+1 := hi-1)
int nodal_bc_fill_phi(const Bx& nbx, V4 phi, const NodalBC& bc, const Bx& ndom, const int per[3], int skipmask, cudaStream_t s) {
  for (int d = 0; d < 3; ++d) {
    if (per[d] || (skipmask & (1 << d))) continue;
    for (int side = -1; side <= 1; side += 2) {
      const int code = side < 0 ? bc.lo[d] : bc.hi[d];
      if (code != IAMRX_LINOP_NEUMANN && code != IAMRX_LINOP_INFLOW) continue;
      if (side < 0 ? nbx.lo[d] != ndom.lo[d] : nbx.hi[d] != ndom.hi[d]) continue;
      Bx R = grow(nbx, 1);   // the full ghost plane incl. its edges (their sources are ghost nodes filled before: x, then y, then z)
      R.lo[d] = R.hi[d] = side < 0 ? ndom.lo[d] - 1 : ndom.hi[d] + 1;
      IX_LAUNCH(mirror_kernel, grid_for(R, R.nz()), dim3(TX, TY, 1), 0, s, R, phi, d, side < 0 ? ndom.lo[d] + 1 : ndom.hi[d] - 1);
      const int rc = check_launch("nodal_bc_fill_phi");
      if (rc) return rc;
    }
  }
  return IAMRX_OK;
}

// sigma ghost cell layer beyond a Neumann / inflow side = the first interior layer (mlndlap_fillbc_cc); zero beyond Dirichlet
// ngt: width of the ghost plane in the tangential directions (the fab's ghost depth; deep-ghost multigrid levels: 4)
int nodal_bc_fill_sigma(const Bx& cbx, V4 sig, const NodalBC& bc, const Bx& dom, const int per[3], cudaStream_t s, int ngt) {
  for (int d = 0; d < 3; ++d) {
    if (per[d]) continue;
    for (int side = -1; side <= 1; side += 2) {
      const int code = side < 0 ? bc.lo[d] : bc.hi[d];
      if (side < 0 ? cbx.lo[d] != dom.lo[d] : cbx.hi[d] != dom.hi[d]) continue;
      Bx R = grow(cbx, ngt);
      R.lo[d] = R.hi[d] = side < 0 ? dom.lo[d] - 1 : dom.hi[d] + 1;
      if (code == IAMRX_LINOP_NEUMANN || code == IAMRX_LINOP_INFLOW) {
        IX_LAUNCH(mirror_kernel, grid_for(R, R.nz()), dim3(TX, TY, 1), 0, s, R, sig, d, side < 0 ? dom.lo[d] : dom.hi[d]);
      } else {
        IX_LAUNCH(scale_plane_kernel, grid_for(R, R.nz()), dim3(TX, TY, 1), 0, s, R, sig, 0.0);
      }
      const int rc = check_launch("nodal_bc_fill_sigma");
      if (rc) return rc;
    }
  }
  return IAMRX_OK;
}

// rows ON a Neumann / inflow side: multiply by f once per such direction (mlndlap_impose_neumann_bc: f = 2; the dot-mask
// weights of the solvability sum: f = 1/2)
int nodal_bc_scale(const Bx& nbx, V4 a, const NodalBC& bc, const Bx& ndom, const int per[3], double f, cudaStream_t s) {
  for (int d = 0; d < 3; ++d) {
    if (per[d]) continue;
    for (int side = -1; side <= 1; side += 2) {
      const int code = side < 0 ? bc.lo[d] : bc.hi[d];
      if (code != IAMRX_LINOP_NEUMANN && code != IAMRX_LINOP_INFLOW) continue;
      if (side < 0 ? nbx.lo[d] != ndom.lo[d] : nbx.hi[d] != ndom.hi[d]) continue;
      Bx R = nbx;
      R.lo[d] = R.hi[d] = side < 0 ? ndom.lo[d] : ndom.hi[d];
      IX_LAUNCH(scale_plane_kernel, grid_for(R, R.nz()), dim3(TX, TY, 1), 0, s, R, a, f);
      const int rc = check_launch("nodal_bc_scale");
      if (rc) return rc;
    }
  }
  return IAMRX_OK;
}

}  // namespace k
}  // namespace ix
