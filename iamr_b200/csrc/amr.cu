// amr.cu -- inter-level transfer operators and the coarse-fine flux register of the two-level coupling:
//   average_down         : NavierStokesBase::avgDown_StatePress (NSB.cpp:4125-4191) -- cells: mean of the 2^3 children
//                          (amrex::average_down), faces: mean of the 2^2 fine faces on the coarse face (average_down_faces),
//                          nodes: injection (average_down_nodal, NSB.cpp:4154)
//   interp cell_cons     : cell_cons_interp of State_Type (NS_setup.cpp:211,228-230): conservative linear interpolation with
//                          monotonised-central slopes per direction, scaled so that no fine value leaves the range of the 3^3
//                          coarse neighbourhood (CellConservativeLinear without linear limiting: mclim + mmlim)
//   interp node_bilinear : node_bilinear_interp of Press_Type (NS_setup.cpp:331)
//   interp face_linear   : face_linear_interp of u_mac in create_umac_grown (NSB.cpp:1127): linear along the face normal
//                          between the two coarse faces, piecewise constant across
//   flux register        : the advective register's CrseAdd / FineAdd / Reflux with area-weighted fluxes and dx := cell volume
//                          (NSB.cpp:4848-4889, 5083-5096; NS.cpp:1794-1795): on the coarse cells that border the fine grids from
//                          outside, reg = dt (sum of fine fluxes - coarse flux) / vol_crse * (+-1); Reflux adds it to the state.
// Refinement ratio 2 (amr.ref_ratio = 2 in every BASELINE config).  These are the building blocks; the two-level time stepping
// itself (FillPatchTwoLevels in time, coarse-fine solver boundaries, mac_sync / level_sync) is not driven by this library.
#include <algorithm>
#include "level.h"
#include "mlmg.h"

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

IX_HD int coarsen2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }   // floor(a/2)

__global__ void __launch_bounds__(TX* TY) node_inject_kernel(Bx cnbx, V4 crse, C4 fine) {
  IDX3(cnbx)
  crse(i, j, k, n) = fine(2 * i, 2 * j, 2 * k, n);
}

IX_D double mc_slope(double um, double u0, double up) {
  const double dc = 0.5 * (up - um), df = 2.0 * (up - u0), db = 2.0 * (u0 - um);
  const double lim = (df * db >= 0.0) ? fmin(fabs(df), fabs(db)) : 0.0;
  return copysign(1.0, dc) * fmin(lim, fabs(dc));
}

// one thread per FINE cell: slopes of its coarse parent are recomputed (27 coarse loads, served by L1: 8 siblings share them)
__global__ void __launch_bounds__(TX* TY) cell_cons_interp_kernel(Bx fbx, V4 fine, C4 crse) {
  IDX3(fbx)
  const int I = coarsen2(i), J = coarsen2(j), K = coarsen2(k);
  const double u = crse(I, J, K, n);
  double sx = mc_slope(crse(I - 1, J, K, n), u, crse(I + 1, J, K, n));
  double sy = mc_slope(crse(I, J - 1, K, n), u, crse(I, J + 1, K, n));
  double sz = mc_slope(crse(I, J, K - 1, n), u, crse(I, J, K + 1, n));
  double cmn = u, cmx = u;
  for (int kk = -1; kk <= 1; ++kk)
    for (int jj = -1; jj <= 1; ++jj)
      for (int ii = -1; ii <= 1; ++ii) { const double v = crse(I + ii, J + jj, K + kk, n); cmn = fmin(cmn, v); cmx = fmax(cmx, v); }
  // largest excursion of the eight children: sum of |slope| / 4
  const double dmax = 0.25 * (fabs(sx) + fabs(sy) + fabs(sz));
  double alpha = 1.0;
  if (dmax > 0.0) {
    if (u + dmax > cmx) alpha = fmin(alpha, (cmx - u) / dmax);
    if (u - dmax < cmn) alpha = fmin(alpha, (u - cmn) / dmax);
  }
  const double ox = (i - 2 * I) ? 0.25 : -0.25, oy = (j - 2 * J) ? 0.25 : -0.25, oz = (k - 2 * K) ? 0.25 : -0.25;
  fine(i, j, k, n) = u + alpha * (sx * ox + sy * oy + sz * oz);
}

// pc_interp (piecewise constant): every fine cell takes its parent's value
__global__ void __launch_bounds__(TX* TY) pc_interp_kernel(Bx fbx, V4 fine, C4 crse, int nz) {
  const int kz = blockIdx.z % nz, n = blockIdx.z / nz;
  const int k = fbx.lo[2] + kz;
  const int j = fbx.lo[1] + blockIdx.y * TY + threadIdx.y;
  const int i = fbx.lo[0] + blockIdx.x * TX + threadIdx.x;
  if (j > fbx.hi[1] || i > fbx.hi[0]) return;
  auto fl = [](int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); };
  fine(i, j, k, n) = crse(fl(i), fl(j), fl(k), n);
}

__global__ void __launch_bounds__(TX* TY) node_bilinear_interp_kernel(Bx fnbx, V4 fine, C4 crse) {
  IDX3(fnbx)
  const int I = coarsen2(i), J = coarsen2(j), K = coarsen2(k);
  const int ox = i - 2 * I, oy = j - 2 * J, oz = k - 2 * K;
  double acc = 0.0;
  for (int dk = 0; dk <= oz; ++dk)
    for (int dj = 0; dj <= oy; ++dj)
      for (int di = 0; di <= ox; ++di) acc += crse(I + di, J + dj, K + dk, n);
  fine(i, j, k, n) = acc / (double)((1 + ox) * (1 + oy) * (1 + oz));
}

__global__ void __launch_bounds__(TX* TY) face_linear_interp_kernel(Bx ffbx, int dir, V4 fine, C4 crse) {
  IDX3(ffbx)
  const int I = coarsen2(i), J = coarsen2(j), K = coarsen2(k);
  const int idx = dir == 0 ? i : (dir == 1 ? j : k);
  const double c0 = crse(I, J, K, n);
  if ((idx & 1) == 0) fine(i, j, k, n) = c0;   // on a coarse face
  else fine(i, j, k, n) = 0.5 * (c0 + crse(I + (dir == 0), J + (dir == 1), K + (dir == 2), n));
}

// ---- coarse-fine boundary values of a fine-level solve (InterpBndryData::setBndryValues, order 3) ---------------------------------
// R = the one-cell layer of fine ghost cells beyond side d of a fine box.  The value of each cell is the coarse field at the fine
// cell's TANGENTIAL position in the plane of the coarse cell centres: c0 + y dy + y^2 d2y + z dz + z^2 d2z + y z dyz, with
// centred differences where both tangential coarse neighbours are usable (mask: inside the domain and not under the fine level),
// one-sided first differences where one is, and the mixed term where all four diagonal neighbours are.
__global__ void __launch_bounds__(TX* TY) cf_bndry_kernel(Bx R, V4 fine, C4 crse, C4 mask, int d) {
  IDX3(R)
  const int f[3] = {i, j, k};
  const int I[3] = {coarsen2(i), coarsen2(j), coarsen2(k)};
  if (mask(I[0], I[1], I[2]) == 0.0) return;   // beyond a physical side, or under another fine box
  const int t[2] = {(d + 1) % 3, (d + 2) % 3};
  const double c0 = crse(I[0], I[1], I[2], n);
  double v = c0, off[2];
  for (int q = 0; q < 2; ++q) {
    int m[3] = {I[0], I[1], I[2]}, p[3] = {I[0], I[1], I[2]};
    m[t[q]] -= 1; p[t[q]] += 1;
    const bool um = mask(m[0], m[1], m[2]) != 0.0, up = mask(p[0], p[1], p[2]) != 0.0;
    const double x = (f[t[q]] - 2 * I[t[q]]) ? 0.25 : -0.25;
    off[q] = x;
    double d1 = 0.0, d2 = 0.0;
    if (um && up) {
      const double cm = crse(m[0], m[1], m[2], n), cp = crse(p[0], p[1], p[2], n);
      d1 = 0.5 * (cp - cm);
      d2 = 0.5 * (cp - 2.0 * c0 + cm);
    } else if (up) d1 = crse(p[0], p[1], p[2], n) - c0;
    else if (um) d1 = c0 - crse(m[0], m[1], m[2], n);
    v += x * d1 + x * x * d2;
  }
  bool all = true;
  double cr[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  for (int sb = 0; sb < 2; ++sb)
    for (int sa = 0; sa < 2; ++sa) {
      int q[3] = {I[0], I[1], I[2]};
      q[t[0]] += sa ? 1 : -1; q[t[1]] += sb ? 1 : -1;
      if (mask(q[0], q[1], q[2]) == 0.0) all = false; else cr[sb][sa] = crse(q[0], q[1], q[2], n);
    }
  if (all) v += off[0] * off[1] * 0.25 * (cr[1][1] - cr[1][0] - cr[0][1] + cr[0][0]);
  fine(i, j, k, n) = v;
}

// ---- sync register (SyncRegister.cpp) ---------------------------------------------------------------------------------------
// mask of InitRHS (SyncRegister.cpp:126-285): 0 on coarse nodes whose eight surrounding coarse cells are all under the fine level
// (cells beyond a non-periodic side count as their mirror images: the "double the cell contributions" step), 1 elsewhere
__global__ void __launch_bounds__(TX* TY) sync_mask_kernel(Bx nbx, V4 mask, C4 covered, double maxcount) {
  IDX3(nbx)
  double sum = 0.0;
  for (int dk = -1; dk <= 0; ++dk) for (int dj = -1; dj <= 0; ++dj) for (int di = -1; di <= 0; ++di) sum += covered(i + di, j + dj, k + dk);
  mask(i, j, k, n) = sum > maxcount ? 0.0 : 1.0;
}
// FineAdd (SyncRegister.cpp:351-607): the coarse nodes of ONE boundary plane (normal `dir`) of a coarsened fine node box take the
// fine residual restricted with the tent weights of the two tangential directions, (r - m)(r - n) r_dir / prod(r^2), halved on
// the centre lines; fine nodes on the edges of the fine box count half and on its corners a third (0.5 * 2/3), nodes outside the
// fine box are zero (the residual's ghost nodes); nodes on non-periodic domain planes are doubled once per such direction.
__global__ void __launch_bounds__(TX* TY) sync_fine_add_kernel(Bx R, V4 acc, C4 fine, Bx fnb, int dir, double mult, Bx cnd, int p0, int p1, int p2) {
  IDX3(R)
  const int dim1 = dir != 0 ? 0 : 1, dim2 = dir != 0 ? (dir == 2 ? 1 : 2) : 2;
  const int ic[3] = {i, j, k};
  const double denom = 2.0 / 64.0;
  auto F = [&](int a, int b, int c) -> double {
    if (a < fnb.lo[0] || a > fnb.hi[0] || b < fnb.lo[1] || b > fnb.hi[1] || c < fnb.lo[2] || c > fnb.hi[2]) return 0.0;
    const int nb = (a == fnb.lo[0] || a == fnb.hi[0]) + (b == fnb.lo[1] || b == fnb.hi[1]) + (c == fnb.lo[2] || c == fnb.hi[2]);
    const double w = nb >= 3 ? 0.5 * (2.0 / 3.0) : (nb == 2 ? 0.5 : 1.0);
    return w * fine(a, b, c, n);
  };
  double v = 0.0;
  for (int nn = 0; nn < 2; ++nn)
    for (int m = 0; m < 2; ++m) {
      double coeff = (2 - m) * (2 - nn) * denom;
      if (nn == 0) coeff *= 0.5;
      if (m == 0) coeff *= 0.5;
      int f0[3] = {2 * ic[0], 2 * ic[1], 2 * ic[2]}, f1[3] = {2 * ic[0], 2 * ic[1], 2 * ic[2]}, f2[3] = {2 * ic[0], 2 * ic[1], 2 * ic[2]},
          f3[3] = {2 * ic[0], 2 * ic[1], 2 * ic[2]};
      f0[dim1] += m; f0[dim2] += nn;
      f1[dim1] -= m; f1[dim2] += nn;
      f2[dim1] += m; f2[dim2] -= nn;
      f3[dim1] -= m; f3[dim2] -= nn;
      v += coeff * (F(f0[0], f0[1], f0[2]) + F(f1[0], f1[1], f1[2]) + F(f2[0], f2[1], f2[2]) + F(f3[0], f3[1], f3[2]));
    }
  const int per[3] = {p0, p1, p2};
  for (int q = 0; q < 3; ++q) if (!per[q] && (ic[q] == cnd.lo[q] || ic[q] == cnd.hi[q])) v *= 2.0;
  acc(i, j, k, n) += mult * v;
}
// reg(p) += onB(p) * sum over the periodic images of p inside the node domain of acc(image)  (FabSet::plusFrom with periodicity)
__global__ void __launch_bounds__(TX* TY) sync_gather_kernel(Bx nbx, V4 reg, C4 onb, C4 acc, Bx cnd, int l0, int l1, int l2) {
  IDX3(nbx)
  if (onb(i, j, k) == 0.0) return;
  double s = 0.0;
  for (int sz = -1; sz <= 1; ++sz) {
    if (sz != 0 && l2 == 0) continue;
    const int kk = k + sz * l2;
    if (kk < cnd.lo[2] || kk > cnd.hi[2]) continue;
    for (int sy = -1; sy <= 1; ++sy) {
      if (sy != 0 && l1 == 0) continue;
      const int jj = j + sy * l1;
      if (jj < cnd.lo[1] || jj > cnd.hi[1]) continue;
      for (int sx = -1; sx <= 1; ++sx) {
        if (sx != 0 && l0 == 0) continue;
        const int ii = i + sx * l0;
        if (ii < cnd.lo[0] || ii > cnd.hi[0]) continue;
        s += acc(ii, jj, kk, n);
      }
    }
  }
  reg(i, j, k, n) += s;
}

// ---- flux register ---------------------------------------------------------------------------------------------------
// R = coarse cells of one interface patch (a one-cell-thick slab just OUTSIDE a fine box, on side `side` of direction d);
// fc = index of the coarse face between the slab and the fine region.
__global__ void __launch_bounds__(TX* TY) fr_crse_add_kernel(Bx R, V4 reg, C4 flux, int d, int fc_off, double scale) {
  IDX3(R)
  // coarse flux through the face the cell shares with the fine region (fc_off = 1: its high face, 0: its low face)
  reg(i, j, k, n) += scale * flux(i + fc_off * (d == 0), j + fc_off * (d == 1), k + fc_off * (d == 2), n);
}
__global__ void __launch_bounds__(TX* TY) fr_fine_add_kernel(Bx R, V4 reg, C4 flux, int d, int fc_off, double scale) {
  IDX3(R)
  // the 2 x 2 fine faces that tile the coarse face
  const int fi = 2 * (i + fc_off * (d == 0)), fj = 2 * (j + fc_off * (d == 1)), fk = 2 * (k + fc_off * (d == 2));
  double sum = 0.0;
  for (int b = 0; b < 2; ++b)
    for (int a = 0; a < 2; ++a) {
      const int ii = fi + (d == 0 ? 0 : a), jj = fj + (d == 1 ? 0 : (d == 0 ? a : b)), kk = fk + (d == 2 ? 0 : b);
      sum += flux(ii, jj, kk, n);
    }
  reg(i, j, k, n) += scale * sum;
}
__global__ void __launch_bounds__(TX* TY) fr_reflux_kernel(Bx R, V4 state, C4 reg, double scale) {
  IDX3(R)
  state(i, j, k, n) += scale * reg(i, j, k, n);
}

}  // namespace

// create_umac_grown's divergence correction (NSB.cpp:1203-1308, the non-RZ branch): a ghost cell that the fine level does not cover
// and that has exactly ONE face neighbour in the valid / covered region gets its OUTER face velocity from the divergence constraint.
// mask: 0 interior (valid cells of this box), 1 covered by another fine box, 2 not covered, 3 outside a physical boundary.
__global__ void __launch_bounds__(TX* TY) umac_divfix_kernel(Bx g1, Bx vb, C4 mask, V4 u, V4 v, V4 w, C4 divu, double dx0, double dx1,
                                                             double dx2) {
  const int k = g1.lo[2] + blockIdx.z;
  const int j = g1.lo[1] + blockIdx.y * TY + threadIdx.y;
  const int i = g1.lo[0] + blockIdx.x * TX + threadIdx.x;
  if (j > g1.hi[1] || i > g1.hi[0]) return;
  if (mask(i, j, k) != 2.0) return;
  auto inside = [&](int ii, int jj, int kk) { const double m = mask(ii, jj, kk); return m == 0.0 || m == 1.0; };
  const int count = (int)inside(i - 1, j, k) + (int)inside(i + 1, j, k) + (int)inside(i, j - 1, k) + (int)inside(i, j + 1, k) +
                    (int)inside(i, j, k - 1) + (int)inside(i, j, k + 1);
  if (count != 1) return;
  const double div = divu.ok() ? divu(i, j, k) : 0.0;
  const double dux = (1.0 / dx0) * (u(i + 1, j, k) - u(i, j, k));
  const double duy = (1.0 / dx1) * (v(i, j + 1, k) - v(i, j, k));
  const double duz = (1.0 / dx2) * (w(i, j, k + 1) - w(i, j, k));
  if (i < vb.lo[0] && mask(i + 1, j, k) != 2.0) u(i, j, k) = u(i + 1, j, k) + dx0 * (duy + duz - div);
  else if (i > vb.hi[0] && mask(i - 1, j, k) != 2.0) u(i + 1, j, k) = u(i, j, k) - dx0 * (duy + duz - div);
  if (j < vb.lo[1] && mask(i, j + 1, k) != 2.0) v(i, j, k) = v(i, j + 1, k) + dx1 * (dux + duz - div);
  else if (j > vb.hi[1] && mask(i, j - 1, k) != 2.0) v(i, j + 1, k) = v(i, j, k) - dx1 * (dux + duz - div);
  if (k < vb.lo[2] && mask(i, j, k + 1) != 2.0) w(i, j, k) = w(i, j, k + 1) + dx2 * (dux + duy - div);
  else if (k > vb.hi[2] && mask(i, j, k - 1) != 2.0) w(i, j, k + 1) = w(i, j, k) - dx2 * (dux + duy - div);
}
int umac_divfix(const Bx& vb, C4 mask, V4 u, V4 v, V4 w, C4 divu, const double dx[3], cudaStream_t s) {
  const Bx g1 = grow(vb, 1);
  IX_LAUNCH(umac_divfix_kernel, dim3(cdiv(g1.nx(), TX), cdiv(g1.ny(), TY), g1.nz()), dim3(TX, TY, 1), 0, s, g1, vb, mask, u, v, w, divu,
            dx[0], dx[1], dx[2]);
  return check_launch("umac_divfix");
}

int average_down_nodal(const Bx& cnbx, V4 crse, C4 fine, int ncomp, cudaStream_t s) {
  if (!cnbx.ok()) return IAMRX_OK;
  IX_LAUNCH(node_inject_kernel, grid_for(cnbx, cnbx.nz() * ncomp), dim3(TX, TY, 1), 0, s, cnbx, crse, fine);
  return check_launch("average_down_nodal");
}
int cell_cons_interp(const Bx& fbx, V4 fine, C4 crse, int ncomp, cudaStream_t s) {
  if (!fbx.ok()) return IAMRX_OK;
  IX_LAUNCH(cell_cons_interp_kernel, grid_for(fbx, fbx.nz() * ncomp), dim3(TX, TY, 1), 0, s, fbx, fine, crse);
  return check_launch("cell_cons_interp");
}
int pc_interp(const Bx& fbx, V4 fine, C4 crse, int ncomp, cudaStream_t s) {
  if (!fbx.ok()) return IAMRX_OK;
  IX_LAUNCH(pc_interp_kernel, dim3(cdiv(fbx.nx(), TX), cdiv(fbx.ny(), TY), fbx.nz() * ncomp), dim3(TX, TY, 1), 0, s, fbx, fine, crse, fbx.nz());
  return check_launch("pc_interp");
}
int sync_mask(const Bx& nbx, V4 mask, C4 covered, double maxcount, cudaStream_t s) {
  IX_LAUNCH(sync_mask_kernel, grid_for(nbx, nbx.nz()), dim3(TX, TY, 1), 0, s, nbx, mask, covered, maxcount);
  return check_launch("sync_mask");
}
int sync_fine_add(const Bx& R, V4 acc, C4 fine, const Bx& fnb, int dir, double mult, const Bx& cnd, const int per[3], cudaStream_t s) {
  if (!R.ok()) return IAMRX_OK;
  IX_LAUNCH(sync_fine_add_kernel, grid_for(R, R.nz()), dim3(TX, TY, 1), 0, s, R, acc, fine, fnb, dir, mult, cnd, per[0], per[1], per[2]);
  return check_launch("sync_fine_add");
}
int sync_gather(const Bx& nbx, V4 reg, C4 onb, C4 acc, const Bx& cnd, const int plen[3], cudaStream_t s) {
  IX_LAUNCH(sync_gather_kernel, grid_for(nbx, nbx.nz()), dim3(TX, TY, 1), 0, s, nbx, reg, onb, acc, cnd, plen[0], plen[1], plen[2]);
  return check_launch("sync_gather");
}
int cf_bndry_interp(const Bx& R, int d, V4 fine, C4 crse, C4 mask, int ncomp, cudaStream_t s) {
  if (!R.ok()) return IAMRX_OK;
  IX_LAUNCH(cf_bndry_kernel, grid_for(R, R.nz() * ncomp), dim3(TX, TY, 1), 0, s, R, fine, crse, mask, d);
  return check_launch("cf_bndry_interp");
}
int node_bilinear_interp(const Bx& fnbx, V4 fine, C4 crse, int ncomp, cudaStream_t s) {
  if (!fnbx.ok()) return IAMRX_OK;
  IX_LAUNCH(node_bilinear_interp_kernel, grid_for(fnbx, fnbx.nz() * ncomp), dim3(TX, TY, 1), 0, s, fnbx, fine, crse);
  return check_launch("node_bilinear_interp");
}
int face_linear_interp(const Bx& ffbx, int dir, V4 fine, C4 crse, int ncomp, cudaStream_t s) {
  if (!ffbx.ok()) return IAMRX_OK;
  IX_LAUNCH(face_linear_interp_kernel, grid_for(ffbx, ffbx.nz() * ncomp), dim3(TX, TY, 1), 0, s, ffbx, dir, fine, crse);
  return check_launch("face_linear_interp");
}
int fr_crse_add(const Bx& R, V4 reg, C4 flux, int d, int fc_off, double scale, int ncomp, cudaStream_t s) {
  if (!R.ok()) return IAMRX_OK;
  IX_LAUNCH(fr_crse_add_kernel, grid_for(R, R.nz() * ncomp), dim3(TX, TY, 1), 0, s, R, reg, flux, d, fc_off, scale);
  return check_launch("fr_crse_add");
}
int fr_fine_add(const Bx& R, V4 reg, C4 flux, int d, int fc_off, double scale, int ncomp, cudaStream_t s) {
  if (!R.ok()) return IAMRX_OK;
  IX_LAUNCH(fr_fine_add_kernel, grid_for(R, R.nz() * ncomp), dim3(TX, TY, 1), 0, s, R, reg, flux, d, fc_off, scale);
  return check_launch("fr_fine_add");
}
int fr_reflux(const Bx& R, V4 state, C4 reg, double scale, int ncomp, cudaStream_t s) {
  if (!R.ok()) return IAMRX_OK;
  IX_LAUNCH(fr_reflux_kernel, grid_for(R, R.nz() * ncomp), dim3(TX, TY, 1), 0, s, R, state, reg, scale);
  return check_launch("fr_reflux");
}

}  // namespace k

// ---------------------------------------------------------------------------------------------------------------------
// Flux register object: coarse level + fine level (ratio 2), every box on this rank (single-rank coupling; see header).
// Interface patches = for every fine box, direction and side, the coarsened one-cell slab just outside it, minus what other
// fine boxes cover (those faces are fine-fine), shifted periodically into the coarse box that owns it.
// ---------------------------------------------------------------------------------------------------------------------
struct FluxReg {
  Level* crse = nullptr; Level* fine = nullptr;
  int ncomp = 0;
  MF reg;                      // on the coarse level, 0 ghost; nonzero only on interface cells
  struct Patch { int cbox; int fbox; int d; int side; Bx R; int shift[3]; };   // R in the index space of coarse box cbox; fine = R - shift
  std::vector<Patch> patches;       // cbox = LOCAL index of a coarse box of this rank; fbox = index into fine->boxes
  // several ranks: the interface regions of ALL coarse boxes (FineAdd goes through one replicated array over the coarse domain:
  // every rank adds the contributions of its own fine boxes, an all-reduce sums them, each rank keeps what lies in its coarse boxes)
  std::vector<Patch> all_patches;   // cbox unused
  std::vector<int> fine_local;      // index into fine->boxes -> local index on this rank, or -1
};

static void subtract_boxes(std::vector<Bx>& pieces, const Bx& cut) {
  std::vector<Bx> out;
  for (const Bx& a : pieces) {
    const Bx is = intersect(a, cut);
    if (!is.ok()) { out.push_back(a); continue; }
    Bx rest = a;
    for (int d = 2; d >= 0; --d) {
      if (rest.lo[d] < is.lo[d]) { Bx p = rest; p.hi[d] = is.lo[d] - 1; out.push_back(p); rest.lo[d] = is.lo[d]; }
      if (rest.hi[d] > is.hi[d]) { Bx p = rest; p.lo[d] = is.hi[d] + 1; out.push_back(p); rest.hi[d] = is.hi[d]; }
    }
  }
  pieces.swap(out);
}

int fluxreg_build(FluxReg& F, Level* crse, Level* fine, int ncomp) {
  F.crse = crse; F.fine = fine; F.ncomp = ncomp;
  F.fine_local.assign(fine->boxes.size(), -1);
  for (int il = 0; il < fine->nlocal(); ++il) F.fine_local[fine->local[il]] = il;
  const bool distributed = comm().nranks > 1 && !crse->replicated;
  F.reg.define(crse, IX_CELL, ncomp, 0);
  std::vector<Bx> cf;   // coarsened fine boxes
  for (const Bx& fb : fine->boxes) {
    Bx c;
    for (int d = 0; d < 3; ++d) {
      if ((fb.lo[d] & 1) || !((fb.hi[d] + 1) % 2 == 0)) { set_error("flux register: fine boxes must be coarsenable by 2"); return IAMRX_ERR_ARG; }
      c.lo[d] = k::coarsen2(fb.lo[d]); c.hi[d] = k::coarsen2(fb.hi[d]);
    }
    cf.push_back(c);
  }
  int plen[3];
  for (int d = 0; d < 3; ++d) plen[d] = crse->domain.hi[d] - crse->domain.lo[d] + 1;
  for (size_t fbi = 0; fbi < cf.size(); ++fbi)
    for (int d = 0; d < 3; ++d)
      for (int side = -1; side <= 1; side += 2) {
        Bx slab = cf[fbi];
        slab.lo[d] = slab.hi[d] = side < 0 ? cf[fbi].lo[d] - 1 : cf[fbi].hi[d] + 1;
        // periodic images of the slab that fall inside the domain; a slab outside a non-periodic side is a physical boundary
        for (int sz = -1; sz <= 1; ++sz) for (int sy = -1; sy <= 1; ++sy) for (int sx = -1; sx <= 1; ++sx) {
          const int sh[3] = {sx * plen[0], sy * plen[1], sz * plen[2]};
          bool ok = true;
          for (int q = 0; q < 3; ++q) if (sh[q] != 0 && !crse->geom.periodic[q]) ok = false;
          if (!ok) continue;
          Bx img = slab;
          for (int q = 0; q < 3; ++q) { img.lo[q] += sh[q]; img.hi[q] += sh[q]; }
          img = intersect(img, crse->domain);
          if (!img.ok()) continue;
          std::vector<Bx> pieces{img};
          // cells covered by (an image of) any fine box are not coarse-fine interface cells
          for (const Bx& other : cf)
            for (int tz = -1; tz <= 1; ++tz) for (int ty = -1; ty <= 1; ++ty) for (int tx = -1; tx <= 1; ++tx) {
              const int th[3] = {tx * plen[0], ty * plen[1], tz * plen[2]};
              bool ok2 = true;
              for (int q = 0; q < 3; ++q) if (th[q] != 0 && !crse->geom.periodic[q]) ok2 = false;
              if (!ok2) continue;
              Bx o = other;
              for (int q = 0; q < 3; ++q) { o.lo[q] += th[q]; o.hi[q] += th[q]; }
              subtract_boxes(pieces, o);
            }
          for (const Bx& pc : pieces) {
            for (int cb = 0; cb < crse->nlocal(); ++cb) {
              const Bx is = intersect(pc, crse->lbox(cb));
              if (!is.ok()) continue;
              FluxReg::Patch P{cb, (int)fbi, d, side, is, {sh[0], sh[1], sh[2]}};
              F.patches.push_back(P);
            }
            if (distributed && F.fine_local[fbi] >= 0) {   // the coarse boxes tile the domain: the whole piece belongs to some rank
              FluxReg::Patch P{-1, (int)fbi, d, side, pc, {sh[0], sh[1], sh[2]}};
              F.all_patches.push_back(P);
            }
          }
        }
      }
  return IAMRX_OK;
}

}  // namespace ix

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
using namespace ix;
namespace ix { Level* level_of(iamrx_level_t h); }
struct iamrx_fluxreg_s { FluxReg F; };

// SyncRegister (SyncRegister.cpp): held densely on the coarse level's nodes -- `reg` is zero off the register's node set B (the
// six boundary planes of every coarsened fine grid: the FabSets bndry[face]), `onb` marks B, `mask` is InitRHS's interior mask.
struct iamrx_syncreg_s {
  Level* crse = nullptr; Level* fine = nullptr;
  MF reg, onb, mask;
  std::vector<Bx> cnb;   // coarsened fine node boxes, one per fine box of the level
};

#define IX_TRY(call) do { int rc_ = (call); if (rc_ != IAMRX_OK) return rc_; } while (0)

extern "C" {

int iamrx_average_down_box(const iamrx_box* cbx, iamrx_fab* crse, const iamrx_fab* fine, int ncomp, int ixtype, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(cbx && crse && fine && ncomp >= 1 && ixtype >= 0 && ixtype <= 4, "average_down arguments");
  const Bx b = ixbox(mkbx(*cbx), ixtype);
  cudaStream_t s = (cudaStream_t)stream;
  if (ixtype == IX_CELL) return k::cc_restrict(b, view(crse), cview(fine), ncomp, s);
  if (ixtype == IX_NODE) return k::average_down_nodal(b, view(crse), cview(fine), ncomp, s);
  return k::face_restrict(b, ixtype - 1, view(crse), cview(fine), ncomp, s);
}

int iamrx_interp_box(int kind, const iamrx_box* fbx, iamrx_fab* fine, const iamrx_fab* crse, int ncomp, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(fbx && fine && crse && ncomp >= 1, "interp arguments");
  const Bx b = mkbx(*fbx);
  cudaStream_t s = (cudaStream_t)stream;
  // the coarse fab must cover the coarsened fine region grown by the stencil (1 cell for the conservative slopes)
  Bx need;
  for (int d = 0; d < 3; ++d) { need.lo[d] = k::coarsen2(b.lo[d]); need.hi[d] = k::coarsen2(b.hi[d]); }
  switch (kind) {
    case IAMRX_INTERP_CELL_CONS: {
      const Bx g = grow(need, 1);
      for (int d = 0; d < 3; ++d) IX_ARG(crse->lo[d] <= g.lo[d] && crse->hi[d] >= g.hi[d], "cell_cons_interp: the coarse fab needs one filled ghost cell around the coarsened fine box");
      return k::cell_cons_interp(b, view(fine), cview(crse), ncomp, s);
    }
    case IAMRX_INTERP_NODE_BILINEAR: {
      const Bx nb = ixbox(b, IX_NODE);
      for (int d = 0; d < 3; ++d) IX_ARG(crse->lo[d] <= k::coarsen2(nb.lo[d]) && crse->hi[d] >= k::coarsen2(nb.hi[d] + 1), "node_bilinear_interp: the coarse fab must cover the coarsened node box");
      return k::node_bilinear_interp(nb, view(fine), cview(crse), ncomp, s);
    }
    case IAMRX_INTERP_FACE_LINEAR_X: case IAMRX_INTERP_FACE_LINEAR_Y: case IAMRX_INTERP_FACE_LINEAR_Z: {
      const int dir = kind - IAMRX_INTERP_FACE_LINEAR_X;
      const Bx fb = ixbox(b, IX_XFACE + dir);
      for (int d = 0; d < 3; ++d) IX_ARG(crse->lo[d] <= k::coarsen2(fb.lo[d]) && crse->hi[d] >= k::coarsen2(fb.hi[d]) + (d == dir ? 1 : 0), "face_linear_interp: the coarse fab must cover the coarsened faces");
      return k::face_linear_interp(fb, dir, view(fine), cview(crse), ncomp, s);
    }
    default: IX_ARG(false, "unknown interpolater");
  }
}

int iamrx_fluxreg_create(iamrx_level_t crse, iamrx_level_t fine, int ncomp, iamrx_fluxreg_t* out) {
  IX_NEED_DEVICE();
  IX_ARG(crse && fine && out && ncomp >= 1, "null argument");
  Level* C = level_of(crse); Level* Fn = level_of(fine);
  for (int d = 0; d < 3; ++d) {
    IX_ARG(Fn->domain.lo[d] == 2 * C->domain.lo[d] && Fn->domain.hi[d] == 2 * C->domain.hi[d] + 1, "the fine level's domain must be the coarse domain refined by 2");
  }
  auto* h = new iamrx_fluxreg_s();
  const int rc = fluxreg_build(h->F, C, Fn, ncomp);
  if (rc) { delete h; return rc; }
  const int rc2 = mf_setval(h->F.reg, 0.0, 0, ncomp, 0, nullptr);
  if (rc2) { delete h; return rc2; }
  *out = h;
  return IAMRX_OK;
}
int iamrx_fluxreg_destroy(iamrx_fluxreg_t r) { delete r; return IAMRX_OK; }
int iamrx_fluxreg_num_patches(iamrx_fluxreg_t r) { return r ? (int)r->F.patches.size() : IAMRX_ERR_ARG; }
int iamrx_fluxreg_reset(iamrx_fluxreg_t r, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(r, "null argument");
  return mf_setval(r->F.reg, 0.0, 0, r->F.ncomp, 0, (cudaStream_t)stream);
}
// fluxes: per coarse local box, per direction (fx[il], fy[il], fz[il]); area-weighted; vol = coarse cell volume ("dx := volume")
int iamrx_fluxreg_crse_add(iamrx_fluxreg_t r, const iamrx_fab* fx, const iamrx_fab* fy, const iamrx_fab* fz, double dt, double vol_crse,
                           void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(r && fx && fy && fz && vol_crse > 0.0, "fluxreg_crse_add arguments");
  const iamrx_fab* f[3] = {fx, fy, fz};
  FluxReg& F = r->F;
  for (const FluxReg::Patch& P : F.patches) {
    // the cell lies below the fine region (side < 0): the shared face is its HIGH face and the coarse flux leaves it (-);
    // above (side > 0): the shared face is its LOW face and the coarse flux enters it (+).  The register holds minus that.
    const int fc_off = P.side < 0 ? 1 : 0;
    const double sgn = P.side < 0 ? 1.0 : -1.0;
    IX_TRY(k::fr_crse_add(P.R, F.reg.v(P.cbox), cview(&f[P.d][P.cbox]), P.d, fc_off, sgn * dt / vol_crse, F.ncomp, (cudaStream_t)stream));
  }
  return IAMRX_OK;
}
// fine fluxes: per FINE local box and direction
int iamrx_fluxreg_fine_add(iamrx_fluxreg_t r, const iamrx_fab* fx, const iamrx_fab* fy, const iamrx_fab* fz, double dt, double vol_crse,
                           void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(r && fx && fy && fz && vol_crse > 0.0, "fluxreg_fine_add arguments");
  const iamrx_fab* f[3] = {fx, fy, fz};
  FluxReg& F = r->F;
  if (comm().nranks > 1 && !F.crse->replicated) {
    // fine and coarse boxes of an interface may live on different ranks
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<Bx> one{F.crse->domain};
    std::vector<int> own{comm().rank};
    std::unique_ptr<Level> RL = make_level(F.crse->geom, one, own);
    RL->replicated = true;
    MF acc(RL.get(), IX_CELL, F.ncomp, 0);
    IX_TRY(mf_setval(acc, 0.0, 0, F.ncomp, 0, s));
    for (const FluxReg::Patch& P : F.all_patches) {
      const int fc_off = P.side < 0 ? 1 : 0;
      const double sgn = P.side < 0 ? -1.0 : 1.0;
      iamrx_fab ff = f[P.d][F.fine_local[P.fbox]];
      for (int q = 0; q < 3; ++q) { ff.lo[q] += 2 * P.shift[q]; ff.hi[q] += 2 * P.shift[q]; }
      IX_TRY(k::fr_fine_add(P.R, acc.v(0), cview(&ff), P.d, fc_off, sgn * dt / vol_crse, F.ncomp, s));
    }
    IX_TRY(comm_allreduce(acc.fabs[0].p, (int)(acc.fabs[0].nstride * F.ncomp), 0, s));
    for (int il = 0; il < F.reg.n(); ++il)
      IX_TRY(k::lincomb(F.reg.vbox(il), F.reg.v(il), 1.0, F.reg.c(il), 1.0, acc.c(0), F.ncomp, s));
    return IAMRX_OK;
  }
  for (const FluxReg::Patch& P : F.patches) {
    const int fc_off = P.side < 0 ? 1 : 0;
    const double sgn = P.side < 0 ? -1.0 : 1.0;
    // the fine box sits at (R - shift) in unshifted coarse index space: hand the kernel a view of the fine flux shifted by 2*shift
    iamrx_fab ff = f[P.d][P.fbox];
    for (int q = 0; q < 3; ++q) { ff.lo[q] += 2 * P.shift[q]; ff.hi[q] += 2 * P.shift[q]; }
    IX_TRY(k::fr_fine_add(P.R, F.reg.v(P.cbox), cview(&ff), P.d, fc_off, sgn * dt / vol_crse, F.ncomp, (cudaStream_t)stream));
  }
  return IAMRX_OK;
}
// state(il, scomp..) += scale * register on the interface cells (NS.cpp:1794-1795 advflux_reg->Reflux)
int iamrx_fluxreg_reflux(iamrx_fluxreg_t r, iamrx_fab* state, int scomp, double scale, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(r && state && scomp >= 0, "fluxreg_reflux arguments");
  FluxReg& F = r->F;
  // a cell can border the fine grids through several faces (several patches), but the register value must be added once:
  // reflux over whole coarse boxes (the register is zero away from the interface)
  for (int il = 0; il < F.crse->nlocal(); ++il)
    IX_TRY(k::fr_reflux(F.crse->lbox(il), view(&state[il], scomp), F.reg.c(il), scale, F.ncomp, (cudaStream_t)stream));
  return IAMRX_OK;
}
int iamrx_fluxreg_field(iamrx_fluxreg_t r, int ilocal, iamrx_fab* out) {
  IX_ARG(r && out && ilocal >= 0 && ilocal < r->F.reg.n(), "fluxreg_field arguments");
  *out = r->F.reg.fabs[ilocal];
  return IAMRX_OK;
}

// amrex::FillPatchTwoLevels for cell-centred data (what AmrLevel::FillPatch does on a level that does not cover the domain;
// NSB.cpp:4399,4435 through FillPatchIterator with the cell_cons_interp interpolater of NS_setup.cpp:211): see iamrx.h.
// The time-interpolated coarse data is gathered into ONE replicated box covering the coarse domain (ghost cells: periodic images
// and the physical boundary fill), so every rank interpolates what its fine boxes need without a second parallel copy -- at 180 GB
// per GPU a replicated coarse level is cheap next to the fine one it serves.
int iamrx_fillpatch_two_levels(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* fine, const iamrx_fab* crse_old,
                               const iamrx_fab* crse_new, double t_old, double t_new, double time, int ncomp, int ngrow,
                               const iamrx_bcrec* bcrec, const double* bcvals, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(fine_lev && crse_lev && fine && crse_new && ncomp >= 1 && ncomp <= 8 && ngrow >= 1, "fillpatch_two_levels arguments");
  Level* FL = level_of(fine_lev);
  Level* CL = level_of(crse_lev);
  cudaStream_t s = (cudaStream_t)stream;
  for (int d = 0; d < 3; ++d) {
    IX_ARG(FL->geom.domain.lo[d] == 2 * CL->geom.domain.lo[d] && FL->geom.domain.hi[d] == 2 * CL->geom.domain.hi[d] + 1,
           "the fine level's domain must be the coarse one refined by 2");
    IX_ARG(FL->geom.periodic[d] == CL->geom.periodic[d], "periodicity differs between the levels");
  }
  double w_new = 1.0;
  if (crse_old && t_new != t_old) w_new = (time - t_old) / (t_new - t_old);
  IX_ARG(w_new >= -1.0e-12 && w_new <= 1.0 + 1.0e-12, "time outside [t_old, t_new]");
  // 1. coarse data at `time` on the coarse layout, gathered into a replicated box with enough ghost cells for the stencil
  const int ngc = (ngrow + 1) / 2 + 1;
  MF cn; cn.alias(CL, IX_CELL, ncomp, 0, const_cast<iamrx_fab*>(crse_new));
  MF ct(CL, IX_CELL, ncomp, 0);
  if (crse_old && w_new != 1.0) {
    MF co; co.alias(CL, IX_CELL, ncomp, 0, const_cast<iamrx_fab*>(crse_old));
    IX_TRY(mf_lincomb(ct, 0, 1.0 - w_new, co, 0, w_new, cn, 0, ncomp, 0, s));
  } else {
    IX_TRY(mf_copy(ct, cn, 0, 0, ncomp, 0, s));
  }
  std::vector<Bx> one{mkbx(CL->geom.domain)};
  std::vector<int> own{comm().rank};
  std::unique_ptr<Level> RL = make_level(CL->geom, one, own);
  RL->replicated = true;
  MF cr(RL.get(), IX_CELL, ncomp, ngc);
  if (CL->replicated || (CL->boxes.size() == 1 && CL->nlocal() == 1)) IX_TRY(mf_copy(cr, ct, 0, 0, ncomp, 0, s));
  else IX_TRY(mf_gather_replicate(cr, ct, ncomp, s));
  IX_TRY(mf_fill_boundary(cr, 0, ncomp, ngc, s));
  k::PhysBC bc{};
  bool walls = false;
  for (int d = 0; d < 3; ++d) if (!CL->geom.periodic[d]) walls = true;
  if (walls) {
    IX_ARG(bcrec != nullptr, "a non-periodic domain needs the BCRec of every component");
    for (int n = 0; n < ncomp; ++n)
      for (int d = 0; d < 3; ++d) { bc.lo[n][d] = bcrec[n].lo[d]; bc.hi[n][d] = bcrec[n].hi[d]; }
    if (bcvals) for (int f = 0; f < 6; ++f) for (int n = 0; n < ncomp; ++n) bc.val[f][n] = bcvals[f * ncomp + n];
    IX_TRY(mf_fill_physbc(cr, 0, ncomp, ngc, bc, s));
  }
  // 2. every ghost cell of every local fine box that lies inside the domain (periodic directions: anywhere) <- interpolation
  MF fm; fm.alias(FL, IX_CELL, ncomp, ngrow, fine);
  const Bx fdom = mkbx(FL->geom.domain);
  for (int il = 0; il < fm.n(); ++il) {
    const Bx vb = fm.vbox(il);
    std::vector<Bx> shell{fm.gbox(il, ngrow)};
    subtract_boxes(shell, vb);
    for (Bx p : shell) {
      for (int d = 0; d < 3; ++d)
        if (!FL->geom.periodic[d]) { p.lo[d] = std::max(p.lo[d], fdom.lo[d]); p.hi[d] = std::min(p.hi[d], fdom.hi[d]); }
      if (!p.ok()) continue;
      IX_TRY(k::cell_cons_interp(p, fm.v(il), cr.c(0), ncomp, s));
    }
  }
  // 3. fine data where fine neighbours (or their periodic images) exist, 4. the physical boundary of the fine level
  IX_TRY(mf_fill_boundary(fm, 0, ncomp, ngrow, s));
  if (walls) IX_TRY(mf_fill_physbc(fm, 0, ncomp, ngrow, bc, s));
  return IAMRX_OK;
}


// The coarse data of a two-level transfer as ONE replicated box over the coarse domain with ngc ghost layers: periodic images and,
// on a non-periodic domain, the physical boundary fill (bc: BCRec + values of every component).
static int replicated_coarse(Level* CL, const iamrx_fab* crse, int scomp, int ncomp, int ixtype, int ngc, const k::PhysBC* bc, cudaStream_t s,
                             std::unique_ptr<Level>& RL, MF& cr) {
  std::vector<Bx> one{mkbx(CL->geom.domain)};
  std::vector<int> own{comm().rank};
  RL = make_level(CL->geom, one, own);
  RL->replicated = true;
  MF cn; cn.alias(CL, ixtype, scomp + ncomp, 0, const_cast<iamrx_fab*>(crse));
  MF ct(CL, ixtype, ncomp, 0);
  IX_TRY(mf_copy(ct, cn, scomp, 0, ncomp, 0, s));
  cr.define(RL.get(), ixtype, ncomp, ngc);
  if (CL->replicated || (CL->boxes.size() == 1 && CL->nlocal() == 1)) IX_TRY(mf_copy(cr, ct, 0, 0, ncomp, 0, s));
  else IX_TRY(mf_gather_replicate(cr, ct, ncomp, s));
  if (ngc > 0) {
    IX_TRY(mf_fill_boundary(cr, 0, ncomp, ngc, s));
    if (bc) IX_TRY(mf_fill_physbc(cr, 0, ncomp, ngc, *bc, s));
  }
  return IAMRX_OK;
}

// amrex::average_down between two levels (NavierStokesBase::avgDown, NSB.cpp:3913-3937 / average_down :3939-3990; avgDown of the
// state and the pressure in NS.cpp:1840-1933): the fine data averaged onto the coarsened fine layout where they live, gathered into
// one replicated coarse-domain box, and copied from there into every local coarse box where a fine box covers it.  Cells: mean of
// the 8 children; nodes: injection (Press_Type, avgDown of the pressure); faces: mean of the 4 fine faces.
int iamrx_average_down(iamrx_level_t fine_lev, iamrx_level_t crse_lev, const iamrx_fab* fine, iamrx_fab* crse, int scomp, int ncomp,
                       int ixtype, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(fine_lev && crse_lev && fine && crse && ncomp >= 1 && scomp >= 0 && ixtype >= 0 && ixtype <= 4, "average_down arguments");
  Level* FL = level_of(fine_lev);
  Level* CL = level_of(crse_lev);
  cudaStream_t s = (cudaStream_t)stream;
  for (int d = 0; d < 3; ++d)
    IX_ARG(FL->geom.domain.lo[d] == 2 * CL->geom.domain.lo[d] && FL->geom.domain.hi[d] == 2 * CL->geom.domain.hi[d] + 1,
           "the fine level's domain must be the coarse one refined by 2");
  // 1. coarsened fine layout (same owners): average where the fine data are
  std::vector<Bx> cfb;
  for (const Bx& b : FL->boxes) {
    Bx c;
    for (int d = 0; d < 3; ++d) {
      IX_ARG((b.lo[d] & 1) == 0 && (b.hi[d] & 1) == 1, "fine boxes must be coarsenable by 2");
      c.lo[d] = b.lo[d] / 2; c.hi[d] = (b.hi[d] - 1) / 2;   // (floor division for negative indices is not needed: lo even, hi odd)
      if (b.lo[d] < 0) c.lo[d] = -((-b.lo[d] + 1) / 2);
      if (b.hi[d] < 0) c.hi[d] = -((-b.hi[d] + 1) / 2);
    }
    cfb.push_back(c);
  }
  std::unique_ptr<Level> CFL = make_level(CL->geom, cfb, FL->owner);
  MF fm; fm.alias(FL, ixtype, scomp + ncomp, 0, const_cast<iamrx_fab*>(fine));
  MF cf(CFL.get(), ixtype, ncomp, 0);
  for (int il = 0; il < cf.n(); ++il) {
    const Bx b = cf.vbox(il);
    if (ixtype == IX_CELL) IX_TRY(k::cc_restrict(b, cf.v(il), fm.c(il, scomp), ncomp, s));
    else if (ixtype == IX_NODE) IX_TRY(k::average_down_nodal(b, cf.v(il), fm.c(il, scomp), ncomp, s));
    else IX_TRY(k::face_restrict(b, ixtype - 1, cf.v(il), fm.c(il, scomp), ncomp, s));
  }
  // 2. one replicated box over the coarse domain
  std::vector<Bx> one{mkbx(CL->geom.domain)};
  std::vector<int> own{comm().rank};
  std::unique_ptr<Level> RL = make_level(CL->geom, one, own);
  RL->replicated = true;
  MF cr(RL.get(), ixtype, ncomp, 0);
  if (comm().nranks <= 1 && cf.n() == 0) return IAMRX_OK;
  IX_TRY(mf_gather_replicate(cr, cf, ncomp, s));
  // 3. into the coarse boxes, only where a fine box covers them
  MF cm; cm.alias(CL, ixtype, scomp + ncomp, 0, crse);
  for (int il = 0; il < cm.n(); ++il) {
    const Bx vb = cm.vbox(il);
    for (const Bx& c : cfb) {
      Bx r = ixbox(c, ixtype);
      for (int d = 0; d < 3; ++d) { r.lo[d] = std::max(r.lo[d], vb.lo[d]); r.hi[d] = std::min(r.hi[d], vb.hi[d]); }
      if (r.ok()) IX_TRY(k::copy(r, cm.v(il, scomp), cr.c(0), ncomp, s));
    }
  }
  return IAMRX_OK;
}

// NavierStokesBase::create_umac_grown on a level > 0 (NSB.cpp:1108-1310): see iamrx.h
int iamrx_create_umac_grown(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* umac_f, iamrx_fab* vmac_f, iamrx_fab* wmac_f,
                            const iamrx_fab* umac_c, const iamrx_fab* vmac_c, const iamrx_fab* wmac_c, const iamrx_fab* divu, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(fine_lev && crse_lev && umac_f && vmac_f && wmac_f && umac_c && vmac_c && wmac_c, "create_umac_grown arguments");
  Level* FL = level_of(fine_lev);
  Level* CL = level_of(crse_lev);
  cudaStream_t s = (cudaStream_t)stream;
  for (int d = 0; d < 3; ++d)
    IX_ARG(FL->geom.domain.lo[d] == 2 * CL->geom.domain.lo[d] && FL->geom.domain.hi[d] == 2 * CL->geom.domain.hi[d] + 1,
           "the fine level's domain must be the coarse one refined by 2");
  const Bx fdom = mkbx(FL->geom.domain);
  iamrx_fab* uf[3] = {umac_f, vmac_f, wmac_f};
  const iamrx_fab* uc[3] = {umac_c, vmac_c, wmac_c};
  // 1. FillPatchTwoLevels of the face velocities with face_linear_interp (piecewise constant in time), one ghost cell: every face
  //    of the grown box from the coarse level, then the box's own faces, then the fine neighbours' (and their periodic images)
  for (int d = 0; d < 3; ++d) {
    std::unique_ptr<Level> RL;
    MF cr;
    IX_TRY(replicated_coarse(CL, uc[d], 0, 1, IX_XFACE + d, 1, nullptr, s, RL, cr));
    MF um; um.alias(FL, IX_XFACE + d, 1, 1, uf[d]);
    MF tmp(FL, IX_XFACE + d, 1, 1);
    for (int il = 0; il < um.n(); ++il) {
      Bx gb = grow(FL->lbox(il), 1);
      for (int q = 0; q < 3; ++q)
        if (!FL->geom.periodic[q]) { gb.lo[q] = std::max(gb.lo[q], fdom.lo[q]); gb.hi[q] = std::min(gb.hi[q], fdom.hi[q]); }
      IX_TRY(k::copy(um.gbox(il, 1), tmp.v(il), um.c(il), 1, s));                       // (cells outside a physical boundary keep the caller's values)
      IX_TRY(k::face_linear_interp(ixbox(gb, IX_XFACE + d), d, tmp.v(il), cr.c(0), 1, s));
      IX_TRY(k::copy(um.vbox(il), tmp.v(il), um.c(il), 1, s));
    }
    IX_TRY(mf_fill_boundary(tmp, 0, 1, 1, s));
    for (int il = 0; il < um.n(); ++il) IX_TRY(k::copy(um.gbox(il, 1), um.v(il), tmp.c(il), 1, s));
  }
  // 2. the coarse-fine mask on the boxes grown by 2 (iMultiFab::BuildMask: interior / covered / not covered / physical boundary)
  MF mask(FL, IX_CELL, 1, 2);
  IX_TRY(mf_setval(mask, 3.0, 0, 1, 2, s));
  for (int il = 0; il < mask.n(); ++il) {
    const Bx vb = FL->lbox(il), g2 = grow(vb, 2);
    Bx in = g2;
    for (int q = 0; q < 3; ++q)
      if (!FL->geom.periodic[q]) { in.lo[q] = std::max(in.lo[q], fdom.lo[q]); in.hi[q] = std::min(in.hi[q], fdom.hi[q]); }
    IX_TRY(k::setval(in, mask.v(il), 1, 2.0, s));
    int sh[3];
    for (sh[2] = -1; sh[2] <= 1; ++sh[2])
      for (sh[1] = -1; sh[1] <= 1; ++sh[1])
        for (sh[0] = -1; sh[0] <= 1; ++sh[0]) {
          bool okshift = true;
          for (int q = 0; q < 3; ++q) if (sh[q] != 0 && !FL->geom.periodic[q]) okshift = false;
          if (!okshift) continue;
          for (const Bx& ob : FL->boxes) {
            Bx r = ob;
            for (int q = 0; q < 3; ++q) { const int n = fdom.hi[q] - fdom.lo[q] + 1; r.lo[q] += sh[q] * n; r.hi[q] += sh[q] * n; }
            for (int q = 0; q < 3; ++q) { r.lo[q] = std::max(r.lo[q], g2.lo[q]); r.hi[q] = std::min(r.hi[q], g2.hi[q]); }
            if (r.ok()) IX_TRY(k::setval(r, mask.v(il), 1, 1.0, s));
          }
        }
    IX_TRY(k::setval(vb, mask.v(il), 1, 0.0, s));
  }
  // 3. the divergence correction of the one-cell halo
  MF U[3];
  for (int d = 0; d < 3; ++d) U[d].alias(FL, IX_XFACE + d, 1, 1, uf[d]);
  MF Dv; if (divu) Dv.alias(FL, IX_CELL, 1, 1, const_cast<iamrx_fab*>(divu));
  for (int il = 0; il < mask.n(); ++il)
    IX_TRY(k::umac_divfix(FL->lbox(il), mask.c(il), U[0].v(il), U[1].v(il), U[2].v(il), divu ? Dv.c(il) : C4{}, FL->geom.dx, s));
  return IAMRX_OK;
}

// MLLinOp::setCoarseFineBC(crse, ratio) (MacProj.cpp:1164-1167): see iamrx.h
int iamrx_set_coarse_fine_bc(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* fine, const iamrx_fab* crse, int ncomp, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(fine_lev && crse_lev && fine && crse && ncomp >= 1 && ncomp <= 8, "set_coarse_fine_bc arguments");
  Level* FL = level_of(fine_lev);
  Level* CL = level_of(crse_lev);
  cudaStream_t s = (cudaStream_t)stream;
  for (int d = 0; d < 3; ++d) {
    IX_ARG(FL->geom.domain.lo[d] == 2 * CL->geom.domain.lo[d] && FL->geom.domain.hi[d] == 2 * CL->geom.domain.hi[d] + 1,
           "the fine level's domain must be the coarse one refined by 2");
    IX_ARG(FL->geom.periodic[d] == CL->geom.periodic[d], "periodicity differs between the levels");
  }
  // 1. the coarse data as one replicated box with two ghost layers (periodic images; nothing beyond physical sides is read)
  std::unique_ptr<Level> RL;
  MF cr;
  IX_TRY(replicated_coarse(CL, crse, 0, ncomp, IX_CELL, 2, nullptr, s, RL, cr));
  // 2. usable coarse cells: inside the domain (periodic images included) and not under the fine level
  MF mask(RL.get(), IX_CELL, 1, 2);
  IX_TRY(mf_setval(mask, 0.0, 0, 1, 2, s));
  IX_TRY(k::setval(mkbx(CL->geom.domain), mask.v(0), 1, 1.0, s));
  for (const Bx& fb : FL->boxes) {
    Bx cb;
    for (int d = 0; d < 3; ++d) { cb.lo[d] = k::coarsen2(fb.lo[d]); cb.hi[d] = k::coarsen2(fb.hi[d]); }
    IX_TRY(k::setval(cb, mask.v(0), 1, 0.0, s));
  }
  IX_TRY(mf_fill_boundary(mask, 0, 1, 2, s));
  // 3. the layer beyond every side of every local fine box
  MF fm; fm.alias(FL, IX_CELL, ncomp, 1, fine);
  const Bx fdom = mkbx(FL->geom.domain);
  for (int il = 0; il < fm.n(); ++il) {
    const Bx vb = fm.vbox(il);
    for (int d = 0; d < 3; ++d)
      for (int side = 0; side < 2; ++side) {
        Bx R = vb;
        R.lo[d] = R.hi[d] = side == 0 ? vb.lo[d] - 1 : vb.hi[d] + 1;
        if (!FL->geom.periodic[d] && (R.lo[d] < fdom.lo[d] || R.hi[d] > fdom.hi[d])) continue;
        IX_TRY(k::cf_bndry_interp(R, d, fm.v(il), cr.c(0), mask.c(0), ncomp, s));
      }
  }
  return IAMRX_OK;
}

// NavierStokesBase::SyncInterp (NSB.cpp:3071-3255): see iamrx.h
int iamrx_sync_interp(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* fine_sync, int dest_comp, const iamrx_fab* crse_sync,
                      int src_comp, int ncomp, int increment, double dt_clev, int which_interp, const iamrx_bcrec* bcrec, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(fine_lev && crse_lev && fine_sync && crse_sync && ncomp >= 1 && ncomp <= 8 && dest_comp >= 0 && src_comp >= 0, "sync_interp arguments");
  IX_ARG(which_interp == IAMRX_SYNC_PC || which_interp == IAMRX_SYNC_CELL_CONS, "sync_interp: PC_T and CellCons_T are implemented");
  Level* FL = level_of(fine_lev);
  Level* CL = level_of(crse_lev);
  cudaStream_t s = (cudaStream_t)stream;
  bool walls = false;
  for (int d = 0; d < 3; ++d) {
    IX_ARG(FL->geom.domain.lo[d] == 2 * CL->geom.domain.lo[d] && FL->geom.domain.hi[d] == 2 * CL->geom.domain.hi[d] + 1,
           "the fine level's domain must be the coarse one refined by 2");
    if (!CL->geom.periodic[d]) walls = true;
  }
  k::PhysBC bc{};   // HomExtDirFill: the BCRec of the original quantity with homogeneous ext_dir values
  if (walls) {
    IX_ARG(bcrec != nullptr, "a non-periodic domain needs the BCRec of every component (bc_orig_qty)");
    for (int n = 0; n < ncomp; ++n) for (int d = 0; d < 3; ++d) { bc.lo[n][d] = bcrec[n].lo[d]; bc.hi[n][d] = bcrec[n].hi[d]; }
  }
  std::unique_ptr<Level> RL;
  MF cr;
  IX_TRY(replicated_coarse(CL, crse_sync, src_comp, ncomp, IX_CELL, which_interp == IAMRX_SYNC_PC ? 0 : 1, walls ? &bc : nullptr, s, RL, cr));
  MF fm; fm.alias(FL, IX_CELL, dest_comp + ncomp, 0, fine_sync);
  MF tmp(FL, IX_CELL, ncomp, 0);
  for (int il = 0; il < fm.n(); ++il) {
    const Bx vb = fm.vbox(il);
    if (which_interp == IAMRX_SYNC_PC) IX_TRY(k::pc_interp(vb, tmp.v(il), cr.c(0), ncomp, s));
    else IX_TRY(k::cell_cons_interp(vb, tmp.v(il), cr.c(0), ncomp, s));
    if (increment) {   // finedata *= dt_clev; fsync += finedata (:3209-3236)
      IX_TRY(k::scale(vb, tmp.v(il), dt_clev, ncomp, s));
      IX_TRY(k::lincomb(vb, fm.v(il, dest_comp), 1.0, fm.c(il, dest_comp), 1.0, tmp.c(il), ncomp, s));
    } else {
      IX_TRY(k::copy(vb, fm.v(il, dest_comp), tmp.c(il), ncomp, s));
    }
  }
  return IAMRX_OK;
}

// NavierStokesBase::SyncProjInterp (NSB.cpp:3258-3336): P_new += I(phi), P_old += I(phi) with node_bilinear_interp
int iamrx_sync_proj_interp(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* p_new, iamrx_fab* p_old, const iamrx_fab* phi_crse,
                           void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(fine_lev && crse_lev && p_new && p_old && phi_crse, "sync_proj_interp arguments");
  Level* FL = level_of(fine_lev);
  Level* CL = level_of(crse_lev);
  cudaStream_t s = (cudaStream_t)stream;
  for (int d = 0; d < 3; ++d)
    IX_ARG(FL->geom.domain.lo[d] == 2 * CL->geom.domain.lo[d] && FL->geom.domain.hi[d] == 2 * CL->geom.domain.hi[d] + 1,
           "the fine level's domain must be the coarse one refined by 2");
  std::unique_ptr<Level> RL;
  MF cr;
  IX_TRY(replicated_coarse(CL, phi_crse, 0, 1, IX_NODE, 0, nullptr, s, RL, cr));
  MF pn; pn.alias(FL, IX_NODE, 1, 0, p_new);
  MF po; po.alias(FL, IX_NODE, 1, 0, p_old);
  MF tmp(FL, IX_NODE, 1, 0);
  for (int il = 0; il < pn.n(); ++il) {
    const Bx nb = pn.vbox(il);
    IX_TRY(k::node_bilinear_interp(nb, tmp.v(il), cr.c(0), 1, s));
    IX_TRY(k::lincomb(nb, pn.v(il), 1.0, pn.c(il), 1.0, tmp.c(il), 1, s));
    IX_TRY(k::lincomb(nb, po.v(il), 1.0, po.c(il), 1.0, tmp.c(il), 1, s));
  }
  return IAMRX_OK;
}

// SyncRegister::SyncRegister (SyncRegister.cpp:20-47) + the mask of InitRHS (:126-272, which depends on the grids only)
int iamrx_syncreg_create(iamrx_level_t crse_lev, iamrx_level_t fine_lev, double mask_maxcount, iamrx_syncreg_t* out) {
  IX_NEED_DEVICE();
  IX_ARG(crse_lev && fine_lev && out, "syncreg_create arguments");
  Level* CL = level_of(crse_lev); Level* FL = level_of(fine_lev);
  for (int d = 0; d < 3; ++d)
    IX_ARG(FL->domain.lo[d] == 2 * CL->domain.lo[d] && FL->domain.hi[d] == 2 * CL->domain.hi[d] + 1, "the fine level's domain must be the coarse domain refined by 2");
  auto h = std::make_unique<iamrx_syncreg_s>();
  h->crse = CL; h->fine = FL;
  std::vector<Bx> ccb;   // coarsened fine cell boxes
  for (const Bx& fb : FL->boxes) {
    Bx c;
    for (int d = 0; d < 3; ++d) {
      IX_ARG(!(fb.lo[d] & 1) && ((fb.hi[d] + 1) % 2 == 0), "sync register: fine boxes must be coarsenable by 2");
      c.lo[d] = k::coarsen2(fb.lo[d]); c.hi[d] = k::coarsen2(fb.hi[d]);
    }
    ccb.push_back(c);
    h->cnb.push_back(ixbox(c, IX_NODE));
  }
  h->reg.define(CL, IX_NODE, 1, 0); h->onb.define(CL, IX_NODE, 1, 0); h->mask.define(CL, IX_NODE, 1, 0);
  cudaStream_t s = nullptr;
  IX_TRY(mf_setval(h->reg, 0.0, 0, 1, 0, s));
  IX_TRY(mf_setval(h->onb, 0.0, 0, 1, 0, s));
  int plen[3];
  for (int d = 0; d < 3; ++d) plen[d] = CL->domain.hi[d] - CL->domain.lo[d] + 1;
  // B: the boundary planes of the coarsened fine node boxes and their periodic images
  for (const Bx& nb : h->cnb)
    for (int d = 0; d < 3; ++d)
      for (int side = 0; side < 2; ++side) {
        Bx pl = nb;
        pl.lo[d] = pl.hi[d] = side == 0 ? nb.lo[d] : nb.hi[d];
        for (int sz = -1; sz <= 1; ++sz) for (int sy = -1; sy <= 1; ++sy) for (int sx = -1; sx <= 1; ++sx) {
          const int sh[3] = {sx * plen[0], sy * plen[1], sz * plen[2]};
          bool ok = true;
          for (int q = 0; q < 3; ++q) if (sh[q] != 0 && !CL->geom.periodic[q]) ok = false;
          if (!ok) continue;
          Bx img = pl;
          for (int q = 0; q < 3; ++q) { img.lo[q] += sh[q]; img.hi[q] += sh[q]; }
          for (int il = 0; il < h->onb.n(); ++il) {
            const Bx is = intersect(img, h->onb.vbox(il));
            if (is.ok()) IX_TRY(k::setval(is, h->onb.v(il), 1, 1.0, s));
          }
        }
      }
  // covered coarse cells with one ghost layer: periodic images, mirror images beyond non-periodic sides
  MF cov(CL, IX_CELL, 1, 1);
  IX_TRY(mf_setval(cov, 0.0, 0, 1, 1, s));
  for (const Bx& c : ccb)
    for (int il = 0; il < cov.n(); ++il) {
      const Bx is = intersect(c, cov.vbox(il));
      if (is.ok()) IX_TRY(k::setval(is, cov.v(il), 1, 1.0, s));
    }
  IX_TRY(mf_fill_boundary(cov, 0, 1, 1, s));
  bool walls = false;
  for (int d = 0; d < 3; ++d) if (!CL->geom.periodic[d]) walls = true;
  if (walls) {
    k::PhysBC bc{};
    for (int d = 0; d < 3; ++d) { bc.lo[0][d] = IAMRX_BC_REFLECT_EVEN; bc.hi[0][d] = IAMRX_BC_REFLECT_EVEN; }
    IX_TRY(mf_fill_physbc(cov, 0, 1, 1, bc, s));
  }
  for (int il = 0; il < h->mask.n(); ++il) IX_TRY(k::sync_mask(h->mask.vbox(il), h->mask.v(il), cov.c(il), mask_maxcount > 0.0 ? mask_maxcount : 26.5, s));
  IX_CUDA(cudaStreamSynchronize(s));
  *out = h.release();
  return IAMRX_OK;
}
int iamrx_syncreg_destroy(iamrx_syncreg_t r) { delete r; return IAMRX_OK; }

// SyncRegister::CrseInit (SyncRegister.cpp:288-300)
int iamrx_syncreg_crse_init(iamrx_syncreg_t r, const iamrx_fab* sync_resid_crse, double mult, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(r && sync_resid_crse, "syncreg_crse_init arguments");
  cudaStream_t s = (cudaStream_t)stream;
  MF src; src.alias(r->crse, IX_NODE, 1, 0, const_cast<iamrx_fab*>(sync_resid_crse));
  for (int il = 0; il < r->reg.n(); ++il) {
    const Bx nb = r->reg.vbox(il);
    IX_TRY(k::lincomb(nb, r->reg.v(il), mult, src.c(il), 0.0, src.c(il), 1, s));
    IX_TRY(k::mult(nb, r->reg.v(il), r->onb.c(il), 1, 1, s));
  }
  return IAMRX_OK;
}

// SyncRegister::FineAdd (SyncRegister.cpp:351-607)
int iamrx_syncreg_fine_add(iamrx_syncreg_t r, const iamrx_fab* sync_resid_fine, double mult, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(r && sync_resid_fine, "syncreg_fine_add arguments");
  cudaStream_t s = (cudaStream_t)stream;
  Level* CL = r->crse; Level* FL = r->fine;
  // the restricted planes of every local fine box, summed into one replicated array over the coarse node domain
  std::vector<Bx> one{mkbx(CL->geom.domain)};
  std::vector<int> own{comm().rank};
  std::unique_ptr<Level> RL = make_level(CL->geom, one, own);
  RL->replicated = true;
  MF acc(RL.get(), IX_NODE, 1, 0);
  IX_TRY(mf_setval(acc, 0.0, 0, 1, 0, s));
  const Bx cnd = ixbox(CL->domain, IX_NODE);
  MF fm; fm.alias(FL, IX_NODE, 1, 0, const_cast<iamrx_fab*>(sync_resid_fine));
  for (int il = 0; il < fm.n(); ++il) {
    const Bx fnb = fm.vbox(il);
    const Bx nb = r->cnb[FL->local[il]];
    for (int d = 0; d < 3; ++d)
      for (int side = 0; side < 2; ++side) {
        Bx pl = nb;
        pl.lo[d] = pl.hi[d] = side == 0 ? nb.lo[d] : nb.hi[d];
        IX_TRY(k::sync_fine_add(pl, acc.v(0), fm.c(il), fnb, d, mult, cnd, CL->geom.periodic, s));
      }
  }
  if (comm().nranks > 1) {
    const Bx ab = acc.vbox(0);
    IX_TRY(comm_allreduce(acc.fabs[0].p, (int)(acc.fabs[0].kstride * ab.nz()), 0, s));
  }
  int plen[3];
  for (int d = 0; d < 3; ++d) plen[d] = CL->geom.periodic[d] ? CL->domain.hi[d] - CL->domain.lo[d] + 1 : 0;
  for (int il = 0; il < r->reg.n(); ++il)
    IX_TRY(k::sync_gather(r->reg.vbox(il), r->reg.v(il), r->onb.c(il), acc.c(0), cnd, plen, s));
  return IAMRX_OK;
}

// SyncRegister::InitRHS (SyncRegister.cpp:49-285)
int iamrx_syncreg_init_rhs(iamrx_syncreg_t r, iamrx_fab* rhs, const int phys_lo[3], const int phys_hi[3], void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(r && rhs, "syncreg_init_rhs arguments");
  cudaStream_t s = (cudaStream_t)stream;
  Level* CL = r->crse;
  const Bx cnd = ixbox(CL->domain, IX_NODE);
  MF out; out.alias(CL, IX_NODE, 1, 0, rhs);
  for (int il = 0; il < out.n(); ++il) {
    const Bx nb = out.vbox(il);
    IX_TRY(k::copy(nb, out.v(il), r->reg.c(il), 1, s));
    for (int d = 0; d < 3; ++d) {
      if (CL->geom.periodic[d]) continue;
      for (int side = 0; side < 2; ++side) {
        const int code = side == 0 ? (phys_lo ? phys_lo[d] : 0) : (phys_hi ? phys_hi[d] : 0);
        if (code != 2) continue;   // PhysBCType::outflow (the ns.lo_bc / ns.hi_bc code, iamrx_ns_params.lo_bc)
        Bx pl = cnd;
        pl.lo[d] = pl.hi[d] = side == 0 ? cnd.lo[d] : cnd.hi[d];
        const Bx is = intersect(pl, nb);
        if (is.ok()) IX_TRY(k::setval(is, out.v(il), 1, 0.0, s));
      }
    }
    IX_TRY(k::mult(nb, out.v(il), r->mask.c(il), 1, 1, s));
  }
  return IAMRX_OK;
}
int iamrx_syncreg_field(iamrx_syncreg_t r, int which, int ilocal, iamrx_fab* out) {
  IX_ARG(r && out && ilocal >= 0 && ilocal < r->reg.n() && which >= 0 && which <= 2, "syncreg_field arguments");
  *out = (which == 0 ? r->reg : (which == 1 ? r->onb : r->mask)).fabs[ilocal];
  return IAMRX_OK;
}

// AmrLevel::FillCoarsePatch of Press_Type (Projection.cpp:236-239): see iamrx.h
int iamrx_fill_coarse_patch_nodal(iamrx_level_t fine_lev, iamrx_level_t crse_lev, iamrx_fab* fine, const iamrx_fab* crse_old,
                                  const iamrx_fab* crse_new, double t_new_start, double t_new_stop, double time, void* stream) {
  IX_NEED_DEVICE();
  IX_ARG(fine_lev && crse_lev && fine && crse_new, "fill_coarse_patch_nodal arguments");
  Level* FL = level_of(fine_lev);
  Level* CL = level_of(crse_lev);
  cudaStream_t s = (cudaStream_t)stream;
  for (int d = 0; d < 3; ++d)
    IX_ARG(FL->geom.domain.lo[d] == 2 * CL->geom.domain.lo[d] && FL->geom.domain.hi[d] == 2 * CL->geom.domain.hi[d] + 1,
           "the fine level's domain must be the coarse one refined by 2");
  // Press_Type is an Interval quantity (NS_setup.cpp:329-331): StateData hands out the data of the time interval that contains
  // `time` -- no interpolation in time.  New interval [t_new_start, t_new_stop], old interval ending at t_new_start; the new one is
  // tried first, with AMReX's tolerance of 1e-3 of its length.
  const double teps = 1.0e-3 * (t_new_stop - t_new_start);
  const bool use_new = time > t_new_start - teps && time < t_new_stop + teps;
  IX_ARG(use_new || (crse_old != nullptr && time <= t_new_start + teps), "time lies in neither pressure interval");
  MF cn; cn.alias(CL, IX_NODE, 1, 0, const_cast<iamrx_fab*>(use_new ? crse_new : crse_old));
  MF ct(CL, IX_NODE, 1, 0);
  IX_TRY(mf_copy(ct, cn, 0, 0, 1, 0, s));
  std::unique_ptr<Level> RL;
  MF cr;
  IX_TRY(replicated_coarse(CL, ct.fabs.data(), 0, 1, IX_NODE, 0, nullptr, s, RL, cr));
  MF fm; fm.alias(FL, IX_NODE, 1, 0, fine);
  for (int il = 0; il < fm.n(); ++il) IX_TRY(k::node_bilinear_interp(fm.vbox(il), fm.v(il), cr.c(0), 1, s));
  return IAMRX_OK;
}

}  // extern "C"
