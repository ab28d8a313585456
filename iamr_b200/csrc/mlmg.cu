// mlmg.cu -- see mlmg.h.  V-cycle drivers; every arithmetic step is one of the
// kernels in abec.cu / nodal.cu / blas.cu launched over the rank's local boxes.
#include <functional>
#include "mlmg.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace ix {

#define IX_TRY(call) do { int rc_ = (call); if (rc_ != IAMRX_OK) return rc_; } while (0)

std::unique_ptr<Level> make_level(const iamrx_geom& g, const std::vector<Bx>& boxes,
                                  const std::vector<int>& owner) {
  auto L = std::make_unique<Level>();
  L->geom = g;
  L->boxes = boxes;
  L->owner = owner;
  L->domain = mkbx(g.domain);
  const int me = comm().rank;
  for (size_t i = 0; i < boxes.size(); ++i) {
    if (owner[i] == me) L->local.push_back((int)i);
    L->ncells_global += boxes[i].npts();
  }
  for (int d = 0; d < 3; ++d) L->dxinv[d] = 1.0 / g.dx[d];
  return L;
}

// thin: directions that are never coarsened (semi-coarsening: a domain two cells thick in z is a "2-D" problem whose multigrid
// hierarchy coarsens x and y only; the transfer operators take the same mask)
int thin_mask(const Level& f) {
  int m = 0;
  for (int d = 0; d < 3; ++d) if (f.geom.domain.hi[d] - f.geom.domain.lo[d] + 1 <= 2) m |= 1 << d;
  return m == 7 ? 0 : m;
}

std::unique_ptr<Level> coarsen_level(const Level& f, int min_width, int thin) {
  std::vector<Bx> cb;
  for (const Bx& b : f.boxes) {
    Bx c;
    for (int d = 0; d < 3; ++d) {
      if (thin & (1 << d)) { c.lo[d] = b.lo[d]; c.hi[d] = b.hi[d]; continue; }
      const int n = b.hi[d] - b.lo[d] + 1;
      if ((n % 2) != 0 || (b.lo[d] % 2) != 0 || n / 2 < min_width) return nullptr;
      c.lo[d] = b.lo[d] / 2;
      c.hi[d] = c.lo[d] + n / 2 - 1;
    }
    cb.push_back(c);
  }
  iamrx_geom g = f.geom;
  for (int d = 0; d < 3; ++d) {
    // a thin direction keeps its cells AND its spacing (the operator stays consistent for the mode that differs between the two
    // layers).  For the point smoother not to meet a dominant coupling across the layers on coarse levels, such a "2-D" run
    // makes the layers THICK: z extent about half the x extent, so dz >= dx on every level.
    if (thin & (1 << d)) continue;
    const int n = f.geom.domain.hi[d] - f.geom.domain.lo[d] + 1;
    if ((n % 2) != 0 || (f.geom.domain.lo[d] % 2) != 0) return nullptr;
    g.domain.lo[d] = f.geom.domain.lo[d] / 2;
    g.domain.hi[d] = g.domain.lo[d] + n / 2 - 1;
    g.dx[d] = 2.0 * f.geom.dx[d];
  }
  auto L = make_level(g, cb, f.owner);
  L->replicated = f.replicated;
  return L;
}

std::unique_ptr<Level> consolidated_level(const Level& c) {
  static int64_t thresh = -1;   // boxes of at most this many cells are consolidated (IAMRX_MG_CONSOLIDATE=0: never)
  if (thresh < 0) { const char* e = getenv("IAMRX_MG_CONSOLIDATE"); thresh = e ? atoll(e) : 32 * 32 * 32; }
  if (c.replicated || c.boxes.size() < 2 || thresh == 0) return nullptr;
  for (const Bx& b : c.boxes) if (b.npts() > thresh) return nullptr;
  int64_t covered = 0;
  for (const Bx& b : c.boxes) covered += b.npts();
  if (covered != mkbx(c.geom.domain).npts()) return nullptr;   // boxes must tile the domain
  std::vector<Bx> one{mkbx(c.geom.domain)};
  std::vector<int> own{comm().rank};
  auto L = make_level(c.geom, one, own);
  L->replicated = true;
  return L;
}

// Sides of box b of level L that border COARSE cells: the one-cell layer beyond the side is inside the domain (or a periodic
// direction) and not completely covered by the boxes of the level and their periodic images.  Bit 2 d + side.
static int coarse_fine_sides(const Level& L, const Bx& b) {
  int mask = 0;
  const Bx dom = L.domain;
  for (int d = 0; d < 3; ++d)
    for (int side = 0; side < 2; ++side) {
      Bx S = b;
      S.lo[d] = S.hi[d] = side == 0 ? b.lo[d] - 1 : b.hi[d] + 1;
      if (!L.geom.periodic[d] && (S.lo[d] < dom.lo[d] || S.hi[d] > dom.hi[d])) continue;   // a physical side
      int64_t covered = 0;
      const int64_t need = S.npts();
      int sh[3];
      for (sh[2] = -1; sh[2] <= 1 && covered < need; ++sh[2])
        for (sh[1] = -1; sh[1] <= 1 && covered < need; ++sh[1])
          for (sh[0] = -1; sh[0] <= 1 && covered < need; ++sh[0]) {
            bool ok = true;
            for (int q = 0; q < 3; ++q) {
              if (sh[q] != 0 && !L.geom.periodic[q]) ok = false;
              // an image shifted by a whole period can only reach the slab if the slab lies at that end of the domain
              if (sh[q] > 0 && S.hi[q] <= dom.hi[q]) ok = false;
              if (sh[q] < 0 && S.lo[q] >= dom.lo[q]) ok = false;
            }
            if (!ok) continue;
            for (const Bx& o : L.boxes) {
              Bx t = o;
              for (int q = 0; q < 3; ++q) { const int len = dom.hi[q] - dom.lo[q] + 1; t.lo[q] += sh[q] * len; t.hi[q] += sh[q] * len; }
              const Bx x = intersect(S, t);
              if (x.ok()) covered += x.npts();
            }
          }
      if (covered < S.npts()) mask |= 1 << (2 * d + side);
    }
  return mask;
}

static bool all_periodic(const Level& L) {
  return L.geom.periodic[0] && L.geom.periodic[1] && L.geom.periodic[2];
}

// ===========================================================================
// CellMG
// ===========================================================================
CellMG::CellMG(Level* fine, int ncomp, bool tensor, int max_coarsening)
    : ncomp_(ncomp), tensor_(tensor) {
  iamrx_mg_info_default(&info_);
  thin_ = thin_mask(*fine);
  lv_.emplace_back();
  lv_[0].lev = fine;
  Level* cur = fine;
  for (int l = 1; l <= max_coarsening; ++l) {
    auto c = coarsen_level(*cur, 2, thin_);
    if (!c) break;
    lv_.emplace_back();
    if (auto r = consolidated_level(*c)) {   // from here down: one replicated box per rank, no ghost traffic
      lv_.back().xfer_lev = std::move(c);
      lv_.back().lev_owned = std::move(r);
    } else {
      lv_.back().lev_owned = std::move(c);
    }
    lv_.back().lev = lv_.back().lev_owned.get();
    cur = lv_.back().lev;
  }
  // IAMRX_CELL_DEEP=0 / 1 forces the deep-ghost sweeps off / on; by default they are used when boxes live on several ranks (they
  // trade a larger ghost copy and a grown red pass for one exchange instead of two: a loss when the exchange is a local copy --
  // 155.5 vs 149.6 ms per step for 8 boxes on one GPU -- and a gain when it is an NCCL round trip)
  static int deep_env = -2;
  if (deep_env == -2) { const char* e = getenv("IAMRX_CELL_DEEP"); deep_env = e ? (e[0] == '0' ? 0 : 1) : -1; }
  const int deep_on = deep_env >= 0 ? deep_env : (comm().nranks > 1 ? 1 : 0);
  for (auto& L : lv_) {
    for (int d = 0; d < 3; ++d) L.dxinv[d] = L.lev->dxinv[d];
    L.deep = deep_on && !L.lev->replicated && L.lev->level_wrapmask() != 7 && all_periodic(*L.lev);
    // deep levels: 2 m ghost layers of the correction, 2 m - 1 of the right-hand side and of the coefficients let m sweeps run on ONE
    // exchange (red / black passes on boxes that shrink by one layer per pass, recomputing what the neighbour computes); m = 2 =
    // the pre- or post-smoothing of a V-cycle (IAMRX_CELL_DEEP_SWEEPS=1: one sweep per exchange)
    static int dsw = -1;
    if (dsw < 0) { const char* e = getenv("IAMRX_CELL_DEEP_SWEEPS"); dsw = (e && atoi(e) == 1) ? 1 : 2; }
    L.deep_sweeps = L.deep ? dsw : 0;
    L.cor.define(L.lev, IX_CELL, ncomp_, L.deep ? 2 * dsw : 1);
    L.res.define(L.lev, IX_CELL, ncomp_, L.deep ? 2 * dsw - 1 : 0);
    L.rescor.define(L.lev, IX_CELL, ncomp_, 0);
    if (L.xfer_lev) L.xfer.define(L.xfer_lev.get(), IX_CELL, ncomp_, 0);
  }
  // a level that does not tile its domain is a fine AMR level: find the coarse-fine sides of every box on every multigrid level
  if (!fine->replicated && fine->ncells_global != mkbx(fine->geom.domain).npts()) {
    cfmask_.resize(lv_.size());
    for (size_t l = 0; l < lv_.size(); ++l) {
      const Level& L = *lv_[l].lev;
      for (int il = 0; il < L.nlocal(); ++il) {
        cfmask_[l].push_back(coarse_fine_sides(L, L.lbox(il)));
        if (cfmask_[l].back()) cf_ = true;
      }
      // every rank must take the same path: look at the boxes of the other ranks too
      for (const Bx& b : L.boxes) if (!cf_ && coarse_fine_sides(L, b)) cf_ = true;
    }
  }
}

void CellMG::set_bc(const k::LinBC& bc) {
  bc_ = bc;
  has_bc_ = !all_periodic(*lv_[0].lev) || cf_;
}

bool CellMG::box_on_boundary(int l, int il) const {
  const Level& L = *lv_[l].lev;
  const Bx& b = L.lbox(il);
  if (cf_ && cfmask_[l][il]) return true;
  for (int d = 0; d < 3; ++d) if (!L.geom.periodic[d] && (b.lo[d] == L.domain.lo[d] || b.hi[d] == L.domain.hi[d])) return true;
  return false;
}

// the homogeneous ghost cell is +/- the adjacent cell: Neumann, reflect_odd, or Dirichlet extrapolated at order 2
static int mirror_kind(int code, int maxorder, int len) {   // 0 no, 1 even, 2 odd
  if (code == IAMRX_LINOP_NEUMANN) return 1;
  if (code == IAMRX_LINOP_REFLECT_ODD) return 2;
  if (code == IAMRX_LINOP_DIRICHLET && k::linop_bc_order(maxorder, len) == 2) return 2;
  return 0;
}

bool CellMG::bc_in_kernel(int l) const {
  if (!has_bc_) return true;
  if (cf_) return false;   // coarse-fine sides: ghost cells are extrapolated between the colours
  const Level& L = *lv_[l].lev;
  for (const Bx& b : L.boxes)   // all boxes of the level: the same decision on every rank
    for (int c = 0; c < ncomp_ && c < 3; ++c)
      for (int d = 0; d < 3; ++d) {
        if (L.geom.periodic[d]) continue;
        const int len = b.hi[d] - b.lo[d] + 1;
        if (b.lo[d] == L.domain.lo[d] && !mirror_kind(bc_.lo[c][d], bc_.maxorder, len)) return false;
        if (b.hi[d] == L.domain.hi[d] && !mirror_kind(bc_.hi[c][d], bc_.maxorder, len)) return false;
      }
  return true;
}

k::GsBC CellMG::gsbc_of(int l, int il) const {
  k::GsBC g{};
  const Level& L = *lv_[l].lev;
  const Bx& b = L.lbox(il);
  const bool ink = bc_in_kernel(l);
  for (int c = 0; c < 3; ++c)
    for (int d = 0; d < 3; ++d) {
      if (L.geom.periodic[d]) continue;
      const int len = b.hi[d] - b.lo[d] + 1;
      for (int side = 0; side < 2; ++side) {
        if (side == 0 ? b.lo[d] != L.domain.lo[d] : b.hi[d] != L.domain.hi[d]) continue;
        const int code = side == 0 ? bc_.lo[c][d] : bc_.hi[c][d];
        g.f0[c][2 * d + side] = k::linop_bc_f0(code, bc_.maxorder, len);
        const int mk = ink ? mirror_kind(code, bc_.maxorder, len) : 0;
        if (mk == 1) g.even[c] |= 1 << (2 * d + side);
        if (mk == 2) g.odd[c] |= 1 << (2 * d + side);
      }
    }
  if (cf_)
    for (int d = 0; d < 3; ++d)
      for (int side = 0; side < 2; ++side)
        if (cfmask_[l][il] & (1 << (2 * d + side)))
          for (int c = 0; c < 3; ++c) g.f0[c][2 * d + side] = k::linop_cf_f0(bc_.maxorder, b.hi[d] - b.lo[d] + 1, cf_x0(l, d));
  return g;
}

int CellMG::fill_ghosts(int l, MF& phi, bool inhomog, int wm, int grow_t, cudaStream_t s) {
  if (cf_) {   // coarse-fine sides first: the exchange below overwrites the ghost cells that fine neighbours cover
    const bool ihc = inhomog && l == 0 && bvals_.ok();
    const double x0[3] = {cf_x0(l, 0), cf_x0(l, 1), cf_x0(l, 2)};
    for (int il = 0; il < phi.n(); ++il)
      IX_TRY(k::linop_cf_fill(phi.vbox(il), phi.v(il), ncomp_, bc_.maxorder, ihc ? bvals_.c(il) : C4{}, cfmask_[l][il], x0, s));
  }
  if (wm != 7) IX_TRY(mf_fill_boundary(phi, 0, ncomp_, 1, s, wm));
  if (!has_bc_) return IAMRX_OK;
  // homogeneous fills of sides the kernels mirror in place are not needed (grow_t > 0: the tensor cross terms read the cells)
  if (!inhomog && grow_t == 0 && bc_in_kernel(l)) return IAMRX_OK;
  const Level& L = *lv_[l].lev;
  const bool ih = inhomog && l == 0 && bvals_.ok();
  for (int il = 0; il < phi.n(); ++il)
    IX_TRY(k::linop_bc_fill(phi.vbox(il), phi.v(il), ncomp_, bc_, ih ? bvals_.c(il) : C4{}, L.domain, L.geom.periodic, grow_t, 0, s));
  return IAMRX_OK;
}

k::Abec CellMG::op_at(int l, int il) const {
  const MGLevelCell& L = lv_[l];
  k::Abec op;
  op.a = a_; op.b = b_;
  op.acoef = (a_ != 0.0 && L.acoef.ok()) ? L.acoef.c(il) : C4{};
  // constant coefficients: the coarse levels' face arrays are never built (set_coeffs) nor read (k::Abec::cc)
  const bool hb = L.b[0].n() > il;
  op.bx = hb ? L.b[0].c(il) : C4{}; op.by = hb ? L.b[1].c(il) : C4{}; op.bz = hb ? L.b[2].c(il) : C4{};
  op.bncomp = tensor_ ? ncomp_ : 1;
  for (int d = 0; d < 3; ++d) op.dxinv[d] = L.dxinv[d];
  op.cc = cc_ ? 1 : 0; op.cac = cac_ ? 1 : 0;
  for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) op.cb[c][d] = cbv_[c][d];
  op.ca = cav_;
  return op;
}

// Are the coefficients handed to set_coeffs constants (constant density / viscosity: TaylorGreen, HIT, DoubleShearLayer, the
// viscous solves of every constant-mu run)?  One pass over each input array (compare with its first element) and one host
// read-back; the ranks agree through max / min reductions.  IAMRX_CONST_COEF=0 switches the detection off.
int CellMG::detect_constant(const MF* acoef, const MF* const bin[3], cudaStream_t s) {
  cc_ = false; cac_ = false;
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("IAMRX_CONST_COEF"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled) return IAMRX_OK;
  const bool ha = a_ != 0.0 && acoef;
  const int na = ha ? 4 : 3;
  struct Guard { double* p; ~Guard() { dev_free(p); } } buf{dev_alloc(12)}, g{nullptr};
  if (!buf.p) return IAMRX_ERR_CUDA;
  const double init[12] = {0, 0, 1e300, 0, 0, 1e300, 0, 0, 1e300, 0, 0, 1e300};
  IX_CUDA(cudaMemcpyAsync(buf.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
  for (int q = 0; q < na; ++q) {
    const MF& m = q < 3 ? *bin[q] : *acoef;
    for (int il = 0; il < m.n(); ++il) IX_TRY(k::const_check(m.vbox(il), m.c(il), buf.p + 3 * q, s));
  }
  double h[12];
  if (!lv_[0].lev->replicated && comm().nranks > 1) {
    // {differs, max} under MAX, {min} under MIN: regroup so that each reduction is one contiguous call
    g.p = dev_alloc(12);
    if (!g.p) return IAMRX_ERR_CUDA;
    for (int q = 0; q < 4; ++q) {
      IX_CUDA(cudaMemcpyAsync(g.p + 2 * q, buf.p + 3 * q, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
      IX_CUDA(cudaMemcpyAsync(g.p + 8 + q, buf.p + 3 * q + 2, sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
    IX_TRY(comm_allreduce(g.p, 8, 2, s));
    IX_TRY(comm_allreduce(g.p + 8, 4, 1, s));
    double t[12];
    IX_CUDA(cudaMemcpyAsync(t, g.p, sizeof(t), cudaMemcpyDeviceToHost, s));
    IX_CUDA(cudaStreamSynchronize(s));
    for (int q = 0; q < 4; ++q) { h[3 * q] = t[2 * q]; h[3 * q + 1] = t[2 * q + 1]; h[3 * q + 2] = t[8 + q]; }
  } else {
    IX_CUDA(cudaMemcpyAsync(h, buf.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    IX_CUDA(cudaStreamSynchronize(s));
  }
  cac_ = false;
  for (int q = 0; q < 3; ++q) if (h[3 * q] != 0.0 || h[3 * q + 1] != h[3 * q + 2]) return IAMRX_OK;
  cac_ = ha && h[9] == 0.0 && h[10] == h[11];
  for (int c = 0; c < 3; ++c)
    for (int d = 0; d < 3; ++d) {
      const double fac = (tensor_ && c == d) ? (4.0 / 3.0) : 1.0;
      cbv_[c][d] = fac * h[3 * d + 1];   // what mf_lincomb(fac, x, 0, x) stores
    }
  cav_ = cac_ ? h[3 * 3 + 1] : 0.0;
  cc_ = true;
  return IAMRX_OK;
}

int CellMG::set_coeffs(const MF* acoef, const MF* bx, const MF* by, const MF* bz, cudaStream_t s,
                       bool finest_only) {
  const MF* bin[3] = {bx, by, bz};
  for (int d = 0; d < 3; ++d) eta_[d] = bin[d];
  const int bn = tensor_ ? ncomp_ : 1;
  IX_TRY(detect_constant(acoef, bin, s));
  // level 0
  {
    MGLevelCell& L = lv_[0];
    if (a_ != 0.0 && acoef) {
      if (!L.acoef.ok()) L.acoef.define(L.lev, IX_CELL, 1, L.deep ? 2 * L.deep_sweeps - 1 : 0);
      IX_TRY(mf_copy(L.acoef, *acoef, 0, 0, 1, 0, s));
      if (L.deep) IX_TRY(mf_fill_boundary(L.acoef, 0, 1, L.acoef.ng, s));
    }
    for (int d = 0; d < 3; ++d) {
      if (!L.b[d].ok()) L.b[d].define(L.lev, IX_XFACE + d, bn, L.deep ? 2 * L.deep_sweeps - 1 : 0);
      // the tensor operator's nine constant face arrays are never read on the constant-coefficient path (the MAC operator's
      // are: mac_update / getFluxes take beta from them)
      for (int c = 0; c < bn && !(tensor_ && cc_); ++c) {
        const double fac = (tensor_ && c == d) ? (4.0 / 3.0) : 1.0;
        IX_TRY(mf_lincomb(L.b[d], c, fac, *bin[d], 0, 0.0, *bin[d], 0, 1, 0, s));
      }
      if (L.deep) IX_TRY(mf_fill_boundary(L.b[d], 0, bn, L.b[d].ng, s));   // deep-ghost sweeps relax the ghost layers too
    }
  }
  // constant coefficients: the coarse levels' arrays are never read (averages of a constant are that constant, exactly)
  for (size_t l = 1; l < lv_.size() && !finest_only && !(cc_ && (cac_ || !(a_ != 0.0 && acoef))); ++l) {
    MGLevelCell& C = lv_[l];
    MGLevelCell& F = lv_[l - 1];
    if (a_ != 0.0 && acoef && !cac_) {
      if (!C.acoef.ok()) C.acoef.define(C.lev, IX_CELL, 1, C.deep ? 2 * C.deep_sweeps - 1 : 0);
      if (C.xfer_lev) {   // restrict on the distributed layout, then gather into the replicated box
        MF tmp(C.xfer_lev.get(), IX_CELL, 1, 0);
        for (int il = 0; il < tmp.n(); ++il) IX_TRY(k::cc_restrict(tmp.vbox(il), tmp.v(il), F.acoef.c(il), 1, s, thin_));
        IX_TRY(mf_gather_replicate(C.acoef, tmp, 1, s));
      } else {
        for (int il = 0; il < C.acoef.n(); ++il)
          IX_TRY(k::cc_restrict(C.acoef.vbox(il), C.acoef.v(il), F.acoef.c(il), 1, s, thin_));
      }
      if (C.deep) IX_TRY(mf_fill_boundary(C.acoef, 0, 1, C.acoef.ng, s));
    }
    for (int d = 0; d < 3; ++d) {
      if (!C.b[d].ok()) C.b[d].define(C.lev, IX_XFACE + d, bn, C.deep ? 2 * C.deep_sweeps - 1 : 0);
      if (cc_) continue;
      if (C.xfer_lev) {
        MF tmp(C.xfer_lev.get(), IX_XFACE + d, bn, 0);
        for (int il = 0; il < tmp.n(); ++il) IX_TRY(k::face_restrict(tmp.vbox(il), d, tmp.v(il), F.b[d].c(il), bn, s, thin_));
        IX_TRY(mf_gather_replicate(C.b[d], tmp, bn, s));
      } else {
        for (int il = 0; il < C.b[d].n(); ++il)
          IX_TRY(k::face_restrict(C.b[d].vbox(il), d, C.b[d].v(il), F.b[d].c(il), bn, s, thin_));
      }
      if (C.deep) IX_TRY(mf_fill_boundary(C.b[d], 0, bn, C.b[d].ng, s));
    }
  }
  // the operator annihilates constants when a = 0 and no side pins the solution (periodic / Neumann everywhere)
  singular_ = (a_ == 0.0);
  if (has_bc_)
    for (int c = 0; c < ncomp_ && c < 3; ++c)
      for (int d = 0; d < 3; ++d) {
        if (lv_[0].lev->geom.periodic[d]) continue;
        for (int code : {bc_.lo[c][d], bc_.hi[c][d]})
          if (code == IAMRX_LINOP_DIRICHLET || code == IAMRX_LINOP_REFLECT_ODD) singular_ = false;
      }
  if (cf_) singular_ = false;   // coarse-fine sides pin the solution
  return IAMRX_OK;
}

int CellMG::smooth(int l, MF& phi, const MF& rhs, int nsweeps, bool zero_init, cudaStream_t s) {
  MGLevelCell& L = lv_[l];
  // a box that spans the periodic domain wraps its neighbour indices inside the kernel:
  // no ghost fill (and no extra launch) per colour
  const int wm = L.lev->level_wrapmask();   // directions wrapped inside the kernels: no ghost traffic there
  const bool wrap = wm == 7;
  if (zero_init && nsweeps <= 0) return mf_setval(phi, 0.0, 0, ncomp_, phi.ng, s);
  bool fused = wrap && k::abec_gsrb_sweep_enabled();
  for (size_t b = 0; b < L.lev->boxes.size() && fused; ++b) fused = k::abec_gsrb_sweep_ok(L.lev->boxes[b], 7);
  if (fused) {
    // one fused launch per sweep, out of place: ping-pong between phi and a second buffer
    if (zero_init) IX_TRY(mf_setval(phi, 0.0, 0, ncomp_, phi.ng, s));
    if (!L.gs_tmp.ok()) L.gs_tmp.define(L.lev, IX_CELL, ncomp_, 1);
    MF* src = &phi; MF* dst = &L.gs_tmp;
    for (int sw = 0; sw < nsweeps; ++sw) {
      for (int il = 0; il < phi.n(); ++il)
        IX_TRY(k::abec_gsrb_sweep(phi.vbox(il), dst->v(il), src->c(il), rhs.c(il), op_at(l, il), info_.omega, 0, ncomp_, s));
      std::swap(src, dst);
    }
    if (src != &phi) IX_TRY(mf_copy(phi, *src, 0, 0, ncomp_, 0, s));
    return IAMRX_OK;
  }
  if (L.deep && !has_bc_ && phi.ng >= 2 && rhs.ng >= 1) {
    // deep-ghost sweeps: ONE exchange of 2 c ghost layers for c sweeps; pass p of the batch (red, black, red, ...) runs on the box
    // grown by 2 c - 1 - p cells towards its neighbours -- the same arithmetic on the same values as the neighbour's own passes --
    // so the last black pass is on the box itself.  A zero initial guess (ghost layers included) needs no exchange for its batch.
    const int cmax = std::max(1, std::min(std::min(phi.ng / 2, (rhs.ng + 1) / 2), L.deep_sweeps));
    if (!L.rhs_ghost_ok) { IX_TRY(mf_fill_boundary(const_cast<MF&>(rhs), 0, ncomp_, std::min(rhs.ng, 2 * cmax - 1), s, wm)); L.rhs_ghost_ok = true; }
    for (int s0 = 0; s0 < nsweeps; s0 += cmax) {
      const int c = std::min(cmax, nsweeps - s0);
      const bool zero = zero_init && s0 == 0;
      if (!zero) IX_TRY(mf_fill_boundary(phi, 0, ncomp_, 2 * c, s, wm));
      for (int p = 0; p < 2 * c; ++p)
        for (int il = 0; il < phi.n(); ++il) {
          Bx b = phi.vbox(il);
          const int g = 2 * c - 1 - p;
          for (int d = 0; d < 3; ++d) if (!(wm & (1 << d))) b = grow(b, d, g);
          // zero initial guess: the first red pass (on the largest box) writes every cell the later passes read -- the coloured
          // cells and zeros in between; the outermost ghost layer is refilled by the next exchange before anything reads it
          IX_TRY(k::abec_gsrb(b, phi.v(il), rhs.c(il), op_at(l, il), info_.omega, p & 1, ncomp_, s, wm, nullptr, zero && p == 0));
        }
    }
    return IAMRX_OK;
  }
  // small single box whose neighbours are all periodic images: every sweep in one launch (IAMRX_GSRB_SMALL=0: colour launches)
  static int small_on = -1;
  if (small_on < 0) { const char* e = getenv("IAMRX_GSRB_SMALL"); small_on = (e && e[0] == '0') ? 0 : 1; }
  if (small_on && wrap && !has_bc_ && phi.n() == 1 && L.lev->boxes.size() == 1 && k::abec_gsrb_small_ok(phi.vbox(0), ncomp_))
    return k::abec_gsrb_small(phi.vbox(0), phi.v(0), rhs.c(0), op_at(l, 0), info_.omega, ncomp_, nsweeps, zero_init, s, wm, nullptr);
  for (int sw = 0; sw < nsweeps; ++sw) {
    for (int rb = 0; rb < 2; ++rb) {
      // zero_init: the caller did NOT clear phi -- the first colour pass writes every cell (the homogeneous ghost cells of a zero
      // field are zero and are not read); the fill before the second colour refreshes all of them
      const bool zero = zero_init && sw == 0 && rb == 0;
      if (!zero) IX_TRY(fill_ghosts(l, phi, false, wm, 0, s));   // corrections: homogeneous boundary conditions
      for (int il = 0; il < phi.n(); ++il) {
        const bool onb = has_bc_ && box_on_boundary(l, il);
        const k::GsBC gb = onb ? gsbc_of(l, il) : k::GsBC{};
        IX_TRY(k::abec_gsrb(phi.vbox(il), phi.v(il), rhs.c(il), op_at(l, il), info_.omega, rb, ncomp_, s, wm, onb ? &gb : nullptr, zero));
      }
    }
  }
  return IAMRX_OK;
}

// with_cross marks the top-level residual of a solve: inhomogeneous boundary conditions (+ the tensor cross terms)
int CellMG::residual(int l, MF& out, MF& phi, const MF& rhs, bool with_cross, cudaStream_t s, double* norm) {
  const bool cross = tensor_ && l == 0 && with_cross;  // the cross terms read edge / corner neighbours: periodic images in the
  // wrapped directions (in-kernel), ghost cells elsewhere; a box on a physical boundary keeps the full ghost fill
  const int wm = (cross && has_bc_) ? 0 : lv_[l].lev->level_wrapmask();
  IX_TRY(fill_ghosts(l, phi, with_cross, wm, cross ? 1 : 0, s));
  const Level& L = *lv_[l].lev;
  double* nd = nullptr;
  bool all_fused = norm != nullptr && !cross && phi.n() > 0;
  if (all_fused) IX_TRY(norm_acc_begin(&nd, s));
  for (int il = 0; il < phi.n(); ++il) {
    const bool mir = has_bc_ && !with_cross && box_on_boundary(l, il) && bc_in_kernel(l);
    const k::GsBC gb = mir ? gsbc_of(l, il) : k::GsBC{};
    bool nf = false;
    IX_TRY(k::abec_apply(phi.vbox(il), out.v(il), phi.c(il), rhs.c(il), op_at(l, il), ncomp_, s, wm, mir ? &gb : nullptr, all_fused ? nd : nullptr, &nf));
    if (all_fused && !nf) {   // this box took a kernel without the fused norm: reduce it into the same scalar
      double* one = nd;
      for (int c = 0; c < ncomp_; ++c) IX_TRY(k::reduce(phi.vbox(il), out.c(il, c), 1, 2, one, s));
    }
    if (cross) {
      if (has_bc_)
        IX_TRY(k::tensor_cross_bc(phi.vbox(il), out.v(il), phi.c(il), bvals_.ok() ? bvals_.c(il) : C4{}, eta_[0]->c(il), eta_[1]->c(il),
                                  eta_[2]->c(il), -b_, lv_[0].dxinv, bc_, L.domain, L.geom.periodic, s));
      else
        IX_TRY(k::tensor_cross(phi.vbox(il), out.v(il), phi.c(il), eta_[0]->c(il), eta_[1]->c(il),
                               eta_[2]->c(il), -b_, lv_[0].dxinv, s, wm));
    }
  }
  if (norm) {
    if (all_fused) IX_TRY(norm_acc_end(nd, L.replicated, norm, s));
    else IX_TRY(mf_norminf(out, 0, ncomp_, norm, s));
  }
  return IAMRX_OK;
}

// the level BC of a solve / apply = the ghost cells of the MF handed in (MLLinOp::setLevelBC(0, &Soln), Diffusion.cpp:743-744,886-887;
// MacProj.cpp:1168): keep a copy, the ghost cells themselves are overwritten by the extrapolated values
static int save_level_bc(MF& bvals, MF& phi, int ncomp, bool needed, cudaStream_t s) {
  if (!needed) { bvals.clear(); return IAMRX_OK; }
  if (!bvals.ok() || bvals.lev != phi.lev || bvals.ncomp != ncomp) bvals.define(phi.lev, IX_CELL, ncomp, 1);
  IX_TRY(mf_copy(bvals, phi, 0, 0, ncomp, 1, s));
  return mf_fill_boundary(bvals, 0, ncomp, 1, s);   // periodic / interior images of the wall ghost cells (tensor: transverse neighbours)
}

int CellMG::apply(MF& out, MF& phi, cudaStream_t s) {
  if (cf_ && tensor_) { set_error("CellMG: the tensor operator on a level with coarse-fine sides is not implemented"); return IAMRX_ERR_ARG; }
  bool dirichlet = false;
  for (int c = 0; c < ncomp_ && c < 3; ++c) for (int d = 0; d < 3; ++d) if (bc_.lo[c][d] == IAMRX_LINOP_DIRICHLET || bc_.hi[c][d] == IAMRX_LINOP_DIRICHLET) dirichlet = true;
  IX_TRY(save_level_bc(bvals_, phi, ncomp_, has_bc_ && (dirichlet || cf_), s));
  // directions every box spans periodically are wrapped inside the kernels (cross terms included): no ghost traffic there
  const int wm = has_bc_ ? 0 : lv_[0].lev->level_wrapmask();
  IX_TRY(fill_ghosts(0, phi, true, wm, tensor_ ? 1 : 0, s));
  const Level& L = *lv_[0].lev;
  for (int il = 0; il < phi.n(); ++il) {
    IX_TRY(k::abec_apply(phi.vbox(il), out.v(il), phi.c(il), C4{}, op_at(0, il), ncomp_, s, wm));
    if (tensor_) {
      if (has_bc_)
        IX_TRY(k::tensor_cross_bc(phi.vbox(il), out.v(il), phi.c(il), bvals_.ok() ? bvals_.c(il) : C4{}, eta_[0]->c(il), eta_[1]->c(il),
                                  eta_[2]->c(il), b_, lv_[0].dxinv, bc_, L.domain, L.geom.periodic, s));
      else
        IX_TRY(k::tensor_cross(phi.vbox(il), out.v(il), phi.c(il), eta_[0]->c(il), eta_[1]->c(il),
                               eta_[2]->c(il), b_, lv_[0].dxinv, s, wm));
    }
  }
  bvals_.clear();
  return IAMRX_OK;
}

int CellMG::make_solvable(int l, MF& rhs, cudaStream_t s) {
  for (int c = 0; c < ncomp_; ++c) {
    double sum = 0;
    IX_TRY(mf_sum(rhs, c, &sum, s));
    const double mean = sum / (double)lv_[l].lev->ncells_global;
    for (int il = 0; il < rhs.n(); ++il) IX_TRY(k::addconst(rhs.vbox(il), rhs.v(il, c), -mean, 1, s));
  }
  return IAMRX_OK;
}


// ===========================================================================
// BiCGStab bottom solver (AMReX MLCGSolver::solve_bicgstab, the first stage of IAMR's default "bicgcg" bottom solver):
// unpreconditioned, zero initial guess, homogeneous boundary conditions, converged when |r|_inf <= rtol |r0|_inf.
// Host-driven over the level operations (apply / dot / axpy), so it works on any layout and for both operators; the dot
// products use the deterministic two-pass sum.  Returns 0 (converged), or the AMReX break-down code (1 rho = 0,
// 2 <rh,v> = 0, 3 <t,t> = 0, 4 omega = 0, 8 iteration cap): the caller then falls back to the smoother as AMReX does.
// ===========================================================================
namespace {
struct KrylovOps {
  std::function<int(MF& out, MF& in)> apply;                       // out = A in (homogeneous BC; may fill in's ghosts)
  std::function<int(const MF& a, const MF& b, double* r)> dot;
  std::function<int(const MF& a, double* r)> norminf;
  int ncomp = 1;
};

int bicgstab(const KrylovOps& op, MF& sol, const MF& rhs, double rtol, int maxiter, int* iters_out, cudaStream_t s) {
  Level* lev = sol.lev;
  const int nc = op.ncomp, it = sol.ixtype;
  MF r(lev, it, nc, 0), rh(lev, it, nc, 0), p(lev, it, nc, sol.ng), v(lev, it, nc, 0), sv(lev, it, nc, sol.ng), t(lev, it, nc, 0);
  IX_TRY(mf_setval(p, 0.0, 0, nc, p.ng, s));
  IX_TRY(mf_setval(sv, 0.0, 0, nc, sv.ng, s));
  IX_TRY(mf_copy(r, rhs, 0, 0, nc, 0, s));       // x0 = 0: r = b
  IX_TRY(mf_copy(rh, r, 0, 0, nc, 0, s));
  double rnorm0 = 0;
  IX_TRY(op.norminf(r, &rnorm0));
  *iters_out = 0;
  if (rnorm0 == 0.0) return 0;
  const double target = rtol * rnorm0;
  double rho_1 = 0, alpha = 0, omega = 0;
  int ret = 8;
  for (int nit = 1; nit <= maxiter; ++nit) {
    *iters_out = nit;
    double rho = 0;
    IX_TRY(op.dot(rh, r, &rho));
    if (rho == 0.0) { ret = 1; break; }
    if (nit == 1) {
      IX_TRY(mf_copy(p, r, 0, 0, nc, 0, s));
    } else {
      const double beta = (rho / rho_1) * (alpha / omega);
      IX_TRY(mf_lincomb(p, 0, 1.0, p, 0, -omega, v, 0, nc, 0, s));     // p = p - omega v
      IX_TRY(mf_lincomb(p, 0, beta, p, 0, 1.0, r, 0, nc, 0, s));       // p = r + beta p
    }
    IX_TRY(op.apply(v, p));
    double rhTv = 0;
    IX_TRY(op.dot(rh, v, &rhTv));
    if (rhTv == 0.0) { ret = 2; break; }
    alpha = rho / rhTv;
    IX_TRY(mf_lincomb(sol, 0, 1.0, sol, 0, alpha, p, 0, nc, 0, s));
    IX_TRY(mf_lincomb(sv, 0, 1.0, r, 0, -alpha, v, 0, nc, 0, s));      // s = r - alpha v
    double rnorm = 0;
    IX_TRY(op.norminf(sv, &rnorm));
    if (!(rnorm == rnorm)) { ret = 9; break; }
    if (rnorm <= target) { ret = 0; break; }
    IX_TRY(op.apply(t, sv));
    double tt = 0, ts = 0;
    IX_TRY(op.dot(t, t, &tt));
    IX_TRY(op.dot(t, sv, &ts));
    if (tt == 0.0) { ret = 3; break; }
    omega = ts / tt;
    IX_TRY(mf_lincomb(sol, 0, 1.0, sol, 0, omega, sv, 0, nc, 0, s));
    IX_TRY(mf_lincomb(r, 0, 1.0, sv, 0, -omega, t, 0, nc, 0, s));      // r = s - omega t
    IX_TRY(op.norminf(r, &rnorm));
    if (!(rnorm == rnorm)) { ret = 9; break; }
    if (rnorm <= target) { ret = 0; break; }
    if (omega == 0.0) { ret = 4; break; }
    rho_1 = rho;
  }
  return ret;
}

// <a, b> over all components: the product goes through a temporary and the deterministic sum (unique nodes for nodal data)
int mf_dot(const MF& a, const MF& b, int ncomp, bool unique_nodes, double* out, cudaStream_t s) {
  MF tmp(a.lev, a.ixtype, ncomp, 0);
  IX_TRY(mf_copy(tmp, a, 0, 0, ncomp, 0, s));
  for (int il = 0; il < tmp.n(); ++il) IX_TRY(k::mult(tmp.vbox(il), tmp.v(il), b.c(il), ncomp, ncomp, s));
  double tot = 0;
  for (int c = 0; c < ncomp; ++c) { double v = 0; IX_TRY(mf_sum(tmp, c, &v, s, unique_nodes)); tot += v; }
  *out = tot;
  return IAMRX_OK;
}
}  // namespace

int CellMG::bottom_solve(cudaStream_t s) {
  const int nl = (int)lv_.size();
  MGLevelCell& B = lv_[nl - 1];
  IX_TRY(mf_setval(B.cor, 0.0, 0, ncomp_, B.cor.ng, s));
  B.rhs_ghost_ok = false;
  if (singular_ && nl > 1) IX_TRY(make_solvable(nl - 1, B.res, s));
  const int nsm = nl == 1 ? info_.nu1 + info_.nu2 : info_.bottom_sweeps;
  if (info_.bottom_solver != 1 || nl == 1) return smooth(nl - 1, B.cor, B.res, nsm, true, s);
  KrylovOps op;
  op.ncomp = ncomp_;
  op.apply = [&](MF& out, MF& in) -> int {       // rhs - A in with no rhs = A in (homogeneous BC, no cross terms: a coarse level)
    const int wm = B.lev->level_wrapmask();
    IX_TRY(fill_ghosts(nl - 1, in, false, wm, 0, s));
    for (int il = 0; il < in.n(); ++il) {
      const bool mir = has_bc_ && box_on_boundary(nl - 1, il) && bc_in_kernel(nl - 1);
      const k::GsBC gb = mir ? gsbc_of(nl - 1, il) : k::GsBC{};
      IX_TRY(k::abec_apply(in.vbox(il), out.v(il), in.c(il), C4{}, op_at(nl - 1, il), ncomp_, s, wm, mir ? &gb : nullptr));
    }
    return IAMRX_OK;
  };
  op.dot = [&](const MF& a, const MF& b, double* r) { return mf_dot(a, b, ncomp_, false, r, s); };
  op.norminf = [&](const MF& a, double* r) { return mf_norminf(a, 0, ncomp_, r, s); };
  int its = 0;
  const int ret = bicgstab(op, B.cor, B.res, info_.bottom_rtol, info_.bottom_maxiter, &its, s);
  if (ret < 0) return ret;
  info_.bottom_iters += its;
  if (ret != 0) {   // MLMG::bottomSolve: a failed CG solve is discarded and replaced by smoothing
    IX_TRY(mf_setval(B.cor, 0.0, 0, ncomp_, 1, s));
    IX_TRY(smooth(nl - 1, B.cor, B.res, nsm, true, s));
  }
  return IAMRX_OK;
}

int CellMG::vcycle(cudaStream_t s) {
  const int nl = (int)lv_.size();
  for (int l = 0; l < nl - 1; ++l) {
    MGLevelCell& L = lv_[l];
    L.rhs_ghost_ok = false;   // L.res was just rewritten (top-level residual or restriction)
    IX_TRY(smooth(l, L.cor, L.res, info_.nu1, true, s));   // zero initial guess: smooth() clears or overwrites L.cor
    IX_TRY(residual(l, L.rescor, L.cor, L.res, false, s));
    MGLevelCell& C = lv_[l + 1];
    if (C.xfer_lev) {
      for (int il = 0; il < C.xfer.n(); ++il)
        IX_TRY(k::cc_restrict(C.xfer.vbox(il), C.xfer.v(il), L.rescor.c(il), ncomp_, s, thin_));
      IX_TRY(mf_gather_replicate(C.res, C.xfer, ncomp_, s));
    } else {
      for (int il = 0; il < C.res.n(); ++il)
        IX_TRY(k::cc_restrict(C.res.vbox(il), C.res.v(il), L.rescor.c(il), ncomp_, s, thin_));
    }
  }
  IX_TRY(bottom_solve(s));
  for (int l = nl - 2; l >= 0; --l) {
    MGLevelCell& L = lv_[l];
    MGLevelCell& C = lv_[l + 1];
    for (int il = 0; il < L.cor.n(); ++il)
      IX_TRY(k::cc_prolong_add(L.cor.vbox(il), L.cor.v(il), C.cor.c(C.xfer_lev ? 0 : il), ncomp_, s, thin_));   // replicated coarse box: read in place
    IX_TRY(smooth(l, L.cor, L.res, info_.nu2, false, s));
  }
  return IAMRX_OK;
}

int CellMG::solve(MF& sol, const MF& rhs_in, iamrx_mg_info* info, cudaStream_t s) {
  if (cf_ && tensor_) { set_error("CellMG: the tensor operator on a level with coarse-fine sides is not implemented"); return IAMRX_ERR_ARG; }
  if (info) info_ = *info;
  info_.bottom_iters = 0;
  MGLevelCell& L0 = lv_[0];
  {
    bool dirichlet = false;
    for (int c = 0; c < ncomp_ && c < 3; ++c) for (int d = 0; d < 3; ++d) if (bc_.lo[c][d] == IAMRX_LINOP_DIRICHLET || bc_.hi[c][d] == IAMRX_LINOP_DIRICHLET) dirichlet = true;
    IX_TRY(save_level_bc(bvals_, sol, ncomp_, has_bc_ && (dirichlet || cf_), s));
  }
  MF rhs(L0.lev, IX_CELL, ncomp_, 0);
  IX_TRY(mf_copy(rhs, rhs_in, 0, 0, ncomp_, 0, s));
  if (singular_) IX_TRY(make_solvable(0, rhs, s));
  double rhsnorm = 0, resnorm0 = 0, resnorm = 0;
  IX_TRY(mf_norminf(rhs, 0, ncomp_, &rhsnorm, s));
  IX_TRY(residual(0, L0.res, sol, rhs, true, s, &resnorm0));
  if (!std::isfinite(rhsnorm) || !std::isfinite(resnorm0)) { bvals_.clear(); set_error("CellMG: NaN in the right-hand side or the initial guess"); return IAMRX_ERR_NAN; }
  const double maxnorm = std::max(rhsnorm, resnorm0);
  const double target = std::max(info_.atol, info_.rtol * maxnorm);
  resnorm = resnorm0;
  int iters = 0;
  int rc = IAMRX_OK;
  if (!(resnorm0 <= target)) {
    rc = info_.max_iter;  // pessimistic: not converged
    for (iters = 1; iters <= info_.max_iter; ++iters) {
      IX_TRY(vcycle(s));
      IX_TRY(mf_lincomb(sol, 0, 1.0, sol, 0, 1.0, L0.cor, 0, ncomp_, 0, s));
      IX_TRY(residual(0, L0.res, sol, rhs, true, s, &resnorm));
      if (info_.verbose > 1)
        fprintf(stderr, "[iamrx] CellMG iter %d resnorm %.6e (target %.3e)\n", iters, resnorm, target);
      if (!std::isfinite(resnorm)) { set_error("CellMG: NaN residual"); rc = IAMRX_ERR_NAN; break; }
      if (resnorm <= target) { rc = IAMRX_OK; break; }
    }
  }
  IX_TRY(fill_ghosts(0, sol, true, 0, 0, s));  // setFinalFillBC(true)
  bvals_.clear();
  if (info_.verbose > 0)
    fprintf(stderr, "[iamrx] CellMG: %d iters, res0 %.3e -> %.3e (rhs %.3e, levels %d)\n", iters, resnorm0,
            resnorm, rhsnorm, nlevels());
  if (info) { info->iters = iters; info->resnorm0 = resnorm0; info->resnorm = resnorm; info->rhsnorm = rhsnorm; info->bottom_iters = info_.bottom_iters; }
  if (rc > 0) set_error("CellMG: failed to converge");
  return rc;
}

// ===========================================================================
// NodeMG
// ===========================================================================
NodeMG::NodeMG(Level* fine, int max_coarsening) {
  iamrx_mg_info_default(&info_);
  thin_ = thin_mask(*fine);
  lv_.emplace_back();
  lv_[0].lev = fine;
  Level* cur = fine;
  for (int l = 1; l <= max_coarsening; ++l) {
    auto c = coarsen_level(*cur, 2, thin_);
    if (!c) break;
    lv_.emplace_back();
    if (auto r = consolidated_level(*c)) {   // from here down: one replicated box per rank, no ghost traffic
      lv_.back().xfer_lev = std::move(c);
      lv_.back().lev_owned = std::move(r);
    } else {
      lv_.back().lev_owned = std::move(c);
    }
    lv_.back().lev = lv_.back().lev_owned.get();
    cur = lv_.back().lev;
  }
  static int deep_on = -1;
  if (deep_on < 0) { const char* e = getenv("IAMRX_NODAL_DEEP"); deep_on = (e && e[0] == '0') ? 0 : 1; }
  for (auto& L : lv_) {
    for (int d = 0; d < 3; ++d) L.dxinv[d] = L.lev->dxinv[d];
    // deep-ghost level (decided from ALL boxes, the same on every rank): some x / y direction is not wrapped in the kernels,
    // it is periodic (every box side has a neighbour or a periodic image; walls in x / y keep the colour path), and every box
    // qualifies for the fused sweep with ghost-layer halos
    const int wm = L.lev->level_wrapmask();
    L.deep = deep_on && !L.lev->replicated && (wm & 3) != 3;
    for (int d = 0; d < 2 && L.deep; ++d) if (!(wm & (1 << d)) && !L.lev->geom.periodic[d]) L.deep = false;
    for (size_t b = 0; b < L.lev->boxes.size() && L.deep; ++b)
      L.deep = k::nodal_gs_sweep_ok(ixbox(L.lev->boxes[b], IX_NODE), wm | k::NODAL_DEEP_GHOSTS);
    L.ngd = L.deep ? 4 : 1;
    L.sigma.define(L.lev, IX_CELL, 1, L.ngd);
    L.cor.define(L.lev, IX_NODE, 1, L.ngd);
    L.res.define(L.lev, IX_NODE, 1, L.ngd);
    L.rescor.define(L.lev, IX_NODE, 1, 1);
    if (L.xfer_lev) L.xfer.define(L.xfer_lev.get(), IX_NODE, 1, 1);
    // nodes ON Dirichlet sides are never written by the kernels (active_nbox): they must hold zero from the start
    for (MF* m : {&L.cor, &L.res, &L.rescor, &L.xfer}) if (m->ok()) mf_setval(*m, 0.0, 0, 1, m->ng, nullptr);
  }
  for (int d = 0; d < 3; ++d) { bc_.lo[d] = IAMRX_LINOP_PERIODIC; bc_.hi[d] = IAMRX_LINOP_PERIODIC; }
  // a level that does not tile its domain is a fine AMR level: the sides of the patch that border coarse cells carry Dirichlet nodes
  if (!fine->replicated && fine->ncells_global != mkbx(fine->geom.domain).npts()) {
    cfmask_.resize(lv_.size());
    for (size_t l = 0; l < lv_.size(); ++l) {
      const Level& L = *lv_[l].lev;
      for (int il = 0; il < L.nlocal(); ++il) cfmask_[l].push_back(coarse_fine_sides(L, L.lbox(il)));
      for (const Bx& b : L.boxes) if (coarse_fine_sides(L, b)) cf_ = true;
    }
    // the boundary nodes are found side by side: the boxes must form ONE rectangular patch (no re-entrant edges, no partly
    // covered sides)
    Bx u = fine->boxes[0];
    for (const Bx& b : fine->boxes) for (int d = 0; d < 3; ++d) { u.lo[d] = std::min(u.lo[d], b.lo[d]); u.hi[d] = std::max(u.hi[d], b.hi[d]); }
    cf_rect_ = u.npts() == fine->ncells_global;
    static int force_mask = -1;   // IAMRX_NODAL_CF_MASK=1: the node-mask path on rectangular patches too (cross-check)
    if (force_mask < 0) { const char* e = getenv("IAMRX_NODAL_CF_MASK"); force_mask = (e && e[0] == '1') ? 1 : 0; }
    cf_mask_ = cf_ && (!cf_rect_ || force_mask);
    if (cf_mask_) {
      // dmask = 1 on nodes with an uncovered cell among the cells around them that lie inside the domain (periodic images count)
      for (auto& M : lv_) {
        Level* LL = M.lev;
        MF cov(LL, IX_CELL, 1, 1);
        mf_setval(cov, 0.0, 0, 1, 1, nullptr);
        for (int il = 0; il < cov.n(); ++il) k::setval(cov.vbox(il), cov.v(il), 1, 1.0, nullptr);
        mf_fill_boundary(cov, 0, 1, 1, nullptr);
        for (int il = 0; il < cov.n(); ++il)
          for (int d = 0; d < 3; ++d) {
            if (LL->geom.periodic[d]) continue;
            for (int side = 0; side < 2; ++side) {   // cells beyond a physical side do not exist: they never make a node a coarse-fine node
              Bx R = grow(cov.vbox(il), 1);
              if (side == 0) R.hi[d] = std::min(R.hi[d], LL->domain.lo[d] - 1); else R.lo[d] = std::max(R.lo[d], LL->domain.hi[d] + 1);
              if (R.ok()) k::setval(R, cov.v(il), 1, 1.0, nullptr);
            }
          }
        M.dmask.define(LL, IX_NODE, 1, 0);
        for (int il = 0; il < M.dmask.n(); ++il) k::sync_mask(M.dmask.vbox(il), M.dmask.v(il), cov.c(il), 7.5, nullptr);
      }
      cudaStreamSynchronize(nullptr);
    }
  }
}

int NodeMG::apply_node_mask(int l, MF& a, cudaStream_t s) const {
  if (!cf_mask_) return IAMRX_OK;
  const MGLevelNode& M = lv_[l];
  for (int il = 0; il < a.n(); ++il) IX_TRY(k::mask_zero(a.vbox(il), a.v(il), M.dmask.c(il), s));
  return IAMRX_OK;
}

void NodeMG::set_bc(const k::NodalBC& bc) {
  bc_ = bc;
  has_bc_ = !all_periodic(*lv_[0].lev) || cf_;
}

Bx NodeMG::active_nbox(int l, int il, bool with_cf) const {
  const Level& L = *lv_[l].lev;
  Bx nb = ixbox(L.lbox(il), IX_NODE);
  if (!has_bc_) return nb;
  if (cf_ && with_cf && !cf_mask_)
    for (int d = 0; d < 3; ++d) {
      if (cfmask_[l][il] & (1 << (2 * d))) nb.lo[d] += 1;
      if (cfmask_[l][il] & (1 << (2 * d + 1))) nb.hi[d] -= 1;
    }
  for (int d = 0; d < 3; ++d) {
    if (L.geom.periodic[d]) continue;
    if (bc_.lo[d] == IAMRX_LINOP_DIRICHLET && L.lbox(il).lo[d] == L.domain.lo[d]) nb.lo[d] += 1;
    if (bc_.hi[d] == IAMRX_LINOP_DIRICHLET && L.lbox(il).hi[d] == L.domain.hi[d]) nb.hi[d] -= 1;
  }
  return nb;
}

int NodeMG::neumann_sides(int l, int il) const {
  if (!has_bc_) return 0;
  const Level& L = *lv_[l].lev;
  int m = 0;
  for (int d = 0; d < 3; ++d) {
    if (L.geom.periodic[d]) continue;
    if ((bc_.lo[d] == IAMRX_LINOP_NEUMANN || bc_.lo[d] == IAMRX_LINOP_INFLOW) && L.lbox(il).lo[d] == L.domain.lo[d]) m |= 1 << (2 * d);
    if ((bc_.hi[d] == IAMRX_LINOP_NEUMANN || bc_.hi[d] == IAMRX_LINOP_INFLOW) && L.lbox(il).hi[d] == L.domain.hi[d]) m |= 1 << (2 * d + 1);
  }
  return m;
}

bool NodeMG::singular() const {
  if (cf_) return false;   // the coarse-fine boundary nodes pin the solution
  const Level& L = *lv_[0].lev;
  for (int d = 0; d < 3; ++d)
    if (!L.geom.periodic[d] && has_bc_ && (bc_.lo[d] == IAMRX_LINOP_DIRICHLET || bc_.hi[d] == IAMRX_LINOP_DIRICHLET)) return false;
  return true;
}

// bc_fill = false: the caller's kernels mirror the Neumann sides in place (wm | neumann_sides << 3), only the exchange is needed
int NodeMG::fill_ghosts(int l, MF& phi, int wm, cudaStream_t s, bool bc_fill, int depth) {
  if (wm != 7) IX_TRY(mf_fill_boundary(phi, 0, 1, depth, s, wm));
  if (!has_bc_ || !bc_fill) return IAMRX_OK;
  const Level& L = *lv_[l].lev;
  const Bx ndom = ixbox(L.domain, IX_NODE);
  for (int il = 0; il < phi.n(); ++il)
    IX_TRY(k::nodal_bc_fill_phi(phi.vbox(il), phi.v(il), bc_, ndom, L.geom.periodic, 0, s));
  return IAMRX_OK;
}

int NodeMG::set_sigma(const MF& sigma, cudaStream_t s) {
  auto sigma_bc = [&](MGLevelNode& M) -> int {   // mlndlap_fillbc_cc: the ghost layer beyond a Neumann / inflow side copies the interior
    if (!has_bc_) return IAMRX_OK;
    const Level& LL = *M.lev;
    for (int il = 0; il < M.sigma.n(); ++il)
      IX_TRY(k::nodal_bc_fill_sigma(M.sigma.vbox(il), M.sigma.v(il), bc_, LL.domain, LL.geom.periodic, s, M.ngd));
    return IAMRX_OK;
  };
  IX_TRY(mf_setval(lv_[0].sigma, 0.0, 0, 1, lv_[0].ngd, s));
  IX_TRY(mf_copy(lv_[0].sigma, sigma, 0, 0, 1, 0, s));
  IX_TRY(mf_fill_boundary(lv_[0].sigma, 0, 1, lv_[0].ngd, s));
  IX_TRY(sigma_bc(lv_[0]));
  for (size_t l = 1; l < lv_.size(); ++l) {
    MGLevelNode& C = lv_[l];
    MGLevelNode& F = lv_[l - 1];
    IX_TRY(mf_setval(C.sigma, 0.0, 0, 1, C.ngd, s));
    if (C.xfer_lev) {
      MF tmp(C.xfer_lev.get(), IX_CELL, 1, 0);
      for (int il = 0; il < tmp.n(); ++il) IX_TRY(k::cc_restrict(tmp.vbox(il), tmp.v(il), F.sigma.c(il), 1, s, thin_));
      IX_TRY(mf_gather_replicate(C.sigma, tmp, 1, s));
    } else {
      for (int il = 0; il < C.sigma.n(); ++il)
        IX_TRY(k::cc_restrict(C.sigma.vbox(il), C.sigma.v(il), F.sigma.c(il), 1, s, thin_));
    }
    IX_TRY(mf_fill_boundary(C.sigma, 0, 1, C.ngd, s));
    IX_TRY(sigma_bc(C));
  }
  return IAMRX_OK;
}

static int nodal_smoother_kind() {
  static int kind = -1;
  if (kind < 0) {
    const char* e = getenv("IAMRX_NODAL_SMOOTHER");
    kind = (e && e[0] == 'j') ? 1 : 0;  // 0 = 8-colour Gauss-Seidel, 1 = damped Jacobi
  }
  return kind;
}

int NodeMG::smooth(int l, MF& phi, const MF& rhs, int nsweeps, cudaStream_t s) {
  MGLevelNode& L = lv_[l];
  const int wm = L.lev->level_wrapmask();   // directions wrapped inside the kernels: no ghost traffic there
  const bool wrap = wm == 7;
  if (nodal_smoother_kind() == 1) {
    MF tmp(L.lev, IX_NODE, 1, 1);
    for (int sw = 0; sw < 2 * nsweeps; ++sw) {
      IX_TRY(fill_ghosts(l, phi, 0, s, true));
      for (int il = 0; il < phi.n(); ++il)
        IX_TRY(k::nodal_jacobi(active_nbox(l, il), tmp.v(il), phi.c(il), rhs.c(il), L.sigma.c(il), L.dxinv,
                               2.0 / 3.0, s));
      IX_TRY(mf_copy(phi, tmp, 0, 0, 1, 0, s));
      IX_TRY(apply_node_mask(l, phi, s));
    }
    return IAMRX_OK;
  }
  // decided from ALL boxes of the level (not just this rank's): the fused and the colour paths issue different
  // numbers of ghost exchanges, so every rank must take the same one
  // deep-ghost levels (block decompositions): same fused sweep, halos from 4 ghost layers of phi / rhs / sigma instead of wraps
  const bool deep = L.deep && phi.ng >= L.ngd && rhs.ng >= L.ngd;
  // a small level that is ONE box (the coarse levels; consolidated levels of multi-rank runs): every sweep in one launch
  if (!cf_mask_ && L.lev->boxes.size() == 1 && phi.n() == 1 && k::nodal_gs_small_ok(active_nbox(l, 0))) {
    IX_TRY(fill_ghosts(l, phi, wm, s, false));
    return k::nodal_gs_small(active_nbox(l, 0), phi.v(0), rhs.c(0), L.sigma.c(0), L.dxinv, nsweeps, s, wm | (neumann_sides(l, 0) << 3));
  }
  const int wmk = deep ? (wm | k::NODAL_DEEP_GHOSTS) : wm;
  const int gd = deep ? L.ngd : 1;
  bool fused = true;
  for (size_t b = 0; b < L.lev->boxes.size() && fused; ++b) fused = k::nodal_gs_sweep_ok(ixbox(L.lev->boxes[b], IX_NODE), wmk);
  // Dirichlet sides shorten the active node box; keep the fused sweep only if its plane pairing survives (both z sides or none)
  if (cf_) fused = false;   // coarse-fine Dirichlet planes shorten the active boxes side by side: colour passes
  if (fused && has_bc_)
    for (int d = 0; d < 3; ++d)
      if (!L.lev->geom.periodic[d] && ((bc_.lo[d] == IAMRX_LINOP_DIRICHLET) != (bc_.hi[d] == IAMRX_LINOP_DIRICHLET))) fused = false;
  if (fused) {
    // out-of-place fused sweeps ping-pong between phi and a second buffer.  Slabs (x and y wrapped in the kernel, z
    // exchanged): the even-plane phase reads the old odd ghost planes of `src`, the odd-plane phase the NEW even ghost
    // planes of `dst` -- two plane exchanges per sweep instead of eight colour fills.
    if (!L.gs_tmp.ok()) { L.gs_tmp.define(L.lev, IX_NODE, 1, L.ngd); IX_TRY(mf_setval(L.gs_tmp, 0.0, 0, 1, L.ngd, s)); }
    MF* src = &phi; MF* dst = &L.gs_tmp;
    bool zeven = true;   // every box starts (and, having an even extent, ends) on an even node plane
    for (const Bx& b : L.lev->boxes) if (b.lo[2] & 1) zeven = false;
    // (the ghost layers of the right-hand side are scratch: filling them does not change the caller's data)
    if (deep) IX_TRY(mf_fill_boundary(const_cast<MF&>(rhs), 0, 1, gd, s, wm));
    IX_TRY(fill_ghosts(l, *src, wm, s, false, gd));
    for (int sw = 0; sw < nsweeps; ++sw) {
      for (int phase = 0; phase < 2; ++phase) {
        for (int il = 0; il < phi.n(); ++il)
          IX_TRY(k::nodal_gs_sweep(active_nbox(l, il), dst->v(il), src->c(il), rhs.c(il), L.sigma.c(il), L.dxinv, s,
                                   wmk | (neumann_sides(l, il) << 3), phase));
        // slabs (x, y wrapped; node boxes start and end on even planes): the ghost planes are ODD planes, which only the second
        // phase changes, and the second phase reads nothing beyond the box's own even planes -- one exchange per sweep.  Boxes
        // with x / y neighbours (deep-ghost halos) need the new even planes of their neighbours for the second phase.
        if (phase == 1 || deep || !zeven) IX_TRY(fill_ghosts(l, *dst, wm, s, false, gd));
      }
      std::swap(src, dst);
    }
    if (src != &phi) IX_TRY(mf_copy(phi, *src, 0, 0, 1, wrap ? 0 : 1, s));
    return IAMRX_OK;
  }
  for (int sw = 0; sw < nsweeps; ++sw) {
    for (int color = 0; color < 8; ++color) {
      IX_TRY(fill_ghosts(l, phi, wm, s, false));
      for (int il = 0; il < phi.n(); ++il)
        IX_TRY(k::nodal_gs_color(active_nbox(l, il), phi.v(il), rhs.c(il), L.sigma.c(il), L.dxinv, color, s, wm | (neumann_sides(l, il) << 3)));
      IX_TRY(apply_node_mask(l, phi, s));   // (the smoother only ever works on corrections: their Dirichlet nodes are zero)
    }
  }
  return IAMRX_OK;
}

int NodeMG::residual(int l, MF& out, MF& phi, const MF& rhs, cudaStream_t s, double* norm) {
  MGLevelNode& L = lv_[l];
  const int wm = L.lev->level_wrapmask();
  IX_TRY(fill_ghosts(l, phi, wm, s, false));
  // the norm is taken inside the residual kernel when every box's active nodes are all of its nodes (no Dirichlet planes whose
  // entries of `out` the kernel does not write) and the kernel supports it
  double* nd = nullptr;
  bool fusedn = norm != nullptr && phi.n() > 0 && !cf_mask_;
  for (int il = 0; il < phi.n() && fusedn; ++il) {
    const Bx a = active_nbox(l, il), v = phi.vbox(il);
    for (int d = 0; d < 3; ++d) if (a.lo[d] != v.lo[d] || a.hi[d] != v.hi[d]) fusedn = false;
  }
  if (fusedn) IX_TRY(norm_acc_begin(&nd, s));
  bool all = fusedn;
  for (int il = 0; il < phi.n(); ++il) {
    bool nf = false;
    IX_TRY(k::nodal_adotx(active_nbox(l, il), out.v(il), phi.c(il), rhs.c(il), L.sigma.c(il), L.dxinv, s, wm | (neumann_sides(l, il) << 3),
                          fusedn ? nd : nullptr, &nf));
    if (fusedn && !nf) all = false;
  }
  IX_TRY(apply_node_mask(l, out, s));
  if (norm) {
    if (fusedn && all) IX_TRY(norm_acc_end(nd, L.lev->replicated, norm, s));
    else IX_TRY(mf_norminf(out, 0, 1, norm, s));
  }
  return IAMRX_OK;
}

// periodic / Neumann everywhere: make rhs solvable (MLNodeLinOp::getSolvabilityOffset / fixSolvabilityByOffset): subtract the
// mean weighted with the dot mask -- unique nodes, 1/2 per Neumann side a node lies on.  The weights sum to the number of
// cells (a periodic direction has n unique nodes, a Neumann-Neumann one n + 1 with two halves).
int NodeMG::make_solvable(int l, MF& rhs, cudaStream_t s) {
  Level& lev = *lv_[l].lev;
  double sum = 0;
  if (has_bc_) {
    MF tmp(&lev, IX_NODE, 1, 0);
    IX_TRY(mf_copy(tmp, rhs, 0, 0, 1, 0, s));
    const Bx ndom = ixbox(lev.domain, IX_NODE);
    for (int il = 0; il < tmp.n(); ++il) IX_TRY(k::nodal_bc_scale(tmp.vbox(il), tmp.v(il), bc_, ndom, lev.geom.periodic, 0.5, s));
    IX_TRY(mf_sum(tmp, 0, &sum, s, true));
  } else {
    IX_TRY(mf_sum(rhs, 0, &sum, s, true));
  }
  const double mean = sum / (double)lev.ncells_global;
  for (int il = 0; il < rhs.n(); ++il) IX_TRY(k::addconst(rhs.vbox(il), rhs.v(il), -mean, 1, s));
  return IAMRX_OK;
}

int NodeMG::bottom_solve(cudaStream_t s) {
  const int nl = (int)lv_.size();
  MGLevelNode& B = lv_[nl - 1];
  IX_TRY(mf_setval(B.cor, 0.0, 0, 1, 1, s));
  const int nsm = nl == 1 ? info_.nu1 + info_.nu2 : info_.bottom_sweeps;
  if (info_.bottom_solver != 1 || nl == 1) return smooth(nl - 1, B.cor, B.res, nsm, s);
  if (singular()) IX_TRY(make_solvable(nl - 1, B.res, s));   // MLMG::bottomSolve: a singular bottom problem is made solvable first
  KrylovOps op;
  op.ncomp = 1;
  op.apply = [&](MF& out, MF& in) -> int {
    const int wm = B.lev->level_wrapmask();
    IX_TRY(fill_ghosts(nl - 1, in, wm, s, false));
    IX_TRY(mf_setval(out, 0.0, 0, 1, 0, s));   // nodes ON Dirichlet sides stay zero (active_nbox)
    for (int il = 0; il < in.n(); ++il)
      IX_TRY(k::nodal_adotx(active_nbox(nl - 1, il), out.v(il), in.c(il), C4{}, B.sigma.c(il), B.dxinv, s, wm | (neumann_sides(nl - 1, il) << 3)));
    IX_TRY(apply_node_mask(nl - 1, out, s));
    return IAMRX_OK;
  };
  op.dot = [&](const MF& a, const MF& b, double* r) { return mf_dot(a, b, 1, true, r, s); };
  op.norminf = [&](const MF& a, double* r) { return mf_norminf(a, 0, 1, r, s); };
  int its = 0;
  const int ret = bicgstab(op, B.cor, B.res, info_.bottom_rtol, info_.bottom_maxiter, &its, s);
  if (ret < 0) return ret;
  info_.bottom_iters += its;
  if (ret != 0) {
    IX_TRY(mf_setval(B.cor, 0.0, 0, 1, 1, s));
    IX_TRY(smooth(nl - 1, B.cor, B.res, nsm, s));
  }
  return IAMRX_OK;
}

int NodeMG::vcycle(cudaStream_t s) {
  const int nl = (int)lv_.size();
  for (int l = 0; l < nl - 1; ++l) {
    MGLevelNode& L = lv_[l];
    IX_TRY(mf_setval(L.cor, 0.0, 0, 1, 1, s));
    IX_TRY(smooth(l, L.cor, L.res, info_.nu1, s));
    IX_TRY(residual(l, L.rescor, L.cor, L.res, s));
    // MLNodeLaplacian::restriction: applyBC on the fine residual (Neumann sides mirrored); directions every box spans are wrapped
    // inside the restriction kernel, so a slab exchanges whole planes in place here too
    const int rwm = L.lev->level_wrapmask();
    IX_TRY(fill_ghosts(l, L.rescor, rwm, s, true));
    MGLevelNode& C = lv_[l + 1];
    if (C.xfer_lev) {
      // (the transfer level shares the boundary flags of the coarse level: same domain, boxes tile it)
      for (int il = 0; il < C.xfer.n(); ++il) {
        Bx cb = C.xfer.vbox(il);
        const Level& XL = *C.xfer_lev;
        if (has_bc_)
          for (int d = 0; d < 3; ++d) {
            if (XL.geom.periodic[d]) continue;
            if (bc_.lo[d] == IAMRX_LINOP_DIRICHLET && XL.lbox(il).lo[d] == XL.domain.lo[d]) cb.lo[d] += 1;
            if (bc_.hi[d] == IAMRX_LINOP_DIRICHLET && XL.lbox(il).hi[d] == XL.domain.hi[d]) cb.hi[d] -= 1;
          }
        const Bx fnb = L.rescor.vbox(il);
        IX_TRY(k::nodal_restrict(cb, C.xfer.v(il), L.rescor.c(il), s, thin_, rwm, &fnb));
      }
      IX_TRY(mf_gather_replicate(C.res, C.xfer, 1, s));
    } else {
      for (int il = 0; il < C.res.n(); ++il) {
        const Bx fnb = L.rescor.vbox(il);
        IX_TRY(k::nodal_restrict(active_nbox(l + 1, il), C.res.v(il), L.rescor.c(il), s, thin_, rwm, &fnb));
      }
      IX_TRY(apply_node_mask(l + 1, C.res, s));
    }
  }
  IX_TRY(bottom_solve(s));
  for (int l = nl - 2; l >= 0; --l) {
    MGLevelNode& L = lv_[l];
    MGLevelNode& C = lv_[l + 1];
    for (int il = 0; il < L.cor.n(); ++il)
      IX_TRY(k::nodal_interp_add(active_nbox(l, il), L.cor.v(il), C.cor.c(C.xfer_lev ? 0 : il), s, thin_));
    IX_TRY(apply_node_mask(l, L.cor, s));
    IX_TRY(smooth(l, L.cor, L.res, info_.nu2, s));
  }
  return IAMRX_OK;
}

int NodeMG::solve(MF& phi, MF& rhs, iamrx_mg_info* info, cudaStream_t s) {
  if (info) info_ = *info;
  info_.bottom_iters = 0;
  MGLevelNode& L0 = lv_[0];
  Level& lev = *L0.lev;
  (void)lev;
  if (singular()) IX_TRY(make_solvable(0, rhs, s));
  double rhsnorm = 0, resnorm0 = 0, resnorm = 0;
  IX_TRY(mf_norminf(rhs, 0, 1, &rhsnorm, s));
  IX_TRY(residual(0, L0.res, phi, rhs, s, &resnorm0));
  if (!std::isfinite(rhsnorm) || !std::isfinite(resnorm0)) { set_error("NodeMG: NaN in the right-hand side or the initial guess"); return IAMRX_ERR_NAN; }
  const double maxnorm = std::max(rhsnorm, resnorm0);
  const double target = std::max(info_.atol, info_.rtol * maxnorm);
  resnorm = resnorm0;
  int iters = 0;
  int rc = IAMRX_OK;
  if (!(resnorm0 <= target)) {
    rc = info_.max_iter;
    for (iters = 1; iters <= info_.max_iter; ++iters) {
      IX_TRY(vcycle(s));
      IX_TRY(mf_lincomb(phi, 0, 1.0, phi, 0, 1.0, L0.cor, 0, 1, 0, s));
      IX_TRY(residual(0, L0.res, phi, rhs, s, &resnorm));
      if (info_.verbose > 1)
        fprintf(stderr, "[iamrx] NodeMG iter %d resnorm %.6e (target %.3e)\n", iters, resnorm, target);
      if (!std::isfinite(resnorm)) { set_error("NodeMG: NaN residual"); rc = IAMRX_ERR_NAN; break; }
      if (resnorm <= target) { rc = IAMRX_OK; break; }
    }
  }
  IX_TRY(fill_ghosts(0, phi, 0, s, true));
  if (info_.verbose > 0)
  {
    int ndeep = 0;
    for (const auto& L : lv_) ndeep += L.deep ? 1 : 0;
    fprintf(stderr, "[iamrx] NodeMG: %d iters, res0 %.3e -> %.3e (rhs %.3e, levels %d, deep-ghost levels %d)\n", iters, resnorm0,
            resnorm, rhsnorm, nlevels(), ndeep);
  }
  if (info) { info->iters = iters; info->resnorm0 = resnorm0; info->resnorm = resnorm; info->rhsnorm = rhsnorm; info->bottom_iters = info_.bottom_iters; }
  if (rc > 0) set_error("NodeMG: failed to converge");
  return rc;
}

}  // namespace ix
