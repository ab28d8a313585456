// mlmg.h -- geometric multigrid drivers (host orchestration).
//   CellMG : (a*alpha - b div beta grad) on cell centres, ncomp components,
//            optional MLTensorOp cross terms at the finest level.
//   NodeMG : div(sigma grad) on nodes (Q1 finite elements).
// Host-side restatement of the subset of AMReX MLMG / MLABecLaplacian /
// MLTensorOp / MLNodeLaplacian that IAMR drives (MacProj.cpp:1084-1184,
// Projection.cpp:2385-2567, Diffusion.cpp:207-957); SURVEY.md Appendix A.6-A.9.
#pragma once
#include "level.h"

namespace ix {

struct MGLevelCell {
  std::unique_ptr<Level> lev_owned;  // coarse levels own their Level
  Level* lev = nullptr;
  // consolidation: this level is the first REPLICATED one (one box = the domain on every rank); xfer_lev is the
  // distributed coarsening of the level above, where restrictions land before they are gathered
  std::unique_ptr<Level> xfer_lev;
  MF xfer;
  MF acoef;            // 1 comp, 0 ghost (only if a != 0)
  MF b[3];             // face coefs, bncomp comps
  MF cor, res, rescor; // ncomp
  MF gs_tmp;           // second phi buffer of the out-of-place fused red-black sweep (lazy)
  double dxinv[3];
  // deep-ghost level (boxes with neighbours, fully periodic domain): cor carries 2 ghost layers, res / coefficients 1, so that a
  // red-black sweep needs ONE ghost exchange (the red pass also updates the first ghost layer, redundantly with the
  // neighbour) instead of one per colour; the right-hand side's ghost layer is exchanged once per V-cycle visit
  bool deep = false;
  int deep_sweeps = 0;   // sweeps per exchange on a deep level (ghost depths: cor 2 m, res / coefficients 2 m - 1)
  bool rhs_ghost_ok = false;
};

class CellMG {
 public:
  // fine-level coefficient MFs are referenced, not copied (caller keeps them alive)
  CellMG(Level* fine, int ncomp, bool tensor, int max_coarsening);
  int set_scalars(double a, double b) { a_ = a; b_ = b; return 0; }
  // MLLinOp::setDomainBC + setMaxOrder (MacProj.cpp:1164,1172; Diffusion.cpp:715-724).  The level BC (setLevelBC: the Dirichlet
  // face values) is taken from the ghost cells of the solution / input MF at the start of solve() / apply().
  void set_bc(const k::LinBC& bc);
  // acoef: cell MF (1 comp); eta[d]: face MFs with 1 comp.  For tensor the solver
  // builds per-component b = eta * (1 + 1/3 delta_{c,d}) (4/3 on the diagonal).
  int set_coeffs(const MF* acoef, const MF* bx, const MF* by, const MF* bz, cudaStream_t s,
                 bool finest_only = false);
  int apply(MF& out, MF& phi, cudaStream_t s);   // out = L(phi) incl. cross terms; fills phi ghosts
  int solve(MF& sol, const MF& rhs, iamrx_mg_info* info, cudaStream_t s);
  int nlevels() const { return (int)lv_.size(); }
  const MF& bcoef(int d) const { return lv_[0].b[d]; }
  double bscalar() const { return b_; }
  k::Abec op_at(int mglev, int il) const;
  // the level does not tile its domain (a fine AMR level): box sides that are neither domain faces nor covered by other boxes
  // are coarse-fine sides -- Dirichlet, data half a coarse cell beyond the face (MLCellLinOp with setCoarseFineBC,
  // MacProj.cpp:1164-1167); the values are the ghost cells of the solution on entry (iamrx_set_coarse_fine_bc)
  bool has_coarse_fine() const { return cf_; }

 private:
  bool cf_ = false;
  std::vector<std::vector<int>> cfmask_;   // [mg level][local box]: bit 2 d + side
  double cf_x0(int l, int d) const { return -0.5 * 2.0 * lv_[l].dxinv[d] / lv_[0].dxinv[d]; }   // refinement ratio 2
  int smooth(int l, MF& phi, const MF& rhs, int nsweeps, bool zero_init, cudaStream_t s);
  // norm (optional): max-norm of `out`, taken inside the residual kernel where it can be
  int residual(int l, MF& out, MF& phi, const MF& rhs, bool with_cross, cudaStream_t s, double* norm = nullptr);
  int vcycle(cudaStream_t s);
  int bottom_solve(cudaStream_t s);   // smoother sweeps or BiCGStab (iamrx_mg_info.bottom_solver) on the coarsest level
  int make_solvable(int l, MF& rhs, cudaStream_t s);
  // ghost cells of phi on level l: FillBoundary (interior / periodic, minus the in-kernel wrapped directions) then the domain
  // boundary conditions; inhomog: Dirichlet values from bvals_ (finest level only), else homogeneous
  int fill_ghosts(int l, MF& phi, bool inhomog, int wm, int grow_t, cudaStream_t s);
  k::GsBC gsbc_of(int l, int il) const;
  bool bc_in_kernel(int l) const;   // every non-periodic side of every box of level l can be mirrored inside the kernels
  bool box_on_boundary(int l, int il) const;
  int detect_constant(const MF* acoef, const MF* const bin[3], cudaStream_t s);
  bool cc_ = false, cac_ = false;   // constant-coefficient fast path (k::Abec::cc, cac)
  double cbv_[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, cav_ = 0;
  k::LinBC bc_{};
  bool has_bc_ = false;
  MF bvals_;   // copy of the level-BC data (valid + 1 ghost) while a solve / apply is in flight
  std::vector<MGLevelCell> lv_;
  int ncomp_;
  bool tensor_;
  double a_ = 0, b_ = 1;
  bool singular_ = false;
  const MF* eta_[3] = {nullptr, nullptr, nullptr};
  iamrx_mg_info info_;
  int thin_ = 0;   // semi-coarsening mask (thin_mask of the finest level)
};

struct MGLevelNode {
  MF dmask;            // nodal: 1 on the coarse-fine boundary nodes of a fine AMR level (NodeMG::node_mask() only)
  std::unique_ptr<Level> lev_owned;
  Level* lev = nullptr;
  std::unique_ptr<Level> xfer_lev;   // see MGLevelCell
  MF xfer;
  MF sigma;            // cell, 1 ghost
  MF cor, res, rescor; // nodal, 1 ghost
  MF gs_tmp;           // second phi buffer of the out-of-place fused Gauss-Seidel sweep (lazy)
  double dxinv[3];
  // deep-ghost level: boxes have neighbours in x / y (block decomposition); sigma / cor / res / gs_tmp carry 4 ghost layers so
  // that the fused two-phase Gauss-Seidel sweep can take its tile halos from them (k::NODAL_DEEP_GHOSTS)
  bool deep = false;
  int ngd = 1;         // ghost depth of sigma, cor, res, gs_tmp
};

class NodeMG {
 public:
  NodeMG(Level* fine, int max_coarsening);
  void set_bc(const k::NodalBC& bc);   // NodalProjector::setDomainBC (Projection.cpp:2436-2464,2513)
  const k::NodalBC& bc() const { return bc_; }
  bool has_bc() const { return has_bc_; }
  // node box of local box il on level l without the planes ON Dirichlet domain sides (those nodes are held at zero)
  // with_cf = false: only the Dirichlet DOMAIN sides are removed (their nodes are held at zero; the nodes ON coarse-fine sides keep
  // the values handed in -- the interpolated coarse pressure, Projection.cpp:236-239)
  Bx active_nbox(int l, int il, bool with_cf = true) const;
  // a fine AMR level (one rectangular patch of boxes that does not tile the domain): the nodes on the patch boundary are Dirichlet
  // nodes of the single-level solve (MLNodeLaplacian: coarse-fine boundary nodes of the coarsest AMR level of a solve)
  bool has_coarse_fine() const { return cf_; }
  // a fine level of general shape (re-entrant edges, partly covered sides): the boundary nodes are found node by node (dmask) and
  // reset to zero after every kernel that may have written them, instead of being cut off side by side
  bool node_mask() const { return cf_mask_; }
  int apply_node_mask(int l, MF& a, cudaStream_t s) const;
  int neumann_sides(int l, int il) const;   // bit 2d / 2d+1: the low / high side of the box is a Neumann / inflow domain side
  int set_sigma(const MF& sigma, cudaStream_t s);  // copies + coarsens
  int solve(MF& phi, MF& rhs, iamrx_mg_info* info, cudaStream_t s);
  int nlevels() const { return (int)lv_.size(); }
  const MF& sigma0() const { return lv_[0].sigma; }

 private:
  int smooth(int l, MF& phi, const MF& rhs, int nsweeps, cudaStream_t s);
  int residual(int l, MF& out, MF& phi, const MF& rhs, cudaStream_t s, double* norm = nullptr);
  int vcycle(cudaStream_t s);
  int bottom_solve(cudaStream_t s);
  int make_solvable(int l, MF& rhs, cudaStream_t s);
  int fill_ghosts(int l, MF& phi, int wm, cudaStream_t s, bool bc_fill, int depth = 1);   // FillBoundary + mirrored ghost nodes of the Neumann sides
  bool singular() const;
  std::vector<MGLevelNode> lv_;
  iamrx_mg_info info_;
  int thin_ = 0;   // semi-coarsening mask (thin_mask of the finest level)
  k::NodalBC bc_{};
  bool has_bc_ = false;
  bool cf_ = false, cf_rect_ = true, cf_mask_ = false;
  std::vector<std::vector<int>> cfmask_;   // [mg level][local box]: bit 2 d + side = that side of the box is a coarse-fine side
};

// coarsen a level by 2 (all boxes must be coarsenable); nullptr if not possible
int thin_mask(const Level& f);   // directions (bit mask) a multigrid hierarchy on this level never coarsens: <= 2 cells thick
std::unique_ptr<Level> coarsen_level(const Level& f, int min_width, int thin = 0);
// consolidation policy: given the distributed coarsening `c` of a level, return the replicated single-box level that
// should replace it (small boxes: ghost exchanges are pure latency there), or nullptr to stay distributed
std::unique_ptr<Level> consolidated_level(const Level& c);
std::unique_ptr<Level> make_level(const iamrx_geom& g, const std::vector<Bx>& boxes,
                                  const std::vector<int>& owner);

}  // namespace ix
