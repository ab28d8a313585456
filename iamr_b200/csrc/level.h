// level.h -- host data model: a level's BoxArray + DistributionMapping, local
// multi-fabs with a caching device allocator, the FillBoundary plan, and the
// communicator (NCCL, loaded at run time).  Host-side stand-in for the subset of
// amrex::{BoxArray, DistributionMapping, MultiFab, FabArray::FillBoundary,
// ParallelDescriptor} the hot path touches.
#pragma once
#include <map>
#include <memory>
#include <vector>
#include "kernels.h"

namespace ix {

enum IxType { IX_CELL = 0, IX_XFACE = 1, IX_YFACE = 2, IX_ZFACE = 3, IX_NODE = 4 };

inline Bx ixbox(const Bx& cell, int ixtype) {
  Bx b = cell;
  if (ixtype >= IX_XFACE && ixtype <= IX_ZFACE) b.hi[ixtype - 1] += 1;
  if (ixtype == IX_NODE) { b.hi[0]++; b.hi[1]++; b.hi[2]++; }
  return b;
}

// ---- device memory pool ---------------------------------------------------
// Size-bucketed free lists over cudaMalloc; blocks are reused without
// synchronisation because every consumer runs on one stream per rank (the
// reference's The_Arena / The_Async_Arena role, NSB.cpp:4735,4780).
double* dev_alloc(size_t ndoubles);
void dev_free(double* p);
void dev_pool_release();
size_t dev_pool_bytes();

// ---- communicator -----------------------------------------------------------
struct Comm {
  int rank = 0, nranks = 1;
  void* nccl = nullptr;  // ncclComm_t
  // host-supplied transport (iamrx_comm_set_transport); overrides NCCL when set
  iamrx_exchange_fn ex = nullptr;
  iamrx_allreduce_fn ar = nullptr;
  void* ctx = nullptr;
};
Comm& comm();
int comm_allreduce(double* dev, int n, int op, cudaStream_t s);
// grouped point-to-point: for each peer p, send sbuf[p] (scount doubles) and
// receive rbuf[p] (rcount doubles)
int comm_exchange(const std::vector<int>& peers, const std::vector<double*>& sbuf,
                  const std::vector<int64_t>& scount, const std::vector<double*>& rbuf,
                  const std::vector<int64_t>& rcount, cudaStream_t s);

// peer-memory transport (p2p.cu): carries the entries of an exchange whose pair fits a mailbox slot over NVLink (CUDA IPC mapped
// mailboxes + device-side sequence flags); handled[i] = 1 for those.  false: transport off (the default; IAMRX_P2P=1 turns it on).
bool p2p_try_exchange(const std::vector<int>& peers, const std::vector<double*>& sbuf, const std::vector<int64_t>& scount,
                      const std::vector<double*>& rbuf, const std::vector<int64_t>& rcount, cudaStream_t s, std::vector<char>& handled, int* rc);
void p2p_finalize();

// ---- copy descriptors (FillBoundary / ParallelCopy data movers) -----------
struct CopyDesc {
  int kind;        // 0 fab->fab, 1 fab->buf, 2 buf->fab
  int dst, src;    // local fab indices
  int lo[3], hi[3];  // region in DST index space (kind 1: in SRC index space)
  int sh[3];       // src index = dst index + sh
  int64_t bufoff;  // doubles, for kind 1/2
};

struct FabTable {  // passed by value to the batched copy kernel
  static constexpr int MAXF = 128;   // local boxes per rank (two tables = 11 KB of kernel parameters; sm_100 allows 32 KB)
  double* p[MAXF];
  int lo[MAXF][3];
  int64_t js[MAXF], ks[MAXF], ns[MAXF];
};

namespace k {
int copy_batch(const CopyDesc* d_desc, int ndesc, int64_t max_pts, const FabTable& dst, const FabTable& src,
               double* buf, int ncomp, int64_t buf_comp_stride, cudaStream_t s);
}

struct FBPlan {
  int ixtype = 0, ng = 0;
  std::vector<CopyDesc> local;      // kind 0
  CopyDesc* d_local = nullptr;
  // remote
  std::vector<int> peers;
  std::vector<std::vector<CopyDesc>> send, recv;  // per peer (kind 1 / kind 2)
  std::vector<int64_t> send_pts, recv_pts;        // points per component per peer
  CopyDesc* d_send = nullptr; int n_send = 0;
  CopyDesc* d_recv = nullptr; int n_recv = 0;
  int64_t send_total = 0, recv_total = 0;         // points per component, all peers
  int64_t max_local = 0, max_send = 0, max_recv = 0;  // largest region (points) per list: sizes the copy grid
  std::vector<int64_t> send_off, recv_off;
  // every remote region is a stack of whole z planes (x, y skipped and spanned, shift in z only) between boxes of equal
  // x-y shape: planes of MF-owned fabs are then contiguous and identical in layout on both sides, and are sent straight
  // from / received straight into the arrays (no pack / unpack kernels)
  bool direct = false;
  ~FBPlan();
};

// gather of a distributed level's data into a replicated single-box level (multigrid consolidation)
struct GatherPlan {
  std::vector<CopyDesc> pack, local, unpack;
  CopyDesc* d_pack = nullptr; CopyDesc* d_local = nullptr; CopyDesc* d_unpack = nullptr;
  std::vector<int> peers;
  std::vector<int64_t> recv_off, recv_pts;   // points per component, per peer
  int64_t send_pts = 0, recv_total = 0, max_pack = 0, max_local = 0, max_unpack = 0;
  ~GatherPlan();
};

struct Level {
  iamrx_geom geom;
  std::vector<Bx> boxes;   // all boxes of the level (cell index space)
  std::vector<int> owner;  // rank of each box
  std::vector<int> local;  // global indices of the boxes this rank owns
  std::map<std::pair<int, int>, std::unique_ptr<FBPlan>> plans;
  std::map<int, std::unique_ptr<GatherPlan>> gplans;   // by ixtype
  // replicated level: ONE box covering the whole domain, held (and computed on) by every rank identically -- the
  // consolidated coarse multigrid levels.  No ghost traffic, and reductions need no all-reduce.
  bool replicated = false;
  GatherPlan& gather_plan(int ixtype);
  double dxinv[3];
  Bx domain;
  int64_t ncells_global = 0;

  int nlocal() const { return (int)local.size(); }
  const Bx& lbox(int il) const { return boxes[local[il]]; }
  // skip: bit d set = no ghost cells are needed in direction d (kernels wrap indices there): regions are clipped to
  // the valid range of direction d and dropped when empty
  FBPlan& plan(int ixtype, int ng, int skip = 0);
  // bit d set iff local box il spans the whole periodic domain in direction d: its periodic
  // neighbour is the box itself and kernels can wrap indices instead of reading ghost cells
  int wrapmask(int il) const {
    int m = 0;
    for (int d = 0; d < 3; ++d)
      if (geom.periodic[d] && lbox(il).lo[d] == domain.lo[d] && lbox(il).hi[d] == domain.hi[d]) m |= (1 << d);
    return m;
  }
  // true when every local box wraps in all three directions (ghost fills can be skipped)
  bool all_wrap() const {
    for (int il = 0; il < nlocal(); ++il) if (wrapmask(il) != 7) return false;
    return true;
  }
  // directions in which EVERY box of the level (on any rank) spans the periodic domain: the same value on all ranks,
  // so it can steer both the kernels (in-kernel wrap) and the exchange plan (no ghost traffic in those directions)
  int level_wrapmask() const {
    int m = 7;
    for (const Bx& b : boxes)
      for (int d = 0; d < 3; ++d)
        if (!(geom.periodic[d] && b.lo[d] == domain.lo[d] && b.hi[d] == domain.hi[d])) m &= ~(1 << d);
    return boxes.empty() ? 0 : m;
  }
};

// Build the list of (dst box, src box, periodic shift, region) ghost copies for
// one level -- pure host logic, also exercised by the CPU tests through
// iamrx_debug_fb_plan.
void build_fb_regions(const Level& L, int ixtype, int ng,
                      std::vector<int>& dst_box, std::vector<int>& src_box,
                      std::vector<Bx>& region, std::vector<int>& shift3, int skip = 0);

// ---- local multifab ---------------------------------------------------------
struct MF {
  Level* lev = nullptr;
  int ixtype = 0, ncomp = 0, ng = 0;
  std::vector<iamrx_fab> fabs;
  std::vector<double*> owned;

  MF() = default;
  MF(Level* L, int ixtype_, int ncomp_, int ng_) { define(L, ixtype_, ncomp_, ng_); }
  MF(const MF&) = delete;
  MF& operator=(const MF&) = delete;
  MF(MF&& o) noexcept { *this = std::move(o); }
  MF& operator=(MF&& o) noexcept;
  ~MF() { clear(); }
  void define(Level* L, int ixtype_, int ncomp_, int ng_);
  void alias(Level* L, int ixtype_, int ncomp_, int ng_, const iamrx_fab* f);  // caller-owned
  void clear();
  bool ok() const { return lev != nullptr; }
  int n() const { return (int)fabs.size(); }
  Bx vbox(int il) const { return ixbox(lev->lbox(il), ixtype); }       // valid index box
  Bx gbox(int il, int g) const { return grow(vbox(il), g); }
  V4 v(int il, int comp = 0) const { return view(&fabs[il], comp); }
  C4 c(int il, int comp = 0) const { return cview(&fabs[il], comp); }
};

// level-wide helpers (loop over local fabs)
int mf_setval(MF& m, double v, int comp, int ncomp, int ng, cudaStream_t s);
int mf_copy(MF& dst, const MF& src, int scomp, int dcomp, int ncomp, int ng, cudaStream_t s);
int mf_lincomb(MF& dst, int dcomp, double a, const MF& x, int xcomp, double b, const MF& y, int ycomp,
               int ncomp, int ng, cudaStream_t s);
int mf_scale(MF& m, double c, int comp, int ncomp, int ng, cudaStream_t s);
int mf_fill_boundary(MF& m, int comp, int ncomp, int ng, cudaStream_t s, int skip = 0);
// AmrLevel::FillPatch outside the domain: physical boundary conditions of cell data (after mf_fill_boundary); bc holds the
// BCRec / ext_dir values of components comp .. comp+ncomp-1 in its slots 0 .. ncomp-1
int mf_fill_physbc(MF& m, int comp, int ncomp, int ng, const k::PhysBC& bc, cudaStream_t s);
// dst (one box covering the domain, on a replicated level) <- valid regions of every box of src (a distributed level
// of the same resolution): local boxes by a copy kernel, remote ones by an all-to-all of packed boxes
int mf_gather_replicate(MF& dst, const MF& src, int ncomp, cudaStream_t s);
// reductions over valid regions, all ranks (blocking: returns host value).
// For face/nodal data shared points are counted once per owning box (norms only).
int mf_norminf(const MF& m, int comp, int ncomp, double* out_host, cudaStream_t s);  // max over comps
int mf_norminf_each(const MF& m, int comp, int ncomp, double* out_host, cudaStream_t s);
int norm_acc_begin(double** dev, cudaStream_t s);   // max-norm accumulated inside a producer kernel: zeroed device scalar ...
int norm_acc_end(double* dev, bool replicated, double* out_host, cudaStream_t s);   // ... reduced over the ranks, to the host
int mf_sum(const MF& m, int comp, double* out_host, cudaStream_t s, bool unique_nodes = false);
int mf_min(const MF& m, int comp, double* out_host, cudaStream_t s);

}  // namespace ix
