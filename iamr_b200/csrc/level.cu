// level.cu -- see level.h.  Memory pool, NCCL communicator (dlopen'd so the
// library loads on CPU-only hosts), FillBoundary plan + batched copy kernel,
// local multifabs and level-wide helpers.
#include <cstdlib>
#include "level.h"
#include <dlfcn.h>
#include <algorithm>
#include <mutex>

namespace ix {

// ---------------------------------------------------------------------------
// errors / accounting
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static std::mutex g_err_mu;
static std::string g_err_shared;
void set_error(const std::string& s) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err_shared = s;
}
const char* last_error_cstr() {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err = g_err_shared;
  return g_err.c_str();
}
std::atomic<int64_t> g_launches{0};

bool device_ok() {
  static int state = -1;
  if (state < 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    state = (e == cudaSuccess && n > 0) ? 1 : 0;
    if (e != cudaSuccess) cudaGetLastError();
  }
  return state == 1;
}

// ---------------------------------------------------------------------------
// per-launch timing
// ---------------------------------------------------------------------------
#if defined(IX_EMUL)
ProfScope::ProfScope(int, int64_t, double, cudaStream_t s_) : s(s_) {}
ProfScope::~ProfScope() {}
int prof_enable(int, int64_t) { return IAMRX_OK; }
void prof_reset() {}
int prof_report(int, double* ms, int64_t* n, double* b) { if (ms) *ms = 0; if (n) *n = 0; if (b) *b = 0; return IAMRX_OK; }
int prof_all(int) { return IAMRX_OK; }
int prof_dump(char* buf, int cap) { if (buf && cap > 0) buf[0] = 0; return 0; }
#else
namespace {
struct ProfRec { int kclass; cudaEvent_t e0, e1; double bytes; };
struct Prof {
  bool on = false;
  int64_t min_points = 0;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> free_events;
  std::mutex mu;
  cudaEvent_t get() {
    if (!free_events.empty()) { cudaEvent_t e = free_events.back(); free_events.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
};
Prof& prof() { static Prof p; return p; }
}  // namespace
ProfScope::ProfScope(int kclass, int64_t points, double algo_bytes, cudaStream_t s_) : s(s_) {
  Prof& P = prof();
  if (!P.on || points < P.min_points) return;
  std::lock_guard<std::mutex> lk(P.mu);
  ProfRec r{kclass, P.get(), P.get(), algo_bytes};
  cudaEventRecord(r.e0, s);
  slot = (int)P.recs.size();
  P.recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  Prof& P = prof();
  std::lock_guard<std::mutex> lk(P.mu);
  cudaEventRecord(P.recs[slot].e1, s);
}
// --- all-kernel timing by name (iamrx_prof_all / iamrx_prof_dump) ---
namespace {
struct NameRec { const char* name; cudaEvent_t e0, e1; };
struct ProfAll { bool on = false; std::vector<NameRec> recs; };
ProfAll& profall() { static ProfAll p; return p; }
}  // namespace
KernelTimer::KernelTimer(const char* name, void* stream) : slot(-1), s(stream) {
  ProfAll& A = profall();
  if (!A.on) return;
  Prof& P = prof();
  std::lock_guard<std::mutex> lk(P.mu);
  NameRec r{name, P.get(), P.get()};
  cudaEventRecord(r.e0, (cudaStream_t)s);
  slot = (int)A.recs.size();
  A.recs.push_back(r);
}
KernelTimer::~KernelTimer() {
  if (slot < 0) return;
  cudaEventRecord(profall().recs[slot].e1, (cudaStream_t)s);
}
int prof_all(int on) { profall().on = on != 0; return IAMRX_OK; }
int prof_dump(char* buf, int cap) {
  ProfAll& A = profall();
  Prof& P = prof();
  IX_CUDA(cudaDeviceSynchronize());
  std::map<std::string, std::pair<double, int64_t>> agg;
  {
    std::lock_guard<std::mutex> lk(P.mu);
    for (auto& r : A.recs) {
      float t = 0;
      if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) { agg[r.name].first += t; agg[r.name].second++; }
      else cudaGetLastError();
      P.free_events.push_back(r.e0); P.free_events.push_back(r.e1);
    }
    A.recs.clear();
  }
  std::string out;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), (long long)kv.second.second, kv.second.first);
    out += line;
  }
  if (buf && cap > 0) { strncpy(buf, out.c_str(), (size_t)cap - 1); buf[cap - 1] = 0; }
  return (int)out.size();
}
int prof_enable(int on, int64_t min_points) {
  Prof& P = prof();
  P.on = on != 0; P.min_points = min_points;
  return IAMRX_OK;
}
void prof_reset() {
  Prof& P = prof();
  std::lock_guard<std::mutex> lk(P.mu);
  for (auto& r : P.recs) { P.free_events.push_back(r.e0); P.free_events.push_back(r.e1); }
  P.recs.clear();
}
int prof_report(int kclass, double* total_ms, int64_t* launches, double* algo_bytes) {
  Prof& P = prof();
  IX_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(P.mu);
  double ms = 0, by = 0; int64_t n = 0;
  for (auto& r : P.recs) {
    if (r.kclass != kclass) continue;
    float t = 0;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); continue; }
    ms += t; by += r.bytes; n++;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  if (algo_bytes) *algo_bytes = by;
  return IAMRX_OK;
}
#endif

// ---------------------------------------------------------------------------
// device pool
// ---------------------------------------------------------------------------
namespace {
struct Pool {
  std::multimap<size_t, double*> free_;
  std::map<double*, size_t> size_;
  size_t total = 0;
  std::mutex mu;
};
Pool& pool() { static Pool p; return p; }
size_t round_up(size_t n) {
  // bucket to 1/8 octave to get reuse across slightly different shapes
  size_t b = 512;
  while (b < n) b <<= 1;
  size_t step = b / 16;
  if (step == 0) return b;
  size_t r = ((n + step - 1) / step) * step;
  return r;
}
}  // namespace

double* dev_alloc(size_t nd) {
  Pool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  size_t want = round_up(nd ? nd : 1);
  auto it = P.free_.lower_bound(want);
  if (it != P.free_.end() && it->first <= want + want / 8) {
    double* p = it->second;
    P.free_.erase(it);
    return p;
  }
  double* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want * sizeof(double));
  if (e != cudaSuccess) {
    // release cached blocks and retry once
    for (auto& kv : P.free_) { cudaFree(kv.second); P.total -= kv.first * 8; P.size_.erase(kv.second); }
    P.free_.clear();
    e = cudaMalloc(&p, want * sizeof(double));
    if (e != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return nullptr; }
  }
  P.size_[p] = want;
  P.total += want * 8;
  return p;
}
void dev_free(double* p) {
  if (!p) return;
  Pool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  auto it = P.size_.find(p);
  if (it == P.size_.end()) return;
  P.free_.insert({it->second, p});
}
void dev_pool_release() {
  Pool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  for (auto& kv : P.free_) { cudaFree(kv.second); P.total -= kv.first * 8; P.size_.erase(kv.second); }
  P.free_.clear();
}
size_t dev_pool_bytes() { return pool().total; }

// ---------------------------------------------------------------------------
// communicator (NCCL through dlopen)
// ---------------------------------------------------------------------------
namespace {
typedef int (*nccl_uid_fn)(void*);
typedef int (*nccl_init_fn)(void**, int, /*ncclUniqueId by value*/ struct Uid128, int);
struct Uid128 { char b[128]; };
typedef int (*nccl_init_fn2)(void**, int, Uid128, int);
typedef int (*nccl_destroy_fn)(void*);
typedef int (*nccl_group_fn)(void);
typedef int (*nccl_sendrecv_fn)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_send_fn)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);

struct Nccl {
  void* h = nullptr;
  nccl_uid_fn uid = nullptr;
  nccl_init_fn2 init = nullptr;
  nccl_destroy_fn destroy = nullptr;
  nccl_group_fn gstart = nullptr, gend = nullptr;
  nccl_send_fn send = nullptr;
  nccl_sendrecv_fn recv = nullptr;
  nccl_allreduce_fn allreduce = nullptr;
  nccl_errstr_fn errstr = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { set_error(std::string("dlopen libnccl failed: ") + dlerror()); return false; }
    uid = (nccl_uid_fn)dlsym(h, "ncclGetUniqueId");
    init = (nccl_init_fn2)dlsym(h, "ncclCommInitRank");
    destroy = (nccl_destroy_fn)dlsym(h, "ncclCommDestroy");
    gstart = (nccl_group_fn)dlsym(h, "ncclGroupStart");
    gend = (nccl_group_fn)dlsym(h, "ncclGroupEnd");
    send = (nccl_send_fn)dlsym(h, "ncclSend");
    recv = (nccl_sendrecv_fn)dlsym(h, "ncclRecv");
    allreduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
    errstr = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
    if (!uid || !init || !gstart || !gend || !send || !recv || !allreduce) {
      set_error("libnccl: missing symbols"); return false;
    }
    return true;
  }
};
Nccl& nccl() { static Nccl n; return n; }
constexpr int NCCL_FLOAT64 = 8;  // ncclDouble
constexpr int NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3;
#define IX_NCCL(call)                                                                    \
  do {                                                                                   \
    int r_ = (call);                                                                     \
    if (r_ != 0) {                                                                       \
      set_error(std::string(#call) + ": " + (nccl().errstr ? nccl().errstr(r_) : "nccl error")); \
      return IAMRX_ERR_COMM;                                                             \
    }                                                                                    \
  } while (0)
}  // namespace

Comm& comm() { static Comm c; return c; }

int comm_unique_id(unsigned char uid[128]) {
  if (!nccl().load()) return IAMRX_ERR_COMM;
  IX_NCCL(nccl().uid(uid));
  return IAMRX_OK;
}
int comm_init(int rank, int nranks, const unsigned char uid[128]) {
  Comm& c = comm();
  c.rank = rank; c.nranks = nranks;
  if (nranks <= 1) return IAMRX_OK;
  if (!nccl().load()) return IAMRX_ERR_COMM;
  Uid128 u; memcpy(u.b, uid, 128);
  IX_NCCL(nccl().init(&c.nccl, nranks, u, rank));
  return IAMRX_OK;
}
int comm_set_transport(int rank, int nranks, iamrx_exchange_fn ex, iamrx_allreduce_fn ar, void* ctx) {
  Comm& c = comm();
  c.rank = rank; c.nranks = nranks; c.ex = ex; c.ar = ar; c.ctx = ctx;
  return IAMRX_OK;
}
int comm_finalize() {
  Comm& c = comm();
  c.ex = nullptr; c.ar = nullptr; c.ctx = nullptr;
  p2p_finalize();
  if (c.nccl) { nccl().destroy(c.nccl); c.nccl = nullptr; }
  c.rank = 0; c.nranks = 1;
  return IAMRX_OK;
}
int comm_allreduce(double* dev, int n, int op, cudaStream_t s) {
  Comm& c = comm();
  if (c.nranks <= 1) return IAMRX_OK;
  if (c.ar) {
    if (c.ar(c.ctx, dev, n, op, (void*)s) != 0) { set_error("host transport: allreduce failed"); return IAMRX_ERR_COMM; }
    return IAMRX_OK;
  }
  if (!c.nccl) { set_error("communicator not initialised"); return IAMRX_ERR_COMM; }
  const int nop = (op == 0) ? NCCL_SUM : (op == 1 ? NCCL_MIN : NCCL_MAX);
  IX_NCCL(nccl().allreduce(dev, dev, (size_t)n, NCCL_FLOAT64, nop, c.nccl, s));
  return IAMRX_OK;
}
int comm_exchange(const std::vector<int>& peers, const std::vector<double*>& sbuf,
                  const std::vector<int64_t>& scount, const std::vector<double*>& rbuf,
                  const std::vector<int64_t>& rcount, cudaStream_t s) {
  Comm& c = comm();
  if (c.nranks <= 1 || peers.empty()) return IAMRX_OK;
  if (c.ex) {
    if (c.ex(c.ctx, (int)peers.size(), peers.data(), sbuf.data(), scount.data(), rbuf.data(), rcount.data(), (void*)s) != 0) {
      set_error("host transport: exchange failed"); return IAMRX_ERR_COMM;
    }
    return IAMRX_OK;
  }
  if (!c.nccl) { set_error("communicator not initialised"); return IAMRX_ERR_COMM; }
  // peer-memory transport (p2p.cu) for the pairs it can carry; NCCL for the rest
  std::vector<char> handled;
  int prc = IAMRX_OK;
  const bool p2p = p2p_try_exchange(peers, sbuf, scount, rbuf, rcount, s, handled, &prc);
  if (prc != IAMRX_OK) return prc;
  bool rest = !p2p;
  if (p2p) for (size_t i = 0; i < peers.size(); ++i) if (!handled[i] && (scount[i] > 0 || rcount[i] > 0)) rest = true;
  if (!rest) return IAMRX_OK;
  IX_NCCL(nccl().gstart());
  for (size_t i = 0; i < peers.size(); ++i) {
    if (p2p && handled[i]) continue;
    if (scount[i] > 0) IX_NCCL(nccl().send(sbuf[i], (size_t)scount[i], NCCL_FLOAT64, peers[i], c.nccl, s));
    if (rcount[i] > 0) IX_NCCL(nccl().recv(rbuf[i], (size_t)rcount[i], NCCL_FLOAT64, peers[i], c.nccl, s));
  }
  IX_NCCL(nccl().gend());
  return IAMRX_OK;
}

// ---------------------------------------------------------------------------
// batched copy kernel: grid = (chunks, ndesc); each CTA walks rows of its
// descriptor's region with x fastest.
// ---------------------------------------------------------------------------
namespace k {
namespace {
constexpr int CB_T = 256;
constexpr int CB_CHUNKS = 1024;

__global__ void __launch_bounds__(CB_T)
copy_batch_kernel(const CopyDesc* __restrict__ desc, IX_KARG(FabTable) dst, IX_KARG(FabTable) src, double* buf, int ncomp,
                  int64_t bstride) {
  const CopyDesc& d = desc[blockIdx.y];
  const int nx = d.hi[0] - d.lo[0] + 1, ny = d.hi[1] - d.lo[1] + 1, nz = d.hi[2] - d.lo[2] + 1;
  const int64_t npts = (int64_t)nx * ny * nz;
  (void)bstride;
  // "rows": x lines of the region, or whole xy planes when the region is thin in x (x-face slabs), so that a
  // CTA always sweeps a long contiguous-in-index run and the divisions are per row, not per point
  const bool flat = nx < 32;
  const int rowlen = flat ? nx * ny : nx;
  const int nrows = (flat ? nz : ny * nz) * ncomp;
  const int rpn = flat ? nz : ny * nz;   // rows per component
  const double* sp = (d.kind != 2) ? src.p[d.src] : nullptr;
  double* dp = (d.kind != 1) ? dst.p[d.dst] : nullptr;
  // threads cover (row, position in row) jointly: RW = power of two >= rowlen (at most CB_T) lanes per row and
  // CB_T / RW rows per CTA pass, so short rows do not serialise into one dependent load-store round trip each
  int RW = 32;
  while (RW < rowlen && RW < CB_T) RW <<= 1;
  const int rpp = CB_T / RW;                       // rows per pass
  const int tx = threadIdx.x & (RW - 1), ty = threadIdx.x / RW;
  for (int row = blockIdx.x * rpp + ty; row < nrows; row += gridDim.x * rpp) {
    const int n = row / rpn, rr = row - n * rpn;
    const int kk = flat ? rr : rr / ny, jj0 = flat ? 0 : rr - kk * ny;
    const int k = d.lo[2] + kk;
    int64_t sbase = 0, dbase = 0;
    if (d.kind != 2) sbase = (int64_t)(k + d.sh[2] - src.lo[d.src][2]) * src.ks[d.src] + (int64_t)n * src.ns[d.src] + (d.sh[0] + d.lo[0] - src.lo[d.src][0]) +
                             (int64_t)(d.sh[1] + d.lo[1] + jj0 - src.lo[d.src][1]) * src.js[d.src];
    if (d.kind != 1) dbase = (int64_t)(k - dst.lo[d.dst][2]) * dst.ks[d.dst] + (int64_t)n * dst.ns[d.dst] + (d.lo[0] - dst.lo[d.dst][0]) +
                             (int64_t)(d.lo[1] + jj0 - dst.lo[d.dst][1]) * dst.js[d.dst];
    const int64_t bbase = d.bufoff * ncomp + (int64_t)n * npts + (int64_t)kk * nx * ny + (int64_t)jj0 * nx;
    const int sjs = (d.kind != 2) ? (int)src.js[d.src] : 0, djs = (d.kind != 1) ? (int)dst.js[d.dst] : 0;
    for (int m = tx; m < rowlen; m += RW) {
      int ii = m, jj = 0;
      if (flat) { jj = m / nx; ii = m - jj * nx; }
      const double v = (d.kind == 2) ? buf[bbase + m] : sp[sbase + ii + jj * sjs];
      if (d.kind == 1) buf[bbase + m] = v;
      else dp[dbase + ii + jj * djs] = v;
    }
  }
}
}  // namespace

int copy_batch(const CopyDesc* d_desc, int ndesc, int64_t max_pts, const FabTable& dst, const FabTable& src,
               double* buf, int ncomp, int64_t bstride, cudaStream_t s) {
  if (ndesc <= 0) return IAMRX_OK;
  // CTAs per region: one point per thread of the largest region (all loads in flight at once: these copies are
  // latency-, not bandwidth-bound), at most CB_CHUNKS
  int64_t chunks = (max_pts * ncomp + CB_T - 1) / CB_T;
  if (chunks < 1) chunks = 1;
  if (chunks > CB_CHUNKS) chunks = CB_CHUNKS;
  IX_LAUNCH(copy_batch_kernel, dim3((unsigned)chunks, ndesc, 1), CB_T, 0, s, d_desc, dst, src, buf, ncomp, bstride);
  return check_launch("copy_batch");
}
}  // namespace k

// ---------------------------------------------------------------------------
// FillBoundary plan
// ---------------------------------------------------------------------------
FBPlan::~FBPlan() {
  if (d_local) cudaFree(d_local);
  if (d_send) cudaFree(d_send);
  if (d_recv) cudaFree(d_recv);
}

// subtract box `b` from `a`: returns up to 6 disjoint boxes covering a \ b
static void box_diff(const Bx& a, const Bx& b, std::vector<Bx>& out) {
  Bx isect = intersect(a, b);
  if (!isect.ok()) { out.push_back(a); return; }
  Bx rest = a;
  for (int d = 2; d >= 0; --d) {  // peel z, then y, then x so x-rows stay long
    if (rest.lo[d] < isect.lo[d]) { Bx p = rest; p.hi[d] = isect.lo[d] - 1; out.push_back(p); rest.lo[d] = isect.lo[d]; }
    if (rest.hi[d] > isect.hi[d]) { Bx p = rest; p.lo[d] = isect.hi[d] + 1; out.push_back(p); rest.hi[d] = isect.hi[d]; }
  }
}

void build_fb_regions(const Level& L, int ixtype, int ng, std::vector<int>& dst_box,
                      std::vector<int>& src_box, std::vector<Bx>& region, std::vector<int>& shift3, int skip) {
  const int nb = (int)L.boxes.size();
  int plen[3];
  for (int d = 0; d < 3; ++d) plen[d] = L.geom.domain.hi[d] - L.geom.domain.lo[d] + 1;
  for (int bi = 0; bi < nb; ++bi) {
    const Bx vdst = ixbox(L.boxes[bi], ixtype);
    const Bx gdst = grow(vdst, ng);
    // ghost region = gdst \ vdst, as disjoint pieces
    std::vector<Bx> todo;
    box_diff(gdst, vdst, todo);
    // Candidate sources in a fixed order: (box index, shift).  The first source
    // covering a point wins; pieces already filled are removed, so every ghost
    // point is written exactly once (deterministic for face/nodal overlaps).
    // periodic images up to the distance the ghost depth reaches: more than one period when the domain is thinner than the
    // ghost layer (two-layer "2-D" domains with 3 ghost cells)
    int smax[3];
    for (int d = 0; d < 3; ++d) smax[d] = L.geom.periodic[d] ? (ng + plen[d] - 1) / plen[d] : 0;
    for (int bj = 0; bj < nb && !todo.empty(); ++bj) {
      const Bx vsrc = ixbox(L.boxes[bj], ixtype);
      for (int sz = -smax[2]; sz <= smax[2]; ++sz) for (int sy = -smax[1]; sy <= smax[1]; ++sy) for (int sx = -smax[0]; sx <= smax[0]; ++sx) {
        const int sh[3] = {sx, sy, sz};
        bool okp = true;
        for (int d = 0; d < 3; ++d) if (sh[d] != 0 && !L.geom.periodic[d]) okp = false;
        // a skipped direction needs no ghost points, and its valid range is covered by the unshifted sources: without this a
        // periodic image would claim the shared node / face column of nodal and face data first and break up the plane
        for (int d = 0; d < 3; ++d) if (sh[d] != 0 && (skip & (1 << d))) okp = false;
        if (!okp) continue;
        if (bj == bi && sx == 0 && sy == 0 && sz == 0) continue;
        Bx s = vsrc;  // source valid box shifted INTO dst index space
        for (int d = 0; d < 3; ++d) { s.lo[d] += sh[d] * plen[d]; s.hi[d] += sh[d] * plen[d]; }
        std::vector<Bx> next;
        for (const Bx& piece : todo) {
          Bx is = intersect(piece, s);
          if (!is.ok()) { next.push_back(piece); continue; }
          dst_box.push_back(bi); src_box.push_back(bj); region.push_back(is);
          for (int d = 0; d < 3; ++d) shift3.push_back(-sh[d] * plen[d]);  // src = dst + shift
          box_diff(piece, is, next);
        }
        todo.swap(next);
        if (todo.empty()) break;
      }
    }
  }
}

static CopyDesc* upload(const std::vector<CopyDesc>& v) {
  if (v.empty()) return nullptr;
  CopyDesc* d = nullptr;
  cudaMalloc(&d, v.size() * sizeof(CopyDesc));
  cudaMemcpy(d, v.data(), v.size() * sizeof(CopyDesc), cudaMemcpyHostToDevice);
  return d;
}

FBPlan& Level::plan(int ixtype, int ng, int skip) {
  auto key = std::make_pair(ixtype, ng + 1024 * skip);
  auto it = plans.find(key);
  if (it != plans.end()) return *it->second;
  auto P = std::make_unique<FBPlan>();
  P->ixtype = ixtype; P->ng = ng;
  std::vector<int> db, sb, sh; std::vector<Bx> rg;
  build_fb_regions(*this, ixtype, ng, db, sb, rg, sh, skip);
  if (skip) {  // clip every region to the valid range of the skipped directions; drop what is left empty
    std::vector<int> db2, sb2, sh2; std::vector<Bx> rg2;
    for (size_t r = 0; r < rg.size(); ++r) {
      const Bx v = ixbox(boxes[db[r]], ixtype);
      Bx c = rg[r];
      for (int d = 0; d < 3; ++d)
        if (skip & (1 << d)) { c.lo[d] = std::max(c.lo[d], v.lo[d]); c.hi[d] = std::min(c.hi[d], v.hi[d]); }
      if (!c.ok()) continue;
      db2.push_back(db[r]); sb2.push_back(sb[r]); rg2.push_back(c);
      for (int q = 0; q < 3; ++q) sh2.push_back(sh[3 * r + q]);
    }
    db.swap(db2); sb.swap(sb2); rg.swap(rg2); sh.swap(sh2);
  }
  const int me = comm().rank;
  std::vector<int> g2l(boxes.size(), -1);
  for (int il = 0; il < nlocal(); ++il) g2l[local[il]] = il;
  std::map<int, std::vector<CopyDesc>> sendm, recvm;
  std::map<int, int64_t> soff, roff;
  for (size_t r = 0; r < rg.size(); ++r) {
    const int od = owner[db[r]], os = owner[sb[r]];
    if (od != me && os != me) continue;
    CopyDesc d{};
    for (int q = 0; q < 3; ++q) { d.lo[q] = rg[r].lo[q]; d.hi[q] = rg[r].hi[q]; d.sh[q] = sh[3 * r + q]; }
    const int64_t npts = rg[r].npts();
    if (od == me && os == me) {
      d.kind = 0; d.dst = g2l[db[r]]; d.src = g2l[sb[r]];
      P->local.push_back(d);
    } else if (os == me) {  // I send: region expressed in src index space
      d.kind = 1; d.src = g2l[sb[r]]; d.dst = -1;
      for (int q = 0; q < 3; ++q) { d.lo[q] += d.sh[q]; d.hi[q] += d.sh[q]; d.sh[q] = 0; }
      d.bufoff = soff[od]; soff[od] += npts;
      sendm[od].push_back(d);
    } else {  // I receive
      d.kind = 2; d.dst = g2l[db[r]]; d.src = -1;
      d.bufoff = roff[os]; roff[os] += npts;
      recvm[os].push_back(d);
    }
  }
  // peers: union of send/recv ranks, ascending
  std::vector<int> peers;
  for (auto& kv : sendm) peers.push_back(kv.first);
  for (auto& kv : recvm) peers.push_back(kv.first);
  std::sort(peers.begin(), peers.end());
  peers.erase(std::unique(peers.begin(), peers.end()), peers.end());
  P->peers = peers;
  std::vector<CopyDesc> all_send, all_recv;
  int64_t so = 0, ro = 0;
  for (int p : peers) {
    P->send_off.push_back(so); P->recv_off.push_back(ro);
    P->send_pts.push_back(soff.count(p) ? soff[p] : 0);
    P->recv_pts.push_back(roff.count(p) ? roff[p] : 0);
    for (CopyDesc d : sendm[p]) { d.bufoff += so; all_send.push_back(d); }
    for (CopyDesc d : recvm[p]) { d.bufoff += ro; all_recv.push_back(d); }
    so += P->send_pts.back(); ro += P->recv_pts.back();
  }
  P->send_total = so; P->recv_total = ro;
  for (int p : peers) { P->send.push_back(sendm[p]); P->recv.push_back(recvm[p]); }
  {
    static int direct_on = -1;
    if (direct_on < 0) { const char* e = getenv("IAMRX_FB_DIRECT"); direct_on = (e && e[0] == '0') ? 0 : 1; }
    bool ok = direct_on && (skip & 3) == 3 && !peers.empty();
    for (size_t r = 0; r < rg.size() && ok; ++r) {
      const int od = owner[db[r]], os = owner[sb[r]];
      if ((od != me && os != me) || (od == me && os == me)) continue;
      const Bx vd = ixbox(boxes[db[r]], ixtype), vs = ixbox(boxes[sb[r]], ixtype);
      for (int d = 0; d < 2; ++d)
        if (sh[3 * r + d] != 0 || rg[r].lo[d] != vd.lo[d] || rg[r].hi[d] != vd.hi[d] || vs.lo[d] != vd.lo[d] || vs.hi[d] != vd.hi[d]) ok = false;
    }
    P->direct = ok;
  }
  for (const CopyDesc& d : P->local) P->max_local = std::max<int64_t>(P->max_local, (int64_t)(d.hi[0] - d.lo[0] + 1) * (d.hi[1] - d.lo[1] + 1) * (d.hi[2] - d.lo[2] + 1));
  for (const CopyDesc& d : all_send) P->max_send = std::max<int64_t>(P->max_send, (int64_t)(d.hi[0] - d.lo[0] + 1) * (d.hi[1] - d.lo[1] + 1) * (d.hi[2] - d.lo[2] + 1));
  for (const CopyDesc& d : all_recv) P->max_recv = std::max<int64_t>(P->max_recv, (int64_t)(d.hi[0] - d.lo[0] + 1) * (d.hi[1] - d.lo[1] + 1) * (d.hi[2] - d.lo[2] + 1));
  P->d_local = upload(P->local);
  P->d_send = upload(all_send); P->n_send = (int)all_send.size();
  P->d_recv = upload(all_recv); P->n_recv = (int)all_recv.size();
  FBPlan& ref = *P;
  plans[key] = std::move(P);
  return ref;
}

// ---------------------------------------------------------------------------
// MF
// ---------------------------------------------------------------------------
MF& MF::operator=(MF&& o) noexcept {
  if (this != &o) {
    clear();
    lev = o.lev; ixtype = o.ixtype; ncomp = o.ncomp; ng = o.ng;
    fabs = std::move(o.fabs); owned = std::move(o.owned);
    o.lev = nullptr; o.fabs.clear(); o.owned.clear();
  }
  return *this;
}

void MF::define(Level* L, int ixtype_, int ncomp_, int ng_) {
  clear();
  lev = L; ixtype = ixtype_; ncomp = ncomp_; ng = ng_;
  fabs.resize(L->nlocal());
  owned.resize(L->nlocal(), nullptr);
  for (int il = 0; il < L->nlocal(); ++il) {
    const Bx g = grow(ixbox(L->lbox(il), ixtype), ng);
    // x rows padded so the first VALID cell of every row sits on a 128-byte line
    const int pad = (16 - (ng % 16)) % 16;
    const int64_t js = ((int64_t)(pad + g.nx()) + 15) / 16 * 16;
    const int64_t ks = js * g.ny();
    const int64_t ns = ks * g.nz();
    double* base = dev_alloc((size_t)(ns * ncomp + 32));
    owned[il] = base;
    iamrx_fab& f = fabs[il];
    // dev_alloc blocks come from cudaMalloc (256-byte aligned)
    f.p = base ? base + pad : nullptr;
    for (int d = 0; d < 3; ++d) { f.lo[d] = g.lo[d]; f.hi[d] = g.hi[d]; }
    f.jstride = js; f.kstride = ks; f.nstride = ns; f.ncomp = ncomp; f.pad_ = 0;
  }
}

void MF::alias(Level* L, int ixtype_, int ncomp_, int ng_, const iamrx_fab* f) {
  clear();
  lev = L; ixtype = ixtype_; ncomp = ncomp_; ng = ng_;
  fabs.assign(f, f + L->nlocal());
  owned.assign(L->nlocal(), nullptr);
}

void MF::clear() {
  for (double* p : owned) if (p) dev_free(p);
  owned.clear(); fabs.clear(); lev = nullptr;
}

static void fill_table(FabTable& t, const MF& m, int comp) {
  for (int il = 0; il < m.n() && il < FabTable::MAXF; ++il) {
    const iamrx_fab& f = m.fabs[il];
    t.p[il] = f.p + (int64_t)comp * f.nstride;
    for (int d = 0; d < 3; ++d) t.lo[il][d] = f.lo[d];
    t.js[il] = f.jstride; t.ks[il] = f.kstride; t.ns[il] = f.nstride;
  }
}

#define IX_TRY(call) do { int rc_ = (call); if (rc_ != IAMRX_OK) return rc_; } while (0)

namespace {
struct DevBuf {   // pool block released on every exit path (stream-ordered reuse: one stream per rank)
  double* p;
  explicit DevBuf(size_t n) : p(dev_alloc(n)) {}
  ~DevBuf() { dev_free(p); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};
}  // namespace

int mf_setval(MF& m, double v, int comp, int ncomp, int ng, cudaStream_t s) {
  for (int il = 0; il < m.n(); ++il) IX_TRY(k::setval(m.gbox(il, ng), m.v(il, comp), ncomp, v, s));
  return IAMRX_OK;
}
int mf_copy(MF& dst, const MF& src, int scomp, int dcomp, int ncomp, int ng, cudaStream_t s) {
  for (int il = 0; il < dst.n(); ++il)
    IX_TRY(k::copy(dst.gbox(il, ng), dst.v(il, dcomp), src.c(il, scomp), ncomp, s));
  return IAMRX_OK;
}
int mf_lincomb(MF& dst, int dcomp, double a, const MF& x, int xcomp, double b, const MF& y, int ycomp,
               int ncomp, int ng, cudaStream_t s) {
  for (int il = 0; il < dst.n(); ++il)
    IX_TRY(k::lincomb(dst.gbox(il, ng), dst.v(il, dcomp), a, x.c(il, xcomp), b, y.c(il, ycomp), ncomp, s));
  return IAMRX_OK;
}
int mf_scale(MF& m, double c, int comp, int ncomp, int ng, cudaStream_t s) {
  for (int il = 0; il < m.n(); ++il) IX_TRY(k::scale(m.gbox(il, ng), m.v(il, comp), c, ncomp, s));
  return IAMRX_OK;
}

std::atomic<int64_t> g_fb_stats[4];   // fills: in-place exchanges, packed exchanges, local-copy launches, local-only fills

int mf_fill_boundary(MF& m, int comp, int ncomp, int ng, cudaStream_t s, int skip) {
  if (ng <= 0 || m.n() == 0 || skip == 7) return IAMRX_OK;
  if (m.n() > FabTable::MAXF) { set_error("mf_fill_boundary: too many local boxes"); return IAMRX_ERR_ARG; }
  Level& L = *m.lev;
  FBPlan& P = L.plan(m.ixtype, ng, skip);
  FabTable t; fill_table(t, m, comp);
  if (P.peers.empty()) {
    if (P.local.empty()) return IAMRX_OK;
    g_fb_stats[3]++;
    return k::copy_batch(P.d_local, (int)P.local.size(), P.max_local, t, t, nullptr, ncomp, 0, s);
  }
  bool owned = true;
  for (double* o : m.owned) if (!o) owned = false;
  if (P.direct && owned && (int)m.owned.size() == m.n()) {
    // whole z planes, identical layout on both sides: exchange them in place
    std::vector<int> pl; std::vector<double*> sb, rb; std::vector<int64_t> sc, rc;
    for (size_t i = 0; i < P.peers.size(); ++i) {
      const size_t ns_ = P.send[i].size() * ncomp, nr_ = P.recv[i].size() * ncomp, ne = std::max(ns_, nr_);
      for (size_t e = 0; e < ne; ++e) {
        pl.push_back(P.peers[i]);
        if (e < ns_) {
          const CopyDesc& d = P.send[i][e / ncomp]; const iamrx_fab& f = m.fabs[d.src];
          sb.push_back(f.p + (int64_t)(comp + (int)(e % ncomp)) * f.nstride + (int64_t)(d.lo[2] - f.lo[2]) * f.kstride - (f.p - m.owned[d.src]));
          sc.push_back((int64_t)(d.hi[2] - d.lo[2] + 1) * f.kstride);
        } else { sb.push_back(nullptr); sc.push_back(0); }
        if (e < nr_) {
          const CopyDesc& d = P.recv[i][e / ncomp]; const iamrx_fab& f = m.fabs[d.dst];
          rb.push_back(f.p + (int64_t)(comp + (int)(e % ncomp)) * f.nstride + (int64_t)(d.lo[2] - f.lo[2]) * f.kstride - (f.p - m.owned[d.dst]));
          rc.push_back((int64_t)(d.hi[2] - d.lo[2] + 1) * f.kstride);
        } else { rb.push_back(nullptr); rc.push_back(0); }
      }
    }
    IX_TRY(comm_exchange(pl, sb, sc, rb, rc, s));
    g_fb_stats[0]++;
    if (!P.local.empty()) { g_fb_stats[2]++; IX_TRY(k::copy_batch(P.d_local, (int)P.local.size(), P.max_local, t, t, nullptr, ncomp, 0, s)); }
    return IAMRX_OK;
  }
  g_fb_stats[1]++;
  { static const char* dbg = getenv("IAMRX_DEBUG_FB"); if (dbg) fprintf(stderr, "[iamrx] packed fill: ixtype %d ng %d skip %d ncomp %d owned %d direct %d\n", m.ixtype, ng, skip, ncomp, (int)owned, (int)P.direct); }
  DevBuf sb_((size_t)(P.send_total * ncomp + 1)), rb_((size_t)(P.recv_total * ncomp + 1));
  double* sbuf = sb_.p; double* rbuf = rb_.p;
  if (!sbuf || !rbuf) return IAMRX_ERR_CUDA;
  IX_TRY(k::copy_batch(P.d_send, P.n_send, P.max_send, t, t, sbuf, ncomp, 0, s));
  std::vector<double*> sb, rb; std::vector<int64_t> sc, rc;
  for (size_t i = 0; i < P.peers.size(); ++i) {
    sb.push_back(sbuf + P.send_off[i] * ncomp); rb.push_back(rbuf + P.recv_off[i] * ncomp);
    sc.push_back(P.send_pts[i] * ncomp); rc.push_back(P.recv_pts[i] * ncomp);
  }
  IX_TRY(comm_exchange(P.peers, sb, sc, rb, rc, s));
  if (!P.local.empty()) IX_TRY(k::copy_batch(P.d_local, (int)P.local.size(), P.max_local, t, t, nullptr, ncomp, 0, s));
  IX_TRY(k::copy_batch(P.d_recv, P.n_recv, P.max_recv, t, t, rbuf, ncomp, 0, s));
  return IAMRX_OK;
}

// physical-boundary part of a FillPatch: cells of every local fab outside the non-periodic sides of the domain
int mf_fill_physbc(MF& m, int comp, int ncomp, int ng, const k::PhysBC& bc, cudaStream_t s) {
  const Level& L = *m.lev;
  if (ng <= 0 || (L.geom.periodic[0] && L.geom.periodic[1] && L.geom.periodic[2])) return IAMRX_OK;
  if (m.ixtype != IX_CELL) { set_error("mf_fill_physbc: cell-centred data only"); return IAMRX_ERR_ARG; }
  for (int il = 0; il < m.n(); ++il)
    IX_TRY(k::fill_physbc(m.gbox(il, ng), m.v(il, comp), ncomp, bc, L.domain, L.geom.periodic, s));
  return IAMRX_OK;
}

// ---- gather into a replicated level ---------------------------------------------------
GatherPlan::~GatherPlan() {
  if (d_pack) cudaFree(d_pack);
  if (d_local) cudaFree(d_local);
  if (d_unpack) cudaFree(d_unpack);
}

GatherPlan& Level::gather_plan(int ixtype) {
  auto it = gplans.find(ixtype);
  if (it != gplans.end()) return *it->second;
  auto P = std::make_unique<GatherPlan>();
  const int me = comm().rank;
  std::vector<int> g2l(boxes.size(), -1);
  for (int il = 0; il < nlocal(); ++il) g2l[local[il]] = il;
  std::map<int, int64_t> roff;                 // points received so far per peer
  std::map<int, std::vector<CopyDesc>> unp;
  for (size_t b = 0; b < boxes.size(); ++b) {  // global box order on every rank: sender and receiver agree on the layout
    const Bx v = ixbox(boxes[b], ixtype);
    CopyDesc d{};
    for (int q = 0; q < 3; ++q) { d.lo[q] = v.lo[q]; d.hi[q] = v.hi[q]; d.sh[q] = 0; }
    const int64_t npts = v.npts();
    if (owner[b] == me) {
      d.kind = 0; d.dst = 0; d.src = g2l[b]; P->local.push_back(d);
      d.kind = 1; d.dst = -1; d.bufoff = P->send_pts; P->send_pts += npts; P->pack.push_back(d);
      P->max_local = std::max(P->max_local, npts);
    } else {
      d.kind = 2; d.dst = 0; d.src = -1; d.bufoff = roff[owner[b]]; roff[owner[b]] += npts;
      unp[owner[b]].push_back(d);
    }
  }
  P->max_pack = P->max_local;
  int64_t ro = 0;
  for (int p = 0; p < comm().nranks; ++p) {   // every other rank is a peer (it needs this rank's boxes even if it owns none)
    if (p == me) continue;
    P->peers.push_back(p);
    P->recv_off.push_back(ro); P->recv_pts.push_back(roff.count(p) ? roff[p] : 0);
    if (!unp.count(p)) continue;
    for (CopyDesc d : unp[p]) {
      d.bufoff += ro; P->unpack.push_back(d);
      P->max_unpack = std::max<int64_t>(P->max_unpack, (int64_t)(d.hi[0] - d.lo[0] + 1) * (d.hi[1] - d.lo[1] + 1) * (d.hi[2] - d.lo[2] + 1));
    }
    ro += roff[p];
  }
  P->recv_total = ro;
  P->d_pack = upload(P->pack); P->d_local = upload(P->local); P->d_unpack = upload(P->unpack);
  GatherPlan& ref = *P;
  gplans[ixtype] = std::move(P);
  return ref;
}

int mf_gather_replicate(MF& dst, const MF& src, int ncomp, cudaStream_t s) {
  if (dst.n() != 1 || src.n() > FabTable::MAXF) { set_error("mf_gather_replicate: bad layout"); return IAMRX_ERR_ARG; }
  GatherPlan& P = src.lev->gather_plan(src.ixtype);
  FabTable td; fill_table(td, dst, 0);
  FabTable ts; fill_table(ts, src, 0);
  if (!P.local.empty()) IX_TRY(k::copy_batch(P.d_local, (int)P.local.size(), P.max_local, td, ts, nullptr, ncomp, 0, s));
  if (P.peers.empty()) return IAMRX_OK;
  DevBuf sb_((size_t)(P.send_pts * ncomp + 1)), rb_((size_t)(P.recv_total * ncomp + 1));
  double* sbuf = sb_.p; double* rbuf = rb_.p;
  if (!sbuf || !rbuf) return IAMRX_ERR_CUDA;
  if (!P.pack.empty()) IX_TRY(k::copy_batch(P.d_pack, (int)P.pack.size(), P.max_pack, td, ts, sbuf, ncomp, 0, s));
  std::vector<double*> sb, rb; std::vector<int64_t> sc, rc;
  for (size_t i = 0; i < P.peers.size(); ++i) {
    sb.push_back(sbuf); sc.push_back(P.send_pts * ncomp);
    rb.push_back(rbuf + P.recv_off[i] * ncomp); rc.push_back(P.recv_pts[i] * ncomp);
  }
  IX_TRY(comm_exchange(P.peers, sb, sc, rb, rc, s));
  if (!P.unpack.empty()) IX_TRY(k::copy_batch(P.d_unpack, (int)P.unpack.size(), P.max_unpack, td, ts, rbuf, ncomp, 0, s));
  return IAMRX_OK;
}

// ---- reductions ---------------------------------------------------------
namespace {
struct RedScratch {
  double* d = nullptr;
  double* h = nullptr;
  RedScratch() {}
  int init() {
    if (d) return IAMRX_OK;
    if (cudaMalloc(&d, 64 * sizeof(double)) != cudaSuccess) return IAMRX_ERR_CUDA;
    if (cudaMallocHost(&h, 64 * sizeof(double)) != cudaSuccess) return IAMRX_ERR_CUDA;
    return IAMRX_OK;
  }
};
RedScratch& red() { static RedScratch r; return r; }
}  // namespace

static Bx unique_box(const MF& m, int il) {
  // points of box il not shared with a higher-index neighbour: drop the upper
  // node/face layer unless it lies on a non-periodic domain boundary.
  Bx b = m.vbox(il);
  const Level& L = *m.lev;
  for (int d = 0; d < 3; ++d) {
    const bool nodal_d = (m.ixtype == IX_NODE) || (m.ixtype == IX_XFACE + d);
    if (!nodal_d) continue;
    const bool at_dom_hi = (L.lbox(il).hi[d] == L.geom.domain.hi[d]);
    if (!(at_dom_hi && !L.geom.periodic[d])) b.hi[d] -= 1;
  }
  return b;
}

static int reduce_common(const MF& m, int comp, int ncomp, int op, double* out, cudaStream_t s,
                         bool uniq) {
  RedScratch& R = red();
  IX_TRY(R.init());
  IX_TRY(k::reduce_init(R.d, ncomp, op, s));
  for (int il = 0; il < m.n(); ++il)
    IX_TRY(k::reduce(uniq ? unique_box(m, il) : m.vbox(il), m.c(il, comp), ncomp, op, R.d, s));
  if (!m.lev->replicated) IX_TRY(comm_allreduce(R.d, ncomp, op, s));   // a replicated level holds the same data on every rank
  IX_CUDA(cudaMemcpyAsync(R.h, R.d, ncomp * sizeof(double), cudaMemcpyDeviceToHost, s));
  IX_CUDA(cudaStreamSynchronize(s));
  for (int n = 0; n < ncomp; ++n) out[n] = R.h[n];
  return IAMRX_OK;
}

// max-norm accumulated by a producer kernel (abec_apply / nodal_adotx with norm_dev): begin hands out a zeroed device scalar,
// end reduces it over the ranks and brings it to the host
int norm_acc_begin(double** dev, cudaStream_t s) {
  RedScratch& R = red();
  IX_TRY(R.init());
  IX_TRY(k::reduce_init(R.d + 32, 1, 2, s));
  *dev = R.d + 32;
  return IAMRX_OK;
}
int norm_acc_end(double* dev, bool replicated, double* out, cudaStream_t s) {
  RedScratch& R = red();
  if (!replicated) IX_TRY(comm_allreduce(dev, 1, 2, s));
  IX_CUDA(cudaMemcpyAsync(R.h + 32, dev, sizeof(double), cudaMemcpyDeviceToHost, s));
  IX_CUDA(cudaStreamSynchronize(s));
  *out = R.h[32];
  return IAMRX_OK;
}

int mf_norminf_each(const MF& m, int comp, int ncomp, double* out, cudaStream_t s) {
  return reduce_common(m, comp, ncomp, 2, out, s, false);
}
int mf_norminf(const MF& m, int comp, int ncomp, double* out, cudaStream_t s) {
  double tmp[16];
  IX_TRY(reduce_common(m, comp, ncomp, 2, tmp, s, false));
  double r = 0;
  for (int n = 0; n < ncomp; ++n) r = (tmp[n] != tmp[n] || r != r) ? std::nan("") : std::max(r, tmp[n]);   // a NaN component is a NaN norm
  *out = r;
  return IAMRX_OK;
}
int mf_sum(const MF& m, int comp, double* out, cudaStream_t s, bool unique_nodes) {
  return reduce_common(m, comp, 1, 0, out, s, unique_nodes);
}
int mf_min(const MF& m, int comp, double* out, cudaStream_t s) {
  return reduce_common(m, comp, 1, 1, out, s, false);
}

}  // namespace ix
