// common.h -- shared device/host plumbing of libiamrx: Array4-style views,
// error handling, launch accounting.  No reference code here; the view mirrors
// the {p, begin, strides} semantics AMReX's Array4 has.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <atomic>
#if defined(IX_EMUL)
// tests-only host emulation of the CUDA runtime (tests/emul/cuda_emul.h); never
// defined when building the product library.
#include "cuda_emul.h"
#else
#include <cuda_runtime.h>
namespace ix { struct KernelTimer { int slot; void* s; KernelTimer(const char* name, void* stream); ~KernelTimer(); }; }
// every launch goes through here; KernelTimer is a no-op unless iamrx_prof_all(1) was called
#define IX_LAUNCH(kern, grid, block, smem, stream, ...)                      \
  do {                                                                       \
    ix::KernelTimer kt_(#kern, (void*)(stream));                             \
    kern<<<grid, block, smem, stream>>>(__VA_ARGS__);                        \
  } while (0)
// large by-value kernel parameter (the host emulation build passes it by reference)
#define IX_KARG(T) T
#endif
#include "../../include/iamrx.h"

#if defined(__CUDACC__)
#define IX_HD __host__ __device__ __forceinline__
#define IX_D __device__ __forceinline__
#else
#define IX_HD inline
#define IX_D inline
#endif

namespace ix {

struct V4 {  // mutable view
  double* p;
  int l0, l1, l2;
  int64_t js, ks, ns;
  IX_HD double& operator()(int i, int j, int k) const {
    return p[(i - l0) + (j - l1) * js + (k - l2) * ks];
  }
  IX_HD double& operator()(int i, int j, int k, int n) const {
    return p[(i - l0) + (j - l1) * js + (k - l2) * ks + n * ns];
  }
  IX_HD bool ok() const { return p != nullptr; }
};

struct C4 {  // read-only view (plain coherent loads: several kernels alias in/out)
  const double* p;
  int l0, l1, l2;
  int64_t js, ks, ns;
  IX_HD double operator()(int i, int j, int k) const {
    return p[(i - l0) + (j - l1) * js + (k - l2) * ks];
  }
  IX_HD double operator()(int i, int j, int k, int n) const {
    return p[(i - l0) + (j - l1) * js + (k - l2) * ks + n * ns];
  }
  IX_HD bool ok() const { return p != nullptr; }
};

inline V4 view(const iamrx_fab* f, int comp = 0) {
  V4 v{};
  if (!f || !f->p) return v;
  v.p = f->p + (int64_t)comp * f->nstride;
  v.l0 = f->lo[0]; v.l1 = f->lo[1]; v.l2 = f->lo[2];
  v.js = f->jstride; v.ks = f->kstride; v.ns = f->nstride;
  return v;
}
inline C4 cview(const iamrx_fab* f, int comp = 0) {
  C4 v{};
  if (!f || !f->p) return v;
  v.p = f->p + (int64_t)comp * f->nstride;
  v.l0 = f->lo[0]; v.l1 = f->lo[1]; v.l2 = f->lo[2];
  v.js = f->jstride; v.ks = f->kstride; v.ns = f->nstride;
  return v;
}

struct Bx {
  int lo[3], hi[3];
  IX_HD int nx() const { return hi[0] - lo[0] + 1; }
  IX_HD int ny() const { return hi[1] - lo[1] + 1; }
  IX_HD int nz() const { return hi[2] - lo[2] + 1; }
  IX_HD int64_t npts() const { return (int64_t)nx() * ny() * nz(); }
  IX_HD bool ok() const { return hi[0] >= lo[0] && hi[1] >= lo[1] && hi[2] >= lo[2]; }
  IX_HD bool contains(int i, int j, int k) const {
    return i >= lo[0] && i <= hi[0] && j >= lo[1] && j <= hi[1] && k >= lo[2] && k <= hi[2];
  }
};
inline Bx mkbx(const iamrx_box& b) {
  Bx r; for (int d = 0; d < 3; ++d) { r.lo[d] = b.lo[d]; r.hi[d] = b.hi[d]; } return r;
}
inline Bx grow(Bx b, int n) { for (int d = 0; d < 3; ++d) { b.lo[d] -= n; b.hi[d] += n; } return b; }
inline Bx grow(Bx b, int d, int n) { b.lo[d] -= n; b.hi[d] += n; return b; }
inline Bx surrounding(Bx b, int d) { b.hi[d] += 1; return b; }  // cell box -> face-d box
inline Bx nodes(Bx b) { for (int d = 0; d < 3; ++d) b.hi[d] += 1; return b; }
inline Bx intersect(const Bx& a, const Bx& b) {
  Bx r; for (int d = 0; d < 3; ++d) { r.lo[d] = a.lo[d] > b.lo[d] ? a.lo[d] : b.lo[d]; r.hi[d] = a.hi[d] < b.hi[d] ? a.hi[d] : b.hi[d]; } return r;
}
inline Bx shift(Bx b, int d, int n) { b.lo[d] += n; b.hi[d] += n; return b; }

// ---- errors / accounting -------------------------------------------------
void set_error(const std::string& s);
extern std::atomic<int64_t> g_launches;
extern std::atomic<int64_t> g_fb_stats[4];   // ghost-fill counters (iamrx_debug_fb_stats)
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool device_ok();

#define IX_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      ix::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));      \
      return IAMRX_ERR_CUDA;                                                  \
    }                                                                         \
  } while (0)

#define IX_NEED_DEVICE()                                                      \
  do {                                                                        \
    if (!ix::device_ok()) {                                                   \
      ix::set_error("no CUDA device: libiamrx has no CPU fallback");          \
      return IAMRX_ERR_NO_DEVICE;                                             \
    }                                                                         \
  } while (0)

#define IX_ARG(cond, msg)                                                     \
  do {                                                                        \
    if (!(cond)) { ix::set_error(std::string("bad argument: ") + msg); return IAMRX_ERR_ARG; } \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return IAMRX_ERR_CUDA;
  }
  count_launch();
  return IAMRX_OK;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- optional per-launch timing (iamrx_prof_*, see include/iamrx.h) ---------
struct ProfScope {
  int slot = -1;
  cudaStream_t s;
  ProfScope(int kclass, int64_t points, double algo_bytes, cudaStream_t s);
  ~ProfScope();
};

}  // namespace ix
