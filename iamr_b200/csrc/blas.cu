// blas.cu -- level BLAS-1, region copies (the FillBoundary data movers) and
// reductions.  Stand-ins for amrex::MultiFab::{setVal,Copy,Saxpy,Xpay,Multiply,
// Divide,mult,norm0,sum} and the pack/unpack loops inside FabArray::FillBoundary
// (IAMR call sites: NSB.cpp:1387,1429,4408, MacProj.cpp:1126-1127,
// Projection.cpp:273,338, Diffusion.cpp:202).
#include "kernels.h"

namespace ix {
namespace k {
namespace {

constexpr int TX = 128;
constexpr int TY = 2;

// Every thread handles up to ZP (plane, component) slices ZSTRIDE = gridDim.z apart: the loads of all its elements are issued
// before the first store (one element per thread and 65 k tiny CTAs reached about a third of the HBM rate on B200).
constexpr int ZP = 4;
inline dim3 grid_for(const Bx& bx, int nzc) { return dim3(cdiv(bx.nx(), TX), cdiv(bx.ny(), TY), cdiv(nzc, ZP)); }
#define IDX3(bx)                                                     \
  const int nz_ = bx.hi[2] - bx.lo[2] + 1;                            \
  const int j = bx.lo[1] + blockIdx.y * TY + threadIdx.y;             \
  const int i = bx.lo[0] + blockIdx.x * TX + threadIdx.x;             \
  if (j > bx.hi[1] || i > bx.hi[0]) return;                           \
  int kk_[ZP], nn_[ZP], cnt_ = 0;                                     \
  _Pragma("unroll") for (int q_ = 0; q_ < ZP; ++q_) {                 \
    const int zz_ = (int)blockIdx.z + q_ * (int)gridDim.z;            \
    kk_[q_] = bx.lo[2] + zz_ % nz_; nn_[q_] = zz_ / nz_;              \
    if (zz_ < nzc) cnt_ = q_ + 1;                                     \
  }
#define FORZ(q) _Pragma("unroll") for (int q = 0; q < ZP; ++q) if (q < cnt_)
#define K_ kk_[q]
#define N_ nn_[q]

__global__ void setval_kernel(Bx bx, int nzc, V4 d, double v) { IDX3(bx) FORZ(q) d(i, j, K_, N_) = v; }
__global__ void copy_kernel(Bx bx, int nzc, V4 d, C4 s) {
  IDX3(bx) double t[ZP];
  FORZ(q) t[q] = s(i, j, K_, N_);
  FORZ(q) d(i, j, K_, N_) = t[q];
}
__global__ void lincomb_kernel(Bx bx, int nzc, V4 d, double a, C4 x, double b, C4 y) {
  IDX3(bx) double tx[ZP], ty[ZP];
  FORZ(q) { tx[q] = x(i, j, K_, N_); ty[q] = y(i, j, K_, N_); }
  FORZ(q) d(i, j, K_, N_) = a * tx[q] + b * ty[q];
}
__global__ void mult_kernel(Bx bx, int nzc, V4 d, C4 s, int sn) {
  IDX3(bx) double td[ZP], ts[ZP];
  FORZ(q) { td[q] = d(i, j, K_, N_); ts[q] = s(i, j, K_, sn > 1 ? N_ : 0); }
  FORZ(q) d(i, j, K_, N_) = td[q] * ts[q];
}
__global__ void div_kernel(Bx bx, int nzc, V4 d, C4 s, int sn) {
  IDX3(bx) double td[ZP], ts[ZP];
  FORZ(q) { td[q] = d(i, j, K_, N_); ts[q] = s(i, j, K_, sn > 1 ? N_ : 0); }
  FORZ(q) d(i, j, K_, N_) = td[q] / ts[q];
}
__global__ void scale_kernel(Bx bx, int nzc, V4 d, double c) {
  IDX3(bx) double td[ZP];
  FORZ(q) td[q] = d(i, j, K_, N_);
  FORZ(q) d(i, j, K_, N_) = td[q] * c;
}
__global__ void addc_kernel(Bx bx, int nzc, V4 d, double c) {
  IDX3(bx) double td[ZP];
  FORZ(q) td[q] = d(i, j, K_, N_);
  FORZ(q) d(i, j, K_, N_) = td[q] + c;
}
__global__ void copy_shift_kernel(Bx bx, int nzc, V4 d, C4 s, int s0, int s1, int s2) {
  IDX3(bx) double t[ZP];
  FORZ(q) t[q] = s(i + s0, j + s1, K_ + s2, N_);
  FORZ(q) d(i, j, K_, N_) = t[q];
}
__global__ void pack_kernel(Bx bx, int nzc, double* buf, C4 s) {
  IDX3(bx)
  const int64_t nx = bx.hi[0] - bx.lo[0] + 1, ny = bx.hi[1] - bx.lo[1] + 1;
  double t[ZP];
  FORZ(q) t[q] = s(i, j, K_, N_);
  FORZ(q) buf[(i - bx.lo[0]) + nx * ((j - bx.lo[1]) + ny * ((K_ - bx.lo[2]) + (int64_t)nz_ * N_))] = t[q];
}
__global__ void unpack_kernel(Bx bx, int nzc, V4 d, const double* buf) {
  IDX3(bx)
  const int64_t nx = bx.hi[0] - bx.lo[0] + 1, ny = bx.hi[1] - bx.lo[1] + 1;
  double t[ZP];
  FORZ(q) t[q] = buf[(i - bx.lo[0]) + nx * ((j - bx.lo[1]) + ny * ((K_ - bx.lo[2]) + (int64_t)nz_ * N_))];
  FORZ(q) d(i, j, K_, N_) = t[q];
}
#undef K_
#undef N_

// max that PROPAGATES NaN (fmax drops it): a NaN anywhere in the data must reach the norm the solvers test, as a positive quiet
// NaN whose bit pattern also wins the unsigned atomicMax below
IX_HD double nanmax(double a, double b) {
  if (a != a || b != b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return fabs(a != a ? a : b);
#endif
  }
  return a > b ? a : b;
}

#if defined(IX_EMUL)
// tests-only serial stand-ins for the cooperative reduction kernels below
static void reduce_serial(Bx bx, C4 src, int ncomp, int op, double* result) {
  for (int n = 0; n < ncomp; ++n) {
    double acc = result[n];
    for (int k = bx.lo[2]; k <= bx.hi[2]; ++k)
      for (int j = bx.lo[1]; j <= bx.hi[1]; ++j)
        for (int i = bx.lo[0]; i <= bx.hi[0]; ++i) {
          const double v = src(i, j, k, n);
          acc = (op == 0) ? acc + v : (op == 1 ? fmin(acc, v) : nanmax(acc, fabs(v)));
        }
    result[n] = acc;
  }
}
static void dot_serial(Bx bx, C4 x, C4 y, C4 mask, double* result) {
  double acc = 0.0;
  for (int k = bx.lo[2]; k <= bx.hi[2]; ++k)
    for (int j = bx.lo[1]; j <= bx.hi[1]; ++j)
      for (int i = bx.lo[0]; i <= bx.hi[0]; ++i) {
        double v = x(i, j, k) * y(i, j, k);
        if (mask.ok()) v *= mask(i, j, k);
        acc += v;
      }
  *result += acc;
}
#else
// ---- reductions ----------------------------------------------------------
// Grid-stride over (j,k) rows; warp-shuffle + smem block reduce; one atomic per
// CTA.  max/min use the monotone bit pattern of non-negative / ordered doubles
// through atomicMax/atomicMin on (unsigned) long long; sum uses atomicAdd.
IX_D double warp_red(double v, int op) {
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, v, o);
    v = (op == 0) ? v + t : (op == 1 ? fmin(v, t) : nanmax(v, t));
  }
  return v;
}

IX_D void atomic_fmax_nonneg(double* addr, double v) {  // v >= 0
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}
IX_D void atomic_fmin(double* addr, double v) {
  // generic CAS loop (rare path: one per CTA)
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) <= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}

constexpr int RT = 256;

// op == 0 (sum) writes one partial per CTA into `partials` (summed in a fixed order by
// reduce_final_kernel: bit-reproducible run to run); min / max go straight to `result`.
__global__ void __launch_bounds__(RT) reduce_kernel(Bx bx, C4 src, int op, double* result, double* partials) {
  const int n = blockIdx.y;
  const int nx = bx.nx(), ny = bx.ny(), nz = bx.nz();
  const int64_t nrows = (int64_t)ny * nz;
  double acc = (op == 0) ? 0.0 : (op == 1 ? 1.0e300 : 0.0);
  const int lane_x = threadIdx.x & 63;
  const int rsub = threadIdx.x >> 6;  // 4 rows per CTA pass
  for (int64_t r = (int64_t)blockIdx.x * 4 + rsub; r < nrows; r += (int64_t)gridDim.x * 4) {
    const int j = bx.lo[1] + (int)(r % ny);
    const int k = bx.lo[2] + (int)(r / ny);
    const double* row = src.p + ((bx.lo[0] - src.l0) + (j - src.l1) * src.js + (k - src.l2) * src.ks + n * src.ns);
    for (int ii = lane_x; ii < nx; ii += 256) {   // four independent loads in flight per thread
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (ii + 64 * u < nx) ? row[ii + 64 * u] : ((op == 1) ? 1.0e300 : 0.0);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = (op == 0) ? acc + v[u] : (op == 1 ? fmin(acc, v[u]) : nanmax(acc, fabs(v[u])));
    }
  }
  __shared__ double sm[RT / 32];
  acc = warp_red(acc, op);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < RT / 32) ? sm[threadIdx.x] : ((op == 0) ? 0.0 : (op == 1 ? 1.0e300 : 0.0));
    v = warp_red(v, op);
    if (threadIdx.x == 0) {
      if (op == 0) partials[(size_t)n * gridDim.x + blockIdx.x] = v;
      else if (op == 1) atomic_fmin(result + n, v);
      else atomic_fmax_nonneg(result + n, v);
    }
  }
}

__global__ void __launch_bounds__(RT) reduce_final_kernel(const double* partials, int nb, double* result) {
  const int n = blockIdx.x;
  double acc = 0.0;
  for (int b = threadIdx.x; b < nb; b += RT) acc += partials[(size_t)n * nb + b];
  __shared__ double sm[RT];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = RT / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) result[n] += sm[0];
}

__global__ void __launch_bounds__(RT) dot_kernel(Bx bx, C4 x, C4 y, C4 mask, double* result) {
  const int nx = bx.nx(), ny = bx.ny(), nz = bx.nz();
  const int64_t nrows = (int64_t)ny * nz;
  double acc = 0.0;
  const int lane_x = threadIdx.x & 63;
  const int rsub = threadIdx.x >> 6;
  for (int64_t r = (int64_t)blockIdx.x * 4 + rsub; r < nrows; r += (int64_t)gridDim.x * 4) {
    const int j = bx.lo[1] + (int)(r % ny);
    const int k = bx.lo[2] + (int)(r / ny);
    for (int ii = lane_x; ii < nx; ii += 64) {
      const int i = bx.lo[0] + ii;
      double v = x(i, j, k) * y(i, j, k);
      if (mask.ok()) v *= mask(i, j, k);
      acc += v;
    }
  }
  __shared__ double sm[RT / 32];
  acc = warp_red(acc, 0);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < RT / 32) ? sm[threadIdx.x] : 0.0;
    v = warp_red(v, 0);
    if (threadIdx.x == 0) atomicAdd(result, v);
  }
}

// every element of comp 0 equal to the first one?  (constant-coefficient detection of the cell-centred multigrid)
__global__ void __launch_bounds__(RT) const_check_kernel(Bx bx, C4 src, double* result) {
  const int nx = bx.nx(), ny = bx.ny(), nz = bx.nz();
  const int64_t nrows = (int64_t)ny * nz;
  const double ref = src(bx.lo[0], bx.lo[1], bx.lo[2]);
  bool diff = false;
  const int lane_x = threadIdx.x & 63;
  const int rsub = threadIdx.x >> 6;
  for (int64_t r = (int64_t)blockIdx.x * 4 + rsub; r < nrows; r += (int64_t)gridDim.x * 4) {
    const int j = bx.lo[1] + (int)(r % ny);
    const int k = bx.lo[2] + (int)(r / ny);
    for (int ii = lane_x; ii < nx; ii += 64) diff |= (src(bx.lo[0] + ii, j, k) != ref);
  }
  if (__syncthreads_or(diff ? 1 : 0) && threadIdx.x == 0) result[0] = 1.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (!(ref > 0.0)) result[0] = 1.0;
    else { atomic_fmax_nonneg(result + 1, ref); atomic_fmin(result + 2, ref); }
  }
}

#endif  // IX_EMUL

__global__ void reduce_init_kernel(double* r, int n, int op) {
  if ((int)threadIdx.x < n) r[threadIdx.x] = (op == 1) ? 1.0e300 : 0.0;
}

}  // namespace

#define LAUNCH3(kern, bx, ncomp, s, ...)                                                  \
  do {                                                                                    \
    if (!(bx).ok() || (ncomp) <= 0) return IAMRX_OK;                                      \
    IX_LAUNCH(kern, grid_for(bx, (bx).nz() * (ncomp)), dim3(TX, TY, 1), 0, s, bx, (bx).nz() * (ncomp), __VA_ARGS__);  \
    return check_launch(#kern);                                                           \
  } while (0)

int setval(const Bx& bx, V4 dst, int ncomp, double val, cudaStream_t s) { LAUNCH3(setval_kernel, bx, ncomp, s, dst, val); }
int copy(const Bx& bx, V4 dst, C4 src, int ncomp, cudaStream_t s) { LAUNCH3(copy_kernel, bx, ncomp, s, dst, src); }
int lincomb(const Bx& bx, V4 dst, double a, C4 x, double b, C4 y, int ncomp, cudaStream_t s) {
  LAUNCH3(lincomb_kernel, bx, ncomp, s, dst, a, x, b, y);
}
int mult(const Bx& bx, V4 dst, C4 src, int ncomp, int sn, cudaStream_t s) { LAUNCH3(mult_kernel, bx, ncomp, s, dst, src, sn); }
int divide(const Bx& bx, V4 dst, C4 src, int ncomp, int sn, cudaStream_t s) { LAUNCH3(div_kernel, bx, ncomp, s, dst, src, sn); }
int scale(const Bx& bx, V4 dst, double c, int ncomp, cudaStream_t s) { LAUNCH3(scale_kernel, bx, ncomp, s, dst, c); }
int addconst(const Bx& bx, V4 dst, double c, int ncomp, cudaStream_t s) { LAUNCH3(addc_kernel, bx, ncomp, s, dst, c); }
int copy_shift(const Bx& bx, V4 dst, C4 src, int s0, int s1, int s2, int ncomp, cudaStream_t s) {
  LAUNCH3(copy_shift_kernel, bx, ncomp, s, dst, src, s0, s1, s2);
}
int pack(const Bx& bx, double* buf, C4 src, int ncomp, cudaStream_t s) { LAUNCH3(pack_kernel, bx, ncomp, s, buf, src); }
int unpack(const Bx& bx, V4 dst, const double* buf, int ncomp, cudaStream_t s) { LAUNCH3(unpack_kernel, bx, ncomp, s, dst, buf); }

int reduce_init(double* result, int n, int op, cudaStream_t s) {
  IX_LAUNCH(reduce_init_kernel, 1, 32, 0, s, result, n, op);
  return check_launch("reduce_init");
}

int reduce(const Bx& bx, C4 src, int ncomp, int op, double* result, cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  const int64_t nrows = (int64_t)bx.ny() * bx.nz();
  int nb = (int)((nrows + 3) / 4);
  if (nb > 148 * 8) nb = 148 * 8;
  if (nb < 1) nb = 1;
#if defined(IX_EMUL)
  (void)nb; reduce_serial(bx, src, ncomp, op, result);
#else
  static double* partials = nullptr;  // 148*8 CTAs x up to 16 components
  if (!partials && cudaMalloc(&partials, sizeof(double) * 148 * 8 * 16) != cudaSuccess) {
    set_error("reduce: cudaMalloc failed"); return IAMRX_ERR_CUDA;
  }
  if (ncomp > 16) { set_error("reduce: ncomp > 16"); return IAMRX_ERR_ARG; }
  IX_LAUNCH(reduce_kernel, dim3(nb, ncomp, 1), RT, 0, s, bx, src, op, result, partials);
  if (op == 0) {
    int rc = check_launch("reduce");
    if (rc) return rc;
    IX_LAUNCH(reduce_final_kernel, ncomp, RT, 0, s, partials, nb, result);
  }
#endif
  return check_launch("reduce");
}

// const_check: result[0] = 1 if some element of comp 0 over bx differs from the first one (or the value is not positive),
// result[1] = max over the calls of the first element, result[2] = min of it.  The caller initialises result = {0, 0, 1e300};
// after all boxes (and an allreduce) the data is one constant iff result[0] == 0 and result[1] == result[2].
int const_check(const Bx& bx, C4 src, double* result, cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
#if defined(IX_EMUL)
  const double ref = src(bx.lo[0], bx.lo[1], bx.lo[2]);
  bool diff = !(ref > 0.0);
  for (int k = bx.lo[2]; k <= bx.hi[2]; ++k)
    for (int j = bx.lo[1]; j <= bx.hi[1]; ++j)
      for (int i = bx.lo[0]; i <= bx.hi[0]; ++i) if (src(i, j, k) != ref) diff = true;
  if (diff) result[0] = 1.0;
  result[1] = fmax(result[1], ref); result[2] = fmin(result[2], ref);
#else
  const int64_t nrows = (int64_t)bx.ny() * bx.nz();
  int nb = (int)((nrows + 3) / 4);
  if (nb > 148 * 8) nb = 148 * 8;
  IX_LAUNCH(const_check_kernel, nb, RT, 0, s, bx, src, result);
#endif
  return check_launch("const_check");
}

int reduce_dot(const Bx& bx, C4 x, C4 y, C4 mask, double* result, cudaStream_t s) {
  if (!bx.ok()) return IAMRX_OK;
  const int64_t nrows = (int64_t)bx.ny() * bx.nz();
  int nb = (int)((nrows + 3) / 4);
  if (nb > 148 * 8) nb = 148 * 8;
  if (nb < 1) nb = 1;
#if defined(IX_EMUL)
  (void)nb; dot_serial(bx, x, y, mask, result);
#else
  IX_LAUNCH(dot_kernel, nb, RT, 0, s, bx, x, y, mask, result);
#endif
  return check_launch("reduce_dot");
}


#if !defined(IX_EMUL)
// ---- diagnostic: measured FP64 pipe rate (the second roofline of the Godunov kernels; BASELINE.md asks for a measured,
// not a nominal, DFMA rate).  Eight independent DFMA chains per thread, 256 threads, 8 CTAs per SM.
namespace {
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, double a, double b, int iters) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 123.456) out[0] = x0;   // keep the chains alive
}
}  // namespace
int fp64_peak(double* dp_ginstr_per_s, cudaStream_t s) {
  double* d = nullptr;
  IX_CUDA(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  IX_CUDA(cudaEventCreate(&e0)); IX_CUDA(cudaEventCreate(&e1));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int iters = 1 << 15, blocks = sms * 8;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    IX_CUDA(cudaEventRecord(e0, s));
    dfma_peak_kernel<<<blocks, 256, 0, s>>>(d, 0.999999, 1.0e-6, iters);
    IX_CUDA(cudaEventRecord(e1, s));
    IX_CUDA(cudaEventSynchronize(e1));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double rate = (double)blocks * 256.0 * 8.0 * iters / (ms * 1e-3) / 1e9;
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *dp_ginstr_per_s = best;
  return check_launch("dfma_peak");
}
#else
int fp64_peak(double* r, cudaStream_t) { *r = 0.0; return IAMRX_OK; }
#endif

}  // namespace k
}  // namespace ix
