// cuda_emul.h -- TEST INFRASTRUCTURE ONLY.  A host stand-in for the handful of
// CUDA runtime entry points and kernel-launch mechanics libiamrx uses, so that
// the host-side logic (FillBoundary plans, multigrid drivers, the time-step
// sequence) and the simple one-thread-per-cell kernels can be exercised by the
// `-m "not gpu"` tests in a container without a GPU.
//
// It is compiled ONLY into tests/emul/_build/libiamrx_emul.so (see
// tests/emul/Makefile) with -DIX_EMUL.  The product library
// iamr_b200/libiamrx.so is built by nvcc without IX_EMUL, contains none of
// this, and has no CPU path.  Nothing under iamr_b200/ loads the emulation
// library; bench.py and __graft_entry__.py never touch it.
#pragma once
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emul { unsigned x, y, z; };

extern thread_local uint3_emul blockIdx, threadIdx;
extern thread_local dim3 blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "success" : "emulated failure"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) {
  size_t nn = (n + 255) / 256 * 256;
  *p = (T*)aligned_alloc(256, nn ? nn : 256);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }

// Kernel "launch": run every (block, thread) of the grid as a plain function
// call.  Only valid for kernels whose threads do not cooperate; kernels that use
// shared memory / shuffles carry an `#ifdef IX_EMUL` serial body instead.
template <class F>
inline void ix_emul_launch(dim3 g, dim3 b, F&& f) {
  const size_t work_ = (size_t)g.x * g.y * g.z * b.x * b.y * b.z;
#pragma omp parallel for collapse(2) schedule(static) if (work_ > 65536)
  for (unsigned bz = 0; bz < g.z; ++bz)
    for (unsigned by = 0; by < g.y; ++by) {
      gridDim = g; blockDim = b;
      for (unsigned bx = 0; bx < g.x; ++bx) {
        blockIdx = {bx, by, bz};
        for (unsigned tz = 0; tz < b.z; ++tz)
          for (unsigned ty = 0; ty < b.y; ++ty)
            for (unsigned tx = 0; tx < b.x; ++tx) {
              threadIdx = {tx, ty, tz};
              f();
            }
      }
    }
}
#define IX_KARG(T) const T&
#define IX_LAUNCH(kern, grid, block, smem, stream, ...) \
  ix_emul_launch(dim3(grid), dim3(block), [&]() { kern(__VA_ARGS__); })
