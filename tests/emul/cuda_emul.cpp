// storage for the emulated CUDA built-ins (tests only; see cuda_emul.h)
#include "cuda_emul.h"
thread_local uint3_emul blockIdx, threadIdx;
thread_local dim3 blockDim, gridDim;
