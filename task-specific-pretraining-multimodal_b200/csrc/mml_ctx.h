// mml_ctx.h -- per-device context of libmml_b200.so (internal).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mml_b200.h"

typedef CUresult (*mml_tmap_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct mml_ctx {
  int device;
  int sm_count;
  int64_t launches;
  mml_tmap_encode_tiled_fn encode_tiled;
  // (the library owns no device memory: scratch for per-CTA / per-split partials is passed in by the caller, one buffer per
  // stream of weight-gradient launches -- see mml_conv_wgrad_workspace)
  // SMs the persistent kernels may occupy (0 = all): a caller that runs a second stream of small kernels next to them can keep a
  // few SMs free so those kernels never wait for a whole persistent grid to drain (mml_ctx_set_sm_budget)
  int sm_budget;
  // programmatic dependent launch: kernels are enqueued with cudaLaunchAttributeProgrammaticStreamSerialization, so that the next
  // kernel of a stream is scheduled -- and runs its prologue (barrier init, TMEM allocation, tensor-map prefetch) -- while the
  // previous one is still executing; every such kernel calls griddepcontrol.wait before it touches global memory.  MML_PDL=0
  // in the environment turns it off (A/B runs).
  int pdl;
  // library-owned NCCL communicator of the data-parallel step (comm.cu); NULL until mml_comm_init
  void* comm;
  int comm_rank, comm_world;
  char err[512];
};

int mml_set_error(mml_ctx* ctx, int code, const char* fmt, ...);

#define MML_CHECK_CUDA(ctx, expr)                                                                          \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess)                                                                                 \
      return mml_set_error(ctx, MML_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define MML_REQUIRE(ctx, cond, ...)                                        \
  do {                                                                     \
    if (!(cond)) return mml_set_error(ctx, MML_ERR_INVALID, __VA_ARGS__);  \
  } while (0)

// after a <<< >>> launch: count it and surface launch-configuration errors
#define MML_LAUNCHED(ctx)                                                                               \
  do {                                                                                                  \
    (ctx)->launches++;                                                                                  \
    cudaError_t _e = cudaPeekAtLastError();                                                             \
    if (_e != cudaSuccess) {                                                                            \
      cudaGetLastError();                                                                               \
      return mml_set_error(ctx, MML_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                                   \
  } while (0)

static inline int64_t mml_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
#include <utility>
// launch of a kernel that follows the PDL protocol (pdl_wait() before its first global-memory access, then pdl_launch_dependents())
template <typename... KArgs, typename... Args>
static inline cudaError_t mml_launch_kernel(const mml_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                            Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (ctx != nullptr && ctx->pdl) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#define MML_LAUNCH(ctx, kernel, grid, block, smem, st, ...)                                                                   \
  do {                                                                                                                        \
    cudaError_t _le = mml_launch_kernel(ctx, kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__);             \
    (ctx)->launches++;                                                                                                        \
    if (_le != cudaSuccess) {                                                                                                 \
      cudaGetLastError();                                                                                                     \
      return mml_set_error(ctx, MML_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_le), __FILE__, __LINE__); \
    }                                                                                                                         \
  } while (0)
#endif
