// comm.cu -- the library-owned NCCL communicator of the data-parallel step (mml_comm_init /
// mml_allreduce_bucket).  The reference has no distributed code at all; this is the collective of section 8e: one sum all-reduce of
// a contiguous range of the flat fp32 gradient buffer, enqueued on the caller's stream (capturable into the step's CUDA graph).
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 the process already has -- PyTorch's), so the library neither links
// against nor ships a second NCCL.  Owning the communicator lets the step cap NCCL's CTA budget (ncclConfig_t::maxCTAs): the
// all-reduce of the image-encoder range runs UNDER the audio encoder's backward, and every SM NCCL takes is one the persistent
// convolution kernels lose.
#include <dlfcn.h>
#include <nccl.h>  // types and the ncclConfig_t initialiser only; no symbol of libnccl is linked
#include <stdlib.h>
#include <string.h>

#include "mml_ctx.h"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;

int load_nccl(mml_ctx* ctx) {
  if (g_nccl.lib != nullptr) return MML_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return mml_set_error(ctx, MML_ERR_UNSUPPORTED, "libnccl.so.2 not found (import torch first: its NCCL is the one used): %s", dlerror());
#define MML_SYM(field, name)                                                                          \
  *(void**)(&g_nccl.field) = dlsym(lib, name);                                                         \
  if (!g_nccl.field) return mml_set_error(ctx, MML_ERR_UNSUPPORTED, "libnccl: symbol %s not found", name)
  MML_SYM(GetUniqueId, "ncclGetUniqueId");
  MML_SYM(CommInitRankConfig, "ncclCommInitRankConfig");
  MML_SYM(AllReduce, "ncclAllReduce");
  MML_SYM(CommDestroy, "ncclCommDestroy");
  MML_SYM(GetErrorString, "ncclGetErrorString");
  MML_SYM(GetVersion, "ncclGetVersion");
#undef MML_SYM
  g_nccl.lib = lib;
  return MML_OK;
}

#define MML_CHECK_NCCL(ctx, expr)                                                                                         \
  do {                                                                                                                    \
    ncclResult_t _r = (expr);                                                                                             \
    if (_r != ncclSuccess) return mml_set_error(ctx, MML_ERR_CUDA, "%s failed: %s", #expr, g_nccl.GetErrorString(_r));   \
  } while (0)

}  // namespace

extern "C" {

int mml_comm_unique_id(mml_ctx* ctx, uint8_t* id_out) {
  MML_REQUIRE(ctx, ctx && id_out, "comm_unique_id: null pointer");
  int rc = load_nccl(ctx);
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == MML_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  MML_CHECK_NCCL(ctx, g_nccl.GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return MML_OK;
}

int mml_comm_init(mml_ctx* ctx, const uint8_t* id_in, int rank, int world, int max_ctas) {
  MML_REQUIRE(ctx, ctx && id_in && world >= 1 && rank >= 0 && rank < world, "comm_init: bad arguments");
  MML_REQUIRE(ctx, ctx->comm == nullptr, "comm_init: this context already owns a communicator");
  int rc = load_nccl(ctx);
  if (rc) return rc;
  MML_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id_in, sizeof(id));
  ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
  if (max_ctas > 0) {
    cfg.minCTAs = 1;
    cfg.maxCTAs = max_ctas;
  }
  ncclComm_t comm = nullptr;
  MML_CHECK_NCCL(ctx, g_nccl.CommInitRankConfig(&comm, world, id, rank, &cfg));
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return MML_OK;
}

int mml_comm_world(const mml_ctx* ctx) { return (ctx && ctx->comm) ? ctx->comm_world : 0; }

int mml_allreduce_bucket(mml_ctx* ctx, float* buf, int64_t count, void* stream) {
  MML_REQUIRE(ctx, ctx && buf && count >= 1, "allreduce_bucket: bad arguments");
  MML_REQUIRE(ctx, ctx->comm != nullptr, "allreduce_bucket: no communicator (mml_comm_init)");
  MML_CHECK_NCCL(ctx, g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm, (cudaStream_t)stream));
  ctx->launches++;
  return MML_OK;
}

int mml_comm_destroy(mml_ctx* ctx) {
  if (!ctx || !ctx->comm) return MML_OK;
  ncclResult_t r = g_nccl.CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  return r == ncclSuccess ? MML_OK : mml_set_error(ctx, MML_ERR_CUDA, "ncclCommDestroy failed: %s", g_nccl.GetErrorString(r));
}

}  // extern "C"
