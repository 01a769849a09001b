// api.cu -- context, error reporting and driver entry points of libmml_b200.so.
#include <stdlib.h>
#include <string.h>

#include "mml_ctx.h"

static char g_create_error[512] = "";

int mml_set_error(mml_ctx* ctx, int code, const char* fmt, ...) {
  char* dst = ctx ? ctx->err : g_create_error;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

extern "C" {

int mml_version(void) { return 200; }

int mml_bn_stat_slots(int C) {
  if (C < 1) return 0;
  const int s = 1024 / C;
  return s < 2 ? 2 : (s > 16 ? 16 : s);
}

const char* mml_last_error(const mml_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int mml_ctx_sm_count(const mml_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int64_t mml_ctx_launch_count(const mml_ctx* ctx) { return ctx ? ctx->launches : 0; }
int mml_ctx_set_sm_budget(mml_ctx* ctx, int sms) {
  if (!ctx) return MML_ERR_INVALID;
  ctx->sm_budget = sms <= 0 || sms > ctx->sm_count ? 0 : (sms < 8 ? 8 : sms);
  return MML_OK;
}

int mml_ctx_set_pdl(mml_ctx* ctx, int enable) {
  if (!ctx) return MML_ERR_INVALID;
  ctx->pdl = enable ? 1 : 0;
  return MML_OK;
}

int mml_ctx_create(int device, mml_ctx** out) {
  if (!out) return mml_set_error(nullptr, MML_ERR_INVALID, "mml_ctx_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return mml_set_error(nullptr, MML_ERR_CUDA, "mml_ctx_create: no CUDA device (%s)", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return mml_set_error(nullptr, MML_ERR_INVALID, "mml_ctx_create: bad device %d", device);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return mml_set_error(nullptr, MML_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return mml_set_error(nullptr, MML_ERR_UNSUPPORTED,
                         "mml_b200 needs a Blackwell sm_100 device (tcgen05/TMEM/TMA); device %d is sm_%d%d and there is no fallback",
                         device, prop.major, prop.minor);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return mml_set_error(nullptr, MML_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaFree(0);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return mml_set_error(nullptr, MML_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  mml_ctx* ctx = new mml_ctx();
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->encode_tiled = (mml_tmap_encode_tiled_fn)fn;
  const char* pdl = getenv("MML_PDL");
  ctx->pdl = (pdl != nullptr && pdl[0] == '0') ? 0 : 1;
  *out = ctx;
  return MML_OK;
}

void mml_ctx_destroy(mml_ctx* ctx) { delete ctx; }

}  // extern "C"
