// optim.cu -- flat-buffer optimizer and aggregation kernels (all HBM-bandwidth bound, 128-bit accesses).
//
//   mml_adam_step        torch.optim.Adam(lr, weight_decay) as built by MML_Suite/config/optimizer_config.py:212-236 from
//                        configs/avmnist/centralised/train_avmnist_resnet.yaml:28-32 and stepped at models/avmnist.py:303;
//                        one launch over the flat fp32 parameter / gradient / moment buffers (28 B per parameter) that
//                        also refreshes the bf16 shadow copy the tensor-core kernels read (+2 B).
//   mml_fedavg           theta = sum_k (n_k / sum n) theta_k.  The reference has NO implementation
//                        (MML_Suite/train_congruent_federated.py is empty); textbook FedAvg.
#include "mml_common.cuh"
#include "mml_ctx.h"

using namespace mml;

namespace {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            uint16_t* __restrict__ pb, long long n, const float* __restrict__ hyper, const long long* __restrict__ step) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  __shared__ float sh[2];
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], gs = hyper[5];
  if (threadIdx.x == 0) {
    const double t = (double)(step[0] + 1);
    const double bc1 = 1.0 - pow((double)b1, t);
    const double bc2 = 1.0 - pow((double)b2, t);
    sh[0] = (float)((double)lr / bc1);       // step_size
    sh[1] = (float)(1.0 / sqrt(bc2));        // 1/sqrt(bias_correction2)
  }
  __syncthreads();
  const float step_size = sh[0], rsq_bc2 = sh[1];
  const float omb1 = 1.f - b1, omb2 = 1.f - b2;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    float pp[4] = {P.x, P.y, P.z, P.w}, gg[4] = {G.x, G.y, G.z, G.w}, mm[4] = {M.x, M.y, M.z, M.w}, vv[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = fmaf(wd, pp[j], gg[j] * gs);             // coupled L2: g + wd*p
      mm[j] = fmaf(omb1, gr, b1 * mm[j]);
      vv[j] = fmaf(omb2 * gr, gr, b2 * vv[j]);
      const float denom = fmaf(sqrtf(vv[j]), rsq_bc2, eps);     // sqrt(v)/sqrt(bc2) + eps
      pp[j] = pp[j] - step_size * (mm[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (pb) reinterpret_cast<uint2*>(pb)[i] = make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]));
  }
  // tail (n not a multiple of 4)
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gr = fmaf(wd, p[i], g[i] * gs);
    const float mi = fmaf(omb1, gr, b1 * m[i]);
    const float vi = fmaf(omb2 * gr, gr, b2 * v[i]);
    m[i] = mi, v[i] = vi;
    const float np = p[i] - step_size * (mi / fmaf(sqrtf(vi), rsq_bc2, eps));
    p[i] = np;
    if (pb) pb[i] = (uint16_t)(pack_bf16x2(np, 0.f) & 0xFFFFu);
  }
}

__global__ void step_inc_kernel(long long* step) {
  pdl_sync();
  step[0] += 1;
}

__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = (uint16_t)(pack_bf16x2(src[i], 0.f) & 0xFFFFu);
}

// bf16 -> fp32 (exact): a GEMM output row handed to an fp32 head (MonomodalEncoder around the MMIMDb encoders)
__global__ void __launch_bounds__(256) widen_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, long long n) {
  pdl_sync();
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(src) + i);
    reinterpret_cast<float4*>(dst)[i] = make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __uint_as_float((uint32_t)src[i] << 16);
}

constexpr int kMaxClients = 64;

__global__ void __launch_bounds__(256)
fedavg_kernel(const float* const* __restrict__ clients, const float* __restrict__ weights, int K, float* __restrict__ out, long long n) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  __shared__ const float* ptr[kMaxClients];
  __shared__ float wk[kMaxClients];
  if ((int)threadIdx.x < K) {
    ptr[threadIdx.x] = clients[threadIdx.x];
    wk[threadIdx.x] = weights[threadIdx.x];
  }
  __syncthreads();
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ptr[k]) + i);
      const float w = wk[k];
      acc.x = fmaf(w, v.x, acc.x), acc.y = fmaf(w, v.y, acc.y), acc.z = fmaf(w, v.z, acc.z), acc.w = fmaf(w, v.w, acc.w);
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(wk[k], ptr[k][i], acc);
    out[i] = acc;
  }
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ x, const float* __restrict__ weights, int idx, long long n) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  const float w = weights[idx];
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    v.x *= w, v.y *= w, v.z *= w, v.w *= w;
    reinterpret_cast<float4*>(x)[i] = v;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= w;
}

int flat_grid(const mml_ctx* ctx, long long n) {
  long long b = mml_ceil_div(mml_ceil_div(n, 4), 256);
  const long long cap = (long long)ctx->sm_count * 8;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace

extern "C" {

int mml_adam_step(mml_ctx* ctx, float* p, const float* g, float* m, float* v, uint16_t* p_bf16, int64_t n, const float* hyper,
                  int64_t* step, int advance_step, void* stream) {
  MML_REQUIRE(ctx, ctx && p && g && m && v && hyper && step && n >= 1, "adam_step: bad arguments");
  MML_REQUIRE(ctx, aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (!p_bf16 || ((uintptr_t)p_bf16 & 7u) == 0),
              "adam_step: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  MML_LAUNCH(ctx, adam_kernel, flat_grid(ctx, n), 256, 0, st, p, g, m, v, p_bf16, n, hyper, (const long long*)step);
  if (advance_step) {
    MML_LAUNCH(ctx, step_inc_kernel, 1, 1, 0, st, (long long*)step);
  }
  return MML_OK;
}

int mml_cast_f32_bf16(mml_ctx* ctx, const float* src, uint16_t* dst, int64_t n, void* stream) {
  MML_REQUIRE(ctx, ctx && src && dst && n >= 1, "cast: bad arguments");
  MML_REQUIRE(ctx, aligned16(src) && ((uintptr_t)dst & 7u) == 0, "cast: buffers must be aligned");
  MML_LAUNCH(ctx, cast_kernel, flat_grid(ctx, n), 256, 0, (cudaStream_t)stream, src, dst, n);
  return MML_OK;
}

int mml_cast_bf16_f32(mml_ctx* ctx, const uint16_t* src, float* dst, int64_t n, void* stream) {
  MML_REQUIRE(ctx, ctx && src && dst && n >= 1, "cast: bad arguments");
  MML_REQUIRE(ctx, aligned16(dst) && ((uintptr_t)src & 7u) == 0, "cast: buffers must be aligned");
  MML_LAUNCH(ctx, widen_kernel, flat_grid(ctx, n), 256, 0, (cudaStream_t)stream, src, dst, n);
  return MML_OK;
}

int mml_fedavg(mml_ctx* ctx, const float* const* clients, const float* weights, int K, float* out, int64_t n, void* stream) {
  MML_REQUIRE(ctx, ctx && clients && weights && out && n >= 1, "fedavg: bad arguments");
  MML_REQUIRE(ctx, K >= 1 && K <= kMaxClients, "fedavg: K must be in [1, %d]", kMaxClients);
  MML_REQUIRE(ctx, aligned16(out), "fedavg: out must be 16-byte aligned");
  MML_LAUNCH(ctx, fedavg_kernel, flat_grid(ctx, n), 256, 0, (cudaStream_t)stream, clients, weights, K, out, n);
  return MML_OK;
}

int mml_scale_inplace(mml_ctx* ctx, float* x, const float* weights, int idx, int64_t n, void* stream) {
  MML_REQUIRE(ctx, ctx && x && weights && idx >= 0 && n >= 1 && aligned16(x), "scale_inplace: bad arguments");
  MML_LAUNCH(ctx, scale_kernel, flat_grid(ctx, n), 256, 0, (cudaStream_t)stream, x, weights, idx, n);
  return MML_OK;
}

}  // extern "C"
