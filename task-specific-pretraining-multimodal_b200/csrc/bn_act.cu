// bn_act.cu -- HBM-bound fused kernels around the convolutions: BatchNorm statistics finalize, BN-apply + ReLU
// (+ residual) forward, its two-pass backward, max / average pooling, and the missing-modality mask.
//
// Reference ops replaced (MML_Suite/models/msa/networks/resnet.py): nn.BatchNorm2d (:26,31,138,177), nn.ReLU (:27,139),
// residual add (:51), nn.MaxPool2d(3,2,1) (:140), nn.AdaptiveAvgPool2d((1,1)) (:149); data/base_dataset.py:70-72 (mask).
// Activations are NHWC bf16 viewed as a [rows, C] matrix; every thread owns 8 consecutive channels (one 16-byte
// access) and, because 256-thread blocks stride by a multiple of C/8, the SAME 8 channels for its whole life, so the
// per-channel coefficients live in registers.
#include "mml_common.cuh"
#include "mml_ctx.h"

using namespace mml;

namespace {

constexpr int kThreads = 256;

struct F8 {
  float v[8];
};

__device__ __forceinline__ F8 unpack8(const uint4& u) {
  F8 f;
  f.v[0] = bf16_lo(u.x), f.v[1] = bf16_hi(u.x);
  f.v[2] = bf16_lo(u.y), f.v[3] = bf16_hi(u.y);
  f.v[4] = bf16_lo(u.z), f.v[5] = bf16_hi(u.z);
  f.v[6] = bf16_lo(u.w), f.v[7] = bf16_hi(u.w);
  return f;
}
__device__ __forceinline__ uint4 pack8(const F8& f) {
  uint4 u;
  u.x = pack_bf16x2(f.v[0], f.v[1]);
  u.y = pack_bf16x2(f.v[2], f.v[3]);
  u.z = pack_bf16x2(f.v[4], f.v[5]);
  u.w = pack_bf16x2(f.v[6], f.v[7]);
  return u;
}
__device__ __forceinline__ F8 load8f(const float* p) {
  F8 f;
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f.v[0] = a.x, f.v[1] = a.y, f.v[2] = a.z, f.v[3] = a.w;
  f.v[4] = b.x, f.v[5] = b.y, f.v[6] = b.z, f.v[7] = b.w;
  return f;
}
__device__ __forceinline__ uint4 ldg16(const uint16_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// ---------------------------------------------------------------------------------------------------------------
// mask: y = x * m[b]  (true multiply: -0.0 / NaN semantics of torch's CPU  original * mask  are preserved)
// ---------------------------------------------------------------------------------------------------------------
__global__ void mask_apply_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ y,
                                  float* __restrict__ yrev, long long batch, long long per_sample) {
  const long long total = batch * per_sample;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float m = mask[i / per_sample];
    const float v = x[i];
    if (y) y[i] = mask_mul(v, m);
    if (yrev) yrev[i] = mask_mul(mask_mul(v, -1.0f), __fsub_rn(m, 1.0f));  // original * -1 * (mask - 1), :72
  }
}

// ---------------------------------------------------------------------------------------------------------------
// BN statistics finalize (training): per-tile (sum, sumsq) partials -> mean / var -> scale / shift, running stats
// ---------------------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float2* __restrict__ partial, int tiles, int C, double inv_count, double unbias,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
                                   float* running_var, float momentum, float eps, float* scale, float* shift, float* save_mean,
                                   float* save_invstd) {
  __shared__ double sh_s[32][33];
  __shared__ double sh_q[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0, q = 0.0;
  if (c < C)
    for (int t = threadIdx.y; t < tiles; t += 32) {
      const float2 v = partial[(size_t)t * C + c];
      s += (double)v.x;
      q += (double)v.y;
    }
  sh_s[threadIdx.y][threadIdx.x] = s;
  sh_q[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int j = 1; j < 32; ++j) {
      s += sh_s[j][threadIdx.x];
      q += sh_q[j][threadIdx.x];
    }
    const double mean = s * inv_count;
    double var = q * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    save_mean[c] = (float)mean;
    save_invstd[c] = invstd;
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * unbias);
    }
  }
}

__global__ void bn_eval_coeffs_kernel(int C, const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                      float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float sc = gamma[c] * rsqrtf(rv[c] + eps);
    scale[c] = sc;
    shift[c] = beta[c] - rm[c] * sc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward: y = relu?(x*scale + shift [+ res*rscale + rshift])
// ---------------------------------------------------------------------------------------------------------------
template <bool HAS_RES, bool RES_AFFINE, bool RELU>
__global__ void __launch_bounds__(kThreads)
bn_act_fwd_kernel(const uint16_t* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                  const uint16_t* __restrict__ res, const float* __restrict__ rscale, const float* __restrict__ rshift,
                  uint16_t* __restrict__ y, long long n8, int c8) {
  const long long stride = (long long)gridDim.x * kThreads;
  long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  const int cg = (int)(i % c8);
  const F8 sc = load8f(scale + cg * 8), sh = load8f(shift + cg * 8);
  F8 rsc, rsh;
  if (HAS_RES && RES_AFFINE) {
    rsc = load8f(rscale + cg * 8);
    rsh = load8f(rshift + cg * 8);
  }
  for (; i < n8; i += stride) {
    F8 v = unpack8(ldg16(x + i * 8));
    F8 r;
    if (HAS_RES) r = unpack8(ldg16(res + i * 8));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float o = fmaf(v.v[j], sc.v[j], sh.v[j]);
      if (HAS_RES) o += RES_AFFINE ? fmaf(r.v[j], rsc.v[j], rsh.v[j]) : r.v[j];
      v.v[j] = RELU ? fmaxf(o, 0.f) : o;
    }
    *reinterpret_cast<uint4*>(y + i * 8) = pack8(v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward pass 1: per-block partials of  sum g  and  sum g*xhat,   g = (dy1 [+ dy2]) * (y > 0)
// ---------------------------------------------------------------------------------------------------------------
template <bool TWO, bool RELU>
__global__ void __launch_bounds__(kThreads)
bn_bwd_reduce_kernel(const uint16_t* __restrict__ dy1, const uint16_t* __restrict__ dy2, const uint16_t* __restrict__ y,
                     const uint16_t* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                     float2* __restrict__ partial, long long n8, int c8) {
  __shared__ float sh[kThreads][17];
  const long long stride = (long long)gridDim.x * kThreads;
  long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  const int cg = (int)(i % c8);
  const F8 mu = load8f(mean + cg * 8), is = load8f(invstd + cg * 8);
  float sg[8], sgx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sg[j] = sgx[j] = 0.f;
  for (; i < n8; i += stride) {
    F8 g = unpack8(ldg16(dy1 + i * 8));
    if (TWO) {
      const F8 g2 = unpack8(ldg16(dy2 + i * 8));
#pragma unroll
      for (int j = 0; j < 8; ++j) g.v[j] += g2.v[j];
    }
    if (RELU) {
      const F8 yy = unpack8(ldg16(y + i * 8));
#pragma unroll
      for (int j = 0; j < 8; ++j) g.v[j] = yy.v[j] > 0.f ? g.v[j] : 0.f;
    }
    const F8 xv = unpack8(ldg16(x + i * 8));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (xv.v[j] - mu.v[j]) * is.v[j];
      sg[j] += g.v[j];
      sgx[j] = fmaf(g.v[j], xh, sgx[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh[threadIdx.x][j] = sg[j];
    sh[threadIdx.x][8 + j] = sgx[j];
  }
  __syncthreads();
  // threads t, t + c8, t + 2*c8, ... share a channel group (kThreads % c8 == 0)
  const int C = c8 * 8;
  for (int o = threadIdx.x; o < C; o += kThreads) {
    const int g = o >> 3, j = o & 7;
    float a = 0.f, b = 0.f;
    for (int t = g; t < kThreads; t += c8) {
      a += sh[t][j];
      b += sh[t][8 + j];
    }
    partial[(size_t)blockIdx.x * C + o] = make_float2(a, b);
  }
}

__global__ void bn_bwd_finalize_kernel(const float2* __restrict__ partial, int blocks, int C, float inv_count,
                                       const float* __restrict__ gamma, const float* __restrict__ invstd, float* dgamma,
                                       float* dbeta, float* coef) {
  __shared__ float sh_a[8][33];
  __shared__ float sh_b[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, b = 0.f;
  if (c < C)
    for (int t = threadIdx.y; t < blocks; t += 8) {
      const float2 v = partial[(size_t)t * C + c];
      a += v.x;
      b += v.y;
    }
  sh_a[threadIdx.y][threadIdx.x] = a;
  sh_b[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int j = 1; j < 8; ++j) {
      a += sh_a[j][threadIdx.x];
      b += sh_b[j][threadIdx.x];
    }
    if (dbeta) dbeta[c] = a;
    if (dgamma) dgamma[c] = b;
    coef[c] = gamma[c] * invstd[c];
    coef[C + c] = a * inv_count;
    coef[2 * C + c] = b * inv_count;
  }
}

// backward pass 2: dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat));  optional g_out = g (skip-path gradient)
template <bool TWO, bool RELU, bool GOUT>
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_kernel(const uint16_t* __restrict__ dy1, const uint16_t* __restrict__ dy2, const uint16_t* __restrict__ y,
                    const uint16_t* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ coef, uint16_t* __restrict__ dx, uint16_t* __restrict__ g_out, long long n8, int c8) {
  const long long stride = (long long)gridDim.x * kThreads;
  long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  const int cg = (int)(i % c8);
  const int C = c8 * 8;
  const F8 mu = load8f(mean + cg * 8), is = load8f(invstd + cg * 8);
  const F8 k0 = load8f(coef + cg * 8), k1 = load8f(coef + C + cg * 8), k2 = load8f(coef + 2 * C + cg * 8);
  for (; i < n8; i += stride) {
    F8 g = unpack8(ldg16(dy1 + i * 8));
    if (TWO) {
      const F8 g2 = unpack8(ldg16(dy2 + i * 8));
#pragma unroll
      for (int j = 0; j < 8; ++j) g.v[j] += g2.v[j];
    }
    if (RELU) {
      const F8 yy = unpack8(ldg16(y + i * 8));
#pragma unroll
      for (int j = 0; j < 8; ++j) g.v[j] = yy.v[j] > 0.f ? g.v[j] : 0.f;
    }
    if (GOUT) *reinterpret_cast<uint4*>(g_out + i * 8) = pack8(g);
    const F8 xv = unpack8(ldg16(x + i * 8));
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (xv.v[j] - mu.v[j]) * is.v[j];
      o.v[j] = k0.v[j] * (g.v[j] - k1.v[j] - xh * k2.v[j]);
    }
    *reinterpret_cast<uint4*>(dx + i * 8) = pack8(o);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// MaxPool2d(kernel 3, stride 2, pad 1), NHWC.  argmax = window position (0..8) of the FIRST maximum in (h, w) scan
// order (torch CPU/CUDA tie rule), NaN wins like in torch.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
maxpool_fwd_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, uint8_t* __restrict__ amax, int N, int H, int W, int C,
                   int P, int Q) {
  const int c8 = C >> 3;
  const long long total = (long long)N * P * Q * c8;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int cg = (int)(i % c8);
    long long t = i / c8;
    const int q = (int)(t % Q);
    t /= Q;
    const int p = (int)(t % P);
    const int n = (int)(t / P);
    float best[8];
    int idx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      idx[j] = -1;
    }
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * p - 1 + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int w = 2 * q - 1 + s;
        if (w < 0 || w >= W) continue;
        const F8 v = unpack8(ldg16(x + (((long long)n * H + h) * W + w) * C + cg * 8));
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (idx[j] < 0 || v.v[j] > best[j] || v.v[j] != v.v[j]) {  // torch: (val > maxval) || isnan(val)
            best[j] = v.v[j];
            idx[j] = r * 3 + s;
          }
      }
    }
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = best[j];
    *reinterpret_cast<uint4*>(y + i * 8) = pack8(o);
    uint2 a;
    a.x = (uint32_t)idx[0] | ((uint32_t)idx[1] << 8) | ((uint32_t)idx[2] << 16) | ((uint32_t)idx[3] << 24);
    a.y = (uint32_t)idx[4] | ((uint32_t)idx[5] << 8) | ((uint32_t)idx[6] << 16) | ((uint32_t)idx[7] << 24);
    *reinterpret_cast<uint2*>(amax + i * 8) = a;
  }
}

__global__ void __launch_bounds__(kThreads)
maxpool_bwd_kernel(const uint16_t* __restrict__ dy, const uint16_t* __restrict__ dy2, const uint8_t* __restrict__ amax,
                   uint16_t* __restrict__ dx, int N, int H, int W, int C, int P, int Q) {
  const int c8 = C >> 3;
  const long long total = (long long)N * H * W * c8;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int cg = (int)(i % c8);
    long long t = i / c8;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    F8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
    // windows p with 2p-1 <= h <= 2p+1
    const int p_lo = max(0, (h) / 2), p_hi = min(P - 1, (h + 1) / 2);
    const int q_lo = max(0, (w) / 2), q_hi = min(Q - 1, (w + 1) / 2);
    for (int p = p_lo; p <= p_hi; ++p) {
      const int r = h - (2 * p - 1);
      if (r < 0 || r > 2) continue;
      for (int q = q_lo; q <= q_hi; ++q) {
        const int s = w - (2 * q - 1);
        if (s < 0 || s > 2) continue;
        const long long o = ((((long long)n * P + p) * Q + q) * c8 + cg) * 8;
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(amax + o));
        F8 g = unpack8(ldg16(dy + o));
        if (dy2 != nullptr) {  // gradient arriving over two paths (conv branch + identity skip)
          const F8 g2 = unpack8(ldg16(dy2 + o));
#pragma unroll
          for (int j = 0; j < 8; ++j) g.v[j] += g2.v[j];
        }
        const uint32_t want = (uint32_t)(r * 3 + s);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t got = ((j < 4 ? a.x : a.y) >> (8 * (j & 3))) & 0xFFu;
          if (got == want) acc.v[j] += g.v[j];
        }
      }
    }
    *reinterpret_cast<uint4*>(dx + i * 8) = pack8(acc);
  }
}

// AdaptiveAvgPool2d((1,1)) + flatten: [N, HW, C] bf16 -> [N, C] fp32
__global__ void avgpool_fwd_kernel(const uint16_t* __restrict__ x, float* __restrict__ y, int N, int HW, int C) {
  const int c8 = C >> 3;
  const long long total = (long long)N * c8;
  const float inv = 1.f / (float)HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c8);
    const long long n = i / c8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int t = 0; t < HW; ++t) {
      const F8 v = unpack8(ldg16(x + ((n * HW + t) * C) + cg * 8));
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v.v[j];
    }
    float* dst = y + n * C + cg * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
  }
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, uint16_t* __restrict__ dx, int N, int HW, int C) {
  const int c8 = C >> 3;
  const long long total = (long long)N * HW * c8;
  const float inv = 1.f / (float)HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c8);
    const long long n = i / ((long long)c8 * HW);
    F8 g = load8f(dy + n * C + cg * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) g.v[j] *= inv;
    *reinterpret_cast<uint4*>(dx + i * 8) = pack8(g);
  }
}

int ew_grid(const mml_ctx* ctx, long long items) {
  long long b = mml_ceil_div(items, kThreads);
  const long long cap = (long long)ctx->sm_count * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int check_rows_c(mml_ctx* ctx, int64_t rows, int C) {
  MML_REQUIRE(ctx, rows >= 1, "rows must be >= 1");
  MML_REQUIRE(ctx, C >= 8 && C <= 2048 && (C % 8) == 0 && (kThreads % (C / 8)) == 0,
              "channel count %d unsupported by the fused BN kernels (need C/8 to divide 256)", C);
  return MML_OK;
}

}  // namespace

extern "C" {

int mml_mask_apply_f32(mml_ctx* ctx, const float* x, const float* mask, float* y, float* y_reverse, int64_t batch,
                       int64_t per_sample, void* stream) {
  MML_REQUIRE(ctx, ctx && x && mask && (y || y_reverse) && batch >= 0 && per_sample >= 0, "mask_apply: bad arguments");
  if (batch * per_sample == 0) return MML_OK;
  mask_apply_kernel<<<ew_grid(ctx, batch * per_sample), kThreads, 0, (cudaStream_t)stream>>>(x, mask, y, y_reverse, batch, per_sample);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_bn_finalize(mml_ctx* ctx, const float* stats_partial, int tiles, int C, int64_t count, const float* gamma,
                    const float* beta, float* running_mean, float* running_var, float momentum, float eps, float* scale,
                    float* shift, float* save_mean, float* save_invstd, void* stream) {
  MML_REQUIRE(ctx, ctx && stats_partial && gamma && beta && scale && shift && save_mean && save_invstd, "bn_finalize: null pointer");
  MML_REQUIRE(ctx, tiles >= 1 && C >= 1 && count >= 1, "bn_finalize: bad sizes");
  const double unbias = count > 1 ? (double)count / (double)(count - 1) : 1.0;
  bn_finalize_kernel<<<(C + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(stats_partial), tiles, C, 1.0 / (double)count, unbias, gamma, beta, running_mean, running_var,
      momentum, eps, scale, shift, save_mean, save_invstd);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_bn_eval_coeffs(mml_ctx* ctx, int C, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, void* stream) {
  MML_REQUIRE(ctx, ctx && gamma && beta && running_mean && running_var && scale && shift && C >= 1, "bn_eval_coeffs: bad arguments");
  bn_eval_coeffs_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(C, gamma, beta, running_mean, running_var, eps, scale, shift);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_bn_act_fwd(mml_ctx* ctx, const uint16_t* x, const float* scale, const float* shift, const uint16_t* res,
                   const float* rscale, const float* rshift, uint16_t* y, int64_t rows, int C, int relu, void* stream) {
  MML_REQUIRE(ctx, ctx && x && scale && shift && y, "bn_act_fwd: null pointer");
  int rc = check_rows_c(ctx, rows, C);
  if (rc) return rc;
  MML_REQUIRE(ctx, (rscale == nullptr) == (rshift == nullptr), "bn_act_fwd: rscale/rshift must be given together");
  const long long n8 = rows * (C / 8);
  const int grid = ew_grid(ctx, n8);
  cudaStream_t st = (cudaStream_t)stream;
#define MML_FWD(HR, RA, RL) bn_act_fwd_kernel<HR, RA, RL><<<grid, kThreads, 0, st>>>(x, scale, shift, res, rscale, rshift, y, n8, C / 8)
  if (res == nullptr) {
    if (relu) MML_FWD(false, false, true); else MML_FWD(false, false, false);
  } else if (rscale == nullptr) {
    if (relu) MML_FWD(true, false, true); else MML_FWD(true, false, false);
  } else {
    if (relu) MML_FWD(true, true, true); else MML_FWD(true, true, false);
  }
#undef MML_FWD
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_bn_bwd_blocks(const mml_ctx* ctx, int64_t rows, int C) {
  if (!ctx || rows < 1 || C < 8) return 0;
  long long b = mml_ceil_div(rows * (C / 8), (long long)kThreads * 4);
  const long long cap = (long long)ctx->sm_count * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int mml_bn_bwd_reduce(mml_ctx* ctx, const uint16_t* dy1, const uint16_t* dy2, const uint16_t* y, const uint16_t* x,
                      const float* mean, const float* invstd, float* partial, int64_t rows, int C, int relu, void* stream) {
  MML_REQUIRE(ctx, ctx && dy1 && x && mean && invstd && partial && (!relu || y), "bn_bwd_reduce: null pointer");
  int rc = check_rows_c(ctx, rows, C);
  if (rc) return rc;
  const long long n8 = rows * (C / 8);
  const int grid = mml_bn_bwd_blocks(ctx, rows, C);
  cudaStream_t st = (cudaStream_t)stream;
  float2* part = reinterpret_cast<float2*>(partial);
#define MML_RED(TW, RL) bn_bwd_reduce_kernel<TW, RL><<<grid, kThreads, 0, st>>>(dy1, dy2, y, x, mean, invstd, part, n8, C / 8)
  if (dy2) {
    if (relu) MML_RED(true, true); else MML_RED(true, false);
  } else {
    if (relu) MML_RED(false, true); else MML_RED(false, false);
  }
#undef MML_RED
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_bn_bwd_finalize(mml_ctx* ctx, const float* partial, int blocks, int C, int64_t count, const float* gamma,
                        const float* invstd, float* dgamma, float* dbeta, float* coef, void* stream) {
  MML_REQUIRE(ctx, ctx && partial && gamma && invstd && coef && blocks >= 1 && C >= 1 && count >= 1, "bn_bwd_finalize: bad arguments");
  bn_bwd_finalize_kernel<<<(C + 31) / 32, dim3(32, 8), 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(partial), blocks, C,
                                                                                 1.0f / (float)count, gamma, invstd, dgamma, dbeta, coef);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_bn_bwd_apply(mml_ctx* ctx, const uint16_t* dy1, const uint16_t* dy2, const uint16_t* y, const uint16_t* x,
                     const float* mean, const float* invstd, const float* coef, uint16_t* dx, uint16_t* g_out, int64_t rows,
                     int C, int relu, void* stream) {
  MML_REQUIRE(ctx, ctx && dy1 && x && mean && invstd && coef && dx && (!relu || y), "bn_bwd_apply: null pointer");
  int rc = check_rows_c(ctx, rows, C);
  if (rc) return rc;
  const long long n8 = rows * (C / 8);
  const int grid = ew_grid(ctx, n8);
  cudaStream_t st = (cudaStream_t)stream;
#define MML_APP(TW, RL, GO) \
  bn_bwd_apply_kernel<TW, RL, GO><<<grid, kThreads, 0, st>>>(dy1, dy2, y, x, mean, invstd, coef, dx, g_out, n8, C / 8)
  const int key = (dy2 ? 4 : 0) | (relu ? 2 : 0) | (g_out ? 1 : 0);
  switch (key) {
    case 0: MML_APP(false, false, false); break;
    case 1: MML_APP(false, false, true); break;
    case 2: MML_APP(false, true, false); break;
    case 3: MML_APP(false, true, true); break;
    case 4: MML_APP(true, false, false); break;
    case 5: MML_APP(true, false, true); break;
    case 6: MML_APP(true, true, false); break;
    default: MML_APP(true, true, true); break;
  }
#undef MML_APP
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_maxpool3x3s2_fwd(mml_ctx* ctx, const uint16_t* x, uint16_t* y, uint8_t* argmax, int N, int H, int W, int C, void* stream) {
  MML_REQUIRE(ctx, ctx && x && y && argmax && N >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "maxpool_fwd: bad arguments");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  maxpool_fwd_kernel<<<ew_grid(ctx, (long long)N * P * Q * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>(x, y, argmax, N, H, W, C, P, Q);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_maxpool3x3s2_bwd(mml_ctx* ctx, const uint16_t* dy, const uint16_t* dy2, const uint8_t* argmax, uint16_t* dx, int N, int H,
                         int W, int C, void* stream) {
  MML_REQUIRE(ctx, ctx && dy && dx && argmax && N >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "maxpool_bwd: bad arguments");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  maxpool_bwd_kernel<<<ew_grid(ctx, (long long)N * H * W * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>(dy, dy2, argmax, dx, N, H, W, C, P, Q);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_avgpool_fwd(mml_ctx* ctx, const uint16_t* x, float* y, int N, int HW, int C, void* stream) {
  MML_REQUIRE(ctx, ctx && x && y && N >= 1 && HW >= 1 && C >= 8 && C % 8 == 0, "avgpool_fwd: bad arguments");
  avgpool_fwd_kernel<<<ew_grid(ctx, (long long)N * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>(x, y, N, HW, C);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_avgpool_bwd(mml_ctx* ctx, const float* dy, uint16_t* dx, int N, int HW, int C, void* stream) {
  MML_REQUIRE(ctx, ctx && dy && dx && N >= 1 && HW >= 1 && C >= 8 && C % 8 == 0, "avgpool_bwd: bad arguments");
  avgpool_bwd_kernel<<<ew_grid(ctx, (long long)N * HW * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>(dy, dx, N, HW, C);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

}  // extern "C"
