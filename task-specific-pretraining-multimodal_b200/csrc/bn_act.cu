// bn_act.cu -- HBM-bound fused kernels around the convolutions: BatchNorm statistics finalize, BN-apply + ReLU
// (+ residual) forward, its two-pass backward, max / average pooling, and the missing-modality mask.
//
// Reference ops replaced (MML_Suite/models/msa/networks/resnet.py): nn.BatchNorm2d (:26,31,138,177), nn.ReLU (:27,139),
// residual add (:51), nn.MaxPool2d(3,2,1) (:140), nn.AdaptiveAvgPool2d((1,1)) (:149); data/base_dataset.py:70-72 (mask).
// Activations are NHWC bf16 viewed as a [rows, C] matrix; every thread owns 8 consecutive channels (one 16-byte
// access) and, because 256-thread blocks stride by a multiple of C/8, the SAME 8 channels for its whole life, so the
// per-channel coefficients live in registers.
#include <mutex>
#include <unordered_map>

#include "mml_common.cuh"
#include "mml_ctx.h"

using namespace mml;

extern "C" int mml_g_bn_one_wave;
int mml_g_bn_one_wave = 2;  // mml_debug_set key 3 (A/B switch): 0 = old grid caps, 1 = BatchNorm grids capped at one resident wave, 2 = ... of the SM budget

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
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  const long long total = batch * per_sample;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float m = mask[i / per_sample];
    const float v = x[i];
    if (y) y[i] = mask_mul(v, m);
    if (yrev) yrev[i] = mask_mul(mask_mul(v, -1.0f), __fsub_rn(m, 1.0f));  // original * -1 * (mask - 1), :72
  }
}

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm forward.  Training mode: every CTA derives scale / shift for all C channels from the fp64 sums of the conv epilogue
// (16 bytes per channel, no finalize launch); block 0 also saves mean / invstd for backward and updates the running statistics:
// running = (1-m)*running + m*batch, unbiased variance (torch.nn.BatchNorm2d, momentum 0.1).  Eval mode: coefficients given.
//   y = relu?(x*scale + shift [+ res | + res*rscale + rshift])
// ---------------------------------------------------------------------------------------------------------------
struct BnTrain {
  const double* stats;   // [stat_slots(C)][C][2]
  const float* gamma;
  const float* beta;
  float* running_mean;   // may be null
  float* running_var;
  float* save_mean;
  float* save_invstd;
};

constexpr int kMaxC = 512;

__device__ __forceinline__ void bn_prologue(const BnTrain& b, int C, double inv_count, double unbias, float momentum, float eps, float* s_scale,
                                            float* s_shift) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double sum, sq;
    stat_load(b.stats, C, c, sum, sq);
    const double mean = sum * inv_count;
    double var = sq * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = b.gamma[c] * invstd;
    const float mu = (float)mean;
    s_scale[c] = sc;
    s_shift[c] = b.beta[c] - mu * sc;
    if (blockIdx.x == 0) {
      b.save_mean[c] = mu;
      b.save_invstd[c] = invstd;
      if (b.running_mean) {
        b.running_mean[c] = (1.f - momentum) * b.running_mean[c] + momentum * mu;
        b.running_var[c] = (1.f - momentum) * b.running_var[c] + momentum * (float)(var * unbias);
      }
    }
  }
}

__global__ void bn_eval_coeffs_kernel(int C, const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                      float* scale, float* shift) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float sc = gamma[c] * rsqrtf(rv[c] + eps);
    scale[c] = sc;
    shift[c] = beta[c] - rm[c] * sc;
  }
}

constexpr int kU = 4;  // independent 16-byte loads in flight per tensor and thread (HBM latency x bandwidth / resident threads)

// RES: 0 = none, 1 = identity residual, 2 = residual with its own affine (training: its own batch statistics -- downsample path)
// TRAIN: coefficients from the statistics (bn / rbn); else from the scale / shift arrays
template <int RES, bool RELU, bool TRAIN>
__global__ void __launch_bounds__(kThreads)
bn_fwd_kernel(const uint16_t* __restrict__ x, BnTrain bn, const float* __restrict__ scale, const float* __restrict__ shift,
              const uint16_t* __restrict__ res, BnTrain rbn, const float* __restrict__ rscale, const float* __restrict__ rshift,
              uint16_t* __restrict__ y, long long n8, int c8, double inv_count, double unbias, float momentum, float eps) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  __shared__ __align__(16) float s_coef[TRAIN ? (RES == 2 ? 4 : 2) * kMaxC : 4];
  const long long stride = (long long)gridDim.x * kThreads;
  long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  const int cg = (int)(i % c8);
  F8 sc, sh, rsc, rsh;
  if (TRAIN) {
    const int C = c8 * 8;
    bn_prologue(bn, C, inv_count, unbias, momentum, eps, s_coef, s_coef + kMaxC);
    if (RES == 2) bn_prologue(rbn, C, inv_count, unbias, momentum, eps, s_coef + 2 * kMaxC, s_coef + 3 * kMaxC);
    __syncthreads();
    sc = load8f(s_coef + cg * 8), sh = load8f(s_coef + kMaxC + cg * 8);
    if (RES == 2) rsc = load8f(s_coef + 2 * kMaxC + cg * 8), rsh = load8f(s_coef + 3 * kMaxC + cg * 8);
  } else {
    sc = load8f(scale + cg * 8), sh = load8f(shift + cg * 8);
    if (RES == 2) rsc = load8f(rscale + cg * 8), rsh = load8f(rshift + cg * 8);
  }
  auto one = [&](const uint4& xu, const uint4& ru) {
    F8 v = unpack8(xu);
    F8 r;
    if (RES != 0) r = unpack8(ru);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float o = fmaf(v.v[j], sc.v[j], sh.v[j]);
      if (RES == 1) o += r.v[j];
      if (RES == 2) o += fmaf(r.v[j], rsc.v[j], rsh.v[j]);
      v.v[j] = RELU ? fmaxf(o, 0.f) : o;
    }
    return pack8(v);
  };
  for (; i + (kU - 1) * stride < n8; i += kU * stride) {
    uint4 xu[kU], ru[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      xu[u] = ldg16(x + (i + u * stride) * 8);
      if (RES != 0) ru[u] = ldg16(res + (i + u * stride) * 8);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) *reinterpret_cast<uint4*>(y + (i + u * stride) * 8) = one(xu[u], ru[u]);
  }
  for (; i < n8; i += stride) {
    const uint4 xu = ldg16(x + i * 8);
    uint4 ru = make_uint4(0, 0, 0, 0);
    if (RES != 0) ru = ldg16(res + i * 8);
    *reinterpret_cast<uint4*>(y + i * 8) = one(xu, ru);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward pass 1: g = (dy1 [+ dy2]) * (y > 0), optionally stored (it IS the gradient of the identity skip path, and pass 2
// then reads g instead of dy1 / dy2 / y);  per-block partials of  sum g  and  sum g*xhat  -> one fp64 atomic per channel and CTA
// ---------------------------------------------------------------------------------------------------------------
// block-level channel reduction of the per-thread partials + atomics (shared by the two reduce kernels)
__device__ __forceinline__ void bn_bwd_block_reduce(const float (&sg)[8], const float (&sgx)[8], float (*sh)[17], double* bstat, int c8) {
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
    stat_add(bstat, C, blockIdx.x, o, a, b);
  }
}

template <bool TWO, bool RELU, bool GOUT>
__global__ void __launch_bounds__(kThreads)
bn_bwd_reduce_kernel(const uint16_t* __restrict__ dy1, const uint16_t* __restrict__ dy2, const uint16_t* __restrict__ y,
                     const uint16_t* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                     double* __restrict__ bstat, uint16_t* __restrict__ g_out, long long n8, int c8) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  __shared__ float sh[kThreads][17];
  const long long stride = (long long)gridDim.x * kThreads;
  long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  const int cg = (int)(i % c8);
  const F8 mu = load8f(mean + cg * 8), is = load8f(invstd + cg * 8);
  float sg[8], sgx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sg[j] = sgx[j] = 0.f;
  auto one = [&](long long idx, const uint4& d1, const uint4& d2, const uint4& yu, const uint4& xu) {
    F8 g = unpack8(d1);
    if (TWO) {
      const F8 g2 = unpack8(d2);
#pragma unroll
      for (int j = 0; j < 8; ++j) g.v[j] += g2.v[j];
    }
    if (RELU) {
      const F8 yy = unpack8(yu);
#pragma unroll
      for (int j = 0; j < 8; ++j) g.v[j] = yy.v[j] > 0.f ? g.v[j] : 0.f;
    }
    if (GOUT) {
      const uint4 gp = pack8(g);
      *reinterpret_cast<uint4*>(g_out + idx * 8) = gp;
      g = unpack8(gp);  // pass 2 reads the ROUNDED g: the sums must describe the same values
    }
    const F8 xv = unpack8(xu);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (xv.v[j] - mu.v[j]) * is.v[j];
      sg[j] += g.v[j];
      sgx[j] = fmaf(g.v[j], xh, sgx[j]);
    }
  };
  constexpr int U = 2;
  for (; i + (U - 1) * stride < n8; i += U * stride) {
    uint4 d1[U], d2[U], yu[U], xu[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long idx = i + u * stride;
      d1[u] = ldg16(dy1 + idx * 8);
      if (TWO) d2[u] = ldg16(dy2 + idx * 8);
      if (RELU) yu[u] = ldg16(y + idx * 8);
      xu[u] = ldg16(x + idx * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) one(i + u * stride, d1[u], d2[u], yu[u], xu[u]);
  }
  for (; i < n8; i += stride) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    one(i, ldg16(dy1 + i * 8), TWO ? ldg16(dy2 + i * 8) : z, RELU ? ldg16(y + i * 8) : z, ldg16(x + i * 8));
  }
  bn_bwd_block_reduce(sg, sgx, sh, bstat, c8);
}

// backward pass 2: dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)); coefficients straight from the fp64 sums of pass 1 (no
// finalize launch); block 0 writes dgamma / dbeta
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_kernel(const uint16_t* __restrict__ g, const uint16_t* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma, const double* __restrict__ bstat, float inv_count,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, uint16_t* __restrict__ dx, long long n8, int c8) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  __shared__ __align__(16) float s_k[2 * kMaxC];
  const int C = c8 * 8;
  for (int c = threadIdx.x; c < C; c += kThreads) {
    double sg, sgx;
    stat_load(bstat, C, c, sg, sgx);
    s_k[c] = (float)sg * inv_count;          // mean(g)
    s_k[kMaxC + c] = (float)sgx * inv_count; // mean(g * xhat)
    if (blockIdx.x == 0) {
      if (dbeta) dbeta[c] = (float)sg;
      if (dgamma) dgamma[c] = (float)sgx;
    }
  }
  __syncthreads();
  // Traverse from the END of the tensors: pass 1 just streamed the same operands front to back, so their tails are what is
  // still resident in the 126 MB L2.
  const long long stride = (long long)gridDim.x * kThreads;
  long long i = n8 - 1 - (blockIdx.x * (long long)kThreads + threadIdx.x);
  const int cg = (int)(((i % c8) + c8) % c8);
  const F8 mu = load8f(mean + cg * 8), is = load8f(invstd + cg * 8);
  F8 k0 = load8f(gamma + cg * 8);
  const F8 k1 = load8f(s_k + cg * 8), k2 = load8f(s_k + kMaxC + cg * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) k0.v[j] *= is.v[j];  // gamma * invstd
  auto one = [&](const uint4& gu, const uint4& xu) {
    const F8 gv = unpack8(gu), xv = unpack8(xu);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (xv.v[j] - mu.v[j]) * is.v[j];
      o.v[j] = k0.v[j] * (gv.v[j] - k1.v[j] - xh * k2.v[j]);
    }
    return pack8(o);
  };
  for (; i - (kU - 1) * stride >= 0; i -= kU * stride) {
    uint4 gu[kU], xu[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      gu[u] = ldg16(g + (i - u * stride) * 8);
      xu[u] = ldg16(x + (i - u * stride) * 8);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) *reinterpret_cast<uint4*>(dx + (i - u * stride) * 8) = one(gu[u], xu[u]);
  }
  for (; i >= 0; i -= stride) *reinterpret_cast<uint4*>(dx + i * 8) = one(ldg16(g + i * 8), ldg16(x + i * 8));
}

// ---------------------------------------------------------------------------------------------------------------
// MaxPool2d(kernel 3, stride 2, pad 1), NHWC.  argmax = window position (0..8) of the FIRST maximum in (h, w) scan
// order (torch CPU/CUDA tie rule), NaN wins like in torch.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
maxpool_fwd_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, uint8_t* __restrict__ amax, int N, int H, int W, int C,
                   int P, int Q) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
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

// One thread = the 2x2 input block (2a..2a+1, 2b..2b+1) x 8 channels.  Those four positions only belong to the windows
// p in {a, a+1}, q in {b, b+1}, which are loaded once (dy [+ dy2] and the argmax bytes) and scattered to the four outputs.
__global__ void __launch_bounds__(kThreads)
maxpool_bwd_kernel(const uint16_t* __restrict__ dy, const uint16_t* __restrict__ dy2, const uint8_t* __restrict__ amax,
                   uint16_t* __restrict__ dx, int N, int H, int W, int C, int P, int Q) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  const int c8 = C >> 3;
  const int HB = (H + 1) >> 1, WB = (W + 1) >> 1;
  const long long total = (long long)N * HB * WB * c8;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int cg = (int)(i % c8);
    long long t = i / c8;
    const int b = (int)(t % WB);
    t /= WB;
    const int a = (int)(t % HB);
    const int n = (int)(t / HB);
    F8 acc[2][2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int v = 0; v < 2; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[u][v].v[j] = 0.f;
#pragma unroll
    for (int dp = 0; dp < 2; ++dp) {
      const int p = a + dp;
      if (p >= P) continue;
#pragma unroll
      for (int dq = 0; dq < 2; ++dq) {
        const int q = b + dq;
        if (q >= Q) continue;
        const long long o = ((((long long)n * P + p) * Q + q) * c8 + cg) * 8;
        const uint2 am = __ldg(reinterpret_cast<const uint2*>(amax + o));
        F8 g = unpack8(ldg16(dy + o));
        if (dy2 != nullptr) {  // gradient arriving over two paths (conv branch + identity skip)
          const F8 g2 = unpack8(ldg16(dy2 + o));
#pragma unroll
          for (int j = 0; j < 8; ++j) g.v[j] += g2.v[j];
        }
        // window (p, q) covers rows 2p-1..2p+1: block row u (h = 2a+u) is window row r = 2a+u-(2p-1) = u + 1 - 2*dp
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int r = u + 1 - 2 * dp;
          if (r < 0 || r > 2) continue;
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const int s2 = v + 1 - 2 * dq;
            if (s2 < 0 || s2 > 2) continue;
            const uint32_t want = (uint32_t)(r * 3 + s2);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t got = ((j < 4 ? am.x : am.y) >> (8 * (j & 3))) & 0xFFu;
              if (got == want) acc[u][v].v[j] += g.v[j];
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int h = 2 * a + u;
      if (h >= H) continue;
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int w = 2 * b + v;
        if (w >= W) continue;
        *reinterpret_cast<uint4*>(dx + ((((long long)n * H + h) * W + w) * c8 + cg) * 8) = pack8(acc[u][v]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Stem tail, fused: BatchNorm (train or eval) + ReLU + MaxPool2d(3,2,1) in one pass over the raw stem output, and its
// backward.  The post-ReLU activation (the largest tensor of the network: 103 MB at batch 256) is never materialised:
// backward recomputes the ReLU mask from raw*scale+shift and scatters the pooled gradient through the saved argmax.
// ---------------------------------------------------------------------------------------------------------------
template <bool TRAIN>
__global__ void __launch_bounds__(kThreads)
stem_bn_pool_fwd_kernel(const uint16_t* __restrict__ x, BnTrain bn, const float* __restrict__ scale_in, const float* __restrict__ shift_in,
                        uint16_t* __restrict__ y, uint8_t* __restrict__ amax, int N, int H, int W, int C, int P, int Q, double inv_count,
                        double unbias, float momentum, float eps) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  __shared__ __align__(16) float s_coef[2 * kMaxC];
  if (TRAIN) {
    bn_prologue(bn, C, inv_count, unbias, momentum, eps, s_coef, s_coef + kMaxC);
  } else {
    for (int c = threadIdx.x; c < C; c += kThreads) {
      s_coef[c] = scale_in[c];
      s_coef[kMaxC + c] = shift_in[c];
    }
  }
  __syncthreads();
  // One CTA walks output rows (n, p); its threads are (8-channel group) x (column q): no per-item 64-bit divisions (they cost more
  // than the rest of the item: round 1 spent five long divisions per 16 output bytes here).
  const int c8 = C >> 3;
  const int cg = (int)threadIdx.x % c8, qs = (int)threadIdx.x / c8, nq = kThreads / c8;
  const F8 sc = load8f(s_coef + cg * 8), sh = load8f(s_coef + kMaxC + cg * 8);
  uint32_t flip[4];  // sign-bit mask per packed channel pair: set where scale < 0 (max of bn(x) = bn(min x))
#pragma unroll
  for (int c2 = 0; c2 < 4; ++c2) flip[c2] = (sc.v[2 * c2] < 0.f ? 0x00008000u : 0u) | (sc.v[2 * c2 + 1] < 0.f ? 0x80000000u : 0u);
  for (int row = blockIdx.x; row < N * P; row += gridDim.x) {
    const int n = row / P, p = row - n * P;
   for (int q = qs; q < Q; q += nq) {
    const long long i = ((long long)row * Q + q) * c8 + cg;
    // all nine window loads are issued before the first one is consumed (ncu: with the load inside the compare loop every window
    // element cost a full memory round trip -- nine serialized long-scoreboard stalls per output vector, 80 us for 140 MB)
    uint4 win[9];
    uint32_t live = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * p - 1 + r;
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int w = 2 * q - 1 + s2;
        const bool ok = (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W;
        live |= (ok ? 1u : 0u) << (r * 3 + s2);
        win[r * 3 + s2] = ok ? ldg16(x + (((long long)n * H + h) * W + w) * C + cg * 8) : make_uint4(0, 0, 0, 0);
      }
    }
    // The window maximum is taken on the RAW bf16 values, two channels per instruction: x -> fma(x, scale, shift) is monotone
    // (non-decreasing for scale >= 0; for scale < 0 the sign bit of the key is flipped, which turns the max into the min of x), so the
    // maximum of the fp32 BatchNorm outputs IS the BatchNorm output of the extreme raw value -- same y bit for bit, at a quarter of
    // the instructions of comparing nine fp32 fma results per channel (ncu: the kernel was instruction-issue bound, 80 us for 140 MB).
    // ReLU and the bf16 rounding are applied once, to the winner.  argmax = first position (h, w scan order) holding the extreme value.
    uint32_t key[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const bool ok = (live >> t) & 1u;
      const uint32_t raw[4] = {win[t].x, win[t].y, win[t].z, win[t].w};
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2) key[t][c2] = ok ? (raw[c2] ^ flip[c2]) : 0xFF80FF80u;  // outside the image: -inf, never wins
    }
    uint32_t best2[4], idx2[4];
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
      __nv_bfloat162 m = *reinterpret_cast<const __nv_bfloat162*>(&key[0][c2]);
#pragma unroll
      for (int t = 1; t < 9; ++t) m = __hmax2_nan(m, *reinterpret_cast<const __nv_bfloat162*>(&key[t][c2]));  // NaN wins, like torch
      uint32_t id = 4u * 0x00010001u;  // the centre of the window is always inside the image
#pragma unroll
      for (int t = 8; t >= 0; --t) {
        uint32_t eq = __hequ2_mask(*reinterpret_cast<const __nv_bfloat162*>(&key[t][c2]), m);  // 0xFFFF per half where equal (or NaN)
        eq = ((live >> t) & 1u) ? eq : 0u;
        id = (eq & ((uint32_t)t * 0x00010001u)) | (~eq & id);
      }
      best2[c2] = *reinterpret_cast<const uint32_t*>(&m) ^ flip[c2];
      idx2[c2] = id;
    }
    F8 o;
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
      const float a0 = fmaf(bf16_lo(best2[c2]), sc.v[2 * c2], sh.v[2 * c2]), a1 = fmaf(bf16_hi(best2[c2]), sc.v[2 * c2 + 1], sh.v[2 * c2 + 1]);
      o.v[2 * c2] = a0 != a0 ? a0 : fmaxf(a0, 0.f);  // relu(NaN) = NaN like torch
      o.v[2 * c2 + 1] = a1 != a1 ? a1 : fmaxf(a1, 0.f);
    }
    *reinterpret_cast<uint4*>(y + i * 8) = pack8(o);
    uint2 a;
    a.x = (idx2[0] & 0xFFu) | ((idx2[0] >> 8) & 0xFF00u) | ((idx2[1] & 0xFFu) << 16) | ((idx2[1] & 0xFF0000u) << 8);
    a.y = (idx2[2] & 0xFFu) | ((idx2[2] >> 8) & 0xFF00u) | ((idx2[3] & 0xFFu) << 16) | ((idx2[3] & 0xFF0000u) << 8);
    *reinterpret_cast<uint2*>(amax + i * 8) = a;
   }
  }
}

// gradient of the pooled output scattered to the 2x2 input block (2a.., 2b..) of one thread (see maxpool_bwd_kernel)
__device__ __forceinline__ void pool_scatter_2x2(const uint16_t* __restrict__ dy, const uint16_t* __restrict__ dy2,
                                                 const uint8_t* __restrict__ amax, int n, int a, int b, int cg, int c8, int P, int Q, F8 (&acc)[2][2]) {
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 2; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[u][v].v[j] = 0.f;
#pragma unroll
  for (int dp = 0; dp < 2; ++dp) {
    const int p = a + dp;
    if (p >= P) continue;
#pragma unroll
    for (int dq = 0; dq < 2; ++dq) {
      const int q = b + dq;
      if (q >= Q) continue;
      const long long o = ((((long long)n * P + p) * Q + q) * c8 + cg) * 8;
      const uint2 am = __ldg(reinterpret_cast<const uint2*>(amax + o));
      F8 g = unpack8(ldg16(dy + o));
      if (dy2 != nullptr) {
        const F8 g2 = unpack8(ldg16(dy2 + o));
#pragma unroll
        for (int j = 0; j < 8; ++j) g.v[j] += g2.v[j];
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = u + 1 - 2 * dp;
        if (r < 0 || r > 2) continue;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int s2 = v + 1 - 2 * dq;
          if (s2 < 0 || s2 > 2) continue;
          const uint32_t want = (uint32_t)(r * 3 + s2);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t got = ((j < 4 ? am.x : am.y) >> (8 * (j & 3))) & 0xFFu;
            if (got == want) acc[u][v].v[j] += g.v[j];
          }
        }
      }
    }
  }
}

// bstat += (sum g, sum g*xhat) and dx = g (bn_bwd_apply_kernel then finishes in place);  g = scatter(dpool) * (bn(x) > 0),
// bn(x) = gamma*(x-mean)*invstd + beta.  One thread = the 2x2 input block (2a.., 2b..) x FOUR channels: with eight channels the
// per-channel coefficients, the four window gradients and the 2x2 accumulators need 140 registers (one 256-thread CTA per SM, 126 us
// for 270 MB); four channels fit in half of that.
__global__ void __launch_bounds__(kThreads, 3)
stem_bn_pool_bwd_kernel(const uint16_t* __restrict__ dy, const uint16_t* __restrict__ dy2, const uint8_t* __restrict__ amax,
                        const uint16_t* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ gamma, const float* __restrict__ beta, double* __restrict__ bstat,
                        uint16_t* __restrict__ dx, int N, int H, int W, int C, int P, int Q) {
  pdl_sync();
  __shared__ float sh[kThreads][9];
  const int c4 = C >> 2;
  const int HB = (H + 1) >> 1, WB = (W + 1) >> 1;
  // One CTA walks block rows (n, a); its threads are (4-channel group) x (block column b): no per-item 64-bit divisions
  const int cg = (int)threadIdx.x % c4, bs = (int)threadIdx.x / c4, nb = kThreads / c4;
  const float4 mu = *reinterpret_cast<const float4*>(mean + cg * 4), is = *reinterpret_cast<const float4*>(invstd + cg * 4);
  const float4 ga = *reinterpret_cast<const float4*>(gamma + cg * 4), be = *reinterpret_cast<const float4*>(beta + cg * 4);
  float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};
  for (int row = blockIdx.x; row < N * HB; row += gridDim.x) {
    const int n = row / HB, a = row - n * HB;
   for (int b = bs; b < WB; b += nb) {
    // gradients and argmax bytes of the (up to) four pooling windows p in {a, a+1}, q in {b, b+1} that cover this block
    // Every global load of this block -- argmax bytes and gradient(s) of its four windows, the four raw values -- is issued before any
    // of them is consumed (ncu: loaded at their first use they cost five to eight serialized memory round trips per block, 141 us for
    // 240 MB; the loads are predicated, so out-of-range windows cost nothing).
    uint32_t am[2][2];
    uint2 g1[2][2], g2[2][2], xblk[2][2];
#pragma unroll
    for (int dp = 0; dp < 2; ++dp)
#pragma unroll
      for (int dq = 0; dq < 2; ++dq) {
        const int p = a + dp, q = b + dq;
        const bool ok = p < P && q < Q;
        const long long o = (((long long)n * P + p) * Q + q) * C + cg * 4;
        am[dp][dq] = ok ? __ldg(reinterpret_cast<const uint32_t*>(amax + o)) : 0xFFFFFFFFu;  // 0xFF matches no window position
        g1[dp][dq] = ok ? __ldg(reinterpret_cast<const uint2*>(dy + o)) : make_uint2(0, 0);
        g2[dp][dq] = (ok && dy2 != nullptr) ? __ldg(reinterpret_cast<const uint2*>(dy2 + o)) : make_uint2(0, 0);  // second gradient path
      }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int h = 2 * a + u, w = 2 * b + v;
        xblk[u][v] = (h < H && w < W) ? __ldg(reinterpret_cast<const uint2*>(x + (((long long)n * H + h) * W + w) * C + cg * 4)) : make_uint2(0, 0);
      }
    float gw[2][2][4];
#pragma unroll
    for (int dp = 0; dp < 2; ++dp)
#pragma unroll
      for (int dq = 0; dq < 2; ++dq) {
        gw[dp][dq][0] = bf16_lo(g1[dp][dq].x) + bf16_lo(g2[dp][dq].x), gw[dp][dq][1] = bf16_hi(g1[dp][dq].x) + bf16_hi(g2[dp][dq].x);
        gw[dp][dq][2] = bf16_lo(g1[dp][dq].y) + bf16_lo(g2[dp][dq].y), gw[dp][dq][3] = bf16_hi(g1[dp][dq].y) + bf16_hi(g2[dp][dq].y);
      }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int h = 2 * a + u;
      if (h >= H) continue;
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int w = 2 * b + v;
        if (w >= W) continue;
        const long long o = (((long long)n * H + h) * W + w) * C + cg * 4;
        const uint2 xu = xblk[u][v];
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        // window (p, q) covers rows 2p-1..2p+1: block row u (h = 2a+u) is window row r = u + 1 - 2*dp (valid: 0..2)
#pragma unroll
        for (int dp = 0; dp < 2; ++dp) {
          const int r = u + 1 - 2 * dp;
          if (r < 0 || r > 2) continue;
#pragma unroll
          for (int dq = 0; dq < 2; ++dq) {
            const int s2 = v + 1 - 2 * dq;
            if (s2 < 0 || s2 > 2) continue;
            const uint32_t want = (uint32_t)(r * 3 + s2);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (((am[dp][dq] >> (8 * j)) & 0xFFu) == want) acc[j] += gw[dp][dq][j];
          }
        }
        const float xv[4] = {bf16_lo(xu.x), bf16_hi(xu.x), bf16_lo(xu.y), bf16_hi(xu.y)};
        const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, isv[4] = {is.x, is.y, is.z, is.w};
        const float gav[4] = {ga.x, ga.y, ga.z, ga.w}, bev[4] = {be.x, be.y, be.z, be.w};
        float xh[4], g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xh[j] = (xv[j] - muv[j]) * isv[j];
          g[j] = fmaf(gav[j], xh[j], bev[j]) > 0.f ? acc[j] : 0.f;  // ReLU mask recomputed
        }
        const uint2 gp = make_uint2(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]));
        *reinterpret_cast<uint2*>(dx + o) = gp;  // materialised once: the apply pass is then a plain elementwise kernel
        const float gr[4] = {bf16_lo(gp.x), bf16_hi(gp.x), bf16_lo(gp.y), bf16_hi(gp.y)};  // sums of the ROUNDED g (what pass 2 reads)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sg[j] += gr[j];
          sgx[j] = fmaf(gr[j], xh[j], sgx[j]);
        }
      }
    }
   }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sh[threadIdx.x][j] = sg[j];
    sh[threadIdx.x][4 + j] = sgx[j];
  }
  __syncthreads();
  // threads t, t + c4, t + 2*c4, ... share a channel group (kThreads % c4 == 0)
  for (int o = threadIdx.x; o < C; o += kThreads) {
    const int gch = o >> 2, j = o & 3;
    float aa = 0.f, bb = 0.f;
    for (int t = gch; t < kThreads; t += c4) {
      aa += sh[t][j];
      bb += sh[t][4 + j];
    }
    stat_add(bstat, C, blockIdx.x, o, aa, bb);
  }
}

// AdaptiveAvgPool2d((1,1)) + flatten: [N, HW, C] bf16 -> [N, C] fp32
__global__ void avgpool_fwd_kernel(const uint16_t* __restrict__ x, float* __restrict__ y, int N, int HW, int C) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
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
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
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

// grid for the kernels whose main loop keeps U 16-byte loads per tensor in flight: enough items per thread for the unrolled
// loop on large tensors, one item per thread spread over the SMs on small ones
int stream_grid(const mml_ctx* ctx, long long items, int U) {
  long long b = mml_ceil_div(items, (long long)kThreads * U);
  if (b < ctx->sm_count) {
    b = mml_ceil_div(items, kThreads);
    if (b > ctx->sm_count) b = ctx->sm_count;
  }
  const long long cap = (long long)ctx->sm_count * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// One resident wave: a grid capped at (CTAs that fit on an SM) x SMs, so every CTA runs its prologue (statistics -> coefficients,
// two dependent loads) once and then streams ~3x more vectors, instead of 2-3 waves of short CTAs each paying that latency
// (measured on the ResNet18 layer1 tensor, mml_debug_set key 3 switches it: see DESIGN.md section 4).
template <typename K>
int wave_cap(const mml_ctx* ctx, int grid, K kernel, cudaStream_t st) {
  if (!mml_g_bn_one_wave) return grid;
  // The occupancy of each kernel is queried ONCE and outside stream capture (a query during capture invalidated a CUDA-graph capture
  // on the GPU box; every schedule runs two eager steps before it is captured, so the cache is warm by then), kept per kernel address.
  static std::mutex mu;
  static std::unordered_map<const void*, int> cache;
  int occ = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find((const void*)kernel);
    if (it != cache.end()) occ = it->second;
  }
  if (occ == 0) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return grid;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, 0) != cudaSuccess || occ < 1) {
      cudaGetLastError();
      return grid;
    }
    std::lock_guard<std::mutex> lock(mu);
    cache[(const void*)kernel] = occ;
  }
  // mode 2: one wave of the SM BUDGET (mml_ctx_set_sm_budget) -- the audio encoder's BatchNorm kernels then fill 116 SMs' worth of
  // register file instead of all 148, which leaves room on every SM for the image stream's GEMM CTAs to co-reside
  const int sms = (mml_g_bn_one_wave == 2 && ctx->sm_budget > 0) ? ctx->sm_budget : ctx->sm_count;
  const int cap = occ * sms;
  return grid < cap ? grid : cap;
}

int check_rows_c(mml_ctx* ctx, int64_t rows, int C) {
  MML_REQUIRE(ctx, rows >= 1, "rows must be >= 1");
  MML_REQUIRE(ctx, C >= 8 && C <= kMaxC && (C % 8) == 0 && (kThreads % (C / 8)) == 0,
              "channel count %d unsupported by the fused BN kernels (need C <= 512 and C/8 to divide 256)", C);
  return MML_OK;
}

}  // namespace

extern "C" {

int mml_mask_apply_f32(mml_ctx* ctx, const float* x, const float* mask, float* y, float* y_reverse, int64_t batch,
                       int64_t per_sample, void* stream) {
  MML_REQUIRE(ctx, ctx && x && mask && (y || y_reverse) && batch >= 0 && per_sample >= 0, "mask_apply: bad arguments");
  if (batch * per_sample == 0) return MML_OK;
  MML_LAUNCH(ctx, mask_apply_kernel, ew_grid(ctx, batch * per_sample), kThreads, 0, (cudaStream_t)stream, x, mask, y, y_reverse, batch, per_sample);
  return MML_OK;
}

int mml_bn_train_fwd(mml_ctx* ctx, const uint16_t* x, const double* stats, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float* save_mean, float* save_invstd, const uint16_t* res, const double* rstats,
                     const float* rgamma, const float* rbeta, float* r_running_mean, float* r_running_var, float* r_save_mean,
                     float* r_save_invstd, uint16_t* y, int64_t rows, int C, int relu, float momentum, float eps, void* stream) {
  MML_REQUIRE(ctx, ctx && x && stats && gamma && beta && save_mean && save_invstd && y, "bn_train_fwd: null pointer");
  MML_REQUIRE(ctx, (running_mean == nullptr) == (running_var == nullptr), "bn_train_fwd: running_mean / running_var must be given together");
  int rc = check_rows_c(ctx, rows, C);
  if (rc) return rc;
  const int mode = res == nullptr ? 0 : (rstats == nullptr ? 1 : 2);
  if (mode == 2) MML_REQUIRE(ctx, rgamma && rbeta && r_save_mean && r_save_invstd, "bn_train_fwd: residual BN needs gamma/beta/save buffers");
  BnTrain bn{stats, gamma, beta, running_mean, running_var, save_mean, save_invstd};
  BnTrain rbn{rstats, rgamma, rbeta, r_running_mean, r_running_var, r_save_mean, r_save_invstd};
  const long long n8 = rows * (C / 8);
  const int grid = stream_grid(ctx, n8, kU);
  const double inv_count = 1.0 / (double)rows;
  const double unbias = rows > 1 ? (double)rows / (double)(rows - 1) : 1.0;
  cudaStream_t st = (cudaStream_t)stream;
#define MML_TR(M, RL) \
  MML_LAUNCH(ctx, (bn_fwd_kernel<M, RL, true>), wave_cap(ctx, grid, bn_fwd_kernel<M, RL, true>, st), kThreads, 0, st, x, bn, nullptr, nullptr, res, rbn, nullptr, nullptr, y, n8, C / 8, inv_count, unbias, momentum, eps)
  switch (mode * 2 + (relu ? 1 : 0)) {
    case 0: MML_TR(0, false); break;
    case 1: MML_TR(0, true); break;
    case 2: MML_TR(1, false); break;
    case 3: MML_TR(1, true); break;
    case 4: MML_TR(2, false); break;
    default: MML_TR(2, true); break;
  }
#undef MML_TR
  return MML_OK;
}

int mml_bn_eval_coeffs(mml_ctx* ctx, int C, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, void* stream) {
  MML_REQUIRE(ctx, ctx && gamma && beta && running_mean && running_var && scale && shift && C >= 1, "bn_eval_coeffs: bad arguments");
  MML_LAUNCH(ctx, bn_eval_coeffs_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, C, gamma, beta, running_mean, running_var, eps, scale, shift);
  return MML_OK;
}

int mml_bn_act_fwd(mml_ctx* ctx, const uint16_t* x, const float* scale, const float* shift, const uint16_t* res,
                   const float* rscale, const float* rshift, uint16_t* y, int64_t rows, int C, int relu, void* stream) {
  MML_REQUIRE(ctx, ctx && x && scale && shift && y, "bn_act_fwd: null pointer");
  int rc = check_rows_c(ctx, rows, C);
  if (rc) return rc;
  MML_REQUIRE(ctx, (rscale == nullptr) == (rshift == nullptr), "bn_act_fwd: rscale/rshift must be given together");
  const long long n8 = rows * (C / 8);
  const int grid = stream_grid(ctx, n8, kU);
  const int mode = res == nullptr ? 0 : (rscale == nullptr ? 1 : 2);
  const BnTrain none{};
  cudaStream_t st = (cudaStream_t)stream;
#define MML_FWD(M, RL) \
  MML_LAUNCH(ctx, (bn_fwd_kernel<M, RL, false>), wave_cap(ctx, grid, bn_fwd_kernel<M, RL, false>, st), kThreads, 0, st, x, none, scale, shift, res, none, rscale, rshift, y, n8, C / 8, 0.0, 0.0, 0.f, 0.f)
  switch (mode * 2 + (relu ? 1 : 0)) {
    case 0: MML_FWD(0, false); break;
    case 1: MML_FWD(0, true); break;
    case 2: MML_FWD(1, false); break;
    case 3: MML_FWD(1, true); break;
    case 4: MML_FWD(2, false); break;
    default: MML_FWD(2, true); break;
  }
#undef MML_FWD
  return MML_OK;
}

static int bn_bwd_grid(const mml_ctx* ctx, int64_t rows, int C) {
  long long b = mml_ceil_div(rows * (C / 8), (long long)kThreads * 4);
  const long long cap = (long long)ctx->sm_count * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int mml_bn_bwd_reduce(mml_ctx* ctx, const uint16_t* dy1, const uint16_t* dy2, const uint16_t* y, const uint16_t* x,
                      const float* mean, const float* invstd, double* bstat, uint16_t* g_out, int64_t rows, int C, int relu, void* stream) {
  MML_REQUIRE(ctx, ctx && dy1 && x && mean && invstd && bstat && (!relu || y), "bn_bwd_reduce: null pointer");
  int rc = check_rows_c(ctx, rows, C);
  if (rc) return rc;
  const long long n8 = rows * (C / 8);
  const int grid = bn_bwd_grid(ctx, rows, C);
  cudaStream_t st = (cudaStream_t)stream;
#define MML_RED(TW, RL, GO) MML_LAUNCH(ctx, (bn_bwd_reduce_kernel<TW, RL, GO>), wave_cap(ctx, grid, bn_bwd_reduce_kernel<TW, RL, GO>, st), kThreads, 0, st, dy1, dy2, y, x, mean, invstd, bstat, g_out, n8, C / 8)
  const int key = (dy2 ? 4 : 0) | (relu ? 2 : 0) | (g_out ? 1 : 0);
  switch (key) {
    case 0: MML_RED(false, false, false); break;
    case 1: MML_RED(false, false, true); break;
    case 2: MML_RED(false, true, false); break;
    case 3: MML_RED(false, true, true); break;
    case 4: MML_RED(true, false, false); break;
    case 5: MML_RED(true, false, true); break;
    case 6: MML_RED(true, true, false); break;
    default: MML_RED(true, true, true); break;
  }
#undef MML_RED
  return MML_OK;
}

int mml_bn_bwd_apply(mml_ctx* ctx, const uint16_t* g, const uint16_t* x, const float* mean, const float* invstd, const float* gamma,
                     const double* bstat, float* dgamma, float* dbeta, uint16_t* dx, int64_t rows, int C, void* stream) {
  MML_REQUIRE(ctx, ctx && g && x && mean && invstd && gamma && bstat && dx, "bn_bwd_apply: null pointer");
  int rc = check_rows_c(ctx, rows, C);
  if (rc) return rc;
  const long long n8 = rows * (C / 8);
  MML_LAUNCH(ctx, bn_bwd_apply_kernel, wave_cap(ctx, stream_grid(ctx, n8, kU), bn_bwd_apply_kernel, (cudaStream_t)stream), kThreads, 0, (cudaStream_t)stream, g, x, mean, invstd, gamma, bstat, 1.0f / (float)rows, dgamma, dbeta,
                                                                                       dx, n8, C / 8);
  return MML_OK;
}

int mml_maxpool3x3s2_fwd(mml_ctx* ctx, const uint16_t* x, uint16_t* y, uint8_t* argmax, int N, int H, int W, int C, void* stream) {
  MML_REQUIRE(ctx, ctx && x && y && argmax && N >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "maxpool_fwd: bad arguments");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  MML_LAUNCH(ctx, maxpool_fwd_kernel, ew_grid(ctx, (long long)N * P * Q * (C / 8)), kThreads, 0, (cudaStream_t)stream, x, y, argmax, N, H, W, C, P, Q);
  return MML_OK;
}

int mml_maxpool3x3s2_bwd(mml_ctx* ctx, const uint16_t* dy, const uint16_t* dy2, const uint8_t* argmax, uint16_t* dx, int N, int H,
                         int W, int C, void* stream) {
  MML_REQUIRE(ctx, ctx && dy && dx && argmax && N >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "maxpool_bwd: bad arguments");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  MML_LAUNCH(ctx, maxpool_bwd_kernel, ew_grid(ctx, (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8)), kThreads, 0, (cudaStream_t)stream, dy, dy2, argmax, dx, N, H, W,
                                                                                                                                   C, P, Q);
  return MML_OK;
}

int mml_stem_bn_pool_fwd(mml_ctx* ctx, const uint16_t* x, const double* stats, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float* save_mean, float* save_invstd, const float* scale, const float* shift, uint16_t* y,
                         uint8_t* argmax, int N, int H, int W, int C, float momentum, float eps, void* stream) {
  MML_REQUIRE(ctx, ctx && x && y && argmax && N >= 1 && H >= 1 && W >= 1, "stem_bn_pool_fwd: bad arguments");
  MML_REQUIRE(ctx, (stats != nullptr) != (scale != nullptr), "stem_bn_pool_fwd: give either batch statistics (train) or scale/shift (eval)");
  int rc = check_rows_c(ctx, (int64_t)N * H * W, C);
  if (rc) return rc;
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const int64_t rows = (int64_t)N * H * W;
  int grid = N * P < ctx->sm_count * 8 ? N * P : ctx->sm_count * 8;  // CTAs walk output rows (n, p) (a one-wave cap was measured here: slower)
  BnTrain bn{stats, gamma, beta, running_mean, running_var, save_mean, save_invstd};
  cudaStream_t st = (cudaStream_t)stream;
  if (stats) {
    MML_REQUIRE(ctx, gamma && beta && save_mean && save_invstd, "stem_bn_pool_fwd: null BN pointer");
    MML_LAUNCH(ctx, stem_bn_pool_fwd_kernel<true>, grid, kThreads, 0, st, x, bn, nullptr, nullptr, y, argmax, N, H, W, C, P, Q, 1.0 / (double)rows,
                                                              rows > 1 ? (double)rows / (double)(rows - 1) : 1.0, momentum, eps);
  } else {
    MML_REQUIRE(ctx, shift != nullptr, "stem_bn_pool_fwd: shift is NULL");
    MML_LAUNCH(ctx, stem_bn_pool_fwd_kernel<false>, grid, kThreads, 0, st, x, bn, scale, shift, y, argmax, N, H, W, C, P, Q, 0.0, 0.0, momentum, eps);
  }
  return MML_OK;
}

int mml_stem_bn_pool_bwd(mml_ctx* ctx, const uint16_t* dy, const uint16_t* dy2, const uint8_t* argmax, const uint16_t* x, const float* mean,
                         const float* invstd, const float* gamma, const float* beta, double* bstat, float* dgamma, float* dbeta, uint16_t* dx,
                         int N, int H, int W, int C, int apply, void* stream) {
  MML_REQUIRE(ctx, ctx && dy && argmax && x && mean && invstd && gamma && beta && bstat && dx && N >= 1 && H >= 1 && W >= 1,
              "stem_bn_pool_bwd: bad arguments");
  int rc = check_rows_c(ctx, (int64_t)N * H * W, C);
  if (rc) return rc;
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  MML_REQUIRE(ctx, kThreads % (C / 4) == 0, "stem_bn_pool_bwd: C/4 must divide %d", kThreads);
  const float inv_count = 1.0f / (float)((int64_t)N * H * W);
  cudaStream_t st = (cudaStream_t)stream;
  const int brows = N * ((H + 1) / 2);  // CTAs walk block rows (n, a)
  int g0 = brows < ctx->sm_count * 6 ? brows : ctx->sm_count * 6;
  MML_LAUNCH(ctx, stem_bn_pool_bwd_kernel, g0, kThreads, 0, st, dy, dy2, argmax, x, mean, invstd, gamma, beta, bstat, dx, N, H, W, C, P, Q);
  // apply == 0: dx keeps g and bstat the two sums -- the caller folds pass 2 into the stem weight gradient (mml_stem_wgrad_bn), the only
  // consumer of this layer's dx
  if (!apply) return MML_OK;
  // pass 2 in place over the buffer that now holds g: dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat))
  const long long n8 = (long long)N * H * W * (C / 8);
  MML_LAUNCH(ctx, bn_bwd_apply_kernel, wave_cap(ctx, stream_grid(ctx, n8, kU), bn_bwd_apply_kernel, st), kThreads, 0, st, dx, x, mean, invstd, gamma, bstat, inv_count, dgamma, dbeta, dx, n8, C / 8);
  return MML_OK;
}

int mml_avgpool_fwd(mml_ctx* ctx, const uint16_t* x, float* y, int N, int HW, int C, void* stream) {
  MML_REQUIRE(ctx, ctx && x && y && N >= 1 && HW >= 1 && C >= 8 && C % 8 == 0, "avgpool_fwd: bad arguments");
  MML_LAUNCH(ctx, avgpool_fwd_kernel, ew_grid(ctx, (long long)N * (C / 8)), kThreads, 0, (cudaStream_t)stream, x, y, N, HW, C);
  return MML_OK;
}

int mml_avgpool_bwd(mml_ctx* ctx, const float* dy, uint16_t* dx, int N, int HW, int C, void* stream) {
  MML_REQUIRE(ctx, ctx && dy && dx && N >= 1 && HW >= 1 && C >= 8 && C % 8 == 0, "avgpool_bwd: bad arguments");
  MML_LAUNCH(ctx, avgpool_bwd_kernel, ew_grid(ctx, (long long)N * HW * (C / 8)), kThreads, 0, (cudaStream_t)stream, dy, dx, N, HW, C);
  return MML_OK;
}

}  // extern "C"
