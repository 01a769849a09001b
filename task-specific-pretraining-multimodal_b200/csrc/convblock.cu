// convblock.cu -- the kernels the ConvBlock encoders need beyond the ResNet ones:
// MML_Suite/models/avmnist.py:34-185 (MNISTAudio / MNISTImage) and models/conv.py:16-59 (ConvBlock = Conv2d 3x3 + bias ->
// BatchNorm2d -> ReLU, twice), configs/avmnist/centralised/train_avmnist.yaml.
//
//   * the first convolution of each encoder has ONE input channel (K = 9 taps): like the ResNet stem it is HBM-bound, so it is a
//     SIMT kernel that reads the ORIGINAL fp32 input, applies the missing-modality mask on load (data/base_dataset.py:71, mask_mul),
//     and writes NHWC bf16 with the channel count padded to 64 (channels >= K_out are exact zeros), plus BatchNorm partial sums;
//     its weight gradient is a two-stage fixed-order reduction (deterministic);
//   * the 32 -> 32 / 32 -> 64 / 64 -> 64 convolutions run on the tcgen05 kernels of conv_tc.cu with channels padded to 64 (the
//     padded weight rows / columns live in the flat parameter buffer and stay exactly zero);
//   * nn.MaxPool2d(kernel = stride = k, no padding) forward / backward, with an optional fp32 output in NCHW-flatten order -- the
//     order nn.Flatten gives the following nn.Linear (avmnist.py:79-86,154-161);
//   * the convolution bias is NOT added to the stored tensor: a per-channel constant in front of a training-mode BatchNorm cancels
//     in the normalised output and in every gradient (its own gradient is identically zero); it only shifts the batch mean, so it is
//     added where the mean is used: the running-mean update and the eval-mode shift (bn_act.cu, conv_bias arguments).
#include "mml_common.cuh"
#include "mml_ctx.h"

using namespace mml;

namespace {

constexpr int kT = 256;
constexpr int kCP = 64;  // padded channel count of the output

// ---------------------------------------------------------------------------------------------------------------
// y[b,h,w,k] = sum_{r,s} w[k][r][s] * (x*mask)[b, h+r-1, w+s-1]     (3x3, stride 1, pad 1, C_in = 1), K in {8, 16, 32, 64}
// thread = (8-channel group cg < K/8, pixel of a row); a CTA walks image rows (b, h); channels >= K of the 64 are written as zeros
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf16_round(float v) { return bf16_lo(pack_bf16x2(v, 0.f)); }

__device__ __forceinline__ void load_taps(float (&in)[9], const float* __restrict__ xb, int h, int wq, int H, int W, bool masked, float mk) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int hh = h + r - 1, ww = wq + c - 1;
      float v = 0.f;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(xb + hh * W + ww);
      in[r * 3 + c] = v;
    }
#pragma unroll
  for (int t = 0; t < 9; ++t) {  // mask after all nine loads are in flight (mask_mul branches on the value)
    float v = in[t];
    if (masked) v = mask_mul(v, mk);  // sample = original * mask, data/base_dataset.py:71
    in[t] = bf16_round(v);            // bf16 operands like every other convolution on the path
  }
}

__global__ void __launch_bounds__(kT)
conv3x3_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ w, uint16_t* __restrict__ y,
                      double* __restrict__ stats, int B, int H, int W, int K) {
  pdl_sync();
  __shared__ float red[kT][17];
  const int ng = K >> 3;                                       // real channel groups (1, 2, 4 or 8)
  const int cg = (int)threadIdx.x % ng, ws = (int)threadIdx.x / ng, nw = kT / ng;
  float wk[8][9];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) wk[j][t] = bf16_round(__ldg(w + (cg * 8 + j) * 9 + t));
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int row = blockIdx.x; row < B * H; row += gridDim.x) {
    const int b = row / H, h = row - b * H;
    const float mk = mask ? __ldg(mask + b) : 1.0f;
    const float* xb = x + (size_t)b * H * W;
    for (int wq = ws; wq < W; wq += nw) {
      float in[9];
      load_taps(in, xb, h, wq, H, W, mask != nullptr, mk);
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) a = fmaf(wk[j][t], in[t], a);
        o[j] = a;
      }
      uint4 pk;
      pk.x = pack_bf16x2(o[0], o[1]), pk.y = pack_bf16x2(o[2], o[3]), pk.z = pack_bf16x2(o[4], o[5]), pk.w = pack_bf16x2(o[6], o[7]);
      uint16_t* dst = y + ((size_t)row * W + wq) * kCP;
      *reinterpret_cast<uint4*>(dst + cg * 8) = pk;
      for (int z = cg + ng; z < 8; z += ng) *reinterpret_cast<uint4*>(dst + z * 8) = zero;  // padded channels
      const float r8[8] = {bf16_lo(pk.x), bf16_hi(pk.x), bf16_lo(pk.y), bf16_hi(pk.y), bf16_lo(pk.z), bf16_hi(pk.z), bf16_lo(pk.w), bf16_hi(pk.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += r8[j];
        q[j] = fmaf(r8[j], r8[j], q[j]);
      }
    }
  }
  if (stats != nullptr) {  // BatchNorm partials of the STORED values; the padded channels contribute exact zeros (the caller zeroed them)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[threadIdx.x][j] = s[j];
      red[threadIdx.x][8 + j] = q[j];
    }
    __syncthreads();
    if ((int)threadIdx.x < K) {
      const int c = threadIdx.x, g = c >> 3, j = c & 7;
      float a = 0.f, b2 = 0.f;
      for (int t = g; t < kT; t += ng) {
        a += red[t][j];
        b2 += red[t][8 + j];
      }
      stat_add(stats, kCP, blockIdx.x, c, a, b2);
    }
  }
}

// dW[k][t] = sum_px dy[px][k] * (x*mask)[px + tap t]: per-thread 8 x 9 accumulators over the CTA's rows, warp-shuffle + shared-memory
// reduction to ONE partial per CTA (ws[cta][K*9]); the second kernel adds the partials in a fixed order (deterministic)
__global__ void __launch_bounds__(kT)
conv3x3_c1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ mask, const uint16_t* __restrict__ dy, float* __restrict__ ws,
                        int B, int H, int W, int K) {
  pdl_sync();
  __shared__ float red[kT / 32][8 * 72];  // [warp][cg][ch*9+t]  (18 KB)
  const int ng = K >> 3;
  const int cg = (int)threadIdx.x % ng, wsl = (int)threadIdx.x / ng, nw = kT / ng;
  float acc[8][9];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[j][t] = 0.f;
  for (int row = blockIdx.x; row < B * H; row += gridDim.x) {
    const int b = row / H, h = row - b * H;
    const float mk = mask ? __ldg(mask + b) : 1.0f;
    const float* xb = x + (size_t)b * H * W;
    for (int wq = wsl; wq < W; wq += nw) {
      const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(dy + ((size_t)row * W + wq) * kCP + cg * 8));
      float in[9];
      load_taps(in, xb, h, wq, H, W, mask != nullptr, mk);
      const float g[8] = {bf16_lo(g4.x), bf16_hi(g4.x), bf16_lo(g4.y), bf16_hi(g4.y), bf16_lo(g4.z), bf16_hi(g4.z), bf16_lo(g4.w), bf16_hi(g4.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[j][t] = fmaf(g[j], in[t], acc[j][t]);
    }
  }
  // lanes of a warp with the same (lane % ng) share a channel group: butterfly over the other lane bits
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float v = acc[j][t];
      for (int o = 16; o >= ng; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[j][t] = v;
    }
  if (lane < ng) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int t = 0; t < 9; ++t) red[warp][lane * 72 + j * 9 + t] = acc[j][t];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * 9; i += kT) {  // i = (cg*8 + j)*9 + t == cg*72 + j*9 + t
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < kT / 32; ++wv) v += red[wv][i];
    ws[(size_t)blockIdx.x * (K * 9) + i] = v;
  }
}

// out[i] = sum_p ws[p][i] in a FIXED order: one warp per output, lane l adds parts l, l+32, ... then the lanes are combined by a butterfly
__global__ void __launch_bounds__(256) partial_sum_kernel(const float* __restrict__ ws, int parts, int n, float* __restrict__ out) {
  pdl_sync();
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  float a = 0.f;
  for (int p = lane; p < parts; p += 32) a += ws[(size_t)p * n + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[i] = a;
}

// ---------------------------------------------------------------------------------------------------------------
// nn.MaxPool2d(kernel_size = k) (stride = k, no padding, floor): P = H / k, Q = W / k.  argmax = window position r*k + s of the
// FIRST maximum (torch tie rule), NaN wins.  Output NHWC bf16 and / or fp32 in NCHW-flatten order ([B][C*P*Q], nn.Flatten).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT)
maxpool_k_fwd_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, float* __restrict__ yflat, uint8_t* __restrict__ amax, int B, int H,
                     int W, int C, int k, int P, int Q) {
  pdl_sync();
  const int c8 = C >> 3;
  const int cg = (int)threadIdx.x % c8, qs = (int)threadIdx.x / c8, nq = kT / c8;
  for (int row = blockIdx.x; row < B * P; row += gridDim.x) {
    const int b = row / P, p = row - b * P;
    for (int q = qs; q < Q; q += nq) {
      float best[8];
      int idx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) best[j] = -INFINITY, idx[j] = -1;
      for (int r = 0; r < k; ++r)
        for (int s2 = 0; s2 < k; ++s2) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + (p * k + r)) * W + (q * k + s2)) * C + cg * 8));
          const float v[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (idx[j] < 0 || v[j] > best[j] || v[j] != v[j]) best[j] = v[j], idx[j] = r * k + s2;  // torch: (val > maxval) || isnan(val)
        }
      const size_t o = (((size_t)b * P + p) * Q + q) * C + cg * 8;
      if (y != nullptr) {
        uint4 pk;
        pk.x = pack_bf16x2(best[0], best[1]), pk.y = pack_bf16x2(best[2], best[3]), pk.z = pack_bf16x2(best[4], best[5]), pk.w = pack_bf16x2(best[6], best[7]);
        *reinterpret_cast<uint4*>(y + o) = pk;
      }
      if (yflat != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) yflat[(size_t)b * C * P * Q + ((size_t)(cg * 8 + j) * P + p) * Q + q] = best[j];
      }
      uint2 a;
      a.x = (uint32_t)idx[0] | ((uint32_t)idx[1] << 8) | ((uint32_t)idx[2] << 16) | ((uint32_t)idx[3] << 24);
      a.y = (uint32_t)idx[4] | ((uint32_t)idx[5] << 8) | ((uint32_t)idx[6] << 16) | ((uint32_t)idx[7] << 24);
      *reinterpret_cast<uint2*>(amax + o) = a;
    }
  }
}

// dx[b,h,w,c] = dy[b,h/k,w/k,c] where (h%k)*k + w%k is that window's argmax, else 0 (also for the rows / columns the floor drops)
__global__ void __launch_bounds__(kT)
maxpool_k_bwd_kernel(const uint16_t* __restrict__ dy, const float* __restrict__ dyflat, const uint8_t* __restrict__ amax, uint16_t* __restrict__ dx,
                     int B, int H, int W, int C, int k, int P, int Q) {
  pdl_sync();
  const int c8 = C >> 3;
  const int cg = (int)threadIdx.x % c8, wsl = (int)threadIdx.x / c8, nw = kT / c8;
  for (int row = blockIdx.x; row < B * H; row += gridDim.x) {
    const int b = row / H, h = row - b * H;
    const int p = h / k, r = h - p * k;
    for (int w = wsl; w < W; w += nw) {
      const int q = w / k, s2 = w - q * k;
      uint4 out = make_uint4(0, 0, 0, 0);
      if (p < P && q < Q) {
        const size_t o = (((size_t)b * P + p) * Q + q) * C + cg * 8;
        const uint2 am = __ldg(reinterpret_cast<const uint2*>(amax + o));
        float g[8];
        if (dy != nullptr) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + o));
          g[0] = bf16_lo(u.x), g[1] = bf16_hi(u.x), g[2] = bf16_lo(u.y), g[3] = bf16_hi(u.y);
          g[4] = bf16_lo(u.z), g[5] = bf16_hi(u.z), g[6] = bf16_lo(u.w), g[7] = bf16_hi(u.w);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = __ldg(dyflat + (size_t)b * C * P * Q + ((size_t)(cg * 8 + j) * P + p) * Q + q);
        }
        const uint32_t want = (uint32_t)(r * k + s2);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = ((((j < 4 ? am.x : am.y) >> (8 * (j & 3))) & 0xFFu) == want) ? g[j] : 0.f;
        out.x = pack_bf16x2(v[0], v[1]), out.y = pack_bf16x2(v[2], v[3]), out.z = pack_bf16x2(v[4], v[5]), out.w = pack_bf16x2(v[6], v[7]);
      }
      *reinterpret_cast<uint4*>(dx + ((size_t)row * W + w) * C + cg * 8) = out;
    }
  }
}

// A convolution bias in front of a BatchNorm never reaches the normalised output (it shifts the batch mean and cancels) -- the conv
// kernels therefore never add it.  It is visible in exactly two places, both per-channel: the running mean of train mode
// (mean(conv + b) = mean(conv) + b) and the folded coefficients of eval mode (shift += b * scale).
__global__ void bn_conv_bias_fold_kernel(const float* __restrict__ bias, int C, float momentum, float* __restrict__ running_mean,
                                         const float* __restrict__ scale, float* __restrict__ shift) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float b = bias[c];
  if (running_mean != nullptr) running_mean[c] = fmaf(momentum, b, running_mean[c]);
  if (shift != nullptr) shift[c] = fmaf(b, scale[c], shift[c]);
}

int row_grid(const mml_ctx* ctx, int rows) { return rows < ctx->sm_count * 8 ? (rows < 1 ? 1 : rows) : ctx->sm_count * 8; }

}  // namespace

extern "C" {

int mml_conv3x3_c1_fprop(mml_ctx* ctx, const float* x, const float* mask, const float* w, uint16_t* y, double* stats, int B, int H, int W, int K,
                         void* stream) {
  MML_REQUIRE(ctx, ctx && x && w && y && B >= 1 && H >= 1 && W >= 1, "conv3x3_c1_fprop: bad arguments");
  MML_REQUIRE(ctx, K == 8 || K == 16 || K == 32 || K == 64, "conv3x3_c1: output channels must be 8, 16, 32 or 64 (got %d)", K);
  MML_LAUNCH(ctx, conv3x3_c1_fwd_kernel, row_grid(ctx, B * H), kT, 0, (cudaStream_t)stream, x, mask, w, y, stats, B, H, W, K);
  return MML_OK;
}

int64_t mml_conv3x3_c1_wgrad_workspace(const mml_ctx* ctx, int B, int H, int K) {
  if (!ctx || B < 1 || H < 1 || K < 1) return 0;
  const int ctas = B * H < ctx->sm_count * 4 ? B * H : ctx->sm_count * 4;
  return (int64_t)ctas * K * 9 * sizeof(float);
}

int mml_conv3x3_c1_wgrad(mml_ctx* ctx, const float* x, const float* mask, const uint16_t* dy, float* dw, float* workspace, int64_t workspace_bytes,
                         int B, int H, int W, int K, void* stream) {
  MML_REQUIRE(ctx, ctx && x && dy && dw && workspace && B >= 1 && H >= 1 && W >= 1, "conv3x3_c1_wgrad: bad arguments");
  MML_REQUIRE(ctx, K == 8 || K == 16 || K == 32 || K == 64, "conv3x3_c1: output channels must be 8, 16, 32 or 64 (got %d)", K);
  MML_REQUIRE(ctx, workspace_bytes >= mml_conv3x3_c1_wgrad_workspace(ctx, B, H, K), "conv3x3_c1_wgrad: workspace too small");
  const int ctas = B * H < ctx->sm_count * 4 ? B * H : ctx->sm_count * 4;
  cudaStream_t st = (cudaStream_t)stream;
  MML_LAUNCH(ctx, conv3x3_c1_wgrad_kernel, ctas, kT, 0, st, x, mask, dy, workspace, B, H, W, K);
  MML_LAUNCH(ctx, partial_sum_kernel, (K * 9 + 7) / 8, 256, 0, st, (const float*)workspace, ctas, K * 9, dw);
  return MML_OK;
}

int mml_bn_conv_bias_fold(mml_ctx* ctx, const float* conv_bias, int C, float momentum, float* running_mean, const float* scale, float* shift,
                          void* stream) {
  MML_REQUIRE(ctx, ctx && conv_bias && C >= 1 && (running_mean || (scale && shift)), "bn_conv_bias_fold: bad arguments");
  MML_LAUNCH(ctx, bn_conv_bias_fold_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, conv_bias, C, momentum, running_mean, scale, shift);
  return MML_OK;
}

int mml_maxpool_k_fwd(mml_ctx* ctx, const uint16_t* x, uint16_t* y, float* y_flat_nchw, uint8_t* argmax, int B, int H, int W, int C, int k,
                      void* stream) {
  MML_REQUIRE(ctx, ctx && x && argmax && (y || y_flat_nchw) && B >= 1 && k >= 1 && k <= 15 && H >= k && W >= k, "maxpool_k_fwd: bad arguments");
  MML_REQUIRE(ctx, C >= 8 && C % 8 == 0 && kT % (C / 8) == 0, "maxpool_k: channel count %d unsupported", C);
  const int P = H / k, Q = W / k;
  MML_LAUNCH(ctx, maxpool_k_fwd_kernel, row_grid(ctx, B * P), kT, 0, (cudaStream_t)stream, x, y, y_flat_nchw, argmax, B, H, W, C, k, P, Q);
  return MML_OK;
}

int mml_maxpool_k_bwd(mml_ctx* ctx, const uint16_t* dy, const float* dy_flat_nchw, const uint8_t* argmax, uint16_t* dx, int B, int H, int W, int C,
                      int k, void* stream) {
  MML_REQUIRE(ctx, ctx && argmax && dx && ((dy != nullptr) != (dy_flat_nchw != nullptr)) && B >= 1 && k >= 1 && k <= 15 && H >= k && W >= k,
              "maxpool_k_bwd: bad arguments");
  MML_REQUIRE(ctx, C >= 8 && C % 8 == 0 && kT % (C / 8) == 0, "maxpool_k: channel count %d unsupported", C);
  MML_LAUNCH(ctx, maxpool_k_bwd_kernel, row_grid(ctx, B * H), kT, 0, (cudaStream_t)stream, dy, dy_flat_nchw, argmax, dx, B, H, W, C, k, H / k, W / k);
  return MML_OK;
}

}  // extern "C"
