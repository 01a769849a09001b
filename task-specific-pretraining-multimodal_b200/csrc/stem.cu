// stem.cu -- the ResNet stem convolution (7x7, stride 2, pad 3, C_in = 1 -> 64) forward and weight gradient.
//
// Reference: MML_Suite/models/msa/networks/resnet.py:137 (self.conv1) applied at :205 to the (already masked) fp32
// input, and data/base_dataset.py:71 (sample = original * mask) which is fused here: the kernel reads the ORIGINAL
// fp32 input and multiplies by the per-sample mask on load (a true fp32 multiply, bit-identical to mml_mask_apply_f32).
// K = 49 is too thin for tcgen05 tiles and the op is bandwidth/latency bound on its 64-channel bf16 output, so this is
// a SIMT fp32 kernel: a CTA owns a 4 x 28 output-pixel tile and all 64 channels, the input patch and the 64x49 filter
// live in shared memory, every thread accumulates 4 pixels x 8 channels in registers.  The epilogue stores bf16 NHWC
// and accumulates the BatchNorm sum / sum of squares of the stored values (fp64 atomics).
#include "mml_common.cuh"
#include "mml_ctx.h"

using namespace mml;

namespace {

constexpr int TP = 4;    // output rows per tile
constexpr int TQ = 28;   // output cols per tile
constexpr int kStemThreads = (TP * TQ / 4) * 8;  // 28 pixel groups x 8 channel groups = 224
constexpr int PATCH_H = 2 * TP + 5;              // 13
constexpr int PATCH_W = 2 * TQ + 5;              // 61
constexpr int PATCH_LD = 64;

struct StemGeom {
  int B, H, W, P, Q, tiles_p, tiles_q;
};

__host__ __device__ inline StemGeom stem_geom(int B, int H, int W) {
  StemGeom g;
  g.B = B, g.H = H, g.W = W;
  g.P = (H + 6 - 7) / 2 + 1;
  g.Q = (W + 6 - 7) / 2 + 1;
  g.tiles_p = (g.P + TP - 1) / TP;
  g.tiles_q = (g.Q + TQ - 1) / TQ;
  return g;
}

__device__ __forceinline__ void load_patch(float* patch, const float* __restrict__ x, const float* __restrict__ mask, const StemGeom& g,
                                           int b, int p0, int q0) {
  const float m = mask ? mask[b] : 1.0f;
  const int h0 = 2 * p0 - 3, w0 = 2 * q0 - 3;
  const float* img = x + (size_t)b * g.H * g.W;
  for (int i = threadIdx.x; i < PATCH_H * PATCH_LD; i += blockDim.x) {
    const int r = i / PATCH_LD, c = i - r * PATCH_LD;
    const int h = h0 + r, w = w0 + c;
    float v = 0.f;
    if (c < PATCH_W && h >= 0 && h < g.H && w >= 0 && w < g.W) {
      v = img[(size_t)h * g.W + w];
      if (mask) v = mask_mul(v, m);  // base_dataset.py:71
    }
    patch[i] = v;
  }
}

__global__ void __launch_bounds__(kStemThreads)
stem_fprop_kernel(const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ w, uint16_t* __restrict__ y,
                  double* __restrict__ stats, StemGeom g) {
  __shared__ __align__(16) float wsm[49][64];          // [tap][k]
  __shared__ __align__(16) float patch[PATCH_H * PATCH_LD];
  __shared__ float red[TP * TQ / 4][64][2];

  int tile = blockIdx.x;
  const int tq = tile % g.tiles_q;
  tile /= g.tiles_q;
  const int tp = tile % g.tiles_p;
  const int b = tile / g.tiles_p;
  const int p0 = tp * TP, q0 = tq * TQ;

  for (int i = threadIdx.x; i < 64 * 49; i += blockDim.x) {
    const int k = i / 49, t = i - k * 49;
    wsm[t][k] = w[i];
  }
  load_patch(patch, x, mask, g, b, p0, q0);
  __syncthreads();

  const int cgrp = threadIdx.x & 7;
  const int pgrp = threadIdx.x >> 3;       // 0..27
  const int prow = pgrp / (TQ / 4);        // 0..3
  const int pq4 = (pgrp % (TQ / 4)) * 4;   // 0,4,..,24
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll 1
  for (int r = 0; r < 7; ++r) {
    const float* prow_ptr = patch + (2 * prow + r) * PATCH_LD + 2 * pq4;
    float xin[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) xin[i] = prow_ptr[i];
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const float4 wa = *reinterpret_cast<const float4*>(&wsm[r * 7 + s][cgrp * 8]);
      const float4 wb = *reinterpret_cast<const float4*>(&wsm[r * 7 + s][cgrp * 8 + 4]);
      const float wk[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xin[2 * i + s], wk[j], acc[i][j]);
    }
  }

  const int p = p0 + prow;
  float sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + pq4 + i;
    if (p < g.P && q < g.Q) {
      uint4 o;
      o.x = pack_bf16x2(acc[i][0], acc[i][1]);
      o.y = pack_bf16x2(acc[i][2], acc[i][3]);
      o.z = pack_bf16x2(acc[i][4], acc[i][5]);
      o.w = pack_bf16x2(acc[i][6], acc[i][7]);
      *reinterpret_cast<uint4*>(y + ((((size_t)b * g.P + p) * g.Q + q) * 64 + cgrp * 8)) = o;
      const float v[8] = {bf16_lo(o.x), bf16_hi(o.x), bf16_lo(o.y), bf16_hi(o.y), bf16_lo(o.z), bf16_hi(o.z), bf16_lo(o.w), bf16_hi(o.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sum[j] += v[j];
        sq[j] = fmaf(v[j], v[j], sq[j]);
      }
    }
  }
  if (stats) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[pgrp][cgrp * 8 + j][0] = sum[j];
      red[pgrp][cgrp * 8 + j][1] = sq[j];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      float a = 0.f, c = 0.f;
      for (int t = 0; t < TP * TQ / 4; ++t) {
        a += red[t][threadIdx.x][0];
        c += red[t][threadIdx.x][1];
      }
      stat_add(stats, 64, blockIdx.x, threadIdx.x, a, c);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad: dw[k][r][s] = sum_{b,p,q} dy[b,p,q,k] * xm[b, 2p+r-3, 2q+s-3]
// thread = (2 output channels, one filter row r) -> 14 register accumulators, persistent over tiles; per-CTA
// partials go to a workspace and a second kernel reduces them in a fixed order (deterministic).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStemThreads)
stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ mask, const uint16_t* __restrict__ dy, float* __restrict__ ws,
                  StemGeom g, int total_tiles) {
  __shared__ __align__(16) float patch[PATCH_H * PATCH_LD];
  __shared__ __align__(16) float dys[TP * TQ][64 + 2];
  const int kg = threadIdx.x & 31;  // channels 2*kg, 2*kg+1
  const int r = threadIdx.x >> 5;   // 0..6
  float acc[2][7];
#pragma unroll
  for (int s = 0; s < 7; ++s) acc[0][s] = acc[1][s] = 0.f;

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    int t = tile;
    const int tq = t % g.tiles_q;
    t /= g.tiles_q;
    const int tp = t % g.tiles_p;
    const int b = t / g.tiles_p;
    const int p0 = tp * TP, q0 = tq * TQ;
    __syncthreads();  // previous iteration done with smem
    load_patch(patch, x, mask, g, b, p0, q0);
    for (int i = threadIdx.x; i < TP * TQ * 8; i += blockDim.x) {
      const int pix = i >> 3, c8 = i & 7;
      const int p = p0 + pix / TQ, q = q0 + pix % TQ;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (p < g.P && q < g.Q) v = __ldg(reinterpret_cast<const uint4*>(dy + ((((size_t)b * g.P + p) * g.Q + q) * 64 + c8 * 8)));
      float* d = &dys[pix][c8 * 8];
      d[0] = bf16_lo(v.x), d[1] = bf16_hi(v.x), d[2] = bf16_lo(v.y), d[3] = bf16_hi(v.y);
      d[4] = bf16_lo(v.z), d[5] = bf16_hi(v.z), d[6] = bf16_lo(v.w), d[7] = bf16_hi(v.w);
    }
    __syncthreads();
#pragma unroll 1
    for (int pr = 0; pr < TP; ++pr) {
      const float* xrow = patch + (2 * pr + r) * PATCH_LD;
#pragma unroll 1
      for (int q4 = 0; q4 < TQ; q4 += 4) {
        float xin[13];
#pragma unroll
        for (int i = 0; i < 13; ++i) xin[i] = xrow[2 * q4 + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 d = *reinterpret_cast<const float2*>(&dys[pr * TQ + q4 + i][2 * kg]);
#pragma unroll
          for (int s = 0; s < 7; ++s) {
            acc[0][s] = fmaf(d.x, xin[2 * i + s], acc[0][s]);
            acc[1][s] = fmaf(d.y, xin[2 * i + s], acc[1][s]);
          }
        }
      }
    }
  }
  float* out = ws + (size_t)blockIdx.x * (64 * 49);
#pragma unroll
  for (int s = 0; s < 7; ++s) {
    out[(2 * kg) * 49 + r * 7 + s] = acc[0][s];
    out[(2 * kg + 1) * 49 + r * 7 + s] = acc[1][s];
  }
}

__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ ws, int parts, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 49) return;
  float a = 0.f;
  for (int p = 0; p < parts; ++p) a += ws[(size_t)p * (64 * 49) + i];
  dw[i] = a;
}

int stem_wgrad_ctas(const mml_ctx* ctx, int total_tiles) {
  int n = ctx->sm_count * 2;
  if (n > total_tiles) n = total_tiles;
  return n < 1 ? 1 : n;
}

}  // namespace

extern "C" {

int mml_stem_fprop(mml_ctx* ctx, const float* x, const float* mask, const float* w, uint16_t* y, double* stats, int B,
                   int H, int W, void* stream) {
  MML_REQUIRE(ctx, ctx && x && w && y, "stem_fprop: null pointer");
  MML_REQUIRE(ctx, B >= 1 && H >= 1 && W >= 1, "stem_fprop: bad dims");
  const StemGeom g = stem_geom(B, H, W);
  stem_fprop_kernel<<<B * g.tiles_p * g.tiles_q, kStemThreads, 0, (cudaStream_t)stream>>>(x, mask, w, y, stats, g);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int64_t mml_stem_wgrad_workspace(const mml_ctx* ctx, int B, int H, int W) {
  if (!ctx) return 0;
  const StemGeom g = stem_geom(B, H, W);
  return (int64_t)stem_wgrad_ctas(ctx, B * g.tiles_p * g.tiles_q) * 64 * 49 * sizeof(float);
}

int mml_stem_wgrad(mml_ctx* ctx, const float* x, const float* mask, const uint16_t* dy, float* dw, float* workspace,
                   int64_t workspace_bytes, int B, int H, int W, void* stream) {
  MML_REQUIRE(ctx, ctx && x && dy && dw && workspace, "stem_wgrad: null pointer");
  MML_REQUIRE(ctx, B >= 1 && H >= 1 && W >= 1, "stem_wgrad: bad dims");
  MML_REQUIRE(ctx, workspace_bytes >= mml_stem_wgrad_workspace(ctx, B, H, W), "stem_wgrad: workspace too small");
  const StemGeom g = stem_geom(B, H, W);
  const int total = B * g.tiles_p * g.tiles_q;
  const int ctas = stem_wgrad_ctas(ctx, total);
  cudaStream_t st = (cudaStream_t)stream;
  stem_wgrad_kernel<<<ctas, kStemThreads, 0, st>>>(x, mask, dy, workspace, g, total);
  MML_LAUNCHED(ctx);
  stem_wgrad_reduce_kernel<<<(64 * 49 + 255) / 256, 256, 0, st>>>(workspace, ctas, dw);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

}  // extern "C"
