// gated_fusion.cu -- the memory-bound kernels of the MMIMDb gated late-fusion step (BASELINE config 3).
//
// The six Linear layers of that model are GEMMs with M = batch (128 = exactly one UMMA M tile) and run on the tcgen05
// kernels of conv_tc.cu as 1x1 "convolutions" over [B,1,1,C] activations (fprop / dgrad / wgrad, bf16 operands, fp32
// accumulation).  Everything BETWEEN the GEMMs is here, fused so that each activation makes one trip through HBM/L2:
//
//   bn1d_fwd   BatchNorm1d over the batch (exact two-pass mean / variance in fp32) with the PRODUCER of its input fused in:
//                INPUT   v = x * mask[b]                      (missing-modality mask, base_dataset.py:71; mmimdb.py:80)
//                GATED   v = g*h1 + (1-g)*h2                  (GMU mix, gated_bimodal.py:59; mmimdb.py:38)
//                MAXOUT  v = max(pre[:, :C], pre[:, C:]) * keep/(1-p)   (maxout.py:37-41 + Dropout; mmimdb.py:41,44)
//              and the CONSUMER's operand format fused out: bf16 rows (next GEMM's A operand) or fp32 (final Linear).
//   bn1d_bwd   dgamma / dbeta / dx of the same, routed back through the producer (max routing + dropout, or plain dz).
//   gmu_fwd / gmu_bwd   tanh, the per-sample scalar gate sigmoid(w_z . [h1|h2]) and their backward (row reductions).
//   bce_head_fwd / bwd  Linear(H -> classes) + BCEWithLogits (mean over B x classes) + sigmoid > threshold, and backward.
//
// Thread layout of the column kernels: 32 columns x 32 row-groups per CTA; a warp reads 32 consecutive columns of one row
// (128 B fp32 / 64 B bf16 segments), the batch reduction is a shared-memory tree over the 32 row-groups, so the statistics
// never leave the CTA (no atomics, deterministic).
#include <cuda_bf16.h>

#include "mml_common.cuh"
#include "mml_ctx.h"

namespace {
using namespace mml;

constexpr int kCols = 32;
constexpr int kRowGroups = 32;

__device__ __forceinline__ float bf16_val(uint16_t v) { return __uint_as_float((uint32_t)v << 16); }
__device__ __forceinline__ uint16_t to_bf16(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }

template <int MODE>
__device__ __forceinline__ float produce(const mml_bn1d_desc& d, int b, int c) {
  if (MODE == MML_BN1D_INPUT) {
    const float x = d.x[(size_t)b * d.ldx + c];
    return d.mask ? mask_mul(x, d.mask[b]) : x;
  } else if (MODE == MML_BN1D_GATED) {
    const size_t i = (size_t)b * d.C + c;
    if (d.gate == nullptr) return d.mix_a * d.h1[i] + d.mix_b * d.h2[i];
    const float g = d.gate[b];
    return g * d.h1[i] + (1.f - g) * d.h2[i];
  } else if (MODE == MML_BN1D_MAX2) {
    const size_t i = (size_t)b * d.C + c;
    return fmaxf(d.h1[i], d.h2[i]);
  } else {
    const size_t i = (size_t)b * 2 * d.C + c;
    float v = fmaxf(bf16_val(d.pre[i]), bf16_val(d.pre[i + d.C]));
    if (d.keep) v = d.keep[(size_t)b * d.C + c] ? v * d.keep_scale : 0.f;
    return v;
  }
}

// column sums over the 32 row-groups; result valid in every thread of the column
__device__ __forceinline__ float column_total(float v, float (*red)[kCols + 1], float* out, int tx, int ty) {
  red[ty][tx] = v;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kRowGroups; ++k) t += red[k][tx];
    out[tx] = t;
  }
  __syncthreads();
  return out[tx];
}

// REG = true: B <= 4 * kRowGroups, every thread keeps its (at most 4) produced values in registers -> one trip to memory
// instead of three (the statistics passes re-run the producer otherwise).
template <int MODE, bool REG>
__global__ void __launch_bounds__(kCols* kRowGroups) bn1d_fwd_kernel(const mml_bn1d_desc d) {
  __shared__ float red[kRowGroups][kCols + 1];
  __shared__ float tot[kCols];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * kCols + tx;
  const bool valid = c < d.C;
  const int cc = valid ? c : d.C - 1;  // out-of-range columns shadow the last one and never store
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (REG) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ty + k * kRowGroups < d.B) v[k] = produce<MODE>(d, ty + k * kRowGroups, cc);
  }
  float mean, inv;
  if (d.train) {
    float s = 0.f;
    if (REG) {
      s = (v[0] + v[1]) + (v[2] + v[3]);
    } else {
      for (int b = ty; b < d.B; b += kRowGroups) s += produce<MODE>(d, b, cc);
    }
    mean = column_total(s, red, tot, tx, ty) / (float)d.B;
    float ss = 0.f;
    if (REG) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ty + k * kRowGroups < d.B) ss += (v[k] - mean) * (v[k] - mean);
    } else {
      for (int b = ty; b < d.B; b += kRowGroups) {
        const float dv = produce<MODE>(d, b, cc) - mean;
        ss += dv * dv;
      }
    }
    const float var = column_total(ss, red, tot, tx, ty) / (float)d.B;  // biased, used for normalisation
    inv = 1.0f / sqrtf(var + d.eps);
    if (ty == 0 && valid) {
      if (d.invstd) d.invstd[c] = inv;
      const float unbiased = d.B > 1 ? var * (float)d.B / (float)(d.B - 1) : var;
      d.running_mean[c] = (1.f - d.momentum) * d.running_mean[c] + d.momentum * mean;
      d.running_var[c] = (1.f - d.momentum) * d.running_var[c] + d.momentum * unbiased;
    }
  } else {
    mean = d.running_mean[cc];
    inv = 1.0f / sqrtf(d.running_var[cc] + d.eps);
  }
  if (!valid) return;
  const float g = d.gamma[c], be = d.beta[c];
  auto emit = [&](int b, float val) {
    const float xh = (val - mean) * inv;
    if (d.xhat) d.xhat[(size_t)b * d.C + c] = xh;
    const float y = xh * g + be;
    if (d.y_bf16) d.y_bf16[(size_t)b * d.ldy + c] = to_bf16(y);
    if (d.y_f32) d.y_f32[(size_t)b * d.C + c] = y;
  };
  if (REG) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ty + k * kRowGroups < d.B) emit(ty + k * kRowGroups, v[k]);
  } else {
    for (int b = ty; b < d.B; b += kRowGroups) emit(b, produce<MODE>(d, b, c));
  }
}

template <int MODE, bool REG>
__global__ void __launch_bounds__(kCols* kRowGroups) bn1d_bwd_kernel(const mml_bn1d_bwd_desc d) {
  __shared__ float red[kRowGroups][kCols + 1];
  __shared__ float tot[kCols];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * kCols + tx;
  const bool valid = c < d.C;
  const int cc = valid ? c : d.C - 1;
  float s1 = 0.f, s2 = 0.f;
  float vdy[4] = {0.f, 0.f, 0.f, 0.f}, vxh[4] = {0.f, 0.f, 0.f, 0.f};
  if (REG) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = ty + k * kRowGroups;
      if (b < d.B) {
        vdy[k] = bf16_val(d.dy[(size_t)b * d.lddy + cc]);
        vxh[k] = d.xhat[(size_t)b * d.C + cc];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) s1 += vdy[k], s2 += vdy[k] * vxh[k];
  } else {
    for (int b = ty; b < d.B; b += kRowGroups) {
      const float dy = bf16_val(d.dy[(size_t)b * d.lddy + cc]);
      s1 += dy;
      s2 += dy * d.xhat[(size_t)b * d.C + cc];
    }
  }
  const float sdy = column_total(s1, red, tot, tx, ty);
  const float sdyx = column_total(s2, red, tot, tx, ty);
  if (!valid) return;
  if (ty == 0) {
    d.dbeta[c] = sdy;
    d.dgamma[c] = sdyx;
  }
  if (MODE == MML_BN1D_INPUT) return;  // the network input needs no gradient
  const float k = d.gamma[c] * d.invstd[c], invB = 1.f / (float)d.B;
  auto emit = [&](int b, float dy, float xh) {
    float dv = k * (dy - sdy * invB - xh * sdyx * invB);
    if (MODE == MML_BN1D_GATED || MODE == MML_BN1D_MAX2) {
      d.dz[(size_t)b * d.C + c] = dv;
    } else {
      if (d.keep) dv = d.keep[(size_t)b * d.C + c] ? dv * d.keep_scale : 0.f;
      const size_t i = (size_t)b * 2 * d.C + c;
      const float a = bf16_val(d.pre[i]), o = bf16_val(d.pre[i + d.C]);
      // torch.max(a, b) backward: the winner takes the gradient, an exact tie splits it evenly
      const float ga = a > o ? dv : (a == o ? 0.5f * dv : 0.f);
      d.dpre[i] = to_bf16(ga);
      d.dpre[i + d.C] = to_bf16(dv - ga);
    }
  };
  if (REG) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (ty + q * kRowGroups < d.B) emit(ty + q * kRowGroups, vdy[q], vxh[q]);
  } else {
    for (int b = ty; b < d.B; b += kRowGroups) emit(b, bf16_val(d.dy[(size_t)b * d.lddy + c]), d.xhat[(size_t)b * d.C + c]);
  }
}

// ---- GMU: one warp per sample ---------------------------------------------------------------------------------------
constexpr int kGmuWarps = 8;

__global__ void __launch_bounds__(kGmuWarps * 32) gmu_fwd_kernel(const uint16_t* __restrict__ h1pre, const uint16_t* __restrict__ h2pre,
                                                                 const float* __restrict__ wz, float* __restrict__ h1,
                                                                 float* __restrict__ h2, float* __restrict__ gate, int B, int H) {
  const int row = blockIdx.x * kGmuWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  float acc = 0.f;
#pragma unroll 4
  for (int c = lane; c < H; c += 32) {
    const size_t i = (size_t)row * H + c;
    const float a = tanhf(bf16_val(h1pre[i])), b = tanhf(bf16_val(h2pre[i]));
    h1[i] = a;
    h2[i] = b;
    acc += a * wz[c] + b * wz[H + c];
  }
  acc = warp_sum(acc);
  if (lane == 0) gate[row] = 1.f / (1.f + expf(-acc));
}

__global__ void __launch_bounds__(kGmuWarps * 32) gmu_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ h1,
                                                                 const float* __restrict__ h2, const float* __restrict__ gate,
                                                                 const float* __restrict__ wz, float* __restrict__ dwz,
                                                                 uint16_t* __restrict__ dh1pre, uint16_t* __restrict__ dh2pre, int B, int H) {
  __shared__ float s_dgp[kGmuWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kGmuWarps, row = row0 + warp;
  float dgp = 0.f;
  if (row < B) {
    float dg = 0.f;
#pragma unroll 8
    for (int c = lane; c < H; c += 32) {
      const size_t i = (size_t)row * H + c;
      dg += dz[i] * (h1[i] - h2[i]);
    }
    dg = warp_sum(dg);
    const float g = gate[row];
    dgp = dg * g * (1.f - g);  // gradient at the gate's pre-activation
#pragma unroll 4
    for (int c = lane; c < H; c += 32) {
      const size_t i = (size_t)row * H + c;
      const float a = h1[i], b = h2[i], z = dz[i];
      dh1pre[i] = to_bf16((z * g + dgp * wz[c]) * (1.f - a * a));
      dh2pre[i] = to_bf16((z * (1.f - g) + dgp * wz[H + c]) * (1.f - b * b));
    }
  }
  if (lane == 0) s_dgp[warp] = dgp;
  __syncthreads();
  // d w_z[c] = sum over samples of dgp * [h1|h2][c]: the CTA's rows are summed here, CTAs meet in global atomics
  for (int c = threadIdx.x; c < 2 * H; c += blockDim.x) {
    const float* h = c < H ? h1 : h2;
    const int cc = c < H ? c : c - H;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kGmuWarps; ++r)
      if (row0 + r < B) s += s_dgp[r] * h[(size_t)(row0 + r) * H + cc];
    atomicAdd(dwz + c, s);
  }
}

// ---- MultimodalPooling branches (pooling.py:92-98) --------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_fwd_kernel(const uint16_t* __restrict__ pre_a, const uint16_t* __restrict__ pre_b,
                                                      const float* __restrict__ bias_a, const float* __restrict__ bias_b,
                                                      const uint8_t* __restrict__ keep_a, const uint8_t* __restrict__ keep_b, float keep_scale,
                                                      float* __restrict__ h_a, float* __restrict__ h_b, uint16_t* __restrict__ comb, int B, int H) {
  const size_t n = (size_t)B * H;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % H);
    float a = tanhf(bf16_val(pre_a[i]) + bias_a[c]), b = tanhf(bf16_val(pre_b[i]) + bias_b[c]);
    if (keep_a) a = keep_a[i] ? a * keep_scale : 0.f;
    if (keep_b) b = keep_b[i] ? b * keep_scale : 0.f;
    h_a[i] = a;
    h_b[i] = b;
    if (comb) {  // [h_a | h_b] rows as the bf16 A operand of the attention / gate GEMM (pooling.py:115)
      const size_t r = i / H;
      comb[r * 2 * H + c] = to_bf16(a);
      comb[r * 2 * H + H + c] = to_bf16(b);
    }
  }
}

// column-sliced like the BatchNorm kernels: the bias gradients are batch sums, reduced inside the CTA
__global__ void __launch_bounds__(kCols* kRowGroups) pool_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ h_a,
                                                                    const float* __restrict__ h_b, const uint8_t* __restrict__ keep_a,
                                                                    const uint8_t* __restrict__ keep_b, float keep_scale, int kind, float mix_a,
                                                                    float mix_b, const float* __restrict__ gate, const uint16_t* __restrict__ dcomb,
                                                                    uint16_t* __restrict__ dpre_a, uint16_t* __restrict__ dpre_b,
                                                                    float* __restrict__ dbias_a, float* __restrict__ dbias_b, int B, int H) {
  __shared__ float red[kRowGroups][kCols + 1];
  __shared__ float tot[kCols];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * kCols + tx;
  const bool valid = c < H;
  float sa = 0.f, sb = 0.f;
  if (valid) {
    const float inv_scale = 1.f / keep_scale;
    for (int b = ty; b < B; b += kRowGroups) {
      const size_t i = (size_t)b * H + c;
      const float z = dz[i], a = h_a[i], o = h_b[i];
      float da, db;
      if (kind == 0) {  // max: winner takes the gradient, an exact tie splits it (torch.max backward)
        da = a > o ? z : (a == o ? 0.5f * z : 0.f);
        db = z - da;
      } else if (gate) {  // attention / gated pooling: per-sample mix + the gradient that came back through the gate network
        const float g = gate[b];
        da = z * g;
        db = z * (1.f - g);
      } else {
        da = z * mix_a;
        db = z * mix_b;
      }
      if (dcomb) {
        da += bf16_val(dcomb[(size_t)b * 2 * H + c]);
        db += bf16_val(dcomb[(size_t)b * 2 * H + H + c]);
      }
      // through dropout (h = tanh * scale where kept) and tanh
      const bool ka = keep_a ? keep_a[i] != 0 : true, kb = keep_b ? keep_b[i] != 0 : true;
      const float sc = keep_a ? keep_scale : 1.f, isc = keep_a ? inv_scale : 1.f;
      const float ta = a * isc, tb = o * isc;
      const float ga = ka ? da * sc * (1.f - ta * ta) : 0.f;
      const float gb = kb ? db * sc * (1.f - tb * tb) : 0.f;
      dpre_a[i] = to_bf16(ga);
      dpre_b[i] = to_bf16(gb);
      sa += ga;
      sb += gb;
    }
  }
  const float ta = column_total(sa, red, tot, tx, ty);
  const float tb = column_total(sb, red, tot, tx, ty);
  if (valid && ty == 0) {
    dbias_a[c] = ta;
    dbias_b[c] = tb;
  }
}

// ---- attention / gated pooling head (pooling.py:55-72, 113-126): t = tanh(hid + b0); s = W2 t + b2;
// attention: (att_a, att_b) = softmax(s0, s1) = (g, 1 - g) with g = sigmoid(s0 - s1);  gated: g = sigmoid(s0).  One warp per sample.
__global__ void __launch_bounds__(kGmuWarps * 32) att_fwd_kernel(const uint16_t* __restrict__ hid, const float* __restrict__ b0,
                                                                 const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ t,
                                                                 float* __restrict__ gate, int B, int Hd, int NS) {
  const int row = blockIdx.x * kGmuWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
  for (int k = lane; k < Hd; k += 32) {
    const float v = tanhf(bf16_val(hid[(size_t)row * Hd + k]) + b0[k]);
    t[(size_t)row * Hd + k] = v;
    s0 += v * w2[k];
    if (NS == 2) s1 += v * w2[Hd + k];
  }
  s0 = warp_sum(s0) + b2[0];
  if (NS == 2) s0 -= warp_sum(s1) + b2[1];
  if (lane == 0) gate[row] = 1.f / (1.f + expf(-s0));
}

__global__ void __launch_bounds__(kGmuWarps * 32) att_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ h_a,
                                                                 const float* __restrict__ h_b, const float* __restrict__ gate,
                                                                 const float* __restrict__ t, const float* __restrict__ w2, float* __restrict__ dw2,
                                                                 float* __restrict__ db2, float* __restrict__ db0, uint16_t* __restrict__ dhid, int B,
                                                                 int H, int Hd, int NS) {
  __shared__ float s_ds[kGmuWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kGmuWarps, row = row0 + warp;
  float ds = 0.f;  // gradient at s0 (attention: d s1 = -d s0)
  if (row < B) {
    float dg = 0.f;
#pragma unroll 4
    for (int c = lane; c < H; c += 32) {
      const size_t i = (size_t)row * H + c;
      dg += dz[i] * (h_a[i] - h_b[i]);
    }
    dg = warp_sum(dg);
    const float g = gate[row];
    ds = dg * g * (1.f - g);
#pragma unroll 4
    for (int k = lane; k < Hd; k += 32) {
      const float tv = t[(size_t)row * Hd + k];
      const float wv = NS == 2 ? w2[k] - w2[Hd + k] : w2[k];
      dhid[(size_t)row * Hd + k] = to_bf16(ds * wv * (1.f - tv * tv));
    }
  }
  if (lane == 0) s_ds[warp] = ds;
  __syncthreads();
  // parameter gradients: the CTA's rows are summed here, CTAs meet in global atomics (the buffers are zeroed once per step)
  float bsum = 0.f;
  for (int k = threadIdx.x; k < Hd; k += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int r = 0; r < kGmuWarps; ++r) {
      if (row0 + r < B) {
        const float tv = t[(size_t)(row0 + r) * Hd + k];
        const float wv = NS == 2 ? w2[k] - w2[Hd + k] : w2[k];
        a += s_ds[r] * tv;
        b += s_ds[r] * wv * (1.f - tv * tv);
      }
    }
    atomicAdd(dw2 + k, a);
    if (NS == 2) atomicAdd(dw2 + Hd + k, -a);
    atomicAdd(db0 + k, b);
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int r = 0; r < kGmuWarps; ++r) bsum += s_ds[r];
    atomicAdd(db2, bsum);
    if (NS == 2) atomicAdd(db2 + 1, -bsum);
  }
}

// ---- final Linear + BCE-with-logits ---------------------------------------------------------------------------------
constexpr int kHeadWarps = 8;

__global__ void __launch_bounds__(kHeadWarps * 32) bce_head_fwd_kernel(const float* __restrict__ xn, const float* __restrict__ w,
                                                                      const float* __restrict__ bias, const float* __restrict__ labels,
                                                                      float* __restrict__ logits, float* __restrict__ loss,
                                                                      float* __restrict__ dlogits, uint8_t* __restrict__ pred,
                                                                      float* __restrict__ scratch, float threshold, float grad_scale,
                                                                      int B, int H, int NC) {
  // [NC][H] weights + [kHeadWarps][H] rows.  Both are staged with wide, fully independent loads (one round trip to L2);
  // the 23 dots per sample then run out of shared memory.  (Dots straight from global were a chain of dependent
  // load->FMA pairs: 132 us for 128 samples.)
  extern __shared__ __align__(16) float s_head[];
  float* s_w = s_head;
  float* s_row = s_head + (size_t)NC * H;
  __shared__ float s_loss[kHeadWarps];
  __shared__ bool s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kHeadWarps, row = row0 + warp;
  {
    const int nw4 = NC * H / 4, rows = min(kHeadWarps, B - row0), nr4 = rows * H / 4;
    const float4* gw = reinterpret_cast<const float4*>(w);
    const float4* gx = reinterpret_cast<const float4*>(xn + (size_t)row0 * H);
    float4* sw4 = reinterpret_cast<float4*>(s_w);
    float4* sr4 = reinterpret_cast<float4*>(s_row);
    for (int i0 = 0; i0 < nw4; i0 += 8 * blockDim.x) {
      float4 t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = i0 + k * blockDim.x + threadIdx.x;
        if (i < nw4) t[k] = gw[i];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = i0 + k * blockDim.x + threadIdx.x;
        if (i < nw4) sw4[i] = t[k];
      }
    }
    for (int i = threadIdx.x; i < nr4; i += blockDim.x) sr4[i] = gx[i];
  }
  __syncthreads();
  float row_loss = 0.f;
  if (row < B) {
    const float* xr = s_row + warp * H;
    float mine = 0.f;
    for (int j = 0; j < NC; ++j) {
      const float* wr = s_w + (size_t)j * H;
      float acc = 0.f;
#pragma unroll 4
      for (int c = lane; c < H; c += 32) acc += xr[c] * wr[c];
      acc = warp_sum(acc);
      if (lane == j) mine = acc + bias[j];
    }
    if (lane < NC) {
      const size_t i = (size_t)row * NC + lane;
      logits[i] = mine;
      const float sig = 1.f / (1.f + expf(-mine));
      if (pred) pred[i] = sig > threshold ? 1 : 0;
      if (labels) {
        const float y = labels[i];
        row_loss = fmaxf(mine, 0.f) - mine * y + log1pf(expf(-fabsf(mine)));
        if (dlogits) dlogits[i] = (sig - y) * grad_scale / ((float)B * (float)NC);
      }
    }
    row_loss = warp_sum(row_loss);
  }
  if (!loss) return;
  if (lane == 0) s_loss[warp] = row_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < kHeadWarps; ++k) t += s_loss[k];
    scratch[1 + blockIdx.x] = t;
    __threadfence();
    const unsigned ticket = atomicAdd((unsigned*)scratch, 1u);
    s_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {  // fixed summation order -> the loss is bit-reproducible
    __threadfence();
    float t = 0.f;
    for (unsigned k = 0; k < gridDim.x; ++k) t += ((volatile float*)scratch)[1 + k];
    loss[0] = t / ((float)B * (float)NC);
    *(unsigned*)scratch = 0u;
  }
}

__global__ void __launch_bounds__(kCols * 8) bce_head_bwd_kernel(const float* __restrict__ dl, const float* __restrict__ xn,
                                                                const float* __restrict__ w, float* __restrict__ dw,
                                                                float* __restrict__ db, uint16_t* __restrict__ dxn, int B, int H, int NC) {
  constexpr int kTile = 128;
  __shared__ float s_dl[kTile * 32];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kCols + tx;
  const int c = blockIdx.x * kCols + tx;
  const bool valid = c < H;
  const int cc = valid ? c : H - 1;
  float wreg[32];  // column cc of W
#pragma unroll
  for (int j = 0; j < 32; ++j) wreg[j] = j < NC ? w[(size_t)j * H + cc] : 0.f;
  float accw[4] = {0.f, 0.f, 0.f, 0.f};  // d W[ty + 8 q][c]
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b0 = 0; b0 < B; b0 += kTile) {
    const int rows = min(kTile, B - b0);
    __syncthreads();
    for (int i = tid; i < rows * NC; i += kCols * 8) s_dl[i] = dl[(size_t)b0 * NC + i];
    __syncthreads();
    // d xn[b][c] = sum_j dl[b][j] W[j][c]   (s_dl reads are warp-wide broadcasts)
    for (int r = ty; r < rows; r += 8) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < NC) acc += s_dl[r * NC + j] * wreg[j];
      if (valid) dxn[(size_t)(b0 + r) * H + c] = to_bf16(acc);
    }
    // d W[j][c] += sum_b dl[b][j] xn[b][c]: 16 rows of the xn column per batch of loads
    for (int r0 = 0; r0 < rows; r0 += 16) {
      float xv[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) xv[k] = r0 + k < rows ? xn[(size_t)(b0 + r0 + k) * H + cc] : 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (r0 + k < rows) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = ty + 8 * q;
            if (j < NC) {
              const float dv = s_dl[(r0 + k) * NC + j];
              accw[q] += dv * xv[k];
              accb[q] += dv;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = ty + 8 * q;
    if (j < NC && valid) dw[(size_t)j * H + c] = accw[q];
    if (j < NC && blockIdx.x == 0 && tx == 0) db[j] = accb[q];
  }
}

}  // namespace

extern "C" {

int mml_bn1d_fwd(mml_ctx* ctx, const mml_bn1d_desc* d, void* stream) {
  MML_REQUIRE(ctx, ctx && d, "bn1d_fwd: null ctx/desc");
  MML_REQUIRE(ctx, d->B >= 1 && d->C >= 1 && d->gamma && d->beta && d->running_mean && d->running_var, "bn1d_fwd: bad arguments");
  MML_REQUIRE(ctx, !d->train || d->B > 1, "bn1d_fwd: Expected more than 1 value per channel when training (batch %d)", d->B);
  MML_REQUIRE(ctx, d->y_bf16 || d->y_f32, "bn1d_fwd: no output");
  MML_REQUIRE(ctx, !d->y_bf16 || d->ldy >= d->C, "bn1d_fwd: ldy < C");
  const dim3 block(kCols, kRowGroups), grid((unsigned)mml_ceil_div(d->C, kCols));
  cudaStream_t st = (cudaStream_t)stream;
  const bool reg = d->B <= 4 * kRowGroups;
  switch (d->mode) {
    case MML_BN1D_INPUT:
      MML_REQUIRE(ctx, d->x && d->ldx >= d->C, "bn1d_fwd(INPUT): x / ldx");
      if (reg) bn1d_fwd_kernel<MML_BN1D_INPUT, true><<<grid, block, 0, st>>>(*d);
      else bn1d_fwd_kernel<MML_BN1D_INPUT, false><<<grid, block, 0, st>>>(*d);
      break;
    case MML_BN1D_GATED:
      MML_REQUIRE(ctx, d->h1 && d->h2, "bn1d_fwd(GATED): h1 / h2");
      if (reg) bn1d_fwd_kernel<MML_BN1D_GATED, true><<<grid, block, 0, st>>>(*d);
      else bn1d_fwd_kernel<MML_BN1D_GATED, false><<<grid, block, 0, st>>>(*d);
      break;
    case MML_BN1D_MAX2:
      MML_REQUIRE(ctx, d->h1 && d->h2, "bn1d_fwd(MAX2): h1 / h2");
      if (reg) bn1d_fwd_kernel<MML_BN1D_MAX2, true><<<grid, block, 0, st>>>(*d);
      else bn1d_fwd_kernel<MML_BN1D_MAX2, false><<<grid, block, 0, st>>>(*d);
      break;
    case MML_BN1D_MAXOUT:
      MML_REQUIRE(ctx, d->pre, "bn1d_fwd(MAXOUT): pre");
      if (reg) bn1d_fwd_kernel<MML_BN1D_MAXOUT, true><<<grid, block, 0, st>>>(*d);
      else bn1d_fwd_kernel<MML_BN1D_MAXOUT, false><<<grid, block, 0, st>>>(*d);
      break;
    default:
      return mml_set_error(ctx, MML_ERR_INVALID, "bn1d_fwd: unknown mode %d", d->mode);
  }
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_bn1d_bwd(mml_ctx* ctx, const mml_bn1d_bwd_desc* d, void* stream) {
  MML_REQUIRE(ctx, ctx && d, "bn1d_bwd: null ctx/desc");
  MML_REQUIRE(ctx, d->B > 1 && d->C >= 1 && d->dy && d->lddy >= d->C && d->xhat && d->dgamma && d->dbeta, "bn1d_bwd: bad arguments");
  const dim3 block(kCols, kRowGroups), grid((unsigned)mml_ceil_div(d->C, kCols));
  cudaStream_t st = (cudaStream_t)stream;
  const bool reg = d->B <= 4 * kRowGroups;
  switch (d->mode) {
    case MML_BN1D_INPUT:
      if (reg) bn1d_bwd_kernel<MML_BN1D_INPUT, true><<<grid, block, 0, st>>>(*d);
      else bn1d_bwd_kernel<MML_BN1D_INPUT, false><<<grid, block, 0, st>>>(*d);
      break;
    case MML_BN1D_GATED:
    case MML_BN1D_MAX2:  // same kernel: both hand the gradient of the mixed value on as dz
      MML_REQUIRE(ctx, d->gamma && d->invstd && d->dz, "bn1d_bwd(GATED): gamma / invstd / dz");
      if (reg) bn1d_bwd_kernel<MML_BN1D_GATED, true><<<grid, block, 0, st>>>(*d);
      else bn1d_bwd_kernel<MML_BN1D_GATED, false><<<grid, block, 0, st>>>(*d);
      break;
    case MML_BN1D_MAXOUT:
      MML_REQUIRE(ctx, d->gamma && d->invstd && d->pre && d->dpre, "bn1d_bwd(MAXOUT): gamma / invstd / pre / dpre");
      if (reg) bn1d_bwd_kernel<MML_BN1D_MAXOUT, true><<<grid, block, 0, st>>>(*d);
      else bn1d_bwd_kernel<MML_BN1D_MAXOUT, false><<<grid, block, 0, st>>>(*d);
      break;
    default:
      return mml_set_error(ctx, MML_ERR_INVALID, "bn1d_bwd: unknown mode %d", d->mode);
  }
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_gmu_fwd(mml_ctx* ctx, const uint16_t* h1pre, const uint16_t* h2pre, const float* wz, float* h1, float* h2, float* gate, int B,
                int H, void* stream) {
  MML_REQUIRE(ctx, ctx && h1pre && h2pre && wz && h1 && h2 && gate && B >= 1 && H >= 1, "gmu_fwd: bad arguments");
  gmu_fwd_kernel<<<(unsigned)mml_ceil_div(B, kGmuWarps), kGmuWarps * 32, 0, (cudaStream_t)stream>>>(h1pre, h2pre, wz, h1, h2, gate, B, H);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_gmu_bwd(mml_ctx* ctx, const float* dz, const float* h1, const float* h2, const float* gate, const float* wz, float* dwz,
                uint16_t* dh1pre, uint16_t* dh2pre, int B, int H, void* stream) {
  MML_REQUIRE(ctx, ctx && dz && h1 && h2 && gate && wz && dwz && dh1pre && dh2pre && B >= 1 && H >= 1, "gmu_bwd: bad arguments");
  gmu_bwd_kernel<<<(unsigned)mml_ceil_div(B, kGmuWarps), kGmuWarps * 32, 0, (cudaStream_t)stream>>>(dz, h1, h2, gate, wz, dwz, dh1pre, dh2pre,
                                                                                                   B, H);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_pool_fwd(mml_ctx* ctx, const uint16_t* pre_a, const uint16_t* pre_b, const float* bias_a, const float* bias_b, const uint8_t* keep_a,
                 const uint8_t* keep_b, float keep_scale, float* h_a, float* h_b, uint16_t* comb, int B, int H, void* stream) {
  MML_REQUIRE(ctx, ctx && pre_a && pre_b && bias_a && bias_b && h_a && h_b && B >= 1 && H >= 1, "pool_fwd: bad arguments");
  MML_REQUIRE(ctx, (keep_a == nullptr) == (keep_b == nullptr), "pool_fwd: both dropout masks or none");
  int grid = (int)mml_ceil_div((int64_t)B * H, 256);
  if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
  pool_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pre_a, pre_b, bias_a, bias_b, keep_a, keep_b, keep_scale, h_a, h_b, comb, B, H);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_pool_bwd(mml_ctx* ctx, const float* dz, const float* h_a, const float* h_b, const uint8_t* keep_a, const uint8_t* keep_b,
                 float keep_scale, int kind, float mix_a, float mix_b, const float* gate, const uint16_t* dcomb, uint16_t* dpre_a,
                 uint16_t* dpre_b, float* dbias_a, float* dbias_b, int B, int H, void* stream) {
  MML_REQUIRE(ctx, ctx && dz && h_a && h_b && dpre_a && dpre_b && dbias_a && dbias_b && B >= 1 && H >= 1, "pool_bwd: bad arguments");
  MML_REQUIRE(ctx, (keep_a == nullptr) == (keep_b == nullptr), "pool_bwd: both dropout masks or none");
  MML_REQUIRE(ctx, kind == 0 || kind == 1, "pool_bwd: kind must be 0 (max) or 1 (linear mix)");
  pool_bwd_kernel<<<(unsigned)mml_ceil_div(H, kCols), dim3(kCols, kRowGroups), 0, (cudaStream_t)stream>>>(
      dz, h_a, h_b, keep_a, keep_b, keep_scale, kind, mix_a, mix_b, gate, dcomb, dpre_a, dpre_b, dbias_a, dbias_b, B, H);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_att_fwd(mml_ctx* ctx, const uint16_t* hid, const float* b0, const float* w2, const float* b2, float* t, float* gate, int B, int Hd,
                int NS, void* stream) {
  MML_REQUIRE(ctx, ctx && hid && b0 && w2 && b2 && t && gate && B >= 1 && Hd >= 1, "att_fwd: bad arguments");
  MML_REQUIRE(ctx, NS == 1 || NS == 2, "att_fwd: 1 (gated) or 2 (attention) scores");
  att_fwd_kernel<<<(unsigned)mml_ceil_div(B, kGmuWarps), kGmuWarps * 32, 0, (cudaStream_t)stream>>>(hid, b0, w2, b2, t, gate, B, Hd, NS);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_att_bwd(mml_ctx* ctx, const float* dz, const float* h_a, const float* h_b, const float* gate, const float* t, const float* w2,
                float* dw2, float* db2, float* db0, uint16_t* dhid, int B, int H, int Hd, int NS, void* stream) {
  MML_REQUIRE(ctx, ctx && dz && h_a && h_b && gate && t && w2 && dw2 && db2 && db0 && dhid && B >= 1 && H >= 1 && Hd >= 1, "att_bwd: bad arguments");
  MML_REQUIRE(ctx, NS == 1 || NS == 2, "att_bwd: 1 (gated) or 2 (attention) scores");
  att_bwd_kernel<<<(unsigned)mml_ceil_div(B, kGmuWarps), kGmuWarps * 32, 0, (cudaStream_t)stream>>>(dz, h_a, h_b, gate, t, w2, dw2, db2, db0, dhid,
                                                                                                   B, H, Hd, NS);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int64_t mml_bce_head_scratch_floats(int B) { return 1 + mml_ceil_div(B, kHeadWarps); }

int mml_bce_head_fwd(mml_ctx* ctx, const float* xn, const float* w, const float* bias, const float* labels, float* logits, float* loss,
                     float* dlogits, uint8_t* pred, float* scratch, float threshold, float grad_scale, int B, int H, int NC, void* stream) {
  MML_REQUIRE(ctx, ctx && xn && w && bias && logits && B >= 1 && H >= 1, "bce_head_fwd: bad arguments");
  MML_REQUIRE(ctx, NC >= 1 && NC <= 32, "bce_head_fwd: 1..32 classes supported (got %d)", NC);
  MML_REQUIRE(ctx, !loss || (labels && scratch), "bce_head_fwd: the loss needs labels and scratch");
  MML_REQUIRE(ctx, H % 4 == 0, "bce_head_fwd: H must be a multiple of 4 (got %d)", H);
  const size_t smem = (size_t)(kHeadWarps + NC) * H * sizeof(float);
  MML_REQUIRE(ctx, smem <= 200 * 1024, "bce_head_fwd: (classes + 8) x H = %zu bytes of shared memory", smem);
  static size_t configured = 0;
  if (smem > configured) {
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(bce_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  bce_head_fwd_kernel<<<(unsigned)mml_ceil_div(B, kHeadWarps), kHeadWarps * 32, smem, (cudaStream_t)stream>>>(
      xn, w, bias, labels, logits, loss, dlogits, pred, scratch, threshold, grad_scale, B, H, NC);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_bce_head_bwd(mml_ctx* ctx, const float* dlogits, const float* xn, const float* w, float* dw, float* db, uint16_t* dxn, int B, int H,
                     int NC, void* stream) {
  MML_REQUIRE(ctx, ctx && dlogits && xn && w && dw && db && dxn && B >= 1 && H >= 1 && NC >= 1, "bce_head_bwd: bad arguments");
  MML_REQUIRE(ctx, NC <= 32, "bce_head_bwd: 1..32 classes supported (got %d)", NC);
  bce_head_bwd_kernel<<<(unsigned)mml_ceil_div(H, kCols), dim3(kCols, 8), 0, (cudaStream_t)stream>>>(dlogits, xn, w, dw, db, dxn, B, H, NC);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

}  // extern "C"
