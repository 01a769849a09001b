// utt_fusion.cu -- kernels of the MOSI / UttFusion step (BASELINE config 4) that are not GEMMs.
//
// Reference: MML_Suite/models/msa/networks/lstm.py:8-64 (one-layer batch_first nn.LSTM, embedding = h_T),
// networks/textcnn.py:10-69 (Conv2d(1,128,(k,768)) -> ReLU -> max over time), networks/classifier.py:83-117
// ([Linear, ReLU, Dropout] x 3 + fc_out), utt_fusion.py:177-186 (CE, clip_grad_norm_, Adam).
//
// What runs where:
//   * TextCNN convolutions: the Conv2d weight [128][1][k][768] IS a K,R,S,C tensor with C = 768, the text input [B][T][768] is an
//     NHWC tensor [B][T][1][768]  ->  mml_conv_fprop / mml_conv_wgrad (tcgen05, k tap-shifted GEMMs), no new GEMM code.
//   * LSTM: the recurrence is latency-bound (T = 50 dependent steps of a 256 x 84 mat-vec per sample): one CTA per sample, one
//     thread per gate row, W_hh / W_ih resident in shared memory for all steps; BPTT in a second kernel of the same shape with the
//     weight-gradient accumulators in registers.
//   * everything else (bias + ReLU + max over time, dropout, the classifier's dense layers, global gradient norm) is small and fp32.
#include <string.h>

#include "mml_common.cuh"
#include "mml_ctx.h"

namespace {
using namespace mml;

constexpr int kLstmMaxIn = 32;
constexpr int kLstmH = 64;
constexpr int kGates = 4 * kLstmH;  // 256 = one thread per gate row (i, f, g, o blocks of 64)
constexpr int kLstmStageT = 64;     // sequences up to this length are staged in shared memory by the forward kernel

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// h_T of a one-layer LSTM started from zeros.  x [B][T][IN] fp32; w_ih [256][IN], w_hh [256][64], b_ih, b_hh [256];
// saved for BPTT: gates [B][T][256] (post-activation i, f, g, o), cs [B][T][64] (cell state after step t), hs [B][T][64].
__global__ void __launch_bounds__(kGates) lstm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w_ih,
                                                         const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                                         const float* __restrict__ b_hh, float* __restrict__ gates, float* __restrict__ cs,
                                                         float* __restrict__ hs, float* __restrict__ h_last, int T, int IN) {
  extern __shared__ float sm[];
  float* s_whh = sm;                          // [256][65]
  float* s_wih = s_whh + kGates * 65;         // [256][IN + 1]
  float* s_h = s_wih + kGates * (kLstmMaxIn + 1);
  float* s_x = s_h + kLstmH;
  float* s_g = s_x + kLstmMaxIn;              // [256]
  const int b = blockIdx.x, g = threadIdx.x;
  float* s_xall = s_g + kGates;               // [kLstmStageT][IN]: the sample's input sequence (no global load inside the recurrence)
  for (int i = g; i < kGates * kLstmH; i += kGates) s_whh[(i / kLstmH) * 65 + (i % kLstmH)] = w_hh[i];
  for (int i = g; i < kGates * IN; i += kGates) s_wih[(i / IN) * (kLstmMaxIn + 1) + (i % IN)] = w_ih[i];
  const bool staged = T <= kLstmStageT;
  if (staged)
    for (int i = g; i < T * IN; i += kGates) s_xall[i] = x[(size_t)b * T * IN + i];
  if (g < kLstmH) s_h[g] = 0.f;
  const float bias = b_ih[g] + b_hh[g];
  float c = 0.f;  // threads 0..63 own one cell each
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const float* xt = s_xall + t * IN;
    if (!staged) {
      if (g < IN) s_x[g] = x[((size_t)b * T + t) * IN + g];
      __syncthreads();
      xt = s_x;
    }
    float p0 = bias, p1 = 0.f, p2 = 0.f, p3 = 0.f;  // four independent chains instead of one 84-deep dependent FMA chain
    for (int k = 0; k < IN; ++k) p1 = fmaf(s_wih[g * (kLstmMaxIn + 1) + k], xt[k], p1);
#pragma unroll
    for (int k = 0; k < kLstmH; k += 4) {
      p0 = fmaf(s_whh[g * 65 + k], s_h[k], p0);
      p1 = fmaf(s_whh[g * 65 + k + 1], s_h[k + 1], p1);
      p2 = fmaf(s_whh[g * 65 + k + 2], s_h[k + 2], p2);
      p3 = fmaf(s_whh[g * 65 + k + 3], s_h[k + 3], p3);
    }
    const float pre = (p0 + p1) + (p2 + p3);
    const float act = (g >= 2 * kLstmH && g < 3 * kLstmH) ? tanhf(pre) : sigmoidf_(pre);
    s_g[g] = act;
    gates[((size_t)b * T + t) * kGates + g] = act;
    __syncthreads();
    if (g < kLstmH) {
      c = s_g[kLstmH + g] * c + s_g[g] * s_g[2 * kLstmH + g];
      const float h = s_g[3 * kLstmH + g] * tanhf(c);
      s_h[g] = h;
      cs[((size_t)b * T + t) * kLstmH + g] = c;
      hs[((size_t)b * T + t) * kLstmH + g] = h;
      if (t == T - 1) h_last[(size_t)b * kLstmH + g] = h;
    }
    __syncthreads();
  }
}

// BPTT from dh_T [B][64]: weight / bias gradients ACCUMULATED into dw_ih, dw_hh, db_ih, db_hh (global fp32 atomics, one batch of
// atomics per sample at the end).  The input needs no gradient.
__global__ void __launch_bounds__(kGates) lstm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w_hh,
                                                         const float* __restrict__ gates, const float* __restrict__ cs,
                                                         const float* __restrict__ hs, const float* __restrict__ dh_last,
                                                         float* __restrict__ dw_ih, float* __restrict__ dw_hh, float* __restrict__ db_ih,
                                                         float* __restrict__ db_hh, int T, int IN) {
  extern __shared__ float sm[];
  float* s_whh = sm;                    // [256][65]
  float* s_dg = s_whh + kGates * 65;    // [256] gradient at the gate pre-activations
  float* s_hp = s_dg + kGates;          // h_{t-1}
  float* s_x = s_hp + kLstmH;
  float* s_part = s_x + kLstmMaxIn;     // [4][64] partial sums of dh_prev
  float* s_dh = s_part + 4 * kLstmH;    // [64]
  const int b = blockIdx.x, g = threadIdx.x;
  for (int i = g; i < kGates * kLstmH; i += kGates) s_whh[(i / kLstmH) * 65 + (i % kLstmH)] = w_hh[i];
  float acc_hh[kLstmH];
  float acc_ih[kLstmMaxIn];
#pragma unroll
  for (int k = 0; k < kLstmH; ++k) acc_hh[k] = 0.f;
#pragma unroll
  for (int k = 0; k < kLstmMaxIn; ++k) acc_ih[k] = 0.f;
  float acc_b = 0.f;
  float dc = 0.f;  // threads 0..63: gradient of the cell state flowing back in time
  if (g < kLstmH) s_dh[g] = dh_last[(size_t)b * kLstmH + g];
  __syncthreads();
  // saved state of step t, loaded one step ahead (the loads of step t-1 are in flight while step t computes)
  float n_i = 0.f, n_f = 0.f, n_g = 0.f, n_o = 0.f, n_c = 0.f, n_cp = 0.f, n_hp = 0.f, n_x = 0.f;
  auto fetch = [&](int t) {
    const size_t bt = (size_t)b * T + t;
    if (g < kLstmH) {
      n_i = gates[bt * kGates + g], n_f = gates[bt * kGates + kLstmH + g], n_g = gates[bt * kGates + 2 * kLstmH + g], n_o = gates[bt * kGates + 3 * kLstmH + g];
      n_c = cs[bt * kLstmH + g];
      n_cp = t > 0 ? cs[(bt - 1) * kLstmH + g] : 0.f;
      n_hp = t > 0 ? hs[(bt - 1) * kLstmH + g] : 0.f;
    }
    if (g < IN) n_x = x[bt * IN + g];
  };
  fetch(T - 1);
  for (int t = T - 1; t >= 0; --t) {
    const float i_ = n_i, f_ = n_f, g_ = n_g, o_ = n_o, c_t = n_c, c_prev = n_cp, h_prev = n_hp, x_t = n_x;
    if (t > 0) fetch(t - 1);
    if (g < kLstmH) {
      const float tc = tanhf(c_t), dh = s_dh[g];
      dc += dh * o_ * (1.f - tc * tc);
      s_dg[g] = dc * g_ * i_ * (1.f - i_);
      s_dg[kLstmH + g] = dc * c_prev * f_ * (1.f - f_);
      s_dg[2 * kLstmH + g] = dc * i_ * (1.f - g_ * g_);
      s_dg[3 * kLstmH + g] = dh * tc * o_ * (1.f - o_);
      dc *= f_;
      s_hp[g] = h_prev;
    }
    if (g < IN) s_x[g] = x_t;
    __syncthreads();
    const float d = s_dg[g];
    acc_b += d;
#pragma unroll
    for (int k = 0; k < kLstmH; ++k) acc_hh[k] = fmaf(d, s_hp[k], acc_hh[k]);
#pragma unroll
    for (int k = 0; k < kLstmMaxIn; ++k)
      if (k < IN) acc_ih[k] = fmaf(d, s_x[k], acc_ih[k]);
    // dh_{t-1}[j] = sum_g W_hh[g][j] * dgate[g]: thread (q, j) sums a quarter of the gate rows
    {
      const int j = g & (kLstmH - 1), q = g >> 6;
      float s = 0.f;
#pragma unroll 16
      for (int r = 0; r < kLstmH; ++r) s = fmaf(s_whh[(q * kLstmH + r) * 65 + j], s_dg[q * kLstmH + r], s);
      s_part[q * kLstmH + j] = s;
    }
    __syncthreads();
    if (g < kLstmH) s_dh[g] = s_part[g] + s_part[kLstmH + g] + s_part[2 * kLstmH + g] + s_part[3 * kLstmH + g];
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < kLstmH; ++k) atomicAdd(dw_hh + (size_t)g * kLstmH + k, acc_hh[k]);
#pragma unroll
  for (int k = 0; k < kLstmMaxIn; ++k)
    if (k < IN) atomicAdd(dw_ih + (size_t)g * IN + k, acc_ih[k]);
  atomicAdd(db_ih + g, acc_b);
  atomicAdd(db_hh + g, acc_b);
}

// TextCNN pooling: y[b][c] = max_t relu(conv[b][t][c] + bias[c]) over the P valid positions (conv bf16 [B][P][C]); arg = position of
// the maximum (first one), or -1 when every position is <= 0 (ReLU inactive: no gradient).  Optional dropout on the pooled value.
__global__ void relumax_fwd_kernel(const uint16_t* __restrict__ conv, const float* __restrict__ bias, const uint8_t* __restrict__ keep,
                                   float keep_scale, float* __restrict__ y, int* __restrict__ arg, int B, int P, int C, int ldy, int y_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  float best = 0.f;
  int at = -1;
  const float bc = bias[c];
  for (int t = 0; t < P; ++t) {
    const float v = __uint_as_float((uint32_t)conv[((size_t)b * P + t) * C + c] << 16) + bc;
    if (v > best) best = v, at = t;
  }
  if (keep) best = keep[(size_t)b * ldy + y_off + c] ? best * keep_scale : 0.f;
  y[(size_t)b * ldy + y_off + c] = best;
  arg[(size_t)b * ldy + y_off + c] = at;
}

// backward of the same: dconv bf16 [B][P][C] is zero except at the arg-max position; dbias[c] = sum_b of those gradients.
// One thread per (b, c): P coalesced 2-byte stores down the time axis; the bias gradient meets in a global atomic (dbias is part of the
// flat gradient buffer, zeroed once per step).
__global__ void relumax_bwd_kernel(const float* __restrict__ dy, const int* __restrict__ arg, const uint8_t* __restrict__ keep, float keep_scale,
                                   uint16_t* __restrict__ dconv, float* __restrict__ dbias, int B, int P, int C, int ldy, int y_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const size_t yi = (size_t)b * ldy + y_off + c;
  float g = dy[yi];
  if (keep) g = keep[yi] ? g * keep_scale : 0.f;
  const int at = arg[yi];
  const __nv_bfloat16 gb = __float2bfloat16_rn(g);
  const uint16_t gbits = *reinterpret_cast<const uint16_t*>(&gb);
  for (int t = 0; t < P; ++t) dconv[((size_t)b * P + t) * C + c] = t == at ? gbits : (uint16_t)0;
  if (at >= 0 && g != 0.f) atomicAdd(dbias + c, g);
}

// y[b][o] = act(bias[o] + sum_k x[b][k] w[o][k]) (* keep * scale): one warp per output, small batch (B = 32)
__global__ void __launch_bounds__(256) dense_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                                       const float* __restrict__ bias, const uint8_t* __restrict__ keep, float keep_scale,
                                                       int relu, float* __restrict__ y, int ldy, int B, int K, int N) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * N) return;
  const int b = warp / N, o = warp - b * N;
  float acc = 0.f;
#pragma unroll 4
  for (int k = lane; k < K; k += 32) acc = fmaf(x[(size_t)b * ldx + k], w[(size_t)o * K + k], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc + bias[o];
    if (relu) v = fmaxf(v, 0.f);
    if (keep) v = keep[(size_t)b * N + o] ? v * keep_scale : 0.f;
    y[(size_t)b * ldy + o] = v;
  }
}

// backward of one dense layer: dpre = dy * (keep ? scale : 1) * (relu ? y > 0 : 1) in place in dy; then dx[b][k] = sum_o dpre[b][o] w[o][k],
// dw[o][k] = sum_b dpre[b][o] x[b][k], db[o] = sum_b dpre[b][o]   (dw / db STORED).  Three small kernels.
__global__ void dense_bwd_act_kernel(float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy, const uint8_t* __restrict__ keep,
                                     float keep_scale, int relu, int B, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N, o = i - b * N;
  float g = dy[(size_t)b * lddy + o];
  if (keep) g = keep[i] ? g * keep_scale : 0.f;
  if (relu && !(y[(size_t)b * ldy + o] > 0.f)) g = 0.f;
  dy[(size_t)b * lddy + o] = g;
}

__global__ void dense_bwd_data_kernel(const float* __restrict__ dpre, int ldd, const float* __restrict__ w, float* __restrict__ dx, int lddx,
                                      int B, int K, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  const int b = i / K, k = i - b * K;
  float acc = 0.f;
#pragma unroll 4
  for (int o = 0; o < N; ++o) acc = fmaf(dpre[(size_t)b * ldd + o], w[(size_t)o * K + k], acc);
  dx[(size_t)b * lddx + k] = acc;
}

// dw[o][k] = sum_b dpre[b][o] * x[b][k], db[o] = sum_b dpre[b][o]: one CTA per 32 (o) x 64 (k) tile, the batch walked in chunks of 32 rows
// staged in shared memory, 2 x 4 outputs per thread; the sum over b runs in a fixed order (bit-reproducible).
// (Was one thread per output looping over the batch: 256 dependent L2 round trips, ~100 us for any layer size.)
constexpr int kDwTO = 32, kDwTK = 64, kDwTB = 32;
__global__ void __launch_bounds__(256) dense_bwd_weight_kernel(const float* __restrict__ dpre, int ldd, const float* __restrict__ x, int ldx,
                                                              float* __restrict__ dw, float* __restrict__ db, int B, int K, int N) {
  __shared__ float ds[kDwTB][kDwTO + 1];
  __shared__ __align__(16) float xs[kDwTB][kDwTK];
  const int k0 = blockIdx.x * kDwTK, o0 = blockIdx.y * kDwTO;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float sb[2] = {0.f, 0.f};
  for (int b0 = 0; b0 < B; b0 += kDwTB) {
    for (int i = threadIdx.x; i < kDwTB * kDwTO; i += 256) {
      const int bb = i / kDwTO, oo = i - bb * kDwTO;
      ds[bb][oo] = (b0 + bb < B && o0 + oo < N) ? dpre[(size_t)(b0 + bb) * ldd + o0 + oo] : 0.f;
    }
    for (int i = threadIdx.x; i < kDwTB * kDwTK; i += 256) {
      const int bb = i / kDwTK, kk = i - bb * kDwTK;
      xs[bb][kk] = (b0 + bb < B && k0 + kk < K) ? x[(size_t)(b0 + bb) * ldx + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int bb = 0; bb < kDwTB; ++bb) {
      const float d0 = ds[bb][ty * 2], d1 = ds[bb][ty * 2 + 1];
      const float4 xv = *reinterpret_cast<const float4*>(&xs[bb][tx * 4]);
      acc[0][0] = fmaf(d0, xv.x, acc[0][0]), acc[0][1] = fmaf(d0, xv.y, acc[0][1]), acc[0][2] = fmaf(d0, xv.z, acc[0][2]), acc[0][3] = fmaf(d0, xv.w, acc[0][3]);
      acc[1][0] = fmaf(d1, xv.x, acc[1][0]), acc[1][1] = fmaf(d1, xv.y, acc[1][1]), acc[1][2] = fmaf(d1, xv.z, acc[1][2]), acc[1][3] = fmaf(d1, xv.w, acc[1][3]);
      sb[0] += d0, sb[1] += d1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int o = o0 + ty * 2 + i;
    if (o >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) dw[(size_t)o * K + k] = acc[i][j];
    }
    if (blockIdx.x == 0 && tx == 0) db[o] = sb[i];
  }
}

// global gradient norm for clip_grad_norm_ (utt_fusion.py:182): two-stage sum of squares, then
// hyper[g][5] = base_scale * min(1, clip / (norm + 1e-6)) for every optimizer group row -- the Adam kernel multiplies gradients by it
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ partial) {
  __shared__ double sh[256];
  double a = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a += (double)g[i] * (double)g[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void clip_scale_kernel(const double* __restrict__ partial, int parts, float clip, float base_scale, float* __restrict__ hyper, int groups,
                                  float* __restrict__ norm_out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < parts; i += 32) s += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x != 0) return;
  const float norm = (float)sqrt(s) * base_scale;  // the norm of the gradients Adam will see (base_scale = 1 / world size)
  float coef = clip / (norm + 1e-6f);
  if (coef > 1.f) coef = 1.f;
  for (int gidx = 0; gidx < groups; ++gidx) hyper[gidx * 8 + 5] = base_scale * coef;
  if (norm_out) norm_out[0] = norm;
}

size_t lstm_fwd_smem() { return sizeof(float) * (kGates * 65 + kGates * (kLstmMaxIn + 1) + kLstmH + kLstmMaxIn + kGates + kLstmStageT * kLstmMaxIn); }
size_t lstm_bwd_smem() { return sizeof(float) * (kGates * 65 + kGates + kLstmH + kLstmMaxIn + 4 * kLstmH + kLstmH); }

}  // namespace

extern "C" {

int mml_lstm_fwd(mml_ctx* ctx, const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* gates,
                 float* cs, float* hs, float* h_last, int B, int T, int IN, int H, void* stream) {
  MML_REQUIRE(ctx, ctx && x && w_ih && w_hh && b_ih && b_hh && gates && cs && hs && h_last && B >= 1 && T >= 1, "lstm_fwd: bad arguments");
  MML_REQUIRE(ctx, H == kLstmH && IN >= 1 && IN <= kLstmMaxIn, "lstm_fwd: hidden size 64 and 1..32 input features are built (got H=%d IN=%d)", H, IN);
  static bool configured = false;
  if (!configured) {
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_fwd_smem()));
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_bwd_smem()));
    configured = true;
  }
  lstm_fwd_kernel<<<B, kGates, lstm_fwd_smem(), (cudaStream_t)stream>>>(x, w_ih, w_hh, b_ih, b_hh, gates, cs, hs, h_last, T, IN);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_lstm_bwd(mml_ctx* ctx, const float* x, const float* w_hh, const float* gates, const float* cs, const float* hs, const float* dh_last,
                 float* dw_ih, float* dw_hh, float* db_ih, float* db_hh, int B, int T, int IN, int H, void* stream) {
  MML_REQUIRE(ctx, ctx && x && w_hh && gates && cs && hs && dh_last && dw_ih && dw_hh && db_ih && db_hh && B >= 1 && T >= 1, "lstm_bwd: bad arguments");
  MML_REQUIRE(ctx, H == kLstmH && IN >= 1 && IN <= kLstmMaxIn, "lstm_bwd: hidden size 64 and 1..32 input features are built (got H=%d IN=%d)", H, IN);
  static bool configured = false;
  if (!configured) {
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_fwd_smem()));
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_bwd_smem()));
    configured = true;
  }
  lstm_bwd_kernel<<<B, kGates, lstm_bwd_smem(), (cudaStream_t)stream>>>(x, w_hh, gates, cs, hs, dh_last, dw_ih, dw_hh, db_ih, db_hh, T, IN);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_relumax_fwd(mml_ctx* ctx, const uint16_t* conv, const float* bias, const uint8_t* keep, float keep_scale, float* y, int32_t* arg, int B,
                    int P, int C, int ldy, int y_off, void* stream) {
  MML_REQUIRE(ctx, ctx && conv && bias && y && arg && B >= 1 && P >= 1 && C >= 1 && ldy >= y_off + C, "relumax_fwd: bad arguments");
  relumax_fwd_kernel<<<(unsigned)mml_ceil_div((int64_t)B * C, 128), 128, 0, (cudaStream_t)stream>>>(conv, bias, keep, keep_scale, y, arg, B, P, C, ldy,
                                                                                                   y_off);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_relumax_bwd(mml_ctx* ctx, const float* dy, const int32_t* arg, const uint8_t* keep, float keep_scale, uint16_t* dconv, float* dbias, int B,
                    int P, int C, int ldy, int y_off, void* stream) {
  MML_REQUIRE(ctx, ctx && dy && arg && dconv && dbias && B >= 1 && P >= 1 && C >= 1 && ldy >= y_off + C, "relumax_bwd: bad arguments");
  relumax_bwd_kernel<<<(unsigned)mml_ceil_div((int64_t)B * C, 128), 128, 0, (cudaStream_t)stream>>>(dy, arg, keep, keep_scale, dconv, dbias, B, P, C,
                                                                                                   ldy, y_off);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_dense_fwd(mml_ctx* ctx, const float* x, int ldx, const float* w, const float* bias, const uint8_t* keep, float keep_scale, int relu,
                  float* y, int ldy, int B, int K, int N, void* stream) {
  MML_REQUIRE(ctx, ctx && x && w && bias && y && B >= 1 && K >= 1 && N >= 1 && ldx >= K && ldy >= N, "dense_fwd: bad arguments");
  dense_fwd_kernel<<<(unsigned)mml_ceil_div((int64_t)B * N * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, w, bias, keep, keep_scale, relu, y, ldy,
                                                                                                      B, K, N);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_dense_bwd(mml_ctx* ctx, float* dy, int lddy, const float* y, int ldy, const uint8_t* keep, float keep_scale, int relu, const float* x, int ldx,
                  const float* w, float* dx, int lddx, float* dw, float* db, int B, int K, int N, void* stream) {
  MML_REQUIRE(ctx, ctx && dy && y && x && w && dw && db && B >= 1 && K >= 1 && N >= 1 && lddy >= N, "dense_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  dense_bwd_act_kernel<<<(unsigned)mml_ceil_div((int64_t)B * N, 256), 256, 0, st>>>(dy, lddy, y, ldy, keep, keep_scale, relu, B, N);
  MML_LAUNCHED(ctx);
  if (dx) {
    dense_bwd_data_kernel<<<(unsigned)mml_ceil_div((int64_t)B * K, 256), 256, 0, st>>>(dy, lddy, w, dx, lddx, B, K, N);
    MML_LAUNCHED(ctx);
  }
  dense_bwd_weight_kernel<<<dim3((unsigned)mml_ceil_div(K, kDwTK), (unsigned)mml_ceil_div(N, kDwTO)), 256, 0, st>>>(dy, lddy, x, ldx, dw, db, B, K, N);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_clip_grad_scale(mml_ctx* ctx, const float* g, int64_t n, float clip, float base_scale, float* hyper, int groups, double* partial,
                        float* norm_out, void* stream) {
  MML_REQUIRE(ctx, ctx && g && n >= 1 && hyper && partial && groups >= 1 && clip > 0.f, "clip_grad_scale: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int parts = MML_CLIP_PARTIALS;
  sumsq_kernel<<<parts, 256, 0, st>>>(g, n, partial);
  MML_LAUNCHED(ctx);
  clip_scale_kernel<<<1, 32, 0, st>>>(partial, parts, clip, base_scale, hyper, groups, norm_out);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

}  // extern "C"
