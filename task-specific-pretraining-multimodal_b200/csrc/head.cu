// head.cu -- the concat late-fusion head, fused: encoder fc (audio, image) -> concat -> Linear/ReLU/Dropout ->
// Linear/ReLU -> Linear -> softmax cross-entropy (mean) + argmax, and its backward.
//
// Reference: MML_Suite/models/msa/networks/resnet.py:150,218 (encoder fc), models/avmnist.py:219-236 (head, concat),
// :266-267 (forward), :305 (softmax.argmax), experiment_utils/loss.py:98-148 (CrossEntropyLoss() x 1.0, mean).
// These are tiny fp32 GEMMs (M = batch, at most 512 x 128): latency-bound, so one kernel does the whole chain for a
// group of 8 samples with the activations in shared memory; the concat is just a column offset.  Weight rows are read
// coalesced by a warp (one output neuron per warp pass) and reduced with shuffles.
#include <string.h>

#include "mml_common.cuh"
#include "mml_ctx.h"

using namespace mml;

namespace {

constexpr int SPC = 4;          // samples per CTA of the MLP kernels (every CTA streams the 132 KB of MLP weights once)
constexpr int kHeadThreads = 256;
constexpr int kMaxFeat = 1024;  // FA + FI
constexpr int kMaxEmb = 512;    // EA + EI
constexpr int kMaxHid = 256;

struct HeadDims {
  int FA, FI, EA, EI, H1, H2, NC;
  __host__ __device__ int emb() const { return EA + EI; }
  // scratch layout per sample: [emb | h1 (post dropout) | h2 | probs | loss | demb | dpre1 | dh2 | dlog]
  __host__ __device__ int off_h1() const { return emb(); }
  __host__ __device__ int off_h2() const { return off_h1() + H1; }
  __host__ __device__ int off_prob() const { return off_h2() + H2; }
  __host__ __device__ int off_loss() const { return off_prob() + NC; }
  __host__ __device__ int off_demb() const { return off_loss() + 1; }
  __host__ __device__ int off_dpre1() const { return off_demb() + emb(); }
  __host__ __device__ int off_dh2() const { return off_dpre1() + H1; }
  __host__ __device__ int off_dlog() const { return off_dh2() + H2; }
  __host__ __device__ int per_sample() const { return off_dlog() + NC; }
};

HeadDims dims_of(const mml_head_params* p) {
  HeadDims d;
  d.FA = p->FA, d.FI = p->FI, d.EA = p->EA, d.EI = p->EI, d.H1 = p->H1, d.H2 = p->H2, d.NC = p->NC;
  return d;
}

// out[s][o] = act(b[o] + sum_k in[s][k] * W[o][k]) for the CTA's SPC samples.
// Thread -> (output neuron o, k-slice kq): consecutive threads own consecutive outputs and the same k range, so the
// activation reads are shared-memory broadcasts and every thread walks its own weight row in 16-byte steps (the 128-byte
// lines stay in L1 for the next steps).  Loads are issued in explicit batches of 8 x float4 before any FMA touches them:
// the first version (a warp per output, shuffle reduction, loads interleaved with their dependent FMAs) was a chain of
// exposed L2 latencies -- 48 us forward / 86 us backward for 256 samples.  Partials of the k-slices meet in ``red``.
__device__ __forceinline__ int pow2_slices(int n_items, int n_reduce) {
  int q = 1;
  while (q * 2 * n_items <= kHeadThreads && n_reduce / (q * 2) >= 8) q *= 2;
  return q;
}

template <bool RELU>
__device__ __forceinline__ void dense_layer(const float* __restrict__ W, const float* __restrict__ bias, int n_out, int n_in,
                                            const float* in_s, int in_ld, float* out_s, int out_ld, int out_off, float* red) {
  const int KQ = pow2_slices(n_out, n_in);
  const int per = kHeadThreads / KQ;  // outputs per pass
  const int ol = threadIdx.x % per, kq = threadIdx.x / per;
  const int klen = ((n_in + KQ - 1) / KQ + 3) & ~3;
  const int k0 = kq * klen, k1 = min(n_in, k0 + klen);
  const bool vec = (n_in & 3) == 0 && (in_ld & 3) == 0 && ((size_t)in_s & 15) == 0 && ((size_t)W & 15) == 0;
  for (int o0 = 0; o0 < n_out; o0 += per) {
    const int o = o0 + ol;
    float acc[SPC];
#pragma unroll
    for (int s = 0; s < SPC; ++s) acc[s] = 0.f;
    if (o < n_out) {
      const float* wr = W + (size_t)o * n_in;
      if (vec) {
        for (int k = k0; k < k1; k += 32) {
          float4 wv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) wv[u] = k + 4 * u < k1 ? __ldg(reinterpret_cast<const float4*>(wr + k + 4 * u)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (k + 4 * u < k1) {
#pragma unroll
              for (int s = 0; s < SPC; ++s) {
                const float4 xv = *reinterpret_cast<const float4*>(in_s + s * in_ld + k + 4 * u);
                acc[s] = fmaf(xv.x, wv[u].x, fmaf(xv.y, wv[u].y, fmaf(xv.z, wv[u].z, fmaf(xv.w, wv[u].w, acc[s]))));
              }
            }
          }
        }
      } else {
        for (int k = k0; k < k1; ++k) {
          const float wv = __ldg(wr + k);
#pragma unroll
          for (int s = 0; s < SPC; ++s) acc[s] = fmaf(in_s[s * in_ld + k], wv, acc[s]);
        }
      }
    }
    if (KQ > 1) {
#pragma unroll
      for (int s = 0; s < SPC; ++s) red[threadIdx.x * SPC + s] = acc[s];
      __syncthreads();
      if (kq == 0) {
        for (int q = 1; q < KQ; ++q)
#pragma unroll
          for (int s = 0; s < SPC; ++s) acc[s] += red[(q * per + ol) * SPC + s];
      }
    }
    if (kq == 0 && o < n_out) {
      const float bv = __ldg(bias + o);
#pragma unroll
      for (int s = 0; s < SPC; ++s) {
        const float v = acc[s] + bv;
        out_s[s * out_ld + out_off + o] = RELU ? fmaxf(v, 0.f) : v;
      }
    }
    if (KQ > 1) __syncthreads();  // ``red`` is reused by the next pass / layer
  }
}

// ---- encoder fc layers as tiled fp32 GEMMs --------------------------------------------------------------------------
// The two encoder fc layers hold 98 k of the head's 131 k weights.  Inside the per-sample-group kernel every CTA had to
// stream all of them (524 KB per CTA, a chain of exposed L2 round trips); as a 32 x 32-tiled GEMM over the whole batch each
// CTA reads 128 KB, in 8 K-chunks whose loads are all issued before the previous chunk is consumed.
struct FcJob {
  const float* x;     // [B][K]     (fwd: pooled features;   bwd: d emb, row pitch ldx)
  const float* w;     // [N][K]     (fwd)  or  [K][N] row-major read as W^T (bwd)
  const float* bias;  // [N] or null
  float* y;           // [B][ldy] (+ column offset applied)
  int K, N, ldx, ldy, tiles_n;
};
struct FcJobs {
  FcJob j[2];
};

constexpr int kFcTile = 32, kFcChunk = 64;

// y[b][n] = bias[n] + sum_k x[b][k] * w[n][k]
__global__ void __launch_bounds__(256) fc_fwd_kernel(FcJobs jobs, int B) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  const FcJob J = jobs.j[blockIdx.z];
  if ((int)blockIdx.y >= J.tiles_n) return;
  __shared__ float xs[kFcTile][kFcChunk + 1], wsm[kFcTile][kFcChunk + 1];
  const int tid = threadIdx.x, b0 = blockIdx.x * kFcTile, n0 = blockIdx.y * kFcTile;
  const int tr = tid / 16, tc = tid % 16;           // micro-tile: rows tr, tr+16; columns tc, tc+16
  const int lr = tid / 8, lc = (tid % 8) * 8;       // loader: row lr, 8 consecutive k
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float px[8], pw[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + lc + u;
      px[u] = (b0 + lr < B && k < J.K) ? __ldg(J.x + (size_t)(b0 + lr) * J.ldx + k) : 0.f;
      pw[u] = (n0 + lr < J.N && k < J.K) ? __ldg(J.w + (size_t)(n0 + lr) * J.K + k) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < J.K; k0 += kFcChunk) {
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 8; ++u) xs[lr][lc + u] = px[u], wsm[lr][lc + u] = pw[u];
    __syncthreads();
    if (k0 + kFcChunk < J.K) fetch(k0 + kFcChunk);
#pragma unroll 16
    for (int k = 0; k < kFcChunk; ++k) {
      const float a0 = xs[tr][k], a1 = xs[tr + 16][k], w0 = wsm[tc][k], w1 = wsm[tc + 16][k];
      acc[0][0] = fmaf(a0, w0, acc[0][0]);
      acc[0][1] = fmaf(a0, w1, acc[0][1]);
      acc[1][0] = fmaf(a1, w0, acc[1][0]);
      acc[1][1] = fmaf(a1, w1, acc[1][1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int b = b0 + tr + 16 * i, n = n0 + tc + 16 * j;
      if (b < B && n < J.N) J.y[(size_t)b * J.ldy + n] = acc[i][j] + (J.bias ? __ldg(J.bias + n) : 0.f);
    }
}

// y[b][n] = sum_k x[b][k] * w[k][n]      (d pooled = d emb . W_fc;  K = embedding width, N = feature width)
__global__ void __launch_bounds__(256) fc_bwd_data_kernel(FcJobs jobs, int B) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  const FcJob J = jobs.j[blockIdx.z];
  if ((int)blockIdx.y >= J.tiles_n) return;
  __shared__ float xs[kFcTile][kFcChunk + 1], wsm[kFcChunk][kFcTile + 1];
  const int tid = threadIdx.x, b0 = blockIdx.x * kFcTile, n0 = blockIdx.y * kFcTile;
  const int tr = tid / 16, tc = tid % 16;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < J.K; k0 += kFcChunk) {
    __syncthreads();
    for (int i = tid; i < kFcTile * kFcChunk; i += 256) {
      const int r = i / kFcChunk, k = i % kFcChunk;
      xs[r][k] = (b0 + r < B && k0 + k < J.K) ? __ldg(J.x + (size_t)(b0 + r) * J.ldx + k0 + k) : 0.f;
      const int kk = i / kFcTile, n = i % kFcTile;  // weights: consecutive threads -> consecutive n (coalesced)
      wsm[kk][n] = (k0 + kk < J.K && n0 + n < J.N) ? __ldg(J.w + (size_t)(k0 + kk) * J.N + n0 + n) : 0.f;
    }
    __syncthreads();
#pragma unroll 16
    for (int k = 0; k < kFcChunk; ++k) {
      const float a0 = xs[tr][k], a1 = xs[tr + 16][k], w0 = wsm[k][tc], w1 = wsm[k][tc + 16];
      acc[0][0] = fmaf(a0, w0, acc[0][0]);
      acc[0][1] = fmaf(a0, w1, acc[0][1]);
      acc[1][0] = fmaf(a1, w0, acc[1][0]);
      acc[1][1] = fmaf(a1, w1, acc[1][1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int b = b0 + tr + 16 * i, n = n0 + tc + 16 * j;
      if (b < B && n < J.N) J.y[(size_t)b * J.ldy + n] = acc[i][j];
    }
}

__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(mml_head_params p, HeadDims d, const float* __restrict__ pooledA, const float* __restrict__ pooledI,
                const long long* __restrict__ labels, const uint8_t* __restrict__ drop, float drop_scale, float* __restrict__ scratch,
                float* __restrict__ logits, int* __restrict__ pred, int B) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  extern __shared__ float sm[];
  float* es = sm;                          // [SPC][emb]   (encoder fc outputs, computed by fc_fwd_kernel into scratch)
  float* h1 = es + SPC * d.emb();          // [SPC][H1]
  float* h2 = h1 + SPC * d.H1;             // [SPC][H2]
  float* lg = h2 + SPC * d.H2;             // [SPC][NC]
  __shared__ float red[kHeadThreads * SPC];
  const int s0 = blockIdx.x * SPC;
  const int PS = d.per_sample();
  for (int i = threadIdx.x; i < SPC * d.emb(); i += kHeadThreads) {
    const int s = i / d.emb(), j = i - s * d.emb();
    es[i] = s0 + s < B ? scratch[(size_t)(s0 + s) * PS + j] : 0.f;  // concat == column offset inside the scratch row
  }
  __syncthreads();
  dense_layer<true>(p.w0, p.b0, d.H1, d.emb(), es, d.emb(), h1, d.H1, 0, red);
  __syncthreads();
  if (drop != nullptr) {
    for (int i = threadIdx.x; i < SPC * d.H1; i += kHeadThreads) {
      const int s = i / d.H1, j = i - s * d.H1;
      const int b = s0 + s;
      if (b < B) h1[i] = drop[(size_t)b * d.H1 + j] ? h1[i] * drop_scale : 0.f;
    }
    __syncthreads();
  }
  dense_layer<true>(p.w3, p.b3, d.H2, d.H1, h1, d.H1, h2, d.H2, 0, red);
  __syncthreads();
  dense_layer<false>(p.w5, p.b5, d.NC, d.H2, h2, d.H2, lg, d.NC, 0, red);
  __syncthreads();
  // save activations for backward (the embeddings are already in scratch)
  for (int i = threadIdx.x; i < SPC * d.H1; i += kHeadThreads) {
    const int s = i / d.H1, j = i - s * d.H1;
    if (s0 + s < B) scratch[(size_t)(s0 + s) * PS + d.off_h1() + j] = h1[i];
  }
  for (int i = threadIdx.x; i < SPC * d.H2; i += kHeadThreads) {
    const int s = i / d.H2, j = i - s * d.H2;
    if (s0 + s < B) scratch[(size_t)(s0 + s) * PS + d.off_h2() + j] = h2[i];
  }
  // softmax / CE / argmax: one warp per sample (NC <= 32)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = s0 + warp;
  if (warp < SPC && b < B) {
    const float v = lane < d.NC ? lg[warp * d.NC + lane] : -INFINITY;
    float mx = v;
    int arg = lane < d.NC ? lane : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ov > mx || (ov == mx && oa < arg)) {
        mx = ov;
        arg = oa;
      }
    }
    const float e = lane < d.NC ? expf(v - mx) : 0.f;
    const float se = warp_sum(e);
    if (lane < d.NC) {
      logits[(size_t)b * d.NC + lane] = v;
      scratch[(size_t)b * PS + d.off_prob() + lane] = e / se;
    }
    if (lane == 0) {
      pred[b] = arg;
      float li = 0.f;
      if (labels != nullptr) {
        const int y = (int)labels[b];
        li = (y >= 0 && y < d.NC) ? (logf(se) + mx - lg[warp * d.NC + y]) : 0.f;
      }
      scratch[(size_t)b * PS + d.off_loss()] = li;
    }
  }
}

__global__ void head_loss_kernel(const float* __restrict__ scratch, int PS, int off_loss, int B, float* __restrict__ loss_out) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  __shared__ float sh[256];
  float a = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) a += scratch[(size_t)b * PS + off_loss];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = sh[0] / (float)B;
}

// dst[s][i] = sum_o W[o][i] * src[s][o]   (W^T product).  Thread -> (input column i, o-slice): consecutive threads read
// consecutive columns of a weight row (coalesced), 16 rows per explicit load batch; o-slices meet in ``red``.
__device__ __forceinline__ void dense_layer_t(const float* __restrict__ W, int n_out, int n_in, const float* src_s, int src_ld,
                                              float* dst_s, int dst_ld, float* red) {
  const int OQ = pow2_slices(n_in, n_out);
  const int per = kHeadThreads / OQ;  // columns per pass
  const int il = threadIdx.x % per, oq = threadIdx.x / per;
  const int olen = (n_out + OQ - 1) / OQ;
  const int o0 = oq * olen, o1 = min(n_out, o0 + olen);
  for (int i0 = 0; i0 < n_in; i0 += per) {
    const int i = i0 + il;
    float acc[SPC];
#pragma unroll
    for (int s = 0; s < SPC; ++s) acc[s] = 0.f;
    if (i < n_in) {
      for (int o = o0; o < o1; o += 16) {
        float wv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) wv[u] = o + u < o1 ? __ldg(W + (size_t)(o + u) * n_in + i) : 0.f;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          if (o + u < o1) {
#pragma unroll
            for (int s = 0; s < SPC; ++s) acc[s] = fmaf(src_s[s * src_ld + o + u], wv[u], acc[s]);
          }
        }
      }
    }
    if (OQ > 1) {
#pragma unroll
      for (int s = 0; s < SPC; ++s) red[threadIdx.x * SPC + s] = acc[s];
      __syncthreads();
      if (oq == 0) {
        for (int q = 1; q < OQ; ++q)
#pragma unroll
          for (int s = 0; s < SPC; ++s) acc[s] += red[(q * per + il) * SPC + s];
      }
    }
    if (oq == 0 && i < n_in) {
#pragma unroll
      for (int s = 0; s < SPC; ++s) dst_s[s * dst_ld + i] = acc[s];
    }
    if (OQ > 1) __syncthreads();
  }
}

__global__ void __launch_bounds__(kHeadThreads)
head_bwd_data_kernel(mml_head_params p, HeadDims d, const long long* __restrict__ labels, const uint8_t* __restrict__ drop,
                     float drop_scale, float* __restrict__ scratch, float loss_scale, float* __restrict__ dpooledA,
                     float* __restrict__ dpooledI, int B) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  extern __shared__ float sm[];
  float* dlog = sm;                       // [SPC][NC]
  float* dh2 = dlog + SPC * d.NC;         // [SPC][H2]
  float* dh1 = dh2 + SPC * d.H2;          // [SPC][H1]
  float* demb = dh1 + SPC * d.H1;         // [SPC][emb]
  __shared__ float red[kHeadThreads * SPC];
  const int s0 = blockIdx.x * SPC;
  const int PS = d.per_sample();
  const float invB = loss_scale / (float)B;
  for (int i = threadIdx.x; i < SPC * d.NC; i += kHeadThreads) {
    const int s = i / d.NC, c = i - s * d.NC;
    const int b = s0 + s;
    float g = 0.f;
    if (b < B) {
      const float pr = scratch[(size_t)b * PS + d.off_prob() + c];
      g = (pr - ((int)labels[b] == c ? 1.f : 0.f)) * invB;
      scratch[(size_t)b * PS + d.off_dlog() + c] = g;
    }
    dlog[i] = g;
  }
  __syncthreads();
  dense_layer_t(p.w5, d.NC, d.H2, dlog, d.NC, dh2, d.H2, red);
  __syncthreads();
  for (int i = threadIdx.x; i < SPC * d.H2; i += kHeadThreads) {
    const int s = i / d.H2, j = i - s * d.H2;
    const int b = s0 + s;
    float g = 0.f;
    if (b < B) {
      g = scratch[(size_t)b * PS + d.off_h2() + j] > 0.f ? dh2[i] : 0.f;
      scratch[(size_t)b * PS + d.off_dh2() + j] = g;
    }
    dh2[i] = g;
  }
  __syncthreads();
  dense_layer_t(p.w3, d.H2, d.H1, dh2, d.H2, dh1, d.H1, red);
  __syncthreads();
  for (int i = threadIdx.x; i < SPC * d.H1; i += kHeadThreads) {
    const int s = i / d.H1, j = i - s * d.H1;
    const int b = s0 + s;
    float g = 0.f;
    if (b < B) {
      // h1 is stored post-dropout: zero where dropped or where ReLU was inactive
      const float sc = drop != nullptr ? drop_scale : 1.f;
      g = scratch[(size_t)b * PS + d.off_h1() + j] > 0.f ? dh1[i] * sc : 0.f;
      scratch[(size_t)b * PS + d.off_dpre1() + j] = g;
    }
    dh1[i] = g;
  }
  __syncthreads();
  dense_layer_t(p.w0, d.H1, d.emb(), dh1, d.H1, demb, d.emb(), red);
  __syncthreads();
  for (int i = threadIdx.x; i < SPC * d.emb(); i += kHeadThreads) {
    const int s = i / d.emb(), j = i - s * d.emb();
    if (s0 + s < B) scratch[(size_t)(s0 + s) * PS + d.off_demb() + j] = demb[i];
  }
}

// weight gradients: one block per output neuron of one of the five Linear layers
struct WgradJob {
  const float* dout;  // [B][dout_ld] (+ column offset applied)
  const float* in;    // [B][in_ld]
  float* dw;          // [n_out][n_in]
  float* db;          // [n_out]
  int dout_ld, in_ld, n_out, n_in, first_block;
};
struct WgradJobs {
  WgradJob j[5];
};

__global__ void __launch_bounds__(256) head_bwd_weights_kernel(WgradJobs jobs, int B) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  int ji = 0;
#pragma unroll
  for (int t = 1; t < 5; ++t)
    if ((int)blockIdx.x >= jobs.j[t].first_block) ji = t;
  const WgradJob J = jobs.j[ji];
  const int o = blockIdx.x - J.first_block;
  __shared__ float dsh[256];
  float bsum = 0.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // n_in <= 1024
  for (int b0 = 0; b0 < B; b0 += 256) {
    const int nb = min(256, B - b0);
    __syncthreads();
    if ((int)threadIdx.x < nb) dsh[threadIdx.x] = J.dout[(size_t)(b0 + threadIdx.x) * J.dout_ld + o];
    __syncthreads();
#pragma unroll 8
    for (int s = 0; s < nb; ++s) {
      const float g = dsh[s];
      bsum += g;
      const float* inrow = J.in + (size_t)(b0 + s) * J.in_ld;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = threadIdx.x + u * 256;
        if (i < J.n_in) acc[u] = fmaf(g, __ldg(inrow + i), acc[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = threadIdx.x + u * 256;
    if (i < J.n_in) J.dw[(size_t)o * J.n_in + i] = acc[u];
  }
  if (threadIdx.x == 0) J.db[o] = bsum;
}

// y[b][o] = bias[o] + sum_k x[b][k] * W[o][k]: the encoder fc used stand-alone (resnet.py:150,218); one warp per output
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                                         float* __restrict__ y, int B, int n_in, int n_out) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * n_out) return;
  const int b = warp / n_out, o = warp - b * n_out;
  float acc = 0.f;
  for (int k = lane; k < n_in; k += 32) acc = fmaf(__ldg(x + (size_t)b * n_in + k), __ldg(W + (size_t)o * n_in + k), acc);
  acc = warp_sum(acc);
  if (lane == 0) y[(size_t)b * n_out + o] = acc + (bias ? bias[o] : 0.f);
}

__global__ void dropout_mask_kernel(uint8_t* __restrict__ mask, long long n, float p, unsigned long long seed,
                                    const long long* __restrict__ step) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  const unsigned long long st = step ? (unsigned long long)step[0] : 0ull;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // splitmix64 of (seed, step, index): counter-based, graph-replay safe (step lives on the device)
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (st * 0x100000001B3ull + (unsigned long long)i + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
    mask[i] = u >= p ? 1 : 0;
  }
}

// softmax cross-entropy over [B][NC] logits (NC <= 32), one warp per sample: probabilities -> dlogits = (p - onehot) * scale / B,
// per-sample loss, argmax (first maximum, like torch.argmax)
__global__ void __launch_bounds__(256) softmax_ce_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                        float* __restrict__ dlogits, float* __restrict__ row_loss, int* __restrict__ pred,
                                                        float scale, int B, int NC) {
  pdl_sync();  // programmatic dependent launch: wait for the previous kernel of the stream, then let the next one be scheduled
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float v = lane < NC ? logits[(size_t)b * NC + lane] : -INFINITY;
  float mx = v;
  int arg = lane < NC ? lane : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ov > mx || (ov == mx && oa < arg)) mx = ov, arg = oa;
  }
  const float e = lane < NC ? expf(v - mx) : 0.f;
  const float se = warp_sum(e);
  const int y = labels ? (int)labels[b] : -1;
  if (lane < NC && dlogits) dlogits[(size_t)b * NC + lane] = (e / se - (lane == y ? 1.f : 0.f)) * scale / (float)B;
  const float vy = __shfl_sync(0xffffffffu, v, y >= 0 && y < NC ? y : 0);
  if (lane == 0) {
    if (pred) pred[b] = arg;
    if (row_loss) row_loss[b] = (y >= 0 && y < NC) ? (logf(se) + mx - vy) : 0.f;
  }
}

int check_head(mml_ctx* ctx, const mml_head_params* p) {
  MML_REQUIRE(ctx, ctx && p, "head: null ctx/params");
  MML_REQUIRE(ctx, p->fcA_w && p->fcA_b && p->fcI_w && p->fcI_b && p->w0 && p->b0 && p->w3 && p->b3 && p->w5 && p->b5,
              "head: null weight pointer");
  MML_REQUIRE(ctx, p->FA >= 1 && p->FI >= 1 && p->FA + p->FI <= kMaxFeat && p->FA <= 1024 && p->FI <= 1024, "head: feature dims unsupported");
  MML_REQUIRE(ctx, p->EA >= 1 && p->EI >= 1 && p->EA + p->EI <= kMaxEmb, "head: embedding dims unsupported");
  MML_REQUIRE(ctx, p->H1 >= 1 && p->H1 <= kMaxHid && p->H2 >= 1 && p->H2 <= kMaxHid && p->NC >= 1 && p->NC <= 32, "head: hidden dims unsupported");
  return MML_OK;
}

}  // namespace

extern "C" {

int mml_head_scratch_per_sample(const mml_head_params* p) { return p ? dims_of(p).per_sample() : 0; }

int mml_head_fwd(mml_ctx* ctx, const mml_head_params* p, const float* pooledA, const float* pooledI, const int64_t* labels,
                 const uint8_t* dropout_mask, float dropout_scale, float* scratch, float* logits, float* loss_out,
                 int32_t* pred, int B, void* stream) {
  int rc = check_head(ctx, p);
  if (rc) return rc;
  MML_REQUIRE(ctx, pooledA && pooledI && scratch && logits && pred && B >= 1, "head_fwd: bad arguments");
  MML_REQUIRE(ctx, labels == nullptr || loss_out != nullptr, "head_fwd: labels given without loss_out");
  const HeadDims d = dims_of(p);
  const int smem = SPC * (d.emb() + d.H1 + d.H2 + d.NC) * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(head_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  MML_REQUIRE(ctx, smem <= 96 * 1024, "head_fwd: dims need %d bytes of shared memory", smem);
  cudaStream_t st = (cudaStream_t)stream;
  {
    FcJobs jobs;
    const int PS = d.per_sample();
    jobs.j[0] = {pooledA, p->fcA_w, p->fcA_b, scratch, d.FA, d.EA, d.FA, PS, (int)mml_ceil_div(d.EA, kFcTile)};
    jobs.j[1] = {pooledI, p->fcI_w, p->fcI_b, scratch + d.EA, d.FI, d.EI, d.FI, PS, (int)mml_ceil_div(d.EI, kFcTile)};
    const int ty = jobs.j[0].tiles_n > jobs.j[1].tiles_n ? jobs.j[0].tiles_n : jobs.j[1].tiles_n;
    MML_LAUNCH(ctx, fc_fwd_kernel, dim3((unsigned)mml_ceil_div(B, kFcTile), ty, 2), 256, 0, st, jobs, B);
  }
  MML_LAUNCH(ctx, head_fwd_kernel, (B + SPC - 1) / SPC, kHeadThreads, smem, st, *p, d, pooledA, pooledI, (const long long*)labels, dropout_mask,
                                                                  dropout_scale, scratch, logits, pred, B);
  if (labels != nullptr) {
    MML_LAUNCH(ctx, head_loss_kernel, 1, 256, 0, st, scratch, d.per_sample(), d.off_loss(), B, loss_out);
  }
  return MML_OK;
}

int mml_head_bwd(mml_ctx* ctx, const mml_head_params* p, const mml_head_grads* g, const float* pooledA, const float* pooledI,
                 const int64_t* labels, const uint8_t* dropout_mask, float dropout_scale, float* scratch,
                 float loss_scale, float* dpooledA, float* dpooledI, int B, int phases, void* stream) {
  int rc = check_head(ctx, p);
  if (rc) return rc;
  MML_REQUIRE(ctx, g && pooledA && pooledI && labels && scratch && dpooledA && dpooledI && B >= 1, "head_bwd: bad arguments");
  MML_REQUIRE(ctx, g->fcA_w && g->fcA_b && g->fcI_w && g->fcI_b && g->w0 && g->b0 && g->w3 && g->b3 && g->w5 && g->b5,
              "head_bwd: null gradient pointer");

  const HeadDims d = dims_of(p);
  const int smem = SPC * (d.NC + d.H2 + d.H1 + d.emb()) * (int)sizeof(float);
  MML_REQUIRE(ctx, smem <= 96 * 1024, "head_bwd: dims need %d bytes of shared memory", smem);
  cudaStream_t st = (cudaStream_t)stream;
  MML_REQUIRE(ctx, phases >= 1 && phases <= 3, "head_bwd: phases must be 1 (data), 2 (weights) or 3 (both)");
  if (phases & 1) {
    MML_LAUNCH(ctx, head_bwd_data_kernel, (B + SPC - 1) / SPC, kHeadThreads, smem, st, *p, d, (const long long*)labels, dropout_mask, dropout_scale,
                                                                         scratch, loss_scale, dpooledA, dpooledI, B);
    FcJobs jobs;
    const int PS = d.per_sample();
    jobs.j[0] = {scratch + d.off_demb(), p->fcA_w, nullptr, dpooledA, d.EA, d.FA, PS, d.FA, (int)mml_ceil_div(d.FA, kFcTile)};
    jobs.j[1] = {scratch + d.off_demb() + d.EA, p->fcI_w, nullptr, dpooledI, d.EI, d.FI, PS, d.FI, (int)mml_ceil_div(d.FI, kFcTile)};
    const int ty = jobs.j[0].tiles_n > jobs.j[1].tiles_n ? jobs.j[0].tiles_n : jobs.j[1].tiles_n;
    MML_LAUNCH(ctx, fc_bwd_data_kernel, dim3((unsigned)mml_ceil_div(B, kFcTile), ty, 2), 256, 0, st, jobs, B);
  }
  if (!(phases & 2)) return MML_OK;
  const int PS = d.per_sample();
  WgradJobs jobs;
  int fb = 0;
  auto set = [&](int i, const float* dout, int dout_ld, const float* in, int in_ld, float* dw, float* db, int n_out, int n_in) {
    jobs.j[i].dout = dout, jobs.j[i].dout_ld = dout_ld, jobs.j[i].in = in, jobs.j[i].in_ld = in_ld;
    jobs.j[i].dw = dw, jobs.j[i].db = db, jobs.j[i].n_out = n_out, jobs.j[i].n_in = n_in, jobs.j[i].first_block = fb;
    fb += n_out;
  };
  set(0, scratch + d.off_demb(), PS, pooledA, d.FA, g->fcA_w, g->fcA_b, d.EA, d.FA);
  set(1, scratch + d.off_demb() + d.EA, PS, pooledI, d.FI, g->fcI_w, g->fcI_b, d.EI, d.FI);
  set(2, scratch + d.off_dpre1(), PS, scratch, PS, g->w0, g->b0, d.H1, d.emb());
  set(3, scratch + d.off_dh2(), PS, scratch + d.off_h1(), PS, g->w3, g->b3, d.H2, d.H1);
  set(4, scratch + d.off_dlog(), PS, scratch + d.off_h2(), PS, g->w5, g->b5, d.NC, d.H2);
  MML_LAUNCH(ctx, head_bwd_weights_kernel, fb, 256, 0, st, jobs, B);
  return MML_OK;
}

int mml_softmax_ce(mml_ctx* ctx, const float* logits, const int64_t* labels, float* dlogits, float* row_loss, float* loss_out, int32_t* pred,
                   float loss_scale, int B, int NC, void* stream) {
  MML_REQUIRE(ctx, ctx && logits && B >= 1, "softmax_ce: bad arguments");
  MML_REQUIRE(ctx, NC >= 1 && NC <= 32, "softmax_ce: 1..32 classes supported (got %d)", NC);
  MML_REQUIRE(ctx, !loss_out || (labels && row_loss), "softmax_ce: the loss needs labels and row_loss");
  MML_REQUIRE(ctx, !dlogits || labels, "softmax_ce: the gradient needs labels");
  cudaStream_t st = (cudaStream_t)stream;
  MML_LAUNCH(ctx, softmax_ce_kernel, (unsigned)mml_ceil_div(B, 8), 256, 0, st, logits, (const long long*)labels, dlogits, row_loss, pred, loss_scale, B, NC);
  if (loss_out) {
    MML_LAUNCH(ctx, head_loss_kernel, 1, 256, 0, st, row_loss, 1, 0, B, loss_out);
  }
  return MML_OK;
}

int mml_mono_head_fwd(mml_ctx* ctx, const float* pooled, const float* fc_w, const float* fc_b, const float* cls_w, const float* cls_b,
                      const int64_t* labels, float* emb, float* logits, float* dlogits, float* row_loss, float* loss_out, int32_t* pred,
                      float loss_scale, int B, int F, int E, int NC, void* stream) {
  MML_REQUIRE(ctx, ctx && pooled && fc_w && fc_b && cls_w && cls_b && emb && logits && B >= 1 && F >= 1 && E >= 1, "mono_head_fwd: bad arguments");
  MML_REQUIRE(ctx, NC >= 1 && NC <= 32, "mono_head_fwd: 1..32 classes supported (got %d)", NC);
  MML_REQUIRE(ctx, !loss_out || (labels && row_loss), "mono_head_fwd: the loss needs labels and row_loss");
  cudaStream_t st = (cudaStream_t)stream;
  FcJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  jobs.j[0] = {pooled, fc_w, fc_b, emb, F, E, F, E, (int)mml_ceil_div(E, kFcTile)};
  MML_LAUNCH(ctx, fc_fwd_kernel, dim3((unsigned)mml_ceil_div(B, kFcTile), jobs.j[0].tiles_n, 1), 256, 0, st, jobs, B);
  jobs.j[0] = {emb, cls_w, cls_b, logits, E, NC, E, NC, (int)mml_ceil_div(NC, kFcTile)};
  MML_LAUNCH(ctx, fc_fwd_kernel, dim3((unsigned)mml_ceil_div(B, kFcTile), jobs.j[0].tiles_n, 1), 256, 0, st, jobs, B);
  MML_LAUNCH(ctx, softmax_ce_kernel, (unsigned)mml_ceil_div(B, 8), 256, 0, st, logits, (const long long*)labels, dlogits, row_loss, pred, loss_scale, B, NC);
  if (loss_out) {
    MML_LAUNCH(ctx, head_loss_kernel, 1, 256, 0, st, row_loss, 1, 0, B, loss_out);
  }
  return MML_OK;
}

int mml_mono_head_bwd(mml_ctx* ctx, const float* pooled, const float* emb, const float* dlogits, const float* fc_w, const float* cls_w,
                      float* d_fc_w, float* d_fc_b, float* d_cls_w, float* d_cls_b, float* demb, float* dpooled, int B, int F, int E, int NC,
                      void* stream) {
  MML_REQUIRE(ctx, ctx && pooled && emb && dlogits && fc_w && cls_w && d_fc_w && d_fc_b && d_cls_w && d_cls_b && demb && dpooled && B >= 1,
              "mono_head_bwd: bad arguments");
  MML_REQUIRE(ctx, F <= 1024 && E <= 1024, "mono_head_bwd: feature / embedding width above 1024");
  cudaStream_t st = (cudaStream_t)stream;
  FcJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  jobs.j[0] = {dlogits, cls_w, nullptr, demb, NC, E, NC, E, (int)mml_ceil_div(E, kFcTile)};     // d emb = d logits . W_cls
  MML_LAUNCH(ctx, fc_bwd_data_kernel, dim3((unsigned)mml_ceil_div(B, kFcTile), jobs.j[0].tiles_n, 1), 256, 0, st, jobs, B);
  jobs.j[0] = {demb, fc_w, nullptr, dpooled, E, F, E, F, (int)mml_ceil_div(F, kFcTile)};        // d pooled = d emb . W_fc
  MML_LAUNCH(ctx, fc_bwd_data_kernel, dim3((unsigned)mml_ceil_div(B, kFcTile), jobs.j[0].tiles_n, 1), 256, 0, st, jobs, B);
  WgradJobs wj;
  memset(&wj, 0, sizeof(wj));
  wj.j[0] = {demb, pooled, d_fc_w, d_fc_b, E, F, E, F, 0};
  wj.j[1] = {dlogits, emb, d_cls_w, d_cls_b, NC, E, NC, E, E};
  for (int t = 2; t < 5; ++t) wj.j[t].first_block = E + NC;  // unused
  MML_LAUNCH(ctx, head_bwd_weights_kernel, E + NC, 256, 0, st, wj, B);
  return MML_OK;
}

int mml_linear_fwd(mml_ctx* ctx, const float* x, const float* w, const float* bias, float* y, int B, int n_in, int n_out, void* stream) {
  MML_REQUIRE(ctx, ctx && x && w && y && B >= 1 && n_in >= 1 && n_out >= 1, "linear_fwd: bad arguments");
  const long long warps = (long long)B * n_out;
  MML_LAUNCH(ctx, linear_fwd_kernel, (unsigned)mml_ceil_div(warps * 32, 256), 256, 0, (cudaStream_t)stream, x, w, bias, y, B, n_in, n_out);
  return MML_OK;
}

int mml_dropout_mask(mml_ctx* ctx, uint8_t* mask, int64_t n, float p, uint64_t seed, const int64_t* step_counter, void* stream) {
  MML_REQUIRE(ctx, ctx && mask && n >= 1 && p >= 0.f && p < 1.f, "dropout_mask: bad arguments");
  int grid = (int)mml_ceil_div(n, 256);
  if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
  MML_LAUNCH(ctx, dropout_mask_kernel, grid, 256, 0, (cudaStream_t)stream, mask, n, p, seed, (const long long*)step_counter);
  return MML_OK;
}

}  // extern "C"
