// head.cu -- the concat late-fusion head, fused: encoder fc (audio, image) -> concat -> Linear/ReLU/Dropout ->
// Linear/ReLU -> Linear -> softmax cross-entropy (mean) + argmax, and its backward.
//
// Reference: MML_Suite/models/msa/networks/resnet.py:150,218 (encoder fc), models/avmnist.py:219-236 (head, concat),
// :266-267 (forward), :305 (softmax.argmax), experiment_utils/loss.py:98-148 (CrossEntropyLoss() x 1.0, mean).
// These are tiny fp32 GEMMs (M = batch, at most 512 x 128): latency-bound, so one kernel does the whole chain for a
// group of 8 samples with the activations in shared memory; the concat is just a column offset.  Weight rows are read
// coalesced by a warp (one output neuron per warp pass) and reduced with shuffles.
#include "mml_common.cuh"
#include "mml_ctx.h"

using namespace mml;

namespace {

constexpr int SPC = 2;          // samples per CTA (B/2 CTAs: one per SM at B = 256..296)
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

// out[s][o] = act(b[o] + sum_k in[s][k] * W[o][k]) for the CTA's SPC samples.  A warp owns OPW output neurons per pass (OPW
// independent coalesced weight-row streams in flight), lanes split K, shuffles reduce.
constexpr int OPW = 4;
template <bool RELU>
__device__ __forceinline__ void dense_layer(const float* __restrict__ W, const float* __restrict__ bias, int n_out, int n_in,
                                            const float* in_s, int in_ld, float* out_s, int out_ld, int out_off) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o0 = warp * OPW; o0 < n_out; o0 += (kHeadThreads / 32) * OPW) {
    float acc[OPW][SPC];
#pragma unroll
    for (int u = 0; u < OPW; ++u)
#pragma unroll
      for (int s = 0; s < SPC; ++s) acc[u][s] = 0.f;
#pragma unroll 2
    for (int k = lane; k < n_in; k += 32) {
      float xv[SPC];
#pragma unroll
      for (int s = 0; s < SPC; ++s) xv[s] = in_s[s * in_ld + k];
#pragma unroll
      for (int u = 0; u < OPW; ++u) {
        const int o = min(o0 + u, n_out - 1);
        const float wv = __ldg(W + (size_t)o * n_in + k);
#pragma unroll
        for (int s = 0; s < SPC; ++s) acc[u][s] = fmaf(xv[s], wv, acc[u][s]);
      }
    }
#pragma unroll
    for (int u = 0; u < OPW; ++u)
#pragma unroll
      for (int s = 0; s < SPC; ++s) acc[u][s] = warp_sum(acc[u][s]);
    if (lane < OPW * SPC) {
      const int u = lane / SPC, sidx = lane % SPC;
      float v = 0.f;
#pragma unroll
      for (int uu = 0; uu < OPW; ++uu)
#pragma unroll
        for (int s = 0; s < SPC; ++s)
          if (uu == u && s == sidx) v = acc[uu][s];
      const int o = o0 + u;
      if (o < n_out) {
        v += __ldg(bias + o);
        out_s[sidx * out_ld + out_off + o] = RELU ? fmaxf(v, 0.f) : v;
      }
    }
  }
}

__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(mml_head_params p, HeadDims d, const float* __restrict__ pooledA, const float* __restrict__ pooledI,
                const long long* __restrict__ labels, const uint8_t* __restrict__ drop, float drop_scale, float* __restrict__ scratch,
                float* __restrict__ logits, int* __restrict__ pred, int B) {
  extern __shared__ float sm[];
  float* xs = sm;                          // [SPC][FA+FI]
  float* es = xs + SPC * (d.FA + d.FI);    // [SPC][emb]
  float* h1 = es + SPC * d.emb();          // [SPC][H1]
  float* h2 = h1 + SPC * d.H1;             // [SPC][H2]
  float* lg = h2 + SPC * d.H2;             // [SPC][NC]
  const int s0 = blockIdx.x * SPC;
  const int F = d.FA + d.FI;
  for (int i = threadIdx.x; i < SPC * F; i += kHeadThreads) {
    const int s = i / F, k = i - s * F;
    const int b = s0 + s;
    float v = 0.f;
    if (b < B) v = k < d.FA ? pooledA[(size_t)b * d.FA + k] : pooledI[(size_t)b * d.FI + (k - d.FA)];
    xs[i] = v;
  }
  __syncthreads();
  dense_layer<false>(p.fcA_w, p.fcA_b, d.EA, d.FA, xs, F, es, d.emb(), 0);
  dense_layer<false>(p.fcI_w, p.fcI_b, d.EI, d.FI, xs + d.FA, F, es, d.emb(), d.EA);  // concat == column offset
  __syncthreads();
  dense_layer<true>(p.w0, p.b0, d.H1, d.emb(), es, d.emb(), h1, d.H1, 0);
  __syncthreads();
  if (drop != nullptr) {
    for (int i = threadIdx.x; i < SPC * d.H1; i += kHeadThreads) {
      const int s = i / d.H1, j = i - s * d.H1;
      const int b = s0 + s;
      if (b < B) h1[i] = drop[(size_t)b * d.H1 + j] ? h1[i] * drop_scale : 0.f;
    }
    __syncthreads();
  }
  dense_layer<true>(p.w3, p.b3, d.H2, d.H1, h1, d.H1, h2, d.H2, 0);
  __syncthreads();
  dense_layer<false>(p.w5, p.b5, d.NC, d.H2, h2, d.H2, lg, d.NC, 0);
  __syncthreads();
  // save activations for backward
  const int PS = d.per_sample();
  for (int i = threadIdx.x; i < SPC * d.emb(); i += kHeadThreads) {
    const int s = i / d.emb(), j = i - s * d.emb();
    if (s0 + s < B) scratch[(size_t)(s0 + s) * PS + j] = es[i];
  }
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

// dst[s][i] = sum_o W[o][i] * src[s][o]   (W^T product; thread per input column i, coalesced over i)
__device__ __forceinline__ void dense_layer_t(const float* __restrict__ W, int n_out, int n_in, const float* src_s, int src_ld,
                                              float* dst_s, int dst_ld) {
  for (int i = threadIdx.x; i < n_in; i += kHeadThreads) {
    float acc[SPC];
#pragma unroll
    for (int s = 0; s < SPC; ++s) acc[s] = 0.f;
#pragma unroll 8
    for (int o = 0; o < n_out; ++o) {
      const float wv = __ldg(W + (size_t)o * n_in + i);
#pragma unroll
      for (int s = 0; s < SPC; ++s) acc[s] = fmaf(src_s[s * src_ld + o], wv, acc[s]);
    }
#pragma unroll
    for (int s = 0; s < SPC; ++s) dst_s[s * dst_ld + i] = acc[s];
  }
}

__global__ void __launch_bounds__(kHeadThreads)
head_bwd_data_kernel(mml_head_params p, HeadDims d, const long long* __restrict__ labels, const uint8_t* __restrict__ drop,
                     float drop_scale, float* __restrict__ scratch, float loss_scale, float* __restrict__ dpooledA,
                     float* __restrict__ dpooledI, int B) {
  extern __shared__ float sm[];
  float* dlog = sm;                       // [SPC][NC]
  float* dh2 = dlog + SPC * d.NC;         // [SPC][H2]
  float* dh1 = dh2 + SPC * d.H2;          // [SPC][H1]
  float* demb = dh1 + SPC * d.H1;         // [SPC][emb]
  float* dpool = demb + SPC * d.emb();    // [SPC][max(FA,FI)]
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
  dense_layer_t(p.w5, d.NC, d.H2, dlog, d.NC, dh2, d.H2);
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
  dense_layer_t(p.w3, d.H2, d.H1, dh2, d.H2, dh1, d.H1);
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
  dense_layer_t(p.w0, d.H1, d.emb(), dh1, d.H1, demb, d.emb());
  __syncthreads();
  for (int i = threadIdx.x; i < SPC * d.emb(); i += kHeadThreads) {
    const int s = i / d.emb(), j = i - s * d.emb();
    if (s0 + s < B) scratch[(size_t)(s0 + s) * PS + d.off_demb() + j] = demb[i];
  }
  const int FM = d.FA > d.FI ? d.FA : d.FI;
  dense_layer_t(p.fcA_w, d.EA, d.FA, demb, d.emb(), dpool, FM);
  __syncthreads();
  for (int i = threadIdx.x; i < SPC * d.FA; i += kHeadThreads) {
    const int s = i / d.FA, k = i - s * d.FA;
    if (s0 + s < B) dpooledA[(size_t)(s0 + s) * d.FA + k] = dpool[s * FM + k];
  }
  __syncthreads();
  dense_layer_t(p.fcI_w, d.EI, d.FI, demb + d.EA, d.emb(), dpool, FM);
  __syncthreads();
  for (int i = threadIdx.x; i < SPC * d.FI; i += kHeadThreads) {
    const int s = i / d.FI, k = i - s * d.FI;
    if (s0 + s < B) dpooledI[(size_t)(s0 + s) * d.FI + k] = dpool[s * FM + k];
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
  const int smem = SPC * (d.FA + d.FI + d.emb() + d.H1 + d.H2 + d.NC) * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(head_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  MML_REQUIRE(ctx, smem <= 96 * 1024, "head_fwd: dims need %d bytes of shared memory", smem);
  cudaStream_t st = (cudaStream_t)stream;
  head_fwd_kernel<<<(B + SPC - 1) / SPC, kHeadThreads, smem, st>>>(*p, d, pooledA, pooledI, (const long long*)labels, dropout_mask,
                                                                  dropout_scale, scratch, logits, pred, B);
  MML_LAUNCHED(ctx);
  if (labels != nullptr) {
    head_loss_kernel<<<1, 256, 0, st>>>(scratch, d.per_sample(), d.off_loss(), B, loss_out);
    MML_LAUNCHED(ctx);
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
  const int FM = d.FA > d.FI ? d.FA : d.FI;
  const int smem = SPC * (d.NC + d.H2 + d.H1 + d.emb() + FM) * (int)sizeof(float);
  MML_REQUIRE(ctx, smem <= 96 * 1024, "head_bwd: dims need %d bytes of shared memory", smem);
  cudaStream_t st = (cudaStream_t)stream;
  MML_REQUIRE(ctx, phases >= 1 && phases <= 3, "head_bwd: phases must be 1 (data), 2 (weights) or 3 (both)");
  if (phases & 1) {
    head_bwd_data_kernel<<<(B + SPC - 1) / SPC, kHeadThreads, smem, st>>>(*p, d, (const long long*)labels, dropout_mask, dropout_scale,
                                                                         scratch, loss_scale, dpooledA, dpooledI, B);
    MML_LAUNCHED(ctx);
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
  head_bwd_weights_kernel<<<fb, 256, 0, st>>>(jobs, B);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_linear_fwd(mml_ctx* ctx, const float* x, const float* w, const float* bias, float* y, int B, int n_in, int n_out, void* stream) {
  MML_REQUIRE(ctx, ctx && x && w && y && B >= 1 && n_in >= 1 && n_out >= 1, "linear_fwd: bad arguments");
  const long long warps = (long long)B * n_out;
  linear_fwd_kernel<<<(unsigned)mml_ceil_div(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, B, n_in, n_out);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int mml_dropout_mask(mml_ctx* ctx, uint8_t* mask, int64_t n, float p, uint64_t seed, const int64_t* step_counter, void* stream) {
  MML_REQUIRE(ctx, ctx && mask && n >= 1 && p >= 0.f && p < 1.f, "dropout_mask: bad arguments");
  int grid = (int)mml_ceil_div(n, 256);
  if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
  dropout_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mask, n, p, seed, (const long long*)step_counter);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

}  // extern "C"
