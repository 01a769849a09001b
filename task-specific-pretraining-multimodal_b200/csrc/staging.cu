// staging.cu -- device side of the input path in front of the encoders: the per-(sample, modality)
// missing-modality mask draw, the per-batch mask gather, and the uint8 image -> colormap -> luminance -> fp32 conversion.
//
// Reference ops replaced (all of them CPU work inside the reference's DataLoader workers):
//   * MML_Suite/data/base_dataset.py:46-59  _initialise_missing_masks: create_missing_mask(n_modalities, n, [P(present)])
//     -- one independent Bernoulli(P(present)) per sample and modality, drawn once per pattern at dataset construction;
//   * MML_Suite/data/avmnist.py:193-224     __getitem__: the mask of sample ``idx`` is looked up per item;
//   * MML_Suite/data/avmnist.py:188-191     _load_image: np.uint8(cm.gist_earth(img) * 255) -> PIL convert("L") ->
//     PILToTensor -> ToDtype(float32, scale=True).  For uint8 pixels that whole chain is a function of the pixel value alone,
//     i.e. a 256-entry table (built on the host from the colormap, mml_b200/data.py::luma_lut), so the device reads 1 byte per
//     pixel from the host instead of 4.
//
// All three are HBM / PCIe-bound byte work: 16-byte accesses, grid = a multiple of the SM count, no tensor cores.
#include "mml_common.cuh"
#include "mml_ctx.h"

using namespace mml;

namespace {

constexpr int kThreads = 256;

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11): counter-based, so sample i of stream s is
// a pure function of (seed, s, i) -- any launch shape, any shard of the sample range and the CPU oracle produce the same bits.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0, k.y += W1;
  }
  return c;
}

// 24 random bits -> fp32 in [0, 1) (exact), present iff u < P(present): P = 1 always keeps, P = 0 always drops
__device__ __forceinline__ float bernoulli_keep(uint32_t bits, float p_present) {
  const float u = (float)(bits >> 8) * 5.9604644775390625e-08f;  // 2^-24
  return u < p_present ? 1.0f : 0.0f;
}

// masks[m][i] for i in [first, first + count): counter = (i / 4 as 64 bits, modality m, stream), key = seed, lane = i % 4
__global__ void mask_draw_kernel(const float* __restrict__ p_present, float* __restrict__ masks, int n_mod, long long first,
                                 long long count, long long ld, uint32_t seed_lo, uint32_t seed_hi, uint32_t stream_id) {
  pdl_sync();
  const long long q0 = first >> 2, q1 = (first + count + 3) >> 2;  // quads of four consecutive samples
  const long long quads = q1 - q0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < quads * n_mod; t += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(t / quads);
    const long long q = q0 + t % quads;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)((unsigned long long)q >> 32), (uint32_t)m, stream_id),
                                  make_uint2(seed_lo, seed_hi));
    const float p = p_present[m];
    const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
    float* row = masks + (long long)m * ld;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const long long i = q * 4 + l;
      if (i >= first && i < first + count) row[i - first] = bernoulli_keep(bits[l], p);
    }
  }
}

// out[m][b] = masks[m][idx[b]]
__global__ void mask_gather_kernel(const float* __restrict__ masks, const long long* __restrict__ idx, float* __restrict__ out, int n_mod,
                                   long long num_samples, long long batch, int* __restrict__ bad) {
  pdl_sync();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < batch * n_mod; t += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(t / batch);
    const long long b = t % batch, i = idx[b];
    if (i < 0 || i >= num_samples) {
      if (bad) atomicExch(bad, 1);
      out[t] = 0.0f;
      continue;
    }
    out[t] = masks[(long long)m * num_samples + i];
  }
}

// dst[i] = lut[src[i]]: 16 source bytes -> four float4 stores per thread and iteration; the table sits in shared memory
__global__ void u8_lut_kernel(const uint8_t* __restrict__ src, const float* __restrict__ lut, float* __restrict__ dst, long long n) {
  __shared__ float s_lut[256];
  pdl_sync();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const long long n16 = n >> 4;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n16; v += (long long)gridDim.x * blockDim.x) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + v);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float4* o = reinterpret_cast<float4*>(dst) + v * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = make_float4(s_lut[w[j] & 0xFF], s_lut[(w[j] >> 8) & 0xFF], s_lut[(w[j] >> 16) & 0xFF], s_lut[w[j] >> 24]);
  }
  // ragged tail (n % 16 bytes)
  for (long long i = (n16 << 4) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = s_lut[src[i]];
}

inline int staging_grid(const mml_ctx* ctx, long long work_items) {
  long long blocks = mml_ceil_div(work_items, kThreads);
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : (int)blocks;
}

}  // namespace

extern "C" {

int mml_missing_mask_draw(mml_ctx* ctx, const float* p_present, float* masks, int n_modalities, int64_t first_sample, int64_t count,
                          int64_t ld, uint64_t seed, uint32_t stream_id, void* stream) {
  MML_REQUIRE(ctx, ctx && n_modalities >= 1 && first_sample >= 0 && count >= 0 && ld >= count, "missing_mask_draw: bad arguments");
  if (count == 0) return MML_OK;  // an empty shard: nothing to draw (the buffers may be NULL)
  MML_REQUIRE(ctx, p_present && masks, "missing_mask_draw: NULL buffer");
  const long long quads = ((first_sample + count + 3) >> 2) - (first_sample >> 2);
  MML_LAUNCH(ctx, mask_draw_kernel, staging_grid(ctx, quads * n_modalities), kThreads, 0, (cudaStream_t)stream, p_present, masks, n_modalities,
             (long long)first_sample, (long long)count, (long long)ld, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id);
  return MML_OK;
}

int mml_missing_mask_gather(mml_ctx* ctx, const float* masks, const int64_t* sample_idx, float* out, int n_modalities, int64_t num_samples,
                            int64_t batch, int* bad_index_flag, void* stream) {
  MML_REQUIRE(ctx, ctx && n_modalities >= 1 && num_samples >= 0 && batch >= 0, "missing_mask_gather: bad arguments");
  if (batch == 0) return MML_OK;
  MML_REQUIRE(ctx, masks && sample_idx && out, "missing_mask_gather: NULL buffer");
  MML_LAUNCH(ctx, mask_gather_kernel, staging_grid(ctx, batch * n_modalities), kThreads, 0, (cudaStream_t)stream, masks,
             reinterpret_cast<const long long*>(sample_idx), out, n_modalities, (long long)num_samples, (long long)batch, bad_index_flag);
  return MML_OK;
}

int mml_stage_u8_lut_f32(mml_ctx* ctx, const uint8_t* src, const float* lut256, float* dst, int64_t n, void* stream) {
  MML_REQUIRE(ctx, ctx && n >= 0, "stage_u8_lut: bad arguments");
  if (n == 0) return MML_OK;
  MML_REQUIRE(ctx, src && lut256 && dst, "stage_u8_lut: NULL buffer");
  MML_REQUIRE(ctx, ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "stage_u8_lut: src and dst must be 16-byte aligned");
  MML_LAUNCH(ctx, u8_lut_kernel, staging_grid(ctx, mml_ceil_div(n, 16)), kThreads, 0, (cudaStream_t)stream, src, lut256, dst, (long long)n);
  return MML_OK;
}

}  // extern "C"
