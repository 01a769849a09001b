// stem.cu -- the ResNet stem convolution (7x7, stride 2, pad 3, C_in = 1 -> 64) forward and weight gradient on tcgen05.
//
// Reference: MML_Suite/models/msa/networks/resnet.py:137 (self.conv1) applied at :205 to the (already masked) fp32
// input, and data/base_dataset.py:71 (sample = original * mask) which is fused here: the kernels read the ORIGINAL
// fp32 input and multiply by the per-sample mask on load (mask_mul: bit-identical to mml_mask_apply_f32), then round
// the product to bf16 (operand precision of every other convolution on the path).
//
// C_in = 1 gives a GEMM-K of only 49, so there is nothing for TMA to tile on the input side.  A tile is `rpt` whole output
// rows of one image (<= 128 pixels).  The 128 builder threads first stage the tile's input patch ((2*rpt+5) x (2Q+6),
// masked, rounded to bf16) in shared memory with coalesced loads -- prefetched one tile ahead into registers -- and then
// each thread gathers the im2col row of "its" pixel: per filter row four 32-bit shared loads give the 7 taps, i.e. one
// 16-byte chunk (GEMM-K index = r*8 + s, 7 x 8 = 56 of 64, zero elsewhere), stored with the SWIZZLE_128B pattern --
// exactly the tile TMA would have produced.  The SAME tile is
//   * the K-major A operand of fprop:   y[px][64 k]  = xcol[px][64] * W[64 k][64]^T
//   * the MN-major B operand of wgrad:  dW[k][tap]  += dy[px][k]^T * xcol[px][tap]     (dy tile via TMA)
// Persistent CTAs (2-3 per SM) loop over tiles with two A buffers and two TMEM accumulators so that building tile i+1,
// the MMA of tile i and the epilogue of tile i-1 overlap.  Warps 0-3: build + epilogue, warp 4: MMA issue.
#include "mml_common.cuh"
#include "mml_ctx.h"

using namespace mml;

namespace {

constexpr int kTileM = 128;
constexpr int kThreadsStem = 160;
constexpr int kTileBytes = kTileM * 128;  // 16 KB
constexpr int kPatchMaxHalves = 2048;     // bf16 elements of one input patch (16 per builder thread)
constexpr int kPatchRegs = kPatchMaxHalves / 128;

struct StemGeom {
  int B, H, W, P, Q;
  int rpt;          // output rows per tile
  int valid;        // rpt * Q pixels per tile (<= 128)
  int tiles_per_img, tiles;
  int PH, PWW;      // patch rows, patch row pitch in 32-bit words (2 bf16 each): 2Q+6 halves
  int patch_halves;
};

inline bool stem_geom(int B, int H, int W, StemGeom* g) {
  g->B = B, g->H = H, g->W = W;
  g->P = (H + 6 - 7) / 2 + 1;
  g->Q = (W + 6 - 7) / 2 + 1;
  if (g->Q > 128 || g->P < 1 || g->Q < 1) return false;
  int rpt = 1;
  for (int d = 1; d <= g->P; ++d)
    if (g->P % d == 0 && d * g->Q <= kTileM && (2 * d + 5) * (2 * g->Q + 6) <= kPatchMaxHalves) rpt = d;
  g->rpt = rpt;
  g->valid = rpt * g->Q;
  g->tiles_per_img = g->P / rpt;
  g->tiles = B * g->tiles_per_img;
  g->PH = 2 * rpt + 5;
  g->PWW = g->Q + 3;
  g->patch_halves = g->PH * 2 * g->PWW;
  return g->patch_halves <= kPatchMaxHalves;
}

// ---- input patch: rows 2*p0-3 .. 2*p0+2*rpt+1, cols -3 .. 2Q+2 of image b, masked, as bf16 ----
// The (row, column) of the 16 patch elements a builder thread owns do not depend on the tile: computed once.
struct PatchMap {
  int off[kPatchRegs];  // r * W + (c - 3)
  int rr[kPatchRegs];   // patch row r, or a huge value for elements outside the patch / outside the image columns
};
__device__ __forceinline__ void patch_map_init(PatchMap& pm, const StemGeom& g) {
  const int pw = 2 * g.PWW;
#pragma unroll
  for (int i = 0; i < kPatchRegs; ++i) {
    const int e = (int)threadIdx.x + i * 128;
    const int r = e / pw, c = e - r * pw;
    const int w = c - 3;
    const bool ok = e < g.patch_halves && w >= 0 && w < g.W;
    pm.off[i] = r * g.W + w;
    pm.rr[i] = ok ? r : (1 << 28);
  }
}
// Issues the 16 global loads of the NEXT tile's patch and nothing else: the values are only consumed by patch_store, after the
// gather + epilogue of the current tile, so all 16 loads are in flight together.  (Round 1 applied the mask here; mask_mul's NaN
// branch made every load's consumer follow it immediately -- 16 serialized DRAM round trips per tile, ~11 of the ~15 k cycles a
// tile took.)
__device__ __forceinline__ float patch_prefetch(float (&reg)[kPatchRegs], const PatchMap& pm, int tile, const float* __restrict__ x,
                                                const float* __restrict__ mask, const StemGeom& g) {
  const int b = tile / g.tiles_per_img;
  const int p0 = (tile - b * g.tiles_per_img) * g.rpt;
  const int hbase = 2 * p0 - 3;
  const float* origin = x + (size_t)b * g.H * g.W + (long long)hbase * g.W;
#pragma unroll
  for (int i = 0; i < kPatchRegs; ++i) reg[i] = ((unsigned)(hbase + pm.rr[i]) < (unsigned)g.H) ? __ldg(origin + pm.off[i]) : 0.f;
  return mask ? __ldg(mask + b) : 1.0f;
}
// sample = original * mask (data/base_dataset.py:71), rounded to bf16, into the shared-memory patch
__device__ __forceinline__ void patch_store(const float (&reg)[kPatchRegs], float mk, bool masked, uint16_t* patch, const StemGeom& g) {
#pragma unroll
  for (int i = 0; i < kPatchRegs; ++i) {
    const int e = (int)threadIdx.x + i * 128;
    if (e < g.patch_halves) {
      const float v = masked ? mask_mul(reg[i], mk) : reg[i];
      patch[e] = (uint16_t)(pack_bf16x2(v, 0.f) & 0xFFFFu);
    }
  }
}
// im2col row of tile pixel `row` -> 128-byte swizzled tile row (7 chunks of [7 taps, 0], last chunk zero)
__device__ __forceinline__ void gather_row(uint32_t tile_addr, int row, const uint32_t* patch_words, int word0, int pww) {
  const uint32_t row_addr = tile_addr + (uint32_t)row * 128u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t* src = patch_words + word0 + r * pww;
    const uint32_t w0 = src[0], w1 = src[1], w2 = src[2], w3 = src[3] & 0x0000FFFFu;
    const uint32_t dst = row_addr + (((uint32_t)r ^ (uint32_t)(row & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
  }
}

struct alignas(64) StemMaps {
  CUtensorMap io;  // 4-D {64, Q, P, B}, box {64, Q, rpt, 1}: fprop stores y through it, wgrad loads dy through it
};

// fprop smem: [A0 | A1 | W | staging | patch0 | patch1 | barriers]
constexpr int kPatchBytes = kPatchMaxHalves * 2;
constexpr int kFOffW = 2 * kTileBytes;
constexpr int kFOffStage = kFOffW + 64 * 128;
constexpr int kFOffPatch = kFOffStage + kTileBytes;
constexpr int kFOffBars = kFOffPatch + 2 * kPatchBytes;
constexpr int kFBytes = kFOffBars + 1024 + 1024;

__global__ void __launch_bounds__(kThreadsStem, 3)
stem_fprop_tc_kernel(const __grid_constant__ StemMaps maps, const float* __restrict__ x, const float* __restrict__ mask,
                     const float* __restrict__ w, double* __restrict__ stats, StemGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bars = smem_base + kFOffBars;
  auto a_full = [&](int b) { return bars + 8u * b; };
  auto mma_done = [&](int b) { return bars + 8u * (2 + b); };
  const uint32_t tmem_slot = bars + 8u * 4;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kFOffBars + 8 * 4);

  {  // tile rows >= valid and chunk 7 of every row are never written by the builders: zero both A buffers once
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* base = reinterpret_cast<uint4*>(smem_gen);
    for (int i = threadIdx.x; i < 2 * kTileBytes / 16; i += blockDim.x) base[i] = z;
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.io);
    for (int b = 0; b < 2; ++b) {
      mbar_init(a_full(b), 128);
      mbar_init(mma_done(b), 1);
    }
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<128>(tmem_slot);
  pdl_wait();  // the weights below were written by an earlier kernel (Adam)
  if (threadIdx.x < 64) {  // weights: row k, chunk r = [w[k][r][0..6], 0], chunk 7 = 0; K-major SWIZZLE_128B
    const int k = threadIdx.x;
    const uint32_t row_addr = smem_base + kFOffW + (uint32_t)k * 128u;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      if (r < 7) {
        const float* wr = w + k * 49 + r * 7;
        c0 = pack_bf16x2(wr[0], wr[1]), c1 = pack_bf16x2(wr[2], wr[3]), c2 = pack_bf16x2(wr[4], wr[5]), c3 = pack_bf16x2(wr[6], 0.f);
      }
      const uint32_t dst = row_addr + (((uint32_t)r ^ (uint32_t)(k & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_launch_dependents();
  const int n_my = (g.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 4) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      const uint32_t b_addr = smem_base + kFOffW;
      for (int i = 0; i < n_my; ++i) {
        const int buf = i & 1;
        mbar_wait(a_full(buf), (uint32_t)(i >> 1) & 1u);
        tc_fence_after();
        const uint32_t a_addr = smem_base + buf * kTileBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + buf * 64, umma_desc_sw128(a_addr + k * 32, 16, 1024), umma_desc_sw128(b_addr + k * 32, 16, 1024), idesc, k != 0);
        umma_commit(mma_done(buf));
      }
    }
  } else {
    const int row = threadIdx.x;  // 0..127: tile row == TMEM lane
    const int pr = row / g.Q, q = row - pr * g.Q;
    const int word0 = 2 * pr * g.PWW + q;
    const bool live = row < g.valid;
    float4 st_acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto epilogue = [&](int j) {
      const int buf = j & 1;
      const int tile = (int)blockIdx.x + j * (int)gridDim.x;
      mbar_wait(mma_done(buf), (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 64);
      tmem_ld_32x32(taddr, r0);
      tmem_ld_32x32(taddr + 32, r1);
      tmem_ld_wait();
      if (threadIdx.x == 0) tma_store_wait_read<0>();  // the previous tile's TMA store has finished reading the staging buffer
      named_bar_sync(1, 128);
      const uint32_t stage = smem_base + kFOffStage;
      const uint32_t row_addr = stage + row * 128;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const uint32_t* src = jj < 4 ? &r0[8 * jj] : &r1[8 * (jj - 4)];
        const uint32_t dst = row_addr + (((uint32_t)jj ^ (uint32_t)(row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16x2(__uint_as_float(src[0]), __uint_as_float(src[1]))),
                     "r"(pack_bf16x2(__uint_as_float(src[2]), __uint_as_float(src[3]))),
                     "r"(pack_bf16x2(__uint_as_float(src[4]), __uint_as_float(src[5]))),
                     "r"(pack_bf16x2(__uint_as_float(src[6]), __uint_as_float(src[7])))
                     : "memory");
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (threadIdx.x == 0) {
        const int b = tile / g.tiles_per_img;
        const int p0 = (tile - b * g.tiles_per_img) * g.rpt;
        tma_store_4d(&maps.io, stage, 0, 0, p0, b);  // box = exactly the tile's rpt x Q pixels
        tma_store_commit();
      }
      if (stats != nullptr) {
        // BatchNorm partials of the STORED (bf16-rounded) tile: thread = (channel pair wc, row quarter rq), one 32-bit shared load
        // per row, accumulated in registers over all tiles of this CTA (one atomic per channel and CTA at the end)
        const int wc = threadIdx.x & 31, rq = threadIdx.x >> 5;
        const int rend = min(g.valid - rq * 32, 32);
        const uint8_t* colp = smem_gen + kFOffStage + (wc & 3) * 4;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
        for (int b = 0; b < rend; ++b) {
          const int r = rq * 32 + b;
          const uint32_t v = *reinterpret_cast<const uint32_t*>(colp + r * 128 + ((((uint32_t)wc >> 2) ^ (uint32_t)(r & 7)) << 4));
          const float a = bf16_lo(v), c = bf16_hi(v);
          s0 += a, s1 += c;
          q0 = fmaf(a, a, q0), q1 = fmaf(c, c, q1);
        }
        st_acc.x += s0, st_acc.y += s1, st_acc.z += q0, st_acc.w += q1;
      }
    };
    float reg[kPatchRegs];
    PatchMap pm;
    patch_map_init(pm, g);
    float mk = 1.0f;
    if (n_my > 0) {
      mk = patch_prefetch(reg, pm, (int)blockIdx.x, x, mask, g);
      patch_store(reg, mk, mask != nullptr, reinterpret_cast<uint16_t*>(smem_gen + kFOffPatch), g);
    }
    named_bar_sync(1, 128);
    for (int i = 0; i < n_my; ++i) {
      const bool more = i + 1 < n_my;
      if (more) mk = patch_prefetch(reg, pm, (int)blockIdx.x + (i + 1) * (int)gridDim.x, x, mask, g);  // loads in flight during gather + epilogue
      if (live) gather_row(smem_base + (i & 1) * kTileBytes, row, reinterpret_cast<const uint32_t*>(smem_gen + kFOffPatch + (i & 1) * kPatchBytes), word0, g.PWW);
      fence_proxy_async_smem();
      tc_fence_before();  // orders this thread's earlier TMEM loads before the MMA that will overwrite that accumulator
      mbar_arrive(a_full(i & 1));
      if (i >= 1) epilogue(i - 1);
      if (more) patch_store(reg, mk, mask != nullptr, reinterpret_cast<uint16_t*>(smem_gen + kFOffPatch + ((i + 1) & 1) * kPatchBytes), g);
      named_bar_sync(1, 128);
    }
    if (n_my >= 1) epilogue(n_my - 1);
    if (stats != nullptr) {
      float4* sc = reinterpret_cast<float4*>(smem_gen + kFOffStage);  // the staging buffer is free once the last TMA store has read it
      if (threadIdx.x == 0) tma_store_wait_read<0>();
      named_bar_sync(1, 128);
      sc[threadIdx.x] = st_acc;
      named_bar_sync(1, 128);
      if (threadIdx.x < 32) {
        const float4 a = sc[threadIdx.x], b2 = sc[threadIdx.x + 32], c2 = sc[threadIdx.x + 64], d2 = sc[threadIdx.x + 96];
        const int c0 = 2 * threadIdx.x;
        stat_add(stats, 64, (int)blockIdx.x, c0, a.x + b2.x + c2.x + d2.x, a.z + b2.z + c2.z + d2.z);
        stat_add(stats, 64, (int)blockIdx.x, c0 + 1, a.y + b2.y + c2.y + d2.y, a.w + b2.w + c2.w + d2.w);
      }
    }
    if (threadIdx.x == 0) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<128>(tmem_base);
}

// wgrad smem: [DY0 | DY1 | X0 | X1 | patch0 | patch1 | barriers]
//
// The MMA is 128 x 64 x 16 with an MN-major A operand whose two 64-row halves sit LBO bytes apart: rows 0..63 = the 64 channels of the
// dy tile, rows 64..127 = the im2col tile ITSELF (LBO = X - DY), whose otherwise unused last column holds a constant 1.  One pass over
// dy therefore yields, per CTA,
//     acc[k][t]       = sum_p dy[p][k] * xcol[p][t]            (the weight gradient w.r.t. whatever dy is)
//     acc[64 + t'][t] = sum_p xcol[p][t'] * xcol[p][t]         (Gram matrix of the input patches)
//     acc[127][t]     = sum_p xcol[p][t]                        (tap sums)
// which is what folds the BatchNorm backward's second pass into this kernel (mml_stem_wgrad_bn below): with dy = g, the gradient after
// ReLU / pooling but BEFORE the BatchNorm correction,
//     dW[k][t] = gamma_k*invstd_k * ( acc[k][t] - mean(g)_k * S[t] - mean(g*xhat)_k * invstd_k * ( sum_t' W[k][t'] X2[t'][t] - mu_k * S[t] ) )
// because the stem output is linear in the patches (raw[p][k] = sum_t' W[k][t'] xcol[p][t']).  The 102 MB dx tensor is never formed.
constexpr int kWOffX = 2 * kTileBytes;
constexpr int kWOffPatch = 4 * kTileBytes;
constexpr int kWOffBars = kWOffPatch + 2 * kPatchBytes;
constexpr int kWBytes = kWOffBars + 1024 + 1024;
constexpr int kOnesCol = 63;  // column of the im2col tile that carries the constant 1 (columns r*8+7 and 56..63 are padding)

__global__ void __launch_bounds__(kThreadsStem, 2)
stem_wgrad_tc_kernel(const __grid_constant__ StemMaps maps, const float* __restrict__ x, const float* __restrict__ mask,
                     float* __restrict__ ws, StemGeom g, int rows_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bars = smem_base + kWOffBars;
  auto full = [&](int b) { return bars + 8u * b; };
  auto freeb = [&](int b) { return bars + 8u * (2 + b); };
  const uint32_t all_done = bars + 8u * 4;
  const uint32_t tmem_slot = bars + 8u * 5;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kWOffBars + 8 * 5);

  {  // rows >= valid of every tile and the padding columns of the im2col tiles are never written: zero everything once ...
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* base = reinterpret_cast<uint4*>(smem_gen);
    for (int i = threadIdx.x; i < kWOffPatch / 16; i += blockDim.x) base[i] = z;
    __syncthreads();
    // ... and put the constant 1 (bf16 0x3F80) into column kOnesCol of every live row of both im2col tiles (gather_row never touches
    // the last 16-byte chunk of a row, so it stays)
    for (int i = threadIdx.x; i < 2 * g.valid; i += blockDim.x) {
      const int b = i / g.valid, row = i - b * g.valid;
      uint8_t* rowp = smem_gen + kWOffX + b * kTileBytes + row * 128;
      *reinterpret_cast<uint16_t*>(rowp + (((kOnesCol >> 3) ^ (row & 7)) << 4) + (kOnesCol & 7) * 2) = 0x3F80;
    }
    fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.io);
    for (int b = 0; b < 2; ++b) {
      mbar_init(full(b), 129);  // 128 builder rows + the arrive.expect_tx of the dy TMA
      mbar_init(freeb(b), 1);
    }
    mbar_init(all_done, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<64>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_sync();
  const int n_my = (g.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 4) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      const int ksteps = (g.valid + 15) >> 4;
      for (int i = 0; i < n_my; ++i) {
        const int buf = i & 1;
        mbar_wait(full(buf), (uint32_t)(i >> 1) & 1u);
        tc_fence_after();
        const uint32_t a_addr = smem_base + buf * kTileBytes;           // [dy^T ; xcol^T], MN-major: M rows 64..127 start kWOffX further
        const uint32_t b_addr = smem_base + kWOffX + buf * kTileBytes;  // xcol, MN-major, N = 64 (r*8+s)
        for (int ks = 0; ks < ksteps; ++ks)
          umma_bf16(tmem_base, umma_desc_sw128(a_addr + ks * 2048, kWOffX, 1024), umma_desc_sw128(b_addr + ks * 2048, kTileBytes, 1024), idesc,
                    (i | ks) != 0);
        umma_commit(freeb(buf));
      }
      umma_commit(all_done);
    }
  } else {
    const int row = threadIdx.x;
    const int pr = row / g.Q, q = row - pr * g.Q;
    const int word0 = 2 * pr * g.PWW + q;
    const bool live = row < g.valid;
    float reg[kPatchRegs];
    PatchMap pm;
    patch_map_init(pm, g);
    float mk = 1.0f;
    if (n_my > 0) {
      mk = patch_prefetch(reg, pm, (int)blockIdx.x, x, mask, g);
      patch_store(reg, mk, mask != nullptr, reinterpret_cast<uint16_t*>(smem_gen + kWOffPatch), g);
    }
    named_bar_sync(1, 128);
    for (int i = 0; i < n_my; ++i) {
      const int buf = i & 1;
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const bool more = i + 1 < n_my;
      if (more) mk = patch_prefetch(reg, pm, tile + (int)gridDim.x, x, mask, g);
      if (i >= 2) mbar_wait(freeb(buf), (uint32_t)((i >> 1) - 1) & 1u);
      if (threadIdx.x == 0) {
        const int b = tile / g.tiles_per_img;
        const int p0 = (tile - b * g.tiles_per_img) * g.rpt;
        mbar_arrive_expect_tx(full(buf), (uint32_t)g.valid * 128u);
        tma_load_4d(&maps.io, full(buf), smem_base + buf * kTileBytes, 0, 0, p0, b);
      }
      if (live) gather_row(smem_base + kWOffX + buf * kTileBytes, row, reinterpret_cast<const uint32_t*>(smem_gen + kWOffPatch + buf * kPatchBytes), word0, g.PWW);
      fence_proxy_async_smem();
      mbar_arrive(full(buf));
      if (more) patch_store(reg, mk, mask != nullptr, reinterpret_cast<uint16_t*>(smem_gen + kWOffPatch + ((i + 1) & 1) * kPatchBytes), g);
      named_bar_sync(1, 128);
    }
    // dW[k][r][s] lives in accumulator row k, column r*8+s; rows 64..127 (rows_out == 128) = Gram matrix / tap sums, same columns
    float* out = ws + ((size_t)blockIdx.x * rows_out + row) * 49;
    if (n_my >= 1) {
      mbar_wait(all_done, 0);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
      tmem_ld_32x32(taddr, r0);
      tmem_ld_32x32(taddr + 32, r1);
      tmem_ld_wait();
      if (row < rows_out) {
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
          for (int s2 = 0; s2 < 7; ++s2) {
            const int col = r * 8 + s2;
            out[r * 7 + s2] = __uint_as_float(col < 32 ? r0[col] : r1[col - 32]);
          }
      }
    } else if (row < rows_out) {
      for (int t = 0; t < 49; ++t) out[t] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<64>(tmem_base);
}

__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ ws, int parts, float* __restrict__ dw) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 49) return;
  float a = 0.f;
  for (int p = 0; p < parts; ++p) a += ws[(size_t)p * (64 * 49) + i];
  dw[i] = a;
}

// Fixed-order fp64 sums of the per-CTA [128][49] partials (rows 0..63: sum g (x) xcol; rows 64..127: Gram matrix, row 127 = tap sums).
// Eight lanes per output: lane l sums partials l, l+8, ... (four independent loads in flight), then a fixed shuffle tree -- the order
// of the additions never depends on timing.  (One thread per output walking ~300 partials took 17 us of exposed latency at the very
// end of the step.)
__global__ void __launch_bounds__(256) stem_wgrad_bn_sum_kernel(const float* __restrict__ ws, int parts, double* __restrict__ red) {
  pdl_sync();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t >> 3, l = t & 7;
  const bool live = i < 128 * 49;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (live) {
    const float* src = ws + i;
    int p = l;
    for (; p + 24 < parts; p += 32) {
      const float v0 = __ldcg(src + (size_t)p * (128 * 49)), v1 = __ldcg(src + (size_t)(p + 8) * (128 * 49));
      const float v2 = __ldcg(src + (size_t)(p + 16) * (128 * 49)), v3 = __ldcg(src + (size_t)(p + 24) * (128 * 49));
      a0 += (double)v0, a1 += (double)v1, a2 += (double)v2, a3 += (double)v3;
    }
    for (; p < parts; p += 8) a0 += (double)__ldcg(src + (size_t)p * (128 * 49));
  }
  double a = (a0 + a1) + (a2 + a3);
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  a += __shfl_xor_sync(0xffffffffu, a, 4);
  if (live && l == 0) red[i] = a;
}

// dW = gamma*invstd*( G - mean(g)*S - mean(g*xhat)*invstd*( W X2 - mu*S ) ), plus dgamma / dbeta (what bn_bwd_apply's block 0 stores).
// One output per thread; every CTA stages the 49 x 49 Gram matrix and the tap sums in shared memory, and everything an output needs
// from global memory is requested before the first use (two round trips per CTA in total).
__global__ void __launch_bounds__(256)
stem_wgrad_bn_combine_kernel(const double* __restrict__ red, const float* __restrict__ w, const double* __restrict__ bstat,
                             const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma, double inv_count,
                             float* __restrict__ dw, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_sync();
  __shared__ double x2[49][49], S[49];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < 64 * 49;
  const int k = live ? i / 49 : 0, t = live ? i - k * 49 : 0;
  for (int j = threadIdx.x; j < 49 * 49; j += blockDim.x) {
    const int tp = j / 49, tt = j - tp * 49;
    x2[tp][tt] = __ldcg(red + (64 + (tp / 7) * 8 + tp % 7) * 49 + tt);
  }
  for (int tt = threadIdx.x; tt < 49; tt += blockDim.x) S[tt] = __ldcg(red + (64 + kOnesCol) * 49 + tt);
  double sg, sgx;
  stat_load(bstat, 64, k, sg, sgx);
  const double G = __ldcg(red + k * 49 + t);
  const double is = (double)invstd[k], mu = (double)mean[k], ga = (double)gamma[k];
  float wk[49];
#pragma unroll
  for (int tp = 0; tp < 49; ++tp) wk[tp] = __ldg(w + k * 49 + tp);
  __syncthreads();
  if (!live) return;
  double r = 0.0;
#pragma unroll
  for (int tp = 0; tp < 49; ++tp) r += (double)bf16_lo(pack_bf16x2(wk[tp], 0.f)) * x2[tp][t];  // the bf16 operand the forward multiplied with
  const double k1 = sg * inv_count, k2 = sgx * inv_count;
  dw[i] = (float)(ga * is * (G - k1 * S[t] - k2 * is * (r - mu * S[t])));
  if (t == 0) {
    if (dbeta) dbeta[k] = (float)sg;
    if (dgamma) dgamma[k] = (float)sgx;
  }
}

// several CTAs per SM: each has only 4 builder warps, latency is hidden across CTAs (fprop 66 KB smem -> 3, wgrad 106 KB -> 2)
int stem_ctas(const mml_ctx* ctx, int tiles, int per_sm) {
  int n = ctx->sm_count * per_sm;  // (applying the SM budget here was measured and costs more than it frees: these kernels are issue-bound)
  if (n > tiles) n = tiles;
  return n < 1 ? 1 : n;
}

int encode_px_map(mml_ctx* ctx, CUtensorMap* map, const void* ptr, const StemGeom& g) {
  cuuint64_t dims[4] = {64, (cuuint64_t)g.Q, (cuuint64_t)g.P, (cuuint64_t)g.B};
  cuuint64_t strides[3] = {128, (cuuint64_t)g.Q * 128, (cuuint64_t)g.P * g.Q * 128};
  cuuint32_t box[4] = {64, (cuuint32_t)g.Q, (cuuint32_t)g.rpt, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ctx->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return mml_set_error(ctx, MML_ERR_CUDA, "stem: cuTensorMapEncodeTiled failed: %d", (int)r);
  return MML_OK;
}

}  // namespace

extern "C" {

int mml_stem_fprop(mml_ctx* ctx, const float* x, const float* mask, const float* w, uint16_t* y, double* stats, int B, int H,
                   int W, void* stream) {
  MML_REQUIRE(ctx, ctx && x && w && y, "stem_fprop: null pointer");
  MML_REQUIRE(ctx, B >= 1 && H >= 1 && W >= 1, "stem_fprop: bad dims");
  StemGeom g;
  MML_REQUIRE(ctx, stem_geom(B, H, W, &g), "stem: output width %d not supported (max 128)", (W - 1) / 2 + 1);
  StemMaps maps;
  int rc = encode_px_map(ctx, &maps.io, y, g);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(stem_fprop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFBytes));
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(stem_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWBytes));
    configured = true;
  }
  MML_LAUNCH(ctx, stem_fprop_tc_kernel, stem_ctas(ctx, g.tiles, 3), kThreadsStem, kFBytes, (cudaStream_t)stream, maps, x, mask, w, stats, g);
  return MML_OK;
}

int64_t mml_stem_wgrad_workspace(const mml_ctx* ctx, int B, int H, int W) {
  StemGeom g;
  if (!ctx || !stem_geom(B, H, W, &g)) return 0;
  // [ctas][128][49] fp32 partials (the BatchNorm-folding variant stores all 128 accumulator rows) + [128][49] fp64 sums
  return (int64_t)stem_ctas(ctx, g.tiles, 2) * 128 * 49 * sizeof(float) + 128 * 49 * sizeof(double);
}

int mml_stem_wgrad(mml_ctx* ctx, const float* x, const float* mask, const uint16_t* dy, float* dw, float* workspace,
                   int64_t workspace_bytes, int B, int H, int W, void* stream) {
  MML_REQUIRE(ctx, ctx && x && dy && dw && workspace, "stem_wgrad: null pointer");
  MML_REQUIRE(ctx, B >= 1 && H >= 1 && W >= 1, "stem_wgrad: bad dims");
  MML_REQUIRE(ctx, workspace_bytes >= mml_stem_wgrad_workspace(ctx, B, H, W), "stem_wgrad: workspace too small");
  StemGeom g;
  MML_REQUIRE(ctx, stem_geom(B, H, W, &g), "stem: output width %d not supported (max 128)", (W - 1) / 2 + 1);
  StemMaps maps;
  int rc = encode_px_map(ctx, &maps.io, dy, g);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(stem_fprop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFBytes));
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(stem_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWBytes));
    configured = true;
  }
  const int ctas = stem_ctas(ctx, g.tiles, 2);
  cudaStream_t st = (cudaStream_t)stream;
  MML_LAUNCH(ctx, stem_wgrad_tc_kernel, ctas, kThreadsStem, kWBytes, st, maps, x, mask, workspace, g, 64);
  MML_LAUNCH(ctx, stem_wgrad_reduce_kernel, (64 * 49 + 255) / 256, 256, 0, st, workspace, ctas, dw);
  return MML_OK;
}

int mml_stem_wgrad_bn(mml_ctx* ctx, const float* x, const float* mask, const uint16_t* g_bf16, const float* w, const double* bstat,
                      const float* mean, const float* invstd, const float* gamma, float* dgamma, float* dbeta, float* dw, float* workspace,
                      int64_t workspace_bytes, int B, int H, int W, void* stream) {
  MML_REQUIRE(ctx, ctx && x && g_bf16 && w && bstat && mean && invstd && gamma && dw && workspace, "stem_wgrad_bn: null pointer");
  MML_REQUIRE(ctx, B >= 1 && H >= 1 && W >= 1, "stem_wgrad_bn: bad dims");
  MML_REQUIRE(ctx, workspace_bytes >= mml_stem_wgrad_workspace(ctx, B, H, W) && ((uintptr_t)workspace & 7) == 0, "stem_wgrad_bn: workspace too small / misaligned");
  StemGeom g;
  MML_REQUIRE(ctx, stem_geom(B, H, W, &g), "stem: output width %d not supported (max 128)", (W - 1) / 2 + 1);
  StemMaps maps;
  int rc = encode_px_map(ctx, &maps.io, g_bf16, g);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(stem_fprop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFBytes));
    MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(stem_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWBytes));
    configured = true;
  }
  const int ctas = stem_ctas(ctx, g.tiles, 2);
  cudaStream_t st = (cudaStream_t)stream;
  double* red = reinterpret_cast<double*>(workspace + (size_t)ctas * 128 * 49);  // ctas * 128 * 49 floats: a multiple of 8 bytes
  MML_LAUNCH(ctx, stem_wgrad_tc_kernel, ctas, kThreadsStem, kWBytes, st, maps, x, mask, workspace, g, 128);
  MML_LAUNCH(ctx, stem_wgrad_bn_sum_kernel, (128 * 49 * 8 + 255) / 256, 256, 0, st, (const float*)workspace, ctas, red);
  const double inv_count = 1.0 / ((double)B * g.P * g.Q);
  MML_LAUNCH(ctx, stem_wgrad_bn_combine_kernel, (64 * 49 + 255) / 256, 256, 0, st, (const double*)red, w, bstat, mean, invstd, gamma, inv_count, dw, dgamma, dbeta);
  return MML_OK;
}

}  // extern "C"
