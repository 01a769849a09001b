// conv_tc.cu -- tcgen05 / TMEM implicit-GEMM convolutions fed by TMA (fprop, dgrad, wgrad) for sm_100a.
//
// Replaces, on the late-fusion hot path, every 3x3 / 1x1 nn.Conv2d of the reference's ResNet encoders
// (MML_Suite/models/msa/networks/resnet.py:25,30,176) and their autograd (cuDNN fp32 in the reference).
//
// Formulation.  Activations are NHWC bf16.  A convolution is a sum over filter taps of shifted GEMMs:
//     out[n, a, b, :] = sum_t  in_view[t.map][n, a + t.dh, b + t.dw, :] @ Wtap[t.widx]
// where every in_view is a plain 4-D TMA tensor map {C, W, H, N} over the activation tensor.  Stride-2 convolutions
// use "phase views" (base offset + doubled strides), so the kernel never needs im2col-mode or element strides:
// zero padding is TMA out-of-bounds fill.  An output tile is a box {64 ch, Wb = OutW, Hb, Nb} of <= 128 pixels; its
// rows land in shared memory as 128-byte rows with SWIZZLE_128B, which is exactly the canonical K-major UMMA operand
// layout.  The same box of the output tensor is written back by a TMA store from a swizzled staging buffer.
//   fprop           : in = x, W = w[K][R][S][C],            out = y      (+ BatchNorm sum / sum-of-squares, fp64 atomics)
//   dgrad           : in = dy, W = the same w[K][R][S][C] read as an MN-major B operand, out = dx (stride 2: one launch per
//                     output phase)
//   wgrad           : dW[k][t][c] += sum_pixels dy[pix][k] * in_view[t][pix][c]   (both operands MN-major)
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread tcgen05.mma issuer,
// warps 2-5 = epilogue (TMEM -> registers -> bf16 -> swizzled smem -> TMA store, BatchNorm partials from smem).
#include "mml_common.cuh"
#include "mml_ctx.h"

using namespace mml;

namespace {

struct Tap {
  int8_t map;   // which input view
  int8_t dh;    // row offset in that view
  int8_t dw;    // column offset
  int8_t widx;  // filter tap index r*S+s in the weight matrix
};

constexpr int kMaxTaps = 9;
constexpr int kMaxViews = 4;
constexpr int kBoxBytes = 128 * 128;  // one 128-row x 64-channel bf16 box

struct alignas(64) IgemmMaps {
  CUtensorMap in[kMaxViews];
  CUtensorMap w;
  CUtensorMap out;
};

struct IgemmParams {
  int num_taps;
  int c_chunks;  // Cin / 64  (GEMM-K chunks per tap)
  int w_tap_stride;  // elements between taps along the weight map's inner dimension
  int cin;
  int tiles_h;     // OutH / Hb
  int Hb, Nb;      // tile = {OutW, Hb, Nb}
  int valid_rows;  // OutW * Hb * Nb  (<= 128)
  int cout;
  double* stats;  // [16 slots][cout][2] (sum, sum of squares) of the stored output, accumulated with fp64 atomics; or nullptr
  Tap taps[kMaxTaps];
};

template <int BLOCK_N, int STAGES>
struct IgemmSmem {
  static constexpr int kA = kBoxBytes;          // 128 rows x 128 B
  static constexpr int kB = BLOCK_N * 128;      // BLOCK_N rows x 128 B
  static constexpr int kStage = kA + kB;
  static constexpr int kStaging = 2 * kBoxBytes;
  static constexpr int kOffStaging = STAGES * kStage;
  static constexpr int kOffBars = kOffStaging + kStaging;
  static constexpr int kBytes = kOffBars + 1024 /*barriers, tmem slot, stats scratch*/ + 1024 /*alignment slack*/;
};

// ------------------------------------------------------------------------------------------------------------------
// fprop / dgrad kernel
// ------------------------------------------------------------------------------------------------------------------
// B_MN = false: weights are [N rows][taps*K inner] (fprop: W[k][r][s][c]), B operand K-major.
// B_MN = true : weights are [K rows][taps*N inner] (dgrad reads the SAME K,R,S,C tensor: GEMM-K = k, GEMM-N = c), B operand
//               MN-major: 64-row x 64-element boxes, 8 KB each, one per 64 output channels.
template <int BLOCK_N, int STAGES, int MIN_BLOCKS, bool B_MN>
__global__ void __launch_bounds__(192, MIN_BLOCKS)
conv_igemm_kernel(const __grid_constant__ IgemmMaps maps, const IgemmParams p) {
  using L = IgemmSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t bars = smem_base + L::kOffBars;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kOffBars + 8 * (2 * STAGES + 1));
  float2* stat_scratch = reinterpret_cast<float2*>(smem_gen + L::kOffBars + 256);  // 64 x float2

  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int n_blk = m_tile / p.tiles_h;
  const int h_blk = m_tile - n_blk * p.tiles_h;
  const int a0 = h_blk * p.Hb;
  const int n0 = n_blk * p.Nb;
  const int iters = p.num_taps * p.c_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.out);
    tma_prefetch_desc(&maps.in[0]);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<BLOCK_N>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)p.valid_rows * 128u + (uint32_t)L::kB;
      int it = 0;
      for (int t = 0; t < p.num_taps; ++t) {
        const Tap tap = p.taps[t];
        const CUtensorMap* in_map = &maps.in[tap.map];
        for (int cc = 0; cc < p.c_chunks; ++cc, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_arrive_expect_tx(full_bar(s), tx_bytes);
          const uint32_t a_dst = smem_base + s * L::kStage;
          tma_load_4d(in_map, full_bar(s), a_dst, cc * 64, tap.dw, a0 + tap.dh, n0);
          if (!B_MN) {
            tma_load_2d(&maps.w, full_bar(s), a_dst + L::kA, tap.widx * p.w_tap_stride + cc * 64, n_tile * BLOCK_N);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(&maps.w, full_bar(s), a_dst + L::kA + j * 8192, tap.widx * p.w_tap_stride + n_tile * BLOCK_N + j * 64, cc * 64);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 0, B_MN ? 1 : 0);
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * L::kStage;
        const uint32_t b_addr = a_addr + L::kA;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = umma_desc_sw128(a_addr + k * 32, 16, 1024);
          // K-major: +32 B per 16 K-elements inside the swizzle atom.  MN-major: 16 K-rows = 2 KB further down; the next
          // 64 N-elements are one 8 KB box away (LBO).
          const uint64_t db = B_MN ? umma_desc_sw128(b_addr + k * 2048, 8192, 1024) : umma_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_bf16(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));  // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;      // output pixel (tile row) owned by this thread
    const int et = threadIdx.x - 64;    // 0..127
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    constexpr int kChunks = BLOCK_N / 64;
#pragma unroll 1
    for (int ch = 0; ch < kChunks; ++ch) {
      const uint32_t buf = smem_base + L::kOffStaging + (ch & 1) * kBoxBytes;
      uint8_t* buf_gen = smem_gen + L::kOffStaging + (ch & 1) * kBoxBytes;
      if (ch >= 2) {  // the TMA store issued two chunks ago must have finished reading this buffer
        if (et == 0) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
      }
      uint32_t r0[32], r1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 64);
      tmem_ld_32x32(taddr, r0);
      tmem_ld_32x32(taddr + 32, r1);
      tmem_ld_wait();
      const uint32_t row_addr = buf + row * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t dst = row_addr + (((uint32_t)j ^ (uint32_t)(row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                     "r"(pack_bf16x2(__uint_as_float(r0[8 * j + 0]), __uint_as_float(r0[8 * j + 1]))),
                     "r"(pack_bf16x2(__uint_as_float(r0[8 * j + 2]), __uint_as_float(r0[8 * j + 3]))),
                     "r"(pack_bf16x2(__uint_as_float(r0[8 * j + 4]), __uint_as_float(r0[8 * j + 5]))),
                     "r"(pack_bf16x2(__uint_as_float(r0[8 * j + 6]), __uint_as_float(r0[8 * j + 7])))
                     : "memory");
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t dst = row_addr + (((uint32_t)(j + 4) ^ (uint32_t)(row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                     "r"(pack_bf16x2(__uint_as_float(r1[8 * j + 0]), __uint_as_float(r1[8 * j + 1]))),
                     "r"(pack_bf16x2(__uint_as_float(r1[8 * j + 2]), __uint_as_float(r1[8 * j + 3]))),
                     "r"(pack_bf16x2(__uint_as_float(r1[8 * j + 4]), __uint_as_float(r1[8 * j + 5]))),
                     "r"(pack_bf16x2(__uint_as_float(r1[8 * j + 6]), __uint_as_float(r1[8 * j + 7])))
                     : "memory");
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy) store
      named_bar_sync(1, 128);
      if (et == 0) {
        tma_store_4d(&maps.out, buf, n_tile * BLOCK_N + ch * 64, 0, a0, n0);
        tma_store_commit();
      }
      if (p.stats != nullptr) {
        // BatchNorm partials of the STORED (bf16-rounded) tile: thread = (channel, row half)
        const int c = et & 63;
        const int half = et >> 6;
        const int rbeg = half * 64;
        const int rend = min(p.valid_rows, rbeg + 64);
        float s = 0.f, ss = 0.f;
        const uint8_t* colp = buf_gen + (c & 7) * 2;
        for (int r = rbeg; r < rend; ++r) {
          const uint16_t raw = *reinterpret_cast<const uint16_t*>(colp + r * 128 + ((((uint32_t)c >> 3) ^ (uint32_t)(r & 7)) << 4));
          const float v = __uint_as_float((uint32_t)raw << 16);
          s += v;
          ss = fmaf(v, v, ss);
        }
        if (half == 1) stat_scratch[c] = make_float2(s, ss);
        named_bar_sync(1, 128);
        if (half == 0) {
          const float2 o = stat_scratch[c];
          // fp64 atomics: the summation order across CTAs then changes the result far below fp32 resolution
          stat_add(p.stats, p.cout, m_tile, n_tile * BLOCK_N + ch * 64 + c, s + o.x, ss + o.y);
        }
      }
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BLOCK_N>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------
// wgrad kernel:  dW[k][widx][c] (+)= sum over pixel tiles of  dY_tile^T (128 k x rows) * X_tile (rows x BLOCK_C c)
// Both operands are "MN-major": shared-memory rows are pixels (the GEMM K dimension), 64 channels per 128-B row.
// ------------------------------------------------------------------------------------------------------------------
struct alignas(64) WgradMaps {
  CUtensorMap in[kMaxViews];
  CUtensorMap dy;
};

struct WgradParams {
  int num_taps;
  int c_blocks;  // Cin / BLOCK_C
  int cin, cout;
  int rs;        // R*S (row pitch of dW in taps)
  int tiles_h, Hb, Nb, valid_rows;
  int m_tiles;   // total pixel tiles
  int splits;    // gridDim.y
  int ka;        // dY boxes per stage: min(cout,128)/64
  float* dw;
  Tap taps[kMaxTaps];
};

template <int BLOCK_C, int STAGES>
struct WgradSmem {
  static constexpr int kA = 2 * kBoxBytes;
  static constexpr int kB = (BLOCK_C / 64) * kBoxBytes;
  static constexpr int kStage = kA + kB;
  static constexpr int kOffBars = STAGES * kStage;
  static constexpr int kBytes = kOffBars + 1024 + 1024;
};

template <int BLOCK_C, int STAGES>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_kernel(const __grid_constant__ WgradMaps maps, const WgradParams p) {
  using L = WgradSmem<BLOCK_C, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t bars = smem_base + L::kOffBars;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kOffBars + 8 * (2 * STAGES + 1));

  // work item
  int w = blockIdx.x;
  const int cc = w % p.c_blocks;
  w /= p.c_blocks;
  const int ti = w % p.num_taps;
  const int kc = w / p.num_taps;
  const Tap tap = p.taps[ti];
  const int split = blockIdx.y;
  const int t_beg = (int)(((long long)p.m_tiles * split) / p.splits);
  const int t_end = (int)(((long long)p.m_tiles * (split + 1)) / p.splits);
  const int iters = t_end - t_beg;

  // rows [valid_rows, 16-aligned) of every box feed the MMA and are never written by TMA: zero all stages once
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* base = reinterpret_cast<uint4*>(smem_gen);
    for (int i = threadIdx.x; i < L::kOffBars / 16; i += blockDim.x) base[i] = z;
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.dy);
    tma_prefetch_desc(&maps.in[tap.map]);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<BLOCK_C>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)(p.ka + BLOCK_C / 64) * (uint32_t)p.valid_rows * 128u;
      for (int it = 0; it < iters; ++it) {
        const int m_tile = t_beg + it;
        const int n_blk = m_tile / p.tiles_h;
        const int h_blk = m_tile - n_blk * p.tiles_h;
        const int a0 = h_blk * p.Hb, n0 = n_blk * p.Nb;
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), tx_bytes);
        const uint32_t a_dst = smem_base + s * L::kStage;
        for (int j = 0; j < p.ka; ++j) tma_load_4d(&maps.dy, full_bar(s), a_dst + j * kBoxBytes, kc * 128 + j * 64, 0, a0, n0);
#pragma unroll
        for (int j = 0; j < BLOCK_C / 64; ++j)
          tma_load_4d(&maps.in[tap.map], full_bar(s), a_dst + L::kA + j * kBoxBytes, cc * BLOCK_C + j * 64, tap.dw, a0 + tap.dh, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_C, 1, 1);
      const int ksteps = (p.valid_rows + 15) >> 4;
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * L::kStage;
        const uint32_t b_addr = a_addr + L::kA;
        for (int ks = 0; ks < ksteps; ++ks) {
          // 16 pixel rows per MMA = two 8-row groups (SBO = 1024 B apart); next 64 channels are LBO = one box apart
          const uint64_t da = umma_desc_sw128(a_addr + ks * 2048, kBoxBytes, 1024);
          const uint64_t db = umma_desc_sw128(b_addr + ks * 2048, kBoxBytes, 1024);
          umma_bf16(tmem_base, da, db, idesc, (it | ks) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int k = kc * 128 + q * 32 + lane;  // output-channel row of dW owned by this thread
    if (iters > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      float* dst_row = p.dw + ((size_t)k * p.rs + tap.widx) * p.cin + cc * BLOCK_C;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_C; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        if (k < p.cout) {
          if (p.splits > 1) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + c0 + j), "f"(__uint_as_float(r[j])),
                           "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 cur = *reinterpret_cast<float4*>(dst_row + c0 + j);
              cur.x += __uint_as_float(r[j]);
              cur.y += __uint_as_float(r[j + 1]);
              cur.z += __uint_as_float(r[j + 2]);
              cur.w += __uint_as_float(r[j + 3]);
              *reinterpret_cast<float4*>(dst_row + c0 + j) = cur;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BLOCK_C>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------
// host side: tiling, tensor maps, tap tables
// ------------------------------------------------------------------------------------------------------------------
struct TileGeom {
  int Wb, Hb, Nb, valid_rows, tiles_h, tiles_n;
};

bool choose_tile(int outW, int outH, int N, TileGeom* t) {
  if (outW < 1 || outH < 1 || N < 1 || outW > 128) return false;
  t->Wb = outW;
  int hb = 1;
  for (int d = 1; d <= outH; ++d)
    if (outH % d == 0 && outW * d <= 128) hb = d;
  t->Hb = hb;
  t->Nb = 1;
  if (hb == outH) {
    int nb = 128 / (outW * outH);
    if (nb > N) nb = N;
    if (nb > 256) nb = 256;
    if (nb < 1) nb = 1;
    t->Nb = nb;
  }
  t->valid_rows = t->Wb * t->Hb * t->Nb;
  t->tiles_h = outH / t->Hb;
  t->tiles_n = (N + t->Nb - 1) / t->Nb;
  return true;
}

// 4-D view {C, Wv, Hv, N} over an NHWC bf16 tensor: rows h = ph + sh*i, cols w = pw + sw*j
struct View {
  const void* base;
  int C, Wv, Hv, N;
  long long strideW, strideH, strideN;  // bytes
};

View make_phase_view(const void* ptr, int N, int H, int W, int C, int st, int ph, int pw) {
  View v;
  v.base = (const char*)ptr + ((long long)ph * W + pw) * C * 2;
  v.C = C;
  v.Wv = (W - pw + st - 1) / st;
  v.Hv = (H - ph + st - 1) / st;
  v.N = N;
  v.strideW = (long long)st * C * 2;
  v.strideH = (long long)st * W * C * 2;
  v.strideN = (long long)H * W * C * 2;
  return v;
}

int encode_view(mml_ctx* ctx, CUtensorMap* map, const View& v, int boxW, int boxH, int boxN) {
  if (v.Wv < 1 || v.Hv < 1) return mml_set_error(ctx, MML_ERR_INVALID, "empty tensor view");
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.Wv, (cuuint64_t)v.Hv, (cuuint64_t)v.N};
  cuuint64_t strides[3] = {(cuuint64_t)v.strideW, (cuuint64_t)v.strideH, (cuuint64_t)v.strideN};
  cuuint32_t box[4] = {64, (cuuint32_t)boxW, (cuuint32_t)boxH, (cuuint32_t)boxN};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ctx->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return mml_set_error(ctx, MML_ERR_CUDA, "cuTensorMapEncodeTiled(4d) failed: %d (dims %d,%d,%d,%d box %d,%d,%d)", (int)r, v.C,
                         v.Wv, v.Hv, v.N, boxW, boxH, boxN);
  return MML_OK;
}

int encode_weights(mml_ctx* ctx, CUtensorMap* map, const void* w, long long inner, int rows, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)inner * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return mml_set_error(ctx, MML_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: %d", (int)r);
  return MML_OK;
}

inline bool tap_reaches(int out_extent, int d, int view_extent) { return d < view_extent && out_extent - 1 + d >= 0; }

template <typename K>
int set_smem_limit(mml_ctx* ctx, K kernel, int bytes) {
  MML_CHECK_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return MML_OK;
}

template <int BLOCK_N, int STAGES, int MIN_BLOCKS, bool B_MN>
int launch_igemm_t(mml_ctx* ctx, const IgemmMaps& maps, const IgemmParams& p, dim3 grid, cudaStream_t st) {
  using L = IgemmSmem<BLOCK_N, STAGES>;
  static bool configured = false;
  if (!configured) {
    int rc = set_smem_limit(ctx, conv_igemm_kernel<BLOCK_N, STAGES, MIN_BLOCKS, B_MN>, L::kBytes);
    if (rc) return rc;
    configured = true;
  }
  conv_igemm_kernel<BLOCK_N, STAGES, MIN_BLOCKS, B_MN><<<grid, 192, L::kBytes, st>>>(maps, p);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

// One "shifted GEMM" launch: out view <- sum over taps of in views @ weights.
//   b_mn = false: w is [cout rows][n_wtaps*cin inner] bf16;   b_mn = true: w is [cin rows][n_wtaps*cout inner] bf16
int run_igemm(mml_ctx* ctx, const View* in_views, int n_views, const void* w, int n_wtaps, int cin, int cout, const View& out,
              const Tap* taps, int num_taps, double* stats, bool b_mn, cudaStream_t st) {
  MML_REQUIRE(ctx, cin % 64 == 0 && cout % 64 == 0, "conv: channel counts must be multiples of 64 (got C=%d K=%d)", cin, cout);
  MML_REQUIRE(ctx, num_taps >= 1 && num_taps <= kMaxTaps && n_views <= kMaxViews, "conv: bad tap table");
  TileGeom tg;
  MML_REQUIRE(ctx, choose_tile(out.Wv, out.Hv, out.N, &tg), "conv: output width %d not supported (max 128)", out.Wv);
  IgemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  for (int i = 0; i < n_views; ++i)
    if ((rc = encode_view(ctx, &maps.in[i], in_views[i], tg.Wb, tg.Hb, tg.Nb))) return rc;
  const int block_n = cout >= 256 ? 256 : cout;
  if (!b_mn) {
    if ((rc = encode_weights(ctx, &maps.w, w, (long long)n_wtaps * cin, cout, block_n))) return rc;
  } else {
    if ((rc = encode_weights(ctx, &maps.w, w, (long long)n_wtaps * cout, cin, 64))) return rc;
  }
  if ((rc = encode_view(ctx, &maps.out, out, tg.Wb, tg.Hb, tg.Nb))) return rc;
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.num_taps = num_taps;
  p.c_chunks = cin / 64;
  p.w_tap_stride = b_mn ? cout : cin;
  p.cin = cin;
  p.tiles_h = tg.tiles_h;
  p.Hb = tg.Hb;
  p.Nb = tg.Nb;
  p.valid_rows = tg.valid_rows;
  p.cout = cout;
  p.stats = stats;
  for (int i = 0; i < num_taps; ++i) p.taps[i] = taps[i];
  dim3 grid(tg.tiles_h * tg.tiles_n, cout / block_n);
  if (!b_mn) {
    switch (block_n) {
      case 64: return launch_igemm_t<64, 3, 2, false>(ctx, maps, p, grid, st);
      case 128: return launch_igemm_t<128, 4, 1, false>(ctx, maps, p, grid, st);
      case 256: return launch_igemm_t<256, 3, 1, false>(ctx, maps, p, grid, st);
    }
  } else {
    switch (block_n) {
      case 64: return launch_igemm_t<64, 3, 2, true>(ctx, maps, p, grid, st);
      case 128: return launch_igemm_t<128, 4, 1, true>(ctx, maps, p, grid, st);
      case 256: return launch_igemm_t<256, 3, 1, true>(ctx, maps, p, grid, st);
    }
  }
  return mml_set_error(ctx, MML_ERR_INVALID, "conv: unsupported K=%d", cout);
}

int check_geom(mml_ctx* ctx, const mml_conv_geom* g, int* P, int* Q) {
  MML_REQUIRE(ctx, ctx && g, "conv: null ctx/geom");
  MML_REQUIRE(ctx, g->N >= 1 && g->H >= 1 && g->W >= 1, "conv: bad input dims");
  MML_REQUIRE(ctx, (g->R == 3 && g->S == 3) || (g->R == 1 && g->S == 1), "conv: only 3x3 and 1x1 filters (got %dx%d)", g->R, g->S);
  MML_REQUIRE(ctx, g->stride == 1 || g->stride == 2, "conv: stride must be 1 or 2");
  MML_REQUIRE(ctx, g->pad >= 0 && g->pad < g->R, "conv: bad padding");
  *P = (g->H + 2 * g->pad - g->R) / g->stride + 1;
  *Q = (g->W + 2 * g->pad - g->S) / g->stride + 1;
  MML_REQUIRE(ctx, *P >= 1 && *Q >= 1, "conv: empty output");
  return MML_OK;
}

inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// fprop-style tap table over phase views of the input
int build_fprop_taps(const mml_conv_geom* g, int P, int Q, const void* x, View* views, int* n_views, Tap* taps) {
  const int st = g->stride;
  *n_views = st * st;
  // a view that holds no element (e.g. odd phase of a 1-pixel-wide tensor) is never referenced by a valid tap
  for (int ph = 0; ph < st; ++ph)
    for (int pw = 0; pw < st; ++pw) views[ph * st + pw] = make_phase_view(x, g->N, g->H, g->W, g->C, st, ph, pw);
  int n = 0;
  for (int r = 0; r < g->R; ++r)
    for (int s = 0; s < g->S; ++s) {
      const int th = r - g->pad, tw = s - g->pad;
      const int ph = ((th % st) + st) % st, pw = ((tw % st) + st) % st;
      const int dh = floor_div(th - ph, st), dw = floor_div(tw - pw, st);
      const View& v = views[ph * st + pw];
      if (v.Hv < 1 || v.Wv < 1) continue;
      if (!tap_reaches(P, dh, v.Hv) || !tap_reaches(Q, dw, v.Wv)) continue;
      taps[n].map = (int8_t)(ph * st + pw);
      taps[n].dh = (int8_t)dh;
      taps[n].dw = (int8_t)dw;
      taps[n].widx = (int8_t)(r * g->S + s);
      ++n;
    }
  return n;
}

template <int BLOCK_C, int STAGES>
int launch_wgrad_t(mml_ctx* ctx, const WgradMaps& maps, const WgradParams& p, dim3 grid, cudaStream_t st) {
  using L = WgradSmem<BLOCK_C, STAGES>;
  static bool configured = false;
  if (!configured) {
    int rc = set_smem_limit(ctx, conv_wgrad_kernel<BLOCK_C, STAGES>, L::kBytes);
    if (rc) return rc;
    configured = true;
  }
  conv_wgrad_kernel<BLOCK_C, STAGES><<<grid, 192, L::kBytes, st>>>(maps, p);
  MML_LAUNCHED(ctx);
  return MML_OK;
}

}  // namespace

extern "C" {

int mml_conv_fprop(mml_ctx* ctx, const mml_conv_geom* g, const uint16_t* x, const uint16_t* w_krsc, uint16_t* y,
                   double* stats, void* stream) {
  int P, Q, rc;
  if ((rc = check_geom(ctx, g, &P, &Q))) return rc;
  View views[kMaxViews];
  Tap taps[kMaxTaps];
  int n_views = 0;
  const int n_taps = build_fprop_taps(g, P, Q, x, views, &n_views, taps);
  MML_REQUIRE(ctx, n_taps >= 1, "conv fprop: no filter tap reaches the input");
  // compact away empty views so that every encoded map is valid
  View used[kMaxViews];
  int remap[kMaxViews], n_used = 0;
  for (int i = 0; i < n_views; ++i) {
    remap[i] = -1;
    if (views[i].Hv >= 1 && views[i].Wv >= 1) {
      remap[i] = n_used;
      used[n_used++] = views[i];
    }
  }
  for (int i = 0; i < n_taps; ++i) taps[i].map = (int8_t)remap[taps[i].map];
  View out = make_phase_view(y, g->N, P, Q, g->K, 1, 0, 0);
  return run_igemm(ctx, used, n_used, w_krsc, g->R * g->S, g->C, g->K, out, taps, n_taps, stats, false, (cudaStream_t)stream);
}

int mml_conv_dgrad(mml_ctx* ctx, const mml_conv_geom* g, const uint16_t* dy, const uint16_t* w_krsc, uint16_t* dx, void* stream) {
  int P, Q, rc;
  if ((rc = check_geom(ctx, g, &P, &Q))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int s2 = g->stride;
  View in = make_phase_view(dy, g->N, P, Q, g->K, 1, 0, 0);
  bool need_zero = false;
  // dx[h] = sum_r dy[(h + pad - r)/stride] * w_t[r]  over taps with (h + pad - r) % stride == 0.
  // Per output phase e = h % stride:  h = stride*a + e,  p = a + (e + pad - r)/stride.
  struct Launch {
    View out;
    Tap taps[kMaxTaps];
    int n;
  } launches[4];
  int n_launch = 0;
  for (int eh = 0; eh < s2; ++eh)
    for (int ew = 0; ew < s2; ++ew) {
      View out = make_phase_view(dx, g->N, g->H, g->W, g->C, s2, eh, ew);
      if (out.Hv < 1 || out.Wv < 1) continue;
      Launch& L = launches[n_launch];
      L.out = out;
      L.n = 0;
      for (int r = 0; r < g->R; ++r)
        for (int s = 0; s < g->S; ++s) {
          const int th = eh + g->pad - r, tw = ew + g->pad - s;
          if (((th % s2) + s2) % s2 != 0 || ((tw % s2) + s2) % s2 != 0) continue;
          const int dh = floor_div(th, s2), dw = floor_div(tw, s2);
          if (!tap_reaches(out.Hv, dh, P) || !tap_reaches(out.Wv, dw, Q)) continue;
          Tap& t = L.taps[L.n++];
          t.map = 0;
          t.dh = (int8_t)dh;
          t.dw = (int8_t)dw;
          t.widx = (int8_t)(r * g->S + s);
        }
      if (L.n == 0)
        need_zero = true;
      else
        ++n_launch;
    }
  if (need_zero) MML_CHECK_CUDA(ctx, cudaMemsetAsync(dx, 0, (size_t)g->N * g->H * g->W * g->C * 2, st));
  for (int i = 0; i < n_launch; ++i) {
    // GEMM-K = k (rows of the K,R,S,C weight matrix), GEMM-N = c: the fprop weights are read as an MN-major B operand
    rc = run_igemm(ctx, &in, 1, w_krsc, g->R * g->S, g->K, g->C, launches[i].out, launches[i].taps, launches[i].n, nullptr, true, st);
    if (rc) return rc;
  }
  return MML_OK;
}

int mml_conv_wgrad(mml_ctx* ctx, const mml_conv_geom* g, const uint16_t* x, const uint16_t* dy, float* dw_krsc, void* stream) {
  int P, Q, rc;
  if ((rc = check_geom(ctx, g, &P, &Q))) return rc;
  MML_REQUIRE(ctx, g->C % 64 == 0 && g->K % 64 == 0, "conv wgrad: channel counts must be multiples of 64");
  View views[kMaxViews];
  Tap taps[kMaxTaps];
  int n_views = 0;
  const int n_taps = build_fprop_taps(g, P, Q, x, views, &n_views, taps);
  MML_REQUIRE(ctx, n_taps >= 1, "conv wgrad: no filter tap reaches the input");
  TileGeom tg;
  MML_REQUIRE(ctx, choose_tile(Q, P, g->N, &tg), "conv wgrad: output width %d not supported", Q);
  WgradMaps maps;
  memset(&maps, 0, sizeof(maps));
  int remap[kMaxViews], n_used = 0;
  for (int i = 0; i < n_views; ++i) {
    remap[i] = -1;
    if (views[i].Hv >= 1 && views[i].Wv >= 1) {
      if ((rc = encode_view(ctx, &maps.in[n_used], views[i], tg.Wb, tg.Hb, tg.Nb))) return rc;
      remap[i] = n_used++;
    }
  }
  View dyv = make_phase_view(dy, g->N, P, Q, g->K, 1, 0, 0);
  if ((rc = encode_view(ctx, &maps.dy, dyv, tg.Wb, tg.Hb, tg.Nb))) return rc;
  const int block_c = g->C >= 256 ? 256 : g->C;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.num_taps = n_taps;
  p.c_blocks = g->C / block_c;
  p.cin = g->C;
  p.cout = g->K;
  p.rs = g->R * g->S;
  p.tiles_h = tg.tiles_h;
  p.Hb = tg.Hb;
  p.Nb = tg.Nb;
  p.valid_rows = tg.valid_rows;
  p.m_tiles = tg.tiles_h * tg.tiles_n;
  p.ka = g->K >= 128 ? 2 : 1;
  p.dw = dw_krsc;
  for (int i = 0; i < n_taps; ++i) {
    p.taps[i] = taps[i];
    p.taps[i].map = (int8_t)remap[taps[i].map];
  }
  const int out_tiles = (int)mml_ceil_div(g->K, 128) * n_taps * p.c_blocks;
  int splits = (2 * ctx->sm_count + out_tiles - 1) / out_tiles;
  if (splits > p.m_tiles) splits = p.m_tiles;
  if (splits < 1) splits = 1;
  p.splits = splits;
  dim3 grid(out_tiles, splits);
  cudaStream_t st = (cudaStream_t)stream;
  switch (block_c) {
    case 64: return launch_wgrad_t<64, 4>(ctx, maps, p, grid, st);
    case 128: return launch_wgrad_t<128, 3>(ctx, maps, p, grid, st);
    case 256: return launch_wgrad_t<256, 2>(ctx, maps, p, grid, st);
  }
  return mml_set_error(ctx, MML_ERR_INVALID, "conv wgrad: unsupported C=%d", g->C);
}

}  // extern "C"
