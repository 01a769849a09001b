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
inline int persistent_sms(const mml_ctx* ctx) { return ctx->sm_budget > 0 ? ctx->sm_budget : ctx->sm_count; }


struct Tap {
  int8_t map;   // which input view
  int8_t dh;    // row offset in that view
  int8_t dw;    // column offset
  int8_t widx;  // filter tap index r*S+s in the weight matrix
};

constexpr int kMaxTaps = 9;
constexpr int kMaxViews = 4;
constexpr int kBoxBytes = 128 * 128;  // one 128-row x 64-channel bf16 box

constexpr int kMaxPhases = 4;  // output phases one launch may cover (stride-2 dgrad: dx[h % 2][w % 2])

struct alignas(64) IgemmMaps {
  CUtensorMap in[kMaxViews];
  CUtensorMap w;
  CUtensorMap out;
  CUtensorMap out_extra[kMaxPhases - 1];  // output views of phases 1..3 (blockIdx.z)
};

// geometry + tap table of an additional output phase (phase 0 lives in the IgemmParams fields themselves)
struct PhaseExtra {
  int num_taps, tiles_h, Hb, Nb, valid_rows, m_tiles;
  Tap taps[kMaxTaps];
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
  double* stats;  // [stat_slots(cout)][cout][2] (sum, sum of squares) of the stored output, fp64 atomics; or nullptr
  // split-K variant only: the output view as plain addresses (its epilogue stores rows directly instead of through TMA)
  uint8_t* out_base;
  long long out_sw, out_sh, out_sn;  // byte strides of the output view along W, H, N
  int out_w, out_n;                  // view extents (Wb == out_w)
  Tap taps[kMaxTaps];
  // multi-phase launches (grid.z = phases): pixel tiles of phase 0 (CTAs with blockIdx.x beyond their phase's count exit at once)
  int m_tiles;
  PhaseExtra ex[kMaxPhases - 1];
};

template <int BLOCK_N, int STAGES>
struct IgemmSmem {
  static constexpr int kA = kBoxBytes;          // 128 rows x 128 B
  static constexpr int kB = BLOCK_N * 128;      // BLOCK_N rows x 128 B
  static constexpr int kStage = kA + kB;
  static constexpr int kStaging = 2 * kBoxBytes;
  static constexpr int kOffStaging = STAGES * kStage;
  static constexpr int kOffBars = kOffStaging + kStaging;
  static constexpr int kBytes = kOffBars + 256 /*barriers, tmem slot*/ + 2048 /*stats scratch*/ + 1024 /*alignment slack*/;
};

// ------------------------------------------------------------------------------------------------------------------
// fprop / dgrad kernel
// ------------------------------------------------------------------------------------------------------------------
// B_MN = false: weights are [N rows][taps*K inner] (fprop: W[k][r][s][c]), B operand K-major.
// B_MN = true : weights are [K rows][taps*N inner] (dgrad reads the SAME K,R,S,C tensor: GEMM-K = k, GEMM-N = c), B operand
//               MN-major: 64-row x 64-element boxes, 8 KB each, one per 64 output channels.
template <int BLOCK_N, int STAGES, int MIN_BLOCKS, bool B_MN>
__global__ void __launch_bounds__(192, MIN_BLOCKS)
conv_igemm_kernel(const __grid_constant__ IgemmMaps maps, const __grid_constant__ IgemmParams p) {
  using L = IgemmSmem<BLOCK_N, STAGES>;
  // blockIdx.z = output phase: the four phase GEMMs of a stride-2 dgrad (different tap subsets, output views and -- for odd
  // tensors -- tile shapes) run as ONE grid instead of four part-filled ones
  int num_taps = p.num_taps, tiles_h = p.tiles_h, Hb = p.Hb, Nb = p.Nb, valid_rows = p.valid_rows, m_tiles = p.m_tiles;
  const Tap* taps = p.taps;
  const CUtensorMap* out_map = &maps.out;
  if (blockIdx.z > 0) {
    const PhaseExtra& e = p.ex[blockIdx.z - 1];
    num_taps = e.num_taps, tiles_h = e.tiles_h, Hb = e.Hb, Nb = e.Nb, valid_rows = e.valid_rows, m_tiles = e.m_tiles;
    taps = e.taps;
    out_map = &maps.out_extra[blockIdx.z - 1];
  }
  if ((int)blockIdx.x >= m_tiles) return;  // whole CTA, before any barrier / TMEM allocation
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
  float4* stat_scratch = reinterpret_cast<float4*>(smem_gen + L::kOffBars + 256);  // 128 x float4 (2 KB)

  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int n_blk = m_tile / tiles_h;
  const int h_blk = m_tile - n_blk * tiles_h;
  const int a0 = h_blk * Hb;
  const int n0 = n_blk * Nb;
  const int iters = num_taps * p.c_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(out_map);
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
  pdl_sync();  // barriers, TMEM and tensor-map prefetch are set up; everything below touches memory earlier kernels produced

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const uint32_t tx_bytes = (uint32_t)valid_rows * 128u + (uint32_t)L::kB;
      int it = 0;
      for (int t = 0; t < num_taps; ++t) {
        const Tap tap = taps[t];
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
    // elect.sync + (lo, hi) descriptor words with constant hi: ~5 uniform instructions per MMA (the issue rate of this one
    // thread is what bounds tiles whose MMAs are short)
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 0, B_MN ? 1 : 0);
      constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
      constexpr uint32_t kLoB = B_MN ? ((8192u >> 4) << 16) : (1u << 16);
      constexpr uint32_t kBStep = B_MN ? (2048u >> 4) : 2u;
      const uint32_t a_lo_base = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo_base = (((smem_base + L::kA) & 0x3FFFFu) >> 4) | kLoB;
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_lo = a_lo_base + (uint32_t)s * (L::kStage >> 4);
        const uint32_t b_lo = b_lo_base + (uint32_t)s * (L::kStage >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = ((uint64_t)kHi << 32) | (uint64_t)(a_lo + 2u * k);
          const uint64_t db = ((uint64_t)kHi << 32) | (uint64_t)(b_lo + kBStep * k);
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
        tma_store_4d(out_map, buf, n_tile * BLOCK_N + ch * 64, 0, a0, n0);
        tma_store_commit();
      }
      if (p.stats != nullptr) {
        // BatchNorm partials of the STORED (bf16-rounded) tile: thread = (channel pair wc, row quarter rq), one 32-bit shared load
        // per row, 8 loads in flight (round 1: one dependent 16-bit load per row and channel, ~1 us per tile)
        const int wc = et & 31, rq = et >> 5;
        const int rend = min(valid_rows - rq * 32, 32);
        const uint8_t* colp = buf_gen + (wc & 3) * 4;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
        for (int b = 0; b < rend; ++b) {
          const int r = rq * 32 + b;
          const uint32_t v = *reinterpret_cast<const uint32_t*>(colp + r * 128 + ((((uint32_t)wc >> 2) ^ (uint32_t)(r & 7)) << 4));
          const float a = bf16_lo(v), c = bf16_hi(v);
          s0 += a, s1 += c;
          q0 = fmaf(a, a, q0), q1 = fmaf(c, c, q1);
        }
        stat_scratch[et] = make_float4(s0, s1, q0, q1);
        named_bar_sync(1, 128);
        if (et < 32) {
          const float4 a = stat_scratch[et], b2 = stat_scratch[et + 32], c2 = stat_scratch[et + 64], d2 = stat_scratch[et + 96];
          const int c0 = n_tile * BLOCK_N + ch * 64 + 2 * et;
          // fp64 atomics: the summation order across CTAs then changes the result far below fp32 resolution
          stat_add(p.stats, p.cout, m_tile, c0, a.x + b2.x + c2.x + d2.x, a.z + b2.z + c2.z + d2.z);
          stat_add(p.stats, p.cout, m_tile, c0 + 1, a.y + b2.y + c2.y + d2.y, a.w + b2.w + c2.w + d2.w);
        }
        if (ch + 1 < kChunks) named_bar_sync(1, 128);  // the scratch is rewritten by the next chunk
      }
    }
    if (et == 0) tma_store_wait_read<0>();  // the source tile must outlive the copy; the global writes complete with the kernel
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BLOCK_N>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------
// split-K variant for grids far below one wave (the ResNet34 image encoder on 4x4 / 2x2 / 1x1 maps: 8..32 pixel tiles).
// A tap-shifted GEMM CTA streams (128 + BLOCK_N) x 128 bytes per 64-deep K step from L2, and ONE SM sustains only ~45 B/clk of
// that (measured: 16 CTAs x 1.2 MB each = 15 us for 1.2 GFLOP).  Here the (tap, channel-chunk) iterations of an output tile are
// split over the S CTAs of a thread-block cluster; every CTA accumulates its share in TMEM, parks the fp32 partial tile in its own
// shared memory, and after a cluster barrier each CTA sums a 128/S-row slice of the tile over all S partials through distributed
// shared memory IN A FIXED ORDER (deterministic, no atomics, no workspace), rounds to bf16, stores the rows and adds the BatchNorm
// partials of what it stored.  grid = (S, pixel tiles, channel tiles), cluster = (S, 1, 1).
// ------------------------------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES, bool B_MN>
__global__ void __launch_bounds__(192, 1)
conv_igemm_splitk_kernel(const __grid_constant__ IgemmMaps maps, const IgemmParams p) {
  using L = IgemmSmem<BLOCK_N, STAGES>;
  static_assert(STAGES * L::kStage >= 128 * BLOCK_N * 4 + 4096, "the fp32 partial tile + statistics scratch reuse the operand stages");
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

  const int S = (int)gridDim.x;  // == cluster size
  const int rank = (int)cluster_ctarank();
  const int m_tile = blockIdx.y;
  const int n_tile = blockIdx.z;
  const int n_blk = m_tile / p.tiles_h;
  const int h_blk = m_tile - n_blk * p.tiles_h;
  const int a0 = h_blk * p.Hb;
  const int n0 = n_blk * p.Nb;
  const int iters_total = p.num_taps * p.c_chunks;
  const int it_beg = (iters_total * rank) / S, it_end = (iters_total * (rank + 1)) / S;  // host guarantees S <= iters_total
  const int iters = it_end - it_beg;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w);
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
  pdl_sync();  // barriers, TMEM and tensor-map prefetch are set up; everything below touches memory earlier kernels produced

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx_bytes = (uint32_t)p.valid_rows * 128u + (uint32_t)L::kB;
      for (int i = 0; i < iters; ++i) {
        const int it = it_beg + i;
        const int t = it / p.c_chunks, cc = it - t * p.c_chunks;
        const Tap tap = p.taps[t];
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), tx_bytes);
        const uint32_t a_dst = smem_base + s * L::kStage;
        tma_load_4d(&maps.in[tap.map], full_bar(s), a_dst, cc * 64, tap.dw, a0 + tap.dh, n0);
        if (!B_MN) {
          tma_load_2d(&maps.w, full_bar(s), a_dst + L::kA, tap.widx * p.w_tap_stride + cc * 64, n_tile * BLOCK_N);
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_2d(&maps.w, full_bar(s), a_dst + L::kA + j * 8192, tap.widx * p.w_tap_stride + n_tile * BLOCK_N + j * 64, cc * 64);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 0, B_MN ? 1 : 0);
      constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
      constexpr uint32_t kLoB = B_MN ? ((8192u >> 4) << 16) : (1u << 16);
      constexpr uint32_t kBStep = B_MN ? (2048u >> 4) : 2u;
      const uint32_t a_lo_base = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo_base = (((smem_base + L::kA) & 0x3FFFFu) >> 4) | kLoB;
      for (int i = 0; i < iters; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_lo = a_lo_base + (uint32_t)s * (L::kStage >> 4);
        const uint32_t b_lo = b_lo_base + (uint32_t)s * (L::kStage >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = ((uint64_t)kHi << 32) | (uint64_t)(a_lo + 2u * k);
          const uint64_t db = ((uint64_t)kHi << 32) | (uint64_t)(b_lo + kBStep * k);
          umma_bf16(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ---- phase A: this CTA's fp32 partial tile TMEM -> own shared memory, [128 rows][BLOCK_N] with the 16-byte chunk index
    // XOR-swizzled by (row & 7) (conflict-free for the row-per-lane writes here and the chunk-per-lane reads of phase C).  All MMAs
    // of this CTA have completed when tmem_full fires, so the operand stages are free.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t prow = smem_base + (uint32_t)row * (uint32_t)(BLOCK_N * 4);
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t c4 = (uint32_t)(c0 >> 2) + (uint32_t)j;
        const uint32_t dst = prow + ((c4 ^ (uint32_t)(row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's partial tile is complete and visible cluster-wide

  if (warp >= 2) {
    // ---- phase C: rows [rank*R, (rank+1)*R) of the tile, summed over the S partials in rank order
    constexpr int CPR = BLOCK_N / 4;  // 16-byte chunks per row
    constexpr int G = 128 / CPR;      // row groups handled concurrently by the 128 threads
    const int et = threadIdx.x - 64;
    const int c4 = et % CPR, rg = et / CPR;
    const int R = 128 / S;
    const int hw = p.Hb * p.out_w;
    float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int rr = rg; rr < R; rr += G) {
      const int tr = rank * R + rr;  // tile row
      const uint32_t off = (uint32_t)tr * (uint32_t)(BLOCK_N * 4) + (((uint32_t)c4 ^ (uint32_t)(tr & 7)) << 4);
      // all S loads in flight before the first add (a distributed-shared-memory load takes ~200 cycles); summed in rank order
      float4 v[8];
#pragma unroll
      for (int sr = 0; sr < 8; ++sr)
        if (sr < S) v[sr] = ld_dsmem_f4(dsmem_addr(smem_base + off, (uint32_t)sr));
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int sr = 0; sr < 8; ++sr)
        if (sr < S) acc.x += v[sr].x, acc.y += v[sr].y, acc.z += v[sr].z, acc.w += v[sr].w;
      const uint32_t lo = pack_bf16x2(acc.x, acc.y), hi = pack_bf16x2(acc.z, acc.w);
      const int ni = tr / hw, rem = tr - ni * hw;
      const int h = rem / p.out_w, w = rem - h * p.out_w;
      if (tr < p.valid_rows && n0 + ni < p.out_n) {
        uint8_t* dst = p.out_base + (long long)(n0 + ni) * p.out_sn + (long long)(a0 + h) * p.out_sh + (long long)w * p.out_sw +
                       (long long)(n_tile * BLOCK_N + c4 * 4) * 2;
        *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
        // BatchNorm partials of the STORED values (rows beyond the tensor are never stored and contribute nothing)
        const float v0 = bf16_lo(lo), v1 = bf16_hi(lo), v2 = bf16_lo(hi), v3 = bf16_hi(hi);
        s4[0] += v0, s4[1] += v1, s4[2] += v2, s4[3] += v3;
        q4[0] = fmaf(v0, v0, q4[0]), q4[1] = fmaf(v1, v1, q4[1]), q4[2] = fmaf(v2, v2, q4[2]), q4[3] = fmaf(v3, v3, q4[3]);
      }
    }
    if (p.stats != nullptr) {
      // combine the G row groups through shared memory (scratch behind the partial tile), then one atomic pair per channel
      float4* sc = reinterpret_cast<float4*>(smem_gen + 128 * BLOCK_N * 4);  // [G][CPR][2] float4
      sc[(rg * CPR + c4) * 2] = make_float4(s4[0], s4[1], s4[2], s4[3]);
      sc[(rg * CPR + c4) * 2 + 1] = make_float4(q4[0], q4[1], q4[2], q4[3]);
      named_bar_sync(1, 128);
      if (rg == 0) {
        float4 a = sc[c4 * 2], b = sc[c4 * 2 + 1];
#pragma unroll
        for (int g2 = 1; g2 < G; ++g2) {
          const float4 a2 = sc[(g2 * CPR + c4) * 2], b2 = sc[(g2 * CPR + c4) * 2 + 1];
          a.x += a2.x, a.y += a2.y, a.z += a2.z, a.w += a2.w;
          b.x += b2.x, b.y += b2.y, b.z += b2.z, b.w += b2.w;
        }
        const int cb = n_tile * BLOCK_N + c4 * 4;
        const int slot = (int)blockIdx.y * S + rank;
        stat_add(p.stats, p.cout, slot, cb, a.x, b.x);
        stat_add(p.stats, p.cout, slot, cb + 1, a.y, b.y);
        stat_add(p.stats, p.cout, slot, cb + 2, a.z, b.z);
        stat_add(p.stats, p.cout, slot, cb + 3, a.w, b.w);
      }
    }
  }
  __syncthreads();
  cluster_sync_all();  // no CTA may exit (and release its shared memory) while a peer is still reading its partial tile
  if (warp == 1) tmem_dealloc<BLOCK_N>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------
// "halo" kernel: stride-1 3x3 convolutions on large feature maps (ResNet18 layer1 / layer2: 28x28x64, 14x14x128), fprop
// and dgrad.  The tap-shifted GEMM above re-reads every activation tile 9 times and every weight tile once per CTA, which
// makes those layers L2-bandwidth bound (172 B/cycle/SM needed, ~45 available).  Here a persistent CTA
//   * loads the (Hb+2) x (Wb+2) x 64 input patch of an output tile ONCE per 64-channel chunk (TMA box with the halo,
//     zero padding = out-of-bounds fill) and feeds all 9 taps from it: output pixel (h, w) is accumulator row
//     m = h*(Wb+2) + w, so tap (dh, dw) is simply the SAME shared-memory tile read (dh+1)*(Wb+2) + (dw+1) rows further
//     down -- a descriptor start-address shift; the 2 junk columns per row are clipped by the TMA store;
//   * processes T tiles per iteration against one weight stage (weights stay resident in shared memory when they fit);
//   * keeps 4 patches in flight (TMA latency) and double-buffers the TMEM accumulators, so the epilogue of iteration i
//     (8 warps in two groups) overlaps the MMAs of i+1.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kPatchRows = 192;
constexpr int kPatchBytes = kPatchRows * 128;  // 24 KB
constexpr int kHaloThreads = 320;              // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two groups of 4)

struct alignas(64) HaloMaps {
  CUtensorMap in;   // {C, W, H, N}, box {64, Wb+2, Hb+2, 1}
  CUtensorMap w;
  CUtensorMap out;  // {K, W, H, N}, box {64, Wb+2, Hb, 1}
};

struct HaloParams {
  int num_taps;
  int w_tap_stride;
  int tiles_h, Hb, Wb;
  int m_tiles, num_super;
  int cout;
  double* stats;
  Tap taps[kMaxTaps];
};

template <int CCH, int BLOCK_N, int T, bool W_RES>
struct HaloSmem {
  static constexpr int kWStage = BLOCK_N * 128;
  static constexpr int kWStages = W_RES ? 9 * CCH : 4;
  static constexpr int kNPS = 4 / T;  // patch stages (ring over the sequence of (iteration, chunk) pairs): 96 KB of patches
  static constexpr int kUnits = T * (BLOCK_N / 64);  // epilogue units (tile, 64-channel chunk) per iteration
  static constexpr int kOffW = kNPS * T * kPatchBytes;
  static constexpr int kOffStaging = kOffW + kWStages * kWStage;
  static constexpr int kOffBars = kOffStaging + 2 * kBoxBytes;
  static constexpr int kBytes = kOffBars + 512 /*barriers*/ + 4096 /*stats scratch*/ + 512 + 1024 /*alignment slack*/;
  static constexpr int kTmemCols = 2 * T * BLOCK_N;
  static_assert(kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM columns");
  static_assert(kBytes <= 232448, "shared memory budget");
};

template <int CCH, int BLOCK_N, int T, bool W_RES, bool B_MN>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_kernel(const __grid_constant__ HaloMaps maps, const HaloParams p) {
  using L = HaloSmem<CCH, BLOCK_N, T, W_RES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bars = smem_base + L::kOffBars;
  constexpr int NPS = L::kNPS;
  auto patch_full = [&](int s) { return bars + 8u * s; };
  auto patch_empty = [&](int s) { return bars + 8u * (4 + s); };
  auto acc_full = [&](int s) { return bars + 8u * (8 + s); };
  auto acc_empty = [&](int s) { return bars + 8u * (10 + s); };
  auto w_full = [&](int s) { return bars + 8u * (12 + s); };
  auto w_empty = [&](int s) { return bars + 8u * (12 + L::kWStages + s); };
  const uint32_t tmem_slot = bars + 8u * (12 + 2 * L::kWStages);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kOffBars + 8 * (12 + 2 * L::kWStages));
  float4* stat_scratch = reinterpret_cast<float4*>(smem_gen + L::kOffBars + 512);  // [2 groups][4 quarters... 32] float4

  auto patch_addr = [&](int ps, int j) { return smem_base + (uint32_t)((ps * T + j) * kPatchBytes); };
  const int n_my = (p.num_super - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int wp2 = p.Wb + 2;
  const int rows_m = p.Hb * wp2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.in);
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.out);
    for (int s = 0; s < NPS; ++s) {
      mbar_init(patch_full(s), 1);
      mbar_init(patch_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), L::kUnits == 1 ? 4 : 8);  // one unit per iteration: the two epilogue groups alternate iterations
    }
    for (int s = 0; s < L::kWStages; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<L::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_sync();  // barriers, TMEM and tensor-map prefetch are set up; everything below touches memory earlier kernels produced

  auto load_w_stage = [&](int s, int cc, const Tap& tap, uint32_t bar) {
    const uint32_t dst = smem_base + L::kOffW + s * L::kWStage;
    if (!B_MN) {
      tma_load_2d(&maps.w, bar, dst, tap.widx * p.w_tap_stride + cc * 64, 0);
    } else {
#pragma unroll
      for (int j = 0; j < BLOCK_N / 64; ++j) tma_load_2d(&maps.w, bar, dst + j * 8192, tap.widx * p.w_tap_stride + j * 64, cc * 64);
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      if (W_RES) {
        mbar_arrive_expect_tx(w_full(0), (uint32_t)(p.num_taps * CCH * L::kWStage));
        for (int cc = 0; cc < CCH; ++cc)
          for (int t = 0; t < p.num_taps; ++t) load_w_stage(cc * p.num_taps + t, cc, p.taps[t], w_full(0));
      }
      const uint32_t patch_tx = (uint32_t)((p.Hb + 2) * wp2) * 128u;
      int wi = 0;
      for (int it = 0; it < n_my; ++it) {
        const int st = (int)blockIdx.x + it * (int)gridDim.x;
        for (int cc = 0; cc < CCH; ++cc) {
          const int seq = it * CCH + cc, ps = seq % NPS;
          mbar_wait(patch_empty(ps), (((uint32_t)(seq / NPS)) & 1u) ^ 1u);
          mbar_arrive_expect_tx(patch_full(ps), patch_tx * T);
#pragma unroll
          for (int j = 0; j < T; ++j) {
            const int m_tile = st * T + j;  // beyond m_tiles -> image index out of range -> TMA zero fill, store clipped
            const int n = m_tile / p.tiles_h, hb = m_tile - n * p.tiles_h;
            tma_load_4d(&maps.in, patch_full(ps), patch_addr(ps, j), cc * 64, -1, hb * p.Hb - 1, n);
          }
          if (!W_RES) {
            for (int t = 0; t < p.num_taps; ++t, ++wi) {
              const int s = wi % L::kWStages;
              mbar_wait(w_empty(s), (((uint32_t)(wi / L::kWStages)) & 1u) ^ 1u);
              mbar_arrive_expect_tx(w_full(s), (uint32_t)L::kWStage);
              load_w_stage(s, cc, p.taps[t], w_full(s));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Measured on B200 (round 2): a 128 x N x 16 MMA with FRESH descriptors costs ~57 + 1.0 N cycles here (122 at N = 64, 187 at
    // N = 128) although it executes in N / 2 (issuing every MMA twice with the same descriptors added exactly 32 cycles each at
    // N = 64).  A second issuing warp (alternate tiles, own accumulator) changed nothing, an 8-row-aligned A window changed
    // nothing: the cost is not in this thread's instruction stream and not in the row shift, it is the operand fetch of a new
    // descriptor pair.  Only N = 256 tiles amortise it; the 64- and 128-channel layers cannot have them (N = output channels).
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 0, B_MN ? 1 : 0);
      constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
      constexpr uint32_t kLoA = 1u << 16;                               // K-major: LBO field = 1 (unused)
      constexpr uint32_t kLoB = B_MN ? ((8192u >> 4) << 16) : (1u << 16);
      constexpr uint32_t kBStep = B_MN ? (2048u >> 4) : 2u;             // per 16 K-elements
      uint32_t tap_off[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_off[t] = (uint32_t)(((p.taps[t].dh + 1) * wp2 + (p.taps[t].dw + 1)) * 8);  // rows * 128 B / 16
      if (W_RES) {
        mbar_wait(w_full(0), 0);
        tc_fence_after();
      }
      int wi = 0;
      for (int it = 0; it < n_my; ++it) {
        const int ab = it & 1;
        mbar_wait(acc_empty(ab), (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < CCH; ++cc) {
          const int seq = it * CCH + cc, ps = seq % NPS;
          mbar_wait(patch_full(ps), ((uint32_t)(seq / NPS)) & 1u);
          tc_fence_after();
          const uint32_t a_lo0 = ((patch_addr(ps, 0) & 0x3FFFFu) >> 4) | kLoA;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            int s;
            if (W_RES) {
              s = cc * 9 + t;
            } else {
              s = wi % L::kWStages;
              mbar_wait(w_full(s), ((uint32_t)(wi / L::kWStages)) & 1u);
              tc_fence_after();
            }
            const uint32_t b_lo = (((smem_base + L::kOffW + (uint32_t)s * L::kWStage) & 0x3FFFFu) >> 4) | kLoB;
#pragma unroll
            for (int j = 0; j < T; ++j) {
              const uint32_t a_lo = a_lo0 + (uint32_t)j * (kPatchBytes >> 4) + tap_off[t];
              const uint32_t d_tmem = tmem_base + (uint32_t)((ab * T + j) * BLOCK_N);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = ((uint64_t)kHi << 32) | (uint64_t)(a_lo + 2u * k);
                const uint64_t db = ((uint64_t)kHi << 32) | (uint64_t)(b_lo + kBStep * k);
                if (cc == 0 && t == 0 && k == 0) umma_bf16(d_tmem, da, db, idesc, 0u);
                else umma_bf16(d_tmem, da, db, idesc, 1u);
              }
            }
            if (!W_RES) {
              umma_commit(w_empty(s));
              ++wi;
            }
          }
          umma_commit(patch_empty(ps));
        }
        umma_commit(acc_full(ab));
      }
    }
  } else {
    // ===================== epilogue: two groups of 4 warps =====================
    const int gi = (warp - 2) >> 2;       // group
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;        // accumulator row == tile row m = h*(Wb+2) + w
    const int eg = ((warp - 2) & 3) * 32 + lane;  // 0..127 within the group
    const uint32_t bar_id = 1 + gi;
    const uint32_t stage = smem_base + L::kOffStaging + gi * kBoxBytes;
    const uint8_t* stage_gen = smem_gen + L::kOffStaging + gi * kBoxBytes;
    // statistics: thread = (word column wc: channels 2wc, 2wc+1) x (row quarter rq); valid rows of the quarter as a bit mask
    const int wc = eg & 31, rq = eg >> 5;
    uint32_t vmask = 0;
    for (int b = 0; b < 32; ++b) {
      const int r = rq * 32 + b;
      if (r < rows_m && (r % wp2) < p.Wb) vmask |= 1u << b;
    }
    constexpr int kChunks = BLOCK_N / 64;
    float4 st_acc0 = make_float4(0.f, 0.f, 0.f, 0.f), st_acc1 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < n_my; ++it) {
      const int ab = it & 1;
      if (L::kUnits == 1 && ab != gi) continue;  // single unit per iteration: group g owns accumulator buffer g
      const int st = (int)blockIdx.x + it * (int)gridDim.x;
      mbar_wait(acc_full(ab), ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int u = 0; u < L::kUnits; ++u) {
        if (L::kUnits > 1 && (u & 1) != gi) continue;  // units alternate between the two groups
        const int j = (T > 1) ? (u % T) : 0;
        const int ch = (T > 1) ? (u / T) : u;
        const int m_tile = st * T + j;
        const int n = m_tile / p.tiles_h, hb = m_tile - n * p.tiles_h;
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((ab * T + j) * BLOCK_N + ch * 64);
        tmem_ld_32x32(taddr, r0);
        tmem_ld_32x32(taddr + 32, r1);
        tmem_ld_wait();
        // The accumulator is in registers now: hand the TMEM buffer back BEFORE packing / storing / statistics, so the MMAs of
        // iteration it + 2 run under the rest of this epilogue (round 1 released it at the end of the iteration, and the MMA
        // warp spent ~500 cycles per tile waiting for it: profiles/r2_ncu_halo_l1_hot.txt).
        if (L::kUnits == 1 || u == L::kUnits - 2 + gi) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(ab));
        }
        if (eg == 0) tma_store_wait_read<0>();  // this group's previous TMA store has finished reading the staging buffer
        named_bar_sync(bar_id, 128);
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
        named_bar_sync(bar_id, 128);
        if (eg == 0 && m_tile < p.m_tiles) {
          tma_store_4d(&maps.out, stage, ch * 64, 0, hb * p.Hb, n);  // box {64, Wb+2, Hb, 1}: the 2 junk columns are clipped
          tma_store_commit();
        }
        if (p.stats != nullptr && m_tile < p.m_tiles) {
          // BatchNorm partials of the STORED (bf16-rounded) tile, accumulated in registers over all tiles of this CTA
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const uint8_t* colp = stage_gen + (wc & 3) * 4;
#pragma unroll 8
          for (int b = 0; b < 32; ++b) {
            if (vmask & (1u << b)) {
              const int r = rq * 32 + b;
              const uint32_t v = *reinterpret_cast<const uint32_t*>(colp + r * 128 + ((((uint32_t)wc >> 2) ^ (uint32_t)(r & 7)) << 4));
              const float a = bf16_lo(v), c = bf16_hi(v);
              s0 += a, s1 += c;
              q0 = fmaf(a, a, q0), q1 = fmaf(c, c, q1);
            }
          }
          if (kChunks == 1 || ch == 0) st_acc0.x += s0, st_acc0.y += s1, st_acc0.z += q0, st_acc0.w += q1;
          else st_acc1.x += s0, st_acc1.y += s1, st_acc1.z += q0, st_acc1.w += q1;
        }
      }
    }
    if (p.stats != nullptr) {  // one atomic per channel, group and CTA
      float4* sc = stat_scratch + gi * 128;
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch) {
        named_bar_sync(bar_id, 128);
        sc[eg] = ch == 0 ? st_acc0 : st_acc1;
        named_bar_sync(bar_id, 128);
        if (eg < 32) {
          const float4 a = sc[eg], b2 = sc[eg + 32], c2 = sc[eg + 64], d2 = sc[eg + 96];
          const int c0 = ch * 64 + 2 * eg;
          stat_add(p.stats, p.cout, 2 * (int)blockIdx.x + gi, c0, a.x + b2.x + c2.x + d2.x, a.z + b2.z + c2.z + d2.z);
          stat_add(p.stats, p.cout, 2 * (int)blockIdx.x + gi, c0 + 1, a.y + b2.y + c2.y + d2.y, a.w + b2.w + c2.w + d2.w);
        }
      }
    }
    if (eg == 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<L::kTmemCols>(tmem_base);
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
  float* dw;     // splits == 1: the gradient itself (overwritten); else partial slabs [splits][K][RS][C], reduced in a fixed order afterwards
  long long slab;  // elements per slab
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
  pdl_sync();  // barriers, TMEM and tensor-map prefetch are set up; everything below touches memory earlier kernels produced

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
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_C, 1, 1);
      constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
      constexpr uint32_t kLo = ((uint32_t)kBoxBytes >> 4) << 16;        // LBO: next 64 channels are one box further
      const uint32_t a_lo_base = ((smem_base & 0x3FFFFu) >> 4) | kLo;
      const uint32_t b_lo_base = (((smem_base + L::kA) & 0x3FFFFu) >> 4) | kLo;
      const int ksteps = (p.valid_rows + 15) >> 4;
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_lo = a_lo_base + (uint32_t)s * (L::kStage >> 4);
        const uint32_t b_lo = b_lo_base + (uint32_t)s * (L::kStage >> 4);
        // 16 pixel rows per MMA = two 8-row groups (SBO = 1024 B apart) = 2 KB further down per step
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (ks < ksteps) {
            const uint64_t da = ((uint64_t)kHi << 32) | (uint64_t)(a_lo + 128u * ks);
            const uint64_t db = ((uint64_t)kHi << 32) | (uint64_t)(b_lo + 128u * ks);
            umma_bf16(tmem_base, da, db, idesc, (it | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int k = kc * 128 + q * 32 + lane;  // output-channel row of dW owned by this thread
    // deterministic: every split stores its partial tile with plain stores (no atomics); a split that owns no pixel tile stores zeros
    if (iters > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    float* dst_row = p.dw + (size_t)split * (size_t)p.slab + ((size_t)k * p.rs + tap.widx) * p.cin + cc * BLOCK_C;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_C; c0 += 32) {
      uint32_t r[32];
      if (iters > 0) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (k < p.cout) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst_row + c0 + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BLOCK_C>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------
// "halo" wgrad: dW of the stride-1 3x3 layers on large feature maps.  Persistent CTAs walk pixel tiles; per tile they load
// the input patch (with halo) and the dy tile ONCE, both in the padded-width row order m = h*(Wb+2) + w (dy's two junk
// columns are out of bounds -> zeros), and feed all filter taps of the CTA from the same patch by row-shifting the
// descriptor, accumulating in TMEM over all tiles of the CTA:
//     D[(tap slot, c)][k] += sum_m  X[m + shift(tap)][c] * dY[m][k]          (A = patch window, B = dy, both MN-major)
//   C = K = 64 : one CTA owns all 9 taps; two taps are stacked in the M = 128 rows of one MMA (the second 64-row group is
//                simply the same patch a few rows further down: LBO = tap-to-tap row distance) -> 5 accumulators x 64 cols
//   C = K = 128: M = 128 = the two 64-channel chunks of one tap (LBO = patch stride), a CTA owns one filter row (3 taps)
//                -> 3 accumulators x 128 cols, grid.y = 3
// Each CTA finally writes its partial dW to a workspace; a second kernel sums the partials in a fixed order into dW.
// ------------------------------------------------------------------------------------------------------------------
struct alignas(64) WHaloMaps {
  CUtensorMap x;   // {C, W, H, N}, box {64, Wb+2, Hb+2, 1}
  CUtensorMap dy;  // {K, W, H, N}, box {64, Wb+2, Hb, 1}
};

struct WHaloParams {
  int tiles_h, Hb, Wb, m_tiles;
  int rows_m;       // Hb * (Wb + 2)
  float* ws;        // [gridDim.x][K][9][C] partials
};

template <int CH>  // channel count / 64 (C == K): 1 or 2
struct WHaloCfg {
  static constexpr int kC = CH * 64;
  static constexpr int kAcc = CH == 1 ? 5 : 3;                 // accumulators per CTA
  static constexpr int kStage = CH * kPatchBytes + CH * kBoxBytes;
  static constexpr int kStages = CH == 1 ? 4 : 2;
  static constexpr int kOffBars = kStages * kStage;
  static constexpr int kBytes = kOffBars + 1024 + 1024;
  static constexpr int kTmemCols = 512;
  static_assert(kAcc * kC <= 512, "TMEM");
};

template <int CH>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_halo_kernel(const __grid_constant__ WHaloMaps maps, const WHaloParams p) {
  using L = WHaloCfg<CH>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bars = smem_base + L::kOffBars;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (L::kStages + s); };
  const uint32_t done_bar = bars + 8u * (2 * L::kStages);
  const uint32_t tmem_slot = bars + 8u * (2 * L::kStages + 1);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kOffBars + 8 * (2 * L::kStages + 1));

  const int wp2 = p.Wb + 2;
  const int n_my = (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int trow = (int)blockIdx.y;  // CH == 2: the filter row owned by this CTA

  {  // rows that TMA never writes (patch rows past the halo box, dy rows >= rows_m) feed the MMA: zero everything once
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* base = reinterpret_cast<uint4*>(smem_gen);
    for (int i = threadIdx.x; i < L::kOffBars / 16; i += blockDim.x) base[i] = z;
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.dy);
    for (int s = 0; s < L::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<L::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_sync();  // barriers, TMEM and tensor-map prefetch are set up; everything below touches memory earlier kernels produced

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx = (uint32_t)CH * (uint32_t)((p.Hb + 2) * wp2 + p.rows_m) * 128u;
      for (int it = 0; it < n_my; ++it) {
        const int m_tile = (int)blockIdx.x + it * (int)gridDim.x;
        const int n = m_tile / p.tiles_h, hb = m_tile - n * p.tiles_h;
        const int s = it % L::kStages;
        mbar_wait(empty_bar(s), (((uint32_t)(it / L::kStages)) & 1u) ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), tx);
        const uint32_t dst = smem_base + s * L::kStage;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          tma_load_4d(&maps.x, full_bar(s), dst + c * kPatchBytes, c * 64, -1, hb * p.Hb - 1, n);
          tma_load_4d(&maps.dy, full_bar(s), dst + CH * kPatchBytes + c * kBoxBytes, c * 64, 0, hb * p.Hb, n);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, L::kC, 1, 1);
      constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
      // per accumulator: row shift of its (first) tap and the LBO between the two 64-row groups of the A operand
      uint32_t a_off[L::kAcc], a_lbo[L::kAcc];
#pragma unroll
      for (int a = 0; a < L::kAcc; ++a) {
        if (CH == 1) {
          const int t0 = 2 * a, t1 = (2 * a + 1 < 9) ? 2 * a + 1 : 2 * a;  // the 10th slot duplicates tap 8 (ignored)
          const int sh0 = (t0 / 3) * wp2 + (t0 % 3), sh1 = (t1 / 3) * wp2 + (t1 % 3);
          a_off[a] = (uint32_t)(sh0 * 8);
          a_lbo[a] = (uint32_t)(((sh1 - sh0) * 128) >> 4) << 16;
        } else {
          a_off[a] = (uint32_t)((trow * wp2 + a) * 8);
          a_lbo[a] = ((uint32_t)kPatchBytes >> 4) << 16;
        }
      }
      constexpr uint32_t kLoB = ((uint32_t)kBoxBytes >> 4) << 16;
      const int ksteps = (p.rows_m + 15) >> 4;
      for (int it = 0; it < n_my; ++it) {
        const int s = it % L::kStages;
        mbar_wait(full_bar(s), ((uint32_t)(it / L::kStages)) & 1u);
        tc_fence_after();
        const uint32_t x_lo = ((smem_base + (uint32_t)s * L::kStage) & 0x3FFFFu) >> 4;
        const uint32_t b_lo0 = (((smem_base + (uint32_t)s * L::kStage + CH * kPatchBytes) & 0x3FFFFu) >> 4) | kLoB;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (ks < ksteps) {
            const uint64_t db = ((uint64_t)kHi << 32) | (uint64_t)(b_lo0 + 128u * ks);  // 16 pixel rows = 2 KB per step
#pragma unroll
            for (int a = 0; a < L::kAcc; ++a) {
              const uint64_t da = ((uint64_t)kHi << 32) | (uint64_t)((x_lo + a_off[a] + 128u * ks) | a_lbo[a]);
              umma_bf16(tmem_base + (uint32_t)(a * L::kC), da, db, idesc, (it | ks) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(done_bar);
    }
  } else {
    // epilogue: accumulator row = (tap slot, c), column = k  ->  ws[cta][k][tap][c]
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* wsb = p.ws + (size_t)blockIdx.x * (size_t)(L::kC * 9 * L::kC);
    if (n_my > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int a = 0; a < L::kAcc; ++a) {
      int tap, c;
      if (CH == 1) {
        tap = 2 * a + (row >> 6);
        c = row & 63;
      } else {
        tap = trow * 3 + a;
        c = row;
      }
      const bool live = tap < 9;
#pragma unroll 1
      for (int k0 = 0; k0 < L::kC; k0 += 32) {
        uint32_t r[32];
        if (n_my > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * L::kC + k0), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (live) {
#pragma unroll
          for (int j = 0; j < 32; ++j) wsb[((size_t)(k0 + j) * 9 + tap) * L::kC + c] = __uint_as_float(r[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<L::kTmemCols>(tmem_base);
}

// dw[i] = sum over partials in a fixed order (deterministic; OVERWRITES dw).  Halo CH == 2: each grid.y slice wrote only its own filter row.
// block = 32 elements (float4) x 8 partial slices; slices are combined through shared memory in a fixed order
__global__ void __launch_bounds__(256) wgrad_partial_reduce_kernel(const float* __restrict__ ws, int parts, long long n, float* __restrict__ dw) {
  pdl_sync();
  __shared__ float4 sh[8][33];
  const long long n4 = n >> 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long i = blockIdx.x * 32ll + tx;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
    const float4* src = reinterpret_cast<const float4*>(ws) + i;
    int pidx = ty;
    for (; pidx + 24 < parts; pidx += 32) {  // 4 independent 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(src + (size_t)(pidx + 8 * u) * n4);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc.x += v[u].x, acc.y += v[u].y, acc.z += v[u].z, acc.w += v[u].w;
    }
    for (; pidx < parts; pidx += 8) {
      const float4 v = __ldg(src + (size_t)pidx * n4);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  }
  sh[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && i < n4) {
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      const float4 v = sh[y][tx];
      tot.x += v.x, tot.y += v.y, tot.z += v.z, tot.w += v.w;
    }
    reinterpret_cast<float4*>(dw)[i] = tot;
  }
}

// the same for a layer whose filter taps do not all reach the input: only the taps the wgrad kernel wrote are summed (grid.y = tap)
__global__ void __launch_bounds__(256) wgrad_tap_reduce_kernel(const float* __restrict__ ws, int parts, long long slab, float* __restrict__ dw, int K,
                                                               int RS, int C, const WgradParams p) {
  pdl_sync();
  const int widx = p.taps[blockIdx.y].widx;
  const long long i = blockIdx.x * 256ll + threadIdx.x;  // float4 index over [K][C]
  if (i >= (long long)K * C / 4) return;
  const int k = (int)(i / (C / 4)), c4 = (int)(i % (C / 4));
  const size_t off = ((size_t)k * RS + widx) * C + (size_t)c4 * 4;
  float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < parts; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(ws + (size_t)s * slab + off));
    tot.x += v.x, tot.y += v.y, tot.z += v.z, tot.w += v.w;
  }
  *reinterpret_cast<float4*>(dw + off) = tot;
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
  MML_LAUNCH(ctx, (conv_igemm_kernel<BLOCK_N, STAGES, MIN_BLOCKS, B_MN>), grid, 192, L::kBytes, st, maps, p);
  return MML_OK;
}

template <int BLOCK_N, int STAGES, bool B_MN>
int launch_igemm_splitk_t(mml_ctx* ctx, const IgemmMaps& maps, const IgemmParams& p, int S, int m_tiles, int n_tiles, cudaStream_t st) {
  using L = IgemmSmem<BLOCK_N, STAGES>;
  static bool configured = false;
  if (!configured) {
    int rc = set_smem_limit(ctx, conv_igemm_splitk_kernel<BLOCK_N, STAGES, B_MN>, L::kBytes);
    if (rc) return rc;
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(S, m_tiles, n_tiles);
  cfg.blockDim = dim3(192, 1, 1);
  cfg.dynamicSmemBytes = L::kBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->pdl ? 2 : 1;
  MML_CHECK_CUDA(ctx, cudaLaunchKernelEx(&cfg, conv_igemm_splitk_kernel<BLOCK_N, STAGES, B_MN>, maps, p));
  MML_LAUNCHED(ctx);
  return MML_OK;
}

int g_wgrad_min_tiles = 48;  // fewest pixel tiles (128 pixels each) per weight-gradient split (mml_debug_set key 4; swept on B200: DESIGN.md section 5)
int g_wgrad_narrow = 32;  // weight gradients of layers with at most this many pixel tiles use 128-wide output tiles (0 = off; key 5)
int g_igemm_narrow = 0;  // fprop / dgrad of layers with at most this many pixel tiles use 64-wide output tiles (0 = off; key 6)
int g_splitk_max = 1;  // largest cluster the split-K variant may use (mml_debug_set key 2; 1 = off, the default: see DESIGN.md)

// One "shifted GEMM" launch: out view <- sum over taps of in views @ weights, for 1..4 output phases (grid.z).
//   b_mn = false: w is [cout rows][n_wtaps*cin inner] bf16;   b_mn = true: w is [cin rows][n_wtaps*cout inner] bf16
// Several phases (stride-2 dgrad): every phase reads the ONE input view through its own tensor map (in[i], boxed for that phase's
// tile shape -- odd tensors have phases of different extents) and writes its own output view.
struct PhaseDesc {
  View out;
  const Tap* taps;
  int num_taps;
};

int run_igemm_phases(mml_ctx* ctx, const View* in_views, int n_views, const void* w, int n_wtaps, int cin, int cout, const PhaseDesc* ph,
                     int n_ph, double* stats, bool b_mn, cudaStream_t st) {
  MML_REQUIRE(ctx, cin % 64 == 0 && cout % 64 == 0, "conv: channel counts must be multiples of 64 (got C=%d K=%d)", cin, cout);
  MML_REQUIRE(ctx, n_ph >= 1 && n_ph <= kMaxPhases && n_views <= kMaxViews && (n_ph == 1 || (n_views == 1 && stats == nullptr)),
              "conv: bad phase table");
  TileGeom tgs[kMaxPhases];
  int total_tiles = 0, max_tiles = 0;
  for (int i = 0; i < n_ph; ++i) {
    MML_REQUIRE(ctx, ph[i].num_taps >= 1 && ph[i].num_taps <= kMaxTaps, "conv: bad tap table");
    MML_REQUIRE(ctx, choose_tile(ph[i].out.Wv, ph[i].out.Hv, ph[i].out.N, &tgs[i]), "conv: output width %d not supported (max 128)", ph[i].out.Wv);
    const int t = tgs[i].tiles_h * tgs[i].tiles_n;
    total_tiles += t;
    if (t > max_tiles) max_tiles = t;
  }
  const TileGeom& tg = tgs[0];
  const View& out = ph[0].out;
  const int num_taps = ph[0].num_taps;
  IgemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  if (n_ph == 1) {
    for (int i = 0; i < n_views; ++i)
      if ((rc = encode_view(ctx, &maps.in[i], in_views[i], tg.Wb, tg.Hb, tg.Nb))) return rc;
  } else {
    for (int i = 0; i < n_ph; ++i)
      if ((rc = encode_view(ctx, &maps.in[i], in_views[0], tgs[i].Wb, tgs[i].Hb, tgs[i].Nb))) return rc;
  }
  // 128x256 tiles have the best operand reuse, but a grid far below one wave (ResNet18 layer4: 32 pixel tiles) runs faster
  // with 128x128 tiles on twice as many SMs
  int block_n = cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64);
  if (block_n == 256 && total_tiles * (cout / 256) * 2 <= ctx->sm_count) block_n = 128;
  // a single pixel tile (the Linear layers of the MMIMDb step: M = batch <= 128): spread the output columns over as many SMs
  // as possible, every CTA streams the whole A operand from L2 anyway
  if (total_tiles == 1) block_n = 64;
  // a handful of pixel tiles (the ResNet34 4x4 / 2x2 / 1x1 maps): 64-wide tiles double the CTA count again (mml_debug_set key 6)
  if (g_igemm_narrow && block_n == 128 && total_tiles <= g_igemm_narrow && total_tiles * (cout / 64) * 2 <= ctx->sm_count) block_n = 64;
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
  p.m_tiles = tg.tiles_h * tg.tiles_n;
  p.cout = cout;
  p.stats = stats;
  for (int i = 0; i < num_taps; ++i) p.taps[i] = ph[0].taps[i];
  for (int i = 1; i < n_ph; ++i) {
    PhaseExtra& e = p.ex[i - 1];
    e.num_taps = ph[i].num_taps, e.tiles_h = tgs[i].tiles_h, e.Hb = tgs[i].Hb, e.Nb = tgs[i].Nb, e.valid_rows = tgs[i].valid_rows;
    e.m_tiles = tgs[i].tiles_h * tgs[i].tiles_n;
    for (int t = 0; t < ph[i].num_taps; ++t) {
      e.taps[t] = ph[i].taps[t];
      e.taps[t].map = (int8_t)i;  // this phase's boxing of the input view
    }
    if ((rc = encode_view(ctx, &maps.out_extra[i - 1], ph[i].out, tgs[i].Wb, tgs[i].Hb, tgs[i].Nb))) return rc;
  }
  // grids far below one wave: split the K loop over a cluster (see conv_igemm_splitk_kernel).  The cluster size is capped
  // (default 4) because a cluster needs that many free SMs in ONE GPC at the same time, which a concurrently running persistent
  // kernel of another stream (the audio encoder) makes unlikely for large clusters.
  if (n_ph == 1) {
    const int tiles = tg.tiles_h * tg.tiles_n * (cout / block_n);
    int S = 1;
    for (int c = 2; c <= g_splitk_max && c <= 8; c *= 2)
      if (tiles * c <= ctx->sm_count && c <= num_taps * p.c_chunks) S = c;
    if (S > 1 && tiles * 2 <= ctx->sm_count) {
      p.out_base = (uint8_t*)const_cast<void*>(out.base);
      p.out_sw = out.strideW, p.out_sh = out.strideH, p.out_sn = out.strideN;
      p.out_w = out.Wv, p.out_n = out.N;
      const int mt = tg.tiles_h * tg.tiles_n, nt = cout / block_n;
      if (!b_mn) {
        switch (block_n) {
          case 64: return launch_igemm_splitk_t<64, 3, false>(ctx, maps, p, S, mt, nt, st);
          case 128: return launch_igemm_splitk_t<128, 4, false>(ctx, maps, p, S, mt, nt, st);
          case 256: return launch_igemm_splitk_t<256, 3, false>(ctx, maps, p, S, mt, nt, st);
        }
      } else {
        switch (block_n) {
          case 64: return launch_igemm_splitk_t<64, 3, true>(ctx, maps, p, S, mt, nt, st);
          case 128: return launch_igemm_splitk_t<128, 4, true>(ctx, maps, p, S, mt, nt, st);
          case 256: return launch_igemm_splitk_t<256, 3, true>(ctx, maps, p, S, mt, nt, st);
        }
      }
    }
  }
  dim3 grid(max_tiles, cout / block_n, n_ph);
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
  return mml_set_error(ctx, MML_ERR_INVALID, "conv: unsupported tile width %d", block_n);
}

int run_igemm(mml_ctx* ctx, const View* in_views, int n_views, const void* w, int n_wtaps, int cin, int cout, const View& out,
              const Tap* taps, int num_taps, double* stats, bool b_mn, cudaStream_t st) {
  PhaseDesc ph;
  ph.out = out, ph.taps = taps, ph.num_taps = num_taps;
  return run_igemm_phases(ctx, in_views, n_views, w, n_wtaps, cin, cout, &ph, 1, stats, b_mn, st);
}

int check_geom(mml_ctx* ctx, const mml_conv_geom* g, int* P, int* Q) {
  MML_REQUIRE(ctx, ctx && g, "conv: null ctx/geom");
  MML_REQUIRE(ctx, g->N >= 1 && g->H >= 1 && g->W >= 1, "conv: bad input dims");
  // 3x3 / 1x1 (ResNet), k x 1 (TextCNN over time, W == 1): any filter of at most kMaxTaps taps is a tap-shifted GEMM
  MML_REQUIRE(ctx, g->R >= 1 && g->S >= 1 && g->R * g->S <= kMaxTaps, "conv: at most %d filter taps (got %dx%d)", kMaxTaps, g->R, g->S);
  MML_REQUIRE(ctx, g->stride == 1 || g->stride == 2, "conv: stride must be 1 or 2");
  MML_REQUIRE(ctx, g->pad >= 0 && g->pad < g->R && (g->pad == 0 || g->pad < g->S), "conv: bad padding");
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

// ---- halo kernel dispatch -------------------------------------------------------------------------------------------
struct HaloGeom {
  int Wb, Hb, tiles_h, m_tiles;
};

// eligible: 3x3 taps with |dh|,|dw| <= 1 on one full-resolution view, C == K in {64, 128}, and a tile of Hb x W pixels whose
// padded-width row index Hb*(W+2) fits the 128 accumulator rows while the deepest tap shift stays inside the 192-row patch
bool halo_geometry(int outW, int outH, int N, int cin, int cout, HaloGeom* hg) {
  if (cin != cout || (cin != 64 && cin != 128)) return false;
  if (outW + 2 > 31 || outW < 3 || outH < 3) return false;
  int hb = 0;
  for (int d = 1; d <= outH; ++d)
    if (outH % d == 0 && (outW + 2) * d <= 128) hb = d;
  if (hb == 0) return false;
  if ((outW + 2) * hb < 96) return false;  // too few live rows per 128-row MMA: the tap-shifted kernel packs better
  if (2 * (outW + 2) + 2 + 128 > kPatchRows) return false;
  hg->Wb = outW, hg->Hb = hb, hg->tiles_h = outH / hb, hg->m_tiles = (outH / hb) * N;
  return true;
}

template <int CCH, int BLOCK_N, int T, bool W_RES, bool B_MN>
int launch_halo_t(mml_ctx* ctx, const HaloMaps& maps, const HaloParams& p, cudaStream_t st) {
  using L = HaloSmem<CCH, BLOCK_N, T, W_RES>;
  static bool configured = false;
  if (!configured) {
    int rc = set_smem_limit(ctx, conv_halo_kernel<CCH, BLOCK_N, T, W_RES, B_MN>, L::kBytes);
    if (rc) return rc;
    configured = true;
  }
  const int sms = persistent_sms(ctx);
  int grid = p.num_super < sms ? p.num_super : sms;
  MML_LAUNCH(ctx, (conv_halo_kernel<CCH, BLOCK_N, T, W_RES, B_MN>), grid, kHaloThreads, L::kBytes, st, maps, p);
  return MML_OK;
}

int g_halo_enable = 1;

int run_halo(mml_ctx* ctx, const View& in, const void* w, int n_wtaps, int cin, int cout, const View& out, const Tap* taps, int num_taps,
             double* stats, bool b_mn, const HaloGeom& hg, cudaStream_t st) {
  HaloMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  if ((rc = encode_view(ctx, &maps.in, in, hg.Wb + 2, hg.Hb + 2, 1))) return rc;
  if ((rc = encode_view(ctx, &maps.out, out, hg.Wb + 2, hg.Hb, 1))) return rc;
  if (!b_mn) {
    if ((rc = encode_weights(ctx, &maps.w, w, (long long)n_wtaps * cin, cout, cout))) return rc;
  } else {
    if ((rc = encode_weights(ctx, &maps.w, w, (long long)n_wtaps * cout, cin, 64))) return rc;
  }
  const int T = cin == 64 ? 1 : 2;  // C = 64: weights resident, one tile per iteration, 4 patches in flight
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.num_taps = num_taps;
  p.w_tap_stride = b_mn ? cout : cin;
  p.tiles_h = hg.tiles_h, p.Hb = hg.Hb, p.Wb = hg.Wb;
  p.m_tiles = hg.m_tiles;
  p.num_super = (hg.m_tiles + T - 1) / T;
  p.cout = cout;
  p.stats = stats;
  for (int i = 0; i < num_taps; ++i) p.taps[i] = taps[i];
  if (cin == 64) return b_mn ? launch_halo_t<1, 64, 1, true, true>(ctx, maps, p, st) : launch_halo_t<1, 64, 1, true, false>(ctx, maps, p, st);
  return b_mn ? launch_halo_t<2, 128, 2, false, true>(ctx, maps, p, st) : launch_halo_t<2, 128, 2, false, false>(ctx, maps, p, st);
}

template <int CH>
int launch_wgrad_halo_t(mml_ctx* ctx, const WHaloMaps& maps, WHaloParams& p, int ctas, float* dw, cudaStream_t st) {
  using L = WHaloCfg<CH>;
  static bool configured = false;
  if (!configured) {
    int rc = set_smem_limit(ctx, conv_wgrad_halo_kernel<CH>, L::kBytes);
    if (rc) return rc;
    configured = true;
  }
  const long long n = (long long)L::kC * 9 * L::kC;
  dim3 grid(ctas, CH == 1 ? 1 : 3);
  MML_LAUNCH(ctx, conv_wgrad_halo_kernel<CH>, grid, 192, L::kBytes, st, maps, p);
  int rgrid = (int)mml_ceil_div(n / 4, 32);
  MML_LAUNCH(ctx, wgrad_partial_reduce_kernel, rgrid, 256, 0, st, p.ws, ctas, n, dw);
  return MML_OK;
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
  MML_LAUNCH(ctx, (conv_wgrad_kernel<BLOCK_C, STAGES>), grid, 192, L::kBytes, st, maps, p);
  return MML_OK;
}

}  // namespace

extern "C" {

/* experiment / A-B switches: key 1 = halo kernel enable (0/1), key 2 = largest split-K cluster (1 = off, 2, 4, 8) */
extern int mml_g_bn_one_wave;  // bn_act.cu

int mml_debug_set(int key, int value) {
  if (key == 1) g_halo_enable = value;
  else if (key == 2 && (value == 1 || value == 2 || value == 4 || value == 8)) g_splitk_max = value;
  else if (key == 3 && value >= 0 && value <= 2) mml_g_bn_one_wave = value;
  else if (key == 4 && value >= 1 && value <= 64) g_wgrad_min_tiles = value;
  else if (key == 5 && value >= 0 && value <= 1024) g_wgrad_narrow = value;
  else if (key == 6 && value >= 0 && value <= 1024) g_igemm_narrow = value;
  else return MML_ERR_INVALID;
  return MML_OK;
}

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
  HaloGeom hg;
  if (g_halo_enable && g->R == 3 && g->stride == 1 && g->pad == 1 && n_taps == 9 && halo_geometry(Q, P, g->N, g->C, g->K, &hg))
    return run_halo(ctx, used[0], w_krsc, 9, g->C, g->K, out, taps, n_taps, stats, false, hg, (cudaStream_t)stream);
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
  HaloGeom hg;
  if (g_halo_enable && g->R == 3 && s2 == 1 && g->pad == 1 && n_launch == 1 && launches[0].n == 9 && halo_geometry(g->W, g->H, g->N, g->K, g->C, &hg))
    return run_halo(ctx, in, w_krsc, 9, g->K, g->C, launches[0].out, launches[0].taps, 9, nullptr, true, hg, st);
  // GEMM-K = k (rows of the K,R,S,C weight matrix), GEMM-N = c: the fprop weights are read as an MN-major B operand.  The output
  // phases of a strided convolution (up to four part-filled grids in round 1) go out as ONE launch, grid.z = phase.
  PhaseDesc ph[kMaxPhases];
  for (int i = 0; i < n_launch; ++i) ph[i].out = launches[i].out, ph[i].taps = launches[i].taps, ph[i].num_taps = launches[i].n;
  if (n_launch == 0) return MML_OK;
  return run_igemm_phases(ctx, &in, 1, w_krsc, g->R * g->S, g->K, g->C, ph, n_launch, nullptr, true, st);
}

// how mml_conv_wgrad will run a geometry: kernel family, grid, splits and the workspace it needs for the per-CTA / per-split partials
struct WgradPlan {
  bool halo;
  HaloGeom hg;
  int halo_ctas;
  TileGeom tg;
  int n_taps, block_c, out_tiles, splits;
  long long slab;  // elements of one dW-shaped partial slab
  size_t ws_bytes;
};

static int plan_wgrad(const mml_ctx* ctx, const mml_conv_geom* g, int P, int Q, int n_taps, WgradPlan* wp) {
  memset(wp, 0, sizeof(*wp));
  wp->n_taps = n_taps;
  wp->slab = (long long)g->K * g->R * g->S * g->C;
  if (g_halo_enable && g->R == 3 && g->stride == 1 && g->pad == 1 && n_taps == 9 && halo_geometry(Q, P, g->N, g->C, g->K, &wp->hg)) {
    const int CH = g->C / 64;
    const int sms = persistent_sms(ctx);
    int ctas = CH == 1 ? sms : sms / 3;
    if (ctas > wp->hg.m_tiles) ctas = wp->hg.m_tiles;
    wp->halo = true;
    wp->halo_ctas = ctas;
    // sized for ALL SMs: the SM budget of the persistent kernels may differ between the query and the launch
    const int max_ctas = CH == 1 ? ctx->sm_count : ctx->sm_count / 3;
    wp->ws_bytes = (size_t)(max_ctas < wp->hg.m_tiles ? max_ctas : wp->hg.m_tiles) * wp->slab * sizeof(float);
    return MML_OK;
  }
  if (!choose_tile(Q, P, g->N, &wp->tg)) return MML_ERR_INVALID;
  wp->block_c = g->C % 256 == 0 ? 256 : (g->C % 128 == 0 ? 128 : 64);
  wp->out_tiles = (int)mml_ceil_div(g->K, 128) * n_taps * (g->C / wp->block_c);
  const int m_tiles = wp->tg.tiles_h * wp->tg.tiles_n;
  // few pixel tiles (the ResNet34 2x2 / 1x1 maps): 128-wide output tiles give twice the CTAs and a three-stage instead of a two-stage
  // operand pipeline (64 KB instead of 96 KB per stage) for a loop that is all load latency (mml_debug_set key 5)
  if (g_wgrad_narrow && wp->block_c == 256 && m_tiles <= g_wgrad_narrow && wp->out_tiles * 2 <= ctx->sm_count) {
    wp->block_c = 128;
    wp->out_tiles *= 2;
  }
  // one wave: every split's partial tile is written to and read back from the workspace, so more CTAs than SMs only add traffic
  int splits = ctx->sm_count / wp->out_tiles;
  // ... and every split should own a few pixel tiles: a split of ONE 128-pixel tile runs 8 MMAs and then writes (and the reduce launch
  // re-reads) a whole fp32 partial tile -- on the ResNet34 2x2 / 1x1 maps the partials were 8x the useful traffic (18.9 MB for a
  // 2.4 MB gradient), and a single split needs no reduce launch at all
  const int by_depth = (m_tiles + g_wgrad_min_tiles - 1) / g_wgrad_min_tiles;
  if (splits > by_depth) splits = by_depth;
  if (splits > m_tiles) splits = m_tiles;
  if (splits < 1) splits = 1;
  wp->splits = splits;
  wp->ws_bytes = splits > 1 ? (size_t)splits * wp->slab * sizeof(float) : 0;
  return MML_OK;
}

int64_t mml_conv_wgrad_workspace(const mml_ctx* ctx, const mml_conv_geom* g) {
  if (!ctx || !g || g->stride < 1) return -1;
  const int P = (g->H + 2 * g->pad - g->R) / g->stride + 1, Q = (g->W + 2 * g->pad - g->S) / g->stride + 1;
  if (P < 1 || Q < 1) return -1;
  View views[kMaxViews];
  Tap taps[kMaxTaps];
  int n_views = 0;
  const int n_taps = build_fprop_taps(g, P, Q, nullptr, views, &n_views, taps);
  WgradPlan wp;
  if (n_taps < 1 || plan_wgrad(ctx, g, P, Q, n_taps, &wp) != MML_OK) return -1;
  return (int64_t)wp.ws_bytes;
}

int mml_conv_wgrad(mml_ctx* ctx, const mml_conv_geom* g, const uint16_t* x, const uint16_t* dy, float* dw_krsc, float* workspace,
                   int64_t workspace_bytes, void* stream) {
  int P, Q, rc;
  if ((rc = check_geom(ctx, g, &P, &Q))) return rc;
  MML_REQUIRE(ctx, g->C % 64 == 0 && g->K % 64 == 0, "conv wgrad: channel counts must be multiples of 64");
  View views[kMaxViews];
  Tap taps[kMaxTaps];
  int n_views = 0;
  const int n_taps = build_fprop_taps(g, P, Q, x, views, &n_views, taps);
  MML_REQUIRE(ctx, n_taps >= 1, "conv wgrad: no filter tap reaches the input");
  WgradPlan wp;
  MML_REQUIRE(ctx, plan_wgrad(ctx, g, P, Q, n_taps, &wp) == MML_OK, "conv wgrad: output width %d not supported", Q);
  MML_REQUIRE(ctx, wp.ws_bytes == 0 || (workspace != nullptr && (size_t)workspace_bytes >= wp.ws_bytes),
              "conv wgrad: workspace of %lld bytes needed (mml_conv_wgrad_workspace), got %lld", (long long)wp.ws_bytes, (long long)workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  if (wp.halo) {
    const HaloGeom& hg = wp.hg;
    const int CH = g->C / 64;
    WHaloMaps hm;
    memset(&hm, 0, sizeof(hm));
    if ((rc = encode_view(ctx, &hm.x, views[0], hg.Wb + 2, hg.Hb + 2, 1))) return rc;
    View dyh = make_phase_view(dy, g->N, P, Q, g->K, 1, 0, 0);
    if ((rc = encode_view(ctx, &hm.dy, dyh, hg.Wb + 2, hg.Hb, 1))) return rc;
    WHaloParams hp;
    hp.tiles_h = hg.tiles_h, hp.Hb = hg.Hb, hp.Wb = hg.Wb, hp.m_tiles = hg.m_tiles;
    hp.rows_m = hg.Hb * (hg.Wb + 2);
    hp.ws = workspace;
    // CH == 2: the three filter-row CTAs of a column write disjoint taps of the same partial slot
    return CH == 1 ? launch_wgrad_halo_t<1>(ctx, hm, hp, wp.halo_ctas, dw_krsc, st) : launch_wgrad_halo_t<2>(ctx, hm, hp, wp.halo_ctas, dw_krsc, st);
  }
  const TileGeom& tg = wp.tg;
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
  const int block_c = wp.block_c;
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
  p.splits = wp.splits;
  p.slab = wp.slab;
  p.dw = wp.splits > 1 ? workspace : dw_krsc;
  for (int i = 0; i < n_taps; ++i) {
    p.taps[i] = taps[i];
    p.taps[i].map = (int8_t)remap[taps[i].map];
  }
  // filter taps that never reach the input (e.g. 8 of the 9 taps on a 1x1 map) are written by no CTA: their gradient is zero
  if (n_taps < g->R * g->S) MML_CHECK_CUDA(ctx, cudaMemsetAsync(dw_krsc, 0, (size_t)wp.slab * sizeof(float), st));
  dim3 grid(wp.out_tiles, wp.splits);
  switch (block_c) {
    case 64: rc = launch_wgrad_t<64, 4>(ctx, maps, p, grid, st); break;
    case 128: rc = launch_wgrad_t<128, 3>(ctx, maps, p, grid, st); break;
    case 256: rc = launch_wgrad_t<256, 2>(ctx, maps, p, grid, st); break;
    default: return mml_set_error(ctx, MML_ERR_INVALID, "conv wgrad: unsupported C=%d", g->C);
  }
  if (rc) return rc;
  if (wp.splits > 1) {
    // fixed-order sum of the split partials into dW; unreached taps hold garbage in the slabs, so only reached taps are summed
    // when some taps are missing (rare: tiny maps), otherwise the whole slab in one launch
    if (n_taps == g->R * g->S) {
      MML_LAUNCH(ctx, wgrad_partial_reduce_kernel, (int)mml_ceil_div(wp.slab / 4, 32), 256, 0, st, workspace, wp.splits, wp.slab, dw_krsc);
    } else {
      MML_LAUNCH(ctx, wgrad_tap_reduce_kernel, dim3((unsigned)mml_ceil_div((long long)g->K * g->C / 4, 256), n_taps), 256, 0, st, workspace, wp.splits, wp.slab, dw_krsc,
                                                                                                                  g->K, g->R * g->S, g->C, p);
    }
  }
  return MML_OK;
}

}  // extern "C"
