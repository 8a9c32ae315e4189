// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue( A[M,K] * W[N,K]^T )
//
// Replaces every nn.Linear on BAGEL's forward path (reference: modeling/bagel/qwen2_navit.py:515-517,
// 529-536, 589-594; modeling/qwen2/modeling_qwen2.py:200-201; modeling/bagel/bagel.py:801-833), which
// the reference runs as cuBLASLt GEMM + separate ATen elementwise launches.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, kStages-deep smem ring)
//   warp 1      : MMA issuer     (one elected lane issues tcgen05.mma 128 x BN x 16, fp32 accum in TMEM)
//   warps 2..5  : epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
//
// TMEM holds two accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile
// i+1. A and W are both K-major ("TN" GEMM: nn.Linear weight layout), fp32 accumulation, bf16 output.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"
#include "gemm_skinny.h"
#include "gemm_params.h"
#include "gemm2.h"

namespace bagel {

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : (BN == 64 ? 8 : 10));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int EPI, bool CONV = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  static_assert(EPI != EPI_SWIGLU || BN == 256, "SwiGLU epilogue pairs 128 gate + 128 up columns");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                  // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;       // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;   // [2]        MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_k = CONV ? p.ksize * p.ksize * p.cin_chunks : (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_lane()) {   // not `lane == 0`: see common.cuh
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t hint_a = (p.hints & 2) ? kEvictFirst : ((p.hints & 8) ? kEvictLast : kEvictNormal);
      const uint64_t hint_w = (p.hints & 1) ? kEvictLast : ((p.hints & 16) ? kEvictFirst : kEvictNormal);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, p.num_m, p.num_n, p.group_m, p.group_n, m_blk, n_blk);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if constexpr (CONV) {
            // m_blk -> (image, tile row, tile col); K block -> (tap, channel chunk). Out-of-image coordinates
            // (the conv's zero padding, ragged tile edges) are zero-filled by TMA.
            const int tiles_img = p.tiles_w * p.tiles_h;
            const int img = m_blk / tiles_img, t_in = m_blk - img * tiles_img;
            const int h0 = (t_in / p.tiles_w) * p.th, w0 = (t_in % p.tiles_w) * p.tw;
            const int tap = kb / p.cin_chunks, cc = kb - tap * p.cin_chunks;
            const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
            tma_load_4d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], cc * BK, w0 * p.stride + kw - p.pad,
                        h0 * p.stride + kh - p.pad, img, hint_a);
          } else {
            tma_load_2d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * BK, m_blk * BM, hint_a);
          }
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * BK, n_blk * BN, hint_w);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one_lane()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle atom: +2 in the (addr>>4) field
            umma_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row_in_tile = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, p.num_m, p.num_n, p.group_m, p.group_n, m_blk, n_blk);
      int row;
      bool row_ok;
      long long out_row;
      if constexpr (CONV) {
        const int tiles_img = p.tiles_w * p.tiles_h;
        const int img = m_blk / tiles_img, t_in = m_blk - img * tiles_img;
        const int ho = (t_in / p.tiles_w) * p.th + row_in_tile / p.tw;
        const int wo = (t_in % p.tiles_w) * p.tw + row_in_tile % p.tw;
        row_ok = (ho < p.Ho) && (wo < p.Wo);
        row = (img * p.Ho + ho) * p.Wo + wo;
        out_row = row;
      } else {
        row = m_blk * BM + row_in_tile;
        row_ok = row < p.M;
        out_row = row;
        if (p.row_map != nullptr && row_ok) out_row = p.row_map[row];
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + acc * BN + (uint32_t(quarter * 32) << 16);

      if constexpr (EPI == EPI_QKV) {
        static_assert(BN == 256, "fused QKV epilogue: two 128-wide heads per tile");
        qkv_epilogue_row(p, t_acc, n_blk, row_ok, out_row);
      } else if constexpr (EPI == EPI_SWIGLU) {
        const int n_out0 = n_blk * (BN / 2);
        __nv_bfloat16* crow = p.C + out_row * p.ldc + n_out0;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t g[32], u[32];
          tmem_ld_x32(t_acc + c * 32, g);
          tmem_ld_x32(t_acc + 128 + c * 32, u);
          tmem_ld_wait();
          if (row_ok) {
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float h0, h1;
              {
                const float gg = bf16_round(__uint_as_float(g[2 * j]));
                const float uu = bf16_round(__uint_as_float(u[2 * j]));
                h0 = bf16_round(silu_f(gg)) * uu;
              }
              {
                const float gg = bf16_round(__uint_as_float(g[2 * j + 1]));
                const float uu = bf16_round(__uint_as_float(u[2 * j + 1]));
                h1 = bf16_round(silu_f(gg)) * uu;
              }
              o[j] = pack_bf16x2(h0, h1);
            }
            const int n0 = n_out0 + c * 32;
            if (n0 + 32 <= p.N / 2) {
              uint4* dst = reinterpret_cast<uint4*>(crow + c * 32);
#pragma unroll
              for (int q = 0; q < 4; ++q) store16(dst + q, make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]), p.hints & 4);
            } else {
              for (int j = 0; j < 16; ++j)
                if (n0 + 2 * j < p.N / 2) *reinterpret_cast<uint32_t*>(crow + c * 32 + 2 * j) = o[j];
            }
          }
        }
      } else {
        const int n0_tile = n_blk * BN;
        __nv_bfloat16* crow = p.C + out_row * p.ldc + n0_tile;
        const __nv_bfloat16* rrow = (EPI == EPI_RESID) ? p.resid + out_row * p.ldr + n0_tile : nullptr;
        const float* rrow32 = (EPI == EPI_RESID_F32) ? p.resid32 + out_row * p.ldr + n0_tile : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_x32(t_acc + c * 32, v);
          tmem_ld_wait();
          const int n0 = n0_tile + c * 32;
          if (row_ok && n0 < p.N) {
            const bool full = (n0 + 32 <= p.N);
            uint32_t rr[16];
            if constexpr (EPI == EPI_RESID) {
              if (full) {
                const uint4* src = reinterpret_cast<const uint4*>(rrow + c * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 t = src[q];
                  rr[4 * q] = t.x; rr[4 * q + 1] = t.y; rr[4 * q + 2] = t.z; rr[4 * q + 3] = t.w;
                }
              } else {
                for (int j = 0; j < 16; ++j)
                  rr[j] = (n0 + 2 * j < p.N) ? *reinterpret_cast<const uint32_t*>(rrow + c * 32 + 2 * j) : 0u;
              }
            }
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x0 = __uint_as_float(v[2 * j]);
              float x1 = __uint_as_float(v[2 * j + 1]);
              if (p.bias != nullptr) {
                const int nb = min(n0 + 2 * j, p.N - 2);
                const uint32_t bb = *reinterpret_cast<const uint32_t*>(p.bias + nb);
                x0 += bf16_lo(bb);
                x1 += bf16_hi(bb);
              }
              if constexpr (EPI == EPI_RESID) {
                x0 = bf16_lo(rr[j]) + bf16_round(x0);
                x1 = bf16_hi(rr[j]) + bf16_round(x1);
              } else if constexpr (EPI == EPI_RESID_F32) {
                if (n0 + 2 * j < p.N) {   // N % 8 == 0, so a pair is in or out together
                  const float2 r2 = *reinterpret_cast<const float2*>(rrow32 + c * 32 + 2 * j);
                  x0 = __fadd_rn(r2.x, bf16_round(x0));
                  x1 = __fadd_rn(r2.y, bf16_round(x1));
                }
              } else if constexpr (EPI == EPI_GELU) {
                x0 = gelu_tanh_f(bf16_round(x0));
                x1 = gelu_tanh_f(bf16_round(x1));
              } else if constexpr (EPI == EPI_SILU) {
                x0 = silu_f(bf16_round(x0));
                x1 = silu_f(bf16_round(x1));
              }
              if constexpr (EPI == EPI_F32 || EPI == EPI_RESID_F32) {
                v[2 * j] = __float_as_uint(x0);     // reuse the accumulator registers for the fp32 result
                v[2 * j + 1] = __float_as_uint(x1);
              }
              o[j] = pack_bf16x2(x0, x1);
            }
            if constexpr (EPI == EPI_F32 || EPI == EPI_RESID_F32) {
              float* d32 = p.C32 + out_row * p.ldc + n0;   // 128 contiguous bytes of this thread's row: 16-byte stores
              if (full) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  reinterpret_cast<uint4*>(d32)[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              } else {
                for (int j = 0; j < 32; ++j)
                  if (n0 + j < p.N) d32[j] = __uint_as_float(v[j]);
              }
            } else if (full) {
              uint4* dst = reinterpret_cast<uint4*>(crow + c * 32);
#pragma unroll
              for (int q = 0; q < 4; ++q) store16(dst + q, make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]), p.hints & 4);
            } else {
              for (int j = 0; j < 16; ++j)
                if (n0 + 2 * j < p.N) *reinterpret_cast<uint32_t*>(crow + c * 32 + 2 * j) = o[j];
            }
          }
        }
      }
      // all of this warp's tcgen05.ld have completed (wait::ld above) -> hand the accumulator stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BN, int EPI, bool CONV = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_kernel<BN, EPI, CONV>;
  static bool attr_done = false;  // per-instantiation; idempotent if raced
  if (!attr_done) {
    BAGEL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  if (!CONV) p.num_m = (p.M + BM - 1) / BM;  // CONV: set by the caller (images x tiles per image)
  p.num_n = (p.N + BN - 1) / BN;
  p.num_tiles = p.num_m * p.num_n;
  // Raster group: group_m M-tiles share one sweep over the N tiles. Measured on B200 at M=65568 (A/B in one process
  // per value, profiles/r01_gemm_group_ab.txt): wide outputs (148 N-tiles, gate|up) are fastest at 16, narrow ones
  // (14-18 N-tiles: qkv, o_proj, down_proj) at 32; 8 and 64 lose 5-15 % either way.
  {
    static const int env_g = [] { const char* e = getenv("BAGEL_GEMM_GROUP_M"); return e ? atoi(e) : 0; }();
    static const int env_n = [] { const char* e = getenv("BAGEL_GEMM_GROUP_N"); return e ? atoi(e) : -1; }();
    static const int env_h = [] { const char* e = getenv("BAGEL_GEMM_HINTS"); return e ? atoi(e) : -1; }();
    p.group_m = env_g > 0 ? env_g : (p.num_n >= 64 ? 16 : 32);
    // N super-tiles only where W does not fit the L2 beside the streams (gate|up: 272 MB); 0 / >= num_n = one sweep
    int gn = env_n >= 0 ? env_n : 0;
    if (gn <= 0 || gn > p.num_n) gn = p.num_n;
    p.group_n = gn;
    p.hints = env_h >= 0 ? env_h : 0;
  }
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <int BN>
static int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                        cudaStream_t s) {
  switch (epi) {
    case EPI_BIAS: return launch_gemm<BN, EPI_BIAS>(tmA, tmB, p, s);
    case EPI_RESID: return launch_gemm<BN, EPI_RESID>(tmA, tmB, p, s);
    case EPI_GELU: return launch_gemm<BN, EPI_GELU>(tmA, tmB, p, s);
    case EPI_SILU: return launch_gemm<BN, EPI_SILU>(tmA, tmB, p, s);
    case EPI_F32: return launch_gemm<BN, EPI_F32>(tmA, tmB, p, s);
    case EPI_RESID_F32: return launch_gemm<BN, EPI_RESID_F32>(tmA, tmB, p, s);
    default: return set_error(BAGEL_ERR_ARG, "bagel_gemm_bf16: unknown epilogue %d", epi);
  }
}

}  // namespace bagel

using namespace bagel;

extern "C" int bagel_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc,
                               int M, int N, int K, const void* bias, const void* resid, long long ldr,
                               const int* row_map, int epilogue, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(BAGEL_ERR_SHAPE, "bagel_gemm_bf16: M,N,K must be > 0");
  if ((lda % 8) || (ldw % 8) || (ldc % 8) || (K % 8) || (N % 8))
    return set_error(BAGEL_ERR_ALIGN, "bagel_gemm_bf16: K, N and leading dims must be multiples of 8 (16 B)");
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)C | (uintptr_t)bias | (uintptr_t)resid) & 15)
    return set_error(BAGEL_ERR_ALIGN, "bagel_gemm_bf16: pointers must be 16-byte aligned");
  if ((epilogue == EPI_RESID || epilogue == EPI_RESID_F32) && (resid == nullptr || (ldr % 8)))
    return set_error(BAGEL_ERR_ARG, "bagel_gemm_bf16: the residual epilogues need resid with ldr %% 8 == 0");
  if (int rc = require_sm100()) return rc;

  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.C = static_cast<__nv_bfloat16*>(C);
  p.ldc = ldc;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.resid = static_cast<const __nv_bfloat16*>(resid);
  p.resid32 = static_cast<const float*>(resid);   // BAGEL_EPI_RESID_F32: resid and C are fp32 [*, ldr] / [*, ldc]
  p.ldr = ldr;
  p.row_map = row_map;
  p.C32 = static_cast<float*>(C);  // used by the fp32-output epilogues (C is then an fp32 [M, ldc] buffer)
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  if (epilogue == EPI_SWIGLU && (N % 256))
    return set_error(BAGEL_ERR_SHAPE, "bagel_gemm_bf16: SwiGLU needs N (=2*I, interleaved) %% 256 == 0");
  // token-by-token decode / und-expert text rows: weight-streaming kernel with swapped operands and cluster split-K
  static const bool skinny_on = [] { const char* e = getenv("BAGEL_GEMM_SKINNY"); return !(e && atoi(e) == 0); }();
  if (skinny_on && gemm_skinny_supported(M, N, K, epilogue))
    return gemm_skinny(A, lda, W, ldw, C, ldc, M, N, K, bias, resid, ldr, row_map, epilogue, s);

  // the large projections: 256 x 256 tiles on CTA pairs (tcgen05 cta_group::2), gemm2.cu
  if (gemm2_supported(M, N, K, epilogue)) {
    CUtensorMap tmA2, tmB2;
    if (int rc = make_tmap_2d_bf16(&tmA2, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM)) return rc;
    if (int rc = make_tmap_2d_bf16(&tmB2, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, BK, 128)) return rc;
    return gemm2_launch(tmA2, tmB2, p, epilogue, s);
  }

  int bn;
  if (epilogue == EPI_SWIGLU) {
    if (N % 256) return set_error(BAGEL_ERR_SHAPE, "bagel_gemm_bf16: SwiGLU needs N (=2*I, interleaved) %% 256 == 0");
    bn = 256;
  } else {
    bn = (N % 256 == 0 || N >= 1024) ? 256 : (N > 64 ? 128 : 64);
    // skinny M (text decode, und-expert rows of a MoT layer): the GEMM is a weight stream, so use narrow tiles to put
    // 4x more CTAs (and TMA pipelines) on the W matrix
    if (M <= BM && N >= 1024) bn = (N <= 8192) ? 32 : 64;  // >= 112 CTAs on the 3584/4608-wide projections
  }
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmB, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, BK, bn)) return rc;

  if (epilogue == EPI_SWIGLU) return launch_gemm<256, EPI_SWIGLU>(tmA, tmB, p, s);
  if (bn == 256) return dispatch_epi<256>(epilogue, tmA, tmB, p, s);
  if (bn == 128) return dispatch_epi<128>(epilogue, tmA, tmB, p, s);
  if (bn == 32) return dispatch_epi<32>(epilogue, tmA, tmB, p, s);
  return dispatch_epi<64>(epilogue, tmA, tmB, p, s);
}


// ---------------------------------------------------------------------------------------------
// Implicit-GEMM 2-D convolution on NHWC bf16 activations (FLUX VAE: modeling/autoencoder.py:76-80, 102-108,
// 114-119, 139, 170, 221, 248 — the reference runs these as cuDNN NCHW convolutions).
//   out[b, ho, wo, :] = epilogue( sum_{kh,kw,c} x[b, ho*s + kh - pad, wo*s + kw - pad, c] * w[:, kh, kw, c] )
// No im2col buffer: each (tap, 64-channel chunk) K-slice of the A tile is one 4-D TMA box whose signed
// coordinates fall outside the image exactly where the convolution pads with zeros.
// ---------------------------------------------------------------------------------------------
extern "C" int bagel_conv2d_nhwc_bf16(const void* x, int B, int Hi, int Wi, int Cin, const void* w, int Cout, int ksize,
                                      int stride, int pad, const void* bias, const void* resid, void* out, int Ho,
                                      int Wo, void* stream) {
  if (ksize != 1 && ksize != 3) return set_error(BAGEL_ERR_ARG, "bagel_conv2d_nhwc_bf16: ksize must be 1 or 3");
  if (stride != 1 && stride != 2) return set_error(BAGEL_ERR_ARG, "bagel_conv2d_nhwc_bf16: stride must be 1 or 2");
  if (Cin % 64) return set_error(BAGEL_ERR_SHAPE, "bagel_conv2d_nhwc_bf16: Cin must be a multiple of 64 (pad channels)");
  if (Cout % 8) return set_error(BAGEL_ERR_SHAPE, "bagel_conv2d_nhwc_bf16: Cout must be a multiple of 8 (pad filters)");
  if (B <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0) return set_error(BAGEL_ERR_SHAPE, "bagel_conv2d_nhwc_bf16: bad sizes");
  if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)out | (uintptr_t)bias | (uintptr_t)resid) & 15)
    return set_error(BAGEL_ERR_ALIGN, "bagel_conv2d_nhwc_bf16: pointers must be 16-byte aligned");
  if (int rc = require_sm100()) return rc;

  GemmParams p{};
  p.Ho = Ho; p.Wo = Wo;
  int tw = 128;
  while (tw > Wo) tw >>= 1;  // largest power of two <= min(Wo, 128)
  if (tw < 1) tw = 1;
  p.tw = tw; p.th = 128 / tw;
  p.tiles_w = (Wo + p.tw - 1) / p.tw;
  p.tiles_h = (Ho + p.th - 1) / p.th;
  p.ksize = ksize; p.pad = pad; p.stride = stride; p.cin_chunks = Cin / 64;
  p.M = B * Ho * Wo; p.N = Cout; p.K = ksize * ksize * Cin;
  p.num_m = B * p.tiles_w * p.tiles_h;
  p.C = static_cast<__nv_bfloat16*>(out);
  p.ldc = Cout;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.resid = static_cast<const __nv_bfloat16*>(resid);
  p.ldr = Cout;
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  CUtensorMap tmA, tmB;
  if (int rc = make_tmap_4d_nhwc_bf16(&tmA, x, B, Hi, Wi, Cin, 64, p.tw, p.th, stride)) return rc;
  const int bn = (Cout % 256 == 0 || Cout >= 1024) ? 256 : (Cout > 64 ? 128 : 64);
  if (int rc = make_tmap_2d_bf16(&tmB, w, (uint64_t)p.K, (uint64_t)Cout, (uint64_t)p.K, BK, bn)) return rc;
  const bool res = resid != nullptr;
  if (bn == 256) return res ? launch_gemm<256, EPI_RESID, true>(tmA, tmB, p, s) : launch_gemm<256, EPI_BIAS, true>(tmA, tmB, p, s);
  if (bn == 128) return res ? launch_gemm<128, EPI_RESID, true>(tmA, tmB, p, s) : launch_gemm<128, EPI_BIAS, true>(tmA, tmB, p, s);
  return res ? launch_gemm<64, EPI_RESID, true>(tmA, tmB, p, s) : launch_gemm<64, EPI_BIAS, true>(tmA, tmB, p, s);
}


// QKV projection with the whole pre-attention tail fused into the epilogue (head_dim 128):
//   [q|k|v] = A W^T + b ; per-head RMSNorm(q,k) with expert-routed weights ; RoPE ; bf16 ; q -> q_out,
//   k/v -> merged KV buffers at kv_rows[row]. Replaces bagel_gemm_bf16 + bagel_qk_norm_rope (and the [N, 4608]
//   round trip through HBM between them).
extern "C" int bagel_gemm_qkv_norm_rope(const void* A, long long lda, const void* W, long long ldw, const void* bias,
                                        int M, int K, const int* row_map, const void* q_w0, const void* k_w0,
                                        const void* q_w1, const void* k_w1, const uint8_t* expert, const float* cos_t,
                                        const float* sin_t, void* q_out, long long ld_q, void* k_out, void* v_out,
                                        long long ld_kv, const int* kv_rows, int Hq, int Hk, float eps, int fp32_flow,
                                        void* stream) {
  const int N = (Hq + 2 * Hk) * 128;
  if (M <= 0 || K <= 0 || Hq <= 0 || Hk <= 0) return set_error(BAGEL_ERR_SHAPE, "bagel_gemm_qkv_norm_rope: bad sizes");
  if (N % 256) return set_error(BAGEL_ERR_SHAPE, "bagel_gemm_qkv_norm_rope: Hq + 2*Hk must be even (two heads per tile)");
  if ((lda % 8) || (ldw % 8) || (K % 8) || (ld_q % 8) || (ld_kv % 8))
    return set_error(BAGEL_ERR_ALIGN, "bagel_gemm_qkv_norm_rope: K and leading dims must be multiples of 8");
  if (bias == nullptr || q_w0 == nullptr || k_w0 == nullptr || cos_t == nullptr || sin_t == nullptr)
    return set_error(BAGEL_ERR_ARG, "bagel_gemm_qkv_norm_rope: bias, norm weights and RoPE tables are required");
  if (int rc = require_sm100()) return rc;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.row_map = row_map;
  if (fp32_flow < 0 || fp32_flow > 3) return set_error(BAGEL_ERR_ARG, "bagel_gemm_qkv_norm_rope: flow must be 0..3");
  p.qkv.qw0 = q_w0; p.qkv.kw0 = k_w0;
  p.qkv.qw1 = q_w1; p.qkv.kw1 = k_w1;
  p.qkv.expert = expert; p.qkv.cos_t = cos_t; p.qkv.sin_t = sin_t;
  p.qkv.q_out = static_cast<__nv_bfloat16*>(q_out); p.qkv.k_out = static_cast<__nv_bfloat16*>(k_out);
  p.qkv.v_out = static_cast<__nv_bfloat16*>(v_out);
  p.qkv.ld_q = ld_q; p.qkv.ld_kv = ld_kv; p.qkv.kv_rows = kv_rows;
  p.qkv.Hq = Hq; p.qkv.Hk = Hk; p.qkv.eps = eps; p.qkv.fp32_flow = fp32_flow;
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM)) return rc;
  // the large calls (all rows of a denoising step): CTA-pair kernel, each CTA stages half of the 256-row W tile (gemm2.cu)
  if (gemm2_supported(M, N, K, EPI_QKV)) {
    if (int rc = make_tmap_2d_bf16(&tmB, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, BK, 128)) return rc;
    return gemm2_launch(tmA, tmB, p, EPI_QKV, static_cast<cudaStream_t>(stream));
  }
  if (int rc = make_tmap_2d_bf16(&tmB, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, BK, 256)) return rc;
  return launch_gemm<256, EPI_QKV>(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}
