// CTA-pair bf16 GEMM for sm_100a:  C[M,N] = epilogue( A[M,K] * W[N,K]^T ), 256 x 256 tiles on TWO SMs.
//
// The large projections of the denoising step (gate|up + SwiGLU, down, qkv-less o_proj: M = 65568 rows) are bound by the
// power budget, not by issue slots: what a 128 x 256 single-CTA tile pays per FLOP is operand movement (48 KB of L2 ->
// shared-memory traffic and 48 KB of shared-memory reads per 128x256x64 MMA block). A CTA PAIR (two CTAs of a 2-CTA
// cluster on one TPC, tcgen05 `cta_group::2`) computes a 256 x 256 tile with ONE MMA stream: each CTA stages its own 128
// rows of A and only HALF of the W tile (128 of the 256 rows), the tensor cores of both SMs read both halves. Per CTA and
// K block that is 32 KB instead of 48 KB (-33 % L2->smem and smem->tensor-core bytes). The smem ring has room for 6 stages;
// 4 are used (see gemm2_launch: prefetching deeper costs more in L2 misses than it hides).
//
//   warp 0 (both CTAs)   TMA producer: A tile [128 x 64] of its M half, W half-tile [128 x 64]; every load completes on
//                        the LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2, peer bit cleared)
//   warp 1 (leader)      one lane issues tcgen05.mma.cta_group::2 (M 256, N 256, K 16) — fp32 accumulators: 128 rows x
//                        256 columns in EACH CTA's TMEM, two stages; tcgen05.commit ... multicast frees the smem slot in
//                        both CTAs / hands the accumulator to both epilogues
//   warps 2..5 (both)    epilogue of the CTA's own 128 rows: tcgen05.ld -> fused bias / residual / SwiGLU -> 32x32 sub-tiles
//                        staged in (swizzled) shared memory -> cp.async.bulk.tensor stores (direct stores with row_map);
//                        both CTAs' warps arrive on the leader's "accumulator drained" barrier
// Reference ops replaced: the same nn.Linear calls as gemm.cu (modeling/bagel/qwen2_navit.py:589-594,
// modeling/qwen2/modeling_qwen2.py:200-201) — this kernel is selected by bagel_gemm_bf16 for large M.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"
#include "gemm2.h"

namespace bagel {

constexpr int kPairBN = 256;                    // N of the pair tile
constexpr int kPairBNH = kPairBN / 2;           // W rows staged by each CTA
constexpr int kPairStages = 6;
constexpr int kPairABytes = BM * BK * 2;        // 16 KB
constexpr int kPairBBytes = kPairBNH * BK * 2;  // 16 KB
constexpr int kPairStageBytes = kPairABytes + kPairBBytes;
constexpr int kPairStgBytes = 4 * 2 * 2048;   // epilogue staging for TMA stores: 4 warps x 2 buffers x [32 rows x 32 bf16]
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + kPairStgBytes + 1024 + 256;

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  constexpr int BN = kPairBN;
  constexpr int kStages = kPairStages;
  extern __shared__ uint8_t smem_raw[];
  // identical offsets in both CTAs (the MMA and the multicast commits address "the same place" in the peer)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kPairABytes;
  uint8_t* smem_stg = smem + kStages * kPairStageBytes;   // 1024-byte aligned (64-byte swizzle atoms are 512 bytes)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + kPairStgBytes);
  uint64_t* full_bar = bars;                 // [kStages]  TMA (both CTAs) -> MMA        (leader's copy is used)
  uint64_t* empty_bar = bars + kStages;      // [kStages]  MMA -> TMA                    (each CTA its own, multicast arrive)
  uint64_t* tfull_bar = bars + 2 * kStages;  // [2]        MMA -> epilogue               (each CTA its own, multicast arrive)
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]        epilogues of both CTAs -> MMA (leader's copy is used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_k = (p.K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int nstages = p.stages;   // <= kStages

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmC);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);   // the leader's arrive.expect_tx (bytes of BOTH CTAs)
      mbar_init(&empty_bar[i], 1);  // one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);  // 4 epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // barrier inits of the peer are visible before any remote arrive / TMA completion targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one_lane()) {   // not `lane == 0`: see common.cuh
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t hint_a = (p.hints & 2) ? kEvictFirst : ((p.hints & 8) ? kEvictLast : kEvictNormal);
      const uint64_t hint_w = (p.hints & 1) ? kEvictLast : ((p.hints & 16) ? kEvictFirst : kEvictNormal);
      for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
        int mp, n_blk;
        tile_coords(tile, p.num_m, p.num_n, p.group_m, p.group_n, mp, n_blk);
        const int m_blk = 2 * mp + (int)rank;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
          tma_load_2d_2sm(smem_a + stage * kPairABytes, &tmA, &full_bar[stage], kb * BK, m_blk * BM, hint_a);
          tma_load_2d_2sm(smem_b + stage * kPairBBytes, &tmB, &full_bar[stage], kb * BK, n_blk * BN + (int)rank * kPairBNH,
                          hint_w);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one_lane()) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, 0);  // M = 256 across the pair
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(smem_a + stage * kPairABytes));
          const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(smem_b + stage * kPairBBytes));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_ss_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2sm_mc(&empty_bar[stage], 0b11);  // slot reusable in both CTAs once these MMAs retire
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm_mc(&tfull_bar[acc], 0b11);  // accumulator complete -> epilogue warps of both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2..5), both CTAs: the CTA's own 128 rows =====================
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const bool streaming = (p.hints & 4) != 0;
    const bool tma_out = p.tma_store != 0;
    uint8_t* stg = smem_stg + quarter * 4096;   // this warp's two 2 KB staging buffers
    int stg_buf = 0;
    // One 32-row x 32-column bf16 sub-tile of C: registers -> swizzled shared memory -> ONE bulk tensor store issued by lane 0
    // (full 64-byte row segments leave the SM as whole sectors instead of 32 scattered 16-byte stores per instruction; rows
    // past M are clipped by the TMA unit). The buffer is reused two stores later: wait until that store has READ it.
    auto store_tile_tma = [&](const uint32_t (&o)[16], int col0, int row0) {
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      uint8_t* b = stg + stg_buf * 2048;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(b + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) =
            make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmC, b, col0, row0);
        tma_store_commit();
      }
      stg_buf ^= 1;
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
      int mp, n_blk;
      tile_coords(tile, p.num_m, p.num_n, p.group_m, p.group_n, mp, n_blk);
      const int row = (2 * mp + (int)rank) * BM + row_in_tile;
      const bool row_ok = row < p.M;
      long long out_row = row;
      if (p.row_map != nullptr && row_ok) out_row = p.row_map[row];

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + acc * BN + (uint32_t(quarter * 32) << 16);

      if constexpr (EPI == EPI_QKV) {
        // fused q/k-norm + RoPE + K/V placement: each CTA's accumulator holds its own 128 rows x the pair's 256 columns = two
        // whole heads, exactly the tile the 1-CTA kernel's epilogue works on (gemm_params.h)
        qkv_epilogue_row(p, t_acc, n_blk, row_ok, out_row);
      } else if constexpr (EPI == EPI_SWIGLU) {
        // W rows are interleaved per 256 (128 gate | 128 up): columns [0,128) of the tile = gate (staged by the leader),
        // [128,256) = up (staged by the peer) of the same 128 output features
        const int n_out0 = n_blk * (BN / 2);
        __nv_bfloat16* crow = p.C + out_row * p.ldc + n_out0;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t g[32], u[32];
          tmem_ld_x32(t_acc + c * 32, g);
          tmem_ld_x32(t_acc + 128 + c * 32, u);
          tmem_ld_wait();
          if (row_ok || tma_out) {
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float g0 = bf16_round(__uint_as_float(g[2 * j])), u0 = bf16_round(__uint_as_float(u[2 * j]));
              const float g1 = bf16_round(__uint_as_float(g[2 * j + 1])), u1 = bf16_round(__uint_as_float(u[2 * j + 1]));
              o[j] = pack_bf16x2(bf16_round(silu_f(g0)) * u0, bf16_round(silu_f(g1)) * u1);
            }
            if (tma_out) {
              store_tile_tma(o, n_out0 + c * 32, (2 * mp + (int)rank) * BM + quarter * 32);
            } else {
              uint4* dst = reinterpret_cast<uint4*>(crow + c * 32);   // N % 256 == 0: always a full chunk
#pragma unroll
              for (int q = 0; q < 4; ++q) store16(dst + q, make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]), streaming);
            }
          }
        }
      } else {
        const int n0_tile = n_blk * BN;
        __nv_bfloat16* crow = p.C + out_row * p.ldc + n0_tile;
        const __nv_bfloat16* rrow = (EPI == EPI_RESID) ? p.resid + out_row * p.ldr + n0_tile : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_x32(t_acc + c * 32, v);
          tmem_ld_wait();
          if (row_ok || tma_out) {
            uint32_t rr[16];
            if constexpr (EPI == EPI_RESID) {
              if (row_ok) {
                const uint4* src = reinterpret_cast<const uint4*>(rrow + c * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 t = src[q];
                  rr[4 * q] = t.x; rr[4 * q + 1] = t.y; rr[4 * q + 2] = t.z; rr[4 * q + 3] = t.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) rr[j] = 0u;   // rows past M: computed but clipped by the TMA store
              }
            }
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x0 = __uint_as_float(v[2 * j]);
              float x1 = __uint_as_float(v[2 * j + 1]);
              if (p.bias != nullptr) {
                const uint32_t bb = *reinterpret_cast<const uint32_t*>(p.bias + n0_tile + c * 32 + 2 * j);
                x0 += bf16_lo(bb);
                x1 += bf16_hi(bb);
              }
              if constexpr (EPI == EPI_RESID) {
                x0 = bf16_lo(rr[j]) + bf16_round(x0);
                x1 = bf16_hi(rr[j]) + bf16_round(x1);
              }
              o[j] = pack_bf16x2(x0, x1);
            }
            if (tma_out) {
              store_tile_tma(o, n0_tile + c * 32, (2 * mp + (int)rank) * BM + quarter * 32);
            } else {
              uint4* dst = reinterpret_cast<uint4*>(crow + c * 32);   // N % 256 == 0: always a full chunk
#pragma unroll
              for (int q = 0; q < 4; ++q) store16(dst + q, make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]), streaming);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);  // the leader's barrier counts both CTAs' epilogue warps
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (tma_out && lane == 0) tma_store_wait<0>();   // the staging buffers must outlive the last bulk stores
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer's shared memory / TMEM must outlive the leader's last MMA and its own last epilogue
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

template <int EPI>
static int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                   cudaStream_t stream) {
  auto kern = gemm2_bf16_kernel<EPI>;
  static bool attr_done = false;
  if (!attr_done) {
    BAGEL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    attr_done = true;
  }
  int clusters = sm_count() / 2;
  if (clusters > p.num_tiles) clusters = p.num_tiles;
  kern<<<2 * clusters, kGemmThreads, kPairSmemBytes, stream>>>(tmA, tmB, tmC, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

bool gemm2_supported(int M, int N, int K, int epilogue) {
  static const int mode = [] { const char* e = getenv("BAGEL_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  if (!mode) return false;
  // the fused-QKV epilogue on the pair kernel passes every test but measured 0.6 % SLOWER per denoising step than the 1-CTA kernel
  // on one box (763.4 / 767.9 / 763.7 ms for off / on / off, profiles/r02_gemm_pair_ab.txt (6)) -> opt-in
  static const int qkv = [] { const char* e = getenv("BAGEL_GEMM_PAIR_QKV"); return e ? atoi(e) : 0; }();
  if (epilogue == EPI_QKV && !qkv) return false;
  if (epilogue != EPI_BIAS && epilogue != EPI_RESID && epilogue != EPI_SWIGLU && epilogue != EPI_QKV) return false;
  return M >= 4 * BM && (N % kPairBN) == 0 && K >= BK;
}

int gemm2_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams p, int epilogue, cudaStream_t stream) {
  p.num_m = ((p.M + BM - 1) / BM + 1) / 2;  // M tile PAIRS (an odd last pair: the peer's rows are out of range, TMA zero-fills)
  p.num_n = p.N / kPairBN;
  p.num_tiles = p.num_m * p.num_n;
  {
    static const int env_g = [] { const char* e = getenv("BAGEL_GEMM_GROUP_M"); return e ? atoi(e) : 0; }();
    static const int env_n = [] { const char* e = getenv("BAGEL_GEMM_GROUP_N"); return e ? atoi(e) : -1; }();
    static const int env_h = [] { const char* e = getenv("BAGEL_GEMM_HINTS"); return e ? atoi(e) : -1; }();
    // 74 clusters in flight cover group_m pairs x (74 / group_m) N tiles. Measured INSIDE the power-capped denoising step on
    // one box (profiles/r02_gemm_pair_ab.txt): groups of 16 pairs (32 M tiles: W is streamed from DRAM 16x instead of 32x
    // per launch, the 29 MB A panel of a group stays L2-resident) 809-811 ms per step, 8 pairs 836 ms, 1-CTA kernel 816 ms.
    p.group_m = env_g > 0 ? (env_g + 1) / 2 : 16;
    // Large K (down_proj, K = 18944): the A panel of 16 pairs is 155 MB and cannot stay in the L2 between the sweeps over
    // the N tiles. A/B knobs for such GEMMs (K > 8192): BAGEL_GEMM_BIGK = "group_pairs,group_n,hints".
    static int bk_gp = 0, bk_gn = 0, bk_h = 0;
    static const bool bk_set = [] {
      const char* e = getenv("BAGEL_GEMM_BIGK");
      return e && sscanf(e, "%d,%d,%d", &bk_gp, &bk_gn, &bk_h) == 3;
    }();
    const bool bigk = bk_set && p.K > 8192;
    if (bigk && bk_gp > 0) p.group_m = bk_gp;
    int gn = env_n >= 0 ? env_n : 0;
    if (bigk) gn = bk_gn;
    if (gn <= 0 || gn > p.num_n) gn = p.num_n;
    p.group_n = gn;
    p.hints = bigk ? bk_h : (env_h >= 0 ? env_h : 0);
    static const int env_s = [] { const char* e = getenv("BAGEL_GEMM_PAIR_STAGES"); return e ? atoi(e) : 0; }();
    // FOUR stages in use although six fit: with a deeper ring the 74 pairs prefetch so far ahead that the A panel of the
    // raster group no longer survives in the L2 between its reuses — DRAM reads 8.6-24 GB per gate|up launch with 6 stages,
    // 11.4 GB with 5, 5.1 GB with 4 (= the compulsory 16 x W + 1 x A of this raster), and the power-capped step follows the
    // DRAM bytes: 813 / 785 / 782 ms (profiles/r02_gemm_pair_ab.txt). Three stages starve the tensor pipe (841 ms).
    p.stages = (env_s >= 2 && env_s <= kPairStages) ? env_s : 4;
  }
  // epilogue through shared memory + bulk tensor stores unless the output rows are scattered (row_map)
  static const bool tma_store_on = [] { const char* e = getenv("BAGEL_GEMM_TMA_STORE"); return !(e && atoi(e) == 0); }();
  p.tma_store = (tma_store_on && p.row_map == nullptr && epilogue != EPI_QKV) ? 1 : 0;   // the QKV epilogue writes q / K / V rows itself
  CUtensorMap tmC{};
  if (p.tma_store) {
    const uint64_t out_cols = epilogue == EPI_SWIGLU ? (uint64_t)p.N / 2 : (uint64_t)p.N;
    if (int rc = make_tmap_2d_bf16_store(&tmC, p.C, out_cols, (uint64_t)p.M, (uint64_t)p.ldc, 32, 32)) return rc;
  }
  switch (epilogue) {
    case EPI_BIAS: return launch2<EPI_BIAS>(tmA, tmB, tmC, p, stream);
    case EPI_RESID: return launch2<EPI_RESID>(tmA, tmB, tmC, p, stream);
    case EPI_SWIGLU: return launch2<EPI_SWIGLU>(tmA, tmB, tmC, p, stream);
    case EPI_QKV: return launch2<EPI_QKV>(tmA, tmB, tmC, p, stream);
    default: return set_error(BAGEL_ERR_ARG, "gemm2: unsupported epilogue %d", epilogue);
  }
}

}  // namespace bagel
