// Single-query ("decode") attention: every sample contributes at most ONE query token (max_seqlen_q == 1), GQA,
// head_dim 128 — the shape of Bagel.generate_text (reference modeling/bagel/bagel.py:930-1010 ->
// qwen2_navit.py:579-588 with causal=True, one new token against the whole cache).
//
// This is an HBM-bound stream over the K/V cache (2 * len * Hk * 128 * 2 bytes per sample), not a GEMM: the
// tcgen05 kernel in attn.cu would spend a 128-row query tile (and a whole CTA) on one valid row per q head and
// walk the keys serially. Here instead:
//   * grid = (split, Hk, batch): the keys of one (sample, kv head) are split over the CTAs of a thread-block
//     CLUSTER; each CTA handles all G = Hq/Hk query heads that share the kv head, so K and V are read exactly once;
//   * scores: one key row per lane (no cross-lane reduction), q broadcast from shared memory in fp32;
//   * P*V: one 4-wide slice of head_dim per lane, V rows read coalesced (256 B per row per warp);
//   * flash-decoding merge: warps -> CTA through shared memory, CTAs -> rank 0 of the cluster through distributed
//     shared memory, fixed order (deterministic), no workspace, one launch.
#include "common.cuh"
#include "host_util.h"
#include "attn_decode.h"

namespace bagel {

namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kD = 128;

struct DecodeParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* out;
  long long ld_q, ld_k, ld_v, ld_out;
  const int *cu_q, *cu_k, *seqused_k;
  int split;
  float scale_log2;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float* local, uint32_t rank) {
  uint32_t remote;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

template <int G>
__global__ void __launch_bounds__(kThreads, 3) attn_decode_kernel(const DecodeParams p) {
  __shared__ __align__(16) float sq[G][kD];            // q * softmax_scale * log2(e)
  __shared__ __align__(16) float sp[kWarps][G][32];    // probabilities of the warp's current 32-key block
  __shared__ __align__(16) float s_acc[kWarps][G][kD];
  __shared__ float s_m[kWarps][G], s_l[kWarps][G];
  __shared__ __align__(16) float part_acc[G][kD];      // this CTA's (un-normalised) partial, read by cluster rank 0
  __shared__ float part_m[G], part_l[G];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  const int rank = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  const int q_row = p.cu_q[b];
  const bool active = (p.cu_q[b + 1] - q_row) > 0;
  const int k_begin = p.cu_k[b];
  int len = p.seqused_k ? p.seqused_k[b] : (p.cu_k[b + 1] - k_begin);
  if (!active || len < 0) len = 0;
  const int chunk = (((len + p.split - 1) / p.split) + 31) & ~31;
  const int r0 = min(rank * chunk, len), r1 = min(r0 + chunk, len);

  if (active) {
#pragma unroll
    for (int h = 0; h < G; ++h)
      sq[h][t] = __bfloat162float(p.q[(long long)q_row * p.ld_q + (hk * G + h) * kD + t]) * p.scale_log2;
  }
  __syncthreads();

  float m[G], l[G], acc[G][4];
#pragma unroll
  for (int h = 0; h < G; ++h) {
    m[h] = -INFINITY;
    l[h] = 0.f;
    acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f;
  }

  const __nv_bfloat16* kbase = p.k + (long long)k_begin * p.ld_k + hk * kD;
  const __nv_bfloat16* vbase = p.v + (long long)k_begin * p.ld_v + hk * kD;

  for (int blk = r0 + warp * 32; blk < r1; blk += kWarps * 32) {
    const int row = blk + lane;
    const bool valid = row < r1;
    // ---- scores: lane <-> key row ----
    const uint4* kp = reinterpret_cast<const uint4*>(kbase + (long long)(valid ? row : r1 - 1) * p.ld_k);
    uint4 kk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) kk[i] = __ldg(kp + i);
    float s[G], s2[G];
#pragma unroll
    for (int h = 0; h < G; ++h) s[h] = s2[h] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {   // 8 channels per step, unpacked once and used by all G heads
      const float k0 = bf16_lo(kk[i].x), k1 = bf16_hi(kk[i].x), k2 = bf16_lo(kk[i].y), k3 = bf16_hi(kk[i].y);
      const float k4 = bf16_lo(kk[i].z), k5 = bf16_hi(kk[i].z), k6 = bf16_lo(kk[i].w), k7 = bf16_hi(kk[i].w);
#pragma unroll
      for (int h = 0; h < G; ++h) {
        const float4 qa = *reinterpret_cast<const float4*>(&sq[h][i * 8]);
        const float4 qb = *reinterpret_cast<const float4*>(&sq[h][i * 8 + 4]);
        s[h] = fmaf(k0, qa.x, s[h]);
        s2[h] = fmaf(k1, qa.y, s2[h]);
        s[h] = fmaf(k2, qa.z, s[h]);
        s2[h] = fmaf(k3, qa.w, s2[h]);
        s[h] = fmaf(k4, qb.x, s[h]);
        s2[h] = fmaf(k5, qb.y, s2[h]);
        s[h] = fmaf(k6, qb.z, s[h]);
        s2[h] = fmaf(k7, qb.w, s2[h]);
      }
    }
#pragma unroll
    for (int h = 0; h < G; ++h) s[h] = valid ? s[h] + s2[h] : -INFINITY;
    // ---- online softmax (block max over the warp; running state replicated in every lane) ----
#pragma unroll
    for (int h = 0; h < G; ++h) {
      float mb = s[h];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
      const float m_new = fmaxf(m[h], mb);                       // finite: the block has at least one valid key
      const float corr = (m[h] == -INFINITY) ? 0.f : ex2f(m[h] - m_new);
      const float pr = valid ? ex2f(s[h] - m_new) : 0.f;
      l[h] = l[h] * corr + pr;
      m[h] = m_new;
      sp[warp][h][lane] = pr;
      acc[h][0] *= corr; acc[h][1] *= corr; acc[h][2] *= corr; acc[h][3] *= corr;
    }
    __syncwarp();
    // ---- P*V: lane <-> 4 consecutive channels; 32 key rows, 8 in flight ----
    const int nrows = min(32, r1 - blk);
#pragma unroll 1
    for (int r8 = 0; r8 < nrows; r8 += 8) {
      uint2 vv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int rr = min(blk + r8 + j, r1 - 1);                // rows past the end carry p = 0
        vv[j] = __ldg(reinterpret_cast<const uint2*>(vbase + (long long)rr * p.ld_v + lane * 4));
      }
#pragma unroll
      for (int h = 0; h < G; ++h) {
        const float4 pa = *reinterpret_cast<const float4*>(&sp[warp][h][r8]);
        const float4 pb = *reinterpret_cast<const float4*>(&sp[warp][h][r8 + 4]);
        const float pj[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[h][0] = fmaf(pj[j], bf16_lo(vv[j].x), acc[h][0]);
          acc[h][1] = fmaf(pj[j], bf16_hi(vv[j].x), acc[h][1]);
          acc[h][2] = fmaf(pj[j], bf16_lo(vv[j].y), acc[h][2]);
          acc[h][3] = fmaf(pj[j], bf16_hi(vv[j].y), acc[h][3]);
        }
      }
    }
    __syncwarp();   // sp[warp] is rewritten by the next block
  }

  // ---- merge: lanes -> warp -> CTA ----
#pragma unroll
  for (int h = 0; h < G; ++h) {
    float ls = l[h];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
    *reinterpret_cast<float4*>(&s_acc[warp][h][lane * 4]) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
    if (lane == 0) {
      s_m[warp][h] = m[h];
      s_l[warp][h] = ls;
    }
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < G; ++h) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) M = fmaxf(M, s_m[w][h]);
    float L = 0.f, A = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const float wt = (s_m[w][h] == -INFINITY) ? 0.f : ex2f(s_m[w][h] - M);
      L = fmaf(s_l[w][h], wt, L);
      A = fmaf(s_acc[w][h][t], wt, A);
    }
    part_acc[h][t] = A;
    if (t == 0) {
      part_m[h] = M;
      part_l[h] = L;
    }
  }

  // ---- merge: CTAs of the cluster -> rank 0 (distributed shared memory), fixed rank order ----
  if (p.split > 1) cluster_sync_all(); else __syncthreads();
  if (rank == 0 && active) {
#pragma unroll
    for (int h = 0; h < G; ++h) {
      float M = -INFINITY;
      for (int r = 0; r < p.split; ++r) M = fmaxf(M, p.split > 1 ? ld_dsmem_f32(&part_m[h], r) : part_m[h]);
      float L = 0.f, A = 0.f;
      for (int r = 0; r < p.split; ++r) {
        const float mr = p.split > 1 ? ld_dsmem_f32(&part_m[h], r) : part_m[h];
        const float lr = p.split > 1 ? ld_dsmem_f32(&part_l[h], r) : part_l[h];
        const float ar = p.split > 1 ? ld_dsmem_f32(&part_acc[h][t], r) : part_acc[h][t];
        const float wt = (mr == -INFINITY) ? 0.f : ex2f(mr - M);
        L = fmaf(lr, wt, L);
        A = fmaf(ar, wt, A);
      }
      const float o = (L > 0.f) ? A / L : 0.f;
      p.out[(long long)q_row * p.ld_out + (hk * G + h) * kD + t] = __float2bfloat16_rn(o);
    }
  }
  if (p.split > 1) cluster_sync_all();   // peers keep their shared memory alive until rank 0 has read it
}

template <int G>
int launch(const DecodeParams& p, int Hk, int B, cudaStream_t stream) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)p.split, (unsigned)Hk, (unsigned)B);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)p.split;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  BAGEL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_decode_kernel<G>, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // namespace

bool attn_decode_supported(int max_seqlen_q, int head_dim, int Hq, int Hk) {
  static const bool on = [] { const char* e = getenv("BAGEL_ATTN_DECODE"); return !(e && atoi(e) == 0); }();
  if (!on || max_seqlen_q != 1 || head_dim != kD || Hk <= 0 || Hq % Hk) return false;
  const int g = Hq / Hk;
  return g == 1 || g == 2 || g == 4 || g == 7 || g == 8;
}

int attn_decode(const void* q, const void* k, const void* v, void* out, const int* cu_seqlens_q,
                const int* cu_seqlens_k, const int* seqused_k, int batch, int Hq, int Hk, int max_seqlen_k,
                float softmax_scale, long long ld_q, long long ld_k, long long ld_v, long long ld_out,
                cudaStream_t stream) {
  DecodeParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ld_q = ld_q; p.ld_k = ld_k; p.ld_v = ld_v; p.ld_out = ld_out;
  p.cu_q = cu_seqlens_q; p.cu_k = cu_seqlens_k; p.seqused_k = seqused_k;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  // keys per CTA ~128 (one 32-key block per warp), cluster size <= 8; unknown max length -> 8
  int split = (max_seqlen_k > 0) ? (max_seqlen_k + 127) / 128 : 8;
  if (split > 8) split = 8;
  if (split < 1) split = 1;
  // small batches: more CTAs per (sample, kv head) do not help below one block per warp; large batches: keep the
  // grid within a few waves
  while (split > 1 && (long long)batch * Hk * split > 16LL * sm_count()) split >>= 1;
  p.split = split;
  switch (Hq / Hk) {
    case 1: return launch<1>(p, Hk, batch, stream);
    case 2: return launch<2>(p, Hk, batch, stream);
    case 4: return launch<4>(p, Hk, batch, stream);
    case 7: return launch<7>(p, Hk, batch, stream);
    case 8: return launch<8>(p, Hk, batch, stream);
    default: return set_error(BAGEL_ERR_SHAPE, "attn_decode: unsupported GQA group %d", Hq / Hk);
  }
}

}  // namespace bagel
