// Shared between the 1-CTA (gemm.cu) and the CTA-pair (gemm2.cu) GEMM kernels: parameters, epilogue ids, rasterisation.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"

namespace bagel {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B atom along K
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 192;

enum GemmEpilogue : int {
  EPI_BIAS = 0,    // C = bf16(acc + bias)
  EPI_RESID = 1,   // C = bf16(resid + bf16(acc + bias))                (o_proj / down_proj + residual add)
  EPI_SWIGLU = 2,  // C[:, j] = bf16(bf16(silu(bf16 g_j)) * bf16 u_j); W rows interleaved per 256 (128 g | 128 u)
  EPI_GELU = 3,    // C = bf16(gelu_tanh(bf16(acc + bias)))             (SigLIP MLP / connector)
  EPI_SILU = 4,    // C = bf16(silu(bf16(acc + bias)))                  (timestep MLP)
  EPI_F32 = 5,     // C32 = acc (+ bias) as fp32                        (attention logits of the VAE mid block)
  EPI_QKV = 6,     // fused q/k RMSNorm + RoPE + bf16 cast + K/V placement  (PackedAttentionMoT, head_dim 128)
  EPI_RESID_F32 = 7,  // C32 = resid32 + bf16(acc + bias): fp32 residual stream of dtype mode B (fp32 master weights)
};

// Extra arguments of the fused QKV epilogue (see bagel_gemm_qkv_norm_rope in include/bagel_b200.h).
struct QkvEpi {
  const void *qw0, *kw0, *qw1, *kw1;  // per-head RMSNorm weights [128]: und / gen expert (may be null); bf16, fp32 for flow >= 2
  const uint8_t* expert;                       // [rows] 1 = gen expert
  const float *cos_t, *sin_t;                  // [rows, 64]
  __nv_bfloat16 *q_out, *k_out, *v_out;
  long long ld_q, ld_kv;
  const int* kv_rows;                          // destination row of each token in the merged K/V buffers
  int Hq, Hk;
  float eps;
  int fp32_flow;  // rounding-point flow 0..3, see qk_norm_rope_kernel in elementwise.cu
};

struct GemmParams {
  int M, N, K;
  __nv_bfloat16* C;
  long long ldc;
  const __nv_bfloat16* bias;   // [N] or null
  const __nv_bfloat16* resid;  // [*, ldr] or null (EPI_RESID)
  const float* resid32;        // EPI_RESID_F32
  long long ldr;
  const int* row_map;  // optional: output (and residual) row of A-row r is row_map[r]
  int num_m, num_n, num_tiles;
  int group_m;  // rasterisation: group_m M-tiles share one sweep over the N tiles (their A panels stay in L2)
  int group_n;  // N super-tiles: all M groups sweep group_n N-tiles before the next group_n (that W sub-panel stays in L2)
  int tma_store;  // CTA-pair kernel: epilogue writes C through shared memory + cp.async.bulk.tensor stores (no row_map)
  int stages;   // CTA-pair kernel: smem ring stages actually used (<= compile-time depth; A/B knob)
  int hints;    // L2 policy bits: 1 W evict_last, 2 A evict_first, 4 streaming (evict-first) output stores, 8 A evict_last,
                // 16 W evict_first
  // --- implicit-GEMM convolution (CONV kernels only): A is an NHWC activation tensor [B, Hi, Wi, Cin] read through a
  // 4-D TMA map; an M tile is a th x tw patch of output pixels (th*tw = 128) of one image; K runs over
  // (tap, 64-channel chunk); output / residual rows are NHWC pixel indices.
  int Ho, Wo;            // output spatial size
  int tw, th;            // tile width / height in output pixels
  int tiles_w, tiles_h;  // tiles per image
  int ksize, pad;        // 1 or 3; left/top zero padding in input pixels
  int stride;            // 1 or 2 (the TMA map carries the element stride)
  int cin_chunks;        // Cin / 64
  float* C32;            // optional fp32 output (EPI_F32)
  QkvEpi qkv;            // EPI_QKV only
};

// Two-level raster. Outer: N super-tiles of group_n N-tiles (the W sub-panel of a super-tile, group_n * BN * K * 2 bytes, is
// what should stay L2-resident while every M group sweeps it). Inner: groups of group_m M-tiles, M fastest, so the ~148
// CTAs running at any moment cover group_m x (148 / group_m) tiles and share their A / W tiles through the L2.
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int group_m, int group_n, int& m_blk,
                                            int& n_blk) {
  const int super_size = group_n * num_m;
  const int s = tile / super_size;
  const int n0 = s * group_n;
  const int nn = min(group_n, num_n - n0);
  const int rem = tile - s * super_size;
  const int group_size = group_m * nn;
  const int g = rem / group_size;
  const int first_m = g * group_m;
  const int gm = min(num_m - first_m, group_m);
  const int local = rem - g * group_size;
  m_blk = first_m + local % gm;
  n_blk = n0 + local / gm;
}

__device__ __forceinline__ void store16(void* dst, const uint4& v, bool streaming) {
  if (streaming) __stcs(reinterpret_cast<uint4*>(dst), v);   // st.global.cs: evict-first, the output is not re-read by this kernel
  else *reinterpret_cast<uint4*>(dst) = v;
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // torch "gelu_pytorch_tanh": 0.5 x (1 + tanh( sqrt(2/pi) (x + 0.044715 x^3) ))
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(inner));
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }


// Fused QKV epilogue of ONE token row (thread = row) over a 256-wide accumulator tile = two heads of 128: everything the
// reference does between the projection and flash-attn (qwen2_navit.py:518-519 / 542-574) on the fp32 accumulators —
// bf16(acc + bias), per-head RMSNorm with expert-routed weights, RoPE, bf16 cast, q / K / V rows written to their final
// places. Shared by the 1-CTA kernel (gemm.cu) and the CTA-pair kernel (gemm2.cu): in both a thread of the epilogue warps owns
// one row of its CTA's 128 x 256 accumulator tile at TMEM address t_acc.
__device__ __forceinline__ void qkv_epilogue_row(const GemmParams& p, uint32_t t_acc, int n_blk, bool row_ok, long long out_row) {
  const QkvEpi& e = p.qkv;
  const bool gen = row_ok && e.expert != nullptr && e.qw1 != nullptr && e.expert[out_row];
  const long long kv_row = (row_ok && e.kv_rows != nullptr) ? (long long)e.kv_rows[out_row] : out_row;
#pragma unroll 1
  for (int hh = 0; hh < 2; ++hh) {
    const int head = n_blk * 2 + hh;
    float x[128];
    {
      uint32_t* xr = reinterpret_cast<uint32_t*>(x);
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(t_acc + hh * 128 + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&xr[c * 32]));
      tmem_ld_wait();
    }
    if (!row_ok) continue;
    // q/k/v_proj output as the reference sees it: bf16(acc + bias)
    const __nv_bfloat16* bh = p.bias + head * 128;
#pragma unroll
    for (int i = 0; i < 128; i += 2) {
      const uint32_t bb = *reinterpret_cast<const uint32_t*>(bh + i);
      x[i] = bf16_round(x[i] + bf16_lo(bb));
      x[i + 1] = bf16_round(x[i + 1] + bf16_hi(bb));
    }
    __nv_bfloat16* dst;
    const bool is_v = head >= e.Hq + e.Hk;
    if (head < e.Hq) dst = e.q_out + out_row * e.ld_q + head * 128;
    else if (!is_v) dst = e.k_out + kv_row * e.ld_kv + (head - e.Hq) * 128;
    else dst = e.v_out + kv_row * e.ld_kv + (head - e.Hq - e.Hk) * 128;
    if (!is_v) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 128; ++i) ss += x[i] * x[i];
      const float r = rsqrtf(ss * (1.0f / 128.0f) + e.eps);
      const void* w = (head < e.Hq) ? (gen ? e.qw1 : e.qw0) : (gen ? e.kw1 : e.kw0);
      const bool wf32 = e.fp32_flow >= 2;
      const float* cs = e.cos_t + out_row * 64;
      const float* sn = e.sin_t + out_row * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(cs + i);
        const float4 s4 = *reinterpret_cast<const float4*>(sn + i);
        const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float wa = wf32 ? static_cast<const float*>(w)[i + u]
                                : __bfloat162float(static_cast<const __nv_bfloat16*>(w)[i + u]);
          const float wb = wf32 ? static_cast<const float*>(w)[64 + i + u]
                                : __bfloat162float(static_cast<const __nv_bfloat16*>(w)[64 + i + u]);
          float ya, yb, oa, ob;
          if (e.fp32_flow) {
            const float na = (e.fp32_flow == 2) ? bf16_round(x[i + u] * r) : __fmul_rn(x[i + u], r);
            const float nb = (e.fp32_flow == 2) ? bf16_round(x[64 + i + u] * r) : __fmul_rn(x[64 + i + u], r);
            ya = __fmul_rn(wa, na);
            yb = __fmul_rn(wb, nb);
            oa = __fadd_rn(__fmul_rn(ya, cc[u]), __fmul_rn(-yb, sv[u]));
            ob = __fadd_rn(__fmul_rn(yb, cc[u]), __fmul_rn(ya, sv[u]));
          } else {
            ya = bf16_round(wa * bf16_round(x[i + u] * r));
            yb = bf16_round(wb * bf16_round(x[64 + i + u] * r));
            oa = bf16_round(ya * cc[u]) + bf16_round(-yb * sv[u]);
            ob = bf16_round(yb * cc[u]) + bf16_round(ya * sv[u]);
          }
          x[i + u] = oa;
          x[64 + i + u] = ob;
        }
      }
    }
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int q = 0; q < 16; ++q)
      d4[q] = make_uint4(pack_bf16x2(x[8 * q], x[8 * q + 1]), pack_bf16x2(x[8 * q + 2], x[8 * q + 3]),
                         pack_bf16x2(x[8 * q + 4], x[8 * q + 5]), pack_bf16x2(x[8 * q + 6], x[8 * q + 7]));
  }
}

}  // namespace bagel
