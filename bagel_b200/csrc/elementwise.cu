// HBM-bound glue kernels of the MoT forward path (everything that is not a GEMM or attention).
// Each replaces a chain of ATen elementwise / index launches in the reference and keeps the reference's bf16
// rounding points ("mode A": bf16 weights under autocast). All are simple streaming kernels:
// 16-byte vector loads/stores, one warp per row (or per head), fp32 math, warp-shuffle reductions.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "host_util.h"

namespace bagel {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// RMSNorm with per-row expert routing (Qwen2RMSNorm, modeling/qwen2/modeling_qwen2.py:54-59; MoT routing
// of input/post-attention/final norms, modeling/bagel/qwen2_navit.py:781-787, 808-815, 1075-1082).
//   y = bf16( w_e * bf16( x * rsqrt(mean(x^2) + eps) ) ),  e = expert[row] (0: und weights, 1: gen weights)
// One warp per row, the row stays in registers between the reduction and the scale (single HBM pass).
// ---------------------------------------------------------------------------------------------
template <int VPL>  // 16-byte vectors per lane
__global__ void __launch_bounds__(128)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w0,
               const __nv_bfloat16* __restrict__ w1, const uint8_t* __restrict__ expert,
               __nv_bfloat16* __restrict__ y, long long ldy, int N, int H, float eps) {
  pdl_launch_dependents();   // a following PDL kernel (skinny GEMM) may begin prefetching its weights
  pdl_wait();                // no-op unless this kernel itself was launched with the PDL attribute (BAGEL_PDL_SMALL)
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int nvec = H >> 3;
  const bool gen = (expert != nullptr && w1 != nullptr) ? (expert[row] != 0) : false;   // in flight with the row loads
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
  // All loads of the row are issued back to back (select, not branch: with a guarded block per vector the compiler
  // serialises load -> use -> next load, 14 dependent L2 round trips = 14 us for a 32-row decode call), and the norm
  // weights are fetched before the reduction rather than after it.
  constexpr bool kPreW = VPL <= 16;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  uint4 v[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int idx = lane + 32 * i;
    v[i] = (idx < nvec) ? __ldg(xr + idx) : zero;
  }
  const uint4* wr = reinterpret_cast<const uint4*>(gen ? w1 : w0);
  uint4 wv[kPreW ? VPL : 1];
  if constexpr (kPreW) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int idx = lane + 32 * i;
      wv[i] = (idx < nvec) ? __ldg(wr + idx) : zero;
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = bf16_lo(u[e]), b = bf16_hi(u[e]);
      ss += a * a + b * b;
    }
  }
  ss = warp_sum(ss);
  const float r = rsqrtf(ss / (float)H + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + (long long)row * ldy);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      uint4 wq;
      if constexpr (kPreW) wq = wv[i];
      else wq = __ldg(wr + idx);
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      const uint32_t ww[4] = {wq.x, wq.y, wq.z, wq.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = bf16_round(bf16_lo(u[e]) * r), b = bf16_round(bf16_hi(u[e]) * r);
        o[e] = pack_bf16x2(bf16_lo(ww[e]) * a, bf16_hi(ww[e]) * b);
      }
      yr[idx] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm with affine (nn.LayerNorm in SiglipEncoderLayer / post_layernorm, modeling/bagel/siglip_navit.py:
// 269-271, 283, 294, 346, 370): y = bf16( (x - mean) * rsqrt(var + eps) * w + b ), biased variance, fp32 math.
// One warp per row, row held in registers (single HBM pass).
// ---------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(128)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                 const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, long long ldy, int N, int H,
                 float eps) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int nvec = H >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
  uint4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      v[i] = xr[idx];
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) s += bf16_lo(u[e]) + bf16_hi(u[e]);
    }
  }
  const float mean = warp_sum(s) / (float)H;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = bf16_lo(u[e]) - mean, c = bf16_hi(u[e]) - mean;
        ss += a * a + c * c;
      }
    }
  }
  const float r = rsqrtf(warp_sum(ss) / (float)H + eps);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  const uint4* br = reinterpret_cast<const uint4*>(b);
  uint4* yr = reinterpret_cast<uint4*>(y + (long long)row * ldy);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const uint4 wv = wr[idx], bv = br[idx];
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        o[e] = pack_bf16x2((bf16_lo(u[e]) - mean) * r * bf16_lo(ww[e]) + bf16_lo(bb[e]),
                           (bf16_hi(u[e]) - mean) * r * bf16_hi(ww[e]) + bf16_hi(bb[e]));
      yr[idx] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RoPE tables (Qwen2RotaryEmbedding.forward, modeling_qwen2.py:130-150): angle = pos * inv_freq in fp32,
// cos/sin optionally rounded to bf16 (the reference casts them to the hidden-stream dtype). Halves are
// duplicated in the reference; we store D/2 columns.
// ---------------------------------------------------------------------------------------------
__global__ void rope_table_kernel(const long long* __restrict__ pos, const float* __restrict__ inv_freq,
                                  float* __restrict__ cos_t, float* __restrict__ sin_t, int N, int half,
                                  int round_bf16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * half) return;
  const int r = (int)(i / half), c = (int)(i % half);
  const float ang = __fmul_rn((float)pos[r], inv_freq[c]);
  float cs = cosf(ang), sn = sinf(ang);
  if (round_bf16) { cs = bf16_round(cs); sn = bf16_round(sn); }
  cos_t[i] = cs;
  sin_t[i] = sn;
}

// ---------------------------------------------------------------------------------------------
// Per-head q/k RMSNorm + RoPE + bf16 cast + K/V placement into the merged KV buffer
// (PackedAttentionMoT.forward_inference, qwen2_navit.py:518-519 / 542-557 and the KV merge :559-574).
//   qkv   [N, (Hq+2Hk)*D]  bf16 output of the fused QKV projection (bias included)
//   q_out [N, Hq*D]; k_out/v_out [rows, Hk*D] written at row kv_rows[r] (the reference's packed_query_indexes)
// `flow` selects the reference's rounding points (SURVEY.md 8a dtype table):
//   0  mode A (bf16 weights), und / dense attention: every op rounds to bf16, bf16-rounded cos/sin
//   1  mode A, MoT gen branch: fp32 norm + RoPE (bf16 norm weights, bf16-rounded cos/sin), one final bf16 cast
//   2  mode B (fp32 weights), und / dense: bf16(x * r) * w_fp32 -> fp32, RoPE in fp32 with fp32 cos/sin
//   3  mode B, MoT gen branch: everything fp32 (fp32 norm weights, fp32 cos/sin)
// flows 2 and 3 read the norm weights as fp32. One warp per (row, head).
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128)
qk_norm_rope_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld_qkv, const void* __restrict__ qw0,
                    const void* __restrict__ kw0, const void* __restrict__ qw1,
                    const void* __restrict__ kw1, const uint8_t* __restrict__ expert,
                    const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                    __nv_bfloat16* __restrict__ q_out, long long ld_q, __nv_bfloat16* __restrict__ k_out,
                    __nv_bfloat16* __restrict__ v_out, long long ld_kv, const int* __restrict__ kv_rows, int N,
                    int Hq, int Hk, float eps, int fp32_flow) {
  pdl_launch_dependents();
  pdl_wait();   // no-op unless launched with the PDL attribute (BAGEL_PDL_SMALL)
  constexpr int E = D / 64;  // elements per lane in each half
  constexpr int HALF = D / 2;
  const int row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nheads = Hq + 2 * Hk;
  const bool gen = (expert != nullptr && qw1 != nullptr && expert[row]);
  const long long dst_row = kv_rows ? (long long)kv_rows[row] : (long long)row;
  const __nv_bfloat16* src_row = qkv + (long long)row * ld_qkv;

  float cs[E], sn[E];
#pragma unroll
  for (int t = 0; t < E; ++t) {
    cs[t] = cos_t[(long long)row * HALF + lane * E + t];
    sn[t] = sin_t[(long long)row * HALF + lane * E + t];
  }

  // gridDim.y > 1 (few rows, e.g. a decode step): one head per warp, so the load -> reduce -> store chain runs once per warp
  for (int hh = blockIdx.y * 4 + warp; hh < nheads; hh += 4 * gridDim.y) {
    const __nv_bfloat16* src = src_row + hh * D;
    float a[E], b[E];  // first-half / second-half elements owned by this lane
#pragma unroll
    for (int t = 0; t < E; ++t) {
      a[t] = __bfloat162float(src[lane * E + t]);
      b[t] = __bfloat162float(src[HALF + lane * E + t]);
    }
    __nv_bfloat16* dst;
    if (hh < Hq) dst = q_out + (long long)row * ld_q + hh * D;
    else if (hh < Hq + Hk) dst = k_out + dst_row * ld_kv + (hh - Hq) * D;
    else dst = v_out + dst_row * ld_kv + (hh - Hq - Hk) * D;

    if (hh >= Hq + Hk) {  // V: plain copy
#pragma unroll
      for (int t = 0; t < E; ++t) {
        dst[lane * E + t] = __float2bfloat16_rn(a[t]);
        dst[HALF + lane * E + t] = __float2bfloat16_rn(b[t]);
      }
      continue;
    }
    const void* w = (hh < Hq) ? (gen ? qw1 : qw0) : (gen ? kw1 : kw0);
    float ss = 0.f;
#pragma unroll
    for (int t = 0; t < E; ++t) ss += a[t] * a[t] + b[t] * b[t];
    ss = warp_sum(ss);
    const float r = rsqrtf(ss / (float)D + eps);
#pragma unroll
    for (int t = 0; t < E; ++t) {
      const float wa = fp32_flow >= 2 ? static_cast<const float*>(w)[lane * E + t]
                                      : __bfloat162float(static_cast<const __nv_bfloat16*>(w)[lane * E + t]);
      const float wb = fp32_flow >= 2 ? static_cast<const float*>(w)[HALF + lane * E + t]
                                      : __bfloat162float(static_cast<const __nv_bfloat16*>(w)[HALF + lane * E + t]);
      float ya, yb, oa, ob;
      if (fp32_flow) {
        const float na = (fp32_flow == 2) ? bf16_round(a[t] * r) : __fmul_rn(a[t], r);   // flow 2: q_norm runs on bf16 q
        const float nb = (fp32_flow == 2) ? bf16_round(b[t] * r) : __fmul_rn(b[t], r);
        ya = __fmul_rn(wa, na);
        yb = __fmul_rn(wb, nb);
        // q*cos + rotate_half(q)*sin, each product rounded separately (torch does not fuse)
        oa = __fadd_rn(__fmul_rn(ya, cs[t]), __fmul_rn(-yb, sn[t]));
        ob = __fadd_rn(__fmul_rn(yb, cs[t]), __fmul_rn(ya, sn[t]));
      } else {
        ya = bf16_round(wa * bf16_round(a[t] * r));
        yb = bf16_round(wb * bf16_round(b[t] * r));
        oa = bf16_round(ya * cs[t]) + bf16_round(-yb * sn[t]);
        ob = bf16_round(yb * cs[t]) + bf16_round(ya * sn[t]);
      }
      dst[lane * E + t] = __float2bfloat16_rn(oa);
      dst[HALF + lane * E + t] = __float2bfloat16_rn(ob);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row gather / scatter: dst[dst_rows[i] or i] = src[src_rows[i] or i]   (token embedding lookup bagel.py:277,796,
// modality gathers qwen2_navit.py:526-548, KV-cache placement :565-569). One warp per row, 16-byte vectors.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
copy_rows_kernel(const __nv_bfloat16* __restrict__ src, long long lds, const int* __restrict__ src_rows,
                 __nv_bfloat16* __restrict__ dst, long long ldd, const int* __restrict__ dst_rows, int M, int H) {
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= M) return;
  const long long sr = src_rows ? src_rows[i] : i;
  const long long dr = dst_rows ? dst_rows[i] : i;
  const uint4* s = reinterpret_cast<const uint4*>(src + sr * lds);
  uint4* d = reinterpret_cast<uint4*>(dst + dr * ldd);
  for (int v = lane; v < (H >> 3); v += 32) d[v] = s[v];
}

// ---------------------------------------------------------------------------------------------
// Latent-in tail (bagel.py:801-806): seq[dst_rows[i]] = bf16( bf16(proj[i] + t_emb) + pos_table[pos_ids[i]] )
// where proj = vae2llm(x_t) (GEMM), t_emb = TimestepEmbedder(t) (one row, identical for every latent token).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
latent_embed_add_kernel(const __nv_bfloat16* __restrict__ proj, long long ldp, const __nv_bfloat16* __restrict__ t_emb,
                        const __nv_bfloat16* __restrict__ pos_table, long long ldt,
                        const long long* __restrict__ pos_ids, __nv_bfloat16* __restrict__ seq, long long lds,
                        const int* __restrict__ dst_rows, int M, int H) {
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= M) return;
  const uint4* pr = reinterpret_cast<const uint4*>(proj + (long long)i * ldp);
  const uint4* te = reinterpret_cast<const uint4*>(t_emb);
  const uint4* pt = reinterpret_cast<const uint4*>(pos_table + pos_ids[i] * ldt);
  uint4* d = reinterpret_cast<uint4*>(seq + (long long)(dst_rows ? dst_rows[i] : i) * lds);
  for (int v = lane; v < (H >> 3); v += 32) {
    const uint4 a = pr[v], b = (t_emb != nullptr) ? te[v] : make_uint4(0u, 0u, 0u, 0u), c = pt[v];
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w}, uc[4] = {c.x, c.y, c.z, c.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float lo = bf16_round(bf16_lo(ua[e]) + bf16_lo(ub[e])) + bf16_lo(uc[e]);
      const float hi = bf16_round(bf16_hi(ua[e]) + bf16_hi(ub[e])) + bf16_hi(uc[e]);
      o[e] = pack_bf16x2(lo, hi);
    }
    d[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// Classifier-free guidance + renorm + Euler step (bagel.py:873-907 and :746), all bf16 rounding points kept:
//   u = vT + sT (v - vT);  w = (sI > 1) ? vI + sI (u - vI) : u;
//   scale = clamp(|v| / (|w| + 1e-8), renorm_min, 1)  over the whole batch ("global", type 0) or per token
//   ("channel", type 1);  "text_channel" (type 2) renormalises u per token against v before the image CFG.
//   x <- x - bf16(bf16(w * scale) * dt)            x fp32 [M, C]
// Pass 1 (global only) accumulates sum v^2 and sum w^2 into norms[0..1]; pass 2 applies.
// v, vT, vI are rows of the llm2vae output selected by `rows` (the latent rows of the packed sequence).
// ---------------------------------------------------------------------------------------------
struct CfgArgs {
  const __nv_bfloat16 *v, *vT, *vI;
  long long ldv;
  const int* rows;
  float* x;
  float* norms;  // [2] fp32, zeroed by the caller before pass 1
  int M, C;
  float sT, sI, renorm_min, dt;
  const float* dt_dev;  // optional: read dt from device memory (lets one CUDA graph serve every step)
  int renorm_type;
};

__device__ __forceinline__ void cfg_combine(const CfgArgs& a, float v, float vT, float vI, float& u, float& w) {
  // bf16 tensor arithmetic with python-float scalars: every op rounds to bf16
  u = bf16_round(vT + bf16_round(a.sT * bf16_round(v - vT)));
  w = (a.sI > 1.0f) ? bf16_round(vI + bf16_round(a.sI * bf16_round(u - vI))) : u;
}

__global__ void __launch_bounds__(256) cfg_norm_kernel(const CfgArgs a) {
  float sv = 0.f, sw = 0.f;
  const long long total = (long long)a.M * a.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / a.C), c = (int)(i % a.C);
    const long long src = (long long)(a.rows ? a.rows[r] : r) * a.ldv + c;
    const float v = __bfloat162float(a.v[src]);
    const float vT = __bfloat162float(a.vT[src]);
    const float vI = a.vI ? __bfloat162float(a.vI[src]) : 0.f;
    float u, w;
    cfg_combine(a, v, vT, vI, u, w);
    sv += v * v;
    sw += w * w;
  }
  sv = warp_sum(sv);
  sw = warp_sum(sw);
  __shared__ float sh[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = sv; sh[1][warp] = sw; }
  __syncthreads();
  if (warp == 0) {
    sv = lane < 8 ? sh[0][lane] : 0.f;
    sw = lane < 8 ? sh[1][lane] : 0.f;
    sv = warp_sum(sv);
    sw = warp_sum(sw);
    if (lane == 0) { atomicAdd(&a.norms[0], sv); atomicAdd(&a.norms[1], sw); }
  }
}

// one warp per latent token (C = 64 channels -> 2 per lane)
__global__ void __launch_bounds__(128) cfg_apply_kernel(const CfgArgs a, int use_cfg) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= a.M) return;
  const long long src = (long long)(a.rows ? a.rows[r] : r) * a.ldv;
  float gscale = 1.f;
  if (use_cfg && a.renorm_type == 0) {
    const float nv = bf16_round(sqrtf(a.norms[0])), nw = bf16_round(sqrtf(a.norms[1]));
    gscale = fminf(fmaxf(bf16_round(nv / bf16_round(nw + 1e-8f)), a.renorm_min), 1.0f);
  }
  float vv[4], ww[4];
  int cnt = 0;
  float sv = 0.f, sw = 0.f;
  for (int c = lane; c < a.C; c += 32, ++cnt) {
    const float v = __bfloat162float(a.v[src + c]);
    float w = v;
    if (use_cfg) {
      const float vT = __bfloat162float(a.vT[src + c]);
      const float vI = a.vI ? __bfloat162float(a.vI[src + c]) : 0.f;
      if (a.renorm_type == 2) {
        w = bf16_round(vT + bf16_round(a.sT * bf16_round(v - vT)));  // u; image CFG applied after renorm
      } else {
        float u;
        cfg_combine(a, v, vT, vI, u, w);
      }
    }
    vv[cnt] = v;
    ww[cnt] = w;
    sv += v * v;
    sw += w * w;
  }
  const float dt = a.dt_dev ? *a.dt_dev : a.dt;
  float scale = gscale;
  if (use_cfg && a.renorm_type != 0) {
    sv = warp_sum(sv);
    sw = warp_sum(sw);
    const float nv = bf16_round(sqrtf(sv)), nw = bf16_round(sqrtf(sw));
    scale = fminf(fmaxf(bf16_round(nv / bf16_round(nw + 1e-8f)), a.renorm_min), 1.0f);
  }
  cnt = 0;
  for (int c = lane; c < a.C; c += 32, ++cnt) {
    float w = ww[cnt];
    if (use_cfg) {
      w = bf16_round(w * scale);
      if (a.renorm_type == 2 && a.sI > 1.0f) {
        const float vI = __bfloat162float(a.vI[src + c]);
        w = bf16_round(vI + bf16_round(a.sI * bf16_round(w - vI)));
      }
    }
    float* xp = a.x + (long long)r * a.C + c;
    *xp = *xp - bf16_round(w * dt);
  }
}

// x fp32 [M, C] -> bf16 (the autocast cast in front of vae2llm)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    *reinterpret_cast<uint32_t*>(y + i) = pack_bf16x2(x[i], x[i + 1]);
  } else if (i < n) {
    y[i] = __float2bfloat16_rn(x[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// text decode bookkeeping (bagel.py:945-992) kept on the device
// ---------------------------------------------------------------------------------------------
__global__ void decode_prepare_kernel(const int* __restrict__ k_begin, const int* __restrict__ seq_len,
                                      int* __restrict__ kv_rows, int* __restrict__ seqused, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    kv_rows[b] = k_begin[b] + seq_len[b];
    seqused[b] = seq_len[b] + 1;
  }
}

__global__ void decode_advance_kernel(int* __restrict__ seq_len, long long* __restrict__ pos,
                                      const long long* __restrict__ tokens, long long* __restrict__ history,
                                      int* __restrict__ step_dev, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int step = step_dev[0];
  if (b < B) {
    seq_len[b] += 1;
    pos[b] += 1;
    history[(long long)step * B + b] = tokens[b];
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) step_dev[0] = step + 1;  // single block launch (B <= 1024)
}

// one block per row: first index of the maximum (torch.argmax tie rule). The row (vocab 152064 = 297 KB) was just
// written by the lm_head GEMM and is L2-resident; 1024 threads x 16-byte loads keep ~16 KB per iteration in flight.
constexpr int kArgmaxThreads = 1024;
__global__ void __launch_bounds__(kArgmaxThreads)
argmax_rows_kernel(const __nv_bfloat16* __restrict__ logits, long long ld, int V, long long* __restrict__ tokens,
                   int* __restrict__ tokens32) {
  const __nv_bfloat16* row = logits + (long long)blockIdx.x * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  auto take = [&](float v, int i) {
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  };
  int vec_end = 0;
  if ((ld % 8) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) {
    vec_end = V & ~7;
    for (int i = threadIdx.x * 8; i < vec_end; i += kArgmaxThreads * 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(row + i);
      take(bf16_lo(u.x), i);     take(bf16_hi(u.x), i + 1);
      take(bf16_lo(u.y), i + 2); take(bf16_hi(u.y), i + 3);
      take(bf16_lo(u.z), i + 4); take(bf16_hi(u.z), i + 5);
      take(bf16_lo(u.w), i + 6); take(bf16_hi(u.w), i + 7);
    }
  }
  for (int i = vec_end + threadIdx.x; i < V; i += kArgmaxThreads) take(__bfloat162float(row[i]), i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  __shared__ float sv[kArgmaxThreads / 32];
  __shared__ int si[kArgmaxThreads / 32];
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = sv[threadIdx.x];
    bi = si[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) {
      if (bi == 0x7fffffff) bi = 0;   // all-NaN / all -inf row: torch returns an in-range index
      tokens[blockIdx.x] = bi;
      if (tokens32) tokens32[blockIdx.x] = bi;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// TaylorSeer step cache (reference modeling/cache_utils/taylorseer.py; hooks qwen2_navit.py:773-777, 824-829).
// Factors are bf16 planes [order][rows][H] of the LAST decoder layer's output (on an extrapolated step each layer's
// output replaces its input, so only the last layer's extrapolation reaches the result). Both kernels are pure HBM
// streams over [rows, H] and reproduce torch's bf16 elementwise rounding: every binary op rounds to bf16, python
// scalars enter as fp32.
// ---------------------------------------------------------------------------------------------
constexpr int kTaylorMaxPlanes = 7;   // max_order 6 -> factors 0..6

// derivative_approximation (:12-32): new[0] = feature; new[i+1] = bf16(bf16(new[i] - old[i]) / dist), i < n_deriv.
// In place: the old factors of an element are read before any plane is rewritten.
__global__ void __launch_bounds__(256)
taylor_update_kernel(const __nv_bfloat16* __restrict__ feature, long long ldf, __nv_bfloat16* __restrict__ factors,
                     long long plane, int n_deriv, float dist, int rows, int vec_per_row) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * vec_per_row) return;
  const int row = (int)(idx / vec_per_row), c = (int)(idx - (long long)row * vec_per_row);
  const uint4 f = *reinterpret_cast<const uint4*>(feature + (long long)row * ldf + c * 8);
  __nv_bfloat16* dst = factors + ((long long)row * vec_per_row + c) * 8;
  uint4 old[kTaylorMaxPlanes - 1];
#pragma unroll
  for (int i = 0; i < kTaylorMaxPlanes - 1; ++i)
    if (i < n_deriv) old[i] = *reinterpret_cast<const uint4*>(dst + i * plane);
  uint32_t cur[4] = {f.x, f.y, f.z, f.w};
  *reinterpret_cast<uint4*>(dst) = f;
#pragma unroll
  for (int i = 0; i < kTaylorMaxPlanes - 1; ++i) {
    if (i < n_deriv) {
      const uint32_t o[4] = {old[i].x, old[i].y, old[i].z, old[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d0 = bf16_round(bf16_lo(cur[e]) - bf16_lo(o[e]));
        const float d1 = bf16_round(bf16_hi(cur[e]) - bf16_hi(o[e]));
        cur[e] = pack_bf16x2(__fdiv_rn(d0, dist), __fdiv_rn(d1, dist));
      }
      *reinterpret_cast<uint4*>(dst + (long long)(i + 1) * plane) = make_uint4(cur[0], cur[1], cur[2], cur[3]);
    }
  }
}

struct TaylorCoef {
  float c[kTaylorMaxPlanes];    // 1 / i!
  float xp[kTaylorMaxPlanes];   // x^i
};

// taylor_formula (:34-47): out = sum_i bf16(bf16(c_i * f_i) * x^i), accumulated in bf16 in order i = 0, 1, ...
__global__ void __launch_bounds__(256)
taylor_eval_kernel(const __nv_bfloat16* __restrict__ factors, long long plane, int n, const TaylorCoef k,
                   __nv_bfloat16* __restrict__ out, long long ldo, int rows, int vec_per_row) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * vec_per_row) return;
  const int row = (int)(idx / vec_per_row), c = (int)(idx - (long long)row * vec_per_row);
  const __nv_bfloat16* src = factors + ((long long)row * vec_per_row + c) * 8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < kTaylorMaxPlanes; ++i) {
    if (i < n) {
      const uint4 f = *reinterpret_cast<const uint4*>(src + i * plane);
      const uint32_t u[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float t0 = bf16_round(bf16_round(bf16_lo(u[e]) * k.c[i]) * k.xp[i]);
        const float t1 = bf16_round(bf16_round(bf16_hi(u[e]) * k.c[i]) * k.xp[i]);
        acc[2 * e] = (i == 0) ? t0 : bf16_round(acc[2 * e] + t0);
        acc[2 * e + 1] = (i == 0) ? t1 : bf16_round(acc[2 * e + 1] + t1);
      }
    }
  }
  *reinterpret_cast<uint4*>(out + (long long)row * ldo + c * 8) =
      make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                 pack_bf16x2(acc[6], acc[7]));
}

}  // namespace bagel

using namespace bagel;

#define COUNT_LAUNCH() g_launches.fetch_add(1, std::memory_order_relaxed)

extern "C" int bagel_rmsnorm_bf16(const void* x, long long ldx, const void* w0, const void* w1, const uint8_t* expert,
                                  void* y, long long ldy, int N, int H, float eps, void* stream) {
  if (N <= 0) return 0;
  if ((H % 8) || (ldx % 8) || (ldy % 8)) return set_error(BAGEL_ERR_ALIGN, "bagel_rmsnorm_bf16: H, ldx, ldy must be multiples of 8");
  const int nvec = H / 8;
  const int vpl = (nvec + 31) / 32;
  dim3 grid((N + 3) / 4), block(128);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto W0 = static_cast<const __nv_bfloat16*>(w0);
  auto W1 = static_cast<const __nv_bfloat16*>(w1);
  auto Y = static_cast<__nv_bfloat16*>(y);
  const bool pdl = pdl_small_enabled() && N <= 64;   // decode-sized calls only
#define RMS_CASE(V) BAGEL_CUDA_CHECK(launch_maybe_pdl(rmsnorm_kernel<V>, grid, block, 0, s, pdl, X, ldx, W0, W1, expert, Y, ldy, N, H, eps))
  if (vpl <= 1) RMS_CASE(1);
  else if (vpl <= 2) RMS_CASE(2);
  else if (vpl <= 4) RMS_CASE(4);
  else if (vpl <= 8) RMS_CASE(8);
  else if (vpl <= 14) RMS_CASE(14);
  else if (vpl <= 32) RMS_CASE(32);
  else return set_error(BAGEL_ERR_SHAPE, "bagel_rmsnorm_bf16: H=%d too large (max 8192)", H);
#undef RMS_CASE
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_layernorm_bf16(const void* x, long long ldx, const void* w, const void* b, void* y,
                                    long long ldy, int N, int H, float eps, void* stream) {
  if (N <= 0) return 0;
  if ((H % 8) || (ldx % 8) || (ldy % 8)) return set_error(BAGEL_ERR_ALIGN, "bagel_layernorm_bf16: H, ldx, ldy must be multiples of 8");
  const int vpl = (H / 8 + 31) / 32;
  dim3 grid((N + 3) / 4), block(128);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto W = static_cast<const __nv_bfloat16*>(w);
  auto B = static_cast<const __nv_bfloat16*>(b);
  auto Y = static_cast<__nv_bfloat16*>(y);
#define LN_CASE(V) layernorm_kernel<V><<<grid, block, 0, s>>>(X, ldx, W, B, Y, ldy, N, H, eps)
  if (vpl <= 1) LN_CASE(1);
  else if (vpl <= 2) LN_CASE(2);
  else if (vpl <= 5) LN_CASE(5);
  else if (vpl <= 8) LN_CASE(8);
  else if (vpl <= 16) LN_CASE(16);
  else return set_error(BAGEL_ERR_SHAPE, "bagel_layernorm_bf16: H=%d too large (max 4096)", H);
#undef LN_CASE
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_rope_table(const long long* pos, const float* inv_freq, float* cos_t, float* sin_t, int N,
                                int half, int round_bf16, void* stream) {
  if (N <= 0) return 0;
  const long long total = (long long)N * half;
  rope_table_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pos, inv_freq, cos_t, sin_t, N, half, round_bf16);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_qk_norm_rope(const void* qkv, long long ld_qkv, const void* q_w0, const void* k_w0,
                                  const void* q_w1, const void* k_w1, const uint8_t* expert, const float* cos_t,
                                  const float* sin_t, void* q_out, long long ld_q, void* k_out, void* v_out,
                                  long long ld_kv, const int* kv_rows, int N, int Hq, int Hk, int D, float eps,
                                  int fp32_flow, void* stream) {
  if (N <= 0) return 0;
  if (D != 64 && D != 128) return set_error(BAGEL_ERR_SHAPE, "bagel_qk_norm_rope: head_dim must be 64 or 128");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (fp32_flow < 0 || fp32_flow > 3) return set_error(BAGEL_ERR_ARG, "bagel_qk_norm_rope: flow must be 0..3");
#define QK_ARGS                                                                                                   \
  static_cast<const __nv_bfloat16*>(qkv), ld_qkv, q_w0, k_w0, q_w1, k_w1, expert, cos_t, sin_t,                   \
      static_cast<__nv_bfloat16*>(q_out), ld_q, static_cast<__nv_bfloat16*>(k_out),                               \
      static_cast<__nv_bfloat16*>(v_out), ld_kv, kv_rows, N, Hq, Hk, eps, fp32_flow
  // few rows (decode): spread the heads of a row over blockIdx.y — the per-head chain (load, warp reduce, store) is pure latency
  static const bool spread = [] { const char* e = getenv("BAGEL_QKROPE_SPREAD"); return !(e && atoi(e) == 0); }();
  const dim3 grid((unsigned)N, (spread && N <= 1024) ? (unsigned)((Hq + 2 * Hk + 3) / 4) : 1u);
  const bool pdl = pdl_small_enabled() && N <= 64;
  if (D == 128) BAGEL_CUDA_CHECK(launch_maybe_pdl(qk_norm_rope_kernel<128>, grid, dim3(128), 0, s, pdl, QK_ARGS));
  else BAGEL_CUDA_CHECK(launch_maybe_pdl(qk_norm_rope_kernel<64>, grid, dim3(128), 0, s, pdl, QK_ARGS));
#undef QK_ARGS
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_copy_rows_bf16(const void* src, long long lds, const int* src_rows, void* dst, long long ldd,
                                    const int* dst_rows, int M, int H, void* stream) {
  if (M <= 0) return 0;
  if ((H % 8) || (lds % 8) || (ldd % 8)) return set_error(BAGEL_ERR_ALIGN, "bagel_copy_rows_bf16: H, lds, ldd %% 8");
  copy_rows_kernel<<<(M + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), lds, src_rows, static_cast<__nv_bfloat16*>(dst), ldd, dst_rows, M, H);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_latent_embed_add(const void* proj, long long ldp, const void* t_emb, const void* pos_table,
                                      long long ldt, const long long* pos_ids, void* seq, long long lds,
                                      const int* dst_rows, int M, int H, void* stream) {
  if (M <= 0) return 0;
  if ((H % 8) || (ldp % 8) || (ldt % 8) || (lds % 8)) return set_error(BAGEL_ERR_ALIGN, "bagel_latent_embed_add: alignment");
  latent_embed_add_kernel<<<(M + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(proj), ldp, static_cast<const __nv_bfloat16*>(t_emb),
      static_cast<const __nv_bfloat16*>(pos_table), ldt, pos_ids, static_cast<__nv_bfloat16*>(seq), lds, dst_rows, M, H);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_cfg_euler_step(const void* v, const void* v_text, const void* v_img, long long ldv,
                                    const int* rows, float* x, float* norms_ws, int M, int C, float cfg_text_scale,
                                    float cfg_img_scale, float renorm_min, int renorm_type, float dt, const float* dt_dev,
                                    void* stream) {
  if (M <= 0) return 0;
  if (C > 128) return set_error(BAGEL_ERR_SHAPE, "bagel_cfg_euler_step: C must be <= 128");
  if (renorm_type < 0 || renorm_type > 2) return set_error(BAGEL_ERR_ARG, "bagel_cfg_euler_step: renorm_type in {0,1,2}");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CfgArgs a{};
  a.v = static_cast<const __nv_bfloat16*>(v);
  a.vT = static_cast<const __nv_bfloat16*>(v_text);
  a.vI = static_cast<const __nv_bfloat16*>(v_img);
  a.ldv = ldv; a.rows = rows; a.x = x; a.norms = norms_ws; a.M = M; a.C = C;
  a.sT = cfg_text_scale; a.sI = cfg_img_scale; a.renorm_min = renorm_min; a.dt = dt; a.dt_dev = dt_dev;
  a.renorm_type = renorm_type;
  const int use_cfg = (cfg_text_scale > 1.0f && v_text != nullptr) ? 1 : 0;
  if (use_cfg && a.sI > 1.0f && a.vI == nullptr) return set_error(BAGEL_ERR_ARG, "bagel_cfg_euler_step: cfg_img_scale > 1 needs v_img");
  if (use_cfg && renorm_type == 0) {
    if (norms_ws == nullptr) return set_error(BAGEL_ERR_ARG, "bagel_cfg_euler_step: global renorm needs norms_ws[2]");
    BAGEL_CUDA_CHECK(cudaMemsetAsync(norms_ws, 0, 2 * sizeof(float), s));
    const long long total = (long long)M * C;
    int blocks = (int)((total + 256 * 8 - 1) / (256 * 8));
    if (blocks > 1184) blocks = 1184;
    if (blocks < 1) blocks = 1;
    cfg_norm_kernel<<<blocks, 256, 0, s>>>(a);
    COUNT_LAUNCH();
  }
  cfg_apply_kernel<<<(M + 3) / 4, 128, 0, s>>>(a, use_cfg);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream) {
  if (n <= 0) return 0;
  cast_f32_bf16_kernel<<<(unsigned)((n / 2 + 256) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), n);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}


extern "C" int bagel_decode_prepare(const int* k_begin, const int* seq_len, int* kv_rows, int* seqused, int B,
                                    void* stream) {
  if (B <= 0) return 0;
  decode_prepare_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(k_begin, seq_len, kv_rows, seqused, B);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_argmax_rows_bf16(const void* logits, long long ld, int B, int V, long long* tokens, int* tokens32,
                                      void* stream) {
  if (B <= 0 || V <= 0) return 0;
  argmax_rows_kernel<<<B, kArgmaxThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(logits), ld, V,
                                                                       tokens, tokens32);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_decode_advance(int* seq_len, long long* pos, const long long* tokens, long long* history,
                                    int* step_dev, int B, void* stream) {
  if (B <= 0) return 0;
  if (B > 1024) return set_error(BAGEL_ERR_SHAPE, "bagel_decode_advance: B must be <= 1024");
  decode_advance_kernel<<<1, ((B + 31) / 32) * 32, 0, static_cast<cudaStream_t>(stream)>>>(seq_len, pos, tokens, history,
                                                                                          step_dev, B);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_taylor_update_bf16(const void* feature, long long ldf, void* factors, long long plane_stride,
                                        int n_deriv, int dist, int rows, int H, void* stream) {
  if (rows <= 0) return 0;
  if ((H % 8) || (ldf % 8) || (plane_stride % 8) || plane_stride < (long long)rows * H)
    return set_error(BAGEL_ERR_ALIGN, "bagel_taylor_update_bf16: H, ldf, plane_stride %% 8 and plane_stride >= rows*H required");
  if (n_deriv < 0 || n_deriv > kTaylorMaxPlanes - 1 || (n_deriv > 0 && dist <= 0))
    return set_error(BAGEL_ERR_ARG, "bagel_taylor_update_bf16: n_deriv must be in [0, 6] with dist > 0");
  const long long n = (long long)rows * (H / 8);
  taylor_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(feature), ldf, static_cast<__nv_bfloat16*>(factors), plane_stride, n_deriv,
      (float)dist, rows, H / 8);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_taylor_eval_bf16(const void* factors, long long plane_stride, int n_factors, int x, void* out,
                                      long long ldo, int rows, int H, void* stream) {
  if (rows <= 0) return 0;
  if ((H % 8) || (ldo % 8) || (plane_stride % 8) || plane_stride < (long long)rows * H)
    return set_error(BAGEL_ERR_ALIGN, "bagel_taylor_eval_bf16: H, ldo, plane_stride %% 8 and plane_stride >= rows*H required");
  if (n_factors < 1 || n_factors > kTaylorMaxPlanes)
    return set_error(BAGEL_ERR_ARG, "bagel_taylor_eval_bf16: n_factors must be in [1, 7]");
  TaylorCoef k{};
  double fact = 1.0, xp = 1.0;
  for (int i = 0; i < kTaylorMaxPlanes; ++i) {
    if (i > 0) { fact *= i; xp *= x; }
    k.c[i] = (float)(1.0 / fact);
    k.xp[i] = (float)xp;
  }
  const long long n = (long long)rows * (H / 8);
  taylor_eval_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(factors), plane_stride, n_factors, k, static_cast<__nv_bfloat16*>(out), ldo, rows,
      H / 8);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// SigLIP 2-D RoPE (modeling/bagel/siglip_navit.py:102-142, 224-230): the first half of every q/k head is rotated by the
// ROW table, the second half by the COLUMN table of the patch position (RotaryEmbedding2D buffers, fp32), in place on
// the q and k heads of the fused QKV buffer. The reference computes q*cos + rotate_half(q)*sin with bf16 q and fp32
// tables => fp32 products and sum (each rounded separately), one final cast to bf16.
// ---------------------------------------------------------------------------------------------------------------------
namespace bagel {
__global__ void siglip_rope2d_kernel(__nv_bfloat16* __restrict__ x, long long ld, int n, int heads, int head_stride, int d,
                                     const long long* __restrict__ pos, const float* __restrict__ cos_h,
                                     const float* __restrict__ sin_h, const float* __restrict__ cos_w,
                                     const float* __restrict__ sin_w) {
  const int half = d / 2, quarter = d / 4;
  const long long total = (long long)n * heads * 2 * quarter;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % quarter);
  long long t = idx / quarter;
  const int hw = (int)(t & 1);
  t >>= 1;
  const int head = (int)(t % heads);
  const long long tok = t / heads;
  __nv_bfloat16* p = x + tok * ld + (long long)head * head_stride + hw * half;
  const long long row = pos[tok] * half;
  const float* ct = (hw ? cos_w : cos_h) + row;
  const float* st = (hw ? sin_w : sin_h) + row;
  const float a = __bfloat162float(p[j]), b = __bfloat162float(p[j + quarter]);
  const float oa = __fadd_rn(__fmul_rn(a, ct[j]), __fmul_rn(-b, st[j]));
  const float ob = __fadd_rn(__fmul_rn(b, ct[j + quarter]), __fmul_rn(a, st[j + quarter]));
  p[j] = __float2bfloat16_rn(oa);
  p[j + quarter] = __float2bfloat16_rn(ob);
}
}  // namespace bagel

extern "C" int bagel_siglip_rope2d_bf16(void* x, long long ld, int n_tokens, int heads, int head_stride, int head_dim,
                                        const long long* pos_ids, const float* cos_h, const float* sin_h,
                                        const float* cos_w, const float* sin_w, void* stream) {
  if (n_tokens <= 0 || heads <= 0) return 0;
  if (head_dim <= 0 || (head_dim % 4) || head_stride < head_dim)
    return bagel::set_error(BAGEL_ERR_SHAPE, "bagel_siglip_rope2d_bf16: head_dim must be a positive multiple of 4 and <= head_stride");
  if (!x || !pos_ids || !cos_h || !sin_h || !cos_w || !sin_w)
    return bagel::set_error(BAGEL_ERR_ARG, "bagel_siglip_rope2d_bf16: null pointer");
  const long long total = (long long)n_tokens * heads * (head_dim / 2);
  bagel::siglip_rope2d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(x), ld, n_tokens, heads, head_stride, head_dim, pos_ids, cos_h, sin_h, cos_w, sin_w);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------------
// dtype mode B (fp32 master weights under autocast: eval/gen/gen_images_mp.py:159-175 + :73; SURVEY.md 8a dtype table):
// the residual stream and the norm outputs are fp32, every nn.Linear still runs bf16 x bf16 -> bf16.
// ---------------------------------------------------------------------------------------------------------------------
namespace bagel {

// Qwen2RMSNorm on an fp32 row with fp32 weights: y = w * (x * rsqrt(mean(x^2) + eps)), both products rounded to fp32
// (modeling_qwen2.py:54-59 with input_dtype = fp32). OUT = float keeps that value (final norm returned to the caller);
// OUT = bf16 adds the autocast cast in front of the next nn.Linear. Block per row, second pass re-reads x from L1/L2.
template <typename OUT>
__global__ void __launch_bounds__(256)
rmsnorm_f32_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w0, const float* __restrict__ w1,
                   const uint8_t* __restrict__ expert, OUT* __restrict__ y, long long ldy, int H, float eps) {
  const long long row = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  const int nvec = H >> 2;
  float ss = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const float4 v = xr[i];
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += sh[i];
  const float r = rsqrtf(tot / (float)H + eps);
  const float* w = (expert != nullptr && w1 != nullptr && expert[row]) ? w1 : w0;
  const float4* wr = reinterpret_cast<const float4*>(w);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const float4 v = xr[i], ww = wr[i];
    const float o0 = __fmul_rn(ww.x, __fmul_rn(v.x, r)), o1 = __fmul_rn(ww.y, __fmul_rn(v.y, r));
    const float o2 = __fmul_rn(ww.z, __fmul_rn(v.z, r)), o3 = __fmul_rn(ww.w, __fmul_rn(v.w, r));
    if constexpr (sizeof(OUT) == 4) {
      reinterpret_cast<float4*>(y + row * ldy)[i] = make_float4(o0, o1, o2, o3);
    } else {
      reinterpret_cast<uint2*>(y + row * ldy)[i] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
    }
  }
}

// latent-in tail with an fp32 hidden stream (mode B): seq32[dst] = fp32( bf16(proj + t_emb) + pos_table_fp32[pos] ).
__global__ void __launch_bounds__(128)
latent_embed_add_f32_kernel(const __nv_bfloat16* __restrict__ proj, long long ldp, const __nv_bfloat16* __restrict__ t_emb,
                            const float* __restrict__ pos_table, long long ldt, const long long* __restrict__ pos_ids,
                            float* __restrict__ seq, long long lds, const int* __restrict__ dst_rows, int M, int H) {
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= M) return;
  const __nv_bfloat16* pr = proj + (long long)i * ldp;
  const float* pt = pos_table + pos_ids[i] * ldt;
  float* d = seq + (long long)(dst_rows ? dst_rows[i] : i) * lds;
  for (int c = lane; c < H; c += 32) {
    const float a = __bfloat162float(pr[c]);
    const float s = t_emb ? bf16_round(a + __bfloat162float(t_emb[c])) : a;
    d[c] = __fadd_rn(s, pt[c]);
  }
}
}  // namespace bagel

extern "C" int bagel_rmsnorm_f32(const float* x, long long ldx, const float* w0, const float* w1, const uint8_t* expert,
                                 void* y, long long ldy, int out_f32, int N, int H, float eps, void* stream) {
  if (N <= 0) return 0;
  if ((H % 4) || (ldx % 4) || (ldy % 4)) return bagel::set_error(BAGEL_ERR_ALIGN, "bagel_rmsnorm_f32: H, ldx, ldy must be multiples of 4");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (out_f32) bagel::rmsnorm_f32_kernel<float><<<N, 256, 0, s>>>(x, ldx, w0, w1, expert, static_cast<float*>(y), ldy, H, eps);
  else bagel::rmsnorm_f32_kernel<__nv_bfloat16><<<N, 256, 0, s>>>(x, ldx, w0, w1, expert, static_cast<__nv_bfloat16*>(y), ldy, H, eps);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_latent_embed_add_f32(const void* proj, long long ldp, const void* t_emb, const float* pos_table,
                                          long long ldt, const long long* pos_ids, float* seq, long long lds,
                                          const int* dst_rows, int M, int H, void* stream) {
  if (M <= 0) return 0;
  bagel::latent_embed_add_f32_kernel<<<(M + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(proj), ldp, static_cast<const __nv_bfloat16*>(t_emb), pos_table, ldt, pos_ids, seq,
      lds, dst_rows, M, H);
  COUNT_LAUNCH();
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
