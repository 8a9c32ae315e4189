// Launcher of the double-buffered-S attention kernel (attn3.cu); called from bagel_attn_varlen_fwd (attn.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace bagel {

bool attn3_enabled();   // BAGEL_ATTN_V3 (read once)
// Tensor maps as built by bagel_attn_varlen_fwd: Q / K / V boxes of [128 rows, 64 columns], 128-byte swizzle.
int attn3_varlen(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, void* out, long long ld_out,
                 const int* cu_q, const int* cu_k, const int* seqused_k, int batch, int Hq, int Hk, int head_dim,
                 int max_seqlen_q, int causal, float scale_log2, cudaStream_t stream);

}  // namespace bagel
