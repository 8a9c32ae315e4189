// Single-query (decode) attention over an appended K/V cache; see attn_decode.cu. Called from bagel_attn_varlen_fwd
// when max_seqlen_q == 1.
#pragma once
#include <cuda_runtime.h>

namespace bagel {

bool attn_decode_supported(int max_seqlen_q, int head_dim, int Hq, int Hk);
int attn_decode(const void* q, const void* k, const void* v, void* out, const int* cu_seqlens_q,
                const int* cu_seqlens_k, const int* seqused_k, int batch, int Hq, int Hk, int max_seqlen_k,
                float softmax_scale, long long ld_q, long long ld_k, long long ld_v, long long ld_out,
                cudaStream_t stream);

}  // namespace bagel
