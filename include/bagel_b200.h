/* bagel_b200 — C ABI of the B200-native (sm_100a) kernels behind BAGEL's inference forward path.
 *
 * The reference (ByteDance-Seed/Bagel) is pure Python; the only native seam it has is
 * `flash_attn_varlen_func` (modeling/bagel/qwen2_navit.py:361,579; modeling/bagel/siglip_navit.py:232).
 * Every other GPU op is reached through torch (nn.Linear -> cuBLASLt, ATen elementwise, cuDNN conv).
 * This header is the boundary a maintainer would bind with ctypes: plain pointers
 * and sizes, no torch types. Conventions for every entry point:
 *
 *   - all pointers are DEVICE pointers unless a parameter is documented as host;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library never allocates device memory;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no internal synchronisation, so calls
 *     are CUDA-graph capturable;
 *   - return value 0 on success, a negative BAGEL_ERR_* otherwise; `bagel_last_error()` gives the message
 *     (thread-local);
 *   - bf16 tensors are row-major with an explicit leading dimension in ELEMENTS.
 */
#ifndef BAGEL_B200_H_
#define BAGEL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAGEL_OK 0
#define BAGEL_ERR_SHAPE (-1)
#define BAGEL_ERR_ALIGN (-2)
#define BAGEL_ERR_ARCH (-3)
#define BAGEL_ERR_CUDA (-4)
#define BAGEL_ERR_ARG (-5)

const char* bagel_last_error(void);
/* ABI version of this header; bump on any signature change. */
int bagel_abi_version(void);
/* Number of kernels launched by this library in this process (monotonic; used by bench.py's gpu_launches). */
long long bagel_launch_count(void);

/* GEMM epilogues (bagel_gemm_bf16 `epilogue`). */
#define BAGEL_EPI_BIAS 0   /* C = bf16(acc + bias)                                   nn.Linear               */
#define BAGEL_EPI_RESID 1  /* C = bf16(resid + bf16(acc + bias))                     Linear + residual add   */
#define BAGEL_EPI_SWIGLU 2 /* C[:,j] = bf16(bf16(silu(bf16 g_j)) * bf16 u_j)         Qwen2MLP gate/up/act    */
#define BAGEL_EPI_GELU 3   /* C = bf16(gelu_tanh(bf16(acc + bias)))                  SiglipMLP / connector   */
#define BAGEL_EPI_SILU 4   /* C = bf16(silu(bf16(acc + bias)))                       TimestepEmbedder.mlp[0:2] */

/* C[M,N] = epilogue(A[M,K] @ W[N,K]^T), bf16 in / fp32 accumulate (tcgen05, TMEM) / bf16 out.
 * Replaces nn.Linear at modeling/bagel/qwen2_navit.py:515-517,529-536 (q/k/v_proj{,_moe_gen}),
 * :589-594 (o_proj{,_moe_gen}), modeling/qwen2/modeling_qwen2.py:200-201 (gate/up/down_proj),
 * modeling/bagel/bagel.py:803,832 (vae2llm, llm2vae), modeling/bagel/modeling_utils.py:84-110,120-124.
 *   W        nn.Linear weight layout [N, K]. For BAGEL_EPI_SWIGLU, W is [2*I, K] with gate/up rows
 *            interleaved in blocks of 128 (rows 256t..256t+127 = gate rows 128t.., next 128 = up rows) and
 *            C is [M, I].
 *   bias     [N] bf16 or NULL.   resid  [*, ldr] bf16 (BAGEL_EPI_RESID only).
 *   row_map  optional int32[M]: A-row r is written to C row row_map[r] (and reads resid row row_map[r]);
 *            used to scatter the und-expert rows of a MoT layer back into the packed sequence.
 * K, N, lda, ldw, ldc, ldr must be multiples of 8; pointers 16-byte aligned. */
int bagel_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M,
                    int N, int K, const void* bias, const void* resid, long long ldr, const int* row_map,
                    int epilogue, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BAGEL_B200_H_ */
