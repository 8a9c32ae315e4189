// Skinny-M (M <= 64) swapped-operand GEMM with cluster split-K; see gemm_skinny.cu. Called from bagel_gemm_bf16.
#pragma once
#include <cuda_runtime.h>

namespace bagel {

bool gemm_skinny_supported(int M, int N, int K, int epilogue);
int gemm_skinny(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N, int K,
                const void* bias, const void* resid, long long ldr, const int* row_map, int epilogue,
                cudaStream_t stream);

}  // namespace bagel
