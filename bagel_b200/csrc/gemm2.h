// CTA-pair (tcgen05 cta_group::2) GEMM for the large projections: launcher used by bagel_gemm_bf16.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "gemm_params.h"

namespace bagel {
// true if the CTA-pair kernel handles this problem (large M, N % 256 == 0, BIAS / RESID / SWIGLU / QKV epilogue)
bool gemm2_supported(int M, int N, int K, int epilogue);
// p: M, N, K, C, ldc, bias, resid, ldr, row_map filled by the caller; tmA: box [128 rows, 64 cols], tmB: box [128, 64]
int gemm2_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams p, int epilogue, cudaStream_t stream);
}  // namespace bagel
