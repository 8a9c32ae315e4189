"""torch.Tensor-facing wrappers over the C ABI (include/bagel_b200.h). PyTorch is used only for device
memory and streams; every call below lands in a hand-written sm_100a kernel or raises."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _cabi

EPI_BIAS, EPI_RESID, EPI_SWIGLU, EPI_GELU, EPI_SILU = 0, 1, 2, 3, 4


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _cabi.BagelB200Error(f"{name}: expected a CUDA tensor (bagel_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _cabi.BagelB200Error(f"{name}: expected {dtype}, got {t.dtype}")
    if t.dim() >= 1 and t.stride(-1) != 1:
        raise _cabi.BagelB200Error(f"{name}: innermost dimension must be contiguous")


def interleave_gate_up(gate_w: torch.Tensor, up_w: torch.Tensor, block: int = 128) -> torch.Tensor:
    """[I,K],[I,K] -> [2I,K] with rows arranged (gate block of 128 | up block of 128) per 256 rows, the
    layout BAGEL_EPI_SWIGLU expects so one 256-wide tile holds matching gate/up columns."""
    I, K = gate_w.shape
    assert up_w.shape == (I, K) and I % block == 0
    g = gate_w.reshape(I // block, block, K)
    u = up_w.reshape(I // block, block, K)
    return torch.stack((g, u), dim=1).reshape(2 * I, K).contiguous()


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: Optional[torch.Tensor] = None,
         resid: Optional[torch.Tensor] = None, row_map: Optional[torch.Tensor] = None,
         epilogue: int = EPI_BIAS, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = epilogue(a @ w.T); a [M,K] bf16, w [N,K] bf16 (nn.Linear layout)."""
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    if out is None:
        assert row_map is None, "row_map scatter needs an explicit `out`"
        out = torch.empty((M, n_out), dtype=torch.bfloat16, device=a.device)
    _req(out, torch.bfloat16, "out")
    assert out.shape[1] == n_out
    if bias is not None:
        _req(bias, torch.bfloat16, "bias")
        assert bias.numel() == N
    ldr = 0
    if resid is not None:
        _req(resid, torch.bfloat16, "resid")
        ldr = resid.stride(0)
    if row_map is not None:
        _req(row_map, torch.int32, "row_map")
        assert row_map.numel() == M
    rc = _cabi.lib().bagel_gemm_bf16(_ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(out), out.stride(0),
                                     M, N, K, _ptr(bias), _ptr(resid), ldr, _ptr(row_map), epilogue, _stream())
    _cabi.check(rc, "bagel_gemm_bf16")
    return out
