"""tcgen05 bf16 GEMM entry points (csrc/gemm_tcgen05.cu).

``linear_bf16(x, w, bias)`` computes ``x @ w.T + bias`` like ``F.linear`` (the op
behind fc1/fc2, train_dist.py:61-62,68,70) on the 5th-gen tensor cores: TMA-fed
128B-swizzled smem tiles, ``tcgen05.mma`` with the fp32 accumulator in TMEM,
bias/ReLU fused in the ``tcgen05.ld`` epilogue.  No cuBLAS on this path."""
from __future__ import annotations

from typing import Optional

import torch

from . import _ext

__all__ = ["linear_bf16"]


def linear_bf16(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
                out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """``relu?(x[M,K] @ w[N,K]^T + bias[N])``; x, w bf16 (cast if needed), fp32 accumulate."""
    C = _ext.C()
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype != torch.bfloat16:
        x2 = x2.to(torch.bfloat16)
    if w.dtype != torch.bfloat16:
        w = w.to(torch.bfloat16)
    K = x2.shape[1]
    if K % 8:                                  # TMA needs 16-byte rows
        pad = 8 - K % 8
        x2 = torch.nn.functional.pad(x2, (0, pad))
        w = torch.nn.functional.pad(w, (0, pad))
    b = None if bias is None else bias.to(torch.float32).contiguous()
    out = C.gemm_bf16(x2.contiguous(), w.contiguous(), b, relu, out_dtype == torch.bfloat16)
    return out.view(*x.shape[:-1], w.shape[0])
