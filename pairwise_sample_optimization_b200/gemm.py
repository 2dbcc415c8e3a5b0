"""Host side of ``psob200_lora_gemm`` (tensor plumbing only; the arithmetic is the tcgen05 kernel)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _mat(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dim() != 2:
        raise _lib.Psob200Error(f"{what}: expected a 2-D matrix, got shape {tuple(t.shape)}")
    if t.stride(1) != 1 or t.stride(0) % 8 or t.data_ptr() % 16:
        # TMA needs a contiguous reduction dimension, a 16-byte row pitch and a 16-byte aligned base
        cols = t.shape[1]
        pad = (-cols) % 8
        if pad:
            buf = torch.zeros(t.shape[0], cols + pad, dtype=t.dtype, device=t.device)
            buf[:, :cols] = t
            return buf[:, :cols]
        t = t.contiguous()
    return t


def lora_gemm(a1, b1, a2=None, b2=None, *, bias=None, alpha: float = 1.0, out=None, out_t=None, out_dtype=None,
              want_out: bool = True, want_out_t: bool = False, a_reduction_major: bool = False,
              accumulate: bool = False, split_k: int = 0, tune_bn: int = 0, diag: int = 0, pdl: int = 0):
    """D[M,N] = alpha * (a1 @ b1.T + a2 @ b2.T) + bias on the tcgen05 tensor cores.

    ``a1``: [M,K1] (or [K1,M] when ``a_reduction_major``), ``b1``: [N,K1], ``a2``: [M,K2], ``b2``: [N,K2]; bf16 or
    fp16.  Returns ``(D, Dt)`` where ``Dt`` is the transposed copy [N,M] if requested (else None).  With
    ``accumulate`` the given fp32 ``out`` / ``out_t`` are accumulated into (split-reduction with atomics).
    """
    dev = _lib.require_cuda(a1, b1, a2, b2, bias, out, out_t)
    if a1.dtype not in (torch.bfloat16, torch.float16) or b1.dtype != a1.dtype:
        raise _lib.Psob200Error(f"lora_gemm operands must both be bf16 or fp16, got {a1.dtype} / {b1.dtype}")
    a1m, b1m = _mat(a1, "a1"), _mat(b1, "b1")
    if a_reduction_major:
        K1, M = a1m.shape
    else:
        M, K1 = a1m.shape
    N = b1m.shape[0]
    if b1m.shape[1] != K1:
        raise _lib.Psob200Error(f"reduction lengths differ: a1 {tuple(a1.shape)} vs b1 {tuple(b1.shape)}")
    g = _lib.GemmArgs()
    g.a1, g.lda1, g.b1, g.ldb1 = a1m.data_ptr(), a1m.stride(0), b1m.data_ptr(), b1m.stride(0)
    keep = [a1m, b1m]
    K2 = 0
    if a2 is not None:
        if a2.dtype != a1.dtype or b2.dtype != a1.dtype:
            raise _lib.Psob200Error("second-segment operands must have the dtype of the first")
        a2m, b2m = _mat(a2, "a2"), _mat(b2, "b2")
        K2 = a2m.shape[1]
        if a2m.shape[0] != M or b2m.shape != (N, K2):
            raise _lib.Psob200Error(f"second segment shapes {tuple(a2.shape)} / {tuple(b2.shape)} do not match M={M}, N={N}")
        g.a2, g.lda2, g.b2, g.ldb2 = a2m.data_ptr(), a2m.stride(0), b2m.data_ptr(), b2m.stride(0)
        keep += [a2m, b2m]
    if accumulate:
        out_dtype = torch.float32
        if out is None and out_t is None:
            raise _lib.Psob200Error("accumulate=True needs the fp32 tensor(s) to accumulate into")
    elif out_dtype is None:
        out_dtype = a1.dtype
    pad8 = lambda n: (n + 7) // 8 * 8  # row pitch a multiple of 16 bytes: the result can feed the next launch's TMA as is
    if out is None and want_out and not (accumulate and out_t is not None):
        out = torch.empty(M, pad8(N), dtype=out_dtype, device=dev)[:, :N]
    if out_t is None and want_out_t and not accumulate:
        out_t = torch.empty(N, pad8(M), dtype=out_dtype, device=dev)[:, :M]
    for t, shape, name in ((out, (M, N), "out"), (out_t, (N, M), "out_t")):
        if t is not None and (tuple(t.shape) != shape or t.dtype != out_dtype or t.stride(1) != 1):
            raise _lib.Psob200Error(f"{name} must be a row-major {shape} {out_dtype} tensor")
    if out is not None:
        g.d, g.ldd = out.data_ptr(), out.stride(0)
    if out_t is not None:
        g.dt, g.lddt = out_t.data_ptr(), out_t.stride(0)
    if bias is not None:
        bias = bias.contiguous()
        if bias.numel() != N:
            raise _lib.Psob200Error(f"bias has {bias.numel()} entries, N = {N}")
        g.bias, g.bias_dtype = bias.data_ptr(), _lib.dtype_code(bias)
        keep.append(bias)
    g.M, g.N, g.K1, g.K2 = M, N, K1, K2
    g.alpha = float(alpha)
    g.ab_dtype = _lib.dtype_code(a1m)
    g.d_dtype = _lib._DTYPES[out_dtype]
    g.a_reduction_major = int(a_reduction_major)
    g.accumulate = int(accumulate)
    g.split_k, g.tune_bn, g.diag, g.pdl = int(split_k), int(tune_bn), int(diag), int(pdl)
    _lib.launch(dev, "psob200_lora_gemm", C.byref(g), _lib.current_stream(dev))
    del keep
    return out, out_t
