"""Gated GELU of the transformer feed-forward on one fused kernel each way (psob200_geglu_forward / _backward).

diffusers==0.27.0 ``GEGLU.forward``::

    hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
    return hidden_states * self.gelu(gate)

``geglu(proj_out)`` replaces the second line (and the ``chunk``): the feed-forward lies between the LoRA-wrapped attention
blocks of every transformer layer and is NOT a row of the PSO hot path -- it is here because its stock
lowering (strided gelu, strided mul; gelu_backward, two strided muls and a concatenation in the backward) was the largest
single item of the measured micro-step.  No PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class _GegluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proj: torch.Tensor) -> torch.Tensor:
        dev = _lib.require_cuda(proj)
        two_i = proj.shape[-1]
        if two_i % 2:
            raise _lib.Psob200Error(f"GEGLU needs an even number of features, got {two_i}")
        p2 = proj.detach().reshape(-1, two_i)
        if p2.stride(-1) != 1 or p2.data_ptr() % 16:
            p2 = p2.contiguous()
        inner = two_i // 2
        out = torch.empty(p2.shape[0], inner, dtype=proj.dtype, device=dev)
        a = _lib.GegluArgs()
        a.proj, a.out, a.M, a.I = p2.data_ptr(), out.data_ptr(), p2.shape[0], inner
        a.ld_proj, a.ld_out, a.dtype = p2.stride(0), inner, _lib.dtype_code(p2)
        _lib.launch(dev, "psob200_geglu_forward", C.byref(a), _lib.current_stream(dev))
        ctx.save_for_backward(p2)
        ctx.shape = proj.shape
        return out.view(*proj.shape[:-1], inner)

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        (p2,) = ctx.saved_tensors
        inner = p2.shape[1] // 2
        d2 = dout.reshape(-1, inner)
        if d2.dtype != p2.dtype:
            d2 = d2.to(p2.dtype)
        if d2.stride(-1) != 1 or d2.data_ptr() % 16:
            d2 = d2.contiguous()
        dproj = torch.empty(p2.shape[0], 2 * inner, dtype=p2.dtype, device=p2.device)
        a = _lib.GegluArgs()
        a.proj, a.dout, a.dproj, a.M, a.I = p2.data_ptr(), d2.data_ptr(), dproj.data_ptr(), p2.shape[0], inner
        a.ld_proj, a.ld_dout, a.ld_dproj, a.dtype = p2.stride(0), d2.stride(0), 2 * inner, _lib.dtype_code(p2)
        _lib.launch(p2.device, "psob200_geglu_backward", C.byref(a), _lib.current_stream(p2.device))
        return dproj.view(ctx.shape)


def geglu(proj_out: torch.Tensor) -> torch.Tensor:
    """``hidden * gelu(gate)`` with ``hidden, gate = proj_out.chunk(2, dim=-1)`` (exact erf GELU), differentiable."""
    return _GegluFn.apply(proj_out)


def install_fused_geglu(model: torch.nn.Module) -> int:
    """Point every GEGLU-shaped module of ``model`` (one with a ``proj`` Linear whose forward is the two lines above, e.g.
    diffusers' ``GEGLU``) at the fused kernel.  Returns the number of modules patched."""
    n = 0
    for m in model.modules():
        if type(m).__name__ == "GEGLU" and hasattr(m, "proj"):
            m.forward = (lambda mod: (lambda x, *a, **k: geglu(mod.proj(x))))(m)
            n += 1
    return n
