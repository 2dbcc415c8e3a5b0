"""Restated LoRA-wrapped projection and attention data flow (TEST INFRASTRUCTURE, CPU).

The arithmetic lives in un-vendored third parties (absent from /root/reference and
not installed here), so it is restated from their published behaviour and anchored
on the reference's call sites:

* peft==0.11.1 (environment.yml:22) ``tuners/lora/layer.py::Linear.forward``:
      result = base_layer(x)
      result = result + lora_B(lora_A(dropout(x.to(lora_A.weight.dtype)))) * scaling
  with ``scaling = lora_alpha / r`` and ``init_lora_weights="gaussian"``:
  ``normal_(lora_A.weight, std=1/r)``, ``zeros_(lora_B.weight)``.
  Call sites: train_online_pso_sdxl_turbo.py:338-345 (r=lora_rank, alpha=lora_rank,
  targets to_k/to_q/to_v/to_out.0), train_online_pso_sdxl_dmd2.py:361-368,
  train_pso_sdxl_turbo_dreambooth.py:1319-1326; adapter switching at
  turbo :790,805 (disable_adapters()/enable_adapters()).
* diffusers==0.27.0 (environment.yml:17) ``AttnProcessor2_0.__call__``: q/k/v
  projections -> [B, heads, L, head_dim] -> scaled_dot_product_attention -> merge
  heads -> to_out[0] -> to_out[1] (dropout) -> (+residual) / rescale_output_factor.

Parity status: third-party, restated -- "parity unpinned" against the packages
themselves; pinned only to autograd of this restatement.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def lora_linear(x, weight, bias, lora_A, lora_B, scaling: float = 1.0, adapters_enabled: bool = True):
    """y = x W^T (+ b) + scaling * (x A^T) B^T.   weight [N,K], lora_A [r,K], lora_B [N,r]."""
    result = F.linear(x, weight, bias)
    if not adapters_enabled:
        return result
    orig = result.dtype
    xa = x.to(lora_A.dtype)
    result = result + F.linear(F.linear(xa, lora_A), lora_B) * scaling
    return result.to(orig)


def lora_linear_grads(x, weight, lora_A, lora_B, grad_y, scaling: float = 1.0):
    """Closed-form backward in fp64:
    U = dY B ; dX = dY W + s U A ; dA = s U^T X ; dB = s dY^T (X A^T)."""
    x2 = x.double().reshape(-1, x.shape[-1])
    g2 = grad_y.double().reshape(-1, grad_y.shape[-1])
    W, A, Bm = weight.double(), lora_A.double(), lora_B.double()
    U = g2 @ Bm
    dX = g2 @ W + scaling * (U @ A)
    dA = scaling * (U.t() @ x2)
    dB = scaling * (g2.t() @ (x2 @ A.t()))
    return dX.reshape(x.shape), dA, dB


def gaussian_lora_init(r: int, in_features: int, out_features: int, generator=None, dtype=torch.float32):
    """peft ``init_lora_weights='gaussian'``: A ~ N(0, (1/r)^2), B = 0."""
    A = torch.randn(r, in_features, generator=generator, dtype=torch.float32) * (1.0 / r)
    Bm = torch.zeros(out_features, r, dtype=torch.float32)
    return A.to(dtype), Bm.to(dtype)


def attention_forward(hidden_states, encoder_hidden_states, proj, heads: int, residual_connection=False,
                      rescale_output_factor: float = 1.0):
    """AttnProcessor2_0 data flow for 3-D inputs.  ``proj(name, x)`` applies projection
    'to_q' | 'to_k' | 'to_v' | 'to_out' (to_out carries the bias)."""
    residual = hidden_states
    enc = hidden_states if encoder_hidden_states is None else encoder_hidden_states
    B = hidden_states.shape[0]
    q = proj("to_q", hidden_states)
    k = proj("to_k", enc)
    v = proj("to_v", enc)
    inner = k.shape[-1]
    hd = inner // heads
    q = q.view(B, -1, heads, hd).transpose(1, 2)
    k = k.view(B, -1, heads, hd).transpose(1, 2)
    v = v.view(B, -1, heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, -1, heads * hd).to(q.dtype)
    o = proj("to_out", o)
    if residual_connection:
        o = o + residual
    return o / rescale_output_factor


def lora_linear_rounded_flow(x, weight, bias, lora_A, lora_B, grad_y, scaling: float = 1.0, dtype=torch.bfloat16):
    """The same forward/backward with the 16-bit ROUNDING POINTS of a half-precision run made explicit (fp64 math
    in between): peft in bf16 materialises ``lora_A(x)`` in bf16 and autograd materialises ``grad @ lora_B`` in
    bf16; outputs y / dX are 16-bit, the adapter gradients are accumulated in fp32.  This is what the tcgen05 path
    is compared with: it must agree to 1 ulp of ``dtype`` on y, dX and to fp32 round-off on dA, dB."""
    rd = lambda t: t.to(dtype).double()
    x2 = x.double().reshape(-1, x.shape[-1])
    g2 = grad_y.double().reshape(-1, grad_y.shape[-1])
    W, A, Bm = weight.double(), rd(lora_A), rd(lora_B)
    T = rd(scaling * (x2 @ A.t()))
    y = x2 @ W.t() + T @ Bm.t()
    if bias is not None:
        y = y + bias.double()
    U = rd(scaling * (g2 @ Bm))
    dX = g2 @ W + U @ A
    dA = U.t() @ x2
    dB = g2.t() @ T
    return {"y": y.reshape(*x.shape[:-1], -1), "dX": dX.reshape(x.shape), "dA": dA, "dB": dB, "T": T, "U": U}


class OracleLoRALinear(torch.nn.Module):
    """peft==0.11.1 ``lora.Linear`` restated as a plain torch module (CPU oracle / reference arm): same attribute
    names as the real one (``base_layer``, ``lora_A``, ``lora_B``, ``scaling``, ``disable_adapters``)."""

    def __init__(self, base_layer: torch.nn.Linear, r: int, lora_alpha: int, adapter_name: str = "default"):
        super().__init__()
        self.base_layer = base_layer
        self.active_adapter = adapter_name
        self.r, self.lora_alpha, self.scaling = {adapter_name: r}, {adapter_name: lora_alpha}, {adapter_name: lora_alpha / r}
        dev, dt = base_layer.weight.device, torch.float32
        a = torch.nn.Linear(base_layer.in_features, r, bias=False, device=dev, dtype=dt)
        b = torch.nn.Linear(r, base_layer.out_features, bias=False, device=dev, dtype=dt)
        torch.nn.init.normal_(a.weight, std=1.0 / r)
        torch.nn.init.zeros_(b.weight)
        self.lora_A = torch.nn.ModuleDict({adapter_name: a})
        self.lora_B = torch.nn.ModuleDict({adapter_name: b})
        self.disable_adapters = False
        base_layer.weight.requires_grad_(False)
        if base_layer.bias is not None:
            base_layer.bias.requires_grad_(False)

    def forward(self, x):
        n = self.active_adapter
        return lora_linear(x, self.base_layer.weight, self.base_layer.bias, self.lora_A[n].weight, self.lora_B[n].weight,
                           self.scaling[n], adapters_enabled=not self.disable_adapters)


def oracle_add_adapter(model, r, lora_alpha, target_modules=("to_k", "to_q", "to_v", "to_out.0")):
    """``unet.add_adapter(LoraConfig(...))`` with the restated module; suffix matching like peft."""
    wrapped = []
    for parent_name, parent in list(model.named_modules()):
        for child_name, child in list(parent.named_children()):
            full = f"{parent_name}.{child_name}" if parent_name else child_name
            if isinstance(child, torch.nn.Linear) and any(full == t or full.endswith("." + t) for t in target_modules):
                new = OracleLoRALinear(child, r, lora_alpha)
                setattr(parent, child_name, new)
                wrapped.append(new)
    return wrapped


def oracle_set_adapters(model, enabled: bool):
    for m in model.modules():
        if isinstance(m, OracleLoRALinear):
            m.disable_adapters = not enabled
