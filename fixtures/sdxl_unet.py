"""SDXL-architecture UNet scaffold in plain torch (random init) -- a FIXTURE, not product code.

diffusers is not installed in this image, so the reference's
``UNet2DConditionModel.from_pretrained(...)`` (train_online_pso_sdxl_turbo.py:288-294) cannot be used to host the
LoRA-wrapped attention projections.  This module restates the topology from the published SDXL ``unet/config.json``
(block_out_channels (320, 640, 1280), transformer_layers_per_block (1, 2, 10), attention_head_dim (5, 10, 20) = heads,
cross_attention_dim 2048, text_time additional embedding) with the diffusers==0.27.0 module surface the PSO hot path
touches: ``Attention`` objects with ``to_q/to_k/to_v/to_out[0]/to_out[1]/heads/processor/set_processor``, qualified
names ending in ``attn1.to_q`` ... ``attn2.to_out.0`` (so peft-style ``target_modules`` match), the forward signature
``unet(sample, timestep, encoder_hidden_states, added_cond_kwargs={"text_embeds", "time_ids"})`` returning an object
with ``.sample``, and ``enable_gradient_checkpointing()``.

Everything here runs on stock torch kernels (convolutions, norms, SDPA: outside the PSO hot path); only the 560
attention projections are replaced by the tcgen05 LoRA path when ``lora.add_adapter`` wraps them.
``tiny_config()`` is the diffusers "dummy SDXL UNet" of BASELINE config 1 (32/64 channels).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    down_block_types: Tuple[str, ...] = ("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D")
    up_block_types: Tuple[str, ...] = ("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D")
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)
    attention_head_dim: Tuple[int, ...] = (5, 10, 20)  # = number of heads (diffusers naming quirk)
    cross_attention_dim: int = 2048
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    norm_num_groups: int = 32
    sample_size: int = 128


def sdxl_config() -> UNetConfig:
    return UNetConfig()


def tiny_config() -> UNetConfig:
    return UNetConfig(block_out_channels=(32, 64), down_block_types=("DownBlock2D", "CrossAttnDownBlock2D"),
                      up_block_types=("CrossAttnUpBlock2D", "UpBlock2D"), transformer_layers_per_block=(1, 2),
                      attention_head_dim=(2, 4), cross_attention_dim=64, addition_time_embed_dim=8,
                      projection_class_embeddings_input_dim=80, sample_size=32)


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers ``Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)``."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class AttnProcessor2_0:
    """Stock diffusers==0.27.0 processor data flow (the oracle-side processor of the tests)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
        b = hidden_states.shape[0]
        enc = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q, k, v = attn.to_q(hidden_states), attn.to_k(enc), attn.to_v(enc)
        hd = k.shape[-1] // attn.heads
        q, k, v = (t.view(b, -1, attn.heads, hd).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, attn.heads * hd).to(q.dtype)
        return attn.to_out[1](attn.to_out[0](o))


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim: Optional[int], heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale = heads, dim_head ** -0.5
        self.residual_connection, self.rescale_output_factor = False, 1.0
        self.is_cross_attention = cross_attention_dim is not None  # (diffusers 0.27.0 Attention sets the same attribute)
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])
        self.processor = AttnProcessor2_0()

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kwargs)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim // heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_dim, heads, dim // heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, enc):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), encoder_hidden_states=enc)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, cross_dim, depth, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, cross_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, enc):
        b, c, h, w = x.shape
        res = x
        y = self.norm(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        y = self.proj_in(y)
        for blk in self.transformer_blocks:
            y = blk(y, enc)
        y = self.proj_out(y).reshape(b, h, w, c).permute(0, 3, 1, 2)
        return y + res


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    """Down / up block: ``resnets[i]`` (+ ``attentions[i]`` when cross-attention) then an optional resampler."""

    def __init__(self, res_io, temb_dim, groups, attn: Optional[tuple], down: bool, resample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(i, o, temb_dim, groups) for (i, o) in res_io])
        cout = res_io[-1][1]
        self.attentions = None
        if attn is not None:
            heads, cross_dim, depth = attn
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim, depth, groups) for _ in res_io])
        self.resampler = (Downsample2D(cout) if down else Upsample2D(cout)) if resample else None
        self.is_down = down


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.config = cfg
        ch = cfg.block_out_channels
        temb_dim = ch[0] * 4
        g = cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb_dim)
        self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, temb_dim)
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        self._skip_channels = [out]
        for i, kind in enumerate(cfg.down_block_types):
            cin, out = out, ch[i]
            last = i == len(ch) - 1
            attn = (cfg.attention_head_dim[i], cfg.cross_attention_dim, cfg.transformer_layers_per_block[i]) \
                if kind.startswith("CrossAttn") else None
            io = [(cin if j == 0 else out, out) for j in range(cfg.layers_per_block)]
            self.down_blocks.append(_Block(io, temb_dim, g, attn, True, not last))
            self._skip_channels += [out] * cfg.layers_per_block + ([out] if not last else [])
        self.mid_block = nn.ModuleDict({
            "resnets": nn.ModuleList([ResnetBlock2D(out, out, temb_dim, g), ResnetBlock2D(out, out, temb_dim, g)]),
            "attentions": nn.ModuleList([Transformer2DModel(out, cfg.attention_head_dim[-1], cfg.cross_attention_dim,
                                                            cfg.transformer_layers_per_block[-1], g)])})
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        rev_heads = list(reversed(cfg.attention_head_dim))
        rev_depth = list(reversed(cfg.transformer_layers_per_block))
        skips = list(self._skip_channels)
        prev = rev[0]
        for i, kind in enumerate(cfg.up_block_types):
            cout = rev[i]
            last = i == len(ch) - 1
            io = []
            for j in range(cfg.layers_per_block + 1):
                skip = skips.pop()
                io.append(((prev if j == 0 else cout) + skip, cout))
            attn = (rev_heads[i], cfg.cross_attention_dim, rev_depth[i]) if kind.startswith("CrossAttn") else None
            self.up_blocks.append(_Block(io, temb_dim, g, attn, False, not last))
            prev = cout
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)
        self.gradient_checkpointing = False

    # ---- diffusers surface used by the trainers
    def enable_gradient_checkpointing(self):
        self.gradient_checkpointing = True

    def attention_modules(self):
        return [m for m in self.modules() if isinstance(m, Attention)]

    def set_attn_processor(self, processor):
        for m in self.attention_modules():
            m.set_processor(processor)

    def _run(self, module, *args):
        if self.gradient_checkpointing and self.training and torch.is_grad_enabled():
            return checkpoint(module, *args, use_reentrant=False, preserve_rng_state=False)
        return module(*args)

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None, return_dict=True, **_):
        cfg = self.config
        b = sample.shape[0]
        t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], device=sample.device)
        t = t.reshape(-1).to(sample.device).expand(b)
        emb = self.time_embedding(timestep_embedding(t, cfg.block_out_channels[0]).to(sample.dtype))
        time_ids = added_cond_kwargs["time_ids"]
        tid = timestep_embedding(time_ids.flatten(), cfg.addition_time_embed_dim).reshape(b, -1)
        add = torch.cat([added_cond_kwargs["text_embeds"], tid.to(sample.dtype)], dim=-1)
        emb = emb + self.add_embedding(add.to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            for j, res in enumerate(blk.resnets):
                x = self._run(res, x, emb)
                if blk.attentions is not None:
                    x = self._run(blk.attentions[j], x, encoder_hidden_states)
                skips.append(x)
            if blk.resampler is not None:
                x = blk.resampler(x)
                skips.append(x)
        x = self._run(self.mid_block["resnets"][0], x, emb)
        x = self._run(self.mid_block["attentions"][0], x, encoder_hidden_states)
        x = self._run(self.mid_block["resnets"][1], x, emb)
        for blk in self.up_blocks:
            for j, res in enumerate(blk.resnets):
                x = torch.cat([x, skips.pop()], dim=1)
                x = self._run(res, x, emb)
                if blk.attentions is not None:
                    x = self._run(blk.attentions[j], x, encoder_hidden_states)
            if blk.resampler is not None:
                x = blk.resampler(x)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        if not return_dict:
            return (x,)
        return SimpleNamespace(sample=x)


def count_lora_targets(unet: nn.Module) -> int:
    return sum(1 for n, m in unet.named_modules() if isinstance(m, nn.Linear) and
               any(n.endswith(s) for s in (".to_q", ".to_k", ".to_v", ".to_out.0")))
