#!/usr/bin/env python
"""Online PSO on the tiny SDXL-architecture fixture, written the way the reference's Turbo trainer is
(human_preference_tuning/train_online_pso_sdxl_turbo.py:544-861), on this package's drop-in API:

    sampling   two trajectories per prompt with ``sdxl_turbo_pipeline_with_logprob`` (TP:53-161; the per-step noise is drawn
               inside the step kernel), last step dropped (T:617,622-623)
    rewards    a synthetic scalar per image (no reward model here) -> ``sample_compare`` (T:401-416)
    training   for every trained timestep: policy forward, frozen-reference forward, ``pso_pair_loss`` (T:810-850 in one
               launch), backward through the LoRA projections (stacked q / k / v), optimizer boundary every
               ``accum`` micro-steps with ``FusedLoRAOptimizer`` (T:857-861)
    checkpoint ``checkpoint.save_state`` / ``load_state`` (T:886-889)

It is an EXAMPLE of the integration, not part of the product or the benchmark.

    python examples/online_pso_tiny.py [--epochs 3] [--prompts 4] [--rank 8]
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import pairwise_sample_optimization_b200 as pso  # noqa: E402
from fixtures import sdxl_unet  # noqa: E402
from pairwise_sample_optimization_b200 import checkpoint, lora  # noqa: E402


def euler_ancestral_schedule(num_steps: int = 4, device="cuda"):
    """The three attributes of diffusers' EulerAncestralDiscreteScheduler the path reads (T:264-267, TP:99-103): SDXL
    `scaled_linear` betas, `trailing` spacing -> timesteps [999, 749, 499, 249], sigmas with a trailing 0."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    ac = torch.cumprod(1.0 - betas, dim=0)
    sig_all = ((1 - ac) / ac) ** 0.5
    ts = torch.round(torch.arange(1000, 0, -1000 / num_steps, dtype=torch.float64)) - 1
    sig = torch.cat([sig_all[ts.long()], torch.zeros(1, dtype=torch.float64)]).float()
    sched = types.SimpleNamespace(timesteps=ts.float().to(device), sigmas=sig.to(device), init_noise_sigma=float(sig.max()),
                                  is_scale_input_called=False)
    sched.set_timesteps = lambda n, device=None: None
    return sched


class _Accelerator:  # the one method of accelerate.Accelerator the sampler pipeline calls (TP:79)
    @staticmethod
    def unwrap_model(m):
        return m


def build(rank: int = 8, seed: int = 0, lr: float = 2e-4):
    torch.manual_seed(seed)
    cfg = sdxl_unet.tiny_config()
    with torch.device("cuda"):
        unet = sdxl_unet.UNet2DConditionModel(cfg)
    unet = unet.to(torch.bfloat16).requires_grad_(False)
    lora.add_adapter(unet, lora.LoraConfig(r=rank, lora_alpha=rank, init_lora_weights="gaussian",
                                           target_modules=["to_k", "to_q", "to_v", "to_out.0"]))       # T:338-345
    unet.set_attn_processor(lora.PSOAttnProcessor2_0())
    lora.fuse_attention_projections(unet)
    lora.fuse_cross_attention_kv(unet)  # k / v of all cross-attention layers: one launch per UNet forward
    opt = lora.FusedLoRAOptimizer(unet, lr=lr, weight_decay=1e-4, max_grad_norm=1.0)                     # T:428-448, :859
    return cfg, unet, opt


def sample_epoch(unet, sched, cfg, n_prompts: int, steps: int = 4):
    """T:555-673: two trajectories per prompt; returns the stored tensors of the training phase."""
    dev = torch.device("cuda")
    pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
    prompt_embeds = torch.randn(n_prompts, 77, cfg.cross_attention_dim, device=dev, dtype=torch.bfloat16)
    text_embeds = torch.randn(n_prompts, pooled, device=dev, dtype=torch.bfloat16)
    time_ids = torch.tensor([[512.0, 512.0, 0.0, 0.0, 512.0, 512.0]], device=dev).repeat(n_prompts, 1)
    unet.eval()
    out = {"prompt_embeds": prompt_embeds, "text_embeds": text_embeds, "time_ids": time_ids}
    for k in (0, 1):
        image, all_latents, all_log_probs, all_inputs = pso.sdxl_turbo_pipeline_with_logprob(
            _Accelerator, None, unet, sched, 512, 512, num_inference_steps=steps, prompt_embeds=prompt_embeds,
            pooled_prompt_embeds=text_embeds, add_time_ids=time_ids.to(torch.bfloat16), output_type="latent")
        lat = torch.stack(all_latents, dim=1)                                      # [B, steps, 4, 64, 64]  (T:587-612)
        out[f"latents_{k}"] = lat[:, :-1]                                          # T:617: the last step is not trained
        out[f"next_latents_{k}"] = lat[:, 1:]
        # the sampler runs under autocast like the reference's (TP:77): cast what the UNet will be fed again to its dtype
        out[f"input_latents_{k}"] = torch.stack([t.to(torch.bfloat16) for t in all_inputs], dim=1)
        out[f"final_{k}"] = image
    unet.train()
    out["timesteps"] = sched.timesteps[:steps - 1].long()
    return out


def synthetic_reward(final_latents: torch.Tensor) -> torch.Tensor:
    """Stand-in for PickScore (T:632-648): prefers images whose first latent channel is brighter.  [B, 1]."""
    return final_latents[:, 0].float().mean(dim=(1, 2)).unsqueeze(1)


def train_epoch(unet, opt, sched, s, accum: int, beta: float = 50.0, eps: float = 0.1):
    """T:731-861 for one sampled batch: every pair at every trained timestep."""
    B, T = s["latents_0"].shape[:2]
    human_prefer = pso.sample_compare(synthetic_reward(s["final_0"]), synthetic_reward(s["final_1"]))   # T:842
    cond = {"text_embeds": s["text_embeds"], "time_ids": s["time_ids"].to(torch.bfloat16)}
    losses, micro = [], 0
    for j in range(T):
        ts = s["timesteps"][j].expand(B)
        preds, refs = [], []
        for k in (0, 1):                                                                               # T:775-787
            preds.append(unet(s[f"input_latents_{k}"][:, j], ts, s["prompt_embeds"], added_cond_kwargs=cond).sample)
        lora.disable_adapters(unet)                                                                    # T:790
        with torch.no_grad():
            for k in (0, 1):
                refs.append(unet(s[f"input_latents_{k}"][:, j], ts, s["prompt_embeds"], added_cond_kwargs=cond).sample)
        lora.enable_adapters(unet)                                                                     # T:805
        loss, stats = pso.pso_pair_loss(preds[0], preds[1], refs[0], refs[1], s["latents_0"][:, j], s["latents_1"][:, j],
                                        s["next_latents_0"][:, j], s["next_latents_1"][:, j], ts, ts, human_prefer,
                                        scheduler=sched, kind="turbo", beta=beta, eps=eps, loss_scale=1.0 / accum,
                                        return_stats=True)                                             # T:810-850
        loss.backward()                                                                                # T:857
        losses.append(loss.detach() * accum)
        micro += 1
        if micro % accum == 0:                                                                         # T:858-861
            opt.all_reduce()
            opt.step()
    return torch.stack(losses), stats[:, 6], human_prefer


def run(epochs: int = 3, prompts: int = 4, rank: int = 8, seed: int = 0, save_dir: str | None = None, verbose: bool = True):
    cfg, unet, opt = build(rank, seed)
    sched = euler_ancestral_schedule(4)
    history = []
    for epoch in range(epochs):
        s = sample_epoch(unet, sched, cfg, prompts)
        losses, z, hp = train_epoch(unet, opt, sched, s, accum=3)
        pso.check_status()
        rec = {"epoch": epoch, "loss": float(losses.mean()), "grad_norm": float(opt.grad_norm), "steps": int(opt.step_dev),
               "reward": float(torch.cat([synthetic_reward(s["final_0"]), synthetic_reward(s["final_1"])]).mean())}
        history.append(rec)
        if verbose:
            print(rec, flush=True)
    if save_dir:
        checkpoint.save_state(save_dir, unet, opt)
    return history, unet, opt


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--prompts", type=int, default=4)
    ap.add_argument("--rank", type=int, default=8)
    ap.add_argument("--save-dir", default="")
    a = ap.parse_args()
    run(a.epochs, a.prompts, a.rank, save_dir=a.save_dir or None)
