"""Drop-in for ``sdxl_dmd_pipeline_with_logprob``
(human_preference_tuning/pso_pytorch/diffusers_patch/sdxl_dmd_with_logprob.py:54-174): the
few-step DMD2 sampler; every step's update + log-prob is one fused launch, the last step's x0
recovery (:158-162) another.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import torch

from ... import _lib, runtime, step_ops


def prepare_latents(scheduler, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
    """:30-50 -- 128x128 latents, scaled by ``scheduler.init_noise_sigma``."""
    shape = (batch_size, num_channels_latents, 128, 128)
    if isinstance(generator, list) and len(generator) != batch_size:
        raise ValueError(
            f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
            f" size of {batch_size}. Make sure the batch size matches the length of the generators."
        )
    if latents is None:
        latents = torch.randn(shape, generator=generator, device=device, dtype=dtype)
    else:
        latents = latents.to(device)
    sigma = float(scheduler.init_noise_sigma)
    return latents if sigma == 1.0 else step_ops.scale(latents, sigma)


@torch.no_grad()
def sdxl_dmd_pipeline_with_logprob(
    accelerator,
    vae,
    unet,
    timesteps,
    noise_scheduler,
    height,
    width,
    num_inference_steps: int = 4,
    guidance_scale: float = 0.0,
    negative_prompt: Optional[Union[str, List[str]]] = None,
    num_images_per_prompt: Optional[int] = 1,
    generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
    latents: Optional[torch.FloatTensor] = None,
    prompt_embeds: Optional[torch.FloatTensor] = None,
    pooled_prompt_embeds: Optional[torch.FloatTensor] = None,
    add_time_ids: Optional[torch.FloatTensor] = None,
    negative_prompt_embeds: Optional[torch.FloatTensor] = None,
    output_type: Optional[str] = "pil",
    return_dict: bool = True,
    callback: Optional[Callable[[int, int, torch.FloatTensor], None]] = None,
    callback_steps: int = 1,
    cross_attention_kwargs: Optional[Dict[str, Any]] = None,
    guidance_rescale: float = 0.0,
):
    """Returns ``(image, all_latents, all_log_probs)`` like :174."""
    dev = prompt_embeds.device
    _lib.require_cuda(prompt_embeds)
    with torch.autocast("cuda"):
        batch_size = prompt_embeds.shape[0]
        num_channels_latents = accelerator.unwrap_model(unet).config.in_channels
        latents = prepare_latents(noise_scheduler, batch_size * num_images_per_prompt, num_channels_latents, height,
                                  width, prompt_embeds.dtype, dev, generator, latents)
        unet_added_conditions = {"time_ids": add_time_ids, "text_embeds": pooled_prompt_embeds}
        if torch.is_tensor(noise_scheduler.alphas_cumprod) and noise_scheduler.alphas_cumprod.device != dev:
            noise_scheduler.alphas_cumprod = noise_scheduler.alphas_cumprod.to(dev)
        timesteps = torch.as_tensor(timesteps).to(dev)
        all_latents = [latents]
        all_log_probs = []
        x0_pred = None
        n_steps = timesteps.shape[0]
        for i in range(n_steps):
            t = timesteps[i]
            current_timesteps = torch.ones(batch_size, device=dev, dtype=torch.long) * t                    # :113
            noise_pred = unet(latents, current_timesteps, prompt_embeds, added_cond_kwargs=unet_added_conditions).sample
            ts = runtime.timesteps_on(t, dev)
            if i != n_steps - 1:                                                                               # :124
                ts_prev = runtime.timesteps_on(timesteps[i + 1], dev).to(ts.dtype)
                sched = runtime.dmd_schedule(noise_scheduler, dev, _lib.ts_dtype_code(ts))
                if runtime.sampler_noise_in_kernel(generator):                                             # DS:123-124
                    log_prob, latents, _ = step_ops.step_forward(sched, noise_pred, latents, ts, ts_prev, noise_rows=1,
                                                                 philox=step_ops.next_philox(), out_dtype=latents.dtype)
                else:
                    noise = torch.randn((1,) + tuple(noise_pred.shape[1:]), generator=generator, device=dev,
                                        dtype=latents.dtype)
                    log_prob, latents, _ = step_ops.step_forward(sched, noise_pred, latents, ts, ts_prev, noise=noise)
                all_latents.append(latents)
                all_log_probs.append(log_prob)
            else:                                                                                              # :154-162
                x0_pred = step_ops.x0_from_noise(runtime.device_table(noise_scheduler.alphas_cumprod, dev),
                                                 noise_pred, latents, ts,
                                                 table_dtype=noise_scheduler.alphas_cumprod.dtype)  # fp32 like DS:36-42
                all_latents.append(x0_pred)

        if not output_type == "latent":
            image = vae.decode(x0_pred / vae.config.scaling_factor, return_dict=False)[0]
        else:
            image = x0_pred
    return image, all_latents, all_log_probs
