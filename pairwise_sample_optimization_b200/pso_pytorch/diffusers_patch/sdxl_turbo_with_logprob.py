"""Drop-in for ``sdxl_turbo_pipeline_with_logprob``
(human_preference_tuning/pso_pytorch/diffusers_patch/sdxl_turbo_with_logprob.py:53-161).

Per denoising step the reference runs ~30 eager kernels around the UNet call (input scaling :121,
the Euler-ancestral update, the noise add and the log-prob).  Here one fused launch per step
produces the next latents, their log-prob AND the next step's scaled UNet input
``latents / sqrt(sigma_next^2 + 1)``.  The UNet, VAE and text conditioning are the caller's
objects and are used exactly as the reference uses them.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import torch

from ... import _lib, runtime, step_ops


def prepare_latents(batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
    """:30-49 -- the turbo UNet's sample size is 64 (hard-coded there too)."""
    shape = (batch_size, num_channels_latents, 64, 64)
    if isinstance(generator, list) and len(generator) != batch_size:
        raise ValueError(
            f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
            f" size of {batch_size}. Make sure the batch size matches the length of the generators."
        )
    if latents is None:
        latents = torch.randn(shape, generator=generator, device=device, dtype=dtype)
    else:
        latents = latents.to(device)
    return latents


@torch.no_grad()
def sdxl_turbo_pipeline_with_logprob(
    accelerator,
    vae,
    unet,
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
    """Returns ``(image, all_latents, all_log_probs, all_model_input_latents)`` like :161."""
    dev = prompt_embeds.device
    _lib.require_cuda(prompt_embeds)
    with torch.autocast("cuda"):
        batch_size = prompt_embeds.shape[0]
        num_channels_latents = accelerator.unwrap_model(unet).config.in_channels
        latents = prepare_latents(batch_size * num_images_per_prompt, num_channels_latents, height, width,
                                  prompt_embeds.dtype, dev, generator, latents)
        init_sigma = noise_scheduler.init_noise_sigma                                        # :99
        if torch.is_tensor(init_sigma) and init_sigma.is_cuda:  # a device scalar: scale without reading it back
            latents = step_ops.scale_by_device_scalar(latents, init_sigma)
        else:
            latents = step_ops.scale(latents, float(init_sigma))
        noise_scheduler.set_timesteps(num_inference_steps, device=dev)                       # :102
        timesteps = noise_scheduler.timesteps
        sigmas = noise_scheduler.sigmas
        unet_added_conditions = {"time_ids": add_time_ids, "text_embeds": pooled_prompt_embeds}

        all_latents = [latents]
        all_model_input_latents = []
        all_log_probs = []
        # :120-121 (i = 0): latents / sqrt(sigma_0^2 + 1) with sigma_0 read on the device (the reference's `sigmas[i]` indexing
        # inside scale_model_input costs a host sync per step; here no step of the loop reads anything back)
        latent_model_input = step_ops.scale_by_device_scalar(latents, torch.rsqrt(sigmas[0].to(dev, torch.float32) ** 2 + 1))
        for i, t in enumerate(timesteps):
            noise_scheduler.is_scale_input_called = True
            noise_pred = unet(
                latent_model_input,
                t,
                encoder_hidden_states=prompt_embeds,
                added_cond_kwargs=unet_added_conditions,
                return_dict=False,
            )[0]
            ts = runtime.timesteps_on(t, dev)
            sched = runtime.turbo_schedule(noise_scheduler, dev, _lib.ts_dtype_code(ts))
            # fused: x_next = mu + sigma_up*noise, log_prob, and the NEXT step's scaled input     :136-142, :121
            if runtime.sampler_noise_in_kernel(generator):
                log_prob, latents_next, scaled_next = step_ops.step_forward(
                    sched, noise_pred, latents, ts, philox=step_ops.next_philox(), out_dtype=noise_pred.dtype,
                    want_scaled_next=True)
            else:
                noise = torch.randn(noise_pred.shape, dtype=noise_pred.dtype, device=dev, generator=generator)
                log_prob, latents_next, scaled_next = step_ops.step_forward(
                    sched, noise_pred, latents, ts, noise=noise, want_scaled_next=True)
            if i != num_inference_steps - 1:                                                   # :146-149
                all_model_input_latents.append(latent_model_input)
                all_latents.append(latents_next)
                all_log_probs.append(log_prob)
            latents = latents_next
            latent_model_input = scaled_next

        if not output_type == "latent":                                                        # :154-157
            image = vae.decode(latents / vae.config.scaling_factor, return_dict=False)[0]
        else:
            image = latents
        return image, all_latents, all_log_probs, all_model_input_latents
