"""Restated few-step sampler loops (TEST INFRASTRUCTURE, CPU).

``turbo_sampler`` follows sdxl_turbo_with_logprob.py:86-149 (VAE decode :152-157 is out of
scope); ``dmd_sampler`` follows sdxl_dmd_with_logprob.py:89-162.  ``unet`` is any callable
``unet(latent_input, t) -> noise_pred`` (conditioning is closed over by the caller);
``noises`` is the list of per-step draws so an implementation under test can be fed
identical noise (the reference draws it inside the step: turbo step :97, distilled step :123).
"""
from __future__ import annotations

import torch

from . import steps


def turbo_sampler(unet, scheduler, latents, noises, num_inference_steps=4, step_fn=None):
    """Returns (final_latents, all_latents, all_log_probs, all_model_input_latents)."""
    latents = latents * scheduler.init_noise_sigma                     # :99
    timesteps = scheduler.timesteps                                    # :103
    all_latents, all_inputs, all_lp = [latents], [], []
    for i, t in enumerate(timesteps):                                  # :116
        sigma = scheduler.sigmas[i]                                    # :120
        latent_in = latents / ((sigma ** 2 + 1) ** 0.5)                # :121
        noise_pred = unet(latent_in, t)                                # :126-132
        if step_fn is None:
            latents, lp = steps.turbo_step(scheduler, noise_pred, t.unsqueeze(0), latents, noise=noises[i])
        else:
            latents, lp = step_fn(scheduler, noise_pred, t.unsqueeze(0), latents, noises[i])
        if i != num_inference_steps - 1:                               # :146-149
            all_inputs.append(latent_in)
            all_latents.append(latents)
            all_lp.append(lp)
    return latents, all_latents, all_lp, all_inputs


def dmd_sampler(unet, scheduler, timesteps, latents, noises, step_fn=None):
    """Returns (x0_pred, all_latents, all_log_probs)."""
    latents = latents * scheduler.init_noise_sigma                     # prepare_latents :49
    B = latents.shape[0]
    all_latents, all_lp = [latents], []
    x0_pred = None
    for i, t in enumerate(timesteps):                                  # :112
        cur = torch.ones(B, dtype=torch.long) * t                      # :113
        noise_pred = unet(latents, cur)                                # :117-122
        if i != timesteps.shape[0] - 1:                                # :124
            prev = torch.ones(B, dtype=torch.long) * timesteps[i + 1]  # :125-126
            if step_fn is None:
                latents, lp = steps.distilled_step(scheduler, noise_pred, cur, prev, latents, noise=noises[i])
            else:
                latents, lp = step_fn(scheduler, noise_pred, cur, prev, latents, noises[i])
            all_latents.append(latents)
            all_lp.append(lp)
        else:                                                          # :154-162
            x0_pred = steps.x0_from_noise(latents, noise_pred, scheduler.alphas_cumprod, cur)
            all_latents.append(x0_pred)
    return x0_pred, all_latents, all_lp
