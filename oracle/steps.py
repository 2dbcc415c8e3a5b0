"""Restated step-with-logprob functions (TEST INFRASTRUCTURE, CPU).

``turbo_step``      follows turbo_inference_with_logprob.py:24-116
``distilled_step``  follows distilled_inference_with_logprob.py:45-137

Both keep the reference's operation order in fp32 so that their rounding is the
reference's rounding; ``tests/test_oracle_golden.py`` pins them against outputs
of the verbatim reference functions.  One deliberate difference, documented in
SURVEY.md finding 5 / App. C: ``distilled_step(..., upcast=True)`` (the default)
evaluates on inputs upcast to fp32 -- "bf16 storage, fp32 math" -- instead of
reproducing the reference's half-precision arithmetic when latents are stored in
half precision (distilled :84-86,99).  ``upcast=False`` reproduces the quirk.

``*_closed_form`` evaluate the same quantities in fp64 through the affine form
mu = k*x + a*eps; they are what fp32 parity is judged
against (finding 4: the reference's own fp32 noise is ~2e-5 on the gradient).
"""
from __future__ import annotations

import math

import torch

from . import schedules

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def turbo_step_indices(scheduler, timestep) -> list[int]:
    """turbo :61-64 -- first index where ``t == scheduler.timesteps`` (IndexError if absent)."""
    out = []
    for _t in timestep:
        hits = (_t == scheduler.timesteps).nonzero()
        out.append(int(hits[0].item()))
    return out


def turbo_step(scheduler, model_output, timestep, sample, generator=None, prev_sample=None, noise=None):
    """Restatement of turbo_step_with_logprob (:24-116).  ``noise`` lets a caller inject the
    draw of :97 so sampling mode is reproducible across implementations."""
    idx = turbo_step_indices(scheduler, timestep)                     # :61-64
    nxt = [i + 1 for i in idx]
    sigma = scheduler.sigmas[idx].reshape(-1, 1, 1, 1)                # :66
    sample = sample.to(torch.float32)                                 # :69
    pred_x0 = sample - sigma * model_output                           # :73
    s_from = scheduler.sigmas[idx]                                    # :77
    s_to = scheduler.sigmas[nxt]                                      # :78
    s_up = (s_to ** 2 * (s_from ** 2 - s_to ** 2) / s_from ** 2) ** 0.5   # :79
    s_down = (s_to ** 2 - s_up ** 2) ** 0.5                           # :80
    s_up = s_up.reshape(-1, 1, 1, 1)
    s_down = s_down.reshape(-1, 1, 1, 1)
    derivative = (sample - pred_x0) / sigma                           # :88
    dt = s_down - sigma                                               # :90
    mean = sample + derivative * dt                                   # :92
    if prev_sample is None:                                           # :94-99
        if noise is None:
            noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
        prev_sample = mean + noise * s_up
    else:
        prev_sample = prev_sample.to(torch.float32)                   # :102
    log_prob = (
        -((prev_sample.detach() - mean) ** 2) / (2 * (s_up ** 2))
        - torch.log(s_up)
        - torch.log(torch.sqrt(2 * torch.as_tensor(math.pi)))
    )                                                                 # :108-112
    log_prob = log_prob.mean(dim=tuple(range(1, log_prob.ndim)))      # :114
    return prev_sample.to(model_output.dtype), log_prob               # :116


def x0_from_noise(sample, model_output, alphas_cumprod, timestep):
    """distilled :36-42."""
    a_t = alphas_cumprod[timestep.long()].reshape(-1, 1, 1, 1)
    b_t = 1 - a_t
    return (sample - b_t ** 0.5 * model_output) / a_t ** 0.5


def distilled_step(scheduler, model_output, timestep, prev_timestep, sample, generator=None,
                   prev_sample=None, noise=None, upcast=True):
    """Restatement of distilled_step_with_logprob (:45-137)."""
    if prev_sample is not None and generator is not None:             # :115-119
        raise ValueError("Cannot pass both generator and prev_sample.")
    out_dtype = sample.dtype
    if upcast:
        sample = sample.float()
        model_output = model_output.float()
        prev_sample = None if prev_sample is None else prev_sample.float()
    x0 = x0_from_noise(sample, model_output, scheduler.alphas_cumprod, timestep).to(sample.dtype)  # :84-86
    ac = scheduler.alphas_cumprod.to(dtype=x0.dtype)                  # :99
    sa = (ac[prev_timestep.long()] ** 0.5).flatten().reshape(-1, 1, 1, 1)        # :102-105
    s1 = ((1 - ac[prev_timestep.long()]) ** 0.5).flatten().reshape(-1, 1, 1, 1)  # :107-110
    mean = sa * x0                                                    # :112
    if prev_sample is None:                                           # :121-126  one draw shared by the batch
        if noise is None:
            noise = torch.randn((1,) + tuple(x0.shape[1:]), generator=generator, dtype=sample.dtype)
        prev_sample = mean + s1 * noise
    log_prob = (
        -((prev_sample.detach() - mean) ** 2) / (2 * (s1 ** 2))
        - torch.log(s1)
        - torch.log(torch.sqrt(2 * torch.as_tensor(math.pi)))
    )                                                                 # :129-133
    log_prob = log_prob.mean(dim=tuple(range(1, log_prob.ndim)))      # :135
    return prev_sample.type(out_dtype), log_prob                      # :137


# ----------------------------------------------------------------------------- fp64 closed forms
def affine_logprob_closed_form(model_output, sample, prev_sample, k, a, s):
    """logp_b = -S_b/(2 s^2 N) - log s - 0.5 log 2pi with S_b = sum (x' - k x - a eps)^2, fp64."""
    x = sample.double()
    e = model_output.double()
    xn = prev_sample.double()
    B = x.shape[0]
    kk, aa, ss = (v.double().reshape(B, *([1] * (x.ndim - 1))) for v in (k, a, s))
    r = xn - kk * x - aa * e
    n = r[0].numel()
    S = (r * r).reshape(B, -1).sum(1)
    return -S / (2 * ss.reshape(B) ** 2 * n) - torch.log(ss.reshape(B)) - _HALF_LOG_2PI, r


def turbo_logprob_closed_form(scheduler, model_output, timestep, sample, prev_sample):
    idx = torch.tensor(turbo_step_indices(scheduler, timestep))
    k, a, s = schedules.turbo_coefficients(scheduler.sigmas, idx)
    return affine_logprob_closed_form(model_output, sample, prev_sample, k, a, s)[0]


def dmd_logprob_closed_form(scheduler, model_output, timestep, prev_timestep, sample, prev_sample):
    k, a, s = schedules.dmd_coefficients(scheduler.alphas_cumprod, timestep, prev_timestep)
    return affine_logprob_closed_form(model_output, sample, prev_sample, k, a, s)[0]
