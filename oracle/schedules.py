"""Scheduler constants the hot path reads (TEST INFRASTRUCTURE).

These live in diffusers==0.27.0 (environment.yml:17; 0.27.2 per
environment_fix.sh:19), which is NOT vendored under /root/reference and not
installed here.  They are restated from the published algorithm of
``EulerAncestralDiscreteScheduler`` / ``LCMScheduler`` / ``EulerDiscreteScheduler``
for the SDXL scheduler config (beta_start=0.00085, beta_end=0.012,
beta_schedule="scaled_linear", num_train_timesteps=1000,
timestep_spacing="trailing" for sdxl-turbo).  Anchors: sigma(t=999)=14.6146 (the
well-known SDXL init_noise_sigma) and SURVEY.md App. A.1's five sigmas.

Reference call sites these objects serve:
  train_online_pso_sdxl_turbo.py:264-267  (EulerAncestral .timesteps/.sigmas)
  sdxl_turbo_with_logprob.py:99-103,120   (.init_noise_sigma, .set_timesteps, .sigmas[i])
  train_online_pso_sdxl_dmd2.py:285-288,542-550 (LCM .alphas_cumprod, distill_timesteps)
  train_pso_sdxl_turbo_dreambooth.py:1235-1237,1675-1685,1787 (EulerDiscrete .sigmas/.timesteps/.add_noise)

The step functions touch only ``.timesteps``, ``.sigmas`` and
``.alphas_cumprod`` (turbo_inference_with_logprob.py:63,66,77-78;
distilled_inference_with_logprob.py:85,98), so a SimpleNamespace is a valid
``self``.
"""
from __future__ import annotations

import types

import numpy as np
import torch

NUM_TRAIN_TIMESTEPS = 1000
BETA_START = 0.00085
BETA_END = 0.012


def sdxl_alphas_cumprod() -> torch.Tensor:
    """scaled_linear betas -> alphas_cumprod, fp32 [1000] (diffusers scheduling_*.py __init__)."""
    betas = torch.linspace(BETA_START ** 0.5, BETA_END ** 0.5, NUM_TRAIN_TIMESTEPS, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def sdxl_sigmas_all() -> np.ndarray:
    """sigma(t) = sqrt((1-abar_t)/abar_t) for t=0..999 (numpy, as diffusers computes it)."""
    ac = sdxl_alphas_cumprod().numpy()
    return np.array(((1 - ac) / ac) ** 0.5)


def trailing_timesteps(num_inference_steps: int) -> np.ndarray:
    """'trailing' spacing: round(arange(N, 0, -N/n)) - 1  ->  [999, 749, 499, 249] for n=4."""
    step_ratio = NUM_TRAIN_TIMESTEPS / num_inference_steps
    return (np.round(np.arange(NUM_TRAIN_TIMESTEPS, 0, -step_ratio)) - 1).astype(np.float32)


def turbo_scheduler(num_inference_steps: int = 4) -> types.SimpleNamespace:
    """Duck-typed EulerAncestralDiscreteScheduler after ``set_timesteps(n)`` (trailing spacing).

    ``sigmas`` has n+1 entries (trailing 0); ``timesteps`` is float32 like diffusers'.
    """
    sig_all = sdxl_sigmas_all()
    ts = trailing_timesteps(num_inference_steps)
    sig = np.interp(ts, np.arange(0, len(sig_all)), sig_all)
    sig = np.concatenate([sig, [0.0]]).astype(np.float32)
    sched = types.SimpleNamespace(
        timesteps=torch.from_numpy(ts),
        sigmas=torch.from_numpy(sig),
        init_noise_sigma=float(sig.max()),  # trailing/linspace spacing: sigmas.max()
        alphas_cumprod=sdxl_alphas_cumprod(),
        num_inference_steps=num_inference_steps,
        is_scale_input_called=False,
    )
    sched.set_timesteps = lambda n, device=None: None  # sdxl_turbo_with_logprob.py:102 (already set)
    return sched


def dmd_scheduler() -> types.SimpleNamespace:
    """Duck-typed LCMScheduler: only ``alphas_cumprod`` (and init_noise_sigma=1) are read."""
    return types.SimpleNamespace(alphas_cumprod=sdxl_alphas_cumprod(), init_noise_sigma=1.0)


def dmd_distill_timesteps(num_steps: int = 4):
    """train_online_pso_sdxl_dmd2.py:542-550 -> (LongTensor [999,749,499,249], step_ratio)."""
    step_ratio = NUM_TRAIN_TIMESTEPS // num_steps
    ts = (np.arange(num_steps, 0, -1) * step_ratio).round() - 1
    return torch.tensor(ts, dtype=torch.float32).long(), step_ratio


def dreambooth_scheduler() -> types.SimpleNamespace:
    """Duck-typed EulerDiscreteScheduler at construction (no set_timesteps call):
    timesteps = [999..0] float32, sigmas = reversed sigma_all + [0]
    (train_pso_sdxl_turbo_dreambooth.py:1675-1685 looks sigma up by timestep equality;
    :1781 indexes ``timesteps`` by position; :1787 add_noise = x0 + sigma*noise)."""
    sig_all = sdxl_sigmas_all()
    ts = np.linspace(0, NUM_TRAIN_TIMESTEPS - 1, NUM_TRAIN_TIMESTEPS, dtype=float)[::-1].copy()
    sig = np.concatenate([sig_all[::-1], [0.0]]).astype(np.float32)
    return types.SimpleNamespace(
        timesteps=torch.from_numpy(ts.astype(np.float32)),
        sigmas=torch.from_numpy(sig),
        alphas_cumprod=sdxl_alphas_cumprod(),
    )


def turbo_coefficients(sigmas: torch.Tensor, step_index: torch.Tensor):
    """Per-sample (k, a, s) of the affine form mu = k*x + a*eps, std s,
    computed in fp64 from turbo_inference_with_logprob.py:77-92."""
    sg = sigmas.double()
    s_from = sg[step_index]
    s_to = sg[step_index + 1]
    s_up = (s_to ** 2 * (s_from ** 2 - s_to ** 2) / s_from ** 2) ** 0.5
    s_down = (s_to ** 2 - s_up ** 2) ** 0.5
    return torch.ones_like(s_from), s_down - s_from, s_up


def dmd_coefficients(alphas_cumprod: torch.Tensor, t: torch.Tensor, t_prev: torch.Tensor):
    """(k, a, s) in fp64 from distilled_inference_with_logprob.py:36-42,102-112."""
    ac = alphas_cumprod.double()
    a_t = ac[t.long()]
    a_p = ac[t_prev.long()]
    k = a_p.sqrt() / a_t.sqrt()
    a = -a_p.sqrt() * (1 - a_t).sqrt() / a_t.sqrt()
    s = (1 - a_p).sqrt()
    return k, a, s
