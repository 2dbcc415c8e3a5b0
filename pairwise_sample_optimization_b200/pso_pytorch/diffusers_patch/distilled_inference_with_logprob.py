"""Drop-in for the reference's ``distilled_step_with_logprob`` and ``_get_x0_from_noise``
(human_preference_tuning/pso_pytorch/diffusers_patch/distilled_inference_with_logprob.py:36-42,45-137).

Differences from the reference:
  * "half storage, fp32 math": the reference evaluates in ``sample.dtype`` (:84-86,:99), which
    destroys the log-prob differences when latents are stored in bf16/fp16 (SURVEY.md finding 5);
    here the stored values are upcast in registers and all arithmetic is fp32;
  * ``alphas_cumprod`` lookups with a negative index wrap like torch indexing (prev_timestep = -1);
  * no CPU path.
"""
from __future__ import annotations

import torch

from ... import _lib, runtime, step_ops


def _get_x0_from_noise(sample, model_output, alphas_cumprod, timestep):
    """x0 = (sample - sqrt(1-abar_t) * model_output) / sqrt(abar_t)   (:36-42)."""
    dev = _lib.require_cuda(sample, model_output)
    ts = runtime.timesteps_on(timestep, dev)
    return step_ops.x0_from_noise(runtime.device_table(alphas_cumprod, dev), model_output, sample, ts,
                                  table_dtype=alphas_cumprod.dtype)


def distilled_step_with_logprob(self, model_output, timestep, prev_timestep, sample, eta: float = 0.0,
                                use_clipped_model_output: bool = False, generator=None, prev_sample=None,
                                device=torch.device("cuda")):
    """One DMD2/LCM update ``x' = sqrt(abar_prev) x0 + sqrt(1-abar_prev) noise`` and its log-prob.

    self: scheduler (duck-typed: only ``.alphas_cumprod`` is read; like :98 it is moved to the
    compute device in place).  Scoring mode when ``prev_sample`` is given; sampling mode draws ONE
    noise tensor of shape (1,C,H,W) in ``sample.dtype`` shared by the whole batch (:123-124).
    Raises ValueError when both ``generator`` and ``prev_sample`` are passed (:115-119).
    Returns ``(prev_sample.type(sample.dtype), log_prob[B])`` (:137).
    """
    if prev_sample is not None and generator is not None:
        raise ValueError(
            "Cannot pass both generator and prev_sample. Please make sure that either `generator` or"
            " `prev_sample` stays `None`."
        )
    dev = _lib.require_cuda(model_output, sample, prev_sample)
    if torch.is_tensor(self.alphas_cumprod) and self.alphas_cumprod.device != dev:
        self.alphas_cumprod = self.alphas_cumprod.to(device=dev)  # :98
    ts = runtime.timesteps_on(timestep, dev)
    ts_prev = runtime.timesteps_on(prev_timestep, dev).to(ts.dtype)
    sched = runtime.dmd_schedule(self, dev, _lib.ts_dtype_code(ts))
    if prev_sample is None:
        if runtime.sampler_noise_in_kernel(generator):  # one in-kernel draw shared by the batch (:123-124)
            log_prob, prev_out, _ = step_ops.step_forward(sched, model_output.detach(), sample, ts, ts_prev,
                                                          philox=step_ops.next_philox(), noise_rows=1, out_dtype=sample.dtype)
            return prev_out, log_prob
        noise = torch.randn((1,) + tuple(model_output.shape[1:]), generator=generator, device=dev,
                            dtype=sample.dtype)  # :123-124
        log_prob, prev_out, _ = step_ops.step_forward(sched, model_output.detach(), sample, ts, ts_prev, noise=noise)
        return prev_out, log_prob
    log_prob = step_ops.StepLogProb.apply(model_output, sample.detach(), prev_sample.detach(), ts, ts_prev, None, sched)
    return prev_sample.type(sample.dtype), log_prob
