"""Drop-in for the reference's ``turbo_step_with_logprob``
(human_preference_tuning/pso_pytorch/diffusers_patch/turbo_inference_with_logprob.py:24-116):
same name, argument meaning, return values and error behaviour, computed by one sm_100a
kernel launch (forward) + one (backward) instead of ~28 eager kernels and B host syncs.

Differences from the reference, all documented in DESIGN.md:
  * the arithmetic is fp32 on the stored values (the reference upcasts too: :69,:102);
  * the schedule lookup ``(_t == self.timesteps).nonzero()[0].item()`` (:63) happens on the
    device; a timestep that is not in the schedule raises IndexError from
    ``runtime.check_status()`` (a deliberate, explicit sync) instead of immediately;
  * no CPU path: non-CUDA tensors raise.
"""
from __future__ import annotations

import torch

from ... import _lib, runtime, step_ops


def turbo_step_with_logprob(self, model_output, timestep, sample, generator=None, prev_sample=None,
                            device=torch.device("cuda")):
    """One Euler-ancestral update and the Gaussian log-prob of ``prev_sample``.

    self:          scheduler (duck-typed: ``.timesteps`` and ``.sigmas`` are read, :63,:66,:77-78)
    model_output:  UNet epsilon prediction [B,C,H,W]
    timestep:      [B] (or [1], broadcast over the batch as in sdxl_turbo_with_logprob.py:139)
    sample:        current latents [B,C,H,W]
    prev_sample:   stored next latents -> scoring mode (:100-102); None -> sampling mode: noise is
                   drawn with ``generator`` exactly like :97 (``model_output.shape/dtype/device``)
    Returns ``(prev_sample.to(model_output.dtype), log_prob[B] float32)`` (:116); ``log_prob`` is
    differentiable w.r.t. ``model_output`` only (:109 detaches prev_sample).
    """
    dev = _lib.require_cuda(model_output, sample, prev_sample)
    ts = runtime.timesteps_on(timestep, dev)
    sched = runtime.turbo_schedule(self, dev, _lib.ts_dtype_code(ts))
    if prev_sample is None:
        if runtime.sampler_noise_in_kernel(generator):  # no generator given: the draw of :97 happens inside the kernel
            log_prob, prev_out, _ = step_ops.step_forward(sched, model_output.detach(), sample, ts, philox=step_ops.next_philox(),
                                                          out_dtype=model_output.dtype)
            return prev_out, log_prob
        noise = torch.randn(model_output.shape, dtype=model_output.dtype, device=dev, generator=generator)  # :97
        log_prob, prev_out, _ = step_ops.step_forward(sched, model_output.detach(), sample, ts, noise=noise)
        return prev_out, log_prob
    # like the reference (:100), a given prev_sample silently wins over a generator
    log_prob = step_ops.StepLogProb.apply(model_output, sample.detach(), prev_sample.detach(), ts, None, None, sched)
    return prev_sample.to(model_output.dtype), log_prob
