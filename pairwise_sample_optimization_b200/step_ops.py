"""Host side of the step-with-logprob kernels (psob200_step_logprob*): tensor plumbing only."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, runtime


def _sample_numel(t: torch.Tensor) -> int:
    n = 1
    for s in t.shape[1:]:
        n *= int(s)
    return n


def _fill_common(args, model_output, sample, ts, ts_prev, coef, n):
    mo, st_mo = _lib.rows(model_output, n)
    sa, st_sa = _lib.rows(sample, n)
    args.model_output = mo.data_ptr()
    args.sample = sa.data_ptr()
    args.stride_model_output = st_mo
    args.stride_sample = st_sa
    args.ts = ts.data_ptr() if ts is not None else None
    args.ts_prev = ts_prev.data_ptr() if ts_prev is not None else None
    args.coef = coef.data_ptr() if coef is not None else None
    args.B = model_output.shape[0]
    args.N = n
    args.ts_rows = 1 if (ts is not None and ts.numel() == 1 and model_output.shape[0] > 1) else model_output.shape[0]
    args.pred_dtype = _lib.dtype_code(mo)
    args.latent_dtype = _lib.dtype_code(sa)
    return mo, sa


_PHILOX = {"offset": 0}


def next_philox(seed=None):
    """(seed, offset) for one in-kernel draw: torch's global seed (or ``seed``) and a process-wide call counter -- no device
    work, no sync.  Re-seeding torch (``torch.manual_seed``) re-keys the stream."""
    _PHILOX["offset"] += 1
    return (int(torch.initial_seed()) if seed is None else int(seed)) & 0xFFFFFFFFFFFFFFFF, _PHILOX["offset"]


def reset_philox(offset: int = 0) -> None:
    """Restart the call counter of the in-kernel noise stream (reproducible runs: seed torch, then reset)."""
    _PHILOX["offset"] = int(offset)


def step_forward(sched: runtime.ScheduleDesc, model_output, sample, ts, ts_prev=None, coef=None, prev_sample=None,
                 noise=None, want_scaled_next=False, tune=(0, 0), philox=None, noise_rows=None, out_dtype=None):
    """One launch of psob200_step_logprob.  Scoring mode if ``prev_sample`` is given, else sampling mode with explicit
    ``noise``, or -- ``philox=(seed, offset)`` -- with N(0,1) draws generated inside the kernel (``noise_rows`` = B or 1 for the
    batch-shared DMD2 draw, ``out_dtype`` = the dtype the noise would have had).  Returns (log_prob[B] fp32, prev_out | None,
    scaled_next | None)."""
    dev = _lib.require_cuda(model_output, sample, prev_sample, noise)
    B, n = model_output.shape[0], _sample_numel(model_output)
    if sample.shape[0] != B or _sample_numel(sample) != n:
        raise _lib.Psob200Error(f"sample shape {tuple(sample.shape)} does not match model_output {tuple(model_output.shape)}")
    if ts is not None and ts.numel() not in (1, B):
        raise _lib.Psob200Error(f"{ts.numel()} timesteps for a batch of {B}")
    args = _lib.StepArgs()
    keep = _fill_common(args, model_output, sample, ts, ts_prev, coef, n)
    log_prob = torch.empty(B, dtype=torch.float32, device=dev)
    args.log_prob = log_prob.data_ptr()
    args.status = runtime.status_word(dev).data_ptr()
    prev_out = scaled = None
    if prev_sample is not None:
        if prev_sample.dtype != sample.dtype:
            prev_sample = prev_sample.to(sample.dtype)
        ps, st_ps = _lib.rows(prev_sample, n)
        args.prev_sample = ps.data_ptr()
        args.stride_prev_sample = st_ps
        keep += (ps,)
    else:
        if philox is not None:
            if noise is not None:
                raise _lib.Psob200Error("pass either a noise tensor or philox=(seed, offset)")
            ndt = out_dtype or model_output.dtype
            args.use_philox, args.philox_seed, args.philox_offset = 1, int(philox[0]), int(philox[1])
            args.noise_rows = B if noise_rows is None else int(noise_rows)
        else:
            noise = noise.contiguous()
            ndt = noise.dtype
            args.noise = noise.data_ptr()
            args.noise_rows = noise.shape[0]
        if ndt not in (model_output.dtype, sample.dtype):
            raise _lib.Psob200Error("noise dtype must be the model_output dtype (turbo) or the sample dtype (DMD)")
        args.out_dtype = _lib._DTYPES[ndt]
        prev_out = torch.empty(model_output.shape, dtype=ndt, device=dev)
        args.prev_out = prev_out.data_ptr()
        if want_scaled_next:
            scaled = torch.empty_like(prev_out)
            args.scaled_next_out = scaled.data_ptr()
    args.tune_threads, args.tune_cluster = tune
    _lib.launch(dev, "psob200_step_logprob", sched.ref(), C.byref(args), _lib.current_stream(dev))
    del keep
    return log_prob, prev_out, scaled


def step_backward(sched: runtime.ScheduleDesc, model_output, sample, prev_sample, ts, ts_prev, coef, grad_log_prob):
    dev = _lib.require_cuda(model_output, sample, prev_sample, grad_log_prob)
    n = _sample_numel(model_output)
    args = _lib.StepBwdArgs()
    keep = _fill_common(args, model_output, sample, ts, ts_prev, coef, n)
    if prev_sample.dtype != sample.dtype:
        prev_sample = prev_sample.to(sample.dtype)
    ps, st_ps = _lib.rows(prev_sample, n)
    args.prev_sample = ps.data_ptr()
    args.stride_prev_sample = st_ps
    glp = grad_log_prob.to(torch.float32).contiguous()
    args.grad_log_prob = glp.data_ptr()
    grad = torch.empty(model_output.shape, dtype=model_output.dtype, device=dev)
    args.grad_model_output = grad.data_ptr()
    args.status = runtime.status_word(dev).data_ptr()
    _lib.launch(dev, "psob200_step_logprob_backward", sched.ref(), C.byref(args), _lib.current_stream(dev))
    del keep, ps, glp
    return grad


class StepLogProb(torch.autograd.Function):
    """log_prob[b] of ``prev_sample`` under the step policy, differentiable w.r.t. ``model_output`` only
    (the reference detaches prev_sample: turbo :109, distilled :130)."""

    @staticmethod
    def forward(ctx, model_output, sample, prev_sample, ts, ts_prev, coef, sched):
        log_prob, _, _ = step_forward(sched, model_output, sample, ts, ts_prev, coef, prev_sample=prev_sample)
        ctx.sched = sched
        ctx.save_for_backward(model_output, sample, prev_sample, ts, ts_prev, coef)
        return log_prob

    @staticmethod
    def backward(ctx, grad_log_prob):
        model_output, sample, prev_sample, ts, ts_prev, coef = ctx.saved_tensors
        grad = step_backward(ctx.sched, model_output, sample, prev_sample, ts, ts_prev, coef, grad_log_prob)
        return grad, None, None, None, None, None, None


def x0_from_noise(alphas_cumprod_dev, model_output, sample, ts, out_dtype=None, table_dtype=torch.float32):
    """psob200_dmd_x0_from_noise: distilled_inference_with_logprob.py:36-42.  The result type follows torch's promotion in
    the reference expression: ``alphas_cumprod[t].reshape(-1,1,1,1)`` is a 4-D tensor, so its dtype (``table_dtype``: fp32
    for every diffusers scheduler) takes part -- 16-bit latents and predictions still give an fp32 ``x0`` (the final DMD2
    latent that goes to the VAE, sdxl_dmd_with_logprob.py:158-162)."""
    dev = _lib.require_cuda(model_output, sample)
    n = _sample_numel(model_output)
    mo, sa = model_output.contiguous(), sample.contiguous()
    out_dtype = out_dtype or torch.promote_types(torch.promote_types(mo.dtype, sa.dtype), table_dtype)
    out = torch.empty(mo.shape, dtype=out_dtype, device=dev)
    _lib.launch(dev, "psob200_dmd_x0_from_noise",
        alphas_cumprod_dev.data_ptr(), alphas_cumprod_dev.numel(), mo.data_ptr(), sa.data_ptr(), ts.data_ptr(),
        _lib.ts_dtype_code(ts), 1 if ts.numel() == 1 and mo.shape[0] > 1 else mo.shape[0], out.data_ptr(),
        mo.shape[0], n, _lib.dtype_code(mo), _lib.dtype_code(sa), _lib.dtype_code(out),
        runtime.status_word(dev).data_ptr(), _lib.current_stream(dev))
    return out


def scale_by_device_scalar(t: torch.Tensor, factor: torch.Tensor) -> torch.Tensor:
    """``t * factor`` with ``factor`` a one-element tensor ON THE DEVICE: no read-back
    (psob200_scale_inplace_by_device_scalar on a copy)."""
    dev = _lib.require_cuda(t, factor)
    out = t.contiguous().clone()
    f32 = factor.detach().reshape(1).to(torch.float32)
    _lib.launch(dev, "psob200_scale_inplace_by_device_scalar", out.data_ptr(), out.numel(), _lib.dtype_code(out),
                f32.data_ptr(), _lib.current_stream(dev))
    return out


def scale(t: torch.Tensor, factor: float, out_dtype=None) -> torch.Tensor:
    """psob200_scale: out = t * factor (sdxl_turbo_with_logprob.py:99,121)."""
    dev = _lib.require_cuda(t)
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=out_dtype or t.dtype, device=dev)
    _lib.launch(dev, "psob200_scale", t.data_ptr(), out.data_ptr(), t.numel(), float(factor), _lib.dtype_code(t),
                                  _lib.dtype_code(out), _lib.current_stream(dev))
    return out
