"""Shared helpers for the GPU parity tests."""
import numpy as np
import torch

from oracle import losses, make_golden

DT = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


def cuda(t, dtype=None):
    return t.to(device="cuda", dtype=dtype) if dtype is not None else t.to("cuda")


def synth(kind, B, shape, seed, pred_noise=0.02, tie_every=0, pred_dtype=torch.float32, latent_dtype=torch.float32):
    """Seeded micro-step inputs, rounded to the storage dtypes on the CPU so that the oracle and the
    kernels see identical values ("half storage, fp32 math")."""
    d = make_golden.synth_online(kind, B, shape, seed, pred_noise, tie_every)
    for key, dt in (("noise_pred", pred_dtype), ("noise_ref_pred", pred_dtype), ("latents", latent_dtype),
                    ("next_latents", latent_dtype)):
        d[key] = [t.to(dt) for t in d[key]]
    return d


def oracle_fp64(d, beta=50.0, eps=0.1, loss_scale=1.0):
    up = {k: [t.double() for t in d[k]] for k in ("noise_pred", "noise_ref_pred", "latents", "next_latents")}
    return losses.online_closed_form(d["kind"], d["sched"], up["noise_pred"], up["noise_ref_pred"], up["latents"],
                                     up["next_latents"], d["timesteps"], d["human_prefer"], beta, eps,
                                     step_ratio=d["step_ratio"], loss_scale=loss_scale)


def run_fused(pso, d, beta=50.0, eps=0.1, loss_scale=1.0, tune=(0, 0), stats=True):
    pred = [cuda(p).requires_grad_(True) for p in d["noise_pred"]]
    out = pso.pso_pair_loss(pred[0], pred[1], cuda(d["noise_ref_pred"][0]), cuda(d["noise_ref_pred"][1]),
                            cuda(d["latents"][0]), cuda(d["latents"][1]), cuda(d["next_latents"][0]),
                            cuda(d["next_latents"][1]), cuda(d["timesteps"][0]), cuda(d["timesteps"][1]),
                            cuda(d["human_prefer"]), scheduler=d["sched"], kind=d["kind"], beta=beta, eps=eps,
                            step_ratio=d["step_ratio"], loss_scale=loss_scale, return_stats=stats, tune=tune)
    loss, st = out if stats else (out, None)
    loss.backward()
    return loss, st, pred[0].grad, pred[1].grad


def rel_max(got, want):
    """max |got - want| / max |want|  (the norm the fp32 1e-5 target is stated in)."""
    got, want = got.double().cpu(), want.double().cpu()
    den = want.abs().max().item()
    return (got - want).abs().max().item() / (den if den > 0 else 1.0)


def assert_rounded_equal(got, want64, dtype, max_mismatch_frac=0.01):
    """``got`` (in ``dtype``) equals the fp64 truth rounded to ``dtype`` up to 1 ulp, and almost everywhere
    exactly -- the strictest statement possible for a half-precision output."""
    want = want64.to(dtype)
    g, w = got.cpu().float(), want.float()
    exact = (g == w)
    frac = 1.0 - exact.float().mean().item()
    assert frac <= max_mismatch_frac, f"{frac:.4%} of elements differ from the correctly rounded value"
    # spacing of representable values around w (fp16 subnormals have a fixed spacing of 2^-24)
    ulp = w.abs() * (2.0 ** (-7 if dtype == torch.bfloat16 else -10))
    ulp = torch.clamp(ulp, min=(2.0 ** -133 if dtype == torch.bfloat16 else 2.0 ** -24))
    # where the residual cancels (|r| << |x'|, |x|, |eps|) the fp32 evaluation carries an ABSOLUTE error of a few
    # fp32 ulps of the operands; grant that on top of the output rounding
    slack = 4e-6 * w.abs().max()
    bad = (g - w).abs() > ulp + slack
    assert not bool(bad.any()), f"{int(bad.sum())} elements are off by more than 1 ulp (+fp32 cancellation slack)"
