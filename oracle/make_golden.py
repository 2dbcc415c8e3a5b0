"""Generate tests/golden/*.npz by running the reference ITSELF (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

What produces each number:
  * log-probs / prev_samples of ``turbo_*`` and ``dmd_*`` cases: the reference's own
    ``turbo_step_with_logprob`` / ``distilled_step_with_logprob`` loaded verbatim by
    ``oracle.reference_loader`` (fp32, CPU).
  * ``online_*`` cases: the trainer's OWN lines executed verbatim (turbo trainer :810-850, dmd2 trainer
    :812-854: four step-with-logprob calls, ``sample_compare`` / ``compare``, the clamped ratios and the
    inline loss -- cut out of the trainer file at run time by ``reference_loader.trainer_loss_block``) and
    ``loss.backward()``; gradients are torch autograd through the reference's code.  The restatement
    ``oracle.losses.online_micro_step`` must reproduce loss and gradients bit for bit (asserted here).
  * ``dreambooth_*`` cases: likewise the DreamBooth trainer's lines :1846-1935 executed verbatim
    (``reference_loader.trainer_dreambooth_block``), restatement asserted bit-identical.
  * ``*_fp64`` arrays: the fp64 closed form (``oracle.losses.*_closed_form``), stored so the
    GPU tests can judge fp32 kernels against the truth rather than the reference's fp32 noise.

Small cases store their full inputs; full-size cases store the seed, input checksums and
output summaries (the inputs are regenerated from the seed by ``synth_online``).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import losses, reference_loader as rl, schedules, steps

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TURBO_TS = [999.0, 749.0, 499.0]
DMD_TS = [999, 749, 499]


def synth_online(kind, B, shape, seed, pred_noise=0.02, tie_every=0):
    """Synthetic micro-step inputs: next latents are drawn from the *reference
    policy's* step so that Delta is small and the clamp gate is open for pred_noise=0.02."""
    g = torch.Generator().manual_seed(seed)
    if kind == "turbo":
        sched = schedules.turbo_scheduler(4)
        step_ratio = None
    else:
        sched = schedules.dmd_scheduler()
        _, step_ratio = schedules.dmd_distill_timesteps(4)
    lat, nxt, pred, ref, ts = [], [], [], [], []
    tsel = torch.randint(0, 3, (B,), generator=g)
    for k in (0, 1):
        if kind == "turbo":
            t = torch.tensor(TURBO_TS)[tsel]
            sig = sched.sigmas[tsel].reshape(-1, 1, 1, 1)
            x = torch.randn(B, *shape, generator=g) * sig
        else:
            t = torch.tensor(DMD_TS)[tsel]
            x = torch.randn(B, *shape, generator=g)
        e_ref = torch.randn(B, *shape, generator=g)
        e_pol = e_ref + pred_noise * torch.randn(B, *shape, generator=g)
        noise = torch.randn(B, *shape, generator=g)
        if kind == "turbo":
            xn, _ = steps.turbo_step(sched, e_ref, t, x, noise=noise)
        else:
            xn, _ = steps.distilled_step(sched, e_ref, t, t - step_ratio, x, noise=noise)
        lat.append(x); nxt.append(xn); pred.append(e_pol); ref.append(e_ref); ts.append(t)
    sign = torch.randint(0, 2, (B,), generator=g).float() * 2 - 1
    h = torch.stack([-sign, sign], 1)
    if tie_every:
        h[::tie_every] = 0.0
    return dict(kind=kind, sched=sched, step_ratio=step_ratio, latents=lat, next_latents=nxt,
                noise_pred=pred, noise_ref_pred=ref, timesteps=ts, human_prefer=h)


def _np(t):
    return t.detach().cpu().numpy()


def make_step_cases():
    shape = (4, 16, 16)
    B = 3
    # ---- turbo
    sched = schedules.turbo_scheduler(4)
    fn = rl.turbo_step_with_logprob()
    g = torch.Generator().manual_seed(11)
    t = torch.tensor(TURBO_TS)
    sig = sched.sigmas[:3].reshape(-1, 1, 1, 1)
    x = torch.randn(B, *shape, generator=g) * sig
    e = torch.randn(B, *shape, generator=g)
    xn_s, lp_s = fn(sched, e, t, x, generator=torch.Generator().manual_seed(5), device="cpu")
    noise = torch.randn(e.shape, generator=torch.Generator().manual_seed(5), dtype=e.dtype)
    e2 = (e + 0.05 * torch.randn(B, *shape, generator=g)).requires_grad_(True)
    xn_out, lp = fn(sched, e2, t, x, prev_sample=xn_s, device="cpu")
    wgt = torch.tensor([1.0, -2.0, 0.5])
    (lp * wgt).sum().backward()
    lp64 = steps.turbo_logprob_closed_form(sched, e2.detach(), t, x, xn_s)
    np.savez_compressed(
        os.path.join(OUT, "turbo_step.npz"), timesteps=_np(t), sample=_np(x), model_output_sampling=_np(e),
        noise=_np(noise), prev_sample_sampling=_np(xn_s), log_prob_sampling=_np(lp_s),
        model_output_scoring=_np(e2), log_prob_scoring=_np(lp), grad_weights=_np(wgt),
        grad_model_output=_np(e2.grad), log_prob_scoring_fp64=_np(lp64), sigmas=_np(sched.sigmas),
        sched_timesteps=_np(sched.timesteps),
    )
    # ---- dmd
    sched = schedules.dmd_scheduler()
    fn = rl.distilled_step_with_logprob()
    _, sr = schedules.dmd_distill_timesteps(4)
    t = torch.tensor(DMD_TS)
    x = torch.randn(B, *shape, generator=g)
    e = torch.randn(B, *shape, generator=g)
    xn_s, lp_s = fn(sched, e, t, t - sr, x, generator=torch.Generator().manual_seed(6), device="cpu")
    noise = torch.randn((1,) + shape, generator=torch.Generator().manual_seed(6), dtype=x.dtype)
    e2 = (e + 0.05 * torch.randn(B, *shape, generator=g)).requires_grad_(True)
    _, lp = fn(sched, e2, t, t - sr, x, prev_sample=xn_s, device="cpu")
    (lp * wgt).sum().backward()
    lp64 = steps.dmd_logprob_closed_form(sched, e2.detach(), t, t - sr, x, xn_s)
    x0_last = rl.get_x0_from_noise()(x, e, sched.alphas_cumprod, torch.tensor([249, 249, 249]))
    np.savez_compressed(
        os.path.join(OUT, "dmd_step.npz"), timesteps=_np(t), prev_timesteps=_np(t - sr), sample=_np(x),
        model_output_sampling=_np(e), noise=_np(noise), prev_sample_sampling=_np(xn_s),
        log_prob_sampling=_np(lp_s), model_output_scoring=_np(e2), log_prob_scoring=_np(lp),
        grad_weights=_np(wgt), grad_model_output=_np(e2.grad), log_prob_scoring_fp64=_np(lp64),
        x0_last_step=_np(x0_last), alphas_cumprod=_np(sched.alphas_cumprod),
    )


def _online_case(name, kind, B, shape, seed, pred_noise, tie_every, full_inputs, beta=50.0, eps=0.1):
    d = synth_online(kind, B, shape, seed, pred_noise, tie_every)
    fn = rl.turbo_step_with_logprob() if kind == "turbo" else rl.distilled_step_with_logprob()
    pred = [p.clone().requires_grad_(True) for p in d["noise_pred"]]
    loss, lps = losses.online_micro_step(kind, d["sched"], pred, d["noise_ref_pred"], d["latents"],
                                         d["next_latents"], d["timesteps"], d["human_prefer"], beta, eps,
                                         step_ratio=d["step_ratio"], step_fns=fn)
    loss.backward()
    # the same numbers from the trainer's own lines: rewards chosen so that its compare function yields human_prefer
    h = d["human_prefer"]
    r0 = torch.where(h[:, 0] < 0, torch.zeros(B), torch.where(h[:, 0] > 0, torch.ones(B), torch.full((B,), 0.5)))
    r1 = 1.0 - r0  # ties (dmd2 only: compare -> [0, 0]) have r0 == r1 == 0.5
    mk = lambda k, r: {"timesteps": d["timesteps"][k][:, None], "latents": d["latents"][k][:, None],
                       "next_latents": d["next_latents"][k][:, None], "rewards": r[:, None] if kind == "turbo" else r}
    if kind == "dmd" or not tie_every:  # the turbo trainer's sample_compare has no tie outcome
        pred_v = [p.clone().requires_grad_(True) for p in d["noise_pred"]]
        loss_v, h_v = rl.trainer_loss_block(kind)(d["sched"], pred_v, d["noise_ref_pred"], mk(0, r0), mk(1, r1), 0, beta, eps,
                                                  d["step_ratio"])
        loss_v.backward()
        assert torch.equal(h_v, h) and torch.equal(loss_v, loss), (name, loss_v.item(), loss.item())
        assert all(torch.equal(a.grad, b.grad) for a, b in zip(pred_v, pred)), name
    cf = losses.online_closed_form(kind, d["sched"], d["noise_pred"], d["noise_ref_pred"], d["latents"],
                                   d["next_latents"], d["timesteps"], d["human_prefer"], beta, eps,
                                   step_ratio=d["step_ratio"])
    out = dict(
        kind=kind, B=B, shape=np.array(shape), seed=seed, pred_noise=pred_noise, tie_every=tie_every,
        beta=beta, eps=eps, loss=_np(loss), logp=np.stack([_np(v) for v in lps]),
        loss_fp64=_np(cf["loss"]), logp_fp64=np.stack([_np(v) for v in cf["logp"]]),
        delta_fp64=np.stack([_np(v) for v in cf["delta"]]), z_fp64=_np(cf["z"]),
        human_prefer=_np(d["human_prefer"]), timesteps=np.stack([_np(v) for v in d["timesteps"]]),
    )
    for k in (0, 1):
        gr, g64 = pred[k].grad, cf["grads"][k]
        out[f"grad{k}_sum"] = _np(gr.double().sum())
        out[f"grad{k}_abs_sum"] = _np(gr.double().abs().sum())
        out[f"grad{k}_row_l2"] = _np(gr.double().reshape(B, -1).norm(dim=1))
        out[f"grad{k}_fp64_row_l2"] = _np(g64.reshape(B, -1).norm(dim=1))
        out[f"grad{k}_head"] = _np(gr.reshape(B, -1)[:, :64])
        out[f"grad{k}_fp64_head"] = _np(g64.reshape(B, -1)[:, :64])
        out[f"in_checksum{k}"] = np.array([float(v.double().sum()) for v in
                                           (d["latents"][k], d["next_latents"][k], d["noise_pred"][k],
                                            d["noise_ref_pred"][k])])
        if full_inputs:
            out[f"latents{k}"] = _np(d["latents"][k]); out[f"next_latents{k}"] = _np(d["next_latents"][k])
            out[f"noise_pred{k}"] = _np(d["noise_pred"][k]); out[f"noise_ref_pred{k}"] = _np(d["noise_ref_pred"][k])
            out[f"grad{k}"] = _np(gr); out[f"grad{k}_fp64"] = _np(g64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def synth_dreambooth(b, shape, seed, pred_noise=0.05):
    """Synthetic DreamBooth-PSO step inputs following train_pso_sdxl_turbo_dreambooth.py:1763-1796."""
    g = torch.Generator().manual_seed(seed)
    sched = schedules.dreambooth_scheduler()
    x0 = torch.randn(2 * b, *shape, generator=g) * 0.8
    noise = torch.randn(2 * b, *shape, generator=g).chunk(2)[0].repeat(2, 1, 1, 1)       # :1763 shared noise
    raw = torch.randint(0, 1000, (b,), generator=g)
    idx = (250 * (raw % 4) + 249).long().repeat(2)                                       # :1769-1779
    timesteps = sched.timesteps[idx]                                                     # :1781
    sig = torch.stack([sched.sigmas[(sched.timesteps == t).nonzero().item()] for t in timesteps]).reshape(-1, 1, 1, 1)
    noisy = x0 + sig * noise                                                             # :1787 (EulerDiscrete.add_noise)
    ref_pred = noise + 0.3 * torch.randn(2 * b, *shape, generator=g)
    model_pred = ref_pred + pred_noise * torch.randn(2 * b, *shape, generator=g)
    return dict(model_pred=model_pred, ref_pred=ref_pred, noisy=noisy, x0=x0, sigmas=sig, timesteps=timesteps)


def _dreambooth_case(name, b, shape, seed, loss_type, beta, nu, lam):
    d = synth_dreambooth(b, shape, seed)
    mp = d["model_pred"].clone().requires_grad_(True)
    loss, lw, ll, logits = losses.dreambooth_pso_loss(mp, d["ref_pred"], d["noisy"], d["x0"], d["sigmas"],
                                                      loss_type, beta, nu, lam)
    loss.backward()
    mp_v = d["model_pred"].clone().requires_grad_(True)
    loss_v, lw_v, ll_v, logits_v = rl.trainer_dreambooth_block()(mp_v, d["ref_pred"], d["noisy"], d["x0"], d["sigmas"], loss_type,
                                                                  beta, nu, lam)
    loss_v.backward()
    assert torch.equal(loss_v, loss) and torch.equal(mp_v.grad, mp.grad) and torch.equal(logits_v, logits), name
    cf = losses.dreambooth_closed_form(d["model_pred"], d["ref_pred"], d["noisy"], d["x0"], d["sigmas"],
                                       loss_type, beta, nu, lam)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), b=b, shape=np.array(shape), seed=seed, loss_type=loss_type,
        beta_pso=beta, neg_defactor=nu, prior_loss_weight=lam, model_pred=_np(d["model_pred"]),
        ref_pred=_np(d["ref_pred"]), noisy=_np(d["noisy"]), x0=_np(d["x0"]), sigmas=_np(d["sigmas"]),
        loss=_np(loss), losses_w=_np(lw), losses_l=_np(ll), logits=_np(logits), grad=_np(mp.grad),
        loss_fp64=_np(cf["loss"]), grad_fp64=_np(cf["grad"]), logits_fp64=_np(cf["logits"]),
        losses_w_fp64=_np(cf["losses_w"]), losses_l_fp64=_np(cf["losses_l"]),
    )


def main():
    if not rl.available():
        raise SystemExit("needs /root/reference (build container only)")
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    make_step_cases()
    small = (4, 16, 16)
    _online_case("online_turbo_small", "turbo", 4, small, 21, 0.02, 0, True)
    _online_case("online_turbo_small_gate", "turbo", 4, small, 22, 0.5, 0, True)
    _online_case("online_dmd_small", "dmd", 4, small, 23, 0.02, 3, True)
    _online_case("online_dmd_small_gate", "dmd", 4, small, 24, 0.5, 0, True)
    _online_case("online_turbo_full", "turbo", 4, (4, 64, 64), 31, 0.02, 0, False)
    _online_case("online_dmd_full", "dmd", 2, (4, 128, 128), 32, 0.02, 0, False)
    _dreambooth_case("dreambooth_pso", 2, small, 41, "pso", 5.0, 0.1, 0.5)
    _dreambooth_case("dreambooth_pso_db", 2, small, 42, "pso_db", 5.0, 0.1, 0.5)
    _dreambooth_case("dreambooth_pso_noprior", 3, small, 43, "pso", 2.0, 1.0, 0.0)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
