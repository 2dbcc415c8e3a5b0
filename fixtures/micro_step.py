"""One online-PSO training micro-step on the SDXL-architecture fixture, written the way the reference's trainers
write it (train_online_pso_sdxl_turbo.py:771-861), once against the product API (CUDA) and once against the oracle
(CPU).  FIXTURE code shared by tests/ and bench.py; the oracle variant is test / baseline infrastructure only."""
from __future__ import annotations

import types

import torch

TURBO_TS = [999, 749, 499]  # trained timesteps: first 3 of the 4-step trailing schedule (turbo :617)


DMD_STEP_RATIO = 250        # 1000 // num_steps (dmd2 trainer :542); prev_timestep = t - step_ratio (:816)


def synth_batch(B, latent_hw, cross_dim, pooled_dim, seed, sigmas=None, dtype=torch.float32, device="cpu", kind="turbo"):
    """Synthetic stored trajectories of one micro-step: per pair and branch the current latent, the
    UNet input, the stored next latent, one trained timestep per pair, prompt embeddings and the reward sign.
    kind="turbo": latents are sigma-scaled noise and the UNet sees latent / sqrt(sigma^2+1) (turbo :121, :776);
    kind="dmd":   unit-variance latents, fed to the UNet unscaled (dmd2 :778), 1024x1024 time ids (:347-355)."""
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, len(TURBO_TS), (B,), generator=g)
    ts = torch.tensor(TURBO_TS)[idx]  # both samplers train the first 3 of [999, 749, 499, 249]
    out = {"timesteps": ts, "step_index": idx}
    if kind == "turbo":
        sig = sigmas[idx].float()[:, None, None, None]
        sig_next = sigmas[idx + 1].float()[:, None, None, None]
        in_scale = 1.0 / (sig ** 2 + 1) ** 0.5
    else:
        sig = sig_next = in_scale = torch.ones(B, 1, 1, 1)
    for k in (0, 1):
        lat = torch.randn(B, 4, latent_hw, latent_hw, generator=g) * sig
        out[f"latents_{k}"] = lat
        out[f"input_latents_{k}"] = lat * in_scale
        out[f"next_latents_{k}"] = torch.randn(B, 4, latent_hw, latent_hw, generator=g) * sig_next
    out["prompt_embeds"] = torch.randn(B, 77, cross_dim, generator=g)
    out["text_embeds"] = torch.randn(B, pooled_dim, generator=g)
    px = 8.0 * latent_hw
    out["time_ids"] = torch.tensor([[px, px, 0.0, 0.0, px, px]]).repeat(B, 1)
    sign = torch.randint(0, 2, (B,), generator=g).float() * 2 - 1
    out["human_prefer"] = torch.stack([-sign, sign], 1)
    cast = lambda t: t.to(device=device, dtype=dtype) if t.is_floating_point() else t.to(device)
    return {k: (cast(v) if k not in ("human_prefer", "time_ids") else v.to(device)) for k, v in out.items()}


def _unet_call(unet, x, ts, batch):
    return unet(x, ts, batch["prompt_embeds"], added_cond_kwargs={"text_embeds": batch["text_embeds"],
                                                                  "time_ids": batch["time_ids"].to(x.dtype)}).sample


def _loss_kind(kind):
    return {"kind": kind, "step_ratio": DMD_STEP_RATIO if kind == "dmd" else None}


def product_micro_step(pso, lora, unet, batch, sched, *, beta=50.0, eps=0.1, loss_scale=1.0, kind="turbo"):
    """Product path: 2 policy forwards with grad, 2 adapter-disabled forwards without (turbo :775-805), the fused
    loss+grad kernel in place of the four step calls + inline loss (:810-850), backward (:857)."""
    ts = batch["timesteps"]
    p0 = _unet_call(unet, batch["input_latents_0"], ts, batch)
    p1 = _unet_call(unet, batch["input_latents_1"], ts, batch)
    lora.disable_adapters(unet)
    with torch.no_grad():
        r0 = _unet_call(unet, batch["input_latents_0"], ts, batch)
        r1 = _unet_call(unet, batch["input_latents_1"], ts, batch)
    lora.enable_adapters(unet)
    loss = pso.pso_pair_loss(p0, p1, r0, r1, batch["latents_0"], batch["latents_1"],
                             batch["next_latents_0"], batch["next_latents_1"], ts, ts, batch["human_prefer"],
                             scheduler=sched, beta=beta, eps=eps, loss_scale=loss_scale, **_loss_kind(kind))
    loss.backward()
    return loss


def batched_view(batch):
    """Both branches stacked along the batch dimension (win rows then lose rows): the two policy forwards become one of
    batch 2B, likewise the two frozen-reference forwards.  Per-sample results are unchanged (the UNet has no
    cross-sample operation); the reference runs them separately (turbo :775-787)."""
    cat2 = lambda a, b: torch.cat([a, b], dim=0).contiguous()
    out = dict(batch)
    out["input_latents_01"] = cat2(batch["input_latents_0"], batch["input_latents_1"])
    for k in ("prompt_embeds", "text_embeds", "time_ids", "timesteps"):
        out[k + "_01"] = cat2(batch[k], batch[k])
    return out


def product_micro_step_batched(pso, lora, unet, batch, sched, *, beta=50.0, eps=0.1, loss_scale=1.0, ref_stream=None,
                               kind="turbo", return_stats=False):
    """Same micro-step with ONE policy forward and ONE frozen-reference forward of batch 2B (``batched_view``).
    With ``ref_stream`` the no-grad reference forward is issued on a second CUDA stream: it is independent of the policy
    forward, and most kernels of a batch-8 forward leave SMs idle, so the two overlap (inside a captured CUDA graph the
    fork / join become graph edges)."""
    B = batch["latents_0"].shape[0]
    cond = {"text_embeds": batch["text_embeds_01"], "time_ids": batch["time_ids_01"].to(batch["input_latents_01"].dtype)}
    cur = torch.cuda.current_stream()
    if ref_stream is not None:
        ref_stream.wait_stream(cur)
        lora.disable_adapters(unet)
        with torch.cuda.stream(ref_stream), torch.no_grad():
            ref = unet(batch["input_latents_01"], batch["timesteps_01"], batch["prompt_embeds_01"], added_cond_kwargs=cond).sample
        lora.enable_adapters(unet)
    pol = unet(batch["input_latents_01"], batch["timesteps_01"], batch["prompt_embeds_01"], added_cond_kwargs=cond).sample
    if ref_stream is not None:
        cur.wait_stream(ref_stream)
        ref.record_stream(cur)
    else:
        lora.disable_adapters(unet)
        with torch.no_grad():
            ref = unet(batch["input_latents_01"], batch["timesteps_01"], batch["prompt_embeds_01"], added_cond_kwargs=cond).sample
        lora.enable_adapters(unet)
    ts = batch["timesteps"]
    out = pso.pso_pair_loss(pol[:B], pol[B:], ref[:B], ref[B:], batch["latents_0"], batch["latents_1"],
                            batch["next_latents_0"], batch["next_latents_1"], ts, ts, batch["human_prefer"],
                            scheduler=sched, beta=beta, eps=eps, loss_scale=loss_scale, return_stats=return_stats,
                            **_loss_kind(kind))
    loss = out[0] if return_stats else out
    loss.backward()
    return out


def oracle_micro_step(olora, olosses, unet, batch, sched, *, beta=50.0, eps=0.1, loss_scale=1.0, kind="turbo"):
    """The reference's flow restated with the oracle pieces (CPU): same four forwards, four step-with-logprob calls,
    inline loss, autograd backward."""
    ts = batch["timesteps"]
    p0 = _unet_call(unet, batch["input_latents_0"], ts, batch)
    p1 = _unet_call(unet, batch["input_latents_1"], ts, batch)
    olora.oracle_set_adapters(unet, False)
    with torch.no_grad():
        r0 = _unet_call(unet, batch["input_latents_0"], ts, batch)
        r1 = _unet_call(unet, batch["input_latents_1"], ts, batch)
    olora.oracle_set_adapters(unet, True)
    loss, _ = olosses.online_micro_step(kind, sched, [p0.float(), p1.float()], [r0.float(), r1.float()],
                                        [batch["latents_0"].float(), batch["latents_1"].float()],
                                        [batch["next_latents_0"].float(), batch["next_latents_1"].float()], [ts, ts],
                                        batch["human_prefer"], beta, eps,
                                        step_ratio=DMD_STEP_RATIO if kind == "dmd" else None)
    (loss * loss_scale).backward()
    return loss * loss_scale


# ----------------------------------------------------------------------------------------------- DreamBooth-style PSO (config 4)
DB_LEVELS = 4  # args.distill_train_timesteps: the 4-level schedule 249 / 499 / 749 / 999 (dreambooth trainer :1769-1779)


def synth_dreambooth_batch(b, latent_hw, cross_dim, pooled_dim, seed, sched, dtype=torch.float32, device="cpu"):
    """Inputs of one DreamBooth-PSO step as the trainer builds them (train_pso_sdxl_turbo_dreambooth.py:1727-1804): 2b rows
    -- b win latents (instance images) then b lose latents (negatives sampled from the base model) --, ONE noise draw shared by
    win and lose (:1763), one of 4 timestep levels per pair (:1769-1781), sigma looked up by timestep equality (:1675-1685),
    ``noisy = x0 + sigma noise`` (EulerDiscrete.add_noise, :1787), UNet input ``noisy / sqrt(sigma^2 + 1)`` (:1796), the
    instance prompt's embeddings repeated for both halves (:1813-1817).  ``sched`` = oracle.schedules.dreambooth_scheduler()
    or any object with ``.timesteps`` / ``.sigmas``."""
    g = torch.Generator().manual_seed(seed)
    shape = (4, latent_hw, latent_hw)
    x0 = torch.randn(2 * b, *shape, generator=g) * 0.8                                   # VAE latents * scaling_factor
    noise = torch.randn(2 * b, *shape, generator=g).chunk(2)[0].repeat(2, 1, 1, 1)       # :1763
    raw = torch.randint(0, 1000, (b,), generator=g)
    stride = 1000 // DB_LEVELS
    idx = (stride * (raw % DB_LEVELS) + stride - 1).long().repeat(2)                     # :1769-1779
    timesteps = sched.timesteps[idx]                                                     # :1781
    sig = torch.stack([sched.sigmas[(sched.timesteps == t).nonzero().item()] for t in timesteps]).reshape(-1, 1, 1, 1)
    noisy = x0 + sig * noise
    inp = noisy / ((sig ** 2 + 1) ** 0.5)
    px = 8.0 * latent_hw
    out = {"model_input": x0, "noisy_model_input": noisy, "inp_noisy_latents": inp, "sigmas": sig.float(),
           "timesteps": timesteps.float(),
           "prompt_embeds": torch.randn(b, 77, cross_dim, generator=g).repeat(2, 1, 1),
           "text_embeds": torch.randn(b, pooled_dim, generator=g).repeat(2, 1),
           "time_ids": torch.tensor([[px, px, 0.0, 0.0, px, px]]).repeat(2 * b, 1)}
    keep32 = ("sigmas", "timesteps", "time_ids")
    return {k: (v.to(device=device, dtype=dtype) if k not in keep32 else v.to(device)) for k, v in out.items()}


def _db_unet(unet, batch):
    x = batch["inp_noisy_latents"]
    return unet(x, batch["timesteps"], batch["prompt_embeds"],
                added_cond_kwargs={"text_embeds": batch["text_embeds"], "time_ids": batch["time_ids"].to(x.dtype)},
                return_dict=False)[0]


def product_dreambooth_micro_step(pso, lora, unet, batch, *, loss_type="pso", beta_pso=5.0, neg_defactor=0.1,
                                  prior_loss_weight=0.5, loss_scale=1.0, ref_stream=None, return_stats=False):
    """Product path of train_pso_sdxl_turbo_dreambooth.py:1812-1953: ONE policy forward of 2b rows with grad, for
    ``loss_type="pso"`` one adapter-disabled forward without (:1894-1905), then the fused DreamBooth-PSO kernel (EDM output
    preconditioning :1855, sigma^-2 weighting :1865, win / lose split, reference branch, hinge / log-sigmoid, prior term :1932)
    and the backward.  With ``ref_stream`` the frozen-reference forward runs on a second stream like the online step."""
    ref = None
    cur = torch.cuda.current_stream()
    if loss_type == "pso" and ref_stream is not None:
        ref_stream.wait_stream(cur)
        lora.disable_adapters(unet)
        with torch.cuda.stream(ref_stream), torch.no_grad():
            ref = _db_unet(unet, batch)
        lora.enable_adapters(unet)
    pred = _db_unet(unet, batch)
    if loss_type == "pso":
        if ref_stream is not None:
            cur.wait_stream(ref_stream)
            ref.record_stream(cur)
        else:
            lora.disable_adapters(unet)
            with torch.no_grad():
                ref = _db_unet(unet, batch)
            lora.enable_adapters(unet)
    loss, lw, ll, logits = pso.pso_db_loss(pred, ref, batch["noisy_model_input"], batch["model_input"], batch["sigmas"],
                                           loss_type=loss_type, beta_pso=beta_pso, neg_defactor=neg_defactor,
                                           prior_loss_weight=prior_loss_weight, loss_scale=loss_scale)
    loss.backward()
    return (loss, (lw, ll, logits)) if return_stats else loss


def oracle_dreambooth_micro_step(olora, olosses, unet, batch, *, loss_type="pso", beta_pso=5.0, neg_defactor=0.1,
                                 prior_loss_weight=0.5, loss_scale=1.0, loss_fn=None):
    """The trainer's flow restated with the oracle pieces (:1812-1953): policy forward, the reference forward with the
    adapters disabled, the loss lines (``oracle.losses.dreambooth_pso_loss``, or ``loss_fn`` = the trainer's own lines from
    ``reference_loader.trainer_dreambooth_block()``), autograd backward."""
    pred = _db_unet(unet, batch)
    ref = None
    if loss_type == "pso":
        olora.oracle_set_adapters(unet, False)
        with torch.no_grad():
            ref = _db_unet(unet, batch)
        olora.oracle_set_adapters(unet, True)
    fn = loss_fn or olosses.dreambooth_pso_loss
    loss, lw, ll, logits = fn(pred.float(), None if ref is None else ref.float(), batch["noisy_model_input"].float(),
                              batch["model_input"].float(), batch["sigmas"].float(), loss_type, beta_pso, neg_defactor,
                              prior_loss_weight)
    (loss * loss_scale).backward()
    return loss * loss_scale, (lw, ll, logits)
