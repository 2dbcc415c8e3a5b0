"""Restated PSO losses (TEST INFRASTRUCTURE, CPU).

The reference has no loss *function*: the expressions are inline in the trainers
(which cannot be imported here -- they need diffusers/peft/accelerate).  Each
function below restates the cited lines; torch autograd of these restatements is
the fp32 "reference-rounding" oracle, the ``*_closed_form`` functions are the fp64
truth parity is judged against.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import schedules, steps


# ----------------------------------------------------------------------------- preference signs
def sample_compare(a: torch.Tensor, b: torch.Tensor, generator=None, reward_indices=None):
    """train_online_pso_sdxl_turbo.py:401-416.  a, b: [bs, m] rewards; one random reward column per
    row; ``a <= b`` -> [-1, +1] (ties favour sample 1), else [+1, -1]."""
    bs, m = a.shape
    if reward_indices is None:
        reward_indices = torch.randint(0, m, (bs,), generator=generator)
    pa = a[torch.arange(bs), reward_indices]
    pb = b[torch.arange(bs), reward_indices]
    a_dom = pa <= pb
    b_dom = pb < pa
    c = torch.zeros([bs, 2], dtype=torch.float)
    c[a_dom] = torch.tensor([-1.0, 1.0])
    c[b_dom] = torch.tensor([1.0, -1.0])
    return c


def compare(a: torch.Tensor, b: torch.Tensor):
    """train_online_pso_sdxl_dmd2.py:420-434.  Pareto dominance; ties -> [0, 0]."""
    if a.ndim == 1:
        a = a[..., None]
        b = b[..., None]
    a_dom = torch.logical_and(torch.all(a <= b, dim=1), torch.any(a < b, dim=1))
    b_dom = torch.logical_and(torch.all(b <= a, dim=1), torch.any(b < a, dim=1))
    c = torch.zeros([a.shape[0], 2], dtype=torch.float)
    c[a_dom] = torch.tensor([-1.0, 1.0])
    c[b_dom] = torch.tensor([1.0, -1.0])
    return c


# ----------------------------------------------------------------------------- online PSO loss
def online_pso_loss(lp0, lpref0, lp1, lpref1, human_prefer, beta: float, eps: float):
    """train_online_pso_sdxl_turbo.py:844-850 (== train_online_pso_sdxl_dmd2.py:848-854)."""
    ratio_0 = torch.clamp(torch.exp(lp0 - lpref0), 1 - eps, 1 + eps)
    ratio_1 = torch.clamp(torch.exp(lp1 - lpref1), 1 - eps, 1 + eps)
    return -torch.log(torch.sigmoid(
        beta * (torch.log(ratio_0)) * human_prefer[:, 0] +
        beta * (torch.log(ratio_1)) * human_prefer[:, 1]
    )).mean()


def online_micro_step(kind, scheduler, noise_pred, noise_ref_pred, latents, next_latents, timesteps,
                      human_prefer, beta, eps, step_ratio=None, step_fns=None, upcast=True):
    """The loss part of one training micro-step: four step calls + inline loss
    (turbo trainer :810-850, dmd2 trainer :812-854).  Each tensor argument is a pair
    ``(branch0, branch1)``.  ``step_fns`` lets tests pass the verbatim reference functions.
    Returns (loss, [lp0, lpref0, lp1, lpref1])."""
    lps = []
    for k in (0, 1):
        for pred in (noise_pred[k], noise_ref_pred[k]):
            if kind == "turbo":
                fn = step_fns or steps.turbo_step
                kw = {"device": "cpu"} if step_fns else {}
                _, lp = fn(scheduler, model_output=pred, timestep=timesteps[k], sample=latents[k],
                           prev_sample=next_latents[k], **kw)
            else:
                fn = step_fns or steps.distilled_step
                kw = {"device": "cpu"} if step_fns else {"upcast": upcast}
                _, lp = fn(scheduler, model_output=pred, timestep=timesteps[k],
                           prev_timestep=timesteps[k] - step_ratio, sample=latents[k],
                           prev_sample=next_latents[k], **kw)
            lps.append(lp)
    loss = online_pso_loss(lps[0], lps[1], lps[2], lps[3], human_prefer, beta, eps)
    return loss, lps


def online_coefficients(kind, scheduler, timesteps, step_ratio=None):
    """(k, a, s) fp64 per sample for either scheduler family."""
    if kind == "turbo":
        idx = torch.tensor(steps.turbo_step_indices(scheduler, timesteps))
        return schedules.turbo_coefficients(scheduler.sigmas, idx)
    return schedules.dmd_coefficients(scheduler.alphas_cumprod, timesteps, timesteps - step_ratio)


def online_closed_form(kind, scheduler, noise_pred, noise_ref_pred, latents, next_latents, timesteps,
                       human_prefer, beta, eps, step_ratio=None, loss_scale=1.0):
    """fp64 closed form of loss and of d loss / d noise_pred_k.

    Returns dict(loss, grads=(g0, g1), logp=[lp0, lpref0, lp1, lpref1], delta=(d0, d1), z)."""
    B = noise_pred[0].shape[0]
    h = human_prefer.double()
    logp, delta, resid, coefs, gate = [], [], [], [], []
    for k in (0, 1):
        kk, aa, ss = online_coefficients(kind, scheduler, timesteps[k], step_ratio)
        lp, r = steps.affine_logprob_closed_form(noise_pred[k], latents[k], next_latents[k], kk, aa, ss)
        lpr, _ = steps.affine_logprob_closed_form(noise_ref_pred[k], latents[k], next_latents[k], kk, aa, ss)
        logp += [lp, lpr]
        d = lp - lpr
        delta.append(d)
        resid.append(r)
        coefs.append((kk, aa, ss))
        e = torch.exp(d)
        gate.append(((e >= 1 - eps) & (e <= 1 + eps)).double())  # torch.clamp passes grad on the closed interval
    logr = [torch.log(torch.clamp(torch.exp(d), 1 - eps, 1 + eps)) for d in delta]
    z = beta * (h[:, 0] * logr[0] + h[:, 1] * logr[1])
    loss = loss_scale * F.softplus(-z).mean()
    grads = []
    for k in (0, 1):
        kk, aa, ss = coefs[k]
        n = resid[k][0].numel()
        g = loss_scale * (-torch.sigmoid(-z) / B) * beta * h[:, k] * gate[k] * aa / (ss ** 2 * n)
        grads.append(g.reshape(B, *([1] * (resid[k].ndim - 1))) * resid[k])
    return dict(loss=loss, grads=tuple(grads), logp=logp, delta=tuple(delta), z=z)


# ----------------------------------------------------------------------------- DreamBooth PSO loss
def dreambooth_pso_loss(model_pred, ref_pred, noisy_model_input, model_input, sigmas, loss_type="pso",
                        beta_pso=1.0, neg_defactor=0.1, prior_loss_weight=0.0):
    """train_pso_sdxl_turbo_dreambooth.py:1847-1865 (EDM-style epsilon preconditioning, non-EDM
    scheduler) + :1881-1935.  Rows [0,b) are the win images, [b,2b) the lose images.
    ``model_pred`` / ``ref_pred`` are raw UNet outputs; ``sigmas`` is [2b,1,1,1].
    Returns (loss, model_losses_w, model_losses_l, logits)."""
    x0_pred = model_pred * (-sigmas) + noisy_model_input              # :1855
    weighting = (sigmas ** -2.0).float()                              # :1865
    target = model_input                                              # :1869
    model_losses = torch.mean(
        (weighting.float() * (x0_pred.float() - target.float()) ** 2).reshape(target.shape[0], -1), 1
    )                                                                 # :1885-1890
    losses_w, losses_l = model_losses.chunk(2)                        # :1891
    model_diff = losses_w - neg_defactor * losses_l                   # :1892
    if loss_type != "pso_db":                                         # :1894-1920
        with torch.no_grad():
            rp = ref_pred * (-sigmas) + noisy_model_input             # :1906
            ref_loss = torch.mean(
                (weighting.float() * (rp.float() - target.float()) ** 2).reshape(target.shape[0], -1), 1
            )
            ref_w, ref_l = ref_loss.chunk(2)
            ref_diff = ref_w - neg_defactor * ref_l
        logits = ref_diff - model_diff
    else:
        logits = -model_diff                                          # :1922
    if loss_type == "pso":
        loss = -1 * F.logsigmoid(beta_pso * logits).mean()            # :1925
    elif loss_type == "pso_db":
        loss = torch.relu(1 - beta_pso * logits).mean()               # :1927
    else:
        raise ValueError(f"Unknown loss type {loss_type}")            # :1929
    if prior_loss_weight > 0.0:                                       # :1932-1935
        loss = loss + prior_loss_weight * losses_l.mean()
    return loss, losses_w, losses_l, logits


def dreambooth_closed_form(model_pred, ref_pred, noisy_model_input, model_input, sigmas, loss_type="pso",
                           beta_pso=1.0, neg_defactor=0.1, prior_loss_weight=0.0, loss_scale=1.0):
    """fp64 closed form of the DreamBooth-PSO loss and its gradient into ``model_pred``
    (SURVEY App. A.4): grad = G_i * 2 w (x0_pred - x0)/N * (-sigma)."""
    mp, nz, x0, sg = (t.double() for t in (model_pred, noisy_model_input, model_input, sigmas))
    b2 = mp.shape[0]
    b = b2 // 2
    n = mp[0].numel()
    w = sg ** -2.0
    r = mp * (-sg) + nz - x0
    L = (w * r * r).reshape(b2, -1).mean(1)
    Lw, Ll = L[:b], L[b:]
    model_diff = Lw - neg_defactor * Ll
    if loss_type == "pso":
        rr = ref_pred.double() * (-sg) + nz - x0
        Lr = (w * rr * rr).reshape(b2, -1).mean(1)
        logits = (Lr[:b] - neg_defactor * Lr[b:]) - model_diff
        per = F.softplus(-beta_pso * logits)
        dl = -beta_pso * torch.sigmoid(-beta_pso * logits) / b        # d loss / d logits
    elif loss_type == "pso_db":
        logits = -model_diff
        per = torch.relu(1 - beta_pso * logits)
        dl = -beta_pso * ((1 - beta_pso * logits) > 0).double() / b
    else:
        raise ValueError(loss_type)
    loss = per.mean()
    lam = prior_loss_weight if prior_loss_weight > 0.0 else 0.0
    loss = loss + lam * Ll.mean()
    G = torch.cat([-dl, neg_defactor * dl + lam / b])                 # d loss / d L_i
    grad = (G.reshape(b2, 1, 1, 1) * 2 * w * r / n * (-sg)) * loss_scale
    return dict(loss=loss * loss_scale, grad=grad, losses_w=Lw, losses_l=Ll, logits=logits)
