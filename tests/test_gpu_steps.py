"""GPU parity: the drop-in step / sampler functions (same names and signatures as the reference's
pso_pytorch.diffusers_patch) vs fixtures produced by the reference itself and vs the oracle."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import losses as olosses, samplers as osamplers, schedules, steps as osteps
from tests import _util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pso(built_lib):
    import pairwise_sample_optimization_b200 as pso
    assert torch.cuda.is_available()
    return pso


def _c(a, dtype=None):
    return U.cuda(torch.from_numpy(np.asarray(a)), dtype)


def test_turbo_step_scoring_and_backward_vs_reference_fixture(pso, golden_dir):
    g = np.load(os.path.join(golden_dir, "turbo_step.npz"))
    sched = schedules.turbo_scheduler(4)
    e = _c(g["model_output_scoring"]).requires_grad_(True)
    prev, lp = pso.turbo_step_with_logprob(sched, e, _c(g["timesteps"]), _c(g["sample"]),
                                           prev_sample=_c(g["prev_sample_sampling"]))
    assert prev.dtype == e.dtype and lp.dtype == torch.float32
    np.testing.assert_allclose(lp.detach().cpu().numpy(), g["log_prob_scoring_fp64"], rtol=1e-6)
    np.testing.assert_allclose(lp.detach().cpu().numpy(), g["log_prob_scoring"], rtol=2e-5)
    (lp * _c(g["grad_weights"])).sum().backward()
    assert U.rel_max(e.grad, torch.from_numpy(g["grad_model_output"])) <= 2e-5
    pso.check_status()


def test_turbo_step_sampling_vs_reference_fixture(pso, golden_dir):
    from pairwise_sample_optimization_b200 import _lib, runtime, step_ops
    g = np.load(os.path.join(golden_dir, "turbo_step.npz"))
    sched = schedules.turbo_scheduler(4)
    ts = _c(g["timesteps"])
    sd = runtime.turbo_schedule(sched, ts.device, _lib.ts_dtype_code(ts))
    lp, prev, scaled = step_ops.step_forward(sd, _c(g["model_output_sampling"]), _c(g["sample"]), ts,
                                             noise=_c(g["noise"]), want_scaled_next=True)
    np.testing.assert_allclose(prev.cpu().numpy(), g["prev_sample_sampling"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(lp.cpu().numpy(), g["log_prob_sampling"], rtol=2e-5)
    sig_next = torch.from_numpy(g["sigmas"])[1:4].reshape(-1, 1, 1, 1)
    want_scaled = torch.from_numpy(g["prev_sample_sampling"]) / ((sig_next ** 2 + 1) ** 0.5)  # TP:121
    np.testing.assert_allclose(scaled.cpu().numpy(), want_scaled.numpy(), rtol=2e-6, atol=2e-6)
    # the public function draws its own noise like TS:97: same generator state -> same draw as torch.randn
    gen = torch.Generator(device="cuda").manual_seed(3)
    e = _c(g["model_output_sampling"])
    prev2, lp2 = pso.turbo_step_with_logprob(sched, e, ts, _c(g["sample"]), generator=gen)
    noise = torch.randn(e.shape, dtype=e.dtype, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    want, want_lp = osteps.turbo_step(sched, e.cpu(), ts.cpu(), torch.from_numpy(g["sample"]), noise=noise.cpu())
    np.testing.assert_allclose(prev2.cpu().numpy(), want.numpy(), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(lp2.cpu().numpy(), want_lp.numpy(), rtol=2e-5)


def test_distilled_step_vs_reference_fixture(pso, golden_dir):
    from pairwise_sample_optimization_b200 import _lib, runtime, step_ops
    g = np.load(os.path.join(golden_dir, "dmd_step.npz"))
    sched = schedules.dmd_scheduler()
    t, tp = _c(g["timesteps"]), _c(g["prev_timesteps"])
    e = _c(g["model_output_scoring"]).requires_grad_(True)
    prev, lp = pso.distilled_step_with_logprob(sched, e, t, tp, _c(g["sample"]), prev_sample=_c(g["prev_sample_sampling"]))
    assert sched.alphas_cumprod.is_cuda  # DS:98 mutates the scheduler the same way
    np.testing.assert_allclose(lp.detach().cpu().numpy(), g["log_prob_scoring_fp64"], rtol=1e-6)
    np.testing.assert_allclose(lp.detach().cpu().numpy(), g["log_prob_scoring"], rtol=2e-5)
    (lp * _c(g["grad_weights"])).sum().backward()
    assert U.rel_max(e.grad, torch.from_numpy(g["grad_model_output"])) <= 2e-5
    sd = runtime.dmd_schedule(sched, t.device, _lib.ts_dtype_code(t))
    lp_s, prev_s, _ = step_ops.step_forward(sd, _c(g["model_output_sampling"]), _c(g["sample"]), t, tp, noise=_c(g["noise"]))
    np.testing.assert_allclose(prev_s.cpu().numpy(), g["prev_sample_sampling"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(lp_s.cpu().numpy(), g["log_prob_sampling"], rtol=2e-5)
    x0 = pso._get_x0_from_noise(_c(g["sample"]), _c(g["model_output_sampling"]), sched.alphas_cumprod,
                                torch.tensor([249, 249, 249], device="cuda"))
    np.testing.assert_allclose(x0.cpu().numpy(), g["x0_last_step"], rtol=2e-6, atol=2e-6)
    # the reference's expression promotes with the fp32 [B,1,1,1] alphas_cumprod gather (DS:36-42): 16-bit latents and
    # predictions still give an fp32 x0 -- the final DMD2 latent is not rounded to bf16 before the VAE
    x0h = pso._get_x0_from_noise(_c(g["sample"]).bfloat16(), _c(g["model_output_sampling"]).bfloat16(), sched.alphas_cumprod,
                                 torch.tensor([249, 249, 249], device="cuda"))
    assert x0h.dtype == torch.float32
    want = osteps.x0_from_noise(torch.from_numpy(g["sample"]).bfloat16(), torch.from_numpy(g["model_output_sampling"]).bfloat16(),
                               sched.alphas_cumprod.cpu(), torch.tensor([249] * 3))
    assert want.dtype == torch.float32
    np.testing.assert_allclose(x0h.cpu().numpy(), want.numpy(), rtol=2e-6, atol=2e-6)
    with pytest.raises(ValueError):  # DS:115-119
        pso.distilled_step_with_logprob(sched, e, t, tp, _c(g["sample"]), generator=torch.Generator(device="cuda"),
                                        prev_sample=prev)
    pso.check_status()


@pytest.mark.parametrize("kind", ["turbo", "dmd"])
@pytest.mark.parametrize("pd,ld", [("fp32", "fp32"), ("fp32", "fp16"), ("bf16", "bf16")])
def test_unmodified_trainer_flow_matches_fused_and_oracle(pso, kind, pd, ld):
    """The trainers' own code shape: four step calls + the inline loss in torch + backward
    (T:810-857) on the drop-in step functions, vs the fused kernel, vs the fp64 oracle."""
    shape = (4, 64, 64) if kind == "turbo" else (4, 128, 128)
    d = U.synth(kind, 2, shape, 21, 0.02, 0, U.DT[pd], U.DT[ld])
    cf = U.oracle_fp64(d)
    pred = [U.cuda(p).requires_grad_(True) for p in d["noise_pred"]]
    lps = []
    for k in (0, 1):
        for mo in (pred[k], U.cuda(d["noise_ref_pred"][k])):
            if kind == "turbo":
                _, lp = pso.turbo_step_with_logprob(d["sched"], model_output=mo, timestep=U.cuda(d["timesteps"][k]),
                                                    sample=U.cuda(d["latents"][k]), prev_sample=U.cuda(d["next_latents"][k]))
            else:
                ts = U.cuda(d["timesteps"][k])
                _, lp = pso.distilled_step_with_logprob(d["sched"], model_output=mo, timestep=ts, prev_timestep=ts - 250,
                                                        sample=U.cuda(d["latents"][k]), prev_sample=U.cuda(d["next_latents"][k]))
            lps.append(lp)
    loss = olosses.online_pso_loss(lps[0], lps[1], lps[2], lps[3], U.cuda(d["human_prefer"]), 50.0, 0.1)  # T:844-850 verbatim
    loss.backward()
    # the inline fp32 exp/log round trip carries the reference's own noise (SURVEY finding 4)
    assert abs(loss.item() - cf["loss"].item()) <= 3e-5 * cf["loss"].item()
    tol = 2e-4 if pd == "fp32" else 8e-3
    assert U.rel_max(pred[0].grad, cf["grads"][0]) <= tol and U.rel_max(pred[1].grad, cf["grads"][1]) <= tol
    fl, _, g0, g1 = U.run_fused(pso, d)
    assert abs(fl.item() - cf["loss"].item()) <= 1e-5 * cf["loss"].item()
    assert U.rel_max(pred[0].grad, g0) <= tol


class _ToyUNet(torch.nn.Module):
    """Stand-in for the UNet call sites TP:126-132 / DP:117-122 (the UNet itself is out of scope).  Purely
    elementwise fp32 so that the GPU (under the pipelines' autocast region) and the CPU oracle agree to rounding."""

    def __init__(self):
        super().__init__()
        self.config = types.SimpleNamespace(in_channels=4)
        self.w = torch.nn.Parameter(torch.tensor([0.7, -0.4, 0.9, 0.3]).reshape(1, 4, 1, 1))
        self.b = torch.nn.Parameter(torch.tensor([0.1, 0.0, -0.2, 0.05]).reshape(1, 4, 1, 1))

    def forward(self, x, t, encoder_hidden_states=None, added_cond_kwargs=None, return_dict=True):
        t = torch.as_tensor(t, device=x.device, dtype=torch.float32).reshape(-1, 1, 1, 1) / 1000.0
        y = torch.tanh(x.float() * self.w + self.b) * (1.0 + t)
        return (y,) if not return_dict else types.SimpleNamespace(sample=y)


class _Acc:
    @staticmethod
    def unwrap_model(m):
        return m


def test_turbo_sampler_pipeline_vs_oracle(pso):
    torch.manual_seed(0)
    unet = _ToyUNet().cuda()
    sched = schedules.turbo_scheduler(4)
    B = 2
    emb = torch.zeros(B, 77, 8, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(7)
    image, all_lat, all_lp, all_in = pso.sdxl_turbo_pipeline_with_logprob(
        _Acc, None, unet, sched, 512, 512, num_inference_steps=4, generator=gen, prompt_embeds=emb,
        pooled_prompt_embeds=emb[:, 0], add_time_ids=emb[:, 0, :6], output_type="latent")
    assert len(all_lat) == 4 and len(all_lp) == 3 and len(all_in) == 3  # TP:146-149 drops the last step
    # same draws, replayed for the oracle on the CPU
    gen2 = torch.Generator(device="cuda").manual_seed(7)
    lat0 = torch.randn(B, 4, 64, 64, generator=gen2, device="cuda", dtype=emb.dtype)
    noises = [torch.randn(B, 4, 64, 64, generator=gen2, device="cuda").cpu() for _ in range(4)]
    cpu_unet = _ToyUNet()
    cpu_unet.load_state_dict(unet.state_dict())
    with torch.no_grad():
        fin, o_lat, o_lp, o_in = osamplers.turbo_sampler(lambda x, t: cpu_unet(x, t, return_dict=False)[0], sched,
                                                         lat0.cpu(), noises, 4)
    for a, b in zip(all_lat, o_lat):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=1e-4, atol=2e-4)
    for a, b in zip(all_in, o_in):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=1e-4, atol=2e-4)
    for a, b in zip(all_lp, o_lp):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=1e-4)
    np.testing.assert_allclose(image.cpu().numpy(), fin.numpy(), rtol=1e-4, atol=2e-4)


def test_dmd_sampler_pipeline_vs_oracle(pso):
    torch.manual_seed(1)
    unet = _ToyUNet().cuda()
    sched = schedules.dmd_scheduler()
    ts, _ = schedules.dmd_distill_timesteps(4)
    B = 2
    emb = torch.zeros(B, 77, 8, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(8)
    image, all_lat, all_lp = pso.sdxl_dmd_pipeline_with_logprob(
        _Acc, None, unet, ts, sched, 1024, 1024, num_inference_steps=4, generator=gen, prompt_embeds=emb,
        pooled_prompt_embeds=emb[:, 0], add_time_ids=emb[:, 0, :6], output_type="latent")
    assert len(all_lat) == 5 and len(all_lp) == 3
    gen2 = torch.Generator(device="cuda").manual_seed(8)
    lat0 = torch.randn(B, 4, 128, 128, generator=gen2, device="cuda", dtype=emb.dtype)
    noises = [torch.randn(1, 4, 128, 128, generator=gen2, device="cuda").cpu() for _ in range(3)]  # DS:123: shared draw
    cpu_unet = _ToyUNet()
    cpu_unet.load_state_dict(unet.state_dict())
    with torch.no_grad():
        x0, o_lat, o_lp = osamplers.dmd_sampler(lambda x, t: cpu_unet(x, t, return_dict=False)[0],
                                                schedules.dmd_scheduler(), ts, lat0.cpu(), noises)
    for a, b in zip(all_lat, o_lat):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=1e-4, atol=2e-4)
    for a, b in zip(all_lp, o_lp):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=1e-4)
    np.testing.assert_allclose(image.cpu().numpy(), x0.numpy(), rtol=1e-4, atol=2e-4)


def test_broadcast_timestep_and_last_turbo_step(pso):
    """TP:139 passes ``t.unsqueeze(0)`` for the whole batch; the last step has sigma_up = 0 and the
    reference's log-prob is non-finite there (it is discarded, TP:146)."""
    sched = schedules.turbo_scheduler(4)
    x = torch.randn(3, 4, 64, 64, device="cuda") * 4
    e = torch.randn(3, 4, 64, 64, device="cuda")
    prev, lp = pso.turbo_step_with_logprob(sched, e, sched.timesteps[1].unsqueeze(0).cuda(), x,
                                           generator=torch.Generator(device="cuda").manual_seed(0))
    want, wlp = osteps.turbo_step(sched, e.cpu(), sched.timesteps[1].unsqueeze(0), x.cpu(),
                                  noise=torch.randn(e.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0)).cpu())
    np.testing.assert_allclose(prev.cpu().numpy(), want.numpy(), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(lp.cpu().numpy(), wlp.numpy(), rtol=2e-5)
    prev, lp = pso.turbo_step_with_logprob(sched, e, sched.timesteps[3].unsqueeze(0).cuda(), x,
                                           generator=torch.Generator(device="cuda").manual_seed(0))
    assert not torch.isfinite(lp).any()
    want0 = x.cpu() + e.cpu() * (0.0 - sched.sigmas[3])  # sigma_down = sigma_to = 0: x0 prediction
    np.testing.assert_allclose(prev.cpu().numpy(), want0.numpy(), rtol=2e-6, atol=2e-6)
    pso.check_status()


@pytest.mark.parametrize("kind,B,shape,dtype", [("turbo", 3, (4, 64, 64), torch.float32), ("turbo", 2, (4, 64, 64), torch.bfloat16),
                                                ("dmd", 3, (4, 32, 32), torch.float32), ("turbo", 2, (3, 5, 12), torch.float32)])
def test_in_kernel_philox_noise_matches_the_oracle_stream(pso, kind, B, shape, dtype):
    """Throughput mode of the sampler (no generator): the N(0,1) draws are produced inside the step kernel.  With
    model_output = sample = 0 the update is x' = s * noise, so the draws can be read back and compared with oracle/philox.py
    (Philox4x32-10 + Box-Muller, pinned to the published known-answer vectors); the log-prob is that of the drawn sample."""
    from oracle import philox
    from pairwise_sample_optimization_b200 import _lib, runtime, step_ops
    n = int(np.prod(shape))
    z = torch.zeros((B,) + shape, device="cuda", dtype=dtype)
    seed, offset = 0x1234567887654321, 5
    if kind == "turbo":
        sched = schedules.turbo_scheduler(4)
        ts = torch.tensor([999, 749, 499][:B] + [999] * max(0, B - 3), device="cuda")
        sd = runtime.turbo_schedule(sched, z.device, _lib.ts_dtype_code(ts))
        lp, prev, scaled = step_ops.step_forward(sd, z, z, ts, philox=(seed, offset), out_dtype=dtype, want_scaled_next=True)
        idx = torch.tensor(osteps.turbo_step_indices(sched, ts.cpu()))
        _, _, s = schedules.turbo_coefficients(sched.sigmas, idx)
        rows = B
    else:
        sched = schedules.dmd_scheduler()
        ts = torch.full((B,), 749, device="cuda")
        sd = runtime.dmd_schedule(sched, z.device, _lib.ts_dtype_code(ts))
        lp, prev, _ = step_ops.step_forward(sd, z, z, ts, ts - 250, philox=(seed, offset), noise_rows=1, out_dtype=dtype)
        _, _, s = schedules.dmd_coefficients(sched.alphas_cumprod, ts.cpu(), ts.cpu() - 250)
        rows = 1
    pso.check_status()
    want = torch.from_numpy(philox.normal(rows * n, seed, offset)).reshape((rows,) + shape)
    if dtype != torch.float32:
        want = want.to(dtype).double()  # the reference draws randn(dtype=...): values representable in the tensor's type
    got = prev.double().cpu() / s.double().reshape(-1, 1, 1, 1)
    tol = 3e-6 if dtype == torch.float32 else 2.0 ** -8
    assert (got - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())
    if kind == "dmd":  # ONE draw shared by the whole batch (DS:123-124)
        assert torch.equal(prev[0], prev[1]) and torch.equal(prev[1], prev[2])
    # the log-prob is the log-prob of the sample it returned (scoring mode on the stored fp32 / rounded sample)
    lp_want = -(want ** 2).reshape(rows, -1).mean(1) / 2 - torch.log(s.double()) - 0.5 * np.log(2 * np.pi)
    assert (lp.double().cpu() - lp_want).abs().max().item() <= 2e-5 * lp_want.abs().max().item()
    # another offset -> another stream; same (seed, offset) -> the same draws, whatever the launch geometry
    _, prev2, _ = (step_ops.step_forward(sd, z, z, ts, philox=(seed, offset + 1), out_dtype=dtype) if kind == "turbo" else
                   step_ops.step_forward(sd, z, z, ts, ts - 250, philox=(seed, offset + 1), noise_rows=1, out_dtype=dtype))
    assert not torch.equal(prev2, prev)
    if kind == "turbo" and n % 8 == 0:
        _, prev3, _ = step_ops.step_forward(sd, z, z, ts, philox=(seed, offset), out_dtype=dtype, tune=(128, 2))
        assert torch.equal(prev3, prev)
    if kind == "turbo":  # a 4-draw group must not straddle two samples: per-sample noise needs N % 4 == 0
        z7 = torch.zeros(B, 3, 5, 7, device="cuda", dtype=dtype)
        with pytest.raises(Exception, match="shape"):
            step_ops.step_forward(sd, z7, z7, ts, philox=(seed, offset), out_dtype=dtype)


def test_pipelines_draw_in_kernel_without_a_generator(pso):
    """generator=None: the sampler loop draws inside the step kernel (runtime.set_sampler_noise("philox"), the default)."""
    from pairwise_sample_optimization_b200 import runtime, step_ops
    B = 2
    unet = _ToyUNet().cuda()
    sched = schedules.turbo_scheduler(4)
    emb = torch.randn(B, 77, 8, device="cuda")
    def run():
        torch.manual_seed(11)      # keys the in-kernel stream (and draws the initial latents)
        step_ops.reset_philox()    # ... together with the process-wide call counter
        return pso.sdxl_turbo_pipeline_with_logprob(_Acc, None, unet, sched, 512, 512, num_inference_steps=4, prompt_embeds=emb,
                                                    pooled_prompt_embeds=emb[:, 0], add_time_ids=emb[:, 0, :6], output_type="latent")
    out1, out1b = run(), run()
    assert len(out1[1]) == 4 and all(torch.isfinite(t).all() for t in out1[1]) and all(torch.isfinite(t).all() for t in out1[2])
    assert all(torch.equal(a, b) for a, b in zip(out1[1], out1b[1]))  # same seed, same call sequence: the same trajectory
    assert not torch.equal(out1[1][1], out1[1][2])
    runtime.set_sampler_noise("torch")
    try:
        out2 = run()
    finally:
        runtime.set_sampler_noise("philox")
    assert not torch.equal(out2[1][1], out1[1][1])
    pso.check_status()
