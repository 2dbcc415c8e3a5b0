"""BASELINE config 4 end to end: the DreamBooth-style PSO micro-step (train_pso_sdxl_turbo_dreambooth.py:1720-1964) on the tiny
SDXL-architecture fixture -- 2b rows (b win + b lose), shared noise, 4-level timesteps, sigma lookup, EDM input / output
preconditioning, ``pso`` (with the adapter-disabled reference forward) and ``pso_db`` (hinge, no reference), prior term,
LoRA rank 4, b = 4 -- product path (bf16 UNet on the GPU, tcgen05 LoRA projections, fused DreamBooth loss+grad kernel) against
the trainer's flow restated on the CPU in fp32 (oracle LoRA module, ``oracle.losses.dreambooth_pso_loss`` = the trainer's loss
lines, pinned bit-identical to the verbatim lines by tests/test_oracle_golden.py).  Statistical tolerances as in
test_gpu_unet_step.py (bf16 activations through ~40 layers): loss 2e-2, gradient cosine >= 0.98, worst element 1e-1."""
import copy

import pytest
import torch

from fixtures import micro_step, sdxl_unet
from oracle import lora as olora, losses as olosses, schedules

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built_lib):
    import pairwise_sample_optimization_b200 as pso
    from pairwise_sample_optimization_b200 import lora
    return pso, lora


def _flat(layers):
    return torch.cat([torch.cat([m.lora_A["default"].weight.grad.flatten(), m.lora_B["default"].weight.grad.flatten()])
                      for m in layers]).double().cpu()


@pytest.mark.parametrize("loss_type,prior,graph", [("pso", 0.5, False), ("pso", 0.0, True), ("pso_db", 0.5, False)])
def test_config4_dreambooth_micro_step_vs_oracle(mods, loss_type, prior, graph):
    pso, lora = mods
    r, b = 4, 4
    torch.manual_seed(0)
    cfg = sdxl_unet.tiny_config()
    base = sdxl_unet.UNet2DConditionModel(cfg).to(torch.bfloat16)
    cpu, gpu = copy.deepcopy(base).float(), copy.deepcopy(base).cuda()
    wo = olora.oracle_add_adapter(cpu, r, r)
    wg = lora.add_adapter(gpu, lora.LoraConfig(r=r, lora_alpha=r))
    g = torch.Generator().manual_seed(1)
    for mo, mg in zip(wo, wg):
        A = (torch.randn(mo.lora_A["default"].weight.shape, generator=g) * (1.0 / r)).bfloat16().float()
        Bm = (torch.randn(mo.lora_B["default"].weight.shape, generator=g) * 0.03).bfloat16().float()
        with torch.no_grad():
            mo.lora_A["default"].weight.copy_(A); mo.lora_B["default"].weight.copy_(Bm)
            mg.lora_A["default"].weight.copy_(A); mg.lora_B["default"].weight.copy_(Bm)
    cpu.train(); gpu.train()
    gpu.set_attn_processor(lora.PSOAttnProcessor2_0())
    bucket = lora.LoRAGradBucket(lora.lora_parameters(gpu))
    sched = schedules.dreambooth_scheduler()
    batch = micro_step.synth_dreambooth_batch(b, 64, cfg.cross_attention_dim, 32, 7, sched)
    keep32 = ("sigmas", "timesteps", "time_ids")
    d = {k: (v.cuda().bfloat16() if k not in keep32 else v.cuda()) for k, v in batch.items()}
    batch = {k: (v.bfloat16().float() if k not in keep32 else v) for k, v in batch.items()}
    # the product's sync-free sigma lookup is the trainer's get_sigmas (:1675-1685)
    dev_sched = type(sched)(timesteps=sched.timesteps.cuda(), sigmas=sched.sigmas.cuda())
    assert torch.equal(pso.edm_sigmas(dev_sched, d["timesteps"]).cpu(), batch["sigmas"])
    kw = dict(loss_type=loss_type, beta_pso=5.0, neg_defactor=0.1, prior_loss_weight=prior)
    if graph:  # as bench.py --config dreambooth64 runs it: reference forward on a second stream, captured and replayed
        ref_stream, side = torch.cuda.Stream(), torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            micro_step.product_dreambooth_micro_step(pso, lora, gpu, d, ref_stream=ref_stream, **kw)
        torch.cuda.synchronize()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg, stream=side):
            loss_g, (lw, ll, logits) = micro_step.product_dreambooth_micro_step(pso, lora, gpu, d, ref_stream=ref_stream,
                                                                                return_stats=True, **kw)
        bucket.zero_()
        cg.replay()
        torch.cuda.synchronize()
    else:
        loss_g, (lw, ll, logits) = micro_step.product_dreambooth_micro_step(pso, lora, gpu, d, return_stats=True, **kw)
    pso.check_status()
    loss_o, (lw_o, ll_o, logits_o) = micro_step.oracle_dreambooth_micro_step(olora, olosses, cpu, batch, **kw)
    assert abs(loss_g.item() - loss_o.item()) <= 2e-2 * abs(loss_o.item()), (loss_g.item(), loss_o.item())
    for got, want in ((lw, lw_o), (ll, ll_o)):
        assert (got.cpu().double() - want.double()).abs().max().item() <= 2e-2 * want.abs().max().item()
    flat_o, flat_g = _flat(wo), _flat(wg)
    assert flat_o.abs().max() > 0, "vacuous: every gradient is zero"
    cos = torch.dot(flat_o, flat_g) / (flat_o.norm() * flat_g.norm())
    assert cos.item() >= 0.98, cos.item()
    worst = (flat_o - flat_g).abs().max().item() / flat_o.abs().max().item()
    assert worst <= 1e-1, worst
    assert abs(bucket.flat.double().norm().item() - flat_g.norm().item()) <= 1e-6 * flat_g.norm().item()
