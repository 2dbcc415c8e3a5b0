"""BASELINE config 1: PSO loss + LoRA step on the tiny random-init SDXL-architecture UNet (32/64 channels), 2 pairs,
64x64 latents.  The product path (bf16 UNet on the GPU, tcgen05 LoRA projections, fused loss+grad kernel) against the
reference's flow restated on the CPU in fp32 (oracle LoRA module, four step-with-logprob calls, inline loss, autograd).

Tolerance: the GPU run carries bf16 activations through ~40 layers, the oracle none, so this end-to-end check is
statistical (loss 2e-2 relative, every adapter gradient: cosine >= 0.98 over the flat vector and 1e-1 of max|grad|);
the bit-level statements live in test_gpu_lora*.py and test_gpu_pair_loss.py."""
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


def _build(lora, r, seed, gradient_checkpointing):
    torch.manual_seed(seed)
    cfg = sdxl_unet.tiny_config()
    base = sdxl_unet.UNet2DConditionModel(cfg).to(torch.bfloat16)  # weights rounded once; both arms share the values
    cpu = copy.deepcopy(base).float()
    gpu = copy.deepcopy(base).cuda()
    wrapped_o = olora.oracle_add_adapter(cpu, r, r)
    wrapped_g = lora.add_adapter(gpu, lora.LoraConfig(r=r, lora_alpha=r))
    assert len(wrapped_o) == len(wrapped_g) == sdxl_unet.count_lora_targets(base) == 96
    g = torch.Generator().manual_seed(seed + 1)
    for mo, mg in zip(wrapped_o, wrapped_g):
        A = torch.randn(mo.lora_A["default"].weight.shape, generator=g) * (1.0 / r)
        Bm = torch.randn(mo.lora_B["default"].weight.shape, generator=g) * 0.03
        A, Bm = A.bfloat16().float(), Bm.bfloat16().float()  # values both arms can represent
        with torch.no_grad():
            mo.lora_A["default"].weight.copy_(A); mo.lora_B["default"].weight.copy_(Bm)
            mg.lora_A["default"].weight.copy_(A); mg.lora_B["default"].weight.copy_(Bm)
    cpu.train(); gpu.train()
    if gradient_checkpointing:
        gpu.enable_gradient_checkpointing()
    return cfg, cpu, gpu, wrapped_o, wrapped_g


@pytest.mark.parametrize("kind,gradient_checkpointing", [("turbo", False), ("turbo", True), ("dmd", True)])
def test_config1_tiny_unet_micro_step(mods, kind, gradient_checkpointing):
    pso, lora = mods
    r, B = 4, 2
    cfg, cpu, gpu, wo, wg = _build(lora, r, 0, gradient_checkpointing)
    sched = schedules.turbo_scheduler(4) if kind == "turbo" else schedules.dmd_scheduler()
    batch = micro_step.synth_batch(B, 64, cfg.cross_attention_dim, 32, 5, getattr(sched, "sigmas", None), kind=kind)
    to_gpu = {k: (v.cuda().bfloat16() if v.is_floating_point() and k not in ("human_prefer", "time_ids") else v.cuda())
              for k, v in batch.items()}
    # the oracle sees the same bf16-rounded inputs
    batch = {k: (v.bfloat16().float() if v.is_floating_point() and k not in ("human_prefer", "time_ids") else v)
             for k, v in batch.items()}
    bucket = lora.LoRAGradBucket(lora.lora_parameters(gpu))
    kw = dict(beta=5.0, eps=0.9, kind=kind)
    loss_g = micro_step.product_micro_step(pso, lora, gpu, to_gpu, sched, **kw)
    loss_o = micro_step.oracle_micro_step(olora, olosses, cpu, batch, sched, **kw)
    pso.check_status()
    assert abs(loss_g.item() - loss_o.item()) <= 2e-2 * abs(loss_o.item()), (loss_g.item(), loss_o.item())
    flat_o = torch.cat([torch.cat([m.lora_A["default"].weight.grad.flatten(), m.lora_B["default"].weight.grad.flatten()])
                        for m in wo]).double()
    flat_g = torch.cat([torch.cat([m.lora_A["default"].weight.grad.flatten(), m.lora_B["default"].weight.grad.flatten()])
                        for m in wg]).double().cpu()
    assert flat_o.abs().max() > 0, "vacuous: the clamp gate closed and every gradient is zero"
    cos = torch.dot(flat_o, flat_g) / (flat_o.norm() * flat_g.norm())
    assert cos.item() >= 0.98, cos.item()
    worst = (flat_o - flat_g).abs().max().item() / flat_o.abs().max().item()
    # the worst single element moves from run to run (0.040 - 0.054 observed over 12 runs: the order of the fp32 atomic
    # accumulation flips bf16 roundings that ~40 layers amplify); the cosine above is stable at 0.999
    assert worst <= 1e-1, worst
    # the bucket IS the gradients (one flat buffer, ready for a single all-reduce)
    assert abs(bucket.flat.double().norm().item() - flat_g.norm().item()) <= 1e-6 * flat_g.norm().item()
