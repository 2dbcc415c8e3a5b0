#!/usr/bin/env python
"""Which ingredient of the bench-timed configuration changes the gradient?  One full-size dmd128 micro-step (B pairs) under
variants of (in-launch deps, wgrad side stream, reference stream, CUDA graph), each compared with the eager one-stream run."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pairwise_sample_optimization_b200 as pso  # noqa: E402
from fixtures import micro_step, sdxl_unet  # noqa: E402
from oracle import schedules  # noqa: E402
from pairwise_sample_optimization_b200 import feed_forward, lora  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
fuse = "--no-fuse" not in sys.argv
torch.manual_seed(1234)
cfg = sdxl_unet.sdxl_config()
with torch.device("cuda"):
    unet = sdxl_unet.UNet2DConditionModel(cfg)
unet = unet.to(torch.bfloat16).requires_grad_(False)
wrapped = lora.add_adapter(unet, lora.LoraConfig(r=64, lora_alpha=64))
for m in wrapped:
    torch.nn.init.normal_(m.lora_B["default"].weight, std=0.01)
unet.set_attn_processor(lora.PSOAttnProcessor2_0())
if fuse:
    lora.fuse_attention_projections(unet)
feed_forward.install_fused_geglu(unet)
unet.train()
opt = lora.FusedLoRAOptimizer(unet)
sched = schedules.dmd_scheduler()
pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
host = micro_step.synth_batch(B, 128, cfg.cross_attention_dim, pooled, 100, None, dtype=torch.bfloat16, kind="dmd")
d = micro_step.batched_view({k: v.cuda() for k, v in host.items()})
kw = dict(beta=50.0, eps=0.1, kind="dmd")


def run(deps, wgrad, ref, graph):
    lora.set_in_launch_dependencies(deps)
    lora.set_wgrad_stream(wgrad)
    rs = torch.cuda.Stream() if ref else None
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            micro_step.product_micro_step_batched(pso, lora, unet, d, sched, ref_stream=rs, **kw)
        torch.cuda.synchronize()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg, stream=side):
            loss = micro_step.product_micro_step_batched(pso, lora, unet, d, sched, ref_stream=rs, **kw)
        opt.bucket.zero_()
        cg.replay()
    else:
        opt.bucket.zero_()
        loss = micro_step.product_micro_step_batched(pso, lora, unet, d, sched, ref_stream=rs, **kw)
    torch.cuda.synchronize()
    pso.check_status()
    return float(loss.item()), opt.bucket.flat.double().clone()


base_loss, base = run(True, False, False, False)
print(f"baseline: deps on, one stream, eager: loss {base_loss:.8f} |g| {base.norm().item():.6e}")
for name, cfg_ in (("same again", (True, False, False, False)), ("deps off", (False, False, False, False)),
                   ("wgrad side stream", (True, True, False, False)), ("reference stream", (True, False, True, False)),
                   ("graph only", (True, False, False, True)), ("wgrad + ref + graph", (True, True, True, True)),
                   ("deps off, wgrad + ref + graph", (False, True, True, True))):
    loss, g = run(*cfg_)
    cos = (torch.dot(base, g) / (base.norm() * g.norm())).item()
    worst = ((base - g).abs().max() / base.abs().max()).item()
    print(f"{name:32s} loss {loss:.8f} (== {loss == base_loss})  cos {cos:.9f}  worst {worst:.3e}  |g| {g.norm().item():.6e}")
lora.set_wgrad_stream(False)
lora.set_in_launch_dependencies(True)
