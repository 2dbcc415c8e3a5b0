"""The callers either side of the hot path, end to end on the tiny fixture (examples/online_pso_tiny.py: sampling with the
drop-in pipeline -> preference signs -> per-timestep policy / frozen-reference forwards -> fused loss -> LoRA backward ->
fused optimizer boundary -> checkpoint), written like train_online_pso_sdxl_turbo.py:544-861."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_online_pso_example_trains_and_resumes(built_lib, tmp_path):
    import online_pso_tiny as ex
    import pairwise_sample_optimization_b200 as pso
    from pairwise_sample_optimization_b200 import checkpoint, lora

    history, unet, opt = ex.run(epochs=2, prompts=4, rank=8, seed=0, save_dir=str(tmp_path), verbose=False)
    assert len(history) == 2 and all(torch.isfinite(torch.tensor([h["loss"], h["grad_norm"], h["reward"]])).all() for h in history)
    assert history[0]["loss"] > 0 and history[-1]["steps"] == 2 and history[-1]["grad_norm"] > 0  # one optimizer boundary per epoch
    assert len(lora.projection_groups(unet)) > 0
    # overfitting ONE sampled batch must push the pairwise objective the right way: z = beta (h0 D0 + h1 D1) grows, loss drops
    sched = ex.euler_ancestral_schedule(4)
    cfg = unet.config
    s = ex.sample_epoch(unet, sched, cfg, 4)
    first = last = None
    for it in range(8):
        losses, z, hp = ex.train_epoch(unet, opt, sched, s, accum=1)
        if it == 0:
            first = float(losses.mean())
        last = float(losses.mean())
    pso.check_status()
    assert last < first - 1e-3, (first, last)
    # resume: a fresh model + optimizer loaded from the checkpoint of the example continue from the saved state
    _, unet2, opt2 = ex.build(rank=8, seed=123)
    checkpoint.load_state(str(tmp_path), unet2, opt2)
    assert int(opt2.step_dev) == 2 and opt2.exp_avg.abs().max() > 0
