"""psob200_clip_preprocess (one launch: quantise -> Pillow-exact bicubic resize -> centre crop -> rescale / normalise ->
channels first) against the CPU oracle and the Pillow-generated fixtures: BIT-EXACT (integer resize, table-driven float step)."""
import os

import numpy as np
import pytest
import torch

from oracle import clip_preprocess as ocp, make_golden_clip

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rp(built_lib):
    from pairwise_sample_optimization_b200 import reward_preprocess
    return reward_preprocess


def test_fixtures_bit_exact(rp, golden_dir):
    g = np.load(os.path.join(golden_dir, "clip_preprocess.npz"))
    for name in ("sq96", "wide", "tall", "up"):
        img = torch.from_numpy(g[f"{name}_image"]).cuda()[None]
        got = rp.clip_image_preprocess(img, size=32, crop_size=32)
        assert got.dtype == torch.float32 and np.array_equal(got[0].cpu().numpy(), g[f"{name}_pixel_values"]), name
    h, w, seed = (int(v) for v in g["train512_seed"])
    img = torch.from_numpy(make_golden_clip.synth_image(h, w, seed)).cuda()[None]
    assert np.array_equal(rp.clip_image_preprocess(img)[0].cpu().numpy(), g["train512_pixel_values"])


@pytest.mark.parametrize("B,h,w", [(4, 512, 512), (2, 1024, 1024), (3, 300, 400), (2, 400, 300), (1, 64, 64), (5, 224, 224),
                                    (2, 257, 511)])
def test_uint8_batches_vs_oracle(rp, B, h, w):
    rng = np.random.default_rng(h * 7 + w)
    imgs = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
    want = ocp.clip_preprocess(list(imgs))
    got = rp.clip_image_preprocess(torch.from_numpy(imgs).cuda())
    assert tuple(got.shape) == (B, 3, 224, 224) and np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_decoded_float_images_are_quantised_like_the_trainer(rp, dtype):
    """turbo :632-633: ((images + 1.0) * 127.5).clamp(0, 255).to(uint8) in the image's own dtype, then the processor."""
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(3, 3, 512, 512, generator=g) * 2.4 - 1.2).to(dtype)  # beyond [-1, 1] on both sides: exercises the clamp
    want = ocp.clip_preprocess(list(ocp.quantize_images(x)))
    got = rp.clip_image_preprocess(x.cuda())
    assert np.array_equal(got.cpu().numpy(), want)


def test_half_outputs_and_processor_object(rp):
    rng = np.random.default_rng(9)
    imgs = rng.integers(0, 256, (2, 128, 160, 3), dtype=np.uint8)
    want = torch.from_numpy(ocp.clip_preprocess(list(imgs), size=64, crop=64))
    for dt in (torch.float16, torch.bfloat16):
        proc = rp.DeviceCLIPImageProcessor(size=64, crop_size=64, out_dtype=dt)
        got = proc(images=torch.from_numpy(imgs).cuda(), return_tensors="pt")["pixel_values"]
        assert got.dtype == dt and torch.equal(got.cpu(), want.to(dt))
    with pytest.raises(Exception):
        rp.clip_image_preprocess(torch.zeros(1, 16, 16, 3, dtype=torch.uint8).cuda(), size=8, crop_size=16)
