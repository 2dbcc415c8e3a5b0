"""Reward-side preprocessing, CPU side: the oracle (oracle/clip_preprocess.py) against the committed Pillow-generated fixtures
and, where the imports exist, against Pillow / transformers themselves; the library's HOST plan functions (no GPU needed)
against the oracle bit for bit."""
import os

import numpy as np
import pytest

from oracle import clip_preprocess as ocp, make_golden_clip

SMALL = ("sq96", "wide", "tall", "up")


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "clip_preprocess.npz"))


def test_oracle_matches_the_pillow_generated_fixtures(golden):
    for name in SMALL:
        got = ocp.clip_preprocess([golden[f"{name}_image"]], size=32, crop=32)[0]
        assert got.dtype == np.float32 and np.array_equal(got, golden[f"{name}_pixel_values"]), name
    h, w, seed = (int(v) for v in golden["train512_seed"])
    img = make_golden_clip.synth_image(h, w, seed)  # the trainers' case: 512 x 512 -> 224 x 224 (turbo :626-640)
    assert np.array_equal(ocp.clip_preprocess([img])[0], golden["train512_pixel_values"])


def test_oracle_resize_is_pillow_bicubic_bit_for_bit():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for (h, w, oh, ow) in [(512, 512, 224, 224), (96, 128, 37, 49), (64, 64, 224, 224), (100, 37, 50, 19), (33, 1, 7, 5)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.array(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
        assert np.array_equal(ocp.pil_resize_u8(img, oh, ow), want), (h, w, oh, ow)


def test_oracle_matches_the_pil_backed_transformers_processor():
    Image = pytest.importorskip("PIL.Image")
    try:
        from transformers.models.clip import CLIPImageProcessorPil
    except Exception:
        pytest.skip("this transformers has no PIL-backed CLIP processor")
    rng = np.random.default_rng(1)
    p = CLIPImageProcessorPil()
    for shape in [(512, 512, 3), (300, 400, 3), (250, 224, 3)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        want = p(images=[Image.fromarray(img)], return_tensors="np")["pixel_values"]
        assert np.array_equal(ocp.clip_preprocess([img]), want), shape


def test_quantisation_line_of_the_trainer():
    import torch
    x = torch.tensor([-1.5, -1.0, -0.999, 0.0, 0.0039, 0.5, 0.9961, 1.0, 1.2]).reshape(1, 1, 3, 3).repeat(1, 3, 1, 1)
    q = ocp.quantize_images(x)
    assert q.shape == (1, 3, 3, 3) and q.dtype == np.uint8
    assert q[0, :, :, 0].reshape(-1).tolist() == [0, 0, 0, 127, 127, 191, 254, 255, 255]  # truncation, clamp


def test_library_host_plans_match_the_oracle(built_lib):
    from pairwise_sample_optimization_b200 import reward_preprocess as rp
    for a, b in [(512, 224), (1024, 224), (300, 168), (64, 224), (224, 224), (37, 19), (1, 5), (7, 1)]:
        bo, co = ocp.resample_plan(a, b)
        bc, cc = rp.resample_plan(a, b)
        assert np.array_equal(bo, bc) and np.array_equal(co, cc), (a, b)
        assert (cc.sum(1) - (1 << 22)).__abs__().max() <= cc.shape[1]  # weights sum to 1 up to rounding
    assert np.array_equal(ocp.norm_table(), rp.norm_table(1 / 255, ocp.OPENAI_CLIP_MEAN, ocp.OPENAI_CLIP_STD))
    assert rp.resize_output_size(512, 512, 224) == (224, 224) == ocp.resize_output_size(512, 512, 224)
    assert rp.resize_output_size(300, 400, 224) == (224, 298) == ocp.resize_output_size(300, 400, 224)
    assert rp.resize_output_size(400, 300, 224) == (298, 224)
    import torch
    with pytest.raises(Exception, match="CUDA"):
        rp.clip_image_preprocess(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))  # no CPU fallback
