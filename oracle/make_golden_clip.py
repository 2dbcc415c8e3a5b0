"""Generates tests/golden/clip_preprocess.npz with the REAL third-party code the reference's reward path runs on the host:
Pillow's ``Image.resize(BICUBIC)`` (what transformers==4.38.1 ``CLIPImageProcessor`` calls; pickscore_utils.py:24-33) followed
by the processor's centre crop / rescale / normalise, here through transformers' PIL-backed CLIP processor when the installed
version still ships one, else through the restated numpy steps with the Pillow resize.  Run in the build container:

    python -m oracle.make_golden_clip

Small cases keep full inputs and outputs; the 512x512 -> 224 case of the trainers keeps the (seeded) input recipe and the output.
"""
import os

import numpy as np

from . import clip_preprocess as ocp

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synth_image(h, w, seed):
    """A seeded image with structure at several scales (smooth gradients + texture + hard edges), uint8 [h, w, 3]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    for c in range(3):
        img[..., c] = 127 + 90 * np.sin(xx / (7.0 + 3 * c) + seed) * np.cos(yy / (11.0 - 2 * c)) + rng.normal(0, 25, (h, w))
    img[h // 4:h // 2, w // 3:2 * w // 3] = rng.integers(0, 256, 3)  # a flat block with hard edges
    return np.clip(img, 0, 255).astype(np.uint8)


def reference_pixel_values(img_u8):
    from PIL import Image
    try:
        from transformers.models.clip import CLIPImageProcessorPil
        return CLIPImageProcessorPil()(images=[Image.fromarray(img_u8)], return_tensors="np")["pixel_values"][0], "transformers-pil"
    except Exception:
        def pil_resize(img, oh, ow):
            return np.array(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
        return ocp.clip_preprocess([img_u8], resize_fn=pil_resize)[0], "pillow+restated"


def main():
    import PIL
    out = {}
    how = None
    for name, (h, w), seed in (("sq96", (96, 96), 1), ("wide", (60, 150), 2), ("tall", (131, 57), 3), ("up", (20, 27), 4),
                               ("train512", (512, 512), 5)):
        img = synth_image(h, w, seed)
        # the processor's size is fixed at 224 in the reference; the small cases are preprocessed at size 32 / crop 32 through the
        # SAME Pillow resize so that fixtures stay small
        if name == "train512":
            pv, how = reference_pixel_values(img)
            out[f"{name}_pixel_values"] = pv
        else:
            from PIL import Image

            def pil_resize(im, oh, ow):
                return np.array(Image.fromarray(im).resize((ow, oh), resample=Image.BICUBIC))
            out[f"{name}_image"] = img
            out[f"{name}_pixel_values"] = ocp.clip_preprocess([img], size=32, crop=32, resize_fn=pil_resize)[0]
        out[f"{name}_seed"] = np.array([h, w, seed])
    out["generated_with"] = np.array(f"Pillow {PIL.__version__}; 224-case through {how}")
    np.savez_compressed(os.path.join(OUT, "clip_preprocess.npz"), **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()}, os.path.getsize(os.path.join(OUT, "clip_preprocess.npz")))


if __name__ == "__main__":
    main()
