"""Reward-side image preprocessing restated on the CPU (TEST INFRASTRUCTURE; numpy integer / float arithmetic).

What the reference does before every reward call:

* train_online_pso_sdxl_turbo.py:632-633 -- ``((images + 1.0) * 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
  .cpu().numpy()`` then ``Image.fromarray`` (:638-640)                                        -> ``quantize_images``
* pso_pytorch/pickscore_utils.py:24-33 -- ``AutoProcessor.from_pretrained("laion/CLIP-ViT-H-14-laion2B-s32B-b79K")(images=...)``
  = transformers==4.38.1 (environment.yml) ``CLIPImageProcessor.preprocess``: resize the shortest edge to 224 with
  ``PIL.Image.resize(resample=BICUBIC)``, centre crop 224 x 224, ``image * (1/255)`` (float64 product, cast to float32),
  ``(image - mean) / std`` in float32 with the OpenAI CLIP statistics, channels first             -> ``clip_preprocess``

Third-party arithmetic, absent from /root/reference, restated from the published algorithms:

* Pillow ``libImaging/Resample.c`` (8 bits per channel): ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` + the horizontal
  then vertical fixed-point passes                                                               -> ``resample_plan`` / ``pil_resize_u8``
* transformers ``image_transforms.get_resize_output_image_size`` / ``center_crop`` / ``rescale`` / ``normalize``.

Pinning: tests/test_oracle_golden.py checks ``pil_resize_u8`` bit for bit against the Pillow installed in this image (the
reference's own dependency) and ``clip_preprocess`` against the installed transformers CLIPImageProcessor where those imports
exist; tests/golden/clip_preprocess.npz holds outputs generated here with Pillow (oracle/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def quantize_images(images):
    """train_online_pso_sdxl_turbo.py:632: float NCHW in [-1, 1] -> uint8 NHWC (torch arithmetic in the tensor's dtype)."""
    import torch
    return ((images + 1.0) * 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).cpu().numpy()


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_plan(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc over the whole axis -> (bounds[out,2] int32, coeffs[out,taps] int32)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coeffs = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        bounds[xx] = (xmin, xmax)
        for x, w in enumerate(k):
            coeffs[xx, x] = int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << PRECISION_BITS))
    return bounds, coeffs


def _pass(img: np.ndarray, bounds, coeffs, axis: int) -> np.ndarray:
    """One fixed-point convolution pass along ``axis`` of an HWC uint8 image."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.uint8)
    for j in range(bounds.shape[0]):
        x0, n = int(bounds[j, 0]), int(bounds[j, 1])
        acc = np.tensordot(coeffs[j, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[j] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``PIL.Image.fromarray(img).resize((out_w, out_h), resample=BICUBIC)`` for an [H, W, 3] uint8 array."""
    h, w = img.shape[:2]
    if w != out_w:
        img = _pass(img, *resample_plan(w, out_w), axis=1)
    if h != out_h:
        img = _pass(img, *resample_plan(h, out_h), axis=0)
    return img


def resize_output_size(h: int, w: int, size: int):
    """transformers get_resize_output_image_size(..., default_to_square=False): shortest edge -> size."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def norm_table(rescale=1 / 255, mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD) -> np.ndarray:
    v = (np.arange(256, dtype=np.uint8)[None, :] * rescale).astype(np.float32)          # rescale(): float64 product -> float32
    return (v - np.array(mean, np.float32)[:, None]) / np.array(std, np.float32)[:, None]  # normalize(): float32


def clip_preprocess(images_u8, size: int = 224, crop: int = 224, mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD,
                    rescale: float = 1 / 255, resize_fn=pil_resize_u8) -> np.ndarray:
    """CLIPImageProcessor.preprocess for a list / batch of [H, W, 3] uint8 images -> float32 [B, 3, crop, crop]."""
    out = []
    for img in images_u8:
        h, w = img.shape[:2]
        rh, rw = resize_output_size(h, w, size)
        r = resize_fn(np.ascontiguousarray(img), rh, rw)
        top, left = (rh - crop) // 2, (rw - crop) // 2
        if top < 0 or left < 0:
            raise ValueError("crop larger than the resized image is not restated (the reference never hits it)")
        r = r[top:top + crop, left:left + crop]
        x = (r * rescale).astype(np.float32)
        x = (x - np.array(mean, dtype=x.dtype)) / np.array(std, dtype=x.dtype)
        out.append(x.transpose(2, 0, 1))
    return np.stack(out)
