"""Reward-side image preprocessing on the device (psob200_clip_preprocess) -- host side.

The reference decodes the sampled latents, converts the images to uint8 ON THE HOST and hands PIL images to the reward
model's ``processor`` (train_online_pso_sdxl_turbo.py:632-640; pso_pytorch/pickscore_utils.py:24-33), which resizes / crops /
normalises with PIL + numpy and copies the result back to the GPU.  ``clip_image_preprocess`` produces the same
``pixel_values`` -- bit for bit -- from the decoded images where they already are, in one kernel launch:

    pixel_values = clip_image_preprocess(images)            # images: float [B,3,H,W] in [-1,1] (the VAE output) or uint8 [B,H,W,3]
    image_embs = model.get_image_features(pixel_values=pixel_values)      # pickscore_utils.py:44

``DeviceCLIPImageProcessor`` is the drop-in for the ``images=`` call of the processor object (returns ``{"pixel_values": ...}``).
The reward models themselves (CLIP-H / PickScore weights) are outside the PSO hot path and are not part of this package.
No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

_plans: dict = {}
_tables: dict = {}


def resize_output_size(h: int, w: int, size: int):
    """transformers ``get_resize_output_image_size(image, size, default_to_square=False)``: shortest edge -> ``size``."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def resample_plan(in_size: int, out_size: int):
    """Pillow's bicubic coefficient table of one axis (psob200_resample_plan, host): (bounds [out,2], coeffs [out,taps]) int32."""
    L = _lib.lib()
    taps = L.psob200_resample_taps(in_size, out_size)
    if taps <= 0:
        raise _lib.Psob200Error(f"invalid resample sizes {in_size} -> {out_size}")
    bounds = np.zeros((out_size, 2), np.int32)
    coeffs = np.zeros((out_size, taps), np.int32)
    _lib.check(L.psob200_resample_plan(in_size, out_size, bounds.ctypes.data_as(C.POINTER(C.c_int32)),
                                       coeffs.ctypes.data_as(C.POINTER(C.c_int32))), "psob200_resample_plan")
    return bounds, coeffs


def _device_plan(in_size: int, out_size: int, dev: torch.device):
    key = (in_size, out_size, dev)
    hit = _plans.get(key)
    if hit is None:
        b, c = resample_plan(in_size, out_size)
        hit = (torch.from_numpy(b).to(dev), torch.from_numpy(c).to(dev), c.shape[1])
        _plans[key] = hit
    return hit


def norm_table(rescale: float, mean, std) -> np.ndarray:
    """table[c, v] = float32((float32(v * rescale) - mean[c]) / std[c]) -- transformers ``rescale`` then ``normalize``."""
    m, s = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    out = np.zeros((3, 256), np.float32)
    _lib.check(_lib.lib().psob200_clip_norm_table(float(rescale), m, s, out.ctypes.data_as(C.POINTER(C.c_float))),
               "psob200_clip_norm_table")
    return out


def _device_table(rescale, mean, std, dev):
    key = (float(rescale), tuple(mean), tuple(std), dev)
    hit = _tables.get(key)
    if hit is None:
        hit = _tables[key] = torch.from_numpy(norm_table(rescale, mean, std)).to(dev)
    return hit


def clip_image_preprocess(images: torch.Tensor, size: int = 224, crop_size: int = 224, image_mean=OPENAI_CLIP_MEAN,
                          image_std=OPENAI_CLIP_STD, rescale_factor: float = 1 / 255, out_dtype=torch.float32) -> torch.Tensor:
    """``CLIPImageProcessor(images=[PIL...])["pixel_values"]`` for a batch that is already on the GPU.

    images: uint8 [B, H, W, 3] (what ``Image.fromarray`` would receive) or float [B, 3, H, W] in [-1, 1] (the decoded VAE
    output; quantised exactly like turbo :632).  Returns ``out_dtype`` [B, 3, crop_size, crop_size] (the processor returns
    float32)."""
    dev = _lib.require_cuda(images)
    a = _lib.ClipPreprocessArgs()
    if images.dtype == torch.uint8:
        if images.dim() != 4 or images.shape[-1] != 3:
            raise _lib.Psob200Error(f"uint8 images must be [B, H, W, 3], got {tuple(images.shape)}")
        B, H, W = images.shape[:3]
        a.src_dtype = _lib.U8
    else:
        if images.dim() != 4 or images.shape[1] != 3:
            raise _lib.Psob200Error(f"float images must be [B, 3, H, W], got {tuple(images.shape)}")
        B, _, H, W = images.shape
        a.src_dtype = _lib.dtype_code(images)
    src = images.contiguous()
    rh, rw = resize_output_size(H, W, size)
    top, left = (rh - crop_size) // 2, (rw - crop_size) // 2
    if top < 0 or left < 0:
        raise _lib.Psob200Error("crop_size larger than the resized image")
    bh, ch, th = _device_plan(W, rw, dev)
    bv, cv, tv = _device_plan(H, rh, dev)
    table = _device_table(rescale_factor, image_mean, image_std, dev)
    out = torch.empty(B, 3, crop_size, crop_size, dtype=out_dtype, device=dev)
    a.src, a.dst = src.data_ptr(), out.data_ptr()
    a.bounds_h, a.coeffs_h, a.bounds_v, a.coeffs_v = bh.data_ptr(), ch.data_ptr(), bv.data_ptr(), cv.data_ptr()
    a.norm_table = table.data_ptr()
    a.B, a.in_h, a.in_w, a.rs_h, a.rs_w = B, H, W, rh, rw
    a.out_h = a.out_w = crop_size
    a.crop_top, a.crop_left, a.taps_h, a.taps_v = top, left, th, tv
    a.dst_dtype = _lib.dtype_code(out)
    _lib.launch(dev, "psob200_clip_preprocess", C.byref(a), _lib.current_stream(dev))
    return out


class DeviceCLIPImageProcessor:
    """Stands in for ``self.processor(images=..., return_tensors="pt")`` of the reward selectors (pickscore_utils.py:26-32):
    takes the decoded image batch (device tensor) instead of a list of PIL images."""

    def __init__(self, size: int = 224, crop_size: int = 224, image_mean=OPENAI_CLIP_MEAN, image_std=OPENAI_CLIP_STD,
                 rescale_factor: float = 1 / 255, out_dtype=torch.float32):
        self.kw = dict(size=size, crop_size=crop_size, image_mean=image_mean, image_std=image_std,
                       rescale_factor=rescale_factor, out_dtype=out_dtype)

    def __call__(self, images: torch.Tensor, **_ignored):
        return {"pixel_values": clip_image_preprocess(images, **self.kw)}
