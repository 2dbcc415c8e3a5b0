"""Load the reference's two step-with-logprob files verbatim (TEST INFRASTRUCTURE).

The files are executed from the read-only mount, never copied:

  /root/reference/human_preference_tuning/pso_pytorch/diffusers_patch/
      turbo_inference_with_logprob.py      (turbo_step_with_logprob, :24-116)
      distilled_inference_with_logprob.py  (distilled_step_with_logprob, :45-137)

Their only non-torch imports are ``randn_tensor`` and three scheduler class
names used purely as annotations (turbo :13-15, distilled :13-15); diffusers is
not installed in this image, so those names are provided by stub modules.

``available()`` is False on the GPU box (no /root/reference there): callers in
``tests/`` skip, and the committed fixtures under ``tests/golden`` stand in.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("PSO_REFERENCE_ROOT", "/root/reference")
_PATCH_DIR = os.path.join(
    REFERENCE_ROOT, "human_preference_tuning", "pso_pytorch", "diffusers_patch"
)
_cache: dict[str, types.ModuleType] = {}


def available() -> bool:
    return os.path.isfile(os.path.join(_PATCH_DIR, "turbo_inference_with_logprob.py"))


def _randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    # diffusers.utils.torch_utils.randn_tensor for a single generator on one device.
    return torch.randn(tuple(shape), generator=generator, device=device, dtype=dtype)


def _install_stub() -> None:
    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "_pso_oracle_stub", False):
        return  # a real diffusers is importable; use it
    names = {
        "diffusers": {"DDPMScheduler": type("DDPMScheduler", (), {})},
        "diffusers.utils": {},
        "diffusers.utils.torch_utils": {"randn_tensor": _randn_tensor},
        "diffusers.schedulers": {},
        "diffusers.schedulers.scheduling_euler_ancestral_discrete": {
            "EulerAncestralDiscreteScheduler": type("EulerAncestralDiscreteScheduler", (), {})
        },
        "diffusers.schedulers.scheduling_ddim": {
            "DDIMSchedulerOutput": type("DDIMSchedulerOutput", (), {}),
            "DDIMScheduler": type("DDIMScheduler", (), {}),
        },
    }
    for name, attrs in names.items():
        mod = types.ModuleType(name)
        mod._pso_oracle_stub = True
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
    for name in names:
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])


def _load(stem: str) -> types.ModuleType:
    if stem in _cache:
        return _cache[stem]
    if not available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stub()
    path = os.path.join(_PATCH_DIR, stem + ".py")
    spec = importlib.util.spec_from_file_location("_pso_reference_" + stem, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[stem] = mod
    return mod


def turbo_step_with_logprob():
    """The reference's own function object (turbo_inference_with_logprob.py:24)."""
    return _load("turbo_inference_with_logprob").turbo_step_with_logprob


def distilled_step_with_logprob():
    """The reference's own function object (distilled_inference_with_logprob.py:45)."""
    return _load("distilled_inference_with_logprob").distilled_step_with_logprob


def get_x0_from_noise():
    """distilled_inference_with_logprob.py:36-42."""
    return _load("distilled_inference_with_logprob")._get_x0_from_noise
