"""Per-device runtime state for the C-ABI calls: schedule tables resident in HBM, the
zero-initialised pair-loss workspace and the device status word.

Nothing here synchronises the host with the device; ``check_status`` is the one explicit,
opt-in sync (it is how a timestep that is not in the schedule surfaces as an exception,
turbo_inference_with_logprob.py:63 raises IndexError there).
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import _lib

_tables: dict = {}
_SAMPLER_NOISE = {"mode": "philox"}


def set_sampler_noise(mode: str) -> None:
    """Where the sampler's N(0,1) draws come from when the caller passes NO generator: "philox" (default) = generated inside the
    step kernel (no noise tensor, no extra launch; keyed by torch's seed and a call counter), "torch" = ``torch.randn`` from the
    global generator, as the reference does.  With an explicit ``generator`` the draw is always ``torch.randn(generator=...)``."""
    if mode not in ("philox", "torch"):
        raise ValueError(mode)
    _SAMPLER_NOISE["mode"] = mode


def sampler_noise_in_kernel(generator) -> bool:
    return generator is None and _SAMPLER_NOISE["mode"] == "philox"

_workspaces: dict = {}
_status: dict = {}


def device_table(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    """fp32 copy of a small scheduler table on ``device``, cached per source tensor version."""
    if t.device == device and t.dtype == torch.float32 and t.is_contiguous():
        return t
    key = (id(t), t._version, device)
    hit = _tables.get(key)
    if hit is not None and hit[0]() is t:
        return hit[1]
    if len(_tables) > 64:
        _tables.clear()
    d = t.detach().to(device=device, dtype=torch.float32).contiguous()
    _tables[key] = (weakref.ref(t), d)
    return d


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Zero-initialised scratch for the pair-loss kernels, one per (device, stream); the kernels
    leave it zeroed, so it is memset only when (re)allocated."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 4096), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def status_word(device: torch.device) -> torch.Tensor:
    st = _status.get(device)
    if st is None:
        st = torch.zeros(1, dtype=torch.int32, device=device)
        _status[device] = st
    return st


def check_status(device=None) -> None:
    """Synchronising check of the device status word; raises like the reference would have."""
    devices = [torch.device(device)] if device is not None else list(_status)
    for dev in devices:
        st = _status.get(dev)
        if st is None:
            continue
        v = int(st.item())
        if v:
            st.zero_()
            if v & _lib.STATUS_TIMESTEP_NOT_IN_SCHEDULE:
                raise IndexError("a timestep passed to the PSO kernels is not in the scheduler's schedule "
                                 "(turbo_inference_with_logprob.py:63 / alphas_cumprod index out of range)")
            raise FloatingPointError("non-finite step coefficient (k, a) in the PSO kernels")


def timesteps_on(ts, device: torch.device) -> torch.Tensor:
    """Per-sample timesteps as a contiguous int64 / float32 / int32 device tensor (no sync for device inputs)."""
    if not torch.is_tensor(ts):
        ts = torch.as_tensor(ts)
    if ts.dim() == 0:
        ts = ts.reshape(1)
    if ts.dtype not in (torch.int64, torch.float32, torch.int32):
        ts = ts.to(torch.float32 if ts.is_floating_point() else torch.int64)
    if ts.device != device:
        ts = ts.to(device, non_blocking=True)
    return ts.contiguous().reshape(-1)


class ScheduleDesc:
    """A ``psob200_schedule`` plus the device tensors that back its pointers."""

    def __init__(self, kind: int, device: torch.device, ts_dtype: int, sched_timesteps=None, table=None):
        self.keep = (sched_timesteps, table)
        self.c = _lib.Schedule()
        self.c.kind = kind
        self.c.ts_dtype = ts_dtype
        if kind == _lib.SCHED_TURBO:
            self.c.n_table = int(sched_timesteps.numel())
            self.c.sched_timesteps = sched_timesteps.data_ptr()
            self.c.table = table.data_ptr()
        elif kind == _lib.SCHED_DMD:
            self.c.n_table = int(table.numel())
            self.c.table = table.data_ptr()

    def ref(self):
        return C.byref(self.c)


def turbo_schedule(scheduler, device: torch.device, ts_dtype: int) -> ScheduleDesc:
    """Reads ``scheduler.timesteps`` / ``scheduler.sigmas`` (turbo_inference_with_logprob.py:63,66,77-78)."""
    tst = device_table(scheduler.timesteps, device)
    sig = device_table(scheduler.sigmas, device)
    if sig.numel() < tst.numel() + 1:
        raise _lib.Psob200Error("scheduler.sigmas must have one more entry than scheduler.timesteps")
    return ScheduleDesc(_lib.SCHED_TURBO, device, ts_dtype, tst, sig)


def dmd_schedule(scheduler, device: torch.device, ts_dtype: int) -> ScheduleDesc:
    """Reads ``scheduler.alphas_cumprod`` (distilled_inference_with_logprob.py:85,98-99)."""
    return ScheduleDesc(_lib.SCHED_DMD, device, ts_dtype, None, device_table(scheduler.alphas_cumprod, device))


def affine_schedule(device: torch.device) -> ScheduleDesc:
    return ScheduleDesc(_lib.SCHED_AFFINE, device, _lib.TS_I64)
