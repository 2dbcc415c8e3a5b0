"""Philox4x32-10 + Box-Muller restated in numpy (TEST INFRASTRUCTURE): the oracle of the sampler kernel's throughput mode
(psob200_step_args.use_philox; SURVEY.md section 7, hard part 8).

The reference draws its sampler noise with torch's generator (``randn_tensor``: turbo_inference_with_logprob.py:97,
distilled_inference_with_logprob.py:123-124); a stream of torch's generator cannot be reproduced inside a custom kernel, so
parity mode takes the noise as an input and this mode is specified here instead: the published Philox4x32-10 round function
(Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) with counter (q_lo, q_hi, offset_lo,
offset_hi), key (seed_lo, seed_hi), uniforms ``((x >> 8) + 0.5) * 2^-24``, Box-Muller pairs ``(r cos 2 pi u2, r sin 2 pi u2)``,
``r = sqrt(-2 ln u1)``; element ``4 q + j`` of the noise tensor gets draw ``j`` of counter ``q``."""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(q: np.ndarray, seed: int, offset: int) -> np.ndarray:
    """q: uint64 counters -> uint32 [len(q), 4]."""
    q = q.astype(np.uint64)
    c0, c1 = (q & MASK), (q >> np.uint64(32))
    c2 = np.full_like(c0, offset & MASK)
    c3 = np.full_like(c0, (offset >> 32) & MASK)
    k0, k1 = seed & MASK, (seed >> 32) & MASK
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def normal(n: int, seed: int, offset: int) -> np.ndarray:
    """The first n draws of the stream (seed, offset) as float64 (the kernel evaluates the same formulas in fp32)."""
    groups = (n + 3) // 4
    x = philox4x32_10(np.arange(groups, dtype=np.uint64), seed, offset).astype(np.float64)
    u = (np.floor(x / 256.0) + 0.5) * 2.0 ** -24
    out = np.empty((groups, 4))
    for h in (0, 1):
        rad = np.sqrt(-2.0 * np.log(u[:, 2 * h]))
        ang = 2.0 * np.pi * u[:, 2 * h + 1]
        out[:, 2 * h] = rad * np.cos(ang)
        out[:, 2 * h + 1] = rad * np.sin(ang)
    return out.reshape(-1)[:n]
