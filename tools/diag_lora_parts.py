#!/usr/bin/env python
"""Device time of each of the six launches behind one LoRA-wrapped projection (forward: t, y; backward: u, dx, dA, dB),
timed separately (CUDA events around CUDA-graph replays) -- where the per-layer time goes.

    python tools/diag_lora_parts.py [--shape 8192,1280,1280] [--rank 64] [--split 0,8,13,15]
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pairwise_sample_optimization_b200 import _lib  # noqa: E402
from tools.bench_lora_gemm import timed  # noqa: E402

BF16, F32 = 1, 0


def gemm(dev, **kw):
    g = _lib.GemmArgs()
    g.alpha, g.ab_dtype, g.d_dtype = 1.0, _lib._DTYPES[torch.bfloat16], _lib._DTYPES[torch.bfloat16]
    for k, v in kw.items():
        setattr(g, k, v.data_ptr() if torch.is_tensor(v) else v)

    def call():
        _lib.check(_lib.lib().psob200_lora_gemm(C.byref(g), _lib.current_stream(dev)), "psob200_lora_gemm")
    return call


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="8192,1280,1280;32768,640,640;616,2048,1280")
    ap.add_argument("--rank", type=int, default=64)
    ap.add_argument("--split", default="0")
    ap.add_argument("--bn", default="0")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=gen) * sc).bfloat16()
    r = args.rank
    f32 = _lib._DTYPES[torch.float32]
    for shp in args.shape.split(";"):
        M, K, N = (int(v) for v in shp.split(","))
        M8 = (M + 7) // 8 * 8
        x, w, dy = rn(M, K), rn(N, K, sc=K ** -0.5), rn(M, N)
        A, Bm = rn(r, K, sc=1 / r), rn(N, r, sc=0.05)
        t, tt, u, ut = rn(M, r), rn(r, M8), rn(M, r), rn(r, M8)
        y, dx = torch.empty(M, N, device=dev, dtype=torch.bfloat16), torch.empty(M, K, device=dev, dtype=torch.bfloat16)
        dA, dB = torch.zeros(r, K, device=dev), torch.zeros(N, r, device=dev)
        parts = {
            "t  = x A^T (+tt)": gemm(dev, a1=x, lda1=K, b1=A, ldb1=K, M=M, N=r, K1=K, d=t, ldd=r, dt=tt, lddt=M8),
            "y  = x W^T + t B^T": gemm(dev, a1=x, lda1=K, b1=w, ldb1=K, a2=t, lda2=r, b2=Bm, ldb2=r, M=M, N=N, K1=K, K2=r, d=y, ldd=N),
            "y  plain (ref fwd)": gemm(dev, a1=x, lda1=K, b1=w, ldb1=K, M=M, N=N, K1=K, d=y, ldd=N),
            "u  = dy B (+ut)": gemm(dev, a1=dy, lda1=N, b1=Bm, ldb1=r, b_reduction_major=1, M=M, N=r, K1=N, d=u, ldd=r, dt=ut, lddt=M8),
            "dx = dy W + u A": gemm(dev, a1=dy, lda1=N, b1=w, ldb1=K, a2=u, lda2=r, b2=A, ldb2=K, b_reduction_major=1, M=M, N=K,
                                    K1=N, K2=r, d=dx, ldd=K),
        }
        for bn in (int(v) for v in args.bn.split(",")):
            if bn == 0:
                continue
            parts[f"y  = x W^T + t B^T  bn={bn}"] = gemm(dev, a1=x, lda1=K, b1=w, ldb1=K, a2=t, lda2=r, b2=Bm, ldb2=r, M=M, N=N, K1=K,
                                                        K2=r, d=y, ldd=N, tune_bn=bn)
            if bn % 128 == 0:
                parts[f"dx = dy W + u A  bn={bn}"] = gemm(dev, a1=dy, lda1=N, b1=w, ldb1=K, a2=u, lda2=r, b2=A, ldb2=K,
                                                         b_reduction_major=1, M=M, N=K, K1=N, K2=r, d=dx, ldd=K, tune_bn=bn)
        for sp in (int(v) for v in args.split.split(",")):
            parts[f"dA += u^T x  split={sp}"] = gemm(dev, a1=x, lda1=K, a_reduction_major=1, b1=ut, ldb1=M8, M=K, N=r, K1=M, dt=dA,
                                                     lddt=K, d_dtype=f32, accumulate=1, split_k=sp)
            parts[f"dB += dy^T t split={sp}"] = gemm(dev, a1=dy, lda1=N, a_reduction_major=1, b1=tt, ldb1=M8, M=N, N=r, K1=M, d=dB,
                                                     ldd=r, d_dtype=f32, accumulate=1, split_k=sp)
        print(f"--- M={M} K={K} N={N} r={r}")
        for name, fn in parts.items():
            print(f"{timed(fn):8.1f} us  {name}", flush=True)
        seq_f = lambda: (parts["t  = x A^T (+tt)"](), parts["y  = x W^T + t B^T"]())
        print(f"{timed(seq_f):8.1f} us  forward, two launches back to back (no PDL)")


if __name__ == "__main__":
    main()
