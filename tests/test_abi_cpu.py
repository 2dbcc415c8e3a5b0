"""CPU: the C-ABI library builds, loads, and exports every symbol include/psob200.h declares;
argument validation that needs no GPU; the product path refuses to run without CUDA."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "psob200.h")).read()
    return sorted(set(re.findall(r"PSOB200_API\s+[\w\s\*]+?\b(psob200_\w+)\s*\(", src)))


def test_header_declares_what_binding_binds():
    from pairwise_sample_optimization_b200 import _lib
    assert _declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(built_lib):
    handle = C.CDLL(built_lib)
    for name in _declared_symbols():
        assert hasattr(handle, name), name


def test_binding_loads_and_struct_sizes_match(built_lib):
    from pairwise_sample_optimization_b200 import _lib
    lib = _lib.lib()  # raises on any ABI mismatch (psob200_struct_size check)
    assert lib.psob200_abi_version() == 2
    assert lib.psob200_strerror(0) == b"ok"
    assert b"workspace" in lib.psob200_strerror(-6)
    assert lib.psob200_pair_loss_workspace_bytes(0) == 16
    assert lib.psob200_pair_loss_workspace_bytes(5) == 16 + 32


def test_argument_validation_without_gpu(built_lib):
    from pairwise_sample_optimization_b200 import _lib
    lib = _lib.lib()
    assert lib.psob200_online_pso_loss_grad(None, None, None) == -1
    assert lib.psob200_dreambooth_pso_loss_grad(None, None) == -1
    assert lib.psob200_step_logprob(None, None, None) == -1
    assert lib.psob200_step_logprob_backward(None, None, None) == -1
    assert lib.psob200_scale(None, None, 8, 1.0, 0, 0, None) == -1
    sched = _lib.Schedule()
    sched.kind = _lib.SCHED_AFFINE
    args = _lib.OnlinePsoArgs()
    args.B, args.N = 0, 16
    assert lib.psob200_online_pso_loss_grad(C.byref(sched), C.byref(args), None) == -1
    # entry points added later in the round: rejected before any CUDA call
    assert lib.psob200_geglu_forward(None, None) == -1 and lib.psob200_geglu_backward(None, None) == -1
    g = _lib.GegluArgs()
    g.proj, g.out, g.M, g.I, g.ld_proj, g.ld_out, g.dtype = 16, 32, 4, 10, 20, 10, _lib._DTYPES[torch.bfloat16]
    assert lib.psob200_geglu_forward(C.byref(g), None) == -5  # half width 10 is not a multiple of 16 bytes
    g.I, g.ld_proj, g.ld_out, g.proj = 16, 32, 16, 8
    assert lib.psob200_geglu_forward(C.byref(g), None) == -3  # misaligned pointer
    assert lib.psob200_flat_allreduce_sumsq(None, None) == -1
    x = _lib.FlatAllreduceArgs()
    x.grad_multicast, x.sumsq_multicast, x.n, x.rank, x.world, x.scale = 16, 64, 1024, 2, 2, 0.5
    assert lib.psob200_flat_allreduce_sumsq(C.byref(x), None) == -1  # rank outside the group
    x.rank, x.n = 1, 1022
    assert lib.psob200_flat_allreduce_sumsq(C.byref(x), None) == -5  # 128-bit multimem accesses need n % 4 == 0
    o = _lib.FlatAdamwArgs()
    o.param = o.grad = o.exp_avg = o.exp_avg_sq = o.workspace = 16
    o.n, o.step, o.n_sumsq_parts = 8, 1, 2  # pieces announced but no pointer to them
    assert lib.psob200_flat_adamw_step(C.byref(o), None) == -1


def test_product_path_has_no_cpu_fallback(built_lib):
    import pairwise_sample_optimization_b200 as pso
    from oracle import schedules
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(pso._lib.Psob200Error, match="CUDA"):
        pso.turbo_step_with_logprob(schedules.turbo_scheduler(4), x, torch.tensor([999.0]), x, prev_sample=x)
    with pytest.raises(pso._lib.Psob200Error, match="CUDA"):
        pso.pso_db_loss(x.repeat(2, 1, 1, 1), None, x.repeat(2, 1, 1, 1), x.repeat(2, 1, 1, 1), torch.ones(2),
                        loss_type="pso_db")
    with pytest.raises(ValueError):
        pso.pso_db_loss(x, None, x, x, torch.ones(1), loss_type="sigmoid")  # P:1929


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pairwise_sample_optimization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)
                assert "/root/reference" not in src, os.path.join(dirpath, f)


def test_compare_host_logic_matches_oracle():
    import pairwise_sample_optimization_b200 as pso
    from oracle import losses
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(64, 3, generator=g).round(decimals=1), torch.randn(64, 3, generator=g).round(decimals=1)
    assert torch.equal(pso.compare(a, b), losses.compare(a, b))
    assert torch.equal(pso.compare(a[:, 0], b[:, 0]), losses.compare(a[:, 0], b[:, 0]))
    got = pso.sample_compare(a, b, generator=torch.Generator().manual_seed(5))
    idx = torch.randint(0, 3, (64,), generator=torch.Generator().manual_seed(5))
    assert torch.equal(got, losses.sample_compare(a, b, reward_indices=idx))
    # a NaN reward satisfies neither `a <= b` nor `b < a`: the reference leaves that row [0, 0] (turbo :401-416)
    a[3], b[7] = float("nan"), float("nan")
    got = pso.sample_compare(a, b, generator=torch.Generator().manual_seed(5))
    assert torch.equal(got, losses.sample_compare(a, b, reward_indices=idx))
    assert got[3].tolist() == [0.0, 0.0] and got[7].tolist() == [0.0, 0.0]
