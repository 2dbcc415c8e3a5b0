"""Property tests (hypothesis) over random batch sizes, latent shapes, storage types and schedules: the fused loss+grad
kernel against the fp64 closed form of the oracle, and the tcgen05 GEMM against a torch fp32 matmul.  Shapes the
reference's configs never exercise (odd N, N < 8, B = 1, ragged K / N tails) must still match: every code path of the
launch heuristics (tensor-memory kernel with cluster 1..8, general vector / scalar kernel; GEMM tile widths and tails)."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from tests import _util as U

pytestmark = pytest.mark.gpu
_SETTINGS = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow],
                 derandomize=True)


@pytest.fixture(scope="module")
def pso(built_lib):
    import pairwise_sample_optimization_b200 as p
    return p


@settings(max_examples=30, **_SETTINGS)
@given(B=st.integers(1, 24), c=st.sampled_from([1, 3, 4]), h=st.integers(1, 48), w=st.integers(1, 48),
       kind=st.sampled_from(["turbo", "dmd"]), pd=st.sampled_from(["fp32", "bf16", "fp16"]),
       ld=st.sampled_from(["fp32", "bf16", "fp16"]), seed=st.integers(0, 10_000), ties=st.sampled_from([0, 2, 3]))
def test_pair_loss_any_shape_matches_closed_form(pso, B, c, h, w, kind, pd, ld, seed, ties):
    d = U.synth(kind, B, (c, h, w), seed, 0.02, ties, U.DT[pd], U.DT[ld])
    cf = U.oracle_fp64(d)
    loss, stt, g0, g1 = U.run_fused(pso, d)
    pso.check_status()
    assert abs(loss.item() - cf["loss"].item()) <= 1e-5 * abs(cf["loss"].item())
    np.testing.assert_allclose(stt[:, 4:6].T.cpu().numpy(), torch.stack(cf["delta"]).numpy(), rtol=5e-5, atol=2e-8)
    for gk, wk in ((g0, cf["grads"][0]), (g1, cf["grads"][1])):
        if float(wk.abs().max()) == 0.0:  # every pair tied or gated: exact zeros
            assert float(gk.abs().max()) == 0.0
        elif pd == "fp32":
            assert U.rel_max(gk, wk) <= 1e-5
        else:
            U.assert_rounded_equal(gk, wk, U.DT[pd])


@settings(max_examples=30, **_SETTINGS)
@given(M=st.integers(1, 700), N=st.integers(1, 400), K=st.integers(1, 300), r=st.sampled_from([0, 4, 8, 20, 64]),
       bias=st.booleans(), dt=st.sampled_from([torch.bfloat16, torch.float16]), seed=st.integers(0, 1000))
def test_gemm_any_shape_matches_fp32_matmul(pso, M, N, K, r, bias, dt, seed):
    from pairwise_sample_optimization_b200 import gemm
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s, sc=1.0: (torch.randn(*s, device="cuda", generator=g) * sc).to(dt)
    a1, b1 = rn(M, K), rn(N, K, sc=K ** -0.5)
    a2, b2 = (rn(M, r), rn(N, r, sc=0.1)) if r else (None, None)
    bv = rn(N) if bias else None
    want = a1.float() @ b1.float().t()
    if r:
        want = want + a2.float() @ b2.float().t()
    want = 0.75 * want + (bv.float() if bias else 0.0)
    got, got_t = gemm.lora_gemm(a1, b1, a2, b2, bias=bv, alpha=0.75, out_dtype=torch.float32, want_out_t=True)
    den = max(want.abs().max().item(), 1e-6)
    assert (got - want).abs().max().item() <= 3e-5 * den
    assert torch.equal(got_t, got.t())
