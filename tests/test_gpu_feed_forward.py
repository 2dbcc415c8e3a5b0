"""GPU parity: fused GEGLU kernels (psob200_geglu_forward / _backward) vs the two lines of diffusers==0.27.0
``GEGLU.forward`` evaluated by torch in fp64 (`hidden, gate = proj.chunk(2, -1); hidden * F.gelu(gate)`) and its autograd."""
import pytest
import torch
import torch.nn.functional as F

from tests import _util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ff(built_lib):
    from pairwise_sample_optimization_b200 import feed_forward
    return feed_forward


def _truth(proj, dout):
    p = proj.detach().double().cpu().requires_grad_(True)
    h, g = p.chunk(2, dim=-1)
    out = h * F.gelu(g)
    out.backward(dout.double().cpu())
    return out.detach(), p.grad


@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
@pytest.mark.parametrize("shape", [(3, 7, 16), (2, 77, 2 * 2560), (1, 5, 2 * 5120), (4, 33, 24)])
def test_geglu_forward_backward(ff, dtype, shape):
    dt = U.DT[dtype]
    if dtype != "fp32" and (shape[-1] // 2) % 8:
        pytest.skip("16-bit rows need a multiple of 8 features per half")
    g = torch.Generator().manual_seed(hash(shape) % 1000)
    proj = (torch.randn(*shape, generator=g) * 2.0).to(dt).cuda().requires_grad_(True)
    dout = torch.randn(*shape[:-1], shape[-1] // 2, generator=g).to(dt).cuda()
    out = ff.geglu(proj)
    out.backward(dout)
    want_out, want_grad = _truth(proj, dout)
    if dtype == "fp32":
        assert U.rel_max(out, want_out) <= 1e-5  # tolerance of BASELINE.json's north_star for fp32
        assert U.rel_max(proj.grad, want_grad) <= 1e-5
    else:
        # fp32 erf / exp against the fp64 truth: a few per cent of the values land on the other side of a rounding
        # boundary; none may be off by more than 1 ulp of the storage type
        U.assert_rounded_equal(out.detach(), want_out, dt, max_mismatch_frac=0.05)
        U.assert_rounded_equal(proj.grad, want_grad, dt, max_mismatch_frac=0.05)


def test_geglu_strided_rows_and_errors(ff):
    from pairwise_sample_optimization_b200 import _lib
    big = torch.randn(6, 64, device="cuda").bfloat16()
    view = big[:, :32]  # row pitch 64, 2 I = 32
    out = ff.geglu(view)
    h, g = view.float().chunk(2, dim=-1)
    U.assert_rounded_equal(out, (h.double() * F.gelu(g.double())).cpu(), torch.bfloat16, max_mismatch_frac=0.05)
    with pytest.raises(_lib.Psob200Error):
        ff.geglu(torch.randn(4, 32))  # CPU tensor: no fallback
    with pytest.raises(_lib.Psob200Error):
        ff.geglu(torch.randn(4, 20, device="cuda").bfloat16())  # half width 10: not a multiple of 16 bytes


def test_install_on_fixture_matches_stock(ff):
    """The patched GEGLU modules of the tiny SDXL-architecture fixture give the stock forward / backward results."""
    from fixtures import sdxl_unet
    torch.manual_seed(0)
    mod = sdxl_unet.GEGLU(64, 256).cuda().bfloat16()
    x = torch.randn(2, 10, 64, device="cuda").bfloat16()
    xs = x.clone().requires_grad_(True)
    want = mod(xs)
    want.sum().backward()
    assert ff.install_fused_geglu(mod) == 1
    xf = x.clone().requires_grad_(True)
    got = mod(xf)
    got.sum().backward()
    assert U.rel_max(got, want) <= 2 ** -7
    assert U.rel_max(xf.grad, xs.grad) <= 2 ** -6
