"""psob200_lora_gemm (tcgen05 / TMEM / TMA) against a plain torch fp32 matmul of the same 16-bit operands.

Tolerances: fp32 outputs 2e-5 of max|D| (fp32 accumulation in a different order); 16-bit outputs must equal the
fp32 reference rounded to the output type up to 1 ulp (north_star: 1e-3 relative for bf16 is looser than that).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gemm(built_lib):
    from pairwise_sample_optimization_b200 import gemm as g
    return g


def _rand(rows, cols, dtype, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(rows, cols, device="cuda", generator=g) * scale).to(dtype)


def _ref(a1, b1, a2=None, b2=None, bias=None, alpha=1.0):
    d = a1.float() @ b1.float().t()
    if a2 is not None:
        d = d + a2.float() @ b2.float().t()
    d = alpha * d
    if bias is not None:
        d = d + bias.float()
    return d


def _check(got, want, dtype):
    assert got.dtype == dtype
    den = want.abs().max().item()
    if dtype == torch.float32:
        assert (got - want).abs().max().item() <= 2e-5 * den
    else:
        ulp = want.abs() * (2.0 ** (-7 if dtype == torch.bfloat16 else -10)) + 1e-6 * den
        assert bool(((got.float() - want).abs() <= ulp).all())


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 128), (200, 72, 100), (1024, 640, 640), (77, 1280, 2048),
                                   (384, 16, 1280), (130, 8, 72)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_single_segment(gemm, M, N, K, dtype):
    a, b = _rand(M, K, dtype, 1), _rand(N, K, dtype, 2, K ** -0.5)
    want = _ref(a, b)
    got32, _ = gemm.lora_gemm(a, b, out_dtype=torch.float32)
    _check(got32, want, torch.float32)
    got16, got16_t = gemm.lora_gemm(a, b, want_out_t=True)
    _check(got16, want, dtype)
    assert torch.equal(got16_t, got16.t())


@pytest.mark.parametrize("M,K,N,r", [(1024, 640, 640, 8), (512, 1280, 1280, 64), (300, 640, 640, 4), (256, 2048, 1280, 128),
                                     (2048, 1280, 1280, 16)])
def test_fused_base_plus_lora_segment(gemm, M, K, N, r):
    """y = x W^T + b + T B^T in one pass (T = s x A^T computed by a first skinny pass of the same kernel)."""
    dt = torch.bfloat16
    x, w = _rand(M, K, dt, 3), _rand(N, K, dt, 4, K ** -0.5)
    A, Bm = _rand(r, K, dt, 5, 1.0 / r), _rand(N, r, dt, 6, 0.05)
    bias = _rand(1, N, dt, 7)[0]
    s = 0.5
    T, Tt = gemm.lora_gemm(x, A, alpha=s, want_out_t=True)
    _check(T, _ref(x, A, alpha=s), dt)
    assert torch.equal(Tt, T.t())
    y, _ = gemm.lora_gemm(x, w, T, Bm, bias=bias)
    _check(y, _ref(x, w, T, Bm, bias=bias), dt)


@pytest.mark.parametrize("Mred,Mo,No", [(1024, 640, 16), (4096, 1280, 64), (1000, 200, 8), (8192, 640, 128)])
@pytest.mark.parametrize("split_k", [0, 1, 3])
def test_reduction_major_accumulate(gemm, Mred, Mo, No, split_k):
    """dA^T[K, r] += X^T U with X given token-major ([Mred, Mo]) and U^T [r, Mred]: the weight-gradient shape."""
    dt = torch.bfloat16
    x, ut = _rand(Mred, Mo, dt, 8), _rand(No, Mred, dt, 9, Mred ** -0.5)
    want = x.float().t() @ ut.float().t()
    acc = torch.ones(Mo, No, device="cuda")
    acc_t = torch.full((No, Mo), 2.0, device="cuda")
    gemm.lora_gemm(x, ut, out=acc, out_t=acc_t, a_reduction_major=True, accumulate=True, split_k=split_k)
    den = want.abs().max().item()
    assert (acc - 1.0 - want).abs().max().item() <= 3e-5 * den + 1e-6
    assert (acc_t - 2.0 - want.t()).abs().max().item() <= 3e-5 * den + 1e-6


def test_many_tiles_persistent_schedule(gemm):
    """More tiles than SMs, two accumulator buffers and the 4-stage ring wrapping many times."""
    dt = torch.bfloat16
    a, b = _rand(128 * 40, 320, dt, 10), _rand(1280, 320, dt, 11, 320 ** -0.5)
    got, _ = gemm.lora_gemm(a, b, out_dtype=torch.float32, tune_bn=128)
    _check(got, _ref(a, b), torch.float32)


def test_argument_errors(gemm):
    from pairwise_sample_optimization_b200 import _lib
    a, b = _rand(64, 64, torch.bfloat16, 1), _rand(64, 64, torch.bfloat16, 2)
    with pytest.raises(_lib.Psob200Error):
        gemm.lora_gemm(a.float(), b.float())
    with pytest.raises(_lib.Psob200Error):
        gemm.lora_gemm(a, b[:, :32])
    with pytest.raises(_lib.Psob200Error):
        gemm.lora_gemm(a, b, accumulate=True)


@pytest.mark.parametrize("M,K,N,r", [(256, 64, 256, 0), (512, 320, 256, 0), (1000, 200, 328, 8), (2048, 1280, 1280, 64),
                                     (300, 640, 640, 16), (4096, 640, 160, 0), (37 * 256, 128, 512, 8)])
@pytest.mark.parametrize("b_mn", [False, True])
def test_cta_pair_kernel(gemm, M, K, N, r, b_mn):
    """The cta_group::2 variant (forced with diag bit 16): 256-row pair tiles, halves of B per CTA, multicast commits."""
    dt = torch.bfloat16
    a1, b1 = _rand(M, K, dt, 21), _rand(N, K, dt, 22, K ** -0.5)
    a2, b2 = (_rand(M, r, dt, 23), _rand(N, r, dt, 24, 0.1)) if r else (None, None)
    bias = _rand(1, N, dt, 25)[0]
    want = _ref(a1, b1, a2, b2, bias=bias, alpha=0.5)
    if b_mn:  # B given reduction-major ([K, N] row-major), as the backward consumes W / lora_A / lora_B
        if N % 8:
            pytest.skip("reduction-major operands need a 16-byte row pitch")
        from pairwise_sample_optimization_b200 import _lib
        import ctypes as C
        b1t = b1.t().contiguous()
        b2t = b2.t().contiguous() if r else None
        out = torch.empty(M, N, device="cuda", dtype=torch.float32)
        g = _lib.GemmArgs()
        g.a1, g.lda1, g.b1, g.ldb1 = a1.data_ptr(), a1.stride(0), b1t.data_ptr(), b1t.stride(0)
        if r:
            a2p = torch.zeros(M, (r + 7) // 8 * 8, device="cuda", dtype=dt)
            a2p[:, :r] = a2
            g.a2, g.lda2, g.b2, g.ldb2, g.K2 = a2p.data_ptr(), a2p.stride(0), b2t.data_ptr(), b2t.stride(0), r
        g.bias, g.bias_dtype = bias.data_ptr(), _lib.BF16
        g.d, g.ldd, g.M, g.N, g.K1 = out.data_ptr(), N, M, N, K
        g.alpha, g.ab_dtype, g.d_dtype, g.b_reduction_major, g.diag = 0.5, _lib.BF16, _lib.F32, 1, 0x10000
        _lib.check(_lib.lib().psob200_lora_gemm(C.byref(g), _lib.current_stream(a1.device)), "psob200_lora_gemm")
        got = out
    else:
        got, _ = gemm.lora_gemm(a1, b1, a2, b2, bias=bias, alpha=0.5, out_dtype=torch.float32, diag=0x10000)
    _check(got, want, torch.float32)
    if not b_mn:
        got16, _ = gemm.lora_gemm(a1, b1, a2, b2, bias=bias, alpha=0.5, diag=0x10000)
        _check(got16, want, dt)


@pytest.mark.parametrize("bn", [16, 48, 80, 144, 208])
def test_tile_widths_that_are_not_multiples_of_the_epilogue_chunk(gemm, bn):
    """Several column tiles whose width is not a multiple of the 32-column epilogue chunk: a chunk must not spill into
    the neighbouring tile."""
    a, b = _rand(300, 200, torch.bfloat16, 31), _rand(500, 200, torch.bfloat16, 32, 200 ** -0.5)
    got, _ = gemm.lora_gemm(a, b, out_dtype=torch.float32, tune_bn=bn)
    _check(got, _ref(a, b), torch.float32)
