"""LoRA-wrapped projections on the tcgen05 path (rows a9-a11 of SURVEY.md section 8) against the oracle's restated
peft / diffusers data flow (oracle/lora.py), fp64 with the 16-bit rounding points of a bf16 run.

Tolerances: 16-bit outputs (y, dX) within 1 ulp of the oracle value (+1e-5 of the tensor's max for cancellation);
fp32 adapter gradients within 1e-3 of max|grad| (north_star: 1e-3 relative for bf16 -- a 1-ulp flip of one bf16
element of T or U is the error source) and 2e-5 when nothing 16-bit sits in between."""
import pytest
import torch

from oracle import lora as olora

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(built_lib):
    from pairwise_sample_optimization_b200 import lora
    return lora


def _mk(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


def _ulp_ok(got, want64, dtype, what, slack=None):
    """|got - want| <= 1 ulp of the output type (+ ``slack``: the rigorous bound on what 1-ulp differences in the
    16-bit intermediate T / U -- accumulated in another order -- can do to this output)."""
    w = want64.float()
    eps = 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    ulp = w.abs() * eps + 1e-5 * w.abs().max()
    if slack is not None:
        ulp = ulp + eps * slack.float().reshape(w.shape)
    bad = (got.float().cpu() - w).abs() > ulp
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} of {bad.numel()} elements off by more than 1 ulp"


def _rel(got, want64):
    return ((got.double().cpu() - want64).abs().max() / want64.abs().max()).item()


def _layer(L, K, N, r, bias, dtype, seed, alpha=None, lora_dtype=torch.float32):
    base = torch.nn.Linear(K, N, bias=bias)
    with torch.no_grad():
        base.weight.copy_(_mk((N, K), seed, K ** -0.5, torch.float32))
        if bias:
            base.bias.copy_(_mk((N,), seed + 1, 0.5, torch.float32))
    base = base.to(device="cuda", dtype=dtype)
    lay = L.LoRALinear(base, r, alpha if alpha is not None else r, lora_dtype=lora_dtype)
    with torch.no_grad():  # the reference initialises B = 0; use non-zero B so the adapter path is exercised
        lay.lora_B["default"].weight.copy_(_mk((N, r), seed + 2, 0.05, torch.float32))
    return lay


@pytest.mark.parametrize("M_shape,K,N,r,bias", [((2, 256), 640, 640, 8, False), ((4, 77), 2048, 640, 4, False),
                                                  ((2, 1024), 640, 640, 64, True), ((300,), 1280, 1280, 16, True),
                                                  ((1, 128), 320, 320, 128, True), ((2, 64), 64, 32, 4, False)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_lora_linear_forward_backward_vs_oracle(L, M_shape, K, N, r, bias, dtype):
    lay = _layer(L, K, N, r, bias, dtype, 11, alpha=2 * r)
    x = _mk((*M_shape, K), 5, 1.0, dtype).cuda().requires_grad_(True)
    dy = _mk((*M_shape, N), 6, 1.0, dtype).cuda()
    y = lay(x)
    assert y.shape == (*M_shape, N) and y.dtype == dtype
    y.backward(dy)
    A, Bm = lay.lora_A["default"].weight, lay.lora_B["default"].weight
    o = olora.lora_linear_rounded_flow(x.detach().cpu(), lay.base_layer.weight.cpu(),
                                       lay.base_layer.bias.cpu() if bias else None, A.detach().cpu(), Bm.detach().cpu(),
                                       dy.cpu(), scaling=2.0, dtype=dtype)
    A16, B16 = A.detach().to(dtype).double().cpu(), Bm.detach().to(dtype).double().cpu()
    _ulp_ok(y.detach(), o["y"], dtype, "y", slack=o["T"].abs() @ B16.abs().t())
    _ulp_ok(x.grad, o["dX"], dtype, "dX", slack=o["U"].abs() @ A16.abs())
    # (2e-3: dA / dB inherit the 1-ulp differences of the 16-bit t / u that the y / dX checks allow -- up to 1.5e-3 of max |gradient|
    # at r = 128 with ~130 rows, depending on the lora_A that the gaussian init draws; see test_gpu_lora_groups.py)
    assert A.grad.dtype == torch.float32 and _rel(A.grad, o["dA"]) <= 2e-3
    assert Bm.grad.dtype == torch.float32 and _rel(Bm.grad, o["dB"]) <= 2e-3
    # the reference's own (unrounded, fp64) arithmetic: same numbers to bf16 accuracy
    dX, dA, dB = olora.lora_linear_grads(x.detach().cpu(), lay.base_layer.weight.cpu(), A.detach().to(dtype).cpu(),
                                         Bm.detach().to(dtype).cpu(), dy.cpu(), scaling=2.0)
    tol = 8e-3 if dtype == torch.bfloat16 else 1e-3
    assert _rel(x.grad, dX) <= tol and _rel(A.grad, dA) <= tol and _rel(Bm.grad, dB) <= tol


def test_gradient_accumulates_in_place_and_bucket_is_flat(L):
    lay = _layer(L, 640, 640, 8, True, torch.bfloat16, 21)
    params = [lay.lora_A["default"].weight, lay.lora_B["default"].weight]
    bucket = L.LoRAGradBucket(params)
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, bucket.views))
    x = _mk((512, 640), 1).cuda().requires_grad_(True)
    dy = _mk((512, 640), 2).cuda()
    lay(x).backward(dy)
    once = bucket.flat.clone()
    lay(x).backward(dy)
    assert torch.allclose(bucket.flat, 2 * once, rtol=1e-5, atol=1e-5 * once.abs().max().item())  # atomics: order varies
    assert params[0].grad.data_ptr() == bucket.views[0].data_ptr()  # still the view: nothing was re-allocated
    norm = bucket.clip_grad_norm_(1.0)
    assert abs(torch.linalg.vector_norm(bucket.flat).item() - min(1.0, norm.item())) < 1e-3
    bucket.zero_()
    assert float(params[1].grad.abs().sum()) == 0.0


def test_disable_adapters_is_the_frozen_reference(L):
    lay = _layer(L, 640, 1280, 8, True, torch.bfloat16, 31)
    x = _mk((2, 200, 640), 3).cuda().requires_grad_(True)
    lay.enable_adapters(False)
    assert lay.disable_adapters
    with torch.no_grad():
        y_ref = lay(x)
    want = x.detach().double().cpu() @ lay.base_layer.weight.double().cpu().t() + lay.base_layer.bias.double().cpu()
    _ulp_ok(y_ref, want, torch.bfloat16, "reference pass")
    y = lay(x)  # with grad, adapters off: dX = dy W only, adapter grads untouched
    y.backward(torch.ones_like(y))
    assert lay.lora_A["default"].weight.grad is None
    dX = torch.ones(400, 1280, dtype=torch.float64) @ lay.base_layer.weight.double().cpu()
    _ulp_ok(x.grad.reshape(400, 640), dX, torch.bfloat16, "dX")
    lay.enable_adapters(True)
    assert not torch.equal(lay(x).detach(), y_ref)


def test_input_without_grad_skips_dx(L):
    """Cross-attention K/V read encoder_hidden_states, which carries no gradient."""
    lay = _layer(L, 2048, 640, 8, False, torch.bfloat16, 41)
    x = _mk((2, 77, 2048), 4).cuda()
    dy = _mk((2, 77, 640), 5).cuda()
    y = lay(x)
    y.backward(dy)
    A, Bm = lay.lora_A["default"].weight, lay.lora_B["default"].weight
    o = olora.lora_linear_rounded_flow(x.cpu(), lay.base_layer.weight.cpu(), None, A.detach().cpu(), Bm.detach().cpu(),
                                       dy.cpu())
    assert _rel(A.grad, o["dA"]) <= 1e-3 and _rel(Bm.grad, o["dB"]) <= 1e-3


class _Attention(torch.nn.Module):
    """The diffusers==0.27.0 ``Attention`` surface the processor reads."""

    def __init__(self, query_dim, cross_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = torch.nn.Linear(query_dim, inner, bias=False)
        self.to_k = torch.nn.Linear(cross_dim or query_dim, inner, bias=False)
        self.to_v = torch.nn.Linear(cross_dim or query_dim, inner, bias=False)
        self.to_out = torch.nn.ModuleList([torch.nn.Linear(inner, query_dim, bias=True), torch.nn.Dropout(0.0)])
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.processor = None

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


@pytest.mark.parametrize("cross", [False, True])
def test_attention_processor_with_adapters_vs_oracle(L, cross):
    torch.manual_seed(0)
    attn = _Attention(640, 2048 if cross else None, 10, 64).to(device="cuda", dtype=torch.bfloat16)
    wrapped = L.add_adapter(attn, L.LoraConfig(r=8, lora_alpha=8))
    assert len(wrapped) == 4 and isinstance(attn.to_out[0], L.LoRALinear) and isinstance(attn.to_q, L.LoRALinear)
    for i, m in enumerate(wrapped):
        with torch.no_grad():
            m.lora_B["default"].weight.copy_(_mk(tuple(m.lora_B["default"].weight.shape), 50 + i, 0.05, torch.float32))
    attn.processor = L.PSOAttnProcessor2_0()
    h = _mk((2, 256, 640), 7).cuda().requires_grad_(True)
    enc = _mk((2, 77, 2048), 8).cuda() if cross else None
    out = attn(h, encoder_hidden_states=enc)
    out.float().square().mean().backward()

    # oracle: restated peft forward per projection + restated AttnProcessor2_0 flow, fp32 on the CPU, autograd
    hp = h.detach().float().cpu().requires_grad_(True)
    mods = {"to_q": attn.to_q, "to_k": attn.to_k, "to_v": attn.to_v, "to_out": attn.to_out[0]}
    leaves = {}

    def proj(name, x):
        m = mods[name]
        A = m.lora_A["default"].weight.detach().float().cpu().requires_grad_(True)
        Bm = m.lora_B["default"].weight.detach().float().cpu().requires_grad_(True)
        leaves[name] = (A, Bm)
        b = m.base_layer.bias
        return olora.lora_linear(x, m.base_layer.weight.float().cpu(), None if b is None else b.float().cpu(), A, Bm, 1.0)

    ref = olora.attention_forward(hp, None if enc is None else enc.float().cpu(), proj, heads=10)
    ref.square().mean().backward()
    assert _rel(out.detach(), ref.detach().double()) <= 1e-2  # bf16 activations through SDPA
    assert _rel(h.grad, hp.grad.double()) <= 2e-2
    for name, m in mods.items():
        A, Bm = leaves[name]
        assert _rel(m.lora_A["default"].weight.grad, A.grad.double()) <= 2e-2, name
        assert _rel(m.lora_B["default"].weight.grad, Bm.grad.double()) <= 2e-2, name
    # policy / frozen-reference switch
    L.disable_adapters(attn)
    with torch.no_grad():
        off = attn(h, encoder_hidden_states=enc)
    L.enable_adapters(attn)
    assert not torch.equal(off, out.detach())


def test_cpu_tensor_is_refused(L):
    from pairwise_sample_optimization_b200 import _lib
    lay = _layer(L, 64, 64, 4, False, torch.bfloat16, 61)
    with pytest.raises(_lib.Psob200Error):
        lay(torch.zeros(2, 64, dtype=torch.bfloat16))


@pytest.mark.parametrize("r", [8, 4])
def test_fused_flat_optimizer_matches_torch_adamw_with_clipping(L, r):
    """psob200_flat_adamw_step (clip + AdamW + zero_grad + operand refresh, two launches) against
    torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW on the same gradients, three steps."""
    torch.manual_seed(3)
    attn = _Attention(640, 2048, 10, 64).to(device="cuda", dtype=torch.bfloat16)
    wrapped = L.add_adapter(attn, L.LoraConfig(r=r, lora_alpha=r))
    for m in wrapped:
        torch.nn.init.normal_(m.lora_B["default"].weight, std=0.05)
    ref_params = [p.detach().clone().requires_grad_(True) for p in L.lora_parameters(attn)]
    ref_opt = torch.optim.AdamW(ref_params, lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    opt = L.FusedLoRAOptimizer(attn, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=1.0)
    params = L.lora_parameters(attn)
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(opt.bucket.params, opt.bucket.views))
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in range(3):
        for p, q in zip(params, ref_params):
            grad = torch.randn(p.shape, device="cuda", generator=g) * (3.0 if step == 0 else 0.01)  # step 0 clips, later ones do not
            p.grad.copy_(grad)
            q.grad = grad.clone()
        want_norm = torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
        ref_opt.step()
        norm = opt.step()
        assert abs(norm.item() - want_norm.item()) <= 1e-5 * want_norm.item()
        assert float(opt.bucket.flat.abs().max()) == 0.0  # zero_grad fused
        for p, q in zip(params, ref_params):
            assert (p.detach() - q.detach()).abs().max().item() <= 2e-6 * max(q.detach().abs().max().item(), 1e-3)
    # the GEMM operands follow the parameters without any per-layer refresh
    for m in wrapped:
        for which, lin in (("a", m.lora_A["default"]), ("b", m.lora_B["default"])):
            op = m._operand(which, torch.bfloat16)
            assert torch.equal(op, lin.weight.detach().to(torch.bfloat16)) and op.stride(0) % 8 == 0
    # and the projection still works (forward + backward into the flat gradient)
    x = _mk((2, 64, 640), 9).cuda().requires_grad_(True)
    attn.to_q(x).float().square().mean().backward()
    assert float(opt.bucket.flat.abs().max()) > 0.0


def test_weight_gradients_on_the_side_stream_are_identical(L):
    """set_wgrad_stream(True): dA / dB leave the critical path (second stream, joined by an autograd-engine callback at the
    end of the backward); results are those of the single-stream order, eagerly and replayed from a CUDA graph."""
    dt = torch.bfloat16
    stack = [_layer(L, 320, 320, 8, True, dt, 40 + i) for i in range(3)]
    x = _mk((4, 200, 320), 7, 1.0, dt).cuda()
    dy = _mk((4, 200, 320), 8, 1.0, dt).cuda()

    def run():
        for lay in stack:
            for p in (lay.lora_A["default"].weight, lay.lora_B["default"].weight):
                if p.grad is not None:
                    p.grad.zero_()
        xg = x.clone().requires_grad_(True)
        h = xg
        for lay in stack:
            h = lay(h) * 0.5
        h.backward(dy)
        torch.cuda.synchronize()
        return [xg.grad.clone()] + [p.grad.clone() for lay in stack for p in (lay.lora_A["default"].weight, lay.lora_B["default"].weight)]

    want = run()
    L.set_wgrad_stream(True)
    try:
        got = run()
        for g, w in zip(got, want):
            assert _rel(g, w.double().cpu()) <= 2e-6  # only the order of the fp32 atomics differs
        # captured: fork / join become graph edges
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            run()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        xs = x.clone().requires_grad_(True)
        with torch.cuda.graph(graph, stream=side):
            h = xs
            for lay in stack:
                h = lay(h) * 0.5
            h.backward(dy)
        for lay in stack:
            lay.lora_A["default"].weight.grad.zero_()
            lay.lora_B["default"].weight.grad.zero_()
        graph.replay()
        torch.cuda.synchronize()
        got = [p.grad for lay in stack for p in (lay.lora_A["default"].weight, lay.lora_B["default"].weight)]
        for g, w in zip(got, want[1:]):
            assert _rel(g, w.double().cpu()) <= 2e-6
    finally:
        L.set_wgrad_stream(False)


# ----------------------------------------------------------------------------------------------- round-2 additions
def _attn_with_opt(L, r=8, seed=3, **kw):
    torch.manual_seed(seed)
    attn = _Attention(640, 2048, 10, 64).to(device="cuda", dtype=torch.bfloat16)
    wrapped = L.add_adapter(attn, L.LoraConfig(r=r, lora_alpha=r))
    for m in wrapped:
        torch.nn.init.normal_(m.lora_B["default"].weight, std=0.05)
    opt = L.FusedLoRAOptimizer(attn, lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0, **kw)
    return attn, wrapped, opt


def test_optimizer_state_round_trip_resumes_bit_identically(L, tmp_path):
    """accelerator.save_state / load_state (turbo :889) also carry the AdamW moments and the step count: a run resumed from
    checkpoint.save_state continues exactly like the uninterrupted one (the state used to be lost)."""
    from pairwise_sample_optimization_b200 import checkpoint
    attn, _, opt = _attn_with_opt(L)
    g = torch.Generator(device="cuda").manual_seed(11)
    grads = [torch.randn(opt.bucket.flat.numel(), device="cuda", generator=g) * 0.02 for _ in range(4)]
    for k in range(2):
        opt.bucket.flat.copy_(grads[k])
        opt.step()
    checkpoint.save_state(str(tmp_path), attn, opt)
    attn2, wrapped2, opt2 = _attn_with_opt(L, seed=99)  # different adapters, fresh moments
    assert not torch.equal(opt2.flat_param, opt.flat_param)
    checkpoint.load_state(str(tmp_path), attn2, opt2)
    assert int(opt2.step_dev.item()) == 2 and torch.equal(opt2.flat_param, opt.flat_param)
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and torch.equal(opt2.exp_avg_sq, opt.exp_avg_sq)
    for m in wrapped2:  # the 16-bit operands the GEMM kernels read follow the loaded parameters
        for which, lin in (("a", m.lora_A["default"]), ("b", m.lora_B["default"])):
            assert torch.equal(m._operand(which, torch.bfloat16), lin.weight.detach().to(torch.bfloat16))
    for k in (2, 3):
        for o in (opt, opt2):
            o.bucket.flat.copy_(grads[k])
            o.step()
    assert torch.equal(opt2.flat_param, opt.flat_param) and torch.equal(opt2.exp_avg_sq, opt.exp_avg_sq)
    # a checkpoint of another layout is refused
    _, _, opt3 = _attn_with_opt(L, r=4)
    with pytest.raises(Exception):
        checkpoint.load_optimizer_state(str(tmp_path), opt3)


def test_zero_grad_set_to_none_keeps_the_flat_bucket_attached(L):
    """torch's default zero_grad(set_to_none=True) sets param.grad = None; the weight-gradient kernels must keep writing into
    the parameter's slice of the flat bucket (they used to accumulate into private tensors, silently)."""
    attn, wrapped, opt = _attn_with_opt(L)
    for p in L.lora_parameters(attn):
        p.grad = None
    x = _mk((2, 64, 640), 9).cuda().requires_grad_(True)
    attn.to_q(x).float().square().mean().backward()
    assert float(opt.bucket.flat.abs().max()) > 0.0
    pa = wrapped[0].lora_A["default"].weight
    q = [m for m in wrapped if m is attn.to_q][0]
    assert q.lora_A["default"].weight.grad.data_ptr() == q.lora_A["default"].weight._psob200_grad_view.data_ptr()
    opt.zero_grad(set_to_none=True)
    assert float(opt.bucket.flat.abs().max()) == 0.0 and pa.grad is not None
    # a foreign tensor in param.grad is an error, not a silent no-op
    q.lora_A["default"].weight.grad = torch.zeros_like(q.lora_A["default"].weight)
    with pytest.raises(Exception, match="flat LoRA gradient bucket"):
        attn.to_q(x).float().square().mean().backward()


def test_nonfinite_gradient_norm_skips_the_update_and_grad_scale_unscales(L):
    """fp16 loss scaling (accelerate's GradScaler around turbo :857-860): grad_scale unscales inside the fused step; an
    overflowed (inf / NaN) gradient skips the whole update, reports found_inf and does not advance the step count."""
    attn, _, opt = _attn_with_opt(L)
    _, _, ref = _attn_with_opt(L)
    g = torch.Generator(device="cuda").manual_seed(2)
    grad = torch.randn(opt.bucket.flat.numel(), device="cuda", generator=g) * 0.01
    scale = 1024.0
    opt.bucket.flat.copy_(grad * scale)
    opt.grad_scale = 1.0 / scale
    ref.bucket.flat.copy_(grad)
    n1, n2 = opt.step().clone(), ref.step().clone()
    assert abs(n1.item() - n2.item()) <= 1e-6 * n2.item() and float(opt.found_inf) == 0.0
    assert (opt.flat_param - ref.flat_param).abs().max().item() <= 1e-7
    before = (opt.flat_param.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), opt.flat_operand.clone())
    for bad in (float("inf"), float("nan")):
        opt.bucket.flat.copy_(grad * scale)
        opt.bucket.flat[12345] = bad
        norm = opt.step()
        assert not torch.isfinite(norm).item() and float(opt.found_inf) == 1.0 and int(opt.step_dev.item()) == 1
        assert float(opt.bucket.flat.abs().max()) == 0.0  # zero_grad still happens
        for a, b in zip(before, (opt.flat_param, opt.exp_avg, opt.exp_avg_sq, opt.flat_operand)):
            assert torch.equal(a, b)
    # the next finite step continues with bias corrections of step 2 -- same as the run that never overflowed
    for o, s in ((opt, scale), (ref, 1.0)):
        o.bucket.flat.copy_(grad * s)
        o.step()
    assert (opt.flat_param - ref.flat_param).abs().max().item() <= 1e-7 and float(opt.found_inf) == 0.0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tensors_on_a_second_device_launch_there(L):
    """Per-device kernel configuration + device guard: cuda:1 tensors while cuda:0 is current."""
    assert torch.cuda.current_device() == 0
    dt = torch.bfloat16
    lay0 = _layer(L, 640, 640, 8, True, dt, 1)
    torch.manual_seed(0)
    lin = torch.nn.Linear(640, 640, bias=True).to(device="cuda:1", dtype=dt)
    lay1 = L.LoRALinear(lin, 8, 8)
    with torch.no_grad():
        lay1.lora_B["default"].weight.normal_(std=0.05)
    x = _mk((2, 300, 640), 7, 1.0, dt)
    y0 = lay0(x.cuda())  # configures the kernels on device 0 first
    y1 = lay1(x.to("cuda:1"))
    want = torch.nn.functional.linear(x.to("cuda:1").float(), lin.weight.float(), lin.bias.float()) + \
        (x.to("cuda:1").float() @ lay1.lora_A["default"].weight.t()) @ lay1.lora_B["default"].weight.t()
    assert y1.device.index == 1 and y0.device.index == 0
    assert (y1.float() - want).abs().max().item() <= 2e-2 * want.abs().max().item()
