"""Stacked LoRA projections (psob200_lora_group_*; lora.LoRAProjectionGroup / fuse_attention_projections): G projections that
share their input as ONE problem list per direction -- t / u as tiles of the same launch as y / dx (flags per row block),
q / k / v stacked along N, dx as one reduction over the G gradients, dA + dB_g as one launch -- against the per-projection
oracle (oracle.lora.lora_linear_rounded_flow: the peft flow with the 16-bit rounding points of a half-precision run) and
against the per-projection product path."""
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


def _layers(L, G, K, N, r, dtype, seed):
    out = []
    for g in range(G):
        base = torch.nn.Linear(K, N, bias=False)
        with torch.no_grad():
            base.weight.copy_(_mk((N, K), seed + 10 * g, K ** -0.5, torch.float32))
        lay = L.LoRALinear(base.to(device="cuda", dtype=dtype), r, 2 * r)
        with torch.no_grad():
            lay.lora_B["default"].weight.copy_(_mk((N, r), seed + 10 * g + 2, 0.05, torch.float32))
        out.append(lay)
    return out


def _ulp_ok(got, want64, dtype, what, slack=None):
    w = want64.float()
    eps = 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    ulp = w.abs() * eps + 1e-5 * w.abs().max()
    if slack is not None:
        ulp = ulp + eps * slack.float().reshape(w.shape)
    bad = (got.float().cpu() - w).abs() > ulp
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} of {bad.numel()} elements off by more than 1 ulp"


def _rel(got, want64):
    return ((got.double().cpu() - want64).abs().max() / want64.abs().max()).item()


@pytest.mark.parametrize("G,M_shape,K,N,r,need_dx", [(3, (2, 300), 640, 640, 8, True), (3, (4, 256), 1280, 1280, 64, True),
                                                     (2, (4, 77), 2048, 640, 16, False), (2, (3, 77), 2048, 1280, 64, True),
                                                     (3, (1, 130), 320, 320, 128, True), (3, (8, 1024), 1280, 1280, 64, True),
                                                     (3, (2, 1024), 640, 640, 8, True),
                                                     # ragged everything: M, K and N off the tile / k-block / box sizes, tiles that
                                                     # would straddle two stacked projections, reduction tails inside the stacked weight
                                                     (2, (3, 77), 72, 200, 8, True), (3, (1, 300), 136, 264, 16, True),
                                                     (3, (1, 5), 64, 48, 8, True),
                                                     # ranks that are not multiples of 8 (BASELINE config 4: r = 4): stacked with
                                                     # 8-wide column groups, zero rows in the stacked lora_a
                                                     (3, (2, 300), 640, 640, 4, True), (2, (4, 77), 2048, 640, 12, False),
                                                     (3, (2, 1024), 1280, 1280, 4, True)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_stacked_group_forward_backward_vs_oracle(L, G, M_shape, K, N, r, need_dx, dtype):
    layers = _layers(L, G, K, N, r, dtype, 3)
    group = L.LoRAProjectionGroup(layers)
    x = _mk(M_shape + (K,), 5, 1.0, dtype).cuda().requires_grad_(need_dx)
    dys = [_mk(M_shape + (N,), 7 + g, 1.0, dtype).cuda() for g in range(G)]
    ys = group(x)
    assert len(ys) == G and ys[0].shape == M_shape + (N,)
    torch.autograd.backward(list(ys), dys)
    torch.cuda.synchronize()
    assert int(L._flags_workspace(torch.device("cuda", 0)).abs().sum()) == 0  # every launch leaves the dependency flags zeroed
    dx_want = 0
    slack_dx = 0
    for g, lay in enumerate(layers):
        A, Bm = lay.lora_A["default"].weight, lay.lora_B["default"].weight
        o = olora.lora_linear_rounded_flow(x.detach().cpu(), lay.base_layer.weight.cpu(), None, A.detach().cpu(), Bm.detach().cpu(),
                                           dys[g].cpu(), scaling=2.0, dtype=dtype)
        # y: 1 ulp of the output type (+ what 1-ulp differences of the 16-bit t can do: |B| summed over r per unit of t's ulp)
        slack_y = (o["T"].abs() @ Bm.detach().cpu().to(dtype).double().abs().t())
        _ulp_ok(ys[g].detach().reshape(-1, N), o["y"].reshape(-1, N), dtype, f"y[{g}]", slack_y)
        dx_want = dx_want + o["dX"]
        slack_dx = slack_dx + (o["U"].abs() @ A.detach().cpu().to(dtype).double().abs())
        # dA / dB inherit the 1-ulp differences of the 16-bit t / u that the y / dx checks allow (fp32 vs fp64 accumulation before the
        # rounding): 1.0e-3 - 1.5e-3 of max |gradient| at r = 128 with 130 rows, depending on the lora_A that LoRALinear's gaussian init
        # draws (80 runs on the SAME inputs give bit-identical y, t^T and dx, and dA / dB that differ by 1e-7: split-K atomics order)
        assert _rel(A.grad, o["dA"]) <= 2e-3 and _rel(Bm.grad, o["dB"]) <= 2e-3, (g, _rel(A.grad, o["dA"]), _rel(Bm.grad, o["dB"]))
    if need_dx:
        _ulp_ok(x.grad.reshape(-1, K), dx_want.reshape(-1, K), dtype, "dx", slack_dx)
    else:
        assert x.grad is None


def test_fused_launches_equal_the_separate_launch_sequence(L):
    """The in-launch dependency (flags) changes WHEN t / u are computed, not how: y and dx are bit-identical to the
    launch-by-launch sequence (t, y / u, dx as launches of their own, ordered by the stream)."""
    dtype = torch.bfloat16
    outs = []
    for deps in (True, False):
        torch.manual_seed(21)  # LoRALinear draws A from the global generator
        layers = _layers(L, 3, 1280, 1280, 64, dtype, 11)
        group = L.LoRAProjectionGroup(layers)
        x = _mk((2, 1024, 1280), 5, 1.0, dtype).cuda().requires_grad_(True)
        dys = [_mk((2, 1024, 1280), 7 + g, 1.0, dtype).cuda() for g in range(3)]
        sink = []
        L.set_in_launch_dependencies(deps)
        L.set_timing_sink(sink)
        try:
            ys = group(x)
            torch.autograd.backward(list(ys), dys)
        finally:
            L.set_timing_sink(None)
            L.set_in_launch_dependencies(True)
        torch.cuda.synchronize()
        assert len(sink) == 3  # forward, input gradient, weight gradients: three host calls for three projections
        outs.append([y.detach().clone() for y in ys] + [x.grad.clone()] +
                    [l.lora_A["default"].weight.grad.clone() for l in layers])
    for a, b in zip(outs[0][:4], outs[1][:4]):
        assert torch.equal(a, b)
    for a, b in zip(outs[0][4:], outs[1][4:]):  # split reductions accumulate with fp32 atomics: order-dependent in the last bits
        assert (a - b).abs().max().item() <= 1e-4 * b.abs().max().item()


@pytest.mark.parametrize("rank", [8, 4])
def test_processor_with_fused_projections_matches_the_per_projection_path(L, rank):
    from tests.test_gpu_lora import _Attention
    dtype = torch.bfloat16
    res = []
    for fuse in (False, True):
        for cross in (None, 2048):
            torch.manual_seed(3)
            attn = _Attention(640, cross, 10, 64).to(device="cuda", dtype=dtype)
            wrapped = L.add_adapter(attn, L.LoraConfig(r=rank, lora_alpha=rank))
            g = torch.Generator().manual_seed(4)
            for m in wrapped:
                with torch.no_grad():
                    m.lora_B["default"].weight.copy_(torch.randn(m.lora_B["default"].weight.shape, generator=g) * 0.05)
            if fuse:
                assert L.fuse_attention_projections(attn) == 1
            attn.processor = L.PSOAttnProcessor2_0()
            opt = L.FusedLoRAOptimizer(attn)
            if fuse:  # the flat layout keeps a group's matrices adjacent: the stacked operands are views, not copies
                grp = L.projection_groups(attn)[0]
                if rank % 8 == 0:
                    assert grp.stacked_operand("a", dtype).data_ptr() == grp.layers[0]._operand("a", dtype).data_ptr()
                else:  # padded private copies: [G * 8, K] with zero rows behind each projection's r rows
                    sa = grp.stacked_operand("a", dtype)
                    assert sa.shape[0] == grp.G * 8 and float(sa[rank:8].abs().max()) == 0.0
                assert grp.stacked_grad("b").data_ptr() == grp.layers[0].lora_B["default"].weight.grad.data_ptr()
            x = _mk((2, 256, 640), 9, 1.0, dtype).cuda().requires_grad_(True)
            enc = None if cross is None else _mk((2, 77, 2048), 10, 1.0, dtype).cuda()
            y = attn(x, encoder_hidden_states=enc)
            y.backward(_mk((2, 256, 640), 12, 1.0, dtype).cuda())
            torch.cuda.synchronize()
            grads = {n: p.grad.clone() for n, p in attn.named_parameters() if p.requires_grad}
            res.append((y.detach().clone(), x.grad.clone(), grads))
            L.disable_adapters(attn)
            with torch.no_grad():
                y0 = attn(x, encoder_hidden_states=enc)
            L.enable_adapters(attn)
            res[-1] = res[-1] + (y0.clone(),)
            if fuse:  # the optimizer boundary keeps the stacked operands current (also the padded private copies)
                opt.lr = 1e-2
                opt.step()
                grp = L.projection_groups(attn)[0]
                for g, lay in enumerate(grp.layers):
                    rs = grp.r_stride
                    assert torch.equal(grp.stacked_operand("a", dtype)[g * rs:g * rs + rank], lay.lora_A["default"].weight.detach().to(dtype))
                    n = lay.out_features
                    assert torch.equal(grp.stacked_operand("b", dtype)[g * n:(g + 1) * n], lay.lora_B["default"].weight.detach().to(dtype))
    for k in (0, 1):  # self-attention, cross-attention
        (y_a, dx_a, g_a, y0_a), (y_b, dx_b, g_b, y0_b) = res[k], res[2 + k]
        assert torch.equal(y0_a, y0_b)  # the frozen-reference pass: same reduction order per element, any tile width
        assert (y_a.float() - y_b.float()).abs().max().item() <= 2e-2 * y_a.float().abs().max().item()
        assert (dx_a.float() - dx_b.float()).abs().max().item() <= 2e-2 * dx_a.float().abs().max().item()
        assert set(g_a) == set(g_b)
        for n in g_a:
            assert (g_a[n] - g_b[n]).abs().max().item() <= 2e-2 * g_a[n].abs().max().item() + 1e-6, n


def test_deterministic_weight_gradients_are_bit_reproducible(L):
    """set_deterministic_wgrad(True): no split-K, one accumulation per gradient element and launch -> identical bits run to run
    (the default split reductions agree only to fp32 round-off), and the same values up to that round-off."""
    dtype = torch.bfloat16
    runs = []
    for det in (True, True, False):
        torch.manual_seed(21)
        layers = _layers(L, 3, 640, 640, 16, dtype, 11)
        group = L.LoRAProjectionGroup(layers)
        x = _mk((4, 1000, 640), 5, 1.0, dtype).cuda().requires_grad_(True)
        dys = [_mk((4, 1000, 640), 7 + g, 1.0, dtype).cuda() for g in range(3)]
        L.set_deterministic_wgrad(det)
        try:
            for _ in range(2):  # two accumulating passes (gradient accumulation)
                torch.autograd.backward(list(group(x)), dys)
        finally:
            L.set_deterministic_wgrad(False)
        torch.cuda.synchronize()
        runs.append([l.lora_A["default"].weight.grad.clone() for l in layers] + [l.lora_B["default"].weight.grad.clone() for l in layers])
    for a, b in zip(runs[0], runs[1]):
        assert torch.equal(a, b)
    for a, c in zip(runs[0], runs[2]):
        assert (a - c).abs().max().item() <= 1e-4 * c.abs().max().item()


@pytest.mark.parametrize("rank", [8, 4])
def test_cross_attention_kv_bank_matches_the_per_block_launches(L, rank):
    """lora.fuse_cross_attention_kv: the k / v projections of every cross-attention layer (same prompt embeddings) as ONE forward
    launch per shape, backward per layer.  Tiny SDXL-architecture UNet: frozen pass bit-identical, policy pass and every adapter
    gradient equal to the per-block path (same reductions per element; the split-K weight-gradient sums to fp32 round-off), the
    stacked adapter operands are views of the flat buffers (no copies) for ranks that need no padding, and an optimizer step
    reaches the bank's operands."""
    from fixtures import sdxl_unet
    cfg = sdxl_unet.tiny_config()
    dtype = torch.bfloat16
    res = []
    for bank in (False, True):
        torch.manual_seed(0)
        unet = sdxl_unet.UNet2DConditionModel(cfg).to(dtype).cuda()
        wrapped = L.add_adapter(unet, L.LoraConfig(r=rank, lora_alpha=rank))
        g = torch.Generator().manual_seed(1)
        for m in wrapped:
            with torch.no_grad():
                m.lora_B["default"].weight.copy_(torch.randn(m.lora_B["default"].weight.shape, generator=g) * 0.05)
        unet.train()
        unet.set_attn_processor(L.PSOAttnProcessor2_0())
        assert L.fuse_attention_projections(unet) > 0
        if bank:
            n_banks = L.fuse_cross_attention_kv(unet)
            assert n_banks >= 1
            assert L.fuse_cross_attention_kv(unet) == n_banks  # idempotent
        L.set_deterministic_wgrad(True)
        try:
            opt = L.FusedLoRAOptimizer(unet, lr=1e-2, weight_decay=0.0, max_grad_norm=1e9)
            if bank:
                for b in unet.__dict__["_psob200_kv_banks"]:
                    assert len(b.members) > 1
                    l0 = b.all.layers[0]
                    assert b.all.stacked_weight().data_ptr() == l0.base_layer.weight.data_ptr()
                    for _, grp in b.members:  # the per-layer groups find their members inside the bank's matrix: no second copy
                        assert grp.stacked_weight().data_ptr() == grp.layers[0].base_layer.weight.data_ptr()
                    if rank % 8 == 0:
                        assert b.all.stacked_operand("a", dtype).data_ptr() == l0._operand("a", dtype).data_ptr()
                        assert not b.all.has_private_operands()
            B = 2
            x = _mk((B, 4, 16, 16), 3, 1.0, dtype).cuda()
            enc = _mk((B, 77, cfg.cross_attention_dim), 4, 1.0, dtype).cuda()
            pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
            cond = {"text_embeds": _mk((B, pooled), 5, 1.0, dtype).cuda(), "time_ids": torch.ones(B, 6, device="cuda", dtype=dtype)}
            ts = torch.tensor([999, 499], device="cuda")
            out = unet(x, ts, enc, added_cond_kwargs=cond).sample
            if bank:  # every cross-attention module got its k / v from the bank launch of THIS forward
                stashed = [m.__dict__["_psob200_kv"] for m in unet.modules() if "_psob200_kv" in m.__dict__]
                assert len(stashed) == sum(len(b.members) for b in unet.__dict__["_psob200_kv_banks"]) and all(p[0] is enc for p in stashed)
            if bank:  # ... and a later direct call of a block with the same tensor does NOT reuse them (the adapters may have changed)
                m = next(mm for mm in unet.modules() if "_psob200_kv" in mm.__dict__)
                e0, k0, v0, state = m.__dict__["_psob200_kv"]
                assert state["live"] is False
                m.__dict__["_psob200_kv"] = (e0, torch.full_like(k0, float("nan")), torch.full_like(v0, float("nan")), state)
                h = _mk((B, 64, m.to_q.in_features), 8, 1.0, dtype).cuda()
                with torch.no_grad():
                    direct = m(h, encoder_hidden_states=enc)
                assert direct.shape == h.shape and bool(torch.isfinite(direct).all())
            out.backward(_mk(tuple(out.shape), 6, 1.0, dtype).cuda())
            torch.cuda.synchronize()
            grads = opt.bucket.flat.clone()
            names = [id(p) for p in L.lora_parameters(unet)]
            by_name = {n: p.grad.clone() for n, p in unet.named_parameters() if p.requires_grad}
            L.disable_adapters(unet)
            with torch.no_grad():
                out0 = unet(x, ts, enc, added_cond_kwargs=cond).sample
            L.enable_adapters(unet)
            opt.step()  # the optimizer boundary must reach the operands the bank launch reads
            with torch.no_grad():
                out1 = unet(x, ts, enc, added_cond_kwargs=cond).sample
            torch.cuda.synchronize()
            res.append((out.detach().clone(), out0.clone(), by_name, out1.clone(), len(names)))
        finally:
            L.set_deterministic_wgrad(False)
    (y_a, y0_a, g_a, y1_a, n_a), (y_b, y0_b, g_b, y1_b, n_b) = res
    assert n_a == n_b
    assert torch.equal(y0_a, y0_b)                      # frozen reference pass
    assert torch.equal(y_a, y_b)                        # policy pass: the same reductions per element
    assert set(g_a) == set(g_b)
    worst = max(((g_a[n] - g_b[n]).abs().max().item() / (g_a[n].abs().max().item() + 1e-12)) for n in g_a)
    assert worst <= 2e-2, worst                         # (the stock attention / convolution backward kernels are not bit-reproducible)
    assert any(v.abs().max().item() > 0 for v in g_a.values())
    assert not torch.equal(y1_a, y_a)                   # the step changed the adapters ...
    assert (y1_a.float() - y1_b.float()).abs().max().item() <= 2e-2 * y1_a.float().abs().max().item()  # ... the same way in both


def test_cross_attention_kv_bank_with_gradient_checkpointing(L):
    """Checkpointed blocks (turbo trainer :358) run their forward twice: both passes take the k / v the bank launch of that forward
    produced (same autograd structure: non-reentrant checkpointing checks it) -- the adapter gradients must be those of the plain
    run (each exactly once)."""
    from fixtures import sdxl_unet
    cfg = sdxl_unet.tiny_config()
    dtype = torch.bfloat16
    res = []
    for ckpt in (False, True):
        torch.manual_seed(0)
        unet = sdxl_unet.UNet2DConditionModel(cfg).to(dtype).cuda()
        wrapped = L.add_adapter(unet, L.LoraConfig(r=8, lora_alpha=8))
        g = torch.Generator().manual_seed(1)
        for m in wrapped:
            with torch.no_grad():
                m.lora_B["default"].weight.copy_(torch.randn(m.lora_B["default"].weight.shape, generator=g) * 0.05)
        unet.train()
        unet.set_attn_processor(L.PSOAttnProcessor2_0())
        assert L.fuse_attention_projections(unet) > 0 and L.fuse_cross_attention_kv(unet) >= 1
        if ckpt:
            unet.enable_gradient_checkpointing()
        L.set_deterministic_wgrad(True)
        try:
            opt = L.FusedLoRAOptimizer(unet)
            B = 2
            x = _mk((B, 4, 16, 16), 3, 1.0, dtype).cuda()
            enc = _mk((B, 77, cfg.cross_attention_dim), 4, 1.0, dtype).cuda()
            pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
            cond = {"text_embeds": _mk((B, pooled), 5, 1.0, dtype).cuda(), "time_ids": torch.ones(B, 6, device="cuda", dtype=dtype)}
            out = unet(x, torch.tensor([999, 499], device="cuda"), enc, added_cond_kwargs=cond).sample
            out.backward(_mk(tuple(out.shape), 6, 1.0, dtype).cuda())
            torch.cuda.synchronize()
            res.append((out.detach().clone(), opt.bucket.flat.clone()))
        finally:
            L.set_deterministic_wgrad(False)
    (y_a, g_a), (y_b, g_b) = res
    assert torch.equal(y_a, y_b)
    assert g_a.abs().max().item() > 0
    cos = torch.dot(g_a.double(), g_b.double()) / (g_a.double().norm() * g_b.double().norm())
    assert cos.item() >= 0.9995 and abs(g_b.norm().item() / g_a.norm().item() - 1.0) <= 1e-2, (cos.item(), g_a.norm().item(), g_b.norm().item())


def test_many_stacked_projections_forward_with_row_block_fastest_tile_order(L):
    """The forward takes any number of stacked projections (a cross-attention bank: here 20 x [1280, 2048] = 105 MB of frozen
    weights against 616 rows).  With a weight matrix larger than L2 can hold and only a few row blocks, the tiles are walked row
    blocks fastest (every weight tile read once); the results must be bit-identical to the 20 single-projection launches, with and
    without adapters, and t^T (the operand of the per-layer dB launches) must be the transpose of t."""
    dtype = torch.bfloat16
    G, M, K, N, r = 20, 616, 2048, 1280, 64
    layers = _layers(L, G, K, N, r, dtype, 31)
    x = _mk((8, 77, K), 5, 1.0, dtype).cuda()
    with torch.no_grad():
        singles = [lay(x).clone() for lay in layers]           # per-projection launches (column tiles fastest, weights fit L2)
        for lay in layers: lay.enable_adapters(False)
        frozen_singles = [lay(x).clone() for lay in layers]
        for lay in layers: lay.enable_adapters(True)
    group = L.LoRAProjectionGroup(layers)
    assert group.stacked_weight().numel() * 2 > 48e6
    with torch.no_grad():
        outs = group(x)
    assert len(outs) == G
    for g in range(G):
        assert torch.equal(outs[g], singles[g]), g
    x2, y, tt, enabled = L._group_forward_launch(x, group, True)   # grad mode: also writes t^T for the weight-gradient launches
    torch.cuda.synchronize()
    assert enabled and tt is not None and tt.shape[0] == G * r
    assert torch.equal(y.view(8, 77, G * N)[..., :N], singles[0])
    t_ref = (x2.float() @ torch.cat([l._operand("a", dtype) for l in layers]).float().t() * float(layers[0].scaling["default"]))
    got = tt[:, :M].float().t()
    assert (got - t_ref).abs().max().item() <= 2.0 ** -7 * t_ref.abs().max().item()
    with torch.no_grad():
        for lay in layers: lay.enable_adapters(False)
        outs0 = group(x)
        for lay in layers: lay.enable_adapters(True)
    for g in range(G):
        assert torch.equal(outs0[g], frozen_singles[g]), g
