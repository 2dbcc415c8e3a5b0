"""Host-side logic of the LoRA path that needs no GPU: adapter injection by peft-style suffix matching, the flat
gradient bucket and its single all-reduce (gloo, world_size 2), the fixture's SDXL inventory, the oracle micro-step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fixtures import micro_step, sdxl_unet
from oracle import lora as olora, losses as olosses, schedules


def test_sdxl_architecture_inventory():
    """The scaffold reproduces the counts SURVEY.md derives for the real SDXL UNet."""
    with torch.device("meta"):
        big = sdxl_unet.UNet2DConditionModel(sdxl_unet.sdxl_config())
    assert sdxl_unet.count_lora_targets(big) == 560
    assert sum(1 for m in big.modules() if isinstance(m, sdxl_unet.BasicTransformerBlock)) == 70
    tot = sum(m.in_features + m.out_features for n, m in big.named_modules() if isinstance(m, torch.nn.Linear) and
              any(n.endswith(s) for s in (".to_q", ".to_k", ".to_v", ".to_out.0")))
    assert tot == 1451520  # x rank = trainable parameters
    assert abs(sum(p.numel() for p in big.parameters()) / 1e9 - 2.567) < 0.01


def test_add_adapter_matches_peft_suffix_rules(built_lib):
    from pairwise_sample_optimization_b200 import _lib, lora
    unet = sdxl_unet.UNet2DConditionModel(sdxl_unet.tiny_config())
    wrapped = lora.add_adapter(unet, lora.LoraConfig(r=4, lora_alpha=4))
    assert len(wrapped) == 96
    names = [n for n, m in unet.named_modules() if isinstance(m, lora.LoRALinear)]
    assert all(n.endswith((".to_q", ".to_k", ".to_v", ".to_out.0")) for n in names)
    assert not any(".ff." in n or "proj_in" in n or "proj_out" in n for n in names)
    m = wrapped[0]
    assert m.scaling["default"] == 1.0 and m.r["default"] == 4 and not m.disable_adapters
    assert float(m.lora_B["default"].weight.detach().abs().sum()) == 0.0  # peft gaussian init: B = 0
    assert abs(float(m.lora_A["default"].weight.detach().std()) - 0.25) < 0.1  # std = 1/r
    trainable = [p for p in unet.parameters() if p.requires_grad]
    frozen_targets = [m.base_layer.weight.requires_grad for m in wrapped]
    assert not any(frozen_targets) and len(lora.lora_parameters(unet)) == 192
    lora.disable_adapters(unet)
    assert all(w.disable_adapters for w in wrapped)
    lora.enable_adapters(unet)
    assert not any(w.disable_adapters for w in wrapped)
    with pytest.raises(_lib.Psob200Error):  # no CPU fallback
        m(torch.zeros(2, m.in_features))
    with pytest.raises(ValueError):
        lora.add_adapter(torch.nn.Sequential(torch.nn.Linear(4, 4)), lora.LoraConfig(target_modules=["to_q"]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pairwise_sample_optimization_b200 import lora
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(4, 6)), torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(2, 2))]
    bucket = lora.LoRAGradBucket(params)
    for i, p in enumerate(params):  # what the dA / dB kernels do: accumulate into the views in place
        p.grad.add_(torch.full_like(p, float((rank + 1) * (i + 1))))
    bucket.all_reduce()
    want = [(1 + 2) / 2 * (i + 1) for i in range(3)]
    ok = all(torch.allclose(p.grad, torch.full_like(p, w)) for p, w in zip(params, want))
    ok = ok and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, bucket.views))
    norm = bucket.clip_grad_norm_(1.0)
    ok = ok and abs(float(torch.linalg.vector_norm(bucket.flat)) - 1.0) < 1e-4 and float(norm) > 1.0
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_bucket_single_allreduce_gloo_world2(built_lib):
    world = 2
    with mp.Manager() as man:
        out = man.dict()
        mp.spawn(_bucket_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_oracle_micro_step_on_tiny_unet_cpu():
    torch.manual_seed(0)
    cfg = sdxl_unet.tiny_config()
    unet = sdxl_unet.UNet2DConditionModel(cfg)
    wrapped = olora.oracle_add_adapter(unet, 4, 4)
    for m in wrapped:
        torch.nn.init.normal_(m.lora_B["default"].weight, std=0.02)
    sched = schedules.turbo_scheduler(4)
    batch = micro_step.synth_batch(2, 32, cfg.cross_attention_dim, 32, 3, sched.sigmas)
    loss = micro_step.oracle_micro_step(olora, olosses, unet, batch, sched, beta=5.0, eps=0.9)
    assert torch.isfinite(loss)
    g = torch.cat([m.lora_A["default"].weight.grad.flatten() for m in wrapped])
    assert torch.isfinite(g).all() and g.abs().max() > 0
    assert all(m.base_layer.weight.grad is None for m in wrapped)


def test_lora_checkpoint_wire_format_round_trip(built_lib, tmp_path):
    """pytorch_lora_weights.safetensors with the diffusers key naming the reference's save hook produces
    (train_online_pso_sdxl_turbo.py:361-379), and back."""
    from safetensors import safe_open
    from pairwise_sample_optimization_b200 import checkpoint, lora
    torch.manual_seed(0)
    a = sdxl_unet.UNet2DConditionModel(sdxl_unet.tiny_config())
    wa = lora.add_adapter(a, lora.LoraConfig(r=4, lora_alpha=4))
    for m in wa:
        torch.nn.init.normal_(m.lora_B["default"].weight, std=0.02)
    path = checkpoint.save_lora_weights(str(tmp_path), a)
    assert os.path.basename(path) == "pytorch_lora_weights.safetensors"
    with safe_open(path, framework="pt") as f:
        keys = sorted(f.keys())
        assert f.metadata() == {"format": "pt"}
    assert len(keys) == 192 and all(k.startswith("unet.") for k in keys)
    assert "unet.down_blocks.1.attentions.0.transformer_blocks.0.attn1.to_q.lora.down.weight" in keys
    assert "unet.mid_block.attentions.0.transformer_blocks.1.attn2.to_out.0.lora.up.weight" in keys
    b = sdxl_unet.UNet2DConditionModel(sdxl_unet.tiny_config())
    wb = lora.add_adapter(b, lora.LoraConfig(r=4, lora_alpha=4))
    assert len(checkpoint.load_lora_weights(str(tmp_path), b)) == 192
    for ma, mb in zip(wa, wb):
        assert torch.equal(ma.lora_A["default"].weight, mb.lora_A["default"].weight)
        assert torch.equal(ma.lora_B["default"].weight, mb.lora_B["default"].weight)
    c = sdxl_unet.UNet2DConditionModel(sdxl_unet.tiny_config())
    lora.add_adapter(c, lora.LoraConfig(r=8, lora_alpha=8))
    with pytest.raises(ValueError):
        checkpoint.load_lora_weights(path, c)


def test_host_logic_of_the_late_additions(built_lib):
    """CPU-checkable parts of the pieces added late in the round: the exchange refuses to exist without a process group, the
    GEGLU installer finds the feed-forward modules of the fixture, the product path still refuses CPU tensors."""
    import pytest
    import torch
    from fixtures import sdxl_unet
    from pairwise_sample_optimization_b200 import _lib, feed_forward, lora
    with pytest.raises(_lib.Psob200Error):
        lora.SymmetricGradExchange()  # no torch.distributed group initialised in this process
    unet = sdxl_unet.UNet2DConditionModel(sdxl_unet.tiny_config())
    n_blocks = sum(1 for m in unet.modules() if type(m).__name__ == "BasicTransformerBlock")
    assert feed_forward.install_fused_geglu(unet) == n_blocks > 0
    with pytest.raises(_lib.Psob200Error):
        feed_forward.geglu(torch.randn(2, 16))
    lora.set_wgrad_stream(True)
    lora.set_wgrad_stream(False)


def test_stacking_and_flat_layout_host_logic(built_lib):
    """fuse_attention_projections / fuse_cross_attention_kv need no GPU for their bookkeeping: which modules are stacked how,
    the frozen weights as views of one matrix (values unchanged), and lora_parameters() ordering the flat buffers so that a
    group's -- and a bank's -- matrices are adjacent."""
    from pairwise_sample_optimization_b200 import lora
    torch.manual_seed(0)
    unet = sdxl_unet.UNet2DConditionModel(sdxl_unet.tiny_config())
    wrapped = lora.add_adapter(unet, lora.LoraConfig(r=8, lora_alpha=8))
    before = {id(m): m.base_layer.weight.detach().clone() for m in wrapped}
    plain_order = [id(p) for p in lora.lora_parameters(unet)]
    n_groups = lora.fuse_attention_projections(unet)
    attn = [m for m in unet.modules() if isinstance(m, sdxl_unet.Attention)]
    assert n_groups == len(attn) and lora.fuse_attention_projections(unet) == 0  # every module once
    for m in attn:  # cross-attention is stacked as k / v even where its widths equal the query's (tiny fixture: 64 = 64)
        assert set(m.__dict__["_psob200_groups"]) == ({"kv"} if m.is_cross_attention else {"qkv"})
    n_banks = lora.fuse_cross_attention_kv(unet)
    assert n_banks >= 1 and lora.fuse_cross_attention_kv(unet) == n_banks
    banks = unet.__dict__["_psob200_kv_banks"]
    cross = [m for m in attn if m.is_cross_attention]
    assert sum(len(b.members) for b in banks) == len(cross)
    for m in wrapped:  # re-homed, not changed
        assert torch.equal(m.base_layer.weight, before[id(m)])
    for b in banks:
        w = b.all.stacked_weight()
        assert w.shape == (b.all.G * b.all.N, b.all.K) and w.data_ptr() == b.all.layers[0].base_layer.weight.data_ptr()
        for _, g in b.members:  # the per-layer groups (backward) see their members inside the bank's matrix
            assert g.stacked_weight().data_ptr() == g.layers[0].base_layer.weight.data_ptr()
            assert g.stacked_weight().untyped_storage().data_ptr() == w.untyped_storage().data_ptr()
    assert [g for g in lora.projection_groups(unet) if g in [b.all for b in banks]]  # the optimizer refreshes their operands too
    # flat layout: same parameters as before, a bank = all its A (layer by layer: k, v adjacent), then all its B
    params = lora.lora_parameters(unet)
    assert sorted(id(p) for p in params) == sorted(plain_order) and len(set(id(p) for p in params)) == len(params)
    pos = {id(p): i for i, p in enumerate(params)}
    for b in banks:
        a_pos = [pos[id(l.lora_A["default"].weight)] for l in b.all.layers]
        b_pos = [pos[id(l.lora_B["default"].weight)] for l in b.all.layers]
        assert a_pos == list(range(a_pos[0], a_pos[0] + len(a_pos)))
        assert b_pos == list(range(a_pos[-1] + 1, a_pos[-1] + 1 + len(b_pos)))
    for m in attn:
        if not m.is_cross_attention:
            g = m.__dict__["_psob200_groups"]["qkv"]
            a_pos = [pos[id(l.lora_A["default"].weight)] for l in g.layers]
            b_pos = [pos[id(l.lora_B["default"].weight)] for l in g.layers]
            assert a_pos == list(range(a_pos[0], a_pos[0] + 3)) and b_pos == list(range(a_pos[0] + 3, a_pos[0] + 6))
    # the bucket built from this order hands every group / bank adjacent fp32 slices
    bucket = lora.LoRAGradBucket(params, align=8)
    for b in banks:
        for _, g in b.members:
            v = [l.lora_A["default"].weight._psob200_grad_view for l in g.layers]
            assert g._adjacent(v)


def test_padded_stacked_operands_refresh_in_one_copy(built_lib):
    """Ranks that are not multiples of 8 (BASELINE config 4: r = 4) are stacked through padded private copies.  Once the members'
    parameters are consecutive slices of the flat buffer, the refresh is ONE strided copy; it must give exactly what the
    member-by-member refresh gives, also after the parameters changed."""
    from pairwise_sample_optimization_b200 import lora
    torch.manual_seed(0)
    dtype = torch.bfloat16
    layers = []
    for _ in range(5):
        lay = lora.LoRALinear(torch.nn.Linear(32, 48, bias=False).to(dtype), 4, 4)
        with torch.no_grad():
            lay.lora_B["default"].weight.normal_(std=0.05)
        layers.append(lay)
    loose = lora.LoRAProjectionGroup(layers)                      # parameters allocated one by one: member-by-member copies
    a_slow, b_slow = loose.stacked_operand("a", dtype).clone(), loose.stacked_operand("b", dtype).clone()
    params = [l.lora_A["default"].weight for l in layers] + [l.lora_B["default"].weight for l in layers]
    flat = torch.cat([p.detach().flatten() for p in params])      # the flat layout of lora_parameters(): all A, then all B
    off = 0
    for p in params:
        p.data = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    packed = lora.LoRAProjectionGroup(layers)
    assert packed._adjacent([l.lora_A["default"].weight.detach() for l in layers])
    a_fast, b_fast = packed.stacked_operand("a", dtype), packed.stacked_operand("b", dtype)
    assert a_fast.shape == (5 * 8, 32) and torch.equal(a_fast, a_slow) and torch.equal(b_fast, b_slow)
    assert float(a_fast.view(5, 8, 32)[:, 4:].abs().max()) == 0.0  # the zero rows behind each projection's r rows stay zero
    with torch.no_grad():
        for p in params:
            p.add_(0.25)                                          # (in-place: bumps the version the refresh looks at)
    a2, b2 = packed.stacked_operand("a", dtype), packed.stacked_operand("b", dtype)
    for g, l in enumerate(layers):
        assert torch.equal(a2[g * 8:g * 8 + 4], l.lora_A["default"].weight.detach().to(dtype))
        assert torch.equal(b2[g * 48:(g + 1) * 48], l.lora_B["default"].weight.detach().to(dtype))
    assert a2.data_ptr() == a_fast.data_ptr()                     # refreshed in place (CUDA-graph friendly)
