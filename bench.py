#!/usr/bin/env python
"""bench.py -- PSO train pairs/sec on synthetic SDXL-shaped data (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--pairs B]

Workload (config.workload): the configuration BASELINE.json's metric is quoted on ("SDXL 128x128 latents") = configs[2]'s
per-GPU slice: SDXL-DMD2 online PSO, LoRA rank 64, 1024x1024 (128x128 latents), bf16 (`--config dmd128`, default); configs[1]
(SDXL-Turbo, rank 8, 64x64 latents) is `--config turbo64`.  Both run on a random-init SDXL-ARCHITECTURE UNet
(fixtures/sdxl_unet.py: 2.57 B parameters, 70 transformer blocks, 560 LoRA-wrapped attention projections; diffusers is not
installed here).  A "step" is one training micro-step over B pairs, written as the reference writes it
(train_online_pso_sdxl_dmd2.py:773-864 / train_online_pso_sdxl_turbo.py:771-861):

    2 UNet forwards with grad (policy) + 2 adapter-disabled forwards without (frozen reference); activations stay resident
        (gradient checkpointing, the reference's memory workaround at :358, is `--grad-checkpointing`: same gradients, +28 % time)
    fused PSO loss+grad kernel  (replaces 4 x turbo_step_with_logprob + the inline loss + its backward)
    backward through the UNet   (LoRA dX / dA / dB on the tcgen05 GEMM path, accumulated into ONE flat fp32 buffer)
    every `accum` = GA x T = 6 steps: exchange of the flat LoRA gradient (N > 1: one multimem kernel of ours over the NVLink
    multicast mapping, fused with the gradient-norm pass; `--exchange nccl` = ncclAllReduce), then clip + AdamW +
    zero_grad + operand refresh fused over the flat buffers (2 launches)

The 560 projections (forward + backward), the loss and the feed-forward's gated GELU run on this repo's sm_100a kernels;
convolutions, norms, the feed-forward GEMMs and the attention core are stock torch kernels (outside the PSO hot path).  One rank per GPU, pairs sharded
over ranks (weak scaling); the only collective is the LoRA-gradient all-reduce.

Numbers on the JSON line:
  value / ms_per_step : exactly K steps, inputs resident in HBM, CUDA events, max over ranks.
  e2e                 : the same K steps with the step's inputs copied from pinned host memory and the loss read back,
                        every step, inside the timed region.
  roofline            : the dominant kernel of OURS in this step -- the tensor-bound main passes of the LoRA-wrapped
                        projections (lora_gemm2_kernel, CTA pairs) -- one CUDA-event pair around EVERY launch in a separate
                        instrumented pass; achieved = 2 M N (K + r) flops per launch / launch time; peak = sustained bf16 TF/s.
  lora_skinny_launches: the rank-r side launches (t, u, dA, dB: HBM / latency bound) of the same pass, as GB/s of their
                        algorithmic bytes against the measured HBM copy bandwidth.
  loss_kernel_roofline / lora_gemm_large : the two kernel-level figures of BASELINE config 5 (fused loss+grad kernel at 256
                        pairs x 128x128 latents against the HBM roofline; the fused base+LoRA GEMM at M = 18944).
  cpu_baseline        : the oracle port of the reference's PyTorch path (fp32, same UNet architecture) on this box's host
                        cores, one micro-step of ONE pair (bounded sample), N = 1 / rank 0 only.
`--impl reference` times that oracle port as the reference arm (the reference is Python that needs diffusers / peft /
accelerate and /root/reference; none exist on the GPU box -- DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pso_train_pairs_per_sec"
UNIT = "pairs/s"
ACCUM = 6  # gradient_accumulation_steps (2) x trained timesteps (3): turbo trainer :232, dmd2 trainer :236
_COMMON = ("online PSO micro-step (2 policy + 2 frozen-reference UNet forwards, fused PSO loss+grad, backward), random-init "
           "SDXL-architecture UNet (2.57 B params, 560 LoRA projections), bf16, optimizer step every 6 micro-steps")
# BASELINE.json's metric is quoted on SDXL 128x128 latents = configs[2] (SDXL-DMD2, rank 64, 1024 px): the default.
# configs[1] (SDXL-Turbo, rank 8, 64x64 latents, the reference's 512-px recipe) is `--config turbo64`.
CONFIGS = {
    "dmd128": {"kind": "dmd", "latent_hw": 128, "rank": 64,
               "workload": "BASELINE configs[2] per-GPU slice: SDXL-DMD2 " + _COMMON + ", LoRA rank 64, 128x128 latents (1024 px)"},
    "turbo64": {"kind": "turbo", "latent_hw": 64, "rank": 8,
                "workload": "BASELINE configs[1]: SDXL-Turbo " + _COMMON + ", LoRA rank 8, 64x64 latents (512 px)"},
}


# ----------------------------------------------------------------------------------------------- utilities
def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML, falls back to nvidia-smi)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _poll(self):
        while not self._stop.is_set():
            try:
                if self.nv is not None:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        self._thread = threading.Thread(target=self._poll, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join()

    def summary(self):
        if not self.samples:
            try:
                import subprocess
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
                sm, mx = (int(v) for v in out.stdout.strip().split(","))
                return {"sm_mhz": sm, "sm_max_mhz": mx, "reasons": sorted(self.reasons), "note": "sampled after the timed region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "tf_burst": float(p["bf16_tflops"]),
                "tf_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def dmd_scheduler():
    """SDXL alphas_cumprod (scaled_linear betas; restated, see oracle/schedules.py): all the DMD2 / LCM step reads."""
    import types
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    return types.SimpleNamespace(alphas_cumprod=torch.cumprod(1.0 - betas, dim=0))


def turbo_scheduler():
    """SDXL-Turbo EulerAncestral schedule, 4 trailing steps (restated; see oracle/schedules.py for the anchored copy)."""
    import types
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    ac = torch.cumprod(1.0 - betas, dim=0)
    ts = torch.tensor([999, 749, 499, 249])
    sig = ((1 - ac) / ac) ** 0.5
    return types.SimpleNamespace(timesteps=ts.float(), sigmas=torch.cat([sig[ts], torch.zeros(1)]).float())


def graph_timed(fn, reps, per_graph=2):
    """Mean device time (ms) of one ``fn()`` call: calls captured in a CUDA graph, one event pair per replay."""
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(per_graph):
            fn()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        graph.replay()
        b.record()
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in evs) / per_graph


# ----------------------------------------------------------------------------------------------- kernel-level figures
def loss_kernel_roofline(pso, dev, peaks, ncu_traffic):
    """Fused loss+grad kernel alone at the top of the config-5 sweep: 256 pairs of 4x128x128 bf16 latents (DMD2 shapes);
    inputs (336 MB) exceed L2.  Algorithmic bytes = 10 N sizeof(bf16) per pair."""
    import types
    B, n = 256, 4 * 128 * 128
    g = torch.Generator(device=dev).manual_seed(7)
    mk = lambda: torch.randn(B, 4, 128, 128, device=dev, generator=g).bfloat16()
    x0, x1, r0, r1, n0, n1 = mk(), mk(), mk(), mk(), mk(), mk()
    p0 = (r0.float() + 0.02 * torch.randn_like(r0, dtype=torch.float32)).bfloat16()
    p1 = (r1.float() + 0.02 * torch.randn_like(r1, dtype=torch.float32)).bfloat16()
    ts = torch.tensor([999, 749, 499], device=dev)[torch.randint(0, 3, (B,), device=dev)]
    h = torch.tensor([[-1.0, 1.0]], device=dev).repeat(B, 1)
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    sched = types.SimpleNamespace(alphas_cumprod=torch.cumprod(1.0 - betas, dim=0).to(dev))

    def call():
        with torch.no_grad():
            pso.pso_pair_loss(p0, p1, r0, r1, x0, x1, n0, n1, ts, ts, h, scheduler=sched, kind="dmd", step_ratio=250)
    ms = graph_timed(call, 20, per_graph=1)
    alg = 10 * n * 2 * B
    ach = alg / (ms * 1e-3) / 1e9
    return {"kernel": "pair_loss_grad_tmem_kernel<bf16,bf16,ref>", "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm"],
            "unit": "GB/s", "frac": round(ach / peaks["hbm"], 4), "traffic": ncu_traffic,
            "workload": "256 pairs x 4x128x128 bf16 (BASELINE config 5 top of sweep)", "algorithmic_bytes_per_launch": alg,
            "avg_launch_us": round(ms * 1e3, 2), "peak_source": peaks["source"]}


def lora_gemm_large(dev, peaks):
    """Fused base+LoRA projection y = x W^T + b + t B^T (one launch) at M = 18944 (148 row tiles), K = N = 1280, r = 64."""
    from pairwise_sample_optimization_b200 import gemm
    M, K, N, r = 18944, 1280, 1280, 64
    g = torch.Generator(device=dev).manual_seed(3)
    rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).bfloat16()
    x, w, b, A, Bm = rn(M, K), rn(N, K, sc=K ** -0.5), rn(N), rn(r, K, sc=1 / r), rn(N, r, sc=0.05)
    T, _ = gemm.lora_gemm(x, A)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ms = graph_timed(lambda: gemm.lora_gemm(x, w, T, Bm, bias=b, out=out), 20)
    fl = 2.0 * M * N * (K + r)
    ach = fl / (ms * 1e-3) / 1e12
    return {"kernel": "lora_gemm_kernel (x W^T + b + t B^T)", "bound": "tensor", "achieved": round(ach, 1),
            "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": round(ach / peaks["tf_burst"], 4),
            "shape": {"M": M, "K": K, "N": N, "r": r}, "avg_launch_us": round(ms * 1e3, 2), "peak_source": peaks["source"]}


# ----------------------------------------------------------------------------------------------- b200 arm
def run_b200(args):
    import pairwise_sample_optimization_b200 as pso
    from fixtures import micro_step, sdxl_unet
    from pairwise_sample_optimization_b200 import _lib, lora

    rank, world, local = dist_env()
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the b200 arm")
    L = _lib.lib()  # fail loudly if the CUDA library is missing
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.pairs, args.steps, max(args.warmup, 3)
    peaks = measured_peaks()
    conf = CONFIGS[args.config]
    KIND, LATENT_HW, RANK = conf["kind"], conf["latent_hw"], conf["rank"]

    # ---- model: random-init SDXL-architecture UNet in bf16, LoRA rank 8 on to_q/to_k/to_v/to_out.0
    torch.manual_seed(1234)  # same base weights on every rank (as a checkpoint would give)
    cfg = sdxl_unet.tiny_config() if args.tiny else sdxl_unet.sdxl_config()
    with torch.device(dev):
        unet = sdxl_unet.UNet2DConditionModel(cfg)
    unet = unet.to(torch.bfloat16).requires_grad_(False)
    wrapped = lora.add_adapter(unet, lora.LoraConfig(r=RANK, lora_alpha=RANK))
    for m in wrapped:  # the reference starts from B = 0; use a small non-zero B so every adapter GEMM does real work
        torch.nn.init.normal_(m.lora_B["default"].weight, std=0.01)
    unet.set_attn_processor(lora.PSOAttnProcessor2_0())
    if args.fused_geglu:  # the feed-forward's gated GELU on one fused kernel each way (outside SURVEY section 8's rows)
        from pairwise_sample_optimization_b200 import feed_forward
        feed_forward.install_fused_geglu(unet)
    lora.set_wgrad_stream(args.wgrad_stream)
    unet.train()
    if args.grad_checkpointing:
        unet.enable_gradient_checkpointing()  # turbo trainer :358
    # parameters, gradients and Adam moments of all 1120 adapter matrices live in four flat fp32 buffers: the optimizer
    # boundary is one all-reduce + two launches (clip + AdamW + zero_grad + 16-bit operand refresh)
    # data-parallel exchange: one kernel of ours per rank over the NVLink multicast mapping (in-switch reduction fused with the
    # norm pass of clip_grad_norm_); `--exchange nccl` (or a group without multicast support) = one NCCL all-reduce
    opt_kw = dict(lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_grad_norm=1.0)
    opt, exchange_kind = None, "none (1 GPU)"
    if world > 1 and args.exchange == "multimem":
        try:
            opt = lora.FusedLoRAOptimizer(unet, exchange=lora.SymmetricGradExchange(), **opt_kw)
            exchange_kind = "multimem.ld_reduce/st kernel fused with the gradient-norm pass (NVLink SHARP), no NCCL call"
        except Exception as e:  # no NVSwitch multicast on this box: same on every rank
            print(f"[bench] symmetric-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
    if opt is None:
        opt = lora.FusedLoRAOptimizer(unet, **opt_kw)
        if world > 1:
            exchange_kind = "one NCCL all-reduce of the flat gradient"
    bucket = opt.bucket
    sched = turbo_scheduler() if KIND == "turbo" else dmd_scheduler()
    pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
    host = micro_step.synth_batch(B, LATENT_HW, cfg.cross_attention_dim, pooled, 100 + rank, getattr(sched, "sigmas", None),
                                  dtype=torch.bfloat16, kind=KIND)
    if not args.separate_forwards:
        host = micro_step.batched_view(host)
    host = {k: v.pin_memory() for k, v in host.items()}
    d = {k: v.to(dev) for k, v in host.items()}
    fwd_bwd = micro_step.product_micro_step if args.separate_forwards else micro_step.product_micro_step_batched

    ref_stream = torch.cuda.Stream() if (args.overlap_reference and not args.separate_forwards) else None

    def micro(batch, overlap=True):
        kw = {"ref_stream": ref_stream} if (ref_stream is not None and overlap) else {}
        return fwd_bwd(pso, lora, unet, batch, sched, beta=50.0, eps=0.1, loss_scale=1.0 / ACCUM, kind=KIND, **kw)

    def optimizer_boundary(i):
        if (i + 1) % ACCUM == 0:  # turbo trainer :858-861 (sync_gradients): all-reduce, clip, AdamW, zero_grad
            opt.all_reduce()
            opt.step()

    graph = None
    static_loss = None
    launches_per_micro = 0

    def step(i, batch):
        if graph is not None:
            graph.replay()  # the captured micro-step reads the static device batch `d`
            loss = static_loss
        else:
            loss = micro(batch)
        optimizer_boundary(i)
        return loss

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    if not args.no_graph:
        # warm up eagerly on a side stream (also fills every host-side cache), then capture ONE micro-step: forward(s),
        # fused loss+grad kernel, backward with in-place accumulation into the flat bucket.  Replays need no Python.
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                micro(d)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        pso.check_status(dev)
        bucket.zero_()
        graph = torch.cuda.CUDAGraph()
        n0 = L.psob200_launch_count()
        with torch.cuda.graph(graph, stream=side):
            static_loss = micro(d)
        launches_per_micro = L.psob200_launch_count() - n0  # kernels of this library inside the captured micro-step
        bucket.zero_()
    for i in range(W):
        loss = step(i, d)
    optimizer_boundary(ACCUM - 1)  # one untimed optimizer boundary: AdamW state allocation, first NCCL all-reduce
    pso.check_status(dev)
    bucket.zero_()

    # ---- timed region: exactly K steps, inputs resident in HBM
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = L.psob200_launch_count()
    with ClockSampler(local) as clocks:
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        ev0.record()
        marks[0].record()
        for i in range(K):
            loss = step(i, d)
            marks[i + 1].record()
        ev1.record()
        barrier()
    launches = L.psob200_launch_count() - launches0 + K * launches_per_micro
    per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(K)]
    ms_per_step = max_over_ranks(ev0.elapsed_time(ev1)) / K
    value = world * B / (ms_per_step * 1e-3)
    loss_value = float(loss.item()) * ACCUM

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H loss, every step
    loss_pinned = torch.empty((), dtype=torch.float32).pin_memory()
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def e2e_step(i):
        for k, v in host.items():
            d[k].copy_(v, non_blocking=True)
        loss = step(i, d)
        loss_pinned.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_pinned)
    bucket.zero_()
    e2e_step(0)
    bucket.zero_()
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    barrier()
    e2e = {"value": round(world * B * K / (e2e_ms * 1e-3), 2), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
           "d2h_bytes_per_step": 4, "ms_per_step": round(e2e_ms / K, 3)}

    # ---- roofline of the dominant kernel of ours: instrumented pass, ONE launch between each pair of events
    peak_hbm_gb = round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2)  # before the second (instrumented) graph
    bucket.zero_()
    sink = []
    n_inst = 1
    lora.set_timing_sink(sink)
    if args.no_graph:
        micro(d, overlap=False)
    else:
        # captured like the timed step, with the event records as graph nodes: an eager pass is host-bound (the GPU idles
        # between a record and the launch behind it) and would charge host latency to the kernels
        side.wait_stream(torch.cuda.current_stream())
        inst_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(inst_graph, stream=side):  # the stream the step was warmed up and captured on
            micro(d, overlap=False)  # one stream: a launch's interval must not include waiting for the other stream's CTAs
        for _ in range(2):
            inst_graph.replay()  # the events keep the timestamps of the last replay
    lora.set_timing_sink(None)
    torch.cuda.synchronize()
    is_main = lambda role: role.startswith("y =") or role.startswith("dx =")
    agg = {True: [0.0, 0.0, 0.0, 0], False: [0.0, 0.0, 0.0, 0]}  # ms, flops, bytes, launches
    by_shape = {}
    for ev0, ev1, fl, by, role, shape in sink:
        ms = ev0.elapsed_time(ev1)
        a_ = agg[is_main(role)]
        a_[0] += ms; a_[1] += fl; a_[2] += by; a_[3] += 1
        t = by_shape.setdefault((role,) + tuple(shape), [0.0, 0, 0.0, 0.0])
        t[0] += ms; t[1] += 1; t[2] += fl; t[3] += by
    m_ms, m_fl, m_by, m_n = agg[True]
    s_ms, s_fl, s_by, s_n = agg[False]
    ach = m_fl / (m_ms * 1e-3) / 1e12 if m_ms > 0 else 0.0
    rows = [{"launch": k[0], "M": k[1], "K": k[2], "N": k[3], "r": k[4], "launches_per_step": v[1] // n_inst,
             "avg_us": round(v[0] * 1e3 / v[1], 2), "ms_per_step": round(v[0] / n_inst, 2),
             **({"tflops": round(v[2] / (v[0] * 1e-3) / 1e12, 1)} if is_main(k[0]) else
                {"gbs": round(v[3] / (v[0] * 1e-3) / 1e9, 1)})}
            for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1][0])]
    roofline = {"kernel": "lora_gemm2_kernel / lora_gemm_kernel main passes: y = x W^T + b + t B^T, dx = dy W + u A (tcgen05, "
                          "frozen weight + adapter in one pass) over the 560 LoRA-wrapped projections",
                "bound": "tensor", "achieved": round(ach, 2), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / peaks["tf_sustained"], 4), "traffic": None,
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "launches_per_step": m_n // n_inst, "avg_launch_us": round(m_ms * 1e3 / max(m_n, 1), 2),
                "share_of_step": round(m_ms / n_inst / ms_per_step, 4), "flops_per_step": m_fl / n_inst,
                "algorithmic_flops": "2 M N (K + r) per launch (r = 0 for the frozen-reference pass)",
                "by_launch": [r_ for r_ in rows if is_main(r_["launch"])],
                "how": "one CUDA-event pair around EVERY launch (psob200 forward_phases / backward_phases issue the launches of "
                       "a projection one at a time; event records captured as graph nodes) in one extra replayed step on ONE "
                       "stream (in the timed step the frozen-reference forward shares the SMs from a second stream, which would "
                       "charge its CTAs' residency to these intervals); includes the graph-node gaps around each launch and "
                       "lacks the programmatic-dependent-launch overlap, so slightly pessimistic"}
    s_ach = s_by / (s_ms * 1e-3) / 1e9 if s_ms > 0 else 0.0
    skinny = {"kernel": "lora_gemm_kernel skinny passes: t = s x A^T, u = s dy B, dA += u^T x, dB += dy^T t (rank-r side of "
                        "every projection; arithmetic intensity <= 84 flop/B, SURVEY.md section 8d)",
              "bound": "hbm", "achieved": round(s_ach, 1), "peak": peaks["hbm"], "unit": "GB/s",
              "frac": round(s_ach / peaks["hbm"], 4), "launches_per_step": s_n // n_inst,
              "avg_launch_us": round(s_ms * 1e3 / max(s_n, 1), 2), "share_of_step": round(s_ms / n_inst / ms_per_step, 4),
              "algorithmic_bytes": "operands read once + result written once per launch",
              "note": "latency-bound: one 128-row tile per CTA walks the whole reduction (13 us per launch at any M)",
              "by_launch": [r_ for r_ in rows if not is_main(r_["launch"])]}
    bucket.zero_()

    extra = {}
    if rank == 0 and not args.no_kernel_figures:
        extra["loss_kernel_roofline"] = loss_kernel_roofline(pso, dev, peaks, args.ncu_traffic)
        extra["lora_gemm_large"] = lora_gemm_large(dev, peaks)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del unet, opt, bucket
        torch.cuda.empty_cache()
        cpu = cpu_reference(args, reps=1, warmup=0)
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_per_step, 3), "ms_per_step_each": [round(v, 1) for v in per_step],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": conf["workload"] if not args.tiny else "TINY fixture (debug run, not the benchmark)",
                       "name": args.config,
                       "pairs_per_gpu_per_step": B, "latent_shape": [4, LATENT_HW, LATENT_HW], "lora_rank": RANK,
                       "beta": 50.0, "eps": 0.1, "accum": ACCUM,
                       "parallelism": f"dp{world} (pairs sharded; one exchange of the flat LoRA gradient per {ACCUM} steps)",
                       "gradient_exchange": exchange_kind,
                       "activations": ("recomputed in the backward (gradient checkpointing, as the reference)" if args.grad_checkpointing
                                       else "resident in HBM (no gradient checkpointing: same gradients, no recompute forward)"),
                       "weight_gradients": "dA / dB launches on a side stream" if args.wgrad_stream else "in stream order",
                       "feed_forward": "fused GEGLU kernels" if args.fused_geglu else "stock torch GEGLU",
                       "l2_policy": "working set larger than L2 (5.1 GB of bf16 weights streamed every forward)",
                       "forwards": "4 separate (as the reference)" if args.separate_forwards else
                                   "win+lose batched: 1 policy + 1 reference forward of batch 2B" +
                                   (", reference forward on a second stream" if ref_stream is not None else ""),
                       "timing": ("eager launches" if args.no_graph else "micro-step replayed from one CUDA graph; optimizer "
                                  "boundary eager") + ", CUDA events around K steps, max over ranks"},
            "gpu_launches": int(launches), "clocks": clocks.summary(), "e2e": e2e, "roofline": roofline,
            "lora_skinny_launches": skinny,
            "loss": round(loss_value, 6), "peak_hbm_gb": peak_hbm_gb,
        }
        line.update(extra)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- reference arm (CPU)
def _make_cpu_step(args, pairs):
    """The reference's micro-step restated with the oracle pieces (oracle/ is checker code; it is executed here only as
    the reported CPU baseline / reference arm): fp32, same UNet architecture, oracle LoRA modules (peft forward restated),
    four step-with-logprob calls, the inline loss, autograd backward; all host threads."""
    from fixtures import micro_step, sdxl_unet
    from oracle import lora as olora, losses as olosses, schedules
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = sdxl_unet.tiny_config() if args.tiny else sdxl_unet.sdxl_config()
    with torch.device("meta"):
        unet = sdxl_unet.UNet2DConditionModel(cfg)
    unet = unet.to_empty(device="cpu")
    g = torch.Generator().manual_seed(0)
    pool = torch.randn(1 << 22, generator=g) * 0.02  # cheap random init: values only need to be finite and small
    with torch.no_grad():
        for prm in unet.parameters():
            n = prm.numel()
            flat = prm.view(-1)
            for off in range(0, n, pool.numel()):
                m = min(pool.numel(), n - off)
                flat[off:off + m] = pool[:m]
            if prm.dim() == 1:
                prm.zero_()
        for mod in unet.modules():
            if isinstance(mod, (torch.nn.GroupNorm, torch.nn.LayerNorm)):
                mod.weight.fill_(1.0)
                mod.bias.zero_()
    unet.requires_grad_(False)
    conf = CONFIGS[args.config]
    KIND, LATENT_HW, RANK = conf["kind"], conf["latent_hw"], conf["rank"]
    wrapped = olora.oracle_add_adapter(unet, RANK, RANK)
    for m in wrapped:
        torch.nn.init.normal_(m.lora_B["default"].weight, std=0.01)
    unet.train()
    unet.enable_gradient_checkpointing()
    sched = schedules.turbo_scheduler(4) if KIND == "turbo" else schedules.dmd_scheduler()
    pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
    batch = micro_step.synth_batch(pairs, LATENT_HW, cfg.cross_attention_dim, pooled, 100, getattr(sched, "sigmas", None),
                                   kind=KIND)

    def one():
        loss = micro_step.oracle_micro_step(olora, olosses, unet, batch, sched, beta=50.0, eps=0.1, loss_scale=1.0 / ACCUM,
                                            kind=KIND)
        for m in wrapped:
            m.lora_A["default"].weight.grad = None
            m.lora_B["default"].weight.grad = None
        return float(loss.detach()) * ACCUM
    return one, threads


def cpu_reference(args, reps, warmup):
    pairs = args.cpu_pairs
    one, threads = _make_cpu_step(args, pairs)
    for _ in range(warmup):
        one()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": round(pairs / med, 4), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{pairs} pair(s) x {len(times)} rep(s) of the same micro-step (fp32, torch {torch.__version__} CPU, "
                      f"{torch.get_num_threads()} threads, no warm-up), {med:.1f} s per step",
            "ms_per_step": round(med * 1e3, 1)}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    K, W = args.steps, min(args.warmup, 1)
    pairs = args.cpu_pairs
    one, threads = _make_cpu_step(args, pairs)
    t_begin = time.perf_counter()
    res = None
    for _ in range(W):
        res = one()
    times = []
    for _ in range(K):  # each step: one micro-step over `pairs` pair(s); bounded by --ref-budget-seconds
        t0 = time.perf_counter()
        res = one()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > args.ref_budget_seconds:
            break
    ms = statistics.mean(times) * 1e3
    value = pairs / (ms * 1e-3)
    sample = (f"{pairs} pair(s) per step (bounded sample of the {args.pairs}-pair micro-step), fp32, torch {torch.__version__} CPU, "
              f"{torch.get_num_threads()} threads; {len(times)} of the {K} requested steps fit the {args.ref_budget_seconds:.0f} s budget")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "steps_requested": K, "warmup": W, "ms_per_step": round(ms, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS[args.config]["workload"] + " -- oracle port of the reference's PyTorch path on the "
                                   "host cores", "name": args.config,
                       "pairs_per_step": pairs,
                       "latent_shape": [4, CONFIGS[args.config]["latent_hw"], CONFIGS[args.config]["latent_hw"]],
                       "lora_rank": CONFIGS[args.config]["rank"], "beta": 50.0, "eps": 0.1},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "loss": round(res, 6)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="dmd128", choices=sorted(CONFIGS),
                    help="dmd128 = the configuration BASELINE.json's metric is quoted on (SDXL 128x128 latents); turbo64 = configs[1]")
    ap.add_argument("--pairs", type=int, default=4, help="pairs per GPU per micro-step (train.batch_size of the shipped recipe)")
    ap.add_argument("--cpu-pairs", type=int, default=1, help="pairs in the bounded CPU sample")
    ap.add_argument("--ref-budget-seconds", type=float, default=150.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-figures", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the micro-step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--separate-forwards", action="store_true", help="4 UNet forwards of batch B instead of 2 of batch 2B")
    ap.add_argument("--overlap-reference", action="store_true", default=True,
                    help="issue the frozen-reference forward on a second stream (default)")
    ap.add_argument("--no-overlap-reference", dest="overlap_reference", action="store_false")
    ap.add_argument("--grad-checkpointing", dest="grad_checkpointing", action="store_true", default=False,
                    help="recompute the blocks in the backward (the reference's memory workaround, turbo trainer :358); default: "
                         "the activations of the 4-pair micro-step (43 GB at 128x128 latents) stay resident in the 180 GB of HBM")
    ap.add_argument("--no-grad-checkpointing", dest="grad_checkpointing", action="store_false")
    ap.add_argument("--wgrad-stream", dest="wgrad_stream", action="store_true", default=True,
                    help="issue the dA / dB launches of every projection on a side stream (joined at the end of the backward)")
    ap.add_argument("--no-wgrad-stream", dest="wgrad_stream", action="store_false")
    ap.add_argument("--exchange", default="multimem", choices=["multimem", "nccl"],
                    help="N > 1: the LoRA-gradient exchange (multimem = this repo's NVLink-multicast kernel; nccl = ncclAllReduce)")
    ap.add_argument("--no-fused-geglu", dest="fused_geglu", action="store_false", default=True,
                    help="leave the feed-forward's GEGLU on the stock torch kernels")
    ap.add_argument("--tiny", action="store_true", help="debug: the 32/64-channel fixture instead of the SDXL architecture")
    ap.add_argument("--ncu-traffic", type=float, default=311483648.0,
                    help="dram__bytes_read.sum + dram__bytes_write.sum per launch of the loss kernel, from the committed "
                         "ncu --set full capture profiles/r01_pair_loss_tmem_ncu_raw.txt (268.56 MB + 42.92 MB)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
