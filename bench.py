#!/usr/bin/env python
"""bench.py -- PSO train pairs/sec on synthetic SDXL-shaped latents (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--pairs B]

Workload (config.workload): the online-PSO hot path of one training micro-step on DMD2-shaped data
(4x128x128 latents, bf16), B pairs per GPU:
    2 sampler-update launches  (distilled_step_with_logprob, sampling mode: next latents of both branches)
    1 fused loss+grad launch   (pso_pair_loss: four log-probs + pairwise log-sigmoid loss + grad into the
                                two policy predictions)
    loss.backward()            (hands the fused gradients to autograd; 2 device-side no-op scale launches)
One rank per GPU; pairs shard across ranks with no data-path collective (weak scaling).

Numbers:
  value / ms_per_step : K CUDA-graph replays of the step, inputs resident in HBM, CUDA events, max over ranks.
  roofline            : the fused loss+grad kernel, one CUDA-event pair around every launch in a second pass of
                        K direct launches on the same (larger-than-L2) inputs; algorithmic bytes = 10*N*2 per pair.
  e2e                 : the same step through the public API from pinned host buffers: H2D of the step's inputs,
                        the step, D2H of the loss, every step, inside the timed region.
  cpu_baseline        : the oracle port of the reference's PyTorch path (oracle/), timed on this box's host cores
                        on a bounded sample (N=1, rank 0 only).
`--impl reference` times that oracle port as the reference arm (the reference is pure Python and cannot be
installed on the GPU box: diffusers/peft/accelerate are absent).
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
SHAPE = (4, 128, 128)
N_ELEM = 4 * 128 * 128
DMD_TS = [999, 749, 499]
STEP_RATIO = 250


# ----------------------------------------------------------------------------------------------- utilities
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


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
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------- synthetic data
def make_host_inputs(B: int, seed: int, pin: bool):
    """Synthetic step inputs on the host: x ~ N(0,1), eps_ref ~ N(0,1),
    eps_policy = eps_ref + 0.02 N(0,1) (keeps exp(delta) inside the clamp), timesteps uniform over the trained
    ones, preferences uniform with 5% ties."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in (0, 1):
        out[f"x{k}"] = torch.randn(B, *SHAPE, generator=g).bfloat16()
        ref = torch.randn(B, *SHAPE, generator=g)
        out[f"ref{k}"] = ref.bfloat16()
        out[f"pred{k}"] = (ref + 0.02 * torch.randn(B, *SHAPE, generator=g)).bfloat16()
    tsel = torch.randint(0, len(DMD_TS), (B,), generator=g)
    out["ts"] = torch.tensor(DMD_TS)[tsel]
    sign = torch.randint(0, 2, (B,), generator=g).float() * 2 - 1
    h = torch.stack([-sign, sign], 1)
    h[torch.rand(B, generator=g) < 0.05] = 0.0
    out["h"] = h
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def alphas_cumprod() -> torch.Tensor:
    """SDXL scaled_linear schedule (diffusers scheduler config; restated, see oracle/schedules.py)."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


# ----------------------------------------------------------------------------------------------- b200 arm
def run_b200(args):
    import types

    import pairwise_sample_optimization_b200 as pso
    from pairwise_sample_optimization_b200 import _lib

    rank, world, local = dist_env()
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the b200 arm")
    _lib.lib()  # fail loudly if the CUDA library is missing
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.pairs, args.steps, args.warmup
    host = make_host_inputs(B, 1234 + rank, pin=True)
    d = {k: v.to(dev) for k, v in host.items()}
    sched = types.SimpleNamespace(alphas_cumprod=alphas_cumprod().to(dev))
    torch.cuda.manual_seed(99 + rank)  # sampler noise comes from the default CUDA generator (graph-capturable)
    launches_per_step = 5

    def step(t):
        """One pass of the hot path over B pairs through the public API."""
        ts, tp = t["ts"], t["ts"] - STEP_RATIO
        nxt = []
        for k in (0, 1):  # sampler update under the frozen reference policy -> stored next latents
            xn, _ = pso.distilled_step_with_logprob(sched, t[f"ref{k}"], ts, tp, t[f"x{k}"])
            nxt.append(xn)
        # fresh autograd leaves every step (what a UNet forward would hand over)
        p0, p1 = t["pred0"].detach().requires_grad_(True), t["pred1"].detach().requires_grad_(True)
        loss = pso.pso_pair_loss(p0, p1, t["ref0"], t["ref1"], t["x0"], t["x1"], nxt[0], nxt[1],
                                 ts, ts, t["h"], scheduler=sched, kind="dmd", beta=50.0, eps=0.1, step_ratio=STEP_RATIO)
        loss.backward()
        return loss, nxt, (p0.grad, p1.grad)

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

    # ---- warm-up (eager), then capture one step in a CUDA graph on a side stream
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        for _ in range(max(W, 3)):
            loss, nxt, grads = step(d)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        g_loss, g_nxt, g_grads = step(d)
    for _ in range(max(W, 3)):
        graph.replay()
    pso.check_status(dev)

    # ---- timed region: exactly K steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(K):
            graph.replay()
        ev1.record()
        barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / K
    value = world * B / (ms_per_step * 1e-3)
    loss_value = float(g_loss.item())

    # ---- roofline pass: the fused loss+grad kernel alone, one event pair per launch
    nxt = [t.detach() for t in g_nxt]
    ts = d["ts"]
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]

    def loss_only(tune=(0, 0)):
        with torch.no_grad():
            return pso.pso_pair_loss(d["pred0"].detach(), d["pred1"].detach(), d["ref0"], d["ref1"], d["x0"], d["x1"],
                                     nxt[0], nxt[1], ts, ts, d["h"], scheduler=sched, kind="dmd", beta=50.0, eps=0.1,
                                     step_ratio=STEP_RATIO, tune=tune)
    # the Python/ctypes cost of one call exceeds the kernel's run time, so the launch is captured in a CUDA graph:
    # the host then runs ahead of the GPU and each event pair brackets device time only
    with torch.cuda.stream(side):
        for _ in range(3):
            loss_only()
    torch.cuda.synchronize()
    kgraph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(kgraph, stream=side):
        loss_only()
    for _ in range(3):
        kgraph.replay()
    torch.cuda.synchronize()
    for a, b in evs:
        a.record()
        kgraph.replay()
        b.record()
    torch.cuda.synchronize()
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
    alg_bytes = 10 * N_ELEM * 2 * B
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"kernel": "pair_loss_grad_kernel<bf16,bf16,ref,vec8>", "bound": "hbm", "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": args.ncu_traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "avg_launch_us": round(kern_ms * 1e3, 2)}

    # ---- end-to-end: pinned host buffers -> H2D -> step -> D2H loss, every step
    loss_pinned = torch.empty((), dtype=torch.float32).pin_memory()
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def e2e_step():
        with torch.no_grad():
            for k, v in host.items():
                d[k].copy_(v, non_blocking=True)
        loss, _, _ = step(d)
        loss_pinned.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_pinned)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    barrier()
    e2e = {"value": round(world * B * K / (e2e_ms * 1e-3), 1), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
           "d2h_bytes_per_step": 4, "ms_per_step": round(e2e_ms / K, 4)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(args.cpu_pairs, budget_s=args.cpu_seconds)
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 storage / fp32 math", "data": "synthetic",
            "config": {"workload": "online-PSO loss hot path, SDXL-DMD2 shapes (BASELINE configs[2] per-GPU slice): "
                                   "2 sampler-update launches + 1 fused loss+grad launch + backward hand-off; "
                                   "LoRA GEMMs not in this step",
                       "pairs_per_gpu": B, "latent_shape": list(SHAPE), "beta": 50.0, "eps": 0.1,
                       "trained_timesteps": DMD_TS, "parallelism": f"dp{world} (pairs sharded, no collective)",
                       "l2_policy": f"inputs larger than L2 ({alg_bytes / 1e6:.0f} MB touched per step vs 126 MB L2)",
                       "timing": "CUDA-graph replay of the step; CUDA events; max over ranks"},
            "gpu_launches": launches_per_step * K, "clocks": clocks.summary(), "e2e": e2e, "roofline": roofline,
            "loss": round(loss_value, 6),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- reference arm
def _make_cpu_step(pairs: int):
    """One micro-step of the reference's PyTorch path as restated in oracle/ (checker code, executed here only as
    the reported CPU baseline / reference arm): two sampling-mode step calls, four scoring-mode step calls, the
    inline loss and loss.backward(); fp32, all host threads."""
    from oracle import losses, schedules, steps
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sched = schedules.dmd_scheduler()
    h = make_host_inputs(pairs, 4321, pin=False)
    x = [h["x0"].float(), h["x1"].float()]
    ref = [h["ref0"].float(), h["ref1"].float()]
    ts = h["ts"]
    gen = torch.Generator().manual_seed(5)

    def one():
        pred = [h["pred0"].float().requires_grad_(True), h["pred1"].float().requires_grad_(True)]
        nxt = []
        with torch.no_grad():
            for k in (0, 1):
                xn, _ = steps.distilled_step(sched, ref[k], ts, ts - STEP_RATIO, x[k], generator=gen)
                nxt.append(xn)
        loss, _ = losses.online_micro_step("dmd", sched, pred, ref, x, nxt, [ts, ts], h["h"], 50.0, 0.1,
                                           step_ratio=STEP_RATIO)
        loss.backward()
        return float(loss.detach())
    return one, threads


def cpu_reference(pairs: int, budget_s: float = 15.0, min_reps: int = 3):
    """The reference's path on the host cores (the reference itself is Python that needs diffusers/peft/accelerate
    and /root/reference, neither of which exists on the GPU box, so the oracle port is what runs)."""
    one, threads = _make_cpu_step(pairs)
    one()
    times = []
    t_start = time.perf_counter()
    while len(times) < min_reps or (time.perf_counter() - t_start) < budget_s:
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
        if len(times) >= 200:
            break
    med = statistics.median(times)
    return {"value": round(pairs / med, 1), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{pairs} pairs x {len(times)} reps of the same micro-step (fp32, torch {torch.__version__} CPU, "
                      f"{torch.get_num_threads()} threads), median {med * 1e3:.1f} ms",
            "ms_per_step": round(med * 1e3, 3)}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    pairs = args.cpu_pairs
    one, threads = _make_cpu_step(pairs)
    res = None
    for _ in range(max(W, 1)):  # each "step" is one bounded micro-step of `pairs` pairs
        one()
    t0 = time.perf_counter()
    for _ in range(K):
        res = one()
    dt = time.perf_counter() - t0
    ms = dt / K * 1e3
    value = pairs / (ms * 1e-3)
    sample = (f"{pairs} pairs per step (bounded sample of the {args.pairs}-pair workload), fp32, torch "
              f"{torch.__version__} CPU, {torch.get_num_threads()} threads")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": max(W, 1), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "online-PSO loss hot path, SDXL-DMD2 shapes: oracle port of the reference's "
                                   "PyTorch path on host cores (2 sampling steps, 4 scoring steps, inline loss, backward)",
                       "pairs_per_step": pairs, "latent_shape": list(SHAPE), "beta": 50.0, "eps": 0.1},
            "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "loss": round(res, 6)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=256, help="pairs per GPU per step")
    ap.add_argument("--cpu-pairs", type=int, default=16, help="pairs in the bounded CPU sample")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-traffic", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from the committed ncu --set full capture")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
