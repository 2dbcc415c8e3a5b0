"""bench.py's b200 arm on the tiny fixture (debug run, seconds): the JSON line carries the keys of the bench contract, the step
runs from the captured graph, and kernels of this library were launched inside the timed region."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_b200_arm_json_line_on_the_tiny_fixture(built_lib):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--tiny", "--config", "turbo64", "--steps", "6", "--warmup", "3",
                          "--no-cpu-baseline", "--no-kernel-figures"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "gpu_launches", "clocks", "e2e", "roofline", "lora_skinny_launches", "peak_hbm_gb"):
        assert key in line, key
    assert line["metric"] == "pso_train_pairs_per_sec" and line["unit"] == "pairs/s" and line["n_gpus"] == 1
    assert line["steps"] == 6 and line["warmup"] >= 3 and line["value"] > 0 and line["scaling"] == "weak"
    assert line["gpu_launches"] > 6 * 100  # 96 LoRA projections, forward + backward, every step
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] == 4
    for r in (line["roofline"], line["lora_skinny_launches"]):
        assert r["bound"] in ("tensor", "hbm") and r["achieved"] > 0 and r["peak"] > 0
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["launches_per_step"] > 0
    assert "workload" in line["config"] and line["clocks"]["sm_max_mhz"]
