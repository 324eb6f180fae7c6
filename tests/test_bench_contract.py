"""bench.py prints ONE JSON line with the keys the driver reads (metric / value / unit / n_gpus / steps / warmup /
ms_per_step / higher_is_better / scaling / vs_baseline / dtype / data / config.workload / e2e / gpu_launches / clocks /
roofline / cpu_baseline).  The reference arm runs on the host CPU only (oracle restatement of ark-ec's MSM), so it is
checked in the CPU suite; the GPU arm in the GPU suite."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "e2e", "cpu_baseline")


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True,
                         timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def _check_common(d):
    for k in COMMON:
        assert k in d, k
    assert d["metric"] == "Pallas MSM Mpts/s @2^20" and d["unit"] == "Mpts/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600)
    _check_common(d)
    assert d["impl"] == "reference" and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--no-open"], 900)
    _check_common(d)
    assert d["steps"] == 3 and d["warmup"] == 3 and d["gpu_launches"] > 0 and d["verified_vs_oracle"] is True
    assert d["e2e"]["h2d_bytes_per_step"] == (1 << 20) * 32 and d["e2e"]["d2h_bytes_per_step"] > 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-3
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
