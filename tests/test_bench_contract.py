"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys and never
loads the product library; our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_reference_arm_line_and_independence_from_the_product_library():
    # any attempt to load the product library would fail: the path does not exist
    env = dict(os.environ, WRFB200_LIB="/nonexistent/libwrfb200.so")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--steps", "3", "--warmup", "3"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "advance_mu_t grid-points/s" and d["unit"] == "grid-points/s"
    assert d["steps"] == 3 and d["warmup"] == 3 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    # the printed ms_per_step is the TIMED one: value x time per step = the points one small step updates
    from oracle import synth_np
    from wrf_model_cuda_sample_b200.advance_mu_t import Grid
    n3, _ = synth_np.updated_points(Grid.from_shape(74, 61, 28, halo=5))
    assert abs(d["value"] * d["ms_per_step"] * 1e-3 / n3 - 1.0) < 1e-6
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "grid-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(_have_gpu(), reason="GPU present")
def test_our_arm_refuses_to_run_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "tiny", "--steps", "1"],
                       capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")], "no JSON line may be printed"
