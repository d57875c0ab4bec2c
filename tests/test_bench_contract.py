"""bench.py prints ONE JSON line with the driver's contract keys (native arm, tiny configuration)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_bench_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "3", "--warmup", "3",
                          "--batch", "8", "--model", "vggf-mini"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    j = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks"):
        assert k in j, k
    assert j["n_gpus"] == 1 and j["steps"] == 3 and j["higher_is_better"] is True and j["scaling"] == "weak"
    assert j["value"] > 0 and j["gpu_launches"] > 0 and j["dtype"] == "bf16"
    assert j["e2e"]["value"] > 0 and j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] == 4
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(j["clocks"])


def test_bench_reference_arm_reports_unavailable_without_install(tmp_path):
    """--impl reference must exit 0 with an 'unavailable' line when baseline/_ref is missing."""
    import shutil

    work = tmp_path / "repo"
    work.mkdir()
    shutil.copy(os.path.join(ROOT, "bench.py"), work / "bench.py")
    out = subprocess.run([sys.executable, str(work / "bench.py"), "--impl", "reference"], cwd=str(work),
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0
    j = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert j["impl"] == "reference" and "unavailable" in j
