"""The reference arm of bench.py runs without a GPU (it times the compiled reference on the host cores): check that it prints ONE JSON
line with the keys the driver reads, on the same config as the GPU arm."""
import json
import os
import subprocess
import sys

import pytest

from common import ROOT, ref_lib


@pytest.mark.skipif(ref_lib() is None, reason="prebuilt compiled reference not available")
def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"].startswith("configs[1]") and len(d["config"]["codecs"]) == 19 and d.get("same_config") is True
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == 1 and cb["value"] == d["value"] and "cpu_model" in cb and "isa_path" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["all_cores"]["cores"] >= 1 and d["all_cores"]["value"] > 0
