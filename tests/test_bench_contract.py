"""bench.py contract checks that need no GPU: the reference (CPU) arm prints one valid JSON line with the keys
the driver reads, the B200 arm refuses to run without a device (no CPU fallback), and oracle/ is only touched
by the CPU legs."""
import json
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec" and d["unit"] == "img/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "resnet50" in d["config"]["workload"]


def test_b200_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "1", timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_oracle_is_only_used_by_the_cpu_legs_of_bench():
    src = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"^\s*from oracle\b|^\s*import oracle\b", src, flags=re.M)]
    assert len(uses) == 1
    # the single import sits inside oracle_forward(), which only cpu_port() and run_reference_arm() call
    fn_start = src.rfind("\ndef ", 0, uses[0])
    assert src[fn_start:].lstrip().startswith("def oracle_forward(")
    callers = re.findall(r"^def (\w+)\(.*?(?=^def |\Z)", src, flags=re.M | re.S)
    for name in ("measure", "main", "build_model", "synthetic_state_dict"):
        body = re.search(r"^def %s\(.*?(?=^def |\Z)" % name, src, flags=re.M | re.S).group(0)
        assert "oracle_forward" not in body or name == "main", name
    assert "oracle_forward" in re.search(r"^def cpu_port\(.*?(?=^def |\Z)", src, flags=re.M | re.S).group(0)
