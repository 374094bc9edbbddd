"""Repository contract checks that need no GPU: the product never touches the oracle, the
reference arm of bench.py prints the agreed JSON line, the docs exist."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may use oracle/."""
    pkg = os.path.join(ROOT, "dynamicscaler_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
                assert "/root/reference" not in text, os.path.join(dirpath, f)
    bench = open(os.path.join(ROOT, "bench.py"), encoding="utf-8").read()
    uses = [m.start() for m in re.finditer(r"from oracle import", bench)]
    assert len(uses) == 1                                   # inside cpu_one_step (cpu_baseline / --impl reference legs) only
    assert bench.rfind("def cpu_one_step", 0, uses[0]) > max(bench.rfind(f"def {n}", 0, uses[0])
                                                               for n in ("run_ours", "parity_check", "make_runner"))
    body = bench[uses[0]:bench.find("\ndef ", uses[0])]
    assert "tp." in body and "gather_all" not in body      # the oracle is the thing timed there, never the GPU path
    assert "/root/reference" not in bench


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--mode", "nearest",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "blends/s" and line["higher_is_better"] is True
    assert line["config"]["workload"] == "cfg2_sphere_128x256" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "blends/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_docs_and_profiles_present():
    for f in ("DESIGN.md", "INTEGRATION.md", "include/dynscaler.h", "profiles/traffic.json"):
        assert os.path.exists(os.path.join(ROOT, f)), f
    assert any(n.endswith("_launches.csv") for n in os.listdir(os.path.join(ROOT, "profiles")))
