"""bench.py contract checks that need no GPU: the reference arm (the CPU port of the reference algorithm) prints one
JSON line with the agreed keys, and the product arm of bench.py touches oracle/ only inside the CPU-baseline leg."""
import ast
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # stdout is reserved for the JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["scaling"] == "weak"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_does_not_import_the_oracle():
    """only cpu_reference_steps (the cpu_baseline / --impl reference leg) may import anything from oracle/"""
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    offenders = []
    for fn in [n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.Module))]:
        body = fn.body if isinstance(fn, ast.Module) else fn.body
        for node in body if isinstance(fn, ast.Module) else ast.walk(fn):
            if isinstance(node, (ast.Import, ast.ImportFrom)):
                names = [a.name for a in node.names] + ([node.module] if isinstance(node, ast.ImportFrom) and node.module else [])
                if any(n and ("mmdfn_oracle" in n or n == "helpers") for n in names):
                    where = "module" if isinstance(fn, ast.Module) else fn.name
                    if where != "cpu_reference_steps":
                        offenders.append((where, names))
    assert not offenders, offenders
    # and nothing under the product package imports it either
    pkg = os.path.join(ROOT, "mm-dfn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                assert "mmdfn_oracle" not in open(os.path.join(dirpath, f)).read(), os.path.join(dirpath, f)
