# -*- coding: utf-8 -*-
"""CPU tier: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints ONE JSON line with the
contract's keys; it times the UNMODIFIED reference from git-ignored baseline/_ref when that travelled with the snapshot
(kind "reference"), the torch-CPU port otherwise (kind "port")."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"], cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900, env=dict(os.environ, PYTHONPATH=ROOT))
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = [ln for ln in r.stdout.decode().splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "wavenet_vocoder", "nets"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
