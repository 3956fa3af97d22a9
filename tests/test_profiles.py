# -*- coding: utf-8 -*-
"""CPU tier: the committed evidence under profiles/ is self-consistent -- the steady-state launch list summarised by
tools/launch_summary.py contains no torch compute kernels (only scalar helpers), its library launch count matches the
bench line's `gpu_launches`, and the bench line carries the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_list_and_bench_line_agree():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"),
                          os.path.join(ROOT, "profiles", "r2_launches_final.csv")], stdout=subprocess.PIPE, check=True).stdout.decode()
    rows = [ln for ln in out.splitlines() if " launches " in ln and ln.rstrip().endswith(("libwnb200", "torch"))]
    lib = sum(int(ln.split(" launches ")[0].split()[-1]) for ln in rows if ln.rstrip().endswith("libwnb200"))
    torch_us = sum(float(ln.split(" launches ")[1].split()[0]) for ln in rows if ln.rstrip().endswith("torch"))
    lib_us = sum(float(ln.split(" launches ")[1].split()[0]) for ln in rows if ln.rstrip().endswith("libwnb200"))
    assert torch_us < 0.005 * lib_us                       # torch launches on the step: scalar helpers only
    assert any("adam_flat_kernel" in ln for ln in rows) and not any("multi_tensor_apply" in ln for ln in rows)
    with open(os.path.join(ROOT, "profiles", "r2_bench_default.json")) as f:
        d = json.loads(f.read().strip().splitlines()[-1])
    assert d["gpu_launches"] == lib * d["steps"]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["e2e"]["h2d_bytes_per_step"] > 0
    assert abs(d["value"] - d["n_gpus"] * 8 * 20000 / (d["ms_per_step"] * 1e-3)) < 1e-3 * d["value"]
