#!/usr/bin/env python
"""Tuning aid: one training step of the bench workload with the per-role wait profilers of the tcgen05 kernels
switched on (WNB_PROF=1 / WNB_FWD_PROF=1: every launch is synchronous and prints one line to stderr).
Usage: python tools/step_prof.py 2>&1 | grep 'wnb200' | sed -n '200,260p'"""
import os
import sys

os.environ["WNB_PROF"] = "1"
os.environ["WNB_FWD_PROF"] = os.environ.get("WNB_FWD_PROF", "0")
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pytorchwavenetvocoder_b200.nets import cross_entropy  # noqa: E402

dev = torch.device("cuda:0")
cfg, net = bench.make_model(dev, "tf32")
net.train()
xh, hh, th = bench.synth_batch(cfg, 0, bench.BATCH, pinned=False)
x, h, t = xh.to(dev), hh.to(dev), th.to(dev)
for _ in range(2):
    y = net(x, h)
    loss = cross_entropy(y, t, cfg.receptive_field)
    net.zero_grad(set_to_none=True)
    sys.stderr.write("wnb200 ---- backward ----\n")
    sys.stderr.flush()
    loss.backward()
    torch.cuda.synchronize()
    sys.stderr.write("wnb200 ---- step done ----\n")
