#!/usr/bin/env python
"""Two training steps of the bench workload, nothing else: the target of `ncu -k regex:... --launch-skip ...` captures.
    ncu --set full --import-source on --clock-control none -k regex:gemm_nt_tc_kernel --launch-skip 100 --launch-count 4 \
        -o gpurun_out/r2_prof_blk python tools/prof_step.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda:0")
cfg, net = bench.make_model(dev, "tf32")
net.train()
xh, hh, th = bench.synth_batch(cfg, 0, bench.BATCH, pinned=False)
x, h, t = xh.to(dev), hh.to(dev), th.to(dev)
from pytorchwavenetvocoder_b200.optim import Adam  # noqa: E402
opt = Adam(net.parameters(), lr=1e-4, module=net)
for _ in range(int(os.environ.get("STEPS", "2"))):
    loss = net.forward_loss(x, h, t, cfg.receptive_field)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
print("loss", float(loss))
