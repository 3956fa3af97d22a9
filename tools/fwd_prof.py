#!/usr/bin/env python
"""Tuning aid: run a few fused residual-block launches with WNB_FWD_PROF=1 (see resblock_tc.cu) and print where
each warp role of the pipeline waited.  Usage: python tools/fwd_prof.py [B T]"""
import os
import sys

os.environ["WNB_FWD_PROF"] = "1"
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorchwavenetvocoder_b200 import _lib  # noqa: E402

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 23040)
R, S, Ap, ks = 64, 512, 32, 2
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, T, R, device=dev, generator=g)
h = torch.randn(B, T, Ap, device=dev, generator=g)
w1 = torch.randn(2 * R, ks * R + Ap, device=dev, generator=g) * 0.05
b1 = torch.zeros(2 * R, device=dev)
w2 = torch.randn(R + S, R, device=dev, generator=g) * 0.05
b2 = torch.zeros(R + S, device=dev)
xo = torch.empty_like(x)
skip = torch.zeros(B, T, S, device=dev)
lib = _lib.load()
for d, init in ((1, 1), (1, 0), (64, 0), (512, 0)):
    for _ in range(2):
        _lib.check(lib.wnb_resblock_fwd(_lib.ptr(x), _lib.ptr(h), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2),
                                        _lib.ptr(xo), _lib.ptr(skip), None, B, T, R, S, Ap, ks, d, init, 1,
                                        _lib.stream()), "resblock_fwd")
torch.cuda.synchronize()
