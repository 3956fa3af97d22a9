#!/usr/bin/env python
"""Timing of the batched MLSA noise-shaping kernel beside the CPU oracle (one host core, the C
restatement of SPTK's recursion): python tools/mlsa_timing.py [n_utts] [samples] -> one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mlsa_oracle as M  # noqa: E402  (checker / CPU baseline only)
from pytorchwavenetvocoder_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 160000
rng = np.random.RandomState(0)
coef = M.convert_mcep_to_mlsa_coef(rng.randn(25) * 0.6, 0.5, 0.41)
x = np.int16(np.clip(np.cumsum(rng.randn(B, n), axis=1) * 30, -30000, 30000))
dev = torch.device("cuda:0")
xd = torch.from_numpy(x.reshape(-1)).to(dev)
off = torch.arange(0, (B + 1) * n, n, dtype=torch.int64, device=dev)
b = torch.from_numpy(coef).to(dev)
y = torch.empty(B * n, dtype=torch.int16, device=dev)
lib = _lib.load()


def run():
    _lib.check(lib.wnb_mlsa_filter(_lib.ptr(xd), 1, _lib.ptr(off), B, _lib.ptr(b), 24, 0.41, 4, float(np.exp(coef[0])),
                                   _lib.ptr(y), 1, _lib.stream()), "mlsa_filter")


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
t0 = time.time()
ref = [M.noise_shaping_one(x[i], coef, 0.41) for i in range(2)]
cpu_s = (time.time() - t0) / 2
yh = y.cpu().numpy().reshape(B, n)
print(json.dumps({"workload": "%d utterances x %d samples, order 24, alpha 0.41, pd 4, int16 in/out" % (B, n),
                  "gpu_ms_per_batch": ms, "gpu_samples_per_s": B * n / (ms * 1e-3),
                  "cpu_oracle_s_per_utterance_1core": cpu_s, "cpu_samples_per_s_1core": n / cpu_s,
                  "bit_exact_vs_oracle_first_two": bool(np.array_equal(yh[0], ref[0]) and np.array_equal(yh[1], ref[1]))}))
