# -*- coding: utf-8 -*-
"""Helpers shared by the GPU parity tests: build our WaveNet from oracle-style parameter dicts."""
import numpy as np
import torch

from oracle import wavenet_oracle as O


def our_model(cfg, params, device="cuda", math_mode="fp32"):
    from pytorchwavenetvocoder_b200.nets import WaveNet
    net = WaveNet(*cfg.as_tuple())
    net.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v).astype(np.float32)) for k, v in params.items()})
    net.math_mode = math_mode
    return net.to(device)


def cfg_of(t):
    return O.Config(*t)
