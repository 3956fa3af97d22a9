# -*- coding: utf-8 -*-
"""CPU tier: oracle/torch_port.py (the timed CPU baseline) reproduces the golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import torch_port as TP
from oracle import wavenet_oracle as O
from tests.golden.cases import FORWARD_CASES, GEN_CASES, make_gen_inputs, make_inputs

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["tiny_up", "ks3_up", "d10x3"])
def test_forward(name):
    cfg_t, seed, B, T, start = FORWARD_CASES[name]
    cfg = O.Config(*cfg_t)
    g = np.load(os.path.join(G, "forward_%s.npz" % name))
    p = TP.params_to_torch(O.make_params(cfg, seed))
    x, h, t = make_inputs(cfg, seed, B, T)
    with torch.no_grad():
        y = TP.forward(cfg, p, torch.from_numpy(x), torch.from_numpy(h)).numpy()
    np.testing.assert_allclose(y, g["logits"], atol=1e-5, rtol=0)


@pytest.mark.parametrize("name", ["ks2", "ks3", "ks2_up"])
def test_generate(name):
    cfg_t, seed, B, T0, n_list, naive = GEN_CASES[name]
    cfg = O.Config(*cfg_t)
    g = np.load(os.path.join(G, "gen_%s.npz" % name))
    p = TP.params_to_torch(O.make_params(cfg, seed))
    x, h = make_gen_inputs(cfg, seed, B, T0, n_list)
    outs = TP.batch_fast_generate(cfg, p, torch.from_numpy(x), torch.from_numpy(h), list(n_list), "argmax")
    for i, o in enumerate(outs):
        assert np.array_equal(o, g["batch_%d" % i])
