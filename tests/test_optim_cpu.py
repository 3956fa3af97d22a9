# -*- coding: utf-8 -*-
"""CPU tier: optim.Adam without the flat layout (CPU parameters, no module) IS torch.optim.Adam -- same updates, same
state_dict keys; the one-kernel path is the GPU tier's (tests/test_gpu_optim.py)."""
import torch


def test_fallback_is_torch_adam():
    from pytorchwavenetvocoder_b200.optim import Adam
    torch.manual_seed(0)
    a = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa, ob = Adam(a, lr=1e-2, weight_decay=0.1), torch.optim.Adam(b, lr=1e-2, weight_decay=0.1)
    for it in range(4):
        for p, q in zip(a, b):
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    assert oa.flat_steps == 0 and all(torch.equal(p, q) for p, q in zip(a, b))
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["state"].keys() == sb["state"].keys() and set(sa["state"][0]) == set(sb["state"][0])
    oc = Adam([torch.nn.Parameter(p.detach().clone()) for p in a], lr=1e-2, weight_decay=0.1)
    oc.load_state_dict(sb)
    assert float(oc.state_dict()["state"][0]["step"]) == 4.0
