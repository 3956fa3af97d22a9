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


def test_flat_layout_detection_on_host_tensors():
    """optim.Adam._offsets: gradients that tile one flat buffer are recognised (any order, parameters without a gradient
    skipped); a gap, a foreign tensor or a wrong dtype is not."""
    from pytorchwavenetvocoder_b200.optim import Adam
    ps = [torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2)),
          torch.nn.Parameter(torch.zeros(3))]
    opt = Adam(ps, lr=1e-3)
    flat = torch.arange(21, dtype=torch.float32)
    ps[1].grad = flat[0:5].view(5)                    # layout order differs from the parameter order
    ps[0].grad = flat[5:17].view(4, 3)
    ps[2].grad = flat[17:21].view(2, 2)               # ps[3] has no gradient: skipped, as torch.optim.Adam does
    plist, offs = opt._offsets(flat)
    assert [id(p) for p in plist] == [id(ps[0]), id(ps[1]), id(ps[2])] and offs == [5, 0, 17]
    ps[2].grad = flat[16:20].view(2, 2)               # overlaps its neighbour, leaves a hole at the end
    assert opt._offsets(flat) is None
    ps[2].grad = torch.zeros(2, 2)                    # not a slice of the flat buffer
    assert opt._offsets(flat) is None
    ps[2].grad = flat[17:21].view(2, 2)
    assert opt._offsets(flat[:20]) is None            # buffer shorter than the gradients it should hold
    assert opt._offsets(flat.double()) is None
