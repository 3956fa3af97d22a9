# -*- coding: utf-8 -*-
"""GPU tier: wnb_pack_weights (csrc/pack.cu + nets/packing.py) against the differentiable torch packing it replaces
(WaveNet._pack / _pack_stack), and the single-node training entry WaveNet.forward_loss against forward + cross_entropy."""
import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests.util import our_model

pytestmark = pytest.mark.gpu


def _net(cfg_t=(256, 28, 64, 512, 3, 2, 2, 80), seed=4):
    cfg = O.Config(*cfg_t)
    return cfg, our_model(cfg, O.make_params(cfg, seed), math_mode="tf32").train()


@pytest.mark.parametrize("cfg_t", [(256, 28, 64, 512, 3, 2, 2, 80), (256, 20, 64, 96, 2, 1, 2, 0)])
def test_pack_kernel_matches_torch_packing(cfg_t):
    from pytorchwavenetvocoder_b200 import _lib
    cfg, net = _net(cfg_t)
    plan = net._stack_plan(torch.device("cuda", torch.cuda.current_device()))
    assert plan is not None
    plan.pack(_lib.stream())
    torch.cuda.synchronize()
    with torch.no_grad():
        wf, bf, W1, b1, W2res, b2res, Wskip, bskip, Wp1, bp1, Wp2, bp2 = net._pack_stack(net._pack())
        L, R = W1.size(0), net.n_resch
        K1 = W1.size(2)
        wgate = torch.zeros(L, 3 * R, K1 + R, device="cuda")
        wgate[:, :2 * R, :K1] = W1
        wgate[:, 2 * R:, K1:] = W2res.transpose(1, 2)
        want = dict(wf=wf, bf=bf, W1=W1, b1=b1, W2res=W2res, b2res=b2res, Wskip=Wskip, Wp1=Wp1, bp1=bp1, Wp2=Wp2, bp2=bp2,
                    w1t=W1.transpose(1, 2), wgate=wgate, wskt=Wskip.t(), wp1t=Wp1.t(), wp2t=Wp2.t())
        for k, v in want.items():
            assert torch.equal(plan.p(k), v.contiguous().view(plan.P.shape[k])), k
        # summed skip biases: table order 0..L-1 in fp32 (torch's reduction order may differ in the last bit)
        torch.testing.assert_close(plan.p("bskip"), bskip, rtol=0, atol=1e-6)
    # a second pack after an in-place parameter update picks the new values up (same storage, no rebuild)
    with torch.no_grad():
        net.dil_tanh[1].conv.weight.mul_(2.0)
    assert not plan.stale()
    plan.pack(_lib.stream())
    torch.cuda.synchronize()
    with torch.no_grad():
        assert torch.equal(plan.p("W1"), net._pack()[2])


def test_unpack_is_the_transpose_of_the_torch_packing():
    """gradient routing: <packed(params), Gbuf> differentiated by autograd == unpack(Gbuf) (x the upstream scale)"""
    from pytorchwavenetvocoder_b200 import _lib
    cfg, net = _net()
    dev = torch.device("cuda", torch.cuda.current_device())
    plan = net._stack_plan(dev)
    torch.manual_seed(0)
    gbuf = torch.randn(plan.G.size, device=dev)
    packed = dict(zip(("wf", "bf", "W1", "b1", "W2res", "b2res", "Wskip", "bskip", "Wp1", "bp1", "Wp2", "bp2"),
                      net._pack_stack(net._pack())))
    obj = sum((packed[k] * plan.G.view(gbuf, k)).sum() for k in packed)
    obj = obj + (net.upsampling.conv.weight.view(-1) * plan.G.view(gbuf, "upw")).sum() \
        + (net.upsampling.conv.bias.view(-1) * plan.G.view(gbuf, "upb")).sum()
    obj.backward()
    flat = torch.full((plan.grad_size,), float("nan"), device=dev)
    scale = torch.tensor(2.0, device=dev)
    plan.unpack(gbuf, flat, scale, _lib.stream())
    torch.cuda.synchronize()
    views = plan.grad_views(flat)
    L = len(net.dilations)
    for (name, p), v in zip(plan.params, views):
        if name.startswith("res_1x1.%d." % (L - 1)):
            assert v is None and p.grad is None, name      # reference: no gradient for the last block's res_1x1
            continue
        torch.testing.assert_close(v, 2.0 * p.grad, rtol=1e-6, atol=1e-6, msg=name)


def test_forward_loss_matches_forward_plus_cross_entropy():
    """WaveNet.forward_loss (one autograd node, loss fused) == cross_entropy(forward(x, h), t, rf): same loss, same
    gradients (up to the order of the weight-gradient atomics), p.grad are views of ONE flat buffer."""
    from pytorchwavenetvocoder_b200.nets import cross_entropy
    cfg, net = _net((256, 28, 64, 512, 4, 2, 2, 16))
    rng = np.random.RandomState(1)
    B, T = 2, 640
    x = torch.from_numpy(rng.randint(0, 256, size=(B, T)).astype(np.int64)).cuda()
    t = torch.from_numpy(rng.randint(0, 256, size=(B, T)).astype(np.int64)).cuda()
    h = torch.from_numpy(rng.standard_normal((B, 28, T // 16)).astype(np.float32)).cuda()
    loss_a = cross_entropy(net(x, h), t, net.receptive_field)
    loss_a.backward()
    ga = {k: (None if v.grad is None else v.grad.clone()) for k, v in net.named_parameters()}
    net.zero_grad(set_to_none=True)
    loss_b = net.forward_loss(x, h, t)
    (3.0 * loss_b).backward()
    assert abs(loss_a.item() - loss_b.item()) < 1e-6
    flat = net._wnb_flat_grad
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
    for k, v in net.named_parameters():
        if ga[k] is None:
            assert v.grad is None, k
            continue
        assert lo <= v.grad.data_ptr() < hi, k                  # stolen by autograd as a view of the flat buffer
        torch.testing.assert_close(v.grad, 3.0 * ga[k], rtol=2e-3, atol=1e-6 * float(ga[k].abs().max()) + 1e-9, msg=k)
    # no-grad inference keeps working (ping-pong residual buffers)
    with torch.no_grad():
        y = net(x, h)
    assert tuple(y.shape) == (B, T, 256)
