# -*- coding: utf-8 -*-
"""CPU tier: the copy tables of nets/packing.py (what wnb_pack_weights executes on the device) emulated in numpy and
checked against the differentiable torch packing (WaveNet._pack / _pack_stack) and its autograd transpose."""
import ctypes

import numpy as np
import pytest
import torch

from pytorchwavenetvocoder_b200.nets import WaveNet
from pytorchwavenetvocoder_b200.nets.packing import ADD2, COPY, DESC, SRC_ABS, SUMPTR, StackPlan


def _f32_at(addr, n):
    return np.ctypeslib.as_array((ctypes.c_float * n).from_address(addr))


def _run_table(table, src_base, dst, scale=1.0):
    """numpy restatement of pack_kernel (csrc/pack.cu)"""
    descs = np.frombuffer(table.numpy().tobytes(), dtype=DESC)
    for d in descs:
        i0, i1, i2 = np.meshgrid(np.arange(d["n0"]), np.arange(d["n1"]), np.arange(d["n2"]), indexing="ij")
        so = (i0 * d["ss0"] + i1 * d["ss1"] + i2 * d["ss2"]).reshape(-1)
        do = (i0 * d["ds0"] + i1 * d["ds1"] + i2 * d["ds2"]).reshape(-1) + d["dst"]
        span = int(so.max()) + 1

        def src(a):
            return _f32_at(int(a), span) if d["flags"] & SRC_ABS else src_base[int(a):int(a) + span]
        if d["op"] == COPY:
            v = src(d["src"])[so]
        elif d["op"] == ADD2:
            v = src(d["src"])[so] + src(d["src2"])[so]
        else:
            assert d["op"] == SUMPTR
            tab = np.ctypeslib.as_array((ctypes.c_int64 * int(d["nsum"])).from_address(int(d["src"])))
            v = np.zeros(so.shape, np.float32)
            for a in tab:
                v = v + _f32_at(int(a), span)[so]
        assert len(np.unique(do)) == len(do)      # an entry never writes an element twice
        dst[do] = v * np.float32(scale)


@pytest.mark.parametrize("cfg_t", [(256, 28, 64, 512, 3, 2, 2, 80), (256, 20, 64, 96, 2, 1, 2, 0)])
def test_pack_and_unpack_tables(cfg_t):
    torch.manual_seed(0)
    net = WaveNet(*cfg_t)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn_like(p))
    plan = StackPlan(net, torch.device("cpu"))
    pbuf = plan.pbuf.numpy()
    _run_table(plan.pack_table, None, pbuf)
    wf, bf, W1, b1, W2res, b2res, Wskip, bskip, Wp1, bp1, Wp2, bp2 = net._pack_stack(net._pack())
    L, R, K1 = W1.size(0), net.n_resch, W1.size(2)
    wgate = torch.zeros(L, 3 * R, K1 + R)
    wgate[:, :2 * R, :K1] = W1.detach()
    wgate[:, 2 * R:, K1:] = W2res.detach().transpose(1, 2)
    want = dict(wf=wf, bf=bf, W1=W1, b1=b1, W2res=W2res, b2res=b2res, Wskip=Wskip, bskip=bskip, Wp1=Wp1, bp1=bp1, Wp2=Wp2,
                bp2=bp2, w1t=W1.transpose(1, 2), wgate=wgate, wskt=Wskip.t(), wp1t=Wp1.t(), wp2t=Wp2.t())
    for k, v in want.items():
        got = plan.p(k).numpy()
        np.testing.assert_allclose(got, v.detach().contiguous().numpy().reshape(got.shape), rtol=0,
                                   atol=1e-5 if k == "bskip" else 0, err_msg=k)
    # every float of the packed buffer outside the written sections is still the zero it was initialised with
    covered = np.zeros(pbuf.shape, bool)
    for k in want:
        o = plan.P.off[k]
        covered[o:o + int(np.prod(plan.P.shape[k]))] = True
    assert not pbuf[~covered].any()

    # unpack = autograd transpose of the packing, scaled
    g = torch.Generator().manual_seed(1)
    gbuf = torch.randn(plan.G.size, generator=g)
    packed = dict(zip(("wf", "bf", "W1", "b1", "W2res", "b2res", "Wskip", "bskip", "Wp1", "bp1", "Wp2", "bp2"),
                      (wf, bf, W1, b1, W2res, b2res, Wskip, bskip, Wp1, bp1, Wp2, bp2)))
    obj = sum((packed[k] * plan.G.view(gbuf, k)).sum() for k in packed)
    if net.upsampling_factor > 0:
        obj = obj + (net.upsampling.conv.weight.view(-1) * plan.G.view(gbuf, "upw")).sum() \
            + (net.upsampling.conv.bias.view(-1) * plan.G.view(gbuf, "upb")).sum()
    obj.backward()
    flat = torch.full((plan.grad_size,), float("nan"))
    _run_table(plan.unpack_table, gbuf.numpy(), flat.numpy(), scale=0.5)
    for (name, p), v in zip(plan.params, plan.grad_views(flat)):
        if name.startswith("res_1x1.%d." % (L - 1)):
            assert v is None and p.grad is None, name
            continue
        np.testing.assert_allclose(v.numpy(), 0.5 * p.grad.numpy(), rtol=1e-6, atol=1e-6, err_msg=name)
    # views are 16-byte aligned slices in parameters() order
    offs = [plan.grad_off[n] for n in plan.grad_names]
    assert offs == sorted(offs) and all(o % 4 == 0 for o in offs)
