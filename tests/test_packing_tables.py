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
    # gap-free slices; every parameter whose size is a multiple of 4 floats starts 16-byte aligned
    offs = [plan.grad_off[n] for n in plan.grad_names]
    sizes = [dict(plan.params)[n].numel() for n in plan.grad_names]
    assert offs == sorted(offs) and offs[0] == 0 and all(o + s == o2 for o, s, o2 in zip(offs, sizes, offs[1:] + [plan.grad_size]))
    assert all(o % 4 == 0 for o, s in zip(offs, sizes) if s % 4 == 0)


def test_decode_cluster_stream_is_a_split_of_the_single_cta_stream():
    """_decode_warp_pack(W, CL=2): CTA r's stream = the tiles of warps r*W/2.. of every matrix + the full biases; sizes
    agree with wnb_decode_warp_floats (csrc/decode_warp.cu)."""
    from pytorchwavenetvocoder_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(1)
    net = WaveNet(256, 28, 64, 512, 3, 2, 2, 80)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn_like(p))
    L, W = 6, 16
    one = net._decode_warp_pack(W, 1)
    two = net._decode_warp_pack(W, 2)
    assert one.numel() == lib.wnb_decode_warp_floats(L, 1) and two.numel() == lib.wnb_decode_warp_floats(L, 2)
    half = two.numel() // 2
    jb1, jb2 = 4096, 2048                    # floats of one W1 j-block: 16 warps / 8 warps
    lay1 = 5 * jb1 + 128 + 4096 + 576 + 64 * 512
    lay2 = 5 * jb2 + 128 + 2048 + 576 + 64 * 256
    for r in range(2):
        st = two[r * half:(r + 1) * half]
        for l in range(L):
            a, b = one[l * lay1:(l + 1) * lay1], st[l * lay2:(l + 1) * lay2]
            for j in range(5):               # W1: [j][warp][...]: this CTA's 8 warps are contiguous inside a j-block
                assert torch.equal(b[j * jb2:(j + 1) * jb2], a[j * jb1 + r * jb2:j * jb1 + (r + 1) * jb2])
            oa, ob = 5 * jb1, 5 * jb2
            assert torch.equal(b[ob:ob + 128], a[oa:oa + 128])                       # b1 (full)
            oa, ob = oa + 128, ob + 128
            for j in range(2):               # W2res: [j][warp][g][lane][4], 128 floats per warp
                assert torch.equal(b[ob + j * 1024:ob + (j + 1) * 1024], a[oa + j * 2048 + r * 1024:oa + j * 2048 + (r + 1) * 1024])
            oa, ob = oa + 4096, ob + 2048
            assert torch.equal(b[ob:ob + 576], a[oa:oa + 576])                       # b2 (full)
            oa, ob = oa + 576, ob + 576
            assert torch.equal(b[ob:].view(64, 256), a[oa:].view(64, 512)[:, r * 256:(r + 1) * 256])
        pa, pb = one[L * lay1:], st[L * lay2:]
        assert torch.equal(pb[:768], pa[:768])                                       # bp1 | bp2
        assert torch.equal(pb[768:768 + 512 * 256].view(512, 256), pa[768:768 + 512 * 512].view(512, 512)[:, r * 256:(r + 1) * 256])
        assert torch.equal(pb[768 + 512 * 256:].view(512, 128), pa[768 + 512 * 512:].view(512, 256)[:, r * 128:(r + 1) * 128])
