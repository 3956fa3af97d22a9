# -*- coding: utf-8 -*-
"""CPU tier: the C-ABI library builds, loads without a GPU and exports every symbol the header declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "wnb200.h")).read()
    return sorted(set(re.findall(r"WNB_API[^;(]*?\b(wnb_\w+)\s*\(", src)))


def test_build_and_exports():
    from pytorchwavenetvocoder_b200 import _lib, build
    build.build()
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 18
    assert sorted(_lib.SIGNATURES) == syms          # binding covers exactly the header
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.wnb_version() >= 100
    assert lib.wnb_launch_count() == 0 or lib.wnb_launch_count() > 0


def test_argument_validation_without_gpu():
    from pytorchwavenetvocoder_b200 import _lib
    lib = _lib.load()
    # invalid shapes are rejected before any CUDA call, with a message
    rc = lib.wnb_front_embed_fwd(None, None, None, None, 0, 0, 256, 8, 2, None)
    assert rc == -1 and b"bad shape" in lib.wnb_last_error()
    rc = lib.wnb_cross_entropy(None, None, None, None, 1, 10, 256, 10, None)
    assert rc == -1


def test_stack_entry_points_host_logic_without_gpu():
    """Shape coverage, workspace arithmetic and argument checks of the deferred-skip stack ABI (no kernels run)."""
    import ctypes
    from pytorchwavenetvocoder_b200 import _lib
    lib = _lib.load()
    TF32, FP32 = _lib.MATH_TF32, _lib.MATH_FP32
    assert lib.wnb_stack_supported(64, 512, 32, 2, 30, TF32) == 1        # BASELINE family
    assert lib.wnb_stack_supported(64, 256, 32, 2, 7, TF32) == 1
    assert lib.wnb_stack_supported(64, 512, 32, 2, 30, FP32) == 0        # fp32 mode: per-block FFMA path
    assert lib.wnb_stack_supported(512, 256, 96, 3, 30, TF32) == 0       # recipe shape: composed per-block path
    assert lib.wnb_stack_supported(64, 384, 32, 2, 30, TF32) == 0        # skip GEMM needs S <= 256 or S == 512
    L, B, T, R = 30, 8, 23040, 64
    bt = B * T
    assert lib.wnb_stack_bwd_workspace(L, B, T, R, 512, 32, 2) == 4 * (bt * L * R + L * R + L * bt * 2 * R + L * bt * R)
    rc = lib.wnb_stack_fwd(None, 2, None, None, None, None, None, None, None, None, None, None, 30, 8, 100, 64, 512,
                           32, 2, 1, None)
    assert rc == -1 and b"null pointer" in lib.wnb_last_error()
    rc = lib.wnb_resblock_fwd_z(None, None, None, None, None, None, None, None, 64, 0, 0, 0, 64, 32, 2, 1, None)
    assert rc == -1 and b"bad shape" in lib.wnb_last_error()
    tot, n = ctypes.c_double(0.0), ctypes.c_int(0)
    assert lib.wnb_profile_read(99, ctypes.byref(tot), ctypes.byref(n)) == -1
    assert lib.wnb_profile_enable(0) == 0 and lib.wnb_profile_read(0, ctypes.byref(tot), ctypes.byref(n)) == 0 and n.value == 0


def test_no_cpu_fallback():
    import torch
    from pytorchwavenetvocoder_b200 import _lib
    from pytorchwavenetvocoder_b200.nets import WaveNet
    net = WaveNet(256, 28, 8, 16, 4, 1, 2, 0)
    with pytest.raises(_lib.WnbError):
        net(torch.zeros(1, 32, dtype=torch.long), torch.zeros(1, 28, 32))
    with pytest.raises(_lib.WnbError):
        net.fast_generate(torch.zeros(1, 1, dtype=torch.long), torch.zeros(1, 28, 10), 5, mode="argmax")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pytorchwavenetvocoder_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, fn


def test_stack_weight_packing_on_cpu():
    """WaveNet._pack_stack (pure torch ops, no kernels): layouts documented in include/wnb200.h, and gradients flow
    back to the reference-shaped parameters (the last block's res_1x1 stays out of the graph, like the reference)."""
    import numpy as np
    import torch
    from pytorchwavenetvocoder_b200.nets import WaveNet
    torch.manual_seed(0)
    net = WaveNet(256, 28, 64, 96, 3, 2, 2, 0)
    for prm in net.parameters():
        torch.nn.init.normal_(prm, std=0.1)
    wf, bf, W1, b1, W2res, b2res, Wskip, bskip, Wp1, bp1, Wp2, bp2 = net._pack_stack(net._pack())
    L, R, S = 6, 64, 96
    assert W1.shape == (L, 2 * R, 2 * R + 32) and W2res.shape == (L, R, R) and Wskip.shape == (S, L * R)
    sd = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    for l in range(L):
        assert np.array_equal(Wskip[:, l * R:(l + 1) * R].detach().numpy(), sd["skip_1x1.%d.weight" % l][:, :, 0])
        assert np.array_equal(W2res[l].detach().numpy(), sd["res_1x1.%d.weight" % l][:, :, 0])
        # W1 rows: sigmoid branch then tanh branch; columns: tap 0 (oldest), tap 1, aux (zero padded to 32)
        assert np.array_equal(W1[l, :R, :R].detach().numpy(), sd["dil_sigmoid.%d.conv.weight" % l][:, :, 0])
        assert np.array_equal(W1[l, R:, R:2 * R].detach().numpy(), sd["dil_tanh.%d.conv.weight" % l][:, :, 1])
        assert np.array_equal(W1[l, :R, 2 * R:2 * R + 28].detach().numpy(), sd["aux_1x1_sigmoid.%d.weight" % l][:, :, 0])
        assert np.all(W1[l, :, 2 * R + 28:].detach().numpy() == 0)
    assert np.allclose(bskip.detach().numpy(), sum(sd["skip_1x1.%d.bias" % l] for l in range(L)), atol=1e-6)
    (Wskip.sum() + W2res.sum() + b2res.sum() + bskip.sum()).backward()
    assert net.skip_1x1[0].weight.grad is not None and net.res_1x1[0].weight.grad is not None
    assert net.res_1x1[L - 1].weight.grad is None and net.res_1x1[L - 1].bias.grad is None
