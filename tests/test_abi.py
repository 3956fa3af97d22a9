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
