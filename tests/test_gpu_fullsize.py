# -*- coding: utf-8 -*-
"""GPU tier, BASELINE.json full sizes (configs[1] train batch 8 x 23 040, configs[3] 64-utterance decode): the oracle
cannot run these in seconds, so the checks are size-independent properties of the path."""
import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests.util import our_model

pytestmark = pytest.mark.gpu
CFG = (256, 28, 64, 512, 10, 3, 2, 80)


@pytest.fixture(scope="module")
def net():
    cfg = O.Config(*CFG)
    return cfg, our_model(cfg, O.make_params(cfg, 5), math_mode="tf32")


def test_train_forward_causality_and_batch_independence(net):
    cfg, model = net
    model.eval()
    g = torch.Generator().manual_seed(1)
    B, T = 8, 23040
    x = torch.randint(0, 256, (B, T), generator=g).cuda()
    h = torch.randn(B, 28, T // 80, generator=g).cuda()
    with torch.no_grad():
        y = model(x, h)
        assert y.shape == (B, T, 256) and torch.isfinite(y).all()
        # causality: changing samples at t >= t0 (and aux frames from t0 on) leaves every earlier logit bit-identical
        t0 = 12800
        x2, h2 = x.clone(), h.clone()
        x2[:, t0:] = (x2[:, t0:] + 11) % 256
        h2[:, :, t0 // 80:] += 1.0
        y2 = model(x2, h2)
        assert torch.equal(y[:, :t0], y2[:, :t0])
        assert not torch.equal(y[:, t0:], y2[:, t0:])
        # batch independence: a row computed alone equals the same row inside the batch (same tiles, same order)
        y1 = model(x[3:4], h[3:4])
        assert torch.equal(y1[0], y[3])
        # receptive field: the residual stack sees receptive_field samples of the front conv's output, and the front
        # conv (kernel_size taps) adds kernel_size - 1 more: logits at t >= t1 + rf - 1 + (ks - 1) cannot depend on
        # samples before t1, earlier ones do
        t1 = 4096
        x3 = x.clone()
        x3[:, :t1] = 128
        y3 = model(x3, h)
        lim = t1 + cfg.receptive_field - 1 + (cfg.kernel_size - 1)
        assert torch.equal(y[:, lim:], y3[:, lim:])
        assert not torch.equal(y[:, t1:lim], y3[:, t1:lim])


def test_train_step_full_size_decreases_loss(net):
    from pytorchwavenetvocoder_b200.nets import cross_entropy
    cfg, _ = net
    model = our_model(cfg, O.make_params(cfg, 6), math_mode="tf32").train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    g = torch.Generator().manual_seed(2)
    B, T = 8, 23040
    tt = torch.arange(T + 1) / 16000.0
    wav = torch.stack([0.4 * torch.sin(2 * np.pi * (150 + 25 * b) * tt) for b in range(B)])
    q = torch.from_numpy(O.encode_mu_law(wav.numpy(), 256))
    x, t = q[:, :-1].contiguous().cuda(), q[:, 1:].contiguous().cuda()
    h = torch.randn(B, 28, T // 80, generator=g).cuda()
    losses = []
    for _ in range(8):
        loss = cross_entropy(model(x, h), t, cfg.receptive_field)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and losses[-1] < losses[0] - 0.5, losses
    # the last block's res_1x1 never receives a gradient (reference wavenet.py:230-238)
    assert model.res_1x1[-1].weight.grad is None and model.res_1x1[0].weight.grad is not None


def test_decode_64_utterances_determinism_and_batching(net):
    """configs[3] shape (64 utterances, seed 128, 30 layers): argmax decoding is deterministic, independent of how the
    utterances are grouped into CTAs (NU = 1 vs 2 vs 4) and of which other utterances share the launch."""
    cfg, model = net
    model.eval()
    g = torch.Generator().manual_seed(3)
    B, n = 64, 200
    x = torch.full((B, 1), 128, dtype=torch.long).cuda()
    h = torch.randn(B, 28, (n + 80) // 80, generator=g).cuda()
    with torch.no_grad():
        a = model._decode(x, h, [n] * B, "argmax").cpu()
        b = model._decode(x, h, [n] * B, "argmax").cpu()
        assert torch.equal(a, b)
        sub = model._decode(x[:5], h[:5], [n] * 5, "argmax").cpu()
        assert torch.equal(sub, a[:5])
        # 320 utterances = the 64 repeated 5x -> NU = 4 path; every copy must reproduce the NU = 1 result
        big = model._decode(x.repeat(5, 1), h.repeat(5, 1, 1), [n] * (5 * B), "argmax").cpu()
        assert torch.equal(big.view(5, B, n)[4], a)
        # ragged lengths: each utterance is a prefix of its full-length run
        nl = [n - (i % 7) * 9 for i in range(B)]
        r = model._decode(x, h, nl, "argmax").cpu()
        for i in range(B):
            assert torch.equal(r[i, :nl[i]], a[i, :nl[i]])
    assert a.min() >= 0 and a.max() <= 255


def test_decode_recipe_shape_vs_oracle():
    """The recipes' shape (512 res / 256 skip) goes through the streaming decode kernel: a few steps vs the oracle."""
    cfg = O.Config(256, 28, 512, 256, 3, 1, 2, 8)
    p = O.make_params(cfg, 21)
    model = our_model(cfg, p).eval()
    rng = np.random.RandomState(4)
    x = np.full((2, 1), 128, np.int64)
    h = rng.standard_normal((2, 28, 3)).astype(np.float32)
    nl = [14, 9]
    outs, olg = O.batch_fast_generate(cfg, p, x, h, list(nl), mode="argmax", return_logits=True)
    with torch.no_grad():
        gen, lg = model._decode(torch.from_numpy(x).cuda(), torch.from_numpy(h).cuda(), nl, "argmax", return_logits=True)
    gen, lg = gen.cpu().numpy(), lg.cpu().numpy()
    assert model.last_decode_kernel in ("stream", "direct")
    order = sorted(range(2), key=lambda b: (nl[b], b))
    for i, b in enumerate(order):
        assert np.array_equal(gen[b, :nl[b]], outs[i])
        np.testing.assert_allclose(lg[b, :nl[b]], olg[b, :nl[b]], atol=2e-4, rtol=0)
