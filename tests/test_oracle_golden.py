# -*- coding: utf-8 -*-
"""Pin the oracle (oracle/wavenet_oracle.py) to vectors recorded from the live reference
(tests/golden/make_golden.py).  CPU only; runs in the default ``-m "not gpu"`` suite."""
import os

import numpy as np
import pytest

from oracle import wavenet_oracle as O
from tests.golden.cases import (FORWARD_CASES, GEN_CASES, make_gen_inputs, make_inputs,
                                mulaw_edge_mask, mulaw_inputs, mulaw_pcm16_domain)

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_mulaw_bit_exact():
    g = np.load(os.path.join(G, "mulaw.npz"))
    x32, x64, codes = mulaw_inputs()
    got32 = O.encode_mu_law(x32, 256)   # exact on the recording host; SIMD-log dependent at edges elsewhere
    assert not np.any((got32 != g["mulaw_enc_f32"]) & ~mulaw_edge_mask(x32))
    assert np.array_equal(O.encode_mu_law(x64, 256), g["mulaw_enc_f64"])
    dec = O.decode_mu_law(codes, 256)   # float64; numpy pow is SVML or libm depending on the host
    fx = np.abs((codes - 0.5) / 255 * 2 - 1)
    assert np.all(np.abs(dec - g["mulaw_dec"]) <= 1.01 * np.spacing(256.0 ** fx) / 255)
    assert np.array_equal(O.encode_mu_law(mulaw_pcm16_domain(), 256), g["mulaw_enc_pcm16"].astype(np.int64))
    # known answers quoted in SURVEY.md 8c
    kat = np.array([-1, -.5, -.1, -.01, -1e-3, 0, 1e-3, .01, .1, .5, .999, 1])
    assert O.encode_mu_law(kat).tolist() == [0, 16, 52, 98, 122, 128, 133, 157, 203, 239, 255, 255]
    assert O.decode_mu_law(np.array([128]))[0] == 0.0


@pytest.mark.parametrize("name", sorted(FORWARD_CASES))
def test_forward_loss_grads(name):
    cfg_t, seed, B, T, start = FORWARD_CASES[name]
    cfg = O.Config(*cfg_t)
    g = np.load(os.path.join(G, "forward_%s.npz" % name))
    p = {k: v.astype(np.float64) for k, v in O.make_params(cfg, seed).items()}
    x, h, t = make_inputs(cfg, seed, B, T)
    y, cache = O.forward(cfg, p, x, h.astype(np.float64), return_cache=True)
    # fp64 oracle vs fp32 reference: the reference's own fp32-vs-fp64 gap is ~2e-6
    np.testing.assert_allclose(y, g["logits"], atol=2e-5, rtol=0)
    loss, dl = O.cross_entropy(y, t, start)
    assert abs(loss - float(g["loss"])) < 1e-5
    grads = O.backward(cfg, p, cache, dl)
    n_checked = 0
    for k in g.files:
        if k.startswith("grad."):
            ref = g[k]
            tol = 1e-6 + 2e-4 * np.abs(ref).max()
            np.testing.assert_allclose(grads[k[5:]], ref, atol=tol, rtol=0, err_msg=k)
            n_checked += 1
        elif k.startswith("nograd."):
            assert np.abs(grads[k[7:]]).max() == 0.0
    assert n_checked >= len(p) - 2


@pytest.mark.parametrize("name", sorted(GEN_CASES))
def test_generation_argmax(name):
    cfg_t, seed, B, T0, n_list, naive = GEN_CASES[name]
    cfg = O.Config(*cfg_t)
    g = np.load(os.path.join(G, "gen_%s.npz" % name))
    p = O.make_params(cfg, seed)   # float32, like the reference
    x, h = make_gen_inputs(cfg, seed, B, T0, n_list)
    U = cfg.upsampling_factor
    for b in range(B):
        n = n_list[b]
        nf = (n + T0 + U - 1) // U if U > 0 else n + T0
        got = O.fast_generate(cfg, p, x[b:b + 1], h[b:b + 1, :, :nf], n, mode="argmax")
        assert np.array_equal(got, g["fast_%d" % b]), (name, b)
    if B > 1:
        outs = O.batch_fast_generate(cfg, p, x, h, list(n_list), mode="argmax")
        for i, o in enumerate(outs):
            assert np.array_equal(o, g["batch_%d" % i]), (name, i)
    if naive and name != "tiny1000":
        got = O.generate_naive(cfg, p, x[:1], h[:1], min(n_list[0], 6), mode="argmax")
        assert np.array_equal(got, g["naive_0"][:len(got)])


def test_sampling_inverse_cdf_distribution():
    rng = np.random.RandomState(3)
    logits = rng.standard_normal(256).astype(np.float32) * 2
    u = rng.uniform(size=20000)
    draws = np.array([O._pick(logits, "sampling", ui) for ui in u])
    pr = np.exp(logits - logits.max())
    pr /= pr.sum()
    cnt = np.bincount(draws, minlength=256)
    chi2 = ((cnt - pr * len(u)) ** 2 / (pr * len(u) + 1e-9))[pr * len(u) > 5].sum()
    assert chi2 < 400  # dof ~ 200
