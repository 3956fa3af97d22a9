# -*- coding: utf-8 -*-
"""CPU tier: the algebra behind the deferred-skip stack (pytorchwavenetvocoder_b200/csrc/stack.cu),
restated in numpy float64 on top of the oracle's per-block caches and checked against the oracle's own per-block
forward / backward (reference wavenet.py:229-238 and :525-536):

  forward    sum_l skip_1x1_l(z_l)                 ==  Z_all Wskip^T + bskip
  backward   dz_l                                  ==  (dskip Wskip)[:, l*R:(l+1)*R] + dout_l W2res_l
             skip_1x1_l.weight.grad                ==  (dskip^T Z_all)[:, l*R:(l+1)*R]
             skip_1x1_l.bias.grad                  ==  colsum(dskip)          (the same for every block)
             [pre | dz_res] as ONE GEMM over the block matrix [[W1, 0], [0, W2res^T]]

The GPU tier (tests/test_gpu_tc.py) checks that the kernels compute exactly these expressions; this file pins the
expressions themselves to the reference's arithmetic, with no GPU.
"""
import numpy as np
import pytest

from oracle import wavenet_oracle as O


def _setup(cfg, seed, B, T):
    p = {k: v.astype(np.float64) for k, v in O.make_params(cfg, seed).items()}
    rng = np.random.RandomState(seed + 1)
    x = rng.randint(0, cfg.n_quantize, size=(B, T)).astype(np.int64)
    t = rng.randint(0, cfg.n_quantize, size=(B, T)).astype(np.int64)
    Tf = T // cfg.upsampling_factor if cfg.upsampling_factor > 0 else T
    h = rng.standard_normal((B, cfg.n_aux, Tf))
    y, cache = O.forward(cfg, p, x, h, return_cache=True)
    _, dl = O.cross_entropy(y, t, 4)
    g = O.backward(cfg, p, cache, dl)
    return p, cache, dl, g


@pytest.mark.parametrize("cfg,B,T", [(O.Config(256, 28, 8, 16, 4, 1, 2, 0), 2, 40),
                                     (O.Config(256, 5, 16, 32, 3, 2, 2, 4), 1, 48)])
def test_skip_sum_is_one_gemm_over_concatenated_gate_outputs(cfg, B, T):
    p, cache, _, _ = _setup(cfg, 3, B, T)
    L, R = len(cfg.dilations), cfg.n_resch
    # Z_all (B, T, L*R) channels-last; Wskip[s][l*R + c] = skip_1x1_l.weight[s][c]
    zall = np.concatenate([np.transpose(cache["caches"][l][2], (0, 2, 1)) for l in range(L)], axis=2)
    wskip = np.concatenate([p["skip_1x1.%d.weight" % l][:, :, 0] for l in range(L)], axis=1)
    bskip = sum(p["skip_1x1.%d.bias" % l] for l in range(L))
    skip = zall @ wskip.T + bskip
    ref = np.transpose(cache["skip_sum"], (0, 2, 1))
    assert zall.shape == (B, T, L * R) and np.abs(skip - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("cfg,B,T", [(O.Config(256, 28, 8, 16, 4, 1, 2, 0), 2, 40),
                                     (O.Config(256, 5, 16, 32, 3, 2, 2, 4), 1, 48)])
def test_hoisted_backward_identities(cfg, B, T):
    p, cache, dl, g = _setup(cfg, 5, B, T)
    L, R, S = len(cfg.dilations), cfg.n_resch, cfg.n_skipch
    r0, h1, r1 = cache["post"]
    # dskip exactly as the oracle forms it (channels-last here)
    dy = np.transpose(dl, (0, 2, 1))
    dr1 = np.einsum("oc,bot->bct", p["conv_post_2.weight"][:, :, 0], dy)
    dh1 = dr1 * (h1 > 0)
    dr0 = np.einsum("oc,bot->bct", p["conv_post_1.weight"][:, :, 0], dh1)
    dskip = np.transpose(dr0 * (cache["skip_sum"] > 0), (0, 2, 1))                      # (B, T, S)
    zall = np.concatenate([np.transpose(cache["caches"][l][2], (0, 2, 1)) for l in range(L)], axis=2)
    wskip = np.concatenate([p["skip_1x1.%d.weight" % l][:, :, 0] for l in range(L)], axis=1)   # (S, L*R)
    dzall = dskip @ wskip                                                               # (B, T, L*R)
    dwskip = np.einsum("bts,btn->sn", dskip, zall)                                      # (S, L*R)
    dbskip = dskip.sum(axis=(0, 1))
    for l in range(L):
        sl = slice(l * R, (l + 1) * R)
        ref_w = g["skip_1x1.%d.weight" % l][:, :, 0]
        assert np.abs(dwskip[:, sl] - ref_w).max() <= 1e-12 * max(1.0, np.abs(ref_w).max()), l
        ref_b = g["skip_1x1.%d.bias" % l]
        assert np.abs(dbskip - ref_b).max() <= 1e-12 * max(1.0, np.abs(ref_b).max()), l
    # the per-block chain with the hoisted dz part: walk the blocks in reverse like wnb_stack_bwd
    h = np.transpose(cache["h"], (0, 2, 1))                                             # (B, T, A)
    dout = None
    for l in reversed(range(L)):
        d = cfg.dilations[l]
        x = np.transpose(cache["xs"][l], (0, 2, 1))                                     # (B, T, R)
        sg, th, z = (np.transpose(a, (0, 2, 1)) for a in cache["caches"][l])
        ks, A = cfg.kernel_size, cfg.n_aux
        # W1 (2R, ks*R + A): both branches, taps oldest first, then aux
        def branch(name, aux):
            wd = p["%s.%d.conv.weight" % (name, l)]                                     # (R, R, ks)
            return np.concatenate([wd[:, :, j] for j in range(ks)] + [p["%s.%d.weight" % (aux, l)][:, :, 0]], axis=1)
        W1 = np.concatenate([branch("dil_sigmoid", "aux_1x1_sigmoid"), branch("dil_tanh", "aux_1x1_tanh")], axis=0)
        K1 = ks * R + A
        W2res = p["res_1x1.%d.weight" % l][:, :, 0]                                     # [o][c]
        wgate = np.zeros((3 * R, K1 + R))
        wgate[:2 * R, :K1] = W1
        wgate[2 * R:, K1:] = W2res.T
        # A operand rows: [x(t-(ks-1)d) ... x(t) | aux(t) | dout(t)], zero for negative time
        taps = []
        for j in range(ks):
            s = (ks - 1 - j) * d
            xs_ = np.zeros_like(x)
            if s < T:
                xs_[:, s:] = x[:, :T - s]
            taps.append(xs_)
        dout_l = np.zeros_like(x) if dout is None else dout
        arow = np.concatenate(taps + [h, dout_l], axis=2)                               # (B, T, K1 + R)
        acc = arow @ wgate.T                                                            # (B, T, 3R)
        b1 = np.concatenate([p["dil_sigmoid.%d.conv.bias" % l] + p["aux_1x1_sigmoid.%d.bias" % l],
                             p["dil_tanh.%d.conv.bias" % l] + p["aux_1x1_tanh.%d.bias" % l]])
        pre = acc[:, :, :2 * R] + b1
        sg2 = 1.0 / (1.0 + np.exp(-pre[:, :, :R]))
        th2 = np.tanh(pre[:, :, R:])
        assert np.abs(sg2 - sg).max() <= 1e-12 and np.abs(th2 - th).max() <= 1e-12      # the gate recompute
        dz = dzall[:, :, l * R:(l + 1) * R] + acc[:, :, 2 * R:]                         # hoisted skip part + residual part
        da = dz * th2 * sg2 * (1 - sg2)
        dg = dz * sg2 * (1 - th2 * th2)
        dpre = np.concatenate([da, dg], axis=2)                                         # (B, T, 2R)
        # weight gradients of this block from dpre (what the segmented dW1 launch accumulates)
        dW1 = np.einsum("btm,btk->mk", dpre, arow[:, :, :K1])
        ref_sig = g["dil_sigmoid.%d.conv.weight" % l]
        for j in range(ks):
            got = dW1[:R, j * R:(j + 1) * R]
            assert np.abs(got - ref_sig[:, :, j]).max() <= 1e-11 * max(1.0, np.abs(ref_sig).max()), (l, j)
        ref_aux_t = g["aux_1x1_tanh.%d.weight" % l][:, :, 0]
        assert np.abs(dW1[R:, ks * R:] - ref_aux_t).max() <= 1e-11 * max(1.0, np.abs(ref_aux_t).max()), l
        if dout is not None:   # dW2res_l = dout_l^T z_l (the last block has no residual output)
            ref_res = g["res_1x1.%d.weight" % l][:, :, 0]
            got = np.einsum("bto,btc->oc", dout_l, z)
            assert np.abs(got - ref_res).max() <= 1e-11 * max(1.0, np.abs(ref_res).max()), l
        # dx_l = dout_l + sum_j dpre(t + (ks-1-j)d) W1[:, tap j]
        dx = dout_l.copy()
        for j in range(ks):
            s = (ks - 1 - j) * d
            contrib = dpre @ W1[:, j * R:(j + 1) * R]                                   # gradient w.r.t. x(t - s)
            if s < T:
                dx[:, :T - s] += contrib[:, s:]
        dout = dx
    # the gradient that reaches the front embedding must be the oracle's
    ref_front_b = g["causal.conv.bias"]
    assert np.abs(dout.sum(axis=(0, 1)) - ref_front_b).max() <= 1e-10 * max(1.0, np.abs(ref_front_b).max())
