# -*- coding: utf-8 -*-
"""B200-native mirror of ``wavenet_vocoder/nets/wavenet.py`` (reference v0.1.1).

Same import surface, constructor arguments, attributes, method signatures, ``state_dict`` keys and
error behaviour as the reference module, but every hot-path computation is a
hand-written sm_100a kernel reached through the C ABI in ``include/wnb200.h``:

* ``WaveNet.forward``  (reference wavenet.py:212-241)  -> front gather, aux up-sampling, ONE fused
  kernel per residual block, post network; backward through ``torch.autograd.Function``.
* ``fast_generate`` / ``batch_fast_generate`` (wavenet.py:309-511) -> one persistent kernel launch.
* ``encode_mu_law`` / ``decode_mu_law`` (wavenet.py:17-47) -> device kernels, numpy in / numpy out.

PyTorch only owns the memory, the parameters and the streams.  There is no CPU fallback: calling the
model with CPU tensors raises.
"""
import ctypes
import logging
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib
from .._lib import MATH_FP32, MATH_TF32, MODE_ARGMAX, MODE_SAMPLING, check, ptr, stream

__all__ = ["encode_mu_law", "decode_mu_law", "initialize", "OneHot", "CausalConv1d", "UpSampling",
           "WaveNet", "cross_entropy", "mulaw_pcm16_table", "codes_to_pcm16"]

# bench.py sets this to a list to get (start, end) CUDA-event pairs around every fused-block launch
PROFILE_EVENTS = None


def tc_supported(cfg):
    """True when the tcgen05 (tf32) fused-block kernel covers WaveNet(*cfg)."""
    lib = _lib.load()
    Q, A, R, S, depth, repeat, ks, U = cfg
    return bool(lib.wnb_resblock_fwd_supported(R, S, _round_up(A, 32), ks, MATH_TF32))


def _device():
    if not torch.cuda.is_available():
        raise _lib.WnbError("no CUDA device: the wnb200 kernels need a B200 (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


# ------------------------------------------------------------------------------------------
# mu-law (reference wavenet.py:17-47): numpy in, numpy out, computed on the GPU
# ------------------------------------------------------------------------------------------
def encode_mu_law(x, mu=256):
    """PERFORM MU-LAW ENCODING (reference wavenet.py:17-30).

    Args:
        x (ndarray): Audio signal with the range from -1 to 1 (float32 or float64).
        mu (int): Quantized level.

    Returns:
        ndarray: int64 quantized signal with the range from 0 to mu - 1.
    """
    lib = _lib.load()
    x = np.asarray(x)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    dev = _device()
    xt = torch.from_numpy(np.ascontiguousarray(x).reshape(-1)).to(dev)
    yt = torch.empty(xt.numel(), dtype=torch.int64, device=dev)
    fn = lib.wnb_mulaw_encode_f32 if x.dtype == np.float32 else lib.wnb_mulaw_encode_f64
    check(fn(ptr(xt), ptr(yt), xt.numel(), int(mu), stream()), "mulaw_encode")
    return yt.cpu().numpy().reshape(x.shape)


def decode_mu_law(y, mu=256):
    """PERFORM MU-LAW DECODING (reference wavenet.py:33-47); returns float64 like numpy does."""
    lib = _lib.load()
    y = np.asarray(y)
    dev = _device()
    yt = torch.from_numpy(np.ascontiguousarray(y).astype(np.int64).reshape(-1)).to(dev)
    xt = torch.empty(yt.numel(), dtype=torch.float64, device=dev)
    check(lib.wnb_mulaw_decode_f64(ptr(yt), ptr(xt), yt.numel(), int(mu), stream()), "mulaw_decode")
    return xt.cpu().numpy().reshape(y.shape)


def mulaw_pcm16_table(mu=256):
    """int16 PCM value of every mu-law code: ``decode_mu_law`` (reference wavenet.py:33-47) followed by the PCM_16
    quantisation of ``utils.write_wav``'s stdlib writer (round-half-even of x * 32768, clipped to int16)."""
    x = decode_mu_law(np.arange(mu), mu)
    return np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)


def codes_to_pcm16(codes, mu=256):
    """(B, n) int32 CUDA tensor of generated codes -> (B, n) int16 CUDA tensor (one gather kernel for the whole batch)."""
    lib = _lib.load()
    codes = codes.contiguous()
    if codes.dtype != torch.int32:
        codes = codes.to(torch.int32)
    tab = torch.from_numpy(mulaw_pcm16_table(mu)).to(codes.device)
    out = torch.empty(codes.shape, dtype=torch.int16, device=codes.device)
    check(lib.wnb_lut_i16(ptr(codes), ptr(tab), ptr(out), codes.numel(), int(mu), stream()), "lut_i16")
    return out


def initialize(m):
    """INITILIZE CONV WITH XAVIER (reference wavenet.py:50-63)."""
    if isinstance(m, nn.Conv1d):
        nn.init.xavier_uniform_(m.weight)
        nn.init.constant_(m.bias, 0.0)
    if isinstance(m, nn.ConvTranspose2d):
        nn.init.constant_(m.weight, 1.0)
        nn.init.constant_(m.bias, 0.0)


# ------------------------------------------------------------------------------------------
# autograd functions wrapping the C ABI
# ------------------------------------------------------------------------------------------
class _UpsampleFn(torch.autograd.Function):
    """aux up-sampling + (B,A,Tf)->(B,T,Ap) layout change (reference wavenet.py:124-154)."""

    @staticmethod
    def forward(ctx, h, w, b, Ap):
        lib = _lib.load()
        h = h.contiguous().float()
        B, A, Tf = h.shape
        U = 0 if w is None else w.numel()
        T = Tf * U if U > 0 else Tf
        out = torch.empty(B, T, Ap, device=h.device, dtype=torch.float32)
        wc = None if w is None else w.contiguous().view(-1)
        bc = None if b is None else b.contiguous().view(-1)
        check(lib.wnb_aux_upsample_fwd(ptr(h), ptr(wc), ptr(bc), ptr(out), B, A, Ap, Tf, U, stream()),
              "aux_upsample_fwd")
        ctx.save_for_backward(h)
        ctx.meta = (B, A, Ap, Tf, U, None if w is None else w.shape, None if b is None else b.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        (h,) = ctx.saved_tensors
        B, A, Ap, Tf, U, wshape, bshape = ctx.meta
        if U == 0:
            return None, None, None, None
        dout = dout.contiguous()
        dw = torch.zeros(U, device=h.device, dtype=torch.float32)
        db = torch.zeros(1, device=h.device, dtype=torch.float32)
        check(lib.wnb_aux_upsample_bwd(ptr(h), ptr(dout), ptr(dw), ptr(db), B, A, Ap, Tf, U, stream()),
              "aux_upsample_bwd")
        return None, dw.view(wshape), db.view(bshape), None


class _CrossEntropyFn(torch.autograd.Function):
    """CrossEntropyLoss(mean) over logits[:, start:] fused with its gradient (bin/train.py:534-536)."""

    @staticmethod
    def forward(ctx, logits, target, start):
        lib = _lib.load()
        logits = logits.contiguous()
        target = target.contiguous()
        B, T, Q = logits.shape
        if target.dtype != torch.int64 or tuple(target.shape) != (B, T) or target.device != logits.device:
            raise ValueError("cross_entropy: target should be an int64 (B, T) = %s tensor on %s, got %s %s on %s"
                             % ((B, T), logits.device, target.dtype, tuple(target.shape), target.device))
        loss = torch.zeros(1, device=logits.device, dtype=torch.float64)
        need = ctx.needs_input_grad[0]
        dl = torch.empty_like(logits) if need else None
        check(lib.wnb_cross_entropy(ptr(logits), ptr(target), ptr(loss), ptr(dl), B, T, Q, int(start), stream()),
              "cross_entropy")
        ctx.dl = dl
        return loss[0].float()

    @staticmethod
    def backward(ctx, g):
        dl = ctx.dl
        ctx.dl = None
        return dl * g, None, None


def cross_entropy(logits, target, start=0):
    """Mean cross entropy over ``logits[:, start:]`` (B,T,Q) vs ``target[:, start:]`` (B,T)."""
    return _CrossEntropyFn.apply(logits, target, start)


class _WaveNetFn(torch.autograd.Function):
    """WaveNet.forward (reference wavenet.py:212-241) on packed weights."""

    @staticmethod
    def forward(ctx, x, haux, wf, bf, W1, b1, W2, b2, Wp1, bp1, Wp2, bp2, meta):
        lib = _lib.load()
        Q, R, S, Ap, ks, dilations, math_mode = meta
        B, T = x.shape
        L = len(dilations)
        dev = x.device
        x = x.contiguous()
        need_grad = any(ctx.needs_input_grad)
        st = stream()
        # residual-stream buffers: all L layer inputs are kept when a backward will follow
        # (the gate is recomputed there, nothing else is saved per layer); otherwise ping-pong.
        nbuf = L if need_grad else min(L, 2)
        xs = torch.empty(nbuf, B, T, R, device=dev, dtype=torch.float32)
        check(lib.wnb_front_embed_fwd(ptr(x), ptr(wf), ptr(bf), ptr(xs[0]), B, T, Q, R, ks, st), "front_embed_fwd")
        skip = torch.empty(B, T, S, device=dev, dtype=torch.float32)
        # shapes outside the fused tcgen05 kernel run the composed tcgen05 path, which needs a z scratch
        zbuf = None
        if math_mode == MATH_TF32 and lib.wnb_resblock_fwd_supported(R, S, Ap, ks, MATH_TF32) == 2:
            zbuf = torch.empty(B, T, R, device=dev, dtype=torch.float32)
        prof = PROFILE_EVENTS
        for l, d in enumerate(dilations):
            xin = xs[l % nbuf]
            xout = xs[(l + 1) % nbuf] if l + 1 < L else None
            if prof is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            check(lib.wnb_resblock_fwd(ptr(xin), ptr(haux), ptr(W1[l]), ptr(b1[l]), ptr(W2[l]), ptr(b2[l]),
                                       ptr(xout), ptr(skip), ptr(zbuf), B, T, R, S, Ap, ks, int(d),
                                       1 if l == 0 else 0, math_mode, st), "resblock_fwd")
            if prof is not None:
                ev[1].record()
                prof.append(ev)
        r1 = torch.empty(B, T, S, device=dev, dtype=torch.float32)
        logits = torch.empty(B, T, Q, device=dev, dtype=torch.float32)
        check(lib.wnb_post_fwd(ptr(skip), ptr(Wp1), ptr(bp1), ptr(Wp2), ptr(bp2), ptr(r1), ptr(logits),
                               B, T, S, Q, math_mode, 0, st), "post_fwd")
        if need_grad:
            ctx.save_for_backward(x, haux, wf, W1, b1, W2, Wp1, Wp2, xs, skip, r1)
            ctx.meta = meta
            ctx.haux_needs_grad = ctx.needs_input_grad[1]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = _lib.load()
        x, haux, wf, W1, b1, W2, Wp1, Wp2, xs, skip, r1 = ctx.saved_tensors
        Q, R, S, Ap, ks, dilations, math_mode = ctx.meta
        B, T = x.shape
        L = len(dilations)
        dev = x.device
        st = stream()
        dlogits = dlogits.contiguous()
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)  # noqa: E731
        K1 = ks * R + Ap
        dWp1, dbp1, dWp2, dbp2 = z(S, S), z(S), z(Q, S), z(Q)
        dskip = torch.empty(B, T, S, device=dev, dtype=torch.float32)
        ws_post = torch.empty(B, T, S, device=dev, dtype=torch.float32)
        wp1t = Wp1.t().contiguous()
        wp2t = Wp2.t().contiguous()
        check(lib.wnb_post_bwd(ptr(skip), ptr(r1), ptr(dlogits), ptr(wp1t), ptr(wp2t), ptr(dskip), ptr(dWp1),
                               ptr(dbp1), ptr(dWp2), ptr(dbp2), ptr(ws_post), B, T, S, Q, math_mode, st), "post_bwd")
        del ws_post
        dW1, db1, dW2, db2 = z(L, 2 * R, K1), z(L, 2 * R), z(L, R + S, R), z(L, R + S)
        dhaux = z(B, T, Ap) if ctx.haux_needs_grad else None
        w1t = W1.transpose(1, 2).contiguous()   # (L, K1, 2R)
        w2t = W2.transpose(1, 2).contiguous()   # (L, R, R+S)
        nbytes = lib.wnb_resblock_bwd_workspace(B, T, R, S, Ap, ks)
        ws = torch.empty(nbytes // 4, device=dev, dtype=torch.float32)
        dbuf = [torch.empty(B, T, R, device=dev, dtype=torch.float32) for _ in range(2)]
        dout = None
        for l in reversed(range(L)):
            dxin = dbuf[l % 2]
            check(lib.wnb_resblock_bwd(ptr(xs[l]), ptr(haux), ptr(dout), ptr(dskip), ptr(W1[l]), ptr(b1[l]),
                                       ptr(w1t[l]), ptr(w2t[l]), ptr(dxin), ptr(dhaux), ptr(dW1[l]), ptr(db1[l]),
                                       ptr(dW2[l]), ptr(db2[l]), ptr(ws), B, T, R, S, Ap, ks, int(dilations[l]),
                                       math_mode, st), "resblock_bwd")
            dout = dxin
        dwf, dbf = z(ks, Q, R), z(R)
        check(lib.wnb_front_embed_bwd(ptr(x), ptr(dout), ptr(dwf), ptr(dbf), B, T, Q, R, ks, st), "front_embed_bwd")
        return None, dhaux, dwf, dbf, dW1, db1, dW2, db2, dWp1, dbp1, dWp2, dbp2, None


class _WaveNetStackFn(torch.autograd.Function):
    """WaveNet.forward (reference wavenet.py:212-241) with the skip path in deferred form (csrc/stack.cu): the blocks
    write their gate outputs z_l into Z_all (B,T,L*R) and ONE GEMM forms skip = Z_all Wskip^T + bskip; the backward
    hoists dZ_all = dskip Wskip and dWskip = dskip^T Z_all out of the block loop the same way.  Same numbers as the
    per-block form (different summation order inside tf32/fp32 accumulation only)."""

    @staticmethod
    def forward(ctx, x, haux, wf, bf, W1, b1, W2res, b2res, Wskip, bskip, Wp1, bp1, Wp2, bp2, meta):
        lib = _lib.load()
        Q, R, S, Ap, ks, dilations, math_mode = meta
        B, T = x.shape
        L = len(dilations)
        dev = x.device
        x = x.contiguous()
        need_grad = any(ctx.needs_input_grad)
        st = stream()
        nbuf = L if need_grad else min(L, 2)
        xs = torch.empty(nbuf, B, T, R, device=dev, dtype=torch.float32)
        check(lib.wnb_front_embed_fwd(ptr(x), ptr(wf), ptr(bf), ptr(xs[0]), B, T, Q, R, ks, st), "front_embed_fwd")
        zall = torch.empty(B, T, L * R, device=dev, dtype=torch.float32)
        skip = torch.empty(B, T, S, device=dev, dtype=torch.float32)
        dil = (ctypes.c_int * L)(*[int(d) for d in dilations])
        prof = PROFILE_EVENTS
        if prof is None:
            check(lib.wnb_stack_fwd(ptr(xs), nbuf, ptr(haux), ptr(W1), ptr(b1), ptr(W2res), ptr(b2res), ptr(Wskip),
                                    ptr(bskip), ptr(zall), ptr(skip), dil, L, B, T, R, S, Ap, ks, 1, st), "stack_fwd")
        else:   # same launches, one ABI call per block so that each can be bracketed by CUDA events
            for l, d in enumerate(dilations):
                xout = xs[(l + 1) % nbuf] if l + 1 < L else None
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
                check(lib.wnb_resblock_fwd_z(ptr(xs[l % nbuf]), ptr(haux), ptr(W1[l]), ptr(b1[l]), ptr(W2res[l]),
                                             ptr(b2res[l]), ptr(xout), ptr(zall), L * R, l * R, B, T, R, Ap, ks,
                                             int(d), st), "resblock_fwd_z")
                ev[1].record()
                prof.append(ev)
            check(lib.wnb_skip_gemm(ptr(zall), ptr(Wskip), ptr(bskip), ptr(skip), B, T, L * R, S, 1, st), "skip_gemm")
        r1 = torch.empty(B, T, S, device=dev, dtype=torch.float32)
        logits = torch.empty(B, T, Q, device=dev, dtype=torch.float32)
        check(lib.wnb_post_fwd(ptr(skip), ptr(Wp1), ptr(bp1), ptr(Wp2), ptr(bp2), ptr(r1), ptr(logits),
                               B, T, S, Q, math_mode, 1, st), "post_fwd")
        if need_grad:
            ctx.save_for_backward(x, haux, wf, W1, b1, W2res, Wskip, Wp1, Wp2, xs, zall, skip, r1)
            ctx.meta = meta
            ctx.haux_needs_grad = ctx.needs_input_grad[1]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = _lib.load()
        x, haux, wf, W1, b1, W2res, Wskip, Wp1, Wp2, xs, zall, skip, r1 = ctx.saved_tensors
        Q, R, S, Ap, ks, dilations, math_mode = ctx.meta
        B, T = x.shape
        L = len(dilations)
        dev = x.device
        st = stream()
        dlogits = dlogits.contiguous()
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)  # noqa: E731
        K1 = ks * R + Ap
        dWp1, dbp1, dWp2, dbp2 = z(S, S), z(S), z(Q, S), z(Q)
        dskip = torch.empty(B, T, S, device=dev, dtype=torch.float32)
        ws_post = torch.empty(B, T, S, device=dev, dtype=torch.float32)
        wp1t = Wp1.t().contiguous()
        wp2t = Wp2.t().contiguous()
        check(lib.wnb_post_bwd(ptr(skip), ptr(r1), ptr(dlogits), ptr(wp1t), ptr(wp2t), ptr(dskip), ptr(dWp1),
                               ptr(dbp1), ptr(dWp2), ptr(dbp2), ptr(ws_post), B, T, S, Q, math_mode, st), "post_bwd")
        del ws_post
        dW1, db1, dW2res, db2res = z(L, 2 * R, K1), z(L, 2 * R), z(L, R, R), z(L, R)
        dWskip, dbskip = z(S, L * R), z(S)
        dhaux = z(B, T, Ap) if ctx.haux_needs_grad else None
        w1t = W1.transpose(1, 2).contiguous()         # (L, K1, 2R)
        wgate = torch.zeros(L, 3 * R, K1 + R, device=dev, dtype=torch.float32)   # [[W1, 0], [0, W2res^T]]
        wgate[:, :2 * R, :K1] = W1
        wgate[:, 2 * R:, K1:] = W2res.transpose(1, 2)
        wskt = Wskip.t().contiguous()                 # (L*R, S)
        nbytes = lib.wnb_stack_bwd_workspace(L, B, T, R, S, Ap, ks)
        ws = torch.empty(nbytes // 4, device=dev, dtype=torch.float32)
        dx0 = torch.empty(B, T, R, device=dev, dtype=torch.float32)
        dil = (ctypes.c_int * L)(*[int(d) for d in dilations])
        check(lib.wnb_stack_bwd(ptr(xs), ptr(haux), ptr(zall), ptr(dskip), ptr(W1), ptr(b1), ptr(w1t), ptr(wgate),
                                ptr(wskt), ptr(dx0), ptr(dhaux), ptr(dW1), ptr(db1), ptr(dW2res), ptr(db2res),
                                ptr(dWskip), ptr(dbskip), ptr(ws), dil, L, B, T, R, S, Ap, ks, st), "stack_bwd")
        del ws
        dwf, dbf = z(ks, Q, R), z(R)
        check(lib.wnb_front_embed_bwd(ptr(x), ptr(dx0), ptr(dwf), ptr(dbf), B, T, Q, R, ks, st), "front_embed_bwd")
        return (None, dhaux, dwf, dbf, dW1, db1, dW2res, db2res, dWskip, dbskip, dWp1, dbp1, dWp2, dbp2, None)


class _StackTrainFn(torch.autograd.Function):
    """WaveNet.forward (reference wavenet.py:212-241) [+ the cross entropy of bin/train.py:534-536] on the deferred-skip
    stack, with NO torch arithmetic in between: parameters are packed by ``wnb_pack_weights`` (one launch), the packed
    gradients are scattered into one flat buffer by the same kernel (one launch) and handed to autograd as views, so
    ``p.grad`` of every parameter aliases one contiguous tensor (``module._wnb_flat_grad``: what the data-parallel
    all-reduce sends).  ``target is None``: returns the (B,T,Q) logits; otherwise the mean CE over ``[:, start:]``."""

    @staticmethod
    def forward(ctx, net, plan, x, h, target, start, *params):
        lib = _lib.load()
        d = plan.dims
        R, S, Q, ks, A, Ap, L, U = d["R"], d["S"], d["Q"], d["ks"], d["A"], d["Ap"], d["L"], d["U"]
        dev = x.device
        st = stream()
        x = x.contiguous()
        h = h.contiguous().float()
        B, T = x.shape
        Tf = h.size(2)
        if h.size(1) != A:
            raise ValueError("aux feature has %d dims, expected %d" % (h.size(1), A))
        if (Tf * U if U > 0 else Tf) != T:
            raise ValueError("aux length %d does not match waveform length %d" % (Tf * max(U, 1), T))
        need_grad = any(ctx.needs_input_grad)
        plan.pack(st)
        P = plan.p
        f32 = dict(device=dev, dtype=torch.float32)
        haux = torch.empty(B, T, Ap, **f32)
        upw = net.upsampling.conv.weight if U > 0 else None
        upb = net.upsampling.conv.bias if U > 0 else None
        check(lib.wnb_aux_upsample_fwd(ptr(h), ptr(upw), ptr(upb), ptr(haux), B, A, Ap, Tf, U, st), "aux_upsample_fwd")
        nbuf = L if need_grad else min(L, 2)
        xs = torch.empty(nbuf, B, T, R, **f32)
        check(lib.wnb_front_embed_fwd(ptr(x), ptr(P("wf")), ptr(P("bf")), ptr(xs[0]), B, T, Q, R, ks, st), "front_embed_fwd")
        zall = torch.empty(B, T, L * R, **f32)
        skip = torch.empty(B, T, S, **f32)
        dil = (ctypes.c_int * L)(*[int(v) for v in net.dilations])
        check(lib.wnb_stack_fwd(ptr(xs), nbuf, ptr(haux), ptr(P("W1")), ptr(P("b1")), ptr(P("W2res")), ptr(P("b2res")),
                                ptr(P("Wskip")), ptr(P("bskip")), ptr(zall), ptr(skip), dil, L, B, T, R, S, Ap, ks, 1, st),
              "stack_fwd")
        r1 = torch.empty(B, T, S, **f32)
        logits = torch.empty(B, T, Q, **f32)
        check(lib.wnb_post_fwd(ptr(skip), ptr(P("Wp1")), ptr(P("bp1")), ptr(P("Wp2")), ptr(P("bp2")), ptr(r1), ptr(logits),
                               B, T, S, Q, MATH_TF32, 1, st), "post_fwd")
        dbg = getattr(net, "_wnb_debug", None)
        if dbg is not None:     # test aid: the ReLU sign patterns of the post network (tests/test_gpu_oracle_direct.py)
            dbg["skip_mask"], dbg["r1_mask"] = skip > 0, r1 > 0
        out = logits
        dl = None
        if target is not None:
            target = target.contiguous()
            if target.dtype != torch.int64 or tuple(target.shape) != (B, T) or target.device != dev:
                raise ValueError("target should be an int64 (B, T) tensor on the model's device")
            loss = torch.empty(1, device=dev, dtype=torch.float64)
            check(lib.wnb_zero(ptr(loss), 8, st), "zero")
            dl = logits if need_grad else None     # dlogits overwrite the logits in place (each row is read, then written)
            check(lib.wnb_cross_entropy(ptr(logits), ptr(target), ptr(loss), ptr(dl), B, T, Q, int(start), st),
                  "cross_entropy")
            out = loss.float().view(())
        if need_grad:
            ctx.save_for_backward(x, h, haux, xs, zall, skip, r1)
            ctx.net, ctx.plan, ctx.dl, ctx.dil = net, plan, dl, dil
            ctx.fused_loss = target is not None
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, h, haux, xs, zall, skip, r1 = ctx.saved_tensors
        net, plan, dil = ctx.net, ctx.plan, ctx.dil
        d = plan.dims
        R, S, Q, ks, A, Ap, L, U = d["R"], d["S"], d["Q"], d["ks"], d["A"], d["Ap"], d["L"], d["U"]
        B, T = x.shape
        Tf = h.size(2)
        dev = x.device
        st = stream()
        f32 = dict(device=dev, dtype=torch.float32)
        if ctx.fused_loss:
            dlogits, scale = ctx.dl, g.contiguous().float()   # d loss / d logits from the CE kernel; g scales the result
            ctx.dl = None
        else:
            dlogits, scale = g.contiguous(), None
        P = plan.p
        gbuf = torch.empty(plan.G.size, **f32)
        check(lib.wnb_zero(ptr(gbuf), gbuf.numel() * 4, st), "zero")
        G = lambda name: plan.G.view(gbuf, name)  # noqa: E731
        dskip = torch.empty(B, T, S, **f32)
        ws_post = torch.empty(B, T, S, **f32)
        check(lib.wnb_post_bwd(ptr(skip), ptr(r1), ptr(dlogits), ptr(P("wp1t")), ptr(P("wp2t")), ptr(dskip), ptr(G("Wp1")),
                               ptr(G("bp1")), ptr(G("Wp2")), ptr(G("bp2")), ptr(ws_post), B, T, S, Q, MATH_TF32, st),
              "post_bwd")
        del ws_post, dlogits
        dhaux = None
        if U > 0:
            dhaux = torch.empty(B, T, Ap, **f32)
            check(lib.wnb_zero(ptr(dhaux), dhaux.numel() * 4, st), "zero")
        nbytes = lib.wnb_stack_bwd_workspace(L, B, T, R, S, Ap, ks)
        ws = torch.empty(nbytes // 4, **f32)
        dx0 = torch.empty(B, T, R, **f32)
        check(lib.wnb_stack_bwd(ptr(xs), ptr(haux), ptr(zall), ptr(dskip), ptr(P("W1")), ptr(P("b1")), ptr(P("w1t")),
                                ptr(P("wgate")), ptr(P("wskt")), ptr(dx0), ptr(dhaux), ptr(G("W1")), ptr(G("b1")),
                                ptr(G("W2res")), ptr(G("b2res")), ptr(G("Wskip")), ptr(G("bskip")), ptr(ws), dil, L, B, T,
                                R, S, Ap, ks, st), "stack_bwd")
        del ws
        check(lib.wnb_front_embed_bwd(ptr(x), ptr(dx0), ptr(G("wf")), ptr(G("bf")), B, T, Q, R, ks, st), "front_embed_bwd")
        if U > 0:
            check(lib.wnb_aux_upsample_bwd(ptr(h), ptr(dhaux), ptr(G("upw")), ptr(G("upb")), B, A, Ap, Tf, U, st),
                  "aux_upsample_bwd")
        flat = torch.empty(plan.grad_size, **f32)
        plan.unpack(gbuf, flat, scale, st)
        net._wnb_flat_grad = flat
        return (None, None, None, None, None, None) + tuple(plan.grad_views(flat))


# ------------------------------------------------------------------------------------------
# modules (same names / parameters / state_dict keys as the reference)
# ------------------------------------------------------------------------------------------
class OneHot(nn.Module):
    """CONVERT TO ONE-HOT VECTOR (reference wavenet.py:66-92).

    Kept for API compatibility only: ``WaveNet`` never materialises the one-hot tensor (the front
    convolution is an embedding gather on the device)."""

    def __init__(self, depth):
        super(OneHot, self).__init__()
        self.depth = depth

    def forward(self, x):
        x = x % self.depth
        x = torch.unsqueeze(x, 2)
        x_onehot = x.new_zeros(x.size(0), x.size(1), self.depth).float()
        return x_onehot.scatter_(2, x, 1)


class CausalConv1d(nn.Module):
    """1D DILATED CAUSAL CONVOLUTION (reference wavenet.py:95-121): parameter container.

    Inside ``WaveNet`` these weights are consumed by the fused block kernel; the module keeps the
    reference's ``.conv.weight`` / ``.conv.bias`` parameter names so checkpoints interchange."""

    def __init__(self, in_channels, out_channels, kernel_size, dilation=1, bias=True):
        super(CausalConv1d, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.dilation = dilation
        self.padding = padding = (kernel_size - 1) * dilation
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size,
                              padding=padding, dilation=dilation, bias=bias)

    def forward(self, x):
        """FORWARD CALCULATION (reference wavenet.py:108-121): x (B, C, T) -> (B, C_out, T), inference only.

        Inside ``WaveNet`` the convolution runs in the fused block kernels; this standalone form exists for API
        compatibility and goes through ``wnb_causal_conv1d_fwd`` (no autograd)."""
        if torch.is_grad_enabled() and (x.requires_grad or self.conv.weight.requires_grad):
            # the reference module is a differentiable Conv1d; this standalone helper has no backward kernel, and a
            # silently detached result would train nothing -- say so instead
            raise RuntimeError("CausalConv1d.forward is inference-only in the B200 build (no standalone backward): "
                               "call it under torch.no_grad(); training goes through WaveNet.forward")
        lib = _lib.load()
        if not x.is_cuda:
            raise _lib.WnbError("CausalConv1d.forward needs CUDA tensors: the B200 build has no CPU fallback")
        B, C, T = x.shape
        xl = x.detach().float().transpose(1, 2).contiguous()                                   # (B, T, C)
        w = self.conv.weight.detach().float().permute(0, 2, 1).reshape(self.out_channels, -1).contiguous()  # [o][j*C+c]
        b = None if self.conv.bias is None else self.conv.bias.detach().float().contiguous()
        out = torch.empty(B, T, self.out_channels, device=x.device, dtype=torch.float32)
        check(lib.wnb_causal_conv1d_fwd(ptr(xl), ptr(w), ptr(b), ptr(out), B, T, C, self.out_channels,
                                        self.kernel_size, self.dilation, stream()), "causal_conv1d_fwd")
        return out.transpose(1, 2)


class UpSampling(nn.Module):
    """UPSAMPLING LAYER WITH DECONVOLUTION (reference wavenet.py:124-154)."""

    def __init__(self, upsampling_factor, bias=True):
        super(UpSampling, self).__init__()
        self.upsampling_factor = upsampling_factor
        self.bias = bias
        self.conv = nn.ConvTranspose2d(1, 1,
                                       kernel_size=(1, self.upsampling_factor),
                                       stride=(1, self.upsampling_factor),
                                       bias=self.bias)

    def forward(self, x):
        """x (B, C, T) -> (B, C, T * upsampling_factor)."""
        C = x.size(1)
        b = self.conv.bias if self.bias else torch.zeros(1, device=x.device)
        out = _UpsampleFn.apply(x, self.conv.weight, b, _round_up(C, 32))
        return out[:, :, :C].transpose(1, 2)


def _round_up(a, m):
    return (a + m - 1) // m * m


class WaveNet(nn.Module):
    """CONDITIONAL WAVENET (reference wavenet.py:157-549).

    Args:
        n_quantize (int): Number of quantization.
        n_aux (int): Number of aux feature dimension.
        n_resch (int): Number of filter channels for residual block.
        n_skipch (int): Number of filter channels for skip connection.
        dilation_depth (int): Number of dilation depth (e.g. if set 10, max dilation = 2^(10-1)).
        dilation_repeat (int): Number of dilation repeat.
        kernel_size (int): Filter size of dilated causal convolution.
        upsampling_factor (int): Upsampling factor.
    """

    def __init__(self, n_quantize=256, n_aux=28, n_resch=512, n_skipch=256,
                 dilation_depth=10, dilation_repeat=3, kernel_size=2, upsampling_factor=0):
        super(WaveNet, self).__init__()
        self.n_aux = n_aux
        self.n_quantize = n_quantize
        self.n_resch = n_resch
        self.n_skipch = n_skipch
        self.kernel_size = kernel_size
        self.dilation_depth = dilation_depth
        self.dilation_repeat = dilation_repeat
        self.upsampling_factor = upsampling_factor

        self.dilations = [2 ** i for i in range(self.dilation_depth)] * self.dilation_repeat
        self.receptive_field = (self.kernel_size - 1) * sum(self.dilations) + 1

        # B200 build: contraction precision of the training forward. "fp32" = FFMA parity path,
        # "tf32" = tcgen05 tensor-core path (fp32 storage and accumulation). Decode is always fp32.
        self.math_mode = "fp32"

        # for preprocessing
        self.onehot = OneHot(self.n_quantize)
        self.causal = CausalConv1d(self.n_quantize, self.n_resch, self.kernel_size)
        if self.upsampling_factor > 0:
            self.upsampling = UpSampling(self.upsampling_factor)

        # for residual blocks
        self.dil_sigmoid = nn.ModuleList()
        self.dil_tanh = nn.ModuleList()
        self.aux_1x1_sigmoid = nn.ModuleList()
        self.aux_1x1_tanh = nn.ModuleList()
        self.skip_1x1 = nn.ModuleList()
        self.res_1x1 = nn.ModuleList()
        for d in self.dilations:
            self.dil_sigmoid += [CausalConv1d(self.n_resch, self.n_resch, self.kernel_size, d)]
            self.dil_tanh += [CausalConv1d(self.n_resch, self.n_resch, self.kernel_size, d)]
            self.aux_1x1_sigmoid += [nn.Conv1d(self.n_aux, self.n_resch, 1)]
            self.aux_1x1_tanh += [nn.Conv1d(self.n_aux, self.n_resch, 1)]
            self.skip_1x1 += [nn.Conv1d(self.n_resch, self.n_skipch, 1)]
            self.res_1x1 += [nn.Conv1d(self.n_resch, self.n_resch, 1)]

        # for postprocessing
        self.conv_post_1 = nn.Conv1d(self.n_skipch, self.n_skipch, 1)
        self.conv_post_2 = nn.Conv1d(self.n_skipch, self.n_quantize, 1)

    # -------------------------------------------------------------------------------------
    # weight packing (include/wnb200.h "Packed weight layouts"); differentiable torch ops so the
    # gradients of the packed matrices flow back to the reference-shaped parameters
    # -------------------------------------------------------------------------------------
    @property
    def n_aux_pad(self):
        return _round_up(self.n_aux, 32)

    def _math(self):
        if self.math_mode == "fp32":
            return MATH_FP32
        if self.math_mode == "tf32":
            return MATH_TF32
        raise ValueError("math_mode should be fp32 or tf32")

    def _pack(self):
        R, A, Ap, ks = self.n_resch, self.n_aux, self.n_aux_pad, self.kernel_size
        L = len(self.dilations)
        st = torch.stack
        wd = torch.cat([st([m.conv.weight for m in self.dil_sigmoid]),
                        st([m.conv.weight for m in self.dil_tanh])], 1)            # (L, 2R, R, ks)
        wd = wd.permute(0, 1, 3, 2).reshape(L, 2 * R, ks * R)                       # [o][j*R + c]
        wa = torch.cat([st([m.weight for m in self.aux_1x1_sigmoid])[..., 0],
                        st([m.weight for m in self.aux_1x1_tanh])[..., 0]], 1)      # (L, 2R, A)
        wa = F.pad(wa, (0, Ap - A))
        W1 = torch.cat([wd, wa], 2).contiguous()                                    # (L, 2R, K1)
        b1 = torch.cat([st([m.conv.bias for m in self.dil_sigmoid]) + st([m.bias for m in self.aux_1x1_sigmoid]),
                        st([m.conv.bias for m in self.dil_tanh]) + st([m.bias for m in self.aux_1x1_tanh])],
                       1).contiguous()                                              # (L, 2R)
        # the last block's residual output is discarded (reference wavenet.py:230-238): its res_1x1
        # gets no gradient there, so keep it out of the graph here as well
        res_w = [m.weight if l + 1 < L else m.weight.detach() for l, m in enumerate(self.res_1x1)]
        res_b = [m.bias if l + 1 < L else m.bias.detach() for l, m in enumerate(self.res_1x1)]
        W2 = torch.cat([st(res_w)[..., 0], st([m.weight for m in self.skip_1x1])[..., 0]], 1).contiguous()
        b2 = torch.cat([st(res_b), st([m.bias for m in self.skip_1x1])], 1).contiguous()
        wf = self.causal.conv.weight.permute(2, 1, 0).contiguous()                  # (ks, Q, R)
        bf = self.causal.conv.bias
        Wp1 = self.conv_post_1.weight[..., 0].contiguous()
        Wp2 = self.conv_post_2.weight[..., 0].contiguous()
        return wf, bf, W1, b1, W2, b2, Wp1, self.conv_post_1.bias, Wp2, self.conv_post_2.bias

    def _pack_stack(self, packed):
        """Operands of the deferred-skip stack (csrc/stack.cu) from ``_pack()``'s: W2 / b2 split into the residual part
        ``W2res (L,R,R)``, ``b2res (L,R)`` and the skip GEMM operand over the concatenated gate outputs
        ``Wskip (S, L*R)`` with ``Wskip[s][l*R + c] = skip_1x1[l].weight[s][c]``, ``bskip (S) = sum_l skip_1x1[l].bias``."""
        wf, bf, W1, b1, W2, b2, Wp1, bp1, Wp2, bp2 = packed
        R, S, L = self.n_resch, self.n_skipch, len(self.dilations)
        W2res, b2res = W2[:, :R].contiguous(), b2[:, :R].contiguous()
        Wskip = W2[:, R:].permute(1, 0, 2).reshape(S, L * R).contiguous()
        bskip = b2[:, R:].sum(0)
        return wf, bf, W1, b1, W2res, b2res, Wskip, bskip, Wp1, bp1, Wp2, bp2

    def _aux(self, h, upsample=True):
        """(B, A, T or T/U) -> channels-last (B, T, Ap) with the up-sampling layer applied."""
        if h.size(1) != self.n_aux:
            raise ValueError("aux feature has %d dims, expected %d" % (h.size(1), self.n_aux))
        if self.upsampling_factor > 0 and upsample:
            return _UpsampleFn.apply(h, self.upsampling.conv.weight, self.upsampling.conv.bias, self.n_aux_pad)
        return _UpsampleFn.apply(h, None, None, self.n_aux_pad)

    def _stack_plan(self, device):
        """Copy tables + packed buffers of the one-launch weight packing (nets/packing.py); None when the deferred-skip
        stack does not cover this model."""
        if not getattr(self, "deferred_skip", True) or self.math_mode != "tf32":
            return None
        if not _lib.load().wnb_stack_supported(self.n_resch, self.n_skipch, self.n_aux_pad, self.kernel_size,
                                               len(self.dilations), MATH_TF32):
            return None
        plan = self.__dict__.get("_wnb_plan")
        if plan is None or plan.device != device or plan.stale():
            from .packing import StackPlan
            plan = StackPlan(self, device)
            self.__dict__["_wnb_plan"] = plan
        return plan

    def forward_loss(self, x, h, t, start=None):
        """Training entry of the B200 build: mean cross entropy of ``forward(x, h)[:, start:]`` against ``t[:, start:]``
        (reference bin/train.py:533-536 with ``start = receptive_field``) as ONE autograd node -- the post network's
        logits never make a round trip through a separate loss kernel's gradient multiply.  Same value and gradients as
        ``cross_entropy(self(x, h), t, start)``."""
        start = self.receptive_field if start is None else int(start)
        if not x.is_cuda:
            raise _lib.WnbError("WaveNet.forward_loss needs CUDA tensors: the B200 build has no CPU fallback")
        plan = self._stack_plan(x.device)
        if plan is None:
            return cross_entropy(self._forward_impl(x, h), t, start)
        return _StackTrainFn.apply(self, plan, x, h, t, start, *[p for _, p in plan.params])

    def _forward_impl(self, x, h, upsample=True):
        if not x.is_cuda:
            raise _lib.WnbError("WaveNet.forward needs CUDA tensors: the B200 build has no CPU fallback")
        if upsample and PROFILE_EVENTS is None:
            plan = self._stack_plan(x.device)
            if plan is not None:
                return _StackTrainFn.apply(self, plan, x, h, None, 0, *[p for _, p in plan.params])
        haux = self._aux(h, upsample)
        if haux.size(1) != x.size(1):
            raise ValueError("aux length %d does not match waveform length %d" % (haux.size(1), x.size(1)))
        packed = self._pack()
        meta = (self.n_quantize, self.n_resch, self.n_skipch, self.n_aux_pad, self.kernel_size,
                tuple(self.dilations), self._math())
        if getattr(self, "deferred_skip", True) and _lib.load().wnb_stack_supported(self.n_resch, self.n_skipch, self.n_aux_pad, self.kernel_size,
                                           len(self.dilations), self._math()):
            return _WaveNetStackFn.apply(x, haux, *self._pack_stack(packed), meta)
        return _WaveNetFn.apply(x, haux, *packed, meta)

    def forward(self, x, h):
        """FORWARD CALCULATION (reference wavenet.py:212-241).

        Args:
            x (Tensor): Long tensor variable with the shape (B, T).
            h (Tensor): Float tensor variable with the shape (B, n_aux, T)
                (or (B, n_aux, T / upsampling_factor) with the upsampling layer).

        Returns:
            Tensor: Float tensor variable with the shape (B, T, n_quantize).
        """
        return self._forward_impl(x, h)

    # -------------------------------------------------------------------------------------
    # generation
    # -------------------------------------------------------------------------------------
    def generate(self, x, h, n_samples, intervals=None, mode="sampling"):
        """GENERATE WAVEFORM WITH NAIVE CALCULATION (reference wavenet.py:243-307).

        O(receptive_field) work per sample through the training forward kernels; kept, like in the
        reference, as the ground truth the fast algorithm is tested against."""
        if mode not in ("sampling", "argmax"):
            logging.error("mode should be sampling or argmax")
            sys.exit(1)
        with torch.no_grad():
            if self.upsampling_factor > 0:
                h = self.upsampling(h)
            n_pad = self.receptive_field - x.size(1)
            if n_pad > 0:
                x = F.pad(x, (n_pad, 0), "constant", self.n_quantize // 2)
                h = F.pad(h, (n_pad, 0), "replicate")
            samples = x[0].tolist()
            start = time.time()
            for i in range(n_samples):
                current_idx = len(samples)
                xx = torch.tensor(samples[-self.receptive_field:], device=h.device).long().view(1, -1)
                h_ = h[:, :, current_idx - self.receptive_field: current_idx].contiguous()
                output = self._forward_impl(xx, h_, upsample=False)[0]
                if mode == "sampling":
                    posterior = F.softmax(output[-1], dim=0)
                    sample = int(torch.distributions.Categorical(posterior).sample())
                else:
                    sample = int(output[-1].argmax())
                samples.append(sample)
                if intervals is not None and (i + 1) % intervals == 0:
                    logging.info("%d/%d estimated time = %.3f sec (%.3f sec / sample)" % (
                        i + 1, n_samples,
                        (n_samples - i - 1) * ((time.time() - start) / intervals),
                        (time.time() - start) / intervals))
                    start = time.time()
        return np.array(samples[-n_samples:])

    def _decode_pack(self):
        def pad4(t):  # pad the output (last) dim of a K-major matrix to a multiple of 4
            o = t.size(-1)
            return F.pad(t, (0, _round_up(o, 4) - o)).contiguous()
        with torch.no_grad():
            wf, bf, W1, b1, W2, b2, Wp1, bp1, Wp2, bp2 = [t.detach().float() for t in self._pack()]
            R = self.n_resch
            return dict(wf=wf.contiguous(), bf=bf.contiguous(), w1d=pad4(W1.transpose(1, 2)), b1=b1,
                        w2d=pad4(W2.transpose(1, 2)), b2=b2, wp1d=pad4(Wp1.t()), bp1=bp1.contiguous(),
                        wp2d=pad4(Wp2.t()), bp2=bp2.contiguous(),
                        w2d_res=pad4(W2[:, :R].transpose(1, 2)), w2d_skip=pad4(W2[:, R:].transpose(1, 2)))

    def _decode_stream_pack(self, w):
        """One contiguous fp32 stream in consumption order (include/wnb200.h, wnb_decode_stream)."""
        R = self.n_resch
        parts = []
        for l in range(len(self.dilations)):
            parts += [w["w1d"][l].reshape(-1), w["w2d_res"][l].reshape(-1), w["w2d_skip"][l].reshape(-1)]
        parts += [w["wp1d"].reshape(-1), w["wp2d"].reshape(-1)]
        return torch.cat(parts).contiguous()

    def _decode_warp_pack(self, W, CL=1):
        """Warp-tile ordered stream for wnb_decode_warp with W (virtual) consumer warps (layout: csrc/decode_warp.cu).
        CL = 2: one stream per CTA of the cluster, back to back -- CTA r gets the tiles of warps r*W/2 .. (r+1)*W/2 - 1
        (= gate / residual channels, skip / post columns of its half) and the full bias vectors."""
        with torch.no_grad():
            wf, bf, W1, b1, W2, b2, Wp1, bp1, Wp2, bp2 = [t.detach().float() for t in self._pack()]
            L = W1.size(0)
            CH = 64 // W                     # gate / residual channels per warp
            WP = W // CL
            # W1 (L,128,160)[o][k] -> [j][w][g][lane][4]: k = 32j+lane; a lane's 2*CH values (index 4g+e) are
            # (sigmoid rows CH*w.., tanh rows CH*w..); consecutive lanes are 16 B apart (conflict-free LDS.128)
            t1 = W1.reshape(L, 2, W, CH, 5, 32).permute(0, 4, 2, 1, 3, 5)          # [L][j][w][br][cc][lane]
            t1 = t1.reshape(L, 5, W, 2 * CH // 4, 4, 32).permute(0, 1, 2, 3, 5, 4)  # [L][j][w][g][lane][4]
            # W2 res rows (L,64,64)[o][k] -> [j][w][g][lane][4]: k = 32j+lane, o = CH*w + 4g + e
            tr = W2[:, :64, :].reshape(L, W, CH, 2, 32).permute(0, 3, 1, 2, 4)     # [L][j][w][cc][lane]
            tr = tr.reshape(L, 2, W, CH // 4, 4, 32).permute(0, 1, 2, 3, 5, 4)      # [L][j][w][g][lane][4]
            # W2 skip rows (L,512,64)[o][k] -> [k][512]
            ts = W2[:, 64:, :].transpose(1, 2)                                      # [L][64 k][512]
            p1, p2 = Wp1.t(), Wp2.t()                                              # [512 k][512], [512 k][256]
            S, Q = p1.size(1), p2.size(1)
            parts = []
            for r in range(CL):
                ws = slice(r * WP, (r + 1) * WP)
                # biases ride in the stream right behind their matrices: [W1 | b1 | W2res | b2 | W2skip] per layer,
                # then [bp1 | bp2 | Wp1^T | Wp2^T]
                per_layer = torch.cat([t1[:, :, ws].reshape(L, -1), b1, tr[:, :, ws].reshape(L, -1), b2,
                                       ts[:, :, r * S // CL:(r + 1) * S // CL].reshape(L, -1)], 1).reshape(-1)
                parts += [per_layer, bp1, bp2, p1[:, r * S // CL:(r + 1) * S // CL].reshape(-1),
                          p2[:, r * Q // CL:(r + 1) * Q // CL].reshape(-1)]
            return torch.cat(parts).contiguous()

    def _decode(self, x, h, n_samples_list, mode, uniforms=None, return_logits=False, seed=None, kernel="auto"):
        lib = _lib.load()
        if mode not in ("sampling", "argmax"):
            logging.error("mode should be sampling or argmax")
            sys.exit(1)
        if not h.is_cuda:
            raise _lib.WnbError("generation needs CUDA tensors: the B200 build has no CPU fallback")
        dev = h.device
        B, T0 = x.shape
        n_list = [int(n) for n in n_samples_list]
        max_n = max(n_list)
        Q, A, Ap, R, S, ks, U = (self.n_quantize, self.n_aux, self.n_aux_pad, self.n_resch, self.n_skipch,
                                 self.kernel_size, self.upsampling_factor)
        P = max(T0, self.receptive_field)
        n_pad = P - T0
        xs = torch.full((B, P + max_n), Q // 2, dtype=torch.int32, device=dev)
        xs[:, n_pad:P] = x.to(dev).to(torch.int32)
        h = h.contiguous().float()
        Th = h.size(2)
        if Th * max(U, 1) < max_n + T0 - 1:
            raise ValueError("aux features too short: %d frames for %d samples" % (Th, max_n + T0))
        w = self._decode_pack()
        dil = (ctypes.c_int32 * len(self.dilations))(*self.dilations)
        L = len(self.dilations)
        qbytes = lib.wnb_decode_workspace(B, R, ks, dil, L)
        queues = torch.empty(max(qbytes // 4, 1), dtype=torch.float32, device=dev)
        nsm = torch.tensor(n_list, dtype=torch.int32, device=dev)
        uni = None if uniforms is None else uniforms.to(dev).float().contiguous()
        lg = torch.empty(B, max_n, Q, dtype=torch.float32, device=dev) if return_logits else None
        if seed is None:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())  # follows torch.manual_seed
        upw = self.upsampling.conv.weight.detach().float().contiguous().view(-1) if U > 0 else None
        upb = self.upsampling.conv.bias.detach().float().contiguous().view(-1) if U > 0 else None
        cmode = MODE_ARGMAX if mode == "argmax" else MODE_SAMPLING
        cseed = ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF)
        rc = -3
        self.last_decode_kernel = None
        if kernel in ("auto", "warp") and lib.wnb_decode_warp_supported(Q, Ap, R, S, ks, L):
            Wc, CLc = lib.wnb_decode_warp_plan(B), lib.wnb_decode_warp_cluster(B)
            wstream = self._decode_warp_pack(Wc, CLc)
            assert wstream.numel() == lib.wnb_decode_warp_floats(L, CLc)
            rc = lib.wnb_decode_warp(ptr(xs), ptr(h), ptr(upw), ptr(upb), ptr(w["wf"]), ptr(w["bf"]), ptr(wstream),
                                     ptr(w["b1"]), ptr(w["b2"]), ptr(w["bp1"]), ptr(w["bp2"]), dil, L, ptr(queues),
                                     ptr(nsm), ptr(uni), ptr(lg), B, P, max_n, n_pad, Th, A, U, cmode, cseed, Wc,
                                     CLc, stream())
            if rc != -3 or kernel == "warp":
                check(rc, "decode_warp")
            if rc == 0:
                self.last_decode_kernel = "warp"
        elif kernel == "warp":
            raise _lib.WnbError("decode kernel 'warp' does not cover this configuration")
        if rc == -3 and kernel in ("auto", "stream"):
            wstream = self._decode_stream_pack(w)
            assert wstream.numel() == lib.wnb_decode_stream_floats(Q, Ap, R, S, ks, L)
            rc = lib.wnb_decode_stream(ptr(xs), ptr(h), ptr(upw), ptr(upb), ptr(w["wf"]), ptr(w["bf"]), ptr(wstream),
                                       ptr(w["b1"]), ptr(w["b2"]), ptr(w["bp1"]), ptr(w["bp2"]), dil, L, ptr(queues),
                                       ptr(nsm), ptr(uni), ptr(lg), B, P, max_n, n_pad, Th, Q, A, Ap, R, S, ks, U,
                                       cmode, cseed, stream())
            if rc != -3 or kernel == "stream":
                check(rc, "decode_stream")
        if rc == -3:   # WNB_ERR_UNSUPPORTED: shape outside the streaming kernel -> direct-from-L2 kernel
            check(lib.wnb_decode(ptr(xs), ptr(h), ptr(upw), ptr(upb), ptr(w["wf"]), ptr(w["bf"]), ptr(w["w1d"]),
                                 ptr(w["b1"]), ptr(w["w2d"]), ptr(w["b2"]), ptr(w["wp1d"]), ptr(w["bp1"]),
                                 ptr(w["wp2d"]), ptr(w["bp2"]), dil, L, ptr(queues), ptr(nsm), ptr(uni), ptr(lg),
                                 B, P, max_n, n_pad, Th, Q, A, Ap, R, S, ks, U, cmode, cseed, stream()), "decode")
        if self.last_decode_kernel is None:
            self.last_decode_kernel = "stream" if rc == 0 else "direct"
        gen = xs[:, P:]
        if return_logits:
            return gen, lg
        return gen

    def fast_generate(self, x, h, n_samples, intervals=None, mode="sampling"):
        """GENERATE WAVEFORM WITH FAST ALGORITHM (reference wavenet.py:309-395).

        One persistent kernel launch; ``intervals`` only controls the summary log line."""
        start = time.time()
        gen = self._decode(x[:1], h[:1], [n_samples], mode)
        out = gen[0, :n_samples].cpu().numpy().astype(np.int64)
        if intervals is not None:
            el = time.time() - start
            logging.info("%d/%d generated in %.3f sec (%.6f sec / sample)" % (n_samples, n_samples, el,
                                                                              el / max(n_samples, 1)))
        return out

    def batch_fast_generate_pcm16(self, x, h, n_samples_list, intervals=None, mode="sampling"):
        """``batch_fast_generate`` + ``decode_mu_law`` + PCM_16 quantisation (reference bin/decode.py:316-319) without
        leaving the device in between: one decode launch, one gather kernel for the whole batch, 2 bytes per sample over
        PCIe.  Returns int16 arrays in COMPLETION order like ``batch_fast_generate``."""
        n_list = [int(n) for n in n_samples_list]
        pcm = codes_to_pcm16(self._decode(x, h, n_list, mode), self.n_quantize).cpu().numpy()
        order = sorted(range(len(n_list)), key=lambda b: (n_list[b], b))
        return [pcm[b, :n_list[b]].copy() for b in order]

    def batch_fast_generate(self, x, h, n_samples_list, intervals=None, mode="sampling"):
        """GENERATE WAVEFORM WITH FAST ALGORITHM IN BATCH MODE (reference wavenet.py:397-511).

        Returns the utterances in COMPLETION order (ascending length, ties by batch index) exactly
        like the reference's retirement loop (:487-509).  The caller's list is not mutated."""
        start = time.time()
        n_list = [int(n) for n in n_samples_list]
        gen = self._decode(x, h, n_list, mode).cpu().numpy().astype(np.int64)
        order = sorted(range(len(n_list)), key=lambda b: (n_list[b], b))
        outs = [gen[b, :n_list[b]].copy() for b in order]
        if intervals is not None:
            el = time.time() - start
            logging.info("%d utterances, %d samples generated in %.3f sec (%.6f sec / sample step)" % (
                len(n_list), sum(n_list), el, el / max(max(n_list), 1)))
        return outs
