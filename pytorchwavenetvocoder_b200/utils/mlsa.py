# -*- coding: utf-8 -*-
"""MLSA noise shaping on the GPU: host side of ``wnb_mlsa_filter``.

Mirrors what the reference does through pysptk (reference bin/noise_shaping.py:28-43, 46-87): ``mc2b`` turns the scaled
average mel-cepstrum into MLSA filter coefficients, and every wav file is passed through the time-invariant MLSA filter
(``Synthesizer(MLSADF(order, alpha), hopsize).synthesis`` with one coefficient row tiled over all frames).  Here a whole
list of utterances is filtered by ONE kernel launch (csrc/mlsa.cu); there is no CPU fallback."""
import numpy as np
import torch

from .._lib import check, load, ptr, stream


def mc2b(mc, alpha):
    """SPTK ``mc2b`` (``pysptk.mc2b``, reference noise_shaping.py:41): b[m] = mc[m]; b[i] = mc[i] - alpha * b[i+1]."""
    mc = np.asarray(mc, dtype=np.float64)
    b = np.empty_like(mc)
    b[-1] = mc[-1]
    for i in range(len(mc) - 2, -1, -1):
        b[i] = mc[i] - alpha * b[i + 1]
    return b


def convert_mcep_to_mlsa_coef(avg_mcep, mag, alpha):
    """CONVERT AVERAGE MEL-CEPTSRUM TO MLSA FILTER COEFFICIENT (reference noise_shaping.py:28-43; like the reference the
    argument is scaled IN PLACE when it is a float64 array)."""
    avg_mcep *= mag
    avg_mcep[0] = 0.0
    coef = mc2b(avg_mcep.astype(np.float64), alpha)
    assert np.isfinite(coef).all()
    return coef


def mlsa_filter_batch(signals, coef, alpha, pd=4, out_int16=None, device=None):
    """Filter a list of 1-D signals with the time-invariant MLSA filter ``coef`` (order + 1 values from ``mc2b``; pass
    ``-coef`` for the inverse filter, reference noise_shaping.py:55-56).

    signals: list of int16 arrays (wav samples; converted like ``np.float64(x)``) or float64 arrays, all of one dtype.
    Returns a list of arrays of the same lengths: int16 (``np.int16(y)``, truncation toward zero, the reference's wav
    output) when the input is int16 or ``out_int16`` is True, float64 otherwise.  One launch for the whole list; every
    utterance starts from a zero filter state."""
    if len(signals) == 0:
        return []
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    is_i16 = all(np.asarray(s).dtype == np.int16 for s in signals)
    if out_int16 is None:
        out_int16 = is_i16
    dt = np.int16 if is_i16 else np.float64
    lens = [len(s) for s in signals]
    off = np.zeros(len(signals) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    flat = np.concatenate([np.asarray(s, dtype=dt) for s in signals]) if off[-1] > 0 else np.zeros(0, dtype=dt)
    if off[-1] == 0:
        return [np.zeros(0, dtype=np.int16 if out_int16 else np.float64) for _ in signals]
    with torch.cuda.device(dev):
        x = torch.from_numpy(flat).to(dev)
        offsets = torch.from_numpy(off).to(dev)
        b = torch.from_numpy(coef).to(dev)
        y = torch.empty(int(off[-1]), dtype=torch.int16 if out_int16 else torch.float64, device=dev)
        check(load().wnb_mlsa_filter(ptr(x), 1 if is_i16 else 0, ptr(offsets), len(signals), ptr(b), len(coef) - 1,
                                     float(alpha), int(pd), float(np.exp(coef[0])), ptr(y), 1 if out_int16 else 0, stream()),
              "mlsa_filter")
        yh = y.cpu().numpy()
    return [yh[off[i]:off[i + 1]] for i in range(len(signals))]
