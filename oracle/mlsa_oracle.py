# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE (oracle): ctypes front end of oracle/mlsa_oracle.c, the CPU restatement of the reference's
pysptk-based MLSA noise shaping (reference bin/noise_shaping.py:28-43, :57-87).  See the header of the C file: parity
is UNPINNED against pysptk (absent from this image); the restatement is pinned by the filter's analytic frequency
response and inverse-filter round trip (tests/test_mlsa_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU leg may import this module."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mlsa_oracle.c")
LIB = os.path.join(HERE, "_build", "libmlsa_oracle.so")
_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off (no FMA contraction: the operation order is the contract) -> oracle/_build/"""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        d, lg, i = ctypes.POINTER(ctypes.c_double), ctypes.c_long, ctypes.c_int
        L.mlsa_mc2b.argtypes = [d, d, i, ctypes.c_double]
        L.mlsa_delay_len.argtypes = [i, i]
        L.mlsa_delay_len.restype = i
        L.mlsa_synthesis.argtypes = [d, lg, d, lg, i, ctypes.c_double, i, i, d, d, d]
        L.mlsa_synthesis.restype = lg
        L.mlsa_filter_const.argtypes = [d, lg, d, i, ctypes.c_double, i, ctypes.c_double, d, d]
        L.mlsa_to_int16.argtypes = [d, lg, ctypes.POINTER(ctypes.c_int16)]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def mc2b(mc, alpha):
    """SPTK mc2b (pysptk.mc2b, reference noise_shaping.py:41)."""
    mc = np.ascontiguousarray(mc, dtype=np.float64)
    b = np.empty_like(mc)
    lib().mlsa_mc2b(_p(mc), _p(b), len(mc) - 1, float(alpha))
    return b


def convert_mcep_to_mlsa_coef(avg_mcep, mag, alpha):
    """reference noise_shaping.py:28-43 (does not modify its argument, unlike the reference)."""
    mc = np.array(avg_mcep, dtype=np.float64) * mag
    mc[0] = 0.0
    coef = mc2b(mc, alpha)
    assert np.isfinite(coef).all()
    return coef


def new_delay(order, pd=4):
    return np.zeros(lib().mlsa_delay_len(int(order), int(pd)), dtype=np.float64)


def filter_const(x, coef, alpha, pd=4, delay=None):
    """time-invariant MLSA filter of one signal: what Synthesizer.synthesis does with a tiled coefficient matrix
    (reference noise_shaping.py:80-85).  x: float64 (n,) -> float64 (n,)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    y = np.empty_like(x)
    if delay is None:
        delay = new_delay(len(coef) - 1, pd)
    lib().mlsa_filter_const(_p(x), len(x), _p(coef), len(coef) - 1, float(alpha), int(pd), float(np.exp(coef[0])), _p(delay),
                            _p(y))
    return y


def synthesis(x, coefs, alpha, hop, pd=4, delay=None):
    """pysptk Synthesizer.synthesis: coefs (nframes, order+1), linear interpolation inside each hop-sample frame."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    y = np.zeros_like(x)
    if delay is None:
        delay = new_delay(coefs.shape[1] - 1, pd)
    # b0 of every sample exactly as the frame loop accumulates it (prev + slope + slope ...), then numpy's exp as pysptk's
    # Python loop takes it
    c0 = np.empty(len(x), dtype=np.float64)
    for f in range(coefs.shape[0]):
        s0, s1 = f * hop, min((f + 1) * hop, len(x))
        if s0 >= len(x):
            break
        prev, curr = coefs[max(f - 1, 0), 0], coefs[f, 0]
        c0[s0:s1] = np.cumsum(np.concatenate([[prev], np.full(s1 - s0 - 1, (curr - prev) / float(hop))]))
    gain = np.exp(c0)
    lib().mlsa_synthesis(_p(x), len(x), _p(coefs), coefs.shape[0], coefs.shape[1] - 1, float(alpha), int(pd), int(hop),
                         _p(gain), _p(delay), _p(y))
    return y


def to_int16(y):
    """np.int16(float64 array) as the reference writes it (noise_shaping.py:87)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty(len(y), dtype=np.int16)
    lib().mlsa_to_int16(_p(y), len(y), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)))
    return out


def noise_shaping_one(x_int16, coef, alpha, fs=16000, shiftms=5, pd=4, delay=None):
    """One file of reference noise_shaping.py:66-87: int16 samples -> float64 -> tiled-coefficient synthesis -> np.int16."""
    x = np.float64(x_int16)
    hop = int(fs / 1000 * shiftms)
    nframes = int(len(x) / hop) + 1
    coefs = np.float64(np.tile(coef, [nframes, 1]))
    return to_int16(synthesis(x, coefs, alpha, hop, pd, delay))
