# -*- coding: utf-8 -*-
"""Seeded case definitions shared by make_golden.py (reference side) and the tests (our side).

Config tuples are WaveNet ctor args (n_quantize, n_aux, n_resch, n_skipch, dilation_depth,
dilation_repeat, kernel_size, upsampling_factor), wavenet.py:172-173.  Shapes follow the
reference's own tests (test/test_wavenet.py:39-66, 80, 104-253) plus BASELINE.json configs[0].
"""
import numpy as np

# name -> (cfg, seed, B, T, loss_start)
FORWARD_CASES = {
    "tiny": ((256, 28, 8, 16, 4, 1, 2, 0), 11, 2, 80, 16),
    "tiny_up": ((256, 28, 8, 16, 4, 1, 2, 10), 12, 2, 80, 16),
    "ks3_up": ((256, 28, 32, 128, 10, 1, 3, 10), 13, 1, 100, 20),      # test_wavenet.py:66
    "ks2_r32": ((256, 28, 32, 128, 10, 1, 2, 0), 14, 2, 100, 10),      # test_wavenet.py:39
    "d10x3": ((256, 28, 4, 4, 10, 3, 2, 0), 15, 2, 64, 8),             # test_wavenet.py:80
    "mini": ((256, 28, 32, 16, 5, 1, 2, 8), 16, 3, 96, 32),            # egs/arctic/sd-mini/run.sh:47-50
}

# name -> (cfg, seed, B, T0 seed length, n_samples_list, also run naive generate)
GEN_CASES = {
    "tiny1000": ((256, 28, 8, 16, 4, 1, 2, 0), 1234, 1, 1, [1000], True),      # BASELINE cfg[0]
    "ks2": ((256, 28, 4, 4, 10, 3, 2, 0), 21, 2, 1, [32, 32], True),            # test_wavenet.py:104-128
    "ks3": ((256, 28, 4, 4, 10, 3, 3, 0), 22, 2, 1, [32, 32], True),            # :133-157
    "ks2_up": ((256, 28, 4, 4, 10, 3, 2, 10), 23, 2, 1, [29, 29], True),        # :169-193
    "ragged": ((256, 28, 4, 4, 10, 3, 2, 0), 24, 4, 1, [30, 16, 25, 16], False),  # :225-253
    "ragged_up": ((256, 28, 8, 16, 4, 2, 3, 5), 25, 3, 4, [34, 19, 24], False),
    "longseed": ((256, 28, 8, 16, 4, 1, 2, 0), 26, 2, 40, [20, 20], True),      # T0 > rf (no pad)
}


def make_inputs(cfg, seed, B, T):
    rng = np.random.RandomState(seed + 1000)
    x = rng.randint(0, cfg.n_quantize, size=(B, T)).astype(np.int64)
    t = rng.randint(0, cfg.n_quantize, size=(B, T)).astype(np.int64)
    U = cfg.upsampling_factor
    Th = T // U if U > 0 else T
    h = rng.standard_normal((B, cfg.n_aux, Th)).astype(np.float32)
    return x, h, t


def make_gen_inputs(cfg, seed, B, T0, n_list):
    rng = np.random.RandomState(seed + 2000)
    x = rng.randint(0, cfg.n_quantize, size=(B, T0)).astype(np.int64)
    if T0 == 1 and B == 1:
        x[:] = cfg.n_quantize // 2
    U = cfg.upsampling_factor
    need = max(n_list) + T0
    Th = (need + U - 1) // U if U > 0 else need
    h = rng.standard_normal((B, cfg.n_aux, Th)).astype(np.float32)
    return x, h


def mulaw_inputs():
    rng = np.random.RandomState(7)
    grid = np.linspace(-1, 1, 20001)
    kat = np.array([-1, -.5, -.1, -.01, -1e-3, 0, 1e-3, .01, .1, .5, .999, 1])
    # bin edges of the quantiser +- a few ulps: where rounding bugs show up
    q = np.arange(1, 256) - 0.5
    fx = q / 255.0 * 2 - 1
    edges = np.sign(fx) / 255.0 * (256.0 ** np.abs(fx) - 1)
    e32 = edges.astype(np.float32)
    near = np.concatenate([np.nextafter(e32, np.float32(-2)), e32, np.nextafter(e32, np.float32(2))])
    x64 = np.concatenate([kat, grid, edges, rng.uniform(-1, 1, 4096)])
    x32 = np.concatenate([x64.astype(np.float32), near])
    codes = np.arange(256, dtype=np.int64)
    return x32, x64, codes


def mulaw_pcm16_domain():
    """Every float32 that sf.read(dtype=float32) can produce from a PCM_16 wav (bin/train.py:121)."""
    return np.arange(-32768, 32768).astype(np.float32) / np.float32(32768)


def mulaw_edge_mask(x32):
    """True where a float32 input lies within 2 ulp of a quantiser edge (reference result there
    depends on numpy's non-correctly-rounded SIMD float32 log; see csrc/elementwise.cu)."""
    q = np.arange(1, 256) - 0.5
    fx = q / 255.0 * 2 - 1
    edges = (np.sign(fx) / 255.0 * (256.0 ** np.abs(fx) - 1))
    d = np.abs(x32.astype(np.float64)[:, None] - edges[None, :]).min(axis=1)
    return d <= 2 * np.spacing(np.abs(x32).astype(np.float32)).astype(np.float64)
