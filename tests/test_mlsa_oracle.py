# -*- coding: utf-8 -*-
"""CPU tier: the MLSA noise-shaping oracle (oracle/mlsa_oracle.c).

pysptk -- the library the reference calls (bin/noise_shaping.py:41, 59-64, 85) -- is not in this image, so the
restatement of SPTK's mc2b / mlsadf is pinned by what an MLSA filter IS: its frequency response is
exp(sum_m c_m e^{-j m w~}) on the alpha-warped frequency axis w~ (to the accuracy of the order-pd Pade approximant of
exp), filtering with -coef inverts filtering with coef, and the filter is linear and time invariant.  When pysptk is
importable the same cases are compared with it directly."""
import numpy as np
import pytest

from oracle import mlsa_oracle as M
from pytorchwavenetvocoder_b200.utils import mlsa as P


def _warp(w, a):
    return w + 2 * np.arctan(a * np.sin(w) / (1 - a * np.cos(w)))


@pytest.mark.parametrize("pd,alpha,order,scale,c0", [(4, 0.41, 24, 0.3, 0.0), (5, 0.41, 24, 0.3, 0.2), (4, 0.455, 59, 0.12, -0.1),
                                                     (4, 0.35, 12, 0.4, 0.0)])
def test_frequency_response_is_exp_of_the_mel_cepstrum(pd, alpha, order, scale, c0):
    rng = np.random.RandomState(order + pd)
    mc = rng.randn(order + 1) * scale
    mc[0] = c0
    b = M.mc2b(mc, alpha)
    n = 8192
    imp = np.zeros(n)
    imp[0] = 1.0
    h = M.filter_const(imp, b, alpha, pd)
    assert abs(h[-1]) < 1e-9 * np.abs(h).max()          # the impulse response has died out inside the window
    H = np.fft.rfft(h)
    wt = _warp(np.linspace(0, np.pi, len(H)), alpha)
    F = sum(mc[m] * np.exp(-1j * m * wt) for m in range(order + 1))
    rel = np.abs(H - np.exp(F)).max() / np.abs(np.exp(F)).max()
    print("pd %d alpha %.3f order %d: max |log H| %.2f, relative error of the response %.2e" % (pd, alpha, order, np.abs(F).max(), rel))
    assert np.abs(F).max() > 1.0                         # a real test of exp(), not of 1 + F
    assert rel < (2e-3 if pd == 4 else 5e-4)


def test_mc2b_and_product_side_coefficients():
    rng = np.random.RandomState(0)
    mc = rng.randn(25)
    b = M.mc2b(mc, 0.41)
    back = b.copy()
    back[:-1] += 0.41 * b[1:]                            # b2mc
    np.testing.assert_allclose(back, mc, rtol=0, atol=1e-15)
    assert np.array_equal(P.mc2b(mc, 0.41), b)           # the product's host-side mc2b: same recursion, bit for bit
    avg = rng.randn(25)
    want = M.convert_mcep_to_mlsa_coef(avg, 0.5, 0.41)
    got = P.convert_mcep_to_mlsa_coef(avg.copy(), 0.5, 0.41)
    assert np.array_equal(got, want) and want[0] == -0.41 * want[1]


def test_inverse_filter_linearity_and_frame_loop():
    rng = np.random.RandomState(1)
    mc = rng.randn(25) * 0.15
    mc[0] = 0.0
    b = M.mc2b(mc, 0.41)
    x = np.cumsum(rng.randn(6000)) * 30.0                # low-pass noise, wav-like amplitudes
    x -= x.mean()
    y = M.filter_const(x, b, 0.41)
    xr = M.filter_const(y, -b, 0.41)
    assert np.abs(xr - x).max() / np.abs(x).max() < 1e-3   # reference noise_shaping.py:55-56: --inv restores the signal
    x2 = rng.randn(6000) * 100.0
    lin = M.filter_const(2.0 * x + x2, b, 0.41) - (2.0 * y + M.filter_const(x2, b, 0.41))
    assert np.abs(lin).max() / np.abs(y).max() < 1e-12
    # pysptk's Synthesizer frame loop with the reference's tiled coefficient matrix == the time-invariant filter
    hop = 80
    coefs = np.tile(b, [len(x) // hop + 1, 1])
    assert np.array_equal(M.synthesis(x, coefs, 0.41, hop), y)
    # delayed input -> delayed output (time invariance, zero initial state)
    xd = np.concatenate([np.zeros(37), x])
    np.testing.assert_allclose(M.filter_const(xd, b, 0.41)[37:], y, rtol=0, atol=1e-9)


def test_int16_path_of_the_cli():
    rng = np.random.RandomState(2)
    x16 = (np.cumsum(rng.randn(4000)) * 40).astype(np.int16)
    coef = M.convert_mcep_to_mlsa_coef(rng.randn(25) * 0.4, 0.5, 0.41)
    out = M.noise_shaping_one(x16, coef, 0.41)
    y = M.filter_const(np.float64(x16), coef, 0.41)
    assert out.dtype == np.int16 and np.array_equal(out, np.int16(y))      # np.int16(): truncation toward zero
    assert np.array_equal(M.to_int16(np.array([-1.7, -0.2, 0.9, 32767.9, 1.0])), np.array([-1, 0, 0, 32767, 1], dtype=np.int16))


def test_against_pysptk_when_available():
    pysptk = pytest.importorskip("pysptk")
    rng = np.random.RandomState(3)
    mc = rng.randn(25) * 0.2
    mc[0] = 0.0
    b = pysptk.mc2b(mc, 0.41)
    assert np.array_equal(b, M.mc2b(mc, 0.41))
    x = rng.randn(2000) * 1000.0
    syn = pysptk.synthesis.Synthesizer(pysptk.synthesis.MLSADF(order=24, alpha=0.41), hopsize=80)
    want = syn.synthesis(x, np.tile(b, [len(x) // 80 + 1, 1]))
    np.testing.assert_array_equal(M.filter_const(x, b, 0.41), want)


def test_cli_writes_the_filter_coefficients_like_the_reference(tmp_path):
    """bin/noise_shaping.py main() up to the filter launch (an empty wav directory: no GPU needed): /mlsa/coef and
    /mlsa/alpha are derived from <feature_type>/mean exactly as reference noise_shaping.py:171-177 does and are not
    recomputed when they already exist."""
    from pytorchwavenetvocoder_b200.bin import noise_shaping as ns
    from pytorchwavenetvocoder_b200.utils import check_hdf5, read_hdf5, write_hdf5
    rng = np.random.RandomState(11)
    stats = str(tmp_path / "stats.npz")
    mean = rng.randn(28)
    write_hdf5(stats, "/world/mean", mean)
    (tmp_path / "wav").mkdir()
    argv = ["--waveforms", str(tmp_path / "wav"), "--stats", stats, "--outdir", str(tmp_path / "out"), "--mag", "0.5",
            "--mcep_alpha", "0.41", "--verbose", "0"]
    ns.main(argv)
    assert check_hdf5(stats, "/mlsa/coef") and (tmp_path / "out").is_dir()
    want = M.convert_mcep_to_mlsa_coef(mean[2:27], 0.5, 0.41)
    assert np.array_equal(read_hdf5(stats, "/mlsa/coef"), want) and float(read_hdf5(stats, "/mlsa/alpha")) == 0.41
    assert np.array_equal(read_hdf5(stats, "/world/mean"), mean)          # the stats themselves are untouched
    write_hdf5(stats, "/mlsa/coef", want * 2.0)                            # an existing coefficient set is used as is
    ns.main(argv)
    assert np.array_equal(read_hdf5(stats, "/mlsa/coef"), want * 2.0)
    # mcep features: the whole mean vector is the mel-cepstrum (reference :173-175)
    stats2 = str(tmp_path / "stats2.npz")
    write_hdf5(stats2, "/mcep/mean", mean[:25])
    ns.main(["--waveforms", str(tmp_path / "wav"), "--stats", stats2, "--outdir", str(tmp_path / "out"), "--feature_type", "mcep",
             "--verbose", "0"])
    assert np.array_equal(read_hdf5(stats2, "/mlsa/coef"), M.convert_mcep_to_mlsa_coef(mean[:25], 0.5, 0.41))
