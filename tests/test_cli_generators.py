# -*- coding: utf-8 -*-
"""CPU tier: host-side data plumbing of the CLI mirrors (reference test/test_generator.py:53-212 shape
contracts, with .npz features and stdlib wav I/O instead of sprocket/librosa/h5py)."""
import os

import numpy as np
import pytest
import torch

from pytorchwavenetvocoder_b200.utils import extend_time, read_hdf5, write_hdf5, write_wav


@pytest.fixture(scope="module")
def corpus(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("corpus"))
    rng = np.random.RandomState(0)
    U, D = 80, 28
    wavs, feats = [], []
    for i, n_frames in enumerate([90, 120, 75, 101]):
        n = n_frames * U + rng.randint(-30, 30)
        w = os.path.join(root, "u%d.wav" % i)
        f = os.path.join(root, "u%d.npz" % i)
        write_wav(w, 0.3 * np.sin(np.arange(n) / (7.0 + i)), 16000)
        write_hdf5(f, "/world", rng.standard_normal((n_frames, D)))
        wavs.append(w)
        feats.append(f)
    return wavs, feats, U, D


def _enc(x):   # CPU stand-in for the device mu-law (the generator only needs integer codes here)
    from oracle.wavenet_oracle import encode_mu_law
    return encode_mu_law(x, 256)


@pytest.mark.parametrize("batch_length,batch_size,use_up", [(2000, 3, True), (2000, 2, False), (None, 1, True),
                                                            (None, 1, False)])
def test_train_generator_shapes(corpus, batch_length, batch_size, use_up):
    from pytorchwavenetvocoder_b200.bin.train import train_generator
    wavs, feats, U, D = corpus
    rf = 1 + 2 * 1023   # 2 x 10 layers, ks 2
    gen = train_generator(wavs, feats, receptive_field=rf, batch_length=batch_length, batch_size=batch_size,
                          wav_transform=_enc, feat_transform=None, shuffle=False, upsampling_factor=U,
                          use_upsampling_layer=use_up)
    for _ in range(3):
        (x, h), t = gen.next()
        assert x.dtype == torch.int64 and h.dtype == torch.float32
        assert x.shape == t.shape and h.size(1) == D
        if batch_length is not None:
            assert x.size(0) == batch_size
        else:
            assert x.size(0) == 1
        if use_up:
            assert x.size(1) == h.size(2) * U          # reference test_generator.py:99-101
        else:
            assert x.size(1) == h.size(2)              # :115-117
        if batch_length is not None and use_up:
            assert x.size(1) == ((rf + batch_length) // U) * U
        assert torch.equal(x[:, 1:], t[:, :-1])        # next-sample targets


@pytest.mark.parametrize("batch_size,use_up", [(1, True), (1, False), (3, True), (3, False)])
def test_decode_generator_shapes(corpus, batch_size, use_up):
    from pytorchwavenetvocoder_b200.bin.decode import decode_generator
    wavs, feats, U, D = corpus
    gen = decode_generator(feats, batch_size=batch_size, wav_transform=_enc, feat_transform=None,
                           upsampling_factor=U, use_upsampling_layer=use_up)
    seen = 0
    for ids, (x, h, n) in gen:
        if batch_size == 1:
            assert x.shape == (1, 1) and int(x) == 128
            assert (h.size(2) * U if use_up else h.size(2)) == n + 1       # reference test_generator.py:171-173
            seen += 1
        else:
            assert x.size(0) == h.size(0) == len(n) == len(ids)
            assert (h.size(2) * U if use_up else h.size(2)) == max(n) + 1  # :190-191
            assert list(n) == sorted(n)                                    # length-sorted batches
            seen += len(ids)
    assert seen == len(feats)


def test_extend_time_and_alias_package():
    f = np.arange(6.0).reshape(3, 2)
    e = extend_time(f, 4)
    assert e.shape == (12, 2) and np.array_equal(e[4:8], np.repeat(f[1:2], 4, axis=0))
    import wavenet_vocoder.nets as N
    import pytorchwavenetvocoder_b200.nets as M
    if N.WaveNet.__module__.startswith("pytorchwavenetvocoder_b200"):   # repo root first on sys.path
        assert N.WaveNet is M.WaveNet
    from pytorchwavenetvocoder_b200.bin import decode, train
    a = train.get_parser().parse_args(["--waveforms", "w", "--feats", "f", "--stats", "s", "--expdir", "e"])
    assert a.n_resch == 512 and a.n_skipch == 256 and a.batch_length == 20000 and a.use_upsampling_layer is True
    b = decode.get_parser().parse_args(["--feats", "f", "--checkpoint", "c", "--outdir", "o"])
    assert b.batch_size == 32 and b.fs == 16000


@pytest.mark.parametrize("batch_length,batch_size,use_up", [(2000, 3, True), (1500, 2, False), (900, 1, True), (2000, 4, True)])
def test_window_planner_reproduces_the_host_generator(corpus, batch_length, batch_size, use_up):
    """utils/device_loader.WindowPlanner (the host bookkeeping of the on-device train_generator) + a plain numpy stream
    give exactly the batches of train_generator, including the reference's dropped incomplete batch at each epoch end."""
    from pytorchwavenetvocoder_b200.bin.train import train_generator
    from pytorchwavenetvocoder_b200.utils.device_loader import WindowPlanner, load_pair_frames
    wavs, feats, U, D = corpus
    rf = 1 + 1023
    host = train_generator(wavs, feats, receptive_field=rf, batch_length=batch_length, batch_size=batch_size,
                           wav_transform=_enc, feat_transform=None, shuffle=False, upsampling_factor=U,
                           use_upsampling_layer=use_up)
    want = [host.next() for _ in range(14)]          # several epochs of this corpus
    if use_up:
        bl = batch_length - (rf + batch_length) % U
        h_bs = (rf + bl) // U
        T, hop, need = h_bs * U, (bl // U) * U, (h_bs + 1) * U
    else:
        T, hop, need = rf + batch_length - 1, batch_length, rf + batch_length + 1
    plan = WindowPlanner(batch_size, hop, need)
    xs, hs = [], []
    for ep in range(12):
        for i, (w, f) in enumerate(zip(wavs, feats)):
            x, h = load_pair_frames(w, f, "world", U, use_up, False)
            xs.append(x)
            hs.append(h if use_up else h[np.arange(len(x)) // U])
            plan.append(len(x), i + 1 == len(wavs))
    sx, sh = np.concatenate(xs), np.concatenate(hs)
    for k, ((xw, hw), tw) in enumerate(want):
        s0 = plan.ready[k]
        rows = np.stack([_enc(sx[s0 + b * hop:s0 + b * hop + T + 1]) for b in range(batch_size)])
        assert np.array_equal(rows[:, :-1], xw.numpy()) and np.array_equal(rows[:, 1:], tw.numpy()), k
        f0 = [(s0 + b * hop) // U if use_up else s0 + b * hop for b in range(batch_size)]
        Tf = T // U if use_up else T
        hh = np.stack([sh[f:f + Tf].T for f in f0]).astype(np.float32)
        assert np.array_equal(hh, hw.numpy()), k
