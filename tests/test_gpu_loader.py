# -*- coding: utf-8 -*-
"""GPU tier: the on-device train_generator (utils/device_loader.py + csrc/loader.cu) yields exactly the
batches of the host generator (bin/train.py train_generator = reference train.py:67-312) on the same file lists."""
import os

import numpy as np
import pytest
import torch

from pytorchwavenetvocoder_b200.utils import write_hdf5, write_wav

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["float64", "float32"])
def corpus(request, tmp_path_factory):
    root = str(tmp_path_factory.mktemp("corpus_" + request.param))
    rng = np.random.RandomState(0)
    U, D = 80, 28
    wavs, feats = [], []
    for i, n_frames in enumerate([90, 120, 75, 101, 33, 140]):
        n = n_frames * U + rng.randint(-30, 30)
        w = os.path.join(root, "u%d.wav" % i)
        f = os.path.join(root, "u%d.npz" % i)
        write_wav(w, 0.6 * np.sin(np.arange(n) / (7.0 + i)) * np.cos(np.arange(n) / 301.0), 16000)
        feat = rng.standard_normal((n_frames, D))
        write_hdf5(f, "/world", feat.astype(request.param))     # both storage dtypes the recipes produce
        write_hdf5(f, "/speaker_code", np.eye(3)[i % 3].astype(request.param))
        wavs.append(w)
        feats.append(f)
    mean = rng.standard_normal(D + 3)
    scale = 0.5 + rng.rand(D + 3)
    return wavs, feats, U, D, mean, scale


@pytest.mark.parametrize("batch_length,batch_size,use_up,shuffle,use_spk", [
    (2000, 3, True, False, False), (1500, 2, False, False, False), (2000, 4, True, True, False), (900, 1, True, False, False),
    (2000, 2, True, False, True), (1200, 2, False, True, True)])      # use_speaker_code: reference train.py:123-128
def test_device_generator_matches_host_generator(corpus, batch_length, batch_size, use_up, shuffle, use_spk):
    from sklearn.preprocessing import StandardScaler
    from pytorchwavenetvocoder_b200.bin.train import train_generator
    from pytorchwavenetvocoder_b200.nets import encode_mu_law
    from pytorchwavenetvocoder_b200.utils.device_loader import DeviceTrainGenerator
    wavs, feats, U, D, mean, scale = corpus
    if not use_spk:
        mean, scale = mean[:D], scale[:D]
    rf = 1 + 1023
    scaler = StandardScaler()
    scaler.mean_, scaler.scale_ = mean, scale
    np.random.seed(5)
    host = train_generator(wavs, feats, receptive_field=rf, batch_length=batch_length, batch_size=batch_size,
                           wav_transform=lambda x: encode_mu_law(x, 256), feat_transform=scaler.transform,
                           shuffle=shuffle, upsampling_factor=U, use_upsampling_layer=use_up, use_speaker_code=use_spk)
    want = [host.next() for _ in range(12)]          # > 1 epoch of this corpus: crosses the reshuffle
    np.random.seed(5)
    devg = DeviceTrainGenerator(wavs, feats, rf, batch_length, batch_size, n_quantize=256, mean=mean, scale=scale,
                                shuffle=shuffle, upsampling_factor=U, use_upsampling_layer=use_up,
                                use_speaker_code=use_spk)
    for k, ((xw, hw), tw) in enumerate(want):
        (x, h), t = devg.next()
        assert x.is_cuda and x.dtype == torch.int64 and h.dtype == torch.float32
        assert torch.equal(x.cpu(), xw.cpu()), k
        assert torch.equal(t.cpu(), tw.cpu()), k
        assert h.size(1) == D + (3 if use_spk else 0)
        assert torch.equal(h.cpu(), hw.cpu()), (k, (h.cpu() - hw.cpu()).abs().max())
