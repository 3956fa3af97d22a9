# -*- coding: utf-8 -*-
"""GPU tier: the CLI mirrors end to end (reference egs/*/run.sh stages 4 and 5): bin/train.py writes model.conf and
checkpoints in the reference format, resumes, and bin/decode.py turns feature files into PCM_16 wavs."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=timeout)
    assert r.returncode == 0, r.stdout.decode()[-3000:]
    return r.stdout.decode()


def test_train_resume_decode(tmp_path):
    from pytorchwavenetvocoder_b200.utils import read_wav, write_hdf5, write_wav
    rng = np.random.RandomState(0)
    U, D = 16, 28
    wavdir, expdir, outdir = str(tmp_path / "wav"), str(tmp_path / "exp"), str(tmp_path / "out")
    os.makedirs(wavdir)
    wavs, feats = [], []
    for i, n_frames in enumerate([60, 75, 50]):
        n = n_frames * U
        w, f = os.path.join(wavdir, "u%d.wav" % i), os.path.join(wavdir, "u%d.npz" % i)
        write_wav(w, 0.4 * np.sin(np.arange(n) / (5.0 + i)), 16000)
        write_hdf5(f, "/world", rng.standard_normal((n_frames, D)))
        wavs.append(w)
        feats.append(f)
    wl, fl, stats = str(tmp_path / "wav.scp"), str(tmp_path / "feats.scp"), str(tmp_path / "stats.npz")
    open(wl, "w").write("\n".join(wavs) + "\n")
    open(fl, "w").write("\n".join(feats) + "\n")
    write_hdf5(stats, "/world/mean", np.zeros(D))
    write_hdf5(stats, "/world/scale", np.ones(D))
    common = ["--waveforms", wl, "--feats", fl, "--stats", stats, "--expdir", expdir, "--n_resch", "64", "--n_skipch",
              "128", "--dilation_depth", "4", "--dilation_repeat", "2", "--upsampling_factor", str(U),
              "--batch_length", "320", "--batch_size", "2", "--intervals", "2", "--checkpoint_interval", "3",
              "--verbose", "1"]
    log = _run(["-m", "pytorchwavenetvocoder_b200.bin.train"] + common + ["--iters", "4"])
    assert "average loss" in log and os.path.exists(expdir + "/checkpoint-3.pkl")
    ck = torch.load(expdir + "/checkpoint-3.pkl", map_location="cpu", weights_only=False)
    assert set(ck) == {"model", "optimizer", "iterations"} and ck["iterations"] == 3
    assert "dil_sigmoid.0.conv.weight" in ck["model"] and "upsampling.conv.weight" in ck["model"]
    final = torch.load(expdir + "/checkpoint-final.pkl", map_location="cpu", weights_only=False)
    assert set(final) == {"model"}
    conf = torch.load(expdir + "/model.conf", weights_only=False)
    assert conf.n_resch == 64 and conf.upsampling_factor == U
    # resume from the 3-iteration checkpoint (reference train.py:503-513)
    log = _run(["-m", "pytorchwavenetvocoder_b200.bin.train"] + common +
               ["--iters", "5", "--resume", expdir + "/checkpoint-3.pkl"])
    assert "restored from 3-iter checkpoint" in log
    # decode (batch mode and single-utterance mode), reference decode.py
    for bs in ("2", "1"):
        od = outdir + bs
        _run(["-m", "pytorchwavenetvocoder_b200.bin.decode", "--feats", fl, "--checkpoint", expdir + "/checkpoint-final.pkl",
              "--stats", stats, "--outdir", od, "--fs", "16000", "--batch_size", bs, "--intervals", "100"])
        for i, n_frames in enumerate([60, 75, 50]):
            x, fs = read_wav(os.path.join(od, "u%d.wav" % i))
            assert fs == 16000 and len(x) == n_frames * U - 1      # reference decode.py:108-111
            assert np.all(np.abs(x) <= 1.03)


def test_train_and_decode_with_speaker_code_without_upsampling_layer(tmp_path):
    """f3: `--use_speaker_code true --use_upsampling_layer false` (reference train.py:123-128, decode.py:83-88): the
    speaker code is tiled onto the aux features (n_aux = 28 + 3), the aux features arrive at sample rate through
    `extend_time`; device loader in training, `batch_fast_generate_pcm16` in decoding."""
    from pytorchwavenetvocoder_b200.utils import read_wav, write_hdf5, write_wav
    rng = np.random.RandomState(1)
    U, D, NS = 16, 28, 3
    wavdir, expdir, outdir = str(tmp_path / "wav"), str(tmp_path / "exp"), str(tmp_path / "out")
    os.makedirs(wavdir)
    wavs, feats = [], []
    for i, n_frames in enumerate([70, 55, 64]):
        n = n_frames * U - 3
        w, f = os.path.join(wavdir, "s%d.wav" % i), os.path.join(wavdir, "s%d.npz" % i)
        write_wav(w, 0.4 * np.sin(np.arange(n) / (4.0 + i)), 16000)
        write_hdf5(f, "/world", rng.standard_normal((n_frames, D)))
        write_hdf5(f, "/speaker_code", np.eye(NS)[i % NS])
        wavs.append(w)
        feats.append(f)
    wl, fl, stats = str(tmp_path / "wav.scp"), str(tmp_path / "feats.scp"), str(tmp_path / "stats.npz")
    open(wl, "w").write("\n".join(wavs) + "\n")
    open(fl, "w").write("\n".join(feats) + "\n")
    write_hdf5(stats, "/world/mean", np.zeros(D + NS))
    write_hdf5(stats, "/world/scale", np.ones(D + NS))
    log = _run(["-m", "pytorchwavenetvocoder_b200.bin.train", "--waveforms", wl, "--feats", fl, "--stats", stats,
                "--expdir", expdir, "--n_aux", str(D + NS), "--n_resch", "64", "--n_skipch", "64", "--dilation_depth", "3",
                "--dilation_repeat", "2", "--upsampling_factor", str(U), "--use_upsampling_layer", "false",
                "--use_speaker_code", "true", "--batch_length", "300", "--batch_size", "2", "--iters", "3",
                "--intervals", "1", "--checkpoint_interval", "3", "--verbose", "1"])
    assert "average loss" in log and os.path.exists(expdir + "/checkpoint-final.pkl")
    conf = torch.load(expdir + "/model.conf", weights_only=False)
    assert conf.use_speaker_code is True and conf.use_upsampling_layer is False and conf.n_aux == D + NS
    _run(["-m", "pytorchwavenetvocoder_b200.bin.decode", "--feats", fl, "--checkpoint", expdir + "/checkpoint-final.pkl",
          "--stats", stats, "--outdir", outdir, "--fs", "16000", "--batch_size", "3", "--intervals", "100"])
    for i, n_frames in enumerate([70, 55, 64]):
        x, fs = read_wav(os.path.join(outdir, "s%d.wav" % i))
        assert fs == 16000 and len(x) == n_frames * U - 1
