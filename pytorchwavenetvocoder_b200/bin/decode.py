#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""DECODE WITH WAVENET VOCODER -- B200 build of ``wavenet_vocoder/bin/decode.py`` (reference v0.1.1).

Same flags (reference decode.py:179-204), same inputs (``model.conf``, ``checkpoint*.pkl`` with key
``"model"``, feature ``.h5`` files, ``stats.h5``) and outputs (``{feat_id}.wav`` PCM_16 at ``--fs``).
Utterances are length-sorted and batched like the reference (``decode_generator`` :52-174) and sharded over
GPUs with ``np.array_split`` (:261-262), one process per GPU, no communication (:330-338).  Each batch
is ONE persistent-kernel launch (``WaveNet.batch_fast_generate``); like the reference the decoding mode is
"sampling".
"""
from __future__ import division

import argparse
import logging
import math
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

from sklearn.preprocessing import StandardScaler

from pytorchwavenetvocoder_b200.nets import decode_mu_law
from pytorchwavenetvocoder_b200.nets import encode_mu_law
from pytorchwavenetvocoder_b200.nets import WaveNet
from pytorchwavenetvocoder_b200.utils import extend_time
from pytorchwavenetvocoder_b200.utils import find_files
from pytorchwavenetvocoder_b200.utils import read_hdf5
from pytorchwavenetvocoder_b200.utils import read_txt
from pytorchwavenetvocoder_b200.utils import shape_hdf5
from pytorchwavenetvocoder_b200.utils import have_soundfile
from pytorchwavenetvocoder_b200.utils import write_wav
from pytorchwavenetvocoder_b200.utils import write_wav_pcm16


def pad_list(batch_list, pad_value=0.0):
    """PAD VALUE (reference decode.py:31-49): list of (T_i, C) -> (B, T_max, C), zero padded like there."""
    maxlen = max(b.shape[0] for b in batch_list)
    out = np.zeros((len(batch_list), maxlen, batch_list[0].shape[-1]))
    for i, b in enumerate(batch_list):
        out[i, :b.shape[0]] = b
    return out


def _load_feat(featfile, feature_type, upsampling_factor, use_upsampling_layer, use_speaker_code, feat_transform):
    h = read_hdf5(featfile, "/" + feature_type)
    if not use_upsampling_layer:
        h = extend_time(h, upsampling_factor)
    if use_speaker_code:
        sc = read_hdf5(featfile, "/speaker_code")
        h = np.concatenate([h, np.tile(sc, [h.shape[0], 1])], axis=1)
    if feat_transform is not None:
        h = feat_transform(h)
    n_samples = h.shape[0] - 1 if not use_upsampling_layer else h.shape[0] * upsampling_factor - 1
    return h, n_samples


def decode_generator(feat_list,
                     batch_size=32,
                     feature_type="world",
                     wav_transform=None,
                     feat_transform=None,
                     upsampling_factor=80,
                     use_upsampling_layer=True,
                     use_speaker_code=False):
    """GENERATE DECODING BATCH (reference decode.py:52-174).

    batch_size == 1: yields ``feat_id, (x (1,1), h (1,C,T), n_samples)`` per file;
    otherwise length-sorted batches ``feat_ids, (x (B,1), h (B,C,T_max), n_samples_list)``.
    The seed sample is ``wav_transform(0)`` (= 128 for mu-law 256)."""
    def seed():
        x = np.zeros((1))
        return wav_transform(x) if wav_transform is not None else x

    def cuda(t):
        return t.cuda() if torch.cuda.is_available() else t

    if batch_size == 1:
        for featfile in feat_list:
            h, n_samples = _load_feat(featfile, feature_type, upsampling_factor, use_upsampling_layer,
                                      use_speaker_code, feat_transform)
            x = torch.from_numpy(np.asarray(seed())).long().unsqueeze(0)
            ht = torch.from_numpy(np.asarray(h)).float().transpose(0, 1).unsqueeze(0)
            feat_id = os.path.basename(featfile).replace(".h5", "").replace(".npz", "")
            yield feat_id, (cuda(x), cuda(ht), n_samples)
    else:
        shape_list = [shape_hdf5(f, "/" + feature_type)[0] for f in feat_list]
        feat_list = [feat_list[i] for i in np.argsort(shape_list)]
        n_batch = math.ceil(len(feat_list) / batch_size)
        for batch_list in [f.tolist() for f in np.array_split(feat_list, n_batch)]:
            batch_x, batch_h, n_samples_list, feat_ids = [], [], [], []
            for featfile in batch_list:
                h, n_samples = _load_feat(featfile, feature_type, upsampling_factor, use_upsampling_layer,
                                          use_speaker_code, feat_transform)
                batch_x.append(seed())
                batch_h.append(h)
                n_samples_list.append(n_samples)
                feat_ids.append(os.path.basename(featfile).replace(".h5", "").replace(".npz", ""))
            bx = torch.from_numpy(np.stack(batch_x, axis=0)).long()
            bh = torch.from_numpy(pad_list(batch_h)).float().transpose(1, 2)
            yield feat_ids, (cuda(bx), cuda(bh), n_samples_list)


def get_parser():
    """Flags of reference decode.py:179-204."""
    parser = argparse.ArgumentParser()
    parser.add_argument("--feats", required=True, type=str, help="list or directory of aux feat files")
    parser.add_argument("--checkpoint", required=True, type=str, help="model file")
    parser.add_argument("--outdir", required=True, type=str, help="directory to save generated samples")
    parser.add_argument("--stats", default=None, type=str, help="hdf5 file including statistics")
    parser.add_argument("--config", default=None, type=str, help="configure file")
    parser.add_argument("--fs", default=16000, type=int, help="sampling rate")
    parser.add_argument("--batch_size", default=32, type=int, help="number of batch size in decoding")
    parser.add_argument("--n_gpus", default=1, type=int, help="number of gpus")
    parser.add_argument("--intervals", default=1000, type=int, help="log interval")
    parser.add_argument("--seed", default=1, type=int, help="seed number")
    parser.add_argument("--verbose", default=1, type=int, help="log level")
    return parser


def gpu_decode(feat_list, gpu, args, config):
    """One process per GPU (reference decode.py:274-327)."""
    # spawned workers do not inherit main()'s RNG state (the reference forks after seeding, decode.py:242):
    # re-seed here so that --seed controls the Philox stream of every GPU (distinct per GPU, reproducible)
    np.random.seed(args.seed + gpu)
    torch.manual_seed(args.seed + gpu)
    with torch.cuda.device(gpu):
        with torch.no_grad():
            upsampling_factor = config.upsampling_factor if config.use_upsampling_layer else 0
            model = WaveNet(
                n_quantize=config.n_quantize,
                n_aux=config.n_aux,
                n_resch=config.n_resch,
                n_skipch=config.n_skipch,
                dilation_depth=config.dilation_depth,
                dilation_repeat=config.dilation_repeat,
                kernel_size=config.kernel_size,
                upsampling_factor=upsampling_factor)
            model.load_state_dict(torch.load(
                args.checkpoint, map_location=lambda storage, loc: storage, weights_only=False)["model"])
            model.eval()
            model.cuda()
            scaler = StandardScaler()
            scaler.mean_ = read_hdf5(args.stats, "/" + config.feature_type + "/mean")
            scaler.scale_ = read_hdf5(args.stats, "/" + config.feature_type + "/scale")
            generator = decode_generator(
                feat_list,
                batch_size=args.batch_size,
                feature_type=config.feature_type,
                wav_transform=lambda x: encode_mu_law(x, config.n_quantize),
                feat_transform=lambda x: scaler.transform(x),
                upsampling_factor=config.upsampling_factor,
                use_upsampling_layer=config.use_upsampling_layer,
                use_speaker_code=config.use_speaker_code)
            if args.batch_size > 1:
                for feat_ids, (batch_x, batch_h, n_samples_list) in generator:
                    logging.info("decoding start")
                    # batch_fast_generate returns completion order = ascending length, ties by batch index
                    order = sorted(range(len(feat_ids)), key=lambda b: (n_samples_list[b], b))
                    if not have_soundfile():
                        # the PCM_16 quantisation is ours (stdlib writer): do decode_mu_law + quantisation for the whole
                        # batch on the device and move 2 bytes per sample (reference :316-319 does both on the host)
                        pcm_list = model.batch_fast_generate_pcm16(batch_x, batch_h, n_samples_list, args.intervals)
                        for b, pcm in zip(order, pcm_list):
                            write_wav_pcm16(args.outdir + "/" + feat_ids[b] + ".wav", pcm, args.fs)
                            logging.info("wrote %s.wav in %s." % (feat_ids[b], args.outdir))
                        continue
                    samples_list = model.batch_fast_generate(batch_x, batch_h, n_samples_list, args.intervals)
                    for b, samples in zip(order, samples_list):
                        wav = decode_mu_law(samples, config.n_quantize)
                        write_wav(args.outdir + "/" + feat_ids[b] + ".wav", wav, args.fs)   # libsndfile quantises
                        logging.info("wrote %s.wav in %s." % (feat_ids[b], args.outdir))
            else:
                for feat_id, (x, h, n_samples) in generator:
                    logging.info("decoding %s (length = %d)" % (feat_id, n_samples))
                    samples = model.fast_generate(x, h, n_samples, args.intervals)
                    wav = decode_mu_law(samples, config.n_quantize)
                    write_wav(args.outdir + "/" + feat_id + ".wav", wav, args.fs)
                    logging.info("wrote %s.wav in %s." % (feat_id, args.outdir))


def main():
    """RUN DECODING."""
    args = get_parser().parse_args()
    fmt = '%(asctime)s (%(module)s:%(lineno)d) %(levelname)s: %(message)s'
    level = logging.INFO if args.verbose > 0 else logging.WARNING
    if args.verbose > 1:
        level = logging.DEBUG
    logging.basicConfig(level=level, format=fmt, datefmt='%m/%d/%Y %I:%M:%S')
    if args.verbose < 1:
        logging.warning("logging is disabled.")
    for key, value in vars(args).items():
        logging.info("%s = %s" % (key, str(value)))
    if args.stats is None:
        args.stats = os.path.dirname(args.checkpoint) + "/stats.h5"      # reference decode.py:226-227
    if args.config is None:
        args.config = os.path.dirname(args.checkpoint) + "/model.conf"   # :228-229
    if not os.path.exists(args.outdir):
        os.makedirs(args.outdir)
    os.environ['PYTHONHASHSEED'] = str(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)

    config = torch.load(args.config, weights_only=False)   # argparse.Namespace written by train.py

    if os.path.isdir(args.feats):
        feat_list = sorted(find_files(args.feats, "*.h5")) or sorted(find_files(args.feats, "*.npz"))
    elif os.path.isfile(args.feats):
        feat_list = read_txt(args.feats)
    else:
        logging.error("--feats should be directory or list.")
        sys.exit(1)
    if not torch.cuda.is_available():
        logging.error("gpu is not available. please check the setting.")
        sys.exit(1)

    # utterance sharding: np.array_split over GPUs, one process each, no communication
    feat_lists = [f.tolist() for f in np.array_split(feat_list, args.n_gpus)]
    if args.n_gpus == 1:
        gpu_decode(feat_lists[0], 0, args, config)
        return
    ctx = mp.get_context("spawn")
    processes = []
    for gpu, f in enumerate(feat_lists):
        p = ctx.Process(target=gpu_decode, args=(f, gpu, args, config))
        p.start()
        processes.append(p)
    code = 0
    for p in processes:
        p.join()
        code = code or p.exitcode
    if code:
        logging.error("a decoding process failed (exit code %s)." % code)
        sys.exit(1)


if __name__ == "__main__":
    main()
