# -*- coding: utf-8 -*-
"""CPU tier: the N>1 host logic (gradient all-reduce, utterance sharding) with gloo, world_size 2."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pytorchwavenetvocoder_b200.parallel import GradAllReduce, shard_utterances
    torch.manual_seed(rank)   # different init per rank: the wrapper must broadcast rank 0's weights
    m = torch.nn.Linear(5, 3)
    sync = GradAllReduce(m)
    w0 = m.weight.detach().clone()
    x = torch.full((4, 5), float(rank + 1))
    m(x).sum().backward()
    # leave one parameter without grad on purpose
    extra = torch.nn.Parameter(torch.zeros(2))
    sync.params.append(extra)
    sync.allreduce()
    q.put((rank, w0.numpy(), m.weight.grad.numpy().copy(), shard_utterances(list("abcdefg"), world, rank)))
    dist.destroy_process_group()


def test_grad_allreduce_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda r: r[0])
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert np.array_equal(res[0][1], res[1][1])                 # weights broadcast from rank 0
    # grad of sum(Wx+b) wrt W is sum over batch of x: rank r -> 4*(r+1); mean over ranks = 6
    assert np.allclose(res[0][2], 6.0) and np.allclose(res[1][2], 6.0)
    assert res[0][3] == list("abcd") and res[1][3] == list("efg")   # np.array_split semantics
