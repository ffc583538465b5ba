"""world_size-2 gloo test of the N>1 host logic: batch dealing, per-bin count gather and
ordered merge.  The per-batch results come from the oracle here (no GPU in this tier); the
`-m gpu` tests cover the kernels, bench.py --gpus N the NCCL path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from orcdemux import shard, synth

N_BATCHES, BATCH = 5, 300


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _bins(rs):
    rec0, rec1, *_ = H.run_oracle(H.m13_rounds(), rs, n_threads=2)
    b = (rec0["adapter"] + 1) + 13 * (rec1["adapter"] + 1)
    return np.bincount(b, minlength=169).astype(np.int64)


def _worker(rank, world, port, out_dir):
    for p in (H.ROOT, H.PKG, os.path.join(H.ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    local = np.zeros(169, dtype=np.int64)
    for b in shard.batches_of_rank(N_BATCHES, rank, world):
        assert shard.owner_of_batch(b, world) == rank
        local += _bins(synth.generate(BATCH, 300, 500, seed=1000 + b))
    total = shard.gather_counts(local)
    np.save(os.path.join(out_dir, "counts_%d.npy" % rank), total)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_count_gather(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    c0 = np.load(tmp_path / "counts_0.npy")
    c1 = np.load(tmp_path / "counts_1.npy")
    assert np.array_equal(c0, c1)
    whole = sum(_bins(synth.generate(BATCH, 300, 500, seed=1000 + b)) for b in range(N_BATCHES))
    assert np.array_equal(c0, whole)
    assert int(c0.sum()) == N_BATCHES * BATCH


def test_dealing_and_merge():
    seen = sorted(b for r in range(3) for b in shard.batches_of_rank(10, r, 3))
    assert seen == list(range(10))
    parts = {2: b"c", 0: b"a", 1: [b"b", b"B"]}
    assert shard.merge_in_order(parts) == b"abBc"
    assert np.array_equal(shard.gather_counts(np.arange(4)), np.arange(4))
