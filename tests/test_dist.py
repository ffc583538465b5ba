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


def _bin_text(batch, b):
    """Deterministic FASTQ-looking text of bin b in batch `batch` (bin 2 stays empty)."""
    if b == 2:
        return b""
    n = 40 + 13 * batch + 7 * b
    return b"".join(b"@b%d_%d_%d\nACGT\n+\nIIII\n" % (batch, b, i) for i in range(n)) * (300 if batch == 3 else 1)


class _Res:
    def __init__(self, parts):
        self.fastq = np.frombuffer(b"".join(parts), dtype=np.uint8)
        self.bin_offsets = np.cumsum([0] + [len(p) for p in parts]).astype(np.uint64)


def _part_writer(rank, world, port, out_dir, gz):
    for p in (H.ROOT, H.PKG, os.path.join(H.ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from orcdemux import fastq as F
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ext = ".fastq.gz" if gz else ".fastq"
    paths = [os.path.join(out_dir, "bin%d%s" % (b, ext)) for b in range(3)] + [None]
    w = F.BinWriters([shard.part_path(p, rank) if p else None for p in paths], 1, threads=2, index=True)
    tickets = [w.write_batch(_Res([_bin_text(batch, b) for b in range(3)] + [b""]))
               for batch in shard.batches_of_rank(N_BATCHES, rank, world)]
    for t in tickets:
        w.wait(t)
    w.close()
    dist.barrier()
    if rank == 0:
        shard.merge_part_files(paths, world, 1)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("gz", [True, False])
def test_two_rank_part_files_merge_in_batch_order(tmp_path, gz):
    """What `orcdemux.cli two-round --gpus 2` does on the host: every rank writes its batches' bins to part
    files, rank 0 stitches them in batch order.  The result equals what one writer fed all batches in order
    produces (decompressed), empty bins are valid files, nothing but the final files is left."""
    import gzip
    from orcdemux import fastq as F
    port = _free_port()
    mp.spawn(_part_writer, args=(2, port, str(tmp_path), gz), nprocs=2, join=True)
    ext = ".fastq.gz" if gz else ".fastq"
    for b in range(3):
        blob = (tmp_path / ("bin%d%s" % (b, ext))).read_bytes()
        got = gzip.decompress(blob) if gz else blob
        assert got == b"".join(_bin_text(batch, b) for batch in range(N_BATCHES)), b
        if gz:      # still a file the member-parallel reader takes (bin 3 of batch 3 spans several chunks)
            with F.FastqReader(str(tmp_path / ("bin%d%s" % (b, ext))), 1 << 16, 1 << 24, pinned=False, threads=3) as rd:
                assert b"".join(tb.text[:tb.n_bytes].tobytes() for tb in rd) == got
    assert sorted(os.listdir(tmp_path)) == sorted("bin%d%s" % (b, ext) for b in range(3))
