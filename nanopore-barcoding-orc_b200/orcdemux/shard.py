"""Multi-GPU host logic of the product command (`python -m orcdemux.cli two-round --gpus N`, or the same
under torchrun): one process per GPU; reads are independent, so the batches of the input are dealt round-robin
to the ranks with no collective on the data path.  Every rank writes its bins to part files; the only
exchanges are the final per-bin count gather (all_reduce of n_bins counters) and a barrier before rank 0
stitches the part files together in batch order -- cutadapt's ordered chunk writer (runners.py) keeps input
order inside every output file, and so does this.
"""
from __future__ import annotations

import ctypes as C
import os
import socket
import subprocess
import sys
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def batches_of_rank(n_batches: int, rank: int, world: int) -> List[int]:
    """Batch ids dealt to `rank`: rank, rank + world, ..."""
    return list(range(rank, n_batches, world))


def owner_of_batch(batch_id: int, world: int) -> int:
    return batch_id % world


def dist_env() -> Tuple[int, int, int]:
    """(rank, world, local rank) from the torchrun-style environment; (0, 1, 0) without one."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    return rank, world, int(os.environ.get("LOCAL_RANK", str(rank)))


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def spawn_ranks(n: int, argv: Sequence[str]) -> int:
    """Run `python -m orcdemux.cli argv` as n ranks of one node (what torchrun would do) and wait for them."""
    port = free_port()
    procs = []
    pkg = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pp = os.environ.get("PYTHONPATH", "")
    for r in range(n):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(n), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PYTHONPATH=pkg + (os.pathsep + pp if pp else ""))
        procs.append(subprocess.Popen([sys.executable, "-m", "orcdemux.cli"] + list(argv), env=env))
    rcs = [p.wait() for p in procs]
    return max(abs(rc) for rc in rcs)


def init_process_group(use_cuda: bool, local_rank: int):
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return
    if use_cuda and torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("gloo")


def gather_counts(local_counts: np.ndarray, device=None) -> np.ndarray:
    """Sum of the per-bin read counts of every rank (all_reduce of n_bins int64 counters)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local_counts, dtype=np.int64).copy()
    t = torch.as_tensor(np.asarray(local_counts, dtype=np.int64))
    if device is None and dist.get_backend() == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def merge_in_order(parts: Dict[int, Iterable[bytes]]) -> bytes:
    """Concatenate per-batch byte strings of one bin in batch-id order."""
    return b"".join(b"".join(parts[k]) if not isinstance(parts[k], (bytes, bytearray)) else parts[k]
                    for k in sorted(parts))


def part_path(path: str, rank: int) -> str:
    """Rank r's part of a bin file; keeps a .gz suffix (the writer compresses by suffix)."""
    if path.endswith(".gz"):
        return "%s.part%d.gz" % (path[:-3], rank)
    return "%s.part%d" % (path, rank)


def merge_part_files(paths: Sequence[Optional[str]], world: int, level: int = 1, keep_parts: bool = False):
    """Stitch PATH.part<r> (r < world) into PATH for every bin path: rank r's k-th write holds batch r + k * world
    (PATH.part<r>.idx lists its chunks as (ticket, bytes) in file order), and PATH receives the chunks in batch
    order.  gzip members are independent, so for .gz files this is a copy of byte ranges."""
    empty = None
    for path in paths:
        if path is None:
            continue
        pieces = []                                     # (batch id, order inside the batch, rank, offset, bytes)
        for r in range(world):
            idx = np.fromfile(part_path(path, r) + ".idx", dtype="<u8").reshape(-1, 2)
            off = 0
            for j, (ticket, nbytes) in enumerate(idx.tolist()):
                pieces.append((r + ticket * world, j, r, off, nbytes))
                off += nbytes
        pieces.sort()
        fds = {}
        with open(path, "wb") as out:
            for _, _, r, off, nbytes in pieces:
                if r not in fds:
                    fds[r] = open(part_path(path, r), "rb")
                fh = fds[r]
                fh.seek(off)
                left = nbytes
                while left:
                    buf = fh.read(min(left, 8 << 20))
                    if not buf:
                        raise OSError("%s is shorter than its index says" % part_path(path, r))
                    out.write(buf)
                    left -= len(buf)
            if not pieces and path.endswith(".gz"):     # a bin without reads: an empty but valid .gz
                if empty is None:
                    from . import lib as _lib
                    buf = (C.c_uint8 * 64)()
                    n = _lib.load().orc_empty_gzip_member(buf, 64, level)
                    empty = bytes(buf[:n])
                out.write(empty)
        for fh in fds.values():
            fh.close()
        if not keep_parts:
            for r in range(world):
                for p in (part_path(path, r), part_path(path, r) + ".idx"):
                    try:
                        os.unlink(p)
                    except OSError:
                        pass
