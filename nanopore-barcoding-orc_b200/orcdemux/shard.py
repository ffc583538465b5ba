"""Multi-GPU host logic: reads are independent, so batches are dealt round-robin to the
ranks (one process per GPU) with no collective on the data path; the only exchange is the
final per-bin count gather.  Output order inside a bin is restored from the
batch ids (cutadapt's ordered chunk writer, runners.py, keeps input order)."""
from __future__ import annotations

from typing import Dict, Iterable, List

import numpy as np


def batches_of_rank(n_batches: int, rank: int, world: int) -> List[int]:
    """Batch ids dealt to `rank`: rank, rank + world, ..."""
    return list(range(rank, n_batches, world))


def owner_of_batch(batch_id: int, world: int) -> int:
    return batch_id % world


def gather_counts(local_counts: np.ndarray, device=None) -> np.ndarray:
    """Sum of the per-bin read counts of every rank (all_reduce of n_bins int64 counters)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local_counts, dtype=np.int64).copy()
    t = torch.as_tensor(np.asarray(local_counts, dtype=np.int64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def merge_in_order(parts: Dict[int, Iterable[bytes]]) -> bytes:
    """Concatenate per-batch byte strings of one bin in batch-id order."""
    return b"".join(b"".join(parts[k]) if not isinstance(parts[k], (bytes, bytearray)) else parts[k]
                    for k in sorted(parts))
