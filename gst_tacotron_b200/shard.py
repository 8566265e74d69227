"""Utterance sharding across the GPUs of one box.

Each utterance's decode depends only on its own encodings and state, so the batch is split contiguously over
the ranks (one process per GPU, weights replicated) and there is NO collective on the data path: every rank
decodes its slice with ``row_offset = start`` (so the counter-based dropout / noise streams are those of the
unsharded batch) and the results are gathered on the host of rank 0."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of `n` utterances for `rank` (first n % world ranks get one more)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_host(local: np.ndarray, n_total: int, group=None) -> Optional[np.ndarray]:
    """Host gather of per-rank row blocks (in rank order) onto rank 0; other ranks get None."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    pieces = [None] * world if rank == 0 else None
    dist.gather_object(np.ascontiguousarray(local), pieces, dst=0, group=group)
    if rank != 0:
        return None
    out = np.concatenate(pieces, axis=0)
    assert out.shape[0] == n_total
    return out


def decode_sharded(engine, enc_text, gst, steps: int, seed: int = 0, want=("mel", "stop"), group=None) -> Optional[Dict[str, np.ndarray]]:
    """Decode the full utterance list [N, ...] given on every rank: this rank runs rows [start, stop) on its GPU and
    rank 0 returns the gathered host arrays."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = int(np.shape(enc_text)[0])
    a, b = shard_range(n, world, rank)
    if n < world:
        # fewer utterances than ranks would leave a rank with an empty slice: its decode would raise while the others
        # already wait in the gather.  Every rank sees the same n, so all of them raise here, before any collective.
        raise ValueError("decode_sharded: {} utterances cannot be split over {} ranks".format(n, world))
    out = engine.decode(enc_text=np.ascontiguousarray(np.asarray(enc_text)[a:b]), gst=np.ascontiguousarray(np.asarray(gst)[a:b]),
                        steps=steps, rng="philox", seed=seed, row_offset=a, want=want, host_outputs=True)
    res = {k: gather_host(np.asarray(v), n, group) for k, v in out.items()}
    return res if rank == 0 else None
