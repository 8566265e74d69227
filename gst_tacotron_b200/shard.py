"""Utterance sharding across the GPUs of one box.

Each utterance's decode depends only on its own encodings and state, so the batch is split contiguously over
the ranks (one process per GPU, weights replicated) and there is NO collective on the data path: every rank
decodes its slice with ``row_offset = start`` (so the counter-based dropout / noise streams are those of the
unsharded batch) and the results are gathered on the host of rank 0."""
from __future__ import annotations

import os
import weakref
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
    """Host gather of per-rank row blocks (in rank order) onto rank 0; other ranks get None.

    On a gloo group the rows travel as plain CPU tensors (``dist.gather`` into views of the output array, no pickling - the
    gathered mels of an 8192-utterance job are 2.6 GB); on any other backend as pickled objects."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local = np.ascontiguousarray(local)
    if dist.get_backend(group) == "gloo":
        tail = local.shape[1:]
        per = int(np.prod(tail)) if tail else 1
        rows = [shard_range(n_total, world, r) for r in range(world)]
        if rank == 0:
            out = np.empty((n_total,) + tail, local.dtype)
            out[rows[0][0]:rows[0][1]] = local
            views = [torch.from_numpy(out[a:b].reshape(-1)) for a, b in rows]
            reqs = [dist.irecv(views[r], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
                    for r in range(1, world) if rows[r][1] > rows[r][0]]
            for q in reqs:
                q.wait()
            return out
        if local.size:
            assert local.size == (rows[rank][1] - rows[rank][0]) * per, "gather_host: this rank's block does not match shard_range"
            dist.send(torch.from_numpy(local.reshape(-1)), dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return None
    pieces = [None] * world if rank == 0 else None
    dist.gather_object(local, pieces, dst=0, group=group)
    if rank != 0:
        return None
    out = np.concatenate(pieces, axis=0)
    assert out.shape[0] == n_total
    return out


class SharedHostArray:
    """float32 array [rows, ...] in POSIX shared memory (/dev/shm), page-locked in THIS process: every rank of the box maps the
    same pages, so each rank's device->host copy of its rows lands directly in the array rank 0 reads - the 'final host
    gather' of the sharded decode costs no extra copy.  Rank 0 creates the file, everybody else opens it after a barrier."""

    def __init__(self, name: str, shape, create: bool):
        self.path = os.path.join("/dev/shm", name)
        self.shape = tuple(int(x) for x in shape)
        n = int(np.prod(self.shape))
        if create:
            with open(self.path, "wb") as f:
                f.truncate(n * 4)
        self.tensor = torch.from_file(self.path, shared=True, size=n, dtype=torch.float32).view(self.shape)
        self._release = None
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), n * 4, 0)
            if int(rc) != 0:
                raise RuntimeError("cudaHostRegister of {} failed ({})".format(self.path, rc))
            # The page lock MUST be dropped before the mapping goes away: a registration that outlives its mapping keeps pointing
            # at the old pages, and the next array mapped at the same address would receive the device's copies there instead of
            # in its own pages.  The finalizer holds the mapping (the tensor) until it has run, whoever drops the array and when.
            self._release = weakref.finalize(self, _unregister, self.tensor)

    def close(self, unlink: bool):
        if self._release is not None:
            self._release()
            self._release = None
        self.tensor = None
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def _unregister(tensor) -> None:
    torch.cuda.cudart().cudaHostUnregister(tensor.data_ptr())


def decode_sharded(engine, enc_text, gst, steps: int, seed: int = 0, want=("mel", "stop"), group=None,
                   gather: str = "auto", kernel: Optional[str] = None) -> Optional[Dict[str, np.ndarray]]:
    """Decode the full utterance list [N, ...] given on every rank: this rank runs rows [start, stop) on its GPU and
    rank 0 returns the gathered host arrays (with gather="shm" they are views of shared pages kept alive by the extra
    "_shared" entry of the dict).

    gather = "shm": the ranks share one box - the outputs are written by every rank's own device->host copies into arrays in
    shared memory (no gather traffic at all); "send": per-rank host buffers + send/recv to rank 0; "auto": "shm" when
    LOCAL_WORLD_SIZE == WORLD_SIZE and the engine is a real one (it accepts ``out_buffers``).

    kernel: Engine.decode's kernel selector.  None = whatever the UNSHARDED decode of the same list would take ("batch" for more
    than 16 utterances, "auto" otherwise), so that the gathered result is bit-identical to the unsharded one even when a rank's
    slice is small enough for the small-batch kernel (whose rows do not depend on how many utterances share a launch)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = int(np.shape(enc_text)[0])
    a, b = shard_range(n, world, rank)
    if n < world:
        # fewer utterances than ranks would leave a rank with an empty slice: its decode would raise while the others
        # already wait in the gather.  Every rank sees the same n, so all of them raise here, before any collective.
        raise ValueError("decode_sharded: {} utterances cannot be split over {} ranks".format(n, world))
    if kernel is None:
        kernel = "batch" if n > 16 else "auto"
    text_l = np.ascontiguousarray(np.asarray(enc_text)[a:b])
    gst_l = np.ascontiguousarray(np.asarray(gst)[a:b])
    one_box = os.environ.get("LOCAL_WORLD_SIZE", str(world)) == str(world)
    if gather == "auto":
        gather = "shm" if (world > 1 and one_box and hasattr(engine, "cfg") and hasattr(engine, "_h")) else "send"
    if gather == "shm" and world > 1:
        cfg = engine.cfg
        shapes = {"mel": (n, steps * cfg.step_reduction, cfg.mel_dim), "stop": (n, steps), "alignment": (n, steps, int(np.shape(enc_text)[1]))}
        box = [("gstk_{}_{}".format(os.getpid(), int.from_bytes(os.urandom(4), "little")))] if rank == 0 else [None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        tag = box[0]
        arrs = {}
        for k in want:
            if rank == 0:
                arrs[k] = SharedHostArray("{}_{}".format(tag, k), shapes[k], create=True)
        dist.barrier(group)
        for k in want:
            if rank != 0:
                arrs[k] = SharedHostArray("{}_{}".format(tag, k), shapes[k], create=False)
        engine.decode(enc_text=text_l, gst=gst_l, steps=steps, rng="philox", seed=seed, row_offset=a, want=want, host_outputs=True,
                      out_buffers={k: arrs[k].tensor[a:b] for k in want}, kernel=kernel)
        dist.barrier(group)   # every rank's rows are in the shared arrays
        if rank != 0:
            for k in want:
                arrs[k].close(unlink=False)
            return None
        res = {k: arrs[k].tensor.numpy() for k in want}   # views of the shared pages: no copy
        for k in want:   # the names can go; the mappings (and the page lock) live as long as the arrays are referenced
            try:
                os.unlink(arrs[k].path)
            except OSError:
                pass
        res["_shared"] = arrs
        return res
    out = engine.decode(enc_text=text_l, gst=gst_l, steps=steps, rng="philox", seed=seed, row_offset=a, want=want, host_outputs=True,
                        **({} if kernel == "auto" else {"kernel": kernel}))
    res = {k: gather_host(np.asarray(v), n, group) for k, v in out.items()}
    return res if rank == 0 else None
