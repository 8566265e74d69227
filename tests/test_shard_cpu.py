"""Host-side sharding logic with a real 2-process gloo group on CPU (no GPU, no data-path collective)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from gst_tacotron_b200.shard import gather_host, shard_range
from oracle import reference_port as O
from tests.util import make_cfg


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 8192):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeEngine:
    """Stands in for the CUDA engine on the CPU tier: 'decodes' with the oracle's philox stream so that the
    row_offset contract can be checked end to end."""

    def __init__(self, cfg):
        self.cfg = cfg

    def decode(self, enc_text, gst, steps, rng, seed, row_offset, want, host_outputs, kernel="auto"):
        B, Tv = enc_text.shape[0], enc_text.shape[1]
        k0, _, nz = O.philox_randomness(self.cfg, seed, steps, B, Tv, b0=row_offset)
        return {"mel": np.transpose(k0[:, :, :80], (1, 0, 2)) + enc_text[:, :1, :1], "stop": np.transpose(nz[:, :, 0], (1, 0))}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gst_tacotron_b200.shard import decode_sharded
    cfg = make_cfg()
    n, Tv, T = 7, 5, 3
    rng = np.random.default_rng(0)
    text = rng.standard_normal((n, Tv, 4)).astype(np.float32)
    gst = np.zeros((n, 2), np.float32)
    res = decode_sharded(_FakeEngine(cfg), text, gst, steps=T, seed=11)
    local = np.full((shard_range(n, world, rank)[1] - shard_range(n, world, rank)[0], 2), rank, np.float32)
    g = gather_host(local, n)
    if rank == 0:
        full = _FakeEngine(cfg).decode(text, gst, T, "philox", 11, 0, ("mel", "stop"), True)
        q.put((np.array_equal(res["mel"], full["mel"]) and np.array_equal(res["stop"], full["stop"]), g[:, 0].tolist()))
    else:
        assert res is None and g is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_decode_matches_unsharded():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    same, order = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert same
    assert order == [0, 0, 0, 0, 1, 1, 1]
