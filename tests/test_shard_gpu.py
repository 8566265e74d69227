"""Sharded decode on REAL engines (BASELINE configs[4]): two processes, one Engine each (on GPU rank % device_count,
so the test also runs on a one-GPU box), gloo group for the host gather - sharded result == unsharded result, bit for bit (the
Philox streams are keyed by the global row index via row_offset)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gst_tacotron_b200.runtime import Engine
        from gst_tacotron_b200.shard import decode_sharded
        from tests.util import make_cfg, make_weights
        cfg = make_cfg("SMA", precision="bf16")
        eng = Engine(cfg, make_weights(cfg), device=rank % torch.cuda.device_count())
        Tv, T = 33, 7
        rng = np.random.default_rng(0)
        bad = []
        # 20 utterances: the slices (10 + 10) would fit the small-batch kernel, the unsharded decode does not - decode_sharded pins the
        # batch-256 kernel; 11 and 5 utterances: whole and slices all run the small-batch kernel (11: two n-tiles whole, one per slice)
        # (the repeats: a dropped result must release its page lock before its mapping - a stale registration would swallow the next copies)
        for n, how in ((20, "shm"), (11, "send"), (5, "shm"), (5, "send"), (5, "shm"), (11, "shm"), (20, "send")):   # shm: arrays written by every rank's own D2H copies | send/recv to rank 0
            text = rng.uniform(-1, 1, (n, Tv, cfg.text_dim)).astype(np.float32)
            gst = rng.uniform(-1, 1, (n, cfg.style_size)).astype(np.float32)
            res = decode_sharded(eng, text, gst, steps=T, seed=11, gather=how)
            if rank == 0:
                full = eng.decode(enc_text=text, gst=gst, steps=T, rng="philox", seed=11, want=("mel", "stop"), host_outputs=True)
                if not (np.array_equal(res["mel"], np.asarray(full["mel"])) and np.array_equal(res["stop"], np.asarray(full["stop"]))):
                    bad.append((n, how, float(np.abs(res["mel"] - np.asarray(full["mel"])).max())))
            else:
                assert res is None
        if rank == 0:
            q.put(bad)
        dist.barrier()
        eng.close()
    finally:
        dist.destroy_process_group()


def test_two_process_sharded_decode_equals_unsharded_on_real_engines():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    bad = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert bad == []
