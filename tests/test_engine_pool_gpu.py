"""runtime.EnginePool: two engines on one device, requests alternate between them (the serving loop bench.py's e2e figure runs
through).  Results must be those of a single engine, whatever the interleaving, and come back in submission order."""
import numpy as np
import pytest
import torch

from tests.util import make_cfg, make_weights

pytestmark = pytest.mark.gpu


def test_pool_results_equal_single_engine():
    from gst_tacotron_b200.runtime import Engine, EnginePool
    cfg = make_cfg("SMA", precision="bf16")
    W = make_weights(cfg)
    rng = np.random.default_rng(0)
    reqs = [(rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32), T, seed)
            for seed, (B, Tv, T) in enumerate([(12, 40, 300), (3, 21, 9), (12, 40, 300), (20, 64, 260), (1, 82, 40), (12, 40, 300)])]
    single = Engine(cfg, W)
    want = [single.decode(encodings=x, steps=T, rng="philox", seed=s, want=("mel", "stop")) for x, T, s in reqs]
    single.close()
    with EnginePool(cfg, W, device=0, depth=2) as pool:
        futs = [pool.submit(lambda e, x=x, T=T, s=s: e.decode(encodings=x, steps=T, rng="philox", seed=s, want=("mel", "stop")))
                for x, T, s in reqs]
        got = [f.result(timeout=300) for f in futs]
        assert len({id(e) for e in pool.engines}) == 2
        for a, b in zip(got, want):
            assert np.array_equal(a["mel"], b["mel"]) and np.array_equal(a["stop"], b["stop"])
        # device tensors in, device tensors out: the result lives on the engine's stream; synchronize() before reading elsewhere
        xd = torch.as_tensor(reqs[1][0], device="cuda:0")
        f = pool.submit(lambda e: e.decode(encodings=xd, steps=9, rng="philox", seed=1, want=("mel",)))
        out = f.result(timeout=120)
        pool.synchronize()
        assert np.array_equal(out["mel"].cpu().numpy(), want[1]["mel"])
        with pytest.raises(Exception):      # an error inside the worker comes back through the future
            pool.submit(lambda e: e.decode(encodings=np.zeros((1, 4, 3), np.float32), steps=2)).result(timeout=60)
