"""BASELINE configs[4], literally: batch-sharded free-running decode of 8192 synthetic utterances (150 tokens, Max_Step 1000) across
the ranks of one box through gst_tacotron_b200.shard.decode_sharded - every rank decodes its contiguous slice on its own GPU
(chunks of 256 utterances), the mels / stop logits are gathered on the HOST of rank 0 (gloo, no collective on the data path).
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_configs4.py [N_UTT]
Prints one JSON line on rank 0: frames/s of the decode alone (device time, max over ranks) and of the whole job incl. the host gather."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from gst_tacotron_b200.hparams import load_config
from gst_tacotron_b200.runtime import Engine
from gst_tacotron_b200.shard import gather_host, shard_range
from gst_tacotron_b200.weights import init_weights

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("gloo")
torch.cuda.set_device(local)
cfg = load_config(precision="bf16")
eng = Engine(cfg, init_weights(cfg, bias_scale=0.05), device=local)
T, Tv = cfg.max_step // cfg.step_reduction, 150
a, b = shard_range(N, world, rank)
rng = np.random.default_rng(100 + rank)
text = rng.uniform(-1, 1, (b - a, Tv, cfg.text_dim)).astype(np.float32)   # this rank's slice of the job (synthetic, generated locally)
gst = rng.uniform(-1, 1, (b - a, cfg.style_size)).astype(np.float32)
text_p, gst_p = torch.from_numpy(text).pin_memory(), torch.from_numpy(gst).pin_memory()
from gst_tacotron_b200.shard import SharedHostArray
GATHER = os.environ.get("GSTK_GATHER", "shm" if world > 1 else "none")   # shm: every rank's D2H lands in shared pages rank 0 reads
if GATHER == "shm":
    box = ["gstk_c4_{}_{}".format(os.getpid(), int.from_bytes(os.urandom(4), "little"))] if rank == 0 else [None]
    dist.broadcast_object_list(box, src=0)
    sh = {}
    if rank == 0:
        sh = {"mel": SharedHostArray(box[0] + "_mel", (N, T, cfg.mel_dim), True), "stop": SharedHostArray(box[0] + "_stop", (N, T), True)}
    dist.barrier()
    if rank != 0:
        sh = {"mel": SharedHostArray(box[0] + "_mel", (N, T, cfg.mel_dim), False), "stop": SharedHostArray(box[0] + "_stop", (N, T), False)}
    mel, stop = sh["mel"].tensor[a:b], sh["stop"].tensor[a:b]
else:
    mel = torch.empty(b - a, T, cfg.mel_dim).pin_memory()
    stop = torch.empty(b - a, T).pin_memory()
import ctypes as C
from gst_tacotron_b200 import _lib


def decode_slice():
    for c0 in range(0, b - a, 256):            # the library takes any batch; explicit chunks keep the pinned views contiguous
        c1 = min(c0 + 256, b - a)
        da = _lib.GstkDecodeArgs()
        da.batch, da.key_time, da.steps, da.mode, da.rng_mode = c1 - c0, Tv, T, _lib.MODE_FREE, _lib.RNG["philox"]
        da.seed, da.row_offset = 7, a + c0
        da.enc_text, da.gst = text_p[c0:c1].data_ptr(), gst_p[c0:c1].data_ptr()
        da.out_mel, da.out_stop = mel[c0:c1].data_ptr(), stop[c0:c1].data_ptr()
        da.stream = torch.cuda.current_stream().cuda_stream
        _lib.raise_for(eng._lib.gstk_decode(eng._h, C.byref(da)), eng._h)


decode_slice()   # warm-up (weight images, staging buffers)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
decode_slice()
torch.cuda.synchronize()
t_dec = time.perf_counter() - t0
if GATHER == "shm":
    dist.barrier()
    res_mel, res_stop = (sh["mel"].tensor.numpy(), sh["stop"].tensor.numpy()) if rank == 0 else (None, None)
else:
    res_mel = gather_host(mel.numpy(), N) if world > 1 else mel.numpy()
    res_stop = gather_host(stop.numpy(), N) if world > 1 else stop.numpy()
t_all = time.perf_counter() - t0
tt = torch.tensor([t_dec, t_all], dtype=torch.float64)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    assert res_mel.shape == (N, T, cfg.mel_dim) and res_stop.shape == (N, T) and np.isfinite(res_mel).all()
    print(json.dumps({"workload": "configs[4]: {} utterances x {} steps, T_v {}, {} rank(s), host gather of mel + stop to rank 0".format(N, T, Tv, world),
                      "n_gpus": world, "decode_s": float(tt[0]), "job_s_incl_host_gather": float(tt[1]),
                      "mel_frames_per_s_decode": N * T / float(tt[0]), "mel_frames_per_s_job": N * T / float(tt[1]),
                      "gather": GATHER, "gathered_bytes": int(res_mel.nbytes + res_stop.nbytes)}))
if GATHER == "shm":
    for v in sh.values():
        v.close(unlink=rank == 0)
eng.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
