#!/usr/bin/env python
"""bench.py - mel frames/s of the batched GST-Tacotron2 decode hot path on N B200s.

One "step" = one pass of the hot path over one synthetic batch per GPU:
    Style_Token_Layer (GST front end) on B reference mels  ->  free-running Tacotron2 decoder loop,
    B utterances x (Max_Step // Step_Reduction) decoder steps, GST concat folded into the value projection.
Workload at every N: BASELINE.json configs[2] per GPU (free-running decode, batch 256, T_v 150,
Max_Step 1000, SMA); utterances are sharded across ranks with no collective ("weak" scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

Prints ONE JSON line (rank 0).  `value` = whole-job mel frames/s with inputs resident in HBM;
`e2e` = the same through the C ABI with pinned HOST buffers (H2D of the inputs and D2H of mel / stop /
alignment inside the timed region); `roofline` = the persistent decoder kernel against the measured
bf16 tensor peak; `cpu_baseline` = the CPU oracle port of the reference (torch fp32, all host threads,
value projection recomputed every step like Steps.py:123) on a bounded sample.
`--impl reference` times that CPU port as the reference arm (TensorFlow cannot be installed here,
DESIGN.md section "Reference arm").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "mel frames/sec, batched GST-Tacotron2 decode"
UNIT = "mel_frames/s"
B_DEC, TV, REF_FRAMES = 256, 150, 188
FLOP_PER_STEP_UTT = 2 * (14367872 + 128 * TV) + 4 * 128 * TV + 22000  # 28.87 MFLOP at T_v=150


def load_traffic():
    """DRAM bytes (read + write) of one persistent-decoder launch of this workload, from the newest committed `ncu --set full`
    capture of this same command (profiles/r*_ncu_traffic.json, written by tools/ncu_traffic.py).  The capture is stamped with
    the SHA-1 of the kernel's source file: `stale` = the kernel has changed since (the number then describes an older kernel).
    Returns (bytes or None, stale flag, file name)."""
    import glob
    import hashlib
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    if not files:
        return None, None, None
    p = files[-1]
    try:
        d = json.load(open(p))
        src = os.path.join(ROOT, "gst_tacotron_b200", "csrc", "decoder_bf16.cuh")
        sha = hashlib.sha1(open(src, "rb").read()).hexdigest()
        return float(d["dram_bytes_per_launch"]), d.get("kernel_source_sha1") != sha, os.path.basename(p)
    except Exception:
        return None, None, os.path.basename(p)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = "/tmp/gstk_clocks_{}.csv".format(os.getpid())

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = smax
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def synth_batch(cfg, seed):
    rng = np.random.default_rng(seed)
    text = rng.uniform(-1.0, 1.0, size=(B_DEC, TV, cfg.text_dim)).astype(np.float32)
    mels = rng.uniform(-4.0, 4.0, size=(B_DEC, REF_FRAMES + 1, cfg.mel_dim)).astype(np.float32)
    mels[:, 0] = 0.0
    lens = rng.integers(REF_FRAMES // 2, REF_FRAMES + 1, size=(B_DEC,)).astype(np.int32)
    return text, mels, lens


def run_reference(args, rank, world):
    """Reference arm: the CPU port of the reference's own path (oracle/reference_port.py, the one place
    outside tests where bench.py may execute oracle code) on all host threads."""
    if rank != 0:
        return
    from gst_tacotron_b200.hparams import load_config
    from gst_tacotron_b200.weights import init_weights
    from oracle import reference_port as O
    cfg = load_config()
    W = O.to_torch(init_weights(cfg, bias_scale=0.05), torch.float32)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_steps = args.ref_decoder_steps
    text, mels, lens = synth_batch(cfg, 2024)
    k0, k1, nz = O.philox_randomness(cfg, 7, sample_steps, B_DEC, TV)

    def one():
        g = O.style_token_layer(W, cfg, mels, lens, dtype=torch.float32)
        enc = O.gst_concat(torch.as_tensor(text), g)
        O.decoder_loop(W, cfg, enc, steps=sample_steps, keep0=k0, keep1=k1, noise=nz, dtype=torch.float32,
                       reproject_every_step=True)

    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = (time.perf_counter() - t0) / args.steps
    val = B_DEC * sample_steps * cfg.step_reduction / dt
    sample = "B={} x {} decoder steps (of {}) + GST on {} ref frames per step".format(
        B_DEC, sample_steps, cfg.max_step // cfg.step_reduction, REF_FRAMES)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args, "fp32"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(cfg, args, precision):
    return {"workload": "configs[2]: free-running decode batch={} per GPU, T_v={}, {} decoder steps, attention={}, "
                        "+ GST front end on {}-frame reference mels".format(
                            B_DEC, TV, cfg.max_step // cfg.step_reduction, cfg.attention_type, REF_FRAMES),
            "batch_per_gpu": B_DEC, "key_time": TV, "decoder_steps": cfg.max_step // cfg.step_reduction,
            "attention": cfg.attention_type, "rng": "philox (in-kernel dropout masks + sigmoid noise)",
            "precision": precision, "sharding": "utterances over ranks, no collective",
            "l2": "per-step inputs+outputs 335 MB > 126 MB L2 (no explicit flush)"}


def cpu_baseline(cfg, W_np, budget_s=12.0):
    from oracle import reference_port as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    W = O.to_torch(W_np, torch.float32)
    text, mels, lens = synth_batch(cfg, 2024)
    n = 10
    k0, k1, nz = O.philox_randomness(cfg, 7, n, B_DEC, TV)
    enc = O.gst_concat(torch.as_tensor(text), O.style_token_layer(W, cfg, mels, lens, dtype=torch.float32))
    O.decoder_loop(W, cfg, enc, steps=2, keep0=k0, keep1=k1, noise=nz, dtype=torch.float32, reproject_every_step=True)
    done, t0 = 0, time.perf_counter()
    while True:
        O.decoder_loop(W, cfg, enc, steps=n, keep0=k0, keep1=k1, noise=nz, dtype=torch.float32,
                       reproject_every_step=True)
        done += n
        if time.perf_counter() - t0 > budget_s or done >= 400:
            break
    dt = time.perf_counter() - t0
    return {"value": B_DEC * done * cfg.step_reduction / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "B={} x {} decoder steps, torch-CPU fp32 port of the reference (value projection every step)".format(
                B_DEC, done)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("GSTK_PRECISION", "auto"), choices=["auto", "bf16", "fp32"])
    ap.add_argument("--ref-decoder-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e through one engine, one step at a time (no EnginePool)")
    ap.add_argument("--no-extras", action="store_true", help="skip the GST configs[3] and early-stop blocks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gst_tacotron_b200 has no CPU fallback")
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on the C-level stdout when the communicator is created; the contract is ONE JSON
        # line on stdout, so fd 1 points at stderr until the first collective has run
        import ctypes
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            try:
                ctypes.CDLL(None).fflush(None)
            except Exception:
                pass
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    from gst_tacotron_b200.hparams import load_config
    from gst_tacotron_b200.runtime import Engine
    from gst_tacotron_b200.weights import init_encoder_weights, init_postnet_weights, init_vocoder_weights, init_weights
    from gst_tacotron_b200 import build as _b
    precision = args.precision
    if precision == "auto":
        precision = "bf16" if _b.BF16_READY else "fp32"
    cfg = load_config(precision=precision)
    W = init_weights(cfg, bias_scale=0.05)
    W_all = dict(W)
    W_all.update(init_postnet_weights(cfg))   # own generator: the decode pack is unchanged
    W_all.update(init_encoder_weights(cfg))
    W_all.update(init_vocoder_weights(cfg))
    eng = Engine(cfg, W_all, device=local_rank)
    dev = torch.device("cuda", local_rank)
    T = cfg.max_step // cfg.step_reduction
    frames_per_step = B_DEC * T * cfg.step_reduction

    text, mels, lens = synth_batch(cfg, 2024 + rank)
    text_d, mels_d, lens_d = torch.as_tensor(text, device=dev), torch.as_tensor(mels, device=dev), torch.as_tensor(lens, device=dev)
    text_h, mels_h = torch.as_tensor(text).pin_memory(), torch.as_tensor(mels).pin_memory()
    lens_h = torch.as_tensor(lens).pin_memory()
    out_h = {"mel": torch.empty(B_DEC, T * cfg.step_reduction, cfg.mel_dim).pin_memory(),
             "stop": torch.empty(B_DEC, T).pin_memory(), "alignment": torch.empty(B_DEC, T, TV).pin_memory()}
    gst_dev = torch.empty(B_DEC, cfg.style_size, device=dev)
    import ctypes as C
    from gst_tacotron_b200 import _lib

    def step_device(i):
        g = eng.gst(mels_d, lens_d, want=("gst",), host_outputs=False)["gst"]
        out = eng.decode(enc_text=text_d, gst=g, steps=T, rng="philox", seed=1000 + i, row_offset=rank * B_DEC,
                         host_outputs=False)
        return out

    def step_host(i, eng=eng, out_h=out_h, gst_h=gst_dev):
        # public API calls with HOST buffers: the library stages H2D / D2H itself (pinned => async copies).  The style embedding is an
        # intermediate: it stays on the device between the two calls (no host round trip, no synchronisation between them), the
        # step's results - mel, stop, alignment - come back to pinned host memory.
        ga = _lib.GstkGstArgs()
        ga.batch, ga.frames, ga.drop_first = B_DEC, REF_FRAMES + 1, 1
        ga.mels, ga.lengths, ga.out_gst = mels_h.data_ptr(), lens_h.data_ptr(), gst_h.data_ptr()
        ga.stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.raise_for(eng._lib.gstk_gst(eng._h, C.byref(ga)), eng._h)
        da = _lib.GstkDecodeArgs()
        da.batch, da.key_time, da.steps, da.mode, da.rng_mode = B_DEC, TV, T, _lib.MODE_FREE, _lib.RNG["philox"]
        da.seed, da.row_offset = 1000 + i, rank * B_DEC
        da.enc_text, da.gst = text_h.data_ptr(), gst_h.data_ptr()
        da.out_mel, da.out_stop, da.out_alignment = out_h["mel"].data_ptr(), out_h["stop"].data_ptr(), out_h["alignment"].data_ptr()
        da.stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.raise_for(eng._lib.gstk_decode(eng._h, C.byref(da)), eng._h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn):
        for i in range(args.warmup):
            fn(i)
        barrier()
        l0 = eng.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            fn(args.warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), eng.launch_count - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, launches = timed(step_device)
    clocks = sampler.stop() if rank == 0 else None
    e2e_serial_ms, _ = timed(step_host)
    # The serving loop (gst_tacotron_b200.runtime.EnginePool): two engines on this GPU, consecutive steps alternate between them,
    # each step still copies ITS inputs from pinned host memory and ITS results back to its own pinned host buffers inside the
    # timed region - the copies of one step overlap the kernels of the other.
    from gst_tacotron_b200.runtime import EnginePool
    e2e_depth = 1 if args.no_pipeline else 2
    e2e_ms = e2e_serial_ms
    if e2e_depth > 1:
        pool = EnginePool(cfg, W_all, device=local_rank, depth=e2e_depth)
        bufs = [({"mel": torch.empty_like(out_h["mel"]).pin_memory(), "stop": torch.empty_like(out_h["stop"]).pin_memory(),
                  "alignment": torch.empty_like(out_h["alignment"]).pin_memory()}, torch.empty_like(gst_dev))
                for _ in range(e2e_depth)]
        pending = []

        def step_pool(i):
            k = i % e2e_depth
            if len(pending) >= e2e_depth:        # the buffers of engine k are free again once its previous step has returned
                pending.pop(0).result()
            pending.append(pool.submit(lambda e, i=i, k=k: step_host(i, e, bufs[k][0], bufs[k][1]), engine=k))

        def timed_pool():
            for i in range(args.warmup):
                step_pool(i)
            while pending:
                pending.pop(0).result()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                step_pool(args.warmup + i)
            while pending:
                pending.pop(0).result()          # every step's results are in its pinned host buffers
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        e2e_ms = timed_pool()
        last = bufs[(args.warmup + args.steps - 1) % e2e_depth][0]
        assert torch.isfinite(last["mel"]).all() and float(last["alignment"].sum()) > 0     # the pooled steps did produce outputs
        pool.close()
        del bufs

    # dominant kernel (persistent decoder): live CUDA-event duration, measured by the library around the
    # cooperative launch on the launching stream
    dk = []
    for i in range(3):
        step_device(100 + i)
        dk.append(eng.last_kernel_ms())
    dec_ms = float(np.median(dk))

    # Postnet (Taco2.py:230) on the decode's own output: reported next to the headline, NOT part of
    # `value` / `e2e` (the metric is the decoder loop + GST front end)
    post = None
    if rank == 0:
        mel_d = step_device(200)["mel"]
        pk = []
        for i in range(5):
            eng.postnet(mel_d)
            pk.append(eng.last_kernel_ms())
        pms = float(np.median(pk[2:]))
        cin = [cfg.mel_dim] + [l[0] for l in cfg.postnet_layers[:-1]]
        pflops = 2.0 * B_DEC * T * cfg.step_reduction * sum(k * ci * co for (co, k, _s, _t), ci in zip(cfg.postnet_layers, cin))
        post = {"ms": pms, "frames_per_s": frames_per_step / (pms * 1e-3), "achieved_tflops": pflops / (pms * 1e-3) / 1e12,
                "algorithmic_flops": pflops, "in_timed_region": False}

    # text Encoder (Taco2.py:12-51) on B_DEC x TV tokens: also reported next to the headline only
    enc_blk = None
    if rank == 0:
        tok = torch.randint(0, cfg.vocab_size, (B_DEC, TV), device=dev, dtype=torch.int32)
        ek = []
        for i in range(5):
            eng.encoder(tok)
            ek.append(eng.last_kernel_ms())
        enc_blk = {"ms": float(np.median(ek[2:])), "tokens": B_DEC * TV, "in_timed_region": False}
        # whole Inference model (Model.py:108-129): Encoder -> GST -> decoder loop -> Postnet -> Vocoder_Taco1, device-resident
        fk = []
        for i in range(4):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            eng.inference(tok, mels_d, lens_d, steps=T, rng="philox", seed=300 + i, host_outputs=False)
            torch.cuda.synchronize(dev)
            fk.append((time.perf_counter() - t0) * 1e3)
        enc_blk["full_inference_ms"] = float(np.median(fk[1:]))
        enc_blk["full_inference_frames_per_s"] = frames_per_step / (enc_blk["full_inference_ms"] * 1e-3)

    # wav side: Vocoder_Taco1 (Taco2.py:234-260) on the Postnet output of the decode above, and what
    # Export_Inference does with its spectrogram (Model.py:412-420): Griffin-Lim, 60 iterations, for all B_DEC utterances at once
    voc_blk = None
    if rank == 0 and not args.no_extras:
        post_d = eng.postnet(mel_d)
        vk = []
        for i in range(4):
            spec_d = eng.vocoder(post_d)
            vk.append(eng.last_kernel_ms())
        ln_d = torch.full((B_DEC,), T * cfg.step_reduction, dtype=torch.int32, device=dev)
        gk = []
        for i in range(3):
            wav_d = eng.griffin_lim(spec_d, lengths=ln_d, rng="philox", seed=i, max_abs_value=cfg.max_abs_mel)
            gk.append(eng.last_kernel_ms())
        vms, gms = float(np.median(vk[1:])), float(np.median(gk[1:]))
        voc_blk = {"vocoder_ms": vms, "vocoder_frames_per_s": frames_per_step / (vms * 1e-3),
                   "griffin_lim_ms": gms, "griffin_lim_iters": cfg.griffin_lim_iters,
                   "audio_seconds_per_launch": float(wav_d.numel()) / cfg.sample_rate,
                   "griffin_lim_x_realtime": float(wav_d.numel()) / cfg.sample_rate / (gms * 1e-3), "in_timed_region": False}
        del post_d, spec_d, wav_d

    # GST front end on BASELINE configs[3] (batch 512 x 1000-frame reference mels; 16-token bank of Hyper_Parameters.json and the
    # 10-token bank configs[3] names): reported next to the headline (the timed step uses 188-frame reference mels)
    gst_blk = None
    if rank == 0 and not args.no_extras:
        GB, GT, GST_FLOP = 512, 1000, 201.4e6   # 201.4 MFLOP per 1000-frame mel
        gm = torch.as_tensor(np.random.default_rng(5).uniform(-4, 4, (GB, GT, cfg.mel_dim)).astype(np.float32), device=dev)
        gl = torch.full((GB,), GT, dtype=torch.int32, device=dev)
        gst_blk = {"workload": "configs[3]: batch {} x {}-frame reference mels".format(GB, GT), "in_timed_region": False}
        for tokens in (cfg.n_tokens, 10):
            e2 = eng if tokens == cfg.n_tokens else Engine(load_config(precision=precision, n_tokens=10),
                                                             init_weights(load_config(precision=precision, n_tokens=10), bias_scale=0.05), device=local_rank)
            for _ in range(3):
                e2.gst(gm, gl)
            torch.cuda.synchronize(dev)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(10):
                e2.gst(gm, gl)
            g1.record()
            torch.cuda.synchronize(dev)
            gms = g0.elapsed_time(g1) / 10
            gtf = GB * GST_FLOP / (gms * 1e-3) / 1e12
            gst_blk["tokens_{}".format(tokens)] = {"ms": gms, "achieved_tflops": gtf, "frac_of_tensor_peak": gtf / load_peaks()["bf16_tflops"],
                                                  "roofline_us": GB * GST_FLOP / (load_peaks()["bf16_tflops"] * 1e12) * 1e6,
                                                  "mels_per_s": GB / (gms * 1e-3)}
            if e2 is not eng:
                e2.close()
        del gm

    # Early stop (Model.py:380 cuts every utterance at argmax(stop < 0) AFTER Max_Step steps).  Random-init weights
    # give no meaningful stop token, so the stop distribution is synthetic and stated: the stop logit is not fed back, hence shifting
    # its bias moves the stop frames without touching the trajectory - the shift is chosen so that the last of the 256 utterances
    # stops at 60 % of Max_Step.
    es_blk = None
    if rank == 0 and not args.no_extras and precision == "bf16":
        full = step_device(300)
        st = full["stop"].float()
        cut = int(0.6 * T)
        shift = -float(st[:, :cut].min(dim=1).values.max().item()) - 1e-3
        W_es = dict(W)
        pb = np.array(W_es["Decoder/Decoder_Step/Projection/bias"], np.float32, copy=True)
        pb[-1] += shift
        W_es["Decoder/Decoder_Step/Projection/bias"] = pb
        e3 = Engine(cfg, W_es, device=local_rank)
        g = e3.gst(mels_d, lens_d, want=("gst",), host_outputs=False)["gst"]
        kw = dict(enc_text=text_d, gst=g, steps=T, rng="philox", seed=1300, row_offset=rank * B_DEC, host_outputs=False)
        ms_es, ms_full, out_es = [], [], None
        for i in range(4):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            out_es = e3.decode(early_stop=True, **kw)
            torch.cuda.synchronize(dev)
            ms_es.append((time.perf_counter() - t0) * 1e3)
            t0 = time.perf_counter()
            e3.decode(**kw)
            torch.cuda.synchronize(dev)
            ms_full.append((time.perf_counter() - t0) * 1e3)
        idx = out_es["stop_index"]
        es_blk = {"stop_distribution": "synthetic: stop-bias shift {:+.4f} puts the LAST utterance's stop frame at {} of {} steps (random-init stop "
                                       "logits hover around zero, so most rows cross it in the first steps: min/median = {}/{}); the loop leaves "
                                       "when the last utterance has stopped".format(shift, int(idx.max()), T, int(idx.min()), int(np.median(idx))),
                  "steps_done": int(out_es["steps_done"]), "ms_early_stop": float(np.median(ms_es[1:])), "ms_full_length": float(np.median(ms_full[1:])),
                  "speedup": float(np.median(ms_full[1:]) / np.median(ms_es[1:])),
                  "frames_per_s_until_exit": float(B_DEC * int(out_es["steps_done"]) / (np.median(ms_es[1:]) * 1e-3)), "in_timed_region": False}
        e3.close()

    lat = None
    if rank == 0 and not args.no_latency:
        # p50 per-step latency at batch 1 (BASELINE configs[0] shape: 80 tokens + <S>,<E>)
        e1 = torch.as_tensor(np.random.default_rng(1).uniform(-1, 1, (1, 82, cfg.text_dim)).astype(np.float32), device=dev)
        g1 = torch.zeros(1, cfg.style_size, device=dev)
        per = []
        for i in range(23):
            eng.decode(enc_text=e1, gst=g1, steps=T, rng="philox", seed=i, want=("mel", "stop"), host_outputs=False)
            if i >= 3:
                per.append(eng.last_kernel_ms() * 1e3 / T)
        lat = {"batch": 1, "key_time": 82, "p50_us_per_step": float(np.median(per)), "runs": len(per),
               "decoder_steps": T, "kernel": "decoder_bf16_sb (batch <= 16 dispatch)" if precision == "bf16" else "decoder_fp32"}
        if precision == "bf16":   # the same shape on the batch-256 kernel, and batch 8 x 150 keys on the latency kernel
            os.environ["GSTK_DECODER"] = "barrier"
            pb = []
            for i in range(8):
                eng.decode(enc_text=e1, gst=g1, steps=T, rng="philox", seed=i, want=("mel", "stop"), host_outputs=False)
                pb.append(eng.last_kernel_ms() * 1e3 / T)
            os.environ.pop("GSTK_DECODER")
            lat["p50_us_per_step_batch256_kernel"] = float(np.median(pb[2:]))
            e8 = torch.as_tensor(np.random.default_rng(2).uniform(-1, 1, (8, TV, cfg.text_dim)).astype(np.float32), device=dev)
            g8 = torch.zeros(8, cfg.style_size, device=dev)
            p8 = []
            for i in range(8):
                eng.decode(enc_text=e8, gst=g8, steps=T, rng="philox", seed=i, want=("mel", "stop"), host_outputs=False)
                p8.append(eng.last_kernel_ms() * 1e3 / T)
            lat["batch8_key_time150_p50_us_per_step"] = float(np.median(p8[2:]))
            e16 = torch.as_tensor(np.random.default_rng(3).uniform(-1, 1, (16, TV, cfg.text_dim)).astype(np.float32), device=dev)
            g16 = torch.zeros(16, cfg.style_size, device=dev)
            p16 = []
            for i in range(8):
                eng.decode(enc_text=e16, gst=g16, steps=T, rng="philox", seed=i, want=("mel", "stop"), host_outputs=False)
                p16.append(eng.last_kernel_ms() * 1e3 / T)
            lat["batch16_key_time150_p50_us_per_step"] = float(np.median(p16[2:]))

    if rank == 0:
        peaks = load_peaks()
        traffic = load_traffic()
        value = world * frames_per_step * args.steps / (total_ms * 1e-3)
        e2e_val = world * frames_per_step * args.steps / (e2e_ms * 1e-3)
        flops = FLOP_PER_STEP_UTT * B_DEC * T
        achieved = flops / (dec_ms * 1e-3) / 1e12
        res = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(cfg, args, precision),
            "e2e": {"value": e2e_val, "unit": UNIT,
                    "h2d_bytes_per_step": int(text_h.numel() * 4 + mels_h.numel() * 4 + lens_h.numel() * 4),
                    "d2h_bytes_per_step": int(sum(v.numel() for v in out_h.values()) * 4),
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "gstk_gst + gstk_decode with pinned host buffers" + (
                        " through runtime.EnginePool(depth=2): consecutive steps alternate between two engines" if e2e_depth > 1 else ""),
                    "pipeline_depth": e2e_depth,
                    "serial_value": world * frames_per_step * args.steps / (e2e_serial_ms * 1e-3),
                    "serial_ms_per_step": e2e_serial_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops"], "traffic": traffic[0] if precision == "bf16" else None,
                         "traffic_stale": traffic[1] if precision == "bf16" else None, "traffic_source": traffic[2],
                         "peak_source": peaks["source"],
                         "kernel": "persistent decoder ({})".format(precision),
                         "kernel_ms": dec_ms, "us_per_decoder_step": dec_ms * 1e3 / T,
                         "algorithmic_flops_per_launch": flops},
            "latency": lat,
            "postnet": post,
            "encoder": enc_blk,
            "gst": gst_blk,
            "vocoder": voc_blk,
            "early_stop": es_blk,
        }
        if post is not None:
            post["frac_of_tensor_peak"] = post["achieved_tflops"] / peaks["bf16_tflops"]
        if not args.no_cpu_baseline:
            res["cpu_baseline"] = cpu_baseline(cfg, W)
        print(json.dumps(res))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
