#!/usr/bin/env python
"""bench.py — the reference's headline metric on the B200-native path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload = BASELINE.json configs[1]: en-us/ljspeech GlowTTS + hifi_gan/universal_large architecture (synthetic
weights, there is no network for checkpoints), ONE 128-phoneme sentence at batch 1, fp32: phoneme ids -> mel (~502
frames, 5.8 s of audio) -> 22.05 kHz PCM.  One "step" = UTT_PER_STEP of those utterances one after the other, each at
batch 1 (a single utterance takes ~3 ms: twenty of them would be a 60 ms timed region; 20 x 24 make ~1.4 s), L2 flushed (untimed) before
every utterance.

Metric: mel-frames/s (whole job, all GPUs); the same line carries RTF and x-real-time.
  value : device-resident — ids/outputs live in HBM, CUDA events on the library's stream around every utterance.
  e2e   : the reference-facing drop-in (GlowTextToSpeech.phonemes_to_mels -> numpy glue ->
          HiFiGanVocoder.mels_to_audio, i.e. larynx._sentence_task's span, larynx/__init__.py:229-268)
          with HOST numpy buffers; wall clock; host<->device copies inside the timed region.
  extra : BASELINE configs 3, 4, 5 (serving shapes through the micro-batch scheduler and the multi-GPU dispatcher), wall
          clock with host buffers; they do not enter `value`.
N > 1 (torchrun, one rank per GPU): utterances shard embarrassingly, each rank serves its own replica (weights packed
once on rank 0 and broadcast over NCCL at load); "weak" scaling.
--impl reference: the oracle port (a restatement of the reference's PyTorch-CPU backend; the reference itself is pure
Python + onnxruntime and neither it nor onnxruntime exists on the GPU box) timed on the host cores on the same sentence.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_IDS = 128
UTT_PER_STEP = 24   # 20 steps x 24 utterances x ~3 ms: a timed region of ~1.4 s
NOISE_SCALE, LENGTH_SCALE = 0.667, 1.0
HIFIGAN_MFLOP_PER_FRAME = 614.105088  # HiFi-GAN universal_large, exact 2*MAC count
FLOW_MFLOP_PER_FRAME = 21.351936      # flow decoder
ENC_MFLOP_PER_PHONEME = 2 * (7.163392 + 2304e-6 * N_IDS)
FP32_SIMT_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e-3  # 148 SMs x 128 FMA lanes x 2 flop x 1.965 GHz = 74.5
VOICE_SYMBOLS = [46, 54, 42, 52, 41, 38, 58, 41]   # config 4: en/de/fr/ru/sv/it/nl/sw (local/*/*-glow_tts/config.json)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the config 3/4/5 serving measurements")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return None


class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions: NVML polled in-process every 10 ms from before the warm-up
    (a `nvidia-smi -lms` child needs longer to start than a short bench runs); samples are kept when a timed region is
    open (`mark(True)` .. `mark(False)`)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self.active = False
        self.stop_flag = False
        self.t = None
        self.err = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.idx
            if vis:
                try:
                    idx = int(vis.split(",")[self.idx])
                except Exception:  # noqa: BLE001
                    pass
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            if self.active:
                try:
                    mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:  # noqa: BLE001
                        bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.samples.append((mhz, bits))
                except Exception as e:  # noqa: BLE001
                    self.err = repr(e)
            time.sleep(0.01)

    def mark(self, on: bool):
        self.active = on

    def stop(self):
        self.stop_flag = True
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + str(self.err)], "samples": 0}
        self.t.join(timeout=1.0)
        sm = [s[0] for s in self.samples]
        reasons = set()
        for _, bits in self.samples:
            for bit, name in self.REASONS.items():
                if bits & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload():
    from larynx_b200 import synth

    gcfg, hcfg = synth.glow_config(46), synth.hifigan_config("universal_large")
    return gcfg, hcfg, synth.bench_ids(N_IDS), synth


def cpu_port_runner():
    """Returns (fn() -> (frames, samples), cores) running the oracle port once per call.  Weight-norm is folded ONCE here
    (the reference folds at load: WN.remove_weight_norm / Generator.remove_weight_norm), not inside the timed forward."""
    import torch

    from oracle import port

    gcfg, hcfg, ids, synth = make_workload()
    gsd = port.prefold(port.to_torch_sd(synth.glow_state_dict(gcfg, 1234)))
    hsd = port.prefold(port.to_torch_sd(synth.hifigan_state_dict(hcfg, 4321)))
    eps = synth.seeded_eps(1024)
    avail = os.cpu_count() or 1
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        pass
    try:  # cgroup v2 CPU quota, if any
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            avail = max(1, min(avail, int(float(q) / float(per) + 0.5)))
    except Exception:  # noqa: BLE001
        pass
    # "all the host threads it can use": more threads than physical cores makes torch's conv1d slower,
    # so pick the fastest thread count on a short probe utterance (a few seconds in total)
    probe_ids = synth.bench_ids(16, seed=5)
    best, best_t = 1, float("inf")
    cands = sorted({c for c in (4, 8, 16, 32, 64, avail) if c <= avail} | {min(avail, 1)})
    for c in cands:
        torch.set_num_threads(c)
        port.sentence_task(gsd, gcfg, hsd, hcfg, probe_ids, NOISE_SCALE, LENGTH_SCALE, eps)
        t0 = time.perf_counter()
        port.sentence_task(gsd, gcfg, hsd, hcfg, probe_ids, NOISE_SCALE, LENGTH_SCALE, eps)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
        if dt > 20.0:
            break
    cores = best
    torch.set_num_threads(cores)

    def run():
        o = port.sentence_task(gsd, gcfg, hsd, hcfg, ids, NOISE_SCALE, LENGTH_SCALE, eps)
        return o["mel"].shape[2], o["audio"].shape[0]

    return run, cores


def bench_reference(args, rank):
    if rank != 0:
        return
    run, cores = cpu_port_runner()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        run()
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    frames = samples = 0
    for _ in range(steps):
        f, s = run()
        frames += f
        samples += s
    dt = time.perf_counter() - t0
    val = frames / dt
    rtf = dt / (samples / 22050.0)
    line = {
        "impl": "reference", "metric": "mel_frames_per_s", "value": val, "unit": "mel-frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "rtf": rtf, "x_realtime": 1.0 / rtf,
        "config": workload_config(1),
        "cpu_baseline": {"value": val, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                         "sample": f"each step = 1 x the same 128-phoneme utterance (a bounded sample of the GPU arm's {UTT_PER_STEP}-utterance step; "
                                   f"oracle/port.py, torch CPU fp32, {cores} threads, weight-norm folded once at load)"},
        "e2e": {"value": val, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of the reference's PyTorch-CPU backend; onnxruntime + generator.onnx are not available offline",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "configs[1]: ljspeech GlowTTS + hifi_gan/universal_large, one 128-phoneme sentence, batch 1, fp32",
            "n_ids": N_IDS, "batch": 1, "utterances_per_step": UTT_PER_STEP, "noise_scale": NOISE_SCALE, "length_scale": LENGTH_SCALE,
            "vocoder": "universal_large", "parallelism": f"replicas x{n_gpus} (utterance sharding, no hot-path collective)",
            "l2": "L2 flushed before every timed utterance (256 MiB write), untimed"}


# ---------------------------------------------------------------------------------------------------------------------
# serving shapes (BASELINE configs 3, 4, 5): wall clock, host buffers in and out, through the scheduler / dispatcher
# ---------------------------------------------------------------------------------------------------------------------
def fixture_length_sampler(n, mean, seed):
    """Sentence lengths ~ the reference's fixture rows (111/28/47/51/61 ids, local/en-us/ljspeech-glow_tts/samples) jittered
    and scaled to the given mean."""
    rs = np.random.RandomState(seed)
    base = np.array([111, 28, 47, 51, 61], dtype=np.float64)
    lens = base[rs.randint(0, len(base), n)] * rs.uniform(0.6, 1.4, n)
    lens = lens * (mean / lens.mean())
    return np.clip(np.round(lens), 4, 400).astype(np.int64)


def run_extras(args, rank, world, local_rank, glow, voc, glue, hblob_factory, dist, torch):
    from larynx_b200 import dispatch, pack, synth
    from larynx_b200.glow_tts import GlowTextToSpeech
    from larynx_b200.scheduler import BatchScheduler

    dev = torch.device(f"cuda:{local_rank}")

    def reduce(wall, utts, frames):
        t = torch.tensor([wall, float(utts), float(frames)], dtype=torch.float64, device=dev)
        if dist is not None:
            w = t[0:1].clone()
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
            s = t[1:3].clone()
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
            return float(w[0]), float(s[0]), float(s[1])
        return wall, float(utts), float(frames)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(sched_list, rows_list, warm_all=False):
        """rows_list[i] goes through sched_list[i]; returns (wall, utterances, frames) of this rank."""
        for sch, rows in zip(sched_list, rows_list):  # warm-up: arenas, attributes
            # ragged lengths: every worker must have seen the longest bucket, or its workspace growth (cudaFree + cudaMalloc) is timed
            if warm_all:
                sch.reserve(rows)
            sch.synthesize(rows if warm_all else rows[: min(len(rows), 2 * sch.max_batch)])
        barrier()
        t0 = time.perf_counter()
        outs = []
        if len(sched_list) == 1:
            outs.append(sched_list[0].synthesize(rows_list[0]))
        else:  # several voices on this GPU: their streams run concurrently
            res = [None] * len(sched_list)

            def work(i):
                res[i] = sched_list[i].synthesize(rows_list[i])

            th = [threading.Thread(target=work, args=(i,)) for i in range(len(sched_list))]
            [t.start() for t in th]
            [t.join() for t in th]
            outs = res
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        n = sum(len(o) for o in outs)
        frames = sum(int(a.shape[0]) // 256 for o in outs for a in o)
        return wall, n, frames

    def all_ok(ok: bool) -> bool:
        """Every rank takes the same path: a failure anywhere ends the extras everywhere (no rank waits in a collective)."""
        if dist is None:
            return ok
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t[0]))

    def record(name, workload, fn):
        err, res = None, (1.0, 0, 0)
        try:
            res = fn()
        except Exception as e:  # noqa: BLE001
            err = repr(e)
        if not all_ok(err is None):
            extra[name] = {"workload": workload, "error": err or "failed on another rank"}
            return False
        w, n, f = reduce(*res)
        extra[name] = {"workload": workload, "utterances_per_s": n / w, "mel_frames_per_s": f / w, "x_realtime": f * 256 / 22050.0 / w,
                       "wall_s": w, "utterances": int(n)}
        return True

    extra = {}

    # ---- config 3: batch 32, ragged U{64..256} phonemes, one replica per GPU
    def config3():
        rs = np.random.RandomState(1234 + rank)
        rows = [synth.bench_ids(int(n), seed=3000 + 97 * rank + i) for i, n in enumerate(rs.randint(64, 257, 128))]
        sch = BatchScheduler(glow, voc, glue, max_batch=32, max_cost=None, workers=2)
        try:
            return timed([sch], [rows], warm_all=True)
        finally:
            sch.close()

    # ---- config 4: 8 voices (num_symbols differ), voice v pinned to rank v % world, shared vocoder
    def config4():
        mine = [v for v in range(len(VOICE_SYMBOLS)) if v % world == rank]
        scheds, rows = [], []
        try:
            for v in mine:
                gm = glow
                if v != 0:
                    vcfg = synth.glow_config(VOICE_SYMBOLS[v])
                    gm = GlowTextToSpeech(None, device=local_rank, blob=pack.pack_glow(synth.glow_state_dict(vcfg, 1234 + v), vcfg), model_config=vcfg)
                scheds.append(BatchScheduler(gm, voc, glue, max_batch=16, workers=1))
                rows.append([synth.bench_ids(N_IDS, seed=4000 + 131 * v + i, num_symbols=VOICE_SYMBOLS[v]) for i in range(64)])
            return timed(scheds, rows)
        finally:
            for s_ in scheds:
                s_.close()

    # ---- config 5: 1250 sentences per GPU (10 k on 8), sharded round-robin, denoiser 0.005, ordered gather
    total = 1250 * world
    state5 = {}

    def config5():
        lens = fixture_length_sampler(total, 100, 1234)
        shard = dispatch.shard_round_robin(total, rank, world)
        rows = [synth.bench_ids(int(lens[i]), seed=5000 + i) for i in shard]
        sch = BatchScheduler(glow, voc, glue, max_batch=32, max_cost=4096, workers=2, denoiser_strength=0.005)
        try:
            sch.synthesize(rows[:64])
            barrier()
            t0 = time.perf_counter()
            audio = sch.synthesize(rows)
            torch.cuda.synchronize()
            state5["local"] = {i: (int(a.shape[0]), int(np.abs(a.astype(np.int32)).sum() % 1000003)) for i, a in zip(shard, audio)}
            state5["t0"] = t0
            return time.perf_counter() - t0, len(rows), sum(int(a.shape[0]) for a in audio) / 256.0
        finally:
            sch.close()

    w3 = "batch 32, U{64..256} phonemes (padded+masked, length-bucketed), 128 utterances per GPU, fp32 encoder"
    w4 = (f"8 voices (num_symbols {VOICE_SYMBOLS}), voice v on GPU v % {world}, shared universal_large vocoder, 64 x 128-phoneme "
          "utterances per voice, micro-batches of 16")
    w5 = (f"{total} sentences (fixture length distribution scaled to mean 100 phonemes), round-robin over {world} GPU(s), universal_large "
          "(-q high), denoiser 0.005 on the device; + ordered gather of (length, checksum) per sentence over the process group")
    if record("config3", w3, config3) and record("config4", w4, config4) and record("config5", w5, config5):
        # the dispatcher's ordered gather (what a book reader needs to stitch chapters): every rank gets all 10 k entries
        t0 = time.perf_counter()
        merged = dispatch.gather_in_order(state5["local"], total)
        extra["config5"]["gather_s"] = time.perf_counter() - t0
        extra["config5"]["gathered_sentences"] = len(merged)
    return extra


def bench_ours(args, rank, world, local_rank):
    import torch

    from larynx_b200 import dispatch, engine, host, pack
    from larynx_b200.glow_tts import GlowTextToSpeech
    from larynx_b200.hifi_gan import HiFiGanVocoder

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    gcfg, hcfg, ids, synth = make_workload()
    # one collective, at load: packed weights broadcast from rank 0 (NCCL over NVLink)
    gblob = dispatch.broadcast_blob((lambda: pack.pack_glow(synth.glow_state_dict(gcfg, 1234), gcfg)) if rank == 0 else None, dev)
    hblob = dispatch.broadcast_blob((lambda: pack.pack_hifigan(synth.hifigan_state_dict(hcfg, 4321), hcfg)) if rank == 0 else None, dev)
    glow = GlowTextToSpeech(None, device=local_rank, blob=gblob, model_config=gcfg)
    voc = HiFiGanVocoder(None, device=local_rank, blob=hblob, model_config=hcfg)
    del gblob, hblob
    glue = host.MelGlue.from_config(gcfg["audio"])
    cglue = glue.c_struct()
    call = engine.CALLS.get(local_rank)
    g, v = glow.model, voc.model

    # ---------------- device-resident utterance (value) ----------------
    ids_d = torch.from_numpy(ids).view(1, -1).to(dev)
    xlen_d = torch.tensor([N_IDS], dtype=torch.int32, device=dev)
    cap = 4096
    pcm_d = torch.zeros(1, cap * v.hop, dtype=torch.int16, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    state = {"frames": 0}

    def utt_device(seed):
        y = g.encode(call, ids_d, xlen_d, 1, N_IDS, LENGTH_SCALE)
        ms = call.last_ms
        engine.tts_finish(g, v, call, cglue, NOISE_SCALE, None, 0, seed, None, 0, None, pcm_d, pcm_d.shape[1])
        state["frames"] = int(y[0])
        return ms + call.last_ms

    ids_host = torch.from_numpy(ids.copy()).pin_memory().numpy()

    def utt_host(seed):
        return host.sentence_task(ids_host, glue, glow, {"noise_scale": NOISE_SCALE, "length_scale": LENGTH_SCALE, "seed": seed}, voc)

    warm = max(args.warmup, 3)
    for i in range(warm):
        for u in range(2):
            utt_device(1000 + 2 * i + u)
            utt_host(2000 + 2 * i + u)
    torch.cuda.synchronize()
    frames = state["frames"]
    samples = frames * v.hop

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: K steps of UTT_PER_STEP utterances; L2 flushed (untimed) before each; device time from CUDA events on
    #      the call's stream around each utterance
    barrier()
    sampler.mark(True)
    l0 = call.launches
    dev_ms = 0.0
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        for u in range(UTT_PER_STEP):
            flush.fill_((i + u) & 0xFF)
            torch.cuda.synchronize()
            dev_ms += utt_device(i * UTT_PER_STEP + u)
    barrier()
    sampler.mark(False)
    wall_dev = time.perf_counter() - t_wall0
    launches = call.launches - l0
    n_utt = args.steps * UTT_PER_STEP
    # ---- vocoder-only pass for the roofline of the dominant kernel family
    mel_d = torch.zeros(1, 80, frames, device=dev)
    g.encode(call, ids_d, xlen_d, 1, N_IDS, LENGTH_SCALE)
    enc_ms = call.last_ms
    g.decode(call, mel_d, frames, NOISE_SCALE, None, 0, 7)
    dec_ms = call.last_ms
    glued_d = torch.from_numpy(glue.apply(mel_d.cpu().numpy())).to(dev)
    frames_d = torch.tensor([frames], dtype=torch.int32, device=dev)
    voc_ms = []
    lv0 = call.launches
    for i in range(10):
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        v.infer(call, glued_d, frames_d, 1, frames, None, pcm_d, pcm_d.shape[1])
        voc_ms.append(call.last_ms)
    voc_launches = (call.launches - lv0) // len(voc_ms)
    # ---- e2e: host buffers through the drop-in boundary, wall clock, back to back
    barrier()
    sampler.mark(True)
    t0 = time.perf_counter()
    for i in range(n_utt):
        r = utt_host(i)
    barrier()
    sampler.mark(False)
    wall_e2e = time.perf_counter() - t0
    assert r.audio.shape[0] == samples

    times = torch.tensor([dev_ms, wall_e2e * 1e3, float(frames)], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = times.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        fsum = times[2:3].clone()
        dist.all_reduce(fsum, op=dist.ReduceOp.SUM)
        sum_dev_ms, wall_e2e_ms, tot_frames = float(tmax[0]), float(tmax[1]), float(fsum[0])
    else:
        sum_dev_ms, wall_e2e_ms, tot_frames = dev_ms, wall_e2e * 1e3, float(frames)

    extra = None
    if not args.no_extra:
        try:
            extra = run_extras(args, rank, world, local_rank, glow, voc, glue, None, dist, torch)
        except Exception as e:  # noqa: BLE001 - the headline line must survive a failure of the extras
            extra = {"error": repr(e)}
            if dist is not None:
                raise
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks, peak_src = measured_peaks()
        K = args.steps
        value = tot_frames * n_utt / (sum_dev_ms * 1e-3)
        ms_per_utt = sum_dev_ms / n_utt
        e2e_val = tot_frames * n_utt / (wall_e2e_ms * 1e-3)
        audio_s = samples / 22050.0
        rtf = (ms_per_utt * 1e-3) / audio_s
        rtf_e2e = (wall_e2e_ms / n_utt * 1e-3) / audio_s
        vms = float(np.mean(voc_ms))
        voc_tflops = HIFIGAN_MFLOP_PER_FRAME * 1e6 * frames / (vms * 1e-3) / 1e12
        tensor_peak = peaks.get("bf16_tflops", 1590.0)
        step_flops = (HIFIGAN_MFLOP_PER_FRAME + FLOW_MFLOP_PER_FRAME) * 1e6 * frames + ENC_MFLOP_PER_PHONEME * 1e6 * N_IDS
        traffic = ncu_traffic()
        # cpu baseline: bounded sample (2 warm + ~6 utterances ~ 10-20 s) on this box's host cores
        if args.no_cpu:
            cores, n_cpu, cf, cpu_dt = 0, 1, 0, 1.0
        else:
            run, cores = cpu_port_runner()
            run()
            n_cpu, t0 = 0, time.perf_counter()
            cf = 0
            while n_cpu < 2 or (time.perf_counter() - t0 < 12.0 and n_cpu < 12):
                f, _ = run()
                cf += f
                n_cpu += 1
            cpu_dt = time.perf_counter() - t0
        line = {
            "metric": "mel_frames_per_s", "value": value, "unit": "mel-frames/s", "n_gpus": world, "steps": K,
            "warmup": warm, "ms_per_step": sum_dev_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "ms_per_utterance": ms_per_utt, "timed_region_s": sum_dev_ms * 1e-3,
            "rtf": rtf, "x_realtime": 1.0 / rtf, "mel_frames_per_utterance": frames, "audio_seconds_per_utterance": audio_s,
            "wall_ms_per_utterance_device_resident_incl_flush": wall_dev / n_utt * 1e3,
            "phase_ms": {"encode": enc_ms, "decode": dec_ms, "vocoder": vms},
            "e2e": {"value": e2e_val, "unit": "mel-frames/s", "ms_per_utterance": wall_e2e_ms / n_utt, "rtf": rtf_e2e,
                    "x_realtime": 1.0 / rtf_e2e,
                    "h2d_bytes_per_step": int(UTT_PER_STEP * (ids.nbytes + 4 + 80 * frames * 4 + 4)),
                    "d2h_bytes_per_step": int(UTT_PER_STEP * (4 + 80 * frames * 4 + samples * 2)),
                    "path": "GlowTextToSpeech.phonemes_to_mels -> numpy glue -> HiFiGanVocoder.mels_to_audio (host buffers), back to back"},
            "gpu_launches": int(launches),
            "gpu_launches_per_utterance": int(launches // n_utt),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": voc_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
                         "frac": voc_tflops / tensor_peak,
                         "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                         "traffic_note": (traffic.get("note") if traffic else "no ncu --set full capture of the current kernels committed yet"),
                         "traffic_unit": "bytes per stage_umma_kernel launch (dram__bytes_read.sum + dram__bytes_write.sum, mean of the 4 stage launches)",
                         "peak_source": peak_src + " bf16 dense (burst)",
                         "kernel": f"stage_umma_kernel<MT> x 4 (one persistent dataflow launch per HiFi-GAN stage: tcgen05 kind::f16, bf16 hi/lo split = 3 products per "
                                   f"algorithmic MAC issued as 2 wide-N MMAs, fp32 TMEM accumulators double buffered) + conv_umma_kernel (conv_pre, upsamplers): HiFi-GAN pass "
                                   f"({voc_launches} launches, {vms:.3f} ms, {HIFIGAN_MFLOP_PER_FRAME:.1f} MFLOP/frame x {frames} frames)",
                         "tensor_flops_issued_tflops": 3 * voc_tflops, "frac_tensor_issued": 3 * voc_tflops / tensor_peak,
                         "fp32_simt_peak_tflops": FP32_SIMT_PEAK_TFLOPS, "frac_of_fp32_simt_peak": voc_tflops / FP32_SIMT_PEAK_TFLOPS,
                         "step_tflops": step_flops / (ms_per_utt * 1e-3) / 1e12,
                         "hbm_compulsory_gbs": (170.0e6 + 8 * N_IDS + 832 * frames) / (ms_per_utt * 1e-3) / 1e9,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "cpu_baseline": {"value": cf / cpu_dt, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                             "rtf": cpu_dt / n_cpu / audio_s,
                             "sample": f"{n_cpu} x the same 128-phoneme utterance (oracle/port.py: torch CPU fp32 restatement of the reference PyTorch backend, weight-norm folded once)"},
            "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        bench_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: re-exec ourselves with one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29533"), os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps",
               str(args.steps), "--warmup", str(args.warmup)] + (["--no-extra"] if args.no_extra else []) + (["--no-cpu"] if args.no_cpu else [])
        sys.exit(subprocess.call(cmd))
    bench_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
