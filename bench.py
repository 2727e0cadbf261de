#!/usr/bin/env python
"""bench.py — the reference's headline metric on the B200-native path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input: the BASELINE.json
configs[1] workload — en-us/ljspeech GlowTTS + hifi_gan/universal_large architecture (synthetic
weights, there is no network for checkpoints), ONE 128-phoneme sentence at batch 1, fp32:
phoneme ids -> mel (~502 frames, 5.8 s of audio) -> 22.05 kHz PCM.

Metric: mel-frames/s (whole job, all GPUs); the same line carries RTF and x-real-time.
  value : device-resident — ids/eps/outputs live in HBM, CUDA events on the library's stream.
  e2e   : the reference-facing drop-in (GlowTextToSpeech.phonemes_to_mels -> numpy glue ->
          HiFiGanVocoder.mels_to_audio, i.e. larynx._sentence_task's span, larynx/__init__.py:229-268)
          with HOST numpy buffers; wall clock; host<->device copies inside the timed region.
N > 1 (torchrun, one rank per GPU): utterances shard embarrassingly, each rank serves its own
replica (weights packed once on rank 0 and broadcast over NCCL at load); "weak" scaling.
--impl reference: the oracle port (a restatement of the reference's PyTorch-CPU backend; the
reference itself is pure Python + onnxruntime and neither it nor onnxruntime exists on the GPU box)
timed on the host cores on the same workload.
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
NOISE_SCALE, LENGTH_SCALE = 0.667, 1.0
HIFIGAN_MFLOP_PER_FRAME = 614.105088  # HiFi-GAN universal_large, exact 2*MAC count
FLOW_MFLOP_PER_FRAME = 21.351936      # flow decoder
ENC_MFLOP_PER_PHONEME = 2 * (7.163392 + 2304e-6 * N_IDS)
FP32_SIMT_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e-3  # 148 SMs x 128 FMA lanes x 2 flop x 1.965 GHz = 74.5


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload():
    from larynx_b200 import synth

    gcfg, hcfg = synth.glow_config(46), synth.hifigan_config("universal_large")
    return gcfg, hcfg, synth.bench_ids(N_IDS), synth


def cpu_port_runner():
    """Returns (fn() -> (frames, samples), cores, torch) running the oracle port once per call."""
    import torch

    from oracle import port

    gcfg, hcfg, ids, synth = make_workload()
    gsd = port.to_torch_sd(synth.glow_state_dict(gcfg, 1234))
    hsd = port.to_torch_sd(synth.hifigan_state_dict(hcfg, 4321))
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
    for _ in range(min(args.warmup, 2)):
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
        "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "rtf": rtf, "x_realtime": 1.0 / rtf,
        "config": workload_config(1),
        "cpu_baseline": {"value": val, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} x the same 128-phoneme utterance (oracle/port.py, torch CPU fp32, {cores} threads)"},
        "e2e": {"value": val, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of the reference's PyTorch-CPU backend; onnxruntime + generator.onnx are not available offline",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "configs[1]: ljspeech GlowTTS + hifi_gan/universal_large, one 128-phoneme sentence, batch 1, fp32",
            "n_ids": N_IDS, "batch": 1, "noise_scale": NOISE_SCALE, "length_scale": LENGTH_SCALE,
            "vocoder": "universal_large", "parallelism": f"replicas x{n_gpus} (utterance sharding, no hot-path collective)",
            "l2": "L2 flushed between timed steps (256 MiB write), untimed"}


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

    # ---------------- device-resident step (value) ----------------
    ids_d = torch.from_numpy(ids).view(1, -1).to(dev)
    xlen_d = torch.tensor([N_IDS], dtype=torch.int32, device=dev)
    cap = 4096
    pcm_d = torch.zeros(1, cap * v.hop, dtype=torch.int16, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    state = {"frames": 0}

    def step_device(seed):
        y = g.encode(call, ids_d, xlen_d, 1, N_IDS, LENGTH_SCALE)
        ms = call.last_ms
        engine.tts_finish(g, v, call, cglue, NOISE_SCALE, None, 0, seed, None, 0, None, pcm_d, pcm_d.shape[1])
        state["frames"] = int(y[0])
        return ms + call.last_ms

    def step_host(seed):
        r = host.sentence_task(ids_host, glue, glow, {"noise_scale": NOISE_SCALE, "length_scale": LENGTH_SCALE, "seed": seed}, voc)
        return r

    ids_host = torch.from_numpy(ids.copy()).pin_memory().numpy()

    for i in range(max(args.warmup, 3)):
        step_device(1000 + i)
        step_host(2000 + i)
    torch.cuda.synchronize()
    frames = state["frames"]
    samples = frames * v.hop

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- value: K steps, L2 flushed (untimed) between steps, device time from CUDA events on the call's stream
    barrier()
    l0 = call.launches
    dev_ms = []
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        dev_ms.append(step_device(i))
    barrier()
    wall_dev = time.perf_counter() - t_wall0
    launches = call.launches - l0
    # ---- vocoder-only pass for the roofline of the dominant kernel family (conv1d_kernel)
    mel_d = torch.zeros(1, 80, frames, device=dev)
    g.encode(call, ids_d, xlen_d, 1, N_IDS, LENGTH_SCALE)
    g.decode(call, mel_d, frames, NOISE_SCALE, None, 0, 7)
    glued_d = torch.from_numpy(glue.apply(mel_d.cpu().numpy())).to(dev)
    frames_d = torch.tensor([frames], dtype=torch.int32, device=dev)
    voc_ms = []
    lv0 = call.launches
    for i in range(max(3, min(args.steps, 10))):
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        v.infer(call, glued_d, frames_d, 1, frames, None, pcm_d, pcm_d.shape[1])
        voc_ms.append(call.last_ms)
    voc_launches = (call.launches - lv0) // len(voc_ms)
    # ---- e2e: host buffers through the drop-in boundary, wall clock
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        r = step_host(i)
    barrier()
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    assert r.audio.shape[0] == samples

    sum_dev_ms = float(np.sum(dev_ms))
    times = torch.tensor([sum_dev_ms, wall_e2e * 1e3, float(frames)], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = times.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        fsum = times[2:3].clone()
        dist.all_reduce(fsum, op=dist.ReduceOp.SUM)
        sum_dev_ms, wall_e2e_ms, tot_frames = float(tmax[0]), float(tmax[1]), float(fsum[0])
    else:
        wall_e2e_ms, tot_frames = wall_e2e * 1e3, float(frames)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        K = args.steps
        value = tot_frames * K / (sum_dev_ms * 1e-3)
        ms_per_step = sum_dev_ms / K
        e2e_val = tot_frames * K / (wall_e2e_ms * 1e-3)
        audio_s = samples / 22050.0
        rtf = (ms_per_step * 1e-3) / audio_s
        rtf_e2e = (wall_e2e_ms / K * 1e-3) / audio_s
        vms = float(np.mean(voc_ms))
        voc_tflops = HIFIGAN_MFLOP_PER_FRAME * 1e6 * frames / (vms * 1e-3) / 1e12
        tensor_peak = peaks.get("bf16_tflops", 1590.0)
        step_flops = (HIFIGAN_MFLOP_PER_FRAME + FLOW_MFLOP_PER_FRAME) * 1e6 * frames + ENC_MFLOP_PER_PHONEME * 1e6 * N_IDS
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
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "rtf": rtf, "x_realtime": 1.0 / rtf, "mel_frames_per_step": frames, "audio_seconds_per_step": audio_s,
            "wall_ms_per_step_device_resident": wall_dev / K * 1e3,
            "e2e": {"value": e2e_val, "unit": "mel-frames/s", "ms_per_step": wall_e2e_ms / K, "rtf": rtf_e2e,
                    "x_realtime": 1.0 / rtf_e2e,
                    "h2d_bytes_per_step": int(ids.nbytes + 4 + 80 * frames * 4 + 4),
                    "d2h_bytes_per_step": int(4 + 80 * frames * 4 + samples * 2),
                    "path": "GlowTextToSpeech.phonemes_to_mels -> numpy glue -> HiFiGanVocoder.mels_to_audio (host buffers)"},
            "gpu_launches": int(launches),
            "gpu_launches_per_step": int(launches // K),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": voc_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
                         "frac": voc_tflops / tensor_peak, "traffic": 6.1e7,
                         "traffic_note": "dram__bytes_read+write of one stage-2 ResBlock launch (3 MRF branches, C=128, L=32128; the largest share of the "
                                         "vocoder) from the ncu --set full capture in profiles/r1h_ncu_full_stage2_resblock.csv: 61 MB (112 MB on the first, "
                                         "cold pass) against 49 MB of activation images it reads + 49 MB it writes: they largely stay in the 126 MB L2",
                         "peak_source": peak_src + " bf16 dense (burst)",
                         "kernel": f"conv_umma_kernel<MT,NSPLIT,EW,WST,AST,G,NBUF> (tcgen05 kind::f16, bf16 hi/lo split = 3 products per algorithmic MAC "
                                   f"issued as 2 wide-N MMAs, fp32 TMEM accumulators, persistent double-buffered launches for the long stages): HiFi-GAN pass ({voc_launches} launches, {vms:.3f} ms, {HIFIGAN_MFLOP_PER_FRAME:.1f} MFLOP/frame x {frames} frames)",
                         "tensor_flops_issued_tflops": 3 * voc_tflops, "frac_tensor_issued": 3 * voc_tflops / tensor_peak,
                         "fp32_simt_peak_tflops": FP32_SIMT_PEAK_TFLOPS, "frac_of_fp32_simt_peak": voc_tflops / FP32_SIMT_PEAK_TFLOPS,
                         "step_tflops": step_flops / (ms_per_step * 1e-3) / 1e12,
                         "hbm_compulsory_gbs": (170.0e6 + 8 * N_IDS + 832 * frames) / (ms_per_step * 1e-3) / 1e9,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "cpu_baseline": {"value": cf / cpu_dt, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                             "rtf": cpu_dt / n_cpu / audio_s,
                             "sample": f"{n_cpu} x the same 128-phoneme utterance (oracle/port.py: torch CPU fp32 restatement of the reference PyTorch backend)"},
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
               str(args.steps), "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    bench_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
