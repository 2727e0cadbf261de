"""Not a test, not the contract bench: whole-GPU throughput of the hot path under the reference's serving shapes.

  python tools/throughput.py                 # sweeps host threads x micro-batch size on cuda:0

The reference submits one `_sentence_task` per sentence to a ThreadPoolExecutor (larynx/__init__.py:146-157); here every
host thread owns one b200tts_call (stream + workspaces) and pulls micro-batches of sentences from a shared list, so the
kernels of different sentences overlap on the device.  Reports utterances/s, mel-frames/s and x real time (wall clock,
host buffers in and out — the same span as the reference's real-time factor).
"""
from __future__ import annotations

import argparse
import json
import sys
import threading
import time

import numpy as np

sys.path.insert(0, ".")
from larynx_b200 import dispatch, host, synth  # noqa: E402
from tests import common  # noqa: E402


def run(glow, voc, glue, rows, n_threads: int, batch: int, denoise: float, bucket: bool = False):
    if bucket:  # length-bucketed micro-batches, longest first (larynx_b200.dispatch.bucket_by_length)
        groups = dispatch.bucket_by_length([len(r) for r in rows], batch)[::-1]
        chunks = [[rows[i] for i in g] for g in groups]
    else:
        chunks = [rows[i : i + batch] for i in range(0, len(rows), batch)]
    nxt = [0]
    lock = threading.Lock()
    frames = [0] * n_threads
    err = []

    def worker(tid):
        try:
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= len(chunks):
                    return
                out = host.synthesize_batch(glow, voc, glue, chunks[i], seed=i, denoiser_strength=denoise)
                frames[tid] += int(out["y_len"].sum())
        except Exception as e:  # noqa: BLE001
            err.append(e)

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    if err:
        raise err[0]
    return dt, sum(frames)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256, help="sentences per measurement")
    ap.add_argument("--ids", default="128", help="'128' for fixed length or 'lo:hi' for U{lo..hi} (seed 1234)")
    ap.add_argument("--threads", default="1,2,4,8")
    ap.add_argument("--batch", default="1,4,16,32")
    ap.add_argument("--denoise", type=float, default=0.0)
    ap.add_argument("--bucket", action="store_true", help="sort sentences into micro-batches of similar length")
    a = ap.parse_args()
    glow, voc = common.product_glow(), common.product_hifi()
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    rs = np.random.RandomState(1234)
    if ":" in a.ids:
        lo, hi = (int(x) for x in a.ids.split(":"))
        lens = rs.randint(lo, hi + 1, a.n)
    else:
        lens = np.full(a.n, int(a.ids))
    rows = [synth.bench_ids(int(n), seed=1000 + i) for i, n in enumerate(lens)]
    for b in (int(x) for x in a.batch.split(",")):
        for t in (int(x) for x in a.threads.split(",")):
            if b * t > 64:  # every call context owns its workspaces (~0.5 GB per batched 6 s sentence): bound the footprint
                continue
            run(glow, voc, glue, rows[: max(2 * b * t, 8)], t, b, a.denoise, a.bucket)  # warm-up: arenas, attributes
            dt, fr = run(glow, voc, glue, rows, t, b, a.denoise, a.bucket)
            print(json.dumps({"threads": t, "batch": b, "ids": a.ids, "denoise": a.denoise, "bucket": a.bucket, "sentences": a.n, "wall_s": round(dt, 4),
                              "utt_per_s": round(a.n / dt, 1), "mel_frames_per_s": round(fr / dt, 1),
                              "x_realtime": round(fr * 256 / 22050 / dt, 1)}), flush=True)


if __name__ == "__main__":
    main()
