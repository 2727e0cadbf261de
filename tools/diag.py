"""Not a test: parity + timing diagnostic on the GPU box (ids -> mel -> PCM, golden fixtures, end-to-end error).

  python tools/diag.py [bench128|ljspeech_fixture28] [reps]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from larynx_b200 import engine, host, synth  # noqa: E402
from tests import common  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "bench128"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = np.load(os.path.join("tests", "golden", name + ".npz"))
eps = synth.seeded_eps(4096, seed=int(g["eps_seed"]))
glow, voc = common.product_glow(), common.product_hifi()
glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
call = engine.CALLS.get(0)
ids = g["ids"]
ns, ls = float(g["noise_scale"]), float(g["length_scale"])

mel = glow.phonemes_to_mels(ids, {"noise_scale": ns, "length_scale": ls, "eps": eps})
logw, w_ceil, idx = glow.model.alignment(call, 1, len(ids), mel.shape[2])
print("w_ceil exact", np.array_equal(w_ceil[0], g["w_ceil"]), "idx exact", np.array_equal(idx[0], g["idx"]),
      "logw", float(np.abs(logw[0] - g["logw"]).max()))
print("mel   max|d| %.3e (absmax %.2f)" % (float(np.abs(mel[0] - g["mel"]).max()), float(np.abs(g["mel"]).max())))
glued = glue.apply(mel)
print("glued max|d| %.3e" % float(np.abs(glued[0] - g["mel_glued"]).max()))
pcm_own = voc.mels_to_float(glued)
pcm_ref_mel = voc.mels_to_float(g["mel_glued"][None])
ref = g["pcm"]
sub = 1 if len(ref) == len(pcm_own) else 16  # bench128 keeps every 16th sample of the reference waveform
pcm_own, pcm_ref_mel = pcm_own[::sub], pcm_ref_mel[::sub]
print("pcm (reference mel) max|d| %.3e" % float(np.abs(pcm_ref_mel - ref).max()))
print("pcm (ids -> pcm, two calls) max|d| %.3e   rms %.3e   pcm absmax %.3f" % (
    float(np.abs(pcm_own - ref).max()), float(np.sqrt(np.mean((pcm_own - ref) ** 2))), float(np.abs(ref).max())))
f = host.synthesize_batch(glow, voc, glue, [ids], ns, ls, eps=eps[None, :, : max(mel.shape[2], 1)], want_float=True, want_mel=True)
print("pcm (ids -> pcm, fused)     max|d| %.3e" % float(np.abs(f["pcm"][0][::sub] - ref).max()))
print("mel fused vs two-call       max|d| %.3e" % float(np.abs(f["mel"][0] - mel[0]).max()))

for r in range(reps):
    t0 = time.perf_counter()
    y = glow.model.encode(call, np.ascontiguousarray(ids.reshape(1, -1)), np.array([len(ids)], np.int32), 1, len(ids), ls)
    enc_ms = call.last_ms
    l0 = call.launches
    m = np.zeros((1, 80, int(y[0])), np.float32)
    glow.model.decode(call, m, int(y[0]), ns, None, 0, 1)
    dec_ms, dec_l = call.last_ms, call.launches - l0
    l0 = call.launches
    a = voc.mels_to_audio(glued)
    voc_ms, voc_l = call.last_ms, call.launches - l0
    print("rep %d: encode %.3f ms  decode %.3f ms (%d launches)  vocoder %.3f ms (%d launches)  wall %.2f ms" % (
        r, enc_ms, dec_ms, dec_l, voc_ms, voc_l, (time.perf_counter() - t0) * 1e3))
