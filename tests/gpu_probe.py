"""Not a test: first-contact diagnostic run on the GPU box (prints parity numbers per stage)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from larynx_b200 import synth  # noqa: E402
from larynx_b200.engine import CALLS  # noqa: E402
from oracle import port  # noqa: E402
from tests import common  # noqa: E402

n_ids = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ids = synth.bench_ids(n_ids)
eps = synth.seeded_eps(4096)
gcfg, hcfg = common.glow_cfg(), common.hifi_cfg()
t = time.time()
o = port.sentence_task(common.oracle_glow_sd(), gcfg, common.oracle_hifi_sd(), hcfg, ids, 0.667, 1.0, eps)
print("oracle %.2fs  T_y=%d" % (time.time() - t, o["mel"].shape[2]), flush=True)

glow = common.product_glow()
call = CALLS.get(0)
call.set_debug(True)
mel = glow.phonemes_to_mels(ids, {"noise_scale": 0.667, "length_scale": 1.0, "eps": eps})
print("mel", mel.shape, "enc+dec last_ms", call.last_ms, "launches", call.launches)
ty = mel.shape[2]
logw, w_ceil, idx = glow.model.alignment(call, 1, len(ids), ty)
print("logw maxdiff", np.abs(logw[0] - o["logw"].reshape(-1)).max())
print("w_ceil equal", np.array_equal(w_ceil[0], o["w_ceil"][0].astype(np.int32)), "idx equal", np.array_equal(idx[0], o["idx"][0]))
if ty == o["mel"].shape[2]:
    taps = {}
    ids_t = torch.from_numpy(ids).unsqueeze(0)
    g = port.glow_forward(common.oracle_glow_sd(), gcfg, ids_t, torch.tensor([len(ids)]), 0.667, 1.0, torch.from_numpy(eps).unsqueeze(0), taps)
    for blk in (11, 10, 6, 0):
        tap = glow.model.flow_tap(call, blk, 1, ty // 2)
        ref = taps[f"flow{blk}"].numpy()
        print("flow tap", blk, "maxdiff", np.abs(tap - ref).max(), "ref absmax", np.abs(ref).max())
    print("mel maxdiff", np.abs(mel - o["mel"]).max())
voc = common.product_hifi()
for rep in range(3):
    pcm = voc.mels_to_float(o["mel_glued"])
    print("hifigan last_ms", call.last_ms)
print("pcm maxdiff", np.abs(pcm - o["pcm"]).max(), "pcm absmax", np.abs(o["pcm"]).max())
aud = voc.mels_to_audio(o["mel_glued"])
print("int16 maxdiff", np.abs(aud.astype(np.int32) - o["audio"].astype(np.int32)).max(), aud.shape, aud.dtype)
for rep in range(3):
    t = time.perf_counter()
    mel = glow.phonemes_to_mels(ids, {"noise_scale": 0.667, "eps": eps})
    t1 = time.perf_counter()
    a = voc.mels_to_audio(port.mel_glue(mel, gcfg["audio"]))
    t2 = time.perf_counter()
    print("wall: glow %.2f ms, voc %.2f ms, audio %.2fs" % ((t1 - t) * 1e3, (t2 - t1) * 1e3, len(a) / 22050))
print("launches total", call.launches)
