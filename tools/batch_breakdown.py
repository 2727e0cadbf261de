"""Not a test: where a ragged micro-batch spends its device time (BASELINE config 3 shape: batch 32, T_x ~ U{64..256}).

  python tools/batch_breakdown.py [batch=32] [lo=64] [hi=256] [reps=4]

Prints encode / decode / vocoder device ms per batch (CUDA events on the call's stream, device-resident buffers) and per
utterance, plus the frames/s of each phase.
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from larynx_b200 import engine, host, synth  # noqa: E402
from tests import common  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 256
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
rs = np.random.RandomState(1234)
rows = sorted((synth.bench_ids(int(n), seed=3000 + i) for i, n in enumerate(rs.randint(lo, hi + 1, B))), key=len)
glow, voc = common.product_glow(), common.product_hifi()
glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
g, v = glow.model, voc.model
call = engine.CALLS.get(0)
dev = torch.device("cuda:0")
ids, x_len = engine.pad_ids(rows)
ids_d, xl_d = torch.from_numpy(ids).to(dev), torch.from_numpy(x_len).to(dev)
for r in range(reps):
    y = g.encode(call, ids_d, xl_d, B, ids.shape[1], 1.0)
    enc = call.last_ms
    ty = int(y.max())
    mel_d = torch.empty(B, 80, ty, device=dev)
    g.decode(call, mel_d, ty, 0.667, None, 0, 7)
    dec = call.last_ms
    glued = torch.from_numpy(glue.apply(mel_d.cpu().numpy())).to(dev)
    frames_d = torch.from_numpy(y.astype(np.int32)).to(dev)
    pcm_d = torch.empty(B, ty * 256, dtype=torch.int16, device=dev)
    v.infer(call, glued, frames_d, B, ty, None, pcm_d, ty * 256)
    vo = call.last_ms
    fr = int(y.sum())
    print("rep %d: batch %d, %d phonemes (max %d), %d frames (max %d): encode %.2f ms  decode %.2f ms  vocoder %.2f ms  | per utterance "
          "%.3f / %.3f / %.3f ms | vocoder %.0f kframes/s, total %.0f utt/s" % (
              r, B, int(x_len.sum()), ids.shape[1], fr, ty, enc, dec, vo, enc / B, dec / B, vo / B, fr / vo, B / (enc + dec + vo) * 1e3))
