"""Not a test: a short workload for ncu (vocoder-only or full path) on synthetic inputs, no oracle."""
import sys

import numpy as np

sys.path.insert(0, ".")
from larynx_b200 import host, synth  # noqa: E402
from larynx_b200.engine import CALLS  # noqa: E402
from tests import common  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "voc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 502
call = CALLS.get(0)
if what == "voc":
    voc = common.product_hifi()
    mel = (np.random.RandomState(0).randn(1, 80, frames) * 4 - 4).astype(np.float32)
    for _ in range(reps):
        voc.mels_to_audio(mel)
        print("voc ms", call.last_ms)
else:
    glow, voc = common.product_glow(), common.product_hifi()
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    ids = synth.bench_ids(128)
    for _ in range(reps):
        r = host.synthesize_batch(glow, voc, glue, [ids], seed=1)
        print("finish ms", call.last_ms)
