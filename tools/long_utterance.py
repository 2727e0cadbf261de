"""Not a test: one 200-phoneme utterance through the drop-in (for an ncu launch list: which attention kernel runs)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from larynx_b200 import host, synth  # noqa: E402
from oracle import port  # noqa: E402
from tests import common  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ids = synth.bench_ids(n, seed=7)
eps = synth.seeded_eps(4096)
glow = common.product_glow()
mel = glow.phonemes_to_mels(ids, {"noise_scale": 0.667, "length_scale": 1.0, "eps": eps})
o = port.sentence_task(common.oracle_glow_sd(), common.glow_cfg(), common.oracle_hifi_sd(), common.hifi_cfg(), ids, 0.667, 1.0, eps)
print("frames", mel.shape[2], "oracle", o["mel"].shape[2], "mel max|d| %.3e" % float(np.abs(mel - o["mel"]).max()))
