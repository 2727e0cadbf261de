"""Not a test: a ragged batch with long items (two-block attention, per-layer encoder path) and the denoiser, for compute-sanitizer."""
import sys

import numpy as np

sys.path.insert(0, ".")
from larynx_b200 import host, synth  # noqa: E402
from tests import common  # noqa: E402

glow, voc = common.product_glow(), common.product_hifi()
glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
rows = [synth.bench_ids(n, seed=300 + n) for n in (150, 40, 200)]
o = host.synthesize_batch(glow, voc, glue, rows, seed=5, denoiser_strength=0.005)
print("ok", [int(a.shape[0]) for a in o["audio"]], [int(np.abs(a.astype(np.int32)).max()) for a in o["audio"]])
