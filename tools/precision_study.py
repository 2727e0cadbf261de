"""Not a test: how many operand bits the HiFi-GAN pass needs to stay inside the 1e-3 waveform contract (CPU, oracle/port.py).

Every conv of the vocoder is rerun with its INPUT rounded to a storage format (weights likewise where stated), fp32 accumulation,
on the 516-frame mel of tests/golden/bench128.npz; the waveform is compared with the fp32 oracle's.  This bounds what a
cheaper operand scheme than the engine's bf16 hi/lo split (3 tensor-core products per MAC, ~16 mantissa bits per operand)
could deliver: one fp16 plane for the activations would be 2 products per MAC with split weights, 1 product with fp16 weights.

  python tools/precision_study.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import port  # noqa: E402
from tests import common  # noqa: E402

g = np.load(os.path.join("tests", "golden", "bench128.npz"))
mel = torch.from_numpy(g["mel_glued"][None])
sd, cfg = common.oracle_hifi_sd(), common.hifi_cfg()
ref = port.hifigan_forward(sd, cfg, mel).reshape(-1).numpy()


def rounder(kind):
    if kind == "fp32":
        return lambda t: t
    if kind == "fp16":
        return lambda t: t.half().float()
    if kind == "bf16":
        return lambda t: t.bfloat16().float()
    if kind == "bf16x2":  # hi + lo planes
        def f(t):
            hi = t.bfloat16().float()
            return hi + (t - hi).bfloat16().float()
        return f
    if kind == "fp16x2":
        def f(t):
            hi = t.half().float()
            return hi + (t - hi).half().float()
        return f
    raise ValueError(kind)


def run(act, wt):
    ra, rw = rounder(act), rounder(wt)
    c1, ct = F.conv1d, F.conv_transpose1d
    F.conv1d = lambda x, w, b=None, **k: c1(ra(x), rw(w), b, **k)
    F.conv_transpose1d = lambda x, w, b=None, **k: ct(ra(x), rw(w), b, **k)
    try:
        y = port.hifigan_forward(sd, cfg, mel).reshape(-1).numpy()
    finally:
        F.conv1d, F.conv_transpose1d = c1, ct
    d = np.abs(y - ref)
    return float(d.max()), float(np.sqrt(np.mean(d ** 2)))


print("waveform peak %.3f, %d samples" % (float(np.abs(ref).max()), ref.size))
for act, wt, what in (("bf16x2", "bf16x2", "engine's scheme (3 products: hh + hl + lh; lo x lo dropped is below this)"),
                      ("fp16x2", "fp16x2", "fp16 hi/lo for both"),
                      ("fp16", "fp16x2", "ONE fp16 activation plane x split weights (2 products)"),
                      ("fp16", "fp32", "one fp16 activation plane, exact weights"),
                      ("fp16", "fp16", "fp16 activations and weights (1 product; the reference's `half`, fp32 accumulate)"),
                      ("bf16x2", "fp16", "split activations x ONE fp16 weight plane (2 products)"),
                      ("bf16", "bf16x2", "one bf16 activation plane x split weights"),
                      ("bf16", "bf16", "plain bf16 (1 product)")):
    mx, rms = run(act, wt)
    print("act %-7s w %-7s max|d| %.3e  rms %.3e   %s" % (act, wt, mx, rms, what))
