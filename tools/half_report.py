"""Not a test: what the reduced-precision encoder (`encoder_bf16`, b200tts_call_set_encoder_bf16) costs and buys.

BASELINE config 3 asks for "bf16 encoder GEMMs (fp32 accumulate, fp32 LN/softmax/duration head) ... report mel parity and
alignment mismatch rate vs fp32".  The fp32 side here is this library's default mode, whose w_ceil / idx equal the
reference modules' bit for bit on every golden fixture (tests/test_parity_gpu.py), so a mismatch against it is a mismatch
against the reference.

  python tools/half_report.py [n_utterances=416] [out.json]

Utterances: ids ~ U{1..45}, T_x ~ U{64..128} (seed 1234) — the single-launch encoder's range; longer batches take the
per-layer path, which has no reduced-precision mode.  Batches of 32, same injected eps in both modes.
"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from larynx_b200 import engine, synth  # noqa: E402
from tests import common  # noqa: E402

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 416
out_path = sys.argv[2] if len(sys.argv) > 2 else None
B = 32
rng = np.random.RandomState(1234)
glow = common.product_glow()
g = glow.model
call = engine.CALLS.get(0)
eps_all = synth.seeded_eps(2048, seed=1234)  # [80, 2048]

ph_total = ph_bad = utt_bad = ylen_bad = 0
dlogw = 0.0
mel_err, mel_n = 0.0, 0
ms = {False: [], True: []}
for b0 in range(0, n_utt, B):
    rows = [rng.randint(1, 46, size=rng.randint(64, 129)).astype(np.int64) for _ in range(min(B, n_utt - b0))]
    ids, x_len = engine.pad_ids(rows)
    res = {}
    for mode in (False, True):
        y_len = g.encode(call, ids, x_len, len(rows), ids.shape[1], 1.0, encoder_bf16=mode)
        ms[mode].append(call.last_ms)
        ty = int(y_len.max())
        mel = np.zeros((len(rows), g.mel_channels, ty), np.float32)
        eps = np.ascontiguousarray(np.broadcast_to(eps_all[None, :, :ty], mel.shape))
        g.decode(call, mel, ty, 0.667, eps, ty, 0)
        logw, w_ceil, _ = g.alignment(call, len(rows), ids.shape[1], ty)  # (the taps exist after decode)
        res[mode] = (y_len.copy(), logw, w_ceil, mel)
    (y0, l0, w0, m0), (y1, l1, w1, m1) = res[False], res[True]
    for i, r in enumerate(rows):
        n = len(r)
        bad = int((w0[i, :n] != w1[i, :n]).sum())
        ph_total += n
        ph_bad += bad
        utt_bad += bad > 0
        ylen_bad += int(y0[i] != y1[i])
        dlogw = max(dlogw, float(np.abs(l0[i, :n] - l1[i, :n]).max()))
        if bad == 0:  # same alignment: the mels are comparable frame by frame
            mel_err = max(mel_err, float(np.abs(m0[i, :, : y0[i]] - m1[i, :, : y0[i]]).max()))
            mel_n += 1

# batch-1 latency at the bench length
ids1 = rng.randint(1, 46, size=(1, 128)).astype(np.int64)
x1 = np.array([128], np.int32)
lat = {}
for mode in (False, True):
    t = []
    for _ in range(12):
        g.encode(call, ids1, x1, 1, 128, 1.0, encoder_bf16=mode)
        t.append(call.last_ms)
    lat[mode] = float(np.median(t[2:]))

rep = {
    "utterances": n_utt, "phonemes": ph_total,
    "w_ceil_mismatch_per_phoneme": ph_bad / max(ph_total, 1),
    "utterances_with_any_mismatch": utt_bad / max(n_utt, 1),
    "utterances_with_other_length": ylen_bad / max(n_utt, 1),
    "max_abs_dlogw": dlogw,
    "mel_max_abs_err_same_alignment": mel_err, "mel_compared_utterances": mel_n,
    "encode_ms_batch32_fp32": float(np.median(ms[False][1:] or ms[False])),
    "encode_ms_batch32_bf16": float(np.median(ms[True][1:] or ms[True])),
    "encode_ms_batch1_128ids_fp32": lat[False], "encode_ms_batch1_128ids_bf16": lat[True],
}
print(json.dumps(rep, indent=1))
if out_path:
    with open(out_path, "w") as f:
        json.dump(rep, f, indent=1)
