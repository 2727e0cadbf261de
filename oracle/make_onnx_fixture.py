"""TEST INFRASTRUCTURE — writes tests/golden/{glow_tiny,hifigan_tiny}.onnx: the reference's OWN modules (glow_tts/models.py
FlowGenerator after store_inverse(), hifi_gan/models.py Generator after remove_weight_norm()) traced by torch's TorchScript ONNX
exporter, at a tiny configuration so that the files stay small.  larynx_b200/onnx_weights.py is tested against them
(tests/test_pack_and_abi.py): real exporter output — scoped constant names, Constant nodes, folded weight norm — not a file
written by the test itself.

Runs only in the build container (needs /root/reference).  The `onnx` package is not installed; torch's exporter needs it only
for a post-processing step that is a no-op for these graphs (`_add_onnxscript_fn`), which is bypassed here.

  python -m oracle.make_onnx_fixture
"""
import json
import os
import sys
import tempfile
import warnings
from pathlib import Path

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larynx_b200 import synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

GLOW_TINY = dict(hidden_channels=16, hidden_channels_enc=16, hidden_channels_dec=24, filter_channels=32, filter_channels_dp=40, n_layers_enc=2,
                 n_blocks_dec=3, n_block_layers=2)
GLOW_SEED, HIFI_SEED = 77, 78


def glow_tiny_config() -> dict:
    return synth.glow_config(46, **GLOW_TINY)


def hifigan_tiny_config() -> dict:
    cfg = synth.hifigan_config("vctk_small")
    cfg["model"]["upsample_initial_channel"] = 32
    return cfg


def main():
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils

    onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes
    warnings.filterwarnings("ignore")
    out = Path(__file__).resolve().parent.parent / "tests" / "golden"
    tmp = Path(tempfile.mkdtemp())
    gdir = synth.write_glow_voice(tmp / "g", seed=GLOW_SEED, config=glow_tiny_config())
    hdir = tmp / "h"
    hdir.mkdir()
    hcfg = hifigan_tiny_config()
    with open(hdir / "config.json", "w", encoding="utf-8") as f:
        json.dump(hcfg, f)
    torch.save({"generator": {k: torch.from_numpy(v) for k, v in synth.hifigan_state_dict(hcfg, HIFI_SEED).items()}}, str(hdir / "generator.pth"))
    glow, gen = ref_harness.load_reference_models(gdir, hdir)
    torch.onnx.export(gen, (torch.randn(1, 80, 20),), str(out / "hifigan_tiny.onnx"), opset_version=12, do_constant_folding=True,
                      input_names=["mel"], output_names=["audio"], dynamic_axes={"mel": {0: "b", 2: "t"}, "audio": {0: "b", 2: "t"}}, dynamo=False)
    for p in glow.parameters():
        p.requires_grad_(False)
    for f in glow.decoder.flows:
        if hasattr(f, "weight_inv"):
            f.weight_inv = f.weight_inv.detach()

    class Infer(torch.nn.Module):  # the three inputs the reference feeds its GlowTTS graph (larynx/glow_tts.py:161-168)
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, x_lengths, scales):
            (mel, *_), _, _ = self.m(x, x_lengths, noise_scale=scales[0], length_scale=scales[1])
            return mel

    with torch.no_grad():
        torch.onnx.export(Infer(glow), (torch.randint(1, 40, (1, 12)), torch.tensor([12]), torch.tensor([0.667, 1.0])), str(out / "glow_tiny.onnx"),
                          opset_version=12, do_constant_folding=True, input_names=["input", "input_lengths", "scales"], output_names=["output"],
                          dynamic_axes={"input": {0: "b", 1: "t"}, "input_lengths": {0: "b"}, "output": {0: "b", 2: "t"}}, dynamo=False)
    for n in ("glow_tiny.onnx", "hifigan_tiny.onnx"):
        print(n, os.path.getsize(out / n), "bytes")


if __name__ == "__main__":
    main()
